#!/usr/bin/env python3
"""bench.py — encrypted candidate distances/sec of the PreFHEtch server-side search hot path.

One "step" = one batch of encrypted queries through the whole hot path on synthetic SIFT-shaped
data: stage 1 (plaintext coarse quantization, top-nprobe) + stage 2 (rotated query sets, ct x pt
multiply-accumulate over every candidate block of the probed lists, add of the norms, inverse NTT,
mod-switch of the results).
  value : whole-job useful candidate distances/s with query ciphertexts already resident in HBM
  e2e   : the same metric through the C-ABI calls with HOST buffers (SEAL-serialized query ciphertexts
          in pinned memory -> SEAL-serialized result ciphertexts in pinned memory), requests pipelined
          with pf_search_submit / pf_search_collect the way a serving loop keeps the GPU busy
  roofline : the ct x pt MAC kernel against the measured HBM copy bandwidth (+ the rotate phase against
          the FP64 / IMAD pipes in `rotate_roofline`)
  cpu_baseline : the CPU oracle (port of the same op sequence) on a bounded sample, all host cores
`--impl reference` times that CPU path alone (the reference's own server cannot be built here:
FAISS fork / SEAL / Drogon are network FetchContent dependencies).

Multi-GPU (torchrun, one process per GPU): ranks form a grid of L list shards x Q query groups
(`--grid LxQ`, default Nx1).  Rank r owns the IVF lists l with l % L == r % L and serves the queries of
group r // L; result ciphertexts are gathered into rank 0's HBM over NVLink (peer DMA + flags).  Default
at N > 1: WEAK scaling of configs[1] (every list shard holds its own 1M vectors) as the headline, plus a
`strong` record: BASELINE configs[2] (fixed 1M index) at N GPUs and at 1 GPU in the same job.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CONFIGS = {
    # BASELINE.json configs[1]: the single-GPU configuration the metric is quoted on
    "sift1m_nlist1024_nprobe16": dict(nb=1_000_000, d=128, nlist=1024, nprobe=16, n=8192, g=8, m=1, tbits=24, nq=64),
    # BASELINE.json configs[2]: lists sharded across 2/4/8 GPUs (strong scaling of a fixed 1M index)
    "sift1m_nlist4096_nprobe64": dict(nb=1_000_000, d=128, nlist=4096, nprobe=64, n=8192, g=32, m=1, tbits=24, nq=64),
    # BASELINE.json configs[0] shape (the reference's own CPU-runnable case)
    "siftsmall_nlist100_nprobe8": dict(nb=10_000, d=128, nlist=100, nprobe=8, n=8192, g=8, m=1, tbits=24, nq=16),
    # BASELINE.json configs[3]: GIST1M, 960-d padded to 1024, 8 query ciphertexts (dimension chunks of 128),
    # 15 rotations per chunk, K = 128 diagonals per block, 27-bit plain modulus
    "gist1m_nlist1024": dict(nb=1_000_000, d=960, nlist=1024, nprobe=16, n=8192, g=8, m=8, tbits=27, nq=16),
    # BASELINE.json configs[4]: 10M x 128, nlist 16384, batch of 256 encrypted queries; nprobe is not stated in
    # BASELINE.json: 64.  --poly-degree 16384 runs the other point of the sweep (L = 8, g = 16: same C).
    "synth10m_nlist16384": dict(nb=10_000_000, d=128, nlist=16384, nprobe=64, n=8192, g=8, m=1, tbits=24, nq=256),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------
# synthetic SIFT-shaped data: uint8-valued vectors from a Gaussian mixture
# ----------------------------------------------------------------------------------------------
def make_dataset(cfg, device, seed=1234, sigma=24.0, spread=160.0, lloyd=0, npool=4096, offset=0.0):
    """Index build (out of the timed path; the reference does it once in Server::init_index).
    torch is used only as plumbing for the k-means-style assignment.  sigma / spread / lloyd shape the
    mixture: the defaults are SURVEY §8d's well-separated clusters; a wide sigma with a few Lloyd
    iterations gives overlapping lists (recall@10 < 1 at the same nprobe)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    nb, d, nlist = cfg["nb"], cfg["d"], cfg["nlist"]
    centres = offset + torch.rand((nlist, d), generator=g, device=device) * spread
    assign = torch.randint(0, nlist, (nb,), generator=g, device=device)
    chunk = 1 << 20
    base = torch.empty((nb, d), device=device)
    for s in range(0, nb, chunk):   # chunked: 10M x 128 noise in one piece would double the footprint
        e = min(nb, s + chunk)
        base[s:e] = torch.clamp(torch.round(centres[assign[s:e]] + torch.randn((e - s, d), generator=g, device=device) * sigma), 0, 255)
    # IVF centroids = mean of the assigned vectors of each true cluster (one Lloyd step from the truth)
    cent = torch.zeros((nlist, d), device=device).index_add_(0, assign, base)
    cnt = torch.bincount(assign, minlength=nlist).clamp(min=1).unsqueeze(1)
    cent = cent / cnt

    def nearest(c):
        lab = torch.empty(nb, dtype=torch.long, device=device)
        c2 = (c * c).sum(1)
        for s in range(0, nb, 65536):
            x = base[s:s + 65536]
            lab[s:s + 65536] = (c2[None, :] - 2.0 * x @ c.T).argmin(1)
        return lab

    # assign every vector to its nearest centroid (what faiss::IndexIVF::add does)
    lab = nearest(cent)
    for _ in range(lloyd):
        cent2 = torch.zeros((nlist, d), device=device).index_add_(0, lab, base)
        c2 = torch.bincount(lab, minlength=nlist).unsqueeze(1)
        cent = torch.where(c2 > 0, cent2 / c2.clamp(min=1), cent)
        lab = nearest(cent)
    order = torch.argsort(lab, stable=True)
    counts = torch.bincount(lab, minlength=nlist)
    offsets = torch.zeros(nlist + 1, dtype=torch.long, device=device)
    offsets[1:] = torch.cumsum(counts, 0)
    vecs = base[order].contiguous()
    # query pool from the same mixture
    qa = torch.randint(0, nlist, (npool,), generator=g, device=device)
    queries = torch.clamp(torch.round(centres[qa] + torch.randn((npool, d), generator=g, device=device) * sigma), 0, 255)
    return dict(centroids=cent.cpu().numpy().astype(np.float32), offsets=offsets.cpu().numpy().astype(np.int64),
                ids=order.cpu().numpy().astype(np.int64, copy=False), vectors=vecs.cpu().numpy().astype(np.float32, copy=False),
                queries=queries.cpu().numpy().astype(np.float32, copy=False))


def make_dataset_shard(cfg, device, rank, world, seed=1234):
    """Weak-scaling data set: the index of configs[1] replicated `world` times — world*nb vectors,
    world*nlist lists, list l owned by rank l % world.  Every rank generates the centres of ALL lists
    (same seed) but the vectors of its own lists only; the lists of other ranks are empty here, which is
    how a sharded deployment loads its shard.  Vectors stay in the list of the centre that generated
    them; the centroid is the mean of the list (exchanged between ranks by the caller)."""
    import torch
    nb, d, nlist = cfg["nb"], cfg["d"], cfg["nlist"]
    G = nlist * world
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    centres = torch.rand((G, d), generator=g, device=device) * 160.0          # identical on every rank
    npool = 4096
    qa = torch.randint(0, G, (npool,), generator=g, device=device)
    queries = torch.clamp(torch.round(centres[qa] + torch.randn((npool, d), generator=g, device=device) * 24.0), 0, 255)
    g2 = torch.Generator(device=device)
    g2.manual_seed(seed + 7919 * (rank + 1))
    own = torch.arange(rank, G, world, device=device)                           # lists of this rank
    assign = own[torch.randint(0, nlist, (nb,), generator=g2, device=device)]
    base = torch.clamp(torch.round(centres[assign] + torch.randn((nb, d), generator=g2, device=device) * 24.0), 0, 255)
    order = torch.argsort(assign, stable=True)
    counts = torch.bincount(assign, minlength=G)
    offsets = torch.zeros(G + 1, dtype=torch.long, device=device)
    offsets[1:] = torch.cumsum(counts, 0)
    vecs = base[order].contiguous()
    cent = torch.zeros((G, d), device=device).index_add_(0, assign, base) / counts.clamp(min=1).unsqueeze(1)
    ids = order + rank * nb                                                      # globally unique ids
    return dict(centroids=cent, counts=counts, offsets=offsets.cpu().numpy().astype(np.int64),
                ids=ids.cpu().numpy().astype(np.int64), vectors=vecs.cpu().numpy().astype(np.float32),
                queries=queries.cpu().numpy().astype(np.float32))


def make_dataset_cpu_light(cfg, seed=1234):
    """CPU-only dataset for the reference arm when no GPU plumbing is wanted: same generator family, numpy."""
    rng = np.random.default_rng(seed)
    nb, d, nlist = cfg["nb"], cfg["d"], cfg["nlist"]
    centres = rng.uniform(0, 160, size=(nlist, d)).astype(np.float32)
    assign = rng.integers(0, nlist, size=nb)
    base = np.empty((nb, d), dtype=np.float32)
    for s in range(0, nb, 1 << 18):
        e = min(nb, s + (1 << 18))
        base[s:e] = np.clip(np.rint(centres[assign[s:e]] + rng.normal(0, 24, size=(e - s, d)).astype(np.float32)), 0, 255)
    cent = np.zeros((nlist, d), dtype=np.float64)
    np.add.at(cent, assign, base)
    cent = (cent / np.maximum(np.bincount(assign, minlength=nlist), 1)[:, None]).astype(np.float32)
    lab = np.empty(nb, dtype=np.int64)
    c2 = (cent * cent).sum(1)
    for s in range(0, nb, 65536):
        x = base[s:s + 65536]
        lab[s:s + 65536] = (c2[None, :] - 2.0 * x @ cent.T).argmin(1)
    order = np.argsort(lab, kind="stable")
    offsets = np.zeros(nlist + 1, dtype=np.int64)
    np.cumsum(np.bincount(lab, minlength=nlist), out=offsets[1:])
    qa = rng.integers(0, nlist, size=4096)
    queries = np.clip(np.rint(centres[qa] + rng.normal(0, 24, size=(4096, d))), 0, 255).astype(np.float32)
    return dict(centroids=cent, offsets=offsets, ids=order.astype(np.int64),
                vectors=np.ascontiguousarray(base[order]), queries=queries)


# ----------------------------------------------------------------------------------------------
# clocks / throttle reasons DURING the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled while the GPU is under load: NVML in-process every 5 ms when
    libnvidia-ml is loadable, else an `nvidia-smi -lms 20` child.  Started BEFORE warm-up (process start-up
    and the first NVML call take longer than a 50 ms timed region); every sample carries its host time and
    `stop(t0, t1)` reports the samples inside the timed window (falling back to all samples under load if
    the window caught none, and saying so)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index, torch_device=None):
        self.gpu, self.dev = gpu_index, torch_device
        self.rows = []            # (host time, sm MHz, set of reasons)
        self.max_mhz, self.proc, self.source = None, None, None
        self.stop_flag = threading.Event()

    # -- NVML ---------------------------------------------------------------------------------
    def _nvml_handle(self, nv):
        import torch
        idx = torch.device(self.dev).index or 0 if self.dev is not None else self.gpu
        cands = []
        try:
            cands.append("GPU-" + str(torch.cuda.get_device_properties(self.dev).uuid))
        except Exception:
            pass
        for u in cands:
            for arg in (u, u.encode()):
                try:
                    return nv.nvmlDeviceGetHandleByUUID(arg)
                except Exception:
                    continue
        return nv.nvmlDeviceGetHandleByIndex(idx)

    def _nvml_poll(self):
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        while not self.stop_flag.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.rows.append((time.perf_counter(), mhz, {k for k, b in bits.items() if r & b}))
            except Exception:
                pass
            self.stop_flag.wait(0.005)

    def _start_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        self.nv = pynvml
        self.h = self._nvml_handle(pynvml)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        float(pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM))   # fails here rather than in the thread
        self.source = "nvml"
        self.th = threading.Thread(target=self._nvml_poll, daemon=True)
        self.th.start()

    # -- nvidia-smi ---------------------------------------------------------------------------
    def _smi_read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if len(r) >= 8 and r[1].replace(".", "").isdigit():
                if r[2].replace(".", "").isdigit():
                    self.max_mhz = max(self.max_mhz or 0.0, float(r[2]))
                self.rows.append((time.perf_counter(), float(r[1]), {self.NAMES[i] for i in range(4) if r[4 + i] == "Active"}))

    def _start_smi(self):
        self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
        self.source = "nvidia-smi"
        self.th = threading.Thread(target=self._smi_read, daemon=True)
        self.th.start()

    def start(self):
        mode = os.environ.get("PF_BENCH_CLOCKS", "auto")
        if mode == "off":
            return
        if mode in ("auto", "nvml"):
            try:
                self._start_nvml()
                return
            except Exception as ex:
                log(f"[clocks] NVML unavailable ({type(ex).__name__}: {ex}); using nvidia-smi")
        try:
            self._start_smi()
        except OSError:
            self.proc, self.source = None, None

    def stop(self, t0=None, t1=None):
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if getattr(self, "th", None):
            self.th.join(timeout=2)
        return self.report(t0, t1)

    def report(self, t0=None, t1=None):
        """clocks over the samples inside [t0, t1] (host perf_counter); the sampler keeps running"""
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source available"], "samples": 0}
        rows = list(self.rows)
        win = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
        window = "timed"
        if not win:   # a very short timed region between two samples: everything sampled under load
            win, window = rows, "load (warm-up + timed + e2e)"
        sm = [r[1] for r in win]
        reasons = sorted(set().union(*[r[2] for r in win])) if win else []
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "samples_total": len(rows), "window": window, "source": self.source}


# ----------------------------------------------------------------------------------------------
# host placement: a rank's threads and its pinned buffers live on the NUMA node of its GPU
# ----------------------------------------------------------------------------------------------
def pin_to_gpu_numa_node(torch_device):
    """sched_setaffinity to the CPUs local to this GPU's PCIe root (sysfs local_cpulist).  Pinned buffers
    allocated afterwards are first-touched, hence placed, on that node: at 8 ranks the D2H copies of ranks
    whose buffers sat on the other socket ran at 60 % of the local ones."""
    try:
        import torch
        p = torch.cuda.get_device_properties(torch_device)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        path = Path("/sys/bus/pci/devices") / bdf / "local_cpulist"
        txt = path.read_text().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
        return {"pci": bdf, "cpus": len(use or allowed), "pinned": bool(use and use != allowed)}
    except Exception as ex:   # placement is an optimisation, never a requirement
        return {"pinned": False, "why": f"{type(ex).__name__}: {ex}"}


def bind_pages_to_gpu_node(arr, torch_device):
    """mbind(2) the pages of a host buffer to the NUMA node of this rank's GPU BEFORE they are first touched
    (the shared response buffer: a rank whose share sat on the other socket downloaded at 60 % of the local
    rate).  Best effort: a container whose cpuset.mems excludes the node, or a kernel without NUMA, says no."""
    try:
        import ctypes
        import torch
        p = torch.cuda.get_device_properties(torch_device)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int((Path("/sys/bus/pci/devices") / bdf / "numa_node").read_text().strip())
        if node < 0:
            return {"bound": False, "why": "no NUMA node reported for the GPU"}
        addr = arr.ctypes.data
        page = 4096
        start = (addr + page - 1) // page * page
        length = (addr + arr.nbytes) // page * page - start
        if length <= 0:
            return {"bound": False, "why": "buffer smaller than a page"}
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        libc = ctypes.CDLL(None, use_errno=True)
        # MPOL_PREFERRED, not MPOL_BIND: if the node cannot supply the pages the kernel falls back to another node
        # instead of failing the first touch
        MPOL_PREFERRED, SYS_mbind = 1, 237
        rc = libc.syscall(SYS_mbind, ctypes.c_void_p(start), ctypes.c_ulong(length), ctypes.c_int(MPOL_PREFERRED), mask,
                          ctypes.c_ulong(16 * 64), ctypes.c_uint(0))
        if rc != 0:
            return {"bound": False, "node": node, "why": f"mbind errno {ctypes.get_errno()}"}
        return {"bound": True, "node": node}
    except Exception as ex:
        return {"bound": False, "why": f"{type(ex).__name__}: {ex}"}


# ----------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle's OpenMP whole-step driver on a bounded sample
# ----------------------------------------------------------------------------------------------
def cpu_sample_queries(nq, nthreads):
    """bounded sample of the workload both CPU legs time: two queries per host thread, at least 8"""
    return int(min(nq, max(8, 2 * nthreads)))


def cpu_pipeline(cfg, data, nq_sample, nthreads, steps=1, warmup=0, seed=5, result_limbs=0):
    """Runs the CPU port of the step on `nq_sample` queries of the workload.  Only this function, the parity
    self-check and --impl reference execute oracle/ (as baseline / checker, never as the product)."""
    from oracle import pf_oracle as O
    n, d, g, m, nprobe = cfg["n"], cfg["d"], cfg["g"], cfg["m"], cfg["nprobe"]
    primes, t = O.BFV_DEFAULT_PRIMES[n], O.BATCHING_T[(n, cfg["tbits"])]
    ctx = O.Context(n, primes, t)
    lay = O.LayoutPlan(n, d, m, g)
    rng = np.random.default_rng(seed)
    q = data["queries"][:nq_sample]
    idx, _ = O.coarse_quantize(q, data["centroids"], nprobe)
    offsets = data["offsets"]
    # blocks touched by the sample
    block_of = {}
    boff, bnv, pair_q, pair_b, useful = [], [], [], [], 0
    for i in range(nq_sample):
        for l in idx[i]:
            n_l = int(offsets[l + 1] - offsets[l])
            for b0 in range(0, n_l, lay.C):
                key = (int(l), b0)
                if key not in block_of:
                    block_of[key] = len(boff)
                    boff.append(int(offsets[l]) + b0)
                    bnv.append(min(lay.C, n_l - b0))
                pair_q.append(i)
                pair_b.append(block_of[key])
                useful += min(lay.C, n_l - b0)
    xs = data["vectors"].astype(np.int32)
    t0 = time.perf_counter()
    diag, norm = O.encode_blocks(ctx, lay, xs, boff, bnv, nthreads)
    enc_s = time.perf_counter() - t0
    # ciphertext-shaped random residues (throughput does not depend on the values) and random keys
    L, k = ctx.L, ctx.k
    cts = np.stack([rng.integers(0, primes[l], size=(nq_sample, m, 2, n), dtype=np.uint64) for l in range(L)], axis=3)
    cts = np.ascontiguousarray(cts)
    keys = []
    for r in range(1, lay.R):
        keys.append(np.ascontiguousarray(
            np.stack([rng.integers(0, primes[j], size=(L, 2, n), dtype=np.uint64) for j in range(k)], axis=2)))
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        idx2, _ = O.coarse_quantize(q, data["centroids"], nprobe)
        out, (rot_s, mac_s) = O.search_pairs(ctx, lay, cts, keys, False, pair_q, pair_b, diag, norm, nthreads,
                                             result_limbs=result_limbs)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append((dt, rot_s, mac_s))
    dt = float(np.mean([x[0] for x in times]))
    return dict(useful=useful, slots=len(pair_q) * lay.C, seconds=dt, rot_s=float(np.mean([x[1] for x in times])),
                mac_s=float(np.mean([x[2] for x in times])), encode_s=enc_s, pairs=len(pair_q), nq=nq_sample,
                distinct_blocks=len(boff))


def reference_plaintext_path(data, nprobe, nq, repeats=3):
    """What the reference server/client pair computes TODAY for BASELINE configs[0] (no HE in the snapshot):
    client-side coarse quantization (ref: src/client/client_lib.cpp:50-81) + every vector of the probed lists
    scored with the exact squared L2 of Server::preciseSearch (ref: src/server/server_lib.cpp:111-167), single
    thread like the reference's one Drogon IO thread; the oracle's restatement of those loops, timed."""
    from oracle import pf_oracle as O
    q = np.ascontiguousarray(data["queries"][:nq], dtype=np.float32)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        idx, _ = O.coarse_quantize(q, data["centroids"], nprobe)
        t1 = time.perf_counter()
        dist, labels, sizes = O.search_lists_plain(q, idx, data["offsets"], data["ids"], data["vectors"])
        t2 = time.perf_counter()
        if best is None or t2 - t0 < best[0]:
            best = (t2 - t0, t1 - t0, t2 - t1, int(sizes.sum()))
    return {"what": "reference plaintext path (sort_nearest_centroids + exact L2 over the probed lists), 1 thread, CPU port",
            "queries": int(nq), "candidates": best[3], "seconds": best[0], "coarse_s": best[1], "lists_s": best[2],
            "plaintext_distances_per_s": best[3] / best[0], "queries_per_s": nq / best[0]}


def reference_functions_timed(data):
    """The reference's OWN plaintext functions — sort_nearest_centroids (src/client/client_lib.cpp:49-81) and
    Server::preciseSearch (src/server/server_lib.cpp:140-167), compiled from its sources into oracle/_ref (recipe:
    oracle/ref_build/Makefile; prebuilt .so on the GPU box) — timed on this host, one thread, in the reference's
    compile-time shape (5 queries, 200 candidates).  kind "reference": no restatement involved."""
    try:
        from oracle import pf_ref as R
        if not R.build():
            return {"unavailable": "oracle/_ref not built and the reference sources are not on this machine"}
        d = data["vectors"].shape[1]
        if d != R.D:
            return {"unavailable": f"the reference is compiled for d = {R.D}"}
        q = np.ascontiguousarray(data["queries"][:R.NQUERY], dtype=np.float32)
        ids = (np.arange(R.NQUERY * R.COARSE_PROBE, dtype=np.int64) * 7919 % len(data["vectors"])).reshape(R.NQUERY, R.COARSE_PROBE)
        t = R.time_reference_functions(q, data["centroids"], ids, data["vectors"])
        t.update({"kind": "reference", "cores": 1,
                  "centroid_distances_per_s": R.NQUERY * len(data["centroids"]) / t["sort_nearest_centroids_s"],
                  "exact_distances_per_s": R.NQUERY * R.COARSE_PROBE / t["precise_search_s"]})
        return t
    except Exception as ex:    # noqa: BLE001 — a record, not a gate
        return {"unavailable": repr(ex)[:300]}


def bench_config(cfg_name, cfg, L, result_limbs, nq, world=1, weak=False, nprobe=None, grid=None, db_gib=None):
    """the `config` object of the JSON line — built by ONE function for both arms so that they name the same
    workload key for key (the CPU arm's bounded sample is described under cpu_baseline.sample)"""
    nprobe = nprobe if nprobe is not None else cfg["nprobe"]
    c = {"workload": cfg_name + (f" x{world} shards (weak: {world}M vectors, nlist {cfg['nlist'] * world}, nprobe {nprobe})" if weak else ""),
         "nb": cfg["nb"] * (world if weak else 1), "d": cfg["d"], "nlist": cfg["nlist"] * (world if weak else 1),
         "nprobe": nprobe, "poly_degree": cfg["n"], "limbs": L, "result_limbs": result_limbs or L, "g": cfg["g"],
         "query_cts": cfg["m"], "queries_per_step": nq,
         "parallelism": (f"grid {grid[0]} list shards x {grid[1]} query groups" if world > 1 else "single"),
         "l2_policy": "inputs larger than L2: the NTT-domain DB (GiBs per rank) is streamed from HBM, query batches rotate through a pool"}
    return c


def run_reference(args, cfg, cfg_name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import pf_oracle as O
    O.build()
    data = make_dataset(cfg, "cpu") if cfg["nb"] <= 200_000 else make_dataset_cpu_light(cfg)
    nthreads = O.host_cores()      # all host cores, whatever OMP_NUM_THREADS torchrun exported
    L = len(O.BFV_DEFAULT_PRIMES[cfg["n"]]) - 1
    rl = args.result_limbs if 0 < args.result_limbs < L else 0
    nq_sample = cpu_sample_queries(cfg["nq"], nthreads)
    # N > 1 without --config: the GPU arm weak-scales this workload (N list shards of 1M vectors, nprobe 16 * N).  The
    # CPU arm takes the same per-query work — nprobe * N probed lists — from ONE shard's lists (the cost of a block does
    # not depend on which vectors are in it), so both arms name and do the same workload.
    world = max(1, int(os.environ.get("WORLD_SIZE", args.gpus or 1)))
    weak = world > 1 and args.config is None
    run_cfg = dict(cfg)
    if weak:
        run_cfg["nprobe"] = min(cfg["nprobe"] * world, cfg["nlist"])
    r = cpu_pipeline(run_cfg, data, nq_sample, nthreads, steps=args.steps, warmup=args.warmup, result_limbs=rl)
    val = r["useful"] / r["seconds"]
    line = {
        "metric": "encrypted candidate distances/sec", "value": val, "unit": "distances/s", "impl": "reference",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": bench_config(cfg_name, cfg, L, rl, cfg["nq"], world, weak, run_cfg["nprobe"], (world, 1)),
        "cpu_baseline": {"value": val, "unit": "distances/s", "cores": nthreads, "kind": "port",
                         "sample": f"{nq_sample} of the {cfg['nq']} queries of a step x {r['pairs']} (query,block) pairs"
                                   + (f" ({run_cfg['nprobe']} probed lists per query, drawn from one 1M-vector shard)" if weak else "") + ", whole hot path "
                                   f"incl. mod-switch to {rl or L} limb(s) (rotations {r['rot_s']:.2f}s + MAC/INTT {r['mac_s']:.2f}s)",
                         "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
        "e2e": {"value": val, "unit": "distances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "queries_per_s": nq_sample / r["seconds"],
    }
    if cfg["nb"] <= 200_000:   # BASELINE configs[0]: also what the reference computes today, in plaintext
        line["reference_plaintext_path"] = reference_plaintext_path(data, cfg["nprobe"], min(100, len(data["queries"])))
        line["reference_functions"] = reference_functions_timed(data)
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------
# recall (untimed bookkeeping)
# ----------------------------------------------------------------------------------------------
def recall_metrics(eng, data, nprobe, dev, nq_r=64, gt_k=100):
    """Both recall@10 definitions of the two-stage search against exact brute force over the whole base set
    (untimed, world 1).  Stage 2 here is the engine's plaintext path: the encrypted results decrypt to exactly
    these distances (tests/test_gpu_parity.py), so both pipelines have this recall.
      standard : |GT_top10 ∩ returned_top10| / 10                     (BASELINE.json's metric)
      reference: |GT_top100 ∩ returned_top10| / 10                    (ref: src/client/client_lib.cpp:272-281,:327)
    Ties are broken by the lower id on both sides."""
    import torch
    x = np.ascontiguousarray(data["queries"][:nq_r], dtype=np.float32)
    idx = eng.coarse_quantize(x, nprobe)
    dist, labels, sizes = eng.coarseSearch(x, idx)
    base = torch.from_numpy(data["vectors"]).to(dev)
    ids = torch.from_numpy(data["ids"]).to(dev)
    xq = torch.from_numpy(x).to(dev).double()
    q2 = (xq * xq).sum(1)
    SH = 1 << 24
    gt_k = min(gt_k, base.shape[0])
    best = None
    for s0 in range(0, base.shape[0], 131072):      # exact integer distances in float64, chunked
        b = base[s0:s0 + 131072].double()
        d2 = q2[:, None] + (b * b).sum(1)[None, :] - 2.0 * xq @ b.T
        key = d2.round().long() * SH + ids[s0:s0 + 131072][None, :]
        key = key if best is None else torch.cat([best, key], dim=1)
        best = key.topk(min(gt_k, key.shape[1]), dim=1, largest=False).values
    gt = (best % SH).cpu().numpy()
    hits10 = hits100 = 0
    off = 0
    for i in range(len(x)):
        n = int(sizes[i])
        k = dist[off:off + n].astype(np.int64) * SH + labels[off:off + n]
        found = set((np.sort(k)[:10] % SH).tolist())
        hits10 += len(found & set(gt[i][:10].tolist()))
        hits100 += len(found & set(gt[i].tolist()))
        off += n
    return {"recall_at_10": hits10 / (10.0 * len(x)), "reference_recall_10": hits100 / (10.0 * len(x)), "queries": len(x)}


def recall_at_10(eng, data, nprobe, dev, nq_r=64):
    return recall_metrics(eng, data, nprobe, dev, nq_r)["recall_at_10"]


# ----------------------------------------------------------------------------------------------
_JSON_OUT = None


def emit(line):
    """the one JSON line goes to the process's original stdout"""
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)


class StageGuard:
    """No optional stage may cost the run its headline.  After the timed region rank 0 publishes the JSON line as
    it stands; every later stage (e2e, parity self-check, CPU baseline, the strong-scaling record) runs between
    enter(stage, limit) and leave().  If a stage hangs (a peer died inside a collective) or a rank reports a
    failure through the abort file, the watchdog thread of every rank fires: rank 0 emits the published line
    with `aborted_stage` filled in, and all ranks leave with exit code 0.  The same thread enforces a limit on
    the whole run."""

    def __init__(self, rank, world, total_limit_s=1500.0):
        self.rank, self.world = rank, world
        # the driver gives one bench run 870 s (SCALE_r01.json: per_n_timeout_s) and then kills it: whatever the stages
        # do, the line that exists by then is emitted BEFORE that (PF_BENCH_RUN_LIMIT_S, counted from here)
        total_limit_s = min(total_limit_s, float(os.environ.get("PF_BENCH_RUN_LIMIT_S", 780.0)))
        self.run_deadline = time.monotonic() + total_limit_s
        self.partial, self.stage, self.deadline = None, "setup + timed region", self.run_deadline
        self.abort_file = Path(f"/tmp/pf_bench_abort_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}")
        self.shm_names = []
        self.done = threading.Event()
        try:
            self.abort_file.unlink()
        except OSError:
            pass
        # torchrun ends the surviving workers with SIGTERM when one of them dies (e.g. killed for memory).  A Python
        # signal handler only runs when the MAIN thread returns to the interpreter — not while it waits inside a
        # collective — so the signal number is also written to a wake-up pipe that the watchdog thread polls.
        self.sig_r = None
        try:
            import signal
            r, w = os.pipe()
            os.set_blocking(r, False)
            os.set_blocking(w, False)
            signal.set_wakeup_fd(w, warn_on_full_buffer=False)
            signal.signal(signal.SIGTERM, lambda signum, frame: self._fire("SIGTERM (another worker of the job died?)"))
            self.sig_r, self.sigterm = r, int(signal.SIGTERM)
        except (ValueError, OSError):
            pass   # not the main thread
        self.th = threading.Thread(target=self._watch, daemon=True)
        self.th.start()

    def publish(self, line):
        self.partial = json.loads(json.dumps(line))     # a deep, serialisable copy

    def enter(self, stage, limit_s):
        limit_s = float(os.environ.get("PF_BENCH_STAGE_LIMIT_S", limit_s))   # tests shorten it
        self.stage, self.deadline = stage, time.monotonic() + limit_s

    def leave(self):
        self.stage, self.deadline = "between stages", time.monotonic() + 600.0

    def remaining(self):
        """seconds left of the whole run's allowance"""
        return self.run_deadline - time.monotonic()

    def abort(self, stage, why):
        """called by the rank on which an optional stage raised: tells every rank's watchdog to fire"""
        log(f"[rank {self.rank}] stage '{stage}' failed: {why}")
        try:
            self.abort_file.write_text(f"rank {self.rank}: {stage}: {why}"[:2000])
        except OSError:
            pass
        self._fire(f"rank {self.rank}: {why}")

    def finish(self):
        self.done.set()

    def _watch(self):
        while not self.done.wait(0.25):
            if self.sig_r is not None:
                try:
                    if self.sigterm in os.read(self.sig_r, 64):
                        self._fire("SIGTERM (another worker of the job died?)")
                except (BlockingIOError, OSError):
                    pass
            if time.monotonic() > min(self.deadline, self.run_deadline):
                self._fire(f"stage exceeded its time limit on rank {self.rank}" if time.monotonic() <= self.run_deadline
                           else f"the run reached its overall time allowance on rank {self.rank}")
            if self.abort_file.exists():
                try:
                    why = self.abort_file.read_text()
                except OSError:
                    why = "abort requested by another rank"
                time.sleep(0.2 if self.rank == 0 else 1.0)
                self._fire(why)

    def _fire(self, why):
        if self.done.is_set():
            return
        self.done.set()
        for name in self.shm_names:
            try:
                os.unlink("/dev/shm/" + name)
            except OSError:
                pass
        if self.rank == 0:
            if self.partial is not None:
                line = dict(self.partial)
                line["aborted_stage"] = {"stage": self.stage, "why": str(why)[:1000]}
                emit(line)
            else:
                log(f"[bench] aborted in '{self.stage}' before a result existed: {why}")
        sys.stderr.flush()
        os._exit(0 if self.partial is not None or self.rank != 0 else 3)


class Comm:
    """torch.distributed plumbing of one job (or of rank 0 alone when `solo`)"""

    def __init__(self, world, rank, local_rank, dev, solo=False):
        self.world, self.rank, self.local_rank, self.dev, self.solo = (1, 0, local_rank, dev, True) if solo else \
            (world, rank, local_rank, dev, False)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()

    def all_reduce(self, t, op="sum"):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return t

    def all_gather_object(self, obj):
        if self.world == 1:
            return [obj]
        import torch.distributed as dist
        out = [None] * self.world
        dist.all_gather_object(out, obj)
        return out

    def broadcast_object(self, obj, src=0):
        if self.world == 1:
            return obj
        import torch.distributed as dist
        box = [obj]
        dist.broadcast_object_list(box, src=src)
        return box[0]


def parse_grid(txt, world):
    if not txt:
        return world, 1
    a, b = txt.lower().split("x")
    Lw, Qw = int(a), int(b)
    if Lw * Qw != world or Lw < 1 or Qw < 1:
        raise SystemExit(f"--grid {txt}: list shards x query groups must equal the number of ranks ({world})")
    return Lw, Qw


def split_queries(nq, Qw, qg):
    """queries [lo, hi) of a step served by query group qg"""
    return nq * qg // Qw, nq * (qg + 1) // Qw


# ----------------------------------------------------------------------------------------------
# one workload on one rank grid
# ----------------------------------------------------------------------------------------------
def run_workload(args, cfg_name, cfg, comm, weak, grid, steps, warmup, want_e2e=True, want_cpu=False, want_recall=False,
                 want_parity=False, sampler=None, tag="", guard=None, on_progress=None):
    """Loads the index, runs the device-resident timed region (`value`), the host-buffer pipeline (`e2e`) and the
    untimed checks.  Returns the record (on rank 0 of `comm`; None elsewhere)."""
    import torch
    import prefhetch_b200 as pf
    world, rank, dev, local_rank = comm.world, comm.rank, comm.dev, comm.local_rank
    Lw, Qw = grid
    lr, qg = rank % Lw, rank // Lw
    n, d, g, m, nprobe, nq = cfg["n"], cfg["d"], cfg["g"], cfg["m"], cfg["nprobe"], cfg["nq"]
    q_lo, q_hi = split_queries(nq, Qw, qg)
    nq_loc = q_hi - q_lo
    t_setup = time.perf_counter()
    if weak:
        nprobe = nprobe * Lw
        data = make_dataset_shard(cfg, dev, lr, Lw)
        cent, counts = data["centroids"], data["counts"]
        if Qw > 1:       # every query group holds a copy of every list shard: sum over ONE group only
            scale = 1.0 / Qw
            cent, counts = cent * scale, counts.double() * scale
        comm.all_reduce(cent)                       # every list is non-zero on exactly one list shard
        comm.all_reduce(counts)
        data["centroids"] = cent.cpu().numpy().astype(np.float32)
        global_list_sizes = counts.round().long().cpu().numpy().astype(np.int64)
    else:
        data = make_dataset(cfg, dev)
        global_list_sizes = (data["offsets"][1:] - data["offsets"][:-1]).astype(np.int64)
    torch.cuda.empty_cache()
    eng = pf.Engine(d, n, pf.bfv_default_primes(n), pf.batching_plain_modulus(n, cfg["tbits"]), m, g,
                    device=local_rank, rank=lr, world=Lw, result_limbs=args.result_limbs)
    info = eng.load_index(data["centroids"], data["offsets"], data["ids"], data["vectors"])
    eng.set_list_sizes(data["offsets"])
    L, k, K, C_ = eng.L, eng.k, info["K"], info["C"]
    ctw = eng.ctw
    stream = torch.cuda.Stream(device=dev)
    eng.set_stream(stream.cuda_stream)
    # synthetic Galois keys / query ciphertexts: uniform residues (timing does not depend on values;
    # parity with real encryptions is what tests/, smoke() and the self-check below verify)
    gen = torch.Generator(device=dev)
    gen.manual_seed(2025)
    primes = eng.primes

    def rand_residues(shape_prefix, limbs):
        cols = [torch.randint(0, primes[j], shape_prefix + (n,), generator=gen, device=dev, dtype=torch.int64)
                for j in limbs]
        return torch.stack(cols, dim=len(shape_prefix)).contiguous()

    for r in range(1, info["R"]):
        key = rand_residues((L, 2), range(k))  # [L][2][k][n]
        eng.set_galois_key(eng.galois_elt(r), key.cpu().numpy().view(np.uint64))
        del key
    npool_steps = 4 if nq_loc * m * ctw * 8 <= (1 << 30) else 2
    ct_pool = [rand_residues((max(nq_loc, 1), m, 2), range(L)) for _ in range(npool_steps)]  # [nq_loc][m][2][L][n]
    queries = data["queries"]
    nsteps_total = warmup + steps
    qsets = [queries[(s * nq) % (len(queries) - nq):][:nq][q_lo:q_hi] for s in range(nsteps_total + 2)]
    bpl_own = eng._blocks_per_list
    if world == 1:
        max_res = int(np.sort(bpl_own)[-min(len(bpl_own), nprobe):].sum()) * nq   # nprobe largest lists, every query
    else:  # probes spread over the list shards: 3x the expected per-rank share is far beyond its fluctuation
        mean_blocks = float(bpl_own[bpl_own > 0].mean())
        max_res = int(3.0 * max(nq_loc, 1) * nprobe / Lw * mean_blocks) + 256
    NBUF = 2 if world > 1 else 1
    d_outs = [torch.empty((max_res, 2, eng.Lr, n), dtype=torch.int64, device=dev) for _ in range(NBUF)]
    log(f"[{tag}rank {rank}] setup {time.perf_counter() - t_setup:.1f}s  index: {info}  max_res {max_res}  grid {Lw}x{Qw} queries [{q_lo},{q_hi})")

    # ---- multi-GPU gather: every rank's result ciphertexts go into RANK 0's HBM through a peer-mapped
    # buffer (CUDA IPC over NVLink): copy-engine DMA on a side stream, overlapped with the next step
    # (rank 0's NVLink ingest is the shared resource: 7 writers bursting from inside their last kernel stall
    # each other, measured 68 % efficiency at 8 GPUs in round 1; the fused-store path was removed from the
    # bench).  Stream-ordered arrival / ack flags (one-thread kernels on peer memory, bounded waits) tell rank 0
    # that a step has landed and the shards that their buffer is free.  NCCL only carries the set-up exchange
    # and the timing reduce.
    comm_stream = torch.cuda.Stream(device=dev, priority=-1) if world > 1 else None
    out_ptrs = [t.data_ptr() for t in d_outs]
    res_words = 2 * eng.Lr * n
    res_bytes = res_words * 8
    ipc_local, ipc_mapped, gather_bufs = [], [], {}
    copied_ev = [None] * NBUF
    FL = 128
    if world > 1:
        sizes_all = comm.all_gather_object(max_res)
        handles = None
        if rank == 0:
            handles = [[None] * world for _ in range(NBUF)]
            for b in range(NBUF):
                for r in range(1, world):
                    ptr, h = eng.ipc_alloc(sizes_all[r] * res_bytes)
                    ipc_local.append(ptr)
                    gather_bufs[(b, r)] = ptr
                    handles[b][r] = h
        handles = comm.broadcast_object(handles)
        if rank != 0:
            for b in range(NBUF):
                ipc_mapped.append(eng.ipc_open(handles[b][rank]))
        arr_ptr, arr_h = eng.ipc_alloc(FL * world) if rank == 0 else (None, None)
        ack_ptr, ack_h = eng.ipc_alloc(FL)
        torch.cuda.synchronize()
        allh = comm.all_gather_object((arr_h, ack_h))
        if rank == 0:
            ipc_local += [arr_ptr, ack_ptr]
            ack_peer = [None] + [eng.ipc_open(allh[r][1]) for r in range(1, world)]
            ipc_mapped += ack_peer[1:]
        else:
            ipc_local.append(ack_ptr)
            arr_ptr = eng.ipc_open(allh[0][0])
            ipc_mapped.append(arr_ptr)
        if rank == 0:
            for r in range(world):
                eng.flag_write(arr_ptr + FL * r, 0)
        eng.flag_write(ack_ptr, 0)
        eng.synchronize()
        comm.barrier()

    next_idx = {}
    host_t = {"wait": 0.0, "search": 0.0, "coarse": 0.0, "gather": 0.0, "n": 0}

    def step(s):
        """stage 1 of batch s+1 is issued right after stage 2 of batch s was enqueued, so the host-side
        planning of the next batch overlaps the GPU work of this one (a serving loop does the same)"""
        b = s % NBUF
        t0 = time.perf_counter()
        if world > 1 and rank != 0 and s >= NBUF and copied_ev[b] is not None:
            stream.wait_event(copied_ev[b])       # the DMA that last read this local buffer is done
        idx = next_idx.pop(s) if s in next_idx else eng.coarse_quantize(qsets[s], nprobe)
        t1 = time.perf_counter()
        rpq, st = eng.search_device(ct_pool[s % npool_steps].data_ptr(), nq_loc, idx, out_ptrs[b], max_res)
        t2 = time.perf_counter()
        # always one stage-1 call per step (the last one quantizes a batch that is never searched)
        if not os.environ.get("PF_BENCH_NO_PREFETCH"):
            next_idx[s + 1] = eng.coarse_quantize(qsets[s + 1], nprobe)
        t3 = time.perf_counter()
        host_t["wait"] += t1 - t0
        host_t["search"] += t2 - t1
        host_t["coarse"] += t3 - t2
        host_t["n"] += 1
        return idx, st

    def gather_results(s, idx, st):
        if world == 1:
            return
        tg = time.perf_counter()
        done = torch.cuda.Event()
        done.record(stream)
        comm_stream.wait_event(done)
        cs = comm_stream.cuda_stream
        if rank != 0:
            b = s % NBUF
            if s >= NBUF:
                eng.flag_wait(ack_ptr, s - NBUF + 1, cs)     # rank 0 is done with the peer buffer
            eng.copy_async(ipc_mapped[b], d_outs[b].data_ptr(), int(st["nresults"]) * res_bytes, cs)
            eng.flag_write(arr_ptr + FL * rank, s + 1, cs)
            ev = torch.cuda.Event()
            ev.record(comm_stream)
            copied_ev[b] = ev
        else:
            for r in range(1, world):
                eng.flag_wait(arr_ptr + FL * r, s + 1, cs)   # every shard's ciphertexts of step s are in HBM
            for r in range(1, world):
                eng.flag_write(ack_peer[r], s + 1, cs)       # ... the response can be assembled; buffer free
        host_t["gather"] += time.perf_counter() - tg

    # ---- value: device-resident timed region --------------------------------------------------
    with torch.cuda.stream(stream):
        for s in range(warmup):
            idx, st = step(s)
            gather_results(s, idx, st)
        eng.synchronize()
        torch.cuda.synchronize()
        comm.barrier()
        eng.timing_enable(True)
        eng.timing_read(reset=True)
        launches0 = eng.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        useful = slots = nres = 0
        blocks_distinct = pairs = 0
        t_host0 = time.perf_counter()
        ev0.record(stream)
        for s in range(warmup, nsteps_total):
            idx, st = step(s)
            gather_results(s, idx, st)
            useful += st["useful_distances"]
            slots += st["slot_distances"]
            nres += st["nresults"]
        if comm_stream is not None:             # the timed region ends when the last gather has landed
            stream.wait_stream(comm_stream)
        ev1.record(stream)
        eng.synchronize()
        torch.cuda.synchronize()
        t_host1 = time.perf_counter()
        comm.barrier()
        ms_total = ev0.elapsed_time(ev1)
        launches = eng.launch_count() - launches0
        phases = eng.timing_read(reset=True)
        eng.timing_enable(False)
        log(f"[{tag}rank {rank}] host ms/step: " + ", ".join(f"{k_} {1e3 * v / max(1, host_t['n']):.3f}" for k_, v in host_t.items() if k_ != "n"))
        log(f"[{tag}rank {rank}] {ms_total / steps:.3f} ms/step; phases " + ", ".join(f"{k_} {v['ms'] / steps:.3f}" for k_, v in phases.items()))

        # ---- gather verification (untimed): what landed in rank 0's buffers is what the ranks computed ----
        gather_verified = None
        if world > 1:
            s = nsteps_total
            idx, st = step(s)
            gather_results(s, idx, st)
            stream.wait_stream(comm_stream)
            eng.synchronize()
            torch.cuda.synchronize()
            nwords = int(st["nresults"]) * res_words
            mine = (int(st["nresults"]), eng.device_checksum(d_outs[s % NBUF].data_ptr(), nwords, stream.cuda_stream))
            allc = comm.all_gather_object(mine)
            if rank == 0:
                ok, bad = 0, []
                for r in range(1, world):
                    got = eng.device_checksum(gather_bufs[(s % NBUF, r)], allc[r][0] * res_words, stream.cuda_stream)
                    if got != allc[r][1]:
                        log(f"[{tag}rank 0] gather verification FAILED: rank {r}'s {allc[r][0]} results: checksum {allc[r][1]:#x} computed, {got:#x} landed")
                        bad.append(r)
                    else:
                        ok += 1
                gather_verified = {"ranks": ok, "results": int(sum(c[0] for c in allc[1:])), "bytes": int(sum(c[0] for c in allc[1:])) * res_bytes}
                if bad:
                    gather_verified["mismatch_ranks"] = bad
            comm.barrier()

    # distinct blocks per step for the algorithmic-bytes formula (host-side bookkeeping, untimed)
    for s in range(warmup, nsteps_total):
        idx = eng.coarse_quantize(qsets[s], nprobe)
        own = idx[(idx % Lw) == lr] if Lw > 1 else idx.reshape(-1)
        pairs += int(bpl_own[own].sum())
        blocks_distinct += int(bpl_own[np.unique(own)].sum())

    t_max = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    tot = torch.tensor([useful, slots, nres, launches, pairs, blocks_distinct], device=dev, dtype=torch.float64)
    comm.all_reduce(t_max, op="max")
    comm.all_reduce(tot)
    ms_total_max = float(t_max.item())
    useful_all, slots_all, nres_all, launches_all = (float(x) for x in tot[:4].tolist())
    ms_step = ms_total_max / steps
    value = useful_all / (ms_total_max * 1e-3)

    # ---- roofline of the MAC kernel (rank-local launch, measured with CUDA events on its stream) --
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # one MAC launch per step, except when the full-level result scratch is capped and a large batch is cut into
    # sub-batches of whole queries (pf_engine.cu search_core): bytes and time are both taken PER STEP, which is
    # bytes per launch / time per launch for equal launches
    mac_launches_per_step = phases["mac"]["launches"] / max(1, steps)
    mac_ms = phases["mac"]["ms"] / max(1, steps)
    LN8 = 8.0 * L * n
    alg_bytes = LN8 * (2.0 * K * nq_loc * steps + (K + 1.0) * blocks_distinct + 2.0 * pairs) / steps  # rank-local, per step
    streamed_bytes = LN8 * (2.0 * K * nq_loc * steps + (K + 1.0) * pairs + 2.0 * pairs) / steps
    achieved = alg_bytes / (mac_ms * 1e-3) / 1e9 if mac_ms > 0 else 0.0
    traffic, mac_name = None, "mac_kernel_occ" if K <= 16 else "mac_kernel"
    tf = ROOT / "profiles" / "mac_traffic.json"
    if tf.exists() and world == 1:   # dram__bytes_read+write per launch from the committed ncu --set full capture
        tj = json.loads(tf.read_text()).get(cfg_name + (f"_n{n}" if n != CONFIGS[cfg_name]["n"] else ""), {})
        traffic, mac_name = tj.get("traffic_bytes_per_launch"), tj.get("kernel", mac_name)
    roofline = {"bound": "hbm", "kernel": mac_name, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": "measured" if peaks else "fallback",
                "algorithmic_bytes_per_launch": alg_bytes / max(1.0, mac_launches_per_step),
                "streamed_bytes_per_launch": streamed_bytes / max(1.0, mac_launches_per_step),
                "algorithmic_bytes_per_step": alg_bytes,
                "streamed_gbs": streamed_bytes / (mac_ms * 1e-3) / 1e9 if mac_ms > 0 else 0.0,
                "ms_per_launch": mac_ms / max(1.0, mac_launches_per_step), "launches_per_step": mac_launches_per_step,
                "ms_per_step": mac_ms}
    # rotate phase (key switch): compute-pipe bound.  Essential work per rotation: 2 + 2L
    # N-point transforms of (N/2) log2 N butterflies at 8 FP64 operations each, and 2 (L+1) L N 64x64
    # multiply-accumulates of 3 IMAD.WIDE each; the two pipes run side by side (64 lanes/clk/SM each), so the
    # bound is the slower of the two at the SM clock seen during the run.
    rot_ms = phases["rotate"]["ms"] / steps
    nrot = nq_loc * m * (info["R"] - 1)
    fp64_ops = nrot * (2 + 2 * L) * (n // 2) * int(np.log2(n)) * 8.0
    imad_ops = nrot * 2.0 * (L + 1) * L * n * 3.0
    rotate_roofline = {"bound": "fp64+imad pipes", "rotations_per_step": nrot, "ms_per_step": rot_ms,
                       "fp64_ops_per_step": fp64_ops, "imad_wide_per_step": imad_ops, "lanes_per_clk_per_sm": 64, "sms": 148}

    rec = {"value": value, "ms_per_step": ms_step, "steps": steps, "warmup": warmup, "useful_per_step": useful_all / steps,
           "slot_distances_per_s": slots_all / (ms_total_max * 1e-3), "result_cts_per_step": nres_all / steps,
           "gpu_launches": int(launches_all), "roofline": roofline, "rotate_roofline": rotate_roofline,
           "phases_ms_per_step": {kk: v["ms"] / steps for kk, v in phases.items()},
           "queries_per_s": nq * steps / (ms_total_max * 1e-3), "timed_window": (t_host0, t_host1),
           "gather_verified": gather_verified, "info": info, "L": L, "Lr": eng.Lr, "nprobe": nprobe,
           "db_gib_per_rank": info["db_bytes"] / 2**30}

    if on_progress is not None:
        on_progress(rec)           # rank 0 publishes the line as it stands: nothing below may cost the headline

    def optional(stage, limit_s, fn):
        """run one optional stage under the guard; at one rank a failure is recorded, at several it aborts the
        run through the guard (the collectives of the ranks are out of step after an exception)"""
        if guard is not None:
            guard.enter(f"{tag}{stage}", limit_s)
        try:
            return fn()
        except Exception as ex:    # noqa: BLE001 — every failure of an optional stage is handled the same way
            if world > 1 and guard is not None:
                guard.abort(f"{tag}{stage}", repr(ex))
            log(f"[{tag}rank {rank}] {stage} failed: {ex!r}")
            return {"error": repr(ex)[:500]}
        finally:
            if guard is not None:
                guard.leave()

    # ---- recall (BASELINE metric, untimed; single GPU: the whole index is local) ----
    if want_recall and world == 1 and cfg["nb"] < (1 << 21):
        def _recall():
            rm = recall_metrics(eng, data, nprobe, dev)
            log(f"[{tag}rank {rank}] recall@10 = {rm['recall_at_10']:.4f} (reference definition {rm['reference_recall_10']:.4f}; nprobe {nprobe} of {cfg['nlist']} lists, "
                f"{rm['queries']} queries, exact brute-force ground truth)")
            return rm
        rec["recall"] = optional("recall", 300, _recall)

    # ---- e2e: host buffers through the public C-ABI calls --------------------------------------
    if want_e2e:
        def _e2e():
            # every rank must take the same path: agree first that the flights' buffers fit beside the resident DB
            free_b, _tot = torch.cuda.mem_get_info(dev)
            depth = int(os.environ.get("PF_BENCH_E2E_DEPTH", "3"))
            need = depth * (max_res * eng.slot_bytes + 2 * max(nq_loc, 1) * m * eng.ct_bytes) + (2 << 30)
            fits = torch.tensor([1.0 if need <= free_b else 0.0], device=dev, dtype=torch.float64)
            comm.all_reduce(fits)
            if int(fits.item()) == world:
                return run_e2e(args, eng, comm, cfg, grid, data, qsets, ct_pool, nprobe, max_res, steps, nsteps_total, tag, guard)
            why = (f"{depth} searches in flight need {need / 2**30:.1f} GiB beside the {info['db_bytes'] / 2**30:.1f} GiB DB; "
                   f"{free_b / 2**30:.1f} GiB free on this GPU (shard the index over more GPUs)")
            log(f"[{tag}rank {rank}] e2e skipped: {why}")
            return {"value": None, "unit": "distances/s", "skipped": why}
        rec["e2e"] = optional("e2e", 300, _e2e)
        if on_progress is not None:
            on_progress(rec)

    # ---- parity self-check on real encryptions (untimed; the oracle is the checker) -------------
    if want_parity and world == 1:
        rec["parity"] = optional("parity self-check", 600, lambda: parity_self_check(eng, cfg, data, nprobe, tag))
        if rec["parity"].get("error"):
            rec["parity"]["parity_checked"] = 0
        if on_progress is not None:
            on_progress(rec)

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) -------------------------------------
    if want_cpu and rank == 0 and world == 1:
        def _cpu():
            from oracle import pf_oracle as O
            O.build()
            nthreads = O.host_cores()
            nq_s = cpu_sample_queries(nq, nthreads)   # ~10 s of CPU work on the host cores
            rl = eng.Lr if eng.Lr < L else 0
            r = cpu_pipeline(cfg, data, nq_s, nthreads, result_limbs=rl)
            r1 = cpu_pipeline(cfg, data, 2, 1, result_limbs=rl)
            out = {"value": r["useful"] / r["seconds"], "unit": "distances/s", "cores": nthreads, "kind": "port",
                   "sample": f"{nq_s} of the {nq} queries of a step / {r['pairs']} (query,block) pairs of the same workload, whole hot path "
                             f"incl. mod-switch to {eng.Lr} limb(s), {nthreads} OpenMP threads ({r['seconds']:.2f}s: rotations {r['rot_s']:.2f}s, "
                             f"MAC+INTT {r['mac_s']:.2f}s)",
                   "single_thread_value": r1["useful"] / r1["seconds"], "queries_per_s": nq_s / r["seconds"]}
            if cfg["nb"] <= 200_000:
                out["reference_plaintext_path"] = reference_plaintext_path(data, cfg["nprobe"], min(100, len(data["queries"])))
            return out
        rec["cpu_baseline"] = optional("cpu baseline", 600, _cpu)
        if rec["cpu_baseline"].get("error"):
            rec["cpu_baseline"]["value"] = None

    # ---- recall on a mixture where it is < 1 (after every stage the headline needs: it builds a second data set and a
    # second engine, and none of that has a claim on the run's time allowance before e2e / parity / the CPU baseline)
    if want_recall and world == 1 and cfg["nb"] < (1 << 21):
        # The SURVEY §8d mixture is well separated (recall@10 = 1 at any sensible nprobe): the same search on an
        # OVERLAPPING mixture — centres in [64,192]^d, sigma 48 (cluster radius ~ centre spacing), lists re-fitted
        # with two Lloyd iterations, same nlist / nprobe — through a second engine (plaintext stages only).
        def _recall_hard():
            hcfg = dict(cfg)
            hcfg["nb"] = min(cfg["nb"], 500_000)
            hd = make_dataset(hcfg, dev, seed=777, sigma=48.0, spread=128.0, lloyd=2, offset=64.0)
            torch.cuda.empty_cache()
            eng2 = pf.Engine(d, n, pf.bfv_default_primes(n), pf.batching_plain_modulus(n, cfg["tbits"]), m, g,
                             device=local_rank, result_limbs=args.result_limbs)
            try:
                eng2.load_index(hd["centroids"], hd["offsets"], hd["ids"], hd["vectors"])
                rm = recall_metrics(eng2, hd, nprobe, dev)
            finally:
                eng2.close()
            rm["dataset"] = (f"{hcfg['nb']} vectors, {cfg['nlist']} overlapping clusters (centres uniform in [64,192]^{d}, sigma 48, "
                             f"2 Lloyd iterations), nprobe {nprobe}")
            log(f"[{tag}rank {rank}] recall@10 on the overlapping mixture = {rm['recall_at_10']:.4f} (reference definition {rm['reference_recall_10']:.4f})")
            return rm
        rec["recall_hard"] = optional("recall (overlapping mixture)", 300, _recall_hard)
        if on_progress is not None:
            on_progress(rec)

    # ---- the e2e pass again with the request a symmetric-key SEAL client sends (seeded streams: half the upload, c1
    # drawn on the device).  One rank, main workload only, LAST: an extra record beside the headline e2e, and nothing
    # it does to the engine can touch the stages above.
    if want_e2e and world == 1 and want_parity and (rec.get("e2e") or {}).get("value"):
        rec["e2e_seeded_requests"] = optional("e2e (seeded requests)", 300, lambda: run_e2e(
            args, eng, comm, cfg, grid, data, qsets, ct_pool, nprobe, max_res, steps, nsteps_total, tag, guard, seeded=True))
        if on_progress is not None:
            on_progress(rec)

    if world > 1:
        torch.cuda.synchronize()
        comm.barrier()
        for p_ in ipc_mapped:
            eng.ipc_close(p_)
        comm.barrier()
        for p_ in ipc_local:
            eng.ipc_free(p_)
    eng.close()
    del d_outs, ct_pool, data
    torch.cuda.empty_cache()
    comm.barrier()
    return rec if rank == 0 else None


class NodeResponse:
    """One request / response buffer for the whole node.  Layout: [flags 4 KiB][query blob][share of rank 0]
    [share of rank 1]...  With one rank it is a pinned allocation; with several it is a POSIX shared-memory
    segment created by rank 0 and attached by the others (each rank page-locks the parts its GPU touches), so
    every GPU moves its share over its own PCIe link and the handler on rank 0 returns one buffer.
    flags[r] (int64) = number of steps rank r has completed; rank 0 polls them.
    shared_data=False (no room in /dev/shm for the whole buffer): only the flags are shared, every rank keeps
    its query copy and its share in private pinned memory."""
    FLAGS = 4096

    def __init__(self, rank, world, name, qbytes, shares, pinned_alloc=None, shared_data=True):
        self.rank, self.world = rank, world
        self.qbytes = int(qbytes)
        self.qpad = (self.qbytes + 4095) // 4096 * 4096
        self.share_off = np.concatenate([[0], np.cumsum(shares)]).astype(np.int64)
        self.shared_data = bool(shared_data) and world > 1
        data_bytes = self.qpad + int(self.share_off[-1])
        self.total = self.FLAGS + (data_bytes if (self.shared_data or world == 1) else 0)
        self.shm = None
        alloc = pinned_alloc or (lambda nbytes: np.zeros(nbytes, dtype=np.uint8))

        def as_np(x):
            return x if isinstance(x, np.ndarray) else x.numpy()
        if world > 1:
            from multiprocessing import shared_memory
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=self.total) if rank == 0 else \
                shared_memory.SharedMemory(name=name)
            self.whole = np.ndarray((self.total,), dtype=np.uint8, buffer=self.shm.buf)
        else:
            self._keep = alloc(self.total)
            self.whole = as_np(self._keep)
        self.flags = self.whole[:self.FLAGS].view(np.int64)
        if self.shared_data or world == 1:
            self.query = self.whole[self.FLAGS:self.FLAGS + self.qbytes]
            o = self.FLAGS + self.qpad + int(self.share_off[rank])
            self.share = self.whole[o:o + int(shares[rank])]
        else:
            self._keep = alloc(self.qpad + int(shares[rank]))
            priv = as_np(self._keep)
            self.query = priv[:self.qbytes]
            self.share = priv[self.qpad:self.qpad + int(shares[rank])]

    def share_of(self, r):
        assert self.shared_data or self.world == 1
        o = self.FLAGS + self.qpad + int(self.share_off[r])
        return self.whole[o:o + int(self.share_off[r + 1] - self.share_off[r])]

    def query_region(self):
        return self.whole[self.FLAGS:self.FLAGS + self.qpad]

    def mark_done(self, step):
        self.flags[self.rank] = step + 1

    def wait_all(self, step, timeout_s=60.0):
        """rank 0: the response of `step` is complete when every rank has marked it"""
        t0 = time.perf_counter()
        while int(self.flags[:self.world].min()) < step + 1:
            if time.perf_counter() - t0 > timeout_s:
                raise TimeoutError(f"ranks {[r for r in range(self.world) if self.flags[r] < step + 1]} did not finish step {step}")

    def close(self):
        self.flags = self.query = self.share = self.whole = None
        if self.shm is not None:
            try:
                self.shm.close()
                if self.rank == 0:
                    self.shm.unlink()
            except Exception:
                pass


def open_node_response(comm, qbytes, seg, pinned_alloc=None, guard=None):
    """Collective: every rank calls it with its share size; the SAME sequence of collectives on every rank
    (all_gather, broadcast, barrier, barrier) — rank 0 creates the segment between the broadcast and the first
    barrier, the others attach after it."""
    segs = comm.all_gather_object(int(seg))
    name, shared = comm.broadcast_object((f"pf_bench_{os.environ.get('MASTER_PORT', '0')}_{os.getpid()}",
                                          shm_has_room(4096 + int(qbytes) + int(sum(segs)))) if comm.rank == 0 else None)
    if comm.world == 1:
        return NodeResponse(0, 1, name, qbytes, segs, pinned_alloc=pinned_alloc)
    if guard is not None:
        guard.shm_names.append(name)
    nr = None
    if comm.rank == 0:
        nr = NodeResponse(0, comm.world, name, qbytes, segs, pinned_alloc=pinned_alloc, shared_data=shared)
    comm.barrier()                                # the segment exists
    if comm.rank != 0:
        nr = NodeResponse(comm.rank, comm.world, name, qbytes, segs, pinned_alloc=pinned_alloc, shared_data=shared)
    comm.barrier()                                # everybody is attached
    return nr


def shm_has_room(nbytes):
    if os.environ.get("PF_BENCH_PRIVATE_RESPONSE") == "1":      # tests: take the fallback
        return False
    try:
        st = os.statvfs("/dev/shm")
        return st.f_bavail * st.f_frsize > nbytes + (256 << 20)
    except OSError:
        return False


def seeded_stream_parts(full_header, half_bytes, seed):
    """(113-byte header, 81-byte PRNG info) of a SEAL seeded ciphertext stream (Serializable<Ciphertext>: c0 + the seed
    of c1) made from the header of the full stream: total size, DynArray size and word count describe ONE polynomial,
    and a nested {SEALHeader, prng_type blake2xb = 1, 64-byte seed} follows the data (layout: DESIGN.md §7 f-3)."""
    import struct
    hdr = np.array(full_header, dtype=np.uint8, copy=True)
    hdr[8:16] = np.frombuffer(struct.pack("<Q", 113 + half_bytes + 81), dtype=np.uint8)
    hdr[97:105] = np.frombuffer(struct.pack("<Q", 16 + 8 + half_bytes), dtype=np.uint8)
    hdr[105:113] = np.frombuffer(struct.pack("<Q", half_bytes // 8), dtype=np.uint8)
    info = np.zeros(81, dtype=np.uint8)
    info[:16] = hdr[:16]
    info[8:16] = np.frombuffer(struct.pack("<Q", 81), dtype=np.uint8)
    info[16] = 1
    info[17:] = np.frombuffer(seed, dtype=np.uint8)
    return hdr, info


def run_e2e(args, eng, comm, cfg, grid, data, qsets, ct_pool, nprobe, max_res, steps, nsteps_total, tag, guard=None, seeded=False):
    """End to end through pf_search_submit / pf_search_collect with HOST buffers.  One response buffer for the
    whole node: at N > 1 it is a POSIX shared-memory segment page-locked by every rank (pf_host_register); the
    query blob of a step is read from it and every rank's GPU writes its share of the response into it over
    its own PCIe link — what a handler on rank 0 returns is complete when every rank has signalled the step.
    Requests are pipelined three deep (one downloading, one computing, one queued).  Timed on the host (it includes the H2D of the query ciphertexts and
    query vectors and the D2H of every result ciphertext and the probe ids), max over ranks by construction:
    rank 0 only advances when every rank has finished the step."""
    import torch
    world, rank, dev = comm.world, comm.rank, comm.dev
    Lw, Qw = grid
    lr, qg = rank % Lw, rank // Lw
    n, m, nq = cfg["n"], cfg["m"], cfg["nq"]
    q_lo, q_hi = split_queries(nq, Qw, qg)
    nq_loc = q_hi - q_lo
    L, ctw, ctb = eng.L, eng.ctw, eng.ct_bytes
    hdr = np.frombuffer(eng.ct_serialize(np.zeros((2, L, n), dtype=np.uint64)), dtype=np.uint8)[:ctb - ctw * 8]
    # seeded=True: the request a symmetric-key SEAL client sends — c0 and the 64-byte seed of c1 (half the bytes); the
    # engine draws c1 on the device.  c0 = the pool's residues, seeds arbitrary: timing does not depend on the values.
    info = None
    if seeded:
        hdr, info = seeded_stream_parts(hdr, ctw * 4, bytes((7 * i + 1) & 255 for i in range(64)))
        ctb = len(hdr) + ctw * 4 + len(info)
    qbytes = nq * m * ctb
    seg = max_res * eng.slot_bytes
    nr = open_node_response(comm, qbytes, seg, pinned_alloc=lambda nbytes: torch.empty(nbytes, dtype=torch.uint8).pin_memory(),
                            guard=guard)
    flags, qnp, out_np = nr.flags, nr.query, nr.share
    numa = bind_pages_to_gpu_node(out_np, dev) if (world > 1 and nr.shared_data) else None
    out_np[:] = 0                                 # first touch by the owner: pages on this rank's NUMA node
    if world > 1:
        log(f"[{tag}rank {rank}] response share: {numa}")
    if rank == 0:
        flags[:] = 0
    # the request of a step: nq*m SEAL streams (the same bytes every step; timing does not depend on values);
    # every query group writes the ciphertexts it holds on its GPU into the shared request blob
    if (lr == 0 or not nr.shared_data) and nq_loc:
        mine = ct_pool[0].cpu().numpy().view(np.uint64).reshape(nq_loc * m, -1)
        for c in range(nq_loc * m):
            o = (q_lo * m + c) * ctb
            qnp[o: o + len(hdr)] = hdr
            if seeded:
                qnp[o + len(hdr): o + len(hdr) + ctw * 4] = mine[c].view(np.uint8)[:ctw * 4]
                info[17 + (c & 63)] = (c * 13 + 5) & 255          # a different seed per ciphertext
                qnp[o + ctb - len(info): o + ctb] = info
            else:
                qnp[o + len(hdr): o + ctb] = mine[c].view(np.uint8)
    registered = world > 1 and nr.shared_data
    if world > 1:
        comm.barrier()
    if registered:
        eng.host_register(nr.query_region())
        eng.host_register(out_np)
    my_q = qnp[q_lo * m * ctb: q_hi * m * ctb]
    offs = (np.arange(nq_loc * m + 1, dtype=np.uint64) * ctb)
    eng.set_search_groups(int(os.environ.get("PF_BENCH_E2E_GROUPS", "1")))
    e_steps = max(2, min(steps, 40))
    DEPTH = int(os.environ.get("PF_BENCH_E2E_DEPTH", "3"))   # searches in flight: one downloading, one computing, one queued
    e_useful, h2d, d2h = 0, 0, 0
    t_cq = t_sub = t_col = 0.0
    pend = []
    comm.barrier()
    torch.cuda.synchronize()

    def finish(p, s):
        nonlocal t_col
        ta = time.perf_counter()
        p.collect()
        nr.mark_done(s)
        if rank == 0 and world > 1:                # the response of step s leaves the node when every share is in
            nr.wait_all(s)
        t_col += time.perf_counter() - ta

    WARM = DEPTH + 2        # every flight's buffers have been allocated (first use) before the clock starts
    for s in range(WARM + e_steps):
        if s == WARM:
            while pend:
                finish(*pend.pop(0))
            comm.barrier()
            t0 = time.perf_counter()
        x = qsets[s % nsteps_total]
        ta = time.perf_counter()
        idx = eng.coarse_quantize(x, nprobe)
        tb = time.perf_counter()
        p = eng.submitSearchEncrypted(my_q, offs, idx, out=out_np) if nq_loc else None
        tc = time.perf_counter()
        if p is not None:
            pend.append((p, s))
        else:
            nr.mark_done(s)
        if len(pend) >= DEPTH:
            finish(*pend.pop(0))
        if s >= WARM:
            t_cq += tb - ta
            t_sub += tc - tb
            if p is not None:
                e_useful += p.result.stats["useful_distances"]
                h2d += nq_loc * m * ctb + x.nbytes
                d2h += p.result.stats["out_bytes"] + idx.nbytes
    while pend:
        finish(*pend.pop(0))
    e_dt = time.perf_counter() - t0
    # one lone request, start to finish (latency of the one-call form, default query groups)
    eng.set_search_groups(0)
    lat = None
    if nq_loc:
        torch.cuda.synchronize()
        ta = time.perf_counter()
        idx = eng.coarse_quantize(qsets[0], nprobe)
        eng.coarseSearchEncrypted(my_q, offs, idx, out=out_np)
        lat = (time.perf_counter() - ta) * 1e3
    log(f"[{tag}rank {rank}] e2e {e_dt / e_steps * 1e3:.3f} ms/step pipelined (host: coarse_quantize {t_cq / e_steps * 1e3:.3f}, "
        f"submit {t_sub / e_steps * 1e3:.3f}, collect+wait {t_col / e_steps * 1e3:.3f}); lone request {lat if lat is None else round(lat, 3)} ms")
    et = torch.tensor([e_dt], device=dev, dtype=torch.float64)
    eu = torch.tensor([e_useful, h2d, d2h], device=dev, dtype=torch.float64)
    comm.all_reduce(et, op="max")
    comm.all_reduce(eu)
    if world > 1:
        comm.barrier()
    if registered:
        eng.host_unregister(nr.query_region())
        eng.host_unregister(out_np)
    shared_note = nr.shared_data
    del flags, qnp, out_np, my_q
    comm.barrier()
    nr.close()
    return {"value": float(eu[0].item()) / float(et.item()), "unit": "distances/s",
            "h2d_bytes_per_step": int(eu[1].item()) // e_steps, "d2h_bytes_per_step": int(eu[2].item()) // e_steps,
            "ms_per_step": float(et.item()) / e_steps * 1e3, "steps": e_steps, "pipeline_depth": DEPTH,
            "lone_request_ms": lat, "request": ("seeded ciphertext streams (c0 + PRNG seed), c1 expanded on the device" if seeded else "full ciphertext streams"),
            "response": ("one host buffer per node" + (" (POSIX shm page-locked by every rank; each GPU writes its share over its own PCIe link)" if world > 1 else " (pinned)"))
            if (shared_note or world == 1) else "per-rank pinned buffers, completion flags shared (/dev/shm too small for one node buffer)"}


def parity_self_check(eng, cfg, data, nprobe, tag, nsample=32):
    """One untimed search of a full batch of REAL encryptions through the engine the timed loops used (same index,
    same plan shapes: key-switch query groups, shared blocks, multi-block lists), checked by the CPU oracle:
    `nsample` results spread over the whole batch byte for byte against the oracle pipeline, and decrypted to
    the exact integer distances.  The oracle plays the client (keygen / encrypt / decrypt) and the checker."""
    from oracle import pf_oracle as O
    from tests.util import OracleClient
    O.build()
    n, d, g, m, nq = cfg["n"], cfg["d"], cfg["g"], cfg["m"], cfg["nq"]
    primes, t = O.BFV_DEFAULT_PRIMES[n], O.BATCHING_T[(n, cfg["tbits"])]
    t0 = time.perf_counter()
    cl = OracleClient(O, n, primes, t, d, m, g)
    keys = cl.step_keys()
    for i, key in enumerate(keys):                 # the timed loops ran on random residues: real keys now
        eng.set_galois_key(eng.galois_elt(i + 1), key)
    query = np.ascontiguousarray(data["queries"][1000:1000 + nq])
    cts = np.stack([cl.encrypt_query(q, 9000 + 7 * i) for i, q in enumerate(query)])
    blob, offs = cl.serialize_queries(cts)
    idx = eng.coarse_quantize(query, nprobe)
    oidx, _ = O.coarse_quantize(query, data["centroids"], nprobe)
    assert np.array_equal(idx, oidx), "probed lists differ from the reference semantics"
    eng.set_search_groups(0)
    res = eng.coarseSearchEncrypted(blob, offs, idx)
    P = res.stats["nresults"]
    offsets, vecs = data["offsets"], data["vectors"]
    C_ = cl.lay.C
    # (query, probe, block) of every result, in response order
    where = []
    for qi in range(nq):
        for p in range(idx.shape[1]):
            l = int(idx[qi, p])
            for b0 in range(0, int(offsets[l + 1] - offsets[l]), C_):
                where.append((qi, l, b0))
    assert len(where) == P, f"{P} results, {len(where)} expected"
    pick = sorted(set(np.linspace(0, P - 1, nsample).astype(int).tolist()))
    rl = eng.Lr if eng.Lr < eng.L else 0
    pid = (0, 0, 0, 0)
    if rl:
        import hashlib
        import struct
        words = [1, n, *primes[:rl], t]
        pid = struct.unpack("<4Q", hashlib.blake2b(struct.pack(f"<{len(words)}Q", *words), digest_size=32).digest())
    rots = {}
    for r in pick:
        qi, l, b0 = where[r]
        if qi not in rots:
            rots[qi] = O.rotate_query_set(cl.ctx, cl.lay, cts[qi], keys, False)
        n_l = int(offsets[l + 1] - offsets[l])
        xs = vecs[offsets[l] + b0: offsets[l] + min(b0 + C_, n_l)].astype(np.int32)
        diag, norm = O.encode_block(cl.ctx, cl.lay, xs)
        want = O.block_distance(cl.ctx, cl.lay, rots[qi], diag, norm)
        if rl:
            want = cl.mod_switch_to(want, rl)
        got = res.result(r)
        assert got == cl.ctx.ct_save(want, parms_id=pid), f"result {r} (query {qi}, list {l}, block {b0}) differs from the oracle"
        ct, _ = eng.ct_deserialize(got)
        dist, budget = cl.distances(ct, query[qi], len(xs))
        assert np.array_equal(dist, ((xs.astype(np.int64) - query[qi].astype(np.int64)) ** 2).sum(1)), f"result {r}: decrypted distances not exact"
        assert budget > 0
    log(f"[{tag}] parity self-check: {len(pick)} of {P} results of a {nq}-query batch bit-exact vs the oracle and decrypted exactly "
        f"({time.perf_counter() - t0:.1f}s)")
    return {"parity_checked": len(pick), "results_in_batch": int(P), "queries": int(nq), "queries_sampled": len(rots),
            "checker": "oracle/ (CPU restatement; parity unpinned by the reference, DESIGN.md §0)"}


def build_line(args, cfg_name, cfg, world, weak, grid, rec, clocks, placement, strong):
    """the one JSON line of the run from the record run_workload returned (rank 0)"""
    recall = rec.get("recall") or {}
    line = {
        "metric": "encrypted candidate distances/sec", "value": rec["value"], "unit": "distances/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_step"],
        "higher_is_better": True, "scaling": "strong" if (world > 1 and not weak) else "weak", "vs_baseline": None,
        "scaling_note": ("N = 1 point of the series: the default run at N GPUs is weak scaling of this workload (N list shards of 1M vectors), "
                         "`--config X` at N > 1 is strong scaling of X") if world == 1 else None,
        "dtype": "u64", "data": "synthetic",
        "config": bench_config(cfg_name, cfg, rec["L"], rec["Lr"], cfg["nq"], world, weak, rec["nprobe"], grid),
        "queries_per_s": rec["queries_per_s"],
        "recall_at_10": recall.get("recall_at_10"), "reference_recall_10": recall.get("reference_recall_10"),
        "recall_overlapping_mixture": rec.get("recall_hard"),
        "slot_distances_per_s": rec["slot_distances_per_s"],
        "result_cts_per_step": rec["result_cts_per_step"],
        "gpu_launches": rec["gpu_launches"],
        "clocks": clocks, "roofline": rec["roofline"], "rotate_roofline": dict(rec["rotate_roofline"]),
        "phases_ms_per_step": rec["phases_ms_per_step"],
        "db_gib_per_rank": rec["db_gib_per_rank"],
        "e2e": rec.get("e2e"), "e2e_seeded_requests": rec.get("e2e_seeded_requests"), "cpu_baseline": rec.get("cpu_baseline"),
        "parity_checked": (rec.get("parity") or {}).get("parity_checked"), "parity": rec.get("parity"),
        "gather_verified": rec.get("gather_verified"), "placement": placement,
    }
    if clocks and clocks.get("sm_mhz"):
        rr = line["rotate_roofline"]
        per_s = 64.0 * 148 * clocks["sm_mhz"] * 1e6
        floor_ms = max(rr["fp64_ops_per_step"], rr["imad_wide_per_step"]) / per_s * 1e3
        rr.update({"floor_ms_per_step": floor_ms, "frac": floor_ms / rr["ms_per_step"] if rr["ms_per_step"] else None,
                   "sm_mhz": clocks["sm_mhz"]})
    if strong is not None:
        line["strong"] = strong
    return line


def main():
    global _JSON_OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=list(CONFIGS))
    ap.add_argument("--nq", type=int, default=None, help="queries per step")
    ap.add_argument("--g", type=int, default=None, help="partial-sum factor of the layout")
    ap.add_argument("--nprobe", type=int, default=None)
    ap.add_argument("--poly-degree", type=int, default=None, choices=[8192, 16384], help="N (configs[4] sweeps 8192 vs 16384)")
    ap.add_argument("--grid", default=None, help="LxQ: list shards x query groups (default Nx1)")
    ap.add_argument("--result-limbs", type=int, default=1,
                    help="limbs of the result ciphertexts (SEAL mod_switch_to before save); 0 = no switching")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="N > 1 default run: skip the configs[2] strong-scaling record")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    # stdout must carry exactly one JSON line: native libraries (NCCL prints its version banner with
    # printf) write to fd 1, so fd 1 is pointed at stderr for the whole run and the line goes to a dup
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args.warmup = max(args.warmup, 3)      # timing rule: at least 3 warm-up steps

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg_name = args.config or "sift1m_nlist1024_nprobe16"
    cfg = dict(CONFIGS[cfg_name])
    # N > 1 default: WEAK scaling of configs[1] — every list shard holds its own 1M-vector / 1024-list
    # share of a world x larger index and the query probes 16*world lists (16 per shard on average) —
    # plus a `strong` record on BASELINE configs[2].  `--config X` at N > 1 runs X as a fixed index (strong).
    weak = world > 1 and args.config is None
    if args.nq:
        cfg["nq"] = args.nq
    if args.nprobe:
        cfg["nprobe"] = args.nprobe
    if args.poly_degree and args.poly_degree != cfg["n"]:
        cfg["g"] = cfg["g"] * args.poly_degree // cfg["n"]      # same candidates per block
        cfg["n"] = args.poly_degree
    if args.g:
        cfg["g"] = args.g

    if args.impl == "reference":
        return run_reference(args, cfg, cfg_name)

    import torch
    import torch.distributed as dist
    # PF_BENCH_DRYRUN=1 (tests/test_bench_dryrun.py): the control flow of this file on the CPU — gloo, a stub engine,
    # no-op streams — to prove that every rank takes the same path through the collectives.  Measures nothing.
    # PF_BENCH_DRYRUN=emul (tests/test_cuda_emulated.py): the same no-op streams, but the REAL prefhetch_b200.Engine over the
    # CPU-emulated build of the CUDA sources (PF_LIB = tests/cuda_emul's library; several ranks with PF_EMUL_IPC=1): the whole flow — timed
    # loops, e2e submit / collect, recall, parity self-check, CPU baseline — with real arithmetic.  Measures nothing either.
    dry_mode = os.environ.get("PF_BENCH_DRYRUN", "")
    dry = dry_mode in ("1", "emul")
    if dry:
        from tests import bench_stub
        stub = bench_stub.install(torch)
        if dry_mode == "1":
            sys.modules["prefhetch_b200"] = stub
        elif not os.environ.get("PF_LIB"):
            raise SystemExit("PF_BENCH_DRYRUN=emul needs PF_LIB (the emulated build of tests/cuda_emul)")
        for key, env in (("nq", "PF_BENCH_DRYRUN_NQ"), ("nprobe", "PF_BENCH_DRYRUN_NPROBE"), ("nlist", "PF_BENCH_DRYRUN_NLIST")):
            if os.environ.get(env):       # emulated runs: every workload of the job (strong record, configs[4]) shrinks with it
                for c in list(CONFIGS.values()) + [cfg]:
                    c[key] = min(c[key], int(os.environ[env]))
        if os.environ.get("PF_BENCH_DRYRUN_NB"):
            for c in CONFIGS.values():
                c["nb"] = min(c["nb"], int(os.environ["PF_BENCH_DRYRUN_NB"]))
            cfg["nb"] = min(cfg["nb"], int(os.environ["PF_BENCH_DRYRUN_NB"]))
    else:
        import prefhetch_b200 as pf   # noqa: F401
        from prefhetch_b200.build import build as build_lib
        if rank == 0:
            build_lib()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cpu") if dry else torch.device("cuda", local_rank)
    placement = pin_to_gpu_numa_node(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line
        from datetime import timedelta
        if dry:
            dist.init_process_group("gloo", timeout=timedelta(minutes=5))
        else:
            dist.init_process_group("nccl", device_id=dev, timeout=timedelta(minutes=40))   # the StageGuard fires long before
        dist.barrier()
    comm = Comm(world, rank, local_rank, dev)
    grid = parse_grid(args.grid, world)
    if weak and grid[1] != 1:
        raise SystemExit("weak scaling shards the lists only: use --config with --grid LxQ")
    log(f"[rank {rank}] placement {placement}")

    sampler = ClockSampler(local_rank, dev)
    if rank == 0:                          # one poller per job: NVML queries take driver locks
        sampler.start()
    guard = StageGuard(rank, world)
    state = {"strong": None, "configs4": None}

    def make_line(rec):
        line = build_line(args, cfg_name, cfg, world, weak, grid, rec, sampler.report(*rec["timed_window"]), placement, state["strong"])
        if state["configs4"] is not None:
            line["configs4"] = state["configs4"]
        return line

    def progress(rec):
        if rank == 0:
            guard.publish(make_line(rec))

    rec = run_workload(args, cfg_name, cfg, comm, weak, grid, args.steps, args.warmup, want_e2e=not args.no_e2e,
                       want_cpu=not args.no_cpu_baseline, want_recall=True, want_parity=not args.no_parity, sampler=sampler,
                       guard=guard, on_progress=progress)
    if rank == 0:
        guard.publish(make_line(rec))

    def time_for(stage, need_s):
        """optional extras start only with enough of the run's allowance left; rank 0 decides for everybody"""
        ok = comm.broadcast_object(guard.remaining() > need_s) if world > 1 else guard.remaining() > need_s
        if not ok:
            log(f"[rank {rank}] skipping {stage}: {guard.remaining():.0f}s of the run's allowance left, {need_s}s wanted")
        return ok

    if weak and not args.no_strong and not time_for("the strong-scaling record", 240):
        state["strong"] = {"skipped": "not enough of the run's time allowance left (PF_BENCH_RUN_LIMIT_S)"}
    elif weak and not args.no_strong:
        # BASELINE configs[2] (fixed 1M index, nlist 4096, nprobe 64): N GPUs, then rank 0 alone (its 1-GPU time).
        # An optional stage: a failure or a hang here still leaves the weak-scaling headline (StageGuard).
        scfg_name = "sift1m_nlist4096_nprobe64"
        scfg = dict(CONFIGS[scfg_name])
        ssteps = max(3, min(args.steps, 60))
        sgrid = parse_grid(os.environ.get("PF_BENCH_STRONG_GRID"), world) if os.environ.get("PF_BENCH_STRONG_GRID") else default_strong_grid(world)
        guard.enter("strong-scaling record (configs[2])", 600)
        try:
            rN = run_workload(args, scfg_name, scfg, comm, False, sgrid, ssteps, args.warmup, want_e2e=not args.no_e2e, tag="strong ",
                              guard=guard)
            guard.enter("strong-scaling record (configs[2], 1-GPU reference on rank 0)", 420)
            r1 = None
            if rank == 0:
                r1 = run_workload(args, scfg_name, scfg, Comm(world, rank, local_rank, dev, solo=True), False, (1, 1), ssteps, args.warmup,
                                  want_e2e=not args.no_e2e, tag="strong-1gpu ", guard=guard)
            comm.barrier()
            if rank == 0:
                e2eN, e2e1 = rN.get("e2e") or {}, r1.get("e2e") or {}
                state["strong"] = {
                    "config": bench_config(scfg_name, scfg, rN["L"], rN["Lr"], scfg["nq"], world, False, grid=sgrid),
                    "scaling": "strong", "n_gpus": world, "value": rN["value"], "ms_per_step": rN["ms_per_step"],
                    "value_1gpu": r1["value"], "ms_per_step_1gpu": r1["ms_per_step"],
                    "efficiency": rN["value"] / (world * r1["value"]), "steps": ssteps,
                    "phases_ms_per_step": rN["phases_ms_per_step"], "phases_ms_per_step_1gpu": r1["phases_ms_per_step"],
                    "gather_verified": rN["gather_verified"], "e2e": rN.get("e2e"), "e2e_1gpu": r1.get("e2e"),
                    "e2e_efficiency": (e2eN["value"] / (world * e2e1["value"])) if e2eN.get("value") and e2e1.get("value") else None}
        except Exception as ex:    # noqa: BLE001
            guard.abort("strong-scaling record (configs[2])", repr(ex))
        guard.leave()

    # BASELINE configs[4] (10M x 128, nlist 16384, batch of 256 encrypted queries) is specified on 8 GPUs: the default
    # 8-GPU run records it as one more optional stage (N = 8192; the N = 16384 point of the sweep is a 1-GPU number in
    # profiles/).  Last, bounded and guarded: whatever happens here, the headline and the strong record are already
    # published.  PF_BENCH_EXTRAS=0 skips it.
    want_configs4 = weak and world >= int(os.environ.get("PF_BENCH_EXTRAS_MIN_GPUS", "8")) and os.environ.get("PF_BENCH_EXTRAS", "1") != "0"
    c4_mem = (True, 0, None)
    if want_configs4:
        c4_mem = host_memory_fits(CONFIGS["synth10m_nlist16384"], world)
        c4_mem = comm.broadcast_object(c4_mem) if world > 1 else c4_mem      # rank 0 decides for everybody
    if want_configs4 and not time_for("the configs[4] record", 300):
        state["configs4"] = {"skipped": "not enough of the run's time allowance left (PF_BENCH_RUN_LIMIT_S)"}
    elif want_configs4 and not c4_mem[0]:
        state["configs4"] = {"skipped": f"host memory: building the 10M-vector data set on {world} ranks needs about {c4_mem[1] / 2**30:.0f} GiB, "
                                        f"{(c4_mem[2] or 0) / 2**30:.0f} GiB available on this host"}
        log(f"[rank {rank}] skipping the configs[4] record: {state['configs4']['skipped']}")
    elif want_configs4:
        if rank == 0:
            guard.publish(make_line(rec))
        xname = "synth10m_nlist16384"
        xcfg = dict(CONFIGS[xname])
        xsteps = max(3, min(args.steps, 6))
        xgrid = default_strong_grid(world)
        guard.enter("configs[4] record (10M x 128, 256-query batches)", 480)
        try:
            rX = run_workload(args, xname, xcfg, comm, False, xgrid, xsteps, args.warmup, want_e2e=not args.no_e2e, tag="configs4 ",
                              guard=guard)
            if rank == 0:
                state["configs4"] = {
                    "config": bench_config(xname, xcfg, rX["L"], rX["Lr"], xcfg["nq"], world, False, grid=xgrid),
                    "scaling": "strong (fixed 10M index)", "n_gpus": world, "value": rX["value"], "ms_per_step": rX["ms_per_step"],
                    "steps": xsteps, "queries_per_s": rX["queries_per_s"], "slot_distances_per_s": rX["slot_distances_per_s"],
                    "result_cts_per_step": rX["result_cts_per_step"], "phases_ms_per_step": rX["phases_ms_per_step"],
                    "roofline": rX["roofline"], "db_gib_per_rank": rX["db_gib_per_rank"], "gather_verified": rX["gather_verified"],
                    "e2e": rX.get("e2e"),
                    "one_gpu_reference": "profiles/r2_bench_synth10m_n8192_1gpu.json: 25.4 ms/step, 394 M distances/s on one B200"}
        except Exception as ex:    # noqa: BLE001
            guard.abort("configs[4] record", repr(ex))
        guard.leave()

    guard.finish()
    if rank == 0:
        sampler.stop()
        emit(make_line(rec))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def host_memory_available():
    """bytes of host memory this job can still take: MemAvailable, capped by the cgroup limit of the container"""
    avail = None
    try:
        for ln in Path("/proc/meminfo").read_text().splitlines():
            if ln.startswith("MemAvailable:"):
                avail = int(ln.split()[1]) * 1024
    except (OSError, ValueError):
        pass
    for lim_f, use_f in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                         ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            lim = Path(lim_f).read_text().strip()
            if lim != "max" and int(lim) < (1 << 60):
                room = int(lim) - int(Path(use_f).read_text().strip())
                avail = room if avail is None else min(avail, room)
        except (OSError, ValueError):
            pass
    return avail


def host_memory_fits(cfg, world):
    """every rank of a fixed-index run builds the WHOLE synthetic data set on the host before it keeps its shard
    (make_dataset): nb x d float32 + ids + slack per rank.  A job that the kernel's OOM killer ends loses its line."""
    need = world * (int(cfg["nb"]) * (cfg["d"] * 4 + 8) * 1.25 + (1 << 30)) + (8 << 30)
    avail = host_memory_available()
    return (avail is None or avail > need), need, avail


def default_strong_grid(world):
    """list shards x query groups for BASELINE configs[2].  Two list shards (the lists stay sharded, every rank
    holds half of the NTT-domain DB) x N/2 query groups: the per-query work (rotations, input NTTs: 0.40 of the
    4.47 ms 1-GPU step, replicated on every rank of an N x 1 grid — 2.53 ms at 2 GPUs = 0.88) is divided by the
    query groups.  Derived from the measured 1- and 2-GPU phase times (profiles/README.md); the 8-GPU A/B run
    of round 2 died on a bench bug before it produced numbers, so `--grid` / PF_BENCH_STRONG_GRID remain."""
    return {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (2, 4)}.get(world, (world, 1))


if __name__ == "__main__":
    sys.exit(main())
