#!/usr/bin/env python3
"""bench.py — encrypted candidate distances/sec of the PreFHEtch server-side search hot path.

One "step" = one batch of encrypted queries through the whole hot path on synthetic SIFT-shaped
data: stage 1 (plaintext coarse quantization, top-nprobe) + stage 2 (rotated query sets, ct x pt
multiply-accumulate over every candidate block of the probed lists, add of the norms, inverse NTT).
  value : whole-job useful candidate distances/s with query ciphertexts already resident in HBM
  e2e   : the same metric through the C-ABI call with HOST buffers (SEAL-serialized query
          ciphertexts in pinned memory -> SEAL-serialized result ciphertexts in pinned memory)
  roofline : the ct x pt MAC kernel against the measured HBM copy bandwidth
  cpu_baseline : the CPU oracle (port of the same op sequence) on a bounded sample, all host cores
`--impl reference` times that CPU path alone (the reference's own server cannot be built here:
FAISS fork / SEAL / Drogon are network FetchContent dependencies).
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
    # BASELINE.json configs[2]: lists sharded across 2/4/8 GPUs
    "sift1m_nlist4096_nprobe64": dict(nb=1_000_000, d=128, nlist=4096, nprobe=64, n=8192, g=32, m=1, tbits=24, nq=64),
    # small smoke configuration (configs[0] shape)
    "siftsmall_nlist100_nprobe8": dict(nb=10_000, d=128, nlist=100, nprobe=8, n=8192, g=8, m=1, tbits=24, nq=16),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------
# synthetic SIFT-shaped data: uint8-valued vectors from a Gaussian mixture
# ----------------------------------------------------------------------------------------------
def make_dataset(cfg, device, seed=1234):
    """Index build (out of the timed path; the reference does it once in Server::init_index).
    torch is used only as plumbing for the k-means-style assignment."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    nb, d, nlist = cfg["nb"], cfg["d"], cfg["nlist"]
    centres = torch.rand((nlist, d), generator=g, device=device) * 160.0
    assign = torch.randint(0, nlist, (nb,), generator=g, device=device)
    base = torch.clamp(torch.round(centres[assign] + torch.randn((nb, d), generator=g, device=device) * 24.0), 0, 255)
    # IVF centroids = mean of the assigned vectors of each true cluster (one Lloyd step from the truth)
    cent = torch.zeros((nlist, d), device=device).index_add_(0, assign, base)
    cnt = torch.bincount(assign, minlength=nlist).clamp(min=1).unsqueeze(1)
    cent = cent / cnt
    # assign every vector to its nearest centroid (what faiss::IndexIVF::add does)
    lab = torch.empty(nb, dtype=torch.long, device=device)
    c2 = (cent * cent).sum(1)
    for s in range(0, nb, 65536):
        x = base[s:s + 65536]
        dist = c2[None, :] - 2.0 * x @ cent.T
        lab[s:s + 65536] = dist.argmin(1)
    order = torch.argsort(lab, stable=True)
    counts = torch.bincount(lab, minlength=nlist)
    offsets = torch.zeros(nlist + 1, dtype=torch.long, device=device)
    offsets[1:] = torch.cumsum(counts, 0)
    vecs = base[order].contiguous()
    # query pool from the same mixture
    npool = 4096
    qa = torch.randint(0, nlist, (npool,), generator=g, device=device)
    queries = torch.clamp(torch.round(centres[qa] + torch.randn((npool, d), generator=g, device=device) * 24.0), 0, 255)
    return dict(centroids=cent.cpu().numpy().astype(np.float32), offsets=offsets.cpu().numpy().astype(np.int64),
                ids=order.cpu().numpy().astype(np.int64), vectors=vecs.cpu().numpy().astype(np.float32),
                queries=queries.cpu().numpy().astype(np.float32))


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


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


class NvmlSampler:
    """Opt-in alternative to ClockSampler (PF_BENCH_CLOCKS=nvml): the same quantities read in-process
    through NVML (what nvidia-smi itself calls) every 100 ms, without a second process."""

    def __init__(self, torch_device):
        self.dev, self.rows, self.ok, self.stop_flag = torch_device, [], False, threading.Event()

    def _handle(self, nv):
        import torch
        idx = torch.device(self.dev).index or 0
        try:   # containers usually expose exactly the assigned GPUs to NVML: same numbering as CUDA
            if nv.nvmlDeviceGetCount() == torch.cuda.device_count():
                return nv.nvmlDeviceGetHandleByIndex(idx)
        except Exception:
            pass
        cands = []
        try:
            cands.append("GPU-" + str(torch.cuda.get_device_properties(self.dev).uuid))
        except Exception:
            pass
        ents = [x.strip() for x in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip()]
        ent = ents[idx] if idx < len(ents) else ""
        if ent.startswith(("GPU-", "MIG-")):
            cands.append(ent)
        for u in cands:
            for arg in (u, u.encode()):
                try:
                    return nv.nvmlDeviceGetHandleByUUID(arg)
                except Exception:
                    continue
        return nv.nvmlDeviceGetHandleByIndex(int(ent) if ent.isdigit() else idx)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = self._handle(pynvml)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
        except Exception:
            self.ok = False
        return self.ok

    def _sample(self):
        nv = self.nv
        try:
            self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                              int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))))
        except Exception:
            pass

    def _poll(self):
        while not self.stop_flag.is_set():
            self._sample()
            self.stop_flag.wait(0.1)

    def stop(self):
        self.stop_flag.set()
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.th.join(timeout=2)
        self._sample()   # one more while the queued steps still run (covers very short timed regions)
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(k for k, bit in names.items() if any(r[1] & bit for r in self.rows))
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "source": "nvml"}


def make_sampler(local_rank, torch_device):
    """PF_BENCH_CLOCKS = smi (default: the nvidia-smi poller every number in profiles/ was taken with) | nvml | off"""
    mode = os.environ.get("PF_BENCH_CLOCKS", "smi")
    if mode == "nvml":
        return NvmlSampler(torch_device)
    s = ClockSampler(local_rank)
    if mode == "off":
        s.start = lambda: None
    return s


# ----------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle's OpenMP whole-step driver on a bounded sample
# ----------------------------------------------------------------------------------------------
def cpu_pipeline(cfg, data, nq_sample, nthreads, steps=1, warmup=0, seed=5):
    """Runs the CPU port of the step on `nq_sample` queries of the workload.  Only this function and
    --impl reference execute oracle/ (as the baseline, never as the product)."""
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
        out, (rot_s, mac_s) = O.search_pairs(ctx, lay, cts, keys, False, pair_q, pair_b, diag, norm, nthreads)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append((dt, rot_s, mac_s))
    dt = float(np.mean([x[0] for x in times]))
    return dict(useful=useful, slots=len(pair_q) * lay.C, seconds=dt, rot_s=float(np.mean([x[1] for x in times])),
                mac_s=float(np.mean([x[2] for x in times])), encode_s=enc_s, pairs=len(pair_q), nq=nq_sample,
                distinct_blocks=len(boff))


def run_reference(args, cfg, cfg_name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import pf_oracle as O
    O.build()
    data = make_dataset(cfg, "cpu") if cfg["nb"] <= 200_000 else make_dataset_cpu_light(cfg)
    nthreads = O.max_threads()
    nq_sample = min(cfg["nq"], 8)     # bounded sample: ~0.1-0.2 s per step on 16 cores, so K steps stay within minutes
    r = cpu_pipeline(cfg, data, nq_sample, nthreads, steps=args.steps, warmup=args.warmup)
    val = r["useful"] / r["seconds"]
    line = {
        "metric": "encrypted candidate distances/sec", "value": val, "unit": "distances/s", "impl": "reference",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        # same keys as the GPU arm's config; the CPU arm times a bounded sample of the same workload
        "config": {"workload": cfg_name, "nb": cfg["nb"], "d": cfg["d"], "nlist": cfg["nlist"], "nprobe": cfg["nprobe"],
                   "poly_degree": cfg["n"], "limbs": len(O.BFV_DEFAULT_PRIMES[cfg["n"]]) - 1,
                   "result_limbs": len(O.BFV_DEFAULT_PRIMES[cfg["n"]]) - 1, "g": cfg["g"], "query_cts": cfg["m"],
                   "queries_per_step": nq_sample, "parallelism": f"{nthreads} host threads (OpenMP)",
                   "l2_policy": "n/a (CPU)"},
        "cpu_baseline": {"value": val, "unit": "distances/s", "cores": nthreads, "kind": "port",
                         "sample": f"{nq_sample} queries x {r['pairs']} (query,block) pairs per step, whole hot path "
                                   f"(rotations {r['rot_s']:.2f}s + MAC/INTT {r['mac_s']:.2f}s)"},
        "e2e": {"value": val, "unit": "distances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "queries_per_s": nq_sample / r["seconds"],
    }
    emit(line)
    return 0


def make_dataset_cpu_light(cfg, seed=1234):
    """CPU-only dataset for the reference arm when no GPU plumbing is wanted: same generator family, numpy."""
    rng = np.random.default_rng(seed)
    nb, d, nlist = cfg["nb"], cfg["d"], cfg["nlist"]
    centres = rng.uniform(0, 160, size=(nlist, d)).astype(np.float32)
    assign = rng.integers(0, nlist, size=nb)
    base = np.clip(np.rint(centres[assign] + rng.normal(0, 24, size=(nb, d)).astype(np.float32)), 0, 255).astype(np.float32)
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


def recall_at_10(eng, data, nprobe, dev, nq_r=64):
    """recall@10 of the two-stage search against exact brute force over the whole base set (untimed, world 1).
    Stage 2 here is the engine's plaintext path: the encrypted results decrypt to exactly these distances
    (tests/test_gpu_parity.py::test_encrypted_search_end_to_end), so both pipelines have this recall.
    Ties are broken by the lower id on both sides."""
    import torch
    x = np.ascontiguousarray(data["queries"][:nq_r], dtype=np.float32)
    idx = eng.coarse_quantize(x, nprobe)
    dist, labels, sizes = eng.coarseSearch(x, idx)
    base = torch.from_numpy(data["vectors"]).to(dev)
    ids = torch.from_numpy(data["ids"]).to(dev)
    xq = torch.from_numpy(x).to(dev).double()
    q2 = (xq * xq).sum(1)
    best = None
    for s0 in range(0, base.shape[0], 131072):      # exact integer distances in float64, chunked
        b = base[s0:s0 + 131072].double()
        d2 = q2[:, None] + (b * b).sum(1)[None, :] - 2.0 * xq @ b.T
        key = d2.round().long() * (1 << 21) + ids[s0:s0 + 131072][None, :]
        key = key if best is None else torch.cat([best, key], dim=1)
        best = key.topk(10, dim=1, largest=False).values
    gt = (best % (1 << 21)).cpu().numpy()
    hits, off = 0, 0
    for i in range(len(x)):
        n = int(sizes[i])
        k = dist[off:off + n].astype(np.int64) * (1 << 21) + labels[off:off + n]
        found = np.sort(k)[:10] % (1 << 21)
        hits += len(set(found.tolist()) & set(gt[i].tolist()))
        off += n
    return hits / (10.0 * len(x))


# ----------------------------------------------------------------------------------------------
_JSON_OUT = None


def emit(line):
    """the one JSON line goes to the process's original stdout"""
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)


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
    ap.add_argument("--result-limbs", type=int, default=1,
                    help="limbs of the result ciphertexts (SEAL mod_switch_to before save); 0 = no switching")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    # stdout must carry exactly one JSON line: native libraries (NCCL prints its version banner with
    # printf) write to fd 1, so fd 1 is pointed at stderr for the whole run and the line goes to a dup
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args.warmup = max(args.warmup, 0)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg_name = args.config or "sift1m_nlist1024_nprobe16"
    cfg = dict(CONFIGS[cfg_name])
    # N > 1 default: WEAK scaling of configs[1] — every rank holds its own 1M-vector / 1024-list shard
    # of a world x larger index and the query probes 16*world lists (16 per shard on average).
    # `--config sift1m_nlist4096_nprobe64` runs BASELINE configs[2] instead (fixed 1M index, strong).
    weak = world > 1 and args.config is None
    if args.nq:
        cfg["nq"] = args.nq
    if args.g:
        cfg["g"] = args.g

    if args.impl == "reference":
        return run_reference(args, cfg, cfg_name)

    import torch
    import torch.distributed as dist
    import prefhetch_b200 as pf
    from prefhetch_b200.build import build as build_lib
    if rank == 0:
        build_lib()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")  # the per-step flag all-reduce must not queue behind compute
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()

    n, d, g, m, nprobe, nq = cfg["n"], cfg["d"], cfg["g"], cfg["m"], cfg["nprobe"], cfg["nq"]
    t_setup = time.perf_counter()
    if weak:
        nprobe = nprobe * world
        data = make_dataset_shard(cfg, dev, rank, world)
        cent, counts = data["centroids"], data["counts"]
        dist.all_reduce(cent)                       # every list is non-zero on exactly one rank
        dist.all_reduce(counts)
        data["centroids"] = cent.cpu().numpy().astype(np.float32)
        global_list_sizes = counts.cpu().numpy().astype(np.int64)
    else:
        data = make_dataset(cfg, dev)
        global_list_sizes = (data["offsets"][1:] - data["offsets"][:-1]).astype(np.int64)
    eng = pf.Engine(d, n, pf.bfv_default_primes(n), pf.batching_plain_modulus(n, cfg["tbits"]), m, g,
                    device=local_rank, rank=rank, world=world, result_limbs=args.result_limbs)
    info = eng.load_index(data["centroids"], data["offsets"], data["ids"], data["vectors"])
    eng.set_list_sizes(data["offsets"])
    L, k, K, C_ = eng.L, eng.k, info["K"], info["C"]
    ctw = eng.ctw
    stream = torch.cuda.Stream(device=dev)
    eng.set_stream(stream.cuda_stream)
    # synthetic Galois keys / query ciphertexts: uniform residues (timing does not depend on values;
    # parity with real encryptions is what tests/ and smoke() check)
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
    npool_steps = 4
    ct_pool = [rand_residues((nq, m, 2), range(L)) for _ in range(npool_steps)]  # [nq][m][2][L][n]
    queries = data["queries"]
    nsteps_total = args.warmup + args.steps
    qsets = [queries[(s * nq) % (len(queries) - nq):][:nq] for s in range(nsteps_total)]
    if world == 1:
        max_res = int(eng._blocks_per_list.max()) * nprobe * nq
    else:  # probes spread over the ranks: 3x the expected per-rank share is far beyond its fluctuation
        mean_blocks = float(eng._blocks_per_list[eng._blocks_per_list > 0].mean())
        max_res = int(3.0 * nq * nprobe / world * mean_blocks) + 256
    NBUF = 2 if world > 1 else 1
    d_outs = [torch.empty((max_res, 2, eng.Lr, n), dtype=torch.int64, device=dev) for _ in range(NBUF)]
    log(f"[rank {rank}] setup {time.perf_counter() - t_setup:.1f}s  index: {info}  max_res {max_res}")
    # result ciphertexts every rank produces for a probe list (all ranks know all list sizes)
    blocks_of_list = (global_list_sizes + C_ - 1) // C_

    def counts_per_rank(idx):
        flat = idx.reshape(-1)
        return [int(blocks_of_list[flat[(flat % world) == r]].sum()) for r in range(world)]

    # ---- multi-GPU gather: every rank's result ciphertexts go into RANK 0's HBM through a peer-mapped
    # buffer (CUDA IPC over NVLink).  Default: copy-engine DMA on a side stream, overlapped with the next
    # step (rank 0's NVLink ingest, ~0.75 TB/s, is the shared resource: 7 writers bursting from inside
    # their last kernel stall each other, measured 68 % efficiency at 8 GPUs).  PF_BENCH_FUSED_GATHER=1
    # lets the last kernel store straight into the peer buffer instead.  Stream-ordered arrival / ack
    # flags (one-thread kernels on peer memory) tell rank 0 that a step has landed and the shards that
    # their buffer is free.  NCCL only carries the set-up exchange and the timing reduce.
    comm_stream = torch.cuda.Stream(device=dev, priority=-1) if world > 1 else None
    out_ptrs = [t.data_ptr() for t in d_outs]
    res_bytes = 2 * eng.Lr * n * 8
    ipc_local, ipc_mapped = [], []
    fused_gather = bool(os.environ.get("PF_BENCH_FUSED_GATHER"))
    copied_ev = [None] * NBUF
    if world > 1:
        handles = [None]
        if rank == 0:
            handles = [[None] * world for _ in range(NBUF)]
            for b in range(NBUF):
                for r in range(1, world):
                    ptr, h = eng.ipc_alloc(max_res * res_bytes)
                    ipc_local.append(ptr)
                    handles[b][r] = h
            handles = [handles]
        dist.broadcast_object_list(handles, src=0)
        if rank != 0:
            for b in range(NBUF):
                p_ = eng.ipc_open(handles[0][b][rank])
                ipc_mapped.append(p_)
                if fused_gather:
                    out_ptrs[b] = p_        # this rank's results go straight into rank 0's gather buffer
        # arrival flags live on rank 0 (one 128-byte line per rank), ack flags on every rank: one-thread
        # kernels write / wait on them in stream order — no NCCL collective on the data path
        FL = 128
        arr_ptr, arr_h = eng.ipc_alloc(FL * world) if rank == 0 else (None, None)
        ack_ptr, ack_h = eng.ipc_alloc(FL)
        torch.cuda.synchronize()
        allh = [None] * world
        dist.all_gather_object(allh, (arr_h, ack_h))
        if rank == 0:
            ipc_local += [arr_ptr, ack_ptr]
            ack_peer = [None] + [eng.ipc_open(allh[r][1]) for r in range(1, world)]
            ipc_mapped += ack_peer[1:]
        else:
            ipc_local.append(ack_ptr)
            arr_ptr = eng.ipc_open(allh[0][0])
            ipc_mapped.append(arr_ptr)
        # zero the flags with the flag kernels (raw IPC pointers, no torch view)
        if rank == 0:
            for r in range(world):
                eng.flag_write(arr_ptr + FL * r, 0)
        eng.flag_write(ack_ptr, 0)
        eng.synchronize()
        dist.barrier()

    next_idx = {}
    host_t = {"wait": 0.0, "search": 0.0, "coarse": 0.0, "gather": 0.0, "n": 0}

    def step(s):
        """stage 1 of batch s+1 is issued right after stage 2 of batch s was enqueued, so the host-side
        planning of the next batch overlaps the GPU work of this one (a serving loop does the same)"""
        b = s % NBUF
        t0 = time.perf_counter()
        if world > 1 and rank != 0 and s >= NBUF:
            if fused_gather:                    # rank 0 has acknowledged the step that last used this buffer
                eng.flag_wait(ack_ptr, s - NBUF + 1)
            elif copied_ev[b] is not None:      # the DMA that last read this local buffer is done
                stream.wait_event(copied_ev[b])
        idx = next_idx.pop(s) if s in next_idx else eng.coarse_quantize(qsets[s], nprobe)
        t1 = time.perf_counter()
        rpq, st = eng.search_device(ct_pool[s % npool_steps].data_ptr(), nq, idx, out_ptrs[b], max_res)
        t2 = time.perf_counter()
        # always one stage-1 call per step (the last one quantizes a batch that is never searched)
        if not os.environ.get("PF_BENCH_NO_PREFETCH"):
            next_idx[s + 1] = eng.coarse_quantize(qsets[(s + 1) % nsteps_total], nprobe)
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
        if rank != 0 and fused_gather:
            eng.flag_write(arr_ptr + FL * rank, s + 1)      # after this rank's last kernel of step s
        elif rank != 0:
            b = s % NBUF
            done = torch.cuda.Event()
            done.record(stream)
            comm_stream.wait_event(done)
            cs = comm_stream.cuda_stream
            if s >= NBUF:
                eng.flag_wait(ack_ptr, s - NBUF + 1, cs)     # rank 0 is done with the peer buffer
            eng.copy_async(ipc_mapped[b], d_outs[b].data_ptr(), int(st["nresults"]) * res_bytes, cs)
            eng.flag_write(arr_ptr + FL * rank, s + 1, cs)
            ev = torch.cuda.Event()
            ev.record(comm_stream)
            copied_ev[b] = ev
        else:
            done = torch.cuda.Event()
            done.record(stream)
            comm_stream.wait_event(done)
            cs = comm_stream.cuda_stream
            for r in range(1, world):
                eng.flag_wait(arr_ptr + FL * r, s + 1, cs)   # every shard's ciphertexts of step s are in HBM
            for r in range(1, world):
                eng.flag_write(ack_peer[r], s + 1, cs)       # ... the response can be assembled; buffer free
        host_t["gather"] += time.perf_counter() - tg

    # ---- value: device-resident timed region --------------------------------------------------
    with torch.cuda.stream(stream):
        for s in range(args.warmup):
            idx, st = step(s)
            gather_results(s, idx, st)
        eng.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        eng.timing_enable(True)
        eng.timing_read(reset=True)
        sampler = make_sampler(local_rank, dev)
        if rank == 0:                          # one poller per job: NVML queries take driver locks
            if sampler.start() is False:       # NVML requested but unusable: fall back to nvidia-smi
                sampler = ClockSampler(local_rank)
                sampler.start()
        launches0 = eng.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        useful = slots = nres = 0
        blocks_distinct = pairs = 0
        ev0.record(stream)
        for s in range(args.warmup, nsteps_total):
            idx, st = step(s)
            gather_results(s, idx, st)
            useful += st["useful_distances"]
            slots += st["slot_distances"]
            nres += st["nresults"]
        if comm_stream is not None:             # the timed region ends when the last gather has landed
            stream.wait_stream(comm_stream)
        ev1.record(stream)
        clocks = sampler.stop() if rank == 0 else None   # last sample while the queued steps still run
        eng.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms_total = ev0.elapsed_time(ev1)
        launches = eng.launch_count() - launches0
        phases = eng.timing_read(reset=True)
        eng.timing_enable(False)
        log(f"[rank {rank}] host ms/step: " + ", ".join(f"{k} {1e3 * v / max(1, host_t['n']):.3f}" for k, v in host_t.items() if k != "n"))
        log(f"[rank {rank}] {ms_total / args.steps:.3f} ms/step; phases " + ", ".join(f"{k} {v['ms'] / args.steps:.3f}" for k, v in phases.items()))

    # distinct blocks per step for the algorithmic-bytes formula (host-side bookkeeping, untimed)
    bpl = eng._blocks_per_list
    for s in range(args.warmup, nsteps_total):
        idx = eng.coarse_quantize(qsets[s], nprobe)
        own = idx[(idx % world) == rank] if world > 1 else idx.reshape(-1)
        pairs += int(bpl[own].sum())
        blocks_distinct += int(bpl[np.unique(own)].sum())

    t_max = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    tot = torch.tensor([useful, slots, nres, launches, pairs, blocks_distinct], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total = float(t_max.item())
    useful_all, slots_all, nres_all, launches_all = (float(x) for x in tot[:4].tolist())
    ms_step = ms_total / args.steps
    value = useful_all / (ms_total * 1e-3)

    # ---- roofline of the MAC kernel (rank-local launch, measured with CUDA events on its stream) --
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    mac_ms = phases["mac"]["ms"] / max(1, phases["mac"]["launches"])
    LN8 = 8.0 * L * n
    alg_bytes = LN8 * (2.0 * K * nq * args.steps + (K + 1.0) * blocks_distinct + 2.0 * pairs) / args.steps  # rank-local
    streamed_bytes = LN8 * (2.0 * K * nq * args.steps + (K + 1.0) * pairs + 2.0 * pairs) / args.steps
    achieved = alg_bytes / (mac_ms * 1e-3) / 1e9 if mac_ms > 0 else 0.0
    traffic, mac_name = None, "mac_kernel_occ" if K <= 16 else "mac_kernel"
    tf = ROOT / "profiles" / "mac_traffic.json"
    if tf.exists() and world == 1:   # dram__bytes_read+write per launch from the committed ncu --set full capture
        tj = json.loads(tf.read_text()).get(cfg_name, {})
        traffic, mac_name = tj.get("traffic_bytes_per_launch"), tj.get("kernel", mac_name)
    roofline = {"bound": "hbm", "kernel": mac_name, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": "measured" if peaks else "fallback",
                "algorithmic_bytes_per_launch": alg_bytes, "streamed_bytes_per_launch": streamed_bytes,
                "streamed_gbs": streamed_bytes / (mac_ms * 1e-3) / 1e9 if mac_ms > 0 else 0.0,
                "ms_per_launch": mac_ms}

    # ---- recall@10 of the search (BASELINE metric, untimed; single GPU: the whole index is local) ----
    recall = None
    if world == 1 and cfg["nb"] < (1 << 21):
        try:
            recall = recall_at_10(eng, data, nprobe, dev)
            log(f"[rank {rank}] recall@10 = {recall:.4f} (nprobe {nprobe} of {cfg['nlist']} lists, 64 queries, exact brute-force ground truth)")
        except Exception as ex:  # the metric line must not depend on this bookkeeping
            log(f"[rank {rank}] recall@10 not computed: {ex}")

    # ---- e2e: host buffers through the public C-ABI call ------------------------------------------
    e2e = None
    if not args.no_e2e:
        ctb = eng.ct_bytes
        hdr = np.frombuffer(eng.ct_serialize(np.zeros((2, L, n), dtype=np.uint64)), dtype=np.uint8)[:ctb - ctw * 8]
        qblob = torch.empty(nq * m * ctb, dtype=torch.uint8).pin_memory()
        qnp = qblob.numpy()
        host_ct = ct_pool[0].cpu().numpy().view(np.uint64).reshape(nq * m, -1)
        for c in range(nq * m):
            qnp[c * ctb: c * ctb + len(hdr)] = hdr
            qnp[c * ctb + len(hdr): (c + 1) * ctb] = host_ct[c].view(np.uint8)
        offs = (np.arange(nq * m + 1, dtype=np.uint64) * ctb)
        out_host = torch.empty(max_res * eng.slot_bytes, dtype=torch.uint8).pin_memory()
        out_np = out_host.numpy()
        e_steps = max(2, min(args.steps, 20))
        e_useful, h2d, d2h = 0, 0, 0
        t_cq = t_se = 0.0
        for s in range(2 + e_steps):
            if s == 2:
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            x = qsets[s % nsteps_total]
            ta = time.perf_counter()
            idx = eng.coarse_quantize(x, nprobe)
            tb = time.perf_counter()
            res = eng.coarseSearchEncrypted(qnp, offs, idx, out=out_np)
            if s >= 2:
                t_cq += tb - ta
                t_se += time.perf_counter() - tb
                e_useful += res.stats["useful_distances"]
                h2d += nq * m * ctb + x.nbytes
                d2h += res.stats["out_bytes"] + idx.nbytes
        torch.cuda.synchronize()
        e_dt = time.perf_counter() - t0
        log(f"[rank {rank}] e2e {e_dt / e_steps * 1e3:.3f} ms/step: coarse_quantize {t_cq / e_steps * 1e3:.3f}, "
            f"coarseSearchEncrypted {t_se / e_steps * 1e3:.3f}")
        et = torch.tensor([e_dt], device=dev, dtype=torch.float64)
        eu = torch.tensor([e_useful], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
            dist.all_reduce(eu, op=dist.ReduceOp.SUM)
        e2e = {"value": float(eu.item()) / float(et.item()), "unit": "distances/s",
               "h2d_bytes_per_step": h2d // e_steps, "d2h_bytes_per_step": d2h // e_steps,
               "ms_per_step": float(et.item()) / e_steps * 1e3, "steps": e_steps}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ---------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import pf_oracle as O
            O.build()
            nthreads = O.max_threads()
            nq_s = min(nq, 32)   # ~10 s of CPU work on the host cores
            r = cpu_pipeline(cfg, data, nq_s, nthreads)
            r1 = cpu_pipeline(cfg, data, 2, 1)
            cpu = {"value": r["useful"] / r["seconds"], "unit": "distances/s", "cores": nthreads, "kind": "port",
                   "sample": f"{nq_s} queries / {r['pairs']} (query,block) pairs of the same workload, whole hot path, "
                             f"{nthreads} OpenMP threads ({r['seconds']:.2f}s: rotations {r['rot_s']:.2f}s, MAC+INTT "
                             f"{r['mac_s']:.2f}s)",
                   "single_thread_value": r1["useful"] / r1["seconds"], "queries_per_s": nq_s / r["seconds"]}
        except Exception as ex:  # the baseline is a report, not the product
            cpu = {"value": None, "error": repr(ex)}

    if rank == 0:
        line = {
            "metric": "encrypted candidate distances/sec", "value": value, "unit": "distances/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak" if (weak or world == 1) else "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": cfg_name + (f" x{world} shards (weak: {world}M vectors, nlist {cfg['nlist'] * world}, "
                                                f"nprobe {nprobe})" if weak else ""),
                       "nb": cfg["nb"] * (world if weak else 1), "d": d,
                       "nlist": cfg["nlist"] * (world if weak else 1), "nprobe": nprobe,
                       "poly_degree": n, "limbs": L, "result_limbs": eng.Lr, "g": g, "query_cts": m, "queries_per_step": nq,
                       "parallelism": f"lists%{world}" if world > 1 else "single",
                       "l2_policy": f"inputs larger than L2: NTT-domain DB {info['db_bytes'] / 2**30:.1f} GiB/rank "
                                    "streamed from HBM, query batches rotate through a pool"},
            "queries_per_s": nq * args.steps / (ms_total * 1e-3),
            "recall_at_10": recall,
            "slot_distances_per_s": slots_all / (ms_total * 1e-3),
            "result_cts_per_step": nres_all / args.steps,
            "gpu_launches": int(launches_all),
            "clocks": clocks, "roofline": roofline,
            "phases_ms_per_step": {kk: v["ms"] / args.steps for kk, v in phases.items()},
            "e2e": e2e, "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        for p_ in ipc_mapped:
            eng.ipc_close(p_)
        dist.barrier()
        for p_ in ipc_local:
            eng.ipc_free(p_)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
