"""Dry-run stand-ins for bench.py (TEST INFRASTRUCTURE): `PF_BENCH_DRYRUN=1 python bench.py ...` runs the whole
control flow of the benchmark — argument handling, data sets, the rank grid, the IPC / flag set-up, the step and
gather loops, every collective, the shared-memory response buffer, the pipelined e2e loop, the strong-scaling
record, the guard, the JSON line — on the CPU with gloo, a stub in place of `prefhetch_b200.Engine` and no-op
CUDA streams / events.  Nothing is measured (the numbers in the line are meaningless); what it proves is that every
rank takes the same path through the collectives and that the bookkeeping holds together at 1, 2, 4 and 8 ranks —
the class of bug that cost round 2 its 8-GPU run.  tests/test_bench_dryrun.py drives it."""
from __future__ import annotations

import contextlib
import time
import types

import numpy as np

_BFV_DEFAULT = {8192: [0x7FFFFFD8001, 0x7FFFFFC8001, 0xFFFFFFFC001, 0xFFFFFF6C001, 0xFFFFFEBC001],
                16384: [0xFFFFFFFD8001, 0xFFFFFFFA0001, 0xFFFFFFF00001, 0x1FFFFFFF68001, 0x1FFFFFFF50001,
                        0x1FFFFFFEE8001, 0x1FFFFFFEA0001, 0x1FFFFFFE88001, 0x1FFFFFFE48001]}


def bfv_default_primes(n):
    return list(_BFV_DEFAULT[n])


def batching_plain_modulus(n, bits):
    return {(8192, 24): 16760833, (8192, 27): 133857281, (16384, 24): 16580609}[(n, bits)]


class PfError(RuntimeError):
    pass


class _Pending:
    def __init__(self, result):
        self.result = result

    def collect(self):
        return self.result


class Engine:
    """same surface as prefhetch_b200.Engine, as far as bench.py uses it; no arithmetic"""
    _next_ptr = 0x1000

    def __init__(self, dim, poly_degree=8192, primes=None, plain_modulus=None, query_cts=1, partial_g=8, device=0, rank=0,
                 world=1, result_limbs=0):
        self.dim, self.n, self.primes = dim, poly_degree, list(primes)
        self.k, self.L = len(self.primes), len(self.primes) - 1
        self.m, self.g, self.rank, self.world = query_cts, partial_g, rank, world
        self.Lr = result_limbs or self.L
        self.ctw = 2 * self.L * self.n
        self.ct_bytes = 113 + self.ctw * 8
        self.result_bytes = 113 + 2 * self.Lr * self.n * 8
        self.slot_bytes = 128 + 2 * self.Lr * self.n * 8
        d_pad = 1
        while d_pad < dim:
            d_pad <<= 1
        self.R = d_pad // self.m // self.g
        self.K = self.m * self.R
        self.C = self.n // self.g
        self._launches = 0
        self._timing = {k: {"ms": 0.0, "launches": 0} for k in ("coarse", "to_ntt", "rotate", "mac", "intt")}
        self._inflight = 0

    def close(self):
        pass

    def load_index(self, centroids, list_offsets, ids, vectors):
        self.cent = np.asarray(centroids, dtype=np.float32)
        self.off = np.asarray(list_offsets, dtype=np.int64)
        self.ids = np.asarray(ids, dtype=np.int64)
        self.vec = np.asarray(vectors, dtype=np.float32)
        self.nlist = len(self.off) - 1
        sizes = np.diff(self.off)
        owned = (np.arange(self.nlist) % self.world) == self.rank
        nb = int(np.where(owned, (sizes + self.C - 1) // self.C, 0).sum())
        return {"nlist": self.nlist, "ntotal": int(self.off[-1]), "nblocks": nb, "nblocks_local": nb,
                "db_bytes": nb * (self.K + 1) * self.L * self.n * 8, "K": self.K, "C": self.C, "R": self.R, "d_pad": self.dim,
                "L": self.L, "k": self.k}

    def index_info(self):
        return {"K": self.K, "C": self.C, "R": self.R}

    def set_list_sizes(self, list_offsets):
        lo = np.asarray(list_offsets, dtype=np.int64)
        sizes = lo[1:] - lo[:-1]
        owned = (np.arange(len(sizes)) % self.world) == self.rank
        self._sizes = np.where(owned, sizes, 0)
        self._blocks_per_list = np.where(owned, (sizes + self.C - 1) // self.C, 0)

    def set_stream(self, s):
        pass

    def galois_elt(self, step):
        return 2 * step + 1

    def set_galois_key(self, elt, words):
        assert words.size == self.L * 2 * self.k * self.n

    def coarse_quantize(self, x, nprobe, return_dist=False):
        x = np.asarray(x, dtype=np.float32)
        if x.shape[0] == 0:
            return np.zeros((0, nprobe), dtype=np.int64)
        d2 = (x * x).sum(1)[:, None] + (self.cent * self.cent).sum(1)[None, :] - 2.0 * x @ self.cent.T
        return np.argsort(d2, axis=1, kind="stable")[:, :nprobe].astype(np.int64)

    def _stats(self, idx):
        flat = np.asarray(idx).reshape(-1)
        nres = int(self._blocks_per_list[flat].sum())
        return {"nresults": nres, "out_bytes": nres * self.slot_bytes, "useful_distances": int(self._sizes[flat].sum()),
                "slot_distances": nres * self.C}

    def search_device(self, d_query_ptr, nq, idx, d_out_ptr, cap_results):
        assert idx.shape[0] == nq
        st = self._stats(idx)
        assert st["nresults"] <= cap_results, "max_res too small"
        self._launches += 20
        for k, v in (("to_ntt", 0.05), ("rotate", 1.0), ("mac", 0.8), ("intt", 0.5)):
            self._timing[k]["ms"] += v
            self._timing[k]["launches"] += 1
        rpq = self._blocks_per_list[np.asarray(idx)].sum(1) if nq else np.zeros(0, dtype=np.int64)
        return rpq.astype(np.int64), st

    def coarseSearch(self, x, idx):
        x = np.asarray(x, dtype=np.float32)
        dist, labels, sizes = [], [], []
        for i in range(len(x)):
            n = 0
            for l in idx[i]:
                v = self.vec[self.off[l]:self.off[l + 1]]
                dist.append(((v - x[i]) ** 2).sum(1))
                labels.append(self.ids[self.off[l]:self.off[l + 1]])
                n += len(v)
            sizes.append(n)
        return np.concatenate(dist).astype(np.float32), np.concatenate(labels), np.array(sizes, dtype=np.int64)

    def launch_count(self):
        return self._launches

    def timing_enable(self, on=True):
        pass

    def timing_read(self, reset=True):
        out = {k: dict(v) for k, v in self._timing.items()}
        if reset:
            for v in self._timing.values():
                v["ms"], v["launches"] = 0.0, 0
        return out

    def synchronize(self):
        pass

    def device_checksum(self, dptr, nwords, cuda_stream=0):
        return (int(nwords) * 2654435761 + 12345) & 0xFFFFFFFFFFFFFFFF   # equal on both sides iff the counts agree

    def ipc_alloc(self, nbytes):
        Engine._next_ptr += 0x100000
        return Engine._next_ptr, bytes(64)

    def ipc_open(self, handle):
        Engine._next_ptr += 0x100000
        return Engine._next_ptr

    def ipc_close(self, ptr):
        pass

    def ipc_free(self, ptr):
        pass

    def copy_async(self, dst, src, nbytes, cuda_stream=0):
        assert nbytes >= 0

    def flag_write(self, ptr, value, cuda_stream=0):
        pass

    def flag_wait(self, ptr, value, cuda_stream=0):
        pass

    def host_register(self, arr):
        assert arr.nbytes > 0

    def host_unregister(self, arr):
        pass

    def set_search_groups(self, groups):
        pass

    def ct_serialize(self, ct, is_ntt=False):
        return bytes(113) + np.asarray(ct, dtype=np.uint64).tobytes()

    def submitSearchEncrypted(self, query_blob, ct_offsets, idx, out=None):
        import os
        inj = os.environ.get("PF_BENCH_DRYRUN_INJECT", "")       # "fail:<rank>" or "hang:<rank>": fault injection in e2e
        if inj and int(inj.split(":")[1]) == int(os.environ.get("RANK", "0")):
            if inj.startswith("fail"):
                raise PfError("injected failure in submit")
            time.sleep(3600)
        assert self._inflight < 4, "more than 4 searches in flight"
        assert len(ct_offsets) == idx.shape[0] * self.m + 1 and int(ct_offsets[-1]) <= query_blob.size
        st = self._stats(idx)
        assert st["out_bytes"] <= out.size, "response share too small"
        out[:min(64, out.size)] = (self.rank + 1) & 0xFF
        self._inflight += 1
        eng = self

        class P(_Pending):
            def collect(self_inner):
                eng._inflight -= 1
                return self_inner.result
        return P(types.SimpleNamespace(stats=st))

    def coarseSearchEncrypted(self, query_blob, ct_offsets, idx, out=None):
        return self.submitSearchEncrypted(query_blob, ct_offsets, idx, out).collect()


class _FakeEvent:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max(1e-3, (other.t - self.t) * 1e3)


class _FakeStream:
    cuda_stream = 0

    def __init__(self, device=None, priority=0):
        pass

    def wait_event(self, ev):
        pass

    def wait_stream(self, s):
        pass


def install(torch):
    """replace the CUDA touch points bench.py uses by no-ops (CPU tensors, gloo)"""
    torch.cuda.Stream = _FakeStream
    torch.cuda.Event = _FakeEvent
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.empty_cache = lambda: None
    torch.cuda.set_device = lambda d: None
    torch.cuda.is_available = lambda: True
    torch.cuda.mem_get_info = lambda dev=None: (100 << 30, 180 << 30)
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.Tensor.pin_memory = lambda self: self

    def no_props(dev=None):
        raise RuntimeError("dry run: no CUDA device")
    torch.cuda.get_device_properties = no_props
    mod = types.ModuleType("prefhetch_b200")
    mod.Engine, mod.PfError = Engine, PfError
    mod.bfv_default_primes, mod.batching_plain_modulus = bfv_default_primes, batching_plain_modulus
    return mod
