"""oracle/_ref — the REFERENCE's own plaintext functions, compiled from /root/reference where they lie (recipe:
oracle/ref_build/Makefile) and loaded here.  TEST INFRASTRUCTURE: used to pin the oracle's restatements (and, through
committed fixtures, the CUDA path) to what the reference itself computes.  Only tests/, tests/golden/make_golden_ref.py
and bench.py's CPU legs may import this; the product never does.  /root/reference does not exist on the GPU box: there
the prebuilt oracle/_ref/*.so (git-ignored, not gpurun-ignored) is used if present, and the committed fixtures always.

What it covers (shapes are the reference's compile-time constants, include/common/client_server_utils.h:10-20:
d = 128, NQUERY = 5, NPROBE = 20, COARSE_PROBE = 200, K = 100):
  sort_nearest_centroids            src/client/client_lib.cpp:49-81
  compute_nearest_coarse_vectors    src/client/client_lib.cpp:122-156
  compute_nearest_precise_vectors   src/client/client_lib.cpp:189-209
  benchmark_results                 src/client/client_lib.cpp:243-337   (numbers read back from its log lines)
  Server::init_index + preciseSearch  src/server/server_lib.cpp:55-99, 140-167  (FAISS stand-in: train / add are no-ops)
Not covered, because its source is absent: Server::coarseSearch's m_Index->search_encrypted (FAISS fork), and all HE."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
REF = Path(os.environ.get("PF_REFERENCE_DIR", "/root/reference"))
D, NPROBE, COARSE_PROBE, K, NBASE, NQUERY, NLIST = 128, 20, 200, 100, 10000, 5, 256


def available() -> bool:
    return (OUT / "libpf_ref_client.so").exists() and (OUT / "libpf_ref_server.so").exists()


def buildable() -> bool:
    return (REF / "src" / "client" / "client_lib.cpp").exists()


def build(force: bool = False) -> bool:
    """compile from the reference's sources when they are present; returns whether the libraries exist afterwards"""
    if buildable() and (force or not available()):
        r = subprocess.run(["make", "-C", str(HERE / "ref_build"), f"REF={REF}"] + (["-B"] if force else []), capture_output=True, text=True)
        if r.returncode:     # test infrastructure: a failure here must not fail the product build
            import sys
            sys.stderr.write("oracle/_ref could not be built from the reference sources:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return available()


_libs = {}


def _lib(which: str) -> C.CDLL:
    if which not in _libs:
        _libs[which] = C.CDLL(str(OUT / f"libpf_ref_{which}.so"))
    return _libs[which]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def constants():
    out = np.zeros(8, dtype=np.int64)
    _lib("client").ref_constants(_p(out, C.c_int64))
    return dict(zip(["d", "nprobe", "coarse_probe", "k", "nbase", "nquery", "nlist", "sub_quantizers"], out.tolist()))


def sort_nearest_centroids(query: np.ndarray, centroids: np.ndarray):
    q = np.ascontiguousarray(query, dtype=np.float32).reshape(NQUERY, D)
    c = np.ascontiguousarray(centroids, dtype=np.float32).reshape(-1, D)
    idx = np.zeros((NQUERY, len(c)), dtype=np.int64)
    dist = np.zeros((NQUERY, len(c)), dtype=np.float32)
    _lib("client").ref_sort_nearest_centroids(_p(q, C.c_float), _p(c, C.c_float), C.c_int64(len(c)), _p(idx, C.c_int64), _p(dist, C.c_float))
    return idx, dist


def compute_nearest_coarse_vectors(scores, indexes, list_sizes):
    s = np.ascontiguousarray(scores, dtype=np.float32)
    ix = np.ascontiguousarray(indexes, dtype=np.int64)
    ls = np.ascontiguousarray(list_sizes, dtype=np.uint64).reshape(NQUERY)
    idx, dist = np.zeros_like(ix), np.zeros_like(s)
    rc = _lib("client").ref_compute_nearest_coarse_vectors(_p(s, C.c_float), _p(ix, C.c_int64), _p(ls, C.c_uint64), _p(idx, C.c_int64), _p(dist, C.c_float))
    if rc:
        raise RuntimeError("the reference threw (fewer than COARSE_PROBE candidates)")
    return idx, dist


def compute_nearest_precise_vectors(precise_scores, coarse_ids):
    ps = np.ascontiguousarray(precise_scores, dtype=np.float32).reshape(NQUERY, COARSE_PROBE)
    ci = np.ascontiguousarray(coarse_ids, dtype=np.int64).reshape(NQUERY, COARSE_PROBE)
    idx, dist = np.zeros_like(ci), np.zeros_like(ps)
    _lib("client").ref_compute_nearest_precise_vectors(_p(ps, C.c_float), _p(ci, C.c_int64), _p(idx, C.c_int64), _p(dist, C.c_float))
    return idx, dist


def _vecs_write(path: Path, a: np.ndarray):
    """.fvecs / .ivecs: every row = int32 dimension + the row (ref: client_server_utils.h:24-56 reads this)"""
    a = np.ascontiguousarray(a)
    n, d = a.shape
    out = np.empty((n, d + 1), dtype=np.int32)
    out[:, 0] = d
    out[:, 1:] = a.view(np.int32) if a.dtype == np.float32 else a.astype(np.int32)
    out.tofile(path)


class _Tree:
    """the directory layout the reference's relative paths expect: <tmp>/sift/siftsmall/*.{f,i}vecs, cwd = <tmp>/build"""

    def __enter__(self):
        self.tmp = tempfile.TemporaryDirectory()
        root = Path(self.tmp.name)
        (root / "sift" / "siftsmall").mkdir(parents=True)
        (root / "build").mkdir()
        self.data, self.old = root / "sift" / "siftsmall", os.getcwd()
        os.chdir(root / "build")
        return self

    def __exit__(self, *exc):
        os.chdir(self.old)
        self.tmp.cleanup()


def benchmark_results(observed: np.ndarray, groundtruth: np.ndarray):
    """-> {'recall': (r1, r10, r100), 'mrr': (m1, m10, m100)} as the reference logs them"""
    obs = np.ascontiguousarray(observed, dtype=np.int64).reshape(NQUERY, K)
    with _Tree() as t:
        _vecs_write(t.data / "siftsmall_groundtruth.ivecs", np.ascontiguousarray(groundtruth, dtype=np.int32))
        need = _lib("client").ref_benchmark_results(_p(obs, C.c_int64), None, C.c_size_t(0))
        buf = C.create_string_buffer(int(need) + 16)
        _lib("client").ref_benchmark_results.restype = C.c_size_t
        _lib("client").ref_benchmark_results(_p(obs, C.c_int64), buf, C.c_size_t(len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        parts = [x.strip() for x in line.split("|")]
        if parts[0].startswith("Recall@1 ="):
            out["recall"] = tuple(float(x) for x in parts[1:4])
        if parts[0].startswith("MRR@1 ="):
            out["mrr"] = tuple(float(x) for x in parts[1:4])
    return out


def precise_search(query: np.ndarray, ids: np.ndarray, base: np.ndarray) -> np.ndarray:
    """Server::init_index (loads `base` through the reference's own vecs_read) + Server::preciseSearch"""
    q = np.ascontiguousarray(query, dtype=np.float32).reshape(NQUERY, D)
    ix = np.ascontiguousarray(ids, dtype=np.int64).reshape(NQUERY, COARSE_PROBE)
    b = np.ascontiguousarray(base, dtype=np.float32).reshape(-1, D)
    out = np.zeros((NQUERY, COARSE_PROBE), dtype=np.float32)
    with _Tree() as t:
        _vecs_write(t.data / "siftsmall_base.fvecs", b)
        _vecs_write(t.data / "siftsmall_learn.fvecs", b[:16])
        rc = _lib("server").ref_precise_search(_p(q, C.c_float), _p(ix, C.c_int64), _p(out, C.c_float))
    if rc:
        raise RuntimeError("the reference threw")
    return out


def time_reference_functions(query, centroids, ids, base, reps: int = 200):
    """seconds per call of the reference's own sort_nearest_centroids (NQUERY queries x len(centroids)) and
    Server::preciseSearch (NQUERY x COARSE_PROBE exact distances), single thread, as compiled here with -O2"""
    q = np.ascontiguousarray(query, dtype=np.float32).reshape(NQUERY, D)
    c = np.ascontiguousarray(centroids, dtype=np.float32).reshape(-1, D)
    ix = np.ascontiguousarray(ids, dtype=np.int64).reshape(NQUERY, COARSE_PROBE)
    b = np.ascontiguousarray(base, dtype=np.float32).reshape(-1, D)
    f1 = _lib("client").ref_sort_nearest_centroids_timed
    f1.restype = C.c_double
    t_sort = f1(_p(q, C.c_float), _p(c, C.c_float), C.c_int64(len(c)), C.c_int(reps))
    f2 = _lib("server").ref_precise_search_timed
    f2.restype = C.c_double
    with _Tree() as t:
        _vecs_write(t.data / "siftsmall_base.fvecs", b)
        _vecs_write(t.data / "siftsmall_learn.fvecs", b[:16])
        t_precise = f2(_p(q, C.c_float), _p(ix, C.c_int64), C.c_int(reps))
    if t_precise < 0:
        raise RuntimeError("the reference threw")
    return {"sort_nearest_centroids_s": t_sort, "precise_search_s": t_precise, "queries": NQUERY, "centroids": len(c),
            "coarse_probe": COARSE_PROBE, "reps": reps}
