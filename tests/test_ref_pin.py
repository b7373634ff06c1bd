"""The plaintext stages pinned to the REFERENCE ITSELF.  tests/golden/ref_plain_v1.json holds outputs of the
reference's own functions — compiled from /root/reference where the sources lie (oracle/ref_build, oracle/pf_ref.py,
generator tests/golden/make_golden_ref.py) — on inputs every test rebuilds from integer formulas.  Checked against
them: the oracle's restatements (bit for bit), the product's C++ client library, and — on the GPU box, in
tests/test_gpu_parity.py::test_plain_stages_against_reference_golden — the CUDA kernels.  When the reference sources
are present (this container) the fixture itself is re-derived and compared, so it cannot drift.

Sorting: the reference sorts with std::ranges::sort, which leaves the order of equal distances open; the product keeps
input order among equals.  Orders are therefore compared up to permutations inside groups of equal distance."""
from __future__ import annotations

import ctypes as C
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests.golden.make_golden_ref import ref_inputs

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "ref_plain_v1.json").read_text())
EXE = ROOT / "prefhetch_b200" / "host" / "pf_client_check"
NQ, D, CP, K = 5, 128, 200, 100


def f32(bits):
    return np.array(bits, dtype=np.uint32).view(np.float32)


def same_up_to_ties(idx_a, idx_b, dist):
    """equal as sequences except for permutations inside runs of equal `dist` (dist ascending)"""
    idx_a, idx_b, dist = np.asarray(idx_a), np.asarray(idx_b), np.asarray(dist)
    start = 0
    for end in range(1, len(dist) + 1):
        if end == len(dist) or dist[end] != dist[start]:
            if sorted(idx_a[start:end].tolist()) != sorted(idx_b[start:end].tolist()):
                return False
            start = end
    return True


@pytest.fixture(scope="module")
def oracle():
    from oracle import pf_oracle
    pf_oracle.build()
    return pf_oracle


@pytest.fixture(scope="module")
def x():
    return ref_inputs()


def test_fixture_is_what_the_reference_computes(x):
    """re-run the reference's functions (needs /root/reference or a prebuilt oracle/_ref) and compare with the fixture"""
    from oracle import pf_ref as R
    if not R.build():
        pytest.skip("reference sources absent and no prebuilt oracle/_ref: the committed fixture stands on its own")
    assert R.constants() == GOLD["constants"] and GOLD["constants"]["nquery"] == NQ and GOLD["constants"]["coarse_probe"] == CP
    for tag in ("int", "frac"):
        idx, dist = R.sort_nearest_centroids(x[f"query_{tag}"], x["cent"])
        g = GOLD[f"sort_nearest_centroids_{tag}"]
        assert idx.reshape(-1).tolist() == g["idx"] and np.array_equal(dist.reshape(-1), f32(g["dist_bits"]))
        ps = R.precise_search(x[f"query_{tag}"], x["ids"], x[f"base_{tag}"])
        assert np.array_equal(ps.reshape(-1).view(np.uint32), np.array(GOLD[f"precise_search_{tag}"]["score_bits"], dtype=np.uint32))
    idx, dist = R.compute_nearest_coarse_vectors(x["scores"], x["labels"], x["sizes"])
    assert idx.tolist() == GOLD["compute_nearest_coarse_vectors"]["idx"]
    b = R.benchmark_results(x["obs"], x["gt"])
    assert list(b["recall"]) == GOLD["benchmark_results"]["recall"] and list(b["mrr"]) == GOLD["benchmark_results"]["mrr"]


@pytest.mark.parametrize("tag", ["int", "frac"])
def test_oracle_stage1_equals_the_reference(oracle, x, tag):
    """pfo_coarse_quantize vs sort_nearest_centroids (client_lib.cpp:49-81): every distance bit for bit, the same order"""
    g = GOLD[f"sort_nearest_centroids_{tag}"]
    nlist = len(x["cent"])
    idx, dist = oracle.coarse_quantize(x[f"query_{tag}"], x["cent"], nlist)
    want_idx, want_dist = np.array(g["idx"]).reshape(NQ, nlist), f32(g["dist_bits"]).reshape(NQ, nlist)
    assert np.array_equal(dist.view(np.uint32), want_dist.view(np.uint32))
    for i in range(NQ):
        assert same_up_to_ties(idx[i], want_idx[i], want_dist[i])
    assert np.array_equal(idx, want_idx)            # no ties in this input: the order itself


@pytest.mark.parametrize("tag", ["int", "frac"])
def test_oracle_exact_l2_equals_the_reference(oracle, x, tag):
    """pfo_l2sqr_ref vs Server::preciseSearch (server_lib.cpp:140-167), incl. fractional data where the float / double
    accumulation order shows in the low bits"""
    want = np.array(GOLD[f"precise_search_{tag}"]["score_bits"], dtype=np.uint32).reshape(NQ, CP)
    q, base, ids = x[f"query_{tag}"], x[f"base_{tag}"], x["ids"]
    fn = oracle.lib().pfo_l2sqr_ref
    fn.restype = C.c_float
    got = np.zeros((NQ, CP), dtype=np.float32)
    for i in range(NQ):
        for j in range(CP):
            row = np.ascontiguousarray(base[ids[i, j]])
            got[i, j] = fn(row.ctypes.data_as(C.POINTER(C.c_float)), q[i].ctypes.data_as(C.POINTER(C.c_float)), C.c_size_t(D))
    assert np.array_equal(got.view(np.uint32), want)
    if tag == "int":        # integer data: exact integers, as DESIGN §1 argues
        exact = ((base[ids].astype(np.int64) - q[:, None, :].astype(np.int64)) ** 2).sum(-1)
        assert np.array_equal(got.astype(np.int64), exact)


def test_oracle_recall_equals_the_reference(oracle, x):
    """pfo_recall vs benchmark_results (client_lib.cpp:243-337): the reference's own recall@1/10/100 and MRR@10"""
    r = oracle.recall(x["obs"], x["gt"])
    want = GOLD["benchmark_results"]
    assert abs(r["ref_recall_1"] - want["recall"][0]) < 1e-6 and abs(r["ref_recall_10"] - want["recall"][1]) < 1e-6
    assert abs(r["ref_recall_100"] - want["recall"][2]) < 1e-6 and abs(r["mrr_10"] - want["mrr"][1]) < 1e-6


def _case(tmp, nprobe, coarse_probe):
    from tests.test_client import write_case
    write_case(tmp, 2048, [12289, 40961], 65537, D, 1, 16, np.zeros((NQ, D), dtype=np.int64), nprobe, coarse_probe, bytes(64))


def test_product_client_equals_the_reference(x, tmp_path):
    """prefhetch::Client's plaintext steps (host/pf_client.hpp) against the reference's outputs"""
    assert EXE.exists(), "run __graft_entry__.build() first"
    nlist = len(x["cent"])
    # sort_nearest_centroids
    for tag in ("int", "frac"):
        d = tmp_path / f"nearest_{tag}"
        _case(d, nlist, 1)
        x[f"query_{tag}"].tofile(d / "queries.f32")
        x["cent"].tofile(d / "centroids.f32")
        r = subprocess.run([str(EXE), "nearest", str(d)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        g = GOLD[f"sort_nearest_centroids_{tag}"]
        assert np.fromfile(d / "nearest_centroids.i64", dtype=np.int64).tolist() == g["idx"]
        assert np.array_equal(np.fromfile(d / "nearest_centroids.f32", dtype=np.float32).view(np.uint32), np.array(g["dist_bits"], dtype=np.uint32))
    # compute_nearest_coarse_vectors (many ties) and its COARSE_PROBE check
    d = tmp_path / "coarse"
    _case(d, 1, CP)
    x["scores"].tofile(d / "scores.f32")
    x["labels"].tofile(d / "labels.i64")
    x["sizes"].tofile(d / "list_sizes.u64")
    r = subprocess.run([str(EXE), "coarse-rank", str(d)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    g = GOLD["compute_nearest_coarse_vectors"]
    got_i, got_d = np.fromfile(d / "coarse_ranked.i64", dtype=np.int64), np.fromfile(d / "coarse_ranked.f32", dtype=np.float32)
    assert np.array_equal(got_d.view(np.uint32), np.array(g["dist_bits"], dtype=np.uint32))
    o = 0
    for n in x["sizes"].astype(int):
        assert same_up_to_ties(got_i[o:o + n], np.array(g["idx"])[o:o + n], got_d[o:o + n])
        o += n
    short = x["sizes"].copy()
    short[1] = 199
    short.tofile(d / "list_sizes.u64")
    x["scores"][:-1].tofile(d / "scores.f32")
    x["labels"][:-1].tofile(d / "labels.i64")
    r = subprocess.run([str(EXE), "coarse-rank", str(d)], capture_output=True, text=True, timeout=120)
    assert GOLD["compute_nearest_coarse_vectors_short"] == "threw" and r.returncode == 1 and "COARSE_PROBE" in r.stderr
    # compute_nearest_precise_vectors + benchmark_results
    d = tmp_path / "rank"
    _case(d, 1, CP)
    x["pscores"].tofile(d / "precise_scores.f32")
    x["ids"].tofile(d / "coarse_ids.i64")
    r = subprocess.run([str(EXE), "rank", str(d)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    g = GOLD["compute_nearest_precise_vectors"]
    got_i, got_d = np.fromfile(d / "ranked.i64", dtype=np.int64).reshape(NQ, CP), np.fromfile(d / "ranked.f32", dtype=np.float32).reshape(NQ, CP)
    assert np.array_equal(got_d.reshape(-1).view(np.uint32), np.array(g["dist_bits"], dtype=np.uint32))
    for i in range(NQ):
        assert same_up_to_ties(got_i[i], np.array(g["idx"]).reshape(NQ, CP)[i], got_d[i])
    # benchmark_results on the reference's own shape: K = 100 returned ids
    d = tmp_path / "bench"
    _case(d, 1, K)
    np.zeros((NQ, K), dtype=np.float32).tofile(d / "precise_scores.f32")       # all equal: the stable sort keeps the order
    x["obs"].tofile(d / "coarse_ids.i64")
    x["gt"].tofile(d / "groundtruth.i32")
    r = subprocess.run([str(EXE), "rank", str(d)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got = [float(v) for v in (d / "benchmark.txt").read_text().split()]
    want = GOLD["benchmark_results"]["recall"] + GOLD["benchmark_results"]["mrr"]
    assert all(abs(a - b) < 1e-6 for a, b in zip(got, want)), (got, want)
