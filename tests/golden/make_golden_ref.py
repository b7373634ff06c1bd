#!/usr/bin/env python
"""Golden vectors from the REFERENCE ITSELF: runs the reference's own plaintext functions — compiled from
/root/reference where the sources lie (oracle/ref_build/Makefile -> oracle/_ref) — on deterministic inputs and writes
their outputs to tests/golden/ref_plain_v1.json.  Needs /root/reference (this container); the fixture travels.

    python tests/golden/make_golden_ref.py

Inputs come from integer formulas (ref_inputs below), not from an RNG, so that the tests rebuild them exactly.
Covered: sort_nearest_centroids (client_lib.cpp:49-81), Server::preciseSearch after Server::init_index
(server_lib.cpp:55-99, 140-167), compute_nearest_coarse_vectors (:122-156, incl. its throw),
compute_nearest_precise_vectors (:189-209), benchmark_results (:243-337).  Floats are stored as their bit patterns."""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))


def _mix(a):
    """a cheap integer hash, vectorised, identical everywhere (uint64 wrap-around arithmetic)"""
    a = np.asarray(a, dtype=np.uint64)
    a = (a ^ (a >> np.uint64(31))) * np.uint64(0x9E3779B97F4A7C15)
    a = (a ^ (a >> np.uint64(29))) * np.uint64(0xBF58476D1CE4E5B9)
    return a ^ (a >> np.uint64(32))


def ref_inputs():
    """deterministic inputs in the reference's shapes (d = 128, NQUERY = 5, COARSE_PROBE = 200, K = 100)"""
    d, nq, cp, k, nlist, nb = 128, 5, 200, 100, 96, 1500
    grid = lambda n, salt: _mix(np.arange(n * d, dtype=np.uint64) + np.uint64(salt)).reshape(n, d)   # noqa: E731
    # SIFT-like integers 0..255, and a fractional variant (eighths + a few thirds) that makes the float / double
    # accumulation order of the reference's loops observable in the low bits
    base_int = (grid(nb, 1) % np.uint64(256)).astype(np.float32)
    base_frac = (base_int + (grid(nb, 2) % np.uint64(8)).astype(np.float32) / np.float32(8.0) + (grid(nb, 3) % np.uint64(3)).astype(np.float32) / np.float32(3.0)).astype(np.float32)
    query_int = (grid(nq, 4) % np.uint64(256)).astype(np.float32)
    query_frac = (query_int + (grid(nq, 5) % np.uint64(16)).astype(np.float32) / np.float32(16.0) + np.float32(0.1)).astype(np.float32)
    cent = ((grid(nlist, 6) % np.uint64(1 << 20)).astype(np.float64) / 4096.0).astype(np.float32)           # 0 .. 256, fractional
    ids = (_mix(np.arange(nq * cp, dtype=np.uint64) + np.uint64(7)) % np.uint64(nb)).astype(np.int64).reshape(nq, cp)
    # coarse stage outputs: per query sizes >= COARSE_PROBE, float scores with many ties
    sizes = np.array([230, 200, 417, 301, 256], dtype=np.uint64)
    tot = int(sizes.sum())
    scores = (_mix(np.arange(tot, dtype=np.uint64) + np.uint64(8)) % np.uint64(90)).astype(np.float32) * np.float32(0.5)
    labels = (_mix(np.arange(tot, dtype=np.uint64) + np.uint64(9)) % np.uint64(1 << 40)).astype(np.int64)
    pscores = (_mix(np.arange(nq * cp, dtype=np.uint64) + np.uint64(10)) % np.uint64(60)).astype(np.float32)
    # ground truth / observed ids for benchmark_results: K returned ids, gt_k = 100 neighbours per query
    gt = np.stack([(_mix(np.arange(100, dtype=np.uint64) + np.uint64(1000 * (i + 1))) % np.uint64(1 << 20)).astype(np.int32) for i in range(nq)])
    obs = gt.astype(np.int64).copy()
    for i in range(nq):       # shuffle deterministically, drop some hits
        perm = np.argsort(_mix(np.arange(k, dtype=np.uint64) + np.uint64(77 * (i + 1))), kind="stable")
        obs[i] = obs[i][perm]
        obs[i, (perm % 7) == 3] += 1 << 21
    obs[0, 0] = gt[0, 0]                                        # one query returns the true nearest first
    return dict(base_int=base_int, base_frac=base_frac, query_int=query_int, query_frac=query_frac, cent=cent, ids=ids, sizes=sizes,
                scores=scores, labels=labels, pscores=pscores, gt=gt, obs=obs)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).reshape(-1).tolist()


def main():
    from oracle import pf_ref as R
    assert R.buildable(), "needs the reference sources (/root/reference)"
    R.build(force=True)
    x = ref_inputs()
    out = {"what": "outputs of the reference's own functions (oracle/_ref, compiled from /root/reference); inputs: ref_inputs()",
           "constants": R.constants()}
    for tag in ("int", "frac"):
        idx, dist = R.sort_nearest_centroids(x[f"query_{tag}"], x["cent"])
        out[f"sort_nearest_centroids_{tag}"] = {"idx": idx.reshape(-1).tolist(), "dist_bits": bits(dist)}
        ps = R.precise_search(x[f"query_{tag}"], x["ids"], x[f"base_{tag}"])
        out[f"precise_search_{tag}"] = {"score_bits": bits(ps)}
    idx, dist = R.compute_nearest_coarse_vectors(x["scores"], x["labels"], x["sizes"])
    out["compute_nearest_coarse_vectors"] = {"idx": idx.tolist(), "dist_bits": bits(dist)}
    short = x["sizes"].copy()
    short[1] = 199
    try:
        R.compute_nearest_coarse_vectors(x["scores"][:-1], x["labels"][:-1], short)
        out["compute_nearest_coarse_vectors_short"] = "returned"
    except RuntimeError:
        out["compute_nearest_coarse_vectors_short"] = "threw"
    idx, dist = R.compute_nearest_precise_vectors(x["pscores"], x["ids"])
    out["compute_nearest_precise_vectors"] = {"idx": idx.reshape(-1).tolist(), "dist_bits": bits(dist)}
    b = R.benchmark_results(x["obs"], x["gt"])
    out["benchmark_results"] = {"recall": list(b["recall"]), "mrr": list(b["mrr"])}
    path = Path(__file__).resolve().parent / "ref_plain_v1.json"
    path.write_text(json.dumps(out, separators=(",", ":")) + "\n")
    print(f"wrote {path} ({path.stat().st_size} bytes)", out["benchmark_results"], out["compute_nearest_coarse_vectors_short"])


if __name__ == "__main__":
    main()
