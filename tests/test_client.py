"""The client side of the encrypted endpoint (prefhetch_b200/host/pf_client.hpp) against the CPU
oracle playing the server: keys, query ciphertexts and GaloisKeys made by the C++ client are consumed by the
oracle's pipeline (the checker the CUDA engine is bit-exact with), and the client decrypts what comes back into the
reference's packed coarse scores (ref: src/client/client_lib.cpp:122-156).  No GPU; the same round trip with the
engine as the server is tests/test_gpu_parity.py::test_cpp_client_round_trip."""
from __future__ import annotations

import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests.util import build_ivf, ntt_primes, sift_like

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "prefhetch_b200" / "host" / "pf_client_check"


@pytest.fixture(scope="module")
def oracle():
    from oracle import pf_oracle
    pf_oracle.build()
    return pf_oracle


@pytest.fixture(scope="module")
def exe():
    src = EXE.with_suffix(".cpp")
    hdr = EXE.parent / "pf_client.hpp"
    if not EXE.exists() or EXE.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["/usr/bin/g++", "-std=c++20", "-O2", "-Wall", "-Wextra", "-o", str(EXE), str(src)], check=True)
    return EXE


def write_case(tmp, n, primes, t, d, m, g, queries, nprobe, coarse_probe, seed):
    tmp.mkdir(parents=True, exist_ok=True)
    (tmp / "params.txt").write_text(" ".join(str(v) for v in [d, n, t, m, g, len(queries), nprobe, coarse_probe, len(primes), *primes]) + "\n")
    (tmp / "seed.bin").write_bytes(seed)
    np.ascontiguousarray(queries, dtype=np.int64).tofile(tmp / "queries.i64")


def parse_galois_keys(blob: bytes, n, k, L):
    """inverse of tests/util.galois_keys_save: {galois_elt: words [L][2][k][n]}, parms_id"""
    assert blob[:8] == bytes([0x5E, 0xA1, 0x10, 4, 1, 0, 0, 0]) and struct.unpack_from("<Q", blob, 8)[0] == len(blob)
    pid = struct.unpack_from("<4Q", blob, 16)
    dim1, = struct.unpack_from("<Q", blob, 48)
    assert dim1 == n
    pos, keys = 56, {}
    per = 113 + 2 * k * n * 8
    for slot in range(n):
        dim2, = struct.unpack_from("<Q", blob, pos)
        pos += 8
        if not dim2:
            continue
        assert dim2 == L
        w = np.zeros((L, 2, k, n), dtype=np.uint64)
        for j in range(L):
            s = blob[pos:pos + per]
            assert s[:8] == bytes([0x5E, 0xA1, 0x10, 4, 1, 0, 0, 0]) and struct.unpack_from("<Q", s, 8)[0] == per
            assert struct.unpack_from("<4Q", s, 16) == pid and s[48] == 1                 # key-level parms_id, NTT form
            assert struct.unpack_from("<3Q", s, 49) == (2, n, k)
            w[j] = np.frombuffer(bytes(s[113:]), dtype=np.uint64).reshape(2, k, n)
            pos += per
        keys[2 * slot + 1] = w
    assert pos == len(blob)
    return keys, pid


def oracle_secret_key(ctx, coeffs):
    sk = np.zeros((ctx.k, ctx.n), dtype=np.uint64)
    for j, q in enumerate(ctx.primes):
        sk[j] = ctx.ntt_fwd(np.where(coeffs < 0, q - 1, coeffs.astype(np.int64)).astype(np.uint64), j)
    return sk


def split(blob: bytes, offs):
    return [bytes(blob[int(offs[i]):int(offs[i + 1])]) for i in range(len(offs) - 1)]


def oracle_server(oracle, ctx, lay, cts, keys, idx, offsets, vecs, result_limbs, parms_id):
    """the encrypted coarseSearch as the oracle computes it: (result streams, probed_sizes, results_per_query, labels)"""
    nlist = len(offsets) - 1
    block_off, block_n, first_block = [], [], []
    for l in range(nlist):
        first_block.append(len(block_off))
        for b0 in range(int(offsets[l]), int(offsets[l + 1]), lay.C):
            block_off.append(b0)
            block_n.append(min(lay.C, int(offsets[l + 1]) - b0))
    first_block.append(len(block_off))
    diag, norm = oracle.encode_blocks(ctx, lay, vecs.astype(np.int32), block_off, block_n)
    pq, pb = [], []
    probed = np.zeros(idx.shape, dtype=np.uint64)
    rpq = np.zeros(len(idx), dtype=np.uint64)
    label_pos = []
    for qi in range(len(idx)):
        for p, l in enumerate(idx[qi]):
            probed[qi, p] = offsets[l + 1] - offsets[l]
            label_pos.append(np.arange(offsets[l], offsets[l + 1]))
            for b in range(first_block[l], first_block[l + 1]):
                pq.append(qi)
                pb.append(b)
                rpq[qi] += 1
    out, _ = oracle.search_pairs(ctx, lay, cts, keys, False, pq, pb, diag, norm, result_limbs=result_limbs)
    streams = [ctx.ct_save(o, parms_id=parms_id) for o in out]
    return streams, probed, rpq, np.concatenate(label_pos)


CASES = [
    # n, primes / t, d, m, g, result_limbs
    pytest.param(2048, None, 128, 1, 16, 1, id="n2048-m1-g16-rl1"),
    pytest.param(2048, None, 128, 2, 8, 0, id="n2048-m2-g8-full"),
    pytest.param(2048, None, 96, 1, 32, 2, id="n2048-d96-g32-rl2"),
    pytest.param(8192, "bfv", 128, 1, 8, 1, id="n8192-bfvdefault-rl1"),
    pytest.param(16384, "bfv", 128, 1, 16, 2, id="n16384-bfvdefault-rl2"),     # L = 8, 9 key primes (configs[4] sweep)
    pytest.param(8192, "bfv27", 960, 8, 8, 1, id="n8192-gist960-m8-rl1"),      # configs[3]: 8 query ciphertexts, 27-bit t
]


@pytest.mark.parametrize("n,pset,d,m,g,rl", CASES)
def test_client_round_trip_against_the_oracle(oracle, exe, tmp_path, n, pset, d, m, g, rl):
    import prefhetch_b200 as pf
    if pset in ("bfv", "bfv27"):
        primes, t = oracle.BFV_DEFAULT_PRIMES[n], oracle.BATCHING_T[(n, 27 if pset == "bfv27" else 24)]
    else:
        primes, t = ntt_primes(n, 40, 3) + ntt_primes(n, 41, 1), ntt_primes(n, 24, 1)[0]
    ctx = oracle.Context(n, primes, t)
    lay = oracle.LayoutPlan(n, d, m, g)
    rng = np.random.default_rng(n + d + m)
    nlist, nprobe, nq, coarse_probe = 6, 2, 2, 10
    base, query, cent = sift_like(rng, 700 if n == 2048 else 1500, d, nlist, nq)
    offsets, ids, vecs = build_ivf(base, cent)
    seed = rng.bytes(64)
    write_case(tmp_path, n, primes, t, d, m, g, query.astype(np.int64), nprobe, coarse_probe, seed)
    r = subprocess.run([str(exe), "keygen", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok keygen"), r.stdout + r.stderr

    # BatchEncoder: the plaintext polynomial itself
    ramp = (np.arange(n, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(17)) % np.uint64(t)
    assert np.array_equal(np.fromfile(tmp_path / "encode_probe.u64", dtype=np.uint64), ctx.encode(ramp))

    # secret key: ternary, roughly balanced
    coeffs = np.fromfile(tmp_path / "sk.i8", dtype=np.int8)
    assert coeffs.shape == (n,) and set(np.unique(coeffs)) <= {-1, 0, 1} and abs(int((coeffs == 0).sum()) - n / 3) < n / 8
    sk = oracle_secret_key(ctx, coeffs)

    data_id = pf.parms_id(n, primes[:-1], t)
    key_id = pf.parms_id(n, primes, t)
    per_ct = 113 + 2 * ctx.L * n * 8
    cts = np.zeros((nq, m, 2, ctx.L, n), dtype=np.uint64)
    for kind in ("seeded", "full"):
        blob = (tmp_path / f"queries_{kind}.bin").read_bytes()
        offs = np.fromfile(tmp_path / f"queries_{kind}.off", dtype=np.uint64)
        parts = split(blob, offs)
        assert len(parts) == nq * m and int(offs[-1]) == len(blob)
        budgets = []
        for i, s in enumerate(parts):
            assert len(s) == (113 + ctx.L * n * 8 + 81 if kind == "seeded" else per_ct)
            full = pf.seal_ct_expand(s, n, primes[:-1])                # the product's own loader of seeded streams
            ct, is_ntt, pid, used = oracle.Context.ct_load(full)
            assert used == per_ct and not is_ntt and pid == data_id and ct.shape == (2, ctx.L, n)
            if kind == "seeded":                                      # c1 is the expansion of the seed in the stream
                assert s[-65] == 1 and np.array_equal(ct[1], ctx.sample_poly_uniform(s[-64:]))
            plain, budget = ctx.decrypt(sk, ct)
            want = lay.query_slots(t, query[i // m].astype(np.int64), i % m)
            assert np.array_equal(ctx.decode(plain), want) and budget > 0
            budgets.append(budget)
            if kind == "seeded":
                cts[i // m, i % m] = ct
        fresh = ctx.decrypt(sk, ctx.encrypt(sk, ctx.encode(want), 5))[1]
        assert abs(min(budgets) - fresh) <= 2 and max(budgets) - min(budgets) <= 2, (budgets, fresh)

    # GaloisKeys: the stream that travels is seeded like Serializable<GaloisKeys> (every key ciphertext = c0 + the
    # seed of c1); the product expands it (pf_seal_galois_keys_expand, also inside pf_load_galois_keys) to exactly the
    # full stream a twin client with the same seed writes; compressed forms too
    gk_seeded, gk_full = (tmp_path / "galois_keys.bin").read_bytes(), (tmp_path / "galois_keys_full.bin").read_bytes()
    nkeys = lay.R - 1
    assert len(gk_full) == 56 + 8 * n + nkeys * ctx.L * (113 + 2 * ctx.k * n * 8)
    assert len(gk_seeded) == 56 + 8 * n + nkeys * ctx.L * (113 + ctx.k * n * 8 + 81)
    assert pf.seal_galois_keys_expand(gk_seeded, n, primes) == gk_full
    assert pf.seal_galois_keys_expand(gk_full, n, primes) == gk_full
    from tests.util import have_zstd, zlib_stream, zstd_stream
    assert pf.seal_galois_keys_expand(zlib_stream(gk_seeded), n, primes) == gk_full
    if have_zstd():
        assert pf.seal_galois_keys_expand(zstd_stream(gk_seeded, streaming=True), n, primes) == gk_full
    if nkeys:
        first = 56 + 8 * ((ctx.galois_elt(1) - 1) // 2 + 1)          # the first key ciphertext of step 1
        with pytest.raises(pf.PfError):
            pf.seal_galois_keys_expand(gk_seeded[:first + 113 + ctx.k * n * 8 + 16] + b"\x02" + gk_seeded[first + 113 + ctx.k * n * 8 + 17:], n, primes)  # shake256
        with pytest.raises(pf.PfError):
            pf.seal_galois_keys_expand(gk_seeded[:-40], n, primes)
    # the R-1 rotation keys, at the key level
    keys_by_elt, pid = parse_galois_keys(gk_full, n, ctx.k, ctx.L)
    assert pid == key_id and sorted(keys_by_elt) == sorted(ctx.galois_elt(r) for r in range(1, lay.R))
    keys = [keys_by_elt[ctx.galois_elt(r)] for r in range(1, lay.R)]
    # one key on its own: rotate_rows with it moves the slots by one
    rot = ctx.rotate_rows(cts[0, 0], 1, keys[0])
    slots = ctx.decode(ctx.decrypt(sk, rot)[0])
    src = ctx.decode(ctx.decrypt(sk, cts[0, 0])[0])
    half = n // 2
    assert np.array_equal(slots[:half], np.roll(src[:half], -1)) and np.array_equal(slots[half:], np.roll(src[half:], -1))

    # the server: probed lists -> result ciphertexts (oracle pipeline), response envelope as the engine returns it
    idx, _ = oracle.coarse_quantize(query, cent, nprobe)
    Lr = rl if rl else ctx.L
    streams, probed, rpq, label_pos = oracle_server(oracle, ctx, lay, cts, keys, idx, offsets, vecs, rl, pf.parms_id(n, primes[:Lr], t))
    roff = np.concatenate([[0], np.cumsum([len(s) for s in streams])]).astype(np.uint64)
    (tmp_path / "results.bin").write_bytes(b"".join(streams))
    roff.tofile(tmp_path / "results.off")
    probed.tofile(tmp_path / "probed_sizes.u64")
    rpq.tofile(tmp_path / "results_per_query.u64")
    ids[label_pos].astype(np.int64).tofile(tmp_path / "labels.i64")
    r = subprocess.run([str(exe), "decrypt", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok decrypt"), r.stdout + r.stderr
    scores = np.fromfile(tmp_path / "scores.f32", dtype=np.float32)
    sizes = np.fromfile(tmp_path / "list_sizes.u64", dtype=np.uint64)
    assert np.array_equal(sizes, probed.sum(1))
    # exact squared L2 of every candidate of every probed list, packed per query like Server::coarseSearch
    want = []
    for qi in range(nq):
        for l in idx[qi]:
            xs = vecs[offsets[l]:offsets[l + 1]].astype(np.int64)
            want.append(((xs - query[qi].astype(np.int64)) ** 2).sum(1))
    want = np.concatenate(want)
    assert np.array_equal(scores.astype(np.int64), want)
    assert int((tmp_path / "budget.txt").read_text()) > 0
    # compute_nearest_coarse_vectors: ascending by distance, ties in response order (stable)
    nearest = np.fromfile(tmp_path / "nearest.i64", dtype=np.int64).reshape(nq, coarse_probe)
    lab = ids[label_pos].astype(np.int64)
    for qi in range(nq):
        a, b = int(sizes[:qi].sum()), int(sizes[:qi + 1].sum())
        order = np.argsort(want[a:b], kind="stable")[:coarse_probe]
        assert np.array_equal(nearest[qi], lab[a:b][order])
    

    # the same exchange through the JSON envelope of POST /coarsesearch-encrypted (pf_query_handlers.hpp): the request
    # body the client writes is read back with Python's json, the response is written by Python's json in the
    # engine's slot layout (128-byte aligned words: 15 pad bytes before every stream)
    import base64
    import json
    idx.astype(np.int64).tofile(tmp_path / "nearest_idx.i64")
    r = subprocess.run([str(exe), "request", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok request"), r.stdout + r.stderr
    req = json.loads((tmp_path / "request.json").read_text())
    assert sorted(req) == ["ctOffsets", "nearestCentroidIndexes", "queryCiphertexts"]
    assert base64.b64decode(req["queryCiphertexts"]) == (tmp_path / "queries_seeded.bin").read_bytes()
    assert req["ctOffsets"] == [int(v) for v in np.fromfile(tmp_path / "queries_seeded.off", dtype=np.uint64)]
    assert req["nearestCentroidIndexes"] == idx.tolist()
    slot = (15 + len(streams[0]) + 127) // 128 * 128
    body = bytearray(slot * len(streams))
    for i, st in enumerate(streams):
        body[i * slot + 15:i * slot + 15 + len(st)] = st
    resp = {"resultCiphertexts": base64.b64encode(bytes(body)).decode(), "resultOffsets": [i * slot + 15 for i in range(len(streams))] + [slot * len(streams)],
            "resultBytes": len(streams[0]), "resultsPerQuery": [int(v) for v in rpq], "coarseVectorIndexes": [int(v) for v in lab],
            "listSizesPerQuery": [int(v) for v in sizes], "probedListSizes": probed.astype(np.int64).tolist()}
    (tmp_path / "response.json").write_text(json.dumps(resp, indent=1))      # whitespace the codec has to skip
    for f in ("scores.f32", "list_sizes.u64", "budget.txt"):
        (tmp_path / f).unlink()
    r = subprocess.run([str(exe), "respond", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok respond"), r.stdout + r.stderr
    assert np.array_equal(np.fromfile(tmp_path / "scores.f32", dtype=np.float32).astype(np.int64), want)
    assert np.array_equal(np.fromfile(tmp_path / "labels_out.i64", dtype=np.int64), lab)
    assert np.array_equal(np.fromfile(tmp_path / "list_sizes.u64", dtype=np.uint64), sizes)
    resp["resultsPerQuery"][0] += 1                                          # an envelope that does not add up
    (tmp_path / "response.json").write_text(json.dumps(resp))
    r = subprocess.run([str(exe), "respond", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 1 and "pf_client_check:" in r.stderr

    # a response of another parameter set, a truncated one and a short envelope are refused
    bad = bytearray(streams[0])
    bad[16] ^= 1
    (tmp_path / "results.bin").write_bytes(bytes(bad) + b"".join(streams[1:]))
    r = subprocess.run([str(exe), "decrypt", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 1 and "parms_id" in r.stderr
    (tmp_path / "results.bin").write_bytes(b"".join(streams)[:-8])
    r = subprocess.run([str(exe), "decrypt", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 1


def test_client_stage1_matches_reference_semantics(oracle, exe, tmp_path):
    """sort_nearest_centroids (ref: client_lib.cpp:49-81) in the client library: the same probed lists and the same
    float bits as the oracle's restatement (and hence as the engine's pf_coarse_quantize), ties by centroid order"""
    rng = np.random.default_rng(4)
    d, nlist, nq, nprobe = 128, 300, 9, 17
    base, query, cent = sift_like(rng, 10, d, nlist, nq)
    cent = (cent + rng.normal(0, 0.37, size=cent.shape)).astype(np.float32)
    cent[7] = cent[3]                                   # an exact tie
    query[0] = cent[3] + 1.0
    write_case(tmp_path, 2048, [12289, 40961], 65537, d, 1, 16, np.zeros((nq, d), dtype=np.int64), nprobe, 1, bytes(64))
    query.astype(np.float32).tofile(tmp_path / "queries.f32")
    cent.tofile(tmp_path / "centroids.f32")
    r = subprocess.run([str(exe), "nearest", str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ok nearest"), r.stdout + r.stderr
    idx = np.fromfile(tmp_path / "nearest_centroids.i64", dtype=np.int64).reshape(nq, nprobe)
    dist = np.fromfile(tmp_path / "nearest_centroids.f32", dtype=np.float32).reshape(nq, nprobe)
    oidx, odist = oracle.coarse_quantize(query, cent, nprobe)
    assert np.array_equal(idx, oidx) and np.array_equal(dist.view(np.uint32), odist.view(np.uint32))
    assert list(idx[0, :2]) == [3, 7]


def test_client_ranking_and_benchmark_bookkeeping(oracle, exe, tmp_path):
    """compute_nearest_precise_vectors + benchmark_results (ref: client_lib.cpp:189-209, 246-330) against numpy and
    the oracle's restatement of the reference's recall definition"""
    rng = np.random.default_rng(12)
    nq, K, gt_k = 7, 20, 100
    gt = np.stack([rng.permutation(5000)[:gt_k] for _ in range(nq)]).astype(np.int32)
    cid = np.stack([np.concatenate([rng.permutation(gt[i, :K])[:K - 6], 6000 + np.arange(6)]) for i in range(nq)]).astype(np.int64)
    cid = np.stack([rng.permutation(row) for row in cid])
    scores = rng.integers(0, 50, size=(nq, K)).astype(np.float32)      # many ties: order among equals is the input order
    write_case(tmp_path, 2048, [12289, 40961], 65537, 128, 1, 16, np.zeros((nq, 128), dtype=np.int64), 1, K, bytes(64))
    scores.tofile(tmp_path / "precise_scores.f32")
    cid.tofile(tmp_path / "coarse_ids.i64")
    gt.tofile(tmp_path / "groundtruth.i32")
    r = subprocess.run([str(exe), "rank", str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ok rank"), r.stdout + r.stderr
    ranked = np.fromfile(tmp_path / "ranked.i64", dtype=np.int64).reshape(nq, K)
    want = np.stack([cid[i][np.argsort(scores[i], kind="stable")] for i in range(nq)])
    assert np.array_equal(ranked, want)
    got = [float(v) for v in (tmp_path / "benchmark.txt").read_text().split()]
    ref = oracle.recall(ranked, gt)
    assert abs(got[0] - ref["ref_recall_1"]) < 1e-6 and abs(got[1] - ref["ref_recall_10"]) < 1e-6 and abs(got[2] - ref["ref_recall_100"]) < 1e-6
    assert abs(got[4] - ref["mrr_10"]) < 1e-6 and got[3] <= got[4] <= got[5]


def test_client_rejects_bad_parameters(exe, tmp_path):
    n = 2048
    primes, t = ntt_primes(n, 40, 3) + ntt_primes(n, 41, 1), ntt_primes(n, 24, 1)[0]
    q = np.zeros((1, 128), dtype=np.int64)
    for name, kw in (("g-too-large", dict(g=256)), ("m-not-pow2", dict(m=3)), ("not-ntt-prime", dict(primes=[primes[0] + 2] + primes[1:]))):
        case = dict(n=n, primes=primes, t=t, d=128, m=1, g=16)
        case.update(kw)
        d = tmp_path / name
        write_case(d, case["n"], case["primes"], case["t"], case["d"], case["m"], case["g"], q, 1, 1, bytes(64))
        r = subprocess.run([str(exe), "keygen", str(d)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 1 and "pf_client_check:" in r.stderr, (name, r.stdout, r.stderr)
