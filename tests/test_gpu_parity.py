"""GPU parity: every C-ABI entry point of libprefhetch_b200.so against the CPU oracle on the same
seeded inputs — bit-exact for all integer / ciphertext work, bit-exact float bits for the
plaintext distances (the kernels restate the reference's float/double arithmetic).

TOLERANCE (BASELINE north star: "decrypted distances and recall@10 must match within a stated tolerance"): ZERO.
Result ciphertexts are compared byte for byte, probed lists index for index, decrypted distances are integers and
must be EQUAL to the exact squared L2 (and hence recall@10 identical to the plaintext pipeline's: the two pipelines
rank the same numbers), and the floating-point outputs of the plaintext stages are compared as float32 BIT PATTERNS.
No assertion in this file uses an epsilon."""
import hashlib

import numpy as np
import pytest

from tests.util import OracleClient, build_ivf, ntt_primes, sift_like

pytestmark = pytest.mark.gpu


def _params(n):
    if n in (8192, 16384):
        from oracle.pf_oracle import BATCHING_T, BFV_DEFAULT_PRIMES
        return BFV_DEFAULT_PRIMES[n], BATCHING_T[(n, 24)]
    k = 4
    return ntt_primes(n, 40, k - 1) + ntt_primes(n, 41, 1), ntt_primes(n, 24, 1)[0]


@pytest.fixture(scope="module")
def pf():
    import prefhetch_b200
    return prefhetch_b200


def _engine(pf, n, d=128, m=1, g=8, primes=None, t=None, **kw):
    p, tt = _params(n)
    return pf.Engine(d, n, primes or p, t or tt, m, g, **kw), (primes or p), (t or tt)


@pytest.mark.parametrize("n", [1024, 2048, 4096, 8192, 16384])
def test_ntt_bit_exact(pf, oracle, n):
    eng, primes, t = _engine(pf, n)
    ctx = oracle.Context(n, primes, t)
    rng = np.random.default_rng(n)
    limbs = list(range(len(primes))) + [-1]
    mods = primes + [t]
    polys = np.stack([rng.integers(0, q, size=n, dtype=np.uint64) for q in mods])
    polys[0, :4] = [0, 1, mods[0] - 1, mods[0] - 2]
    fwd = eng.ntt_forward(polys, limbs)
    for i, l in enumerate(limbs):
        assert np.array_equal(fwd[i], ctx.ntt_fwd(polys[i], l)), f"fwd limb {l}"
    inv = eng.ntt_inverse(polys, limbs)
    for i, l in enumerate(limbs):
        assert np.array_equal(inv[i], ctx.ntt_inv(polys[i], l)), f"inv limb {l}"
    assert np.array_equal(eng.ntt_inverse(fwd, limbs), polys)
    # all-(q-1) and all-zero edge polynomials
    edge = np.stack([np.full(n, mods[0] - 1, dtype=np.uint64), np.zeros(n, dtype=np.uint64)])
    out = eng.ntt_forward(edge, [0, 0])
    assert np.array_equal(out[0], ctx.ntt_fwd(edge[0], 0)) and not out[1].any()
    eng.close()


def test_ntt_against_golden_full_size(pf, golden):
    """the independent Python big-int vectors, straight through the GPU kernels"""
    for case in golden["ntt_sparse"]:
        n, q = case["n"], case["q"]
        eng, primes, _ = _engine(pf, n)
        a = np.zeros(n, dtype=np.uint64)
        for p, c in zip(case["pos"], case["coef"]):
            a[p] = c
        out = eng.ntt_forward(a[None], [primes.index(q)])[0]
        assert out[:4].tolist() == case["first4"]
        assert hashlib.sha256(out.tobytes()).hexdigest() == case["sha256"]
        eng.close()


@pytest.mark.parametrize("n,g,bits", [(2048, 8, 40), (8192, 8, 0), (8192, 1, 0), (16384, 16, 0), (2048, 4, 59)])
def test_ct_pt_mac_bit_exact(pf, oracle, n, g, bits):
    primes, t = _params(n)
    if bits:
        primes = ntt_primes(n, bits, 3) + ntt_primes(n, bits, 4)[3:]
    eng, primes, t = _engine(pf, n, g=g, primes=primes, t=t)
    ctx = oracle.Context(n, primes, t)
    K = eng.index_info()["K"]
    rng = np.random.default_rng(n + g)
    L = len(primes) - 1
    cts = np.stack([np.stack([np.stack([rng.integers(0, primes[l], size=n, dtype=np.uint64) for l in range(L)])
                              for _ in range(2)]) for _ in range(K)])
    pts = np.stack([np.stack([rng.integers(0, primes[l], size=n, dtype=np.uint64) for l in range(L)])
                    for _ in range(K)])
    cts[0, 0, 0, :2] = primes[0] - 1
    pts[0, 0, :2] = primes[0] - 1
    add = np.stack([rng.integers(0, primes[l], size=n, dtype=np.uint64) for l in range(L)])
    want = ctx.mac_plain_ntt(cts, pts)
    assert np.array_equal(eng.ct_pt_mac(cts, pts), want)
    want2 = want.copy()
    for l in range(L):
        want2[0, l] = (want[0, l].astype(object) + add[l].astype(object)) % primes[l]
    assert np.array_equal(eng.ct_pt_mac(cts, pts, add), want2)
    assert np.array_equal(eng.ct_add(cts[0], cts[1]), ctx.add(cts[0], cts[1]))
    assert np.array_equal(eng.ct_to_ntt(cts[:2]), np.stack([ctx.ct_to_ntt(cts[0]), ctx.ct_to_ntt(cts[1])]))
    assert np.array_equal(eng.ct_from_ntt(cts[:1])[0], ctx.ct_from_ntt(cts[0]))
    eng.close()


@pytest.mark.parametrize("n,d,m,g", [(2048, 128, 1, 8), (8192, 128, 1, 8), (8192, 128, 1, 16), (8192, 960, 8, 8),
                                     (2048, 100, 1, 4)])
def test_encode_bit_exact(pf, oracle, n, d, m, g):
    primes, t = _params(n)
    if d > 128:
        t = 133857281 if n == 8192 else t
    eng, primes, t = _engine(pf, n, d=d, m=m, g=g, t=t)
    ctx = oracle.Context(n, primes, t)
    lay = oracle.LayoutPlan(n, d, m, g)
    rng = np.random.default_rng(d + g)
    vals = rng.integers(0, t, size=n, dtype=np.uint64)
    assert np.array_equal(eng.batch_encode(vals), ctx.encode(vals))
    for nvec in (lay.C, lay.C - 5, 1, 0):
        if nvec < 0:
            continue
        xs = rng.integers(0, 256, size=(nvec, d), dtype=np.int32)
        if nvec:
            xs[0] = 255
        diag, norm = eng.encode_block(xs)
        odiag, onorm = oracle.encode_block(ctx, lay, xs)
        assert np.array_equal(diag, odiag), f"diag nvec={nvec}"
        assert np.array_equal(norm, onorm), f"norm nvec={nvec}"
    eng.close()


@pytest.mark.parametrize("n", [2048, 8192])
def test_rotate_rows_bit_exact(pf, oracle, n):
    primes, t = _params(n)
    eng, primes, t = _engine(pf, n)
    cl = OracleClient(oracle, n, primes, t, 128, 1, 8)
    rng = np.random.default_rng(n)
    vals = rng.integers(0, t, size=n, dtype=np.uint64)
    ct = cl.ctx.encrypt(cl.sk, cl.ctx.encode(vals), 5)
    for step in (1, 3, -1):
        key = cl.galois_key(step)
        eng.set_galois_key(eng.galois_elt(step), key)
        assert eng.galois_elt(step) == cl.ctx.galois_elt(step)
        got = eng.rotate_rows(ct, step)
        want = cl.ctx.rotate_rows(ct, step, key)
        assert np.array_equal(got, want), f"step {step}"
        plain, budget = cl.ctx.decrypt(cl.sk, got)
        half = n // 2
        assert np.array_equal(cl.ctx.decode(plain),
                              np.concatenate([np.roll(vals[:half], -step), np.roll(vals[half:], -step)]))
    # a c1 with zero coefficients leaves the hoisted path (its residue identity needs x != 0) and must
    # still match SEAL's semantics bit for bit
    ctz = ct.copy()
    ctz[1, 0, 5] = 0
    ctz[1, 2, n - 1] = 0
    ctz[1, 1, 0] = 0
    for step in (1, 3):
        assert np.array_equal(eng.rotate_rows(ctz, step), cl.ctx.rotate_rows(ctz, step, cl.galois_key(step)))
    with pytest.raises(pf.PfError):
        eng.rotate_rows(ct, 7)  # no key loaded for this step
    eng.close()


@pytest.mark.parametrize("n,d,m,g,chain", [(2048, 128, 1, 16, False), (2048, 128, 1, 16, True),
                                           (8192, 128, 1, 8, False), (2048, 256, 2, 32, False)])
def test_rotated_query_set_bit_exact(pf, oracle, n, d, m, g, chain):
    primes, t = _params(n)
    eng, primes, t = _engine(pf, n, d=d, m=m, g=g)
    cl = OracleClient(oracle, n, primes, t, d, m, g)
    keys = [cl.galois_key(1)] if chain else cl.step_keys()
    for i, key in enumerate(keys):
        eng.set_galois_key(eng.galois_elt(i + 1), key)
    q = np.random.default_rng(1).integers(0, 256, size=d)
    cts = cl.encrypt_query(q, 77)
    got = eng.rotate_query_set(cts, chain)
    want = oracle.rotate_query_set(cl.ctx, cl.lay, cts, keys, chain)
    assert np.array_equal(got, want)
    eng.close()


def _dataset(seed, nb=6000, d=128, nlist=48, nq=7, frac_centroids=True):
    rng = np.random.default_rng(seed)
    base, query, cent = sift_like(rng, nb, d, nlist, nq)
    if frac_centroids:
        cent = (cent + rng.normal(0, 0.3, size=cent.shape)).astype(np.float32)
    offsets, ids, vecs = build_ivf(base, cent)
    return base, query, cent, offsets, ids, vecs


def test_plain_stages_match_reference_semantics(pf, oracle):
    base, query, cent, offsets, ids, vecs = _dataset(3)
    cent[5] = cent[9]  # exact tie between two centroids -> order pinned by index
    eng, _, _ = _engine(pf, 2048)
    eng.load_index(cent, offsets, ids, vecs)
    assert np.array_equal(eng.retrieve_centroids(), cent)
    for nprobe in (1, 5, 48):
        idx, dist = eng.coarse_quantize(query, nprobe, return_dist=True)
        oidx, odist = oracle.coarse_quantize(query, cent, nprobe)
        assert np.array_equal(idx, oidx)
        assert np.array_equal(dist.view(np.uint32), odist.view(np.uint32))
    with pytest.raises(pf.PfError):
        eng.coarse_quantize(query, 49)  # ref: client_lib.cpp:96-99 throws
    idx = eng.coarse_quantize(query, 6)
    dist, labels, sizes = eng.coarseSearch(query, idx)
    odist, olabels, osizes = oracle.search_lists_plain(query, idx, offsets, ids, vecs)
    assert np.array_equal(sizes, osizes) and np.array_equal(labels, olabels)
    assert np.array_equal(dist.view(np.uint32), odist.view(np.uint32))
    # Server::preciseSearch on arbitrary base rows
    rows = np.random.default_rng(4).integers(0, len(base), size=(len(query), 50))
    got = eng.preciseSearch(query, rows)
    want = ((base[rows].astype(np.int64) - query[:, None, :].astype(np.int64)) ** 2).sum(-1)
    assert np.array_equal(got.astype(np.int64), want)
    # non-integer queries still follow the float/double arithmetic exactly
    fq = (query + np.float32(0.37)).astype(np.float32)
    d2, _, _ = eng.coarseSearch(fq, idx)
    o2, _, _ = oracle.search_lists_plain(fq, idx, offsets, ids, vecs)
    assert np.array_equal(d2.view(np.uint32), o2.view(np.uint32))
    with pytest.raises(pf.PfError):
        eng.coarseSearch(query, np.full((len(query), 2), 48, dtype=np.int64))  # list id out of range
    eng.close()


@pytest.mark.parametrize("n,g,chain,world,rl", [(2048, 16, False, 1, 0), (8192, 8, False, 1, 0), (2048, 16, True, 1, 0),
                                                (2048, 16, False, 2, 0), (8192, 8, False, 1, 2), (2048, 16, False, 1, 1),
                                                (8192, 8, False, 1, 1)])
def test_encrypted_search_end_to_end(pf, oracle, n, g, chain, world, rl):
    """serialized query ciphertexts -> pf_search_lists_encrypted -> bytes identical to the oracle's
    pipeline; decrypted distances == exact integer squared L2 of the plaintext path."""
    d, nprobe = 128, 5
    base, query, cent, offsets, ids, vecs = _dataset(n + g, nb=5000 if n == 2048 else 9000, nlist=24, nq=4)
    offsets = offsets.copy()
    primes, t = _params(n)
    cl = OracleClient(oracle, n, primes, t, d, 1, g)
    keys = [cl.galois_key(1)] if chain else cl.step_keys()
    cts = np.stack([cl.encrypt_query(q, 100 + i) for i, q in enumerate(query)])
    blob, offs = cl.serialize_queries(cts)
    C_ = cl.lay.C
    all_results = {}
    for rank in range(world):
        eng, _, _ = _engine(pf, n, g=g, rank=rank, world=world, result_limbs=rl)
        eng.load_index(cent, offsets, ids, vecs)
        eng.set_list_sizes(offsets)
        for i, key in enumerate(keys):
            eng.set_galois_key(eng.galois_elt(i + 1), key)
        idx = eng.coarse_quantize(query, nprobe)
        res = eng.coarseSearchEncrypted(blob, offs, idx)
        # oracle pipeline on the same inputs
        r = 0
        lab_off = 0
        for qi in range(len(query)):
            rot = oracle.rotate_query_set(cl.ctx, cl.lay, cts[qi], keys, chain)
            nres_q = 0
            for p in range(nprobe):
                l = idx[qi, p]
                if l % world != rank:
                    assert res.probed_sizes[qi, p] == 0
                    continue
                n_l = int(offsets[l + 1] - offsets[l])
                assert res.probed_sizes[qi, p] == n_l
                assert np.array_equal(res.labels[lab_off:lab_off + n_l], ids[offsets[l]:offsets[l + 1]])
                lab_off += n_l
                for b0 in range(0, n_l, C_):
                    xs = vecs[offsets[l] + b0: offsets[l] + min(b0 + C_, n_l)].astype(np.int32)
                    diag, norm = oracle.encode_block(cl.ctx, cl.lay, xs)
                    want_ct = oracle.block_distance(cl.ctx, cl.lay, rot, diag, norm)
                    pid = (0, 0, 0, 0)
                    if rl:
                        want_ct = cl.mod_switch_to(want_ct, rl)  # SEAL mod_switch_to_inplace before save
                        pid = _parms_id_py(n, primes[:rl], t)     # parms_id of the level switched to
                    got_bytes = res.result(r)
                    assert got_bytes == cl.ctx.ct_save(want_ct, parms_id=pid), f"rank {rank} query {qi} result {r}"
                    got_ct, is_ntt = eng.ct_deserialize(got_bytes)
                    dist, budget = cl.distances(got_ct, query[qi], len(xs))
                    want = ((xs.astype(np.int64) - query[qi].astype(np.int64)) ** 2).sum(1)
                    assert np.array_equal(dist, want) and budget > 0 and not is_ntt
                    for j, vid in enumerate(ids[offsets[l] + b0: offsets[l] + b0 + len(xs)]):
                        all_results.setdefault(qi, {})[int(vid)] = int(dist[j])
                    r += 1
                    nres_q += 1
            assert res.results_per_query[qi] == nres_q
        assert r == res.stats["nresults"] and lab_off == res.list_sizes.sum()
        eng.close()
    # union over ranks == the plaintext stage-2 result, hence identical recall for both pipelines
    eng, _, _ = _engine(pf, n, g=g)
    eng.load_index(cent, offsets, ids, vecs)
    idx = eng.coarse_quantize(query, nprobe)
    dist, labels, sizes = eng.coarseSearch(query, idx)
    off = 0
    for qi in range(len(query)):
        plain = dict(zip(labels[off:off + sizes[qi]].tolist(), dist[off:off + sizes[qi]].astype(np.int64).tolist()))
        assert plain == all_results[qi]
        off += sizes[qi]
    eng.close()


@pytest.mark.parametrize("old", [False, True])
def test_coarse_quantize_large_nlist(pf, oracle, old, monkeypatch):
    """stage 1 at nlist / nprobe sizes of the multi-GPU configs (radix-select top-k and the iterative
    kernel): same probed lists, same order, same float distances as the reference arithmetic."""
    if old:
        monkeypatch.setenv("PF_TOPK_ITER", "1")
    rng = np.random.default_rng(77)
    nlist, d = 1500, 128
    cent = rng.integers(0, 200, size=(nlist, d)).astype(np.float32)
    cent[700] = cent[3]          # ties
    cent[1499] = cent[3]
    query = np.concatenate([rng.integers(0, 200, size=(5, d)), cent[3:4], cent[900:901]]).astype(np.float32)
    offsets = np.arange(nlist + 1, dtype=np.int64)
    ids = np.arange(nlist, dtype=np.int64)
    eng, _, _ = _engine(pf, 2048)
    eng.load_index(cent, offsets, ids, cent.copy())
    for nprobe in (1, 100, 257, 1500) if not old else (100,):
        idx, dist = eng.coarse_quantize(query, nprobe, return_dist=True)
        oidx, odist = oracle.coarse_quantize(query, cent, nprobe)
        assert np.array_equal(idx, oidx)
        assert np.array_equal(dist.view(np.uint32), odist.view(np.uint32))
    eng.close()


def _parms_id_py(n, primes, t):
    import hashlib
    import struct
    words = [1, n, *primes, t]
    return struct.unpack("<4Q", hashlib.blake2b(struct.pack(f"<{len(words)}Q", *words), digest_size=32).digest())


def _zlib_stream(raw: bytes) -> bytes:
    """what SEAL writes with compr_mode_type::zlib: header (compr_mode 1, new size) + deflate of the body"""
    import struct
    import zlib
    body = zlib.compress(raw[16:], 6)
    return raw[:5] + b"\x01" + raw[6:8] + struct.pack("<Q", 16 + len(body)) + body


def test_zlib_compressed_streams(pf, oracle):
    """SEAL compr_mode zlib on the request side: query ciphertexts, single ciphertexts and
    GaloisKeys saved compressed give the same bytes out as their uncompressed form."""
    n, g, d, nprobe = 2048, 16, 128, 3
    base, query, cent, offsets, ids, vecs = _dataset(21, nb=3000, nlist=16, nq=2)
    primes, t = _params(n)
    cl = OracleClient(oracle, n, primes, t, d, 1, g)
    keys = cl.step_keys()
    cts = np.stack([cl.encrypt_query(q, 300 + i) for i, q in enumerate(query)])
    blob, offs = cl.serialize_queries(cts)
    zparts = [_zlib_stream(bytes(blob[offs[i]:offs[i + 1]])) for i in range(len(cts))]
    zblob = np.frombuffer(b"".join(zparts), dtype=np.uint8)
    zoffs = np.concatenate([[0], np.cumsum([len(z) for z in zparts])]).astype(np.uint64)
    assert len(zblob) < 0.8 * len(blob)
    eng, _, _ = _engine(pf, n, g=g)
    eng.load_index(cent, offsets, ids, vecs)
    eng.set_list_sizes(offsets)
    for i, key in enumerate(keys):
        eng.set_galois_key(eng.galois_elt(i + 1), key)
    idx = eng.coarse_quantize(query, nprobe)
    plain = eng.coarseSearchEncrypted(blob, offs, idx)
    want = [plain.result(r) for r in range(plain.stats["nresults"])]
    comp = eng.coarseSearchEncrypted(zblob, zoffs, idx)
    assert comp.stats["nresults"] == len(want) and all(comp.result(r) == want[r] for r in range(len(want)))
    got, is_ntt = eng.ct_deserialize(zparts[0])
    assert np.array_equal(got, cts[0][0]) and not is_ntt
    with pytest.raises(pf.PfError):
        eng.ct_deserialize(zparts[0][:-20])          # truncated deflate stream
    with pytest.raises(pf.PfError):
        eng.ct_deserialize(zparts[0][:5] + b"\x02" + zparts[0][6:])   # a deflate body labelled zstd
    eng.close()


def test_index_from_faiss_file(pf, oracle, tmp_path):
    """Cached-index branch of Server::init_index (ref: src/server/server_lib.cpp:88-100)."""
    from prefhetch_b200 import faiss_io
    base, query, cent, offsets, ids, vecs = _dataset(11, nb=3000, nlist=32)
    lists = [ids[offsets[l]:offsets[l + 1]] for l in range(32)]
    f = faiss_io.IVFPQFile(128, len(ids), 32, 20, cent, lists, [np.zeros((len(x), 32), np.uint8) for x in lists])
    faiss_io.write_ivfpq(str(tmp_path / "idx.faiss"), f)
    eng, _, _ = _engine(pf, 2048)
    info = eng.load_index_from_faiss(str(tmp_path / "idx.faiss"), base)
    assert info["nlist"] == 32 and info["ntotal"] == len(ids)
    idx = eng.coarse_quantize(query, 4)
    dist, labels, sizes = eng.coarseSearch(query, idx)
    odist, olabels, osizes = oracle.search_lists_plain(query, idx, offsets, ids, vecs)
    assert np.array_equal(labels, olabels) and np.array_equal(dist.view(np.uint32), odist.view(np.uint32))
    eng.close()


def test_encrypted_search_errors(pf, oracle):
    n, g = 2048, 16
    base, query, cent, offsets, ids, vecs = _dataset(9, nb=2000, nlist=8, nq=2)
    primes, t = _params(n)
    cl = OracleClient(oracle, n, primes, t, 128, 1, g)
    eng, _, _ = _engine(pf, n, g=g)
    eng.load_index(cent, offsets, ids, vecs)
    eng.set_list_sizes(offsets)
    cts = np.stack([cl.encrypt_query(q, 5 + i) for i, q in enumerate(query)])
    blob, offs = cl.serialize_queries(cts)
    idx = eng.coarse_quantize(query, 2)
    with pytest.raises(pf.PfError) as ei:   # no Galois keys yet
        eng.coarseSearchEncrypted(blob, offs, idx)
    assert ei.value.code == 4
    for i, key in enumerate(cl.step_keys()):
        eng.set_galois_key(eng.galois_elt(i + 1), key)
    bad = blob.copy()
    bad[5] = 2                              # compr_mode zstd
    with pytest.raises(pf.PfError) as ei:
        eng.coarseSearchEncrypted(bad, offs, idx)
    assert ei.value.code == 5
    bad = blob.copy()
    bad[48] = 1                             # is_ntt_form flag on a BFV query
    with pytest.raises(pf.PfError):
        eng.coarseSearchEncrypted(bad, offs, idx)
    with pytest.raises(pf.PfError) as ei:   # output buffer too small
        eng.coarseSearchEncrypted(blob, offs, idx, out=np.zeros(16, dtype=np.uint8))
    assert ei.value.code == 3
    res = eng.coarseSearchEncrypted(blob, offs, idx)
    assert res.stats["nresults"] > 0
    # an empty list contributes no result ciphertext
    offsets2 = offsets.copy()
    eng.close()
    # float-valued base vectors: plaintext stages load, encrypted search refuses
    eng, _, _ = _engine(pf, n, g=g)
    eng.load_index(cent, offsets, ids, vecs + np.float32(0.5))
    eng.set_list_sizes(offsets)
    d1, _, _ = eng.coarseSearch(query, idx)
    assert len(d1)
    with pytest.raises(pf.PfError):
        eng.coarseSearchEncrypted(blob, offs, idx)
    eng.close()
    # parameter validation
    with pytest.raises(pf.PfError):
        pf.Engine(128, 8192, [17, 19], 16760833)
    with pytest.raises(pf.PfError):
        pf.Engine(128, 8192, partial_g=3)


def test_cpp_host_mirror():
    """prefhetch::Server (C++ mirror of the reference Server class) over the same C ABI"""
    import subprocess
    from pathlib import Path
    exe = Path(__file__).resolve().parent.parent / "prefhetch_b200" / "host" / "pf_server_check"
    assert exe.exists(), "run __graft_entry__.build() first"
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("n,d,m,g,tbits,rl,macv", [(16384, 128, 1, 16, 24, 0, None), (8192, 960, 8, 8, 27, 0, None),
                                                   (8192, 960, 8, 8, 27, 1, "0"), (4096, 64, 2, 4, 24, 0, None),
                                                   (16384, 128, 1, 16, 24, 2, None), (4096, 64, 2, 4, 24, 1, None),
                                                   (8192, 256, 1, 8, 24, 1, None)])
def test_encrypted_search_other_shapes(pf, oracle, monkeypatch, n, d, m, g, tbits, rl, macv):
    """poly degree 16384 (L = 8, 49-bit primes: FP64 NTT with mid-pass reductions), GIST-shaped 960-d
    vectors with 8 query ciphertexts (K = 128 diagonals per block), and a 2-ciphertext small case:
    bytes identical to the oracle's pipeline, decrypted distances exact.  rl > 0: results mod-switched
    to rl limbs (8 -> 2 on the FP64 kernel with 49-bit primes; 2 -> 1 on the generic integer kernel).
    K = 128 runs the narrow-slice one-block-per-lane MAC (32 coefficients, 3 CTAs/SM) by default and the
    two-blocks-per-lane kernel with PF_MAC_VARIANT=0; d = 256 with one ciphertext gives K = 32 (128-wide)."""
    from oracle.pf_oracle import BATCHING_T, BFV_DEFAULT_PRIMES
    if macv is not None:
        monkeypatch.setenv("PF_MAC_VARIANT", macv)
    primes = BFV_DEFAULT_PRIMES[n]
    t = BATCHING_T[(n, tbits)] if (n, tbits) in BATCHING_T else ntt_primes(n, tbits, 1)[0]
    rng = np.random.default_rng(n + d)
    nlist, nq, nprobe = 3, 2, 2
    lay = oracle.LayoutPlan(n, d, m, g)
    nb = lay.C + lay.C // 3
    base, query, cent = sift_like(rng, nb, d, nlist, nq)
    offsets, ids, vecs = build_ivf(base, cent)
    eng = pf.Engine(d, n, primes, t, m, g, result_limbs=rl)
    eng.load_index(cent, offsets, ids, vecs)
    eng.set_list_sizes(offsets)
    cl = OracleClient(oracle, n, primes, t, d, m, g)
    keys = cl.step_keys()
    for i, key in enumerate(keys):
        eng.set_galois_key(eng.galois_elt(i + 1), key)
    cts = np.stack([cl.encrypt_query(q, 300 + 10 * i) for i, q in enumerate(query)])
    blob, offs = cl.serialize_queries(cts)
    idx = eng.coarse_quantize(query, nprobe)
    res = eng.coarseSearchEncrypted(blob, offs, idx)
    r = 0
    for qi in range(nq):
        rot = oracle.rotate_query_set(cl.ctx, cl.lay, cts[qi], keys, False)
        for l in idx[qi]:
            n_l = int(offsets[l + 1] - offsets[l])
            for b0 in range(0, n_l, lay.C):
                xs = vecs[offsets[l] + b0: offsets[l] + min(b0 + lay.C, n_l)].astype(np.int32)
                diag, norm = oracle.encode_block(cl.ctx, cl.lay, xs)
                want_ct = oracle.block_distance(cl.ctx, cl.lay, rot, diag, norm)
                pid = (0, 0, 0, 0)
                if rl:
                    want_ct = cl.mod_switch_to(want_ct, rl)
                    pid = _parms_id_py(n, primes[:rl], t)
                assert res.result(r) == cl.ctx.ct_save(want_ct, parms_id=pid), f"query {qi} result {r}"
                got_ct, _ = eng.ct_deserialize(res.result(r))
                dist, budget = cl.distances(got_ct, query[qi], len(xs))
                assert np.array_equal(dist, ((xs.astype(np.int64) - query[qi].astype(np.int64)) ** 2).sum(1))
                assert budget > 0
                r += 1
    assert r == res.stats["nresults"] and r > 0
    eng.close()


# ---------------------------------------------------------------------------------------------------------
# round 2: the benchmarked shape, the Galois-key loader, the kept kernel variants, submit / collect
# ---------------------------------------------------------------------------------------------------------
def _bench_shape_dataset(seed=5, nlist=256, d=128, C_=1024):
    """SIFT-shaped index with bench.py's structure: ~1 block per list, a few lists of two blocks."""
    rng = np.random.default_rng(seed)
    centres = rng.uniform(0, 160, size=(nlist, d))
    sizes = rng.integers(120, 200, size=nlist)
    sizes[[3, 77, 200]] = [C_ + 40, C_ + 1, 2 * C_]        # multi-block lists (one of them exactly 2 blocks)
    sizes[9] = 0                                           # an empty list
    assign = np.repeat(np.arange(nlist), sizes)
    base = np.clip(np.rint(centres[assign] + rng.normal(0, 24, size=(len(assign), d))), 0, 255).astype(np.float32)
    offsets = np.zeros(nlist + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    ids = rng.permutation(len(assign)).astype(np.int64)
    cent = np.stack([base[offsets[l]:offsets[l + 1]].mean(0) if sizes[l] else centres[l] for l in range(nlist)]).astype(np.float32)
    qa = rng.integers(0, nlist, size=64)
    query = np.clip(np.rint(centres[qa] + rng.normal(0, 24, size=(64, d))), 0, 255).astype(np.float32)
    return cent, offsets, ids, base, query


@pytest.mark.parametrize("rl", [1, 0])
def test_encrypted_search_bench_shape(pf, oracle, rl):
    """The shape bench.py times (64 queries x 16 probes, N = 8192, g = 8, ~1100 (query, block) pairs, real
    encryptions): key-switch query groups of 16, the 4 e2e query groups (5/10/14/16 sixteenths), blocks
    probed by several queries of the batch, multi-block lists.  EVERY result is decrypted and compared with
    the exact integer distances; a sample of >= 32 results spread over all query groups is compared byte
    for byte with the oracle pipeline (ref contract: src/server/server_lib.cpp:111-138)."""
    n, d, g, nprobe, nq = 8192, 128, 8, 16, 64
    primes, t = _params(n)
    cent, offsets, ids, vecs, query = _bench_shape_dataset()
    cl = OracleClient(oracle, n, primes, t, d, 1, g)
    keys = cl.step_keys()
    eng, _, _ = _engine(pf, n, g=g, result_limbs=rl)
    eng.load_index(cent, offsets, ids, vecs)
    eng.set_list_sizes(offsets)
    for i, key in enumerate(keys):
        eng.set_galois_key(eng.galois_elt(i + 1), key)
    cts = np.stack([cl.encrypt_query(q, 1000 + 3 * i) for i, q in enumerate(query)])
    blob, offs = cl.serialize_queries(cts)
    idx = eng.coarse_quantize(query, nprobe)
    idx[5, 2] = 3; idx[21, 0] = 3; idx[41, 15] = 77; idx[60, 7] = 200; idx[63, 1] = 9   # force the special lists in
    res = eng.coarseSearchEncrypted(blob, offs, idx)
    C_ = cl.lay.C
    blocks_of = (offsets[1:] - offsets[:-1] + C_ - 1) // C_
    assert res.stats["nresults"] == int(blocks_of[idx.reshape(-1)].sum()) > 900
    probes_per_list = np.bincount(idx.reshape(-1), minlength=len(cent))
    assert (probes_per_list > 1).sum() > 50            # blocks shared between queries of the batch
    # walk the response in order; decrypt everything, byte-compare the sample
    sample_q = {0, 4, 5, 15, 16, 19, 20, 21, 31, 32, 39, 40, 41, 47, 48, 55, 56, 60, 63}
    checked, r = 0, 0
    pid = _parms_id_py(n, primes[:rl], t) if rl else (0, 0, 0, 0)
    for qi in range(nq):
        rot = oracle.rotate_query_set(cl.ctx, cl.lay, cts[qi], keys, False) if qi in sample_q else None
        assert res.results_per_query[qi] == int(blocks_of[idx[qi]].sum())
        for p in range(nprobe):
            l = idx[qi, p]
            n_l = int(offsets[l + 1] - offsets[l])
            for b0 in range(0, n_l, C_):
                xs = vecs[offsets[l] + b0: offsets[l] + min(b0 + C_, n_l)].astype(np.int32)
                got = res.result(r)
                got_ct, is_ntt = eng.ct_deserialize(got)
                dist, budget = cl.distances(got_ct, query[qi], len(xs))
                assert np.array_equal(dist, ((xs.astype(np.int64) - query[qi].astype(np.int64)) ** 2).sum(1)), (qi, p, b0)
                assert budget > 0 and not is_ntt
                special = l in (3, 77, 200)
                if rot is not None and (p in (0, 7, 15) or special):
                    diag, norm = oracle.encode_block(cl.ctx, cl.lay, xs)
                    want_ct = oracle.block_distance(cl.ctx, cl.lay, rot, diag, norm)
                    if rl:
                        want_ct = cl.mod_switch_to(want_ct, rl)
                    assert got == cl.ctx.ct_save(want_ct, parms_id=pid), f"query {qi} probe {p} block {b0}"
                    checked += 1
                r += 1
    assert r == res.stats["nresults"] and checked >= 32
    eng.close()


def test_load_galois_keys_stream(pf, oracle):
    """pf_load_galois_keys: a SEAL GaloisKeys stream (compr_mode none and zlib) gives the same rotated query
    set as the raw-word setter and the oracle; malformed streams are refused."""
    from tests.util import galois_keys_save, zlib_stream
    n, d, g = 2048, 128, 16
    primes, t = _params(n)
    cl = OracleClient(oracle, n, primes, t, d, 1, g)
    keys = cl.step_keys()
    by_elt = {cl.ctx.galois_elt(i + 1): k for i, k in enumerate(keys)}
    blob = galois_keys_save(cl.ctx, by_elt)
    q = np.random.default_rng(2).integers(0, 256, size=d)
    cts = cl.encrypt_query(q, 9)
    want = oracle.rotate_query_set(cl.ctx, cl.lay, cts, keys, False)
    for stream in (blob, zlib_stream(blob)):
        eng, _, _ = _engine(pf, n, g=g)
        eng.load_galois_keys(stream)
        assert np.array_equal(eng.rotate_query_set(cts, False), want)
        eng.close()
    eng, _, _ = _engine(pf, n, g=g)
    with pytest.raises(pf.PfError) as ei:
        eng.load_galois_keys(blob[:len(blob) // 2])        # truncated
    assert ei.value.code == 5
    with pytest.raises(pf.PfError):
        eng.load_galois_keys(zlib_stream(blob)[:-9])       # truncated deflate stream
    bad = bytearray(blob)
    bad[16 + 32:16 + 40] = (n + 1).to_bytes(8, "little")    # more key slots than Galois elements exist
    with pytest.raises(pf.PfError):
        eng.load_galois_keys(bytes(bad))
    with pytest.raises(pf.PfError):                         # no keys loaded by any of the failed calls
        eng.rotate_query_set(cts, False)
    eng.close()


_VARIANTS = [{"PF_MAC_VARIANT": "0"}, {"PF_MAC_VARIANT": "4"}, {"PF_MAC_VARIANT": "5"}, {"PF_MAC_VARIANT": "6"},
             {"PF_NTT_FP": "0"}, {"PF_MS_INT": "1"}, {"PF_KS_NO_FUSED_PREP": "1"}, {"PF_MAC_NO_FPRED": "1"},
             {"PF_KS_NO_FPRED": "1"}, {"PF_FULL_SCRATCH_MB": "3"}]
# (the query-group count of a search — PF_E2E_GROUPS is read once per process — is exercised through
# pf_search_set_groups in test_submit_collect_pipelined)


@pytest.mark.parametrize("n,g,rl", [(8192, 8, 1), (16384, 16, 2)])
def test_kernel_variants_bit_identical(pf, oracle, monkeypatch, n, g, rl):
    """every kept non-default kernel (environment switches of DESIGN.md) gives the bytes of the default path,
    which is itself checked against the oracle: MAC variants 0/4/5/6, integer NTT at N >= 8192, integer
    mod-switch, un-fused mod-down prep, Barrett instead of FP64-assisted reductions, sub-batched result scratch."""
    d, nprobe, nq = 128, 3, 5
    primes, t = _params(n)
    lay = oracle.LayoutPlan(n, d, 1, g)
    rng = np.random.default_rng(n)
    base, query, cent = sift_like(rng, lay.C + 700, d, 4, nq)
    offsets, ids, vecs = build_ivf(base, cent)
    cl = OracleClient(oracle, n, primes, t, d, 1, g)
    keys = cl.step_keys()
    cts = np.stack([cl.encrypt_query(q, 40 + i) for i, q in enumerate(query)])
    blob, offs = cl.serialize_queries(cts)

    def run():
        eng, _, _ = _engine(pf, n, g=g, result_limbs=rl)     # PF_NTT_FP / *_NO_FPRED are read at creation
        eng.load_index(cent, offsets, ids, vecs)
        eng.set_list_sizes(offsets)
        for i, key in enumerate(keys):
            eng.set_galois_key(eng.galois_elt(i + 1), key)
        idx = eng.coarse_quantize(query, nprobe)
        res = eng.coarseSearchEncrypted(blob, offs, idx)
        out = [res.result(r) for r in range(res.stats["nresults"])]
        eng.close()
        return idx, out

    idx, ref = run()
    # the default path against the oracle (first result of every query)
    r = 0
    for qi in range(nq):
        rot = oracle.rotate_query_set(cl.ctx, cl.lay, cts[qi], keys, False)
        l = idx[qi, 0]
        xs = vecs[offsets[l]: offsets[l] + min(lay.C, int(offsets[l + 1] - offsets[l]))].astype(np.int32)
        diag, norm = oracle.encode_block(cl.ctx, cl.lay, xs)
        want = cl.mod_switch_to(oracle.block_distance(cl.ctx, cl.lay, rot, diag, norm), rl)
        assert ref[r] == cl.ctx.ct_save(want, parms_id=_parms_id_py(n, primes[:rl], t))
        r += sum(int((offsets[x + 1] - offsets[x] + lay.C - 1) // lay.C) for x in idx[qi])
    for env in _VARIANTS:
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        try:
            _, got = run()
        except Exception as ex:
            raise AssertionError(f"variant {env} failed: {ex}") from ex
        for k in env:
            monkeypatch.delenv(k)
        assert len(got) == len(ref) and all(a == b for a, b in zip(got, ref)), f"variant {env} differs from the default path"


def test_submit_collect_pipelined(pf, oracle):
    """pf_search_submit / pf_search_collect: searches in flight give the bytes of the one-call form; a
    fifth submit is refused with PF_ERR_STATE until one is collected; an unknown ticket is refused."""
    n, g, d, nprobe = 2048, 16, 128, 4
    base, query, cent, offsets, ids, vecs = _dataset(31, nb=4000, nlist=16, nq=6)
    primes, t = _params(n)
    cl = OracleClient(oracle, n, primes, t, d, 1, g)
    eng, _, _ = _engine(pf, n, g=g)
    eng.load_index(cent, offsets, ids, vecs)
    eng.set_list_sizes(offsets)
    for i, key in enumerate(cl.step_keys()):
        eng.set_galois_key(eng.galois_elt(i + 1), key)
    batches = []
    for b in range(3):
        qs = query[2 * b: 2 * b + 2]
        cts = np.stack([cl.encrypt_query(q, 500 + 10 * b + i) for i, q in enumerate(qs)])
        blob, offs = cl.serialize_queries(cts)
        idx = eng.coarse_quantize(qs, nprobe)
        one = eng.coarseSearchEncrypted(blob, offs, idx)
        batches.append((blob, offs, idx, [one.result(r) for r in range(one.stats["nresults"])], one.labels.copy()))
    for groups in (0, 1):
        eng.set_search_groups(groups)
        p0 = eng.submitSearchEncrypted(*batches[0][:3])
        p1 = eng.submitSearchEncrypted(*batches[1][:3])
        extra = [eng.submitSearchEncrypted(*batches[2][:3]) for _ in range(2)]   # 4 in flight: the limit
        with pytest.raises(pf.PfError) as ei:
            eng.submitSearchEncrypted(*batches[2][:3])
        assert ei.value.code == 4
        r0 = p0.collect()
        for x in extra:
            x.collect()
        p2 = eng.submitSearchEncrypted(*batches[2][:3])
        for pend, (_, _, _, want, labels) in zip((p0, p1, p2), batches):
            res = pend.collect()
            assert [res.result(r) for r in range(res.stats["nresults"])] == want
            assert np.array_equal(res.labels, labels)
        assert r0 is p0.result
    assert eng.lib.pf_search_collect(eng.h, 12345) == 4
    eng.close()


def test_hostile_inputs(pf, oracle):
    """offsets outside the blob, a deflate bomb in place of a query ciphertext, a peer flag that
    never arrives — each is an error code, never an out-of-bounds read, unbounded allocation or a hang."""
    import struct
    import zlib
    n, g = 2048, 16
    base, query, cent, offsets, ids, vecs = _dataset(9, nb=2000, nlist=8, nq=2)
    primes, t = _params(n)
    cl = OracleClient(oracle, n, primes, t, 128, 1, g)
    eng, _, _ = _engine(pf, n, g=g)
    eng.load_index(cent, offsets, ids, vecs)
    eng.set_list_sizes(offsets)
    for i, key in enumerate(cl.step_keys()):
        eng.set_galois_key(eng.galois_elt(i + 1), key)
    cts = np.stack([cl.encrypt_query(q, 5 + i) for i, q in enumerate(query)])
    blob, offs = cl.serialize_queries(cts)
    idx = eng.coarse_quantize(query, 2)
    for bad_offs in ([0, offs[1], offs[2] + 8], [0, offs[2], offs[1]], [offs[1], 0, offs[2]], [0, 2 ** 63, 2 ** 63 + 5]):
        with pytest.raises(pf.PfError) as ei:
            eng.coarseSearchEncrypted(blob, np.array(bad_offs, dtype=np.uint64), idx)
        assert ei.value.code == 1
    # 200 MB of zeros deflate to ~200 KB: must be refused at the size of one ciphertext, not inflated
    bomb_body = zlib.compress(bytes(200 << 20), 9)
    bomb = bytes(blob[:5]) + b"\x01" + bytes(blob[6:8]) + struct.pack("<Q", 16 + len(bomb_body)) + bomb_body
    with pytest.raises(pf.PfError) as ei:
        eng.ct_deserialize(bomb)
    assert ei.value.code == 5
    zb = np.frombuffer(bomb + bytes(blob[offs[1]:offs[2]]), dtype=np.uint8)
    zo = np.array([0, len(bomb), len(zb)], dtype=np.uint64)
    with pytest.raises(pf.PfError) as ei:
        eng.coarseSearchEncrypted(zb, zo, idx)
    assert ei.value.code == 5
    res = eng.coarseSearchEncrypted(blob, offs, idx)       # the engine still works
    assert res.stats["nresults"] > 0
    eng.close()


def test_flag_wait_times_out(pf, monkeypatch):
    """a peer flag that never arrives: the wait kernel gives up after PF_FLAG_TIMEOUT_MS and the engine reports
    PF_ERR_CUDA from then on instead of hanging the stream"""
    monkeypatch.setenv("PF_FLAG_TIMEOUT_MS", "200")
    eng, _, _ = _engine(pf, 2048)
    ptr, _h = eng.ipc_alloc(128)
    eng.flag_write(ptr, 3)
    eng.flag_wait(ptr, 3)          # already there: returns at once
    eng.synchronize()
    words = np.arange(16, dtype=np.uint64)
    assert eng.device_checksum(ptr, 0) == 0
    eng.flag_wait(ptr, 4)          # never written
    with pytest.raises(pf.PfError) as ei:
        eng.synchronize()
    assert ei.value.code == 2 and "timed out" in str(ei.value)
    with pytest.raises(pf.PfError):
        eng.flag_write(ptr, 5)     # sticky
    eng.ipc_free(ptr)
    eng.close()


def _write_cpp_case(tmp_path, cl, keys, d, n, g, rl, nprobe, query, t, primes, cent, offsets, ids, vecs, blob, offs, idx):
    """the files host/pf_server_check.cpp reads (its header lists them)"""
    from tests.util import galois_keys_save
    (tmp_path / "params.txt").write_text(" ".join(str(x) for x in [d, n, g, 1, rl, nprobe, len(query), t, len(primes), *primes]))
    cent.astype(np.float32).tofile(tmp_path / "centroids.f32")
    offsets.astype(np.int64).tofile(tmp_path / "offsets.i64")
    ids.astype(np.int64).tofile(tmp_path / "ids.i64")
    vecs.astype(np.float32).tofile(tmp_path / "vectors.f32")
    (tmp_path / "galois_keys.bin").write_bytes(galois_keys_save(cl.ctx, {cl.ctx.galois_elt(i + 1): k for i, k in enumerate(keys)}))
    blob.tofile(tmp_path / "queries.bin")
    offs.astype(np.uint64).tofile(tmp_path / "query_offsets.u64")
    idx.astype(np.int64).tofile(tmp_path / "probes.i64")
    np.ascontiguousarray(query, dtype=np.float32).tofile(tmp_path / "queries.f32")


def test_cpp_encrypted_search_matches_python(pf, oracle, tmp_path):
    """the reference-side binding in C++ (prefhetch::Server: loadGaloisKeys from a SEAL stream, result_limbs,
    coarseSearchEncrypted and submit / collect) gives the bytes of the Python path, which is checked against
    the oracle here as well (the encrypted call was never made from C++)"""
    import subprocess
    from pathlib import Path
    from tests.util import galois_keys_save
    exe = Path(__file__).resolve().parent.parent / "prefhetch_b200" / "host" / "pf_server_check"
    assert exe.exists(), "run __graft_entry__.build() first"
    n, g, d, nprobe, rl = 2048, 16, 128, 3, 1
    base, query, cent, offsets, ids, vecs = _dataset(41, nb=3000, nlist=12, nq=3)
    primes, t = _params(n)
    cl = OracleClient(oracle, n, primes, t, d, 1, g)
    keys = cl.step_keys()
    cts = np.stack([cl.encrypt_query(q, 700 + i) for i, q in enumerate(query)])
    blob, offs = cl.serialize_queries(cts)
    eng, _, _ = _engine(pf, n, g=g, result_limbs=rl)
    eng.load_index(cent, offsets, ids, vecs)
    eng.set_list_sizes(offsets)
    for i, key in enumerate(keys):
        eng.set_galois_key(eng.galois_elt(i + 1), key)
    idx = eng.coarse_quantize(query, nprobe)
    res = eng.coarseSearchEncrypted(blob, offs, idx)
    want = b"".join(res.result(r) for r in range(res.stats["nresults"]))
    # oracle check of the first result
    rot = oracle.rotate_query_set(cl.ctx, cl.lay, cts[0], keys, False)
    l = idx[0, 0]
    xs = vecs[offsets[l]: offsets[l] + min(cl.lay.C, int(offsets[l + 1] - offsets[l]))].astype(np.int32)
    diag, norm = oracle.encode_block(cl.ctx, cl.lay, xs)
    first = cl.ctx.ct_save(cl.mod_switch_to(oracle.block_distance(cl.ctx, cl.lay, rot, diag, norm), rl),
                           parms_id=_parms_id_py(n, primes[:rl], t))
    assert res.result(0) == first
    labels = res.labels.copy()
    eng.close()
    # the same request from C++
    _write_cpp_case(tmp_path, cl, keys, d, n, g, rl, nprobe, query, t, primes, cent, offsets, ids, vecs, blob, offs, idx)
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "results.bin").read_bytes() == want
    assert np.array_equal(np.fromfile(tmp_path / "labels.i64", dtype=np.int64), labels)


def test_cpp_handlers_end_to_end(pf, oracle, tmp_path):
    """the handler bodies of host/pf_query_handlers.hpp (ref: src/server/controllers/Query.cc:10-98 + the additive
    encrypted endpoint): JSON request bodies in the reference's shape in, JSON out, equal to direct calls on the
    same prefhetch::Server; run by `pf_server_check --handlers` on files written here"""
    import subprocess
    from pathlib import Path
    exe = Path(__file__).resolve().parent.parent / "prefhetch_b200" / "host" / "pf_server_check"
    assert exe.exists(), "run __graft_entry__.build() first"
    n, g, d, nprobe, rl = 2048, 16, 128, 3, 1
    base, query, cent, offsets, ids, vecs = _dataset(61, nb=3000, nlist=12, nq=3, frac_centroids=False)
    primes, t = _params(n)
    cl = OracleClient(oracle, n, primes, t, d, 1, g)
    keys = cl.step_keys()
    cts = np.stack([cl.encrypt_query(q, 800 + i) for i, q in enumerate(query)])
    blob, offs = cl.serialize_queries(cts)
    oidx, _ = oracle.coarse_quantize(query, cent, nprobe)
    _write_cpp_case(tmp_path, cl, keys, d, n, g, rl, nprobe, query, t, primes, cent, offsets, ids, vecs, blob, offs, oidx)
    r = subprocess.run([str(exe), "--handlers", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "handlers ok" in r.stdout, r.stdout + r.stderr


def test_zstd_compressed_streams(pf, oracle):
    """SEAL compr_mode zstd on the request side (f-3; SEAL's default when built with zstd): query ciphertexts, single
    ciphertexts and GaloisKeys saved as Zstandard frames (one-shot and streamed, made by pyarrow's bundled libzstd)
    give the same bytes out as their uncompressed form.  The decoder itself is covered on the CPU (tests/test_abi.py)."""
    from tests.util import galois_keys_save, have_zstd, zstd_stream
    if not have_zstd():
        pytest.skip("pyarrow without the zstd codec: no independent encoder to make vectors with")
    n, g, d, nprobe = 2048, 16, 128, 3
    base, query, cent, offsets, ids, vecs = _dataset(23, nb=3000, nlist=16, nq=3)
    primes, t = _params(n)
    cl = OracleClient(oracle, n, primes, t, d, 1, g)
    keys = cl.step_keys()
    cts = np.stack([cl.encrypt_query(q, 400 + i) for i, q in enumerate(query)])
    blob, offs = cl.serialize_queries(cts)
    sparts = [zstd_stream(bytes(blob[offs[i]:offs[i + 1]]), streaming=bool(i & 1)) for i in range(len(cts))]
    sblob = np.frombuffer(b"".join(sparts), dtype=np.uint8)
    soffs = np.concatenate([[0], np.cumsum([len(z) for z in sparts])]).astype(np.uint64)
    eng, _, _ = _engine(pf, n, g=g)
    eng.load_index(cent, offsets, ids, vecs)
    eng.set_list_sizes(offsets)
    kblob = galois_keys_save(cl.ctx, {cl.ctx.galois_elt(i + 1): k for i, k in enumerate(keys)})
    eng.load_galois_keys(zstd_stream(kblob, streaming=True))
    idx = eng.coarse_quantize(query, nprobe)
    plain = eng.coarseSearchEncrypted(blob, offs, idx)
    want = [plain.result(r) for r in range(plain.stats["nresults"])]
    assert want
    comp = eng.coarseSearchEncrypted(sblob, soffs, idx)
    assert comp.stats["nresults"] == len(want) and all(comp.result(r) == want[r] for r in range(len(want)))
    # the keys that came in as a zstd stream rotate like the raw-word ones
    assert np.array_equal(eng.rotate_query_set(cts[0], False), oracle.rotate_query_set(cl.ctx, cl.lay, cts[0], keys, False))
    got, is_ntt = eng.ct_deserialize(sparts[1])
    assert np.array_equal(got, cts[1][0]) and not is_ntt
    with pytest.raises(pf.PfError):
        eng.ct_deserialize(sparts[0][:-20])          # truncated frame
    with pytest.raises(pf.PfError):
        eng.load_galois_keys(zstd_stream(kblob)[:-9])
    eng.close()


@pytest.mark.parametrize("n,m,g,rl", [(2048, 1, 16, 1), (8192, 1, 8, 1), (2048, 2, 16, 0)])
def test_cpp_client_round_trip(pf, tmp_path, n, m, g, rl):
    """f-4, no oracle in the loop: the C++ client (host/pf_client.hpp) makes the secret key, the GaloisKeys stream and
    seeded query ciphertexts; the engine loads them as any SEAL client's and answers; the client decrypts the response
    into the reference's packed coarse scores — the exact squared L2 of every candidate of every probed list — and
    ranks them like client_lib.cpp:122-156.  (The same client against the oracle as the server: tests/test_client.py.)"""
    import subprocess
    from pathlib import Path
    from tests.test_client import write_case
    exe = Path(__file__).resolve().parent.parent / "prefhetch_b200" / "host" / "pf_client_check"
    assert exe.exists(), "run __graft_entry__.build() first"
    d, nprobe, coarse_probe = 128, 3, 10
    base, query, cent, offsets, ids, vecs = _dataset(61 + n, nb=2500, nlist=10, nq=3)
    primes, t = _params(n)
    write_case(tmp_path, n, primes, t, d, m, g, query.astype(np.int64), nprobe, coarse_probe, np.random.default_rng(n).bytes(64))
    r = subprocess.run([str(exe), "keygen", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok keygen"), r.stdout + r.stderr
    eng = pf.Engine(d, n, primes, t, m, g, result_limbs=rl)
    eng.load_index(cent, offsets, ids, vecs)
    eng.set_list_sizes(offsets)
    # the keys travel seeded (Serializable<GaloisKeys>); one case loads the full stream of the same key set instead
    eng.load_galois_keys((tmp_path / ("galois_keys_full.bin" if m == 2 else "galois_keys.bin")).read_bytes())
    idx = eng.coarse_quantize(query, nprobe)
    blob = np.fromfile(tmp_path / "queries_seeded.bin", dtype=np.uint8)
    offs = np.fromfile(tmp_path / "queries_seeded.off", dtype=np.uint64)
    res = eng.coarseSearchEncrypted(blob, offs, idx)
    nres = res.stats["nresults"]
    assert nres == int(res.results_per_query.sum()) and nres > 0
    streams = [res.result(i) for i in range(nres)]
    (tmp_path / "results.bin").write_bytes(b"".join(streams))
    np.concatenate([[0], np.cumsum([len(s) for s in streams])]).astype(np.uint64).tofile(tmp_path / "results.off")
    res.probed_sizes.astype(np.uint64).tofile(tmp_path / "probed_sizes.u64")
    res.results_per_query.astype(np.uint64).tofile(tmp_path / "results_per_query.u64")
    np.ascontiguousarray(res.labels, dtype=np.int64).tofile(tmp_path / "labels.i64")
    r = subprocess.run([str(exe), "decrypt", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok decrypt"), r.stdout + r.stderr
    scores = np.fromfile(tmp_path / "scores.f32", dtype=np.float32)
    sizes = np.fromfile(tmp_path / "list_sizes.u64", dtype=np.uint64)
    assert np.array_equal(sizes.astype(np.int64), res.list_sizes)
    want, labels = [], []
    for qi in range(len(query)):
        for l in idx[qi]:
            xs = vecs[offsets[l]:offsets[l + 1]].astype(np.int64)
            want.append(((xs - query[qi].astype(np.int64)) ** 2).sum(1))
            labels.append(ids[offsets[l]:offsets[l + 1]])
    want, labels = np.concatenate(want), np.concatenate(labels)
    assert np.array_equal(np.asarray(res.labels), labels)
    assert np.array_equal(scores.astype(np.int64), want)
    assert int((tmp_path / "budget.txt").read_text()) > 0
    nearest = np.fromfile(tmp_path / "nearest.i64", dtype=np.int64).reshape(len(query), coarse_probe)
    for qi in range(len(query)):
        a, b = int(sizes[:qi].sum()), int(sizes[:qi + 1].sum())
        assert np.array_equal(nearest[qi], labels[a:b][np.argsort(want[a:b], kind="stable")[:coarse_probe]])
    # the same scores through the plaintext endpoint of the reference (Server::coarseSearch)
    pdist, plabels, psizes = eng.coarseSearch(query, idx)
    assert np.array_equal(np.asarray(pdist, dtype=np.float32), scores) and np.array_equal(plabels, labels)
    # a ciphertext stamped for other encryption parameters is refused, as seal::Ciphertext::load(context, ...) would
    other = blob.copy()
    other[int(offs[1]) + 16] ^= 0x40
    with pytest.raises(pf.PfError) as ei:
        eng.coarseSearchEncrypted(other, offs, idx)
    assert ei.value.code == 5 and "parameters" in str(ei.value)
    eng.close()


def test_plain_stages_against_reference_golden(pf):
    """The CUDA plaintext stages against outputs of the REFERENCE ITSELF (tests/golden/ref_plain_v1.json: the
    reference's sort_nearest_centroids and Server::preciseSearch compiled from its own sources and run on these
    inputs — tests/test_ref_pin.py has the CPU side): pf_coarse_quantize returns the reference's order and the
    reference's float bits for all centroids, pf_precise_search the reference's float bits, on integer and on
    fractional data (where the float / double accumulation order shows in the low bits)."""
    from tests.golden.make_golden_ref import ref_inputs
    from tests.test_ref_pin import GOLD, NQ, CP, f32, same_up_to_ties
    x = ref_inputs()
    cent = x["cent"]
    nlist = len(cent)
    for tag in ("int", "frac"):
        base, query = x[f"base_{tag}"], x[f"query_{tag}"]
        offsets, ids, vecs = build_ivf(base, cent)
        eng, _, _ = _engine(pf, 2048)
        eng.load_index(cent, offsets, ids, vecs)        # fractional data: plaintext stages only (no encrypted index)
        g = GOLD[f"sort_nearest_centroids_{tag}"]
        want_idx, want_dist = np.array(g["idx"]).reshape(NQ, nlist), f32(g["dist_bits"]).reshape(NQ, nlist)
        idx, dist = eng.coarse_quantize(query, nlist, return_dist=True)
        assert np.array_equal(dist.view(np.uint32), want_dist.view(np.uint32)), tag
        assert all(same_up_to_ties(idx[i], want_idx[i], want_dist[i]) for i in range(NQ)) and np.array_equal(idx, want_idx), tag
        idx20 = eng.coarse_quantize(query, 20)          # the reference's NPROBE
        assert np.array_equal(idx20, want_idx[:, :20]), tag
        want = np.array(GOLD[f"precise_search_{tag}"]["score_bits"], dtype=np.uint32).reshape(NQ, CP)
        got = eng.preciseSearch(query, x["ids"])
        assert np.array_equal(got.view(np.uint32), want), tag
        eng.close()


def test_cpp_roundtrip_example():
    """host/pf_roundtrip_example.cpp: the reference's client main (src/client/client.cpp) with the coarse step
    encrypted — C++ client, handler bodies with their JSON, prefhetch::Server over the C ABI — in one process, no
    oracle and no Python in the loop: decrypted scores equal the plaintext endpoint's, the re-ranked distances equal
    exact brute force"""
    import subprocess
    from pathlib import Path
    exe = Path(__file__).resolve().parent.parent / "prefhetch_b200" / "host" / "pf_roundtrip_example"
    assert exe.exists(), "run __graft_entry__.build() first"
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "pf_roundtrip_example ok" in r.stdout, r.stdout + r.stderr


def _pq_case(oracle, seed, nb, d, nlist, nq, M):
    """IVF data set + a product quantizer on the residuals (sub-centroids sampled from the residuals themselves: what
    k-means would start from; the arithmetic under test does not care how the codebook was trained)"""
    rng = np.random.default_rng(seed)
    base, query, cent = sift_like(rng, nb, d, nlist, nq)
    cent = (cent + rng.uniform(-0.5, 0.5, size=cent.shape)).astype(np.float32)
    offsets, ids, vecs = build_ivf(base, cent)
    list_of = np.repeat(np.arange(nlist), np.diff(offsets))
    resid = vecs - cent[list_of]
    dsub = d // M
    pick = rng.integers(0, len(vecs), size=(M, 256))
    pqc = np.stack([resid[pick[m], m * dsub:(m + 1) * dsub] for m in range(M)]).astype(np.float32)      # [M][256][dsub]
    pqc += rng.normal(0, 0.25, size=pqc.shape).astype(np.float32)
    codes = oracle.pq_encode_residuals(vecs, offsets, cent, M, pqc)
    return query, cent, offsets, ids, vecs, pqc, codes


# The tests from here on are about kernels written after the round's GPU budget ended (seeded expansion on the device,
# PQ-ADC): they run last so that, under -x, a failure in them hides as little as possible of what is above
# (test_cpp_client_round_trip also sends seeded requests, from the C++ client).
def test_seeded_query_ciphertexts(pf, oracle, monkeypatch):
    """f-3: seeded query streams (Serializable<Ciphertext>: c0 + the PRNG seed of c1), plain and zlib, give the result
    bytes of the same ciphertexts sent in full; pf_ct_deserialize expands them too; a shake256 seed is refused"""
    from tests.util import zlib_stream
    n, g, d, nprobe = 2048, 16, 128, 3
    base, query, cent, offsets, ids, vecs = _dataset(51, nb=3000, nlist=12, nq=3)
    primes, t = _params(n)
    cl = OracleClient(oracle, n, primes, t, d, 1, g)
    rng = np.random.default_rng(8)
    seeds = [rng.bytes(64) for _ in query]
    cts = [cl.ctx.encrypt_seeded(cl.sk, cl.ctx.encode(cl.lay.query_slots(t, q.astype(np.int64), 0)), 900 + i, seeds[i])
           for i, q in enumerate(query)]
    full = [cl.ctx.ct_save(ct) for ct in cts]
    seeded = [cl.ctx.ct_save_seeded(ct, s) for ct, s in zip(cts, seeds)]
    assert all(len(a) < 0.51 * len(b) + 200 for a, b in zip(seeded, full))

    def blob_of(parts):
        return (np.frombuffer(b"".join(parts), dtype=np.uint8).copy(),
                np.concatenate([[0], np.cumsum([len(x) for x in parts])]).astype(np.uint64))
    eng, _, _ = _engine(pf, n, g=g)
    eng.load_index(cent, offsets, ids, vecs)
    eng.set_list_sizes(offsets)
    for i, key in enumerate(cl.step_keys()):
        eng.set_galois_key(eng.galois_elt(i + 1), key)
    idx = eng.coarse_quantize(query, nprobe)
    ref = eng.coarseSearchEncrypted(*blob_of(full), idx)
    want = [ref.result(r) for r in range(ref.stats["nresults"])]
    assert want
    for parts in (seeded, [zlib_stream(x) for x in seeded], [seeded[0], full[1], zlib_stream(seeded[2])]):
        got = eng.coarseSearchEncrypted(*blob_of(parts), idx)
        assert [got.result(r) for r in range(got.stats["nresults"])] == want
    # an all-seeded uncompressed request is expanded ON THE DEVICE (3 more launches per query group than the host
    # expansion needs for its header strip: memset aside, strip + expand + fix-up instead of strip); PF_SEEDED_HOST=1
    # keeps it on the host; same bytes either way
    l0 = eng.launch_count()
    eng.coarseSearchEncrypted(*blob_of(seeded), idx)
    l1 = eng.launch_count()
    monkeypatch.setenv("PF_SEEDED_HOST", "1")
    got = eng.coarseSearchEncrypted(*blob_of(seeded), idx)
    l2 = eng.launch_count()
    monkeypatch.delenv("PF_SEEDED_HOST")
    assert [got.result(r) for r in range(got.stats["nresults"])] == want and (l1 - l0) > (l2 - l1)
    # decrypted distances of the first result are exact (the seeded encryption is a valid one)
    l = idx[0, 0]
    xs = vecs[offsets[l]: offsets[l] + min(cl.lay.C, int(offsets[l + 1] - offsets[l]))].astype(np.int32)
    ct0, _ = eng.ct_deserialize(want[0])
    dist, budget = cl.distances(ct0, query[0], len(xs))
    assert np.array_equal(dist, ((xs.astype(np.int64) - query[0].astype(np.int64)) ** 2).sum(1)) and budget > 0
    back, is_ntt = eng.ct_deserialize(seeded[1])
    assert np.array_equal(back, cts[1]) and not is_ntt
    with pytest.raises(pf.PfError) as ei:
        eng.coarseSearchEncrypted(*blob_of([cl.ctx.ct_save_seeded(cts[0], seeds[0], prng_type=2), full[1], full[2]]), idx)
    assert ei.value.code == 5
    eng.close()


@pytest.mark.parametrize("d,M,nlist", [(128, 32, 24), (64, 8, 9), (128, 16, 5)])
def test_pq_adc_matches_oracle(pf, oracle, tmp_path, d, M, nlist):
    """the distance the reference's FAISS fork computes today (PQ-ADC over every code of the
    given lists) through pf_search_lists_pq == the oracle's restatement, float bit patterns and labels, with the
    quantizer read from a .faiss IndexIVFPQ file (128-d / 32 x 8 bits is the reference's shape: SUB_QUANTIZERS,
    SUB_QUANTIZER_SIZE); ragged lists, an empty list, M not a multiple of 16."""
    from prefhetch_b200 import faiss_io
    query, cent, offsets, ids, vecs, pqc, codes = _pq_case(oracle, 100 + M, 2500, d, nlist, 5, M)
    # empty the last list (its vectors go nowhere: a list FAISS never added to)
    keep = int(offsets[-2])
    offsets = offsets.copy()
    offsets[-1] = keep
    ids, vecs, codes = ids[:keep], vecs[:keep], codes[:keep]
    lists = [ids[offsets[l]:offsets[l + 1]] for l in range(nlist)]
    f = faiss_io.IVFPQFile(d, len(ids), nlist, 20, cent, lists, [codes[offsets[l]:offsets[l + 1]] for l in range(nlist)],
                           code_size=M, pq_M=M, pq_nbits=8, pq_centroids=pqc.reshape(-1))
    faiss_io.write_ivfpq(str(tmp_path / "pq.faiss"), f)
    base = np.zeros((int(ids.max()) + 1, d), dtype=np.float32)
    base[ids] = vecs
    eng = pf.Engine(d, 2048, *_params(2048)[:1], _params(2048)[1], 1, 16)
    with pytest.raises(pf.PfError) as ei:
        eng.coarseSearchPQ(query, np.zeros((len(query), 1), np.int64))
    assert ei.value.code == 4          # PF_ERR_STATE: no index
    eng.load_index(cent, offsets, ids, vecs)
    with pytest.raises(pf.PfError) as ei:
        eng.coarseSearchPQ(query, np.zeros((len(query), 1), np.int64))
    assert ei.value.code == 4 and "pf_load_pq" in str(ei.value)
    with pytest.raises(pf.PfError):
        eng.load_pq(7, 8, pqc, np.zeros((len(ids), 7), np.uint8))          # 7 does not divide d
    eng.load_index_from_faiss(str(tmp_path / "pq.faiss"), base)            # loads the quantizer with the index
    nprobe = min(6, nlist)
    idx = eng.coarse_quantize(query, nprobe)
    idx[0, 0] = nlist - 1                                                  # the empty list is probed too
    dist, labels, sizes = eng.coarseSearchPQ(query, idx)
    odist, olabels, osizes = oracle.search_lists_pq(query, idx, cent, offsets, ids, M, pqc, codes)
    assert np.array_equal(sizes, osizes) and np.array_equal(labels, olabels)
    assert np.array_equal(dist.view(np.uint32), odist.view(np.uint32))
    # ADC is the exact squared L2 to the DECODED vector: float64 evaluation agrees to float rounding
    list_of = np.repeat(np.arange(nlist), np.diff(offsets))
    dsub = d // M
    dec = cent[list_of].astype(np.float64) + np.concatenate([pqc[m, codes[:, m]] for m in range(M)], axis=1).astype(np.float64)
    pos = {int(i): k for k, i in enumerate(ids)}
    o = 0
    for qi in range(len(query)):
        n = int(sizes[qi])
        rows = np.array([pos[int(i)] for i in labels[o:o + n]], dtype=np.int64)
        want = ((dec[rows] - query[qi].astype(np.float64)) ** 2).sum(1) if n else np.zeros(0)
        assert np.allclose(dist[o:o + n], want, rtol=2e-5, atol=1e-2)
        o += n
    # a new index drops the quantizer of the old one
    eng.load_index(cent, offsets, ids, vecs)
    with pytest.raises(pf.PfError):
        eng.coarseSearchPQ(query, idx)
    eng.close()


@pytest.mark.parametrize("n,bits", [(2048, 40), (8192, None), (1024, 57)])
def test_seeded_expansion_on_the_device(pf, oracle, n, bits):
    """f-3: the device-side expansion of a seeded ciphertext (Blake2xb leaves in parallel, rejected words re-drawn in
    stream order) == the host expansion == the oracle's, word for word.  57-bit primes make a draw exceed the largest
    multiple of q below 2^64 with probability ~ 2^-7, so the re-draw path runs (dozens of times per ciphertext,
    including re-draws of re-draws); BFVDefault primes almost never take it."""
    from tests.util import ntt_primes
    if bits is None:
        primes, t = _params(n)
    else:
        primes, t = ntt_primes(n, bits, 3) + ntt_primes(n, bits + 1, 1), ntt_primes(n, 20, 1)[0]
        if bits == 57:   # far from a power of two: 2^64 mod q is a large fraction of q
            from tests.util import is_prime
            primes, c = [], (3 << 55) // (2 * n) * (2 * n) + 1
            while len(primes) < 4:
                if is_prime(c):
                    primes.append(c)
                c -= 2 * n
    L = len(primes) - 1
    ctx = oracle.Context(n, primes, t)
    sk = ctx.keygen(77)
    rng = np.random.default_rng(n + (bits or 0))
    eng = pf.Engine(128, n, primes, t, 1, 16 if n <= 2048 else 8)
    rejected_total = 0
    for trial in range(3):
        seed = rng.bytes(64)
        pt = ctx.encode(rng.integers(0, t, size=n))
        ct = ctx.encrypt_seeded(sk, pt, 500 + trial, seed)
        stream = ctx.ct_save_seeded(ct, seed)
        got = eng.ct_expand_seeded_device(stream)
        assert np.array_equal(got, ct), f"trial {trial}: device expansion differs from the oracle's ciphertext"
        back, is_ntt = eng.ct_deserialize(stream)          # the host expansion (pf_seal_prng.h)
        assert np.array_equal(back, got) and not is_ntt
        # how many words of the first L*n stream words were rejected (test bookkeeping, from the oracle's PRNG)
        stream_bytes = b"".join(oracle.blake2xb(4096, c.to_bytes(8, "little"), seed) for c in range(L * n * 8 // 4096))
        words = np.frombuffer(stream_bytes, dtype=np.uint64).reshape(L, n)
        for j in range(L):
            mm = (2**64 - 1) - ((2**64 - 1) % primes[j]) - 1
            rejected_total += int((words[j] >= np.uint64(mm)).sum())
    if bits == 57:
        assert 10 < rejected_total <= 3 * 96, rejected_total       # the re-draw path ran, within the list's capacity
    # refusals: a full stream, a shake256 seed, a truncated stream
    with pytest.raises(pf.PfError):
        eng.ct_expand_seeded_device(ctx.ct_save(ct))
    with pytest.raises(pf.PfError):
        eng.ct_expand_seeded_device(ctx.ct_save_seeded(ct, seed, prng_type=2))
    with pytest.raises(pf.PfError):
        eng.ct_expand_seeded_device(stream[:-1])
    eng.close()
