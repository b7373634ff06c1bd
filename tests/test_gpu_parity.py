"""GPU parity: every C-ABI entry point of libprefhetch_b200.so against the CPU oracle on the same
seeded inputs — bit-exact for all integer / ciphertext work, bit-exact float bits for the
plaintext distances (the kernels restate the reference's float/double arithmetic)."""
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
        eng.ct_deserialize(zparts[0][:5] + b"\x02" + zparts[0][6:])   # zstd: not available
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


@pytest.mark.parametrize("n,d,m,g,tbits,rl", [(16384, 128, 1, 16, 24, 0), (8192, 960, 8, 8, 27, 0), (4096, 64, 2, 4, 24, 0),
                                              (16384, 128, 1, 16, 24, 2), (4096, 64, 2, 4, 24, 1)])
def test_encrypted_search_other_shapes(pf, oracle, n, d, m, g, tbits, rl):
    """poly degree 16384 (L = 8, 49-bit primes: FP64 NTT with mid-pass reductions), GIST-shaped 960-d
    vectors with 8 query ciphertexts (K = 128 diagonals per block), and a 2-ciphertext small case:
    bytes identical to the oracle's pipeline, decrypted distances exact.  rl > 0: results mod-switched
    to rl limbs (8 -> 2 on the FP64 kernel with 49-bit primes; 2 -> 1 on the generic integer kernel)."""
    from oracle.pf_oracle import BATCHING_T, BFV_DEFAULT_PRIMES
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
