"""Math-level (convention-free) checks of the oracle's BFV path: decrypt correctness, slot
semantics of multiply_plain / rotate_rows, the distance layout, and the whole encrypted pipeline
against the plaintext reference path."""
import numpy as np
import pytest

from tests.util import build_ivf, sift_like


def _centered(v, t):
    v = v.astype(np.int64)
    return np.where(v > t // 2, v - t, v)


def test_encrypt_decrypt_roundtrip(oracle, toy_ctx):
    ctx = toy_ctx
    rng = np.random.default_rng(1)
    sk = ctx.keygen(7)
    vals = rng.integers(0, ctx.t, size=ctx.n, dtype=np.uint64)
    ct = ctx.encrypt(sk, ctx.encode(vals), 11)
    plain, budget = ctx.decrypt(sk, ct)
    assert np.array_equal(ctx.decode(plain), vals)
    assert budget > 60  # 3 x 40-bit data primes, 20-bit t, fresh noise


def test_multiply_plain_and_add_are_slotwise(oracle, toy_ctx):
    ctx = toy_ctx
    rng = np.random.default_rng(2)
    sk = ctx.keygen(3)
    a = rng.integers(0, ctx.t, size=ctx.n, dtype=np.uint64)
    b = rng.integers(0, ctx.t, size=ctx.n, dtype=np.uint64)
    c = rng.integers(0, ctx.t, size=ctx.n, dtype=np.uint64)
    ct = ctx.ct_to_ntt(ctx.encrypt(sk, ctx.encode(a), 5))
    pb = ctx.plain_to_ntt(ctx.encode(b))
    pc = ctx.plain_to_ntt(ctx.encode(c))
    prod = ctx.add(ctx.multiply_plain_ntt(ct, pb), ctx.multiply_plain_ntt(ct, pc))
    mac = ctx.mac_plain_ntt(np.stack([ct, ct]), np.stack([pb, pc]))
    assert np.array_equal(prod, mac)  # lazy MAC == multiply_plain + add chain, bit for bit
    plain, budget = ctx.decrypt(sk, ctx.ct_from_ntt(prod))
    want = (a.astype(object) * (b.astype(object) + c.astype(object))) % ctx.t
    assert ctx.decode(plain).tolist() == [int(x) for x in want]
    assert budget > 0


@pytest.mark.parametrize("step", [1, 2, 5, -1, -3])
def test_rotate_rows_semantics(oracle, toy_ctx, step):
    ctx = toy_ctx
    rng = np.random.default_rng(3)
    sk = ctx.keygen(9)
    vals = rng.integers(0, ctx.t, size=ctx.n, dtype=np.uint64)
    ct = ctx.encrypt(sk, ctx.encode(vals), 13)
    key = ctx.galois_keygen(sk, ctx.galois_elt(step), 17)
    rot = ctx.rotate_rows(ct, step, key)
    plain, budget = ctx.decrypt(sk, rot)
    got = ctx.decode(plain)
    half = ctx.n // 2
    want = np.concatenate([np.roll(vals[:half], -step), np.roll(vals[half:], -step)])  # positive = left
    assert np.array_equal(got, want)
    assert budget > 40


def test_rotation_chain_equals_direct_in_plaintext(oracle, toy_ctx):
    ctx = toy_ctx
    rng = np.random.default_rng(4)
    sk = ctx.keygen(21)
    vals = rng.integers(0, ctx.t, size=ctx.n, dtype=np.uint64)
    ct = ctx.encrypt(sk, ctx.encode(vals), 1)
    k1 = ctx.galois_keygen(sk, ctx.galois_elt(1), 2)
    cur = ct
    for _ in range(7):
        cur = ctx.rotate_rows(cur, 1, k1)
    k7 = ctx.galois_keygen(sk, ctx.galois_elt(7), 3)
    direct = ctx.rotate_rows(ct, 7, k7)
    p1, b1 = ctx.decrypt(sk, cur)
    p2, b2 = ctx.decrypt(sk, direct)
    assert np.array_equal(p1, p2) and b1 > 30 and b2 > 30
    assert not np.array_equal(cur, direct)  # different ciphertexts, same plaintext


def test_mod_switch_preserves_plaintext(oracle, toy_ctx):
    ctx = toy_ctx
    sk = ctx.keygen(5)
    vals = np.arange(ctx.n, dtype=np.uint64) % ctx.t
    ct = ctx.encrypt(sk, ctx.encode(vals), 8)
    low = ctx.mod_switch_next(ct)
    ctx2 = oracle.Context(ctx.n, ctx.primes[: ctx.L - 1] + [ctx.primes[-1]], ctx.t)
    plain, budget = ctx2.decrypt(sk[[*range(ctx.L - 1), ctx.k - 1]], low)
    assert np.array_equal(ctx2.decode(plain), vals) and budget > 20


@pytest.mark.parametrize("d,m,g", [(128, 1, 8), (128, 1, 1), (128, 1, 128), (100, 1, 16), (256, 4, 8), (960, 8, 8)])
def test_layout_slot_algebra(oracle, d, m, g):
    """Plain integer check of the generalised-diagonal identity (no encryption)."""
    n, t = 8192 if d > 256 else 2048, 133857281
    lay = oracle.LayoutPlan(n, d, m, g)
    rng = np.random.default_rng(d + g)
    nvec = min(lay.C, 37)
    xs = rng.integers(0, 256, size=(nvec, d), dtype=np.int32)
    q = rng.integers(0, 256, size=d).astype(np.int64)
    half = n // 2
    acc = np.zeros(n, dtype=object)
    for a in range(lay.m):
        qs = lay.query_slots(t, q, a).astype(object)
        for r in range(lay.R):
            rot = np.concatenate([np.roll(qs[:half], -r), np.roll(qs[half:], -r)])
            acc = (acc + rot * lay.diag_slots(t, xs, a, r).astype(object)) % t
    acc = (acc + lay.norm_slots(t, xs).astype(object)) % t
    qq = int((q * q).sum())
    for u in range(lay.C):
        s = sum(int(acc[lay.slot(u, j)]) for j in range(lay.g)) % t
        got = (s + qq) % t
        want = int(((xs[u].astype(np.int64) - q) ** 2).sum()) if u < nvec else qq
        assert got == want % t, (u, got, want)
    # every slot is owned exactly once
    tab = lay.slot_table().reshape(-1)
    assert sorted(tab.tolist()) == list(range(n))


def _client_distances(ctx, lay, sk, result_ct, qq, nvec):
    plain, budget = ctx.decrypt(sk, result_ct)
    slots = ctx.decode(plain).astype(np.int64)
    tab = lay.slot_table()
    d = (slots[tab].sum(axis=1) + qq) % ctx.t
    return d[:nvec], budget


@pytest.mark.parametrize("chain", [False, True])
def test_encrypted_block_distance_exact(oracle, chain):
    """encrypt -> rotate set -> MAC over a block -> decrypt == exact integer squared L2."""
    n = 2048
    from tests.util import ntt_primes
    primes = ntt_primes(n, 43, 3) + ntt_primes(n, 44, 1)
    t = ntt_primes(n, 24, 1)[0]
    ctx = oracle.Context(n, primes, t)
    d, g = 128, 16
    lay = oracle.LayoutPlan(n, d, 1, g)
    rng = np.random.default_rng(5)
    xs = rng.integers(0, 256, size=(lay.C - 3, d), dtype=np.int32)
    q = rng.integers(0, 256, size=d).astype(np.int64)
    sk = ctx.keygen(1)
    keys = [ctx.galois_keygen(sk, ctx.galois_elt(1 if chain else r), 100 + r) for r in range(1, lay.R)]
    ct = ctx.encrypt(sk, ctx.encode(lay.query_slots(t, q, 0)), 2)
    rot = oracle.rotate_query_set(ctx, lay, ct[None], keys[:1] if chain else keys, chain)
    diag, norm = oracle.encode_block(ctx, lay, xs)
    res = oracle.block_distance(ctx, lay, rot, diag, norm)
    got, budget = _client_distances(ctx, lay, sk, res, int((q * q).sum()), len(xs))
    want = ((xs.astype(np.int64) - q) ** 2).sum(axis=1)
    assert np.array_equal(got, want)
    assert budget > 10


def test_result_mod_switched_to_one_limb_keeps_exact_distances(oracle):
    """The bench / server default ships results at 1 data limb (SEAL mod_switch_to_inplace before save).
    At the production parameters (N=8192, BFVDefault, t=Batching(N,24)) the switch adds at most
    t*(1+N)/2 / q0 = 2^-7 of invariant noise in the WORST case, so decryption stays exact; the test pins
    the budget that is actually left with extreme inputs (all-255 data against an all-0 query)."""
    n, d, g = 8192, 128, 8
    primes, t = oracle.BFV_DEFAULT_PRIMES[n], oracle.BATCHING_T[(n, 24)]
    ctx = oracle.Context(n, primes, t)
    lay = oracle.LayoutPlan(n, d, 1, g)
    rng = np.random.default_rng(9)
    xs = rng.integers(0, 256, size=(lay.C, d), dtype=np.int32)
    xs[:8] = 255
    xs[8:16] = 0
    sk = ctx.keygen(1)
    keys = [ctx.galois_keygen(sk, ctx.galois_elt(r), 100 + r) for r in range(1, lay.R)]
    diag, norm = oracle.encode_block(ctx, lay, xs)
    low = oracle.Context(n, [primes[0], primes[-1]], t)
    sk_low = np.ascontiguousarray(sk[[0, ctx.k - 1]])
    for q in (np.zeros(d, np.int64), np.full(d, 255, np.int64), rng.integers(0, 256, size=d).astype(np.int64)):
        ct = ctx.encrypt(sk, ctx.encode(lay.query_slots(t, q, 0)), 2)
        rot = oracle.rotate_query_set(ctx, lay, ct[None], keys, False)
        res = oracle.block_distance(ctx, lay, rot, diag, norm)
        _, full_budget = _client_distances(ctx, lay, sk, res, int((q * q).sum()), len(xs))
        while res.shape[1] > 1:
            res = ctx.mod_switch_next(res) if res.shape[1] == ctx.L else \
                oracle.Context(n, primes[:res.shape[1]] + [primes[-1]], t).mod_switch_next(res)
        got, budget = _client_distances(low, lay, sk_low, res, int((q * q).sum()), len(xs))
        want = ((xs.astype(np.int64) - q) ** 2).sum(axis=1)
        assert np.array_equal(got, want)
        assert full_budget > 60 and budget >= 6, (full_budget, budget)


@pytest.mark.parametrize("n,d,m,g,tbits,floor", [(8192, 128, 1, 8, 24, 90), (16384, 128, 1, 16, 24, 300),
                                                  (8192, 960, 8, 8, 27, 80)])
def test_noise_budget_at_baseline_parameter_sets(oracle, n, d, m, g, tbits, floor):
    """exact distances and a comfortable noise budget at every BASELINE parameter set
    (SIFT N=8192 / N=16384 with Batching(N,24), GIST-shaped 960-d with the 27-bit plain modulus)."""
    from tests.util import OracleClient
    primes, t = oracle.BFV_DEFAULT_PRIMES[n], oracle.BATCHING_T[(n, tbits)]
    cl = OracleClient(oracle, n, primes, t, d, m, g)
    rng = np.random.default_rng(n + d)
    xs = rng.integers(0, 256, size=(min(cl.lay.C, 96), d), dtype=np.int32)
    xs[0], xs[1] = 255, 0
    q = rng.integers(0, 256, size=d).astype(np.int64)
    keys = cl.step_keys()
    cts = cl.encrypt_query(q.astype(np.float32), 77)
    rot = oracle.rotate_query_set(cl.ctx, cl.lay, cts, keys, False)
    diag, norm = oracle.encode_block(cl.ctx, cl.lay, xs)
    res = oracle.block_distance(cl.ctx, cl.lay, rot, diag, norm)
    dist, budget = cl.distances(res, q, len(xs))
    assert np.array_equal(dist, ((xs.astype(np.int64) - q) ** 2).sum(axis=1))
    assert budget >= floor, budget


def test_plain_path_matches_numpy(oracle):
    rng = np.random.default_rng(6)
    base, query, cent = sift_like(rng, 3000, 128, 32, 9)
    cent = cent + rng.normal(0, 0.37, size=cent.shape).astype(np.float32)  # fractional centroids
    offsets, ids, vecs = build_ivf(base, cent)
    idx, dist = oracle.coarse_quantize(query, cent, 5)
    # independent emulation of the reference arithmetic in numpy scalars
    for i in range(len(query)):
        ds = []
        for j in range(len(cent)):
            acc = np.float32(0)
            for k in range(128):
                diff = np.float32(query[i, k] - cent[j, k])
                acc = np.float32(np.float64(acc) + np.float64(diff) * np.float64(diff))
            ds.append((float(acc), j))
        ds.sort()
        assert [j for _, j in ds[:5]] == idx[i].tolist()
        assert np.allclose([x for x, _ in ds[:5]], dist[i], rtol=0, atol=0)
    dd, labels, sizes = oracle.search_lists_plain(query, idx, offsets, ids, vecs)
    assert sizes.sum() == len(dd)
    off = 0
    for i in range(len(query)):
        want_ids = np.concatenate([ids[offsets[l]:offsets[l + 1]] for l in idx[i]])
        assert np.array_equal(labels[off:off + sizes[i]], want_ids)
        want_d = ((base[want_ids].astype(np.int64) - query[i].astype(np.int64)) ** 2).sum(1)
        assert np.array_equal(dd[off:off + sizes[i]].astype(np.int64), want_d)  # exact for uint8-valued data
        off += sizes[i]


def test_recall_definitions(oracle):
    gt = np.tile(np.arange(100, dtype=np.int32), (2, 1))
    ret = np.tile(np.arange(100, dtype=np.int64), (2, 1))
    r = oracle.recall(ret, gt)
    assert r["ref_recall_1"] == 1.0 and r["ref_recall_10"] == 1.0 and r["std_recall_10"] == 1.0 and r["mrr_10"] == 1.0
    ret2 = ret.copy()
    ret2[:, :10] = np.arange(10, 20)  # top-10 returned are GT ranks 10..19
    ret2[:, 10:20] = np.arange(0, 10)
    r2 = oracle.recall(ret2, gt)
    assert r2["ref_recall_10"] == 1.0   # reference counts any of GT top-100 inside returned top-10
    assert r2["std_recall_10"] == 0.0   # standard recall@10 does not
    assert r2["ref_recall_1"] == 1.0 and r2["mrr_10"] == 0.0


def test_vecs_read(oracle, tmp_path):
    rng = np.random.default_rng(7)
    a = rng.random((5, 7), dtype=np.float32)
    raw = np.zeros((5, 8), dtype=np.float32)
    raw[:, 0] = np.array([7], dtype=np.int32).view(np.float32)[0]
    raw[:, 1:] = a
    p = tmp_path / "x.fvecs"
    raw.tofile(p)
    assert np.array_equal(oracle.vecs_read(str(p)), a)
    (tmp_path / "bad.fvecs").write_bytes(raw.tobytes()[:-4])
    with pytest.raises(IOError):
        oracle.vecs_read(str(tmp_path / "bad.fvecs"))
    with pytest.raises(IOError):
        oracle.vecs_read(str(tmp_path / "missing.fvecs"))
