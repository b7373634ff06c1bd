"""CPU oracle vs the independent math-level golden vectors (tests/golden/kat_v1.json)."""
import ctypes as C
import hashlib

import numpy as np
import pytest


def test_default_primes_minimal_roots(oracle, golden):
    for key, psi in golden["minimal_psi"].items():
        n, q = (int(x) for x in key.split(":"))
        m = oracle.modulus(q)
        assert oracle.lib().pfo_minimal_primitive_root(2 * n, C.byref(m)) == psi, key
    # SURVEY.md §8c (ii) probe values
    assert golden["minimal_psi"]["8192:%d" % 0x7FFFFFD8001] == 1734247217
    assert golden["minimal_psi"]["8192:%d" % 0x7FFFFFC8001] == 304486499
    for key, t in golden["batching_t"].items():
        n, bits = (int(x) for x in key.split(":"))
        assert oracle.BATCHING_T[(n, bits)] == t


def test_mulmod_barrett_shoup(oracle, golden):
    l = oracle.lib()
    for q, a, b, r in golden["mulmod"]:
        m = oracle.modulus(q)
        assert l.pfo_mulmod(a, b, C.byref(m)) == r
    for q, lo, hi, r in golden["barrett128"]:
        m = oracle.modulus(q)
        assert l.pfo_barrett128(lo, hi, C.byref(m)) == r
    for q, ratio in golden["ratio"]:
        m = oracle.modulus(q)
        assert m.ratio[0] + (m.ratio[1] << 64) == ratio
    for q, w, s in golden["shoup"]:
        assert l.pfo_shoup(w, q) == s


def _tables_ctx(oracle, n, q):
    # a context whose limb 0 is q (needs >= 2 primes; reuse q-compatible second prime = q itself is
    # not allowed to repeat in SEAL but the oracle's tables are per limb, so pick t = q for NTT mod t)
    from tests.util import ntt_primes
    other = [p for p in ntt_primes(n, 30, 2) if p != q][:1]
    return oracle.Context(n, [q] + other, ntt_primes(n, 20 if n <= 1024 else 24, 1)[0])


def test_ntt_matches_definition(oracle, golden):
    for case in golden["ntt_def"]:
        ctx = _tables_ctx(oracle, case["n"], case["q"])
        a = np.array(case["a"], dtype=np.uint64)
        out = ctx.ntt_fwd(a, 0)
        assert out.tolist() == case["ntt"], f"n={case['n']}"
        assert ctx.ntt_inv(out, 0).tolist() == case["a"]


def test_ntt_full_size_sparse(oracle, golden):
    for case in golden["ntt_sparse"]:
        n, q = case["n"], case["q"]
        ctx = _tables_ctx(oracle, n, q)
        a = np.zeros(n, dtype=np.uint64)
        for p, c in zip(case["pos"], case["coef"]):
            a[p] = c
        out = ctx.ntt_fwd(a, 0)
        assert out[:4].tolist() == case["first4"] and int(out[-1]) == case["last"]
        assert hashlib.sha256(out.tobytes()).hexdigest() == case["sha256"]
        assert np.array_equal(ctx.ntt_inv(out, 0), a)


def test_ntt_product_is_negacyclic(oracle, golden):
    g = golden["negacyclic"]
    ctx = _tables_ctx(oracle, g["n"], g["q"])
    fa = ctx.ntt_fwd(np.array(g["a"], dtype=np.uint64), 0)
    fb = ctx.ntt_fwd(np.array(g["b"], dtype=np.uint64), 0)
    prod = np.array([int(x) * int(y) % g["q"] for x, y in zip(fa, fb)], dtype=np.uint64)
    assert ctx.ntt_inv(prod, 0).tolist() == g["prod"]


def test_batch_encoder_definition(oracle, golden):
    from tests.util import ntt_primes
    for case in golden["batch_encode"]:
        n, t = case["n"], case["t"]
        ctx = oracle.Context(n, ntt_primes(n, 30, 2), t)
        plain = ctx.encode(np.array(case["values"], dtype=np.uint64))
        assert plain.tolist() == case["plain"]
        assert ctx.decode(plain).tolist() == case["values"]


def test_galois_coeff_and_ntt(oracle, golden):
    g = golden["galois"]
    ctx = _tables_ctx(oracle, g["n"], g["q"])
    a = np.array(g["a"], dtype=np.uint64)
    fa = ctx.ntt_fwd(a, 0)
    assert fa.tolist() == g["ntt_a"]
    for case in g["cases"]:
        assert ctx.apply_galois(a, case["elt"], 0).tolist() == case["coeff"]
        assert ctx.apply_galois_ntt(fa, case["elt"]).tolist() == case["ntt_of_coeff"]
    # step -> element (util/galois.cpp): left rotation by s is 3^s, right by s is 3^(n/2-s), 0 is 2n-1
    n = g["n"]
    assert ctx.galois_elt(1) == 3 and ctx.galois_elt(2) == 9
    assert ctx.galois_elt(0) == 2 * n - 1
    assert ctx.galois_elt(-1) == pow(3, n // 2 - 1, 2 * n)


def test_plain_lift_and_scaling_variant(oracle, golden):
    g = golden["plain_ops"]
    ctx = oracle.Context(g["n"], g["primes"], g["t"])
    m = np.array(g["m"], dtype=np.uint64)
    assert ctx.plain_to_ntt(m).tolist() == g["lift_ntt"]
    z = np.zeros((ctx.L, ctx.n), dtype=np.uint64)
    assert ctx.add_plain_scaled(m, z).tolist() == g["scaled"]


def test_switch_key_formula(oracle, golden):
    g = golden["switch_key"]
    ctx = oracle.Context(g["n"], g["primes"], g["t"])
    out = ctx.switch_key(np.array(g["ct"], dtype=np.uint64), np.array(g["target"], dtype=np.uint64),
                         np.array(g["key"], dtype=np.uint64))
    assert out.tolist() == g["out"]


def test_mod_switch_formula(oracle, golden):
    g = golden["mod_switch"]
    ctx = oracle.Context(g["n"], g["primes"] + [  # context needs one more prime than the ct has limbs
        __import__("tests.util", fromlist=["ntt_primes"]).ntt_primes(g["n"], 33, 1)[0]], 97)
    out = ctx.mod_switch_next(np.array(g["ct"], dtype=np.uint64))
    assert out.tolist() == g["out"]


def test_wire_format(oracle, golden):
    g = golden["wire"]
    ct = np.array(g["data"], dtype=np.uint64).reshape(2, g["L"], g["n"])
    ctx_less_save = oracle.Context.ct_save
    blob = ctx_less_save(None, ct, False, tuple(g["parms_id"]))
    assert blob.hex() == g["hex"]
    back, is_ntt, pid, used = oracle.Context.ct_load(bytes.fromhex(g["hex"]))
    assert used == len(blob) and not is_ntt and list(pid) == g["parms_id"]
    assert np.array_equal(back, ct)
    with pytest.raises(ValueError):
        oracle.Context.ct_load(blob[:-8])
    bad = bytearray(blob)
    bad[0] ^= 1
    with pytest.raises(ValueError):
        oracle.Context.ct_load(bytes(bad))


def test_reference_distance_arithmetic(oracle, golden):
    for case in golden["l2_ref"]:
        q = np.array(case["q"], dtype=np.float32)
        c = np.array(case["c"], dtype=np.float32)
        d = oracle.lib().pfo_l2sqr_ref(q.ctypes.data_as(oracle.f32p), c.ctypes.data_as(oracle.f32p), len(q))
        assert int(np.float32(d).view(np.uint32)) == case["dist_bits"]
