"""CPU checks of the seeded-ciphertext path: SEAL's Serializable<Ciphertext> of a symmetric
encryption carries c0 and the 64-byte seed of the Blake2xb PRNG that drew c1.  Three implementations are compared:
the product's host code (prefhetch_b200/csrc/pf_seal_prng.h through pf_seal_ct_expand — no GPU needed), the oracle's
independent C restatement (oracle/pf_oracle_seeded.c) and a pure-Python BLAKE2b / BLAKE2Xb written here from
RFC 7693 and the BLAKE2X paper (checked against hashlib where hashlib can express the parameters)."""
import hashlib
import struct

import numpy as np
import pytest

from tests.util import have_zstd, is_prime, ntt_primes, zlib_stream, zstd_stream

from tests.golden.make_golden_seeded import M64, param_block as _param, py_blake2b as _py_blake2b, py_blake2xb as _py_blake2xb


def test_blake2b_three_way(oracle):
    rng = np.random.default_rng(5)
    for msglen, keylen, outlen in [(0, 0, 64), (3, 0, 32), (128, 0, 64), (129, 16, 48), (1000, 64, 64), (8, 64, 64)]:
        msg, key = rng.bytes(msglen), rng.bytes(keylen)
        want = hashlib.blake2b(msg, digest_size=outlen, key=key).digest()
        p = _param(outlen, keylen)
        assert _py_blake2b(p, key, msg, outlen) == want
        assert oracle.blake2b_param(p, key, msg, outlen) == want
    # a tree-mode parameter block hashlib can express (node offset, leaf length, inner length, fanout 2, depth 2)
    want = hashlib.blake2b(b"xyz", digest_size=40, fanout=2, depth=2, leaf_size=64, node_offset=7, node_depth=0, inner_size=64).digest()
    p = _param(40, 0, 2, 2, 64, 7, 0, 0, 64)
    assert _py_blake2b(p, b"", b"xyz", 40) == want and oracle.blake2b_param(p, b"", b"xyz", 40) == want


def test_blake2xb_oracle_matches_python(oracle):
    rng = np.random.default_rng(6)
    for outlen, msglen, keylen in [(1, 0, 0), (64, 8, 64), (65, 8, 64), (200, 3, 0), (4096, 8, 64), (1000, 300, 17)]:
        msg, key = rng.bytes(msglen), rng.bytes(keylen)
        assert oracle.blake2xb(outlen, msg, key) == _py_blake2xb(outlen, msg, key), (outlen, msglen, keylen)
    # prefixes of different output lengths differ (the XOF length is part of every parameter block)
    assert oracle.blake2xb(128, b"m", b"")[:64] != oracle.blake2xb(64, b"m", b"")


from tests.golden.make_golden_seeded import seal_prng_words as _seal_prng_words  # noqa: E402


@pytest.mark.parametrize("bits", [40, 60])
def test_sample_poly_uniform_oracle_matches_python(oracle, bits):
    """60-bit primes reject a visible share of the 64-bit draws: the re-draws come from the stream AFTER the
    bulk fill, one word at a time, in limb / coefficient order"""
    n, L = 1024, 3
    if bits == 60:   # primes in the middle of the 60-bit range: 2^64 mod q is a sizeable fraction of q
        primes, cand = [], (int(0.71 * 2 ** 60) // (2 * n)) * (2 * n) + 1
        while len(primes) < L + 1:
            if is_prime(cand):
                primes.append(cand)
            cand -= 2 * n
    else:
        primes = ntt_primes(n, bits, L) + ntt_primes(n, bits + 1, 1)
    ctx = oracle.Context(n, primes, ntt_primes(n, 20, 1)[0])
    seed = bytes(range(64))
    got = ctx.sample_poly_uniform(seed)
    words = _seal_prng_words(seed, L * n + 4096)
    pos = L * n
    want = np.zeros((L, n), dtype=np.uint64)
    redraws = 0
    for j in range(L):
        q = primes[j]
        max_multiple = M64 - (M64 % q) - 1
        for i in range(n):
            r = words[j * n + i]
            while r >= max_multiple:
                r = words[pos]
                pos += 1
                redraws += 1
            want[j, i] = r % q
    assert np.array_equal(got, want)
    if bits == 60:
        assert redraws > 20


def _ctx(oracle, n=2048, bits=40):
    primes = ntt_primes(n, bits, 3) + ntt_primes(n, bits + 1, 1)
    t = ntt_primes(n, 24, 1)[0]
    return oracle.Context(n, primes, t), primes, t


@pytest.mark.parametrize("n,bits", [(2048, 40), (1024, 60), (8192, 0)])
def test_product_expansion_matches_oracle(oracle, n, bits):
    """oracle: encrypt with a seeded c1, save seeded -> product: pf_seal_ct_expand -> the oracle's full save of the
    same ciphertext, byte for byte; also through zlib; the ciphertext decrypts to the message"""
    import prefhetch_b200 as pf
    if bits:
        ctx, primes, t = _ctx(oracle, n, bits) if bits < 60 else (None, None, None)
        if bits == 60:   # mid-range 60-bit primes: the rejection branch of sample_poly_uniform is taken
            primes, cand = [], (int(0.71 * 2 ** 60) // (2 * n)) * (2 * n) + 1
            while len(primes) < 4:
                if is_prime(cand):
                    primes.append(cand)
                cand -= 2 * n
            t = ntt_primes(n, 20, 1)[0]
            ctx = oracle.Context(n, primes, t)
    else:
        primes, t = oracle.BFV_DEFAULT_PRIMES[n], oracle.BATCHING_T[(n, 24)]
        ctx = oracle.Context(n, primes, t)
    rng = np.random.default_rng(n)
    sk = ctx.keygen(11)
    vals = rng.integers(0, t, size=n, dtype=np.uint64)
    seed = rng.bytes(64)
    ct = ctx.encrypt_seeded(sk, ctx.encode(vals), 77, seed)
    assert np.array_equal(ct[1], ctx.sample_poly_uniform(seed))
    plain, budget = ctx.decrypt(sk, ct)
    assert np.array_equal(ctx.decode(plain), vals) and budget > 0
    pid = (1, 2, 3, 4)
    seeded = ctx.ct_save_seeded(ct, seed, parms_id=pid)
    full = ctx.ct_save(ct, parms_id=pid)
    assert len(seeded) == 113 + ctx.L * n * 8 + 81 and len(full) == 113 + 2 * ctx.L * n * 8
    data_primes = primes[:-1]
    assert pf.seal_ct_expand(seeded, n, data_primes) == full
    assert pf.seal_ct_expand(zlib_stream(seeded), n, data_primes) == full
    assert pf.seal_ct_expand(full, n, data_primes) == full                       # not seeded: passes through
    assert pf.seal_ct_expand(zlib_stream(full), n, data_primes) == full
    if have_zstd():     # SEAL's default compr_mode when built with zstd
        assert pf.seal_ct_expand(zstd_stream(seeded, streaming=True), n, data_primes) == full
        assert pf.seal_ct_expand(zstd_stream(full), n, data_primes) == full
    assert pf.seal_ct_expand(seeded + b"next stream", n, data_primes) == full    # only its own bytes are consumed


def test_product_expansion_rejects_malformed(oracle):
    import prefhetch_b200 as pf
    ctx, primes, t = _ctx(oracle)
    n = ctx.n
    sk = ctx.keygen(3)
    ct = ctx.encrypt_seeded(sk, ctx.encode(np.arange(n, dtype=np.uint64) % t), 1, bytes(64))
    good = ctx.ct_save_seeded(ct, bytes(64))
    data_primes = primes[:-1]
    assert len(pf.seal_ct_expand(good, n, data_primes)) == 113 + 2 * ctx.L * n * 8
    shake = ctx.ct_save_seeded(ct, bytes(64), prng_type=2)                       # shake256 PRNG: not supported
    bad_info = bytearray(good)
    bad_info[-81] ^= 0xFF                                                        # magic of the PRNG info stream
    for bad in (shake, good[:-1], good[:200], bytes(bad_info), good[:5] + b"\x02" + good[6:]):
        with pytest.raises(pf.PfError) as ei:
            pf.seal_ct_expand(bad, n, data_primes)
        assert ei.value.code == 5
    with pytest.raises(pf.PfError):
        pf.seal_ct_expand(good, 2 * n, data_primes)                              # another poly degree
    with pytest.raises(pf.PfError):
        pf.seal_ct_expand(good, n, data_primes[:-1])                             # another limb count


def test_committed_golden_vectors(oracle):
    """tests/golden/kat_seeded_v1.json (made by tests/golden/make_golden_seeded.py, pure Python): the oracle's BLAKE2Xb,
    PRNG stream and sampler reproduce the committed vectors; so does the product's expansion (through a stream)"""
    import json
    from pathlib import Path
    import prefhetch_b200 as pf
    kat = json.loads((Path(__file__).resolve().parent / "golden" / "kat_seeded_v1.json").read_text())
    for v in kat["blake2xb"]:
        assert oracle.blake2xb(v["outlen"], bytes.fromhex(v["msg"]), bytes.fromhex(v["key"])).hex() == v["out"]
    for v in kat["prng_words"]:
        seed = bytes.fromhex(v["seed"])
        blk0 = oracle.blake2xb(4096, struct.pack("<Q", 0), seed)
        blk1 = oracle.blake2xb(4096, struct.pack("<Q", 1), seed)
        w = list(struct.unpack("<1024Q", blk0 + blk1))
        assert w[:4] == v["first4"] and w[511:514] == v["word511_512_513"]
        assert hashlib.sha256(struct.pack("<1024Q", *w)).hexdigest() == v["sha256_first_1024"]
    for v in kat["sample_poly_uniform"]:
        n, primes, seed = v["n"], v["primes"], bytes.fromhex(v["seed"])
        # the product: a seeded stream with c0 = 0 over these primes -> c1 of the expanded stream
        L = len(primes)
        hdr = bytearray(113)
        hdr[0:8] = bytes([0x5E, 0xA1, 0x10, 4, 1, 0, 0, 0])
        struct.pack_into("<Q", hdr, 8, 113 + L * n * 8 + 81)
        struct.pack_into("<QQQ", hdr, 49, 2, n, L)
        struct.pack_into("<d", hdr, 73, 1.0)
        struct.pack_into("<Q", hdr, 81, 1)
        hdr[89:97] = bytes([0x5E, 0xA1, 0x10, 4, 1, 0, 0, 0])
        struct.pack_into("<QQ", hdr, 97, 16 + 8 + L * n * 8, L * n)
        info = bytes([0x5E, 0xA1, 0x10, 4, 1, 0, 0, 0]) + struct.pack("<Q", 81) + b"\x01" + seed
        full = pf.seal_ct_expand(bytes(hdr) + bytes(L * n * 8) + info, n, primes)
        c1 = np.frombuffer(full[113 + L * n * 8:], dtype=np.uint64)
        assert c1[:4].tolist() == v["first4"]
        assert hashlib.sha256(c1.tobytes()).hexdigest() == v["sha256"]
    assert kat["sample_poly_uniform"][1]["redraws"] > 10


def test_batch_expansion_is_the_per_stream_expansion(oracle):
    """pf_seal_ct_expand_batch — the slow path of the search calls for a whole request, spread over host threads —
    returns exactly what pf_seal_ct_expand returns stream by stream, for any mix of full / seeded / zlib / zstd streams
    and any thread count; untouched streams are copied; one bad stream fails the batch"""
    import time
    import prefhetch_b200 as pf
    n = 4096
    primes, t = ntt_primes(n, 40, 3) + ntt_primes(n, 41, 1), ntt_primes(n, 24, 1)[0]
    ctx = oracle.Context(n, primes, t)
    rng = np.random.default_rng(77)
    sk = ctx.keygen(1)
    data_primes = primes[:-1]
    parts, want = [], []
    for i in range(24):
        seed = rng.bytes(64)
        ct = ctx.encrypt_seeded(sk, ctx.encode(rng.integers(0, t, size=n, dtype=np.uint64)), 100 + i, seed)
        full, seeded = ctx.ct_save(ct), ctx.ct_save_seeded(ct, seed)
        kind = i % 6
        s = [full, seeded, zlib_stream(seeded), zlib_stream(full), seeded, full][kind]
        if have_zstd() and kind >= 4:
            s = zstd_stream(seeded if kind == 4 else full, streaming=bool(i & 8))
        parts.append(s)
        want.append(full)
    blob = b"".join(parts)
    offs = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.uint64)
    woffs = np.concatenate([[0], np.cumsum([len(p) for p in want])]).astype(np.uint64)
    for threads in (1, 3, 8, 0):
        t0 = time.perf_counter()
        out, ooffs = pf.seal_ct_expand_batch(blob, offs, n, data_primes, threads)
        dt = time.perf_counter() - t0
        assert out == b"".join(want) and np.array_equal(ooffs, woffs), threads
        print(f"batch of {len(parts)} streams, threads={threads}: {dt * 1e3:.1f} ms")
    # all-full batch: nothing to do, streams copied; empty batch
    fblob = b"".join(want)
    out, ooffs = pf.seal_ct_expand_batch(fblob, woffs, n, data_primes, 4)
    assert out == fblob and np.array_equal(ooffs, woffs)
    out, ooffs = pf.seal_ct_expand_batch(b"\0", np.zeros(1, dtype=np.uint64), n, data_primes, 4)
    assert out == b"" and list(ooffs) == [0]
    # one stream seeded with shake256, one truncated, offsets out of range
    bad = list(parts)
    bad[7] = parts[7][:-65] + b"\x02" + parts[7][-64:]
    assert len(parts[7]) == 113 + len(data_primes) * n * 8 + 81        # index 7 is an uncompressed seeded stream
    boffs = np.concatenate([[0], np.cumsum([len(p) for p in bad])]).astype(np.uint64)
    with pytest.raises(pf.PfError):
        pf.seal_ct_expand_batch(b"".join(bad), boffs, n, data_primes, 4)
    bad = list(parts)
    bad[2] = parts[2][:-30]
    boffs = np.concatenate([[0], np.cumsum([len(p) for p in bad])]).astype(np.uint64)
    with pytest.raises(pf.PfError):
        pf.seal_ct_expand_batch(b"".join(bad), boffs, n, data_primes, 4)
    with pytest.raises(pf.PfError):
        pf.seal_ct_expand_batch(blob[:-5], offs, n, data_primes, 4)
