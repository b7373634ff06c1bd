"""CPU checks of the seeded-ciphertext path: SEAL's Serializable<Ciphertext> of a symmetric
encryption carries c0 and the 64-byte seed of the Blake2xb PRNG that drew c1.  Three implementations are compared:
the product's host code (prefhetch_b200/csrc/pf_seal_prng.h through pf_seal_ct_expand — no GPU needed), the oracle's
independent C restatement (oracle/pf_oracle_seeded.c) and a pure-Python BLAKE2b / BLAKE2Xb written here from
RFC 7693 and the BLAKE2X paper (checked against hashlib where hashlib can express the parameters)."""
import hashlib
import struct

import numpy as np
import pytest

from tests.util import is_prime, ntt_primes, zlib_stream

IV = [0x6a09e667f3bcc908, 0xbb67ae8584caa73b, 0x3c6ef372fe94f82b, 0xa54ff53a5f1d36f1,
      0x510e527fade682d1, 0x9b05688c2b3e6c1f, 0x1f83d9abfb41bd6b, 0x5be0cd19137e2179]
SIGMA = [[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15], [14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3],
         [11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4], [7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8],
         [9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13], [2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9],
         [12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11], [13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10],
         [6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5], [10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0]]
M64 = (1 << 64) - 1


def _py_blake2b(param: bytes, key: bytes, msg: bytes, outlen: int) -> bytes:
    """RFC 7693 BLAKE2b with an explicit 64-byte parameter block (pure Python, big integers)"""
    h = [IV[i] ^ struct.unpack_from("<Q", param, 8 * i)[0] for i in range(8)]
    data = (key.ljust(128, b"\0") if key else b"") + msg
    total = len(data)
    blocks = [data[i:i + 128].ljust(128, b"\0") for i in range(0, max(total, 1), 128)]
    rotr = lambda x, n: ((x >> n) | (x << (64 - n))) & M64
    for bi, blk in enumerate(blocks):
        last = bi == len(blocks) - 1
        t = total if last else (bi + 1) * 128
        m = struct.unpack("<16Q", blk)
        v = h + IV[:]
        v[12] ^= t
        if last:
            v[14] ^= M64

        def G(a, b, c, d, x, y):
            v[a] = (v[a] + v[b] + x) & M64
            v[d] = rotr(v[d] ^ v[a], 32)
            v[c] = (v[c] + v[d]) & M64
            v[b] = rotr(v[b] ^ v[c], 24)
            v[a] = (v[a] + v[b] + y) & M64
            v[d] = rotr(v[d] ^ v[a], 16)
            v[c] = (v[c] + v[d]) & M64
            v[b] = rotr(v[b] ^ v[c], 63)
        for r in range(12):
            s = SIGMA[r % 10]
            G(0, 4, 8, 12, m[s[0]], m[s[1]])
            G(1, 5, 9, 13, m[s[2]], m[s[3]])
            G(2, 6, 10, 14, m[s[4]], m[s[5]])
            G(3, 7, 11, 15, m[s[6]], m[s[7]])
            G(0, 5, 10, 15, m[s[8]], m[s[9]])
            G(1, 6, 11, 12, m[s[10]], m[s[11]])
            G(2, 7, 8, 13, m[s[12]], m[s[13]])
            G(3, 4, 9, 14, m[s[14]], m[s[15]])
        h = [h[i] ^ v[i] ^ v[i + 8] for i in range(8)]
    return struct.pack("<8Q", *h)[:outlen]


def _param(digest, keylen=0, fanout=1, depth=1, leaf=0, node_offset=0, xof=0, node_depth=0, inner=0) -> bytes:
    return struct.pack("<BBBBIIIBB", digest, keylen, fanout, depth, leaf, node_offset, xof, node_depth, inner) + bytes(46)


def _py_blake2xb(outlen: int, msg: bytes, key: bytes) -> bytes:
    """BLAKE2X paper section 2 / reference blake2xb.c"""
    h0 = _py_blake2b(_param(64, len(key), 1, 1, 0, 0, outlen), key, msg, 64)
    out = b""
    i = 0
    while len(out) < outlen:
        want = min(64, outlen - len(out))
        out += _py_blake2b(_param(want, 0, 0, 0, 64, i, outlen, 0, 64), b"", h0, want)
        i += 1
    return out


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


def _seal_prng_words(seed: bytes, count: int):
    """seal::Blake2xbPRNG as a word stream (pure Python): 4096-byte blocks blake2xb(., 4096, counter_le64, seed)"""
    out, ctr = [], 0
    while len(out) < count:
        blk = _py_blake2xb(4096, struct.pack("<Q", ctr), seed)
        out += list(struct.unpack("<512Q", blk))
        ctr += 1
    return out


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
