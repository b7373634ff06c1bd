#!/usr/bin/env python3
"""Generate tests/golden/kat_seeded_v1.json — known-answer vectors for the seeded-ciphertext path:
BLAKE2b with an explicit parameter block, BLAKE2Xb, the word stream of seal::Blake2xbPRNG and
util::sample_poly_uniform, computed by the pure-Python implementation in this file (RFC 7693; the BLAKE2X paper,
section 2; [EXT] SEAL 4.1 randomgen.cpp / util/rlwe.cpp restated).  It shares no code with oracle/pf_oracle_seeded.c or
prefhetch_b200/csrc/pf_seal_prng.h; its BLAKE2b is checked against hashlib by tests/test_seal_seeded.py.  These are
NOT outputs of SEAL (absent here): parity stays unpinned by the reference.  Run: python tests/golden/make_golden_seeded.py"""
import hashlib
import json
import struct
from pathlib import Path

OUT = Path(__file__).resolve().parent / "kat_seeded_v1.json"

IV = [0x6a09e667f3bcc908, 0xbb67ae8584caa73b, 0x3c6ef372fe94f82b, 0xa54ff53a5f1d36f1,
      0x510e527fade682d1, 0x9b05688c2b3e6c1f, 0x1f83d9abfb41bd6b, 0x5be0cd19137e2179]
SIGMA = [[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15], [14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3],
         [11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4], [7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8],
         [9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13], [2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9],
         [12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11], [13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10],
         [6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5], [10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0]]
M64 = (1 << 64) - 1


def py_blake2b(param: bytes, key: bytes, msg: bytes, outlen: int) -> bytes:
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


def param_block(digest, keylen=0, fanout=1, depth=1, leaf=0, node_offset=0, xof=0, node_depth=0, inner=0) -> bytes:
    return struct.pack("<BBBBIIIBB", digest, keylen, fanout, depth, leaf, node_offset, xof, node_depth, inner) + bytes(46)


def py_blake2xb(outlen: int, msg: bytes, key: bytes) -> bytes:
    """BLAKE2X paper section 2 / reference blake2xb.c"""
    h0 = py_blake2b(param_block(64, len(key), 1, 1, 0, 0, outlen), key, msg, 64)
    out = b""
    i = 0
    while len(out) < outlen:
        want = min(64, outlen - len(out))
        out += py_blake2b(param_block(want, 0, 0, 0, 64, i, outlen, 0, 64), b"", h0, want)
        i += 1
    return out


def seal_prng_words(seed: bytes, count: int):
    """seal::Blake2xbPRNG as a word stream: 4096-byte blocks blake2xb(., 4096, counter_le64, seed), counter = 0, 1, ..."""
    out, ctr = [], 0
    while len(out) < count:
        blk = py_blake2xb(4096, struct.pack("<Q", ctr), seed)
        out += list(struct.unpack("<512Q", blk))
        ctr += 1
    return out


def sample_poly_uniform(seed: bytes, primes, n: int):
    """util::sample_poly_uniform: bulk fill, then per limb re-draw words >= the largest multiple of q and reduce"""
    L = len(primes)
    words = seal_prng_words(seed, L * n + 8192)
    pos, out, redraws = L * n, [], 0
    for j, q in enumerate(primes):
        max_multiple = M64 - (M64 % q) - 1
        row = []
        for i in range(n):
            r = words[j * n + i]
            while r >= max_multiple:
                r = words[pos]
                pos += 1
                redraws += 1
            row.append(r % q)
        out.append(row)
    return out, redraws


def main():
    kat = {"blake2xb": [], "prng_words": [], "sample_poly_uniform": []}
    msgs = [(1, b"", b""), (64, struct.pack("<Q", 0), bytes(range(64))), (65, struct.pack("<Q", 1), bytes(range(64))),
            (200, b"abc", b""), (1000, bytes(range(200)), bytes(range(17)))]
    for outlen, msg, key in msgs:
        kat["blake2xb"].append({"outlen": outlen, "msg": msg.hex(), "key": key.hex(), "out": py_blake2xb(outlen, msg, key).hex()})
    for seed in (bytes(64), bytes(range(64)), bytes((7 * i + 3) & 0xFF for i in range(64))):
        w = seal_prng_words(seed, 1024)
        kat["prng_words"].append({"seed": seed.hex(), "first4": w[:4], "word511_512_513": w[511:514],
                                  "sha256_first_1024": hashlib.sha256(struct.pack("<1024Q", *w[:1024])).hexdigest()})
    # BFVDefault data primes of N = 8192 on a short polynomial (the sampler does not depend on N being the ring degree)
    # and mid-range 60-bit moduli (the sampler needs no primality), where a visible share of the draws is rejected
    cases = [("bfv8192_data_primes", [0x7FFFFFD8001, 0x7FFFFFC8001, 0xFFFFFFFC001, 0xFFFFFF6C001], 256),
             ("midrange_60bit", [818575470775332865 - 2048 * k for k in (0, 1, 2)], 512)]
    for name, primes, n in cases:
        seed = bytes(range(64))
        rows, redraws = sample_poly_uniform(seed, primes, n)
        flat = [x for row in rows for x in row]
        kat["sample_poly_uniform"].append({"name": name, "primes": primes, "n": n, "seed": seed.hex(), "redraws": redraws,
                                           "first4": flat[:4], "sha256": hashlib.sha256(struct.pack(f"<{len(flat)}Q", *flat)).hexdigest()})
    OUT.write_text(json.dumps(kat, indent=1))
    print(OUT, {k: len(v) for k, v in kat.items()}, [c["redraws"] for c in kat["sample_poly_uniform"]])


if __name__ == "__main__":
    main()
