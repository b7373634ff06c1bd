// pf_blake2b.h — BLAKE2b (RFC 7693, unkeyed) on the host, for SEAL's parms_id:
// EncryptionParameters::compute_parms_id hashes the uint64 array {scheme, poly_modulus_degree,
// coeff_modulus values..., plain_modulus} with blake2b to 32 bytes = 4 little-endian uint64
// [EXT: SEAL 4.1 encryptionparams.cpp / util/hash.h, restated from the published source; the hash
// itself is checked against Python's hashlib in tests/test_abi.py].
#pragma once
#include <cstdint>
#include <cstring>

namespace pfh {

inline void blake2b(const void *in, size_t inlen, void *out, size_t outlen) {
    static const uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                                   0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                                   0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    static const uint8_t SIGMA[12][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
    uint64_t h[8];
    for (int i = 0; i < 8; i++) h[i] = IV[i];
    h[0] ^= 0x01010000ULL ^ (uint64_t)outlen;
    auto rotr = [](uint64_t x, int r) { return (x >> r) | (x << (64 - r)); };
    auto compress = [&](const uint8_t *block, uint64_t t, bool last) {
        uint64_t m[16], v[16];
        memcpy(m, block, 128); // little-endian host
        for (int i = 0; i < 8; i++) {
            v[i] = h[i];
            v[i + 8] = IV[i];
        }
        v[12] ^= t; // message lengths here stay far below 2^64: the high counter word is 0
        if (last) v[14] = ~v[14];
        auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
            v[a] = v[a] + v[b] + x;
            v[d] = rotr(v[d] ^ v[a], 32);
            v[c] = v[c] + v[d];
            v[b] = rotr(v[b] ^ v[c], 24);
            v[a] = v[a] + v[b] + y;
            v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];
            v[b] = rotr(v[b] ^ v[c], 63);
        };
        for (int r = 0; r < 12; r++) {
            const uint8_t *s = SIGMA[r];
            G(0, 4, 8, 12, m[s[0]], m[s[1]]);
            G(1, 5, 9, 13, m[s[2]], m[s[3]]);
            G(2, 6, 10, 14, m[s[4]], m[s[5]]);
            G(3, 7, 11, 15, m[s[6]], m[s[7]]);
            G(0, 5, 10, 15, m[s[8]], m[s[9]]);
            G(1, 6, 11, 12, m[s[10]], m[s[11]]);
            G(2, 7, 8, 13, m[s[12]], m[s[13]]);
            G(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
    };
    const uint8_t *p = static_cast<const uint8_t *>(in);
    uint64_t t = 0;
    while (inlen > 128) {
        t += 128;
        compress(p, t, false);
        p += 128;
        inlen -= 128;
    }
    uint8_t last[128] = {0};
    memcpy(last, p, inlen);
    t += inlen;
    compress(last, t, true);
    memcpy(out, h, outlen);
}

// SEAL parms_id of {scheme bfv = 1, N, primes[0..n), t}
inline void seal_parms_id(uint64_t poly_degree, const uint64_t *primes, uint32_t nprimes, uint64_t plain_modulus,
                          uint64_t out[4]) {
    uint64_t buf[3 + 64];
    uint32_t w = 0;
    buf[w++] = 1; // scheme_type::bfv
    buf[w++] = poly_degree;
    for (uint32_t i = 0; i < nprimes && i < 64; i++) buf[w++] = primes[i];
    buf[w++] = plain_modulus;
    blake2b(buf, (size_t)w * 8, out, 32);
}

} // namespace pfh
