// pf_seal_prng.h — host-side expansion of SEAL "seeded" ciphertexts (Serializable<Ciphertext> saved by a
// symmetric-key Encryptor): the stream carries c0 and, instead of the uniformly random c1, the 64-byte seed of
// the PRNG that generated it; seal::Ciphertext::load re-creates c1 with Ciphertext::expand_seed
// [EXT: SEAL 4.1 ciphertext.cpp, util/rlwe.cpp sample_poly_uniform, randomgen.cpp Blake2xbPRNG,
// util/blake2xb.c — restated from the published sources; SEAL is not in /root/reference].
// Request-side wire compatibility, on the same slow path as zlib streams: the expansion
// runs on the host before the upload.  Cross-checked on the CPU against an independent restatement in the
// oracle and a pure-Python BLAKE2b (tests/test_seal_seeded.py).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace pfh {

// BLAKE2b with an explicit 64-byte parameter block and an optional key (RFC 7693 / BLAKE2 reference blake2b.c)
struct Blake2bState {
    uint64_t h[8];
    uint64_t t = 0;
    uint8_t buf[128];
    size_t buflen = 0;
    size_t outlen = 0;

    static uint64_t rotr(uint64_t x, int r) { return (x >> r) | (x << (64 - r)); }

    void compress(const uint8_t *block, bool last) {
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
        uint64_t m[16], v[16];
        memcpy(m, block, 128); // little-endian host
        for (int i = 0; i < 8; i++) {
            v[i] = h[i];
            v[i + 8] = IV[i];
        }
        v[12] ^= t;
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
    }

    // param: the 64-byte BLAKE2b parameter block (digest_length at byte 0)
    void init(const uint8_t param[64]) {
        static const uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                                       0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                                       0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        for (int i = 0; i < 8; i++) {
            uint64_t w;
            memcpy(&w, param + 8 * i, 8);
            h[i] = IV[i] ^ w;
        }
        t = 0;
        buflen = 0;
        outlen = param[0];
    }
    void update(const void *in, size_t inlen) {
        const uint8_t *p = static_cast<const uint8_t *>(in);
        while (inlen) {
            if (buflen == 128) { // the buffer is only compressed once more input is known to follow
                t += 128;
                compress(buf, false);
                buflen = 0;
            }
            const size_t take = inlen < 128 - buflen ? inlen : 128 - buflen;
            memcpy(buf + buflen, p, take);
            buflen += take;
            p += take;
            inlen -= take;
        }
    }
    void final(void *out) {
        t += buflen;
        memset(buf + buflen, 0, 128 - buflen);
        compress(buf, true);
        memcpy(out, h, outlen);
    }
};

// BLAKE2Xb (BLAKE2 reference blake2xb.c: blake2xb_init_key / update / final), output length < 2^32 - 1
inline void blake2xb(void *out, size_t outlen, const void *in, size_t inlen, const void *key, size_t keylen) {
    uint8_t P[64] = {0};
    P[0] = 64;                 // digest_length
    P[1] = (uint8_t)keylen;    // key_length
    P[2] = 1;                  // fanout
    P[3] = 1;                  // depth
    const uint32_t xof = (uint32_t)outlen;
    memcpy(P + 12, &xof, 4);   // xof_length (bytes 8..11 = node_offset = 0)
    Blake2bState S;
    S.init(P);
    if (keylen) {
        uint8_t block[128] = {0};
        memcpy(block, key, keylen);
        S.update(block, 128);
    }
    S.update(in, inlen);
    uint8_t root[64];
    S.final(root);
    // expansion nodes: key_length 0, fanout 0, depth 0, leaf_length 64, node_offset i, inner_length 64
    P[1] = 0;
    P[2] = 0;
    P[3] = 0;
    const uint32_t leaf = 64;
    memcpy(P + 4, &leaf, 4);
    P[16] = 0;  // node_depth
    P[17] = 64; // inner_length
    uint8_t *o = static_cast<uint8_t *>(out);
    for (uint32_t i = 0; outlen > 0; i++) {
        const size_t block = outlen < 64 ? outlen : 64;
        P[0] = (uint8_t)block;
        memcpy(P + 8, &i, 4);
        Blake2bState C;
        C.init(P);
        C.update(root, 64);
        C.final(o + (size_t)i * 64);
        outlen -= block;
    }
}

// seal::Blake2xbPRNG: a 4096-byte buffer refilled with blake2xb(buffer, 4096, &counter, 8, seed, 64), counter++
struct SealBlake2xbPrng {
    uint8_t seed[64];
    uint64_t counter = 0;
    uint8_t buffer[4096];
    size_t head = 4096; // empty at construction: the first generate() refills
    explicit SealBlake2xbPrng(const uint8_t s[64]) { memcpy(seed, s, 64); }
    void refill() {
        blake2xb(buffer, sizeof(buffer), &counter, sizeof(counter), seed, 64);
        counter++;
        head = 0;
    }
    void generate(size_t n, uint8_t *dst) {
        while (n) {
            if (head == sizeof(buffer)) refill();
            const size_t take = n < sizeof(buffer) - head ? n : sizeof(buffer) - head;
            memcpy(dst, buffer + head, take);
            head += take;
            dst += take;
            n -= take;
        }
    }
};

// seal::util::sample_poly_uniform (SEAL 4.x rlwe.cpp): fill [L][N] words with PRNG output, then per limb map every
// word into [0, q) — words at or above the largest multiple of q below 2^64 are re-drawn one at a time
inline void seal_sample_poly_uniform(SealBlake2xbPrng &prng, const uint64_t *primes, uint32_t L, uint64_t N, uint64_t *dst) {
    prng.generate((size_t)L * N * 8, reinterpret_cast<uint8_t *>(dst));
    for (uint32_t j = 0; j < L; j++) {
        const uint64_t q = primes[j];
        const uint64_t max_multiple = 0xFFFFFFFFFFFFFFFFULL - (0xFFFFFFFFFFFFFFFFULL % q) - 1;
        uint64_t *p = dst + (size_t)j * N;
        for (uint64_t i = 0; i < N; i++) {
            uint64_t r = p[i];
            while (r >= max_multiple) prng.generate(8, reinterpret_cast<uint8_t *>(&r));
            p[i] = r % q;
        }
    }
}

} // namespace pfh
