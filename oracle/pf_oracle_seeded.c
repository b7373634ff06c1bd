/*
 * pf_oracle_seeded.c — oracle restatement of SEAL's seeded ciphertexts (TEST INFRASTRUCTURE ONLY, see
 * pf_oracle.h): BLAKE2b / BLAKE2Xb, seal::Blake2xbPRNG, util::sample_poly_uniform, the symmetric-key
 * encryption that stores the seed of c1 (Encryptor::encrypt_symmetric(...).save()) and the wire format of
 * such a ciphertext.  [EXT] SEAL 4.1: util/blake2b.c, util/blake2xb.c (the BLAKE2 reference code), randomgen.cpp,
 * util/rlwe.cpp (encrypt_zero_symmetric with save_seed, sample_poly_uniform), ciphertext.cpp (save_members).
 * Written independently of prefhetch_b200/csrc/pf_seal_prng.h; the two are compared byte for byte, and the
 * BLAKE2b core of both against a pure-Python implementation and hashlib (tests/test_seal_seeded.py).
 * PARITY UNPINNED by the reference (no SEAL build exists here).
 */
#include <stdlib.h>
#include <string.h>

#include "pf_oracle.h"

/* ---- BLAKE2b, RFC 7693 section 3 ------------------------------------------------------------- */
static const uint64_t B2_IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                  0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
static const uint8_t B2_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

#define B2_ROTR(x, n) (((x) >> (n)) | ((x) << (64 - (n))))
#define B2_MIX(a, b, c, d, x, y)     \
    do {                             \
        v[a] += v[b] + (x);          \
        v[d] = B2_ROTR(v[d] ^ v[a], 32); \
        v[c] += v[d];                \
        v[b] = B2_ROTR(v[b] ^ v[c], 24); \
        v[a] += v[b] + (y);          \
        v[d] = B2_ROTR(v[d] ^ v[a], 16); \
        v[c] += v[d];                \
        v[b] = B2_ROTR(v[b] ^ v[c], 63); \
    } while (0)

static void b2_f(uint64_t h[8], const uint8_t block[128], uint64_t t, int final_block) {
    uint64_t v[16], m[16];
    for (int i = 0; i < 16; i++) { /* little-endian words, byte by byte: no alignment or host-order assumption */
        uint64_t w = 0;
        for (int b = 7; b >= 0; b--) w = (w << 8) | block[8 * i + b];
        m[i] = w;
    }
    for (int i = 0; i < 8; i++) {
        v[i] = h[i];
        v[8 + i] = B2_IV[i];
    }
    v[12] ^= t;
    if (final_block) v[14] ^= ~(uint64_t)0;
    for (int r = 0; r < 12; r++) {
        const uint8_t *s = B2_SIGMA[r % 10];
        B2_MIX(0, 4, 8, 12, m[s[0]], m[s[1]]);
        B2_MIX(1, 5, 9, 13, m[s[2]], m[s[3]]);
        B2_MIX(2, 6, 10, 14, m[s[4]], m[s[5]]);
        B2_MIX(3, 7, 11, 15, m[s[6]], m[s[7]]);
        B2_MIX(0, 5, 10, 15, m[s[8]], m[s[9]]);
        B2_MIX(1, 6, 11, 12, m[s[10]], m[s[11]]);
        B2_MIX(2, 7, 8, 13, m[s[12]], m[s[13]]);
        B2_MIX(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[8 + i];
}

/* One-shot BLAKE2b over (optional key block) || msg with an explicit parameter block.  The whole padded input
 * is laid out first, which keeps "the last block gets the finalisation flag" trivially right. */
void pfo_blake2b_param(const uint8_t param[64], const uint8_t *key, size_t keylen, const uint8_t *msg, size_t msglen,
                       uint8_t *out, size_t outlen) {
    uint64_t h[8];
    for (int i = 0; i < 8; i++) {
        uint64_t w = 0;
        for (int b = 7; b >= 0; b--) w = (w << 8) | param[8 * i + b];
        h[i] = B2_IV[i] ^ w;
    }
    const size_t total = (keylen ? 128 : 0) + msglen;
    size_t nblocks = total ? (total + 127) / 128 : 1;
    uint8_t *buf = (uint8_t *)calloc(nblocks, 128);
    if (keylen) memcpy(buf, key, keylen);
    if (msglen) memcpy(buf + (keylen ? 128 : 0), msg, msglen);
    for (size_t b = 0; b < nblocks; b++) {
        const int last = b + 1 == nblocks;
        const uint64_t t = last ? (uint64_t)total : (uint64_t)(b + 1) * 128;
        b2_f(h, buf + b * 128, t, last);
    }
    free(buf);
    for (size_t i = 0; i < outlen; i++) out[i] = (uint8_t)(h[i / 8] >> (8 * (i % 8)));
}

/* BLAKE2Xb (Aumasson, Neves, Wilcox-O'Hearn, Winnerlein: "BLAKE2X", section 2): H0 = BLAKE2b of the keyed
 * message with the XOF length in the parameter block; output block i = BLAKE2b(H0) with node offset i,
 * fanout 0, depth 0, leaf length 64, inner length 64 and digest length = bytes still wanted (at most 64). */
void pfo_blake2xb(uint8_t *out, size_t outlen, const uint8_t *in, size_t inlen, const uint8_t *key, size_t keylen) {
    uint8_t P[64];
    memset(P, 0, sizeof(P));
    P[0] = 64;
    P[1] = (uint8_t)keylen;
    P[2] = 1;
    P[3] = 1;
    for (int b = 0; b < 4; b++) P[12 + b] = (uint8_t)((uint32_t)outlen >> (8 * b));
    uint8_t h0[64];
    pfo_blake2b_param(P, key, keylen, in, inlen, h0, 64);
    P[1] = 0;
    P[2] = 0;
    P[3] = 0;
    P[4] = 64; /* leaf length (32-bit little endian) */
    P[17] = 64; /* inner length */
    size_t done = 0;
    for (uint32_t i = 0; done < outlen; i++) {
        const size_t want = outlen - done < 64 ? outlen - done : 64;
        P[0] = (uint8_t)want;
        for (int b = 0; b < 4; b++) P[8 + b] = (uint8_t)(i >> (8 * b));
        pfo_blake2b_param(P, NULL, 0, h0, 64, out + done, want);
        done += want;
    }
}

/* ---- seal::Blake2xbPRNG (randomgen.cpp): 4096-byte blocks blake2xb(., 4096, &counter, 8, seed, 64) -------- */
typedef struct {
    uint8_t seed[64];
    uint64_t counter;
    uint8_t block[4096];
    size_t used;
} seal_prng;

static void prng_init(seal_prng *g, const uint8_t seed[64]) {
    memcpy(g->seed, seed, 64);
    g->counter = 0;
    g->used = sizeof(g->block);
}
static void prng_bytes(seal_prng *g, uint8_t *dst, size_t n) {
    for (size_t i = 0; i < n; i++) {
        if (g->used == sizeof(g->block)) {
            uint8_t ctr[8];
            for (int b = 0; b < 8; b++) ctr[b] = (uint8_t)(g->counter >> (8 * b));
            pfo_blake2xb(g->block, sizeof(g->block), ctr, 8, g->seed, 64);
            g->counter++;
            g->used = 0;
        }
        dst[i] = g->block[g->used++];
    }
}
static uint64_t prng_u64(seal_prng *g) {
    uint8_t b[8];
    prng_bytes(g, b, 8);
    uint64_t w = 0;
    for (int i = 7; i >= 0; i--) w = (w << 8) | b[i];
    return w;
}

/* SEAL util/rlwe.cpp sample_poly_uniform (4.x): the whole [L][n] buffer is filled first; then, limb by limb, every
 * word at or above the largest multiple of q representable is replaced by fresh single words until it is below,
 * and reduced mod q.  out[L][n] over the first L primes of the context. */
void pfo_seal_sample_poly_uniform(const pfo_context *c, int L, const uint8_t seed[64], uint64_t *out) {
    seal_prng g;
    prng_init(&g, seed);
    const uint64_t n = c->n;
    for (size_t i = 0; i < (size_t)L * n; i++) out[i] = prng_u64(&g);
    for (int j = 0; j < L; j++) {
        const uint64_t q = c->q[j].q;
        const uint64_t max_multiple = UINT64_MAX - (UINT64_MAX % q) - 1;
        for (uint64_t i = 0; i < n; i++) {
            uint64_t r = out[(size_t)j * n + i];
            while (r >= max_multiple) r = prng_u64(&g);
            out[(size_t)j * n + i] = r % q;
        }
    }
}

/* Encryptor::encrypt_symmetric with save_seed (BFV): c1 = sample_poly_uniform(PRNG(seed)) taken as the
 * COEFFICIENT form of a (what Ciphertext::expand_seed re-creates on load), c0 = -(a s + e) + round(Q/t m).
 * e comes from the oracle's own noise RNG (not a parity quantity).  ct[2][L][n] coefficient form. */
void pfo_encrypt_symmetric_seeded(const pfo_context *c, const uint64_t *sk, const uint64_t *plain, uint64_t noise_seed,
                                  const uint8_t seed[64], uint64_t *ct) {
    const uint64_t n = c->n;
    const int L = c->L;
    uint64_t *c0 = ct, *c1 = ct + (size_t)L * n;
    pfo_seal_sample_poly_uniform(c, L, seed, c1);
    uint64_t s = noise_seed * 0x9E3779B97F4A7C15ULL + 0x5EEDEDULL;
    uint64_t *a_ntt = (uint64_t *)malloc(n * sizeof(uint64_t));
    int64_t *e = (int64_t *)malloc(n * sizeof(int64_t));
    for (uint64_t i = 0; i < n; i++) { /* centred binomial noise, 21 + 21 coins (SEAL sample_poly_cbd) */
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        e[i] = (int64_t)__builtin_popcountll(s & 0x1FFFFF) - (int64_t)__builtin_popcountll((s >> 21) & 0x1FFFFF);
    }
    for (int j = 0; j < L; j++) {
        const pfo_modulus *m = &c->q[j];
        memcpy(a_ntt, c1 + (size_t)j * n, n * sizeof(uint64_t));
        pfo_ntt_fwd(a_ntt, &c->ntt[j]);
        uint64_t *o = c0 + (size_t)j * n;
        for (uint64_t i = 0; i < n; i++) o[i] = e[i] >= 0 ? (uint64_t)e[i] : m->q - (uint64_t)(-e[i]);
        pfo_ntt_fwd(o, &c->ntt[j]);
        const uint64_t *sj = sk + (size_t)j * n;
        for (uint64_t i = 0; i < n; i++) {
            uint64_t v = pfo_mulmod(a_ntt[i], sj[i], m) + o[i];
            if (v >= m->q) v -= m->q;
            o[i] = v ? m->q - v : 0;
        }
        pfo_ntt_inv(o, &c->ntt[j]);
    }
    free(a_ntt);
    free(e);
    pfo_add_plain_scaled(c, plain, c0);
}

/* Ciphertext::save_members of a seeded ciphertext (compr_mode none): members, DynArray of c0 only, then the
 * UniformRandomGeneratorInfo stream {SEALHeader(81 bytes), type = 1 (blake2xb), seed}. */
size_t pfo_ct_save_seeded_size(uint64_t n, int L) { return 16 + 32 + 1 + 8 * 5 + 16 + 8 + (size_t)L * n * 8 + 16 + 1 + 64; }

static void seeded_put_header(uint8_t *p, uint64_t total) {
    p[0] = 0x5E;
    p[1] = 0xA1;
    p[2] = 0x10;
    p[3] = 4;
    p[4] = 1;
    p[5] = 0;
    p[6] = p[7] = 0;
    memcpy(p + 8, &total, 8);
}

size_t pfo_ct_save_seeded(const uint64_t *c0, uint64_t n, int L, const uint64_t parms_id[4], const uint8_t seed[64],
                          uint8_t prng_type, uint8_t *out) {
    const size_t total = pfo_ct_save_seeded_size(n, L), words = (size_t)L * n;
    uint8_t *p = out;
    seeded_put_header(p, total);
    p += 16;
    memcpy(p, parms_id, 32);
    p += 32;
    *p++ = 0; /* BFV: coefficient form */
    uint64_t v = 2;
    memcpy(p, &v, 8);
    p += 8;
    v = n;
    memcpy(p, &v, 8);
    p += 8;
    v = (uint64_t)L;
    memcpy(p, &v, 8);
    p += 8;
    const double scale = 1.0;
    memcpy(p, &scale, 8);
    p += 8;
    v = 1;
    memcpy(p, &v, 8);
    p += 8;
    seeded_put_header(p, 16 + 8 + words * 8);
    p += 16;
    v = words;
    memcpy(p, &v, 8);
    p += 8;
    memcpy(p, c0, words * 8);
    p += words * 8;
    seeded_put_header(p, 16 + 1 + 64);
    p += 16;
    *p++ = prng_type;
    memcpy(p, seed, 64);
    p += 64;
    return (size_t)(p - out);
}
