/*
 * pf_oracle.c — CPU oracle for the PreFHEtch server-side search hot path.  See pf_oracle.h.
 * TEST INFRASTRUCTURE ONLY.  The plaintext functions (pfo_l2sqr_ref, pfo_coarse_quantize, pfo_search_lists_plain,
 * pfo_recall) are PINNED to outputs of the reference's own code (oracle/_ref, tests/golden/ref_plain_v1.json,
 * tests/test_ref_pin.py); the HE part is PARITY UNPINNED by the reference (no HE code / tests / vectors there).
 *
 * Citation convention: "ref:" paths are relative to /root/reference (in-tree reference code);
 * "SEAL:" paths name files of Microsoft SEAL 4.1 (native/src/seal/...), the third-party
 * dependency the reference pins but does not vendor — those are restatements of the published
 * algorithm, checked by the math-level golden vectors under tests/golden/.
 */
#include "pf_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------------------------
 * Modular arithmetic — SEAL: util/uintarithsmallmod.h, modulus.cpp
 * ------------------------------------------------------------------------------------------ */

int pfo_modulus_init(pfo_modulus *m, uint64_t q) {
    if (q < 2 || (q >> 61)) return -1;
    m->q = q;
    /* floor(2^128 / q) by long division of the 3-word number 2^128 (SEAL Modulus::set_value) */
    u128 num_hi = ((u128)1 << 64); /* words: [0]=0,[1]=0,[2]=1 -> top two words as u128 */
    uint64_t q1 = (uint64_t)(num_hi / q);
    u128 rem = num_hi % q;
    uint64_t q0 = (uint64_t)(((rem << 64)) / q);
    m->ratio[1] = q1;
    m->ratio[0] = q0;
    return 0;
}

/* SEAL: barrett_reduce_64 */
uint64_t pfo_barrett64(uint64_t x, const pfo_modulus *m) {
    uint64_t qh = (uint64_t)(((u128)x * m->ratio[1]) >> 64);
    uint64_t r = x - qh * m->q;
    return r >= m->q ? r - m->q : r;
}

/* SEAL: barrett_reduce_128 (two-word Barrett with const_ratio) */
uint64_t pfo_barrett128(uint64_t lo, uint64_t hi, const pfo_modulus *m) {
    uint64_t carry = (uint64_t)(((u128)lo * m->ratio[0]) >> 64);
    u128 t2 = (u128)lo * m->ratio[1];
    uint64_t tmp1 = (uint64_t)t2 + carry;
    uint64_t tmp3 = (uint64_t)(t2 >> 64) + (tmp1 < carry);
    t2 = (u128)hi * m->ratio[0];
    uint64_t s = tmp1 + (uint64_t)t2;
    carry = (uint64_t)(t2 >> 64) + (s < tmp1);
    tmp1 = hi * m->ratio[1] + tmp3 + carry;
    uint64_t r = lo - tmp1 * m->q;
    return r >= m->q ? r - m->q : r;
}

uint64_t pfo_mulmod(uint64_t a, uint64_t b, const pfo_modulus *m) {
    u128 p = (u128)a * b;
    return pfo_barrett128((uint64_t)p, (uint64_t)(p >> 64), m);
}

uint64_t pfo_powmod(uint64_t a, uint64_t e, const pfo_modulus *m) {
    uint64_t r = 1 % m->q;
    a = pfo_barrett64(a, m);
    while (e) {
        if (e & 1) r = pfo_mulmod(r, a, m);
        a = pfo_mulmod(a, a, m);
        e >>= 1;
    }
    return r;
}

uint64_t pfo_invmod(uint64_t a, const pfo_modulus *m) { return pfo_powmod(a, m->q - 2, m); }

/* SEAL: MultiplyUIntModOperand::set_quotient */
uint64_t pfo_shoup(uint64_t w, uint64_t q) { return (uint64_t)((((u128)w) << 64) / q); }

static inline uint64_t addmod(uint64_t a, uint64_t b, uint64_t q) {
    uint64_t s = a + b;
    return s >= q ? s - q : s;
}
static inline uint64_t submod(uint64_t a, uint64_t b, uint64_t q) { return a >= b ? a - b : a + q - b; }
static inline uint64_t negmod(uint64_t a, uint64_t q) { return a ? q - a : 0; }

/* SEAL: multiply_uint_mod(x, MultiplyUIntModOperand) — Shoup/Harvey, canonical result */
static inline uint64_t mul_shoup(uint64_t x, uint64_t w, uint64_t wsh, uint64_t q) {
    uint64_t qh = (uint64_t)(((u128)x * wsh) >> 64);
    uint64_t r = x * w - qh * q;
    return r >= q ? r - q : r;
}
static inline uint64_t mul_shoup_lazy(uint64_t x, uint64_t w, uint64_t wsh, uint64_t q) {
    uint64_t qh = (uint64_t)(((u128)x * wsh) >> 64);
    return x * w - qh * q; /* [0, 2q) */
}

uint32_t pfo_bitrev(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}

/* splitmix64 — the oracle's own deterministic RNG (client-side randomness is not part of parity) */
static inline uint64_t rng_next(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* SEAL: util/numth.cpp try_primitive_root + try_minimal_primitive_root.  The result (the
 * numerically smallest primitive two_n-th root) does not depend on which root the search starts
 * from, so the starting root is found deterministically instead of with SEAL's random device. */
uint64_t pfo_minimal_primitive_root(uint64_t two_n, const pfo_modulus *m) {
    uint64_t q = m->q;
    if ((q - 1) % two_n) return 0;
    uint64_t quotient = (q - 1) / two_n, root = 0, seed = 0x5EA1;
    for (int tries = 0; tries < 1000 && !root; tries++) {
        uint64_t x = rng_next(&seed) % q;
        uint64_t r = pfo_powmod(x, quotient, m);
        if (r && pfo_powmod(r, two_n >> 1, m) == q - 1) root = r;
    }
    if (!root) return 0;
    uint64_t gen_sq = pfo_mulmod(root, root, m), cur = root, best = root;
    for (uint64_t i = 0; i < (two_n >> 1); i++) {
        if (cur < best) best = cur;
        cur = pfo_mulmod(cur, gen_sq, m);
    }
    return best;
}

/* ------------------------------------------------------------------------------------------
 * NTT — SEAL: util/ntt.cpp (NTTTables::initialize), util/dwthandler.h (transform_to_rev /
 * transform_from_rev).  Forward: Cooley-Tukey, natural in -> bit-reversed out, output index i
 * holds a(psi^(2*bitrev(i)+1)).  Inverse: Gentleman-Sande, bit-reversed in -> natural out with
 * n^{-1} applied.  Harvey lazy butterflies inside, stored results fully reduced.
 * ------------------------------------------------------------------------------------------ */

static int ntt_tables_init(pfo_ntt_tables *T, uint64_t n, int logn, uint64_t q) {
    memset(T, 0, sizeof(*T));
    T->n = n;
    T->logn = logn;
    if (pfo_modulus_init(&T->mod, q)) return -1;
    T->psi = pfo_minimal_primitive_root(2 * n, &T->mod);
    if (!T->psi) return -1;
    uint64_t psi_inv = pfo_invmod(T->psi, &T->mod);
    T->n_inv = pfo_invmod(n % q, &T->mod);
    T->rp = (uint64_t *)malloc(4 * n * sizeof(uint64_t));
    if (!T->rp) return -1;
    T->rp_sh = T->rp + n;
    T->irp = T->rp + 2 * n;
    T->irp_sh = T->rp + 3 * n;
    uint64_t p = 1, ip = 1;
    for (uint64_t i = 0; i < n; i++) {
        uint32_t r = pfo_bitrev((uint32_t)i, logn);
        T->rp[r] = p;
        T->rp_sh[r] = pfo_shoup(p, q);
        T->irp[r] = ip;
        T->irp_sh[r] = pfo_shoup(ip, q);
        p = pfo_mulmod(p, T->psi, &T->mod);
        ip = pfo_mulmod(ip, psi_inv, &T->mod);
    }
    return 0;
}

void pfo_ntt_fwd(uint64_t *a, const pfo_ntt_tables *T) {
    const uint64_t n = T->n, q = T->mod.q, two_q = 2 * q;
    for (uint64_t m = 1; m < n; m <<= 1) {
        uint64_t gap = n / (2 * m);
        for (uint64_t i = 0; i < m; i++) {
            const uint64_t w = T->rp[m + i], wsh = T->rp_sh[m + i];
            uint64_t *x = a + 2 * i * gap, *y = x + gap;
            for (uint64_t j = 0; j < gap; j++) {
                uint64_t X = x[j];
                X -= (X >= two_q) ? two_q : 0;
                uint64_t Tm = mul_shoup_lazy(y[j], w, wsh, q);
                x[j] = X + Tm;
                y[j] = X - Tm + two_q;
            }
        }
    }
    for (uint64_t i = 0; i < n; i++) {
        uint64_t v = a[i];
        v -= (v >= two_q) ? two_q : 0;
        v -= (v >= q) ? q : 0;
        a[i] = v;
    }
}

void pfo_ntt_inv(uint64_t *a, const pfo_ntt_tables *T) {
    const uint64_t n = T->n, q = T->mod.q, two_q = 2 * q;
    for (uint64_t m = n >> 1; m >= 1; m >>= 1) {
        uint64_t gap = n / (2 * m);
        for (uint64_t i = 0; i < m; i++) {
            const uint64_t w = T->irp[m + i], wsh = T->irp_sh[m + i];
            uint64_t *x = a + 2 * i * gap, *y = x + gap;
            for (uint64_t j = 0; j < gap; j++) {
                uint64_t U = x[j], V = y[j]; /* both in [0, 2q) */
                uint64_t s = U + V;
                s -= (s >= two_q) ? two_q : 0;
                x[j] = s;
                y[j] = mul_shoup_lazy(U - V + two_q, w, wsh, q);
            }
        }
    }
    const uint64_t ninv = T->n_inv, ninv_sh = pfo_shoup(ninv, q);
    for (uint64_t i = 0; i < n; i++) a[i] = mul_shoup(a[i], ninv, ninv_sh, q);
}

const pfo_ntt_tables *pfo_tables(const pfo_context *c, int limb) { return limb < 0 ? &c->ntt_t : &c->ntt[limb]; }

/* ------------------------------------------------------------------------------------------
 * small fixed-width big numbers (little-endian 64-bit words) for Q = prod q_j
 * ------------------------------------------------------------------------------------------ */
#define BW (PFO_MAX_PRIMES + 2)

static void big_mul_word(uint64_t *a, int w, uint64_t x) { /* a[w] *= x, in place */
    uint64_t carry = 0;
    for (int i = 0; i < w; i++) {
        u128 p = (u128)a[i] * x + carry;
        a[i] = (uint64_t)p;
        carry = (uint64_t)(p >> 64);
    }
}
static void big_mul_word_to(const uint64_t *a, int w, uint64_t x, uint64_t *out) {
    uint64_t carry = 0;
    for (int i = 0; i < w; i++) {
        u128 p = (u128)a[i] * x + carry;
        out[i] = (uint64_t)p;
        carry = (uint64_t)(p >> 64);
    }
}
static void big_add(uint64_t *a, const uint64_t *b, int w) {
    unsigned carry = 0;
    for (int i = 0; i < w; i++) {
        u128 s = (u128)a[i] + b[i] + carry;
        a[i] = (uint64_t)s;
        carry = (unsigned)(s >> 64);
    }
}
static void big_sub(uint64_t *a, const uint64_t *b, int w) {
    unsigned borrow = 0;
    for (int i = 0; i < w; i++) {
        u128 d = (u128)a[i] - b[i] - borrow;
        a[i] = (uint64_t)d;
        borrow = (unsigned)((d >> 64) & 1);
    }
}
static int big_cmp(const uint64_t *a, const uint64_t *b, int w) {
    for (int i = w - 1; i >= 0; i--) {
        if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    }
    return 0;
}
static uint64_t big_div_word(uint64_t *a, int w, uint64_t d) { /* a /= d, returns remainder */
    u128 rem = 0;
    for (int i = w - 1; i >= 0; i--) {
        u128 cur = (rem << 64) | a[i];
        a[i] = (uint64_t)(cur / d);
        rem = cur % d;
    }
    return (uint64_t)rem;
}
static uint64_t big_mod_word(const uint64_t *a, int w, const pfo_modulus *m) {
    uint64_t r = 0;
    for (int i = w - 1; i >= 0; i--) r = pfo_barrett128(a[i], r, m);
    return r;
}
static long double big_to_ld(const uint64_t *a, int w) {
    long double r = 0;
    for (int i = w - 1; i >= 0; i--) r = r * 18446744073709551616.0L + (long double)a[i];
    return r;
}
static int big_bits(const uint64_t *a, int w) {
    for (int i = w - 1; i >= 0; i--) {
        if (a[i]) return 64 * i + (64 - __builtin_clzll(a[i]));
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * context — SEAL: context.cpp (SEALContext::validate), batchencoder.cpp ctor, util/rns.cpp
 * ------------------------------------------------------------------------------------------ */

pfo_context *pfo_context_create(uint64_t n, const uint64_t *primes, int k, uint64_t t) {
    if (k < 2 || k > PFO_MAX_PRIMES || n < 2 || (n & (n - 1))) return NULL;
    pfo_context *c = (pfo_context *)calloc(1, sizeof(pfo_context));
    if (!c) return NULL;
    c->n = n;
    c->logn = __builtin_ctzll(n);
    c->k = k;
    c->L = k - 1;
    for (int j = 0; j < k; j++) {
        if (pfo_modulus_init(&c->q[j], primes[j]) || ntt_tables_init(&c->ntt[j], n, c->logn, primes[j])) {
            pfo_context_destroy(c);
            return NULL;
        }
    }
    if (pfo_modulus_init(&c->t, t) || ntt_tables_init(&c->ntt_t, n, c->logn, t)) {
        pfo_context_destroy(c);
        return NULL;
    }
    /* SEAL: BatchEncoder::populate_matrix_reps_index_map */
    c->index_map = (uint64_t *)malloc(n * sizeof(uint64_t));
    {
        uint64_t row = n >> 1, m2 = 2 * n, pos = 1;
        for (uint64_t i = 0; i < row; i++) {
            uint64_t i1 = (pos - 1) >> 1, i2 = (m2 - pos - 1) >> 1;
            c->index_map[i] = pfo_bitrev((uint32_t)i1, c->logn);
            c->index_map[row | i] = pfo_bitrev((uint32_t)i2, c->logn);
            pos = (pos * 3) & (m2 - 1);
        }
    }
    /* Q = prod of data primes; floor(Q/t), Q mod t (SEAL: context.cpp coeff_div_plain_modulus) */
    int w = c->L + 1;
    c->qwords = w;
    memset(c->Qbig, 0, sizeof(c->Qbig));
    c->Qbig[0] = 1;
    for (int j = 0; j < c->L; j++) big_mul_word(c->Qbig, w, c->q[j].q);
    {
        uint64_t quo[BW];
        memcpy(quo, c->Qbig, sizeof(uint64_t) * w);
        c->q_mod_t = big_div_word(quo, w, t);
        for (int j = 0; j < c->L; j++) c->delta_mod_q[j] = big_mod_word(quo, w, &c->q[j]);
    }
    c->upper_half_threshold = (t + 1) >> 1;
    for (int j = 0; j < c->L; j++) {
        uint64_t tmp[BW];
        memcpy(tmp, c->Qbig, sizeof(uint64_t) * w);
        big_div_word(tmp, w, c->q[j].q);
        memcpy(c->qhat[j], tmp, sizeof(uint64_t) * w);
        c->qhat_inv[j] = pfo_invmod(big_mod_word(tmp, w, &c->q[j]), &c->q[j]);
    }
    /* special prime constants (SEAL: RNSTool::inv_q_last_mod_q at key level) */
    uint64_t P = c->q[k - 1].q;
    c->p_half = P >> 1;
    for (int j = 0; j < c->L; j++) {
        c->p_mod_q[j] = pfo_barrett64(P, &c->q[j]);
        c->p_inv_mod_q[j] = pfo_invmod(c->p_mod_q[j], &c->q[j]);
        c->p_half_mod_q[j] = pfo_barrett64(c->p_half, &c->q[j]);
    }
    return c;
}

void pfo_context_destroy(pfo_context *c) {
    if (!c) return;
    for (int j = 0; j < PFO_MAX_PRIMES; j++) free(c->ntt[j].rp);
    free(c->ntt_t.rp);
    free(c->index_map);
    free(c);
}

/* ------------------------------------------------------------------------------------------
 * BatchEncoder — SEAL: batchencoder.cpp encode()/decode()
 * ------------------------------------------------------------------------------------------ */

void pfo_batch_encode(const pfo_context *c, const uint64_t *values, uint64_t nvalues, uint64_t *plain) {
    memset(plain, 0, c->n * sizeof(uint64_t));
    for (uint64_t i = 0; i < nvalues && i < c->n; i++) plain[c->index_map[i]] = values[i];
    pfo_ntt_inv(plain, &c->ntt_t);
}

void pfo_batch_decode(const pfo_context *c, const uint64_t *plain, uint64_t *values) {
    uint64_t *tmp = (uint64_t *)malloc(c->n * sizeof(uint64_t));
    memcpy(tmp, plain, c->n * sizeof(uint64_t));
    pfo_ntt_fwd(tmp, &c->ntt_t);
    for (uint64_t i = 0; i < c->n; i++) values[i] = tmp[c->index_map[i]];
    free(tmp);
}

/* SEAL: Evaluator::transform_to_ntt_inplace(Plaintext&, parms_id) — centred lift per limb
 * (plain_upper_half_increment = q_j - t on the fast-plain-lift path, same residue otherwise). */
void pfo_plain_to_ntt(const pfo_context *c, const uint64_t *plain, uint64_t *out) {
    const uint64_t n = c->n, t = c->t.q, thr = c->upper_half_threshold;
    for (int j = 0; j < c->L; j++) {
        uint64_t *o = out + (size_t)j * n, inc = c->q[j].q - t;
        for (uint64_t i = 0; i < n; i++) o[i] = plain[i] >= thr ? plain[i] + inc : plain[i];
        pfo_ntt_fwd(o, &c->ntt[j]);
    }
}

/* SEAL: util/scalingvariant.cpp multiply_add_plain_with_scaling_variant */
void pfo_add_plain_scaled(const pfo_context *c, const uint64_t *plain, uint64_t *poly0) {
    const uint64_t n = c->n, t = c->t.q;
    for (uint64_t i = 0; i < n; i++) {
        u128 num = (u128)plain[i] * c->q_mod_t + c->upper_half_threshold;
        uint64_t fix = (uint64_t)(num / t);
        for (int j = 0; j < c->L; j++) {
            u128 v = (u128)plain[i] * c->delta_mod_q[j] + fix;
            uint64_t s = pfo_barrett128((uint64_t)v, (uint64_t)(v >> 64), &c->q[j]);
            uint64_t *p = poly0 + (size_t)j * n + i;
            *p = addmod(*p, s, c->q[j].q);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Keys / encrypt / decrypt — SEAL: keygenerator.cpp, util/rlwe.cpp (encrypt_zero_symmetric,
 * sample_poly_ternary / _uniform / _cbd), encryptor.cpp, decryptor.cpp.  Randomness comes from
 * the oracle's own RNG (not SEAL's Blake2xb PRNG): fresh randomness is not a parity quantity.
 * ------------------------------------------------------------------------------------------ */

static void sample_uniform(const pfo_context *c, int nlimbs_data, int include_special, uint64_t *s, uint64_t *out) {
    /* out[limbs][n], limbs = nlimbs_data (+1 if include_special, using prime index k-1) */
    int limbs = nlimbs_data + (include_special ? 1 : 0);
    for (int j = 0; j < limbs; j++) {
        int pj = (j == nlimbs_data) ? c->k - 1 : j;
        uint64_t q = c->q[pj].q, lim = UINT64_MAX - (UINT64_MAX % q) - 1;
        for (uint64_t i = 0; i < c->n; i++) {
            uint64_t r;
            do r = rng_next(s);
            while (r > lim);
            out[(size_t)j * c->n + i] = r % q;
        }
    }
}
static void sample_cbd(const pfo_context *c, uint64_t *s, int64_t *e) { /* SEAL sample_poly_cbd: 21+21 coins */
    for (uint64_t i = 0; i < c->n; i++) {
        uint64_t r = rng_next(s);
        e[i] = (int64_t)__builtin_popcountll(r & 0x1FFFFF) - (int64_t)__builtin_popcountll((r >> 21) & 0x1FFFFF);
    }
}
static inline uint64_t signed_to_mod(int64_t v, uint64_t q) { return v >= 0 ? (uint64_t)v : q - (uint64_t)(-v); }

void pfo_keygen(const pfo_context *c, uint64_t seed, uint64_t *sk) {
    uint64_t s = seed ^ 0x5ECBE7ULL;
    int8_t *tern = (int8_t *)malloc(c->n);
    for (uint64_t i = 0; i < c->n; i++) {
        uint64_t r;
        do r = rng_next(&s) & 3;
        while (r == 3);
        tern[i] = (int8_t)r - 1;
    }
    for (int j = 0; j < c->k; j++) {
        uint64_t *o = sk + (size_t)j * c->n;
        for (uint64_t i = 0; i < c->n; i++) o[i] = signed_to_mod(tern[i], c->q[j].q);
        pfo_ntt_fwd(o, &c->ntt[j]);
    }
    free(tern);
}

/* encrypt zero under sk over primes {0..nl-1} (+ special if with_special), NTT form:
 * c1 = a (uniform), c0 = -(a*s + e).  Layout out[2][limbs][n]. */
static void encrypt_zero_ntt(const pfo_context *c, const uint64_t *sk, int nl, int with_special, uint64_t *s,
                             uint64_t *out) {
    const uint64_t n = c->n;
    int limbs = nl + (with_special ? 1 : 0);
    uint64_t *c0 = out, *c1 = out + (size_t)limbs * n;
    int64_t *e = (int64_t *)malloc(n * sizeof(int64_t));
    sample_uniform(c, nl, with_special, s, c1);
    sample_cbd(c, s, e);
    for (int j = 0; j < limbs; j++) {
        int pj = (j == nl) ? c->k - 1 : j;
        const pfo_modulus *m = &c->q[pj];
        uint64_t *en = c0 + (size_t)j * n;
        for (uint64_t i = 0; i < n; i++) en[i] = signed_to_mod(e[i], m->q);
        pfo_ntt_fwd(en, &c->ntt[pj]);
        const uint64_t *a = c1 + (size_t)j * n, *sj = sk + (size_t)pj * n;
        for (uint64_t i = 0; i < n; i++) en[i] = negmod(addmod(pfo_mulmod(a[i], sj[i], m), en[i], m->q), m->q);
    }
    free(e);
}

/* SEAL: KeyGenerator::generate_one_kswitch_key with new_key = apply_galois_ntt(sk, elt) */
void pfo_galois_keygen(const pfo_context *c, const uint64_t *sk, uint32_t elt, uint64_t seed, uint64_t *key) {
    const uint64_t n = c->n;
    const int k = c->k, L = c->L;
    uint64_t s = seed ^ ((uint64_t)elt << 32) ^ 0x6A101EULL;
    uint64_t *rot = (uint64_t *)malloc((size_t)k * n * sizeof(uint64_t));
    for (int j = 0; j < k; j++) pfo_apply_galois_ntt(c, sk + (size_t)j * n, elt, rot + (size_t)j * n);
    for (int J = 0; J < L; J++) {
        uint64_t *kj = key + (size_t)J * 2 * k * n;
        encrypt_zero_ntt(c, sk, L, 1, &s, kj);
        /* c0 limb J += (P mod q_J) * sigma(s) limb J */
        const pfo_modulus *m = &c->q[J];
        uint64_t f = c->p_mod_q[J];
        uint64_t *dst = kj + (size_t)J * n;
        const uint64_t *r = rot + (size_t)J * n;
        for (uint64_t i = 0; i < n; i++) dst[i] = addmod(dst[i], pfo_mulmod(r[i], f, m), m->q);
    }
    free(rot);
}

/* SEAL: Encryptor::encrypt_symmetric (BFV): encrypt_zero_symmetric (coefficient form) then
 * multiply_add_plain_with_scaling_variant on c0 */
void pfo_encrypt_symmetric(const pfo_context *c, const uint64_t *sk, const uint64_t *plain, uint64_t seed,
                           uint64_t *ct) {
    uint64_t s = seed ^ 0xE2C0DEULL;
    encrypt_zero_ntt(c, sk, c->L, 0, &s, ct);
    pfo_ct_from_ntt(c, ct, 2);
    pfo_add_plain_scaled(c, plain, ct);
}

/* SEAL: Decryptor::bfv_decrypt + invariant_noise_budget.  x = c0 + c1*s mod Q composed to a big
 * integer per coefficient; m = round(t*x/Q) mod t (what RNSTool::decrypt_scale_and_round
 * computes); noise = |t*x - round(t*x/Q)*Q|. */
int pfo_decrypt(const pfo_context *c, const uint64_t *sk, const uint64_t *ct, uint64_t *plain) {
    const uint64_t n = c->n, t = c->t.q;
    const int L = c->L, w = c->qwords + 1;
    uint64_t *x = (uint64_t *)malloc((size_t)L * n * sizeof(uint64_t));
    for (int j = 0; j < L; j++) {
        const pfo_modulus *m = &c->q[j];
        uint64_t *xj = x + (size_t)j * n;
        memcpy(xj, ct + (size_t)(L + j) * n, n * sizeof(uint64_t));
        pfo_ntt_fwd(xj, &c->ntt[j]);
        const uint64_t *sj = sk + (size_t)j * n;
        for (uint64_t i = 0; i < n; i++) xj[i] = pfo_mulmod(xj[i], sj[i], m);
        pfo_ntt_inv(xj, &c->ntt[j]);
        const uint64_t *c0 = ct + (size_t)j * n;
        for (uint64_t i = 0; i < n; i++) xj[i] = addmod(xj[i], c0[i], m->q);
    }
    uint64_t Q[BW] = {0}, Qhalf[BW] = {0}, maxnoise[BW] = {0};
    memcpy(Q, c->Qbig, sizeof(uint64_t) * c->qwords);
    memcpy(Qhalf, Q, sizeof(Q));
    big_div_word(Qhalf, w, 2);
    const long double Qld = big_to_ld(Q, w);
    for (uint64_t i = 0; i < n; i++) {
        uint64_t X[BW] = {0}, tmp[BW];
        for (int j = 0; j < L; j++) {
            uint64_t y = pfo_mulmod(x[(size_t)j * n + i], c->qhat_inv[j], &c->q[j]);
            memset(tmp, 0, sizeof(tmp));
            big_mul_word_to(c->qhat[j], c->qwords, y, tmp);
            big_add(X, tmp, w);
        }
        while (big_cmp(X, Q, w) >= 0) big_sub(X, Q, w);
        big_mul_word(X, w, t); /* t*x < t*Q */
        uint64_t est = (uint64_t)(big_to_ld(X, w) / Qld);
        if (est >= t) est = t - 1;
        memset(tmp, 0, sizeof(tmp));
        big_mul_word_to(Q, w, est, tmp);
        while (big_cmp(tmp, X, w) > 0) {
            big_sub(tmp, Q, w);
            est--;
        }
        big_sub(X, tmp, w); /* remainder candidate */
        while (big_cmp(X, Q, w) >= 0) {
            big_sub(X, Q, w);
            est++;
        }
        uint64_t noise[BW];
        if (big_cmp(X, Qhalf, w) > 0) { /* round up */
            est++;
            memcpy(noise, Q, sizeof(noise));
            big_sub(noise, X, w);
        } else {
            memcpy(noise, X, sizeof(noise));
        }
        if (big_cmp(noise, maxnoise, w) > 0) memcpy(maxnoise, noise, sizeof(noise));
        plain[i] = est % t;
    }
    free(x);
    int budget = big_bits(Q, w) - big_bits(maxnoise, w) - 1;
    return budget < 0 ? 0 : budget;
}

/* ------------------------------------------------------------------------------------------
 * Evaluator — SEAL: evaluator.cpp
 * ------------------------------------------------------------------------------------------ */

void pfo_ct_to_ntt(const pfo_context *c, uint64_t *ct, int size) {
    for (int p = 0; p < size; p++)
        for (int j = 0; j < c->L; j++) pfo_ntt_fwd(ct + ((size_t)p * c->L + j) * c->n, &c->ntt[j]);
}
void pfo_ct_from_ntt(const pfo_context *c, uint64_t *ct, int size) {
    for (int p = 0; p < size; p++)
        for (int j = 0; j < c->L; j++) pfo_ntt_inv(ct + ((size_t)p * c->L + j) * c->n, &c->ntt[j]);
}

/* SEAL: Evaluator::multiply_plain_ntt — dyadic product of each ct polynomial with the plaintext */
void pfo_multiply_plain_ntt(const pfo_context *c, const uint64_t *ct, const uint64_t *pt_ntt, uint64_t *out) {
    const uint64_t n = c->n;
    for (int p = 0; p < 2; p++)
        for (int j = 0; j < c->L; j++) {
            const uint64_t *a = ct + ((size_t)p * c->L + j) * n, *b = pt_ntt + (size_t)j * n;
            uint64_t *o = out + ((size_t)p * c->L + j) * n;
            for (uint64_t i = 0; i < n; i++) o[i] = pfo_mulmod(a[i], b[i], &c->q[j]);
        }
}

/* SEAL: Evaluator::add_inplace */
void pfo_add(const pfo_context *c, uint64_t *a, const uint64_t *b) {
    const uint64_t n = c->n;
    for (int p = 0; p < 2; p++)
        for (int j = 0; j < c->L; j++) {
            size_t off = ((size_t)p * c->L + j) * n;
            for (uint64_t i = 0; i < n; i++) a[off + i] = addmod(a[off + i], b[off + i], c->q[j].q);
        }
}

/* sum_k multiply_plain_ntt(ct_k, pt_k) accumulated lazily in 128 bits, one Barrett-128 at the end —
 * equal to SEAL's multiply_plain + add_inplace chain because every step is exact mod q. */
void pfo_mac_plain_ntt(const pfo_context *c, const uint64_t *cts, const uint64_t *pts, size_t pt_stride, int K,
                       uint64_t *acc) {
    const uint64_t n = c->n;
    const size_t ct_stride = (size_t)2 * c->L * n;
    for (int p = 0; p < 2; p++)
        for (int j = 0; j < c->L; j++) {
            uint64_t *o = acc + ((size_t)p * c->L + j) * n;
            for (uint64_t i = 0; i < n; i++) {
                u128 s = 0;
                for (int k = 0; k < K; k++) {
                    uint64_t a = cts[(size_t)k * ct_stride + ((size_t)p * c->L + j) * n + i];
                    uint64_t b = pts[(size_t)k * pt_stride + (size_t)j * n + i];
                    s += (u128)a * b;
                    if ((k & 63) == 63) s = pfo_barrett128((uint64_t)s, (uint64_t)(s >> 64), &c->q[j]);
                }
                o[i] = pfo_barrett128((uint64_t)s, (uint64_t)(s >> 64), &c->q[j]);
            }
        }
}

/* SEAL: util/galois.cpp GaloisTool::get_elt_from_step */
uint32_t pfo_galois_elt_from_step(const pfo_context *c, int step) {
    uint32_t n = (uint32_t)c->n, m2 = 2 * n, row = n >> 1;
    if (step == 0) return m2 - 1;
    uint32_t pos = (uint32_t)(step < 0 ? -step : step);
    if (pos >= row) return 0;
    uint32_t s = step < 0 ? row - pos : pos;
    uint64_t e = 1;
    for (uint32_t i = 0; i < s; i++) e = (e * 3) & (m2 - 1);
    return (uint32_t)e;
}

/* SEAL: GaloisTool::apply_galois (coefficient form, one limb) */
void pfo_apply_galois(const pfo_context *c, const uint64_t *in, uint32_t elt, int limb, uint64_t *out) {
    const uint64_t n = c->n, q = c->q[limb].q;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t raw = i * elt, idx = raw & (n - 1);
        uint64_t v = in[i];
        if ((raw >> c->logn) & 1) v = negmod(v, q);
        out[idx] = v;
    }
}

/* SEAL: GaloisTool::generate_table_ntt */
void pfo_galois_ntt_table(const pfo_context *c, uint32_t elt, uint32_t *table) {
    const uint32_t n = (uint32_t)c->n;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t rev = pfo_bitrev(i, c->logn);
        uint64_t raw = ((uint64_t)elt * (2 * (uint64_t)rev + 1)) >> 1;
        table[i] = pfo_bitrev((uint32_t)(raw & (n - 1)), c->logn);
    }
}

/* SEAL: GaloisTool::apply_galois_ntt — pure permutation */
void pfo_apply_galois_ntt(const pfo_context *c, const uint64_t *in, uint32_t elt, uint64_t *out) {
    uint32_t *tab = (uint32_t *)malloc(c->n * sizeof(uint32_t));
    pfo_galois_ntt_table(c, elt, tab);
    for (uint64_t i = 0; i < c->n; i++) out[i] = in[tab[i]];
    free(tab);
}

/* SEAL: Evaluator::switch_key_inplace (BFV branch, ciphertext at the top data level, one special
 * prime).  key layout [J<L][component 2][prime k][n], NTT form. */
void pfo_switch_key(const pfo_context *c, uint64_t *ct, const uint64_t *target, const uint64_t *key) {
    const uint64_t n = c->n;
    const int L = c->L, k = c->k, rns = L + 1;
    uint64_t *prod = (uint64_t *)malloc((size_t)2 * rns * n * sizeof(uint64_t)); /* [comp][I][n] */
    uint64_t *tntt = (uint64_t *)malloc(n * sizeof(uint64_t));
    u128 *lazy = (u128 *)malloc((size_t)2 * n * sizeof(u128));
    for (int I = 0; I < rns; I++) {
        int ki = (I == L) ? k - 1 : I;
        const pfo_modulus *mi = &c->q[ki];
        memset(lazy, 0, (size_t)2 * n * sizeof(u128));
        for (int J = 0; J < L; J++) {
            const uint64_t *tj = target + (size_t)J * n;
            if (c->q[J].q <= mi->q)
                memcpy(tntt, tj, n * sizeof(uint64_t));
            else
                for (uint64_t i = 0; i < n; i++) tntt[i] = pfo_barrett64(tj[i], mi);
            pfo_ntt_fwd(tntt, &c->ntt[ki]);
            for (int comp = 0; comp < 2; comp++) {
                const uint64_t *kp = key + (((size_t)J * 2 + comp) * k + ki) * n;
                u128 *lz = lazy + (size_t)comp * n;
                for (uint64_t i = 0; i < n; i++) lz[i] += (u128)tntt[i] * kp[i];
            }
        }
        for (int comp = 0; comp < 2; comp++) {
            uint64_t *o = prod + ((size_t)comp * rns + I) * n;
            const u128 *lz = lazy + (size_t)comp * n;
            for (uint64_t i = 0; i < n; i++) o[i] = pfo_barrett128((uint64_t)lz[i], (uint64_t)(lz[i] >> 64), mi);
        }
    }
    const pfo_modulus *mp = &c->q[k - 1];
    for (int comp = 0; comp < 2; comp++) {
        uint64_t *last = prod + ((size_t)comp * rns + L) * n;
        pfo_ntt_inv(last, &c->ntt[k - 1]);
        for (uint64_t i = 0; i < n; i++) last[i] = pfo_barrett64(last[i] + c->p_half, mp);
        for (int j = 0; j < L; j++) {
            const pfo_modulus *mj = &c->q[j];
            uint64_t *sj = prod + ((size_t)comp * rns + j) * n;
            pfo_ntt_inv(sj, &c->ntt[j]);
            uint64_t *dst = ct + ((size_t)comp * L + j) * n;
            for (uint64_t i = 0; i < n; i++) {
                uint64_t v = submod(pfo_barrett64(last[i], mj), c->p_half_mod_q[j], mj->q);
                uint64_t o = pfo_mulmod(submod(sj[i], v, mj->q), c->p_inv_mod_q[j], mj);
                dst[i] = addmod(dst[i], o, mj->q);
            }
        }
    }
    free(lazy);
    free(tntt);
    free(prod);
}

/* SEAL: Evaluator::apply_galois_inplace (BFV): c0' = sigma(c0); key-switch sigma(c1) into (c0', 0) */
void pfo_apply_galois_ct(const pfo_context *c, uint64_t *ct, uint32_t elt, const uint64_t *key) {
    const uint64_t n = c->n;
    const int L = c->L;
    uint64_t *tmp = (uint64_t *)malloc((size_t)L * n * sizeof(uint64_t));
    for (int j = 0; j < L; j++) pfo_apply_galois(c, ct + (size_t)j * n, elt, j, tmp + (size_t)j * n);
    memcpy(ct, tmp, (size_t)L * n * sizeof(uint64_t));
    for (int j = 0; j < L; j++) pfo_apply_galois(c, ct + (size_t)(L + j) * n, elt, j, tmp + (size_t)j * n);
    memset(ct + (size_t)L * n, 0, (size_t)L * n * sizeof(uint64_t));
    pfo_switch_key(c, ct, tmp, key);
    free(tmp);
}

/* SEAL: Evaluator::mod_switch_scale_to_next (BFV) = RNSTool::divide_and_round_q_last_inplace */
void pfo_mod_switch_next(const pfo_context *c, const uint64_t *ct, int Lin, uint64_t *out) {
    const uint64_t n = c->n;
    const int Lout = Lin - 1;
    const pfo_modulus *ml = &c->q[Lin - 1];
    const uint64_t half = ml->q >> 1;
    uint64_t *last = (uint64_t *)malloc(n * sizeof(uint64_t));
    for (int p = 0; p < 2; p++) {
        const uint64_t *src = ct + (size_t)p * Lin * n;
        for (uint64_t i = 0; i < n; i++) last[i] = addmod(src[(size_t)(Lin - 1) * n + i], half, ml->q);
        for (int j = 0; j < Lout; j++) {
            const pfo_modulus *mj = &c->q[j];
            uint64_t half_mod = pfo_barrett64(half, mj);
            uint64_t inv = pfo_invmod(pfo_barrett64(ml->q, mj), mj);
            uint64_t *o = out + ((size_t)p * Lout + j) * n;
            for (uint64_t i = 0; i < n; i++) {
                uint64_t tmp = submod(pfo_barrett64(last[i], mj), half_mod, mj->q);
                o[i] = pfo_mulmod(submod(src[(size_t)j * n + i], tmp, mj->q), inv, mj);
            }
        }
    }
    free(last);
}

/* ------------------------------------------------------------------------------------------
 * Wire format — SEAL: serialization.h (SEALHeader), ciphertext.cpp (save_members/load_members),
 * dynarray.h (DynArray::save_members); compr_mode_type::none only.
 * ------------------------------------------------------------------------------------------ */

static void put_header(uint8_t *p, uint64_t total) {
    p[0] = 0x5E;
    p[1] = 0xA1; /* magic 0xA15E little-endian */
    p[2] = 0x10; /* header size */
    p[3] = 4;    /* version major */
    p[4] = 1;    /* version minor */
    p[5] = 0;    /* compr_mode none */
    p[6] = p[7] = 0;
    memcpy(p + 8, &total, 8);
}

size_t pfo_ct_save_size(uint64_t n, int L, int size) {
    return 16 + 32 + 1 + 8 * 5 + 16 + 8 + (size_t)size * L * n * 8;
}

size_t pfo_ct_save(const uint64_t *ct, uint64_t n, int L, int size, int is_ntt, const uint64_t parms_id[4],
                   uint8_t *out) {
    size_t total = pfo_ct_save_size(n, L, size), words = (size_t)size * L * n;
    uint8_t *p = out;
    put_header(p, total);
    p += 16;
    memcpy(p, parms_id, 32);
    p += 32;
    *p++ = (uint8_t)(is_ntt ? 1 : 0);
    uint64_t v;
    v = (uint64_t)size;
    memcpy(p, &v, 8);
    p += 8;
    v = n;
    memcpy(p, &v, 8);
    p += 8;
    v = (uint64_t)L;
    memcpy(p, &v, 8);
    p += 8;
    double scale = 1.0;
    memcpy(p, &scale, 8);
    p += 8;
    v = 1; /* correction_factor */
    memcpy(p, &v, 8);
    p += 8;
    put_header(p, 16 + 8 + words * 8);
    p += 16;
    v = words;
    memcpy(p, &v, 8);
    p += 8;
    memcpy(p, ct, words * 8);
    p += words * 8;
    return (size_t)(p - out);
}

static int check_header(const uint8_t *p, size_t len, uint64_t *total) {
    if (len < 16 || p[0] != 0x5E || p[1] != 0xA1 || p[2] != 0x10 || p[3] != 4 || p[5] != 0) return -1;
    memcpy(total, p + 8, 8);
    return (*total <= len && *total >= 16) ? 0 : -1;
}

size_t pfo_ct_load(const uint8_t *in, size_t len, uint64_t *n, int *L, int *size, int *is_ntt, uint64_t parms_id[4],
                   uint64_t *ct, size_t ct_cap_words) {
    uint64_t total, inner, v, words;
    if (check_header(in, len, &total)) return 0;
    if (total < 16 + 32 + 1 + 40 + 16 + 8) return 0;
    const uint8_t *p = in + 16;
    memcpy(parms_id, p, 32);
    p += 32;
    *is_ntt = *p++ ? 1 : 0;
    memcpy(&v, p, 8);
    *size = (int)v;
    p += 8;
    memcpy(n, p, 8);
    p += 8;
    memcpy(&v, p, 8);
    *L = (int)v;
    p += 8;
    p += 16; /* scale, correction_factor */
    if (check_header(p, total - (size_t)(p - in), &inner)) return 0;
    p += 16;
    memcpy(&words, p, 8);
    p += 8;
    if (words != (uint64_t)*size * (uint64_t)*L * *n || words > ct_cap_words) return 0;
    if ((size_t)(p - in) + words * 8 > total) return 0;
    memcpy(ct, p, words * 8);
    return (size_t)total;
}

/* ------------------------------------------------------------------------------------------
 * Plaintext reference path
 * ------------------------------------------------------------------------------------------ */

/* ref: include/common/client_server_utils.h:24-56 (vecs_read<T>).  Returns rows compacted [n][d]. */
int pfo_vecs_read(const char *fname, size_t *d_out, size_t *n_out, void **data_out) {
    FILE *f = fopen(fname, "rb");
    if (!f) return -1;
    int d = 0;
    if (fread(&d, 1, sizeof(int), f) != sizeof(int) || d <= 0 || d >= 1000000) {
        fclose(f);
        return -2;
    }
    fseek(f, 0, SEEK_SET);
    struct stat st;
    fstat(fileno(f), &st);
    size_t sz = (size_t)st.st_size;
    if (sz % ((size_t)(d + 1) * 4)) {
        fclose(f);
        return -3;
    }
    size_t n = sz / ((size_t)(d + 1) * 4);
    uint32_t *raw = (uint32_t *)malloc(n * (size_t)(d + 1) * 4);
    if (fread(raw, 4, n * (size_t)(d + 1), f) != n * (size_t)(d + 1)) {
        free(raw);
        fclose(f);
        return -4;
    }
    fclose(f);
    for (size_t i = 0; i < n; i++) memmove(raw + i * d, raw + 1 + i * (size_t)(d + 1), (size_t)d * 4);
    *d_out = (size_t)d;
    *n_out = n;
    *data_out = raw;
    return 0;
}

/* ref: src/client/client_lib.cpp:59-62 and src/server/server_lib.cpp:153-160 —
 * `float dist += std::pow(float - float, 2)`: the difference is a float, std::pow(float,int)
 * promotes to double (exact square of a float), the sum is formed in double and rounded back to
 * float every step. */
float pfo_l2sqr_ref(const float *a, const float *b, size_t d) {
    float dist = 0.0f;
    for (size_t k = 0; k < d; k++) {
        float diff = a[k] - b[k];
        double p = (double)diff * (double)diff;
        dist = (float)((double)dist + p);
    }
    return dist;
}

typedef struct {
    float dist;
    int64_t idx;
} dist_idx;
static int cmp_dist_idx(const void *a, const void *b) {
    const dist_idx *x = (const dist_idx *)a, *y = (const dist_idx *)b;
    if (x->dist < y->dist) return -1;
    if (x->dist > y->dist) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx); /* reference sort is unstable: ties pinned by index */
}

/* ref: src/client/client_lib.cpp:50-81 (distance to every centroid, ascending sort) and :93-103
 * (first NPROBE ids).  Distance is (query - centroid) as at :61. */
void pfo_coarse_quantize(size_t nq, size_t d, size_t nlist, const float *x, const float *centroids, size_t nprobe,
                         int64_t *out_idx, float *out_dist) {
    dist_idx *v = (dist_idx *)malloc(nlist * sizeof(dist_idx));
    for (size_t i = 0; i < nq; i++) {
        for (size_t j = 0; j < nlist; j++) {
            v[j].dist = pfo_l2sqr_ref(x + i * d, centroids + j * d, d);
            v[j].idx = (int64_t)j;
        }
        qsort(v, nlist, sizeof(dist_idx), cmp_dist_idx);
        for (size_t p = 0; p < nprobe; p++) {
            out_idx[i * nprobe + p] = v[p].idx;
            if (out_dist) out_dist[i * nprobe + p] = v[p].dist;
        }
    }
    free(v);
}

/* ref: src/server/server_lib.cpp:111-138 — for each query, for each of its given lists in order,
 * one (distance, id) per stored vector, packed back to back, list_sizes[i] = sum of list lengths
 * (consumer contract: src/client/client_lib.cpp:129-148).  The distance is the exact squared L2
 * of src/server/server_lib.cpp:140-167 (the semantics the HE evaluation reproduces); the fork's
 * PQ-ADC approximation is restated separately: pfo_search_lists_pq below. */
size_t pfo_search_lists_plain(size_t nq, size_t d, const float *x, const int64_t *idx, size_t nprobe,
                              const int64_t *list_offsets, const int64_t *ids, const float *vectors, float *dist,
                              int64_t *labels, size_t cap, size_t *list_sizes) {
    size_t w = 0;
    for (size_t i = 0; i < nq; i++) {
        size_t cnt = 0;
        for (size_t p = 0; p < nprobe; p++) {
            int64_t l = idx[i * nprobe + p];
            if (l < 0) continue;
            for (int64_t o = list_offsets[l]; o < list_offsets[l + 1]; o++) {
                if (w < cap) {
                    dist[w] = pfo_l2sqr_ref(vectors + (size_t)o * d, x + i * d, d);
                    labels[w] = ids[o];
                }
                w++;
                cnt++;
            }
        }
        list_sizes[i] = cnt;
    }
    return w;
}

/* What m_Index->search_encrypted computes in the reference TODAY (ref: src/server/server_lib.cpp:126-130 on the
 * faiss::IndexIVFPQ built at :34-36 with SUB_QUANTIZERS sub-quantizers of SUB_QUANTIZER_SIZE bits,
 * include/common/client_server_utils.h:19-20): the product-quantizer asymmetric distance (ADC) of the query to every
 * stored code of every given list, all of them returned (no top-k), packed like pfo_search_lists_plain.
 * [EXT] The fork's source is absent (PES-Innovation-Lab/PreFHEtch-faiss @ 49c5b57c); this restates the
 * published FAISS algorithm for IndexIVFPQ with by_residual = true and METRIC_L2 in its defining form
 * (faiss/IndexIVFPQ.cpp, IVFPQScannerT::precompute_list_tables with use_precomputed_table = 0 and scan_list_with_table;
 * faiss/impl/ProductQuantizer.cpp compute_distance_table; faiss/utils/distances_simd.cpp fvec_L2sqr_ref):
 *     r        = x - centroid[l]                                   (float)
 *     tab[m][j] = sum_k (r[m dsub + k] - pq[m][j][k])^2            (float, k ascending, multiply then add)
 *     dis(code) = sum_m tab[m][code[m]]                            (float, m ascending, from 0)
 * FAISS's default build evaluates the same sums with SIMD partial sums (and, when the table fits, through its
 * precomputed ||c||^2 + 2<c, pq> tables), which rounds differently: equality with a FAISS build is to float rounding,
 * not bit for bit.  PARITY UNPINNED by the reference.  nbits = 8 (one byte per sub-quantizer), as the reference builds it.
 * pq_centroids [M][256][dsub] (FAISS ProductQuantizer::centroids layout), codes [ntotal][M] in list order. */
__attribute__((optimize("fp-contract=off")))
size_t pfo_search_lists_pq(size_t nq, size_t d, const float *x, const int64_t *idx, size_t nprobe, const float *centroids,
                           const int64_t *list_offsets, const int64_t *ids, size_t M, const float *pq_centroids,
                           const uint8_t *codes, float *dist, int64_t *labels, size_t cap, size_t *list_sizes) {
    const size_t ksub = 256, dsub = d / M;
    float *r = (float *)malloc(d * sizeof(float));
    float *tab = (float *)malloc(M * ksub * sizeof(float));
    size_t w = 0;
    for (size_t i = 0; i < nq; i++) {
        size_t cnt = 0;
        for (size_t p = 0; p < nprobe; p++) {
            int64_t l = idx[i * nprobe + p];
            if (l < 0) continue;
            if (list_offsets[l + 1] > list_offsets[l]) {
                for (size_t k = 0; k < d; k++) r[k] = x[i * d + k] - centroids[(size_t)l * d + k];
                for (size_t m = 0; m < M; m++)
                    for (size_t j = 0; j < ksub; j++) {
                        const float *c = pq_centroids + (m * ksub + j) * dsub;
                        float acc = 0.0f;
                        for (size_t k = 0; k < dsub; k++) {
                            const float t = r[m * dsub + k] - c[k];
                            const float t2 = t * t;
                            acc = acc + t2;
                        }
                        tab[m * ksub + j] = acc;
                    }
            }
            for (int64_t o = list_offsets[l]; o < list_offsets[l + 1]; o++) {
                if (w < cap) {
                    const uint8_t *code = codes + (size_t)o * M;
                    float acc = 0.0f;
                    for (size_t m = 0; m < M; m++) acc = acc + tab[m * ksub + code[m]];
                    dist[w] = acc;
                    labels[w] = ids[o];
                }
                w++;
                cnt++;
            }
        }
        list_sizes[i] = cnt;
    }
    free(r);
    free(tab);
    return w;
}

/* FAISS ProductQuantizer::compute_code for nbits = 8 (faiss/impl/ProductQuantizer.cpp: nearest sub-centroid per
 * sub-vector by fvec_L2sqr, first minimum wins) applied to the residual of a vector to the centroid of its list —
 * what IndexIVFPQ::add stores (ref: src/server/server_lib.cpp:80).  Test / bench data generation only. */
__attribute__((optimize("fp-contract=off")))
void pfo_pq_encode_residuals(size_t n, size_t d, const float *vectors, const int64_t *list_of, const float *centroids, size_t M,
                             const float *pq_centroids, uint8_t *codes) {
    const size_t ksub = 256, dsub = d / M;
    for (size_t v = 0; v < n; v++)
        for (size_t m = 0; m < M; m++) {
            float best = 0.0f;
            size_t bj = 0;
            for (size_t j = 0; j < ksub; j++) {
                const float *c = pq_centroids + (m * ksub + j) * dsub;
                float acc = 0.0f;
                for (size_t k = 0; k < dsub; k++) {
                    const float rk = vectors[v * d + m * dsub + k] - centroids[(size_t)list_of[v] * d + m * dsub + k];
                    const float t = rk - c[k];
                    const float t2 = t * t;
                    acc = acc + t2;
                }
                if (j == 0 || acc < best) best = acc, bj = j;
            }
            codes[v * M + m] = (uint8_t)bj;
        }
}

/* ref: src/client/client_lib.cpp:272-291,325-330.  returned[nq][k_ret], gt[nq][gt_k].  The
 * reference loops j,k < K with K = k_ret <= gt_k. */
void pfo_recall(size_t nq, size_t k_ret, const int64_t *returned, size_t gt_k, const int32_t *gt, double *r1,
                double *r10, double *r100, double *std10, double *mrr10) {
    long c1 = 0, c10 = 0, c100 = 0, s10 = 0;
    double m10 = 0;
    for (size_t i = 0; i < nq; i++) {
        for (size_t j = 0; j < k_ret && j < gt_k; j++) {
            for (size_t k = 0; k < k_ret; k++) {
                if ((int64_t)gt[i * gt_k + j] == returned[i * k_ret + k]) {
                    if (k < 1) c1++;
                    if (k < 10) c10++;
                    if (k < 100) c100++;
                    if (j == 0 && k < 10) m10 += 1.0 / (double)(k + 1);
                    if (j < 10 && k < 10) s10++;
                    break;
                }
            }
        }
    }
    if (r1) *r1 = (double)c1 / (1.0 * nq);
    if (r10) *r10 = (double)c10 / (10.0 * nq);
    if (r100) *r100 = (double)c100 / (100.0 * nq);
    if (std10) *std10 = (double)s10 / (10.0 * nq);
    if (mrr10) *mrr10 = m10 / (double)nq;
}

/* ------------------------------------------------------------------------------------------
 * Encrypted-distance layout (SURVEY.md §7.1 "generalised diagonal with partial-sum factor g").
 * This layout is defined by this repository (the reference has no HE protocol); the oracle and
 * the CUDA engine implement it independently.
 * ------------------------------------------------------------------------------------------ */

int pfo_layout_init(pfo_layout *lay, uint64_t n, uint32_t d, uint32_t m, uint32_t g) {
    uint32_t dp = 1;
    while (dp < d) dp <<= 1;
    if (!m || !g || (m & (m - 1)) || (g & (g - 1)) || dp % m) return -1;
    uint32_t dc = dp / m;
    if (dc % g || dc > n / 2) return -1;
    lay->n = n;
    lay->d = d;
    lay->d_pad = dp;
    lay->m = m;
    lay->g = g;
    lay->dc = dc;
    lay->R = dc / g;
    lay->K = m * lay->R;
    lay->C = (uint32_t)(n / g);
    return 0;
}

static inline uint64_t int_to_mod(int64_t v, uint64_t t) {
    int64_t r = v % (int64_t)t;
    return (uint64_t)(r < 0 ? r + (int64_t)t : r);
}

void pfo_layout_query_slots(const pfo_layout *lay, uint64_t t, const int64_t *q, uint32_t a, uint64_t *slots) {
    uint32_t S = (uint32_t)(lay->n / 2);
    for (uint32_t row = 0; row < 2; row++)
        for (uint32_t s = 0; s < S; s++) {
            uint32_t dim = a * lay->dc + (s % lay->dc);
            slots[row * S + s] = dim < lay->d ? int_to_mod(q[dim], t) : 0;
        }
}

uint32_t pfo_layout_slot(const pfo_layout *lay, uint32_t u, uint32_t j) {
    uint32_t S = (uint32_t)(lay->n / 2), per_row = S / lay->g;
    uint32_t row = u / per_row, c = u % per_row;
    uint32_t s0 = (c % lay->R) + (c / lay->R) * lay->dc;
    return row * S + s0 + j * lay->R;
}

/* owner of slot s within a row: candidate c with (s mod dc) mod R == c mod R, (s div dc) == c div R */
void pfo_layout_diag_slots(const pfo_layout *lay, uint64_t t, const int32_t *xs, uint32_t nvec, uint32_t a,
                           uint32_t r, uint64_t *slots) {
    uint32_t S = (uint32_t)(lay->n / 2), per_row = S / lay->g;
    for (uint32_t row = 0; row < 2; row++)
        for (uint32_t s = 0; s < S; s++) {
            uint32_t within = s % lay->dc, grp = s / lay->dc;
            uint32_t c = grp * lay->R + (within % lay->R);
            uint32_t u = row * per_row + c;
            uint32_t dim = a * lay->dc + ((within + r) % lay->dc);
            int64_t v = (u < nvec && dim < lay->d) ? -2 * (int64_t)xs[(size_t)u * lay->d + dim] : 0;
            slots[row * S + s] = int_to_mod(v, t);
        }
}

void pfo_layout_norm_slots(const pfo_layout *lay, uint64_t t, const int32_t *xs, uint32_t nvec, uint64_t *slots) {
    memset(slots, 0, lay->n * sizeof(uint64_t));
    for (uint32_t u = 0; u < nvec && u < lay->C; u++) {
        int64_t s = 0;
        for (uint32_t k = 0; k < lay->d; k++) s += (int64_t)xs[(size_t)u * lay->d + k] * xs[(size_t)u * lay->d + k];
        slots[pfo_layout_slot(lay, u, 0)] = int_to_mod(s, t);
    }
}

void pfo_encode_block(const pfo_context *c, const pfo_layout *lay, const int32_t *xs, uint32_t nvec, uint64_t *diag,
                      uint64_t *norm) {
    const uint64_t n = c->n;
    uint64_t *slots = (uint64_t *)malloc(n * sizeof(uint64_t));
    uint64_t *plain = (uint64_t *)malloc(n * sizeof(uint64_t));
    for (uint32_t a = 0; a < lay->m; a++)
        for (uint32_t r = 0; r < lay->R; r++) {
            pfo_layout_diag_slots(lay, c->t.q, xs, nvec, a, r, slots);
            pfo_batch_encode(c, slots, n, plain);
            pfo_plain_to_ntt(c, plain, diag + (size_t)(a * lay->R + r) * c->L * n);
        }
    pfo_layout_norm_slots(lay, c->t.q, xs, nvec, slots);
    pfo_batch_encode(c, slots, n, plain);
    memset(norm, 0, (size_t)c->L * n * sizeof(uint64_t));
    pfo_add_plain_scaled(c, plain, norm);
    for (int j = 0; j < c->L; j++) pfo_ntt_fwd(norm + (size_t)j * n, &c->ntt[j]);
    free(plain);
    free(slots);
}

/* SEAL op sequence: for r>=1  rot_r = rotate_rows(ct_a, r, gk)   [or chained step-1 rotations],
 * then transform_to_ntt_inplace on every member. */
void pfo_rotate_query_set(const pfo_context *c, const pfo_layout *lay, const uint64_t *cts,
                          const uint64_t *const *keys, int chain, uint64_t *rot) {
    const size_t ctw = (size_t)2 * c->L * c->n;
    for (uint32_t a = 0; a < lay->m; a++) {
        uint64_t *base = rot + (size_t)a * lay->R * ctw;
        memcpy(base, cts + (size_t)a * ctw, ctw * sizeof(uint64_t));
        for (uint32_t r = 1; r < lay->R; r++) {
            uint64_t *dst = base + (size_t)r * ctw;
            if (chain) {
                memcpy(dst, dst - ctw, ctw * sizeof(uint64_t));
                pfo_apply_galois_ct(c, dst, pfo_galois_elt_from_step(c, 1), keys[0]);
            } else {
                memcpy(dst, base, ctw * sizeof(uint64_t));
                pfo_apply_galois_ct(c, dst, pfo_galois_elt_from_step(c, (int)r), keys[r - 1]);
            }
        }
        for (uint32_t r = 0; r < lay->R; r++) pfo_ct_to_ntt(c, base + (size_t)r * ctw, 2);
    }
}

/* SEAL op sequence: acc = sum_k multiply_plain(rot_k, diag_k); transform_from_ntt_inplace(acc);
 * add_plain_inplace(acc, norm_plain).  The norm arrives pre-scaled in NTT form, which is the
 * same value because the NTT is linear and every step is exact mod q. */
void pfo_block_distance(const pfo_context *c, const pfo_layout *lay, const uint64_t *rot, const uint64_t *diag,
                        const uint64_t *norm, uint64_t *out) {
    const uint64_t n = c->n;
    pfo_mac_plain_ntt(c, rot, diag, (size_t)c->L * n, (int)lay->K, out);
    for (int j = 0; j < c->L; j++)
        for (uint64_t i = 0; i < n; i++)
            out[(size_t)j * n + i] = addmod(out[(size_t)j * n + i], norm[(size_t)j * n + i], c->q[j].q);
    pfo_ct_from_ntt(c, out, 2);
}
