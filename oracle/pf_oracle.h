/*
 * pf_oracle.h — CPU restatement ("oracle") of PreFHEtch's server-side search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under prefhetch_b200/ links, imports or executes this
 * code; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs do, and there only as the checker or the CPU baseline.
 *
 * HE half: PARITY UNPINNED by the reference — the reference snapshot holds no HE code, no tests and no
 * golden vectors.  The plaintext half of this file restates in-tree
 * reference code (cited per function, paths relative to /root/reference) and IS pinned: the reference's
 * own functions, compiled from its sources (oracle/ref_build -> oracle/_ref), produce
 * tests/golden/ref_plain_v1.json, which tests/test_ref_pin.py holds this file to bit for bit.  The HE half
 * restates the published algorithms of Microsoft SEAL 4.1 (pinned by the reference at
 * commit 7a931d55ba84a40b85938f6ca3ac206f18654093, CMakeLists.txt:33-38; source NOT in the
 * reference tree) — BFV evaluator semantics, negacyclic Harvey NTT with the numerically
 * smallest primitive 2N-th root, BatchEncoder slot map, Galois automorphisms, hybrid key
 * switching with one special prime, and the uncompressed wire format.  It is pinned instead
 * by math-level known-answer vectors generated with an independent Python big-integer
 * implementation (tests/golden/make_golden.py).
 */
#ifndef PF_ORACLE_H
#define PF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFO_MAX_PRIMES 16

typedef struct {
    uint64_t q;
    uint64_t ratio[2]; /* floor(2^128 / q): [0] low word, [1] high word (SEAL Modulus::const_ratio) */
} pfo_modulus;

typedef struct {
    uint64_t n;
    int logn;
    pfo_modulus mod;
    uint64_t psi;       /* numerically smallest primitive 2n-th root of unity mod q */
    uint64_t n_inv;     /* n^{-1} mod q */
    uint64_t *rp;       /* rp[i]  = psi^{bitrev(i)}        (i in [0,n)) */
    uint64_t *rp_sh;    /* Shoup quotients floor(rp[i] * 2^64 / q) */
    uint64_t *irp;      /* irp[i] = psi^{-bitrev(i)} */
    uint64_t *irp_sh;
} pfo_ntt_tables;

typedef struct pfo_context {
    uint64_t n;
    int logn;
    int k; /* number of primes in the key-level modulus chain (data primes + special prime) */
    int L; /* data limbs = k-1 (ciphertexts at SEAL's first_parms_id) */
    pfo_modulus q[PFO_MAX_PRIMES];
    pfo_ntt_tables ntt[PFO_MAX_PRIMES];
    pfo_modulus t;
    pfo_ntt_tables ntt_t;
    uint64_t *index_map; /* BatchEncoder matrix_reps_index_map_ (size n) */
    /* scaling variant (SEAL util/scalingvariant.cpp), Q = q_0..q_{L-1} */
    uint64_t delta_mod_q[PFO_MAX_PRIMES]; /* floor(Q/t) mod q_j */
    uint64_t q_mod_t;                     /* Q mod t */
    uint64_t upper_half_threshold;        /* (t+1)>>1 */
    /* key switching (special prime P = q[k-1]) */
    uint64_t p_mod_q[PFO_MAX_PRIMES];     /* P mod q_j */
    uint64_t p_inv_mod_q[PFO_MAX_PRIMES]; /* P^{-1} mod q_j */
    uint64_t p_half;                      /* P >> 1 */
    uint64_t p_half_mod_q[PFO_MAX_PRIMES];
    /* big-number view of Q for decryption */
    int qwords;
    uint64_t Qbig[PFO_MAX_PRIMES + 1];
    uint64_t qhat_inv[PFO_MAX_PRIMES];               /* (Q/q_j)^{-1} mod q_j */
    uint64_t qhat[PFO_MAX_PRIMES][PFO_MAX_PRIMES + 1]; /* Q/q_j as words */
} pfo_context;

/* ---- modular arithmetic (SEAL util/uintarithsmallmod.h) ---- */
int pfo_modulus_init(pfo_modulus *m, uint64_t q);
uint64_t pfo_barrett64(uint64_t x, const pfo_modulus *m);
uint64_t pfo_barrett128(uint64_t lo, uint64_t hi, const pfo_modulus *m);
uint64_t pfo_mulmod(uint64_t a, uint64_t b, const pfo_modulus *m);
uint64_t pfo_powmod(uint64_t a, uint64_t e, const pfo_modulus *m);
uint64_t pfo_invmod(uint64_t a, const pfo_modulus *m); /* prime modulus */
uint64_t pfo_shoup(uint64_t w, uint64_t q);            /* floor(w * 2^64 / q) */
uint64_t pfo_minimal_primitive_root(uint64_t two_n, const pfo_modulus *m); /* 0 on failure */
uint32_t pfo_bitrev(uint32_t x, int bits);

/* ---- context ---- */
pfo_context *pfo_context_create(uint64_t n, const uint64_t *primes, int k, uint64_t t);
void pfo_context_destroy(pfo_context *c);

/* ---- NTT (SEAL util/ntt.cpp, dwthandler.h): in place, natural -> bit-reversed and back;
 * stored results fully reduced to [0,q). */
void pfo_ntt_fwd(uint64_t *a, const pfo_ntt_tables *T);
void pfo_ntt_inv(uint64_t *a, const pfo_ntt_tables *T);
const pfo_ntt_tables *pfo_tables(const pfo_context *c, int limb); /* limb == -1 -> plain modulus */

/* ---- BatchEncoder (SEAL batchencoder.cpp) ---- */
void pfo_batch_encode(const pfo_context *c, const uint64_t *values, uint64_t nvalues, uint64_t *plain);
void pfo_batch_decode(const pfo_context *c, const uint64_t *plain, uint64_t *values);

/* ---- plaintext <-> RNS ---- */
/* Evaluator::transform_to_ntt_inplace(Plaintext, parms_id): centred lift then NTT; out[L][n] */
void pfo_plain_to_ntt(const pfo_context *c, const uint64_t *plain, uint64_t *out);
/* scalingvariant.cpp multiply_add_plain_with_scaling_variant: poly0[L][n] += round(Q/t * m) */
void pfo_add_plain_scaled(const pfo_context *c, const uint64_t *plain, uint64_t *poly0);

/* ---- keys, encryption, decryption (client side; test infrastructure) ---- */
/* secret key stored in NTT form over all k primes: sk[k][n] */
void pfo_keygen(const pfo_context *c, uint64_t seed, uint64_t *sk);
/* Galois key for element elt: key[L][2][k][n] (NTT form), KeyGenerator::generate_one_kswitch_key */
void pfo_galois_keygen(const pfo_context *c, const uint64_t *sk, uint32_t elt, uint64_t seed, uint64_t *key);
/* symmetric BFV encryption, output ct[2][L][n] coefficient form */
void pfo_encrypt_symmetric(const pfo_context *c, const uint64_t *sk, const uint64_t *plain, uint64_t seed,
                           uint64_t *ct);
/* decrypt coefficient-form ct[2][L][n] -> plain[n]; returns invariant noise budget in bits */
int pfo_decrypt(const pfo_context *c, const uint64_t *sk, const uint64_t *ct, uint64_t *plain);

/* ---- evaluator ops on the path (SEAL evaluator.cpp) ---- */
void pfo_ct_to_ntt(const pfo_context *c, uint64_t *ct, int size);   /* transform_to_ntt_inplace */
void pfo_ct_from_ntt(const pfo_context *c, uint64_t *ct, int size); /* transform_from_ntt_inplace */
void pfo_multiply_plain_ntt(const pfo_context *c, const uint64_t *ct, const uint64_t *pt_ntt, uint64_t *out);
void pfo_add(const pfo_context *c, uint64_t *a, const uint64_t *b); /* a += b, ct size 2 */
/* acc[2][L][n] = sum_{k<K} ct_k (.) pt_k with lazy 128-bit accumulation (ct stride 2*L*n, pt stride pt_stride words) */
void pfo_mac_plain_ntt(const pfo_context *c, const uint64_t *cts, const uint64_t *pts, size_t pt_stride, int K,
                       uint64_t *acc);
uint32_t pfo_galois_elt_from_step(const pfo_context *c, int step); /* util/galois.cpp */
void pfo_apply_galois(const pfo_context *c, const uint64_t *in, uint32_t elt, int limb, uint64_t *out);
void pfo_apply_galois_ntt(const pfo_context *c, const uint64_t *in, uint32_t elt, uint64_t *out);
void pfo_galois_ntt_table(const pfo_context *c, uint32_t elt, uint32_t *table);
/* Evaluator::switch_key_inplace: ct[2][L][n] (coefficient form) += keyswitch(target[L][n] coefficient form) */
void pfo_switch_key(const pfo_context *c, uint64_t *ct, const uint64_t *target, const uint64_t *key);
/* Evaluator::apply_galois_inplace on a coefficient-form ciphertext with the key of that element */
void pfo_apply_galois_ct(const pfo_context *c, uint64_t *ct, uint32_t elt, const uint64_t *key);
/* Evaluator::mod_switch_to_next_inplace (BFV, coefficient form): ct[2][Lin][n] -> out[2][Lin-1][n] */
void pfo_mod_switch_next(const pfo_context *c, const uint64_t *ct, int Lin, uint64_t *out);

/* ---- SEAL 4.1 wire format, compr_mode none (serialization.cpp, ciphertext.cpp, dynarray.h) ---- */
size_t pfo_ct_save_size(uint64_t n, int L, int size);
size_t pfo_ct_save(const uint64_t *ct, uint64_t n, int L, int size, int is_ntt, const uint64_t parms_id[4],
                   uint8_t *out);
/* returns bytes consumed, 0 on error */
size_t pfo_ct_load(const uint8_t *in, size_t len, uint64_t *n, int *L, int *size, int *is_ntt, uint64_t parms_id[4],
                   uint64_t *ct, size_t ct_cap_words);

/* ---- plaintext reference path ---- */
/* .fvecs/.ivecs reader, include/common/client_server_utils.h:24-56 ; returns 0 on success */
int pfo_vecs_read(const char *fname, size_t *d_out, size_t *n_out, void **data_out);
/* Stage 1, src/client/client_lib.cpp:50-81 (+ top-NPROBE slice :93-103).  Ties broken by index. */
void pfo_coarse_quantize(size_t nq, size_t d, size_t nlist, const float *x, const float *centroids, size_t nprobe,
                         int64_t *out_idx, float *out_dist);
/* exact squared L2, src/server/server_lib.cpp:151-164 arithmetic */
float pfo_l2sqr_ref(const float *base_vec, const float *query, size_t d);
/* Stage 2 plaintext semantics: server_lib.cpp:111-138 output layout with server_lib.cpp:140-167 distances.
 * lists are given CSR style (list_offsets[nlist+1], ids, vectors in list order).  Returns total written. */
size_t pfo_search_lists_plain(size_t nq, size_t d, const float *x, const int64_t *idx, size_t nprobe,
                              const int64_t *list_offsets, const int64_t *ids, const float *vectors, float *dist,
                              int64_t *labels, size_t cap, size_t *list_sizes);
/* PQ-ADC restatement of the FAISS fork's search_encrypted (see pf_oracle.c) and the matching encoder */
size_t pfo_search_lists_pq(size_t nq, size_t d, const float *x, const int64_t *idx, size_t nprobe, const float *centroids,
                           const int64_t *list_offsets, const int64_t *ids, size_t M, const float *pq_centroids,
                           const uint8_t *codes, float *dist, int64_t *labels, size_t cap, size_t *list_sizes);
void pfo_pq_encode_residuals(size_t n, size_t d, const float *vectors, const int64_t *list_of, const float *centroids, size_t M,
                             const float *pq_centroids, uint8_t *codes);
/* recall as the reference counts it (client_lib.cpp:272-281,325-328) and the standard definition */
void pfo_recall(size_t nq, size_t k_ret, const int64_t *returned, size_t gt_k, const int32_t *gt, double *ref_recall_1,
                double *ref_recall_10, double *ref_recall_100, double *std_recall_10, double *mrr_10);

/* ---- encrypted-distance layout ("generalised diagonal"; defined by this repo) ---- */
typedef struct {
    uint64_t n;     /* poly degree */
    uint32_t d;     /* true dimension */
    uint32_t d_pad; /* next pow2 >= d */
    uint32_t m;     /* query ciphertexts (dimension chunks) */
    uint32_t g;     /* partial-sum factor */
    uint32_t dc;    /* d_pad / m */
    uint32_t R;     /* dc / g rotations per chunk */
    uint32_t K;     /* m * R plaintext diagonals per block */
    uint32_t C;     /* candidates per block = n / g */
} pfo_layout;
int pfo_layout_init(pfo_layout *lay, uint64_t n, uint32_t d, uint32_t m, uint32_t g);
/* query chunk a -> slot vector (n values mod t), replicated in both rows */
void pfo_layout_query_slots(const pfo_layout *lay, uint64_t t, const int64_t *q, uint32_t a, uint64_t *slots);
/* slot owned by candidate u (0<=u<C) for partial sum j (0<=j<g) */
uint32_t pfo_layout_slot(const pfo_layout *lay, uint32_t u, uint32_t j);
/* diagonal (a, r) of a block of nvec<=C integer vectors xs[nvec][d]: slot values (-2 x) mod t */
void pfo_layout_diag_slots(const pfo_layout *lay, uint64_t t, const int32_t *xs, uint32_t nvec, uint32_t a,
                           uint32_t r, uint64_t *slots);
/* norm plaintext slots: ||x_u||^2 mod t in slot (u, j=0) */
void pfo_layout_norm_slots(const pfo_layout *lay, uint64_t t, const int32_t *xs, uint32_t nvec, uint64_t *slots);
/* encode one block: diag[K][L][n] NTT form, norm[L][n] = NTT(round(Q/t * norm_plain)) */
void pfo_encode_block(const pfo_context *c, const pfo_layout *lay, const int32_t *xs, uint32_t nvec, uint64_t *diag,
                      uint64_t *norm);
/* rotated query set: in cts[m][2][L][n] coefficient form; out rot[K][2][L][n] NTT form.
 * keys[R-1] : Galois key for step r (index r-1), or chain!=0: keys[0] is the step-1 key applied repeatedly. */
void pfo_rotate_query_set(const pfo_context *c, const pfo_layout *lay, const uint64_t *cts,
                          const uint64_t *const *keys, int chain, uint64_t *rot);
/* result of one block: out[2][L][n] coefficient form = INTT( sum_k rot_k (.) diag_k + norm on c0 ) */
void pfo_block_distance(const pfo_context *c, const pfo_layout *lay, const uint64_t *rot, const uint64_t *diag,
                        const uint64_t *norm, uint64_t *out);

/* ---- SEAL seeded ciphertexts (pf_oracle_seeded.c): BLAKE2b / BLAKE2Xb, Blake2xbPRNG, sample_poly_uniform ---- */
void pfo_blake2b_param(const uint8_t param[64], const uint8_t *key, size_t keylen, const uint8_t *msg, size_t msglen,
                       uint8_t *out, size_t outlen);
void pfo_blake2xb(uint8_t *out, size_t outlen, const uint8_t *in, size_t inlen, const uint8_t *key, size_t keylen);
void pfo_seal_sample_poly_uniform(const pfo_context *c, int L, const uint8_t seed[64], uint64_t *out);
void pfo_encrypt_symmetric_seeded(const pfo_context *c, const uint64_t *sk, const uint64_t *plain, uint64_t noise_seed,
                                  const uint8_t seed[64], uint64_t *ct);
size_t pfo_ct_save_seeded_size(uint64_t n, int L);
size_t pfo_ct_save_seeded(const uint64_t *c0, uint64_t n, int L, const uint64_t parms_id[4], const uint8_t seed[64],
                          uint8_t prng_type, uint8_t *out);

/* ---- whole-step drivers (pf_oracle_pipeline.c), OpenMP over queries / (query, block) pairs ---- */
int pfo_max_threads(void);
void pfo_encode_blocks(const pfo_context *c, const pfo_layout *lay, size_t nblocks, const int32_t *xs,
                       const int64_t *block_vec_offset, const uint32_t *nvec, uint64_t *diag, uint64_t *norm,
                       int nthreads);
void pfo_search_pairs(const pfo_context *c, const pfo_layout *lay, size_t nq, const uint64_t *cts,
                      const uint64_t *const *keys, int chain, size_t P, const int32_t *pair_query,
                      const int64_t *pair_block, const uint64_t *diag, const uint64_t *norm, uint64_t *rot,
                      uint64_t *out, int nthreads, double *times);
/* the same with every result mod-switched down to result_limbs (SEAL Evaluator::mod_switch_to_inplace)
 * before it is stored: out_ms[P][2][result_limbs][n] */
void pfo_search_pairs_ms(const pfo_context *c, const pfo_layout *lay, size_t nq, const uint64_t *cts,
                         const uint64_t *const *keys, int chain, size_t P, const int32_t *pair_query,
                         const int64_t *pair_block, const uint64_t *diag, const uint64_t *norm, uint64_t *rot,
                         int result_limbs, uint64_t *out_ms, int nthreads, double *times);

#ifdef __cplusplus
}
#endif
#endif
