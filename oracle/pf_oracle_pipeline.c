/*
 * pf_oracle_pipeline.c — the oracle's whole-step driver: the same SEAL-semantics op sequence the
 * CUDA engine runs (rotated query sets, ct x pt multiply-accumulate over every candidate block of
 * the probed lists, add_plain of the norms, inverse NTT), threaded with OpenMP so it can serve as
 * the CPU baseline on the GPU box's host cores.  TEST INFRASTRUCTURE ONLY (see pf_oracle.h).
 */
#include "pf_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int pfo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Encode nblocks blocks; xs_blocks[b] points at nvec[b] integer vectors.  diag[nblocks][K][L][n],
 * norm[nblocks][L][n]. */
void pfo_encode_blocks(const pfo_context *c, const pfo_layout *lay, size_t nblocks, const int32_t *xs,
                       const int64_t *block_vec_offset, const uint32_t *nvec, uint64_t *diag, uint64_t *norm,
                       int nthreads) {
    const size_t dw = (size_t)lay->K * c->L * c->n, nw = (size_t)c->L * c->n;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
    for (size_t b = 0; b < nblocks; b++)
        pfo_encode_block(c, lay, xs + (size_t)block_vec_offset[b] * lay->d, nvec[b], diag + b * dw, norm + b * nw);
}

/*
 * One step over a batch: nq queries (cts[nq][m][2][L][n], coefficient form) and P (query, block)
 * pairs.  rot is scratch [nq][K][2][L][n]; out[P][2][L][n] coefficient form.  times[0] = rotation
 * seconds, times[1] = MAC + INTT seconds.
 */
void pfo_search_pairs(const pfo_context *c, const pfo_layout *lay, size_t nq, const uint64_t *cts,
                      const uint64_t *const *keys, int chain, size_t P, const int32_t *pair_query,
                      const int64_t *pair_block, const uint64_t *diag, const uint64_t *norm, uint64_t *rot,
                      uint64_t *out, int nthreads, double *times) {
    const size_t ctw = (size_t)2 * c->L * c->n;
    const size_t dw = (size_t)lay->K * c->L * c->n, nw = (size_t)c->L * c->n;
    if (nthreads < 1) nthreads = 1;
    double t0 = now_s();
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
    for (size_t i = 0; i < nq; i++)
        pfo_rotate_query_set(c, lay, cts + i * lay->m * ctw, keys, chain, rot + i * lay->K * ctw);
    double t1 = now_s();
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
    for (size_t p = 0; p < P; p++)
        pfo_block_distance(c, lay, rot + (size_t)pair_query[p] * lay->K * ctw, diag + (size_t)pair_block[p] * dw,
                           norm + (size_t)pair_block[p] * nw, out + p * ctw);
    double t2 = now_s();
    if (times) {
        times[0] = t1 - t0;
        times[1] = t2 - t1;
    }
}

/*
 * The same step with SEAL Evaluator::mod_switch_to_inplace applied to every result before it is
 * stored (what the CUDA engine does for pf_params.result_limbs < L): out_ms[P][2][Lr][n].  scratch
 * per thread is allocated inside.  times as above (mod-switch time is part of times[1]).
 */
void pfo_search_pairs_ms(const pfo_context *c, const pfo_layout *lay, size_t nq, const uint64_t *cts,
                         const uint64_t *const *keys, int chain, size_t P, const int32_t *pair_query,
                         const int64_t *pair_block, const uint64_t *diag, const uint64_t *norm, uint64_t *rot,
                         int result_limbs, uint64_t *out_ms, int nthreads, double *times) {
    const size_t ctw = (size_t)2 * c->L * c->n;
    const size_t dw = (size_t)lay->K * c->L * c->n, nw = (size_t)c->L * c->n;
    const int Lr = (result_limbs < 1 || result_limbs > c->L) ? c->L : result_limbs;
    const size_t outw = (size_t)2 * Lr * c->n;
    if (nthreads < 1) nthreads = 1;
    double t0 = now_s();
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
    for (size_t i = 0; i < nq; i++)
        pfo_rotate_query_set(c, lay, cts + i * lay->m * ctw, keys, chain, rot + i * lay->K * ctw);
    double t1 = now_s();
#pragma omp parallel num_threads(nthreads)
    {
        uint64_t *a = (uint64_t *)malloc(ctw * sizeof(uint64_t)), *b = (uint64_t *)malloc(ctw * sizeof(uint64_t));
#pragma omp for schedule(dynamic)
        for (size_t p = 0; p < P; p++) {
            pfo_block_distance(c, lay, rot + (size_t)pair_query[p] * lay->K * ctw, diag + (size_t)pair_block[p] * dw,
                               norm + (size_t)pair_block[p] * nw, a);
            uint64_t *cur = a, *nxt = b;
            for (int Lin = c->L; Lin > Lr; Lin--) {
                pfo_mod_switch_next(c, cur, Lin, nxt);
                uint64_t *t = cur;
                cur = nxt;
                nxt = t;
            }
            memcpy(out_ms + p * outw, cur, outw * sizeof(uint64_t));
        }
        free(a);
        free(b);
    }
    double t2 = now_s();
    if (times) {
        times[0] = t1 - t0;
        times[1] = t2 - t1;
    }
}
