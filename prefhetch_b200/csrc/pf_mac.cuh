// pf_mac.cuh — the bandwidth-bound ciphertext x plaintext multiply-accumulate:
//   acc[pair][2][L][N] = sum_{k<K} rot[query(pair)][k][2][L][N] (.) diag[block(pair)][k][L][N]  (+ norm on c0)
// SEAL semantics: Evaluator::multiply_plain (NTT form) + add_inplace chain + (pre-transformed)
// add_plain; every step is exact mod q, so one lazy accumulation with a single Barrett-128 at the
// end is bit-identical.
//
// Mapping (HBM roofline kernel; algorithmic bytes 8*L*N*(2*K*Q + (K+1)*B + 2*P)):
//  * a CTA owns (chunk of pairs of ONE query, limb l, slice of T coefficients).  The query's rotated
//    ciphertext slice rot[q][0..K)[0..1][l][slice] (K*2*T*8 B) is staged ONCE in shared memory and
//    reused for every block of the chunk, so the plaintext diagonals — read exactly once, with
//    128-bit L1-bypassing loads, a warp covering 512 contiguous bytes — are the only HBM stream.
//  * each thread owns 2 adjacent coefficients and BT blocks at a time; accumulators stay in
//    registers across the whole k loop.
//  * integer work is trimmed for primes <= 2^50 (all SEAL BFVDefault primes up to N=16384): the four
//    32x32 partial products are accumulated in three separate lazy sums (lo with carry count,
//    mid 64-bit, hi 64-bit) — 4 IMAD(.WIDE) + 1 carry add per 64x64 MAC, no 128-bit carry chains.
#pragma once
#include "pf_common.cuh"

struct MacChunk {
    int query;      // index into rot
    int pair_start; // first pair of the chunk
    int pair_count;
    int pad;
};

struct MacParams {
    const u64 *rot;  // [nq][K][2][L][N] NTT form
    const u64 *diag; // DB: block b, diagonal k, limb l, coefficient i at diag[b*diag_sb + k*diag_sk + l*N + i]
    const u64 *norm; // [nblocks][L][N] or nullptr
    long long diag_sb, diag_sk, norm_sb;
    const MacChunk *chunks;
    const long long *pair_block; // [P] block index of every pair
    u64 *out;                    // [P][2][L][N] NTT form
    const DevModulus *mods;      // limb l uses mods[l]
    int K, L, N;
};

struct LazyAcc {
    u32 l0, l1, c; // sum of lo*lo partial products, with carry count
    u64 mid;       // sum of lo*hi + hi*lo
    u64 hi;        // sum of hi*hi
};

__device__ __forceinline__ void lazy_zero(LazyAcc &a) {
    a.l0 = a.l1 = a.c = 0;
    a.mid = a.hi = 0;
}

__device__ __forceinline__ void lazy_mac(LazyAcc &a, u64 x, u64 y) {
    const u32 x0 = (u32)x, x1 = (u32)(x >> 32), y0 = (u32)y, y1 = (u32)(y >> 32);
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(a.l0), "+r"(a.l1), "+r"(a.c)
        : "r"(x0), "r"(y0));
    a.mid += (u64)x0 * y1;
    a.mid += (u64)x1 * y0;
    a.hi += (u64)x1 * y1;
}

__device__ __forceinline__ u64 lazy_reduce(const LazyAcc &a, const DevModulus &m) {
    const u64 lo64 = ((u64)a.l1 << 32) | a.l0;
    const u64 vlo = lo64 + (a.mid << 32);
    const u64 vhi = (u64)a.c + (a.mid >> 32) + a.hi + (vlo < lo64 ? 1ull : 0ull);
    return barrett128(vlo, vhi, m.q, m.ratio0, m.ratio1);
}

// generic accumulator for primes > 2^50: full 128-bit sum, reduced every `period` terms by the caller
struct WideAcc {
    u64 lo, hi;
};
__device__ __forceinline__ void wide_mac(WideAcc &a, u64 x, u64 y) {
    asm("mad.lo.cc.u64 %0, %2, %3, %0;\n\t"
        "madc.hi.u64 %1, %2, %3, %1;"
        : "+l"(a.lo), "+l"(a.hi)
        : "l"(x), "l"(y));
}

template <int T, int BT, int UNROLL, bool WIDE>
__global__ void __launch_bounds__(256) mac_kernel(const MacParams p) {
    constexpr int TX = T / 2;      // threads along the slice (2 coefficients each)
    constexpr int BY = 256 / TX;   // block lanes
    extern __shared__ __align__(16) u64 smem_ct[]; // [K][2][T]
    const MacChunk ch = p.chunks[blockIdx.x];
    const int slices = p.N / T;
    const int l = blockIdx.y / slices, s = blockIdx.y % slices;
    const int tx = threadIdx.x % TX, by = threadIdx.x / TX;
    const DevModulus m = p.mods[l];
    const size_t LN = (size_t)p.L * p.N;
    const size_t coef0 = (size_t)l * p.N + (size_t)s * T;

    // stage the query's rotated-ciphertext slice
    {
        const u64 *src = p.rot + (size_t)ch.query * p.K * 2 * LN + coef0;
        const int rows = p.K * 2;
        for (int i = threadIdx.x; i < rows * TX; i += 256) {
            const int row = i / TX, c = i % TX;
            const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(src + (size_t)row * LN) + c);
            reinterpret_cast<ulonglong2 *>(smem_ct)[(size_t)row * TX + c] = v;
        }
    }
    __syncthreads();

    const ulonglong2 *sct = reinterpret_cast<const ulonglong2 *>(smem_ct) + tx;
    for (int pi = by * BT; pi < ch.pair_count; pi += BY * BT) {
        const ulonglong2 *bp[BT];
        bool valid[BT];
#pragma unroll
        for (int j = 0; j < BT; j++) {
            valid[j] = (pi + j) < ch.pair_count;
            const long long b = p.pair_block[ch.pair_start + (valid[j] ? pi + j : pi)];
            bp[j] = reinterpret_cast<const ulonglong2 *>(p.diag + (size_t)b * p.diag_sb + coef0) + tx;
        }
        if (!WIDE) {
            LazyAcc acc[BT][2][2];
#pragma unroll
            for (int j = 0; j < BT; j++)
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    lazy_zero(acc[j][c][0]);
                    lazy_zero(acc[j][c][1]);
                }
            const size_t sk2 = (size_t)p.diag_sk / 2;
            for (int k0 = 0; k0 < p.K; k0 += UNROLL) {
                ulonglong2 pt[UNROLL][BT];
#pragma unroll
                for (int u = 0; u < UNROLL; u++)
#pragma unroll
                    for (int j = 0; j < BT; j++)
                        if (k0 + u < p.K) pt[u][j] = ldg_stream(bp[j] + (size_t)(k0 + u) * sk2);
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    if (k0 + u < p.K) {
                        const ulonglong2 c0 = sct[(size_t)((k0 + u) * 2 + 0) * TX];
                        const ulonglong2 c1 = sct[(size_t)((k0 + u) * 2 + 1) * TX];
#pragma unroll
                        for (int j = 0; j < BT; j++) {
                            lazy_mac(acc[j][0][0], c0.x, pt[u][j].x);
                            lazy_mac(acc[j][0][1], c0.y, pt[u][j].y);
                            lazy_mac(acc[j][1][0], c1.x, pt[u][j].x);
                            lazy_mac(acc[j][1][1], c1.y, pt[u][j].y);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < BT; j++) {
                if (!valid[j]) continue;
                const size_t pair = (size_t)ch.pair_start + pi + j;
                ulonglong2 r0, r1;
                r0.x = lazy_reduce(acc[j][0][0], m);
                r0.y = lazy_reduce(acc[j][0][1], m);
                r1.x = lazy_reduce(acc[j][1][0], m);
                r1.y = lazy_reduce(acc[j][1][1], m);
                if (p.norm) {
                    const long long b = p.pair_block[pair];
                    const ulonglong2 nv =
                        ldg_stream(reinterpret_cast<const ulonglong2 *>(p.norm + (size_t)b * p.norm_sb + coef0) + tx);
                    r0.x = addmod(r0.x, nv.x, m.q);
                    r0.y = addmod(r0.y, nv.y, m.q);
                }
                u64 *o = p.out + pair * 2 * LN + coef0;
                stg_stream(reinterpret_cast<ulonglong2 *>(o) + tx, r0);
                stg_stream(reinterpret_cast<ulonglong2 *>(o + LN) + tx, r1);
            }
        } else {
            // generic path: 128-bit accumulators, one Barrett-128 every 16 terms (16 * q^2 < 2^128 for q < 2^62)
            WideAcc acc[BT][2][2];
#pragma unroll
            for (int j = 0; j < BT; j++)
#pragma unroll
                for (int c = 0; c < 2; c++) acc[j][c][0].lo = acc[j][c][0].hi = acc[j][c][1].lo = acc[j][c][1].hi = 0;
            const size_t sk2 = (size_t)p.diag_sk / 2;
            for (int k = 0; k < p.K; k++) {
                const ulonglong2 c0 = sct[(size_t)(k * 2 + 0) * TX];
                const ulonglong2 c1 = sct[(size_t)(k * 2 + 1) * TX];
#pragma unroll
                for (int j = 0; j < BT; j++) {
                    const ulonglong2 pt = ldg_stream(bp[j] + (size_t)k * sk2);
                    wide_mac(acc[j][0][0], c0.x, pt.x);
                    wide_mac(acc[j][0][1], c0.y, pt.y);
                    wide_mac(acc[j][1][0], c1.x, pt.x);
                    wide_mac(acc[j][1][1], c1.y, pt.y);
                    if ((k & 15) == 15) {
#pragma unroll
                        for (int c = 0; c < 2; c++)
#pragma unroll
                            for (int h = 0; h < 2; h++) {
                                acc[j][c][h].lo = barrett128(acc[j][c][h].lo, acc[j][c][h].hi, m.q, m.ratio0, m.ratio1);
                                acc[j][c][h].hi = 0;
                            }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < BT; j++) {
                if (!valid[j]) continue;
                const size_t pair = (size_t)ch.pair_start + pi + j;
                ulonglong2 r0, r1;
                r0.x = barrett128(acc[j][0][0].lo, acc[j][0][0].hi, m.q, m.ratio0, m.ratio1);
                r0.y = barrett128(acc[j][0][1].lo, acc[j][0][1].hi, m.q, m.ratio0, m.ratio1);
                r1.x = barrett128(acc[j][1][0].lo, acc[j][1][0].hi, m.q, m.ratio0, m.ratio1);
                r1.y = barrett128(acc[j][1][1].lo, acc[j][1][1].hi, m.q, m.ratio0, m.ratio1);
                if (p.norm) {
                    const long long b = p.pair_block[pair];
                    const ulonglong2 nv =
                        ldg_stream(reinterpret_cast<const ulonglong2 *>(p.norm + (size_t)b * p.norm_sb + coef0) + tx);
                    r0.x = addmod(r0.x, nv.x, m.q);
                    r0.y = addmod(r0.y, nv.y, m.q);
                }
                u64 *o = p.out + pair * 2 * LN + coef0;
                stg_stream(reinterpret_cast<ulonglong2 *>(o) + tx, r0);
                stg_stream(reinterpret_cast<ulonglong2 *>(o + LN) + tx, r1);
            }
        }
    }
}
