// pf_ntt.cuh — negacyclic NTT / inverse NTT, one CTA per N-point transform, whole polynomial
// resident in shared memory (N*8 B + padding: 66 KiB @8192, 132 KiB @16384), three register-radix
// passes (2^(LOGN-9), 16, 32 points per thread group), Harvey lazy butterflies with Shoup twiddles.
// Conventions are SEAL 4.1's (util/ntt.cpp, util/dwthandler.h): forward = Cooley-Tukey, natural in
// -> bit-reversed out, twiddle table index m+i at stage m holds psi^bitrev(m+i); inverse =
// Gentleman-Sande with N^{-1} folded into the last stage; stored results fully reduced to [0,q).
// Integer-pipe bound: ~30 integer issues per butterfly, (N/2)log2 N butterflies.
#pragma once
#include "pf_common.cuh"

#define PF_NTT_MAXMAP 17

// NTT_IN_MODDOWN (FP64 kernels only): the input is u = INTT_P(S_c[P]); limb j transforms
// W_c[j] = ((u + P/2) mod P mod q_j) - (P/2 mod q_j)  (pf_keyswitch.cuh step 3), entered as the centred
// representative of u mod P, which is congruent to it mod q_j
enum { NTT_IN_PLAIN = 0, NTT_IN_REDUCE = 1, NTT_IN_LIFT = 2, NTT_IN_GALOIS_REDUCE = 3, NTT_IN_MODDOWN = 4 };

struct NttParams {
    const u64 *in;
    u64 *out;
    long long in_sx, in_sy, in_sz;
    long long out_sx, out_sy, out_sz;
    const DevModulus *mods;
    const Twiddle *tw; // [nmod][2][N]
    const double *tw_fp; // [nmod][2][N] centred twiddles as doubles (FP64 NTT)
    const double *tw_fp_lane; // [nmod][2][N] lane-major twiddles of the pass over bits [4..0]
    int mod_map[PF_NTT_MAXMAP];  // blockIdx.x -> modulus index of the transform
    int src_map[PF_NTT_MAXMAP];  // blockIdx.y -> modulus index the input limb is reduced under (GALOIS_REDUCE)
    u64 lift_t, lift_thr;        // NTT_IN_LIFT: plain modulus and (t+1)/2
    const struct RotJob *jobs;   // NTT_IN_GALOIS_REDUCE: per blockIdx.z source polynomial and element
    int out_split;               // forward only: store results in the MAC's split operand format
    int *zero_flags;             // NTT_IN_REDUCE: zero_flags[blockIdx.z] = 1 if an input coefficient is 0
    u64 ks_p_half;               // NTT_IN_MODDOWN: floor(P/2)
    int md_pmod;                 // NTT_IN_MODDOWN: index of the special prime in mods
    int hoisted_jobs;            // NTT_IN_GALOIS_REDUCE: host hint, the jobs carry hoisted digits (RotJob.D)
    int njobs, job_group;        // NTT_IN_GALOIS_REDUCE (FP64 kernels): CTA z covers jobs [z*job_group, ...) < njobs
};

// one rotation = Evaluator::apply_galois_inplace on one ciphertext (see pf_keyswitch.cuh)
struct RotJob {
    const u64 *c1_coef; // [L][N] coefficient form of the input c1
    const u64 *c0_ntt;  // [L][N] NTT form of the input c0
    const u64 *key;     // [L][2][k][N] Galois key of the element, NTT form
    const u32 *perm;    // NTT-domain permutation table of the element
    u64 *out;           // [2][L][N] NTT form
    u32 einv;           // element^{-1} mod 2N
    u32 pad;
    // hoisted path (pf_keyswitch.cuh): digits of the un-rotated c1, shared by every rotation of it
    const u64 *D;       // [L][L+1][N]: NTT_I(c1_J mod q_I); nullptr -> exact per-rotation digits
    const u64 *KM;      // [2][L+1][N]: per-key correction term
    const int *flag;    // *flag != 0 -> c1 has a zero coefficient: use the exact per-rotation digits
};

__device__ __forceinline__ int sm_phys(int i) { return i + (i >> 5); }

template <int LOGN>
struct NttCfg {
    static constexpr int N = 1 << LOGN;
    static constexpr int NT = N / 32;
    static constexpr int K1 = LOGN - 9; // top bits [LOGN-1 .. 9]
    static constexpr int K2 = 4;        // bits [8 .. 5]
    static constexpr int K3 = 5;        // bits [4 .. 0]
    static constexpr size_t SMEM = (size_t)(N + N / 32) * sizeof(u64);
};

// One forward butterfly stage U of a register pass (compile-time U so every index is static).
template <int LOGN, int K, int P0, int U>
struct FwdStages {
    static __device__ __forceinline__ void run(u64 (&v)[32], const Twiddle *__restrict__ tw, u64 q, u64 two_q) {
        constexpr int LB = P0 - K + 1, G = 32 >> K, S0 = LOGN - 1 - P0, NT = 1 << (LOGN - 5), E = 1 << K;
        constexpr int half = 1 << (K - 1 - U);
#pragma unroll
        for (int g = 0; g < G; g++) {
            const int hi = (g * NT + (int)threadIdx.x) >> LB;
#pragma unroll
            for (int r = 0; r < (1 << U); r++) {
                const Twiddle w = __ldg(tw + ((1 << (S0 + U)) + (hi << U) + r));
#pragma unroll
                for (int i = 0; i < half; i++) {
                    const int e = g * E + ((r << (K - U)) | i);
                    u64 X = v[e], Y = v[e + half];
                    X = X >= two_q ? X - two_q : X;
                    const u64 T = mul_shoup_lazy(Y, w.x, w.y, q);
                    v[e] = X + T;
                    v[e + half] = X - T + two_q;
                }
            }
        }
        if constexpr (U + 1 < K) FwdStages<LOGN, K, P0, U + 1>::run(v, tw, q, two_q);
    }
};

// Forward register pass over element bits [P0 .. P0-K+1]; every thread owns 32 points = (32>>K) groups.
template <int LOGN, int K, int P0, class Load, class Store>
__device__ __forceinline__ void ntt_fwd_pass(const Twiddle *__restrict__ tw, u64 q, Load load, Store store) {
    constexpr int LB = P0 - K + 1, G = 32 >> K, NT = 1 << (LOGN - 5), E = 1 << K;
    const u64 two_q = q << 1;
    u64 v[32];
#pragma unroll
    for (int g = 0; g < G; g++) {
        const int c = g * NT + (int)threadIdx.x, lo = c & ((1 << LB) - 1), hi = c >> LB;
        const int base = (hi << (P0 + 1)) | lo;
#pragma unroll
        for (int e = 0; e < E; e++) v[g * E + e] = load(base | (e << LB));
    }
    FwdStages<LOGN, K, P0, 0>::run(v, tw, q, two_q);
#pragma unroll
    for (int g = 0; g < G; g++) {
        const int c = g * NT + (int)threadIdx.x, lo = c & ((1 << LB) - 1), hi = c >> LB;
        const int base = (hi << (P0 + 1)) | lo;
#pragma unroll
        for (int e = 0; e < E; e++) store(base | (e << LB), v[g * E + e]);
    }
}

// One inverse (Gentleman-Sande) stage B of a register pass; LAST folds N^{-1} into the top stage.
template <int LOGN, int K, int LB, bool LAST, int B>
struct InvStages {
    static __device__ __forceinline__ void run(u64 (&v)[32], const Twiddle *__restrict__ itw, const DevModulus &m) {
        constexpr int G = 32 >> K, NT = 1 << (LOGN - 5), E = 1 << K;
        constexpr int half = 1 << B;
        constexpr int s = LOGN - 1 - LB - B; // stage: m = 2^s
        constexpr bool fold = LAST && (B == K - 1);
        const u64 q = m.q, two_q = q << 1;
#pragma unroll
        for (int g = 0; g < G; g++) {
            const int hi = (g * NT + (int)threadIdx.x) >> LB;
#pragma unroll
            for (int r = 0; r < (1 << (K - 1 - B)); r++) {
                Twiddle w;
                if (fold) {
                    w.x = m.inv_last_w;
                    w.y = m.inv_last_w_sh;
                } else {
                    w = __ldg(itw + ((1 << s) + (hi << (K - 1 - B)) + r));
                }
#pragma unroll
                for (int i = 0; i < half; i++) {
                    const int e = g * E + ((r << (B + 1)) | i);
                    const u64 U = v[e], V = v[e + half];
                    u64 sum = U + V;
                    sum = sum >= two_q ? sum - two_q : sum;
                    if (fold) sum = mul_shoup_lazy(sum, m.n_inv, m.n_inv_sh, q);
                    v[e] = sum;
                    v[e + half] = mul_shoup_lazy(U - V + two_q, w.x, w.y, q);
                }
            }
        }
        if constexpr (B + 1 < K) InvStages<LOGN, K, LB, LAST, B + 1>::run(v, itw, m);
    }
};

// Inverse register pass over element bits [LB .. LB+K-1].
template <int LOGN, int K, int LB, bool LAST, class Load, class Store>
__device__ __forceinline__ void ntt_inv_pass(const Twiddle *__restrict__ itw, const DevModulus &m, Load load,
                                             Store store) {
    constexpr int G = 32 >> K, NT = 1 << (LOGN - 5), E = 1 << K, P0 = LB + K - 1;
    const u64 q = m.q;
    u64 v[32];
#pragma unroll
    for (int g = 0; g < G; g++) {
        const int c = g * NT + (int)threadIdx.x, lo = c & ((1 << LB) - 1), hi = c >> LB;
        const int base = (hi << (P0 + 1)) | lo;
#pragma unroll
        for (int e = 0; e < E; e++) v[g * E + e] = load(base | (e << LB));
    }
    InvStages<LOGN, K, LB, LAST, 0>::run(v, itw, m);
#pragma unroll
    for (int g = 0; g < G; g++) {
        const int c = g * NT + (int)threadIdx.x, lo = c & ((1 << LB) - 1), hi = c >> LB;
        const int base = (hi << (P0 + 1)) | lo;
#pragma unroll
        for (int e = 0; e < E; e++) {
            u64 x = v[g * E + e];
            if (LAST) x = x >= q ? x - q : x;
            store(base | (e << LB), x);
        }
    }
}

template <int LOGN, int INMODE>
__global__ void __launch_bounds__(NttCfg<LOGN>::NT, (LOGN <= 13 ? 2 : 1)) ntt_fwd_kernel(const NttParams p) {
    using Cfg = NttCfg<LOGN>;
    extern __shared__ __align__(16) u64 sm[];
    const int mi = p.mod_map[blockIdx.x];
    const DevModulus m = p.mods[mi];
    const Twiddle *tw = p.tw + (size_t)mi * 2 * Cfg::N;
    const u64 *in = (INMODE == NTT_IN_GALOIS_REDUCE)
                        ? p.jobs[blockIdx.z].c1_coef + blockIdx.y * p.in_sy
                        : p.in + blockIdx.x * p.in_sx + blockIdx.y * p.in_sy + blockIdx.z * p.in_sz;
    u64 *out = p.out + blockIdx.x * p.out_sx + blockIdx.y * p.out_sy + blockIdx.z * p.out_sz;
    const u64 q = m.q;
    const u32 gal_einv = (INMODE == NTT_IN_GALOIS_REDUCE) ? p.jobs[blockIdx.z].einv : 0u;
    if (INMODE == NTT_IN_GALOIS_REDUCE) { // exact digits are only needed where hoisting does not apply
        const RotJob &jb = p.jobs[blockIdx.z];
        if (jb.D && !*jb.flag) return;
    }

    auto gload = [&](int idx) -> u64 {
        if (INMODE == NTT_IN_PLAIN) return in[idx];
        if (INMODE == NTT_IN_REDUCE) {
            const u64 x = in[idx];
            if (x == 0 && p.zero_flags) p.zero_flags[blockIdx.z] = 1;
            return barrett64(x, q, m.ratio1);
        }
        if (INMODE == NTT_IN_LIFT) {
            u64 x = in[idx];
            return x >= p.lift_thr ? x + (q - p.lift_t) : x;
        }
        // NTT_IN_GALOIS_REDUCE: coefficient idx of sigma(a) = +-a[i], i = idx * e^{-1} mod 2N
        const u32 i0 = (u32)(((u64)idx * gal_einv) & (2u * Cfg::N - 1));
        u64 x = in[i0 & (Cfg::N - 1)];
        if (i0 >= (u32)Cfg::N) {
            const u64 qs = p.mods[p.src_map[blockIdx.y]].q;
            x = x ? qs - x : 0;
        }
        return barrett64(x, q, m.ratio1);
    };
    auto sload = [&](int idx) -> u64 { return sm[sm_phys(idx)]; };
    auto sstore = [&](int idx, u64 x) { sm[sm_phys(idx)] = x; };

    ntt_fwd_pass<LOGN, Cfg::K1, LOGN - 1>(tw, q, gload, sstore);
    __syncthreads();
    ntt_fwd_pass<LOGN, Cfg::K2, 8>(tw, q, sload, sstore);
    __syncthreads();
    auto fstore = [&](int idx, u64 x) {
        const u64 two_q = q << 1;
        x = x >= two_q ? x - two_q : x;
        x = x >= q ? x - q : x;
        if (p.out_split) x = ((x >> m.split_shift) << 32) | (x & ((1ull << m.split_shift) - 1));
        sm[sm_phys(idx)] = x;
    };
    ntt_fwd_pass<LOGN, Cfg::K3, 4>(tw, q, sload, fstore);
    __syncthreads();
#pragma unroll 8
    for (int i = 0; i < 32; i++) {
        const int idx = i * Cfg::NT + threadIdx.x;
        out[idx] = sm[sm_phys(idx)];
    }
}

template <int LOGN>
__global__ void __launch_bounds__(NttCfg<LOGN>::NT, (LOGN <= 13 ? 2 : 1)) ntt_inv_kernel(const NttParams p) {
    using Cfg = NttCfg<LOGN>;
    extern __shared__ __align__(16) u64 sm[];
    const int mi = p.mod_map[blockIdx.x];
    const DevModulus m = p.mods[mi];
    const Twiddle *itw = p.tw + (size_t)mi * 2 * Cfg::N + Cfg::N;
    const u64 *in = p.in + blockIdx.x * p.in_sx + blockIdx.y * p.in_sy + blockIdx.z * p.in_sz;
    u64 *out = p.out + blockIdx.x * p.out_sx + blockIdx.y * p.out_sy + blockIdx.z * p.out_sz;
#pragma unroll 8
    for (int i = 0; i < 32; i++) {
        const int idx = i * Cfg::NT + threadIdx.x;
        sm[sm_phys(idx)] = in[idx];
    }
    __syncthreads();
    auto sload = [&](int idx) -> u64 { return sm[sm_phys(idx)]; };
    auto sstore = [&](int idx, u64 x) { sm[sm_phys(idx)] = x; };
    auto gstore = [&](int idx, u64 x) { out[idx] = x; };
    ntt_inv_pass<LOGN, Cfg::K3, 0, false>(itw, m, sload, sstore);
    __syncthreads();
    ntt_inv_pass<LOGN, Cfg::K2, 5, false>(itw, m, sload, sstore);
    __syncthreads();
    ntt_inv_pass<LOGN, Cfg::K1, 9, true>(itw, m, sload, gstore);
}
