// pf_ntt_fp.cuh — the same negacyclic NTT / inverse NTT as pf_ntt.cuh (same passes, same SEAL
// ordering, same load modes, bit-identical canonical results), with the butterflies
// moved from the integer pipes to the FP64 pipe.
//
// Why: B200 (sm_100a) issues 64 DFMA lanes/clk/SM — the same rate as IMAD.WIDE — and the pipe is
// idle in an integer NTT (measured: tools/pipe_ubench.cu).  A Shoup/Harvey butterfly on 64-bit
// residues costs ~33 issued integer instructions (4 IMAD.WIDE for each 64x64 high product, carries,
// compares, selects); with residues held as exact integers in doubles it is 8 FP64 operations and
// no compares.  (B300/sm_103a would not have this option: its FP64 rate is vestigial.)
//
// Exactness (q < 2^50, every quantity an integer-valued double, no rounding anywhere that matters):
//   mulmod(a, w), |a| < 2^52, |w| <= q/2 (twiddles are stored centred):
//     h = a*w (rounded)            l = fma(a, w, -h)      -> a*w = h + l exactly (error-free product)
//     k = rint(h / q) via fma(h, 1/q, 1.5*2^52) - 1.5*2^52  (|h/q| < 2^51)
//     r = fma(-k, q, h)            exact: h, k*q are integers and |h - k*q| <= 0.75 q < 2^50
//     result = r + l               |l| <= ulp(h)/2 <= q/8  =>  |result| < q,  == a*w (mod q)
//   forward (Cooley-Tukey) butterflies add/subtract such values: magnitudes grow by < q per stage;
//   inverse (Gentleman-Sande) sums double per stage.  Values are pulled back to [-q/2, q/2]
//   (x - rint(x/q) q) at pass boundaries (inverse always, forward when N = 16384) and every second
//   inverse stage when N = 16384, which keeps everything below 2^52.  Eligibility (host): primes
//   <= 44 bits for N <= 8192, <= 49 bits for N = 16384 — all SEAL BFVDefault sets; other parameter
//   sets use the integer kernels.  A final reduce + one conditional +q stores the canonical residue,
//   identical to the integer kernels'.
#pragma once
#include "pf_keyswitch.cuh"
#include "pf_ntt.cuh"

template <int LOGN, int K, int P0, int U>
struct FpFwdStages {
    static __device__ __forceinline__ void run(double (&v)[32], const double *__restrict__ tw, double q, double qinv) {
        constexpr int LB = P0 - K + 1, G = 32 >> K, S0 = LOGN - 1 - P0, NT = 1 << (LOGN - 5), E = 1 << K;
        constexpr int half = 1 << (K - 1 - U);
        constexpr bool LANE = (LB == 0 && K == 5); // tw is the lane-major table: entry j of thread c at [j*NT + c]
#pragma unroll
        for (int g = 0; g < G; g++) {
            const int hi = (g * NT + (int)threadIdx.x) >> LB;
#pragma unroll
            for (int r = 0; r < (1 << U); r++) {
                const double w = LANE ? __ldg(tw + (((1 << U) - 1 + r) * NT + (int)threadIdx.x))
                                      : __ldg(tw + ((1 << (S0 + U)) + (hi << U) + r));
#pragma unroll
                for (int i = 0; i < half; i++) {
                    const int e = g * E + ((r << (K - U)) | i);
                    const double X = v[e];
                    const double T = fp_mulmod(v[e + half], w, q, qinv);
                    v[e] = __dadd_rn(X, T);
                    v[e + half] = __dadd_rn(X, -T);
                }
            }
        }
        if constexpr (U + 1 < K) FpFwdStages<LOGN, K, P0, U + 1>::run(v, tw, q, qinv);
    }
};

template <int LOGN, int K, int P0, class Load, class Store>
__device__ __forceinline__ void fp_fwd_pass(const double *__restrict__ tw, double q, double qinv, Load load,
                                            Store store) {
    constexpr int LB = P0 - K + 1, G = 32 >> K, NT = 1 << (LOGN - 5), E = 1 << K;
    double v[32];
#pragma unroll
    for (int g = 0; g < G; g++) {
        const int c = g * NT + (int)threadIdx.x, lo = c & ((1 << LB) - 1), hi = c >> LB;
        const int base = (hi << (P0 + 1)) | lo;
#pragma unroll
        for (int e = 0; e < E; e++) v[g * E + e] = load(base | (e << LB));
    }
    FpFwdStages<LOGN, K, P0, 0>::run(v, tw, q, qinv);
#pragma unroll
    for (int g = 0; g < G; g++) {
        const int c = g * NT + (int)threadIdx.x, lo = c & ((1 << LB) - 1), hi = c >> LB;
        const int base = (hi << (P0 + 1)) | lo;
#pragma unroll
        for (int e = 0; e < E; e++) store(base | (e << LB), v[g * E + e]);
    }
}

template <int LOGN, int K, int LB, bool LAST, int B>
struct FpInvStages {
    static __device__ __forceinline__ void run(double (&v)[32], const double *__restrict__ itw, double q, double qinv,
                                               double ninv, double last_w) {
        constexpr int G = 32 >> K, NT = 1 << (LOGN - 5), E = 1 << K;
        constexpr int half = 1 << B;
        constexpr int s = LOGN - 1 - LB - B;
        constexpr bool fold = LAST && (B == K - 1);
        constexpr bool LANE = (LB == 0 && K == 5); // itw is the lane-major table (see FpFwdStages)
        constexpr int lane0 = 32 - (32 >> B);      // first entry of stage B: 0, 16, 24, 28, 30
#pragma unroll
        for (int g = 0; g < G; g++) {
            const int hi = (g * NT + (int)threadIdx.x) >> LB;
#pragma unroll
            for (int r = 0; r < (1 << (K - 1 - B)); r++) {
                const double w = fold ? last_w
                                      : (LANE ? __ldg(itw + ((lane0 + r) * NT + (int)threadIdx.x))
                                              : __ldg(itw + ((1 << s) + (hi << (K - 1 - B)) + r)));
#pragma unroll
                for (int i = 0; i < half; i++) {
                    const int e = g * E + ((r << (B + 1)) | i);
                    const double U = v[e], V = v[e + half];
                    double sum = __dadd_rn(U, V);
                    // the sum path doubles every stage: with 49-bit primes pull it back every second stage
                    if (LOGN >= 14 && (B & 1) && !fold) sum = fp_reduce(sum, q, qinv);
                    if (fold) sum = fp_mulmod(sum, ninv, q, qinv);
                    v[e] = sum;
                    v[e + half] = fp_mulmod(__dadd_rn(U, -V), w, q, qinv);
                }
            }
        }
        if constexpr (B + 1 < K) FpInvStages<LOGN, K, LB, LAST, B + 1>::run(v, itw, q, qinv, ninv, last_w);
    }
};

template <int LOGN, int K, int LB, bool LAST, class Load, class Store>
__device__ __forceinline__ void fp_inv_pass(const double *__restrict__ itw, double q, double qinv, double ninv,
                                            double last_w, Load load, Store store) {
    constexpr int G = 32 >> K, NT = 1 << (LOGN - 5), E = 1 << K, P0 = LB + K - 1;
    double v[32];
#pragma unroll
    for (int g = 0; g < G; g++) {
        const int c = g * NT + (int)threadIdx.x, lo = c & ((1 << LB) - 1), hi = c >> LB;
        const int base = (hi << (P0 + 1)) | lo;
#pragma unroll
        for (int e = 0; e < E; e++) v[g * E + e] = load(base | (e << LB));
    }
    FpInvStages<LOGN, K, LB, LAST, 0>::run(v, itw, q, qinv, ninv, last_w);
#pragma unroll
    for (int g = 0; g < G; g++) {
        const int c = g * NT + (int)threadIdx.x, lo = c & ((1 << LB) - 1), hi = c >> LB;
        const int base = (hi << (P0 + 1)) | lo;
#pragma unroll
        for (int e = 0; e < E; e++) store(base | (e << LB), v[g * E + e]);
    }
}

// p.tw_fp_lane: the same twiddles of the pass over bits [4..0] in lane-major order: thread c's j-th twiddle
// at [j*NT + c].  In natural order a warp's 32 loads of one stage-U twiddle are 2^U doubles apart (up to
// 32 cache lines per instruction; ncu: 682 L1 wavefronts per warp per transform, half the kernel's LSU
// traffic); lane-major makes every load one 256-byte run (62 wavefronts).
// p.tw_fp: [nmod][2][N] doubles (centred twiddles); p.fp_consts: [nmod] {q, 1/q, centred N^-1, centred irp[1]*N^-1}
template <int LOGN, int INMODE>
__device__ __forceinline__ void ntt_fwd_fp_body(const NttParams &p, const unsigned bz) {
    using Cfg = NttCfg<LOGN>;
    extern __shared__ __align__(16) double smd[];
    const int mi = p.mod_map[blockIdx.x];
    const DevModulus m = p.mods[mi];
    const double *tw = p.tw_fp + (size_t)mi * 2 * Cfg::N;
    const double *twl = p.tw_fp_lane + (size_t)mi * 2 * Cfg::N;
    const u64 *in = (INMODE == NTT_IN_GALOIS_REDUCE)
                        ? p.jobs[bz].c1_coef + blockIdx.y * p.in_sy
                        : p.in + blockIdx.x * p.in_sx + blockIdx.y * p.in_sy + bz * p.in_sz;
    u64 *out = p.out + blockIdx.x * p.out_sx + blockIdx.y * p.out_sy + bz * p.out_sz;
    const double q = m.fq, qinv = m.fqinv;
    const u32 gal_einv = (INMODE == NTT_IN_GALOIS_REDUCE) ? p.jobs[bz].einv : 0u;
    constexpr bool MIDRED = LOGN >= 14; // keep magnitudes below 2^52 for 49-bit primes
    const double md_P = (INMODE == NTT_IN_MODDOWN) ? p.mods[p.md_pmod].fq : 0.0;
    bool saw_zero = false;

    auto gload = [&](int idx) -> double {
        if (INMODE == NTT_IN_PLAIN) return fp_from_u64(in[idx]);
        if (INMODE == NTT_IN_MODDOWN) {
            // W_c[j] = ((u + P/2) mod P mod q_j) - (P/2 mod q_j) is congruent mod q_j to the centred
            // representative of u mod P, and the transform takes any representative below 2^52: the
            // canonical output is the same, no Barrett reduction is needed on the way in
            const u64 u = in[idx];
            return u > p.ks_p_half ? __dadd_rn(fp_from_u64(u), -md_P) : fp_from_u64(u);
        }
        if (INMODE == NTT_IN_REDUCE) {
            const u64 x = in[idx];
            saw_zero |= (x == 0); // one flag store after the pass: a store per element serialises the loads
            return fp_reduce(fp_from_u64(x), q, qinv);
        }
        if (INMODE == NTT_IN_LIFT) {
            const u64 x = in[idx];
            return x >= p.lift_thr ? __dadd_rn(fp_from_u64(x), -(double)p.lift_t) : fp_from_u64(x); // centred value
        }
        const u32 i0 = (u32)(((u64)idx * gal_einv) & (2u * Cfg::N - 1));
        u64 x = in[i0 & (Cfg::N - 1)];
        if (i0 >= (u32)Cfg::N) {
            const u64 qs = p.mods[p.src_map[blockIdx.y]].q;
            x = x ? qs - x : 0;
        }
        return fp_reduce(fp_from_u64(x), q, qinv);
    };
    auto sload = [&](int idx) -> double {
        const double x = smd[sm_phys(idx)];
        return MIDRED ? fp_reduce(x, q, qinv) : x;
    };
    auto sstore = [&](int idx, double x) { smd[sm_phys(idx)] = x; };

    fp_fwd_pass<LOGN, Cfg::K1, LOGN - 1>(tw, q, qinv, gload, sstore);
    if (INMODE == NTT_IN_REDUCE && saw_zero && p.zero_flags) p.zero_flags[bz] = 1;
    __syncthreads();
    fp_fwd_pass<LOGN, Cfg::K2, 8>(tw, q, qinv, sload, sstore);
    __syncthreads();
    u64 *smu = reinterpret_cast<u64 *>(smd);
    auto fstore = [&](int idx, double x) {
        u64 r = fp_canonical(x, q, qinv);
        if (p.out_split) r = ((r >> m.split_shift) << 32) | (r & ((1ull << m.split_shift) - 1));
        smu[sm_phys(idx)] = r;
    };
    fp_fwd_pass<LOGN, Cfg::K3, 4>(twl, q, qinv, sload, fstore);
    __syncthreads();
#pragma unroll 8
    for (int i = 0; i < 32; i++) {
        const int idx = i * Cfg::NT + threadIdx.x;
        out[idx] = smu[sm_phys(idx)];
    }
}

// NTT_IN_GALOIS_REDUCE computes the exact per-rotation digits, needed only by jobs whose ciphertext has a
// zero coefficient in c1 (pf_keyswitch.cuh): CTA z looks at jobs [z*G, z*G+G) (G = p.job_group <= 32) and
// transforms the flagged ones, so the usual case costs one flag read per 32 jobs instead of one CTA per job.
template <int LOGN, int INMODE>
__global__ void __launch_bounds__(NttCfg<LOGN>::NT, (LOGN <= 13 ? 2 : 1)) ntt_fwd_fp_kernel(const NttParams p) {
    if (INMODE == NTT_IN_GALOIS_REDUCE) {
        __shared__ unsigned need;
        if (threadIdx.x < 32) {
            const unsigned j = blockIdx.z * p.job_group + threadIdx.x;
            bool n = false;
            if ((int)threadIdx.x < p.job_group && j < (unsigned)p.njobs) {
                const RotJob &jb = p.jobs[j];
                n = !(jb.D && !*jb.flag);
            }
            const unsigned mask = __ballot_sync(0xffffffffu, n);
            if (threadIdx.x == 0) need = mask;
        }
        __syncthreads();
        unsigned mask = need;
        while (mask) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1;
            ntt_fwd_fp_body<LOGN, INMODE>(p, blockIdx.z * p.job_group + b);
            __syncthreads();
        }
        return;
    }
    ntt_fwd_fp_body<LOGN, INMODE>(p, blockIdx.z);
}

template <int LOGN>
__global__ void __launch_bounds__(NttCfg<LOGN>::NT, (LOGN <= 13 ? 2 : 1)) ntt_inv_fp_kernel(const NttParams p) {
    using Cfg = NttCfg<LOGN>;
    extern __shared__ __align__(16) double smd[];
    const int mi = p.mod_map[blockIdx.x];
    const DevModulus m = p.mods[mi];
    const double *itw = p.tw_fp + (size_t)mi * 2 * Cfg::N + Cfg::N;
    const double *itwl = p.tw_fp_lane + (size_t)mi * 2 * Cfg::N + Cfg::N;
    const u64 *in = p.in + blockIdx.x * p.in_sx + blockIdx.y * p.in_sy + blockIdx.z * p.in_sz;
    u64 *out = p.out + blockIdx.x * p.out_sx + blockIdx.y * p.out_sy + blockIdx.z * p.out_sz;
    const double q = m.fq, qinv = m.fqinv;
    constexpr bool MIDRED = true; // Gentleman-Sande sums double per stage: reduce at every pass boundary
#pragma unroll 8
    for (int i = 0; i < 32; i++) {
        const int idx = i * Cfg::NT + threadIdx.x;
        smd[sm_phys(idx)] = fp_from_u64(in[idx]);
    }
    __syncthreads();
    auto sload0 = [&](int idx) -> double { return smd[sm_phys(idx)]; };
    auto sload = [&](int idx) -> double {
        const double x = smd[sm_phys(idx)];
        return MIDRED ? fp_reduce(x, q, qinv) : x;
    };
    auto sstore = [&](int idx, double x) { smd[sm_phys(idx)] = x; };
    auto gstore = [&](int idx, double x) {
        u64 r = fp_canonical(x, q, qinv);
        out[idx] = r;
    };
    fp_inv_pass<LOGN, Cfg::K3, 0, false>(itwl, q, qinv, m.fninv, m.flast_w, sload0, sstore);
    __syncthreads();
    fp_inv_pass<LOGN, Cfg::K2, 5, false>(itw, q, qinv, m.fninv, m.flast_w, sload, sstore);
    __syncthreads();
    fp_inv_pass<LOGN, Cfg::K1, 9, true>(itw, q, qinv, m.fninv, m.flast_w, sload, gstore);
}

// ---- result mod-switch on the FP64 pipe -------------------------------------------------------------
// Same arithmetic as modswitch_kernel_t (pf_keyswitch.cuh; SEAL divide_and_round_q_last per dropped limb),
// with every residue held as an exact integer in a double: dropping limb c,
//   last = (x_c + (q_c >> 1)) mod q_c  (canonical),   x_j <- (x_j - last + ((q_c >> 1) mod q_j)) * q_c^{-1}  mod q_j
// needs no reduction of `last` modulo q_j (|x_j - last + h| < 3 * 2^49 is fine for fp_mulmod) and no
// 64x64 high products: 8 FP64 operations per step instead of ~35 integer ones.  The integer kernel was
// IMAD-pipe bound (0.31 ms for the 1100 results of a step); this one is bound by its 0.7 GB of traffic.
// Valid for primes <= 49 bits (every BFVDefault set); tabf[(c*MS_MAXL + j)*2] = {(q_c>>1) mod q_j, centred q_c^{-1} mod q_j}.
// grid (N/512, 2, results), two coefficients per thread.
template <int L, int LR>
__global__ void __launch_bounds__(256) modswitch_fp_kernel_t(const u64 *in, size_t in_stride, u64 *out, size_t out_stride,
                                                             const DevModulus *mods, const double *tabf, int N) {
    const int i2 = blockIdx.x * 256 + threadIdx.x, p = blockIdx.y;
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(in + (size_t)blockIdx.z * in_stride + (size_t)p * L * N) + i2;
    double x0[L], x1[L], q[L], qinv[L];
#pragma unroll
    for (int j = 0; j < L; j++) {
        const ulonglong2 v = src[(size_t)j * (N / 2)];
        x0[j] = fp_from_u64(v.x);
        x1[j] = fp_from_u64(v.y);
        q[j] = mods[j].fq;
        qinv[j] = mods[j].fqinv;
    }
#pragma unroll
    for (int c = L - 1; c >= LR; c--) {
        const double half = (q[c] - 1.0) * 0.5; // q_c >> 1 (q_c is odd)
        double l0 = fp_reduce(__dadd_rn(x0[c], half), q[c], qinv[c]);
        double l1 = fp_reduce(__dadd_rn(x1[c], half), q[c], qinv[c]);
        l0 = l0 < 0.0 ? __dadd_rn(l0, q[c]) : l0;
        l1 = l1 < 0.0 ? __dadd_rn(l1, q[c]) : l1;
#pragma unroll
        for (int j = 0; j < c; j++) {
            const double hm = tabf[((size_t)c * MS_MAXL + j) * 2], w = tabf[((size_t)c * MS_MAXL + j) * 2 + 1];
            x0[j] = fp_mulmod(__dadd_rn(__dadd_rn(x0[j], -l0), hm), w, q[j], qinv[j]);
            x1[j] = fp_mulmod(__dadd_rn(__dadd_rn(x1[j], -l1), hm), w, q[j], qinv[j]);
        }
    }
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(out + (size_t)blockIdx.z * out_stride + (size_t)p * LR * N) + i2;
#pragma unroll
    for (int j = 0; j < LR; j++) {
        ulonglong2 r;
        r.x = fp_canonical(x0[j], q[j], qinv[j]);
        r.y = fp_canonical(x1[j], q[j], qinv[j]);
        dst[(size_t)j * (N / 2)] = r;
    }
}
