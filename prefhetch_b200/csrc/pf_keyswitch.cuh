// pf_keyswitch.cuh — Galois automorphism + hybrid key switching with one special prime, restating
// SEAL 4.1 Evaluator::apply_galois_inplace / switch_key_inplace (BFV branch) so that the rotated
// ciphertext is bit-identical to SEAL's, delivered directly in NTT form for the MAC:
//   1. d[J][I]   = NTT_I( sigma(c1)_J mod q_I )            (pf_ntt.cuh, NTT_IN_GALOIS_REDUCE)
//   2. S_c[I]    = sum_J d[J][I] (.) key_J[c][I]            (ks_accumulate_kernel, lazy sums)
//      HOISTED (exactly the same d, without the L(L+1) NTTs per rotation):  with
//      D[J][I] = NTT_I(c1_J mod q_I) computed once per query ciphertext, neg = 0/1 polynomial of the
//      positions sigma negates, M[I] = NTT_I(neg):
//         d[J][I] = perm_sigma(D[J][I]) + (q_J mod q_I) * M[I]        whenever c1_J has no zero coefficient,
//      because sigma(c1)_J takes the unsigned residue q_J - x at negated positions and
//      (q_J - x) mod q_I = -(x mod q_I) + q_J (mod q_I).  Hence
//         S_c[I] = sum_J perm_sigma(D[J][I]) (.) key_J[c][I] + KM_c[I],   KM_c[I] = M[I] (.) sum_J (q_J mod q_I) key_J[c][I]
//      with KM precomputed per key.  A ciphertext whose c1 has a zero coefficient (probability ~N*L/q)
//      raises its flag and takes the exact path, so the result is SEAL's bit for bit in every case.
//   3. u_c       = INTT_P(S_c[P]);  W_c[j] = ((u_c + P/2) mod P mod q_j) - (P/2 mod q_j)
//   4. out_c[j]  = (S_c[j] - NTT_j(W_c[j])) * P^{-1}  (+ sigma_ntt(c0)[j] for c = 0)
// SEAL performs step 4 in coefficient form; doing it in NTT form is the same value because the
// NTT is linear and every operation is exact mod q_j.
#pragma once
#include "pf_common.cuh"
#include "pf_mac.cuh"
#include "pf_ntt.cuh"

struct KsParams {
    const RotJob *jobs;
    const DevModulus *mods;
    u64 *d;   // [z][L][L+1][N]
    u64 *S;   // [z][2][L+1][N]
    u64 *W;   // [z][2][L][N]
    int L, k, N;
    u64 p_half;
    u64 p_half_mod_q[PF_NTT_MAXMAP];
    u64 p_inv_mod_q[PF_NTT_MAXMAP], p_inv_mod_q_sh[PF_NTT_MAXMAP];
    int out_split; // write the rotated ciphertext in the MAC's split operand format
};

// S_c[I] for KS_QT consecutive jobs per CTA.  Jobs are ordered rotation-major, so consecutive jobs
// share the Galois key: its 2*L words per coefficient stay in registers across the KS_QT queries and
// the key stream (2.5 MiB per rotation at N=8192) is read once per KS_QT jobs instead of once per job.
// Keys, hoisted digits D and exact digits d are all stored in the split operand format of their limb.
// grid (N/512, L+1, ceil(z/KS_QT)); thread = 2 coefficients of output limb I, both key components
#define KS_QT 8
#define KS_MAXL 15
// FINISH = false: only the special-prime limb I = L (grid.y = 1); S_c[L] is written for the mod-down.
// FINISH = true : data limbs I < L (grid.y = L) with step 4 fused: W holds NTT_I(W_c[I]) and the
//                 rotated ciphertext (S_c[I] - W) * P^{-1} (+ sigma_ntt(c0) for c = 0) is written
//                 directly, so S of the data limbs never touches memory.
// FPRED: FP64-assisted reduction of the lazy sums (pf_mac.cuh lazy_reduce_fp; host enables it when
//        bits(q) + ceil(log2 L) <= 50).  W and the rotated ciphertext are touched once: L2 evict-first.
// LT <= 4: the key words of a thread live in its own shared-memory slots instead of 32 registers
//          (no barrier: a thread reads back only what it wrote), which brings the kernel to 64
//          registers and 4 CTAs/SM — it is bound by the latency of the digit gathers and W loads.
// The NTT-domain Galois permutation maps every aligned pair of positions {2u, 2u+1} onto an aligned pair
// (possibly swapped): bit 0 of a position is the top bit of its bit-reversed exponent index, and e * N = N
// (mod 2N) for odd e.  A thread's two coefficients are therefore ONE 16-byte gather (pair perm[2u] >> 1,
// swap flag perm[2u] & 1) instead of two 8-byte ones, for the hoisted digits and for sigma(c0).
// FPRED kernels finish on the FP64 pipe (idle here): the lazy sums are reduced to doubles congruent mod q
// (lazy_reduce_fp_d), the hoisting term, the mod-down correction W, the multiplication by P^{-1} (fp_mulmod,
// centred constant) and sigma(c0) are applied as exact integer-valued doubles and only the stored word is
// canonicalised: ~20 FP64 + ~8 integer instructions per output instead of ~50 integer ones (ncu r2: the
// kernel was issue-bound at 66 %, ALU 48 %, FP64 5 %).
template <int LT, bool FINISH, bool FPRED>
__global__ void __launch_bounds__(256, (LT <= 4 ? 4 : (LT <= 8 ? 2 : 1))) ks_accumulate_kernel(const KsParams p, int njobs) {
    constexpr bool SMEMKEY = LT <= 4;
    __shared__ ulonglong2 sk0[SMEMKEY ? LT : 1][SMEMKEY ? 256 : 1], sk1[SMEMKEY ? LT : 1][SMEMKEY ? 256 : 1];
    const int L = p.L, N = p.N;
    const int kk = p.k;
    const int I = FINISH ? (int)blockIdx.y : L;
    const int ki = (I == L) ? kk - 1 : I;
    const DevModulus m = p.mods[ki];
    const int c2 = blockIdx.x * 256 + threadIdx.x; // pair index
    const int sh = (int)m.split_shift;
    const double qinv = m.fqinv, fq = m.fq;
    const u64 once = l2_evict_first_policy();
    const int z0 = blockIdx.z * KS_QT, z1 = min(z0 + KS_QT, njobs);
    ulonglong2 k0[SMEMKEY ? 1 : LT], k1[SMEMKEY ? 1 : LT];
    const u64 *cur_key = nullptr;
    u32 pp = 0;       // pair this thread's two coefficients are gathered from: constant per key
    bool swp = false; // ... in swapped order
    for (int z = z0; z < z1; z++) {
        const RotJob job = p.jobs[z];
        if (job.key != cur_key) { // uniform across the CTA
            cur_key = job.key;
            const u32 p0 = __ldg(job.perm + 2 * c2);
            pp = p0 >> 1;
            swp = p0 & 1;
#pragma unroll
            for (int J = 0; J < LT; J++) {
                if (J < L) {
                    const u64 *kj = job.key + (size_t)J * 2 * kk * N;
                    const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2 *>(kj + (size_t)ki * N) + c2);
                    const ulonglong2 b = __ldg(reinterpret_cast<const ulonglong2 *>(kj + (size_t)(kk + ki) * N) + c2);
                    if (SMEMKEY) {
                        sk0[J][threadIdx.x] = a;
                        sk1[J][threadIdx.x] = b;
                    } else {
                        k0[J] = a;
                        k1[J] = b;
                    }
                }
            }
        }
        const bool hoisted = job.D && !*job.flag;
        LazyAcc a00, a01, a10, a11;
        lazy_zero(a00);
        lazy_zero(a01);
        lazy_zero(a10);
        lazy_zero(a11);
        const u64 *dz = p.d + (size_t)z * L * (L + 1) * N;
#pragma unroll
        for (int J = 0; J < LT; J++) {
            if (J < L) {
                ulonglong2 dv;
                if (hoisted) {
                    dv = __ldg(reinterpret_cast<const ulonglong2 *>(job.D + ((size_t)J * (L + 1) + I) * N) + pp);
                    if (swp) {
                        const u64 t = dv.x;
                        dv.x = dv.y;
                        dv.y = t;
                    }
                } else {
                    dv = reinterpret_cast<const ulonglong2 *>(dz + ((size_t)J * (L + 1) + I) * N)[c2];
                }
                const SplitOp dx = make_op(dv.x), dy = make_op(dv.y);
                const ulonglong2 ka = SMEMKEY ? sk0[J][threadIdx.x] : k0[J], kb = SMEMKEY ? sk1[J][threadIdx.x] : k1[J];
                const SplitOp k0x = make_op(ka.x), k0y = make_op(ka.y);
                const SplitOp k1x = make_op(kb.x), k1y = make_op(kb.y);
                lazy_mac(a00, dx.x0, dx.x1, dx.xs, k0x.x0, k0x.x1, k0x.xs);
                lazy_mac(a01, dy.x0, dy.x1, dy.xs, k0y.x0, k0y.x1, k0y.xs);
                lazy_mac(a10, dx.x0, dx.x1, dx.xs, k1x.x0, k1x.x1, k1x.xs);
                lazy_mac(a11, dy.x0, dy.x1, dy.xs, k1y.x0, k1y.x1, k1y.xs);
            }
        }
        u64 *Sz = p.S + (size_t)z * 2 * (L + 1) * N;
        if (FPRED) {
            // everything after the inner product as exact integer-valued doubles (|values| < 2^48)
            double f00 = lazy_reduce_fp_d(a00, sh, fq, qinv), f01 = lazy_reduce_fp_d(a01, sh, fq, qinv);
            double f10 = lazy_reduce_fp_d(a10, sh, fq, qinv), f11 = lazy_reduce_fp_d(a11, sh, fq, qinv);
            if (hoisted) {
                const ulonglong2 m0 = __ldg(reinterpret_cast<const ulonglong2 *>(job.KM + (size_t)I * N) + c2);
                const ulonglong2 m1 = __ldg(reinterpret_cast<const ulonglong2 *>(job.KM + (size_t)(L + 1 + I) * N) + c2);
                f00 = __dadd_rn(f00, fp_from_u64(m0.x));
                f01 = __dadd_rn(f01, fp_from_u64(m0.y));
                f10 = __dadd_rn(f10, fp_from_u64(m1.x));
                f11 = __dadd_rn(f11, fp_from_u64(m1.y));
            }
            if (!FINISH) {
                ulonglong2 r0, r1;
                r0.x = fp_canonical(f00, fq, qinv);
                r0.y = fp_canonical(f01, fq, qinv);
                r1.x = fp_canonical(f10, fq, qinv);
                r1.y = fp_canonical(f11, fq, qinv);
                reinterpret_cast<ulonglong2 *>(Sz + (size_t)I * N)[c2] = r0;
                reinterpret_cast<ulonglong2 *>(Sz + (size_t)(L + 1 + I) * N)[c2] = r1;
            } else {
                const u64 *Wz = p.W + (size_t)z * 2 * L * N;
                const ulonglong2 w0 = ldg_once(reinterpret_cast<const ulonglong2 *>(Wz + (size_t)I * N) + c2, once);
                const ulonglong2 w1 = ldg_once(reinterpret_cast<const ulonglong2 *>(Wz + (size_t)(L + I) * N) + c2, once);
                ulonglong2 c0 = __ldg(reinterpret_cast<const ulonglong2 *>(job.c0_ntt + (size_t)I * N) + pp);
                if (swp) {
                    const u64 t = c0.x;
                    c0.x = c0.y;
                    c0.y = t;
                }
                const double pinv = m.fpinv;
                const double g00 = __dadd_rn(fp_mulmod(__dadd_rn(f00, -fp_from_u64(w0.x)), pinv, fq, qinv), fp_from_u64(c0.x));
                const double g01 = __dadd_rn(fp_mulmod(__dadd_rn(f01, -fp_from_u64(w0.y)), pinv, fq, qinv), fp_from_u64(c0.y));
                const double g10 = fp_mulmod(__dadd_rn(f10, -fp_from_u64(w1.x)), pinv, fq, qinv);
                const double g11 = fp_mulmod(__dadd_rn(f11, -fp_from_u64(w1.y)), pinv, fq, qinv);
                ulonglong2 o0, o1;
                o0.x = fp_canonical(g00, fq, qinv);
                o0.y = fp_canonical(g01, fq, qinv);
                o1.x = fp_canonical(g10, fq, qinv);
                o1.y = fp_canonical(g11, fq, qinv);
                if (p.out_split) {
                    o0.x = split_word(o0.x, sh);
                    o0.y = split_word(o0.y, sh);
                    o1.x = split_word(o1.x, sh);
                    o1.y = split_word(o1.y, sh);
                }
                stg_once(reinterpret_cast<ulonglong2 *>(job.out + (size_t)I * N) + c2, o0, once);
                stg_once(reinterpret_cast<ulonglong2 *>(job.out + (size_t)(L + I) * N) + c2, o1, once);
            }
            continue;
        }
        ulonglong2 r0, r1;
        r0.x = lazy_reduce(a00, sh, m);
        r0.y = lazy_reduce(a01, sh, m);
        r1.x = lazy_reduce(a10, sh, m);
        r1.y = lazy_reduce(a11, sh, m);
        if (hoisted) {
            const ulonglong2 m0 = __ldg(reinterpret_cast<const ulonglong2 *>(job.KM + (size_t)I * N) + c2);
            const ulonglong2 m1 = __ldg(reinterpret_cast<const ulonglong2 *>(job.KM + (size_t)(L + 1 + I) * N) + c2);
            r0.x = addmod(r0.x, m0.x, m.q);
            r0.y = addmod(r0.y, m0.y, m.q);
            r1.x = addmod(r1.x, m1.x, m.q);
            r1.y = addmod(r1.y, m1.y, m.q);
        }
        if (!FINISH) {
            reinterpret_cast<ulonglong2 *>(Sz + (size_t)I * N)[c2] = r0;
            reinterpret_cast<ulonglong2 *>(Sz + (size_t)(L + 1 + I) * N)[c2] = r1;
        } else {
            const u64 *Wz = p.W + (size_t)z * 2 * L * N;
            const ulonglong2 w0 = ldg_once(reinterpret_cast<const ulonglong2 *>(Wz + (size_t)I * N) + c2, once);
            const ulonglong2 w1 = ldg_once(reinterpret_cast<const ulonglong2 *>(Wz + (size_t)(L + I) * N) + c2, once);
            ulonglong2 c0 = __ldg(reinterpret_cast<const ulonglong2 *>(job.c0_ntt + (size_t)I * N) + pp);
            if (swp) {
                const u64 t = c0.x;
                c0.x = c0.y;
                c0.y = t;
            }
            ulonglong2 o0, o1;
            o0.x = addmod(mul_shoup(submod(r0.x, w0.x, m.q), m.p_inv, m.p_inv_sh, m.q), c0.x, m.q);
            o0.y = addmod(mul_shoup(submod(r0.y, w0.y, m.q), m.p_inv, m.p_inv_sh, m.q), c0.y, m.q);
            o1.x = mul_shoup(submod(r1.x, w1.x, m.q), m.p_inv, m.p_inv_sh, m.q);
            o1.y = mul_shoup(submod(r1.y, w1.y, m.q), m.p_inv, m.p_inv_sh, m.q);
            if (p.out_split) {
                o0.x = split_word(o0.x, sh);
                o0.y = split_word(o0.y, sh);
                o1.x = split_word(o1.x, sh);
                o1.y = split_word(o1.y, sh);
            }
            stg_once(reinterpret_cast<ulonglong2 *>(job.out + (size_t)I * N) + c2, o0, once);
            stg_once(reinterpret_cast<ulonglong2 *>(job.out + (size_t)(L + I) * N) + c2, o1, once);
        }
    }
}

// neg[i] = 1 where coefficient position i of sigma(a) holds a negated coefficient.  grid (N/256)
__global__ void __launch_bounds__(256) galois_negmask_kernel(u64 *out, u32 einv, int N) {
    const u32 idx = blockIdx.x * 256 + threadIdx.x;
    const u32 i0 = (u32)(((u64)idx * einv) & (2u * N - 1));
    out[idx] = i0 >= (u32)N ? 1ull : 0ull;
}

// KM_c[I] = M[I] (.) sum_J (q_J mod q_I) key_J[c][I].  M[L+1][N] NTT form; key in split format.  grid (N/256, L+1, 2)
__global__ void __launch_bounds__(256) galois_km_kernel(const u64 *M, const u64 *key, u64 *KM, const DevModulus *mods,
                                                        int L, int k, int N) {
    const int i = blockIdx.x * 256 + threadIdx.x, I = blockIdx.y, c = blockIdx.z;
    const int ki = (I == L) ? k - 1 : I;
    const DevModulus m = mods[ki];
    u64 acc = 0;
    for (int J = 0; J < L; J++) {
        const u64 f = mods[J].q % m.q;
        const u64 kv = unsplit_word(key[(((size_t)J * 2 + c) * k + ki) * N + i], (int)m.split_shift);
        acc = addmod(acc, mulmod(f, kv, m), m.q);
    }
    KM[((size_t)c * (L + 1) + I) * N + i] = mulmod(acc, M[(size_t)I * N + i], m);
}

// after INTT_P of S_c[L] (in place): W_c[j] = ((u + P/2) mod P mod q_j) - (P/2 mod q_j).  grid (N/256, 2, z)
__global__ void __launch_bounds__(256) ks_moddown_prep_kernel(const KsParams p) {
    const int c = blockIdx.y, z = blockIdx.z, L = p.L, N = p.N;
    const int i = blockIdx.x * 256 + threadIdx.x;
    const DevModulus mp = p.mods[p.k - 1];
    const u64 u = p.S[((size_t)z * 2 + c) * (L + 1) * N + (size_t)L * N + i];
    const u64 v = barrett64(u + p.p_half, mp.q, mp.ratio1);
    for (int j = 0; j < L; j++) {
        const DevModulus mj = p.mods[j];
        const u64 w = submod(barrett64(v, mj.q, mj.ratio1), p.p_half_mod_q[j], mj.q);
        p.W[(((size_t)z * 2 + c) * L + j) * N + i] = w;
    }
}

// element-wise ciphertext add: limb of polynomial y is y % L.  grid (N/256, 2*L)
__global__ void __launch_bounds__(256) ct_add_kernel(const u64 *a, const u64 *b, u64 *out, const DevModulus *mods,
                                                     int L, int N) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const size_t o = (size_t)blockIdx.y * N + i;
    out[o] = addmod(a[o], b[o], mods[blockIdx.y % L].q);
}

// Galois key words [L][2][k][N] canonical -> split operand format of each limb.  grid (N/256, L*2*k)
__global__ void __launch_bounds__(256) key_split_kernel(u64 *key, const DevModulus *mods, int k, int N) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int limb = blockIdx.y % k;
    u64 *w = key + (size_t)blockIdx.y * N + i;
    *w = split_word(*w, (int)mods[limb].split_shift);
}

// SEAL Evaluator::mod_switch_to_inplace on coefficient-form BFV results: drop limbs L-1 .. Lr
// (RNSTool::divide_and_round_q_last_inplace per step).  tab[(c*16 + j)*3 + {0,1,2}] =
// {(q_c >> 1) mod q_j, q_c^{-1} mod q_j, its Shoup quotient} for j < c.  grid (N/256, 2, results)
#define MS_MAXL 16
__global__ void __launch_bounds__(256) modswitch_kernel(const u64 *in, size_t in_stride, u64 *out, size_t out_stride,
                                                        const DevModulus *mods, const u64 *tab, int L, int Lr, int N) {
    const int i = blockIdx.x * 256 + threadIdx.x, p = blockIdx.y;
    const u64 *src = in + (size_t)blockIdx.z * in_stride + (size_t)p * L * N + i;
    u64 x[MS_MAXL];
#pragma unroll
    for (int j = 0; j < MS_MAXL; j++)
        if (j < L) x[j] = src[(size_t)j * N];
    for (int c = L - 1; c >= Lr; c--) {
        u64 last = 0;
#pragma unroll
        for (int j = 0; j < MS_MAXL; j++)
            if (j == c) last = x[j];
        const u64 qc = mods[c].q;
        last = addmod(last, qc >> 1, qc);
#pragma unroll
        for (int j = 0; j < MS_MAXL; j++) {
            if (j < c) {
                const u64 qj = mods[j].q;
                const u64 *t = tab + ((size_t)c * MS_MAXL + j) * 3;
                const u64 tmp = submod(barrett64(last, qj, mods[j].ratio1), t[0], qj);
                x[j] = mul_shoup(submod(x[j], tmp, qj), t[1], t[2], qj);
            }
        }
    }
    u64 *dst = out + (size_t)blockIdx.z * out_stride + (size_t)p * Lr * N + i;
#pragma unroll
    for (int j = 0; j < MS_MAXL; j++)
        if (j < Lr) dst[(size_t)j * N] = x[j];
}

// Compile-time (L, Lr) version of modswitch_kernel, two coefficients per thread.  grid (N/512, 2, results)
template <int L, int LR>
__global__ void __launch_bounds__(256) modswitch_kernel_t(const u64 *in, size_t in_stride, u64 *out, size_t out_stride,
                                                          const DevModulus *mods, const u64 *tab, int N) {
    const int i2 = blockIdx.x * 256 + threadIdx.x, p = blockIdx.y;
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(in + (size_t)blockIdx.z * in_stride + (size_t)p * L * N) + i2;
    ulonglong2 x[L];
    u64 q[L], ratio[L];
#pragma unroll
    for (int j = 0; j < L; j++) {
        x[j] = src[(size_t)j * (N / 2)];
        q[j] = mods[j].q;
        ratio[j] = mods[j].ratio1;
    }
#pragma unroll
    for (int c = L - 1; c >= LR; c--) {
        const u64 half = q[c] >> 1;
        const u64 lx = addmod(x[c].x, half, q[c]), ly = addmod(x[c].y, half, q[c]);
#pragma unroll
        for (int j = 0; j < c; j++) {
            const u64 *t = tab + ((size_t)c * MS_MAXL + j) * 3;
            const u64 hm = t[0], inv = t[1], inv_sh = t[2];
            const u64 tx = submod(barrett64(lx, q[j], ratio[j]), hm, q[j]);
            const u64 ty = submod(barrett64(ly, q[j], ratio[j]), hm, q[j]);
            x[j].x = mul_shoup(submod(x[j].x, tx, q[j]), inv, inv_sh, q[j]);
            x[j].y = mul_shoup(submod(x[j].y, ty, q[j]), inv, inv_sh, q[j]);
        }
    }
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(out + (size_t)blockIdx.z * out_stride + (size_t)p * LR * N) + i2;
#pragma unroll
    for (int j = 0; j < LR; j++) dst[(size_t)j * (N / 2)] = x[j];
}
