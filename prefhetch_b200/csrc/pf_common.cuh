// pf_common.cuh — shared device types and modular arithmetic for the sm_100a kernels.
// Semantics follow SEAL 4.1 util/uintarithsmallmod.h (Barrett 64/128, Shoup operand); results are
// always the canonical representative in [0,q) where they are stored.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;
typedef unsigned int u32;

struct DevModulus {
    u64 q;
    u64 ratio0, ratio1;  // floor(2^128/q) low / high words
    u64 n_inv, n_inv_sh; // N^{-1} mod q and its Shoup quotient
    u64 inv_last_w, inv_last_w_sh; // irp[1] * N^{-1} (last inverse stage with the scaling folded in)
    u64 split_shift; // s = ceil(bits(q)/2): operand split point of the MAC storage format (0 = canonical)
    // key-switch mod-down constants of a data limb (special prime P = last prime of the chain)
    u64 p_half_mod;        // (P >> 1) mod q
    u64 p_inv, p_inv_sh;   // P^{-1} mod q and its Shoup quotient
    u64 pad2;
    // FP64 NTT constants (pf_ntt_fp.cuh): q, 1/q, centred N^{-1}, centred irp[1]*N^{-1}
    double fq, fqinv, fninv, flast_w;
    double fpinv; // centred P^{-1} mod q (key-switch finish on the FP64 pipe, pf_keyswitch.cuh)
    double fpad;
};

// twiddle tables per modulus: fwd[N] then inv[N], each entry {w, floor(w*2^64/q)}
typedef ulonglong2 Twiddle; // .x = w, .y = Shoup quotient

__device__ __forceinline__ u64 barrett64(u64 x, u64 q, u64 ratio1) {
    u64 qh = __umul64hi(x, ratio1);
    u64 r = x - qh * q;
    return r >= q ? r - q : r;
}

// SEAL barrett_reduce_128
__device__ __forceinline__ u64 barrett128(u64 lo, u64 hi, u64 q, u64 ratio0, u64 ratio1) {
    u64 carry = __umul64hi(lo, ratio0);
    u64 t2lo = lo * ratio1, t2hi = __umul64hi(lo, ratio1);
    u64 tmp1 = t2lo + carry;
    u64 tmp3 = t2hi + (tmp1 < carry);
    u64 t3lo = hi * ratio0, t3hi = __umul64hi(hi, ratio0);
    u64 s = tmp1 + t3lo;
    carry = t3hi + (s < tmp1);
    tmp1 = hi * ratio1 + tmp3 + carry;
    u64 r = lo - tmp1 * q;
    return r >= q ? r - q : r;
}

__device__ __forceinline__ u64 mulmod(u64 a, u64 b, const DevModulus &m) {
    return barrett128(a * b, __umul64hi(a, b), m.q, m.ratio0, m.ratio1);
}

// x*w mod q with precomputed wsh = floor(w*2^64/q); lazy result in [0,2q) for any 64-bit x
__device__ __forceinline__ u64 mul_shoup_lazy(u64 x, u64 w, u64 wsh, u64 q) {
    u64 qh = __umul64hi(x, wsh);
    return x * w - qh * q;
}
__device__ __forceinline__ u64 mul_shoup(u64 x, u64 w, u64 wsh, u64 q) {
    u64 r = mul_shoup_lazy(x, w, wsh, q);
    return r >= q ? r - q : r;
}
__device__ __forceinline__ u64 addmod(u64 a, u64 b, u64 q) {
    u64 s = a + b;
    return s >= q ? s - q : s;
}
__device__ __forceinline__ u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }

// 128-bit streaming load that does not allocate in L1 (plaintext diagonals are read exactly once)
__device__ __forceinline__ ulonglong2 ldg_stream(const ulonglong2 *p) {
    ulonglong2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}
// L2 evict-first policy for data that is touched once per launch, so that re-used operands
// (Galois keys, hoisted digits) keep their L2 lines
__device__ __forceinline__ u64 l2_evict_first_policy() {
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ u64 l2_evict_normal_policy() {
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ u64 l2_evict_last_policy() {
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ ulonglong2 ldg_once(const ulonglong2 *p, u64 pol) {
    ulonglong2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;"
                 : "=l"(r.x), "=l"(r.y)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void stg_once(ulonglong2 *p, ulonglong2 v, u64 pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.u64 [%0], {%1, %2}, %3;" ::"l"(p), "l"(v.x), "l"(v.y), "l"(pol));
}
__device__ __forceinline__ void stg_stream(ulonglong2 *p, ulonglong2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y));
}

// ---- exact modular arithmetic on the FP64 pipe (residues as integer-valued doubles; derivation and
// bounds in pf_ntt_fp.cuh) ------------------------------------------------------------------------------
#define PF_FP_MAGIC 6755399441055744.0 /* 1.5 * 2^52 */

__device__ __forceinline__ double fp_mulmod(double a, double w, double q, double qinv) {
    const double h = __dmul_rn(a, w);
    const double l = __fma_rn(a, w, -h);
    const double k = __dadd_rn(__fma_rn(h, qinv, PF_FP_MAGIC), -PF_FP_MAGIC);
    const double r = __fma_rn(-k, q, h);
    return __dadd_rn(r, l);
}
__device__ __forceinline__ double fp_reduce(double x, double q, double qinv) { // -> [-q/2, q/2]
    const double k = __dadd_rn(__fma_rn(x, qinv, PF_FP_MAGIC), -PF_FP_MAGIC);
    return __fma_rn(-k, q, x);
}
__device__ __forceinline__ double fp_from_u64(u64 x) { // exact for x < 2^52
    return __dadd_rn(__longlong_as_double((long long)(x | 0x4330000000000000ull)), -4503599627370496.0);
}
__device__ __forceinline__ u64 fp_to_u64(double x) { // exact for integer x in [0, 2^51)
    return (u64)__double_as_longlong(__dadd_rn(x, 4503599627370496.0)) & 0x000fffffffffffffull;
}
__device__ __forceinline__ u64 fp_canonical(double x, double q, double qinv) {
    double r = fp_reduce(x, q, qinv);
    r = r < 0.0 ? __dadd_rn(r, q) : r;
    return fp_to_u64(r);
}

