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
//  * integer work is trimmed by the split operand format below: 3 IMAD.WIDE.U32 per 64x64 MAC.
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
    const long long *pair_block; // [P] block index of every pair (processing order)
    const int *pair_out;         // [P] result slot of every pair (response order)
    u64 *out;                    // result of pair i at out + i*out_stride: [2][L][N] NTT form
    long long out_stride;        // words between consecutive results (>= 2*L*N)
    const DevModulus *mods;      // limb l uses mods[l]
    int K, L, N;
    int query_base;              // MacChunk.query - query_base indexes rot
    int l2hint;                  // 1: results and the ciphertext slice are L2 evict-first; 2: + plaintexts evict-last
};

// ---- split-operand lazy accumulation -------------------------------------------------------------
// Operands of the MAC (rotated query ciphertexts and plaintext diagonals) are stored in HBM in
// "split" form  w = (x >> s) << 32 | (x & (2^s - 1)),  s = ceil(bits(q)/2)  (both halves < 2^25 for
// every SEAL BFVDefault prime up to N = 16384).  A 64x64 product then needs three 32x32->64
// multiply-adds (Karatsuba) into plain 64-bit sums that cannot overflow over K <= 128 terms:
//   lo += x0*y0;  hi += x1*y1;  kz += (x0+x1)*(y0+y1);      x*y = lo + (kz-lo-hi)*2^s + hi*2^2s
// = 3 IMAD.WIDE.U32 per MAC, no carries, accumulators in aligned register pairs.
struct LazyAcc {
    u64 lo, kz, hi;
};

__device__ __forceinline__ void lazy_zero(LazyAcc &a) { a.lo = a.kz = a.hi = 0; }

// acc += a*b (32x32 -> 64-bit accumulate).  Spelled as a carry chain: ptxas keeps it as ONE fused
// IMAD.WIDE.U32 acc, a, b, acc.  A plain mad.wide.u32 gets re-associated across consecutive k into
// 2 x IMAD.WIDE(+RZ) + IADD3 + IADD3.X, i.e. twice the issue slots (seen in SASS, profiles/).
__device__ __forceinline__ void mad_wide(u64 &acc, u32 a, u32 b) {
    asm("{\n\t"
        ".reg .u32 l0, l1;\n\t"
        "mov.b64 {l0, l1}, %0;\n\t"
        "mad.lo.cc.u32 l0, %1, %2, l0;\n\t"
        "madc.hi.u32 l1, %1, %2, l1;\n\t"
        "mov.b64 %0, {l0, l1};\n\t"
        "}"
        : "+l"(acc)
        : "r"(a), "r"(b));
}

// xs / ys = x0 + x1, y0 + y1 precomputed by the caller (shared by several MACs)
__device__ __forceinline__ void lazy_mac(LazyAcc &a, u32 x0, u32 x1, u32 xs, u32 y0, u32 y1, u32 ys) {
    mad_wide(a.lo, x0, y0);
    mad_wide(a.hi, x1, y1);
    mad_wide(a.kz, xs, ys);
}

__device__ __forceinline__ void add128(u64 &lo, u64 &hi, u64 alo, u64 ahi) {
    lo += alo;
    hi += ahi + (lo < alo ? 1ull : 0ull);
}

__device__ __forceinline__ u64 lazy_reduce(const LazyAcc &a, int s, const DevModulus &m) {
    const u64 mid = a.kz - a.lo - a.hi; // exact: kz >= lo + hi term by term
    u64 vlo = a.lo, vhi = 0;
    add128(vlo, vhi, mid << s, mid >> (64 - s));
    const int s2 = 2 * s;
    if (s2 < 64) add128(vlo, vhi, a.hi << s2, a.hi >> (64 - s2));
    else vhi += a.hi << (s2 - 64);
    return barrett128(vlo, vhi, m.q, m.ratio0, m.ratio1);
}

// FP64-assisted final reduction (B200 has full-rate FP64: 64 DFMA lanes/clk/SM, measured with
// tools/pipe_ubench.cu, and the pipe is otherwise idle here).  Valid when V/q <= 2^50, i.e.
// bits(q) + ceil(log2 K) <= 50 (every BFVDefault prime at N = 8192 with K <= 64): the quotient estimate
// floor(double(V) * (1/q)) is then within +-1 of the true quotient, and the remainder is finished in
// 64-bit integers with two corrections.  3 IMAD instead of ~26 per reduction: with K = 16 terms the
// Barrett-128 epilogue was 35 % of the kernel's IMAD work.
__device__ __forceinline__ double u52_to_double(u64 x) { // exact for x < 2^52
    return __longlong_as_double((long long)(x | 0x4330000000000000ull)) - 4503599627370496.0;
}
__device__ __forceinline__ u64 lazy_reduce_fp(const LazyAcc &a, int s, u64 q, double qinv) {
    const u64 mid = a.kz - a.lo - a.hi;
    const double p2s = __longlong_as_double((long long)(1023 + s) << 52);
    const double v = fma(u52_to_double(a.hi) * p2s, p2s, fma(u52_to_double(mid), p2s, u52_to_double(a.lo)));
    const u64 qh = (u64)__double2ll_rd(v * qinv);
    const u64 vlo = a.lo + (mid << s) + (a.hi << (2 * s)); // low 64 bits of V
    u64 r = vlo - qh * q;                                  // true remainder is in [-q, 2q)
    r = ((long long)r < 0) ? r + q : r;
    return r >= q ? r - q : r;
}
// The reduction done entirely on the FP64 pipe (the integer pipes are the busy ones in the MAC and in the
// key-switch inner product): V = lo + mid 2^s + hi 2^2s with lo, mid, hi < 2^52 exact in doubles.  A power-of-
// two scaling is exact, and x - rint(x/q) q is exact for any integer-valued x below 2^94 (the FMA forms
// k q exactly and the difference is an integer below q), so
//   V  ==  lo + reduce(mid 2^s) + reduce(hi 2^2s)   (mod q),   |result| < 2^52 + 2q
// in 11 FP64 operations and no integer multiply, shift or compare.  Returns that double (callers that go
// on in FP64 use it as it is; lazy_reduce_fp2 canonicalises it).  Valid when 2s + 2 + ceil(log2 K) <= 52.
__device__ __forceinline__ double fp_reduce_big(double x, double q, double qinv) { // -> about [-q/2, q/2], |x| < 2^94
    const double k = __dadd_rn(__fma_rn(x, qinv, 6755399441055744.0), -6755399441055744.0);
    return __fma_rn(-k, q, x);
}
__device__ __forceinline__ double lazy_reduce_fp_d(const LazyAcc &a, int s, double q, double qinv) {
    const u64 mid = a.kz - a.lo - a.hi;
    const double p2s = __longlong_as_double((long long)(1023 + s) << 52);
    const double p22s = __longlong_as_double((long long)(1023 + 2 * s) << 52);
    const double m = fp_reduce_big(__dmul_rn(u52_to_double(mid), p2s), q, qinv);
    const double h = fp_reduce_big(__dmul_rn(u52_to_double(a.hi), p22s), q, qinv);
    return __dadd_rn(__dadd_rn(u52_to_double(a.lo), m), h);
}
template <bool FPRED>
__device__ __forceinline__ u64 lazy_reduce_sel(const LazyAcc &a, int s, const DevModulus &m, double qinv) {
    return FPRED ? lazy_reduce_fp(a, s, m.q, qinv) : lazy_reduce(a, s, m);
}

__device__ __forceinline__ u64 split_word(u64 x, int s) { return ((x >> s) << 32) | (x & ((1ull << s) - 1)); }
__device__ __forceinline__ u64 unsplit_word(u64 w, int s) { return ((w >> 32) << s) | (w & 0xffffffffull); }

// generic accumulator for primes whose split sums could overflow: canonical operands, full 128-bit sum
struct WideAcc {
    u64 lo, hi;
};
__device__ __forceinline__ void wide_mac(WideAcc &a, u64 x, u64 y) {
    asm("mad.lo.cc.u64 %0, %2, %3, %0;\n\t"
        "madc.hi.u64 %1, %2, %3, %1;"
        : "+l"(a.lo), "+l"(a.hi)
        : "l"(x), "l"(y));
}

struct SplitOp {
    u32 x0, x1, xs;
};
__device__ __forceinline__ SplitOp make_op(u64 w) {
    SplitOp o;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(o.x0), "=r"(o.x1) : "l"(w));
    o.xs = o.x0 + o.x1;
    return o;
}

template <int BT, int UNROLL, int TX, bool FPRED>
__device__ __forceinline__ void mac_pairs_split(const MacParams &p, const ulonglong2 *sct, const DevModulus &m,
                                                int split, size_t coef0, size_t LN, int tx, size_t pair0,
                                                u64 pol_once, u64 pol_keep) {
    const double qinv = 1.0 / (double)m.q;
    const ulonglong2 *bp[BT];
#pragma unroll
    for (int j = 0; j < BT; j++) {
        const long long b = p.pair_block[pair0 + j];
        bp[j] = reinterpret_cast<const ulonglong2 *>(p.diag + (size_t)b * p.diag_sb + coef0) + tx;
    }
    LazyAcc acc[BT][2][2]; // [block][poly][coef]
#pragma unroll
    for (int j = 0; j < BT; j++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
            lazy_zero(acc[j][c][0]);
            lazy_zero(acc[j][c][1]);
        }
    const size_t sk2 = (size_t)p.diag_sk / 2;
    // software pipeline over k in groups of UNROLL: the loads of group i+1 are issued before the MACs
    // of group i, so every thread keeps UNROLL*BT..2*UNROLL*BT 128-bit loads in flight at all times.
    ulonglong2 pa[UNROLL][BT], pb[UNROLL][BT];
    auto load_group = [&](ulonglong2(&dst)[UNROLL][BT]) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
#pragma unroll
            for (int j = 0; j < BT; j++) {
                dst[u][j] = ldg_once(bp[j], pol_keep);
                bp[j] += sk2;
            }
    };
    auto mac_group = [&](const ulonglong2(&pt)[UNROLL][BT], int k0) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const ulonglong2 c0 = sct[(size_t)((k0 + u) * 2 + 0) * TX];
            const ulonglong2 c1 = sct[(size_t)((k0 + u) * 2 + 1) * TX];
            const SplitOp a00 = make_op(c0.x), a01 = make_op(c0.y), a10 = make_op(c1.x), a11 = make_op(c1.y);
#pragma unroll
            for (int j = 0; j < BT; j++) {
                const SplitOp b0 = make_op(pt[u][j].x), b1 = make_op(pt[u][j].y);
                lazy_mac(acc[j][0][0], a00.x0, a00.x1, a00.xs, b0.x0, b0.x1, b0.xs);
                lazy_mac(acc[j][0][1], a01.x0, a01.x1, a01.xs, b1.x0, b1.x1, b1.xs);
                lazy_mac(acc[j][1][0], a10.x0, a10.x1, a10.xs, b0.x0, b0.x1, b0.xs);
                lazy_mac(acc[j][1][1], a11.x0, a11.x1, a11.xs, b1.x0, b1.x1, b1.xs);
            }
        }
    };
    load_group(pa);
    int k0 = 0;
    for (; k0 + 2 * UNROLL <= p.K; k0 += 2 * UNROLL) {
        load_group(pb);
        mac_group(pa, k0);
        if (k0 + 2 * UNROLL < p.K) load_group(pa);
        mac_group(pb, k0 + UNROLL);
    }
    if (k0 < p.K) mac_group(pa, k0); // K == UNROLL (odd number of groups only happens for one group)
#pragma unroll
    for (int j = 0; j < BT; j++) {
        const size_t pair = pair0 + j;
        const size_t slot = (size_t)p.pair_out[pair];
        ulonglong2 r0, r1;
        r0.x = lazy_reduce_sel<FPRED>(acc[j][0][0], split, m, qinv);
        r0.y = lazy_reduce_sel<FPRED>(acc[j][0][1], split, m, qinv);
        r1.x = lazy_reduce_sel<FPRED>(acc[j][1][0], split, m, qinv);
        r1.y = lazy_reduce_sel<FPRED>(acc[j][1][1], split, m, qinv);
        if (p.norm) {
            const long long b = p.pair_block[pair];
            const ulonglong2 nv =
                ldg_stream(reinterpret_cast<const ulonglong2 *>(p.norm + (size_t)b * p.norm_sb + coef0) + tx);
            r0.x = addmod(r0.x, nv.x, m.q);
            r0.y = addmod(r0.y, nv.y, m.q);
        }
        u64 *o = p.out + slot * (size_t)p.out_stride + coef0;
        stg_once(reinterpret_cast<ulonglong2 *>(o) + tx, r0, pol_once);
        stg_once(reinterpret_cast<ulonglong2 *>(o + LN) + tx, r1, pol_once);
    }
}

template <int TX>
__device__ __forceinline__ void mac_pair_wide(const MacParams &p, const ulonglong2 *sct, const DevModulus &m,
                                              size_t coef0, size_t LN, int tx, size_t pair) {
    const long long b = p.pair_block[pair];
    const ulonglong2 *bp = reinterpret_cast<const ulonglong2 *>(p.diag + (size_t)b * p.diag_sb + coef0) + tx;
    WideAcc a00{0, 0}, a01{0, 0}, a10{0, 0}, a11{0, 0};
    const size_t sk2 = (size_t)p.diag_sk / 2;
    for (int k = 0; k < p.K; k++) {
        const ulonglong2 c0 = sct[(size_t)(k * 2 + 0) * TX];
        const ulonglong2 c1 = sct[(size_t)(k * 2 + 1) * TX];
        const ulonglong2 pt = ldg_stream(bp + (size_t)k * sk2);
        wide_mac(a00, c0.x, pt.x);
        wide_mac(a01, c0.y, pt.y);
        wide_mac(a10, c1.x, pt.x);
        wide_mac(a11, c1.y, pt.y);
        if ((k & 15) == 15) { // 16 * q^2 + q < 2^128 for q < 2^61
            a00.lo = barrett128(a00.lo, a00.hi, m.q, m.ratio0, m.ratio1);
            a01.lo = barrett128(a01.lo, a01.hi, m.q, m.ratio0, m.ratio1);
            a10.lo = barrett128(a10.lo, a10.hi, m.q, m.ratio0, m.ratio1);
            a11.lo = barrett128(a11.lo, a11.hi, m.q, m.ratio0, m.ratio1);
            a00.hi = a01.hi = a10.hi = a11.hi = 0;
        }
    }
    ulonglong2 r0, r1;
    r0.x = barrett128(a00.lo, a00.hi, m.q, m.ratio0, m.ratio1);
    r0.y = barrett128(a01.lo, a01.hi, m.q, m.ratio0, m.ratio1);
    r1.x = barrett128(a10.lo, a10.hi, m.q, m.ratio0, m.ratio1);
    r1.y = barrett128(a11.lo, a11.hi, m.q, m.ratio0, m.ratio1);
    if (p.norm) {
        const ulonglong2 nv =
            ldg_stream(reinterpret_cast<const ulonglong2 *>(p.norm + (size_t)b * p.norm_sb + coef0) + tx);
        r0.x = addmod(r0.x, nv.x, m.q);
        r0.y = addmod(r0.y, nv.y, m.q);
    }
    u64 *o = p.out + (size_t)p.pair_out[pair] * (size_t)p.out_stride + coef0;
    stg_stream(reinterpret_cast<ulonglong2 *>(o) + tx, r0);
    stg_stream(reinterpret_cast<ulonglong2 *>(o + LN) + tx, r1);
}

// K is a power of two and UNROLL divides it: K/UNROLL is 1 or even (host picks UNROLL = min(K, 2)).
template <int T, int UNROLL, bool WIDE, bool FPRED = false>
__global__ void __launch_bounds__(256, 2) mac_kernel(const MacParams p) {
    constexpr int TX = T / 2;    // threads along the slice (2 coefficients each)
    constexpr int BY = 256 / TX; // block lanes (warp-uniform: TX >= 32)
    extern __shared__ __align__(16) u64 smem_ct[]; // [K][2][T]
    const MacChunk ch = p.chunks[blockIdx.x];
    const int slices = p.N / T;
    const int l = blockIdx.y / slices, s = blockIdx.y % slices;
    const int tx = threadIdx.x % TX, by = threadIdx.x / TX;
    const DevModulus m = p.mods[l];
    const size_t LN = (size_t)p.L * p.N;
    const size_t coef0 = (size_t)l * p.N + (size_t)s * T;

    const u64 pol_once = p.l2hint ? l2_evict_first_policy() : l2_evict_normal_policy();
    const u64 pol_keep = p.l2hint == 2 ? l2_evict_last_policy() : l2_evict_normal_policy();
    // stage the query's rotated-ciphertext slice once; every block of the chunk reuses it
    {
        const u64 *src = p.rot + (size_t)(ch.query - p.query_base) * p.K * 2 * LN + coef0;
        const int rows = p.K * 2;
        for (int i = threadIdx.x; i < rows * TX; i += 256) {
            const int row = i / TX, c = i % TX;
            const ulonglong2 v = ldg_once(reinterpret_cast<const ulonglong2 *>(src + (size_t)row * LN) + c, pol_once);
            reinterpret_cast<ulonglong2 *>(smem_ct)[(size_t)row * TX + c] = v;
        }
    }
    __syncthreads();

    const ulonglong2 *sct = reinterpret_cast<const ulonglong2 *>(smem_ct) + tx;
    if (WIDE) {
        for (int pi = by; pi < ch.pair_count; pi += BY)
            mac_pair_wide<TX>(p, sct, m, coef0, LN, tx, (size_t)ch.pair_start + pi);
    } else {
        const int split = (int)m.split_shift;
        int pi = by * 2;
        for (; pi + 1 < ch.pair_count; pi += BY * 2)
            mac_pairs_split<2, UNROLL, TX, FPRED>(p, sct, m, split, coef0, LN, tx, (size_t)ch.pair_start + pi, pol_once, pol_keep);
        if (pi < ch.pair_count) // odd tail: one pair left for this lane
            mac_pairs_split<1, UNROLL, TX, FPRED>(p, sct, m, split, coef0, LN, tx, (size_t)ch.pair_start + pi, pol_once, pol_keep);
    }
}

// Occupancy variant (PF_MAC_VARIANT=4): one block per lane at a time (12 accumulators instead of 24) so
// that three CTAs fit an SM (<= 85 registers, 3 x 64 KiB of ciphertext slices): 24 warps instead of 16
// to cover the load latency, at the price of reading the ciphertext slice from shared memory once per
// block instead of once per two.  UNROLL4: four diagonals per load group.
template <int T, int UNROLL, bool FPRED, int MINCTA = 3>
__global__ void __launch_bounds__(256, MINCTA) mac_kernel_occ(const MacParams p) {
    constexpr int TX = T / 2, BY = 256 / TX;
    extern __shared__ __align__(16) u64 smem_ct[]; // [K][2][T]
    const MacChunk ch = p.chunks[blockIdx.x];
    const int slices = p.N / T;
    const int l = blockIdx.y / slices, s = blockIdx.y % slices;
    const int tx = threadIdx.x % TX, by = threadIdx.x / TX;
    const DevModulus m = p.mods[l];
    const size_t LN = (size_t)p.L * p.N;
    const size_t coef0 = (size_t)l * p.N + (size_t)s * T;
    const u64 pol = l2_evict_normal_policy();
    {
        const u64 *src = p.rot + (size_t)(ch.query - p.query_base) * p.K * 2 * LN + coef0;
        const int rows = p.K * 2;
        for (int i = threadIdx.x; i < rows * TX; i += 256) {
            const int row = i / TX, c = i % TX;
            const ulonglong2 v = ldg_once(reinterpret_cast<const ulonglong2 *>(src + (size_t)row * LN) + c, pol);
            reinterpret_cast<ulonglong2 *>(smem_ct)[(size_t)row * TX + c] = v;
        }
    }
    __syncthreads();
    const ulonglong2 *sct = reinterpret_cast<const ulonglong2 *>(smem_ct) + tx;
    const int split = (int)m.split_shift;
    // (A lane loop that issues the next pair's first loads and norm before the reduction of the current pair,
    // with the reduction moved to the FP64 pipe, was tried in round 2: 5 % fewer instructions but 0.92 ms
    // against 0.834 ms — at 64 registers the values held across the epilogue are paid for in local-memory
    // traffic and long-scoreboard stalls; profiles/README.md.)
    for (int pi = by; pi < ch.pair_count; pi += BY)
        mac_pairs_split<1, UNROLL, TX, FPRED>(p, sct, m, split, coef0, LN, tx, (size_t)ch.pair_start + pi, pol, pol);
}

// canonical <-> split conversion of polynomial arrays: chunk y (blockIdx.y) of `chunk_words` words is
// read at in + y*in_stride and written at out + y*out_stride; limb of word i of a chunk = (i / N) % L
__global__ void __launch_bounds__(256) split_convert_kernel(const u64 *in, u64 *out, size_t chunk_words,
                                                            size_t in_stride, size_t out_stride,
                                                            const DevModulus *mods, int L, int N, int to_split) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= chunk_words) return;
    const int s = (int)mods[(i / N) % L].split_shift;
    const u64 x = in[(size_t)blockIdx.y * in_stride + i];
    out[(size_t)blockIdx.y * out_stride + i] = to_split ? split_word(x, s) : unsplit_word(x, s);
}
