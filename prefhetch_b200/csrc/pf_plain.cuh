// pf_plain.cuh — plaintext stages: coarse quantization (query -> centroid distances, top-nprobe)
// and exact squared L2 over the probed lists.  The arithmetic restates the reference exactly:
// `float dist += std::pow(float - float, 2)` (ref: src/client/client_lib.cpp:59-62,
// src/server/server_lib.cpp:153-160) = float difference, exact double square, double add, round
// to float every step, in dimension order.  Ordering: ascending distance, ties by lower index
// (the reference's std::ranges::sort is unstable, ref: src/client/client_lib.cpp:70-75).
#pragma once
#include "pf_common.cuh"

__device__ __forceinline__ float ref_l2_step(float dist, float a, float b) {
    const float diff = __fsub_rn(a, b);
    const double p = __dmul_rn((double)diff, (double)diff);
    return __double2float_rn(__dadd_rn((double)dist, p));
}

// dist[q][j] = sum_k (x[q][k] - c[j][k])^2 in the reference arithmetic.  grid (ceil(nlist/128), nq)
__global__ void __launch_bounds__(128) coarse_dist_kernel(const float *__restrict__ x, const float *__restrict__ cent,
                                                          float *__restrict__ dist, int nlist, int d) {
    extern __shared__ float sq[]; // the query
    const int qi = blockIdx.y;
    for (int k = threadIdx.x; k < d; k += 128) sq[k] = x[(size_t)qi * d + k];
    __syncthreads();
    const int j = blockIdx.x * 128 + threadIdx.x;
    if (j >= nlist) return;
    const float *c = cent + (size_t)j * d;
    float acc = 0.0f;
    int k = 0;
    if ((d & 3) == 0) {
        for (; k < d; k += 4) {
            const float4 cv = __ldg(reinterpret_cast<const float4 *>(c + k));
            acc = ref_l2_step(acc, sq[k], cv.x);
            acc = ref_l2_step(acc, sq[k + 1], cv.y);
            acc = ref_l2_step(acc, sq[k + 2], cv.z);
            acc = ref_l2_step(acc, sq[k + 3], cv.w);
        }
    }
    for (; k < d; k++) acc = ref_l2_step(acc, sq[k], __ldg(c + k));
    dist[(size_t)qi * nlist + j] = acc;
}

__device__ __forceinline__ u64 warp_min_u64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const u64 other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other < v ? other : v;
    }
    return v;
}

// top-nprobe per query by repeated block-wide arg-min with warp-shuffle reductions.
// keys scratch [nq][nlist] u64 = (float bits << 32) | index; one CTA of 256 threads per query.
__global__ void __launch_bounds__(256) topk_select_kernel(const float *__restrict__ dist, u64 *__restrict__ keys,
                                                          long long *__restrict__ out_idx,
                                                          float *__restrict__ out_dist, int nlist, int nprobe) {
    __shared__ u64 warp_best[8];
    __shared__ u64 winner;
    const int qi = blockIdx.x, t = threadIdx.x;
    u64 *kq = keys + (size_t)qi * nlist;
    u64 local = ~0ull;
    for (int j = t; j < nlist; j += 256) {
        const u64 key = ((u64)__float_as_uint(dist[(size_t)qi * nlist + j]) << 32) | (u32)j;
        kq[j] = key;
        local = key < local ? key : local;
    }
    for (int r = 0; r < nprobe; r++) {
        const u64 wb = warp_min_u64(local);
        if ((t & 31) == 0) warp_best[t >> 5] = wb;
        __syncthreads();
        if (t < 32) {
            u64 v = t < 8 ? warp_best[t] : ~0ull;
            v = warp_min_u64(v);
            if (t == 0) {
                winner = v;
                out_idx[(size_t)qi * nprobe + r] = (long long)(u32)v;
                if (out_dist) out_dist[(size_t)qi * nprobe + r] = __uint_as_float((u32)(v >> 32));
            }
        }
        __syncthreads();
        const u64 w = winner;
        if (local == w) { // the owner retires its key and rescans its stripe
            kq[(u32)w] = ~0ull;
            local = ~0ull;
            for (int j = t; j < nlist; j += 256) {
                const u64 key = kq[j];
                local = key < local ? key : local;
            }
        }
        __syncthreads();
    }
}

// top-nprobe per query by radix select: 8 passes of an 8-bit histogram over the 64-bit keys find the
// nprobe-th smallest key exactly (keys are unique: the index is part of the key), the keys not above it
// are gathered (exactly nprobe of them) and sorted with a bitonic network in shared memory.  Same
// result and order as topk_select_kernel (ascending distance bits, ties by lower index) in
// O(8 + log^2 nprobe) block-wide steps instead of nprobe: 0.6 ms -> 0.02 ms at nlist 8192, nprobe 128.
// dynamic smem: M = next power of two >= nprobe keys.  One CTA of 256 threads per query.
__global__ void __launch_bounds__(256) topk_radix_kernel(const float *__restrict__ dist, u64 *__restrict__ keys,
                                                         long long *__restrict__ out_idx,
                                                         float *__restrict__ out_dist, int nlist, int nprobe, int M) {
    extern __shared__ u64 sel[]; // [M]
    __shared__ unsigned hist[256];
    __shared__ u64 s_prefix;
    __shared__ unsigned s_k, s_count;
    const int qi = blockIdx.x, t = threadIdx.x;
    u64 *kq = keys + (size_t)qi * nlist;
    for (int j = t; j < nlist; j += 256) kq[j] = ((u64)__float_as_uint(dist[(size_t)qi * nlist + j]) << 32) | (u32)j;
    if (t == 0) {
        s_prefix = 0;
        s_k = (unsigned)nprobe;
        s_count = 0;
    }
    __syncthreads();
    u64 mask = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        hist[t] = 0;
        __syncthreads();
        const u64 prefix = s_prefix;
        for (int j = t; j < nlist; j += 256) {
            const u64 key = kq[j];
            if ((key & mask) == prefix) atomicAdd(&hist[(unsigned)(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (t < 32) { // warp 0: lane l owns bins 8l..8l+7; find the bin holding the s_k-th key of this prefix
            unsigned c[8], sum = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                c[i] = hist[t * 8 + i];
                sum += c[i];
            }
            unsigned incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                if (t >= o) incl += v;
            }
            const unsigned k = s_k;
            const unsigned owner = __ffs(__ballot_sync(0xffffffffu, incl >= k)) - 1;
            if ((unsigned)t == owner) {
                unsigned below = incl - sum;
                int b = 0;
                while (below + c[b] < k) below += c[b++];
                s_k = k - below;
                s_prefix = prefix | ((u64)(t * 8 + b) << shift);
            }
        }
        mask |= (u64)255 << shift;
        __syncthreads();
    }
    const u64 kth = s_prefix; // the nprobe-th smallest key
    for (int j = t; j < M; j += 256) sel[j] = ~0ull;
    __syncthreads();
    for (int j = t; j < nlist; j += 256) {
        const u64 key = kq[j];
        if (key <= kth) sel[atomicAdd(&s_count, 1u)] = key;
    }
    __syncthreads();
    for (int size = 2; size <= M; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = t; i < M / 2; i += 256) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const u64 a = sel[lo], b = sel[hi];
                if ((a > b) == up) {
                    sel[lo] = b;
                    sel[hi] = a;
                }
            }
            __syncthreads();
        }
    for (int r = t; r < nprobe; r += 256) {
        const u64 v = sel[r];
        out_idx[(size_t)qi * nprobe + r] = (long long)(u32)v;
        if (out_dist) out_dist[(size_t)qi * nprobe + r] = __uint_as_float((u32)(v >> 32));
    }
}

struct ListJob {
    long long vec_begin; // first vector of the list in the list-ordered base
    long long out_begin; // where its results start in dist/labels
    int count;
    int query;
};

// exact squared L2 of every vector of every probed list; one CTA per (query, list) job
__global__ void __launch_bounds__(128) list_l2_kernel(const float *__restrict__ x, const float *__restrict__ base,
                                                      const long long *__restrict__ ids, const ListJob *jobs,
                                                      float *__restrict__ dist, long long *__restrict__ labels,
                                                      int d, unsigned long long cap) {
    extern __shared__ float sq[];
    const ListJob job = jobs[blockIdx.x];
    for (int k = threadIdx.x; k < d; k += 128) sq[k] = x[(size_t)job.query * d + k];
    __syncthreads();
    for (int v = threadIdx.x; v < job.count; v += 128) {
        const float *b = base + (size_t)(job.vec_begin + v) * d;
        float acc = 0.0f;
        int k = 0;
        if ((d & 3) == 0) {
            for (; k < d; k += 4) {
                const float4 bv = __ldg(reinterpret_cast<const float4 *>(b + k));
                acc = ref_l2_step(acc, bv.x, sq[k]);
                acc = ref_l2_step(acc, bv.y, sq[k + 1]);
                acc = ref_l2_step(acc, bv.z, sq[k + 2]);
                acc = ref_l2_step(acc, bv.w, sq[k + 3]);
            }
        }
        for (; k < d; k++) acc = ref_l2_step(acc, __ldg(b + k), sq[k]);
        const unsigned long long o = (unsigned long long)(job.out_begin + v);
        if (o < cap) {
            dist[o] = acc;
            labels[o] = ids[job.vec_begin + v];
        }
    }
}

// Product-quantizer ADC over every code of every probed list (what the reference's FAISS fork computes in
// search_encrypted today; restated by the CPU checker (pfo_search_lists_pq), same float operations in the same
// order: r = x - centroid, tab[m][j] = sum_k (r - pq)^2, dis = sum_m tab[m][code[m]]; multiply and add are separate
// roundings, as in FAISS's reference fvec_L2sqr).  One CTA per (query, list) job: the residual and the M x 256
// distance table live in shared memory (M = 32: 32 KiB), then every thread scores codes of the list — M table
// look-ups per code, the code bytes read 16 at a time when M allows.
__global__ void __launch_bounds__(256) list_pq_adc_kernel(const float *__restrict__ x, const float *__restrict__ cent,
                                                          const float *__restrict__ pqc, const unsigned char *__restrict__ codes,
                                                          const long long *__restrict__ ids, const ListJob *jobs,
                                                          const int *__restrict__ job_list, float *__restrict__ dist,
                                                          long long *__restrict__ labels, int d, int M,
                                                          unsigned long long cap) {
    extern __shared__ float pq_sm[]; // [d] residual, then [M][256] table
    float *r = pq_sm, *tab = pq_sm + d;
    const ListJob job = jobs[blockIdx.x];
    const int l = job_list[blockIdx.x];
    const int dsub = d / M;
    for (int k = threadIdx.x; k < d; k += 256) r[k] = __fsub_rn(x[(size_t)job.query * d + k], __ldg(cent + (size_t)l * d + k));
    __syncthreads();
    for (int e = threadIdx.x; e < M * 256; e += 256) { // e = m * 256 + j: consecutive threads read consecutive sub-centroids
        const int m = e >> 8;
        const float *c = pqc + (size_t)e * dsub;
        const float *rm = r + m * dsub;
        float acc = 0.0f;
        for (int k = 0; k < dsub; k++) {
            const float t = __fsub_rn(rm[k], __ldg(c + k));
            acc = __fadd_rn(acc, __fmul_rn(t, t));
        }
        tab[e] = acc;
    }
    __syncthreads();
    for (int v = threadIdx.x; v < job.count; v += 256) {
        const unsigned char *code = codes + (size_t)(job.vec_begin + v) * M;
        float acc = 0.0f;
        int m = 0;
        if ((M & 15) == 0) {
            for (; m < M; m += 16) {
                const uint4 w = __ldg(reinterpret_cast<const uint4 *>(code + m));
                const unsigned ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int b = 0; b < 16; b++) acc = __fadd_rn(acc, tab[((m + b) << 8) + ((ws[b >> 2] >> (8 * (b & 3))) & 255u)]);
            }
        }
        for (; m < M; m++) acc = __fadd_rn(acc, tab[(m << 8) + code[m]]);
        const unsigned long long o = (unsigned long long)(job.out_begin + v);
        if (o < cap) {
            dist[o] = acc;
            labels[o] = ids[job.vec_begin + v];
        }
    }
}

// Server::preciseSearch: ids are base row numbers; pos_of_id maps them to list-ordered positions
__global__ void __launch_bounds__(128) precise_l2_kernel(const float *__restrict__ x, const float *__restrict__ base,
                                                         const long long *__restrict__ pos_of_id,
                                                         const long long *__restrict__ ids, float *__restrict__ out,
                                                         int d, int nids, long long ntotal) {
    extern __shared__ float sq[];
    const int qi = blockIdx.y;
    for (int k = threadIdx.x; k < d; k += 128) sq[k] = x[(size_t)qi * d + k];
    __syncthreads();
    const int j = blockIdx.x * 128 + threadIdx.x;
    if (j >= nids) return;
    const long long id = ids[(size_t)qi * nids + j];
    if (id < 0 || id >= ntotal) {
        out[(size_t)qi * nids + j] = __int_as_float(0x7fc00000);
        return;
    }
    const float *b = base + (size_t)pos_of_id[id] * d;
    float acc = 0.0f;
    for (int k = 0; k < d; k++) acc = ref_l2_step(acc, __ldg(b + k), sq[k]);
    out[(size_t)qi * nids + j] = acc;
}

// stream-ordered flags for the multi-GPU result gather (see include/prefhetch_b200.h)
__global__ void flag_write_kernel(volatile unsigned *flag, unsigned value) {
    __threadfence_system();
    *flag = value;
    __threadfence_system();
}
// Bounded wait: a peer that died, hit an error or lost protocol order must not hang this stream (and
// every later implicit device synchronisation of the process) for ever.  After timeout_ns the kernel
// gives up, records {flag value seen, value wanted} in the engine's host-mapped error word and ends;
// the next API call on the engine returns PF_ERR_CUDA (pf_engine.cu check_device_error).
__global__ void flag_wait_kernel(const volatile unsigned *flag, unsigned value, unsigned long long timeout_ns,
                                 volatile unsigned long long *err_word) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned seen = *flag;
    while (seen < value) {
        __nanosleep(500);
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {
            *err_word = 0x8000000000000000ull | ((unsigned long long)seen << 32) | value;
            __threadfence_system();
            return;
        }
        seen = *flag;
    }
    __threadfence_system();
}

// 64-bit additive checksum of `n` words (gather verification: a rank's results against the copy that
// landed in rank 0's buffer).  grid (blocks), 256 threads; out must be zeroed.
__global__ void __launch_bounds__(256) checksum_kernel(const u64 *__restrict__ p, size_t n, u64 *out) {
    u64 acc = 0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        acc += p[i] * (2 * (u64)(i & 0xffff) + 1); // position-weighted: a permuted copy does not pass
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// The 113-byte SEAL stream header (SEALHeader + Ciphertext members + DynArray header, compr_mode none)
// is the same for every result of a call: it is written in front of the 128-byte aligned ciphertext
// words on the device, so the response leaves the GPU complete with one copy per query group instead of
// the host writing ~1100 headers after the last copy has landed.
struct ResultHeader {
    unsigned char b[128];
};
__global__ void __launch_bounds__(128) stamp_headers_kernel(unsigned char *blob, size_t slot, size_t pad, size_t nresults,
                                                            const ResultHeader hd) {
    const size_t r = (size_t)blockIdx.x * 128 + threadIdx.x;
    if (r >= nresults) return;
    unsigned char *dst = blob + r * slot + pad;
    for (int i = 0; i < 113; i++) dst[i] = hd.b[i];
    for (size_t i = 0; i < pad; i++) blob[r * slot + i] = 0; // the alignment pad travels with the response: keep it deterministic
}

// Query ciphertexts arrive as SEAL streams: 113 header bytes, then 2*L*N little-endian words — at a byte offset
// that is not a multiple of 8.  The blob is uploaded as it is and this kernel moves the words of ciphertext y
// (raw + src_off[y], any alignment) to dst + y*words: two aligned 8-byte loads and a funnel shift per word.
// raw must be readable 8 bytes past the last word.  grid (ceil(words/256), ncts)
__global__ void __launch_bounds__(256) strip_headers_kernel(const unsigned char *__restrict__ raw, const u64 *__restrict__ src_off,
                                                            u64 *__restrict__ dst, size_t words) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= words) return;
    const size_t off = (size_t)src_off[blockIdx.y] + 8 * i;
    const unsigned sh = (unsigned)(off & 7) * 8;
    const u64 *a = reinterpret_cast<const u64 *>(raw + (off & ~(size_t)7));
    const u64 lo = a[0];
    const u64 v = sh ? (lo >> sh) | (a[1] << (64 - sh)) : lo;
    dst[(size_t)blockIdx.y * words + i] = v;
}
