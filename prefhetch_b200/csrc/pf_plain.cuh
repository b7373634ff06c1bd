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
__global__ void flag_wait_kernel(const volatile unsigned *flag, unsigned value) {
    while (*flag < value) __nanosleep(500);
    __threadfence_system();
}
