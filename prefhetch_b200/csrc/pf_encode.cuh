// pf_encode.cuh — index-build kernels: lay a block of integer vectors out as the K plaintext
// diagonals (+ the norm plaintext) of the generalised-diagonal layout, in
// BatchEncoder slot order scattered through SEAL's matrix_reps_index_map, ready for the inverse
// NTT mod t (BatchEncoder::encode), the centred lift + forward NTT per limb
// (Evaluator::transform_to_ntt_inplace(Plaintext)) and, for the norms, the BFV scaling variant
// (util/scalingvariant.cpp multiply_add_plain_with_scaling_variant).
#pragma once
#include "pf_common.cuh"

struct EncodeBlock {
    long long vec_offset; // first vector of the block in the list-ordered base array
    u32 nvec;
    u32 pad;
};

struct EncodeParams {
    const unsigned char *base; // [ntotal][d] uint8, list order
    const EncodeBlock *blocks; // per blockIdx.z
    const u32 *inv_index_map;  // coefficient position -> slot
    u64 *plain;                // [z][K+1][N] values mod t placed at their pre-INTT positions
    u64 t;
    int N, d, dc, R, K, g;
};

// grid (N/256, K+1, zblocks)
__global__ void __launch_bounds__(256) encode_slots_kernel(const EncodeParams p) {
    const int pos = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    const EncodeBlock b = p.blocks[z];
    const int S = p.N >> 1, per_row = S / p.g;
    const int slot = (int)p.inv_index_map[pos];
    const int row = slot / S, s = slot % S;
    const int within = s % p.dc, grp = s / p.dc;
    const int c = grp * p.R + (within % p.R);
    const u32 u = (u32)(row * per_row + c);
    u64 v = 0;
    if (u < b.nvec) {
        const unsigned char *x = p.base + (size_t)(b.vec_offset + u) * p.d;
        if (y < p.K) {
            const int a = y / p.R, r = y % p.R;
            const int dim = a * p.dc + ((within + r) % p.dc);
            if (dim < p.d) {
                const u64 two_x = 2ull * x[dim];
                v = two_x ? p.t - (two_x % p.t) : 0;
                if (v == p.t) v = 0;
            }
        } else if (within / p.R == 0) {
            u64 sq = 0;
            for (int k = 0; k < p.d; k++) sq += (u64)x[k] * x[k];
            v = sq % p.t;
        }
    }
    p.plain[((size_t)z * (p.K + 1) + y) * p.N + pos] = v;
}

struct ScaleParams {
    const u64 *plain; // [z][..][N]; the norm plaintext of block z is at plain + z*plain_sz + plain_off
    long long plain_sz, plain_off;
    u64 *out;         // [z][L][N] coefficient form, = round(Q/t * m) per limb
    const DevModulus *mods;
    u64 t, q_mod_t, half_t; // half_t = (t+1)>>1
    u64 delta_mod_q[17];
    int N, L;
};

// grid (N/256, 1, zblocks)
__global__ void __launch_bounds__(256) scale_plain_kernel(const ScaleParams p) {
    const int i = blockIdx.x * 256 + threadIdx.x, z = blockIdx.z;
    const u64 mval = p.plain[(size_t)z * p.plain_sz + p.plain_off + i];
    const u64 fix = (mval * p.q_mod_t + p.half_t) / p.t; // t < 2^32 enforced on the host
    for (int l = 0; l < p.L; l++) {
        const DevModulus m = p.mods[l];
        u64 lo = mval * p.delta_mod_q[l], hi = __umul64hi(mval, p.delta_mod_q[l]);
        lo += fix;
        hi += (lo < fix);
        p.out[((size_t)z * p.L + l) * p.N + i] = barrett128(lo, hi, m.q, m.ratio0, m.ratio1);
    }
}

// values[N] (slot order) -> pre-INTT coefficient positions (BatchEncoder::encode scatter). grid (N/256)
__global__ void __launch_bounds__(256) slot_scatter_kernel(const u64 *values, const u32 *inv_index_map, u64 *plain) {
    const int pos = blockIdx.x * 256 + threadIdx.x;
    plain[pos] = values[inv_index_map[pos]];
}

// float [n] -> uint8, flags any value that is not an integer in [0,255]
__global__ void __launch_bounds__(256) quantize_u8_kernel(const float *in, unsigned char *out, size_t n, int *bad) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float v = in[i];
    const int iv = (int)v;
    if (!(v >= 0.0f && v <= 255.0f) || (float)iv != v) {
        *bad = 1;
        out[i] = 0;
    } else {
        out[i] = (unsigned char)iv;
    }
}
