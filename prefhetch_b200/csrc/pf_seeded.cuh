// pf_seeded.cuh — expansion of SEAL "seeded" query ciphertexts ON THE DEVICE.
// A symmetric-key SEAL client sends Serializable<Ciphertext>: c0 and, instead of the uniformly random c1, the
// 64-byte seed of the Blake2xb PRNG that drew it — half the upload.  seal::Ciphertext::load re-creates c1 with
// expand_seed = util::sample_poly_uniform over a Blake2xbPRNG [EXT: SEAL 4.1 ciphertext.cpp, util/rlwe.cpp,
// randomgen.cpp, util/blake2xb.c; restated on the host in pf_seal_prng.h, which this file must equal bit for bit].
//
// The PRNG stream is embarrassingly parallel: refill number c of the 4096-byte buffer is
// blake2xb(4096, in = counter c, key = seed) = 64 independent leaves BLAKE2b(node_offset i; root_c), where root_c
// is one keyed BLAKE2b of the counter (2 compressions).  One CTA of 64 threads per refill: thread 0 derives the
// root, every thread one leaf = 8 words of the stream.  sample_poly_uniform first fills all L*N words and only then
// re-draws, limb by limb in position order, the words at or above the largest multiple of q below 2^64 (probability
// q / 2^64 ~ 2^-20 each) from the words that FOLLOW in the stream: stream word p < L*N lands at position p, rejected
// positions are listed (atomic counter, up to SEED_MAX_REJECT per ciphertext), the refill after the last full one
// is kept raw as the "tail", and a one-thread-per-ciphertext fix-up walks the sorted list drawing from the tail.
#pragma once
#include "pf_common.cuh"

#define SEED_MAX_REJECT 96  // rejected words per ciphertext the fix-up can hold (expected: 0.03 at N = 8192)
#define SEED_TAIL_WORDS 512 // one refill
#define SEED_SCRATCH_WORDS (2 + SEED_MAX_REJECT / 2 + SEED_TAIL_WORDS) // per ciphertext: count, pad, u32 list, tail

__device__ __forceinline__ u64 b2_rotr(u64 x, int r) { return (x >> r) | (x << (64 - r)); }

// one BLAKE2b compression (RFC 7693); m = the 128-byte block as 16 words, t = bytes hashed so far incl. this block
__device__ __forceinline__ void blake2b_compress(u64 h[8], const u64 m[16], u64 t, bool last) {
    const u64 IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                       0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    u64 v[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        v[i] = h[i];
        v[i + 8] = IV[i];
    }
    v[12] ^= t;
    if (last) v[14] = ~v[14];
#define B2_G(a, b, c, d, x, y)          \
    v[a] = v[a] + v[b] + (x);           \
    v[d] = b2_rotr(v[d] ^ v[a], 32);    \
    v[c] = v[c] + v[d];                 \
    v[b] = b2_rotr(v[b] ^ v[c], 24);    \
    v[a] = v[a] + v[b] + (y);           \
    v[d] = b2_rotr(v[d] ^ v[a], 16);    \
    v[c] = v[c] + v[d];                 \
    v[b] = b2_rotr(v[b] ^ v[c], 63);
#define B2_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    B2_G(0, 4, 8, 12, m[s0], m[s1])                                                       \
    B2_G(1, 5, 9, 13, m[s2], m[s3])                                                       \
    B2_G(2, 6, 10, 14, m[s4], m[s5])                                                      \
    B2_G(3, 7, 11, 15, m[s6], m[s7])                                                      \
    B2_G(0, 5, 10, 15, m[s8], m[s9])                                                      \
    B2_G(1, 6, 11, 12, m[s10], m[s11])                                                    \
    B2_G(2, 7, 8, 13, m[s12], m[s13])                                                     \
    B2_G(3, 4, 9, 14, m[s14], m[s15])
    B2_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B2_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    B2_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
    B2_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
    B2_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
    B2_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
    B2_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
    B2_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
    B2_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
    B2_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
    B2_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B2_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
#undef B2_ROUND
#undef B2_G
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}

// the byte stream may sit at any alignment inside the uploaded blob
__device__ __forceinline__ u64 load_u64_unaligned(const unsigned char *p) {
    u64 v = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) v |= (u64)p[i] << (8 * i);
    return v;
}

struct SeededParams {
    const unsigned char *raw; // uploaded request bytes
    const u64 *src_off;       // per ciphertext: byte offset of its first data word (stream start + 113) inside raw
    u64 *dst;                 // ciphertext c at dst + c * 2*L*N words: c0 copied by the caller, c1 written here
    u64 *scratch;             // per ciphertext SEED_SCRATCH_WORDS: [0] rejected count, [2..] u32 positions, then the tail
    const DevModulus *mods;
    int L, N;
};

__device__ __forceinline__ u64 seed_max_multiple(u64 q) { return 0xFFFFFFFFFFFFFFFFULL - (0xFFFFFFFFFFFFFFFFULL % q) - 1; }

// grid (L*N/512 + 1, ncts), 64 threads: CTA (r, c) = refill r of ciphertext c's PRNG; the last r is the tail
__global__ void __launch_bounds__(64) seeded_expand_kernel(const SeededParams p) {
    __shared__ u64 root[8];
    const int c = blockIdx.y;
    const u64 half = (u64)p.L * p.N; // words of one polynomial
    const unsigned char *seed = p.raw + p.src_off[c] + half * 8 + 17; // after c0: nested SEALHeader (16), prng type (1)
    u64 *scr = p.scratch + (size_t)c * SEED_SCRATCH_WORDS;
    const u64 IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                       0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    if (threadIdx.x == 0) {
        // root = BLAKE2b(digest 64, key 64 bytes, fanout 1, depth 1, xof_length 4096; key block, then the counter)
        u64 h[8], m[16];
#pragma unroll
        for (int i = 0; i < 8; i++) h[i] = IV[i];
        h[0] ^= 0x0000000001014040ULL;          // digest_length 64 | key_length 64 << 8 | fanout 1 << 16 | depth 1 << 24
        h[1] ^= (u64)4096 << 32;                // node_offset 0 (low 32 bits), xof_length 4096 (high 32 bits)
#pragma unroll
        for (int i = 0; i < 8; i++) m[i] = load_u64_unaligned(seed + 8 * i);
#pragma unroll
        for (int i = 8; i < 16; i++) m[i] = 0;
        blake2b_compress(h, m, 128, false);
        m[0] = (u64)blockIdx.x;                 // the PRNG's refill counter, 8 bytes little endian
#pragma unroll
        for (int i = 1; i < 8; i++) m[i] = 0;
        blake2b_compress(h, m, 136, true);
#pragma unroll
        for (int i = 0; i < 8; i++) root[i] = h[i];
    }
    __syncthreads();
    // leaf i: digest 64, key 0, fanout 0, depth 0, leaf_length 64, node_offset i, xof_length 4096, node_depth 0, inner 64
    u64 h[8], m[16];
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = IV[i];
    h[0] ^= 0x0000004000000040ULL;              // digest_length 64, leaf_length 64 << 32
    h[1] ^= (u64)threadIdx.x | ((u64)4096 << 32);
    h[2] ^= (u64)64 << 8;                       // node_depth 0, inner_length 64
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = root[i];
#pragma unroll
    for (int i = 8; i < 16; i++) m[i] = 0;
    blake2b_compress(h, m, 64, true);
    const u64 t0 = (u64)blockIdx.x * 512 + (u64)threadIdx.x * 8; // stream index of this thread's first word
    if (t0 >= half) {                           // the tail refill: kept raw for the fix-up
        u64 *tail = scr + 2 + SEED_MAX_REJECT / 2;
#pragma unroll
        for (int k = 0; k < 8; k++) tail[threadIdx.x * 8 + k] = h[k];
        return;
    }
    const int j = (int)(t0 / (u64)p.N);         // 8 consecutive words never straddle a limb (N is a multiple of 8)
    const u64 q = p.mods[j].q, ratio1 = p.mods[j].ratio1, mm = seed_max_multiple(q);
    u64 *c1 = p.dst + (size_t)c * 2 * half + half;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const u64 r = h[k];
        if (r >= mm) {
            const unsigned slot = (unsigned)atomicAdd(reinterpret_cast<unsigned long long *>(scr), 1ULL);
            if (slot < SEED_MAX_REJECT) reinterpret_cast<u32 *>(scr + 2)[slot] = (u32)(t0 + k);
            c1[t0 + k] = 0;
        } else {
            c1[t0 + k] = barrett64(r, q, ratio1);
        }
    }
}

// one thread per ciphertext: re-draw the rejected words in position order from the tail.  *err gets bit 62 set
// when a ciphertext had more rejections than the list or the tail holds (expected never: p ~ 1e-100)
__global__ void __launch_bounds__(64) seeded_fixup_kernel(const SeededParams p, int ncts, volatile unsigned long long *err) {
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c >= ncts) return;
    u64 *scr = p.scratch + (size_t)c * SEED_SCRATCH_WORDS;
    const unsigned n = (unsigned)scr[0];
    if (!n) return;
    if (n > SEED_MAX_REJECT) {
        *err = 0x4000000000000000ULL | (unsigned long long)n;
        return;
    }
    u32 *pos = reinterpret_cast<u32 *>(scr + 2);
    for (unsigned a = 1; a < n; a++) {          // insertion sort: the atomics filled the list in arrival order
        const u32 x = pos[a];
        unsigned b = a;
        for (; b > 0 && pos[b - 1] > x; b--) pos[b] = pos[b - 1];
        pos[b] = x;
    }
    const u64 half = (u64)p.L * p.N;
    const u64 *tail = scr + 2 + SEED_MAX_REJECT / 2;
    u64 *c1 = p.dst + (size_t)c * 2 * half + half;
    unsigned u = 0;
    for (unsigned a = 0; a < n; a++) {
        const int j = (int)(pos[a] / (u32)p.N);
        const u64 q = p.mods[j].q, mm = seed_max_multiple(q);
        u64 r;
        do {
            if (u >= SEED_TAIL_WORDS) {
                *err = 0x4000000000000000ULL | (unsigned long long)n;
                return;
            }
            r = tail[u++];
        } while (r >= mm);
        c1[pos[a]] = barrett64(r, q, p.mods[j].ratio1);
    }
}

// c0 of seeded ciphertext y: half = L*N words from raw + src_off[y] (any alignment) to dst + y * 2*half.
// raw must be readable 8 bytes past the last word.  grid (ceil(half/256), ncts)
__global__ void __launch_bounds__(256) strip_seeded_kernel(const unsigned char *__restrict__ raw, const u64 *__restrict__ src_off,
                                                           u64 *__restrict__ dst, size_t half) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= half) return;
    const size_t off = (size_t)src_off[blockIdx.y] + 8 * i;
    const unsigned sh = (unsigned)(off & 7) * 8;
    const u64 *a = reinterpret_cast<const u64 *>(raw + (off & ~(size_t)7));
    const u64 lo = a[0];
    const u64 v = sh ? (lo >> sh) | (a[1] << (64 - sh)) : lo;
    dst[(size_t)blockIdx.y * 2 * half + i] = v;
}
