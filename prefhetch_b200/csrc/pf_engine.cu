// pf_engine.cu — engine and C ABI (include/prefhetch_b200.h) of the B200-native PreFHEtch search
// hot path.  Host orchestration only: every arithmetic step runs in the sm_100a kernels of
// pf_ntt.cuh / pf_mac.cuh / pf_keyswitch.cuh / pf_encode.cuh / pf_plain.cuh.  There is no CPU
// fallback: without a CUDA device pf_engine_create fails with PF_ERR_CUDA.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/prefhetch_b200.h"
#include "pf_common.cuh"
#include "pf_encode.cuh"
#include "pf_blake2b.h"
#include "pf_host_math.h"
#include "pf_seal_prng.h"
#include <dlfcn.h>
#include <zlib.h>
#include "pf_keyswitch.cuh"
#include "pf_mac.cuh"
#include "pf_ntt.cuh"
#include "pf_ntt_fp.cuh"
#include "pf_plain.cuh"
#include "pf_seeded.cuh"

#ifndef PF_MAC_DEFAULT_VARIANT
#define PF_MAC_DEFAULT_VARIANT 0
#endif

namespace {

thread_local std::string g_tls_error;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    cudaError_t ensure(size_t n) {
        if (n <= bytes) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    // per-step scratch whose size follows the batch (pairs, jobs): grow with 25 % headroom, so that a
    // step slightly larger than every step before it does not pay cudaFree (a device-wide
    // synchronisation) + cudaMalloc of hundreds of MB in the middle of a serving loop (measured: one
    // such reallocation stalled the host for 170-670 ms)
    cudaError_t ensure_grow(size_t n) {
        if (n <= bytes) return cudaSuccess;
        return ensure(n + n / 4 + ((size_t)1 << 20));
    }
    template <class T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

// host memory the GPU reads / writes in place (zero-copy): small results that must not queue behind a
// large response download on the device-to-host copy engine
struct MappedBuf {
    void *h = nullptr, *d = nullptr;
    size_t bytes = 0;
    ~MappedBuf() {
        if (h) cudaFreeHost(h);
    }
    cudaError_t ensure(size_t n) {
        if (n <= bytes) return cudaSuccess;
        if (h) cudaFreeHost(h);
        h = d = nullptr;
        bytes = 0;
        n += n / 4 + 4096;
        cudaError_t e = cudaHostAlloc(&h, n, cudaHostAllocMapped);
        if (e != cudaSuccess) return e;
        e = cudaHostGetDevicePointer(&d, h, 0);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
};

struct GaloisKey {
    DevBuf key;  // [L][2][k][N]
    DevBuf perm; // u32[N]
    DevBuf km;   // [2][L+1][N] hoisting correction term (pf_keyswitch.cuh)
    u32 einv = 0;
};

struct BlockInfo {
    long long vec_offset;
    u32 nvec;
    u32 list;
};

constexpr size_t SEAL_CT_HEADER = 16 + 32 + 1 + 8 * 5 + 16 + 8; // bytes before the data words (113)
constexpr size_t PF_RESULT_DATA_OFFSET = 128;                      // words of result r start at r*slot + 128
constexpr size_t PF_RESULT_PAD = PF_RESULT_DATA_OFFSET - SEAL_CT_HEADER; // its SEAL stream starts here
static_assert(SEAL_CT_HEADER == 113, "stamp_headers_kernel copies 113 bytes");

} // namespace

// inflated-size ceilings for client-supplied zlib streams (inflate_seal_stream)
#define PF_MAX_GALOIS_KEYS 256
#define PF_INFLATE_HARD_CAP ((size_t)6 << 30)
#define PF_E2E_GROUPS 8
#define PF_MAX_FLIGHTS 4 // max query groups of pf_search_lists_encrypted (copy / compute overlap)

struct pf_engine {
    pf_params prm{};
    int N = 0, logn = 0, k = 0, L = 0;
    int Lr = 0; // limbs of result ciphertexts (<= L)
    uint64_t result_pid[4] = {0, 0, 0, 0};
    bool result_pid_set = false;
    DevBuf d_mstab, s_full, s_cksum, s_ctoff;
    DevBuf s_seed; // device-side seeded expansion: SEED_SCRATCH_WORDS per ciphertext (pf_seeded.cuh)
    u32 d = 0, d_pad = 0, m = 0, g = 0, dc = 0, R = 0, K = 0, C = 0;
    u64 t = 0;
    std::mutex mu;
    mutable std::string err;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::atomic<uint64_t> launches{0};
    // host-mapped error word written by device code that gives up (flag_wait_kernel time-out); every
    // API call that enqueues or waits for work looks at it first (check_device_error)
    unsigned long long *h_err_word = nullptr, *d_err_word = nullptr;
    unsigned long long flag_timeout_ns = 20ull * 1000 * 1000 * 1000;

    // device tables
    DevBuf d_mods; // DevModulus[k+1] (index k = plain modulus)
    DevBuf d_tw;   // Twiddle[k+1][2][N]
    DevBuf d_tw_fp; // double[k+1][2][N] centred twiddles
    DevBuf d_mstab_fp; // double[MS_MAXL][MS_MAXL][2] mod-switch constants for the FP64 kernel
    DevBuf d_tw_fp_lane; // double[k+1][2][N] lane-major copies for the pass over bits [4..0]
    bool ntt_fp = false; // FP64-pipe NTT kernels eligible (pf_ntt_fp.cuh)
    DevBuf d_inv_index_map;
    std::vector<u64> h_q;
    u64 q_mod_t = 0;
    std::vector<u64> delta_mod_q, p_half_mod_q, p_inv_mod_q, p_inv_mod_q_sh;
    u64 p_half = 0;
    int max_prime_bits = 0;
    bool mac_wide = false;
    bool mac_fpred = false; // FP64-assisted final reduction applies (pf_mac.cuh)
    bool ks_fpred = false;  // same for the key-switch inner product (L terms)
    uint64_t level_pid[PF_MAX_PRIMES + 1][4] = {}; // SEAL parms_id of the level with i data limbs (pf_blake2b.h)

    std::map<u32, GaloisKey> gkeys;

    // index
    bool has_index = false;
    u64 nlist = 0, ntotal = 0;
    std::vector<float> h_centroids;
    std::vector<long long> h_list_offsets;
    std::vector<long long> h_ids;
    DevBuf d_centroids, d_ids, d_base_f32, d_base_u8, d_pos_of_id;
    DevBuf d_pq_cent, d_pq_codes; // product quantizer of the loaded index (pf_load_pq): [M][256][d/M] floats, [ntotal][M] bytes
    u32 pq_M = 0;                 // 0 = none loaded
    bool ids_are_rows = false;
    std::vector<BlockInfo> blocks;            // local blocks
    std::vector<long long> list_block_start;  // [nlist+1] into blocks (0 length for lists not owned)
    DevBuf d_diag, d_norm;
    size_t diag_block_words = 0, norm_block_words = 0;

    // scratch
    DevBuf s_x, s_cx, s_dist, s_keys, s_jobs, s_pl_dist, s_pl_labels, s_ids;
    DevBuf s_rot, s_cqntt, s_hoistD, s_flags, s_ks_d, s_ks_S, s_ks_W, s_rotjobs, s_c1coef, s_chunks, s_pairblock, s_pairout, s_qcts, s_tmp, s_plain,
        s_encblocks;
    MappedBuf m_cx, m_cidx, m_cdist; // stage 1: query vectors in, probe ids / distances out (zero-copy)
    // pinned upload arena: pageable cudaMemcpyAsync would synchronise the stream (and the host) on every
    // small table upload; 4 call slots, a slot is reused only after the call that used it has finished
    MappedBuf h_arena; // host-mapped: the GPU fetches the tables with a kernel, no copy engine involved
    size_t arena_slot_bytes = 0, arena_off = 0;
    int arena_slot = 0;
    cudaEvent_t arena_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool arena_ev_used[4] = {false, false, false, false};
    cudaStream_t copy_stream = nullptr;
    cudaStream_t coarse_stream = nullptr; // stage 1 runs beside the encrypted pipeline of the previous batch
    cudaEvent_t ev_group[2] = {nullptr, nullptr};
    cudaStream_t upload_stream = nullptr; // H2D of the query ciphertexts (pf_search_submit)
    // searches in flight (pf_search_submit .. pf_search_collect): own query / result buffers each
    struct Flight {
        bool busy = false;
        uint64_t id = 0;
        DevBuf qcts, out, qraw; // qraw: the uploaded byte range of the query blob (headers still in place)
        cudaEvent_t ev_up[PF_E2E_GROUPS] = {};
        cudaEvent_t done = nullptr;
        cudaEvent_t tl[6] = {}; // PF_DEBUG_TIMELINE: upload begin/end, compute begin/end, download begin/end
    } flights[PF_MAX_FLIGHTS];
    bool timeline = false;
    cudaEvent_t tl_base = nullptr;
    uint64_t next_ticket = 0;
    int groups_hint = 0; // pf_search_set_groups

    // timing
    bool timing = false;
    float t_ms[PF_T_COUNT] = {0};
    uint64_t t_launch[PF_T_COUNT] = {0};
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending_events;
    std::vector<cudaEvent_t> event_pool;

    int fail(int code, const char *fmt, ...) const {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        err = buf;
        g_tls_error = buf;
        return code;
    }
};

namespace {

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return e->fail(PF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                           __LINE__);                                                                    \
    } while (0)

// PF_DEBUG_HOST=1: host-side time of every section of a search call, printed as it happens;
// PF_DEBUG_HOST=2: per-section totals (count, sum, max), printed when the process exits.
struct HostTickTotals {
    std::mutex mu;
    std::map<std::string, std::array<double, 3>> t;
    ~HostTickTotals() {
        for (auto &kv : t)
            fprintf(stderr, "[pf host total] %-18s n=%6.0f sum=%10.1f us avg=%8.1f us max=%9.1f us\n", kv.first.c_str(), kv.second[0],
                    kv.second[1], kv.second[1] / std::max(1.0, kv.second[0]), kv.second[2]);
    }
};
static HostTickTotals g_tick_totals;
struct HostTick {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    const char *what;
    explicit HostTick(const char *w) : what(w) {}
    ~HostTick() {
        static const int mode = getenv("PF_DEBUG_HOST") ? std::max(1, atoi(getenv("PF_DEBUG_HOST"))) : 0;
        if (!mode) return;
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        if (mode == 1) {
            fprintf(stderr, "[pf host] %-18s %8.1f us\n", what, us);
        } else {
            std::lock_guard<std::mutex> lk(g_tick_totals.mu);
            auto &a = g_tick_totals.t[what];
            a[0] += 1;
            a[1] += us;
            a[2] = std::max(a[2], us);
        }
    }
};

// A device-side give-up (flag wait that timed out) is sticky: it surfaces as PF_ERR_CUDA on every later call.
int check_device_error(pf_engine *e) {
    if (!e->h_err_word) return PF_OK;
    const unsigned long long w = *(volatile unsigned long long *)e->h_err_word;
    if (!w) return PF_OK;
    if (!(w >> 63) && ((w >> 62) & 1))
        return e->fail(PF_ERR_CUDA, "device-side expansion of a seeded ciphertext gave up: %u rejected PRNG words in one ciphertext (set PF_SEEDED_HOST=1)",
                       (unsigned)(w & 0xffffffffu));
    return e->fail(PF_ERR_CUDA, "peer flag wait timed out after %.1f s (flag value %u, waiting for %u): a peer rank is dead or out of protocol order",
                   (double)e->flag_timeout_ns * 1e-9, (unsigned)((w >> 32) & 0x7fffffffu), (unsigned)(w & 0xffffffffu));
}

// ---- pinned upload arena --------------------------------------------------------------------
constexpr size_t ARENA_SLOT = (size_t)8 << 20;

int arena_begin(pf_engine *e) { // call once at the start of an API call that uploads tables
    if (!e->h_arena.h) {
        CK(e->h_arena.ensure(4 * ARENA_SLOT));
        for (auto &ev : e->arena_ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    e->arena_slot = (e->arena_slot + 1) & 3;
    if (e->arena_ev_used[e->arena_slot]) CK(cudaEventSynchronize(e->arena_ev[e->arena_slot]));
    e->arena_off = 0;
    return PF_OK;
}

void arena_end(pf_engine *e) { // after the last upload of the call has been enqueued
    cudaEventRecord(e->arena_ev[e->arena_slot], e->stream);
    e->arena_ev_used[e->arena_slot] = true;
}

// Small per-call tables (pair plan, rotation jobs; tens of KB) reach the device WITHOUT the copy engine:
// they are written into a host-mapped arena slot and a few-CTA kernel on the engine stream pulls them into
// device scratch.  A cudaMemcpyAsync here queues behind the bulk query-ciphertext uploads of the following
// searches on the H2D engine — measured with PF_DEBUG_TIMELINE: the whole compute of a search stalled 2-3 ms
// behind 34 MB of uploads it did not depend on.  (Falls back to a direct copy, which synchronises, when the
// slot is exhausted.)
__global__ void __launch_bounds__(256) fetch_words_kernel(u32 *dst, const u32 *src, size_t nwords) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < nwords) dst[i] = src[i];
}

cudaError_t upload_async(pf_engine *e, void *dst, const void *src, size_t bytes) {
    const size_t aligned = (bytes + 255) & ~(size_t)255;
    if (!e->h_arena.h || e->arena_off + aligned > ARENA_SLOT || (bytes & 3))
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, e->stream);
    const size_t off = (size_t)e->arena_slot * ARENA_SLOT + e->arena_off;
    memcpy((char *)e->h_arena.h + off, src, bytes);
    e->arena_off += aligned;
    const size_t nwords = bytes / 4;
    if (!nwords) return cudaSuccess;
    fetch_words_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, e->stream>>>((u32 *)dst, (const u32 *)((char *)e->h_arena.d + off), nwords);
    e->launches++;
    return cudaGetLastError();
}

// ---- timing -------------------------------------------------------------------------------
struct PhaseTimer {
    pf_engine *e;
    int phase;
    cudaEvent_t a = nullptr, b = nullptr;
    uint64_t launches0;
    cudaStream_t st;
    PhaseTimer(pf_engine *e_, int phase_, cudaStream_t st_ = nullptr)
        : e(e_), phase(phase_), launches0(e_->launches), st(st_ ? st_ : e_->stream) {
        if (!e->timing) return;
        auto get = [&]() {
            cudaEvent_t ev;
            if (!e->event_pool.empty()) {
                ev = e->event_pool.back();
                e->event_pool.pop_back();
            } else {
                cudaEventCreate(&ev);
            }
            return ev;
        };
        a = get();
        b = get();
        cudaEventRecord(a, st);
    }
    ~PhaseTimer() {
        e->t_launch[phase] += e->launches - launches0;
        if (!e->timing) return;
        cudaEventRecord(b, st);
        e->pending_events.push_back({phase, {a, b}});
    }
};

void drain_events(pf_engine *e) {
    for (auto &pe : e->pending_events) {
        float ms = 0;
        if (cudaEventSynchronize(pe.second.second) == cudaSuccess &&
            cudaEventElapsedTime(&ms, pe.second.first, pe.second.second) == cudaSuccess)
            e->t_ms[pe.first] += ms;
        e->event_pool.push_back(pe.second.first);
        e->event_pool.push_back(pe.second.second);
    }
    e->pending_events.clear();
}

// ---- NTT launch helpers -------------------------------------------------------------------
template <int LOGN>
cudaError_t set_ntt_attrs() {
    cudaError_t r;
    const int smem = (int)NttCfg<LOGN>::SMEM;
#define SETATTR(K)                                                                  \
    r = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    if (r != cudaSuccess) return r;
    SETATTR((ntt_fwd_kernel<LOGN, NTT_IN_PLAIN>));
    SETATTR((ntt_fwd_kernel<LOGN, NTT_IN_LIFT>));
    SETATTR((ntt_fwd_kernel<LOGN, NTT_IN_REDUCE>));
    SETATTR((ntt_fwd_kernel<LOGN, NTT_IN_GALOIS_REDUCE>));
    SETATTR((ntt_inv_kernel<LOGN>));
    SETATTR((ntt_fwd_fp_kernel<LOGN, NTT_IN_PLAIN>));
    SETATTR((ntt_fwd_fp_kernel<LOGN, NTT_IN_LIFT>));
    SETATTR((ntt_fwd_fp_kernel<LOGN, NTT_IN_REDUCE>));
    SETATTR((ntt_fwd_fp_kernel<LOGN, NTT_IN_GALOIS_REDUCE>));
    SETATTR((ntt_fwd_fp_kernel<LOGN, NTT_IN_MODDOWN>));
    SETATTR((ntt_inv_fp_kernel<LOGN>));
#undef SETATTR
    return cudaSuccess;
}

template <int LOGN>
void launch_ntt_fp_t(int inmode, bool inverse, const NttParams &p, dim3 grid, cudaStream_t s) {
    const size_t smem = NttCfg<LOGN>::SMEM;
    const int nt = NttCfg<LOGN>::NT;
    if (inverse)
        ntt_inv_fp_kernel<LOGN><<<grid, nt, smem, s>>>(p);
    else if (inmode == NTT_IN_PLAIN)
        ntt_fwd_fp_kernel<LOGN, NTT_IN_PLAIN><<<grid, nt, smem, s>>>(p);
    else if (inmode == NTT_IN_LIFT)
        ntt_fwd_fp_kernel<LOGN, NTT_IN_LIFT><<<grid, nt, smem, s>>>(p);
    else if (inmode == NTT_IN_REDUCE)
        ntt_fwd_fp_kernel<LOGN, NTT_IN_REDUCE><<<grid, nt, smem, s>>>(p);
    else if (inmode == NTT_IN_MODDOWN)
        ntt_fwd_fp_kernel<LOGN, NTT_IN_MODDOWN><<<grid, nt, smem, s>>>(p);
    else
        ntt_fwd_fp_kernel<LOGN, NTT_IN_GALOIS_REDUCE><<<grid, nt, smem, s>>>(p);
}

template <int LOGN>
void launch_ntt_t(int inmode, bool inverse, const NttParams &p, dim3 grid, cudaStream_t s) {
    const size_t smem = NttCfg<LOGN>::SMEM;
    const int nt = NttCfg<LOGN>::NT;
    if (inverse)
        ntt_inv_kernel<LOGN><<<grid, nt, smem, s>>>(p);
    else if (inmode == NTT_IN_PLAIN)
        ntt_fwd_kernel<LOGN, NTT_IN_PLAIN><<<grid, nt, smem, s>>>(p);
    else if (inmode == NTT_IN_LIFT)
        ntt_fwd_kernel<LOGN, NTT_IN_LIFT><<<grid, nt, smem, s>>>(p);
    else if (inmode == NTT_IN_REDUCE)
        ntt_fwd_kernel<LOGN, NTT_IN_REDUCE><<<grid, nt, smem, s>>>(p);
    else
        ntt_fwd_kernel<LOGN, NTT_IN_GALOIS_REDUCE><<<grid, nt, smem, s>>>(p);
}

void launch_ntt(pf_engine *e, int inmode, bool inverse, NttParams p, dim3 grid) {
    p.mods = e->d_mods.as<DevModulus>();
    p.tw = e->d_tw.as<Twiddle>();
    p.tw_fp = e->d_tw_fp.as<double>();
    p.tw_fp_lane = e->d_tw_fp_lane.as<double>();
    e->launches++;
    if (e->ntt_fp) {
        if (!inverse && inmode == NTT_IN_GALOIS_REDUCE) {
            // hoisted jobs need the exact digits only when flagged: 32 jobs per CTA; un-hoisted jobs
            // (chain mode) always do: one job per CTA
            p.njobs = (int)grid.z;
            p.job_group = p.hoisted_jobs ? 32 : 1;
            grid.z = (grid.z + p.job_group - 1) / p.job_group;
        }
        switch (e->logn) {
        case 10: launch_ntt_fp_t<10>(inmode, inverse, p, grid, e->stream); break;
        case 11: launch_ntt_fp_t<11>(inmode, inverse, p, grid, e->stream); break;
        case 12: launch_ntt_fp_t<12>(inmode, inverse, p, grid, e->stream); break;
        case 13: launch_ntt_fp_t<13>(inmode, inverse, p, grid, e->stream); break;
        case 14: launch_ntt_fp_t<14>(inmode, inverse, p, grid, e->stream); break;
        }
        return;
    }
    // gridDim.y/z limits: z <= 65535, y <= 65535
    switch (e->logn) {
    case 10: launch_ntt_t<10>(inmode, inverse, p, grid, e->stream); break;
    case 11: launch_ntt_t<11>(inmode, inverse, p, grid, e->stream); break;
    case 12: launch_ntt_t<12>(inmode, inverse, p, grid, e->stream); break;
    case 13: launch_ntt_t<13>(inmode, inverse, p, grid, e->stream); break;
    case 14: launch_ntt_t<14>(inmode, inverse, p, grid, e->stream); break;
    }
}

// transform `count` consecutive [L][N] polynomials (limb of polynomial y*L + x is x), in place or not
void ntt_limbs(pf_engine *e, const u64 *in, u64 *out, size_t count, bool inverse) {
    const size_t N = e->N;
    size_t done = 0;
    while (done < count) {
        const size_t n = std::min<size_t>(count - done, 32768);
        NttParams p{};
        p.in = in + done * e->L * N;
        p.out = out + done * e->L * N;
        p.in_sx = p.out_sx = (long long)N;
        p.in_sy = p.out_sy = (long long)(e->L * N);
        for (int i = 0; i < e->L; i++) p.mod_map[i] = i;
        launch_ntt(e, NTT_IN_PLAIN, inverse, p, dim3(e->L, (unsigned)n, 1));
        done += n;
    }
}

// ---- engine construction -------------------------------------------------------------------
int build_tables(pf_engine *e) {
    using pfh::minimal_primitive_root; using pfh::invmod; using pfh::shoup; using pfh::ratio128; using pfh::bitrev;
    using pfh::big_mul_word; using pfh::big_div_word; using pfh::big_mod_word;
    const int N = e->N, k = e->k, L = e->L, logn = e->logn;
    std::vector<DevModulus> mods(k + 1);
    std::vector<Twiddle> tw((size_t)(k + 1) * 2 * N);
    std::vector<double> twd((size_t)(k + 1) * 2 * N);
    auto centred = [](u64 x, u64 q) { return x > q / 2 ? -(double)(q - x) : (double)x; };
    for (int j = 0; j <= k; j++) {
        const u64 q = (j == k) ? e->t : e->h_q[j];
        const u64 psi = minimal_primitive_root(2ull * N, q);
        if (!psi) return e->fail(PF_ERR_INVALID, "modulus %llu has no primitive 2N-th root", (unsigned long long)q);
        const u64 psi_inv = invmod(psi, q), n_inv = invmod((u64)N % q, q);
        DevModulus &m = mods[j];
        m.q = q;
        u64 r0, r1;
        ratio128(q, r0, r1);
        m.ratio0 = r0;
        m.ratio1 = r1;
        m.n_inv = n_inv;
        m.n_inv_sh = shoup(n_inv, q);
        Twiddle *f = tw.data() + (size_t)j * 2 * N, *inv = f + N;
        u64 p = 1, ip = 1;
        for (int i = 0; i < N; i++) {
            const uint32_t r = bitrev((uint32_t)i, logn);
            f[r].x = p;
            f[r].y = shoup(p, q);
            inv[r].x = ip;
            inv[r].y = shoup(ip, q);
            p = pfh::mulmod(p, psi, q);
            ip = pfh::mulmod(ip, psi_inv, q);
        }
        m.inv_last_w = pfh::mulmod(inv[1].x, n_inv, q);
        m.inv_last_w_sh = shoup(m.inv_last_w, q);
        m.fq = (double)q;
        m.fqinv = 1.0 / (double)q;
        m.fninv = centred(n_inv, q);
        m.flast_w = centred(m.inv_last_w, q);
        {
            double *fd = twd.data() + (size_t)j * 2 * N;
            for (int i = 0; i < N; i++) {
                fd[i] = centred(f[i].x, q);
                fd[N + i] = centred(inv[i].x, q);
            }
        }
        m.split_shift = (u64)((64 - __builtin_clzll(q) + 1) / 2);
        m.p_half_mod = m.p_inv = m.p_inv_sh = m.pad2 = 0;
        m.fpinv = m.fpad = 0.0;
        if (j < L) {
            const u64 P = e->h_q[k - 1];
            m.p_half_mod = (P >> 1) % q;
            m.p_inv = invmod(P % q, q);
            m.p_inv_sh = shoup(m.p_inv, q);
            m.fpinv = centred(m.p_inv, q);
        }
    }
    CK(e->d_mods.ensure(mods.size() * sizeof(DevModulus)));
    CK(cudaMemcpy(e->d_mods.p, mods.data(), mods.size() * sizeof(DevModulus), cudaMemcpyHostToDevice));
    CK(e->d_tw.ensure(tw.size() * sizeof(Twiddle)));
    CK(cudaMemcpy(e->d_tw.p, tw.data(), tw.size() * sizeof(Twiddle), cudaMemcpyHostToDevice));
    CK(e->d_tw_fp.ensure(twd.size() * sizeof(double)));
    CK(cudaMemcpy(e->d_tw_fp.p, twd.data(), twd.size() * sizeof(double), cudaMemcpyHostToDevice));
    {
        // Lane-major copies of the twiddles of the 5-stage pass over bits [4..0] (thread c owns points
        // 32c..32c+31): entry j of thread c at [j*NT + c], so that every twiddle load of a warp is one
        // contiguous 256-byte run instead of 32 addresses 2^U doubles apart (pf_ntt_fp.cuh).
        const int NT = N / 32;
        std::vector<double> lane((size_t)(k + 1) * 2 * N, 0.0);
        for (int j = 0; j <= k; j++) {
            const double *fd = twd.data() + (size_t)j * 2 * N, *id = fd + N;
            double *lf = lane.data() + (size_t)j * 2 * N, *li = lf + N;
            for (int c = 0; c < NT; c++) {
                int slot = 0;
                for (int U = 0; U < 5; U++) // forward: stage U uses tw[2^(logn-5+U) + (c << U) + r]
                    for (int r = 0; r < (1 << U); r++) lf[(size_t)(slot++) * NT + c] = fd[(1 << (logn - 5 + U)) + (c << U) + r];
                slot = 0;
                for (int B = 0; B < 5; B++) // inverse: stage B uses itw[2^(logn-1-B) + (c << (4-B)) + r]
                    for (int r = 0; r < (1 << (4 - B)); r++) li[(size_t)(slot++) * NT + c] = id[(1 << (logn - 1 - B)) + (c << (4 - B)) + r];
            }
        }
        CK(e->d_tw_fp_lane.ensure(lane.size() * sizeof(double)));
        CK(cudaMemcpy(e->d_tw_fp_lane.p, lane.data(), lane.size() * sizeof(double), cudaMemcpyHostToDevice));
    }

    // BatchEncoder index map (SEAL batchencoder.cpp) and its inverse
    std::vector<u32> inv_map(N);
    {
        const u64 row = N >> 1, m2 = 2ull * N;
        u64 pos = 1;
        for (u64 i = 0; i < row; i++) {
            const u64 i1 = (pos - 1) >> 1, i2 = (m2 - pos - 1) >> 1;
            inv_map[bitrev((uint32_t)i1, logn)] = (u32)i;
            inv_map[bitrev((uint32_t)i2, logn)] = (u32)(row | i);
            pos = (pos * 3) & (m2 - 1);
        }
    }
    CK(e->d_inv_index_map.ensure(N * sizeof(u32)));
    CK(cudaMemcpy(e->d_inv_index_map.p, inv_map.data(), N * sizeof(u32), cudaMemcpyHostToDevice));

    // BFV scaling constants: floor(Q/t) mod q_j, Q mod t
    std::vector<u64> Q{1};
    for (int j = 0; j < L; j++) big_mul_word(Q, e->h_q[j]);
    std::vector<u64> quo = Q;
    e->q_mod_t = big_div_word(quo, e->t);
    e->delta_mod_q.resize(L);
    for (int j = 0; j < L; j++) e->delta_mod_q[j] = big_mod_word(quo, e->h_q[j]);
    // mod-switch tables (divide_and_round_q_last): dropping limb c, for j < c
    {
        std::vector<u64> tab((size_t)MS_MAXL * MS_MAXL * 3, 0);
        for (int c = 1; c < L; c++)
            for (int j = 0; j < c; j++) {
                const u64 qc = e->h_q[c], qj = e->h_q[j];
                u64 *t = tab.data() + ((size_t)c * MS_MAXL + j) * 3;
                t[0] = (qc >> 1) % qj;
                t[1] = invmod(qc % qj, qj);
                t[2] = shoup(t[1], qj);
            }
        CK(e->d_mstab.ensure(tab.size() * 8));
        CK(cudaMemcpy(e->d_mstab.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice));
        std::vector<double> tabf((size_t)MS_MAXL * MS_MAXL * 2, 0.0); // the same constants for modswitch_fp_kernel_t
        for (int c = 1; c < L; c++)
            for (int j = 0; j < c; j++) {
                const u64 *t = tab.data() + ((size_t)c * MS_MAXL + j) * 3;
                tabf[((size_t)c * MS_MAXL + j) * 2] = (double)t[0];
                tabf[((size_t)c * MS_MAXL + j) * 2 + 1] = centred(t[1], e->h_q[j]);
            }
        CK(e->d_mstab_fp.ensure(tabf.size() * 8));
        CK(cudaMemcpy(e->d_mstab_fp.p, tabf.data(), tabf.size() * 8, cudaMemcpyHostToDevice));
    }
    // special prime constants
    const u64 P = e->h_q[k - 1];
    e->p_half = P >> 1;
    e->p_half_mod_q.resize(L);
    e->p_inv_mod_q.resize(L);
    e->p_inv_mod_q_sh.resize(L);
    for (int j = 0; j < L; j++) {
        const u64 q = e->h_q[j];
        e->p_half_mod_q[j] = e->p_half % q;
        e->p_inv_mod_q[j] = invmod(P % q, q);
        e->p_inv_mod_q_sh[j] = shoup(e->p_inv_mod_q[j], q);
    }
    return PF_OK;
}

u32 galois_elt_from_step(const pf_engine *e, int step) { // SEAL util/galois.cpp get_elt_from_step
    const u32 n = (u32)e->N, m2 = 2 * n, row = n >> 1;
    if (step == 0) return m2 - 1;
    const u32 pos = (u32)(step < 0 ? -step : step);
    if (pos >= row) return 0;
    const u32 s = step < 0 ? row - pos : pos;
    u64 x = 1;
    for (u32 i = 0; i < s; i++) x = (x * 3) & (m2 - 1);
    return (u32)x;
}

int set_galois_key_words(pf_engine *e, u32 elt, const u64 *words, bool device_src) {
    const u32 N = (u32)e->N;
    if (!(elt & 1) || elt >= 2 * N) return e->fail(PF_ERR_INVALID, "Galois element %u is not odd and < 2N", elt);
    GaloisKey &gk = e->gkeys[elt];
    const size_t words_n = (size_t)e->L * 2 * e->k * N;
    CK(gk.key.ensure(words_n * 8));
    CK(cudaMemcpyAsync(gk.key.p, words, words_n * 8, device_src ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                       e->stream));
    key_split_kernel<<<dim3(N / 256, e->L * 2 * e->k), 256, 0, e->stream>>>(gk.key.as<u64>(), e->d_mods.as<DevModulus>(), e->k, (int)N);
    e->launches++;
    std::vector<u32> perm(N); // SEAL GaloisTool::generate_table_ntt
    for (u32 i = 0; i < N; i++) {
        const u32 rev = pfh::bitrev(i, e->logn);
        const u64 raw = ((u64)elt * (2ull * rev + 1)) >> 1;
        perm[i] = pfh::bitrev((u32)(raw & (N - 1)), e->logn);
    }
    CK(gk.perm.ensure(N * sizeof(u32)));
    CK(cudaMemcpyAsync(gk.perm.p, perm.data(), N * sizeof(u32), cudaMemcpyHostToDevice, e->stream));
    // inverse of elt modulo 2N (elt is odd): Newton iteration on 2-adic inverse
    u32 inv = elt;
    for (int i = 0; i < 5; i++) inv *= 2 - elt * inv;
    gk.einv = inv & (2 * N - 1);
    // hoisting term: M[I] = NTT_I(negation mask of sigma), KM_c[I] = M[I] (.) sum_J (q_J mod q_I) key_J[c][I]
    const int L = e->L, k = e->k;
    CK(gk.km.ensure((size_t)2 * (L + 1) * N * 8));
    CK(e->s_tmp.ensure_grow((size_t)(L + 2) * N * 8));
    u64 *mask = e->s_tmp.as<u64>(), *M = mask + N;
    galois_negmask_kernel<<<N / 256, 256, 0, e->stream>>>(mask, gk.einv, (int)N);
    e->launches++;
    NttParams mp{};
    mp.in = mask;
    mp.out = M;
    mp.out_sx = N;
    for (int I = 0; I <= L; I++) mp.mod_map[I] = (I == L) ? k - 1 : I;
    launch_ntt(e, NTT_IN_PLAIN, false, mp, dim3(L + 1, 1, 1));
    galois_km_kernel<<<dim3(N / 256, L + 1, 2), 256, 0, e->stream>>>(M, gk.key.as<u64>(), gk.km.as<u64>(),
                                                                    e->d_mods.as<DevModulus>(), L, k, (int)N);
    e->launches++;
    CK(cudaStreamSynchronize(e->stream)); // perm is a stack vector
    CK(cudaGetLastError());
    return PF_OK;
}

// ---- rotations ------------------------------------------------------------------------------
template <bool FINISH>
void launch_ks_accumulate(pf_engine *e, const KsParams &kp, dim3 g, int nz) {
    const int L = e->L;
    if (e->ks_fpred) {
        if (L <= 4) ks_accumulate_kernel<4, FINISH, true><<<g, 256, 0, e->stream>>>(kp, nz);
        else if (L <= 8) ks_accumulate_kernel<8, FINISH, true><<<g, 256, 0, e->stream>>>(kp, nz);
        else ks_accumulate_kernel<KS_MAXL, FINISH, true><<<g, 256, 0, e->stream>>>(kp, nz);
    } else {
        if (L <= 4) ks_accumulate_kernel<4, FINISH, false><<<g, 256, 0, e->stream>>>(kp, nz);
        else if (L <= 8) ks_accumulate_kernel<8, FINISH, false><<<g, 256, 0, e->stream>>>(kp, nz);
        else ks_accumulate_kernel<KS_MAXL, FINISH, false><<<g, 256, 0, e->stream>>>(kp, nz);
    }
}

// Run a batch of rotation jobs (already filled on the host) through the key-switch pipeline.
int run_rot_jobs(pf_engine *e, const std::vector<RotJob> &jobs, bool out_split) {
    const int L = e->L, N = e->N, k = e->k;
    const size_t per_d = (size_t)L * (L + 1) * N, per_S = (size_t)2 * (L + 1) * N, per_W = (size_t)2 * L * N;
    // jobs per pass: as many as a 3 GiB workspace allows.  Smaller, L2-resident passes were measured
    // SLOWER (64 jobs: 2.86 ms, 128: 2.24 ms, 512: 1.82 ms per step): launch count and wave tails win.
    static const size_t env_zb = getenv("PF_KS_BATCH") ? (size_t)atoi(getenv("PF_KS_BATCH")) : 0;
    const size_t zmax = env_zb ? env_zb
                               : std::max<size_t>(1, std::min<size_t>(2048, ((size_t)3 << 30) / ((per_d + per_S + per_W) * 8)));
    CK(e->s_rotjobs.ensure_grow(jobs.size() * sizeof(RotJob)));
    CK(upload_async(e, e->s_rotjobs.p, jobs.data(), jobs.size() * sizeof(RotJob)));
    const size_t zb = std::min(zmax, jobs.size());
    CK(e->s_ks_d.ensure_grow(zb * per_d * 8));
    CK(e->s_ks_S.ensure_grow(zb * per_S * 8));
    CK(e->s_ks_W.ensure_grow(zb * per_W * 8));
    KsParams kp{};
    kp.mods = e->d_mods.as<DevModulus>();
    kp.d = e->s_ks_d.as<u64>();
    kp.S = e->s_ks_S.as<u64>();
    kp.W = e->s_ks_W.as<u64>();
    kp.L = L;
    kp.k = k;
    kp.N = N;
    kp.p_half = e->p_half;
    kp.out_split = out_split ? 1 : 0;
    for (int j = 0; j < L; j++) {
        kp.p_half_mod_q[j] = e->p_half_mod_q[j];
        kp.p_inv_mod_q[j] = e->p_inv_mod_q[j];
        kp.p_inv_mod_q_sh[j] = e->p_inv_mod_q_sh[j];
    }
    for (size_t z0 = 0; z0 < jobs.size(); z0 += zb) {
        const unsigned nz = (unsigned)std::min(zb, jobs.size() - z0);
        const RotJob *dj = e->s_rotjobs.as<RotJob>() + z0;
        kp.jobs = dj;
        // 1. d[J][I] = NTT_I(sigma(c1)_J mod q_I): grid (I<L+1, J<L, z)
        NttParams np{};
        np.jobs = dj;
        np.in_sy = N;
        np.out = kp.d;
        np.out_sx = N;
        np.out_sy = (long long)(L + 1) * N;
        np.out_sz = (long long)per_d;
        for (int I = 0; I <= L; I++) np.mod_map[I] = (I == L) ? k - 1 : I;
        for (int J = 0; J < L; J++) np.src_map[J] = J;
        np.out_split = 1;
        np.hoisted_jobs = jobs[z0].D ? 1 : 0;
        launch_ntt(e, NTT_IN_GALOIS_REDUCE, false, np, dim3(L + 1, L, nz));
        // 2. S_c[I]
        // 2a. S_c[P] (special-prime limb only)
        {
            const dim3 g(N / 512, 1, (nz + KS_QT - 1) / KS_QT);
            launch_ks_accumulate<false>(e, kp, g, (int)nz);
        }
        e->launches++;
        // 3. u_c = INTT_P(S_c[L]) in place, then W_c[j] = ((u + P/2) mod P mod q_j) - (P/2 mod q_j).
        // (An inverse NTT with this prep fused into its store measured slower than the pair on B200:
        // 0.69 ms vs 0.40 ms per step, profiles/r1_launches_step_v4.txt; the variant was removed.)
        NttParams ip{};
        ip.in = kp.S + (size_t)L * N;
        ip.out = kp.S + (size_t)L * N;
        ip.in_sy = ip.out_sy = (long long)(L + 1) * N;
        ip.in_sz = ip.out_sz = (long long)per_S;
        ip.mod_map[0] = k - 1;
        launch_ntt(e, NTT_IN_PLAIN, true, ip, dim3(1, 2, nz));
        // 4. W_c[j] and NTT_j(W_c[j]): with the FP64 kernels W is formed while the transform loads u
        //    (NTT_IN_MODDOWN), so neither the prep kernel nor a coefficient-form W exists; the integer
        //    kernels keep the separate prep + in-place transform.  Then 2b. S_c[j] for the data limbs
        //    with the finish fused: out_c[j] = (S_c[j] - NTT_j(W_c[j])) * P^{-1} (+ sigma_ntt(c0)[j]);
        //    S_c[j] never hits memory.  (Fusing the finish into the NTT copy-out instead measured
        //    slower and was removed.)
        const bool no_fuse = getenv("PF_KS_NO_FUSED_PREP") != nullptr; // per call: tests flip it
        NttParams wp{};
        wp.out = kp.W;
        wp.out_sx = N;
        wp.out_sy = (long long)L * N;
        wp.out_sz = (long long)per_W;
        for (int j = 0; j < L; j++) wp.mod_map[j] = j;
        if (e->ntt_fp && !no_fuse) {
            wp.in = kp.S + (size_t)L * N;
            wp.in_sx = 0;
            wp.in_sy = (long long)(L + 1) * N;
            wp.in_sz = (long long)per_S;
            wp.ks_p_half = e->p_half;
            wp.md_pmod = k - 1;
            launch_ntt(e, NTT_IN_MODDOWN, false, wp, dim3(L, 2, nz));
        } else {
            ks_moddown_prep_kernel<<<dim3(N / 256, 2, nz), 256, 0, e->stream>>>(kp);
            e->launches++;
            wp.in = kp.W;
            wp.in_sx = N;
            wp.in_sy = (long long)L * N;
            wp.in_sz = (long long)per_W;
            launch_ntt(e, NTT_IN_PLAIN, false, wp, dim3(L, 2, nz));
        }
        {
            const dim3 g(N / 512, L, (nz + KS_QT - 1) / KS_QT);
            launch_ks_accumulate<true>(e, kp, g, (int)nz);
        }
        e->launches++;
    }
    CK(cudaGetLastError());
    return PF_OK;
}

// D[c][J][I] = NTT_I(c1_J mod q_I) for `ncts` ciphertexts (c1 at d_cts + c*ct_stride + L*N), and the
// zero-coefficient flags that select the exact path (pf_keyswitch.cuh).
int hoist_digits(pf_engine *e, const u64 *d_cts, size_t ncts, size_t ct_stride) {
    const int L = e->L, N = e->N, k = e->k;
    const size_t per_d = (size_t)L * (L + 1) * N;
    CK(e->s_hoistD.ensure_grow(ncts * per_d * 8));
    CK(e->s_flags.ensure_grow(std::max<size_t>(4, ncts * sizeof(int))));
    CK(cudaMemsetAsync(e->s_flags.p, 0, ncts * sizeof(int), e->stream));
    for (size_t off = 0; off < ncts; off += 16384) {
        const size_t cnt = std::min<size_t>(16384, ncts - off);
        NttParams p{};
        p.in = d_cts + off * ct_stride + (size_t)L * N;
        p.in_sy = N;
        p.in_sz = (long long)ct_stride;
        p.out = e->s_hoistD.as<u64>() + off * per_d;
        p.out_sx = N;
        p.out_sy = (long long)(L + 1) * N;
        p.out_sz = (long long)per_d;
        p.zero_flags = e->s_flags.as<int>() + off;
        p.out_split = 1;
        for (int I = 0; I <= L; I++) p.mod_map[I] = (I == L) ? k - 1 : I;
        launch_ntt(e, NTT_IN_REDUCE, false, p, dim3(L + 1, L, (unsigned)cnt));
    }
    return PF_OK;
}

const GaloisKey *find_key(pf_engine *e, int step) {
    auto it = e->gkeys.find(galois_elt_from_step(e, step));
    return it == e->gkeys.end() ? nullptr : &it->second;
}

void split_convert_chunks(pf_engine *e, const u64 *in, u64 *out, size_t chunk_words, size_t nchunks, size_t in_stride,
                          size_t out_stride, bool to_split) {
    for (size_t c0 = 0; c0 < nchunks; c0 += 32768) {
        const unsigned ny = (unsigned)std::min<size_t>(32768, nchunks - c0);
        split_convert_kernel<<<dim3((unsigned)((chunk_words + 255) / 256), ny), 256, 0, e->stream>>>(
            in + c0 * in_stride, out + c0 * out_stride, chunk_words, in_stride, out_stride, e->d_mods.as<DevModulus>(),
            e->L, e->N, to_split ? 1 : 0);
        e->launches++;
    }
}

void split_convert(pf_engine *e, const u64 *in, u64 *out, size_t nwords, bool to_split) {
    // whole [..][L][N] arrays: chunk per polynomial group keeps blockIdx.x small
    const size_t chunk = (size_t)e->L * e->N;
    split_convert_chunks(e, in, out, chunk, nwords / chunk, chunk, chunk, to_split);
}

// rot[nq][K][2][L][N] (NTT form; split operand format unless the engine is in wide mode or
// `canonical_out`) from d_cts[nq][m][2][L][N] (coefficient form, device)
int build_rotated_sets(pf_engine *e, const u64 *d_cts, size_t nq, u64 *rot, int force_chain, bool canonical_out) {
    const int L = e->L, N = e->N;
    const size_t ctw = (size_t)2 * L * N, m = e->m, R = e->R, K = e->K;
    const bool want_split = !e->mac_wide && !canonical_out;
    bool direct = !force_chain;
    for (size_t r = 1; r < R && direct; r++) direct = find_key(e, (int)r) != nullptr;
    if (R > 1 && !direct && !find_key(e, 1))
        return e->fail(PF_ERR_STATE, "no usable Galois keys: need steps 1..%u or step 1", (unsigned)(R - 1));
    // NTT of the input ciphertexts.  Direct mode with split output keeps a canonical copy (the rotation
    // jobs read c0 from it) and writes the r = 0 members split; otherwise they go straight into rot.
    const bool side_copy = want_split && (direct || R == 1);
    u64 *ntt_dst = rot;
    size_t dst_stride = R * ctw;
    if (side_copy) {
        CK(e->s_cqntt.ensure_grow(nq * m * ctw * 8));
        ntt_dst = e->s_cqntt.as<u64>();
        dst_stride = ctw;
    }
    {
        PhaseTimer pt(e, PF_T_TONTT);
        for (size_t off = 0; off < nq * m; off += 16384) {
            const size_t cnt = std::min<size_t>(16384, nq * m - off);
            NttParams p{};
            p.in = d_cts + off * ctw;
            p.out = ntt_dst + off * dst_stride; // ciphertext (i, a) -> rot member (i*K + a*R) = (i*m + a)*R
            p.in_sx = p.out_sx = N;
            p.in_sy = p.out_sy = (long long)L * N; // y = polynomial 0 / 1
            p.in_sz = (long long)ctw;
            p.out_sz = (long long)dst_stride;
            for (int i = 0; i < L; i++) p.mod_map[i] = i;
            launch_ntt(e, NTT_IN_PLAIN, false, p, dim3(L, 2, (unsigned)cnt));
        }
        if (side_copy) {
            split_convert_chunks(e, ntt_dst, rot, ctw, nq * m, ctw, R * ctw, true);
        }
    }
    if (R == 1) return PF_OK;
    PhaseTimer pt(e, PF_T_ROTATE);
    std::vector<RotJob> jobs;
    if (direct) {
        int hrc;
        {
            HostTick ht("hoist_digits");
            hrc = hoist_digits(e, d_cts, nq * m, ctw);
        }
        if (hrc) return hrc;
        const size_t per_d = (size_t)L * (L + 1) * N;
        jobs.reserve(nq * m * (R - 1));
        // Order: groups of QG queries, rotation-major inside a group.  Consecutive jobs share the Galois
        // key (ks_accumulate keeps it in registers across KS_QT jobs), and the hoisted digits D of a group
        // (QG*m x 1.3 MB at N=8192) stay in L2 across its R-1 rotations; with rotation-major order over the
        // whole batch, D (84 MB for 64 queries) was evicted by the streamed W / output between two uses
        // and re-read from HBM for every rotation (ncu: 1.8 GB read per launch instead of 0.6 GB).
        static const size_t env_qg = getenv("PF_KS_QGROUP") ? (size_t)atoi(getenv("PF_KS_QGROUP")) : 0;
        const size_t QG = env_qg ? env_qg : 16;
        for (size_t i0 = 0; i0 < nq; i0 += QG)
          for (size_t r = 1; r < R; r++)
            for (size_t i = i0; i < std::min(nq, i0 + QG); i++)
                for (size_t a = 0; a < m; a++) {
                    const GaloisKey *gk = find_key(e, (int)r);
                    RotJob j{};
                    j.c1_coef = d_cts + (i * m + a) * ctw + (size_t)L * N;
                    j.c0_ntt = side_copy ? ntt_dst + (i * m + a) * ctw : rot + (i * K + a * R) * ctw;
                    j.key = gk->key.as<u64>();
                    j.perm = gk->perm.as<u32>();
                    j.out = rot + (i * K + a * R + r) * ctw;
                    j.einv = gk->einv;
                    j.D = e->s_hoistD.as<u64>() + (i * m + a) * per_d;
                    j.KM = gk->km.as<u64>();
                    j.flag = e->s_flags.as<int>() + (i * m + a);
                    jobs.push_back(j);
                }
        HostTick ht("run_rot_jobs");
        return run_rot_jobs(e, jobs, want_split);
    }
    // chain: rot_r = rotate(rot_{r-1}, 1); needs c1 of the previous member in coefficient form
    const GaloisKey *gk = find_key(e, 1);
    CK(e->s_c1coef.ensure_grow(nq * m * (size_t)L * N * 8));
    u64 *c1c = e->s_c1coef.as<u64>();
    for (size_t r = 1; r < R; r++) {
        jobs.clear();
        for (size_t i = 0; i < nq; i++)
            for (size_t a = 0; a < m; a++) {
                RotJob j{};
                const u64 *prev = rot + (i * K + a * R + r - 1) * ctw;
                if (r == 1) {
                    j.c1_coef = d_cts + (i * m + a) * ctw + (size_t)L * N;
                } else {
                    u64 *dst = c1c + (i * m + a) * (size_t)L * N;
                    ntt_limbs(e, prev + (size_t)L * N, dst, 1, true);
                    j.c1_coef = dst;
                }
                j.c0_ntt = prev;
                j.key = gk->key.as<u64>();
                j.perm = gk->perm.as<u32>();
                j.out = rot + (i * K + a * R + r) * ctw;
                j.einv = gk->einv;
                jobs.push_back(j);
            }
        int rc = run_rot_jobs(e, jobs, false);
        if (rc) return rc;
    }
    if (want_split) split_convert(e, rot, rot, nq * K * ctw, true);
    return PF_OK;
}

// ---- MAC launch -----------------------------------------------------------------------------
template <int T, int UNROLL, bool WIDE, bool FPRED = false>
void launch_mac_t(pf_engine *e, const MacParams &p, unsigned nchunks) {
    const size_t smem = (size_t)p.K * 2 * T * 8;
    auto kern = mac_kernel<T, UNROLL, WIDE, FPRED>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<dim3(nchunks, p.L * (p.N / T)), 256, smem, e->stream>>>(p);
    e->launches++;
}

template <int T, int UNROLL, int MINCTA, bool FPRED>
void launch_mac_occ_k(pf_engine *e, const MacParams &p, unsigned nchunks) {
    const size_t smem = (size_t)p.K * 2 * T * 8;
    auto kern = mac_kernel_occ<T, UNROLL, FPRED, MINCTA>;
    // the attribute is a MAXIMUM: raise it when a launch needs more than any launch of this instantiation before
    // (setting it per launch cost a few microseconds each; lowering it makes a later, larger launch invalid)
    static size_t attr_max = 48 * 1024;
    if (smem > attr_max) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_max = smem;
    }
    kern<<<dim3(nchunks, p.L * (p.N / T)), 256, smem, e->stream>>>(p);
    e->launches++;
}

template <int T, int UNROLL, int MINCTA = 3>
void launch_mac_occ_t(pf_engine *e, const MacParams &p, unsigned nchunks) {
    if (e->mac_fpred) launch_mac_occ_k<T, UNROLL, MINCTA, true>(e, p, nchunks);
    else launch_mac_occ_k<T, UNROLL, MINCTA, false>(e, p, nchunks);
}

// MAC variant (PF_MAC_VARIANT, read per call so tests can flip it): 0 = two blocks per lane, 2 CTAs/SM
// (mac_kernel), 4 / 5 = one block per lane, 256-coefficient slices, 3 CTAs/SM (mac_kernel_occ, 2 / 4
// diagonals per load group), 6 = the same with 128-coefficient slices, 64 registers and 4 CTAs/SM.
// Default: 6 whenever the slices fit an SM's shared memory (K <= 16), else 0.  Measured at K = 16:
// 1.005 / 0.89 / 0.835 ms for 0 / 4 / 6.  (cp.async and TMA bulk-copy rings were measured slower — 1.5 /
// 1.22 ms — and removed; profiles/README.md keeps the numbers.)
int mac_variant(const pf_engine *e) {
    const char *v = getenv("PF_MAC_VARIANT");
    const int env = v ? atoi(v) : -1;
    if (e->mac_wide || e->K < 4) return 0;
    if (env == 0) return 0;
    if (e->K > 16) { // 7 = the one-block-per-lane kernel with the widest slice of which three fit an SM (K = 128: 32 coefficients)
        const bool fits7 = (size_t)e->K * 2 * 32 * 8 * 3 <= (size_t)227 * 1024;
        return fits7 ? 7 : 0;
    }
    const bool occ_fits = (size_t)e->K * 2 * 256 * 8 * 3 <= (size_t)227 * 1024;
    if (env >= 4 && env <= 6) return occ_fits ? env : 0;
    return occ_fits ? 6 : PF_MAC_DEFAULT_VARIANT;
}

int mac_tile7(const pf_engine *e) { // variant 7: slice width
    for (int T : {128, 64, 32})
        if ((size_t)e->K * 2 * T * 8 * 3 <= (size_t)227 * 1024) return T;
    return 32;
}

int mac_tile(const pf_engine *e) {
    const int v = mac_variant(e);
    if (v == 6) return 128;
    if (v == 7) return mac_tile7(e);
    if (v == 4 || v == 5) return 256;
    const int env_t = getenv("PF_MAC_TILE") ? atoi(getenv("PF_MAC_TILE")) : 0;
    if (env_t == 256 || env_t == 128 || env_t == 64) return env_t;
    return e->K <= 32 ? 256 : (e->K <= 64 ? 128 : 64);
}

template <int T>
void launch_mac_tile(pf_engine *e, const MacParams &p, unsigned nchunks) {
    if (e->mac_wide) launch_mac_t<T, 1, true>(e, p, nchunks);
    else if (p.K >= 4 && e->mac_fpred) launch_mac_t<T, 2, false, true>(e, p, nchunks);
    else if (p.K >= 4) launch_mac_t<T, 2, false>(e, p, nchunks);
    else if (p.K == 2) launch_mac_t<T, 1, false>(e, p, nchunks);
    else launch_mac_t<T, 1, false>(e, p, nchunks);
}

void launch_mac(pf_engine *e, const MacParams &p, unsigned nchunks) {
    const int v = mac_variant(e);
    if (v == 4) return launch_mac_occ_t<256, 2>(e, p, nchunks);
    if (v == 5) return (p.K % 8 == 0) ? launch_mac_occ_t<256, 4>(e, p, nchunks) : launch_mac_occ_t<256, 2>(e, p, nchunks);
    if (v == 6) return launch_mac_occ_t<128, 2, 4>(e, p, nchunks);
    if (v == 7) {
        const int T7 = mac_tile7(e);
        if (T7 == 128) return launch_mac_occ_t<128, 2, 3>(e, p, nchunks);
        if (T7 == 64) return launch_mac_occ_t<64, 2, 3>(e, p, nchunks);
        return launch_mac_occ_t<32, 2, 3>(e, p, nchunks);
    }
    const int T = mac_tile(e);
    if (T == 256) launch_mac_tile<256>(e, p, nchunks);
    else if (T == 128) launch_mac_tile<128>(e, p, nchunks);
    else launch_mac_tile<64>(e, p, nchunks);
}

// Build the (query, block) pair list for this rank and the chunk table.  Returns PF_OK.
struct PairPlan {
    std::vector<long long> pair_block; // processing order: per query sorted by block (L2 reuse across queries)
    std::vector<int> pair_out;         // result slot (response order) of every processed pair
    std::vector<MacChunk> chunks;
    std::vector<uint64_t> results_per_query;
    uint64_t useful = 0;
};

int plan_pairs(pf_engine *e, uint64_t nq, const int64_t *idx, uint32_t nprobe, PairPlan &pl) {
    pl.results_per_query.assign(nq, 0);
    std::vector<int> pair_query;
    for (uint64_t i = 0; i < nq; i++)
        for (uint32_t p = 0; p < nprobe; p++) {
            const int64_t l = idx[i * nprobe + p];
            if (l < 0 || (uint64_t)l >= e->nlist)
                return e->fail(PF_ERR_INVALID, "list id %lld of query %llu out of range [0,%llu)", (long long)l,
                               (unsigned long long)i, (unsigned long long)e->nlist);
            for (long long b = e->list_block_start[l]; b < e->list_block_start[l + 1]; b++) {
                pl.pair_block.push_back(b);
                pair_query.push_back((int)i);
                pl.results_per_query[i]++;
                pl.useful += e->blocks[b].nvec;
            }
        }
    const size_t P = pl.pair_block.size();
    // Within a query, process blocks in ascending block order: CTAs of different queries that run
    // concurrently (same limb / slice) then reach a shared block at about the same time, so its
    // second fetch hits L2.  The response order (probe order) is kept through pair_out.
    pl.pair_out.resize(P);
    {
        size_t a = 0;
        std::vector<std::pair<long long, int>> tmp;
        while (a < P) {
            size_t b = a;
            while (b < P && pair_query[b] == pair_query[a]) b++;
            tmp.clear();
            for (size_t i = a; i < b; i++) tmp.push_back({pl.pair_block[i], (int)i});
            std::sort(tmp.begin(), tmp.end());
            for (size_t i = a; i < b; i++) {
                pl.pair_block[i] = tmp[i - a].first;
                pl.pair_out[i] = tmp[i - a].second;
            }
            a = b;
        }
    }
    const int T = mac_tile(e);
    const size_t ctas_per_chunk = (size_t)e->L * (e->N / T);
    const size_t want_chunks = (4 * 148 + ctas_per_chunk - 1) / ctas_per_chunk;
    size_t CH = std::max<size_t>(4, std::min<size_t>(64, P / std::max<size_t>(1, want_chunks)));
    static const size_t env_ch = getenv("PF_MAC_CHUNK") ? (size_t)atoi(getenv("PF_MAC_CHUNK")) : 0;
    if (env_ch) CH = env_ch;
    size_t s = 0;
    while (s < P) {
        size_t epos = s;
        while (epos < P && pair_query[epos] == pair_query[s] && epos - s < CH) epos++;
        pl.chunks.push_back(MacChunk{pair_query[s], (int)s, (int)(epos - s), 0});
        s = epos;
    }
    return PF_OK;
}

// chunk table and pair->block map of a whole plan to the device (pageable sources: the copies are
// complete on return)
int upload_plan(pf_engine *e, const PairPlan &pl) {
    const size_t P = pl.pair_block.size();
    if (!P) return PF_OK;
    CK(e->s_chunks.ensure_grow(pl.chunks.size() * sizeof(MacChunk)));
    CK(e->s_pairblock.ensure_grow(P * sizeof(long long)));
    CK(upload_async(e, e->s_chunks.p, pl.chunks.data(), pl.chunks.size() * sizeof(MacChunk)));
    CK(upload_async(e, e->s_pairblock.p, pl.pair_block.data(), P * sizeof(long long)));
    CK(e->s_pairout.ensure_grow(P * sizeof(int)));
    CK(upload_async(e, e->s_pairout.p, pl.pair_out.data(), P * sizeof(int)));
    return PF_OK;
}

// The device-resident step for queries [q0, q0+nq) of a plan: rotations, MAC, inverse NTT.  Result of
// pair i lands at d_out + i*out_stride (coefficient form); d_cts points at query q0's ciphertexts.
int search_core(pf_engine *e, uint64_t q0, uint64_t nq, const u64 *d_cts, const PairPlan &pl, u64 *d_out,
                size_t out_stride) {
    const int L = e->L, N = e->N;
    const size_t ctw = (size_t)2 * L * N;
    if (!nq) return PF_OK; // an empty query range (a rank whose query group is empty): nothing to launch
    CK(e->s_rot.ensure_grow(nq * e->K * ctw * 8));
    u64 *rot = e->s_rot.as<u64>();
    int rc;
    {
        HostTick ht("rotated_sets");
        rc = build_rotated_sets(e, d_cts, nq, rot, 0, false);
    }
    if (rc) return rc;
    HostTick ht2("mac+intt");
    // chunks / pairs of this query range (chunks are ordered by query)
    size_t cb = 0, ce = 0;
    while (cb < pl.chunks.size() && (uint64_t)pl.chunks[cb].query < q0) cb++;
    ce = cb;
    while (ce < pl.chunks.size() && (uint64_t)pl.chunks[ce].query < q0 + nq) ce++;
    if (ce == cb) return PF_OK;
    // Sub-batches of whole queries: the full-level results of a sub-batch live in scratch until the
    // mod-switch has written the response slots, and that scratch is capped (a 256-query batch at
    // N = 16384 would otherwise hold 35 GiB of it).  One sub-batch in every configuration up to 64 queries.
    const size_t cap_bytes = getenv("PF_FULL_SCRATCH_MB") ? (size_t)atoll(getenv("PF_FULL_SCRATCH_MB")) << 20 : (size_t)6 << 30; // per call: tests flip it
    const size_t maxP = std::max<size_t>(1, cap_bytes / (ctw * 8));
    for (size_t c0 = cb; c0 < ce;) {
    size_t c1 = c0, np = 0;
    while (c1 < ce) { // take whole queries while they fit (always at least one)
        size_t cq = c1, nqp = 0;
        while (cq < ce && pl.chunks[cq].query == pl.chunks[c1].query) nqp += (size_t)pl.chunks[cq++].pair_count;
        if (np && e->Lr < L && np + nqp > maxP) break;
        np += nqp;
        c1 = cq;
    }
    const size_t p0 = (size_t)pl.chunks[c0].pair_start;
    const size_t p1 = (size_t)pl.chunks[c1 - 1].pair_start + pl.chunks[c1 - 1].pair_count;
    const size_t P = p1 - p0;
    // with result_limbs < L the full-level results live in scratch and the mod-switch writes the slots
    const bool ms = e->Lr < L;
    u64 *full = d_out;
    size_t full_stride = out_stride, full_base = p0;
    if (ms) {
        CK(e->s_full.ensure_grow(P * ctw * 8));
        full = e->s_full.as<u64>();
        full_stride = ctw;
        full_base = 0;
    }
    {
        PhaseTimer pt(e, PF_T_MAC);
        MacParams mp{};
        mp.rot = rot;
        mp.diag = e->d_diag.as<u64>();
        mp.norm = e->d_norm.as<u64>();
        mp.diag_sb = (long long)e->diag_block_words;
        mp.diag_sk = (long long)L * N;
        mp.norm_sb = (long long)e->norm_block_words;
        mp.chunks = e->s_chunks.as<MacChunk>() + c0; // absolute pair / query indices (upload_plan)
        mp.pair_block = e->s_pairblock.as<long long>();
        mp.pair_out = e->s_pairout.as<int>();
        mp.out = ms ? full - p0 * full_stride : d_out; // slot index is absolute
        mp.out_stride = (long long)full_stride;
        mp.mods = e->d_mods.as<DevModulus>();
        mp.K = e->K;
        mp.L = L;
        mp.N = N;
        mp.query_base = (int)q0;
        static const int env_hint = getenv("PF_MAC_L2HINT") ? atoi(getenv("PF_MAC_L2HINT")) : 0;
        mp.l2hint = env_hint;
        launch_mac(e, mp, (unsigned)(c1 - c0));
    }
    {
        PhaseTimer pt(e, PF_T_INTT);
        u64 *base = full + full_base * full_stride;
        // (an inverse NTT with the mod-switch folded into its store was bit-exact but 2-6x slower than the
        // pair INTT + modswitch kernel — register pressure in the 32-point pass — and was removed)
        for (size_t off = 0; off < P; off += 32768) {
            const unsigned cnt = (unsigned)std::min<size_t>(32768, P - off);
            NttParams ip{};
            ip.in = ip.out = base + off * full_stride;
            ip.in_sx = ip.out_sx = N;
            ip.in_sy = ip.out_sy = (long long)L * N;
            ip.in_sz = ip.out_sz = (long long)full_stride;
            for (int i = 0; i < L; i++) ip.mod_map[i] = i;
            {
                launch_ntt(e, NTT_IN_PLAIN, true, ip, dim3(L, 2, cnt));
                if (ms) {
                    const u64 *src = base + off * full_stride;
                    u64 *dst = d_out + (p0 + off) * out_stride;
                    const DevModulus *dm = e->d_mods.as<DevModulus>();
                    const u64 *tab = e->d_mstab.as<u64>();
                    const double *tabf = e->d_mstab_fp.as<double>();
                    const dim3 g2(N / 512, 2, cnt);
                    // primes <= 49 bits: the FP64-pipe kernel (pf_ntt_fp.cuh); otherwise the integer one
                    const bool ms_int = getenv("PF_MS_INT") != nullptr; // per call: tests flip it
                    const bool ms_fp = e->max_prime_bits <= 49 && !ms_int;
#define MS_CASE(LL, RR)                                                                                            \
    else if (L == LL && e->Lr == RR && ms_fp) modswitch_fp_kernel_t<LL, RR><<<g2, 256, 0, e->stream>>>(src, full_stride, dst, \
                                                                                                      out_stride, dm, tabf, N); \
    else if (L == LL && e->Lr == RR) modswitch_kernel_t<LL, RR><<<g2, 256, 0, e->stream>>>(src, full_stride, dst,     \
                                                                                          out_stride, dm, tab, N)
                    if (false) {
                    }
                    MS_CASE(4, 3);
                    MS_CASE(4, 2);
                    MS_CASE(4, 1);
                    MS_CASE(8, 4);
                    MS_CASE(8, 2);
                    MS_CASE(8, 1);
                    MS_CASE(3, 2);
                    MS_CASE(3, 1);
                    else modswitch_kernel<<<dim3(N / 256, 2, cnt), 256, 0, e->stream>>>(src, full_stride, dst, out_stride,
                                                                                        dm, tab, L, e->Lr, N);
#undef MS_CASE
                    e->launches++;
                }
            }
        }
    }
    c0 = c1;
    } // sub-batches
    CK(cudaGetLastError());
    return PF_OK;
}

void write_seal_header(uint8_t *p, uint64_t total) {
    p[0] = 0x5E;
    p[1] = 0xA1;
    p[2] = 0x10;
    p[3] = 4;
    p[4] = 1;
    p[5] = 0;
    p[6] = p[7] = 0;
    memcpy(p + 8, &total, 8);
}

// SEAL Ciphertext::save_members, compr_mode none (ciphertext.cpp, dynarray.h)
void write_ct_prefix(const pf_engine *e, uint8_t *p, int is_ntt, const uint64_t parms_id[4], int limbs) {
    const uint64_t words = (uint64_t)2 * limbs * e->N, total = SEAL_CT_HEADER + words * 8;
    write_seal_header(p, total);
    p += 16;
    memcpy(p, parms_id, 32);
    p += 32;
    *p++ = (uint8_t)(is_ntt ? 1 : 0);
    uint64_t v = 2;
    memcpy(p, &v, 8);
    p += 8;
    v = (uint64_t)e->N;
    memcpy(p, &v, 8);
    p += 8;
    v = (uint64_t)limbs;
    memcpy(p, &v, 8);
    p += 8;
    const double scale = 1.0;
    memcpy(p, &scale, 8);
    p += 8;
    v = 1;
    memcpy(p, &v, 8);
    p += 8;
    write_seal_header(p, 16 + 8 + words * 8);
    p += 16;
    memcpy(p, &words, 8);
}

// returns 0 on success; data_off = offset of the words, nlimbs = coeff_modulus_size
// SEAL streams saved with compr_mode_type::zlib or ::zstd: the 16-byte SEALHeader is followed by one zlib
// (RFC 1950) stream / one Zstandard frame holding what an uncompressed save would have written after its
// header; nested objects inside are saved uncompressed [EXT: SEAL 4.1 serialization.cpp / util/ztools.cpp].
// Inflates into `out` as the equivalent compr_mode none stream.  Returns 0, 1 (compr_mode none, untouched),
// <0 error (-3: unknown mode, or zstd without libzstd.so.1 on this host).
// `max_out` bounds the inflated size (header included): the streams come from clients, and a few KB of
// deflate can expand a thousandfold — callers pass the size the object can legitimately have (-7 beyond
// it).  Input and output are fed to zlib in chunks below its 32-bit avail_in / avail_out.
// compr_mode_type::zstd (SEAL's default when built with it): the body is one Zstandard frame written by
// ZSTD_compressStream2 [EXT: SEAL 4.1 util/ztools.cpp zstd_deflate_array_inplace].  The image ships the
// runtime library without its header, so the five stable entry points of the streaming decoder (zstd.h,
// stable since v1.3) are declared here and bound with dlopen on first use; without the library a zstd
// stream is refused (-3) like any other unsupported mode.
struct ZstdInBuffer {
    const void *src;
    size_t size, pos;
};
struct ZstdOutBuffer {
    void *dst;
    size_t size, pos;
};
struct ZstdApi {
    void *(*create)() = nullptr;
    size_t (*release)(void *) = nullptr;
    size_t (*init)(void *) = nullptr;
    size_t (*run)(void *, ZstdOutBuffer *, ZstdInBuffer *) = nullptr;
    unsigned (*is_error)(size_t) = nullptr;
    bool ok = false;
    ZstdApi() {
        const char *forced = getenv("PF_ZSTD_LIB"); // tests: point at a missing file to exercise the refusal
        void *lib = forced ? dlopen(forced, RTLD_NOW | RTLD_LOCAL) : nullptr;
        if (!forced)
            for (const char *name : {"libzstd.so.1", "libzstd.so"})
                if ((lib = dlopen(name, RTLD_NOW | RTLD_LOCAL))) break;
        if (!lib) return;
        create = reinterpret_cast<void *(*)()>(dlsym(lib, "ZSTD_createDStream"));
        release = reinterpret_cast<size_t (*)(void *)>(dlsym(lib, "ZSTD_freeDStream"));
        init = reinterpret_cast<size_t (*)(void *)>(dlsym(lib, "ZSTD_initDStream"));
        run = reinterpret_cast<size_t (*)(void *, ZstdOutBuffer *, ZstdInBuffer *)>(dlsym(lib, "ZSTD_decompressStream"));
        is_error = reinterpret_cast<unsigned (*)(size_t)>(dlsym(lib, "ZSTD_isError"));
        ok = create && release && init && run && is_error;
    }
};
const ZstdApi &zstd_api() {
    static const ZstdApi api;
    return api;
}

// body of a compr_mode zstd stream -> out (which already holds the 16 header bytes); same contract and
// the same output ceiling as the zlib loop below
int inflate_zstd_body(const uint8_t *body, size_t body_len, std::vector<uint8_t> &out, size_t max_out) {
    const ZstdApi &z = zstd_api();
    if (!z.ok) return -3;
    struct DGuard {
        const ZstdApi &z;
        void *ds;
        ~DGuard() {
            if (ds) z.release(ds);
        }
    } g{z, z.create()};
    if (!g.ds || z.is_error(z.init(g.ds))) return -5;
    ZstdInBuffer in{body, body_len, 0};
    constexpr size_t ZCHUNK = (size_t)1 << 30;
    for (;;) {
        const size_t have = out.size();
        uint8_t probe = 0;
        const bool at_cap = have >= max_out; // the object cannot be larger: one probe byte tells "ended" from "more"
        const size_t grow = at_cap ? 1 : std::min(std::min(max_out - have, ZCHUNK), std::max<size_t>(1 << 16, body_len * 4));
        if (!at_cap) out.resize(have + grow);
        ZstdOutBuffer o{at_cap ? static_cast<void *>(&probe) : static_cast<void *>(out.data() + have), grow, 0};
        const size_t r = z.run(g.ds, &o, &in);
        if (z.is_error(r)) return -6;
        if (at_cap && o.pos) return -7;
        if (!at_cap) out.resize(have + o.pos);
        if (r == 0) return 0;                                  // the frame is decoded and flushed
        if (in.pos == in.size && o.pos < o.size) return -1;    // truncated: input exhausted, output not full
    }
}

int inflate_seal_stream(const uint8_t *p, size_t len, std::vector<uint8_t> &out, size_t *consumed, size_t max_out) {
    if (len < 16 || p[0] != 0x5E || p[1] != 0xA1) return -2;
    if (p[5] == 0) return 1;
    if (p[5] != 1 && p[5] != 2) return -3; // compr_mode_type: 0 none, 1 zlib, 2 zstd
    uint64_t total;
    memcpy(&total, p + 8, 8);
    if (total > len || total < 16) return -1;
    if (max_out < 16) return -7;
    out.assign(16, 0);
    auto finish = [&]() {
        memcpy(out.data(), p, 16);
        out[5] = 0;
        const uint64_t new_total = out.size();
        memcpy(out.data() + 8, &new_total, 8);
        if (consumed) *consumed = (size_t)total;
        return 0;
    };
    if (p[5] == 2) {
        const int zr = inflate_zstd_body(p + 16, (size_t)(total - 16), out, max_out);
        return zr ? zr : finish();
    }
    struct ZGuard {
        z_stream zs{};
        bool live = false;
        ~ZGuard() {
            if (live) inflateEnd(&zs);
        }
    } zg;
    z_stream &zs = zg.zs;
    if (inflateInit(&zs) != Z_OK) return -5;
    zg.live = true;
    const uint8_t *in_pos = p + 16;
    size_t in_left = (size_t)(total - 16);
    constexpr size_t ZCHUNK = (size_t)1 << 30;
    for (;;) {
        if (zs.avail_in == 0 && in_left) {
            const size_t take = std::min(in_left, ZCHUNK);
            zs.next_in = const_cast<Bytef *>(in_pos);
            zs.avail_in = (uInt)take;
            in_pos += take;
            in_left -= take;
        }
        const size_t have = out.size();
        Bytef probe = 0;
        const bool at_cap = have >= max_out; // the object cannot be larger: one probe byte tells "ended" from "more"
        const size_t grow = at_cap ? 1 : std::min(std::min(max_out - have, ZCHUNK), std::max<size_t>(1 << 16, (size_t)(total - 16) * 2));
        if (!at_cap) out.resize(have + grow);
        zs.next_out = at_cap ? &probe : out.data() + have;
        zs.avail_out = (uInt)grow;
        const int zr = inflate(&zs, Z_NO_FLUSH);
        const size_t produced = grow - zs.avail_out;
        if (at_cap && produced) return -7;
        if (!at_cap) out.resize(have + produced);
        if (zr == Z_STREAM_END) break;
        if (zr != Z_OK && zr != Z_BUF_ERROR) return -6;
        if (zs.avail_in == 0 && in_left == 0 && zs.avail_out != 0) return -1; // truncated
    }
    return finish();
}

int parse_ct_prefix(const pf_engine *e, const uint8_t *p, size_t len, int *is_ntt, uint64_t parms_id[4],
                    uint64_t *nlimbs, size_t *total_out) {
    if (len < SEAL_CT_HEADER) return -1;
    if (p[0] != 0x5E || p[1] != 0xA1 || p[2] != 0x10 || p[3] != 4) return -2;
    if (p[5] != 0) return -3; // compressed streams are inflated by the callers first
    uint64_t total, size, n, cms, words;
    memcpy(&total, p + 8, 8);
    if (total > len) return -1;
    memcpy(parms_id, p + 16, 32);
    *is_ntt = p[48] ? 1 : 0;
    memcpy(&size, p + 49, 8);
    memcpy(&n, p + 57, 8);
    memcpy(&cms, p + 65, 8);
    const uint8_t *in = p + 89;
    if (in[0] != 0x5E || in[1] != 0xA1 || in[5] != 0) return -2;
    memcpy(&words, in + 16, 8);
    if (size != 2 || n != (uint64_t)e->N || cms < 1 || cms > PF_MAX_PRIMES) return -4; // bounded before multiplying
    if (words != size * n * cms) return -4;
    if (SEAL_CT_HEADER + words * 8 != total) return -4;
    *nlimbs = cms;
    *total_out = (size_t)total;
    return 0;
}

// SEAL seeded ciphertext (Serializable<Ciphertext> of a symmetric-key encryption): Ciphertext::save_members
// writes the members, the DynArray of the FIRST polynomial only (L*N words) and then the
// UniformRandomGeneratorInfo of the PRNG that drew c1 as a nested stream {SEALHeader, uint8 type, 64-byte seed}
// [EXT: SEAL 4.1 ciphertext.cpp save_members / load_members / expand_seed].  `p` is an uncompressed stream.
// Returns 1 when the stream is not seeded (untouched), 0 when `out` holds the equivalent full stream (c1 =
// sample_poly_uniform of a Blake2xb PRNG with that seed, coefficient form), < 0 on malformed input
// (-8: a PRNG type other than blake2xb, e.g. shake256).
constexpr size_t SEAL_PRNG_INFO_BYTES = 16 + 1 + 64;
int expand_seeded_stream(const uint8_t *p, size_t len, uint64_t N, const uint64_t *primes, uint32_t L, std::vector<uint8_t> &out) {
    if (len < SEAL_CT_HEADER) return -1;
    if (p[0] != 0x5E || p[1] != 0xA1 || p[2] != 0x10 || p[3] != 4 || p[5] != 0) return -2;
    uint64_t total, size, n, cms, words;
    memcpy(&total, p + 8, 8);
    if (total > len) return -1;
    memcpy(&size, p + 49, 8);
    memcpy(&n, p + 57, 8);
    memcpy(&cms, p + 65, 8);
    const uint8_t *in = p + 89;
    if (in[0] != 0x5E || in[1] != 0xA1 || in[5] != 0) return -2;
    memcpy(&words, in + 16, 8);
    if (size != 2 || n != N || cms != (uint64_t)L) return -4;
    if (words == 2 * n * cms) return total == SEAL_CT_HEADER + words * 8 ? 1 : -4; // both polynomials present
    if (words != n * cms) return -4;
    const size_t half = (size_t)words * 8;
    if (total != SEAL_CT_HEADER + half + SEAL_PRNG_INFO_BYTES) return -4;
    const uint8_t *info = p + SEAL_CT_HEADER + half;
    uint64_t info_total;
    memcpy(&info_total, info + 8, 8);
    if (info[0] != 0x5E || info[1] != 0xA1 || info[5] != 0 || info_total != SEAL_PRNG_INFO_BYTES) return -4;
    if (info[16] != 1) return -8; // prng_type::blake2xb = 1 (shake256 = 2)
    out.resize(SEAL_CT_HEADER + 2 * half);
    memcpy(out.data(), p, SEAL_CT_HEADER + half);
    const uint64_t new_total = SEAL_CT_HEADER + 2 * half, arr_total = 16 + 8 + 2 * half, new_words = 2 * words;
    memcpy(out.data() + 8, &new_total, 8);
    memcpy(out.data() + 89 + 8, &arr_total, 8);
    memcpy(out.data() + 89 + 16, &new_words, 8);
    pfh::SealBlake2xbPrng prng(info + 17);
    pfh::seal_sample_poly_uniform(prng, primes, L, N, reinterpret_cast<uint64_t *>(out.data() + SEAL_CT_HEADER + half));
    return 0;
}

// An uncompressed seeded stream of exactly the shape expand_seeded_stream expands, drawn with blake2xb: what the
// device-side expansion takes (anything else — compressed, shake256, malformed — stays on the host path, which
// names the problem).
bool is_seeded_raw_stream(const uint8_t *p, size_t len, uint64_t N, uint32_t L) {
    const size_t half = (size_t)N * L * 8;
    if (len != SEAL_CT_HEADER + half + SEAL_PRNG_INFO_BYTES) return false;
    if (p[0] != 0x5E || p[1] != 0xA1 || p[2] != 0x10 || p[3] != 4 || p[5] != 0) return false;
    uint64_t total, size, n, cms, words, info_total;
    memcpy(&total, p + 8, 8);
    memcpy(&size, p + 49, 8);
    memcpy(&n, p + 57, 8);
    memcpy(&cms, p + 65, 8);
    const uint8_t *in = p + 89;
    if (in[0] != 0x5E || in[1] != 0xA1 || in[5] != 0) return false;
    memcpy(&words, in + 16, 8);
    if (total != len || size != 2 || n != N || cms != (uint64_t)L || words != n * cms) return false;
    const uint8_t *info = p + SEAL_CT_HEADER + half;
    memcpy(&info_total, info + 8, 8);
    return info[0] == 0x5E && info[1] == 0xA1 && info[5] == 0 && info_total == SEAL_PRNG_INFO_BYTES && info[16] == 1;
}

// c0 from the uploaded bytes + c1 from the seed, for `nc` ciphertexts whose data offsets sit in d_off: three launches
// on the engine stream (pf_seeded.cuh).  scratch: nc * SEED_SCRATCH_WORDS words, zeroed here.
int expand_seeded_on_device(pf_engine *e, const uint8_t *d_raw, const u64 *d_off, u64 *d_dst, u64 *d_scratch, size_t nc) {
    if (!nc) return PF_OK;
    const size_t half = (size_t)e->L * e->N;
    CK(cudaMemsetAsync(d_scratch, 0, nc * SEED_SCRATCH_WORDS * 8, e->stream));
    strip_seeded_kernel<<<dim3((unsigned)((half + 255) / 256), (unsigned)nc), 256, 0, e->stream>>>(d_raw, d_off, d_dst, half);
    SeededParams sp{d_raw, d_off, d_dst, d_scratch, e->d_mods.as<DevModulus>(), e->L, e->N};
    seeded_expand_kernel<<<dim3((unsigned)(half / 512 + 1), (unsigned)nc), 64, 0, e->stream>>>(sp);
    seeded_fixup_kernel<<<(unsigned)((nc + 63) / 64), 64, 0, e->stream>>>(sp, (int)nc, e->d_err_word);
    e->launches += 3;
    CK(cudaGetLastError());
    return PF_OK;
}

// The request's slow path, for a whole batch: ciphertext c needs host work when its stream is compressed (zlib /
// zstd) or seeded; out[c] receives its full compr_mode none form and stays EMPTY for streams that are already
// that (the fast path uploads those straight from the caller's blob).  `out` is left empty altogether when no
// stream needs work.  Streams are independent: the work is spread over up to `threads` host threads (Blake2xb
// expansion of one N = 8192 ciphertext is ~0.3 ms of one core; a batch of 64 would otherwise cost more than the
// GPU step).  Returns 0, or the error of the lowest failing index (codes of inflate_seal_stream /
// expand_seeded_stream) in *err_index, *err_mode (the stream's compr_mode byte).
int normalize_ct_streams(const uint8_t *blob, const uint64_t *offs, size_t ncts, uint64_t N, const uint64_t *primes, uint32_t L,
                         std::vector<std::vector<uint8_t>> &out, unsigned threads, size_t *err_index, int *err_mode) {
    const size_t full_bytes = SEAL_CT_HEADER + (size_t)2 * L * N * 8;
    std::vector<size_t> work;
    for (size_t c = 0; c < ncts; c++) {
        const uint8_t *src = blob + offs[c];
        const size_t len = (size_t)(offs[c + 1] - offs[c]);
        bool slow = len >= 16 && src[5] != 0;
        if (!slow && len >= SEAL_CT_HEADER) { // seeded: the DynArray holds one polynomial
            uint64_t words;
            memcpy(&words, src + 105, 8);
            slow = words == N * L;
        }
        if (slow) work.push_back(c);
    }
    out.clear();
    if (work.empty()) return 0;
    out.resize(ncts);
    std::vector<int> code(work.size(), 0);
    auto run = [&](size_t first, size_t step) {
        for (size_t w = first; w < work.size(); w += step) {
            const size_t c = work[w];
            const uint8_t *src = blob + offs[c];
            size_t len = (size_t)(offs[c + 1] - offs[c]);
            std::vector<uint8_t> plain, full;
            if (src[5] != 0) {
                const int zr = inflate_seal_stream(src, len, plain, nullptr, full_bytes);
                if (zr) {
                    code[w] = zr;
                    continue;
                }
                src = plain.data();
                len = plain.size();
            }
            const int er = len >= SEAL_CT_HEADER ? expand_seeded_stream(src, len, N, primes, L, full) : 1;
            if (er == -8) {
                code[w] = -8;
                continue;
            }
            if (er == 0) out[c] = std::move(full);
            else if (!plain.empty()) out[c] = std::move(plain); // inflated, not seeded (or malformed: the strict parser names it)
            // else: looked seeded but is not expandable — left to the strict parser on the original bytes
        }
    };
    const size_t nt = std::max<size_t>(1, std::min<size_t>(threads ? threads : 1, work.size()));
    if (nt == 1) {
        run(0, 1);
    } else {
        std::vector<std::thread> pool;
        for (size_t t = 1; t < nt; t++) pool.emplace_back(run, t, nt);
        run(0, nt);
        for (auto &th : pool) th.join();
    }
    for (size_t w = 0; w < work.size(); w++)
        if (code[w]) {
            if (err_index) *err_index = work[w];
            if (err_mode) *err_mode = (int)blob[offs[work[w]] + 5];
            return code[w];
        }
    return 0;
}

unsigned host_threads() {
    static const unsigned n = [] {
        const char *env = getenv("PF_HOST_THREADS");
        unsigned v = env ? (unsigned)atoi(env) : std::thread::hardware_concurrency();
        return std::max(1u, std::min(v ? v : 1u, 16u));
    }();
    return n;
}

// Serializable<GaloisKeys> (what KeyGenerator::create_galois_keys returns without a destination, the usual way
// keys travel): every key ciphertext — NTT form, all k primes — is saved seeded like a symmetric ciphertext, c1
// replaced by the seed it expands from [EXT: SEAL 4.1 keygenerator.cpp generate_one_kswitch_key with save_seed,
// kswitchkeys.h save_members].  `p` is an uncompressed KSwitchKeys stream; returns 1 when no key inside is seeded
// (untouched), 0 when `out` holds the equivalent full stream, < 0 on malformed input (-8: a PRNG other than
// blake2xb).  Output is bounded by twice the input.
int expand_galois_keys_stream(const uint8_t *p, size_t len, uint64_t N, const uint64_t *primes, uint32_t k, std::vector<uint8_t> &out) {
    if (len < 16 + 32 + 8 || p[0] != 0x5E || p[1] != 0xA1 || p[5] != 0) return -2;
    uint64_t total, dim1;
    memcpy(&total, p + 8, 8);
    if (total > len || total < 16 + 32 + 8) return -1;
    memcpy(&dim1, p + 48, 8);
    if (dim1 > N) return -4;
    out.assign(p, p + 56);
    size_t off = 56;
    bool any = false;
    std::vector<uint8_t> full;
    for (uint64_t index = 0; index < dim1; index++) {
        if (off + 8 > total) return -1;
        uint64_t dim2;
        memcpy(&dim2, p + off, 8);
        out.insert(out.end(), p + off, p + off + 8);
        off += 8;
        if (dim2 > PF_MAX_PRIMES) return -4;
        for (uint64_t j = 0; j < dim2; j++) {
            if (off + 16 > total) return -1;
            uint64_t ctotal;
            memcpy(&ctotal, p + off + 8, 8);
            if (ctotal < SEAL_CT_HEADER || ctotal > total - off) return -1;
            const int er = expand_seeded_stream(p + off, (size_t)ctotal, N, primes, k, full);
            if (er == -8) return -8;
            if (er == 0) {
                any = true;
                out.insert(out.end(), full.begin(), full.end());
            } else { // full already, or something the strict parser of the caller will name
                out.insert(out.end(), p + off, p + off + ctotal);
            }
            off += (size_t)ctotal;
        }
    }
    if (!any) return 1;
    const uint64_t new_total = out.size();
    memcpy(out.data() + 8, &new_total, 8);
    return 0;
}

} // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int pf_abi_version(void) { return PF_ABI_VERSION; }

int pf_seal_stream_inflate(const uint8_t *in, size_t len, uint8_t *out, size_t cap, size_t *written, size_t *consumed) {
    if (!in || !written) return PF_ERR_INVALID;
    if (len < 16 || in[0] != 0x5E || in[1] != 0xA1) return PF_ERR_FORMAT;
    uint64_t total;
    memcpy(&total, in + 8, 8);
    if (total > len || total < 16) return PF_ERR_FORMAT;
    if (in[5] == 0) { // already compr_mode none
        *written = (size_t)total;
        if (consumed) *consumed = (size_t)total;
        if (!out || cap < total) return PF_ERR_CAPACITY;
        memcpy(out, in, (size_t)total);
        return PF_OK;
    }
    std::vector<uint8_t> plain;
    size_t used = 0;
    // engine-less utility: the caller's capacity bounds the output once it is known; the sizing call
    // (out == NULL) is bounded by the largest object of this protocol (GaloisKeys at N = 16384)
    const size_t bound = (out && cap) ? cap : PF_INFLATE_HARD_CAP;
    const int zr = inflate_seal_stream(in, len, plain, &used, bound);
    if (zr == -7 && out && cap) { // larger than the caller's buffer: report the need without inflating past the ceiling
        if (inflate_seal_stream(in, len, plain, &used, PF_INFLATE_HARD_CAP) != 0) return PF_ERR_FORMAT;
        *written = plain.size();
        if (consumed) *consumed = used;
        return PF_ERR_CAPACITY;
    }
    if (zr != 0) return PF_ERR_FORMAT; // zstd, truncated, corrupt or beyond the ceiling
    *written = plain.size();
    if (consumed) *consumed = used;
    if (!out || cap < plain.size()) return PF_ERR_CAPACITY;
    memcpy(out, plain.data(), plain.size());
    return PF_OK;
}

int pf_seal_ct_expand_batch(const uint8_t *in, size_t in_bytes, const uint64_t *offsets, uint64_t ncts, uint64_t poly_degree,
                            const uint64_t *data_primes, uint32_t nprimes, uint8_t *out, size_t cap, uint64_t *out_offsets,
                            uint32_t threads) {
    if (!in || !offsets || !out_offsets || !data_primes || !nprimes || nprimes > PF_MAX_PRIMES) return PF_ERR_INVALID;
    if (poly_degree < 2 || poly_degree > 32768 || (poly_degree & (poly_degree - 1))) return PF_ERR_INVALID;
    for (uint32_t j = 0; j < nprimes; j++)
        if (data_primes[j] < 2 || data_primes[j] >> 61) return PF_ERR_INVALID;
    for (uint64_t c = 0; c < ncts; c++)
        if (offsets[c] > offsets[c + 1] || offsets[c + 1] > in_bytes) return PF_ERR_INVALID;
    std::vector<std::vector<uint8_t>> norm;
    const int nr = normalize_ct_streams(in, offsets, (size_t)ncts, poly_degree, data_primes, nprimes, norm, threads ? threads : host_threads(), nullptr, nullptr);
    if (nr) return PF_ERR_FORMAT;
    size_t pos = 0;
    out_offsets[0] = 0;
    for (uint64_t c = 0; c < ncts; c++) {
        const bool n = !norm.empty() && !norm[c].empty();
        const uint8_t *src = n ? norm[c].data() : in + offsets[c];
        size_t len = n ? norm[c].size() : (size_t)(offsets[c + 1] - offsets[c]);
        if (!n) { // an untouched stream: only its own bytes
            uint64_t total;
            if (len < 16 || src[0] != 0x5E || src[1] != 0xA1) return PF_ERR_FORMAT;
            memcpy(&total, src + 8, 8);
            if (total > len || total < 16) return PF_ERR_FORMAT;
            len = (size_t)total;
        }
        if (out && pos + len <= cap) memcpy(out + pos, src, len);
        pos += len;
        out_offsets[c + 1] = pos;
    }
    return (!out || pos > cap) ? PF_ERR_CAPACITY : PF_OK;
}

int pf_seal_galois_keys_expand(const uint8_t *in, size_t len, uint64_t poly_degree, const uint64_t *key_primes, uint32_t nprimes,
                               uint8_t *out, size_t cap, size_t *written) {
    if (!in || !written || !key_primes || nprimes < 2 || nprimes > PF_MAX_PRIMES) return PF_ERR_INVALID;
    if (poly_degree < 2 || poly_degree > 32768 || (poly_degree & (poly_degree - 1))) return PF_ERR_INVALID;
    for (uint32_t j = 0; j < nprimes; j++)
        if (key_primes[j] < 2 || key_primes[j] >> 61) return PF_ERR_INVALID;
    if (len < 16 || in[0] != 0x5E || in[1] != 0xA1) return PF_ERR_FORMAT;
    const size_t key_bytes = (size_t)(nprimes - 1) * (SEAL_CT_HEADER + (size_t)2 * nprimes * poly_degree * 8);
    const size_t bound = 16 + 32 + 8 + (size_t)poly_degree * 8 + (size_t)PF_MAX_GALOIS_KEYS * key_bytes;
    std::vector<uint8_t> plain, expanded;
    const uint8_t *src = in;
    size_t slen = len;
    const int zr = inflate_seal_stream(in, len, plain, nullptr, bound);
    if (zr < 0) return PF_ERR_FORMAT;
    if (zr == 0) {
        src = plain.data();
        slen = plain.size();
    }
    const int er = expand_galois_keys_stream(src, slen, poly_degree, key_primes, nprimes, expanded);
    if (er < 0) return PF_ERR_FORMAT;
    if (er == 0) {
        src = expanded.data();
        slen = expanded.size();
    } else {
        uint64_t total;
        memcpy(&total, src + 8, 8);
        slen = (size_t)total;
    }
    *written = slen;
    if (!out || cap < slen) return PF_ERR_CAPACITY;
    memcpy(out, src, slen);
    return PF_OK;
}

int pf_seal_ct_expand(const uint8_t *in, size_t len, uint64_t poly_degree, const uint64_t *data_primes, uint32_t nprimes,
                      uint8_t *out, size_t cap, size_t *written, size_t *consumed) {
    if (!in || !written || !data_primes || !nprimes || nprimes > PF_MAX_PRIMES) return PF_ERR_INVALID;
    if (poly_degree < 2 || poly_degree > 32768 || (poly_degree & (poly_degree - 1))) return PF_ERR_INVALID;
    for (uint32_t j = 0; j < nprimes; j++)
        if (data_primes[j] < 2 || data_primes[j] >> 61) return PF_ERR_INVALID;
    if (len < 16 || in[0] != 0x5E || in[1] != 0xA1) return PF_ERR_FORMAT;
    const size_t full = SEAL_CT_HEADER + (size_t)2 * nprimes * poly_degree * 8;
    std::vector<uint8_t> plain, expanded;
    size_t used = 0;
    const uint8_t *src = in;
    size_t slen = len;
    const int zr = inflate_seal_stream(in, len, plain, &used, full);
    if (zr < 0) return PF_ERR_FORMAT;
    if (zr == 0) {
        src = plain.data();
        slen = plain.size();
    } else {
        uint64_t total;
        memcpy(&total, in + 8, 8);
        if (total > len || total < 16) return PF_ERR_FORMAT;
        used = (size_t)total;
    }
    const int er = expand_seeded_stream(src, slen, poly_degree, data_primes, nprimes, expanded);
    if (er < 0) return PF_ERR_FORMAT;
    const uint8_t *res = er == 0 ? expanded.data() : src;
    uint64_t res_len;
    if (er == 0) res_len = expanded.size();
    else memcpy(&res_len, src + 8, 8);
    *written = (size_t)res_len;
    if (consumed) *consumed = used;
    if (!out || cap < res_len) return PF_ERR_CAPACITY;
    memcpy(out, res, (size_t)res_len);
    return PF_OK;
}

int pf_parms_id(uint64_t poly_degree, const uint64_t *coeff_primes, uint32_t nprimes, uint64_t plain_modulus, uint64_t out[4]) {
    if (!coeff_primes || !out || !nprimes || nprimes > 64) return PF_ERR_INVALID;
    pfh::seal_parms_id(poly_degree, coeff_primes, nprimes, plain_modulus, out);
    return PF_OK;
}

const char *pf_last_error(const pf_engine *e) { return e ? e->err.c_str() : g_tls_error.c_str(); }

int pf_engine_create(const pf_params *prm, pf_engine **out) {
    if (!out) return PF_ERR_INVALID;
    *out = nullptr;
    auto tls_fail = [](int code, const std::string &msg) {
        g_tls_error = msg;
        return code;
    };
    if (!prm || prm->struct_size != sizeof(pf_params)) return tls_fail(PF_ERR_INVALID, "pf_params.struct_size mismatch");
    const uint64_t N = prm->poly_degree;
    if (N < 1024 || N > 16384 || (N & (N - 1))) return tls_fail(PF_ERR_INVALID, "poly_degree must be a power of two in [1024,16384]");
    if (prm->num_primes < 2 || prm->num_primes > PF_MAX_PRIMES) return tls_fail(PF_ERR_INVALID, "num_primes must be in [2,16]");
    if (!prm->world || prm->rank >= prm->world) return tls_fail(PF_ERR_INVALID, "rank/world invalid");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev <= 0 || prm->device < 0 || prm->device >= ndev)
        return tls_fail(PF_ERR_CUDA, std::string("no usable CUDA device (this engine has no CPU path): ") +
                                         (ce != cudaSuccess ? cudaGetErrorString(ce) : "device ordinal out of range"));
    pf_engine *e = new pf_engine();
    e->prm = *prm;
    e->N = (int)N;
    e->logn = __builtin_ctzll(N);
    e->k = (int)prm->num_primes;
    e->L = e->k - 1;
    e->t = prm->plain_modulus;
    e->Lr = prm->result_limbs ? (int)prm->result_limbs : e->L;
    auto bail = [&](int code) {
        g_tls_error = e->err;
        delete e;
        return code;
    };
    if (e->Lr < 1 || e->Lr > e->L) return bail(e->fail(PF_ERR_INVALID, "result_limbs must be in [1, L]"));
    for (int j = 0; j < e->k; j++) {
        const u64 q = prm->primes[j];
        if (q >> 61 || q < 2 || (q - 1) % (2 * N) || !pfh::is_prime(q))
            return bail(e->fail(PF_ERR_INVALID, "prime %d (%llu) must be a prime = 1 mod 2N below 2^61", j, (unsigned long long)q));
        for (int i = 0; i < j; i++)
            if (prm->primes[i] == q) return bail(e->fail(PF_ERR_INVALID, "primes must be distinct"));
        e->h_q.push_back(q);
        e->max_prime_bits = std::max(e->max_prime_bits, 64 - __builtin_clzll(q));
    }
    if (e->t < 2 || e->t >> 32 || (e->t - 1) % (2 * N) || !pfh::is_prime(e->t))
        return bail(e->fail(PF_ERR_INVALID, "plain_modulus must be a prime = 1 mod 2N below 2^32"));
    // layout
    e->d = prm->dim;
    u32 dp = 1;
    while (dp < e->d) dp <<= 1;
    e->d_pad = dp;
    e->m = prm->query_cts ? prm->query_cts : 1;
    e->g = prm->partial_g ? prm->partial_g : 1;
    if (!e->d || (e->m & (e->m - 1)) || (e->g & (e->g - 1)) || dp % e->m || (dp / e->m) % e->g || dp / e->m > N / 2)
        return bail(e->fail(PF_ERR_INVALID, "layout invalid: need m | d_pad, g | d_pad/m, d_pad/m <= N/2 (powers of two)"));
    e->dc = dp / e->m;
    e->R = e->dc / e->g;
    e->K = e->m * e->R;
    e->C = (u32)(N / e->g);
    if (e->K > 128) return bail(e->fail(PF_ERR_INVALID, "K = m*d_pad/(m*g) = %u diagonals per block exceeds 128; raise g", e->K));
    // split-operand lazy sums (pf_mac.cuh) need 2*ceil(bits/2) + 2 + log2(terms) <= 64
    const int sbits = 2 * ((e->max_prime_bits + 1) / 2) + 2;
    int logk = 0, logl = 0;
    while ((1u << logk) < e->K) logk++;
    while ((1 << logl) < e->L) logl++;
    e->mac_wide = sbits + logk > 64;
    {
        const int fp_limit = e->logn <= 13 ? 44 : 49;
        const char *env = getenv("PF_NTT_FP");
        e->ntt_fp = e->max_prime_bits <= fp_limit && (!env || atoi(env) != 0);
    }
    e->mac_fpred = !e->mac_wide && e->max_prime_bits + logk <= 50 && !getenv("PF_MAC_NO_FPRED");
    e->ks_fpred = e->max_prime_bits + logl <= 50 && !getenv("PF_KS_NO_FPRED");
    for (int i = 1; i <= e->L; i++) pfh::seal_parms_id(N, prm->primes, (uint32_t)i, prm->plain_modulus, e->level_pid[i]);
    if (sbits + logl > 64)
        return bail(e->fail(PF_ERR_INVALID, "coefficient primes of %d bits with %d limbs overflow the key-switch accumulator", e->max_prime_bits, e->L));
    if (cudaSetDevice(prm->device) != cudaSuccess) return bail(e->fail(PF_ERR_CUDA, "cudaSetDevice(%d) failed", prm->device));
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(e->fail(PF_ERR_CUDA, "cudaStreamCreate failed"));
    e->own_stream = true;
    if (cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithPriority(&e->coarse_stream, cudaStreamNonBlocking, -1) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_group[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_group[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->upload_stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(e->fail(PF_ERR_CUDA, "copy stream / event creation failed"));
    for (auto &fl : e->flights) {
        for (auto &ev : fl.ev_up)
            if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess)
                return bail(e->fail(PF_ERR_CUDA, "event creation failed"));
        if (cudaEventCreateWithFlags(&fl.done, cudaEventDisableTiming) != cudaSuccess)
            return bail(e->fail(PF_ERR_CUDA, "event creation failed"));
        for (auto &ev : fl.tl) cudaEventCreate(&ev);
    }
    e->timeline = getenv("PF_DEBUG_TIMELINE") != nullptr;
    cudaEventCreate(&e->tl_base);
    if (cudaHostAlloc((void **)&e->h_err_word, sizeof(unsigned long long), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&e->d_err_word, e->h_err_word, 0) != cudaSuccess)
        return bail(e->fail(PF_ERR_CUDA, "error word allocation failed"));
    *e->h_err_word = 0;
    if (const char *ft = getenv("PF_FLAG_TIMEOUT_MS")) e->flag_timeout_ns = (unsigned long long)std::max(1, atoi(ft)) * 1000000ull;
    cudaError_t ar = cudaSuccess;
    switch (e->logn) {
    case 10: ar = set_ntt_attrs<10>(); break;
    case 11: ar = set_ntt_attrs<11>(); break;
    case 12: ar = set_ntt_attrs<12>(); break;
    case 13: ar = set_ntt_attrs<13>(); break;
    case 14: ar = set_ntt_attrs<14>(); break;
    }
    if (ar != cudaSuccess) return bail(e->fail(PF_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(ar)));
    int rc = build_tables(e);
    if (rc) return bail(rc);
    *out = e;
    return PF_OK;
}

void pf_engine_destroy(pf_engine *e) {
    if (!e) return;
    cudaSetDevice(e->prm.device);
    cudaStreamSynchronize(e->stream);
    drain_events(e);
    for (auto ev : e->event_pool) cudaEventDestroy(ev);
    for (auto ev : e->arena_ev)
        if (ev) cudaEventDestroy(ev);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->upload_stream) cudaStreamDestroy(e->upload_stream);
    for (auto &fl : e->flights) {
        if (fl.busy && fl.done) cudaEventSynchronize(fl.done);
        for (auto ev : fl.ev_up)
            if (ev) cudaEventDestroy(ev);
        if (fl.done) cudaEventDestroy(fl.done);
    }
    if (e->coarse_stream) cudaStreamDestroy(e->coarse_stream);
    for (auto ev : e->ev_group)
        if (ev) cudaEventDestroy(ev);
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    if (e->h_err_word) cudaFreeHost(e->h_err_word);
    delete e;
}

void *pf_engine_stream(pf_engine *e) { return e ? (void *)e->stream : nullptr; }

int pf_engine_set_stream(pf_engine *e, void *s) {
    if (!e) return PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    cudaStreamSynchronize(e->stream);
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    e->stream = (cudaStream_t)s;
    e->own_stream = false;
    return PF_OK;
}

int pf_engine_synchronize(pf_engine *e) {
    if (!e) return PF_ERR_INVALID;
    CK(cudaStreamSynchronize(e->stream));
    return check_device_error(e);
}

int pf_timing_enable(pf_engine *e, int on) {
    if (!e) return PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    e->timing = on != 0;
    return PF_OK;
}

int pf_timing_read(pf_engine *e, float *ms, uint64_t *launches, int reset) {
    if (!e) return PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaStreamSynchronize(e->stream));
    drain_events(e);
    for (int i = 0; i < PF_T_COUNT; i++) {
        if (ms) ms[i] = e->t_ms[i];
        if (launches) launches[i] = e->t_launch[i];
        if (reset) {
            e->t_ms[i] = 0;
            e->t_launch[i] = 0;
        }
    }
    return PF_OK;
}

uint64_t pf_launch_count(pf_engine *e) { return e ? e->launches.load() : 0; }

// ---- peer-memory gather buffers (CUDA IPC over NVLink) -------------------------------------------
int pf_ipc_alloc(pf_engine *e, size_t bytes, void **dptr, uint8_t handle[PF_IPC_HANDLE_BYTES]) {
    if (!e || !dptr || !handle || !bytes) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == PF_IPC_HANDLE_BYTES, "IPC handle size");
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->prm.device));
    void *p = nullptr;
    CK(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t r = cudaIpcGetMemHandle(&h, p);
    if (r != cudaSuccess) {
        cudaFree(p);
        return e->fail(PF_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(r));
    }
    memcpy(handle, &h, sizeof(h));
    *dptr = p;
    return PF_OK;
}

int pf_ipc_open(pf_engine *e, const uint8_t handle[PF_IPC_HANDLE_BYTES], void **dptr) {
    if (!e || !dptr || !handle) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->prm.device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    CK(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PF_OK;
}

int pf_ipc_close(pf_engine *e, void *dptr) {
    if (!e || !dptr) return PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->prm.device));
    CK(cudaIpcCloseMemHandle(dptr));
    return PF_OK;
}

int pf_copy_async(pf_engine *e, void *dst, const void *src, size_t bytes, void *cuda_stream) {
    if (!e || !dst || !src) return PF_ERR_INVALID;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : e->stream;
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st));
    return PF_OK;
}

int pf_flag_write(pf_engine *e, void *flag, uint32_t value, void *cuda_stream) {
    if (!e || !flag) return PF_ERR_INVALID;
    int rc = check_device_error(e);
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : e->stream;
    flag_write_kernel<<<1, 1, 0, st>>>((volatile unsigned *)flag, value);
    CK(cudaGetLastError());
    e->launches++;
    return PF_OK;
}

int pf_flag_wait(pf_engine *e, const void *flag, uint32_t value, void *cuda_stream) {
    if (!e || !flag) return PF_ERR_INVALID;
    int rc = check_device_error(e);
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : e->stream;
    flag_wait_kernel<<<1, 1, 0, st>>>((const volatile unsigned *)flag, value, e->flag_timeout_ns, e->d_err_word);
    CK(cudaGetLastError());
    e->launches++;
    return PF_OK;
}

int pf_device_checksum(pf_engine *e, const void *dptr, uint64_t nwords, uint64_t *out, void *cuda_stream) {
    if (!e || !dptr || !out) return PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->prm.device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : e->stream;
    CK(e->s_cksum.ensure(8));
    CK(cudaMemsetAsync(e->s_cksum.p, 0, 8, st));
    if (nwords) {
        const unsigned blocks = (unsigned)std::min<uint64_t>(148 * 8, (nwords + 255) / 256);
        checksum_kernel<<<blocks, 256, 0, st>>>((const u64 *)dptr, (size_t)nwords, e->s_cksum.as<u64>());
        CK(cudaGetLastError());
        e->launches++;
    }
    CK(cudaMemcpyAsync(out, e->s_cksum.p, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return check_device_error(e);
}

int pf_ipc_free(pf_engine *e, void *dptr) {
    if (!e || !dptr) return PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->prm.device));
    CK(cudaFree(dptr));
    return PF_OK;
}

uint32_t pf_galois_elt_from_step(pf_engine *e, int step) { return e ? galois_elt_from_step(e, step) : 0; }

// ---- index ------------------------------------------------------------------------------------
int pf_load_index(pf_engine *e, uint64_t nlist, const float *centroids, const int64_t *list_offsets,
                  const int64_t *ids, const float *vectors) {
    if (!e || !nlist || !centroids || !list_offsets || !ids || !vectors) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->prm.device));
    const u32 d = e->d;
    const int N = e->N, L = e->L;
    const uint64_t ntotal = (uint64_t)list_offsets[nlist];
    for (uint64_t l = 0; l < nlist; l++)
        if (list_offsets[l + 1] < list_offsets[l] || list_offsets[0] != 0)
            return e->fail(PF_ERR_INVALID, "list_offsets must start at 0 and be non-decreasing");
    e->has_index = false;
    e->nlist = nlist;
    e->ntotal = ntotal;
    e->h_centroids.assign(centroids, centroids + nlist * d);
    e->h_list_offsets.assign(list_offsets, list_offsets + nlist + 1);
    e->h_ids.assign(ids, ids + ntotal);
    e->pq_M = 0; // a product quantizer belongs to the index it was trained on
    CK(e->d_centroids.ensure(nlist * d * sizeof(float)));
    CK(cudaMemcpyAsync(e->d_centroids.p, centroids, nlist * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    CK(e->d_ids.ensure(std::max<size_t>(8, ntotal * 8)));
    CK(cudaMemcpyAsync(e->d_ids.p, ids, ntotal * 8, cudaMemcpyHostToDevice, e->stream));
    CK(e->d_base_f32.ensure(std::max<size_t>(4, ntotal * d * sizeof(float))));
    CK(cudaMemcpyAsync(e->d_base_f32.p, vectors, ntotal * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    // id -> list-ordered position, when ids are exactly the base row numbers (ref: server_lib.cpp:154-156)
    e->ids_are_rows = true;
    {
        std::vector<long long> pos(ntotal, -1);
        for (uint64_t i = 0; i < ntotal && e->ids_are_rows; i++) {
            const int64_t id = ids[i];
            if (id < 0 || (uint64_t)id >= ntotal || pos[id] != -1) e->ids_are_rows = false;
            else pos[id] = (long long)i;
        }
        if (e->ids_are_rows) {
            CK(e->d_pos_of_id.ensure(std::max<size_t>(8, ntotal * 8)));
            CK(cudaMemcpy(e->d_pos_of_id.p, pos.data(), ntotal * 8, cudaMemcpyHostToDevice));
        }
    }
    // uint8 copy for the encoder; rejects non-integer / out-of-range values
    CK(e->d_base_u8.ensure(std::max<size_t>(4, ntotal * d)));
    CK(e->s_tmp.ensure_grow(sizeof(int)));
    CK(cudaMemsetAsync(e->s_tmp.p, 0, sizeof(int), e->stream));
    if (ntotal) {
        const size_t n = ntotal * d;
        quantize_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->d_base_f32.as<float>(),
                                                                            e->d_base_u8.as<unsigned char>(), n,
                                                                            e->s_tmp.as<int>());
        e->launches++;
    }
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, e->s_tmp.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    const bool encodable = !bad;

    // blocks of the lists this rank owns
    e->blocks.clear();
    e->list_block_start.assign(nlist + 1, 0);
    for (uint64_t l = 0; l < nlist; l++) {
        e->list_block_start[l] = (long long)e->blocks.size();
        if (l % e->prm.world != e->prm.rank) continue;
        const long long n = list_offsets[l + 1] - list_offsets[l];
        for (long long o = 0; o < n; o += e->C)
            e->blocks.push_back(BlockInfo{list_offsets[l] + o, (u32)std::min<long long>(e->C, n - o), (u32)l});
    }
    e->list_block_start[nlist] = (long long)e->blocks.size();
    const size_t nb = e->blocks.size();
    e->diag_block_words = (size_t)e->K * L * N;
    e->norm_block_words = (size_t)L * N;
    if (!encodable) {
        // plaintext stages remain usable; the encrypted path needs integer data
        e->blocks.clear();
        e->list_block_start.assign(nlist + 1, 0);
        e->d_diag.release();
        e->d_norm.release();
        e->has_index = true;
        return e->fail(PF_ERR_INVALID, "base vectors are not integers in [0,255]: encrypted search disabled, plaintext stages loaded"), PF_OK;
    }
    CK(e->d_diag.ensure(std::max<size_t>(8, nb * e->diag_block_words * 8)));
    CK(e->d_norm.ensure(std::max<size_t>(8, nb * e->norm_block_words * 8)));

    // encode in batches of zb blocks
    const size_t plain_words = (size_t)(e->K + 1) * N;
    const size_t zb = std::max<size_t>(1, std::min<size_t>(nb ? nb : 1, ((size_t)256 << 20) / (plain_words * 8)));
    CK(e->s_plain.ensure_grow(zb * plain_words * 8));
    CK(e->s_encblocks.ensure_grow(std::max<size_t>(1, nb) * sizeof(EncodeBlock)));
    {
        std::vector<EncodeBlock> eb(nb);
        for (size_t b = 0; b < nb; b++) eb[b] = EncodeBlock{e->blocks[b].vec_offset, e->blocks[b].nvec, 0};
        if (nb) CK(cudaMemcpy(e->s_encblocks.p, eb.data(), nb * sizeof(EncodeBlock), cudaMemcpyHostToDevice));
    }
    for (size_t b0 = 0; b0 < nb; b0 += zb) {
        const unsigned nz = (unsigned)std::min(zb, nb - b0);
        EncodeParams ep{};
        ep.base = e->d_base_u8.as<unsigned char>();
        ep.blocks = e->s_encblocks.as<EncodeBlock>() + b0;
        ep.inv_index_map = e->d_inv_index_map.as<u32>();
        ep.plain = e->s_plain.as<u64>();
        ep.t = e->t;
        ep.N = N;
        ep.d = (int)d;
        ep.dc = (int)e->dc;
        ep.R = (int)e->R;
        ep.K = (int)e->K;
        ep.g = (int)e->g;
        encode_slots_kernel<<<dim3(N / 256, e->K + 1, nz), 256, 0, e->stream>>>(ep);
        e->launches++;
        // BatchEncoder::encode: inverse NTT mod t, in place, all K+1 plaintexts of every block
        NttParams ip{};
        ip.in = ip.out = e->s_plain.as<u64>();
        ip.in_sy = ip.out_sy = N;
        ip.in_sz = ip.out_sz = (long long)plain_words;
        ip.mod_map[0] = e->k;
        launch_ntt(e, NTT_IN_PLAIN, true, ip, dim3(1, e->K + 1, nz));
        // diagonals: centred lift + forward NTT per limb
        NttParams fp{};
        fp.in = e->s_plain.as<u64>();
        fp.in_sx = 0;
        fp.in_sy = N;
        fp.in_sz = (long long)plain_words;
        fp.out = e->d_diag.as<u64>() + b0 * e->diag_block_words;
        fp.out_sx = N;
        fp.out_sy = (long long)L * N;
        fp.out_sz = (long long)e->diag_block_words;
        fp.lift_t = e->t;
        fp.lift_thr = (e->t + 1) >> 1;
        fp.out_split = e->mac_wide ? 0 : 1;
        for (int i = 0; i < L; i++) fp.mod_map[i] = i;
        launch_ntt(e, NTT_IN_LIFT, false, fp, dim3(L, e->K, nz));
        // norms: BFV scaling variant, then forward NTT
        ScaleParams sp{};
        sp.plain = e->s_plain.as<u64>();
        sp.plain_sz = (long long)plain_words;
        sp.plain_off = (long long)e->K * N;
        sp.out = e->d_norm.as<u64>() + b0 * e->norm_block_words;
        sp.mods = e->d_mods.as<DevModulus>();
        sp.t = e->t;
        sp.q_mod_t = e->q_mod_t;
        sp.half_t = (e->t + 1) >> 1;
        for (int i = 0; i < L; i++) sp.delta_mod_q[i] = e->delta_mod_q[i];
        sp.N = N;
        sp.L = L;
        scale_plain_kernel<<<dim3(N / 256, 1, nz), 256, 0, e->stream>>>(sp);
        e->launches++;
        ntt_limbs(e, sp.out, sp.out, nz, false);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    e->has_index = true;
    return PF_OK;
}

int pf_get_index_info(pf_engine *e, pf_index_info *o) {
    if (!e || !o) return PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    memset(o, 0, sizeof(*o));
    o->nlist = e->nlist;
    o->ntotal = e->ntotal;
    o->nblocks_local = e->blocks.size();
    uint64_t nb = 0;
    if (e->has_index)
        for (uint64_t l = 0; l < e->nlist; l++) {
            const uint64_t n = (uint64_t)(e->h_list_offsets[l + 1] - e->h_list_offsets[l]);
            nb += (n + e->C - 1) / e->C;
        }
    o->nblocks = nb;
    o->db_bytes = e->blocks.size() * (e->diag_block_words + e->norm_block_words) * 8;
    o->K = e->K;
    o->C = e->C;
    o->R = e->R;
    o->d_pad = e->d_pad;
    o->L = (uint32_t)e->L;
    o->k = (uint32_t)e->k;
    return PF_OK;
}

int pf_retrieve_centroids(pf_engine *e, float *out, uint64_t cap) {
    if (!e || !out) return PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->has_index) return e->fail(PF_ERR_STATE, "no index loaded");
    if (cap < e->h_centroids.size()) return e->fail(PF_ERR_CAPACITY, "need %zu floats", e->h_centroids.size());
    memcpy(out, e->h_centroids.data(), e->h_centroids.size() * sizeof(float));
    return PF_OK;
}

// ---- stage 1 ------------------------------------------------------------------------------------
int pf_coarse_quantize(pf_engine *e, uint64_t nq, const float *x, uint32_t nprobe, int64_t *out_idx, float *out_dist) {
    if (!e || !x || !out_idx) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->has_index) return e->fail(PF_ERR_STATE, "no index loaded");
    // ref: src/client/client_lib.cpp:96-99 throws when NPROBE exceeds the centroid count
    if (!nprobe || nprobe > e->nlist) return e->fail(PF_ERR_INVALID, "Centroids count is not equal to NPROBE (nprobe %u, nlist %llu)", nprobe, (unsigned long long)e->nlist);
    if (!nq) return PF_OK;
    CK(cudaSetDevice(e->prm.device));
    const u32 d = e->d;
    const int nlist = (int)e->nlist;
    // Stage 1 only reads the centroid table: it runs on its own stream so that the next batch can be
    // quantized while the encrypted pipeline of the current batch still occupies the engine stream.
    cudaStream_t cs = e->coarse_stream;
    PhaseTimer pt(e, PF_T_COARSE, cs);
    // Inputs go through a pinned staging buffer (asynchronous H2D), outputs are written by the kernel straight
    // into mapped host memory: a D2H copy of these few KB would queue behind the response download of the
    // previous search on the copy engine (measured: 1.3 ms per call instead of 0.1 ms, and a GPU idle meanwhile).
    CK(e->s_cx.ensure_grow(nq * d * sizeof(float)));
    CK(e->s_dist.ensure_grow(nq * (size_t)nlist * sizeof(float)));
    CK(e->s_keys.ensure_grow(nq * (size_t)nlist * sizeof(u64)));
    CK(e->m_cx.ensure(nq * d * sizeof(float)));
    CK(e->m_cidx.ensure(nq * nprobe * sizeof(long long)));
    CK(e->m_cdist.ensure(nq * nprobe * sizeof(float)));
    memcpy(e->m_cx.h, x, nq * d * sizeof(float));
    CK(cudaMemcpyAsync(e->s_cx.p, e->m_cx.h, nq * d * sizeof(float), cudaMemcpyHostToDevice, cs));
    long long *o_idx = (long long *)e->m_cidx.d;
    float *o_dist = (float *)e->m_cdist.d;
    for (uint64_t q0 = 0; q0 < nq; q0 += 32768) {
        const unsigned nqb = (unsigned)std::min<uint64_t>(32768, nq - q0);
        coarse_dist_kernel<<<dim3((nlist + 127) / 128, nqb), 128, d * sizeof(float), cs>>>(
            e->s_cx.as<float>() + q0 * d, e->d_centroids.as<float>(), e->s_dist.as<float>() + q0 * nlist, nlist, (int)d);
        int M = 1;
        while (M < (int)nprobe) M <<= 1;
        const bool old_topk = getenv("PF_TOPK_ITER") != nullptr; // per call: tests flip it
        if (M <= 4096 && !old_topk) // radix select + bitonic sort of the selected keys in shared memory
            topk_radix_kernel<<<nqb, 256, (size_t)M * 8, cs>>>(e->s_dist.as<float>() + q0 * nlist, e->s_keys.as<u64>() + q0 * nlist,
                                                              o_idx + q0 * nprobe, o_dist + q0 * nprobe, nlist, (int)nprobe, M);
        else
            topk_select_kernel<<<nqb, 256, 0, cs>>>(e->s_dist.as<float>() + q0 * nlist, e->s_keys.as<u64>() + q0 * nlist,
                                                     o_idx + q0 * nprobe, o_dist + q0 * nprobe, nlist, (int)nprobe);
        e->launches += 2;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(cs));
    memcpy(out_idx, e->m_cidx.h, nq * nprobe * sizeof(long long));
    if (out_dist) memcpy(out_dist, e->m_cdist.h, nq * nprobe * sizeof(float));
    return PF_OK;
}

// ---- stage 2, plaintext ---------------------------------------------------------------------------
int pf_search_lists_plain(pf_engine *e, uint64_t nq, const float *x, const int64_t *idx, uint32_t nprobe, float *dist,
                          int64_t *labels, uint64_t cap, uint64_t *list_sizes, uint64_t *total) {
    if (!e || !x || !idx || !list_sizes) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->has_index) return e->fail(PF_ERR_STATE, "no index loaded");
    CK(cudaSetDevice(e->prm.device));
    std::vector<ListJob> jobs;
    uint64_t w = 0;
    for (uint64_t i = 0; i < nq; i++) {
        uint64_t cnt = 0;
        for (uint32_t p = 0; p < nprobe; p++) {
            const int64_t l = idx[i * nprobe + p];
            if (l < 0 || (uint64_t)l >= e->nlist)
                return e->fail(PF_ERR_INVALID, "list id %lld out of range", (long long)l);
            const long long n = e->h_list_offsets[l + 1] - e->h_list_offsets[l];
            if (n > 0) jobs.push_back(ListJob{e->h_list_offsets[l], (long long)w, (int)n, (int)i});
            w += (uint64_t)n;
            cnt += (uint64_t)n;
        }
        list_sizes[i] = cnt;
    }
    if (total) *total = w;
    if (w > cap || (w && (!dist || !labels))) return e->fail(PF_ERR_CAPACITY, "output needs %llu entries, capacity %llu", (unsigned long long)w, (unsigned long long)cap);
    if (!w) return PF_OK;
    const u32 d = e->d;
    CK(e->s_x.ensure_grow(nq * d * sizeof(float)));
    CK(e->s_jobs.ensure_grow(jobs.size() * sizeof(ListJob)));
    CK(e->s_pl_dist.ensure_grow(w * sizeof(float)));
    CK(e->s_pl_labels.ensure_grow(w * sizeof(long long)));
    CK(cudaMemcpyAsync(e->s_x.p, x, nq * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->s_jobs.p, jobs.data(), jobs.size() * sizeof(ListJob), cudaMemcpyHostToDevice, e->stream));
    list_l2_kernel<<<(unsigned)jobs.size(), 128, d * sizeof(float), e->stream>>>(
        e->s_x.as<float>(), e->d_base_f32.as<float>(), e->d_ids.as<long long>(), e->s_jobs.as<ListJob>(),
        e->s_pl_dist.as<float>(), e->s_pl_labels.as<long long>(), (int)d, w);
    e->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(dist, e->s_pl_dist.p, w * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(labels, e->s_pl_labels.p, w * sizeof(long long), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return PF_OK;
}

// ---- stage 2 as the reference's FAISS fork computes it today: PQ-ADC --------------------------------
int pf_load_pq(pf_engine *e, uint32_t M, uint32_t nbits, const float *pq_centroids, const uint8_t *codes) {
    if (!e || !pq_centroids || !codes) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->has_index) return e->fail(PF_ERR_STATE, "no index loaded");
    if (nbits != 8) return e->fail(PF_ERR_INVALID, "product quantizer with %u bits per sub-quantizer (only 8, as the reference builds it)", nbits);
    if (!M || e->d % M) return e->fail(PF_ERR_INVALID, "%u sub-quantizers do not divide the dimension %u", M, e->d);
    if (((size_t)e->d + (size_t)M * 256) * sizeof(float) > 200 * 1024)
        return e->fail(PF_ERR_INVALID, "distance table of %u sub-quantizers does not fit shared memory", M);
    CK(cudaSetDevice(e->prm.device));
    const size_t ntotal = (size_t)e->h_list_offsets[e->nlist];
    CK(e->d_pq_cent.ensure((size_t)256 * e->d * sizeof(float)));
    CK(e->d_pq_codes.ensure(std::max<size_t>(16, ntotal * M + 16))); // + 16: the 16-byte code loads of the last row stay inside
    CK(cudaMemcpyAsync(e->d_pq_cent.p, pq_centroids, (size_t)256 * e->d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    if (ntotal) CK(cudaMemcpyAsync(e->d_pq_codes.p, codes, ntotal * M, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->pq_M = M;
    return PF_OK;
}

int pf_search_lists_pq(pf_engine *e, uint64_t nq, const float *x, const int64_t *idx, uint32_t nprobe, float *dist,
                       int64_t *labels, uint64_t cap, uint64_t *list_sizes, uint64_t *total) {
    if (!e || !x || !idx || !list_sizes) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->has_index) return e->fail(PF_ERR_STATE, "no index loaded");
    if (!e->pq_M) return e->fail(PF_ERR_STATE, "no product quantizer loaded (pf_load_pq)");
    CK(cudaSetDevice(e->prm.device));
    std::vector<ListJob> jobs;
    std::vector<int> job_list;
    uint64_t w = 0;
    for (uint64_t i = 0; i < nq; i++) {
        uint64_t cnt = 0;
        for (uint32_t p = 0; p < nprobe; p++) {
            const int64_t l = idx[i * nprobe + p];
            if (l < 0 || (uint64_t)l >= e->nlist)
                return e->fail(PF_ERR_INVALID, "list id %lld out of range", (long long)l);
            const long long n = e->h_list_offsets[l + 1] - e->h_list_offsets[l];
            if (n > 0) {
                jobs.push_back(ListJob{e->h_list_offsets[l], (long long)w, (int)n, (int)i});
                job_list.push_back((int)l);
            }
            w += (uint64_t)n;
            cnt += (uint64_t)n;
        }
        list_sizes[i] = cnt;
    }
    if (total) *total = w;
    if (w > cap || (w && (!dist || !labels))) return e->fail(PF_ERR_CAPACITY, "output needs %llu entries, capacity %llu", (unsigned long long)w, (unsigned long long)cap);
    if (!w) return PF_OK;
    const u32 d = e->d, M = e->pq_M;
    const size_t jobs_bytes = (jobs.size() * sizeof(ListJob) + 15) & ~(size_t)15;
    CK(e->s_x.ensure_grow(nq * d * sizeof(float)));
    CK(e->s_jobs.ensure_grow(jobs_bytes + job_list.size() * sizeof(int)));
    CK(e->s_pl_dist.ensure_grow(w * sizeof(float)));
    CK(e->s_pl_labels.ensure_grow(w * sizeof(long long)));
    CK(cudaMemcpyAsync(e->s_x.p, x, nq * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->s_jobs.p, jobs.data(), jobs.size() * sizeof(ListJob), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->s_jobs.as<char>() + jobs_bytes, job_list.data(), job_list.size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    const size_t smem = ((size_t)d + (size_t)M * 256) * sizeof(float);
    static size_t attr_max = 48 * 1024;
    if (smem > attr_max) {
        CK(cudaFuncSetAttribute(list_pq_adc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_max = smem;
    }
    for (size_t j0 = 0; j0 < jobs.size(); j0 += 1u << 30) { // grid.x limit is 2^31 - 1; one launch in practice
        const size_t nj = std::min<size_t>(jobs.size() - j0, 1u << 30);
        list_pq_adc_kernel<<<(unsigned)nj, 256, smem, e->stream>>>(
            e->s_x.as<float>(), e->d_centroids.as<float>(), e->d_pq_cent.as<float>(), e->d_pq_codes.as<unsigned char>(),
            e->d_ids.as<long long>(), e->s_jobs.as<ListJob>() + j0, reinterpret_cast<const int *>(e->s_jobs.as<char>() + jobs_bytes) + j0,
            e->s_pl_dist.as<float>(), e->s_pl_labels.as<long long>(), (int)d, (int)M, w);
        e->launches++;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(dist, e->s_pl_dist.p, w * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(labels, e->s_pl_labels.p, w * sizeof(long long), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return PF_OK;
}

int pf_precise_search(pf_engine *e, uint64_t nq, const float *x, const int64_t *ids, uint32_t nids, float *out) {
    if (!e || !x || !ids || !out) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->has_index) return e->fail(PF_ERR_STATE, "no index loaded");
    if (!e->ids_are_rows) return e->fail(PF_ERR_STATE, "index ids are not base row numbers 0..ntotal-1");
    if (!nq || !nids) return PF_OK;
    CK(cudaSetDevice(e->prm.device));
    const u32 d = e->d;
    CK(e->s_x.ensure_grow(nq * d * sizeof(float)));
    CK(e->s_ids.ensure_grow(nq * nids * sizeof(long long)));
    CK(e->s_pl_dist.ensure_grow(nq * nids * sizeof(float)));
    CK(cudaMemcpyAsync(e->s_x.p, x, nq * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->s_ids.p, ids, nq * nids * sizeof(long long), cudaMemcpyHostToDevice, e->stream));
    for (uint64_t q0 = 0; q0 < nq; q0 += 32768) { // gridDim.y <= 65535
        const unsigned nqb = (unsigned)std::min<uint64_t>(32768, nq - q0);
        precise_l2_kernel<<<dim3((nids + 127) / 128, nqb), 128, d * sizeof(float), e->stream>>>(
            e->s_x.as<float>() + q0 * d, e->d_base_f32.as<float>(), e->d_pos_of_id.as<long long>(),
            e->s_ids.as<long long>() + q0 * nids, e->s_pl_dist.as<float>() + q0 * nids, (int)d, (int)nids,
            (long long)e->ntotal);
        e->launches++;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, e->s_pl_dist.p, nq * nids * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return PF_OK;
}

// ---- Galois keys ------------------------------------------------------------------------------------
int pf_set_galois_key(pf_engine *e, uint32_t elt, const uint64_t *words) {
    if (!e || !words) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->prm.device));
    return set_galois_key_words(e, elt, (const u64 *)words, false);
}

// SEAL KSwitchKeys::save_members inside a Serialization::Save envelope (kswitchkeys.h), compr none
int pf_load_galois_keys(pf_engine *e, const uint8_t *bytes, size_t len) {
    if (!e || !bytes) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->prm.device));
    std::vector<uint8_t> plain;
    {
        // a GaloisKeys object of this parameter set: header + parms_id + dim1 + one count per possible
        // element + at most PF_MAX_GALOIS_KEYS keys of L ciphertexts over k primes
        const size_t key_bytes = (size_t)e->L * (SEAL_CT_HEADER + (size_t)2 * e->k * e->N * 8);
        const size_t cap = 16 + 32 + 8 + (size_t)e->N * 8 + (size_t)PF_MAX_GALOIS_KEYS * key_bytes;
        const int zr = inflate_seal_stream(bytes, len, plain, nullptr, cap);
        if (zr == -7) return e->fail(PF_ERR_FORMAT, "compressed GaloisKeys stream inflates beyond %zu bytes (%d keys of this parameter set)", cap, PF_MAX_GALOIS_KEYS);
        if (zr < 0) return e->fail(PF_ERR_FORMAT, "malformed compressed GaloisKeys stream (code %d)", zr);
        if (zr == 0) {
            bytes = plain.data();
            len = plain.size();
        }
    }
    if (len < 16 + 32 + 8 || bytes[0] != 0x5E || bytes[1] != 0xA1 || bytes[5] != 0)
        return e->fail(PF_ERR_FORMAT, "not a SEAL stream (compr_mode none, zlib or zstd)");
    // Serializable<GaloisKeys>: seeded key ciphertexts are expanded first (host); a stream without any is parsed as it is
    std::vector<uint8_t> expanded;
    {
        const int xr = expand_galois_keys_stream(bytes, len, (uint64_t)e->N, reinterpret_cast<const uint64_t *>(e->h_q.data()), (uint32_t)e->k, expanded);
        if (xr == -8) return e->fail(PF_ERR_FORMAT, "GaloisKeys seeded with a PRNG other than blake2xb (unsupported)");
        if (xr == 0) {
            bytes = expanded.data();
            len = expanded.size();
        }
    }
    uint64_t total;
    memcpy(&total, bytes + 8, 8);
    if (total > len) return e->fail(PF_ERR_FORMAT, "truncated GaloisKeys stream");
    size_t off = 16 + 32;
    uint64_t dim1;
    memcpy(&dim1, bytes + off, 8);
    off += 8;
    if (dim1 > (uint64_t)e->N) return e->fail(PF_ERR_FORMAT, "GaloisKeys stream holds %llu slots, at most N = %d possible", (unsigned long long)dim1, e->N);
    const size_t key_ct_words = (size_t)2 * e->k * e->N;
    std::vector<u64> words((size_t)e->L * key_ct_words);
    for (uint64_t index = 0; index < dim1; index++) {
        if (off + 8 > total) return e->fail(PF_ERR_FORMAT, "truncated GaloisKeys stream");
        uint64_t dim2;
        memcpy(&dim2, bytes + off, 8);
        off += 8;
        if (!dim2) continue;
        if (dim2 != (uint64_t)e->L) return e->fail(PF_ERR_FORMAT, "key %llu has %llu parts, expected %d", (unsigned long long)index, (unsigned long long)dim2, e->L);
        for (uint64_t j = 0; j < dim2; j++) {
            int is_ntt;
            uint64_t pid[4], cms;
            size_t ctotal;
            if (parse_ct_prefix(e, bytes + off, total - off, &is_ntt, pid, &cms, &ctotal) || cms != (uint64_t)e->k || !is_ntt)
                return e->fail(PF_ERR_FORMAT, "malformed key ciphertext (index %llu part %llu)", (unsigned long long)index, (unsigned long long)j);
            memcpy(words.data() + j * key_ct_words, bytes + off + SEAL_CT_HEADER, key_ct_words * 8);
            off += ctotal;
        }
        int rc = set_galois_key_words(e, (u32)(2 * index + 1), words.data(), false);
        if (rc) return rc;
    }
    return PF_OK;
}

// ---- stage 2, encrypted -------------------------------------------------------------------------------
int pf_search_device(pf_engine *e, uint64_t nq, const uint64_t *d_query_cts, const int64_t *idx, uint32_t nprobe,
                     uint64_t *d_out, uint64_t cap_results, uint64_t *results_per_query, pf_search_stats *stats) {
    if (!e || !d_query_cts || !idx) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    HostTick htall("pf_search_device");
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->has_index || (!e->d_diag.p && e->ntotal)) return e->fail(PF_ERR_STATE, "no encodable index loaded");
    {
        HostTick ht("cudaSetDevice");
        CK(cudaSetDevice(e->prm.device));
    }
    PairPlan pl;
    int rc;
    {
        HostTick ht("plan_pairs");
        rc = plan_pairs(e, nq, idx, nprobe, pl);
    }
    if (rc) return rc;
    const uint64_t P = pl.pair_block.size();
    if (stats) {
        stats->nresults = P;
        stats->out_bytes = P * 2ull * e->Lr * e->N * 8;
        stats->useful_distances = pl.useful;
        stats->slot_distances = P * e->C;
    }
    if (results_per_query) memcpy(results_per_query, pl.results_per_query.data(), nq * sizeof(uint64_t));
    if (P > cap_results || (P && !d_out)) return e->fail(PF_ERR_CAPACITY, "need room for %llu result ciphertexts, capacity %llu", (unsigned long long)P, (unsigned long long)cap_results);
    {
        HostTick ht("arena_begin");
        rc = arena_begin(e);
    }
    if (rc) return rc;
    {
        HostTick ht("upload_plan");
        rc = upload_plan(e, pl);
    }
    if (rc) return rc;
    rc = search_core(e, 0, nq, (const u64 *)d_query_cts, pl, (u64 *)d_out, (size_t)2 * e->Lr * e->N);
    arena_end(e);
    return rc;
}

// The encrypted search is enqueued by pf_search_submit and completed by pf_search_collect: two calls may
// be in flight, so the upload and the compute of call i+1 overlap the device-to-host copy of call i
// (separate query / result buffers per flight; three streams: upload, engine, copy).  Everything that
// touches the caller's host buffers on the CPU (labels, sizes, offsets) is done inside submit while the
// GPU works; the SEAL stream headers in front of every result are stamped on the device, so collect is
// just a wait on the last copy.
static int submit_search(pf_engine *e, uint64_t nq, const uint8_t *query_cts, uint64_t query_bytes,
                         const uint64_t *ct_offsets, const int64_t *idx, uint32_t nprobe, uint8_t *out_cts,
                         uint64_t out_cap, uint64_t *result_offsets, uint64_t max_results, uint64_t *results_per_query,
                         int64_t *labels, uint64_t label_cap, uint64_t *list_sizes, uint64_t *probed_sizes,
                         pf_search_stats *stats, uint64_t *ticket) {
    if (!e || !query_cts || !ct_offsets || !idx || !ticket) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    HostTick htall("search_submit");
    std::lock_guard<std::mutex> lk(e->mu);
    int rc = check_device_error(e);
    if (rc) return rc;
    if (!e->has_index || (!e->d_diag.p && e->ntotal)) return e->fail(PF_ERR_STATE, "no encodable index loaded");
    int fi = -1;
    for (int i = 0; i < PF_MAX_FLIGHTS; i++)
        if (!e->flights[i].busy) {
            fi = i;
            break;
        }
    if (fi < 0) return e->fail(PF_ERR_STATE, "%d searches already in flight: collect one first", PF_MAX_FLIGHTS);
    pf_engine::Flight &fl = e->flights[fi];
    CK(cudaSetDevice(e->prm.device));
    const int L = e->L, N = e->N;
    const size_t ctw = (size_t)2 * L * N, ncts = nq * e->m;
    // the ciphertext streams sit inside [0, query_bytes): offsets ascending, every stream inside the blob
    for (size_t c = 0; c < ncts; c++)
        if (ct_offsets[c + 1] < ct_offsets[c] || ct_offsets[c + 1] > query_bytes)
            return e->fail(PF_ERR_INVALID, "ct_offsets[%zu..%zu] = [%llu, %llu) is not an ascending range inside the %llu-byte query blob", c,
                           c + 1, (unsigned long long)ct_offsets[c], (unsigned long long)ct_offsets[c + 1], (unsigned long long)query_bytes);
    PairPlan pl;
    {
        HostTick ht("plan_pairs");
        rc = plan_pairs(e, nq, idx, nprobe, pl);
    }
    if (rc) return rc;
    const uint64_t P = pl.pair_block.size();
    const size_t rw = (size_t)2 * e->Lr * N; // words of a result ciphertext
    // Result r occupies a slot of `slot` bytes; its SEAL stream starts at r*slot + RESULT_PAD so that
    // the ciphertext words sit at a 128-byte aligned offset (one aligned D2H per query group).
    const size_t slot = PF_RESULT_DATA_OFFSET + rw * 8;
    // labels / sizes of the owned probed lists (same packing as pf_search_lists_plain)
    // (the label copy itself happens further down, while the GPU works)
    uint64_t nlabels = 0;
    for (uint64_t i = 0; i < nq; i++) {
        uint64_t cnt = 0;
        for (uint32_t p = 0; p < nprobe; p++) {
            const int64_t l = idx[i * nprobe + p];
            const bool owned = (uint64_t)l % e->prm.world == e->prm.rank;
            const uint64_t n = owned ? (uint64_t)(e->h_list_offsets[l + 1] - e->h_list_offsets[l]) : 0;
            if (probed_sizes) probed_sizes[i * nprobe + p] = n;
            nlabels += n;
            cnt += n;
        }
        if (list_sizes) list_sizes[i] = cnt;
    }
    if (stats) {
        stats->nresults = P;
        stats->out_bytes = P * slot;
        stats->useful_distances = pl.useful;
        stats->slot_distances = P * e->C;
    }
    if (results_per_query) memcpy(results_per_query, pl.results_per_query.data(), nq * sizeof(uint64_t));
    if (P > max_results || P * slot > out_cap || (P && !out_cts) || (labels && nlabels > label_cap))
        return e->fail(PF_ERR_CAPACITY, "need %llu results / %llu bytes / %llu labels", (unsigned long long)P, (unsigned long long)(P * slot), (unsigned long long)nlabels);
    // parse the query ciphertexts
    CK(fl.qcts.ensure_grow(std::max<size_t>(8, ncts * ctw * 8)));
    uint64_t parms_id[4] = {0, 0, 0, 0};
    std::vector<const uint8_t *> ct_src(ncts);
    std::vector<std::vector<uint8_t>> inflated; // full form of compressed / seeded queries (slow path, host threads); empty = none
    // Seeded requests (what a symmetric-key SEAL client sends: half the bytes) are expanded ON THE DEVICE when every
    // stream of the call is an uncompressed blake2xb-seeded one; PF_SEEDED_HOST=1 keeps the host expansion (A/B, tests)
    bool seeded_dev = ncts > 0 && e->d_err_word && (size_t)L * N % 512 == 0 && !getenv("PF_SEEDED_HOST");
    for (size_t c = 0; c < ncts && seeded_dev; c++)
        seeded_dev = is_seeded_raw_stream(query_cts + ct_offsets[c], (size_t)(ct_offsets[c + 1] - ct_offsets[c]), (uint64_t)N, (uint32_t)L);
    if (seeded_dev) {
        for (size_t c = 0; c < ncts; c++) {
            const uint8_t *src = query_cts + ct_offsets[c];
            ct_src[c] = src;
            if (src[48]) return e->fail(PF_ERR_FORMAT, "query ciphertext %zu is in NTT form; BFV ciphertexts must be in coefficient form", c);
            memcpy(parms_id, src + 16, 32);
            if ((parms_id[0] | parms_id[1] | parms_id[2] | parms_id[3]) && memcmp(parms_id, e->level_pid[L], 32) != 0)
                return e->fail(PF_ERR_FORMAT, "query ciphertext %zu was made for other encryption parameters (parms_id differs from this engine's top data level)", c);
        }
    } else {
        HostTick ht("parse_queries");
        size_t bad = 0;
        int bad_mode = 0;
        const int nr = normalize_ct_streams(query_cts, ct_offsets, ncts, (uint64_t)N, reinterpret_cast<const uint64_t *>(e->h_q.data()), (uint32_t)L,
                                            inflated, host_threads(), &bad, &bad_mode);
        if (nr == -3) return e->fail(PF_ERR_FORMAT, "query ciphertext %zu: compr_mode %d is not supported on this host (none, zlib; zstd needs libzstd.so.1)", bad, bad_mode);
        if (nr == -8) return e->fail(PF_ERR_FORMAT, "query ciphertext %zu is seeded with a PRNG other than blake2xb (unsupported)", bad);
        if (nr) return e->fail(PF_ERR_FORMAT, "query ciphertext %zu: malformed compressed stream (code %d)", bad, nr);
        for (size_t c = 0; c < ncts; c++) {
            const bool norm = !inflated.empty() && !inflated[c].empty();
            const uint8_t *src = norm ? inflated[c].data() : query_cts + ct_offsets[c];
            const size_t len = norm ? inflated[c].size() : (size_t)(ct_offsets[c + 1] - ct_offsets[c]);
            ct_src[c] = src;
            int is_ntt;
            uint64_t cms;
            size_t total;
            const int pr = parse_ct_prefix(e, src, len, &is_ntt, parms_id, &cms, &total);
            if (pr || cms != (uint64_t)L) return e->fail(PF_ERR_FORMAT, "query ciphertext %zu malformed (code %d)", c, pr);
            if (is_ntt) return e->fail(PF_ERR_FORMAT, "query ciphertext %zu is in NTT form; BFV ciphertexts must be in coefficient form", c);
            // seal::Ciphertext::load(context, ...) refuses a ciphertext of another parameter set; so does this (an
            // all-zero parms_id = "not stamped", as hand-packed test streams have it)
            if ((parms_id[0] | parms_id[1] | parms_id[2] | parms_id[3]) && memcmp(parms_id, e->level_pid[L], 32) != 0)
                return e->fail(PF_ERR_FORMAT, "query ciphertext %zu was made for other encryption parameters (parms_id differs from this engine's top data level)", c);
        }
    }
    // Query groups: the H2D of group i+1 (upload stream) and the D2H of group i-1 (copy stream) overlap
    // the compute of group i (engine stream).  With calls pipelined through submit / collect the overlap
    // comes from the neighbouring calls, and one group (whole-batch kernels) is the most efficient.
    static const uint64_t env_groups = getenv("PF_E2E_GROUPS") ? (uint64_t)atoi(getenv("PF_E2E_GROUPS")) : 0;
    const uint64_t want_groups = env_groups ? env_groups : (e->groups_hint ? (uint64_t)e->groups_hint : 4);
    const uint64_t ngroups = std::max<uint64_t>(1, std::min<uint64_t>(nq, std::min<uint64_t>(want_groups, PF_E2E_GROUPS)));
    // group boundaries: with 4 groups the last one is the smallest (its D2H is the only copy nothing hides)
    uint64_t q_end[PF_E2E_GROUPS + 1];
    q_end[0] = 0;
    for (uint64_t gi = 0; gi < ngroups; gi++) {
        static const uint64_t w4[4] = {5, 10, 14, 16};
        q_end[gi + 1] = (ngroups == 4 && nq >= 16) ? nq * w4[gi] / 16 : nq * (gi + 1) / ngroups;
    }
    // From here on copies that touch the caller's buffers are in flight on three streams: if this
    // function fails they are drained before it returns (and the upload arena slot is closed).
    struct DrainOnError {
        pf_engine *e;
        bool armed = true, arena_open = false;
        ~DrainOnError() {
            if (arena_open) arena_end(e);
            if (!armed) return;
            cudaStreamSynchronize(e->upload_stream);
            cudaStreamSynchronize(e->stream);
            cudaStreamSynchronize(e->copy_stream);
        }
    } drain{e};
    if (e->timeline) {
        if (e->next_ticket == 0) cudaEventRecord(e->tl_base, e->upload_stream);
        cudaEventRecord(fl.tl[0], e->upload_stream);
    }
    // Upload.  Uncompressed streams (the fast path) go up as ONE copy per query group — the byte range of
    // the blob that holds the group's streams, SEAL headers included — and strip_headers_kernel moves the
    // ciphertext words (at a byte offset of 113 inside every stream) to their aligned place; 64 separate
    // 512 KB copies reached 37 GB/s alone and crawled next to a response download.  Streams that were
    // inflated on the host are copied one by one.
    rc = arena_begin(e);
    if (rc) return rc;
    drain.arena_open = true;
    const bool raw_path = inflated.empty() && ncts > 0;
    std::vector<u64> rel_off(ncts);
    if (raw_path) {
        const uint64_t lo = ct_offsets[0], hi = ct_offsets[ncts];
        CK(fl.qraw.ensure_grow((size_t)(hi - lo) + 64));
        CK(e->s_ctoff.ensure_grow(ncts * 8));
        if (seeded_dev) CK(e->s_seed.ensure_grow(ncts * SEED_SCRATCH_WORDS * 8));
        for (size_t c = 0; c < ncts; c++) rel_off[c] = ct_offsets[c] - lo + SEAL_CT_HEADER;
        CK(upload_async(e, e->s_ctoff.p, rel_off.data(), ncts * 8)); // engine stream, ahead of the strip kernels
    }
    {
        HostTick ht("enqueue_h2d");
        for (uint64_t gi = 0; gi < ngroups; gi++) {
            const size_t c_lo = q_end[gi] * e->m, c_hi = q_end[gi + 1] * e->m;
            if (raw_path && c_hi > c_lo) {
                const uint64_t b_lo = ct_offsets[c_lo] - ct_offsets[0], b_hi = ct_offsets[c_hi] - ct_offsets[0];
                CK(cudaMemcpyAsync(fl.qraw.as<uint8_t>() + b_lo, query_cts + ct_offsets[c_lo], (size_t)(b_hi - b_lo),
                                   cudaMemcpyHostToDevice, e->upload_stream));
            } else {
                for (size_t c = c_lo; c < c_hi; c++)
                    CK(cudaMemcpyAsync(fl.qcts.as<u64>() + c * ctw, ct_src[c] + SEAL_CT_HEADER, ctw * 8,
                                       cudaMemcpyHostToDevice, e->upload_stream));
            }
            CK(cudaEventRecord(fl.ev_up[gi], e->upload_stream));
        }
    }
    if (e->timeline) {
        cudaEventRecord(fl.tl[1], e->upload_stream);
        cudaEventRecord(fl.tl[2], e->stream);
    }
    CK(fl.out.ensure_grow(std::max<size_t>(8, P * slot)));
    uint8_t *d_blob = fl.out.as<uint8_t>();
    u64 *d_words = reinterpret_cast<u64 *>(d_blob + PF_RESULT_DATA_OFFSET);
    rc = upload_plan(e, pl);
    if (rc) return rc;
    // SEAL stream headers (113 bytes in front of the aligned words of every result), written on the device
    // so that they travel with the one D2H per group.  Result parms_id: caller's override, else the
    // query's at full level, else SEAL's hash of the parameters of the level the results were switched to
    if (P) {
        const uint64_t *out_pid = e->result_pid_set ? e->result_pid : (e->Lr == L ? parms_id : e->level_pid[e->Lr]);
        ResultHeader hd{};
        write_ct_prefix(e, hd.b, 0, out_pid, e->Lr);
        stamp_headers_kernel<<<(unsigned)((P + 127) / 128), 128, 0, e->stream>>>(d_blob, slot, PF_RESULT_PAD, P, hd);
        e->launches++;
    }
    uint64_t pair_lo = 0, q_lo = 0;
    for (uint64_t gi = 0; gi < ngroups; gi++) {
        const uint64_t q_hi = q_end[gi + 1];
        CK(cudaStreamWaitEvent(e->stream, fl.ev_up[gi], 0));
        if (raw_path && seeded_dev && q_hi > q_lo) {
            const size_t c_lo = q_lo * e->m, nc = (q_hi - q_lo) * e->m;
            rc = expand_seeded_on_device(e, fl.qraw.as<uint8_t>(), e->s_ctoff.as<u64>() + c_lo, fl.qcts.as<u64>() + c_lo * ctw,
                                         e->s_seed.as<u64>() + c_lo * SEED_SCRATCH_WORDS, nc);
            if (rc) return rc;
        } else if (raw_path && q_hi > q_lo) {
            const size_t c_lo = q_lo * e->m, nc = (q_hi - q_lo) * e->m;
            strip_headers_kernel<<<dim3((unsigned)((ctw + 255) / 256), (unsigned)nc), 256, 0, e->stream>>>(
                fl.qraw.as<uint8_t>(), e->s_ctoff.as<u64>() + c_lo, fl.qcts.as<u64>() + c_lo * ctw, ctw);
            e->launches++;
        }
        uint64_t pair_hi = pair_lo;
        for (uint64_t q = q_lo; q < q_hi; q++) pair_hi += pl.results_per_query[q];
        rc = search_core(e, q_lo, q_hi - q_lo, fl.qcts.as<u64>() + q_lo * e->m * ctw, pl, d_words, slot / 8);
        if (rc) return rc;
        if (pair_hi > pair_lo) {
            cudaEvent_t ev = e->ev_group[gi & 1];
            CK(cudaEventRecord(ev, e->stream));
            CK(cudaStreamWaitEvent(e->copy_stream, ev, 0));
            CK(cudaMemcpyAsync(out_cts + pair_lo * slot, d_blob + pair_lo * slot, (pair_hi - pair_lo) * slot,
                               cudaMemcpyDeviceToHost, e->copy_stream));
        }
        pair_lo = pair_hi;
        q_lo = q_hi;
    }
    // the flight is complete when the engine stream (rotations of an empty plan included) and the last
    // copy are done
    CK(cudaEventRecord(e->ev_group[0], e->stream));
    if (e->timeline) cudaEventRecord(fl.tl[3], e->stream);
    CK(cudaStreamWaitEvent(e->copy_stream, e->ev_group[0], 0));
    CK(cudaEventRecord(fl.done, e->copy_stream));
    if (e->timeline) cudaEventRecord(fl.tl[5], e->copy_stream);
    arena_end(e);
    drain.arena_open = false;
    drain.armed = false;
    fl.busy = true;
    fl.id = ++e->next_ticket;
    *ticket = fl.id;
    // host-side share of the response, while the GPU works: result offsets and the labels of the owned
    // probed lists (same packing as pf_search_lists_plain)
    HostTick htl("labels+offsets");
    if (result_offsets) {
        for (uint64_t r = 0; r < P; r++) result_offsets[r] = r * slot + PF_RESULT_PAD;
        result_offsets[P] = P * slot;
    }
    if (labels) {
        uint64_t pos = 0;
        for (uint64_t i = 0; i < nq * nprobe; i++) {
            const int64_t l = idx[i];
            if ((uint64_t)l % e->prm.world != e->prm.rank) continue;
            const uint64_t n = (uint64_t)(e->h_list_offsets[l + 1] - e->h_list_offsets[l]);
            memcpy(labels + pos, e->h_ids.data() + e->h_list_offsets[l], n * sizeof(int64_t));
            pos += n;
        }
    }
    return PF_OK;
}

int pf_search_submit(pf_engine *e, uint64_t nq, const uint8_t *query_cts, uint64_t query_bytes, const uint64_t *ct_offsets,
                     const int64_t *idx, uint32_t nprobe, uint8_t *out_cts, uint64_t out_cap, uint64_t *result_offsets,
                     uint64_t max_results, uint64_t *results_per_query, int64_t *labels, uint64_t label_cap,
                     uint64_t *list_sizes, uint64_t *probed_sizes, pf_search_stats *stats, uint64_t *ticket) {
    return submit_search(e, nq, query_cts, query_bytes, ct_offsets, idx, nprobe, out_cts, out_cap, result_offsets,
                         max_results, results_per_query, labels, label_cap, list_sizes, probed_sizes, stats, ticket);
}

int pf_search_collect(pf_engine *e, uint64_t ticket) {
    if (!e) return PF_ERR_INVALID;
    HostTick ht("search_collect");
    cudaEvent_t done = nullptr;
    int fi = -1;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        for (int i = 0; i < PF_MAX_FLIGHTS; i++)
            if (e->flights[i].busy && e->flights[i].id == ticket) fi = i;
        if (fi < 0) return e->fail(PF_ERR_STATE, "no search in flight with ticket %llu", (unsigned long long)ticket);
        done = e->flights[fi].done;
        CK(cudaSetDevice(e->prm.device));
    }
    const cudaError_t ce = cudaEventSynchronize(done); // not under the lock: another thread may submit meanwhile
    std::lock_guard<std::mutex> lk(e->mu);
    e->flights[fi].busy = false;
    if (ce != cudaSuccess) return e->fail(PF_ERR_CUDA, "search %llu failed: %s", (unsigned long long)ticket, cudaGetErrorString(ce));
    if (e->timeline) {
        float t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 6; i++)
            if (i != 4) cudaEventElapsedTime(&t[i], e->tl_base, e->flights[fi].tl[i]);
        fprintf(stderr, "[pf timeline] ticket %3llu  upload %8.3f..%8.3f  compute %8.3f..%8.3f  download ..%8.3f ms\n",
                (unsigned long long)ticket, t[0], t[1], t[2], t[3], t[5]);
    }
    return check_device_error(e);
}

int pf_search_set_groups(pf_engine *e, uint32_t groups) {
    if (!e || groups > PF_E2E_GROUPS) return PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    e->groups_hint = (int)groups;
    return PF_OK;
}

int pf_search_lists_encrypted(pf_engine *e, uint64_t nq, const uint8_t *query_cts, uint64_t query_bytes,
                              const uint64_t *ct_offsets, const int64_t *idx, uint32_t nprobe, uint8_t *out_cts,
                              uint64_t out_cap, uint64_t *result_offsets, uint64_t max_results, uint64_t *results_per_query,
                              int64_t *labels, uint64_t label_cap, uint64_t *list_sizes, uint64_t *probed_sizes,
                              pf_search_stats *stats) {
    uint64_t ticket = 0;
    const int rc = submit_search(e, nq, query_cts, query_bytes, ct_offsets, idx, nprobe, out_cts, out_cap, result_offsets,
                                 max_results, results_per_query, labels, label_cap, list_sizes, probed_sizes, stats, &ticket);
    if (rc) return rc;
    return pf_search_collect(e, ticket);
}

int pf_host_register(pf_engine *e, void *ptr, size_t bytes) {
    if (!e || !ptr || !bytes) return PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->prm.device));
    CK(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return PF_OK;
}

int pf_host_unregister(pf_engine *e, void *ptr) {
    if (!e || !ptr) return PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->prm.device));
    CK(cudaHostUnregister(ptr));
    return PF_OK;
}

// ---- primitives -----------------------------------------------------------------------------------------
static int ntt_host(pf_engine *e, uint64_t *polys, uint64_t npoly, const int32_t *limb, bool inverse) {
    if (!e || !polys || !limb) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->prm.device));
    const size_t N = e->N;
    CK(e->s_tmp.ensure_grow(std::max<size_t>(8, npoly * N * 8)));
    CK(cudaMemcpyAsync(e->s_tmp.p, polys, npoly * N * 8, cudaMemcpyHostToDevice, e->stream));
    for (uint64_t i = 0; i < npoly; i++) {
        const int li = limb[i] < 0 ? e->k : limb[i];
        if (li > e->k) return e->fail(PF_ERR_INVALID, "limb index %d out of range", limb[i]);
        NttParams p{};
        p.in = p.out = e->s_tmp.as<u64>() + i * N;
        p.mod_map[0] = li;
        launch_ntt(e, NTT_IN_PLAIN, inverse, p, dim3(1, 1, 1));
    }
    CK(cudaMemcpyAsync(polys, e->s_tmp.p, npoly * N * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaGetLastError());
    return PF_OK;
}

int pf_ntt_forward(pf_engine *e, uint64_t *polys, uint64_t npoly, const int32_t *limb) { return ntt_host(e, polys, npoly, limb, false); }
int pf_ntt_inverse(pf_engine *e, uint64_t *polys, uint64_t npoly, const int32_t *limb) { return ntt_host(e, polys, npoly, limb, true); }

static int ct_ntt_host(pf_engine *e, uint64_t *cts, uint64_t ncts, bool inverse) {
    if (!e || !cts) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->prm.device));
    const size_t ctw = (size_t)2 * e->L * e->N;
    CK(e->s_tmp.ensure_grow(std::max<size_t>(8, ncts * ctw * 8)));
    CK(cudaMemcpyAsync(e->s_tmp.p, cts, ncts * ctw * 8, cudaMemcpyHostToDevice, e->stream));
    ntt_limbs(e, e->s_tmp.as<u64>(), e->s_tmp.as<u64>(), 2 * ncts, inverse);
    CK(cudaMemcpyAsync(cts, e->s_tmp.p, ncts * ctw * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaGetLastError());
    return PF_OK;
}
int pf_ct_to_ntt(pf_engine *e, uint64_t *cts, uint64_t ncts) { return ct_ntt_host(e, cts, ncts, false); }
int pf_ct_from_ntt(pf_engine *e, uint64_t *cts, uint64_t ncts) { return ct_ntt_host(e, cts, ncts, true); }

int pf_ct_pt_mac(pf_engine *e, const uint64_t *cts, const uint64_t *pts, uint32_t K, const uint64_t *addend, uint64_t *out) {
    if (!e || !cts || !pts || !out || !K) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    if (K != e->K) return e->fail(PF_ERR_INVALID, "K must equal the engine layout's K = %u", e->K);
    CK(cudaSetDevice(e->prm.device));
    const int L = e->L, N = e->N;
    const size_t ctw = (size_t)2 * L * N, ptw = (size_t)L * N;
    DevBuf dc, dp, dn, dout, dch, dpb, dpo;
    CK(dc.ensure(K * ctw * 8));
    CK(dp.ensure(K * ptw * 8));
    CK(dn.ensure(ptw * 8));
    CK(dout.ensure(ctw * 8));
    CK(dch.ensure(sizeof(MacChunk)));
    CK(dpb.ensure(8));
    CK(dpo.ensure(4));
    CK(cudaMemsetAsync(dpo.p, 0, 4, e->stream));
    CK(cudaMemcpyAsync(dc.p, cts, K * ctw * 8, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(dp.p, pts, K * ptw * 8, cudaMemcpyHostToDevice, e->stream));
    if (addend) CK(cudaMemcpyAsync(dn.p, addend, ptw * 8, cudaMemcpyHostToDevice, e->stream));
    if (!e->mac_wide) {
        split_convert(e, dc.as<u64>(), dc.as<u64>(), K * ctw, true);
        split_convert(e, dp.as<u64>(), dp.as<u64>(), K * ptw, true);
    }
    const MacChunk ch{0, 0, 1, 0};
    const long long pb = 0;
    CK(cudaMemcpyAsync(dch.p, &ch, sizeof(ch), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(dpb.p, &pb, 8, cudaMemcpyHostToDevice, e->stream));
    MacParams mp{};
    mp.rot = dc.as<u64>();
    mp.diag = dp.as<u64>();
    mp.norm = addend ? dn.as<u64>() : nullptr;
    mp.diag_sb = (long long)(K * ptw);
    mp.diag_sk = (long long)ptw;
    mp.norm_sb = (long long)ptw;
    mp.chunks = dch.as<MacChunk>();
    mp.pair_block = dpb.as<long long>();
    mp.pair_out = dpo.as<int>();
    mp.out = dout.as<u64>();
    mp.out_stride = (long long)ctw;
    mp.mods = e->d_mods.as<DevModulus>();
    mp.K = (int)K;
    mp.L = L;
    mp.N = N;
    launch_mac(e, mp, 1);
    CK(cudaMemcpyAsync(out, dout.p, ctw * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaGetLastError());
    return PF_OK;
}

int pf_ct_add(pf_engine *e, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    if (!e || !a || !b || !out) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->prm.device));
    const size_t ctw = (size_t)2 * e->L * e->N;
    CK(e->s_tmp.ensure_grow(3 * ctw * 8));
    u64 *da = e->s_tmp.as<u64>(), *db = da + ctw, *dc = db + ctw;
    CK(cudaMemcpyAsync(da, a, ctw * 8, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(db, b, ctw * 8, cudaMemcpyHostToDevice, e->stream));
    ct_add_kernel<<<dim3(e->N / 256, 2 * e->L), 256, 0, e->stream>>>(da, db, dc, e->d_mods.as<DevModulus>(), e->L, e->N);
    e->launches++;
    CK(cudaMemcpyAsync(out, dc, ctw * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaGetLastError());
    return PF_OK;
}

int pf_rotate_rows(pf_engine *e, const uint64_t *ct, int step, uint64_t *out) {
    if (!e || !ct || !out) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->prm.device));
    const GaloisKey *gk = find_key(e, step);
    if (!gk) return e->fail(PF_ERR_STATE, "no Galois key for step %d", step);
    const int L = e->L, N = e->N;
    const size_t ctw = (size_t)2 * L * N;
    CK(e->s_tmp.ensure_grow(3 * ctw * 8));
    u64 *din = e->s_tmp.as<u64>(), *dntt = din + ctw, *dout = dntt + ctw;
    CK(cudaMemcpyAsync(din, ct, ctw * 8, cudaMemcpyHostToDevice, e->stream));
    ntt_limbs(e, din, dntt, 2, false);
    std::vector<RotJob> jobs(1);
    jobs[0].c1_coef = din + (size_t)L * N;
    jobs[0].c0_ntt = dntt;
    jobs[0].key = gk->key.as<u64>();
    jobs[0].perm = gk->perm.as<u32>();
    jobs[0].out = dout;
    jobs[0].einv = gk->einv;
    int rc = hoist_digits(e, din, 1, ctw);
    if (rc) return rc;
    jobs[0].D = e->s_hoistD.as<u64>();
    jobs[0].KM = gk->km.as<u64>();
    jobs[0].flag = e->s_flags.as<int>();
    rc = run_rot_jobs(e, jobs, false);
    if (rc) return rc;
    ntt_limbs(e, dout, dout, 2, true); // SEAL returns BFV ciphertexts in coefficient form
    CK(cudaMemcpyAsync(out, dout, ctw * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaGetLastError());
    return PF_OK;
}

int pf_rotate_query_set(pf_engine *e, const uint64_t *cts, int chain, uint64_t *rot) {
    if (!e || !cts || !rot) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->prm.device));
    const size_t ctw = (size_t)2 * e->L * e->N;
    CK(e->s_qcts.ensure_grow(e->m * ctw * 8));
    CK(e->s_rot.ensure_grow(e->K * ctw * 8));
    CK(cudaMemcpyAsync(e->s_qcts.p, cts, e->m * ctw * 8, cudaMemcpyHostToDevice, e->stream));
    int rc = build_rotated_sets(e, e->s_qcts.as<u64>(), 1, e->s_rot.as<u64>(), chain, true);
    if (rc) return rc;
    CK(cudaMemcpyAsync(rot, e->s_rot.p, e->K * ctw * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaGetLastError());
    return PF_OK;
}

int pf_batch_encode(pf_engine *e, const uint64_t *values, uint64_t *plain) {
    if (!e || !values || !plain) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->prm.device));
    const size_t N = e->N;
    CK(e->s_tmp.ensure_grow(2 * N * 8));
    u64 *dv = e->s_tmp.as<u64>(), *dp = dv + N;
    CK(cudaMemcpyAsync(dv, values, N * 8, cudaMemcpyHostToDevice, e->stream));
    slot_scatter_kernel<<<(unsigned)(N / 256), 256, 0, e->stream>>>(dv, e->d_inv_index_map.as<u32>(), dp);
    e->launches++;
    NttParams p{};
    p.in = p.out = dp;
    p.mod_map[0] = e->k;
    launch_ntt(e, NTT_IN_PLAIN, true, p, dim3(1, 1, 1));
    CK(cudaMemcpyAsync(plain, dp, N * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaGetLastError());
    return PF_OK;
}

int pf_encode_block(pf_engine *e, const int32_t *xs, uint32_t nvec, uint64_t *diag, uint64_t *norm) {
    if (!e || !xs || !diag || !norm) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    if (nvec > e->C) return e->fail(PF_ERR_INVALID, "block holds at most %u vectors", e->C);
    // a one-list, one-block index built through the same kernels as pf_load_index
    const u32 d = e->d;
    std::vector<float> v((size_t)std::max<u32>(nvec, 1) * d, 0.f), cent(d, 0.f);
    for (size_t i = 0; i < (size_t)nvec * d; i++) v[i] = (float)xs[i];
    std::vector<int64_t> ids(std::max<u32>(nvec, 1)), off{0, (int64_t)nvec};
    for (u32 i = 0; i < nvec; i++) ids[i] = i;
    pf_params p2 = e->prm;
    p2.rank = 0;
    p2.world = 1;
    pf_engine *tmp = nullptr;
    int rc = pf_engine_create(&p2, &tmp);
    if (rc) return e->fail(rc, "%s", pf_last_error(nullptr));
    rc = pf_load_index(tmp, 1, cent.data(), off.data(), ids.data(), v.data());
    if (rc == PF_OK && !tmp->d_diag.p) rc = tmp->fail(PF_ERR_INVALID, "vectors must be integers in [0,255]");
    if (rc == PF_OK && nvec) {
        if (!tmp->mac_wide) split_convert(tmp, tmp->d_diag.as<u64>(), tmp->d_diag.as<u64>(), tmp->diag_block_words, false);
        cudaStreamSynchronize(tmp->stream);
        cudaMemcpy(diag, tmp->d_diag.p, tmp->diag_block_words * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(norm, tmp->d_norm.p, tmp->norm_block_words * 8, cudaMemcpyDeviceToHost);
    } else if (rc == PF_OK) {
        memset(diag, 0, tmp->diag_block_words * 8);
        memset(norm, 0, tmp->norm_block_words * 8);
    }
    if (rc) e->fail(rc, "%s", pf_last_error(tmp));
    pf_engine_destroy(tmp);
    return rc;
}

size_t pf_ct_serialized_size(pf_engine *e) { return e ? SEAL_CT_HEADER + (size_t)2 * e->L * e->N * 8 : 0; }
size_t pf_result_slot_size(pf_engine *e) { return e ? PF_RESULT_DATA_OFFSET + (size_t)2 * e->Lr * e->N * 8 : 0; }
size_t pf_result_serialized_size(pf_engine *e) { return e ? SEAL_CT_HEADER + (size_t)2 * e->Lr * e->N * 8 : 0; }
int pf_set_result_parms_id(pf_engine *e, const uint64_t parms_id[4]) {
    if (!e || !parms_id) return PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    memcpy(e->result_pid, parms_id, 32);
    e->result_pid_set = true;
    return PF_OK;
}

int pf_ct_serialize(pf_engine *e, const uint64_t *ct, int is_ntt, uint8_t *out, size_t cap, size_t *written) {
    if (!e || !ct || !out) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    const size_t need = pf_ct_serialized_size(e);
    if (written) *written = need;
    if (cap < need) return e->fail(PF_ERR_CAPACITY, "need %zu bytes", need);
    const uint64_t pid[4] = {0, 0, 0, 0};
    write_ct_prefix(e, out, is_ntt, pid, e->L);
    memcpy(out + SEAL_CT_HEADER, ct, need - SEAL_CT_HEADER);
    return PF_OK;
}

int pf_seal_ct_expand_device(pf_engine *e, const uint8_t *in, size_t len, uint64_t *ct_words, size_t cap_words) {
    if (!e || !in || !ct_words) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->mu);
    int rc = check_device_error(e);
    if (rc) return rc;
    const size_t ctw = (size_t)2 * e->L * e->N;
    if (cap_words < ctw) return e->fail(PF_ERR_CAPACITY, "need %zu words", ctw);
    if ((size_t)e->L * e->N % 512 || !e->d_err_word) return e->fail(PF_ERR_STATE, "device-side expansion unavailable for these parameters");
    if (!is_seeded_raw_stream(in, len, (uint64_t)e->N, (uint32_t)e->L))
        return e->fail(PF_ERR_FORMAT, "not an uncompressed blake2xb-seeded ciphertext stream of this engine's top level");
    CK(cudaSetDevice(e->prm.device));
    DevBuf raw, out, off, scratch;
    CK(raw.ensure(len + 64));
    CK(out.ensure(ctw * 8));
    CK(off.ensure(8));
    CK(scratch.ensure(SEED_SCRATCH_WORDS * 8));
    const u64 data_off = SEAL_CT_HEADER;
    CK(cudaMemcpyAsync(raw.p, in, len, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(off.p, &data_off, 8, cudaMemcpyHostToDevice, e->stream));
    rc = expand_seeded_on_device(e, raw.as<uint8_t>(), off.as<u64>(), out.as<u64>(), scratch.as<u64>(), 1);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ct_words, out.p, ctw * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return check_device_error(e);
}

int pf_ct_deserialize(pf_engine *e, const uint8_t *in, size_t len, uint64_t *ct, size_t cap_words, int *limbs,
                      int *is_ntt, size_t *consumed) {
    if (!e || !in || !ct) return e ? e->fail(PF_ERR_INVALID, "null argument") : PF_ERR_INVALID;
    int ntt;
    uint64_t pid[4], cms;
    size_t total, zconsumed = 0;
    std::vector<uint8_t> plain;
    const int zr = inflate_seal_stream(in, len, plain, &zconsumed, SEAL_CT_HEADER + (size_t)2 * e->L * e->N * 8);
    if (zr < 0) return e->fail(PF_ERR_FORMAT, "malformed compressed ciphertext (code %d)", zr);
    if (zr == 0) {
        in = plain.data();
        len = plain.size();
    }
    // a seeded stream (top level only: that is where symmetric encryptions are made) is expanded first
    std::vector<uint8_t> full;
    size_t seeded_consumed = 0;
    if (len >= SEAL_CT_HEADER && in[5] == 0) {
        uint64_t stream_total;
        memcpy(&stream_total, in + 8, 8);
        const int er = expand_seeded_stream(in, len, (uint64_t)e->N, reinterpret_cast<const uint64_t *>(e->h_q.data()), (uint32_t)e->L, full);
        if (er == -8) return e->fail(PF_ERR_FORMAT, "ciphertext is seeded with a PRNG other than blake2xb (unsupported)");
        if (er == 0) {
            seeded_consumed = (size_t)stream_total;
            in = full.data();
            len = full.size();
        }
    }
    const int pr = parse_ct_prefix(e, in, len, &ntt, pid, &cms, &total);
    if (pr || cms < 1 || cms > (uint64_t)e->L) return e->fail(PF_ERR_FORMAT, "malformed ciphertext (code %d)", pr);
    if ((total - SEAL_CT_HEADER) / 8 > cap_words) return e->fail(PF_ERR_CAPACITY, "need %zu words", (total - SEAL_CT_HEADER) / 8);
    memcpy(ct, in + SEAL_CT_HEADER, total - SEAL_CT_HEADER);
    if (limbs) *limbs = (int)cms;
    if (is_ntt) *is_ntt = ntt;
    if (consumed) *consumed = zr == 0 ? zconsumed : (seeded_consumed ? seeded_consumed : total);
    return PF_OK;
}

} // extern "C"
