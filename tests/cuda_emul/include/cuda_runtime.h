// tests/cuda_emul/include/cuda_runtime.h — TEST INFRASTRUCTURE, never part of the product.
//
// A minimal CPU interpreter of the CUDA programming model, just large enough to run the product's own
// kernel sources (prefhetch_b200/csrc/*.cu, *.cuh, rewritten by tests/cuda_emul/build_emul.py: launch
// syntax, dynamic shared memory, the dozen inline-PTX statements) on a machine WITHOUT a GPU, so that the
// `-m gpu` parity tests can be dry-run against the oracle before they reach a B200.  What it is for:
// control flow, indexing, buffer sizing, the host side of pf_engine.cu and the integer / FP64 arithmetic
// of the kernels (IEEE-754 doubles and fma() behave the same on the host).  What it cannot show: timing,
// memory-model races between warps, anything about sm_100a code generation.
//
// Execution model: one launch at a time (streams are synchronous), blocks distributed over a pool of host
// threads, the threads of one block are fibers of one host thread; __syncthreads() and the warp collectives
// are counted barriers between fibers.  `__shared__` variables are thread_local statics of the host thread
// that runs the block.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <time.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <map>
#include <string>
#include <mutex>
#include <thread>
#include <type_traits>
#include <vector>

#define PF_CUDA_EMUL 1

// ---- qualifiers ---------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types -------------------------------------------------------------------------------------
struct uint3 {
    unsigned x, y, z;
};
struct dim3 {
    unsigned x, y, z;
    constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __attribute__((aligned(8))) uint2 {
    unsigned x, y;
};
struct __attribute__((aligned(16))) uint4 {
    unsigned x, y, z, w;
};
struct __attribute__((aligned(8))) int2 {
    int x, y;
};
struct __attribute__((aligned(16))) int4 {
    int x, y, z, w;
};
struct __attribute__((aligned(8))) float2 {
    float x, y;
};
struct __attribute__((aligned(16))) float4 {
    float x, y, z, w;
};
struct __attribute__((aligned(16))) double2 {
    double x, y;
};
struct __attribute__((aligned(16))) ulonglong2 {
    unsigned long long x, y;
};
struct __attribute__((aligned(16))) longlong2 {
    long long x, y;
};
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }
static inline longlong2 make_longlong2(long long x, long long y) { return longlong2{x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

// ---- runtime API types --------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamDefault = 0, cudaStreamNonBlocking = 1 };
enum { cudaEventDefault = 0, cudaEventBlockingSync = 1, cudaEventDisableTiming = 2 };
enum { cudaHostAllocDefault = 0, cudaHostAllocPortable = 1, cudaHostAllocMapped = 2, cudaHostAllocWriteCombined = 4 };
enum { cudaHostRegisterDefault = 0, cudaHostRegisterPortable = 1, cudaHostRegisterMapped = 2 };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
struct CUstream_st {
    int id;
};
struct CUevent_st {
    double t;
};
typedef CUstream_st *cudaStream_t;
typedef CUevent_st *cudaEvent_t;
struct cudaIpcMemHandle_t {
    char reserved[64];
};

namespace pf_emul {

// ---- fibers: callee-saved register switch (System V x86-64) --------------------------------------------
extern "C" void pf_emul_switch(void **save_sp, void *load_sp);
#if defined(__x86_64__)
asm(R"(
.text
.globl pf_emul_switch
.type pf_emul_switch,@function
pf_emul_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size pf_emul_switch,.-pf_emul_switch
)");
#else
#error "tests/cuda_emul needs x86-64"
#endif

// AddressSanitizer build (build_emul.py --asan): a memcheck of every kernel — device allocations are exact-size
// heap blocks, the dynamic shared memory of a block is an exact-size heap block, and the fiber switches are
// announced to the sanitizer so that its stack bookkeeping follows them.
#if defined(__SANITIZE_ADDRESS__)
#define PF_EMUL_ASAN 1
extern "C" void __sanitizer_start_switch_fiber(void **fake_stack_save, const void *bottom, size_t size);
extern "C" void __sanitizer_finish_switch_fiber(void *fake_stack_save, const void **bottom_old, size_t *size_old);
#define PF_ASAN_START(save, bottom, size) __sanitizer_start_switch_fiber(save, bottom, size)
#define PF_ASAN_FINISH(save, bo, so) __sanitizer_finish_switch_fiber(save, bo, so)
#else
#define PF_ASAN_START(save, bottom, size) ((void)0)
#define PF_ASAN_FINISH(save, bo, so) ((void)0)
#endif

#if defined(PF_EMUL_ASAN)
constexpr size_t kStack = 128 * 1024;   // instrumented frames are larger
#else
constexpr size_t kStack = 96 * 1024;
#endif
constexpr int kMaxThreads = 1024;
constexpr size_t kMaxDynSmem = 232 * 1024;

// Fresh device memory and the dynamic shared memory of a block hold garbage on a GPU; here they hold 0xCD bytes
// (PF_EMUL_POISON=0 turns it off), so that code which relies on zeroes it never wrote fails the parity tests.
inline bool poison_enabled() {
    static const bool on = !(getenv("PF_EMUL_POISON") && atoi(getenv("PF_EMUL_POISON")) == 0);
    return on;
}
inline void *poison(void *p, size_t bytes) {
    if (p && bytes && poison_enabled()) memset(p, 0xCD, bytes);
    return p;
}

struct Barrier {
    int arrived = 0;
    unsigned gen = 0;
};

struct Block {
    const std::function<void()> *body = nullptr;
    int nthreads = 0, live = 0;
    Barrier bar;                  // __syncthreads
    Barrier wbar[kMaxThreads / 32];
    int wlive[kMaxThreads / 32];
    unsigned long long xch[kMaxThreads / 32][2][32];   // warp exchange slots, double-buffered
    unsigned char xpar[kMaxThreads];                    // per-thread shuffle parity
    void *sp[kMaxThreads];
    bool done[kMaxThreads];
    void *sched_sp = nullptr;
    int cur = 0;
    dim3 bdim;
    // sanitizer bookkeeping
    void *fake[kMaxThreads];
    void *sched_fake = nullptr;
    const void *sched_bottom = nullptr;
    size_t sched_size = 0;
};

inline std::atomic<unsigned long long> g_launch_id{0};

struct Worker {
    char *stacks = nullptr;       // kMaxThreads fiber stacks
    char *dyn = nullptr;          // dynamic shared memory of the running block
    char *exact = nullptr;        // sanitizer build: exact-size dynamic shared memory of the current launch
    unsigned long long exact_launch = ~0ull;
    Block blk;
    void ensure() {
        if (stacks) return;
        stacks = (char *)mmap(nullptr, kStack * kMaxThreads, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (stacks == MAP_FAILED) {
            fprintf(stderr, "[cuda_emul] mmap of the fiber stacks failed\n");
            abort();
        }
        if (posix_memalign((void **)&dyn, 1024, kMaxDynSmem)) abort();
    }
};

inline thread_local Worker tl_worker;
inline thread_local uint3 tl_threadIdx, tl_blockIdx;
inline thread_local dim3 tl_blockDim, tl_gridDim;

inline void yield_to_scheduler() {
    Block &b = tl_worker.blk;
    const int me = b.cur;
    PF_ASAN_START(&b.fake[me], b.sched_bottom, b.sched_size);
    pf_emul_switch(&b.sp[me], b.sched_sp);
    PF_ASAN_FINISH(tl_worker.blk.fake[me], nullptr, nullptr);
}

inline void release_if_complete(Barrier &bar, int live) {
    if (live > 0 && bar.arrived >= live) {
        bar.arrived = 0;
        bar.gen++;
    }
}

inline void barrier_wait(Barrier &bar, const int &live) {
    const unsigned g = bar.gen;
    bar.arrived++;
    release_if_complete(bar, live);
    while (bar.gen == g) yield_to_scheduler();
}

extern "C" inline void pf_emul_fiber_entry() {
    Block &b = tl_worker.blk;
    const int me = b.cur;
    PF_ASAN_FINISH(nullptr, &b.sched_bottom, &b.sched_size);
    (*b.body)();
    b.done[me] = true;
    b.live--;
    b.wlive[me / 32]--;
    release_if_complete(b.bar, b.live);           // threads that exited count as arrived
    release_if_complete(b.wbar[me / 32], b.wlive[me / 32]);
    PF_ASAN_START(nullptr, b.sched_bottom, b.sched_size);    // nullptr: this fiber's fake stack is released
    pf_emul_switch(&b.sp[me], b.sched_sp);
    abort();                                       // a finished fiber is never resumed
}

inline void run_block(const std::function<void()> &body, dim3 grid, dim3 block, size_t smem, unsigned bx, unsigned by, unsigned bz) {
    Worker &w = tl_worker;
    w.ensure();
#if defined(PF_EMUL_ASAN)
    // an exact-size block per (worker, launch): a kernel that indexes past the dynamic shared memory it was
    // launched with is reported.  Not per CTA: freed blocks sit in the sanitizer's quarantine, and thousands of
    // 100 KB blocks per launch turn into page-fault time.
    if (w.exact_launch != g_launch_id.load() || !w.exact) {
        free(w.exact);
        w.exact = nullptr;
        if (posix_memalign((void **)&w.exact, 16, smem ? smem : 16)) abort();   // 16: what the kernels may assume
        w.exact_launch = g_launch_id.load();
    }
    char *const pool_dyn = w.dyn;
    w.dyn = w.exact;
#else
    (void)smem;
#endif
    Block &b = w.blk;
    const int T = (int)(block.x * block.y * block.z);
    poison(w.dyn, smem);
    b.body = &body;
    b.nthreads = b.live = T;
    b.bar = Barrier();
    b.bdim = block;
    for (int wi = 0; wi < (T + 31) / 32; wi++) {
        b.wbar[wi] = Barrier();
        b.wlive[wi] = (T - wi * 32) < 32 ? (T - wi * 32) : 32;
    }
    memset(b.xpar, 0, (size_t)T);
    tl_blockIdx = uint3{bx, by, bz};
    tl_blockDim = block;
    tl_gridDim = grid;
    for (int t = 0; t < T; t++) {
        b.done[t] = false;
        char *top = w.stacks + (size_t)(t + 1) * kStack;
        void **sp = (void **)top;
        *--sp = nullptr;                            // keeps the entry frame 16-byte aligned after `ret`
        *--sp = (void *)&pf_emul_fiber_entry;
        for (int r = 0; r < 6; r++) *--sp = nullptr;
        b.sp[t] = (void *)sp;
    }
    int remaining = T;
    long passes = 0;
    // PF_EMUL_ORDER: the order in which the threads of a block are resumed in a scheduling pass — "forward"
    // (default), "reverse", or "random[:seed]" (a fresh permutation every pass).  Results must not depend on it:
    // a kernel that does is missing a barrier (the racecheck of this emulation).
    static const int order_mode = [] {
        const char *e = getenv("PF_EMUL_ORDER");
        return !e ? 0 : (!strncmp(e, "reverse", 7) ? 1 : (!strncmp(e, "random", 6) ? 2 : 0));
    }();
    static const unsigned long long order_seed = [] {
        const char *e = getenv("PF_EMUL_ORDER");
        const char *c = e ? strchr(e, ':') : nullptr;
        return c ? strtoull(c + 1, nullptr, 10) : 12345ull;
    }();
    unsigned long long rng = order_seed * 0x9E3779B97F4A7C15ull + ((unsigned long long)bx << 40) + ((unsigned long long)by << 20) + bz + 1;
    static thread_local int perm[kMaxThreads];
    for (int t = 0; t < T; t++) perm[t] = order_mode == 1 ? T - 1 - t : t;
    while (remaining > 0) {
        int progressed = 0;
        if (order_mode == 2)
            for (int t = T - 1; t > 0; t--) {     // Fisher-Yates with a xorshift generator
                rng ^= rng << 13, rng ^= rng >> 7, rng ^= rng << 17;
                const int j = (int)(rng % (unsigned long long)(t + 1));
                const int tmp = perm[t];
                perm[t] = perm[j], perm[j] = tmp;
            }
        for (int ti = 0; ti < T; ti++) {
            const int t = perm[ti];
            if (b.done[t]) continue;
            b.cur = t;
            const unsigned tx = (unsigned)t % block.x, ty = ((unsigned)t / block.x) % block.y, tz = (unsigned)t / (block.x * block.y);
            tl_threadIdx = uint3{tx, ty, tz};
            PF_ASAN_START(&b.sched_fake, w.stacks + (size_t)t * kStack, kStack);
            pf_emul_switch(&b.sched_sp, b.sp[t]);
            PF_ASAN_FINISH(b.sched_fake, nullptr, nullptr);
            if (b.done[t]) {
                remaining--;
                progressed++;
            }
        }
        if (++passes > 50000000L) {
            fprintf(stderr, "[cuda_emul] a block did not finish: deadlocked barrier?\n");
            abort();
        }
        (void)progressed;
    }
#if defined(PF_EMUL_ASAN)
    w.dyn = pool_dyn;
#endif
}

// ---- the block pool ---------------------------------------------------------------------------------------
struct Pool {
    std::mutex launch_mu;          // one launch at a time
    std::mutex mu;
    std::condition_variable cv, cv_done;
    std::vector<std::thread> workers;
    const std::function<void()> *body = nullptr;
    dim3 grid, block;
    size_t smem_bytes = 0;
    std::atomic<unsigned long long> next{0};
    unsigned long long total = 0;
    unsigned long long epoch = 0;
    int active = 0;
    bool stop = false;

    void work() {
        for (;;) {
            const unsigned long long i = next.fetch_add(1);
            if (i >= total) break;
            const unsigned bx = (unsigned)(i % grid.x), by = (unsigned)((i / grid.x) % grid.y), bz = (unsigned)(i / ((unsigned long long)grid.x * grid.y));
            run_block(*body, grid, block, smem_bytes, bx, by, bz);
        }
    }
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || epoch != seen; });
                if (stop) return;
                seen = epoch;
            }
            work();
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--active == 0) cv_done.notify_all();
            }
        }
    }
    void start() {
        if (!workers.empty()) return;
        int n = (int)std::thread::hardware_concurrency();
        if (const char *s = getenv("PF_EMUL_THREADS")) n = atoi(s);
        if (n < 1) n = 1;
        for (int i = 0; i + 1 < n; i++) {
            workers.emplace_back([this] { loop(); });
            workers.back().detach();
        }
    }
    void run(dim3 g, dim3 b, size_t smem, const std::function<void()> &fn) {
        const unsigned long long nblocks = (unsigned long long)g.x * g.y * g.z;
        const unsigned T = b.x * b.y * b.z;
        if (!nblocks || !T || T > (unsigned)kMaxThreads || smem > kMaxDynSmem || g.y > 65535 || g.z > 65535) {
            fprintf(stderr, "[cuda_emul] invalid launch configuration: grid (%u,%u,%u) block (%u,%u,%u) smem %zu\n", g.x, g.y, g.z, b.x, b.y, b.z, smem);
            last_error = cudaErrorInvalidValue;
            return;
        }
        std::lock_guard<std::mutex> launch_lock(launch_mu);
        start();
        body = &fn;
        grid = g;
        block = b;
        smem_bytes = smem;
        g_launch_id++;
        total = nblocks;
        next.store(0);
        const bool fan_out = nblocks > 1 && !workers.empty();
        if (fan_out) {
            std::lock_guard<std::mutex> lk(mu);
            active = (int)workers.size();
            epoch++;
            cv.notify_all();
        }
        work();
        if (fan_out) {
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&] { return active == 0; });
        }
        launches++;
    }
    std::atomic<int> last_error{0};
    std::atomic<unsigned long long> launches{0};
};
inline Pool &pool() {
    static Pool *p = new Pool();   // leaked on purpose: worker threads outlive static destruction
    return *p;
}

inline dim3 mkdim(dim3 d) { return d; }
template <class I, class = typename std::enable_if<std::is_integral<I>::value>::type>
inline dim3 mkdim(I x) { return dim3((unsigned)x); }

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F &&f) {
    const std::function<void()> fn(std::forward<F>(f));
    pool().run(grid, block, smem, fn);
}

inline void *dyn_smem() { return tl_worker.dyn; }

inline unsigned long long globaltimer() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}

// ---- PF_EMUL_IPC=1: allocations up to PF_EMUL_IPC_MAX_MB (default 256) live in POSIX shared memory, so that
// cudaIpcGetMemHandle / cudaIpcOpenMemHandle work ACROSS PROCESSES (the multi-rank gather protocol of bench.py:
// peer buffers, arrival / ack flags) --------------------------------------------------------------------
struct ShmRec {
    std::string name;
    size_t bytes;
    bool owner;
};
struct ShmTable {
    std::mutex mu;
    std::map<void *, ShmRec> recs;
    unsigned long long seq = 0;
    ~ShmTable() {
        for (auto &kv : recs)
            if (kv.second.owner) shm_unlink(kv.second.name.c_str());
    }
};
inline ShmTable &shm_table() {
    static ShmTable t;
    return t;
}
inline bool ipc_enabled() {
    static const bool on = getenv("PF_EMUL_IPC") && atoi(getenv("PF_EMUL_IPC")) != 0;
    return on;
}
inline size_t ipc_max_bytes() {
    static const size_t mb = getenv("PF_EMUL_IPC_MAX_MB") ? (size_t)atoll(getenv("PF_EMUL_IPC_MAX_MB")) : 256;
    return mb << 20;
}
inline void *shm_alloc(size_t bytes) {
    ShmTable &t = shm_table();
    std::lock_guard<std::mutex> lk(t.mu);
    char name[64];
    snprintf(name, sizeof name, "/pf_emul_%d_%llu", (int)getpid(), t.seq++);
    const int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return nullptr;
    const size_t len = (bytes + 4095) & ~(size_t)4095;
    if (ftruncate(fd, (off_t)len)) {
        close(fd);
        shm_unlink(name);
        return nullptr;
    }
    void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) {
        shm_unlink(name);
        return nullptr;
    }
    t.recs[p] = ShmRec{name, len, true};
    return p;
}
inline bool shm_release(void *p) {
    ShmTable &t = shm_table();
    std::lock_guard<std::mutex> lk(t.mu);
    auto it = t.recs.find(p);
    if (it == t.recs.end()) return false;
    munmap(p, it->second.bytes);
    if (it->second.owner) shm_unlink(it->second.name.c_str());
    t.recs.erase(it);
    return true;
}

// ---- device allocations with red zones (a write past either end aborts at cudaFree) -------------------
constexpr size_t kRed = 256;
struct AllocHdr {
    size_t bytes;
    unsigned long long magic;
};
#if defined(PF_EMUL_ASAN)
inline void *dev_alloc(size_t bytes) {
    if (ipc_enabled() && bytes <= ipc_max_bytes()) return poison(shm_alloc(bytes ? bytes : 1), bytes);
    void *p = nullptr;
    return posix_memalign(&p, 256, bytes ? bytes : 1) ? nullptr : poison(p, bytes);
}
inline void dev_free(void *p) {
    if (p && !shm_release(p)) free(p);
}
#else
inline void *dev_alloc(size_t bytes) {
    if (ipc_enabled() && bytes <= ipc_max_bytes()) return poison(shm_alloc(bytes ? bytes : 1), bytes);
    char *raw = nullptr;
    const size_t total = kRed + bytes + kRed;
    if (posix_memalign((void **)&raw, 256, total ? total : 256)) return nullptr;
    memset(raw, 0xA5, kRed);
    memset(raw + kRed + bytes, 0xA5, kRed);
    AllocHdr h{bytes, 0x50464D454D554C21ull};
    memcpy(raw, &h, sizeof h);
    return poison(raw + kRed, bytes);
}
inline void dev_free(void *p) {
    if (!p || shm_release(p)) return;
    char *raw = (char *)p - kRed;
    AllocHdr h;
    memcpy(&h, raw, sizeof h);
    if (h.magic != 0x50464D454D554C21ull) {
        fprintf(stderr, "[cuda_emul] cudaFree of a pointer cudaMalloc did not return (or the front red zone was overwritten)\n");
        abort();
    }
    for (size_t i = sizeof h; i < kRed; i++)
        if ((unsigned char)raw[i] != 0xA5) {
            fprintf(stderr, "[cuda_emul] write BEFORE a device allocation of %zu bytes\n", h.bytes);
            abort();
        }
    for (size_t i = 0; i < kRed; i++)
        if ((unsigned char)raw[kRed + h.bytes + i] != 0xA5) {
            fprintf(stderr, "[cuda_emul] write PAST a device allocation of %zu bytes (offset +%zu)\n", h.bytes, i);
            abort();
        }
    free(raw);
}
#endif

}  // namespace pf_emul

// ---- built-in variables ----------------------------------------------------------------------------------
#define threadIdx (pf_emul::tl_threadIdx)
#define blockIdx (pf_emul::tl_blockIdx)
#define blockDim (pf_emul::tl_blockDim)
#define gridDim (pf_emul::tl_gridDim)
static const int warpSize = 32;

// ---- synchronisation and warp collectives ----------------------------------------------------------------
static inline void __syncthreads() {
    pf_emul::Block &b = pf_emul::tl_worker.blk;
    pf_emul::barrier_wait(b.bar, b.live);
}
static inline void __syncwarp(unsigned = 0xffffffffu) {
    pf_emul::Block &b = pf_emul::tl_worker.blk;
    const int w = b.cur / 32;
    pf_emul::barrier_wait(b.wbar[w], b.wlive[w]);
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __nanosleep(unsigned ns) {
    timespec ts{0, (long)ns};
    nanosleep(&ts, nullptr);
}

namespace pf_emul {
// every lane of the warp publishes a 64-bit value, then reads the lane it wants (full-mask collectives only)
inline unsigned long long warp_exchange(unsigned long long mine, int src_lane, bool &valid) {
    Block &b = tl_worker.blk;
    const int me = b.cur, w = me / 32, lane = me % 32;
    const int par = b.xpar[me];
    b.xpar[me] ^= 1;
    b.xch[w][par][lane] = mine;
    barrier_wait(b.wbar[w], b.wlive[w]);
    const int lanes = (b.nthreads - w * 32) < 32 ? (b.nthreads - w * 32) : 32;
    valid = src_lane >= 0 && src_lane < lanes;
    return valid ? b.xch[w][par][src_lane] : mine;
}
template <class T>
inline T shfl(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of a type wider than 64 bits");
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    bool valid;
    raw = warp_exchange(raw, src_lane, valid);
    T out;
    memcpy(&out, &raw, sizeof(T));
    return valid ? out : v;
}
inline int lane_id() { return tl_worker.blk.cur % 32; }
}  // namespace pf_emul

template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    const int lane = pf_emul::lane_id();
    return pf_emul::shfl(v, (lane / width) * width + (src % width));
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int mask, int width = 32) {
    (void)width;
    return pf_emul::shfl(v, pf_emul::lane_id() ^ mask);
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32) {
    const int lane = pf_emul::lane_id();
    const int src = lane - (int)delta;
    return pf_emul::shfl(v, (src < (lane / width) * width) ? lane : src);
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
    const int lane = pf_emul::lane_id();
    const int src = lane + (int)delta;
    return pf_emul::shfl(v, (src >= (lane / width + 1) * width) ? lane : src);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    // 32 exchanges would do; one is enough: every lane publishes its predicate, lane order is fixed
    pf_emul::Block &b = pf_emul::tl_worker.blk;
    const int me = b.cur, w = me / 32, lane = me % 32;
    const int par = b.xpar[me];
    b.xpar[me] ^= 1;
    b.xch[w][par][lane] = pred ? 1ull : 0ull;
    pf_emul::barrier_wait(b.wbar[w], b.wlive[w]);
    const int lanes = (b.nthreads - w * 32) < 32 ? (b.nthreads - w * 32) : 32;
    unsigned m = 0;
    for (int l = 0; l < lanes; l++)
        if (b.xch[w][par][l]) m |= 1u << l;
    return m;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) {
    pf_emul::Block &b = pf_emul::tl_worker.blk;
    const int w = b.cur / 32;
    const int lanes = (b.nthreads - w * 32) < 32 ? (b.nthreads - w * 32) : 32;
    const unsigned full = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1u);
    return (__ballot_sync(m, pred) & full) == full;
}

// ---- atomics (blocks run on several host threads) --------------------------------------------------------
template <class T>
static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned *p, int v) { return __atomic_fetch_add(p, (unsigned)v, __ATOMIC_SEQ_CST); }
template <class T>
static inline T atomicMax(T *p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
    }
    return old;
}
template <class T>
static inline T atomicMin(T *p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
    }
    return old;
}
template <class T>
static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T>
static inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T>
static inline T atomicCAS(T *p, T cmp, T v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}

// ---- arithmetic intrinsics (compile with -ffp-contract=off: a*b+c must not fuse on its own) ---------------
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
static inline long long __mul64hi(long long a, long long b) { return (long long)(((__int128)a * b) >> 64); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __fma_rn(double a, double b, double c) { return __builtin_fma(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmaf_rn(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
static inline long long __double2ll_rd(double x) { return (long long)floor(x); }
static inline long long __double2ll_rn(double x) { return (long long)nearbyint(x); }
static inline long long __double2ll_rz(double x) { return (long long)x; }
static inline unsigned long long __double2ull_rd(double x) { return (unsigned long long)floor(x); }
static inline unsigned long long __double2ull_rz(double x) { return (unsigned long long)x; }
static inline double __ll2double_rn(long long x) { return (double)x; }
static inline double __ull2double_rn(unsigned long long x) { return (double)x; }
static inline float __double2float_rn(double x) { return (float)x; }
static inline double __longlong_as_double(long long x) {
    double d;
    memcpy(&d, &x, 8);
    return d;
}
static inline long long __double_as_longlong(double d) {
    long long x;
    memcpy(&x, &d, 8);
    return x;
}
static inline float __int_as_float(int x) {
    float f;
    memcpy(&f, &x, 4);
    return f;
}
static inline int __float_as_int(float f) {
    int x;
    memcpy(&x, &f, 4);
    return x;
}
static inline unsigned __float_as_uint(float f) {
    unsigned x;
    memcpy(&x, &f, 4);
    return x;
}
static inline float __uint_as_float(unsigned x) {
    float f;
    memcpy(&f, &x, 4);
    return f;
}
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
template <class T>
static inline T __ldg(const T *p) { return *p; }
template <class T>
static inline T __ldcs(const T *p) { return *p; }
template <class T>
static inline void __stcs(T *p, T v) { *p = v; }

// CUDA's global min / max overloads (mixed signedness promotes the way nvcc's headers do: to the wider unsigned)
template <class A, class B, class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
static inline typename std::common_type<A, B>::type min(A a, B b) {
    typedef typename std::common_type<A, B>::type R;
    return (R)b < (R)a ? (R)b : (R)a;
}
template <class A, class B, class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
static inline typename std::common_type<A, B>::type max(A a, B b) {
    typedef typename std::common_type<A, B>::type R;
    return (R)a < (R)b ? (R)b : (R)a;
}

// ---- runtime API: synchronous streams, host memory is device memory ---------------------------------------
namespace pf_emul {
inline int device_count() {
    static const int n = getenv("PF_EMUL_DEVICES") ? atoi(getenv("PF_EMUL_DEVICES")) : 1;   // all of them are this host
    return n < 1 ? 1 : n;
}
}  // namespace pf_emul
static inline cudaError_t cudaGetDeviceCount(int *n) {
    *n = pf_emul::device_count();
    return cudaSuccess;
}
static inline cudaError_t cudaSetDevice(int d) { return d >= 0 && d < pf_emul::device_count() ? cudaSuccess : cudaErrorInvalidValue; }
static inline cudaError_t cudaGetDevice(int *d) {
    *d = 0;
    return cudaSuccess;
}
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return pf_emul::pool().last_error.exchange(0); }
static inline cudaError_t cudaPeekAtLastError() { return pf_emul::pool().last_error.load(); }
static inline const char *cudaGetErrorString(cudaError_t e) {
    return e == cudaSuccess ? "no error" : (e == cudaErrorNotSupported ? "operation not supported (CUDA emulation)" : (e == cudaErrorMemoryAllocation ? "out of memory" : "invalid value (CUDA emulation)"));
}
static inline cudaError_t cudaMalloc(void **p, size_t bytes) {
    *p = pf_emul::dev_alloc(bytes);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T>
static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc((void **)p, bytes); }
static inline cudaError_t cudaFree(void *p) {
    pf_emul::dev_free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned) {
    *p = nullptr;
    if (posix_memalign(p, 4096, bytes ? bytes : 4096)) return cudaErrorMemoryAllocation;
    return cudaSuccess;
}
template <class T>
static inline cudaError_t cudaHostAlloc(T **p, size_t bytes, unsigned f) { return cudaHostAlloc((void **)p, bytes, f); }
static inline cudaError_t cudaMallocHost(void **p, size_t bytes) { return cudaHostAlloc(p, bytes, 0); }
static inline cudaError_t cudaFreeHost(void *p) {
    free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned) {
    *d = h;
    return cudaSuccess;
}
template <class T>
static inline cudaError_t cudaHostGetDevicePointer(T **d, void *h, unsigned f) { return cudaHostGetDevicePointer((void **)d, h, f); }
static inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind) {
    if (n) memmove(dst, src, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
    if (n) memmove(dst, src, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void *p, int v, size_t n) {
    if (n) memset(p, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t = nullptr) {
    if (n) memset(p, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) {
    *s = new CUstream_st{1};
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { return cudaStreamCreate(s); }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { return cudaStreamCreate(s); }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) {
    *lo = 0;
    *hi = -1;
    return cudaSuccess;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) {
    delete s;
    return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) {
    *e = new CUevent_st{0.0};
    return cudaSuccess;
}
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
    delete e;
    return cudaSuccess;
}
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
    e->t = (double)pf_emul::globaltimer() * 1e-6;
    return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = (float)(b->t - a->t);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
// CUDA IPC.  PF_EMUL_IPC=1: the handle names the POSIX shared-memory segment behind the allocation and another
// process maps it; otherwise same-process only (the handle carries the pointer).
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) {
    memset(h, 0, sizeof *h);
    if (pf_emul::ipc_enabled()) {
        pf_emul::ShmTable &t = pf_emul::shm_table();
        std::lock_guard<std::mutex> lk(t.mu);
        auto it = t.recs.find(p);
        if (it == t.recs.end() || it->second.name.size() > 47) return cudaErrorInvalidValue;
        h->reserved[0] = 'S';
        memcpy(h->reserved + 1, it->second.name.c_str(), it->second.name.size() + 1);
        memcpy(h->reserved + 56, &it->second.bytes, 8);
        return cudaSuccess;
    }
    h->reserved[0] = 'P';
    memcpy(h->reserved + 8, &p, sizeof p);
    return cudaSuccess;
}
static inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) {
    if (h.reserved[0] == 'P') {
        memcpy(p, h.reserved + 8, sizeof *p);
        return *p ? cudaSuccess : cudaErrorInvalidValue;
    }
    if (h.reserved[0] != 'S') return cudaErrorInvalidValue;
    h.reserved[55] = 0;
    size_t bytes = 0;
    memcpy(&bytes, h.reserved + 56, 8);
    const int fd = shm_open(h.reserved + 1, O_RDWR, 0600);
    if (fd < 0) return cudaErrorInvalidValue;
    void *m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return cudaErrorMemoryAllocation;
    pf_emul::ShmTable &t = pf_emul::shm_table();
    std::lock_guard<std::mutex> lk(t.mu);
    t.recs[m] = pf_emul::ShmRec{std::string(h.reserved + 1), bytes, false};
    *p = m;
    return cudaSuccess;
}
static inline cudaError_t cudaIpcCloseMemHandle(void *p) {
    if (pf_emul::ipc_enabled()) return pf_emul::shm_release(p) ? cudaSuccess : cudaErrorInvalidValue;
    return cudaSuccess;
}
