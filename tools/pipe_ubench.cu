// pipe_ubench.cu — instruction-throughput microbenchmark for the pipes the MAC / NTT kernels lean on:
// IMAD.WIDE.U32 (fmaheavy), IADD3 (alu), DFMA (fp64), and IMAD.WIDE + DFMA interleaved.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_ubench pipe_ubench.cu
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;

template <int MODE>
__global__ void k(u64 *out, int iters, unsigned a0, unsigned b0, double da, double db) {
    u64 acc[8];
    double dacc[8];
    unsigned a = a0 + threadIdx.x, b = b0 + blockIdx.x;
    double x = da + threadIdx.x, y = db;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        acc[i] = i;
        dacc[i] = i;
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0 || MODE == 3) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a), "r"(b));
            if (MODE == 1) asm volatile("add.u64 %0, %0, %1;" : "+l"(acc[i]) : "l"((u64)a));
            if (MODE == 2 || MODE == 3) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(dacc[i]) : "d"(x), "d"(y));
            if (MODE == 4) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(*(unsigned *)&acc[i]) : "r"(a), "r"(b));
        }
    }
    u64 s = 0;
    double ds = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        s += acc[i];
        ds += dacc[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (u64)ds;
}

template <int MODE>
void run(const char *name, int ops_per_iter) {
    u64 *out;
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    cudaMalloc(&out, (size_t)blocks * threads * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 16, 3, 5, 1.5, 2.5);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 3, 5, 1.5, 2.5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * iters * ops_per_iter;
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %8.3f ms  %7.2f Tops/s  ~%5.1f lanes/clk/SM (at %.0f MHz nominal)\n", name, ms, ops / ms * 1e-9,
           ops / (ms * 1e-3) / 148.0 / (clk * 1e3), clk * 1e-3);
    cudaFree(out);
}

int main() {
    run<0>("IMAD.WIDE.U32 (64-bit acc)", 8);
    run<4>("IMAD.LO (32-bit)", 8);
    run<1>("IADD 64-bit (2 IADD3)", 8);
    run<2>("DFMA", 8);
    run<3>("IMAD.WIDE + DFMA interleaved", 16);
    return 0;
}
