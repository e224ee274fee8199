// Microbenchmark (GPU box): issue/pipe throughput of scalar FFMA vs packed fma.rn.f32x2 (FFMA2) on sm_100a, alone and
// interleaved with MUFU.EX2, to decide whether the bilateral kernel should use packed fp32 arithmetic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fma2bench tools/fma2bench.cu && tools/fma2bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int MODE>
__global__ void __launch_bounds__(256) bench(float *out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    unsigned long long X[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) X[i] = pk(x[2 * i], x[2 * i + 1]);
    const unsigned long long A = pk(a, a), B = pk(b, b);
    float m[4] = {0.1f, 0.2f, 0.3f, 0.4f};
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {  // 16 scalar FFMA
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
        } else if (MODE == 1) {  // 8 FFMA2 = 16 fp32 FMAs
#pragma unroll
            for (int i = 0; i < 8; ++i) X[i] = fma2(X[i], A, B);
        } else if (MODE == 2) {  // 16 scalar FFMA + 2 MUFU
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
            m[0] = ex2f(m[0]); m[1] = ex2f(m[1]);
        } else if (MODE == 3) {  // 8 FFMA2 + 2 MUFU
#pragma unroll
            for (int i = 0; i < 8; ++i) X[i] = fma2(X[i], A, B);
            m[0] = ex2f(m[0]); m[1] = ex2f(m[1]);
        } else if (MODE == 4) {  // 4 MUFU only
#pragma unroll
            for (int i = 0; i < 4; ++i) m[i] = ex2f(m[i]);
        }
    }
    float s = m[0] + m[1] + m[2] + m[3];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)(X[i] & 0xffffffffu)) + __uint_as_float((unsigned)(X[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char *name, double fma_per_iter, double mufu_per_iter) {
    float *out;
    const int blocks = 148 * 8, iters = 20000;
    cudaMalloc(&out, blocks * 256 * sizeof(float));
    bench<MODE><<<blocks, 256>>>(out, 100, 1.0001f, 0.5f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<MODE><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double thr = (double)blocks * 256 * iters;
    printf("%-28s %8.3f ms  %7.2f Gfma/s (%.1f fma/clk/SM @1.965GHz)  %7.2f Gmufu/s (%.1f /clk/SM)\n", name, ms,
           thr * fma_per_iter / ms / 1e6, thr * fma_per_iter / ms / 1e6 / 148 / 1.965,
           thr * mufu_per_iter / ms / 1e6, thr * mufu_per_iter / ms / 1e6 / 148 / 1.965);
    cudaFree(out);
}

int main() {
    run<0>("16 FFMA", 16, 0);
    run<1>("8 FFMA2", 16, 0);
    run<2>("16 FFMA + 2 MUFU.EX2", 16, 2);
    run<3>("8 FFMA2 + 2 MUFU.EX2", 16, 2);
    run<4>("4 MUFU.EX2", 0, 4);
    return 0;
}
