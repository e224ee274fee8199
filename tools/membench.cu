// HBM microbenchmark: what a trivial streaming kernel achieves for the traffic mixes of the à trous path.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/membench tools/membench.cu && tools/membench
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

__global__ void k_read(const float4 *a, float *sink, size_t n) {
    float acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = a[i];
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 123.456f) *sink = acc;
}
__global__ void k_write(float4 *a, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        a[i] = make_float4(1, 2, 3, 4);
}
__global__ void k_copy(const float4 *a, float4 *b, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
template <int CS>
__global__ void k_r1w2(const float4 *a, float4 *b, float4 *c, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = a[i];
        float4 u = make_float4(v.x * 0.5f, v.y * 0.5f, v.z * 0.5f, v.w * 0.5f);
        b[i] = u;
        float4 w = make_float4(v.x - u.x, v.y - u.y, v.z - u.z, v.w - u.w);
        if (CS) __stcs(c + i, w); else c[i] = w;
    }
}

int main() {
    const size_t plane = 4096ull * 4096ull;      // floats
    const size_t n4 = plane / 4;
    float *A, *B, *C, *sink;
    cudaMalloc(&A, plane * 4); cudaMalloc(&B, plane * 4); cudaMalloc(&C, plane * 4 * 11); cudaMalloc(&sink, 4);
    cudaMemset(A, 0, plane * 4); cudaMemset(B, 0, plane * 4); cudaMemset(C, 0, plane * 4 * 11);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char *name, double bytes, auto fn) {
        for (int i = 0; i < 3; ++i) fn(i);
        cudaDeviceSynchronize();
        const int reps = 50;
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) fn(i);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-44s %8.2f us/launch  %7.1f GB/s\n", name, ms / reps * 1e3, bytes / (ms / reps * 1e-3) / 1e9);
    };
    for (int grid : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
        printf("--- grid %d x 256\n", grid);
        // working sets cycle through the 11-plane buffer so nothing is L2 resident unless stated
        timeit("read  64MiB (rotating planes)", plane * 4.0, [&](int i) { k_read<<<grid, 256>>>((float4 *)(C + (i % 11) * plane), sink, n4); });
        timeit("write 64MiB (rotating planes)", plane * 4.0, [&](int i) { k_write<<<grid, 256>>>((float4 *)(C + (i % 11) * plane), n4); });
        timeit("copy  64MiB->64MiB (rotating)", plane * 8.0, [&](int i) { k_copy<<<grid, 256>>>((float4 *)(C + (i % 5) * plane), (float4 *)(C + (5 + i % 5) * plane), n4); });
        timeit("1R:2W rotating, plain stores", plane * 12.0, [&](int i) { k_r1w2<0><<<grid, 256>>>((float4 *)(C + (i % 3) * plane), (float4 *)(C + (3 + i % 3) * plane), (float4 *)(C + (6 + i % 5) * plane), n4); });
        // the transform's pattern: c ping-pong A<->B (just written by the previous launch), w streams to 11 planes
        timeit("1R:2W ping-pong c + streamed w (plain)", plane * 12.0, [&](int i) { k_r1w2<0><<<grid, 256>>>((float4 *)((i & 1) ? B : A), (float4 *)((i & 1) ? A : B), (float4 *)(C + (i % 11) * plane), n4); });
        timeit("1R:2W ping-pong c + streamed w (__stcs)", plane * 12.0, [&](int i) { k_r1w2<1><<<grid, 256>>>((float4 *)((i & 1) ? B : A), (float4 *)((i & 1) ? A : B), (float4 *)(C + (i % 11) * plane), n4); });
    }
    // big copy like the driver's peak measurement (1 GiB -> beyond L2)
    float *X, *Y; const size_t big = 1ull << 28;  // floats = 1 GiB
    cudaMalloc(&X, big * 4); cudaMalloc(&Y, big * 4); cudaMemset(X, 0, big * 4);
    timeit("copy 1GiB->1GiB kernel", big * 8.0, [&](int) { k_copy<<<148 * 16, 256>>>((float4 *)X, (float4 *)Y, big / 4); });
    timeit("cudaMemcpyAsync 1GiB D2D", big * 8.0, [&](int) { cudaMemcpyAsync(Y, X, big * 4, cudaMemcpyDeviceToDevice); });
    timeit("write 1GiB", big * 4.0, [&](int) { k_write<<<148 * 16, 256>>>((float4 *)Y, big / 4); });
    timeit("read 1GiB", big * 4.0, [&](int) { k_read<<<148 * 16, 256>>>((float4 *)X, sink, big / 4); });
    return 0;
}
