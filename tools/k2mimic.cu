// Microbenchmark (GPU box): what bounds the inner loop of the fp32 bilateral kernel (K2)?  It replays the kernel's
// per-output-row instruction mix on a register-resident 5x5 window of packed pixel pairs, adding one ingredient of the
// real kernel per mode, and reports the cycles per warp-row and the share of the MUFU pipe (8 cycles per warp MUFU per
// SM sub-partition) they correspond to.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/k2mimic tools/k2mimic.cu && /tmp/k2mimic
//   base: window variance from row statistics (37 packed operations), 4 rcp, the 24 range-weighted taps (sub2, mul2, fma2,
//   2 ex2, add2, fma2 per packed tap); then one memory-side ingredient of the real kernel at a time (see F below).
// It also holds the dispatch probe: N packed FFMA2 (or 2 N scalar FFMA) plus M independent ALU instructions per
// iteration -- a packed instruction holds the issue port for TWO cycles (the ALU instructions add one cycle each on top
// of 2 N instead of hiding under the FMA pipe), i.e. the kernel's issue budget is instructions + packed instructions.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up2(u64 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// F: bit mask of memory-side ingredients: 1 = the new window row comes from shared memory (5 LDS.64; otherwise it is
// synthesised in registers with 5 packed adds), 2 = 2 STG.64, 4 = 2 STS.64 (instead of the global stores), 8 = one mbarrier
// try_wait on a completed barrier per step (all lanes), 16 = the same by lane 0 only + __syncwarp, 32 = 6 local-memory
// loads (the spills of the 96-register kernel).  ARITH: 0 packed fp32x2, 1 scalar, 2 packed with every second exponential
// as a polynomial on the FMA pipe.
// Like the real kernel the step loop is unrolled by 5 so that the rotating window row is a compile-time register
// index; every step brings in a new bottom row, so nothing is loop-invariant.
template <int I> struct IC { static constexpr int value = I; };

template <int ARITH, int F, int THREADS, int BLOCKS>
__global__ void __launch_bounds__(THREADS, BLOCKS) mimic(float *out, int iters, float seed) {
    extern __shared__ float smem[];
    __shared__ unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)));
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(&bar)) : "memory");
    }
    volatile float lmem[8];
    if (F & 32) for (int i = 0; i < 8; ++i) lmem[i] = seed * i;
    u64 X[5][5], SA[5], SB[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
        for (int k = 0; k < 5; ++k) X[i][k] = pk2(seed * (threadIdx.x + 7 * i + k), seed * (threadIdx.x + 3 * i + 2 * k + 1));
        SA[i] = pk2(seed * i, seed * (i + 1));
        SB[i] = pk2(seed * seed * i, seed * seed * (i + 2));
    }
    for (int i = threadIdx.x; i < 8 * 1024; i += blockDim.x) smem[i] = seed * i;
    __syncthreads();
    u64 acc = pk2(seed, seed);
    const float lk[5] = {-4.0f, -2.0f, -1.4150375f, -2.0f, -4.0f};
    const float h[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
    float2 *dst = reinterpret_cast<float2 *>(out) + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned row0 = (unsigned)__cvta_generic_to_shared(smem) + (2 * threadIdx.x % 512) * 4;

    auto step = [&](auto ic, int it) {
        constexpr int I = decltype(ic)::value;  // slot of the new bottom row; the window is slots I+1 .. I+5 (mod 5)
        if (F & 8) {
            unsigned ok;
            do {
                asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}"
                             : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(0u) : "memory");
            } while (!ok);
        }
        if (F & 16) {
            if ((threadIdx.x & 31) == 0) {
                unsigned ok;
                do {
                    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}"
                                 : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(0u) : "memory");
                } while (!ok);
            }
            __syncwarp();
        }
        // the new row
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            if (F & 1) {
                u64 t;
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(t) : "r"(row0 + (unsigned)(((it & 7) * 1024 + 64 * k) * 4)));
                X[I][k] = add2(t, acc);  // (one packed add more than the kernel: keeps the loads from being hoisted)
            } else {
                X[I][k] = add2(X[I][k], acc);
            }
        }
        // row statistics of the new row
        {
            u64 a = 0ull, b = 0ull;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const u64 dl = sub2(X[I][k], X[I][2]), dr = sub2(X[I][4 - k], X[I][2]);
                const u64 sum = add2(dl, dr), sq = fma2(dr, dr, mul2(dl, dl));
                a = fma2(pk2(h[k], h[k]), sum, a);
                b = fma2(pk2(h[k], h[k]), sq, b);
            }
            SA[I] = a;
            SB[I] = b;
        }
        constexpr int RC = (I + 5 - 2) % 5;
        const u64 xc = X[RC][2];
        u64 s1 = 0ull, s2 = 0ull;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            constexpr int dummy = 0;
            const int rs = (I + 1 + i) % 5;
            if (i == 2) {
                s1 = fma2(pk2(-h[i], -h[i]), SA[rs], s1);
                s2 = fma2(pk2(h[i], h[i]), SB[rs], s2);
            } else {
                const u64 dc = sub2(xc, X[rs][2]);
                const u64 t = fma2(pk2(-2.0f, -2.0f), SA[rs], dc);
                const u64 u = fma2(dc, t, SB[rs]);
                s1 = fma2(pk2(h[i], h[i]), sub2(dc, SA[rs]), s1);
                s2 = fma2(pk2(h[i], h[i]), u, s2);
            }
            (void)dummy;
        }
        float v0, v1;
        up2(sub2(s2, mul2(s1, s1)), v0, v1);
        v0 = (v0 <= 0.0f) ? 1e-20f : v0;
        v1 = (v1 <= 0.0f) ? 1e-20f : v1;
        const float h0 = -0.72f * rcpf(fmaxf(v0 * seed, 1e-37f)), h1 = -0.72f * rcpf(fmaxf(v1 * seed, 1e-37f));
        const u64 nhi = pk2(h0, h1);
        u64 num = 0ull, den = pk2(0.14f, 0.14f);
        if (ARITH == 1) {
            float x0, x1, n0 = 0, n1 = 0, d0 = 0.14f, d1 = 0.14f;
            up2(xc, x0, x1);
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    if (i == 2 && k == 2) continue;
                    const int rs = (I + 1 + i) % 5;
                    float t0, t1;
                    up2(X[rs][k], t0, t1);
                    const float e0 = x0 - t0, e1 = x1 - t1;
                    const float g0 = ex2f(fmaf(e0 * e0, h0, lk[i] + lk[k])), g1 = ex2f(fmaf(e1 * e1, h1, lk[i] + lk[k]));
                    d0 += g0; d1 += g1;
                    n0 = fmaf(g0, e0, n0); n1 = fmaf(g1, e1, n1);
                }
            num = pk2(n0, n1);
            den = pk2(d0, d1);
        } else {
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    if (i == 2 && k == 2) continue;
                    const int rs = (I + 1 + i) % 5;
                    const float l = lk[i] + lk[k];
                    const u64 dd = sub2(xc, X[rs][k]);
                    const u64 arg = fma2(mul2(dd, dd), nhi, pk2(l, l));
                    u64 gw;
                    if (ARITH == 2 && ((i * 5 + k) & 1)) {
                        // 2^arg on the FMA pipe: arg = n + f, degree-5 polynomial of 2^f, exponent add
                        const u64 magic = pk2(12582912.0f, 12582912.0f);
                        const u64 t = add2(arg, magic);
                        const u64 f = sub2(arg, sub2(t, magic));
                        u64 pz = fma2(f, pk2(1.3333558e-3f, 1.3333558e-3f), pk2(9.6181291e-3f, 9.6181291e-3f));
                        pz = fma2(f, pz, pk2(5.5504109e-2f, 5.5504109e-2f));
                        pz = fma2(f, pz, pk2(2.4022651e-1f, 2.4022651e-1f));
                        pz = fma2(f, pz, pk2(6.9314718e-1f, 6.9314718e-1f));
                        pz = fma2(f, pz, pk2(1.0f, 1.0f));
                        float p0, p1, t0, t1;
                        up2(pz, p0, p1);
                        up2(t, t0, t1);
                        gw = pk2(__int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23)),
                                 __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23)));
                    } else {
                        float a0, a1;
                        up2(arg, a0, a1);
                        gw = pk2(ex2f(a0), ex2f(a1));
                    }
                    den = add2(den, gw);
                    num = fma2(gw, dd, num);
                }
        }
        float n0, n1, d0, d1, x0, x1;
        up2(num, n0, n1);
        up2(den, d0, d1);
        up2(xc, x0, x1);
        const float c0 = x0 - n0 * rcpf(d0), c1 = x1 - n1 * rcpf(d1);
        if (F & 2) {
            dst[0] = make_float2(c0, c1);
            __stcs(dst + 1, make_float2(x0 - c0, x1 - c1));
        }
        if (F & 4) {
            float2 *sd = reinterpret_cast<float2 *>(smem + (it & 7) * 1024) + threadIdx.x % 256;
            sd[0] = make_float2(c0, c1);
            sd[256] = make_float2(x0 - c0, x1 - c1);
        }
        float ls = 0.0f;
        if (F & 32) {
#pragma unroll
            for (int i = 0; i < 6; ++i) ls += lmem[i];
        }
        acc = pk2(c0 * 1e-6f + ls, c1 * 1e-6f);
    };
#pragma unroll 1
    for (int it = 0; it < iters; it += 5) {
        step(IC<0>{}, it);
        step(IC<1>{}, it + 1);
        step(IC<2>{}, it + 2);
        step(IC<3>{}, it + 3);
        step(IC<4>{}, it + 4);
    }
    float a0, a1;
    up2(acc, a0, a1);
    if (a0 + a1 == 123.456f) out[threadIdx.x] = a0;
}

template <int ARITH, int F, int THREADS, int BLOCKS> void run(const char *name, int mufu_per_iter) {
    float *out;
    const int iters = 4000;  // multiple of 5
    const int blocks = 148 * BLOCKS;
    cudaMalloc(&out, (size_t)blocks * THREADS * 16 + 4096);
    const size_t smem = 8 * 1024 * 4;
    mimic<ARITH, F, THREADS, BLOCKS><<<blocks, THREADS, smem>>>(out, 50, 0.001f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    mimic<ARITH, F, THREADS, BLOCKS><<<blocks, THREADS, smem>>>(out, iters, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    const double clk = 1.965e9;
    const double warps_per_smsp = (double)THREADS / 32 * BLOCKS / 4;
    const double cyc_per_warp_row = ms * 1e-3 * clk / iters / warps_per_smsp;  // SMSP cycles per warp-row
    printf("%-44s %3d thr x %d blk/SM  %8.3f ms  %7.1f cycles/warp-row/SMSP  MUFU pipe %5.1f %%  (%s)\n", name, THREADS, BLOCKS, ms,
           cyc_per_warp_row, 100.0 * mufu_per_iter * 8 / cyc_per_warp_row, cudaGetErrorString(err));
    cudaFree(out);
}


// Dispatch probe: NF independent FFMA2 (or 2 NF scalar FFMA when SCALAR) plus NI independent integer LOP3/IADD per
// iteration.  If a packed instruction held the dispatch port for one cycle only, the integer work would hide under the
// two FMA-pipe cycles each FFMA2 needs.
template <int NF, int NI, bool SCALAR>
__global__ void __launch_bounds__(256, 2) probe(float *out, int iters, float a, int q) {
    u64 X[8];
    float x[16];
    int n[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { X[i] = pk2(threadIdx.x * 0.001f + i, i); n[i] = threadIdx.x + i; }
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    const u64 A = pk2(a, a), B = pk2(0.5f, 0.25f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NF; ++i) {
            if (SCALAR) { x[2 * (i % 8)] = fmaf(x[2 * (i % 8)], a, 0.5f); x[2 * (i % 8) + 1] = fmaf(x[2 * (i % 8) + 1], a, 0.25f); }
            else X[i % 8] = fma2(X[i % 8], A, B);
        }
#pragma unroll
        for (int i = 0; i < NI; ++i) n[i % 8] = __funnelshift_l(n[i % 8], n[i % 8], q);  // one ALU-pipe instruction each
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo, hi; up2(X[i], lo, hi); s += lo + hi + n[i] + x[2 * i] + x[2 * i + 1]; }
    if (s == 123.456f) out[threadIdx.x] = s;
}
template <int NF, int NI, bool SCALAR> void run_probe() {
    float *out;
    cudaMalloc(&out, 4096);
    const int iters = 20000, blocks = 148 * 2;
    probe<NF, NI, SCALAR><<<blocks, 256>>>(out, 10, 1.0001f, 3);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<NF, NI, SCALAR><<<blocks, 256>>>(out, iters, 1.0001f, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9 / iters / 4.0;  // 16 warps per SM = 4 per SMSP: cycles per warp-iteration per SMSP
    printf("probe: %2d %s + %2d integer ops per iteration: %6.1f SMSP cycles per warp-iteration\n", SCALAR ? 2 * NF : NF,
           SCALAR ? "FFMA " : "FFMA2", NI, cyc);
    cudaFree(out);
}

int main() {
    run_probe<8, 0, false>(); run_probe<8, 2, false>(); run_probe<8, 4, false>(); run_probe<8, 6, false>(); run_probe<8, 8, false>(); run_probe<8, 0, true>(); run_probe<8, 4, true>(); run_probe<8, 8, true>(); run_probe<0, 8, false>();
    run<0, 0, 256, 2>("variance + 24 taps + 4 rcp, packed", 52);
    run<0, 0, 128, 3>("  the same, 12 warps per SM", 52);
    run<0, 0, 256, 1>("  the same, 8 warps per SM", 52);
    run<1, 0, 256, 2>("  the same, scalar arithmetic", 52);
    run<2, 0, 256, 2>("  the same, every second exp2 on the FMA pipe", 28);
    run<0, 1, 256, 2>("+ 5 LDS.64", 52);
    run<0, 2, 256, 2>("+ 2 STG.64", 52);
    run<0, 4, 256, 2>("+ 2 STS.64", 52);
    run<0, 8, 256, 2>("+ mbarrier try_wait (all lanes)", 52);
    run<0, 16, 256, 2>("+ mbarrier try_wait (lane 0) + syncwarp", 52);
    run<0, 32, 256, 2>("+ 6 LDL", 52);
    run<0, 1 + 2 + 8 + 32, 256, 2>("+ all of the kernel's (LDS, STG, try_wait, LDL)", 52);
    run<0, 1 + 2 + 8 + 32, 128, 3>("  the same, 12 warps per SM", 52);
    return 0;
}
