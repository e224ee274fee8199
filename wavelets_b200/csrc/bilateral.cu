// K2 -- one scale of the BILATERAL à trous cascade, fused: local variance + range-weighted 25-tap gather + w_s.
//
// Replaces, per scale, watroo/wavelets.py:433-442:  sdev_loc (two dense filter2D, :24-32), the scaling by sigma_b^2
// (:434-436), atrous_convolution with bilateral_variance (np.pad + 24 full-image numexpr passes, :74-105) and the
// subtraction (:442).  One read of c_s, one write of c_{s+1}, one write of w_s.
//
// Maths (x = c_s at the output pixel, x_t the K^2-1 off-centre dilated taps through the symmetric border, k_t = h_i h_j):
//     var = S[x^2] - S[x]^2  ==  sum_t k_t D_t^2 - (sum_t k_t D_t)^2,   D_t = x - x_t      (shift invariance)
//     V   = max(var, 1e-20) * sigma_b^2 * (s+1 if bilateral_scaling)
//     g_t = k_t exp(-D_t^2 / V / 2),  N = k_c + sum_t g_t
//     c_{s+1} = (k_c x + sum_t g_t x_t) / N  ==  x - (sum_t g_t D_t) / N
// The centred form has no catastrophic cancellation, so the fp32 kernel follows the reference's float64 result
// more closely than the reference's own fp32 path does.  Not HBM-bound: 24 exponentials
// per pixel put it on the MUFU/FMA pipes; it shares the TMA row pipeline of K1 (taps rows resident in the ring).
#include "pipeline.cuh"

namespace wb {

struct BilateralParams {
    ScaleParams sp;
    double var_factor;  // sigma_b[s]^2 * (s + 1 if bilateral_scaling else 1)
};

__device__ __forceinline__ float exp2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// g = k * exp(-D^2 / V / 2) with nhalf_inv = -log2(e) / (2 V) (fp32) or -1 / (2 V) (fp64)
__device__ __forceinline__ float range_weight(float k, float d2, float nhalf_inv) { return k * exp2_fast(d2 * nhalf_inv); }
__device__ __forceinline__ double range_weight(double k, double d2, double nhalf_inv) { return k * exp(d2 * nhalf_inv); }
template <typename T> __device__ __forceinline__ T nhalf_inverse(T v);
template <> __device__ __forceinline__ float nhalf_inverse<float>(float v) { return -1.4426950408889634f / (2.0f * v); }
template <> __device__ __forceinline__ double nhalf_inverse<double>(double v) { return -1.0 / (2.0 * v); }

// Values of the TAPS horizontal taps of one staged row for the V columns of a vector.
template <typename T, int TAPS, int DMODE>
__device__ __forceinline__ void row_taps(const T *srow, const TapPlan<PlanSize<TAPS, DMODE>::NV> &tp,
                                         T (&out)[TAPS][VecOf<T>::V]) {
    constexpr int V = VecOf<T>::V;
    constexpr int C = TAPS / 2;
    if constexpr (DMODE == 0) {
#pragma unroll
        for (int k = 0; k < TAPS; ++k) {
            Pack<T, V> t = ld_vec(srow + tp.off[k]);
            if ((tp.rev >> k) & 1u) reverse_vec<T, V>(t);
#pragma unroll
            for (int e = 0; e < V; ++e) out[k][e] = t.v[e];
        }
    } else {
        T win[3 * V];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            Pack<T, V> t = ld_vec(srow + tp.off[k]);
            if ((tp.rev >> k) & 1u) reverse_vec<T, V>(t);
#pragma unroll
            for (int e = 0; e < V; ++e) win[k * V + e] = t.v[e];
        }
#pragma unroll
        for (int k = 0; k < TAPS; ++k)
#pragma unroll
            for (int e = 0; e < V; ++e) out[k][e] = win[V + e + (k - C) * DMODE];
    }
}

template <typename T, int TAPS, int DMODE>
__global__ void __launch_bounds__(288) bilateral_rows_kernel(const BilateralParams bp) {
    const ScaleParams &p = bp.sp;
    constexpr int V = VecOf<T>::V;
    constexpr int C = TAPS / 2;
    constexpr int NV = PlanSize<TAPS, DMODE>::NV;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *rows = reinterpret_cast<T *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)p.slots * p.row_stride * sizeof(T));
    uint64_t *empty = full + p.slots;

    const int nt = blockDim.x - 32;
    const int nwc = nt >> 5;
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    int bx = blockIdx.x;
    const int strip = bx % p.n_strips;
    bx /= p.n_strips;
    const int r = bx % p.d;
    const int g = bx / p.d;
    const int frame = blockIdx.y;

    const int n_chain = (r < p.H) ? (p.H - r + p.d - 1) / p.d : 0;
    const int i0 = g * p.seg;
    const int n_out = min(p.seg, n_chain - i0);
    if (n_out <= 0) return;
    const int n_load = n_out + 2 * C;

    const int x0 = strip * p.wt;
    const int lo = max(0, x0 - p.halo_al);
    const int hi = min(p.W, x0 + p.wt + p.halo_al);
    const uint32_t row_bytes = (uint32_t)(hi - lo) * (uint32_t)sizeof(T);

    if (tid == 0) {
        for (int s = 0; s < p.slots; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], nwc);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == nwc) {
        if (lane == 0) {
            const T *src = reinterpret_cast<const T *>(p.in) + (long long)frame * p.in_bstride + lo;
            int slot = 0;
            uint32_t round = 0;
            for (int j = 0; j < n_load; ++j) {
                if (round > 0) mbar_wait(&empty[slot], (round - 1) & 1);
                const long long y = reflect_any(p.gwy0 + r + (long long)(i0 - C + j) * p.d, p.Hg) - p.gwy0 + p.row_off_in;
                mbar_arrive_expect_tx(&full[slot], row_bytes);
                tma_load_1d(rows + (size_t)slot * p.row_stride, src + y * p.in_pitch, row_bytes,
                            &full[slot]);
                if (++slot == p.slots) { slot = 0; ++round; }
            }
        }
        return;
    }

    T *out_c = reinterpret_cast<T *>(p.out_c);
    T *out_w = reinterpret_cast<T *>(p.out_w);
    if (out_c) out_c += (long long)frame * p.c_bstride;
    if (out_w) out_w += (long long)frame * p.w_bstride;

    const int xg = x0 + tid * V;
    const bool act = xg < p.W;
    const TapPlan<NV> plan = make_tap_plan<V, NV>(act ? xg : x0, DMODE == 0 ? p.d : V, p.W, lo);
    const T var_factor = (T)bp.var_factor;

    long long orow = (long long)r + (long long)i0 * p.d;
    int slot = 0, fslot = 0;  // slot of row j, slot of row j - 2C (first row of the window, next to be released)
    uint32_t parity = 0;
    for (int j = 0; j < n_load; ++j) {
        mbar_wait(&full[slot], parity);
        if (j >= 2 * C) {
            if (act) {
                // centre pixel: row j - C of the window, columns xg .. xg+V-1
                int cs = fslot + C;
                if (cs >= p.slots) cs -= p.slots;
                Pack<T, V> xc = ld_vec(rows + (size_t)cs * p.row_stride + (xg - lo));
                T dlt[TAPS][TAPS][V];
                T s1[V], s2[V];
#pragma unroll
                for (int e = 0; e < V; ++e) { s1[e] = T(0); s2[e] = T(0); }
                int ws = fslot;
#pragma unroll
                for (int i = 0; i < TAPS; ++i) {
                    T tv[TAPS][V];
                    row_taps<T, TAPS, DMODE>(rows + (size_t)ws * p.row_stride, plan, tv);
#pragma unroll
                    for (int k = 0; k < TAPS; ++k) {
                        const T kk = Taps<T, TAPS>::h(i) * Taps<T, TAPS>::h(k);
#pragma unroll
                        for (int e = 0; e < V; ++e) {
                            const T dd = xc.v[e] - tv[k][e];
                            dlt[i][k][e] = dd;
                            const T kd = kk * dd;
                            s1[e] += kd;
                            s2[e] = fma_t<T>(kd, dd, s2[e]);
                        }
                    }
                    if (++ws == p.slots) ws = 0;
                }
                Pack<T, V> cn, wv;
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    T var = s2[e] - s1[e] * s1[e];
                    if (var <= T(0)) var = T(1e-20);
                    const T nhi = nhalf_inverse<T>(var * var_factor);
                    T num = T(0);
                    T den = Taps<T, TAPS>::h(C) * Taps<T, TAPS>::h(C);
#pragma unroll
                    for (int i = 0; i < TAPS; ++i)
#pragma unroll
                        for (int k = 0; k < TAPS; ++k) {
                            if (i == C && k == C) continue;
                            const T dd = dlt[i][k][e];
                            const T gw = range_weight(Taps<T, TAPS>::h(i) * Taps<T, TAPS>::h(k), dd * dd, nhi);
                            den += gw;
                            num = fma_t<T>(gw, dd, num);
                        }
                    cn.v[e] = xc.v[e] - num / den;
                    wv.v[e] = xc.v[e] - cn.v[e];
                }
                if (out_c) st_vec(out_c + orow * p.c_pitch + xg, cn);
                if (out_w) st_vec_cs(out_w + orow * p.w_pitch + xg, wv);
            }
            orow += p.d;
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[fslot]);
            if (++fslot == p.slots) fslot = 0;
        }
        if (++slot == p.slots) { slot = 0; parity ^= 1; }
    }
}

template <typename T, int TAPS>
__global__ void __launch_bounds__(256) bilateral_generic_kernel(const BilateralParams bp) {
    const ScaleParams &p = bp.sp;
    constexpr int C = TAPS / 2;
    const long long n = (long long)p.H * p.W;
    const int frame = blockIdx.y;
    const T *in = reinterpret_cast<const T *>(p.in) + (long long)frame * p.in_bstride;
    T *out_c = reinterpret_cast<T *>(p.out_c);
    T *out_w = reinterpret_cast<T *>(p.out_w);
    const T var_factor = (T)bp.var_factor;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(idx / p.W), x = (int)(idx % p.W);
        const T xc = in[(long long)y * p.in_pitch + x];
        T dlt[TAPS][TAPS];
        T s1 = T(0), s2 = T(0);
#pragma unroll
        for (int i = 0; i < TAPS; ++i) {
            const T *row = in + (long long)reflect_any((long long)y + (long long)(i - C) * p.d, p.Hg) * p.in_pitch;
#pragma unroll
            for (int k = 0; k < TAPS; ++k) {
                const T dd = xc - row[reflect_any((long long)x + (long long)(k - C) * p.d, p.W)];
                dlt[i][k] = dd;
                const T kd = Taps<T, TAPS>::h(i) * Taps<T, TAPS>::h(k) * dd;
                s1 += kd;
                s2 = fma_t<T>(kd, dd, s2);
            }
        }
        T var = s2 - s1 * s1;
        if (var <= T(0)) var = T(1e-20);
        const T nhi = nhalf_inverse<T>(var * var_factor);
        T num = T(0), den = Taps<T, TAPS>::h(C) * Taps<T, TAPS>::h(C);
#pragma unroll
        for (int i = 0; i < TAPS; ++i)
#pragma unroll
            for (int k = 0; k < TAPS; ++k) {
                if (i == C && k == C) continue;
                const T gw = range_weight(Taps<T, TAPS>::h(i) * Taps<T, TAPS>::h(k), dlt[i][k] * dlt[i][k], nhi);
                den += gw;
                num = fma_t<T>(gw, dlt[i][k], num);
            }
        const T cn = xc - num / den;
        if (out_c) out_c[(long long)frame * p.c_bstride + (long long)y * p.c_pitch + x] = cn;
        if (out_w) out_w[(long long)frame * p.w_bstride + (long long)y * p.w_pitch + x] = xc - cn;
    }
}

template <typename T, int TAPS, int DMODE>
static int launch_bilateral(const BilateralParams &bp, int batch, int nt, cudaStream_t st) {
    auto kern = bilateral_rows_kernel<T, TAPS, DMODE>;
    const ScaleParams &p = bp.sp;
    const size_t smem = (size_t)p.slots * p.row_stride * sizeof(T) + 16 * (size_t)p.slots;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    dim3 grid((unsigned)((long long)p.n_strips * p.d * p.n_seg), (unsigned)batch);
    kern<<<grid, nt + 32, smem, st>>>(bp);
    return launch_status();
}

// Geometry for K2: one vector per thread, 256 consumer threads, the ring holds the `taps` window rows plus prefetch.
static bool plan_bilateral(ScaleParams &p, int taps, int esize, int batch) {
    const int V = 16 / esize;
    const int c = taps / 2;
    const int nt = 256;
    p.wt = nt * V;
    p.n_strips = (p.W + p.wt - 1) / p.wt;
    p.halo_al = round_up(c * p.d, V);
    long long rs = (long long)p.wt + 2LL * p.halo_al;
    if (rs > p.W) rs = p.W;
    p.row_stride = (int)rs;
    int slots = taps + 3;
    const int min_slots = taps + 1;
    while (slots > min_slots && (long long)slots * p.row_stride * esize + 16LL * slots > kMaxSmem) --slots;
    if ((long long)slots * p.row_stride * esize + 16LL * slots > kMaxSmem) return false;
    p.slots = slots;
    const int n_max = (p.H + p.d - 1) / p.d;
    // compute-bound kernel: many short segments balance better than few long ones; halo rows only cost L2 reads
    const long long chains = (long long)p.n_strips * (p.d < p.H ? p.d : p.H) * batch;
    const long long target = 8LL * device_sm_count();
    long long per_chain = (target + chains - 1) / chains;
    if (per_chain < 1) per_chain = 1;
    int seg = (int)((n_max + per_chain - 1) / per_chain);
    if (seg < 8 * c) seg = 8 * c;
    if (seg > n_max) seg = n_max;
    p.seg = seg;
    p.n_seg = (n_max + seg - 1) / seg;
    return true;
}

template <typename T, int TAPS>
static int dispatch_bilateral(BilateralParams &bp, int batch, cudaStream_t st) {
    constexpr int V = VecOf<T>::V;
    ScaleParams &p = bp.sp;
    if (fast_path_ok(p, TAPS, (int)sizeof(T)) && plan_bilateral(p, TAPS, (int)sizeof(T), batch)) {
        const int dmode = (p.d % V == 0) ? 0 : p.d;
        if (dmode == 0) return launch_bilateral<T, TAPS, 0>(bp, batch, 256, st);
        if (dmode == 1) return launch_bilateral<T, TAPS, 1>(bp, batch, 256, st);
        if constexpr (V == 4) {
            if (dmode == 2) return launch_bilateral<T, TAPS, 2>(bp, batch, 256, st);
        }
    }
    const long long n = (long long)p.H * p.W;
    long long blocks = (n + 255) / 256;
    const long long cap = 32LL * device_sm_count();
    if (blocks > cap) blocks = cap;
    bilateral_generic_kernel<T, TAPS><<<dim3((unsigned)blocks, (unsigned)batch), 256, 0, st>>>(bp);
    return launch_status();
}

}  // namespace wb

extern "C" {

int wb_atrous_scale_bilateral(const void *in, void *out_c, void *out_w, int batch, int H, int W, long long in_pitch,
                              long long in_bstride, long long out_c_pitch, long long out_c_bstride,
                              long long out_w_pitch, long long out_w_bstride, int scale, int taps, int dtype,
                              double var_factor, void *stream) {
    int rc = wb::check_common(batch, H, W, taps, dtype);
    if (rc) return rc;
    if (scale < 0 || scale > 30) return WB_EINVAL_SCALE;
    if (!in || (!out_c && !out_w)) return WB_EINVAL_POINTER;
    if (in == out_c || in == out_w) return WB_EINVAL_POINTER;
    if (in_pitch < W || (out_c && out_c_pitch < W) || (out_w && out_w_pitch < W) || !(var_factor > 0)) return WB_EINVAL_ARG;
    wb::BilateralParams bp;
    memset(&bp, 0, sizeof(bp));
    wb::ScaleParams &p = bp.sp;
    p.in = in; p.out_c = out_c; p.out_w = out_w;
    p.H = H; p.W = W; p.d = 1 << scale; p.Hg = H;
    p.in_pitch = in_pitch; p.in_bstride = in_bstride;
    p.c_pitch = out_c_pitch; p.c_bstride = out_c_bstride;
    p.w_pitch = out_w_pitch; p.w_bstride = out_w_bstride;
    bp.var_factor = var_factor;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == WB_F32)
        return taps == 3 ? wb::dispatch_bilateral<float, 3>(bp, batch, st) : wb::dispatch_bilateral<float, 5>(bp, batch, st);
    return taps == 3 ? wb::dispatch_bilateral<double, 3>(bp, batch, st) : wb::dispatch_bilateral<double, 5>(bp, batch, st);
}

}  // extern "C"
