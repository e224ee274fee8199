// K4 -- global reductions of the hot path, entirely on the device (no host synchronisation):
//   * exact median of |x| over a plane  (Coefficients.get_noise, watroo/wavelets.py:126-127: np.median(np.abs(w_0)))
//   * population mean / variance / std of a plane  (np.std of the residual plane, watroo/utils.py:187; per-plane std
//     of compute_noise_weights, watroo/wavelets.py:227)
//
// Median: order statistics of |x| are order statistics of the IEEE bit patterns of |x| (unsigned integers).  A
// sampled bracket [lo, hi] around the median is refined by full passes that only COUNT: keys below the bracket go
// to a per-thread counter, keys inside it to a 2048-bin shared-memory histogram with per-bin min/max (few percent
// of the keys, spread over many bins, so atomics do not contend).  A one-block "decide" kernel narrows the bracket
// to the bin(s) holding the two middle ranks; once few keys remain they are collected and sorted in shared memory.
// Ties, bimodal data and a missed initial bracket are handled exactly (the bracket then restarts from the half
// line that holds the ranks); every step reads its state from device memory, so a fixed launch sequence is issued
// and converged steps exit at once.  Typical cost: 2 streaming reads of the plane.
#include "common.cuh"

namespace wb {

constexpr int kBins = 2048;
constexpr int kCollectCap = 4096;
constexpr int kSample = 2048;

template <typename K> struct SelState {
    K lo, hi;                      // inclusive key bracket
    K res_lo, res_hi;              // keys of ranks k_lo, k_hi once done
    unsigned long long below;      // number of keys < lo (known exactly when below_valid)
    unsigned long long k_lo, k_hi; // target ranks (0-based)
    unsigned long long acc_below;  // accumulated by the pass kernel
    unsigned int n_collected;
    int shift;                     // bin = (key - lo) >> shift
    int mode;                      // 0 = count pass, 1 = collect pass
    int done;
    int fresh;                     // 1 = bracket comes from the sample (below unknown)
};

template <typename T> struct KeyOf;
template <> struct KeyOf<float> {
    using type = uint32_t;
    static __device__ __forceinline__ uint32_t key(float x) { return __float_as_uint(fabsf(x)); }
    static __device__ __forceinline__ float val(uint32_t k) { return __uint_as_float(k); }
    static constexpr uint32_t kmax = 0xFFFFFFFFu;
};
template <> struct KeyOf<double> {
    using type = unsigned long long;
    static __device__ __forceinline__ unsigned long long key(double x) {
        return (unsigned long long)__double_as_longlong(fabs(x));
    }
    static __device__ __forceinline__ double val(unsigned long long k) { return __longlong_as_double((long long)k); }
    static constexpr unsigned long long kmax = 0xFFFFFFFFFFFFFFFFull;
};

template <typename K> struct Workspace {
    SelState<K> st;
    unsigned int hist[kBins];
    K bmin[kBins];
    K bmax[kBins];
    K collect[kCollectCap];
};

template <typename K> __device__ __forceinline__ int shift_for(K lo, K hi) {
    // smallest shift with ((hi - lo) >> shift) < kBins
    K width = hi - lo;
    int s = 0;
    while ((width >> s) >= (K)kBins) ++s;
    return s;
}

// In-place bitonic sort of n (power of two) keys in shared memory by the whole block.
template <typename K> __device__ void bitonic_sort(K *a, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    K x = a[i], y = a[ixj];
                    bool up = ((i & k) == 0);
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(1024) select_init_kernel(const T *x, long long n, long long bstride,
                                                           Workspace<typename KeyOf<T>::type> *ws_all) {
    using K = typename KeyOf<T>::type;
    __shared__ K s[kSample];
    Workspace<K> *ws = ws_all + blockIdx.x;
    const T *xf = x + (long long)blockIdx.x * bstride;
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) {
        ws->hist[i] = 0;
        ws->bmin[i] = KeyOf<T>::kmax;
        ws->bmax[i] = 0;
    }
    const int m = n < kSample ? 0 : kSample;  // tiny inputs: skip sampling, bracket = everything
    if (m) {
        const long long stride = n / m;
        for (int i = threadIdx.x; i < m; i += blockDim.x) s[i] = KeyOf<T>::key(xf[(long long)i * stride + stride / 2]);
        __syncthreads();
        bitonic_sort<K>(s, m);
    }
    if (threadIdx.x == 0) {
        SelState<K> &st = ws->st;
        st.k_lo = (unsigned long long)((n - 1) / 2);
        st.k_hi = (unsigned long long)(n / 2);
        st.acc_below = 0;
        st.n_collected = 0;
        st.done = 0;
        st.res_lo = st.res_hi = 0;
        if (m) {
            const int margin = 113;  // 5 sigma of the sample-median rank for m = 2048 (sqrt(m) / 2 per sigma)
            st.lo = s[m / 2 - margin];
            st.hi = s[m / 2 + margin];
            st.fresh = 1;
            st.below = 0;
            st.mode = 0;
        } else {
            st.lo = 0;
            st.hi = KeyOf<T>::kmax;
            st.fresh = 0;
            st.below = 0;
            st.mode = (n <= kCollectCap) ? 1 : 0;
        }
        st.shift = shift_for<K>(st.lo, st.hi);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) select_pass_kernel(const T *x, long long n, long long bstride,
                                                          Workspace<typename KeyOf<T>::type> *ws_all) {
    using K = typename KeyOf<T>::type;
    Workspace<K> *ws = ws_all + blockIdx.y;
    if (ws->st.done) return;
    const T *xf = x + (long long)blockIdx.y * bstride;
    const K lo = ws->st.lo, hi = ws->st.hi;
    const int shift = ws->st.shift, mode = ws->st.mode;
    __shared__ unsigned int h[kBins];
    __shared__ K hmin[kBins];
    __shared__ K hmax[kBins];
    __shared__ unsigned long long blk_below;
    if (mode == 0) {
        for (int i = threadIdx.x; i < kBins; i += blockDim.x) { h[i] = 0; hmin[i] = KeyOf<T>::kmax; hmax[i] = 0; }
    }
    if (threadIdx.x == 0) blk_below = 0;
    __syncthreads();
    unsigned long long below = 0;
    constexpr int V = VecOf<T>::V;
    const long long nvec = n / V;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(xf) & 15u) == 0);
    auto visit = [&](T v) {
        const K key = KeyOf<T>::key(v);
        if (key < lo) {
            ++below;
        } else if (key <= hi) {
            if (mode == 0) {
                const int b = (int)((key - lo) >> shift);
                atomicAdd(&h[b], 1u);
                atomicMin(&hmin[b], key);
                atomicMax(&hmax[b], key);
            } else {
                const unsigned int pos = atomicAdd(&ws->st.n_collected, 1u);
                if (pos < (unsigned)kCollectCap) ws->collect[pos] = key;
            }
        }
    };
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    if (vec_ok) {
        // four independent 16-byte loads in flight per thread: one load per iteration left the pass latency-bound
        // (64 MiB in 34 us); the visits only touch registers and (rarely) shared-memory atomics
        long long i = tid;
        for (; i + 3 * nth < nvec; i += 4 * nth) {
            Pack<T, V> p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) p[u] = ld_vec(xf + (i + u * nth) * V);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int e = 0; e < V; ++e) visit(p[u].v[e]);
        }
        for (; i < nvec; i += nth) {
            Pack<T, V> p = ld_vec(xf + i * V);
#pragma unroll
            for (int e = 0; e < V; ++e) visit(p.v[e]);
        }
        for (long long i2 = nvec * V + tid; i2 < n; i2 += nth) visit(xf[i2]);
    } else {
        for (long long i = tid; i < n; i += nth) visit(xf[i]);
    }
    // block reduction of the below counter
    for (int o = 16; o > 0; o >>= 1) below += __shfl_down_sync(0xffffffffu, below, o);
    if ((threadIdx.x & 31) == 0 && below) atomicAdd(&blk_below, below);
    __syncthreads();
    if (threadIdx.x == 0 && blk_below) atomicAdd(&ws->st.acc_below, blk_below);
    if (mode == 0) {
        for (int i = threadIdx.x; i < kBins; i += blockDim.x) {
            if (h[i]) {
                atomicAdd(&ws->hist[i], h[i]);
                atomicMin(&ws->bmin[i], hmin[i]);
                atomicMax(&ws->bmax[i], hmax[i]);
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(1024) select_decide_kernel(long long n, Workspace<typename KeyOf<T>::type> *ws_all,
                                                             T *out_median, double *out_noise, double sigma_e0,
                                                             int last) {
    using K = typename KeyOf<T>::type;
    Workspace<K> *ws = ws_all + blockIdx.x;
    SelState<K> &st = ws->st;
    // one buffer, two uses: candidate keys (collect mode) or the cumulative histogram (count mode)
    __shared__ unsigned long long buf[kCollectCap];
    K *s = reinterpret_cast<K *>(buf);
    unsigned long long *cum = buf;
    __shared__ int fin;
    if (st.done) return;
    if (threadIdx.x == 0) fin = 0;
    __syncthreads();

    if (st.mode == 1) {
        // ---- collect pass finished: sort the candidates and read the two ranks off --------------------------
        const unsigned int cnt = st.n_collected;
        const unsigned long long below = st.fresh ? st.acc_below : st.below;
        // cnt <= kCollectCap is guaranteed by construction (mode 1 is only entered with a known small count)
        int npow = 1;
        while (npow < (int)cnt) npow <<= 1;
        for (int i = threadIdx.x; i < npow; i += blockDim.x) s[i] = (i < (int)cnt) ? ws->collect[i] : KeyOf<T>::kmax;
        __syncthreads();
        bitonic_sort<K>(s, npow);
        if (threadIdx.x == 0) {
            st.res_lo = s[st.k_lo - below];
            st.res_hi = s[st.k_hi - below];
            st.done = 1;
            fin = 1;
        }
    } else {
        // ---- count pass finished: locate the bins of the two ranks ---------------------------------------------
        const unsigned long long below = st.fresh ? st.acc_below : st.below;
        {
            // inclusive prefix of the histogram by the whole block (a serial loop over 2048 bins in global memory took
            // 60 us): each thread scans its kBins / blockDim consecutive bins, warp shuffles scan the per-thread
            // totals, one more pass over the warp totals finishes it
            constexpr int PER = kBins / 1024;
            static_assert(kBins % 1024 == 0, "decide kernel: 1024 threads x PER bins");
            __shared__ unsigned long long warp_tot[32];
            unsigned long long loc[PER];
            unsigned long long run = 0;
#pragma unroll
            for (int e = 0; e < PER; ++e) { run += ws->hist[threadIdx.x * PER + e]; loc[e] = run; }
            unsigned long long inc = run;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
                if ((threadIdx.x & 31) >= o) inc += v;
            }
            if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = inc;
            __syncthreads();
            if (threadIdx.x < 32) {
                unsigned long long w = warp_tot[threadIdx.x];
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned long long v = __shfl_up_sync(0xffffffffu, w, o);
                    if (threadIdx.x >= o) w += v;
                }
                warp_tot[threadIdx.x] = w;
            }
            __syncthreads();
            const unsigned long long base = below + (inc - run) + ((threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0ull);
#pragma unroll
            for (int e = 0; e < PER; ++e) cum[threadIdx.x * PER + e] = base + loc[e];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long inside_end = cum[kBins - 1];
            if (st.k_lo < below) {
                // sampled bracket missed: both ranks are below it (k_hi may equal the first key of the bracket;
                // keep lo itself inside the new bracket so that case stays covered)
                st.hi = st.lo;
                st.lo = 0;
                st.below = 0;
                st.mode = 0;
            } else if (st.k_hi >= inside_end) {
                // missed on the other side; keep hi inside for the straddling case
                unsigned long long b = below;
                // keys < hi: everything below the bracket plus all bins except the keys equal to hi -- unknown,
                // so restart the count from the last bin's lower edge, whose cumulative count is known
                int lastbin = (int)((st.hi - st.lo) >> st.shift);
                b = (lastbin > 0) ? cum[lastbin - 1] : below;
                st.lo = st.lo + ((K)lastbin << st.shift);
                st.hi = KeyOf<T>::kmax;
                st.below = b;
                st.mode = 0;
            } else {
                int b_lo = 0, b_hi = 0;
                // first bin whose inclusive cumulative count exceeds the rank
                int l = 0, r = kBins - 1;
                while (l < r) { int m = (l + r) >> 1; if (cum[m] > st.k_lo) r = m; else l = m + 1; }
                b_lo = l;
                l = 0; r = kBins - 1;
                while (l < r) { int m = (l + r) >> 1; if (cum[m] > st.k_hi) r = m; else l = m + 1; }
                b_hi = l;
                if (b_lo != b_hi) {
                    // rank k_lo is the largest key of its bin, rank k_hi the smallest key of a later bin
                    st.res_lo = ws->bmax[b_lo];
                    st.res_hi = ws->bmin[b_hi];
                    st.done = 1;
                    fin = 1;
                } else {
                    const K nlo = ws->bmin[b_lo], nhi = ws->bmax[b_lo];
                    const unsigned long long nbelow = (b_lo > 0) ? cum[b_lo - 1] : below;
                    if (nlo == nhi) {
                        st.res_lo = st.res_hi = nlo;
                        st.done = 1;
                        fin = 1;
                    } else {
                        st.lo = nlo;
                        st.hi = nhi;
                        st.below = nbelow;
                        st.mode = (ws->hist[b_lo] <= (unsigned)kCollectCap) ? 1 : 0;
                    }
                }
            }
            if (!st.done) {
                st.fresh = 0;
                st.shift = shift_for<K>(st.lo, st.hi);
                st.acc_below = 0;
                st.n_collected = 0;
            }
        }
        __syncthreads();
        if (!st.done)
            for (int i = threadIdx.x; i < kBins; i += blockDim.x) {
                ws->hist[i] = 0;
                ws->bmin[i] = KeyOf<T>::kmax;
                ws->bmax[i] = 0;
            }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (fin) {
            // np.median: mean of the two middle values, in the plane dtype
            const T a = KeyOf<T>::val(st.res_lo), b = KeyOf<T>::val(st.res_hi);
            const T med = (a + b) / T(2);
            if (out_median) out_median[blockIdx.x] = med;
            if (out_noise) {
                // get_noise (watroo/wavelets.py:127) with NumPy>=2 promotion: '/0.6745' in the plane dtype,
                // '/sigma_e[0]' in float64
                const T q = med / T(0.6745);
                out_noise[blockIdx.x] = (double)q / sigma_e0;
            }
        } else if (last) {
            // not converged within the launch budget (cannot happen for finite data): flag with NaN
            if (out_median) out_median[blockIdx.x] = T(NAN);
            if (out_noise) out_noise[blockIdx.x] = NAN;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Moments: shifted one-pass sums in float64, deterministic two-stage reduction
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMomBlocks = 592;  // 4 per SM on a 148-SM part
constexpr int kMomSamples = 256;

template <typename T>
__global__ void __launch_bounds__(256) moments_partial_kernel(const T *x, long long n, long long bstride,
                                                              double *partial /* [batch][kMomBlocks][2] */,
                                                              double *shift_out /* [batch] */) {
    const T *xf = x + (long long)blockIdx.y * bstride;
    // shift K: mean of up to 256 strided samples -- identical in every block, keeps the sums well conditioned
    __shared__ double sh[8];
    __shared__ double kshift;
    {
        const long long m = n < kMomSamples ? n : kMomSamples;
        const long long stride = n / m;
        double v = (threadIdx.x < m) ? (double)xf[(long long)threadIdx.x * stride] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0;
            for (int i = 0; i < 8; ++i) t += sh[i];
            kshift = t / (double)m;
            if (blockIdx.x == 0) shift_out[blockIdx.y] = kshift;
        }
        __syncthreads();
    }
    const double K = kshift;
    double s1 = 0, s2 = 0;
    constexpr int V = VecOf<T>::V;
    const long long nvec = n / V;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    if ((reinterpret_cast<uintptr_t>(xf) & 15u) == 0) {
        // (same summation order as a one-load-per-iteration loop: the four vectors are consumed in index order)
        long long i = tid;
        for (; i + 3 * nth < nvec; i += 4 * nth) {
            Pack<T, V> p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) p[u] = ld_vec(xf + (i + u * nth) * V);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int e = 0; e < V; ++e) { double dlt = (double)p[u].v[e] - K; s1 += dlt; s2 = fma(dlt, dlt, s2); }
        }
        for (; i < nvec; i += nth) {
            Pack<T, V> p = ld_vec(xf + i * V);
#pragma unroll
            for (int e = 0; e < V; ++e) { double dlt = (double)p.v[e] - K; s1 += dlt; s2 = fma(dlt, dlt, s2); }
        }
        for (long long i = nvec * V + tid; i < n; i += nth) { double dlt = (double)xf[i] - K; s1 += dlt; s2 = fma(dlt, dlt, s2); }
    } else {
        for (long long i = tid; i < n; i += nth) { double dlt = (double)xf[i] - K; s1 += dlt; s2 = fma(dlt, dlt, s2); }
    }
    __shared__ double r1[8], r2[8];
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int i = 0; i < 8; ++i) { a += r1[i]; b += r2[i]; }
        double *dst = partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 2;
        dst[0] = a;
        dst[1] = b;
    }
}

__global__ void __launch_bounds__(256) moments_final_kernel(const double *partial, const double *shift, long long n,
                                                            int nblocks, double *out /* [batch][3] mean,var,std */) {
    const double *src = partial + (long long)blockIdx.x * nblocks * 2;
    double s1 = 0, s2 = 0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) { s1 += src[2 * i]; s2 += src[2 * i + 1]; }
    __shared__ double r1[8], r2[8];
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int i = 0; i < 8; ++i) { a += r1[i]; b += r2[i]; }
        const double m1 = a / (double)n;
        double var = b / (double)n - m1 * m1;
        if (var < 0) var = 0;
        out[blockIdx.x * 3 + 0] = shift[blockIdx.x] + m1;
        out[blockIdx.x * 3 + 1] = var;
        out[blockIdx.x * 3 + 2] = sqrt(var);
    }
}

template <typename T>
static int median_impl(const void *x, long long n, int batch, long long bstride, void *out_median, double *out_noise,
                       double sigma_e0, void *workspace, cudaStream_t st) {
    using K = typename KeyOf<T>::type;
    auto *ws = reinterpret_cast<Workspace<K> *>(workspace);
    const T *xp = reinterpret_cast<const T *>(x);
    select_init_kernel<T><<<batch, 1024, 0, st>>>(xp, n, bstride, ws);
    long long blocks = (n / VecOf<T>::V + 255) / 256;
    int sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
    }
    if (blocks > 8LL * sms) blocks = 8LL * sms;
    if (blocks < 1) blocks = 1;
    const int passes = sizeof(T) == 4 ? 5 : 9;
    for (int p = 0; p < passes; ++p) {
        select_pass_kernel<T><<<dim3((unsigned)blocks, (unsigned)batch), 256, 0, st>>>(xp, n, bstride, ws);
        select_decide_kernel<T><<<batch, 1024, 0, st>>>(n, ws, reinterpret_cast<T *>(out_median), out_noise, sigma_e0,
                                                        p == passes - 1);
    }
    return launch_status();
}

}  // namespace wb

extern "C" {

size_t wb_abs_median_workspace_bytes(int dtype, int batch) {
    if (batch < 1) batch = 1;
    size_t per = dtype == WB_F64 ? sizeof(wb::Workspace<unsigned long long>) : sizeof(wb::Workspace<uint32_t>);
    return per * (size_t)batch;
}

int wb_abs_median(const void *x, long long n, int batch, long long bstride, int dtype, void *out_median,
                  double *out_noise, double sigma_e0, void *workspace, void *stream) {
    if (dtype != WB_F32 && dtype != WB_F64) return WB_EINVAL_DTYPE;
    if (n < 1 || batch < 1 || batch > 65535) return WB_EINVAL_SHAPE;
    if (!x || !workspace || (!out_median && !out_noise)) return WB_EINVAL_POINTER;
    if (out_noise && !(sigma_e0 > 0)) return WB_EINVAL_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == WB_F32
               ? wb::median_impl<float>(x, n, batch, bstride, out_median, out_noise, sigma_e0, workspace, st)
               : wb::median_impl<double>(x, n, batch, bstride, out_median, out_noise, sigma_e0, workspace, st);
}

size_t wb_plane_moments_workspace_bytes(int batch) {
    if (batch < 1) batch = 1;
    return (size_t)batch * (wb::kMomBlocks * 2 + 1) * sizeof(double);
}

int wb_plane_moments(const void *x, long long n, int batch, long long bstride, int dtype, double *out,
                     void *workspace, void *stream) {
    if (dtype != WB_F32 && dtype != WB_F64) return WB_EINVAL_DTYPE;
    if (n < 1 || batch < 1 || batch > 65535) return WB_EINVAL_SHAPE;
    if (!x || !out || !workspace) return WB_EINVAL_POINTER;
    cudaStream_t st = (cudaStream_t)stream;
    double *partial = reinterpret_cast<double *>(workspace);
    double *shift = partial + (size_t)batch * wb::kMomBlocks * 2;
    dim3 grid(wb::kMomBlocks, (unsigned)batch);
    if (dtype == WB_F32)
        wb::moments_partial_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float *>(x), n, bstride, partial, shift);
    else
        wb::moments_partial_kernel<double><<<grid, 256, 0, st>>>(reinterpret_cast<const double *>(x), n, bstride, partial, shift);
    wb::moments_final_kernel<<<batch, 256, 0, st>>>(partial, shift, n, wb::kMomBlocks, out);
    return wb::launch_status();
}

}  // extern "C"
