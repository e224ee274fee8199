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
// and converged steps exit at once.
// Round 2: the FIRST pass is a filter -- it counts the keys below the sampled bracket and copies the values inside it
// (about 11 % of the plane for the 5-sigma bracket of a 2048-key sample) into a compact buffer of the workspace, staged
// through shared memory; every later pass reads that buffer instead of the plane.  Typical cost: ONE streaming read of
// the plane plus two passes over ~1/9 of it (before: two full reads with three shared-memory atomics per key inside
// the bracket, 164 us for a 4096^2 fp32 plane).  A bracket that misses the ranks or overflows the buffer (heavy
// ties) falls back to full-plane counting passes, still exact.
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
    // filter pass (mode 2): values inside the bracket are copied to the compact buffer of the workspace
    unsigned long long acc_inside; // number of keys inside the bracket (counted even when the buffer overflows)
    unsigned long long cur_n;      // use_compact: number of values in the compact buffer
    unsigned long long gcap;       // capacity of the compact buffer (values)
    int use_compact;               // 1 = the passes read the compact buffer instead of the plane
    // edge pass (mode 3): the two ranks sit in different bins -> largest key <= edge_a and smallest key >= edge_b
    K edge_a, edge_b, found_max, found_min;
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
    K collect[kCollectCap];
};

template <typename K> __device__ __forceinline__ int shift_for(K lo, K hi) {
    // smallest shift with ((hi - lo) >> shift) < kBins
    K width = hi - lo;
    int s = 0;
    while ((width >> s) >= (K)kBins) ++s;
    return s;
}

// Key of rank `rank` (0-based) among the n keys of a[] (shared memory), by the whole block: most-significant-digit
// radix select, 8 bits per round, a 256-bin shared-memory histogram of the keys that still match the prefix.  4 rounds
// of ~4 block barriers for 32-bit keys -- the bitonic sort it replaces needed 66 - 78 barrier-separated stages and
// cost 20 us per call.  Every thread returns the key.  `bins` / `sel`: 256 + 4 words of shared scratch.
template <typename K>
__device__ K smem_select(const K *a, int n, unsigned long long rank, unsigned int *bins, unsigned long long *sel) {
    constexpr int kBits = (int)sizeof(K) * 8;
    K prefix = 0, mask = 0;
    for (int sh = kBits - 8; sh >= 0; sh -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) bins[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const K k = a[i];
            if ((k & mask) == prefix) atomicAdd(&bins[(unsigned)((k >> sh) & (K)0xFF)], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // warp 0: inclusive prefix over the 256 bins (8 per lane), find the digit whose range holds the rank
            unsigned int loc[8], run = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) { run += bins[threadIdx.x * 8 + e]; loc[e] = run; }
            unsigned int inc = run;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
                if ((int)threadIdx.x >= o) inc += v;
            }
            const unsigned int base = inc - run;
            if (rank >= base && rank < inc) {  // exactly one lane
                int e = 0;
                while (rank >= base + loc[e]) ++e;
                sel[0] = (unsigned long long)(threadIdx.x * 8 + e);
                sel[1] = rank - (base + (e ? loc[e - 1] : 0u));
            }
        }
        __syncthreads();
        prefix |= (K)sel[0] << sh;
        mask |= (K)0xFF << sh;
        rank = sel[1];
        __syncthreads();
    }
    return prefix;
}

template <typename T>
__global__ void __launch_bounds__(1024) select_init_kernel(const T *x, long long n, long long bstride,
                                                           Workspace<typename KeyOf<T>::type> *ws_all,
                                                           unsigned long long gcap) {
    pdl_launch_dependents();
    pdl_wait();  // launched with programmatic stream serialisation: nothing is read before the predecessor is complete
    using K = typename KeyOf<T>::type;
    __shared__ K s[kSample];
    Workspace<K> *ws = ws_all + blockIdx.x;
    const T *xf = x + (long long)blockIdx.x * bstride;
    __shared__ unsigned int bins[256];
    __shared__ unsigned long long sel[2];
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) ws->hist[i] = 0;
    const int m = n < kSample ? 0 : kSample;  // tiny inputs: skip sampling, bracket = everything
    constexpr int margin = 113;  // 5 sigma of the sample-median rank for m = 2048 (sqrt(m) / 2 per sigma)
    K s_lo = 0, s_hi = 0;
    if (m) {
        const long long stride = n / m;
        for (int i = threadIdx.x; i < m; i += blockDim.x) s[i] = KeyOf<T>::key(xf[(long long)i * stride + stride / 2]);
        __syncthreads();
        s_lo = smem_select<K>(s, m, (unsigned long long)(m / 2 - margin), bins, sel);
        s_hi = smem_select<K>(s, m, (unsigned long long)(m / 2 + margin), bins, sel);
    }
    if (threadIdx.x == 0) {
        SelState<K> &st = ws->st;
        st.k_lo = (unsigned long long)((n - 1) / 2);
        st.k_hi = (unsigned long long)(n / 2);
        st.acc_below = 0;
        st.n_collected = 0;
        st.done = 0;
        st.res_lo = st.res_hi = 0;
        st.acc_inside = 0;
        st.edge_a = st.edge_b = st.found_max = 0;
        st.found_min = KeyOf<T>::kmax;
        st.cur_n = 0;
        st.gcap = gcap;
        st.use_compact = 0;
        if (m) {
            st.lo = s_lo;
            st.hi = s_hi;
            st.fresh = 1;
            st.below = 0;
            st.mode = gcap ? 2 : 0;  // filter pass first when the workspace has a compact buffer
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


// Filter pass (mode 2): one streaming read of the plane.  Keys below the bracket are counted; values inside it are
// staged in shared memory (one shared-memory atomic per value) and flushed to the compact buffer in coalesced runs
// (one global atomic per flush).
constexpr int kStage = 6144;  // staged values per block; a block-iteration adds at most 256 threads x 16 values
template <typename T>
__global__ void __launch_bounds__(256) select_filter_kernel(const T *x, long long n, long long bstride,
                                                            Workspace<typename KeyOf<T>::type> *ws_all, T *compact_all) {
    pdl_launch_dependents();
    pdl_wait();  // launched with programmatic stream serialisation: nothing is read before the predecessor is complete
    using K = typename KeyOf<T>::type;
    Workspace<K> *ws = ws_all + blockIdx.y;
    if (ws->st.done || ws->st.mode != 2) return;
    const T *xf = x + (long long)blockIdx.y * bstride;
    const unsigned long long gcap = ws->st.gcap;
    T *compact = compact_all + (size_t)blockIdx.y * gcap;
    const K lo = ws->st.lo, hi = ws->st.hi;
    extern __shared__ __align__(16) unsigned char stage_raw[];
    T *stage = reinterpret_cast<T *>(stage_raw);
    __shared__ unsigned int scount;
    __shared__ unsigned long long sbase;
    __shared__ unsigned long long blk_below, blk_inside;
    if (threadIdx.x == 0) { scount = 0; blk_below = 0; blk_inside = 0; }
    __syncthreads();
    unsigned long long below = 0;
    constexpr int V = VecOf<T>::V;
    constexpr int U = 16 / V;  // vectors per thread per iteration: 16 values
    const long long nvec = n / V;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(xf) & 15u) == 0);
    // A thread classifies its (up to) 16 values, the warp prefix-sums the in-bracket counts and reserves its run of the
    // staging buffer with ONE shared-memory atomic (one atomic per value on the block's counter serialised the 32 lanes
    // of every warp: 34 us per 64 MiB plane), then every thread writes its values behind its prefix.
    const unsigned lane = threadIdx.x & 31u;
    auto stage_values = [&](const T *vals, int cnt_vals) {  // cnt_vals <= 16, warp-convergent call
        unsigned inside = 0;  // bit j: value j is inside the bracket
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (j < cnt_vals) {
                const K key = KeyOf<T>::key(vals[j]);
                below += key < lo;
                inside |= (key >= lo && key <= hi) ? (1u << j) : 0u;
            }
        }
        const unsigned c = __popc(inside);
        unsigned inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += v;
        }
        const unsigned total = __shfl_sync(0xffffffffu, inc, 31);
        unsigned base = 0;
        if (lane == 31 && total) base = atomicAdd(&scount, total);
        base = __shfl_sync(0xffffffffu, base, 31);
        unsigned pos = base + inc - c;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if ((inside >> j) & 1u) stage[pos++] = vals[j];  // < kStage: see flush()
    };
    auto flush = [&]() {  // block-uniform call
        __syncthreads();
        const unsigned int cnt = scount;
        if (cnt > (unsigned)(kStage - 4096)) {
            if (threadIdx.x == 0) {
                sbase = atomicAdd(&ws->st.acc_inside, (unsigned long long)cnt);
            }
            __syncthreads();
            const unsigned long long base = sbase;
            for (unsigned int i = threadIdx.x; i < cnt; i += blockDim.x)
                if (base + i < gcap) compact[base + i] = stage[i];
            __syncthreads();
            if (threadIdx.x == 0) scount = 0;
            __syncthreads();
        }
    };
    const long long nth = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec_ok) {
        // block-uniform trip count so that the shuffles and the flush barriers are reached by every thread
        const long long per_it = nth * U;
        const long long iters = (nvec + per_it - 1) / per_it;
        for (long long it = 0; it < iters; ++it) {
            T vals[16];
            Pack<T, V> pk[U];
            int present = 0;  // the U vectors of a thread are ordered: the present ones come first
            const long long i0 = it * per_it + tid;
            if (i0 + (U - 1) * nth < nvec) {  // all U vectors exist (every iteration but the last): loads back to back
#pragma unroll
                for (int u = 0; u < U; ++u) pk[u] = ld_vec(xf + (i0 + u * nth) * V);
                present = 16;
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (i0 + u * nth < nvec) { pk[u] = ld_vec(xf + (i0 + u * nth) * V); present += V; }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int e = 0; e < V; ++e) vals[u * V + e] = pk[u].v[e];  // absent slots are never looked at
            stage_values(vals, present);  // warp-convergent: the shuffles inside use the full mask
            flush();
        }
        if (blockIdx.x == 0 && threadIdx.x < 32) {  // < V leftover values: one warp, one value per lane at most
            T vals[16];
            const long long i2 = nvec * V + threadIdx.x;
            vals[0] = i2 < n ? xf[i2] : T(0);
            stage_values(vals, i2 < n ? 1 : 0);
        }
    } else {
        const long long per_it = nth * 16;
        const long long iters = (n + per_it - 1) / per_it;
        for (long long it = 0; it < iters; ++it) {
            T vals[16];
            int present = 0;
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const long long i = it * per_it + u * nth + tid;
                vals[u] = i < n ? xf[i] : T(0);
                present += i < n ? 1 : 0;
            }
            stage_values(vals, present);
            flush();
        }
    }
    // final flush of whatever is staged
    __syncthreads();
    {
        const unsigned int cnt = scount;
        if (cnt) {
            if (threadIdx.x == 0) sbase = atomicAdd(&ws->st.acc_inside, (unsigned long long)cnt);
            __syncthreads();
            const unsigned long long base = sbase;
            for (unsigned int i = threadIdx.x; i < cnt; i += blockDim.x)
                if (base + i < gcap) compact[base + i] = stage[i];
        }
    }
    for (int o = 16; o > 0; o >>= 1) below += __shfl_down_sync(0xffffffffu, below, o);
    if ((threadIdx.x & 31) == 0 && below) atomicAdd(&blk_below, below);
    __syncthreads();
    if (threadIdx.x == 0 && blk_below) atomicAdd(&ws->st.acc_below, blk_below);
}

template <typename T>
__global__ void __launch_bounds__(256) select_pass_kernel(const T *x, long long n, long long bstride,
                                                          Workspace<typename KeyOf<T>::type> *ws_all, const T *compact_all) {
    pdl_launch_dependents();
    pdl_wait();  // launched with programmatic stream serialisation: nothing is read before the predecessor is complete
    using K = typename KeyOf<T>::type;
    Workspace<K> *ws = ws_all + blockIdx.y;
    if (ws->st.done) return;
    const T *xf = x + (long long)blockIdx.y * bstride;
    if (ws->st.use_compact) {  // the filter pass left the candidates in the compact buffer
        xf = compact_all + (size_t)blockIdx.y * ws->st.gcap;
        n = (long long)ws->st.cur_n;
    }
    // Every block that takes part merges a 2048-bin histogram into the global one with atomics, so a pass over few
    // values must not use the whole grid (1184 blocks x ~1100 non-empty bins were 1.3 M global atomics for the 1.85 M
    // values of a compact pass: 19 us).  At least 16384 values per block; the grid-stride loops below use `nblk`.
    long long nblk = (n + 16383) / 16384;
    if (nblk > (long long)gridDim.x) nblk = gridDim.x;
    if (nblk < 1) nblk = 1;
    if ((long long)blockIdx.x >= nblk) return;
    const K lo = ws->st.lo, hi = ws->st.hi;
    const int shift = ws->st.shift, mode = ws->st.mode;
    __shared__ unsigned int h[kBins];
    __shared__ unsigned long long blk_below;
    if (mode == 0) {
        for (int i = threadIdx.x; i < kBins; i += blockDim.x) h[i] = 0;
    }
    if (threadIdx.x == 0) blk_below = 0;
    __syncthreads();
    unsigned long long below = 0;
    constexpr int V = VecOf<T>::V;
    const long long nvec = n / V;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(xf) & 15u) == 0);
    const K edge_a = ws->st.edge_a, edge_b = ws->st.edge_b;
    K loc_max = 0, loc_min = KeyOf<T>::kmax;
    auto visit = [&](T v) {
        const K key = KeyOf<T>::key(v);
        if (mode == 3) {
            if (key <= edge_a && key > loc_max) loc_max = key;
            if (key >= edge_b && key < loc_min) loc_min = key;
        } else if (key < lo) {
            ++below;
        } else if (key <= hi) {
            if (mode == 0) {
                const int b = (int)((key - lo) >> shift);
                atomicAdd(&h[b], 1u);  // one atomic per key (round 1 also tracked per-bin min / max: three)
            } else {
                const unsigned int pos = atomicAdd(&ws->st.n_collected, 1u);
                if (pos < (unsigned)kCollectCap) ws->collect[pos] = key;
            }
        }
    };
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = nblk * blockDim.x;
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
    if (mode == 3) {
        for (int o = 16; o > 0; o >>= 1) {
            const K a = __shfl_down_sync(0xffffffffu, loc_max, o), b = __shfl_down_sync(0xffffffffu, loc_min, o);
            loc_max = a > loc_max ? a : loc_max;
            loc_min = b < loc_min ? b : loc_min;
        }
        if ((threadIdx.x & 31) == 0) {
            if (loc_max) atomicMax(&ws->st.found_max, loc_max);  // a key of 0 needs no update: found_max starts at 0
            if (loc_min != KeyOf<T>::kmax) atomicMin(&ws->st.found_min, loc_min);
        }
        return;
    }
    // block reduction of the below counter
    for (int o = 16; o > 0; o >>= 1) below += __shfl_down_sync(0xffffffffu, below, o);
    if ((threadIdx.x & 31) == 0 && below) atomicAdd(&blk_below, below);
    __syncthreads();
    if (threadIdx.x == 0 && blk_below) atomicAdd(&ws->st.acc_below, blk_below);
    if (mode == 0) {
        for (int i = threadIdx.x; i < kBins; i += blockDim.x) {
            if (h[i]) atomicAdd(&ws->hist[i], h[i]);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(1024) select_decide_kernel(long long n, Workspace<typename KeyOf<T>::type> *ws_all,
                                                             T *out_median, double *out_noise, double sigma_e0,
                                                             int last) {
    pdl_launch_dependents();
    pdl_wait();  // launched with programmatic stream serialisation: nothing is read before the predecessor is complete
    using K = typename KeyOf<T>::type;
    Workspace<K> *ws = ws_all + blockIdx.x;
    SelState<K> &st = ws->st;
    // one buffer, two uses: candidate keys (collect mode) or the cumulative histogram (count mode)
    __shared__ unsigned long long buf[kCollectCap];
    K *s = reinterpret_cast<K *>(buf);
    unsigned long long *cum = buf;
    __shared__ int fin;
    __shared__ unsigned int sel_bins[256];
    __shared__ unsigned long long sel_out[2];
    if (st.done) return;
    if (threadIdx.x == 0) fin = 0;
    __syncthreads();

    if (st.mode == 3) {
        // ---- edge pass finished: rank k_lo is the largest key of its bin, rank k_hi the smallest key of a later bin ----
        if (threadIdx.x == 0) {
            st.res_lo = st.found_max;
            st.res_hi = st.found_min;
            st.done = 1;
            fin = 1;
        }
    } else if (st.mode == 2) {
        // ---- filter pass finished: the ranks lie inside the bracket (then the later passes read the compact buffer) or
        //      the sampled bracket missed / the buffer overflowed (then they count over the plane as before) -----------
        if (threadIdx.x == 0) {
            const unsigned long long below = st.acc_below, inside = st.acc_inside;
            st.use_compact = 0;
            st.fresh = 0;
            if (st.k_lo < below) {
                // both ranks below the bracket (k_hi may be the first key of it: keep lo inside)
                st.hi = st.lo;
                st.lo = 0;
                st.below = 0;
                st.mode = 0;
            } else if (st.k_hi >= below + inside) {
                // both ranks above it (k_lo may be the last key of it: keep hi inside); the keys below the new bracket
                // are counted by the next pass
                st.lo = st.hi;
                st.hi = KeyOf<T>::kmax;
                st.below = 0;
                st.fresh = 1;
                st.mode = 0;
            } else if (inside > st.gcap) {
                // heavy ties / a flat distribution around the median: count over the plane with the same bracket
                st.below = below;
                st.mode = 0;
            } else {
                // the compact buffer holds exactly the keys of [lo, hi]: ranks relative to it
                st.use_compact = 1;
                st.cur_n = inside;
                st.k_lo -= below;
                st.k_hi -= below;
                st.below = 0;
                st.mode = (inside <= (unsigned long long)kCollectCap) ? 1 : 0;
            }
            st.shift = shift_for<K>(st.lo, st.hi);
            st.acc_below = 0;
            st.n_collected = 0;
        }
    } else if (st.mode == 1) {
        // ---- collect pass finished: sort the candidates and read the two ranks off --------------------------
        const unsigned int cnt = st.n_collected;
        const unsigned long long below = st.fresh ? st.acc_below : st.below;
        // cnt <= kCollectCap is guaranteed by construction (mode 1 is only entered with a known small count)
        for (int i = threadIdx.x; i < (int)cnt; i += blockDim.x) s[i] = ws->collect[i];
        __syncthreads();
        const K r_lo = smem_select<K>(s, (int)cnt, st.k_lo - below, sel_bins, sel_out);
        const K r_hi = (st.k_hi == st.k_lo) ? r_lo : smem_select<K>(s, (int)cnt, st.k_hi - below, sel_bins, sel_out);
        if (threadIdx.x == 0) {
            st.res_lo = r_lo;
            st.res_hi = r_hi;
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
                if (st.shift == 0) {
                    // one key value per bin: the bins ARE the keys of the two ranks
                    st.res_lo = st.lo + (K)b_lo;
                    st.res_hi = st.lo + (K)b_hi;
                    st.done = 1;
                    fin = 1;
                } else if (b_lo != b_hi) {
                    // the two middle ranks straddle a bin edge (possibly across empty bins: bimodal data): rank k_lo is
                    // the largest key of bin b_lo, rank k_hi the smallest key of bin b_hi -- one min / max pass
                    st.edge_a = st.lo + ((K)b_lo << st.shift) + (((K)1 << st.shift) - 1);
                    st.edge_b = st.lo + ((K)b_hi << st.shift);
                    st.found_max = 0;
                    st.found_min = KeyOf<T>::kmax;
                    st.mode = 3;
                } else {
                    // narrow the bracket to the key range of the bin holding both ranks; collect once few keys are left
                    const K nlo = st.lo + ((K)b_lo << st.shift);
                    K nhi = nlo + (((K)1 << st.shift) - 1);
                    if (nhi > st.hi || nhi < nlo) nhi = st.hi;  // last bin is cut by hi (or the add wrapped)
                    const unsigned long long nbelow = (b_lo > 0) ? cum[b_lo - 1] : below;
                    const unsigned long long inside = cum[b_lo] - nbelow;
                    st.lo = nlo;
                    st.hi = nhi;
                    st.below = nbelow;
                    st.mode = (inside <= (unsigned long long)kCollectCap) ? 1 : 0;
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
            for (int i = threadIdx.x; i < kBins; i += blockDim.x) ws->hist[i] = 0;
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
    pdl_launch_dependents();
    pdl_wait();  // launched with programmatic stream serialisation: nothing is read before the predecessor is complete
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
    pdl_launch_dependents();
    pdl_wait();
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

template <typename K> static size_t median_state_bytes(int batch) {
    return (sizeof(Workspace<K>) * (size_t)batch + 255) & ~(size_t)255;  // the compact buffers follow, 256-byte aligned
}

template <typename T>
static int median_impl(const void *x, long long n, int batch, long long bstride, void *out_median, double *out_noise,
                       double sigma_e0, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    using K = typename KeyOf<T>::type;
    auto *ws = reinterpret_cast<Workspace<K> *>(workspace);
    const T *xp = reinterpret_cast<const T *>(x);
    const size_t state = median_state_bytes<K>(batch);
    if (workspace_bytes < sizeof(Workspace<K>) * (size_t)batch) return WB_EINVAL_ARG;
    // compact buffer: whatever the caller provides beyond the selection state, split evenly over the frames
    unsigned long long gcap = 0;
    if (workspace_bytes > state) gcap = ((workspace_bytes - state) / (size_t)batch / sizeof(T)) & ~(unsigned long long)15;
    if (gcap < 4096) gcap = 0;
    T *compact = reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(workspace) + state);
    launch_pdl_v(select_init_kernel<T>, dim3((unsigned)batch), dim3(1024), 0, st, xp, n, bstride, ws, gcap);
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
    // Launch budget (converged steps exit at once).  A counting pass narrows the key bracket by 11 bits and finishes when
    // a bin is one key wide: at most 3 (fp32) / 6 (fp64) of them, plus one collect / edge pass; without the filter pass
    // one more may be spent on a sampled bracket that missed the ranks.
    const int passes = gcap ? (sizeof(T) == 4 ? 4 : 7) : (sizeof(T) == 4 ? 5 : 9);
    if (gcap) {
        long long fblocks = (n / 16 + 255) / 256;  // 16 values per thread per iteration
        if (fblocks > 8LL * sms) fblocks = 8LL * sms;
        if (fblocks < 1) fblocks = 1;
        const size_t smem = (size_t)kStage * sizeof(T);
        static bool configured[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !configured[dev]) {
            cudaError_t e = cudaFuncSetAttribute(select_filter_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            if (dev >= 0 && dev < 64) configured[dev] = true;
        }
        launch_pdl_v(select_filter_kernel<T>, dim3((unsigned)fblocks, (unsigned)batch), dim3(256), smem, st, xp, n, bstride, ws, compact);
        launch_pdl_v(select_decide_kernel<T>, dim3((unsigned)batch), dim3(1024), 0, st, n, ws, reinterpret_cast<T *>(out_median), out_noise, sigma_e0, 0);
    }
    for (int p = 0; p < passes; ++p) {
        launch_pdl_v(select_pass_kernel<T>, dim3((unsigned)blocks, (unsigned)batch), dim3(256), 0, st, xp, n, bstride, ws, (const T *)compact);
        launch_pdl_v(select_decide_kernel<T>, dim3((unsigned)batch), dim3(1024), 0, st, n, ws, reinterpret_cast<T *>(out_median),
                     out_noise, sigma_e0, (int)(p == passes - 1));
    }
    return launch_status();
}

}  // namespace wb

extern "C" {

size_t wb_abs_median_workspace_bytes(int dtype, int batch, long long n) {
    if (batch < 1) batch = 1;
    const size_t esz = dtype == WB_F64 ? 8 : 4;
    const size_t state = dtype == WB_F64 ? wb::median_state_bytes<unsigned long long>(batch) : wb::median_state_bytes<uint32_t>(batch);
    // compact buffer of the filter pass: a quarter of the plane per frame (the 5-sigma bracket of the 2048-key sample
    // holds ~11 % of the keys); n <= 0: selection state only (every pass then reads the plane)
    size_t cap = n > 0 ? (((size_t)n / 4 + 15) & ~(size_t)15) : 0;
    if (cap < 4096) cap = 0;
    return state + cap * esz * (size_t)batch;
}

int wb_abs_median(const void *x, long long n, int batch, long long bstride, int dtype, void *out_median,
                  double *out_noise, double sigma_e0, void *workspace, size_t workspace_bytes, void *stream) {
    if (dtype != WB_F32 && dtype != WB_F64) return WB_EINVAL_DTYPE;
    if (n < 1 || batch < 1 || batch > 65535) return WB_EINVAL_SHAPE;
    if (!x || !workspace || (!out_median && !out_noise)) return WB_EINVAL_POINTER;
    if (out_noise && !(sigma_e0 > 0)) return WB_EINVAL_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == WB_F32
               ? wb::median_impl<float>(x, n, batch, bstride, out_median, out_noise, sigma_e0, workspace, workspace_bytes, st)
               : wb::median_impl<double>(x, n, batch, bstride, out_median, out_noise, sigma_e0, workspace, workspace_bytes, st);
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
        wb::launch_pdl_v(wb::moments_partial_kernel<float>, grid, dim3(256), 0, st, reinterpret_cast<const float *>(x), n, bstride, partial, shift);
    else
        wb::launch_pdl_v(wb::moments_partial_kernel<double>, grid, dim3(256), 0, st, reinterpret_cast<const double *>(x), n, bstride, partial, shift);
    return wb::launch_pdl_v(wb::moments_final_kernel, dim3((unsigned)batch), dim3(256), 0, st, (const double *)partial, (const double *)shift, n, (int)wb::kMomBlocks, out);
}

}  // extern "C"
