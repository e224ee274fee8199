// K1 -- one scale of the plain à trous cascade, fused: c_{s+1} = S_s[c_s] and w_s = c_s - c_{s+1} in one pass.
//
// Replaces watroo/wavelets.py:35-45 (cv2.filter2D with the dense dilated kernel of :191-197, BORDER_REFLECT) and the
// in-place subtraction of :442.  HBM-bound: 3*sizeof(T) algorithmic bytes per pixel (read c_s, write c_{s+1}, w_s).
//
// Design ("K1"):
//   * The dilated separable filter only couples pixels that are 2^s apart, so the image splits into 2^s row
//     "chains" (rows r, r+d, r+2d, ... with d = 2^s).  A thread block owns a segment of one chain over one column
//     strip and walks down it with a sliding window: every input row is fetched ONCE per block, as one contiguous
//     TMA bulk copy (cp.async.bulk -> UBLKCP) into a shared-memory ring guarded by full/empty mbarriers, issued by a
//     dedicated producer warp.  The column halo costs (taps-1) extra row loads per segment whatever the dilation,
//     so deep scales are as cheap as shallow ones (no 2^s-wide halo, no dense kernel).
//   * Row pass: each consumer thread owns NG 16-byte vectors of columns and reads its taps x + k*d straight from
//     the staged row with LDS.128 (conflict-free: consecutive threads read consecutive vectors).  When the strip is
//     the whole row the x halo is free -- the symmetric border is an index reflection inside the same staged row
//     (a reflected aligned vector is the mirrored vector read backwards).
//   * Column pass: the last `taps` row-filtered vectors live in a register ring; c_{s+1} comes out of registers,
//     w_s = raw - c_{s+1} uses the raw centre row still resident in the ring, and both are written with 128-bit
//     coalesced stores (w_s with an evict-first hint so it does not push c_{s+1} out of L2 before scale s+1 reads it).
//   * Anything the vector path cannot take (W % V != 0, unaligned pointers, 2^s*c > W so more than one reflection)
//     goes to a generic gather kernel with the full modular reflection.
#include "common.cuh"

namespace wb {

struct ScaleParams {
    const void *in;
    void *out_c;
    void *out_w;
    int H, W, d;
    long long in_pitch, in_bstride, c_pitch, c_bstride, w_pitch, w_bstride;
    int wt;          // strip width in elements (= consumer threads * V * NG)
    int n_strips;    // column strips per row
    int seg;         // chain rows produced per thread block
    int n_seg;       // segments per chain
    int slots;       // depth of the shared-memory row ring
    int row_stride;  // elements per ring slot
    int halo_al;     // x halo kept in shared memory on each side of a strip (multiple of V)
};

// Load the V-wide aligned vector of columns starting at logical column p (p % V == 0, -W <= p < 2W) of the staged
// row `srow` (which holds global columns [lo, hi)), through the symmetric border.
template <typename T, int V>
__device__ __forceinline__ Pack<T, V> load_tap(const T *srow, int p, int W, int lo) {
    const bool left = p < 0, right = p >= W;
    const int q = left ? (-V - p) : (right ? (2 * W - V - p) : p);
    Pack<T, V> t = ld_vec(srow + (q - lo));
    if (left || right) {
#pragma unroll
        for (int e = 0; e < V / 2; ++e) {
            T a = t.v[e];
            t.v[e] = t.v[V - 1 - e];
            t.v[V - 1 - e] = a;
        }
    }
    return t;
}

// Row pass for one vector of columns starting at x.  DMODE == 0: d % V == 0, taps are whole aligned vectors.
// DMODE == d in {1, 2}: d < V, all taps lie in the previous/current/next vector.  SQUARE: filter the squares.
template <typename T, int TAPS, int DMODE, bool SQUARE>
__device__ __forceinline__ Pack<T, VecOf<T>::V> row_pass(const T *srow, int x, int d, int W, int lo) {
    constexpr int V = VecOf<T>::V;
    constexpr int C = TAPS / 2;
    Pack<T, V> acc;
    if constexpr (DMODE == 0) {
#pragma unroll
        for (int k = 0; k < TAPS; ++k) {
            Pack<T, V> t = load_tap<T, V>(srow, x + (k - C) * d, W, lo);
#pragma unroll
            for (int e = 0; e < V; ++e) {
                T v = SQUARE ? t.v[e] * t.v[e] : t.v[e];
                acc.v[e] = (k == 0) ? Taps<T, TAPS>::h(0) * v : fma_t<T>(Taps<T, TAPS>::h(k), v, acc.v[e]);
            }
        }
    } else {
        T win[3 * V];
        Pack<T, V> a = load_tap<T, V>(srow, x - V, W, lo);
        Pack<T, V> b = load_tap<T, V>(srow, x, W, lo);
        Pack<T, V> c = load_tap<T, V>(srow, x + V, W, lo);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            win[e] = SQUARE ? a.v[e] * a.v[e] : a.v[e];
            win[V + e] = SQUARE ? b.v[e] * b.v[e] : b.v[e];
            win[2 * V + e] = SQUARE ? c.v[e] * c.v[e] : c.v[e];
        }
#pragma unroll
        for (int e = 0; e < V; ++e) {
#pragma unroll
            for (int k = 0; k < TAPS; ++k) {
                T v = win[V + e + (k - C) * DMODE];
                acc.v[e] = (k == 0) ? Taps<T, TAPS>::h(0) * v : fma_t<T>(Taps<T, TAPS>::h(k), v, acc.v[e]);
            }
        }
    }
    return acc;
}

template <typename T, int TAPS, int DMODE, int NG>
__global__ void __launch_bounds__(544, 1) atrous_rows_kernel(const ScaleParams p) {
    constexpr int V = VecOf<T>::V;
    constexpr int C = TAPS / 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *rows = reinterpret_cast<T *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)p.slots * p.row_stride * sizeof(T));
    uint64_t *empty = full + p.slots;

    const int nt = blockDim.x - 32;  // consumer threads; the last warp is the TMA producer
    const int nwc = nt >> 5;
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    // block -> (strip, chain residue r, segment g); r varies fastest so concurrent blocks touch adjacent rows
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
        // ---------------- producer warp: one lane streams the chain rows into the ring ----------------
        if (lane == 0) {
            const T *src = reinterpret_cast<const T *>(p.in) + (long long)frame * p.in_bstride + lo;
            int slot = 0;
            uint32_t round = 0;
            for (int j = 0; j < n_load; ++j) {
                if (round > 0) mbar_wait(&empty[slot], (round - 1) & 1);
                const int y = reflect_any((long long)r + (long long)(i0 - C + j) * p.d, p.H);
                mbar_arrive_expect_tx(&full[slot], row_bytes);
                tma_load_1d(rows + (size_t)slot * p.row_stride, src + (long long)y * p.in_pitch, row_bytes,
                            &full[slot]);
                if (++slot == p.slots) { slot = 0; ++round; }
            }
        }
        return;
    }

    // ---------------- consumer warps ----------------
    T *out_c = reinterpret_cast<T *>(p.out_c);
    T *out_w = reinterpret_cast<T *>(p.out_w);
    if (out_c) out_c += (long long)frame * p.c_bstride;
    if (out_w) out_w += (long long)frame * p.w_bstride;

    int xg[NG];
    bool act[NG];
#pragma unroll
    for (int q = 0; q < NG; ++q) {
        xg[q] = x0 + (q * nt + tid) * V;
        act[q] = xg[q] < p.W;
    }

    Pack<T, V> ring[TAPS][NG];
#pragma unroll
    for (int k = 0; k < TAPS; ++k)
#pragma unroll
        for (int q = 0; q < NG; ++q)
#pragma unroll
            for (int e = 0; e < V; ++e) ring[k][q].v[e] = T(0);

    int slot = 0, cslot = 0, rslot = 0;  // slot of row j, of the centre row j-C, of the row being released
    uint32_t parity = 0;
    for (int j = 0; j < n_load; ++j) {
        mbar_wait(&full[slot], parity);
        const T *srow = rows + (size_t)slot * p.row_stride;

#pragma unroll
        for (int k = 0; k + 1 < TAPS; ++k)
#pragma unroll
            for (int q = 0; q < NG; ++q) ring[k][q] = ring[k + 1][q];
#pragma unroll
        for (int q = 0; q < NG; ++q)
            if (act[q]) ring[TAPS - 1][q] = row_pass<T, TAPS, DMODE, false>(srow, xg[q], p.d, p.W, lo);

        if (j >= 2 * C) {
            const long long y = (long long)r + (long long)(i0 + j - 2 * C) * p.d;
            const T *crow = rows + (size_t)cslot * p.row_stride;
#pragma unroll
            for (int q = 0; q < NG; ++q) {
                if (!act[q]) continue;
                Pack<T, V> c;
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    T a = Taps<T, TAPS>::h(0) * ring[0][q].v[e];
#pragma unroll
                    for (int k = 1; k < TAPS; ++k) a = fma_t<T>(Taps<T, TAPS>::h(k), ring[k][q].v[e], a);
                    c.v[e] = a;
                }
                if (out_c) st_vec(out_c + y * p.c_pitch + xg[q], c);
                if (out_w) {
                    Pack<T, V> raw = ld_vec(crow + (xg[q] - lo));
#pragma unroll
                    for (int e = 0; e < V; ++e) raw.v[e] -= c.v[e];
                    st_vec_cs(out_w + y * p.w_pitch + xg[q], raw);
                }
            }
        }
        if (j >= C) {
            // the raw row j-C is not needed any more: hand its slot back to the producer
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[rslot]);
            if (++rslot == p.slots) rslot = 0;
            if (++cslot == p.slots) cslot = 0;
        }
        if (++slot == p.slots) { slot = 0; parity ^= 1; }
    }
}

// Generic path: one thread per output pixel, full modular reflection, any shape / alignment.
template <typename T, int TAPS>
__global__ void __launch_bounds__(256) atrous_generic_kernel(const ScaleParams p) {
    constexpr int C = TAPS / 2;
    const long long n = (long long)p.H * p.W;
    const int frame = blockIdx.y;
    const T *in = reinterpret_cast<const T *>(p.in) + (long long)frame * p.in_bstride;
    T *out_c = reinterpret_cast<T *>(p.out_c);
    T *out_w = reinterpret_cast<T *>(p.out_w);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(idx / p.W), x = (int)(idx % p.W);
        int xs[TAPS];
#pragma unroll
        for (int k = 0; k < TAPS; ++k) xs[k] = reflect_any((long long)x + (long long)(k - C) * p.d, p.W);
        T acc = T(0);
#pragma unroll
        for (int i = 0; i < TAPS; ++i) {
            const T *row = in + (long long)reflect_any((long long)y + (long long)(i - C) * p.d, p.H) * p.in_pitch;
            T ra = Taps<T, TAPS>::h(0) * row[xs[0]];
#pragma unroll
            for (int k = 1; k < TAPS; ++k) ra = fma_t<T>(Taps<T, TAPS>::h(k), row[xs[k]], ra);
            acc = (i == 0) ? Taps<T, TAPS>::h(0) * ra : fma_t<T>(Taps<T, TAPS>::h(i), ra, acc);
        }
        if (out_c) out_c[(long long)frame * p.c_bstride + (long long)y * p.c_pitch + x] = acc;
        if (out_w)
            out_w[(long long)frame * p.w_bstride + (long long)y * p.w_pitch + x] = in[(long long)y * p.in_pitch + x] - acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Host: configuration and dispatch
// ---------------------------------------------------------------------------------------------------------------
struct K1Config { int nt, ng, slots, seg; };

static K1Config g_override[32];
static bool g_override_set[32];

static constexpr int kMaxSmem = 227 * 1024;

static int device_sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

static bool fast_path_ok(const ScaleParams &p, int taps, int esize) {
    const int V = 16 / esize;
    const int c = taps / 2;
    if (p.W % V || p.W < 8 * V) return false;
    if ((long long)c * p.d > p.W) return false;  // more than one reflection in x
    if (p.in_pitch % V || p.in_bstride % V || !aligned16(p.in)) return false;
    if (p.out_c && (p.c_pitch % V || p.c_bstride % V || !aligned16(p.out_c))) return false;
    if (p.out_w && (p.w_pitch % V || p.w_bstride % V || !aligned16(p.out_w))) return false;
    return true;
}

// Fill in the strip/segment/ring geometry.  Returns false if no geometry fits in shared memory.
static bool plan_fast(ScaleParams &p, int taps, int esize, int batch, int scale, K1Config *cfg_out) {
    const int V = 16 / esize;
    const int c = taps / 2;
    K1Config cfg;
    if (scale < 32 && g_override_set[scale]) {
        cfg = g_override[scale];
    } else {
        cfg.ng = 2;
        const int vecs = (p.W + V - 1) / V;
        cfg.nt = round_up((vecs + cfg.ng - 1) / cfg.ng, 32);
        // a strip is at most 16 KiB of row data (4096 fp32 / 2048 fp64 columns)
        const int nt_cap = 16384 / (16 * cfg.ng);
        if (cfg.nt > nt_cap) cfg.nt = nt_cap;
        if (cfg.nt > 512) cfg.nt = 512;
        cfg.slots = 8;
        cfg.seg = 0;  // decided below
    }
    if (cfg.nt % 32 || cfg.nt < 32 || cfg.nt > 512 || (cfg.ng != 1 && cfg.ng != 2)) return false;
    p.wt = cfg.nt * V * cfg.ng;
    p.n_strips = (p.W + p.wt - 1) / p.wt;
    p.halo_al = round_up(c * p.d, V);
    long long rs = (long long)p.wt + 2LL * p.halo_al;
    if (rs > p.W) rs = p.W;
    p.row_stride = (int)rs;
    const int min_slots = c + 2;
    int slots = cfg.slots;
    while (slots > min_slots && (long long)slots * p.row_stride * esize + 16LL * slots > kMaxSmem) --slots;
    if ((long long)slots * p.row_stride * esize + 16LL * slots > kMaxSmem || slots < min_slots) return false;
    p.slots = slots;
    const int n_max = (p.H + p.d - 1) / p.d;  // longest chain
    int seg = cfg.seg;
    if (seg <= 0) {
        // aim at ~2 blocks per SM in flight, but never let the (taps-1)-row halo exceed ~25 % of a segment
        const long long chains = (long long)p.n_strips * (p.d < p.H ? p.d : p.H) * batch;
        const long long target = 2LL * device_sm_count();
        long long per_chain = (target + chains - 1) / chains;
        if (per_chain < 1) per_chain = 1;
        seg = (int)((n_max + per_chain - 1) / per_chain);
        if (seg < 8 * c) seg = 8 * c;
    }
    if (seg > n_max) seg = n_max;
    p.seg = seg;
    p.n_seg = (n_max + seg - 1) / seg;
    cfg.slots = slots;
    cfg.seg = seg;
    if (cfg_out) *cfg_out = cfg;
    return true;
}

template <typename T, int TAPS, int DMODE, int NG>
static int launch_rows(const ScaleParams &p, int batch, int nt, cudaStream_t st) {
    auto kern = atrous_rows_kernel<T, TAPS, DMODE, NG>;
    const size_t smem = (size_t)p.slots * p.row_stride * sizeof(T) + 16 * (size_t)p.slots;
    static bool configured[64] = {};  // per instantiation, per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    dim3 grid((unsigned)((long long)p.n_strips * p.d * p.n_seg), (unsigned)batch);
    kern<<<grid, nt + 32, smem, st>>>(p);
    return launch_status();
}

template <typename T, int TAPS>
static int dispatch(ScaleParams &p, int batch, int scale, cudaStream_t st) {
    constexpr int V = VecOf<T>::V;
    K1Config cfg;
    if (fast_path_ok(p, TAPS, (int)sizeof(T)) && plan_fast(p, TAPS, (int)sizeof(T), batch, scale, &cfg)) {
        const int dmode = (p.d % V == 0) ? 0 : p.d;
#define WB_LAUNCH(DM)                                                               \
    (cfg.ng == 1 ? launch_rows<T, TAPS, DM, 1>(p, batch, cfg.nt, st) : launch_rows<T, TAPS, DM, 2>(p, batch, cfg.nt, st))
        if (dmode == 0) return WB_LAUNCH(0);
        if (dmode == 1) return WB_LAUNCH(1);
        if constexpr (V == 4) {
            if (dmode == 2) return WB_LAUNCH(2);
        }
#undef WB_LAUNCH
    }
    const long long n = (long long)p.H * p.W;
    long long blocks = (n + 255) / 256;
    const long long cap = 32LL * device_sm_count();
    if (blocks > cap) blocks = cap;
    atrous_generic_kernel<T, TAPS><<<dim3((unsigned)blocks, (unsigned)batch), 256, 0, st>>>(p);
    return launch_status();
}

static int scale_impl(const void *in, void *out_c, void *out_w, int batch, int H, int W, long long in_pitch,
                      long long in_bstride, long long c_pitch, long long c_bstride, long long w_pitch,
                      long long w_bstride, int scale, int taps, int dtype, cudaStream_t st) {
    int rc = check_common(batch, H, W, taps, dtype);
    if (rc) return rc;
    if (scale < 0 || scale > 30) return WB_EINVAL_SCALE;
    if (!in || (!out_c && !out_w)) return WB_EINVAL_POINTER;
    if (in == out_c || in == out_w) return WB_EINVAL_POINTER;
    if (in_pitch < W || (out_c && c_pitch < W) || (out_w && w_pitch < W)) return WB_EINVAL_ARG;
    ScaleParams p;
    memset(&p, 0, sizeof(p));
    p.in = in; p.out_c = out_c; p.out_w = out_w;
    p.H = H; p.W = W; p.d = 1 << scale;
    p.in_pitch = in_pitch; p.in_bstride = in_bstride;
    p.c_pitch = c_pitch; p.c_bstride = c_bstride;
    p.w_pitch = w_pitch; p.w_bstride = w_bstride;
    if (dtype == WB_F32)
        return taps == 3 ? dispatch<float, 3>(p, batch, scale, st) : dispatch<float, 5>(p, batch, scale, st);
    return taps == 3 ? dispatch<double, 3>(p, batch, scale, st) : dispatch<double, 5>(p, batch, scale, st);
}

}  // namespace wb

extern "C" {

int wb_atrous_scale_path(int H, int W, long long in_pitch, long long out_pitch, int scale, int taps, int dtype,
                         const void *in, const void *out_c, const void *out_w) {
    if (wb::check_common(1, H, W, taps, dtype) || scale < 0 || scale > 30) return -1;
    wb::ScaleParams p;
    memset(&p, 0, sizeof(p));
    p.in = in; p.out_c = const_cast<void *>(out_c); p.out_w = const_cast<void *>(out_w);
    p.H = H; p.W = W; p.d = 1 << scale;
    p.in_pitch = in_pitch; p.c_pitch = out_pitch; p.w_pitch = out_pitch;
    const int esize = wb::dtype_size(dtype);
    return (wb::fast_path_ok(p, taps, esize) && wb::plan_fast(p, taps, esize, 1, scale, nullptr)) ? 1 : 0;
}

// Tuning hook (benchmark sweeps only): override the K1 geometry of one scale; nt == 0 clears the override.
int wb_tune_k1(int scale, int nt, int ng, int slots, int seg) {
    if (scale < 0 || scale >= 32) return WB_EINVAL_SCALE;
    if (nt == 0) { wb::g_override_set[scale] = false; return WB_OK; }
    wb::g_override[scale] = wb::K1Config{nt, ng, slots, seg};
    wb::g_override_set[scale] = true;
    return WB_OK;
}

int wb_atrous_scale(const void *in, void *out_c, void *out_w, int batch, int H, int W, long long in_pitch,
                    long long in_bstride, long long out_c_pitch, long long out_c_bstride, long long out_w_pitch,
                    long long out_w_bstride, int scale, int taps, int dtype, void *stream) {
    return wb::scale_impl(in, out_c, out_w, batch, H, W, in_pitch, in_bstride, out_c_pitch, out_c_bstride,
                          out_w_pitch, out_w_bstride, scale, taps, dtype, (cudaStream_t)stream);
}

int wb_atrous_transform(const void *in, void *planes, void *scratch, int batch, int H, int W, long long in_pitch,
                        long long in_bstride, int levels, int taps, int dtype, void *stream) {
    int rc = wb::check_common(batch, H, W, taps, dtype);
    if (rc) return rc;
    if (levels < 0 || levels > 30) return WB_EINVAL_SCALE;
    if (!in || !planes || (levels > 1 && !scratch)) return WB_EINVAL_POINTER;
    if (in_pitch < W) return WB_EINVAL_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t es = (size_t)wb::dtype_size(dtype);
    const long long plane = (long long)H * W;
    const long long fstride = (long long)(levels + 1) * plane;  // frame stride inside `planes`
    char *pl = reinterpret_cast<char *>(planes);
    char *sc = reinterpret_cast<char *>(scratch);
    if (levels == 0) {
        cudaError_t e = cudaMemcpy2DAsync(pl, (size_t)W * es, in, (size_t)in_pitch * es, (size_t)W * es, (size_t)H,
                                          cudaMemcpyDeviceToDevice, st);
        for (int b = 1; b < batch && e == cudaSuccess; ++b)
            e = cudaMemcpy2DAsync(pl + (size_t)b * fstride * es, (size_t)W * es,
                                  reinterpret_cast<const char *>(in) + (size_t)b * in_bstride * es,
                                  (size_t)in_pitch * es, (size_t)W * es, (size_t)H, cudaMemcpyDeviceToDevice, st);
        return (int)e;
    }
    for (int s = 0; s < levels; ++s) {
        // c_s: the image itself for s == 0, else the scratch half written by the previous scale
        const void *src = (s == 0) ? in : (const void *)(sc + (size_t)((s - 1) & 1) * batch * plane * es);
        const long long src_pitch = (s == 0) ? in_pitch : W;
        const long long src_bstride = (s == 0) ? in_bstride : plane;
        // c_{s+1}: the other scratch half, or plane L of the output for the last scale
        const bool last = (s == levels - 1);
        void *dst_c = last ? (void *)(pl + (size_t)levels * plane * es) : (void *)(sc + (size_t)(s & 1) * batch * plane * es);
        const long long c_bstride = last ? fstride : plane;
        void *dst_w = pl + (size_t)s * plane * es;
        rc = wb::scale_impl(src, dst_c, dst_w, batch, H, W, src_pitch, src_bstride, W, c_bstride, W, fstride, s, taps,
                            dtype, st);
        if (rc) return rc;
    }
    return WB_OK;
}

}  // extern "C"
