// K1+K3 fused -- one WOW scale in ONE pass over HBM: c_{s+1} = S_s[c_s], w_s = c_s - c_{s+1},
// P = S_s[w_s^2], w'_s = w_s * significance(w_s) * (weight / sqrt(max(P, 1e-15))).
//
// Replaces, per scale, watroo/wavelets.py:35-45 + :442 (smooth + subtraction) AND the loop body of
// watroo/utils.py:177,193-203 (power = c**2, convolution(power), clamp, sqrt, significance, weighting).  The raw detail
// plane w_s never reaches HBM: 3*sizeof(T) bytes per pixel per scale (read c_s, write c_{s+1}, write w'_s) instead of
// the 5*sizeof(T) of K1 followed by K3.
//
// Same chain walk as K1 (see atrous_scale.cu): the local power uses the SAME dilation as the smooth, so it only couples
// rows of the same chain.  A thread block walks a chain segment once, with two specialised halves that form a
// pipeline through shared memory (each half owns every column vector of the row, NG vectors per thread):
//   SMOOTH warps (first half):  wait for input row j (TMA, mbarrier) -> row pass -> column pass (running partial
//       sums in registers) -> c_{s+1} row (128-bit store, only rows of the segment) -> w_s = raw - c_{s+1} into the
//       shared-memory w ring -> arrive "w row full"; release input row j-C.  Lane 0 of warp 0 is also the TMA loader.
//   POWER warps (second half):  wait "w row full" -> row pass over the squares -> column pass -> P; epilogue with the
//       raw w_s of the centre row (same ring) -> streaming 128-bit store of w'_s; release w row t-C.
// All hand-offs are mbarriers with one arrival per warp, so warps drift freely (no block-wide barrier in the loop), and
// with one ring of partial sums per thread both halves fit in 64 registers: 1024 threads = 32 warps per SM.
// Rows outside the image are symmetric reflections: the loader fetches the reflected rows; because the symmetric
// extension commutes with the symmetric smooth, c_{s+1} / w_s evaluated at such a virtual row equal their reflected
// values (up to the rounding of a reversed summation order), which is what the power filter needs at the border.
// The strip is always the whole row, so the x border is an index reflection inside the staged row for both row
// passes (same tap plan).  Away from the top/bottom border the arithmetic is exactly K1 followed by K3.
#include "pipeline.cuh"
#include "whiten.cuh"

namespace wb {

static constexpr int kInRing = 8;  // input rows staged by TMA (C+1 live, the rest prefetch)
static constexpr int kWRing = 4;   // raw w_s rows (read back by the writing thread C+1 steps later)
static constexpr int kW2Ring = 2;  // w_s^2 rows (cross-thread, guarded by the split-phase barriers)

template <typename T, int TAPS, int DMODE, int NG, bool HINTS>
__global__ void __launch_bounds__(512, 1) wow_rows_kernel(const ScaleParams p) {
    constexpr int V = VecOf<T>::V;
    constexpr int C = TAPS / 2;
    constexpr int NV = PlanSize<TAPS, DMODE>::NV;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t RB = (uint32_t)p.row_stride * (uint32_t)sizeof(T);  // bytes per staged row
    const uint32_t in_base = smem_u32(smem_raw);                       // kInRing rows of c_s (TMA)
    const uint32_t w_base = in_base + (uint32_t)kInRing * RB;           // kWRing rows of raw w_s
    const uint32_t w2_base = w_base + (uint32_t)kWRing * RB;            // kW2Ring rows of w_s^2
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)(kInRing + kWRing + kW2Ring) * RB);
    uint64_t *wbar = full + kInRing;  // split-phase "w / w^2 row complete" barriers, one arrival per warp

    const int tid = threadIdx.x;
    const int lane = tid & 31;

    int bx = blockIdx.x;
    const int r = bx % p.d;
    const int g = bx / p.d;
    const int frame = blockIdx.y;

    const int n_chain = (r < p.H) ? (p.H - r + p.d - 1) / p.d : 0;
    const int i0 = g * p.seg;
    const int n_out = min(p.seg, n_chain - i0);
    if (n_out <= 0) return;
    const int n_load = n_out + 4 * C;  // input rows i0-2C .. i0+n_out+2C-1 of the chain (virtual rows are reflected)
    const uint32_t row_bytes = (uint32_t)p.W * (uint32_t)sizeof(T);
    const T *src = reinterpret_cast<const T *>(p.in) + (long long)frame * p.in_bstride;
    const long long y_first = (long long)r + (long long)(i0 - 2 * C) * p.d;  // image row of input row 0

    const uint64_t pol_in = policy_evict_first();   // c_s is dead once this launch has read it
    const uint64_t pol_keep = policy_evict_last();  // c_{s+1}: the next scale reads it back
    constexpr bool hints = HINTS, hints_c = HINTS;
    int next_load = 0;  // thread 0: next chain row to request
    if (tid == 0) {
        for (int s = 0; s < kInRing; ++s) mbar_init(&full[s], 1);
        for (int s = 0; s < kW2Ring; ++s) mbar_init(&wbar[s], blockDim.x >> 5);
        fence_mbar_init();
        const int n0 = min(kInRing, n_load);
        for (; next_load < n0; ++next_load) {
            const long long y = reflect_any(y_first + (long long)next_load * p.d, p.H);
            mbar_arrive_expect_tx(&full[next_load], row_bytes);
            if (hints) tma_load_1d_hint(smem_raw + (size_t)next_load * RB, src + y * p.in_pitch, row_bytes, &full[next_load], pol_in);
            else tma_load_1d(smem_raw + (size_t)next_load * RB, src + y * p.in_pitch, row_bytes, &full[next_load]);
        }
    }
    __syncthreads();

    WhitenEpilogue<T> epi;
    epi.init(p, frame);

    uint32_t xb[NG];  // byte offset of this thread's own vector inside a staged row
    int xg[NG];
    bool act[NG];
    BytePlan<NV> plan[NG];
#pragma unroll
    for (int q = 0; q < NG; ++q) {
        xg[q] = (q * (int)blockDim.x + tid) * V;
        act[q] = xg[q] < p.W;
        if (!act[q]) xg[q] = 0;  // idle threads shadow vector 0; only their stores are masked
        xb[q] = (uint32_t)xg[q] * (uint32_t)sizeof(T);
        const TapPlan<NV> tp = make_tap_plan<V, NV>(xg[q], DMODE == 0 ? p.d : V, p.W, 0);
#pragma unroll
        for (int k = 0; k < NV; ++k) plan[q].off[k] = (uint32_t)tp.off[k] * (uint32_t)sizeof(T);
        plan[q].rev = tp.rev;
    }

    T SA[NG][V][TAPS - 1], SB[NG][V][TAPS - 1];  // running column sums of the smooth and of the power filter
#pragma unroll
    for (int q = 0; q < NG; ++q)
#pragma unroll
        for (int e = 0; e < V; ++e)
#pragma unroll
            for (int t = 0; t < TAPS - 1; ++t) { SA[q][e][t] = T(0); SB[q][e][t] = T(0); }

    T *c_ptr = reinterpret_cast<T *>(p.out_c) + (long long)frame * p.c_bstride +
               ((long long)r + (long long)(i0 - C) * p.d) * p.c_pitch;  // row of the first c produced (part C)
    T *o_ptr = reinterpret_cast<T *>(p.out_w) + (long long)frame * p.w_bstride +
               ((long long)r + (long long)i0 * p.d) * p.w_pitch;        // row of the first output
    const long long c_step = (long long)p.d * p.c_pitch, o_step = (long long)p.d * p.w_pitch;

    // Step j: A  input row j lands -> row pass -> column feed (c of the row C steps back)
    //         B  [for the w^2 row written during step j-1] row pass -> column feed -> P, epilogue, store w'
    //         C  c row m = j-2C: store, w = raw - c and w^2 -> shared rings, arrive on wbar
    // One extra drain step (j == n_load) runs part B only.  All ring slots are j-derived with power-of-two rings.
    for (int j = 0; j <= n_load; ++j) {
        Pack<T, V> cv[NG];
        if (j < n_load) {
            mbar_wait(&full[j & (kInRing - 1)], (uint32_t)(j / kInRing) & 1u);
            const uint32_t row = in_base + (uint32_t)(j & (kInRing - 1)) * RB;
#pragma unroll
            for (int q = 0; q < NG; ++q) {
                const Pack<T, V> v = row_pass_b<T, TAPS, DMODE, false>(row, plan[q]);
#pragma unroll
                for (int e = 0; e < V; ++e) cv[q].v[e] = col_feed<T, TAPS>(SA[q][e], v.v[e]);
            }
        }
        if (j > 2 * C) {
            const int mb = j - 1 - 2 * C;  // count index of the w row handled here (written during step j-1)
            mbar_wait(&wbar[mb & (kW2Ring - 1)], (uint32_t)(mb / kW2Ring) & 1u);
            if (tid == 0) {
                // every warp is past part C of step j-1: input rows <= j-1-C are free
                while (next_load < n_load && next_load - kInRing <= j - 1 - C) {
                    const int sl = next_load & (kInRing - 1);
                    const long long y = reflect_any(y_first + (long long)next_load * p.d, p.H);
                    mbar_arrive_expect_tx(&full[sl], row_bytes);
                    if (hints) tma_load_1d_hint(smem_raw + (size_t)sl * RB, src + y * p.in_pitch, row_bytes, &full[sl], pol_in);
                    else tma_load_1d(smem_raw + (size_t)sl * RB, src + y * p.in_pitch, row_bytes, &full[sl]);
                    ++next_load;
                }
            }
            const uint32_t row = w2_base + (uint32_t)(mb & (kW2Ring - 1)) * RB;
            Pack<T, V> pw[NG];
#pragma unroll
            for (int q = 0; q < NG; ++q) {
                const Pack<T, V> v = row_pass_b<T, TAPS, DMODE, false>(row, plan[q]);
#pragma unroll
                for (int e = 0; e < V; ++e) pw[q].v[e] = col_feed<T, TAPS>(SB[q][e], v.v[e]);
            }
            if (j > 4 * C) {
                // P row i = j-1-4C (relative to i0); its raw w row has count index i + C
                const uint32_t wrow = w_base + (uint32_t)((j - 1 - 3 * C) & (kWRing - 1)) * RB;
#pragma unroll
                for (int q = 0; q < NG; ++q) {
                    Pack<T, V> raw = lds_vec<T>(wrow + xb[q]);
#pragma unroll
                    for (int e = 0; e < V; ++e) raw.v[e] = epi.apply(raw.v[e], pw[q].v[e]);
                    if (act[q]) st_vec_cs(o_ptr + xg[q], raw);
                }
                o_ptr += o_step;
            }
        }
        if (j >= 2 * C && j < n_load) {
            const int mc = j - 2 * C;                       // count index of this c / w row
            const bool store_c = (mc >= C) && (mc - C < n_out);  // chain index relative to i0 is mc - C
            const uint32_t crow = in_base + (uint32_t)((j - C) & (kInRing - 1)) * RB;
            const uint32_t wrow = w_base + (uint32_t)(mc & (kWRing - 1)) * RB;
            const uint32_t w2row = w2_base + (uint32_t)(mc & (kW2Ring - 1)) * RB;
#pragma unroll
            for (int q = 0; q < NG; ++q) {
                if (store_c && act[q]) {
                    if (hints_c) st_vec_hint(c_ptr + xg[q], cv[q], pol_keep);
                    else st_vec(c_ptr + xg[q], cv[q]);
                }
                Pack<T, V> raw = lds_vec<T>(crow + xb[q]);
                Pack<T, V> sq;
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    raw.v[e] -= cv[q].v[e];
                    sq.v[e] = raw.v[e] * raw.v[e];
                }
                if (act[q]) {
                    sts_vec(wrow + xb[q], raw);
                    sts_vec(w2row + xb[q], sq);
                }
            }
            c_ptr += c_step;
            __syncwarp();
            if (lane == 0) mbar_arrive(&wbar[mc & (kW2Ring - 1)]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// fp32, d % 4 == 0 or d == 2: the same kernel on PACKED pixel pairs (FFMA2 / FMUL2 / FADD2).  The scalar kernel is
// bound by instruction issue, not by HBM (37 M warp instructions per 4096^2 plane, 57 % issue-active, DRAM 30 %); with
// every tap of a pixel pair being an aligned pair again, each 16-byte vector is carried as two 64-bit register pairs
// from LDS.128 to STG.128 and every filter FMA, the subtraction, the square and the epilogue multiplies issue once
// per two pixels.  Operation order and roundings are those of the scalar kernel: the planes are bit-identical.
// ---------------------------------------------------------------------------------------------------------------
template <int TAPS, int DMODE, bool HINTS>
__global__ void __launch_bounds__(512, 1) wow_rows_packed_kernel(const ScaleParams p) {
    using T = float;
    constexpr int V = 4, NG = 2;
    constexpr int C = TAPS / 2;
    constexpr int NV = PlanSize<TAPS, DMODE>::NV;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t RB = (uint32_t)p.row_stride * (uint32_t)sizeof(T);
    const uint32_t in_base = smem_u32(smem_raw);
    const uint32_t w_base = in_base + (uint32_t)kInRing * RB;
    const uint32_t w2_base = w_base + (uint32_t)kWRing * RB;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)(kInRing + kWRing + kW2Ring) * RB);
    uint64_t *wbar = full + kInRing;

    const int tid = threadIdx.x;
    const int lane = tid & 31;

    int bx = blockIdx.x;
    const int r = bx % p.d;
    const int g = bx / p.d;
    const int frame = blockIdx.y;

    const int n_chain = (r < p.H) ? (p.H - r + p.d - 1) / p.d : 0;
    const int i0 = g * p.seg;
    const int n_out = min(p.seg, n_chain - i0);
    if (n_out <= 0) return;
    const int n_load = n_out + 4 * C;
    const uint32_t row_bytes = (uint32_t)p.W * (uint32_t)sizeof(T);
    const T *src = reinterpret_cast<const T *>(p.in) + (long long)frame * p.in_bstride;
    const long long y_first = (long long)r + (long long)(i0 - 2 * C) * p.d;

    const uint64_t pol_in = policy_evict_first();
    const uint64_t pol_keep = policy_evict_last();
    int next_load = 0;
    if (tid == 0) {
        for (int s = 0; s < kInRing; ++s) mbar_init(&full[s], 1);
        for (int s = 0; s < kW2Ring; ++s) mbar_init(&wbar[s], blockDim.x >> 5);
        fence_mbar_init();
        const int n0 = min(kInRing, n_load);
        for (; next_load < n0; ++next_load) {
            const long long y = reflect_any(y_first + (long long)next_load * p.d, p.H);
            mbar_arrive_expect_tx(&full[next_load], row_bytes);
            if (HINTS) tma_load_1d_hint(smem_raw + (size_t)next_load * RB, src + y * p.in_pitch, row_bytes, &full[next_load], pol_in);
            else tma_load_1d(smem_raw + (size_t)next_load * RB, src + y * p.in_pitch, row_bytes, &full[next_load]);
        }
    }
    __syncthreads();

    WhitenEpilogue<T> epi;
    epi.init(p, frame);
    const PackedTaps<TAPS> H;

    uint32_t xb[NG];
    int xg[NG];
    bool act[NG];
    BytePlan<NV> plan[NG];
#pragma unroll
    for (int q = 0; q < NG; ++q) {
        xg[q] = (q * (int)blockDim.x + tid) * V;
        act[q] = xg[q] < p.W;
        if (!act[q]) xg[q] = 0;
        xb[q] = (uint32_t)xg[q] * (uint32_t)sizeof(T);
        const TapPlan<NV> tp = make_tap_plan<V, NV>(xg[q], DMODE == 0 ? p.d : V, p.W, 0);
#pragma unroll
        for (int k = 0; k < NV; ++k) plan[q].off[k] = (uint32_t)tp.off[k] * (uint32_t)sizeof(T);
        plan[q].rev = tp.rev;
    }

    u64 SA[NG][2][TAPS - 1], SB[NG][2][TAPS - 1];  // running column sums, one pixel pair per entry
#pragma unroll
    for (int q = 0; q < NG; ++q)
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int t = 0; t < TAPS - 1; ++t) { SA[q][e][t] = 0ull; SB[q][e][t] = 0ull; }

    T *c_ptr = reinterpret_cast<T *>(p.out_c) + (long long)frame * p.c_bstride +
               ((long long)r + (long long)(i0 - C) * p.d) * p.c_pitch;
    T *o_ptr = reinterpret_cast<T *>(p.out_w) + (long long)frame * p.w_bstride +
               ((long long)r + (long long)i0 * p.d) * p.w_pitch;
    const long long c_step = (long long)p.d * p.c_pitch, o_step = (long long)p.d * p.w_pitch;

    // same step structure as wow_rows_kernel: A (input row j), B (power of the w^2 row of step j-1), C (c / w / w^2)
    for (int j = 0; j <= n_load; ++j) {
        P4 cv[NG];
        if (j < n_load) {
            mbar_wait(&full[j & (kInRing - 1)], (uint32_t)(j / kInRing) & 1u);
            const uint32_t row = in_base + (uint32_t)(j & (kInRing - 1)) * RB;
#pragma unroll
            for (int q = 0; q < NG; ++q) {
                const P4 v = row_pass_p<TAPS, DMODE>(row, plan[q], H);
                cv[q].lo = col_feed_p<TAPS>(SA[q][0], v.lo, H);
                cv[q].hi = col_feed_p<TAPS>(SA[q][1], v.hi, H);
            }
        }
        if (j > 2 * C) {
            const int mb = j - 1 - 2 * C;
            mbar_wait(&wbar[mb & (kW2Ring - 1)], (uint32_t)(mb / kW2Ring) & 1u);
            if (tid == 0) {
                while (next_load < n_load && next_load - kInRing <= j - 1 - C) {
                    const int sl = next_load & (kInRing - 1);
                    const long long y = reflect_any(y_first + (long long)next_load * p.d, p.H);
                    mbar_arrive_expect_tx(&full[sl], row_bytes);
                    if (HINTS) tma_load_1d_hint(smem_raw + (size_t)sl * RB, src + y * p.in_pitch, row_bytes, &full[sl], pol_in);
                    else tma_load_1d(smem_raw + (size_t)sl * RB, src + y * p.in_pitch, row_bytes, &full[sl]);
                    ++next_load;
                }
            }
            const uint32_t row = w2_base + (uint32_t)(mb & (kW2Ring - 1)) * RB;
            P4 pw[NG];
#pragma unroll
            for (int q = 0; q < NG; ++q) {
                const P4 v = row_pass_p<TAPS, DMODE>(row, plan[q], H);
                pw[q].lo = col_feed_p<TAPS>(SB[q][0], v.lo, H);
                pw[q].hi = col_feed_p<TAPS>(SB[q][1], v.hi, H);
            }
            if (j > 4 * C) {
                const uint32_t wrow = w_base + (uint32_t)((j - 1 - 3 * C) & (kWRing - 1)) * RB;
#pragma unroll
                for (int q = 0; q < NG; ++q) {
                    P4 raw = lds_p4(wrow + xb[q]);
                    raw.lo = epi.apply2(raw.lo, pw[q].lo);
                    raw.hi = epi.apply2(raw.hi, pw[q].hi);
                    if (act[q]) stg_p4_cs(o_ptr + xg[q], raw);
                }
                o_ptr += o_step;
            }
        }
        if (j >= 2 * C && j < n_load) {
            const int mc = j - 2 * C;
            const bool store_c = (mc >= C) && (mc - C < n_out);
            const uint32_t crow = in_base + (uint32_t)((j - C) & (kInRing - 1)) * RB;
            const uint32_t wrow = w_base + (uint32_t)(mc & (kWRing - 1)) * RB;
            const uint32_t w2row = w2_base + (uint32_t)(mc & (kW2Ring - 1)) * RB;
#pragma unroll
            for (int q = 0; q < NG; ++q) {
                if (store_c && act[q]) {
                    if (HINTS) stg_p4_hint(c_ptr + xg[q], cv[q], pol_keep);
                    else stg_p4(c_ptr + xg[q], cv[q]);
                }
                P4 raw = lds_p4(crow + xb[q]);
                raw.lo = sub2(raw.lo, cv[q].lo);
                raw.hi = sub2(raw.hi, cv[q].hi);
                const P4 sq{mul2(raw.lo, raw.lo), mul2(raw.hi, raw.hi)};
                if (act[q]) {
                    sts_p4(wrow + xb[q], raw);
                    sts_p4(w2row + xb[q], sq);
                }
            }
            c_ptr += c_step;
            __syncwarp();
            if (lane == 0) mbar_arrive(&wbar[mc & (kW2Ring - 1)]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Host
// ---------------------------------------------------------------------------------------------------------------
struct WowGeom { int nt, ng; size_t smem; };

static bool plan_wow(ScaleParams &p, int taps, int esize, int batch, WowGeom *geo) {
    const int V = 16 / esize;
    const int c = taps / 2;
    if (!fast_path_ok(p, taps, esize)) return false;
    if (!p.out_c || !p.out_w) return false;
    const int vecs = p.W / V;
    const int ng = 2;
    const int nt = round_up((vecs + ng - 1) / ng, 32);
    if (nt > 512) return false;  // whole-row strips only: W <= 4096 (fp32) / 2048 (fp64)
    p.wt = p.W;
    p.n_strips = 1;
    p.halo_al = 0;
    p.row_stride = p.W;
    const size_t row = (size_t)p.W * esize;
    const int slots = kInRing;
    const size_t smem = (size_t)(kInRing + kWRing + kW2Ring) * row + 8 * (size_t)(kInRing + kW2Ring);
    if (smem > (size_t)kMaxSmem) return false;
    p.slots = slots;
    const int n_max = (p.H + p.d - 1) / p.d;
    const long long chains = (long long)(p.d < p.H ? p.d : p.H) * batch;
    int occ = (int)(kMaxSmem / (smem + 1024));
    const int occ_regs = 65536 / (nt * 128);
    if (occ > occ_regs) occ = occ_regs;
    if (occ < 1) occ = 1;
    const long long slots_total = (long long)device_sm_count() * occ;
    long long per_chain = slots_total / chains;
    if (per_chain < 1) per_chain = 1;
    int seg = (int)((n_max + per_chain - 1) / per_chain);
    if (seg < 12 * c) seg = 12 * c;
    if (seg > n_max) seg = n_max;
    p.seg = seg;
    p.n_seg = (n_max + seg - 1) / seg;
    geo->nt = nt;
    geo->ng = ng;
    geo->smem = smem;
    return true;
}

// WB_WOW_PACKED=0 in the environment selects the scalar fp32 kernel (A/B measurements, bit-identity tests).
static bool wow_packed_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("WB_WOW_PACKED");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

template <typename T, int TAPS, int DMODE, bool HINTS>
static auto wow_kernel_for(bool packed) -> void (*)(const ScaleParams) {
    if constexpr (sizeof(T) == 4 && (DMODE == 0 || DMODE == 2)) {
        if (packed) return wow_rows_packed_kernel<TAPS, DMODE, HINTS>;
    }
    return wow_rows_kernel<T, TAPS, DMODE, 2, HINTS>;
}

template <typename T, int TAPS, int DMODE, bool HINTS>
static int launch_wow_h(const ScaleParams &p, int batch, const WowGeom &geo, cudaStream_t st) {
    const bool packed = wow_packed_enabled();
    auto kern = wow_kernel_for<T, TAPS, DMODE, HINTS>(packed);
    static bool configured[2][64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[packed][dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < 64) configured[packed][dev] = true;
    }
    dim3 grid((unsigned)((long long)p.d * p.n_seg), (unsigned)batch);
    kern<<<grid, geo.nt, geo.smem, st>>>(p);
    return launch_status();
}

template <typename T, int TAPS, int DMODE>
static int launch_wow(const ScaleParams &p, int batch, const WowGeom &geo, cudaStream_t st) {
    return p.l2_hints ? launch_wow_h<T, TAPS, DMODE, true>(p, batch, geo, st)
                      : launch_wow_h<T, TAPS, DMODE, false>(p, batch, geo, st);
}

template <typename T, int TAPS>
static int dispatch_wow(const ScaleParams &p, int batch, const WowGeom &geo, cudaStream_t st) {
    constexpr int V = VecOf<T>::V;
    const int dmode = (p.d % V == 0) ? 0 : p.d;
    if (dmode == 0) return launch_wow<T, TAPS, 0>(p, batch, geo, st);
    if (dmode == 1) return launch_wow<T, TAPS, 1>(p, batch, geo, st);
    if constexpr (V == 4) {
        if (dmode == 2) return launch_wow<T, TAPS, 2>(p, batch, geo, st);
    }
    return WB_ENOT_FUSABLE;
}

static int fill_wow_params(ScaleParams &p, const void *in, void *out_c, void *out_w, int batch, int H, int W,
                           long long in_pitch, long long in_bstride, long long c_pitch, long long c_bstride,
                           long long w_pitch, long long w_bstride, int scale, int taps, int dtype) {
    int rc = check_common(batch, H, W, taps, dtype);
    if (rc) return rc;
    if (scale < 0 || scale > 30) return WB_EINVAL_SCALE;
    if (!in || !out_c || !out_w || in == out_c || in == out_w || out_c == out_w) return WB_EINVAL_POINTER;
    if (in_pitch < W || c_pitch < W || w_pitch < W) return WB_EINVAL_ARG;
    memset(&p, 0, sizeof(p));
    p.in = in; p.out_c = out_c; p.out_w = out_w;
    p.H = H; p.W = W; p.d = 1 << scale; p.Hg = H;
    p.in_pitch = in_pitch; p.in_bstride = in_bstride;
    p.c_pitch = c_pitch; p.c_bstride = c_bstride;
    p.w_pitch = w_pitch; p.w_bstride = w_bstride;
    {
        static int v = -1;  // WB_L2_HINTS_WOW=0/1/2/3 (bit 0: input evict-first, bit 1: c_{s+1} evict-last); A/B measurements
        if (v < 0) {
            const char *e = getenv("WB_L2_HINTS_WOW");
            v = e ? atoi(e) : 3;
        }
        p.l2_hints = l2_hints_enabled() ? v : 0;
    }
    return WB_OK;
}

}  // namespace wb

extern "C" {

int wb_wow_scale_path(int batch, int H, int W, long long in_pitch, long long out_c_pitch, long long out_w_pitch,
                      int scale, int taps, int dtype, const void *in, const void *out_c, const void *out_w) {
    wb::ScaleParams p;
    if (wb::fill_wow_params(p, in, const_cast<void *>(out_c), const_cast<void *>(out_w), batch, H, W, in_pitch, 0,
                            out_c_pitch, 0, out_w_pitch, 0, scale, taps, dtype))
        return 0;
    wb::WowGeom geo;
    return wb::plan_wow(p, taps, wb::dtype_size(dtype), batch, &geo) ? 1 : 0;
}

int wb_wow_scale(const void *in, void *out_c, void *out_w, int batch, int H, int W, long long in_pitch,
                 long long in_bstride, long long out_c_pitch, long long out_c_bstride, long long out_w_pitch,
                 long long out_w_bstride, int scale, int taps, int dtype, int sig_mode, double sigma, double sigma_e,
                 double noise_host, const double *noise_dev, double weight, void *stream) {
    wb::ScaleParams p;
    int rc = wb::fill_wow_params(p, in, out_c, out_w, batch, H, W, in_pitch, in_bstride, out_c_pitch, out_c_bstride,
                                 out_w_pitch, out_w_bstride, scale, taps, dtype);
    if (rc) return rc;
    if (sig_mode < 0 || sig_mode > 2) return WB_EINVAL_ARG;
    p.sig_mode = sig_mode; p.sigma = sigma; p.sigma_e = sigma_e;
    p.noise_host = noise_host; p.noise_dev = noise_dev; p.weight = weight;
    wb::WowGeom geo;
    if (!wb::plan_wow(p, taps, wb::dtype_size(dtype), batch, &geo)) return WB_ENOT_FUSABLE;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == WB_F32)
        return taps == 3 ? wb::dispatch_wow<float, 3>(p, batch, geo, st) : wb::dispatch_wow<float, 5>(p, batch, geo, st);
    return taps == 3 ? wb::dispatch_wow<double, 3>(p, batch, geo, st) : wb::dispatch_wow<double, 5>(p, batch, geo, st);
}

}  // extern "C"
