// Whitening epilogue shared by K3 (wb_wow_whiten_scale) and the fused K1+K3 kernel (wb_wow_scale).
#pragma once

#include "pipeline.cuh"

namespace wb {

__device__ __forceinline__ float rsqrt_fast(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Whitening of one coefficient (watroo/utils.py:195-203 + watroo/wavelets.py:129-143), P = S_s[w^2] at that pixel:
//     w' = w * significance(w) * (weight / sqrt(P > 0 ? P : 1e-15))
// fp32: weight / sqrt(P) is evaluated as weight * rsqrt.approx(P) (MUFU.RSQ, <= 2 ulp) instead of the reference's
// sqrt -> divide -> multiply chain (three roundings): a few 1e-7 relative, far inside the 1e-5 parity budget, and
// 5 instructions per pixel instead of ~30 (the fused kernel is issue-bound, not HBM-bound, otherwise).
// fp64: weight * rsqrt(P) (CUDA's double-precision rsqrt, <= 1 ulp; the sqrt -> divide chain cost ~40 FP64 instructions per
// pixel and made the fp64 whitening pass FP64-pipe-bound at twice its HBM time); the difference to the reference's
// two roundings is <= 3 ulp, four orders of magnitude inside the 1e-12 budget.  Soft threshold: erf(|w / thr|) with the reciprocal of thr hoisted (fp32);
// hard threshold: NumPy >= 2 compares |w| (fp32) with the float64 threshold; |w| > thr in float64 is equivalent to
// |w| > RD(thr) in fp32 (RD = round towards -inf), so the mask stays bit-exact without float64 instructions.
template <typename T> struct WhitenEpilogue {
    int mode;
    T thr_cmp;  // hard threshold in the plane dtype, rounded down
    T thr;      // fp64: the threshold itself (soft); fp32: unused
    T inv_thr;  // soft: 1 / thr
    T weight;
    __device__ __forceinline__ void init(const ScaleParams &p, int frame) {
        mode = p.sig_mode;
        double t = 0.0;
        if (mode) {
            const double noise = p.noise_dev ? p.noise_dev[frame] : p.noise_host;
            if (noise == 0.0) mode = 0;  // scalar noise == 0 -> significance is all ones (wavelets.py:134-135)
            t = (p.sigma * noise) * p.sigma_e;  // (sigma * noise) * sigma_e, float64 as in the reference
        }
        if constexpr (sizeof(T) == 4) {
            thr_cmp = __double2float_rd(t);
            thr = (T)t;
            // a denormal threshold must not turn 1 / thr into inf (0 * inf would poison pixels where w == 0, for which the
            // reference's erf(|w / thr|) is 0): clamp to the largest finite float
            inv_thr = (T)fmin(1.0 / t, 3.4028234663852886e38);
        } else {
            thr_cmp = t;
            thr = t;
            inv_thr = fmin(1.0 / t, 1.7976931348623157e308);  // hoisted reciprocal (one rounding more than w / thr)
        }
        weight = (T)p.weight;
    }
    __device__ __forceinline__ T apply(T w, T power) const {
        power = (power <= T(0)) ? T(1e-15) : power;
        T g;
        if constexpr (sizeof(T) == 4) g = weight * rsqrt_fast(power);
        else g = weight * rsqrt(power);  // MUFU.RSQ64H seed + Newton steps, <= 1 ulp: ~10 FP64 instructions instead of the ~40 of sqrt + divide
        if (mode == 1) {
            // the reference multiplies by erf() evaluated in float64 and rounds the product to the plane dtype
            if constexpr (sizeof(T) == 4) w = w * erff(fabsf(w * inv_thr));
            else w = w * erf(fabs(w * inv_thr));
        } else if (mode == 2) {
            w = (fabs(w) > thr_cmp) ? w : T(0);
        }
        return w * g;
    }
    // Two pixels at once (fp32 only): the same operations as apply() with the multiplies packed.  MODE is the
    // significance mode the kernel was compiled for; `mode` can still be 0 at run time (noise == 0).
    template <int MODE>
    __device__ __forceinline__ u64 apply2(u64 w2, u64 power2) const {
        float p0, p1, w0, w1;
        up2(power2, p0, p1);
        p0 = (p0 <= 0.f) ? 1e-15f : p0;
        p1 = (p1 <= 0.f) ? 1e-15f : p1;
        const u64 g2 = mul2(pk2((float)weight, (float)weight), pk2(rsqrt_fast(p0), rsqrt_fast(p1)));
        if (MODE == 1 && mode == 1) {
            up2(w2, w0, w1);
            w2 = mul2(w2, pk2(erff(fabsf(w0 * (float)inv_thr)), erff(fabsf(w1 * (float)inv_thr))));
        } else if (MODE == 2 && mode == 2) {
            up2(w2, w0, w1);
            w2 = pk2((fabsf(w0) > (float)thr_cmp) ? w0 : 0.f, (fabsf(w1) > (float)thr_cmp) ? w1 : 0.f);
        }
        return mul2(w2, g2);
    }
    // One 64-bit lane of a lean kernel: two packed fp32 pixels (apply2) or one double (apply with the compiled-in mode).
    template <int MODE>
    __device__ __forceinline__ u64 apply_lane(u64 w, u64 power) const {
        if constexpr (sizeof(T) == 4) {
            return apply2<MODE>(w, power);
        } else {
            double x = Lane<double>::d(w), pw = Lane<double>::d(power);
            pw = (pw <= 0.0) ? 1e-15 : pw;
            const double g = weight * rsqrt(pw);
            if (MODE == 1 && mode == 1) x = x * erf(fabs(x * inv_thr));
            else if (MODE == 2 && mode == 2) x = (fabs(x) > thr_cmp) ? x : 0.0;
            return Lane<double>::u(x * g);
        }
    }
};

}  // namespace wb
