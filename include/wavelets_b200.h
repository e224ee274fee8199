/*
 * wavelets_b200.h -- C ABI of the B200 (sm_100a) à trous / WOW hot path.
 *
 * The reference (frederic-auchere/wavelets, "watroo" 0.0.4) is pure Python: it has no FFI of its own, its hot path
 * bottoms out in cv2.filter2D / numexpr / NumPy calls.  Each entry point below replaces one of those call sites;
 * the reference-side binding a maintainer would add is a ctypes stub.  The host layer that
 * mirrors the reference's Python API on top of this ABI is wavelets_b200/{wavelets,utils}.py.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer on the current CUDA device; the caller allocates every output and every
 *     workspace (the library never allocates or frees device memory);
 *   - images are row-major, innermost stride 1; `*_pitch` is the row stride and `*_bstride` the frame (batch)
 *     stride, both in ELEMENTS; `batch` frames are processed by one launch;
 *   - dtype: WB_F32 or WB_F64 (the arithmetic type of the path: the reference computes in the image dtype);
 *   - taps: WB_TRIANGLE (3 taps, [1/4,1/2,1/4]) or WB_B3SPLINE (5 taps, [1/16,1/4,3/8,1/4,1/16])
 *     (watroo/wavelets.py:239, :268);
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous and never synchronise the device;
 *   - return value: 0 on success, a negative WB_E* code for a rejected argument, or a positive cudaError_t;
 *     wb_error_string() renders both.  Nothing is written when a negative code is returned.
 *   - border rule everywhere: half-sample symmetric reflection (cv2.BORDER_REFLECT == np.pad 'symmetric'),
 *     any number of reflections.
 */
#ifndef WAVELETS_B200_H
#define WAVELETS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define WB_ABI_VERSION 1

enum { WB_F32 = 0, WB_F64 = 1 };
enum { WB_TRIANGLE = 3, WB_B3SPLINE = 5 };

enum {
    WB_OK = 0,
    WB_EINVAL_DTYPE = -1,
    WB_EINVAL_TAPS = -2,
    WB_EINVAL_SHAPE = -3,
    WB_EINVAL_SCALE = -4,
    WB_EINVAL_POINTER = -5,
    WB_EINVAL_ARG = -6
};

/* ABI version of the loaded library (== WB_ABI_VERSION of the header it was built from). */
int wb_abi_version(void);

/* Human-readable text for a status returned by any entry point. */
const char *wb_error_string(int status);

/* Which kernel variant wb_atrous_scale would launch for this problem: 1 = TMA row-pipeline kernel,
 * 0 = generic gather kernel (unaligned / tiny images).  For tests and the benchmark report. */
int wb_atrous_scale_path(int H, int W, long long in_pitch, long long out_pitch, int scale, int taps, int dtype,
                         const void *in, const void *out_c, const void *out_w);

/*
 * One scale of the plain cascade.  Replaces watroo/wavelets.py:35-45 (`convolution`, 2-D branch: cv2.filter2D with
 * the dilated dense kernel of :191-197 and BORDER_REFLECT) fused with the subtraction of :442:
 *     out_c(y,x) = sum_i sum_j h_i h_j in(R(y+(i-c)2^scale), R(x+(j-c)2^scale))          (c_{s+1})
 *     out_w      = in - out_c                                                            (w_s)
 * `in` is never written.  `out_w` may be NULL (smooth only); `out_c` may be NULL (detail plane only).
 * `out_c`/`out_w` must not alias `in` (rows of `in` are re-read as halo by neighbouring thread blocks).
 */
int wb_atrous_scale(const void *in, void *out_c, void *out_w, int batch, int H, int W,
                    long long in_pitch, long long in_bstride,
                    long long out_c_pitch, long long out_c_bstride,
                    long long out_w_pitch, long long out_w_bstride,
                    int scale, int taps, int dtype, void *stream);

/*
 * The whole plain cascade.  Replaces AtrousTransform.atrous_standard with bilateral=None
 * (watroo/wavelets.py:408-432,442-444).  `planes` is the (batch, levels+1, H, W) C-contiguous coefficient array
 * ([w_0 .. w_{L-1}, c_L] per frame); `scratch` is a caller-provided workspace of 2*batch*H*W elements that holds
 * the running smooth planes (c_s ping-pong).  `in` (pitch in_pitch, frame stride in_bstride) is not modified.
 * levels == 0 copies the image into plane 0.
 */
int wb_atrous_transform(const void *in, void *planes, void *scratch, int batch, int H, int W,
                        long long in_pitch, long long in_bstride, int levels, int taps, int dtype, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* WAVELETS_B200_H */
