// The WOW pipeline of a stack of frames in ONE library call: the loop of watroo/utils.py:172-205 (whitening on) on the
// plain or the bilateral cascade, i.e. what wavelets_b200.utils._wow_stack issues scale by scale -- per scale wb_wow_scale
// (or, for bilateral scales, where the fused kernel declines, or when the MAD noise has to be estimated from the raw w_0
// first: wb_atrous_scale / wb_atrous_scale_bilateral [+ wb_abs_median] + wb_wow_whiten_scale), then wb_plane_moments +
// wb_residual_rescale of the residual plane and wb_synthesis.  Host code
// only: it exists because a 512^2 .. 2048^2 frame is finished on the device long before a Python loop has issued its
// fifteen launches (0.27 - 0.30 ms of interpreter and ctypes time per wow() call, measured; profiles/r2_sizes.json).
#include "common.cuh"

namespace wb {

static size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

}  // namespace wb

extern "C" {

size_t wb_wow_cascade_workspace_bytes(int dtype, int batch, long long n) {
    if (batch < 1) batch = 1;
    return wb::align256(wb_abs_median_workspace_bytes(dtype, batch, n)) + wb::align256(wb_plane_moments_workspace_bytes(batch)) +
           wb::align256((size_t)batch * 3 * sizeof(double));
}

int wb_wow_cascade(const void *in, long long in_pitch, long long in_bstride, void *planes, void *scratch, void *recon,
                   int batch, int H, int W, int n_scales, int taps, int dtype, const double *weights, const double *sigmas,
                   const double *sigma_e, const double *var_factors, int soft, double noise_host, double *noise_dev,
                   int estimate_noise, void *workspace, size_t workspace_bytes, void *stream) {
    int rc = wb::check_common(batch, H, W, taps, dtype);
    if (rc) return rc;
    if (n_scales < 1 || n_scales > 30) return WB_EINVAL_SCALE;
    if (!in || !planes || !scratch || !recon || !weights || !sigmas || !sigma_e || !workspace) return WB_EINVAL_POINTER;
    if (estimate_noise && !noise_dev) return WB_EINVAL_POINTER;
    const long long n = (long long)H * W;
    if (workspace_bytes < wb_wow_cascade_workspace_bytes(dtype, batch, n)) return WB_EINVAL_ARG;
    const size_t es = (size_t)wb::dtype_size(dtype);
    const int L = n_scales;
    char *pl = reinterpret_cast<char *>(planes);    // (batch, L + 1, H, W)
    char *sc = reinterpret_cast<char *>(scratch);   // (3, batch, H, W): two ping-pong smooth planes, one raw detail plane
    const long long plane_bs = (long long)(L + 1) * n;  // elements between the same plane of consecutive frames
    auto plane = [&](int s) { return pl + (size_t)s * (size_t)n * es; };
    auto scr = [&](int i) { return sc + (size_t)i * (size_t)batch * (size_t)n * es; };
    char *ws = reinterpret_cast<char *>(workspace);
    const size_t med_bytes = wb::align256(wb_abs_median_workspace_bytes(dtype, batch, n));
    void *med_ws = ws;
    void *mom_ws = ws + med_bytes;
    double *mom = reinterpret_cast<double *>(ws + med_bytes + wb::align256(wb_plane_moments_workspace_bytes(batch)));

    bool have_noise = !estimate_noise;  // a given noise: host scalar, or device scalars when noise_dev != NULL
    const double *nz_dev = estimate_noise ? nullptr : noise_dev;
    const void *src = in;
    long long src_pitch = in_pitch, src_bs = in_bstride;
    for (int s = 0; s < L; ++s) {
        void *dst_c = (s == L - 1) ? (void *)plane(L) : (void *)scr(s & 1);
        const long long c_bs = (s == L - 1) ? plane_bs : n;
        const bool need_sig = sigmas[s] != 0.0;
        const int mode = need_sig ? (soft ? 1 : 2) : 0;
        const double sg = need_sig ? sigmas[s] : 0.0, se = need_sig ? sigma_e[s] : 1.0;
        // bilateral scale (var_factors[s] = sigma_bilateral[s]^2 (s + 1 if bilateral_scaling else 1)): the range-weighted
        // gather K2 writes the raw detail plane, K3 whitens it
        const bool bilateral = var_factors != nullptr;
        bool two_pass = bilateral, median_of_raw = false;
        if (need_sig && !have_noise) {
            if (s == 0) {
                two_pass = median_of_raw = true;  // from the raw w_0: a grid-wide dependency inside the scale
            } else {
                // lazily, from the current state of plane 0 (already whitened here) -- watroo/wavelets.py:131-132
                rc = wb_abs_median(plane(0), n, batch, plane_bs, dtype, nullptr, noise_dev, sigma_e[0], med_ws, med_bytes, stream);
                if (rc) return rc;
                have_noise = true;
                nz_dev = noise_dev;
            }
        }
        if (!two_pass) {
            rc = wb_wow_scale(src, dst_c, plane(s), batch, H, W, src_pitch, src_bs, W, c_bs, W, plane_bs, s, taps, dtype, mode,
                              sg, se, need_sig ? noise_host : 0.0, need_sig ? nz_dev : nullptr, weights[s], stream);
            if (rc == WB_ENOT_FUSABLE) two_pass = true;
            else if (rc) return rc;
        }
        if (two_pass) {
            void *raw = scr(2);
            rc = bilateral ? wb_atrous_scale_bilateral(src, dst_c, raw, batch, H, W, src_pitch, src_bs, W, c_bs, W, n, s, taps,
                                                       dtype, var_factors[s], stream)
                           : wb_atrous_scale(src, dst_c, raw, batch, H, W, src_pitch, src_bs, W, c_bs, W, n, s, taps, dtype, stream);
            if (rc) return rc;
            if (median_of_raw) {
                rc = wb_abs_median(raw, n, batch, n, dtype, nullptr, noise_dev, sigma_e[0], med_ws, med_bytes, stream);
                if (rc) return rc;
                have_noise = true;
                nz_dev = noise_dev;
            }
            rc = wb_wow_whiten_scale(raw, plane(s), batch, H, W, W, n, W, plane_bs, s, taps, dtype, mode, sg, se,
                                     need_sig ? noise_host : 0.0, need_sig ? nz_dev : nullptr, weights[s], stream);
            if (rc) return rc;
        }
        src = dst_c;
        src_pitch = W;
        src_bs = c_bs;
    }
    // residual plane: c_L *= weight_L / std(c_L)  (watroo/utils.py:185-189, :203), then the sum of the planes (:205)
    rc = wb_plane_moments(plane(L), n, batch, plane_bs, dtype, mom, mom_ws, stream);
    if (rc) return rc;
    rc = wb_residual_rescale(plane(L), n, batch, plane_bs, dtype, mom, weights[L], stream);
    if (rc) return rc;
    return wb_synthesis(planes, L + 1, n, n, batch, plane_bs, recon, n, dtype, stream);
}

}  // extern "C"
