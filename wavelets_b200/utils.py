"""Applications on top of the transform: ``denoise`` and ``wow`` (mirror of watroo/utils.py:83-102, :105-219).

Both keep the reference's signatures and quirks (``denoise``'s ``weights`` are the sigma thresholds; ``wow`` returns
``(recon, coefficients)`` with the planes overwritten by their whitened values).  NumPy in -> NumPy out for the image
results; torch tensor in -> torch CUDA tensor out.  ``coefficients.data`` always stays on the device.
"""
from __future__ import annotations

import copy
import os
import warnings

import numpy as np
import torch

from . import _lib
from .scaling import B3spline
from .wavelets import (AtrousTransform, Coefficients, _Noise, _frame_layout, abs_median_noise, atrous_scale,
                       bilateral_list, plane_moments, synthesis, to_device_image)

__all__ = ["denoise", "wow", "wow_batch", "wow_stream", "generalized_anscombe", "enhance", "richardson_lucy"]

# Development switch (tests compare the fused one-pass WOW scales against the two-pass route bit for bit).
FUSED_WOW = True
# The cascade with whitening in ONE library call (wb_wow_cascade) instead of a Python loop over the scales: the same
# launches, without the interpreter / ctypes time per launch (tests switch it off to compare the two bit for bit).
C_CASCADE = os.environ.get("WB_C_CASCADE", "1") != "0"
# Bilateral WOW, Python loop only: run the memory-bound half of every scale (exact median, whitening K3) on a
# high-priority side stream while the compute-bound bilateral kernel K2 of the next scale runs on the caller's stream.
# Measured: 1 % on 4096^2 fp32 frames (K2 is issue-bound: a co-resident kernel is paid in full) against 0.4 ms of host
# time per frame on small ones -- off by default (WB_OVERLAP_BILATERAL=1 turns it on and bypasses wb_wow_cascade).
OVERLAP_BILATERAL = os.environ.get("WB_OVERLAP_BILATERAL", "0") != "0"
_SIDE_STREAMS: dict = {}


def _side_stream(device):
    """One high-priority stream per device for the whitening passes that overlap with K2 (created once)."""
    key = (device.type, device.index)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device, priority=-1)
    return st


_PIPELINE_STREAMS: dict = {}


def _pipeline_streams(device):
    """The (upload, compute, download) streams of the host-frame pipelines, created once per device: the caching
    allocator keeps one pool per stream, so fresh streams per call would start every call with cudaMalloc."""
    key = (device.type, device.index)
    st = _PIPELINE_STREAMS.get(key)
    if st is None:
        st = _PIPELINE_STREAMS[key] = tuple(torch.cuda.Stream(device) for _ in range(3))
    return st


def generalized_anscombe(signal, alpha=1, g=0, sigma=0, inverse=False):
    """Variance-stabilising transform and its algebraic inverse (watroo/wavelets.py:14-21), element-wise on a torch
    tensor (device or host)."""
    if inverse:
        return ((alpha * signal / 2) ** 2 + alpha * g - sigma ** 2 - 3 * alpha / 8) / alpha
    dum = alpha * signal + 3 * alpha ** 2 / 8 + sigma ** 2 - alpha * g
    return 2 * torch.sqrt(torch.clamp(dum, min=0)) / alpha


class _NullContext:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


def _result(t, as_numpy):
    return t.cpu().numpy() if as_numpy else t


def denoise(data, weights, scaling_function=B3spline, noise=None, bilateral=None, soft_threshold=True,
            anscombe=False):
    """Convenience denoiser (watroo/utils.py:83-102).  NB ``weights`` are the sigma thresholds of each scale and
    their count sets the number of scales; the result is the sum of the thresholded planes."""
    img, was_numpy = to_device_image(data, ndim_ok=(1, 2, 3))  # 1-D signals and 3-D volumes: plain cascade only
    if anscombe:
        img = generalized_anscombe(img)
    transform = AtrousTransform(scaling_function, bilateral=bilateral)
    coefficients = transform(img, len(weights))
    coefficients.noise = noise
    coefficients.denoise(weights, soft_threshold=soft_threshold)
    planes = coefficients.data
    out = synthesis(planes if img.ndim == 2 else planes.reshape(planes.shape[0], 1, -1)).reshape(img.shape)
    if anscombe:
        out = generalized_anscombe(out, inverse=True)
    return _result(out, was_numpy)


def _wow_plan(shape, scaling_function, n_scales, weights, denoise_coefficients, bilateral, from_coefficients=None,
              h=0, n_dims=2):
    """Scale-count and per-scale parameter lists of wow() (watroo/utils.py:121-146, :160-170)."""
    n_taps = len(scaling_function.coefficients_1d)
    if from_coefficients is None:
        max_scales = int(np.round(np.log2(min(shape)) - np.log2(n_taps)))
        if n_scales is None:
            n_scales = max_scales if h < 1 else len(denoise_coefficients)  # utils.py:123-124
        elif n_scales > max_scales:
            n_scales = max_scales
    else:
        n_scales = len(from_coefficients) - 1
    table_len = len(scaling_function(n_dims).sigma_e(bilateral=bilateral))
    if len(denoise_coefficients) >= table_len:
        warnings.warn(f"Required number of scales lager then the maximum for scaling function. Using {table_len}.")
        n_scales = table_len
    sigma_bilateral = None if bilateral is None else bilateral_list(bilateral, n_scales)
    wts = copy.copy(list(weights))
    if len(wts) <= n_scales:
        wts.extend([1, ] * (n_scales - len(wts) + 1))
    dns = copy.copy(list(denoise_coefficients))
    if len(dns) < n_scales:
        dns.extend([0, ] * (n_scales - len(dns)))
    if len(dns) == n_scales:
        dns.extend([1, ])
    return n_scales, sigma_bilateral, wts, dns


def _whiten_scale(lib, w_raw, out, s, scaling_function, sig_mode, sigma, sigma_e, nz, weight):
    b, h, w, pitch, bstride = _frame_layout(w_raw)
    _, _, _, o_pitch, o_bstride = _frame_layout(out)
    with torch.cuda.device(w_raw.device):
        _lib.check(lib.wb_wow_whiten_scale(w_raw.data_ptr(), out.data_ptr(), b, h, w, pitch, bstride, o_pitch,
                                           o_bstride, s, scaling_function.taps_code, _lib.dtype_code(w_raw.dtype),
                                           sig_mode, float(sigma), float(sigma_e), nz.host, nz.dev_ptr, float(weight),
                                           _lib.stream_ptr(w_raw.device)))


def _wow_scale_fused(lib, src, dst_c, out_w, s, scaling_function, sig_mode, sigma, sigma_e, nz, weight):
    """One fused WOW scale (wb_wow_scale).  Returns False without launching anything when the problem is outside the
    fused kernel (the caller then takes the two-pass route)."""
    b, h, w, pitch, bstride = _frame_layout(src)
    _, _, _, c_pitch, c_bstride = _frame_layout(dst_c)
    _, _, _, w_pitch, w_bstride = _frame_layout(out_w)
    with torch.cuda.device(src.device):
        rc = lib.wb_wow_scale(src.data_ptr(), dst_c.data_ptr(), out_w.data_ptr(), b, h, w, pitch, bstride, c_pitch,
                              c_bstride, w_pitch, w_bstride, s, scaling_function.taps_code, _lib.dtype_code(src.dtype),
                              sig_mode, float(sigma), float(sigma_e), nz.host, nz.dev_ptr, float(weight),
                              _lib.stream_ptr(src.device))
    if rc == _lib.WB_ENOT_FUSABLE:
        return False
    _lib.check(rc)
    return True


def _scalar_noise(noise, device):
    """None -> None; number / 0-d array / 1-element tensor -> _Noise with a host or device scalar."""
    if noise is None:
        return None
    if isinstance(noise, torch.Tensor):
        if noise.numel() == 1 and noise.ndim <= 1 and noise.is_cuda and noise.dtype == torch.float64:
            return _Noise(dev=noise)
        if noise.numel() == 1:
            return _Noise(host=float(noise.item()))
        return "map"
    if isinstance(noise, np.ndarray) and noise.ndim > 0:
        return "map"
    return _Noise(host=float(noise))


def _wow_stack(stack, scaling_function_class, n_scales, wts, dns, sigma_bilateral, bilateral_scaling, whitening,
               soft_threshold, noise, buffers=None):
    """The fused WOW pipeline on a stack (B, H, W) of frames.  Returns (recon (B,H,W), planes (B,L+1,H,W), noise).

    ``buffers``: optional dict of preallocated device tensors ("planes", "scratch", "raw_planes", "recon"), filled in
    on first use and reused by later calls with the same geometry (``wow_stream`` keeps one per frame in flight, so
    that its steady state allocates nothing large).

    Per plain scale ONE fused launch (wb_wow_scale: smooth, detail, local power, significance, whitening; the raw
    detail plane never exists in HBM).  Bilateral scales, and scale 0 when the MAD noise must first be estimated from
    the raw w_0, take two passes: K1/K2 writes c_{s+1} into a ping-pong scratch and the raw w_s into a one-plane
    scratch, K3 whitens it into the output.  The noise is estimated on the device (no host synchronisation anywhere),
    the residual plane is rescaled in place and K5 sums the planes."""
    lib = _lib.load(require_cuda=True)
    sf = scaling_function_class(2)
    b, h, w = stack.shape
    dev, dt = stack.device, stack.dtype
    L = n_scales
    if buffers is None:
        buffers = {}

    def buffer(name, shape):
        t = buffers.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dt or t.device != dev:
            t = buffers[name] = torch.empty(shape, dtype=dt, device=dev)
        return t

    planes = buffer("planes", (b, L + 1, h, w))
    bilateral = sigma_bilateral
    sigma_e = sf.sigma_e(bilateral=bilateral)
    transform = AtrousTransform(scaling_function_class, bilateral=bilateral, bilateral_scaling=bilateral_scaling)
    factors = transform.var_factors(L) if bilateral is not None else [None] * L
    if L == 0:
        planes[:, 0].copy_(stack)
    # (bilateral cascades: in one call unless the side-stream overlap of K3 with K2 is asked for -- it gains 1 % on 4096^2
    # frames and costs 0.4 ms of host time on small ones)
    if C_CASCADE and FUSED_WOW and whitening and (bilateral is None or not OVERLAP_BILATERAL) and L >= 1 and \
            dev.type == "cuda":
        # one call: wb_wow_cascade issues exactly the launches of the loop below
        import ctypes
        scratch3 = buffer("scratch3", (3, b, h, w))
        recon = buffer("recon", (b, h, w))
        code = _lib.dtype_code(dt)
        ws_bytes = lib.wb_wow_cascade_workspace_bytes(code, b, h * w)
        ws = buffers.get("cascade_ws")
        if ws is None or ws.numel() < ws_bytes or ws.device != dev:
            ws = buffers["cascade_ws"] = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        any_sig = any(d != 0 for d in dns[:L])
        estimate = noise is None and any_sig
        nz = noise
        if estimate:
            nz = _Noise(dev=torch.empty(b, dtype=torch.float64, device=dev))
        use = nz if nz is not None else _Noise()
        arr = ctypes.c_double * (L + 1)
        _, _, _, pitch, bstride = _frame_layout(stack)
        with torch.cuda.device(dev):
            _lib.check(lib.wb_wow_cascade(stack.data_ptr(), pitch, bstride, planes.data_ptr(), scratch3.data_ptr(),
                                          recon.data_ptr(), b, h, w, L, sf.taps_code, code,
                                          arr(*[float(x) for x in wts[:L + 1]]),
                                          arr(*([float(x) for x in dns[:L]] + [0.0])),
                                          arr(*([float(x) for x in sigma_e[:L]] + [0.0])),
                                          None if bilateral is None else arr(*([float(x) for x in factors[:L]] + [0.0])),
                                          1 if soft_threshold else 0, float(use.host), use.dev_ptr, 1 if estimate else 0,
                                          ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)))
        return recon, planes, nz
    scratch = buffer("scratch", (2, b, h, w)) if L > 1 else None
    raw_planes = None  # scratch for the raw w_s of the two-pass route, allocated on first use
    nz = noise  # _Noise or None
    src = stack
    # Bilateral cascade with whitening: K2 is bound by the MUFU/FMA pipes and leaves HBM idle, the exact median and the
    # whitening pass K3 are bound by HBM.  They go to a high-priority side stream: K3 of scale s then runs under K2 of
    # scale s+1 (its blocks are scheduled as K2 blocks retire).  Three raw planes rotate (with two, K2 of scale 2 had to
    # wait for the whitening of scale 0, which itself waits for the exact median); events order the hand-over.
    overlap = OVERLAP_BILATERAL and whitening and bilateral is not None and L > 1 and dev.type == "cuda"
    main = torch.cuda.current_stream(dev) if overlap else None
    side = _side_stream(dev) if overlap else None
    ev_k3 = [None, None, None]  # whitening of the scale that last used raw_planes[i]
    for s in range(L):
        dst_c = planes[:, L] if s == L - 1 else scratch[s & 1]
        d, wt = dns[s], wts[s]
        need_sig = d != 0
        mode = (1 if soft_threshold else 2) if need_sig else 0
        if whitening:
            if need_sig and nz is None and s > 0:
                # lazily, from the current state of plane 0 (already whitened here) -- watroo/wavelets.py:131-132
                if overlap:
                    main.wait_stream(side)
                nz = _Noise(dev=abs_median_noise(planes[:, 0], sigma_e[0]))
            # One pass (K1+K3 fused, raw w_s stays on chip) unless the scale is bilateral, the MAD noise has to be
            # estimated from the raw w_0 first (grid-wide dependency), or the shape is outside the fused kernel.
            fused = FUSED_WOW and factors[s] is None and not (need_sig and nz is None) and \
                _wow_scale_fused(lib, src, dst_c, planes[:, s], s, sf, mode, d, sigma_e[s] if need_sig else 1.0,
                                 nz if need_sig else _Noise(), wt)
            if not fused:
                if raw_planes is None:
                    raw_planes = buffer("raw_planes", (3 if overlap else 1, b, h, w))
                raw_plane = raw_planes[s % 3] if overlap else raw_planes[0]
                if overlap and ev_k3[s % 3] is not None:
                    main.wait_event(ev_k3[s % 3])  # the whitening that read this raw plane three scales ago is done
                atrous_scale(src, s, sf, out_c=dst_c, out_w=raw_plane, var_factor=factors[s])
                if overlap:
                    side.wait_stream(main)
                with torch.cuda.stream(side) if overlap else _NullContext():
                    if need_sig and nz is None:
                        nz = _Noise(dev=abs_median_noise(raw_plane, sigma_e[0]))  # s == 0: from the raw w_0
                    _whiten_scale(lib, raw_plane, planes[:, s], s, sf, mode, d, sigma_e[s] if need_sig else 1.0,
                                  nz if need_sig else _Noise(), wt)
                    if overlap:
                        ev_k3[s % 3] = torch.cuda.Event()
                        ev_k3[s % 3].record(side)
        else:
            atrous_scale(src, s, sf, out_c=dst_c, out_w=planes[:, s], var_factor=factors[s])
            if need_sig and nz is None:
                nz = _Noise(dev=abs_median_noise(planes[:, 0], sigma_e[0]))
            if mode or wt != 1:
                use = nz if need_sig else _Noise()
                with torch.cuda.device(dev):
                    _lib.check(lib.wb_denoise_plane(planes[:, s].data_ptr(), h * w, b, (L + 1) * h * w,
                                                    _lib.dtype_code(dt), mode, float(d),
                                                    float(sigma_e[s]) if need_sig else 1.0, use.host, use.dev_ptr, 0,
                                                    float(wt), _lib.stream_ptr(dev)))
        src = dst_c
    if overlap:
        main.wait_stream(side)  # every whitened plane is complete before the tail (and before any buffer is freed)
    # residual plane: c_L *= wt_L / std(c_L)  (watroo/utils.py:185-189, :203)
    last = planes[:, L]
    if whitening:
        # (a fused rescale + synthesis pass measured slower: the separate rescale finds c_L still in L2)
        mom = plane_moments(last)
        with torch.cuda.device(dev):
            _lib.check(lib.wb_residual_rescale(last.data_ptr(), h * w, b, (L + 1) * h * w, _lib.dtype_code(dt),
                                               mom.data_ptr(), float(wts[L]), _lib.stream_ptr(dev)))
    elif wts[L] != 1:
        last.mul_(wts[L])
    recon = synthesis(planes, out=buffer("recon", (b, h, w)))
    return recon, planes, nz


def wow(data, scaling_function=B3spline, n_scales=None, weights=[], whitening=True, denoise_coefficients=[],
        noise=None, bilateral=None, bilateral_scaling=False, soft_threshold=True, preserve_variance=False, gamma=3.2,
        gamma_min=None, gamma_max=None, h=0):
    """Wavelets Optimized Whitening (watroo/utils.py:105-219): ``(recon, coefficients)``.

    ``data`` is a 2-D image (ndarray or torch tensor) or a ``Coefficients`` object (which is then whitened in
    place, as in the reference).  The default call and the ``bilateral`` / ``denoise_coefficients`` / ``weights`` /
    ``whitening`` options run on the fused kernels; ``h > 0`` (gamma blending, utils.py:157-158, :207-217) and
    ``preserve_variance`` (utils.py:178-184) take a plane-by-plane route over the same kernels."""
    general = h != 0 or preserve_variance
    if isinstance(data, Coefficients):
        if general:
            return _wow_general(data, weights, whitening, denoise_coefficients, bilateral, soft_threshold,
                                preserve_variance, gamma, gamma_min, gamma_max, h)
        return _wow_coefficients(data, weights, whitening, denoise_coefficients, bilateral, soft_threshold)
    if not isinstance(data, (np.ndarray, torch.Tensor)):
        raise ValueError("Unknown input type")  # watroo/utils.py:133
    img, was_numpy = to_device_image(data, ndim_ok=(1, 2, 3))
    if img.ndim != 2:
        # 1-D signals and 3-D volumes: the reference's loop is dimension-generic (watroo/utils.py:148-150, :194)
        n_scales, sigma_bilateral, wts, dns = _wow_plan(img.shape, scaling_function, n_scales, weights,
                                                        denoise_coefficients, bilateral, h=h, n_dims=img.ndim)
        transform = AtrousTransform(scaling_function, bilateral=sigma_bilateral, bilateral_scaling=bilateral_scaling)
        co = transform(img, n_scales)
        co.noise = noise
        recon = _wow_nd(co, wts, dns, whitening, soft_threshold, preserve_variance, gamma, gamma_min, gamma_max, h)
        return _result(recon, was_numpy), co
    n_scales, sigma_bilateral, wts, dns = _wow_plan(img.shape, scaling_function, n_scales, weights,
                                                    denoise_coefficients, bilateral, h=h)
    if general:
        transform = AtrousTransform(scaling_function, bilateral=sigma_bilateral, bilateral_scaling=bilateral_scaling)
        co = transform(img, n_scales)
        co.noise = noise
        recon, co = _wow_general(co, weights, whitening, denoise_coefficients, bilateral, soft_threshold,
                                 preserve_variance, gamma, gamma_min, gamma_max, h, plan=(n_scales, wts, dns))
        return _result(recon, was_numpy), co
    nz = _scalar_noise(noise, img.device)
    if nz == "map":
        # per-pixel noise maps take the unfused route: transform, then whiten plane by plane
        transform = AtrousTransform(scaling_function, bilateral=sigma_bilateral, bilateral_scaling=bilateral_scaling)
        co = transform(img, n_scales)
        co.noise = noise
        recon, co = _wow_coefficients(co, weights, whitening, denoise_coefficients, bilateral, soft_threshold,
                                      plan=(n_scales, sigma_bilateral, wts, dns))
        return _result(recon, was_numpy), co
    recon, planes, nz = _wow_stack(img.unsqueeze(0), scaling_function, n_scales, wts, dns, sigma_bilateral,
                                   bilateral_scaling, whitening, soft_threshold, nz)
    co = Coefficients(planes[0], scaling_function(2), sigma_bilateral)
    if nz is not None:
        co.noise = nz.dev if nz.dev is not None else nz.host
    else:
        co.noise = None
    return _result(recon[0], was_numpy), co


def _wow_nd(co, wts, dns, whitening, soft_threshold, preserve_variance, gamma, gamma_min, gamma_max, h):
    """The loop of wow() (watroo/utils.py:172-217) on the coefficients of a 1-D signal or a 3-D volume, every option
    included: the transform and the n-D smooth of the local power run in the library's kernels, the plane-wise factors
    are element-wise device arithmetic in the reference's order and dtypes.  A parity path, whitened in place."""
    from .wavelets import _smooth_nd
    data = co.data
    dt = data.dtype
    L = len(co) - 1
    white = whitening and h < 1
    gamma_scaled = torch.zeros_like(data[0]) if h > 0 else None
    for s in range(L + 1):
        c = data[s]
        power = c * c
        power_norm = None
        if preserve_variance:  # utils.py:178-184
            power_norm = c.to(torch.float64).std(unbiased=False).to(dt) if s == L else torch.sqrt(power.mean())
        if s == L:
            local_power = None
            if white:
                sd = plane_moments(c.reshape(1, -1))[0, 2].to(dt)
                local_power = torch.where(sd <= 0, torch.full_like(sd, 1e-15), sd)
        else:
            local_power = None
            if white:
                local_power = _smooth_nd(power.contiguous(), s, co.scaling_function)  # plain smooth, utils.py:194
                local_power = torch.sqrt(torch.where(local_power <= 0, torch.full_like(local_power, 1e-15), local_power))
            sig = co.significance(dns[s], s, soft_threshold=soft_threshold)
            if sig.dtype == torch.bool or sig.dtype == torch.float64:
                c.copy_((c.to(torch.float64) * sig.to(torch.float64)).to(dt))  # product in float64, rounded once
        if gamma_scaled is not None:
            gamma_scaled += c
        factor = torch.as_tensor(wts[s], dtype=dt, device=c.device)
        if power_norm is not None:
            factor = factor * power_norm
        if local_power is not None:
            factor = factor / local_power
        c.mul_(factor)
    recon = synthesis(data.reshape(L + 1, 1, -1)).reshape(data.shape[1:])
    if gamma_scaled is not None:  # utils.py:207-217
        lo = gamma_scaled.min() if gamma_min is None else gamma_min
        hi = gamma_scaled.max() if gamma_max is None else gamma_max
        gamma_scaled -= lo
        gamma_scaled /= (hi - lo)
        gamma_scaled.clamp_(0, 1)
        gamma_scaled.pow_(1 / gamma)
        recon = (1 - h) * recon + h * gamma_scaled
    return recon


def wow_batch(frames, scaling_function=B3spline, n_scales=None, weights=[], whitening=True, denoise_coefficients=[],
              noise=None, bilateral=None, bilateral_scaling=False, soft_threshold=True):
    """NEW entry point (no reference equivalent): WOW of a stack ``(B, H, W)`` of independent frames, every kernel
    launched once for the whole stack.  Returns ``(recon (B,H,W), planes (B,L+1,H,W), noise (B,) or None)`` as device
    tensors; per-frame results are identical to ``wow(frame, ...)``."""
    stack, _ = to_device_image(frames, ndim_ok=(3,))
    n_scales, sigma_bilateral, wts, dns = _wow_plan(stack.shape[1:], scaling_function, n_scales, weights,
                                                    denoise_coefficients, bilateral)
    nz = _scalar_noise(noise, stack.device)
    if nz == "map":
        raise NotImplementedError("wow_batch() takes a scalar noise or None")
    recon, planes, nz = _wow_stack(stack, scaling_function, n_scales, wts, dns, sigma_bilateral, bilateral_scaling,
                                   whitening, soft_threshold, nz)
    noise_out = None if nz is None else (nz.dev if nz.dev is not None else nz.host)
    return recon, planes, noise_out


def wow_stream(frames, out=None, depth=2, scaling_function=B3spline, **kwargs):
    """NEW entry point (no reference equivalent): ``wow()`` of a sequence of HOST frames ``(N, H, W)`` into a host array
    ``(N, H, W)`` of reconstructions, overlapping the host->device copy of frame n+1, the WOW of frame n and the
    device->host copy of the reconstruction of frame n-1 on three CUDA streams (``depth`` device buffers in flight).

    ``frames`` / ``out``: torch CPU tensors (pinned memory is used as is, pageable memory is staged through a pinned copy)
    or NumPy arrays; ``kwargs`` are those of ``wow`` (scalar ``noise`` only).  Returns ``out`` (allocated pinned when None;
    a NumPy view of it for NumPy input).  Per-frame results are identical to ``wow(frame, ...)[0]``."""
    was_numpy = not isinstance(frames, torch.Tensor)
    host = torch.from_numpy(np.ascontiguousarray(frames)) if was_numpy else frames
    if host.ndim != 3:
        raise ValueError("wow_stream() takes a stack of frames (N, H, W)")
    if host.is_cuda:
        raise ValueError("wow_stream() takes host frames; use wow_batch() for device-resident stacks")
    if host.dtype not in (torch.float32, torch.float64):
        host = host.to(torch.float64)  # the reference's recast rule for integer inputs
    if not host.is_pinned():
        host = host.contiguous().pin_memory()
    n, h, w = host.shape
    if out is None:
        out = torch.empty((n, h, w), dtype=host.dtype).pin_memory()
    out_t = torch.from_numpy(out) if isinstance(out, np.ndarray) else out
    if tuple(out_t.shape) != (n, h, w) or out_t.dtype != host.dtype:
        raise ValueError("out must be (N, H, W) of the frame dtype")
    staged = None if out_t.is_pinned() else torch.empty(out_t.shape, dtype=out_t.dtype).pin_memory()
    dst = out_t if staged is None else staged
    _lib.load(require_cuda=True)
    dev = torch.device("cuda", torch.cuda.current_device())
    depth = max(1, min(int(depth), n))
    s_in, s_cmp, s_out = _pipeline_streams(dev)
    caller = torch.cuda.current_stream(dev)
    for st in (s_in, s_cmp, s_out):
        st.wait_stream(caller)
    # The options of the fused pipeline are planned ONCE and every frame in flight owns one set of device buffers
    # (planes, ping-pong scratch, reconstruction): after the first `depth` frames nothing large is allocated -- with
    # wow() called per frame the caching allocator split and re-grew its blocks while the host ran ahead of the device
    # (a cudaMalloc of 0.7 GB per frame: 25 ms instead of 1.3).  h > 0, preserve_variance and per-pixel noise maps take
    # wow() itself, frame by frame.
    opts = dict(kwargs)
    unknown = set(opts) - {"n_scales", "weights", "whitening", "denoise_coefficients", "noise", "bilateral",
                           "bilateral_scaling", "soft_threshold", "preserve_variance", "gamma", "gamma_min", "gamma_max", "h"}
    if unknown:
        raise TypeError(f"wow_stream() got unexpected keyword arguments {sorted(unknown)}")
    plan = None
    if opts.get("h", 0) == 0 and not opts.get("preserve_variance", False):
        nz0 = _scalar_noise(opts.get("noise"), dev)
        if not (isinstance(nz0, str) and nz0 == "map"):
            n_scales, sigma_bilateral, wts, dns = _wow_plan((h, w), scaling_function, opts.get("n_scales"),
                                                            opts.get("weights", []), opts.get("denoise_coefficients", []),
                                                            opts.get("bilateral"), h=0)
            plan = (n_scales, wts, dns, sigma_bilateral, opts.get("bilateral_scaling", False),
                    opts.get("whitening", True), opts.get("soft_threshold", True), nz0)
    d_in = [torch.empty((h, w), dtype=host.dtype, device=dev) for _ in range(depth)]
    d_out = [None] * depth
    bufs = [dict() for _ in range(depth)]
    ev_in = [torch.cuda.Event() for _ in range(depth)]
    ev_cmp = [torch.cuda.Event() for _ in range(depth)]
    ev_out = [torch.cuda.Event() for _ in range(depth)]
    for i in range(n):
        b = i % depth
        with torch.cuda.stream(s_in):
            if i >= depth:
                s_in.wait_event(ev_cmp[b])       # the WOW that read d_in[b] is done
            d_in[b].copy_(host[i], non_blocking=True)
            ev_in[b].record(s_in)
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev_in[b])
            if i >= depth:
                s_cmp.wait_event(ev_out[b])      # the download of d_out[b] is done (its memory is reused)
            if plan is not None:
                L, wts, dns, sigma_bilateral, bil_scaling, whitening, soft, nz0 = plan
                recon, _, _ = _wow_stack(d_in[b].unsqueeze(0), scaling_function, L, wts, dns, sigma_bilateral,
                                         bil_scaling, whitening, soft, nz0, buffers=bufs[b])
                d_out[b] = recon[0]
            else:
                d_out[b], _ = wow(d_in[b], scaling_function=scaling_function, **kwargs)
            ev_cmp[b].record(s_cmp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp[b])
            dst[i].copy_(d_out[b], non_blocking=True)
            if plan is None:
                d_out[b].record_stream(s_out)
            ev_out[b].record(s_out)
    for st in (s_in, s_cmp, s_out):
        caller.wait_stream(st)
    s_out.synchronize()  # the result lives in host memory: hand it back complete
    if staged is not None:
        out_t.copy_(staged)
    return out if (isinstance(out, np.ndarray) or not was_numpy) else out_t.numpy()


def _wow_coefficients(co, weights, whitening, denoise_coefficients, bilateral, soft_threshold, plan=None):
    """wow() on already computed coefficients (watroo/utils.py:128-131, :152-153): planes are whitened in place,
    in plane order, with the lazily estimated noise of the object."""
    lib = _lib.load(require_cuda=True)
    sf = co.scaling_function
    if plan is None:
        n_scales, _, wts, dns = _wow_plan(co.data.shape[1:], sf.__class__, None, weights, denoise_coefficients,
                                          bilateral, from_coefficients=co)
        if n_scales != len(co) - 1:
            n_scales = len(co) - 1  # the warning branch cannot add planes to existing coefficients
    else:
        n_scales, _, wts, dns = plan
    data = co.data
    L = n_scales
    tmp = torch.empty_like(data[0])
    for s in range(L):
        d, wt = dns[s], wts[s]
        plane = data[s]
        if whitening:
            # local power from the raw plane, significance from the raw plane, then both factors (utils.py:194-203)
            nz = co._noise_arg(d) if d != 0 else _Noise()
            if nz.map is not None:
                sig = co.significance(d, s, soft_threshold=soft_threshold)
                _whiten_scale(lib, plane, tmp, s, sf, 0, 0.0, 1.0, _Noise(), wt)
                plane.copy_((tmp.to(torch.float64) * sig.to(torch.float64)).to(plane.dtype))
            else:
                mode = (1 if soft_threshold else 2) if d != 0 else 0
                _whiten_scale(lib, plane, tmp, s, sf, mode, d, co.sigma_e[s] if d != 0 else 1.0, nz, wt)
                plane.copy_(tmp)
        else:
            co._denoise_plane(lib, s, d, wt, soft_threshold)
    last = data[L]
    if whitening:
        mom = plane_moments(last)
        with torch.cuda.device(last.device):
            _lib.check(lib.wb_residual_rescale(last.data_ptr(), last.numel(), 1, 0, _lib.dtype_code(last.dtype),
                                               mom.data_ptr(), float(wts[L]), _lib.stream_ptr(last.device)))
    elif wts[L] != 1:
        last.mul_(wts[L])
    return synthesis(data), co


def _wow_general(co, weights, whitening, denoise_coefficients, bilateral, soft_threshold, preserve_variance, gamma,
                 gamma_min, gamma_max, h, plan=None):
    """wow() with the gamma blend (h > 0) and / or preserve_variance (watroo/utils.py:157-158, :174-217), plane by
    plane on raw coefficients: the same kernels as the fused route (K3 whitening, significance, moments, synthesis)
    plus a few device-scalar multiplications; nothing synchronises with the host."""
    lib = _lib.load(require_cuda=True)
    sf = co.scaling_function
    if plan is None:
        n_scales, _, wts, dns = _wow_plan(co.data.shape[1:], sf.__class__, None, weights, denoise_coefficients,
                                          bilateral, from_coefficients=co, h=h)
        n_scales = len(co) - 1
    else:
        n_scales, wts, dns = plan
    data = co.data
    dt = data.dtype
    L = n_scales
    white = whitening and h < 1  # utils.py:186, :193
    gamma_scaled = torch.zeros_like(data[0]) if h > 0 else None
    tmp = torch.empty_like(data[0]) if white else None
    for s in range(L):
        d, wt = dns[s], wts[s]
        plane = data[s]
        power_norm = None
        if preserve_variance:
            mom = plane_moments(plane)[0]                       # [mean, var, std] of the raw plane, float64
            power_norm = torch.sqrt(mom[1] + mom[0] * mom[0]).to(dt)  # sqrt(mean(c**2)), utils.py:182
        nz = co._noise_arg(d) if d != 0 else _Noise()           # lazily, from the current plane 0 (wavelets.py:131)
        if nz.map is not None:
            sig = co.significance(d, s, soft_threshold=soft_threshold).to(torch.float64)
            if white:
                _whiten_scale(lib, plane, tmp, s, sf, 0, 0.0, 1.0, _Noise(), wt)
                tmp.copy_((tmp.to(torch.float64) * sig).to(dt))
            plane.copy_((plane.to(torch.float64) * sig).to(dt))
        else:
            mode = (1 if soft_threshold else 2) if d != 0 else 0
            if white:
                _whiten_scale(lib, plane, tmp, s, sf, mode, d, co.sigma_e[s] if d != 0 else 1.0, nz, wt)
            co._denoise_plane(lib, s, d, 1, soft_threshold)     # plane <- raw * significance
        if gamma_scaled is not None:
            gamma_scaled += plane                               # utils.py:200-201: after significance, before weighting
        if white:
            plane.copy_(tmp if power_norm is None else tmp * power_norm)
        else:
            factor = wt if power_norm is None else power_norm * wt
            if power_norm is not None or wt != 1:
                plane.mul_(factor)
    last = data[L]
    mom = plane_moments(last)[0] if (preserve_variance or white) else None
    if gamma_scaled is not None:
        gamma_scaled += last
    factor = torch.ones((), dtype=torch.float64, device=last.device) * wts[L]
    if preserve_variance:
        factor = factor * mom[2].to(dt).to(torch.float64)       # power_norm = np.std(c), utils.py:180
    if white:
        std = mom[2].to(dt).to(torch.float64)
        factor = factor / torch.where(std <= 0, torch.full_like(std, 1e-15), std)  # utils.py:187-189
    last.mul_(factor.to(dt))
    recon = synthesis(data)
    if gamma_scaled is not None:  # utils.py:207-217
        lo = gamma_scaled.min() if gamma_min is None else gamma_min
        hi = gamma_scaled.max() if gamma_max is None else gamma_max
        gamma_scaled -= lo
        gamma_scaled /= (hi - lo)
        gamma_scaled.clamp_(0, 1)
        gamma_scaled.pow_(1 / gamma)
        recon = (1 - h) * recon + h * gamma_scaled
    return recon, co


# ---------------------------------------------------------------------------------------------------------------
# Callers of the transform beyond denoise / wow: enhance, richardson_lucy
# ---------------------------------------------------------------------------------------------------------------
def prepare_params(param, ndims):
    """Per-channel parameter lists of ``enhance`` (watroo/utils.py:10-33), quirks included (a scalar given for a
    3-channel image becomes the SAME inner list for every channel)."""
    if ndims == 2:
        if param is None:
            return []
        if type(param) is not list:
            return [param]
        return copy.copy(param)
    if type(param) is not list:
        return [[], ] * ndims if param is None else [[param], ] * ndims
    if len(param) != ndims:
        raise ValueError("Invalid number of parameters")
    out = [prepare_params(p, 2) for p in param]
    if None in out:
        out[out.index(None)] = []
    return out


def enhance(*args, weights=None, denoise=None, soft_threshold=True, out=None, **kwargs):
    """De-noising and / or enhancement by modification of the wavelet coefficients (watroo/utils.py:36-80): per
    channel (a 3-D input is three channels, channel first) transform -> ``Coefficients.denoise(denoise,
    weights=weights)`` -> sum of the planes.  ``args[1]``, when given, is the noise (per channel for 3 channels);
    ``kwargs`` go to ``AtrousTransform``.  NumPy in -> NumPy out (``out`` is filled when given)."""
    img = args[0]
    was_numpy = not isinstance(img, torch.Tensor)
    ndim = img.ndim
    channels = [0, 1, 2] if ndim == 3 else [Ellipsis]
    weights = prepare_params(weights, ndim)
    denoise_p = prepare_params(denoise, ndim)
    atrous = AtrousTransform(**kwargs)
    results = []
    for c in channels:
        dns = denoise_p if c is Ellipsis else denoise_p[c]
        wgt = weights if c is Ellipsis else weights[c]
        if len(wgt) < len(dns):
            wgt.extend([1] * (len(dns) - len(wgt)))
        elif len(dns) < len(wgt):
            dns.extend([0] * (len(wgt) - len(dns)))
        coeffs = atrous(img[c], len(wgt))
        if len(args) == 2:
            coeffs.noise = args[1] if c is Ellipsis else args[1][c]
        else:
            coeffs.noise = coeffs.get_noise()
        coeffs.denoise(dns, weights=wgt, soft_threshold=soft_threshold)
        results.append(synthesis(coeffs.data))
    res = results[0] if ndim != 3 else torch.stack(results)
    if out is not None:
        if isinstance(out, torch.Tensor):
            out.copy_(res)
        else:
            out[...] = res.cpu().numpy()
        return out
    return _result(res, was_numpy)


def _filter2d(img, kernel_dev, flip):
    """cv2.filter2D(img, -1, kernel, (-1,-1), 0, BORDER_REFLECT) on the device (wb_filter2d); flip rotates the kernel."""
    lib = _lib.load(require_cuda=True)
    out = torch.empty_like(img)
    kh, kw = kernel_dev.shape
    with torch.cuda.device(img.device):
        _lib.check(lib.wb_filter2d(img.data_ptr(), out.data_ptr(), img.shape[0], img.shape[1], img.stride(0),
                                   out.stride(0), kernel_dev.data_ptr(), kh, kw, 1 if flip else 0,
                                   _lib.dtype_code(img.dtype), _lib.stream_ptr(img.device)))
    return out


def richardson_lucy(data, psf, iterations=10, denoise_coefficients=(5, 2, 1), threshold_type='soft',
                    uniform_init=False, persistent_mrs=True, fft=False):
    """Wavelet-regularised Richardson-Lucy deconvolution (watroo/utils.py:222-290), every step on the device.

    Per iteration: phi = psf (*) psi (cv2.filter2D with the flipped PSF and the symmetric border, or a circular FFT
    convolution when ``fft``), residual ``data - phi`` -> à trous transform -> significance of every scale against the
    noise of the FIRST transform -> multiresolution support (persistent or not; hard: a mask that only grows, soft: a
    running product applied with the exponent 1/(iteration+1)) -> synthesis -> ``(res + phi) / phi`` -> correlation
    with the PSF -> multiplicative update of psi.  Returns psi (NumPy for NumPy input)."""
    img, was_numpy = to_device_image(data)
    soft = threshold_type == 'soft'
    level = len(denoise_coefficients)
    psf_t = torch.as_tensor(np.ascontiguousarray(psf) if not isinstance(psf, torch.Tensor) else psf)
    psf_dev = psf_t.to(device=img.device, dtype=img.dtype).contiguous()
    transform = AtrousTransform()
    coefficients = transform(img, level)
    if uniform_init:
        psi = torch.ones(img.shape, dtype=torch.float32, device=img.device)
        psi *= (img.sum() / img.numel()).to(torch.float32)
    else:
        coefficients.denoise(denoise_coefficients, soft_threshold=soft)
        psi = synthesis(coefficients.data)
    mrs = (torch.ones if soft else torch.zeros)((level,) + tuple(img.shape), dtype=torch.float64, device=img.device)
    if fft:
        h, w = psi.shape
        kh, kw = psf_dev.shape
        padded = torch.zeros_like(psi)
        padded[h // 2 - kh // 2: h // 2 - kh // 2 + kh, w // 2 - kw // 2: w // 2 - kw // 2 + kw] = psf_dev
        # np.fft computes in float64 whatever the input dtype: the reference's fft route runs in double from here on
        fft_psf = torch.fft.rfft2(torch.roll(padded, (h // 2, w // 2), dims=(0, 1)).to(torch.float64))
        psf_conj = fft_psf.conj()
    for iteration in range(iterations):
        if fft:
            phi = torch.fft.irfft2(torch.fft.rfft2(psi.to(torch.float64)) * fft_psf, s=psi.shape)
        else:
            phi = _filter2d(psi.to(img.dtype) if psi.dtype != img.dtype else psi, psf_dev, flip=True)
        res = img - phi
        res_coefficients = transform(res, level)
        res_coefficients._noise = coefficients._noise  # None when uniform_init: then estimated from the residual
        for s, c in enumerate(denoise_coefficients):
            sig = res_coefficients.significance(c, s, soft_threshold=soft)
            plane = res_coefficients.data[s]
            if not soft:
                sig = sig.to(torch.bool) if sig.dtype != torch.bool else sig
                if persistent_mrs:
                    mrs[s][sig] = 1
                else:
                    mrs[s] = sig.to(torch.float64)
                plane.copy_((plane.to(torch.float64) * mrs[s]).to(plane.dtype))
            else:
                if persistent_mrs:
                    mrs[s] *= sig.to(torch.float64)
                else:
                    mrs[s] = sig.to(torch.float64)
                plane.copy_((plane.to(torch.float64) * mrs[s] ** (1 / (iteration + 1))).to(plane.dtype))
        res = synthesis(res_coefficients.data)
        res += phi
        res /= phi
        if fft:
            conv = torch.fft.irfft2(torch.fft.rfft2(res.to(torch.float64)) * psf_conj, s=res.shape)
        else:
            conv = _filter2d(res, psf_dev, flip=False)
        psi = psi * conv.to(psi.dtype)
    return _result(psi, was_numpy)
