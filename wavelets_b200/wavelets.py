"""Host-side mirror of the reference's transform API (watroo/wavelets.py:108-149, :290-444) on top of the C ABI.

``AtrousTransform(Triangle|B3spline)(img, n_scales)`` returns a ``Coefficients`` object whose ``.data`` is a torch
CUDA tensor of shape ``(n_scales + 1, H, W)`` holding ``[w_0 .. w_{L-1}, c_L]``.  Inputs may be NumPy arrays (copied
to the current CUDA device) or torch tensors.  All arithmetic runs in the sm_100a kernels of
``libwavelets_b200.so``; there is no CPU path.
"""
from __future__ import annotations

import copy

import numpy as np
import torch

from . import _lib
from .scaling import AbstractScalingFunction, B3spline, Triangle

__all__ = ["AtrousTransform", "B3spline", "Triangle", "Coefficients", "atrous_scale"]

# integer / big-endian inputs are recast to float64 (watroo/wavelets.py:297, :319-320)
_RECAST_NUMPY = tuple(np.dtype(t) for t in (np.int16, np.uint16, np.int32, np.uint32, np.int64, ">f4", ">f8"))
_RECAST_TORCH = (torch.int16, torch.int32, torch.int64)


def _device():
    _lib.load(require_cuda=True)
    return torch.device("cuda", torch.cuda.current_device())


def to_device_image(arr, ndim_ok=(2,)):
    """Bring an image to the device as a float32/float64 tensor with unit inner stride.

    Returns (tensor, was_numpy).  Follows the reference's dtype rule: float32/float64 pass through, the integer
    and big-endian types of ``_recasting_types`` become float64; anything else is rejected (the reference silently
    computes in uint8 or crashes inside OpenCV for float16 -- neither is worth reproducing)."""
    was_numpy = not isinstance(arr, torch.Tensor)
    if was_numpy:
        arr = np.asarray(arr)
        if arr.ndim > 3:
            raise ValueError("Unsupported number of dimensions")  # watroo/wavelets.py:316-317
        if arr.dtype in _RECAST_NUMPY:
            arr = arr.astype(np.float64)
        if arr.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError(f"unsupported image dtype {arr.dtype}: expected float32/float64 or a recast integer type")
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(_device(), non_blocking=False)
    else:
        if arr.ndim > 3:
            raise ValueError("Unsupported number of dimensions")
        t = arr
        if t.dtype in _RECAST_TORCH:
            t = t.to(torch.float64)
        if t.dtype not in (torch.float32, torch.float64):
            raise TypeError(f"unsupported image dtype {t.dtype}: expected float32/float64 or a recast integer type")
        if not t.is_cuda:
            t = t.to(_device())
        if t.stride(-1) != 1:
            t = t.contiguous()
    if t.ndim not in ndim_ok:
        if t.ndim == 3 or t.ndim == 1:
            raise NotImplementedError(
                "wavelets_b200 accelerates the 2-D path; a 3-D input is a volume in the reference "
                "(watroo/wavelets.py:322), not a batch -- use AtrousTransform.batch() for stacks of frames")
        raise ValueError("Unsupported number of dimensions")
    return t, was_numpy


def _frame_layout(t):
    """(batch, H, W, pitch, bstride) of a 2-D image or a 3-D stack with unit inner stride."""
    if t.ndim == 2:
        return 1, t.shape[0], t.shape[1], t.stride(0), 0
    return t.shape[0], t.shape[1], t.shape[2], t.stride(1), t.stride(0)


def atrous_scale(src, scale, scaling_function, out_c=None, out_w=None):
    """One scale of the plain cascade on device tensors: ``out_c = S_scale[src]``, ``out_w = src - out_c``.

    Thin wrapper over ``wb_atrous_scale`` (replaces ``convolution`` watroo/wavelets.py:35-45 plus the subtraction of
    :442).  ``src`` is (H, W) or (B, H, W); outputs are allocated when not given (pass ``False`` to skip one)."""
    lib = _lib.load(require_cuda=True)
    b, h, w, pitch, bstride = _frame_layout(src)
    if out_c is None:
        out_c = torch.empty(src.shape, dtype=src.dtype, device=src.device)
    if out_w is None:
        out_w = torch.empty(src.shape, dtype=src.dtype, device=src.device)
    pc = pw = 0
    lc = lw = (0, 0)
    if out_c is not False:
        pc, lc = out_c.data_ptr(), _frame_layout(out_c)[3:]
    if out_w is not False:
        pw, lw = out_w.data_ptr(), _frame_layout(out_w)[3:]
    with torch.cuda.device(src.device):
        _lib.check(lib.wb_atrous_scale(src.data_ptr(), pc, pw, b, h, w, pitch, bstride, lc[0], lc[1], lw[0], lw[1],
                                       int(scale), scaling_function.taps_code, _lib.dtype_code(src.dtype),
                                       _lib.stream_ptr(src.device)))
    return (None if out_c is False else out_c), (None if out_w is False else out_w)


class Coefficients:
    """Wavelet planes plus the metadata needed to threshold them (mirror of watroo/wavelets.py:108-149).

    ``data``: torch CUDA tensor ``(L+1, H, W)``; ``scaling_function``: instance; ``bilateral``: what the transform
    was run with (selects the sigma_e table); ``noise``: scalar / map, estimated lazily when first needed."""

    def __init__(self, data, scaling_function, bilateral=None):
        self.data = data
        self.scaling_function = scaling_function
        self.bilateral = bilateral
        self.noise = None

    def __len__(self):
        return len(self.data)

    def __array__(self, dtype=None, copy=None):
        """Host copy, so that ``np.sum(coefficients, axis=0)`` of the README keeps working."""
        host = self.data.detach().cpu().numpy()
        return host if dtype is None else host.astype(dtype, copy=False)

    @property
    def sigma_e(self):
        return self.scaling_function.sigma_e(bilateral=self.bilateral)


class AtrousTransform:
    """Dyadic 'à trous' transform (mirror of watroo/wavelets.py:290-328, standard algorithm :408-444)."""

    def __init__(self, scaling_function_class=B3spline, bilateral=None, bilateral_scaling=False):
        self.scaling_function_class = scaling_function_class
        self.bilateral = bilateral
        self.bilateral_scaling = bilateral_scaling

    def __call__(self, arr, level, recursive=False):
        """Transform a 2-D image over ``level`` scales -> ``Coefficients`` with ``level + 1`` planes.

        ``recursive`` is accepted for signature compatibility and ignored: the reference's recursive variant is a
        CPU-side optimisation of the same transform (it differs from the standard one only near the borders,
        watroo/wavelets.py:394-395); the device kernels always implement the standard algorithm."""
        img, _ = to_device_image(arr)
        scaling_function = self.scaling_function_class(img.ndim)
        planes = self._run(img, int(level), scaling_function)
        return Coefficients(planes, scaling_function, self.bilateral)

    def batch(self, frames, level):
        """NEW entry point (no reference equivalent): transform a stack ``(B, H, W)`` of independent frames in one
        launch per scale.  Returns a ``(B, level + 1, H, W)`` tensor.  Plain (non-bilateral) cascade only."""
        stack, _ = to_device_image(frames, ndim_ok=(3,))
        if self.bilateral is not None:
            raise NotImplementedError("batch() supports the plain cascade only")
        return self._run(stack, int(level), self.scaling_function_class(2))

    # -- internals ----------------------------------------------------------------------------------------------
    def _run(self, img, level, scaling_function):
        if level < 0:
            raise ValueError("level must be >= 0")
        lib = _lib.load(require_cuda=True)
        b, h, w, pitch, bstride = _frame_layout(img)
        shape = (level + 1, h, w) if img.ndim == 2 else (b, level + 1, h, w)
        planes = torch.empty(shape, dtype=img.dtype, device=img.device)
        if self.bilateral is None:
            scratch = torch.empty((2, b, h, w), dtype=img.dtype, device=img.device) if level > 1 else None
            with torch.cuda.device(img.device):
                _lib.check(lib.wb_atrous_transform(
                    img.data_ptr(), planes.data_ptr(), 0 if scratch is None else scratch.data_ptr(), b, h, w, pitch,
                    bstride, level, scaling_function.taps_code, _lib.dtype_code(img.dtype),
                    _lib.stream_ptr(img.device)))
            return planes
        raise NotImplementedError("bilateral cascade: kernel K2 not built yet")


def noise_weights(scaling_function, n_scales, n_trials=100, bilateral=None, fields=None, seed=None):
    raise NotImplementedError("compute_noise_weights: reduction kernels not built yet")
