"""wavelets_b200 -- B200 (sm_100a) implementation of watroo's à trous / WOW hot path behind the reference's API.

    from wavelets_b200 import AtrousTransform, B3spline, Triangle, Coefficients, denoise, wow

Same call surface as ``watroo`` (frederic-auchere/wavelets 0.0.4) for the 2-D path; all arithmetic runs in the CUDA
kernels of ``libwavelets_b200.so`` (C ABI: include/wavelets_b200.h).  There is no CPU fallback.
"""
from .scaling import AbstractScalingFunction, B3spline, Triangle  # noqa: F401
from .wavelets import AtrousTransform, Coefficients, atrous_scale, convolution  # noqa: F401
from .utils import denoise, enhance, generalized_anscombe, richardson_lucy, wow, wow_batch, wow_stream  # noqa: F401

__version__ = "0.1.0"
__all__ = ["AtrousTransform", "B3spline", "Triangle", "Coefficients", "generalized_anscombe", "convolution",
           "denoise", "wow", "wow_batch", "wow_stream", "atrous_scale", "enhance", "richardson_lucy"]
