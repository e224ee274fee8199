"""wavelets_b200 -- B200 (sm_100a) implementation of watroo's à trous / WOW hot path behind the reference's API.

    from wavelets_b200 import AtrousTransform, B3spline, Triangle, Coefficients, denoise, wow
"""
from .scaling import AbstractScalingFunction, B3spline, Triangle  # noqa: F401
from .wavelets import AtrousTransform, Coefficients, atrous_scale  # noqa: F401

__version__ = "0.1.0"
__all__ = ["AtrousTransform", "B3spline", "Triangle", "Coefficients", "atrous_scale"]
