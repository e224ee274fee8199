"""Stand-in for numexpr, used ONLY by make_golden.py when the real package is not installed.

The reference has a single numexpr call site (watroo/wavelets.py:97), an element-wise expression over arrays in
the caller's frame.  Evaluating it with NumPy gives the same values to ~1 ulp of the image dtype.
"""
import sys

import numpy as np

__version__ = "0.0-shim"


def evaluate(expr, out=None, **_ignored):
    frame = sys._getframe(1)
    names = dict(frame.f_globals)
    names.update(frame.f_locals)
    names.update(exp=np.exp, sqrt=np.sqrt, log=np.log, abs=np.abs)
    result = eval(expr, {"__builtins__": {}}, names)  # noqa: S307 - fixed expression from the reference
    if out is not None:
        out[...] = result
        return out
    return result
