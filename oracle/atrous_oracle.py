"""CPU oracle for the à trous / WOW hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module is a NumPy restatement of the algorithm of watroo 0.0.4 (frederic-auchere/wavelets) for the one
hot path this repository accelerates.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import it; nothing under ``wavelets_b200/`` does, and the product path raises if its CUDA library is
missing instead of falling back to this file.

Parity status: PINNED.  ``tests/golden/*.npz`` were produced by running the real reference (``/root/reference``,
with a 10-line ``numexpr`` shim because numexpr is not installed in this image) through
``tests/golden/make_golden.py``; ``tests/test_oracle.py`` checks every function below against those vectors, against
the reference's own known-answer test (constant image -> zero detail planes, ``tests/test_wavelets.py:8-13``), the
perfect-reconstruction identity, and the README equivalences (``README.md:46-61``).

Third-party arithmetic the reference leans on (not vendored under /root/reference; only lower bounds are pinned in
``requirements.txt:1-5``; versions below are the ones installed in this image and used to make the goldens):
  * opencv-python-headless 4.13.0.92  ``cv2.filter2D(..., BORDER_REFLECT)``  (``watroo/wavelets.py:39-45``).
    Published semantics: *correlation* of the image with the kernel, anchor at the kernel centre, border pixels
    taken by half-sample symmetric reflection ``fedcba|abcdefgh|hgfedcb``.  For float32 images OpenCV sums in
    float32 for small kernels and goes through a double-precision DFT for big ones (the result is then the
    correctly rounded double answer).  The ``numpy`` backend below restates that as: accumulate in float64, round
    once to the image dtype.  The ``cv2`` backend issues the very same call as the reference.
  * numexpr (absent here)  ``ne.evaluate('k*exp(-((image - shifted)**2)/bilateral_variance/2)')``
    (``watroo/wavelets.py:97``): an element-wise expression in the image dtype; restated with ``np.exp``.
  * scipy 1.18.1 ``special.erf`` (``watroo/wavelets.py:138``), numpy 2.3.5 ``median/std/sum/pad`` with NEP-50
    promotion (an fp32 array divided by an fp64 NumPy scalar yields fp64 -- this matters for ``significance``).
"""
from __future__ import annotations

import copy
import math
import warnings

import numpy as np
from scipy import special

try:  # the reference's own smoothing primitive; optional so the oracle still runs where OpenCV is absent
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

__all__ = [
    "TAPS", "SIGMA_E_2D", "SIGMA_E_2D_BILATERAL", "reflect_index", "smooth", "local_variance", "bilateral_smooth",
    "atrous_transform", "get_noise", "significance", "denoise_planes", "denoise", "wow", "compute_noise_weights",
    "default_backend", "wow_default_scales", "solar_like", "emax", "SIGMA_E_1D", "SIGMA_E_3D", "mirror_index",
    "smooth_nd", "atrous_transform_recursive", "filter2d_reflect", "enhance", "richardson_lucy",
]

# ---------------------------------------------------------------------------------------------------------------
# Constants (watroo/wavelets.py:239-252 Triangle, :268-281 B3spline).  The sigma_e tables are data recorded by the
# reference authors with compute_noise_weights(); parity of significance/denoise/wow requires the same numbers.
# ---------------------------------------------------------------------------------------------------------------
TAPS = {
    "triangle": np.array([1 / 4, 1 / 2, 1 / 4]),
    "b3spline": np.array([1 / 16, 1 / 4, 3 / 8, 1 / 4, 1 / 16]),
}
SIGMA_E_1D_LEN = {"triangle": 11, "b3spline": 11}  # len(sigma_e_1d): sets the noise-field side (wavelets.py:225)
SIGMA_E_2D = {
    "triangle": np.array([0.7999247, 0.27308452, 0.11998217, 0.05793947, 0.0288104, 0.01447795, 0.00733832,
                          0.0037203, 0.00192882, 0.00098568, 0.00048533]),
    "b3spline": np.array([8.907e-01, 2.0072e-01, 8.5551e-02, 4.1261e-02, 2.0470e-02, 1.0232e-02, 5.1435e-03,
                          2.6008e-03, 1.3161e-03, 6.7359e-04, 4.0040e-04]),
}
SIGMA_E_2D_BILATERAL = {
    "triangle": np.array([0.31063172, 0.34575647, 0.23712331, 0.13559906, 0.07172004, 0.03665405, 0.01850046,
                          0.00928768, 0.00465967, 0.00234445, 0.00119249]),
    "b3spline": np.array([0.38234752, 0.24305799, 0.16012153, 0.10633541, 0.07083733, 0.04728659, 0.03163678,
                          0.02122341, 0.01429102, 0.00952376]),
}


def default_backend() -> str:
    return "cv2" if cv2 is not None else "numpy"


SIGMA_E_1D = {
    "triangle": np.array([0.60840933, 0.33000059, 0.21157957, 0.145824, 0.10158388, 0.07155912, 0.04902655,
                          0.03529812, 0.02409187, 0.01722846, 0.01144442]),
    "b3spline": np.array([0.72514976, 0.28538683, 0.17901161, 0.12222841, 0.08469601, 0.06027006, 0.04242257,
                          0.02919823, 0.01805671, 0.01383672, 0.00943623]),
}
SIGMA_E_3D_BILATERAL = {
    "triangle": np.array([0.3828863, 0.36182913, 0.19520299, 0.08498861, 0.03363142]),
    "b3spline": np.array([0.44111772, 0.3552894, 0.16137159, 0.05769064, 0.01932497]),
}
SIGMA_E_3D = {
    "triangle": np.array([0.89736751, 0.19514386, 0.06239262, 0.02311278, 0.00939645]),
    "b3spline": np.array([0.95633954, 0.12491933, 0.03933029, 0.01489642, 0.0064108]),
}


def sigma_e(name: str, bilateral=None, ndim: int = 2) -> np.ndarray:
    """wavelets.py:199-219: the table of the data's dimensionality; the bilateral table (restated for 2-D only) is
    selected whenever ``bilateral is not None``."""
    if bilateral is not None:
        if ndim == 1:
            raise AttributeError("sigma_e_1d_bilateral")  # the reference has no such table
        return {2: SIGMA_E_2D_BILATERAL, 3: SIGMA_E_3D_BILATERAL}[ndim][name]
    return {1: SIGMA_E_1D, 2: SIGMA_E_2D, 3: SIGMA_E_3D}[ndim][name]


# ---------------------------------------------------------------------------------------------------------------
# Border rule and per-scale smoothing  (watroo/wavelets.py:35-45 + :191-197)
# ---------------------------------------------------------------------------------------------------------------
def reflect_index(i, n):
    """Half-sample symmetric reflection of integer index(es) ``i`` into [0, n): ``fedcba|abcdefgh|hgfedcb``.

    Same rule as cv2.BORDER_REFLECT (wavelets.py:45) and np.pad(mode='symmetric') (wavelets.py:77), for any number
    of reflections (period 2n)."""
    m = np.mod(i, 2 * n)
    return np.where(m < n, m, 2 * n - 1 - m)


def _dense_kernel(taps: np.ndarray, s: int) -> np.ndarray:
    """The dilated dense 2-D kernel of wavelets.py:191-197: outer(h,h) scattered with stride 2**s."""
    k = len(taps)
    side = (k - 1) * 2 ** s + 1
    dense = np.zeros((side, side))
    dense[:: 2 ** s, :: 2 ** s] = np.outer(taps, taps)
    return dense


def smooth(arr: np.ndarray, name: str, s: int = 0, backend: str | None = None) -> np.ndarray:
    """c_{s+1} = S_s[c_s]   (wavelets.py:35-45, 2-D branch).

    out(y,x) = sum_i sum_j h_i h_j arr(R(y+(i-c)2^s), R(x+(j-c)2^s)), R = reflect_index; output dtype = input dtype.
    backend 'cv2'   : the reference's own call (dense dilated kernel, filter2D, BORDER_REFLECT).
    backend 'numpy' : separable gather, float64 accumulation, one rounding to the image dtype.
    """
    if arr.ndim != 2:
        return smooth_nd(arr, name, s)
    backend = backend or default_backend()
    taps = TAPS[name]
    if backend == "cv2":
        out = np.empty_like(arr)
        cv2.filter2D(arr, -1, _dense_kernel(taps, s).astype(arr.dtype), out, (-1, -1), 0, cv2.BORDER_REFLECT)
        return out
    h, w = arr.shape
    c = len(taps) // 2
    d = 2 ** s
    a64 = arr.astype(np.float64)
    rows = np.zeros((h, w), dtype=np.float64)
    xs = np.arange(w)
    for j, t in enumerate(taps):
        rows += t * a64[:, reflect_index(xs + (j - c) * d, w)]
    out = np.zeros((h, w), dtype=np.float64)
    ys = np.arange(h)
    for i, t in enumerate(taps):
        out += t * rows[reflect_index(ys + (i - c) * d, h), :]
    return out.astype(arr.dtype)


def mirror_index(i, n):
    """Whole-sample reflection of integer index(es) ``i`` into [0, n): ``dcb|abcd|cba`` -- scipy.ndimage mode='mirror'
    (wavelets.py:69, the 1-D branch), for any number of reflections (period 2n - 2)."""
    if n == 1:
        return np.zeros_like(np.asarray(i))
    m = np.mod(i, 2 * n - 2)
    return np.where(m < n, m, 2 * n - 2 - m)


def smooth_nd(arr: np.ndarray, name: str, s: int = 0) -> np.ndarray:
    """c_{s+1} = S_s[c_s] for 1-D signals and 3-D volumes (wavelets.py:46-69), float64 accumulation, one rounding.

    1-D (:64-69): scipy.ndimage.convolve(arr, atrous_kernel(s), mode='mirror') -- whole-sample reflection.
    3-D (:46-63): the 2-D dilated smooth of every arr[i] slice (cv2.BORDER_REFLECT), then the dilated 1-D filter
    along axis 0 of every [:, :, i] slice (cv2.filter2D with the (K, 1) kernel, BORDER_REFLECT): half-sample
    symmetric reflection on all three axes."""
    taps = TAPS[name]
    c = len(taps) // 2
    d = 2 ** s
    a64 = arr.astype(np.float64)
    if arr.ndim == 1:
        xs = np.arange(arr.shape[0])
        out = np.zeros(arr.shape, dtype=np.float64)
        for j, t in enumerate(taps):
            out += t * a64[mirror_index(xs + (j - c) * d, arr.shape[0])]
        return out.astype(arr.dtype)
    if arr.ndim != 3:
        raise ValueError("Unsupported number of dimensions")
    out = a64
    for axis in (2, 1, 0):
        n = arr.shape[axis]
        idx = np.arange(n)
        acc = np.zeros(arr.shape, dtype=np.float64)
        for j, t in enumerate(taps):
            acc += t * np.take(out, reflect_index(idx + (j - c) * d, n), axis=axis)
        out = acc
        if axis == 1:
            out = out.astype(arr.dtype).astype(np.float64)  # the reference rounds to the volume dtype between its passes
    return out.astype(arr.dtype)


def atrous_transform_recursive(arr: np.ndarray, level: int, name: str = "b3spline", bilateral=None,
                               bilateral_scaling: bool = False) -> np.ndarray:
    """AtrousTransform(sf, bilateral, bilateral_scaling)(arr, level, recursive=True).data for a 2-D transform
    (wavelets.py:330-406), restated without the recursion: symmetric pad by (K // 2) * 2**(level - 1); at scale s every
    pixel's taps reflect (half-sample symmetric) inside its own decimated sub-array {o, o + 2^s, o + 2*2^s, ...}, which
    is what filtering each `conv[oy::2^s, ox::2^s]` sub-array with BORDER_REFLECT (plain) or through sdev_loc +
    atrous_convolution(mode='symmetric') (bilateral, :371-378) does; crop the pad at the end."""
    if arr.dtype in _RECAST:
        arr = np.float64(arr)
    taps = TAPS[name]
    c = len(taps) // 2
    hw = c * 2 ** (level - 1)
    cur = np.pad(arr, hw, mode="symmetric").astype(np.float64)
    planes = np.empty((level + 1,) + cur.shape, dtype=arr.dtype)
    sb = _bilateral_list(bilateral, level)

    def lattice(n, off, d):
        i = np.arange(n)
        o, t = i % d, i // d
        n_sub = (n - o + d - 1) // d
        m = np.mod(t + off, 2 * n_sub)
        return o + np.where(m < n_sub, m, 2 * n_sub - 1 - m) * d

    def smooth_lattice(a64, d):
        rows = np.zeros_like(a64)
        for j, t in enumerate(taps):
            rows += t * a64[:, lattice(a64.shape[1], j - c, d)]
        out = np.zeros_like(a64)
        for i, t in enumerate(taps):
            out += t * rows[lattice(a64.shape[0], i - c, d), :]
        return out

    cur = cur.astype(arr.dtype)
    for s in range(level):
        d = 2 ** s
        if bilateral is None:
            nxt = smooth_lattice(cur.astype(np.float64), d).astype(arr.dtype)
        else:
            # sdev_loc in the image dtype (squares, subtraction), then the range-weighted gather on the same lattice
            mean2 = smooth_lattice(cur.astype(np.float64), d).astype(arr.dtype) ** 2
            vari = smooth_lattice((cur ** 2).astype(np.float64), d).astype(arr.dtype)
            vari -= mean2
            vari[vari <= 0] = 1e-20
            vari = vari * sb[s] ** 2
            if bilateral_scaling:
                vari *= s + 1
            k2d = np.outer(taps, taps).astype(arr.dtype)
            out = k2d[c, c] * cur
            norm = np.full_like(cur, k2d[c, c])
            for i in range(len(taps)):
                yy = lattice(cur.shape[0], i - c, d)
                for j in range(len(taps)):
                    if i == c and j == c:
                        continue
                    shifted = cur[yy][:, lattice(cur.shape[1], j - c, d)]
                    weight = k2d[i, j] * np.exp(-((cur - shifted) ** 2) / vari / 2)
                    norm += weight
                    out += shifted * weight
            nxt = (out / norm).astype(arr.dtype)
        planes[s] = cur - nxt
        cur = nxt
    planes[level] = cur
    return planes[:, hw:hw + arr.shape[0], hw:hw + arr.shape[1]].copy()


def local_variance(arr: np.ndarray, name: str, s: int, backend: str | None = None) -> np.ndarray:
    """sdev_loc(..., variance=True)  (wavelets.py:24-32): S_s[x^2] - (S_s[x])^2, non-positive -> 1e-20.

    Squares and the subtraction are done in the image dtype exactly as the reference does (this cancels badly in
    fp32 for bright pixels)."""
    mean2 = smooth(arr, name, s, backend) ** 2
    vari = smooth(arr ** 2, name, s, backend)
    vari -= mean2
    vari[vari <= 0] = 1e-20
    return vari


def bilateral_smooth(arr: np.ndarray, name: str, variance: np.ndarray, s: int) -> np.ndarray:
    """atrous_convolution(image, kernel, bilateral_variance, s, 'symmetric')  (wavelets.py:74-105).

    out = (k_c x + sum_t g_t x_t) / (k_c + sum_t g_t),  g_t = k_t exp(-(x - x_t)^2 / V / 2) over the K^2-1
    off-centre taps, x_t read through the symmetric border, V taken at the output pixel.  All arithmetic in the
    image dtype (the kernel is cast with .astype(arr.dtype), wavelets.py:438)."""
    if arr.ndim != 2:
        return bilateral_smooth_nd(arr, name, variance, s)
    taps = TAPS[name]
    k2d = np.outer(taps, taps).astype(arr.dtype)
    n = len(taps)
    c = n // 2
    d = 2 ** s
    h, w = arr.shape
    padded = np.pad(arr, [(c * d, c * d), (c * d, c * d)], mode="symmetric")
    out = k2d[c, c] * arr
    norm = np.full_like(arr, k2d[c, c])
    for i in range(n):
        for j in range(n):
            if i == c and j == c:
                continue
            shifted = padded[i * d:i * d + h, j * d:j * d + w]
            k = k2d[i, j]
            weight = k * np.exp(-((arr - shifted) ** 2) / variance / 2)
            norm += weight
            out += shifted * weight
    out /= norm
    return out


def bilateral_smooth_nd(arr: np.ndarray, name: str, variance: np.ndarray, s: int) -> np.ndarray:
    """atrous_convolution on a 1-D signal or a 3-D volume (wavelets.py:74-105 is dimension-generic): the kernel is the
    n-fold tensor product of the taps, the border is np.pad 'symmetric' on every axis (also in 1-D, where the plain
    smooth -- and hence the variance -- uses 'mirror')."""
    taps = TAPS[name]
    n = len(taps)
    c = n // 2
    d = 2 ** s
    kern = taps
    for _ in range(arr.ndim - 1):
        kern = np.multiply.outer(kern, taps)
    kern = kern.astype(arr.dtype)
    padded = np.pad(arr, [(c * d, c * d)] * arr.ndim, mode="symmetric")
    centre = (c,) * arr.ndim
    out = kern[centre] * arr
    norm = np.full_like(arr, kern[centre])
    for index in np.ndindex(*kern.shape):
        if index == centre:
            continue
        shifted = padded[tuple(slice(i * d, i * d + m) for i, m in zip(index, arr.shape))]
        weight = kern[index] * np.exp(-((arr - shifted) ** 2) / variance / 2)
        norm += weight
        out += shifted * weight
    out /= norm
    return out


# ---------------------------------------------------------------------------------------------------------------
# Cascade  (watroo/wavelets.py:307-328, :408-444)
# ---------------------------------------------------------------------------------------------------------------
_RECAST = [np.dtype(t) for t in (np.int16, np.uint16, np.int32, np.uint32, np.int64, ">f4", ">f8")]


def _bilateral_list(bilateral, level):
    """wavelets.py:421-424: scalar -> [b]*(level+1); list -> copy padded with 1 up to level+1 entries."""
    sb = copy.copy(bilateral) if type(bilateral) is list else [bilateral, ] * (level + 1)
    if len(sb) <= level:
        sb.extend([1, ] * (level - len(sb) + 1))
    return sb


def atrous_transform(arr: np.ndarray, level: int, name: str = "b3spline", bilateral=None,
                     bilateral_scaling: bool = False, backend: str | None = None) -> np.ndarray:
    """AtrousTransform(sf, bilateral, bilateral_scaling)(arr, level).data  (wavelets.py:307-328, :408-444).

    Planes [w_0 .. w_{L-1}, c_L] with c_0 = arr, c_{s+1} = S_s[c_s] (or its bilateral variant), w_s = c_s - c_{s+1}.
    Integer and big-endian inputs are recast to float64 (wavelets.py:297,319-320); the input is never modified."""
    if arr.ndim > 3:
        raise ValueError("Unsupported number of dimensions")  # wavelets.py:316-317
    if arr.dtype in _RECAST:
        arr = np.float64(arr)
    sb = _bilateral_list(bilateral, level)
    coeffs = np.empty((level + 1,) + arr.shape, dtype=arr.dtype)
    coeffs[0] = arr
    for s in range(level):
        if bilateral is None:
            coeffs[s + 1] = smooth(coeffs[s], name, s, backend)
        else:
            variance = local_variance(coeffs[s], name, s, backend) * sb[s] ** 2
            if bilateral_scaling:
                variance *= s + 1
            coeffs[s + 1] = bilateral_smooth(coeffs[s], name, variance, s)
        coeffs[s] -= coeffs[s + 1]
    return coeffs


# ---------------------------------------------------------------------------------------------------------------
# Noise, significance, denoise  (watroo/wavelets.py:126-149, watroo/utils.py:83-102)
# ---------------------------------------------------------------------------------------------------------------
def get_noise(planes: np.ndarray, name: str, bilateral=None):
    """Coefficients.get_noise (wavelets.py:126-127): MAD estimate median(|w_0|)/0.6745/sigma_e[0].

    dtype flow under NumPy>=2: median keeps the plane dtype, '/0.6745' keeps it (weak Python float), '/sigma_e[0]'
    promotes to float64 (sigma_e is a float64 NumPy scalar)."""
    return np.median(np.abs(planes[0])) / 0.6745 / sigma_e(name, bilateral, planes.ndim - 1)[0]


def significance(planes: np.ndarray, name: str, sigma, scale: int, noise, bilateral=None,
                 soft_threshold: bool = True) -> np.ndarray:
    """Coefficients.significance (wavelets.py:129-143) for an already known ``noise`` (scalar or map).

    sigma == 0 or scalar noise == 0 -> ones (plane dtype).  soft: erf(|w_s / (sigma*noise*sigma_e[s])|) -- a float64
    array even for fp32 planes; hard: |w_s| > sigma*noise*sigma_e[s] compared in float64 (bool array)."""
    if sigma != 0:
        if type(noise) is not np.ndarray:
            if noise == 0:
                return np.ones_like(planes[0])
        thr = sigma * noise * sigma_e(name, bilateral, planes.ndim - 1)[scale]
        if soft_threshold:
            return special.erf(np.abs(planes[scale] / thr))
        return np.abs(planes[scale]) > thr
    return np.ones_like(planes[0])


def denoise_planes(planes: np.ndarray, name: str, sigma, weights=None, noise=None, bilateral=None,
                   soft_threshold: bool = True):
    """Coefficients.denoise (wavelets.py:145-149), in place on ``planes``; returns the noise that was used.

    Only the first len(sigma) planes are touched; noise is estimated lazily from the *unmodified* plane 0 at the
    first scale whose sigma is non-zero (wavelets.py:131-132)."""
    if weights is None:
        weights = (1,) * len(sigma)
    for scl, (c, sig, wgt) in enumerate(zip(planes, sigma, weights)):
        if sig != 0 and noise is None:
            noise = get_noise(planes, name, bilateral)
        c *= wgt * significance(planes, name, sig, scl, noise, bilateral, soft_threshold)
    return noise


def generalized_anscombe(signal: np.ndarray, alpha=1, g=0, sigma=0, inverse: bool = False) -> np.ndarray:
    """Variance-stabilising transform of Poisson-Gaussian data and its algebraic inverse (wavelets.py:14-21)."""
    if inverse:
        return ((alpha * signal / 2) ** 2 + alpha * g - sigma ** 2 - 3 * alpha / 8) / alpha
    dum = alpha * signal + 3 * alpha ** 2 / 8 + sigma ** 2 - alpha * g
    dum[dum <= 0] = 0
    return 2 * np.sqrt(dum) / alpha


def denoise(data: np.ndarray, weights, name: str = "b3spline", noise=None, bilateral=None,
            soft_threshold: bool = True, backend: str | None = None, anscombe: bool = False) -> np.ndarray:
    """utils.denoise (utils.py:83-102): ``weights`` are the sigma thresholds and their count is the number of
    scales; result = sum of the planes in plane order (np.sum(axis=0)); ``anscombe`` wraps the whole thing in the
    generalized Anscombe transform and its inverse (utils.py:93-94, :99-100)."""
    if anscombe:
        data = generalized_anscombe(data)
    planes = atrous_transform(data, len(weights), name, bilateral=bilateral, backend=backend)
    denoise_planes(planes, name, weights, noise=noise, bilateral=bilateral, soft_threshold=soft_threshold)
    out = np.sum(planes, axis=0)
    return generalized_anscombe(out, inverse=True) if anscombe else out


# ---------------------------------------------------------------------------------------------------------------
# WOW  (watroo/utils.py:105-219; the h == 0, preserve_variance == False path plus `weights` / `whitening`)
# ---------------------------------------------------------------------------------------------------------------
def wow_default_scales(shape, name: str) -> int:
    """utils.py:122: round(log2(min(shape)) - log2(len(taps)))."""
    return int(np.round(np.log2(min(shape)) - np.log2(len(TAPS[name]))))


def wow(data: np.ndarray, name: str = "b3spline", n_scales=None, weights=(), whitening: bool = True,
        denoise_coefficients=(), noise=None, bilateral=None, bilateral_scaling: bool = False,
        soft_threshold: bool = True, preserve_variance: bool = False, gamma: float = 3.2, gamma_min=None,
        gamma_max=None, h: float = 0, backend: str | None = None):
    """utils.wow (utils.py:105-219) for an image input.

    Returns (recon, planes, noise): the synthesis (blended with the gamma-scaled image when h > 0, utils.py:207-217),
    the whitened planes (what ``coefficients.data`` holds after the call) and the noise value that was used (None if
    no significance was evaluated)."""
    weights = list(weights)
    denoise_coefficients = list(denoise_coefficients)
    max_scales = wow_default_scales(data.shape, name)
    if n_scales is None:
        n_scales = max_scales if h < 1 else len(denoise_coefficients)  # utils.py:123-124
    elif n_scales > max_scales:
        n_scales = max_scales
    table_len = len(sigma_e(name, bilateral, data.ndim))
    if len(denoise_coefficients) >= table_len:  # utils.py:135-138
        warnings.warn(f"Required number of scales lager then the maximum for scaling function. Using {table_len}.")
        n_scales = table_len
    sigma_bilateral = None if bilateral is None else _bilateral_list(bilateral, n_scales)  # utils.py:140-146

    planes = atrous_transform(data, n_scales, name, bilateral=sigma_bilateral, bilateral_scaling=bilateral_scaling,
                              backend=backend)
    gamma_scaled = np.zeros_like(planes[0]) if h > 0 else None  # utils.py:157-158

    wts = copy.copy(weights)  # utils.py:160-163
    if len(wts) <= n_scales:
        wts.extend([1, ] * (n_scales - len(wts) + 1))
    dns = copy.copy(denoise_coefficients)  # utils.py:165-170
    if len(dns) < n_scales:
        dns.extend([0, ] * (n_scales - len(dns)))
    if len(dns) == n_scales:
        dns.extend([1, ])

    for s, (c, w, d) in enumerate(zip(planes, wts, dns)):  # utils.py:174-203
        power = c ** 2
        if preserve_variance:  # utils.py:178-184
            power_norm = np.std(c) if s == n_scales else np.sqrt(np.mean(power))
        else:
            power_norm = 1
        if s == n_scales:
            if whitening and h < 1:
                local_power = np.std(c)
                if local_power <= 0:
                    local_power = 1e-15
            else:
                local_power = 1
        else:
            if whitening and h < 1:
                local_power = smooth(power, name, s, backend)  # plain smooth even if the transform was bilateral
                local_power[local_power <= 0] = 1e-15
                np.sqrt(local_power, out=local_power)
            else:
                local_power = 1
            if d != 0 and noise is None:
                noise = get_noise(planes, name, sigma_bilateral)
            c *= significance(planes, name, d, s, noise, sigma_bilateral, soft_threshold)
        if h > 0:
            gamma_scaled += c
        c *= w * power_norm / local_power
    recon = np.sum(planes, axis=0)
    if h > 0:  # utils.py:207-217
        if gamma_min is None:
            gamma_min = gamma_scaled.min()
        if gamma_max is None:
            gamma_max = gamma_scaled.max()
        gamma_scaled -= gamma_min
        gamma_scaled /= gamma_max - gamma_min
        gamma_scaled[gamma_scaled < 0] = 0
        gamma_scaled[gamma_scaled > 1] = 1
        gamma_scaled **= 1 / gamma
        recon = (1 - h) * recon + h * gamma_scaled
    return recon, planes, noise


# ---------------------------------------------------------------------------------------------------------------
# Monte-Carlo noise weights  (watroo/wavelets.py:221-229)
# ---------------------------------------------------------------------------------------------------------------
def compute_noise_weights(name: str, n_scales: int, n_trials: int = 100, bilateral=None, fields=None,
                          backend: str | None = None) -> np.ndarray:
    """Mean over trials of the per-plane population std of planes 0..n-1 of the transform of fp32 N(0,1) noise of
    side len(sigma_e_1d)*2**n_scales.  ``fields`` (an iterable of ready-made fp32 noise images) replaces the global
    NumPy RNG so that a test can feed identical fields to the GPU path."""
    std = np.zeros(n_scales)
    side = SIGMA_E_1D_LEN[name] * 2 ** n_scales
    it = iter(fields) if fields is not None else None
    for _ in range(n_trials):
        field = next(it) if it is not None else np.random.normal(size=(side, side)).astype(np.float32)
        planes = atrous_transform(field, n_scales, name, bilateral=bilateral, backend=backend)
        std += planes[:-1].std(axis=(1, 2))
    return std / n_trials


# ---------------------------------------------------------------------------------------------------------------
# Synthetic inputs and the parity metric shared by tests and bench  
# ---------------------------------------------------------------------------------------------------------------
def filter2d_reflect(img: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """cv2.filter2D(img, -1, kernel, (-1,-1), 0, cv2.BORDER_REFLECT) restated: correlation, anchor at the kernel
    centre (k // 2), half-sample symmetric border, float64 accumulation, one rounding (utils.py:252-255, :283-286)."""
    kh, kw = kernel.shape
    h, w = img.shape
    a64 = img.astype(np.float64)
    out = np.zeros((h, w), dtype=np.float64)
    ys, xs = np.arange(h), np.arange(w)
    for i in range(kh):
        rows = a64[reflect_index(ys + i - kh // 2, h), :]
        for j in range(kw):
            out += float(kernel[i, j]) * rows[:, reflect_index(xs + j - kw // 2, w)]
    return out.astype(img.dtype)


def enhance(img: np.ndarray, name: str = "b3spline", weights=None, denoise=None, noise=None,
            soft_threshold: bool = True) -> np.ndarray:
    """utils.py:36-80 for one channel or three (channel first): transform with max(len(weights), len(denoise)) scales,
    Coefficients.denoise(denoise, weights=weights), sum of the planes.  ``weights`` / ``denoise`` are per-channel
    lists of lists for a 3-D input (scalars are broadcast like prepare_params, utils.py:10-33)."""
    def one(ch, wgt, dns, nz):
        wgt, dns = list(wgt), list(dns)
        wgt += [1] * (len(dns) - len(wgt))
        dns += [0] * (len(wgt) - len(dns))
        planes = atrous_transform(ch, len(wgt), name)
        if nz is None:
            nz = get_noise(planes, name)
        denoise_planes(planes, name, dns, weights=wgt, noise=nz, soft_threshold=soft_threshold)
        return planes.sum(axis=0)

    def norm(p, nd):
        if nd == 2:
            return [] if p is None else ([p] if not isinstance(p, list) else list(p))
        if not isinstance(p, list):
            return [[] if p is None else [p]] * nd
        return [norm(q, 2) for q in p]

    if img.ndim == 2:
        return one(img, norm(weights, 2), norm(denoise, 2), noise)
    w3, d3 = norm(weights, 3), norm(denoise, 3)
    return np.stack([one(img[c], w3[c], d3[c], None if noise is None else noise[c]) for c in range(3)])


def richardson_lucy(data: np.ndarray, psf: np.ndarray, iterations: int = 10, denoise_coefficients=(5, 2, 1),
                    threshold_type: str = "soft", uniform_init: bool = False, persistent_mrs: bool = True) -> np.ndarray:
    """utils.py:222-290 (fft=False route), B3spline, line by line in terms of the pieces above."""
    soft = threshold_type == "soft"
    level = len(denoise_coefficients)
    name = "b3spline"
    planes0 = atrous_transform(data, level, name)
    noise0 = None
    if uniform_init:
        psi = np.ones_like(data, np.float32)
        psi *= data.sum() / data.size
    else:
        noise0 = get_noise(planes0, name)
        denoise_planes(planes0, name, list(denoise_coefficients), noise=noise0, soft_threshold=soft)
        psi = np.sum(planes0, axis=0)
    mrs = (np.ones if soft else np.zeros)((level,) + data.shape)
    for iteration in range(iterations):
        phi = filter2d_reflect(psi.astype(data.dtype), psf[::-1, ::-1])
        res = data - phi
        rc = atrous_transform(res, level, name)
        nz = noise0 if noise0 is not None else get_noise(rc, name)
        for s, c in enumerate(denoise_coefficients):
            sig = significance(rc, name, c, s, nz, soft_threshold=soft)
            if not soft:
                if persistent_mrs:
                    mrs[s][sig] = 1
                else:
                    mrs[s] = sig
                rc[s] *= mrs[s]
            else:
                if persistent_mrs:
                    mrs[s] *= sig
                else:
                    mrs[s] = sig
                rc[s] *= mrs[s] ** (1 / (iteration + 1))
        res = np.sum(rc, axis=0)
        res += phi
        res /= phi
        conv = filter2d_reflect(res, psf)
        psi *= conv
    return psi


def solar_like(n: int, seed: int = 2, flux: float = 1.0, dtype=np.float32, m: int | None = None) -> np.ndarray:
    """Synthetic solar-like frame: limb-darkened disk + exponential off-limb corona + 30 Gaussian active regions
    + background, Poisson noise.  Values are integers, hence identical in fp32 and fp64."""
    m = n if m is None else m
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:n, 0:m].astype(np.float64)
    scale = min(n, m)
    r = np.hypot(x - m / 2, y - n / 2) / (0.4 * scale)
    img = np.where(r < 1, 2000 * (0.4 + 0.6 * np.sqrt(np.clip(1 - r ** 2, 0, None))),
                   800 * np.exp(-(np.clip(r, 1, None) - 1) / 0.15))
    for _ in range(30):
        cx, cy = rng.uniform(0.2 * m, 0.8 * m), rng.uniform(0.2 * n, 0.8 * n)
        sg = rng.uniform(3, 30) * scale / 1024
        amp = rng.uniform(500, 8000)
        img += amp * np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2 * sg ** 2))
    img = (img + 20) * flux
    return rng.poisson(img).astype(dtype)


def emax(new: np.ndarray, ref: np.ndarray) -> float:
    """Parity metric of SURVEY 8(d): max|new-ref| / max|ref| over one plane (never element-wise relative)."""
    ref64 = np.asarray(ref, dtype=np.float64)
    den = np.abs(ref64).max()
    num = np.abs(np.asarray(new, dtype=np.float64) - ref64).max()
    return float(num / den) if den > 0 else float(num)
