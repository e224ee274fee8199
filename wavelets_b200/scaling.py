"""Scaling functions of the à trous transform: taps, 2-D kernels and noise-normalisation tables.

Mirror of the reference's ``AbstractScalingFunction`` / ``Triangle`` / ``B3spline`` (watroo/wavelets.py:152-287)
(2-D hot path, plus the 1-D / 3-D tables): same attribute and method names (``name``, ``n_dim``, ``kernel``, ``coefficients_1d``,
``coefficients_2d``, ``atrous_kernel(scale)``, ``sigma_e(bilateral)``, ``compute_noise_weights``).  The device
kernels never build the dense dilated kernel -- ``atrous_kernel`` exists for API compatibility only.
"""
from __future__ import annotations

import numpy as np

from . import _lib

__all__ = ["AbstractScalingFunction", "Triangle", "B3spline"]


class AbstractScalingFunction:
    """Base class; subclasses provide ``coefficients_1d`` and the ``sigma_e_*`` tables."""

    coefficients_1d = None
    taps_code = None          # WB_TRIANGLE / WB_B3SPLINE of the C ABI
    sigma_e_1d = None
    sigma_e_2d = None
    sigma_e_3d = None
    sigma_e_1d_bilateral = None
    sigma_e_2d_bilateral = None
    sigma_e_3d_bilateral = None

    def __init__(self, name, n_dim):
        if n_dim not in (1, 2, 3):
            raise ValueError("Unsupported number of dimensions")  # watroo/wavelets.py:189
        self.name = name
        self.n_dim = n_dim
        self.kernel = self.make_kernel()

    @property
    def coefficients_2d(self):
        return np.outer(self.coefficients_1d, self.coefficients_1d)

    @property
    def coefficients_3d(self):
        h = self.coefficients_1d
        return h[:, None, None] * h[None, :, None] * h[None, None, :]

    def make_kernel(self):
        return {1: self.coefficients_1d, 2: self.coefficients_2d, 3: self.coefficients_3d}[self.n_dim]

    def atrous_kernel(self, scale):
        """Dense kernel with 2**scale - 1 zeros ("trous") between taps (watroo/wavelets.py:191-197)."""
        step = 2 ** scale
        dense = np.zeros([(n - 1) * step + 1 for n in self.kernel.shape])
        dense[(slice(None, None, step),) * self.n_dim] = self.kernel
        return dense

    def sigma_e(self, bilateral=None):
        """Std of each wavelet plane for unit white noise (watroo/wavelets.py:199-219): the table of this
        dimensionality; the bilateral table whenever ``bilateral is not None`` (None where the reference has none)."""
        plain = {1: self.sigma_e_1d, 2: self.sigma_e_2d, 3: self.sigma_e_3d}
        bil = {1: self.sigma_e_1d_bilateral, 2: self.sigma_e_2d_bilateral, 3: self.sigma_e_3d_bilateral}
        table = (plain if bilateral is None else bil)[self.n_dim]
        if table is None:  # the reference has no such attribute (e.g. sigma_e_1d_bilateral): same exception type
            raise AttributeError(f"'{type(self).__name__}' object has no attribute "
                                 f"'sigma_e_{self.n_dim}d{'_bilateral' if bilateral is not None else ''}'")
        return table

    def compute_noise_weights(self, n_scales, n_trials=100, bilateral=None, fields=None, seed=None):
        """Monte-Carlo estimate of ``sigma_e`` (watroo/wavelets.py:221-229), entirely on the GPU.

        Each trial transforms an fp32 N(0,1) field of side ``len(sigma_e_1d) * 2**n_scales`` and takes the
        population std of planes 0..n_scales-1; the result is the mean over trials (float64 ndarray).
        ``fields`` (iterable of ready-made fp32 images, host or device) replaces the device RNG -- a test hook to
        feed the very same noise to the reference; ``seed`` seeds the device generator."""
        from .wavelets import noise_weights  # local import: wavelets.py imports this module
        return noise_weights(self, n_scales, n_trials=n_trials, bilateral=bilateral, fields=fields, seed=seed)


class Triangle(AbstractScalingFunction):
    """Triangle scaling function, 3 taps [1/4, 1/2, 1/4] (watroo/wavelets.py:232-258)."""

    coefficients_1d = np.array([1 / 4, 1 / 2, 1 / 4])
    taps_code = _lib.WB_TRIANGLE
    sigma_e_1d = np.array([0.60840933, 0.33000059, 0.21157957, 0.145824, 0.10158388, 0.07155912, 0.04902655,
                           0.03529812, 0.02409187, 0.01722846, 0.01144442])
    sigma_e_2d = np.array([0.7999247, 0.27308452, 0.11998217, 0.05793947, 0.0288104, 0.01447795, 0.00733832,
                           0.0037203, 0.00192882, 0.00098568, 0.00048533])
    sigma_e_3d = np.array([0.89736751, 0.19514386, 0.06239262, 0.02311278, 0.00939645])
    sigma_e_3d_bilateral = np.array([0.3828863, 0.36182913, 0.19520299, 0.08498861, 0.03363142])
    sigma_e_2d_bilateral = np.array([0.31063172, 0.34575647, 0.23712331, 0.13559906, 0.07172004, 0.03665405,
                                     0.01850046, 0.00928768, 0.00465967, 0.00234445, 0.00119249])

    def __init__(self, n_dim=2):
        super().__init__("triangle", n_dim)


class B3spline(AbstractScalingFunction):
    """B3-spline scaling function, 5 taps [1/16, 1/4, 3/8, 1/4, 1/16] (watroo/wavelets.py:261-287)."""

    coefficients_1d = np.array([1 / 16, 1 / 4, 3 / 8, 1 / 4, 1 / 16])
    taps_code = _lib.WB_B3SPLINE
    sigma_e_1d = np.array([0.72514976, 0.28538683, 0.17901161, 0.12222841, 0.08469601, 0.06027006, 0.04242257,
                           0.02919823, 0.01805671, 0.01383672, 0.00943623])
    sigma_e_2d = np.array([8.907e-01, 2.0072e-01, 8.5551e-02, 4.1261e-02, 2.0470e-02, 1.0232e-02, 5.1435e-03,
                           2.6008e-03, 1.3161e-03, 6.7359e-04, 4.0040e-04])
    sigma_e_3d = np.array([0.95633954, 0.12491933, 0.03933029, 0.01489642, 0.0064108])
    sigma_e_3d_bilateral = np.array([0.44111772, 0.3552894, 0.16137159, 0.05769064, 0.01932497])
    # NB: 10 entries, one fewer than Triangle's (watroo/wavelets.py:280-281)
    sigma_e_2d_bilateral = np.array([0.38234752, 0.24305799, 0.16012153, 0.10633541, 0.07083733, 0.04728659,
                                     0.03163678, 0.02122341, 0.01429102, 0.00952376])

    def __init__(self, n_dim=2):
        super().__init__("b3spline", n_dim)
