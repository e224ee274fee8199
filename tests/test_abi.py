"""CPU checks of the boundary: the C-ABI library loads, exports every symbol the header declares, rejects bad
arguments before touching the device, and the host-side planning logic matches the reference's."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from wavelets_b200.build import build_library
    build_library()
    from wavelets_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from wavelets_b200 import _lib
    header = open(os.path.join(ROOT, "include", "wavelets_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(wb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.wb_abi_version() == _lib.ABI_VERSION


def test_argument_validation_needs_no_device(lib):
    ok = ctypes.c_void_p(16)
    # bad dtype / taps / shape / scale / aliasing / pitch are rejected with negative codes, nothing is launched
    assert lib.wb_atrous_scale(ok, ok, ok, 1, 8, 8, 8, 0, 8, 0, 8, 0, 0, 5, 7, None) == -1
    assert lib.wb_atrous_scale(ok, ctypes.c_void_p(32), None, 1, 8, 8, 8, 0, 8, 0, 8, 0, 0, 4, 0, None) == -2
    assert lib.wb_atrous_scale(ok, ctypes.c_void_p(32), None, 1, 0, 8, 8, 0, 8, 0, 8, 0, 0, 5, 0, None) == -3
    assert lib.wb_atrous_scale(ok, ctypes.c_void_p(32), None, 1, 8, 8, 8, 0, 8, 0, 8, 0, 31, 5, 0, None) == -4
    assert lib.wb_atrous_scale(ok, ok, None, 1, 8, 8, 8, 0, 8, 0, 8, 0, 0, 5, 0, None) == -5
    assert lib.wb_atrous_scale(ok, ctypes.c_void_p(32), None, 1, 8, 8, 4, 0, 8, 0, 8, 0, 0, 5, 0, None) == -6
    assert lib.wb_atrous_transform(None, ok, ok, 1, 8, 8, 8, 0, 2, 5, 0, None) == -5
    assert lib.wb_wow_whiten_scale(ok, ok, 1, 8, 8, 8, 0, 8, 0, 0, 5, 0, 0, 0.0, 1.0, 0.0, None, 1.0, None) == -5
    assert lib.wb_abs_median(ok, 0, 1, 0, 0, ok, None, 1.0, ok, 1 << 20, None) == -3
    # the whole-cascade entry validates everything before its first launch
    arr = (ctypes.c_double * 3)(1.0, 1.0, 1.0)
    ws = lib.wb_wow_cascade_workspace_bytes(0, 1, 64)
    assert ws >= lib.wb_abs_median_workspace_bytes(0, 1, 64) + lib.wb_plane_moments_workspace_bytes(1)
    assert lib.wb_wow_cascade(ok, 8, 0, ok, ok, ok, 1, 8, 8, 2, 5, 7, arr, arr, arr, None, 1, 0.0, None, 0, ok, ws, None) == -1
    assert lib.wb_wow_cascade(ok, 8, 0, ok, ok, ok, 1, 8, 8, 0, 5, 0, arr, arr, arr, None, 1, 0.0, None, 0, ok, ws, None) == -4
    assert lib.wb_wow_cascade(ok, 8, 0, None, ok, ok, 1, 8, 8, 2, 5, 0, arr, arr, arr, None, 1, 0.0, None, 0, ok, ws, None) == -5
    assert lib.wb_wow_cascade(ok, 8, 0, ok, ok, ok, 1, 8, 8, 2, 5, 0, arr, arr, arr, None, 1, 0.0, None, 1, ok, ws, None) == -5
    assert lib.wb_wow_cascade(ok, 8, 0, ok, ok, ok, 1, 8, 8, 2, 5, 0, arr, arr, arr, None, 1, 0.0, None, 0, ok, 16, None) == -6
    assert b"dtype" in lib.wb_error_string(-1)
    assert lib.wb_abs_median_workspace_bytes(0, 2, 0) >= 2 * lib.wb_abs_median_workspace_bytes(0, 1, 0) - 512
    # + a quarter of the plane per frame for the filter pass
    assert lib.wb_abs_median_workspace_bytes(0, 2, 1 << 20) >= lib.wb_abs_median_workspace_bytes(0, 2, 0) + 2 * (1 << 20)


def test_kernel_path_selection(lib):
    # aligned 4096^2: TMA row pipeline at every scale of BASELINE cfg2 / cfg3; odd widths and tiny images: generic
    for s in range(10):
        assert lib.wb_atrous_scale_path(4096, 4096, 4096, 4096, s, 5, 0, None, None, None) == 1
        assert lib.wb_atrous_scale_path(4096, 4096, 4096, 4096, s, 5, 1, None, None, None) == 1
    assert lib.wb_atrous_scale_path(37, 53, 53, 53, 0, 5, 0, None, None, None) == 0
    assert lib.wb_atrous_scale_path(64, 64, 64, 64, 5, 5, 0, None, None, None) == 1      # 2*32 <= 64
    assert lib.wb_atrous_scale_path(64, 64, 64, 64, 6, 5, 0, None, None, None) == 0      # 2*64 > 64: multi-reflection
    assert lib.wb_atrous_scale_path(64, 64, 64, 64, 6, 3, 0, None, None, None) == 1      # Triangle: 64 <= 64


def test_scaling_function_surface():
    import wavelets_b200 as wb
    b3, tri = wb.B3spline(2), wb.Triangle(2)
    assert b3.name == "b3spline" and tri.name == "triangle" and b3.n_dim == 2
    assert np.allclose(b3.coefficients_1d, [1 / 16, 1 / 4, 3 / 8, 1 / 4, 1 / 16])
    assert np.allclose(b3.kernel, np.outer(b3.coefficients_1d, b3.coefficients_1d))
    k = tri.atrous_kernel(3)
    assert k.shape == (17, 17) and np.isclose(k.sum(), 1) and np.count_nonzero(k) == 9
    assert len(b3.sigma_e()) == 11 and len(b3.sigma_e(bilateral=1)) == 10 and len(tri.sigma_e(bilateral=1)) == 11
    with pytest.raises(ValueError, match="Unsupported number of dimensions"):
        wb.B3spline(4)
    # the tables are the reference's recorded constants (oracle holds an independent copy)
    from oracle import atrous_oracle as orc
    assert np.array_equal(b3.sigma_e(), orc.SIGMA_E_2D["b3spline"])
    assert np.array_equal(tri.sigma_e(bilateral=True), orc.SIGMA_E_2D_BILATERAL["triangle"])


def test_wow_plan_matches_reference_logic():
    import wavelets_b200 as wb
    from wavelets_b200.utils import _wow_plan
    n, sb, wts, dns = _wow_plan((4096, 4096), wb.B3spline, None, [], [5, 2], 1)
    assert n == 10 and sb == [1] * 11 and wts == [1] * 11 and dns == [5, 2] + [0] * 8 + [1]
    n, sb, wts, dns = _wow_plan((512, 512), wb.B3spline, 50, [0.5], [], None)
    assert n == 7 and sb is None and wts == [0.5] + [1] * 7
    n, sb, wts, dns = _wow_plan((512, 512), wb.Triangle, 3, [], [], [2, 3])
    assert n == 3 and sb == [2, 3, 1, 1]
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        n, _, _, _ = _wow_plan((64, 64), wb.B3spline, None, [], [0] * 10, 1)  # bilateral table has 10 entries
    assert n == 10 and any("lager" in str(w.message) for w in rec)


def test_no_cpu_fallback():
    import torch
    import wavelets_b200 as wb
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wb.AtrousTransform()(np.ones((16, 16), dtype=np.float32), 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wb.wow(np.ones((16, 16), dtype=np.float32))
