"""GPU parity of the callers beyond denoise / wow: enhance, richardson_lucy and the PSF
filter wb_filter2d, against golden vectors of the real reference and the oracle."""
import numpy as np
import pytest
import torch

from oracle import atrous_oracle as orc
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_filter2d_matches_oracle(dt):
    """wb_filter2d == cv2.filter2D(..., BORDER_REFLECT) restated by the oracle: odd / even / non-square kernels,
    kernels larger than the image (several reflections), flipped and not."""
    from wavelets_b200 import utils
    rng = np.random.default_rng(0)
    for (h, w), (kh, kw) in (((40, 56), (5, 5)), ((33, 17), (7, 3)), ((6, 9), (15, 21)), ((64, 64), (4, 6))):
        img = rng.standard_normal((h, w)).astype(dt)
        ker = rng.standard_normal((kh, kw)).astype(dt)
        dev = torch.from_numpy(img).cuda()
        kd = torch.from_numpy(ker).cuda()
        tol = 1e-5 if dt == np.float32 else 1e-12
        if kh % 2 and kw % 2:  # the oracle restates the centred anchor of odd kernels (all the reference uses)
            ref = orc.filter2d_reflect(img, ker)
            assert orc.emax(utils._filter2d(dev, kd, flip=False).cpu().numpy(), ref) < tol
            ref_f = orc.filter2d_reflect(img, ker[::-1, ::-1])
            assert orc.emax(utils._filter2d(dev, kd, flip=True).cpu().numpy(), ref_f) < tol
        else:
            got = utils._filter2d(dev, kd, flip=False).cpu().numpy()
            pad = np.pad(img.astype(np.float64), ((kh, kh), (kw, kw)), mode="symmetric")
            ref = np.zeros((h, w))
            for i in range(kh):
                for j in range(kw):
                    ref += ker[i, j] * pad[kh + i - kh // 2: kh + i - kh // 2 + h, kw + j - kw // 2: kw + j - kw // 2 + w]
            assert orc.emax(got, ref.astype(dt)) < tol


def test_enhance_golden():
    import wavelets_b200 as wb
    g = load_golden("enhance")
    img = g["img"]
    a = wb.enhance(img.copy(), weights=[1.5, 1.2, 1.0], denoise=[3, 2], scaling_function_class=wb.B3spline)
    assert isinstance(a, np.ndarray) and a.dtype == np.float32 and orc.emax(a, g["enh_a"]) < 2e-5
    b = wb.enhance(img.copy(), np.float64(2.0), weights=[2.0], denoise=[4, 2, 1], soft_threshold=False,
                   scaling_function_class=wb.Triangle)
    assert (np.abs(b - g["enh_b"]) > 1e-4 * np.abs(g["enh_b"]).max()).mean() < 1e-3
    out = np.empty_like(g["rgb"])
    r = wb.enhance(g["rgb"].copy(), weights=[[1.2, 1.1], [1.0], [1.5, 1.0, 1.0]], denoise=[[3], [4, 2], [2]], out=out)
    assert r is out and orc.emax(out, g["enh_rgb"]) < 1e-11
    with pytest.raises(ValueError, match="Invalid number of parameters"):
        wb.enhance(g["rgb"], weights=[[1], [1]])


@pytest.mark.parametrize("dt", ["float32", "float64"])
def test_richardson_lucy_golden(dt):
    import wavelets_b200 as wb
    g = load_golden("richardson_lucy")
    data, psf = g[f"data_{dt}"], g[f"psf_{dt}"]
    tol = 3e-4 if dt == "float32" else 1e-9
    soft = wb.richardson_lucy(data.copy(), psf, iterations=4)
    assert isinstance(soft, np.ndarray) and soft.dtype == data.dtype and orc.emax(soft, g[f"soft_{dt}"]) < tol
    nonp = wb.richardson_lucy(data.copy(), psf, iterations=3, persistent_mrs=False)
    assert orc.emax(nonp, g[f"soft_np_{dt}"]) < tol
    hard = wb.richardson_lucy(data.copy(), psf, iterations=3, denoise_coefficients=(4, 2), threshold_type="hard")
    ref = g[f"hard_{dt}"]
    assert (np.abs(hard - ref) > tol * np.abs(ref).max()).mean() < 5e-3
    if dt == "float32":
        uni = wb.richardson_lucy(data.copy(), psf, iterations=3, uniform_init=True)
        assert uni.dtype == np.float32 and orc.emax(uni, g["uniform_float32"]) < tol
    else:
        even = data[:72, :88]
        fft = wb.richardson_lucy(even.copy(), psf, iterations=3, fft=True)
        assert orc.emax(fft, g["fft_float64"]) < 1e-9
    # device tensor in -> device tensor out
    t = wb.richardson_lucy(torch.from_numpy(data).cuda(), psf, iterations=1)
    assert isinstance(t, torch.Tensor) and t.is_cuda
