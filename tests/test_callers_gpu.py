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


def test_f2_noise_weights_bilateral_and_anscombe_golden():
    """SURVEY 8(f) rank 2 against the real reference (tests/golden/make_golden.py --f2): compute_noise_weights through
    the bilateral cascade fed with the reference's own noise fields; denoise(anscombe=True) (soft, hard + explicit
    noise, bilateral); generalized_anscombe forward and inverse."""
    import wavelets_b200 as wb
    from wavelets_b200.utils import generalized_anscombe
    g = load_golden("f2_noise_weights_anscombe")
    for name, sf in (("b3spline", wb.B3spline), ("triangle", wb.Triangle)):
        for key, bil in (("nw_bil1", 1), ("nw_bil2", 2.5)):
            out = sf(2).compute_noise_weights(3, n_trials=2, bilateral=bil, fields=g[f"{name}_fields"])
            assert isinstance(out, np.ndarray) and out.dtype == np.float64 and out.shape == (3,)
            # fp32 fields: the reference's own fp32 variance (S[x^2] - S[x]^2) is noisier than the kernel's centred form
            assert np.abs(out / g[f"{name}_{key}"] - 1).max() < 2e-5, (name, key, out / g[f"{name}_{key}"])
    for dt in ("float32", "float64"):
        img = g[f"ans_in_{dt}"]
        img64 = img.astype(np.float64)
        cases = (("ans_soft", "b3spline", wb.B3spline, dict(weights=[4, 2, 1])),
                 ("ans_bil", "b3spline", wb.B3spline, dict(weights=[3, 2], bilateral=1)))
        for key, name, sf, kw in cases:
            out = wb.denoise(img, scaling_function=sf, anscombe=True, **kw)
            ref = g[f"{key}_{dt}"]
            assert isinstance(out, np.ndarray) and out.dtype == ref.dtype and out.shape == ref.shape
            ref64 = orc.denoise(img64, name=name, anscombe=True, backend="numpy", **kw)
            tol = max(1e-5, 2 * orc.emax(ref, ref64)) if dt == "float32" else (1e-10 if "bilateral" in kw else 1e-12)
            assert orc.emax(out, ref64) <= tol, (key, dt, orc.emax(out, ref64), tol)
        hard = wb.denoise(img, [3, 2], wb.Triangle, anscombe=True, soft_threshold=False, noise=1.0)
        ref = g[f"ans_tri_hard_{dt}"]
        assert (np.abs(hard - ref) > 1e-5 * np.abs(ref).max()).mean() < (2e-3 if dt == "float32" else 1e-9)
        dev = torch.from_numpy(img).cuda()
        fwd = generalized_anscombe(dev, alpha=2.0, g=1.5, sigma=0.7)
        assert orc.emax(fwd.cpu().numpy(), g[f"ans_fwd_{dt}"]) < (2e-7 if dt == "float32" else 1e-15)
        inv = generalized_anscombe(fwd, alpha=2.0, g=1.5, sigma=0.7, inverse=True)
        assert orc.emax(inv.cpu().numpy(), g[f"ans_inv_{dt}"]) < (5e-7 if dt == "float32" else 1e-15)


def test_convolution_export_numpy_and_nd():
    """the exported convolution() mirrors the reference's (arr, sf, s, output): NumPy in -> NumPy out, an
    ndarray `output` is filled and returned, 1-D and 3-D inputs take the reference's borders."""
    import wavelets_b200 as wb
    rng = np.random.default_rng(2)
    img = rng.standard_normal((40, 64)).astype(np.float32)
    out = wb.convolution(img, wb.B3spline(2), s=2)
    assert isinstance(out, np.ndarray) and out.dtype == np.float32
    ref = orc.smooth(img.astype(np.float64), "b3spline", 2, backend="numpy")
    assert orc.emax(out, ref) < 1e-6
    buf = np.zeros_like(img)
    assert wb.convolution(img, wb.B3spline(2), s=2, output=buf) is buf and np.array_equal(buf, out)
    dev_out = wb.convolution(torch.from_numpy(img).cuda(), wb.B3spline(2), s=2)
    assert isinstance(dev_out, torch.Tensor) and np.array_equal(dev_out.cpu().numpy(), out)
    sig = rng.standard_normal(257)
    assert orc.emax(wb.convolution(sig, wb.Triangle(1), s=3), orc.smooth(sig, "triangle", 3)) < 1e-14
    vol = rng.standard_normal((6, 20, 24))
    assert orc.emax(wb.convolution(vol, wb.B3spline(3), s=1), orc.smooth(vol, "b3spline", 1)) < 1e-14
