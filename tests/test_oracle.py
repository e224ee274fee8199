"""Pins oracle/atrous_oracle.py against vectors produced by the real reference (tests/golden/make_golden.py),
the reference's own known-answer test and the identities listed in SURVEY.md section 4.  CPU only."""
import warnings

import numpy as np
import pytest

from oracle import atrous_oracle as orc
from tests.conftest import load_golden

BACKENDS = ["numpy"] + (["cv2"] if orc.cv2 is not None else [])
TOL = {"float32": 1e-5, "float64": 1e-13}  # north_star E_max tolerances; the cv2 backend must be bit-identical


def check(new, ref, dt, backend, exact_with_cv2=True, scale=0.0):
    """Per plane: max|new-ref| <= max(TOL*max|ref_p|, 4*eps*scale).  The second term is the rounding floor of the
    fp32 data itself (w_s = c_s - c_{s+1} carries a few ulp of c_s, whatever the size of w_s)."""
    assert new.dtype == ref.dtype and new.shape == ref.shape
    if backend == "cv2" and exact_with_cv2:
        assert np.array_equal(new, ref)
    else:
        floor = 4 * np.finfo(ref.dtype).eps * scale
        for p in range(new.shape[0]) if new.ndim == 3 else [None]:
            a, b = (new, ref) if p is None else (new[p], ref[p])
            err = np.abs(a.astype(np.float64) - b).max()
            assert err <= max(TOL[dt] * np.abs(b).max(), floor), (p, err, np.abs(b).max())


def test_reflect_index_matches_np_pad_symmetric():
    for n in (1, 2, 5, 8):
        base = np.arange(n)
        padded = np.pad(base, (3 * n + 1, 3 * n + 2), mode="symmetric")
        idx = np.arange(-(3 * n + 1), n + 3 * n + 2)
        assert np.array_equal(orc.reflect_index(idx, n), padded)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("dt", ["float32", "float64"])
@pytest.mark.parametrize("sf", ["b3spline", "triangle"])
def test_transform_golden(sf, dt, backend):
    g = load_golden(f"transform_{sf}_{dt}")
    for k in range(int(g["n"])):
        img = g[f"in{k}"]
        keep = img.copy()
        out = orc.atrous_transform(img, int(g[f"level{k}"]), sf, backend=backend)
        assert np.array_equal(img, keep)  # input never modified (wavelets.py:427)
        check(out, g[f"out{k}"], dt, backend, scale=np.abs(img).max())
        # perfect reconstruction
        assert orc.emax(out.sum(axis=0), img) < (1e-5 if dt == "float32" else 1e-13)


def test_mirror_index_matches_np_pad_reflect():
    for n in (1, 2, 5, 8):
        base = np.arange(n)
        padded = np.pad(base, (3 * n + 1, 3 * n + 2), mode="reflect") if n > 1 else np.zeros(7 * n + 3, dtype=int)
        idx = np.arange(-(3 * n + 1), n + 3 * n + 2)
        assert np.array_equal(orc.mirror_index(idx, n), padded)


@pytest.mark.parametrize("dt", ["float32", "float64"])
@pytest.mark.parametrize("sf", ["b3spline", "triangle"])
def test_transform_nd_golden(sf, dt):
    """1-D signals (scipy 'mirror' border) and 3-D volumes (2-D smooth per slice + depth pass) of the real reference
    (wavelets.py:46-69), incl. MAD noise and soft denoise of the planes."""
    g = load_golden(f"transform_nd_{sf}_{dt}")
    for k in range(int(g["n"])):
        arr = g[f"in{k}"]
        level = int(g[f"level{k}"])
        out = orc.atrous_transform(arr, level, sf)
        ref = g[f"out{k}"]
        assert out.shape == ref.shape and out.dtype == ref.dtype
        floor = 4 * np.finfo(ref.dtype).eps * np.abs(arr).max()
        for p in range(level + 1):
            err = np.abs(out[p].astype(np.float64) - ref[p]).max()
            assert err <= max(TOL[dt] * np.abs(ref[p]).max(), floor), (k, p, err)
        assert orc.emax(out.sum(axis=0), arr) < (1e-5 if dt == "float32" else 1e-13)
        dn = orc.denoise(arr, [3, 2, 1][:level], sf)
        assert orc.emax(dn, g[f"dn{k}"]) < (2e-5 if dt == "float32" else 1e-12)


@pytest.mark.parametrize("backend", BACKENDS)
def test_integer_input_is_recast_to_float64(backend):
    g = load_golden("transform_int16")
    out = orc.atrous_transform(g["img"], 3, "b3spline", backend=backend)
    assert out.dtype == np.float64
    check(out, g["out"], "float64", backend)


@pytest.mark.parametrize("backend", BACKENDS)
def test_reference_kat_constant_image(backend):
    """tests/test_wavelets.py:8-13 of the reference: ones -> zero detail planes, unit residual."""
    out = orc.atrous_transform(np.ones((128, 128)), 4, "b3spline", backend=backend)
    expected = np.zeros(out.shape)
    expected[-1] = 1
    assert np.isclose(out, expected).all()
    assert np.isclose(load_golden("kat_ones")["out"], expected).all()


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("dt", ["float32", "float64"])
def test_bilateral_golden(dt, backend):
    g = load_golden(f"bilateral_{dt}")
    tol = 5e-5 if dt == "float32" else 1e-11  # exp() ulp differences get amplified by the fp32 variance
    cases = [("in_solar", "solar_b1", dict(level=4, name="b3spline", bilateral=1)),
             ("in_solar", "solar_tri_b2", dict(level=3, name="triangle", bilateral=2.0)),
             ("in_gauss", "gauss_list_scaling", dict(level=4, name="b3spline", bilateral=[2, 1.5],
                                                     bilateral_scaling=True))]
    for src, key, kw in cases:
        out = orc.atrous_transform(g[src], backend=backend, **kw)
        ref = g[key]
        assert out.dtype == ref.dtype
        for p in range(len(ref)):
            assert orc.emax(out[p], ref[p]) <= tol, (key, p, orc.emax(out[p], ref[p]))


def test_cfg1_readme_denoise():
    """README.md:37-61 (BASELINE configs[0]): Triangle, 512x512 np.random.normal, denoise([5, 3])."""
    g = load_golden("cfg1_denoise_triangle_512")
    np.random.seed(int(g["seed"]))
    img = np.random.normal(size=(512, 512))
    for backend in BACKENDS:
        planes = orc.atrous_transform(img, 2, "triangle", backend=backend)
        assert orc.emax(planes[:, ::4, ::4], g["raw_sub"]) < 1e-13
        noise = orc.get_noise(planes, "triangle")
        assert abs(noise - float(g["noise"])) <= 1e-13 * float(g["noise"])
        hard = orc.significance(planes, "triangle", 3, 1, noise, soft_threshold=False)
        assert hard.dtype == np.bool_ and int(hard.sum()) == int(g["hard_count"])
        assert np.array_equal(hard[::4, ::4], g["hard_sub"])
        den = orc.denoise(img, [5, 3], "triangle", backend=backend)
        assert orc.emax(den[::4, ::4], g["out_sub"]) < 1e-13
        assert abs(den.sum() - float(g["out_sum"])) < 1e-8
        # README equivalences: np.sum(coefficients, axis=0) == data.sum(axis=0) == denoise(...)
        orc.denoise_planes(planes, "triangle", [5, 3])
        assert np.array_equal(planes.sum(axis=0), den)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("dt", ["float32", "float64"])
def test_denoise_and_significance_golden(dt, backend):
    g = load_golden(f"denoise_{dt}")
    img = g["img"]
    for sf in ("b3spline", "triangle"):
        planes = orc.atrous_transform(img, 3, sf, backend=backend)
        noise = orc.get_noise(planes, sf)
        assert isinstance(noise, np.float64)  # NEP-50 promotion
        assert abs(noise - float(g[f"{sf}_noise"])) <= (2e-6 if dt == "float32" else 1e-13) * noise
        soft = orc.significance(planes, sf, 2.5, 2, noise)
        assert soft.dtype == np.float64 == g[f"{sf}_soft2"].dtype
        assert np.abs(soft - g[f"{sf}_soft2"]).max() < (1e-4 if dt == "float32" else 1e-11)
        hard = orc.significance(planes, sf, 3, 0, noise, soft_threshold=False)
        assert hard.dtype == np.bool_
        if backend == "cv2":
            assert np.array_equal(hard, g[f"{sf}_hard0"])
        else:  # a 1-ulp difference in w_0 may flip a pixel sitting on the threshold
            assert (hard != g[f"{sf}_hard0"]).mean() < 1e-3
        for key, kw in (("den_soft", dict(weights=[5, 3, 2])),
                        ("den_hard", dict(weights=[4, 3, 0], soft_threshold=False)),
                        ("den_noise", dict(weights=[3, 2], noise=0.8)),
                        ("den_bilateral", dict(weights=[3, 2], bilateral=1))):
            out = orc.denoise(img, name=sf, backend=backend, **kw)
            ref = g[f"{sf}_{key}"]
            assert out.dtype == ref.dtype
            if key == "den_hard" and backend != "cv2":
                assert (np.abs(out - ref) > 1e-5 * np.abs(ref).max()).mean() < 1e-3
            else:
                tol = (5e-5 if key == "den_bilateral" else 2e-6) if dt == "float32" else 1e-11
                assert orc.emax(out, ref) <= tol, (sf, key, orc.emax(out, ref))


WOW_CASES = {
    "default": {},
    "den": dict(denoise_coefficients=[5, 2]),
    "den_hard": dict(denoise_coefficients=[5, 2], soft_threshold=False),
    "bil": dict(bilateral=1),
    "bil_den": dict(bilateral=1, denoise_coefficients=[5, 2]),
    "weights": dict(weights=[0.5, 2.0, 1.5], n_scales=3),
    "nowhite": dict(whitening=False, denoise_coefficients=[3, 1], noise=1.3),
    "tri": dict(name="triangle", denoise_coefficients=[0, 3]),
}


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("dt", ["float32", "float64"])
def test_wow_golden(dt, backend):
    g = load_golden(f"wow_{dt}")
    for tag in ("gauss", "solar", "rect"):
        img = g[f"{tag}_in"]
        for key, kw in WOW_CASES.items():
            if f"{tag}_{key}_recon" not in g:
                continue
            recon, planes, noise = orc.wow(img.copy(), backend=backend, **kw)
            ref = g[f"{tag}_{key}_recon"]
            assert recon.dtype == ref.dtype and recon.shape == ref.shape
            bil = "bil" in key
            if dt == "float64":
                tol = 1e-9 if bil else 1e-11
            else:
                tol = 2e-3 if bil else 2e-5
            if key == "den_hard" and backend != "cv2":
                assert (np.abs(recon - ref) > tol * np.abs(ref).max()).mean() < 2e-3
            else:
                assert orc.emax(recon, ref) <= tol, (tag, key, orc.emax(recon, ref))
            ref_noise = float(g[f"{tag}_{key}_noise"])
            if np.isnan(ref_noise):
                assert noise is None
            else:
                assert abs(noise - ref_noise) <= (1e-4 if dt == "float32" else 1e-11) * abs(ref_noise)
            if f"{tag}_{key}_planes" in g and not (key == "den_hard" and backend != "cv2"):
                rp = g[f"{tag}_{key}_planes"]
                assert planes.shape == rp.shape
                for p in range(len(rp)):
                    assert orc.emax(planes[p], rp[p]) <= tol * 10, (tag, key, p, orc.emax(planes[p], rp[p]))


WOW_OPTION_CASES = {
    "gamma": dict(h=0.4, denoise_coefficients=[5, 2], gamma=2.5),
    "gamma_one": dict(h=1, denoise_coefficients=[3, 2, 1]),
    "gamma_range": dict(h=0.3, gamma_min=10.0, gamma_max=60.0, n_scales=3),
    "pv": dict(preserve_variance=True),
    "pv_den_gamma": dict(preserve_variance=True, h=0.25, denoise_coefficients=[4, 2], weights=[1.5, 0.5]),
}


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("dt", ["float32", "float64"])
def test_wow_options_golden(dt, backend):
    """gamma blend (h > 0) and preserve_variance (utils.py:157-158, :178-184, :207-217) against the real reference."""
    g = load_golden(f"wow_options_{dt}")
    for tag in ("gauss", "solar"):
        img = g[f"{tag}_in"]
        for key, kw in WOW_OPTION_CASES.items():
            recon, planes, _ = orc.wow(img.copy(), backend=backend, **kw)
            ref, rp = g[f"{tag}_{key}_recon"], g[f"{tag}_{key}_planes"]
            assert recon.dtype == ref.dtype and recon.shape == ref.shape and planes.shape == rp.shape
            tol = 1e-11 if dt == "float64" else 2e-5
            assert orc.emax(recon, ref) <= tol, (tag, key, orc.emax(recon, ref))
            for p in range(len(rp)):
                assert orc.emax(planes[p], rp[p]) <= tol * 10, (tag, key, p, orc.emax(planes[p], rp[p]))


def test_wow_scale_count_logic():
    assert orc.wow_default_scales((4096, 4096), "b3spline") == 10  # utils.py:122 at BASELINE's size
    assert orc.wow_default_scales((4096, 4096), "triangle") == 10
    assert orc.wow_default_scales((512, 512), "b3spline") == 7
    img = np.random.default_rng(0).standard_normal((64, 64))
    _, planes, _ = orc.wow(img, n_scales=50, backend="numpy")  # clipped to the default (utils.py:125-126)
    assert len(planes) == orc.wow_default_scales(img.shape, "b3spline") + 1
    with warnings.catch_warnings(record=True) as rec:  # utils.py:135-138
        warnings.simplefilter("always")
        _, planes, _ = orc.wow(np.ones((16, 16)) + img[:16, :16], denoise_coefficients=[0] * 11, noise=1.0,
                               backend="numpy")
    assert len(planes) == 12 and any("lager" in str(w.message) for w in rec)


@pytest.mark.parametrize("sf", ["b3spline", "triangle"])
def test_noise_weights_golden_and_tables(sf):
    g = load_golden(f"noise_weights_{sf}")
    out = orc.compute_noise_weights(sf, 3, n_trials=2, fields=g["fields"])
    assert np.abs(out / g["out"] - 1).max() < 1e-6
    # statistical KAT: the recorded sigma_e_2d table (wavelets.py:245-247, :274-276), 2 trials of 88x88
    assert np.abs(out / orc.SIGMA_E_2D[sf][:3] - 1).max() < 0.08


def test_backends_agree_fp64():
    if orc.cv2 is None:
        pytest.skip("OpenCV not installed")
    img = orc.solar_like(96, seed=1, flux=1.0, dtype=np.float64)
    a = orc.atrous_transform(img, 5, "b3spline", backend="numpy")
    b = orc.atrous_transform(img, 5, "b3spline", backend="cv2")
    for p in range(6):
        assert orc.emax(a[p], b[p]) < 1e-13


def test_enhance_golden():
    """enhance (utils.py:36-80): single channel (estimated and given noise, soft and hard) and three channels."""
    g = load_golden("enhance")
    img = g["img"]
    a = orc.enhance(img, "b3spline", weights=[1.5, 1.2, 1.0], denoise=[3, 2])
    assert orc.emax(a, g["enh_a"]) < 2e-5
    b = orc.enhance(img, "triangle", weights=[2.0], denoise=[4, 2, 1], noise=np.float64(2.0), soft_threshold=False)
    assert (np.abs(b - g["enh_b"]) > 1e-4 * np.abs(g["enh_b"]).max()).mean() < 1e-3  # hard mask: rare threshold flips
    rgb = orc.enhance(g["rgb"], "b3spline", weights=[[1.2, 1.1], [1.0], [1.5, 1.0, 1.0]], denoise=[[3], [4, 2], [2]])
    assert orc.emax(rgb, g["enh_rgb"]) < 1e-12


@pytest.mark.parametrize("dt", ["float32", "float64"])
def test_richardson_lucy_golden(dt):
    """richardson_lucy (utils.py:222-290), direct-filter route: soft / hard support, persistent or not, uniform init."""
    g = load_golden("richardson_lucy")
    data, psf = g[f"data_{dt}"], g[f"psf_{dt}"]
    tol = 2e-4 if dt == "float32" else 1e-9
    assert orc.emax(orc.richardson_lucy(data, psf, iterations=4), g[f"soft_{dt}"]) < tol
    assert orc.emax(orc.richardson_lucy(data, psf, iterations=3, persistent_mrs=False), g[f"soft_np_{dt}"]) < tol
    hard = orc.richardson_lucy(data, psf, iterations=3, denoise_coefficients=(4, 2), threshold_type="hard")
    ref = g[f"hard_{dt}"]
    assert (np.abs(hard - ref) > tol * np.abs(ref).max()).mean() < 2e-3  # a flipped mask pixel spreads over one PSF
    if dt == "float32":
        uni = orc.richardson_lucy(data, psf, iterations=3, uniform_init=True)
        assert uni.dtype == np.float32 and orc.emax(uni, g["uniform_float32"]) < tol


@pytest.mark.parametrize("sf", ["b3spline", "triangle"])
def test_transform_recursive_golden(sf):
    """recursive=True (wavelets.py:330-406) restated as a per-sub-lattice border rule: equals the reference's recursion
    and differs from the standard algorithm near the borders."""
    g = load_golden(f"transform_recursive_{sf}")
    for k in range(int(g["n"])):
        img, level, ref = g[f"in{k}"], int(g[f"level{k}"]), g[f"out{k}"]
        out = orc.atrous_transform_recursive(img, level, sf)
        assert out.dtype == ref.dtype and out.shape == ref.shape
        dt = str(ref.dtype)
        floor = 4 * np.finfo(ref.dtype).eps * np.abs(img).max()
        for p in range(level + 1):
            err = np.abs(out[p].astype(np.float64) - ref[p]).max()
            assert err <= max(TOL[dt] * np.abs(ref[p]).max(), floor), (k, p, err)
        std = orc.atrous_transform(img, level, sf)
        assert max(orc.emax(std[p], ref[p]) for p in range(level + 1)) > 1e-2  # the two algorithms really differ


def test_oracle_f2_noise_weights_bilateral_and_anscombe():
    """SURVEY 8(f) rank 2 against the real reference (tests/golden/make_golden.py --f2): compute_noise_weights with a
    bilateral cascade on stored noise fields, denoise(anscombe=True), generalized_anscombe and its inverse."""
    g = load_golden("f2_noise_weights_anscombe")
    for sf in ("b3spline", "triangle"):
        for key, bil in (("nw_bil1", 1), ("nw_bil2", 2.5)):
            out = orc.compute_noise_weights(sf, 3, n_trials=2, bilateral=bil, fields=g[f"{sf}_fields"], backend="cv2")
            assert np.abs(out / g[f"{sf}_{key}"] - 1).max() < 2e-5, (sf, key, out, g[f"{sf}_{key}"])
    for dt in ("float32", "float64"):
        img = g[f"ans_in_{dt}"]
        tol = 2e-5 if dt == "float32" else 1e-12
        assert orc.emax(orc.denoise(img.copy(), [4, 2, 1], "b3spline", anscombe=True, backend="cv2"), g[f"ans_soft_{dt}"]) < tol
        assert orc.emax(orc.denoise(img.copy(), [3, 2], "b3spline", anscombe=True, bilateral=1, backend="cv2"),
                        g[f"ans_bil_{dt}"]) < (1e-4 if dt == "float32" else 1e-11)
        hard = orc.denoise(img.copy(), [3, 2], "triangle", anscombe=True, soft_threshold=False, noise=1.0, backend="cv2")
        assert (np.abs(hard - g[f"ans_tri_hard_{dt}"]) > 1e-5 * np.abs(hard).max()).mean() < 1e-3
        fwd = orc.generalized_anscombe(img.copy(), alpha=2.0, g=1.5, sigma=0.7)
        assert np.array_equal(fwd, g[f"ans_fwd_{dt}"])
        assert np.array_equal(orc.generalized_anscombe(fwd.copy(), alpha=2.0, g=1.5, sigma=0.7, inverse=True), g[f"ans_inv_{dt}"])


@pytest.mark.parametrize("dt", ["float32", "float64"])
def test_oracle_wide_golden(dt):
    """The oracle against the real reference at benchmark widths (96x4096, 72x1536; make_golden.py --wide): the cv2
    backend is bit-identical on the plain path, within summation-order rounding on the bilateral / WOW paths, and the
    separable numpy backend (what the GPU tests use) agrees with it."""
    g = load_golden(f"wide_{dt}")
    npdt = np.dtype(dt).type
    for tag in ("a", "b"):
        h, w, seed = (int(v) for v in g[f"{tag}_shape"])
        img = orc.solar_like(h, seed=seed, flux=float(g[f"{tag}_flux"]), dtype=npdt, m=w)
        assert img.astype(np.float64).sum() == float(g[f"{tag}_in_sum"]) and np.array_equal(img[:2], g[f"{tag}_in_rows"])
        cols = g[f"{tag}_cols"]
        plain = orc.atrous_transform(img, 5, "b3spline", backend="cv2")[:, :, cols]
        assert np.array_equal(plain, g[f"{tag}_plain"])
        sep = orc.atrous_transform(img.astype(np.float64), 5, "b3spline", backend="numpy")[:, :, cols]
        for p in range(6):
            assert orc.emax(sep[p], g[f"{tag}_plain"][p]) < (1e-5 if dt == "float32" else 1e-13), (tag, p)
        if dt == "float64":
            bil = orc.atrous_transform(img, 5, "b3spline", bilateral=1, backend="cv2")[:, :, cols]
            for p in range(6):
                assert orc.emax(bil[p], g[f"{tag}_bil"][p]) < 1e-12, (tag, p)
            for key, kw in (("default", {}), ("bil_den", dict(bilateral=1, denoise_coefficients=[5, 2]))):
                recon, planes, noise = orc.wow(img.copy(), backend="cv2", **kw)
                # bilateral WOW: the 24-term weighted sum is accumulated in a different order than the reference's
                # in-place loop -- measured 1.1e-12 (96x4096, flux 0.05) and 1.2e-11 (72x1536, flux 1: S[x^2] - S[x]^2
                # cancels ~8 digits of float64 on counts of 1e4), which is the floor any "1e-12" bilateral claim has
                assert orc.emax(recon[:, cols], g[f"{tag}_{key}_recon"]) < (1e-10 if kw else 1e-12), (tag, key)
                if noise is not None:
                    assert abs(noise - float(g[f"{tag}_{key}_noise"])) <= 1e-13 * noise


ND_BILATERAL_CASES = [("b3spline", dict(bilateral=1)), ("triangle", dict(bilateral=[2, 1.5], bilateral_scaling=True)),
                      ("b3spline", dict(bilateral=1)), ("triangle", dict(bilateral=1.5))]


@pytest.mark.parametrize("dt", ["float32", "float64"])
def test_oracle_nd_bilateral_golden(dt):
    """Bilateral cascade of 1-D signals and 3-D volumes against the real reference (make_golden.py --nd-bilateral)."""
    g = load_golden(f"transform_nd_bilateral_{dt}")
    for k, (sf, kw) in enumerate(ND_BILATERAL_CASES):
        arr, level = g[f"in{k}"], int(g[f"level{k}"])
        out = orc.atrous_transform(arr, level, sf, backend="cv2" if arr.ndim == 3 else None, **kw)
        for p in range(level + 1):
            # fp32: the reference's variance S[x^2] - S[x]^2 cancels in float32 and the shim's exp is NumPy's
            assert orc.emax(out[p], g[f"out{k}"][p]) < (5e-5 if dt == "float32" else 1e-12), (k, p)
        if arr.ndim == 3:
            assert abs(orc.get_noise(out, sf, kw["bilateral"]) / float(g[f"noise{k}"]) - 1) < (5e-6 if dt == "float32" else 1e-13)
            dn = orc.denoise(arr.copy(), [3, 2][:level], sf, bilateral=kw["bilateral"], backend="cv2")
            assert orc.emax(dn, g[f"dn{k}"]) < (5e-6 if dt == "float32" else 1e-13)
    with pytest.raises(AttributeError):  # the reference has no sigma_e_1d_bilateral table
        orc.sigma_e("b3spline", bilateral=1, ndim=1)


WOW_ND_CASES = [{}, dict(denoise_coefficients=[4, 2], weights=[1.5, 1.0, 0.5]),
                dict(scaling_function="triangle", h=0.3, preserve_variance=True, denoise_coefficients=[3], noise=1.2),
                {}, dict(denoise_coefficients=[4, 2], soft_threshold=False),
                dict(bilateral=1, denoise_coefficients=[3, 1], n_scales=2)]


@pytest.mark.parametrize("dt", ["float32", "float64"])
def test_oracle_wow_nd_golden(dt):
    """wow() on 1-D signals and 3-D volumes against the real reference (make_golden.py --wow-nd)."""
    g = load_golden(f"wow_nd_{dt}")
    for k, kw in enumerate(WOW_ND_CASES):
        kw = dict(kw)
        name = kw.pop("scaling_function", "b3spline")
        recon, planes, noise = orc.wow(g[f"in{k}"].copy(), name=name, backend="cv2", **kw)
        bil = "bilateral" in kw
        assert planes.shape == g[f"planes{k}"].shape
        if kw.get("soft_threshold", True):
            # fp32 volumes: the oracle's n-D smooth rounds once, the reference's slice-wise cv2 passes round twice
            assert orc.emax(recon, g[f"recon{k}"]) < ((2e-5 if bil else 1e-5) if dt == "float32" else 1e-12), k
            for p in range(len(planes)):
                assert orc.emax(planes[p], g[f"planes{k}"][p]) < ((2e-4 if bil else 5e-5) if dt == "float32" else 1e-12), (k, p)
        else:
            assert (np.abs(recon - g[f"recon{k}"]) > 1e-5 * np.abs(recon).max()).mean() < 1e-3


RECURSIVE_BILATERAL_CASES = [("b3spline", dict(bilateral=1)), ("triangle", dict(bilateral=[2, 1.5], bilateral_scaling=True)),
                             ("b3spline", dict(bilateral=1.5))]


def test_oracle_recursive_bilateral_golden():
    """recursive=True with a bilateral cascade (wavelets.py:371-378) against the real reference."""
    g = load_golden("transform_recursive_bilateral")
    for k, (sf, kw) in enumerate(RECURSIVE_BILATERAL_CASES):
        ref = g[f"out{k}"]
        out = orc.atrous_transform_recursive(g[f"in{k}"], int(g[f"level{k}"]), sf, **kw)
        for p in range(len(ref)):
            assert orc.emax(out[p], ref[p]) < (5e-5 if ref.dtype == np.float32 else 1e-12), (k, p)
