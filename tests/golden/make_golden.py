"""Generate tests/golden/*.npz by running the REAL reference (watroo 0.0.4 under /root/reference).

Run here (the build container), never on the GPU box:   python tests/golden/make_golden.py
Each .npz stores the exact inputs and the reference outputs, so the GPU tests need neither /root/reference nor a
matching RNG.  numexpr is not installed in this image; its single call site is served by _numexpr_shim.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("WATROO_REFERENCE", "/root/reference")
try:
    import numexpr  # noqa: F401
except ImportError:
    sys.path.insert(0, os.path.join(HERE, "_numexpr_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import watroo  # noqa: E402
from watroo import AtrousTransform, B3spline, Triangle, denoise, wow  # noqa: E402
from oracle.atrous_oracle import solar_like  # noqa: E402  (input generator only)

SF = {"b3spline": B3spline, "triangle": Triangle}


def gaussian(shape, seed, dtype):
    return np.random.default_rng(seed).standard_normal(shape).astype(dtype)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name:36s} {os.path.getsize(path) / 1024:8.1f} KiB")


def main_nd():
    """1-D signals and 3-D volumes (wavelets.py:46-69): plain transform + denoise of the planes."""
    for sf in SF:
        for dt in ("float32", "float64"):
            out = {}
            cases = [((257,), 5), ((64,), 3), ((9,), 2), ((12, 20, 28), 3), ((5, 33, 16), 2)]
            for k, (shape, level) in enumerate(cases):
                arr = gaussian(shape, 20 + k, dt) * 3 + 10
                co = AtrousTransform(SF[sf])(arr, level)
                out[f"in{k}"] = arr
                out[f"out{k}"] = co.data.copy()
                out[f"level{k}"] = np.int64(level)
                out[f"noise{k}"] = np.float64(co.get_noise())
                co.denoise([3, 2][:level], soft_threshold=True)
                out[f"den{k}"] = co.data.copy()
                out[f"dn{k}"] = denoise(arr.copy(), [3, 2, 1][:level], scaling_function=SF[sf])  # utils.denoise, n-D
            save(f"transform_nd_{sf}_{dt}", n=np.int64(len(cases)), **out)


def main_recursive():
    """AtrousTransform(...)(img, level, recursive=True) (wavelets.py:330-406): differs from the standard planes near the
    borders."""
    for sf in SF:
        out = {}
        cases = [((64, 64), 4, "float64"), ((37, 53), 3, "float64"), ((96, 80), 5, "float32"), ((12, 9), 2, "float32")]
        for k, (shape, level, dt) in enumerate(cases):
            img = gaussian(shape, 40 + k, dt) * 3 + 10
            out[f"in{k}"] = img
            out[f"out{k}"] = AtrousTransform(SF[sf])(img.copy(), level, recursive=True).data
            out[f"level{k}"] = np.int64(level)
        save(f"transform_recursive_{sf}", n=np.int64(len(cases)), **out)
    # recursive algorithm of the BILATERAL cascade (wavelets.py:371-378)
    out = {}
    cases = [((64, 64), 3, "float64", "b3spline", dict(bilateral=1)),
             ((40, 56), 3, "float32", "triangle", dict(bilateral=[2, 1.5], bilateral_scaling=True)),
             ((37, 53), 2, "float64", "b3spline", dict(bilateral=1.5))]
    for k, (shape, level, dt, sf, kw) in enumerate(cases):
        img = gaussian(shape, 50 + k, dt) * 3 + 10
        out[f"in{k}"] = img
        out[f"out{k}"] = AtrousTransform(SF[sf], **kw)(img.copy(), level, recursive=True).data
        out[f"level{k}"] = np.int64(level)
    save("transform_recursive_bilateral", n=np.int64(len(cases)), **out)


def gaussian_psf(n, sigma):
    ax = np.arange(n) - n // 2
    g = np.exp(-(ax[:, None] ** 2 + ax[None, :] ** 2) / (2 * sigma ** 2))
    g[0, 1] *= 1.3  # not symmetric: tells convolution from correlation
    return g / g.sum()


def main_callers():
    """enhance (utils.py:36-80, not exported by the reference) and richardson_lucy (utils.py:222-290)."""
    from watroo.utils import enhance, richardson_lucy
    out = {}
    img = solar_like(96, seed=5, flux=0.05, dtype=np.float32, m=80)
    out["img"] = img
    out["enh_a"] = enhance(img.copy(), weights=[1.5, 1.2, 1.0], denoise=[3, 2], scaling_function_class=B3spline)
    out["enh_b"] = enhance(img.copy(), np.float64(2.0), weights=[2.0], denoise=[4, 2, 1], soft_threshold=False,
                           scaling_function_class=Triangle)
    rgb = np.stack([solar_like(64, seed=6 + c, flux=0.05, dtype=np.float64) for c in range(3)])
    out["rgb"] = rgb
    out["enh_rgb"] = enhance(rgb.copy(), weights=[[1.2, 1.1], [1.0], [1.5, 1.0, 1.0]], denoise=[[3], [4, 2], [2]])
    save("enhance", **out)

    out = {}
    rng = np.random.default_rng(11)
    truth = solar_like(72, seed=9, flux=0.05, dtype=np.float64, m=88) + 5.0
    for dt in ("float32", "float64"):
        psf = gaussian_psf(7, 1.4).astype(dt)
        import cv2
        blurred = cv2.filter2D(truth.astype(dt), -1, psf[::-1, ::-1].copy(), None, (-1, -1), 0, cv2.BORDER_REFLECT)
        data = (blurred + rng.standard_normal(blurred.shape) * 0.5).astype(dt)
        out[f"data_{dt}"] = data
        out[f"psf_{dt}"] = psf
        out[f"soft_{dt}"] = richardson_lucy(data.copy(), psf, iterations=4)
        out[f"hard_{dt}"] = richardson_lucy(data.copy(), psf, iterations=3, denoise_coefficients=(4, 2),
                                            threshold_type='hard')
        out[f"soft_np_{dt}"] = richardson_lucy(data.copy(), psf, iterations=3, persistent_mrs=False)
    out["uniform_float32"] = richardson_lucy(out["data_float32"].copy(), out["psf_float32"], iterations=3,
                                             uniform_init=True)
    even = out["data_float64"][:72, :88]
    out["fft_float64"] = richardson_lucy(even.copy(), out["psf_float64"], iterations=3, fft=True)
    save("richardson_lucy", **out)



def wide_columns(w):
    """Column subset stored by the wide fixtures: 4 columns either side of every 512-column strip boundary of the
    CUDA kernels, the first and last 8 columns, and every 64th column in between."""
    cols = set(range(0, 8)) | set(range(w - 8, w)) | set(range(0, w, 64))
    for b in range(512, w, 512):
        cols |= set(range(b - 4, b + 4))
    return np.array(sorted(c for c in cols if 0 <= c < w), dtype=np.int64)


def main_wide():
    """Frames at the widths the benchmark runs (>= 1024 columns, several 512-column strips): plain and bilateral
    transform, wow default / bilateral + denoise.  Inputs are regenerated from the seed by the tests (checked against
    the stored sum and a few rows); outputs are stored on `wide_columns` only to keep the fixtures small."""
    for dt in ("float32", "float64"):
        out = {}
        for tag, (h, w, seed, flux) in (("a", (96, 4096, 12, 0.05)), ("b", (72, 1536, 13, 1.0))):
            img = solar_like(h, seed=seed, flux=flux, dtype=dt, m=w)
            cols = wide_columns(w)
            out[f"{tag}_shape"] = np.array([h, w, seed], dtype=np.int64)
            out[f"{tag}_flux"] = np.float64(flux)
            out[f"{tag}_cols"] = cols
            out[f"{tag}_in_sum"] = np.float64(img.astype(np.float64).sum())
            out[f"{tag}_in_rows"] = img[:2].copy()
            out[f"{tag}_plain"] = AtrousTransform(B3spline)(img.copy(), 5).data[:, :, cols]
            out[f"{tag}_bil"] = AtrousTransform(B3spline, bilateral=1)(img.copy(), 5).data[:, :, cols]
            out[f"{tag}_tri_bil"] = AtrousTransform(Triangle, bilateral=[1.5, 1], bilateral_scaling=True)(img.copy(), 6).data[:, :, cols]
            for key, kw in (("default", {}), ("den", dict(denoise_coefficients=[5, 2])),
                            ("bil_den", dict(bilateral=1, denoise_coefficients=[5, 2]))):
                recon, co = wow(img.copy(), **kw)
                out[f"{tag}_{key}_recon"] = recon[:, cols]
                out[f"{tag}_{key}_planes"] = co.data[:, :, cols]
                out[f"{tag}_{key}_noise"] = np.float64(np.nan if co.noise is None else co.noise)
        save(f"wide_{dt}", **out)


def main_f2():
    """compute_noise_weights(bilateral=...) on stored noise fields and denoise(anscombe=True)."""
    import watroo.wavelets as ww
    ww.tqdm = lambda it: it  # silence the progress bar
    out = {}
    side = 11 * 2 ** 3
    for sf in SF:
        np.random.seed(8)
        state = np.random.get_state()
        out[f"{sf}_fields"] = np.stack([np.random.normal(size=(side, side)).astype(np.float32) for _ in range(2)])
        np.random.set_state(state)
        out[f"{sf}_nw_bil1"] = SF[sf](2).compute_noise_weights(3, n_trials=2, bilateral=1)
        np.random.set_state(state)
        out[f"{sf}_nw_bil2"] = SF[sf](2).compute_noise_weights(3, n_trials=2, bilateral=2.5)
    for dt in ("float32", "float64"):
        img = solar_like(96, seed=15, flux=0.05, dtype=dt, m=128)
        out[f"ans_in_{dt}"] = img
        out[f"ans_soft_{dt}"] = denoise(img.copy(), [4, 2, 1], B3spline, anscombe=True)
        out[f"ans_tri_hard_{dt}"] = denoise(img.copy(), [3, 2], Triangle, anscombe=True, soft_threshold=False, noise=1.0)
        out[f"ans_bil_{dt}"] = denoise(img.copy(), [3, 2], B3spline, anscombe=True, bilateral=1)
        from watroo.wavelets import generalized_anscombe
        out[f"ans_fwd_{dt}"] = generalized_anscombe(img.copy(), alpha=2.0, g=1.5, sigma=0.7)
        out[f"ans_inv_{dt}"] = generalized_anscombe(out[f"ans_fwd_{dt}"].copy(), alpha=2.0, g=1.5, sigma=0.7, inverse=True)
    save("f2_noise_weights_anscombe", **out)


def main_nd_bilateral():
    """Bilateral cascade of 1-D signals and 3-D volumes (wavelets.py:433-442 on n-D input: variance through the n-D
    branches of convolution, gather through the dimension-generic atrous_convolution), plus denoise of a volume (the
    3-D bilateral sigma_e table exists; the reference has no 1-D one)."""
    for dt in ("float32", "float64"):
        out = {}
        cases = [((200,), 4, "b3spline", dict(bilateral=1)), ((64,), 3, "triangle", dict(bilateral=[2, 1.5], bilateral_scaling=True)),
                 ((10, 18, 22), 3, "b3spline", dict(bilateral=1)), ((6, 20, 12), 2, "triangle", dict(bilateral=1.5))]
        for k, (shape, level, sf, kw) in enumerate(cases):
            arr = (gaussian(shape, 60 + k, "float64") * 3 + 10 + 4 * np.sin(np.arange(shape[-1]) / 7.0)).astype(dt)
            co = AtrousTransform(SF[sf], **kw)(arr.copy(), level)
            out[f"in{k}"] = arr
            out[f"out{k}"] = co.data.copy()
            out[f"level{k}"] = np.int64(level)
            if arr.ndim == 3:
                out[f"noise{k}"] = np.float64(co.get_noise())
                out[f"dn{k}"] = denoise(arr.copy(), [3, 2][:level], scaling_function=SF[sf], bilateral=kw["bilateral"])
        save(f"transform_nd_bilateral_{dt}", n=np.int64(len(cases)), **out)


WOW_ND_CASES = [((300,), {}), ((300,), dict(denoise_coefficients=[4, 2], weights=[1.5, 1.0, 0.5])),
                ((257,), dict(scaling_function="triangle", h=0.3, preserve_variance=True, denoise_coefficients=[3], noise=1.2)),
                ((24, 40, 44), {}), ((24, 40, 44), dict(denoise_coefficients=[4, 2], soft_threshold=False)),
                ((24, 36, 40), dict(bilateral=1, denoise_coefficients=[3, 1], n_scales=2))]


def main_wow_nd():
    """wow() on 1-D signals and 3-D volumes (utils.py:121-219 is dimension-generic)."""
    for dt in ("float32", "float64"):
        out = {}
        for k, (shape, kw) in enumerate(WOW_ND_CASES):
            kw = dict(kw)
            sf = SF[kw.pop("scaling_function", "b3spline")]
            arr = (gaussian(shape, 80 + k, "float64") * 4 + 30 + 10 * np.sin(np.arange(shape[-1]) / 9.0)).astype(dt)
            recon, co = wow(arr.copy(), scaling_function=sf, **kw)
            out[f"in{k}"] = arr
            out[f"recon{k}"] = recon
            out[f"planes{k}"] = co.data.copy()
            out[f"noise{k}"] = np.float64(np.nan if co.noise is None else co.noise)
        save(f"wow_nd_{dt}", n=np.int64(len(WOW_ND_CASES)), **out)


def main():
    assert watroo.__version__ == "0.0.4", watroo.__version__
    if "--callers" in sys.argv:
        warnings.simplefilter("ignore")
        return main_callers()
    if "--recursive" in sys.argv:
        return main_recursive()
    if "--wide" in sys.argv:  # benchmark-width fixtures (round 2; the others are unchanged)
        warnings.simplefilter("ignore")
        return main_wide()
    if "--f2" in sys.argv:
        warnings.simplefilter("ignore")
        return main_f2()
    if "--nd-bilateral" in sys.argv:
        warnings.simplefilter("ignore")
        return main_nd_bilateral()
    if "--wow-nd" in sys.argv:
        warnings.simplefilter("ignore")
        return main_wow_nd()
    warnings.simplefilter("ignore")
    if "--nd" in sys.argv:  # only the 1-D / 3-D fixtures (added later; the others are unchanged)
        return main_nd()
    main_nd()
    main_callers()
    main_recursive()
    main_wide()
    main_f2()
    main_nd_bilateral()
    main_wow_nd()

    # ---- plain transform: wavelets.py:408-444 via :307 ------------------------------------------------------
    cases = [((64, 64), 4), ((37, 53), 4), ((6, 7), 3), ((96, 64), 6), ((24, 256), 5)]
    for sf in SF:
        for dt in ("float32", "float64"):
            out = {}
            for k, (shape, level) in enumerate(cases):
                img = gaussian(shape, k, dt) * 3 + 10
                out[f"in{k}"] = img
                out[f"out{k}"] = AtrousTransform(SF[sf])(img, level).data
                out[f"level{k}"] = np.int64(level)
            save(f"transform_{sf}_{dt}", n=np.int64(len(cases)), **out)

    # integer input is recast to float64 (wavelets.py:297,319-320)
    img = (gaussian((32, 48), 7, "float64") * 100).astype(np.int16)
    save("transform_int16", img=img, out=AtrousTransform(B3spline)(img, 3).data)

    # known-answer test of the reference's own suite (tests/test_wavelets.py:8-13)
    ones = np.ones((128, 128))
    save("kat_ones", out=AtrousTransform()(ones, 4).data)

    # ---- bilateral transform: wavelets.py:433-440, :24-32, :74-105 -------------------------------------------
    for dt in ("float32", "float64"):
        out = {}
        img = solar_like(64, seed=3, flux=0.05, dtype=dt)
        out["in_solar"] = img
        out["solar_b1"] = AtrousTransform(B3spline, bilateral=1)(img, 4).data
        out["solar_tri_b2"] = AtrousTransform(Triangle, bilateral=2.0)(img, 3).data
        img = gaussian((48, 80), 5, dt)
        out["in_gauss"] = img
        out["gauss_list_scaling"] = AtrousTransform(B3spline, bilateral=[2, 1.5], bilateral_scaling=True)(img, 4).data
        save(f"bilateral_{dt}", **out)

    # ---- denoise / significance: wavelets.py:126-149, utils.py:83-102 ---------------------------------------
    np.random.seed(0)
    img = np.random.normal(size=(512, 512))  # README.md:37-51 (cfg1)
    co = AtrousTransform(Triangle)(img, 2)
    raw = co.data.copy()
    hard = co.significance(3, 1, soft_threshold=False)
    noise = np.float64(co.noise)
    co.denoise([5, 3])
    den = denoise(img, [5, 3], Triangle)
    assert np.array_equal(den, co.data.sum(axis=0)) and np.array_equal(den, np.sum(co, axis=0))
    save("cfg1_denoise_triangle_512", seed=np.int64(0), noise=noise, raw_sub=raw[:, ::4, ::4],
         hard_sub=hard[::4, ::4], hard_count=np.int64(hard.sum()), out_sub=den[::4, ::4],
         out_sum=np.float64(den.sum()), out_abs_sum=np.float64(np.abs(den).sum()))

    for dt in ("float32", "float64"):
        out = {}
        img = gaussian((96, 128), 11, dt) + 0.02 * np.arange(128, dtype=dt)[None, :]
        out["img"] = img
        for sf in SF:
            co = AtrousTransform(SF[sf])(img, 3)
            out[f"{sf}_noise"] = np.float64(co.get_noise())
            out[f"{sf}_soft2"] = co.significance(2.5, 2)
            out[f"{sf}_hard0"] = co.significance(3, 0, soft_threshold=False)
            out[f"{sf}_den_soft"] = denoise(img, [5, 3, 2], SF[sf])
            out[f"{sf}_den_hard"] = denoise(img, [4, 3, 0], SF[sf], soft_threshold=False)
            out[f"{sf}_den_noise"] = denoise(img, [3, 2], SF[sf], noise=0.8)
            out[f"{sf}_den_bilateral"] = denoise(img, [3, 2], SF[sf], bilateral=1)
        save(f"denoise_{dt}", **out)

    # ---- wow: utils.py:105-219 ---------------------------------------------------------------------------------
    for dt in ("float32", "float64"):
        out = {}
        for tag, img in (("gauss", gaussian((64, 64), 21, dt) * 5 + 50),
                         ("solar", solar_like(64, seed=2, flux=0.01, dtype=dt)),
                         ("rect", solar_like(48, seed=4, flux=0.05, dtype=dt, m=80))):
            out[f"{tag}_in"] = img
            for key, kw in (("default", {}),
                            ("den", dict(denoise_coefficients=[5, 2])),
                            ("den_hard", dict(denoise_coefficients=[5, 2], soft_threshold=False)),
                            ("bil", dict(bilateral=1)),
                            ("bil_den", dict(bilateral=1, denoise_coefficients=[5, 2])),
                            ("weights", dict(weights=[0.5, 2.0, 1.5], n_scales=3)),
                            ("nowhite", dict(whitening=False, denoise_coefficients=[3, 1], noise=1.3)),
                            ("tri", dict(scaling_function=Triangle, denoise_coefficients=[0, 3]))):
                if tag == "rect" and key not in ("default", "bil_den"):
                    continue
                recon, co = wow(img.copy(), **kw)
                out[f"{tag}_{key}_recon"] = recon
                if key in ("default", "den_hard", "bil_den", "tri"):
                    out[f"{tag}_{key}_planes"] = co.data
                out[f"{tag}_{key}_noise"] = np.float64(np.nan if co.noise is None else co.noise)
        save(f"wow_{dt}", **out)

    # ---- wow options of SURVEY 8(f) rank 1: gamma blend (h > 0) and preserve_variance, utils.py:157-158,178-184,207-217
    for dt in ("float32", "float64"):
        out = {}
        for tag, img in (("gauss", gaussian((64, 64), 23, dt) * 5 + 50),
                         ("solar", solar_like(64, seed=6, flux=0.01, dtype=dt))):
            out[f"{tag}_in"] = img
            for key, kw in (("gamma", dict(h=0.4, denoise_coefficients=[5, 2], gamma=2.5)),
                            ("gamma_one", dict(h=1, denoise_coefficients=[3, 2, 1])),
                            ("gamma_range", dict(h=0.3, gamma_min=10.0, gamma_max=60.0, n_scales=3)),
                            ("pv", dict(preserve_variance=True)),
                            ("pv_den_gamma", dict(preserve_variance=True, h=0.25, denoise_coefficients=[4, 2],
                                                  weights=[1.5, 0.5]))):
                recon, co = wow(img.copy(), **kw)
                out[f"{tag}_{key}_recon"] = recon
                out[f"{tag}_{key}_planes"] = co.data
        save(f"wow_options_{dt}", **out)

    # ---- compute_noise_weights: wavelets.py:221-229 -------------------------------------------------------------
    for sf in SF:
        np.random.seed(5)
        side = 11 * 2 ** 3
        state = np.random.get_state()
        fields = np.stack([np.random.normal(size=(side, side)).astype(np.float32) for _ in range(2)])
        np.random.set_state(state)
        import watroo.wavelets as ww
        ww.tqdm = lambda it: it  # silence the progress bar
        out = SF[sf](2).compute_noise_weights(3, n_trials=2)
        save(f"noise_weights_{sf}", fields=fields, out=out)


if __name__ == "__main__":
    main()
