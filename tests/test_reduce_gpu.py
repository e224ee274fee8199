"""K4 reductions on the GPU: exact |x| median (bit-identical to np.median), moments, synthesis, RNG."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _median_cases():
    rng = np.random.default_rng(0)
    yield "gauss_even", rng.standard_normal((512, 384))
    yield "gauss_odd", rng.standard_normal((255, 257))
    yield "tiny", rng.standard_normal((3, 5))
    yield "single", np.array([[-2.5]])
    yield "two", np.array([[1.0, -3.0]])
    yield "small_4097", rng.standard_normal((1, 4097))
    yield "ties_integers", rng.poisson(3.0, (300, 400)).astype(np.float64) - 3
    yield "constant", np.full((64, 64), 7.0)
    yield "zeros", np.zeros((100, 100))
    yield "bimodal_gap", np.concatenate([np.full(5000, 1.0), np.full(5000, 1000.0)]).reshape(100, 100)
    yield "bimodal_odd", np.concatenate([rng.uniform(0, 1, 5001), rng.uniform(1e6, 2e6, 5000)]).reshape(1, -1)
    yield "heavy_tail", rng.standard_cauchy((400, 400)) * 1e3
    yield "lognormal_wide", np.exp(rng.normal(0, 10, (300, 300)))
    yield "mostly_zero", np.where(rng.uniform(size=(256, 256)) < 0.7, 0.0, rng.standard_normal((256, 256)))
    yield "denormal_mix", np.where(rng.uniform(size=(128, 128)) < 0.5, 1e-42, rng.standard_normal((128, 128)))
    yield "sorted_ramp", np.arange(200 * 200, dtype=np.float64).reshape(200, 200)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_abs_median_is_bit_exact(dt):
    from wavelets_b200.wavelets import abs_median
    for name, arr in _median_cases():
        a = arr.astype(dt)
        want = np.median(np.abs(a))
        for compact in (True, False):  # with / without the filter pass's compact buffer: same exact result
            got = abs_median(torch.from_numpy(a).cuda(), compact=compact).cpu().numpy()[0]
            assert got.dtype == want.dtype
            assert got == want or (np.isnan(got) and np.isnan(want)), (name, dt, compact, got, want)


def test_abs_median_batched_and_large():
    from wavelets_b200.wavelets import abs_median, abs_median_noise
    gen = torch.Generator(device="cuda").manual_seed(3)
    stack = torch.randn((3, 1024, 2048), generator=gen, device="cuda") * torch.tensor([1.0, 10.0, 0.1], device="cuda")[:, None, None]
    got = abs_median(stack).cpu().numpy()
    host = stack.cpu().numpy()
    for b in range(3):
        assert got[b] == np.median(np.abs(host[b]))
    noise = abs_median_noise(stack, 0.8907).cpu().numpy()
    for b in range(3):
        want = np.median(np.abs(host[b])) / 0.6745 / np.float64(0.8907)
        assert isinstance(want, np.float64) and noise[b] == want
    # full-size plane (4096^2), the size the MAD estimate runs on in BASELINE cfg3
    big = torch.randn((4096, 4096), generator=gen, device="cuda")
    assert abs_median(big).cpu().numpy()[0] == np.median(np.abs(big.cpu().numpy()))
    assert abs_median(big, compact=False).cpu().numpy()[0] == np.median(np.abs(big.cpu().numpy()))
    # heavy ties around the median (more than a quarter of the plane inside the sampled bracket: the compact buffer
    # overflows and the passes fall back to the plane), odd element counts, unaligned views
    ties = torch.round(torch.randn((1500, 1501), generator=gen, device="cuda") * 2)
    assert abs_median(ties).cpu().numpy()[0] == np.median(np.abs(ties.cpu().numpy()))
    tall = torch.randn((2049, 1023), generator=gen, device="cuda", dtype=torch.float64) ** 3
    assert abs_median(tall).cpu().numpy()[0] == np.median(np.abs(tall.cpu().numpy()))
    off = torch.randn(3_000_001, generator=gen, device="cuda")[1:].reshape(1, -1)  # 4-byte aligned only
    assert abs_median(off).cpu().numpy()[0] == np.median(np.abs(off.cpu().numpy()))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_plane_moments(dt):
    from wavelets_b200.wavelets import plane_moments
    rng = np.random.default_rng(1)
    planes = (rng.standard_normal((4, 300, 500)) * np.array([1, 5, 0.01, 100])[:, None, None]
              + np.array([0, 1e4, -3, 1e6])[:, None, None]).astype(dt)
    got = plane_moments(torch.from_numpy(planes).cuda()).cpu().numpy()
    p64 = planes.astype(np.float64)
    assert np.allclose(got[:, 0], p64.mean(axis=(1, 2)), rtol=1e-13, atol=0)
    assert np.allclose(got[:, 2], p64.std(axis=(1, 2)), rtol=1e-9, atol=0)
    const = torch.full((64, 64), 3.25, dtype=torch.float64 if dt == np.float64 else torch.float32, device="cuda")
    assert plane_moments(const).cpu().numpy()[0, 2] == 0.0


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_synthesis_matches_numpy_order(dt):
    from wavelets_b200.wavelets import synthesis
    rng = np.random.default_rng(2)
    planes = (rng.standard_normal((7, 130, 257)) * 10.0 ** rng.integers(-3, 4, (7, 1, 1))).astype(dt)
    got = synthesis(torch.from_numpy(planes).cuda()).cpu().numpy()
    assert np.array_equal(got, np.sum(planes, axis=0))  # same order, same dtype -> bit-identical
    stack = torch.from_numpy(np.stack([planes, planes[::-1].copy()])).cuda()
    got2 = synthesis(stack).cpu().numpy()
    assert np.array_equal(got2[0], np.sum(planes, axis=0)) and np.array_equal(got2[1], np.sum(planes[::-1], axis=0))


def test_randn_field_statistics():
    from wavelets_b200.wavelets import randn_field
    a = randn_field((2048, 2048), seed=7).cpu().numpy().astype(np.float64)
    b = randn_field((2048, 2048), seed=7, offset=2048 * 2048 // 4).cpu().numpy().astype(np.float64)
    c = randn_field((2048, 2048), seed=7).cpu().numpy().astype(np.float64)
    assert np.array_equal(a, c) and not np.array_equal(a, b)
    n = a.size
    for x in (a, b):
        assert abs(x.mean()) < 5 / np.sqrt(n) and abs(x.std() - 1) < 5 / np.sqrt(2 * n)
        assert abs((x ** 3).mean()) < 5 * np.sqrt(15 / n) and abs((x ** 4).mean() - 3) < 5 * np.sqrt(96 / n)
        assert np.isfinite(x).all() and np.abs(x).max() < 7
    assert abs(np.mean(a * b)) < 5 / np.sqrt(n)
    assert abs(np.mean(a[:, 1:] * a[:, :-1])) < 5 / np.sqrt(n) and abs(np.mean(a[1:] * a[:-1])) < 5 / np.sqrt(n)
