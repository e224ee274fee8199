"""Host-side logic of the multi-GPU modes, on CPU: frame sharding, band geometry, and the per-scale halo exchange
run for real over the gloo backend with world sizes 2 and 3 (the per-band arithmetic is done by the oracle here; on
GPUs it is wb_atrous_scale_band -- see tests/test_sharded_gpu.py and tools/check_banded.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import atrous_oracle as orc
from wavelets_b200.sharded import BandedTransform, band_range, exchange_plan, frame_shard, halo_rows


def test_frame_shard_partitions_all_frames():
    for n, world in ((2048, 8), (7, 3), (2, 4), (0, 2)):
        seen = sorted(i for r in range(world) for i in frame_shard(n, r, world))
        assert seen == list(range(n))
    assert list(frame_shard(2048, 3, 8))[:3] == [3, 11, 19] and len(frame_shard(2048, 3, 8)) == 256


def test_band_range_and_halo():
    for h, world in ((32768, 8), (100, 3), (5, 4)):
        bands = [band_range(h, r, world) for r in range(world)]
        assert bands[0][0] == 0 and bands[-1][1] == h
        assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in bands]
        assert max(sizes) - min(sizes) <= 1
    assert band_range(32768, 3, 8) == (12288, 16384)
    assert halo_rows(11, 5) == 4096 and halo_rows(0, 3) == 1   # cfg5: s = 11 needs one whole neighbour band


def test_exchange_plan_is_symmetric_and_complete():
    for h, world, halo in ((64, 2, 4), (60, 3, 8), (60, 3, 32), (32, 4, 16)):
        plans = [exchange_plan(h, world, r, halo) for r in range(world)]
        for r in range(world):
            recv, send = plans[r]
            for peer, g0, g1 in recv:          # whatever I receive, the owner sends
                assert (r, g0, g1) in plans[peer][1]
            y0, y1 = band_range(h, r, world)
            need = set(range(max(0, y0 - halo), y0)) | set(range(y1, min(h, y1 + halo)))
            got = set()
            for _, g0, g1 in recv:
                got |= set(range(g0, g1))
            assert got == need


def _oracle_band_scale(ext_in, pad, out_c, out_pad, w_out, rows, width, height, y0, scale, taps_code, var_factor=None):
    """One scale of one band with the oracle's arithmetic (float64), reading only rows the band is entitled to."""
    name = "b3spline" if taps_code == 5 else "triangle"
    taps = orc.TAPS[name]
    c, d = len(taps) // 2, 2 ** scale
    a = ext_in.numpy()
    xs = np.arange(width)
    gy = np.arange(y0, y0 + rows)
    if var_factor is not None:
        # bilateral scale (wavelets.py:433-442) on the band: variance and range-weighted gather from the same 25 taps
        x = a[pad:pad + rows]
        tap = {}
        for i in range(len(taps)):
            src = a[orc.reflect_index(gy + (i - c) * d, height) - y0 + pad]
            for j in range(len(taps)):
                tap[i, j] = src[:, orc.reflect_index(xs + (j - c) * d, width)]
        assert all(np.isfinite(t).all() for t in tap.values()), "a halo row that was never exchanged has been read"
        mean = sum(taps[i] * taps[j] * tap[i, j] for i, j in tap)
        var = sum(taps[i] * taps[j] * tap[i, j] ** 2 for i, j in tap) - mean ** 2
        var[var <= 0] = 1e-20
        var = var * var_factor
        num, den = taps[c] ** 2 * x, np.full_like(x, taps[c] ** 2)
        for (i, j), t in tap.items():
            if (i, j) != (c, c):
                g = taps[i] * taps[j] * np.exp(-((x - t) ** 2) / var / 2)
                num = num + g * t
                den = den + g
        res = num / den
        out_c[out_pad:out_pad + rows] = torch.from_numpy(res)
        w_out[:] = torch.from_numpy(x - res)
        return
    out = np.zeros((rows, width))
    for i, ti in enumerate(taps):
        src = a[orc.reflect_index(gy + (i - c) * d, height) - y0 + pad]
        rowf = np.zeros((rows, width))
        for j, tj in enumerate(taps):
            rowf += tj * src[:, orc.reflect_index(xs + (j - c) * d, width)]
        out += ti * rowf
    assert np.isfinite(out).all(), "a halo row that was never exchanged has been read"
    out_c[out_pad:out_pad + rows] = torch.from_numpy(out)
    w_out[:] = torch.from_numpy(a[pad:pad + rows] - out)


def _worker(rank, world, port, height, width, level, sf_name, result_dir, bilateral=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import wavelets_b200 as wb
        img = np.random.default_rng(42).standard_normal((height, width))
        y0, y1 = band_range(height, rank, world)
        sf = {"b3spline": wb.B3spline, "triangle": wb.Triangle}[sf_name]
        planes = BandedTransform(sf, scale_fn=_oracle_band_scale, poison=True, bilateral=bilateral,
                                 bilateral_scaling=bilateral is not None)(torch.from_numpy(img[y0:y1].copy()), level, height)
        np.save(os.path.join(result_dir, f"band{rank}.npy"), planes.numpy())
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,height,width,level,sf", [(2, 64, 48, 4, "b3spline"), (3, 50, 40, 4, "triangle"),
                                                         (3, 36, 32, 4, "b3spline")])
def test_banded_cascade_over_gloo(tmp_path, world, height, width, level, sf):
    """world_size-2/3 gloo run: bands + per-scale halo exchange reproduce the unsharded oracle cascade.  The third
    case has halos taller than a band (multi-hop exchange) and reflections reaching into neighbour bands."""
    mp.spawn(_worker, args=(world, _free_port(), height, width, level, sf, str(tmp_path)), nprocs=world, join=True)
    img = np.random.default_rng(42).standard_normal((height, width))
    want = orc.atrous_transform(img, level, sf, backend="numpy")
    got = np.concatenate([np.load(tmp_path / f"band{r}.npy") for r in range(world)], axis=1)
    assert got.shape == want.shape
    for p in range(level + 1):
        assert orc.emax(got[p], want[p]) < 1e-13, (p, orc.emax(got[p], want[p]))


@pytest.mark.parametrize("world,height,width,level,sf,bilateral", [(2, 64, 48, 3, "b3spline", 1),
                                                                   (3, 60, 40, 4, "triangle", [2, 1.5])])
def test_banded_bilateral_cascade_over_gloo(tmp_path, world, height, width, level, sf, bilateral):
    """Row bands + per-scale halo exchange reproduce the unsharded BILATERAL oracle cascade (same halo as the plain one)."""
    mp.spawn(_worker, args=(world, _free_port(), height, width, level, sf, str(tmp_path), bilateral), nprocs=world, join=True)
    img = np.random.default_rng(42).standard_normal((height, width))
    want = orc.atrous_transform(img, level, sf, bilateral=bilateral, bilateral_scaling=True, backend="numpy")
    got = np.concatenate([np.load(tmp_path / f"band{r}.npy") for r in range(world)], axis=1)
    for p in range(level + 1):
        assert orc.emax(got[p], want[p]) < 1e-12, (p, orc.emax(got[p], want[p]))


class _OracleWowBackend:
    """Per-band arithmetic of BandedWow with the oracle's formulas (float64) -- the CUDA kernels' stand-in on CPU."""

    scale = staticmethod(_oracle_band_scale)

    @staticmethod
    def whiten(ext_w, pad, out, rows, width, height, y0, scale, taps_code, sig_mode, sigma, sigma_e, noise, weight):
        from scipy import special
        name = "b3spline" if taps_code == 5 else "triangle"
        taps = orc.TAPS[name]
        c, d = len(taps) // 2, 2 ** scale
        a = ext_w.numpy()
        xs = np.arange(width)
        gy = np.arange(y0, y0 + rows)
        power = np.zeros((rows, width))
        for i, ti in enumerate(taps):
            src = a[orc.reflect_index(gy + (i - c) * d, height) - y0 + pad] ** 2
            rowf = np.zeros((rows, width))
            for j, tj in enumerate(taps):
                rowf += tj * src[:, orc.reflect_index(xs + (j - c) * d, width)]
            power += ti * rowf
        assert np.isfinite(power).all(), "a halo row of w_s that was never exchanged has been read"
        power[power <= 0] = 1e-15
        w = a[pad:pad + rows].copy()
        if sig_mode == 1:
            w *= special.erf(np.abs(w / (sigma * noise * sigma_e)))
        elif sig_mode == 2:
            w *= np.abs(w) > sigma * noise * sigma_e
        out[:] = torch.from_numpy(w * (weight / np.sqrt(power)))

    @staticmethod
    def moments(plane):
        p = plane.numpy()
        return torch.tensor([p.size, p.mean(), p.var()], dtype=torch.float64)

    @staticmethod
    def synthesis(planes):
        return planes.sum(dim=0)


def _wow_worker(rank, world, port, height, width, kw, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import wavelets_b200 as wb
        from wavelets_b200.sharded import BandedWow
        img = orc.solar_like(height, seed=4, flux=0.05, dtype=np.float64, m=width)
        y0, y1 = band_range(height, rank, world)
        recon, planes = BandedWow(wb.B3spline, backend=_OracleWowBackend, poison=True)(
            torch.from_numpy(img[y0:y1].copy()), height, **kw)
        np.save(os.path.join(result_dir, f"recon{rank}.npy"), recon.numpy())
        np.save(os.path.join(result_dir, f"planes{rank}.npy"), planes.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,height,width,kw", [
    (2, 64, 48, dict()),
    (3, 72, 64, dict(n_scales=3, weights=[1.5, 1.0, 0.5], denoise_coefficients=[4, 2], noise=1.7)),
    (3, 48, 40, dict(denoise_coefficients=[3], noise=2.0, soft_threshold=False)),
    (3, 60, 52, dict(denoise_coefficients=[4, 2])),   # noise=None: distributed exact MAD estimate of the raw w_0
    (2, 60, 52, dict(denoise_coefficients=[0, 3])),   # first threshold at scale 1: MAD of the WHITENED plane 0
    (2, 64, 56, dict(bilateral=1, denoise_coefficients=[4, 2])),   # bilateral cascade on bands, bilateral sigma_e table
    (3, 64, 64, dict(denoise_coefficients=[0, 0, 2], soft_threshold=False)),
])
def test_banded_wow_over_gloo(tmp_path, world, height, width, kw):
    """Row-band WOW over gloo (two halo exchanges per scale + the all-gather of the residual moments) reproduces the
    unsharded oracle wow(): recon and every whitened plane."""
    mp.spawn(_wow_worker, args=(world, _free_port(), height, width, kw, str(tmp_path)), nprocs=world, join=True)
    img = orc.solar_like(height, seed=4, flux=0.05, dtype=np.float64, m=width)
    recon, planes, _ = orc.wow(img, "b3spline", backend="numpy", **kw)
    got_p = np.concatenate([np.load(tmp_path / f"planes{r}.npy") for r in range(world)], axis=1)
    got_r = np.concatenate([np.load(tmp_path / f"recon{r}.npy") for r in range(world)], axis=0)
    assert got_p.shape == planes.shape
    for p in range(planes.shape[0]):
        assert orc.emax(got_p[p], planes[p]) < 1e-11, (p, orc.emax(got_p[p], planes[p]))
    assert orc.emax(got_r, recon) < 1e-11


def _median_worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from wavelets_b200.sharded import distributed_abs_median
        out = {}
        for name, full in _median_cases().items():
            y0, y1 = band_range(full.shape[0], rank, world)
            out[name] = distributed_abs_median(torch.from_numpy(full[y0:y1].copy())).numpy()
        if rank == 0:
            np.savez(os.path.join(result_dir, "med.npz"), **out)
    finally:
        dist.destroy_process_group()


def _median_cases():
    rng = np.random.default_rng(3)
    cases = {
        "odd32": rng.standard_normal((31, 33)).astype(np.float32),          # odd count
        "even32": rng.standard_normal((30, 34)).astype(np.float32) * 1e-3,  # even count: mean of the two middle values
        "even64": rng.standard_normal((24, 50)) * 1e5,
        "ties32": np.round(rng.standard_normal((40, 40)) * 2).astype(np.float32),  # heavy ties, many exact zeros
        "const64": np.full((9, 9), -2.5),
    }
    cases["half_zero32"] = np.concatenate([np.zeros((20, 16), np.float32), cases["odd32"][:20, :16]])
    return cases


@pytest.mark.parametrize("world", [1, 3])
def test_distributed_abs_median_is_exact(tmp_path, world):
    """The radix select over the ranks returns np.median(np.abs(x)) bit for bit (odd / even counts, ties, zeros)."""
    if world == 1:
        from wavelets_b200.sharded import distributed_abs_median
        got = {k: distributed_abs_median(torch.from_numpy(v)).numpy() for k, v in _median_cases().items()}
    else:
        mp.spawn(_median_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
        got = dict(np.load(tmp_path / "med.npz"))
    for name, full in _median_cases().items():
        want = np.median(np.abs(full))
        assert got[name].dtype == full.dtype and got[name] == want, (name, got[name], want)
