"""Row-band mode on real GPUs.  The single-GPU part (a band window of a taller image through wb_atrous_scale_band
equals the rows of the full transform) runs everywhere; the 2-rank NCCL part needs two devices."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_band_window_equals_rows_of_full_transform(dt):
    """One process plays every rank in turn: copies the neighbours' rows into the halo zones by hand, runs the band
    kernel, and must reproduce the unsharded cascade bit for bit (incl. bands whose halo exceeds their height)."""
    import wavelets_b200 as wb
    from wavelets_b200.sharded import _cuda_band_scale, band_range, halo_rows
    # 512 columns: generic row-pipeline kernel; 2048 columns: the lean fp32 kernel (rows wider than 1024)
    for h, w, level, world in ((384, 512, 6, 3), (192, 2048, 5, 2)):
        _band_window_case(dt, h, w, level, world)


def _band_window_case(dt, h, w, level, world):
    import wavelets_b200 as wb
    from wavelets_b200.sharded import _cuda_band_scale, band_range, halo_rows
    gen = torch.Generator(device="cuda").manual_seed(1)
    img = torch.randn((h, w), generator=gen, device="cuda", dtype=torch.float32).to(dt)
    for sf in (wb.B3spline, wb.Triangle):
        full = wb.AtrousTransform(sf)(img, level).data
        taps = len(sf.coefficients_1d)
        pad = halo_rows(level - 1, taps)
        # smooth planes c_s of the unsharded run, to fill halos from
        c = [img]
        for s in range(level):
            c.append(wb.atrous_scale(c[-1], s, sf(2), out_w=False)[0])
        for rank in range(world):
            y0, y1 = band_range(h, rank, world)
            rows = y1 - y0
            for s in range(level):
                halo = halo_rows(s, taps)
                ext = torch.full((rows + 2 * pad, w), float("nan"), dtype=dt, device="cuda")
                g0, g1 = max(0, y0 - halo), min(h, y1 + halo)
                ext[pad + (g0 - y0): pad + (g1 - y0)] = c[s][g0:g1]
                out_c = torch.empty((rows + 2 * pad, w), dtype=dt, device="cuda")
                out_w = torch.empty((rows, w), dtype=dt, device="cuda")
                _cuda_band_scale(ext, pad, out_c, pad, out_w, rows, w, h, y0, s, sf.taps_code)
                assert torch.equal(out_w, full[s, y0:y1]), (sf.__name__, rank, s)
                assert torch.equal(out_c[pad:pad + rows], c[s + 1][y0:y1]), (sf.__name__, rank, s)


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("shape,world", [((384, 512), 3), ((200, 264), 5), ((96, 130), 2), ((160, 2048), 3)])
def test_peer_window_scale_equals_rows_of_full_transform(dt, shape, world):
    """wb_atrous_scale_band_p2p with the ranks' band buffers as separate allocations of ONE device (the address
    arithmetic is the same as with NVLink-mapped peers): every band of every scale must equal the unsharded cascade
    bit for bit, including halos that span several bands / several reflections and the generic (unaligned) kernel."""
    import wavelets_b200 as wb
    from wavelets_b200.sharded import band_range, band_scale_p2p
    h, w = shape
    level = 6
    gen = torch.Generator(device="cuda").manual_seed(3)
    img = torch.randn((h, w), generator=gen, device="cuda", dtype=torch.float32).to(dt)
    for sf in (wb.B3spline, wb.Triangle):
        full = wb.AtrousTransform(sf)(img, level).data
        c = [img]
        for s in range(level):
            c.append(wb.atrous_scale(c[-1], s, sf(2), out_w=False)[0])
        y0s = [band_range(h, k, world)[0] for k in range(world)] + [h]
        for s in range(level):
            bands = [c[s][y0s[k]:y0s[k + 1]].clone() for k in range(world)]  # separate allocations
            ptrs = [b.data_ptr() for b in bands]
            for rank in range(world):
                rows = y0s[rank + 1] - y0s[rank]
                out_c = torch.empty((rows, w), dtype=dt, device="cuda")
                out_w = torch.empty((rows, w), dtype=dt, device="cuda")
                band_scale_p2p(ptrs, y0s, rank, out_c, out_w, w, w, s, sf.taps_code, dt, img.device)
                assert torch.equal(out_w, full[s, y0s[rank]:y0s[rank + 1]]), (sf.__name__, rank, s)
                assert torch.equal(out_c, c[s + 1][y0s[rank]:y0s[rank + 1]]), (sf.__name__, rank, s)


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("shape,level,world", [((384, 512), 6, 3), ((256, 2048), 6, 2), ((200, 130), 5, 3),
                                               ((1024, 4096), 8, 4)])
def test_halo_push_scale_equals_rows_of_full_transform(dt, shape, level, world):
    """wb_atrous_scale_band_push with every rank's padded buffers as separate allocations of ONE device (the address
    arithmetic is the same as with NVLink-mapped peers; launching the ranks one after the other stands for the
    per-scale barrier): each scale kernel stores the next scale's halo rows into the neighbours' buffers, no other
    exchange happens, and every band of every plane equals the unsharded cascade bit for bit.  Buffers start
    NaN-filled, so a halo row that was never pushed would show up."""
    import wavelets_b200 as wb
    from wavelets_b200.sharded import band_range, band_scale_push, halo_rows, push_plan
    h, w = shape
    gen = torch.Generator(device="cuda").manual_seed(4)
    img = torch.randn((h, w), generator=gen, device="cuda", dtype=torch.float32).to(dt)
    for sf in (wb.B3spline, wb.Triangle):
        taps = len(sf.coefficients_1d)
        pad = halo_rows(level - 1, taps)
        bands = [band_range(h, k, world) for k in range(world)]
        if pad > min(b - a for a, b in bands):
            continue  # multi-hop halos are outside the push mode (BandedTransform falls back to the exchange)
        full = wb.AtrousTransform(sf)(img, level).data
        ext = [torch.full((2, (b - a) + 2 * pad, w), float("nan"), dtype=dt, device="cuda") for a, b in bands]
        planes = [torch.empty((level + 1, b - a, w), dtype=dt, device="cuda") for a, b in bands]
        h0 = halo_rows(0, taps)
        for k, (a, b) in enumerate(bands):
            ext[k][0, pad:pad + (b - a)] = img[a:b]
            if a > 0:
                ext[k][0, pad - h0:pad] = img[a - h0:a]
            if b < h:
                ext[k][0, pad + (b - a):pad + (b - a) + h0] = img[b:b + h0]
        row_bytes = w * img.element_size()
        for s in range(level):
            last = s == level - 1
            for k, (a, b) in enumerate(bands):
                rows = b - a
                if last:
                    band_scale_push(ext[k][s & 1], pad, planes[k][level], 0, planes[k][s], rows, w, h, a, s, sf.taps_code)
                else:
                    up, n_up, dn, n_dn = push_plan(h, world, k, halo_rows(s + 1, taps), pad, row_bytes,
                                                   [e[(s + 1) & 1].data_ptr() for e in ext])
                    band_scale_push(ext[k][s & 1], pad, ext[k][(s + 1) & 1], pad, planes[k][s], rows, w, h, a, s,
                                    sf.taps_code, up, n_up, dn, n_dn)
        for k, (a, b) in enumerate(bands):
            assert torch.equal(planes[k], full[:, a:b]), (sf.__name__, k)


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_band_bilateral_equals_rows_of_full_transform(dt):
    """wb_atrous_scale_bilateral_band on every band of every scale (halo rows copied in by hand) equals the rows of the
    unsharded bilateral cascade bit for bit: multi-strip fp32 pair kernel, fp64 row kernel and the generic kernel."""
    import wavelets_b200 as wb
    from wavelets_b200.sharded import _cuda_band_scale, band_range, halo_rows
    for h, w, level, world in ((192, 1024, 4, 3), (96, 130, 3, 2)):
        gen = torch.Generator(device="cuda").manual_seed(6)
        img = (torch.randn((h, w), generator=gen, device="cuda", dtype=torch.float32) * 3 + 20).to(dt)
        for sf, bil in ((wb.B3spline, 1), (wb.Triangle, [2, 1.5])):
            tr = wb.AtrousTransform(sf, bilateral=bil, bilateral_scaling=True)
            full = tr(img, level).data
            factors = tr.var_factors(level)
            taps = len(sf.coefficients_1d)
            pad = halo_rows(level - 1, taps)
            c = [img]
            for s in range(level):
                c.append(wb.atrous_scale(c[-1], s, sf(2), out_w=False, var_factor=factors[s])[0])
            for rank in range(world):
                y0, y1 = band_range(h, rank, world)
                rows = y1 - y0
                for s in range(level):
                    halo = halo_rows(s, taps)
                    ext = torch.full((rows + 2 * pad, w), float("nan"), dtype=dt, device="cuda")
                    g0, g1 = max(0, y0 - halo), min(h, y1 + halo)
                    ext[pad + (g0 - y0): pad + (g1 - y0)] = c[s][g0:g1]
                    out_c = torch.empty((rows + 2 * pad, w), dtype=dt, device="cuda")
                    out_w = torch.empty((rows, w), dtype=dt, device="cuda")
                    _cuda_band_scale(ext, pad, out_c, pad, out_w, rows, w, h, y0, s, sf.taps_code, var_factor=factors[s])
                    assert torch.equal(out_w, full[s, y0:y1]), (sf.__name__, rank, s)
                    assert torch.equal(out_c[pad:pad + rows], c[s + 1][y0:y1]), (sf.__name__, rank, s)


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_band_whitening_equals_rows_of_unsharded_whitening(dt):
    """wb_wow_whiten_scale_band on every band of every scale (halo rows of the raw w_s copied in by hand) equals the
    rows of wb_wow_whiten_scale on the whole plane bit for bit, for the three significance modes."""
    import wavelets_b200 as wb
    from wavelets_b200 import _lib, utils
    from wavelets_b200.sharded import _CudaWowBackend, band_range, halo_rows
    lib = _lib.load(require_cuda=True)
    for h, w, world in ((200, 264, 3), (128, 2048, 2)):
        gen = torch.Generator(device="cuda").manual_seed(5)
        img = torch.randn((h, w), generator=gen, device="cuda", dtype=torch.float32).to(dt) * 3
        sf = wb.B3spline(2)
        raw = wb.AtrousTransform(wb.B3spline)(img, 5).data
        for s in range(5):
            for mode, noise in ((0, 0.0), (1, 0.9), (2, 1.3)):
                ref = torch.empty_like(raw[s])
                utils._whiten_scale(lib, raw[s].contiguous(), ref, s, sf, mode, 2.5, 0.3,
                                    utils._Noise(host=noise) if mode else utils._Noise(), 1.25)
                halo = halo_rows(s, 5)
                for rank in range(world):
                    y0, y1 = band_range(h, rank, world)
                    rows = y1 - y0
                    ext = torch.full((rows + 2 * halo, w), float("nan"), dtype=dt, device="cuda")
                    g0, g1 = max(0, y0 - halo), min(h, y1 + halo)
                    ext[halo + (g0 - y0): halo + (g1 - y0)] = raw[s][g0:g1]
                    out = torch.empty((rows, w), dtype=dt, device="cuda")
                    _CudaWowBackend.whiten(ext, halo, out, rows, w, h, y0, s, sf.taps_code, mode, 2.5, 0.3, noise, 1.25)
                    assert torch.equal(out, ref[y0:y1]), (h, w, s, mode, rank)


def test_peer_window_rejects_bad_arguments():
    from wavelets_b200.sharded import band_scale_p2p
    a = torch.zeros((8, 64), device="cuda")
    o = torch.zeros((8, 64), device="cuda")
    with pytest.raises(RuntimeError):  # boundaries must start at 0 and every rank must own a row
        band_scale_p2p([a.data_ptr(), a.data_ptr()], [0, 8, 8], 0, o, None, 64, 64, 0, 5, torch.float32, a.device)
    with pytest.raises(RuntimeError):  # output aliases an input window
        band_scale_p2p([a.data_ptr()], [0, 8], 0, a, None, 64, 64, 0, 5, torch.float32, a.device)


def test_banded_two_ranks_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "check_banded.py"), "--side", "2048",
           "--levels", "8", "--reps", "2"]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert '"bit_identical": true' in proc.stdout
    assert "p2p   banded == unsharded on all ranks: True" in proc.stdout  # in-kernel NVLink halo reads
    assert "push  banded == unsharded on all ranks: True" in proc.stdout  # in-kernel NVLink halo writes
    assert "banded wow == unsharded two-pass wow on all ranks: True" in proc.stdout
