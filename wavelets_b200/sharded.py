"""Multi-GPU sharding of the hot path; the reference itself is single-process.

Two modes, one process per GPU (``torch.distributed``, NCCL on GPUs, gloo in the CPU tests):

* **Frames** (BASELINE cfg4): independent frames are dealt round-robin to the ranks (``frame_shard``); every global
  reduction of the path (MAD median, residual std) is per frame, so there is NO data-path collective.
* **Row bands** (BASELINE cfg5): one very tall image is cut into contiguous row bands (``band_range``).  The row (x)
  pass of a scale is local; the column (y) pass at scale s needs ``c * 2**s`` rows of ``c_s`` from the bands above
  and below (``c`` = taps // 2), obtained by one neighbour exchange per scale (``exchange_halos``: non-blocking
  send/recv of exactly the rows each peer needs, multi-hop when a halo is taller than a band).  Rows beyond the
  global top/bottom come from the symmetric reflection about the GLOBAL height, which always lands inside a band's
  own window.  Each band then runs ``wb_atrous_scale_band`` -- the same kernel and arithmetic as the single-device
  path, so the sharded planes are bit-identical to the unsharded ones.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from .scaling import B3spline

__all__ = ["frame_shard", "band_range", "halo_rows", "exchange_plan", "exchange_halos", "BandedTransform"]


def frame_shard(n_frames: int, rank: int, world: int) -> range:
    """Indices of the frames rank ``rank`` processes: i -> GPU i mod world."""
    return range(rank, n_frames, world)


def band_range(height: int, rank: int, world: int) -> tuple[int, int]:
    """Global rows [y0, y1) owned by ``rank``: contiguous bands whose sizes differ by at most one row."""
    base, extra = divmod(height, world)
    y0 = rank * base + min(rank, extra)
    return y0, y0 + base + (1 if rank < extra else 0)


def halo_rows(scale: int, n_taps: int) -> int:
    """Rows of c_s a band needs from beyond each of its edges at this scale."""
    return (n_taps // 2) * 2 ** scale


def exchange_plan(height: int, world: int, rank: int, halo: int):
    """Row intervals to receive from / send to every other rank for one scale.

    Returns (recv, send): lists of (peer, g0, g1) global-row intervals.  ``recv``: rows of the halo zones
    [y0 - halo, y0) and [y1, y1 + halo) (clipped to the image) owned by ``peer``; ``send``: rows of my band that lie
    in a halo zone of ``peer``."""
    y0, y1 = band_range(height, rank, world)

    def zones(a, b):
        return [(max(0, a - halo), a), (b, min(height, b + halo))]

    recv, send = [], []
    for peer in range(world):
        if peer == rank:
            continue
        a, b = band_range(height, peer, world)
        for z0, z1 in zones(y0, y1):
            g0, g1 = max(a, z0), min(b, z1)
            if g0 < g1:
                recv.append((peer, g0, g1))
        for z0, z1 in zones(a, b):
            g0, g1 = max(y0, z0), min(y1, z1)
            if g0 < g1:
                send.append((peer, g0, g1))
    return recv, send


def exchange_halos(ext: torch.Tensor, pad: int, height: int, halo: int, group=None) -> None:
    """Fill the halo rows of ``ext`` for one scale.

    ``ext`` has ``pad`` rows above and below the band: row ``pad + i`` holds global row ``y0 + i``.  After the call,
    rows ``[pad - halo, pad)`` and ``[pad + band, pad + band + halo)`` hold the neighbours' rows wherever those rows
    exist in the image."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1 or halo == 0:
        return
    assert halo <= pad, "halo taller than the padding of the band buffer"
    y0, _ = band_range(height, rank, world)
    recv, send = exchange_plan(height, world, rank, halo)
    # one batched group of non-blocking sends and receives (ncclGroupStart/End on NCCL: no ordering deadlock)
    ops = []
    for peer, g0, g1 in send:
        rows = ext[pad + (g0 - y0): pad + (g1 - y0)]
        ops.append(dist.P2POp(dist.isend, rows, dist.get_global_rank(group, peer) if group is not None else peer,
                              group))
    for peer, g0, g1 in recv:
        rows = ext[pad + (g0 - y0): pad + (g1 - y0)]
        ops.append(dist.P2POp(dist.irecv, rows, dist.get_global_rank(group, peer) if group is not None else peer,
                              group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()


def _cuda_band_scale(ext_in, pad, ext_out, out_pad, w_out, band_rows, width, height, y0, scale, taps_code):
    """One scale of one band on the device: wb_atrous_scale_band."""
    lib = _lib.load(require_cuda=True)
    with torch.cuda.device(ext_in.device):
        _lib.check(lib.wb_atrous_scale_band(ext_in.data_ptr(), ext_out.data_ptr(), w_out.data_ptr(), band_rows, width,
                                            height, y0, pad, ext_in.stride(0), out_pad, ext_out.stride(0), 0,
                                            w_out.stride(0), scale, taps_code, _lib.dtype_code(ext_in.dtype),
                                            _lib.stream_ptr(ext_in.device)))


class BandedTransform:
    """Plain à trous cascade of ONE image sharded by row bands over the ranks of a process group.

    ``band`` is this rank's rows ``band_range(global_height, rank, world)`` of the image (device tensor on NCCL,
    CPU tensor with a custom ``scale_fn`` in the gloo tests).  Returns this rank's rows of the coefficient planes,
    shape ``(level + 1, band_rows, W)``.  ``scale_fn`` computes one scale of one band (default: the CUDA kernel)."""

    def __init__(self, scaling_function_class=B3spline, group=None, scale_fn=None, poison=False):
        self.scaling_function_class = scaling_function_class
        self.group = group
        self.scale_fn = scale_fn or _cuda_band_scale
        self.poison = poison  # tests: NaN-fill the padded buffers so that any read of an unfilled halo row shows up

    def __call__(self, band: torch.Tensor, level: int, global_height: int) -> torch.Tensor:
        sf = self.scaling_function_class(2)
        n_taps = len(sf.coefficients_1d)
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        y0, y1 = band_range(global_height, rank, world)
        rows, width = band.shape
        assert rows == y1 - y0, f"rank {rank}: band has {rows} rows, expected {y1 - y0}"
        planes = torch.empty((level + 1, rows, width), dtype=band.dtype, device=band.device)
        if level == 0:
            planes[0].copy_(band)
            return planes
        pad = halo_rows(level - 1, n_taps) if world > 1 else 0
        pad = min(pad, global_height)
        # two padded buffers hold the running smooth plane c_s (ping-pong); row pad + i <-> global row y0 + i
        ext = [torch.empty((rows + 2 * pad, width), dtype=band.dtype, device=band.device) for _ in range(2)]
        if self.poison:
            for e in ext:
                e.fill_(float("nan"))
        ext[0][pad:pad + rows].copy_(band)
        for s in range(level):
            cur, nxt = ext[s & 1], ext[(s + 1) & 1]
            halo = min(halo_rows(s, n_taps), pad)
            if world > 1:
                exchange_halos(cur, pad, global_height, halo, self.group)
            last = s == level - 1
            out_c, out_pad = (planes[level], 0) if last else (nxt, pad)
            self.scale_fn(cur, pad, out_c, out_pad, planes[s], rows, width, global_height, y0, s, sf.taps_code)
        return planes
