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
  ``BandedTransform(p2p=True)`` removes the exchange altogether: the ping-pong band buffers live in symmetric memory
  (``torch.distributed._symmetric_memory``: every rank's buffer is mapped into every process over NVLink) and
  ``wb_atrous_scale_band_p2p`` fetches each halo row straight from its owner's buffer with the same TMA bulk copy as
  a local row -- compute and communication are one kernel; the only synchronisation is one device-side barrier per
  scale on the launch stream (no host sync, no staging copy, no padded buffer).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from .scaling import B3spline

__all__ = ["frame_shard", "band_range", "halo_rows", "exchange_plan", "exchange_halos", "BandedTransform",
           "band_scale_p2p", "band_scale_push", "push_plan", "PeerBandBuffers", "BandedWow", "distributed_abs_median"]


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


def _cuda_band_scale(ext_in, pad, ext_out, out_pad, w_out, band_rows, width, height, y0, scale, taps_code,
                     var_factor=None):
    """One scale of one band on the device: wb_atrous_scale_band, or wb_atrous_scale_bilateral_band when ``var_factor``
    (sigma_b[s]**2 * (s + 1 if bilateral_scaling), watroo/wavelets.py:434-436) is given."""
    lib = _lib.load(require_cuda=True)
    with torch.cuda.device(ext_in.device):
        if var_factor is None:
            _lib.check(lib.wb_atrous_scale_band(ext_in.data_ptr(), ext_out.data_ptr(), w_out.data_ptr(), band_rows, width,
                                                height, y0, pad, ext_in.stride(0), out_pad, ext_out.stride(0), 0,
                                                w_out.stride(0), scale, taps_code, _lib.dtype_code(ext_in.dtype),
                                                _lib.stream_ptr(ext_in.device)))
        else:
            _lib.check(lib.wb_atrous_scale_bilateral_band(
                ext_in.data_ptr(), ext_out.data_ptr(), w_out.data_ptr(), band_rows, width, height, y0, pad,
                ext_in.stride(0), out_pad, ext_out.stride(0), 0, w_out.stride(0), scale, taps_code,
                _lib.dtype_code(ext_in.dtype), float(var_factor), _lib.stream_ptr(ext_in.device)))


def band_scale_p2p(peer_ptrs, peer_y0, rank, out_c, out_w, width, pitch, scale, taps_code, dtype, device):
    """One scale of band ``rank`` with every input row read from its owner's buffer: wb_atrous_scale_band_p2p.

    ``peer_ptrs[k]`` is the address (valid on ``device``) of row 0 of rank k's band of c_s, ``peer_y0`` the
    ``world + 1`` band boundaries.  ``out_c`` / ``out_w`` are local ``(band_rows, width)`` tensors (or None)."""
    import ctypes
    lib = _lib.load(require_cuda=True)
    n = len(peer_ptrs)
    ptrs = (ctypes.c_void_p * n)(*[int(a) for a in peer_ptrs])
    y0s = (ctypes.c_longlong * (n + 1))(*[int(y) for y in peer_y0])
    with torch.cuda.device(device):
        _lib.check(lib.wb_atrous_scale_band_p2p(
            ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(y0s, ctypes.c_void_p), n, rank,
            out_c.data_ptr() if out_c is not None else None, out_w.data_ptr() if out_w is not None else None,
            width, int(peer_y0[-1]), pitch, out_c.stride(0) if out_c is not None else 0,
            out_w.stride(0) if out_w is not None else 0, scale, taps_code, _lib.dtype_code(dtype),
            _lib.stream_ptr(device)))


def band_scale_push(ext_in, pad, out_c, out_pad, w_out, band_rows, width, height, y0, scale, taps_code,
                    push_up_ptr=0, push_up_rows=0, push_dn_ptr=0, push_dn_rows=0):
    """One scale of one band that also stores the next scale's halo rows into the neighbours' padded buffers:
    wb_atrous_scale_band_push.  ``push_up_ptr`` / ``push_dn_ptr``: address (valid on this device) of the row that this
    band's output row 0 occupies in the upper / lower neighbour's padded c buffer (0: no neighbour)."""
    lib = _lib.load(require_cuda=True)
    with torch.cuda.device(ext_in.device):
        _lib.check(lib.wb_atrous_scale_band_push(
            ext_in.data_ptr(), out_c.data_ptr(), w_out.data_ptr() if w_out is not None else None, band_rows, width,
            height, y0, pad, ext_in.stride(0), out_pad, out_c.stride(0), 0,
            w_out.stride(0) if w_out is not None else 0, push_up_ptr or None, int(push_up_rows), push_dn_ptr or None,
            int(push_dn_rows), scale, taps_code, _lib.dtype_code(ext_in.dtype), _lib.stream_ptr(ext_in.device)))


def push_plan(height: int, world: int, rank: int, halo_next: int, pad: int, row_bytes: int, base_ptrs):
    """Where band ``rank`` pushes the halo rows of the NEXT scale (``halo_next`` rows beyond each band edge).

    ``base_ptrs[k]``: address of row 0 of rank k's padded buffer (the half that receives c_{s+1}); row ``pad + i`` of a
    buffer holds that rank's band row i.  Returns (push_up_ptr, push_up_rows, push_dn_ptr, push_dn_rows): my output row
    0 sits ``rows_up`` rows below the upper neighbour's band start and ``rows_mine`` rows above the lower neighbour's."""
    y0, y1 = band_range(height, rank, world)
    rows = y1 - y0
    up = dn = 0
    n_up = n_dn = 0
    if halo_next > 0 and rank > 0:
        a, b = band_range(height, rank - 1, world)
        n_up = min(halo_next, rows)
        up = int(base_ptrs[rank - 1]) + (pad + (b - a)) * row_bytes
    if halo_next > 0 and rank < world - 1:
        n_dn = min(halo_next, rows)
        dn = int(base_ptrs[rank + 1]) + (pad - rows) * row_bytes
    return up, n_up, dn, n_dn


_PEER_BUFFERS: dict = {}
TRACE_EVENTS = None  # tools/trace_push.py: a list that receives one CUDA event per scale boundary of the push cascade


class PeerBandBuffers:
    """Ping-pong buffers for the running smooth plane c_s of one band, allocated in symmetric memory so that every
    rank of the group can address every other rank's rows (NVLink peer mapping).  ``buf[b]`` is this rank's
    ``(rows_max, width)`` half b; ``ptrs(b)[k]`` is the address of rank k's half b as seen from this device."""

    def __init__(self, rows_max: int, width: int, dtype: torch.dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.shape = (2, rows_max, width)
        self.buf = symm_mem.empty(self.shape, dtype=dtype, device=device)
        self.handle = symm_mem.rendezvous(self.buf, group)
        half = rows_max * width * self.buf.element_size()
        self._ptrs = [[int(base) + b * half for base in self.handle.buffer_ptrs] for b in range(2)]

    def peer(self, k: int) -> torch.Tensor:
        """Rank k's whole buffer ``(2, rows_max, width)`` as a tensor on THIS device (peer-mapped memory)."""
        return self.handle.get_buffer(k, self.shape, self.buf.dtype)

    def ptrs(self, b: int):
        return self._ptrs[b]

    def barrier(self):
        """Device-side barrier of all ranks on the current stream (signal pads in symmetric memory)."""
        self.handle.barrier(channel=0)


class BandedTransform:
    """À trous cascade (plain, or bilateral with ``bilateral=``) of ONE image sharded by row bands over the ranks of a
    process group.

    ``band`` is this rank's rows ``band_range(global_height, rank, world)`` of the image (device tensor on NCCL,
    CPU tensor with a custom ``scale_fn`` in the gloo tests).  Returns this rank's rows of the coefficient planes,
    shape ``(level + 1, band_rows, W)``.  ``scale_fn`` computes one scale of one band (default: the CUDA kernel)."""

    def __init__(self, scaling_function_class=B3spline, group=None, scale_fn=None, poison=False, p2p=False, push=False,
                 bilateral=None, bilateral_scaling=False):
        self.scaling_function_class = scaling_function_class
        self.bilateral = bilateral                  # as AtrousTransform: None, a scalar or a per-scale list
        self.bilateral_scaling = bilateral_scaling  # (the bilateral cascade takes the halo-exchange transport)
        self.group = group
        self.scale_fn = scale_fn or _cuda_band_scale
        self.poison = poison  # tests: NaN-fill the padded buffers so that any read of an unfilled halo row shows up
        self.p2p = p2p        # halo rows read in place from the neighbours' buffers (NVLink), no exchange
        self.push = push      # halo rows of the next scale stored into the neighbours' buffers by the scale kernel

    def _call_push(self, band, planes, level, global_height, sf, rank, world):
        """Compute and communication in ONE kernel per scale, transfers as posted writes: the scale kernel stores the
        rows of c_{s+1} that the neighbours need at scale s+1 straight into their padded buffers (symmetric memory over
        NVLink) while it streams its own band; every scale then reads local memory only.  One device-side barrier per
        scale orders the ranks (no host synchronisation).  Single-hop halos only (the caller checks)."""
        rows, width = band.shape
        n_taps = len(sf.coefficients_1d)
        rows_max = band_range(global_height, 0, world)[1]
        pad = halo_rows(level - 1, n_taps)
        key = ("push", id(self.group), rows_max + 2 * pad, width, band.dtype, band.device.index)
        peer = _PEER_BUFFERS.get(key)
        if peer is None:
            peer = _PEER_BUFFERS[key] = PeerBandBuffers(rows_max + 2 * pad, width, band.dtype, band.device, self.group)
        if self.poison:
            peer.buf.fill_(float("nan"))  # safe: the previous call ended with a barrier
            peer.barrier()                # ... and no neighbour may write its scale-0 halo rows before the fill
        row_bytes = width * band.element_size()
        ext = peer.buf
        ext[0, pad:pad + rows].copy_(band)
        # halo of scale 0: my edge rows go to the neighbours by plain peer copies
        h0 = halo_rows(0, n_taps)
        if rank > 0:
            a, b = band_range(global_height, rank - 1, world)
            peer.peer(rank - 1)[0, pad + (b - a): pad + (b - a) + h0].copy_(band[:h0])
        if rank < world - 1:
            peer.peer(rank + 1)[0, pad - h0: pad].copy_(band[rows - h0:])
        y0 = band_range(global_height, rank, world)[0]
        for s in range(level):
            if TRACE_EVENTS is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                TRACE_EVENTS.append(ev)
            # every rank has received its halo of c_s and has finished reading the half this scale overwrites
            peer.barrier()
            last = s == level - 1
            cur = ext[s & 1]
            if last:
                band_scale_push(cur, pad, planes[level], 0, planes[s], rows, width, global_height, y0, s, sf.taps_code)
            else:
                up, n_up, dn, n_dn = push_plan(global_height, world, rank, halo_rows(s + 1, n_taps), pad, row_bytes,
                                               peer.ptrs((s + 1) & 1))
                band_scale_push(cur, pad, ext[(s + 1) & 1], pad, planes[s], rows, width, global_height, y0, s,
                                sf.taps_code, up, n_up, dn, n_dn)
        peer.barrier()  # nobody may refill the buffers (next call) while a neighbour still reads them
        return planes

    def _call_p2p(self, band, planes, level, global_height, sf, rank, world):
        rows, width = band.shape
        rows_max = band_range(global_height, 0, world)[1]
        # symmetric allocations are collective and never shrink: keep one set per (group, geometry) for the process
        key = (id(self.group), rows_max, width, band.dtype, band.device.index)
        peer = _PEER_BUFFERS.get(key)
        if peer is None:
            peer = _PEER_BUFFERS[key] = PeerBandBuffers(rows_max, width, band.dtype, band.device, self.group)
        if self.poison:
            peer.buf.fill_(float("nan"))  # safe: the previous call ended with a barrier
        y0s = [band_range(global_height, k, world)[0] for k in range(world)] + [global_height]
        peer.buf[0, :rows].copy_(band)
        for s in range(level):
            # every rank has written its c_s and has finished reading the half this scale overwrites
            peer.barrier()
            last = s == level - 1
            out_c = planes[level] if last else peer.buf[(s + 1) & 1]
            band_scale_p2p(peer.ptrs(s & 1), y0s, rank, out_c, planes[s], width, width, s, sf.taps_code, band.dtype,
                           band.device)
        if TRACE_EVENTS is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            TRACE_EVENTS.append(ev)
        peer.barrier()  # nobody may refill the buffers (next call) while a neighbour still reads them
        return planes

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
        factors = None
        if self.bilateral is not None:
            from .wavelets import AtrousTransform
            factors = AtrousTransform(self.scaling_function_class, self.bilateral, self.bilateral_scaling).var_factors(level)
        if self.push and world > 1 and factors is None:
            min_rows = min(band_range(global_height, k, world)[1] - band_range(global_height, k, world)[0]
                           for k in range(world))
            if halo_rows(level - 1, n_taps) <= min_rows:  # single-hop halos; otherwise the exchange below (multi-hop)
                return self._call_push(band, planes, level, global_height, sf, rank, world)
        if self.p2p and world > 1 and factors is None:
            return self._call_p2p(band, planes, level, global_height, sf, rank, world)
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
            if factors is None:
                self.scale_fn(cur, pad, out_c, out_pad, planes[s], rows, width, global_height, y0, s, sf.taps_code)
            else:  # same halo (the bilateral gather reads the same taps), the range-weighted kernel on the band
                self.scale_fn(cur, pad, out_c, out_pad, planes[s], rows, width, global_height, y0, s, sf.taps_code,
                              var_factor=factors[s])
        return planes


# ---------------------------------------------------------------------------------------------------------------
# WOW of ONE image sharded by row bands ("WOW on bands additionally needs tiny all-reduces")
# ---------------------------------------------------------------------------------------------------------------
class _CudaWowBackend:
    """Per-band arithmetic of BandedWow on the device (the gloo tests substitute an oracle backend)."""

    scale = staticmethod(_cuda_band_scale)

    @staticmethod
    def whiten(ext_w, pad, out, rows, width, height, y0, scale, taps_code, sig_mode, sigma, sigma_e, noise, weight):
        lib = _lib.load(require_cuda=True)
        with torch.cuda.device(ext_w.device):
            _lib.check(lib.wb_wow_whiten_scale_band(ext_w.data_ptr(), out.data_ptr(), rows, width, height, y0, pad,
                                                    ext_w.stride(0), 0, out.stride(0), scale, taps_code,
                                                    _lib.dtype_code(ext_w.dtype), sig_mode, float(sigma), float(sigma_e),
                                                    float(noise), None, float(weight), _lib.stream_ptr(ext_w.device)))

    @staticmethod
    def moments(plane):
        """(count, mean, population variance) of this rank's rows, float64 tensor on the plane's device."""
        from .wavelets import plane_moments
        m = plane_moments(plane)[0]
        return torch.stack([torch.tensor(float(plane.numel()), dtype=torch.float64, device=plane.device), m[0], m[1]])

    @staticmethod
    def synthesis(planes):
        from .wavelets import synthesis
        return synthesis(planes)


def distributed_abs_median(x: torch.Tensor, group=None) -> torch.Tensor:
    """Exact ``np.median(np.abs(X))`` of a tensor X whose elements are spread over the ranks of ``group`` (this rank
    holds ``x``), as a 0-dim tensor of x's dtype -- bit-identical to NumPy, including the mean of the two middle values
    for an even count.

    Order statistics of |x| are order statistics of the IEEE bit patterns of |x| read as integers, so this is a
    radix select on those integers, 11 bits per pass from the top: every rank histograms the digit of the keys that
    match the prefix found so far, one all-reduce of 2048 counts per pass locates the bin that holds the wanted rank
    (3 passes for float32, 6 for float64; twice that when the two middle ranks of an even count part ways).  Plumbing
    only (torch ops + ``all_reduce``): the MAD estimate is computed once per image."""
    if x.dtype == torch.float32:
        keys, bits = x.detach().abs().contiguous().view(torch.int32).reshape(-1).to(torch.int64), 31
    elif x.dtype == torch.float64:
        keys, bits = x.detach().abs().contiguous().view(torch.int64).reshape(-1), 63
    else:
        raise TypeError(f"distributed_abs_median: float32 or float64, got {x.dtype}")
    distributed = dist.is_initialized() and dist.get_world_size(group) > 1
    n = torch.tensor([keys.numel()], dtype=torch.int64, device=x.device)
    if distributed:
        dist.all_reduce(n, group=group)
    n = int(n.item())
    if n == 0:
        return torch.full((), float("nan"), dtype=x.dtype, device=x.device)

    def select(k):
        prefix, below, remaining = 0, 0, bits
        while remaining > 0:
            b = min(11, remaining)
            shift = remaining - b
            sel = keys if remaining == bits else keys[(keys >> remaining) == prefix]
            digit = (sel >> shift) & ((1 << b) - 1)
            hist = torch.bincount(digit, minlength=1 << b)
            if distributed:
                dist.all_reduce(hist, group=group)
            cum = torch.cumsum(hist, 0)
            bin_ = int(torch.searchsorted(cum, torch.tensor(k - below, dtype=torch.int64, device=cum.device),
                                          right=True).item())
            below += int(cum[bin_ - 1].item()) if bin_ > 0 else 0
            prefix = (prefix << b) | bin_
            remaining = shift
        return prefix

    k_lo, k_hi = (n - 1) // 2, n // 2
    key_lo = select(k_lo)
    key_hi = key_lo if k_hi == k_lo else select(k_hi)
    pair = torch.tensor([key_lo, key_hi], dtype=torch.int64, device=x.device)
    vals = pair.to(torch.int32).view(torch.float32) if x.dtype == torch.float32 else pair.view(torch.float64)
    return (vals[0] + vals[1]) / 2


class BandedWow:
    """``wow(image)`` (watroo/utils.py:105-219, plain or bilateral cascade, whitening on) of ONE image sharded by row
    bands.

    Per scale: halo exchange of ``c_s`` -> band scale kernel (``c_{s+1}``, raw ``w_s``) -> halo exchange of the raw
    ``w_s`` (the local power is a second dilated filter) -> band whitening kernel.  The residual plane needs the
    population std of the WHOLE plane: every rank contributes (count, mean, variance) of its rows, combined after one
    all-gather of three doubles.  The synthesis sum is local.  When thresholds (``denoise_coefficients``) are requested
    without ``noise``, the MAD estimate is the exact median of |w_0| over the whole image (``distributed_abs_median``:
    a radix select with one 2048-bin all-reduce per pass).  Returns ``(recon_band, planes_band)``; the planes equal the two-pass single-device route bit for bit, the
    residual plane up to the rounding of the combined std."""

    def __init__(self, scaling_function_class=B3spline, group=None, backend=None, poison=False):
        self.scaling_function_class = scaling_function_class
        self.group = group
        self.backend = backend or _CudaWowBackend
        self.poison = poison

    def __call__(self, band, global_height, n_scales=None, weights=(), denoise_coefficients=(), noise=None,
                 soft_threshold=True, bilateral=None, bilateral_scaling=False):
        from .utils import _wow_plan
        from .wavelets import AtrousTransform
        sf = self.scaling_function_class(2)
        n_taps = len(sf.coefficients_1d)
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        y0, y1 = band_range(global_height, rank, world)
        rows, width = band.shape
        assert rows == y1 - y0, f"rank {rank}: band has {rows} rows, expected {y1 - y0}"
        level, sigma_bilateral, wts, dns = _wow_plan((global_height, width), self.scaling_function_class, n_scales,
                                                     list(weights), list(denoise_coefficients), bilateral)
        sigma_e = sf.sigma_e(bilateral=sigma_bilateral)  # the bilateral table when the cascade is bilateral
        factors = None
        if sigma_bilateral is not None:
            factors = AtrousTransform(self.scaling_function_class, sigma_bilateral, bilateral_scaling).var_factors(level)
        be = self.backend
        planes = torch.empty((level + 1, rows, width), dtype=band.dtype, device=band.device)
        pad = min(halo_rows(max(level - 1, 0), n_taps), global_height) if world > 1 else 0
        fill = float("nan") if self.poison else 0.0
        ext = [torch.full((rows + 2 * pad, width), fill, dtype=band.dtype, device=band.device) for _ in range(2)]
        wext = torch.full((rows + 2 * pad, width), fill, dtype=band.dtype, device=band.device)
        ext[0][pad:pad + rows].copy_(band)
        if level == 0:
            planes[0].copy_(band)
        for s in range(level):
            cur, nxt = ext[s & 1], ext[(s + 1) & 1]
            halo = min(halo_rows(s, n_taps), pad)
            if world > 1:
                exchange_halos(cur, pad, global_height, halo, self.group)
            last = s == level - 1
            out_c, out_pad = (planes[level], 0) if last else (nxt, pad)
            if factors is None:
                be.scale(cur, pad, out_c, out_pad, wext[pad:pad + rows], rows, width, global_height, y0, s, sf.taps_code)
            else:
                be.scale(cur, pad, out_c, out_pad, wext[pad:pad + rows], rows, width, global_height, y0, s, sf.taps_code,
                         var_factor=factors[s])
            d = dns[s]
            if d != 0 and noise is None:
                # MAD estimate over the WHOLE image, lazily at the first scale that thresholds (watroo/wavelets.py:126-127,
                # :131-132): from the raw w_0 when that is scale 0, else from plane 0 as it stands by then -- already
                # whitened (the quirk the reference's 'tri' golden with [0, 3] pins; wow()/_wow_stack do the same).
                # NumPy >= 2 promotion: '/ 0.6745' in the plane dtype, '/ sigma_e[0]' in float64.
                med = distributed_abs_median(wext[pad:pad + rows] if s == 0 else planes[0], self.group)
                noise = float((med / torch.tensor(0.6745, dtype=med.dtype)).item()) / float(sigma_e[0])
            if world > 1:
                exchange_halos(wext, pad, global_height, halo, self.group)
            mode = (1 if soft_threshold else 2) if d != 0 else 0
            be.whiten(wext, pad, planes[s], rows, width, global_height, y0, s, sf.taps_code, mode, d,
                      sigma_e[s] if d != 0 else 1.0, noise if d != 0 else 0.0, wts[s])
        # residual plane: c_L *= wt_L / std(c_L) over the WHOLE plane (utils.py:185-189, :203)
        mine = be.moments(planes[level])
        if world > 1:
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine, group=self.group)
        else:
            parts = [mine]
        stat = torch.stack(parts).to(torch.float64)
        n_tot = stat[:, 0].sum()
        mean = (stat[:, 0] * stat[:, 1]).sum() / n_tot
        var = (stat[:, 0] * (stat[:, 2] + (stat[:, 1] - mean) ** 2)).sum() / n_tot  # no E[x^2] - mean^2 cancellation
        sd = torch.sqrt(torch.clamp(var, min=0)).to(band.dtype)
        sd = torch.where(sd <= 0, torch.full_like(sd, 1e-15), sd)
        planes[level].mul_((torch.tensor(wts[level], dtype=band.dtype, device=sd.device) / sd).to(planes.device))
        return be.synthesis(planes), planes
