"""Where the time of the halo-push row-band cascade goes (N GPUs of one box, under torchrun):

    torchrun --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 tools/trace_push.py [--side 32768 --levels 12]

(1) cost of one device-side barrier of the symmetric-memory group, (2) device time of every scale of
BandedTransform(push=True) (CUDA events around each scale's barrier + kernel, max over ranks), with the HBM and NVLink
time each scale would take at the measured peaks beside it, (3) what plain peer copies of a band to both neighbours
reach (every rank copying at once): the NVLink ceiling the pushes compete with."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wavelets_b200 as wb  # noqa: E402
from wavelets_b200 import sharded  # noqa: E402
from wavelets_b200.sharded import BandedTransform, band_range, halo_rows  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=32768)
    ap.add_argument("--levels", type=int, default=12)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    side, L = args.side, args.levels
    y0, y1 = band_range(side, rank, world)
    rows = y1 - y0
    gen = torch.Generator(device=dev).manual_seed(5 + rank)
    band = torch.randn((rows, side), generator=gen, device=dev)
    bt = BandedTransform(wb.B3spline, push=True)
    bt(band, L, side)  # allocates the symmetric buffers
    peer = next(v for k, v in sharded._PEER_BUFFERS.items() if k[0] == "push")
    res = {"n_gpus": world, "side": side, "levels": L}

    def max_over_ranks(x):
        t = torch.tensor(x, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    # (1) barrier alone
    for _ in range(5):
        peer.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        peer.barrier()
    e1.record()
    torch.cuda.synchronize()
    res["barrier_us"] = max_over_ranks([e0.elapsed_time(e1) * 10])[0]

    # (2) per-scale device time (events recorded by the traced cascade)
    sharded.TRACE_EVENTS = []
    per_scale = None
    for rep in range(args.reps):
        sharded.TRACE_EVENTS = []
        dist.barrier()
        bt(band, L, side)
        torch.cuda.synchronize()
        ev = sharded.TRACE_EVENTS
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(len(ev) - 1)]
        per_scale = ms if per_scale is None else [min(a, b) for a, b in zip(per_scale, ms)]
    sharded.TRACE_EVENTS = None
    per_scale = max_over_ranks(per_scale)
    hbm = 3 * 4 * rows * side / 6550.7e9 * 1e3
    res["per_scale_ms"] = per_scale
    res["total_ms"] = sum(per_scale)
    res["hbm_ms_per_scale"] = hbm
    res["push_bytes_per_scale"] = [2 * min(halo_rows(s + 1, 5), rows) * side * 4 if s < L - 1 else 0 for s in range(L)]
    res["link_ms_per_scale_at_770"] = [b / 770e9 * 1e3 for b in res["push_bytes_per_scale"]]

    # (3) plain peer copies of the whole band to both neighbours, every rank at once
    up = peer.peer(rank - 1)[1, :rows] if rank > 0 else None
    dn = peer.peer(rank + 1)[1, :rows] if rank < world - 1 else None
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def copies():
        cur = torch.cuda.current_stream(dev)
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        if up is not None:
            with torch.cuda.stream(s1):
                up.copy_(band, non_blocking=True)
        if dn is not None:
            with torch.cuda.stream(s2):
                dn.copy_(band, non_blocking=True)
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    copies()
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        copies()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = max_over_ranks([e0.elapsed_time(e1) / 3])[0]
    res["peer_copy_band_to_both_neighbours_ms"] = ms
    res["peer_copy_egress_gbs_interior_rank"] = 2 * rows * side * 4 / ms / 1e6
    if rank == 0:
        line = json.dumps(res)
        print(line, flush=True)
        if args.out:
            with open(args.out, "w") as fh:
                fh.write(line + "\n")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
