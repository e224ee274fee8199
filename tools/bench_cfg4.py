"""BASELINE cfg4: a batch of 4096^2 frames, B3spline 8 scales + WOW, frames sharded over the ranks (no collective).

    python tools/bench_cfg4.py [--frames 64 --chunk 16]                       # one GPU
    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/bench_cfg4.py   # N GPUs, frame i -> rank i mod N

Every rank whitens its own frames with wow_batch() in chunks of --chunk frames (one launch per scale for the whole
chunk), device-resident; the time is taken with CUDA events and the maximum over ranks is reported.  Frames: 8 distinct
solar-like frames (bench.solar_like_device) tiled with fresh device noise -- generating 128 GiB on the host would only
measure the host.  Roofline: (5L+3)*4 = 172 algorithmic bytes per pixel."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import wavelets_b200 as wb  # noqa: E402
from wavelets_b200.sharded import frame_shard  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64, help="frames PER RANK (cfg4 has 2048 / 8 = 256)")
    ap.add_argument("--chunk", type=int, default=16)
    ap.add_argument("--side", type=int, default=4096)
    ap.add_argument("--scales", type=int, default=8)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.side
    total = args.frames * world
    mine = list(frame_shard(total, rank, world))
    base = torch.stack([bench.solar_like_device(n, torch.float32, dev, seed=2 + k) for k in range(8)])
    gen = torch.Generator(device=dev).manual_seed(100 + rank)

    def make_chunk(ids):
        fr = base[[i % 8 for i in ids]].clone()
        fr += torch.sqrt(fr.clamp(min=1)) * 0.1 * torch.randn(fr.shape, generator=gen, device=dev)
        return fr

    chunks = [make_chunk(mine[i:i + args.chunk]) for i in range(0, min(len(mine), 2 * args.chunk), args.chunk)]

    def run_all():
        done = 0
        k = 0
        while done < len(mine):
            fr = chunks[k % len(chunks)]
            take = min(fr.shape[0], len(mine) - done)
            wb.wow_batch(fr[:take], n_scales=args.scales)
            done += take
            k += 1

    run_all()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        run_all()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        fps = total / (ms.item() / 1e3)
        algo = (5 * args.scales + 3) * 4 * n * n
        peak = bench.measured_peaks()[0] if hasattr(bench, "measured_peaks") else 6650.0
        line = {"workload": f"cfg4: {total} x {n}^2 fp32 frames, B3spline {args.scales} scales + WOW, frames sharded "
                            f"round-robin over {world} rank(s), chunks of {args.chunk}",
                "n_gpus": world, "frames": total, "ms_total": ms.item(), "frames_per_s": fps,
                "frames_per_s_per_gpu": fps / world, "algorithmic_bytes_per_frame": algo,
                "achieved_gbs_per_gpu": algo * fps / world / 1e9,
                "frac_of_hbm_peak": algo * fps / world / 1e9 / float(peak), "peak_gbs": float(peak)}
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "w") as fh:
                fh.write(json.dumps(line) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
