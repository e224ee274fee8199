"""Sweep the K1 geometry (consumer threads, vectors/thread, ring depth, segment length) per scale on the GPU box.

    python tools/tune_k1.py [--dtype f32|f64] [--side 4096] [--levels 10] [--out gpurun_out/tune_k1.json]

Coordinate descent: the whole transform is timed (realistic L2 state), one scale's geometry is varied at a time.
"""
import argparse
import itertools
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wavelets_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--side", type=int, default=4096)
    ap.add_argument("--levels", type=int, default=10)
    ap.add_argument("--taps", type=int, default=5)
    ap.add_argument("--reps", type=int, default=15)
    ap.add_argument("--out", default="gpurun_out/tune_k1.json")
    args = ap.parse_args()
    lib = _lib.load(require_cuda=True)
    tdt = torch.float32 if args.dtype == "f32" else torch.float64
    code = _lib.dtype_code(tdt)
    n, L = args.side, args.levels
    img = torch.randn((n, n), device="cuda", dtype=tdt)
    planes = torch.empty((L + 1, n, n), device="cuda", dtype=tdt)
    scratch = torch.empty((2, n, n), device="cuda", dtype=tdt)
    st = torch.cuda.current_stream().cuda_stream

    def run():
        _lib.check(lib.wb_atrous_transform(img.data_ptr(), planes.data_ptr(), scratch.data_ptr(), 1, n, n, n, 0, L,
                                           args.taps, code, st))

    def timed(reps):
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    base = timed(args.reps)
    print(f"default geometry: {base:.4f} ms per transform", flush=True)
    V = 4 if args.dtype == "f32" else 2
    vecs = n // V
    geos = []
    for nt, ng in itertools.product((128, 256, 384, 512), (1, 2)):
        if nt * ng > vecs:
            continue
        geos.append((nt, ng))
    results = {"default_ms": base, "scales": {}}
    best_cfg = {}
    for s in range(L):
        d = 2 ** s
        chain = (n + d - 1) // d
        rows = []
        for (nt, ng), slots, seg in itertools.product(geos, (5, 8, 12), (8, 16, 32, 64, 128, 256)):
            if seg > chain and seg != 8:
                continue
            seg_eff = min(seg, chain)
            lib.wb_tune_k1(s, nt, ng, slots, seg_eff)
            try:
                t = timed(args.reps)
            except RuntimeError as e:  # geometry rejected (shared memory) -> generic kernel; skip
                t = float("inf")
            rows.append({"nt": nt, "ng": ng, "slots": slots, "seg": seg_eff, "ms": t})
        rows.sort(key=lambda r: r["ms"])
        best = rows[0]
        lib.wb_tune_k1(s, best["nt"], best["ng"], best["slots"], best["seg"])
        best_cfg[s] = best
        results["scales"][s] = rows[:12]
        print(f"scale {s}: best {best}", flush=True)
    final = timed(args.reps * 2)
    results["best"] = best_cfg
    results["final_ms"] = final
    print(f"tuned: {final:.4f} ms per transform (default {base:.4f})", flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(results, fh, indent=1)


if __name__ == "__main__":
    main()
