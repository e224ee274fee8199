"""BASELINE cfg2, second half: ScalingFunction.compute_noise_weights(10) at the reference's size (n_trials = 100 fp32
N(0,1) fields of side 11 * 2**10 = 11264, 10 scales each; watroo/wavelets.py:221-229), entirely on the device.

    python tools/bench_noise_weights.py [--scales 10 --trials 100 --out gpurun_out/noise_weights.json]

Reports the wall time (the call returns a host array, so it includes the final synchronisation), the implied
Mpixel*scales/s and the deviation from the reference's published sigma_e_2d table (a statistical known-answer test:
the table itself was produced by this Monte-Carlo procedure)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wavelets_b200 as wb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scales", type=int, default=10)
    ap.add_argument("--trials", type=int, default=100)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    res = {}
    for sf in (wb.B3spline, wb.Triangle):
        f = sf(2)
        f.compute_noise_weights(args.scales, n_trials=2, seed=1)  # warm-up (allocator, kernel attributes)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        w = f.compute_noise_weights(args.scales, n_trials=args.trials, seed=0)
        dt = time.perf_counter() - t0
        side = len(f.sigma_e_1d) * 2 ** args.scales
        table = f.sigma_e_2d[:args.scales]
        res[f.name] = {"seconds": dt, "side": side, "trials": args.trials, "scales": args.scales,
                       "mpx_scales_per_s": side * side * args.scales * args.trials / dt / 1e6,
                       "weights": [float(x) for x in w],
                       "max_rel_dev_from_sigma_e_2d": float(np.max(np.abs(w - table) / table))}
        print(f.name, json.dumps(res[f.name]), flush=True)
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
