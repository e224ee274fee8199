"""Regenerate every sigma_e table of the reference on the GPU (watroo/wavelets.py:221-229,
tables :245-258 and :274-287) and compare with the published constants -- including the 11th entry of the B3spline
2-D bilateral table, which the reference's table lacks (watroo/wavelets.py:280-281 has 10).

    python tools/regen_sigma_e.py [--trials 100] [--out profiles/r2_sigma_e_tables.json]

Every table: n_trials fp32 N(0,1) fields of side len(sigma_e_1d) * 2**n (device Philox RNG), the transform, the
population std of each detail plane, mean over trials -- exactly compute_noise_weights().  The bilateral tables use
bilateral=1 (the value the reference's own wow(bilateral=1) examples use)."""
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
    ap.add_argument("--trials", type=int, default=100)
    ap.add_argument("--out", default="gpurun_out/sigma_e_tables.json")
    args = ap.parse_args()
    res = {"trials_2d": args.trials, "how": "ScalingFunction.compute_noise_weights on the device (Philox N(0,1) fields)"}
    for sf in (wb.B3spline, wb.Triangle):
        for nd, n, bil, trials in ((2, 11, None, args.trials), (2, 11, 1, args.trials),
                                   (1, 11, None, 4000), (3, 5, None, max(4, args.trials // 10))):
            f = sf(nd)
            t0 = time.perf_counter()
            w = f.compute_noise_weights(n, n_trials=trials, bilateral=bil, seed=2026)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            table = f.sigma_e(bilateral=bil)
            m = min(len(table), len(w))
            key = f"{f.name}_{nd}d" + ("_bilateral" if bil is not None else "")
            res[key] = {"n_scales": n, "trials": trials, "side": len(f.sigma_e_1d) * 2 ** n, "seconds": dt,
                        "regenerated": [float(x) for x in w], "reference_table": [float(x) for x in table],
                        "reference_entries": len(table),
                        "max_rel_dev_over_common_entries": float(np.max(np.abs(w[:m] - table[:m]) / table[:m]))}
            print(key, json.dumps(res[key]), flush=True)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
