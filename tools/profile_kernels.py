"""Launch each kernel family a few times at one scale so that `ncu -k regex:<name>` can capture it (GPU box only).

    ncu --set full --clock-control none --import-source on -k regex:wow_rows -s 2 -c 1 -o gpurun_out/prof_fused \
        python tools/profile_kernels.py --which fused --scale 3
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wavelets_b200 as wb  # noqa: E402
from wavelets_b200 import _lib, utils  # noqa: E402
from wavelets_b200.wavelets import abs_median_noise, atrous_scale, plane_moments, synthesis  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="all")
    ap.add_argument("--scale", type=int, default=3)
    ap.add_argument("--side", type=int, default=4096)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    lib = _lib.load(require_cuda=True)
    tdt = torch.float32 if args.dtype == "f32" else torch.float64
    n, s = args.side, args.scale
    sf = wb.B3spline(2)
    src = (torch.randn((1, n, n), device="cuda", dtype=tdt) * 10 + 100)
    c, w, o = torch.empty_like(src), torch.empty_like(src), torch.empty_like(src)
    nz = utils._Noise(dev=torch.tensor([1.0], dtype=torch.float64, device="cuda"))
    which = args.which.split(",")

    def want(name):
        return "all" in which or name in which

    for _ in range(args.reps):
        if want("k1"):
            atrous_scale(src, s, sf, out_c=c, out_w=w)
        if want("k2"):
            atrous_scale(src, s, sf, out_c=c, out_w=w, var_factor=1.0)
        if want("k3"):
            atrous_scale(src, s, sf, out_c=c, out_w=w)
            utils._whiten_scale(lib, w, o, s, sf, 0, 0.0, 1.0, utils._Noise(), 1.0)
        if want("k3soft"):
            atrous_scale(src, s, sf, out_c=c, out_w=w)
            utils._whiten_scale(lib, w, o, s, sf, 1, 5.0, 0.2, nz, 1.0)
        if want("fused"):
            utils._wow_scale_fused(lib, src, c, o, s, sf, 0, 0.0, 1.0, utils._Noise(), 1.0)
        if want("fusedsoft"):
            utils._wow_scale_fused(lib, src, c, o, s, sf, 1, 5.0, 0.2, nz, 1.0)
        if want("median"):
            abs_median_noise(src, 0.89)
        if want("moments"):
            plane_moments(src)
        if want("synthesis"):
            synthesis(torch.cat([src, c, w, o]))
        if want("transform"):
            wb.AtrousTransform(wb.B3spline)(src[0], 10)
        if want("wow"):
            wb.wow(src[0])
        if want("wowbil"):
            wb.wow(src[0], bilateral=1, denoise_coefficients=[5, 2])
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
