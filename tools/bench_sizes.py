import sys, os, json, torch
sys.path.insert(0, os.getcwd())
import wavelets_b200 as wb
from tools.bench_wow import solar_like_device, timed
res = {}
for n in (512, 1024, 2048, 3072, 4096):
    img = solar_like_device(n, torch.float32)
    tr = wb.AtrousTransform(wb.B3spline)
    L = int(round(__import__('math').log2(n) - __import__('math').log2(5)))
    res[n] = {"L": L, "transform_us": round(1e3 * timed(lambda: tr(img, L), 50), 1), "wow_us": round(1e3 * timed(lambda: wb.wow(img), 30), 1),
              "wow_bil_us": round(1e3 * timed(lambda: wb.wow(img, bilateral=1, denoise_coefficients=[5, 2]), 10), 1)}
    px = n * n
    res[n]["transform_GBs"] = round(12 * px * L / res[n]["transform_us"] / 1e3, 0)
    res[n]["wow_frac_model"] = round((5 * L + 3) * 4 * px / (res[n]["wow_us"] * 1e-6) / 6550.7e9, 3)
print(json.dumps(res))
