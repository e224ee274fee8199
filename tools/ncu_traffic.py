"""Summarise an ncu --csv log of per-launch DRAM traffic taken INSIDE a running cascade (no cache flush between launches).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
        --cache-control none --clock-control none -k regex:<kernel> -s <warm launches> -c <n> --csv \
        --log-file gpurun_out/traffic.csv python tools/profile_kernels.py --which transform --reps 4
    python tools/ncu_traffic.py gpurun_out/traffic.csv [--algorithmic-bytes 201326592] [--out profiles/k1_traffic.json]

Prints one row per launch and the mean over the captured launches; with --out writes the JSON bench.py reads for
`roofline.traffic` (DRAM bytes per launch, mean over a whole cascade step)."""
import argparse
import csv
import json
import sys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--algorithmic-bytes", type=float, default=0.0)
    ap.add_argument("--out", default="")
    ap.add_argument("--note", default="")
    args = ap.parse_args()
    rows = []
    with open(args.csv) as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    launches = {}
    for r in csv.DictReader(lines):
        key = r["ID"]
        d = launches.setdefault(key, {"kernel": r["Kernel Name"], "grid": r.get("Grid Size"), "block": r.get("Block Size")})
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ns": 1e-3, "ms": 1e3}.get(unit, 1.0)
        d[r["Metric Name"]] = val * scale
    for key in sorted(launches, key=int):
        d = launches[key]
        rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
        rows.append({"id": int(key), "kernel": d["kernel"][:60], "us": d.get("gpu__time_duration.sum"),
                     "dram_read_mb": rd / 1e6, "dram_write_mb": wr / 1e6, "dram_total_mb": (rd + wr) / 1e6,
                     "l2_hit_pct": d.get("lts__t_sector_hit_rate.pct")})
    for r in rows:
        print(r)
    if not rows:
        sys.exit("no launches found")
    n = len(rows)
    mean = {k: sum(r[k] for r in rows if r[k] is not None) / n for k in ("us", "dram_read_mb", "dram_write_mb", "dram_total_mb", "l2_hit_pct")}
    print("mean over", n, "launches:", mean)
    if args.out:
        out = {"dram_bytes_per_launch": mean["dram_total_mb"] * 1e6, "dram_read_bytes_per_launch": mean["dram_read_mb"] * 1e6,
               "dram_write_bytes_per_launch": mean["dram_write_mb"] * 1e6, "l2_sector_hit_rate_pct": mean["l2_hit_pct"],
               "launches_averaged": n, "us_per_launch_under_ncu": mean["us"],
               "algorithmic_bytes_per_launch": args.algorithmic_bytes or None,
               "how": "ncu --cache-control none --clock-control none, consecutive launches of one cascade step (no flush, no "
                      "replay: the metrics fit one pass)", "note": args.note, "per_launch": rows}
        with open(args.out, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
