"""Aggregate an `ncu --page source --csv` dump: executed warp-instructions and stall samples by SASS opcode.

    ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_mix.py src.csv [top]
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    ex, smp = collections.Counter(), collections.Counter()
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    stalls = collections.Counter()
    total = 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        sass = r[col["Source"]].strip()
        toks = sass.split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        op = op.rstrip(";")
        if not (r[col["Instructions Executed"]] or "0").isdigit():
            continue  # repeated header of the next captured launch
        n = int(r[col["Instructions Executed"]] or 0)
        ex[op] += n
        total += n
        smp[op] += int(r[col["# Samples"]] or 0)
        for sc in stall_cols:
            stalls[sc] += int(r[col[sc]] or 0)
    print("total warp-instructions", total)
    for op, n in ex.most_common(top):
        print(f"{op:28s} {n:12d} {100.0 * n / total:6.2f}%   samples {smp[op]}")
    tot_s = sum(stalls.values())
    print("stall samples:", ", ".join(f"{k[6:]} {100.0 * v / tot_s:.1f}%" for k, v in stalls.most_common(8)))


if __name__ == "__main__":
    main()
