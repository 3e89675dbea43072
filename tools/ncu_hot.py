#!/usr/bin/env python
"""Top stall sites of an ncu source-page CSV (ncu -i x.ncu-rep --page source --csv > x.csv):
    python tools/ncu_hot.py x.csv [N]   -> N hottest SASS instructions with their dominant stall reasons,
plus the sample histogram over 64-instruction windows so hot regions are visible."""
import csv
import sys


def main():
    path = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for li, r in enumerate(rows[2:]):
        if len(r) < len(hdr):
            continue
        samples = int(r[idx["# Samples"]] or 0)
        st = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
        data.append((li, r[idx["Source"]].strip(), samples, int(r[idx["Instructions Executed"]] or 0), st))
    tot = sum(d[2] for d in data)
    print(f"total samples {tot}, instructions {len(data)}")
    for li, src, s, ex, st in sorted(data, key=lambda d: -d[2])[:n]:
        print(f"{li:5d} {100 * s / tot:5.2f}% ex={ex:9d} {src[:70]:70s} " + " ".join(f"{k}:{v}" for v, k in st if v))
    print("--- windows of 64 instructions")
    for w in range(0, len(data), 64):
        s = sum(d[2] for d in data[w:w + 64])
        if s * 100 > tot:
            print(f"{w:5d}-{w + 63:5d} {100 * s / tot:5.1f}%  first: {data[w][1][:60]}")


if __name__ == "__main__":
    main()
