#!/usr/bin/env python
"""Stall-reason totals and SASS size of every launch of an .ncu-rep captured with --import-source on:
    python tools/ncu_stalls.py gpurun_out/x.ncu-rep     (run where ncu is installed; no GPU needed)
One line per launch: number of SASS instructions (code footprint = 16 B each), warp samples, top stall reasons in %."""
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    i = 0
    while True:
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--launch-skip", str(i), "--launch-count", "1"],
                             capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3 or rows[0][0] != "Kernel Name":
            break
        end = next((k for k in range(2, len(rows)) if rows[k] and rows[k][0] == "Kernel Name"), len(rows))   # SASS view only
        hdr = rows[1]
        idx = {h: j for j, h in enumerate(hdr)}
        cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = dict.fromkeys(cols, 0)
        n = 0
        for r in rows[2:end]:
            if len(r) < len(hdr):
                continue
            n += 1
            for c in cols:
                tot[c] += int(r[idx[c]] or 0)
        s = max(sum(tot.values()), 1)
        top = "  ".join(f"{c[6:]} {100 * v / s:.1f}" for c, v in sorted(tot.items(), key=lambda x: -x[1])[:7])
        print(f"launch {i}: {rows[0][1][:90]}\n    SASS instructions {n} ({n * 16 // 1024} KB)   samples {s}   stalls %: {top}")
        i += 1


if __name__ == "__main__":
    main()
