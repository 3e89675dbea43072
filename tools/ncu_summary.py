#!/usr/bin/env python
"""Curated summary of an .ncu-rep (run where ncu is installed; no GPU needed):
    python tools/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg",
    "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    path = sys.argv[1]
    hdr, units, launches = raw(path)
    for li, vals in enumerate(launches):
        d = dict(zip(hdr, zip(units, vals)))
        print(f"--- launch {li}: {d.get('Kernel Name', ('', '?'))[1][:100]}  grid {d.get('Grid Size', ('', ''))[1]} block {d.get('Block Size', ('', ''))[1]}")
        for k in KEYS:
            if k in d:
                print(f"  {k:75s} {d[k][1]:>16s} {d[k][0]}")
        st = sorted(((float(v[1].replace(',', '')), k) for k, v in d.items()
                     if "issue_stalled" in k and k.endswith("_per_warp_active.pct") and v[1]), reverse=True)
        for v, k in st[:8]:
            print(f"  stall {k.split('issue_stalled_')[1].split('_per')[0]:40s} {v:8.2f} % of warp-active cycles")


if __name__ == "__main__":
    main()
