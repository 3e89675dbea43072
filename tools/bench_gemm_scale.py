#!/usr/bin/env python
"""Per-tile vs fixed cost of oeh_linear_fwd: the same (N, K) at growing M (us per call, warm L2 where it fits)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import ops

def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

for (N, K) in ((2304, 768), (3072, 768), (768, 3072)):
    w = torch.randn(N, K, device="cuda", dtype=torch.float16) * K ** -0.5
    b = torch.randn(N, device="cuda", dtype=torch.float16)
    for M in (4096, 16384, 65536):
        x = torch.randn(M, K, device="cuda", dtype=torch.float16); y = torch.empty(M, N, device="cuda", dtype=torch.float16)
        row = dict(M=M, N=N, K=K, gflop=round(2.0 * M * N * K / 1e9, 1))
        row["cublas_us"] = round(t(lambda: torch.nn.functional.linear(x, w, b)), 1)
        for tile in (256, 2256, 2128, 128):
            row[f"tile{tile}_us"] = round(t(lambda: ops.linear_fwd(x, w, b, out=y, tile_n=tile)), 1)
        row["gelu_tile2256_us"] = round(t(lambda: ops.linear_fwd(x, w, b, act="gelu", out=y, tile_n=2256)), 1)
        best = min(row["tile256_us"], row["tile2256_us"], row["tile2128_us"], row["tile128_us"])
        row["best_tflops"] = round(2.0 * M * N * K / best / 1e6, 1); row["cublas_tflops"] = round(2.0 * M * N * K / row["cublas_us"] / 1e6, 1)
        print(json.dumps(row), flush=True)
