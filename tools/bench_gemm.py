#!/usr/bin/env python
"""oeh_linear_fwd (hand-written tcgen05 GEMM + fused epilogue) vs the library path it replaces (cuBLASLt F.linear followed by
the separate activation / fake-quant kernels) on the Linear shapes of BERT-base (M = 4096) and OPT-125m (M = 2048).
us per call, inputs rotating over a pool larger than L2, back-to-back launches in one CUDA graph, CUDA events."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import ops  # noqa: E402


def time_graph(fn, pool, reps=5):
    for s in pool:
        fn(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for s in pool:
                fn(s)
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / (reps * len(pool)))
    return min(ts) * 1e3


def main():
    torch.manual_seed(0)
    dt = torch.float16
    qp = torch.tensor([0.03, 128.0], device="cuda")
    tiles = [int(t) for t in os.environ.get("TILES", "0,128,256,2128,2256").split(",")]
    for (M, N, K, act, tap) in ((4096, 2304, 768, "none", 0), (4096, 2304, 768, "none", 3), (4096, 768, 768, "none", 0),
                                (4096, 3072, 768, "gelu", 0), (4096, 3072, 768, "gelu", 1), (4096, 768, 3072, "none", 0),
                                (2048, 2304, 768, "none", 3), (2048, 3072, 768, "relu", 1), (2048, 768, 3072, "none", 1),
                                (2048, 768, 768, "none", 1)):
        n_pool = max(2, min(24, (260 << 20) // ((M * K + M * N) * 2)))
        pool = [(torch.randn(M, K, device="cuda", dtype=dt), torch.empty(M, N, device="cuda", dtype=dt)) for _ in range(n_pool)]
        w = torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5
        b = torch.randn(N, device="cuda", dtype=dt)
        qps = None if tap == 0 else (qp if tap == 1 else [qp] * tap)

        def lib(s):
            if act == "relu":
                y = torch._addmm_activation(b, s[0], w.t())
            else:
                y = torch.nn.functional.linear(s[0], w, b)
                if act == "gelu":
                    y = ops.gelu(y)
            if tap == 1:
                y = ops.fake_quant_asym(y, qp, 8)
            elif tap > 1:
                y = ops.fake_quant_asym_segments(y, [qp] * tap, 8)
            return y

        row = dict(M=M, N=N, K=K, act=act, tap_segments=tap, gflop=round(2.0 * M * N * K / 1e9, 2))
        row["library_us"] = round(time_graph(lib, pool), 2)
        row["gemm_only_library_us"] = round(time_graph(lambda s: torch.nn.functional.linear(s[0], w, b), pool), 2)
        for t in tiles:
            try:
                us = time_graph(lambda s: ops.linear_fwd(s[0], w, b, act=act, qparams=qps, out=s[1], tile_n=t), pool)
                row[f"oeh_tile{t}_us"] = round(us, 2)
            except Exception as e:   # a tile width that does not divide the segment width
                row[f"oeh_tile{t}_us"] = str(e)[:40]
        best = min([v for k2, v in row.items() if k2.startswith("oeh_tile") and isinstance(v, float)] or [float("nan")])
        row["oeh_best_tflops"] = round(2.0 * M * N * K / best / 1e6, 1)
        row["speedup_vs_library"] = round(row["library_us"] / best, 3)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
