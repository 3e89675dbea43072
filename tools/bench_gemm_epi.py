#!/usr/bin/env python
"""Epilogue cost of oeh_linear_fwd at the BERT FFN1 / QKV shapes: {none, gelu, gelu+tap, tap} x {stores on, off} in-graph."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import ops
from tools.bench_gemm import time_graph
torch.manual_seed(0)
qp = torch.tensor([0.03, 128.0], device="cuda")
for (M, N, K) in ((4096, 3072, 768), (4096, 2304, 768)):
    pool = [(torch.randn(M, K, device="cuda", dtype=torch.float16), torch.empty(M, N, device="cuda", dtype=torch.float16)) for _ in range(12)]
    w = torch.randn(N, K, device="cuda", dtype=torch.float16) * K ** -0.5
    b = torch.randn(N, device="cuda", dtype=torch.float16)
    row = dict(M=M, N=N, K=K, cublas_us=round(time_graph(lambda s: torch.nn.functional.linear(s[0], w, b), pool), 2))
    for tile in (2256, 2192, 192):
        for name, kw in (("none", {}), ("relu", dict(act="relu")), ("gelu", dict(act="gelu")), ("tap", dict(qparams=qp)), ("gelu_tap", dict(act="gelu", qparams=qp))):
            for dbg in (0, 1):
                row[f"t{tile}_{name}{'_nostore' if dbg else ''}"] = round(time_graph(lambda s: ops.linear_fwd(s[0], w, b, out=s[1], tile_n=tile + 10000 * dbg, **kw), pool), 2)
    print(json.dumps(row), flush=True)
