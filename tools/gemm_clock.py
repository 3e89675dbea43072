"""SM clock and us per launch of oeh_linear_fwd (4096 x 3072 x 768) with / without the GELU epilogue: back to back vs one at a time."""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from outeffhop_b200 import ops
M, N, K = 4096, 3072, 768
x = torch.randn(M, K, device="cuda", dtype=torch.float16); w = torch.randn(N, K, device="cuda", dtype=torch.float16) * K ** -0.5
b = torch.randn(N, device="cuda", dtype=torch.float16); y = torch.empty(M, N, device="cuda", dtype=torch.float16)
for name, f in (("cublas", lambda: torch.nn.functional.linear(x, w, b)), ("none", lambda: ops.linear_fwd(x, w, b, out=y)),
                ("gelu", lambda: ops.linear_fwd(x, w, b, out=y, act="gelu")), ("relu", lambda: ops.linear_fwd(x, w, b, out=y, act="relu"))):
    time.sleep(0.3)
    best = min(bench._event_time(torch, f, reps=1, warm=(3 if i == 0 else 0)) for i in range(20))
    row = {"case": name, "single_best_us": round(best * 1e6, 2)}
    for reps in (20, 200, 2000):
        mhz = []
        t = bench._event_time(torch, f, reps=reps, warm=0, clock=mhz)
        row[f"b2b{reps}_us"] = round(t * 1e6, 2); row[f"b2b{reps}_mhz"] = round(mhz[0])
    print(json.dumps(row), flush=True)
