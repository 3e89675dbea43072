"""softmax1 attention TF/s at T=2048 / 8192 (config-5 shapes) for the library named by OEH_LIB_PATH (poly-exp2 share experiments:
tools/build_variant.sh polyN -DOEH_POLY_PAIRS=N)."""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from outeffhop_b200 import ops
res = {"lib": os.path.basename(os.environ.get("OEH_LIB_PATH", "liboeh.so"))}
for T, reps in ((2048, 40), (8192, 12)):
    B = 65536 // T
    q = torch.randn(B, 12, T, 64, device="cuda", dtype=torch.float16); k = torch.randn_like(q); v = torch.randn_like(q)
    for causal in (False, True):
        mhz = []
        t = bench._event_time(torch, lambda: ops.attention_fwd(q, k, v, scale=0.125, kind=1, causal=causal), reps=reps, warm=10, clock=mhz)
        fl = 2.0 * B * 12 * T * (T + 1) * 64 if causal else 4.0 * B * 12 * T * T * 64
        res[f"T{T}{'c' if causal else ''}"] = round(fl / t / 1e12, 1)
    del q, k, v
print(json.dumps(res))
