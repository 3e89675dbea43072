import json, sys, torch
sys.path.insert(0, ".")
import bench
from outeffhop_b200 import ops
peaks = bench.load_peaks()
r = bench.attention_sweep(torch, ops, peaks)
for k, v in r.items():
    print(k, json.dumps(v))
# the same T=2048 softmax1 launch timed ONE call at a time after 50 ms of idle
import time
B, T = 32, 2048
q = torch.randn(B, 12, T, 64, device="cuda", dtype=torch.float16); k = torch.randn_like(q); v = torch.randn_like(q)
for reps in (1, 5, 20, 100):
    time.sleep(0.2)
    mhz = []
    t = bench._event_time(torch, lambda: ops.attention_fwd(q, k, v, scale=0.125, kind=1), reps=reps, warm=1, clock=mhz)
    print("T=2048 softmax1 reps", reps, "TF/s", round(4.0 * B * 12 * T * T * 64 / t / 1e12, 1), "MHz", round(mhz[0]))
