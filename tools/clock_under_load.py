#!/usr/bin/env python
"""SM clock / power while the fused attention kernel runs back to back for a few seconds (nvidia-smi -lms 100)."""
import argparse, os, subprocess, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=8192)
ap.add_argument("--B", type=int, default=8)
ap.add_argument("--kind", type=int, default=1)
ap.add_argument("--seconds", type=float, default=3.0)
a = ap.parse_args()
q = torch.randn(a.B, 12, a.T, 64, device="cuda", dtype=torch.float16)
k, v = torch.randn_like(q), torch.randn_like(q)
out = torch.empty(a.B, a.T, 768, device="cuda", dtype=torch.float16)
f = lambda: ops.attention_fwd(q, k, v, scale=0.125, kind=a.kind, gamma=-0.025, eta=1.1, out=out)
f(); torch.cuda.synchronize()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                     stdout=subprocess.PIPE, text=True)
time.sleep(0.5)
t0 = time.time(); n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < a.seconds:
    for _ in range(20):
        f()
    n += 20
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
time.sleep(0.3)
p.terminate()
rows = [l.strip() for l in p.stdout.read().splitlines() if l.strip()]
fl = 4.0 * a.B * 12 * a.T * a.T * 64
print(f"T={a.T} B={a.B} kind={a.kind}: {ms:.4f} ms/launch sustained, {fl / ms / 1e9:.1f} TF/s over {n} launches")
print("clocks.sm MHz, power W, sw_power_cap:")
print("  " + " | ".join(rows))
