#!/usr/bin/env python
"""Per-phase cycle trace of the warp-specialised attention kernel (needs a library built with -DOEH_TRACE=1:
tools/build_variant.sh trace -DOEH_TRACE=1; OEH_LIB_PATH=.../liboeh_trace.so python tools/trace_ws.py)."""
import argparse, ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import ops, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=8192)
ap.add_argument("--B", type=int, default=8)
ap.add_argument("--kind", type=int, default=1)
a = ap.parse_args()
q = torch.randn(a.B, 12, a.T, 64, device="cuda", dtype=torch.float16)
k, v = torch.randn_like(q), torch.randn_like(q)
for _ in range(2):
    ops.attention_fwd(q, k, v, scale=0.125, kind=a.kind, gamma=-0.025, eta=1.1)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 240)()
h = ctypes.CDLL(_lib.LIB_PATH)
assert h.oeh_trace_read(buf) == 0
names = {"softmax": "top s_full ld max exp pack+arrive", "issuer": "loop kv_full p_full issue"}
for w in range(20):
    r = buf[w * 12:(w + 1) * 12]
    if r[0] == 0:
        continue
    n = r[0]
    role = "softmax" if w < 16 else "issuer"
    print(f"warp {w:2d} {role:8s} steps {n} total/step {r[1] / n:7.1f} | " + " ".join(f"{x / n:6.1f}" for x in r[2:10]) + f"   ({names[role]})")
