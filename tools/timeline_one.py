"""Per-CTA timeline of the one-tile attention kernel (B=32, H=12, T=128) from %globaltimer stamps.

Needs the experiment build:  tools/build_variant.sh timeline -DOEH_TIMELINE
Run:  OEH_LIB_PATH=outeffhop_b200/lib/liboeh_timeline.so python tools/timeline_one.py [kind]
Stamps: 0 entry, 1 TMEM allocated + CTA barrier, 2 S ready (loads + QK^T), 3 P written, 4 O ready, 5 O in registers + dealloc, 6 stored.
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from outeffhop_b200 import ops, _lib

kind = sys.argv[1] if len(sys.argv) > 1 else "clippedsoftmax1"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
H, T, d = 12, 128, 64
if len(sys.argv) > 3:
    _lib.set_override("OEH_ONE_PERSISTENT", sys.argv[3])
dev = torch.device("cuda")
torch.manual_seed(0)
sets = [[torch.randn(B, T, H, d, device=dev, dtype=torch.float16).transpose(1, 2) for _ in range(3)] for _ in range(8)]
out = torch.empty(B, T, H * d, device=dev, dtype=torch.float16)
buf = torch.zeros(B * H * 8, dtype=torch.int64, device=dev)
lib = _lib.lib()
KINDS = {"vanilla": 0, "softmax1": 1, "clipped": 2, "clippedsoftmax1": 3}
kw = dict(scale=0.125, kind=KINDS[kind])
if kind.startswith("clipped"):
    kw.update(gamma=-0.025, eta=1.0)
for i in range(5):
    ops.attention_fwd(*sets[i % 8], out=out, **kw)
torch.cuda.synchronize()
import ctypes
lib.oeh_set_timeline.argtypes = [ctypes.c_void_p]
assert lib.oeh_set_timeline(buf.data_ptr()) == 0
rows = []
for i in range(8):
    buf.zero_()
    ops.attention_fwd(*sets[i], out=out, **kw)
    torch.cuda.synchronize()
    t = buf.cpu().numpy().reshape(B * H, 8).astype(np.int64)
    sm = t[:, 7]
    st = t[:, :7] - t[:, 0].min()
    rows.append(st)
    if i == 7:
        names = ["entry", "alloc'd", "S ready", "P written", "O ready", "dealloc", "stored"]
        print(f"kind {kind}; times in ns since the first CTA's entry; n = {B * H} CTAs")
        for k, n in enumerate(names):
            print(f"  {n:10s} min {st[:, k].min():6d}  median {int(np.median(st[:, k])):6d}  p90 {int(np.percentile(st[:, k], 90)):6d}  max {st[:, k].max():6d}")
        ph = np.diff(st, axis=1)
        for k in range(6):
            print(f"  phase {names[k]:>9s} -> {names[k + 1]:10s} median {int(np.median(ph[:, k])):6d}  p90 {int(np.percentile(ph[:, k], 90)):6d}  max {ph[:, k].max():6d}")
        # order of arrival per SM
        for s in sorted(set(sm.tolist()))[:4]:
            idx = np.where(sm == s)[0]
            print(f"  SM {s}: " + " | ".join(f"cta {j}: " + ",".join(str(int(x)) for x in st[j]) for j in idx))
        per_sm = np.bincount(sm)
        print("  CTAs per SM histogram:", np.bincount(per_sm).tolist())
print("span per launch (ns):", [int(r[:, 6].max()) for r in rows])
