"""Item-boundary timeline of the persistent one-pass attention kernel (attn_fwd_wsp.cuh), CTA 0, from %globaltimer stamps.
Needs the experiment build:  tools/build_variant.sh timeline -DOEH_TIMELINE
Run:  OEH_LIB_PATH=outeffhop_b200/lib/liboeh_timeline.so python tools/timeline_wsp.py [T]
MMA issuer (tile 0): 0 item start, 1 Q landed, 2 first two QK^T issued, 3 / 4 PV of step 0 / 1 issued (P there, previous epilogue read
out), 5 / 6 PV of the last two steps issued.  Softmax stream (tile 0, parity p): 0 S of its first step there, 1 of its second,
2 of its last, 3 last P written, 4 statistics merged, 5 last PV complete, 6 stored."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import _lib, ops
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
B = 65536 // T
q = torch.randn(B, 12, T, 64, device="cuda", dtype=torch.float16); k = torch.randn_like(q); v = torch.randn_like(q)
buf = torch.zeros(8 * 4 * 8, dtype=torch.int64, device="cuda")
lib = _lib.lib()
lib.oeh_set_wsp_timeline.argtypes = [ctypes.c_void_p]
for _ in range(3):
    ops.attention_fwd(q, k, v, scale=0.125, kind=1)
torch.cuda.synchronize()
assert lib.oeh_set_wsp_timeline(buf.data_ptr()) == 0
buf.zero_()
ops.attention_fwd(q, k, v, scale=0.125, kind=1)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(8, 4, 8).astype(np.int64)
t0 = t[t > 0].min()
names = ["mma0", "mma0+", "sm p0", "sm p1"]   # mma0+ (now: epilogue of stream p0): 0 O in registers, 1 accumulators combined, 2 finish_row done
for it in range(6):
    for who in (0, 1, 2, 3):
        print(f"item {it} {names[who]:6s} " + " ".join(f"{(x - t0) if x else -1:7d}" for x in t[it, who, :7]))
