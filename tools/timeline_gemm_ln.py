"""Per-CTA timeline of the fused Linear -> + residual -> LayerNorm cluster kernel from %globaltimer stamps.
Needs the experiment build:  tools/build_variant.sh timeline -DOEH_TIMELINE
Run:  OEH_LIB_PATH=outeffhop_b200/lib/liboeh_timeline.so python tools/timeline_gemm_ln.py [M N K]
Stamps: 0 entry, 1 barriers + TMEM + first cluster barrier, 2 accumulator complete, 3 row sums written, 4 squared deviations written,
5 y staged, 6 stores done, 7 last cluster barrier."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import _lib, ops  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (4096, 768, 768)
torch.manual_seed(0)
x = torch.randn(M, K, device="cuda").half()
w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
b, g, be = (torch.randn(N, device="cuda").half() for _ in range(3))
res = torch.randn(M, N, device="cuda").half()
n_cta = (N // 192 if N % 192 == 0 else N // 256) * ((M + 127) // 128)
buf = torch.zeros(n_cta * 8, dtype=torch.int64, device="cuda")
lib = _lib.lib()
lib.oeh_set_ln_timeline.argtypes = [ctypes.c_void_p]
for _ in range(5):
    ops.linear_add_layer_norm(x, w, b, res, g, be, 1e-5)
torch.cuda.synchronize()
assert lib.oeh_set_ln_timeline(buf.data_ptr()) == 0
for rep in range(3):
    buf.zero_()
    ops.linear_add_layer_norm(x, w, b, res, g, be, 1e-5)
    torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(n_cta, 8).astype(np.int64)
st = t - t[:, 0].min()
names = ["entry", "setup", "acc done", "sums", "ssq", "y staged", "stored", "exit"]
print(f"M={M} N={N} K={K}: {n_cta} CTAs; ns since the first CTA's entry")
for k, n in enumerate(names):
    print(f"  {n:9s} min {st[:, k].min():6d}  median {int(np.median(st[:, k])):6d}  max {st[:, k].max():6d}")
