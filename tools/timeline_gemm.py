"""Per-CTA timeline of the persistent GEMM (oeh_linear_fwd) from %globaltimer stamps: per tile, the MMA warp's first issue and last
commit, and an epilogue warp's start (accumulator complete) and end.
Needs the experiment build:  tools/build_variant.sh timeline -DOEH_TIMELINE
Run:  OEH_LIB_PATH=outeffhop_b200/lib/liboeh_timeline.so python tools/timeline_gemm.py [act] [tile_n]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import _lib, ops  # noqa: E402

act = sys.argv[1] if len(sys.argv) > 1 else "gelu"
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 0
M, N, K = 4096, 3072, 768
torch.manual_seed(0)
x = torch.randn(M, K, device="cuda").half()
w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
b = torch.randn(N, device="cuda").half()
y = torch.empty(M, N, device="cuda", dtype=torch.float16)
buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
lib = _lib.lib()
lib.oeh_set_gemm_timeline.argtypes = [ctypes.c_void_p]
for _ in range(5):
    ops.linear_fwd(x, w, b, out=y, act=act, tile_n=tile)
torch.cuda.synchronize()
assert lib.oeh_set_gemm_timeline(buf.data_ptr()) == 0
for rep in range(3):
    buf.zero_()
    ops.linear_fwd(x, w, b, out=y, act=act, tile_n=tile)
    torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(148, 16).astype(np.int64)
t0 = t[t > 0].min()
print(f"act={act} tile_n={tile}: ns since the first stamp; per CTA: tile k -> [mma first issue, mma last commit | epilogue start, epilogue end]")
for cta in (0, 1, 2, 3, 100, 146, 147):
    row = t[cta]
    parts = []
    for k in range(4):
        v = row[4 * k: 4 * k + 4]
        if v.max() == 0:
            continue
        parts.append(f"tile{k}: mma {v[0] - t0 if v[0] else -1}..{v[1] - t0 if v[1] else -1} | epi {v[2] - t0 if v[2] else -1}..{v[3] - t0 if v[3] else -1}")
    print(f"  cta {cta:3d}  " + "   ".join(parts))
end = t.max() - t0
print("last stamp:", end)
