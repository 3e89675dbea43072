#!/usr/bin/env python
"""Drive the standalone quantizer / range / LayerNorm kernels once each at a bandwidth-bound size, for
    ncu --set full --clock-control none --import-source on -k regex:'fake_quant_kernel|minmax_stage1|add_layer_norm_kernel' ...
(fake-quant = N*(size_in+size_out) algorithmic bytes, min/max = N*size_in, add+LayerNorm = rows*cols*3*size)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import ops  # noqa: E402

n = 1 << 28
for dt in (torch.float32, torch.float16):
    x = torch.randn(n, device="cuda", dtype=dt)
    y = torch.empty_like(x)
    mm = torch.empty(2, device="cuda")
    st = ops.new_qstate("cuda")
    for _ in range(2):
        ops.minmax(x, out=mm)
        ops.range_update(st, mm, 0.9, 8, 1e-8)
        ops.fake_quant_asym(x, st[4:6], 8, out=y)
    torch.cuda.synchronize()
    del x, y
rows, cols = 262144, 768
a = torch.randn(rows, cols, device="cuda", dtype=torch.float16)
b = torch.randn_like(a)
w, bias = torch.ones(cols, device="cuda", dtype=torch.float16), torch.zeros(cols, device="cuda", dtype=torch.float16)
out = torch.empty_like(a)
for _ in range(2):
    ops.add_layer_norm(a, b, w, bias, 1e-12, out=out)
torch.cuda.synchronize()
print("ok")
