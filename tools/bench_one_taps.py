#!/usr/bin/env python
"""One-tile attention kernel (Tk <= 128) at the BERT INT8 shape (B=32, H=12, T=128, bthd strides, three taps):
register-resident S row (OEH_ONE_REGS=1) vs chunked TMEM passes (=0).  us per launch, cold pool, CUDA graph."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import _lib, ops
from tools.bench_mid import time_launch
torch.manual_seed(0)
B, T, H = 32, 128, 12
qp = torch.tensor([0.05, 128.0], device="cuda"); pp = torch.tensor([1.0 / 255, 0.0], device="cuda"); cp = torch.tensor([0.01, 128.0], device="cuda")
mk = lambda: torch.randn(B, T, H, 64, device="cuda", dtype=torch.float16).permute(0, 2, 1, 3)
pool = [(mk(), mk(), mk(), torch.empty(B, T, H * 64, device="cuda", dtype=torch.float16)) for _ in range(8)]
for name, kw in (("clippedsoftmax1+3taps", dict(kind=3, gamma=-0.025, eta=1.1, qdq_scores=qp, qdq_probs=pp, qdq_ctx=cp, n_bits=8)),
                 ("softmax1+3taps", dict(kind=1, qdq_scores=qp, qdq_probs=pp, qdq_ctx=cp, n_bits=8)),
                 ("clippedsoftmax1", dict(kind=3, gamma=-0.025, eta=1.1))):
    row = dict(case=name)
    for regs in ("0", "1"):
        _lib.set_override("OEH_ONE_REGS", regs)
        row[f"regs{regs}_us"] = round(time_launch(lambda s: ops.attention_fwd(s[0], s[1], s[2], scale=0.125, out=s[3], **kw), pool), 2)
    _lib.set_override("OEH_ONE_REGS", None)
    print(json.dumps(row), flush=True)
