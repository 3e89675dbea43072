#!/usr/bin/env python
"""A/B of the one-tile attention kernels at the BERT eval shape (B=32, H=12, T=128; BASELINE configs 1 / 2) and neighbours:
two threads per query row (attn_fwd_one2.cuh, OEH_ONE_SPLIT=1) against one (attn_fwd_one.cuh); us per launch, cold (inputs
rotate over a pool larger than L2), CUDA graph of back-to-back launches, CUDA events."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_mid import time_launch  # noqa: E402
from outeffhop_b200 import _lib, ops  # noqa: E402


def main():
    torch.manual_seed(0)
    qp = torch.tensor([0.05, 128.0], device="cuda")
    pp = torch.tensor([1.0 / 255, 0.0], device="cuda")
    cp = torch.tensor([0.01, 128.0], device="cuda")
    for (B, H, T) in ((32, 12, 128), (8, 12, 128), (16, 12, 128), (64, 12, 128), (128, 12, 128), (32, 12, 64), (32, 16, 100)):
        n_pool = max(2, min(64, (200 << 20) // (B * H * T * 64 * 2 * 4)))
        mk = lambda: torch.randn(B, T, H, 64, device="cuda", dtype=torch.float16).permute(0, 2, 1, 3)
        pool = [(mk(), mk(), mk(), torch.randn(B, H, T, device="cuda"), torch.empty(B, T, H * 64, device="cuda", dtype=torch.float16))
                for _ in range(n_pool)]
        for name, kw in (("softmax1", dict(kind=1)), ("clippedsoftmax1", dict(kind=3, gamma=-0.025, eta=1.0)),
                         ("clippedsoftmax1+3taps", dict(kind=3, gamma=-0.025, eta=1.0, qdq_scores=qp, qdq_probs=pp, qdq_ctx=cp, n_bits=8)),
                         ("softmax1+gate", dict(kind=1, gate=True))):
            gate = kw.pop("gate", False)
            row = dict(B=B, H=H, T=T, case=name)
            run = lambda s: ops.attention_fwd(s[0], s[1], s[2], scale=0.125, gate_logits=s[3] if gate else None, out=s[4], **kw)
            for form in ("1", "0"):
                _lib.set_override("OEH_ONE_SPLIT", form)
                row["two_per_row_us" if form == "1" else "one_per_row_us"] = round(time_launch(run, pool), 2)
            _lib.set_override("OEH_ONE_SPLIT", None)
            alg = B * H * T * 64 * 2 * 4
            row["two_per_row_GBps"] = round(alg / row["two_per_row_us"] / 1e3, 1)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
