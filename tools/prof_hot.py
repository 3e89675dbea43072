#!/usr/bin/env python
"""One profiled launch of every hot kernel of the path at its BASELINE shape (the command `ncu --set full` wraps):
persistent one-pass attention (softmax1, T=2048, B=32), two-pass clipped attention (dense, T=2048), the single-pass
mid-length kernel with the three INT8 taps (B=4, H=12, T=512 causal, config 4), the one-tile kernel at the BERT shape
(B=32, T=128, model strides, clippedsoftmax1) and the tcgen05 GEMM with the erf-GELU epilogue (4096 x 768 -> 3072).
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/hot python tools/prof_hot.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import ops  # noqa: E402


def main():
    torch.manual_seed(0)
    f16 = torch.float16
    cases = []
    mk = lambda B, T: torch.randn(B, 12, T, 64, device="cuda", dtype=f16)
    q, k, v = mk(32, 2048), mk(32, 2048), mk(32, 2048)
    cases.append(lambda: ops.attention_fwd(q, k, v, scale=0.125, kind=1))
    cases.append(lambda: ops.attention_fwd(q, k, v, scale=0.125, kind=1, causal=True))
    cases.append(lambda: ops.attention_fwd(q, k, v, scale=0.125, kind=3, gamma=0.0, eta=1.0))
    mkb = lambda B, T: torch.randn(B, T, 12, 64, device="cuda", dtype=f16).permute(0, 2, 1, 3)
    q4, k4, v4 = mkb(4, 512), mkb(4, 512), mkb(4, 512)
    g4 = torch.randn(4, 12, 512, device="cuda")
    o4 = torch.empty(4, 512, 768, device="cuda", dtype=f16)
    qp = torch.tensor([0.05, 128.0], device="cuda")
    pp = torch.tensor([1.0 / 255, 0.0], device="cuda")
    cp = torch.tensor([0.01, 128.0], device="cuda")
    cases.append(lambda: ops.attention_fwd(q4, k4, v4, scale=0.125, causal=True, gate_logits=g4, out=o4, kind=1, qdq_scores=qp,
                                           qdq_probs=pp, qdq_ctx=cp, n_bits=8, ctx_qdq_before_gate=True))
    cases.append(lambda: ops.attention_fwd(q4, k4, v4, scale=0.125, causal=True, gate_logits=g4, out=o4, kind=0))
    q1, k1, v1 = mkb(32, 128), mkb(32, 128), mkb(32, 128)
    o1 = torch.empty(32, 128, 768, device="cuda", dtype=f16)
    cases.append(lambda: ops.attention_fwd(q1, k1, v1, scale=0.125, kind=3, gamma=-0.025, eta=1.1, out=o1))
    M, N, K = 4096, 3072, 768
    x = torch.randn(M, K, device="cuda", dtype=f16)
    w = torch.randn(N, K, device="cuda", dtype=f16) * K ** -0.5
    b = torch.randn(N, device="cuda", dtype=f16)
    y = torch.empty(M, N, device="cuda", dtype=f16)
    cases.append(lambda: ops.linear_fwd(x, w, b, out=y, act="gelu"))
    for _ in range(3):
        for c in cases:
            c()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for c in cases:
        c()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
