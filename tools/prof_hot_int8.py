#!/usr/bin/env python
"""One profiled launch of the hot kernels of the INT8 steps at their model shapes (the command `ncu --set full` wraps):
two-threads-per-row one-tile attention with the three taps (BERT INT8: B=32, H=12, T=128, clippedsoftmax1), the tcgen05 GEMM
with per-segment taps (q|k|v projection) and with erf GELU + tap (FFN1), residual add + LayerNorm with taps (BERT: a + res + out
taps; OPT: + sum_out), and the standalone fake-quant kernel.
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/hot_int8 python tools/prof_hot_int8.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import ops  # noqa: E402


def main():
    torch.manual_seed(0)
    f16 = torch.float16
    qp = torch.tensor([0.05, 128.0], device="cuda")
    pp = torch.tensor([1.0 / 255, 0.0], device="cuda")
    cp = torch.tensor([0.01, 128.0], device="cuda")
    mkb = lambda B, T: torch.randn(B, T, 12, 64, device="cuda", dtype=f16).permute(0, 2, 1, 3)
    q1, k1, v1 = mkb(32, 128), mkb(32, 128), mkb(32, 128)
    o1 = torch.empty(32, 128, 768, device="cuda", dtype=f16)
    cases = [lambda: ops.attention_fwd(q1, k1, v1, scale=0.125, kind=3, gamma=-0.025, eta=1.1, out=o1, qdq_scores=qp, qdq_probs=pp,
                                       qdq_ctx=cp, n_bits=8)]
    M, K = 4096, 768
    x = torch.randn(M, K, device="cuda", dtype=f16)
    wq = torch.randn(2304, K, device="cuda", dtype=f16) * K ** -0.5
    bq = torch.randn(2304, device="cuda", dtype=f16)
    yq = torch.empty(M, 2304, device="cuda", dtype=f16)
    w1 = torch.randn(3072, K, device="cuda", dtype=f16) * K ** -0.5
    b1 = torch.randn(3072, device="cuda", dtype=f16)
    y1 = torch.empty(M, 3072, device="cuda", dtype=f16)
    cases.append(lambda: ops.linear_fwd(x, wq, bq, out=yq, qparams=[qp, qp, qp]))
    cases.append(lambda: ops.linear_fwd(x, w1, b1, out=y1, act="gelu", qparams=qp))
    a, r = torch.randn(32, 128, 768, device="cuda", dtype=f16), torch.randn(32, 128, 768, device="cuda", dtype=f16)
    o = torch.empty_like(a)
    w, bias = torch.randn(768, device="cuda", dtype=f16), torch.randn(768, device="cuda", dtype=f16)
    cases.append(lambda: ops.add_layer_norm(a, r, w, bias, 1e-12, out=o, qparams_a=qp, qparams_res=cp, qparams_out=qp))
    cases.append(lambda: ops.add_layer_norm(a, r, w, bias, 1e-12, out=o))
    cases.append(lambda: ops.fake_quant_asym(a, qp, 8, out=o))
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
