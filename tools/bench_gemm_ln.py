#!/usr/bin/env python
"""Linear -> + residual -> LayerNorm: the fused cluster kernel (oeh_linear_add_layer_norm) against the two launches it replaces
(library GEMM -> oeh_add_layer_norm) at the model shapes; us per call, L2-warm pool of 2 input sets, CUDA graph, CUDA events."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_mid import time_launch  # noqa: E402
from outeffhop_b200 import ops  # noqa: E402


def main():
    torch.manual_seed(0)
    qa = torch.tensor([0.05, 128.0], device="cuda")
    qr = torch.tensor([0.06, 127.0], device="cuda")
    qo = torch.tensor([0.03, 120.0], device="cuda")
    for M, N, K in ((4096, 768, 768), (4096, 768, 3072), (2048, 768, 768), (2048, 768, 3072), (4096, 1024, 4096)):
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
        b = torch.randn(N, device="cuda").half()
        gamma, beta = torch.randn(N, device="cuda").half(), torch.randn(N, device="cuda").half()
        pool = [(torch.randn(M, K, device="cuda").half(), torch.randn(M, N, device="cuda").half(),
                 torch.empty(M, N, device="cuda", dtype=torch.float16)) for _ in range(2)]
        for name, kw in (("fp", {}), ("taps+sum", dict(qparams_a=qa, qparams_res=qr, qparams_out=qo, return_sum=True))):
            row = dict(M=M, N=N, K=K, case=name)
            row["fused_us"] = round(time_launch(lambda s: ops.linear_add_layer_norm(s[0], w, b, s[1], gamma, beta, 1e-5, out=s[2], **kw), pool, reps=20), 2)
            row["gemm_then_ln_us"] = round(time_launch(lambda s: ops.add_layer_norm(torch.nn.functional.linear(s[0], w, b), s[1], gamma, beta, 1e-5, out=s[2], **kw), pool, reps=20), 2)
            row["gemm_only_us"] = round(time_launch(lambda s: torch.nn.functional.linear(s[0], w, b), pool, reps=20), 2)
            row["fused_tflops"] = round(2.0 * M * N * K / row["fused_us"] / 1e6, 1)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
