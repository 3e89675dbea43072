#!/usr/bin/env python
"""us per launch of the residual-add + LayerNorm kernel (oeh_add_layer_norm) at the model shapes, L2-warm as inside a model step
(inputs rotate over a small pool) and cold (pool larger than L2): CUDA graph of back-to-back launches, CUDA events."""
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
    res = {"lib": os.path.basename(os.environ.get("OEH_LIB_PATH", "liboeh.so"))}
    qa = torch.tensor([0.05, 128.0], device="cuda")
    qr = torch.tensor([0.06, 127.0], device="cuda")
    qo = torch.tensor([0.03, 120.0], device="cuda")
    gw, gb = torch.randn(12, 64, device="cuda"), torch.randn(12, device="cuda")
    for rows, B in ((4096, 32), (2048, 4)):
        w, bias = torch.randn(768, device="cuda", dtype=torch.float16), torch.randn(768, device="cuda", dtype=torch.float16)
        for warm in (True, False):
            n_pool = 2 if warm else max(2, (200 << 20) // (rows * 768 * 2 * 4))
            pool = [(torch.randn(B, rows // B, 768, device="cuda", dtype=torch.float16), torch.randn(B, rows // B, 768, device="cuda", dtype=torch.float16),
                     torch.empty(B, rows // B, 768, device="cuda", dtype=torch.float16)) for _ in range(n_pool)]
            for name, kw in (("fp", {}), ("taps_res_out", dict(qparams_res=qr, qparams_out=qo)),
                             ("taps_a_res_out_sum", dict(qparams_a=qa, qparams_res=qr, qparams_out=qo, return_sum=True)),
                             ("fp_sum_gate", dict(return_sum=True, gate=(gw, gb)))):
                run = lambda s: ops.add_layer_norm(s[0], s[1], w, bias, 1e-5, out=s[2], **kw)
                res[f"{rows}x768 {name} {'warm' if warm else 'cold'}"] = round(time_launch(run, pool, reps=max(1, 40 // n_pool)), 2)
            del pool
    print(json.dumps(res))


if __name__ == "__main__":
    main()
