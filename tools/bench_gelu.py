#!/usr/bin/env python
"""Launch duration of the erf GELU kernel (oeh_gelu_erf) vs ATen's at the BERT-base FFN shape (4096 x 3072 fp16):
  l2   : one 25 MB tensor over and over (the fc1 output is L2-resident when the kernel runs in the model)
  cold : a pool of tensors larger than L2, launches back to back inside one CUDA-event pair
    python tools/bench_gelu.py [--rows 4096] [--cols 3072]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import ops  # noqa: E402


def timed(fn, pool, reps):
    for x, y in pool:
        fn(x, y)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        for x, y in pool:
            fn(x, y)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * len(pool))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4096)
    ap.add_argument("--cols", type=int, default=3072)
    a = ap.parse_args()
    torch.manual_seed(0)
    for dtype in (torch.float16, torch.bfloat16):
        mk = lambda: (torch.randn(a.rows, a.cols, device="cuda", dtype=dtype) * 2,
                      torch.empty(a.rows, a.cols, device="cuda", dtype=dtype))
        nbytes = 2 * a.rows * a.cols * 2
        for name, pool, reps in (("l2", [mk()], 200), ("cold", [mk() for _ in range(12)], 20)):
            us_oeh = timed(lambda x, y: ops.gelu(x, out=y), pool, reps)
            us_aten = timed(lambda x, y: torch.nn.functional.gelu(x), pool, reps)
            print(json.dumps(dict(kernel="gelu_erf", dtype=str(dtype), rows=a.rows, cols=a.cols, mode=name,
                                  us=us_oeh, gbs=nbytes / us_oeh * 1e-3, aten_us=us_aten, aten_gbs=nbytes / us_aten * 1e-3)),
                  flush=True)


if __name__ == "__main__":
    main()
