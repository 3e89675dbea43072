#!/usr/bin/env python
"""Launch duration of the fused attention kernel at the C-BERT shape (B=32, H=12, T=128, d=64; (B,T,H,d) strided
views in, (B,T,H*d) out) -- the roofline kernel of bench.py.  Two timings per (kernel, kind):
  cold  : a pool of input sets larger than L2 (8 x 25 MB), launches back to back inside one CUDA-event pair
  flush : one launch per event pair with an L2 flush in between (includes the launch gap)
OEH_ATTN_KERNEL=tile selects the multi-tile kernel for comparison.
    python tools/bench_bert_attn.py [--B 32] [--T 128] [--kernels one,tile]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--T", type=int, default=128)
    ap.add_argument("--H", type=int, default=12)
    ap.add_argument("--kernels", type=str, default="default,tile")
    ap.add_argument("--kinds", type=str, default="1,3")
    ap.add_argument("--pool", type=int, default=8)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    B, T, H = a.B, a.T, a.H
    torch.manual_seed(0)
    mk = lambda: torch.randn(B, T, H, 64, device="cuda", dtype=torch.float16).permute(0, 2, 1, 3)
    pool = [(mk(), mk(), mk(), torch.empty(B, T, H * 64, device="cuda", dtype=torch.float16)) for _ in range(a.pool)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    nbytes = 4 * B * H * T * 64 * 2
    for kern in a.kernels.split(","):
        from outeffhop_b200 import _lib
        _lib.set_override("OEH_ATTN_KERNEL", None if kern == "default" else kern)
        for kind in [int(x) for x in a.kinds.split(",")]:
            run = lambda s: ops.attention_fwd(s[0], s[1], s[2], scale=0.125, kind=kind, gamma=-0.025, eta=1.1, out=s[3])
            for s in pool:
                run(s)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            graph = torch.cuda.CUDAGraph()      # back-to-back launches without the host-side call overhead
            with torch.cuda.graph(graph):
                for _ in range(a.reps):
                    for s in pool:
                        run(s)
            graph.replay()
            torch.cuda.synchronize()
            flush.zero_()
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            cold_us = e0.elapsed_time(e1) * 1e3 / (a.reps * len(pool))
            ts = []
            for i in range(20):
                flush.zero_()
                e0.record()
                run(pool[i % len(pool)])
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            e0.record()
            for _ in range(50):
                run(pool[0])
            e1.record()
            torch.cuda.synchronize()
            hot_us = e0.elapsed_time(e1) * 1e3 / 50
            print(json.dumps(dict(kernel=kern, kind=kind, B=B, T=T, cold_us=round(cold_us, 2), flush_us_median=round(ts[10], 2),
                                  hot_us=round(hot_us, 2), cold_gbs=round(nbytes / cold_us / 1e3, 1),
                                  tflops_cold=round(4.0 * B * H * T * T * 64 / cold_us / 1e6, 1))), flush=True)


if __name__ == "__main__":
    main()
