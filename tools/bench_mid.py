#!/usr/bin/env python
"""A/B of the single-pass mid-length attention kernel (attn_fwd_mid.cuh) against the two-pass / persistent kernels at the
OPT eval shape (B=4, H=12, T=512 causal; BASELINE configs 3 / 4) and neighbours: us per launch, cold (inputs rotate over
a pool larger than L2), CUDA graph of back-to-back launches, CUDA events."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import _lib, ops  # noqa: E402


def time_launch(run, pool, reps=10):
    for s in pool:
        run(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for s in pool:
                run(s)
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / (reps * len(pool)))
    return min(ts) * 1e3


def main():
    torch.manual_seed(0)
    qp = torch.tensor([0.05, 128.0], device="cuda")
    pp = torch.tensor([1.0 / 255, 0.0], device="cuda")
    cp = torch.tensor([0.01, 128.0], device="cuda")
    for (B, H, T, causal) in ((4, 12, 512, True), (4, 12, 512, False), (8, 12, 256, True), (16, 12, 384, False), (32, 12, 512, True)):
        n_pool = max(2, min(64, (200 << 20) // (B * H * T * 64 * 2 * 4)))
        mk = lambda: torch.randn(B, T, H, 64, device="cuda", dtype=torch.float16).permute(0, 2, 1, 3)
        pool = [(mk(), mk(), mk(), torch.randn(B, H, T, device="cuda"), torch.empty(B, T, H * 64, device="cuda", dtype=torch.float16))
                for _ in range(n_pool)]
        for name, kw in (("softmax1+3taps+gate", dict(kind=1, qdq_scores=qp, qdq_probs=pp, qdq_ctx=cp, n_bits=8, ctx_qdq_before_gate=True)),
                         ("clippedsoftmax1", dict(kind=3, gamma=-0.025, eta=1.1)),
                         ("vanilla+gate", dict(kind=0)), ("softmax1", dict(kind=1))):
            row = dict(B=B, H=H, T=T, causal=causal, case=name)
            run = lambda s: ops.attention_fwd(s[0], s[1], s[2], scale=0.125, causal=causal, gate_logits=s[3], out=s[4], **kw)
            for mid in ("0", "1", "3"):
                _lib.set_override("OEH_MID_KERNEL", mid)
                row[{"0": "other_us", "1": "mid_us", "3": "mid_2perrow_us"}[mid]] = round(time_launch(run, pool), 2)
            _lib.set_override("OEH_MID_KERNEL", None)
            _lib.set_override("OEH_ATTN_KERNEL", "tile")          # one 128-row Q tile per 128-thread CTA, two CTAs per SM
            row["tile_us"] = round(time_launch(run, pool), 2)
            _lib.set_override("OEH_ATTN_KERNEL", None)
            for pers, tiles in (("0", "1"), ("1", "1"), ("1", "2")):
                _lib.set_override("OEH_MID_KERNEL", "0")
                _lib.set_override("OEH_WS_PERSISTENT", pers)
                _lib.set_override("OEH_WS_TILES", tiles)
                row[f"ws_p{pers}_t{tiles}_us"] = round(time_launch(run, pool), 2)
            for k2 in ("OEH_MID_KERNEL", "OEH_WS_PERSISTENT", "OEH_WS_TILES"):
                _lib.set_override(k2, None)
            flops = (2.0 * B * H * T * (T + 1) * 64) if causal else 4.0 * B * H * T * T * 64
            row["mid_tflops"] = round(flops / row["mid_us"] / 1e6, 1)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
