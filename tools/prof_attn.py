#!/usr/bin/env python
"""Launch the fused attention kernel a few times on one shape (the command ncu wraps).
    python tools/prof_attn.py --T 4096 --B 4 --kind 1 [--causal] [--iters 3] [--bert-layout]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=4096)
    ap.add_argument("--B", type=int, default=4)
    ap.add_argument("--H", type=int, default=12)
    ap.add_argument("--kind", type=int, default=1)
    ap.add_argument("--causal", action="store_true")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--dtype", type=str, default="float16")
    ap.add_argument("--bert-layout", action="store_true", help="(B,T,H,d) memory viewed as (B,H,T,d)")
    a = ap.parse_args()
    dt = getattr(torch, a.dtype)
    torch.manual_seed(0)
    if a.bert_layout:
        mk = lambda: torch.randn(a.B, a.T, a.H, 64, device="cuda", dtype=dt).permute(0, 2, 1, 3)
    else:
        mk = lambda: torch.randn(a.B, a.H, a.T, 64, device="cuda", dtype=dt)
    q, k, v = mk(), mk(), mk()
    for _ in range(a.iters):
        ops.attention_fwd(q, k, v, scale=0.125, kind=a.kind, gamma=-0.025, eta=1.1, causal=a.causal)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        ops.attention_fwd(q, k, v, scale=0.125, kind=a.kind, gamma=-0.025, eta=1.1, causal=a.causal)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    fl = 4.0 * a.B * a.H * a.T * a.T * 64 * (0.5 if a.causal else 1.0)
    print(f"T={a.T} B={a.B} kind={a.kind} causal={a.causal}: {ms:.4f} ms  {fl / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
