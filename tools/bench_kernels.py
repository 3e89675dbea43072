#!/usr/bin/env python
"""Kernel-level throughput sweep (run on the GPU box): attention TFLOP/s per (T, kind, causal) and
fake-quant / min-max GB/s, CUDA-event timed with an L2 flush between iterations."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import ops  # noqa: E402

FLUSH = None


def flush_l2():
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    FLUSH.zero_()


def time_fn(fn, iters=10, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=65536)
    ap.add_argument("--seqs", type=str, default="128,256,512,1024,2048,4096,8192")
    ap.add_argument("--kinds", type=str, default="0,1,3")
    ap.add_argument("--dtype", type=str, default="float16")
    ap.add_argument("--out", type=str, default="")
    ap.add_argument("--no-quant", action="store_true")
    ap.add_argument("--no-causal", action="store_true")
    ap.add_argument("--gamma", type=float, default=-0.025, help="clipped kinds: gamma (>= 0 disables the sparse skip)")
    ap.add_argument("--eta", type=float, default=1.1)
    ap.add_argument("--qscale", type=float, default=1.0, help="multiply q by this (peaked attention rows)")
    a = ap.parse_args()
    dt = getattr(torch, a.dtype)
    res = {"attn": [], "quant": []}
    for T in [int(x) for x in a.seqs.split(",")]:
        B = max(1, a.tokens // T)
        H = 12
        torch.manual_seed(1000 + T)
        q = (torch.randn(B, H, T, 64, device="cuda", dtype=torch.float32) * a.qscale).to(dt)
        k = torch.randn(B, H, T, 64, device="cuda", dtype=dt)
        v = torch.randn(B, H, T, 64, device="cuda", dtype=dt)
        for kind in [int(x) for x in a.kinds.split(",")]:
            for causal in ((False,) if a.no_causal else (False, True)):
                fn = lambda: ops.attention_fwd(q, k, v, scale=0.125, kind=kind, gamma=a.gamma, eta=a.eta, causal=causal)
                med, best = time_fn(fn)
                flops = 4.0 * B * H * T * T * 64 if not causal else 2.0 * B * H * T * (T + 1) * 64
                r = dict(T=T, B=B, kind=kind, causal=causal, gamma=a.gamma, qscale=a.qscale, ms=med, ms_best=best, tflops=flops / med / 1e9,
                         tflops_best=flops / best / 1e9)
                res["attn"].append(r)
                print(json.dumps(r), flush=True)
    if not a.no_quant:
        for dtn in ("float32", "float16"):
            for logn in (20, 24, 28):
                n = 1 << logn
                x = torch.randn(n, device="cuda", dtype=getattr(torch, dtn))
                y = torch.empty_like(x)
                qp = torch.tensor([0.03, 128.0], device="cuda")
                med, best = time_fn(lambda: ops.fake_quant_asym(x, qp, out=y), flush=logn < 27)
                gb = n * x.element_size() * 2 / 1e9
                r = dict(op="fake_quant", dtype=dtn, n=n, ms=med, gbs=gb / med * 1e3, gbs_best=gb / best * 1e3)
                res["quant"].append(r)
                print(json.dumps(r), flush=True)
                mm = torch.empty(2, device="cuda")
                med, best = time_fn(lambda: ops.minmax(x, out=mm), flush=logn < 27)
                gb = n * x.element_size() / 1e9
                r = dict(op="minmax", dtype=dtn, n=n, ms=med, gbs=gb / med * 1e3, gbs_best=gb / best * 1e3)
                res["quant"].append(r)
                print(json.dumps(r), flush=True)
    if not a.no_quant:
        # fused residual add + LayerNorm (C-BERT rows x 768 and a large case), fused mean cross-entropy (C-BERT logits)
        for rows, cols in ((4096, 768), (262144, 768)):
            xa = torch.randn(rows, cols, device="cuda", dtype=torch.float16)
            xb = torch.randn_like(xa)
            w, bb = torch.ones(cols, device="cuda", dtype=torch.float16), torch.zeros(cols, device="cuda", dtype=torch.float16)
            y = torch.empty_like(xa)
            med, best = time_fn(lambda: ops.add_layer_norm(xa, xb, w, bb, 1e-12, out=y), flush=rows < 100000)
            gb = rows * cols * 2 * 3 / 1e9
            r = dict(op="add_layer_norm", dtype="float16", rows=rows, cols=cols, ms=med, gbs=gb / med * 1e3, gbs_best=gb / best * 1e3)
            res["quant"].append(r)
            print(json.dumps(r), flush=True)
        for rows, vocab, pad, frac in ((4096, 30522, 6, 0.15), (4096, 30522, 6, 1.0), (2048, 50272, 0, 1.0)):
            logits = torch.randn(rows, vocab + pad, device="cuda", dtype=torch.float16)
            labels = torch.randint(0, vocab, (rows,), device="cuda")
            labels[torch.rand(rows, device="cuda") >= frac] = -100
            labels[0] = 1
            med, best = time_fn(lambda: ops.ce_loss(logits, labels, vocab=vocab))
            n_lab = int((labels != -100).sum())
            gb = n_lab * vocab * 2 / 1e9                   # algorithmic bytes: the labelled rows, read once
            r = dict(op="ce_loss", dtype="float16", rows=rows, vocab=vocab, labelled=n_lab, ms=med, gbs=gb / med * 1e3,
                     gbs_best=gb / best * 1e3)
            res["quant"].append(r)
            print(json.dumps(r), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
