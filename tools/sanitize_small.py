#!/usr/bin/env python
"""Small invocations of every round-2 kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import _lib, ops
torch.manual_seed(0)
qp = torch.tensor([0.05, 128.0], device="cuda"); pp = torch.tensor([1.0 / 255, 0.0], device="cuda"); cp = torch.tensor([0.01, 128.0], device="cuda")
# mid kernel: causal + taps, ragged, 1..4 tiles
# (the (4, 12, 512) launch has 192 items for at most 148 CTAs: several items per CTA, buffers and barrier phases carried over)
for (B, H, Tq, Tk, causal) in ((1, 2, 512, 512, True), (1, 1, 200, 333, False), (1, 2, 130, 130, True), (4, 12, 512, 512, True), (7, 12, 300, 300, False)):
    q = torch.randn(B, H, Tq, 64, device="cuda").half(); k = torch.randn(B, H, Tk, 64, device="cuda").half(); v = torch.randn_like(k)
    _lib.set_override("OEH_MID_KERNEL", "1")
    ops.attention_fwd(q, k, v, scale=0.125, kind=3, gamma=-0.025, eta=1.1, causal=causal, qdq_scores=qp, qdq_probs=pp, qdq_ctx=cp, n_bits=8)
    ops.attention_fwd(q, k, v, scale=0.125, kind=1, causal=causal)
    _lib.set_override("OEH_MID_KERNEL", None)
# fp32 path
q = torch.randn(1, 2, 100, 64, device="cuda"); ops.attention_fwd(q, q, q, scale=0.125, kind=1)
# GEMM: every tile shape, activations, taps, ragged
for (M, N, K) in ((300, 520, 264), (128, 256, 64), (513, 768, 136)):
    x = torch.randn(M, K, device="cuda").half(); w = torch.randn(N, K, device="cuda").half() * 0.05; b = torch.randn(N, device="cuda").half()
    for tile in (128, 192, 256, 2128, 2192, 2256):
        ops.linear_fwd(x, w, b, act="gelu", qparams=qp, tile_n=tile)
        ops.linear_fwd(x, w, None, tile_n=tile)
x = torch.randn(256, 768, device="cuda").half(); w = torch.randn(768 * 3, 768, device="cuda").half() * 0.03
ops.linear_fwd(x, w, None, qparams=[qp, pp, cp])
# LayerNorm + gate, embeddings
a = torch.randn(2, 70, 768, device="cuda").half(); g = torch.ones(768, device="cuda").half(); be = torch.zeros(768, device="cuda").half()
ops.add_layer_norm(a, a, g, be, 1e-5, return_sum=True, gate=(torch.randn(12, 64, device="cuda"), torch.randn(12, device="cuda")))
ops.add_layer_norm(a, a, g, be, 1e-5, qparams_a=qp, qparams_res=cp, qparams_out=qp, return_sum=True)   # taps: the common-chain body
ops.add_layer_norm(a, a, g, be, 1e-5, qparams_res=torch.tensor([1e-38, 3.0], device="cuda"), n_bits=12)     # taps: the general body
ops.fake_quant_asym(a, qp, 8)
ids = torch.randint(0, 1000, (2, 70), device="cuda")
ops.embed_layer_norm(ids, torch.randn(1000, 768, device="cuda").half(), torch.randn(128, 768, device="cuda").half(), torch.randn(2, 768, device="cuda").half(), g, be, 1e-12)
# one-tile kernels (one / two threads per row, coalesced stores), (B, T, H*d) layout with ragged rows
q = torch.randn(2, 100, 3, 64, device="cuda").half().permute(0, 2, 1, 3); k = torch.randn(2, 100, 3, 64, device="cuda").half().permute(0, 2, 1, 3)
for split in ("0", "1"):
    _lib.set_override("OEH_ONE_SPLIT", split)
    ops.attention_fwd(q, k, k, scale=0.125, kind=3, gamma=-0.025, eta=1.1)
    ops.attention_fwd(q, k, k, scale=0.125, kind=1, qdq_scores=qp, qdq_probs=pp, qdq_ctx=cp, n_bits=8)
_lib.set_override("OEH_ONE_SPLIT", None)
# warp-specialised kernels: persistent (several items per CTA, cross-item QK^T issue, staged stores) and per-item, one / two tiles
q = torch.randn(3, 40, 600, 64, device="cuda").half()
for pers in ("0", "1"):
    for tiles in ("1", "2"):
        _lib.set_override("OEH_WS_PERSISTENT", pers); _lib.set_override("OEH_WS_TILES", tiles)
        ops.attention_fwd(q, q, q, scale=0.125, kind=1, causal=True)
        ops.attention_fwd(q, q, q, scale=0.125, kind=3, gamma=-0.025, eta=1.1)
_lib.set_override("OEH_WS_PERSISTENT", None); _lib.set_override("OEH_WS_TILES", None)
# Linear -> + residual -> LayerNorm cluster kernel
x = torch.randn(300, 264, device="cuda").half(); w = torch.randn(768, 264, device="cuda").half() * 0.05
r = torch.randn(300, 768, device="cuda").half(); gb = torch.randn(768, device="cuda").half()
ops.linear_add_layer_norm(x, w, gb, r, g, be, 1e-5, qparams_a=qp, qparams_res=cp, qparams_out=qp, return_sum=True)
torch.cuda.synchronize()
print("sanitize_small ok")
