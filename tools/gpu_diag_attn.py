#!/usr/bin/env python
"""First-light diagnostics for the tcgen05 attention kernel (run on the GPU box).
Dumps raw S (tile 0) and raw O accumulators next to the final output and compares each stage with a
torch fp32 restatement, so a descriptor/layout bug can be localised from one run."""
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import ops  # noqa: E402


def ref_attn(q, k, v, scale, kind, gamma, eta, causal):
    qf, kf, vf = q.float(), k.float(), v.float()
    s = qf @ kf.transpose(-1, -2) * scale
    if causal:
        Tq, Tk = s.shape[-2:]
        tri = torch.ones(Tq, Tk, dtype=torch.bool, device=s.device).triu(1 + Tk - Tq)
        s = s.masked_fill(tri, float("-inf"))
    m = s.max(-1, keepdim=True).values
    e = torch.exp(s - m)
    den = e.sum(-1, keepdim=True)
    if kind in (1, 3):
        den = den + torch.exp(-m)
    p = e / den
    if kind >= 2:
        p = torch.clip(p * (eta - gamma) + gamma, 0, 1)
    return p @ vf, s


def run(B, H, Tq, Tk, dtype, kind, causal=False, layout="bhtd", gamma=-0.025, eta=1.1, seed=0):
    torch.manual_seed(seed)
    dev = "cuda"
    if layout == "bhtd":
        q = torch.randn(B, H, Tq, 64, device=dev, dtype=dtype)
        k = torch.randn(B, H, Tk, 64, device=dev, dtype=dtype)
        v = torch.randn(B, H, Tk, 64, device=dev, dtype=dtype)
    else:   # BERT: (B,T,H,d) memory viewed as (B,H,T,d)
        q = torch.randn(B, Tq, H, 64, device=dev, dtype=dtype).permute(0, 2, 1, 3)
        k = torch.randn(B, Tk, H, 64, device=dev, dtype=dtype).permute(0, 2, 1, 3)
        v = torch.randn(B, Tk, H, 64, device=dev, dtype=dtype).permute(0, 2, 1, 3)
    dbg_s = torch.full((B, H, Tq, 128), float("nan"), device=dev)
    dbg_o = torch.full((B, H, Tq, 64), float("nan"), device=dev)
    out = ops.attention_fwd(q, k, v, scale=0.125, kind=kind, gamma=gamma, eta=eta, causal=causal,
                            out_layout="bhtd", _debug=(dbg_s, dbg_o))
    torch.cuda.synchronize()
    ref, s_ref = ref_attn(q, k, v, 0.125, kind, gamma, eta, causal)
    n0 = min(Tk, 128)
    s_raw_ref = (q.float() @ k.float().transpose(-1, -2))[..., :n0]
    e_s = (dbg_s[..., :n0] - s_raw_ref).abs().max().item()
    e_o = (out.float() - ref).abs().max().item()
    tag = f"B{B} H{H} Tq{Tq} Tk{Tk} {str(dtype)[6:]} kind{kind} causal{int(causal)} {layout}"
    print(f"{tag:60s} S-err {e_s:.3e}  out-err {e_o:.3e}  out-nan {int(torch.isnan(out).sum())}", flush=True)
    return e_s, e_o, out, ref, dbg_s, s_raw_ref, dbg_o


def main():
    print(torch.cuda.get_device_name(0), torch.version.cuda)
    cases = [
        dict(B=1, H=1, Tq=128, Tk=128, dtype=torch.float16, kind=0),
        dict(B=1, H=1, Tq=128, Tk=128, dtype=torch.float16, kind=1),
        dict(B=1, H=1, Tq=128, Tk=128, dtype=torch.bfloat16, kind=1),
        dict(B=2, H=3, Tq=128, Tk=128, dtype=torch.float16, kind=1, layout="bthd"),
        dict(B=1, H=1, Tq=128, Tk=256, dtype=torch.float16, kind=0),
        dict(B=1, H=2, Tq=256, Tk=512, dtype=torch.float16, kind=1),
        dict(B=1, H=2, Tq=100, Tk=77, dtype=torch.float16, kind=1),
        dict(B=2, H=2, Tq=300, Tk=300, dtype=torch.float16, kind=1, causal=True),
        dict(B=1, H=1, Tq=128, Tk=128, dtype=torch.float16, kind=3),
        dict(B=1, H=2, Tq=256, Tk=384, dtype=torch.float16, kind=2),
        dict(B=2, H=12, Tq=512, Tk=512, dtype=torch.float16, kind=3, causal=True, layout="bthd"),
        dict(B=1, H=12, Tq=2048, Tk=2048, dtype=torch.bfloat16, kind=1),
    ]
    first_fail = None
    for c in cases:
        try:
            e_s, e_o, out, ref, dbg_s, s_ref, dbg_o = run(**c)
            if first_fail is None and not (e_o < 2e-2):
                first_fail = (c, out, ref, dbg_s, s_ref, dbg_o)
        except Exception:
            traceback.print_exc()
            print("case failed:", c, flush=True)
            break
    if first_fail is not None:
        c, out, ref, dbg_s, s_ref, dbg_o = first_fail
        print("---- first failing case detail:", c)
        torch.set_printoptions(precision=4, linewidth=200, sci_mode=False)
        print("S kernel [0,0,:4,:8]\n", dbg_s[0, 0, :4, :8])
        print("S ref    [0,0,:4,:8]\n", s_ref[0, 0, :4, :8])
        print("S kernel rows 32..34\n", dbg_s[0, 0, 32:35, :8])
        print("S ref    rows 32..34\n", s_ref[0, 0, 32:35, :8])
        print("out kernel [0,0,:4,:8]\n", out[0, 0, :4, :8].float())
        print("out ref    [0,0,:4,:8]\n", ref[0, 0, :4, :8])
        print("O raw kernel [0,0,:4,:8]\n", dbg_o[0, 0, :4, :8])
        # where are the S errors?
        err = (dbg_s[..., :s_ref.shape[-1]] - s_ref).abs()
        print("S err per column block of 8:", err[0, 0].reshape(err.shape[2], -1, 8).amax((0, 2)))
        print("S err per row block of 8:", err[0, 0].reshape(-1, 8, err.shape[3]).amax((1, 2)))


if __name__ == "__main__":
    main()
