"""Two-pass (clipped, gamma = 0: nothing skipped) attention TF/s at T = 2048 / 8192 for the launch forms of the warp-specialised
kernel: per-item CTAs (default for these kinds) vs persistent CTAs, one or two Q tiles per CTA."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from outeffhop_b200 import _lib, ops
for T, reps in ((2048, 20), (8192, 5)):
    B = 65536 // T
    q = torch.randn(B, 12, T, 64, device="cuda", dtype=torch.float16); k = torch.randn_like(q); v = torch.randn_like(q)
    row = {"T": T}
    for pers, tiles in ((None, None), ("0", "1"), ("0", "2"), ("1", "1"), ("1", "2")):
        _lib.set_override("OEH_WS_PERSISTENT", pers)
        _lib.set_override("OEH_WS_TILES", tiles)
        for causal in (False, True):
            t = bench._event_time(torch, lambda: ops.attention_fwd(q, k, v, scale=0.125, kind=3, gamma=0.0, eta=1.1, causal=causal), reps=reps, warm=reps)
            fl = 2.0 * B * 12 * T * (T + 1) * 64 if causal else 4.0 * B * 12 * T * T * 64
            row[f"p{pers}_t{tiles}{'_causal' if causal else ''}"] = round(fl / t / 1e12, 1)
    _lib.set_override("OEH_WS_PERSISTENT", None); _lib.set_override("OEH_WS_TILES", None)
    print(json.dumps(row), flush=True)
    del q, k, v
