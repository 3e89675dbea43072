import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import ops
torch.manual_seed(0)
for (M, N, K, tile, use_bias) in ((4096, 2304, 768, 256, True), (4096, 2304, 768, 256, False), (4096, 2304, 768, 128, True), (256, 512, 64, 256, False),
                                  (256, 512, 128, 256, False), (128, 256, 768, 256, False), (4096, 768, 768, 128, True), (4096, 256, 64, 256, False)):
    x = torch.randn(M, K, device="cuda").half(); w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
    b = torch.randn(N, device="cuda").half() if use_bias else None
    for rep in range(2):
        y = ops.linear_fwd(x, w, b, tile_n=tile)
        ref = (x.double() @ w.double().t() + (b.double() if use_bias else 0)).float()
        err = (y.float() - ref).abs()
        bad = err > 4e-3 * ref.abs().clamp_min(0.25)
        rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
        print(f"M={M} N={N} K={K} tile={tile} bias={use_bias} rep={rep}: max err {err.max().item():.4f} bad {int(bad.sum())}"
              f" rows[{rows[:6].tolist()}..{rows[-3:].tolist() if len(rows) else []}] n_rows={len(rows)}"
              f" cols[{cols[:6].tolist()}..{cols[-3:].tolist() if len(cols) else []}] n_cols={len(cols)}")
        if int(bad.sum()):
            r, c = bad.nonzero()[0].tolist()
            print("   first bad", r, c, "got", y[r, c].item(), "ref", ref[r, c].item(), " row-tile", r // 128, "col-in-tile", c % tile)
