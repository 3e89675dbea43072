"""Per-CTA timeline of the per-item warp-specialised attention kernel at the config-3 shape (B=4, H=12, T=512 causal, vanilla softmax
+ per-token gate; one 128-row Q tile per CTA, two CTAs per SM) from %globaltimer stamps.
Needs the experiment build:  tools/build_variant.sh timeline -DOEH_TIMELINE
Run:  OEH_LIB_PATH=outeffhop_b200/lib/liboeh_timeline.so python tools/timeline_ws.py
Stamps (stream (tile 0, parity 0) of each CTA): 0 entry, 1 TMEM allocated, 2 first S ready, 3 last P written, 4 statistics merged,
5 last PV complete, 6 stored."""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from outeffhop_b200 import ops, _lib

B, H, T = 4, 12, 512
dev = torch.device("cuda")
torch.manual_seed(0)
mk = lambda: torch.randn(B, T, H, 64, device=dev, dtype=torch.float16).permute(0, 2, 1, 3)
sets = [(mk(), mk(), mk()) for _ in range(8)]
gl = torch.randn(B, H, T, device=dev)
out = torch.empty(B, T, H * 64, device=dev, dtype=torch.float16)
kw = dict(scale=0.125, causal=True, gate_logits=gl, out=out, kind=0)
n_cta = B * H * (T // 128)
buf = torch.zeros(n_cta * 8, dtype=torch.int64, device=dev)
lib = _lib.lib()
lib.oeh_set_ws_timeline.argtypes = [ctypes.c_void_p]
for i in range(5):
    ops.attention_fwd(*sets[i % 8], **kw)
torch.cuda.synchronize()
assert lib.oeh_set_ws_timeline(buf.data_ptr()) == 0
names = ["entry", "alloc'd", "S ready", "last P", "merged", "O final", "stored"]
for i in range(8):
    buf.zero_()
    ops.attention_fwd(*sets[i], **kw)
    torch.cuda.synchronize()
    t = buf.cpu().numpy().reshape(n_cta, 8).astype(np.int64)
    if i < 6:
        continue
    sm, nkv = t[:, 7] & 0xffff, t[:, 7] >> 16
    st = t[:, :7] - t[:, 0].min()
    print(f"launch {i}: span {st[:, 6].max()} ns, {n_cta} CTAs, CTAs per SM histogram {np.bincount(np.bincount(sm)).tolist()}")
    for n in sorted(set(nkv.tolist()), reverse=True):
        sel = nkv == n
        print(f"  n_kv={n}: {sel.sum()} CTAs; median " + " ".join(f"{names[k]}={int(np.median(st[sel, k]))}" for k in range(7))
              + f" | latest stored {st[sel, 6].max()} | median own time {int(np.median(st[sel, 6] - st[sel, 0]))}")
    for j in np.argsort(-st[:, 6])[:5]:
        print(f"  late cta {j}: n_kv={nkv[j]} sm={sm[j]} shares its SM with {int((sm == sm[j]).sum()) - 1} " + ",".join(str(int(x)) for x in st[j]))
