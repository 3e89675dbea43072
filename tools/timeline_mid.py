"""Per-item timeline of the single-pass mid-length attention kernel (B=4, H=12, T=512 causal + 3 taps + gate: config 4) from
%globaltimer stamps.  Needs the experiment build:  tools/build_variant.sh timeline -DOEH_TIMELINE
Run:  OEH_LIB_PATH=outeffhop_b200/lib/liboeh_timeline.so python tools/timeline_mid.py
Stamps: 0 entry, 1 TMEM allocated, 2 first S tile ready, 3 pass A done, 4 P written, 5 O ready, 6 stored."""
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
qp = torch.tensor([0.05, 128.0], device=dev); pp = torch.tensor([1.0 / 255, 0.0], device=dev); cp = torch.tensor([0.01, 128.0], device=dev)
kw = dict(scale=0.125, causal=True, gate_logits=gl, out=out, kind=1, qdq_scores=qp, qdq_probs=pp, qdq_ctx=cp, n_bits=8, ctx_qdq_before_gate=True)
n_cta = B * H * (T // 128)
buf = torch.zeros(n_cta * 8, dtype=torch.int64, device=dev)
lib = _lib.lib()
lib.oeh_set_mid_timeline.argtypes = [ctypes.c_void_p]
for i in range(5):
    ops.attention_fwd(*sets[i % 8], **kw)
torch.cuda.synchronize()
assert lib.oeh_set_mid_timeline(buf.data_ptr()) == 0
names = ["entry", "alloc'd", "S0 ready", "passA done", "P written", "O ready", "stored"]
for i in range(8):
    buf.zero_()
    ops.attention_fwd(*sets[i], **kw)
    torch.cuda.synchronize()
    t = buf.cpu().numpy().reshape(n_cta, 8).astype(np.int64)
    if i < 6:
        continue
    sm, nkv = t[:, 7] & 0xffff, t[:, 7] >> 16
    st = t[:, :7] - t[:, 0].min()
    print(f"launch {i}: span {st[:, 6].max()} ns, {n_cta} CTAs")
    for n in (4, 3, 2, 1):
        sel = nkv == n
        print(f"  n_kv={n}: {sel.sum()} CTAs; median stamps " + " ".join(f"{names[k]}={int(np.median(st[sel, k]))}" for k in range(7))
              + f" | latest stored {st[sel, 6].max()} | median own time {int(np.median(st[sel, 6] - st[sel, 0]))}")
    late = np.argsort(-st[:, 6])[:6]
    for j in late:
        print(f"  late cta {j}: n_kv={nkv[j]} sm={sm[j]} " + ",".join(str(int(x)) for x in st[j]))
    second = st[:, 0] > 1500
    print(f"  items starting after 1.5 us (second item of a CTA / second wave): {second.sum()}, their n_kv histogram {np.bincount(nkv[second], minlength=5).tolist()}, "
          f"start median {int(np.median(st[second, 0])) if second.any() else -1}, stored max {st[second, 6].max() if second.any() else -1}")
