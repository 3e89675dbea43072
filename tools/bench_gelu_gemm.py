"""us per launch of oeh_linear_fwd with the erf-GELU epilogue at the BERT FFN1 shape (in-graph, rotating inputs) for the library
named by OEH_LIB_PATH, plus the max difference against GELU(library GEMM) (reciprocal / polynomial experiments)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import ops
from tools.bench_gemm import time_graph
torch.manual_seed(0)
M, N, K = 4096, 3072, 768
pool = [(torch.randn(M, K, device="cuda", dtype=torch.float16), torch.empty(M, N, device="cuda", dtype=torch.float16)) for _ in range(12)]
w = torch.randn(N, K, device="cuda", dtype=torch.float16) * K ** -0.5
b = torch.randn(N, device="cuda", dtype=torch.float16)
qp = torch.tensor([0.03, 128.0], device="cuda")
row = {"lib": os.path.basename(os.environ.get("OEH_LIB_PATH", "liboeh.so"))}
for name, kw in (("none", {}), ("gelu", dict(act="gelu")), ("gelu_tap", dict(act="gelu", qparams=qp))):
    row[name + "_us"] = round(time_graph(lambda s: ops.linear_fwd(s[0], w, b, out=s[1], **kw), pool), 2)
x = pool[0][0]
got = ops.linear_fwd(x, w, b, act="gelu")
ref = torch.nn.functional.gelu(torch.nn.functional.linear(x, w, b).float()).half()
d = (got.float() - ref.float()).abs()
row["mismatch_frac"] = round((got != ref).float().mean().item(), 6)
row["max_abs_diff"] = d.max().item()
print(json.dumps(row))
