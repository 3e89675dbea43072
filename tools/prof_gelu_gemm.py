"""One profiled launch set of oeh_linear_fwd with the erf-GELU epilogue at the BERT FFN1 shape (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outeffhop_b200 import ops
torch.manual_seed(0)
M, N, K = 4096, 3072, 768
x = torch.randn(M, K, device="cuda", dtype=torch.float16)
w = torch.randn(N, K, device="cuda", dtype=torch.float16) * K ** -0.5
b = torch.randn(N, device="cuda", dtype=torch.float16)
y = torch.empty(M, N, device="cuda", dtype=torch.float16)
for _ in range(3):
    ops.linear_fwd(x, w, b, out=y, act="gelu")
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ops.linear_fwd(x, w, b, out=y, act="gelu")
ops.linear_fwd(x, w, b, out=y)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
