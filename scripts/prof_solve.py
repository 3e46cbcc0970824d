"""ncu workload: block-inverse solve and GEMV at n = 5000 (per-kernel durations of one chord step's building blocks)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppbo_b200 import ops
dev = torch.device("cuda", 0)
n = 5000
A0 = torch.randn(n, n, dtype=torch.float64, device=dev)
A = A0 @ A0.T + n * torch.eye(n, dtype=torch.float64, device=dev)
W = A.clone()
info, ws = ops.potrf_lower(W)
Wb = ops.blockinv_build(W, ws)
b = torch.randn(n, dtype=torch.float64, device=dev)
for _ in range(3):
    x = ops.potrs_vec_blockinv(W, Wb, b)
    y = ops.gemv(A, b)
torch.cuda.synchronize()
print("resid %.2e" % float((A @ x - b).abs().max() / b.abs().max()))
