"""small workload for ncu: Cholesky n=5000, one Newton iteration of the GP fit, one of the RFF fit"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppbo_b200 import ops, synthetic
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
A0 = torch.randn(n, n, dtype=torch.float64, device=dev)
A = A0 @ A0.T + n * torch.eye(n, dtype=torch.float64, device=dev)
W = A.clone()
ops.potrf_lower(W)
W.copy_(A)
torch.cuda.synchronize()
info, ws = ops.potrf_lower(W)
b = torch.randn(n, dtype=torch.float64, device=dev)
ops.potrs_vec(W, ws, b)
torch.cuda.synchronize()
if "--fit" in sys.argv:
    prob = synthetic.make_problem("ackley20d")
    X = ops.to_dev(prob["X"]); th = prob["theta"]; Q, m = prob["Q"], prob["m"]
    Sigma = ops.gram_regularized("SE_kernel", X, th[1], th[2], 1e-6)
    ops.laplace_fit(Sigma, Q, m, th[0], max_iter=2)
    Wf = ops.to_dev(prob["W"]); bf = ops.to_dev(prob["b"])
    Phi = ops.rff_features(Wf, bf, X, th[2], feature_major=True)
    ops.rff_fit(Phi, Q, m, th[0], max_iter=2)
    torch.cuda.synchronize()
