"""GPU diagnostic: the factor grown by ppbo_factor_extend against a from-scratch Cholesky (torch, diagnostics only) of the same
I + s G s, region by region."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from ppbo_b200 import iteration, ops, synthetic  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "levy10d"
cfg = synthetic.CONFIGS[name]
Q0 = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["Q"]
prob = synthetic.make_problem(name, Q=Q0 + 1)
m, theta = prob["m"], prob["theta"]
dev = torch.device("cuda", 0)
X = ops.to_dev(prob["X"])
n0 = Q0 * (m + 1)
st = iteration.GPState(prob["kernel"], theta, prob["D"], m, Q0 + 1, dev, tol=1e-8)
st.cold(X[:n0])
lap = st.lap
M0, M1, cap = Q0 * m, (Q0 + 1) * m, lap.cap
L = lap._Lfac[:cap * cap].view(cap, cap)
sa = lap.sa_fac
G = lap.G
A0 = torch.eye(M0, dtype=torch.float64, device=dev) + sa[:M0, None] * G[:M0, :M0] * sa[None, :M0]
Lref0 = torch.linalg.cholesky(A0)
print("cold factor vs torch: max abs diff %.3e (|L| max %.3e)" % ((torch.tril(L[:M0, :M0]) - Lref0).abs().max().item(), Lref0.abs().max().item()))
# the append (bordered warm start); PPBO_TRACE=1 prints the iteration
blk = X[n0:]
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
st.append(blk)
e.record()
torch.cuda.synchronize()
print("append: %.2f ms  stats %s" % (a.elapsed_time(e), st.lap.stats))
A1 = torch.eye(M1, dtype=torch.float64, device=dev) + sa[:M1, None] * G[:M1, :M1] * sa[None, :M1]
Lref1 = torch.linalg.cholesky(A1)
Lx = torch.tril(L[:M1, :M1])
b0 = (M0 // 128) * 128
for nm, r0, r1, c0, c1 in (("old rows < b0", 0, b0, 0, b0), ("old rows [b0,M0) x [0,b0)", b0, M0, 0, b0), ("old rows [b0,M0) x [b0,M0)", b0, M0, b0, M0),
                           ("new rows x [0,b0)", M0, M1, 0, b0), ("new rows x [b0,M1)", M0, M1, b0, M1)):
    if r1 > r0 and c1 > c0:
        print("grown factor, %-32s max abs diff %.3e" % (nm, (Lx[r0:r1, c0:c1] - Lref1[r0:r1, c0:c1]).abs().max().item()))
g = iteration.gp_fit(X, prob["kernel"], theta, Q0 + 1, m, tol=1e-9)
f1, f2 = st.f_map.cpu().numpy(), g.f_map.cpu().numpy()
print("warm vs cold mode: rel %.2e" % (np.abs(f1 - f2).max() / np.abs(f2).max()))
