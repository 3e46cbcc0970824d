"""per-step wall/device times of run_iteration over many steps (resident and host-input modes) + allocator state"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ppbo_b200 import iteration, synthetic
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
prob = synthetic.make_problem("ackley20d")
inputs = iteration.IterationInputs(prob["X"], None, prob["W"], prob["b"], None, prob["grids"])
B = prob["grids"].shape[0]
sums_host = torch.empty((B, 3), dtype=torch.float64).pin_memory()
res = inputs.to_device(dev)
def step(resident):
    d = res if resident else inputs.to_device(dev)
    sums, gp, rff = iteration.run_iteration(d, prob["kernel"], prob["theta"], prob["Q"], prob["m"], prob["S"], seed=1234)
    if not resident:
        sums_host.copy_(sums, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    return sums
for mode in (True, False, True, False):
    ts = []
    for i in range(14):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        step(mode)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    st = torch.cuda.memory_stats()
    print("resident" if mode else "host-in ", " ".join("%.1f" % t for t in ts), "| reserved %.1f GB, cudaMalloc calls %d, frees %d" % (
        torch.cuda.memory_reserved() / 2**30, st["num_device_alloc"], st["num_device_free"]))
