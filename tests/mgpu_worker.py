"""Worker of tests/test_multi_gpu.py (launched by torch.distributed.run, one rank per GPU, NCCL): the N-rank iteration -- cold and
with an appended comparison set, with the default partition and with a share for rank 0 -- against the same iteration on one rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppbo_b200 import iteration, synthetic  # noqa: E402


class _Single(iteration.Shard):
    def __init__(self):
        self.dist, self.group, self.rank, self.world = None, None, 0, 1


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))
    shard = iteration.Shard()
    prob = synthetic.make_problem("levy10d", Q=13, S=4096, P=128, F=256)          # large enough for the INT8 sampling engine
    m, theta, kernel, S = prob["m"], prob["theta"], prob["kernel"], prob["S"]
    Q0 = 12
    n0 = Q0 * (m + 1)
    inp = iteration.IterationInputs(prob["X"][:n0], None, prob["W"], prob["b"], None, prob["grids"])
    d = inp.to_device(dev)
    block = torch.from_numpy(prob["X"][n0:]).to(dev)
    results = {}
    for tag, shares in (("default", None), ("rank0_samples", [0.5] + [1.0] * (shard.world - 1))):
        st = iteration.IterationState(kernel, theta, prob["D"], m, Q0 + 1, dev, d["W"], d["b"], shard=shard)
        sums_cold, _, _ = iteration.run_iteration(d, kernel, theta, Q0, m, S, shard=shard, seed=11, state=st, shares=shares)
        dw = {"block": block, "W": d["W"], "b": d["b"], "grids": d["grids"]}
        sums_warm, _, _ = iteration.run_iteration(dw, kernel, theta, Q0 + 1, m, S, shard=shard, seed=11, state=st, shares=shares)
        results[tag] = (sums_cold.cpu().numpy(), sums_warm.cpu().numpy())
    torch.cuda.synchronize()
    dist.barrier()
    if shard.rank == 0:
        one = _Single()
        st = iteration.IterationState(kernel, theta, prob["D"], m, Q0 + 1, dev, d["W"], d["b"], shard=one)
        ref_cold, _, _ = iteration.run_iteration(d, kernel, theta, Q0, m, S, shard=one, seed=11, state=st)
        dw = {"block": block, "W": d["W"], "b": d["b"], "grids": d["grids"]}
        ref_warm, _, _ = iteration.run_iteration(dw, kernel, theta, Q0 + 1, m, S, shard=one, seed=11, state=st)
        ref_cold, ref_warm = ref_cold.cpu().numpy(), ref_warm.cpu().numpy()
        for tag, (c, w) in results.items():
            for name, got, ref in (("cold", c, ref_cold), ("warm", w, ref_warm)):
                err = np.abs(got - ref).max() / np.abs(ref).max()
                # mu* (world >= 3: sharded candidates, max is exact) and the fits are identical; the sums differ by the order of the
                # partial sums only
                assert err <= 1e-12, (tag, name, err)
                assert int(np.argmax(got[:, 0])) == int(np.argmax(ref[:, 0])), (tag, name)
        print("MGPU_OK world=%d" % shard.world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
