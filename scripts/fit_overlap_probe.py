"""How much of the weight-space fit hides behind the GP fit (scripts/fit_overlap_probe.py): iteration time for the
sequential order and for the concurrent order with / without a high-priority GP stream."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppbo_b200 import iteration, synthetic  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    prob = synthetic.make_problem("ackley20d")
    inp = iteration.IterationInputs(prob["X"], None, prob["W"], prob["b"], None, prob["grids"])
    d = inp.to_device(dev)
    for conc, prio in ((False, None), (True, None), (True, -1), (True, -2)):
        iteration.CONCURRENT_FITS = conc
        iteration.GP_STREAM_PRIORITY = prio
        ts = []
        for i in range(7):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            timers = []
            try:
                sums, gp, rff = iteration.run_iteration(d, prob["kernel"], prob["theta"], prob["Q"], prob["m"], prob["S"], seed=1, timers=timers)
            except Exception as e:
                print("concurrent=%s priority=%s failed: %s" % (conc, prio, e))
                break
            b.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append([a.elapsed_time(b)] + [x[1].elapsed_time(y[1]) for x, y in zip(timers[:-1], timers[1:])])
        if ts:
            print("concurrent=%s gp priority=%s: total %.2f ms; stages %s %s" % (
                conc, prio, np.mean([t[0] for t in ts]), [n for n, _ in timers[1:]], np.round(np.mean(ts, axis=0)[1:], 2)), flush=True)


if __name__ == "__main__":
    main()
