"""GPModel.mu_star on the full-size goldens (BASELINE configs 1-3): the scipy call ('de-scipy') against its C++ replay ('de', the
default) with one trial per launch and with speculative windows of 32: time, evaluations, time per evaluation, launches, and that all
return the same bits.  Run on the GPU box:
  python scripts/mustar_probe.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "src"), os.path.join(ROOT, "tests")]
from conftest import golden_full_names, load_golden  # noqa: E402
from test_src_gpu import _model  # noqa: E402

for name in golden_full_names():
    g = load_golden(name)
    st, gp = _model(g)
    gp.mu_pred(g["xstar"])                                         # first launch outside the timed region
    res = {}
    for tag, method, window in (("scipy", "de-scipy", 1), ("w1", "de", 1), ("w32", "de", 32), ("w32", "de", 32)):
        gp.mustar_method, gp.mustar_window = method, window
        gp.mu_pred_calls = gp.mu_star_launches = 0
        np.random.seed(3)
        t0 = time.perf_counter()
        xs, mu, loc = gp.mu_star()
        dt = time.perf_counter() - t0
        res[tag] = (xs, mu, dt, gp.mu_pred_calls, gp.mu_star_launches)
    a, b, c = res["scipy"], res["w1"], res["w32"]
    same = all(bool(np.array_equal(a[0], r[0]) and a[1] == r[1] and a[3] == r[3]) for r in (b, c))
    print("%-16s N=%5d D=%2d  scipy %7.1f ms | replay, 1 trial per launch %7.1f ms | windows of 32: %6.1f ms, %5d launches  "
          "(%6d evaluations: %4.1f -> %4.1f -> %4.1f us each)  identical: %s" % (
              name, gp.N, gp.D, 1e3 * a[2], 1e3 * b[2], 1e3 * c[2], c[4], c[3], 1e6 * a[2] / a[3], 1e6 * b[2] / b[3],
              1e6 * c[2] / c[3], same), flush=True)
