"""GPModel.mu_star on the full-size goldens (BASELINE configs 1-3): the scipy call ('de-scipy') against its C++ replay ('de', the
default): time, evaluations, time per evaluation, and that both return the same bits.  Run on the GPU box:
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
    for method in ("de-scipy", "de", "de-scipy", "de"):
        gp.mustar_method = method
        gp.mu_pred_calls = 0
        np.random.seed(3)
        t0 = time.perf_counter()
        xs, mu, loc = gp.mu_star()
        dt = time.perf_counter() - t0
        res[method] = (xs, mu, dt, gp.mu_pred_calls)
    a, b = res["de-scipy"], res["de"]
    print("%-16s N=%5d D=%2d  scipy %8.1f ms  replay %8.1f ms  (%6d evaluations: %5.1f -> %5.1f us each)  identical: %s" % (
        name, gp.N, gp.D, 1e3 * a[2], 1e3 * b[2], b[3], 1e6 * a[2] / a[3], 1e6 * b[2] / b[3],
        bool(np.array_equal(a[0], b[0]) and a[1] == b[1] and a[3] == b[3])))
