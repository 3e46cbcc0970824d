import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
SRC = os.path.join(ROOT, "src")          # the drop-in package: flat modules with the reference's names
if SRC not in sys.path:
    sys.path.insert(1, SRC)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    for k in ("D", "Q", "m", "mc_samples", "Lambda_MAP_nnz"):
        if k in g:
            g[k] = int(g[k])
    g["kernel"] = str(g["kernel"])
    g["theta"] = [float(t) for t in g["theta"]]
    g["bounds"] = tuple((float(a), float(b)) for a, b in g["bounds"])
    return g


@pytest.fixture(params=golden_names())
def golden(request):
    return load_golden(request.param)


def relerr(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
