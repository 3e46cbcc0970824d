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


def _all_golden():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def golden_names():
    """small cases recorded from the executed reference (every N x N matrix stored)"""
    return [n for n in _all_golden() if not n.startswith("full_") and not n.endswith("_full")]


def golden_full_names():
    """BASELINE configs 1-3 at the reference drivers' own sizes (N = 1014 / 1066 / 520), recorded from the executed reference;
    N x N matrices as sampled rows"""
    return [n for n in _all_golden() if n.endswith("_full")]


def oracle_full_names():
    """BASELINE configs 4-5 (synthetic bench problems, N = 2080 / 5200): oracle fixtures of oracle/make_full_fixtures.py"""
    return [n[len("full_"):] for n in _all_golden() if n.startswith("full_")]


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


@pytest.fixture(params=golden_full_names())
def golden_full(request):
    return load_golden(request.param)


def load_oracle_full(name):
    z = np.load(os.path.join(GOLDEN_DIR, "full_" + name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def relerr(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
