"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (AaltoPML/PPBO).

The reference is pure Python and lives read-only at /root/reference (this container only;
it does not exist on the GPU box).  It cannot be imported as-is on numpy 2 / scipy 1.18
because three third-party packages are absent and two scipy call signatures changed.
This module installs the smallest possible *test-side* compatibility shim
and imports the reference's flat modules under a private prefix so they never collide
with this repo's own ``src/`` modules of the same names.

Nothing in the product path (``ppbo_b200/``, ``src/``) may import this file.  Only
``oracle/make_golden.py`` and ``tests/`` (CPU-side, this container) use it.

Shim items (reference files are untouched):
  1. fake ``arspy.ars``            (imported by src/TGN_distribution.py:1, only used for 'TGN' grids)
  2. fake ``GPyOpt.methods``       (imported by src/gp_model.py:12, src/acquisition.py:4)
  3. ``scipy.linalg.solve``        sym_pos=True -> assume_a='pos'   (src/misc.py:93,100)
  4. ``scipy.optimize.minimize``   ravel() a (N,1) x0               (src/gp_model.py:374,381)
"""
import importlib
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PPBO_REFERENCE_ROOT", "/root/reference")
_REF_MODULES = ("misc", "kernels", "TGN_distribution", "feedback_processing",
                "ppbo_settings", "gp_model", "acquisition", "random_fourier_sampler")
_PREFIX = "_ppbo_ref_"
_loaded = {}


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def _install_fakes():
    if "arspy" not in sys.modules:
        arspy = types.ModuleType("arspy")
        ars = types.ModuleType("arspy.ars")

        def adaptive_rejection_sampling(*a, **k):
            raise NotImplementedError("arspy is absent; use alpha_grid_distribution='equispaced'")
        ars.adaptive_rejection_sampling = adaptive_rejection_sampling
        arspy.ars = ars
        sys.modules["arspy"] = arspy
        sys.modules["arspy.ars"] = ars
    if "GPyOpt" not in sys.modules:
        gpyopt = types.ModuleType("GPyOpt")
        methods = types.ModuleType("GPyOpt.methods")

        class BayesianOptimization:  # only the ctor/run signature; never a real optimiser
            def __init__(self, *a, **k):
                raise NotImplementedError("GPyOpt is absent; outer BO strategies are not oracled")
        methods.BayesianOptimization = BayesianOptimization
        gpyopt.methods = methods
        sys.modules["GPyOpt"] = gpyopt
        sys.modules["GPyOpt.methods"] = methods


def _patch_scipy():
    import numpy as np
    import scipy.linalg
    import scipy.optimize
    if not getattr(scipy.linalg.solve, "_ppbo_shim", False):
        _solve = scipy.linalg.solve

        def solve(a, b, *args, sym_pos=None, **kw):
            if sym_pos:
                kw.setdefault("assume_a", "pos")
            return _solve(a, b, *args, **kw)
        solve._ppbo_shim = True
        scipy.linalg.solve = solve
    if not getattr(scipy.optimize.minimize, "_ppbo_shim", False):
        _minimize = scipy.optimize.minimize

        def minimize(fun, x0, *args, **kw):
            return _minimize(fun, np.asarray(x0, dtype=float).ravel(), *args, **kw)
        minimize._ppbo_shim = True
        scipy.optimize.minimize = minimize
    if not hasattr(np, "product"):
        np.product = np.prod


class _FlatImportHook:
    """While the reference modules are executed, their flat ``from misc import ...`` must
    resolve to the reference's own siblings, not to this repo's ``src/`` modules."""

    def __enter__(self):
        self.saved = {m: sys.modules.get(m) for m in _REF_MODULES}
        for m in _REF_MODULES:
            if m in _loaded:
                sys.modules[m] = _loaded[m]
            else:
                sys.modules.pop(m, None)
        return self

    def __exit__(self, *exc):
        for m, old in self.saved.items():
            if old is None:
                sys.modules.pop(m, None)
            else:
                sys.modules[m] = old


def load():
    """Return a namespace with the reference modules: ref.gp_model.GPModel, ref.kernels.SE_kernel ..."""
    if not available():
        raise RuntimeError("reference not present at %s (expected on the GPU box)" % REFERENCE_ROOT)
    if len(_loaded) == len(_REF_MODULES):
        return types.SimpleNamespace(**_loaded)
    _install_fakes()
    _patch_scipy()
    with _FlatImportHook():
        for name in _REF_MODULES:
            if name in _loaded:
                continue
            path = os.path.join(REFERENCE_ROOT, "src", name + ".py")
            spec = importlib.util.spec_from_file_location(_PREFIX + name, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[_PREFIX + name] = mod
            sys.modules[name] = mod          # siblings import it by its flat name
            spec.loader.exec_module(mod)
            _loaded[name] = mod
    return types.SimpleNamespace(**_loaded)
