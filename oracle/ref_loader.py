"""Load the UNMODIFIED reference (cchandre/RG) from /root/reference.

TEST INFRASTRUCTURE ONLY.  Usable only in the build container (the GPU box has
no /root/reference).  Used by tests/golden/make_golden.py to freeze golden
vectors and by oracle validation tests that are skipped when the reference is
absent.  Nothing in the product package (rg_b200/) may import this file.

The reference imports matplotlib unconditionally (RG.py:29, RG_modules.py:30-32,39)
and opens theta_taylor.mat by relative path (RG.py:98), so we stub matplotlib and
chdir into a scratch directory holding symlinks to the reference files.
"""
import os
import sys
import tempfile
import types

REFERENCE_DIR = os.environ.get("RG_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "RG.py"))


class _Any:
    def __getattr__(self, k):
        return _Any()

    def __call__(self, *a, **k):
        return _Any()

    def __iter__(self):
        return iter((_Any(), _Any()))


_loaded = None


def load():
    """Return (RG_module, RG_modules_module, RG_dict_module); chdir's into a scratch dir."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_DIR)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__getattr__ = lambda k: _Any()
            sys.modules[name] = m
    sys.modules["matplotlib.pyplot"].rcParams = {}
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    scratch = tempfile.mkdtemp(prefix="rg_ref_")
    for fn in ("RG.py", "RG_modules.py", "RG_dict.py", "theta_taylor.mat"):
        os.symlink(os.path.join(REFERENCE_DIR, fn), os.path.join(scratch, fn))
    os.chdir(scratch)
    sys.path.insert(0, scratch)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import RG_dict
        import RG
        import RG_modules
    RG_modules.tqdm = lambda it, **k: it
    _loaded = (RG, RG_modules, RG_dict)
    return _loaded


def make_params(**over):
    """param_dict of the reference's golden-mean block with overrides.

    Remember the prefixes: Method carries 'compute_' (RG_dict.py:76) and
    CanonicalTransformation carries 'self.' (RG_dict.py:96).
    """
    import numpy as xp
    RG, RG_modules, RG_dict = load()
    p = dict(RG_dict.param_dict)
    p["AmpInf"] = list(p["AmpInf"])
    p["AmpSup"] = list(p["AmpSup"])
    p["SaveData"] = False
    p["PlotResults"] = False
    ct = over.pop("CT", None)
    if ct is not None:
        p["CanonicalTransformation"] = "self." + ct
    torus = over.pop("torus", "golden")
    if torus == "spiral":
        sigma = 1.3247179572447460259
        p.update(N=xp.asarray([[0, 1, 0], [0, 0, 1], [1, 0, -1]], dtype=int),
                 omega0=xp.asarray([sigma, sigma ** 2, 1.0], dtype=xp.float64),
                 Omega=xp.asarray([1, 1, -1], dtype=xp.float64),
                 K=((0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1)),
                 AmpInf=[0, 0, 0], AmpSup=[0.05, 0.25, 0.1],
                 CoordLine=[0, 0.05], ModesLine=xp.asarray((1, 1, 0)),
                 DirLine=xp.asarray([1, 5, 0.1]))
        p["ModesK"] = [p["K"][_] for _ in xp.nonzero(p["ModesLine"])[0]]
    elif torus == "taumean":   # RG_dict.py:36-46
        tau = 0.445041867912629
        p.update(N=xp.asarray([[0, 1, -1], [1, -1, 1], [0, -1, 2]], dtype=int),
                 omega0=xp.asarray([1.0, tau, 1.0 - tau - tau ** 2], dtype=xp.float64),
                 Omega=xp.asarray([1, 1, -1], dtype=xp.float64),
                 K=((0, 0, -1, 1), (0, 1, -1, -1), (0, 0, 0, 1)),
                 AmpInf=[0.0, 0.0, 0.01], AmpSup=[7e-4, 6e-3, 0.01],
                 CoordLine=[0.0, 0.002], ModesLine=xp.asarray((1, 1, 0)),
                 DirLine=xp.asarray([1, 5, 0.01]))
        p["ModesK"] = [p["K"][_] for _ in xp.nonzero(p["ModesLine"])[0]]
    p.update(over)
    return p
