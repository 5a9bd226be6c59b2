"""compute_surface (BASELINE config 3: Nxy=128, L=J=10, RelDist=1e-7) timing: all 128 columns on one GPU and the 16
columns one of 8 ranks owns (device-side bisection, rg_approach): python tools/probe_surface.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rg_b200 import RG, make_param_dict
from rg_b200.rg_modules import compute_cr_batch, shard_indices

eps = np.linspace(0, 1, 128)
ref = None
for ncols, label in ((16, "16 columns (one of 8 ranks)"), (128, "128 columns (one GPU)")):
    idx = shard_indices(128, 0, 128 // ncols)
    for depth in (1,):
        c = RG(make_param_dict(L=10, J=10, CanonicalTransformation="expm_onestep", RelDist=1e-7, Nxy=128, Method="surface", SaveData=False))
        c.norm(np.zeros(c.r_jl))
        calls = {"n": 0, "ms": 0.0, "liou": 0}
        for name in ("converge_batch", "approach_batch"):
            def wrapped(*a, _o=getattr(c, name)):
                r = _o(*a)
                calls["n"] += 1; calls["ms"] += c.last_stats.kernel_ms; calls["liou"] += c.last_stats.liouville
                return r
            setattr(c, name, wrapped)
        t0 = time.perf_counter()
        res = np.array(compute_cr_batch([eps[i] for i in idx], c), dtype=np.float64)
        dt = time.perf_counter() - t0
        if ncols == 16 and depth == 1:
            ref = res
        print("%s: %.2f s wall, %d launches, kernel %.2f s, %d liouvilles, bisection levels %s" %
              (label, dt, calls["n"], calls["ms"] * 1e-3, calls["liou"], np.bincount(c.last_levels).nonzero()[0].tolist()))
        c.close()
