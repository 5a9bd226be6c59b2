"""compute_line (BASELINE config 2) timing for several speculation depths: python tools/probe_line.py"""
import os, sys, time, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rg_b200 import RG, make_param_dict, compute_line

for depth in (1, 2, 3, 1, 2):
    os.environ["RG_B200_SPEC_DEPTH"] = str(depth)
    c = RG(make_param_dict(L=10, J=10, CanonicalTransformation="expm_onestep", RelDist=1e-8, Method="line", SaveData=False))
    c.norm(np.zeros(c.r_jl))       # build the plan outside the timed region
    calls = {"n": 0, "liou": 0, "ms": 0.0}
    orig = c.converge_batch
    def wrapped(hl, _o=orig):
        r = _o(hl)
        calls["n"] += 1; calls["liou"] += c.last_stats.liouville; calls["ms"] += c.last_stats.kernel_ms
        return r
    c.converge_batch = wrapped
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        hl = compute_line(c)
    dt = time.perf_counter() - t0
    print("depth %d: %.2f s wall, %d converge launches, kernel %.2f s, %d liouvilles, eps=%.12f counts %s" %
          (depth, dt, calls["n"], calls["ms"] * 1e-3, calls["liou"], 2 * hl[0].f[c.K[0]], [h.count for h in hl]))
    c.close()
