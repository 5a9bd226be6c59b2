"""Smallest launches that exercise every synchronisation pattern of the fused kernel, for compute-sanitizer
(one tool per gpurun call):  compute-sanitizer --tool racecheck|memcheck python tools/sanitize_probe.py
  * L=J=5 (four CTAs per SM, 256 threads) and L=J=10 (one CTA per SM, 384 threads, two-stage schedule) region launches
    -- conv_all's barrier-free combine (per-box counters in shared memory + __threadfence_block), the series epilogue,
    the block reductions;
  * the same on 2- and 4-CTA clusters (RG_B200_CLUSTER): DSMEM row broadcast between cluster barriers;
  * expm_adapt and expm_multiply once each (estimator columns in the global workspace);
  * the stand-alone HBM operators.
Prints the maps so that a sanitised run can be compared with a plain one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rg_b200 import RG, make_param_dict
from rg_b200.rg_modules import region_maps

def run(L, J, ct, xs, ys, cluster=None):
    if cluster:
        os.environ["RG_B200_CLUSTER"] = str(cluster)
    else:
        os.environ.pop("RG_B200_CLUSTER", None)
    c = RG(make_param_dict(L=L, J=J, CanonicalTransformation=ct, SaveData=False), device=0)
    data, info = region_maps(c, xs, ys)
    print("L=%d J=%d %s cluster=%s count=%s err=%s liouvilles=%d" % (L, J, ct, cluster, data[:, :, 1].ravel().tolist(),
                                                                     info.ravel().tolist(), c.last_stats.liouville), flush=True)
    f = np.random.default_rng(0).standard_normal((3,) + c.r_jl) * 1e-3
    f[:, 2, 0, 0] = 0.5
    c.norm_batch(f); c.sym_batch(f); c.reindex_rescale_batch(f, c.Omega)
    c.close()

quick = np.array([0.02, 0.05])
run(5, 5, "expm_onestep", quick, np.array([0.03, 0.06]))
run(5, 5, "expm_onestep", quick, np.array([0.03]), cluster=2)
run(5, 5, "expm_onestep", quick, np.array([0.03]), cluster=4)
run(5, 5, "expm_adapt", quick[:1], np.array([0.03]))
run(5, 5, "expm_multiply", quick[:1], np.array([0.03]))
run(10, 10, "expm_onestep", np.array([0.09]), np.array([0.30]))           # diverges in one rg_map: a few brackets
run(10, 10, "expm_onestep", np.array([0.09]), np.array([0.30]), cluster=4)
print("done")
