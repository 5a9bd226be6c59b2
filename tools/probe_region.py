"""Quick throughput probe of the region path (not the bench): python tools/probe_region.py [Nxy] [ct] [L]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rg_b200 import RG, make_param_dict, _lib
if os.environ.get("PROBE_SO"):   # experiment builds: a variant of librg_b200.so kept next to the product one
    _lib.SO_PATH = os.path.abspath(os.environ["PROBE_SO"])
from rg_b200.rg_modules import region_amps

nxy = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ct = sys.argv[2] if len(sys.argv) > 2 else "expm_onestep"
L = int(sys.argv[3]) if len(sys.argv) > 3 else 10
stride = 256 // nxy
case = RG(make_param_dict(L=L, J=L, CanonicalTransformation=ct, SaveData=False))
xs = np.linspace(0, 0.12, 256)[stride // 2::stride]
ys = np.linspace(0, 0.35, 256)[stride // 2::stride]
amps = region_amps(case, xs, ys)
if os.environ.get("PROBE_SAME"):
    # uniform batch: every CTA gets the same point -> no load imbalance, pure kernel speed
    x, y = (float(v) for v in os.environ["PROBE_SAME"].split(","))
    amps = np.tile(np.array([[x, y]]), (int(os.environ.get("PROBE_N", "296")), 1))
print(case.plan_info())
for rep in range(int(os.environ.get('PROBE_REPS', '2'))):
    t0 = time.perf_counter()
    conv, count, err = case.points(amps)
    dt = time.perf_counter() - t0
    st = case.last_stats
    info = case.plan_info()
    fl_useful = 2.0 * st.conv_fma_useful
    fl_exec = 2.0 * st.conv_fma_executed
    print("n=%d wall %.3fs kernel %.1f ms  pts/s %.1f  liouvilles %d (%.1f/pt) lie %d maps %d | useful %.2f TF/s executed %.2f TF/s dense-equiv %.2f TF/s | us/liouville/SM %.1f"
          % (len(amps), dt, st.kernel_ms, len(amps) / (st.kernel_ms * 1e-3), st.liouville, st.liouville / len(amps), st.lie_steps, st.rg_maps,
             fl_useful / st.kernel_ms * 1e-9, fl_exec / st.kernel_ms * 1e-9, st.liouville * 28.924e6 / st.kernel_ms * 1e-9,
             st.kernel_ms * 1e3 * 148 * info['ctas_per_sm'] / max(st.liouville, 1)))
    print("   clocks per liouville: conv %.0f (warp 0 busy %.0f)  epilogue %.0f  total %.0f" % (st.clk_conv / max(st.liouville, 1), st.clk_conv_busy / max(st.liouville, 1), st.clk_epilogue / max(st.liouville, 1), st.clk_total / max(st.liouville, 1)))
print("converged %d, errors %s" % (conv.sum(), dict(zip(*np.unique(err, return_counts=True)))))
if os.environ.get("PROBE_BUSY"):
    import ctypes
    from rg_b200 import _lib
    buf = (ctypes.c_ulonglong * 32)()
    _lib.lib().rg_debug_busy(buf)
    n = max(buf[31], 1)
    print("per-warp busy clocks per bracket (CTA 0):", [round(buf[i] / n) for i in range(12)])
    ls = max(buf[19], 1)
    print("clocks per Lie step (CTA 0, thread 0): generate_y %d, build_Y+build_G %d, sym_and_norms %d; brackets per Lie step %.2f"
          % (buf[16] / ls, buf[17] / ls, buf[18] / ls, n / ls))
    print("   whole Lie step %d, of which expm_series %d" % (buf[20] / ls, buf[21] / ls))
