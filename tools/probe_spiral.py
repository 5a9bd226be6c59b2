"""BASELINE config 5 (spiral mean n=3, L=J=6) on a strided sub-grid: python tools/probe_spiral.py [n] [ct]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rg_b200 import RG, make_param_dict
from rg_b200.rg_dict import SPIRAL_MEAN
from rg_b200.rg_modules import region_amps

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ct = sys.argv[2] if len(sys.argv) > 2 else "expm_multiply"
L = int(sys.argv[3]) if len(sys.argv) > 3 else 6
case = RG(make_param_dict(L=L, J=L, CanonicalTransformation=ct, SaveData=False, **SPIRAL_MEAN))
stride = 128 // n
xs = np.linspace(case.AmpInf[0], case.AmpSup[0], 128)[stride // 2::stride]
ys = np.linspace(case.AmpInf[1], case.AmpSup[1], 128)[stride // 2::stride]
amps = region_amps(case, xs, ys)
info = case.plan_info()
print(info)
work = np.zeros(len(amps), np.int64)
t0 = time.perf_counter()
conv, count, err = case.points(amps, work=work)
dt = time.perf_counter() - t0
st = case.last_stats
print("n=%d ct=%s wall %.2fs kernel %.1f ms  %.2f pts/s  liouvilles %d (mean %.0f max %d)  useful %.2f TF/s issued %.2f TF/s" %
      (len(amps), ct, dt, st.kernel_ms, len(amps) / (st.kernel_ms * 1e-3), st.liouville, work.mean(), work.max(),
       2.0 * st.conv_fma_useful / st.kernel_ms * 1e-9, 2.0 * st.conv_fma_executed / st.kernel_ms * 1e-9))
print("   clocks per liouville: conv %.0f (warp 0 busy %.0f) epilogue %.0f total %.0f" %
      (st.clk_conv / max(st.liouville, 1), st.clk_conv_busy / max(st.liouville, 1), st.clk_epilogue / max(st.liouville, 1), st.clk_total / max(st.liouville, 1)))
print("converged %d errors %s" % (conv.sum(), dict(zip(*np.unique(err, return_counts=True)))))
print(count.reshape(len(ys), len(xs)))
