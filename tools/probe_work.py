"""Per-point work distribution of a strided sub-grid of config 4: python tools/probe_work.py [n] [ct] [L=J]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rg_b200 import RG, make_param_dict
from rg_b200.rg_modules import region_amps

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ct = sys.argv[2] if len(sys.argv) > 2 else "expm_multiply"
LJ = int(sys.argv[3]) if len(sys.argv) > 3 else 10
stride = 256 // n
case = RG(make_param_dict(L=LJ, J=LJ, CanonicalTransformation=ct, SaveData=False))
xs = np.linspace(0, 0.12, 256)[stride // 2::stride]
ys = np.linspace(0, 0.35, 256)[stride // 2::stride]
amps = region_amps(case, xs, ys)
work = np.zeros(len(amps), np.int64)
t0 = time.perf_counter()
conv, count, err = case.points(amps, work=work)
dt = time.perf_counter() - t0
st = case.last_stats
print("n=%d ct=%s wall %.2fs kernel %.1f ms liouvilles %d mean %.1f max %d" % (len(amps), ct, dt, st.kernel_ms, work.sum(), work.mean(), work.max()))
print("errors", dict(zip(*np.unique(err, return_counts=True))))
order = np.argsort(-work)[:12]
for i in order:
    print("  pt %5d x=%.5f y=%.5f work %8d conv %d count %3d err %d" % (i, amps[i, 0], amps[i, 1], work[i], conv[i], count[i], err[i]))
np.savez_compressed(os.path.join("gpurun_out", "work_%s_%d%s.npz" % (ct, n, "" if LJ == 10 else "_L%d" % LJ)), work=work, conv=conv, count=count, err=err, amps=amps)
print("work percentiles", np.percentile(work, [50, 90, 99, 99.9]).astype(int))
