"""HBM roofline of the stand-alone batched operators: rg_norm, rg_sym, rg_reindex_rescale on
device-resident batches, CUDA-event timed.  python tools/bench_ops.py [B] > profiles/ops_hbm_rNN.json"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rg_b200 import RG, make_param_dict, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
peak = float(peaks.get("hbm_gbs", peaks.get("hbm_gbs_burst", 0)) or 0)
case = RG(make_param_dict(L=10, J=10, CanonicalTransformation="expm_onestep", SaveData=False))
vecjl = int(np.prod(case.r_jl))
L = _lib.lib()
plan = case._plan()
nrm, onu, oms, tab = case._tables(case.Omega)
g = torch.Generator(device="cuda").manual_seed(1)
f = torch.rand((B, vecjl), dtype=torch.float64, device="cuda", generator=g) - 0.5
f.view(B, case.J + 1, -1)[:, 2, 0] = 0.5
out_v = torch.empty_like(f)
out_s = torch.empty(B, dtype=torch.float64, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
res = {"B": B, "vecjl": vecjl, "peak_gbs": peak, "ops": {}}


def timed(fn, nbytes, reps=5):
    ts = []
    for _ in range(reps + 2):
        flush.fill_(1)                      # > L2 between repetitions
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # the ABI call runs on the plan's own stream and returns synchronised: bracket it with events on the
        # legacy default stream (which both events and the plan stream's sync order against) and keep the best
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts[2:])
    return {"ms": t, "GBps": nbytes / t * 1e-6, "frac_of_peak": (nbytes / t * 1e-6) / peak if peak else None, "bytes": nbytes}


for which, name in ((0, "norm"), (1, "norm_int"), (2, "norm_iminus")):
    res["ops"]["rg_" + name] = timed(lambda: _lib.check(L.rg_norm(plan, B, f.data_ptr(), which, out_s.data_ptr())), 8 * vecjl * B)
res["ops"]["rg_sym"] = timed(lambda: _lib.check(L.rg_sym(plan, B, f.data_ptr(), out_v.data_ptr())), 16 * vecjl * B)
res["ops"]["rg_reindex_rescale"] = timed(lambda: _lib.check(L.rg_reindex_rescale(plan, tab, B, f.data_ptr(), None, out_v.data_ptr())), 16 * vecjl * B)
# plain device copy of the same size as the yardstick
res["ops"]["torch_copy"] = timed(lambda: out_v.copy_(f), 16 * vecjl * B)
print(json.dumps(res))
