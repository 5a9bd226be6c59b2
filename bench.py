#!/usr/bin/env python
"""Headline benchmark: region grid points/s (Nxy=256, L=J=10, golden-mean torus) -- BASELINE.json config 4.

    python bench.py --gpus N --steps K --warmup W            # this repo (B200 kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

One "step" = one pass of the hot path over the whole 256x256 grid (65 536 independent Hamiltonians,
`point` of RG_modules.py:121-125), sharded cyclically over the N ranks (strong scaling: the grid is
fixed), one CTA per Hamiltonian, one final all_gather of 3 x int32 per point.

  value : whole-job points/s from the device time of the kernel (CUDA events on the launching stream,
          inputs already resident in HBM when the events start), max over ranks.
  e2e   : the same metric through the public API (rg_b200.RG.points + gather) with HOST pinned buffers:
          H2D of the amplitudes, kernel, D2H of (conv, count, error), all_gather -- wall clock, max over ranks.
  roofline : fp64 FMA pipe. achieved = 2 x useful FMAs of the convolution loops (in-kernel liouville
          counter x the schedule's in-range, structurally non-zero FMA count) / kernel time, per GPU;
          peak = DFMA throughput measured on this pool by tools/fp64_peak.cu (profiles/fp64_peak_r01.json;
          fp64 is not in MEASURED_PEAKS.json).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = dict(L=10, J=10, Nxy=256, AmpInf=[0, 0], AmpSup=[0.12, 0.35])
DENSE_FLOP_PER_LIOUVILLE = 28.924e6   # 2 conv x 2 x T(10)=66 x F(10)^2=331^2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ct", default=os.environ.get("RG_BENCH_CT", "expm_onestep"),
                    choices=["expm_onestep", "expm_adapt", "expm_multiply"])
    ap.add_argument("--nxy", type=int, default=GRID["Nxy"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-ct", action="store_true", help="skip the expm_multiply / expm_adapt side measurements")
    ap.add_argument("--cpu-sample", type=int, default=8, help="side of the strided CPU-baseline subsample")
    ap.add_argument("--neighbour-queue", action="store_true", help="experimental: heavy points queue their grid neighbours")
    ap.add_argument("--emulate-world", type=int, default=0,
                    help="diagnostic: run rank 0's shard of an N-rank job on one GPU (tail / load-balance study)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons every 200 ms during the timed region (NVML)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while not self.stop_flag and self.nv is not None:
            try:
                self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ---------------------------------------------------------------------------------------------- reference arm
def cpu_sample_points(nxy, side):
    """Strided subsample of the Nxy grid: ix,iy in {s + s/2*.., ...}; side=8 at Nxy=256 -> {24,56,...,248}."""
    stride = nxy // side
    off = stride - stride // 4
    xs = np.linspace(GRID["AmpInf"][0], GRID["AmpSup"][0], nxy)[off::stride]
    ys = np.linspace(GRID["AmpInf"][1], GRID["AmpSup"][1], nxy)[off::stride]
    return xs, ys, "%dx%d strided subsample (ix,iy in {%d,%d,...}) of the Nxy=%d grid" % (len(xs), len(ys), off, off + stride, nxy)


def run_cpu(ct, nxy, side, steps, warmup):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import rg_oracle as ro
    cores = ro._lib().oracle_max_threads()
    orc = ro.OracleRG(ro.default_params(L=GRID["L"], J=GRID["J"], CT=ct))
    xs, ys, desc = cpu_sample_points(nxy, side)
    n = len(xs) * len(ys)
    for _ in range(warmup):
        orc.region(xs[:2], ys[:2], nthreads=cores)      # warm the pool; a bounded sample pass is the step
    t0 = time.perf_counter()
    nl = 0
    for _ in range(steps):
        conv, count, err, liou = orc.region(xs, ys, nthreads=cores)
        nl += int(liou.sum())
    dt = time.perf_counter() - t0
    return dict(value=steps * n / dt, unit="points/s", cores=cores, kind="port",
                sample=desc + "; oracle/rg_oracle.c (direct-convolution port of the reference's numpy/scipy path), "
                "OpenMP over points, %d liouville evaluations in %.1f s" % (nl, dt)), dt / steps, n


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, ms, n = run_cpu(a.ct, a.nxy, a.cpu_sample, a.steps, min(a.warmup, 1))
    line = {"impl": "reference", "metric": "region grid points/sec (Nxy=%d, L=J=10)" % a.nxy, "value": cb["value"],
            "unit": "points/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "region Nxy=%d L=J=10 golden-mean torus, %s; each step = %s" % (a.nxy, a.ct, cb["sample"])},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- our arm
def other_transformations(a, device):
    """The three canonical transformations of the reference side by side (RG_dict's default is expm_multiply) on the
    64x64 strided sub-grid of the same region, device-timed, outside the headline's timed region.  A per-point work
    budget bounds the few points where expm_multiply's scaling parameter explodes (the reference would spend
    CPU-hours on each); they are counted in `budget_capped`."""
    from rg_b200 import RG, make_param_dict
    from rg_b200 import rg_modules
    side = 64
    stride = a.nxy // side
    out = {}
    os.environ["RG_B200_MAX_LIOUVILLE"] = "300000"
    try:
        for ct in ("expm_onestep", "expm_multiply", "expm_adapt"):
            case = RG(make_param_dict(L=GRID["L"], J=GRID["J"], Nxy=a.nxy, AmpInf=GRID["AmpInf"], AmpSup=GRID["AmpSup"],
                                      CanonicalTransformation=ct, SaveData=False), device=device)
            xs = np.linspace(case.AmpInf[0], case.AmpSup[0], a.nxy)[stride // 2::stride]
            ys = np.linspace(case.AmpInf[1], case.AmpSup[1], a.nxy)[stride // 2::stride]
            amps = rg_modules.region_amps(case, xs, ys)
            case.points(amps[:296])                              # warm-up (plan upload, first launch)
            conv, count, err = case.points(amps)
            st = case.last_stats
            out[ct] = {"value": len(amps) / (st.kernel_ms * 1e-3), "unit": "points/s",
                       "sample": "%dx%d strided sub-grid (ix,iy = %d, %d, ...) of the Nxy=%d region, work budget 300000 liouville per point"
                                 % (side, side, stride // 2, stride // 2 + stride, a.nxy),
                       "liouville_per_point": st.liouville / len(amps),
                       "useful_tflops": 2.0 * st.conv_fma_useful / st.kernel_ms * 1e-9,
                       "budget_capped": int((err == -98).sum()), "converged": int(conv.sum()), "count_sum": int(count.sum())}
            case.close()
    finally:
        os.environ.pop("RG_B200_MAX_LIOUVILLE", None)
    return out


def main_b200(a):
    import torch
    import torch.distributed as dist
    from rg_b200 import RG, make_param_dict
    from rg_b200 import rg_modules

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    # stdout carries exactly one JSON line: everything else written to fd 1 while the job runs (NCCL prints its
    # version banner there at the first collective) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    case = RG(make_param_dict(L=GRID["L"], J=GRID["J"], Nxy=a.nxy, AmpInf=GRID["AmpInf"], AmpSup=GRID["AmpSup"],
                              CanonicalTransformation=a.ct, SaveData=False), device=local)
    x_vec, y_vec = (np.linspace(case.AmpInf[_], case.AmpSup[_], case.Nxy, dtype=case.Precision) for _ in range(2))
    amps = rg_modules.region_amps(case, x_vec, y_vec)
    n_total = amps.shape[0]
    idx = rg_modules.shard_indices(n_total, rank, world)
    if a.emulate_world > 1:
        idx = rg_modules.shard_indices(n_total, 0, a.emulate_world)
        n_total = len(idx)
    n_loc = len(idx)
    eff_world = a.emulate_world if a.emulate_world > 1 else world
    row_len = a.nxy // eff_world if (a.nxy % eff_world == 0 and a.neighbour_queue) else 0
    # host pinned buffers (the public-API path copies them in and out inside the timed region)
    h_amps = torch.from_numpy(amps[idx].copy()).pin_memory()
    h_out = [torch.zeros(n_loc, dtype=torch.int32).pin_memory() for _ in range(3)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    info = case.plan_info()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        flush.fill_(1)                       # L2 flush between iterations
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        case.points(h_amps, *h_out, row_len=row_len)   # H2D + kernel + D2H through the C ABI
        loc = np.stack([t.numpy() for t in h_out], axis=1)
        full = rg_modules.gather_sharded(loc, n_total, rank, world)
        wall = time.perf_counter() - t0
        st = case.last_stats
        return wall, st.kernel_ms * 1e-3, st, full

    for _ in range(a.warmup):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t_begin = time.perf_counter()
    wall_sum = dev_sum = 0.0
    liou = useful = executed = 0
    full = None
    for _ in range(a.steps):
        w, d, st, full = step()
        wall_sum += w
        dev_sum += d
        liou += st.liouville
        useful += st.conv_fma_useful
        executed += st.conv_fma_executed
    barrier()
    bracket = time.perf_counter() - t_begin
    sampler.stop_flag = True
    sampler.join(timeout=2)

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    dev_max, wall_max, bracket_max = reduce_max(dev_sum), reduce_max(wall_sum), reduce_max(bracket)
    liou_all = reduce_sum(liou)
    if rank == 0:
        peak = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak_r01.json")))["dfma_tflops"]
        ach = 2.0 * useful / dev_sum * 1e-12            # this rank's kernel, this rank's time
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_bench_traffic_r01.json")
        if world == 1 and a.nxy == GRID["Nxy"] and a.ct == "expm_onestep" and os.path.exists(tpath):
            tj = json.load(open(tpath))                 # ncu capture of this very launch (one launch per step)
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        line = {
            "metric": "region grid points/sec (Nxy=%d, L=J=10)" % a.nxy,
            "value": a.steps * n_total / dev_max, "unit": "points/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": bracket_max / a.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "region Nxy=%d, L=J=10, golden-mean torus N=[[1,1],[1,0]], K=((0,1,0),(0,1,1)), "
                                   "AmpSup=[0.12,0.35], TolMin=1e-10, TolMax=1e4, MaxLie=5000, NormChoice=sum, %s; "
                                   "%d points sharded cyclically over %d rank(s)" % (a.nxy, a.ct, n_total, world),
                       "canonical_transformation": a.ct, "l2_flush_between_steps": True,
                       "threads_per_cta": info["threads_per_cta"], "smem_bytes_per_cta": info["smem_bytes"],
                       "ctas_per_sm": info["ctas_per_sm"]},
            "e2e": {"value": a.steps * n_total / wall_max, "unit": "points/s",
                    "h2d_bytes_per_step": int(h_amps.numel() * 8), "d2h_bytes_per_step": int(3 * n_loc * 4)},
            "gpu_launches": a.steps * world,
            "roofline": {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                         "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write, profiles/ncu_bench_traffic_r01.json; "
                                                             "compute-bound kernel: 0.08 % of the HBM peak, L2 write-back of the workspace)",
                         "kernel": "rg_kernel<10, false> (fused converge: conv_all + series epilogue)",
                         "peak_source": "measured DFMA peak, tools/fp64_peak.cu on this pool (profiles/fp64_peak_r01.json)",
                         "liouville_evaluations": int(liou_all),
                         "useful_flop_per_liouville": 2 * info["fma_useful_per_liouville"],
                         "issued_flop_per_liouville": 2 * info["fma_executed_per_liouville"],
                         "issued_tflops": 2.0 * executed / dev_sum * 1e-12,
                         "dense_equivalent_tflops": liou * DENSE_FLOP_PER_LIOUVILLE / dev_sum * 1e-12},
            "clocks": sampler.summary(),
            "result_checksum": {"converged": int(full[:, 0].sum()), "count_sum": int(full[:, 1].sum()),
                                "errors": {str(k): int(v) for k, v in zip(*np.unique(full[:, 2], return_counts=True))}},
        }
        if world == 1 and not a.no_cpu_baseline:
            cb, _, _ = run_cpu(a.ct, a.nxy, a.cpu_sample, 1, 0)
            line["cpu_baseline"] = cb
        if world == 1 and not a.no_other_ct:
            line["other_transformations"] = other_transformations(a, local)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    case.close()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_b200(args)
