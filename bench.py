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
    ap.add_argument("--port", action="store_true", help="reference arm: time the C port (oracle/rg_oracle.c) instead of the reference files")
    ap.add_argument("--dump-map", type=int, default=0, help="reference arm: add the sample's (conv, count, error) map to the JSON line")
    ap.add_argument("--config", type=int, default=4, choices=[3, 4, 5],
                    help="BASELINE.json config: 4 = region Nxy=256 L=J=10 (the headline, default); 3 = surface Nxy=128 L=J=10 "
                         "sharded over the ranks; 5 = spiral-mean n=3 region L=J=6 expm_multiply (sub-grid, see --sub)")
    ap.add_argument("--sub", type=int, default=32, help="config 5: side of the strided sub-grid of the Nxy=128 region (128 = all)")
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
# The reference's own CPU path (RG_modules.point -> RG.converge -> ... -> scipy.signal.convolve, RG_modules.py:121-143):
# the UNMODIFIED reference files, copied by __graft_entry__.build() into the git-ignored oracle/_ref/ so that they
# travel to the GPU box, loaded through oracle/ref_loader.py (matplotlib stub + CWD holding theta_taylor.mat) and
# fanned out over a process pool like RG_modules.compute_region does (one task per point, chunksize 1, stock order).
SUB16 = 16   # the 16x16 strided subsample ix, iy in {8, 24, ..., 248} of the Nxy=256 grid


def sub16_points(nxy):
    stride = nxy // SUB16
    idx = np.arange(stride // 2, nxy, stride)
    xs = np.linspace(GRID["AmpInf"][0], GRID["AmpSup"][0], nxy)[idx]
    ys = np.linspace(GRID["AmpInf"][1], GRID["AmpSup"][1], nxy)[idx]
    pts = [(float(x), float(y)) for y in ys for x in xs]          # y is the outer loop (RG_modules.py:138-139)
    flat = np.array([iy * nxy + ix for iy in idx for ix in idx])
    return pts, flat, "16x16 strided subsample (ix,iy in {%d,%d,...,%d}) of the Nxy=%d grid" % (idx[0], idx[1], idx[-1], nxy)


def reference_dir():
    d = os.path.join(ROOT, "oracle", "_ref")
    return d if os.path.isfile(os.path.join(d, "RG.py")) and os.path.isfile(os.path.join(d, "theta_taylor.mat")) else None


_REF = {}


def _ref_init(ct, refdir):
    """Pool worker: build the reference's RG object once (np.random.seed(12345) per worker)."""
    import warnings
    warnings.simplefilter("ignore")
    os.environ["RG_REFERENCE_DIR"] = refdir
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    ref_loader.REFERENCE_DIR = refdir
    RG, RG_modules, RG_dict = ref_loader.load()
    np.random.seed(12345)
    prm = ref_loader.make_params(L=GRID["L"], J=GRID["J"], CT=ct, AmpInf=[float(v) for v in GRID["AmpInf"]],
                                 AmpSup=[float(v) for v in GRID["AmpSup"]], Nxy=GRID["Nxy"])
    _REF["case"] = RG.RG(prm)
    _REF["point"] = RG_modules.point


def _ref_point(xy):
    t0 = time.perf_counter()
    (conv, count), err = _REF["point"](xy[0], xy[1], _REF["case"])
    return int(conv), int(count), int(err), time.perf_counter() - t0


def _port_init(ct, _unused):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import rg_oracle as ro
    _REF["orc"] = ro.OracleRG(ro.default_params(L=GRID["L"], J=GRID["J"], CT=ct))


def _port_point(xy):
    t0 = time.perf_counter()
    conv, count, err, _ = _REF["orc"].points([xy], nthreads=1)
    return int(conv[0]), int(count[0]), int(err[0]), time.perf_counter() - t0


def run_reference(ct, nxy, steps, warmup, kind=None):
    """K timed steps: the sample's points in stock order, cut into K contiguous chunks (step s = chunk s), streamed
    through ONE pool.imap(chunksize=1) -- no barrier between steps, as on the full grid where a row of 256 points
    keeps every core busy.  Returns the JSON fields of the run."""
    import multiprocessing as mp
    refdir = reference_dir()
    if kind is None:
        kind = "reference" if refdir else "port"
    cores = len(os.sched_getaffinity(0))          # never from OpenMP: torchrun exports OMP_NUM_THREADS=1
    pts, flat, desc = sub16_points(nxy)
    lim = int(os.environ.get("RG_BENCH_REF_LIMIT", "0"))      # smoke tests of this script only
    if lim > 0:
        pts, flat, desc = pts[-lim:], flat[-lim:], desc + " [last %d points only]" % lim
    init, fn = (_ref_init, _ref_point) if kind == "reference" else (_port_init, _port_point)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=init, initargs=(ct, refdir)) as pool:
        for _ in range(max(1, warmup)):           # warm-up step: one cheap point (diverges in one rg_map) per worker
            pool.map(fn, [pts[-1]] * cores, chunksize=1)
        bounds = np.linspace(0, len(pts), steps + 1).round().astype(int)
        t0 = time.perf_counter()
        res, step_end = [], []
        for i, r in enumerate(pool.imap(fn, pts, chunksize=1)):
            res.append(r)
            while len(step_end) < steps and i + 1 >= bounds[len(step_end) + 1]:
                step_end.append(time.perf_counter() - t0)
        wall = time.perf_counter() - t0
    res = np.array(res, dtype=np.float64)
    core_s = float(res[:, 3].sum())
    out = dict(value=len(pts) / wall, unit="points/s", cores=cores, kind=kind, wall_s=wall,
               core_seconds_per_point=core_s / len(pts), slowest_point_s=float(res[:, 3].max()),
               balanced_points_per_s=cores * len(pts) / core_s,
               extrapolated_full_grid_hours=(nxy * nxy / (len(pts) / wall)) / 3600.0,
               sample=desc + ", %s, every point once over the %d step(s); " % (ct, steps) +
               ("the unmodified reference (oracle/_ref: RG.py, RG_modules.py point(), scipy FFT convolution)" if kind == "reference"
                else "oracle/rg_oracle.c (direct-convolution C port; the reference files are absent)") +
               ", fork pool of %d processes (os.sched_getaffinity), imap chunksize=1, stock order; %.1f core-s in %.1f s wall"
               % (cores, core_s, wall))
    gpath = os.path.join(ROOT, "tests", "golden", "region_L10J10_sub16_%s.npz" % ct)
    if nxy == GRID["Nxy"] and os.path.exists(gpath) and lim <= 0:
        g = np.load(gpath)
        ok = (np.array_equal(res[:, 0].reshape(16, 16), g["conv"]) and np.array_equal(res[:, 1].reshape(16, 16), g["count"])
              and np.array_equal(res[:, 2].reshape(16, 16), g["error"]))
        out["matches_frozen_golden"] = bool(ok)
    return out, res[:, :3].astype(np.int64), flat, wall / steps


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, smap, flat, per_step = run_reference(a.ct, a.nxy, a.steps, a.warmup, kind=("port" if a.port else None))
    line = {"impl": "reference", "metric": "region grid points/sec (Nxy=%d, L=J=10)" % a.nxy, "value": cb["value"],
            "unit": "points/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": per_step * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "region Nxy=%d, L=J=10, golden-mean torus N=[[1,1],[1,0]], K=((0,1,0),(0,1,1)), AmpSup=[0.12,0.35], "
                                   "TolMin=1e-10, TolMax=1e4, MaxLie=5000, NormChoice=sum, %s; bounded sample: %s" % (a.nxy, a.ct, cb["sample"]),
                       "canonical_transformation": a.ct},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if a.dump_map:
        line["sample_map"] = {"flat_index": flat.tolist(), "conv_count_error": smap.tolist()}
    print(json.dumps(line))


def cpu_baseline_subprocess(ct, nxy):
    """The reference arm in a child process (this one holds a CUDA context; the pool forks).  Returns its JSON line."""
    import subprocess
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ct", ct, "--nxy", str(nxy), "--dump-map", "1"], capture_output=True, text=True,
                       env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
    for ln in r.stdout.splitlines()[::-1]:
        if ln.startswith("{"):
            return json.loads(ln)
    raise RuntimeError("reference arm failed: " + r.stderr[-2000:])


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
        case.points(h_amps, *h_out)   # H2D + kernel + D2H through the C ABI
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
        tpath = os.path.join(ROOT, "profiles", "ncu_bench_traffic_r02.json")
        if not os.path.exists(tpath):
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
                         "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write, %s; compute-bound kernel: "
                                                             "< 0.1 %% of the HBM peak, L2 write-back of the workspace)" % os.path.relpath(tpath, ROOT),
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
        if world == 1 and not a.no_cpu_baseline and a.emulate_world <= 1:
            # the reference's CPU path on this box's host cores, same run, and the parity gate of
            # the GPU map restricted to the sample must equal the CPU map bit for bit
            ref = cpu_baseline_subprocess(a.ct, a.nxy)
            line["cpu_baseline"] = ref["cpu_baseline"]
            sm = ref.get("sample_map")
            if sm:
                cpu_map = np.array(sm["conv_count_error"], dtype=np.int64)
                gpu_map = full[np.array(sm["flat_index"])].astype(np.int64)
                bad = np.flatnonzero((cpu_map != gpu_map).any(axis=1))
                line["parity_gate"] = {"points": int(len(cpu_map)), "mismatches": int(len(bad)),
                                       "mismatching_flat_indices": [int(sm["flat_index"][b]) for b in bad],
                                       "what": "(conv, count, error) of the GPU map at the CPU sample's points vs the "
                                               "%s arm of this run" % ref["cpu_baseline"]["kind"]}
        if world == 1 and not a.no_other_ct:
            line["other_transformations"] = other_transformations(a, local)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    case.close()


# ---------------------------------------------------------------------------------------------- configs 3 and 5
def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return torch, dist, world, rank, local, json_fd


def main_config3(a):
    """BASELINE config 3: compute_surface, Nxy=128 columns of the critical surface, L=J=10, RelDist=1e-7, columns sharded
    cyclically over the ranks, one all_gather of the float64[2][Nxy] result (RG_modules.py:92-119).  One step = the
    whole surface.  value = columns/s (wall clock between barriers, max over ranks: the two launches per rank --
    batched converge of the brackets, batched on-device bisection -- are synchronous ABI calls)."""
    torch, dist, world, rank, local, json_fd = _dist_setup()
    from rg_b200 import RG, make_param_dict
    from rg_b200 import rg_modules
    import io, contextlib
    nxy = 128
    case = RG(make_param_dict(L=GRID["L"], J=GRID["J"], Nxy=nxy, AmpInf=GRID["AmpInf"], AmpSup=GRID["AmpSup"], RelDist=1e-7,
                              CanonicalTransformation=a.ct, SaveData=False, Method="surface"), device=local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step():
        with contextlib.redirect_stdout(io.StringIO()):
            return rg_modules.compute_surface(case)
    for _ in range(a.warmup):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    data = None
    for _ in range(a.steps):
        data = step()
    barrier()
    dt = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    if rank == 0:
        import hashlib
        line = {"metric": "critical-surface columns/sec (Nxy=128, L=J=10, RelDist=1e-7)", "value": a.steps * nxy / dt, "unit": "columns/s",
                "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "BASELINE config 3: compute_surface Nxy=128, L=J=10, golden mean, %s, RelDist=1e-7; %d columns per rank "
                                       "(cyclic), 2 launches per rank (bracket converge, on-device bisection)" % (a.ct, (nxy + world - 1) // world)},
                "gpu_launches": 2 * a.steps * world, "clocks": sampler.summary(),
                "result": {"columns_with_threshold": int(np.isfinite(data[1]).sum()), "eps2_min": float(np.nanmin(data[1])),
                           "eps2_max": float(np.nanmax(data[1])), "sha1_of_data": hashlib.sha1(np.ascontiguousarray(data).tobytes()).hexdigest()}}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    case.close()


def main_config5(a):
    """BASELINE config 5: spiral-mean torus n=3, region Nxy=128, L=J=6, expm_multiply -- the global-record (BIG) kernel.
    One step = the --sub x --sub strided sub-grid of the region (default 32x32 = 1024 points; 128 = the whole grid),
    sharded cyclically over the ranks."""
    torch, dist, world, rank, local, json_fd = _dist_setup()
    from rg_b200 import RG, make_param_dict
    from rg_b200 import rg_modules
    from rg_b200.rg_dict import SPIRAL_MEAN
    nxy, side = 128, a.sub
    stride = nxy // side
    case = RG(make_param_dict(L=6, J=6, Nxy=nxy, CanonicalTransformation=a.ct, SaveData=False, **SPIRAL_MEAN), device=local)
    xs = np.linspace(case.AmpInf[0], case.AmpSup[0], nxy)[stride // 2::stride]
    ys = np.linspace(case.AmpInf[1], case.AmpSup[1], nxy)[stride // 2::stride]
    amps = rg_modules.region_amps(case, xs, ys)
    idx = rg_modules.shard_indices(len(amps), rank, world)
    h_amps = torch.from_numpy(amps[idx].copy()).pin_memory()
    h_out = [torch.zeros(len(idx), dtype=torch.int32).pin_memory() for _ in range(3)]
    info = case.plan_info()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step():
        t0 = time.perf_counter()
        case.points(h_amps, *h_out)
        loc = np.stack([t.numpy() for t in h_out], axis=1)
        full = rg_modules.gather_sharded(loc, len(amps), rank, world)
        return time.perf_counter() - t0, case.last_stats, full
    for _ in range(a.warmup):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t_begin = time.perf_counter()
    wall = dev = 0.0
    liou = useful = executed = 0
    for _ in range(a.steps):
        w, st, full = step()
        wall += w
        dev += st.kernel_ms * 1e-3
        liou += st.liouville
        useful += st.conv_fma_useful
        executed += st.conv_fma_executed
    barrier()
    bracket = time.perf_counter() - t_begin
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if world > 1:
        t = torch.tensor([dev, wall, bracket], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_max, wall_max, bracket = (float(v) for v in t.tolist())
    else:
        dev_max, wall_max = dev, wall
    if rank == 0:
        peak = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak_r01.json")))["dfma_tflops"]
        ach = 2.0 * useful / dev * 1e-12
        line = {"metric": "region grid points/sec (spiral mean n=3, Nxy=128, L=J=6, %s)" % a.ct, "value": a.steps * len(amps) / dev_max,
                "unit": "points/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": bracket / a.steps * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "BASELINE config 5: spiral-mean torus N=[[0,1,0],[0,0,1],[1,0,-1]], region Nxy=128, L=J=6, %s; "
                                       "%dx%d strided sub-grid (ix,iy = %d, %d, ...) = %d points over %d rank(s)"
                                       % (a.ct, side, side, stride // 2, stride // 2 + stride, len(amps), world),
                           "threads_per_cta": info["threads_per_cta"], "smem_bytes_per_cta": info["smem_bytes"],
                           "ctas_per_sm": info["ctas_per_sm"], "records": "global (BIG kernel): 246 KB of g records per Hamiltonian"},
                "e2e": {"value": a.steps * len(amps) / wall_max, "unit": "points/s", "h2d_bytes_per_step": int(h_amps.numel() * 8),
                        "d2h_bytes_per_step": int(3 * len(idx) * 4)},
                "gpu_launches": a.steps * world,
                "roofline": {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                             "kernel": "rg_kernel<6, true> (global-record variant)", "liouville_evaluations": int(liou),
                             "useful_flop_per_liouville": 2 * info["fma_useful_per_liouville"],
                             "issued_flop_per_liouville": 2 * info["fma_executed_per_liouville"],
                             "issued_tflops": 2.0 * executed / dev * 1e-12,
                             "dense_equivalent_tflops": liou * 229.4e6 / dev * 1e-12,
                             "peak_source": "measured DFMA peak, tools/fp64_peak.cu (profiles/fp64_peak_r01.json)"},
                "clocks": sampler.summary(),
                "result_checksum": {"converged": int(full[:, 0].sum()), "count_sum": int(full[:, 1].sum()),
                                    "errors": {str(k): int(v) for k, v in zip(*np.unique(full[:, 2], return_counts=True))}}}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    case.close()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        main_reference(args)
    elif args.config == 3:
        main_config3(args)
    elif args.config == 5:
        if "--ct" not in sys.argv and "RG_BENCH_CT" not in os.environ:
            args.ct = "expm_multiply"          # BASELINE config 5 names it
        main_config5(args)
    else:
        main_b200(args)
