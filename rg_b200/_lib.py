"""ctypes binding of librg_b200.so (C ABI declared in include/rg_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device is visible, every compute call
raises.  `build()` compiles the library in-tree for sm_100a (nvcc cross-compiles without a GPU).
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "librg_b200.so")
_LIB = None

c_int_p = ctypes.POINTER(ctypes.c_int32)
c_dbl_p = ctypes.POINTER(ctypes.c_double)
vp = ctypes.c_void_p


class RGParams(ctypes.Structure):
    """rg_params of include/rg_b200.h (field order and types must match)."""
    _fields_ = [("dim", ctypes.c_int32), ("L", ctypes.c_int32), ("J", ctypes.c_int32),
                ("norm_choice", ctypes.c_int32), ("norm_analytic", ctypes.c_double),
                ("tol_min", ctypes.c_double), ("tol_max", ctypes.c_double),
                ("max_lie", ctypes.c_int32), ("ct", ctypes.c_int32),
                ("min_step", ctypes.c_double), ("abs_tol", ctypes.c_double), ("rel_tol", ctypes.c_double),
                ("eigenvalue", ctypes.c_double),
                ("omega0_nu", vp), ("iminus", vp), ("reidx", vp), ("theta", vp),
                ("n_theta", ctypes.c_int32), ("device", ctypes.c_int32), ("max_series", ctypes.c_int32),
                ("max_liouville", ctypes.c_int64)]


class RGStats(ctypes.Structure):
    """rg_stats of include/rg_b200.h."""
    _fields_ = [("liouville", ctypes.c_int64), ("lie_steps", ctypes.c_int64), ("rg_maps", ctypes.c_int64),
                ("conv_fma_executed", ctypes.c_int64), ("conv_fma_useful", ctypes.c_int64),
                ("clk_conv", ctypes.c_int64), ("clk_epilogue", ctypes.c_int64), ("clk_total", ctypes.c_int64),
                ("clk_conv_busy", ctypes.c_int64),
                ("kernel_ms", ctypes.c_float), ("launches", ctypes.c_int32), ("liouville_full", ctypes.c_int64)]


# every symbol include/rg_b200.h declares, with its ctypes signature
SIGNATURES = {
    "rg_last_error": (ctypes.c_char_p, []),
    "rg_device_count": (ctypes.c_int, [c_int_p]),
    "rg_plan_create": (ctypes.c_int, [ctypes.POINTER(RGParams), ctypes.POINTER(vp)]),
    "rg_plan_destroy": (ctypes.c_int, [vp]),
    "rg_plan_info": (ctypes.c_int, [vp, c_int_p, c_int_p, c_int_p, c_int_p, ctypes.POINTER(ctypes.c_int64),
                                    ctypes.POINTER(ctypes.c_int64), c_int_p]),
    "rg_tables_create": (ctypes.c_int, [vp, ctypes.c_int32, vp, vp, ctypes.POINTER(vp)]),
    "rg_tables_destroy": (ctypes.c_int, [vp]),
    "rg_region": (ctypes.c_int, [vp, vp, ctypes.c_int32, vp, vp, ctypes.c_int32, vp, vp, vp, ctypes.POINTER(RGStats)]),
    "rg_region_work": (ctypes.c_int, [vp, vp, ctypes.c_int32, vp, vp, ctypes.c_int32, vp, vp, vp, vp, ctypes.POINTER(RGStats)]),
    "rg_converge": (ctypes.c_int, [vp, vp, ctypes.c_int32, vp, vp, vp, vp, vp, ctypes.POINTER(RGStats)]),
    "rg_approach": (ctypes.c_int, [vp, vp, ctypes.c_int32, vp, vp, ctypes.c_double, vp, vp, vp, vp, vp, ctypes.POINTER(RGStats)]),
    "rg_map": (ctypes.c_int, [vp, vp, ctypes.c_int32, vp, vp, ctypes.c_int32, vp, vp, ctypes.POINTER(RGStats)]),
    "rg_expm": (ctypes.c_int, [vp, ctypes.c_int32, ctypes.c_int32, vp, vp, vp, vp, vp, vp, ctypes.c_double, vp, vp,
                               ctypes.POINTER(RGStats)]),
    "rg_liouville": (ctypes.c_int, [vp, ctypes.c_int32, vp, vp, vp, vp, vp, vp]),
    "rg_conv_product": (ctypes.c_int, [vp, ctypes.c_int32, vp, vp, vp]),
    "rg_generate_y": (ctypes.c_int, [vp, ctypes.c_int32, vp, vp, vp, vp, vp, vp]),
    "rg_norm": (ctypes.c_int, [vp, ctypes.c_int32, vp, ctypes.c_int32, vp]),
    "rg_sym": (ctypes.c_int, [vp, ctypes.c_int32, vp, vp]),
    "rg_reindex_rescale": (ctypes.c_int, [vp, vp, ctypes.c_int32, vp, vp, vp]),
    "rg_compute_m_s": (ctypes.c_int, [vp, ctypes.c_int32, vp, vp, vp, vp, ctypes.c_double, ctypes.c_double, vp, vp, vp]),
    "rg_pnorm": (ctypes.c_int, [vp, ctypes.c_int32, vp, vp, vp, vp, ctypes.c_int32, vp]),
    "rg_debug_schedule": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, vp, ctypes.c_int32,
                                         c_int_p, c_int_p, ctypes.POINTER(c_int_p), ctypes.POINTER(c_int_p), c_int_p,
                                         ctypes.POINTER(ctypes.POINTER(ctypes.c_uint32)), c_int_p,
                                         ctypes.POINTER(c_int_p), ctypes.POINTER(ctypes.c_int64), c_int_p, c_int_p]),
    "rg_debug_piece_runs": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, vp, ctypes.c_int32,
                                           ctypes.POINTER(c_int_p), c_int_p, c_int_p]),
    "rg_debug_free": (None, [vp]),
}


def build(force=False, verbose=False):
    """Compile librg_b200.so in-tree for sm_100a."""
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in ("rg_b200.cu", "rg_kernel_inst.cu", "rg_kernel.cuh", "rg_device.cuh", "rg_multiply.cuh",
                                            "rg_plan.hpp", "Makefile")]
    srcs.append(os.path.join(_HERE, "..", "include", "rg_b200.h"))
    if not force and os.path.exists(SO_PATH) and all(os.path.getmtime(SO_PATH) >= os.path.getmtime(s) for s in srcs):
        return SO_PATH
    if force:
        subprocess.run(["make", "-C", csrc, "clean"], capture_output=True, text=True)
    r = subprocess.run(["make", "-C", csrc], capture_output=True, text=True)   # one translation unit per Taylor order, in parallel
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("nvcc build of librg_b200.so failed; see rg_b200/csrc/build.log")
    return SO_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError("librg_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                               "rg_b200 has no CPU fallback")
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


class RGError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = lib().rg_last_error()
        raise RGError("librg_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def device_count():
    n = ctypes.c_int32(0)
    rc = lib().rg_device_count(ctypes.byref(n))
    return n.value if rc == 0 else 0
