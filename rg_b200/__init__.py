"""rg_b200 -- B200-native renormalization-group map, drop-in for the hot path of cchandre/RG.

    from rg_b200 import RG, Hamiltonian, param_dict, compute_region, compute_surface, compute_line

The Python surface mirrors the reference (RG.py, RG_modules.py, RG_dict.py); the numerics run in
hand-written sm_100a kernels behind the C ABI of include/rg_b200.h (librg_b200.so, built in-tree).
"""
from .rg import RG, Hamiltonian  # noqa: F401
from .rg_dict import param_dict, make_param_dict  # noqa: F401
from .rg_modules import (compute_iterates, compute_line, compute_surface, compute_region, compute_cr,  # noqa: F401
                         point, save_data, plot_fun)
from ._lib import build, RGError, device_count  # noqa: F401

__all__ = ["RG", "Hamiltonian", "param_dict", "make_param_dict", "compute_iterates", "compute_line",
           "compute_surface", "compute_region", "compute_cr", "point", "save_data", "plot_fun", "build",
           "RGError", "device_count", "main"]


def main(parameters=None):
    """RG.py:39-42"""
    case = RG(param_dict if parameters is None else parameters)
    return {"compute_iterates": compute_iterates, "compute_line": compute_line,
            "compute_surface": compute_surface, "compute_region": compute_region}[case.Method](case)
