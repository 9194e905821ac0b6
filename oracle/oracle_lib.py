"""Binding of the CPU oracle (oracle/liboracle.so, prefix orc_) to the same ctypes harness as the CUDA library.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. Nothing under breeze.jl_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import breeze_b200  # noqa: E402
from breeze_b200 import abi  # noqa: E402

LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    inc = os.path.join(os.path.dirname(_HERE), "include")
    import glob
    deps = glob.glob(os.path.join(_HERE, "*.c")) + glob.glob(os.path.join(_HERE, "*.h")) + [os.path.join(_HERE, "Makefile")]
    deps += glob.glob(os.path.join(inc, "*.h"))
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(d) for d in deps):
        return LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB_PATH


_LIB = None


def load_oracle_library() -> abi.Library:
    global _LIB
    if _LIB is None:
        build()
        _LIB = abi.Library(LIB_PATH, "orc_", cuda=False)
        d = _LIB.dll
        d.orc_saturation_specific_humidity.restype = C.c_double
        d.orc_saturation_specific_humidity.argtypes = [C.POINTER(abi.bz_config), C.c_double, C.c_double, C.c_double]
        d.orc_saturation_vapor_pressure.restype = C.c_double
        d.orc_saturation_vapor_pressure.argtypes = [C.POINTER(abi.bz_config), C.c_double, C.c_double]
        d.orc_density.restype = C.c_double
        d.orc_density.argtypes = [C.POINTER(abi.bz_config), C.c_double, C.c_double, C.c_double]
        d.orc_weno5_biased.restype = C.c_double
        d.orc_weno5_biased.argtypes = [C.POINTER(C.c_double)]
        d.orc_weno3_biased.restype = C.c_double
        d.orc_weno3_biased.argtypes = [C.POINTER(C.c_double)]
        d.orc_weno_biased_window.restype = C.c_double
        d.orc_weno_biased_window.argtypes = [C.POINTER(C.c_double), C.c_int]
        d.orc_centered_window.restype = C.c_double
        d.orc_centered_window.argtypes = [C.POINTER(C.c_double), C.c_int]
        d.orc_num_threads.restype = C.c_int
        d.orc_set_beta_form.argtypes = [C.c_int]
        d.orc_set_num_threads.argtypes = [C.c_int]
    return _LIB


class CPUOracle:
    """Architecture object for breeze_b200.RectilinearGrid(architecture=...) that routes the host mirror to the oracle."""
    device, rank, n_ranks, nccl_unique_id, use_tma, z_chunks = 0, 0, 1, None, 0, 0

    def library(self):
        return load_oracle_library()


def set_beta_form(form: int):
    """0: the reference's quadratic-form smoothness indicators (default); 1: difference form (what the CUDA kernels use)."""
    load_oracle_library().dll.orc_set_beta_form(int(form))
