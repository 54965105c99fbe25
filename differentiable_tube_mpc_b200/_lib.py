"""ctypes binding of the C ABI declared in include/dtmpc.h (libdtmpc_b200.so, built in-tree).

There is no CPU fallback: if the CUDA library is missing, loading fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DTMPC_LIB selects an alternative build of the same ABI (kernel tuning experiments); default: the in-tree library
LIB_PATH = os.environ.get("DTMPC_LIB") or os.path.join(_HERE, "libdtmpc_b200.so")

MAX_OBS = 16
MAX_ALPHAS = 8
NTHETA = 12

AGG_SMOOTHMIN, AGG_MIN, AGG_SINGLE = 0, 1, 2
OK, E_BADARG, E_WORKSPACE, E_CUDA = 0, 1, 2, 3
ST_NONFINITE, ST_NO_CANDIDATE = 1, 2


class Env(C.Structure):
    """struct dtmpc_env"""
    _fields_ = [
        ("dt", C.c_double),
        ("u_min", C.c_double * 2),
        ("u_max", C.c_double * 2),
        ("active_tol", C.c_double),
        ("n_obs", C.c_int32),
        ("obs_agg", C.c_int32),
        ("obs_cx", C.c_double * MAX_OBS),
        ("obs_cy", C.c_double * MAX_OBS),
        ("obs_r", C.c_double * MAX_OBS),
        ("obs_beta", C.c_double),
    ]


class Ocp(C.Structure):
    """struct dtmpc_ocp"""
    _fields_ = [
        ("horizon", C.c_int32),
        ("max_iter", C.c_int32),
        ("n_alpha", C.c_int32),
        ("angle_wrap", C.c_int32),
        ("tol", C.c_double),
        ("reg", C.c_double),
        ("alphas", C.c_double * MAX_ALPHAS),
        ("eps_dyn", C.c_double),
        ("eps_jac", C.c_double),
        ("tight_in_jac", C.c_int32),
        ("reserved", C.c_int32),
    ]


class SolveIO(C.Structure):
    """struct dtmpc_solve_io"""
    _fields_ = [
        ("x0", C.c_void_p), ("V_init", C.c_void_p), ("theta", C.c_void_p), ("theta_stride", C.c_int64),
        ("target", C.c_void_p), ("target_stride", C.c_int64), ("Xref", C.c_void_p), ("Uref", C.c_void_p),
        ("X", C.c_void_p), ("V", C.c_void_p), ("K", C.c_void_p), ("kff", C.c_void_p),
        ("iters", C.c_void_p), ("status", C.c_void_p), ("alpha_idx", C.c_void_p),
        ("costs", C.c_void_p), ("margins", C.c_void_p), ("fail_at", C.c_void_p),
    ]


class SensIO(C.Structure):
    """struct dtmpc_sens_io"""
    _fields_ = [
        ("X", C.c_void_p), ("V", C.c_void_p), ("theta", C.c_void_p), ("theta_stride", C.c_int64),
        ("target", C.c_void_p), ("target_stride", C.c_int64), ("Xref", C.c_void_p), ("Uref", C.c_void_p),
        ("xref_row", C.c_int64), ("gx", C.c_void_p), ("gu", C.c_void_p), ("Xbar", C.c_void_p),
        ("dX", C.c_void_p), ("dV", C.c_void_p), ("dL", C.c_void_p), ("g_theta", C.c_void_p),
        ("gXref", C.c_void_p), ("gUref", C.c_void_p), ("loss", C.c_void_p), ("status", C.c_void_p),
        ("gx_row", C.c_int64),
    ]


class Disturbance(C.Structure):
    """struct dtmpc_disturbance"""
    _fields_ = [("low", C.c_double * 3), ("high", C.c_double * 3), ("seed", C.c_uint64), ("first_instance", C.c_int64),
                ("step", C.c_void_p)]


# every symbol include/dtmpc.h declares (tests/test_abi.py checks the header against this list)
_PER_DTYPE = ("ilqr_solve", "tube_solve", "tube_step", "sensitivity", "ift_accumulate", "reduce_columns", "rollout", "primitives",
              "plant_step")
EXPORTS = ["dtmpc_abi_version", "dtmpc_last_error", "dtmpc_kernel_launches", "dtmpc_solve_workspace_bytes", "dtmpc_sens_workspace_bytes", "dtmpc_step_workspace_bytes",
           "dtmpc_reduce_workspace_bytes", "dtmpc_measure_fp64_peak_tflops", "dtmpc_measure_fp32_peak_tflops",
           "dtmpc_math_selftest_f64"] + [
    f"dtmpc_{n}_{s}" for n in _PER_DTYPE for s in ("f64", "f32")]

ABI_VERSION = 3   # include/dtmpc.h DTMPC_ABI_VERSION
_lib = None


class LibraryMissing(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libdtmpc_b200.so; raises LibraryMissing (never falls back to a CPU path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). differentiable_tube_mpc_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, sz = C.c_void_p, C.c_int64, C.c_int32, C.c_size_t
    pe, po = C.POINTER(Env), C.POINTER(Ocp)
    L.dtmpc_abi_version.restype = C.c_int
    L.dtmpc_last_error.restype = C.c_char_p
    L.dtmpc_kernel_launches.restype = C.c_uint64
    L.dtmpc_solve_workspace_bytes.restype = sz
    L.dtmpc_solve_workspace_bytes.argtypes = [i32, i64, i32, i32]
    L.dtmpc_step_workspace_bytes.restype = sz
    L.dtmpc_step_workspace_bytes.argtypes = [i32, i64, i32, i32]
    L.dtmpc_sens_workspace_bytes.restype = sz
    L.dtmpc_sens_workspace_bytes.argtypes = [i32, i64, i32]
    L.dtmpc_reduce_workspace_bytes.restype = sz
    L.dtmpc_reduce_workspace_bytes.argtypes = [i64, i32, i32]
    for s in ("f64", "f32"):
        f = getattr(L, f"dtmpc_ilqr_solve_{s}")
        f.restype, f.argtypes = C.c_int, [pe, po, i64, C.POINTER(SolveIO), vp, sz, vp]
        f = getattr(L, f"dtmpc_tube_solve_{s}")
        f.restype, f.argtypes = C.c_int, [pe, po, po, i64, C.POINTER(SolveIO), C.POINTER(SolveIO), vp, sz, vp]
        f = getattr(L, f"dtmpc_tube_step_{s}")
        f.restype, f.argtypes = C.c_int, [pe, po, po, i64, C.POINTER(SolveIO), C.POINTER(SolveIO), C.POINTER(SensIO),
                                          C.POINTER(SensIO), vp, sz, vp]
        f = getattr(L, f"dtmpc_sensitivity_{s}")
        f.restype, f.argtypes = C.c_int, [pe, po, i64, C.POINTER(SensIO), vp, sz, vp]
        f = getattr(L, f"dtmpc_ift_accumulate_{s}")
        f.restype, f.argtypes = C.c_int, [pe, po, i64, C.POINTER(SensIO), vp]
        f = getattr(L, f"dtmpc_reduce_columns_{s}")
        f.restype, f.argtypes = C.c_int, [vp, i64, i32, vp, vp, sz, vp]
        f = getattr(L, f"dtmpc_rollout_{s}")
        f.restype, f.argtypes = C.c_int, [pe, po, i64, vp, vp, vp, i64, vp, vp]
        f = getattr(L, f"dtmpc_primitives_{s}")
        f.restype, f.argtypes = C.c_int, [pe, po, i64, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        f = getattr(L, f"dtmpc_plant_step_{s}")
        f.restype, f.argtypes = C.c_int, [pe, po, po, i64, vp, vp, vp, i64, vp, i64, vp, C.POINTER(Disturbance), vp, vp, vp,
                                          vp, vp]
    L.dtmpc_math_selftest_f64.restype = C.c_int
    L.dtmpc_math_selftest_f64.argtypes = [vp, i64, vp, vp]
    L.dtmpc_measure_fp64_peak_tflops.restype = C.c_double
    L.dtmpc_measure_fp64_peak_tflops.argtypes = [i32, vp]
    L.dtmpc_measure_fp32_peak_tflops.restype = C.c_double
    L.dtmpc_measure_fp32_peak_tflops.argtypes = [i32, vp]
    if L.dtmpc_abi_version() != ABI_VERSION:
        raise RuntimeError("libdtmpc_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != OK:
        msg = load().dtmpc_last_error().decode("utf-8", "replace")
        kind = {E_BADARG: ValueError, E_WORKSPACE: MemoryError}.get(rc, RuntimeError)
        raise kind(f"{what} failed (code {rc}): {msg}")
