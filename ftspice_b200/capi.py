"""ctypes binding of libftsb.so (the C-ABI declared in include/ftsb.h).

The library is the product's only compute path.  If it is missing, importing this module raises: there is
no Python / CPU fallback.
"""
from __future__ import annotations

import ctypes as ct
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libftsb.so")

# enums of include/ftsb.h
MEM_HOST, MEM_DEVICE = 0, 1
LAYOUT_INSTANCE_MAJOR, LAYOUT_PARAM_MAJOR = 0, 1
OK, NOT_CONVERGED, DIVERGED, TRAN_H_UNDERFLOW, NMOS_UNREACHABLE = range(5)
STATUS = {0: "OK", 1: "NOT_CONVERGED", 2: "DIVERGED", 3: "TRAN_H_UNDERFLOW", 4: "NMOS_UNREACHABLE"}
E_ARG, E_CUDA, E_UNSUPPORTED, E_STATE = -1, -2, -3, -4

EXPORTS = [
    "ftsb_abi_version", "ftsb_device_count", "ftsb_last_error", "ftsb_compile", "ftsb_destroy", "ftsb_n_elems",
    "ftsb_n_run", "ftsb_n_start", "ftsb_n_fn", "ftsb_set_option", "ftsb_set_params", "ftsb_run_op", "ftsb_run_dc",
    "ftsb_run_tran", "ftsb_sync", "ftsb_stream", "ftsb_set_stream", "ftsb_get_counters", "ftsb_launch_count",
    "ftsb_last_variant", "ftsb_lu_solve", "ftsb_measure_fp64_tflops", "ftsb_solve_op", "ftsb_solve_dc",
    "ftsb_multi_compile", "ftsb_multi_destroy", "ftsb_multi_device_count", "ftsb_multi_last_error", "ftsb_multi_last_variant",
    "ftsb_multi_set_option", "ftsb_multi_solve_op", "ftsb_multi_solve_dc", "ftsb_multi_solve_tran", "ftsb_multi_get_counters",
    "ftsb_multi_launch_count",
]


class FtsbElem(ct.Structure):
    _fields_ = [("kind", ct.c_int32), ("node", ct.c_int32 * 3), ("branch", ct.c_int32), ("fn_kind", ct.c_int32)]


class FtsbTranOpts(ct.Structure):
    _fields_ = [("capacity", ct.c_int64), ("rec_stride", ct.c_int32), ("reserved", ct.c_int32)]


class FtsbCounters(ct.Structure):
    _fields_ = [("solve_points", ct.c_int64), ("newton_iters", ct.c_int64), ("lu_factor", ct.c_int64),
                ("lu_solve", ct.c_int64), ("rej_newton", ct.c_int64), ("rej_plte", ct.c_int64),
                ("redo", ct.c_int64), ("sparse_flops", ct.c_int64)]


class FtsbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ftsb error {code}: {msg}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile libftsb.so in-tree with nvcc for sm_100a (ftspice_b200/csrc/Makefile)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc")] + (["-B"] if force else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libftsb.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


_lib = None


def lib() -> ct.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = ct.CDLL(LIB_PATH)
    vp, i32, i64, d = ct.c_void_p, ct.c_int32, ct.c_int64, ct.c_double
    L.ftsb_abi_version.restype = ct.c_int
    L.ftsb_device_count.restype = ct.c_int
    L.ftsb_last_error.argtypes = [vp]
    L.ftsb_last_error.restype = ct.c_char_p
    L.ftsb_compile.argtypes = [ct.POINTER(FtsbElem), i32, i32, i32, ct.POINTER(i32), ct.c_int, ct.POINTER(vp)]
    L.ftsb_compile.restype = ct.c_int
    L.ftsb_destroy.argtypes = [vp]
    L.ftsb_destroy.restype = None
    for f in ("ftsb_n_elems", "ftsb_n_run", "ftsb_n_start", "ftsb_n_fn"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = i32
    L.ftsb_set_option.argtypes = [vp, ct.c_char_p, i64]
    L.ftsb_set_option.restype = ct.c_int
    L.ftsb_set_params.argtypes = [vp, i64, vp, vp, ct.c_int, ct.c_int]
    L.ftsb_set_params.restype = ct.c_int
    L.ftsb_run_op.argtypes = [vp, vp, vp, vp, ct.c_int]
    L.ftsb_run_op.restype = ct.c_int
    L.ftsb_run_dc.argtypes = [vp, i32, vp, vp, vp, i64, vp, vp, vp, vp, ct.c_int]
    L.ftsb_run_dc.restype = ct.c_int
    L.ftsb_solve_op.argtypes = [vp, i64, vp, vp, ct.c_int, vp, vp, vp]
    L.ftsb_solve_op.restype = ct.c_int
    L.ftsb_solve_dc.argtypes = [vp, i64, vp, vp, ct.c_int, i32, vp, vp, vp, i64, vp, vp, vp, vp]
    L.ftsb_solve_dc.restype = ct.c_int
    L.ftsb_run_tran.argtypes = [vp, d, d, ct.POINTER(FtsbTranOpts), vp, vp, vp, vp, vp, vp, ct.c_int]
    L.ftsb_run_tran.restype = ct.c_int
    L.ftsb_sync.argtypes = [vp]
    L.ftsb_sync.restype = ct.c_int
    L.ftsb_stream.argtypes = [vp]
    L.ftsb_stream.restype = vp
    L.ftsb_set_stream.argtypes = [vp, vp]
    L.ftsb_set_stream.restype = ct.c_int
    L.ftsb_get_counters.argtypes = [vp, ct.POINTER(FtsbCounters)]
    L.ftsb_get_counters.restype = ct.c_int
    L.ftsb_launch_count.argtypes = [vp]
    L.ftsb_launch_count.restype = i64
    L.ftsb_last_variant.argtypes = [vp]
    L.ftsb_last_variant.restype = ct.c_char_p
    L.ftsb_lu_solve.argtypes = [ct.c_int, i32, i64, vp, vp, vp, ct.c_int]
    L.ftsb_lu_solve.restype = ct.c_int
    L.ftsb_measure_fp64_tflops.argtypes = [ct.c_int]
    L.ftsb_measure_fp64_tflops.restype = d
    L.ftsb_multi_compile.argtypes = [ct.POINTER(FtsbElem), i32, i32, i32, ct.POINTER(i32), ct.POINTER(i32), i32, ct.POINTER(vp)]
    L.ftsb_multi_compile.restype = ct.c_int
    L.ftsb_multi_destroy.argtypes = [vp]
    L.ftsb_multi_destroy.restype = None
    L.ftsb_multi_device_count.argtypes = [vp]
    L.ftsb_multi_device_count.restype = i32
    L.ftsb_multi_last_error.argtypes = [vp]
    L.ftsb_multi_last_error.restype = ct.c_char_p
    L.ftsb_multi_last_variant.argtypes = [vp]
    L.ftsb_multi_last_variant.restype = ct.c_char_p
    L.ftsb_multi_set_option.argtypes = [vp, ct.c_char_p, i64]
    L.ftsb_multi_set_option.restype = ct.c_int
    L.ftsb_multi_solve_op.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    L.ftsb_multi_solve_op.restype = ct.c_int
    L.ftsb_multi_solve_dc.argtypes = [vp, i64, vp, vp, i32, vp, vp, vp, i64, vp, vp, vp, vp]
    L.ftsb_multi_solve_dc.restype = ct.c_int
    L.ftsb_multi_solve_tran.argtypes = [vp, i64, vp, vp, d, d, ct.POINTER(FtsbTranOpts), vp, vp, vp, vp, vp, vp]
    L.ftsb_multi_solve_tran.restype = ct.c_int
    L.ftsb_multi_get_counters.argtypes = [vp, ct.POINTER(FtsbCounters)]
    L.ftsb_multi_get_counters.restype = ct.c_int
    L.ftsb_multi_launch_count.argtypes = [vp]
    L.ftsb_multi_launch_count.restype = i64
    _lib = L
    return L


def check(rc: int, handle=None) -> None:
    if rc != 0:
        msg = lib().ftsb_last_error(handle)
        raise FtsbError(rc, msg.decode() if msg else "")


def check_multi(rc: int, handle=None) -> None:
    if rc != 0:
        msg = lib().ftsb_multi_last_error(handle)
        raise FtsbError(rc, msg.decode() if msg else "")
