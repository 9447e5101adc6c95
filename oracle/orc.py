"""ctypes binding of the CPU oracle (oracle/libftsp_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the product package.
"""
from __future__ import annotations

import ctypes as ct
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libftsp_oracle.so")

STATUS = {0: "OK", 1: "NOT_CONVERGED", 2: "DIVERGED", 3: "TRAN_H_UNDERFLOW", 4: "NMOS_UNREACHABLE"}


class OrcElem(ct.Structure):
    _fields_ = [("kind", ct.c_int32), ("node", ct.c_int32 * 3), ("branch", ct.c_int32), ("fn_kind", ct.c_int32),
                ("val", ct.c_double), ("fn", ct.c_double * 7)]


class OrcCircuit(ct.Structure):
    _fields_ = [("n_elems", ct.c_int32), ("n_run", ct.c_int32), ("n_start", ct.c_int32), ("pad_", ct.c_int32),
                ("elems", ct.POINTER(OrcElem)), ("cls", ct.POINTER(ct.c_int32)),
                ("order_run", ct.POINTER(ct.c_int32)), ("order_start", ct.POINTER(ct.c_int32))]


class OrcStats(ct.Structure):
    _fields_ = [("newton_iters", ct.c_int64), ("lu_solves", ct.c_int64), ("rej_newton", ct.c_int64),
                ("rej_plte", ct.c_int64)]


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc -O2 -ffp-contract=off)."""
    src = os.path.join(_HERE, "ftsp_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> ct.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ct.CDLL(_LIB_PATH)
        d, i64, i32 = ct.c_double, ct.c_int64, ct.c_int32
        P = ct.POINTER
        _lib.orc_lu_solve.argtypes = [ct.c_int, P(d), P(d), P(d)]
        _lib.orc_lu_solve.restype = None
        _lib.orc_norm.argtypes = [P(d), ct.c_int]
        _lib.orc_norm.restype = d
        _lib.orc_dampen_step.argtypes = [P(d), P(d), ct.c_int]
        _lib.orc_dampen_step.restype = None
        _lib.orc_converged.argtypes = [d] * 8
        _lib.orc_converged.restype = ct.c_int
        _lib.orc_unrolled_dot.argtypes = [P(d), P(d), ct.c_int]
        _lib.orc_unrolled_dot.restype = d
        _lib.orc_spice_fn_eval.argtypes = [ct.c_int, P(d), d]
        _lib.orc_spice_fn_eval.restype = d
        _lib.orc_stamp_linear.argtypes = [P(OrcElem), ct.c_int, P(d), P(d), ct.c_int]
        _lib.orc_stamp_linear.restype = None
        _lib.orc_diode_model.argtypes = [d, d, P(d), P(d), P(d)]
        _lib.orc_diode_model.restype = None
        _lib.orc_npn_model.argtypes = [d, d, d, P(d)]
        _lib.orc_npn_model.restype = None
        _lib.orc_nmos_model.argtypes = [d, d, d, P(d)]
        _lib.orc_nmos_model.restype = ct.c_int
        _lib.orc_elem_apply.argtypes = [P(OrcElem), d, d, ct.c_int, ct.c_int, P(d), d, P(d), P(d)]
        _lib.orc_elem_apply.restype = ct.c_int
        _lib.orc_elem_funcs.argtypes = [P(OrcElem), ct.c_int, ct.c_int, ct.c_int, P(d), P(d), P(d)]
        _lib.orc_elem_funcs.restype = ct.c_int
        _lib.orc_get_err_raw.argtypes = [ct.c_int, ct.c_int, P(d), P(d), P(d), P(d), P(d), P(d)]
        _lib.orc_get_err_raw.restype = None
        _lib.orc_dc_count.argtypes = [d, d, d]
        _lib.orc_dc_count.restype = i64
        _lib.orc_run_op.argtypes = [P(OrcCircuit), P(d), P(i64), P(OrcStats)]
        _lib.orc_run_dc.argtypes = [P(OrcCircuit), ct.c_int, d, d, d, P(d), P(i64), P(d), P(i64), P(OrcStats)]
        _lib.orc_run_tran.argtypes = [P(OrcCircuit), d, d, i64, P(d), P(d), P(i64), P(i64), P(d), P(OrcStats)]
        _lib.orc_batch_op.argtypes = [P(OrcCircuit), i64, P(d), P(d), P(d), P(i64), P(i32), ct.c_int]
        _lib.orc_batch_dc.argtypes = [P(OrcCircuit), i64, P(d), P(d), ct.c_int, P(d), P(d), P(d), i64, P(d),
                                      P(i64), P(i64), P(i32), ct.c_int]
        _lib.orc_batch_tran.argtypes = [P(OrcCircuit), i64, P(d), P(d), d, d, i64, P(d), P(d), P(i64), P(i64),
                                        P(d), P(i32), P(OrcStats), ct.c_int]
    return _lib


def _dp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ct.POINTER(ct.c_double))


def _ip(a: Optional[np.ndarray], t):
    return None if a is None else a.ctypes.data_as(ct.POINTER(t))


class Oracle:
    """One compiled-for-the-oracle circuit (keeps the ctypes arrays alive)."""

    def __init__(self, circ):
        self.circ = circ
        ne = circ.n_elems
        self._elems = (OrcElem * max(ne, 1))()
        for i in range(ne):
            e = self._elems[i]
            e.kind = int(circ.kind[i])
            for j in range(3):
                e.node[j] = int(circ.node[i, j])
            e.branch = int(circ.branch[i])
            e.fn_kind = int(circ.fn_kind[i])
            e.val = float(circ.vals[i])
            for j in range(7):
                e.fn[j] = float(circ.fns[i, j])
        self._cls = np.ascontiguousarray(circ.cls, np.int32)
        self._or = np.ascontiguousarray(circ.order_run, np.int32)
        self._os = np.ascontiguousarray(circ.order_start, np.int32)
        self.c = OrcCircuit(ne, circ.n_run, circ.n_start, 0, self._elems, _ip(self._cls, ct.c_int32),
                            _ip(self._or, ct.c_int32), _ip(self._os, ct.c_int32))

    # ---- single instance (nominal values) ----
    def run_op(self):
        x = np.zeros(self.circ.n_start)
        it = ct.c_int64(0)
        st = OrcStats()
        s = lib().orc_run_op(ct.byref(self.c), _dp(x), ct.byref(it), ct.byref(st))
        return s, int(it.value), x, st

    def run_dc(self, sweep_elem: int, start: float, stop: float, step: float):
        npts = int(lib().orc_dc_count(start, stop, step))
        x = np.zeros((max(npts, 1), self.circ.n_run))
        it = np.zeros(max(npts, 1), np.int64)
        sw = np.zeros(max(npts, 1))
        done = ct.c_int64(0)
        st = OrcStats()
        s = lib().orc_run_dc(ct.byref(self.c), sweep_elem, start, stop, step, _dp(x), _ip(it, ct.c_int64), _dp(sw),
                             ct.byref(done), ct.byref(st))
        k = int(done.value)
        return s, npts, it[:k], x[:k], sw[:k], st

    def run_tran(self, stop: float, step_max: float, cap: int = 1 << 16):
        t = np.zeros(cap)
        x = np.zeros((cap, self.circ.n_run))
        it = np.zeros(cap, np.int64)
        nrec = ct.c_int64(0)
        xf = np.zeros(self.circ.n_run)
        st = OrcStats()
        s = lib().orc_run_tran(ct.byref(self.c), stop, step_max, cap, _dp(t), _dp(x), _ip(it, ct.c_int64),
                               ct.byref(nrec), _dp(xf), ct.byref(st))
        k = min(int(nrec.value), cap)
        return s, int(nrec.value), it[:k], t[:k], x[:k], xf, st

    # ---- batches (per-instance values) ----
    def batch_op(self, vals: np.ndarray, fns: Optional[np.ndarray] = None, threads: int = 1):
        vals = np.ascontiguousarray(vals, np.float64)
        B = vals.shape[0]
        fns = None if fns is None else np.ascontiguousarray(fns, np.float64)
        x = np.zeros((B, self.circ.n_start))
        it = np.zeros(B, np.int64)
        stt = np.zeros(B, np.int32)
        lib().orc_batch_op(ct.byref(self.c), B, _dp(vals), _dp(fns), _dp(x), _ip(it, ct.c_int64),
                           _ip(stt, ct.c_int32), threads)
        return stt, it, x

    def batch_dc(self, vals, fns, sweep_elem, start, stop, step, max_points, threads: int = 1):
        vals = np.ascontiguousarray(vals, np.float64)
        B = vals.shape[0]
        fns = None if fns is None else np.ascontiguousarray(fns, np.float64)
        start = np.ascontiguousarray(start, np.float64)
        stop = np.ascontiguousarray(stop, np.float64)
        step = np.ascontiguousarray(step, np.float64)
        x = np.zeros((B, max_points, self.circ.n_run))
        it = np.zeros((B, max_points), np.int64)
        done = np.zeros(B, np.int64)
        stt = np.zeros(B, np.int32)
        lib().orc_batch_dc(ct.byref(self.c), B, _dp(vals), _dp(fns), sweep_elem, _dp(start), _dp(stop), _dp(step),
                           max_points, _dp(x), _ip(it, ct.c_int64), _ip(done, ct.c_int64), _ip(stt, ct.c_int32),
                           threads)
        return stt, done, it, x

    def batch_tran(self, vals, fns, stop, step_max, cap, threads: int = 1, want_wave: bool = True):
        vals = np.ascontiguousarray(vals, np.float64)
        B = vals.shape[0]
        fns = None if fns is None else np.ascontiguousarray(fns, np.float64)
        t = np.zeros((B, cap)) if want_wave else None
        x = np.zeros((B, cap, self.circ.n_run)) if want_wave else None
        it = np.zeros((B, cap), np.int64) if want_wave else None
        nrec = np.zeros(B, np.int64)
        xf = np.zeros((B, self.circ.n_run))
        stt = np.zeros(B, np.int32)
        stats = (OrcStats * B)()
        lib().orc_batch_tran(ct.byref(self.c), B, _dp(vals), _dp(fns), stop, step_max, cap, _dp(t), _dp(x),
                             _ip(it, ct.c_int64) if it is not None else None, _ip(nrec, ct.c_int64), _dp(xf),
                             _ip(stt, ct.c_int32), stats, threads)
        return stt, nrec, it, t, x, xf, stats
