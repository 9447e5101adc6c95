"""Host-side mirror of the reference's ``Engine`` (src/engine.rs:22-206) on top of the C-ABI.

``BatchEngine`` is the batched form (one topology, many instances); ``Engine`` mirrors the reference's
single-netlist interface — ``Engine(elems, cmds)``, ``run_op()``, ``run_dc()``, ``run_tran()`` returning a
``SimResult`` whose ``print()`` output has the reference's CSV format (src/engine/sim_result.rs:24-36) and
whose failures raise ``NotConvergedError`` (src/engine/error.rs:3-10).  All arithmetic happens in
libftsb.so (CUDA); nothing here computes.
"""
from __future__ import annotations

import ctypes as ct
import math
from dataclasses import dataclass
from decimal import Decimal
from typing import List, Optional, Sequence, Union

import numpy as np

from . import capi
from . import netlist as nl


class NotConvergedError(RuntimeError):
    """src/engine/error.rs:3-10"""

    def __init__(self, status: int = capi.NOT_CONVERGED):
        super().__init__("Couldn't converge simulation.")
        self.status = status


class EnginePanic(RuntimeError):
    """The reference process would have panicked (newtons_method.rs:53-55, nmos/model.rs:38-40)."""

    def __init__(self, status: int):
        super().__init__({capi.DIVERGED: "v_err or i_err diverged",
                          capi.NMOS_UNREACHABLE: "internal error: entered unreachable code"}.get(status, str(status)))
        self.status = status


def rust_f64_str(x: float) -> str:
    """Rust's ``f64::to_string``: shortest round-trip digits, positional, integers without '.0'."""
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    s = format(Decimal(repr(float(x))), "f")
    if "." in s:
        s = s.rstrip("0").rstrip(".")
    if s in ("0", "-0"):
        return "-0" if math.copysign(1.0, x) < 0 else "0"
    return s


class SimResult:
    """src/engine/sim_result.rs:6-46"""

    def __init__(self, headers: Sequence[str]):
        self.headers = list(headers)
        self.data: List[List[float]] = []

    def push(self, record: dict) -> None:
        self.data.append([record[h] for h in self.headers])

    def get(self, label: str) -> np.ndarray:
        idx = self.headers.index(label)
        return np.array([row[idx] for row in self.data])

    def to_csv(self) -> str:
        lines = [",".join(self.headers)]
        for row in self.data:
            lines.append(",".join(rust_f64_str(v) for v in row))
        return "\n".join(lines) + "\n"

    def print(self) -> None:
        print(self.to_csv(), end="")


@dataclass
class OpResult:
    status: np.ndarray    # [B] int32
    n_iters: np.ndarray   # [B] int32
    x: np.ndarray         # [B, n_start]


@dataclass
class DcResult:
    status: np.ndarray       # [B]
    points_done: np.ndarray  # [B] int64
    n_iters: np.ndarray      # [B, P] int32
    x: np.ndarray            # [B, P, n_run]
    sweep: np.ndarray        # [B, P] sweep values start + k*step


@dataclass
class TranResult:
    status: np.ndarray   # [B]
    n_rec: np.ndarray    # [B] int64, ALL accepted steps
    n_iters: Optional[np.ndarray]  # [B, cap]
    t: Optional[np.ndarray]        # [B, cap]
    x: Optional[np.ndarray]        # [B, cap, n_run]
    x_final: np.ndarray  # [B, n_run]


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ct.c_void_p)


class BatchEngine:
    """One compiled topology on one GPU; runs batches of independent instances."""

    def __init__(self, circ: nl.Circuit, device: int = 0):
        self.circ = circ
        self.device = device
        L = capi.lib()
        ne = circ.n_elems
        elems = (capi.FtsbElem * max(ne, 1))()
        for i in range(ne):
            elems[i].kind = int(circ.kind[i])
            for j in range(3):
                elems[i].node[j] = int(circ.node[i, j])
            elems[i].branch = int(circ.branch[i])
            elems[i].fn_kind = int(circ.fn_kind[i])
        cls = np.ascontiguousarray(circ.cls, np.int32)
        h = ct.c_void_p()
        rc = L.ftsb_compile(elems, ne, circ.n_run, circ.n_start, cls.ctypes.data_as(ct.POINTER(ct.c_int32)), device,
                            ct.byref(h))
        capi.check(rc, None)
        self.h = h
        self.n_fn = int(L.ftsb_n_fn(h))
        self.fn_elems = [i for i in range(ne) if circ.kind[i] in (nl.V, nl.I) and circ.fn_kind[i] != nl.FN_NONE]
        self.batch = 0

    def close(self) -> None:
        if getattr(self, "h", None):
            capi.lib().ftsb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- options / introspection ----
    def set_option(self, name: str, value: int) -> None:
        capi.check(capi.lib().ftsb_set_option(self.h, name.encode(), int(value)), self.h)

    @property
    def variant(self) -> str:
        return capi.lib().ftsb_last_variant(self.h).decode()

    @property
    def launch_count(self) -> int:
        return int(capi.lib().ftsb_launch_count(self.h))

    @property
    def stream(self) -> int:
        return int(capi.lib().ftsb_stream(self.h) or 0)

    def sync(self) -> None:
        capi.check(capi.lib().ftsb_sync(self.h), self.h)

    def counters(self) -> dict:
        c = capi.FtsbCounters()
        capi.check(capi.lib().ftsb_get_counters(self.h, ct.byref(c)), self.h)
        return {k: int(getattr(c, k)) for k in ("solve_points", "newton_iters", "lu_factor", "lu_solve", "rej_newton",
                                                "rej_plte", "redo", "sparse_flops")}

    # ---- parameters ----
    def compact_fns(self, fns: np.ndarray) -> np.ndarray:
        """[B, n_elems, 7] -> [B, n_fn, 7] (waveform sources in element order)."""
        return np.ascontiguousarray(fns[:, self.fn_elems, :], np.float64)

    def set_params(self, vals: Optional[np.ndarray] = None, fns: Optional[np.ndarray] = None, batch: Optional[int] = None
                   ) -> None:
        """vals [B, n_elems]; fns [B, n_elems, 7] or already compact [B, n_fn, 7].  None = nominal netlist values."""
        c = self.circ
        if vals is None:
            B = batch or 1
            vals = np.broadcast_to(c.vals, (B, c.n_elems))
        vals = np.ascontiguousarray(vals, np.float64)
        B = vals.shape[0]
        if fns is None:
            fns = np.broadcast_to(c.fns, (B, c.n_elems, 7))
        fns = np.asarray(fns)
        if fns.shape[1] == c.n_elems:  # full per-element table -> waveform sources only
            fns = fns[:, self.fn_elems, :]
        elif fns.shape[1] != self.n_fn:
            raise ValueError(f"fns must be [B, {c.n_elems}, 7] or [B, {self.n_fn}, 7]")
        fns = np.ascontiguousarray(fns, np.float64)
        self._keep = (vals, fns)
        capi.check(capi.lib().ftsb_set_params(self.h, B, _ptr(vals), _ptr(fns) if self.n_fn else None,
                                              capi.LAYOUT_INSTANCE_MAJOR, capi.MEM_HOST), self.h)
        self.batch = B

    def set_params_device(self, batch: int, vals_ptr: int, fns_ptr: int, layout: int = capi.LAYOUT_INSTANCE_MAJOR) -> None:
        capi.check(capi.lib().ftsb_set_params(self.h, batch, ct.c_void_p(vals_ptr), ct.c_void_p(fns_ptr) if fns_ptr else None,
                                              layout, capi.MEM_DEVICE), self.h)
        self.batch = batch

    # ---- analyses (host buffers) ----
    def run_op(self) -> OpResult:
        B, n = self.batch, self.circ.n_start
        x = np.zeros((B, n))
        it = np.zeros(B, np.int32)
        st = np.zeros(B, np.int32)
        capi.check(capi.lib().ftsb_run_op(self.h, _ptr(x), _ptr(it), _ptr(st), capi.MEM_HOST), self.h)
        return OpResult(st, it, x)

    def run_dc(self, source: Union[str, int], start, stop, step, max_points: Optional[int] = None) -> DcResult:
        B, n = self.batch, self.circ.n_run
        se = self.circ.elem_index(source) if isinstance(source, str) else int(source)
        start = np.ascontiguousarray(np.broadcast_to(np.asarray(start, np.float64), (B,)))
        stop = np.ascontiguousarray(np.broadcast_to(np.asarray(stop, np.float64), (B,)))
        step = np.ascontiguousarray(np.broadcast_to(np.asarray(step, np.float64), (B,)))
        counts = np.ceil((stop - start) / step)
        counts = np.where(counts >= 0, counts, 0).astype(np.int64)
        P = int(max_points or max(int(counts.max()), 1))
        x = np.zeros((B, P, n))
        it = np.zeros((B, P), np.int32)
        done = np.zeros(B, np.int64)
        st = np.zeros(B, np.int32)
        capi.check(capi.lib().ftsb_run_dc(self.h, se, _ptr(start), _ptr(stop), _ptr(step), P, _ptr(x), _ptr(it),
                                          _ptr(done), _ptr(st), capi.MEM_HOST), self.h)
        sweep = start[:, None] + step[:, None] * np.arange(P, dtype=np.float64)[None, :]
        return DcResult(st, done, it, x, sweep)

    def run_tran(self, stop: float, step: float, capacity: int = 0, rec_stride: int = 1, want_wave: bool = True
                 ) -> TranResult:
        B, n = self.batch, self.circ.n_run
        cap = int(capacity) if want_wave else 0
        t = np.zeros((B, cap)) if cap else None
        x = np.zeros((B, cap, n)) if cap else None
        it = np.zeros((B, cap), np.int32) if cap else None
        nrec = np.zeros(B, np.int64)
        xf = np.zeros((B, n))
        st = np.zeros(B, np.int32)
        opts = capi.FtsbTranOpts(cap, rec_stride, 0)
        capi.check(capi.lib().ftsb_run_tran(self.h, float(stop), float(step), ct.byref(opts), _ptr(t), _ptr(x), _ptr(it),
                                            _ptr(nrec), _ptr(xf), _ptr(st), capi.MEM_HOST), self.h)
        return TranResult(st, nrec, it, t, x, xf)

    # ---- parameters + analysis in one call (host buffers; chunk-pipelined copies once the handle is warm) ----
    def _host_params(self, vals: np.ndarray, fns: Optional[np.ndarray]):
        c = self.circ
        vals = np.ascontiguousarray(vals, np.float64)
        B = vals.shape[0]
        if fns is None:
            fns = np.broadcast_to(c.fns, (B, c.n_elems, 7))
        fns = np.asarray(fns)
        if fns.shape[1] == c.n_elems:
            fns = fns[:, self.fn_elems, :]
        elif fns.shape[1] != self.n_fn:
            raise ValueError(f"fns must be [B, {c.n_elems}, 7] or [B, {self.n_fn}, 7]")
        return vals, np.ascontiguousarray(fns, np.float64), B

    def solve_op(self, vals: np.ndarray, fns: Optional[np.ndarray] = None) -> OpResult:
        """ftsb_solve_op: same results as set_params(vals, fns) followed by run_op()."""
        vals, fns, B = self._host_params(vals, fns)
        n = self.circ.n_start
        x = np.zeros((B, n))
        it = np.zeros(B, np.int32)
        st = np.zeros(B, np.int32)
        capi.check(capi.lib().ftsb_solve_op(self.h, B, _ptr(vals), _ptr(fns) if self.n_fn else None,
                                            capi.LAYOUT_INSTANCE_MAJOR, _ptr(x), _ptr(it), _ptr(st)), self.h)
        self.batch = B
        return OpResult(st, it, x)

    def solve_dc(self, vals: np.ndarray, fns: Optional[np.ndarray], source: Union[str, int], start, stop, step,
                 max_points: Optional[int] = None) -> DcResult:
        """ftsb_solve_dc: same results as set_params(vals, fns) followed by run_dc(...)."""
        vals, fns, B = self._host_params(vals, fns)
        n = self.circ.n_run
        se = self.circ.elem_index(source) if isinstance(source, str) else int(source)
        start = np.ascontiguousarray(np.broadcast_to(np.asarray(start, np.float64), (B,)))
        stop = np.ascontiguousarray(np.broadcast_to(np.asarray(stop, np.float64), (B,)))
        step = np.ascontiguousarray(np.broadcast_to(np.asarray(step, np.float64), (B,)))
        counts = np.ceil((stop - start) / step)
        counts = np.where(counts >= 0, counts, 0).astype(np.int64)
        P = int(max_points or max(int(counts.max()), 1))
        x = np.zeros((B, P, n))
        it = np.zeros((B, P), np.int32)
        done = np.zeros(B, np.int64)
        st = np.zeros(B, np.int32)
        capi.check(capi.lib().ftsb_solve_dc(self.h, B, _ptr(vals), _ptr(fns) if self.n_fn else None,
                                            capi.LAYOUT_INSTANCE_MAJOR, se, _ptr(start), _ptr(stop), _ptr(step), P,
                                            _ptr(x), _ptr(it), _ptr(done), _ptr(st)), self.h)
        self.batch = B
        sweep = start[:, None] + step[:, None] * np.arange(P, dtype=np.float64)[None, :]
        return DcResult(st, done, it, x, sweep)

    # ---- analyses (device pointers; asynchronous on the handle's stream) ----
    def run_op_device(self, x_ptr: int, it_ptr: int, st_ptr: int) -> None:
        capi.check(capi.lib().ftsb_run_op(self.h, ct.c_void_p(x_ptr), ct.c_void_p(it_ptr), ct.c_void_p(st_ptr),
                                          capi.MEM_DEVICE), self.h)

    def run_dc_device(self, sweep_elem: int, start_ptr: int, stop_ptr: int, step_ptr: int, max_points: int, x_ptr: int,
                      it_ptr: int, done_ptr: int, st_ptr: int) -> None:
        v = ct.c_void_p
        capi.check(capi.lib().ftsb_run_dc(self.h, sweep_elem, v(start_ptr), v(stop_ptr), v(step_ptr), max_points, v(x_ptr),
                                          v(it_ptr), v(done_ptr), v(st_ptr), capi.MEM_DEVICE), self.h)

    def run_tran_device(self, stop: float, step: float, capacity: int, rec_stride: int, t_ptr: int, x_ptr: int, it_ptr: int,
                        nrec_ptr: int, xf_ptr: int, st_ptr: int) -> None:
        v = lambda p: ct.c_void_p(p) if p else None
        opts = capi.FtsbTranOpts(capacity, rec_stride, 0)
        capi.check(capi.lib().ftsb_run_tran(self.h, float(stop), float(step), ct.byref(opts), v(t_ptr), v(x_ptr), v(it_ptr),
                                            v(nrec_ptr), v(xf_ptr), v(st_ptr), capi.MEM_DEVICE), self.h)


class MultiEngine:
    """One compiled topology on a LIST of GPUs (ftsb_multi_*): the batch is sharded contiguously inside the library, one
    host thread per device, results written straight into the caller's arrays — the device count does not show in the
    output layout.  Host buffers only."""

    def __init__(self, circ: nl.Circuit, devices: Optional[Sequence[int]] = None):
        self.circ = circ
        L = capi.lib()
        ne = circ.n_elems
        elems = (capi.FtsbElem * max(ne, 1))()
        for i in range(ne):
            elems[i].kind = int(circ.kind[i])
            for j in range(3):
                elems[i].node[j] = int(circ.node[i, j])
            elems[i].branch = int(circ.branch[i])
            elems[i].fn_kind = int(circ.fn_kind[i])
        cls = np.ascontiguousarray(circ.cls, np.int32)
        devs = np.ascontiguousarray(devices if devices is not None else [], np.int32)
        h = ct.c_void_p()
        rc = L.ftsb_multi_compile(elems, ne, circ.n_run, circ.n_start, cls.ctypes.data_as(ct.POINTER(ct.c_int32)),
                                  devs.ctypes.data_as(ct.POINTER(ct.c_int32)) if len(devs) else None, len(devs), ct.byref(h))
        capi.check_multi(rc, None)
        self.h = h
        self.fn_elems = [i for i in range(ne) if circ.kind[i] in (nl.V, nl.I) and circ.fn_kind[i] != nl.FN_NONE]
        self.n_fn = len(self.fn_elems)

    def close(self) -> None:
        if getattr(self, "h", None):
            capi.lib().ftsb_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_devices(self) -> int:
        return int(capi.lib().ftsb_multi_device_count(self.h))

    @property
    def variant(self) -> str:
        return capi.lib().ftsb_multi_last_variant(self.h).decode()

    @property
    def launch_count(self) -> int:
        return int(capi.lib().ftsb_multi_launch_count(self.h))

    def set_option(self, name: str, value: int) -> None:
        capi.check_multi(capi.lib().ftsb_multi_set_option(self.h, name.encode(), int(value)), self.h)

    def counters(self) -> dict:
        c = capi.FtsbCounters()
        capi.check_multi(capi.lib().ftsb_multi_get_counters(self.h, ct.byref(c)), self.h)
        return {k: int(getattr(c, k)) for k in ("solve_points", "newton_iters", "lu_factor", "lu_solve", "rej_newton",
                                                "rej_plte", "redo", "sparse_flops")}

    def _host_params(self, vals: np.ndarray, fns: Optional[np.ndarray]):
        c = self.circ
        vals = np.ascontiguousarray(vals, np.float64)
        B = vals.shape[0]
        if fns is None:
            fns = np.broadcast_to(c.fns, (B, c.n_elems, 7))
        fns = np.asarray(fns)
        if fns.shape[1] == c.n_elems:
            fns = fns[:, self.fn_elems, :]
        return vals, np.ascontiguousarray(fns, np.float64), B

    def solve_op(self, vals: np.ndarray, fns: Optional[np.ndarray] = None) -> OpResult:
        vals, fns, B = self._host_params(vals, fns)
        x = np.zeros((B, self.circ.n_start))
        it = np.zeros(B, np.int32)
        st = np.zeros(B, np.int32)
        capi.check_multi(capi.lib().ftsb_multi_solve_op(self.h, B, _ptr(vals), _ptr(fns) if self.n_fn else None, _ptr(x), _ptr(it),
                                                        _ptr(st)), self.h)
        return OpResult(st, it, x)

    def solve_dc(self, vals: np.ndarray, fns: Optional[np.ndarray], source: Union[str, int], start, stop, step,
                 max_points: Optional[int] = None) -> DcResult:
        vals, fns, B = self._host_params(vals, fns)
        n = self.circ.n_run
        se = self.circ.elem_index(source) if isinstance(source, str) else int(source)
        start = np.ascontiguousarray(np.broadcast_to(np.asarray(start, np.float64), (B,)))
        stop = np.ascontiguousarray(np.broadcast_to(np.asarray(stop, np.float64), (B,)))
        step = np.ascontiguousarray(np.broadcast_to(np.asarray(step, np.float64), (B,)))
        counts = np.ceil((stop - start) / step)
        counts = np.where(counts >= 0, counts, 0).astype(np.int64)
        P = int(max_points or max(int(counts.max()), 1))
        x = np.zeros((B, P, n))
        it = np.zeros((B, P), np.int32)
        done = np.zeros(B, np.int64)
        st = np.zeros(B, np.int32)
        capi.check_multi(capi.lib().ftsb_multi_solve_dc(self.h, B, _ptr(vals), _ptr(fns) if self.n_fn else None, se, _ptr(start),
                                                        _ptr(stop), _ptr(step), P, _ptr(x), _ptr(it), _ptr(done), _ptr(st)), self.h)
        sweep = start[:, None] + step[:, None] * np.arange(P, dtype=np.float64)[None, :]
        return DcResult(st, done, it, x, sweep)

    def solve_tran(self, vals: np.ndarray, fns: Optional[np.ndarray], stop: float, step: float, capacity: int = 0,
                   rec_stride: int = 1) -> TranResult:
        vals, fns, B = self._host_params(vals, fns)
        n = self.circ.n_run
        cap = int(capacity)
        t = np.zeros((B, cap)) if cap else None
        x = np.zeros((B, cap, n)) if cap else None
        it = np.zeros((B, cap), np.int32) if cap else None
        nrec = np.zeros(B, np.int64)
        xf = np.zeros((B, n))
        st = np.zeros(B, np.int32)
        opts = capi.FtsbTranOpts(cap, rec_stride, 0)
        capi.check_multi(capi.lib().ftsb_multi_solve_tran(self.h, B, _ptr(vals), _ptr(fns) if self.n_fn else None, float(stop),
                                                          float(step), ct.byref(opts), _ptr(t), _ptr(x), _ptr(it), _ptr(nrec),
                                                          _ptr(xf), _ptr(st)), self.h)
        return TranResult(st, nrec, it, t, x, xf)


def _raise_for(status: int) -> None:
    if status in (capi.NOT_CONVERGED, capi.TRAN_H_UNDERFLOW):
        raise NotConvergedError(status)
    if status != capi.OK:
        raise EnginePanic(status)


class Engine:
    """Mirror of the reference's ``Engine`` for ONE netlist (a batch of one instance).

    ``Engine::new`` keeps the first command of each kind (src/engine.rs:32-43); ``run_*`` return a ``SimResult``
    with the reference's headers (``n_iters[,t]`` + unknown names in BTreeMap key order) or raise.
    """

    def __init__(self, elems: Sequence[nl.Elem], cmds: Sequence[nl.Command], device: int = 0):
        self.circ = nl.lower(elems, cmds)
        self.op_cmd = self.circ.command("op")
        self.dc_cmd = self.circ.command("dc")
        self.tran_cmd = self.circ.command("tran")
        self._be = BatchEngine(self.circ, device)
        self._be.set_params()

    @classmethod
    def from_netlist(cls, text: str, device: int = 0) -> "Engine":
        elems, cmds = nl.parse_netlist(text)
        nl.check_elems(elems)
        return cls(elems, cmds, device)

    def run_op(self) -> SimResult:
        r = self._be.run_op()
        _raise_for(int(r.status[0]))
        names = self.circ.headers(True)
        res = SimResult(["n_iters"] + names)
        rec = {"n_iters": float(r.n_iters[0])}
        for i in self.circ.order_start:
            rec[self.circ.names[i]] = float(r.x[0, i])
        res.push(rec)
        return res

    def run_dc(self) -> SimResult:
        if self.dc_cmd is None:
            raise RuntimeError("DC simulation wrongly configured.")
        d = self.dc_cmd
        try:
            se = self.circ.elem_index(d.source)
        except KeyError:
            raise RuntimeError("Sweep source not found") from None
        r = self._be.run_dc(se, d.start, d.stop, d.step)
        _raise_for(int(r.status[0]))  # `?` on the first failing point: the reference prints nothing
        names = self.circ.headers(False)
        res = SimResult(["n_iters"] + names)
        for k in range(int(r.points_done[0])):
            rec = {"n_iters": float(r.n_iters[0, k])}
            for i in self.circ.order_run:
                rec[self.circ.names[i]] = float(r.x[0, k, i])
            res.push(rec)
        return res

    def run_tran(self, capacity: int = 1 << 16) -> SimResult:
        if self.tran_cmd is None:
            raise RuntimeError("DC simulation wrongly configured.")  # sic, src/engine.rs:145
        c = self.tran_cmd
        r = self._be.run_tran(c.stop, c.step, capacity)
        _raise_for(int(r.status[0]))
        if int(r.n_rec[0]) > capacity:
            raise RuntimeError(f"{int(r.n_rec[0])} accepted steps exceed capacity {capacity}")
        names = self.circ.headers(False)
        res = SimResult(["n_iters", "t"] + names)
        for k in range(int(r.n_rec[0])):
            rec = {"n_iters": float(r.n_iters[0, k]), "t": float(r.t[0, k])}
            for i in self.circ.order_run:
                rec[self.circ.names[i]] = float(r.x[0, k, i])
            res.push(rec)
        return res


def lu_solve(a: np.ndarray, b: np.ndarray, device: int = 0) -> np.ndarray:
    """Batched dense solve on the GPU with the reference's pivoting rule (src/engine/gauss_lu.rs:3-54)."""
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    if a.ndim == 2:
        a, b = a[None], b[None]
    B, n, _ = a.shape
    x = np.zeros((B, n))
    capi.check(capi.lib().ftsb_lu_solve(device, n, B, _ptr(a), _ptr(b), _ptr(x), capi.MEM_HOST), None)
    return x
