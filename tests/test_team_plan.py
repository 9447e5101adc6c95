"""CPU test of the static elimination plan of the per-team kernels (ftspice_b200/csrc/team_plan.cpp): the plan's lists
(fill-in, targets, updates, substitution entries) are executed in numpy on random values and must give, bit for bit,
what a dense elimination with the same pivot sequence gives (gauss_lu.rs:29-53: reciprocal scaling, separate multiply
and subtract, column-oriented back substitution).  No GPU needed."""
import ctypes as ct

import numpy as np
import pytest

from ftspice_b200 import capi
from ftspice_b200 import netlist as nl

from test_q9_isource import mid_netlist

CIRCUITS = {
    "ladder64": lambda: nl.ladder_netlist(64, "SIN( 0.0 1.0 10M )", stop="40n", step="1n"),
    "ladder24_isrc": lambda: mid_netlist(24),
    "inverter40": lambda: nl.inverter_chain_netlist(40, stop="2n", step="50p"),
    "inverter258": lambda: nl.inverter_chain_netlist(256, stop="1n", step="50p"),
}


def team_plan(circ, an=2):
    lib = capi.lib()
    fn = lib.ftsb_debug_team_plan
    fn.argtypes = [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_void_p, ct.c_int, ct.c_void_p, ct.c_int, ct.c_void_p,
                   ct.c_void_p, ct.c_void_p]
    fn.restype = ct.c_int
    elems = (capi.FtsbElem * max(circ.n_elems, 1))()
    for i in range(circ.n_elems):
        elems[i].kind = int(circ.kind[i])
        for j in range(3):
            elems[i].node[j] = int(circ.node[i, j])
        elems[i].branch = int(circ.branch[i])
        elems[i].fn_kind = int(circ.fn_kind[i])
    cls = np.ascontiguousarray(circ.cls, np.int32)
    cap = 1 << 22
    ibuf = np.zeros(cap, np.int32)
    meta = np.zeros(32, np.int32)
    jdst = np.zeros(1 << 18, np.int32)
    piv = np.zeros(1 << 12, np.int32)
    rc = fn(elems, circ.n_elems, circ.n_run, circ.n_start, cls.ctypes.data, an, ibuf.ctypes.data, cap, meta.ctypes.data,
            jdst.ctypes.data, piv.ctypes.data)
    assert rc == 0, rc
    names = ["n", "nA", "S", "nJ", "nT", "nU", "nB"]
    p = {k: int(v) for k, v in zip(names, meta[:7])}
    offs = ["arow", "acol", "aptr", "acon", "sptr", "scon", "pivrow", "pive", "tptr", "tgt", "tgtr", "uptr", "updm", "updds",
            "bptr", "bwde", "bwdr"]
    lens = {"arow": p["n"] + 1, "acol": p["nA"], "aptr": p["nA"] + 1, "sptr": p["nJ"] + 1, "pivrow": p["n"], "pive": p["n"],
            "tptr": p["n"] + 1, "tgt": p["nT"], "tgtr": p["nT"], "uptr": p["nT"] + 1, "updm": p["nU"], "updds": p["nU"],
            "bptr": p["n"] + 1, "bwde": p["nB"], "bwdr": p["nB"]}
    for k, o in zip(offs, meta[7:24]):
        if k in lens:
            p[k] = ibuf[int(o):int(o) + lens[k]].copy()
    p["jdst"] = jdst[:p["nJ"]].copy()
    p["piv"] = piv[:p["n"]].copy()
    return p


def run_plan(p, vals, rhs):
    """The kernel's factor / forward / backward (solver_team.cuh) in numpy float64 scalars."""
    n, S = p["n"], p["S"]
    sv = np.zeros(S)
    sv[:p["nJ"]] = vals
    alpha = np.zeros(n)
    for k in range(n):
        al = np.float64(1.0) / sv[p["pive"][k]]
        alpha[k] = al
        t0, t1 = p["tptr"][k], p["tptr"][k + 1]
        for tt in range(t0, t1):
            e = p["tgt"][tt] & 0xffff
            sv[e] = al * sv[e]
        for u in range(p["uptr"][t0], p["uptr"][t1]):
            m = sv[p["updm"][u]]
            if m != 0.0:
                d, s = p["updds"][u] & 0xffff, (p["updds"][u] >> 16) & 0xffff
                sv[d] = sv[d] - m * sv[s]
    y = rhs.copy()
    for k in range(n):
        for tt in range(p["tptr"][k], p["tptr"][k + 1]):
            m = sv[p["tgt"][tt] & 0xffff]
            if m != 0.0:
                r = p["tgtr"][tt]
                y[r] = y[r] - m * y[p["pivrow"][k]]
    x = np.zeros(n)
    for i in range(n - 1, -1, -1):
        x[i] = y[p["pivrow"][i]] / sv[p["pive"][i]]
        for bb in range(p["bptr"][i], p["bptr"][i + 1]):
            r = p["bwdr"][bb]
            y[r] = y[r] - x[i] * sv[p["bwde"][bb]]
    return x


def dense_same_pivots(n, a, rhs, piv):
    """gauss_lu.rs:3-54 with the pivot row of every step prescribed (rows really swapped, like the reference)."""
    a = a.copy()
    b = rhs.copy()
    where = list(range(n))  # where[r] = current position of original row r
    at = list(range(n))
    for k in range(n):
        pp = where[piv[k]]
        if pp != k:
            a[[k, pp]] = a[[pp, k]]
            b[[k, pp]] = b[[pp, k]]
            rk = at[k]
            at[k], at[pp] = piv[k], rk
            where[piv[k]], where[rk] = k, pp
        alpha = np.float64(1.0) / a[k, k]
        for i in range(k + 1, n):
            a[i, k] = a[i, k] * alpha
        for i in range(k + 1, n):
            if a[i, k] != 0.0:
                a[i, k + 1:] = a[i, k + 1:] - a[i, k] * a[k, k + 1:]
                b[i] = b[i] - a[i, k] * b[k]
    x = b.copy()
    for i in range(n - 1, -1, -1):
        x[i] = x[i] / a[i, i]
        x[:i] = x[:i] - x[i] * a[:i, i]
    return x


@pytest.mark.parametrize("which", sorted(CIRCUITS))
def test_plan_equals_dense_elimination_bitwise(which):
    circ = nl.load(CIRCUITS[which]())
    p = team_plan(circ)
    n = p["n"]
    assert n == circ.n_run and sorted(p["piv"].tolist()) == list(range(n))
    rng = np.random.default_rng(7)
    rows, cols = p["jdst"] & 0xffff, p["jdst"] >> 16
    for trial in range(3):
        vals = rng.uniform(0.5, 2.0, p["nJ"]) * rng.choice([-1.0, 1.0], p["nJ"])
        if trial == 2:
            vals[rng.integers(0, p["nJ"], 5)] = 0.0  # exact zeros among the off-pivot entries: the zero-multiplier skip
            vals[p["pive"][p["pive"] < p["nJ"]]] = 3.0
        a = np.zeros((n, n))
        a[rows, cols] = vals
        rhs = rng.standard_normal(n)
        x_plan = run_plan(p, vals, rhs)
        x_dense = dense_same_pivots(n, a, rhs, p["piv"].tolist())
        assert np.all(np.isfinite(x_dense))
        assert np.array_equal(x_plan, x_dense), (which, trial, np.abs(x_plan - x_dense).max())
        assert np.abs(a @ x_plan - rhs).max() < 1e-6 * max(1.0, np.abs(x_plan).max())


def test_plan_a_entries_are_the_linear_part():
    """A = linear + companion stamps: for the ladder every entry of J is an entry of A (no nonlinear device)."""
    circ = nl.load(CIRCUITS["ladder64"]())
    p = team_plan(circ)
    assert p["nA"] == p["nJ"] and p["S"] >= p["nJ"]
    a_keys = set()
    for r in range(p["n"]):
        for c in range(p["arow"][r], p["arow"][r + 1]):
            a_keys.add((r, int(p["acol"][c])))
    j_keys = {(int(d) & 0xffff, int(d) >> 16) for d in p["jdst"]}
    assert a_keys == j_keys
    circ = nl.load(CIRCUITS["inverter40"]())
    p = team_plan(circ)
    assert p["nA"] < p["nJ"]  # the NMOS gate columns exist in J only
