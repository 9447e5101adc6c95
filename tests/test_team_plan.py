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

CHUNK = 16  # TEAM_CHUNK of team_plan.h

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
    meta = np.zeros(96, np.int32)
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
    for k, v in zip(["ellW", "nSlotA", "chain", "nFop", "nBop", "nFchunk", "nBchunk"], meta[25:32]):
        p[k] = int(v)
    lens2 = {"fop": p["nFop"], "fvidx": p["nFchunk"] * CHUNK, "fchunk": p["nFchunk"] + 1, "bop": 2 * p["nBop"],
             "bvidx": p["nBchunk"] * CHUNK, "bchunk": p["nBchunk"] + 1, "ellcol": 2 * (p["n"] + 1), "ovfptr": p["n"] + 1}
    for k, o in zip(["fop", "fvidx", "fchunk", "bop", "bvidx", "bchunk", "ellcol", "ovfptr", "ovfcol"], meta[32:41]):
        if k in lens2:
            p[k] = ibuf[int(o):int(o) + lens2[k]].copy()
        else:
            p[k] = ibuf[int(o):int(o) + int(p["ovfptr"][-1])].copy()
    for k, v in zip(["nEarly", "nXop", "nDFop", "nDBop"], meta[41:45]):
        p[k] = int(v)
    for k, o, ln in zip(["early", "xop", "dfop", "dbop"], meta[45:49], [p["nEarly"], 2 * p["nXop"], 2 * p["nDFop"], 2 * p["nDBop"]]):
        p[k] = ibuf[int(o):int(o) + ln].copy()
    for k, v in zip(["nXlev", "nFlev", "nBlev"], meta[49:52]):
        p[k] = int(v)
    for k, o, ln in zip(["xlptr", "xlop", "flptr", "flop", "blptr", "blop"], meta[52:58],
                        [p["nXlev"] + 1, 2 * (p["n"] + p["nU"] + p["nDFop"]), p["nFlev"] + 1, 0, p["nBlev"] + 1, 2 * p["nDBop"]]):
        p[k] = ibuf[int(o):int(o) + ln].copy()
    for key, (ns, o_seg, o_ops) in {"x": meta[59:62], "b": meta[62:65]}.items():
        seg = ibuf[int(o_seg):int(o_seg) + 2 * int(ns)].reshape(-1, 2)
        p[key + "seg"] = seg.copy()
        n_ops = max([int(off) + (int(c) * NARROW if c > 0 else -int(c)) for c, off in seg] + [0])
        p[key + "sop"] = ibuf[int(o_ops):int(o_ops) + 2 * n_ops].copy()
    p["spre"] = ibuf[int(meta[58]):int(meta[58]) + p["nJ"]].copy()
    p["scon"] = ibuf[int(meta[12]):int(meta[12]) + int(p["sptr"][-1])].copy()
    p["aptr"] = ibuf[int(meta[9]):int(meta[9]) + p["nSlotA"] + 1].copy()  # indexed by ELL slot (padding slots: empty lists)
    p["acon"] = ibuf[int(meta[10]):int(meta[10]) + int(p["aptr"][-1])].copy()
    p["jdst"] = jdst[:p["nJ"]].copy()
    p["piv"] = piv[:p["n"]].copy()
    return p


def run_streams(p, sv, alpha, rhs):
    """The kernel's chain_forward / chain_backward: op streams, values staged in chunks of CHUNK slots."""
    n, S = p["n"], p["S"]
    sva = np.concatenate([sv, alpha])
    y = rhs.copy()
    for c in range(p["nFchunk"]):
        buf = sva[p["fvidx"][CHUNK * c:CHUNK * c + CHUNK]]
        slot = 0
        for op in range(p["fchunk"][c], p["fchunk"][c + 1]):
            w0 = int(p["fop"][op])
            r, pk = w0 & 0xffff, (w0 >> 16) & 0xffff
            m = buf[slot]
            slot += 1
            if m != 0.0:
                y[r] = y[r] - m * y[pk]
    x = np.full(n, np.nan)
    xi = 0.0
    for c in range(p["nBchunk"]):
        buf = sva[p["bvidx"][CHUNK * c:CHUNK * c + CHUNK]]
        slot = 0
        for op in range(p["bchunk"][c], p["bchunk"][c + 1]):
            w0, w1 = int(p["bop"][2 * op]), int(p["bop"][2 * op + 1])
            if w1 < 0:
                pi, i = w0 & 0xffff, (w0 >> 16) & 0xffff
                d, al = buf[slot], buf[slot + 1]
                slot += 2
                assert al == alpha[i]
                xi = y[pi] / d
                x[i] = xi
            else:
                y[w0] = y[w0] - xi * buf[slot]
                slot += 1
        assert slot <= CHUNK
    assert p["fchunk"][-1] == p["nFop"] and p["bchunk"][-1] == p["nBop"]
    return x


def run_direct(p, vals, rhs):
    """factor_direct / forward_direct / backward_direct of solver_team.cuh: early steps first (any order), then the op stream."""
    n, S = p["n"], p["S"]
    sv = np.zeros(S)
    sv[:p["nJ"]] = vals
    alpha = np.zeros(n)

    def head(k):
        al = np.float64(1.0) / sv[p["pive"][k]]
        alpha[k] = al
        for tt in range(p["tptr"][k], p["tptr"][k + 1]):
            e = p["tgt"][tt] & 0xffff
            sv[e] = al * sv[e]

    for k in reversed(p["early"].tolist()):  # the kernel does them lanes-in-parallel: order must not matter
        head(k)
    for op in range(p["nXop"]):
        w0, w1 = int(p["xop"][2 * op]), int(p["xop"][2 * op + 1])
        if w1 < 0:
            head(w0)
        else:
            m = sv[w1]
            if m != 0.0:
                d, s = w0 & 0xffff, (w0 >> 16) & 0xffff
                sv[d] = sv[d] - m * sv[s]
    y = rhs.copy()
    for op in range(p["nDFop"]):
        w0, w1 = int(p["dfop"][2 * op]), int(p["dfop"][2 * op + 1])
        m = sv[w1]
        if m != 0.0:
            r, pk = w0 & 0xffff, (w0 >> 16) & 0xffff
            y[r] = y[r] - m * y[pk]
    x = np.full(n, np.nan)
    xi = 0.0
    for op in range(p["nDBop"]):
        w0, w1 = int(p["dbop"][2 * op]), int(p["dbop"][2 * op + 1])
        if w1 < 0:
            pi, i = w0 & 0xffff, (w0 >> 16) & 0xffff
            xi = y[pi] / sv[w1 & 0xffff]
            x[i] = xi
        else:
            y[w0] = y[w0] - xi * sv[w1]
    return x


NARROW, OP_NOP = 4, 0x20000000


def levels_of_segments(seg, sop):
    """The segmented form the kernel walks (seg_walk): runs of narrow levels padded to NARROW slots, wide levels plain."""
    levels = []
    for cnt, off in seg:
        cnt, off = int(cnt), int(off)
        if cnt < 0:
            levels.append([(int(sop[2 * o]), int(sop[2 * o + 1])) for o in range(off, off - cnt)])
        else:
            for j in range(cnt):
                ops = [(int(sop[2 * o]), int(sop[2 * o + 1])) for o in range(off + j * NARROW, off + (j + 1) * NARROW)]
                levels.append([w for w in ops if not (w[1] >= 0 and (w[1] & OP_NOP))])
    return levels


def test_segmented_schedules_are_the_level_schedules():
    for which in ("inverter258", "inverter40", "ladder64"):
        p = team_plan(nl.load(CIRCUITS[which]()))
        for key, lp, lo in (("x", "xlptr", "xlop"), ("b", "blptr", "blop")):
            want = [[(int(p[lo][2 * o]), int(p[lo][2 * o + 1])) for o in range(p[lp][L], p[lp][L + 1])] for L in range(len(p[lp]) - 1)]
            assert levels_of_segments(p[key + "seg"], p[key + "sop"]) == want, (which, key)
            assert all(len(lv) > 0 for lv in want)


def run_levels(p, vals, rhs, rng):
    """factor_levels / forward_levels / backward_levels of solver_team.cuh: the ops of a level run on different lanes at the
    same time — here in a RANDOM order inside the level, reading the state as it stood when the level began and writing at
    its end (what lanes between two warp barriers may observe at worst): the result must not depend on it."""
    n, S = p["n"], p["S"]
    sv = np.zeros(S)
    sv[:p["nJ"]] = vals
    alpha = np.zeros(n)
    y = rhs.copy()
    x = np.full(n, np.nan)

    def walk(lptr, ops, nops, exec_op, arrays):
        assert lptr[0] == 0 and lptr[-1] == nops and np.all(np.diff(lptr) > 0), (lptr[-1], nops)
        for L in range(len(lptr) - 1):
            snap = [a.copy() for a in arrays]  # operands as they stood at the barrier
            writes = {}
            for o in rng.permutation(np.arange(lptr[L], lptr[L + 1])):
                for key, v in exec_op(int(ops[2 * o]), int(ops[2 * o + 1]), *snap):
                    assert key not in writes, "two ops of a level write the same location"
                    writes[key] = v
            for (ai, idx), v in writes.items():
                arrays[ai][idx] = v

    def x_op(w0, w1, sv_, alpha_, y_):
        out = []
        if w1 >= 0 and (w1 & 0x40000000):  # forward substitution op riding along with the elimination
            m = sv_[w1 & 0xffff]
            if m != 0.0:
                r, pk = w0 & 0xffff, (w0 >> 16) & 0xffff
                out.append(((2, r), y_[r] - m * y_[pk]))
        elif w1 < 0:
            k = w0
            al = np.float64(1.0) / sv_[p["pive"][k]]
            out.append(((1, k), al))
            for tt in range(p["tptr"][k], p["tptr"][k + 1]):
                e = p["tgt"][tt] & 0xffff
                out.append(((0, e), al * sv_[e]))
        else:
            m = sv_[w1]
            if m != 0.0:
                d, s_ = w0 & 0xffff, (w0 >> 16) & 0xffff
                out.append(((0, d), sv_[d] - m * sv_[s_]))
        return out

    walk(p["xlptr"], p["xlop"], p["n"] + p["nU"] + p["nDFop"], x_op, [sv, alpha, y])  # every head, update and forward op

    def f_op(w0, w1, y_):
        m = sv[w1]
        if m != 0.0:
            r, pk = w0 & 0xffff, (w0 >> 16) & 0xffff
            return [((0, r), y_[r] - m * y_[pk])]
        return []

    assert p["nFlev"] == 0  # (a separate forward schedule: unused)

    def b_op(w0, w1, y_, x_):
        row, i = w0 & 0xffff, (w0 >> 16) & 0xffff
        if w1 < 0:
            return [((1, i), y_[row] / sv[w1 & 0xffff])]
        return [((0, row), y_[row] - x_[i] * sv[w1])]

    walk(p["blptr"], p["blop"], p["nDBop"], b_op, [y, x])
    return x


def run_plan(p, vals, rhs, streams=False):
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
    if streams:
        return run_streams(p, sv, alpha, rhs)
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
        assert np.array_equal(run_plan(p, vals, rhs, streams=True), x_dense), (which, trial, "op streams")
        assert np.array_equal(run_direct(p, vals, rhs), x_dense), (which, trial, "direct streams")
        assert np.array_equal(run_levels(p, vals, rhs, rng), x_dense), (which, trial, "level schedules")
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
    # the residual's ELL form: a tridiagonal ladder needs three slots per row and no overflow; the columns of a row are
    # its CSR columns in order, padding slots point at the row itself; row n is the zero row
    assert p["ellW"] == 3 and p["nSlotA"] == 3 * (p["n"] + 1) and p["ovfptr"][-1] == 0 and p["chain"] == 1
    n = p["n"]
    assert [int(p["ellcol"][2 * n]) & 0xffff, int(p["ellcol"][2 * n]) >> 16, int(p["ellcol"][2 * n + 1]) & 0xffff] == [n, n, n]
    for r in range(p["n"]):
        cols = [int(p["ellcol"][2 * r]) & 0xffff, (int(p["ellcol"][2 * r]) >> 16) & 0xffff, int(p["ellcol"][2 * r + 1]) & 0xffff]
        want = [int(c) for c in p["acol"][p["arow"][r]:p["arow"][r + 1]]]
        assert cols[:len(want)] == want and all(c == r for c in cols[len(want):])
    circ = nl.load(CIRCUITS["inverter40"]())
    p = team_plan(circ)
    assert p["nA"] < p["nJ"]  # the NMOS gate columns exist in J only
    assert p["chain"] == 0 and p["ovfptr"][-1] > 0 and p["ellW"] == 4  # the supply rail: wide elimination steps, a long row
    assert p["nEarly"] >= 36  # every stage of the chain is an early step (nothing writes its pivot or its column entries)
    circ = nl.load(CIRCUITS["ladder64"]())
    assert team_plan(circ)["nEarly"] <= 3  # a ladder: every step updates the next pivot


@pytest.mark.gpu
def test_team_exact_division_equals_plain_division_bitwise():
    """The back substitution divides by a pivot whose correctly rounded reciprocal is known (solver_team.cuh,
    div_known_rcp): the result must be the IEEE quotient bit for bit — random operands over the guarded exponent range
    and beyond it (zeros, denormals, huge, infinities, NaN take the plain division)."""
    lib = capi.lib()
    fn = lib.ftsb_debug_div
    fn.argtypes = [ct.c_int, ct.c_longlong, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p]
    fn.restype = ct.c_int
    rng = np.random.default_rng(11)
    n = 1 << 23

    def rnd(emin, emax, special):
        man = rng.integers(0, 1 << 52, n, dtype=np.uint64)
        if special:
            k = rng.integers(0, 4, n)
            man = np.where(k == 0, np.uint64((1 << 52) - 1) - rng.integers(0, 16, n, dtype=np.uint64), man)
            man = np.where(k == 1, rng.integers(0, 16, n, dtype=np.uint64), man)
            man = np.where(k == 2, rng.integers(0, 256, n, dtype=np.uint64) << np.uint64(44), man)
        e = rng.integers(emin + 1023, emax + 1024, n, dtype=np.uint64)
        s = rng.integers(0, 2, n, dtype=np.uint64)
        return ((s << np.uint64(63)) | (e << np.uint64(52)) | man).view(np.float64)

    for (e0, e1, special) in ((-60, 60, True), (-399, 399, False), (-1022, 1023, False)):
        y, d = rnd(e0, e1, special), rnd(e0, e1, special)
        if e0 < -1000:  # sprinkle the operands the guard must route to the plain division
            y[:64] = np.resize([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, 1e-310, 1.7e308], 64)
            d[64:128] = np.resize([5e-324, 1e-310, 1.7e308, np.inf], 64)
        fast, ref = np.zeros(n), np.zeros(n)
        assert fn(0, n, y.ctypes.data, d.ctypes.data, fast.ctypes.data, ref.ctypes.data) == 0
        same = (fast.view(np.uint64) == ref.view(np.uint64)) | (np.isnan(fast) & np.isnan(ref))
        assert same.all(), (e0, e1, int((~same).sum()), y[~same][:3], d[~same][:3])
        with np.errstate(all="ignore"):
            host = y / d
        assert ((ref.view(np.uint64) == host.view(np.uint64)) | (np.isnan(ref) & np.isnan(host))).all()


@pytest.mark.parametrize("which", ["inverter40", "inverter258", "ladder64"])
def test_linear_prefix_of_the_factor_entries(which):
    """spre[e] = how many leading contributions of static entry e are linear / companion stamps: exactly the contributions
    of A's entry at the same position, in the same order (the per-attempt partial sum the Newton iterations start from)."""
    p = team_plan(nl.load(CIRCUITS[which]()))
    n, W = p["n"], p["ellW"]
    a_list = {}
    ovf = 0
    for r in range(n):
        cols = [int(p["acol"][c]) for c in range(p["arow"][r], p["arow"][r + 1])]
        for w, c in enumerate(cols):
            if w < W:
                slot = r * W + w
            else:
                slot = W * (n + 1) + ovf
                ovf += 1
            a_list[(r, c)] = p["acon"][p["aptr"][slot]:p["aptr"][slot + 1]].tolist()
    longest = seen = 0
    for d in range(p["nJ"]):
        r, c = int(p["jdst"][d]) & 0xffff, int(p["jdst"][d]) >> 16
        full = p["scon"][p["sptr"][d]:p["sptr"][d + 1]].tolist()
        k, idx = int(p["spre"][d]) >> 16, int(p["spre"][d]) & 0xffff
        al = a_list.get((r, c), [])
        assert full[:k] == al[:k]
        if k:  # a kept sum: at least two contributions, the whole common prefix, compact numbering
            assert k >= 2 and (k == len(full) or k == len(al) or full[k] != al[k])
            assert idx == seen
            seen += 1
        assert [w for w in full if w in al] == al  # A's list is J's list without the nonlinear stamps, in order
        longest = max(longest, k)
    if which == "inverter258":
        assert longest >= 254  # the rail's diagonal entry: one conductance per stage
    if which == "ladder64":
        assert all((int(p["spre"][d]) >> 16) in (0, p["sptr"][d + 1] - p["sptr"][d]) for d in range(p["nJ"]))  # linear circuit
