"""Host-side logic without a GPU: netlist reader / lowering (mirrors src/parser.rs, src/node_collection.rs),
result formatting (src/engine/sim_result.rs), the C-ABI library's exports, Monte-Carlo sharding."""
import ctypes as ct
import os
import re

import numpy as np
import pytest

from ftspice_b200 import capi
from ftspice_b200 import netlist as nl
from ftspice_b200.engine import SimResult, rust_f64_str

from harness import ROOT, golden, load, netlist_names


# ---- values: src/spice.pest:44-46, src/parser.rs:310-340 ----------------------------------------------
@pytest.mark.parametrize("tok,val", [("2200", 2200.0), ("2.2k", 2.2 * 1e3), ("10p", 10 * 1e-12), ("0.3u", 0.3 * 1e-6),
                                      ("100M", 100 * 1e6), ("-1.5", -1.5), ("100u", 100 * 1e-6), ("2.", 2.0),
                                      ("1da", 10.0), ("1d", 0.1), ("3G", 3e9), ("5f", 5 * 1e-15), ("7h", 700.0),
                                      ("4c", 4 * 1e-2)])
def test_parse_value(tok, val):
    assert nl.parse_value(tok) == val


def test_parse_value_units_and_errors():
    assert nl.parse_value("4000mV", "V") == 4000 * 1e-3
    assert nl.parse_value("10mA", "A") == 10 * 1e-3
    assert nl.parse_value("3V", "V") == 3.0
    for bad in ("1e3", "+1", "01", "abc", "1kk"):
        with pytest.raises(nl.NetlistError):
            nl.parse_value(bad)


def test_all_reference_netlists_parse():
    """src/parser.rs:346-426 load test/*.sp and check element names and command kinds."""
    g = golden()
    assert len(netlist_names()) == 17
    c = load("proj1")
    assert [e.name for e in c.elems] == ["V50", "V76", "V32", "I1", "I2", "R1", "R2", "R3", "R4", "R5", "R6", "R7", "R8"]
    assert [k.kind for k in c.cmds] == ["op"]
    c = load("proj2")
    assert [k.kind for k in c.cmds] == ["op", "dc"] and c.command("dc").source == "V01"
    assert c.command("dc").step == 10 * 1e-3
    c = load("proj3")  # "*.OP" is a comment
    assert [k.kind for k in c.cmds] == ["tran"] and c.command("tran").stop == 40 * 1e-9
    assert c.elems[0].fn_kind == nl.FN_EXP and c.elems[0].fn == (3.0, 0.0, 0.0, 2 * 1e-9, 40 * 1e-9, 2 * 1e-9)
    assert c.elems[0].val == 3.0  # f(0), parser.rs:95-99
    for n in netlist_names():
        assert g[n]["n_run"] == load(n).n_run and g[n]["n_start"] == load(n).n_start


def test_terminal_order_swaps():
    """parser.rs:83-84,114-115,145-146,161-162,177-178: V, I, C, L, D store [second, first]; R, Q, M keep order."""
    txt = "V1 a 0 1V\nI1 a b 1mA\nR1 a b R=1\nC1 a b C=1p\nL1 a b L=1n\nD1 a b d_model\nQ1 a b c 0 q_model\nM1 a b c 0 t_model\n.end\n"
    el, _ = nl.parse_netlist(txt)
    assert el[0].nodes == ["0", "a"] and el[1].nodes == ["b", "a"]
    assert el[2].nodes == ["a", "b"]
    assert el[3].nodes == ["b", "a"] and el[4].nodes == ["b", "a"] and el[5].nodes == ["b", "a"]
    assert el[6].nodes == ["a", "b", "c"] and el[7].nodes == ["a", "b", "c"]


def test_node_collection_ordering():
    """node_collection.rs:13-74: voltage nodes sorted bytewise ("10" < "2", digits < upper < lower), then V
    names, then (start-up) inductor names; headers in BTreeMap key order."""
    txt = "V9 10 0 1V\nVa 2 0 1V\nR1 10 b R=1\nR2 b B R=1\nR3 B 2 R=1\nL1 2 x L=1n\nLa x 0 L=1n\n.op\n.end\n"
    c = nl.load(txt)
    assert c.names[:5] == ["10", "2", "B", "b", "x"]
    assert c.names[5:7] == ["V9", "Va"] and c.names[7:] == ["L1", "La"]
    assert c.n_run == 7 and c.n_start == 9
    assert list(c.cls) == [0] * 5 + [1] * 4
    assert c.headers(True) == ["10", "2", "B", "L1", "La", "V9", "Va", "b", "x"]
    assert c.headers(False) == ["10", "2", "B", "V9", "Va", "b", "x"]
    li = c.elem_index("L1")
    assert c.branch[li] == 7 and list(c.node[li][:2]) == [c.index_of["x"], c.index_of["2"]]


def test_check_elems():
    """src/parser/check_elems.rs:29-76."""
    with pytest.raises(nl.NetlistError, match="Duplicate"):
        nl.load("R1 1 0 R=1\nR1 1 0 R=2\n.end\n")
    with pytest.raises(nl.NetlistError, match="GND"):
        nl.load("R1 1 2 R=1\n.end\n")


def test_rust_f64_display():
    """src/engine/sim_result.rs:24-36 prints with Rust's f64 Display."""
    assert rust_f64_str(26.0) == "26"
    assert rust_f64_str(1e-18) == "0.000000000000000001"
    assert rust_f64_str(3.999999999643776) == "3.999999999643776"
    assert rust_f64_str(-0.0009090909090909089) == "-0.0009090909090909089"
    assert rust_f64_str(1e22) == "10000000000000000000000"
    assert rust_f64_str(0.0) == "0" and rust_f64_str(-0.0) == "-0"
    assert rust_f64_str(float("nan")) == "NaN" and rust_f64_str(float("inf")) == "inf"
    r = SimResult(["n_iters", "1"])
    r.push({"1": 0.5, "n_iters": 3.0})
    assert r.to_csv() == "n_iters,1\n3,0.5\n" and list(r.get("1")) == [0.5]


# ---- the C-ABI library: loads, exports everything include/ftsb.h declares, fails loudly without a GPU ----
def declared_functions():
    hdr = open(os.path.join(ROOT, "include", "ftsb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ftsb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"libftsb.so does not export {n}"
    assert set(names) == set(capi.EXPORTS)
    assert L.ftsb_abi_version() == 1


def test_abi_struct_layout():
    assert ct.sizeof(capi.FtsbElem) == 24
    assert ct.sizeof(capi.FtsbTranOpts) == 16
    assert ct.sizeof(capi.FtsbCounters) == 64


def test_compile_validates_arguments_or_reports_no_device():
    """Without a CUDA device every compute entry point fails loudly (no CPU fallback); argument errors are
    reported before any device work."""
    L = capi.lib()
    h = ct.c_void_p()
    rc = L.ftsb_compile(None, 0, 0, 0, None, 0, None)
    assert rc == capi.E_ARG
    e = (capi.FtsbElem * 1)()
    e[0].kind = 0
    e[0].node[0], e[0].node[1], e[0].node[2] = 0, -1, -1
    e[0].branch = -1
    cls = (ct.c_int32 * 1)(0)
    rc = L.ftsb_compile(e, 1, 1, 1, cls, 0, ct.byref(h))
    if L.ftsb_device_count() == 0:
        assert rc == capi.E_CUDA
        assert b"no CPU fallback" in L.ftsb_last_error(None)
        assert L.ftsb_lu_solve(0, 2, 1, None, None, None, 0) != 0
        assert L.ftsb_measure_fp64_tflops(0) < 0
    else:
        assert rc == 0
        L.ftsb_destroy(h)


# ---- Monte-Carlo parameters: instance 0 nominal, +-10 %, shard-independent ------------------------------
def test_monte_carlo_properties():
    c = nl.load(nl.PROJ1_NL)
    vals, fns = nl.monte_carlo(c, 64)
    assert np.array_equal(vals[0], c.vals)
    scal = np.isin(c.kind, (nl.R, nl.C, nl.L, nl.V, nl.I))
    ratio = vals[1:, scal] / c.vals[scal]
    assert ratio.min() >= 0.9 and ratio.max() <= 1.1 and ratio.std() > 0.03
    assert np.array_equal(vals[:, ~scal], np.broadcast_to(c.vals[~scal], (64, (~scal).sum())))
    a, _ = nl.monte_carlo(c, 16, first=16)
    assert np.array_equal(a, vals[16:32])  # shards of one batch draw the same numbers


def test_synthetic_configs_shapes():
    c4 = nl.load(nl.ladder_netlist(64))
    assert c4.n_run == 65 and c4.n_start == 80 and c4.command("tran").stop == 5.4 * 1e-6
    c5 = nl.load(nl.inverter_chain_netlist(256))
    assert c5.n_run == 258 and sum(1 for e in c5.elems if e.kind == nl.M) == 254
    c2 = nl.load(nl.PROJ1_NL)
    assert c2.n_start == 14 and sum(1 for e in c2.elems if e.kind in (nl.D, nl.Q)) == 2


@pytest.mark.parametrize("count", [1, 2, 3, 7, 33, 64, 65, 129, 1000])
@pytest.mark.parametrize("tol", [1e-6, 1e-9])
def test_norm_convergence_test_is_a_threshold_on_the_sum_of_squares(count, tol):
    """solver_team.cuh (compile-time-rows kernels) replaces `s < 0.001 * s + tol`, s = sqrt(q) / count, by `q < threshold`
    with the threshold found by bisection on the bit pattern (TeamOps::norm_threshold).  That is exact iff the test is
    monotone in q: checked here with IEEE double arithmetic (numpy: correctly rounded sqrt / divide, no contraction) on
    40 000 neighbouring bit patterns around the threshold and on 200 000 random magnitudes."""
    def ok(q):
        s = np.sqrt(q) / np.float64(count)
        return s < np.float64(0.001) * s + np.float64(tol)

    lo, hi = 0, 0x7FF0000000000000
    assert ok(np.float64(0.0)) and not ok(np.float64(np.inf))
    while hi - lo > 1:
        mid = lo + (hi - lo) // 2
        if ok(np.array([mid], dtype=np.uint64).view(np.float64)[0]):
            lo = mid
        else:
            hi = mid
    thr = np.array([hi], dtype=np.uint64).view(np.float64)[0]
    near = (np.arange(-20000, 20000, dtype=np.int64) + hi).astype(np.uint64).view(np.float64)
    assert np.array_equal(ok(near), near < thr)
    rng = np.random.default_rng(count)
    q = np.exp(rng.uniform(np.log(thr) - 40.0, np.log(thr) + 40.0, 200000))
    q = np.concatenate([q, thr * (1.0 + rng.uniform(-1e-9, 1e-9, 100000)), [0.0, np.inf]])
    assert np.array_equal(ok(q), q < thr)
    with np.errstate(invalid="ignore"):
        assert not ok(np.float64(np.nan)) and not (np.float64(np.nan) < thr)
