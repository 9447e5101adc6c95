"""The reference's own known-answer tests for the hot path, run against the CPU ORACLE (no GPU).

Every test cites the reference unit test it restates (paths relative to /root/reference/).  These pin the
oracle's building blocks; end-to-end .op/.dc/.tran results are pinned by nothing in the reference
 — see tests/test_oracle_golden.py for the cross-checks against SURVEY App. A."""
import ctypes as ct

import numpy as np
import pytest

from oracle import orc

D = ct.POINTER(ct.c_double)


def dp(a):
    return a.ctypes.data_as(D)


@pytest.fixture(scope="module")
def L():
    orc.build()
    return orc.lib()


def lu(L, a, b):
    a = np.array(a, np.float64)
    b = np.array(b, np.float64)
    x = np.zeros(len(b))
    L.orc_lu_solve(len(b), dp(a), dp(b), dp(x))
    return x


# ---- src/engine/gauss_lu.rs:60-109 -------------------------------------------------------------------
def test_lu_2x2_zero(L):
    assert (lu(L, [[1.0, 0.0], [0.0, 1.0]], [0.0, 0.0]) == [0.0, 0.0]).all()


def test_lu_2x2_nonzero(L):
    assert (lu(L, [[1.0, 0.0], [0.0, 1.0]], [1.0, 2.0]) == [1.0, 2.0]).all()


def test_lu_2x2_nontrivial(L):
    x = lu(L, [[5.0, 2.0], [-1.0, 3.0]], [1.0, 2.0])
    assert abs(x[0] - (-1.0 / 17.0)) < 1e-15 and abs(x[1] - 11.0 / 17.0) < 1e-15


def test_lu_3x3_nontrivial(L):
    x = lu(L, [[5.0, 2.0, 1.0], [-1.0, 3.0, -1.0], [0.0, 2.0, -1.0]], [1.0, 2.0, 1.0])
    assert abs(x[0] - (-2.0 / 9.0)) < 1e-15 and abs(x[1] - 7.0 / 9.0) < 1e-15 and abs(x[2] - 5.0 / 9.0) < 1e-15


def test_lu_pivot_rule_first_strict_max(L):
    """gauss_lu.rs:6-16: ties keep the FIRST row; verified through the multiplier pattern left in `a`."""
    a = np.array([[1.0, 2.0, 3.0], [-1.0, 1.0, 0.0], [1.0, 0.0, 2.0]])
    b = np.array([1.0, 2.0, 3.0])
    x = np.zeros(3)
    aa = a.copy()
    L.orc_lu_solve(3, dp(aa), dp(b.copy()), dp(x))
    assert (aa[0] == a[0]).all()  # |1| == |-1| == |1|: row 0 stays the pivot
    np.testing.assert_allclose(a @ x, b, rtol=0, atol=1e-14)


# ---- src/engine/mna.rs:36-52 --------------------------------------------------------------------------
def test_get_err(L):
    x = np.array([1.0, 2.0])
    a = np.array([[1.0, 2.0], [3.0, 4.0]])
    b = np.array([8.0, 12.5])
    h = np.eye(2)
    g = np.array([x.sum(), x.sum() / 2])
    out = np.zeros(2)
    L.orc_get_err_raw(2, 2, dp(a), dp(b), dp(h), dp(g), dp(x), dp(out))
    assert (out == [0.0, 0.0]).all()


# ---- src/engine/node_vec_norm.rs:49-77 ----------------------------------------------------------------
def test_norm_zero(L):
    assert L.orc_norm(dp(np.zeros(3)), 3) == 0.0


def test_norm_pos(L):
    assert L.orc_norm(dp(np.ones(3)), 3) > 0.0


def test_norm_scale(L):
    v = np.ones(3)
    assert abs(3.0 * L.orc_norm(dp(v), 3) - L.orc_norm(dp(v * 3.0), 3)) < 1e-12


def test_norm_triangular(L):
    v1, v2 = np.ones(3), 2.0 * np.ones(3)
    assert L.orc_norm(dp(v1 + v2), 3) >= L.orc_norm(dp(v1), 3) + L.orc_norm(dp(v2), 3)


def test_norm_is_sqrt_sum_over_len_and_nan_when_empty(L):
    """node_vec_norm.rs:12-14: sqrt(sum x^2)/len (not RMS); an empty class gives 0/0 = NaN (quirk Q3)."""
    assert L.orc_norm(dp(np.array([3.0, 4.0])), 2) == 2.5
    assert np.isnan(L.orc_norm(dp(np.zeros(1)), 0))


# ---- src/engine/newtons_method.rs:95-153 --------------------------------------------------------------
def test_dampen_step(L):
    step = np.array([1.0, 2.0, 3.0])
    out = np.zeros(3)
    L.orc_dampen_step(dp(step), dp(out), 3)
    assert (step > out).all()
    small = np.array([1e-9, -1e-9, 0.0])
    L.orc_dampen_step(dp(small), dp(out), 3)
    np.testing.assert_allclose(out[:2], 1.3 * small[:2], rtol=1e-7)  # over-relaxed for small steps (Q2)
    assert out[2] == 0.0


@pytest.mark.parametrize("err,step,want", [
    ((1e-9, 1e-9), (1e-9, 1e-9), 1),   # test_converged_success
    ((1.0, 1e-9), (1e-9, 1e-9), 0),    # test_converged_fail_err_v
    ((1e-9, 1.0), (1e-9, 1e-9), 0),    # test_converged_fail_err_i
    ((1e-9, 1e-9), (1.0, 1e-9), 0),    # test_converged_fail_step_v
    ((1e-9, 1e-9), (1e-9, 1.0), 0),    # test_converged_fail_step_i
])
def test_converged(L, err, step, want):
    assert L.orc_converged(err[0], err[1], step[0], step[1], 1.0, 1.0, 1.0, 1.0) == want


def test_converged_is_strict_and_nan_false(L):
    assert L.orc_converged(1e-6, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0) == 0  # 1e-6 < 1e-6 is false
    assert L.orc_converged(float("nan"), 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0) == 0


# ---- device stamps ------------------------------------------------------------------------------------
def elem(kind, nodes, val=0.0, branch=-1):
    e = orc.OrcElem()
    e.kind = kind
    for j in range(3):
        e.node[j] = nodes[j] if j < len(nodes) else -1
    e.branch = branch
    e.val = val
    return e


def apply(L, e, op, n, x=None, h=0.0, u=0.0, i=0.0, a=None, b=None):
    a = np.zeros((n, n)) if a is None else a
    b = np.zeros(n) if b is None else b
    x = np.zeros(max(n, 1)) if x is None else np.array(x, np.float64)
    rc = L.orc_elem_apply(ct.byref(e), u, i, op, n, dp(x), h, dp(a), dp(b))
    return a, b, rc


def test_res_stamps(L):
    """src/device/res.rs:80-155 (nodes[0] = "0" -> GND = -1)."""
    a, b, _ = apply(L, elem(0, [-1, 0], 1e3), 0, 1)
    assert a[0, 0] == 1e-3 and b[0] == 0.0
    a, b, _ = apply(L, elem(0, [0, -1], 1e3), 0, 1)
    assert a[0, 0] == 1e-3
    a, b, _ = apply(L, elem(0, [0, 1], 1e3), 0, 2)
    assert (a == np.array([[1e-3, -1e-3], [-1e-3, 1e-3]])).all() and (b == 0).all()
    a, b, _ = apply(L, elem(0, [0, 1], 1e3), 0, 3)
    a, b, _ = apply(L, elem(0, [0, 1], 1e3), 1, 3, a=a, b=b)
    assert (a == 0).all() and (b == 0).all()
    a, b, _ = apply(L, elem(0, [0, 1], 1e-3), 5, 2, x=[1.0, 2.0])  # nonlinear_stamp is a no-op
    assert (a == 0).all() and (b == 0).all()


def test_vdd_stamps(L):
    """src/device/vdd.rs:108-193: nodes = [vneg, vpos]; branch row = index of "V1"."""
    a, b, _ = apply(L, elem(3, [-1, 0], 1e-3, branch=1), 0, 2)   # nodes ["0","1"]
    m = np.zeros((2, 2)); m[0, 1] = 1.0; m[1, 0] = 1.0
    assert (a == m).all() and (b == [0.0, 1e-3]).all()
    a, b, _ = apply(L, elem(3, [0, -1], 1e-3, branch=1), 0, 2)   # nodes ["1","0"]
    assert (a == -m).all() and (b == [0.0, 1e-3]).all()
    a, b, _ = apply(L, elem(3, [0, 1], 1e-3, branch=2), 0, 3)    # nodes ["1","2"]
    m = np.zeros((3, 3)); m[2, 1] = m[1, 2] = 1.0; m[2, 0] = m[0, 2] = -1.0
    assert (a == m).all() and (b == [0.0, 0.0, 1e-3]).all()
    a, b, _ = apply(L, elem(3, [0, 1], 1e-3, branch=2), 1, 3, a=a, b=b)
    assert (a == 0).all() and (b == 0).all()


def test_idd_stamps(L):
    """src/device/idd.rs:90-148."""
    a, b, _ = apply(L, elem(4, [-1, 0], 1e-3), 0, 1)
    assert (a == 0).all() and b[0] == 1e-3
    a, b, _ = apply(L, elem(4, [0, -1], 1e-3), 0, 1)
    assert b[0] == -1e-3
    a, b, _ = apply(L, elem(4, [0, 1], 1e-3), 0, 2)
    assert (a == 0).all() and (b == [-1e-3, 1e-3]).all()
    a, b, _ = apply(L, elem(4, [0, 1], 1e-3), 1, 2, a=a, b=b)
    assert (b == 0).all()


@pytest.mark.parametrize("kind,val", [(1, 1e-6), (2, 1e-3)])
def test_cap_ind_dynamic_stamps(L, kind, val):
    """src/device/cap.rs:166-202, src/device/ind.rs:216-252: u_curr = 3, i_curr = 0, h = 1e-8."""
    e = elem(kind, [0, 1], val)
    a, b, _ = apply(L, e, 3, 2, x=[1.0, 2.0], h=1e-8, u=3.0, i=0.0)
    assert a[0, 0] > 0 and a[0, 1] < 0 and a[1, 0] < 0 and a[1, 1] > 0
    assert b[0] < 0 and b[1] > 0
    a, b, _ = apply(L, e, 4, 2, x=[1.0, 2.0], h=1e-8, u=3.0, i=0.0, a=a, b=b)
    assert (a == 0).all() and (b == 0).all()


def funcs(L, e, n, m, x):
    h = np.zeros((n, m))
    g = np.zeros(8)
    x = np.array(x, np.float64)
    cnt = L.orc_elem_funcs(ct.byref(e), n, m, 0, dp(x), dp(h), dp(g))
    return cnt, h, g[:cnt]


def test_diode(L):
    """src/device/diode.rs:138-276."""
    cnt, h, g = funcs(L, elem(5, [-1, 0]), 1, 1, [1.5, 1.0])
    assert cnt == 1 and h[0, 0] == -1.0 and g[0] < 0
    cnt, h, g = funcs(L, elem(5, [0, -1]), 1, 1, [1.5, 1.0])
    assert h[0, 0] == 1.0 and g[0] > 0
    cnt, h, g = funcs(L, elem(5, [0, 1]), 2, 1, [1.5, 1.0])
    assert (h[:, 0] == [1.0, -1.0]).all() and g[0] > 0
    a, b, _ = apply(L, elem(5, [-1, 0]), 5, 1, x=[1.0])
    assert a[0, 0] > 0 and b[0] < 0
    a, b, _ = apply(L, elem(5, [0, -1]), 5, 1, x=[1.0])
    assert a[0, 0] > 0 and b[0] > 0
    a, b, _ = apply(L, elem(5, [0, 1]), 5, 2, x=[1.0, 2.0])
    assert a[0, 0] > 0 and a[0, 1] < 0 and a[1, 0] < 0 and a[1, 1] > 0 and b[0] > 0 and b[1] < 0
    a, b, _ = apply(L, elem(5, [0, 1]), 0, 2)  # linear_stamp: no-op
    assert (a == 0).all() and (b == 0).all()


def test_npn(L):
    """src/device/npn.rs:193-254."""
    e = elem(6, [0, 1, 2])
    cnt, h, g = funcs(L, e, 3, 3, [2.0, 1.0, 0.0])
    assert cnt == 3 and (h == np.eye(3)).all()
    assert g[0] > 0 and g[1] > 0 and g[2] < 0
    a, b, _ = apply(L, e, 5, 3, x=[2.0, 1.0, 0.0])
    assert a[0, 0] > 0 and a[1, 1] > 0 and a[2, 2] > 0 and b[0] > 0 and b[1] > 0 and b[2] < 0


def test_nmos(L):
    """src/device/nmos.rs:193-262: gate row of the Jacobian all zero, ig == 0."""
    e = elem(7, [0, 1, 2])
    cnt, h, g = funcs(L, e, 3, 3, [2.0, 1.0, 0.0])
    assert cnt == 3 and (h == np.eye(3)).all()
    assert g[0] > 0 and g[1] == 0.0 and g[2] < 0
    a, b, _ = apply(L, e, 5, 3, x=[2.0, 1.0, 0.0])
    assert a[0, 0] > 0 and a[0, 1] > 0 and a[0, 2] < 0
    assert (a[1] == 0).all()
    assert a[2, 0] < 0 and a[2, 1] < 0 and a[2, 2] > 0
    assert b[0] > 0 and b[1] == 0.0 and b[2] < 0


def test_nmos_nan_is_unreachable(L):
    """src/device/nmos/model.rs:38-40 (quirk Q19)."""
    out = (ct.c_double * 4)()
    assert L.orc_nmos_model(float("nan"), 1.0, 0.0, out) == 4
    assert L.orc_nmos_model(2.0, 0.5, 0.0, out) == 0 and out[0] == 0.0  # cut-off


# ---- src/spice_fn.rs:8-124 (no reference test; pinned by construction) ----------------------------------
def test_spice_fn(L):
    p = (ct.c_double * 7)(0.0, 3.0, 1e8, 0, 0, 0, 0)
    assert L.orc_spice_fn_eval(1, p, 0.0) == 0.0
    assert abs(L.orc_spice_fn_eval(1, p, 2.5e-9) - 3.0) < 1e-12
    p = (ct.c_double * 7)(0.0, 3.0, 0.0, 10e-12, 10e-12, 5e-9, 10e-9)
    assert L.orc_spice_fn_eval(2, p, 5e-12) == 0.0 + (5e-12 - 0.0) / 10e-12 * 3.0
    assert L.orc_spice_fn_eval(2, p, 1e-9) == 3.0
    assert L.orc_spice_fn_eval(2, p, 4.995e-9) < 3.0   # the high phase ends at delay + pw - t_fall (spice_fn.rs:62-64)
    assert L.orc_spice_fn_eval(2, p, 6e-9) == 0.0
    assert L.orc_spice_fn_eval(2, p, 11e-9) == 3.0     # periodic
    p = (ct.c_double * 7)(3.0, 0.0, 0.0, 2e-9, 40e-9, 2e-9, 0)
    assert L.orc_spice_fn_eval(3, p, 0.0) == 3.0
    assert abs(L.orc_spice_fn_eval(3, p, 2e-9) - (3.0 + 3.0 * np.expm1(-1.0))) < 1e-15


def test_dc_count(L):
    """ndarray Array::range (src/engine.rs:120): ceil((stop-start)/step), stop exclusive."""
    assert L.orc_dc_count(0.0, 10.0, 0.1) == 100
    assert L.orc_dc_count(0.0, 3.0, 10 * 1e-3) == 300
    assert L.orc_dc_count(0.0, 10 * 1e-3, 100 * 1e-6) == 101  # 100u = 9.999999999999999e-05 (parser.rs:336)
