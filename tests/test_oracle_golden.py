"""Oracle vs the committed golden fixtures, and vs the independent scratch restatement quoted in
SURVEY.md App. A (two independent restatements of the reference agreeing is the best pin available without
a Rust toolchain)."""
import numpy as np
import pytest

from harness import fl, golden, load, netlist_names


@pytest.mark.parametrize("name", netlist_names())
def test_oracle_reproduces_golden(name, oracle_mod):
    g = golden()[name]
    c = load(name)
    o = oracle_mod.Oracle(c)
    if "op" in g:
        s, it, x, _ = o.run_op()
        assert s == g["op"]["status"]
        if s == 0:
            assert it == g["op"]["n_iters"] and np.array_equal(x, fl(g["op"]["x"]))
    if "dc" in g:
        d = c.command("dc")
        s, npts, it, x, sw, _ = o.run_dc(c.elem_index(d.source), d.start, d.stop, d.step)
        assert s == g["dc"]["status"] and npts == g["dc"]["points"]
        assert list(it) == g["dc"]["n_iters"]
        if len(it):
            assert np.array_equal(x, fl(g["dc"]["x"])) and np.array_equal(sw, fl(g["dc"]["sweep"]))
    if "tran" in g:
        t = c.command("tran")
        s, nrec, it, tt, x, xf, st = o.run_tran(t.stop, t.step)
        assert s == g["tran"]["status"] and nrec == g["tran"]["n_rec"]
        assert list(it) == g["tran"]["n_iters"]
        if nrec:
            assert np.array_equal(tt, fl(g["tran"]["t"])) and np.array_equal(x, fl(g["tran"]["x"]))
        assert st.rej_plte == g["tran"]["rej_plte"] and st.rej_newton == g["tran"]["rej_newton"]


# expectations of the survey's independent Python restatement
SURVEY_OP = {
    "v_divider": (30, {"1": 3.999999999643776, "2": 1.9999999999997917, "V1": -0.0009090909090909089}),
    "r_d_direct": (30, {"1": 3.999999999643776, "2": 0.5504989896412221, "V1": -0.0015679550047085322}),
    "r_d_inverse": (30, {"1": 3.999999999643776, "2": 3.999999997443776, "V1": -9.999999960040607e-13}),
    "nmos_test": (32, {"1": 0.8, "2": 5.000000001149253, "3": 4.989501051007703, "V01": 0.0,
                       "V02": -1.0498950104988994e-05}),
    "npn_test": (26, {"3": 0.11212054153535668, "4": 0.7087901806474356, "V01": -0.009120981935187396}),
    "proj1": (29, {"1": 1.8852722965078312, "8": 1.000000000000007, "V50": -0.19887991691548698, "V76": 0.0}),
    "proj2": (26, {"2": 0.6927424204716578, "7": 5.5088690808140654e-08, "V03": -0.004490469869405123}),
}
SURVEY_FAIL = ["i_divider", "i_divider_sweep", "rc_i_sine"]  # quirk Q3: no Current-type unknown -> NaN norm
SURVEY_DC = {"v_divider_sweep": (100, 1486, 1, 15), "proj2": (300, 3959, 13, 58)}
SURVEY_TRAN = {  # records, Newton iterations of the records, PLTE rejections
    "rc": (106, 106, 0), "rl": (123, 313, 0), "rc_sine": (213, 3060, 9), "rc_exp": (134, 1632, 6),
    "rc_pulse": (505, 4164, 233), "proj3": (125, 1399, 14),
}


@pytest.mark.parametrize("name", sorted(SURVEY_OP))
def test_op_matches_survey_restatement(name):
    g = golden()[name]["op"]
    it, vals = SURVEY_OP[name]
    assert g["status"] == 0 and g["n_iters"] == it
    c = load(name)
    x = fl(g["x"])  # stored by unknown index
    got = {h: float(x[c.index_of[h]]) for h in g["headers"]}
    for k, v in vals.items():
        assert abs(got[k] - v) <= 1e-12 * max(1.0, abs(v)), (name, k, got[k], v)


@pytest.mark.parametrize("name", SURVEY_FAIL)
def test_reference_failures_are_reproduced(name):
    g = golden()[name]
    for an in ("op", "dc", "tran"):
        if an in g:
            assert g[an]["status"] == 1


@pytest.mark.parametrize("name", sorted(SURVEY_DC))
def test_dc_counts_match_survey(name):
    g = golden()[name]["dc"]
    pts, tot, lo, hi = SURVEY_DC[name]
    it = np.array(g["n_iters"])
    assert len(it) == pts and it.sum() == tot and it.min() == lo and it.max() == hi


def test_dc_spot_values_match_survey():
    g = golden()["v_divider_sweep"]["dc"]
    x = fl(g["x"])
    sw = fl(g["sweep"])
    assert sw[3] == 0.30000000000000004  # start + step*k, not k/10
    np.testing.assert_allclose(x[1], [0.09999999964858593, 0.04999999990388598, -2.272727305301746e-05], rtol=1e-13)
    np.testing.assert_allclose(x[99], [9.899999999648587, 4.949999999903886, -0.002250000000325745], rtol=1e-13)


@pytest.mark.parametrize("name", sorted(SURVEY_TRAN))
def test_tran_counts_match_survey(name):
    g = golden()[name]["tran"]
    rec, iters, rej = SURVEY_TRAN[name]
    assert g["status"] == 0 and g["n_rec"] == rec and sum(g["n_iters"]) == iters and g["rej_plte"] == rej
    assert g["rej_newton"] == 0


def test_tran_spot_values_match_survey():
    g = golden()
    t = fl(g["rc"]["tran"]["t"])
    assert t[0] == 1e-18 and t[3] == 4e-18 and t[40] == 5.3687091220000015e-09 and t[-1] == 4.0265318402000016e-08
    assert fl(g["rl"]["tran"]["t"])[-1] == 0.04053239664633448
    assert fl(g["rc_sine"]["tran"]["t"])[40] == 1.8790481940000002e-09
    assert g["rc_sine"]["tran"]["n_iters"][40] == 16
    x40 = fl(g["rc_sine"]["tran"]["x"])[40]
    np.testing.assert_allclose(x40[:2], [2.7745497576841225, 0.2768839209877408], rtol=1e-12)
    assert fl(g["proj3"]["tran"]["t"])[-1] == 4.000946585800001e-08
