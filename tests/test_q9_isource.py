"""Quirk Q9: time-varying CURRENT sources in .tran.

`transient::step` restores `b` from the backup taken at entry, which still holds the I(0) stamp of `run_tran`
(src/engine.rs:152-155), while the element keeps the value of its last evaluation (src/device/idd.rs:40-44); the
undo / redo of src/engine/transient.rs:36-42 therefore leaves `I(0) - I(t) + I(t + h)` in `b` for every attempt of
a step.  The reference's only netlist with such a source (test/rc_i_sine.sp) never gets that far (no V source:
NaN norm, quirk Q3), so these netlists add a V source: the .tran converges and the drift formula is exercised in
every kernel family.  CPU part: the oracle shows the drift; GPU part: every family against the oracle."""
import numpy as np
import pytest

from ftspice_b200 import netlist as nl

from harness import assert_close


def _small(src, extra="", stop="40n", step="1n"):
    return (f"* V source + time-varying I source\nV1 1 0 3V\nR12 1 2 R=1k\nC20 2 0 C=1p\nI20 2 0 {src}\n{extra}"
            f".TRAN {stop} {step}\n.END\n")


SMALL = {
    "rc_vi_sin": _small("SIN( 0.0 1m 50M )"),
    "rc_vi_pulse": _small("PULSE( 0.0 1m 2n 4n 4n 12n 24n )"),
    "rc_vi_exp": _small("EXP( 0.0 1m 2n 3n 15n 4n )"),
    "rl_vi_sin": _small("SIN( 0.5m 1m 50M )", "L23 2 3 L=1u\nR30 3 0 R=500\n"),
}


def mid_netlist(nodes=24, stop="40n"):
    """RC/RL ladder driven by a V source, with SIN / PULSE / EXP current sources injected along it (n > 16)."""
    text = nl.ladder_netlist(nodes, "SIN( 0.0 1.0 10M )", stop=stop, step="1n")
    w = len(str(nodes - 1))
    inj = (f"I1 n{5:0{w}d} 0 SIN( 0.0 2m 40M )\nI2 n{11:0{w}d} n{3:0{w}d} PULSE( 0.0 1m 2n 4n 4n 12n 24n )\n"
           f"I3 0 n{17:0{w}d} EXP( 0.2m 1m 2n 3n 15n 4n )\n")
    return text.replace(".TRAN", inj + ".TRAN")


@pytest.mark.parametrize("name", sorted(SMALL))
def test_oracle_shows_the_current_source_drift(name, oracle_mod):
    """With the drift the injected current is I(0) - I(t) + I(t+h) ~ h dI/dt, not I(t+h): the node the source drives moves
    by a fraction of the 1 V (1 mA x 1 kOhm) a drift-free source would give."""
    circ = nl.load(SMALL[name])
    tr = circ.command("tran")
    s, nrec, it, t, x, xf, st = oracle_mod.Oracle(circ).run_tran(tr.stop, tr.step)
    assert s == 0 and nrec > 100
    v2 = x[:, 1]  # unknown order: voltage nodes sorted ("1", "2", ...), then the V-source rows
    assert np.ptp(v2) > 0.05            # the source does act on the node ...
    assert np.ptp(v2) < 0.45            # ... but far less than the ~1 V (rc) / ~0.7 V (rl) swing without the drift


def _run(circ, vals, fns, opts, cap):
    from ftspice_b200.engine import BatchEngine
    be = BatchEngine(circ)
    for k, v in opts.items():
        be.set_option(k, v)
    be.set_params(vals, fns)
    tr = circ.command("tran")
    r = be.run_tran(tr.stop, tr.step, capacity=cap)
    c = be.counters()
    return be.variant, (r.status, r.n_rec, r.n_iters, r.t, r.x, r.x_final), c


def _check_vs_oracle(circ, vals, fns, out, oracle_mod, cap, what):
    tr = circ.command("tran")
    st, nrec, it, t, x, xf, _ = oracle_mod.Oracle(circ).batch_tran(vals, fns, tr.stop, tr.step, cap, threads=4)
    status, n_rec, n_iters, tt, xx, x_final = out
    assert (st == 0).all(), "the netlist must converge in the reference, or Q9 is not exercised"
    assert (status == st).all() and (n_rec == nrec).all(), (what, status, n_rec, nrec)
    assert (nrec < cap).all()
    for i in range(len(st)):
        k = int(nrec[i])
        assert (n_iters[i, :k] == it[i, :k]).all(), (what, i)
        assert (tt[i, :k] == t[i, :k]).all(), (what, i)
        assert_close(xx[i, :k], x[i, :k], f"{what} instance {i}")
    assert_close(x_final, xf, f"{what} x_final")


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SMALL))
def test_tran_current_source_small_circuits_all_families(name, oracle_mod):
    """n <= 16: k_lanes with 1 / 2 / 4 lanes per instance, the generic sub-warp kernel and the CTA kernel."""
    circ = nl.load(SMALL[name])
    B, cap = 5, 400
    vals, fns = nl.monte_carlo(circ, B)
    ref = None
    seen = set()
    for opts in ({}, {"lanes": 1}, {"lanes": 2}, {"lanes": 4}, {"force_generic": 1}, {"force_cta": 1}):
        v, out, c = _run(circ, vals, fns, opts, cap)
        seen.add(v.split("/")[0].split("+")[0])
        _check_vs_oracle(circ, vals, fns, out, oracle_mod, cap, f"{name} {opts} [{v}]")
        if ref is None:
            ref = out
        else:  # same solution-path arithmetic in every family: the same bits
            for a, b in zip(ref, out):
                assert np.array_equal(a, b, equal_nan=True), (name, opts, v)
    assert len(seen) >= 3, seen  # lanes kernel, generic sub-warp kernel, CTA kernel


@pytest.mark.gpu
def test_tran_current_source_mid_circuit_all_families(oracle_mod):
    """n > 16: the structure-exploiting / per-team kernels (default), the dense warp kernel, the generic kernel."""
    circ = nl.load(mid_netlist())
    B, cap = 6, 400
    vals, fns = nl.monte_carlo(circ, B)
    ref = None
    seen = set()
    for opts in ({}, {"team": 0}, {"sparse_warp": 0, "team": 0}, {"force_generic": 1}, {"force_cta": 1}):
        v, out, c = _run(circ, vals, fns, opts, cap)
        seen.add(v.split("/")[0].split("_")[0])
        _check_vs_oracle(circ, vals, fns, out, oracle_mod, cap, f"mid {opts} [{v}]")
        if ref is None:
            ref = out
        else:
            for a, b in zip(ref, out):
                assert np.array_equal(a, b, equal_nan=True), (opts, v)
    assert len(seen) >= 3, seen
