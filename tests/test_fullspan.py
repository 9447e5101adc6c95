"""Full-span parity of the two named .tran configs: C4 — 64-node ladder, 5.4 us / 1 ns, ~10 k accepted
steps — and C5 — 256-node NMOS inverter chain, 168 ns / 50 ps, ~5 k steps, where about 40 % of the +-10 % Monte-Carlo
instances END WITH THE REFERENCE'S ERROR (NMOS saturation / linear discontinuity, quirk Q18, then h < 1e-18).

Sampled instances of the full-size batches (global Monte-Carlo indices, so exactly the draws the 65 536 / 16 384
instance runs make) against (a) the committed oracle fixtures tests/golden/fullspan_c{4,5}.json (make_fullspan.py):
status, number of accepted steps, rejection and iteration totals, SHA-256 of all time stamps and of all per-record
iteration counts (bit for bit), unknowns at sampled records and at the end within 1e-9 / 1e-12; and (b) the live oracle on
a few of them, every record.  The failing SET of C5 must be the oracle's failing set."""
import hashlib
import json
import os

import numpy as np
import pytest

from ftspice_b200 import netlist as nl
from ftspice_b200.engine import BatchEngine

from harness import ROOT, assert_close

pytestmark = pytest.mark.gpu


def _fixture(which):
    with open(os.path.join(ROOT, "tests", "golden", f"fullspan_{which}.json")) as f:
        return json.load(f)


def _circuit(which):
    if which == "c4":
        return nl.load(nl.ladder_netlist(64, "SIN( 0.0 1.0 10M )", stop="5.4u", step="1n")), 12000
    return nl.load(nl.inverter_chain_netlist(256, stop="168n", step="50p")), 7000


def _params(circ, idx):
    vs, fs = zip(*(nl.monte_carlo(circ, 1, first=int(i)) for i in idx))
    return np.concatenate(vs), np.concatenate(fs)


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _run_and_check(which, opts, oracle_mod, live):
    fx = _fixture(which)
    circ, cap = _circuit(which)
    inst = fx["instances"]
    idx = [i["index"] for i in inst]
    vals, fns = _params(circ, idx)
    be = BatchEngine(circ)
    for k, v in opts.items():
        be.set_option(k, v)
    be.set_params(vals, fns)
    r = be.run_tran(fx["stop"], fx["step"], capacity=cap)
    c = be.counters()
    want_status = np.array([i["status"] for i in inst])
    want_nrec = np.array([i["n_rec"] for i in inst])
    assert (r.status == want_status).all(), (be.variant, np.flatnonzero(r.status != want_status), r.status, want_status)
    assert (r.n_rec == want_nrec).all(), (be.variant, np.flatnonzero(r.n_rec != want_nrec))
    assert c["solve_points"] == int(want_nrec.sum())
    assert c["newton_iters"] >= sum(i["newton_iters"] for i in inst)  # + the iterations of rejected attempts
    assert c["rej_plte"] == sum(i["rej_plte"] for i in inst) and c["rej_newton"] == sum(i["rej_newton"] for i in inst)
    for j, i in enumerate(inst):
        k = i["n_rec"]
        assert _sha(r.t[j, :k]) == i["t_sha256"], (which, i["index"], "time stamps differ from the oracle's")
        assert _sha(r.n_iters[j, :k].astype(np.int32)) == i["n_iters_sha256"], (which, i["index"], "iteration counts differ")
        assert int(r.n_iters[j, :k].sum()) == i["newton_iters"]
        for row, xr in zip(i["rows"], i["x_rows"]):
            assert_close(r.x[j, row], np.array([float(v) for v in xr]), f"{which} instance {i['index']} record {row}")
        if i["status"] == 0:  # (a failing instance ends inside a Newton limit cycle: its last iterate is not an output row)
            assert_close(r.x_final[j], np.array([float(v) for v in i["x_final"]]), f"{which} instance {i['index']} final")
    # live oracle on a few instances (failing ones included): every record
    st, nrec, it, t, x, xf, _ = oracle_mod.Oracle(circ).batch_tran(vals[live], fns[live], fx["stop"], fx["step"], cap,
                                                                   threads=len(live))
    for a, j in enumerate(live):
        k = int(nrec[a])
        assert st[a] == r.status[j] and k == r.n_rec[j]
        assert np.array_equal(t[a, :k], r.t[j, :k]) and np.array_equal(it[a, :k], r.n_iters[j, :k])
        assert_close(r.x[j, :k], x[a, :k], f"{which} instance {idx[j]} vs live oracle")
    return be.variant, r


@pytest.mark.parametrize("team", [1, 0])
def test_c4_full_span_sampled_instances(team, oracle_mod):
    v, r = _run_and_check("c4", {"team": team}, oracle_mod, live=[0, 1, 17, 40, 63])
    assert v.startswith("team" if team else "spw_n65"), v
    assert (r.status == 0).all() and r.n_rec.min() > 10000


def test_c5_full_span_sampled_instances_failing_set_equals_the_oracles(oracle_mod):
    fx = _fixture("c5")
    status = [i["status"] for i in fx["instances"]]
    ok = [j for j, s in enumerate(status) if s == 0]
    bad = [j for j, s in enumerate(status) if s != 0]
    assert len(ok) >= 8 and len(bad) >= 8  # the sample holds both kinds
    v, r = _run_and_check("c5", {}, oracle_mod, live=[ok[1], bad[0], bad[-1]])
    assert set(np.flatnonzero(r.status != 0)) == set(bad)
