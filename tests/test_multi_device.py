"""The multi-device entry points (ftsb_multi_*, include/ftsb.h): one handle over a LIST of devices, the batch sharded
contiguously inside the library, shards written straight into the caller's buffers.  On the one-GPU test box the list
names device 0 several times: every shard then has its own handle, host thread, stream and device buffers, which
exercises the sharding exactly as several GPUs would (ragged batches, empty shards, .op / .dc / .tran)."""
import numpy as np
import pytest

from ftspice_b200 import capi
from ftspice_b200 import netlist as nl
from ftspice_b200.engine import BatchEngine, MultiEngine

from harness import load


def test_multi_exports_and_no_device_fails_loudly():
    L = capi.lib()
    for s in capi.EXPORTS:
        assert hasattr(L, s), s
    if L.ftsb_device_count() == 0:  # the CPU container: the call must fail, not fall back
        with pytest.raises(capi.FtsbError) as e:
            MultiEngine(load("v_divider"))
        assert e.value.code == capi.E_CUDA and "no CPU fallback" in str(e.value)


@pytest.mark.gpu
@pytest.mark.parametrize("devices,B", [([0, 0, 0], 1000), ([0, 0], 37), ([0, 0, 0, 0], 3), (None, 64)])
def test_multi_solve_op_equals_single_device_bitwise(devices, B):
    circ = nl.load(nl.PROJ1_NL)
    vals, fns = nl.monte_carlo(circ, B, rel=0.3)
    me = MultiEngine(circ, devices)
    assert me.n_devices == (len(devices) if devices else capi.lib().ftsb_device_count())
    rm = me.solve_op(vals, fns)
    be = BatchEngine(circ)
    rs = be.solve_op(vals, fns)
    for a, b in ((rm.status, rs.status), (rm.n_iters, rs.n_iters), (rm.x, rs.x)):
        assert np.array_equal(a, b, equal_nan=True)
    assert me.counters()["solve_points"] == int((rs.status == 0).sum())
    assert me.launch_count >= 1 and "devices x [" in me.variant
    rm2 = me.solve_op(vals, fns)  # the handles are warm now (plans learned, per-topology kernel): same bits
    assert np.array_equal(rm2.x, rs.x, equal_nan=True) and np.array_equal(rm2.n_iters, rs.n_iters)


@pytest.mark.gpu
def test_multi_solve_dc_and_tran_equal_single_device_bitwise():
    from bench import NMOS_TEST
    circ = nl.load(NMOS_TEST)
    B, P = 50, 8
    vals = np.broadcast_to(circ.vals, (B, circ.n_elems)).copy()
    start = np.linspace(0.0, 3.0, B)
    step = np.full(B, 0.01)
    stop = start + P * step - 0.25 * step
    me = MultiEngine(circ, [0, 0, 0])
    rm = me.solve_dc(vals, None, "V01", start, stop, step, P)
    rs = BatchEngine(circ).solve_dc(vals, None, "V01", start, stop, step, P)
    for a, b in ((rm.status, rs.status), (rm.points_done, rs.points_done), (rm.n_iters, rs.n_iters), (rm.x, rs.x)):
        assert np.array_equal(a, b, equal_nan=True)

    circ = nl.load(nl.ladder_netlist(24, "PULSE( 0.0 1.0 0.0 10n 10n 50n 100n )", stop="30n", step="1n"))
    B = 11
    vals, fns = nl.monte_carlo(circ, B)
    tr = circ.command("tran")
    me = MultiEngine(circ, [0, 0, 0, 0])
    rm = me.solve_tran(vals, fns, tr.stop, tr.step, capacity=300)
    be = BatchEngine(circ)
    be.set_params(vals, fns)
    rs = be.run_tran(tr.stop, tr.step, capacity=300)
    for a, b in ((rm.status, rs.status), (rm.n_rec, rs.n_rec), (rm.n_iters, rs.n_iters), (rm.t, rs.t), (rm.x, rs.x),
                 (rm.x_final, rs.x_final)):
        assert np.array_equal(a, b, equal_nan=True)
    assert me.counters()["solve_points"] == int(rs.n_rec.sum())
