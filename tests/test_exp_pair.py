"""fts_exp.h — exp(a) and exp_m1(a) of a junction voltage from one argument reduction, as the diode and NPN models of
the kernels compute them (src/device/diode/model.rs:12-22, src/device/npn/model.rs:24-47).  The reference calls the
platform libm (Rust ``f64::exp`` / ``f64::exp_m1``, ~1 ulp functions); these must stay within 1 ulp of the exact value.
Host evaluation of the same header the kernels include; the GPU suite checks that the device returns the same bits."""
import ctypes as ct

import mpmath as mp
import numpy as np

from ftspice_b200 import capi


def exp_pair(a: np.ndarray, device: int = -1):
    L = capi.lib()
    fn = L.ftsb_debug_exp_pair
    fn.argtypes = [ct.c_int, ct.c_int64, ct.c_void_p, ct.c_void_p, ct.c_void_p]
    fn.restype = ct.c_int
    a = np.ascontiguousarray(a, np.float64)
    e = np.empty_like(a)
    m = np.empty_like(a)
    assert fn(device, a.size, a.ctypes.data, e.ctypes.data, m.ctypes.data) == 0
    return e, m


def sample(seed: int = 1, n: int = 200000) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return np.concatenate([
        rng.uniform(-690, 690, n), rng.uniform(-40, 40, n), rng.uniform(-1, 1, n), rng.uniform(-0.02, 0.02, n),
        rng.uniform(-5, 5, n) / 26e-3 / 8,  # junction voltages / Vt
        np.sign(rng.uniform(-1, 1, n)) * 10.0 ** rng.uniform(-300, 0, n),
        (np.arange(-2000, 2000) + 0.5) * np.log(2) / 128,  # reduction boundaries
        [0.0, -0.0, 690.0, -690.0, 1e-320, -1e-320]])


def test_specials_and_fallback_range():
    e, m = exp_pair(np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 700.0, -745.0, -800.0, 710.0]))
    assert e[0] == 1.0 and e[1] == 1.0 and m[0] == 0.0 and m[1] == 0.0 and np.signbit(m[1])
    assert e[2] == np.inf and m[2] == np.inf and e[3] == 0.0 and m[3] == -1.0
    assert np.isnan(e[4]) and np.isnan(m[4])
    assert e[5] == np.exp(700.0) and m[6] == -1.0 and e[7] == 0.0 and e[8] == np.inf


def test_error_below_one_ulp():
    a = sample()
    e, m = exp_pair(a)
    al = a.astype(np.longdouble)

    def ulps(o, ref):
        u = np.spacing(np.abs(ref.astype(np.float64))).astype(np.longdouble)
        return np.abs((o.astype(np.longdouble) - ref) / u).astype(np.float64)

    ee = ulps(e, np.exp(al))
    rm = np.expm1(al)
    em = ulps(m, rm)
    em[rm == 0] = 0
    mp.mp.prec = 200
    we = wm = 0.0
    for j in np.argsort(ee)[-60:]:
        ex = mp.exp(mp.mpf(float(a[j])))
        we = max(we, abs(float((mp.mpf(float(e[j])) - ex) / mp.mpf(float(np.spacing(float(ex)))))))
    for j in np.argsort(em)[-60:]:
        ex = mp.expm1(mp.mpf(float(a[j])))
        wm = max(wm, abs(float((mp.mpf(float(m[j])) - ex) / mp.mpf(float(np.spacing(abs(float(ex))))))))
    assert we < 0.6 and wm < 0.6, (we, wm)
    # never more than one ulp away from the platform libm the reference calls
    ge, gm = np.exp(a), np.expm1(a)
    assert (np.abs(e - ge) <= np.spacing(ge)).all()
    assert (np.abs(m - gm) <= np.spacing(np.abs(gm))).all()
