"""fts_log1p.h — the damping function's ln_1p (src/engine/newtons_method.rs:73-77) as the kernels compute it.
The reference calls the platform libm (Rust ``f64::ln_1p``), a ~1 ulp function; this one must stay under 1 ulp of
the exact value everywhere on its domain (u >= 0, +inf, NaN).  Host evaluation of the same header the kernels
include (``ftsb_debug_log1p(device=-1)``); the GPU suite checks that the device returns the same bits."""
import ctypes as ct

import mpmath as mp
import numpy as np

from ftspice_b200 import capi


def log1p_host(u: np.ndarray, device: int = -1) -> np.ndarray:
    L = capi.lib()
    fn = L.ftsb_debug_log1p
    fn.argtypes = [ct.c_int, ct.c_int64, ct.c_void_p, ct.c_void_p]
    fn.restype = ct.c_int
    u = np.ascontiguousarray(u, np.float64)
    o = np.empty_like(u)
    assert fn(device, u.size, u.ctypes.data, o.ctypes.data) == 0
    return o


def sample(seed: int = 1, n: int = 200000) -> np.ndarray:
    rng = np.random.default_rng(seed)
    edges = (1 + np.arange(0, 129) / 128.0)[:, None] * (2.0 ** np.arange(0, 40, 3))[None, :]
    edges = (edges[..., None] + np.spacing(edges)[..., None] * np.arange(-3, 4)[None, None, :]).ravel() - 1.0
    return np.concatenate([
        10.0 ** rng.uniform(-20, 20, n), rng.uniform(0, 2, n), rng.uniform(0, 1 / 64, n), 2.0 ** rng.uniform(-9, -5, n),
        16.0 * 10.0 ** rng.uniform(-12, 1, n),  # 16 |delta| of Newton steps from 1e-12 V to 10 V
        edges[edges >= 0],
        [0.0, 5e-324, 1e-320, 1e-300, 2.0 ** -53, 2.0 ** -52, 2.0 ** -7, 1.0, 3.0, 7.0, 2.0 ** 52, 2.0 ** 53, 2.0 ** 60, 1e308,
         1.7976931348623157e308]])


def test_specials():
    o = log1p_host(np.array([0.0, np.inf, np.nan, 5e-324, 1e-300]))
    assert o[0] == 0.0 and not np.signbit(o[0])
    assert o[1] == np.inf and np.isnan(o[2])
    assert o[3] == 5e-324 and o[4] == 1e-300


def test_error_below_one_ulp():
    u = sample()
    o = log1p_host(u)
    ref = np.log1p(u.astype(np.longdouble))  # 64-bit mantissa: resolves 2^-11 ulp of a double
    ulp = np.spacing(np.abs(ref.astype(np.float64))).astype(np.longdouble)
    err = np.abs((o.astype(np.longdouble) - ref) / ulp).astype(np.float64)
    assert np.isfinite(err).all()
    # exact ranking of the worst cases
    mp.mp.prec = 200
    worst = 0.0
    for j in np.argsort(err)[-100:]:
        ex = mp.log1p(mp.mpf(float(u[j])))
        worst = max(worst, abs(float((mp.mpf(float(o[j])) - ex) / mp.mpf(float(np.spacing(abs(float(ex))))))))
    assert worst < 0.8, worst
    # and it is, like glibc's, faithful: never more than one ulp away from the platform libm the reference calls
    g = np.log1p(u)
    assert (np.abs(o - g) <= np.spacing(np.abs(g))).all()


def test_monotone_on_a_fine_grid():
    u = np.sort(np.concatenate([np.linspace(0, 0.05, 200001), np.linspace(0.9, 1.1, 200001), np.linspace(15.9, 16.1, 100001)]))
    o = log1p_host(u)
    assert (np.diff(o) >= 0).all()
