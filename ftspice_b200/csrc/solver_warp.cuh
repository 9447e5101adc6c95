// solver_warp.cuh — mid-size circuits (16 < n <= 96): one WARP per instance, RPT rows per lane.
//
// The whole warp is the team, so every shuffle / barrier uses the constant full mask.  Only J lives in shared
// memory (column-major, leading dimension = the analysis mode's n): like the small-circuit kernels, J is
// re-assembled from the slot values (J_full) and the residual's A*x comes from the flat map Ax, so neither A nor b is
// stored and (n*n + ~10n) doubles per instance suffice (n = 65: 39 KB, five instances per SM).
// Lane `rank` owns rows rank, rank+32, ... (RPT of them); rows never move, each carries its position; pivot =
// arg-max of (|a|, -position) (gauss_lu.rs:6-16); the pivot row is read from shared memory as a broadcast.
// Elimination uses separate multiply / subtract (gauss_lu.rs:35-41, no FMA contraction): bit-identical to the
// generic solver.  For linear circuits the factors are kept until the matrix changes (skip_refactor).
#pragma once
#include "engine.cuh"

namespace ftsb {

using WarpTeam = SubWarp<32>;

template <int RPT>
struct WarpSolver {
    using Team = WarpTeam;
    static constexpr bool kHasA = false;
    static constexpr bool kHasB = false;
    static constexpr bool kExternalStartup = true;
    struct State {
        bool jvalid;
        int pos[RPT]; // position (= elimination step) of this lane's rows in the cached factorisation
        __device__ State() : jvalid(false) {}
    };
    template <class W>
    static __device__ __forceinline__ void matrix_changed(const Team &, const ModeProg &, const W &, State &s)
    {
        s.jvalid = false;
    }

    static __device__ __forceinline__ void factor(const Team &tm, int n, double *J, double *y, int *piv, int (&pos)[RPT])
    {
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        int row[RPT];
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            row[q] = tm.rank + 32 * q;
            pos[q] = row[q];
        }
        for (int k = 0; k < n; ++k) {
            tm.sync();
            double v = -1.0;
            int p = INT_MAX;
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                if (row[q] < n && pos[q] >= k) {
                    double vv = fabs(J[k * n + row[q]]);
                    if (vv != vv) vv = (pos[q] == k) ? INF : -0.5;
                    if (vv > v || (vv == v && pos[q] < p)) {
                        v = vv;
                        p = pos[q];
                    }
                }
            }
            tm.argmax(v, p);
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                if (row[q] < n) {
                    if (pos[q] == p) {
                        pos[q] = k;
                        piv[k] = row[q];
                    } else if (pos[q] == k) {
                        pos[q] = p;
                    }
                }
            }
            tm.sync();
            const int P = piv[k];
            const double alpha = 1.0 / J[k * n + P];
            const double yk = y[P];
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                if (row[q] < n && pos[q] > k) {
                    const int r = row[q];
                    const double m = alpha * J[k * n + r];
                    J[k * n + r] = m;
                    for (int j = k + 1; j < n; ++j) J[j * n + r] = J[j * n + r] - m * J[j * n + P];
                    y[r] = y[r] - m * yk;
                }
            }
        }
        tm.sync();
    }

    static __device__ __forceinline__ void forward(const Team &tm, int n, const double *J, double *y, const int *piv,
                                                   const int (&pos)[RPT])
    {
        for (int k = 0; k < n - 1; ++k) {
            tm.sync();
            const double yk = y[piv[k]];
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                const int r = tm.rank + 32 * q;
                if (r < n && pos[q] > k) y[r] = y[r] - J[k * n + r] * yk;
            }
        }
        tm.sync();
    }

    static __device__ __forceinline__ void backward(const Team &tm, int n, const double *J, double *y, double *xp,
                                                    const int *piv, const int (&pos)[RPT])
    {
        for (int i = n - 1; i >= 0; --i) {
            const int ri = piv[i];
            const double xi = y[ri] / J[i * n + ri];
            if (tm.rank == 0) xp[i] = xi;
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                const int r = tm.rank + 32 * q;
                if (r < n && pos[q] < i) y[r] = y[r] - xi * J[i * n + r];
            }
            tm.sync();
        }
    }

    // f = A x + H g - b and its class norms from the flat maps
    template <class W, class V>
    static __device__ __forceinline__ void resid(const Team &tm, const ModeProg &M, const W &w, V x, int nbad, double &nv,
                                                 double &ni)
    {
        const int n = M.n;
        double sv = 0.0, si = 0.0;
        for (int row = tm.rank; row < n; row += 32) {
            double ax = 0.0;
            {
                const int c0 = __ldg(&M.Ax.ptr[row]), c1 = __ldg(&M.Ax.ptr[row + 1]);
                for (int c = c0; c < c1; ++c) {
                    const unsigned wd = (unsigned)__ldg(&M.Ax.con[c]);
                    const double v = w.val[wd & AX_SLOT_MASK];
                    const double xc = x[(wd & ~CON_NEG) >> AX_COL_SHIFT];
                    ax = ax + ((wd & CON_NEG) ? -v : v) * xc; // separate multiply and add, like the reference's dot (mna.rs:28)
                }
            }
            double hg = 0.0;
            if (M.n_gfun > 0) {
                int own_bad = 0;
                const int c0 = __ldg(&M.f_nl.ptr[row]), c1 = __ldg(&M.f_nl.ptr[row + 1]);
                for (int c = c0; c < c1; ++c) {
                    const unsigned wd = (unsigned)__ldg(&M.f_nl.con[c]);
                    const double v = w.val[wd & ~CON_NEG];
                    own_bad += isfinite(v) ? 0 : 1;
                    hg += (wd & CON_NEG) ? -v : v;
                }
                if (nbad > own_bad) hg = __longlong_as_double(0x7ff8000000000000LL);
            }
            double bb = 0.0;
            {
                const int c0 = __ldg(&M.b_all.ptr[row]), c1 = __ldg(&M.b_all.ptr[row + 1]);
                for (int c = c0; c < c1; ++c) {
                    const unsigned wd = (unsigned)__ldg(&M.b_all.con[c]);
                    const double v = w.val[wd & ~CON_NEG];
                    bb += (wd & CON_NEG) ? -v : v;
                }
            }
            const double f = ax + hg - bb;
            if (__ldg(&M.cls[row]))
                si += f * f;
            else
                sv += f * f;
        }
        tm.sum2(sv, si);
        nv = sqrt(sv) / (double)M.nv;
        ni = sqrt(si) / (double)M.ni;
    }

    // damped Newton (newtons_method.rs:19-71).  The initial get_err(x) of the reference never influences a decision
    // (err_old/step_old are overwritten before they are compared, :62-63), so only the device values at x are needed.
    template <class W>
    static __device__ __forceinline__ int newton(const Team &tm, const ModeProg &M, const W &w, State &s, Counters &cn,
                                                 int skip)
    {
        const int n = M.n;
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        const bool linear = (M.n_nl == 0);
        double *__restrict__ J = w.J.p;
        double *__restrict__ y = w.r.p;
        double *__restrict__ xp = w.xp.p;
        int *__restrict__ piv = w.piv.p;

        int flags = eval_nl(tm, M, w.x, w.val);
        if (flags >> 20) return -ST_NMOS_UNREACHABLE;
        double sov = INF, soi = INF, eov = INF, eoi = INF;
        double ev = INF, ei = INF, sv = INF, si = INF;
        bool xp_ready = false;
        int it = 0;
        while (it < 100 && !(sv < 0.001 * sov + 1e-6 && si < 0.001 * soi + 1e-9 && ev < 0.001 * eov + 1e-6 &&
                             ei < 0.001 * eoi + 1e-9)) {
            const bool do_factor = !(linear && skip && s.jvalid);
            const bool do_solve = !(linear && skip && xp_ready);
            if (do_solve) {
                for (int row = tm.rank; row < n; row += 32) { // r = b + nonlinear rhs, reference accumulation order
                    double acc = 0.0;
                    int c0 = __ldg(&M.b_all.ptr[row]), c1 = __ldg(&M.b_all.ptr[row + 1]);
                    for (int c = c0; c < c1; ++c) {
                        const unsigned wd = (unsigned)__ldg(&M.b_all.con[c]);
                        const double v = w.val[wd & ~CON_NEG];
                        acc += (wd & CON_NEG) ? -v : v;
                    }
                    c0 = __ldg(&M.r_nl.ptr[row]);
                    c1 = __ldg(&M.r_nl.ptr[row + 1]);
                    for (int c = c0; c < c1; ++c) {
                        const unsigned wd = (unsigned)__ldg(&M.r_nl.con[c]);
                        const double v = w.val[wd & ~CON_NEG];
                        acc += (wd & CON_NEG) ? -v : v;
                    }
                    y[row] = acc;
                }
                if (do_factor) {
                    for (int e = tm.rank; e < n * n; e += 32) J[e] = 0.0;
                    tm.sync();
                    gather(tm, M.J_full, w.J, decltype(w.J){nullptr}, w.val, n);
                    factor(tm, n, J, y, piv, s.pos);
                    s.jvalid = linear;
                    cn.c[CN_FACTOR]++;
                } else {
                    forward(tm, n, J, y, piv, s.pos);
                }
                backward(tm, n, J, y, xp, piv, s.pos);
                xp_ready = true;
                cn.c[CN_SOLVE]++;
            }
            double qv = 0.0, qi = 0.0;
            for (int i = tm.rank; i < n; i += 32) { // damped step (newtons_method.rs:73-77)
                const double xi = w.x[i];
                const double d = xp[i] - xi;
                const double t = (1.3 / 16.0) * copysign(1.0, d) * damp_log1p(16.0 * fabs(d));
                w.xn[i] = xi + t;
                if (__ldg(&M.cls[i]))
                    qi += t * t;
                else
                    qv += t * t;
            }
            tm.sum2(qv, qi);
            sv = sqrt(qv) / (double)M.nv;
            si = sqrt(qi) / (double)M.ni;
            tm.sync();

            flags = eval_nl(tm, M, w.xn, w.val);
            if (flags >> 20) return -ST_NMOS_UNREACHABLE;
            resid(tm, M, w, w.xn, flags & 0xfffff, ev, ei);
            if (isinf(ev) || isinf(ei)) return -ST_DIVERGED;
            for (int i = tm.rank; i < n; i += 32) w.x[i] = w.xn[i];
            tm.sync();
            ++it;
            cn.c[CN_NEWTON]++;
            eov = ev;
            eoi = ei;
            sov = sv;
            soi = si;
        }
        return it < 100 ? it : -ST_NOT_CONVERGED;
    }
};

// per-instance layout (contiguous), in doubles; ints follow
struct WarpLayout {
    int oJ, or_, ox, oxn, oxp, oval, oprev, ostate, oring, doubles, ints;
    size_t bytes;
};

__host__ __device__ inline WarpLayout warp_layout(int n, int n_slots, int n_dyn, int n_src, bool tran)
{
    WarpLayout t;
    int o = 0;
    auto take = [&](int c) {
        int r = o;
        o += (c + 1) & ~1;
        return r;
    };
    t.oJ = take(n * n);
    t.or_ = take(n);
    t.ox = take(n);
    t.oxn = take(n);
    t.oxp = take(n);
    t.oval = take(n_slots);
    t.oprev = take(tran ? 3 * n_src : 0);
    t.ostate = take(tran ? 2 * n_dyn : 0);
    t.oring = take(tran ? 3 * n : 0);
    t.doubles = o;
    t.ints = (n + 3) & ~3;
    t.bytes = (size_t)t.doubles * 8 + (size_t)t.ints * 4;
    return t;
}

template <int RPT, int AN>
__global__ void __launch_bounds__(32) k_warp(const __grid_constant__ RunArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevProg &P = a.prog;
    const ModeProg &M = AN == AN_OP ? P.op : AN == AN_DC ? P.dc : P.tran;
    const int n = M.n;
    const WarpLayout T = warp_layout(n, P.n_slots, P.n_dyn, P.n_src, AN == AN_TRAN);
    double *base = reinterpret_cast<double *>(smem_raw);
    Ws w;
    w.ld = n;
    w.A.p = nullptr;
    w.b.p = nullptr;
    w.J.p = base + T.oJ;
    w.r.p = base + T.or_;
    w.x.p = base + T.ox;
    w.xn.p = base + T.oxn;
    w.xp.p = base + T.oxp;
    w.val.p = base + T.oval;
    w.prev.p = AN == AN_TRAN ? base + T.oprev : nullptr;
    w.state.p = base + T.ostate;
    w.ring.p = base + T.oring;
    w.piv.p = reinterpret_cast<int *>(base + T.doubles);
    WarpTeam tm;
    Counters cn;
    for (;;) {
        long long inst = 0;
        if (tm.rank == 0) inst = next_instance(a);
        inst = __shfl_sync(0xffffffffu, inst, 0);
        if (inst < 0) break;
        run_instance<WarpTeam, WarpSolver<RPT>, AN>(tm, a, w, inst, cn);
        tm.sync();
    }
    flush_counters(tm, a, cn);
}

typedef void (*KernelFn)(const RunArgs);
KernelFn warp_kernel(int rpt, int an); // kern_warp.cu

} // namespace ftsb
