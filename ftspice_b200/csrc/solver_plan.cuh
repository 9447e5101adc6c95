// solver_plan.cuh — small circuits (n <= 16), .op and .dc: one THREAD per instance interpreting learned elimination
// plans (plan.h).  32 instances per warp in lock-step rounds like k_sync (solver_lanes.cuh); the per-instance working
// set holds only the structurally non-zero matrix entries (compact numbering), so more warps fit per SM than with the
// dense thread-per-instance mapping, and the elimination executes only the updates on structural non-zeros.
//
// Exactness: the interpreter performs the reference's operations (gauss_lu.rs:29-53: reciprocal of the pivot, multiplier
// = alpha * a, separate multiply / subtract, column-oriented back substitution) on every structurally non-zero entry;
// what it leaves out are updates whose multiplier or pivot-row entry is a structural zero (`a - 0*p == a` for finite
// p).  At every step it verifies the planned pivot against the reference's rule over the structural candidates; a
// zero / non-finite pivot candidate, an unknown pivot sequence or a non-finite solution component hands the instance
// to the dense kernel (redo list), which repeats it with the full dense arithmetic.
#pragma once
#include "plan.h"
#include "solver_lanes.cuh"
#include "solver_sparse.cuh" // ST_REDO

namespace ftsb {

// G lanes per instance (SyncTeam<G>: all collectives on the full, converged warp).  Every team of the warp executes
// the same number of steps, barriers and shuffles whatever plan it follows; a team whose factorisation has failed
// keeps going on its (garbage) data with `dead` set, so that the warp stays converged.
template <int G>
struct PlanExec {
    static constexpr int TS = 32 / G;

    // gauss_lu.rs:3-42 along a plan, rhs carried along.  Returns false -> redo.  `plan` = the plan finally followed.
    static __device__ __forceinline__ bool factor(const SyncTeam<G> &ts, const PlanProg &pp, double *__restrict__ sv,
                                                  double *__restrict__ y, int &plan)
    {
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        const int n = pp.n;
        int p = 0;
        bool dead = false;
        for (int k = 0; k < n; ++k) {
            int4 rec = __ldg(&pp.rec[p * n + k]);
            ts.sync(); // the updates of step k-1
            // pivot search over the structural candidates: first position with the strictly largest |a| (:6-16)
            const int nc = (rec.x >> 16) & 255, coff = rec.y & 0xffff;
            double bv = -1.0;
            int key = 0x7fffff00; // pos << 8 | row, bit 30 = a non-finite candidate; smaller pos wins a tie
            int bad = 0;
            for (int c = ts.rank; c < nc; c += G) {
                const int wd = __ldg(&pp.cand[coff + c]);
                const double v = fabs(sv[(wd & 255) * TS]);
                const int kk = (((wd >> 16) & 255) << 8) | ((wd >> 8) & 255);
                bad |= !(v < INF);
                if (v > bv || (v == bv && kk < key)) {
                    bv = v;
                    key = kk;
                }
            }
#pragma unroll
            for (int o = 16; o >= TS; o >>= 1) {
                const double v2 = __shfl_xor_sync(0xffffffffu, bv, o);
                const int k2 = __shfl_xor_sync(0xffffffffu, key, o);
                bad |= __shfl_xor_sync(0xffffffffu, bad, o);
                if (v2 > bv || (v2 == bv && k2 < key)) {
                    bv = v2;
                    key = k2;
                }
            }
            dead = dead || bad || !(bv > 0.0);
            const int brow = key & 255;
            if (!dead && brow != ((rec.x >> 8) & 255)) { // another row wins: follow the plan that takes it, if any
                const int q = pp.next[(p * n + k) * 16 + brow];
                if (q < 0)
                    dead = true;
                else {
                    p = q;
                    rec = __ldg(&pp.rec[p * n + k]);
                }
            }
            const double alpha = 1.0 / sv[(rec.x & 255) * TS]; // (:31)
            const double yk = y[((rec.x >> 8) & 255) * TS];
            const int nm = (rec.x >> 24) & 255, moff = (rec.y >> 16) & 0xffff;
            for (int t = ts.rank; t < nm; t += G) {
                const int wd = __ldg(&pp.mult[moff + t]);
                const int e = (wd & 255) * TS, r = ((wd >> 8) & 255) * TS;
                const double m = alpha * sv[e];
                sv[e] = m;
                y[r] = y[r] - m * yk; // (:40)
            }
            ts.sync();
            const int nu = (rec.z >> 16) & 0xffff, uoff = rec.z & 0xffff;
            for (int u = ts.rank; u < nu; u += G) {
                const int wd = __ldg(&pp.upd[uoff + u]);
                const int d = (wd & 255) * TS;
                sv[d] = sv[d] - sv[((wd >> 8) & 255) * TS] * sv[((wd >> 16) & 255) * TS]; // (:35-39)
            }
        }
        ts.sync();
        plan = p;
        return !dead;
    }

    // a new right-hand side through the stored multipliers of plan p
    static __device__ __forceinline__ void forward(const SyncTeam<G> &ts, const PlanProg &pp, const double *__restrict__ sv,
                                                   double *__restrict__ y, int p)
    {
        const int n = pp.n;
        for (int k = 0; k < n - 1; ++k) {
            const int4 rec = __ldg(&pp.rec[p * n + k]);
            ts.sync();
            const double yk = y[((rec.x >> 8) & 255) * TS];
            const int nm = (rec.x >> 24) & 255, moff = (rec.y >> 16) & 0xffff;
            for (int t = ts.rank; t < nm; t += G) {
                const int wd = __ldg(&pp.mult[moff + t]);
                const int r = ((wd >> 8) & 255) * TS;
                y[r] = y[r] - sv[(wd & 255) * TS] * yk;
            }
        }
        ts.sync();
    }

    // gauss_lu.rs:44-53; false when a solution component is not finite
    static __device__ __forceinline__ bool backward(const SyncTeam<G> &ts, const PlanProg &pp, const double *__restrict__ sv,
                                                    double *__restrict__ y, double *xp, int p)
    {
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        const int n = pp.n;
        bool ok = true;
        for (int i = n - 1; i >= 0; --i) {
            const int4 rec = __ldg(&pp.rec[p * n + i]);
            const double xi = y[((rec.x >> 8) & 255) * TS] / sv[(rec.x & 255) * TS];
            if (ts.rank == 0) xp[i * TS] = xi;
            ok = ok && (fabs(xi) < INF);
            const int nb = (rec.w >> 16) & 0xffff, boff = rec.w & 0xffff;
            for (int t = ts.rank; t < nb; t += G) {
                const int wd = __ldg(&pp.bu[boff + t]);
                const int r = ((wd >> 8) & 255) * TS;
                y[r] = y[r] - xi * sv[(wd & 255) * TS];
            }
            ts.sync();
        }
        return ok;
    }
};

// per-instance layout in doubles (each element strided by 32/G inside the warp's region)
struct PlanLayout {
    int osv, oy, ox, oxp, oxn, oval, doubles;
};
__host__ __device__ inline PlanLayout plan_layout(int n, int S, int n_slots, bool linear)
{
    PlanLayout t;
    int o = 0;
    t.osv = o;
    o += S;
    t.oy = o;
    o += n;
    t.ox = o;
    o += n;
    t.oxp = o;
    o += n;
    // x_new overwrites the proposed solution in place, except for linear circuits, which reuse the proposal of a solve
    t.oxn = linear ? o : t.oxp;
    o += linear ? n : 0;
    t.oval = o;
    o += n_slots;
    t.doubles = o;
    return t;
}

template <int G, int AN>
__global__ void __launch_bounds__(32) k_plan(const __grid_constant__ RunArgs a)
{
    static_assert(AN == AN_OP || AN == AN_DC, "the plan kernel covers .op and .dc");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TS = 32 / G;
    const DevProg &P = a.prog;
    const ModeProg &M = AN == AN_OP ? P.op : P.dc;
    const PlanProg &pp = a.plan;
    const int n = M.n, S = pp.S;
    const bool linear = (M.n_nl == 0);
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const PlanLayout T = plan_layout(n, S, P.n_slots, linear);
    LaneTeam<G> tl; // run-time mask: used only in the (divergent, rare) transition code
    SyncTeam<G> ts; // constant full mask: used in the lock-step body
    double *base = reinterpret_cast<double *>(smem_raw) + tl.slot;
    WsT<TS> w;
    w.ld = n;
    w.A.p = nullptr;
    w.b.p = nullptr;
    w.J.p = base + T.osv * TS;
    w.r.p = base + T.oy * TS;
    w.x.p = base + T.ox * TS;
    w.xp.p = base + T.oxp * TS;
    w.xn.p = base + T.oxn * TS;
    w.val.p = base + T.oval * TS;
    w.state.p = nullptr;
    w.ring.p = nullptr;
    w.prev.p = nullptr;
    w.piv.p = nullptr;
    w.aux = nullptr;
    double *__restrict__ sv_ = w.J.p;
    double *__restrict__ yp = w.r.p;
    double *xpp = w.xp.p; // may alias x_new (see plan_layout)
    Counters cn; // finished instances
    Counters ci; // the instance in flight (dropped when the instance goes to the dense kernel, which re-counts it)
    const int lane = threadIdx.x & 31;

    // per-team state (identical in the lanes of a team)
    int phase = PH_FETCH;
    long long inst = -1;
    bool jvalid = false, xp_ready = false, redo = false;
    int plan = 0, it = 0;
    double sv = INF, si = INF, ev = INF, ei = INF, sov = INF, soi = INF, eov = INF, eoi = INF;
    long long k = 0, count = 0;
    double dc_start = 0.0, dc_step = 0.0;
    const int sweep_slot = AN == AN_DC ? __ldg(&P.src[5 * a.sweep_src + 1]) : 0;

    for (;;) {
        // ---------------- transitions ----------------
        {
            const unsigned need = __ballot_sync(0xffffffffu, phase == PH_FETCH && tl.rank == 0);
            if (need) { // one atomic per warp for all teams that need an instance
                long long first = 0;
                if (lane == 0) first = (long long)atomicAdd(a.work, (unsigned long long)__popc(need));
                first = __shfl_sync(0xffffffffu, first, 0);
                const unsigned below = need & ((1u << tl.slot) - 1u);
                if (phase == PH_FETCH) inst = first + __popc(below);
            }
        }
        if (phase == PH_FETCH) {
            if (inst < a.batch) {
                load_instance(tl, a, w, inst);
                for (int i = tl.rank; i < n; i += G) {
                    w.x[i] = 0.0; // x0 = 0 (mna.rs:21-23)
                    w.xn[i] = 0.0;
                }
                tl.sync();
                if (AN == AN_DC) {
                    dc_start = a.start[inst];
                    dc_step = a.step[inst];
                    const double cnt_d = ceil((a.stop[inst] - dc_start) / dc_step); // ndarray Array::range
                    count = cnt_d >= 0.0 ? (long long)cnt_d : 0;
                    if (count > a.max_points) count = a.max_points;
                    k = 0;
                    if (tl.rank == 0) w.val[sweep_slot] = dc_start + dc_step * 0.0;
                    tl.sync();
                }
                const int fl = eval_nl(tl, M, w.x, w.val); // device values at x0
                ci = Counters();
                jvalid = false;
                xp_ready = false;
                it = 0;
                sv = si = ev = ei = sov = soi = eov = eoi = INF;
                phase = PH_SOLVE;
                if (AN == AN_DC && count == 0) {
                    if (tl.rank == 0) {
                        a.points_done[inst] = 0;
                        a.status[inst] = ST_OK;
                    }
                    phase = PH_FETCH;
                } else if (fl >> 20) {
                    if (tl.rank == 0) {
                        a.status[inst] = ST_NMOS_UNREACHABLE;
                        if (AN == AN_OP) a.n_iters[inst] = 0;
                        if (AN == AN_DC) a.points_done[inst] = 0;
                    }
                    phase = PH_FETCH;
                }
            } else {
                phase = PH_IDLE;
            }
        }
        if (__all_sync(0xffffffffu, phase != PH_SOLVE)) {
            if (__all_sync(0xffffffffu, phase == PH_IDLE)) break;
            continue;
        }

        // ---------------- C: assemble, solve, damped step ----------------
        const bool solving = phase == PH_SOLVE;
        const bool do_factor = !(linear && a.skip_refactor && jvalid);
        const bool do_solve = !(linear && a.skip_refactor && xp_ready);
        const bool any_solve = __any_sync(0xffffffffu, solving && do_solve);
        const bool any_factor = __any_sync(0xffffffffu, solving && do_factor);
        if (any_solve) {
            for (int row = ts.rank; row < n; row += G) { // r = b + nonlinear rhs, reference accumulation order
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
                yp[row * TS] = acc;
            }
            bool ok = true;
            if (any_factor) {
                // a team whose cached factors are still valid re-creates exactly the same factors: harmless.
                // Structural entries from the slot table; the plans' fill-in entries start at +0.0.
                const Gather &g = M.J_full;
                for (int d = ts.rank; d < g.n_dst; d += G) {
                    double acc = 0.0;
                    const int c0 = __ldg(&g.ptr[d]), c1 = __ldg(&g.ptr[d + 1]);
                    for (int c = c0; c < c1; ++c) {
                        const unsigned wd = (unsigned)__ldg(&g.con[c]);
                        const double v = w.val[wd & ~CON_NEG];
                        acc += (wd & CON_NEG) ? -v : v;
                    }
                    sv_[__ldg(&pp.jdst[d]) * TS] = acc;
                }
                for (int e = pp.nnz0 + ts.rank; e < S; e += G) sv_[e * TS] = 0.0;
                ok = PlanExec<G>::factor(ts, pp, sv_, yp, plan);
                jvalid = linear && ok;
                if (solving) ci.c[CN_FACTOR]++;
            } else {
                PlanExec<G>::forward(ts, pp, sv_, yp, plan);
            }
            ok = PlanExec<G>::backward(ts, pp, sv_, yp, xpp, plan) && ok;
            if (solving) {
                redo = !ok;
                ci.c[CN_SOLVE]++;
            }
            xp_ready = true;
        }
        double qv = 0.0, qi = 0.0;
        for (int i = ts.rank; i < n; i += G) { // damped step (newtons_method.rs:73-77)
            const double xi = w.x[i];
            const double d = xpp[i * TS] - xi;
            const double t = (1.3 / 16.0) * copysign(1.0, d) * damp_log1p(16.0 * fabs(d));
            w.xn[i] = xi + t;
            if (__ldg(&M.cls[i]))
                qi += t * t;
            else
                qv += t * t;
        }
        ts.sum2(qv, qi);
        sv = sqrt(qv) / (double)M.nv;
        si = sqrt(qi) / (double)M.ni;
        ts.sync();

        // ---------------- A: device evaluation and residual at x_new ----------------
        const int flags = eval_nl(ts, M, w.xn, w.val);
        LaneSolver<G, 1>::resid(ts, M, w, w.xn, flags & 0xfffff, ev, ei);

        // ---------------- B: accept, count, decide (per team, no collectives) ----------------
        if (solving) {
            int fin = -1; // -1: keep iterating
            if (redo)
                fin = ST_REDO;
            else if (flags >> 20)
                fin = ST_NMOS_UNREACHABLE;
            else if (isinf(ev) || isinf(ei))
                fin = ST_DIVERGED;
            else {
                for (int i = ts.rank; i < n; i += G) w.x[i] = w.xn[i];
                ++it;
                ci.c[CN_NEWTON]++;
                eov = ev;
                eoi = ei;
                sov = sv;
                soi = si;
                const bool conv = sv < 0.001 * sov + 1e-6 && si < 0.001 * soi + 1e-9 && ev < 0.001 * eov + 1e-6 &&
                                  ei < 0.001 * eoi + 1e-9;
                if (!(it < 100 && !conv)) fin = it < 100 ? ST_OK : ST_NOT_CONVERGED;
            }
            if (fin == ST_REDO) { // the dense kernel repeats the whole instance (for .dc: the whole chain)
                if (ts.rank == 0) {
                    a.redo_list[atomicAdd(a.redo_count, 1ULL)] = inst;
                    a.status[inst] = ST_REDO;
                }
                redo = false;
                phase = PH_FETCH;
            } else if (fin >= 0) {
                if (AN == AN_OP) {
                    for (int i = ts.rank; i < n; i += G) a.x[inst * n + i] = w.x[i];
                    if (ts.rank == 0) {
                        a.n_iters[inst] = fin == ST_OK ? it : 0;
                        a.status[inst] = fin;
                    }
                    if (fin == ST_OK) ci.c[CN_POINTS]++;
                    for (int i = 0; i < CN_LOCAL; ++i) cn.c[i] += ci.c[i];
                    phase = PH_FETCH;
                } else {
                    if (fin == ST_OK) {
                        double *xo = a.x + (inst * a.max_points + k) * n;
                        for (int i = ts.rank; i < n; i += G) xo[i] = w.x[i];
                        if (ts.rank == 0) a.n_iters[inst * a.max_points + k] = it;
                        ci.c[CN_POINTS]++;
                        ++k;
                    }
                    if (fin != ST_OK || k >= count) { // `?` aborts the sweep at the first failure (engine.rs:127)
                        if (ts.rank == 0) {
                            a.points_done[inst] = k;
                            a.status[inst] = fin;
                        }
                        for (int i = 0; i < CN_LOCAL; ++i) cn.c[i] += ci.c[i];
                        phase = PH_FETCH;
                    } else { // next sweep point, warm start (engine.rs:116-127); device values at x stay valid
                        if (ts.rank == 0) w.val[sweep_slot] = dc_start + dc_step * (double)k;
                        xp_ready = false;
                        it = 0;
                        sv = si = ev = ei = sov = soi = eov = eoi = INF;
                    }
                }
            }
        }
        ts.sync();
    }
#pragma unroll
    for (int i = 0; i < CN_LOCAL; ++i) { // counters are kept by rank 0 of each team
        unsigned long long v = tl.rank == 0 ? cn.c[i] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&a.counters[i], v);
    }
}

KernelFn plan_kernel(int g, int an); // kern_plan.cu

} // namespace ftsb
