// solver_team.cuh — .tran for mid-size circuits (n > 16): G lanes per instance, 32 / G instances per warp in lock step.
//
// Why another family.  ncu on the warp-per-instance kernel (solver_sparse.cuh) at the named C4 size (64-node ladder,
// n = 65, linear) shows ~21 000 warp instructions per accepted step for what is, per step, one pair of sparse
// substitutions and ~13 damped updates of 65 unknowns with a 192-term residual: 86 % of the instructions walk maps,
// shuffle five-level butterflies for 65-element norms and run 65-row loops three lanes-rounds deep with the third round
// 1/32 full.  Here:
//   * a TEAM of G lanes (4 / 8 / 16 / 32) owns an instance; rows, devices and companion elements are dealt to the
//     lanes round-robin (n = 65 on 8 lanes: 9 rounds, 90 % full), norms need log2(G) shuffle levels;
//   * the 32 / G teams of a warp run in lock step at the level of a transient ATTEMPT (transient.rs:35-91): every team
//     builds its own right-hand side at its own t + h, all solve, all iterate Newton until the slowest has converged,
//     every team accepts / rejects on its own (PLTE, h control, state update) and a team whose instance is finished
//     fetches the next one — so the teams' data sit side by side in shared memory ([index][team]: conflict free) and a
//     step of the list-walking code serves 32 / G instances;
//   * the elimination follows a static plan (team_plan.h): pivot sequence fixed, fill-in / targets / updates /
//     substitution lists worked out on the host, the reference's pivot rule CHECKED at every step; an instance that
//     leaves the plan goes to the warp kernel (dynamic pivoting) and from there to the dense kernels;
//   * the residual uses the ASSEMBLED entries of A (sum of the stamps in the reference's order, then a(r,c) * x(c) with a
//     separate multiply and add, mna.rs:28) instead of one multiply-add per stamp term;
//   * what is touched once per accepted step — C/L state, the ring of the last three accepted x, for linear circuits
//     also the slot table and the factors — lives in a per-warp global scratch region with the same [index][team]
//     layout (coalesced, L2 resident): 3.6 KB of shared memory per instance at n = 65.
//   * LINEAR circuits with 3..12 rows per lane of 8 (template parameter RC; kern_team_rows*.cu) run a Newton loop with the rows
//     at compile-time offsets, x / x_proposed in registers, the convergence tests as thresholds on the class sums of squares and
//     the residual evaluated only in the iterations whose step test passes (see the loop and TeamOps::norm_threshold);
//   * NONLINEAR circuits (factors in shared memory) eliminate and substitute along LEVEL SCHEDULES (team_plan.h) with all lanes
//     of the team, keep the partial sums of the linear prefixes of their factor entries per attempt, and sum their longest
//     residual row (a supply rail) with the whole team.
// The solution-path arithmetic (assembly order, pivot rule, reciprocal scaling, separate multiply / subtract, damping,
// state update, PLTE) is the one of the other families, operation for operation: outputs compare equal bit for bit
// (tests/test_gpu_parity.py).  Reference: transient.rs:18-105, newtons_method.rs:19-89, gauss_lu.rs:3-54,
// state_history.rs:40-56, engine.rs:142-206.
#pragma once
#include "solver_sparse.cuh"
#include "team_plan.h"

namespace ftsb {

// offsets in doubles per team; an array lives in shared memory or in the warp's global region ("g" offsets)
struct TeamLayout {
    int ox, oxp, ob, oy, oA, oprev, ostg; // shared (stg: two chunks of substitution operands, team_plan.h)
    int oval, osv, oalpha;               // shared (PLACE bits 0 / 1) or global
    int ostate, oring, odynv, osvp;      // global (svp: partial sums of the linear prefixes of the factor entries, nonlinear circuits)
    int sm_doubles, gl_doubles;
};

// place: bit 0 = slot table in shared memory, bit 1 = factors in shared memory (nonlinear circuits: both are touched in
// every Newton iteration), bit 2 = the entries of A in the global region (large circuits: read once per iteration)
__host__ __device__ inline TeamLayout team_layout(int n, int nA, int S, int n_slots, int n_dyn, int n_src, int place, int n_pre = 0)
{
    TeamLayout t;
    int so = 0, go = 0;
    auto ts = [&](int c) {
        int r = so;
        so += c;
        return r;
    };
    auto tg = [&](int c) {
        int r = go;
        go += c;
        return r;
    };
    t.ox = ts(n + 1); // + the zero row (team_plan.h)
    t.oxp = ts(n + 1);
    t.ob = ts(n + 1);
    t.oy = ts(n);
    t.oA = (place & 4) ? tg(nA) : ts(nA);
    t.oprev = ts(3 * n_src);
    t.ostg = ts(2 * TEAM_CHUNK);
    t.oval = (place & 1) ? ts(n_slots) : tg(n_slots);
    t.osv = (place & 2) ? ts(S) : tg(S);
    t.oalpha = (place & 2) ? ts(n) : tg(n);
    t.ostate = tg(2 * n_dyn);
    t.oring = tg(3 * n);
    t.odynv = tg(n_dyn);
    t.osvp = tg(n_pre);
    t.sm_doubles = so;
    t.gl_doubles = go > 0 ? go : 1;
    return t;
}

struct TeamArgs {
    TeamPlan plan;
    TeamLayout lay;
    double *gscratch; // [grid * warps per block][gl_doubles][T]
    long long flops_factor, flops_solve;
};

constexpr int TEAM_WARPS_MAX = 8;  // warps per block (they share the ln_1p table); the launcher picks the count
constexpr int TEAM_WARPS_ROWS = 12; // the same for the kernels with compile-time rows per lane (RC > 0): three warps per scheduler leave 168 registers (four: 128)

template <int G>
struct TeamOps {
    static constexpr int T = 32 / G;
    using V = Vec<T>;

    static __device__ __forceinline__ void sum2(double &a, double &b)
    {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
    }
    static __device__ __forceinline__ int sumi(int v)
    {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    static __device__ __forceinline__ bool any(bool p, unsigned tmask) { return (__ballot_sync(0xffffffffu, p) & tmask) != 0u; }

    // row-indexed gather: out[row] = sum of +-val[slot] (reference accumulation order)
    static __device__ __forceinline__ double gsum(const Gather &g, int row, V val, double acc)
    {
        const int c0 = __ldg(&g.ptr[row]), c1 = __ldg(&g.ptr[row + 1]);
        for (int c = c0; c < c1; ++c) {
            const unsigned wd = (unsigned)__ldg(&g.con[c]);
            const double v = val[wd & ~CON_NEG];
            acc += (wd & CON_NEG) ? -v : v;
        }
        return acc;
    }

    // entries of A and of J from the slot table (fill-in entries of the plan start at 0.0)
    static __device__ __forceinline__ void assemble(const TeamPlan &P, V val, V A, V sv, int l, bool on)
    {
        if (on) {
            for (int e = l; e < P.nSlotA; e += G) { // (ELL padding slots have no contribution: 0.0)
                double acc = 0.0;
                const int c0 = __ldg(&P.aptr[e]), c1 = __ldg(&P.aptr[e + 1]);
                for (int c = c0; c < c1; ++c) {
                    const unsigned wd = (unsigned)__ldg(&P.acon[c]);
                    const double v = val[wd & ~CON_NEG];
                    acc += (wd & CON_NEG) ? -v : v;
                }
                A[e] = acc;
            }
        }
    }
    // PRE = 0: the whole contribution list of every static entry.  PRE = 1: only the linear prefixes, into svp (once per attempt
    // of a nonlinear circuit).  PRE = 2: continue from svp with the contributions behind the prefix (every Newton iteration) —
    // the same additions in the same order as PRE = 0, without re-summing what cannot have changed (a supply rail's diagonal
    // entry sums one conductance per stage: 254 dependent loads by one lane in every iteration of an inverter chain).
    template <int PRE>
    static __device__ __forceinline__ void assemble_sv(const TeamPlan &P, V val, V sv, V svp, int l, bool on)
    {
        if (on) {
            constexpr int EB = 4; // entries in flight per lane (the slot table may live in L2: overlap the gathers)
            for (int e0 = l; e0 < P.nJ; e0 += EB * G) {
                int c0[EB], len[EB], pix[EB], maxlen = 0;
                double acc[EB];
#pragma unroll
                for (int u = 0; u < EB; ++u) {
                    const int e = e0 + u * G;
                    const bool oke = e < P.nJ;
                    c0[u] = oke ? __ldg(&P.sptr[e]) : 0;
                    len[u] = oke ? __ldg(&P.sptr[e + 1]) - c0[u] : 0;
                    acc[u] = 0.0;
                    pix[u] = -1;
                    if (PRE != 0 && oke) {
                        const int pw = __ldg(&P.spre[e]), pre = pw >> 16;
                        if (PRE == 1) {
                            len[u] = pre;
                            if (pre > 0) pix[u] = pw & 0xffff;
                        } else if (pre > 0) {
                            c0[u] += pre;
                            len[u] -= pre;
                            acc[u] = svp[pw & 0xffff];
                        }
                    }
                    maxlen = max(maxlen, len[u]);
                }
                for (int tt = 0; tt < maxlen; ++tt) {
#pragma unroll
                    for (int u = 0; u < EB; ++u) {
                        if (tt < len[u]) {
                            const unsigned wd = (unsigned)__ldg(&P.scon[c0[u] + tt]);
                            const double v = val[wd & ~CON_NEG];
                            acc[u] += (wd & CON_NEG) ? -v : v;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < EB; ++u) {
                    const int e = e0 + u * G;
                    if (PRE == 1) {
                        if (pix[u] >= 0) svp[pix[u]] = acc[u];
                    } else if (e < P.nJ)
                        sv[e] = acc[u];
                }
            }
            if (PRE != 1)
                for (int e = P.nJ + l; e < P.S; e += G) sv[e] = 0.0;
        }
    }

    // gauss_lu.rs:3-42 along the plan.  Returns true when the team must leave the plan: the planned row is not the
    // reference's pivot (first position with the strictly largest |a|), or a candidate is zero-only / infinite / NaN.
    static __device__ bool factor(const TeamPlan &P, V sv, V alpha, int l, bool on, unsigned tmask)
    {
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        bool bad = false;
        const int n = P.n;
        int t0 = __ldg(&P.tptr[0]);
        for (int k = 0; k < n; ++k) {
            const int t1 = __ldg(&P.tptr[k + 1]);
            const int eP = __ldg(&P.pive[k]);
            double ap = 1.0;
            if (on) ap = sv[eP];
            const double abs_p = fabs(ap);
            bad |= !(abs_p > 0.0 && abs_p < INF);
            const double al = 1.0 / ap; // (:31)
            for (int tt = t0 + l; tt < t1; tt += G) {
                const int w = __ldg(&P.tgt[tt]);
                if (on) {
                    const int e = w & 0xffff;
                    const double a = sv[e];
                    const double av = fabs(a);
                    bad |= (w >> 30) ? !(av < abs_p) : !(av <= abs_p);
                    sv[e] = al * a; // (:32)
                }
            }
            if (on && l == 0) alpha[k] = al;
            if (t1 > t0) {
                __syncwarp();
                const int u0 = __ldg(&P.uptr[t0]), u1 = __ldg(&P.uptr[t1]);
                for (int u = u0 + l; u < u1; u += G) {
                    const int em = __ldg(&P.updm[u]);
                    const int ds = __ldg(&P.updds[u]);
                    if (on) {
                        const double m = sv[em];
                        if (m != 0.0) { // a - 0 * p == a
                            const int d = ds & 0xffff, s = (ds >> 16) & 0xffff;
                            sv[d] = sv[d] - m * sv[s]; // (:35-39)
                        }
                    }
                }
                __syncwarp();
            }
            t0 = t1;
        }
        return any(on && bad, tmask);
    }

    // y through the stored multipliers (gauss_lu.rs:40, in step order)
    static __device__ void forward(const TeamPlan &P, V sv, V y, int l, bool on)
    {
        const int n = P.n;
        int t0 = __ldg(&P.tptr[0]);
        for (int k = 0; k < n; ++k) {
            const int t1 = __ldg(&P.tptr[k + 1]);
            if (t1 > t0) {
                const int pk = __ldg(&P.pivrow[k]);
                for (int tt = t0 + l; tt < t1; tt += G) {
                    const int e = __ldg(&P.tgt[tt]) & 0xffff, r = __ldg(&P.tgtr[tt]);
                    if (on) {
                        const double m = sv[e];
                        if (m != 0.0) y[r] = y[r] - m * y[pk];
                    }
                }
                __syncwarp();
            }
            t0 = t1;
        }
    }

    // gauss_lu.rs:44-53; false when a solution component is not finite
    static __device__ bool backward(const TeamPlan &P, V sv, V alpha, V y, V xp, int l, bool on)
    {
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        const int n = P.n;
        bool ok = true;
        int b1 = __ldg(&P.bptr[n]);
        for (int i = n - 1; i >= 0; --i) {
            const int b0 = __ldg(&P.bptr[i]);
            const int pi = __ldg(&P.pivrow[i]), eP = __ldg(&P.pive[i]);
            double xi = 0.0;
            if (on) {
                xi = y[pi] / sv[eP];
                if (l == 0) xp[i] = xi;
            }
            ok = ok && (fabs(xi) < INF);
            if (b1 > b0) {
                for (int bb = b0 + l; bb < b1; bb += G) {
                    const int e = __ldg(&P.bwde[bb]), r = __ldg(&P.bwdr[bb]);
                    if (on) y[r] = y[r] - xi * sv[e];
                }
                __syncwarp();
            }
            b1 = b0;
        }
        return ok;
    }

    // y / d for a d whose correctly rounded reciprocal al = 1.0 / d is at hand: q0 = y * al is within an ulp of the
    // quotient, the exact remainder r = y - d * q0 (one FMA) corrects it, and q0 + r * al rounds to the correctly
    // rounded quotient (Markstein) as long as nothing overflows / underflows — checked on the exponents; everything
    // else (zero, tiny, huge, infinite, NaN operands) takes the division itself.  Bit-identical to y / d
    // (tests: test_team_exact_division_*), five instructions instead of ~30 on the critical path of the back substitution.
    static __device__ __noinline__ double div_plain(double y, double d) { return y / d; }
    static __device__ __forceinline__ double div_known_rcp(double y, double d, double al)
    {
        const unsigned ey = ((unsigned)__double2hiint(y) >> 20) & 0x7ffu, ed = ((unsigned)__double2hiint(d) >> 20) & 0x7ffu;
        if (ey - 623u < 800u && ed - 623u < 800u) { // both magnitudes in [2^-400, 2^400)
            const double q0 = y * al;
            const double r0 = fma(-d, q0, y);
            return fma(r0, al, q0);
        }
        return div_plain(y, d); // (out of line: the streams inline this function several times)
    }
    // s / c for a small positive integer c (class sizes): same construction; s >= 0 is a square root, zero or normal
    static __device__ __forceinline__ double div_count(double s, double c, double rc)
    {
        if (s < __longlong_as_double(0x7ff0000000000000LL)) {
            const double q0 = s * rc;
            const double r0 = fma(-c, q0, s);
            return fma(r0, rc, q0);
        }
        return s / c; // inf, NaN
    }

    // The class norm s = sqrt(q) / count of a sum of squares q (node_vec_norm.rs:12-33) enters the convergence test only as
    // `s < 0.001 * s + tol` (newtons_method.rs:79-89 with err_old == err, step_old == step: they are cloned at the end of the
    // previous iteration, :62-63).  sqrt and the division are correctly rounded and monotone in q, and s -> fl(fl(0.001 s) + tol)
    // is non-decreasing with slope << 1, so the test holds exactly for the q below one threshold: found here by bisection on
    // the bit pattern with the very operations the generic loop uses.  count = 0 (an empty class: 0 / 0 = NaN, never
    // converged) gives 0.0, which no q >= 0 is below.  tests/test_host_logic.py checks the monotonicity argument in numpy.
    static __device__ double norm_threshold(double dn, double rn, double tol)
    {
        auto ok = [&](double q) {
            const double s = div_count(sqrt(q), dn, rn);
            return s < 0.001 * s + tol;
        };
        if (!ok(0.0)) return 0.0;
        unsigned long long lo = 0ULL, hi = 0x7ff0000000000000ULL; // ok(lo), !ok(hi)
        while (hi - lo > 1ULL) {
            const unsigned long long mid = lo + ((hi - lo) >> 1);
            if (ok(__longlong_as_double((long long)mid)))
                lo = mid;
            else
                hi = mid;
        }
        return __longlong_as_double((long long)hi);
    }

    // The substitutions as op streams (team_plan.h): lane 0 of the team walks the ops in order — a chain of dependent
    // updates on a ladder — with the value it just wrote kept in a register; all lanes of the team fetch the factor
    // entries the ops need one chunk ahead (one L2 latency per TEAM_CHUNK ops instead of one per op) and hand them over
    // through a double-buffered staging area in shared memory.
    // forward: y through the stored multipliers (gauss_lu.rs:40)
    static __device__ void chain_forward(const TeamPlan &P, V sva, V y, V stg, int l, bool on)
    {
        constexpr int VPL = TEAM_CHUNK / G > 0 ? TEAM_CHUNK / G : 1; // (G <= TEAM_CHUNK: the kernel does not use the streams otherwise)
        double vreg[VPL];
        const int nchunk = P.nFchunk;
#pragma unroll
        for (int u = 0; u < VPL; ++u) {
            const int idx = __ldg(&P.fvidx[u * G + l]);
            vreg[u] = on ? sva[idx] : 0.0;
        }
        double last_v = 0.0;
        int last_r = -1;
        int o0 = __ldg(&P.fchunk[0]);
        for (int c = 0; c < nchunk; ++c) {
            const V buf = stg + (c & 1) * TEAM_CHUNK;
#pragma unroll
            for (int u = 0; u < VPL; ++u) buf[u * G + l] = vreg[u];
            __syncwarp();
            if (c + 1 < nchunk) {
#pragma unroll
                for (int u = 0; u < VPL; ++u) {
                    const int idx = __ldg(&P.fvidx[(c + 1) * TEAM_CHUNK + u * G + l]);
                    vreg[u] = on ? sva[idx] : 0.0;
                }
            }
            const int o1 = __ldg(&P.fchunk[c + 1]);
            if (on && l == 0) {
                int slot = 0;
                for (int op = o0; op < o1; ++op, ++slot) {
                    const int w0 = __ldg(&P.fop[op]);
                    const int r = w0 & 0xffff, pk = (w0 >> 16) & 0xffff;
                    const double m = buf[slot];
                    if (m != 0.0) {
                        const double ypk = pk == last_r ? last_v : y[pk];
                        const double yr = r == last_r ? last_v : y[r];
                        const double v = yr - m * ypk;
                        y[r] = v;
                        last_r = r;
                        last_v = v;
                    }
                }
            }
            o0 = o1;
        }
        __syncwarp();
    }
    // backward: gauss_lu.rs:44-53; false when a solution component is not finite
    static __device__ bool chain_backward(const TeamPlan &P, V sva, V y, V xp, V stg, int l, bool on, unsigned tmask)
    {
        constexpr int VPL = TEAM_CHUNK / G > 0 ? TEAM_CHUNK / G : 1;
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        double vreg[VPL];
        const int nchunk = P.nBchunk;
#pragma unroll
        for (int u = 0; u < VPL; ++u) {
            const int idx = __ldg(&P.bvidx[u * G + l]);
            vreg[u] = on ? sva[idx] : 0.0;
        }
        double last_v = 0.0, xi = 0.0;
        int last_r = -1;
        bool bad = false;
        int o0 = __ldg(&P.bchunk[0]);
        for (int c = 0; c < nchunk; ++c) {
            const V buf = stg + (c & 1) * TEAM_CHUNK;
#pragma unroll
            for (int u = 0; u < VPL; ++u) buf[u * G + l] = vreg[u];
            __syncwarp();
            if (c + 1 < nchunk) {
#pragma unroll
                for (int u = 0; u < VPL; ++u) {
                    const int idx = __ldg(&P.bvidx[(c + 1) * TEAM_CHUNK + u * G + l]);
                    vreg[u] = on ? sva[idx] : 0.0;
                }
            }
            const int o1 = __ldg(&P.bchunk[c + 1]);
            if (on && l == 0) {
                int slot = 0;
                for (int op = o0; op < o1; ++op) {
                    const int2 w = __ldg(reinterpret_cast<const int2 *>(P.bop) + op);
                    if (w.y < 0) { // x_i = y[pivot row] / pivot
                        const int pi = w.x & 0xffff, i = (w.x >> 16) & 0xffff;
                        const double d = buf[slot], al = buf[slot + 1];
                        slot += 2;
                        const double yv = pi == last_r ? last_v : y[pi];
                        xi = div_known_rcp(yv, d, al);
                        xp[i] = xi;
                        bad |= !(fabs(xi) < INF);
                    } else { // y[r] -= x_i * u(r, i)
                        const int r = w.x;
                        const double uv = buf[slot];
                        slot += 1;
                        const double yr = r == last_r ? last_v : y[r];
                        const double v = yr - xi * uv;
                        y[r] = v;
                        last_r = r;
                        last_v = v;
                    }
                }
            }
            o0 = o1;
        }
        __syncwarp();
        return !any(bad, tmask);
    }

    // ---- direct streams (team_plan.h): the factors live in shared memory, one lane walks the ops, no staging ----
    // pivot check, reciprocal and multipliers of elimination step k (gauss_lu.rs:6-32); returns true when the planned row
    // is not the reference's pivot or a candidate is zero-only / infinite / NaN
    static __device__ __forceinline__ bool step_head(const TeamPlan &P, V sv, V alpha, int k)
    {
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        const int t0 = __ldg(&P.tptr[k]), t1 = __ldg(&P.tptr[k + 1]);
        const double ap = sv[__ldg(&P.pive[k])];
        const double abs_p = fabs(ap);
        bool bad = !(abs_p > 0.0 && abs_p < INF);
        const double al = 1.0 / ap; // (:31)
        alpha[k] = al;
        for (int tt = t0; tt < t1; ++tt) {
            const int w = __ldg(&P.tgt[tt]);
            const int e = w & 0xffff;
            const double a = sv[e];
            const double av = fabs(a);
            bad |= (w >> 30) ? !(av < abs_p) : !(av <= abs_p);
            sv[e] = al * a; // (:32)
        }
        return bad;
    }
    static __device__ bool factor_direct(const TeamPlan &P, V sv, V alpha, int l, bool on, unsigned tmask)
    {
        bool bad = false;
        for (int idx = l; idx < P.nEarly; idx += G) { // steps no earlier update reaches: all at once, lanes in parallel
            const int k = __ldg(&P.early[idx]);
            if (on) bad |= step_head(P, sv, alpha, k);
        }
        __syncwarp();
        if (on && l == 0) { // the rest, in elimination order
            const int nop = P.nXop;
            const int2 *ops = reinterpret_cast<const int2 *>(P.xop);
            for (int o0 = 0; o0 < nop; o0 += 8) {
                int2 w[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) w[k] = o0 + k < nop ? __ldg(ops + o0 + k) : make_int2(0, 0);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (o0 + k < nop) {
                        if (w[k].y < 0)
                            bad |= step_head(P, sv, alpha, w[k].x);
                        else {
                            const double m = sv[w[k].y];
                            if (m != 0.0) { // a - 0 * p == a
                                const int d = w[k].x & 0xffff, s_ = (w[k].x >> 16) & 0xffff;
                                sv[d] = sv[d] - m * sv[s_]; // (:35-39)
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
        return any(on && bad, tmask);
    }
    static __device__ void forward_direct(const TeamPlan &P, V sv, V y, int l, bool on)
    {
        if (on && l == 0) {
            const int nop = P.nDFop;
            const int2 *ops = reinterpret_cast<const int2 *>(P.dfop);
            double last_v = 0.0;
            int last_r = -1;
            for (int o0 = 0; o0 < nop; o0 += 8) {
                int2 w[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) w[k] = o0 + k < nop ? __ldg(ops + o0 + k) : make_int2(0, 0);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (o0 + k < nop) {
                        const double m = sv[w[k].y];
                        if (m != 0.0) {
                            const int r = w[k].x & 0xffff, pk = (w[k].x >> 16) & 0xffff;
                            const double ypk = pk == last_r ? last_v : y[pk];
                            const double yr = r == last_r ? last_v : y[r];
                            const double v = yr - m * ypk; // (:40)
                            y[r] = v;
                            last_r = r;
                            last_v = v;
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
    static __device__ bool backward_direct(const TeamPlan &P, V sv, V alpha, V y, V xp, int l, bool on, unsigned tmask)
    {
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        bool bad = false;
        if (on && l == 0) {
            const int nop = P.nDBop;
            const int2 *ops = reinterpret_cast<const int2 *>(P.dbop);
            double last_v = 0.0, xi = 0.0;
            int last_r = -1;
            for (int o0 = 0; o0 < nop; o0 += 8) {
                int2 w[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) w[k] = o0 + k < nop ? __ldg(ops + o0 + k) : make_int2(0, 0);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (o0 + k < nop) {
                        const int row = w[k].x & 0xffff;
                        const double yv = row == last_r ? last_v : y[row];
                        if (w[k].y < 0) { // x_i = y[pivot row] / pivot (gauss_lu.rs:47)
                            const int i = (w[k].x >> 16) & 0xffff;
                            xi = div_known_rcp(yv, sv[w[k].y & 0xffff], alpha[i]);
                            xp[i] = xi;
                            bad |= !(fabs(xi) < INF);
                        } else { // y[r] -= x_i * u(r, i) (:49-52)
                            const double v = yv - xi * sv[w[k].y];
                            y[row] = v;
                            last_r = row;
                            last_v = v;
                        }
                    }
                }
            }
        }
        __syncwarp();
        return !any(bad, tmask);
    }

    // ---- level schedules (team_plan.h): the ops of a level on the lanes of the team, a warp barrier between levels.  The level
    // pointers are loaded two levels ahead and each lane's first op word of the NEXT level one level ahead, so that on the narrow
    // levels of a recurrence (one or two ops) no global load sits between two barriers.
    template <class F>
    static __device__ __forceinline__ void seg_walk(const int *seg, int nseg, const int2 *ops, int l, F &&exec)
    {
        const int2 *sg = reinterpret_cast<const int2 *>(seg);
        int2 sn = nseg > 0 ? __ldg(sg) : make_int2(0, 0);
        for (int si = 0; si < nseg; ++si) {
            const int2 sc = sn;
            if (si + 1 < nseg) sn = __ldg(sg + si + 1);
            if (sc.x < 0) { // a wide level
                const int2 *o = ops + sc.y;
                for (int k = l; k < -sc.x; k += G) exec(__ldg(o + k));
                __syncwarp();
            } else { // a run of narrow levels: the lane's op word of the next level is in flight while this one executes
                const int2 *o = ops + sc.y + l;
                const bool mine = l < TEAM_NARROW;
                int2 wn = mine ? __ldg(o) : make_int2(0, TEAM_OP_NOP);
                for (int j = 0; j < sc.x; ++j) {
                    const int2 w = wn;
                    if (mine && j + 1 < sc.x) wn = __ldg(o + (j + 1) * TEAM_NARROW);
                    if (mine) exec(w);
                    __syncwarp();
                }
            }
        }
    }
    // elimination + forward substitution (the right-hand side as one more column).  An update and a forward op are the same
    // statement on different arrays — d -= m * s — so they share one code path (the arrays sit in the same shared memory).
    static __device__ bool factor_levels(const TeamPlan &P, V sv, V alpha, V y, int l, bool on, unsigned tmask)
    {
        bool bad = false;
        seg_walk(P.xseg, P.nXseg, reinterpret_cast<const int2 *>(P.xsop), l, [&](const int2 w) {
            if (on) {
                if (w.y < 0)
                    bad |= step_head(P, sv, alpha, w.x);
                else if (!(w.y & TEAM_OP_NOP)) {
                    const double m = sv[w.y & 0xffff];
                    if (m != 0.0) { // a - 0 * p == a
                        const V t{(w.y & TEAM_OP_FWD) ? y.p : sv.p};
                        const int d = w.x & 0xffff, s_ = (w.x >> 16) & 0xffff;
                        t[d] = t[d] - m * t[s_]; // (:35-39, :40)
                    }
                }
            }
        });
        return any(on && bad, tmask);
    }
    static __device__ bool backward_levels(const TeamPlan &P, V sv, V alpha, V y, V xp, int l, bool on, unsigned tmask)
    {
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        bool bad = false;
        seg_walk(P.bseg, P.nBseg, reinterpret_cast<const int2 *>(P.bsop), l, [&](const int2 w) {
            if (on && !(w.y >= 0 && (w.y & TEAM_OP_NOP))) {
                const int row = w.x & 0xffff, i = (w.x >> 16) & 0xffff;
                if (w.y < 0) { // x_i = y[pivot row] / pivot (gauss_lu.rs:47)
                    const double xi = div_known_rcp(y[row], sv[w.y & 0xffff], alpha[i]);
                    xp[i] = xi;
                    bad |= !(fabs(xi) < INF);
                } else // y[r] -= x_i * u(r, i) (:49-52)
                    y[row] = y[row] - xp[i] * sv[w.y];
            }
        });
        return !any(on && bad, tmask);
    }

    // nonlinear devices at x -> slot values; returns team-wide (#non-finite currents) | panic << 20
    static __device__ __forceinline__ int eval_nl(const ModeProg &M, V x, V val, int l, bool on)
    {
        if (M.n_nl == 0) return 0;
        int acc = 0;
        for (int d = l; d < M.n_nl; d += G) {
            int desc[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) desc[k] = __ldg(&M.nl[d * 5 + k]);
            if (on) acc += eval_device(desc, x, val);
        }
        acc = sumi(acc);
        __syncwarp();
        return acc;
    }
};

// place: see team_layout.  RC > 0: LINEAR circuits whose rows per lane (n + G - 1) / G equal RC — the Newton loop keeps the
// lane's rows of x and x_proposed in registers, addresses its rows at compile-time offsets, decides convergence on the sums
// of squares (norm_threshold) and evaluates the residual only in the iterations whose step test passes (see the loop).
template <int G, int PLACE, int W, int RC = 0>
__global__ void __launch_bounds__(32 * (RC > 0 ? TEAM_WARPS_ROWS : TEAM_WARPS_MAX)) __maxnreg__(RC > 0 ? 168 : 128) k_team_tran(const __grid_constant__ RunArgs a, const __grid_constant__ TeamArgs ta)
{
    using Ops = TeamOps<G>;
    constexpr int T = Ops::T;
    using V = Vec<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *tab = reinterpret_cast<double *>(smem_raw);
    {
        unsigned long long *t_ = reinterpret_cast<unsigned long long *>(smem_raw);
        for (int i = threadIdx.x; i < 2 * FTS_LOG1P_N; i += blockDim.x) t_[i] = k_log1p_tab[i];
    }
    __syncthreads();
    const DevProg &DP = a.prog;
    const ModeProg &M = DP.tran;
    const TeamPlan &P = ta.plan;
    const TeamLayout &L = ta.lay;
    const int n = M.n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane / G, l = lane % G;
    const unsigned tmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (q * G));
    double *sw = tab + 2 * FTS_LOG1P_N + (size_t)warp * L.sm_doubles * T + q;
    double *gw = ta.gscratch + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * (size_t)L.gl_doubles * T + q;
    const V x{sw + L.ox * T}, xp{sw + L.oxp * T}, b{sw + L.ob * T}, y{sw + L.oy * T}, A{((PLACE & 4) ? gw : sw) + L.oA * T}, prev{sw + L.oprev * T};
    const V stg{sw + L.ostg * T};
    const V val{((PLACE & 1) ? sw : gw) + L.oval * T};
    const V sv{((PLACE & 2) ? sw : gw) + L.osv * T}, alpha{((PLACE & 2) ? sw : gw) + L.oalpha * T};
    const V state{gw + L.ostate * T}, ring{gw + L.oring * T}, dynv{gw + L.odynv * T}, svp{gw + L.osvp * T};
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const bool linear = RC > 0 || M.n_nl == 0;                       // (RC > 0: guaranteed by the launcher, as is the chain plan)
    const bool chain = RC > 0 || (P.chain != 0 && G <= TEAM_CHUNK); // single-lane op streams (ladders); wider teams: lane-parallel lists
    const int R = (n + G - 1) / G; // rows per lane: row i = j * G + l, j < R
    unsigned clsmask = 0;          // bit j: row j * G + l is a current row (NodeType::Current); R <= 32 (launcher)
    for (int j = 0; j < R && j < 32; ++j) {
        const int i = j * G + l;
        if (i < n && __ldg(&M.cls[i])) clsmask |= 1u << j;
    }
    auto is_cur = [&](int j, int) -> bool { return ((clsmask >> j) & 1u) != 0u; };
    const double dnv = (double)M.nv, dni = (double)M.ni, rnv = 1.0 / dnv, rni = 1.0 / dni;
    const bool has_ovf = P.nSlotA > W * (n + 1);
    const int long_row = RC > 0 ? -1 : P.longRow; // generic Newton loop: the row whose overflow entries the team sums together
    double thr_v = 0.0, thr_i = 0.0; // RC > 0: the convergence tests as thresholds on the class sums of squares
    if constexpr (RC > 0) {
        thr_v = Ops::norm_threshold(dnv, rnv, 1e-6);
        thr_i = Ops::norm_threshold(dni, rni, 1e-9);
    }
    // the zero row (index n of x, x_proposed, b; row n of A): written once, never touched by an instance
    if (l == 0) {
        x[n] = 0.0;
        xp[n] = 0.0;
        b[n] = 0.0;
    }
    __syncwarp();

    // per-team state (uniform over the lanes of a team)
    bool have = false, exhausted = false, new_step = true;
    long long inst = -1, nrec = 0, stored = 0;
    int status = ST_OK;
    double t = 0.0, h = 1e-18, hA = -1.0, hV = -1.0, h_first = 1e-18, tr0 = 0.0, tr1 = 0.0, tr2 = 0.0;
    unsigned long long ci[CN_LOCAL]; // the instance in flight (added to the global counters when it finishes)
#pragma unroll
    for (int i = 0; i < CN_LOCAL; ++i) ci[i] = 0;
    const unsigned long long total = a.in_list ? *a.in_count : (unsigned long long)a.batch;

    for (;;) {
        // ---------------------------------------------------------------- refill: idle teams fetch an instance
        {
            const bool need = !have && !exhausted;
            const unsigned nb = __ballot_sync(0xffffffffu, need && l == 0);
            if (nb) {
                unsigned long long first = 0;
                if (lane == 0) first = atomicAdd(a.work, (unsigned long long)__popc(nb));
                first = __shfl_sync(0xffffffffu, first, 0);
                if (need) {
                    const unsigned long long idx = first + (unsigned)__popc(nb & ((1u << (q * G)) - 1u));
                    if (idx >= total)
                        exhausted = true;
                    else {
                        inst = a.in_list ? a.in_list[idx] : (long long)idx;
                        have = true;
                    }
                }
            }
            const bool fresh = need && have;
            if (!__any_sync(0xffffffffu, have)) break;
            if (__any_sync(0xffffffffu, fresh)) {
                if (fresh) { // load_instance (engine.cuh) + start-up solution + init_state (cap.rs:38-47, ind.rs:42-49)
                    const ParamView vals = vals_of(a, inst), fn = fns_of(a, inst);
                    if (l == 0) val[SLOT_ONE] = 1.0;
                    for (int i = l; i < DP.n_res; i += G) {
                        const int e = __ldg(&DP.res[2 * i]), s = __ldg(&DP.res[2 * i + 1]);
                        val[s] = 1.0 / vals[e];
                    }
                    for (int i = l; i < DP.n_src; i += G) {
                        const int *sd = &DP.src[5 * i];
                        const int e = __ldg(&sd[0]), s = __ldg(&sd[1]), fk = __ldg(&sd[3]), fs = __ldg(&sd[4]);
                        const double v = fk != FN_NONE ? spice_fn_eval(fk, fn + fs * 7, 0.0) : vals[e];
                        val[s] = v;
                        prev[3 * i] = v;
                        prev[3 * i + 1] = v;
                        prev[3 * i + 2] = v;
                    }
                    const double *xs = a.x_start + inst * a.n_start;
                    for (int i = l; i < n; i += G) x[i] = xs[i];
                    for (int i = l; i < DP.n_dyn; i += G) {
                        const int *dd = &DP.dyn[6 * i];
                        const int e = __ldg(&dd[0]), kind = __ldg(&dd[2]), vp = __ldg(&dd[3]), vn = __ldg(&dd[4]), br = __ldg(&dd[5]);
                        dynv[i] = vals[e];
                        if (kind == K_C) {
                            state[2 * i] = (vp >= 0 ? xs[vp] : 0.0) - (vn >= 0 ? xs[vn] : 0.0);
                            state[2 * i + 1] = 0.0;
                        } else {
                            state[2 * i] = 0.0;
                            state[2 * i + 1] = xs[br];
                        }
                    }
                    const int st0 = a.start_status[inst];
                    status = st0;
                    t = 0.0;
                    h = 1e-18;
                    hA = -1.0;
                    hV = -1.0;
                    tr0 = tr1 = tr2 = 0.0;
                    nrec = stored = 0;
                    new_step = true;
#pragma unroll
                    for (int i = 0; i < CN_LOCAL; ++i) ci[i] = 0;
                }
                __syncwarp();
            }
        }
        bool act = have && status == ST_OK && t < a.t_stop; // this team runs an attempt in this round
        bool redo = false;
        int r_newton = 0; // n_iters >= 0, or -status

        if (__any_sync(0xffffffffu, act)) {
            // ------------------------------------------------------------ attempt prologue (transient.rs:35-46)
            if (act && new_step) {
                h_first = h;
                // companion models at the first attempt's h (cap/model.rs:10-17, ind/model.rs:10-17; quirk Q10)
                for (int i = l; i < DP.n_dyn; i += G) {
                    const int *dd = &DP.dyn[6 * i];
                    const int s = __ldg(&dd[1]), kind = __ldg(&dd[2]);
                    const double u = state[2 * i], ic = state[2 * i + 1];
                    double g; // kept in the slot table while h does not change (hV): same expression, same operands, same bits
                    if (h_first == hV)
                        g = val[s];
                    else {
                        const double pv = dynv[i];
                        g = kind == K_C ? 2.0 * pv / h_first : h_first / (2.0 * pv);
                        val[s] = g;
                    }
                    val[s + 1] = kind == K_C ? -(ic + g * u) : ic + g * u;
                }
                hV = h_first;
            }
            const double te = t + h;
            if (act) { // time-varying sources at t + h (transient.rs:36-42); I sources drift (quirk Q9)
                const ParamView fn = fns_of(a, inst);
                for (int i = l; i < DP.n_src; i += G) {
                    const int *sd = &DP.src[5 * i];
                    const int s = __ldg(&sd[1]), kind = __ldg(&sd[2]), fk = __ldg(&sd[3]), fs = __ldg(&sd[4]);
                    if (fk != FN_NONE) {
                        const double v = spice_fn_eval(fk, fn + fs * 7, te);
                        prev[3 * i + 2] = v;
                        val[s] = kind == K_V ? v : (prev[3 * i] - prev[3 * i + 1]) + v;
                    }
                }
            }
            __syncwarp();
            const bool mchg = act && h_first != hA; // the companion conductances changed: A (and, linear, its factors)
            if (__any_sync(0xffffffffu, mchg)) {
                Ops::assemble(P, val, A, sv, l, mchg);
                if (!linear) Ops::template assemble_sv<1>(P, val, sv, svp, l, mchg); // the linear prefixes of the factor entries
                if (linear) {
                    Ops::template assemble_sv<0>(P, val, sv, svp, l, mchg);
                    __syncwarp();
                    if (Ops::factor(P, sv, alpha, l, mchg, tmask)) redo = true;
                    if (mchg) {
                        ci[CN_FACTOR]++;
                        ci[CN_FLOPS] += (unsigned long long)ta.flops_factor;
                    }
                }
                if (mchg) hA = h_first;
                __syncwarp();
            }
            if (act) { // b of this attempt (sources do not change inside a Newton call)
                if constexpr (RC > 0) { // EBR rows per batch: their index and slot-table loads overlap (same sums, same order)
                    constexpr int EBR = 3;
                    for (int i0 = l; i0 < n; i0 += EBR * G) {
                        int c0[EBR], len[EBR], maxlen = 0;
                        double acc[EBR];
#pragma unroll
                        for (int e = 0; e < EBR; ++e) {
                            const int i = i0 + e * G;
                            c0[e] = i < n ? __ldg(&M.b_all.ptr[i]) : 0;
                            len[e] = i < n ? __ldg(&M.b_all.ptr[i + 1]) - c0[e] : 0;
                            maxlen = max(maxlen, len[e]);
                            acc[e] = 0.0;
                        }
                        for (int tt = 0; tt < maxlen; ++tt) {
                            unsigned wd[EBR];
#pragma unroll
                            for (int e = 0; e < EBR; ++e) wd[e] = tt < len[e] ? (unsigned)__ldg(&M.b_all.con[c0[e] + tt]) : 0u;
                            double v[EBR];
#pragma unroll
                            for (int e = 0; e < EBR; ++e) v[e] = tt < len[e] ? val[wd[e] & ~CON_NEG] : 0.0;
#pragma unroll
                            for (int e = 0; e < EBR; ++e)
                                if (tt < len[e]) acc[e] += (wd[e] & CON_NEG) ? -v[e] : v[e];
                        }
#pragma unroll
                        for (int e = 0; e < EBR; ++e) {
                            const int i = i0 + e * G;
                            if (i < n) {
                                b[i] = acc[e];
                                y[i] = acc[e];
                            }
                        }
                    }
                } else {
                    for (int i = l; i < n; i += G) {
                        const double bi = Ops::gsum(M.b_all, i, val, 0.0);
                        b[i] = bi;
                        if (linear) y[i] = bi;
                    }
                }
            }
            __syncwarp();
            bool it_on = act && !redo; // still iterating
            if (linear) { // one pair of substitutions per attempt: J == A, the proposed solution does not move
                if (chain) {
                    Ops::chain_forward(P, sv, y, stg, l, it_on);
                    if (!Ops::chain_backward(P, sv, y, xp, stg, l, it_on, tmask) && it_on) redo = true;
                } else {
                    Ops::forward(P, sv, y, l, it_on);
                    if (!Ops::backward(P, sv, alpha, y, xp, l, it_on) && it_on) redo = true;
                }
                if (it_on) {
                    ci[CN_SOLVE]++;
                    ci[CN_FLOPS] += (unsigned long long)ta.flops_solve;
                }
                it_on = it_on && !redo;
                __syncwarp();
            }
            // ------------------------------------------------------------ Newton (newtons_method.rs:19-71), all teams in rounds
            if constexpr (RC > 0) {
                // ---- linear circuit, RC rows per lane (row j * G + l, the last one clamped to the zero row).  The proposed solution
                // does not move inside the call, so every unknown iterates on its own — x_i += damp(xp_i - x_i) — and the unknowns
                // meet only in the norms: x and x_proposed stay in registers.  converged() (newtons_method.rs:79-89) is a
                // conjunction of the step test and the residual test, each on the current iteration's own norms; the residual is
                // therefore evaluated only in the iterations whose step test passes.  What the reference does with a residual
                // that is not evaluated here is the `is_infinite` panic (:53-55): it cannot fire when every |a(r,c)|, |b_r|,
                // |x_i| and |xp_i| of the call is below 1e70 (|x_k - xp| never grows under the damping, so every iterate stays
                // below 3e70 and a sum of n < 65 536 squares of |f_r| < 2e145 is finite); an instance that fails this guard
                // evaluates the residual in every iteration like the generic loop.
                const int il = min((RC - 1) * G + l, n);
                const int2 *ellc = reinterpret_cast<const int2 *>(P.ellcol);
                constexpr bool XP_REG = RC <= 10; // x_proposed in registers too (otherwise one shared-memory load per row and iteration)
                double xr[RC], xpr[XP_REG ? RC : 1];
                bool big = false;
#pragma unroll
                for (int j = 0; j < RC; ++j) {
                    const int i = j < RC - 1 ? j * G + l : il;
                    xr[j] = x[i];
                    const double xpj = xp[i];
                    if (XP_REG) xpr[j] = xpj;
                    big |= !(fabs(xr[j]) < 1e70) || !(fabs(xpj) < 1e70) || !(fabs(b[i]) < 1e70);
#pragma unroll
                    for (int w = 0; w < W; ++w) big |= !(fabs(A[i * W + w]) < 1e70);
                    if (has_ovf && i < n) {
                        const int v0 = __ldg(&P.ovfptr[i]), v1 = __ldg(&P.ovfptr[i + 1]);
                        for (int c = v0; c < v1; ++c) big |= !(fabs(A[W * (n + 1) + c]) < 1e70);
                    }
                }
                const bool lazy = !Ops::any(it_on && big, tmask);
                bool sok = false, eok = false;
                int it = 0;
                for (;;) {
                    if (it_on && !(it < 100 && !(sok && eok))) {
                        it_on = false;
                        r_newton = it < 100 ? it : -ST_NOT_CONVERGED;
                    }
                    if (!__any_sync(0xffffffffu, it_on)) break;
                    double qv = 0.0, qi = 0.0;
                    if (it_on) { // damped step (newtons_method.rs:73-77); x = x_new (:57-59)
                        constexpr int RB = 3; // rows in flight: their ln_1p chains interleave
#pragma unroll
                        for (int j0 = 0; j0 < RC; j0 += RB) {
                            double tt_[RB];
#pragma unroll
                            for (int u = 0; u < RB; ++u)
                                if (j0 + u < RC) {
                                    const double d = (XP_REG ? xpr[j0 + u] : xp[j0 + u < RC - 1 ? (j0 + u) * G + l : il]) - xr[j0 + u];
                                    tt_[u] = (1.3 / 16.0) * copysign(1.0, d) * fts_log1p_pos(16.0 * fabs(d), tab);
                                }
#pragma unroll
                            for (int u = 0; u < RB; ++u)
                                if (j0 + u < RC) {
                                    xr[j0 + u] = xr[j0 + u] + tt_[u];
                                    if (is_cur(j0 + u, 0))
                                        qi += tt_[u] * tt_[u];
                                    else
                                        qv += tt_[u] * tt_[u];
                                }
                        }
                    }
                    Ops::sum2(qv, qi);
                    sok = qv < thr_v && qi < thr_i;
                    const bool need = it_on && (sok || !lazy);
                    double fv = 0.0, fi = 0.0;
                    if (__any_sync(0xffffffffu, need)) {
                        if (need) {
#pragma unroll
                            for (int j = 0; j < RC; ++j) x[j < RC - 1 ? j * G + l : il] = xr[j];
                        }
                        __syncwarp();
                        if (need) { // f = A x - b (mna.rs:25-29) and its class sums (node_vec_norm.rs:12-33)
#pragma unroll
                            for (int j = 0; j < RC; ++j) {
                                const int row = j < RC - 1 ? j * G + l : il;
                                const int2 cols = __ldg(ellc + row);
                                double ax = 0.0;
#pragma unroll
                                for (int w = 0; w < W; ++w) { // the row's entries in column order, separate multiply and add
                                    const int cw = w < 2 ? cols.x : cols.y;
                                    const int col = (w & 1) ? (cw >> 16) & 0xffff : cw & 0xffff;
                                    ax = ax + A[row * W + w] * x[col];
                                }
                                if (has_ovf && row < n) {
                                    const int v0 = __ldg(&P.ovfptr[row]), v1 = __ldg(&P.ovfptr[row + 1]);
                                    for (int c = v0; c < v1; ++c) ax = ax + A[W * (n + 1) + c] * x[__ldg(&P.ovfcol[c])]; // long rows
                                }
                                const double f = ax + 0.0 - b[row];
                                if (is_cur(j, 0))
                                    fi += f * f;
                                else
                                    fv += f * f;
                            }
                        }
                        Ops::sum2(fv, fi);
                        __syncwarp();
                    }
                    if (it_on) {
                        eok = need && fv < thr_v && fi < thr_i;
                        ++it;
                        ci[CN_NEWTON]++;
                        if (need && (fv == INF || fi == INF)) { // the reference panics: reproduced from scratch by the warp kernel
                            redo = true;
                            it_on = false;
                        }
                    }
                }
                if (act) { // the last iterate (x is not restored after a rejected attempt either, quirk Q11)
#pragma unroll
                    for (int j = 0; j < RC; ++j) x[j < RC - 1 ? j * G + l : il] = xr[j];
                }
                __syncwarp();
            } else {
            int flags = Ops::eval_nl(M, x, val, l, it_on);
            if (it_on && (flags >> 20)) {
                redo = true; // NMOS_UNREACHABLE: the warp kernel reproduces the reference's panic from scratch
                it_on = false;
            }
            double sov = INF, soi = INF, eov = INF, eoi = INF, ev = INF, ei = INF, svn = INF, sin_ = INF;
            int it = 0;
            for (;;) {
                if (it_on && !(it < 100 && !(svn < 0.001 * sov + 1e-6 && sin_ < 0.001 * soi + 1e-9 && ev < 0.001 * eov + 1e-6 &&
                                             ei < 0.001 * eoi + 1e-9))) {
                    it_on = false;
                    r_newton = it < 100 ? it : -ST_NOT_CONVERGED;
                }
                if (!__any_sync(0xffffffffu, it_on)) break;
                if (!linear) {
                    if (it_on)
                        for (int i = l; i < n; i += G) y[i] = Ops::gsum(M.r_nl, i, val, b[i]);
                    Ops::template assemble_sv<2>(P, val, sv, svp, l, it_on);
                    __syncwarp();
                    if constexpr ((PLACE & 2) != 0) { // factors in shared memory: early steps in parallel, then one op stream
                        if (P.levels) { // level schedules: all lanes of the team, a barrier per level
                            if (Ops::factor_levels(P, sv, alpha, y, l, it_on, tmask) && it_on) redo = true; // (+ forward substitution)
                            if (!Ops::backward_levels(P, sv, alpha, y, xp, l, it_on, tmask) && it_on) redo = true;
                        } else { // early steps in parallel, then one lane walks the rest
                            if (Ops::factor_direct(P, sv, alpha, l, it_on, tmask) && it_on) redo = true;
                            Ops::forward_direct(P, sv, y, l, it_on);
                            if (!Ops::backward_direct(P, sv, alpha, y, xp, l, it_on, tmask) && it_on) redo = true;
                        }
                    } else {
                        if (Ops::factor(P, sv, alpha, l, it_on, tmask) && it_on) redo = true;
                        if (chain) {
                            Ops::chain_forward(P, sv, y, stg, l, it_on);
                            if (!Ops::chain_backward(P, sv, y, xp, stg, l, it_on, tmask) && it_on) redo = true;
                        } else {
                            Ops::forward(P, sv, y, l, it_on);
                            if (!Ops::backward(P, sv, alpha, y, xp, l, it_on) && it_on) redo = true;
                        }
                    }
                    if (it_on) {
                        ci[CN_FACTOR]++;
                        ci[CN_SOLVE]++;
                        ci[CN_FLOPS] += (unsigned long long)(ta.flops_factor + ta.flops_solve);
                    }
                    if (redo) it_on = false;
                    __syncwarp();
                }
                double qv = 0.0, qi = 0.0;
                if (it_on) { // damped step (newtons_method.rs:73-77); x is updated in place (x = x_new, :57-59)
                    constexpr int RB = 3; // rows in flight per lane: their ln_1p chains interleave
                    for (int j0 = 0; j0 < R; j0 += RB) {
                        double xi_[RB], d_[RB], tt_[RB];
                        int i_[RB];
#pragma unroll
                        for (int u = 0; u < RB; ++u) {
                            i_[u] = min((j0 + u) * G + l, n); // past the last row: the zero row (d = 0, t = 0)
                            xi_[u] = x[i_[u]];
                            d_[u] = xp[i_[u]] - xi_[u];
                        }
#pragma unroll
                        for (int u = 0; u < RB; ++u)
                            tt_[u] = (1.3 / 16.0) * copysign(1.0, d_[u]) * fts_log1p_pos(16.0 * fabs(d_[u]), tab);
#pragma unroll
                        for (int u = 0; u < RB; ++u) {
                            x[i_[u]] = xi_[u] + tt_[u];
                            if (is_cur(j0 + u, 0))
                                qi += tt_[u] * tt_[u];
                            else
                                qv += tt_[u] * tt_[u];
                        }
                    }
                }
                Ops::sum2(qv, qi);
                svn = Ops::div_count(sqrt(qv), dnv, rnv);
                sin_ = Ops::div_count(sqrt(qi), dni, rni);
                __syncwarp();
                flags = Ops::eval_nl(M, x, val, l, it_on);
                double fv = 0.0, fi = 0.0, ax_long = 0.0;
                const int nbad = flags & 0xfffff;
                auto row_tail = [&](int row, double axv, bool cur) { // + H g - b, squared into the row's class sum
                    double hg = 0.0;
                    if (!linear && row < n) {
                        int own_bad = 0;
                        const int f0 = __ldg(&M.f_nl.ptr[row]), f1 = __ldg(&M.f_nl.ptr[row + 1]);
                        for (int c = f0; c < f1; ++c) {
                            const unsigned wd = (unsigned)__ldg(&M.f_nl.con[c]);
                            const double v = val[wd & ~CON_NEG];
                            own_bad += isfinite(v) ? 0 : 1;
                            hg += (wd & CON_NEG) ? -v : v;
                        }
                        if (nbad > own_bad) hg = __longlong_as_double(0x7ff8000000000000LL);
                    }
                    const double f = axv + hg - b[row];
                    if (cur)
                        fi += f * f;
                    else
                        fv += f * f;
                };
                if (it_on) { // f = A x + H g - b (mna.rs:25-29) and its class norms (node_vec_norm.rs:12-33)
                    constexpr int RB = 3;
                    for (int j0 = 0; j0 < R; j0 += RB) {
                        int2 cols[RB];
                        int row_[RB];
                        double ax[RB];
#pragma unroll
                        for (int u = 0; u < RB; ++u) {
                            row_[u] = min((j0 + u) * G + l, n); // past the last row: the zero row (f = 0)
                            cols[u] = __ldg(reinterpret_cast<const int2 *>(P.ellcol) + row_[u]);
                            ax[u] = 0.0;
                        }
#pragma unroll
                        for (int w = 0; w < W; ++w) { // the row's entries in column order, a(r, c) * x(c) with separate multiply and add
#pragma unroll
                            for (int u = 0; u < RB; ++u) {
                                const int cw = w < 2 ? cols[u].x : cols[u].y;
                                const int col = (w & 1) ? (cw >> 16) & 0xffff : cw & 0xffff;
                                ax[u] = ax[u] + A[row_[u] * W + w] * x[col];
                            }
                        }
#pragma unroll
                        for (int u = 0; u < RB; ++u) {
                            const int row = row_[u];
                            if (row == long_row) { // (summed by the whole team below)
                                ax_long = ax[u];
                                continue;
                            }
                            if (has_ovf && row < n) {
                                const int v0 = __ldg(&P.ovfptr[row]), v1 = __ldg(&P.ovfptr[row + 1]);
                                for (int c = v0; c < v1; ++c) ax[u] = ax[u] + A[W * (n + 1) + c] * x[__ldg(&P.ovfcol[c])]; // long rows
                            }
                            row_tail(row, ax[u], is_cur(j0 + u, 0));
                        }
                    }
                }
                if (long_row >= 0) { // the longest row (a supply rail: an entry per stage): the lanes of the team form the products,
                                     // the sum runs over them in column order through shuffles — the same additions in the same
                                     // order as one lane's loop, without its dependent L2 loads
                    const int owner = long_row % G;
                    const int v0 = __ldg(&P.ovfptr[long_row]), v1 = __ldg(&P.ovfptr[long_row + 1]);
                    double acc = __shfl_sync(0xffffffffu, ax_long, owner, G);
                    for (int c0 = v0; c0 < v1; c0 += G) {
                        const int c = c0 + l;
                        double pr = 0.0;
                        if (it_on && c < v1) pr = A[W * (n + 1) + c] * x[__ldg(&P.ovfcol[c])];
                        const int m = min(G, v1 - c0);
                        for (int k = 0; k < m; ++k) acc = acc + __shfl_sync(0xffffffffu, pr, k, G);
                    }
                    if (it_on && l == owner) row_tail(long_row, acc, is_cur(long_row / G, 0));
                }
                Ops::sum2(fv, fi);
                if (it_on) {
                    ev = Ops::div_count(sqrt(fv), dnv, rnv);
                    ei = Ops::div_count(sqrt(fi), dni, rni);
                    if ((flags >> 20) || isinf(ev) || isinf(ei)) { // the reference panics: reproduced from scratch by the warp kernel
                        redo = true;
                        it_on = false;
                    }
                    ++it;
                    ci[CN_NEWTON]++;
                    eov = ev;
                    eoi = ei;
                    sov = svn;
                    soi = sin_;
                }
                __syncwarp();
            }
            }
            // ------------------------------------------------------------ accept / reject (transient.rs:48-91)
            bool accepted = false;
            double next_h = h;
            int nit = 0;
            const bool solved = act && !redo;
            const bool want_plte = solved && r_newton >= 0 && nrec >= 3;
            double pv = 0.0, pi = 0.0, xv = 0.0, xi2 = 0.0;
            if (__any_sync(0xffffffffu, want_plte)) {
                if (want_plte) { // PLTE = -0.5 h^3 dd3 over (ring0, ring1, ring2, candidate)  (state_history.rs:40-56)
                    const double hn = te - tr2;
                    const double c3 = -1.0 / 12.0;
                    const double alph = 6.0 * c3 * (hn * hn * hn);
                    const double a01 = 1.0 / (tr1 - tr0), a12 = 1.0 / (tr2 - tr1), a23 = 1.0 / (te - tr2);
                    const double a02 = 1.0 / (tr2 - tr0), a13 = 1.0 / (te - tr1), a03 = 1.0 / (te - tr0);
                    const int o0 = (int)(nrec % 3), o1 = (int)((nrec + 1) % 3), o2 = (int)((nrec + 2) % 3);
                    auto plte_row = [&](int j, int i) {
                        const double x0 = ring[o0 * n + i], x1 = ring[o1 * n + i], x2 = ring[o2 * n + i];
                        const double x3 = x[i];
                        const double d01 = a01 * (x1 - x0), d12 = a12 * (x2 - x1), d23 = a23 * (x3 - x2);
                        const double d02 = a02 * (d12 - d01), d13 = a13 * (d23 - d12);
                        const double d03 = a03 * (d13 - d02);
                        const double pl = alph * d03;
                        if (is_cur(j, i)) {
                            pi += pl * pl;
                            xi2 += x3 * x3;
                        } else {
                            pv += pl * pl;
                            xv += x3 * x3;
                        }
                    };
                    if constexpr (RC > 0) { // all rows unrolled: the ring loads (L2) of the lane's rows are in flight together
#pragma unroll
                        for (int j = 0; j < RC; ++j) {
                            const int i = j * G + l;
                            if (j < RC - 1 || i < n) plte_row(j, i);
                        }
                    } else {
#pragma unroll 3
                        for (int j = 0; j < R; ++j) {
                            const int i = j * G + l;
                            if (i >= n) break;
                            plte_row(j, i);
                        }
                    }
                }
                Ops::sum2(pv, pi);
                Ops::sum2(xv, xi2);
            }
            if (solved) {
                if (r_newton == -ST_NOT_CONVERGED) {
                    h *= 0.5;
                    ci[CN_REJ_NEWTON]++;
                } else if (nrec < 3) {
                    accepted = true;
                    nit = r_newton;
                } else {
                    const double plv = Ops::div_count(sqrt(pv), dnv, rnv), pli = Ops::div_count(sqrt(pi), dni, rni);
                    const double xnv = Ops::div_count(sqrt(xv), dnv, rnv), xni = Ops::div_count(sqrt(xi2), dni, rni);
                    const bool too_big = plv > xnv * 0.001 + 1e-3 || pli > xni * 0.001 + 1e-6; // transient.rs:99-101
                    if (too_big) {
                        h *= 0.5;
                        ci[CN_REJ_PLTE]++;
                    } else {
                        accepted = true;
                        nit = r_newton;
                        const bool can_grow = plv < 0.1 * 1e-3 && pli < 0.1 * 1e-6; // :103-105
                        next_h = (can_grow && h <= a.step_max / 2.0) ? h * 2.0 : h;
                    }
                }
                if (!accepted && h < 1e-18) status = ST_TRAN_H_UNDERFLOW; // transient.rs:88-90
                new_step = accepted;
            }
            if (__any_sync(0xffffffffu, accepted)) {
                if (accepted) { // push the record (state_history.rs:24-30), write it out per the recording policy
                    const int slot = (int)(nrec % 3);
                    for (int i = l; i < n; i += G) ring[slot * n + i] = x[i];
                    tr0 = tr1;
                    tr1 = tr2;
                    tr2 = te;
                    const int stride = a.rec_stride > 1 ? a.rec_stride : 1;
                    if (stored < a.cap && (nrec % stride) == 0) {
                        const long long row = inst * a.cap + stored;
                        if (a.x)
                            for (int i = l; i < n; i += G) a.x[row * n + i] = x[i];
                        if (l == 0) {
                            if (a.t_out) a.t_out[row] = te;
                            if (a.n_iters) a.n_iters[row] = nit;
                        }
                        ++stored;
                    }
                    ++nrec;
                    ci[CN_POINTS]++;
                    for (int i = l; i < DP.n_src; i += G) prev[3 * i + 1] = prev[3 * i + 2]; // vdd.rs:40-44, idd.rs:40-44
                    t += h;
                    // update_state with the accepted h (engine.rs:183-185; cap/model.rs:19-25, ind/model.rs:19-25)
                    for (int i = l; i < DP.n_dyn; i += G) {
                        const int *dd = &DP.dyn[6 * i];
                        const int kind = __ldg(&dd[2]), vp = __ldg(&dd[3]), vn = __ldg(&dd[4]);
                        const double u_old = state[2 * i], i_old = state[2 * i + 1], pvv = dynv[i];
                        const double u_new = xat(x, vp) - xat(x, vn);
                        double i_new;
                        if (kind == K_C) {
                            const double g = 2.0 * pvv / h;
                            i_new = g * (u_new - u_old) - i_old;
                        } else {
                            const double g = h / (2.0 * pvv);
                            i_new = g * (u_new + u_old) + i_old;
                        }
                        state[2 * i] = u_new;
                        state[2 * i + 1] = i_new;
                    }
                    h = nrec <= 3 ? h : next_h; // the first three records keep h (transient.rs:55-59)
                }
                __syncwarp();
            }
        }
        // ---------------------------------------------------------------- finished (or abandoned) instances
        const bool done = have && (redo || status != ST_OK || !(t < a.t_stop));
        if (__any_sync(0xffffffffu, done)) {
            if (done) {
                if (redo) { // left the plan: the warp kernel runs the instance from scratch (dynamic pivoting)
                    if (l == 0) {
                        a.redo_list[atomicAdd(a.redo_count, 1ULL)] = inst;
                        a.status[inst] = ST_REDO;
                    }
                } else {
                    if (a.x_final)
                        for (int i = l; i < n; i += G) a.x_final[inst * n + i] = x[i];
                    if (l == 0) {
                        a.n_rec[inst] = nrec;
                        a.status[inst] = status;
#pragma unroll
                        for (int i = 0; i < CN_LOCAL; ++i)
                            if (ci[i]) atomicAdd(&a.counters[i], ci[i]);
                    }
                }
                have = false;
            }
            __syncwarp();
        }
    }
}

typedef void (*TeamKernelFn)(const RunArgs, const TeamArgs);
TeamKernelFn team_kernel(int g, int place, int ell_w);          // kern_team.cu
TeamKernelFn team_kernel_rows(int g, int place, int ell_w, int rows); // kern_team_rows*.cu: linear circuits, rows per lane at compile time

} // namespace ftsb
