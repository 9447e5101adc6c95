// solver_lanes.cuh — small circuits (n <= 16): G = 1, 2 or 4 lanes per instance, working set in shared memory.
//
// A warp runs 32/G instances in lock step.  With G = 1 one THREAD owns one instance: no shuffles, no reductions, no
// barriers and no divergence in the device models (all lanes evaluate the same device at the same time) — the
// mapping with the fewest instructions per flop.  Its weakness is occupancy: the per-instance working set (J is
// n*n doubles) lives in shared memory, so only a few warps fit per SM.  G = 2 or 4 lanes share one instance (rows
// split by their *position* modulo G) and buy 2-4x the warps for ~10-20 % more instructions.
//
// Layout: interleaved — element e of instance slot s at 8-byte word e*(32/G) + s; the lanes of an instance are
// s, s + 32/G, ... (strided), so a half-warp touches every slot's bank pair once: accesses are conflict free even
// when instances touch different rows (different pivots).
// Elimination: pivot row and the row being updated are held in registers (compile-time column count NC == leading
// dimension), one LDS and one STS per element update, separate multiply / subtract like the reference
// (gauss_lu.rs:35-41; no FMA contraction) => bit-identical to the generic solver and to the oracle's LU.
// Neither A nor b is kept: J is re-assembled from the slot values through J_full and the residual's A*x comes from
// the flat per-row map Ax (the residual only feeds convergence decisions).
#pragma once
#include <type_traits>

#include "engine.cuh"

namespace ftsb {

// compile-time loops (static register indexing by construction; `#pragma unroll` is refused by LLVM beyond its
// pragma-unroll size threshold, which would demote the register arrays to local memory)
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// G lanes per instance, strided across the warp
template <int G>
struct LaneTeam {
    static constexpr bool kIsCta = false;
    static constexpr int kSlots = 32 / G; // instances per warp == interleave stride
    int rank, slot;
    unsigned mask;
    __device__ LaneTeam()
    {
        const int lane = threadIdx.x & 31;
        slot = lane % kSlots;
        rank = lane / kSlots;
        mask = 0;
#pragma unroll
        for (int j = 0; j < G; ++j) mask |= 1u << (slot + j * kSlots);
    }
    __device__ int size() const { return G; }
    __device__ void sync() const
    {
        if (G > 1) __syncwarp(mask);
    }
    __device__ void sum2(double &a, double &b) const
    {
#pragma unroll
        for (int o = 16; o >= kSlots; o >>= 1) {
            a += __shfl_xor_sync(mask, a, o);
            b += __shfl_xor_sync(mask, b, o);
        }
    }
    __device__ double sum(double v) const
    {
#pragma unroll
        for (int o = 16; o >= kSlots; o >>= 1) v += __shfl_xor_sync(mask, v, o);
        return v;
    }
    __device__ int sumi(int v) const
    {
#pragma unroll
        for (int o = 16; o >= kSlots; o >>= 1) v += __shfl_xor_sync(mask, v, o);
        return v;
    }
    __device__ void argmax(double &v, int &pos) const
    {
#pragma unroll
        for (int o = 16; o >= kSlots; o >>= 1) {
            const double v2 = __shfl_xor_sync(mask, v, o);
            const int p2 = __shfl_xor_sync(mask, pos, o);
            if (v2 > v || (v2 == v && p2 < pos)) {
                v = v2;
                pos = p2;
            }
        }
    }
    __device__ long long bcast_ll(long long v) const { return G > 1 ? __shfl_sync(mask, v, slot) : v; }
};

// ---------------------------------------------------------------------------------------------------------
// Zero-skipping elimination for the thread-per-instance kernels (G = 1).
//
// MNA matrices are sparse (C2: 40 structural non-zeros of 196; 819 dense element updates per factorisation, 49 of
// them on non-zeros).  The reference's dense elimination (gauss_lu.rs:29-42) spends the rest on `a - 0*p`, which
// leaves `a` unchanged, so skipping those updates reproduces its result (up to the sign of a zero) as long as every
// value involved is finite.  The structure is tracked at run time, per instance, because partial pivoting is value
// dependent: one 32-bit word per index i = rowmask(i) | colmask(i) << 16 (which columns row i holds; which rows
// hold column i).  J stays in its dense column-major slots, but entries outside the masks are never read, so the
// matrix is not zero-filled before assembly.  Same pivot rule (first position with the strictly largest |a|), same
// reciprocal scaling, same separate multiply / subtract, same update order as the dense path.
// A factorisation that meets a zero, infinite or NaN pivot candidate, or a non-finite solution component, returns
// false: the caller then repeats it with the dense arithmetic (0*inf = NaN poisoning and all), so the observable
// behaviour in those cases is the reference's too.
// ---------------------------------------------------------------------------------------------------------
template <int NC>
struct SparseLU {
    static_assert(NC <= 16, "masks are 16 bits wide");
    static constexpr int TS = 32;
    static constexpr unsigned FULL = (1u << NC) - 1u;
    static __device__ __forceinline__ int nib_get(unsigned long long p, int i) { return (int)((p >> (4 * i)) & 15ull); }
    static __device__ __forceinline__ unsigned long long nib_set(unsigned long long p, int i, int r)
    {
        return (p & ~(15ull << (4 * i))) | ((unsigned long long)r << (4 * i));
    }
    static __device__ __forceinline__ double &at(double *Jp, int r, int j) { return Jp[(j * NC + r) * TS]; }

    // mk[i], i < NC: structure words (reset from the static pattern mk0); mk[NC + k]: rows eliminated at step k
    // (low half) and rows of U holding column k (high half), for the substitutions
    static __device__ __forceinline__ bool factor(double *__restrict__ Jp, double *__restrict__ yp, int *__restrict__ mk,
                                                  const int *__restrict__ mk0, unsigned long long &perm)
    {
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
#pragma unroll
        for (int i = 0; i < NC; ++i) mk[i * TS] = mk0[i];
        perm = 0xfedcba9876543210ull;                    // position -> row
        unsigned long long posn = 0xfedcba9876543210ull; // row -> position
        unsigned active = FULL;                          // rows not yet chosen as a pivot
        bool ok = true;
        for (int k = 0; k < NC; ++k) {
            const unsigned colk = ((unsigned)mk[k * TS] >> 16) & FULL;
            const unsigned cand = colk & active;
            // pivot: first position with the strictly largest |a| (gauss_lu.rs:6-16); structural zeros cannot win
            double vmax = -1.0;
            int P = nib_get(perm, k), ppos = k;
            for (unsigned c = cand; c; c &= c - 1u) {
                const int r = __ffs((int)c) - 1;
                const double v = fabs(at(Jp, r, k));
                const int p = nib_get(posn, r);
                ok = ok && (v < INF);
                if (v > vmax || (v == vmax && p < ppos)) {
                    vmax = v;
                    P = r;
                    ppos = p;
                }
            }
            ok = ok && (vmax > 0.0);
            if (!ok) break;
            const int rk = nib_get(perm, k); // row swap (:19-27)
            perm = nib_set(nib_set(perm, ppos, rk), k, P);
            posn = nib_set(nib_set(posn, rk, ppos), P, k);
            const unsigned done = FULL & ~active;
            active &= ~(1u << P);
            const double alpha = 1.0 / at(Jp, P, k); // (:31)
            const double yk = yp[P * TS];
            const unsigned cols = (unsigned)mk[P * TS] & FULL & ~((2u << k) - 1u); // pivot row, columns > k
            unsigned lrows = 0;
            for (unsigned c = cand & active; c; c &= c - 1u) {
                const int r = __ffs((int)c) - 1;
                const double a = at(Jp, r, k);
                const double m = alpha * a;
                at(Jp, r, k) = m;
                if (a != 0.0) {
                    lrows |= 1u << r;
                    yp[r * TS] = yp[r * TS] - m * yk; // (:40)
                    const unsigned w = (unsigned)mk[r * TS];
                    mk[r * TS] = (int)(w | cols);
                    for (unsigned cc = cols; cc; cc &= cc - 1u) {
                        const int j = __ffs((int)cc) - 1;
                        const double t = m * at(Jp, P, j);
                        if ((w >> j) & 1u) {
                            at(Jp, r, j) = at(Jp, r, j) - t; // (:35-39)
                        } else { // fill-in: the dense entry was +0.0
                            at(Jp, r, j) = 0.0 - t;
                            mk[j * TS] = (int)((unsigned)mk[j * TS] | (1u << (16 + r)));
                        }
                    }
                }
            }
            mk[(NC + k) * TS] = (int)(lrows | ((colk & done) << 16));
        }
        return ok;
    }

    // new right-hand side through the stored multipliers
    static __device__ __forceinline__ void forward(const double *__restrict__ Jp, double *__restrict__ yp,
                                                   const int *__restrict__ mk, unsigned long long perm)
    {
        for (int k = 0; k < NC - 1; ++k) {
            const double yk = yp[nib_get(perm, k) * TS];
            for (unsigned c = (unsigned)mk[(NC + k) * TS] & FULL; c; c &= c - 1u) {
                const int r = __ffs((int)c) - 1;
                yp[r * TS] = yp[r * TS] - Jp[(k * NC + r) * TS] * yk;
            }
        }
    }

    // gauss_lu.rs:44-53; false when a solution component is not finite
    static __device__ __forceinline__ bool backward(const double *__restrict__ Jp, double *__restrict__ yp,
                                                    double *__restrict__ xp, const int *__restrict__ mk,
                                                    unsigned long long perm)
    {
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        bool ok = true;
        for (int i = NC - 1; i >= 0; --i) {
            const int ri = nib_get(perm, i);
            const double xi = yp[ri * TS] / Jp[(i * NC + ri) * TS];
            xp[i * TS] = xi;
            ok = ok && (fabs(xi) < INF);
            for (unsigned c = ((unsigned)mk[(NC + i) * TS] >> 16) & FULL; c; c &= c - 1u) {
                const int r = __ffs((int)c) - 1;
                yp[r * TS] = yp[r * TS] - xi * Jp[(i * NC + r) * TS];
            }
        }
        return ok;
    }
};

// Dense elimination with run-time loops (compact code): the exact-arithmetic path the zero-skipping solver falls
// back to.  Same operations in the same order as LaneSolver::factor / forward / backward with G = 1.
template <int NC>
struct DenseDyn {
    static constexpr int TS = 32;
    static __device__ __forceinline__ int nib_get(unsigned long long p, int i) { return (int)((p >> (4 * i)) & 15ull); }
    static __device__ __forceinline__ unsigned long long nib_set(unsigned long long p, int i, int r)
    {
        return (p & ~(15ull << (4 * i))) | ((unsigned long long)r << (4 * i));
    }
    static __device__ __noinline__ void factor(double *Jp, double *yp, unsigned long long &perm_out)
    {
        unsigned long long perm = 0xfedcba9876543210ull;
        for (int k = 0; k < NC; ++k) {
            int imax = k;
            double vmax = fabs(Jp[(k * NC + nib_get(perm, k)) * TS]);
            for (int i = k + 1; i < NC; ++i) {
                const double v = fabs(Jp[(k * NC + nib_get(perm, i)) * TS]);
                if (v > vmax) {
                    vmax = v;
                    imax = i;
                }
            }
            const int rk = nib_get(perm, k), rp = nib_get(perm, imax);
            perm = nib_set(nib_set(perm, imax, rk), k, rp);
            const double alpha = 1.0 / Jp[(k * NC + rp) * TS];
            const double yk = yp[rp * TS];
            for (int i = k + 1; i < NC; ++i) {
                const int r = nib_get(perm, i);
                const double m = alpha * Jp[(k * NC + r) * TS];
                Jp[(k * NC + r) * TS] = m;
                for (int j = k + 1; j < NC; ++j) Jp[(j * NC + r) * TS] = Jp[(j * NC + r) * TS] - m * Jp[(j * NC + rp) * TS];
                yp[r * TS] = yp[r * TS] - m * yk;
            }
        }
        perm_out = perm;
    }
    static __device__ __noinline__ void forward(const double *Jp, double *yp, unsigned long long perm)
    {
        for (int k = 0; k < NC - 1; ++k) {
            const double yk = yp[nib_get(perm, k) * TS];
            for (int i = k + 1; i < NC; ++i) {
                const int r = nib_get(perm, i);
                yp[r * TS] = yp[r * TS] - Jp[(k * NC + r) * TS] * yk;
            }
        }
    }
    static __device__ __noinline__ void backward(const double *Jp, double *yp, double *xp, unsigned long long perm)
    {
        for (int i = NC - 1; i >= 0; --i) {
            const int ri = nib_get(perm, i);
            const double xi = yp[ri * TS] / Jp[(i * NC + ri) * TS];
            xp[i * TS] = xi;
            for (int j = i - 1; j >= 0; --j) {
                const int rj = nib_get(perm, j);
                yp[rj * TS] = yp[rj * TS] - xi * Jp[(i * NC + rj) * TS];
            }
        }
    }
};

// N = the analysis mode's unknown count, known at compile time (= leading dimension of J): the statically unrolled
// elimination has no run-time guards (a run-time `j < n` inside the unrolled code makes the compiler track the
// validity of every cached pivot-row element with dozens of predicate instructions per load).
// SP: zero-skipping elimination (SparseLU above; G = 1 only).
template <int G, int NC, bool SP = false>
struct LaneSolver {
    static_assert(!SP || G == 1, "the zero-skipping solver is thread-per-instance");
    using Team = LaneTeam<G>;
    static constexpr int TS = 32 / G; // element stride inside the warp's region
    static constexpr bool kHasA = false;
    static constexpr bool kHasB = false;
    static constexpr bool kExternalStartup = true; // .tran takes its start-up solution from a previous launch
    struct State {
        bool jvalid;             // J holds the factorisation of the current matrix (linear circuits only)
        bool dense;              // SP: that factorisation was made by the dense fall-back (masks are not valid)
        unsigned long long perm; // row at position i in nibble i (n <= 16); identical in all lanes of the team
        __device__ State() : jvalid(false), dense(false), perm(0) {}
    };
    template <class W>
    static __device__ __forceinline__ void matrix_changed(const Team &, const ModeProg &, const W &, State &s)
    {
        s.jvalid = false;
    }

    static __device__ __forceinline__ int perm_get(unsigned long long p, int i) { return (int)((p >> (4 * i)) & 15ull); }
    static __device__ __forceinline__ unsigned long long perm_set(unsigned long long p, int i, int r)
    {
        return (p & ~(15ull << (4 * i))) | ((unsigned long long)r << (4 * i));
    }
    // element (row r, column j) of the column-major J with leading dimension NC
    static __device__ __forceinline__ double &at(double *Jp, int r, int j) { return Jp[(j * NC + r) * TS]; }

    // gauss_lu.rs:3-42 with the rhs carried along; rows are addressed through the position -> row map and
    // distributed over the team's lanes by position
    template <class TeamT>
    static __device__ __forceinline__ void factor(const TeamT &tm, double *Jp, double *yp, unsigned long long &perm)
    {
        constexpr int n = NC;
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        perm = 0xfedcba9876543210ull;
        static_for<0, NC>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            tm.sync(); // rows written by the other lanes in step k-1
            // pivot: first position with the strictly largest |a| (:6-16)
            int imax = INT_MAX;
            double vmax = -1.0;
            for (int i = k + tm.rank; i < n; i += G) {
                double v = fabs(at(Jp, perm_get(perm, i), k));
                if (G > 1 && v != v) v = (i == k) ? INF : -0.5; // a NaN wins only at position k (:8,12)
                if (imax == INT_MAX || v > vmax) {
                    vmax = v;
                    imax = i;
                }
            }
            tm.argmax(vmax, imax);
            const int rk = perm_get(perm, k);
            const int rp = perm_get(perm, imax);
            perm = perm_set(perm_set(perm, imax, rk), k, rp); // row swap (:19-27)
            double prow[NC];
            static_for<k, NC>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                prow[j] = at(Jp, rp, j);
            });
            const double yk = yp[rp * TS];
            const double alpha = 1.0 / prow[k]; // (:31)
            for (int i = k + 1 + tm.rank; i < n; i += G) {
                const int r = perm_get(perm, i);
                const double m = alpha * at(Jp, r, k);
                at(Jp, r, k) = m;
                static_for<k + 1, NC>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    at(Jp, r, j) = at(Jp, r, j) - m * prow[j]; // (:35-39)
                });
                yp[r * TS] = yp[r * TS] - m * yk; // (:40)
            }
        });
        tm.sync();
    }

    // new right-hand side through the stored multipliers (bit-identical to carrying it through factor())
    template <class TeamT>
    static __device__ __forceinline__ void forward(const TeamT &tm, const double *Jp, double *yp, unsigned long long perm)
    {
        constexpr int n = NC;
        for (int k = 0; k < n - 1; ++k) {
            tm.sync();
            const double yk = yp[perm_get(perm, k) * TS];
            for (int i = k + 1 + tm.rank; i < n; i += G) {
                const int r = perm_get(perm, i);
                yp[r * TS] = yp[r * TS] - Jp[(k * NC + r) * TS] * yk;
            }
        }
        tm.sync();
    }

    // gauss_lu.rs:44-53
    template <class TeamT>
    static __device__ __forceinline__ void backward(const TeamT &tm, const double *Jp, double *yp, double *xp,
                                                    unsigned long long perm)
    {
        constexpr int n = NC;
        for (int i = n - 1; i >= 0; --i) {
            const int ri = perm_get(perm, i);
            const double xi = yp[ri * TS] / Jp[(i * NC + ri) * TS];
            if (tm.rank == 0) xp[i * TS] = xi;
            for (int j = i - 1 - tm.rank; j >= 0; j -= G) {
                const int rj = perm_get(perm, j);
                yp[rj * TS] = yp[rj * TS] - xi * Jp[(i * NC + rj) * TS];
            }
            tm.sync();
        }
    }

    // f = A x + H g - b and its class norms, from the flat maps (no A, no b in memory)
    template <class TeamT, class W, class V>
    static __device__ __forceinline__ void resid(const TeamT &tm, const ModeProg &M, const W &w, V x, int nbad, double &nv,
                                                 double &ni)
    {
        const int n = M.n;
        double sv = 0.0, si = 0.0;
        for (int row = tm.rank; row < n; row += G) {
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

    // r = b + nonlinear rhs, reference accumulation order (row-indexed maps)
    template <class TeamT, class W>
    static __device__ __forceinline__ void assemble_rhs(const TeamT &tm, const ModeProg &M, const W &w, double *__restrict__ yp)
    {
        for (int row = tm.rank; row < NC; row += G) {
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
    }

    // SP: the zero-skipping factorisation met a zero / non-finite value: redo this solve with the dense arithmetic
    template <class TeamT, class W>
    static __device__ __noinline__ void dense_fallback(const TeamT &tm, const ModeProg &M, const W &w, State &s)
    {
        double *Jp = w.J.p, *yp = w.r.p, *xpp = w.xp.p;
        assemble_rhs(tm, M, w, yp);
        for (int e = 0; e < NC * NC; ++e) Jp[e * TS] = 0.0;
        gather(tm, M.J_full, w.J, decltype(w.J){nullptr}, w.val, NC);
        DenseDyn<NC>::factor(Jp, yp, s.perm);
        DenseDyn<NC>::backward(Jp, yp, xpp, s.perm);
        s.dense = true;
    }

    template <class W>
    static __device__ __forceinline__ int newton(const Team &tm, const ModeProg &M, const W &w, State &s, Counters &cn,
                                                 int skip)
    {
        constexpr int n = NC; // == M.n
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        const bool linear = (M.n_nl == 0);
        double *__restrict__ Jp = w.J.p;
        double *__restrict__ yp = w.r.p;
        double *__restrict__ xpp = w.xp.p;

        int flags = eval_nl(tm, M, w.x, w.val);
        if (flags >> 20) return -ST_NMOS_UNREACHABLE;
        double eov, eoi;
        resid(tm, M, w, w.x, flags & 0xfffff, eov, eoi);
        double sov = eov, soi = eoi;
        double ev = INF, ei = INF, sv = INF, si = INF;
        bool xp_ready = false;
        int it = 0;
        while (it < 100 && !(sv < 0.001 * sov + 1e-6 && si < 0.001 * soi + 1e-9 && ev < 0.001 * eov + 1e-6 &&
                             ei < 0.001 * eoi + 1e-9)) {
            const bool do_factor = !(linear && skip && s.jvalid);
            const bool do_solve = !(linear && skip && xp_ready);
            if (do_solve) {
                assemble_rhs(tm, M, w, yp);
                if constexpr (SP) {
                    int *__restrict__ mk = w.piv.p;
                    bool ok = true;
                    if (do_factor) {
                        gather(tm, M.J_full, w.J, decltype(w.J){nullptr}, w.val, NC);
                        ok = SparseLU<NC>::factor(Jp, yp, mk, w.aux, s.perm);
                        s.dense = false;
                        s.jvalid = linear;
                        cn.c[CN_FACTOR]++;
                    } else if (s.dense) {
                        DenseDyn<NC>::forward(Jp, yp, s.perm);
                    } else {
                        SparseLU<NC>::forward(Jp, yp, mk, s.perm);
                    }
                    if (ok) {
                        if (s.dense)
                            DenseDyn<NC>::backward(Jp, yp, xpp, s.perm);
                        else
                            ok = SparseLU<NC>::backward(Jp, yp, xpp, mk, s.perm);
                    }
                    if (!ok) dense_fallback(tm, M, w, s);
                } else {
                    if (do_factor) {
                        for (int c = 0; c < n; ++c)
                            for (int r = tm.rank; r < n; r += G) Jp[(c * NC + r) * TS] = 0.0;
                        tm.sync();
                        gather(tm, M.J_full, w.J, decltype(w.J){nullptr}, w.val, NC);
                        factor(tm, Jp, yp, s.perm);
                        s.jvalid = linear;
                        cn.c[CN_FACTOR]++;
                    } else {
                        forward(tm, Jp, yp, s.perm);
                    }
                    backward(tm, Jp, yp, xpp, s.perm);
                }
                xp_ready = true;
                cn.c[CN_SOLVE]++;
            }
            // damped step (newtons_method.rs:73-77) and its norms
            double qv = 0.0, qi = 0.0;
            for (int i = tm.rank; i < n; i += G) {
                const double xi = w.x[i];
                const double d = xpp[i * TS] - xi;
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
            tm.sync();
            for (int i = tm.rank; i < n; i += G) w.x[i] = w.xn[i];
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

// per-instance layout of the lane kernels, in doubles (each element is strided by 32/G inside the warp's region)
struct LaneLayout {
    int oJ, or_, ox, oxn, oxp, oval, ostate, oring, oprev, omk, doubles;
};

__host__ __device__ inline LaneLayout lane_layout(int nc, int n_slots, int n_dyn, int n_src, bool tran, bool sparse = false)
{
    LaneLayout t;
    int o = 0;
    t.oJ = o;
    o += nc * nc;
    t.or_ = o;
    o += nc;
    t.ox = o;
    o += nc;
    t.oxn = o;
    o += nc;
    t.oxp = o;
    o += nc;
    t.oval = o;
    o += n_slots;
    t.oprev = o; // source history: only .tran reads it
    o += tran ? 3 * n_src : 0;
    t.ostate = o;
    o += tran ? 2 * n_dyn : 0;
    t.oring = o;
    o += tran ? 3 * nc : 0;
    t.omk = o; // structure words of the zero-skipping solver: 2 * nc ints
    o += sparse ? nc : 0;
    t.doubles = o;
    return t;
}

// static structure words of J_full (rowmask(i) | colmask(i) << 16), one per index, computed once per block
template <int NC>
__device__ __forceinline__ void static_structure(const Gather &g, int *mk0)
{
    const int i = threadIdx.x & 31;
    if (i < NC) {
        unsigned rm = 0, cm = 0;
        for (int d = 0; d < g.n_dst; ++d) {
            const int off = __ldg(&g.dst[d]);
            const int row = off & 0xffff, col = off >> 16;
            if (row == i) rm |= 1u << col;
            if (col == i) cm |= 1u << row;
        }
        mk0[i] = (int)(rm | (cm << 16));
    }
    __syncwarp();
}

// .tran for small circuits: starts from the start-up solution a.x_start left by a k_sync<.., AN_OP> launch
template <int G, int NC, int AN, bool SP = false>
__global__ void __launch_bounds__(32) k_lanes(const __grid_constant__ RunArgs a)
{
    static_assert(AN == AN_TRAN, "k_lanes is the .tran kernel; .op and .dc use k_sync");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int mk0[16];
    constexpr int TS = 32 / G;
    const DevProg &P = a.prog;
    const LaneLayout T = lane_layout(NC, P.n_slots, P.n_dyn, P.n_src, AN == AN_TRAN, SP);
    if constexpr (SP) static_structure<NC>(P.tran.J_full, mk0);
    LaneTeam<G> tm;
    double *base = reinterpret_cast<double *>(smem_raw) + tm.slot;
    WsT<TS> w;
    w.ld = NC;
    w.A.p = nullptr;
    w.b.p = nullptr;
    w.J.p = base + T.oJ * TS;
    w.r.p = base + T.or_ * TS;
    w.x.p = base + T.ox * TS;
    w.xn.p = base + T.oxn * TS;
    w.xp.p = base + T.oxp * TS;
    w.val.p = base + T.oval * TS;
    w.state.p = base + T.ostate * TS;
    w.ring.p = base + T.oring * TS;
    w.prev.p = AN == AN_TRAN ? base + T.oprev * TS : nullptr;
    // the structure words are ints interleaved like the doubles (word e of slot s at e * TS + s)
    w.piv.p = SP ? reinterpret_cast<int *>(reinterpret_cast<double *>(smem_raw) + T.omk * TS) + tm.slot : nullptr;
    w.aux = mk0;
    Counters cn;
    const int lane = threadIdx.x & 31;
    for (;;) {
        long long first = 0;
        if (lane == 0) first = (long long)atomicAdd(a.work, (unsigned long long)TS);
        first = __shfl_sync(0xffffffffu, first, 0);
        if (first >= a.batch) break;
        const long long inst = first + tm.slot;
        if (inst < a.batch) run_instance<LaneTeam<G>, LaneSolver<G, NC, SP>, AN>(tm, a, w, inst, cn);
        __syncwarp();
    }
    // one atomic per counter per warp (counters are kept by rank 0 of each team)
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        unsigned long long v = tm.rank == 0 ? cn.c[i] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&a.counters[i], v);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Warp-synchronous variant for .op and .dc: every team of the warp executes ONE Newton iteration per round, in lock
// step, so all shuffles / barriers use the constant full mask (a run-time sub-warp mask costs a MATCH/REDUX/VOTE
// preamble per collective), there is a single iteration body (small instruction footprint), and a team whose
// instance is finished fetches the next one at once instead of idling until the slowest instance of the warp
// converges ("converged lanes retire", per-instance masking).
//
// Round = C (assemble, LU, damped step -> x_new) ; A (device evaluation + residual at x_new) ; B (accept, count,
// decide).  The reference's initial get_err(x) (newtons_method.rs:28-30) only seeds err_old/step_old, which are
// overwritten before they are ever compared (:62-63, quirk Q1), so a solve needs no initial residual: only device
// values at the current x, which the previous round's A (or the instance set-up) has left in the slot table.
// A team without a pending solve still executes C/A/B on its own (garbage) data: nothing it owns is observable.
// ---------------------------------------------------------------------------------------------------------
template <int G>
struct SyncTeam { // same lanes as LaneTeam<G>, collectives on the whole (converged) warp
    static constexpr bool kIsCta = false;
    static constexpr int kSlots = 32 / G;
    int rank, slot;
    __device__ SyncTeam()
    {
        const int lane = threadIdx.x & 31;
        slot = lane % kSlots;
        rank = lane / kSlots;
    }
    __device__ int size() const { return G; }
    __device__ void sync() const
    {
        if (G > 1) __syncwarp();
    }
    __device__ void sum2(double &a, double &b) const
    {
#pragma unroll
        for (int o = 16; o >= kSlots; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
    }
    __device__ int sumi(int v) const
    {
#pragma unroll
        for (int o = 16; o >= kSlots; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ void argmax(double &v, int &pos) const
    {
#pragma unroll
        for (int o = 16; o >= kSlots; o >>= 1) {
            const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
            const int p2 = __shfl_xor_sync(0xffffffffu, pos, o);
            if (v2 > v || (v2 == v && p2 < pos)) {
                v = v2;
                pos = p2;
            }
        }
    }
};

enum : int { PH_FETCH = 0, PH_SOLVE = 1, PH_IDLE = 2 };

template <int G, int NC, int AN, bool SP = false>
__global__ void __launch_bounds__(32) k_sync(const __grid_constant__ RunArgs a)
{
    static_assert(AN == AN_OP || AN == AN_DC, "the warp-synchronous kernel covers .op and .dc");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int mk0[16];
    constexpr int TS = 32 / G;
    using Solver = LaneSolver<G, NC, SP>;
    using LSync = LaneSolver<G, NC, SP>; // factor/forward/backward/resid are templated on the team object below
    const DevProg &P = a.prog;
    const ModeProg &M = AN == AN_OP ? P.op : P.dc;
    constexpr int n = NC; // == M.n (the host picks the instantiation)
    const bool linear = (M.n_nl == 0);
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const LaneLayout T = lane_layout(NC, P.n_slots, P.n_dyn, P.n_src, false, SP);
    if constexpr (SP) static_structure<NC>(M.J_full, mk0);
    LaneTeam<G> tl;  // run-time mask: used only in the (divergent, rare) transition code
    SyncTeam<G> ts;  // constant full mask: used in the lock-step body
    double *base = reinterpret_cast<double *>(smem_raw) + tl.slot;
    WsT<TS> w;
    w.ld = NC;
    w.A.p = nullptr;
    w.b.p = nullptr;
    w.J.p = base + T.oJ * TS;
    w.r.p = base + T.or_ * TS;
    w.x.p = base + T.ox * TS;
    w.xn.p = base + T.oxn * TS;
    w.xp.p = base + T.oxp * TS;
    w.val.p = base + T.oval * TS;
    w.state.p = nullptr;
    w.ring.p = nullptr;
    w.prev.p = nullptr;
    w.piv.p = SP ? reinterpret_cast<int *>(reinterpret_cast<double *>(smem_raw) + T.omk * TS) + tl.slot : nullptr;
    w.aux = mk0;
    int *__restrict__ mk = w.piv.p;
    if constexpr (SP) { // lanes that never get an instance still run the lock-step body: give them an empty structure
        for (int i = 0; i < 2 * NC; ++i) mk[i * TS] = 0;
    }
    double *__restrict__ Jp = w.J.p;
    double *__restrict__ yp = w.r.p;
    double *__restrict__ xpp = w.xp.p;
    Counters cn;
    const int lane = threadIdx.x & 31;

    // per-team state (identical in the lanes of a team)
    int phase = PH_FETCH;
    long long inst = -1;
    typename Solver::State s;
    bool xp_ready = false;
    int it = 0;
    double sv = INF, si = INF, ev = INF, ei = INF, sov = INF, soi = INF, eov = INF, eoi = INF;
    // .dc chain state
    long long k = 0, count = 0;
    double dc_start = 0.0, dc_step = 0.0;
    const int sweep_slot = AN == AN_DC ? __ldg(&P.src[5 * a.sweep_src + 1]) : 0;
    int tcnt = 0; // factorisations of the current instance recorded in the pivot trace (pilot / redo runs only)
    long long tidx = 0;

    for (;;) {
        // ---------------- transitions ----------------
        {
            const unsigned need = __ballot_sync(0xffffffffu, phase == PH_FETCH && tl.rank == 0);
            if (need) { // one atomic per warp for all teams that need an instance
                long long first = 0;
                if (lane == 0) first = (long long)atomicAdd(a.work, (unsigned long long)__popc(need));
                first = __shfl_sync(0xffffffffu, first, 0);
                // my team's index among the needy ones (slot bits live in lanes 0..kSlots-1)
                const unsigned below = need & ((1u << tl.slot) - 1u);
                if (phase == PH_FETCH) {
                    inst = first + __popc(below);
                    tidx = inst; // trace row: the position in the work list
                    if (a.in_list) inst = (unsigned long long)inst < *a.in_count ? a.in_list[inst] : a.batch; // list mode
                    tcnt = 0;
                }
            }
        }
        if (phase == PH_FETCH) {
            if (inst < a.batch) {
                load_instance(tl, a, w, inst);
                for (int i = tl.rank; i < n; i += G) {
                    w.x[i] = 0.0;  // x0 = 0 (mna.rs:21-23)
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
                s.jvalid = false;
                xp_ready = false;
                it = 0;
                sv = si = ev = ei = sov = soi = eov = eoi = INF;
                phase = PH_SOLVE;
                if (AN == AN_DC && count == 0) { // empty sweep: nothing to solve
                    if (tl.rank == 0) {
                        a.points_done[inst] = 0;
                        a.status[inst] = ST_OK;
                    }
                    phase = PH_FETCH;
                } else if (fl >> 20) { // cannot happen at x0 = 0; kept for symmetry with the reference's panic
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
            continue; // some team still has to fetch (its previous fetch produced an empty chain)
        }

        // ---------------- C: assemble, solve, damped step ----------------
        const bool solving = phase == PH_SOLVE;
        const bool do_factor = !(linear && a.skip_refactor && s.jvalid);
        const bool do_solve = !(linear && a.skip_refactor && xp_ready);
        const bool any_solve = __any_sync(0xffffffffu, solving && do_solve);
        const bool any_factor = __any_sync(0xffffffffu, solving && do_factor);
        if (any_solve) {
            LSync::assemble_rhs(ts, M, w, yp);
            if constexpr (SP) {
                bool ok = true;
                if (any_factor) {
                    // a team whose cached factors are still valid re-creates exactly the same factors: harmless
                    gather(ts, M.J_full, w.J, decltype(w.J){nullptr}, w.val, NC);
                    ok = SparseLU<NC>::factor(Jp, yp, mk, mk0, s.perm);
                    s.dense = false;
                    s.jvalid = linear;
                    if (solving) cn.c[CN_FACTOR]++;
                } else if (s.dense) {
                    DenseDyn<NC>::forward(Jp, yp, s.perm);
                } else {
                    SparseLU<NC>::forward(Jp, yp, mk, s.perm);
                }
                if (ok) {
                    if (s.dense)
                        DenseDyn<NC>::backward(Jp, yp, xpp, s.perm);
                    else
                        ok = SparseLU<NC>::backward(Jp, yp, xpp, mk, s.perm);
                }
                if (solving && !ok) LSync::dense_fallback(ts, M, w, s); // rare: zero / non-finite values
            } else {
                if (any_factor) {
                    // a team whose cached factors are still valid re-creates exactly the same factors: harmless
                    for (int c = 0; c < n; ++c)
                        for (int r = ts.rank; r < n; r += G) Jp[(c * NC + r) * TS] = 0.0;
                    ts.sync();
                    gather(ts, M.J_full, w.J, decltype(w.J){nullptr}, w.val, NC);
                    LSync::factor(ts, Jp, yp, s.perm);
                    s.jvalid = linear;
                    if (solving) cn.c[CN_FACTOR]++;
                    if (a.trace && solving && do_factor && ts.rank == 0 && tidx < a.trace_rows) // remember the pivot sequence
                        a.trace[tidx * PLAN_TRACE_MAX + (tcnt++ & (PLAN_TRACE_MAX - 1))] = s.perm;
                } else {
                    LSync::forward(ts, Jp, yp, s.perm);
                }
                LSync::backward(ts, Jp, yp, xpp, s.perm);
            }
            xp_ready = true;
            if (solving) cn.c[CN_SOLVE]++;
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
        LSync::resid(ts, M, w, w.xn, flags & 0xfffff, ev, ei);

        // ---------------- B: accept, count, decide (per team, no collectives) ----------------
        if (solving) {
            int fin = -1; // -1: keep iterating
            if (flags >> 20)
                fin = ST_NMOS_UNREACHABLE;
            else if (isinf(ev) || isinf(ei))
                fin = ST_DIVERGED;
            else {
                for (int i = ts.rank; i < n; i += G) w.x[i] = w.xn[i];
                ++it;
                cn.c[CN_NEWTON]++;
                eov = ev;
                eoi = ei;
                sov = sv;
                soi = si;
                const bool conv = sv < 0.001 * sov + 1e-6 && si < 0.001 * soi + 1e-9 && ev < 0.001 * eov + 1e-6 &&
                                  ei < 0.001 * eoi + 1e-9;
                if (!(it < 100 && !conv)) fin = it < 100 ? ST_OK : ST_NOT_CONVERGED;
            }
            if (fin >= 0) {
                if (AN == AN_OP) {
                    for (int i = ts.rank; i < n; i += G) a.x[inst * n + i] = w.x[i];
                    if (ts.rank == 0) {
                        a.n_iters[inst] = fin == ST_OK ? it : 0;
                        a.status[inst] = fin;
                    }
                    if (fin == ST_OK) cn.c[CN_POINTS]++;
                    phase = PH_FETCH;
                } else {
                    if (fin == ST_OK) {
                        double *xo = a.x + (inst * a.max_points + k) * n;
                        for (int i = ts.rank; i < n; i += G) xo[i] = w.x[i];
                        if (ts.rank == 0) a.n_iters[inst * a.max_points + k] = it;
                        cn.c[CN_POINTS]++;
                        ++k;
                    }
                    if (fin != ST_OK || k >= count) { // `?` aborts the sweep at the first failure (engine.rs:127)
                        if (ts.rank == 0) {
                            a.points_done[inst] = k;
                            a.status[inst] = fin;
                        }
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
    for (int i = 0; i < 6; ++i) {
        unsigned long long v = tl.rank == 0 ? cn.c[i] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&a.counters[i], v);
    }
}

typedef void (*KernelFn)(const RunArgs);
// defined in kern_lanes_*.cu (scripts/gen_lanes.py); nullptr when (g, nc, sp) is not instantiated there
KernelFn lanes_kernel_a(int g, int nc, int an, int sp);
KernelFn lanes_kernel_b(int g, int nc, int an, int sp);
KernelFn lanes_kernel_c(int g, int nc, int an, int sp);
KernelFn lanes_kernel_d(int g, int nc, int an, int sp);
KernelFn lanes_kernel_e(int g, int nc, int an, int sp);
KernelFn lanes_kernel_s1(int g, int nc, int an, int sp);
KernelFn lanes_kernel_s2(int g, int nc, int an, int sp);
KernelFn lanes_kernel_s3(int g, int nc, int an, int sp);
KernelFn lanes_kernel_s4(int g, int nc, int an, int sp);

} // namespace ftsb
