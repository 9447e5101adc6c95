/*
 * ftsp_oracle.c — CPU ORACLE (test infrastructure, NOT product code; see ftsp_oracle.h).
 *
 * Plain-C restatement of ftspice's `src/engine` path, written to follow the
 * reference statement by statement so that the iteration PATH (not only the
 * root) is reproduced: same x0, same stamping order, same dense LU with the
 * same pivot rule, same damping, same stopping rule, same stamp/un-stamp
 * sequences in .dc and .tran (quirks Q1-Q22 of SURVEY.md §8a).
 *
 * Build: gcc -O2 -ffp-contract=off (rustc never contracts a*b+c into an FMA).
 * Third-party arithmetic restated from its published source: ndarray 0.15
 * (`numeric_util::unrolled_dot` / `unrolled_fold`: eight partial accumulators,
 * combined as (p0+p4)+(p1+p5)+(p2+p6)+(p3+p7), then the <8 tail; and
 * `linspace::range`: len = ceil((stop-start)/step), item k = start + step*k).
 * ndarray's source is NOT under /root/reference (Cargo.toml:8-11) — restated
 * from memory of the 0.15 sources; it only decides last-bit summation order.
 *
 * PARITY UNPINNED for end-to-end results: the reference has no golden .op/.dc/
 * .tran output and cannot be built here (no cargo).  Pinned: the reference's own
 * unit-test KATs (gauss_lu.rs:60-109, mna.rs:36-52, node_vec_norm.rs:49-77,
 * newtons_method.rs:95-153, res.rs/vdd.rs/idd.rs stamp tests) — tests/test_oracle_kats.py.
 *
 * All `file:line` citations are relative to /root/reference/.
 */
#include "ftsp_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * ndarray 0.15 numeric_util restatements
 * ---------------------------------------------------------------------------------------- */

double orc_unrolled_dot(const double *xs, const double *ys, int len)
{
    double sum = 0.0;
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0, p4 = 0.0, p5 = 0.0, p6 = 0.0, p7 = 0.0;
    while (len >= 8) {
        p0 = p0 + xs[0] * ys[0];
        p1 = p1 + xs[1] * ys[1];
        p2 = p2 + xs[2] * ys[2];
        p3 = p3 + xs[3] * ys[3];
        p4 = p4 + xs[4] * ys[4];
        p5 = p5 + xs[5] * ys[5];
        p6 = p6 + xs[6] * ys[6];
        p7 = p7 + xs[7] * ys[7];
        xs += 8;
        ys += 8;
        len -= 8;
    }
    sum = sum + (p0 + p4);
    sum = sum + (p1 + p5);
    sum = sum + (p2 + p6);
    sum = sum + (p3 + p7);
    for (int i = 0; i < len; i++) {
        if (i >= 7) break;
        sum = sum + xs[i] * ys[i];
    }
    return sum;
}

static double unrolled_sum(const double *xs, int len)
{
    double acc = 0.0;
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0, p4 = 0.0, p5 = 0.0, p6 = 0.0, p7 = 0.0;
    while (len >= 8) {
        p0 = p0 + xs[0];
        p1 = p1 + xs[1];
        p2 = p2 + xs[2];
        p3 = p3 + xs[3];
        p4 = p4 + xs[4];
        p5 = p5 + xs[5];
        p6 = p6 + xs[6];
        p7 = p7 + xs[7];
        xs += 8;
        len -= 8;
    }
    acc = acc + (p0 + p4);
    acc = acc + (p1 + p5);
    acc = acc + (p2 + p6);
    acc = acc + (p3 + p7);
    for (int i = 0; i < len; i++) {
        if (i >= 7) break;
        acc = acc + xs[i];
    }
    return acc;
}

/* ------------------------------------------------------------------------------------------
 * src/engine/gauss_lu.rs:3-54
 * ---------------------------------------------------------------------------------------- */
void orc_lu_solve(int n, double *a, double *b, double *x)
{
    for (int k = 0; k < n; k++) {
        /* pivot: first row with the strictly largest |a[r,k]|  (:6-16) */
        int max_idx = k;
        double max_val = fabs(a[k * n + k]);
        for (int r = k + 1; r < n; r++) {
            double v = fabs(a[r * n + k]);
            if (v > max_val) {
                max_idx = r;
                max_val = v;
            }
        }
        /* full-row swap, multipliers included, and b  (:19-27) */
        if (max_idx != k) {
            for (int j = 0; j < n; j++) {
                double t = a[k * n + j];
                a[k * n + j] = a[max_idx * n + j];
                a[max_idx * n + j] = t;
            }
            double t = b[k];
            b[k] = b[max_idx];
            b[max_idx] = t;
        }
        /* scale under the diagonal by the reciprocal  (:31-32) */
        double alpha = 1.0 / a[k * n + k];
        for (int i = k + 1; i < n; i++) a[i * n + k] = alpha * a[i * n + k];
        /* subtract  (:35-41) */
        for (int i = k + 1; i < n; i++) {
            double m = a[i * n + k];
            for (int j = k + 1; j < n; j++) a[i * n + j] -= m * a[k * n + j];
            b[i] -= m * b[k];
        }
    }
    /* backwards substitution, column oriented  (:44-53) */
    for (int i = 0; i < n; i++) x[i] = b[i];
    for (int i = n - 1; i >= 0; i--) {
        x[i] /= a[i * n + i];
        for (int j = i - 1; j >= 0; j--) x[j] -= x[i] * a[j * n + i];
    }
}

/* ------------------------------------------------------------------------------------------
 * src/engine/node_vec_norm.rs:12-41
 * ---------------------------------------------------------------------------------------- */
double orc_norm(const double *v, int len)
{
    /* v.mapv(|x| x.powi(2)).sum().sqrt() / v.len() as f64 ; empty → 0/0 = NaN (Q3) */
    double stackbuf[64];
    double *sq = len <= 64 ? stackbuf : (double *)malloc(sizeof(double) * (size_t)len);
    for (int i = 0; i < len; i++) sq[i] = v[i] * v[i];
    double s = unrolled_sum(sq, len);
    if (sq != stackbuf) free(sq);
    return sqrt(s) / (double)len;
}

typedef struct {
    double v, i;
} nvn_t;

/* ------------------------------------------------------------------------------------------
 * src/engine/newtons_method.rs:10-17, 73-89
 * ---------------------------------------------------------------------------------------- */
#define MAX_ITERS 100
static const double TOL_REL = 0.001;
static const double TOL_ABS_V = 1e-6;
static const double TOL_ABS_A = 1e-9;
static const double DAMPING_GAMMA = 1.3;
static const double DAMPING_K = 16.0;

static double rust_signum(double x)
{
    if (isnan(x)) return x;
    return copysign(1.0, x);
}

void orc_dampen_step(const double *step, double *out, int n)
{
    for (int i = 0; i < n; i++) {
        double x = step[i];
        out[i] = DAMPING_GAMMA / DAMPING_K * rust_signum(x) * log1p(DAMPING_K * fabs(x));
    }
}

int orc_converged(double err_v, double err_i, double step_v, double step_i, double err_old_v,
                  double err_old_i, double step_old_v, double step_old_i)
{
    return step_v < TOL_REL * step_old_v + TOL_ABS_V && step_i < TOL_REL * step_old_i + TOL_ABS_A &&
           err_v < TOL_REL * err_old_v + TOL_ABS_V && err_i < TOL_REL * err_old_i + TOL_ABS_A;
}

/* ------------------------------------------------------------------------------------------
 * src/spice_fn.rs:8-124
 * ---------------------------------------------------------------------------------------- */
double orc_spice_fn_eval(int kind, const double *p, double t)
{
    switch (kind) {
    case ORC_FN_SIN: { /* :26-28 */
        double offset = p[0], amplitude = p[1], freq = p[2];
        return offset + amplitude * sin(2.0 * 3.14159265358979323846264338327950288 * freq * t);
    }
    case ORC_FN_PULSE: { /* :52-79 */
        double v1 = p[0], v2 = p[1], delay = p[2], t_rise = p[3], t_fall = p[4], pw = p[5], period = p[6];
        double t_norm = fabs(fmod(t, period));
        if (t_norm < delay) {
            return v1;
        } else if (t_norm < delay + t_rise) {
            double frac = (t_norm - delay) / t_rise;
            return v1 + frac * (v2 - v1);
        } else if (t_norm < delay + pw - t_fall) {
            return v2;
        } else if (t_norm < delay + pw) {
            double frac = (t_norm - (delay + pw - t_fall)) / t_fall;
            return v2 + frac * (v1 - v2);
        } else {
            return v1;
        }
    }
    case ORC_FN_EXP: { /* :99-124 */
        double v1 = p[0], v2 = p[1], rd = p[2], rtau = p[3], fd = p[4], ftau = p[5];
        if (t < rd) {
            return v1;
        } else if (t < fd) {
            return v1 + (v1 - v2) * expm1(-(t - rd) / rtau);
        } else {
            return v1 + (v1 - v2) * expm1(-(t - rd) / rtau) + (v2 - v1) * expm1(-(t - fd) / ftau);
        }
    }
    default:
        return 0.0;
    }
}

/* ------------------------------------------------------------------------------------------
 * device models
 * ---------------------------------------------------------------------------------------- */
/* src/device/diode/model.rs:7-22 */
static const double D_ISAT = 1.0e-12, D_ETA = 1.0, D_VT = 26e-3;
static double dio_i(double vp, double vn) { return D_ISAT * expm1((vp - vn) / (D_ETA * D_VT)); }
static double dio_g(double vp, double vn) { return D_ISAT / (D_ETA * D_VT) * exp((vp - vn) / (D_ETA * D_VT)); }
static double dio_ieq(double vp, double vn) { return dio_i(vp, vn) - dio_g(vp, vn) * (vp - vn); }
void orc_diode_model(double vpos, double vneg, double *i, double *g_eq, double *i_eq)
{
    *i = dio_i(vpos, vneg);
    *g_eq = dio_g(vpos, vneg);
    *i_eq = dio_ieq(vpos, vneg);
}

/* src/device/npn/model.rs:8-54 */
static const double Q_IES = 2e-14, Q_ICS = 99e-14, Q_VTE = 26e-3, Q_VTC = 26e-3, Q_AR = 0.02, Q_AF = 0.99;
typedef struct {
    double vc, vb, ve;
} npn_t;
static double q_vbe(const npn_t *q) { return q->vb - q->ve; }
static double q_vbc(const npn_t *q) { return q->vb - q->vc; }
static double q_ie(const npn_t *q)
{
    return -Q_IES * expm1(q_vbe(q) / Q_VTE) + Q_AR * Q_ICS * expm1(q_vbc(q) / Q_VTC);
}
static double q_ic(const npn_t *q)
{
    return Q_AF * Q_IES * expm1(q_vbe(q) / Q_VTE) - Q_ICS * expm1(q_vbc(q) / Q_VTC);
}
static double q_ib(const npn_t *q) { return -(q_ie(q) + q_ic(q)); }
static double q_gee(const npn_t *q) { return Q_IES / Q_VTE * exp(q_vbe(q) / Q_VTE); }
static double q_gec(const npn_t *q) { return Q_AR * Q_ICS / Q_VTC * exp(q_vbc(q) / Q_VTC); }
static double q_gce(const npn_t *q) { return Q_AF * Q_IES / Q_VTE * exp(q_vbe(q) / Q_VTE); }
static double q_gcc(const npn_t *q) { return Q_ICS / Q_VTC * exp(q_vbc(q) / Q_VTC); }
static double q_ie_eq(const npn_t *q) { return q_ie(q) + q_gee(q) * q_vbe(q) - q_gec(q) * q_vbc(q); }
static double q_ic_eq(const npn_t *q) { return q_ic(q) - q_gce(q) * q_vbe(q) + q_gcc(q) * q_vbc(q); }
void orc_npn_model(double vc, double vb, double ve, double out[9])
{
    npn_t q = {vc, vb, ve};
    out[0] = q_ie(&q);
    out[1] = q_ic(&q);
    out[2] = q_ib(&q);
    out[3] = q_gee(&q);
    out[4] = q_gec(&q);
    out[5] = q_gce(&q);
    out[6] = q_gcc(&q);
    out[7] = q_ie_eq(&q);
    out[8] = q_ic_eq(&q);
}

/* src/device/nmos/model.rs:15-79 */
static const double M_BETA = 0.5e-3, M_VT = 0.6, M_LAMBDA = 0.01;
typedef struct {
    double vd, vg, vs;
} nmos_t;
enum { M_CUTOFF, M_LINEAR, M_SAT, M_UNREACHABLE };
static __thread int tl_panic; /* set when the reference would hit unreachable!() (Q19) */

static double m_vgs(const nmos_t *m) { return m->vg - m->vs; }
static double m_vds(const nmos_t *m) { return m->vd - m->vs; }
static int m_state(const nmos_t *m)
{
    double vgs = m_vgs(m), vds = m_vds(m);
    if (vgs <= M_VT) {
        return M_CUTOFF;
    } else if (0.0 <= vds && vds <= vgs - M_VT) {
        return M_LINEAR;
    } else if (0.0 <= vgs - M_VT && vgs - M_VT <= vds) {
        return M_SAT;
    } else {
        tl_panic = ORC_NMOS_UNREACHABLE;
        return M_UNREACHABLE;
    }
}
static double m_id(const nmos_t *m)
{
    switch (m_state(m)) {
    case M_CUTOFF:
        return 0.0;
    case M_LINEAR:
        return M_BETA * ((m_vgs(m) - M_VT) * m_vds(m) - 0.5 * (m_vds(m) * m_vds(m)));
    case M_SAT:
        return 0.5 * M_BETA * (m_vgs(m) - M_VT) * (m_vgs(m) - M_VT) * (1.0 + M_LAMBDA * m_vds(m));
    default:
        return NAN;
    }
}
static double m_gds(const nmos_t *m)
{
    switch (m_state(m)) {
    case M_CUTOFF:
        return 0.0;
    case M_LINEAR:
        return M_BETA * (m_vgs(m) - M_VT - m_vds(m));
    case M_SAT:
        return 0.5 * M_BETA * M_LAMBDA * ((m_vgs(m) - M_VT) * (m_vgs(m) - M_VT));
    default:
        return NAN;
    }
}
static double m_gm(const nmos_t *m)
{
    switch (m_state(m)) {
    case M_CUTOFF:
        return 0.0;
    case M_LINEAR:
        return M_BETA * m_vds(m);
    case M_SAT:
        return M_BETA * (m_vgs(m) - M_VT) * (1.0 + M_LAMBDA * m_vds(m));
    default:
        return NAN;
    }
}
static double m_ieq(const nmos_t *m) { return m_id(m) - m_gds(m) * m_vds(m) - m_gm(m) * m_vgs(m); }
int orc_nmos_model(double vd, double vg, double vs, double out[4])
{
    nmos_t m = {vd, vg, vs};
    tl_panic = 0;
    out[0] = m_id(&m);
    out[1] = m_gds(&m);
    out[2] = m_gm(&m);
    out[3] = m_ieq(&m);
    int p = tl_panic;
    tl_panic = 0;
    return p;
}

/* ------------------------------------------------------------------------------------------
 * element objects (the Rust structs' mutable fields) and the Stamp trait (src/device.rs:20-114)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t kind, node[3], branch, fn_kind;
    double val;
    double fn[7];
    double u_curr, i_curr; /* Cap/Ind state (cap.rs:13-14, ind.rs:13-14) */
} elem_t;

static double xat(const double *x, int idx) { return idx >= 0 ? x[idx] : 0.0; } /* map_or(0.0, |i| x[i]) */

static int has_tran(const elem_t *e) { return (e->kind == ORC_V || e->kind == ORC_I) && e->fn_kind != ORC_FN_NONE; }
static void eval_tran(elem_t *e, double t)
{
    if (has_tran(e)) e->val = orc_spice_fn_eval(e->fn_kind, e->fn, t); /* vdd.rs:40-44, idd.rs:40-44 */
}

/* linear_stamp / undo_linear_stamp; `n` = row length of a */
static void linear_stamp(const elem_t *e, int n, double *a, double *b)
{
    switch (e->kind) {
    case ORC_R: { /* res.rs:34-50 */
        double g = 1.0 / e->val;
        int vneg = e->node[0], vpos = e->node[1];
        if (vneg >= 0) a[vneg * n + vneg] += g;
        if (vpos >= 0) a[vpos * n + vpos] += g;
        if (vpos >= 0 && vneg >= 0) {
            a[vpos * n + vneg] -= g;
            a[vneg * n + vpos] -= g;
        }
        break;
    }
    case ORC_V: { /* vdd.rs:46-64 */
        int vneg = e->node[0], vpos = e->node[1], is = e->branch;
        b[is] = e->val;
        if (vpos >= 0) {
            a[is * n + vpos] += 1.0;
            a[vpos * n + is] += 1.0;
        }
        if (vneg >= 0) {
            a[is * n + vneg] -= 1.0;
            a[vneg * n + is] -= 1.0;
        }
        break;
    }
    case ORC_I: { /* idd.rs:46-57 */
        int vneg = e->node[0], vpos = e->node[1];
        double val = e->val;
        if (vpos >= 0) b[vpos] += val;
        if (vneg >= 0) b[vneg] -= val;
        break;
    }
    default: /* Cap, Ind, Diode, NPN, NMOS: trait default, no-op (device.rs:42) */
        break;
    }
}

static void undo_linear_stamp(const elem_t *e, int n, double *a, double *b)
{
    switch (e->kind) {
    case ORC_R: { /* res.rs:52-68 */
        double g = 1.0 / e->val;
        int vneg = e->node[0], vpos = e->node[1];
        if (vneg >= 0) a[vneg * n + vneg] -= g;
        if (vpos >= 0) a[vpos * n + vpos] -= g;
        if (vpos >= 0 && vneg >= 0) {
            a[vpos * n + vneg] += g;
            a[vneg * n + vpos] += g;
        }
        break;
    }
    case ORC_V: { /* vdd.rs:66-84 */
        int vneg = e->node[0], vpos = e->node[1], is = e->branch;
        b[is] = 0.0;
        if (vpos >= 0) {
            a[is * n + vpos] -= 1.0;
            a[vpos * n + is] -= 1.0;
        }
        if (vneg >= 0) {
            a[is * n + vneg] += 1.0;
            a[vneg * n + is] += 1.0;
        }
        break;
    }
    case ORC_I: { /* idd.rs:59-70 */
        int vneg = e->node[0], vpos = e->node[1];
        double val = e->val;
        if (vpos >= 0) b[vpos] -= val;
        if (vneg >= 0) b[vneg] += val;
        break;
    }
    default:
        break;
    }
}

static void linear_startup_stamp(const elem_t *e, int n, double *a, double *b)
{
    if (e->kind == ORC_L) { /* ind.rs:70-92 */
        int vpos = e->node[0], vneg = e->node[1], is = e->branch;
        b[is] = 0.0;
        if (vpos >= 0) {
            a[is * n + vpos] += 1.0;
            a[vpos * n + is] += 1.0;
        }
        if (vneg >= 0) {
            a[is * n + vneg] -= 1.0;
            a[vneg * n + is] -= 1.0;
        }
    } else {
        linear_stamp(e, n, a, b); /* device.rs:52-59 */
    }
}

void orc_stamp_linear(const orc_elem_t *oe, int n, double *a, double *b, int mode)
{
    elem_t e;
    memset(&e, 0, sizeof e);
    e.kind = oe->kind;
    memcpy(e.node, oe->node, sizeof e.node);
    e.branch = oe->branch;
    e.fn_kind = oe->fn_kind;
    e.val = oe->val;
    if (mode == 0)
        linear_stamp(&e, n, a, b);
    else if (mode == 1)
        undo_linear_stamp(&e, n, a, b);
    else
        linear_startup_stamp(&e, n, a, b);
}

/* cap/model.rs:10-25, ind/model.rs:10-25 */
static double cap_g(const elem_t *e, double h) { return 2.0 * e->val / h; }
static double cap_ieq(const elem_t *e, double h) { return -(e->i_curr + cap_g(e, h) * e->u_curr); }
static double ind_g(const elem_t *e, double h) { return h / (2.0 * e->val); }
static double ind_ieq(const elem_t *e, double h) { return e->i_curr + ind_g(e, h) * e->u_curr; }

/* dynamic_stamp (sign=+1) / undo_dynamic_stamp (sign=-1): cap.rs:68-144, ind.rs:118-194 */
static void dynamic_stamp(const elem_t *e, double h, int n, double *a, double *b, int undo)
{
    int vpos, vneg;
    double g_eq, i_eq;
    if (e->kind == ORC_C) {
        vneg = e->node[0];
        vpos = e->node[1];
        g_eq = cap_g(e, h);
        i_eq = cap_ieq(e, h);
    } else if (e->kind == ORC_L) {
        vpos = e->node[0];
        vneg = e->node[1];
        g_eq = ind_g(e, h);
        i_eq = ind_ieq(e, h);
    } else {
        return;
    }
    if (!undo) {
        if (vpos >= 0) {
            a[vpos * n + vpos] += g_eq;
            b[vpos] -= i_eq;
        }
        if (vneg >= 0) {
            a[vneg * n + vneg] += g_eq;
            b[vneg] += i_eq;
        }
        if (vneg >= 0 && vpos >= 0) {
            a[vneg * n + vpos] -= g_eq;
            a[vpos * n + vneg] -= g_eq;
        }
    } else {
        if (vpos >= 0) {
            a[vpos * n + vpos] -= g_eq;
            b[vpos] += i_eq;
        }
        if (vneg >= 0) {
            a[vneg * n + vneg] -= g_eq;
            b[vneg] -= i_eq;
        }
        if (vneg >= 0 && vpos >= 0) {
            a[vneg * n + vpos] += g_eq;
            a[vpos * n + vneg] += g_eq;
        }
    }
}

static void init_state(elem_t *e, const double *x)
{
    if (e->kind == ORC_C) { /* cap.rs:38-47 */
        double vpos = xat(x, e->node[1]), vneg = xat(x, e->node[0]);
        e->u_curr = vpos - vneg;
        e->i_curr = 0.0;
    } else if (e->kind == ORC_L) { /* ind.rs:42-49 */
        e->u_curr = 0.0;
        e->i_curr = x[e->branch];
    }
}

static void update_state(elem_t *e, const double *x, double h)
{
    if (e->kind == ORC_C) { /* cap.rs:49-66, cap/model.rs:19-25 */
        double vpos = xat(x, e->node[1]), vneg = xat(x, e->node[0]);
        double u_new = vpos - vneg;
        double i_new = cap_g(e, h) * (u_new - e->u_curr) - e->i_curr;
        e->u_curr = u_new;
        e->i_curr = i_new;
    } else if (e->kind == ORC_L) { /* ind.rs:51-68, ind/model.rs:19-25 */
        double vpos = xat(x, e->node[0]), vneg = xat(x, e->node[1]);
        double u_new = vpos - vneg;
        double i_new = ind_g(e, h) * (u_new + e->u_curr) + e->i_curr;
        e->u_curr = u_new;
        e->i_curr = i_new;
    }
}

static int count_nonlinear_funcs(const elem_t *e)
{
    switch (e->kind) {
    case ORC_D:
        return 1;
    case ORC_Q:
    case ORC_M:
        return 3;
    default:
        return 0;
    }
}

/* nonlinear_stamp: diode.rs:64-93, npn.rs:86-134, nmos.rs:90-134 */
static void nonlinear_stamp(const elem_t *e, const double *x, int n, double *a, double *b)
{
    switch (e->kind) {
    case ORC_D: {
        int vpos_idx = e->node[0], vneg_idx = e->node[1];
        double vp = xat(x, vpos_idx), vn = xat(x, vneg_idx);
        double g_eq = dio_g(vp, vn), i_eq = dio_ieq(vp, vn);
        if (vpos_idx >= 0) {
            a[vpos_idx * n + vpos_idx] += g_eq;
            b[vpos_idx] -= i_eq;
        }
        if (vneg_idx >= 0) {
            a[vneg_idx * n + vneg_idx] += g_eq;
            b[vneg_idx] += i_eq;
        }
        if (vpos_idx >= 0 && vneg_idx >= 0) {
            a[vpos_idx * n + vneg_idx] -= g_eq;
            a[vneg_idx * n + vpos_idx] -= g_eq;
        }
        break;
    }
    case ORC_Q: {
        int c = e->node[0], bb = e->node[1], em = e->node[2];
        npn_t q = {xat(x, c), xat(x, bb), xat(x, em)};
        double gee = q_gee(&q), gec = q_gec(&q), gce = q_gce(&q), gcc = q_gcc(&q);
        double i_e = q_ie_eq(&q), i_c = q_ic_eq(&q);
        if (c >= 0) {
            a[c * n + c] += gcc;
            b[c] -= i_c;
        }
        if (bb >= 0) {
            a[bb * n + bb] += gcc + gee - gce - gec;
            b[bb] += i_e + i_c;
        }
        if (em >= 0) {
            a[em * n + em] += gee;
            b[em] -= i_e;
        }
        if (em >= 0 && c >= 0) {
            a[em * n + c] -= gec;
            a[c * n + em] -= gce;
        }
        if (em >= 0 && bb >= 0) {
            a[em * n + bb] += gec - gee;
            a[bb * n + em] += gce - gee;
        }
        if (c >= 0 && bb >= 0) {
            a[c * n + bb] += gce - gcc;
            a[bb * n + c] += gec - gcc;
        }
        break;
    }
    case ORC_M: {
        int vd_idx = e->node[0], vg_idx = e->node[1], vs_idx = e->node[2];
        double vd = xat(x, vd_idx), vg = xat(x, vg_idx), vs = xat(x, vs_idx);
        if (vs > vd) { /* nmos.rs:105-108: voltages AND indices swap */
            double tv = vd;
            vd = vs;
            vs = tv;
            int ti = vd_idx;
            vd_idx = vs_idx;
            vs_idx = ti;
        }
        nmos_t m = {vd, vg, vs};
        double gds = m_gds(&m), gm = m_gm(&m), ieq = m_ieq(&m);
        if (vd_idx >= 0) {
            a[vd_idx * n + vd_idx] += gds;
            b[vd_idx] -= ieq;
        }
        if (vs_idx >= 0) {
            a[vs_idx * n + vs_idx] += gds + gm;
            b[vs_idx] += ieq;
        }
        if (vd_idx >= 0 && vs_idx >= 0) {
            a[vd_idx * n + vs_idx] -= gds + gm;
            a[vs_idx * n + vd_idx] -= gds;
        }
        if (vd_idx >= 0 && vg_idx >= 0) a[vd_idx * n + vg_idx] += gm;
        if (vs_idx >= 0 && vg_idx >= 0) a[vs_idx * n + vg_idx] -= gm;
        break;
    }
    default:
        break;
    }
}

/* ------------------------------------------------------------------------------------------
 * src/engine/mna.rs:4-30
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int elem;  /* index into elems */
    int which; /* 0,1,2: which closure of the element (npn.rs:72-83, nmos.rs:76-87) */
} gfn_t;

typedef struct {
    int n, m;
    double *a, *b, *h; /* a[n*n], b[n], h[n*m] row-major */
    gfn_t *g;
    int g_len;
} mna_t;

static void mna_new(mna_t *mna, int n, int m)
{
    mna->n = n;
    mna->m = m;
    mna->a = (double *)calloc((size_t)n * n + 1, sizeof(double));
    mna->b = (double *)calloc((size_t)n + 1, sizeof(double));
    mna->h = (double *)calloc((size_t)n * m + 1, sizeof(double));
    mna->g = (gfn_t *)calloc((size_t)m + 1, sizeof(gfn_t));
    mna->g_len = 0;
}
static void mna_free(mna_t *mna)
{
    free(mna->a);
    free(mna->b);
    free(mna->h);
    free(mna->g);
}

/* nonlinear_funcs: diode.rs:39-62, npn.rs:39-84, nmos.rs:39-88 */
static void nonlinear_funcs(const elem_t *e, int ei, mna_t *mna)
{
    int m = mna->m;
    int k = mna->g_len;
    switch (e->kind) {
    case ORC_D:
        if (e->node[0] >= 0) mna->h[e->node[0] * m + k] = 1.0;
        if (e->node[1] >= 0) mna->h[e->node[1] * m + k] = -1.0;
        mna->g[mna->g_len++] = (gfn_t){ei, 0};
        break;
    case ORC_Q:
    case ORC_M:
        if (e->node[0] >= 0) mna->h[e->node[0] * m + k] = 1.0;
        if (e->node[1] >= 0) mna->h[e->node[1] * m + k + 1] = 1.0;
        if (e->node[2] >= 0) mna->h[e->node[2] * m + k + 2] = 1.0;
        mna->g[mna->g_len++] = (gfn_t){ei, 0};
        mna->g[mna->g_len++] = (gfn_t){ei, 1};
        mna->g[mna->g_len++] = (gfn_t){ei, 2};
        break;
    default:
        break;
    }
}

static double eval_g(const elem_t *elems, gfn_t gf, const double *x)
{
    const elem_t *e = &elems[gf.elem];
    switch (e->kind) {
    case ORC_D: /* diode.rs:55-61 */
        return dio_i(xat(x, e->node[0]), xat(x, e->node[1]));
    case ORC_Q: { /* npn.rs:72-83 */
        npn_t q = {xat(x, e->node[0]), xat(x, e->node[1]), xat(x, e->node[2])};
        return gf.which == 0 ? q_ic(&q) : gf.which == 1 ? q_ib(&q) : q_ie(&q);
    }
    case ORC_M: { /* nmos.rs:59-87: voltages swap, indices do not (Q17) */
        double vd = xat(x, e->node[0]), vg = xat(x, e->node[1]), vs = xat(x, e->node[2]);
        if (vs > vd) {
            double t = vs;
            vs = vd;
            vd = t;
        }
        nmos_t m = {vd, vg, vs};
        return gf.which == 0 ? m_id(&m) : gf.which == 1 ? 0.0 : -m_id(&m);
    }
    default:
        return 0.0;
    }
}

/* mna.rs:25-29:  a.dot(x) + h.dot(g_val) - b  */
static void get_err(const mna_t *mna, const elem_t *elems, const double *x, double *out, double *g_val)
{
    int n = mna->n, m = mna->m;
    for (int k = 0; k < mna->g_len; k++) g_val[k] = eval_g(elems, mna->g[k], x);
    for (int i = 0; i < n; i++) {
        double ax = orc_unrolled_dot(&mna->a[(size_t)i * n], x, n);
        double hg = orc_unrolled_dot(&mna->h[(size_t)i * m], g_val, m);
        out[i] = ax + hg - mna->b[i];
    }
}


/* ------------------------------------------------------------------------------------------
 * element-level hooks for the reference's device unit tests (src/device/<kind>.rs `mod tests`)
 * op: 0 linear_stamp, 1 undo_linear_stamp, 2 linear_startup_stamp, 3 dynamic_stamp,
 *     4 undo_dynamic_stamp, 5 nonlinear_stamp.  u, i = the Cap/Ind state (u_curr, i_curr).
 * ---------------------------------------------------------------------------------------- */
static void elem_from(elem_t *e, const orc_elem_t *oe, double u, double i)
{
    memset(e, 0, sizeof *e);
    e->kind = oe->kind;
    memcpy(e->node, oe->node, sizeof e->node);
    e->branch = oe->branch;
    e->fn_kind = oe->fn_kind;
    e->val = oe->val;
    memcpy(e->fn, oe->fn, sizeof e->fn);
    e->u_curr = u;
    e->i_curr = i;
}

int orc_elem_apply(const orc_elem_t *oe, double u, double i, int op, int n, const double *x, double h, double *a,
                   double *b)
{
    elem_t e;
    elem_from(&e, oe, u, i);
    tl_panic = 0;
    switch (op) {
    case 0: linear_stamp(&e, n, a, b); break;
    case 1: undo_linear_stamp(&e, n, a, b); break;
    case 2: linear_startup_stamp(&e, n, a, b); break;
    case 3: dynamic_stamp(&e, h, n, a, b, 0); break;
    case 4: dynamic_stamp(&e, h, n, a, b, 1); break;
    case 5: nonlinear_stamp(&e, x, n, a, b); break;
    default: return -1;
    }
    return tl_panic;
}

/* nonlinear_funcs into hmat[n*m] starting at column col0; evaluates the closures at x into gvals.
 * Returns count_nonlinear_funcs(). */
int orc_elem_funcs(const orc_elem_t *oe, int n, int m, int col0, const double *x, double *hmat, double *gvals)
{
    elem_t e;
    elem_from(&e, oe, 0.0, 0.0);
    mna_t mna;
    mna.n = n;
    mna.m = m;
    mna.a = NULL;
    mna.b = NULL;
    mna.h = hmat;
    gfn_t g[64];
    if (col0 < 0 || col0 > 60) return -1;
    mna.g = g;
    mna.g_len = col0; /* g_vec.len() at the time of the call */
    nonlinear_funcs(&e, 0, &mna);
    int cnt = mna.g_len - col0;
    tl_panic = 0;
    if (x && gvals)
        for (int k = 0; k < cnt; k++) gvals[k] = eval_g(&e, (gfn_t){0, k}, x);
    return cnt;
}

/* mna.rs:25-29 with caller-supplied g values: out = a.dot(x) + h.dot(g) - b  (KAT mna.rs:36-52) */
void orc_get_err_raw(int n, int m, const double *a, const double *b, const double *h, const double *gval,
                     const double *x, double *out)
{
    for (int i = 0; i < n; i++)
        out[i] = orc_unrolled_dot(&a[(size_t)i * n], x, n) + orc_unrolled_dot(&h[(size_t)i * m], gval, m) - b[i];
}

/* ------------------------------------------------------------------------------------------
 * node collection (index-lowered): class + BTreeMap value order
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int n;
    const int32_t *cls;   /* by unknown index */
    const int32_t *order; /* nodes.values() order */
} nodes_t;

/* node_vec_norm.rs:17-33 */
static nvn_t nvn_new(const nodes_t *nodes, const double *v, double *scratch)
{
    int nv = 0, ni = 0;
    double *vv = scratch, *iv = scratch + nodes->n;
    for (int k = 0; k < nodes->n; k++) {
        int idx = nodes->order[k];
        if (nodes->cls[idx] == 0) vv[nv++] = v[idx];
    }
    for (int k = 0; k < nodes->n; k++) {
        int idx = nodes->order[k];
        if (nodes->cls[idx] == 1) iv[ni++] = v[idx];
    }
    nvn_t r;
    r.v = orc_norm(vv, nv);
    r.i = orc_norm(iv, ni);
    return r;
}

/* ------------------------------------------------------------------------------------------
 * src/engine/newtons_method.rs:19-71
 * returns n_iters (>=0) on Ok, -(status) on error
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    double *f0, *jf, *bt, *xp, *stepp, *stept, *xn, *gval, *scr;
} nws_t;

static void nws_alloc(nws_t *w, int n, int m)
{
    w->f0 = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    w->jf = (double *)malloc(sizeof(double) * ((size_t)n * n + 1));
    w->bt = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    w->xp = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    w->stepp = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    w->stept = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    w->xn = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    w->gval = (double *)malloc(sizeof(double) * (size_t)(m + 1));
    w->scr = (double *)malloc(sizeof(double) * (size_t)(2 * n + 2));
}
static void nws_free(nws_t *w)
{
    free(w->f0);
    free(w->jf);
    free(w->bt);
    free(w->xp);
    free(w->stepp);
    free(w->stept);
    free(w->xn);
    free(w->gval);
    free(w->scr);
}

static int64_t newton_solve(const nodes_t *nodes, const elem_t *elems, int n_elems, double *x, const mna_t *mna,
                            nws_t *w, orc_stats_t *st)
{
    int n = mna->n;
    nvn_t err = {INFINITY, INFINITY}, step = {INFINITY, INFINITY};

    tl_panic = 0;
    get_err(mna, elems, x, w->f0, w->gval);
    if (tl_panic) return -(int64_t)tl_panic;
    nvn_t err_old = nvn_new(nodes, w->f0, w->scr);
    nvn_t step_old = err_old;

    int64_t n_iters = 0;
    while (n_iters < MAX_ITERS &&
           !orc_converged(err.v, err.i, step.v, step.i, err_old.v, err_old.i, step_old.v, step_old.i)) {
        memcpy(w->jf, mna->a, sizeof(double) * (size_t)n * n);
        memcpy(w->bt, mna->b, sizeof(double) * (size_t)n);
        memcpy(w->xp, x, sizeof(double) * (size_t)n);

        for (int e = 0; e < n_elems; e++) nonlinear_stamp(&elems[e], w->xp, n, w->jf, w->bt);
        if (tl_panic) return -(int64_t)tl_panic;

        orc_lu_solve(n, w->jf, w->bt, w->xp);
        if (st) st->lu_solves++;

        for (int i = 0; i < n; i++) w->stepp[i] = w->xp[i] - x[i];
        orc_dampen_step(w->stepp, w->stept, n);
        step = nvn_new(nodes, w->stept, w->scr);

        for (int i = 0; i < n; i++) w->xn[i] = x[i] + w->stept[i];

        get_err(mna, elems, w->xn, w->f0, w->gval);
        if (tl_panic) return -(int64_t)tl_panic;
        err = nvn_new(nodes, w->f0, w->scr);
        if (isinf(err.v) || isinf(err.i)) return -(int64_t)ORC_DIVERGED; /* panic, :53-55 */

        for (int i = 0; i < n; i++) x[i] = w->xn[i];

        n_iters += 1;
        if (st) st->newton_iters++;
        err_old = err;
        step_old = step;
    }
    if (n_iters < MAX_ITERS) return n_iters;
    return -(int64_t)ORC_NOT_CONVERGED;
}

/* ------------------------------------------------------------------------------------------
 * Engine (src/engine.rs:22-206)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const orc_circuit_t *c;
    elem_t *elems;
    int n_elems;
    int num_nonlinear_funcs;
} engine_t;

static void engine_new(engine_t *eng, const orc_circuit_t *c, const double *vals, const double *fns)
{
    eng->c = c;
    eng->n_elems = c->n_elems;
    eng->elems = (elem_t *)calloc((size_t)c->n_elems + 1, sizeof(elem_t));
    int m = 0;
    for (int i = 0; i < c->n_elems; i++) {
        elem_t *e = &eng->elems[i];
        const orc_elem_t *o = &c->elems[i];
        e->kind = o->kind;
        memcpy(e->node, o->node, sizeof e->node);
        e->branch = o->branch;
        e->fn_kind = o->fn_kind;
        e->val = vals ? vals[i] : o->val;
        memcpy(e->fn, fns ? &fns[(size_t)i * 7] : o->fn, sizeof e->fn);
        e->u_curr = NAN; /* Option::None */
        e->i_curr = NAN;
        m += count_nonlinear_funcs(e);
    }
    eng->num_nonlinear_funcs = m; /* engine.rs:45 */
    for (int i = 0; i < c->n_elems; i++) eval_tran(&eng->elems[i], 0.0); /* engine.rs:47-51 */
}
static void engine_free(engine_t *eng) { free(eng->elems); }

/* engine.rs:62-90; x_out has n_start entries */
static int engine_run_op(engine_t *eng, double *x_out, int64_t *n_iters_out, orc_stats_t *st)
{
    const orc_circuit_t *c = eng->c;
    nodes_t nodes = {c->n_start, c->cls, c->order_start};
    int n = c->n_start;
    mna_t mna;
    mna_new(&mna, n, eng->num_nonlinear_funcs);
    for (int i = 0; i < eng->n_elems; i++) {
        linear_startup_stamp(&eng->elems[i], n, mna.a, mna.b);
        nonlinear_funcs(&eng->elems[i], i, &mna);
    }
    double *x = (double *)calloc((size_t)n + 1, sizeof(double));
    nws_t w;
    nws_alloc(&w, n, mna.m);
    int64_t r = newton_solve(&nodes, eng->elems, eng->n_elems, x, &mna, &w, st);
    int status = ORC_OK;
    if (r < 0) {
        status = (int)-r;
    } else {
        for (int i = 0; i < eng->n_elems; i++) init_state(&eng->elems[i], x);
        if (n_iters_out) *n_iters_out = r;
    }
    if (x_out) memcpy(x_out, x, sizeof(double) * (size_t)n); /* on error: last iterate, diagnostic only */
    nws_free(&w);
    free(x);
    mna_free(&mna);
    return status;
}

int64_t orc_dc_count(double start, double stop, double step)
{
    /* ndarray 0.15 linspace::range: steps = ceil((b - a) / step).to_usize() */
    double steps = ceil((stop - start) / step);
    if (!(steps >= 0.0)) return 0; /* the reference would panic in to_usize().expect() */
    return (int64_t)steps;
}

/* engine.rs:92-140 */
static int engine_run_dc(engine_t *eng, int sweep_idx, double start, double stop, double step, double *x_out,
                         int64_t *n_iters_out, double *sweep_out, int64_t *points_done, orc_stats_t *st)
{
    const orc_circuit_t *c = eng->c;
    nodes_t nodes = {c->n_run, c->cls, c->order_run};
    int n = c->n_run;
    mna_t mna;
    mna_new(&mna, n, eng->num_nonlinear_funcs);
    for (int i = 0; i < eng->n_elems; i++) {
        linear_stamp(&eng->elems[i], n, mna.a, mna.b);
        nonlinear_funcs(&eng->elems[i], i, &mna);
    }
    double *x = (double *)calloc((size_t)n + 1, sizeof(double));
    nws_t w;
    nws_alloc(&w, n, mna.m);
    double val_bkp = eng->elems[sweep_idx].val;
    int64_t count = orc_dc_count(start, stop, step);
    int status = ORC_OK;
    int64_t done = 0;
    for (int64_t k = 0; k < count; k++) {
        double sweep_val = start + step * (double)k;
        undo_linear_stamp(&eng->elems[sweep_idx], n, mna.a, mna.b);
        eng->elems[sweep_idx].val = sweep_val;
        linear_stamp(&eng->elems[sweep_idx], n, mna.a, mna.b);
        int64_t r = newton_solve(&nodes, eng->elems, eng->n_elems, x, &mna, &w, st);
        if (r < 0) {
            status = (int)-r; /* `?` aborts the whole sweep (Q6) */
            break;
        }
        memcpy(&x_out[(size_t)k * n], x, sizeof(double) * (size_t)n);
        n_iters_out[k] = r;
        if (sweep_out) sweep_out[k] = sweep_val;
        done++;
    }
    eng->elems[sweep_idx].val = val_bkp;
    *points_done = done;
    nws_free(&w);
    free(x);
    mna_free(&mna);
    return status;
}

/* ------------------------------------------------------------------------------------------
 * src/engine/transient/state_history.rs:4-57
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int n;
    int64_t len, cap_alloc;
    int64_t *n_iters;
    double *t;
    double *x; /* [len][n] */
} hist_t;

static void hist_push(hist_t *hst, int64_t n_iters, const double *x, double t)
{
    if (hst->len == hst->cap_alloc) {
        hst->cap_alloc = hst->cap_alloc ? hst->cap_alloc * 2 : 256;
        hst->n_iters = (int64_t *)realloc(hst->n_iters, sizeof(int64_t) * (size_t)hst->cap_alloc);
        hst->t = (double *)realloc(hst->t, sizeof(double) * (size_t)hst->cap_alloc);
        hst->x = (double *)realloc(hst->x, sizeof(double) * (size_t)hst->cap_alloc * hst->n);
    }
    hst->n_iters[hst->len] = n_iters;
    hst->t[hst->len] = t;
    memcpy(&hst->x[(size_t)hst->len * hst->n], x, sizeof(double) * (size_t)hst->n);
    hst->len++;
}

/* divided_diff(n_max, n_min), recursive like :48-56; out[n] */
static void divided_diff(const hist_t *hst, int64_t n_max, int64_t n_min, double *out)
{
    int n = hst->n;
    if (n_max == n_min) {
        memcpy(out, &hst->x[(size_t)n_max * n], sizeof(double) * (size_t)n);
    } else {
        double alpha = 1.0 / (hst->t[n_max] - hst->t[n_min]);
        double *a1 = (double *)malloc(sizeof(double) * (size_t)(2 * n + 2));
        double *a2 = a1 + n;
        divided_diff(hst, n_max, n_min + 1, a1);
        divided_diff(hst, n_max - 1, n_min, a2);
        for (int i = 0; i < n; i++) out[i] = alpha * (a1[i] - a2[i]);
        free(a1);
    }
}

static double powi3(double a) { return a * a * a; } /* f64::powi(3): (a*a)*a == a*(a*a) bitwise */

static void hist_plte(const hist_t *hst, int64_t nn, double *out)
{
    double c3 = -1.0 / 12.0;
    double h_next = hst->t[nn + 1] - hst->t[nn];
    double alpha = 6.0 * c3 * powi3(h_next);
    divided_diff(hst, nn + 1, nn - 2, out);
    for (int i = 0; i < hst->n; i++) out[i] = alpha * out[i];
}

/* ------------------------------------------------------------------------------------------
 * src/engine/transient.rs:12-105
 * ---------------------------------------------------------------------------------------- */
static const double T_STEP_MIN = 1e-18;
static const double TR_TOL_REL = 0.001, TR_TOL_ABS_V = 1e-3, TR_TOL_ABS_A = 1e-6;

static int plte_is_too_big(nvn_t plte, nvn_t x)
{
    return plte.v > x.v * TR_TOL_REL + TR_TOL_ABS_V || plte.i > x.i * TR_TOL_REL + TR_TOL_ABS_A;
}
static int plte_can_grow(nvn_t plte) { return plte.v < 0.1 * TR_TOL_ABS_V && plte.i < 0.1 * TR_TOL_ABS_A; }

/* returns status; on OK *h_out, *next_h_out set */
static int tran_step(const nodes_t *nodes, elem_t *elems, int n_elems, mna_t *mna, double t, double h_in, double *x,
                     hist_t *hst, double step_max, nws_t *w, double *a_bkp, double *b_bkp, double *plte_buf,
                     double *h_out, double *next_h_out, orc_stats_t *st)
{
    int n = mna->n;
    double h = h_in;
    double next_h = h;
    int step_accepted = 0;

    memcpy(a_bkp, mna->a, sizeof(double) * (size_t)n * n);
    memcpy(b_bkp, mna->b, sizeof(double) * (size_t)n);

    while (!step_accepted) {
        for (int e = 0; e < n_elems; e++) {
            if (has_tran(&elems[e])) {
                undo_linear_stamp(&elems[e], n, mna->a, mna->b);
                eval_tran(&elems[e], t + h);
                linear_stamp(&elems[e], n, mna->a, mna->b);
            }
        }
        for (int e = 0; e < n_elems; e++) dynamic_stamp(&elems[e], h, n, mna->a, mna->b, 0);

        int64_t r = newton_solve(nodes, elems, n_elems, x, mna, w, st);

        if (r == -(int64_t)ORC_NOT_CONVERGED) {
            h /= 2.0;
            step_accepted = 0;
            if (st) st->rej_newton++;
        } else if (r < 0) {
            return (int)-r; /* panic */
        } else if (hst->len < 3) {
            hist_push(hst, r, x, t + h);
            next_h = h;
            step_accepted = 1;
        } else {
            hist_push(hst, r, x, t + h);
            hist_plte(hst, hst->len - 2, plte_buf);
            nvn_t plte_norm = nvn_new(nodes, plte_buf, w->scr);
            nvn_t x_norm = nvn_new(nodes, x, w->scr);
            step_accepted = !plte_is_too_big(plte_norm, x_norm);
            if (!step_accepted) {
                h /= 2.0;
                hst->len--; /* pop */
                if (st) st->rej_plte++;
            } else {
                next_h = (plte_can_grow(plte_norm) && h <= step_max / 2.0) ? h * 2.0 : h;
            }
        }

        if (!step_accepted) {
            /* NOTE: h has ALREADY been halved (Q10) */
            for (int e = 0; e < n_elems; e++) dynamic_stamp(&elems[e], h, n, mna->a, mna->b, 1);
        }

        if (h < T_STEP_MIN) return ORC_TRAN_H_UNDERFLOW;
    }

    memcpy(mna->a, a_bkp, sizeof(double) * (size_t)n * n);
    memcpy(mna->b, b_bkp, sizeof(double) * (size_t)n);
    *h_out = h;
    *next_h_out = next_h;
    return ORC_OK;
}

/* engine.rs:142-206 */
static int engine_run_tran(engine_t *eng, double stop, double step_max, int64_t cap, double *t_out, double *x_out,
                           int64_t *n_iters_out, int64_t *n_rec, double *x_final, orc_stats_t *st)
{
    const orc_circuit_t *c = eng->c;
    nodes_t nodes = {c->n_run, c->cls, c->order_run};
    int n = c->n_run;
    mna_t mna;
    mna_new(&mna, n, eng->num_nonlinear_funcs);
    double *x = (double *)calloc((size_t)n + 1, sizeof(double));
    for (int i = 0; i < eng->n_elems; i++) {
        linear_stamp(&eng->elems[i], n, mna.a, mna.b);
        nonlinear_funcs(&eng->elems[i], i, &mna);
    }
    *n_rec = 0;

    /* start-up solution: a full run_op with its own numbering (Q13) */
    double *xs = (double *)calloc((size_t)c->n_start + 1, sizeof(double));
    int64_t it0 = 0;
    int status = engine_run_op(eng, xs, &it0, st);
    if (status != ORC_OK) {
        free(xs);
        free(x);
        mna_free(&mna);
        return status;
    }
    for (int i = 0; i < n; i++) x[i] = xs[i]; /* copy by name == copy by index for the shared prefix */
    free(xs);

    hist_t hst;
    memset(&hst, 0, sizeof hst);
    hst.n = n;
    nws_t w;
    nws_alloc(&w, n, mna.m);
    double *a_bkp = (double *)malloc(sizeof(double) * ((size_t)n * n + 1));
    double *b_bkp = (double *)malloc(sizeof(double) * ((size_t)n + 1));
    double *plte_buf = (double *)malloc(sizeof(double) * ((size_t)n + 1));

    double t = 0.0; /* tran_params.start, parser.rs:247 */
    double h = T_STEP_MIN;
    double next_h;
    while (t < stop) {
        status = tran_step(&nodes, eng->elems, eng->n_elems, &mna, t, h, x, &hst, step_max, &w, a_bkp, b_bkp,
                           plte_buf, &h, &next_h, st);
        if (status != ORC_OK) break;
        t += h;
        for (int e = 0; e < eng->n_elems; e++) update_state(&eng->elems[e], x, h);
        h = next_h;
    }

    /* the reference prints nothing on error (`?`); we still hand back what was accepted so far */
    *n_rec = hst.len;
    for (int64_t r = 0; r < hst.len && r < cap; r++) {
        if (t_out) t_out[r] = hst.t[r];
        if (n_iters_out) n_iters_out[r] = hst.n_iters[r];
        if (x_out) memcpy(&x_out[(size_t)r * n], &hst.x[(size_t)r * n], sizeof(double) * (size_t)n);
    }
    if (x_final) memcpy(x_final, x, sizeof(double) * (size_t)n);

    free(hst.n_iters);
    free(hst.t);
    free(hst.x);
    free(a_bkp);
    free(b_bkp);
    free(plte_buf);
    nws_free(&w);
    free(x);
    mna_free(&mna);
    return status;
}

/* ------------------------------------------------------------------------------------------
 * public single-instance entry points
 * ---------------------------------------------------------------------------------------- */
int orc_run_op(const orc_circuit_t *c, double *x_out, int64_t *n_iters_out, orc_stats_t *st)
{
    engine_t eng;
    engine_new(&eng, c, NULL, NULL);
    int s = engine_run_op(&eng, x_out, n_iters_out, st);
    engine_free(&eng);
    return s;
}

int orc_run_dc(const orc_circuit_t *c, int sweep_elem, double start, double stop, double step, double *x_out,
               int64_t *n_iters_out, double *sweep_out, int64_t *points_done, orc_stats_t *st)
{
    engine_t eng;
    engine_new(&eng, c, NULL, NULL);
    int s = engine_run_dc(&eng, sweep_elem, start, stop, step, x_out, n_iters_out, sweep_out, points_done, st);
    engine_free(&eng);
    return s;
}

int orc_run_tran(const orc_circuit_t *c, double stop, double step_max, int64_t cap, double *t_out, double *x_out,
                 int64_t *n_iters_out, int64_t *n_rec, double *x_final, orc_stats_t *st)
{
    engine_t eng;
    engine_new(&eng, c, NULL, NULL);
    int s = engine_run_tran(&eng, stop, step_max, cap, t_out, x_out, n_iters_out, n_rec, x_final, st);
    engine_free(&eng);
    return s;
}

/* ------------------------------------------------------------------------------------------
 * batched drivers (instances are independent; one reference process per instance)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int mode; /* 0 op, 1 dc, 2 tran */
    const orc_circuit_t *c;
    int64_t batch;
    const double *vals, *fns;
    /* dc */
    int sweep_elem;
    const double *start, *stop, *step;
    int64_t max_points;
    int64_t *points_done;
    /* tran */
    double t_stop, step_max;
    int64_t cap;
    double *t_out;
    int64_t *n_rec;
    double *x_final;
    orc_stats_t *stats_out;
    /* common out */
    double *x_out;
    int64_t *n_iters_out;
    int32_t *status_out;
    /* work distribution */
    int64_t next;
    pthread_mutex_t mu;
} job_t;

static void run_one(job_t *j, int64_t i)
{
    const orc_circuit_t *c = j->c;
    engine_t eng;
    engine_new(&eng, c, j->vals ? &j->vals[(size_t)i * c->n_elems] : NULL,
               j->fns ? &j->fns[(size_t)i * c->n_elems * 7] : NULL);
    orc_stats_t st;
    memset(&st, 0, sizeof st);
    int s = 0;
    if (j->mode == 0) {
        int64_t it = 0;
        s = engine_run_op(&eng, &j->x_out[(size_t)i * c->n_start], &it, &st);
        j->n_iters_out[i] = s == ORC_OK ? it : 0;
    } else if (j->mode == 1) {
        int64_t done = 0;
        s = engine_run_dc(&eng, j->sweep_elem, j->start[i], j->stop[i], j->step[i],
                          &j->x_out[(size_t)i * j->max_points * c->n_run],
                          &j->n_iters_out[(size_t)i * j->max_points], NULL, &done, &st);
        j->points_done[i] = done;
    } else {
        int64_t nrec = 0;
        s = engine_run_tran(&eng, j->t_stop, j->step_max, j->cap, j->t_out ? &j->t_out[(size_t)i * j->cap] : NULL,
                            j->x_out ? &j->x_out[(size_t)i * j->cap * c->n_run] : NULL,
                            j->n_iters_out ? &j->n_iters_out[(size_t)i * j->cap] : NULL, &nrec,
                            j->x_final ? &j->x_final[(size_t)i * c->n_run] : NULL, &st);
        j->n_rec[i] = nrec;
    }
    j->status_out[i] = s;
    if (j->stats_out) j->stats_out[i] = st;
    engine_free(&eng);
}

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int64_t i = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (i >= j->batch) break;
        run_one(j, i);
    }
    return NULL;
}

static int run_job(job_t *j, int threads)
{
    j->next = 0;
    pthread_mutex_init(&j->mu, NULL);
    if (threads <= 1) {
        worker(j);
    } else {
        if (threads > 256) threads = 256;
        pthread_t th[256];
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, j);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    }
    pthread_mutex_destroy(&j->mu);
    return 0;
}

int orc_batch_op(const orc_circuit_t *c, int64_t batch, const double *vals, const double *fns, double *x_out,
                 int64_t *n_iters_out, int32_t *status_out, int threads)
{
    job_t j;
    memset(&j, 0, sizeof j);
    j.mode = 0;
    j.c = c;
    j.batch = batch;
    j.vals = vals;
    j.fns = fns;
    j.x_out = x_out;
    j.n_iters_out = n_iters_out;
    j.status_out = status_out;
    return run_job(&j, threads);
}

int orc_batch_dc(const orc_circuit_t *c, int64_t batch, const double *vals, const double *fns, int sweep_elem,
                 const double *start, const double *stop, const double *step, int64_t max_points, double *x_out,
                 int64_t *n_iters_out, int64_t *points_done, int32_t *status_out, int threads)
{
    job_t j;
    memset(&j, 0, sizeof j);
    j.mode = 1;
    j.c = c;
    j.batch = batch;
    j.vals = vals;
    j.fns = fns;
    j.sweep_elem = sweep_elem;
    j.start = start;
    j.stop = stop;
    j.step = step;
    j.max_points = max_points;
    j.points_done = points_done;
    j.x_out = x_out;
    j.n_iters_out = n_iters_out;
    j.status_out = status_out;
    return run_job(&j, threads);
}

int orc_batch_tran(const orc_circuit_t *c, int64_t batch, const double *vals, const double *fns, double stop,
                   double step_max, int64_t cap, double *t_out, double *x_out, int64_t *n_iters_out, int64_t *n_rec,
                   double *x_final, int32_t *status_out, orc_stats_t *stats_out, int threads)
{
    job_t j;
    memset(&j, 0, sizeof j);
    j.mode = 2;
    j.c = c;
    j.batch = batch;
    j.vals = vals;
    j.fns = fns;
    j.t_stop = stop;
    j.step_max = step_max;
    j.cap = cap;
    j.t_out = t_out;
    j.x_out = x_out;
    j.n_iters_out = n_iters_out;
    j.n_rec = n_rec;
    j.x_final = x_final;
    j.status_out = status_out;
    j.stats_out = stats_out;
    return run_job(&j, threads);
}
