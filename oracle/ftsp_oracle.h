/*
 * ftsp_oracle.h — CPU ORACLE for the ftspice `src/engine` hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`,
 * `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
 * legs may load it.  The product (libftsb.so, the CUDA path) never calls it and
 * has no CPU fallback.
 *
 * It is a plain-C restatement of the reference's algorithm, file by file,
 * quirks included (the reference is Rust; no Rust toolchain exists in this
 * image, so the reference itself cannot be compiled: "parity unpinned" for the
 * end-to-end .op/.dc/.tran results — the reference holds no golden output for
 * them.  What IS pinned: every known-answer test the reference's own unit tests
 * hold for this path, see tests/test_oracle_kats.py).
 *
 * Citations are relative to /root/reference/.
 */
#ifndef FTSP_ORACLE_H
#define FTSP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element kinds (one per file under src/device/) */
enum { ORC_R = 0, ORC_C = 1, ORC_L = 2, ORC_V = 3, ORC_I = 4, ORC_D = 5, ORC_Q = 6, ORC_M = 7 };
/* waveform kinds (src/spice_fn.rs:2-6) */
enum { ORC_FN_NONE = 0, ORC_FN_SIN = 1, ORC_FN_PULSE = 2, ORC_FN_EXP = 3 };
/* result of an analysis */
enum {
    ORC_OK = 0,
    ORC_NOT_CONVERGED = 1,    /* Err(NotConvergedError) from newtons_method.rs:66-70           */
    ORC_DIVERGED = 2,         /* panic!("v_err or i_err diverged")   newtons_method.rs:53-55   */
    ORC_TRAN_H_UNDERFLOW = 3, /* Err(NotConvergedError) from transient.rs:88-90                */
    ORC_NMOS_UNREACHABLE = 4  /* unreachable!() in nmos/model.rs:38-40 (NaN terminal voltage)  */
};

/* One element, already index-lowered the way NodeCollection would do it
 * (src/node_collection.rs:13-74).  node[] is the `nodes` Vec of the Rust struct
 * (i.e. AFTER the parser's terminal swaps, src/parser.rs:83-84,114-115,145-146,
 * 161-162,177-178); -1 = GND. */
typedef struct {
    int32_t kind;
    int32_t node[3];
    int32_t branch;  /* V: its Current-type row (both numberings); L: its start-up row; else -1 */
    int32_t fn_kind;
    double val;      /* R, C, L, V, I value (V/I with a waveform: f(0), parser.rs:95-99) */
    double fn[7];    /* SIN: offset, amplitude, freq; PULSE: v1, v2, delay, t_rise, t_fall,
                        pulse_width, period; EXP: v1, v2, rise_delay, rise_tau, fall_delay, fall_tau */
} orc_elem_t;

typedef struct {
    int32_t n_elems;
    int32_t n_run;   /* NodeCollection::from_elems(..).len()          */
    int32_t n_start; /* NodeCollection::from_startup_elems(..).len()  */
    int32_t pad_;
    const orc_elem_t *elems;    /* netlist order */
    const int32_t *cls;         /* [n_start] 0 = Voltage, 1 = Current */
    const int32_t *order_run;   /* [n_run]   unknown indices in BTreeMap key order (nodes.values()) */
    const int32_t *order_start; /* [n_start] same for the start-up numbering */
} orc_circuit_t;

typedef struct {
    int64_t newton_iters;   /* total Newton iterations executed */
    int64_t lu_solves;      /* total gauss_lu::solve calls      */
    int64_t rej_newton;     /* .tran attempts rejected because Newton failed */
    int64_t rej_plte;       /* .tran attempts rejected by the PLTE test      */
} orc_stats_t;

/* ---- building blocks, exported for the reference's known-answer tests ---- */
void orc_lu_solve(int n, double *a /*n*n row-major, destroyed*/, double *b /*destroyed*/, double *x);
double orc_norm(const double *v, int len);                     /* node_vec_norm.rs:12-14 */
void orc_dampen_step(const double *step, double *out, int n);  /* newtons_method.rs:73-77 */
int orc_converged(double err_v, double err_i, double step_v, double step_i, double err_old_v,
                  double err_old_i, double step_old_v, double step_old_i); /* :79-89 */
double orc_unrolled_dot(const double *x, const double *y, int n);
double orc_spice_fn_eval(int fn_kind, const double *p, double t); /* spice_fn.rs:8-124 */
/* stamps one element into dense a[n*n], b[n]; mode 0 linear_stamp, 1 undo_linear_stamp,
 * 2 linear_startup_stamp.  For the stamp KATs of res.rs/vdd.rs/idd.rs. */
void orc_stamp_linear(const orc_elem_t *e, int n, double *a, double *b, int mode);
/* device model values for the sign/pattern KATs: out = {i/ic/id, g/..}. */
void orc_diode_model(double vpos, double vneg, double *i, double *g_eq, double *i_eq);
void orc_npn_model(double vc, double vb, double ve, double out[9]);
/* out: ie, ic, ib, gee, gec, gce, gcc, ie_eq, ic_eq */
int orc_nmos_model(double vd, double vg, double vs, double out[4]); /* id, gds, gm, ieq; ret !=0 → unreachable */

/* element-level hooks for the reference's device unit tests; op: 0 linear_stamp, 1 undo_linear_stamp,
 * 2 linear_startup_stamp, 3 dynamic_stamp, 4 undo_dynamic_stamp, 5 nonlinear_stamp */
int orc_elem_apply(const orc_elem_t *e, double u_curr, double i_curr, int op, int n, const double *x, double h,
                   double *a, double *b);
int orc_elem_funcs(const orc_elem_t *e, int n, int m, int col0, const double *x, double *hmat, double *gvals);
void orc_get_err_raw(int n, int m, const double *a, const double *b, const double *h, const double *gval,
                     const double *x, double *out);

/* ---- the three analyses (src/engine.rs:62-206) ---- */
/* x_out[n_start]; n_iters_out single */
int orc_run_op(const orc_circuit_t *c, double *x_out, int64_t *n_iters_out, orc_stats_t *st);
/* points = ceil((stop-start)/step) (ndarray Array::range); returns status; rows written = *points_done */
int64_t orc_dc_count(double start, double stop, double step);
int orc_run_dc(const orc_circuit_t *c, int sweep_elem, double start, double stop, double step,
               double *x_out /*[points][n_run]*/, int64_t *n_iters_out /*[points]*/,
               double *sweep_out /*[points] or NULL*/, int64_t *points_done, orc_stats_t *st);
/* records up to cap accepted steps (all are counted in *n_rec even beyond cap). */
int orc_run_tran(const orc_circuit_t *c, double stop, double step_max, int64_t cap,
                 double *t_out /*[cap]*/, double *x_out /*[cap][n_run]*/, int64_t *n_iters_out /*[cap]*/,
                 int64_t *n_rec, double *x_final /*[n_run] or NULL*/, orc_stats_t *st);

/* batched helpers used by bench.py's cpu_baseline and by the parity tests: run `batch`
 * instances that share the topology but have their own element values
 * (vals[batch][n_elems], fn[batch][n_elems][7]).  threads > 1 uses pthreads over instances. */
int orc_batch_op(const orc_circuit_t *c, int64_t batch, const double *vals, const double *fns,
                 double *x_out /*[batch][n_start]*/, int64_t *n_iters_out, int32_t *status_out,
                 int threads);
int orc_batch_dc(const orc_circuit_t *c, int64_t batch, const double *vals, const double *fns,
                 int sweep_elem, const double *start, const double *stop, const double *step,
                 int64_t max_points, double *x_out /*[batch][max_points][n_run]*/,
                 int64_t *n_iters_out /*[batch][max_points]*/, int64_t *points_done,
                 int32_t *status_out, int threads);
int orc_batch_tran(const orc_circuit_t *c, int64_t batch, const double *vals, const double *fns,
                   double stop, double step_max, int64_t cap, double *t_out, double *x_out,
                   int64_t *n_iters_out, int64_t *n_rec, double *x_final, int32_t *status_out,
                   orc_stats_t *stats_out /*[batch] or NULL*/, int threads);

#ifdef __cplusplus
}
#endif
#endif
