/*
 * ftsb.h — C-ABI of the B200 batched circuit-solve engine (libftsb.so).
 *
 * Drop-in boundary for ftspice's `src/engine` hot path.  The reference has no FFI; the seam is the
 * method set of `Engine` (src/engine.rs:22-206).  Each entry point below names the reference
 * interface it replaces.  A front end (the reference's Rust parser + NodeCollection, or the C++
 * mirror in ftspice_b200/host/) lowers its element list to indices and calls:
 *
 *   ftsb_compile      <- Engine::new(elems, cmds)                    src/engine.rs:31-60
 *                        + NodeCollection::from_elems / from_startup_elems   src/node_collection.rs:13-74
 *   ftsb_set_params   <- the per-element mutable values: Stamp::set_value (src/device.rs:27),
 *                        Res/Cap/Ind/Vdd/Idd `val` (res.rs:10, cap.rs:12, ind.rs:12, vdd.rs:11, idd.rs:11),
 *                        SpiceFn parameters (src/spice_fn.rs:19-90) — one set per batched instance
 *   ftsb_run_op       <- Engine::run_op                              src/engine.rs:62-90
 *   ftsb_run_dc       <- Engine::run_dc                              src/engine.rs:92-140
 *   ftsb_run_tran     <- Engine::run_tran                            src/engine.rs:142-206
 *   per-instance `status` <- Result<_, NotConvergedError> (src/engine/error.rs:3-10) and the two panics
 *                        (src/engine/newtons_method.rs:53-55, src/device/nmos/model.rs:38-40)
 *
 * Everything is plain C: pointers and sizes, no C++ or torch types.  Every call returns 0 on success
 * or a negative FTSB_E_* code; ftsb_last_error() gives the text.  An ftsb_circuit handle = one CUDA device
 * and one stream (an ftsb_multi handle = a list of devices, sharding internal); a handle is used by one
 * thread at a time.  There is NO CPU fallback: without a usable CUDA
 * device ftsb_compile fails with FTSB_E_CUDA.
 *
 * A *batch* is a set of independent circuit instances sharing one topology: Monte-Carlo parameter
 * sets, chunks of a .dc sweep ("chains"), parallel .tran runs.  Instance i of a batch gives exactly the
 * result the reference gives for a netlist with instance i's element values.
 */
#ifndef FTSB_H
#define FTSB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FTSB_ABI_VERSION 1

/* element kinds — one per Stamp implementor (src/device/{res,cap,ind,vdd,idd,diode,npn,nmos}.rs) */
enum { FTSB_R = 0, FTSB_C = 1, FTSB_L = 2, FTSB_V = 3, FTSB_I = 4, FTSB_D = 5, FTSB_Q = 6, FTSB_M = 7 };
/* source waveforms — SpiceFn (src/spice_fn.rs:2-6) */
enum { FTSB_FN_NONE = 0, FTSB_FN_SIN = 1, FTSB_FN_PULSE = 2, FTSB_FN_EXP = 3 };
/* row class — NodeType (src/node.rs:9-13) */
enum { FTSB_ROW_VOLTAGE = 0, FTSB_ROW_CURRENT = 1 };

/* per-instance outcome */
enum {
    FTSB_OK = 0,
    FTSB_NOT_CONVERGED = 1,    /* Err(NotConvergedError): 100 Newton iterations, newtons_method.rs:66-70 */
    FTSB_DIVERGED = 2,         /* panic "v_err or i_err diverged", newtons_method.rs:53-55              */
    FTSB_TRAN_H_UNDERFLOW = 3, /* Err(NotConvergedError): h < 1e-18, transient.rs:88-90                  */
    FTSB_NMOS_UNREACHABLE = 4  /* unreachable!() on a NaN terminal voltage, nmos/model.rs:38-40          */
};

/* call-level errors */
enum {
    FTSB_E_ARG = -1,        /* bad argument (message says which) */
    FTSB_E_CUDA = -2,       /* CUDA runtime error / no device     */
    FTSB_E_UNSUPPORTED = -3,/* circuit too large for the built kernels */
    FTSB_E_STATE = -4       /* call order (e.g. run before set_params) */
};

/* where the caller's buffers live */
enum {
    FTSB_MEM_HOST = 0,  /* host pointers; the call copies in/out and returns when results are in place */
    FTSB_MEM_DEVICE = 1 /* pointers the handle's device can dereference — device memory, or page-locked host memory
                           (unified addressing): with the latter .tran records STREAM into host memory while the
                           kernel runs, which is how waveforms larger than HBM leave the GPU; the call only enqueues
                           work on the handle's stream and returns; use ftsb_sync().  Exception: the first .op / .dc calls
                           of a handle on a small circuit learn elimination plans from a pilot run and compile
                           a per-topology kernel (about a second, once); those calls synchronise the stream */
};

/* layout of the per-instance parameter arrays handed to ftsb_set_params */
enum {
    FTSB_LAYOUT_INSTANCE_MAJOR = 0, /* vals[instance][elem]        fn[instance][fn_slot][7] */
    FTSB_LAYOUT_PARAM_MAJOR = 1     /* vals[elem][instance]        fn[fn_slot][7][instance] */
};

/* One element of the topology, index-lowered.  node[] is the Rust struct's `nodes` Vec (AFTER the
 * parser's terminal swaps, src/parser.rs:83-84,114-115,145-146,161-162,177-178), as unknown indices of
 * NodeCollection (voltage nodes first, then V-source rows, then — start-up only — inductor rows);
 * -1 = GND ("0") or unused. */
typedef struct ftsb_elem {
    int32_t kind;    /* FTSB_R .. FTSB_M */
    int32_t node[3]; /* R,C,V,I: {vneg, vpos}; L,D: {vpos, vneg}; Q: {c, b, e}; M: {d, g, s} */
    int32_t branch;  /* V: its current row; L: its start-up current row; others -1 */
    int32_t fn_kind; /* V, I: FTSB_FN_*; others 0 */
} ftsb_elem;

typedef struct ftsb_circuit ftsb_circuit; /* opaque */

/* .tran recording policy.  The reference keeps every accepted step (StateHistory,
 * src/engine/transient/state_history.rs:24-30); at 64k instances x 10k steps that is 333 GB, so the
 * caller chooses what is written out.  n_rec[] always counts ALL accepted steps. */
typedef struct ftsb_tran_opts {
    int64_t capacity;   /* rows available per instance in t/x/n_iters (0 = record nothing)        */
    int32_t rec_stride; /* store accepted step k (0-based) iff k % rec_stride == 0; <=1 = every one */
    int32_t reserved;
} ftsb_tran_opts;

/* work counters, summed over the instances of the last run (for flop/byte accounting) */
typedef struct ftsb_counters {
    int64_t solve_points;   /* rows the reference would print: .op rows, sweep points, accepted steps */
    int64_t newton_iters;   /* Newton iterations executed                                             */
    int64_t lu_factor;      /* dense LU factorisations executed                                       */
    int64_t lu_solve;       /* forward/back substitutions executed                                    */
    int64_t rej_newton;     /* .tran attempts rejected by Newton failure                              */
    int64_t rej_plte;       /* .tran attempts rejected by the PLTE test                               */
    int64_t redo;           /* instances the structure-exploiting kernel handed to the dense kernels     */
    int64_t sparse_flops;   /* flops executed by the structure-exploiting solver (factor + substitutions) */
} ftsb_counters;

/* ---- library ---- */
int ftsb_abi_version(void);
int ftsb_device_count(void);
/* text of the last error of `c` (or of the last failed ftsb_compile when c == NULL) */
const char *ftsb_last_error(const ftsb_circuit *c);

/* ---- Engine::new ----
 * n_run   = NodeCollection::from_elems(..).len()           (used by .dc and .tran)
 * n_start = NodeCollection::from_startup_elems(..).len()   (used by .op: inductors get a current row)
 * row_class[n_start] = FTSB_ROW_* per unknown.  device = CUDA ordinal. */
int ftsb_compile(const ftsb_elem *elems, int32_t n_elems, int32_t n_run, int32_t n_start,
                 const int32_t *row_class, int device, ftsb_circuit **out);
void ftsb_destroy(ftsb_circuit *c);

/* topology queries */
int32_t ftsb_n_elems(const ftsb_circuit *c);
int32_t ftsb_n_run(const ftsb_circuit *c);
int32_t ftsb_n_start(const ftsb_circuit *c);
int32_t ftsb_n_fn(const ftsb_circuit *c);   /* number of V/I elements with a waveform, in element order */
/* opt-in/out of skipping bit-identical refactorisations (linear circuits; default 1 = skip).
 * Results are identical either way; only ftsb_counters.lu_factor changes. */
int ftsb_set_option(ftsb_circuit *c, const char *name, int64_t value);

/* ---- per-instance parameters ----
 * vals: one f64 per element per instance (R ohm, C farad, L henry, V volt, I ampere; ignored for
 *       D/Q/M which have no parameters, src/device/diode.rs:27-33).  For V/I with a waveform the value is
 *       recomputed as f(0) like Engine::new does (src/engine.rs:47-51).
 * fn:   7 f64 per waveform source per instance, SpiceFn field order (src/spice_fn.rs:19-90):
 *       SIN offset, amplitude, freq | PULSE v1, v2, delay, t_rise, t_fall, pulse_width, period |
 *       EXP v1, v2, rise_delay, rise_tau, fall_delay, fall_tau.  May be NULL when ftsb_n_fn()==0.
 * `mem` says where vals/fn live.  The library keeps its own device copy. */
int ftsb_set_params(ftsb_circuit *c, int64_t batch, const double *vals, const double *fn, int layout, int mem);

/* ---- Engine::run_op ----  x[batch][n_start], n_iters[batch], status[batch] */
int ftsb_run_op(ftsb_circuit *c, double *x, int32_t *n_iters, int32_t *status, int mem);

/* ---- Engine::run_dc ----
 * Instance i sweeps element `sweep_elem` (a V or I source) over start[i] + k*step[i],
 * k < ceil((stop[i]-start[i])/step[i])  (ndarray Array::range, src/engine.rs:120), warm-starting each
 * point from the previous one and stopping at the first failure (src/engine.rs:116-135).
 * x[batch][max_points][n_run], n_iters[batch][max_points], points_done[batch], status[batch].
 * An instance with more than max_points points is an FTSB_E_ARG error (nothing runs). */
int ftsb_run_dc(ftsb_circuit *c, int32_t sweep_elem, const double *start, const double *stop, const double *step,
                int64_t max_points, double *x, int32_t *n_iters, int64_t *points_done, int32_t *status, int mem);

/* ---- Engine::new values + Engine::run_op / run_dc in one blocking call, HOST buffers only ----
 * Same results as ftsb_set_params(..., FTSB_MEM_HOST) followed by ftsb_run_op / ftsb_run_dc(..., FTSB_MEM_HOST)
 * (src/engine.rs:31-60 then :62-90 / :92-140).  Once a handle runs its per-topology kernel (small circuits, from
 * the third call on) the batch is processed in chunks on separate streams so that the host<->device copies of one
 * chunk overlap the solve of the others (option "pipe_chunks", default 4; 1 = serial); pass page-locked buffers
 * for the copies to be asynchronous.  Instance-major layout pipelines; parameter-major takes the serial path. */
int ftsb_solve_op(ftsb_circuit *c, int64_t batch, const double *vals, const double *fn, int layout, double *x,
                  int32_t *n_iters, int32_t *status);
int ftsb_solve_dc(ftsb_circuit *c, int64_t batch, const double *vals, const double *fn, int layout, int32_t sweep_elem,
                  const double *start, const double *stop, const double *step, int64_t max_points, double *x,
                  int32_t *n_iters, int64_t *points_done, int32_t *status);

/* ---- Engine::run_tran ----  (.TRAN stop step: start is always 0, src/parser.rs:246-250)
 * t[batch][capacity], x[batch][capacity][n_run], n_iters[batch][capacity] (any may be NULL),
 * n_rec[batch] = accepted steps, x_final[batch][n_run] = unknowns at the last accepted step (may be
 * NULL), status[batch]. */
int ftsb_run_tran(ftsb_circuit *c, double stop, double step_max, const ftsb_tran_opts *opts, double *t, double *x,
                  int32_t *n_iters, int64_t *n_rec, double *x_final, int32_t *status, int mem);

/* ---- several devices behind one handle ----
 * The instances of a batch are independent: the batch is cut into contiguous shards, one per device of
 * the list (devices == NULL or n_devices <= 0: every visible device; a device may be listed more than once), each
 * device has its own compiled copy of the topology and its own host thread (pinned to the cores of the GPU's NUMA node;
 * FTSB_NUMA_PIN=0 in the environment switches that off), and every shard's results are written straight into the caller's HOST
 * buffers at the shard's offset: the output layout is that of a single-device call, the device count is invisible.
 * No collective, no process group: this is what the Rust side of src/main.rs:29-44 calls on a multi-GPU box.
 * Instance-major parameters only.  ftsb_multi_solve_* = Engine::new values + Engine::run_op / run_dc / run_tran
 * (src/engine.rs:31-60, :62-90, :92-140, :142-206) for every instance, blocking. */
typedef struct ftsb_multi ftsb_multi; /* opaque */
int ftsb_multi_compile(const ftsb_elem *elems, int32_t n_elems, int32_t n_run, int32_t n_start, const int32_t *row_class,
                       const int32_t *devices, int32_t n_devices, ftsb_multi **out);
void ftsb_multi_destroy(ftsb_multi *m);
int32_t ftsb_multi_device_count(const ftsb_multi *m);
const char *ftsb_multi_last_error(const ftsb_multi *m);
const char *ftsb_multi_last_variant(const ftsb_multi *m);
int ftsb_multi_set_option(ftsb_multi *m, const char *name, int64_t value); /* forwarded to every device's handle */
int ftsb_multi_solve_op(ftsb_multi *m, int64_t batch, const double *vals, const double *fn, double *x, int32_t *n_iters,
                        int32_t *status);
int ftsb_multi_solve_dc(ftsb_multi *m, int64_t batch, const double *vals, const double *fn, int32_t sweep_elem,
                        const double *start, const double *stop, const double *step, int64_t max_points, double *x,
                        int32_t *n_iters, int64_t *points_done, int32_t *status);
int ftsb_multi_solve_tran(ftsb_multi *m, int64_t batch, const double *vals, const double *fn, double stop, double step_max,
                          const ftsb_tran_opts *opts, double *t, double *x, int32_t *n_iters, int64_t *n_rec, double *x_final,
                          int32_t *status);
int ftsb_multi_get_counters(ftsb_multi *m, ftsb_counters *out); /* summed over the devices */
int64_t ftsb_multi_launch_count(const ftsb_multi *m);

/* ---- plumbing ---- */
int ftsb_sync(ftsb_circuit *c);                          /* wait for the handle's stream */
void *ftsb_stream(ftsb_circuit *c);                      /* the cudaStream_t work is enqueued on */
int ftsb_set_stream(ftsb_circuit *c, void *cuda_stream); /* use the caller's stream instead */
int ftsb_get_counters(ftsb_circuit *c, ftsb_counters *out); /* syncs; counters of the last run */
int64_t ftsb_launch_count(const ftsb_circuit *c);        /* kernels launched by this handle so far */
/* name of the kernel variant the last run used, e.g. "warp16", "cta128" */
const char *ftsb_last_variant(const ftsb_circuit *c);

/* ---- standalone batched dense LU solve (src/engine/gauss_lu.rs:3-54) ----
 * Solves batch systems a[i] (n x n row-major) x = b[i]; pivoting rule, reciprocal scaling and
 * column-oriented back-substitution as the reference.  Used by the LU known-answer tests. */
int ftsb_lu_solve(int device, int32_t n, int64_t batch, const double *a, const double *b, double *x, int mem);

/* fp64 FMA micro-benchmark: returns measured TFLOP/s of dependent-chain-free DFMA on `device`
 * (the roofline denominator for the LU kernels, measured in the same run as the bench). */
double ftsb_measure_fp64_tflops(int device);

#ifdef __cplusplus
}
#endif
#endif /* FTSB_H */
