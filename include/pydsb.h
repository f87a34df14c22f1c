/* pydsb.h -- C ABI of libpyds_b200.so, the B200 (sm_100a) feasibility engine behind
 * pyds_b200.run.TwoStageManager / ThreeStageManager.
 *
 * The reference (cornelmarck/pyds) has NO FFI of its own: its hot path is a Python callable
 * g(d, p) handed to DEUS (reference run_twostage.py:145-165) that loops over design points and
 * launches one GAMS/CONOPT process per point (reference pyds/run.py:47-69, pyds/solver.py:50-55).
 * Each entry point below names the reference code it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative pydsb_status on error; the message is
 *     available from pydsb_last_error() (thread local);
 *   - no exceptions cross the boundary; no torch / C++ types in any signature;
 *   - unless the name ends in _host, array arguments are DEVICE pointers to contiguous row-major
 *     fp64 on the handle's device, and `stream` is a cudaStream_t passed as void* (NULL = legacy
 *     default stream); calls are asynchronous with respect to the host;
 *   - one handle per (process, device); a handle is not thread-safe.  The handles of one process share the
 *     program window in constant memory (every evaluation call re-uploads its model's program, stream ordered):
 *     calls on DIFFERENT handles of the same device must be issued on the same stream or be synchronised.
 *     On ONE handle the *_host entry points (which work on a private stream) are ordered after every earlier
 *     asynchronous call of that handle by an event; asynchronous calls on different streams are the caller's to order.
 */
#ifndef PYDSB_H
#define PYDSB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pydsb_status {
    PYDSB_OK = 0,
    PYDSB_ERR_INVALID = -1,     /* bad argument / malformed blob */
    PYDSB_ERR_CUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
    PYDSB_ERR_UNSUPPORTED = -3, /* model shape outside what the kernels are instantiated for */
    PYDSB_ERR_NO_DEVICE = -4    /* no CUDA device: there is deliberately no CPU fallback */
} pydsb_status;

/* control-vector sources for pydsb_eval (`z_mode`) */
enum {
    PYDSB_Z_MODEL = 0,        /* model defaults: warm-start profile / parameter-defined controls
                                 (reference run_twostage.py:25-26 F_in_default)                    */
    PYDSB_Z_SHARED = 1,       /* z is (nz): one profile for every scenario                          */
    PYDSB_Z_PER_DESIGN = 2,   /* z is (n_d, nz): shared across theta -- the reference's two-stage
                                 formulation (stage-0 F_in, reference pyds/utils.py:92-112)         */
    PYDSB_Z_PER_SCENARIO = 3  /* z is (n_d, n_t, nz): per-scenario recourse                         */
};

/* indices into the array returned by pydsb_model_info */
enum {
    PYDSB_INFO_NS = 0, PYDSB_INFO_NQ, PYDSB_INFO_NG, PYDSB_INFO_D_DIM, PYDSB_INFO_P_DIM, PYDSB_INFO_NZ,
    PYDSB_INFO_NFE, PYDSB_INFO_NCP, PYDSB_INFO_N_UNITS, PYDSB_INFO_SMEM_BYTES, PYDSB_INFO_BLOCKS_PER_SM,
    PYDSB_INFO_NUM_SMS, PYDSB_INFO_REGS_PER_THREAD, PYDSB_INFO_SCENARIOS_PER_THREAD, PYDSB_INFO_PROGRAM_LENGTH,
    PYDSB_INFO_CHAIN_BLOCKS,          /* blocks of the fused element loop of the feasibility program (0: interpreted)   */
    PYDSB_INFO_RECOURSE_CHAIN_BLOCKS, /* blocks of the chain evaluator of the control optimisation (0: sens_kernel)     */
    PYDSB_INFO_COUNT
};

/* indices into the array returned by pydsb_get_counters */
enum {
    PYDSB_CNT_SCENARIOS = 0,     /* (d,theta) evaluations completed                                  */
    PYDSB_CNT_NEWTON_ITERS,      /* Newton iterations until each scenario's own convergence          */
    PYDSB_CNT_NEWTON_WARP_ITERS, /* iterations actually executed (warp-uniform loop), x32 lanes      */
    PYDSB_CNT_FAILED,            /* scenarios whose state solve failed (counted infeasible; cf.
                                    reference pyds/solver.py:39-44, 63-69)                           */
    PYDSB_CNT_LAUNCHES,          /* kernel launches issued by this handle                            */
    PYDSB_CNT_NEWTON_FLOPS,      /* sum over scenarios and blocks of (own iterations x algorithmic flops
                                    of one Newton iteration of that block): the variable part of F_alg  */
    PYDSB_CNT_COUNT
};

const char* pydsb_last_error(void);
int pydsb_version(void);

/* Replaces Manager._build_model / utils.create_EF (reference pyds/run.py:32-34, pyds/utils.py:92-112):
 * instead of rebuilding a Pyomo block tree whenever n_p changes, the model is lowered once to a
 * tape blob (pyds_b200/model.py) and uploaded to the device here. */
int pydsb_model_create(const void* blob, size_t nbytes, int device, void** handle);
int pydsb_model_destroy(void* handle);
/* Host-only half of pydsb_model_create (needs no CUDA device): assembles the blob's tapes into the
 * pre-decoded per-scenario program feas_kernel interprets (csrc/assembler.h, csrc/feas_kernel.cuh) and copies up
 * to `cap` 32-bit words of it into `words`; *n_words = the full length.  info = { scenarios per thread,
 * register-file units, shared-memory bytes per CTA, program length in 16-byte words, blocks of the fused chain (0: the
 * element loop is interpreted), the two shape words of the chain, 1 when that shape has a compiled instantiation };
 * info[8 ...] (optional, n_info > 8): the chain descriptor word by word (csrc/poly_chain.cuh).  What it replaces in
 * the reference: the GAMS model file Pyomo writes for every solve (reference pyds/solver.py:50-55). */
int pydsb_assemble(const void* blob, size_t nbytes, uint32_t* words, size_t cap, size_t* n_words, long long* info,
                   int n_info);
int pydsb_model_info(void* handle, long long* out, int n);
/* Host-only: the bounding ellipsoid the nested-sampling proposals are drawn from ("suob-ellipsoid" sampling of the DEUS
 * activity form, reference run_twostage.py:182-200 [3P]).  live: (n, dim) design points, lo / hi: the design box.  In
 * box-scaled coordinates u = (d - lo) / (hi - lo): mu_out (dim) = mean, L_out (dim x dim, row-major, lower) = Cholesky
 * factor of the sample covariance (+ 1e-12 I; 0.25 I when n <= dim), *r_out = max_i |L^-1 (u_i - mu)|, so that
 * { mu + L z : |z| <= r } is the smallest ellipsoid of that shape holding every live point. */
int pydsb_ns_ellipsoid(const double* live, long long n, int dim, const double* lo, const double* hi, double* mu_out,
                       double* L_out, double* r_out);
/* Host-only: the replace-the-worst pass of one nested-sampling iteration (the live-point bookkeeping DEUS does around
 * the constraint function, reference run_twostage.py:167-180 `points_schedule` / nlive / nreplacements [3P]).  For each
 * candidate in order: if its probability cand_p[c] is above the worst live value -- or equal to it and at least `target`
 * -- it takes the slot of the worst live point (lowest slot among equal minima): live_p is updated in place and
 * slot_out[c] = that slot, else slot_out[c] = -1.  A candidate accepted early can be evicted by a later one. */
int pydsb_ns_replace_worst(double* live_p, long long n_live, const double* cand_p, long long n_cand, double target,
                           long long* slot_out);

/* The whole hot path in one launch sequence: state solve (collocation Newton), CQAs g, big-M
 * indicator and the weighted P(feasible) reduction over theta.
 * Replaces the body of TwoStageManager.g / ThreeStageManager.g for given controls
 * (reference pyds/run.py:55-67, 85-96: load_input -> simulate -> solver.solve -> -indicator) and
 * DEUS's efp reduction  P_i = sum_j w_j 1[all_k g_ijk >= 0].
 *   d      (n_d, d_dim)      theta (n_t, p_dim)      w (n_t) or NULL (= 1/n_t)
 *   z      per z_mode (may be NULL for PYDSB_Z_MODEL)
 *   g_out  NULL or (n_d, n_t, n_g)   constraint values g_k (feasible <=> all g_k <= 0)
 *   ind_out NULL or (n_d, n_t)       -indicator in {-0.0, -1.0}: exactly what Manager.g returns
 *   P_out  NULL or (n_d)             weighted probability of feasibility
 */
int pydsb_eval(void* handle, const double* d, long long n_d, const double* theta, const double* w,
               long long n_t, const double* z, int z_mode, double* g_out, double* ind_out, double* P_out,
               void* stream);

/* pydsb_eval with the classification rule of the relaxed big-M problem: a scenario is feasible when
 * g_k / M_k <= feas_tol for every k (big-M-scaled units; reference pyds/solver.py:71-80 snaps the relaxed indicator
 * `value == 0 -> 0 else 1`, and a solver returns an indicator on its bound as exactly 0 while the big-M rows hold to its
 * feasibility tolerance).  feas_tol = 0 is pydsb_eval.  Lets the managers reduce P(feasible) at the optimised controls
 * on the device instead of copying g to the host. */
int pydsb_eval_tol(void* handle, const double* d, long long n_d, const double* theta, const double* w, long long n_t,
                   const double* z, int z_mode, double feas_tol, double* g_out, double* ind_out, double* P_out, void* stream);

/* The state trajectories behind g: the converged collocation solution of every scenario, what the reference reads back
 * from the solved Pyomo model with utils.parse_value(scen, 'c') for its output records (reference pyds/solver.py:100-117,
 * pyds/utils.py:136-156) and what pyomo.dae.Simulator returns for the warm start (reference pyds/simulator.py:112-119).
 * x_out: (n_d, n_t, 1 + ncp * nfe, ns + nq) -- time points tau = 0, then the ncp Radau points of every finite element
 * in order; columns: the ns dynamic states, then the nq quadrature states (model order within each; pydsb_model_info
 * gives ns, nq, nfe, ncp).  failed_out (n_d, n_t), optional: 1.0 where the Newton solve failed, else 0.0.
 * z / z_mode as pydsb_eval.  Cold path: same kernels and program as pydsb_eval with the stores added. */
int pydsb_state_solve(void* handle, const double* d, long long n_d, const double* theta, long long n_t, const double* z,
                      int z_mode, double* x_out, double* failed_out, void* stream);

/* Convenience forms with the names SURVEY.md section 8b proposes. */
int pydsb_eval_g(void* handle, const double* d, long long n_d, const double* theta, long long n_t,
                 double* g_out, void* stream);
int pydsb_efp(void* handle, const double* d, long long n_d, const double* theta, const double* w,
              long long n_t, double* P_out, void* stream);

/* Same as pydsb_eval but with HOST pointers: stages through pinned buffers, copies in, runs,
 * copies the requested outputs back and synchronises.  This is the call the drop-in Manager.g
 * makes for numpy inputs (the reference's g(d, p) takes and returns numpy arrays). */
int pydsb_eval_host(void* handle, const double* d, long long n_d, const double* theta, const double* w,
                    long long n_t, const double* z, int z_mode, double* g_out, double* ind_out, double* P_out);

/* Control (recourse) optimisation: replaces Solver.solve / _solve_relaxation (reference
 * pyds/solver.py:26-55), i.e. the GAMS/CONOPT solve of the big-M relaxation
 *     max (1/N) sum_s (1 - y_s)  s.t.  g_k <= M_k y_s                  (pyds/utils.py:7-51)
 * in its reduced form  min_z sum_s w_s max(0, max_k g_ks(z)/M_k)  over the control box.
 *   per_scenario = 0: one control vector per design point, shared by all theta (the reference's
 *                     two-stage formulation); G = n_d groups
 *   per_scenario = 1: one control vector per (d, theta) scenario; G = n_d*n_t groups, w ignored
 *   per_scenario = 2: ONE control vector for the whole (d, theta) grid, G = 1 -- the reference's
 *                     three-stage formulation (stage-0 F_in above the design stage, pyds/run.py:76-91)
 *   z0      NULL (zeros) or (z0_rows, nz) start profile, z0_rows in {1, G}
 *   opts    NULL or 6 doubles {mu0, mu_min, lam0, step_tol, max_iter, check_every} (<= 0: default)
 *   z_out   (G, nz) optimised controls          phi_out  NULL or (G) objective value at z_out
 *   status_out NULL or (G): 1 all scenarios feasible, 2 converged, 3 iteration cap
 *   iters_out  NULL or (G) evaluations used
 * All pointers are DEVICE pointers; the call synchronises `stream` while it iterates. */
int pydsb_recourse_solve(void* handle, int per_scenario, const double* d, long long n_d, const double* theta,
                         const double* w, long long n_t, const double* z0, long long z0_rows, const double* opts,
                         double* z_out, double* phi_out, int* status_out, int* iters_out, void* stream);

/* The three-stage formulation on several GPUs: the ONE control vector couples every design point (reference
 * pyds/run.py:76-91), so with the design points sharded over processes each optimiser iteration sums the
 * {Phi_mu, Phi, grad, H} statistics of all processes.  `reduce(buf, n, user)` must all-reduce (sum) the n doubles at the
 * DEVICE pointer `buf` (= reduce_buf) over the processes and return 0 once the result is in place (NCCL / MPI in the
 * caller's hands; the stream is synchronised before the call).  Every process passes its own n_d_local rows (0 is
 * allowed), the same theta / w / z0 / opts, and receives the same z_out (nz), phi_out, status_out, iters_out (1 each).
 * reduce_buf: device buffer of at least 6 * (2 + nz + nz (nz + 1) / 2) doubles. */
typedef int (*pydsb_allreduce_fn)(double* device_buf, int n, void* user);
int pydsb_recourse_solve_global_dist(void* handle, const double* d, long long n_d_local, long long n_d_total,
                                     const double* theta, const double* w, long long n_t, const double* z0,
                                     const double* opts, double* z_out, double* phi_out, int* status_out, int* iters_out,
                                     double* reduce_buf, pydsb_allreduce_fn reduce, void* user, void* stream);

/* Counters since the last reset (device-side totals; synchronises the handle's work). */
int pydsb_get_counters(void* handle, unsigned long long* out, int n, int reset);

/* Per-kernel device time (CUDA events on the call's stream around every launch of the three compute kernels) since the
 * last reset -- what bench.py's roofline figures divide by.  Off by default (two event records per launch).
 * pydsb_get_profile synchronises the device; out[i] per the enum below (ms / launch counts). */
enum {
    PYDSB_PROF_FEAS_MS = 0, PYDSB_PROF_FEAS_LAUNCHES, PYDSB_PROF_SENS_MS, PYDSB_PROF_SENS_LAUNCHES, PYDSB_PROF_LM_MS,
    PYDSB_PROF_LM_LAUNCHES, PYDSB_PROF_COUNT
};
int pydsb_set_profiling(void* handle, int on);
int pydsb_get_profile(void* handle, double* out, int n, int reset);

/* Measured FP64 FMA throughput of the device (dependent-chain-free DFMA microkernel, CUDA-event
 * timed): the roofline denominator for this path.  Returns TFLOP/s in *tflops (2 flop per FMA). */
int pydsb_dfma_peak(int device, int repeats, double* tflops, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* PYDSB_H */
