/* macb200.h -- C-ABI of libmacb200.so: the B200 (sm_100a) Frank-Wolfe / Fiedler hot path of MAC.
 *
 * The reference (MarineRoboticsGroup/mac) is pure Python and has no FFI layer; its boundary for
 * this path is the Python call surface of `mac.solvers.mac.MAC` and `mac.optimization.frankwolfe`.
 * Each entry point below names the reference site it replaces (paths relative to the reference
 * root; `nx:` = networkx/linalg/algebraicconnectivity.py, the third-party module fiedler.py:42
 * calls).  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer argument is HOST memory owned by the caller;
 *   - every call returns an int status (MACB_OK = 0); `macb_last_error` gives the message;
 *   - a handle owns all device memory, one CUDA stream and its CUDA-graph cache; it is bound to
 *     one device and is NOT thread-safe (one handle per thread / GPU);
 *   - calls are synchronous: results are in the caller's buffers when the call returns;
 *   - all floating point is IEEE double; indices are int32; counts are int64.
 */
#ifndef MACB200_H
#define MACB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct macb_ctx* macb_handle;

enum {
    MACB_OK = 0,
    MACB_NOT_CONVERGED = 1,   /* eigen-iteration hit max_steps; outputs hold the best estimate   */
    MACB_ERR_ARG = -1,        /* bad argument (the reference raises AssertionError)              */
    MACB_ERR_CUDA = -2,       /* CUDA runtime error; message has the cudaError string            */
    MACB_ERR_STATE = -3,      /* call order violated (e.g. gradient before fiedler)              */
    MACB_ERR_NOMEM = -4
};

/* ---- lifetime -------------------------------------------------------------------------------- */

/* Replaces MAC.__init__ (mac/solvers/mac.py:22-72): uploads the fixed and candidate edge lists,
 * builds the fixed union sparsity pattern of L(w) = L_fixed + sum_k w_k kappa_k L_k once (CSR,
 * off-diagonals only; the diagonal is the weighted degree), and the slot -> edge map the assembly
 * kernel uses.  Self loops contribute nothing (graphs.py:77-96 nets them to zero) and are skipped.
 * The feasibility asserts of mac.py:46-52 are checked by the Python mirror, not here.
 * `device` < 0 keeps the current device. */
int macb_create(int32_t n,
                int64_t nf, const int32_t* fi, const int32_t* fj, const double* fw,
                int64_t m, const int32_t* ci, const int32_t* cj, const double* ckappa,
                int device, macb_handle* out);
int macb_destroy(macb_handle h);

/* Message for the last non-OK status on `h`; with h == NULL, the last macb_create failure. */
const char* macb_last_error(macb_handle h);

/* ---- Laplacian L(x) -------------------------------------------------------------------------- */

/* Replaces MAC.laplacian (mac.py:74-89) + weight_graph_lap_from_edges (graphs.py:58-98):
 * uploads x[m] and rewrites the CSR values of L(x) in place on the device; candidate k
 * contributes x_k kappa_k iff x_k > min_sel_tol (mac.py:85). */
int macb_set_x(macb_handle h, const double* x, double min_sel_tol);
int macb_get_x(macb_handle h, double* x);

/* y = L(x) v for host vectors of length n (replaces the scipy `L @ X` at nx:237); test / roofline. */
int macb_spmv(macb_handle h, const double* v, double* y);

/* max_i sum_j |L_ij| of the current L(x) (nx:229). */
int macb_lnorm(macb_handle h, double* lnorm);

/* ---- Fiedler pair ---------------------------------------------------------------------------- */

/* Start vector of the eigen-iteration (fiedler.py:27-32 seeds a fresh RandomState(7) block per
 * call; the Python mirror passes that block's first column).  NULL selects the built-in
 * deterministic generator.  The vector is copied. */
int macb_set_start(macb_handle h, const double* x0);

/* Replaces find_fiedler_pair (fiedler.py:9-44) -> networkx _tracemin_fiedler (nx:149-253) on the
 * current L(x).  Deflated Lanczos on the complement of the all-ones vector; stops when the TRUE
 * residual meets the reference's test ||L v - lambda v||_1 / ||L||_inf < tol (nx:243-245).
 * v has unit 2-norm and zero mean.  warm != 0 starts from the previous Fiedler vector.
 * Any output pointer may be NULL. */
int macb_fiedler(macb_handle h, double tol, int max_steps, int warm,
                 double* lambda2, double* v, int* steps, double* resid);

/* ---- gradient / LP step ---------------------------------------------------------------------- */

/* Replaces the gradient loop of MAC.problem (mac.py:117-124): g_k = kappa_k (v_i - v_j)^2 from the
 * Fiedler vector of the last macb_fiedler call.  g may be NULL (device-only). */
int macb_gradient(macb_handle h, double* g);

/* Replaces solve_subset_box_lp(g, k) (constraints.py:12-22 -> rounding.py:21-28) on the device
 * gradient: s[m] in {0,1} marks the k largest entries; ties at the k-th value go to the lowest
 * index.  s may be NULL. */
int macb_topk(macb_handle h, int64_t k, double* s);

/* Same LP oracle for an arbitrary host vector g[m] (no handle state involved). */
int macb_topk_dense(int device, const double* g, int64_t m, int64_t k, double* s);

/* macb_topk_dense / macb_round_nearest_dense keep one internal handle per (device, m) (at most four, least recently used first
 * out), so that a caller-side Frank-Wolfe loop (frankwolfe.py:60 calls solve_lp every iteration) does not pay a handle per call.
 * This releases them. */
void macb_dense_cache_clear(void);

/* Replaces round_nearest(w, k, weights=kappa, break_ties_decimal_tol=decimals) (rounding.py:30-42, called at
 * mac.py:207): the k largest entries in the lexicographic order (numpy.round(w, decimals), kappa); entries equal in
 * both keys are taken lowest index first (numpy's argpartition leaves that order unspecified).  w[m] in,
 * rounded[m] in {0,1} out.  The _dense form takes the tie-break weights explicitly and needs no handle. */
int macb_round_nearest(macb_handle h, const double* w, int64_t k, int decimals, double* rounded);
int macb_round_nearest_dense(int device, const double* w, const double* weights, int64_t m, int64_t k, int decimals,
                             double* rounded);

/* lambda2(L(x_b)) for nb iterates xs[nb][m], cold solves at `tol`, ONE host synchronisation for the whole batch (every solve
 * is enqueued behind the previous one; the stop decisions are taken on the device).  Replaces the loop of
 * `value_fn(x)` = MAC.evaluate_objective calls in round_madow (rounding.py:63-75, mac.py:203-204) and the 3-5
 * evaluate_objective calls per budget of examples/g2o_experiment.py:347-376.  resid (may be NULL) receives the residuals. */
int macb_evaluate_batch(macb_handle h, const double* xs, int nb, double tol, double min_sel_tol, int max_steps,
                        double* lambda2, double* resid);

/* ---- whole Frank-Wolfe loop ------------------------------------------------------------------- */

/* Replaces frank_wolfe (frankwolfe.py:10-79) specialised as MAC.solve calls it (mac.py:186-200):
 * problem = MAC.problem, solve_lp = top-k, step 2/(i+2), dual bound u = min(u, f + g.(s - x)),
 * stop tests ||g||_2 < grad_norm_tol and (u - f) < rel_gap_tol |f|.  x_init[m] in, w[m] out.
 * f_hist / u_hist (length >= max_iters, may be NULL) receive f and u per iteration. */
int macb_fw_run(macb_handle h, int64_t k, const double* x_init, int max_iters,
                double rel_gap_tol, double grad_norm_tol, double fiedler_tol, double min_sel_tol,
                int fiedler_max_steps, int warm,
                double* w, double* u, int* iters_done, double* f_hist, double* u_hist);

/* ---- measurement ----------------------------------------------------------------------------- */

enum {
    MACB_T_ASSEMBLE = 0, MACB_T_FIEDLER, MACB_T_GRADIENT, MACB_T_TOPK, MACB_T_UPDATE, MACB_T_COPY,
    MACB_T_COUNT
};
/* Cumulative counters since the last reset: kernels launched (graph nodes counted individually),
 * SpMV launches, Lanczos steps, Fiedler solves, and -- when profiling is on -- CUDA-event
 * milliseconds per phase (MACB_T_*).  Any pointer may be NULL. */
int macb_counters(macb_handle h, int64_t* kernel_launches, int64_t* spmv_launches,
                  int64_t* lanczos_steps, int64_t* fiedler_solves, double* phase_ms /*[MACB_T_COUNT]*/);
int macb_reset_counters(macb_handle h);
int macb_set_profile(macb_handle h, int on);

/* `reps` back-to-back launches of the SpMV kernel on the current L(x), timed with CUDA events on
 * the handle's stream; flush_l2 != 0 rewrites a > L2-sized buffer before every launch (each
 * launch then timed on its own).  Returns the average milliseconds per launch and the
 * algorithmic bytes of one launch. */
int macb_spmv_bench(macb_handle h, int reps, int flush_l2, double* avg_ms, double* algo_bytes);

/* Sizes of the device-resident problem: n, m, off-diagonal slots of the union pattern, and the
 * slots active (non-zero) in the current L(x). */
int macb_sizes(macb_handle h, int64_t* n, int64_t* m, int64_t* nnz_union, int64_t* nnz_active);

/* Writes a buffer larger than L2 (bench hygiene between timed steps). */
int macb_l2_flush(macb_handle h);

/* Bench mode for macb_fw_run: every FW iteration is bracketed by its own pair of CUDA events on the
 * handle's stream, and -- if flush_l2_between_iters != 0 -- a buffer larger than L2 is rewritten
 * between iterations, outside those brackets.  macb_iter_ms returns the per-iteration device
 * milliseconds of the last macb_fw_run (up to cap entries; *count = iterations recorded). */
int macb_set_bench(macb_handle h, int time_iters, int flush_l2_between_iters);
int macb_iter_ms(macb_handle h, double* ms, int cap, int* count);

/* Bench mode only (macb_set_bench time_iters != 0): cumulative CUDA-event time of the Lanczos kernel launches
 * (events recorded on the handle's stream around each launch), the Lanczos steps they ran, and the algorithmic
 * bytes of one step.  This is the dominant kernel of the path; bench.py derives its roofline line from it. */
int macb_lanczos_kernel_time(macb_handle h, double* ms, int64_t* phases, double* algo_bytes_per_phase);

/* Selects the kernel behind macb_spmv / macb_spmv_bench: 0 = k_spmv (CSR, W lanes per row; default), 1 = k_spmv_jds (chunked
 * jagged-diagonal layout with column-sorted slots, built on first use; for matrices far larger than L2 with locality it is bound
 * by HBM instead of by the gather rate).  Same result up to the order of summation inside a row.  Replaces the same site as
 * macb_spmv (`L @ X`, nx:237). */
int macb_spmv_engine(macb_handle h, int engine);

/* Name of the Lanczos kernel this handle launches (chosen from the graph's size at the first eigen-solve):
 * "k_lanczos_pipe" (default), "k_lanczos_small2" (graphs that fit one SM), "k_lanczos_slots" (chunked fall-back) or "k_lanczos_persist"
 * (CUDA-graph engine); "" before the first solve.  The pointer stays valid for the life of the handle. */
const char* macb_lanczos_kernel_name(macb_handle h);

/* CTAs (= SMs: 1024 threads and most of the shared memory each) one eigen-solve launch of this handle occupies, and the SMs of its
 * device: how many independent solves (budgets of a K-sweep, g2o_experiment.py:306) fit on the GPU side by side.  Builds the engine
 * (layout, buffers) if no solve has done so yet. */
int macb_lanczos_footprint(macb_handle h, int32_t* ctas, int32_t* sm_count);

/* Measured L2 -> SM read bandwidth of `device` (-1: current): every SM streams a `bytes`-sized, L2-resident buffer
 * `reps` times with 16-byte loads that bypass L1.  The Lanczos kernels' matrix is L2-resident at the BASELINE sizes, so
 * this -- not the HBM copy bandwidth -- is the roofline that binds them (bench.py `roofline.l2`).  Measurement only;
 * replaces nothing in the reference. */
int macb_measure_l2_bandwidth(int device, int64_t bytes, int reps, double* gbs);

/* On-device Rayleigh-Ritz (the stop decision of the Lanczos iteration taken by an extra CTA of the solver launch; replaces
 * the 4x4 `eigh` + residual test of nx:239-245): whether this handle uses it, how many eigen-solves since the last counter
 * reset fell back to the host-driven path, and status / order k / number of checks of the last solve. */
int macb_device_rr_stats(macb_handle h, int* enabled, int64_t* fallbacks, int* last_status, int* last_k, int* last_checks,
                         double* last_theta_est_target /* [13]: smallest Ritz value, residual estimate, its target, cycles the
                                                           Rayleigh-Ritz CTA waited / computed, coefficients published beyond k when
                                                           it decided, multisection rounds, cycles per stage (6); may be NULL */);

/* cudaDeviceSynchronize on the handle's device. */
int macb_device_sync(macb_handle h);

/* ---- K-sweep farm (multi-GPU) ----------------------------------------------------------------- */

/* The budget sweep of examples/g2o_experiment.py:284,306-336 has no loop-carried state: rank r of a group of processes (one
 * per GPU, every rank holding the same graph in its own handle) takes its share of the budgets, and ONE ncclAllGather of
 * fixed-size records collects the results on every rank.  NCCL is loaded at run time (dlopen of libnccl.so.2, or
 * $MACB_NCCL_LIB); the 128-byte unique id of macb_comm_unique_id (rank 0) reaches the other ranks by whatever out-of-band
 * channel the launcher has (mac_b200/farm.py: a TCP socket on MASTER_ADDR). */
typedef struct macb_comm* macb_comm_t;
int macb_comm_unique_id(char* id /*[128]*/);
int macb_comm_init(int nranks, int rank, const char* id /*[128]*/, int device, macb_comm_t* out);
/* send[bytes_per_rank] (host) from every rank -> recv[nranks * bytes_per_rank] (host) on every rank. */
int macb_comm_allgather(macb_comm_t c, const void* send, void* recv, int64_t bytes_per_rank);
int macb_comm_destroy(macb_comm_t c);
const char* macb_comm_last_error(macb_comm_t c);

/* owner[i] = rank that solves budget ks[i]: longest-processing-time-first on the cost estimate 1 + (m - k)/m (low budgets
 * run all their Frank-Wolfe iterations, high budgets exit early). */
int macb_sweep_owner(const int64_t* ks, int nk, int64_t m, int nranks, int32_t* owner);

/* For every budget ks[i], i < nk: MAC.solve(ks[i], x_inits[i], max_iters, rounding="nearest") as g2o_experiment.py:319 calls it
 * (mac.py:130-225), plus lambda2 of the relaxed solution (g2o_experiment.py:347).  comm == NULL: this process solves all of
 * them; otherwise every rank of `comm` calls this with the same arguments, solves the budgets macb_sweep_owner gives it and
 * receives all results.  x_inits[nk][m]; out: rounded[nk][m] (0/1), w[nk][m] (relaxed solutions, may be NULL), u[nk] (dual
 * bounds), lambda_unrounded[nk], iters[nk]. */
int macb_sweep(macb_handle h, macb_comm_t comm, const int64_t* ks, int nk, const double* x_inits, int max_iters,
               double rel_gap_tol, double grad_norm_tol, double fiedler_tol, double min_sel_tol, int fiedler_max_steps,
               uint8_t* rounded, double* w, double* u, double* lambda_unrounded, int32_t* iters);

/* ---- host-only helpers (no GPU needed; exported for the CPU test-suite) ----------------------- */

/* Host-side layout builder of the Lanczos kernel k_lanczos_pipe (32-row slices padded to their longest row with a lane stride of
 * 33 doubles, column-sorted slots that carry the position of their product, bank-fitted positions; see build_slice_layout in
 * csrc/api.cu for the format).  rp/col/eid: the union pattern of macb_host_build_pattern; CTA b owns rows row_start[b] ..
 * row_start[b+1] (at most 1024 of them, fewer than 2^15 product positions; n < 2^17 - 1).  Outputs: jrow[n], jlen[n], jcol[nnz],
 * jeid[nnz], jw[ncta*64] = (first position, entries per row) of every slice, positions[ncta] = product positions per CTA.
 * Replaces nothing in the reference (which rebuilds a CSR per iteration, graphs.py:58-98): exported so that the CPU test-suite
 * can check the layout's invariants without a GPU. */
int macb_host_build_slices(int32_t n, const int32_t* rp, const int32_t* col, const int32_t* eid, int32_t ncta, const int32_t* row_start,
                           int bankfit, int32_t* jrow, int32_t* jlen, int32_t* jcol, int32_t* jeid, int32_t* jw, int32_t* positions);

/* Smallest eigenpair of the symmetric tridiagonal T_k (diagonal a[0..k), off-diagonal b[1..k)):
 * bisection on the Sturm count + twisted factorisation.  This is the Rayleigh-Ritz step the
 * Lanczos driver runs on the host (the reference's counterpart is the 4x4 `eigh` at nx:239). */
int macb_tridiag_smallest(const double* a, const double* b, int k, double* theta, double* s);

/* Union-pattern CSR builder used by macb_create, host arrays out (row_ptr[n+1], col[nnz],
 * eid[nnz]; nnz = 2 * (#non-loop edges)); edge ids: fixed e -> e, candidate k -> nf + k. */
int macb_host_build_pattern(int32_t n, int64_t nf, const int32_t* fi, const int32_t* fj,
                            int64_t m, const int32_t* ci, const int32_t* cj,
                            int32_t* row_ptr, int32_t* col, int32_t* eid, int64_t* nnz);

const char* macb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MACB200_H */
