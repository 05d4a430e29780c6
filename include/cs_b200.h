/* cs_b200.h -- C ABI of libcausalstump_b200.so
 *
 * B200-native (sm_100a) replacement for the RcppArmadillo hot path of
 * mazphilip/CausalStump.  This header is the drop-in boundary: every entry point
 * names the reference interface it replaces (file:line under /root/reference).
 * The reference's boundary is R's .Call into 11 registered routines
 * (src/RcppExports.cpp:187-205, R/RcppExports.R:4-46); INTEGRATION.md shows the
 * Rcpp / .Call glue that binds them to the functions below.
 *
 * Conventions
 *   - all matrices column-major fp64 (R / Armadillo layout), all pointers HOST memory
 *     unless stated otherwise; nothing is retained across stateless calls
 *   - return value: 0 = OK; >0 = matrix not positive definite, value is the 1-based
 *     failing pivot (LAPACK dpotrf `info`, what makes arma::inv_sympd throw at
 *     src/kernel_GP_SE_cpp.cpp:56); <0 = CS_ERR_* below, text in cs_last_error()
 *   - flat parameter vector = the reference's list order
 *       GP (kind 0): sigma, lambdam, lambdaa, Lm[p], La[p], mu        (2p+4)
 *       TP (kind 1): sigma, lambdam, nu, lambda0, Lm[p], La[p], mu    (2p+5)
 *     gradients use the same layout (clone(parameters), kernel_GP_SE_cpp.cpp:145)
 *   - there is NO CPU fallback: every call needs a CUDA device and fails with
 *     CS_ERR_NODEVICE otherwise
 */
#ifndef CS_B200_H
#define CS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CS_GP 0
#define CS_TP 1

#define CS_OPT_ADAM 0
#define CS_OPT_NADAM 1
#define CS_OPT_NESTEROV 2 /* "GD" is Nesterov with momentum 0 (R/main.R:68-70) */

#define CS_ERR_ARG (-1)
#define CS_ERR_CUDA (-2)
#define CS_ERR_NOMEM (-3)
#define CS_ERR_NODEVICE (-4)
#define CS_ERR_UNSUPPORTED (-5)
#define CS_ERR_NONFINITE (-6) /* |dLML| is NaN inside the optimisation loop: R stops at `if (change < tol ...)`, R/main.R:82 */

#define CS_MAX_P 128 /* p <= 32: tensor-path Gram / gradient kernels; p <= 43: element-wise kernels, one launch; beyond: the
                        * gradient kernel runs one launch per window of dimensions (its accumulators live in shared memory) */

typedef struct cs_fit cs_fit; /* opaque device-resident fit handle */

/* mirrors the arguments of CausalStump() (R/main.R:35) that reach the native code */
typedef struct cs_fit_config {
  int kind;        /* CS_GP / CS_TP          <- prior=                              */
  int optim;       /* CS_OPT_*               <- myoptim=                            */
  double lr;       /* learning_rate                                                 */
  double beta1;    /* beta1                                                         */
  double beta2;    /* beta2                                                         */
  double eps;      /* 1e-8, hard-coded at R/optimizer_classes.R:23,44               */
  double momentum; /* momentum (Nesterov / GD)                                      */
  int tile;        /* 0 = auto; 64 or 128 selects the dense tile edge               */
  int chunk;       /* optimizer iterations enqueued between host polls (0 = auto)   */
  int use_graph;   /* 0 = replay one captured iteration as a CUDA graph; -1 = eager  */
  int gemm_tile;   /* 0 / 64 = 64x64 CTA tiles (default); 128 = 128x128 CTA tiles    */
  int inv_pipeline;/* 1 = inverse pipelined behind the Cholesky on its own stream;   */
                   /* 0 = recursive triangular inverse + X^T X after the Cholesky    */
  int partition;   /* SM partition for the Cholesky's critical chain (CUDA green contexts):   */
                   /* 0 = auto (padded n >= 4096), 1 = on, -1 = off                            */
} cs_fit_config;

/* ---- library / device ---------------------------------------------------------- */
int cs_version(void);
const char* cs_last_error(void);
int cs_device_count(void);
int cs_set_device(int device);

/* ---- stateless mirrors of the 11 registered routines --------------------------- */

/* kernmat_GP_SE_cpp (src/kernel_GP_SE_cpp.cpp:72-140) and kernmat_TP_SE_cpp
 * (src/kernel_TP_SE_cpp.cpp:46-90; lambda_a carries lambda0).  X1 n1 x p, X2 n2 x p,
 * outputs n1 x n2; any of K/Km/Ka may be NULL. */
int cs_kernmat(int n1, int n2, int p, const double* X1, const double* X2, const double* Z1, const double* Z2,
               const double* Lm, const double* La, double lambdam, double lambda_a, double* K, double* Km,
               double* Ka);

/* invkernel_cpp (src/kernel_GP_SE_cpp.cpp:46-59): invK = inv_sympd(K + diag(exp(sigma)/w)).
 * logdet (nullable) receives log det of the noisy matrix. */
int cs_invkernel(int n, const double* w, const double* K, double sigma, double* invK, double* logdet);

/* mu_solution_cpp (src/kernel_GP_SE_cpp.cpp:61-70) */
int cs_mu_solution(int n, const double* y, const double* invK, double* mu);

/* grad_GP_SE_cpp (src/kernel_GP_SE_cpp.cpp:143-202) / grad_TP_SE_cpp
 * (src/kernel_TP_SE_cpp.cpp:100-163), same arguments in the same order.
 *   invK == NULL : Gram, Cholesky and inverse are computed on the device from
 *                  (X, z, w, params) -- the fast path; Kmat/Km/Ka must be NULL too.
 *   invK != NULL : used exactly as passed (log det taken from a Cholesky of it, the
 *                  reference uses an LU of it, :197).  Kmat/Km/Ka either all NULL
 *                  (recomputed from X, z, params -- they are pure functions of those at
 *                  every reference call site, R/kernel_GP_SE_class.R:50) or all given
 *                  (then read from memory exactly like the reference).
 * grad has the parameter layout; stats[2] = (squared residual, log marginal likelihood). */
int cs_grad(int kind, int n, int p, const double* y, const double* X, const double* z, const double* w,
            const double* Kmat, const double* Km, const double* Ka, const double* invK, const double* params,
            double* grad, double* stats);

/* stats_GP_SE (src/kernel_GP_SE_cpp.cpp:204-227) / stats_TP_SE (kernel_TP_SE_cpp.cpp:165-192):
 * mu = parameters$mu; nu = exp(parameters$nu) (TP only). */
int cs_stats(int kind, int n, const double* y, const double* Kmat, const double* invK, double mu, double nu,
             double* stats);

/* Adam_cpp / Nadam_cpp / Nesterov_cpp (src/optimizer_cpp.cpp:46-63, 28-44, 9-26) on flat
 * vectors of length nparam.  For Nesterov `m` is the velocity `nu`, `v` is unused and
 * beta1 carries the momentum.  Updates m, v, para in place. */
int cs_opt_step(int optim, int nparam, double iter, double lr, double beta1, double beta2, double eps, double* m,
                double* v, const double* grad, double* para);

/* ---- stateful fast path: KernelClass_*_SE + the loop of R/main.R:79-88 ---------- */

/* Uploads y, X (n x p), z, w once; init_params as produced by parainit
 * (R/kernel_GP_SE_class.R:10-22 / R/kernel_TP_SE_class.R:11-22). */
int cs_fit_create(const cs_fit_config* cfg, int n, int p, const double* y, const double* X, const double* z,
                  const double* w, const double* init_params, cs_fit** out);
void cs_fit_destroy(cs_fit* f);
int cs_fit_nparam(const cs_fit* f);

/* Batched handle: nbatch independent fits of the same (n, p, kind, optimizer) -- the replicates x
 * restarts of config 4 (test/sec5-3_Hill2011.R:611-616 runs them as forked R workers).  Every launch
 * of the evaluation / optimisation loop serves all members (one CTA slice per member), which removes
 * the launch-rate bound of n ~ 700 fits.  y, X, z, w, init_params are arrays of nbatch host pointers
 * (w, or any w[i], may be NULL = ones).  cs_fit_eval_async / cs_fit_run act on all members;
 * host transfers (set/get params, set_data, fetch, get_matrices, predict, treat) act on the member
 * chosen with cs_fit_select (initially 0).  cs_fit_run then fills iters[nbatch] and nbatch
 * consecutive 2 x (maxiter+2) trace blocks; maxiter <= 6000 for nbatch > 1. */
int cs_fit_create_batch(const cs_fit_config* cfg, int nbatch, int n, int p, const double* const* y, const double* const* X,
                        const double* const* z, const double* const* w, const double* const* init_params, cs_fit** out);
int cs_fit_batch_size(const cs_fit* f);
int cs_fit_select(cs_fit* f, int member);

/* One "LML + gradient evaluation" (SURVEY.md 8d): Gram -> Cholesky -> inverse -> fused
 * trace gradients -> LML/stats -> closed-form mu.  params NULL = current parameters.
 * No parameter update.  Outputs nullable. */
int cs_fit_eval(cs_fit* f, const double* params, double* grad, double* stats, double* mu_solution);

/* Same evaluation, device-resident: enqueue only (no host copies, no sync). */
int cs_fit_eval_async(cs_fit* f);
int cs_fit_sync(cs_fit* f);
int cs_fit_fetch(cs_fit* f, double* grad, double* stats, double* mu_solution);

/* The optimisation loop of R/main.R:79-88 entirely on the device: para_update per
 * iteration (gradient, optimizer step, closed-form mu, stopping rule |dLML| < tol &&
 * iter > 3), then the final get_train_stats.  stats_trace (nullable) is 2 x (maxiter+2)
 * column-major like `stats` in R/main.R:77; *iters receives the last iteration run. */
int cs_fit_run(cs_fit* f, int maxiter, double tol, int* iters, double* stats_trace);
/* Failure of one member (K_n not positive definite: the reference's inv_sympd throws, src/kernel_GP_SE_cpp.cpp:56; or a
 * non-finite LML) stops THAT member on the device; the others run to completion and all outputs are filled.  cs_fit_run /
 * cs_fit_run_many then return the first failing member's code (> 0 pivot, CS_ERR_NONFINITE) and status[nbatch] tells
 * which members are affected (0 = OK), so a replicate driver drops or retries only those units. */
int cs_fit_member_status(const cs_fit* f, int* status);
/* The same loop for several independent fits at once (replicates x restarts of config 4):
 * their iterations are interleaved on separate streams so the small kernels of one fit
 * fill the SMs another leaves idle.  iters[nfits]; stats_traces[i] nullable. */
int cs_fit_run_many(cs_fit** fits, int nfits, int maxiter, double tol, int* iters, double** stats_traces);

/* Replace the training data of an existing handle (same n, p): one pinned-staged
 * host->device copy of y, X, z, w (w NULL = ones). */
int cs_fit_set_data(cs_fit* f, const double* y, const double* X, const double* z, const double* w);

int cs_fit_get_params(cs_fit* f, double* params);
int cs_fit_set_params(cs_fit* f, const double* params);
/* Reset optimizer moments (initOpt, R/optimizer_classes.R:29-33) */
int cs_fit_reset_optimizer(cs_fit* f);
/* RefClass fields Kmat, Km, Ka, invKmatn (n x n each, nullable) for the CURRENT
 * parameters as left by the last eval / run. */
int cs_fit_get_matrices(cs_fit* f, double* Kmat, double* Km, double* Ka, double* invK);

/* KernelClass_*_SE$predict (R/kernel_GP_SE_class.R:74-85, R/kernel_TP_SE_class.R:72-94):
 * map[n2] = K_xX K^-1 (y-mu) + mu, var_diag[n2] = diag(cov), cov (nullable, n2 x n2)
 * = K_xx - K_xX K^-1 K_xX' + exp(sigma) I.  Normalised scale. */
int cs_predict(cs_fit* f, int n2, const double* X2, const double* z2, double* map, double* var_diag, double* cov);

/* KernelClass_*_SE$predict_treat (R/kernel_GP_SE_class.R:86-105, R/kernel_TP_SE_class.R:95-125):
 * treatment-interaction block with z2 = 1.  ate = mean(map); ate_var = sum(cov)/n^2 with
 * the TRAINING n (quirk Q8).  cov_jitter is added to diag(cov) (1e-6 for TP, :110). */
int cs_treat(cs_fit* f, int n2, const double* X2, double cov_jitter, double* map, double* var_diag, double* ate,
             double* ate_var, double* cov);

/* mvnfast::rmvt + per-column order statistic (R/kernel_TP_SE_class.R:84-93, 112-124) on
 * the device: N_rep = ceil(nsampling/1000) rounds of 1000 multivariate-t draws with
 * scale `cov` (n2 x n2) and df degrees of freedom; U[n2] = mean over rounds of the 900th
 * order statistic per column, *ate_U likewise for the row means (nullable). */
int cs_tp_sample(int n2, const double* cov, double df, int nsampling, unsigned long long seed, double* U,
                 double* ate_U);

/* ---- simulation driver on the device: the replicate loop of the reference's simulation scripts
 * (test/sec5-3_Hill2011.R: Hill11_surfaceB :49-66, run(i) :157-176, update.results :104-152; norm_variables R/utilities.R:1-33;
 * parainit R/kernel_GP_SE_class.R:10-20) for a list of (replicate, restart) units.  Response surface, 10 % test split,
 * normalisation and initial parameters are produced by one kernel per batch straight into the members of batched fit
 * handles; the fits run as cs_fit_run_many; predict(all rows) / treatment(train) / treatment(test) keep their inputs and
 * outputs on the device; one kernel per batch evaluates the 22 metrics.  Only the metric block comes back.
 * Random numbers: Philox-4x32-10 counters keyed by `seed` (host mirror: causalstump_b200/hill.py).  GP (prior=FALSE) arm. */
typedef struct cs_sim_config {
  int n, p;          /* rows / covariates of one replicate before the split (747 x 25)          */
  double test_frac;  /* 0.1: idx.test = sample.int(n, ceiling(n*0.1)), :171                     */
  int maxiter;       /* fit settings of the GP arm (:344-345: 3000, tol 1e-1, lr .01, Nadam)    */
  double tol, lr;
  int optim;         /* CS_OPT_*                                                                */
  int batch;         /* members per batched fit handle (0 = 32)                                 */
  unsigned long long seed;
} cs_sim_config;
#define CS_SIM_NMETRIC 22 /* row order of result.item, :163-164: 11 train rows then 11 test rows */
/* X (n x p, raw) and z (n) = covariates / treatment shared by every replicate, like the reference's fixed IHDP data; both
 * NULL: drawn per replicate with the IHDP marginals.  metrics[nunits][22]; iters, lml, status (0 OK, > 0 pivot of a non
 * positive definite K_n, CS_ERR_NONFINITE) nullable, per unit.  A failing unit fails alone (its metrics are NaN). */
int cs_sim_run(const cs_sim_config* cfg, int nunits, const int* replicate, const int* restart, const double* X, const double* z,
               double* metrics, int* iters, double* lml, int* status);
/* the generator alone for one unit (tests): every output nullable; Xtrain is ntr x p with ntr = n - ceil(n*test_frac) */
int cs_sim_generate(const cs_sim_config* cfg, int replicate, int restart, const double* X, const double* z, double* Xraw,
                    double* zraw, double* y, double* y1, double* y0, int* is_test, double* Xtrain, double* ytrain, double* ztrain,
                    double* Xall, double* moments, double* params);

/* ---- config 5: one large evaluation over a P x Q process grid (2-D block-cyclic tiles,
 * NCCL panel broadcasts; no reference equivalent) --------------------- */
typedef struct cs_dist cs_dist;
/* rank 0 obtains a 128-byte NCCL unique id and ships it to the other ranks by any means */
int cs_dist_unique_id(char* id128);
/* collective: every rank passes the same data (X, y, z, w are replicated: n(p+3) doubles)
 * and its own rank; world == P*Q, rank = pr*Q + pc. */
int cs_dist_create(int kind, int n, int p, const double* y, const double* X, const double* z, const double* w,
                   const double* params, int rank, int world, int P, int Q, const char* id128, cs_dist** out);
/* collective: one LML + gradient evaluation; every rank receives the same outputs;
 * *ms (nullable) = device time of this rank's stream for the evaluation. */
int cs_dist_eval(cs_dist* d, const double* params, double* grad, double* stats, double* mu_solution, float* ms);
void cs_dist_destroy(cs_dist* d);
/* the same driver with all P*Q ranks inside the calling process (collectives = device
 * copies): multi-rank algorithm test on one device.  T = 64 or 128 is the block size. */
int cs_dist_sim_eval(int kind, int n, int p, int T, int P, int Q, const double* y, const double* X, const double* z,
                     const double* w, const double* params, double* grad, double* stats, double* mu_solution);

/* ---- measurement helpers (bench.py) -------------------------------------------- */
/* CUDA events on the handle's stream; slots 0..15 */
int cs_fit_event_record(cs_fit* f, int slot);
int cs_fit_event_elapsed_ms(cs_fit* f, int slot_a, int slot_b, float* ms);
/* per-phase device time of one evaluation: gram, potrf, trtri, xtx, matvec, grad, final */
#define CS_NPHASE 7
int cs_fit_eval_profile(cs_fit* f, float phase_ms[CS_NPHASE]);
/* Dense phase (Cholesky + inverse: every launch an instance of the DMMA GEMM kernel or the leaf kernel) of the plan that is
 * actually being timed: while on, every cs_fit_eval_async / cs_fit_eval brackets that phase with a pair of CUDA events on the
 * handle's stream; cs_fit_dense_ms synchronises, returns the summed duration and the number of evaluations, and resets. */
int cs_fit_dense_timing(cs_fit* f, int on);
int cs_fit_dense_ms(cs_fit* f, float* total_ms, int* evaluations);
/* kernels launched by this library since load (claimed launch count for bench.py) */
long long cs_launch_count(void);
/* cs_predict / cs_treat keep up to 256 MiB of device scratch per host thread and device for the next call (a
 * replicate loop makes three such calls per fit); this releases it.  It is also released when the host thread exits. */
int cs_release_scratch(void);
/* test hook: happens-before check of the multi-stream Cholesky launch plan (0 = no races) */
int cs_debug_check_plan(int n, int tile, int gemm_tile, int verbose);
/* debug: per-launch timeline of the dense plan (rows of 6 floats: stream, type, tiles, K | block,
 * t_begin_ms, t_end_ms); used by tools/timeline.py to find idle gaps on the critical path */
int cs_debug_timeline(cs_fit* f, int cap, int* n_out, float* rows);
/* write `bytes` of device memory (> L2) to flush the cache between timed iterations */
int cs_flush_l2(cs_fit* f);

#ifdef __cplusplus
}
#endif
#endif /* CS_B200_H */
