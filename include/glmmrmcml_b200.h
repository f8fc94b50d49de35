/* glmmrmcml_b200.h — C-ABI of the B200-native Monte-Carlo E-step + random-effect sampler of glmmrMCML.
 *
 * This is the drop-in boundary: the entry points below are what the reference's Rcpp
 * exports (src/RcppExports.cpp:291-305) bind to once their bodies are replaced by thin adapters; see
 * INTEGRATION.md for the adapter code.  Conventions:
 *   - every pointer is a HOST pointer owned by the caller; matrices are column-major double
 *     (what R / Eigen hand over), integer matrices column-major int32;
 *   - the library never keeps a host pointer past return; opaque handles own device memory;
 *   - every function returns 0 on success or a GMB_E* code; gmb_last_error() gives the message
 *     (thread-local).  No C++ exception crosses this boundary;
 *   - there is NO CPU fallback: without a CUDA device every call that computes fails with GMB_ECUDA.
 *
 * All paths cited are relative to the reference tree.
 */
#ifndef GLMMRMCML_B200_H
#define GLMMRMCML_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMB_OK        0
#define GMB_EINVAL    1   /* bad argument / size mismatch */
#define GMB_EFAMILY   2   /* unknown family+link (the reference throws std::out_of_range, mcmlmodel.h:89) */
#define GMB_ECUDA     3   /* CUDA runtime / no device */
#define GMB_ENOTPD    4   /* D(theta) not positive definite; message carries the pivot index */
#define GMB_ENCCL     5   /* NCCL failure */
#define GMB_ESTATE    6   /* call order (e.g. loglik before set_u) */
#define GMB_ECOV      7   /* unsupported covariance function id */
#define GMB_ENUMERIC  8   /* singular system in a Newton step */

typedef struct gmb_ctx   gmb_ctx;     /* one per process: device, streams, scratch, optional NCCL communicator */
typedef struct gmb_model gmb_model;   /* replaces glmmr::mcmlModel (inst/include/glmmrmcml/mcmlmodel.h:28-307) */
typedef struct gmb_cov   gmb_cov;     /* replaces glmmr::DData + glmmr::MCMLDmatrix (mcmldmatrix.h:18-79) */

const char* gmb_last_error(void);
const char* gmb_version(void);

/* ---- context ------------------------------------------------------------------------------------------- */
int  gmb_ctx_create(int device, gmb_ctx** out);
void gmb_ctx_destroy(gmb_ctx* ctx);
int  gmb_ctx_sync(gmb_ctx* ctx);
/* Number of kernels this library launched on this context so far (bench.py's "gpu_launches"). */
int64_t gmb_ctx_launch_count(gmb_ctx* ctx);
/* Raw cudaStream_t the kernels are launched on (so callers can bracket it with CUDA events). */
void* gmb_ctx_stream(gmb_ctx* ctx);
/* CUDA-event stopwatch on that stream: start records an event, stop records a second one, waits for it and returns
 * the elapsed device time in milliseconds. */
int gmb_ctx_timer_start(gmb_ctx* ctx);
int gmb_ctx_timer_stop(gmb_ctx* ctx, double* ms);
/* Overwrites a 256 MiB scratch buffer (larger than the 126 MB L2) on the stream: cold-cache timing of the next kernel. */
int gmb_ctx_flush_l2(gmb_ctx* ctx);

/* Multi-GPU: one process per GPU; Monte-Carlo samples and chains are sharded over ranks, the
 * per-evaluation sufficient sums are summed with ncclAllReduce on the context's stream.  The 128-byte NCCL
 * unique id is created on rank 0 and distributed by the caller (torch.distributed / MPI / R sockets). */
int gmb_comm_unique_id(void* id128);
int gmb_comm_init(gmb_ctx* ctx, const void* id128, int rank, int world);
int gmb_comm_rank(gmb_ctx* ctx, int* rank, int* world);
/* Sum `count` doubles over ranks, in place, host buffer (used by host-side callers; no-op when world == 1). */
int gmb_comm_allreduce_host(gmb_ctx* ctx, double* buf, int count);
/* Broadcast `count` doubles from rank 0 (the (beta, theta, sigma) broadcast of each MCML iteration). */
int gmb_comm_bcast_host(gmb_ctx* ctx, double* buf, int count);

/* ---- model: replaces glmmr::mcmlModel ------------------------------------------------------------------ */
/* mcmlModel ctor, mcmlmodel.h:51-98.  X is n x P, Z is n x Q, y length n.  family/link as in :74-87;
 * device kernels: codes 1-8 (1 poisson/log, 3 binomial/logit, 7 gaussian/identity on every kernel family; 2, 4, 5, 6, 8 on the general ones). */
int  gmb_model_create(gmb_ctx* ctx, int n, int P, int Q, const double* X, const double* Z, const double* y,
                      const char* family, const char* link, gmb_model** out);
/* The same with a storage precision for the streamed E-step matrices zd = Z u and (binomial/logit) F = exp(+-zd): 64 (what gmb_model_create
 * uses) or 32.  In fp32 mode they are held as float — the E-step kernels are HBM bound, so an evaluation moves 4 n m + 8 n bytes instead of
 * 8 n m + 16 n — while every accumulation, the per-row terms, the sampler and the covariance path stay fp64.  Results agree
 * with the fp64 mode to ~1e-7 relative (tolerance of the fp32 mode: 1e-5).  Family/link codes 1, 3, 7. */
int  gmb_model_create_prec(gmb_ctx* ctx, int n, int P, int Q, const double* X, const double* Z, const double* y,
                           const char* family, const char* link, int precision, gmb_model** out);
void gmb_model_destroy(gmb_model* mdl);
int  gmb_model_flink(gmb_model* mdl);

/* Set the Monte-Carlo sample matrix u (Q x m_local columns of THIS rank; m_total = columns over all ranks) and
 * build zd = Z u once (the reference rebuilds it on every evaluation, mcmlmodel.h:286).
 * niter_total = number of leading columns the E-step averages over (mcmlmodel.h:73 niter_; pass m_total unless
 * reproducing the Q x (m+1) quirk of mhmcmc.h:126,155 where niter_ = m). */
int gmb_model_set_u(gmb_model* mdl, const double* U, int Q, int m_local, int m_total, int niter_total);
/* Same, but the samples are already on the device from gmb_hmc_sample(..., keep_on_device=1). */
int gmb_model_use_device_u(gmb_model* mdl, int niter_total);

/* Copies `ncols` columns, starting at column col0, of this rank's device-resident sample matrix u (Q x m_local: what gmb_model_set_u
 * uploaded or gmb_hmc_sample(..., keep_on_device) left behind) into U_out (Q x ncols, column-major). */
int gmb_model_get_u(gmb_model* mdl, int col0, int ncols, double* U_out);

/* Re-forms zd = Z u from the device-resident samples (the work gmb_model_set_u does after its upload); for timing the contraction alone. */
int gmb_model_rebuild_zd(gmb_model* mdl);

/* E-step objective, mcmlModel::log_likelihood mcmlmodel.h:284-304 after update_beta(beta) (:100-102):
 * mean_j sum_i l(y_i, (X beta)_i + zd_ij ; var_par).  All-reduced over ranks. */
int gmb_model_loglik(gmb_model* mdl, const double* beta, double var_par, double* out);
/* Batched: evaluates n_eval parameter vectors (columns of beta_mat P x n_eval, var_par[n_eval]) back to back with
 * one device->host read at the end (the pattern of f_hess's 4k^2 stencil, mcmloptim.h:333-355). */
int gmb_model_loglik_batch(gmb_model* mdl, const double* beta_mat, const double* var_par, int n_eval, double* out);

/* MCNR sufficient sums and Newton step, mcmloptim::mcnr mcmloptim.h:198-236 (serial semantics):
 * xtwx (P x P) = mean_j X^T W_j X, score (P) = X^T mean_j Wu_j, beta_incr = xtwx^{-1} score, sigma = mean_j sd(resid_j).
 * Any output pointer may be NULL. */
int gmb_model_mcnr(gmb_model* mdl, const double* beta, double var_par,
                   double* xtwx, double* score, double* beta_incr, double* sigma);

/* ---- covariance: replaces glmmr::DData / MCMLDmatrix --------------------------------------------------- */
/* cov: rows x 5 int32 column-major [block, block dim, function id, n vars, first parameter index]
 * (src/mcml_optim.cpp:20-22); data, eff_range as produced by Covariance$get_D_data(). */
int  gmb_cov_create(gmb_ctx* ctx, const int32_t* cov, int rows, const double* data, int n_data,
                    const double* eff_range, int n_eff, gmb_cov** out);
void gmb_cov_destroy(gmb_cov* cv);
int  gmb_cov_dims(gmb_cov* cv, int* B, int* Q, int* R);
/* DMatrix::genD(0, chol, false): dense block-diagonal Q x Q D(theta) or its lower Cholesky factor
 * (call sites src/mcml_full.cpp:68,121).  L_out may be NULL to only factorise on the device. */
int gmb_cov_gen(gmb_cov* cv, const double* theta, int chol, double* L_out);
/* MCMLDmatrix::loglik(u), mcmldmatrix.h:23-41: (1/m) sum_b sum_j log N(u_j[b]; 0, D_b(theta)); U is Q x m_local host
 * columns of this rank (m_total over ranks).  All-reduced over ranks. */
int gmb_cov_mvn_ll(gmb_cov* cv, const double* theta, const double* U, int Q, int m_local, int m_total, double* out);
/* Same on the device-resident samples of a model (set by gmb_model_set_u / gmb_hmc_sample); ncols_total = how many
 * leading columns to average (m+1 in mcml_full, mcmldmatrix.h:24,40). */
int gmb_cov_mvn_ll_model(gmb_cov* cv, const double* theta, gmb_model* mdl, int ncols_total, double* out);
/* The same at the k columns of thetas (R x k): one launch and one synchronisation for the whole batch when every block is <= 16 (the
 * optimiser's stencils and the optimhess points of mcml_hess arrive as batches).  out[e] = -inf where D(theta_e) is not positive definite. */
int gmb_cov_mvn_ll_model_batch(gmb_cov* cv, const double* thetas, int k, gmb_model* mdl, int ncols_total, double* out);
/* MCMLDmatrix::logdet, mcmldmatrix.h:43-54. */
int gmb_cov_logdet(gmb_cov* cv, const double* theta, double* out);

/* ---- sampler: replaces glmmr::mcmc::mcmcRunHMC (mhmcmc.h:16-160) and the Stan programs in inst/stan -------- */
typedef struct gmb_hmc_stats {
    double accept_rate;      /* mean over chains of accept_/(warmup+nsamp), mhmcmc.h:152 */
    double step_size_mean;   /* mean final e_ over chains */
    double steps_mean;       /* mean leapfrog steps per proposal */
    double leapfrog_total;   /* total leapfrog steps over chains and proposals */
    double kernel_ms;        /* device time of the sampling kernel(s) */
    int    n_chains;
    int    nsamp_per_chain;
    int    rows_used;        /* rows the sampler ran on: n, or the number of distinct rows of [X | Z] when aggregated */
    int    kernel_variant;   /* 1 = two-GEMM, 2 = on-chip, 3 = structure-aware (sparse Z L) */
    double zl_nonzeros;      /* entries of Z L the kernel works on per leapfrog step and chain: non-zeros (variant 3) or rows_used * Q */
    int    component_groups; /* variant 3 on a large model: groups of connected components of Z L the trajectory is decomposed into (else 0) */
    int    factored;         /* variant 1 with Z applied in sparse form and L as the dense operand (Q x Q contractions instead of n x Q) */
    int    lane_components;  /* variant 3 on a small block-structured model: connected components of Z L, one lane each (else 0) */
} gmb_hmc_stats;

/* Runs n_chains independent copies of mcmcRunHMC::sample(warmup, .) (mhmcmc.h:121-157), each with its own
 * step-size adaptation (:107-117), as one batched kernel.  L is the Q x Q lower Cholesky factor of D (host),
 * xb = X beta is formed from beta.  Each chain yields nsamp_per_chain + 1 whitened states (column 0 = state after
 * warm-up, :142); U_out (Q x n_chains*(nsamp_per_chain+1), may be NULL) receives L v, chain-major.
 * V_out (same shape, may be NULL) receives the whitened states v.  RNG: Philox4x32-10 keyed by `seed`
 * with counter (idx, iteration, chain_offset + chain, stream) — reproducible, unlike mhmcmc.h:55.
 * keep_on_device != 0 keeps L v on the device as the model's sample matrix (then call gmb_model_use_device_u). */
int gmb_hmc_sample(gmb_model* mdl, const double* L, const double* beta, double var_par,
                   int warmup, int nsamp_per_chain, double lambda, int max_steps, double target_accept, int adapt,
                   int n_chains, uint32_t chain_offset, uint64_t seed, int keep_on_device,
                   double* U_out, double* V_out, gmb_hmc_stats* stats);

/* Sampler kernel selection: 0 = automatic (the structure-aware kernels when Z L is sparse enough — indicator Z, block-diagonal D;
 * else the on-chip kernel, Z L resident in shared memory, when the model fits one SM; otherwise two fused-epilogue GEMMs per
 * leapfrog step), 1 = force the two-GEMM variant, 2 = force the on-chip one, 3 = force the structure-aware one.
 * All variants follow the same chain arithmetic and the same random streams. */
int gmb_hmc_set_variant(int variant);

/* Structure-aware sampler on large models: 1 (default) = decompose the trajectory over the connected components of Z L (one warp per chain and
 * group of components, two launches per proposal), 0 = one CTA per chain streaming the sparse Z L on every leapfrog step. */
int gmb_hmc_set_components(int on);

/* Structure-aware sampler on small models: 1 (default) = when the view's Z L falls apart into at most 32 connected components of at most 6 rows
 * and 6 columns (cluster designs: one per cluster), a lane integrates a whole component in registers and a leapfrog step needs no exchange
 * between lanes (hmc_lane.cu); 0 = one warp per chain with the rows / columns spread over its lanes (hmc_sparse.cu). */
int gmb_hmc_set_lane(int on);

/* Two-GEMM sampler: 1 (default) = when Z is sparse, Z L is dense and n >= 2 Q, apply Z and L separately (W = L V', gather by the rows of Z,
 * residual, gather by its columns, G = L^T T: contractions of size Q x Q instead of n x Q); 0 = always contract with the dense n x Q matrix Z L. */
int gmb_hmc_set_factored(int on);

/* On-chip sampler: CTAs per group of 8 chains.  0 = automatic (per run: the cluster size with the shortest estimated
 * leapfrog step among those whose share of Z L fits one SM's shared memory), 1 = one CTA per group, 2 / 4 = the
 * observations of a group are split over a thread-block cluster of 2 / 4 SMs that exchange partial gradients through
 * distributed shared memory.  All settings follow the same chain arithmetic and random streams. */
int gmb_hmc_set_cluster_size(int cs);

/* E-step evaluation strategy for poisson/log and gaussian/identity: 1 (default) = O(n) evaluations from row statistics of zd built once per
 * sample matrix; 0 = stream zd on every evaluation (used by the roofline probes and the parity tests of the streaming kernel). */
int gmb_estep_set_rowstats(int on);

/* zd = Z u: 1 (default) = gather through the sparse form of Z when Z is sparse (indicator designs) and Q >= 64, 0 = always the dense contraction. */
int gmb_estep_set_sparse_zd(int on);

/* fp32 mode with a dense Z: 1 (default) = zd = Z u on the tensor cores (tcgen05.mma kind::tf32, operands split 3xTF32, fp32 accumulation in
 * tensor memory), 0 = the fp64 DMMA product narrowed to float. */
int gmb_estep_set_tf32(int on);

/* Batched binomial/logit evaluations (gmb_model_loglik_batch, batches of >= 8): 1 (default) = one launch for the whole batch, 8 parameter
 * vectors share each pass over the factor matrix; 0 = one launch per evaluation.  Both deterministic; they partition the sum differently, so
 * values agree to rounding (1e-13 relative), not bit for bit. */
int gmb_estep_set_multi(int on);

/* mvn_ll on a model's device-resident samples: 1 (default) = through sufficient statistics of the samples, built once per sample matrix —
 * when every covariance block is <= 16, their Gram matrices (each evaluation independent of the number of samples); for a large block
 * (> 64) on a single rank with at least twice as many samples as rows, the Cholesky factor of its Gram matrix (n^3 / 3 flop per
 * evaluation instead of n^2 m; falls back to the samples when that matrix is not positive definite); 0 = stream the samples on every
 * evaluation. */
int gmb_cov_set_gram(int on);

/* Gram-matrix path: 1 (default) = blocks that are IDENTICAL (same size, same function rows, same data — gr(cl)*ar1(t) repeats one block per
 * cluster; mcmldmatrix.h:26-36 loops over all of them) are built and factorised once per class, on the class's summed Gram matrix;
 * 0 = one factorisation per block.  Same sums in a different order.  gmb_cov_block_classes: number of distinct classes found. */
int gmb_cov_set_block_classes(int on);

/* E-step (log-likelihood, MCNR sums) of poisson/log, binomial/logit (0/1 responses) and gaussian/identity models: 1 (default) = when at most
 * a quarter of the rows of [X | Z] are distinct, zd = Z u is formed for the distinct rows only and every kernel runs on those rows with the
 * sufficient statistics of their observations (config C2: 500 rows -> 50; R6ModelExtMCML.R:283,297 densifies); 0 = one row per observation.
 * Same sums in a different order.  Needs the sampler's row view (gmb_hmc_set_row_aggregation(1), the default).
 * gmb_model_estep_rows: rows of the zd the model currently holds. */
int gmb_estep_set_row_aggregation(int on);

/* The reference-named entry points (gmb_mcmc_sample, gmb_mcml_optim, gmb_mcml_hess, ...) keep the device objects of their last few
 * (X, Z, y, family, link) and covariance specifications between calls — an R loop passes the same arrays again and again — and reuse them when
 * a byte-for-byte comparison of those arrays matches: 1 (default) on, 0 off (and drop what is kept).  Results do not depend on it. */
int gmb_set_object_cache(int on);
int gmb_model_estep_rows(gmb_model* mdl, int* rows);
int gmb_cov_block_classes(gmb_cov* cv, int* ncls);

/* On-chip sampler: 1 (default) = observations that share their row of [X | Z] (hence their linear predictor) are aggregated into one
 * weighted row (aggregate.cu; config C2: 500 rows -> 50), 0 = one row per observation.  Same sums in a different order. */
int gmb_hmc_set_row_aggregation(int on);

/* mcmlModel::log_prob / log_grad (mcmlmodel.h:138-153, 156-279, usezl = true) for C whitened states V (Q x C):
 * lp[C], grad (Q x C).  Either output may be NULL.  Used by the parity tests and by mcml_la. */
int gmb_model_logprob_grad(gmb_model* mdl, const double* L, const double* beta, double var_par,
                           const double* V, int C, double* lp, double* grad);

/* ---- host-side optimiser and finite-difference stencils: stand in for rminqa (Rbobyqa, Functor::Gradient/Hessian;
 * call sites mcmloptim.h:58-66,73-85,93-109,300-304,335-352).  The objective is BATCHED: it receives k points (columns of
 * X, n x k) and fills f[k], so that a GPU objective costs one synchronisation per batch instead of one per point.
 * It returns GMB_OK or an error code, which aborts the optimisation.  These three functions run on the host only. */
typedef int (*gmb_objective_batch)(const double* X, int n, int k, double* f, void* user);
/* Minimises f over [lower, upper] (NULL = unbounded) by projected BFGS on batched central-difference gradients.
 * rhobeg <= 0 selects BOBYQA's default min(0.95, 0.2 max|x0|).  x is updated in place. */
int gmb_minimize_bounded(gmb_objective_batch f, void* user, int n, double* x, const double* lower, const double* upper,
                         double rhobeg, double xtol, int maxit, double* fmin, int* nfev);
/* Bounded central differences with steps ndeps (rminqa Functor::Gradient, mcmloptim.h:296-317). */
int gmb_fd_gradient(gmb_objective_batch f, void* user, int n, const double* x, const double* ndeps,
                    const double* lower, const double* upper, int usebounds, double* grad);
/* optimhess stencil, 4 n^2 points in one batch (rminqa Functor::Hessian, mcmloptim.h:333-355); hess is n x n. */
int gmb_fd_hessian(gmb_objective_batch f, void* user, int n, const double* x, const double* ndeps,
                   const double* lower, const double* upper, int usebounds, double* hess, int* nfev);

/* ---- reference-named entry points (what the Rcpp exports forward to) -----------------------------------
 * These run on the process-wide default context: created lazily on the current CUDA device, or installed with
 * gmb_set_default_ctx (e.g. a context that gmb_comm_init joined to an NCCL communicator, so that mcml_full shards its
 * chains over the ranks).  The library does not take ownership of an installed context. */
int gmb_set_default_ctx(gmb_ctx* ctx);
/* Number of blocks B, total dimension Q and number of covariance parameters R (DData::n_cov_pars(), mcmloptim.h:26) of a
 * cov matrix; host arithmetic only, so adapters can size their outputs before calling an entry point. */
int gmb_cov_shape(const int32_t* cov, int rows, int* B, int* Q, int* R);

/* mvn_ll, src/mcml_optim.cpp:406-414 */
int gmb_mvn_ll(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
               const double* gamma, int n_gamma, const double* u, int Q, int m, double* out);

/* mcmc_sample, src/mcml_full.cpp:314-338: returns Q x (nsamp+1) samples (column 0 = state after warm-up).
 * n_chains <= 0 lets the library choose; n_chains = 1 reproduces the reference's single chain. */
int gmb_mcmc_sample(const double* Z, const double* L, const double* X, const double* y, const double* beta,
                    int n, int P, int Q, const char* family, const char* link, int warmup, int nsamp, double lambda,
                    double var_par, int trace, int refresh, int maxsteps, double target_accept,
                    int n_chains, uint64_t seed, double* samples_out);

/* mcml_optim, src/mcml_optim.cpp:35-68: one M-step on fixed samples u (Q x m).
 * start has length >= P + R (+1 for gaussian).  Outputs beta[P], theta[R], sigma. */
int gmb_mcml_optim(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                   const double* Z, const double* X, const double* y, const double* u, int n, int P, int Q, int m,
                   const char* family, const char* link, const double* start, int n_start, int trace, int mcnr,
                   double* beta_out, double* theta_out, double* sigma_out);

/* mcml_simlik, src/mcml_optim.cpp:90-117 (joint optimisation of beta, theta; F_likelihood likelihood.h:67-110). */
int gmb_mcml_simlik(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                    const double* Z, const double* X, const double* y, const double* u, int n, int P, int Q, int m,
                    const char* family, const char* link, const double* start, int n_start, int trace,
                    double* beta_out, double* theta_out, double* sigma_out);

/* Importance-weighted objective of mcml_simlik (F_likelihood with importance = true, likelihood.h:101-105): 0 (default) = evaluated in
 * log space, -(ll + logl - denomD); 1 = exactly as the reference writes it, -log(exp(ll + logl) / exp(denomD)), which is the same number
 * while exp() stays in range and -log(0 / 0) = NaN beyond (|ll + logl| > ~745: any model with more than a few hundred observations). */
int gmb_mcml_set_importance_form(int reference_form);

/* mcml_hess, src/mcml_optim.cpp:263-285: (P+R) x (P+R) finite-difference Hessian of F_likelihood (optimhess stencil). */
int gmb_mcml_hess(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                  const double* Z, const double* X, const double* y, const double* u, int n, int P, int Q, int m,
                  const char* family, const char* link, const double* start, int n_start, double tol, int trace,
                  double* hess_out);

/* aic_mcml, src/mcml_optim.cpp:356-392 */
int gmb_aic_mcml(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                 const double* Z, const double* X, const double* y, const double* u, int n, int P, int Q, int m,
                 const char* family, const char* link, const double* beta_par, int n_beta_par,
                 const double* cov_par, int n_cov_par, double* out);

/* mcml_full, src/mcml_full.cpp:41-148: the whole MCML loop with the native sampler; samples never leave the device
 * between the E- and M-step.  u_out is Q x (m+1) (may be NULL).  n_chains <= 0 lets the library choose. */
int gmb_mcml_full(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                  const double* Z, const double* X, const double* y, int n, int P, int Q,
                  const char* family, const char* link, const double* start, int n_start,
                  int mcnr, int m, int maxiter, int warmup, double tol, int verbose, double lambda, int trace,
                  int refresh, int maxsteps, double target_accept, int n_chains, uint64_t seed,
                  double* beta_out, double* theta_out, double* sigma_out, int* converged_out, int* iter_out,
                  double* u_out);

/* mcml_la, src/mcml_la.cpp:28-155: Laplace-approximation fit (la_optim over (beta, v), la_optim_cov over theta, final joint
 * la_optim_bcov; optional finite-difference standard errors, hess_la).  start holds (beta, theta, sigma), n_start >= P + R + 1.
 * se_out: n_start values (zeros unless usehess), u_out: Q values (u = L v).  m = 1: not a throughput path. */
int gmb_mcml_la(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                const double* Z, const double* X, const double* y, int n, int P, int Q, const char* family, const char* link,
                const double* start, int n_start, int usehess, double tol, int verbose, int trace, int maxiter,
                double* beta_out, double* theta_out, double* sigma_out, double* se_out, double* u_out, int* iter_out);
/* mcml_la_nr, src/mcml_la.cpp:178-290: the same loop with the Newton-Raphson step mcnr_b (mcmloptim.h:238-293) for (beta, v). */
int gmb_mcml_la_nr(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                   const double* Z, const double* X, const double* y, int n, int P, int Q, const char* family, const char* link,
                   const double* start, int n_start, int usehess, double tol, int verbose, int trace, int maxiter,
                   double* beta_out, double* theta_out, double* sigma_out, double* se_out, double* u_out, int* iter_out);
/* Parity hook for the Laplace path: out3 = (LA_likelihood(beta, v), LA_likelihood_cov(theta[, sigma]), LA_likelihood_btheta(beta, theta
 * [, sigma])) of likelihood.h:112-230 at the given state, with W formed at xb + Z v (w_use_l = 0, update_W()) or xb + Z L v
 * (w_use_l = 1); beta_nr / v_nr / sigma_nr (may be NULL) receive the state after one mcnr_b step (mcmloptim.h:238-293). */
int gmb_la_objectives(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                      const double* Z, const double* X, const double* y, int n, int P, int Q, const char* family, const char* link,
                      const double* beta, const double* theta, int R, const double* v, double sigma, int w_use_l,
                      double* out3, double* beta_nr, double* v_nr, double* sigma_nr);

#ifdef __cplusplus
}
#endif
#endif /* GLMMRMCML_B200_H */
