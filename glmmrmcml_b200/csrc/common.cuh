// common.cuh — internal declarations shared by the CUDA translation units of libglmmrmcml_b200.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <string>
#include <vector>
#include <cmath>

#include "../../include/glmmrmcml_b200.h"
#ifdef __CUDACC__
#include "exp_table.cuh"
#endif

// ------------------------------------------------------------------------------------------------
// errors (thread-local message, integer codes; nothing throws across the C boundary)
// ------------------------------------------------------------------------------------------------
int gmb_set_error(int code, const char* fmt, ...);

#define GMB_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return gmb_set_error(GMB_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define GMB_TRY(call)                       \
    do {                                    \
        int rc__ = (call);                  \
        if (rc__ != GMB_OK) return rc__;    \
    } while (0)

#define GMB_RESULT_DOUBLES 16384

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline size_t round_up_sz(size_t x, size_t m) { return (x + m - 1) / m * m; }

// Host-side phase timer for the entry points: with GMB_TRACE=1 in the environment every GMB_PHASE("name") prints the milliseconds since the
// previous one to stderr (the stream is synchronised first, so the cost lands on the phase that issued it); without it, one branch.
#include <chrono>
#include <cstdlib>
struct GmbPhase {
    bool on; cudaStream_t st; std::chrono::steady_clock::time_point t0;
    explicit GmbPhase(cudaStream_t s) : st(s) { static const bool e = getenv("GMB_TRACE") != nullptr; on = e; if (on) { cudaStreamSynchronize(st); t0 = std::chrono::steady_clock::now(); } }
    void mark(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[gmb trace] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// ------------------------------------------------------------------------------------------------
// host-side objects
// ------------------------------------------------------------------------------------------------
struct gmb_ctx {
    int device = 0;
    int sms = 148;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    // device scratch for partial sums / small results, and its pinned host mirror
    double* d_scratch = nullptr;
    size_t scratch_doubles = 0;
    double* h_pinned = nullptr;
    size_t pinned_doubles = 0;
    // fixed-address device buffers: small results (never reallocated) and the "last CTA done" counter
    double* d_result = nullptr;          // GMB_RESULT_DOUBLES doubles
    unsigned int* d_counter = nullptr;   // zero between kernels
    // NCCL (loaded with dlopen; see comm.cu)
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // internal (sampler kernel time)
    cudaEvent_t ev2 = nullptr, ev3 = nullptr;   // gmb_ctx_timer_start / _stop
    cudaStream_t stream2 = nullptr;             // high-priority side stream: panel factorisations running ahead of the trailing updates (cov_large.cu)
    cudaEvent_t evp = nullptr, evn = nullptr, evj = nullptr;   // panel-ready / narrow-update-done / join events of that look-ahead
    cudaStream_t stream3 = nullptr;             // second high-priority stream: work of the panel chain that is off its critical path (block inverses)
    cudaEvent_t evd[4] = {nullptr, nullptr, nullptr, nullptr}, evx = nullptr;   // diagonal-block-ready (per panel of an outer block) / inverse-ready
    void* d_flush = nullptr; int flush_val = 0; // gmb_ctx_flush_l2
};

int gmb_ctx_scratch(gmb_ctx* ctx, size_t doubles);   // grow d_scratch to at least `doubles`
// Device memory comes from the device's stream-ordered pool (cudaMallocAsync / cudaFreeAsync on the context's stream, release
// threshold = keep everything): the reference-named entry points build and tear down their device objects on every call, and
// a plain cudaFree costs between 1 ms and more than a second per call once the process holds larger allocations.
cudaError_t gmb_dmalloc_raw(gmb_ctx* ctx, void** p, size_t bytes);
void gmb_dfree(gmb_ctx* ctx, void* p);                // no-op for NULL
template <class T> static inline cudaError_t gmb_dmalloc(gmb_ctx* ctx, T** p, size_t bytes) { return gmb_dmalloc_raw(ctx, reinterpret_cast<void**>(p), bytes); }
int gmb_comm_allreduce_dev(gmb_ctx* ctx, double* dbuf, int count);   // in place on ctx->stream; no-op if world==1

// the on-chip sampler's view of a model: distinct rows of [X | Z] with weights (aggregate.cu); identity view when rows are distinct
struct gmb_agg {
    bool built = false, active = false, zl_valid = false;
    int flag = -1;               // gmb_agg_enabled() at build time
    int ng = 0, ldn = 0;         // rows of the view and their padded leading dimension
    double *dX = nullptr, *dZ = nullptr, *dZL = nullptr, *dxb = nullptr;   // aggregated copies (active only)
    double* dvec = nullptr;      // 6 x ldn: the arrays below
    double *dcnt = nullptr, *dys = nullptr;                               // residual: observations per row, response term
    double *dlcnt = nullptr, *dlys = nullptr, *dlsq = nullptr, *dlrc = nullptr;   // log-likelihood: count, response sum / mean, within-row SS, sum of lf(y)
    // E-step on the aggregated rows (estep.cu: *_agg kernels): response sum, within-row sum of squares about the row mean (every family),
    // row of each observation / representative observation of each row; binary_ok: every response of a binomial model is 0 or 1
    double *deys = nullptr, *dess = nullptr;
    int *dgid = nullptr, *drep = nullptr;
    bool binary_ok = true;
};

// sparse form of the sampler's Z L view (hmc_sparse.cu): row-wise and column-wise ELL, entry w of row i at [w * ngp + i]
// (of column j at [w * qp + j]), in ascending column (row) order, padded with zero values
struct gmb_ell {
    bool checked = false;        // the Z L the model currently holds has been analysed
    bool valid = false;          // ... and the structure-aware kernels apply (arrays below are filled)
    int ng = 0, Q = 0, ngp = 0, qp = 0, wr = 0, wc = 0;
    long long nnz = 0;
    double* rv = nullptr; int* rc = nullptr;     // wr x ngp
    double* cv = nullptr; int* cr = nullptr;     // wc x qp
    size_t r_cap = 0, c_cap = 0;                 // capacities (entries)
    int* dcnt = nullptr; size_t cnt_cap = 0;     // per-row / per-column counts
};

// connected components of the view's Z L packed into groups of <= 32 rows / columns (hmc_comp.cu)
struct gmb_comp {
    bool checked = false, valid = false;
    int G = 0, W = 0, ncomp = 0;
    int* dint = nullptr; size_t int_cap = 0;     // grow | gcol | lrc | lcr
    double* dval = nullptr; size_t val_cap = 0;  // lrv | lcv
};

// connected components of the view's Z L for the lane-per-component sampler (hmc_lane.cu): small block-structured models
struct gmb_lane {
    bool checked = false, valid = false;
    bool tri = false; int ncomp = 0, L = 0, slots = 0;                // components; lanes per component; rows = columns owned by a lane                    // components (lanes per chain); padded rows = columns per component (4 or 6)
    int* dint = nullptr; size_t int_cap = 0;     // rowid | colid
    double* dval = nullptr; size_t val_cap = 0;  // dense blocks
};

struct gmb_model {
    gmb_ctx* ctx = nullptr;
    int n = 0, P = 0, Q = 0, flink = 0;
    int ldn = 0;                 // padded leading dimension of n-row matrices (multiple of 4)
    int ldq = 0;                 // padded leading dimension of Q-row matrices (multiple of 4)
    double* dX = nullptr;        // ldn x P
    double* dZ = nullptr;        // ldn x Q
    double* dy = nullptr;        // ldn
    double* drowc = nullptr;     // ldn : per-row constant of the family term (poisson: log y! approx)
    double* dxb = nullptr;       // ldn : X beta of the last evaluation
    double* dbeta = nullptr;     // beta_cap doubles
    int beta_cap = 0;
    // samples (this rank's columns)
    double* dU = nullptr;        // ldq x m_cap
    int prec = 64;               // 64: zd / F stored as double; 32 (gmb_model_create_prec): as float — half the E-step bytes, arithmetic stays fp64
    float* dzd32 = nullptr;      // ldn x m_cap (fp32 mode)
    float* dF32 = nullptr;
    float *dZt_hi = nullptr, *dZt_lo = nullptr;   // fp32 mode, dense Z: K-major 3xTF32 split of Z (n rows of ldk floats), built once (gemm_tf32.cu)
    float *dU_hi = nullptr, *dU_lo = nullptr; size_t u32_cap = 0;   // ... and of the current sample matrix (m rows of ldk floats)
    double* dzd = nullptr;       // ldn x m_cap
    double* dF = nullptr;        // ldn x m_cap, binomial/logit only: exp(s_i zd_ij), s_i = -1 (y_i = 1) / +1 (y_i = 0); see estep.cu
    bool f_valid = false;
    double* dstat = nullptr;     // 2 x ldn row statistics of zd for the poisson / gaussian E-step (estep.cu: ensure_rowstats)
    bool stat_valid = false; int stat_cols = 0;
    int m_cap = 0;
    int m_local = 0, niter_local = 0, m_total = 0, niter_total = 0;
    bool zd_valid = false;
    unsigned long long u_version = 1;   // bumped whenever dU changes (caches keyed on the sample matrix: cov.cu Gram matrices)
    // sampler state
    double* dZL = nullptr;       // ldn x Q   (Z L)
    double* dL = nullptr;        // ldq x Q
    bool zl_valid = false;
    bool l_lower = false;        // the factor uploaded last is lower triangular (verified on the host): its zero k tiles are skipped in the contractions
    double* dV = nullptr;        // whitened samples of the last gmb_hmc_sample, ldq x v_cap
    size_t v_cap = 0;
    gmb_agg agg;                 // row aggregation for the on-chip sampler
    bool eagg = false;           // zd (and the E-step kernels) run on the agg.ng distinct rows of [X | Z], leading dimension agg.ldn (model.cu)
    gmb_ell ell;                 // sparse form of the view's Z L (structure-aware sampler)
    gmb_comp comp;               // its connected components (large sparse models)
    gmb_lane lane;               // ... and (small block-structured models)
    gmb_ell zell;                // sparse form of Z itself (factored sampler: Z sparse, L dense)
    double* hmc_work = nullptr;  // chain state + work buffers of the sampler
    size_t hmc_work_doubles = 0;
};

struct CovFn { int id, nvar, par0, col0; double eff; };
struct CovBlock { int n, start, fn0, nfn, ncol, all_gr; long long data0; long long l0; };
// leading dimension of a block's factor inside gmb_cov::d_Lblk: packed for warp-sized blocks, padded otherwise
#define GMB_COV_SMALL_MAX 32
__host__ __device__ static inline int gmb_cov_ld(int n) { return n <= GMB_COV_SMALL_MAX ? n : (n + 3) / 4 * 4; }

struct gmb_cov {
    gmb_ctx* ctx = nullptr;
    int B = 0, Q = 0, R = 0, max_n = 0;
    std::vector<CovBlock> blocks;
    std::vector<CovFn> fns;
    CovBlock* d_blocks = nullptr;
    CovFn* d_fns = nullptr;
    double* d_data = nullptr;
    double* d_theta = nullptr;
    double* d_Lblk = nullptr;    // concatenated per-block factors (col-major n_b x n_b each), offsets CovBlock::l0
    long long lblk_doubles = 0;
    double* d_logdet = nullptr;  // per-block sum of 2 log L_ii
    int* d_status = nullptr;     // first non-PD pivot (global row index + 1), 0 if fine
    std::vector<double> theta_cached;
    bool factor_valid = false;
    // buffers for host-provided U
    double* dU = nullptr; size_t dU_doubles = 0;
    // large-block workspace
    double* d_work = nullptr; size_t work_doubles = 0;
    // inverses of the 64 x 64 diagonal blocks of the large blocks' factors (cov_large.cu), 64 x 64 col-major each
    double* d_linv = nullptr; size_t linv_doubles = 0;
    std::vector<long long> linv_off;     // per block: offset into d_linv (-1 for blocks that take the small / medium path)
    // inverses of the 512 x 512 diagonal blocks of the large blocks' factors (cov_large.cu: formed beside the factorisation, used by the
    // forward substitution), 512 x 512 col-major each
    double* d_x512 = nullptr; std::vector<long long> x512_off;
    // sufficient statistics of a model's samples for the LARGE blocks (cov_large.cu: gmb_cov_gram_large): C_b = chol(U_b U_b^T), so that
    // sum_j ||L_b^-1 u_bj||^2 = ||L_b^-1 C_b||_F^2 costs n_b^3 / 3 flop per theta instead of n_b^2 m
    struct GramLarge { double* C = nullptr; double* linv = nullptr; int state = 0; /* 0 none, 1 valid, -1 not usable */ };
    std::vector<GramLarge> gram_large;
    const gmb_model* gramL_model = nullptr; unsigned long long gramL_version = 0; int gramL_cols = 0;
    // Gram matrices of a model's samples (cov.cu: cov_ensure_gram), laid out like d_Lblk
    double* d_batch = nullptr; size_t batch_bytes = 0;     // work area of gmb_cov_mvn_ll_model_batch
    double* d_gram = nullptr; const gmb_model* gram_model = nullptr; unsigned long long gram_version = 0; int gram_cols = 0;
    // classes of IDENTICAL blocks (same size, function rows and data: gr(cl)*ar1(t) repeats one block per cluster): the Gram
    // path builds and factorises one block per class on the class's summed Gram matrix
    int ncls = 0;
    int *d_cls_rep = nullptr, *d_cls_ptr = nullptr, *d_cls_mem = nullptr;   // representative block / CSR member lists
    double* d_gram_cls = nullptr;                                           // class sums, stored at the representative's l0
};

// ------------------------------------------------------------------------------------------------
// kernels' host launchers (defined in the respective .cu files)
// ------------------------------------------------------------------------------------------------
// gemm_f64.cu : C (M x N, ldc) = alpha * op(A) * B + beta * C ; A is M x K col-major (transA=0) or K x M (transA=1)
int gmb_dgemm(gmb_ctx* ctx, int transA, int transB, int M, int N, int K, double alpha, const double* A, int lda,
              const double* B, int ldb, double beta, double* C, int ldc);
int gmb_dgemm_tri(gmb_ctx* ctx, int transA, int transB, int M, int N, int K, double alpha, const double* A, int lda,
                  const double* B, int ldb, double beta, double* C, int ldc, int lower_a);
int gmb_dsyrk_lower_sub(gmb_ctx* ctx, int M, int K, const double* Pm, int ldp, double* C, int ldc, int c0, int c1);   // C[:, c0:c1) -= P P^T, lower tiles
int gmb_dsyrk_lower_rest(gmb_ctx* ctx, int M, int K, const double* Pm, int ldp, double* C, int ldc, int skip, int max_ctas);
int gmb_dgemm_rtri(gmb_ctx* ctx, int M, int N, const double* A, int lda, const double* T, int ldt, double* C, int ldc);
bool gmb_gemm_tma_available();
int gmb_dgemm_rowpanel_small(gmb_ctx* ctx, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double* C, int ldc);
int gmb_dgemm_nt_small(gmb_ctx* ctx, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C, int ldc);
int gmb_dsyrk_lower_small(gmb_ctx* ctx, int M, int K, const double* Pm, int ldp, double* C, int ldc);
int gmb_dgemm_rowpanel(gmb_ctx* ctx, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double* C, int ldc);   // C may alias A
int gmb_dgemm_colpanel(gmb_ctx* ctx, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double* C, int ldc);   // C may alias B

// gemm_tf32.cu: tcgen05 (kind::tf32, 3xTF32 split) contraction of the fp32 mode
bool gmb_tf32_enabled();
int gmb_split_tf32(gmb_ctx* ctx, int rows, int cols, int ld, const double* src, int transpose, int ldo, float* hi, float* lo);
int gmb_sgemm3_tf32(gmb_ctx* ctx, int M, int N, int K, const float* Ahi, const float* Alo, int lda, const float* Bhi, const float* Blo, int ldb,
                    float* Cm, int ldc);

// estep.cu
int gmb_launch_xb(gmb_model* mdl, const double* d_beta, double* d_xb);
int gmb_launch_loglik(gmb_model* mdl, const double* d_beta, double var_par, double* d_out /* 1 double: sum over local cols */);
int gmb_launch_loglik_cols(gmb_model* mdl, const double* d_beta, double var_par, const double* d_zd, int ncols, double* d_out);
int gmb_launch_mcnr(gmb_model* mdl, const double* d_xb, double var_par, double* d_out /* P*P + P + 1 doubles: local sums */);
int gmb_launch_build_factor(gmb_model* mdl, int ncols);
#define GMB_LOGLIK_NB 8
int gmb_launch_loglik_multi(gmb_model* mdl, const double* d_beta, int n_eval, double* d_out, int* done);
int gmb_estep_rowstats_enabled();

// cov.cu
int gmb_cov_factor(gmb_cov* cv, const double* theta);   // builds + factorises all blocks on the device
int gmb_cov_quad(gmb_cov* cv, const double* dU, int ldu, int ncols, double* d_out /* 1 double: sum_j sum_b l_b(u_j) */, gmb_model* gram_mdl = nullptr);
int gmb_dsyrk_lower_set(gmb_ctx* ctx, int M, int K, const double* Pm, int ldp, double* C, int ldc);   // lower tiles of C = P P^T
int gmb_cov_gen_device(gmb_cov* cv, const double* theta, int chol, double* d_out, int ld);   // dense D(theta) or chol D on the device
// cov_large.cu: in-place blocked Cholesky of a raw device matrix (see the definition)
int gmb_chol_blocked(gmb_ctx* ctx, double* A, int ld, int n, int row_offset, int* d_status, double* linv, double* d_logdet, double* x512 = nullptr);
size_t gmb_chol_linv_doubles(int n);   // size of the `linv` argument for an n x n matrix

// model.cu
int gmb_model_reserve_samples(gmb_model* mdl, int m);
int gmb_model_build_zd(gmb_model* mdl);
int gmb_solve_small(int P, const double* A, const double* b, double* x);

// hmc.cu
int gmb_hmc_prepare(gmb_model* mdl, const double* L_host);   // uploads L and forms ZL = Z L
// aggregate.cu
int gmb_agg_ensure(gmb_model* mdl);
void gmb_agg_free(gmb_model* mdl);
int gmb_agg_enabled();
// hmc_sparse.cu
void gmb_ell_free(gmb_model* mdl);
int gmb_ell_ensure(gmb_model* mdl);
int gmb_zell_ensure(gmb_model* mdl);
// hmc_comp.cu
void gmb_comp_free(gmb_model* mdl);
int gmb_comp_ensure(gmb_model* mdl);
// hmc_lane.cu
void gmb_lane_free(gmb_model* mdl);
int gmb_lane_ensure(gmb_model* mdl);
// the sparse forms describe the view's current Z L: call whenever it changes
static inline void gmb_sparse_invalidate(gmb_model* mdl) {
    mdl->ell.checked = mdl->ell.valid = false; mdl->comp.checked = mdl->comp.valid = false; mdl->lane.checked = mdl->lane.valid = false;
}

// optim.cpp: gmb_minimize_bounded / gmb_fd_gradient / gmb_fd_hessian are declared in the public header

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

#define GMB_PI_FAMILY 3.141593   /* moremaths.h:21,76 write pi like this */

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum; result valid in thread 0.  `red` must hold >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    int nw = (blockDim.x * blockDim.y * blockDim.z + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (l < nw) ? red[l] : 0.0;
        v = warp_sum(v);
    }
    return v;
}

// exp(x) for the streaming kernels.  Same algorithm class as the library exp (Cody-Waite reduction by ln 2, polynomial,
// exponent insertion; error < 1 ulp on the fast path) but with every constant taken from the constant bank as a direct
// DFMA operand: the library version materialises each 64-bit coefficient with two UMOV instructions, which made uniform
// moves 34% of the issued instructions of the E-step kernel (profiles/r01_ncu_summary.md).
static __constant__ double GMB_EXPC[20] = {
    1.4426950408889634074,          // 0: log2(e)
    6755399441055744.0,             // 1: 1.5 * 2^52 (round-to-nearest-integer by addition)
    -6.93147180369123816490e-01,    // 2: -ln2 high part (trailing zeros: k * ln2_hi is exact)
    -1.90821492927058770002e-10,    // 3: -ln2 low part
    1.0 / 6227020800.0,             // 4: 1/13!
    1.0 / 479001600.0,              // 5: 1/12!
    1.0 / 39916800.0,               // 6
    1.0 / 3628800.0,                // 7
    1.0 / 362880.0,                 // 8
    1.0 / 40320.0,                  // 9
    1.0 / 5040.0,                   // 10
    1.0 / 720.0,                    // 11
    1.0 / 120.0,                    // 12
    1.0 / 24.0,                     // 13
    1.0 / 6.0,                      // 14
    0.5,                            // 15
    1.0,                            // 16
    700.0,                          // 17: beyond this the library exp handles overflow / subnormal results
    0.0, 0.0};

__device__ __forceinline__ double dev_exp(double x) {
    const double* c = GMB_EXPC;
    double t = fma(x, c[0], c[1]);
    const int k = __double2loint(t);
    t -= c[1];
    double r = fma(t, c[2], x);
    r = fma(t, c[3], r);
    double p = c[4];
#pragma unroll
    for (int i = 5; i <= 16; i++) p = fma(p, r, c[i]);
    p = fma(p, r, c[16]);
    double res = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    if (!(fabs(x) < c[17])) res = exp(x);
    return res;
}

// Table-driven exp for the sampler's inner loop: exp(x) = 2^e * T[j] * exp(r) with x = (64 e + j) ln2/64 + r,
// |r| <= ln2/128, T = 2^(j/64) (64 doubles staged in shared memory from GMB_EXP2_TAB) and a degree-5 polynomial for
// exp(r) - 1 (truncation error r^6/720 < 4e-17).  10 FP64-pipe operations and no branch, against ~18 + a range branch for
// the library exp; error within 1.2 ulp of the exact value (tests/test_device_math_constants.py).  The binary exponent is saturated at +-1008 (e^+-698.7) with integer min/max: beyond
// that the family residuals that use it are already saturated (1/(1 + e^698) + y - 1 == y - 1 in double precision);
// valid for |x| < 2e7 (the integer part must fit 32 bits), NaN for NaN.
__device__ __forceinline__ double dev_exp_tab(double x, const double* __restrict__ tab) {
    double t = fma(x, 92.33248261689366, 6755399441055744.0);   // 64/ln2 ; 1.5 * 2^52 rounds to the nearest integer
    int k = __double2loint(t);
    t -= 6755399441055744.0;
    double r = fma(t, -0.010830424667801708, x);               // ln2/64 high part (24 trailing zero bits: t * hi is exact)
    r = fma(t, -2.8447437476627285e-11, r);                    // ln2/64 low part
    const double T = tab[k & 63];
    double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = q * r;                                                 // exp(r) - 1
    const double m = fma(T, q, T);                             // in [1, 2)
    k = min(max(k, -64512), 64512);                            // saturate at exp(+-698.7) (integer pipe; valid for |x| < 2e7)
    return __hiloint2double(__double2hiint(m) + ((k >> 6) << 20), __double2loint(m));
}

// 1/d for finite normal d >= 1: hardware seed (MUFU.RCP64H) + two Newton steps; error <= 1 ulp, no slow path.
__device__ __forceinline__ double dev_rcp_fast(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    return fma(y, e, y);
}

// Gradient residual r(eta) of mcmlmodel.h:170-175 (FL 1: y - e^eta), :184-193 (FL 3: 1/(1 + e^eta) + y - 1), :233-238 (FL 7:
// y - eta, without the 1/sigma^2) evaluated with dev_exp_tab / dev_rcp_fast, cut into four stages so that a caller can
// interleave the stages of one tile with the tensor instructions of another (the sampler's software pipeline).
template <int FL>
struct ResidStages {
    double y, x, t, r, T, q;
    int k;
    __device__ __forceinline__ void s0(double y_, double eta, const double* __restrict__ tab) {
        y = y_; x = eta;
        if (FL == 7) return;
        t = fma(x, 92.33248261689366, 6755399441055744.0);
        k = __double2loint(t);
        t -= 6755399441055744.0;
        r = fma(t, -0.010830424667801708, x);
        r = fma(t, -2.8447437476627285e-11, r);
        T = tab[k & 63];
    }
    __device__ __forceinline__ void s1() {
        if (FL == 7) return;
        q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
        q = fma(q, r, 1.0 / 6.0);
        q = fma(q, r, 0.5);
        q = fma(q, r, 1.0);
    }
    __device__ __forceinline__ void s2() {
        if (FL == 7) return;
        q = q * r;
        const double m = fma(T, q, T);
        k = min(max(k, -64512), 64512);
        q = __hiloint2double(__double2hiint(m) + ((k >> 6) << 20), __double2loint(m));   // e^eta
        if (FL == 3) {
            t = q + 1.0;                                                                    // d = e^eta + 1
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(t));                           // y0
            q = fma(-t, r, 1.0);
            r = fma(r, q, r);                                                               // y1
        }
    }
    __device__ __forceinline__ double s3() {
        if (FL == 7) return y - x;
        if (FL == 1) return y - q;
        q = fma(-t, r, 1.0);
        r = fma(r, q, r);                                                                   // y2 = 1/(e^eta + 1)
        return r + y - 1.0;
    }
};

// gradient residual in one call (same arithmetic as the four stages)
template <int FL>
__device__ __forceinline__ double dev_family_resid_tab(double y, double eta, const double* __restrict__ tab) {
    ResidStages<FL> st;
    st.s0(y, eta, tab); st.s1(); st.s2();
    return st.s3();
}

// moremaths.h:16-24
__device__ __forceinline__ double dev_log_factorial_approx(double n) {
    if (n == 0) return 0.0;
    return n * log(n) - n + log(n * (1 + 4 * n * (1 + 2 * n))) / 6 + log(GMB_PI_FAMILY) / 2;
}

// family/link codes with device kernels (mcmlmodel.h:74-87): 1 poisson/log, 2 poisson/identity, 3 binomial/logit, 4 binomial/log,
// 5 binomial/identity, 6 binomial/probit, 7 gaussian/identity, 8 gaussian/log.  1, 3 and 7 (the north-star's) run on every kernel family;
// the others on the general kernels (streaming E-step, two-contraction sampler).
__host__ __device__ static inline bool gmb_flink_supported(int fl) { return fl >= 1 && fl <= 8; }
__host__ __device__ static inline bool gmb_flink_core(int fl) { return fl == 1 || fl == 3 || fl == 7; }
__host__ __device__ static inline bool gmb_flink_gaussian(int fl) { return fl == 7 || fl == 8; }

// Per-observation log-density with the reference's algebra (moremaths.h:26-102).
//   FL 1: poisson/log      : y*eta - exp(eta) - rowc          (rowc = log_factorial_approx(y), hoisted per row)
//   FL 2: poisson/identity : y*log(eta) - eta - rowc
//   FL 3: binomial/logit   : y==1: log(1/(1+exp(-eta))) ; y==0: log(1 - 1/(1+exp(-eta)))
//   FL 4: binomial/log     : y==1: eta ; y==0: log(1 - exp(eta))
//   FL 5: binomial/identity: y==1: log(eta) ; y==0: log(1 - eta)
//   FL 6: binomial/probit  : y==1: log Phi(eta) ; y==0: log(1 - Phi(eta))      (R::pnorm, :69-75)
//   FL 7: gaussian/identity: c0 - 0.5*((y-eta)/sigma)^2         (c0 = -log(sigma) - 0.5*log(2*3.141593))
//   FL 8: gaussian/log     : c0 - 0.5*((log(y)-eta)/sigma)^2    (y here is what the model stores: the constructor already replaced it
//                            by log y, mcmlmodel.h:90-92, so the reference takes the logarithm twice, moremaths.h:81 — kept as is)
// Binomial codes leave responses other than 0 / 1 out of the sum (the reference leaves logl unset there).
template <int FL>
__device__ __forceinline__ double dev_family_ll(double y, double eta, double rowc, double c0, double sigma) {
    if (FL == 1) {
        return y * eta - exp(eta) - rowc;
    } else if (FL == 2) {
        return y * log(eta) - eta - rowc;
    } else if (FL == 3) {
        double p = 1.0 / (1.0 + exp(-1.0 * eta));
        double r = 0.0;
        if (y == 1.0) r = log(p);
        else if (y == 0.0) r = log(1.0 - p);
        return r;
    } else if (FL == 4) {
        double r = 0.0;
        if (y == 1.0) r = eta;
        else if (y == 0.0) r = log(1.0 - exp(eta));
        return r;
    } else if (FL == 5) {
        double r = 0.0;
        if (y == 1.0) r = log(eta);
        else if (y == 0.0) r = log(1.0 - eta);
        return r;
    } else if (FL == 6) {
        double r = 0.0;
        if (y == 1.0) r = log(normcdf(eta));
        else if (y == 0.0) r = log(1.0 - normcdf(eta));
        return r;
    } else if (FL == 7) {
        double z = (y - eta) / sigma;
        return c0 - 0.5 * z * z;
    } else {
        double z = (log(y) - eta) / sigma;
        return c0 - 0.5 * z * z;
    }
}

// gradient residual r(eta) of mcmlmodel.h:170-175 (FL 1), :176-183 (2), :184-193 (3), :194-206 (4), :207-219 (5), :220-232 (6), :233-244 (7, 8:
// without the 1/sigma^2).  Codes 4-6 as the reference writes them (e.g. the y == 0 branch of code 4 carries the reference's sign).
template <int FL, bool FAST = false>
__device__ __forceinline__ double dev_family_resid(double y, double eta) {
    if (FAST) {
        if (FL == 1) return y - dev_exp(eta);
        if (FL == 3) return __drcp_rn(dev_exp(eta) + 1.0) + y - 1.0;
    } else {
        if (FL == 1) return y - exp(eta);
        if (FL == 3) return 1.0 / (exp(eta) + 1.0) + y - 1.0;
    }
    if (FL == 2) return y * (1.0 / eta) - 1.0;
    if (FL == 4) { if (y == 1.0) return 1.0; if (y == 0.0) return exp(eta) / (1.0 - exp(eta)); return eta; }
    if (FL == 5) { if (y == 1.0) return 1.0 / eta; if (y == 0.0) return -1.0 / (1.0 - eta); return eta; }
    if (FL == 6) {
        const double pdf = 0.3989422804014327 * exp(-0.5 * eta * eta);      // R::dnorm(eta, 0, 1)
        if (y == 1.0) return pdf / normcdf(eta);
        if (y == 0.0) return -1.0 * pdf / (1.0 - normcdf(eta));
        return eta;
    }
    return y - eta;
}

// Row-aggregated forms (aggregate.cu): a row stands for c observations that share eta; ys is the response term of the residual
// (binomial: sum y - c, others: sum y).  With c = 1 these are the per-observation formulas of mcmlmodel.h:170-175, :184-193, :233-238.
template <int FL>
__device__ __forceinline__ double dev_family_resid_w(double c, double ys, double eta, const double* __restrict__ tab) {
    if (FL == 1) return fma(-c, dev_exp_tab(eta, tab), ys);
    if (FL == 3) return fma(c, dev_rcp_fast(dev_exp_tab(eta, tab) + 1.0), ys);
    return fma(-c, eta, ys);
}
// log-density of a row: lc observations counted, lys = sum of y (poisson, binomial) or their mean (gaussian), lsq = within-row sum of
// squares (gaussian), lrc = sum of lf(y) (poisson).  With lc = 1: dev_family_ll.
template <int FL>
__device__ __forceinline__ double dev_family_ll_w(double lc, double lys, double lsq, double lrc, double eta, double c0, double sigma) {
    if (FL == 1) {
        return lys * eta - lc * exp(eta) - lrc;
    } else if (FL == 3) {
        const double p = 1.0 / (1.0 + exp(-1.0 * eta));
        double r = 0.0;
        if (lys != 0.0) r += lys * log(p);
        if (lc - lys != 0.0) r += (lc - lys) * log(1.0 - p);
        return r;
    } else {
        const double z = (lys - eta) / sigma;
        return lc * c0 - 0.5 * (lsq / (sigma * sigma) + lc * z * z);
    }
}

// Covariance kernel functions (glmmrBase DSubMatrix::get_val; SURVEY.md App. C.2 — reconstructed, the table lives here for the device side and
// in oracle/oracle.cpp cov_fn for the oracle): entry = prod_k f_k(dist_k; theta), eff = the function's effective range (compact support).
//   1 gr  2 fexp0  3 ar1  4 sqexp  7 wend0  8 wend1  9 wend2  13 fexp  14 sqexp0
__host__ __device__ static inline bool gmb_cov_fn_supported(int id) { return id == 1 || id == 2 || id == 3 || id == 4 || id == 7 || id == 8 || id == 9 || id == 13 || id == 14; }
__device__ __forceinline__ double dev_cov_fn(int id, double d, const double* th, double eff) {
    switch (id) {
    case 1:  return d == 0.0 ? th[0] * th[0] : 0.0;                 // gr
    case 2:  return exp(-d / th[0]);                                // fexp0
    case 3:  return pow(th[0], d);                                  // ar1
    case 4:  return th[0] * exp(-d * d / (th[1] * th[1]));          // sqexp
    case 7:  { const double x = d / eff; return x < 1.0 ? th[0] * pow(1.0 - x, th[1]) : 0.0; }                                                   // wend0
    case 8:  { const double x = d / eff; return x < 1.0 ? th[0] * (1.0 + th[1] * x) * pow(1.0 - x, th[1]) : 0.0; }                               // wend1
    case 9:  { const double x = d / eff; return x < 1.0 ? th[0] * (1.0 + th[1] * x + (th[1] * th[1] - 1.0) * (1.0 / 3.0) * x * x) * pow(1.0 - x, th[1]) : 0.0; }   // wend2
    case 13: return th[0] * exp(-d / th[1]);                        // fexp
    case 14: return exp(-d * d / (th[0] * th[0]));                  // sqexp0
    }
    return nan("");
}

// Philox4x32-10; counter = (idx, iteration, chain, stream), key = seed.  Must match oracle/oracle.cpp.
__device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

__device__ __forceinline__ double dev_u01(uint32_t lo, uint32_t hi) {
    unsigned long long x = ((unsigned long long)hi << 32) | lo;
    return ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void dev_rng_uniform2(unsigned long long seed, uint32_t idx, uint32_t iter, uint32_t chain,
                                                 uint32_t stream, double& u1, double& u2) {
    uint32_t c0 = idx, c1 = iter, c2 = chain, c3 = stream;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    u1 = dev_u01(c0, c1); u2 = dev_u01(c2, c3);
}

// Box-Muller pair p (elements 2p, 2p+1 of the normal vector)
__device__ __forceinline__ void dev_rng_normal2(unsigned long long seed, uint32_t p, uint32_t iter, uint32_t chain,
                                                uint32_t stream, double& z0, double& z1) {
    double u1, u2;
    dev_rng_uniform2(seed, p, iter, chain, stream, u1, u2);
    double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    z0 = r * c; z1 = r * s;
}

#endif  // __CUDACC__
