// cov_large.cu — covariance blocks larger than a warp: two-level blocked right-looking Cholesky and two-level blocked
// forward substitution with many right-hand sides, both built on the DMMA GEMM (gemm_f64.cu).
//
// Replaces, for one dense block (e.g. the fexp Gaussian-process block of configs C3/C5), glmmrBase
// gen_block_mat(b, true, false) (unblocked Cholesky–Banachiewicz) and the per-sample
// algo::forward_sub of mcmldmatrix.h:67-75 / moremaths.h:166-179.
#include "common.cuh"

namespace {

constexpr int NB = 64;   // panel width

__device__ __forceinline__ double cov_fn_eval_l(int id, double d, const double* th) {
    switch (id) {
    case 1:  return d == 0.0 ? th[0] * th[0] : 0.0;
    case 2:  return exp(-d / th[0]);
    case 3:  return pow(th[0], d);
    case 4:  return th[0] * exp(-d * d / (th[1] * th[1]));
    case 13: return th[0] * exp(-d / th[1]);
    case 14: return exp(-d * d / (th[0] * th[0]));
    }
    return nan("");
}

// fill the lower triangle of the block with D_b(i,j) (upper triangle zero)
__global__ void build_block_kernel(CovBlock b, const CovFn* __restrict__ fns, const double* __restrict__ data,
                                   const double* __restrict__ theta, double* __restrict__ A, int ld) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= b.n || j >= b.n) return;
    double v = 0.0;
    if (j <= i) {
        v = 1.0;
        const double* dat = data + b.data0;
        for (int f = 0; f < b.nfn; f++) {
            const CovFn fn = fns[b.fn0 + f];
            double d2 = 0.0;
            for (int k = 0; k < fn.nvar; k++) {
                double di = dat[i + (size_t)(fn.col0 + k) * b.n] - dat[j + (size_t)(fn.col0 + k) * b.n];
                d2 += di * di;
            }
            v *= cov_fn_eval_l(fn.id, sqrt(d2), theta + fn.par0);
        }
    }
    A[i + (size_t)j * ld] = v;
}

// unblocked Cholesky of the kb x kb diagonal block at (k0, k0), one CTA of 64 threads (thread = row), in shared memory;
// also writes the inverse of the factor (64 x 64 col-major, zero padded) to Linv: the panel solve of the factorisation and
// the diagonal solves of the forward substitution then run as DMMA GEMMs instead of one serial recurrence per thread.
__global__ void __launch_bounds__(NB) potf2_kernel(double* __restrict__ A, int ld, int k0, int kb, int row_offset, int* __restrict__ status,
                                                   double* __restrict__ Linv) {
    __shared__ double s[NB][NB + 1];
    const int t = threadIdx.x;
    for (int j = 0; j < NB; j++) s[t][j] = (t < kb && j <= t && j < kb) ? A[(k0 + t) + (size_t)(k0 + j) * ld] : (t == j ? 1.0 : 0.0);
    __syncthreads();
    for (int j = 0; j < kb; j++) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        if (t >= j && t < kb) {
            int k = 0;
            for (; k + 3 < j; k += 4) { s0 += s[t][k] * s[j][k]; s1 += s[t][k + 1] * s[j][k + 1]; s2 += s[t][k + 2] * s[j][k + 2]; s3 += s[t][k + 3] * s[j][k + 3]; }
            for (; k < j; k++) s0 += s[t][k] * s[j][k];
        }
        const double sum = (s0 + s1) + (s2 + s3);
        __shared__ double djj;
        if (t == j) djj = s[j][j] - sum;
        __syncthreads();
        if (!(djj > 0.0)) { if (t == 0) atomicCAS(status, 0, row_offset + k0 + j + 1); return; }
        double d = sqrt(djj);
        if (t == j) s[j][j] = d;
        else if (t > j && t < kb) s[t][j] = (s[t][j] - sum) / d;
        __syncthreads();
    }
    if (t < kb) for (int j = 0; j < kb; j++) A[(k0 + t) + (size_t)(k0 + j) * ld] = (j <= t) ? s[t][j] : 0.0;
    // inverse: thread t solves L x = e_t (rows/cols >= kb form an identity block); every thread runs the same recurrence on
    // broadcast reads of L, entries above the diagonal come out as exact zeros
    double x[NB];
#pragma unroll
    for (int i = 0; i < NB; i++) {
        double a0 = (i == t) ? 1.0 : 0.0, a1 = 0.0;
#pragma unroll
        for (int k = 0; k < i; k++) { if (k & 1) a1 -= s[i][k] * x[k]; else a0 -= s[i][k] * x[k]; }
        x[i] = (a0 + a1) / s[i][i];
    }
#pragma unroll
    for (int i = 0; i < NB; i++) Linv[i + (size_t)t * NB] = (i >= t && i < kb && t < kb) ? x[i] : 0.0;
}

__global__ void logdet_diag_kernel(const double* __restrict__ A, int ld, int n, double* __restrict__ out) {
    __shared__ double red[32];
    double c = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) c += 2.0 * log(A[i + (size_t)i * ld]);
    c = block_sum(c, red);
    if (threadIdx.x == 0) out[0] = c;
}

__global__ void __launch_bounds__(256) sumsq_kernel(const double* __restrict__ W, int ldw, int n, int ncols, double* __restrict__ partials) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int j = blockIdx.x; j < ncols; j += gridDim.x) {
        const double* w = W + (size_t)j * ldw;
        for (int i = threadIdx.x; i < n; i += blockDim.x) acc += w[i] * w[i];
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

__global__ void copy_rows_kernel(const double* __restrict__ U, int ldu, int start, int n, int ncols, double* __restrict__ W, int ldw) {
    int i = blockIdx.y * blockDim.x + threadIdx.x;      // grid.x runs over the columns (no 65535 limit)
    int j = blockIdx.x;
    if (i < ldw && j < ncols) W[(size_t)j * ldw + i] = (i < n) ? U[(size_t)j * ldu + start + i] : 0.0;
}

}  // namespace

// Two-level blocking.  Panels of NB = 64 columns are factorised in shared memory (potf2 + panel solve); their updates are
// applied right away only INSIDE the current outer block of NBO = 256 columns, and the trailing matrix receives one
// rank-256 update per outer block, split into block columns so that only the lower trapezoid is computed.  The rank-64
// updates of the plain right-looking form ran the DMMA GEMM with a 4-iteration k loop (pipeline fill/drain dominated) and
// computed the full square.
constexpr int NBO = 256;   // outer block
constexpr int NBC = 512;   // block-column width of the trailing update

// storage for the inverted diagonal blocks of large block `bi` (allocated for all large blocks at first use)
static int linv_buffer(gmb_cov* cv, int bi, double** out) {
    gmb_ctx* ctx = cv->ctx;
    if (cv->linv_off.empty()) {
        long long off = 0;
        cv->linv_off.assign(cv->blocks.size(), -1);
        for (size_t k = 0; k < cv->blocks.size(); k++)
            if (cv->blocks[k].n > GMB_COV_SMALL_MAX) { cv->linv_off[k] = off; off += (long long)((cv->blocks[k].n + NB - 1) / NB) * NB * NB; }
        cv->linv_doubles = (size_t)off;
        GMB_CUDA(gmb_dmalloc(ctx, &cv->d_linv, sizeof(double) * (off > 0 ? off : 1)));
    }
    if (cv->linv_off[bi] < 0) return gmb_set_error(GMB_ESTATE, "block %d has no inverse-diagonal storage", bi);
    *out = cv->d_linv + cv->linv_off[bi];
    return GMB_OK;
}

// in-place lower Cholesky of the n x n device matrix A (lower triangle read; the strict upper triangle of the 64 x 64 diagonal
// blocks is zeroed, the rest of the upper triangle is left as scratch).  linv: ceil(n/64) * 64 * 64 doubles (inverted diagonal
// blocks), status: first non-PD pivot + 1 + row_offset (0 if fine), d_logdet: sum of 2 log L_ii.
int gmb_chol_blocked(gmb_ctx* ctx, double* A, int ld, int n, int row_offset, int* d_status, double* linv, double* d_logdet) {
    for (int K0 = 0; K0 < n; K0 += NBO) {
        const int KB = n - K0 < NBO ? n - K0 : NBO;
        const int Kend = K0 + KB;
        for (int k0 = K0; k0 < Kend; k0 += NB) {
            const int kb = Kend - k0 < NB ? Kend - k0 : NB;
            double* Li = linv + (size_t)(k0 / NB) * NB * NB;
            potf2_kernel<<<1, NB, 0, ctx->stream>>>(A, ld, k0, kb, row_offset, d_status, Li);
            ctx->launches++;
            const int rest = n - k0 - kb;
            if (rest > 0) {
                // panel solve P <- P L_kk^{-T}, in place: a CTA tile spans all kb <= 64 columns, so it only reads its own rows
                double* Pp = A + (k0 + kb) + (size_t)k0 * ld;
                GMB_TRY(gmb_dgemm(ctx, 0, 1, rest, kb, kb, 1.0, Pp, ld, Li, NB, 0.0, Pp, ld));
                // inside the outer block: A[k0+kb:n, k0+kb:Kend] -= P Ptop^T, P = A[k0+kb:n, k0:k0+kb]
                const int ncols_in = Kend - (k0 + kb);
                if (ncols_in > 0) {
                    const double* Pm = A + (k0 + kb) + (size_t)k0 * ld;
                    GMB_TRY(gmb_dgemm(ctx, 0, 1, rest, ncols_in, kb, -1.0, Pm, ld, Pm, ld, 1.0, A + (k0 + kb) + (size_t)(k0 + kb) * ld, ld));
                }
            }
        }
        // trailing matrix: A[J:n, J:J+w] -= Pn[J:n, :] Pn[J:J+w, :]^T for block columns J, Pn = A[Kend:n, K0:Kend]
        for (int J = Kend; J < n; J += NBC) {
            const int w = n - J < NBC ? n - J : NBC;
            const double* PJ = A + J + (size_t)K0 * ld;
            GMB_TRY(gmb_dgemm(ctx, 0, 1, n - J, w, KB, -1.0, PJ, ld, PJ, ld, 1.0, A + J + (size_t)J * ld, ld));
        }
    }
    if (d_logdet) {
        logdet_diag_kernel<<<1, 256, 0, ctx->stream>>>(A, ld, n, d_logdet);
        ctx->launches++;
    }
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}

int gmb_cov_factor_large(gmb_cov* cv, int bi) {
    gmb_ctx* ctx = cv->ctx;
    const CovBlock& b = cv->blocks[bi];
    const int n = b.n, ld = gmb_cov_ld(n);
    double* A = cv->d_Lblk + b.l0;
    double* linv = nullptr;
    GMB_TRY(linv_buffer(cv, bi, &linv));
    dim3 blk(32, 8), grd((n + 31) / 32, (n + 7) / 8);
    build_block_kernel<<<grd, blk, 0, ctx->stream>>>(b, cv->d_fns, cv->d_data, cv->d_theta, A, ld);
    ctx->launches++;
    return gmb_chol_blocked(ctx, A, ld, n, b.start, cv->d_status, linv, cv->d_logdet + bi);
}

// d_partial: 64 doubles (zeroed by the caller); receives partial sums of ||L^{-1} u_j||^2 over the columns
int gmb_cov_quad_large(gmb_cov* cv, int bi, const double* dU, int ldu, int ncols, double* d_partial) {
    gmb_ctx* ctx = cv->ctx;
    const CovBlock& b = cv->blocks[bi];
    const int n = b.n, ld = gmb_cov_ld(n), ldw = round_up(n, 4);
    const double* A = cv->d_Lblk + b.l0;
    // workspace: a copy of the block's rows of U for a chunk of columns (bounded to ~2 GiB)
    size_t max_cols = ((size_t)1 << 28) / (size_t)ldw;
    if (max_cols < 128) max_cols = 128;
    int chunk = ncols < (int)max_cols ? ncols : (int)max_cols;
    size_t need = (size_t)ldw * chunk;
    if (need > cv->work_doubles) {
        if (cv->d_work) { GMB_CUDA(cudaStreamSynchronize(ctx->stream)); gmb_dfree(ctx, cv->d_work); cv->d_work = nullptr; }
        GMB_CUDA(gmb_dmalloc(ctx, &cv->d_work, need * sizeof(double)));
        cv->work_doubles = need;
    }
    double* W = cv->d_work;
    double* linv = nullptr;
    GMB_TRY(linv_buffer(cv, bi, &linv));
    int nchunks = (ncols + chunk - 1) / chunk;
    int slots = 64 / nchunks; if (slots < 1) slots = 1;
    for (int c = 0; c < nchunks; c++) {
        int c0 = c * chunk, nc = ncols - c0 < chunk ? ncols - c0 : chunk;
        copy_rows_kernel<<<dim3(nc, (ldw + 255) / 256), 256, 0, ctx->stream>>>(dU + (size_t)c0 * ldu, ldu, b.start, n, nc, W, ldw);
        ctx->launches++;
        // blocked forward substitution, two levels: 64-row diagonal solves and rank-64 updates inside an outer block of 256
        // rows, one rank-256 update of all rows below per outer block
        for (int K0 = 0; K0 < n; K0 += NBO) {
            const int KB = n - K0 < NBO ? n - K0 : NBO;
            const int Kend = K0 + KB;
            for (int k0 = K0; k0 < Kend; k0 += NB) {
                const int kb = Kend - k0 < NB ? Kend - k0 : NB;
                // diagonal solve W[k0:k0+kb, :] <- L_kk^{-1} W[k0:k0+kb, :], in place (a CTA tile spans all kb <= 64 rows)
                GMB_TRY(gmb_dgemm(ctx, 0, 0, kb, nc, kb, 1.0, linv + (size_t)(k0 / NB) * NB * NB, NB, W + k0, ldw, 0.0, W + k0, ldw));
                const int rows_in = Kend - (k0 + kb);
                if (rows_in > 0)   // W[k0+kb:Kend, :] -= L[k0+kb:Kend, k0:k0+kb] W[k0:k0+kb, :]
                    GMB_TRY(gmb_dgemm(ctx, 0, 0, rows_in, nc, kb, -1.0, A + (k0 + kb) + (size_t)k0 * ld, ld, W + k0, ldw, 1.0, W + k0 + kb, ldw));
            }
            const int rest = n - Kend;
            if (rest > 0)          // W[Kend:, :] -= L[Kend:, K0:Kend] W[K0:Kend, :]
                GMB_TRY(gmb_dgemm(ctx, 0, 0, rest, nc, KB, -1.0, A + Kend + (size_t)K0 * ld, ld, W + K0, ldw, 1.0, W + Kend, ldw));
        }
        if (c < 64) {
            int s0 = (c * slots) % 64;
            sumsq_kernel<<<slots, 256, 0, ctx->stream>>>(W, ldw, n, nc, d_partial + s0);
            ctx->launches++;
        } else {
            return gmb_set_error(GMB_EINVAL, "too many column chunks for a large covariance block");
        }
    }
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}
