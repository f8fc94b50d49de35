// cov_large.cu — covariance blocks larger than a warp: blocked right-looking Cholesky and blocked forward
// substitution with many right-hand sides, both built on the DMMA GEMM (gemm_f64.cu).
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

// unblocked Cholesky of the kb x kb diagonal block at (k0, k0), one CTA of 64 threads (thread = row), in shared memory
__global__ void __launch_bounds__(NB) potf2_kernel(double* __restrict__ A, int ld, int k0, int kb, int row_offset, int* __restrict__ status) {
    __shared__ double s[NB][NB + 1];
    const int t = threadIdx.x;
    if (t < kb) for (int j = 0; j <= t; j++) s[t][j] = A[(k0 + t) + (size_t)(k0 + j) * ld];
    __syncthreads();
    for (int j = 0; j < kb; j++) {
        double sum = 0.0;
        if (t >= j && t < kb) for (int k = 0; k < j; k++) sum += s[t][k] * s[j][k];
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
}

// panel solve: rows r >= k0+kb of columns [k0, k0+kb) <- A[r, k0:k0+kb] * L_kk^{-T} ; thread = row
__global__ void __launch_bounds__(128) trsm_panel_kernel(double* __restrict__ A, int ld, int n, int k0, int kb) {
    __shared__ double Lkk[NB][NB + 1];
    for (int e = threadIdx.x; e < kb * kb; e += blockDim.x) { int i = e % kb, j = e / kb; Lkk[i][j] = A[(k0 + i) + (size_t)(k0 + j) * ld]; }
    __syncthreads();
    int r = k0 + kb + blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double x[NB];
#pragma unroll 1
    for (int j = 0; j < kb; j++) {
        double v = A[r + (size_t)(k0 + j) * ld];
        for (int k = 0; k < j; k++) v -= x[k] * Lkk[j][k];
        x[j] = v / Lkk[j][j];
        A[r + (size_t)(k0 + j) * ld] = x[j];
    }
}

__global__ void logdet_diag_kernel(const double* __restrict__ A, int ld, int n, double* __restrict__ out) {
    __shared__ double red[32];
    double c = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) c += 2.0 * log(A[i + (size_t)i * ld]);
    c = block_sum(c, red);
    if (threadIdx.x == 0) out[0] = c;
}

// diagonal-block solve for the blocked forward substitution: W[k0:k0+kb, j] <- L_kk^{-1} W[k0:k0+kb, j]; thread = column
__global__ void __launch_bounds__(128) trsv_cols_kernel(const double* __restrict__ A, int ld, int k0, int kb,
                                                        double* __restrict__ W, int ldw, int ncols) {
    extern __shared__ double sm[];
    double* Lkk = sm;                  // kb x kb (col-major), reciprocal diagonal
    double* z = sm + NB * NB;          // [kb][128]
    for (int e = threadIdx.x; e < kb * kb; e += blockDim.x) {
        int i = e % kb, j = e / kb;
        double v = A[(k0 + i) + (size_t)(k0 + j) * ld];
        Lkk[e] = (i == j) ? 1.0 / v : v;
    }
    __syncthreads();
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    double* w = W + (size_t)j * ldw + k0;
    for (int i = 0; i < kb; i++) {
        double lsum = 0.0;
        for (int k = 0; k < i; k++) lsum += Lkk[i + k * kb] * z[k * 128 + threadIdx.x];
        double zi = (w[i] - lsum) * Lkk[i + i * kb];
        z[i * 128 + threadIdx.x] = zi;
        w[i] = zi;
    }
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
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y;
    if (i < ldw && j < ncols) W[(size_t)j * ldw + i] = (i < n) ? U[(size_t)j * ldu + start + i] : 0.0;
}

}  // namespace

int gmb_cov_factor_large(gmb_cov* cv, int bi) {
    gmb_ctx* ctx = cv->ctx;
    const CovBlock& b = cv->blocks[bi];
    const int n = b.n, ld = gmb_cov_ld(n);
    double* A = cv->d_Lblk + b.l0;
    dim3 blk(32, 8), grd((n + 31) / 32, (n + 7) / 8);
    build_block_kernel<<<grd, blk, 0, ctx->stream>>>(b, cv->d_fns, cv->d_data, cv->d_theta, A, ld);
    ctx->launches++;
    for (int k0 = 0; k0 < n; k0 += NB) {
        int kb = n - k0 < NB ? n - k0 : NB;
        potf2_kernel<<<1, NB, 0, ctx->stream>>>(A, ld, k0, kb, b.start, cv->d_status);
        ctx->launches++;
        int rest = n - k0 - kb;
        if (rest > 0) {
            trsm_panel_kernel<<<(rest + 127) / 128, 128, 0, ctx->stream>>>(A, ld, n, k0, kb);
            ctx->launches++;
            // trailing update A22 -= P P^T  (P = A[k0+kb:, k0:k0+kb]); full square, the upper part is never read
            const double* Pm = A + (k0 + kb) + (size_t)k0 * ld;
            double* A22 = A + (k0 + kb) + (size_t)(k0 + kb) * ld;
            GMB_TRY(gmb_dgemm(ctx, 0, 1, rest, rest, kb, -1.0, Pm, ld, Pm, ld, 1.0, A22, ld));
        }
    }
    logdet_diag_kernel<<<1, 256, 0, ctx->stream>>>(A, ld, n, cv->d_logdet + bi);
    ctx->launches++;
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
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
    static bool configured = false;
    size_t smem = (size_t)(NB * NB + NB * 128) * sizeof(double);
    if (!configured) { GMB_CUDA(cudaFuncSetAttribute(trsv_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured = true; }
    int nchunks = (ncols + chunk - 1) / chunk;
    int slots = 64 / nchunks; if (slots < 1) slots = 1;
    for (int c = 0; c < nchunks; c++) {
        int c0 = c * chunk, nc = ncols - c0 < chunk ? ncols - c0 : chunk;
        copy_rows_kernel<<<dim3((ldw + 255) / 256, nc), 256, 0, ctx->stream>>>(dU + (size_t)c0 * ldu, ldu, b.start, n, nc, W, ldw);
        ctx->launches++;
        for (int k0 = 0; k0 < n; k0 += NB) {
            int kb = n - k0 < NB ? n - k0 : NB;
            trsv_cols_kernel<<<(nc + 127) / 128, 128, smem, ctx->stream>>>(A, ld, k0, kb, W, ldw, nc);
            ctx->launches++;
            int rest = n - k0 - kb;
            if (rest > 0)   // W[k0+kb:, :] -= L[k0+kb:, k0:k0+kb] * W[k0:k0+kb, :]
                GMB_TRY(gmb_dgemm(ctx, 0, 0, rest, nc, kb, -1.0, A + (k0 + kb) + (size_t)k0 * ld, ld, W + k0, ldw, 1.0, W + k0 + kb, ldw));
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
