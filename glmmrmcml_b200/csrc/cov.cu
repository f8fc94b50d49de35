// cov.cu — covariance update: D(theta) build, block Cholesky and the multivariate-normal objective.
//
// K4: glmmrBase DMatrix::genD / gen_block_mat / DSubMatrix::get_val as reconstructed in SURVEY.md App. C
//     (call sites mcmldmatrix.h:46,59 ; src/mcml_full.cpp:68,121) — entries D_b(i,j) = prod_k f_k(dist_k(i,j); theta),
//     lower Cholesky by Cholesky–Banachiewicz.
// K5: MCMLDmatrix::loglik / loglik_block (mcmldmatrix.h:23-41, 57-78) + algo::forward_sub (moremaths.h:166-179).
//
// The reference re-factorises every block once per SAMPLE (mcmldmatrix.h:33-36 -> :59).  Here every block is
// factorised once per theta; the per-sample work is only the forward substitution, which for small blocks is an
// HBM stream of u (8 Q m bytes) and for large blocks a blocked TRSM on the DMMA GEMM.
#include "common.cuh"
#include <map>
#include <string>
#include <algorithm>

namespace {

// SURVEY.md App. C.2 — the id -> function table lives here (and in oracle/oracle.cpp cov_fn) so that it can be
// corrected in one place should the glmmrBase sources become available.
__host__ __device__ inline int cov_fn_npar(int id) {
    switch (id) { case 1: case 2: case 3: case 6: case 14: return 1; case 4: case 5: case 7: case 8: case 9: case 10: case 11: case 12: case 13: return 2; }
    return 0;
}
__host__ __device__ inline bool cov_fn_supported(int id) { return gmb_cov_fn_supported(id); }

// DSubMatrix::get_val(i, j)
__device__ __forceinline__ double block_val(const CovBlock& b, const CovFn* __restrict__ fns, const double* __restrict__ data,
                                            const double* __restrict__ theta, int i, int j) {
    double v = 1.0;
    const double* dat = data + b.data0;
    for (int f = 0; f < b.nfn; f++) {
        const CovFn fn = fns[b.fn0 + f];
        double d2 = 0.0;
        for (int k = 0; k < fn.nvar; k++) {
            double di = dat[i + (size_t)(fn.col0 + k) * b.n] - dat[j + (size_t)(fn.col0 + k) * b.n];
            d2 += di * di;
        }
        v *= dev_cov_fn(fn.id, sqrt(d2), theta + fn.par0, fn.eff);
    }
    return v;
}

// ---------------------------------------------------------------------------------------------------
// K4 small: one warp per block with n_b <= 32.  Lane i owns row i.  Writes the factor (col-major n_b x n_b),
// the block's log-determinant sum_i 2 log L_ii and a non-PD status.
// ---------------------------------------------------------------------------------------------------
constexpr int SMALL_MAX = GMB_COV_SMALL_MAX;
constexpr int FACT_WARPS = 4;

__global__ void __launch_bounds__(FACT_WARPS * 32) factor_small_kernel(int B, const CovBlock* __restrict__ blocks,
                                                                        const CovFn* __restrict__ fns, const double* __restrict__ data,
                                                                        const double* __restrict__ theta, double* __restrict__ Lblk,
                                                                        double* __restrict__ logdet, int* __restrict__ status) {
    __shared__ double sL[FACT_WARPS][SMALL_MAX][SMALL_MAX + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bi = blockIdx.x * FACT_WARPS + warp;
    if (bi >= B) return;
    const CovBlock b = blocks[bi];
    if (b.n > SMALL_MAX) return;
    const int n = b.n;
    double (*L)[SMALL_MAX + 1] = sL[warp];
    // entries of the lower triangle, row = lane
    if (lane < n) for (int j = 0; j <= lane; j++) L[lane][j] = block_val(b, fns, data, theta, lane, j);
    __syncwarp();
    bool bad = false;
    for (int j = 0; j < n; j++) {
        // Cholesky–Banachiewicz column j: s accumulated in ascending k exactly like the oracle
        double s = 0.0;
        if (lane >= j && lane < n) for (int k = 0; k < j; k++) s += L[lane][k] * L[j][k];
        double djj = __shfl_sync(0xffffffffu, L[j][j] - s, j);   // lane j holds a_jj - sum
        if (!(djj > 0.0)) { bad = true; if (lane == 0) atomicCAS(status, 0, b.start + j + 1); break; }
        double d = sqrt(djj);
        __syncwarp();
        if (lane == j) L[j][j] = d;
        else if (lane > j && lane < n) L[lane][j] = (L[lane][j] - s) / d;
        __syncwarp();
    }
    if (bad) { if (lane == 0) logdet[bi] = nan(""); return; }
    double ld = (lane < n) ? 2.0 * log(L[lane][lane]) : 0.0;
    ld = warp_sum(ld);
    if (lane == 0) logdet[bi] = ld;
    double* out = Lblk + b.l0;
    for (int e = lane; e < n * n; e += 32) {
        int i = e % n, j = e / n;
        out[e] = (j <= i) ? L[i][j] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------------
// K5 small: forward substitution for blocks with n_b <= 16.  CTA = 8 warps handles a group of 32 consecutive blocks
// (lane = block) and a chunk of sample columns (warp w takes columns j0 + w, j0 + w + 8, ...).  The group's factors
// are staged in shared memory once and reused for every column, so the only HBM traffic is the stream of u.
// Emits sum over (blocks in group, columns in chunk) of ||L_b^{-1} u_bj||^2.
// ---------------------------------------------------------------------------------------------------
constexpr int QUAD_SMALL_MAX = 16;

__global__ void __launch_bounds__(256) quad_small_kernel(int B, const CovBlock* __restrict__ blocks, const double* __restrict__ Lblk,
                                                         const double* __restrict__ U, int ldu, int ncols, int cols_per_cta,
                                                         double* __restrict__ partials) {
    extern __shared__ double sL[];                  // factors of the group, block stride padded to an odd count
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b0 = blockIdx.x * 32;
    const int nb = min(32, B - b0);
    const int j0 = blockIdx.y * cols_per_cta, j1 = min(j0 + cols_per_cta, ncols);
    const int stride = QUAD_SMALL_MAX * QUAD_SMALL_MAX + 1;
    // stage factors; store 1/L_ii on the diagonal so the inner loop multiplies
    for (int bb = 0; bb < nb; bb++) {
        const CovBlock b = blocks[b0 + bb];
        if (b.n > QUAD_SMALL_MAX) continue;
        for (int e = threadIdx.x; e < b.n * b.n; e += 256) {
            int i = e % b.n, j = e / b.n;
            double v = Lblk[b.l0 + e];
            sL[bb * stride + i * QUAD_SMALL_MAX + j] = (i == j) ? 1.0 / v : v;
        }
    }
    __syncthreads();
    double acc = 0.0;
    if (lane < nb) {
        const CovBlock b = blocks[b0 + lane];
        if (b.n <= QUAD_SMALL_MAX) {
            const double* L = sL + lane * stride;
            const int n = b.n;
            for (int j = j0 + warp; j < j1; j += 8) {
                const double* u = U + (size_t)j * ldu + b.start;
                double z[QUAD_SMALL_MAX];
                double q = 0.0;
#pragma unroll
                for (int i = 0; i < QUAD_SMALL_MAX; i++) {
                    if (i < n) {
                        double lsum = 0.0;
#pragma unroll
                        for (int k = 0; k < i; k++) lsum += L[i * QUAD_SMALL_MAX + k] * z[k];     // moremaths.h:172-175
                        z[i] = (u[i] - lsum) * L[i * QUAD_SMALL_MAX + i];                         // :176 (reciprocal staged)
                        q += z[i] * z[i];
                    }
                }
                acc += q;
            }
        }
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.y * gridDim.x + blockIdx.x] = acc;
}

// K5 medium: one block with 16 < n_b <= 64, factor in shared memory, one thread per sample column.
constexpr int QUAD_MED_MAX = 64;

__global__ void __launch_bounds__(128) quad_medium_kernel(CovBlock b, const double* __restrict__ Lblk, const double* __restrict__ U,
                                                          int ldu, int ncols, double* __restrict__ partials) {
    extern __shared__ double smm[];
    __shared__ double red[32];
    const int n = b.n;
    double* L = smm;                         // n x n col-major, diagonal holds reciprocals
    double* z = smm + n * n;                 // [n][128]
    const int ldb = gmb_cov_ld(n);
    for (int e = threadIdx.x; e < n * n; e += 128) {
        int i = e % n, j = e / n;
        double v = Lblk[b.l0 + i + (size_t)j * ldb];
        L[e] = (i == j) ? 1.0 / v : v;
    }
    __syncthreads();
    double acc = 0.0;
    for (int j = blockIdx.x * 128 + threadIdx.x; j < ncols; j += gridDim.x * 128) {
        const double* u = U + (size_t)j * ldu + b.start;
        double q = 0.0;
        for (int i = 0; i < n; i++) {
            double lsum = 0.0;
            for (int k = 0; k < i; k++) lsum += L[i + k * n] * z[k * 128 + threadIdx.x];
            double zi = (u[i] - lsum) * L[i + i * n];
            z[i * 128 + threadIdx.x] = zi;
            q += zi * zi;
        }
        acc += q;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// ---------------------------------------------------------------------------------------------------
// K5 through sufficient statistics (all blocks <= 16).  The samples are fixed while d_optim / f_hess evaluate mvn_ll at many theta
// (likelihood.h:40-45), and sum_j ||L_b^-1 u_bj||^2 = tr(D_b^-1 S_b) with the Gram matrix S_b = sum_j u_bj u_bj'.  S_b is built by ONE
// stream over U per sample matrix (gram_small_kernel + gram_reduce_kernel, cached per (model, sample version)); an evaluation is then a
// single launch that builds D_b(theta), factorises it, forms row k of L_b^-1 by back substitution and accumulates x' S_b x — independent
// of the number of samples.  Same sums as mcmldmatrix.h:23-78 in a different order (relative error ~ kappa(D_b) eps).
// ---------------------------------------------------------------------------------------------------
constexpr int GRAM_WARPS = 8;

__global__ void __launch_bounds__(GRAM_WARPS * 32) gram_small_kernel(int B, const CovBlock* __restrict__ blocks, const double* __restrict__ U,
                                                                      int ldu, int ncols, int cols_per_cta, long long gram_doubles,
                                                                      double* __restrict__ part /* [gridDim.y][gram_doubles] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bi = blockIdx.x * GRAM_WARPS + warp;
    if (bi >= B) return;
    const CovBlock b = blocks[bi];
    const int n = b.n, nn = n * n;
    const int j0 = blockIdx.y * cols_per_cta, j1 = min(j0 + cols_per_cta, ncols);
    double acc[8];
    int ei[8], ej[8];
#pragma unroll
    for (int t = 0; t < 8; t++) { const int e = lane + 32 * t; acc[t] = 0.0; ei[t] = e < nn ? e % n : 0; ej[t] = e < nn ? e / n : 0; }
    for (int j = j0; j < j1; j++) {
        const double* u = U + (size_t)j * ldu + b.start;
#pragma unroll
        for (int t = 0; t < 8; t++) if (lane + 32 * t < nn) acc[t] = fma(u[ei[t]], u[ej[t]], acc[t]);
    }
    double* out = part + (size_t)blockIdx.y * gram_doubles + b.l0;
#pragma unroll
    for (int t = 0; t < 8; t++) if (lane + 32 * t < nn) out[lane + 32 * t] = acc[t];
}

__global__ void gram_reduce_kernel(long long gram_doubles, int nchunks, const double* __restrict__ part, double* __restrict__ gram) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= gram_doubles) return;
    double s = 0.0;
    for (int c = 0; c < nchunks; c++) s += part[(size_t)c * gram_doubles + e];
    gram[e] = s;
}

// class sums of the Gram matrices: element e of class c = sum over the member blocks in list order (deterministic)
__global__ void gram_class_kernel(int ncls, const CovBlock* __restrict__ blocks, const int* __restrict__ rep, const int* __restrict__ ptr,
                                  const int* __restrict__ mem, const double* __restrict__ gram, double* __restrict__ gram_cls) {
    const int c = blockIdx.x;
    if (c >= ncls) return;
    const CovBlock b = blocks[rep[c]];
    const int nn = b.n * b.n;
    for (int e = threadIdx.x; e < nn; e += blockDim.x) {
        double s = 0.0;
        for (int k = ptr[c]; k < ptr[c + 1]; k++) s += gram[blocks[mem[k]].l0 + e];
        gram_cls[b.l0 + e] = s;
    }
}

// blockIdx.y = evaluation e of a batch: theta, status, partials, counter and out are indexed by e (R parameters per evaluation); Lblk /
// logdet (the factor cache of a single evaluation) may be NULL.
__global__ void __launch_bounds__(FACT_WARPS * 32) mvn_gram_kernel(int B, const CovBlock* __restrict__ blocks, const CovFn* __restrict__ fns,
                                                                    const double* __restrict__ data, const double* __restrict__ theta, int R,
                                                                    const double* __restrict__ gram, double ncols, double* __restrict__ Lblk,
                                                                    double* __restrict__ logdet, int* __restrict__ status,
                                                                    double* __restrict__ partials, unsigned int* __restrict__ counter,
                                                                    double* __restrict__ out, const int* __restrict__ cls_rep,
                                                                    const int* __restrict__ cls_ptr) {
    // cls_rep != NULL: B counts CLASSES of identical blocks; `gram` holds the class sums and the log-determinant term counts once per member
    theta += (size_t)blockIdx.y * R; status += blockIdx.y; partials += (size_t)blockIdx.y * gridDim.x; counter += blockIdx.y; out += blockIdx.y;
    __shared__ double sL[FACT_WARPS][QUAD_SMALL_MAX][QUAD_SMALL_MAX + 1];
    __shared__ double sS[FACT_WARPS][QUAD_SMALL_MAX][QUAD_SMALL_MAX + 1];
    __shared__ double red[32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bi = blockIdx.x * FACT_WARPS + warp;
    double contrib = 0.0;
    if (bi < B) {
        const CovBlock b = blocks[cls_rep ? cls_rep[bi] : bi];
        const double mult = cls_rep ? (double)(cls_ptr[bi + 1] - cls_ptr[bi]) : 1.0;
        const int n = b.n;
        double (*L)[QUAD_SMALL_MAX + 1] = sL[warp];
        double (*S)[QUAD_SMALL_MAX + 1] = sS[warp];
        if (lane < n) for (int j = 0; j <= lane; j++) L[lane][j] = block_val(b, fns, data, theta, lane, j);
        for (int e = lane; e < n * n; e += 32) S[e % n][e / n] = gram[b.l0 + e];
        __syncwarp();
        bool bad = false;
        for (int j = 0; j < n; j++) {                   // Cholesky–Banachiewicz, as factor_small_kernel
            double s = 0.0;
            if (lane >= j && lane < n) for (int k = 0; k < j; k++) s += L[lane][k] * L[j][k];
            double djj = __shfl_sync(0xffffffffu, L[j][j] - s, j);
            if (!(djj > 0.0)) { bad = true; if (lane == 0) atomicCAS(status, 0, b.start + j + 1); break; }
            double d = sqrt(djj);
            __syncwarp();
            if (lane == j) L[j][j] = d;
            else if (lane > j && lane < n) L[lane][j] = (L[lane][j] - s) / d;
            __syncwarp();
        }
        if (bad) {
            contrib = nan("");
            if (lane == 0 && logdet) logdet[bi] = nan("");
        } else {
            double ld = (lane < n) ? 2.0 * log(L[lane][lane]) : 0.0;
            ld = warp_sum(ld);
            if (lane == 0 && logdet) logdet[bi] = ld;
            if (Lblk) for (int e = lane; e < n * n; e += 32) { const int i = e % n, j = e / n; Lblk[b.l0 + e] = (j <= i) ? L[i][j] : 0.0; }
            // lane k: x = row k of L^-1 (solve L' x = e_k by back substitution), q_k = x' S x
            double q = 0.0;
            if (lane < n) {
                double x[QUAD_SMALL_MAX];
                const int k = lane;
#pragma unroll
                for (int i = QUAD_SMALL_MAX - 1; i >= 0; i--) {
                    double v = 0.0;
                    if (i < n && i <= k) {
                        double s = (i == k) ? 1.0 : 0.0;
#pragma unroll
                        for (int j = i + 1; j < QUAD_SMALL_MAX; j++) if (j <= k && j < n) s -= L[j][i] * x[j];
                        v = s / L[i][i];
                    }
                    x[i] = v;
                }
#pragma unroll
                for (int i = 0; i < QUAD_SMALL_MAX; i++) {
                    if (i < n) {
                        double t = 0.0;
#pragma unroll
                        for (int j = 0; j < QUAD_SMALL_MAX; j++) if (j < n) t = fma(S[i][j], x[j], t);
                        q = fma(x[i], t, q);
                    }
                }
            }
            q = warp_sum(q);
            contrib = mult * ncols * (-0.5 * n * log(2 * 3.14159265358979323846) - 0.5 * ld) - 0.5 * q;   // mcmldmatrix.h:63,75 (M_PI)
        }
    }
    // deterministic grid sum: lane 0 of each warp holds its block's contribution
    if (lane == 0) red[warp] = (bi < B) ? contrib : 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int w = 0; w < FACT_WARPS; w++) v += red[w];
        partials[blockIdx.x] = v;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double s = 0.0;
        for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) s += partials[k];
        s = block_sum(s, red);
        if (threadIdx.x == 0) { out[0] = s; *counter = 0u; }
    }
}

// final: out[0] = ncols * sum_b (-0.5 n_b log(2 pi) - 0.5 logdet_b) - 0.5 * sum(partials)   (mcmldmatrix.h:63,75: M_PI)
__global__ void __launch_bounds__(256) mvn_finish_kernel(int B, const CovBlock* __restrict__ blocks, const double* __restrict__ logdet,
                                                         const double* __restrict__ partials, int npart, int ncols,
                                                         double* __restrict__ out) {
    __shared__ double red[32];
    double c = 0.0, q = 0.0;
    for (int b = threadIdx.x; b < B; b += 256) c += -0.5 * blocks[b].n * log(2 * 3.14159265358979323846) - 0.5 * logdet[b];
    for (int k = threadIdx.x; k < npart; k += 256) q += partials[k];
    c = block_sum(c, red);
    q = block_sum(q, red);
    if (threadIdx.x == 0) out[0] = (double)ncols * c - 0.5 * q;
}

__global__ void logdet_sum_kernel(int B, const double* __restrict__ logdet, double* __restrict__ out) {
    __shared__ double red[32];
    double c = 0.0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) c += logdet[b];
    c = block_sum(c, red);
    if (threadIdx.x == 0) out[0] = c;
}

// expand the per-block factors (or, with chol == 0, the blocks of D) into a dense ldq x Q matrix
__global__ void expand_blocks_kernel(int B, const CovBlock* __restrict__ blocks, const CovFn* __restrict__ fns,
                                     const double* __restrict__ data, const double* __restrict__ theta,
                                     const double* __restrict__ Lblk, int chol, double* __restrict__ out, int ld) {
    const CovBlock b = blocks[blockIdx.x];
    const long long nn = (long long)b.n * b.n;
    for (long long e = (long long)blockIdx.y * blockDim.x + threadIdx.x; e < nn; e += (long long)gridDim.y * blockDim.x) {
        const int i = (int)(e % b.n), j = (int)(e / b.n);
        double v = chol ? ((j <= i) ? Lblk[b.l0 + i + (size_t)j * gmb_cov_ld(b.n)] : 0.0) : block_val(b, fns, data, theta, i, j);
        out[(size_t)(b.start + j) * ld + b.start + i] = v;
    }
}

}  // namespace

int gmb_cov_factor_large(gmb_cov* cv, int bi);                                        // cov_large.cu
int gmb_cov_quad_large(gmb_cov* cv, int bi, const double* dU, int ldu, int ncols, double* d_partial, int lower_rhs);
int gmb_cov_gram_large(gmb_cov* cv, int bi, gmb_model* mdl, const double** C_out, int* ldc_out);
void gmb_cov_gram_large_free(gmb_cov* cv);

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
extern "C" int gmb_cov_create(gmb_ctx* ctx, const int32_t* cov, int rows, const double* data, int n_data,
                              const double* eff_range, int n_eff, gmb_cov** out) {
    if (!ctx || !cov || !data || !out || rows <= 0) return gmb_set_error(GMB_EINVAL, "gmb_cov_create: bad arguments");
    gmb_cov* cv = new gmb_cov();
    cv->ctx = ctx;
    int maxb = -1;
    for (int r = 0; r < rows; r++) { if (cov[r] < 0) { delete cv; return gmb_set_error(GMB_EINVAL, "negative block id"); } if (cov[r] > maxb) maxb = cov[r]; }
    cv->B = maxb + 1;
    cv->blocks.assign(cv->B, CovBlock{0, 0, -1, 0, 0, 1, 0, 0});
    // rows of one block are expected to be contiguous (that is how get_D_data() emits them)
    for (int r = 0; r < rows; r++) {
        int b = cov[r], nb = cov[r + rows], id = cov[r + 2 * rows], nv = cov[r + 3 * rows], p0 = cov[r + 4 * rows];
        if (!cov_fn_supported(id)) { delete cv; return gmb_set_error(GMB_ECOV, "covariance function id %d is not supported (supported: 1 gr, 2 fexp0, 3 ar1, 4 sqexp, 7-9 wend0/1/2, 13 fexp, 14 sqexp0)", id); }
        if (nb <= 0 || nv < 0 || p0 < 0) { delete cv; return gmb_set_error(GMB_EINVAL, "bad covariance row %d", r); }
        CovBlock& blk = cv->blocks[b];
        if (blk.fn0 < 0) blk.fn0 = (int)cv->fns.size();
        else if (blk.fn0 + blk.nfn != (int)cv->fns.size()) { delete cv; return gmb_set_error(GMB_EINVAL, "rows of block %d are not contiguous", b); }
        blk.n = nb;
        const double eff = (eff_range && r < n_eff) ? eff_range[r] : 0.0;
        if (id >= 7 && id <= 9 && !(eff > 0.0)) { delete cv; return gmb_set_error(GMB_EINVAL, "covariance row %d: the compact-support function %d needs eff_range > 0", r, id); }
        cv->fns.push_back(CovFn{id, nv, p0, blk.ncol, eff});
        blk.nfn++; blk.ncol += nv;
        if (id != 1) blk.all_gr = 0;
        if (p0 + cov_fn_npar(id) > cv->R) cv->R = p0 + cov_fn_npar(id);
    }
    long long off = 0, loff = 0; int start = 0;
    for (auto& b : cv->blocks) {
        if (b.n == 0) { delete cv; return gmb_set_error(GMB_EINVAL, "block ids must be 0..B-1 without gaps"); }
        b.data0 = off; b.start = start; b.l0 = loff;
        off += (long long)b.n * b.ncol; start += b.n;
        long long sz = (long long)gmb_cov_ld(b.n) * b.n;     // packed for n <= 32, ld = round_up(n, 4) otherwise
        loff += (sz + 1) & ~1LL;                              // keep every block 16-byte aligned
        if (b.n > cv->max_n) cv->max_n = b.n;
    }
    if (off > n_data) { delete cv; return gmb_set_error(GMB_EINVAL, "covariance data has %d values, %lld required", n_data, off); }
    if (cv->R > 64) { const int R = cv->R; delete cv; return gmb_set_error(GMB_EINVAL, "%d covariance parameters exceed the staging area (64)", R); }
    cv->Q = start;
    cv->lblk_doubles = loff;
    cudaSetDevice(ctx->device);
    GMB_CUDA(gmb_dmalloc(ctx, &cv->d_blocks, sizeof(CovBlock) * cv->B));
    GMB_CUDA(gmb_dmalloc(ctx, &cv->d_fns, sizeof(CovFn) * cv->fns.size()));
    GMB_CUDA(gmb_dmalloc(ctx, &cv->d_data, sizeof(double) * (off > 0 ? off : 1)));
    GMB_CUDA(gmb_dmalloc(ctx, &cv->d_theta, sizeof(double) * (cv->R > 0 ? cv->R : 1)));
    GMB_CUDA(gmb_dmalloc(ctx, &cv->d_Lblk, sizeof(double) * (loff > 0 ? loff : 1)));
    GMB_CUDA(gmb_dmalloc(ctx, &cv->d_logdet, sizeof(double) * cv->B));
    GMB_CUDA(gmb_dmalloc(ctx, &cv->d_status, 2 * sizeof(int)));      // [1]: the Gram factorisations of large blocks
    GMB_CUDA(cudaMemcpyAsync(cv->d_blocks, cv->blocks.data(), sizeof(CovBlock) * cv->B, cudaMemcpyHostToDevice, ctx->stream));
    GMB_CUDA(cudaMemcpyAsync(cv->d_fns, cv->fns.data(), sizeof(CovFn) * cv->fns.size(), cudaMemcpyHostToDevice, ctx->stream));
    GMB_CUDA(cudaMemcpyAsync(cv->d_data, data, sizeof(double) * off, cudaMemcpyHostToDevice, ctx->stream));
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (cv->max_n <= QUAD_SMALL_MAX && cv->B > 1) {
        // classes of identical blocks: exact comparison of size, function rows (id, variables, parameter offsets, effective range) and data differences
        std::map<std::string, int> seen;
        std::vector<int> rep, cls(cv->B);
        for (int b = 0; b < cv->B; b++) {
            const CovBlock& blk = cv->blocks[b];
            std::string key(reinterpret_cast<const char*>(&blk.n), sizeof(int));
            for (int f = 0; f < blk.nfn; f++) {
                const CovFn& fn = cv->fns[blk.fn0 + f];
                key.append(reinterpret_cast<const char*>(&fn.id), sizeof(int)); key.append(reinterpret_cast<const char*>(&fn.nvar), sizeof(int));
                key.append(reinterpret_cast<const char*>(&fn.par0), sizeof(int)); key.append(reinterpret_cast<const char*>(&fn.col0), sizeof(int));
                key.append(reinterpret_cast<const char*>(&fn.eff), sizeof(double));
            }
            // the functions see the data only through the differences dat[i] - dat[j] of each column (block_val): equal differences, computed
            // exactly as the kernels compute them, give bitwise equal blocks — gr(cl) * ar1(t) carries a different label per cluster, same block
            const double* dat = data + blk.data0;
            for (int c = 0; c < blk.ncol; c++)
                for (int i = 1; i < blk.n; i++)
                    for (int j = 0; j < i; j++) {
                        double di = dat[i + (size_t)c * blk.n] - dat[j + (size_t)c * blk.n];
                        if (di == 0.0) di = 0.0;                               // -0.0
                        key.append(reinterpret_cast<const char*>(&di), sizeof(double));
                    }
            auto it = seen.find(key);
            if (it == seen.end()) { it = seen.emplace(std::move(key), (int)rep.size()).first; rep.push_back(b); }
            cls[b] = it->second;
        }
        if ((int)rep.size() < cv->B) {
            cv->ncls = (int)rep.size();
            std::vector<int> ptr(cv->ncls + 1, 0), mem(cv->B);
            for (int b = 0; b < cv->B; b++) ptr[cls[b] + 1]++;
            for (int c = 0; c < cv->ncls; c++) ptr[c + 1] += ptr[c];
            std::vector<int> fill(ptr.begin(), ptr.end() - 1);
            for (int b = 0; b < cv->B; b++) mem[fill[cls[b]]++] = b;
            GMB_CUDA(gmb_dmalloc(ctx, &cv->d_cls_rep, sizeof(int) * cv->ncls));
            GMB_CUDA(gmb_dmalloc(ctx, &cv->d_cls_ptr, sizeof(int) * (cv->ncls + 1)));
            GMB_CUDA(gmb_dmalloc(ctx, &cv->d_cls_mem, sizeof(int) * cv->B));
            GMB_CUDA(cudaMemcpyAsync(cv->d_cls_rep, rep.data(), sizeof(int) * cv->ncls, cudaMemcpyHostToDevice, ctx->stream));
            GMB_CUDA(cudaMemcpyAsync(cv->d_cls_ptr, ptr.data(), sizeof(int) * (cv->ncls + 1), cudaMemcpyHostToDevice, ctx->stream));
            GMB_CUDA(cudaMemcpyAsync(cv->d_cls_mem, mem.data(), sizeof(int) * cv->B, cudaMemcpyHostToDevice, ctx->stream));
            GMB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
    *out = cv;
    return GMB_OK;
}

static int g_cov_classes = 1;
extern "C" int gmb_cov_set_block_classes(int on) { g_cov_classes = on ? 1 : 0; return GMB_OK; }
extern "C" int gmb_cov_block_classes(gmb_cov* cv, int* ncls) {
    if (!cv || !ncls) return gmb_set_error(GMB_EINVAL, "gmb_cov_block_classes: bad arguments");
    *ncls = cv->ncls > 0 ? cv->ncls : cv->B;
    return GMB_OK;
}

extern "C" void gmb_cov_destroy(gmb_cov* cv) {
    if (!cv) return;
    cudaSetDevice(cv->ctx->device);
    cudaStreamSynchronize(cv->ctx->stream);
    gmb_dfree(cv->ctx, cv->d_blocks); gmb_dfree(cv->ctx, cv->d_fns); gmb_dfree(cv->ctx, cv->d_data); gmb_dfree(cv->ctx, cv->d_theta); gmb_dfree(cv->ctx, cv->d_Lblk);
    gmb_dfree(cv->ctx, cv->d_logdet); gmb_dfree(cv->ctx, cv->d_status);
    if (cv->dU) gmb_dfree(cv->ctx, cv->dU);
    if (cv->d_work) gmb_dfree(cv->ctx, cv->d_work);
    if (cv->d_linv) gmb_dfree(cv->ctx, cv->d_linv);
    if (cv->d_x512) gmb_dfree(cv->ctx, cv->d_x512);
    gmb_cov_gram_large_free(cv);
    if (cv->d_gram) gmb_dfree(cv->ctx, cv->d_gram);
    if (cv->d_gram_cls) gmb_dfree(cv->ctx, cv->d_gram_cls);
    gmb_dfree(cv->ctx, cv->d_cls_rep); gmb_dfree(cv->ctx, cv->d_cls_ptr); gmb_dfree(cv->ctx, cv->d_cls_mem);
    if (cv->d_batch) gmb_dfree(cv->ctx, cv->d_batch);
    delete cv;
}

extern "C" int gmb_cov_dims(gmb_cov* cv, int* B, int* Q, int* R) {
    if (!cv) return gmb_set_error(GMB_EINVAL, "cov is NULL");
    if (B) *B = cv->B;
    if (Q) *Q = cv->Q;
    if (R) *R = cv->R;
    return GMB_OK;
}

// Builds and factorises every block for `theta` (cached: repeated calls with the same theta are free).
int gmb_cov_factor(gmb_cov* cv, const double* theta) {
    gmb_ctx* ctx = cv->ctx;
    if (cv->factor_valid && (int)cv->theta_cached.size() == cv->R &&
        memcmp(cv->theta_cached.data(), theta, sizeof(double) * cv->R) == 0)
        return GMB_OK;
    cv->factor_valid = false;
    for (int r = 0; r < cv->R; r++) ctx->h_pinned[r] = theta[r];
    GMB_CUDA(cudaMemcpyAsync(cv->d_theta, ctx->h_pinned, sizeof(double) * cv->R, cudaMemcpyHostToDevice, ctx->stream));
    GMB_CUDA(cudaMemsetAsync(cv->d_status, 0, sizeof(int), ctx->stream));
    bool any_small = false;
    for (const auto& b : cv->blocks) if (b.n <= SMALL_MAX) { any_small = true; break; }
    if (any_small) {
        factor_small_kernel<<<(cv->B + FACT_WARPS - 1) / FACT_WARPS, FACT_WARPS * 32, 0, ctx->stream>>>(
            cv->B, cv->d_blocks, cv->d_fns, cv->d_data, cv->d_theta, cv->d_Lblk, cv->d_logdet, cv->d_status);
        ctx->launches++;
        GMB_CUDA(cudaGetLastError());
    }
    for (int bi = 0; bi < cv->B; bi++)
        if (cv->blocks[bi].n > SMALL_MAX) GMB_TRY(gmb_cov_factor_large(cv, bi));
    int* hstat = reinterpret_cast<int*>(ctx->h_pinned + 64);
    GMB_CUDA(cudaMemcpyAsync(hstat, cv->d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (*hstat != 0) return gmb_set_error(GMB_ENOTPD, "D(theta) is not positive definite: pivot %d", *hstat - 1);
    cv->theta_cached.assign(theta, theta + cv->R);
    cv->factor_valid = true;
    return GMB_OK;
}

// d_out[0] = sum over the given columns of sum_b log N(u_j[b]; 0, D_b)   (raw sum; caller divides by m)
static int g_cov_gram = 1;
extern "C" int gmb_cov_set_gram(int on) { g_cov_gram = on ? 1 : 0; return GMB_OK; }

// gram_mdl != NULL: dU are that model's device-resident samples; large blocks then go through the Cholesky factor of their Gram matrix
// (gmb_cov_gram_large) when there are at least twice as many samples as rows, this rank holds all of them and the Gram path is on
int gmb_cov_quad(gmb_cov* cv, const double* dU, int ldu, int ncols, double* d_out, gmb_model* gram_mdl) {
    gmb_ctx* ctx = cv->ctx;
    if (ncols <= 0) { GMB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double), ctx->stream)); return GMB_OK; }
    int n_small = 0, n_other = 0;
    for (const auto& b : cv->blocks) { if (b.n <= QUAD_SMALL_MAX) n_small++; else n_other++; }
    // Gram factors of the large blocks first: building one uses the context's scratch area, which the partial sums below live in
    std::vector<const double*> gramC(cv->B, nullptr);
    std::vector<int> gramLd(cv->B, 0);
    if (gram_mdl && g_cov_gram && ctx->world == 1)
        for (int bi = 0; bi < cv->B; bi++)
            if (cv->blocks[bi].n > QUAD_MED_MAX && ncols >= 2 * cv->blocks[bi].n) GMB_TRY(gmb_cov_gram_large(cv, bi, gram_mdl, &gramC[bi], &gramLd[bi]));
    int groups = (cv->B + 31) / 32;
    int CC = 1, cols_per_cta = ncols;
    if (n_small) {
        int want = (ctx->sms * 4 + groups - 1) / groups;
        int maxcc = (ncols + 15) / 16;
        CC = want < maxcc ? want : maxcc; if (CC < 1) CC = 1;
        cols_per_cta = (ncols + CC - 1) / CC;
        CC = (ncols + cols_per_cta - 1) / cols_per_cta;
    }
    const int MED_CTAS = 64;
    size_t npart_small = n_small ? (size_t)groups * CC : 0;
    size_t npart = npart_small + (size_t)n_other * MED_CTAS;
    GMB_TRY(gmb_ctx_scratch(ctx, npart));
    double* partials = ctx->d_scratch;
    if (n_other) GMB_CUDA(cudaMemsetAsync(partials + npart_small, 0, sizeof(double) * n_other * MED_CTAS, ctx->stream));
    if (n_small) {
        size_t smem = (size_t)32 * (QUAD_SMALL_MAX * QUAD_SMALL_MAX + 1) * sizeof(double);
        static bool configured = false;
        if (!configured) { GMB_CUDA(cudaFuncSetAttribute(quad_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured = true; }
        quad_small_kernel<<<dim3(groups, CC), 256, smem, ctx->stream>>>(cv->B, cv->d_blocks, cv->d_Lblk, dU, ldu, ncols, cols_per_cta, partials);
        ctx->launches++;
        GMB_CUDA(cudaGetLastError());
    }
    int k = 0;
    for (int bi = 0; bi < cv->B; bi++) {
        const CovBlock& b = cv->blocks[bi];
        if (b.n <= QUAD_SMALL_MAX) continue;
        double* dst = partials + npart_small + (size_t)k * MED_CTAS;
        if (b.n <= QUAD_MED_MAX) {
            size_t smem = ((size_t)b.n * b.n + (size_t)b.n * 128) * sizeof(double);
            static bool configured = false;
            if (!configured) { GMB_CUDA(cudaFuncSetAttribute(quad_medium_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((64 * 64 + 64 * 128) * sizeof(double)))); configured = true; }
            int ctas = (ncols + 127) / 128; if (ctas > MED_CTAS) ctas = MED_CTAS;
            quad_medium_kernel<<<ctas, 128, smem, ctx->stream>>>(b, cv->d_Lblk, dU, ldu, ncols, dst);
            ctx->launches++;
            GMB_CUDA(cudaGetLastError());
        } else {
            if (gramC[bi]) GMB_TRY(gmb_cov_quad_large(cv, bi, gramC[bi], gramLd[bi], b.n, dst, 1));
            else GMB_TRY(gmb_cov_quad_large(cv, bi, dU, ldu, ncols, dst, 0));
        }
        k++;
    }
    mvn_finish_kernel<<<1, 256, 0, ctx->stream>>>(cv->B, cv->d_blocks, cv->d_logdet, partials, (int)npart, ncols, d_out);
    ctx->launches++;
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}

// CTAs per block of expand_blocks_kernel: one for small blocks, enough to fill the machine for a large one
static int expand_chunks(const gmb_cov* cv) {
    const long long per = ((long long)cv->max_n * cv->max_n + 256 * 8 - 1) / (256 * 8);
    return (int)std::max<long long>(1, std::min<long long>(per, 2048));
}

// dense Q x Q D(theta) (chol = 0) or its lower Cholesky factor (chol = 1) into a device matrix with leading dimension ld
// (zero outside the blocks).  Returns GMB_ENOTPD when a block is not positive definite.
int gmb_cov_gen_device(gmb_cov* cv, const double* theta, int chol, double* d_out, int ld) {
    gmb_ctx* ctx = cv->ctx;
    if (chol) GMB_TRY(gmb_cov_factor(cv, theta));
    else {
        for (int r = 0; r < cv->R; r++) ctx->h_pinned[r] = theta[r];
        GMB_CUDA(cudaMemcpyAsync(cv->d_theta, ctx->h_pinned, sizeof(double) * cv->R, cudaMemcpyHostToDevice, ctx->stream));
        cv->factor_valid = false;
    }
    GMB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t)ld * cv->Q, ctx->stream));
    expand_blocks_kernel<<<dim3(cv->B, expand_chunks(cv)), 256, 0, ctx->stream>>>(cv->B, cv->d_blocks, cv->d_fns, cv->d_data, cv->d_theta, cv->d_Lblk, chol, d_out, ld);
    ctx->launches++;
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}

extern "C" int gmb_cov_gen(gmb_cov* cv, const double* theta, int chol, double* L_out) {
    if (!cv || !theta) return gmb_set_error(GMB_EINVAL, "gmb_cov_gen: bad arguments");
    gmb_ctx* ctx = cv->ctx;
    cudaSetDevice(ctx->device);
    if (chol) GMB_TRY(gmb_cov_factor(cv, theta));
    else {
        for (int r = 0; r < cv->R; r++) ctx->h_pinned[r] = theta[r];
        GMB_CUDA(cudaMemcpyAsync(cv->d_theta, ctx->h_pinned, sizeof(double) * cv->R, cudaMemcpyHostToDevice, ctx->stream));
        cv->factor_valid = false;
    }
    if (!L_out) return GMB_OK;
    size_t Q = cv->Q;
    double* dense = nullptr;
    GMB_CUDA(gmb_dmalloc(ctx, &dense, sizeof(double) * Q * Q));
    GMB_CUDA(cudaMemsetAsync(dense, 0, sizeof(double) * Q * Q, ctx->stream));
    expand_blocks_kernel<<<dim3(cv->B, expand_chunks(cv)), 256, 0, ctx->stream>>>(cv->B, cv->d_blocks, cv->d_fns, cv->d_data, cv->d_theta, cv->d_Lblk, chol, dense, (int)Q);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(L_out, dense, sizeof(double) * Q * Q, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    gmb_dfree(ctx, dense);
    if (e != cudaSuccess) return gmb_set_error(GMB_ECUDA, "gmb_cov_gen: %s", cudaGetErrorString(e));
    return GMB_OK;
}

static int cov_upload_u(gmb_cov* cv, const double* U, int Q, int m, int* ldu) {
    gmb_ctx* ctx = cv->ctx;
    int ld = round_up(Q, 4);
    size_t need = (size_t)ld * (m > 0 ? m : 1);
    if (need > cv->dU_doubles) {
        if (cv->dU) { GMB_CUDA(cudaStreamSynchronize(ctx->stream)); gmb_dfree(ctx, cv->dU); cv->dU = nullptr; }
        GMB_CUDA(gmb_dmalloc(ctx, &cv->dU, need * sizeof(double)));
        cv->dU_doubles = need;
    }
    if (m > 0)
        GMB_CUDA(cudaMemcpy2DAsync(cv->dU, ld * sizeof(double), U, Q * sizeof(double), Q * sizeof(double), m, cudaMemcpyHostToDevice, ctx->stream));
    *ldu = ld;
    return GMB_OK;
}

static int cov_finish_ll(gmb_cov* cv, double* d_out, int m_total, double* out) {
    gmb_ctx* ctx = cv->ctx;
    GMB_TRY(gmb_comm_allreduce_dev(ctx, d_out, 1));
    GMB_CUDA(cudaMemcpyAsync(ctx->h_pinned, d_out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = ctx->h_pinned[0] / m_total;      // mcmldmatrix.h:40
    return GMB_OK;
}

// 1 = mvn_ll on a model's samples goes through the Gram matrices when every block is <= 16 (default); 0 = always stream U

// Gram matrices S_b = sum_j u_bj u_bj' of the model's local sample columns, cached per (model, sample version)
static int cov_ensure_gram(gmb_cov* cv, gmb_model* mdl) {
    gmb_ctx* ctx = cv->ctx;
    if (cv->gram_model == mdl && cv->gram_version == mdl->u_version && cv->gram_cols == mdl->m_local && cv->d_gram) return GMB_OK;
    const long long gd = cv->lblk_doubles;
    if (!cv->d_gram) GMB_CUDA(gmb_dmalloc(ctx, &cv->d_gram, sizeof(double) * (gd > 0 ? gd : 1)));
    const int ncols = mdl->m_local;
    const int gx = (cv->B + GRAM_WARPS - 1) / GRAM_WARPS;
    int CC = (ctx->sms * 8 + gx - 1) / gx;
    const int max_cc = (ncols + 63) / 64;
    if (CC > max_cc) CC = max_cc;
    if (CC < 1) CC = 1;
    const int cols_per_cta = (ncols + CC - 1) / CC;
    CC = (ncols + cols_per_cta - 1) / cols_per_cta;
    GMB_TRY(gmb_ctx_scratch(ctx, (size_t)CC * gd));
    GMB_CUDA(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(double) * (size_t)CC * gd, ctx->stream));
    gram_small_kernel<<<dim3(gx, CC), GRAM_WARPS * 32, 0, ctx->stream>>>(cv->B, cv->d_blocks, mdl->dU, mdl->ldq, ncols, cols_per_cta, gd, ctx->d_scratch);
    gram_reduce_kernel<<<(unsigned)((gd + 255) / 256), 256, 0, ctx->stream>>>(gd, CC, ctx->d_scratch, cv->d_gram);
    ctx->launches += 2;
    if (cv->ncls > 0) {
        if (!cv->d_gram_cls) GMB_CUDA(gmb_dmalloc(ctx, &cv->d_gram_cls, sizeof(double) * (gd > 0 ? gd : 1)));
        gram_class_kernel<<<cv->ncls, 128, 0, ctx->stream>>>(cv->ncls, cv->d_blocks, cv->d_cls_rep, cv->d_cls_ptr, cv->d_cls_mem, cv->d_gram, cv->d_gram_cls);
        ctx->launches++;
    }
    GMB_CUDA(cudaGetLastError());
    cv->gram_model = mdl; cv->gram_version = mdl->u_version; cv->gram_cols = ncols;
    return GMB_OK;
}

extern "C" int gmb_cov_mvn_ll(gmb_cov* cv, const double* theta, const double* U, int Q, int m_local, int m_total, double* out) {
    if (!cv || !theta || !out || (m_local > 0 && !U)) return gmb_set_error(GMB_EINVAL, "gmb_cov_mvn_ll: bad arguments");
    if (Q != cv->Q) return gmb_set_error(GMB_EINVAL, "u has %d rows, covariance has %d", Q, cv->Q);
    if (m_total <= 0 || m_local < 0) return gmb_set_error(GMB_EINVAL, "bad column counts");
    cudaSetDevice(cv->ctx->device);
    GMB_TRY(gmb_cov_factor(cv, theta));
    int ldu;
    GMB_TRY(cov_upload_u(cv, U, Q, m_local, &ldu));
    GMB_TRY(gmb_cov_quad(cv, cv->dU, ldu, m_local, cv->ctx->d_result));
    return cov_finish_ll(cv, cv->ctx->d_result, m_total, out);
}

extern "C" int gmb_cov_mvn_ll_model(gmb_cov* cv, const double* theta, gmb_model* mdl, int ncols_total, double* out) {
    if (!cv || !theta || !mdl || !out) return gmb_set_error(GMB_EINVAL, "gmb_cov_mvn_ll_model: bad arguments");
    if (mdl->Q != cv->Q) return gmb_set_error(GMB_EINVAL, "model has Q=%d, covariance has %d", mdl->Q, cv->Q);
    if (!mdl->dU || mdl->m_local < 0) return gmb_set_error(GMB_ESTATE, "the model holds no samples (call gmb_model_set_u or gmb_hmc_sample first)");
    cudaSetDevice(cv->ctx->device);
    if (cv->max_n <= QUAD_SMALL_MAX && g_cov_gram && mdl->m_local > 0) {
        // sufficient-statistics path: Gram matrices of the model's samples (cached), one launch per evaluation
        gmb_ctx* ctx = cv->ctx;
        for (int r = 0; r < cv->R; r++) if (!(theta[r] == theta[r])) return gmb_set_error(GMB_ENOTPD, "D(theta) is not positive definite: theta is NaN");
        GMB_TRY(cov_ensure_gram(cv, mdl));
        GMB_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int r = 0; r < cv->R; r++) ctx->h_pinned[r] = theta[r];
        GMB_CUDA(cudaMemcpyAsync(cv->d_theta, ctx->h_pinned, sizeof(double) * cv->R, cudaMemcpyHostToDevice, ctx->stream));
        GMB_CUDA(cudaMemsetAsync(cv->d_status, 0, sizeof(int), ctx->stream));
        const bool by_class = cv->ncls > 0 && g_cov_classes;       // one factorisation per class of identical blocks
        const int nb = by_class ? cv->ncls : cv->B;
        const int ctas = (nb + FACT_WARPS - 1) / FACT_WARPS;
        GMB_TRY(gmb_ctx_scratch(ctx, (size_t)ctas));
        mvn_gram_kernel<<<ctas, FACT_WARPS * 32, 0, ctx->stream>>>(nb, cv->d_blocks, cv->d_fns, cv->d_data, cv->d_theta, cv->R,
                                                                  by_class ? cv->d_gram_cls : cv->d_gram, (double)mdl->m_local,
                                                                  by_class ? nullptr : cv->d_Lblk, by_class ? nullptr : cv->d_logdet, cv->d_status,
                                                                  ctx->d_scratch, ctx->d_counter, ctx->d_result, by_class ? cv->d_cls_rep : nullptr,
                                                                  by_class ? cv->d_cls_ptr : nullptr);
        ctx->launches++;
        GMB_CUDA(cudaGetLastError());
        if (by_class) cv->factor_valid = false;                                          // the per-block factor cache was not refreshed
        else { cv->theta_cached.assign(theta, theta + cv->R); cv->factor_valid = true; }  // Lblk / logdet now hold this theta's factor
        GMB_TRY(gmb_comm_allreduce_dev(ctx, ctx->d_result, 1));
        int* hstat = reinterpret_cast<int*>(ctx->h_pinned + 64);
        GMB_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_result, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        GMB_CUDA(cudaMemcpyAsync(hstat, cv->d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GMB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (*hstat != 0) { cv->factor_valid = false; return gmb_set_error(GMB_ENOTPD, "D(theta) is not positive definite: pivot %d", *hstat - 1); }
        *out = ctx->h_pinned[0] / (ncols_total > 0 ? ncols_total : mdl->m_total);   // mcmldmatrix.h:40
        return GMB_OK;
    }
    GMB_TRY(gmb_cov_factor(cv, theta));
    GMB_TRY(gmb_cov_quad(cv, mdl->dU, mdl->ldq, mdl->m_local, cv->ctx->d_result, mdl));
    return cov_finish_ll(cv, cv->ctx->d_result, ncols_total > 0 ? ncols_total : mdl->m_total, out);
}

// k evaluations at the columns of thetas (R x k) with one launch and one synchronisation (sufficient-statistics path); otherwise one
// after the other.  An evaluation whose D(theta) is not positive definite yields -inf instead of an error.
extern "C" int gmb_cov_mvn_ll_model_batch(gmb_cov* cv, const double* thetas, int k, gmb_model* mdl, int ncols_total, double* out) {
    if (!cv || !thetas || !mdl || !out || k < 0) return gmb_set_error(GMB_EINVAL, "gmb_cov_mvn_ll_model_batch: bad arguments");
    if (k == 0) return GMB_OK;
    const int R = cv->R;
    if (!(cv->max_n <= QUAD_SMALL_MAX && g_cov_gram && mdl->dU && mdl->m_local > 0 && mdl->Q == cv->Q)) {
        for (int e = 0; e < k; e++) {
            int rc = gmb_cov_mvn_ll_model(cv, thetas + (size_t)e * R, mdl, ncols_total, out + e);
            if (rc == GMB_ENOTPD) { out[e] = -INFINITY; rc = GMB_OK; }
            GMB_TRY(rc);
        }
        return GMB_OK;
    }
    gmb_ctx* ctx = cv->ctx;
    cudaSetDevice(ctx->device);
    GMB_TRY(cov_ensure_gram(cv, mdl));
    const bool by_class = cv->ncls > 0 && g_cov_classes;
    const int nb = by_class ? cv->ncls : cv->B;
    const int ctas = (nb + FACT_WARPS - 1) / FACT_WARPS;
    const double denom = (double)(ncols_total > 0 ? ncols_total : mdl->m_total);
    for (int off = 0; off < k; off += 2048) {
        const int kb = std::min(2048, k - off);
        // device layout: theta [R kb] | out [kb] | partials [ctas kb] | status [kb] ints | counters [kb] uints
        const size_t nd = (size_t)R * kb + kb + (size_t)ctas * kb, need = nd * sizeof(double) + (size_t)kb * 8;
        if (need > cv->batch_bytes) {
            if (cv->d_batch) { GMB_CUDA(cudaStreamSynchronize(ctx->stream)); gmb_dfree(ctx, cv->d_batch); cv->d_batch = nullptr; cv->batch_bytes = 0; }
            GMB_CUDA(gmb_dmalloc(ctx, &cv->d_batch, need));
            cv->batch_bytes = need;
        }
        if ((size_t)(R + 1) * kb + kb > ctx->pinned_doubles) return gmb_set_error(GMB_EINVAL, "batch too large");
        double* d_th = cv->d_batch; double* d_out = d_th + (size_t)R * kb; double* d_part = d_out + kb;
        int* d_stat = reinterpret_cast<int*>(d_part + (size_t)ctas * kb); unsigned int* d_cnt = reinterpret_cast<unsigned int*>(d_stat + kb);
        GMB_CUDA(cudaStreamSynchronize(ctx->stream));
        memcpy(ctx->h_pinned, thetas + (size_t)off * R, sizeof(double) * R * kb);
        GMB_CUDA(cudaMemcpyAsync(d_th, ctx->h_pinned, sizeof(double) * R * kb, cudaMemcpyHostToDevice, ctx->stream));
        GMB_CUDA(cudaMemsetAsync(d_stat, 0, (size_t)kb * 8, ctx->stream));
        mvn_gram_kernel<<<dim3(ctas, kb), FACT_WARPS * 32, 0, ctx->stream>>>(nb, cv->d_blocks, cv->d_fns, cv->d_data, d_th, R,
                                                                             by_class ? cv->d_gram_cls : cv->d_gram, (double)mdl->m_local, nullptr, nullptr,
                                                                             d_stat, d_part, d_cnt, d_out, by_class ? cv->d_cls_rep : nullptr,
                                                                             by_class ? cv->d_cls_ptr : nullptr);
        ctx->launches++;
        GMB_CUDA(cudaGetLastError());
        GMB_TRY(gmb_comm_allreduce_dev(ctx, d_out, kb));
        double* h_out = ctx->h_pinned + (size_t)R * kb; int* h_stat = reinterpret_cast<int*>(h_out + kb);
        GMB_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof(double) * kb, cudaMemcpyDeviceToHost, ctx->stream));
        GMB_CUDA(cudaMemcpyAsync(h_stat, d_stat, sizeof(int) * kb, cudaMemcpyDeviceToHost, ctx->stream));
        GMB_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int e = 0; e < kb; e++) out[off + e] = (h_stat[e] != 0 || !(h_out[e] == h_out[e])) ? -INFINITY : h_out[e] / denom;   // mcmldmatrix.h:40
    }
    return GMB_OK;
}

extern "C" int gmb_cov_logdet(gmb_cov* cv, const double* theta, double* out) {
    if (!cv || !theta || !out) return gmb_set_error(GMB_EINVAL, "gmb_cov_logdet: bad arguments");
    gmb_ctx* ctx = cv->ctx;
    cudaSetDevice(ctx->device);
    GMB_TRY(gmb_cov_factor(cv, theta));
    logdet_sum_kernel<<<1, 256, 0, ctx->stream>>>(cv->B, cv->d_logdet, ctx->d_result);
    ctx->launches++;
    GMB_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_result, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = ctx->h_pinned[0];
    return GMB_OK;
}
