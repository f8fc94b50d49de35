// model.cu — host side of gmb_model: the device-resident replacement of glmmr::mcmlModel
// (inst/include/glmmrmcml/mcmlmodel.h:28-307).
//
// What the reference object caches per evaluation is hoisted here to "once per sample matrix":
//   zd = Z u  is built once in gmb_model_set_u (the reference rebuilds it in every log_likelihood call,
//   mcmlmodel.h:286, and m times per mcnr() call, mcmloptim.h:213 -> mcmlmodel.h:121).
// The dense n x n weight matrix W_ (mcmlmodel.h:36,62) is never materialised.
#include "common.cuh"
#include <algorithm>

namespace {

// family+link -> flink, mcmlmodel.h:74-87
int flink_from_strings(const char* family, const char* link) {
    static const char* keys[12] = {"poissonlog", "poissonidentity", "binomiallogit", "binomiallog",
                                   "binomialidentity", "binomialprobit", "gaussianidentity", "gaussianlog",
                                   "gammalog", "gammainverse", "gammaidentity", "betalogit"};
    std::string k = std::string(family ? family : "") + std::string(link ? link : "");
    for (int i = 0; i < 12; i++) if (k == keys[i]) return i + 1;
    return 0;
}

__global__ void rowc_kernel(int n, int flink, const double* __restrict__ y, double* __restrict__ rowc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rowc[i] = (flink == 1 || flink == 2) ? dev_log_factorial_approx(y[i]) : 0.0;   // moremaths.h:34-39, :43
}

// gaussian/log: the constructor replaces y by log y (mcmlmodel.h:90-92)
__global__ void logy_kernel(int n, double* __restrict__ y) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = log(y[i]);
}

int upload_matrix(gmb_ctx* ctx, double* dst, int ld, const double* src, int rows, int cols) {
    if (rows == 0 || cols == 0) return GMB_OK;
    GMB_CUDA(cudaMemcpy2DAsync(dst, (size_t)ld * sizeof(double), src, (size_t)rows * sizeof(double),
                               (size_t)rows * sizeof(double), cols, cudaMemcpyHostToDevice, ctx->stream));
    return GMB_OK;
}

int upload_params(gmb_model* mdl, const double* beta, int count) {
    gmb_ctx* ctx = mdl->ctx;
    if ((size_t)count > ctx->pinned_doubles / 2) return gmb_set_error(GMB_EINVAL, "too many parameter values in one call (%d)", count);
    if (count > mdl->beta_cap) {
        if (mdl->dbeta) { GMB_CUDA(cudaStreamSynchronize(ctx->stream)); gmb_dfree(ctx, mdl->dbeta); mdl->dbeta = nullptr; }
        GMB_CUDA(gmb_dmalloc(ctx, &mdl->dbeta, sizeof(double) * count));
        mdl->beta_cap = count;
    }
    // the pinned staging area is reused by every call: make sure the previous copy has been consumed
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(ctx->h_pinned, beta, sizeof(double) * count);
    GMB_CUDA(cudaMemcpyAsync(mdl->dbeta, ctx->h_pinned, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->stream));
    return GMB_OK;
}

}  // namespace

extern "C" int gmb_model_create(gmb_ctx* ctx, int n, int P, int Q, const double* X, const double* Z, const double* y,
                                const char* family, const char* link, gmb_model** out) {
    return gmb_model_create_prec(ctx, n, P, Q, X, Z, y, family, link, 64, out);
}

extern "C" int gmb_model_create_prec(gmb_ctx* ctx, int n, int P, int Q, const double* X, const double* Z, const double* y,
                                     const char* family, const char* link, int precision, gmb_model** out) {
    if (precision != 64 && precision != 32) return gmb_set_error(GMB_EINVAL, "precision must be 64 or 32 (got %d)", precision);
    if (!ctx || !out || !X || !Z || !y || n <= 0 || P <= 0 || Q <= 0)
        return gmb_set_error(GMB_EINVAL, "gmb_model_create: bad arguments (n=%d P=%d Q=%d)", n, P, Q);
    int fl = flink_from_strings(family, link);
    if (fl == 0)
        return gmb_set_error(GMB_EFAMILY, "unknown family/link '%s'/'%s' (mcmlmodel.h:74-87 lists the valid pairs)",
                             family ? family : "", link ? link : "");
    if (precision == 32 && !gmb_flink_core(fl))
        return gmb_set_error(GMB_EFAMILY, "fp32 mode is implemented for poisson/log, binomial/logit and gaussian/identity (got code %d)", fl);
    if (!gmb_flink_supported(fl))
        return gmb_set_error(GMB_EFAMILY, "family/link '%s'/'%s' (code %d) has no device kernel (supported: codes 1-8; the Gamma codes are "
                             "unreachable in the reference, whose family string is 'Gamma', and beta/logit is not implemented)", family, link, fl);
    GMB_CUDA(cudaSetDevice(ctx->device));
    gmb_model* mdl = new gmb_model();
    mdl->ctx = ctx; mdl->n = n; mdl->P = P; mdl->Q = Q; mdl->flink = fl; mdl->prec = precision;
    mdl->ldn = round_up(n, 4); mdl->ldq = round_up(Q, 4);
    const size_t ldn = mdl->ldn;
    cudaError_t e = cudaSuccess;
    auto alloc0 = [&](double** p, size_t doubles) {
        if (e != cudaSuccess) return;
        e = gmb_dmalloc(ctx, p, doubles * sizeof(double));
        if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, doubles * sizeof(double), ctx->stream);
    };
    alloc0(&mdl->dX, ldn * P); alloc0(&mdl->dZ, ldn * Q); alloc0(&mdl->dy, ldn); alloc0(&mdl->drowc, ldn); alloc0(&mdl->dxb, ldn);
    if (e != cudaSuccess) { gmb_model_destroy(mdl); return gmb_set_error(GMB_ECUDA, "gmb_model_create: %s", cudaGetErrorString(e)); }
    int rc = upload_matrix(ctx, mdl->dX, mdl->ldn, X, n, P);
    if (!rc) rc = upload_matrix(ctx, mdl->dZ, mdl->ldn, Z, n, Q);
    if (!rc) rc = upload_matrix(ctx, mdl->dy, mdl->ldn, y, n, 1);
    if (rc) { gmb_model_destroy(mdl); return rc; }
    rowc_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, fl, mdl->dy, mdl->drowc);
    ctx->launches++;
    if (fl == 8) { logy_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, mdl->dy); ctx->launches++; }
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { gmb_model_destroy(mdl); return gmb_set_error(GMB_ECUDA, "gmb_model_create: %s", cudaGetErrorString(e)); }
    *out = mdl;
    return GMB_OK;
}

extern "C" void gmb_model_destroy(gmb_model* mdl) {
    if (!mdl) return;
    cudaSetDevice(mdl->ctx->device);
    cudaStreamSynchronize(mdl->ctx->stream);
    gmb_dfree(mdl->ctx, mdl->dX); gmb_dfree(mdl->ctx, mdl->dZ); gmb_dfree(mdl->ctx, mdl->dy); gmb_dfree(mdl->ctx, mdl->drowc); gmb_dfree(mdl->ctx, mdl->dxb); gmb_dfree(mdl->ctx, mdl->dbeta);
    gmb_dfree(mdl->ctx, mdl->dstat);
    gmb_agg_free(mdl);
    gmb_ell_free(mdl);
    gmb_comp_free(mdl);
    gmb_lane_free(mdl);
    gmb_dfree(mdl->ctx, mdl->dzd32); gmb_dfree(mdl->ctx, mdl->dF32);
    gmb_dfree(mdl->ctx, mdl->dZt_hi); gmb_dfree(mdl->ctx, mdl->dZt_lo); gmb_dfree(mdl->ctx, mdl->dU_hi); gmb_dfree(mdl->ctx, mdl->dU_lo);
    gmb_dfree(mdl->ctx, mdl->dU); gmb_dfree(mdl->ctx, mdl->dzd); gmb_dfree(mdl->ctx, mdl->dF); gmb_dfree(mdl->ctx, mdl->dZL); gmb_dfree(mdl->ctx, mdl->dL);
    gmb_dfree(mdl->ctx, mdl->dV); gmb_dfree(mdl->ctx, mdl->hmc_work);
    delete mdl;
}

extern "C" int gmb_model_flink(gmb_model* mdl) { return mdl ? mdl->flink : 0; }

// grow the sample buffers (dU: ldq x m_cap, dzd: ldn x m_cap); contents are NOT preserved
int gmb_model_reserve_samples(gmb_model* mdl, int m) {
    gmb_ctx* ctx = mdl->ctx;
    if (m <= mdl->m_cap) return GMB_OK;
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (mdl->dU) { gmb_dfree(ctx, mdl->dU); mdl->dU = nullptr; }
    if (mdl->dzd) { gmb_dfree(ctx, mdl->dzd); mdl->dzd = nullptr; }
    if (mdl->dF) { gmb_dfree(ctx, mdl->dF); mdl->dF = nullptr; }
    mdl->f_valid = false;
    mdl->m_cap = 0;
    size_t cap = (size_t)m;
    GMB_CUDA(gmb_dmalloc(ctx, &mdl->dU, sizeof(double) * mdl->ldq * cap));
    if (mdl->dzd32) { gmb_dfree(ctx, mdl->dzd32); mdl->dzd32 = nullptr; }
    if (mdl->dF32) { gmb_dfree(ctx, mdl->dF32); mdl->dF32 = nullptr; }
    // padding rows must hold finite values: the streaming kernels load them (and mask the result)
    GMB_CUDA(cudaMemsetAsync(mdl->dU, 0, sizeof(double) * mdl->ldq * cap, ctx->stream));
    if (mdl->prec == 32) {
        GMB_CUDA(gmb_dmalloc(ctx, &mdl->dzd32, sizeof(float) * mdl->ldn * cap));
        if (mdl->flink == 3) GMB_CUDA(gmb_dmalloc(ctx, &mdl->dF32, sizeof(float) * mdl->ldn * cap));
        GMB_CUDA(cudaMemsetAsync(mdl->dzd32, 0, sizeof(float) * mdl->ldn * cap, ctx->stream));
    } else {
        GMB_CUDA(gmb_dmalloc(ctx, &mdl->dzd, sizeof(double) * mdl->ldn * cap));
        if (mdl->flink == 3) GMB_CUDA(gmb_dmalloc(ctx, &mdl->dF, sizeof(double) * mdl->ldn * cap));
        GMB_CUDA(cudaMemsetAsync(mdl->dzd, 0, sizeof(double) * mdl->ldn * cap, ctx->stream));
    }
    mdl->m_cap = m;
    return GMB_OK;
}

// zd = Z u for the first m_local columns of dU (mcmlmodel.h:286 / :117, hoisted)
// zd = Z u with Z in ELL form (indicator designs: one or two non-zeros per row): a gather instead of a 2 n Q m flop contraction
template <class T>
__global__ void __launch_bounds__(256) zd_sparse_kernel(int n, int ldn, int ngp, int wr, int ldq, const double* __restrict__ rv, const int* __restrict__ rc,
                                                        const int* __restrict__ rowmap /* row of Z that output row i reads; NULL: i */,
                                                        const double* __restrict__ U, T* __restrict__ zd) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const size_t j = blockIdx.y;
    if (i >= ldn) return;
    double s = 0.0;
    if (i < n) {
        const int zi = rowmap ? rowmap[i] : i;
        for (int w = 0; w < wr; w++) s = fma(rv[(size_t)w * ngp + zi], U[rc[(size_t)w * ngp + zi] + j * ldq], s);
    }
    zd[i + j * ldn] = (T)s;
}

// fp32 mode, dense Z: a block of columns of the fp64 product, narrowed
__global__ void narrow_kernel(size_t count, const double* __restrict__ src, float* __restrict__ dst) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < count) dst[e] = (float)src[e];
}

// 1 (default) = build zd by gathering when Z is sparse (same criterion as the samplers' sparse forms), 0 = always the dense contraction
static int g_zd_sparse = 1;
extern "C" int gmb_estep_set_sparse_zd(int on) { g_zd_sparse = on ? 1 : 0; return GMB_OK; }

// 1 (default) = the E-step runs on the distinct rows of [X | Z] when at most a quarter of the rows are distinct (poisson/log, binomial/logit
// with 0/1 responses, gaussian/identity); 0 = one row of zd per observation
static int g_estep_agg = 1;
extern "C" int gmb_estep_set_row_aggregation(int on) { g_estep_agg = on ? 1 : 0; return GMB_OK; }
extern "C" int gmb_model_estep_rows(gmb_model* mdl, int* rows) {
    if (!mdl || !rows) return gmb_set_error(GMB_EINVAL, "gmb_model_estep_rows: bad arguments");
    *rows = (mdl->zd_valid && mdl->eagg) ? mdl->agg.ng : mdl->n;
    return GMB_OK;
}

int gmb_model_build_zd(gmb_model* mdl) {
    bool sparse = false;
    if (mdl->m_local > 0 && g_zd_sparse && mdl->Q >= 64) {
        GMB_TRY(gmb_zell_ensure(mdl));
        sparse = mdl->zell.valid && mdl->m_local <= 65535 * 16;
    }
    // rows that share [X | Z] share zd — form it (and run the E-step kernels, estep.cu) on the distinct rows only
    mdl->eagg = false;
    if (g_estep_agg && mdl->m_local > 0 && gmb_flink_core(mdl->flink) && gmb_agg_enabled()) {
        GMB_TRY(gmb_agg_ensure(mdl));
        const gmb_agg& a = mdl->agg;
        mdl->eagg = a.active && (long long)a.ng * 4 <= mdl->n && (mdl->flink != 3 || a.binary_ok);
    }
    const int ne = mdl->eagg ? mdl->agg.ng : mdl->n, lde = mdl->eagg ? mdl->agg.ldn : mdl->ldn;
    const double* Ze = mdl->eagg ? mdl->agg.dZ : mdl->dZ;
    const int* rowmap = mdl->eagg ? mdl->agg.drep : nullptr;
    if (sparse) {
        const gmb_ell& e = mdl->zell;
        for (int j0 = 0; j0 < mdl->m_local; j0 += 65535) {
            const int nc = std::min(65535, mdl->m_local - j0);
            if (mdl->prec == 32)
                zd_sparse_kernel<float><<<dim3((lde + 255) / 256, nc), 256, 0, mdl->ctx->stream>>>(ne, lde, e.ngp, e.wr, mdl->ldq, e.rv, e.rc, rowmap,
                                                                                                 mdl->dU + (size_t)j0 * mdl->ldq, mdl->dzd32 + (size_t)j0 * lde);
            else
                zd_sparse_kernel<double><<<dim3((lde + 255) / 256, nc), 256, 0, mdl->ctx->stream>>>(ne, lde, e.ngp, e.wr, mdl->ldq, e.rv, e.rc, rowmap,
                                                                                                  mdl->dU + (size_t)j0 * mdl->ldq, mdl->dzd + (size_t)j0 * lde);
            mdl->ctx->launches++;
        }
        GMB_CUDA(cudaGetLastError());
    } else if (mdl->m_local > 0 && mdl->prec == 32 && gmb_tf32_enabled() && !mdl->eagg) {
        // dense Z in fp32 mode: tensor cores, tcgen05.mma kind::tf32 with the 3xTF32 split (gemm_tf32.cu)
        gmb_ctx* ctx = mdl->ctx;
        const int ldk = round_up(mdl->Q, 4);
        if (!mdl->dZt_hi) {
            GMB_CUDA(gmb_dmalloc(ctx, &mdl->dZt_hi, sizeof(float) * (size_t)mdl->n * ldk));
            GMB_CUDA(gmb_dmalloc(ctx, &mdl->dZt_lo, sizeof(float) * (size_t)mdl->n * ldk));
            GMB_TRY(gmb_split_tf32(ctx, mdl->n, mdl->Q, mdl->ldn, mdl->dZ, 1, ldk, mdl->dZt_hi, mdl->dZt_lo));
        }
        const size_t need = (size_t)mdl->m_local * ldk;
        if (need > mdl->u32_cap) {
            if (mdl->dU_hi) { GMB_CUDA(cudaStreamSynchronize(ctx->stream)); gmb_dfree(ctx, mdl->dU_hi); gmb_dfree(ctx, mdl->dU_lo); mdl->dU_hi = mdl->dU_lo = nullptr; }
            GMB_CUDA(gmb_dmalloc(ctx, &mdl->dU_hi, sizeof(float) * need));
            GMB_CUDA(gmb_dmalloc(ctx, &mdl->dU_lo, sizeof(float) * need));
            mdl->u32_cap = need;
        }
        GMB_TRY(gmb_split_tf32(ctx, mdl->Q, mdl->m_local, mdl->ldq, mdl->dU, 0, ldk, mdl->dU_hi, mdl->dU_lo));
        GMB_TRY(gmb_sgemm3_tf32(ctx, mdl->n, mdl->m_local, mdl->Q, mdl->dZt_hi, mdl->dZt_lo, ldk, mdl->dU_hi, mdl->dU_lo, ldk, mdl->dzd32, mdl->ldn));
    } else if (mdl->m_local > 0 && mdl->prec == 32) {
        // dense Z in fp32 mode: the fp64 DMMA product in column chunks through the context's scratch area, narrowed to float
        gmb_ctx* ctx = mdl->ctx;
        const int chunk = std::max(1, std::min(mdl->m_local, (int)(((size_t)1 << 27) / (size_t)lde)));
        GMB_TRY(gmb_ctx_scratch(ctx, (size_t)lde * chunk));
        for (int j0 = 0; j0 < mdl->m_local; j0 += chunk) {
            const int nc = std::min(chunk, mdl->m_local - j0);
            GMB_CUDA(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(double) * (size_t)lde * nc, ctx->stream));
            GMB_TRY(gmb_dgemm(ctx, 0, 0, ne, nc, mdl->Q, 1.0, Ze, lde, mdl->dU + (size_t)j0 * mdl->ldq, mdl->ldq, 0.0, ctx->d_scratch, lde));
            const size_t cnt = (size_t)lde * nc;
            narrow_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(cnt, ctx->d_scratch, mdl->dzd32 + (size_t)j0 * lde);
            ctx->launches++;
        }
        GMB_CUDA(cudaGetLastError());
    } else if (mdl->m_local > 0) {
        if (mdl->eagg) GMB_CUDA(cudaMemsetAsync(mdl->dzd, 0, sizeof(double) * (size_t)lde * mdl->m_local, mdl->ctx->stream));    // padding rows
        GMB_TRY(gmb_dgemm(mdl->ctx, 0, 0, ne, mdl->m_local, mdl->Q, 1.0, Ze, lde, mdl->dU, mdl->ldq, 0.0, mdl->dzd, lde));
    }
    mdl->f_valid = false;
    mdl->stat_valid = false;
    if (mdl->flink == 3 && (mdl->dF || mdl->dF32) && mdl->m_local > 0) {        // factor matrix of the binomial/logit E-step (estep.cu); aggregated rows: exp(zd)
        GMB_TRY(gmb_launch_build_factor(mdl, mdl->m_local));
        mdl->f_valid = true;
    }
    mdl->zd_valid = true;
    return GMB_OK;
}

static int set_counts(gmb_model* mdl, int m_local, int m_total, int niter_total) {
    gmb_ctx* ctx = mdl->ctx;
    if (m_total <= 0) m_total = m_local;
    if (niter_total <= 0 || niter_total > m_total) niter_total = m_total;
    if (ctx->world == 1 && m_total != m_local) return gmb_set_error(GMB_EINVAL, "m_total (%d) != m_local (%d) on a single rank", m_total, m_local);
    // the E-step uses the leading niter_total columns (mcmlmodel.h:73,295); ranks own consecutive column ranges,
    // so only the trailing rank(s) drop columns
    int drop = m_total - niter_total;
    int nl = m_local;
    if (drop > 0) {
        if (ctx->world == 1) nl = m_local - drop;
        else if (ctx->rank == ctx->world - 1) {
            if (drop > m_local) return gmb_set_error(GMB_EINVAL, "niter_total leaves the last rank with a negative column count");
            nl = m_local - drop;
        }
    }
    mdl->m_local = m_local; mdl->m_total = m_total; mdl->niter_total = niter_total; mdl->niter_local = nl;
    return GMB_OK;
}

extern "C" int gmb_model_set_u(gmb_model* mdl, const double* U, int Q, int m_local, int m_total, int niter_total) {
    if (!mdl || m_local < 0 || (m_local > 0 && !U)) return gmb_set_error(GMB_EINVAL, "gmb_model_set_u: bad arguments");
    if (Q != mdl->Q) return gmb_set_error(GMB_EINVAL, "u has %d rows, Z has %d columns", Q, mdl->Q);
    gmb_ctx* ctx = mdl->ctx;
    GMB_CUDA(cudaSetDevice(ctx->device));
    mdl->zd_valid = false;
    GMB_TRY(set_counts(mdl, m_local, m_total, niter_total));
    GMB_TRY(gmb_model_reserve_samples(mdl, m_local > 0 ? m_local : 1));
    GMB_TRY(upload_matrix(ctx, mdl->dU, mdl->ldq, U, Q, m_local));
    mdl->u_version++;
    GMB_TRY(gmb_model_build_zd(mdl));
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));   // U is a caller buffer: do not return before it has been read
    return GMB_OK;
}

extern "C" int gmb_model_use_device_u(gmb_model* mdl, int niter_total) {
    if (!mdl) return gmb_set_error(GMB_EINVAL, "model is NULL");
    if (!mdl->dU || mdl->m_local <= 0 && mdl->ctx->world == 1) return gmb_set_error(GMB_ESTATE, "the model holds no device samples (run gmb_hmc_sample with keep_on_device)");
    GMB_CUDA(cudaSetDevice(mdl->ctx->device));
    GMB_TRY(set_counts(mdl, mdl->m_local, mdl->m_total, niter_total));
    if (!mdl->zd_valid) GMB_TRY(gmb_model_build_zd(mdl));
    return GMB_OK;
}

// Re-forms zd = Z u (and the factor matrix) from the device-resident samples: what gmb_model_set_u does after its upload — exposed so that the
// contraction can be timed without the host-to-device copy (benchmarks, profiles).
extern "C" int gmb_model_rebuild_zd(gmb_model* mdl) {
    if (!mdl || !mdl->dU || mdl->m_local <= 0) return gmb_set_error(GMB_ESTATE, "gmb_model_rebuild_zd: the model holds no samples");
    GMB_CUDA(cudaSetDevice(mdl->ctx->device));
    return gmb_model_build_zd(mdl);
}

extern "C" int gmb_model_get_u(gmb_model* mdl, int col0, int ncols, double* U_out) {
    if (!mdl || !U_out || col0 < 0 || ncols < 0) return gmb_set_error(GMB_EINVAL, "gmb_model_get_u: bad arguments");
    if (!mdl->dU || col0 + ncols > mdl->m_local) return gmb_set_error(GMB_ESTATE, "the model holds %d sample columns, asked for [%d, %d)", mdl->dU ? mdl->m_local : 0, col0, col0 + ncols);
    if (ncols == 0) return GMB_OK;
    GMB_CUDA(cudaSetDevice(mdl->ctx->device));
    GMB_CUDA(cudaMemcpy2DAsync(U_out, (size_t)mdl->Q * sizeof(double), mdl->dU + (size_t)col0 * mdl->ldq, (size_t)mdl->ldq * sizeof(double),
                               (size_t)mdl->Q * sizeof(double), ncols, cudaMemcpyDeviceToHost, mdl->ctx->stream));
    GMB_CUDA(cudaStreamSynchronize(mdl->ctx->stream));
    return GMB_OK;
}

static int check_ready(gmb_model* mdl, const double* beta) {
    if (!mdl || !beta) return gmb_set_error(GMB_EINVAL, "model or beta is NULL");
    if (!mdl->zd_valid) return gmb_set_error(GMB_ESTATE, "no samples set: call gmb_model_set_u (or gmb_hmc_sample + gmb_model_use_device_u) first");
    if (mdl->niter_total <= 0) return gmb_set_error(GMB_ESTATE, "the sample matrix has no columns");
    return GMB_OK;
}

// 1 (default): batched binomial/logit evaluations share their pass over the factor matrix (estep.cu: loglik_logit_factor_multi_kernel)
static int g_loglik_multi = 1;
extern "C" int gmb_estep_set_multi(int on) { g_loglik_multi = on ? 1 : 0; return GMB_OK; }

// one chunk of a batch: at most GMB_RESULT_DOUBLES evaluations whose parameters fit the pinned staging area
static int loglik_chunk(gmb_model* mdl, const double* beta_mat, const double* var_par, int n_eval, double* out) {
    gmb_ctx* ctx = mdl->ctx;
    GMB_TRY(upload_params(mdl, beta_mat, mdl->P * n_eval));
    int e0 = 0;
    if (n_eval >= GMB_LOGLIK_NB && g_loglik_multi) {
        // binomial/logit on the factor matrix: the whole batch in one launch, GMB_LOGLIK_NB evaluations per pass over F
        int done = 0;
        GMB_TRY(gmb_launch_loglik_multi(mdl, mdl->dbeta, n_eval, ctx->d_result, &done));
        if (done) e0 = n_eval;
    }
    for (int e = e0; e < n_eval; e++)
        GMB_TRY(gmb_launch_loglik(mdl, mdl->dbeta + (size_t)e * mdl->P, var_par[e], ctx->d_result + e));
    GMB_TRY(gmb_comm_allreduce_dev(ctx, ctx->d_result, n_eval));
    double* hres = ctx->h_pinned + ctx->pinned_doubles / 2;
    GMB_CUDA(cudaMemcpyAsync(hres, ctx->d_result, sizeof(double) * n_eval, cudaMemcpyDeviceToHost, ctx->stream));
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int e = 0; e < n_eval; e++) out[e] = hres[e] / mdl->niter_total;   // mcmlmodel.h:303 ll.mean()
    return GMB_OK;
}

extern "C" int gmb_model_loglik_batch(gmb_model* mdl, const double* beta_mat, const double* var_par, int n_eval, double* out) {
    GMB_TRY(check_ready(mdl, beta_mat));
    if (!var_par || !out || n_eval <= 0) return gmb_set_error(GMB_EINVAL, "gmb_model_loglik_batch: bad arguments");
    gmb_ctx* ctx = mdl->ctx;
    GMB_CUDA(cudaSetDevice(ctx->device));
    if (mdl->flink == 7 || mdl->flink == 8) for (int e = 0; e < n_eval; e++) if (!(var_par[e] > 0.0)) return gmb_set_error(GMB_EINVAL, "gaussian var_par must be > 0 (got %g)", var_par[e]);
    // any batch size: chunks bounded by the result buffer and by the pinned staging area (P values per evaluation; the reference
    // has no limit on P or on the number of optimhess points, mcmloptim.h:333-355)
    size_t cmax = (ctx->pinned_doubles / 2) / (size_t)mdl->P;
    if (cmax > GMB_RESULT_DOUBLES) cmax = GMB_RESULT_DOUBLES;
    if (cmax < 1) return gmb_set_error(GMB_EINVAL, "P = %d exceeds the parameter staging area", mdl->P);
    for (int off = 0; off < n_eval; off += (int)cmax) {
        const int nb = n_eval - off < (int)cmax ? n_eval - off : (int)cmax;
        GMB_TRY(loglik_chunk(mdl, beta_mat + (size_t)off * mdl->P, var_par + off, nb, out + off));
    }
    return GMB_OK;
}

extern "C" int gmb_model_loglik(gmb_model* mdl, const double* beta, double var_par, double* out) {
    return gmb_model_loglik_batch(mdl, beta, &var_par, 1, out);
}

// Solves A x = b for a small P x P system by Gaussian elimination with partial pivoting
// (stands in for Eigen's .inverse() at mcmloptim.h:230).  Returns non-zero when singular.
int gmb_solve_small(int P, const double* A_in, const double* b_in, double* x) {
    std::vector<double> A(A_in, A_in + (size_t)P * P), b(b_in, b_in + P);
    for (int c = 0; c < P; c++) {
        int piv = c; double best = fabs(A[c + (size_t)c * P]);
        for (int r = c + 1; r < P; r++) if (fabs(A[r + (size_t)c * P]) > best) { best = fabs(A[r + (size_t)c * P]); piv = r; }
        if (!(best > 0.0)) return 1;
        if (piv != c) { for (int k = 0; k < P; k++) std::swap(A[c + (size_t)k * P], A[piv + (size_t)k * P]); std::swap(b[c], b[piv]); }
        for (int r = c + 1; r < P; r++) {
            double f = A[r + (size_t)c * P] / A[c + (size_t)c * P];
            for (int k = c; k < P; k++) A[r + (size_t)k * P] -= f * A[c + (size_t)k * P];
            b[r] -= f * b[c];
        }
    }
    for (int r = P - 1; r >= 0; r--) {
        double s = b[r];
        for (int k = r + 1; k < P; k++) s -= A[r + (size_t)k * P] * x[k];
        x[r] = s / A[r + (size_t)r * P];
    }
    return 0;
}

extern "C" int gmb_model_mcnr(gmb_model* mdl, const double* beta, double var_par,
                              double* xtwx, double* score, double* beta_incr, double* sigma) {
    GMB_TRY(check_ready(mdl, beta));
    gmb_ctx* ctx = mdl->ctx;
    GMB_CUDA(cudaSetDevice(ctx->device));
    const int P = mdl->P, nout = P * P + P + 1;
    if (nout > GMB_RESULT_DOUBLES) return gmb_set_error(GMB_EINVAL, "P = %d is too large for the MCNR sums buffer", P);
    if (mdl->n < 2) return gmb_set_error(GMB_EINVAL, "MCNR needs n >= 2 (sd of the residuals, mcmloptim.h:216)");
    GMB_TRY(upload_params(mdl, beta, P));
    GMB_TRY(gmb_launch_xb(mdl, mdl->dbeta, mdl->dxb));
    GMB_TRY(gmb_launch_mcnr(mdl, mdl->dxb, var_par, ctx->d_result));
    GMB_TRY(gmb_comm_allreduce_dev(ctx, ctx->d_result, nout));      // the P^2 + P + 1 sufficient sums, SURVEY §8e
    double* hres = ctx->h_pinned + ctx->pinned_doubles / 2;
    GMB_CUDA(cudaMemcpyAsync(hres, ctx->d_result, sizeof(double) * nout, cudaMemcpyDeviceToHost, ctx->stream));
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    const double inv = 1.0 / mdl->niter_total;
    std::vector<double> A((size_t)P * P), sc(P), incr(P);
    for (int k = 0; k < P * P; k++) A[k] = hres[k] * inv;          // mcmloptim.h:227-229
    for (int k = 0; k < P; k++) sc[k] = hres[P * P + k] * inv;     // :231-232
    if (xtwx) memcpy(xtwx, A.data(), sizeof(double) * P * P);
    if (score) memcpy(score, sc.data(), sizeof(double) * P);
    if (sigma) *sigma = hres[P * P + P] * inv;                     // :235
    if (beta_incr) {
        if (gmb_solve_small(P, A.data(), sc.data(), incr.data())) return gmb_set_error(GMB_EINVAL, "MCNR: X^T W X is singular");
        memcpy(beta_incr, incr.data(), sizeof(double) * P);
    }
    return GMB_OK;
}
