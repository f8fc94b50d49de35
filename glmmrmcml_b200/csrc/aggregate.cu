// aggregate.cu — row aggregation for the on-chip sampler ("structure-aware Z").
//
// In a cluster design many observations share their row of [X | Z] (config C2: 10 individuals per cluster-period, 500 rows but 50
// distinct ones), hence their linear predictor eta_i = x_i' beta + (Z L v)_i for EVERY beta and v.  The sampler's log-density and
// gradient only need, per distinct row g, the number of observations c_g and sums of their responses:
//   binomial/logit : sum_i l_i = ys_g log p + (c_g - ys_g) log(1 - p),     sum_i r_i = c_g / (1 + e^eta) + (ys_g - c_g)
//   poisson/log    : sum_i l_i = ys_g eta - c_g e^eta - sum_i lf(y_i),     sum_i r_i = ys_g - c_g e^eta
//   gaussian/id    : sum_i (y_i - eta)^2 = ss_g + c_g (ybar_g - eta)^2,    sum_i r_i = ys_g - c_g eta
// (the same sums as mcmlmodel.h:138-279 in a different order).  The sampler then runs on the n_g distinct rows: for C2 a tenth of the
// tensor work per leapfrog step, and a model ten times larger fits the on-chip variant.
//
// Grouping: a 64-bit hash of every row of [X | Z] is computed on the device, rows are grouped by hash on the host, and a second kernel
// verifies every row against its group's representative element by element — on any mismatch (a hash collision) aggregation is
// switched off and the sampler runs on the original rows.  Built lazily at the first sampling call of a model, cached.
#include "common.cuh"
#include <algorithm>
#include <numeric>

namespace {

__device__ __forceinline__ unsigned long long mix64(unsigned long long h, unsigned long long v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 29;
    return h;
}

__global__ void row_hash_kernel(int n, int P, int Q, int ldn, const double* __restrict__ X, const double* __restrict__ Z,
                                unsigned long long* __restrict__ hash) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long h = 0x243F6A8885A308D3ull;
    for (int p = 0; p < P; p++) { const double v = X[i + (size_t)p * ldn]; h = mix64(h, (unsigned long long)__double_as_longlong(v == 0.0 ? 0.0 : v)); }
    for (int q = 0; q < Q; q++) {
        const double v = Z[i + (size_t)q * ldn];
        if (v != 0.0) h = mix64(mix64(h, (unsigned long long)q), (unsigned long long)__double_as_longlong(v));
    }
    hash[i] = h;
}

// every row equals the representative row of its group?  (exact comparison; -0.0 == 0.0)
__global__ void row_verify_kernel(int n, int P, int Q, int ldn, const double* __restrict__ X, const double* __restrict__ Z,
                                  const int* __restrict__ gid, const int* __restrict__ rep, int* __restrict__ mismatch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = rep[gid[i]];
    if (r == i) return;
    bool bad = false;
    for (int p = 0; p < P; p++) bad |= X[i + (size_t)p * ldn] != X[r + (size_t)p * ldn];
    for (int q = 0; q < Q; q++) bad |= Z[i + (size_t)q * ldn] != Z[r + (size_t)q * ldn];
    if (bad) atomicExch(mismatch, 1);
}

__global__ void gather_rows_kernel(int ng, int ldng, int ncol, int ldn, const double* __restrict__ A, const int* __restrict__ rep,
                                   double* __restrict__ Ag) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (g < ldng && c < ncol) Ag[g + (size_t)c * ldng] = (g < ng) ? A[rep[g] + (size_t)c * ldn] : 0.0;
}

}  // namespace

void gmb_agg_free(gmb_model* mdl) {
    gmb_agg& a = mdl->agg;
    gmb_ctx* ctx = mdl->ctx;
    if (a.built && a.active) { gmb_dfree(ctx, a.dX); gmb_dfree(ctx, a.dZ); gmb_dfree(ctx, a.dZL); gmb_dfree(ctx, a.dxb); }
    gmb_dfree(ctx, a.dvec); gmb_dfree(ctx, a.dgid); gmb_dfree(ctx, a.drep);
    mdl->eagg = false;
    a = gmb_agg();
}

// Builds mdl->agg (once per model).  When fewer than 3/4 of the rows are distinct the sampler's view is the aggregated model,
// otherwise the view aliases the model's own arrays with unit weights.
int gmb_agg_ensure(gmb_model* mdl) {
    gmb_agg& a = mdl->agg;
    if (a.built && a.flag == gmb_agg_enabled()) return GMB_OK;
    const bool rebuild_zd = a.built && mdl->eagg && mdl->zd_valid;   // the E-step's zd lives on the view that is about to change
    if (a.built) gmb_agg_free(mdl);
    gmb_sparse_invalidate(mdl);                      // the sparse forms describe the view
    a.flag = gmb_agg_enabled();
    gmb_ctx* ctx = mdl->ctx;
    const int n = mdl->n, P = mdl->P, Q = mdl->Q, ldn = mdl->ldn, fl = mdl->flink;
    std::vector<double> y(n), rowc(n);
    GMB_CUDA(cudaMemcpyAsync(y.data(), mdl->dy, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    GMB_CUDA(cudaMemcpyAsync(rowc.data(), mdl->drowc, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<int> gid(n), rep;
    bool want = gmb_agg_enabled() && n >= 16;
    if (want) {
        unsigned long long* d_hash = nullptr;
        GMB_CUDA(gmb_dmalloc(ctx, &d_hash, sizeof(unsigned long long) * n));
        row_hash_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, P, Q, ldn, mdl->dX, mdl->dZ, d_hash);
        ctx->launches++;
        std::vector<unsigned long long> h(n);
        GMB_CUDA(cudaMemcpyAsync(h.data(), d_hash, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost, ctx->stream));
        GMB_CUDA(cudaStreamSynchronize(ctx->stream));
        gmb_dfree(ctx, d_hash);
        // groups in order of first appearance (deterministic)
        std::vector<int> order(n);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int p, int q) { return h[p] < h[q]; });
        std::vector<int> first(n, -1);                 // representative (smallest row index) of the hash class of each row
        for (int k = 0; k < n;) {
            int e = k; int r0 = order[k];
            while (e < n && h[order[e]] == h[order[k]]) { r0 = std::min(r0, order[e]); e++; }
            for (int t = k; t < e; t++) first[order[t]] = r0;
            k = e;
        }
        std::vector<int> gindex(n, -1);
        for (int i = 0; i < n; i++) {
            if (first[i] == i) { gindex[i] = (int)rep.size(); rep.push_back(i); }
            gid[i] = gindex[first[i]];
        }
        want = (long long)rep.size() * 4 <= (long long)n * 3;
        if (want) {                                    // exact verification on the device
            int *d_gid = nullptr, *d_rep = nullptr, *d_bad = nullptr;
            GMB_CUDA(gmb_dmalloc(ctx, &d_gid, sizeof(int) * n));
            GMB_CUDA(gmb_dmalloc(ctx, &d_rep, sizeof(int) * rep.size()));
            GMB_CUDA(gmb_dmalloc(ctx, &d_bad, sizeof(int)));
            GMB_CUDA(cudaMemcpyAsync(d_gid, gid.data(), sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
            GMB_CUDA(cudaMemcpyAsync(d_rep, rep.data(), sizeof(int) * rep.size(), cudaMemcpyHostToDevice, ctx->stream));
            GMB_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
            row_verify_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, P, Q, ldn, mdl->dX, mdl->dZ, d_gid, d_rep, d_bad);
            ctx->launches++;
            int bad = 0;
            GMB_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            GMB_CUDA(cudaStreamSynchronize(ctx->stream));
            if (bad) want = false;
            else {
                a.ng = (int)rep.size(); a.ldn = round_up(a.ng, 4);
                GMB_CUDA(gmb_dmalloc(ctx, &a.dX, sizeof(double) * (size_t)a.ldn * P));
                GMB_CUDA(gmb_dmalloc(ctx, &a.dZ, sizeof(double) * (size_t)a.ldn * Q));
                GMB_CUDA(gmb_dmalloc(ctx, &a.dZL, sizeof(double) * (size_t)a.ldn * Q));
                GMB_CUDA(gmb_dmalloc(ctx, &a.dxb, sizeof(double) * (size_t)a.ldn));
                GMB_CUDA(cudaMemsetAsync(a.dZL, 0, sizeof(double) * (size_t)a.ldn * Q, ctx->stream));
                GMB_CUDA(cudaMemsetAsync(a.dxb, 0, sizeof(double) * (size_t)a.ldn, ctx->stream));
                gather_rows_kernel<<<dim3((a.ldn + 127) / 128, P), 128, 0, ctx->stream>>>(a.ng, a.ldn, P, ldn, mdl->dX, d_rep, a.dX);
                gather_rows_kernel<<<dim3((a.ldn + 127) / 128, Q), 128, 0, ctx->stream>>>(a.ng, a.ldn, Q, ldn, mdl->dZ, d_rep, a.dZ);
                ctx->launches += 2;
                GMB_CUDA(cudaStreamSynchronize(ctx->stream));
            }
            if (want) { a.dgid = d_gid; a.drep = d_rep; } else { gmb_dfree(ctx, d_gid); gmb_dfree(ctx, d_rep); }
            gmb_dfree(ctx, d_bad);
        }
    } else {
        GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (!want) {                                       // identity view
        a.ng = n; a.ldn = ldn;
        for (int i = 0; i < n; i++) gid[i] = i;
    }
    a.active = want;
    // per-row weights and response sums, accumulated in row order on the host (deterministic)
    const int ng = a.ng, ldg = a.ldn;
    std::vector<double> hv((size_t)8 * ldg, 0.0);
    double *cnt = hv.data(), *ys = cnt + ldg, *lcnt = ys + ldg, *lys = lcnt + ldg, *lsq = lys + ldg, *lrc = lsq + ldg, *eys = lrc + ldg, *ess = eys + ldg;
    std::vector<double> ysum(ng, 0.0);
    for (int i = 0; i < n; i++) { cnt[gid[i]] += 1.0; ysum[gid[i]] += y[i]; lrc[gid[i]] += rowc[i]; }
    for (int i = 0; i < n; i++) {                      // log-likelihood weights
        const int g = gid[i];
        if (fl == 3) { if (y[i] == 1.0 || y[i] == 0.0) { lcnt[g] += 1.0; lys[g] += y[i]; } }      // other y contribute nothing (moremaths.h:47-53)
        else lcnt[g] += 1.0;
    }
    for (int g = 0; g < ng; g++) {
        if (fl == 3) ys[g] = ysum[g] - cnt[g];                                                   // r = c / (1 + e^eta) + (ys - c)
        else ys[g] = ysum[g];
        if (fl == 1) lys[g] = ysum[g];
        if (fl == 7) lys[g] = ysum[g] / cnt[g];                                                   // group mean
    }
    if (fl == 7) for (int i = 0; i < n; i++) { const double d = y[i] - lys[gid[i]]; lsq[gid[i]] += d * d; }   // within-group sum of squares
    a.binary_ok = true;
    for (int g = 0; g < ng; g++) eys[g] = ysum[g];
    for (int i = 0; i < n; i++) {
        const int g = gid[i];
        const double d = y[i] - ysum[g] / cnt[g];
        ess[g] += d * d;
        if (fl == 3 && !(y[i] == 0.0 || y[i] == 1.0)) a.binary_ok = false;
    }
    GMB_CUDA(gmb_dmalloc(ctx, &a.dvec, sizeof(double) * hv.size()));
    GMB_CUDA(cudaMemcpyAsync(a.dvec, hv.data(), sizeof(double) * hv.size(), cudaMemcpyHostToDevice, ctx->stream));
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    a.dcnt = a.dvec; a.dys = a.dvec + ldg; a.dlcnt = a.dvec + 2 * (size_t)ldg; a.dlys = a.dvec + 3 * (size_t)ldg;
    a.dlsq = a.dvec + 4 * (size_t)ldg; a.dlrc = a.dvec + 5 * (size_t)ldg; a.deys = a.dvec + 6 * (size_t)ldg; a.dess = a.dvec + 7 * (size_t)ldg;
    if (!a.active) { a.dX = mdl->dX; a.dZ = mdl->dZ; a.dZL = nullptr; a.dxb = nullptr; }          // aliases: resolved by the accessors below
    a.built = true;
    if (rebuild_zd) GMB_TRY(gmb_model_build_zd(mdl));
    return GMB_OK;
}

// 1 = aggregate duplicate rows for the on-chip sampler (default), 0 = never
static int g_agg = 1;
int gmb_agg_enabled() { return g_agg; }
extern "C" int gmb_hmc_set_row_aggregation(int on) { g_agg = on ? 1 : 0; return GMB_OK; }
