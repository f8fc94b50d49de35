// laplace.cu — Laplace-approximation fits mcml_la / mcml_la_nr (src/mcml_la.cpp:28-155, 178-290) on the device primitives of
// the Monte-Carlo path: D(theta) build + Cholesky (cov.cu), Z L and (Z L)^T W (Z L) as DMMA GEMMs (gemm_f64.cu), the family
// log-likelihood stream (estep.cu) on single columns, the blocked Cholesky (cov_large.cu) for log det((Z L)^T W (Z L) + I).
// m = 1 here, so nothing is a throughput target: the objectives run one evaluation per launch
// sequence and the O(n P^2), O(Q^3) pieces of the Newton step run on the host.
//
//   LA_likelihood         (likelihood.h:112-141)  -> LaFit::obj_bv       over (beta, v)
//   LA_likelihood_cov     (likelihood.h:143-181)  -> LaFit::obj_cov      over theta (, sigma)
//   LA_likelihood_btheta  (likelihood.h:183-230)  -> LaFit::obj_btheta   over (beta, theta (, sigma))
//   la_optim / la_optim_cov / la_optim_bcov / hess_la / mcnr_b (mcmloptim.h:116-195, 238-293) -> methods of the same names
// The optimised state is the whitened v (u = L v only for the return value); update_W() without arguments forms the weights
// at xb + Z v and mcnr_b mixes Z v and Z L v exactly as the reference does (oracle/laplace.py lists the quirks).
#include "common.cuh"
#include <limits>
#include <algorithm>

int gmb_default_ctx(gmb_ctx** out);   // api.cpp

namespace {

const double kInf = std::numeric_limits<double>::infinity();

// W_ii = 1 / (dhdmu(xb_i + zu_i) * nvar), mcmlmodel.h:120-134 with glmmrBase's dhdmu (reconstructed)
__global__ void la_w_kernel(int n, int flink, const double* __restrict__ xb, const double* __restrict__ zu, double nvar,
                            double* __restrict__ W) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double eta = xb[i] + zu[i];
    double w = 1.0;
    if (flink == 1) w = exp(-eta);
    else if (flink == 3) { const double p = exp(eta) / (1 + exp(eta)); w = 1 / (p * (1 - p)); }
    W[i] = 1 / (w * nvar);
}

// B = diag(W) ZL
__global__ void la_scale_rows_kernel(int n, int Q, int ldn, const double* __restrict__ W, const double* __restrict__ ZL,
                                     double* __restrict__ B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int q = blockIdx.y;
    if (i < ldn && q < Q) B[i + (size_t)q * ldn] = (i < n) ? W[i] * ZL[i + (size_t)q * ldn] : 0.0;
}

__global__ void la_identity_kernel(int Q, int ld, double* __restrict__ M) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y;
    if (i < ld && j < Q) M[i + (size_t)j * ld] = (i == j) ? 1.0 : 0.0;
}

struct LaFit {
    gmb_ctx* ctx = nullptr; gmb_model* M = nullptr; gmb_cov* D = nullptr;
    int n = 0, P = 0, Q = 0, R = 0, ldn = 0, ldq = 0;
    bool gaussian = false;
    const double* hX = nullptr; const double* hy = nullptr;   // the caller's buffers (valid for the duration of the call)
    std::vector<double> beta, theta, v, Lhost, D0;             // D0 = L L' of the INITIAL theta (mcmlmodel.h:66; never refreshed)
    double var_par = 1.0, sigma = 1.0;
    std::vector<double> theta_L;                               // theta at which dL / dZL currently stand
    // device buffers
    double *d_par = nullptr, *d_V = nullptr, *d_W = nullptr, *d_zu = nullptr, *d_B = nullptr, *d_M = nullptr, *d_linv = nullptr, *d_vec = nullptr;
    int* d_status = nullptr;
    int kmax = 0;

    ~LaFit() {
        if (!ctx) return;
        gmb_dfree(ctx, d_par); gmb_dfree(ctx, d_V); gmb_dfree(ctx, d_W); gmb_dfree(ctx, d_zu); gmb_dfree(ctx, d_B); gmb_dfree(ctx, d_M);
        gmb_dfree(ctx, d_linv); gmb_dfree(ctx, d_vec); gmb_dfree(ctx, d_status);
    }

    int init(gmb_ctx* c, gmb_cov* d, gmb_model* m, const double* X, const double* y, const double* start, int n_start, const char* family) {
        ctx = c; D = d; M = m; n = m->n; P = m->P; Q = m->Q; ldn = m->ldn; ldq = m->ldq; hX = X; hy = y;
        int B;
        GMB_TRY(gmb_cov_dims(d, &B, &Q, &R));
        if (Q > 65535) return gmb_set_error(GMB_EINVAL, "Laplace path: Q = %d exceeds 65535", Q);
        if (!gmb_flink_core(m->flink)) return gmb_set_error(GMB_EFAMILY, "Laplace path: family/link code %d is not implemented (poisson/log, binomial/logit, gaussian/identity)", m->flink);
        // pinned staging layout of this file: beta at [0, P), v at [1024, 1024 + Q), results from pinned_doubles / 2
        if (P > 1024 || (size_t)Q + 1024 > ctx->pinned_doubles / 2)
            return gmb_set_error(GMB_EINVAL, "Laplace path: P = %d / Q = %d exceed the staging area (P <= 1024, Q <= %zu)", P, Q, ctx->pinned_doubles / 2 - 1024);
        gaussian = std::string(family ? family : "") == "gaussian";
        if (n_start < P + R) return gmb_set_error(GMB_EINVAL, "start has %d values, needs P + R = %d", n_start, P + R);
        beta.assign(start, start + P);
        theta.assign(start + P, start + P + R);
        v.assign(Q, 0.0);                                                          // src/mcml_la.cpp:48
        var_par = 1.0;                                                             // :46
        sigma = gaussian && n_start > P + R ? start[P + R] : 1.0;                  // mcmloptim ctor, mcmloptim.h:30
        kmax = std::max(1, std::min(2048, (int)(ctx->pinned_doubles / 4) / (P + Q)));
        GMB_CUDA(gmb_dmalloc(ctx, &d_par, sizeof(double) * (size_t)(P + 1) * kmax));
        GMB_CUDA(gmb_dmalloc(ctx, &d_V, sizeof(double) * (size_t)ldq * kmax));
        GMB_CUDA(gmb_dmalloc(ctx, &d_W, sizeof(double) * ldn));
        GMB_CUDA(gmb_dmalloc(ctx, &d_zu, sizeof(double) * ldn));
        GMB_CUDA(gmb_dmalloc(ctx, &d_B, sizeof(double) * (size_t)ldn * Q));
        GMB_CUDA(gmb_dmalloc(ctx, &d_M, sizeof(double) * (size_t)ldq * Q));
        GMB_CUDA(gmb_dmalloc(ctx, &d_linv, sizeof(double) * gmb_chol_linv_doubles(Q)));       // inverted diagonal blocks, sized by the factorisation itself
        GMB_CUDA(gmb_dmalloc(ctx, &d_vec, sizeof(double) * (size_t)(ldn + ldq) * 2));
        GMB_CUDA(gmb_dmalloc(ctx, &d_status, sizeof(int)));
        GMB_CUDA(cudaMemsetAsync(d_V, 0, sizeof(double) * (size_t)ldq * kmax, ctx->stream));
        GMB_CUDA(cudaMemsetAsync(d_zu, 0, sizeof(double) * ldn, ctx->stream));
        GMB_TRY(gmb_model_reserve_samples(M, kmax));
        Lhost.assign((size_t)Q * Q, 0.0);
        GMB_TRY(set_L(theta.data(), true));
        D0.assign((size_t)Q * Q, 0.0);                                             // D_ = L L', mcmlmodel.h:67
        for (int j = 0; j < Q; j++)
            for (int k = 0; k <= j; k++) {
                const double ljk = Lhost[j + (size_t)k * Q];
                if (ljk == 0.0) continue;
                for (int i = 0; i < Q; i++) D0[i + (size_t)j * Q] += Lhost[i + (size_t)k * Q] * ljk;
            }
        GMB_TRY(upload_beta(beta.data()));
        GMB_TRY(update_W(false));                                                  // ctor: update_W(), mcmlmodel.h:94
        return GMB_OK;
    }

    // L = chol D(theta) on the device, Z L (mcmlmodel.h:104-106); keep_host also copies L back (for u = L v and D0)
    int set_L(const double* th, bool keep_host) {
        if (!keep_host && (int)theta_L.size() == R && memcmp(theta_L.data(), th, sizeof(double) * R) == 0) return GMB_OK;
        GMB_TRY(gmb_hmc_prepare(M, nullptr));                                      // allocates dL / dZL
        theta_L.clear();
        GMB_TRY(gmb_cov_gen_device(D, th, 1, M->dL, ldq));
        GMB_TRY(gmb_dgemm(ctx, 0, 0, n, Q, Q, 1.0, M->dZ, ldn, M->dL, ldq, 0.0, M->dZL, ldn));
        if (keep_host) {
            GMB_CUDA(cudaMemcpy2DAsync(Lhost.data(), Q * sizeof(double), M->dL, ldq * sizeof(double), Q * sizeof(double), Q, cudaMemcpyDeviceToHost, ctx->stream));
            GMB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        theta_L.assign(th, th + R);
        return GMB_OK;
    }

    int upload_beta(const double* b) {          // update_beta, mcmlmodel.h:100-102: d_par[0:P] = beta, M->dxb = X beta
        GMB_CUDA(cudaStreamSynchronize(ctx->stream));
        memcpy(ctx->h_pinned, b, sizeof(double) * P);
        GMB_CUDA(cudaMemcpyAsync(d_par, ctx->h_pinned, sizeof(double) * P, cudaMemcpyHostToDevice, ctx->stream));
        GMB_TRY(gmb_launch_xb(M, d_par, M->dxb));
        return GMB_OK;
    }

    int upload_v() {                            // column 0 of d_V = v
        GMB_CUDA(cudaStreamSynchronize(ctx->stream));
        memcpy(ctx->h_pinned + 1024, v.data(), sizeof(double) * Q);
        GMB_CUDA(cudaMemcpyAsync(d_V, ctx->h_pinned + 1024, sizeof(double) * Q, cudaMemcpyHostToDevice, ctx->stream));
        return GMB_OK;
    }

    // update_W(0, useL), mcmlmodel.h:120-134, at the current beta (M->dxb), v, var_par
    int update_W(bool useL) {
        GMB_TRY(upload_v());
        GMB_TRY(gmb_dgemm(ctx, 0, 0, n, 1, Q, 1.0, useL ? M->dZL : M->dZ, ldn, d_V, ldq, 0.0, d_zu, ldn));
        const double nvar = gaussian ? var_par * var_par : 1.0;
        la_w_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, M->flink, M->dxb, d_zu, nvar, d_W);
        ctx->launches++;
        GMB_CUDA(cudaGetLastError());
        return GMB_OK;
    }

    double vtv() const { double s = 0; for (int q = 0; q < Q; q++) s += v[q] * v[q]; return s; }

    // ---- LA_likelihood over k points (beta, v), likelihood.h:121-140 ----
    int obj_bv(const double* X, int stride, int k, double* f) {
        for (int off = 0; off < k; off += kmax) {
            const int nb = std::min(kmax, k - off);
            GMB_CUDA(cudaStreamSynchronize(ctx->stream));
            double* hp = ctx->h_pinned;                       // [P x nb] betas, then [Q x nb] states
            for (int c = 0; c < nb; c++) {
                const double* x = X + (size_t)(off + c) * stride;
                memcpy(hp + (size_t)c * P, x, sizeof(double) * P);
                memcpy(hp + (size_t)nb * P + (size_t)c * Q, x + P, sizeof(double) * Q);
            }
            GMB_CUDA(cudaMemcpyAsync(d_par, hp, sizeof(double) * (size_t)nb * P, cudaMemcpyHostToDevice, ctx->stream));
            GMB_CUDA(cudaMemcpy2DAsync(d_V, ldq * sizeof(double), hp + (size_t)nb * P, Q * sizeof(double), Q * sizeof(double), nb,
                                       cudaMemcpyHostToDevice, ctx->stream));
            GMB_TRY(gmb_dgemm(ctx, 0, 0, n, nb, Q, 1.0, M->dZL, ldn, d_V, ldq, 0.0, M->dzd, ldn));            // Z L v, :129
            for (int c = 0; c < nb; c++)
                GMB_TRY(gmb_launch_loglik_cols(M, d_par + (size_t)c * P, var_par, M->dzd + (size_t)c * ldn, 1, ctx->d_result + c));   // :130-133
            double* hres = ctx->h_pinned + ctx->pinned_doubles / 2;
            GMB_CUDA(cudaMemcpyAsync(hres, ctx->d_result, sizeof(double) * nb, cudaMemcpyDeviceToHost, ctx->stream));
            GMB_CUDA(cudaStreamSynchronize(ctx->stream));
            for (int c = 0; c < nb; c++) {
                const double* vv = X + (size_t)(off + c) * stride + P;
                double logl = 0; for (int q = 0; q < Q; q++) logl += vv[q] * vv[q];                              // :126
                f[off + c] = -1.0 * (hres[c] - 0.5 * logl);                                                     // :135
            }
        }
        return GMB_OK;
    }

    // ---- LA_likelihood_cov at (theta, sigma) with the current xb (d_par[0:P] = beta), v, W — likelihood.h:153-180 ----
    int obj_cov_at(const double* th, double sg, double* f) {
        for (int r = 0; r < R; r++) if (!(th[r] == th[r])) { *f = kInf; return GMB_OK; }
        if (gaussian && !(sg > 0.0)) { *f = kInf; return GMB_OK; }
        int rc = set_L(th, false);
        if (rc == GMB_ENOTPD) { *f = kInf; theta_L.clear(); return GMB_OK; }
        GMB_TRY(rc);
        GMB_TRY(upload_v());
        GMB_TRY(gmb_dgemm(ctx, 0, 0, n, 1, Q, 1.0, M->dZL, ldn, d_V, ldq, 0.0, M->dzd, ldn));                  // zd = Z L u, :165
        GMB_TRY(gmb_launch_loglik_cols(M, d_par, sg, M->dzd, 1, ctx->d_result));                                 // :166-169
        dim3 g1((ldn + 255) / 256, Q);
        la_scale_rows_kernel<<<g1, 256, 0, ctx->stream>>>(n, Q, ldn, d_W, M->dZL, d_B);
        dim3 g2((ldq + 255) / 256, Q);
        la_identity_kernel<<<g2, 256, 0, ctx->stream>>>(Q, ldq, d_M);                                            // + I, :176-177
        ctx->launches += 2;
        GMB_TRY(gmb_dgemm(ctx, 1, 0, Q, Q, n, 1.0, M->dZL, ldn, d_B, ldn, 1.0, d_M, ldq));                      // (Z L)' W (Z L), :175
        GMB_CUDA(cudaMemsetAsync(d_status, 0, sizeof(int), ctx->stream));
        GMB_TRY(gmb_chol_blocked(ctx, d_M, ldq, Q, 0, d_status, d_linv, ctx->d_result + 1));                     // logdet, :178
        double* hres = ctx->h_pinned + ctx->pinned_doubles / 2;
        int* hstat = reinterpret_cast<int*>(hres + 4);
        GMB_CUDA(cudaMemcpyAsync(hres, ctx->d_result, sizeof(double) * 2, cudaMemcpyDeviceToHost, ctx->stream));
        GMB_CUDA(cudaMemcpyAsync(hstat, d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GMB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (*hstat != 0) { *f = kInf; return GMB_OK; }
        *f = -1 * (hres[0] - 0.5 * vtv() - 0.5 * hres[1]);                                                      // :180
        return GMB_OK;
    }

    static int cb_bv(const double* X, int nn, int k, double* f, void* user) { return static_cast<LaFit*>(user)->obj_bv(X, nn, k, f); }
    static int cb_cov(const double* X, int nn, int k, double* f, void* user) {
        LaFit* s = static_cast<LaFit*>(user);
        for (int c = 0; c < k; c++) {
            const double* x = X + (size_t)c * nn;
            GMB_TRY(s->obj_cov_at(x, s->gaussian ? x[s->R] : s->var_par, f + c));
        }
        return GMB_OK;
    }
    // LA_likelihood_btheta, likelihood.h:193-229: update_beta, update_W() (Z v), then the covariance objective
    static int cb_btheta(const double* X, int nn, int k, double* f, void* user) {
        LaFit* s = static_cast<LaFit*>(user);
        for (int c = 0; c < k; c++) {
            const double* x = X + (size_t)c * nn;
            const double sg = s->gaussian ? x[s->P + s->R] : s->var_par;
            if (s->gaussian && !(sg > 0.0)) { f[c] = kInf; continue; }
            const double keep = s->var_par;
            s->var_par = sg;                                   // M_->var_par_ = par2.back(), :203
            GMB_TRY(s->upload_beta(x));
            GMB_TRY(s->update_W(false));
            int rc = s->obj_cov_at(x + s->P, sg, f + c);
            s->var_par = keep;
            GMB_TRY(rc);
        }
        return GMB_OK;
    }

    // ---- M-steps, mcmloptim.h:116-177 ----
    int la_optim() {
        std::vector<double> x(beta);
        x.insert(x.end(), v.begin(), v.end());
        GMB_TRY(gmb_minimize_bounded(cb_bv, this, P + Q, x.data(), nullptr, nullptr, 0.0, 1e-8, 300, nullptr, nullptr));
        beta.assign(x.begin(), x.begin() + P);
        v.assign(x.begin() + P, x.end());
        return GMB_OK;
    }
    int la_optim_cov() {
        std::vector<double> x(theta), lo(R, 1e-6);
        if (gaussian) { x.push_back(sigma); lo.push_back(0.0); }
        GMB_TRY(gmb_minimize_bounded(cb_cov, this, (int)x.size(), x.data(), lo.data(), nullptr, 0.0, 1e-8, 200, nullptr, nullptr));
        theta.assign(x.begin(), x.begin() + R);
        if (gaussian) sigma = x[R];
        return GMB_OK;
    }
    int la_optim_bcov() {
        std::vector<double> x(beta), lo(P, -kInf);
        for (int r = 0; r < R; r++) { x.push_back(theta[r]); lo.push_back(1e-6); }
        if (gaussian) { x.push_back(sigma); lo.push_back(0.0); }
        GMB_TRY(gmb_minimize_bounded(cb_btheta, this, (int)x.size(), x.data(), lo.data(), nullptr, 0.0, 1e-8, 300, nullptr, nullptr));
        beta.assign(x.begin(), x.begin() + P);
        theta.assign(x.begin() + P, x.begin() + P + R);
        if (gaussian) sigma = x[P + R];
        return GMB_OK;
    }
    int hess_la(double tol, std::vector<double>& H, int* nvar_out) {               // mcmloptim.h:179-195
        const int nvar = P + R + (gaussian ? 1 : 0);
        std::vector<double> x(beta), nd(nvar, tol);
        x.insert(x.end(), theta.begin(), theta.end());
        if (gaussian) x.push_back(sigma);
        H.assign((size_t)nvar * nvar, 0.0);
        *nvar_out = nvar;
        return gmb_fd_hessian(cb_btheta, this, nvar, x.data(), nd.data(), nullptr, nullptr, 0, H.data(), nullptr);
    }

    // ---- mcnr_b, mcmloptim.h:238-293 (the O(n P^2) and O(Q^3) pieces on the host) ----
    int mcnr_b() {
        std::vector<double> zd(n), zv(n), W(n), Mh((size_t)Q * Q), xb(n, 0.0);
        GMB_TRY(upload_v());
        GMB_TRY(gmb_dgemm(ctx, 0, 0, n, 1, Q, 1.0, M->dZL, ldn, d_V, ldq, 0.0, d_vec, ldn));                    // zd = Z L u, :244
        GMB_TRY(gmb_dgemm(ctx, 0, 0, n, 1, Q, 1.0, M->dZ, ldn, d_V, ldq, 0.0, d_vec + ldn, ldn));               // Z u for log_grad(.., false)
        dim3 g1((ldn + 255) / 256, Q);
        la_scale_rows_kernel<<<g1, 256, 0, ctx->stream>>>(n, Q, ldn, d_W, M->dZL, d_B);
        dim3 g2((ldq + 255) / 256, Q);
        la_identity_kernel<<<g2, 256, 0, ctx->stream>>>(Q, ldq, d_M);
        ctx->launches += 2;
        GMB_TRY(gmb_dgemm(ctx, 1, 0, Q, Q, n, 1.0, M->dZL, ldn, d_B, ldn, 1.0, d_M, ldq));                      // LZWZL + I, :248-251
        GMB_CUDA(cudaMemcpyAsync(zd.data(), d_vec, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
        GMB_CUDA(cudaMemcpyAsync(zv.data(), d_vec + ldn, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
        GMB_CUDA(cudaMemcpyAsync(W.data(), d_W, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
        GMB_CUDA(cudaMemcpy2DAsync(Mh.data(), Q * sizeof(double), d_M, ldq * sizeof(double), Q * sizeof(double), Q, cudaMemcpyDeviceToHost, ctx->stream));
        GMB_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int p = 0; p < P; p++) for (int i = 0; i < n; i++) xb[i] += hX[i + (size_t)p * n] * beta[p];
        const int fl = M->flink;
        std::vector<double> resid(n), Wu(n), r2(n);
        double mean = 0;
        for (int i = 0; i < n; i++) {
            const double eta = xb[i] + zd[i];
            double mu = eta, dmu = 1.0;                                                                          // mod_inv_func / detadmu
            if (fl == 1) { mu = std::exp(eta); dmu = std::exp(-1.0 * eta); }
            else if (fl == 3) { mu = std::exp(eta) / (1 + std::exp(eta)); dmu = 1 / (mu * (1.0 - mu)); }
            resid[i] = hy[i] - mu; mean += resid[i];                                                             // :255-256
            Wu[i] = W[i] * dmu * resid[i];                                                                       // :260-262
            const double eta2 = xb[i] + zv[i];                                                                   // log_grad(v, false): mu = xb + Z v
            if (fl == 1) r2[i] = hy[i] - std::exp(eta2);
            else if (fl == 3) r2[i] = 1.0 / (std::exp(eta2) + 1.0) + hy[i] - 1.0;
            else r2[i] = (hy[i] - eta2) / (var_par * var_par);
        }
        mean /= n;
        double ss = 0; for (int i = 0; i < n; i++) ss += (resid[i] - mean) * (resid[i] - mean);
        const double sigmas = std::sqrt(ss / (n - 1));                                                           // :257
        std::vector<double> XtWX((size_t)P * P, 0.0), XtWu(P, 0.0), bincr(P);
        for (int a = 0; a < P; a++) {
            for (int b = 0; b < P; b++) { double s = 0; for (int i = 0; i < n; i++) s += hX[i + (size_t)a * n] * W[i] * hX[i + (size_t)b * n]; XtWX[a + (size_t)b * P] = s; }
            double s = 0; for (int i = 0; i < n; i++) s += hX[i + (size_t)a * n] * Wu[i]; XtWu[a] = s;
        }
        if (gmb_solve_small(P, XtWX.data(), XtWu.data(), bincr.data())) return gmb_set_error(GMB_ENUMERIC, "mcnr_b: X'WX is singular");   // :264-267
        // vgrad = -D v + (Z L)' r, :268 ; vincr = (LZWZL + I)^-1 vgrad, :269
        GMB_CUDA(cudaStreamSynchronize(ctx->stream));
        if ((size_t)n > ctx->pinned_doubles / 2) return gmb_set_error(GMB_EINVAL, "mcnr_b: n = %d exceeds the staging buffer", n);
        memcpy(ctx->h_pinned, r2.data(), sizeof(double) * n);
        GMB_CUDA(cudaMemsetAsync(d_vec, 0, sizeof(double) * ldn, ctx->stream));
        GMB_CUDA(cudaMemcpyAsync(d_vec, ctx->h_pinned, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
        GMB_TRY(gmb_dgemm(ctx, 1, 0, Q, 1, n, 1.0, M->dZL, ldn, d_vec, ldn, 0.0, d_vec + 2 * (size_t)ldn, ldq));
        std::vector<double> g(Q), vincr(Q);
        GMB_CUDA(cudaMemcpyAsync(g.data(), d_vec + 2 * (size_t)ldn, sizeof(double) * Q, cudaMemcpyDeviceToHost, ctx->stream));
        GMB_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < Q; i++) { double s = 0; for (int j = 0; j < Q; j++) s += D0[i + (size_t)j * Q] * v[j]; g[i] += -1.0 * s; }
        if (gmb_solve_small(Q, Mh.data(), g.data(), vincr.data())) return gmb_set_error(GMB_ENUMERIC, "mcnr_b: (ZL)'W(ZL) + I is singular");
        for (int q = 0; q < Q; q++) v[q] += vincr[q];                                                            // :290
        for (int p = 0; p < P; p++) beta[p] += bincr[p];                                                         // :291
        sigma = sigmas;                                                                                          // :292
        return GMB_OK;
    }
};

// common driver of mcml_la (nr = 0) and mcml_la_nr (nr = 1)
int run_la(int nr, const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
           const double* Z, const double* X, const double* y, int n, int P, int Q, const char* family, const char* link,
           const double* start, int n_start, int usehess, double tol, int verbose, int maxiter,
           double* beta_out, double* theta_out, double* sigma_out, double* se_out, double* u_out, int* iter_out) {
    if (!cov || !data || !Z || !X || !y || !start) return gmb_set_error(GMB_EINVAL, "mcml_la: NULL argument");
    gmb_ctx* ctx; GMB_TRY(gmb_default_ctx(&ctx));
    struct H { gmb_cov* cv = nullptr; gmb_model* mdl = nullptr; ~H() { if (mdl) gmb_model_destroy(mdl); if (cv) gmb_cov_destroy(cv); } } h;
    GMB_TRY(gmb_cov_create(ctx, cov, cov_rows, data, n_data, eff_range, n_eff, &h.cv));
    int B, Qc, R;
    GMB_TRY(gmb_cov_dims(h.cv, &B, &Qc, &R));
    if (Qc != Q) return gmb_set_error(GMB_EINVAL, "covariance has %d random effects, Z has %d columns", Qc, Q);
    if (n_start < P + R + 1) return gmb_set_error(GMB_EINVAL, "start needs at least P + R + 1 = %d values (src/mcml_la.cpp:50,81)", P + R + 1);
    GMB_TRY(gmb_model_create(ctx, n, P, Q, X, Z, y, family, link, &h.mdl));
    const std::string fam(family ? family : "");
    const bool has_var = fam == "gaussian" || fam == "Gamma";
    LaFit mc;
    GMB_TRY(mc.init(ctx, h.cv, h.mdl, X, y, start, n_start, family));
    if (nr) GMB_TRY(mc.update_W(true));                                            // src/mcml_la.cpp:199
    std::vector<double> beta(mc.beta), theta(mc.theta), newbeta(P), newtheta(R);
    double var_par = 1.0, new_var_par = 1.0, maxdiff = 1.0;
    int iter = 1; bool converged = false;
    while (maxdiff > tol && iter <= maxiter) {                                     // :62 / :213
        if (!nr) GMB_TRY(mc.la_optim()); else GMB_TRY(mc.mcnr_b());               // :64 / :216
        newbeta = mc.beta;
        GMB_TRY(mc.upload_beta(newbeta.data()));                                   // model.update_beta(newbeta), :66
        GMB_TRY(mc.update_W(nr != 0));                                             // :67 / :219
        GMB_TRY(mc.la_optim_cov());                                                // :68
        newtheta = mc.theta;
        if (has_var) new_var_par = mc.sigma;                                       // :70
        maxdiff = 0.0;
        for (int p = 0; p < P; p++) maxdiff = std::max(maxdiff, std::fabs(beta[p] - newbeta[p]));
        for (int r = 0; r < R; r++) maxdiff = std::max(maxdiff, std::fabs(theta[r] - newtheta[r]));
        maxdiff = std::max(maxdiff, std::fabs(var_par - new_var_par));
        if (maxdiff < tol) converged = true;                                       // :77
        beta = newbeta; theta = newtheta; var_par = new_var_par;
        if (!converged) {                                                          // :83-91 / :236-244
            // the reference refreshes W BEFORE Z L: mcml_la forms it at xb + Z v with the previous var_par_ (:90-93), mcml_la_nr at
            // xb + (Z L_old) v with the new one (:239-243); the new factor only enters with update_L() afterwards
            GMB_TRY(mc.upload_beta(beta.data()));
            if (nr) mc.var_par = new_var_par;
            GMB_TRY(mc.update_W(nr != 0));
            mc.var_par = new_var_par;
            GMB_TRY(mc.set_L(theta.data(), true));
        }
        if (verbose) {
            fprintf(stderr, "Iter %d  beta:", iter);
            for (int p = 0; p < P; p++) fprintf(stderr, " %.5f", beta[p]);
            fprintf(stderr, "  theta:");
            for (int r = 0; r < R; r++) fprintf(stderr, " %.5f", theta[r]);
            fprintf(stderr, "  sigma: %.5f  max diff: %.3g%s\n", var_par, maxdiff, converged ? "  CONVERGED" : "");
        }
        iter++;
    }
    GMB_TRY(mc.la_optim_bcov());                                                   // :107 / :260
    beta = mc.beta; theta = mc.theta;
    if (has_var) var_par = mc.sigma;
    if (se_out) {
        for (int i = 0; i < n_start; i++) se_out[i] = 0.0;                         // :121
        if (usehess) {                                                             // :123-129
            std::vector<double> Hm; int nvar = 0;
            GMB_TRY(mc.hess_la(1e-4, Hm, &nvar));
            // hess.llt().solve(I): column i of the inverse by a small solve (stands in for Eigen's LLT)
            std::vector<double> e(nvar), col(nvar);
            for (int i = 0; i < nvar && i < n_start; i++) {
                std::fill(e.begin(), e.end(), 0.0); e[i] = 1.0;
                if (gmb_solve_small(nvar, Hm.data(), e.data(), col.data())) { se_out[i] = std::numeric_limits<double>::quiet_NaN(); continue; }
                se_out[i] = std::sqrt(col[i]);
            }
        }
    }
    if (beta_out) memcpy(beta_out, beta.data(), sizeof(double) * P);
    if (theta_out) memcpy(theta_out, theta.data(), sizeof(double) * R);
    if (sigma_out) *sigma_out = var_par;
    if (iter_out) *iter_out = iter - 1;
    if (u_out)                                                                     // u = L * u with the L of the last refresh, :148
        for (int i = 0; i < Q; i++) { double s = 0; for (int j = 0; j < Q; j++) s += mc.Lhost[i + (size_t)j * Q] * mc.v[j]; u_out[i] = s; }
    return GMB_OK;
}

}  // namespace

// test hooks: the three Laplace objectives and one mcnr_b step at caller-supplied states (parity tests against oracle/laplace.py)
extern "C" int gmb_la_objectives(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                                 const double* Z, const double* X, const double* y, int n, int P, int Q, const char* family, const char* link,
                                 const double* beta, const double* theta, int R_in, const double* v, double sigma, int w_use_l,
                                 double* out3, double* beta_nr, double* v_nr, double* sigma_nr) {
    if (!cov || !data || !Z || !X || !y || !beta || !theta || !v || !out3) return gmb_set_error(GMB_EINVAL, "gmb_la_objectives: NULL argument");
    gmb_ctx* ctx; GMB_TRY(gmb_default_ctx(&ctx));
    struct H { gmb_cov* cv = nullptr; gmb_model* mdl = nullptr; ~H() { if (mdl) gmb_model_destroy(mdl); if (cv) gmb_cov_destroy(cv); } } h;
    GMB_TRY(gmb_cov_create(ctx, cov, cov_rows, data, n_data, eff_range, n_eff, &h.cv));
    int B, Qc, R;
    GMB_TRY(gmb_cov_dims(h.cv, &B, &Qc, &R));
    if (Qc != Q || R != R_in) return gmb_set_error(GMB_EINVAL, "gmb_la_objectives: covariance dimensions do not match");
    GMB_TRY(gmb_model_create(ctx, n, P, Q, X, Z, y, family, link, &h.mdl));
    std::vector<double> start(beta, beta + P);
    start.insert(start.end(), theta, theta + R);
    start.push_back(sigma);
    LaFit mc;
    GMB_TRY(mc.init(ctx, h.cv, h.mdl, X, y, start.data(), (int)start.size(), family));
    mc.v.assign(v, v + Q);
    mc.var_par = mc.gaussian ? sigma : 1.0;
    mc.sigma = sigma;
    GMB_TRY(mc.upload_beta(beta));
    GMB_TRY(mc.update_W(w_use_l != 0));
    std::vector<double> x(beta, beta + P);
    x.insert(x.end(), v, v + Q);
    GMB_TRY(mc.obj_bv(x.data(), P + Q, 1, out3));                                  // LA_likelihood(beta, v)
    GMB_TRY(mc.obj_cov_at(theta, mc.var_par, out3 + 1));                           // LA_likelihood_cov(theta [, sigma]) with W as set above
    std::vector<double> xb(beta, beta + P);
    xb.insert(xb.end(), theta, theta + R);
    if (mc.gaussian) xb.push_back(sigma);
    GMB_TRY(LaFit::cb_btheta(xb.data(), (int)xb.size(), 1, out3 + 2, &mc));        // LA_likelihood_btheta (refreshes W with Z v)
    if (beta_nr && v_nr) {
        GMB_TRY(mc.upload_beta(beta));
        GMB_TRY(mc.set_L(theta, false));
        GMB_TRY(mc.update_W(w_use_l != 0));
        GMB_TRY(mc.mcnr_b());
        memcpy(beta_nr, mc.beta.data(), sizeof(double) * P);
        memcpy(v_nr, mc.v.data(), sizeof(double) * Q);
        if (sigma_nr) *sigma_nr = mc.sigma;
    }
    return GMB_OK;
}

// mcml_la, src/mcml_la.cpp:28-155
extern "C" int gmb_mcml_la(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                           const double* Z, const double* X, const double* y, int n, int P, int Q, const char* family, const char* link,
                           const double* start, int n_start, int usehess, double tol, int verbose, int trace, int maxiter,
                           double* beta_out, double* theta_out, double* sigma_out, double* se_out, double* u_out, int* iter_out) {
    (void)trace;
    return run_la(0, cov, cov_rows, data, n_data, eff_range, n_eff, Z, X, y, n, P, Q, family, link, start, n_start, usehess, tol, verbose, maxiter,
                  beta_out, theta_out, sigma_out, se_out, u_out, iter_out);
}

// mcml_la_nr, src/mcml_la.cpp:178-290
extern "C" int gmb_mcml_la_nr(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                              const double* Z, const double* X, const double* y, int n, int P, int Q, const char* family, const char* link,
                              const double* start, int n_start, int usehess, double tol, int verbose, int trace, int maxiter,
                              double* beta_out, double* theta_out, double* sigma_out, double* se_out, double* u_out, int* iter_out) {
    (void)trace;
    return run_la(1, cov, cov_rows, data, n_data, eff_range, n_eff, Z, X, y, n, P, Q, family, link, start, n_start, usehess, tol, verbose, maxiter,
                  beta_out, theta_out, sigma_out, se_out, u_out, iter_out);
}
