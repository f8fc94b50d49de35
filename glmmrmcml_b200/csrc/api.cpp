// api.cpp — the reference-named entry points (what src/RcppExports.cpp:291-305 forwards to once the Rcpp bodies are
// thin adapters; INTEGRATION.md shows the adapters).  Each function restates the BODY of its reference counterpart in
// src/mcml_full.cpp / src/mcml_optim.cpp on top of the device objects gmb_model / gmb_cov; the class `Fit` below plays
// the role of glmmr::mcmloptim<MCMLDmatrix> (inst/include/glmmrmcml/mcmloptim.h:16-372) and the three objective
// callbacks the role of likelihood.h's D_/L_/F_likelihood functors (likelihood.h:31-110).
#include "common.cuh"
#include <map>
#include <cstring>
#include <string>
#include <vector>
#include <limits>
#include <algorithm>

namespace {

gmb_ctx* g_default_ctx = nullptr;
bool g_default_owned = false;

int default_ctx(gmb_ctx** out) {
    if (!g_default_ctx) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
        GMB_TRY(gmb_ctx_create(dev, &g_default_ctx));
        g_default_owned = true;
    }
    *out = g_default_ctx;
    return GMB_OK;
}

const double kInf = std::numeric_limits<double>::infinity();
int g_importance_reference_form = 0;     // gmb_mcml_set_importance_form

struct Fit {
    gmb_ctx* ctx = nullptr;
    gmb_cov* D = nullptr;
    gmb_model* M = nullptr;
    int P = 0, R = 0;
    bool gaussian = false;
    std::vector<double> beta, theta, cov_par_fix;
    double sigma = 0.0;          // mcmloptim::sigma_  (mcmloptim.h:30)
    double model_var_par = 1.0;  // mcmlModel::var_par_ as the entry point constructed it (e.g. src/mcml_optim.cpp:52)
    int d_cols_total = 0;        // columns MCMLDmatrix::loglik averages (u.cols(), mcmldmatrix.h:24)
    int nfev_ll = 0, nfev_d = 0;
    std::map<std::vector<double>, double> memo_ll, memo_d;   // identical parameter vectors are evaluated once per sample matrix

    void new_samples() { memo_ll.clear(); memo_d.clear(); }

    // mcmloptim ctor, mcmloptim.h:19-39
    int init(gmb_ctx* c, gmb_cov* d, gmb_model* m, const double* start, int n_start, const char* family) {
        ctx = c; D = d; M = m; P = m->P;
        int B, Q;
        GMB_TRY(gmb_cov_dims(d, &B, &Q, &R));
        gaussian = std::string(family ? family : "") == "gaussian";
        if (n_start < P + R + (gaussian ? 1 : 0))
            return gmb_set_error(GMB_EINVAL, "start has %d values, needs %d (P=%d beta, R=%d theta%s)", n_start, P + R + (gaussian ? 1 : 0), P, R, gaussian ? ", sigma" : "");
        beta.assign(start, start + P);
        theta.assign(start + P, start + P + R);
        cov_par_fix = theta;
        sigma = gaussian ? start[P + R] : 0.0;
        return GMB_OK;
    }

    // -mean log-likelihood for k parameter columns (beta, optional sigma); memoised
    int eval_ll(const double* X, int stride, int k, bool sigma_in_par, double fixed_sigma, double* out) {
        std::vector<int> todo;
        std::vector<double> B, S;
        std::vector<std::vector<double>> keys(k);
        for (int c = 0; c < k; c++) {
            const double* x = X + (size_t)c * stride;
            double sg = sigma_in_par ? x[P] : fixed_sigma;
            keys[c].assign(x, x + P); keys[c].push_back(sg);
            if (gaussian && !(sg > 0.0)) { out[c] = -kInf; continue; }
            auto itf = memo_ll.find(keys[c]);
            if (itf != memo_ll.end()) { out[c] = itf->second; continue; }
            bool dup = false;
            for (int t : todo) if (keys[t] == keys[c]) { dup = true; break; }
            if (!dup) { todo.push_back(c); B.insert(B.end(), x, x + P); S.push_back(sg); }
            out[c] = std::numeric_limits<double>::quiet_NaN();
        }
        for (size_t off = 0; off < todo.size(); off += 4096) {
            int nb = (int)std::min<size_t>(4096, todo.size() - off);
            std::vector<double> res(nb);
            GMB_TRY(gmb_model_loglik_batch(M, B.data() + off * P, S.data() + off, nb, res.data()));
            for (int t = 0; t < nb; t++) memo_ll[keys[todo[off + t]]] = res[t];
            nfev_ll += nb;
        }
        for (int c = 0; c < k; c++) if (out[c] != out[c]) { auto itf = memo_ll.find(keys[c]); out[c] = itf != memo_ll.end() ? itf->second : out[c]; }
        return GMB_OK;
    }

    // MCMLDmatrix::loglik at theta on the model's samples; -inf when D(theta) is not positive definite
    int eval_d(const double* th, double* out) {
        std::vector<double> key(th, th + R);
        auto itf = memo_d.find(key);
        if (itf != memo_d.end()) { *out = itf->second; return GMB_OK; }
        for (int r = 0; r < R; r++) if (!(th[r] == th[r])) { *out = -kInf; return GMB_OK; }
        int rc = gmb_cov_mvn_ll_model(D, th, M, d_cols_total, out);
        if (rc == GMB_ENOTPD) { *out = -kInf; rc = GMB_OK; }
        if (rc == GMB_OK) { memo_d[key] = *out; nfev_d++; }
        return rc;
    }

    // the same for k parameter columns (theta at X + c * stride + offset): points not seen before go to the device as one batch
    int eval_d_batch(const double* X, int stride, int offset, int k, double* out) {
        std::vector<int> todo;
        std::vector<double> T;
        std::vector<std::vector<double>> keys(k);
        for (int c = 0; c < k; c++) {
            const double* th = X + (size_t)c * stride + offset;
            keys[c].assign(th, th + R);
            auto itf = memo_d.find(keys[c]);
            if (itf != memo_d.end()) { out[c] = itf->second; continue; }
            bool nan_par = false;
            for (int r = 0; r < R; r++) nan_par |= !(th[r] == th[r]);
            if (nan_par) { out[c] = -kInf; continue; }
            bool dup = false;
            for (int t : todo) if (keys[t] == keys[c]) { dup = true; break; }
            if (!dup) { todo.push_back(c); T.insert(T.end(), th, th + R); }
            out[c] = std::numeric_limits<double>::quiet_NaN();
        }
        if (!todo.empty()) {
            std::vector<double> res(todo.size());
            GMB_TRY(gmb_cov_mvn_ll_model_batch(D, T.data(), (int)todo.size(), M, d_cols_total, res.data()));
            for (size_t t = 0; t < todo.size(); t++) memo_d[keys[todo[t]]] = res[t];
            nfev_d += (int)todo.size();
        }
        for (int c = 0; c < k; c++) if (out[c] != out[c]) { auto itf = memo_d.find(keys[c]); if (itf != memo_d.end()) out[c] = itf->second; }
        return GMB_OK;
    }

    // ---- objectives (likelihood.h) ----
    static int L_obj(const double* X, int n, int k, double* f, void* user) {      // L_likelihood, likelihood.h:57-64
        Fit* self = static_cast<Fit*>(user);
        GMB_TRY(self->eval_ll(X, n, k, self->gaussian, 0.0, f));
        for (int c = 0; c < k; c++) f[c] = -1 * f[c];
        return GMB_OK;
    }
    static int D_obj(const double* X, int n, int k, double* f, void* user) {      // D_likelihood, likelihood.h:40-45
        Fit* self = static_cast<Fit*>(user);
        GMB_TRY(self->eval_d_batch(X, n, 0, k, f));
        for (int c = 0; c < k; c++) f[c] = -1 * f[c];
        return GMB_OK;
    }
    struct FArgs { Fit* self; bool importance; double fix_var_par; double denomD; };
    static int F_obj(const double* X, int n, int k, double* f, void* user) {      // F_likelihood, likelihood.h:88-108 (fix_var = true)
        FArgs* a = static_cast<FArgs*>(user);
        Fit* self = a->self;
        std::vector<double> ll(k);
        GMB_TRY(self->eval_ll(X, n, k, false, a->fix_var_par, ll.data()));
        std::vector<double> dl(k);
        GMB_TRY(self->eval_d_batch(X, n, self->P, k, dl.data()));
        for (int c = 0; c < k; c++) {
            const double logl = dl[c];
            if (!a->importance) f[c] = -1.0 * (ll[c] + logl);
            else if (g_importance_reference_form) {                  // likelihood.h:101-105 as written: du = exp(ll + logl) / exp(denomD); -log(du)
                const double du = std::exp(ll[c] + logl) / std::exp(a->denomD);
                f[c] = -1.0 * std::log(du);
            } else f[c] = -1.0 * (ll[c] + logl - a->denomD);          // the same quantity in log space (default: the written form underflows to
                                                                      // -log(0 / 0) once |ll + logl| exceeds ~745)
        }
        return GMB_OK;
    }

    // ---- M-steps (mcmloptim.h) ----
    int l_optim() {                                                               // mcmloptim.h:71-88
        std::vector<double> x(beta), lo(P, -kInf);
        if (gaussian) { x.push_back(sigma); lo.push_back(0.0); }
        GMB_TRY(gmb_minimize_bounded(L_obj, this, (int)x.size(), x.data(), lo.data(), nullptr, 0.0, 1e-8, 200, nullptr, nullptr));
        beta.assign(x.begin(), x.begin() + P);
        if (gaussian) sigma = x[P];
        return GMB_OK;
    }
    int d_optim() {                                                               // mcmloptim.h:56-68
        std::vector<double> x(theta), lo(R, 1e-6);
        GMB_TRY(gmb_minimize_bounded(D_obj, this, R, x.data(), lo.data(), nullptr, 0.0, 1e-8, 200, nullptr, nullptr));
        theta = x;
        return GMB_OK;
    }
    int f_optim() {                                                               // mcmloptim.h:91-113
        FArgs a{this, true, sigma, 0.0};
        GMB_TRY(eval_d(cov_par_fix.data(), &a.denomD));
        std::vector<double> x(beta), lo(P, -kInf);
        for (int r = 0; r < R; r++) { x.push_back(theta[r]); lo.push_back(1e-6); }
        // the reference also hands sigma to BOBYQA for gaussian models, but the objective never reads it (fix_var = true,
        // likelihood.h:95): the direction is flat and sigma comes back unchanged up to the optimiser's wandering
        GMB_TRY(gmb_minimize_bounded(F_obj, &a, P + R, x.data(), lo.data(), nullptr, 0.0, 1e-8, 300, nullptr, nullptr));
        beta.assign(x.begin(), x.begin() + P);
        theta.assign(x.begin() + P, x.begin() + P + R);
        return GMB_OK;
    }
    int mcnr() {                                                                  // mcmloptim.h:198-236
        std::vector<double> incr(P);
        double sg = 0.0;
        GMB_TRY(gmb_model_mcnr(M, beta.data(), model_var_par, nullptr, nullptr, incr.data(), &sg));
        for (int p = 0; p < P; p++) beta[p] += incr[p];
        sigma = sg;
        return GMB_OK;
    }
    int f_hess(double tol, double* hess) {                                        // mcmloptim.h:333-355
        FArgs a{this, false, sigma, 0.0};
        const int k = P + R;
        std::vector<double> x(beta), lo(P, -kInf), up(k, kInf), nd(k, tol);
        for (int r = 0; r < R; r++) { x.push_back(theta[r]); lo.push_back(1e-6); }
        return gmb_fd_hessian(F_obj, &a, k, x.data(), nd.data(), lo.data(), up.data(), 1, hess, nullptr);
    }
};

// ---- device objects kept between calls -----------------------------------------------------------------------------------------
// The reference-named entry points take host arrays and (in the reference) rebuild their C++ objects on every call; an R loop calls
// mcmc_sample / mcml_optim / mcml_hess again and again with the SAME X, Z, y and covariance specification.  Building the device objects
// (uploads, row aggregation, sparse forms of Z, block classes) costs ~0.5-1 ms per call at C2 — 13 % of the end-to-end step — so the last few
// models and covariance objects of a context are kept, keyed by an EXACT comparison (memcmp) of their defining arrays; everything that depends
// on the call's other arguments (u, L, beta, theta) is set by the call as before.  Objects of more than 8 MB of host data are not kept.
struct CachedModel { gmb_ctx* ctx; int n, P, Q; std::string fam, link; std::vector<double> X, Z, y; gmb_model* mdl; bool busy; unsigned long long stamp; };
struct CachedCov { gmb_ctx* ctx; int rows, n_data, n_eff; std::vector<int32_t> cov; std::vector<double> data, eff; gmb_cov* cv; bool busy; unsigned long long stamp; };
std::vector<CachedModel> g_models;
std::vector<CachedCov> g_covs;
unsigned long long g_stamp = 0;
int g_object_cache = 1;
constexpr size_t kCacheMaxDoubles = (size_t)1 << 20;
constexpr size_t kCacheEntries = 4;

int acquire_model(gmb_ctx* ctx, int n, int P, int Q, const double* X, const double* Z, const double* y, const char* family, const char* link,
                  gmb_model** out, bool* cached) {
    *cached = false;
    const size_t total = (size_t)n * ((size_t)P + Q + 1);
    if (!g_object_cache || total > kCacheMaxDoubles || !family || !link) return gmb_model_create(ctx, n, P, Q, X, Z, y, family, link, out);
    for (auto& e : g_models)
        if (!e.busy && e.ctx == ctx && e.n == n && e.P == P && e.Q == Q && e.fam == family && e.link == link &&
            memcmp(e.y.data(), y, sizeof(double) * n) == 0 && memcmp(e.X.data(), X, sizeof(double) * (size_t)n * P) == 0 &&
            memcmp(e.Z.data(), Z, sizeof(double) * (size_t)n * Q) == 0) {
            e.busy = true; e.stamp = ++g_stamp; *out = e.mdl; *cached = true;
            return GMB_OK;
        }
    GMB_TRY(gmb_model_create(ctx, n, P, Q, X, Z, y, family, link, out));
    if (g_models.size() >= kCacheEntries) {                                   // evict the least recently used idle entry
        int victim = -1;
        for (size_t k = 0; k < g_models.size(); k++) if (!g_models[k].busy && (victim < 0 || g_models[k].stamp < g_models[victim].stamp)) victim = (int)k;
        if (victim < 0) return GMB_OK;                                        // all busy: this one is not kept
        gmb_model_destroy(g_models[victim].mdl);
        g_models.erase(g_models.begin() + victim);
    }
    g_models.push_back(CachedModel{ctx, n, P, Q, family, link, std::vector<double>(X, X + (size_t)n * P), std::vector<double>(Z, Z + (size_t)n * Q),
                                   std::vector<double>(y, y + n), *out, true, ++g_stamp});
    *cached = true;
    return GMB_OK;
}
void release_model(gmb_model* mdl, bool cached) {
    if (!mdl) return;
    if (cached) { for (auto& e : g_models) if (e.mdl == mdl) { e.busy = false; return; } }
    gmb_model_destroy(mdl);
}

int acquire_cov(gmb_ctx* ctx, const int32_t* cov, int rows, const double* data, int n_data, const double* eff, int n_eff, gmb_cov** out, bool* cached) {
    *cached = false;
    if (!g_object_cache || !cov || !data || rows <= 0 || n_data < 0 || (size_t)n_data > kCacheMaxDoubles || (size_t)rows > kCacheMaxDoubles)
        return gmb_cov_create(ctx, cov, rows, data, n_data, eff, n_eff, out);
    const int ne = eff ? n_eff : 0;
    for (auto& e : g_covs)
        if (!e.busy && e.ctx == ctx && e.rows == rows && e.n_data == n_data && e.n_eff == ne &&
            memcmp(e.cov.data(), cov, sizeof(int32_t) * (size_t)rows * 5) == 0 && memcmp(e.data.data(), data, sizeof(double) * n_data) == 0 &&
            (ne == 0 || memcmp(e.eff.data(), eff, sizeof(double) * ne) == 0)) {
            e.busy = true; e.stamp = ++g_stamp; *out = e.cv; *cached = true;
            return GMB_OK;
        }
    GMB_TRY(gmb_cov_create(ctx, cov, rows, data, n_data, eff, n_eff, out));
    if (g_covs.size() >= kCacheEntries) {
        int victim = -1;
        for (size_t k = 0; k < g_covs.size(); k++) if (!g_covs[k].busy && (victim < 0 || g_covs[k].stamp < g_covs[victim].stamp)) victim = (int)k;
        if (victim < 0) return GMB_OK;
        gmb_cov_destroy(g_covs[victim].cv);
        g_covs.erase(g_covs.begin() + victim);
    }
    g_covs.push_back(CachedCov{ctx, rows, n_data, ne, std::vector<int32_t>(cov, cov + (size_t)rows * 5), std::vector<double>(data, data + n_data),
                               ne ? std::vector<double>(eff, eff + ne) : std::vector<double>(), *out, true, ++g_stamp});
    *cached = true;
    return GMB_OK;
}
void release_cov(gmb_cov* cv, bool cached) {
    if (!cv) return;
    if (cached) { for (auto& e : g_covs) if (e.cv == cv) { e.busy = false; return; } }
    gmb_cov_destroy(cv);
}

struct Handles {
    gmb_cov* cv = nullptr; gmb_model* mdl = nullptr; int m_total = 0;
    bool cv_cached = false, mdl_cached = false;
    int make_model(gmb_ctx* ctx, int n, int P, int Q, const double* X, const double* Z, const double* y, const char* family, const char* link) {
        return acquire_model(ctx, n, P, Q, X, Z, y, family, link, &mdl, &mdl_cached);
    }
    int make_cov(gmb_ctx* ctx, const int32_t* cov, int rows, const double* data, int n_data, const double* eff, int n_eff) {
        return acquire_cov(ctx, cov, rows, data, n_data, eff, n_eff, &cv, &cv_cached);
    }
    ~Handles() { release_model(mdl, mdl_cached); release_cov(cv, cv_cached); }
};

int check_common(const void* cov, const void* data, const void* Z, const void* X, const void* y, int n, int P, int Q) {
    if (!cov || !data || !Z || !X || !y) return gmb_set_error(GMB_EINVAL, "NULL input array");
    if (n <= 0 || P <= 0 || Q <= 0) return gmb_set_error(GMB_EINVAL, "bad dimensions n=%d P=%d Q=%d", n, P, Q);
    return GMB_OK;
}

// shared front half of mcml_optim / mcml_simlik / mcml_hess / aic_mcml: DData + MCMLDmatrix + mcmlModel(Z, nullptr, X, y, &u, beta, var_par)
int setup_fixed_u(gmb_ctx* ctx, const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                  const double* Z, const double* X, const double* y, const double* u, int n, int P, int Q, int m,
                  const char* family, const char* link, Handles& h) {
    GMB_TRY(check_common(cov, data, Z, X, y, n, P, Q));
    if (!u || m <= 0) return gmb_set_error(GMB_EINVAL, "u must have at least one column");
    GMB_TRY(h.make_cov(ctx, cov, cov_rows, data, n_data, eff_range, n_eff));
    int B, Qc, R;
    GMB_TRY(gmb_cov_dims(h.cv, &B, &Qc, &R));
    if (Qc != Q) return gmb_set_error(GMB_EINVAL, "covariance has %d random effects, Z has %d columns", Qc, Q);
    GMB_TRY(h.make_model(ctx, n, P, Q, X, Z, y, family, link));
    // with an NCCL-enabled default context every rank passes ITS columns of u; the averages run over all of them
    double mt = (double)m;
    GMB_TRY(gmb_comm_allreduce_host(ctx, &mt, 1));
    h.m_total = (int)(mt + 0.5);
    GMB_TRY(gmb_model_set_u(h.mdl, u, Q, m, h.m_total, h.m_total));
    return GMB_OK;
}

int default_chains(int m) {
    int c = (m + 1 + 31) / 32;
    return c < 1 ? 1 : (c > 1024 ? 1024 : c);
}

}  // namespace

int gmb_default_ctx(gmb_ctx** out) { return default_ctx(out); }   // for laplace.cu

// objects kept for `ctx` are destroyed with it (called by gmb_ctx_destroy before the context's streams go away)
void gmb_api_release_ctx(gmb_ctx* ctx) {
    for (size_t k = g_models.size(); k-- > 0;) if (g_models[k].ctx == ctx) { gmb_model_destroy(g_models[k].mdl); g_models.erase(g_models.begin() + k); }
    for (size_t k = g_covs.size(); k-- > 0;) if (g_covs[k].ctx == ctx) { gmb_cov_destroy(g_covs[k].cv); g_covs.erase(g_covs.begin() + k); }
    if (g_default_ctx == ctx) { g_default_ctx = nullptr; g_default_owned = false; }
}
extern "C" int gmb_set_object_cache(int on) {
    g_object_cache = on ? 1 : 0;
    if (!on) {
        for (size_t k = g_models.size(); k-- > 0;) if (!g_models[k].busy) { gmb_model_destroy(g_models[k].mdl); g_models.erase(g_models.begin() + k); }
        for (size_t k = g_covs.size(); k-- > 0;) if (!g_covs[k].busy) { gmb_cov_destroy(g_covs[k].cv); g_covs.erase(g_covs.begin() + k); }
    }
    return GMB_OK;
}

extern "C" int gmb_mcml_set_importance_form(int reference_form) { g_importance_reference_form = reference_form ? 1 : 0; return GMB_OK; }

// DData::n_cov_pars() and the total block dimension from the cov matrix alone (host arithmetic; lets an adapter size
// its theta output before calling an entry point).  Parameter counts per function id: R/R6ModelExtMCML.R:430.
extern "C" int gmb_cov_shape(const int32_t* cov, int rows, int* B_out, int* Q_out, int* R_out) {
    if (!cov || rows <= 0) return gmb_set_error(GMB_EINVAL, "gmb_cov_shape: bad arguments");
    static const int np[15] = {0, 1, 1, 1, 2, 2, 1, 2, 2, 2, 2, 2, 2, 2, 1};
    int B = 0, R = 0;
    for (int r = 0; r < rows; r++) {
        const int b = cov[r], id = cov[r + 2 * rows], p0 = cov[r + 4 * rows];
        if (b < 0 || id < 1 || id > 14 || p0 < 0) return gmb_set_error(GMB_EINVAL, "bad covariance row %d", r);
        if (b + 1 > B) B = b + 1;
        if (p0 + np[id] > R) R = p0 + np[id];
    }
    std::vector<int> dim(B, 0);
    for (int r = 0; r < rows; r++) dim[cov[r]] = cov[r + rows];
    long long Q = 0;
    for (int b = 0; b < B; b++) { if (dim[b] <= 0) return gmb_set_error(GMB_EINVAL, "block ids must be 0..B-1 without gaps"); Q += dim[b]; }
    if (B_out) *B_out = B;
    if (Q_out) *Q_out = (int)Q;
    if (R_out) *R_out = R;
    return GMB_OK;
}

extern "C" int gmb_set_default_ctx(gmb_ctx* ctx) {
    if (g_default_ctx && g_default_owned && g_default_ctx != ctx) gmb_ctx_destroy(g_default_ctx);
    g_default_ctx = ctx; g_default_owned = false;
    return GMB_OK;
}

// mvn_ll, src/mcml_optim.cpp:406-414
extern "C" int gmb_mvn_ll(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                          const double* gamma, int n_gamma, const double* u, int Q, int m, double* out) {
    if (!cov || !data || !gamma || !u || !out) return gmb_set_error(GMB_EINVAL, "gmb_mvn_ll: NULL argument");
    gmb_ctx* ctx; GMB_TRY(default_ctx(&ctx));
    Handles h;
    GMB_TRY(h.make_cov(ctx, cov, cov_rows, data, n_data, eff_range, n_eff));
    int B, Qc, R;
    GMB_TRY(gmb_cov_dims(h.cv, &B, &Qc, &R));
    if (n_gamma < R) return gmb_set_error(GMB_EINVAL, "gamma has %d values, the covariance has %d parameters", n_gamma, R);
    return gmb_cov_mvn_ll(h.cv, gamma, u, Q, m, m, out);
}

// mcmc_sample, src/mcml_full.cpp:314-338
extern "C" int gmb_mcmc_sample(const double* Z, const double* L, const double* X, const double* y, const double* beta,
                               int n, int P, int Q, const char* family, const char* link, int warmup, int nsamp, double lambda,
                               double var_par, int trace, int refresh, int maxsteps, double target_accept,
                               int n_chains, uint64_t seed, double* samples_out) {
    (void)trace; (void)refresh;
    if (!Z || !L || !X || !y || !beta || !samples_out) return gmb_set_error(GMB_EINVAL, "gmb_mcmc_sample: NULL argument");
    if (nsamp < 0) return gmb_set_error(GMB_EINVAL, "nsamp must be >= 0");
    gmb_ctx* ctx; GMB_TRY(default_ctx(&ctx));
    GmbPhase ph(ctx->stream);
    Handles h;
    GMB_TRY(h.make_model(ctx, n, P, Q, X, Z, y, family, link));
    ph.mark("mcmc_sample: model upload");
    const int want = nsamp + 1;                                  // Q x (nsamp + 1), mhmcmc.h:126
    int C = n_chains > 0 ? n_chains : default_chains(nsamp);
    if (C > want) C = want;
    const int per = (want + C - 1) / C;                          // columns per chain (k + 1)
    if ((size_t)C * per == (size_t)want) {                       // the chains' columns fill the result exactly: straight into the caller's buffer
        GMB_TRY(gmb_hmc_sample(h.mdl, L, beta, var_par, warmup, per - 1, lambda, maxsteps, target_accept, 100, C, 0u, seed, 0,
                               samples_out, nullptr, nullptr));
        return GMB_OK;
    }
    std::vector<double> U((size_t)Q * C * per);
    GMB_TRY(gmb_hmc_sample(h.mdl, L, beta, var_par, warmup, per - 1, lambda, maxsteps, target_accept, 100, C, 0u, seed, 0,
                           U.data(), nullptr, nullptr));
    memcpy(samples_out, U.data(), sizeof(double) * (size_t)Q * want);
    return GMB_OK;
}

// mcml_optim, src/mcml_optim.cpp:35-68
extern "C" int gmb_mcml_optim(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                              const double* Z, const double* X, const double* y, const double* u, int n, int P, int Q, int m,
                              const char* family, const char* link, const double* start, int n_start, int trace, int mcnr,
                              double* beta_out, double* theta_out, double* sigma_out) {
    (void)trace;
    if (!start) return gmb_set_error(GMB_EINVAL, "start is NULL");
    gmb_ctx* ctx; GMB_TRY(default_ctx(&ctx));
    GmbPhase ph(ctx->stream);
    Handles h;
    GMB_TRY(setup_fixed_u(ctx, cov, cov_rows, data, n_data, eff_range, n_eff, Z, X, y, u, n, P, Q, m, family, link, h));
    ph.mark("optim: objects, upload u, zd");
    Fit mc;
    GMB_TRY(mc.init(ctx, h.cv, h.mdl, start, n_start, family));
    mc.model_var_par = 1.0;                                       // :52
    mc.d_cols_total = h.m_total;
    if (!mcnr) GMB_TRY(mc.l_optim()); else GMB_TRY(mc.mcnr());    // :55-59
    ph.mark("optim: beta step");
    GMB_TRY(mc.d_optim());                                        // :60
    ph.mark("optim: theta step");
    if (beta_out) memcpy(beta_out, mc.beta.data(), sizeof(double) * P);
    if (theta_out) memcpy(theta_out, mc.theta.data(), sizeof(double) * mc.R);
    if (sigma_out) *sigma_out = mc.sigma;
    return GMB_OK;
}

// mcml_simlik, src/mcml_optim.cpp:90-117
extern "C" int gmb_mcml_simlik(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                               const double* Z, const double* X, const double* y, const double* u, int n, int P, int Q, int m,
                               const char* family, const char* link, const double* start, int n_start, int trace,
                               double* beta_out, double* theta_out, double* sigma_out) {
    (void)trace;
    if (!start) return gmb_set_error(GMB_EINVAL, "start is NULL");
    gmb_ctx* ctx; GMB_TRY(default_ctx(&ctx));
    Handles h;
    GMB_TRY(setup_fixed_u(ctx, cov, cov_rows, data, n_data, eff_range, n_eff, Z, X, y, u, n, P, Q, m, family, link, h));
    Fit mc;
    GMB_TRY(mc.init(ctx, h.cv, h.mdl, start, n_start, family));
    mc.d_cols_total = h.m_total;
    GMB_TRY(mc.f_optim());                                        // :108
    if (beta_out) memcpy(beta_out, mc.beta.data(), sizeof(double) * P);
    if (theta_out) memcpy(theta_out, mc.theta.data(), sizeof(double) * mc.R);
    if (sigma_out) *sigma_out = mc.sigma;
    return GMB_OK;
}

// mcml_hess, src/mcml_optim.cpp:263-285
extern "C" int gmb_mcml_hess(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                             const double* Z, const double* X, const double* y, const double* u, int n, int P, int Q, int m,
                             const char* family, const char* link, const double* start, int n_start, double tol, int trace,
                             double* hess_out) {
    (void)trace;
    if (!start || !hess_out) return gmb_set_error(GMB_EINVAL, "start or hess_out is NULL");
    if (!(tol > 0.0)) return gmb_set_error(GMB_EINVAL, "tol must be > 0");
    gmb_ctx* ctx; GMB_TRY(default_ctx(&ctx));
    Handles h;
    GMB_TRY(setup_fixed_u(ctx, cov, cov_rows, data, n_data, eff_range, n_eff, Z, X, y, u, n, P, Q, m, family, link, h));
    Fit mc;
    GMB_TRY(mc.init(ctx, h.cv, h.mdl, start, n_start, family));
    mc.d_cols_total = h.m_total;
    return mc.f_hess(tol, hess_out);                              // :283
}

// aic_mcml, src/mcml_optim.cpp:356-392
extern "C" int gmb_aic_mcml(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                            const double* Z, const double* X, const double* y, const double* u, int n, int P, int Q, int m,
                            const char* family, const char* link, const double* beta_par, int n_beta_par,
                            const double* cov_par, int n_cov_par, double* out) {
    if (!beta_par || !cov_par || !out) return gmb_set_error(GMB_EINVAL, "gmb_aic_mcml: NULL argument");
    gmb_ctx* ctx; GMB_TRY(default_ctx(&ctx));
    const std::string fam(family ? family : "");
    const bool has_var = fam == "gaussian" || fam == "Gamma" || fam == "beta";     // :374
    if (n_beta_par < P + (has_var ? 1 : 0)) return gmb_set_error(GMB_EINVAL, "beta_par has %d values, needs %d", n_beta_par, P + (has_var ? 1 : 0));
    Handles h;
    GMB_TRY(setup_fixed_u(ctx, cov, cov_rows, data, n_data, eff_range, n_eff, Z, X, y, u, n, P, Q, m, family, link, h));
    int B, Qc, R;
    GMB_TRY(gmb_cov_dims(h.cv, &B, &Qc, &R));
    if (n_cov_par < R) return gmb_set_error(GMB_EINVAL, "cov_par has %d values, the covariance has %d parameters", n_cov_par, R);
    const double var_par = has_var ? beta_par[P] : 0.0;            // :375-382
    const int dof = n_beta_par + n_cov_par;                        // :371
    double dmvvec, ll;
    GMB_TRY(gmb_cov_mvn_ll_model(h.cv, cov_par, h.mdl, h.m_total, &dmvvec));   // :386
    GMB_TRY(gmb_model_loglik(h.mdl, beta_par, var_par, &ll));          // :387
    *out = -2 * (ll + dmvvec) + 2 * dof;                           // :389
    return GMB_OK;
}

// mcml_full, src/mcml_full.cpp:41-148
extern "C" int gmb_mcml_full(const int32_t* cov, int cov_rows, const double* data, int n_data, const double* eff_range, int n_eff,
                             const double* Z, const double* X, const double* y, int n, int P, int Q,
                             const char* family, const char* link, const double* start, int n_start,
                             int mcnr, int m, int maxiter, int warmup, double tol, int verbose, double lambda, int trace,
                             int refresh, int maxsteps, double target_accept, int n_chains, uint64_t seed,
                             double* beta_out, double* theta_out, double* sigma_out, int* converged_out, int* iter_out,
                             double* u_out) {
    (void)trace; (void)refresh;
    if (!start) return gmb_set_error(GMB_EINVAL, "start is NULL");
    if (m <= 0) return gmb_set_error(GMB_EINVAL, "m must be positive");
    GMB_TRY(check_common(cov, data, Z, X, y, n, P, Q));
    gmb_ctx* ctx; GMB_TRY(default_ctx(&ctx));
    Handles h;
    GMB_TRY(h.make_cov(ctx, cov, cov_rows, data, n_data, eff_range, n_eff));
    int B, Qc, R;
    GMB_TRY(gmb_cov_dims(h.cv, &B, &Qc, &R));
    if (Qc != Q) return gmb_set_error(GMB_EINVAL, "covariance has %d random effects, Z has %d columns", Qc, Q);
    if (n_start < P + R + 1) return gmb_set_error(GMB_EINVAL, "start needs at least P + R + 1 = %d values (src/mcml_full.cpp:73,110)", P + R + 1);
    GMB_TRY(h.make_model(ctx, n, P, Q, X, Z, y, family, link));
    const std::string fam(family ? family : "");
    std::vector<double> theta(start + P, start + P + R), beta(start, start + P);                   // :63-64
    double var_par = (fam == "gaussian" || fam == "Gamma") ? start[n_start - 1] : 1.0;             // :65
    Fit mc;
    GMB_TRY(mc.init(ctx, h.cv, h.mdl, start, n_start, family));                                    // :71
    int world = ctx->world, rank = ctx->rank;
    // chains: n_chains counts chains over all ranks; each yields `per` columns (column 0 = state after warm-up)
    const int want = m + 1;
    int C_total = n_chains > 0 ? n_chains : default_chains(m);
    if (C_total > want) C_total = want;
    int C_local = (C_total + world - 1) / world;
    C_total = C_local * world;
    const int per = (want + C_total - 1) / C_total;
    const bool single_chain = (C_total == 1);
    std::vector<double> Lhost((size_t)Q * Q);
    double maxdiff = 1.0; int iter = 1; bool converged = false;
    std::vector<double> newbeta(P), newtheta(R);
    double new_var_par = 1.0;                                                                      // :79
    GMB_TRY(gmb_cov_gen(h.cv, theta.data(), 1, Lhost.data()));                                     // :68
    while (maxdiff > tol && iter <= maxiter) {                                                     // :83
        // u = mcmc.sample(warmup, m)  (:92) — chain restarted and re-adapted every iteration (mhmcmc.h:127)
        gmb_hmc_stats st;
        GMB_TRY(gmb_hmc_sample(h.mdl, Lhost.data(), beta.data(), var_par, warmup, per - 1, lambda, maxsteps, target_accept, 100,
                               C_local, (uint32_t)(rank * C_local), seed + (uint64_t)iter * 0x9E3779B97F4A7C15ull, 1, nullptr, nullptr, &st));
        // niter_ stays m while u has m + 1 columns (App. B #1) for the reference's single chain; with several chains
        // every generated column is used by both steps
        const int tot = C_total * per;
        GMB_TRY(gmb_model_use_device_u(h.mdl, single_chain ? m : tot));
        mc.d_cols_total = tot;
        mc.model_var_par = var_par;
        mc.new_samples();
        if (!mcnr) GMB_TRY(mc.l_optim()); else GMB_TRY(mc.mcnr());                                 // :95-99
        GMB_TRY(mc.d_optim());                                                                     // :101
        newbeta = mc.beta; newtheta = mc.theta;                                                    // :103-104
        if (fam == "gaussian" || fam == "Gamma") new_var_par = mc.sigma;                           // :105
        // every rank holds the same all-reduced sums, hence the same estimates; the broadcast pins that down
        std::vector<double> pack(newbeta); pack.insert(pack.end(), newtheta.begin(), newtheta.end()); pack.push_back(new_var_par);
        GMB_TRY(gmb_comm_bcast_host(ctx, pack.data(), (int)pack.size()));
        for (int p = 0; p < P; p++) newbeta[p] = pack[p];
        for (int r = 0; r < R; r++) newtheta[r] = pack[P + r];
        new_var_par = pack[P + R];
        mc.beta = newbeta; mc.theta = newtheta;
        maxdiff = 0.0;                                                                             // :108-111
        for (int p = 0; p < P; p++) maxdiff = std::max(maxdiff, std::fabs(beta[p] - newbeta[p]));
        for (int r = 0; r < R; r++) maxdiff = std::max(maxdiff, std::fabs(theta[r] - newtheta[r]));
        maxdiff = std::max(maxdiff, std::fabs(var_par - new_var_par));
        if (maxdiff < tol) converged = true;                                                       // :113
        beta = newbeta; theta = newtheta; var_par = new_var_par;                                   // :116-118
        if (!converged) GMB_TRY(gmb_cov_gen(h.cv, theta.data(), 1, Lhost.data()));                 // :119-126
        if (verbose && rank == 0) {
            fprintf(stderr, "Iter %d  beta:", iter);
            for (int p = 0; p < P; p++) fprintf(stderr, " %.5f", beta[p]);
            fprintf(stderr, "  theta:");
            for (int r = 0; r < R; r++) fprintf(stderr, " %.5f", theta[r]);
            fprintf(stderr, "  sigma: %.5f  max diff: %.3g  accept: %.3f  eps: %.4f%s\n", var_par, maxdiff, st.accept_rate, st.step_size_mean,
                    converged ? "  CONVERGED" : "");
        }
        iter++;
    }
    if (beta_out) memcpy(beta_out, beta.data(), sizeof(double) * P);
    if (theta_out) memcpy(theta_out, theta.data(), sizeof(double) * R);
    if (sigma_out) *sigma_out = var_par;
    if (converged_out) *converged_out = converged ? 1 : 0;
    if (iter_out) *iter_out = iter - 1;
    if (u_out) {
        // the first m + 1 columns of this rank's samples (Q x (m+1), :144-145); zero-filled when the rank holds fewer
        const int have = std::min(want, h.mdl->m_local);
        memset(u_out, 0, sizeof(double) * (size_t)Q * want);
        GMB_CUDA(cudaMemcpy2D(u_out, Q * sizeof(double), h.mdl->dU, h.mdl->ldq * sizeof(double), Q * sizeof(double), have, cudaMemcpyDeviceToHost));
    }
    return GMB_OK;
}
