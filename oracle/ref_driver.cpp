// ref_driver.cpp — builds oracle/_ref/libref.so: the REFERENCE'S OWN numeric headers, compiled unmodified from where
// they lie under /root/reference/inst/include, against the stand-in headers in oracle/shim/ (TEST INFRASTRUCTURE).
//
// What this pins: the arithmetic the reference itself wrote — maths::log_likelihood, log_factorial_approx, detadmu,
// forward_sub (moremaths.h), mcmlModel::{log_likelihood, log_prob, log_grad, update_W} (mcmlmodel.h),
// MCMLDmatrix::{loglik, loglik_block, logdet} (mcmldmatrix.h), mcmloptim::mcnr (mcmloptim.h) and
// mcmcRunHMC::{initialise_u, new_proposal, sample} (mhmcmc.h) — is executed as written and compared with
// oracle/oracle.cpp (tests/test_oracle_vs_ref.py) and frozen into tests/golden/ (tests/golden/make_golden.py).
// What it does NOT pin: Eigen's own kernels (the shim evaluates the same expressions eagerly, in plain loops) and the
// un-vendored glmmrBase pieces (DData/DMatrix/dhdmu/mod_inv_func are reconstructions, oracle/shim/glmmr*.h).
//
// The reference draws normals from R's RNG and uniforms from a std::minstd_rand seeded by std::random_device
// (mhmcmc.h:48-55,62,85).  Here both are redirected — by name, without touching the header — to the counter-based
// Philox stream the device sampler and oracle.cpp use, so that chains can be compared state by state.
#include <cstdint>
#include <cstring>
#include <random>
#include <string>

namespace refrng {
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
static inline double u01(uint32_t lo, uint32_t hi) { uint64_t x = ((uint64_t)hi << 32) | lo; return ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static inline void uniform2(uint64_t seed, uint32_t idx, uint32_t iter, uint32_t chain, uint32_t stream, double* a, double* b) {
    uint32_t c[4] = {idx, iter, chain, stream};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    *a = u01(c[0], c[1]); *b = u01(c[2], c[3]);
}
static inline void normal_vec(uint64_t seed, uint32_t iter, uint32_t chain, uint32_t stream, int Q, double* z) {
    for (int p = 0; p < (Q + 1) / 2; p++) {
        double u1, u2; uniform2(seed, (uint32_t)p, iter, chain, stream, &u1, &u2);
        double r = std::sqrt(-2.0 * std::log(u1)), a = 2.0 * 3.14159265358979323846 * u2;
        z[2 * p] = r * std::cos(a);
        if (2 * p + 1 < Q) z[2 * p + 1] = r * std::sin(a);
    }
}
struct State { uint64_t seed = 0; uint32_t chain = 0; int normal_calls = 0; int uniform_calls = 0; };
static State g;
// call 0,1: constructor's initialise_u (mhmcmc.h:44) ; 2: sample()'s u_ (:127 -> :48) ; 3: its r_ (:50) ; 4+t: proposal t (:62)
static void normal_source(int n, double* out) {
    int k = g.normal_calls++;
    if (k == 2) normal_vec(g.seed, 0, g.chain, 0, n, out);
    else if (k >= 4) normal_vec(g.seed, (uint32_t)(k - 4), g.chain, 2, n, out);
    else for (int i = 0; i < n; i++) out[i] = 0.0;
}
static double next_uniform() { double a, b; uniform2(g.seed, 0, (uint32_t)g.uniform_calls++, g.chain, 3, &a, &b); return a; }
}  // namespace refrng

// redirect the three <random> names used by mhmcmc.h:27-28,55-56,85
namespace std {
struct gmb_shim_rd { unsigned operator()() { return 0u; } };
struct gmb_shim_rng { gmb_shim_rng() {} explicit gmb_shim_rng(unsigned) {} };
template <class T> struct gmb_shim_dist { gmb_shim_dist() {} gmb_shim_dist(T, T) {} T operator()(gmb_shim_rng&) { return (T)refrng::next_uniform(); } };
}
#define random_device gmb_shim_rd
#define minstd_rand gmb_shim_rng
#define uniform_real_distribution gmb_shim_dist

#include <RcppEigen.h>
#include "glmmrmcml/mcmldmatrix.h"
#include "glmmrmcml/mcmloptim.h"
#include "glmmrmcml/mhmcmc.h"

#undef random_device
#undef minstd_rand
#undef uniform_real_distribution

#define REF_API extern "C" __attribute__((visibility("default")))

static Eigen::MatrixXd mat(const double* p, int r, int c) { Eigen::MatrixXd m(r, c); std::memcpy(m.data(), p, sizeof(double) * r * c); return m; }
static Eigen::VectorXd vec(const double* p, int n) { Eigen::VectorXd v(n); std::memcpy(v.data(), p, sizeof(double) * n); return v; }
static Eigen::ArrayXd arrd(const double* p, int n) { Eigen::ArrayXd v(n); if (n) std::memcpy(v.data(), p, sizeof(double) * n); return v; }
static Eigen::ArrayXXi covmat(const int32_t* p, int rows) { Eigen::ArrayXXi m(rows, 5); for (int i = 0; i < rows * 5; i++) m.d[i] = p[i]; return m; }

REF_API const char* ref_root() { return REF_ROOT; }
REF_API double ref_family_ll(double y, double mu, double var_par, int flink) { return glmmr::maths::log_likelihood(y, mu, var_par, flink); }
REF_API double ref_log_factorial_approx(double n) { return glmmr::maths::log_factorial_approx(n); }
REF_API void ref_detadmu(const double* xb, int n, const char* link, double* out) {
    Eigen::VectorXd w = glmmr::maths::detadmu(vec(xb, n), link);
    std::memcpy(out, w.data(), sizeof(double) * n);
}
REF_API void ref_forward_sub(const double* L, const double* u, int n, double* out) {
    Eigen::MatrixXd Lm = mat(L, n, n); Eigen::VectorXd uv = vec(u, n);
    Eigen::VectorXd y = glmmr::algo::forward_sub(&Lm, &uv, n);
    std::memcpy(out, y.data(), sizeof(double) * n);
}

// mcmlModel(Z, nullptr, X, y, &u, beta, var_par, family, link).log_likelihood()  — as src/mcml_optim.cpp:52,387
REF_API double ref_loglik(int n, int P, int Q, int m, const double* X, const double* Z, const double* U, const double* y,
                          const double* beta, double var_par, const char* family, const char* link) {
    Eigen::MatrixXd Xm = mat(X, n, P), Zm = mat(Z, n, Q), u = mat(U, Q, m);
    glmmr::mcmlModel model(Zm, nullptr, Xm, vec(y, n), &u, vec(beta, P), var_par, family, link);
    return model.log_likelihood();
}

// the same model, log_likelihood() called `reps` times (what an optimiser does between two sample draws); returns the last value.
// Used by bench.py's reference arm to time one objective evaluation without the model construction.
REF_API double ref_loglik_reps(int n, int P, int Q, int m, const double* X, const double* Z, const double* U, const double* y,
                               const double* beta, double var_par, const char* family, const char* link, int reps) {
    Eigen::MatrixXd Xm = mat(X, n, P), Zm = mat(Z, n, Q), u = mat(U, Q, m);
    glmmr::mcmlModel model(Zm, nullptr, Xm, vec(y, n), &u, vec(beta, P), var_par, family, link);
    double v = 0.0;
    for (int r = 0; r < reps; r++) v = model.log_likelihood();
    return v;
}

// log_prob / log_grad with L given — as src/mcml_full.cpp:332
REF_API double ref_log_prob(int n, int P, int Q, const double* X, const double* Z, const double* L, const double* y,
                            const double* beta, double var_par, const char* family, const char* link, const double* v) {
    Eigen::MatrixXd Xm = mat(X, n, P), Zm = mat(Z, n, Q), Lm = mat(L, Q, Q), u = Eigen::MatrixXd::Zero(Q, 1);
    glmmr::mcmlModel model(Zm, &Lm, Xm, vec(y, n), &u, vec(beta, P), var_par, family, link);
    return model.log_prob(vec(v, Q));
}
REF_API void ref_log_grad(int n, int P, int Q, const double* X, const double* Z, const double* L, const double* y,
                          const double* beta, double var_par, const char* family, const char* link, const double* v, double* grad) {
    Eigen::MatrixXd Xm = mat(X, n, P), Zm = mat(Z, n, Q), Lm = mat(L, Q, Q), u = Eigen::MatrixXd::Zero(Q, 1);
    glmmr::mcmlModel model(Zm, &Lm, Xm, vec(y, n), &u, vec(beta, P), var_par, family, link);
    Eigen::VectorXd g = model.log_grad(vec(v, Q));
    std::memcpy(grad, g.data(), sizeof(double) * Q);
}

// MCMLDmatrix(&dat, gamma).loglik(u) — src/mcml_optim.cpp:411-413
REF_API double ref_mvn_loglik(const int32_t* cov, int rows, const double* data, int n_data, const double* eff, int n_eff,
                              const double* theta, int R, const double* U, int Q, int m) {
    glmmr::DData dat(covmat(cov, rows), arrd(data, n_data), arrd(eff, n_eff));
    glmmr::MCMLDmatrix dmat(&dat, arrd(theta, R));
    return dmat.loglik(mat(U, Q, m));
}
REF_API double ref_logdet(const int32_t* cov, int rows, const double* data, int n_data, const double* eff, int n_eff,
                          const double* theta, int R) {
    glmmr::DData dat(covmat(cov, rows), arrd(data, n_data), arrd(eff, n_eff));
    glmmr::MCMLDmatrix dmat(&dat, arrd(theta, R));
    return dmat.logdet();
}

// mcml_optim(..., mcnr = true) up to and including mc.mcnr() — src/mcml_optim.cpp:48-58.  Serial (the OpenMP loop of
// mcmloptim.h:210 races on the shared model; libref is built without -fopenmp).
REF_API void ref_mcnr(const int32_t* cov, int rows, const double* data, int n_data, const double* eff, int n_eff,
                      int n, int P, int Q, int m, const double* X, const double* Z, const double* U, const double* y,
                      const char* family, const char* link, const double* start, int n_start, double* beta_out, double* sigma_out) {
    glmmr::DData dat(covmat(cov, rows), arrd(data, n_data), arrd(eff, n_eff));
    Eigen::ArrayXd st = arrd(start, n_start);
    Eigen::ArrayXd thetapars = st.segment(P, dat.n_cov_pars());
    glmmr::MCMLDmatrix dmat(&dat, thetapars);
    Eigen::VectorXd beta = st.segment(0, P);
    Eigen::MatrixXd Xm = mat(X, n, P), Zm = mat(Z, n, Q), u = mat(U, Q, m);
    glmmr::mcmlModel model(Zm, nullptr, Xm, vec(y, n), &u, beta, 1, family, link);
    glmmr::mcmloptim<glmmr::MCMLDmatrix> mc(&dmat, &model, st, 0);
    mc.mcnr();
    Eigen::VectorXd b = mc.get_beta();
    std::memcpy(beta_out, b.data(), sizeof(double) * P);
    *sigma_out = mc.get_sigma();
}

// D_likelihood / L_likelihood / F_likelihood functors at one point (likelihood.h:31-110)
REF_API void ref_objectives(const int32_t* cov, int rows, const double* data, int n_data, const double* eff, int n_eff,
                            int n, int P, int Q, int m, const double* X, const double* Z, const double* U, const double* y,
                            const char* family, const char* link, const double* par, int n_par, double fix_var_par, double* out3) {
    glmmr::DData dat(covmat(cov, rows), arrd(data, n_data), arrd(eff, n_eff));
    const int R = dat.n_cov_pars();
    Eigen::ArrayXd st = arrd(par, n_par);
    Eigen::ArrayXd thetapars = st.segment(P, R);
    glmmr::MCMLDmatrix dmat(&dat, thetapars);
    Eigen::MatrixXd Xm = mat(X, n, P), Zm = mat(Z, n, Q), u = mat(U, Q, m);
    glmmr::mcmlModel model(Zm, nullptr, Xm, vec(y, n), &u, st.segment(0, P), 1, family, link);
    std::vector<double> pb(par, par + P), pt(par + P, par + P + R), pf(par, par + P + R);
    if (std::string(family) == "gaussian") pb.push_back(fix_var_par);
    glmmr::likelihood::L_likelihood ll(&model);
    glmmr::likelihood::D_likelihood<glmmr::MCMLDmatrix> dl(&dmat, &u);
    glmmr::likelihood::F_likelihood<glmmr::MCMLDmatrix> fl(&dmat, &model, thetapars, false, true, fix_var_par);
    out3[0] = ll(pb); out3[1] = dl(pt); out3[2] = fl(pf);
}

// Laplace path: the three functors of likelihood.h:112-230 at one state and one mcmloptim::mcnr_b step (mcmloptim.h:238-293), on a model
// built as in src/mcml_la.cpp:42-52 with theta as the initial covariance parameters (so D_ = D(theta)) and u = v.
// w_use_l selects update_W(0, true) (mcml_la_nr) or update_W() (mcml_la).  out3 = (LA_likelihood, LA_likelihood_cov, LA_likelihood_btheta).
REF_API void ref_la_objectives(const int32_t* cov, int rows, const double* data, int n_data, const double* eff, int n_eff,
                               int n, int P, int Q, const double* X, const double* Z, const double* y, const char* family, const char* link,
                               const double* beta, const double* theta, int R, const double* v, double sigma, int w_use_l,
                               double* out3, double* beta_nr, double* v_nr, double* sigma_nr) {
    glmmr::DData dat(covmat(cov, rows), arrd(data, n_data), arrd(eff, n_eff));
    const bool gaussian = std::string(family) == "gaussian";
    Eigen::ArrayXd st(P + R + 1);
    for (int i = 0; i < P; i++) st(i) = beta[i];
    for (int i = 0; i < R; i++) st(P + i) = theta[i];
    st(P + R) = sigma;
    glmmr::MCMLDmatrix dmat(&dat, arrd(theta, R));
    Eigen::MatrixXd Xm = mat(X, n, P), Zm = mat(Z, n, Q), u = mat(v, Q, 1);
    Eigen::MatrixXd L = dmat.genD(0, true, false);
    glmmr::mcmlModel model(Zm, &L, Xm, vec(y, n), &u, vec(beta, P), gaussian ? sigma : 1.0, family, link);
    glmmr::mcmloptim<glmmr::MCMLDmatrix> mc(&dmat, &model, st, 0);
    model.update_W(0, w_use_l != 0);
    std::vector<double> pbv(beta, beta + P), pt(theta, theta + R), pbt(beta, beta + P);
    pbv.insert(pbv.end(), v, v + Q);
    pbt.insert(pbt.end(), theta, theta + R);
    if (gaussian) { pt.push_back(sigma); pbt.push_back(sigma); }
    glmmr::likelihood::LA_likelihood<glmmr::MCMLDmatrix> la(&model, &dmat);
    glmmr::likelihood::LA_likelihood_cov<glmmr::MCMLDmatrix> lc(&model, &dmat);
    glmmr::likelihood::LA_likelihood_btheta<glmmr::MCMLDmatrix> lb(&model, &dmat);
    out3[0] = la(pbv);
    out3[1] = lc(pt);
    out3[2] = lb(pbt);
    if (beta_nr && v_nr) {
        for (int q = 0; q < Q; q++) u(q, 0) = v[q];
        model.update_beta(vec(beta, P));
        model.update_W(0, w_use_l != 0);
        mc.mcnr_b();
        Eigen::VectorXd b = mc.get_beta();
        std::memcpy(beta_nr, b.data(), sizeof(double) * P);
        for (int q = 0; q < Q; q++) v_nr[q] = u(q, 0);
        if (sigma_nr) *sigma_nr = mc.get_sigma();
    }
}

// mcmc_sample(Z, L, X, y, beta, family, link, warmup, nsamp, lambda, var_par, 0, refresh, maxsteps, target_accept)
// — src/mcml_full.cpp:329-337 — driven by the Philox stream (seed, chain).  out is Q x (nsamp+1).
// stats: accept_/(warmup+nsamp), e_, ebar_, steps_ of the last proposal.
REF_API void ref_mcmc_sample(int n, int P, int Q, const double* X, const double* Z, const double* L, const double* y,
                             const double* beta, const char* family, const char* link, int warmup, int nsamp, double lambda,
                             double var_par, int maxsteps, double target_accept, uint64_t seed, uint32_t chain, double* out, double* stats) {
    refrng::g = refrng::State(); refrng::g.seed = seed; refrng::g.chain = chain;
    Rcpp::normal_source() = refrng::normal_source;
    Eigen::MatrixXd Xm = mat(X, n, P), Zm = mat(Z, n, Q), L_ = mat(L, Q, Q), u = Eigen::MatrixXd::Zero(Q, nsamp);
    glmmr::mcmlModel model(Zm, &L_, Xm, vec(y, n), &u, vec(beta, P), var_par, family, link);
    glmmr::mcmc::mcmcRunHMC mcmc(&model, 0, lambda, 500, maxsteps, target_accept);
    Eigen::ArrayXXd samples = mcmc.sample(warmup, nsamp);
    std::memcpy(out, samples.data(), sizeof(double) * Q * (nsamp + 1));
    if (stats) { stats[0] = (double)mcmc.accept_ / (warmup + nsamp); stats[1] = mcmc.e_; stats[2] = mcmc.ebar_; stats[3] = mcmc.steps_; }
}
