// adapters_driver.cpp — TEST INFRASTRUCTURE: calls the Rcpp adapters of src/ the way src/RcppExports.cpp does (Eigen / Rcpp objects in,
// Rcpp::List / Eigen objects out), compiled against the stand-in Rcpp / RcppEigen headers of oracle/shim (R is not installed here).
// Each drv_* function converts plain buffers to the adapter's argument types (the job of Rcpp::traits::input_parameter), calls the
// adapter, unpacks the result (the job of Rcpp::wrap) and maps an R error (Rcpp::stop -> exception, BEGIN_RCPP / END_RCPP) to rc = 1 + text.
#include <RcppEigen.h>
#include <cstring>
#include <string>

Rcpp::List mcml_full(const Eigen::ArrayXXi&, const Eigen::ArrayXd&, const Eigen::ArrayXd&, const Eigen::MatrixXd&, const Eigen::MatrixXd&, const Eigen::VectorXd&,
                     std::string, std::string, Eigen::ArrayXd, bool, int, int, int, double, bool, double, int, int, int, double);
Eigen::ArrayXXd mcmc_sample(const Eigen::MatrixXd&, const Eigen::MatrixXd&, const Eigen::MatrixXd&, const Eigen::VectorXd&, const Eigen::VectorXd&, std::string,
                            std::string, int, int, double, double, int, int, int, double);
Rcpp::List mcml_la(const Eigen::ArrayXXi&, const Eigen::ArrayXd&, const Eigen::ArrayXd&, const Eigen::MatrixXd&, const Eigen::MatrixXd&, const Eigen::VectorXd&,
                   std::string, std::string, Eigen::ArrayXd, bool, double, bool, int, int);
Rcpp::List mcml_la_nr(const Eigen::ArrayXXi&, const Eigen::ArrayXd&, const Eigen::ArrayXd&, const Eigen::MatrixXd&, const Eigen::MatrixXd&, const Eigen::VectorXd&,
                      std::string, std::string, Eigen::ArrayXd, bool, double, bool, int, int);
Rcpp::List mcml_optim(const Eigen::ArrayXXi&, const Eigen::ArrayXd&, const Eigen::ArrayXd&, const Eigen::MatrixXd&, const Eigen::MatrixXd&, const Eigen::VectorXd&,
                      Eigen::MatrixXd, std::string, std::string, Eigen::ArrayXd, int, bool);
Rcpp::List mcml_simlik(const Eigen::ArrayXXi&, const Eigen::ArrayXd&, const Eigen::ArrayXd&, const Eigen::MatrixXd&, const Eigen::MatrixXd&, const Eigen::VectorXd&,
                       Eigen::MatrixXd, std::string, std::string, Eigen::ArrayXd, int);
Eigen::MatrixXd mcml_hess(const Eigen::ArrayXXi&, const Eigen::ArrayXd&, const Eigen::ArrayXd&, const Eigen::MatrixXd&, const Eigen::MatrixXd&, const Eigen::VectorXd&,
                          Eigen::MatrixXd, std::string, std::string, Eigen::ArrayXd, double, int);
double aic_mcml(const Eigen::ArrayXXi&, const Eigen::ArrayXd&, const Eigen::ArrayXd&, const Eigen::MatrixXd&, const Eigen::MatrixXd&, const Eigen::VectorXd&,
                Eigen::MatrixXd, std::string, std::string, const Eigen::VectorXd&, const Eigen::VectorXd&);
double mvn_ll(const Eigen::ArrayXXi&, const Eigen::ArrayXd&, const Eigen::ArrayXd&, const Eigen::ArrayXd&, const Eigen::MatrixXd&);

namespace {
thread_local std::string g_msg;
template <class M> M mat(const double* p, int r, int c) { M m(r, c); std::memcpy(m.data(), p, sizeof(double) * (size_t)r * c); return m; }
template <class V> V vec(const double* p, int n) { V v(n); std::memcpy(v.data(), p, sizeof(double) * (size_t)n); return v; }
Eigen::ArrayXXi imat(const int* p, int r, int c) { Eigen::ArrayXXi m(r, c); std::memcpy(m.data(), p, sizeof(int) * (size_t)r * c); return m; }
void put(const Rcpp::List& l, const char* name, double* out) { const Rcpp::ListEntry& e = l[name]; std::memcpy(out, e.v.data(), sizeof(double) * e.v.size()); }
struct Common {
    Eigen::ArrayXXi cov; Eigen::ArrayXd data, eff; Eigen::MatrixXd Z, X; Eigen::VectorXd y;
    Common(const int* c, int rows, const double* d, int nd, const double* e, int ne, const double* Zp, const double* Xp, const double* yp, int n, int P, int Q)
        : cov(imat(c, rows, 5)), data(vec<Eigen::ArrayXd>(d, nd)), eff(vec<Eigen::ArrayXd>(e, ne)), Z(mat<Eigen::MatrixXd>(Zp, n, Q)), X(mat<Eigen::MatrixXd>(Xp, n, P)),
          y(vec<Eigen::VectorXd>(yp, n)) {}
};
}  // namespace

#define DRV_TRY try {
#define DRV_CATCH } catch (const std::exception& ex) { g_msg = ex.what(); return 1; } catch (...) { g_msg = "unknown C++ exception"; return 1; } return 0;
#define DRV extern "C" __attribute__((visibility("default")))

DRV const char* drv_last_error() { return g_msg.c_str(); }
DRV void drv_set_seed(unsigned long long s) { Rcpp::unif_state() = s; }

DRV int drv_mvn_ll(const int* cov, int rows, const double* data, int nd, const double* eff, int ne, const double* gamma, int ng, const double* u, int Q, int m, double* out) {
    DRV_TRY
    *out = mvn_ll(imat(cov, rows, 5), vec<Eigen::ArrayXd>(data, nd), vec<Eigen::ArrayXd>(eff, ne), vec<Eigen::ArrayXd>(gamma, ng), mat<Eigen::MatrixXd>(u, Q, m));
    DRV_CATCH
}

DRV int drv_mcmc_sample(const double* Z, const double* L, const double* X, const double* y, const double* beta, int n, int P, int Q, const char* family, const char* link,
                        int warmup, int nsamp, double lambda, double var_par, int trace, int refresh, int maxsteps, double target_accept, double* out) {
    DRV_TRY
    Eigen::ArrayXXd s = mcmc_sample(mat<Eigen::MatrixXd>(Z, n, Q), mat<Eigen::MatrixXd>(L, Q, Q), mat<Eigen::MatrixXd>(X, n, P), vec<Eigen::VectorXd>(y, n),
                                    vec<Eigen::VectorXd>(beta, P), family, link, warmup, nsamp, lambda, var_par, trace, refresh, maxsteps, target_accept);
    if (s.rows() != Q || s.cols() != nsamp + 1) { g_msg = "mcmc_sample: unexpected shape"; return 2; }
    std::memcpy(out, s.data(), sizeof(double) * (size_t)Q * (nsamp + 1));
    DRV_CATCH
}

DRV int drv_mcml_optim(const int* cov, int rows, const double* data, int nd, const double* eff, int ne, const double* Z, const double* X, const double* y, const double* u,
                       int n, int P, int Q, int m, const char* family, const char* link, const double* start, int ns, int trace, int mcnr, int simlik,
                       double* beta, double* theta, double* sigma) {
    DRV_TRY
    Common c(cov, rows, data, nd, eff, ne, Z, X, y, n, P, Q);
    Rcpp::List l = simlik ? mcml_simlik(c.cov, c.data, c.eff, c.Z, c.X, c.y, mat<Eigen::MatrixXd>(u, Q, m), family, link, vec<Eigen::ArrayXd>(start, ns), trace)
                          : mcml_optim(c.cov, c.data, c.eff, c.Z, c.X, c.y, mat<Eigen::MatrixXd>(u, Q, m), family, link, vec<Eigen::ArrayXd>(start, ns), trace, mcnr != 0);
    put(l, "beta", beta); put(l, "theta", theta); put(l, "sigma", sigma);
    DRV_CATCH
}

DRV int drv_mcml_hess(const int* cov, int rows, const double* data, int nd, const double* eff, int ne, const double* Z, const double* X, const double* y, const double* u,
                      int n, int P, int Q, int m, const char* family, const char* link, const double* start, int ns, double tol, int trace, double* hess, int k) {
    DRV_TRY
    Common c(cov, rows, data, nd, eff, ne, Z, X, y, n, P, Q);
    Eigen::MatrixXd H = mcml_hess(c.cov, c.data, c.eff, c.Z, c.X, c.y, mat<Eigen::MatrixXd>(u, Q, m), family, link, vec<Eigen::ArrayXd>(start, ns), tol, trace);
    if (H.rows() != k || H.cols() != k) { g_msg = "mcml_hess: unexpected shape"; return 2; }
    std::memcpy(hess, H.data(), sizeof(double) * (size_t)k * k);
    DRV_CATCH
}

DRV int drv_aic_mcml(const int* cov, int rows, const double* data, int nd, const double* eff, int ne, const double* Z, const double* X, const double* y, const double* u,
                     int n, int P, int Q, int m, const char* family, const char* link, const double* bp, int nbp, const double* cp, int ncp, double* out) {
    DRV_TRY
    Common c(cov, rows, data, nd, eff, ne, Z, X, y, n, P, Q);
    *out = aic_mcml(c.cov, c.data, c.eff, c.Z, c.X, c.y, mat<Eigen::MatrixXd>(u, Q, m), family, link, vec<Eigen::VectorXd>(bp, nbp), vec<Eigen::VectorXd>(cp, ncp));
    DRV_CATCH
}

DRV int drv_mcml_full(const int* cov, int rows, const double* data, int nd, const double* eff, int ne, const double* Z, const double* X, const double* y,
                      int n, int P, int Q, const char* family, const char* link, const double* start, int ns, int mcnr, int m, int maxiter, int warmup, double tol,
                      double lambda, int maxsteps, double target_accept, double* beta, double* theta, double* sigma, int* converged, double* u) {
    DRV_TRY
    Common c(cov, rows, data, nd, eff, ne, Z, X, y, n, P, Q);
    Rcpp::List l = mcml_full(c.cov, c.data, c.eff, c.Z, c.X, c.y, family, link, vec<Eigen::ArrayXd>(start, ns), mcnr != 0, m, maxiter, warmup, tol, false, lambda, 0, 500,
                             maxsteps, target_accept);
    put(l, "beta", beta); put(l, "theta", theta); put(l, "sigma", sigma);
    *converged = l["converged"].v[0] != 0.0;
    if (l["u"].rows != Q || l["u"].cols != m + 1) { g_msg = "mcml_full: unexpected shape of u"; return 2; }
    put(l, "u", u);
    DRV_CATCH
}

DRV int drv_mcml_la(const int* cov, int rows, const double* data, int nd, const double* eff, int ne, const double* Z, const double* X, const double* y,
                    int n, int P, int Q, const char* family, const char* link, const double* start, int ns, int nr, int usehess, double tol, int maxiter,
                    double* beta, double* theta, double* sigma, double* se, double* u) {
    DRV_TRY
    Common c(cov, rows, data, nd, eff, ne, Z, X, y, n, P, Q);
    Rcpp::List l = nr ? mcml_la_nr(c.cov, c.data, c.eff, c.Z, c.X, c.y, family, link, vec<Eigen::ArrayXd>(start, ns), usehess != 0, tol, false, 0, maxiter)
                      : mcml_la(c.cov, c.data, c.eff, c.Z, c.X, c.y, family, link, vec<Eigen::ArrayXd>(start, ns), usehess != 0, tol, false, 0, maxiter);
    put(l, "beta", beta); put(l, "theta", theta); put(l, "sigma", sigma); put(l, "se", se); put(l, "u", u);
    DRV_CATCH
}
