// Rcpp adapters for the Laplace-approximation fits mcml_la and mcml_la_nr.
// Signatures as declared in the reference's src/RcppExports.cpp:71,95; the bodies forward to the C-ABI.
#include "gmb_adapter.h"
using namespace gmb_adapter;

namespace {
typedef int (*la_entry)(const int32_t*, int, const double*, int, const double*, int, const double*, const double*, const double*, int, int, int,
                        const char*, const char*, const double*, int, int, double, int, int, int, double*, double*, double*, double*, double*, int*);

Rcpp::List run(la_entry fn, const Eigen::ArrayXXi& cov, const Eigen::ArrayXd& data, const Eigen::ArrayXd& eff_range, const Eigen::MatrixXd& Z,
               const Eigen::MatrixXd& X, const Eigen::VectorXd& y, const std::string& family, const std::string& link, const Eigen::ArrayXd& start,
               bool usehess, double tol, bool verbose, int trace, int maxiter) {
    const Shape s = cov_shape(cov);
    const int P = (int)X.cols(), Q = (int)Z.cols();
    Eigen::VectorXd beta(P), theta(s.R), se = Eigen::VectorXd::Zero((int)start.size());
    Eigen::MatrixXd u(Q, 1);
    double sigma = 0.0;
    int iter = 0;
    check(fn(cov.data(), cov.rows(), data.data(), (int)data.size(), eff_range.data(), (int)eff_range.size(), Z.data(), X.data(), y.data(), (int)X.rows(), P, Q,
             family.c_str(), link.c_str(), start.data(), (int)start.size(), usehess ? 1 : 0, tol, verbose ? 1 : 0, trace, maxiter,
             beta.data(), theta.data(), &sigma, se.data(), u.data(), &iter));
    return Rcpp::List::create(Rcpp::_["beta"] = beta, Rcpp::_["theta"] = theta, Rcpp::_["sigma"] = sigma, Rcpp::_["se"] = se, Rcpp::_["u"] = u);
}
}  // namespace

// [[Rcpp::export]]
Rcpp::List mcml_la(const Eigen::ArrayXXi& cov, const Eigen::ArrayXd& data, const Eigen::ArrayXd& eff_range, const Eigen::MatrixXd& Z,
                   const Eigen::MatrixXd& X, const Eigen::VectorXd& y, std::string family, std::string link, Eigen::ArrayXd start,
                   bool usehess = false, double tol = 1e-3, bool verbose = true, int trace = 0, int maxiter = 10) {
    return run(gmb_mcml_la, cov, data, eff_range, Z, X, y, family, link, start, usehess, tol, verbose, trace, maxiter);
}

// [[Rcpp::export]]
Rcpp::List mcml_la_nr(const Eigen::ArrayXXi& cov, const Eigen::ArrayXd& data, const Eigen::ArrayXd& eff_range, const Eigen::MatrixXd& Z,
                      const Eigen::MatrixXd& X, const Eigen::VectorXd& y, std::string family, std::string link, Eigen::ArrayXd start,
                      bool usehess = false, double tol = 1e-3, bool verbose = true, int trace = 0, int maxiter = 10) {
    return run(gmb_mcml_la_nr, cov, data, eff_range, Z, X, y, family, link, start, usehess, tol, verbose, trace, maxiter);
}
