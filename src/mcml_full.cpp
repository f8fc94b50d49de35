// Rcpp adapters for the sampling entry points: mcml_full (whole MCML loop, samples stay on the device) and mcmc_sample.
// Signatures as declared in the reference's src/RcppExports.cpp:16,46; the bodies forward to the C-ABI.
#include "gmb_adapter.h"
using namespace gmb_adapter;

// [[Rcpp::export]]
Rcpp::List mcml_full(const Eigen::ArrayXXi& cov, const Eigen::ArrayXd& data, const Eigen::ArrayXd& eff_range, const Eigen::MatrixXd& Z,
                     const Eigen::MatrixXd& X, const Eigen::VectorXd& y, std::string family, std::string link, Eigen::ArrayXd start,
                     bool mcnr = false, int m = 500, int maxiter = 30, int warmup = 500, double tol = 1e-3, bool verbose = true,
                     double lambda = 0.05, int trace = 0, int refresh = 500, int maxsteps = 100, double target_accept = 0.9) {
    const Shape s = cov_shape(cov);
    const int P = (int)X.cols(), Q = (int)Z.cols();
    Eigen::VectorXd beta(P), theta(s.R);
    Eigen::MatrixXd u(Q, m + 1);                 // the reference returns Q x (m + 1) (mhmcmc.h:126,155)
    double sigma = 0.0;
    int converged = 0, iter = 0;
    check(gmb_mcml_full(cov.data(), cov.rows(), data.data(), (int)data.size(), eff_range.data(), (int)eff_range.size(), Z.data(), X.data(), y.data(),
                        (int)X.rows(), P, Q, family.c_str(), link.c_str(), start.data(), (int)start.size(), mcnr ? 1 : 0, m, maxiter, warmup, tol,
                        verbose ? 1 : 0, lambda, trace, refresh, maxsteps, target_accept, default_chains(), seed_from_r(),
                        beta.data(), theta.data(), &sigma, &converged, &iter, u.data()));
    return Rcpp::List::create(Rcpp::_["beta"] = beta, Rcpp::_["theta"] = theta, Rcpp::_["sigma"] = sigma, Rcpp::_["converged"] = (converged != 0),
                              Rcpp::_["u"] = u);
}

// [[Rcpp::export]]
Eigen::ArrayXXd mcmc_sample(const Eigen::MatrixXd& Z, const Eigen::MatrixXd& L, const Eigen::MatrixXd& X, const Eigen::VectorXd& y,
                            const Eigen::VectorXd& beta, std::string family, std::string link, int warmup, int nsamp, double lambda,
                            double var_par = 1, int trace = 0, int refresh = 500, int maxsteps = 100, double target_accept = 0.9) {
    Eigen::ArrayXXd samples((int)Z.cols(), nsamp + 1);
    check(gmb_mcmc_sample(Z.data(), L.data(), X.data(), y.data(), beta.data(), (int)X.rows(), (int)X.cols(), (int)Z.cols(), family.c_str(), link.c_str(),
                          warmup, nsamp, lambda, var_par, trace, refresh, maxsteps, target_accept, default_chains(), seed_from_r(), samples.data()));
    return samples;
}
