// Rcpp adapters for the fixed-sample entry points: mcml_optim, mcml_simlik, mcml_hess, aic_mcml, mvn_ll.
// Signatures as declared in the reference's src/RcppExports.cpp:119,141,209,255,276; the bodies forward to the C-ABI.
// (mcml_optim_sparse / mcml_simlik_sparse / mcml_hess_sparse keep their reference bodies: sparse D is out of scope)
#include "gmb_adapter.h"
using namespace gmb_adapter;

#define GMB_COV(cov, data, eff) cov.data(), cov.rows(), data.data(), (int)data.size(), eff.data(), (int)eff.size()
#define GMB_FIXED_U() Z.data(), X.data(), y.data(), u.data(), (int)X.rows(), (int)X.cols(), (int)Z.cols(), (int)u.cols(), family.c_str(), link.c_str()

// [[Rcpp::export]]
Rcpp::List mcml_optim(const Eigen::ArrayXXi& cov, const Eigen::ArrayXd& data, const Eigen::ArrayXd& eff_range, const Eigen::MatrixXd& Z,
                      const Eigen::MatrixXd& X, const Eigen::VectorXd& y, Eigen::MatrixXd u, std::string family, std::string link,
                      Eigen::ArrayXd start, int trace, bool mcnr = false) {
    const Shape s = cov_shape(cov);
    Eigen::VectorXd beta((int)X.cols()), theta(s.R);
    double sigma = 0.0;
    check(gmb_mcml_optim(GMB_COV(cov, data, eff_range), GMB_FIXED_U(), start.data(), (int)start.size(), trace, mcnr ? 1 : 0, beta.data(), theta.data(), &sigma));
    return Rcpp::List::create(Rcpp::_["beta"] = beta, Rcpp::_["theta"] = theta, Rcpp::_["sigma"] = sigma);
}

// [[Rcpp::export]]
Rcpp::List mcml_simlik(const Eigen::ArrayXXi& cov, const Eigen::ArrayXd& data, const Eigen::ArrayXd& eff_range, const Eigen::MatrixXd& Z,
                       const Eigen::MatrixXd& X, const Eigen::VectorXd& y, Eigen::MatrixXd u, std::string family, std::string link,
                       Eigen::ArrayXd start, int trace) {
    const Shape s = cov_shape(cov);
    Eigen::VectorXd beta((int)X.cols()), theta(s.R);
    double sigma = 0.0;
    check(gmb_mcml_simlik(GMB_COV(cov, data, eff_range), GMB_FIXED_U(), start.data(), (int)start.size(), trace, beta.data(), theta.data(), &sigma));
    return Rcpp::List::create(Rcpp::_["beta"] = beta, Rcpp::_["theta"] = theta, Rcpp::_["sigma"] = sigma);
}

// [[Rcpp::export]]
Eigen::MatrixXd mcml_hess(const Eigen::ArrayXXi& cov, const Eigen::ArrayXd& data, const Eigen::ArrayXd& eff_range, const Eigen::MatrixXd& Z,
                          const Eigen::MatrixXd& X, const Eigen::VectorXd& y, Eigen::MatrixXd u, std::string family, std::string link,
                          Eigen::ArrayXd start, double tol = 1e-5, int trace = 0) {
    const Shape s = cov_shape(cov);
    const int k = (int)X.cols() + s.R;
    Eigen::MatrixXd hess(k, k);
    check(gmb_mcml_hess(GMB_COV(cov, data, eff_range), GMB_FIXED_U(), start.data(), (int)start.size(), tol, trace, hess.data()));
    return hess;
}

// [[Rcpp::export]]
double aic_mcml(const Eigen::ArrayXXi& cov, const Eigen::ArrayXd& data, const Eigen::ArrayXd& eff_range, const Eigen::MatrixXd& Z,
                const Eigen::MatrixXd& X, const Eigen::VectorXd& y, Eigen::MatrixXd u, std::string family, std::string link,
                const Eigen::VectorXd& beta_par, const Eigen::VectorXd& cov_par) {
    double aic = 0.0;
    check(gmb_aic_mcml(GMB_COV(cov, data, eff_range), GMB_FIXED_U(), beta_par.data(), (int)beta_par.size(), cov_par.data(), (int)cov_par.size(), &aic));
    return aic;
}

// [[Rcpp::export]]
double mvn_ll(const Eigen::ArrayXXi& cov, const Eigen::ArrayXd& data, const Eigen::ArrayXd& eff_range, const Eigen::ArrayXd& gamma,
              const Eigen::MatrixXd& u) {
    double ll = 0.0;
    check(gmb_mvn_ll(GMB_COV(cov, data, eff_range), gamma.data(), (int)gamma.size(), u.data(), (int)u.rows(), (int)u.cols(), &ll));
    return ll;
}
