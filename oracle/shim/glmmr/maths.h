// oracle/shim/glmmr/maths.h — STAND-IN for glmmrBase's <glmmr/maths.h> (TEST INFRASTRUCTURE).
// glmmrBase is an un-vendored dependency of the reference (DESCRIPTION:19); these are the reconstructions of
// SURVEY.md App. C.3 — UNVERIFIED against the glmmrBase sources, which are not available offline.
#pragma once
#include <RcppEigen.h>
#include <string>

namespace glmmr {
namespace maths {

inline Eigen::VectorXd exp_vec(const Eigen::VectorXd& x) { return x.exp(); }

inline Eigen::VectorXd mod_inv_func(const Eigen::VectorXd& eta, const std::string& link) {
    Eigen::VectorXd mu(eta.size());
    for (int i = 0; i < eta.size(); i++) {
        if (link == "logit") mu(i) = std::exp(eta(i)) / (1 + std::exp(eta(i)));
        else if (link == "log") mu(i) = std::exp(eta(i));
        else if (link == "probit") mu(i) = R::pnorm(eta(i), 0, 1, true, false);
        else if (link == "inverse") mu(i) = 1 / eta(i);
        else mu(i) = eta(i);
    }
    return mu;
}

inline Eigen::ArrayXd gaussian_pdf_vec(const Eigen::VectorXd& x) {
    Eigen::ArrayXd o(x.size());
    for (int i = 0; i < x.size(); i++) o(i) = R::dnorm(x(i), 0, 1, false);
    return o;
}

// reciprocal IRLS weight without the dispersion: V(mu) (d eta / d mu)^2
inline Eigen::VectorXd dhdmu(const Eigen::VectorXd& eta, const std::string& family, const std::string& link) {
    Eigen::VectorXd w(eta.size());
    for (int i = 0; i < eta.size(); i++) {
        const double e = eta(i);
        if (family == "poisson" && link == "log") w(i) = std::exp(-e);
        else if (family == "poisson" && link == "identity") w(i) = e;
        else if (family == "binomial" && link == "logit") { double p = std::exp(e) / (1 + std::exp(e)); w(i) = 1 / (p * (1 - p)); }
        else if (family == "binomial" && link == "log") { double p = std::exp(e); w(i) = (1 - p) / p; }
        else if (family == "binomial" && link == "identity") w(i) = e * (1 - e);
        else w(i) = 1.0;
    }
    return w;
}

}  // namespace maths
namespace algo {
inline double inner_sum(const double* a, const double* b, int n) { double s = 0; for (int i = 0; i < n; i++) s += a[i] * b[i]; return s; }
}
}  // namespace glmmr
