// oracle/shim/Rcpp.h — STAND-IN for <Rcpp.h> and the R API pieces the reference headers touch (TEST INFRASTRUCTURE).
#pragma once
#include <cmath>
#include <limits>
#include <ostream>
#include <string>
#include <vector>

#define R_NegInf (-std::numeric_limits<double>::infinity())
#define R_PosInf (std::numeric_limits<double>::infinity())

namespace R {
inline double pnorm(double x, double mu, double sd, bool lower, bool logp) {
    double z = (x - mu) / sd;
    double p = lower ? 0.5 * std::erfc(-z / std::sqrt(2.0)) : 0.5 * std::erfc(z / std::sqrt(2.0));
    return logp ? std::log(p) : p;
}
inline double dnorm(double x, double mu, double sd, bool logp) {
    double z = (x - mu) / sd;
    double l = -0.5 * z * z - std::log(sd) - 0.5 * std::log(2 * 3.14159265358979323846);
    return logp ? l : std::exp(l);
}
}  // namespace R

namespace Rcpp {
struct NumericVector { std::vector<double> v; };
// normal draws come from the driver (ref_driver.cpp) so that the reference chain can be fed a known stream
typedef void (*NormalSource)(int n, double* out);
inline NormalSource& normal_source() { static NormalSource s = nullptr; return s; }
inline NumericVector rnorm(int n) {
    NumericVector z; z.v.assign(n, 0.0);
    if (normal_source()) normal_source()(n, z.v.data());
    return z;
}
struct NullStream {
    template <class T> NullStream& operator<<(const T&) { return *this; }
    NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
static NullStream Rcout;
}  // namespace Rcpp
