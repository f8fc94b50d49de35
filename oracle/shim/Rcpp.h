// oracle/shim/Rcpp.h — STAND-IN for <Rcpp.h> and the R API pieces the reference headers touch (TEST INFRASTRUCTURE).
#pragma once
#include <cmath>
#include <limits>
#include <ostream>
#include <string>
#include <vector>
#include <stdexcept>
#include <type_traits>

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
// ---- pieces the Rcpp ADAPTERS in src/ use (Rcpp::stop, Rcpp::List::create(Rcpp::_["name"] = value, ...), unif_rand) -----------------
// Values are kept as (name, rows, cols, doubles); tests/adapters_driver.cpp reads them back by name.
struct ListEntry { std::string name; int rows = 0, cols = 0; std::vector<double> v; };
struct Named {
    std::string name;
    template <class T>
    ListEntry operator=(const std::vector<T>& x) const {          // the *_sparse exports return std::vector members (src/mcml_optim.cpp:181-182)
        ListEntry e; e.name = name; e.rows = (int)x.size(); e.cols = 1; e.v.assign(x.begin(), x.end());
        return e;
    }
    template <class T>
    ListEntry operator=(const T& x) const {
        ListEntry e; e.name = name;
        if constexpr (std::is_arithmetic<T>::value) { e.rows = e.cols = 1; e.v.assign(1, (double)x); }
        else { e.rows = x.rows(); e.cols = x.cols(); e.v.assign(x.data(), x.data() + (size_t)x.rows() * x.cols()); }
        return e;
    }
};
struct NamedPlaceholder { Named operator[](const char* n) const { return Named{n}; } };
static const NamedPlaceholder _{};
struct List {
    std::vector<ListEntry> e;
    template <class... A> static List create(const A&... a) { List l; (l.e.push_back(a), ...); return l; }
    const ListEntry& operator[](const std::string& n) const { for (const auto& x : e) if (x.name == n) return x; throw std::runtime_error("shim: no list entry " + n); }
};
[[noreturn]] inline void stop(const std::string& msg) { throw std::runtime_error(msg); }
// R's uniform RNG (R_ext/Random.h unif_rand, valid under RNGScope): a seedable 64-bit LCG here
inline unsigned long long& unif_state() { static unsigned long long s = 0x9E3779B97F4A7C15ull; return s; }
}  // namespace Rcpp
inline double unif_rand() {
    unsigned long long& s = Rcpp::unif_state();
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return (double)(s >> 11) * (1.0 / 9007199254740992.0);
}
namespace Rcpp {
struct NullStream {
    template <class T> NullStream& operator<<(const T&) { return *this; }
    NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
static NullStream Rcout;
}  // namespace Rcpp
