// gmb_adapter.h — what every Rcpp adapter in src/ shares: status check -> R error, seed from R's RNG, the covariance triple.
// The adapters keep the reference's exported C++ signatures (src/RcppExports.cpp:16-276 declares them) so that RcppExports.cpp, R/RcppExports.R
// and all R code stay as they are; only the bodies change: each forwards to its C-ABI entry point in include/glmmrmcml_b200.h with the
// Eigen objects' own storage (column-major double / int, zero copy).
#pragma once
#include <RcppEigen.h>
#include <cstdint>
#include <string>
#include "glmmrmcml_b200.h"

namespace gmb_adapter {

// no exception leaves the library: a non-zero status becomes an R error here (BEGIN_RCPP / END_RCPP turn it into a condition)
inline void check(int rc) { if (rc != GMB_OK) Rcpp::stop(std::string(gmb_last_error())); }

// 64 bits from R's generator (every export runs under Rcpp::RNGScope, src/RcppExports.cpp:20): set.seed() makes the device sampler reproducible
inline uint64_t seed_from_r() {
    const uint64_t hi = (uint64_t)(unif_rand() * 4294967296.0), lo = (uint64_t)(unif_rand() * 4294967296.0);
    return (hi << 32) | (lo & 0xffffffffu);
}

// number of chains: 0 = the library chooses (ceil((nsamp + 1) / 32), at most 1024); 1 reproduces the reference's single chain
inline int default_chains() { return 0; }

struct Shape { int B = 0, Q = 0, R = 0; };
inline Shape cov_shape(const Eigen::ArrayXXi& cov) {
    Shape s;
    check(gmb_cov_shape(cov.data(), cov.rows(), &s.B, &s.Q, &s.R));
    return s;
}

}  // namespace gmb_adapter
