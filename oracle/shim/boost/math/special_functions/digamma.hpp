// oracle/shim stand-in for boost::math::digamma (only the beta family, out of scope, touches it).
#pragma once
#include <cmath>
namespace boost { namespace math {
inline double digamma(double x) {
    double r = 0;
    while (x < 6) { r -= 1 / x; x += 1; }
    double f = 1 / (x * x);
    return r + std::log(x) - 0.5 / x - f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f / 132))));
}
}}
