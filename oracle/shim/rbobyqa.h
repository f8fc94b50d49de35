// oracle/shim/rbobyqa.h — STAND-IN for rminqa's <rbobyqa.h> (TEST INFRASTRUCTURE): just enough for likelihood.h and
// mcmloptim.h to compile.  The optimiser itself is not reproduced: only mcmloptim::mcnr() (no optimiser involved) and
// the objective functors are exercised through oracle/_ref.
#pragma once
#include <stdexcept>
#include <vector>
namespace rminqa {
template <class V>
class Functor {
public:
    struct { std::vector<double> ndeps_, lower_, upper_; int usebounds_ = 0; } os;
    virtual double operator()(const V& par) = 0;
    virtual ~Functor() {}
    void Gradient(const V&, V&) { throw std::runtime_error("shim: rminqa Gradient not available"); }
    void Hessian(const V&, V&) { throw std::runtime_error("shim: rminqa Hessian not available"); }
};
template <class F, class V>
class Rbobyqa {
public:
    struct { int iprint = 0; } control;
    V par_;
    void set_upper(const V&) {}
    void set_lower(const V&) {}
    void minimize(F&, V&) { throw std::runtime_error("shim: rminqa BOBYQA not available"); }
    V par() const { return par_; }
};
}  // namespace rminqa
