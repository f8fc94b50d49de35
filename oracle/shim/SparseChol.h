// oracle/shim/SparseChol.h — STAND-IN for the SparseChol package's header (TEST INFRASTRUCTURE): just enough for the reference's
// sparsedmatrix.h and the *_sparse exports of src/mcml_optim.cpp to COMPILE.  The sparse-D variants are out of scope;
// every member that would compute throws.
#pragma once
#include <stdexcept>
#include <vector>
class sparse {
public:
    std::vector<int> Ap, Ai;
    std::vector<double> Ax;
    explicit sparse(const std::vector<int>& p) : Ap(p) {}
};
class SparseChol {
public:
    std::vector<double> D;
    sparse* L;
    explicit SparseChol(sparse* m) : L(m) {}
    void ldl_numeric() { throw std::runtime_error("shim: SparseChol is not available (sparse-D variants are out of scope)"); }
    void ldl_lsolve(double*) { throw std::runtime_error("shim: SparseChol is not available"); }
    void ldl_d2solve(double*) { throw std::runtime_error("shim: SparseChol is not available"); }
};
