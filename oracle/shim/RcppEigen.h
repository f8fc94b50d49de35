// oracle/shim/RcppEigen.h — STAND-IN for <RcppEigen.h> (TEST INFRASTRUCTURE).
//
// The reference's numeric headers (inst/include/glmmrmcml/*.h) are written against Eigen + Rcpp, neither of which is
// installed here.  This header supplies the small, eagerly evaluated subset of the Eigen API those headers use, so that
// oracle/ref_driver.cpp can compile the reference headers UNMODIFIED, from where they lie under /root/reference, into
// oracle/_ref/libref.so.  Semantics follow Eigen (column-major, matrix vs array flavour of operator* and inverse()).
// It is not a general Eigen replacement and nothing under glmmrmcml_b200/ includes it.
#pragma once
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include <algorithm>
#include "Rcpp.h"

namespace Eigen {

const int Dynamic = -1;

template <class T> struct Base;
typedef Base<double> BD;

template <class T>
struct Base {
    int r = 0, c = 0;
    bool arr = false;                 // array flavour: operator* and inverse() act element-wise
    std::vector<T> d;
    typedef T Scalar;
    Base() {}
    Base(int rows, int cols, bool a = false) : r(rows), c(cols), arr(a), d((size_t)rows * cols) {}
    int rows() const { return r; }
    int cols() const { return c; }
    int size() const { return r * c; }
    T* data() { return d.data(); }
    const T* data() const { return d.data(); }
    T& operator()(int i, int j) { return d[(size_t)j * r + i]; }
    const T& operator()(int i, int j) const { return d[(size_t)j * r + i]; }
    T& operator()(int i) { return d[i]; }
    const T& operator()(int i) const { return d[i]; }
    T& operator[](int i) { return d[i]; }
    const T& operator[](int i) const { return d[i]; }
    Base& noalias() { return *this; }
    operator T() const { if (r * c != 1) throw std::runtime_error("shim: 1x1 expected"); return d[0]; }

    struct ColRef {
        Base* m; int j;
        operator Base() const { Base o(m->r, 1, m->arr); for (int i = 0; i < m->r; i++) o.d[i] = (*m)(i, j); return o; }
        ColRef& operator=(const Base& v) { for (int i = 0; i < m->r; i++) (*m)(i, j) = v.d[i]; return *this; }
        ColRef& operator=(const ColRef& v) { Base t = v; return *this = t; }
        ColRef& operator+=(const Base& v) { for (int i = 0; i < m->r; i++) (*m)(i, j) += v.d[i]; return *this; }
        T& operator()(int i) { return (*m)(i, j); }
        Base segment(int a, int n) const { Base o(n, 1, m->arr); for (int i = 0; i < n; i++) o.d[i] = (*m)(a + i, j); return o; }
        Base transpose() const { Base t = *this; return t.transpose(); }
        Base array() const { Base t = *this; t.arr = true; return t; }
        Base<int> operator==(T v) const { Base t = *this; return t == v; }
    };
    struct ConstColRef {
        const Base* m; int j;
        operator Base() const { Base o(m->r, 1, m->arr); for (int i = 0; i < m->r; i++) o.d[i] = (*m)(i, j); return o; }
        const T& operator()(int i) const { return (*m)(i, j); }
        Base segment(int a, int n) const { Base o(n, 1, m->arr); for (int i = 0; i < n; i++) o.d[i] = (*m)(a + i, j); return o; }
        Base transpose() const { Base t = *this; return t.transpose(); }
        Base array() const { Base t = *this; t.arr = true; return t; }
        Base<int> operator==(T v) const { Base t = *this; return t == v; }
    };
    ColRef col(int j) { return ColRef{this, j}; }
    ConstColRef col(int j) const { return ConstColRef{this, j}; }

    struct BlockRef {
        Base* m; int i0, j0, p, q;
        operator Base() const { Base o(p, q, m->arr); for (int j = 0; j < q; j++) for (int i = 0; i < p; i++) o(i, j) = (*m)(i0 + i, j0 + j); return o; }
        BlockRef& operator=(const Base& v) { for (int j = 0; j < q; j++) for (int i = 0; i < p; i++) (*m)(i0 + i, j0 + j) = v(i, j); return *this; }
    };
    BlockRef block(int i0, int j0, int p, int q) { return BlockRef{this, i0, j0, p, q}; }
    struct SegRef {
        Base* m; int a, n;
        operator Base() const { Base o(n, 1, m->arr); for (int i = 0; i < n; i++) o.d[i] = m->d[a + i]; return o; }
        SegRef& operator=(const Base& v) { for (int i = 0; i < n; i++) m->d[a + i] = v.d[i]; return *this; }
        Base matrix() const { Base o = *this; o.arr = false; return o; }
    };
    SegRef segment(int a, int n) { return SegRef{this, a, n}; }
    Base segment(int a, int n) const { Base o(n, 1, arr); for (int i = 0; i < n; i++) o.d[i] = d[a + i]; return o; }
    Base head(int n) const { Base o(1, std::min(n, size()), arr); for (int i = 0; i < o.c; i++) o.d[i] = d[i]; return o; }

    Base transpose() const { Base o(c, r, arr); for (int j = 0; j < c; j++) for (int i = 0; i < r; i++) o(j, i) = (*this)(i, j); return o; }
    Base array() const { Base o = *this; o.arr = true; return o; }
    Base matrix() const { Base o = *this; o.arr = false; return o; }
    template <class F> Base map(F f) const { Base o(r, c, arr); for (size_t k = 0; k < d.size(); k++) o.d[k] = f(d[k]); return o; }
    Base exp() const { return map([](T x) { return std::exp(x); }); }
    Base log() const { return map([](T x) { return std::log(x); }); }
    Base square() const { return map([](T x) { return x * x; }); }
    Base cwiseAbs() const { return map([](T x) { return std::abs(x); }); }
    T sum() const { T s = 0; for (const T& x : d) s += x; return s; }
    T mean() const { return sum() / (T)d.size(); }
    T maxCoeff() const { return *std::max_element(d.begin(), d.end()); }
    bool all() const { for (const T& x : d) if (!x) return false; return true; }
    Base<int> operator==(T v) const { Base<int> o(r, c, true); for (size_t k = 0; k < d.size(); k++) o.d[k] = (d[k] == v); return o; }
    struct Rowwise { const Base* m; Base mean() const { Base o(m->r, 1, m->arr); for (int i = 0; i < m->r; i++) { T s = 0; for (int j = 0; j < m->c; j++) s += (*m)(i, j); o.d[i] = s / m->c; } return o; } };
    Rowwise rowwise() const { return Rowwise{this}; }

    // element-wise inverse for arrays; matrix inverse (Gauss-Jordan, partial pivoting) for matrices
    Base inverse() const {
        if (arr) return map([](T x) { return (T)1 / x; });
        if (r != c) throw std::runtime_error("shim: inverse of a non-square matrix");
        int n = r; Base A = *this, I(n, n);
        for (int i = 0; i < n; i++) I(i, i) = 1;
        for (int k = 0; k < n; k++) {
            int piv = k; for (int i = k + 1; i < n; i++) if (std::abs(A(i, k)) > std::abs(A(piv, k))) piv = i;
            if (piv != k) for (int j = 0; j < n; j++) { std::swap(A(k, j), A(piv, j)); std::swap(I(k, j), I(piv, j)); }
            T p = A(k, k);
            for (int j = 0; j < n; j++) { A(k, j) /= p; I(k, j) /= p; }
            for (int i = 0; i < n; i++) if (i != k) { T f = A(i, k); if (f != 0) for (int j = 0; j < n; j++) { A(i, j) -= f * A(k, j); I(i, j) -= f * I(k, j); } }
        }
        return I;
    }
    struct LLTs {
        Base L;
        const Base& matrixL() const { return L; }
        Base solve(const Base& B) const {
            int n = L.r; Base X = B;
            for (int j = 0; j < B.c; j++) {
                for (int i = 0; i < n; i++) { T s = X(i, j); for (int k = 0; k < i; k++) s -= L(i, k) * X(k, j); X(i, j) = s / L(i, i); }
                for (int i = n - 1; i >= 0; i--) { T s = X(i, j); for (int k = i + 1; k < n; k++) s -= L(k, i) * X(k, j); X(i, j) = s / L(i, i); }
            }
            return X;
        }
    };
    LLTs llt() const {
        int n = r; LLTs o; o.L = Base(n, n);
        for (int j = 0; j < n; j++) {
            T s = (*this)(j, j); for (int k = 0; k < j; k++) s -= o.L(j, k) * o.L(j, k);
            o.L(j, j) = std::sqrt(s);
            for (int i = j + 1; i < n; i++) { T t = (*this)(i, j); for (int k = 0; k < j; k++) t -= o.L(i, k) * o.L(j, k); o.L(i, j) = t / o.L(j, j); }
        }
        return o;
    }
    Base& operator+=(const Base& o) { for (size_t k = 0; k < d.size(); k++) d[k] += o.d[k]; return *this; }
    Base& operator-=(const Base& o) { for (size_t k = 0; k < d.size(); k++) d[k] -= o.d[k]; return *this; }
    Base& operator*=(T s) { for (T& x : d) x *= s; return *this; }
};

// ---- arithmetic on double containers (non-template so that the proxy types convert implicitly) ----
inline BD operator+(const BD& a, const BD& b) { BD o = a; for (size_t k = 0; k < o.d.size(); k++) o.d[k] += b.d[k]; return o; }
inline BD operator-(const BD& a, const BD& b) { BD o = a; for (size_t k = 0; k < o.d.size(); k++) o.d[k] -= b.d[k]; return o; }
inline BD operator-(const BD& a, double s) { BD o = a; for (double& x : o.d) x -= s; return o; }
inline BD operator+(const BD& a, double s) { BD o = a; for (double& x : o.d) x += s; return o; }
inline BD operator*(double s, const BD& a) { BD o = a; for (double& x : o.d) x *= s; return o; }
inline BD operator*(const BD& a, double s) { return s * a; }
inline BD operator/(const BD& a, double s) { BD o = a; for (double& x : o.d) x /= s; return o; }
inline BD operator*(const BD& a, const BD& b) {
    if (a.arr || b.arr) { BD o = a; for (size_t k = 0; k < o.d.size(); k++) o.d[k] *= b.d[k]; return o; }
    if (a.c != b.r) throw std::runtime_error("shim: matrix product size mismatch");
    BD o(a.r, b.c);
    for (int j = 0; j < b.c; j++)
        for (int k = 0; k < a.c; k++) { const double bv = b(k, j); for (int i = 0; i < a.r; i++) o(i, j) += a(i, k) * bv; }
    return o;
}

// ---- the named Eigen types ----
#define SHIM_TYPE(NAME, T, ARR, VEC)                                                                          \
    struct NAME : Base<T> {                                                                                    \
        NAME() { this->arr = ARR; }                                                                            \
        explicit NAME(int n) : Base<T>(VEC ? n : n, VEC ? 1 : n, ARR) {}                                       \
        NAME(int rr, int cc) : Base<T>(rr, cc, ARR) {}                                                         \
        NAME(const Base<T>& o) : Base<T>(o) { this->arr = ARR; }                                               \
        NAME(const typename Base<T>::ColRef& o) : Base<T>(Base<T>(o)) { this->arr = ARR; }                     \
        NAME(const typename Base<T>::ConstColRef& o) : Base<T>(Base<T>(o)) { this->arr = ARR; }                \
        NAME(const typename Base<T>::BlockRef& o) : Base<T>(Base<T>(o)) { this->arr = ARR; }                   \
        NAME(const typename Base<T>::SegRef& o) : Base<T>(Base<T>(o)) { this->arr = ARR; }                     \
        NAME& operator=(const Base<T>& o) { Base<T>::operator=(o); this->arr = ARR; return *this; }            \
        static NAME Zero(int n) { NAME o(n); return o; }                                                       \
        static NAME Zero(int rr, int cc) { NAME o(rr, cc); return o; }                                         \
        static NAME Ones(int n) { NAME o(n); for (T& x : o.d) x = 1; return o; }                               \
        static NAME Identity(int rr, int cc) { NAME o(rr, cc); for (int i = 0; i < std::min(rr, cc); i++) o(i, i) = 1; return o; } \
        struct MapRef { T* p; int n; MapRef& operator=(const Base<T>& v) { for (int i = 0; i < n; i++) p[i] = v.d[i]; return *this; } }; \
        static MapRef Map(T* p, int n) { return MapRef{p, n}; }                                                \
    };
SHIM_TYPE(MatrixXd, double, false, false)
SHIM_TYPE(VectorXd, double, false, true)
SHIM_TYPE(ArrayXd, double, true, true)
SHIM_TYPE(ArrayXXd, double, true, false)
SHIM_TYPE(ArrayXXi, int, true, false)
SHIM_TYPE(ArrayXi, int, true, true)
#undef SHIM_TYPE

template <class S, int R, int C> struct Matrix : MatrixXd { using MatrixXd::MatrixXd; Matrix(const BD& o) : MatrixXd(o) {} };

// Eigen::Map<M>(ptr, n [, m]) : read-only use in the reference (copy semantics suffice)
template <class M>
struct Map : M {
    Map(typename M::Scalar* p, int n) : M(n), wp_(p) { for (int i = 0; i < n; i++) this->d[i] = p[i]; }
    Map(typename M::Scalar* p, int rr, int cc) : M(rr, cc) { for (size_t i = 0; i < this->d.size(); i++) this->d[i] = p[i]; }
    Map(const Rcpp::NumericVector& z) : M((int)z.v.size()) { for (size_t i = 0; i < z.v.size(); i++) this->d[i] = z.v[i]; }
    // Eigen::Map<T>(ptr, n) = value : writes through to the mapped memory (sparsedmatrix.h:32-34)
    typename M::Scalar* wp_ = nullptr;
    Map& operator=(const M& o) { if (wp_) for (size_t i = 0; i < o.d.size() && i < this->d.size(); i++) wp_[i] = o.d[i]; this->d = o.d; return *this; }
};

template <class M>
struct LLT {
    BD L;
    LLT(const BD& A) { L = A.llt().L; }
    const BD& matrixL() const { return L; }
};

}  // namespace Eigen

namespace Rcpp {
template <class T> T as(const NumericVector& z) { return T(z); }
}
