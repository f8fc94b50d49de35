// oracle/shim/rbobyqa.h — STAND-IN for rminqa's <rbobyqa.h> (TEST INFRASTRUCTURE).
//
// rminqa (Powell's BOBYQA + R's optim()-style finite differences) is not available offline.  What the reference needs from it
// (call sites mcmloptim.h:58-66,73-85,93-109,117-147,160-173,183-192,296-355) is provided as follows:
//   * Functor::Gradient / Functor::Hessian — R's optim() numerical gradient (fmingr: central differences with steps ndeps_, clamped to
//     [lower_, upper_] when usebounds_, the denominator being the step actually taken) and R's optimhess() (rows from differences of
//     that gradient at x +- ndeps_[i] e_i, then symmetrised): SURVEY App. C.4.  These are deterministic formulas, restated exactly.
//   * Rbobyqa::minimize — NOT Powell's algorithm: a projected BFGS on central-difference gradients with a backtracking line search and a
//     trust-region-like cap on the step length (initially BOBYQA's rhobeg), so that it stays in the basin of its starting point.
//     Any correct bounded minimiser reaches the same optimum to far below the MCML tolerance; results obtained through it are
//     compared at optimiser tolerance, never bit for bit.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <stdexcept>
#include <vector>
namespace rminqa {
template <class V>
class Functor {
public:
    struct { std::vector<double> ndeps_, lower_, upper_; int usebounds_ = 0; } os;
    virtual double operator()(const V& par) = 0;
    virtual ~Functor() {}
    // R: src/appl/optim.c fmingr, numerical branch (parscale = 1, fnscale = 1)
    void Gradient(const V& par, V& grad) {
        const size_t n = par.size();
        V x = par;
        if (grad.size() != n) grad.resize(n);
        for (size_t i = 0; i < n; i++) {
            const double eps = os.ndeps_.size() > i ? os.ndeps_[i] : 1e-3;
            if (!os.usebounds_) {
                x[i] = par[i] + eps; const double v1 = (*this)(x);
                x[i] = par[i] - eps; const double v2 = (*this)(x);
                grad[i] = (v1 - v2) / (2 * eps);
            } else {
                double epsused = eps, eps2 = eps, tmp = par[i] + eps;
                if (os.upper_.size() > i && tmp > os.upper_[i]) { tmp = os.upper_[i]; epsused = tmp - par[i]; }
                x[i] = tmp; const double v1 = (*this)(x);
                tmp = par[i] - eps;
                if (os.lower_.size() > i && tmp < os.lower_[i]) { tmp = os.lower_[i]; eps2 = par[i] - tmp; }
                x[i] = tmp; const double v2 = (*this)(x);
                grad[i] = (v1 - v2) / (epsused + eps2);
            }
            x[i] = par[i];
        }
    }
    // R: src/library/stats/src/optim.c optimhess
    void Hessian(const V& par, V& hess) {
        const size_t n = par.size();
        if (hess.size() != n * n) hess.assign(n * n, 0.0);
        V dpar = par, df1(n), df2(n);
        for (size_t i = 0; i < n; i++) {
            const double eps = os.ndeps_.size() > i ? os.ndeps_[i] : 1e-3;
            dpar[i] = par[i] + eps; Gradient(dpar, df1);
            dpar[i] = par[i] - eps; Gradient(dpar, df2);
            for (size_t j = 0; j < n; j++) hess[i * n + j] = (df1[j] - df2[j]) / (2 * eps);
            dpar[i] = par[i];
        }
        for (size_t i = 0; i < n; i++)
            for (size_t j = 0; j < i; j++) { const double t = 0.5 * (hess[i * n + j] + hess[j * n + i]); hess[i * n + j] = hess[j * n + i] = t; }
    }
};

template <class F, class V>
class Rbobyqa {
public:
    struct { int iprint = 0; } control;
    V par_, lower_, upper_;
    double fval_ = 0;
    int feval_ = 0;
    void set_upper(const V& u) { upper_ = u; }
    void set_lower(const V& l) { lower_ = l; }
    V par() const { return par_; }
    double fval() const { return fval_; }

    void minimize(F& f, V& x0) {
        const size_t n = x0.size();
        const double inf = std::numeric_limits<double>::infinity();
        std::vector<double> lo(n, -inf), up(n, inf), x(x0.begin(), x0.end());
        for (size_t i = 0; i < n; i++) { if (lower_.size() > i) lo[i] = lower_[i]; if (upper_.size() > i) up[i] = upper_[i]; }
        auto proj = [&](std::vector<double>& v) { for (size_t i = 0; i < n; i++) v[i] = std::min(std::max(v[i], lo[i]), up[i]); };
        auto eval = [&](const std::vector<double>& v) {
            V p(v.begin(), v.end()); feval_++;
            const double r = f(p);
            return std::isfinite(r) ? r : 1e300;
        };
        auto grad = [&](const std::vector<double>& v, std::vector<double>& g) {
            std::vector<double> y = v;
            for (size_t i = 0; i < n; i++) {
                const double h = 1e-6 * std::max(1.0, std::fabs(v[i]));
                const double a = std::min(v[i] + h, up[i]), b = std::max(v[i] - h, lo[i]);
                y[i] = a; const double fa = eval(y);
                y[i] = b; const double fb = eval(y);
                y[i] = v[i];
                g[i] = (a > b) ? (fa - fb) / (a - b) : 0.0;
            }
        };
        proj(x);
        std::vector<double> H(n * n, 0.0), g(n), gn(n), d(n), xn(n), s(n), yv(n), Hy(n);
        auto reset = [&] { std::fill(H.begin(), H.end(), 0.0); for (size_t i = 0; i < n; i++) H[i * n + i] = 1.0; };
        reset();
        double fx = eval(x);
        grad(x, g);
        int small = 0; bool fresh = true;
        // step cap in the spirit of BOBYQA's trust region (rhobeg = min(0.95, 0.2 max|x0|)): a first steepest-descent step
        // of full length can leave the basin of the starting point (e.g. land on a bound where the objective happens to be lower)
        double delta = 0; for (size_t i = 0; i < n; i++) delta = std::max(delta, std::fabs(x[i]));
        delta = std::max(1e-3, std::min(0.95, 0.2 * delta));
        for (int it = 0; it < 400 && small < 2; it++) {
            // variables held at a bound whose gradient points outward stay there
            std::vector<char> fixed(n, 0);
            for (size_t i = 0; i < n; i++) fixed[i] = (x[i] <= lo[i] && g[i] > 0) || (x[i] >= up[i] && g[i] < 0);
            double dg = 0;
            for (size_t i = 0; i < n; i++) {
                double t = 0;
                if (!fixed[i]) for (size_t j = 0; j < n; j++) if (!fixed[j]) t -= H[i * n + j] * g[j];
                d[i] = t; dg += t * g[i];
            }
            if (!(dg < 0)) {
                if (!fresh) { reset(); fresh = true; continue; }
                break;                                                   // projected gradient vanishes
            }
            double dmax = 0; for (size_t i = 0; i < n; i++) dmax = std::max(dmax, std::fabs(d[i]));
            double t = dmax > delta ? delta / dmax : 1.0, fn = fx; bool ok = false;
            const double t0 = t;
            for (int ls = 0; ls < 60; ls++, t *= 0.5) {
                for (size_t i = 0; i < n; i++) xn[i] = x[i] + t * d[i];
                proj(xn);
                fn = eval(xn);
                double lin = 0; for (size_t i = 0; i < n; i++) lin += g[i] * (xn[i] - x[i]);
                if (fn <= fx + 1e-4 * lin && fn < 1e299) { ok = true; break; }
            }
            if (!ok) {
                if (!fresh) { reset(); fresh = true; continue; }
                break;
            }
            delta = (t == t0) ? 2 * delta : std::max(1e-12, 2 * t * dmax);        // grow after an accepted first trial, shrink to the accepted length
            grad(xn, gn);
            double sy = 0, step = 0, scale = 1.0;
            for (size_t i = 0; i < n; i++) { s[i] = xn[i] - x[i]; yv[i] = gn[i] - g[i]; sy += s[i] * yv[i]; step = std::max(step, std::fabs(s[i])); scale = std::max(scale, std::fabs(xn[i])); }
            if (sy > 1e-14) {                                            // BFGS update of the inverse Hessian approximation
                double yHy = 0;
                for (size_t i = 0; i < n; i++) { double a = 0; for (size_t j = 0; j < n; j++) a += H[i * n + j] * yv[j]; Hy[i] = a; yHy += a * yv[i]; }
                for (size_t i = 0; i < n; i++)
                    for (size_t j = 0; j < n; j++) H[i * n + j] += (1 + yHy / sy) * s[i] * s[j] / sy - (Hy[i] * s[j] + s[i] * Hy[j]) / sy;
                fresh = false;
            } else { reset(); fresh = true; }
            small = (step < 1e-9 * scale || std::fabs(fx - fn) < 1e-15 * (1 + std::fabs(fx))) ? small + 1 : 0;
            x = xn; g = gn; fx = fn;
        }
        par_ = V(x.begin(), x.end());
        fval_ = fx;
        x0 = par_;
    }
};
}  // namespace rminqa
