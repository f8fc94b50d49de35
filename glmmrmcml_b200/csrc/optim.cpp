// optim.cpp — host-side optimiser and finite-difference stencils that stand in for rminqa (Rbobyqa, Functor::Gradient,
// Functor::Hessian; call sites mcmloptim.h:58-66,73-85,93-109,300-304,335-352).  rminqa is an un-vendored dependency
// of the reference (DESCRIPTION:26); any bounded optimiser that reaches the same optimum within the MCML tolerance is a
// valid replacement.
//
// B200-first difference: BOBYQA evaluates ONE point at a time, which on a GPU costs one launch + one device->host read
// per evaluation.  Here every phase evaluates a BATCH of points per device round trip — the 2k+1 points of a central
// finite-difference gradient, the candidates of a line search, all 4k^2 points of the optimhess stencil — through the
// batched objective callback, so an M-step is a handful of synchronisations instead of hundreds.
#include "common.cuh"
#include <algorithm>
#include <limits>

namespace {

struct Evaluator {
    gmb_objective_batch f; void* user; int n; int nfev = 0; int rc = GMB_OK;
    // evaluates k points (columns of X, n x k)
    bool eval(const std::vector<double>& X, int k, std::vector<double>& out) {
        out.assign(k, 0.0);
        rc = f(X.data(), n, k, out.data(), user);
        nfev += k;
        return rc == GMB_OK;
    }
};

inline double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

// central differences with the step clamped to the bounds, denominators = actual step sum (optim.c fmingr with bounds,
// the scheme SURVEY.md App. C.4 attributes to rminqa's Functor::Gradient).  Returns f(x) in *f0 as well.
bool fd_gradient(Evaluator& ev, const std::vector<double>& x, const std::vector<double>& h, const double* lower, const double* upper,
                 double* f0, std::vector<double>& g) {
    const int n = ev.n;
    std::vector<double> X((size_t)n * (2 * n + 1)), den(n), F;
    for (int c = 0; c < 2 * n + 1; c++) for (int i = 0; i < n; i++) X[(size_t)c * n + i] = x[i];
    for (int i = 0; i < n; i++) {
        double up = x[i] + h[i], dn = x[i] - h[i];
        if (upper && up > upper[i]) up = upper[i];
        if (lower && dn < lower[i]) dn = lower[i];
        X[(size_t)(1 + 2 * i) * n + i] = up;
        X[(size_t)(2 + 2 * i) * n + i] = dn;
        den[i] = up - dn;
    }
    if (!ev.eval(X, 2 * n + 1, F)) return false;
    *f0 = F[0];
    g.resize(n);
    for (int i = 0; i < n; i++) g[i] = den[i] > 0 ? (F[1 + 2 * i] - F[2 + 2 * i]) / den[i] : 0.0;
    return true;
}

}  // namespace

// Minimises f over the box [lower, upper] (either may be NULL = unbounded; entries may be +-inf).
// Projected BFGS on batched central-difference gradients with a batched line search.
extern "C" int gmb_minimize_bounded(gmb_objective_batch f, void* user, int n, double* x_io, const double* lower, const double* upper,
                         double rhobeg, double xtol, int maxit, double* fmin, int* nfev) {
    if (n <= 0) return gmb_set_error(GMB_EINVAL, "gmb_minimize_bounded: n must be positive");
    Evaluator ev{f, user, n};
    std::vector<double> x(x_io, x_io + n), g, gn, h(n), d(n), H((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) {
        if (lower) x[i] = std::max(x[i], lower[i]);
        if (upper) x[i] = std::min(x[i], upper[i]);
    }
    if (!(rhobeg > 0)) {
        double mx = 0; for (int i = 0; i < n; i++) mx = std::max(mx, std::fabs(x[i]));
        rhobeg = std::min(0.95, 0.2 * mx);                     // BOBYQA's default trust radius (App. C.4)
        if (!(rhobeg > 0)) rhobeg = 0.1;
    }
    auto set_steps = [&]() { for (int i = 0; i < n; i++) h[i] = 1e-5 * std::max(1.0, std::fabs(x[i])); };
    set_steps();
    double fx;
    if (!fd_gradient(ev, x, h, lower, upper, &fx, g)) return ev.rc;
    if (!(fx == fx) || std::isinf(fx)) return gmb_set_error(GMB_EINVAL, "objective is not finite at the starting point");
    bool fresh = true;   // H is a scaled identity
    auto reset_H = [&](double scale) { std::fill(H.begin(), H.end(), 0.0); for (int i = 0; i < n; i++) H[(size_t)i * n + i] = scale; fresh = true; };
    {
        double gn2 = 0; for (int i = 0; i < n; i++) gn2 += g[i] * g[i];
        reset_H(gn2 > 0 ? rhobeg / std::sqrt(gn2) : 1.0);
    }
    const int NT = 12;
    std::vector<double> XT((size_t)n * NT), FT, xn(n), s(n), yv(n), Hy(n);
    for (int it = 0; it < maxit; it++) {
        // active set: at a bound with the descent direction pointing outwards
        std::vector<char> fixed(n, 0);
        for (int i = 0; i < n; i++) {
            if (lower && x[i] <= lower[i] && g[i] > 0) fixed[i] = 1;
            if (upper && x[i] >= upper[i] && g[i] < 0) fixed[i] = 1;
        }
        double pg = 0;
        for (int i = 0; i < n; i++) if (!fixed[i]) pg = std::max(pg, std::fabs(g[i]));
        if (pg == 0) break;
        for (int i = 0; i < n; i++) {
            double v = 0;
            if (!fixed[i]) for (int j = 0; j < n; j++) if (!fixed[j]) v -= H[(size_t)j * n + i] * g[j];
            d[i] = v;
        }
        double slope = 0; for (int i = 0; i < n; i++) slope += d[i] * g[i];
        if (!(slope < 0)) {   // not a descent direction: fall back to steepest descent
            double gn2 = 0; for (int i = 0; i < n; i++) if (!fixed[i]) gn2 += g[i] * g[i];
            reset_H(rhobeg / std::sqrt(gn2));
            for (int i = 0; i < n; i++) d[i] = fixed[i] ? 0.0 : -H[(size_t)i * n + i] * g[i];
        }
        // batched line search: t = 2, 1, 1/2, ... ; take the best point
        double t = 2.0;
        for (int c = 0; c < NT; c++, t *= (c < 8 ? 0.5 : 0.125))
            for (int i = 0; i < n; i++) {
                double v = x[i] + t * d[i];
                if (lower) v = std::max(v, lower[i]);
                if (upper) v = std::min(v, upper[i]);
                XT[(size_t)c * n + i] = v;
            }
        if (!ev.eval(XT, NT, FT)) return ev.rc;
        int best = -1; double fb = fx;
        for (int c = 0; c < NT; c++) if (FT[c] == FT[c] && FT[c] < fb) { fb = FT[c]; best = c; }
        if (best < 0) {
            if (!fresh) { double gn2 = 0; for (int i = 0; i < n; i++) gn2 += g[i] * g[i]; reset_H(rhobeg * 0.1 / std::sqrt(gn2)); rhobeg *= 0.1; continue; }
            break;            // no decrease along steepest descent at steps down to ~1e-9: converged to FD accuracy
        }
        for (int i = 0; i < n; i++) { xn[i] = XT[(size_t)best * n + i]; s[i] = xn[i] - x[i]; }
        double fnew;
        std::vector<double> xv(xn);
        for (int i = 0; i < n; i++) h[i] = 1e-5 * std::max(1.0, std::fabs(xn[i]));
        if (!fd_gradient(ev, xv, h, lower, upper, &fnew, gn)) return ev.rc;
        double sy = 0, yy = 0, ss = 0, smax = 0;
        for (int i = 0; i < n; i++) { yv[i] = gn[i] - g[i]; sy += s[i] * yv[i]; yy += yv[i] * yv[i]; ss += s[i] * s[i];
                                      smax = std::max(smax, std::fabs(s[i]) / std::max(1.0, std::fabs(xn[i]))); }
        if (sy > 1e-12 * std::sqrt(ss * yy) && yy > 0) {
            if (fresh) { reset_H(sy / yy); fresh = false; }
            // BFGS inverse update: H <- (I - rho s y^T) H (I - rho y s^T) + rho s s^T
            const double rho = 1.0 / sy;
            for (int i = 0; i < n; i++) { double v = 0; for (int j = 0; j < n; j++) v += H[(size_t)j * n + i] * yv[j]; Hy[i] = v; }
            double yHy = 0; for (int i = 0; i < n; i++) yHy += yv[i] * Hy[i];
            for (int j = 0; j < n; j++)
                for (int i = 0; i < n; i++)
                    H[(size_t)j * n + i] += -rho * (s[i] * Hy[j] + Hy[i] * s[j]) + rho * rho * yHy * s[i] * s[j] + rho * s[i] * s[j];
        }
        const double df = fx - fnew;
        x = xn; g = gn; fx = fnew;
        if (smax <= xtol) break;
        if (df <= 1e-14 * std::max(1.0, std::fabs(fx))) break;
    }
    for (int i = 0; i < n; i++) x_io[i] = x[i];
    if (fmin) *fmin = fx;
    if (nfev) *nfev = ev.nfev;
    return GMB_OK;
}

// Gradient by bounded central differences (rminqa Functor::Gradient as used by mcmloptim::f_grad, mcmloptim.h:296-317).
extern "C" int gmb_fd_gradient(gmb_objective_batch f, void* user, int n, const double* x, const double* ndeps,
                    const double* lower, const double* upper, int usebounds, double* grad) {
    Evaluator ev{f, user, n};
    std::vector<double> xv(x, x + n), h(ndeps, ndeps + n), g;
    double f0;
    if (!fd_gradient(ev, xv, h, usebounds ? lower : nullptr, usebounds ? upper : nullptr, &f0, g)) return ev.rc;
    for (int i = 0; i < n; i++) grad[i] = g[i];
    return GMB_OK;
}

// optimhess stencil (R's optim.c optimhess; rminqa Functor::Hessian as used by mcmloptim::f_hess, mcmloptim.h:333-355):
//   H[i, .] = (Gradient(x + e_i h_i) - Gradient(x - e_i h_i)) / (2 h_i), Gradient = bounded central differences with the
//   same steps; then H <- (H + H^T)/2.  All 4 n^2 points are evaluated in one batch.  hess is n x n column-major.
extern "C" int gmb_fd_hessian(gmb_objective_batch f, void* user, int n, const double* x, const double* ndeps,
                   const double* lower, const double* upper, int usebounds, double* hess, int* nfev) {
    if (n <= 0) return gmb_set_error(GMB_EINVAL, "gmb_fd_hessian: n must be positive");
    Evaluator ev{f, user, n};
    const int npts = 4 * n * n;
    std::vector<double> X((size_t)n * npts), den((size_t)2 * n * n), F;
    // point index: ((i * 2 + side) * n + j) * 2 + dir ; side: 0 => x + h_i e_i, 1 => x - h_i e_i ; dir: 0 up, 1 down
    for (int i = 0; i < n; i++)
        for (int side = 0; side < 2; side++) {
            std::vector<double> base(x, x + n);
            base[i] += side == 0 ? ndeps[i] : -ndeps[i];
            for (int j = 0; j < n; j++) {
                double up = base[j] + ndeps[j], dn = base[j] - ndeps[j];
                if (usebounds && upper && up > upper[j]) up = upper[j];
                if (usebounds && lower && dn < lower[j]) dn = lower[j];
                const size_t p = (((size_t)i * 2 + side) * n + j) * 2;
                for (int k = 0; k < n; k++) { X[p * n + k] = base[k]; X[(p + 1) * n + k] = base[k]; }
                X[p * n + j] = up; X[(p + 1) * n + j] = dn;
                den[((size_t)i * 2 + side) * n + j] = up - dn;
            }
        }
    if (!ev.eval(X, npts, F)) return ev.rc;
    std::vector<double> H((size_t)n * n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            const size_t p0 = (((size_t)i * 2 + 0) * n + j) * 2, p1 = (((size_t)i * 2 + 1) * n + j) * 2;
            const double g1 = (F[p0] - F[p0 + 1]) / den[((size_t)i * 2 + 0) * n + j];
            const double g2 = (F[p1] - F[p1 + 1]) / den[((size_t)i * 2 + 1) * n + j];
            H[(size_t)j * n + i] = (g1 - g2) / (2 * ndeps[i]);
        }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) hess[(size_t)j * n + i] = 0.5 * (H[(size_t)j * n + i] + H[(size_t)i * n + j]);
    if (nfev) *nfev = ev.nfev;
    return GMB_OK;
}
