// estep.cu — Monte-Carlo E-step kernels over the cached n x m matrix zd = Z u.
//
// K2: E-step objective, mcmlModel::log_likelihood (mcmlmodel.h:284-304) + maths::log_likelihood (moremaths.h:26-102)
// K3: MCNR sufficient sums, mcmloptim::mcnr (mcmloptim.h:198-236) + update_W (mcmlmodel.h:120-134) + detadmu
//     (moremaths.h:118-161), using mean_j X^T W_j X = X^T diag(mean_j w_j) X.
//
// Both stream zd exactly once (algorithmic bytes 8 n m + 16 n): thread-fixed rows, vectorised
// 16-byte loads that are contiguous across the warp, xb/y held in registers, column loop unrolled for
// memory-level parallelism, deterministic two-level reduction (no floating-point atomics).
#include "common.cuh"
#include "gemm_tma.cuh"    // mbarrier / TMA helpers and the tensor-map encoder

namespace {

// storage type of the streamed matrices (zd, F): double, or float in fp32 mode (values are widened on load; all arithmetic stays fp64)
template <class T> struct Pair;
template <> struct Pair<double> { typedef double2 type; };
template <> struct Pair<float> { typedef float2 type; };
template <class T> __device__ __forceinline__ double2 ld2(const T* p) {
    const typename Pair<T>::type v = *reinterpret_cast<const typename Pair<T>::type*>(p);
    return make_double2((double)v.x, (double)v.y);
}

__global__ void xb_kernel(int n, int P, int ldn, const double* __restrict__ X, const double* __restrict__ beta,
                          double* __restrict__ xb) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int p = 0; p < P; p++) s += X[i + (size_t)p * ldn] * beta[p];   // update_beta, mcmlmodel.h:100-102
    xb[i] = s;
}

// ---------------------------------------------------------------------------------------------------
// K2.  blockDim = (TX, TY): thread (tx, ty) owns rows 2*(rt*TX + tx) + {0,1} and columns j0 + ty + k*TY.
// X beta (update_beta, mcmlmodel.h:100-102) is recomputed per thread for its two rows (P fused multiply-adds, X is
// L2-resident) so that one evaluation is ONE launch.
// Output: raw sum over the local columns [0, ncols) in out[0] (divide by niter on the host after the all-reduce).
//
// binomial/logit: l = log(1/(1+exp(-eta))) (y = 1) or log(1 - 1/(1+exp(-eta))) (y = 0), moremaths.h:47-53; both equal
// -log(1 + exp(s eta)) with s = -1 (y = 1) or +1 (y = 0).  The FP64 log is the slowest instruction sequence of the
// stream (profiles/r01_microbench_fp64.txt: log 0.35 Telem/s vs the 0.82 Telem/s that HBM can feed), so the kernel
// multiplies the 8 factors (1 + exp(s eta)) of one unrolled step and takes ONE log of the product — the same sum to
// within a few ulp per term, at one eighth of the log count.
// ---------------------------------------------------------------------------------------------------
constexpr int LOGIT_K_GUARD = 7386;   // 64 * 80 / ln 2: factors below exp(80) — 8 of them cannot overflow a double

template <int FL>
struct RowTerm {
    double xb, y, rowc, mask;
    int smask;      // binomial/logit: 0 for y = 0 (x = eta), -1 for y = 1 (x = -eta)
    __device__ __forceinline__ void init(double xb_, double y_, double rowc_) {
        xb = xb_; y = y_; rowc = rowc_;
        smask = (y_ == 1.0) ? -1 : 0;
        mask = (y_ == 1.0 || y_ == 0.0) ? 1.0 : 0.0;   // other y contribute nothing (the reference leaves logl unset)
    }
};

// accumulates the contribution of one element into (acc, prod); kmax tracks the largest scaled exponent of the group
template <int FL>
__device__ __forceinline__ void ll_accum(const RowTerm<FL>& r, double z, double c0, double inv_sigma, const double* __restrict__ tab,
                                         double& acc, double& prod, int& kmax) {
    const double eta = r.xb + z;                                     // mcmlmodel.h:298
    if (FL == 1) {
        acc += r.y * eta - dev_exp_tab(eta, tab) - r.rowc;           // moremaths.h:33-40
    } else if (FL == 3) {
        // 1 + exp(x), x = -eta (y = 1) or eta (y = 0): dev_exp_tab's steps with the sign applied to the reduced argument and the
        // scaled exponent (round(-a) = -round(a)), so that the sign costs integer instructions only
        double t = fma(eta, 92.33248261689366, 6755399441055744.0);
        int k = __double2loint(t);
        t -= 6755399441055744.0;
        double rr = fma(t, -0.010830424667801708, eta);
        rr = fma(t, -2.8447437476627285e-11, rr);
        k = (k ^ r.smask) - r.smask;
        rr = __hiloint2double(__double2hiint(rr) ^ (r.smask & 0x80000000), __double2loint(rr));
        const double T = tab[k & 63];
        double q = fma(rr, 1.0 / 120.0, 1.0 / 24.0);
        q = fma(q, rr, 1.0 / 6.0);
        q = fma(q, rr, 0.5);
        q = fma(q, rr, 1.0);
        q = q * rr;
        const double m = fma(T, q, T);
        kmax = max(kmax, k);
        const int kc = min(max(k, -64512), 64512);
        const double e = __hiloint2double(__double2hiint(m) + ((kc >> 6) << 20), __double2loint(m));
        prod *= fma(e, r.mask, 1.0);
    } else if (FL == 7) {
        const double zz = (r.y - eta) * inv_sigma;                   // moremaths.h:75-78
        acc += c0 - 0.5 * zz * zz;
    } else {                                                          // codes 2, 4, 5, 6, 8: the generic per-observation term
        acc += dev_family_ll<FL>(r.y, eta, r.rowc, c0, 1.0 / inv_sigma);
    }
}

// binomial/logit, rare path: a group whose product could overflow is redone term by term
template <int FL>
__device__ __forceinline__ double ll_direct(const RowTerm<FL>& r, double z) {
    const double eta = r.xb + z;
    const double x = r.smask ? -eta : eta;
    return r.mask != 0.0 ? -log(1.0 + exp(x)) : 0.0;
}

// deterministic grid-wide sum: per-CTA partials, the last CTA to arrive adds them in a fixed order (bitwise reproducible)
__device__ __forceinline__ void grid_sum_finish(double acc, int TX, double* __restrict__ partials, unsigned int* __restrict__ counter,
                                                double* __restrict__ out, double* red, bool* is_last) {
    double v = warp_sum(acc);
    const int t = threadIdx.y * TX + threadIdx.x;
    const int w = t >> 5, l = t & 31;
    if (l == 0) red[w] = v;
    __syncthreads();
    const int bid = blockIdx.y * gridDim.x + blockIdx.x;
    const int nblocks = gridDim.x * gridDim.y;
    if (w == 0) {
        v = (l < 8) ? red[l] : 0.0;
        v = warp_sum(v);
        if (l == 0) {
            partials[bid] = v;
            __threadfence();
            unsigned int done = atomicAdd(counter, 1u);
            *is_last = (done == (unsigned)nblocks - 1);
        }
    }
    __syncthreads();
    if (*is_last) {
        __threadfence();
        double s = 0.0;
        for (int k = t; k < nblocks; k += 256) s += partials[k];
        s = warp_sum(s);
        __syncthreads();
        if (l == 0) red[w] = s;
        __syncthreads();
        if (w == 0) {
            s = (l < 8) ? red[l] : 0.0;
            s = warp_sum(s);
            if (l == 0) { out[0] = s; *counter = 0u; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K2, binomial/logit on the FACTOR matrix.  The sample matrix is fixed while the M-step evaluates the objective hundreds of times
// (l_optim, f_hess), and exp(s eta) = exp(s xb_i) exp(s zd_ij) with the sign s fixed by y_i: F_ij = exp(s_i zd_ij) is built once per
// sample matrix (build_factor_kernel, next to zd), after which an evaluation needs one exp per ROW and per element
//     prod *= 1 + A_i F_ij          (A_i = exp(s_i xb_i); A_i = 0 drops rows whose y is neither 0 nor 1)
// i.e. two FP64 instructions instead of a 20-instruction exp, and one log per 8 elements: the kernel streams at the rate of the
// Gaussian one instead of being issue bound.  A group whose product leaves the finite range is redone term by term.
// ---------------------------------------------------------------------------------------------------
template <class T>
__global__ void build_factor_kernel(int n, int ldn, int ncols, const T* __restrict__ zd, const double* __restrict__ y,
                                    T* __restrict__ F) {
    const int i = blockIdx.y * blockDim.x + threadIdx.x;          // grid.x (up to 2^31 - 1) runs over the columns
    const int j = blockIdx.x;
    if (i >= ldn || j >= ncols) return;
    double f = 1.0;
    if (i < n) {
        const double yi = y ? y[i] : 0.0;                 // y == NULL (aggregated rows): F = exp(zd) for every row
        const double z = (double)zd[i + (size_t)j * ldn];
        f = (yi == 1.0) ? exp(-1.0 * z) : exp(z);
    }
    F[i + (size_t)j * ldn] = (T)f;
}

template <class T>
__global__ void __launch_bounds__(256) loglik_logit_factor_kernel(int n, int P, int ldn, int ncols, int cols_per_cta,
                                                                  const T* __restrict__ F, const double* __restrict__ X,
                                                                  const double* __restrict__ beta, const double* __restrict__ y,
                                                                  double* __restrict__ partials, unsigned int* __restrict__ counter,
                                                                  double* __restrict__ out) {
    __shared__ double red[32];
    __shared__ bool is_last;
    const int TX = blockDim.x, TY = blockDim.y;
    const int i0 = 2 * (blockIdx.x * TX + threadIdx.x);
    const int j0 = blockIdx.y * cols_per_cta;
    const int j1 = min(j0 + cols_per_cta, ncols);
    double acc = 0.0;
    if (i0 < n) {
        const bool two = (i0 + 1 < n);
        double xb0 = 0.0, xb1 = 0.0;
        for (int p = 0; p < P; p++) {                                // same order as xb_kernel
            const double b = beta[p];
            xb0 += X[i0 + (size_t)p * ldn] * b;
            if (two) xb1 += X[i0 + 1 + (size_t)p * ldn] * b;
        }
        const double y0 = y[i0], y1 = two ? y[i0 + 1] : -1.0;
        const double A0 = (y0 == 1.0) ? exp(-1.0 * xb0) : ((y0 == 0.0) ? exp(xb0) : 0.0);
        const double A1 = (y1 == 1.0) ? exp(-1.0 * xb1) : ((y1 == 0.0) ? exp(xb1) : 0.0);
        const T* col = F + i0;
        int j = j0 + threadIdx.y;
        double2 z[4], zn[4];
        bool have = j + 3 * TY < j1;
        if (have) {
#pragma unroll
            for (int u = 0; u < 4; u++) z[u] = ld2(col + (size_t)(j + u * TY) * ldn);
        }
        while (have) {
            const int jn = j + 4 * TY;
            const bool have_n = jn + 3 * TY < j1;
            if (have_n) {
#pragma unroll
                for (int u = 0; u < 4; u++) zn[u] = ld2(col + (size_t)(jn + u * TY) * ldn);
            }
            double p0 = 1.0, p1 = 1.0;
#pragma unroll
            for (int u = 0; u < 4; u++) { p0 *= fma(A0, z[u].x, 1.0); p1 *= fma(A1, z[u].y, 1.0); }
            const double prod = p0 * p1;
            if (prod <= 1e300) {
                acc -= log(prod);
            } else {                                                  // overflow (or NaN): term by term
#pragma unroll
                for (int u = 0; u < 4; u++) acc -= log(fma(A0, z[u].x, 1.0)) + log(fma(A1, z[u].y, 1.0));
            }
#pragma unroll
            for (int u = 0; u < 4; u++) z[u] = zn[u];
            j = jn; have = have_n;
        }
        for (; j < j1; j += TY) {
            const double2 zz = ld2(col + (size_t)j * ldn);
            acc -= log(fma(A0, zz.x, 1.0)) + log(fma(A1, zz.y, 1.0));
        }
    }
    grid_sum_finish(acc, TX, partials, counter, out, red, &is_last);
}

// The same for a whole batch of parameter vectors in ONE launch (the optimiser's stencils and the 4 k^2 optimhess points arrive as batches):
// blockIdx.z selects a group of NB evaluations, a thread keeps the NB row constants A_i of its two rows in registers and applies them to every
// element it loads, and the grid is sized for one wave over all groups — so F is read once per NB evaluations and the per-thread set-up (x'beta,
// two exp per evaluation) and the grid reduction are amortised over hundreds of columns instead of a dozen.  At C2 (n = 500: one row tile) a
// single evaluation is bound by launch + set-up + reduction latency (14.6 us for 5 10^6 elements); the batch is bound by its FP64 work.
// Deterministic (fixed partition per problem size and batch), but the partition differs from the single evaluation's: values agree to rounding,
// not bit for bit.  beta: P x n_eval; group z covers evaluations min(z NB, n_eval - NB) .. + NB - 1 (a short last group overlaps the one
// before); partials: [group][NB][blocks]; counters: one per group, zero on entry; out: n_eval values.
template <int NB>
__global__ void __launch_bounds__(256, 2) loglik_logit_factor_multi_kernel(int n, int P, int ldn, int ncols, int cols_per_cta, int n_eval,
                                                                        const double* __restrict__ F, const double* __restrict__ X,
                                                                        const double* __restrict__ beta, const double* __restrict__ y,
                                                                        double* __restrict__ partials, unsigned int* __restrict__ counter,
                                                                        double* __restrict__ out) {
    __shared__ double red[32];
    __shared__ bool is_last;
    const int TX = blockDim.x, TY = blockDim.y;
    {
        const int e_start = min((int)blockIdx.z * NB, n_eval - NB);
        beta += (size_t)e_start * P; out += e_start;
        partials += (size_t)blockIdx.z * NB * gridDim.x * gridDim.y; counter += blockIdx.z;
    }
    const int i0 = 2 * (blockIdx.x * TX + threadIdx.x);
    const int j0 = blockIdx.y * cols_per_cta;
    const int j1 = min(j0 + cols_per_cta, ncols);
    double acc[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) acc[b] = 0.0;
    if (i0 < n) {
        const bool two = (i0 + 1 < n);
        const double y0 = y[i0], y1 = two ? y[i0 + 1] : -1.0;
        double A0[NB], A1[NB];
#pragma unroll
        for (int b = 0; b < NB; b++) {
            double xb0 = 0.0, xb1 = 0.0;
            for (int p = 0; p < P; p++) {                            // same order as xb_kernel
                const double bb = beta[p + (size_t)b * P];
                xb0 += X[i0 + (size_t)p * ldn] * bb;
                if (two) xb1 += X[i0 + 1 + (size_t)p * ldn] * bb;
            }
            A0[b] = (y0 == 1.0) ? exp(-1.0 * xb0) : ((y0 == 0.0) ? exp(xb0) : 0.0);
            A1[b] = (y1 == 1.0) ? exp(-1.0 * xb1) : ((y1 == 0.0) ? exp(xb1) : 0.0);
        }
        const double* col = F + i0;
        int j = j0 + threadIdx.y;
        double2 z[4], zn[4];
        bool have = j + 3 * TY < j1;
        if (have) {
#pragma unroll
            for (int u = 0; u < 4; u++) z[u] = *reinterpret_cast<const double2*>(col + (size_t)(j + u * TY) * ldn);
        }
        while (have) {
            const int jn = j + 4 * TY;
            const bool have_n = jn + 3 * TY < j1;
            if (have_n) {
#pragma unroll
                for (int u = 0; u < 4; u++) zn[u] = *reinterpret_cast<const double2*>(col + (size_t)(jn + u * TY) * ldn);
            }
#pragma unroll
            for (int b = 0; b < NB; b++) {
                double p0 = 1.0, p1 = 1.0;
#pragma unroll
                for (int u = 0; u < 4; u++) { p0 *= fma(A0[b], z[u].x, 1.0); p1 *= fma(A1[b], z[u].y, 1.0); }
                const double prod = p0 * p1;
                if (prod <= 1e300) {
                    acc[b] -= log(prod);
                } else {                                              // overflow (or NaN): term by term
#pragma unroll
                    for (int u = 0; u < 4; u++) acc[b] -= log(fma(A0[b], z[u].x, 1.0)) + log(fma(A1[b], z[u].y, 1.0));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) z[u] = zn[u];
            j = jn; have = have_n;
        }
        for (; j < j1; j += TY) {
            const double2 zz = *reinterpret_cast<const double2*>(col + (size_t)j * ldn);
#pragma unroll
            for (int b = 0; b < NB; b++) acc[b] -= log(fma(A0[b], zz.x, 1.0)) + log(fma(A1[b], zz.y, 1.0));
        }
    }
    // per evaluation: the deterministic grid sum of grid_sum_finish (per-CTA partial, last CTA sums the partials in index order)
    const int t = threadIdx.y * TX + threadIdx.x, w = t >> 5, l = t & 31;
    const int bid = blockIdx.y * gridDim.x + blockIdx.x, nblocks = gridDim.x * gridDim.y;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        double v = warp_sum(acc[b]);
        __syncthreads();
        if (l == 0) red[w] = v;
        __syncthreads();
        if (w == 0) {
            v = (l < 8) ? red[l] : 0.0;
            v = warp_sum(v);
            if (l == 0) partials[(size_t)b * nblocks + bid] = v;
        }
    }
    if (t == 0) {
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == (unsigned)nblocks - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
#pragma unroll
        for (int b = 0; b < NB; b++) {
            double s = 0.0;
            for (int k = t; k < nblocks; k += 256) s += partials[(size_t)b * nblocks + k];
            s = warp_sum(s);
            __syncthreads();
            if (l == 0) red[w] = s;
            __syncthreads();
            if (w == 0) {
                s = (l < 8) ? red[l] : 0.0;
                s = warp_sum(s);
                if (l == 0) out[b] = s;
            }
        }
        if (t == 0) *counter = 0u;
    }
}

template <int FL, class T>
__global__ void __launch_bounds__(256) loglik_kernel(int n, int P, int ldn, int ncols, int cols_per_cta,
                                                     const T* __restrict__ zd, const double* __restrict__ X,
                                                     const double* __restrict__ beta,
                                                     const double* __restrict__ y, const double* __restrict__ rowc,
                                                     double sigma, double* __restrict__ partials,
                                                     unsigned int* __restrict__ counter, double* __restrict__ out) {
    __shared__ double red[32];
    __shared__ double stab[64];
    __shared__ bool is_last;
    const int TX = blockDim.x, TY = blockDim.y;
    {
        const int tt = threadIdx.y * TX + threadIdx.x;
        if (tt < 64) stab[tt] = GMB_EXP2_TAB[tt];
        __syncthreads();
    }
    const int i0 = 2 * (blockIdx.x * TX + threadIdx.x);
    const int j0 = blockIdx.y * cols_per_cta;
    const int j1 = min(j0 + cols_per_cta, ncols);
    const double c0 = (FL == 7 || FL == 8) ? (-1.0 * log(sigma) - 0.5 * log(2 * GMB_PI_FAMILY)) : 0.0;
    const double inv_sigma = (FL == 7 || FL == 8) ? 1.0 / sigma : 1.0;

    double acc = 0.0;
    if (i0 < n) {
        const bool two = (i0 + 1 < n);
        double xb0 = 0.0, xb1 = 0.0;
        for (int p = 0; p < P; p++) {                                // same order as xb_kernel
            const double b = beta[p];
            xb0 += X[i0 + (size_t)p * ldn] * b;
            if (two) xb1 += X[i0 + 1 + (size_t)p * ldn] * b;
        }
        RowTerm<FL> r0, r1;
        r0.init(xb0, y[i0], (FL == 1 || FL == 2) ? rowc[i0] : 0.0);
        r1.init(xb1, two ? y[i0 + 1] : 0.0, ((FL == 1 || FL == 2) && two) ? rowc[i0 + 1] : 0.0);
        if (!two) r1.mask = 0.0;
        if ((FL == 4 || FL == 5 || FL == 6) && !two) r1.y = -1.0;      // neither 0 nor 1: the generic binomial terms contribute nothing
        if ((FL == 2 || FL == 8) && !two) { r1.y = 1.0; r1.xb = 1.0; }  // finite dummy, discarded below
        const T* col = zd + i0;
        int j = j0 + threadIdx.y;
        // 4 independent 16-byte loads per group of columns, and the next group's loads are issued before the current group's
        // arithmetic (register double buffer): 128 bytes in flight per thread while the FP64 chains run
        double2 z[4], zn[4];
        bool have = j + 3 * TY < j1;
        if (have) {
#pragma unroll
            for (int u = 0; u < 4; u++) z[u] = ld2(col + (size_t)(j + u * TY) * ldn);
        }
        while (have) {
            const int jn = j + 4 * TY;
            const bool have_n = jn + 3 * TY < j1;
            if (have_n) {
#pragma unroll
                for (int u = 0; u < 4; u++) zn[u] = ld2(col + (size_t)(jn + u * TY) * ldn);
            }
            double prod = 1.0, a1 = 0.0;
            int kmax = -(1 << 30);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                ll_accum<FL>(r0, z[u].x, c0, inv_sigma, stab, acc, prod, kmax);
                ll_accum<FL>(r1, z[u].y, c0, inv_sigma, stab, a1, prod, kmax);
            }
            if (FL == 3) {
                if (kmax > LOGIT_K_GUARD) {
                    double d = 0.0;
#pragma unroll
                    for (int u = 0; u < 4; u++) d += ll_direct<FL>(r0, z[u].x) + ll_direct<FL>(r1, z[u].y);
                    acc += d;
                } else {
                    acc -= log(prod);
                }
            } else if (two) acc += a1;
#pragma unroll
            for (int u = 0; u < 4; u++) z[u] = zn[u];
            j = jn; have = have_n;
        }
        for (; j < j1; j += TY) {
            double2 z = ld2(col + (size_t)j * ldn);
            double prod = 1.0, a1 = 0.0;
            int kmax = -(1 << 30);
            ll_accum<FL>(r0, z.x, c0, inv_sigma, stab, acc, prod, kmax);
            ll_accum<FL>(r1, z.y, c0, inv_sigma, stab, a1, prod, kmax);
            if (FL == 3) {
                if (kmax > LOGIT_K_GUARD) acc += ll_direct<FL>(r0, z.x) + ll_direct<FL>(r1, z.y);
                else acc -= log(prod);
            } else if (two) acc += a1;
        }
    }
    grid_sum_finish(acc, TX, partials, counter, out, red, &is_last);
}

// ---------------------------------------------------------------------------------------------------
// K2 for poisson/log and gaussian/identity through ROW STATISTICS of the sample matrix.  With zd fixed,
//   poisson : sum_j [y eta_ij - exp(eta_ij) - lf(y)] = y (m xb_i + T_i) - exp(xb_i) S_i - m lf(y_i),   S_i = sum_j exp(zd_ij), T_i = sum_j zd_ij
//   gaussian: sum_j [c0 - (y - eta_ij)^2 / (2 s^2)]  = m c0 - [m d_i^2 - 2 d_i T_i + T2_i] / (2 s^2),  d_i = y_i - xb_i,    T2_i = sum_j zd_ij^2
// so one evaluation costs O(n) once (S, T) or (T, T2) are known; they are built by ONE stream over zd per sample matrix (rowstat_kernel,
// cached per (sample matrix, niter)) instead of one stream per evaluation.  Same sums as mcmlmodel.h:295-300, different order.
// ---------------------------------------------------------------------------------------------------
template <int FL, class T>
__global__ void __launch_bounds__(256) rowstat_kernel(int n, int ldn, int ncols, int cols_per_cta, const T* __restrict__ zd,
                                                      double* __restrict__ rowpart /* [gridDim.y][2][ldn] */) {
    extern __shared__ double sm[];           // [TY][2][2 TX]
    const int TX = blockDim.x, TY = blockDim.y;
    const int i0 = 2 * (blockIdx.x * TX + threadIdx.x);
    const int j0 = blockIdx.y * cols_per_cta, j1 = min(j0 + cols_per_cta, ncols);
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;       // (S or T2, T) of rows i0, i0 + 1
    if (i0 < ldn) {
        const T* col = zd + i0;
        for (int j = j0 + threadIdx.y; j < j1; j += TY) {
            const double2 z = ld2(col + (size_t)j * ldn);
            if (FL == 1) { a0 += exp(z.x); a1 += exp(z.y); }
            else { a0 = fma(z.x, z.x, a0); a1 = fma(z.y, z.y, a1); }
            b0 += z.x; b1 += z.y;
        }
    }
    double* s = sm + (size_t)threadIdx.y * 4 * TX;
    s[2 * threadIdx.x] = a0; s[2 * threadIdx.x + 1] = a1;
    s[2 * TX + 2 * threadIdx.x] = b0; s[2 * TX + 2 * threadIdx.x + 1] = b1;
    __syncthreads();
    if (threadIdx.y == 0 && i0 < ldn) {
        for (int k = 1; k < TY; k++) {
            const double* sk = sm + (size_t)k * 4 * TX;
            a0 += sk[2 * threadIdx.x]; a1 += sk[2 * threadIdx.x + 1]; b0 += sk[2 * TX + 2 * threadIdx.x]; b1 += sk[2 * TX + 2 * threadIdx.x + 1];
        }
        double* ra = rowpart + ((size_t)blockIdx.y * 2 + 0) * ldn, * rb = rowpart + ((size_t)blockIdx.y * 2 + 1) * ldn;
        ra[i0] = a0; rb[i0] = b0;
        if (i0 + 1 < ldn) { ra[i0 + 1] = a1; rb[i0 + 1] = b1; }
    }
}

template <int FL>
__global__ void __launch_bounds__(256) loglik_rowstat_kernel(int n, int P, int ldn, double ncols, const double* __restrict__ statA,
                                                             const double* __restrict__ statT, const double* __restrict__ X,
                                                             const double* __restrict__ beta, const double* __restrict__ y,
                                                             const double* __restrict__ rowc, double sigma, double* __restrict__ partials,
                                                             unsigned int* __restrict__ counter, double* __restrict__ out) {
    __shared__ double red[32];
    __shared__ bool is_last;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    if (i < n) {
        double xb = 0.0;
        for (int p = 0; p < P; p++) xb += X[i + (size_t)p * ldn] * beta[p];
        if (FL == 1) {
            acc = y[i] * (ncols * xb + statT[i]) - exp(xb) * statA[i] - ncols * rowc[i];
        } else {
            const double c0 = -1.0 * log(sigma) - 0.5 * log(2 * GMB_PI_FAMILY);
            const double d = y[i] - xb;
            acc = ncols * c0 - 0.5 * (ncols * d * d - 2.0 * d * statT[i] + statA[i]) / (sigma * sigma);
        }
    }
    grid_sum_finish(acc, blockDim.x, partials, counter, out, red, &is_last);
}

// ---------------------------------------------------------------------------------------------------
// K3 pass 1.  CTA = 8 warps, row tile of 256 rows (lane owns rows 64k + 2 lane + {0,1}, k = 0..3), warp w walks
// columns j0 + w, j0 + w + 8, ...  Per element: IRLS weight w_ij, score term wu_ij, residual r_ij.
// Emits per-(column-chunk) row partial sums and per-(row-tile) column sums of r and r^2.
// ---------------------------------------------------------------------------------------------------
template <int FL>
__device__ __forceinline__ void mcnr_terms(double y, double eta, double inv_phi, const double* __restrict__ tab, double& w, double& wu, double& r) {
    if (FL == 1) {               // poisson/log : dhdmu = exp(-eta), detadmu = exp(-eta)
        double mu = dev_exp_tab(eta, tab);
        r = y - mu; w = mu; wu = r;
    } else if (FL == 3) {        // binomial/logit : dhdmu = detadmu = 1/(p(1-p)); p = e/(1+e) = 1 - 1/(1+e)
        const double e = dev_exp_tab(eta, tab);
        const double rc = dev_rcp_fast(1.0 + e);
        const double p = e * rc;
        r = y - p; w = p * rc; wu = r;                   // p (1 - p) = e / (1 + e)^2
    } else if (FL == 7) {        // gaussian/identity : W = 1/sigma^2
        r = y - eta; w = inv_phi; wu = inv_phi * r;
    } else if (FL == 2) {        // poisson/identity : dhdmu = eta, detadmu = 1
        r = y - eta; w = 1.0 / eta; wu = w * r;
    } else if (FL == 4) {        // binomial/log : dhdmu = (1 - p)/p with p = e^eta, detadmu = e^-eta
        const double p = exp(eta);
        r = y - p; w = 1.0 / ((1.0 - p) / p); wu = w * exp(-1.0 * eta) * r;
    } else if (FL == 5) {        // binomial/identity : dhdmu = eta (1 - eta), detadmu = 1
        r = y - eta; w = 1.0 / (eta * (1.0 - eta)); wu = w * r;
    } else {                     // gaussian/log (8) : dhdmu = 1 (gaussian), W = 1/sigma^2, h^-1 = e^eta, detadmu = e^-eta
        r = y - exp(eta); w = inv_phi; wu = w * exp(-1.0 * eta) * r;
    }
}

constexpr int MCNR_STAGES = 4;
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem) : "memory");
}

// FACTOR (binomial/logit only): zd points to the factor matrix F = exp(s zd) (see loglik_logit_factor_kernel) and the row constant
// is A_i = exp(s_i xb_i): 1/(1 + A F) is 1 - p or p, so the element costs a Newton reciprocal instead of an exp and a division.
template <int FL, bool FACTOR>
__global__ void __launch_bounds__(256) mcnr_pass1_kernel(int n, int ldn, int ncols, int cols_per_cta,
                                                         const double* __restrict__ zd, const double* __restrict__ xb,
                                                         const double* __restrict__ y, double inv_phi,
                                                         double* __restrict__ rowpart /* [gridDim.y][2][ldn] */,
                                                         double* __restrict__ colpart /* [gridDim.x][2][ncols] */) {
    extern __shared__ __align__(16) double sm[];   // [8 warps][MCNR_STAGES][256] column rings, then [8 warps][2][256] for the row reduction
    __shared__ double stab[64];
    if (threadIdx.x < 64) stab[threadIdx.x] = GMB_EXP2_TAB[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rbase = blockIdx.x * 256;
    const int j0 = blockIdx.y * cols_per_cta, j1 = min(j0 + cols_per_cta, ncols);

    double xbr[8], yr[8], wacc[8], sacc[8];
    bool ok[8];
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int v = 0; v < 2; v++) {
            int i = rbase + 64 * k + 2 * lane + v;
            ok[2 * k + v] = i < n;
            xbr[2 * k + v] = ok[2 * k + v] ? xb[i] : 0.0;
            yr[2 * k + v] = ok[2 * k + v] ? y[i] : 0.0;
            if (FACTOR) xbr[2 * k + v] = (yr[2 * k + v] == 1.0) ? exp(-1.0 * xbr[2 * k + v]) : exp(xbr[2 * k + v]);   // A_i
            wacc[2 * k + v] = 0.0; sacc[2 * k + v] = 0.0;
        }
    // Shared-memory staging: every warp keeps a private ring of MCNR_STAGES columns (256 rows = 2 KB each) filled with cp.async, three columns
    // ahead of the arithmetic — 12 KB in flight per warp, 96 KB per CTA, against 2 KB per warp with the register double buffer this replaces
    // (profiles: 41 % of the stall samples were the first use of a loaded value, 41 % of the HBM peak).  A lane reads back exactly the 16-byte
    // chunks it copied itself, so completion of its own cp.async groups is all the synchronisation the ring needs.
    double2 z[4];
    bool inrow[4];
#pragma unroll
    for (int k = 0; k < 4; k++) inrow[k] = rbase + 64 * k + 2 * lane < ldn;
    const bool partial_tile = rbase + 256 > n;
    double* ring = sm + (size_t)warp * MCNR_STAGES * 256;
    auto issue = [&](int jc, int stage) {
        if (jc < j1) {
            const double* col = zd + (size_t)jc * ldn + rbase + 2 * lane;
            double* dst = ring + stage * 256 + 2 * lane;
#pragma unroll
            for (int k = 0; k < 4; k++) if (inrow[k]) cp_async16(dst + 64 * k, col + 64 * k);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int s = 0; s < MCNR_STAGES - 1; s++) issue(j0 + warp + 8 * s, s);
    int it = 0;
    for (int j = j0 + warp; j < j1; j += 8, it++) {
        issue(j + 8 * (MCNR_STAGES - 1), (it + MCNR_STAGES - 1) % MCNR_STAGES);
        asm volatile("cp.async.wait_group %0;" :: "n"(MCNR_STAGES - 1) : "memory");
        {
            const double* src = ring + (it % MCNR_STAGES) * 256 + 2 * lane;
#pragma unroll
            for (int k = 0; k < 4; k++) z[k] = inrow[k] ? *reinterpret_cast<const double2*>(src + 64 * k) : make_double2(0.0, 0.0);
        }
        // straight-line arithmetic over the lane's 8 rows (no per-element branch: the dependent chains of the 8 elements overlap); rows beyond
        // n exist only in the last row tile and are zeroed by one warp-uniform branch
        double w[8], wu[8], r[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const double ze = (e & 1) ? z[e >> 1].y : z[e >> 1].x;
            if (FACTOR) {
                const double rc = dev_rcp_fast(fma(xbr[e], ze, 1.0));                               // 1/(1 + A F): p (y = 1) or 1 - p
                const double p = (yr[e] == 1.0) ? rc : 1.0 - rc;
                r[e] = yr[e] - p; w[e] = fma(-rc, rc, rc); wu[e] = r[e];                            // p (1 - p) = rc - rc^2
            } else {
                mcnr_terms<FL>(yr[e], xbr[e] + ze, inv_phi, stab, w[e], wu[e], r[e]);
            }
        }
        if (partial_tile) {
#pragma unroll
            for (int e = 0; e < 8; e++) if (!ok[e]) { w[e] = 0.0; wu[e] = 0.0; r[e] = 0.0; }
        }
        double sr = 0.0, sr2 = 0.0;
#pragma unroll
        for (int e = 0; e < 8; e++) { wacc[e] += w[e]; sacc[e] += wu[e]; sr += r[e]; sr2 += r[e] * r[e]; }
        sr = warp_sum(sr); sr2 = warp_sum(sr2);
        if (lane == 0) {
            colpart[((size_t)blockIdx.x * 2 + 0) * ncols + j] = sr;
            colpart[((size_t)blockIdx.x * 2 + 1) * ncols + j] = sr2;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                   // the reduction buffer below aliases the rings
    // cross-warp reduction of the row accumulators
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int v = 0; v < 2; v++) {
            int r = 64 * k + 2 * lane + v;
            sm[(warp * 2 + 0) * 256 + r] = wacc[2 * k + v];
            sm[(warp * 2 + 1) * 256 + r] = sacc[2 * k + v];
        }
    __syncthreads();
    {
        int r = threadIdx.x;   // 256 threads <-> 256 rows
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) { a += sm[(w * 2 + 0) * 256 + r]; b += sm[(w * 2 + 1) * 256 + r]; }
        int i = rbase + r;
        if (i < n) {
            rowpart[((size_t)blockIdx.y * 2 + 0) * ldn + i] = a;
            rowpart[((size_t)blockIdx.y * 2 + 1) * ldn + i] = b;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K3 pass 1, TMA version (poisson/log, binomial/logit on the factor matrix, gaussian/identity).
// The cp.async version above issues ~37 instructions per element (per-lane copies with their predicates and addresses, selects, a 10-shuffle
// warp reduction per column): it is ISSUE bound at 55-64 % of the HBM rate.  Here
//   * the sample columns arrive as TMA boxes of 256 rows x 8 columns (16 KB, one column per warp) in a ring of MCNR_TMA_STAGES stages: one
//     elected thread arms the stage's mbarrier and issues one cp.async.bulk.tensor per stage; rows beyond n and columns beyond the matrix are
//     zero-filled by the TMA unit, and the per-row constants are chosen so that zero-filled entries contribute exactly nothing (no masks);
//   * the arithmetic works on sums that need no per-element select: binomial t = 1/(1 + A F) (p or 1 - p by the sign folded into F and A),
//     u = 1 - t, w = t u, residual r = c + s u with per-row constants (c, s) — the row sums kept are sum(t u) and sum(u); poisson mu = A e^z
//     with A = e^xb hoisted, row sum kept is sum(mu); gaussian r = d - z, row sum kept is sum(z);
//   * the per-column sums of r and r^2 of FOUR columns are reduced together: an 8-value fold (16, 8, 4 lanes) followed by two butterfly steps —
//     9 shuffles for 4 columns instead of 40.
// Outputs as mcnr_pass1_kernel: rowpart [gridDim.y][2][ldn] (sum_j w, sum_j wu) and colpart [gridDim.x][2][ncols] (sum_i r, sum_i r^2).
// ---------------------------------------------------------------------------------------------------
constexpr int MCNR_TMA_STAGES = 6;
constexpr int MCNR_TMA_COLS = 8;                 // columns per stage = warps per CTA
constexpr size_t MCNR_TMA_SMEM = (size_t)MCNR_TMA_STAGES * 256 * MCNR_TMA_COLS * sizeof(double) + 1024 + 128;

// cubic-convergence reciprocal for d >= 1: hardware seed (about 20 bits) and one third-order step, error ~ e^3 (<= 1 ulp)
__device__ __forceinline__ double dev_rcp_cubic(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);
    const double t = fma(e, e, e);
    return fma(y, t, y);
}

template <int FL, class T>
__global__ void __launch_bounds__(256, 2) mcnr_tma_kernel(const __grid_constant__ CUtensorMap tm, int n, int ldn, int ncols, int cols_per_cta,
                                                          const double* __restrict__ xb, const double* __restrict__ y, double inv_phi,
                                                          double* __restrict__ rowpart, double* __restrict__ colpart) {
    extern __shared__ unsigned char smraw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~(uintptr_t)1023);
    T* ring = reinterpret_cast<T*>(base);                                            // [stage][col][256]
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)MCNR_TMA_STAGES * 256 * MCNR_TMA_COLS * sizeof(double));
    uint64_t* empty = full + MCNR_TMA_STAGES;
    __shared__ double stab[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 64) stab[tid] = GMB_EXP2_TAB[tid];
    if (tid == 0) {
        for (int s = 0; s < MCNR_TMA_STAGES; s++) { gmbtma::mbar_init(&full[s], 1); gmbtma::mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int rbase = blockIdx.x * 256;
    const int j0 = blockIdx.y * cols_per_cta, j1 = min(j0 + cols_per_cta, ncols);
    const int ntile = (j1 - j0 + MCNR_TMA_COLS - 1) / MCNR_TMA_COLS;

    // per-row constants of the lane's 8 rows (rbase + lane + 32 k); rows beyond n get constants that make a zero-filled entry contribute nothing
    double ca[8], cb[8], cs[8];      // FL 3: A, c, s ; FL 1: A = e^xb, y ; FL 7: d = y - xb
    double acc0[8], acc1[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int i = rbase + lane + 32 * k;
        const bool ok = i < n;
        const double xbi = ok ? xb[i] : 0.0, yi = ok ? y[i] : 0.0;
        if (FL == 3) {
            ca[k] = ok ? ((yi == 1.0) ? exp(-1.0 * xbi) : exp(xbi)) : 0.0;          // F is zero-filled there: t = 1, u = 0
            cs[k] = (yi == 1.0) ? 1.0 : -1.0;
            cb[k] = (yi == 1.0) ? 0.0 : yi;                                           // r = y - p: y = 1: u ; otherwise y - u
        } else if (FL == 1) {
            ca[k] = ok ? exp(xbi) : 0.0; cb[k] = yi; cs[k] = 0.0;
        } else {
            ca[k] = yi - xbi; cb[k] = 0.0; cs[k] = 0.0;
        }
        acc0[k] = 0.0; acc1[k] = 0.0;
    }
    auto produce = [&](int t) {
        const int s = t % MCNR_TMA_STAGES;
        if (t >= MCNR_TMA_STAGES) gmbtma::mbar_wait(&empty[s], ((t / MCNR_TMA_STAGES) - 1) & 1);
        gmbtma::mbar_expect_tx(&full[s], 256 * MCNR_TMA_COLS * sizeof(T));
        gmbtma::tma_load_2d(ring + (size_t)s * 256 * MCNR_TMA_COLS, &tm, &full[s], rbase, j0 + t * MCNR_TMA_COLS);
    };
    if (tid == 0) for (int t = 0; t < MCNR_TMA_STAGES - 1 && t < ntile; t++) produce(t);

    // (sum r, sum r^2) of 4 columns of this warp (4 successive stages) are reduced together: value index q = 2 * slot + (0: sum r, 1: sum r^2)
    for (int t4 = 0; t4 < ntile; t4 += 4) {
        double cv[8];
        int cols[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            cv[2 * q] = 0.0; cv[2 * q + 1] = 0.0; cols[q] = -1;
            const int t = t4 + q;
            if (t >= ntile) continue;                                                // block-uniform
            const int s = t % MCNR_TMA_STAGES;
            if (tid == 0 && t + MCNR_TMA_STAGES - 1 < ntile) produce(t + MCNR_TMA_STAGES - 1);
            __syncwarp();
            gmbtma::mbar_wait(&full[s], (t / MCNR_TMA_STAGES) & 1);
            const int j = j0 + t * MCNR_TMA_COLS + warp;
            if (j < j1) {                                                            // warp-uniform
                const T* col = ring + ((size_t)s * MCNR_TMA_COLS + warp) * 256 + lane;
                double z[8];
#pragma unroll
                for (int k = 0; k < 8; k++) z[k] = (double)col[32 * k];
                double sr = 0.0, sr2 = 0.0;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    double r;
                    if (FL == 3) {
                        const double tt = dev_rcp_cubic(fma(ca[k], z[k], 1.0));       // p (y = 1) or 1 - p
                        const double u = 1.0 - tt;
                        acc0[k] = fma(tt, u, acc0[k]);                                // sum_j p (1 - p)
                        acc1[k] += u;
                        r = fma(cs[k], u, cb[k]);
                    } else if (FL == 1) {
                        const double mu = ca[k] * dev_exp_tab(z[k], stab);
                        acc0[k] += mu;
                        r = cb[k] - mu;
                    } else {
                        acc0[k] += z[k];
                        r = ca[k] - z[k];
                    }
                    sr += r; sr2 = fma(r, r, sr2);
                }
                cv[2 * q] = sr; cv[2 * q + 1] = sr2; cols[q] = j;
            }
            __syncwarp();
            if (lane == 0) gmbtma::mbar_arrive(&empty[s]);
        }
        // fold the 8 values over the warp: 16-, 8- and 4-lane halves keep one half of the values each, then two butterfly steps
        double v4[4], v2[2], v1;
        const bool hi16 = lane & 16, hi8 = lane & 8, hi4 = lane & 4;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const double send = hi16 ? cv[q] : cv[q + 4], keep = hi16 ? cv[q + 4] : cv[q];
            v4[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const double send = hi8 ? v4[q] : v4[q + 2], keep = hi8 ? v4[q + 2] : v4[q];
            v2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        {
            const double send = hi4 ? v2[0] : v2[1], keep = hi4 ? v2[1] : v2[0];
            v1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
        v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
        // lane l holds the total of value index 4 * bit4(l) + 2 * bit3(l) + bit2(l)
        const int qi = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        int cj = cols[0];
        if ((qi >> 1) == 1) cj = cols[1]; else if ((qi >> 1) == 2) cj = cols[2]; else if ((qi >> 1) == 3) cj = cols[3];
        if ((lane & 3) == 0 && cj >= 0) colpart[((size_t)blockIdx.x * 2 + (qi & 1)) * ncols + cj] = v1;
    }
    __syncthreads();                                   // every warp is done with the ring: reuse it for the cross-warp row reduction
    double* red = reinterpret_cast<double*>(base);     // [8 warps][2][256]
    const double nc = (double)(j1 - j0);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int r = lane + 32 * k;
        red[(warp * 2 + 0) * 256 + r] = acc0[k];
        red[(warp * 2 + 1) * 256 + r] = acc1[k];
    }
    __syncthreads();
    {
        const int r = tid, i = rbase + r;              // 256 threads <-> 256 rows
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) { a += red[(w * 2 + 0) * 256 + r]; b += red[(w * 2 + 1) * 256 + r]; }
        if (i < n) {
            const double xbi = xb[i], yi = y[i];
            double wsum, ssum;
            if (FL == 3) { wsum = a; ssum = (yi == 1.0) ? b : fma(nc, yi, -b); }     // sum_j (y - p)
            else if (FL == 1) { wsum = a; ssum = fma(nc, yi, -a); }                  // W = mu, Wu = y - mu
            else { wsum = nc * inv_phi; ssum = inv_phi * fma(nc, yi - xbi, -a); }    // W = 1/sigma^2, Wu = (y - eta)/sigma^2
            rowpart[((size_t)blockIdx.y * 2 + 0) * ldn + i] = wsum;
            rowpart[((size_t)blockIdx.y * 2 + 1) * ldn + i] = ssum;
        }
    }
}

// K3 tail in ONE launch.  CTAs [0, 8 RT): CTA (tile, sub) reduces the row partials of 32 rows of a 256-row tile over the column chunks (lane = row,
// the 8 warps stride over the chunks, fixed-order combination through shared memory) and assembles those rows' share of X' diag(w) X and X' s
// (one warp per output entry).  CTAs [8 RT, 8 RT + NSIG): 32 columns each per pass (lane = column, warps stride over the row tiles) -> their share
// of sum_j sd(resid_j) (mcmloptim.h:216).  The last CTA to finish adds the per-CTA partials in index order -> out [P*P + P + 1].  Deterministic.
__global__ void __launch_bounds__(256) mcnr_tail_kernel(int n, int nobs, int P, int ldn, int ncols, int RT, int RTC, int CC, int NSIG, const double* __restrict__ X,
                                                        const double* __restrict__ rowpart, const double* __restrict__ colpart,
                                                        double* __restrict__ part /* [8 RT][P*P+P] then [NSIG] */, unsigned int* __restrict__ counter,
                                                        double* __restrict__ out) {
    __shared__ double sa[8][32], sb[8][32];
    __shared__ double sw[32], ss[32];
    __shared__ bool is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nout = P * P + P, NR = 8 * RT;
    if ((int)blockIdx.x < NR) {
        const int i0 = ((int)blockIdx.x >> 3) * 256 + ((int)blockIdx.x & 7) * 32, i = i0 + lane;
        double a = 0.0, b = 0.0;
        if (i < n) for (int c = warp; c < CC; c += 8) { a += rowpart[((size_t)c * 2 + 0) * ldn + i]; b += rowpart[((size_t)c * 2 + 1) * ldn + i]; }
        sa[warp][lane] = a; sb[warp][lane] = b;
        __syncthreads();
        if (warp == 0) {
            double ta = 0.0, tb = 0.0;
#pragma unroll
            for (int w = 0; w < 8; w++) { ta += sa[w][lane]; tb += sb[w][lane]; }
            sw[lane] = ta; ss[lane] = tb;
        }
        __syncthreads();
        double* dst = part + (size_t)blockIdx.x * nout;
        const double wi = sw[lane], si = ss[lane];
        for (int e = warp; e < nout; e += 8) {
            double acc = 0.0;
            if (i < n) {
                if (e < P * P) acc = X[i + (size_t)(e % P) * ldn] * wi * X[i + (size_t)(e / P) * ldn];
                else acc = X[i + (size_t)(e - P * P) * ldn] * si;
            }
            acc = warp_sum(acc);
            if (lane == 0) dst[e] = acc;
        }
    } else {
        const int b = blockIdx.x - NR;
        double tot = 0.0;                                   // meaningful in warp 0
        for (int j0 = b * 32; j0 < ncols; j0 += NSIG * 32) {
            const int j = j0 + lane;
            double sr = 0.0, sr2 = 0.0;
            if (j < ncols) for (int t = warp; t < RTC; t += 8) { sr += colpart[((size_t)t * 2 + 0) * ncols + j]; sr2 += colpart[((size_t)t * 2 + 1) * ncols + j]; }
            __syncthreads();
            sa[warp][lane] = sr; sb[warp][lane] = sr2;
            __syncthreads();
            if (warp == 0 && j < ncols) {
                double r1 = 0.0, r2 = 0.0;
#pragma unroll
                for (int w = 0; w < 8; w++) { r1 += sa[w][lane]; r2 += sb[w][lane]; }
                const double mean = r1 / nobs;                 // nobs = observations (n, or more than n rows when the rows are aggregated)
                const double q = r2 - nobs * mean * mean;
                tot += sqrt(fmax(q, 0.0) / (nobs - 1));
            }
        }
        if (warp == 0) {
            tot = warp_sum(tot);
            if (lane == 0) part[(size_t)NR * nout + b] = tot;
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int e = tid; e < nout; e += 256) {
            double s = 0.0;
            for (int t = 0; t < NR; t++) s += part[(size_t)t * nout + e];
            out[e] = s;
        }
        if (tid == 0) {
            double s = 0.0;
            for (int b = 0; b < NSIG; b++) s += part[(size_t)NR * nout + b];
            out[nout] = s;
            *counter = 0u;
        }
    }
}

// K3 pass 2a: rows — sum the column-chunk partials;  2b: columns — sigma_j = sd(resid_j) (mcmloptim.h:216), summed.
__global__ void __launch_bounds__(1024) mcnr_rows_kernel(int n, int ldn, int nchunks, const double* __restrict__ rowpart,
                                                         double* __restrict__ wsum, double* __restrict__ ssum) {
    // block (32 rows, 32 chunk lanes): thread (tx, ty) adds chunks ty, ty + 32, ... of row tx; thread (tx, 0) then adds the 32
    // lane sums in order — a fixed order of addition, ~nchunks/32 dependent loads deep instead of nchunks
    __shared__ double sa[32][33], sb[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.x * 32 + tx;
    double a = 0.0, b = 0.0;
    if (i < n) {
        for (int c = ty; c < nchunks; c += 32) {
            a += rowpart[((size_t)c * 2 + 0) * ldn + i];
            b += rowpart[((size_t)c * 2 + 1) * ldn + i];
        }
    }
    sa[ty][tx] = a; sb[ty][tx] = b;
    __syncthreads();
    if (ty == 0 && i < n) {
        double ta = 0.0, tb = 0.0;
#pragma unroll 8
        for (int k = 0; k < 32; k++) { ta += sa[k][tx]; tb += sb[k][tx]; }
        wsum[i] = ta; ssum[i] = tb;
    }
}

__global__ void __launch_bounds__(256) mcnr_sigma_kernel(int n, int ncols, int ntiles, const double* __restrict__ colpart,
                                                         double* __restrict__ partial /* [gridDim.x] */) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ncols; j += gridDim.x * blockDim.x) {
        double sr = 0.0, sr2 = 0.0;
        for (int t = 0; t < ntiles; t++) {
            sr += colpart[((size_t)t * 2 + 0) * ncols + j];
            sr2 += colpart[((size_t)t * 2 + 1) * ncols + j];
        }
        double mean = sr / n;
        double ss = sr2 - n * mean * mean;          // sum (r - mean)^2
        acc += sqrt(fmax(ss, 0.0) / (n - 1));
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// K3 pass 3: out[a + b P] = sum_i X_ia wsum_i X_ib ; out[P*P + a] = sum_i X_ia ssum_i ; out[P*P+P] = sum_j sigma_j.
__global__ void __launch_bounds__(256) mcnr_assemble_kernel(int n, int P, int ldn, const double* __restrict__ X,
                                                            const double* __restrict__ wsum, const double* __restrict__ ssum,
                                                            const double* __restrict__ sigpart, int nsig,
                                                            double* __restrict__ out) {
    __shared__ double red[32];
    int e = blockIdx.x;
    double acc = 0.0;
    if (e < P * P) {
        int a = e % P, b = e / P;
        for (int i = threadIdx.x; i < n; i += blockDim.x) acc += X[i + (size_t)a * ldn] * wsum[i] * X[i + (size_t)b * ldn];
    } else if (e < P * P + P) {
        int a = e - P * P;
        for (int i = threadIdx.x; i < n; i += blockDim.x) acc += X[i + (size_t)a * ldn] * ssum[i];
    } else {
        for (int i = threadIdx.x; i < nsig; i += blockDim.x) acc += sigpart[i];
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[e] = acc;
}

// ---------------------------------------------------------------------------------------------------
// E-step on AGGREGATED rows.  Observations that share their row of [X | Z] share eta = x'beta + z'u for every beta and u, so
// zd is formed for the ng distinct rows only (model.cu) and a row carries the sufficient statistics of its c observations (aggregate.cu):
//   log-likelihood   sum_i l(y_i, eta) = dev_family_ll_w(c, sum y | mean y, within-row SS, sum lf(y); eta)
//   MCNR             W is a function of eta alone: sum_i w_i = c w;  the working residual is linear in y: sum_i wu_i = c wu(ybar);
//                    sum_i r_i = c (ybar - mu),  sum_i r_i^2 = SS_within + c (ybar - mu)^2
// — the sums of mcmlmodel.h:284-304 / mcmloptim.h:190-231 in a different order, on n / ng times fewer bytes (config C2: 500 -> 50 rows).
// One warp spans 32 rows, the 8 warps of a CTA take every 8th column of the CTA's column chunk.
// ---------------------------------------------------------------------------------------------------
template <int FL, class T>
__global__ void __launch_bounds__(256) loglik_agg_kernel(int ng, int P, int ldg, int ncols, int cols_per_cta, int CC, const T* __restrict__ zd,
                                                         const double* __restrict__ Xg, const double* __restrict__ beta /* P x gridDim.y */,
                                                         const double* __restrict__ lc, const double* __restrict__ lys, const double* __restrict__ lsq,
                                                         const double* __restrict__ lrc, double var_par, double* __restrict__ partials /* [gridDim.y][gridDim.x] */) {
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rt = blockIdx.x / CC, cc = blockIdx.x % CC;
    const int g = rt * 32 + lane;
    const int j0 = cc * cols_per_cta, j1 = min(j0 + cols_per_cta, ncols);
    const double* b = beta + (size_t)blockIdx.y * P;
    double acc = 0.0;
    if (g < ng) {
        double xb = 0.0;
        for (int p = 0; p < P; p++) xb += Xg[g + (size_t)p * ldg] * b[p];          // same order as xb_kernel
        const double c = lc[g], ys = lys[g], sq = (FL == 7) ? lsq[g] : 0.0, rc = (FL == 1) ? lrc[g] : 0.0;
        const double c0 = (FL == 7) ? (-1.0 * log(var_par) - 0.5 * log(2 * GMB_PI_FAMILY)) : 0.0;   // as loglik_kernel
        const T* col = zd + g;
        for (int j = j0 + warp; j < j1; j += 8)
            acc += dev_family_ll_w<FL>(c, ys, sq, rc, xb + (double)col[(size_t)j * ldg], c0, var_par);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = acc;
}

// Binomial/logit on the aggregated rows, through the factor matrix F = exp(zd) (ng x m, built with zd): with c observations of which n1 are
// ones in a row,   sum_i l_i = n1 eta - c log(1 + e^eta),   e^eta = A F,  A = exp(x'beta), so
//     ll(beta) = sum_g n1_g (m xb_g + T_g) - sum_g c_g log prod_j (1 + A_g F_gj),        T_g = sum_j zd_gj (a row statistic of the sample matrix)
// — per element one FMA and one multiply per evaluation and one log per 8 columns, as loglik_logit_factor_multi_kernel but with the product
// running along a ROW (the weight c_g factors out of the log), on ng instead of n rows.  NB evaluations share each load; blockIdx.y = group of NB
// evaluations (a short last group overlaps the one before).  partials: [gridDim.y][NB][gridDim.x].
template <int NB, class T>
__global__ void __launch_bounds__(256, 2) loglik_logit_agg_kernel(int ng, int P, int ldg, int ncols, int cols_per_cta, int CC, int n_eval,
                                                               const T* __restrict__ F, const double* __restrict__ Xg, const double* __restrict__ beta,
                                                               const double* __restrict__ lc, const double* __restrict__ lys, const double* __restrict__ Tsum,
                                                               double* __restrict__ partials) {
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rt = blockIdx.x / CC, cc = blockIdx.x % CC;
    const int g = rt * 32 + lane;
    const int j0 = cc * cols_per_cta, j1 = min(j0 + cols_per_cta, ncols);
    const int e0 = min((int)blockIdx.y * NB, n_eval - NB);
    double A[NB], acc[NB];
#pragma unroll
    for (int e = 0; e < NB; e++) { A[e] = 0.0; acc[e] = 0.0; }
    if (g < ng) {
#pragma unroll
        for (int e = 0; e < NB; e++) {
            const double* b = beta + (size_t)(e0 + e) * P;
            double xb = 0.0;
            for (int p = 0; p < P; p++) xb += Xg[g + (size_t)p * ldg] * b[p];      // same order as xb_kernel
            A[e] = exp(xb);
        }
        const T* row = F + g;
        int j = j0 + warp;
        for (; j + 120 < j1; j += 128) {               // 16 of this warp's columns per group: one log per 16 factors
            double pr[NB];
#pragma unroll
            for (int e = 0; e < NB; e++) pr[e] = 1.0;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                double f[8];
#pragma unroll
                for (int u = 0; u < 8; u++) f[u] = (double)row[(size_t)(j + 8 * (8 * h + u)) * ldg];
#pragma unroll
                for (int e = 0; e < NB; e++) {
#pragma unroll
                    for (int u = 0; u < 8; u++) pr[e] *= fma(A[e], f[u], 1.0);
                }
            }
#pragma unroll
            for (int e = 0; e < NB; e++) {
                if (pr[e] <= 1e300) acc[e] += log(pr[e]);
                else {                                  // overflow (or NaN): term by term, from memory again
#pragma unroll 1
                    for (int u = 0; u < 16; u++) acc[e] += log(fma(A[e], (double)row[(size_t)(j + 8 * u) * ldg], 1.0));
                }
            }
        }
        for (; j < j1; j += 8) {
            const double f = (double)row[(size_t)j * ldg];
#pragma unroll
            for (int e = 0; e < NB; e++) acc[e] += log(fma(A[e], f, 1.0));
        }
        const double c = lc[g];
        const bool first = (cc == 0 && warp == 0);             // the linear term n1 (m xb + T) once per row
        const double n1 = first ? lys[g] : 0.0, ts = first ? Tsum[g] : 0.0;
#pragma unroll
        for (int e = 0; e < NB; e++) {
            double xb = 0.0;
            if (first) {
                const double* b = beta + (size_t)(e0 + e) * P;
                for (int p = 0; p < P; p++) xb += Xg[g + (size_t)p * ldg] * b[p];
            }
            acc[e] = n1 * ((double)ncols * xb + ts) - c * acc[e];
        }
    }
#pragma unroll
    for (int e = 0; e < NB; e++) {
        const double v = block_sum(acc[e], red);
        if (threadIdx.x == 0) partials[((size_t)blockIdx.y * NB + e) * gridDim.x + blockIdx.x] = v;
        __syncthreads();
    }
}

// out[min(z NB, n_eval - NB) + e] = sum of the nb partials of (group z, evaluation e); block index = z NB + e
__global__ void __launch_bounds__(256) agg_finish_groups_kernel(int nb, int NB, int n_eval, const double* __restrict__ partials, double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int k = threadIdx.x; k < nb; k += 256) s += partials[(size_t)blockIdx.x * nb + k];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[min((int)blockIdx.x / NB * NB, n_eval - NB) + (int)blockIdx.x % NB] = s;
}

// out[e] = sum of the nb partials of evaluation e, in a fixed order
__global__ void __launch_bounds__(256) agg_finish_kernel(int nb, const double* __restrict__ partials, double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int k = threadIdx.x; k < nb; k += 256) s += partials[(size_t)blockIdx.x * nb + k];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// MCNR pass 1 on the aggregated rows; writes the same partial sums as mcnr_pass1_kernel: rowpart[cc][2][ldg] (sum_j of c w and c wu per row),
// colpart[rt][2][ncols] (sum over the tile's rows of sum_i r_i and sum_i r_i^2 per column), rt = tiles of 32 rows
// FACTOR (binomial/logit): zd points to F = exp(zd) and e^eta = exp(xb) F — a multiply instead of the table exp
template <int FL, class T, bool FACTOR>
__global__ void __launch_bounds__(256) mcnr_agg_kernel(int ng, int ldg, int ncols, int cols_per_cta, const T* __restrict__ zd,
                                                       const double* __restrict__ xb /* per observation */, const int* __restrict__ rep,
                                                       const double* __restrict__ cnt, const double* __restrict__ eys, const double* __restrict__ ess,
                                                       double inv_phi, double* __restrict__ rowpart, double* __restrict__ colpart) {
    __shared__ double stab[64];
    __shared__ double sw[8][32], ss[8][32];
    if (threadIdx.x < 64) stab[threadIdx.x] = GMB_EXP2_TAB[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * 32 + lane;
    const int j0 = blockIdx.y * cols_per_cta, j1 = min(j0 + cols_per_cta, ncols);
    const bool ok = g < ng;
    const double c = ok ? cnt[g] : 0.0;
    const double ybar = ok ? eys[g] / c : 0.0, sq = ok ? ess[g] : 0.0, xbg = ok ? xb[rep[g]] : 0.0;
    const double Ag = FACTOR ? exp(xbg) : 0.0;
    const T* col = zd + (ok ? g : 0);
    double wacc = 0.0, sacc = 0.0;
    for (int j = j0 + warp; j < j1; j += 8) {
        double w = 0.0, wu = 0.0, r = 0.0;
        if (FACTOR) {
            if (ok) {                                    // p = e / (1 + e), p (1 - p) = e / (1 + e)^2 with e = A F (as mcnr_terms<3>)
                const double e = Ag * (double)col[(size_t)j * ldg];
                const double rc = dev_rcp_fast(1.0 + e);
                const double p = e * rc;
                r = ybar - p; w = p * rc; wu = r;
            }
        } else if (ok) mcnr_terms<FL>(ybar, xbg + (double)col[(size_t)j * ldg], inv_phi, stab, w, wu, r);
        wacc += w; sacc += wu;
        const double sr = warp_sum(c * r), sr2 = warp_sum(ok ? fma(c * r, r, sq) : 0.0);
        if (lane == 0) {
            colpart[((size_t)blockIdx.x * 2 + 0) * ncols + j] = sr;
            colpart[((size_t)blockIdx.x * 2 + 1) * ncols + j] = sr2;
        }
    }
    sw[warp][lane] = c * wacc; ss[warp][lane] = c * sacc;
    __syncthreads();
    if (warp == 0 && ok) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) { a += sw[k][lane]; b += ss[k][lane]; }
        rowpart[((size_t)blockIdx.y * 2 + 0) * ldg + g] = a;
        rowpart[((size_t)blockIdx.y * 2 + 1) * ldg + g] = b;
    }
}

// row statistics of the aggregated rows -> per observation (rows of one group share zd, hence their statistics)
__global__ void stat_expand_kernel(int n, int ldn, int ldg, const int* __restrict__ gid, const double* __restrict__ sg, double* __restrict__ st) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st[i] = sg[gid[i]];
    st[ldn + i] = sg[ldg + gid[i]];
}

}  // namespace

// 1 (default) = MCNR pass 1 through the TMA kernel where it applies; 0 = the cp.async kernel (GMB_MCNR_TMA=0; parity tests compare both)
static int g_mcnr_tma = [] { const char* e = getenv("GMB_MCNR_TMA"); return e ? atoi(e) : 1; }();

// 1 = poisson/gaussian evaluations go through the row statistics (default); 0 = always stream zd (roofline probes, parity tests of the stream)
static int g_rowstats = 1;
int gmb_estep_rowstats_enabled() { return g_rowstats; }
extern "C" int gmb_estep_set_rowstats(int on) { g_rowstats = on ? 1 : 0; return GMB_OK; }

int gmb_launch_xb(gmb_model* mdl, const double* d_beta, double* d_xb) {
    gmb_ctx* ctx = mdl->ctx;
    xb_kernel<<<(mdl->n + 255) / 256, 256, 0, ctx->stream>>>(mdl->n, mdl->P, mdl->ldn, mdl->dX, d_beta, d_xb);
    ctx->launches++;
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}

// row statistics of the model's zd over its first niter_local columns (poisson: S, T; gaussian: T2, T), cached
static int ensure_rowstats(gmb_model* mdl) {
    gmb_ctx* ctx = mdl->ctx;
    const int n = mdl->n, ldn = mdl->ldn, ncols = mdl->niter_local;
    if (mdl->stat_valid && mdl->stat_cols == ncols) return GMB_OK;
    if (!mdl->dstat) GMB_CUDA(gmb_dmalloc(ctx, &mdl->dstat, sizeof(double) * 2 * ldn));
    if (mdl->eagg) {
        // the statistics of the ng aggregated rows, then one copy per observation (the O(n) evaluation kernel is unchanged)
        const gmb_agg& a = mdl->agg;
        const int ng = a.ng, ldg = a.ldn, halfg = (ldg + 1) / 2;
        int TXg = 32; while (TXg < 128 && TXg < halfg) TXg <<= 1;
        const int TYg = 256 / TXg, RTg = (halfg + TXg - 1) / TXg;
        int CCg = (ctx->sms * 4 + RTg - 1) / RTg;
        const int max_ccg = (ncols + 4 * TYg - 1) / (4 * TYg);
        if (CCg > max_ccg) CCg = max_ccg;
        if (CCg < 1) CCg = 1;
        int cpc = (ncols + CCg - 1) / CCg;
        CCg = (ncols + cpc - 1) / cpc;
        GMB_TRY(gmb_ctx_scratch(ctx, (size_t)CCg * 2 * ldg + 2 * (size_t)ldg));
        double* rp = ctx->d_scratch; double* sg = rp + (size_t)CCg * 2 * ldg;
        dim3 gridg(RTg, CCg), blockg(TXg, TYg);
        const size_t smemg = sizeof(double) * 4 * TXg * TYg;
        if (mdl->prec == 32) {
            if (mdl->flink == 1) rowstat_kernel<1, float><<<gridg, blockg, smemg, ctx->stream>>>(ng, ldg, ncols, cpc, mdl->dzd32, rp);
            else rowstat_kernel<7, float><<<gridg, blockg, smemg, ctx->stream>>>(ng, ldg, ncols, cpc, mdl->dzd32, rp);
        } else {
            if (mdl->flink == 1) rowstat_kernel<1, double><<<gridg, blockg, smemg, ctx->stream>>>(ng, ldg, ncols, cpc, mdl->dzd, rp);
            else rowstat_kernel<7, double><<<gridg, blockg, smemg, ctx->stream>>>(ng, ldg, ncols, cpc, mdl->dzd, rp);
        }
        if (mdl->flink == 3)      // binomial/logit: only T_g = sum_j zd_gj is used, by the aggregated rows themselves (loglik_logit_agg_kernel): dstat = [T2 | T] x ldg
            mcnr_rows_kernel<<<(ng + 31) / 32, dim3(32, 32), 0, ctx->stream>>>(ng, ldg, CCg, rp, mdl->dstat, mdl->dstat + ldg);
        else {
            mcnr_rows_kernel<<<(ng + 31) / 32, dim3(32, 32), 0, ctx->stream>>>(ng, ldg, CCg, rp, sg, sg + ldg);
            stat_expand_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, ldn, ldg, a.dgid, sg, mdl->dstat);
        }
        ctx->launches += 3;
        GMB_CUDA(cudaGetLastError());
        mdl->stat_valid = true; mdl->stat_cols = ncols;
        return GMB_OK;
    }
    const int half = (ldn + 1) / 2;
    int TX = 32; while (TX < 128 && TX < half) TX <<= 1;
    const int TY = 256 / TX;
    const int RT = (half + TX - 1) / TX;
    int CC = (ctx->sms * 4 + RT - 1) / RT;
    const int max_cc = (ncols + 4 * TY - 1) / (4 * TY);
    if (CC > max_cc) CC = max_cc;
    if (CC < 1) CC = 1;
    int cols_per_cta = (ncols + CC - 1) / CC;
    CC = (ncols + cols_per_cta - 1) / cols_per_cta;
    GMB_TRY(gmb_ctx_scratch(ctx, (size_t)CC * 2 * ldn));
    double* rowpart = ctx->d_scratch;
    dim3 grid(RT, CC), block(TX, TY);
    const size_t smem = sizeof(double) * 4 * TX * TY;
    if (mdl->prec == 32) {
        if (mdl->flink == 1) rowstat_kernel<1, float><<<grid, block, smem, ctx->stream>>>(n, ldn, ncols, cols_per_cta, mdl->dzd32, rowpart);
        else rowstat_kernel<7, float><<<grid, block, smem, ctx->stream>>>(n, ldn, ncols, cols_per_cta, mdl->dzd32, rowpart);
    } else {
        if (mdl->flink == 1) rowstat_kernel<1, double><<<grid, block, smem, ctx->stream>>>(n, ldn, ncols, cols_per_cta, mdl->dzd, rowpart);
        else rowstat_kernel<7, double><<<grid, block, smem, ctx->stream>>>(n, ldn, ncols, cols_per_cta, mdl->dzd, rowpart);
    }
    mcnr_rows_kernel<<<(n + 31) / 32, dim3(32, 32), 0, ctx->stream>>>(n, ldn, CC, rowpart, mdl->dstat, mdl->dstat + ldn);
    ctx->launches += 2;
    GMB_CUDA(cudaGetLastError());
    mdl->stat_valid = true; mdl->stat_cols = ncols;
    return GMB_OK;
}

// fp32 mode: the whole sample matrix of the model (float storage), one evaluation
static int launch_loglik_f32(gmb_model* mdl, const double* d_beta, double var_par, int ncols, double* d_out) {
    gmb_ctx* ctx = mdl->ctx;
    const int n = mdl->n;
    if (ncols <= 0) { GMB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double), ctx->stream)); return GMB_OK; }
    int half = (n + 1) / 2;
    int TX = 32; while (TX < 256 && TX < half) TX <<= 1;
    int TY = 256 / TX;
    int RT = (half + TX - 1) / TX;
    int want_cc = (ctx->sms * 6 + RT - 1) / RT;
    int max_cc = (ncols + 4 * TY - 1) / (4 * TY);
    int CC = want_cc < max_cc ? want_cc : max_cc; if (CC < 1) CC = 1;
    int cols_per_cta = (ncols + CC - 1) / CC;
    cols_per_cta = round_up(cols_per_cta, TY);
    CC = (ncols + cols_per_cta - 1) / cols_per_cta;
    GMB_TRY(gmb_ctx_scratch(ctx, (size_t)RT * CC));
    double* partials = ctx->d_scratch;
    unsigned int* counter = ctx->d_counter;
    dim3 grid(RT, CC), block(TX, TY);
    if (mdl->flink == 3 && mdl->f_valid)
        loglik_logit_factor_kernel<float><<<grid, block, 0, ctx->stream>>>(n, mdl->P, mdl->ldn, ncols, cols_per_cta, mdl->dF32, mdl->dX, d_beta, mdl->dy, partials, counter, d_out);
    else switch (mdl->flink) {
    case 1: loglik_kernel<1, float><<<grid, block, 0, ctx->stream>>>(n, mdl->P, mdl->ldn, ncols, cols_per_cta, mdl->dzd32, mdl->dX, d_beta, mdl->dy, mdl->drowc, var_par, partials, counter, d_out); break;
    case 3: loglik_kernel<3, float><<<grid, block, 0, ctx->stream>>>(n, mdl->P, mdl->ldn, ncols, cols_per_cta, mdl->dzd32, mdl->dX, d_beta, mdl->dy, mdl->drowc, var_par, partials, counter, d_out); break;
    case 7: loglik_kernel<7, float><<<grid, block, 0, ctx->stream>>>(n, mdl->P, mdl->ldn, ncols, cols_per_cta, mdl->dzd32, mdl->dX, d_beta, mdl->dy, mdl->drowc, var_par, partials, counter, d_out); break;
    default: return gmb_set_error(GMB_EFAMILY, "fp32 mode: family/link code %d is not implemented (poisson/log, binomial/logit, gaussian/identity)", mdl->flink);
    }
    ctx->launches++;
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}

// n_eval evaluations (d_beta: P x n_eval, one var_par) of the whole-model objective on the aggregated rows: two launches
static int launch_loglik_agg(gmb_model* mdl, const double* d_beta, int n_eval, double var_par, double* d_out) {
    gmb_ctx* ctx = mdl->ctx;
    const gmb_agg& a = mdl->agg;
    const int ng = a.ng, ldg = a.ldn, ncols = mdl->niter_local;
    if (ncols <= 0) { GMB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * n_eval, ctx->stream)); return GMB_OK; }
    const int RT = (ng + 31) / 32;
    if (mdl->flink == 3 && mdl->f_valid) {
        GMB_TRY(ensure_rowstats(mdl));                        // T_g
        const bool f32 = mdl->prec == 32;
        for (int e0 = 0; e0 < n_eval;) {
            // batches of >= 8 evaluations: groups of 8 per pass over F; what is left (< 8): one evaluation per pass
            const int NBk = (n_eval - e0 >= GMB_LOGLIK_NB) ? GMB_LOGLIK_NB : 1;
            const int ne = NBk == 1 ? 1 : std::min(n_eval - e0, 8 * 32768);
            const int groups = (ne + NBk - 1) / NBk;
            int CC = (ctx->sms * 2) / (RT * groups);            // one full wave of 2 CTAs per SM (rounded down: no tail wave)
            const int max_cc = (ncols + 63) / 64;
            if (CC > max_cc) CC = max_cc;
            if (CC < 1) CC = 1;
            const int cpc = round_up((ncols + CC - 1) / CC, 128);      // whole groups of 16 columns per warp
            CC = (ncols + cpc - 1) / cpc;
            const int nb = RT * CC;
            GMB_TRY(gmb_ctx_scratch(ctx, (size_t)nb * groups * NBk));
            dim3 grid(nb, groups);
            const double* bp = d_beta + (size_t)e0 * mdl->P;
#define GMB_LFA(NBV, TT, FP) loglik_logit_agg_kernel<NBV, TT><<<grid, 256, 0, ctx->stream>>>(ng, mdl->P, ldg, ncols, cpc, CC, ne, FP, a.dX, bp, a.dlcnt, a.dlys, \
                                                                                           mdl->dstat + ldg, ctx->d_scratch)
            if (NBk == 1) { if (f32) GMB_LFA(1, float, mdl->dF32); else GMB_LFA(1, double, mdl->dF); }
            else { if (f32) GMB_LFA(GMB_LOGLIK_NB, float, mdl->dF32); else GMB_LFA(GMB_LOGLIK_NB, double, mdl->dF); }
#undef GMB_LFA
            agg_finish_groups_kernel<<<groups * NBk, 256, 0, ctx->stream>>>(nb, NBk, ne, ctx->d_scratch, d_out + e0);
            ctx->launches += 2;
            e0 += ne;
        }
        GMB_CUDA(cudaGetLastError());
        return GMB_OK;
    }
    for (int e0 = 0; e0 < n_eval; e0 += 32768) {
        const int ne = std::min(32768, n_eval - e0);
        int CC = (ctx->sms * 4 + RT * ne - 1) / (RT * ne);
        const int max_cc = (ncols + 31) / 32;
        if (CC > max_cc) CC = max_cc;
        if (CC < 1) CC = 1;
        int cpc = round_up((ncols + CC - 1) / CC, 8);
        CC = (ncols + cpc - 1) / cpc;
        const int nb = RT * CC;
        GMB_TRY(gmb_ctx_scratch(ctx, (size_t)nb * ne));
        dim3 grid(nb, ne);
#define GMB_LLA(F, TT, ZD) loglik_agg_kernel<F, TT><<<grid, 256, 0, ctx->stream>>>(ng, mdl->P, ldg, ncols, cpc, CC, ZD, a.dX, d_beta + (size_t)e0 * mdl->P, a.dlcnt, \
                                                                                  a.dlys, a.dlsq, a.dlrc, var_par, ctx->d_scratch)
        if (mdl->prec == 32) { if (mdl->flink == 1) GMB_LLA(1, float, mdl->dzd32); else if (mdl->flink == 3) GMB_LLA(3, float, mdl->dzd32); else GMB_LLA(7, float, mdl->dzd32); }
        else { if (mdl->flink == 1) GMB_LLA(1, double, mdl->dzd); else if (mdl->flink == 3) GMB_LLA(3, double, mdl->dzd); else GMB_LLA(7, double, mdl->dzd); }
#undef GMB_LLA
        agg_finish_kernel<<<ne, 256, 0, ctx->stream>>>(nb, ctx->d_scratch, d_out + e0);
        ctx->launches += 2;
    }
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}

int gmb_launch_loglik(gmb_model* mdl, const double* d_beta, double var_par, double* d_out) {
    gmb_ctx* ctx = mdl->ctx;
    if (mdl->eagg && !((mdl->flink == 1 || mdl->flink == 7) && mdl->niter_local > 0 && gmb_estep_rowstats_enabled()))
        return launch_loglik_agg(mdl, d_beta, 1, var_par, d_out);
    if ((mdl->flink == 1 || mdl->flink == 7) && mdl->niter_local > 0 && gmb_estep_rowstats_enabled()) {
        GMB_TRY(ensure_rowstats(mdl));
        const int n = mdl->n, nb = (n + 255) / 256;
        GMB_TRY(gmb_ctx_scratch(ctx, (size_t)nb));
        if (mdl->flink == 1)
            loglik_rowstat_kernel<1><<<nb, 256, 0, ctx->stream>>>(n, mdl->P, mdl->ldn, (double)mdl->niter_local, mdl->dstat, mdl->dstat + mdl->ldn, mdl->dX,
                                                                  d_beta, mdl->dy, mdl->drowc, var_par, ctx->d_scratch, ctx->d_counter, d_out);
        else
            loglik_rowstat_kernel<7><<<nb, 256, 0, ctx->stream>>>(n, mdl->P, mdl->ldn, (double)mdl->niter_local, mdl->dstat, mdl->dstat + mdl->ldn, mdl->dX,
                                                                  d_beta, mdl->dy, mdl->drowc, var_par, ctx->d_scratch, ctx->d_counter, d_out);
        ctx->launches++;
        GMB_CUDA(cudaGetLastError());
        return GMB_OK;
    }
    if (mdl->prec == 32) return launch_loglik_f32(mdl, d_beta, var_par, mdl->niter_local, d_out);
    return gmb_launch_loglik_cols(mdl, d_beta, var_par, mdl->dzd, mdl->niter_local, d_out);
}

// the same sum over an explicit block of columns of a zd-like matrix (ldn x ncols); the Laplace objectives use single columns
int gmb_launch_loglik_cols(gmb_model* mdl, const double* d_beta, double var_par, const double* d_zd, int ncols, double* d_out) {
    gmb_ctx* ctx = mdl->ctx;
    const int n = mdl->n;
    if (ncols <= 0) { GMB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double), ctx->stream)); return GMB_OK; }
    int half = (n + 1) / 2;
    int TX = 32; while (TX < 256 && TX < half) TX <<= 1;
    int TY = 256 / TX;
    int RT = (half + TX - 1) / TX;
    // enough CTAs for ~6 per SM, but keep >= 4*TY columns per CTA so the unrolled loop is used
    int want_cc = (ctx->sms * 6 + RT - 1) / RT;
    int max_cc = (ncols + 4 * TY - 1) / (4 * TY);
    int CC = want_cc < max_cc ? want_cc : max_cc; if (CC < 1) CC = 1;
    int cols_per_cta = (ncols + CC - 1) / CC;
    cols_per_cta = round_up(cols_per_cta, TY);
    CC = (ncols + cols_per_cta - 1) / cols_per_cta;
    size_t nblocks = (size_t)RT * CC;
    GMB_TRY(gmb_ctx_scratch(ctx, nblocks));
    double* partials = ctx->d_scratch;
    unsigned int* counter = ctx->d_counter;   // zeroed at ctx creation and re-zeroed by the last CTA
    dim3 grid(RT, CC), block(TX, TY);
    if (mdl->flink == 3 && mdl->f_valid && !mdl->eagg && d_zd >= mdl->dzd && d_zd < mdl->dzd + (size_t)mdl->ldn * mdl->m_cap) {
        const double* d_f = mdl->dF + (d_zd - mdl->dzd);              // the same block of columns of the factor matrix
        loglik_logit_factor_kernel<double><<<grid, block, 0, ctx->stream>>>(n, mdl->P, mdl->ldn, ncols, cols_per_cta, d_f, mdl->dX, d_beta, mdl->dy,
                                                                   partials, counter, d_out);
        ctx->launches++;
        GMB_CUDA(cudaGetLastError());
        return GMB_OK;
    }
    switch (mdl->flink) {
    case 1: loglik_kernel<1, double><<<grid, block, 0, ctx->stream>>>(n, mdl->P, mdl->ldn, ncols, cols_per_cta, d_zd, mdl->dX, d_beta, mdl->dy, mdl->drowc, var_par, partials, counter, d_out); break;
    case 3: loglik_kernel<3, double><<<grid, block, 0, ctx->stream>>>(n, mdl->P, mdl->ldn, ncols, cols_per_cta, d_zd, mdl->dX, d_beta, mdl->dy, mdl->drowc, var_par, partials, counter, d_out); break;
    case 7: loglik_kernel<7, double><<<grid, block, 0, ctx->stream>>>(n, mdl->P, mdl->ldn, ncols, cols_per_cta, d_zd, mdl->dX, d_beta, mdl->dy, mdl->drowc, var_par, partials, counter, d_out); break;
#define GMB_LL_CASE(F) case F: loglik_kernel<F, double><<<grid, block, 0, ctx->stream>>>(n, mdl->P, mdl->ldn, ncols, cols_per_cta, d_zd, mdl->dX, d_beta, mdl->dy, mdl->drowc, var_par, partials, counter, d_out); break;
    GMB_LL_CASE(2) GMB_LL_CASE(4) GMB_LL_CASE(5) GMB_LL_CASE(6) GMB_LL_CASE(8)
#undef GMB_LL_CASE
    default: return gmb_set_error(GMB_EFAMILY, "family/link code %d has no device kernel", mdl->flink);
    }
    ctx->launches++;
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}

// n_eval (>= GMB_LOGLIK_NB) evaluations of the binomial/logit objective on the model's factor matrix in one launch; d_beta: P x n_eval,
// d_out: n_eval values.  Sets *done = 0 (and launches nothing) when the factor path does not apply.
int gmb_launch_loglik_multi(gmb_model* mdl, const double* d_beta, int n_eval, double* d_out, int* done) {
    gmb_ctx* ctx = mdl->ctx;
    *done = 0;
    const int n = mdl->n, ncols = mdl->niter_local;
    if (mdl->eagg && mdl->flink == 3 && ncols > 0) {             // aggregated rows: the whole batch in two launches (var_par plays no role)
        GMB_TRY(launch_loglik_agg(mdl, d_beta, n_eval, 1.0, d_out));
        *done = 1;
        return GMB_OK;
    }
    if (!(mdl->flink == 3 && mdl->f_valid && ncols > 0 && n_eval >= GMB_LOGLIK_NB) || mdl->prec == 32) return GMB_OK;
    const int groups = (n_eval + GMB_LOGLIK_NB - 1) / GMB_LOGLIK_NB;
    if (groups > 65535) return GMB_OK;
    int half = (n + 1) / 2;
    int TX = 32; while (TX < 256 && TX < half) TX <<= 1;
    int TY = 256 / TX;
    int RT = (half + TX - 1) / TX;
    // one wave of 2 CTAs per SM over all groups, at least 4 TY columns per CTA
    int want_cc = (ctx->sms * 2) / (RT * groups); if (want_cc < 1) want_cc = 1;
    int max_cc = (ncols + 4 * TY - 1) / (4 * TY);
    int CC = want_cc < max_cc ? want_cc : max_cc; if (CC < 1) CC = 1;
    int cols_per_cta = (ncols + CC - 1) / CC;
    cols_per_cta = round_up(cols_per_cta, TY);
    CC = (ncols + cols_per_cta - 1) / cols_per_cta;
    const size_t nblocks = (size_t)RT * CC;
    const size_t part = nblocks * GMB_LOGLIK_NB * groups;
    GMB_TRY(gmb_ctx_scratch(ctx, part + (size_t)(groups + 1) / 2 + 1));
    unsigned int* counters = reinterpret_cast<unsigned int*>(ctx->d_scratch + part);
    GMB_CUDA(cudaMemsetAsync(counters, 0, sizeof(unsigned int) * groups, ctx->stream));
    dim3 grid(RT, CC, groups), block(TX, TY);
    loglik_logit_factor_multi_kernel<GMB_LOGLIK_NB><<<grid, block, 0, ctx->stream>>>(n, mdl->P, mdl->ldn, ncols, cols_per_cta, n_eval, mdl->dF, mdl->dX,
                                                                                      d_beta, mdl->dy, ctx->d_scratch, counters, d_out);
    ctx->launches++;
    GMB_CUDA(cudaGetLastError());
    *done = 1;
    return GMB_OK;
}

// F = exp(s zd) for the first ncols columns (binomial/logit models; called by gmb_model_build_zd)
int gmb_launch_build_factor(gmb_model* mdl, int ncols) {
    gmb_ctx* ctx = mdl->ctx;
    if (ncols <= 0) return GMB_OK;
    const int ne = mdl->eagg ? mdl->agg.ng : mdl->n, lde = mdl->eagg ? mdl->agg.ldn : mdl->ldn;
    const double* ye = mdl->eagg ? nullptr : mdl->dy;
    dim3 grid(ncols, (lde + 255) / 256);
    if (mdl->prec == 32) build_factor_kernel<float><<<grid, 256, 0, ctx->stream>>>(ne, lde, ncols, mdl->dzd32, ye, mdl->dF32);
    else build_factor_kernel<double><<<grid, 256, 0, ctx->stream>>>(ne, lde, ncols, mdl->dzd, ye, mdl->dF);
    ctx->launches++;
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}

int gmb_launch_mcnr(gmb_model* mdl, const double* d_xb, double var_par, double* d_out) {
    gmb_ctx* ctx = mdl->ctx;
    const int n = mdl->n, P = mdl->P, ldn = mdl->ldn, ncols = mdl->niter_local;
    const int nout = P * P + P + 1;
    if (ncols <= 0) { GMB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * nout, ctx->stream)); return GMB_OK; }
    int RT = (n + 255) / 256;
    int want_cc = (ctx->sms * 4 + RT - 1) / RT;
    int max_cc = (ncols + 31) / 32;
    int CC = want_cc < max_cc ? want_cc : max_cc; if (CC < 1) CC = 1;
    int cols_per_cta = (ncols + CC - 1) / CC;
    cols_per_cta = round_up(cols_per_cta, 32);              // whole TMA boxes (8 columns) and whole 4-column reduction groups per warp
    CC = (ncols + cols_per_cta - 1) / cols_per_cta;
    const int NSIG = 64;
    size_t need = (size_t)CC * 2 * ldn + (size_t)RT * 2 * ncols + 2 * (size_t)ldn + NSIG;
    GMB_TRY(gmb_ctx_scratch(ctx, need));
    double* rowpart = ctx->d_scratch;
    double* colpart = rowpart + (size_t)CC * 2 * ldn;
    double* wsum = colpart + (size_t)RT * 2 * ncols;
    double* ssum = wsum + ldn;
    double* sigpart = ssum + ldn;
    double inv_phi = gmb_flink_gaussian(mdl->flink) ? 1.0 / (var_par * var_par) : 1.0;   // mcmlmodel.h:123-133
    dim3 grid(RT, CC);
    const int fl = mdl->flink;
    if (mdl->eagg) {
        const gmb_agg& a = mdl->agg;
        const int ng = a.ng, ldg = a.ldn;
        const int RTC = (ng + 31) / 32, RTg = (ng + 255) / 256;
        int CCg = (ctx->sms * 4 + RTC - 1) / RTC;
        const int max_ccg = (ncols + 31) / 32;
        if (CCg > max_ccg) CCg = max_ccg;
        if (CCg < 1) CCg = 1;
        const int cpc = round_up((ncols + CCg - 1) / CCg, 8);
        CCg = (ncols + cpc - 1) / cpc;
        const int NSIG2 = (ncols + 31) / 32 < 512 ? (ncols + 31) / 32 : 512;
        const size_t npart = (size_t)8 * RTg * (P * P + P) + NSIG2;
        GMB_TRY(gmb_ctx_scratch(ctx, (size_t)CCg * 2 * ldg + (size_t)RTC * 2 * ncols + npart));
        rowpart = ctx->d_scratch; colpart = rowpart + (size_t)CCg * 2 * ldg;
        double* part = colpart + (size_t)RTC * 2 * ncols;
        dim3 gridg(RTC, CCg);
#define GMB_NRA(F, TT, FAC, ZD) mcnr_agg_kernel<F, TT, FAC><<<gridg, 256, 0, ctx->stream>>>(ng, ldg, ncols, cpc, ZD, d_xb, a.drep, a.dcnt, a.deys, a.dess, inv_phi, rowpart, colpart)
        const bool fac = fl == 3 && mdl->f_valid;
        if (mdl->prec == 32) {
            if (fl == 1) GMB_NRA(1, float, false, mdl->dzd32); else if (fac) GMB_NRA(3, float, true, mdl->dF32); else if (fl == 3) GMB_NRA(3, float, false, mdl->dzd32);
            else GMB_NRA(7, float, false, mdl->dzd32);
        } else {
            if (fl == 1) GMB_NRA(1, double, false, mdl->dzd); else if (fac) GMB_NRA(3, double, true, mdl->dF); else if (fl == 3) GMB_NRA(3, double, false, mdl->dzd);
            else GMB_NRA(7, double, false, mdl->dzd);
        }
#undef GMB_NRA
        mcnr_tail_kernel<<<8 * RTg + NSIG2, 256, 0, ctx->stream>>>(ng, n, P, ldg, ncols, RTg, RTC, CCg, NSIG2, a.dX, rowpart, colpart, part, ctx->d_counter, d_out);
        ctx->launches += 2;
        GMB_CUDA(cudaGetLastError());
        return GMB_OK;
    }
    if (mdl->prec == 32 && !(gmbtma::get_encode() && (fl == 1 || fl == 7 || (fl == 3 && mdl->f_valid))))
        return gmb_set_error(GMB_EFAMILY, "fp32 mode: the MCNR step is implemented for poisson/log, binomial/logit and gaussian/identity");
    if ((g_mcnr_tma || mdl->prec == 32) && gmbtma::get_encode() && (fl == 1 || fl == 7 || (fl == 3 && mdl->f_valid))) {
        // TMA pass + one tail launch
        CUtensorMap tm;
        const bool f32 = mdl->prec == 32;
        const void* src = f32 ? (const void*)((fl == 3) ? mdl->dF32 : mdl->dzd32) : (const void*)((fl == 3) ? mdl->dF : mdl->dzd);
        GMB_TRY(gmbtma::make_map(&tm, src, n, ncols, ldn, 256, MCNR_TMA_COLS, false, f32));
        const int NSIG2 = (ncols + 31) / 32 < 512 ? (ncols + 31) / 32 : 512;
        const size_t npart = (size_t)8 * RT * (P * P + P) + NSIG2;
        GMB_TRY(gmb_ctx_scratch(ctx, (size_t)CC * 2 * ldn + (size_t)RT * 2 * ncols + npart));
        rowpart = ctx->d_scratch; colpart = rowpart + (size_t)CC * 2 * ldn;
        double* part = colpart + (size_t)RT * 2 * ncols;
        static bool configured = false;
        if (!configured) {
            GMB_CUDA(cudaFuncSetAttribute(mcnr_tma_kernel<1, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MCNR_TMA_SMEM));
            GMB_CUDA(cudaFuncSetAttribute(mcnr_tma_kernel<3, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MCNR_TMA_SMEM));
            GMB_CUDA(cudaFuncSetAttribute(mcnr_tma_kernel<7, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MCNR_TMA_SMEM));
            GMB_CUDA(cudaFuncSetAttribute(mcnr_tma_kernel<1, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MCNR_TMA_SMEM));
            GMB_CUDA(cudaFuncSetAttribute(mcnr_tma_kernel<3, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MCNR_TMA_SMEM));
            GMB_CUDA(cudaFuncSetAttribute(mcnr_tma_kernel<7, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MCNR_TMA_SMEM));
            configured = true;
        }
#define GMB_NR_TMA(F, TT) mcnr_tma_kernel<F, TT><<<grid, 256, MCNR_TMA_SMEM, ctx->stream>>>(tm, n, ldn, ncols, cols_per_cta, d_xb, mdl->dy, inv_phi, rowpart, colpart)
        if (f32) { if (fl == 1) GMB_NR_TMA(1, float); else if (fl == 3) GMB_NR_TMA(3, float); else GMB_NR_TMA(7, float); }
        else { if (fl == 1) GMB_NR_TMA(1, double); else if (fl == 3) GMB_NR_TMA(3, double); else GMB_NR_TMA(7, double); }
#undef GMB_NR_TMA
        mcnr_tail_kernel<<<8 * RT + NSIG2, 256, 0, ctx->stream>>>(n, n, P, ldn, ncols, RT, RT, CC, NSIG2, mdl->dX, rowpart, colpart, part, ctx->d_counter, d_out);
        ctx->launches += 2;
        GMB_CUDA(cudaGetLastError());
        return GMB_OK;
    }
    size_t smem = 8 * MCNR_STAGES * 256 * sizeof(double);     // 64 KB: the column rings (the row-reduction buffer aliases them)
    {   // per device and cheap: set on every call
        GMB_CUDA(cudaFuncSetAttribute(mcnr_pass1_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GMB_CUDA(cudaFuncSetAttribute(mcnr_pass1_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GMB_CUDA(cudaFuncSetAttribute(mcnr_pass1_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GMB_CUDA(cudaFuncSetAttribute(mcnr_pass1_kernel<7, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    switch (mdl->flink) {
    case 1: mcnr_pass1_kernel<1, false><<<grid, 256, smem, ctx->stream>>>(n, ldn, ncols, cols_per_cta, mdl->dzd, d_xb, mdl->dy, inv_phi, rowpart, colpart); break;
    case 3:
        if (mdl->f_valid) mcnr_pass1_kernel<3, true><<<grid, 256, smem, ctx->stream>>>(n, ldn, ncols, cols_per_cta, mdl->dF, d_xb, mdl->dy, inv_phi, rowpart, colpart);
        else mcnr_pass1_kernel<3, false><<<grid, 256, smem, ctx->stream>>>(n, ldn, ncols, cols_per_cta, mdl->dzd, d_xb, mdl->dy, inv_phi, rowpart, colpart);
        break;
    case 7: mcnr_pass1_kernel<7, false><<<grid, 256, smem, ctx->stream>>>(n, ldn, ncols, cols_per_cta, mdl->dzd, d_xb, mdl->dy, inv_phi, rowpart, colpart); break;
#define GMB_NR_CASE(F) case F: GMB_CUDA(cudaFuncSetAttribute(mcnr_pass1_kernel<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        mcnr_pass1_kernel<F, false><<<grid, 256, smem, ctx->stream>>>(n, ldn, ncols, cols_per_cta, mdl->dzd, d_xb, mdl->dy, inv_phi, rowpart, colpart); break;
    GMB_NR_CASE(2) GMB_NR_CASE(4) GMB_NR_CASE(5) GMB_NR_CASE(8)
#undef GMB_NR_CASE
    case 6: return gmb_set_error(GMB_EFAMILY, "MCNR for binomial/probit needs glmmrBase's dhdmu, which is not part of the reference tree (use method = 'mcem')");
    default: return gmb_set_error(GMB_EFAMILY, "family/link code %d has no device kernel", mdl->flink);
    }
    mcnr_rows_kernel<<<(n + 31) / 32, dim3(32, 32), 0, ctx->stream>>>(n, ldn, CC, rowpart, wsum, ssum);
    mcnr_sigma_kernel<<<NSIG, 256, 0, ctx->stream>>>(n, ncols, RT, colpart, sigpart);
    mcnr_assemble_kernel<<<nout, 256, 0, ctx->stream>>>(n, P, ldn, mdl->dX, wsum, ssum, sigpart, NSIG, d_out);
    ctx->launches += 4;
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}
