// hmc.cu — batched random-effect sampler: C independent copies of glmmr::mcmc::mcmcRunHMC (mhmcmc.h:16-160) advance
// in lock step.  This is the general ("two-contraction") variant: each leapfrog step of all chains is
//
//   RES = r(xb + ZL * V')          one DMMA GEMM (n x Q)(Q x C) with the family residual — and, on a chain's last
//                                  step, the family log-likelihood column sums — fused into its epilogue
//                                  (mcmlmodel.h:156-279 log_grad, :138-153 log_prob)
//   G   = -V' + s * ZL^T * RES     one DMMA GEMM (Q x n)(n x C) with the leapfrog update of momentum and position
//                                  fused into its epilogue (mhmcmc.h:73-78)
//
// so the whitened states never leave the device and nothing is re-read between the contraction and the
// element-wise work.  Per-chain step size adaptation follows mhmcmc.h:107-117 exactly; chains whose trajectory
// (steps = clamp(round(lambda/e), 1, max_steps), :69-70) is shorter than the longest one in the batch are masked.
// Compared with the reference, per proposal: the gradient at the current state and log_prob(u_) are carried over
// from the previous proposal instead of being recomputed (:64, :82), and log_prob(up_) reuses the linear predictor of
// the last leapfrog gradient (:83) — same values, no extra contractions.
//
// RNG: Philox4x32-10, counter (idx, iteration, chain, stream), key = seed; stream 0 = initial state (:48-49),
// 2 = momentum (:62-63), 3 = accept uniform (:85).  Reproducible, unlike std::random_device at :55.
#include "gemm_tma.cuh"

// 1 (default) = the factored variant of the two-contraction sampler when Z is sparse, Z L is not and n >= 2 Q; 0 = always contract with the dense Z L
static int g_hmc_factored = 1;
extern "C" int gmb_hmc_set_factored(int on) { g_hmc_factored = on ? 1 : 0; return GMB_OK; }

namespace {

// per-chain scalar state, stored as rows of a [CS_COUNT][C] double array
enum { CS_EPS = 0, CS_EBAR, CS_H, CS_LLCUR, CS_K0, CS_ACCEPT, CS_TOTSTEPS, CS_LASTPROB, CS_COUNT };

template <int FL>
struct EpiResid {
    static constexpr bool COLSUM = true;
    const double* __restrict__ xb; const double* __restrict__ y; const double* __restrict__ rowc;
    double* RES; int ldr;
    const int* __restrict__ steps;   // nullptr: every column active and wants its log-likelihood
    int s;
    double c0, sigma;
    double* llpart; int C;
    __device__ __forceinline__ bool column_active(int n) const { return steps == nullptr || s < steps[n]; }
    __device__ __forceinline__ bool wants_ll(int n) const { return steps == nullptr || s == steps[n] - 1; }
    __device__ __forceinline__ void store(int m, int n, double acc) const {
        if (!column_active(n)) return;
        RES[(size_t)n * ldr + m] = dev_family_resid<FL>(y[m], xb[m] + acc);
    }
    __device__ __forceinline__ double colterm(int m, int n, double acc) const {
        if (!wants_ll(n)) return 0.0;
        return dev_family_ll<FL>(y[m], xb[m] + acc, (FL == 1 || FL == 2) ? rowc[m] : 0.0, c0, sigma);
    }
    __device__ __forceinline__ void colsum_out(int rt, int n, double v) const {
        if (wants_ll(n)) llpart[(size_t)rt * C + n] = v;
    }
};

struct EpiLeapfrog {
    static constexpr bool COLSUM = false;
    double* VP; double* R; double* G; int ldq;
    const int* __restrict__ steps; const double* __restrict__ eps;
    int s; double sc;
    int init;      // 1: only G = grad(VP) (start of sampling)
    __device__ __forceinline__ bool column_active(int n) const { return init || s < steps[n]; }
    __device__ __forceinline__ void store(int q, int c, double acc) const {
        if (!column_active(c)) return;
        const size_t o = (size_t)c * ldq + q;
        const double v = VP[o];
        const double g = -1.0 * v + sc * acc;                 // mcmlmodel.h:163 + :173/:191/:235
        G[o] = g;
        if (init) return;
        const double e = eps[c];
        double r = R[o] + (e / 2) * g;                        // mhmcmc.h:77
        if (s < steps[c] - 1) {
            r = r + (e / 2) * g;                              // :74 of the next step
            VP[o] = v + e * r;                                // :75
        }
        R[o] = r;
    }
    __device__ __forceinline__ double colterm(int, int, double) const { return 0.0; }
    __device__ __forceinline__ void colsum_out(int, int, double) const {}
};

// initialise_u, mhmcmc.h:47-59 : v ~ N(0, I); accept = 0; H = 0; e = 0.001; ebar = 1
__global__ void __launch_bounds__(128) hmc_init_kernel(int Q, int ldq, int C, uint32_t chain_offset, unsigned long long seed,
                                                       double* __restrict__ V, double* __restrict__ VP, double* __restrict__ cs) {
    const int c = blockIdx.x;
    for (int p = threadIdx.x; p < (Q + 1) / 2; p += blockDim.x) {
        double z0, z1;
        dev_rng_normal2(seed, (uint32_t)p, 0u, chain_offset + c, 0u, z0, z1);
        V[(size_t)c * ldq + 2 * p] = z0; VP[(size_t)c * ldq + 2 * p] = z0;
        if (2 * p + 1 < Q) { V[(size_t)c * ldq + 2 * p + 1] = z1; VP[(size_t)c * ldq + 2 * p + 1] = z1; }
    }
    if (threadIdx.x == 0) {
        cs[CS_EPS * C + c] = 0.001; cs[CS_EBAR * C + c] = 1.0; cs[CS_H * C + c] = 0.0;
        cs[CS_ACCEPT * C + c] = 0.0; cs[CS_TOTSTEPS * C + c] = 0.0; cs[CS_K0 * C + c] = 0.0; cs[CS_LASTPROB * C + c] = 0.0;
    }
}

// after the initial gradient evaluation: GC = G, llcur = sum of the row-tile partials
__global__ void __launch_bounds__(128) hmc_init_finish_kernel(int Q, int ldq, int C, int row_tiles, const double* __restrict__ G,
                                                              double* __restrict__ GC, const double* __restrict__ llpart,
                                                              double* __restrict__ cs) {
    const int c = blockIdx.x;
    for (int q = threadIdx.x; q < Q; q += blockDim.x) GC[(size_t)c * ldq + q] = G[(size_t)c * ldq + q];
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int t = 0; t < row_tiles; t++) s += llpart[(size_t)t * C + c];
        cs[CS_LLCUR * C + c] = s;
    }
}

// start of new_proposal, mhmcmc.h:61-75: momentum draw, kinetic energy, number of steps, first half step
__global__ void __launch_bounds__(128) hmc_begin_kernel(int Q, int ldq, int C, uint32_t chain_offset, unsigned long long seed, uint32_t t,
                                                        double lambda, int max_steps, const double* __restrict__ V,
                                                        const double* __restrict__ GC, double* __restrict__ VP, double* __restrict__ R,
                                                        double* __restrict__ cs, int* __restrict__ steps, int* __restrict__ max_steps_seen) {
    __shared__ double red[32];
    const int c = blockIdx.x;
    const double e = cs[CS_EPS * C + c];
    double k0 = 0.0;
    for (int p = threadIdx.x; p < (Q + 1) / 2; p += blockDim.x) {
        double z[2];
        dev_rng_normal2(seed, (uint32_t)p, t, chain_offset + c, 2u, z[0], z[1]);          // :62-63
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int q = 2 * p + h;
            if (q < Q) {
                const size_t o = (size_t)c * ldq + q;
                k0 += z[h] * z[h];
                const double r = z[h] + (e / 2) * GC[o];                                  // :74 (first step)
                R[o] = r;
                VP[o] = V[o] + e * r;                                                     // :67, :75
            }
        }
    }
    k0 = block_sum(k0, red);
    if (threadIdx.x == 0) {
        cs[CS_K0 * C + c] = 0.5 * k0;                                                     // :66
        double sd = round(lambda / e);                                                    // :69
        int st = sd >= (double)max_steps ? max_steps : (sd < 1.0 ? 1 : (int)sd);          // :69-70 (clamped before the cast)
        if (!(sd == sd)) st = max_steps;
        steps[c] = st;
        cs[CS_TOTSTEPS * C + c] += st;
        atomicMax(max_steps_seen, st);
    }
}

// end of new_proposal, mhmcmc.h:80-117: accept/reject, step-size adaptation, sample storage (:142,:147)
__global__ void __launch_bounds__(128) hmc_end_kernel(int Q, int ldq, int C, uint32_t chain_offset, unsigned long long seed, uint32_t t,
                                                      int row_tiles, int do_adapt, double target_accept,
                                                      double* __restrict__ V, const double* __restrict__ VP, const double* __restrict__ R,
                                                      const double* __restrict__ G, double* __restrict__ GC,
                                                      const double* __restrict__ llpart, double* __restrict__ cs,
                                                      double* __restrict__ out_col /* nullptr or ldq x (cols per chain) x C */,
                                                      int out_stride_cols, int out_col_index, int* __restrict__ max_steps_seen) {
    __shared__ double red[32];
    __shared__ int s_accept;
    const int c = blockIdx.x;
    double k1 = 0.0, pv = 0.0, pvp = 0.0;
    const double pc = -1.0 * log(1.0) - 0.5 * log(2 * GMB_PI_FAMILY);   // log_likelihood(v, 0, 1, 7), mcmlmodel.h:149
    for (int q = threadIdx.x; q < Q; q += blockDim.x) {
        const size_t o = (size_t)c * ldq + q;
        const double r = R[o], v = V[o], vp = VP[o];
        k1 += r * r;
        pv += pc - 0.5 * v * v;
        pvp += pc - 0.5 * vp * vp;
    }
    k1 = block_sum(k1, red);
    pv = block_sum(pv, red);
    pvp = block_sum(pvp, red);
    if (threadIdx.x == 0) {
        double llnew = 0.0;
        for (int tt = 0; tt < row_tiles; tt++) llnew += llpart[(size_t)tt * C + c];
        const double l1 = cs[CS_LLCUR * C + c] + pv;                                      // :82
        const double l2 = llnew + pvp;                                                    // :83
        const double prob = fmin(1.0, exp(-l1 + cs[CS_K0 * C + c] + l2 - 0.5 * k1));      // :84
        double u1, u2;
        dev_rng_uniform2(seed, 0u, t, chain_offset + c, 3u, u1, u2);                      // :85
        const int acc = u1 < prob;                                                        // :86 (false for NaN)
        s_accept = acc;
        if (acc) { cs[CS_LLCUR * C + c] = llnew; cs[CS_ACCEPT * C + c] += 1.0; }
        cs[CS_LASTPROB * C + c] = prob;
        if (do_adapt) {                                                                   // :107-114
            const int iter = (int)t + 1;
            const double f1 = 1.0 / (iter + 10);
            const double pr = (prob == prob) ? prob : 0.0;
            const double H = (1 - f1) * cs[CS_H * C + c] + f1 * (target_accept - pr);
            const double loge = -4.60517 - sqrt((double)iter / 0.05) * H;
            const double powm = pow((double)iter, -0.75);
            const double logbare = powm * loge + (1 - powm) * log(cs[CS_EBAR * C + c]);
            cs[CS_H * C + c] = H;
            cs[CS_EPS * C + c] = exp(loge);
            cs[CS_EBAR * C + c] = exp(logbare);
        } else {
            cs[CS_EPS * C + c] = cs[CS_EBAR * C + c];                                     // :115-117
        }
        if (c == 0) *max_steps_seen = 0;
    }
    __syncthreads();
    const int acc = s_accept;
    for (int q = threadIdx.x; q < Q; q += blockDim.x) {
        const size_t o = (size_t)c * ldq + q;
        double v = V[o];
        if (acc) { v = VP[o]; V[o] = v; GC[o] = G[o]; }                                   // :102-105
        if (out_col) out_col[((size_t)c * out_stride_cols + out_col_index) * ldq + q] = v;
    }
}

__global__ void __launch_bounds__(128) hmc_store_kernel(int Q, int ldq, const double* __restrict__ V, double* __restrict__ out,
                                                        int out_stride_cols, int out_col_index) {
    const int c = blockIdx.x;
    for (int q = threadIdx.x; q < Q; q += blockDim.x)
        out[((size_t)c * out_stride_cols + out_col_index) * ldq + q] = V[(size_t)c * ldq + q];
}

__global__ void xb_kernel2(int n, int P, int ldn, const double* __restrict__ X, const double* __restrict__ beta, double* __restrict__ xb) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int p = 0; p < P; p++) s += X[i + (size_t)p * ldn] * beta[p];
    xb[i] = s;
}

// ---- factored variant: Z sparse, L dense (Z L is dense but n >> Q, config C5: Z = indicator of the location, L a dense 5000 x 5000 factor) ----
// eta = xb + Z (L v') and grad = -v' + s L^T (Z^T r(eta)): the two dense contractions shrink from n x Q to Q x Q (W = L V', G = L^T T) and Z is
// applied by two gather kernels on its ELL form (by rows for eta and the residual, by columns for T = Z^T RES).
constexpr int ZSP_CH = 4;        // chains per thread of the row kernel
template <int FL>
__global__ void __launch_bounds__(256) zsp_resid_kernel(int n, int ngp, int wr, int C, int ldn, int ldq, const double* __restrict__ rv,
                                                        const int* __restrict__ rc, const double* __restrict__ W, const double* __restrict__ xb,
                                                        const double* __restrict__ y, const double* __restrict__ rowc, double* __restrict__ RES,
                                                        const int* __restrict__ steps, int s, double c0, double sigma, double* __restrict__ llpart) {
    __shared__ double red[32];
    const int i = blockIdx.x * 256 + threadIdx.x;
    const bool ok = i < n;
    const double xbi = ok ? xb[i] : 0.0, yi = ok ? y[i] : 0.0, rci = (ok && (FL == 1 || FL == 2)) ? rowc[i] : 0.0;
    for (int cc = 0; cc < ZSP_CH; cc++) {
        const int c = blockIdx.y * ZSP_CH + cc;
        if (c >= C) break;
        const bool active = steps == nullptr || s < steps[c], want_ll = steps == nullptr || s == steps[c] - 1;   // as EpiResid
        if (!active) continue;                                       // block-uniform
        double ll = 0.0;
        if (ok) {
            double eta = xbi;
            for (int w = 0; w < wr; w++) eta = fma(rv[(size_t)w * ngp + i], W[rc[(size_t)w * ngp + i] + (size_t)c * ldq], eta);
            RES[i + (size_t)c * ldn] = dev_family_resid<FL>(yi, eta);
            if (want_ll) ll = dev_family_ll<FL>(yi, eta, rci, c0, sigma);
        }
        if (want_ll) {
            ll = block_sum(ll, red);
            if (threadIdx.x == 0) llpart[(size_t)blockIdx.x * C + c] = ll;
            __syncthreads();
        }
    }
}
// T = Z^T RES for the active chains
__global__ void __launch_bounds__(256) zsp_gather_kernel(int Q, int qp, int wc, int ldn, int ldq, const double* __restrict__ cv,
                                                         const int* __restrict__ cr, const double* __restrict__ RES, double* __restrict__ T,
                                                         const int* __restrict__ steps, int s) {
    const int q = blockIdx.x * 256 + threadIdx.x, c = blockIdx.y;
    if (q >= Q || !(steps == nullptr || s < steps[c])) return;
    double t = 0.0;
    for (int w = 0; w < wc; w++) t = fma(cv[(size_t)w * qp + q], RES[cr[(size_t)w * qp + q] + (size_t)c * ldn], t);
    T[q + (size_t)c * ldq] = t;
}

struct HmcBuffers {
    double *V, *VP, *R, *G, *GC, *RES, *llpart, *cs, *W, *T;
    int *steps, *max_seen;
    int row_tiles;
    bool factored;
};

static bool factored_applicable(const gmb_model* mdl) {
    // worth it when the Q x Q contractions are at most half the n x Q ones: n >= 2 Q, or n >= Q with a triangular factor (its zero tiles are skipped)
    return g_hmc_factored && mdl->zell.checked && mdl->zell.valid && mdl->dL != nullptr &&
           (mdl->n >= 2 * mdl->Q || (mdl->l_lower && mdl->n >= mdl->Q));
}

int hmc_layout(gmb_model* mdl, int C, HmcBuffers& b) {
    gmb_ctx* ctx = mdl->ctx;
    const size_t ldq = mdl->ldq, ldn = mdl->ldn;
    const int rt = gmbtma::row_tile(ctx, mdl->n, C, mdl->Q);
    b.factored = factored_applicable(mdl);
    b.row_tiles = b.factored ? (mdl->n + 255) / 256 : (mdl->n + rt - 1) / rt;
    size_t need = 7 * ldq * C + ldn * C + (size_t)b.row_tiles * C + (size_t)CS_COUNT * C + (size_t)C /*steps as ints*/ + 16;
    if (need > mdl->hmc_work_doubles) {
        if (mdl->hmc_work) { GMB_CUDA(cudaStreamSynchronize(ctx->stream)); gmb_dfree(ctx, mdl->hmc_work); mdl->hmc_work = nullptr; mdl->hmc_work_doubles = 0; }
        GMB_CUDA(gmb_dmalloc(ctx, &mdl->hmc_work, need * sizeof(double)));
        mdl->hmc_work_doubles = need;
    }
    GMB_CUDA(cudaMemsetAsync(mdl->hmc_work, 0, need * sizeof(double), ctx->stream));
    double* p = mdl->hmc_work;
    b.V = p; p += ldq * C; b.VP = p; p += ldq * C; b.R = p; p += ldq * C; b.G = p; p += ldq * C; b.GC = p; p += ldq * C;
    b.W = p; p += ldq * C; b.T = p; p += ldq * C;
    b.RES = p; p += ldn * C; b.llpart = p; p += (size_t)b.row_tiles * C; b.cs = p; p += (size_t)CS_COUNT * C;
    b.steps = reinterpret_cast<int*>(p); p += (C + 1) / 2 + 1;
    b.max_seen = reinterpret_cast<int*>(p);
    return GMB_OK;
}

template <int FL>
int launch_resid(gmb_model* mdl, int C, const HmcBuffers& b, double var_par, const int* steps, int s) {
    EpiResid<FL> epi;
    epi.xb = mdl->dxb; epi.y = mdl->dy; epi.rowc = mdl->drowc; epi.RES = b.RES; epi.ldr = mdl->ldn;
    epi.steps = steps; epi.s = s;
    epi.c0 = (FL == 7 || FL == 8) ? (-1.0 * log(var_par) - 0.5 * log(2 * GMB_PI_FAMILY)) : 0.0;
    epi.sigma = var_par; epi.llpart = b.llpart; epi.C = C;
    return gmbtma::dispatch<false, true>(mdl->ctx, mdl->n, C, mdl->Q, mdl->dZL, mdl->ldn, b.VP, mdl->ldq, epi);
}

template <int FL>
int launch_resid_factored(gmb_model* mdl, int C, const HmcBuffers& b, double var_par, const int* steps, int s) {
    gmb_ctx* ctx = mdl->ctx;
    const gmb_ell& e = mdl->zell;
    GMB_TRY(gmb_dgemm_tri(ctx, 0, 0, mdl->Q, C, mdl->Q, 1.0, mdl->dL, mdl->ldq, b.VP, mdl->ldq, 0.0, b.W, mdl->ldq, mdl->l_lower ? 1 : 0));   // W = L V'
    const double c0 = (FL == 7 || FL == 8) ? (-1.0 * log(var_par) - 0.5 * log(2 * GMB_PI_FAMILY)) : 0.0;
    zsp_resid_kernel<FL><<<dim3((mdl->n + 255) / 256, (C + ZSP_CH - 1) / ZSP_CH), 256, 0, ctx->stream>>>(
        mdl->n, e.ngp, e.wr, C, mdl->ldn, mdl->ldq, e.rv, e.rc, b.W, mdl->dxb, mdl->dy, mdl->drowc, b.RES, steps, s, c0, var_par, b.llpart);
    zsp_gather_kernel<<<dim3((mdl->Q + 255) / 256, C), 256, 0, ctx->stream>>>(mdl->Q, e.qp, e.wc, mdl->ldn, mdl->ldq, e.cv, e.cr, b.RES, b.T, steps, s);
    ctx->launches += 2;
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}

int launch_resid_fl(gmb_model* mdl, int C, const HmcBuffers& b, double var_par, const int* steps, int s) {
    if (b.factored) {
        switch (mdl->flink) {
        case 1: return launch_resid_factored<1>(mdl, C, b, var_par, steps, s);
        case 3: return launch_resid_factored<3>(mdl, C, b, var_par, steps, s);
        case 7: return launch_resid_factored<7>(mdl, C, b, var_par, steps, s);
        case 2: return launch_resid_factored<2>(mdl, C, b, var_par, steps, s);
        case 4: return launch_resid_factored<4>(mdl, C, b, var_par, steps, s);
        case 5: return launch_resid_factored<5>(mdl, C, b, var_par, steps, s);
        case 6: return launch_resid_factored<6>(mdl, C, b, var_par, steps, s);
        case 8: return launch_resid_factored<8>(mdl, C, b, var_par, steps, s);
        }
        return gmb_set_error(GMB_EFAMILY, "family/link code %d has no device kernel", mdl->flink);
    }
    switch (mdl->flink) {
    case 1: return launch_resid<1>(mdl, C, b, var_par, steps, s);
    case 3: return launch_resid<3>(mdl, C, b, var_par, steps, s);
    case 7: return launch_resid<7>(mdl, C, b, var_par, steps, s);
    case 2: return launch_resid<2>(mdl, C, b, var_par, steps, s);
    case 4: return launch_resid<4>(mdl, C, b, var_par, steps, s);
    case 5: return launch_resid<5>(mdl, C, b, var_par, steps, s);
    case 6: return launch_resid<6>(mdl, C, b, var_par, steps, s);
    case 8: return launch_resid<8>(mdl, C, b, var_par, steps, s);
    }
    return gmb_set_error(GMB_EFAMILY, "family/link code %d has no device kernel", mdl->flink);
}

int launch_leap(gmb_model* mdl, int C, const HmcBuffers& b, double var_par, int s, int init) {
    EpiLeapfrog epi;
    epi.VP = b.VP; epi.R = b.R; epi.G = b.G; epi.ldq = mdl->ldq; epi.steps = b.steps; epi.eps = b.cs + (size_t)CS_EPS * C;
    epi.s = s; epi.sc = gmb_flink_gaussian(mdl->flink) ? 1.0 / (var_par * var_par) : 1.0; epi.init = init;
    if (b.factored)     // G = -V' + s L^T T, T = Z^T RES
        return gmbtma::dispatch<true, true>(mdl->ctx, mdl->Q, C, mdl->Q, mdl->dL, mdl->ldq, b.T, mdl->ldq, epi, mdl->l_lower ? 2 : 0);
    return gmbtma::dispatch<true, true>(mdl->ctx, mdl->Q, C, mdl->n, mdl->dZL, mdl->ldn, b.RES, mdl->ldn, epi);
}

}  // namespace

// uploads L (Q x Q host, col-major) and forms ZL = Z L (mcmlmodel.h:67,105)
int gmb_hmc_prepare(gmb_model* mdl, const double* L_host) {
    gmb_ctx* ctx = mdl->ctx;
    const size_t ldq = mdl->ldq, ldn = mdl->ldn, Q = mdl->Q;
    if (!mdl->dL) {
        GMB_CUDA(gmb_dmalloc(ctx, &mdl->dL, sizeof(double) * ldq * Q));
        GMB_CUDA(cudaMemsetAsync(mdl->dL, 0, sizeof(double) * ldq * Q, ctx->stream));
    }
    if (!mdl->dZL) {
        GMB_CUDA(gmb_dmalloc(ctx, &mdl->dZL, sizeof(double) * ldn * Q));
        GMB_CUDA(cudaMemsetAsync(mdl->dZL, 0, sizeof(double) * ldn * Q, ctx->stream));
    }
    if (L_host) {
        GMB_CUDA(cudaMemcpy2DAsync(mdl->dL, ldq * sizeof(double), L_host, Q * sizeof(double), Q * sizeof(double), Q,
                                   cudaMemcpyHostToDevice, ctx->stream));
        bool lower = true;                                       // a Cholesky factor is; any other square root of D takes the full contractions
        for (size_t c = 1; c < Q && lower; c++)
            for (size_t r = 0; r < c; r++) if (L_host[r + c * Q] != 0.0) { lower = false; break; }
        mdl->l_lower = lower;
    } else {
        mdl->l_lower = false;                                    // factor written on the device by the caller: not verified
    }
    GMB_TRY(gmb_dgemm(ctx, 0, 0, mdl->n, mdl->Q, mdl->Q, 1.0, mdl->dZ, mdl->ldn, mdl->dL, mdl->ldq, 0.0, mdl->dZL, mdl->ldn));
    if (L_host) GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    mdl->zl_valid = true;
    mdl->agg.zl_valid = false;
    gmb_sparse_invalidate(mdl);
    return GMB_OK;
}

// Runs the chains; dV_out: ldq x (n_chains * (nsamp + 1)), chain-major columns.  d_stats: CS_COUNT x n_chains.
static int hmc_run(gmb_model* mdl, double var_par, int warmup, int nsamp, double lambda, int max_steps, double target_accept,
                   int adapt, int C, uint32_t chain_offset, uint64_t seed, double* dV_out, std::vector<double>* host_cs, float* ms) {
    gmb_ctx* ctx = mdl->ctx;
    HmcBuffers b;
    GMB_TRY(hmc_layout(mdl, C, b));
    const int Q = mdl->Q, ldq = mdl->ldq;
    GMB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    hmc_init_kernel<<<C, 128, 0, ctx->stream>>>(Q, ldq, C, chain_offset, seed, b.V, b.VP, b.cs);
    ctx->launches++;
    GMB_TRY(launch_resid_fl(mdl, C, b, var_par, nullptr, 0));
    GMB_TRY(launch_leap(mdl, C, b, var_par, 0, 1));
    hmc_init_finish_kernel<<<C, 128, 0, ctx->stream>>>(Q, ldq, C, b.row_tiles, b.G, b.GC, b.llpart, b.cs);
    ctx->launches++;
    const int total = warmup + nsamp;
    int* h_max = reinterpret_cast<int*>(ctx->h_pinned + 128);
    if (warmup == 0) {   // samples.col(0) = u_ before any proposal (mhmcmc.h:142)
        hmc_store_kernel<<<C, 128, 0, ctx->stream>>>(Q, ldq, b.V, dV_out, nsamp + 1, 0);
        ctx->launches++;
    }
    for (int t = 0; t < total; t++) {
        const int do_adapt = (t < warmup) && (t < adapt);                                  // mhmcmc.h:131-136
        hmc_begin_kernel<<<C, 128, 0, ctx->stream>>>(Q, ldq, C, chain_offset, seed, (uint32_t)t, lambda, max_steps,
                                                     b.V, b.GC, b.VP, b.R, b.cs, b.steps, b.max_seen);
        ctx->launches++;
        GMB_CUDA(cudaMemcpyAsync(h_max, b.max_seen, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GMB_CUDA(cudaStreamSynchronize(ctx->stream));
        const int S = *h_max;
        for (int s = 0; s < S; s++) {
            GMB_TRY(launch_resid_fl(mdl, C, b, var_par, b.steps, s));
            GMB_TRY(launch_leap(mdl, C, b, var_par, s, 0));
        }
        const int col = t - warmup + 1;                                                    // :142 (col 0), :147
        double* out = (col >= 0) ? dV_out : nullptr;
        hmc_end_kernel<<<C, 128, 0, ctx->stream>>>(Q, ldq, C, chain_offset, seed, (uint32_t)t, b.row_tiles, do_adapt, target_accept,
                                                   b.V, b.VP, b.R, b.G, b.GC, b.llpart, b.cs, out, nsamp + 1, col < 0 ? 0 : col, b.max_seen);
        ctx->launches++;
    }
    GMB_CUDA(cudaGetLastError());
    GMB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    if (host_cs) {
        host_cs->resize((size_t)CS_COUNT * C);
        GMB_CUDA(cudaMemcpyAsync(host_cs->data(), b.cs, sizeof(double) * CS_COUNT * C, cudaMemcpyDeviceToHost, ctx->stream));
    }
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ms) GMB_CUDA(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return GMB_OK;
}

// hmc_fused.cu
bool gmb_hmc_fused_applicable(const gmb_model* mdl, int C);
size_t gmb_hmc_fused_cs_doubles(int C);
size_t gmb_hmc_fused_scratch_doubles(const gmb_model* mdl, int C);
int gmb_hmc_run_fused(gmb_model* mdl, double var_par, int warmup, int nsamp, double lambda, int max_steps, double target_accept,
                      int adapt, int C, uint32_t chain_offset, uint64_t seed, double* dV_out, double* d_cs);

// hmc_sparse.cu
bool gmb_hmc_sparse_applicable(const gmb_model* mdl);
size_t gmb_hmc_sparse_work_doubles(const gmb_model* mdl, int C);
int gmb_hmc_run_sparse(gmb_model* mdl, double var_par, int warmup, int nsamp, double lambda, int max_steps, double target_accept,
                       int adapt, int C, uint32_t chain_offset, uint64_t seed, double* dV_out, double* d_cs);

// hmc_lane.cu
bool gmb_hmc_lane_applicable(const gmb_model* mdl);
int gmb_hmc_run_lane(gmb_model* mdl, double var_par, int warmup, int nsamp, double lambda, int max_steps, double target_accept,
                     int adapt, int C, uint32_t chain_offset, uint64_t seed, double* dV_out, double* d_cs);

// hmc_comp.cu
bool gmb_hmc_comp_applicable(const gmb_model* mdl);
size_t gmb_hmc_comp_work_doubles(const gmb_model* mdl, int C);
int gmb_hmc_run_comp(gmb_model* mdl, double var_par, int warmup, int nsamp, double lambda, int max_steps, double target_accept,
                     int adapt, int C, uint32_t chain_offset, uint64_t seed, double* dV_out, double* work);
// 1 (default) = large sparse models run the component-decomposed kernels (hmc_comp.cu), 0 = the CTA-per-chain kernel (hmc_sparse.cu)
static int g_hmc_components = 1;
extern "C" int gmb_hmc_set_components(int on) { g_hmc_components = on ? 1 : 0; return GMB_OK; }

// 0 = choose (structure-aware kernels when Z L is sparse enough, else the on-chip variant when the model fits one SM's shared memory,
// else two GEMMs per step), 1 = force the two-GEMM variant, 2 = force on-chip, 3 = force structure-aware
static int g_hmc_variant = 0;
extern "C" int gmb_hmc_set_variant(int variant) {
    if (variant < 0 || variant > 3) return gmb_set_error(GMB_EINVAL, "variant must be 0 (auto), 1 (two-GEMM), 2 (on-chip) or 3 (structure-aware)");
    g_hmc_variant = variant;
    return GMB_OK;
}

static int hmc_run_fused_timed(gmb_model* mdl, bool sparse, double var_par, int warmup, int nsamp, double lambda, int max_steps, double target_accept,
                               int adapt, int C, uint32_t chain_offset, uint64_t seed, double* dV_out, std::vector<double>* host_cs, float* ms) {
    gmb_ctx* ctx = mdl->ctx;
    const bool lane = sparse && gmb_hmc_lane_applicable(mdl);
    const bool comp = sparse && !lane && g_hmc_components && gmb_hmc_comp_applicable(mdl);
    const size_t need = comp ? gmb_hmc_comp_work_doubles(mdl, C)
                             : (sparse ? gmb_hmc_sparse_work_doubles(mdl, C) : gmb_hmc_fused_cs_doubles(C) + gmb_hmc_fused_scratch_doubles(mdl, C));
    if (need > mdl->hmc_work_doubles) {
        if (mdl->hmc_work) { GMB_CUDA(cudaStreamSynchronize(ctx->stream)); gmb_dfree(ctx, mdl->hmc_work); mdl->hmc_work = nullptr; mdl->hmc_work_doubles = 0; }
        GMB_CUDA(gmb_dmalloc(ctx, &mdl->hmc_work, need * sizeof(double)));
        mdl->hmc_work_doubles = need;
    }
    GMB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    if (lane) GMB_TRY(gmb_hmc_run_lane(mdl, var_par, warmup, nsamp, lambda, max_steps, target_accept, adapt, C, chain_offset, seed, dV_out, mdl->hmc_work));
    else if (comp) GMB_TRY(gmb_hmc_run_comp(mdl, var_par, warmup, nsamp, lambda, max_steps, target_accept, adapt, C, chain_offset, seed, dV_out, mdl->hmc_work));
    else if (sparse) GMB_TRY(gmb_hmc_run_sparse(mdl, var_par, warmup, nsamp, lambda, max_steps, target_accept, adapt, C, chain_offset, seed, dV_out, mdl->hmc_work));
    else GMB_TRY(gmb_hmc_run_fused(mdl, var_par, warmup, nsamp, lambda, max_steps, target_accept, adapt, C, chain_offset, seed, dV_out, mdl->hmc_work));
    GMB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    if (host_cs) {
        host_cs->resize((size_t)CS_COUNT * C);
        GMB_CUDA(cudaMemcpyAsync(host_cs->data(), mdl->hmc_work, sizeof(double) * CS_COUNT * C, cudaMemcpyDeviceToHost, ctx->stream));
    }
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ms) GMB_CUDA(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return GMB_OK;
}

static int set_xb(gmb_model* mdl, const double* beta) {
    gmb_ctx* ctx = mdl->ctx;
    if (mdl->beta_cap < mdl->P) {
        if (mdl->dbeta) { GMB_CUDA(cudaStreamSynchronize(ctx->stream)); gmb_dfree(ctx, mdl->dbeta); mdl->dbeta = nullptr; }
        GMB_CUDA(gmb_dmalloc(ctx, &mdl->dbeta, sizeof(double) * mdl->P));
        mdl->beta_cap = mdl->P;
    }
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(ctx->h_pinned, beta, sizeof(double) * mdl->P);
    GMB_CUDA(cudaMemcpyAsync(mdl->dbeta, ctx->h_pinned, sizeof(double) * mdl->P, cudaMemcpyHostToDevice, ctx->stream));
    xb_kernel2<<<(mdl->n + 255) / 256, 256, 0, ctx->stream>>>(mdl->n, mdl->P, mdl->ldn, mdl->dX, mdl->dbeta, mdl->dxb);
    ctx->launches++;
    if (mdl->agg.built && mdl->agg.active) {       // the sampler's aggregated view (aggregate.cu)
        xb_kernel2<<<(mdl->agg.ng + 255) / 256, 256, 0, ctx->stream>>>(mdl->agg.ng, mdl->P, mdl->agg.ldn, mdl->agg.dX, mdl->dbeta, mdl->agg.dxb);
        ctx->launches++;
    }
    return GMB_OK;
}

// Z L of the sampler's aggregated view, from the factor the model currently holds
static int agg_update_zl(gmb_model* mdl) {
    gmb_agg& a = mdl->agg;
    if (!a.built || !a.active || a.zl_valid) return GMB_OK;
    GMB_TRY(gmb_dgemm(mdl->ctx, 0, 0, a.ng, mdl->Q, mdl->Q, 1.0, a.dZ, a.ldn, mdl->dL, mdl->ldq, 0.0, a.dZL, a.ldn));
    a.zl_valid = true;
    return GMB_OK;
}

extern "C" int gmb_hmc_sample(gmb_model* mdl, const double* L, const double* beta, double var_par,
                              int warmup, int nsamp_per_chain, double lambda, int max_steps, double target_accept, int adapt,
                              int n_chains, uint32_t chain_offset, uint64_t seed, int keep_on_device,
                              double* U_out, double* V_out, gmb_hmc_stats* stats) {
    if (!mdl || !beta) return gmb_set_error(GMB_EINVAL, "gmb_hmc_sample: model or beta is NULL");
    if (!L && !mdl->zl_valid) return gmb_set_error(GMB_ESTATE, "gmb_hmc_sample: L is NULL and the model holds no factor yet");
    if (warmup < 0 || nsamp_per_chain < 0 || n_chains <= 0 || max_steps < 1 || !(lambda > 0.0))
        return gmb_set_error(GMB_EINVAL, "gmb_hmc_sample: bad sampler settings (warmup=%d nsamp=%d chains=%d max_steps=%d lambda=%g)",
                             warmup, nsamp_per_chain, n_chains, max_steps, lambda);
    if (gmb_flink_gaussian(mdl->flink) && !(var_par > 0.0)) return gmb_set_error(GMB_EINVAL, "gaussian var_par must be > 0");
    gmb_ctx* ctx = mdl->ctx;
    GMB_CUDA(cudaSetDevice(ctx->device));
    GmbPhase ph(ctx->stream);
    // family/link codes beyond the north-star's three run on the general two-contraction kernels only
    const bool core = gmb_flink_core(mdl->flink);
    if (!core && (g_hmc_variant == 2 || g_hmc_variant == 3))
        return gmb_set_error(GMB_EFAMILY, "family/link code %d runs on the two-contraction sampler only (variant 0 or 1)", mdl->flink);
    if (g_hmc_variant != 1 && core) GMB_TRY(gmb_agg_ensure(mdl));       // row view of the on-chip sampler (not used by the two-GEMM variant)
    ph.mark("hmc: row aggregation");
    if (L) GMB_TRY(gmb_hmc_prepare(mdl, L));
    GMB_TRY(agg_update_zl(mdl));
    ph.mark("hmc: upload L, Z L");
    const bool try_sparse = core && (g_hmc_variant == 0 || g_hmc_variant == 3);
    if (try_sparse) { GMB_TRY(gmb_ell_ensure(mdl)); GMB_TRY(gmb_comp_ensure(mdl)); GMB_TRY(gmb_lane_ensure(mdl)); }
    ph.mark("hmc: sparse forms");
    GMB_TRY(set_xb(mdl, beta));
    const int C = n_chains, cols = nsamp_per_chain + 1;
    const size_t ncol = (size_t)C * cols;
    if (ncol > (size_t)1 << 30) return gmb_set_error(GMB_EINVAL, "too many sample columns");
    if (ncol > mdl->v_cap) {
        if (mdl->dV) { GMB_CUDA(cudaStreamSynchronize(ctx->stream)); gmb_dfree(ctx, mdl->dV); mdl->dV = nullptr; mdl->v_cap = 0; }
        GMB_CUDA(gmb_dmalloc(ctx, &mdl->dV, sizeof(double) * mdl->ldq * ncol));
        GMB_CUDA(cudaMemsetAsync(mdl->dV, 0, sizeof(double) * mdl->ldq * ncol, ctx->stream));
        mdl->v_cap = ncol;
    }
    std::vector<double> hcs;
    float ms = 0.f;
    const bool sparse = try_sparse && gmb_hmc_sparse_applicable(mdl);
    if (g_hmc_variant == 3 && !sparse)
        return gmb_set_error(GMB_EINVAL, "the structure-aware sampler variant was forced but does not fit: Z L (%d x %d) is not sparse enough", mdl->agg.ng, mdl->Q);
    const bool fits = core && !sparse && g_hmc_variant != 3 && gmb_hmc_fused_applicable(mdl, C);
    if (g_hmc_variant == 2 && !fits) return gmb_set_error(GMB_EINVAL, "the on-chip sampler variant was forced but Z L (%d x %d) does not fit in shared memory", mdl->n, mdl->Q);
    if (sparse || (fits && g_hmc_variant != 1))
        GMB_TRY(hmc_run_fused_timed(mdl, sparse, var_par, warmup, nsamp_per_chain, lambda, max_steps, target_accept, adapt, C, chain_offset, seed,
                                    mdl->dV, stats ? &hcs : nullptr, &ms));
    else {
        if (g_hmc_factored && mdl->n >= mdl->Q) GMB_TRY(gmb_zell_ensure(mdl));
        GMB_TRY(hmc_run(mdl, var_par, warmup, nsamp_per_chain, lambda, max_steps, target_accept, adapt, C, chain_offset, seed,
                        mdl->dV, stats ? &hcs : nullptr, &ms));
    }
    ph.mark("hmc: sampler kernels");
    if (stats) {
        double acc = 0, eps = 0, tot = 0;
        for (int c = 0; c < C; c++) { acc += hcs[(size_t)CS_ACCEPT * C + c]; eps += hcs[(size_t)CS_EPS * C + c]; tot += hcs[(size_t)CS_TOTSTEPS * C + c]; }
        const int total = warmup + nsamp_per_chain;
        stats->accept_rate = total > 0 ? acc / ((double)C * total) : 0.0;      // mhmcmc.h:152
        stats->step_size_mean = eps / C;
        stats->steps_mean = total > 0 ? tot / ((double)C * total) : 0.0;
        stats->leapfrog_total = tot;
        stats->kernel_ms = ms;
        stats->n_chains = C;
        stats->nsamp_per_chain = nsamp_per_chain;
        const bool fused = fits && g_hmc_variant != 1;
        stats->kernel_variant = sparse ? 3 : (fused ? 2 : 1);
        stats->rows_used = ((sparse || fused) && mdl->agg.built) ? mdl->agg.ng : mdl->n;
        stats->zl_nonzeros = sparse ? (double)mdl->ell.nnz : (double)stats->rows_used * mdl->Q;
        stats->factored = (!sparse && !fused && factored_applicable(mdl)) ? 1 : 0;
        stats->lane_components = (sparse && gmb_hmc_lane_applicable(mdl)) ? mdl->lane.ncomp : 0;
        stats->component_groups = (sparse && !stats->lane_components && g_hmc_components && gmb_hmc_comp_applicable(mdl)) ? mdl->comp.G : 0;
    }
    if (V_out)
        GMB_CUDA(cudaMemcpy2DAsync(V_out, mdl->Q * sizeof(double), mdl->dV, mdl->ldq * sizeof(double), mdl->Q * sizeof(double), ncol,
                                   cudaMemcpyDeviceToHost, ctx->stream));
    if (U_out || keep_on_device) {
        // u = L v (mhmcmc.h:155) straight into the model's sample matrix
        mdl->zd_valid = false;
        mdl->u_version++;
        GMB_TRY(gmb_model_reserve_samples(mdl, (int)ncol));
        GMB_TRY(gmb_dgemm_tri(ctx, 0, 0, mdl->Q, (int)ncol, mdl->Q, 1.0, mdl->dL, mdl->ldq, mdl->dV, mdl->ldq, 0.0, mdl->dU, mdl->ldq, mdl->l_lower ? 1 : 0));
        mdl->m_local = (int)ncol; mdl->m_total = (int)ncol * ctx->world; mdl->niter_local = mdl->m_local; mdl->niter_total = mdl->m_total;
        if (U_out)
            GMB_CUDA(cudaMemcpy2DAsync(U_out, mdl->Q * sizeof(double), mdl->dU, mdl->ldq * sizeof(double), mdl->Q * sizeof(double), ncol,
                                       cudaMemcpyDeviceToHost, ctx->stream));
    }
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    ph.mark("hmc: u = L v, copies out");
    return GMB_OK;
}

// mcmlModel::log_prob / log_grad for C states at once (mcmlmodel.h:138-153, :156-279 with usezl = true)
extern "C" int gmb_model_logprob_grad(gmb_model* mdl, const double* L, const double* beta, double var_par,
                                      const double* V, int C, double* lp, double* grad) {
    if (!mdl || !beta || !V || C <= 0) return gmb_set_error(GMB_EINVAL, "gmb_model_logprob_grad: bad arguments");
    if (!L && !mdl->zl_valid) return gmb_set_error(GMB_ESTATE, "gmb_model_logprob_grad: L is NULL and the model holds no factor yet");
    if (gmb_flink_gaussian(mdl->flink) && !(var_par > 0.0)) return gmb_set_error(GMB_EINVAL, "gaussian var_par must be > 0");
    gmb_ctx* ctx = mdl->ctx;
    GMB_CUDA(cudaSetDevice(ctx->device));
    if (L) GMB_TRY(gmb_hmc_prepare(mdl, L));
    GMB_TRY(set_xb(mdl, beta));
    HmcBuffers b;
    GMB_TRY(hmc_layout(mdl, C, b));
    const int Q = mdl->Q, ldq = mdl->ldq;
    GMB_CUDA(cudaMemcpy2DAsync(b.VP, ldq * sizeof(double), V, Q * sizeof(double), Q * sizeof(double), C, cudaMemcpyHostToDevice, ctx->stream));
    GMB_TRY(launch_resid_fl(mdl, C, b, var_par, nullptr, 0));
    GMB_TRY(launch_leap(mdl, C, b, var_par, 0, 1));
    std::vector<double> part((size_t)b.row_tiles * C);
    GMB_CUDA(cudaMemcpyAsync(part.data(), b.llpart, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, ctx->stream));
    if (grad)
        GMB_CUDA(cudaMemcpy2DAsync(grad, Q * sizeof(double), b.G, ldq * sizeof(double), Q * sizeof(double), C, cudaMemcpyDeviceToHost, ctx->stream));
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (lp) {
        const double pc = -1.0 * std::log(1.0) - 0.5 * std::log(2 * 3.141593);
        for (int c = 0; c < C; c++) {
            double ll = 0.0;
            for (int t = 0; t < b.row_tiles; t++) ll += part[(size_t)t * C + c];
            double pr = 0.0;
            for (int q = 0; q < Q; q++) { double v = V[(size_t)c * Q + q]; pr += pc - 0.5 * ((v - 0) / 1.0) * ((v - 0) / 1.0); }   // mcmlmodel.h:148-150
            lp[c] = ll + pr;
        }
    }
    return GMB_OK;
}
