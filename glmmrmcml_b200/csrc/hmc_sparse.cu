// hmc_sparse.cu — structure-aware variant of the batched random-effect sampler (chain semantics of
// mhmcmc.h:47-157 exactly as in hmc_fused.cuh / hmc.cu, same random streams).
//
// In every cluster design of the reference Z is an indicator matrix and D(theta) is block diagonal, so Z L is sparse:
// config C2's view (50 distinct rows x 50 random effects) holds 150 non-zeros in 10 lower-triangular 5 x 5 blocks, C1 two per
// row, C4 (n = Q = 10^4, 1000 blocks of 10 x 10) 55 000 of 10^8.  The reference densifies Z (R/R6ModelExtMCML.R:283,297) and the
// dense kernels of this library run those zeros through the tensor pipe.  Here Z L is held twice in ELL form — by rows for
// eta = xb + (Z L) v, by columns for grad = -v + s (Z L)^T r(eta) — and ONE launch runs the whole sample(warmup, nsamp) call:
//
//   small models (rows of the view and Q <= 128): one WARP per chain.  Lane t owns rows / columns t, t + 32, ... : their ELL
//     entries, xb and the row weights sit in its registers, as do momentum, gradient, current and candidate state of its
//     columns; v' and r(eta) are exchanged through 2 x 8 bytes of shared memory per row / column and two __syncwarp per
//     leapfrog step.  No tensor instruction, no block-wide barrier, no global memory inside a trajectory.
//   large models: one CTA per chain, ELL read from global memory (L2 resident), v' and r(eta) in shared memory, the
//     per-column state in a global scratch area, two __syncthreads per leapfrog step.
//
// Sums run over the non-zeros in ascending column (row) order — the dense contraction without its zero terms (the register variants
// accumulate even and odd ELL positions separately to halve the dependent chain).
#include "hmc_sparse_common.cuh"
#include <algorithm>
#include <type_traits>
#include <initializer_list>

namespace {


constexpr int SP_CPB = 4;        // chains (warps) per CTA, warp-per-chain variants
constexpr int SP_MAXW = 8;       // ELL width up to which the entries of a lane's rows / columns are kept in registers

struct SparseParams {
    int ng, Q, ngp, qp, wr, wc, ldq;
    const double* rv; const int* rc; const double* cv; const int* cr;
    const double* xb; const double* cnt; const double* ys;
    const double* lcnt; const double* lys; const double* lsq; const double* lrc;
    double var_par, lambda, target_accept;
    int warmup, nsamp, max_steps, adapt, C;
    uint32_t chain_offset; unsigned long long seed;
    double* dV_out; double* cs_out;
    double* scratch;             // [C][4][qp]: momentum, gradient, current state and its gradient (CTA-per-chain variant)
    int ell_smem;                // warp variants without register ELL: 1 = the kernel stages the ELL arrays in shared memory
};

// ---- ELL construction from the dense view of Z L (column-major, leading dimension ld) ----
__global__ void ell_count_rows_kernel(int ng, int Q, int ld, const double* __restrict__ A, int* __restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ng) return;
    int c = 0;
    for (int q = 0; q < Q; q++) c += A[i + (size_t)q * ld] != 0.0;
    cnt[i] = c;
}
// one warp per column
__global__ void ell_count_cols_kernel(int ng, int Q, int ld, const double* __restrict__ A, int* __restrict__ cnt) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= Q) return;
    int c = 0;
    for (int i0 = 0; i0 < ng; i0 += 32) {
        const int i = i0 + lane;
        const bool nz = i < ng && A[i + (size_t)j * ld] != 0.0;
        c += __popc(__ballot_sync(0xffffffffu, nz));
    }
    if (lane == 0) cnt[j] = c;
}
__global__ void ell_fill_rows_kernel(int ng, int ngp, int Q, int ld, int wr, const double* __restrict__ A,
                                     double* __restrict__ rv, int* __restrict__ rc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ngp) return;
    int w = 0, first = 0;
    if (i < ng) {
        for (int q = 0; q < Q; q++) {
            const double v = A[i + (size_t)q * ld];
            if (v != 0.0) { if (w == 0) first = q; rv[(size_t)w * ngp + i] = v; rc[(size_t)w * ngp + i] = q; w++; }
        }
    }
    for (; w < wr; w++) { rv[(size_t)w * ngp + i] = 0.0; rc[(size_t)w * ngp + i] = first; }
}
__global__ void ell_fill_cols_kernel(int ng, int Q, int qp, int ld, int wc, const double* __restrict__ A,
                                     double* __restrict__ cv, int* __restrict__ cr) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= qp) return;
    int base = 0, first = 0;
    if (j < Q) {
        for (int i0 = 0; i0 < ng; i0 += 32) {
            const int i = i0 + lane;
            const double v = i < ng ? A[i + (size_t)j * ld] : 0.0;
            const unsigned mask = __ballot_sync(0xffffffffu, v != 0.0);
            if (mask && base == 0) first = i0 + __ffs(mask) - 1;
            if (v != 0.0) {
                const int pos = base + __popc(mask & ((1u << lane) - 1u));
                cv[(size_t)pos * qp + j] = v; cr[(size_t)pos * qp + j] = i;
            }
            base += __popc(mask);
        }
    }
    for (int w = base + lane; w < wc; w += 32) { cv[(size_t)w * qp + j] = 0.0; cr[(size_t)w * qp + j] = first; }
}

// ---- the sampler ----
// TPC threads per chain: 32 = one warp per chain, SP_CPB chains per CTA; otherwise one CTA of TPC threads per chain.
// NPT > 0: every thread owns at most NPT rows and NPT columns and keeps their state in registers (NPT == 0: run-time counts, state
// in shared / global memory).  W > 0: the ELL entries of those rows and columns are in registers too, padded to the compile-time
// width W (<= SP_MAXW) — the leapfrog step is then straight-line code.
template <int FL, int TPC, int NPT, int W>
__global__ void __launch_bounds__(TPC == 32 ? 32 * SP_CPB : TPC) hmc_sparse_kernel(const SparseParams p) {
    constexpr bool WARP = (TPC == 32);
    constexpr bool REGELL = W > 0;
    constexpr int CPB = WARP ? SP_CPB : 1;
    constexpr int NR = NPT > 0 ? NPT : 1;
    constexpr int NE = REGELL ? NPT : 1, NW = REGELL ? W : 2;
    static_assert(W == 0 || NPT > 0, "register ELL needs a compile-time row count");
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x;
    const int t = WARP ? (tid & 31) : tid;
    const int slot = WARP ? (tid >> 5) : 0;
    const int chain = blockIdx.x * CPB + slot;
    const int Q = p.Q, ng = p.ng, ngp = p.ngp, qp = p.qp, wr = p.wr, wc = p.wc;
    double* sTab = sm;
    double* s_vp = sm + 64 + (size_t)slot * (qp + ngp);
    double* s_res = s_vp + qp;
    double* s_red = sm + 64 + (size_t)CPB * (qp + ngp);          // 32 doubles (CTA-per-chain reductions)
    // ELL arrays: global memory, or staged in shared memory behind s_red (small models without register ELL)
    const double* rv = p.rv; const int* rc = p.rc; const double* cv = p.cv; const int* cr = p.cr;
    if (REGELL) { if (tid < 16) sTab[tid] = GMB_EXP2_TAB[4 * tid]; }           // 2^(j/16): the conflict-free 128-byte table of dev_family_resid_w_vec16
    else if (tid < 64) sTab[tid] = GMB_EXP2_TAB[tid];
    if (WARP && !REGELL && p.ell_smem) {
        double* s_rv = s_red + 32; double* s_cv = s_rv + (size_t)wr * ngp;
        int* s_rc = reinterpret_cast<int*>(s_cv + (size_t)wc * qp); int* s_cr = s_rc + (size_t)wr * ngp;
        for (int k = tid; k < wr * ngp; k += blockDim.x) { s_rv[k] = p.rv[k]; s_rc[k] = p.rc[k]; }
        for (int k = tid; k < wc * qp; k += blockDim.x) { s_cv[k] = p.cv[k]; s_cr[k] = p.cr[k]; }
        rv = s_rv; rc = s_rc; cv = s_cv; cr = s_cr;
    }
    __syncthreads();
    if (chain >= p.C) return;                                    // whole warps only (CTA-per-chain: grid = C)
    const uint32_t gchain = p.chain_offset + (uint32_t)chain;

    auto chain_sync = [&]() { if (WARP) __syncwarp(); else __syncthreads(); };
    // sum over the threads of a chain, bitwise identical in every thread (xor butterfly; fixed-order sum of the warp totals)
    auto chain_sum = [&](double v) -> double {
        v = warp_sum(v);
        if (!WARP) {
            __syncthreads();
            if ((tid & 31) == 0) s_red[tid >> 5] = v;
            __syncthreads();
            v = s_red[0];
            for (int w = 1; w < TPC / 32; w++) v += s_red[w];
        }
        return v;
    };

    const int KQ = NPT > 0 ? NPT : (Q + TPC - 1) / TPC;
    const int KR = NPT > 0 ? NPT : (ng + TPC - 1) / TPC;
    const double sigma = p.var_par;
    const double sc = (FL == 7) ? 1.0 / (sigma * sigma) : 1.0;
    const double c0 = (FL == 7) ? (-1.0 * log(sigma) - 0.5 * log(2 * GMB_PI_FAMILY)) : 0.0;
    const double pc = -1.0 * log(1.0) - 0.5 * log(2 * GMB_PI_FAMILY);   // log_likelihood(v, 0, 1, 7), mcmlmodel.h:149

    // per-column state: registers (NPT > 0) or this chain's scratch rows
    double r_reg[NR], g_reg[NR], vp_reg[NR], vc_reg[NR], gc_reg[NR];
    double* r_mem = p.scratch + (size_t)chain * 4 * qp;
    double* g_mem = r_mem + qp; double* vc_mem = g_mem + qp; double* gc_mem = vc_mem + qp;
#define SP_GET(reg, mem, k, j) ((NPT > 0) ? reg[(NPT > 0) ? (k) : 0] : mem[j])
#define SP_SET(reg, mem, k, j, val) do { if (NPT > 0) reg[(NPT > 0) ? (k) : 0] = (val); else mem[j] = (val); } while (0)
    // per-row constants and ELL entries in registers
    double xb_reg[NR], cn_reg[NR], ys_reg[NR];
    double erv[NE][NW], ecv[NE][NW]; int erc[NE][NW], ecr[NE][NW];
    if (NPT > 0) {
#pragma unroll
        for (int k = 0; k < NR; k++) {
            const int i = t + k * TPC;
            const bool ok = i < ng;
            xb_reg[k] = ok ? p.xb[i] : 0.0; cn_reg[k] = ok ? p.cnt[i] : 0.0; ys_reg[k] = ok ? p.ys[i] : 0.0;
            r_reg[k] = g_reg[k] = vp_reg[k] = vc_reg[k] = gc_reg[k] = 0.0;
        }
    }
    if (REGELL) {
        // indices are kept as byte offsets into s_vp / s_res; entries beyond the run-time widths are (0.0, offset 0)
#pragma unroll
        for (int k = 0; k < NE; k++) {
            const int i = t + k * TPC;
#pragma unroll
            for (int w = 0; w < NW; w++) {
                const bool okr = i < ng && w < wr, okc = i < Q && w < wc;
                erv[k][w] = okr ? p.rv[(size_t)w * ngp + i] : 0.0; erc[k][w] = okr ? 8 * p.rc[(size_t)w * ngp + i] : 0;
                ecv[k][w] = okc ? p.cv[(size_t)w * qp + i] : 0.0;  ecr[k][w] = okc ? 8 * p.cr[(size_t)w * qp + i] : 0;
            }
        }
    }
    const char* s_vp_b = reinterpret_cast<const char*>(s_vp);
    const char* s_res_b = reinterpret_cast<const char*>(s_res);
    double* my_vp = s_vp + t; double* my_res = s_res + t;

    double eps = 0.001, ebar = 1.0, H = 0.0, llcur = 0.0, llnew = 0.0;   // initialise_u, mhmcmc.h:47-59
    int accept = 0, steps = 1;
    double totsteps = 0.0, lastprob = 0.0;

    // gradient of the log-density at the v' in s_vp (visible to the whole chain) -> g; with_ll: llnew = family log-likelihood there
    // (with_ll is a compile-time tag: the steps of a trajectory before the last one carry no log-likelihood code and no branch)
    auto grad_eval = [&](auto ll_tag) {
        constexpr bool with_ll = decltype(ll_tag)::value;
        double ll = 0.0;
        if constexpr (REGELL) {
            // straight-line code for all slots of the lane (padding rows / columns carry zero weights, zero ELL values and offset 0), written
            // slot-interleaved so that the dependent chains of the NPT slots overlap; even and odd ELL positions accumulate separately
            double eta[NE], eo[NE];
#pragma unroll
            for (int k = 0; k < NE; k++) { eta[k] = xb_reg[k]; eo[k] = 0.0; }
#pragma unroll
            for (int w = 0; w + 1 < NW; w += 2)
#pragma unroll
                for (int k = 0; k < NE; k++) {
                    eta[k] = fma(erv[k][w], *reinterpret_cast<const double*>(s_vp_b + erc[k][w]), eta[k]);
                    eo[k] = fma(erv[k][w + 1], *reinterpret_cast<const double*>(s_vp_b + erc[k][w + 1]), eo[k]);
                }
            if (NW & 1) {
#pragma unroll
                for (int k = 0; k < NE; k++) eta[k] = fma(erv[k][NW - 1], *reinterpret_cast<const double*>(s_vp_b + erc[k][NW - 1]), eta[k]);
            }
#pragma unroll
            for (int k = 0; k < NE; k++) eta[k] += eo[k];
            double res[NE];
            dev_family_resid_w_vec16<FL, NE>(cn_reg, ys_reg, eta, sTab, res);
#pragma unroll
            for (int k = 0; k < NE; k++) my_res[k * TPC] = res[k];
            if (with_ll) {
#pragma unroll
                for (int k = 0; k < NE; k++) {
                    const int i = t + k * TPC;
                    if (i < ng) {
                        const double lq = (FL == 7) ? p.lsq[i] : 0.0, lr = (FL == 1) ? p.lrc[i] : 0.0;
                        ll += dev_family_ll_w<FL>(p.lcnt[i], p.lys[i], lq, lr, eta[k], c0, sigma);
                    }
                }
            }
            chain_sync();                                          // r(eta) of every row is visible
            double gs[NE], go[NE];
#pragma unroll
            for (int k = 0; k < NE; k++) gs[k] = go[k] = 0.0;
#pragma unroll
            for (int w = 0; w + 1 < NW; w += 2)
#pragma unroll
                for (int k = 0; k < NE; k++) {
                    gs[k] = fma(ecv[k][w], *reinterpret_cast<const double*>(s_res_b + ecr[k][w]), gs[k]);
                    go[k] = fma(ecv[k][w + 1], *reinterpret_cast<const double*>(s_res_b + ecr[k][w + 1]), go[k]);
                }
            if (NW & 1) {
#pragma unroll
                for (int k = 0; k < NE; k++) gs[k] = fma(ecv[k][NW - 1], *reinterpret_cast<const double*>(s_res_b + ecr[k][NW - 1]), gs[k]);
            }
#pragma unroll
            for (int k = 0; k < NE; k++) g_reg[k] = -1.0 * vp_reg[k] + sc * (gs[k] + go[k]);       // mcmlmodel.h:163 + :173/:191/:235
        } else {
#pragma unroll
            for (int k = 0; k < KR; k++) {
                const int i = t + k * TPC;
                if (i < ng) {
                    double eta = (NPT > 0) ? xb_reg[(NPT > 0) ? k : 0] : p.xb[i];
#pragma unroll 4
                    for (int w = 0; w < wr; w++) eta = fma(rv[(size_t)w * ngp + i], s_vp[rc[(size_t)w * ngp + i]], eta);
                    const double cn = (NPT > 0) ? cn_reg[(NPT > 0) ? k : 0] : p.cnt[i];
                    const double yy = (NPT > 0) ? ys_reg[(NPT > 0) ? k : 0] : p.ys[i];
                    s_res[i] = dev_family_resid_w<FL>(cn, yy, eta, sTab);
                    if (with_ll) {
                        const double lq = (FL == 7) ? p.lsq[i] : 0.0, lr = (FL == 1) ? p.lrc[i] : 0.0;
                        ll += dev_family_ll_w<FL>(p.lcnt[i], p.lys[i], lq, lr, eta, c0, sigma);
                    }
                }
            }
            chain_sync();                                          // r(eta) of every row is visible
#pragma unroll
            for (int k = 0; k < KQ; k++) {
                const int j = t + k * TPC;
                if (j < Q) {
                    double gs = 0.0;
#pragma unroll 4
                    for (int w = 0; w < wc; w++) gs = fma(cv[(size_t)w * qp + j], s_res[cr[(size_t)w * qp + j]], gs);
                    const double vpj = (NPT > 0) ? vp_reg[(NPT > 0) ? k : 0] : s_vp[j];
                    SP_SET(g_reg, g_mem, k, j, -1.0 * vpj + sc * gs);  // mcmlmodel.h:163 + :173/:191/:235
                }
            }
        }
        if (with_ll) llnew = chain_sum(ll);
    };

    // initial state (:48-49) and its gradient / log-likelihood (carried over between proposals instead of recomputed, :64,:82)
#pragma unroll
    for (int k = 0; k < KQ; k++) {
        const int j = t + k * TPC;
        if (j < Q) {
            double z0, z1;
            dev_rng_normal2(p.seed, (uint32_t)(j >> 1), 0u, gchain, 0u, z0, z1);
            const double v = (j & 1) ? z1 : z0;
            s_vp[j] = v;
            if (NPT > 0) vp_reg[(NPT > 0) ? k : 0] = v;
            SP_SET(vc_reg, vc_mem, k, j, v);
        }
    }
    chain_sync();
    grad_eval(std::true_type{});
    llcur = llnew;
#pragma unroll
    for (int k = 0; k < KQ; k++) {
        const int j = t + k * TPC;
        if (j < Q) SP_SET(gc_reg, gc_mem, k, j, SP_GET(g_reg, g_mem, k, j));
    }
    const int total = p.warmup + p.nsamp, cols = p.nsamp + 1;
    if (p.warmup == 0) {                                                                    // samples.col(0) = u_, mhmcmc.h:142
#pragma unroll
        for (int k = 0; k < KQ; k++) {
            const int j = t + k * TPC;
            if (j < Q) p.dV_out[((size_t)chain * cols) * p.ldq + j] = SP_GET(vc_reg, vc_mem, k, j);
        }
    }

    for (int it = 0; it < total; it++) {
        // ---- new_proposal, mhmcmc.h:61-75 ----
        double k0 = 0.0, pv = 0.0;
        chain_sync();                                              // every thread is done reading the previous v'
#pragma unroll
        for (int k = 0; k < KQ; k++) {
            const int j = t + k * TPC;
            if (j < Q) {
                double z0, z1;
                dev_rng_normal2(p.seed, (uint32_t)(j >> 1), (uint32_t)it, gchain, 2u, z0, z1);      // :62-63
                const double z = (j & 1) ? z1 : z0;
                const double vq = SP_GET(vc_reg, vc_mem, k, j), gq = SP_GET(gc_reg, gc_mem, k, j);
                k0 += z * z;
                pv += pc - 0.5 * vq * vq;
                const double rr = z + (eps / 2) * gq;                                               // :74 (first step)
                const double vn = vq + eps * rr;                                                    // :67, :75
                SP_SET(r_reg, r_mem, k, j, rr);
                if (NPT > 0) vp_reg[(NPT > 0) ? k : 0] = vn;
                s_vp[j] = vn;
            }
        }
        k0 = 0.5 * chain_sum(k0);                                                                   // :66
        pv = chain_sum(pv);
        {
            const double sd = round(p.lambda / eps);                                                // :69
            steps = sd >= (double)p.max_steps ? p.max_steps : (sd < 1.0 ? 1 : (int)sd);             // :69-70
            if (!(sd == sd)) steps = p.max_steps;
            totsteps += steps;
        }
        chain_sync();
        // ---- leapfrog integrator, :73-78 ----
        auto leap = [&](auto more_tag) {
            constexpr bool more = decltype(more_tag)::value;
            grad_eval(std::integral_constant<bool, !more>{});
            if constexpr (REGELL) {
                // straight-line (padding columns: g = r = v' = 0 throughout)
#pragma unroll
                for (int k = 0; k < NE; k++) {
                    double rr = r_reg[k] + (eps / 2) * g_reg[k];                                    // :77
                    if (more) {
                        rr = rr + (eps / 2) * g_reg[k];                                             // :74 of the next step
                        vp_reg[k] = vp_reg[k] + eps * rr;                                           // :75
                        my_vp[k * TPC] = vp_reg[k];
                    }
                    r_reg[k] = rr;
                }
            } else {
#pragma unroll
                for (int k = 0; k < KQ; k++) {
                    const int j = t + k * TPC;
                    if (j < Q) {
                        const double gj = SP_GET(g_reg, g_mem, k, j);
                        double rr = SP_GET(r_reg, r_mem, k, j) + (eps / 2) * gj;                    // :77
                        if (more) {
                            rr = rr + (eps / 2) * gj;                                               // :74 of the next step
                            const double vn = ((NPT > 0) ? vp_reg[(NPT > 0) ? k : 0] : s_vp[j]) + eps * rr;   // :75
                            if (NPT > 0) vp_reg[(NPT > 0) ? k : 0] = vn;
                            s_vp[j] = vn;
                        }
                        SP_SET(r_reg, r_mem, k, j, rr);
                    }
                }
            }
            chain_sync();                                          // the new v' is visible; every read of r(eta) is done
        };
        for (int s = 0; s + 1 < steps; s++) leap(std::true_type{});
        leap(std::false_type{});
        // ---- Metropolis test and adaptation, :80-117 ----
        double k1 = 0.0, pvp = 0.0;
#pragma unroll
        for (int k = 0; k < KQ; k++) {
            const int j = t + k * TPC;
            if (j < Q) {
                const double rr = SP_GET(r_reg, r_mem, k, j);
                const double vj = (NPT > 0) ? vp_reg[(NPT > 0) ? k : 0] : s_vp[j];
                k1 += rr * rr; pvp += pc - 0.5 * vj * vj;
            }
        }
        k1 = 0.5 * chain_sum(k1); pvp = chain_sum(pvp);
        const double l1 = llcur + pv, l2 = llnew + pvp;                                            // :82-83
        const double prob = fmin(1.0, exp(-l1 + k0 + l2 - k1));                                    // :84
        double u1, u2;
        dev_rng_uniform2(p.seed, 0u, (uint32_t)it, gchain, 3u, u1, u2);                            // :85
        const bool acc = u1 < prob;                                                                // :86
        lastprob = prob;
        if (acc) {                                                                                 // :102-105
            accept++; llcur = llnew;
#pragma unroll
            for (int k = 0; k < KQ; k++) {
                const int j = t + k * TPC;
                if (j < Q) {
                    SP_SET(vc_reg, vc_mem, k, j, (NPT > 0) ? vp_reg[(NPT > 0) ? k : 0] : s_vp[j]);
                    SP_SET(gc_reg, gc_mem, k, j, SP_GET(g_reg, g_mem, k, j));
                }
            }
        }
        if (it < p.warmup && it < p.adapt) {                                                       // :107-114, :131-136
            const int iter = it + 1;
            const double f1 = 1.0 / (iter + 10);
            const double pr = (prob == prob) ? prob : 0.0;
            H = (1 - f1) * H + f1 * (p.target_accept - pr);
            const double loge = -4.60517 - sqrt((double)iter / 0.05) * H;
            const double powm = pow((double)iter, -0.75);
            const double logbare = powm * loge + (1 - powm) * log(ebar);
            eps = exp(loge);
            ebar = exp(logbare);
        } else {
            eps = ebar;                                                                            // :115-117
        }
        const int col = it - p.warmup + 1;                                                         // :142 (col 0), :147
        if (col >= 0) {
#pragma unroll
            for (int k = 0; k < KQ; k++) {
                const int j = t + k * TPC;
                if (j < Q) p.dV_out[((size_t)chain * cols + col) * p.ldq + j] = SP_GET(vc_reg, vc_mem, k, j);
            }
        }
    }
    if (t == 0) {
        const int C = p.C;
        p.cs_out[SP_EPS * C + chain] = eps; p.cs_out[SP_EBAR * C + chain] = ebar; p.cs_out[SP_H * C + chain] = H;
        p.cs_out[SP_LLCUR * C + chain] = llcur; p.cs_out[SP_K0 * C + chain] = 0.0; p.cs_out[SP_ACCEPT * C + chain] = (double)accept;
        p.cs_out[SP_TOTSTEPS * C + chain] = totsteps; p.cs_out[SP_LASTPROB * C + chain] = lastprob;
    }
#undef SP_GET
#undef SP_SET
}

constexpr size_t SP_SMEM_LIMIT = (size_t)227 * 1024;

struct SparsePlan { int kind = 0; /* 0 none, 1 warp+regELL NPT2, 2 warp NPT2, 3 warp NPT4, 4 CTA 128, 5 CTA 512 */ int ell_smem = 0; int w = 0; size_t smem = 0; };

SparsePlan sparse_plan(const gmb_ell& e) {
    SparsePlan pl;
    if (!e.valid) return pl;
    const size_t per_chain = (size_t)e.qp + e.ngp;
    const int big = std::max(e.ng, e.Q);
    if (big <= 128) {
        const size_t base = sizeof(double) * (64 + SP_CPB * per_chain + 32);
        if (big <= 64 && e.wr <= SP_MAXW && e.wc <= SP_MAXW) { pl.kind = 1; pl.w = std::max(1, std::max(e.wr, e.wc)); pl.smem = base; return pl; }
        const size_t ell = (size_t)12 * ((size_t)e.wr * e.ngp + (size_t)e.wc * e.qp) + 16;
        pl.kind = big <= 64 ? 2 : 3;
        pl.ell_smem = base + ell <= 96 * 1024 ? 1 : 0;
        pl.smem = base + (pl.ell_smem ? ell : 0);
        return pl;
    }
    const size_t smem = sizeof(double) * (64 + per_chain + 32);
    if (smem > SP_SMEM_LIMIT) return pl;
    pl.kind = big <= 1024 ? 4 : 5; pl.smem = smem;
    return pl;
}

template <int FL, int TPC, int NPT, int W>
int launch_sparse(gmb_ctx* ctx, const SparseParams& p, size_t smem) {
    auto kern = hmc_sparse_kernel<FL, TPC, NPT, W>;
    GMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    constexpr int CPB = TPC == 32 ? SP_CPB : 1;
    const int grid = (p.C + CPB - 1) / CPB;
    kern<<<grid, TPC == 32 ? 32 * SP_CPB : TPC, smem, ctx->stream>>>(p);
    ctx->launches++;
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}

template <int FL>
int launch_sparse_kind(gmb_ctx* ctx, const SparseParams& p, const SparsePlan& pl) {
    switch (pl.kind) {
    case 1:
        switch (pl.w) {
        case 1: return launch_sparse<FL, 32, 2, 1>(ctx, p, pl.smem);
        case 2: return launch_sparse<FL, 32, 2, 2>(ctx, p, pl.smem);
        case 3: return launch_sparse<FL, 32, 2, 3>(ctx, p, pl.smem);
        case 4: return launch_sparse<FL, 32, 2, 4>(ctx, p, pl.smem);
        case 5: return launch_sparse<FL, 32, 2, 5>(ctx, p, pl.smem);
        case 6: return launch_sparse<FL, 32, 2, 6>(ctx, p, pl.smem);
        case 7: return launch_sparse<FL, 32, 2, 7>(ctx, p, pl.smem);
        case 8: return launch_sparse<FL, 32, 2, 8>(ctx, p, pl.smem);
        }
        break;
    case 2: return launch_sparse<FL, 32, 2, 0>(ctx, p, pl.smem);
    case 3: return launch_sparse<FL, 32, 4, 0>(ctx, p, pl.smem);
    case 4: return launch_sparse<FL, 128, 0, 0>(ctx, p, pl.smem);
    case 5: return launch_sparse<FL, 512, 0, 0>(ctx, p, pl.smem);
    }
    return gmb_set_error(GMB_EINVAL, "structure-aware sampler: no kernel for this model");
}

template <class T>
int ensure_cap(gmb_ctx* ctx, T** ptr, size_t* cap, size_t need) {
    if (need <= *cap) return GMB_OK;
    if (*ptr) { GMB_CUDA(cudaStreamSynchronize(ctx->stream)); gmb_dfree(ctx, *ptr); *ptr = nullptr; *cap = 0; }
    GMB_CUDA(gmb_dmalloc(ctx, ptr, sizeof(T) * need));
    *cap = need;
    return GMB_OK;
}

}  // namespace

void gmb_ell_free(gmb_model* mdl) {
    gmb_ctx* ctx = mdl->ctx;
    for (gmb_ell* e : {&mdl->ell, &mdl->zell}) {
        gmb_dfree(ctx, e->rv); gmb_dfree(ctx, e->rc); gmb_dfree(ctx, e->cv); gmb_dfree(ctx, e->cr); gmb_dfree(ctx, e->dcnt);
        *e = gmb_ell();
    }
}

// Sparse (ELL) form of a dense device matrix; gmb_ell_ensure: of the sampler's view of Z L (rebuilt whenever the factor changes).  The structure-aware kernels apply when few enough
// entries of the view are non-zero and the ELL padding stays within a small multiple of the non-zeros (no dense row or column).
int gmb_ell_build(gmb_ctx* ctx, const double* A, int ld, int ng, int Q, gmb_ell& e) {
    e.valid = false;
    e.ng = ng; e.Q = Q; e.ngp = round_up(ng, 32); e.qp = round_up(Q, 32);
    // two slots of run-time widths per thread in the register variants need zero-filled padding rows / columns up to 64
    if (std::max(ng, Q) <= 64) { e.ngp = 64; e.qp = 64; }
    size_t cap = e.cnt_cap;
    GMB_TRY(ensure_cap(ctx, &e.dcnt, &cap, (size_t)ng + Q)); e.cnt_cap = cap;
    ell_count_rows_kernel<<<(ng + 127) / 128, 128, 0, ctx->stream>>>(ng, Q, ld, A, e.dcnt);
    ell_count_cols_kernel<<<(Q * 32 + 127) / 128, 128, 0, ctx->stream>>>(ng, Q, ld, A, e.dcnt + ng);
    ctx->launches += 2;
    std::vector<int> cnt((size_t)ng + Q);
    GMB_CUDA(cudaMemcpyAsync(cnt.data(), e.dcnt, sizeof(int) * cnt.size(), cudaMemcpyDeviceToHost, ctx->stream));
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    long long nnz = 0; int wr = 0, wc = 0;
    for (int i = 0; i < ng; i++) { nnz += cnt[i]; wr = std::max(wr, cnt[i]); }
    for (int j = 0; j < Q; j++) wc = std::max(wc, cnt[(size_t)ng + j]);
    e.nnz = nnz; e.wr = wr; e.wc = wc;
    const long long dense = (long long)ng * Q;
    const long long padded = (long long)wr * e.ngp + (long long)wc * e.qp;
    // break-even against the dense kernels: about 1/8 non-zeros for the register-resident variants, 1/16 when the ELL arrays stream from L2
    const long long dens = std::max(ng, Q) <= 128 ? 8 : 16;
    const bool sparse_enough = nnz * dens <= dense && padded <= 8 * std::max(nnz, 64LL) + 2 * ((long long)e.ngp + e.qp);
    e.checked = true;
    if (!sparse_enough) return GMB_OK;
    cap = e.r_cap;
    { size_t c2 = e.r_cap; GMB_TRY(ensure_cap(ctx, &e.rv, &cap, (size_t)std::max(wr, 1) * e.ngp)); GMB_TRY(ensure_cap(ctx, &e.rc, &c2, (size_t)std::max(wr, 1) * e.ngp)); e.r_cap = cap; }
    cap = e.c_cap;
    { size_t c2 = e.c_cap; GMB_TRY(ensure_cap(ctx, &e.cv, &cap, (size_t)std::max(wc, 1) * e.qp)); GMB_TRY(ensure_cap(ctx, &e.cr, &c2, (size_t)std::max(wc, 1) * e.qp)); e.c_cap = cap; }
    ell_fill_rows_kernel<<<(e.ngp + 127) / 128, 128, 0, ctx->stream>>>(ng, e.ngp, Q, ld, wr, A, e.rv, e.rc);
    ell_fill_cols_kernel<<<(e.qp * 32 + 127) / 128, 128, 0, ctx->stream>>>(ng, Q, e.qp, ld, wc, A, e.cv, e.cr);
    ctx->launches += 2;
    GMB_CUDA(cudaGetLastError());
    e.valid = true;
    return GMB_OK;
}

int gmb_ell_ensure(gmb_model* mdl) {
    gmb_ell& e = mdl->ell;
    if (e.checked) return GMB_OK;
    const gmb_agg& a = mdl->agg;
    e.valid = false;
    if (!a.built) return gmb_set_error(GMB_ESTATE, "structure-aware sampler: the row view of the model has not been built");
    return gmb_ell_build(mdl->ctx, a.active ? a.dZL : mdl->dZL, a.ldn, a.ng, mdl->Q, e);
}

// sparse form of Z itself (all n rows; Z never changes): the factored two-contraction sampler of hmc.cu applies Z and L separately
int gmb_zell_ensure(gmb_model* mdl) {
    gmb_ell& e = mdl->zell;
    if (e.checked) return GMB_OK;
    return gmb_ell_build(mdl->ctx, mdl->dZ, mdl->ldn, mdl->n, mdl->Q, e);
}

// true when the structure-aware kernels can run the model's current Z L (gmb_ell_ensure must have been called)
bool gmb_hmc_sparse_applicable(const gmb_model* mdl) { return sparse_plan(mdl->ell).kind != 0; }

size_t gmb_hmc_sparse_work_doubles(const gmb_model* mdl, int C) {
    return round_up_sz((size_t)SP_COUNT * C + 16, 16) + (size_t)C * 4 * mdl->ell.qp;
}

// Same contract as gmb_hmc_run_fused: dV_out is ldq x (C * (nsamp + 1)) chain-major, d_cs is SP_COUNT x C followed by the scratch area.
int gmb_hmc_run_sparse(gmb_model* mdl, double var_par, int warmup, int nsamp, double lambda, int max_steps, double target_accept,
                       int adapt, int C, uint32_t chain_offset, uint64_t seed, double* dV_out, double* d_cs) {
    gmb_ctx* ctx = mdl->ctx;
    const gmb_ell& e = mdl->ell;
    const gmb_agg& a = mdl->agg;
    const SparsePlan pl = sparse_plan(e);
    if (!pl.kind) return gmb_set_error(GMB_EINVAL, "the structure-aware sampler does not apply to this model (Z L %d x %d, %lld non-zeros)", e.ng, e.Q, e.nnz);
    SparseParams p;
    p.ng = e.ng; p.Q = e.Q; p.ngp = e.ngp; p.qp = e.qp; p.wr = e.wr; p.wc = e.wc; p.ldq = mdl->ldq;
    p.rv = e.rv; p.rc = e.rc; p.cv = e.cv; p.cr = e.cr;
    p.xb = a.active ? a.dxb : mdl->dxb;
    p.cnt = a.dcnt; p.ys = a.dys; p.lcnt = a.dlcnt; p.lys = a.dlys; p.lsq = a.dlsq; p.lrc = a.dlrc;
    p.var_par = var_par; p.lambda = lambda; p.target_accept = target_accept;
    p.warmup = warmup; p.nsamp = nsamp; p.max_steps = max_steps; p.adapt = adapt; p.C = C;
    p.chain_offset = chain_offset; p.seed = seed; p.dV_out = dV_out; p.cs_out = d_cs;
    p.scratch = d_cs + round_up_sz((size_t)SP_COUNT * C + 16, 16);
    p.ell_smem = pl.ell_smem;
    switch (mdl->flink) {
    case 1: return launch_sparse_kind<1>(ctx, p, pl);
    case 3: return launch_sparse_kind<3>(ctx, p, pl);
    case 7: return launch_sparse_kind<7>(ctx, p, pl);
    }
    return gmb_set_error(GMB_EFAMILY, "family/link code %d has no device kernel", mdl->flink);
}
