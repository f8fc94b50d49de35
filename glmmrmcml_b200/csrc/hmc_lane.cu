// hmc_lane.cu — structure-aware sampler for SMALL block-structured models: one LANE per connected component of Z L (K6l).
//
// In the cluster designs of the reference (README.md:14-43, 70-77) the sampler's view of Z L — distinct rows of [X | Z] x random effects — falls
// apart into a few small connected components: config C2 (gr(cl)*ar1(t)) has 10 components of 5 rows x 5 columns, C1 ((1|gr(cl)) + (1|gr(cl,t)))
// 10 components of 5 rows x 6 columns.  The warp-per-chain kernel of hmc_sparse.cu spreads one chain over 32 lanes and pays, on every leapfrog
// step, two shared-memory exchanges (v' and r(eta)) with a __syncwarp each on top of the dependent exp / reciprocal chain: ~550 cycles per step
// with one or two warps per scheduler to hide it (ncu, round 1: 10.8 % of the warp slots busy, FP64 pipe at 41 %).
//
// A leapfrog trajectory only couples the random effects of one component (step size and trajectory length are per-chain scalars), so here a
// lane integrates a WHOLE component: its dense R x CQ block of Z L, the per-row constants and the momentum / position / gradient of its columns
// live in that lane's registers, and a leapfrog step is straight-line code without any exchange, barrier or memory access — the R rows give
// the instruction-level parallelism that hides the exp / reciprocal latency.  A chain occupies as many lanes as the view has components
// (C2: 10), a warp hosts floor(32 / components) chains; lanes of different chains simply run different trip counts.  Only the Metropolis
// test needs sums over the lanes of a chain: once per proposal, through shared memory, added in lane order by every lane (bitwise identical
// replicas of the per-chain scalars).
//
// The FP64 unit of a scheduler issues one warp instruction every two cycles whatever the number of active lanes, and 1000 chains of 10
// components are only 334 full warps for 592 schedulers: with one lane per component every busy scheduler is bound by its single warp's FP64
// instruction stream (~170 per step) while 40 % of the schedulers idle.  A component is therefore SPLIT over L lanes (L = 3 for 10
// components: 30 lanes per chain, one chain per warp, ~65 FP64 instructions per lane and step): lane j of a component owns its rows and columns
// j, j + L, ...; the component's v' and r(eta) are all-gathered inside the L-lane group with warp shuffles (no shared memory, no barrier).
//
// Chain semantics (mhmcmc.h:47-157), random streams and outputs are those of every other sampler kernel; sums over the entries of a row /
// column run over the component's dense block in ascending order instead of over the ELL entries (the zero entries add nothing).
#include "hmc_sparse_common.cuh"
#include <algorithm>
#include <numeric>
#include <type_traits>

static int g_hmc_lane = 1;
extern "C" int gmb_hmc_set_lane(int on) { g_hmc_lane = on ? 1 : 0; return GMB_OK; }

namespace {

constexpr int LN_WPB = 1;        // warps per CTA (one: 334 CTAs spread over all SMs measured 2 % faster than 4 warps on 84 SMs)

struct LaneParams {
    int lpc, cpw;                // lanes per chain (components x L), chains per warp
    const int* rowid; const int* colid;          // [lpc][RL] view row, [lpc][CL] column of Z L owned by a lane; -1 = padding
    const double* brow; const double* bcol;      // [lpc][RL][CL * L]: the lane's rows of its component's block; [lpc][RL * L][CL]: its columns
    const double* xb; const double* cnt; const double* ys;
    const double* lcnt; const double* lys; const double* lsq; const double* lrc;
    double var_par, lambda, target_accept;
    int warmup, nsamp, max_steps, adapt, C, ldq;
    uint32_t chain_offset; unsigned long long seed;
    double* dV_out; double* cs_out;
};

// L lanes per component; a lane owns RL rows and CL columns of it (component-local row r <-> lane r % L, slot r / L; same for columns)
// TRI (L = 1 only): every component's block is lower triangular in its local numbering (identity-like Z on the component, Cholesky factor of its
// covariance block): the products skip the zero triangle.
template <int FL, int L, int RL, int CL, bool TRI = false>
__global__ void __launch_bounds__(32 * LN_WPB) hmc_lane_kernel(const LaneParams p) {
    constexpr int R = RL * L, CQ = CL * L;                         // padded component size
    __shared__ double sTab[16];
    __shared__ double sRed[LN_WPB][3][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 16) sTab[tid] = GMB_EXP2_TAB[4 * tid];                // 2^(j/16): conflict-free table of dev_family_resid_w_vec16
    __syncthreads();
    const int lpc = p.lpc, cpw = p.cpw;
    const int sub = lane / lpc, cl = lane - sub * lpc;              // chain slot inside the warp, lane of the chain
    const int chain = (blockIdx.x * LN_WPB + warp) * cpw + sub;
    const bool live = sub < cpw && chain < p.C;                     // idle lanes run the same code on an empty component
    const int lane0 = sub * lpc;                                    // first lane of this lane's chain
    const int grp0 = lane - (cl % L);                               // first lane of this lane's component group
    // lanes that execute the same trajectory (one chain): the shuffles inside a trajectory name exactly these
    const unsigned cmask = (sub < cpw) ? (unsigned)(((1ull << lpc) - 1ull) << lane0) : (unsigned)(0xffffffffull << (cpw * lpc));
    const uint32_t gchain = p.chain_offset + (uint32_t)(live ? chain : 0);

    const double sigma = p.var_par;
    const double sc = (FL == 7) ? 1.0 / (sigma * sigma) : 1.0;
    const double c0 = (FL == 7) ? (-1.0 * log(sigma) - 0.5 * log(2 * GMB_PI_FAMILY)) : 0.0;
    const double pc = -1.0 * log(1.0) - 0.5 * log(2 * GMB_PI_FAMILY);   // log_likelihood(v, 0, 1, 7), mcmlmodel.h:149

    // this lane's share of its component, in registers
    double Br[RL][CQ], Bc[R][CL], xb[RL], cn[RL], ys[RL];
    int rid[RL], cid[CL];
#pragma unroll
    for (int a = 0; a < RL; a++) {
        rid[a] = live ? p.rowid[cl * RL + a] : -1;
        const bool ok = rid[a] >= 0;
        xb[a] = ok ? p.xb[rid[a]] : 0.0; cn[a] = ok ? p.cnt[rid[a]] : 0.0; ys[a] = ok ? p.ys[rid[a]] : 0.0;
#pragma unroll
        for (int c = 0; c < CQ; c++) Br[a][c] = live ? p.brow[(cl * RL + a) * CQ + c] : 0.0;
    }
#pragma unroll
    for (int b = 0; b < CL; b++) cid[b] = live ? p.colid[cl * CL + b] : -1;
    if constexpr (L > 1) {                                          // (with one lane per component the rows of the block ARE its columns: Br serves both)
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int b = 0; b < CL; b++) Bc[r][b] = live ? p.bcol[(cl * R + r) * CL + b] : 0.0;
    }

    double rr[CL], g[CL], vp[CL], vc[CL], gc[CL];
    double eps = 0.001, ebar = 1.0, H = 0.0, llcur = 0.0, llnew = 0.0;   // initialise_u, mhmcmc.h:47-59
    int accept = 0, steps = 1;
    double totsteps = 0.0, lastprob = 0.0;

    // sums over the lanes of a chain: every lane adds its chain's values in lane order (identical bits in every replica)
    auto chain_sum3 = [&](double& a, double& b, double& c) {
        __syncwarp();
        sRed[warp][0][lane] = a; sRed[warp][1][lane] = b; sRed[warp][2][lane] = c;
        __syncwarp();
        double sa = 0.0, sb = 0.0, sc3 = 0.0;
        for (int l = 0; l < lpc; l++) { const int src = min(lane0 + l, 31); sa += sRed[warp][0][src]; sb += sRed[warp][1][src]; sc3 += sRed[warp][2][src]; }
        a = sa; b = sb; c = sc3;
    };
    // normal draw of column j (pairs of columns share one Philox block: the pair is redrawn only when it changes)
    auto draws = [&](uint32_t it, uint32_t stream, double (&z)[CL]) {
        int last = -1; double z0 = 0.0, z1 = 0.0;
#pragma unroll
        for (int b = 0; b < CL; b++) {
            z[b] = 0.0;
            if (cid[b] >= 0) {
                if ((cid[b] >> 1) != last) { last = cid[b] >> 1; dev_rng_normal2(p.seed, (uint32_t)last, it, gchain, stream, z0, z1); }
                z[b] = (cid[b] & 1) ? z1 : z0;
            }
        }
    };
    // gradient of the log-density at vp -> g; with_ll: *ll_out = this lane's share of the family log-likelihood there
    auto grad_eval = [&](auto ll_tag, double* ll_out) {
        constexpr bool with_ll = decltype(ll_tag)::value;
        double vfull[CQ], eta[RL], res[RL], rfull[R];
#pragma unroll
        for (int c = 0; c < CQ; c++) vfull[c] = (L == 1) ? vp[c] : __shfl_sync(cmask, vp[c / L], grp0 + (c % L));     // all-gather v' of the component
#pragma unroll
        for (int a = 0; a < RL; a++) eta[a] = xb[a];
#pragma unroll
        for (int c = 0; c < CQ; c++)
#pragma unroll
            for (int a = 0; a < RL; a++) if (!TRI || c <= a) eta[a] = fma(Br[a][c], vfull[c], eta[a]);
        dev_family_resid_w_vec16<FL, RL>(cn, ys, eta, sTab, res);
        if (with_ll) {
            double ll = 0.0;
#pragma unroll
            for (int a = 0; a < RL; a++) {
                if (rid[a] >= 0) {
                    const int i = rid[a];
                    const double lq = (FL == 7) ? p.lsq[i] : 0.0, lr = (FL == 1) ? p.lrc[i] : 0.0;
                    ll += dev_family_ll_w<FL>(p.lcnt[i], p.lys[i], lq, lr, eta[a], c0, sigma);
                }
            }
            *ll_out = ll;
        }
#pragma unroll
        for (int r = 0; r < R; r++) rfull[r] = (L == 1) ? res[r] : __shfl_sync(cmask, res[r / L], grp0 + (r % L));   // all-gather r(eta)
        double gs[CL];
#pragma unroll
        for (int b = 0; b < CL; b++) gs[b] = 0.0;
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int b = 0; b < CL; b++) if (!TRI || b <= r) gs[b] = fma((L == 1) ? Br[(L == 1) ? r : 0][b] : Bc[r][b], rfull[r], gs[b]);
#pragma unroll
        for (int b = 0; b < CL; b++) g[b] = -1.0 * vp[b] + sc * gs[b];                              // mcmlmodel.h:163 + :173/:191/:235
    };

    // initial state (:48-49) and its gradient / log-likelihood (carried over between proposals instead of recomputed, :64,:82)
    {
        double z[CL];
        draws(0u, 0u, z);
#pragma unroll
        for (int b = 0; b < CL; b++) { vp[b] = z[b]; vc[b] = z[b]; rr[b] = 0.0; }
        double ll = 0.0, d1 = 0.0, d2 = 0.0;
        grad_eval(std::true_type{}, &ll);
        chain_sum3(ll, d1, d2);
        llcur = ll;
#pragma unroll
        for (int b = 0; b < CL; b++) gc[b] = g[b];
    }
    const int total = p.warmup + p.nsamp, cols = p.nsamp + 1;
    if (p.warmup == 0) {                                                                            // samples.col(0) = u_, mhmcmc.h:142
#pragma unroll
        for (int b = 0; b < CL; b++) if (cid[b] >= 0) p.dV_out[((size_t)chain * cols) * p.ldq + cid[b]] = vc[b];
    }

    for (int it = 0; it < total; it++) {
        // ---- new_proposal, mhmcmc.h:61-75 ----
        double k0 = 0.0, pv = 0.0;
        {
            double z[CL];
            draws((uint32_t)it, 2u, z);                                                             // :62-63
#pragma unroll
            for (int b = 0; b < CL; b++) {
                if (cid[b] >= 0) { k0 += z[b] * z[b]; pv += pc - 0.5 * vc[b] * vc[b]; }
                rr[b] = z[b] + (eps / 2) * gc[b];                                                   // :74 (first step)
                vp[b] = vc[b] + eps * rr[b];                                                        // :67, :75
            }
        }
        {
            const double sd = round(p.lambda / eps);                                                // :69
            steps = sd >= (double)p.max_steps ? p.max_steps : (sd < 1.0 ? 1 : (int)sd);             // :69-70
            if (!(sd == sd)) steps = p.max_steps;
            totsteps += steps;
        }
        // ---- leapfrog integrator, :73-78: the only exchange is the all-gather inside a component group ----
        for (int s = 0; s + 1 < steps; s++) {
            grad_eval(std::false_type{}, nullptr);
#pragma unroll
            for (int b = 0; b < CL; b++) {
                double q = rr[b] + (eps / 2) * g[b];                                                // :77
                q = q + (eps / 2) * g[b];                                                           // :74 of the next step
                vp[b] = vp[b] + eps * q;                                                            // :75
                rr[b] = q;
            }
        }
        double llp = 0.0;
        grad_eval(std::true_type{}, &llp);
#pragma unroll
        for (int b = 0; b < CL; b++) rr[b] = rr[b] + (eps / 2) * g[b];                              // :77
        // ---- Metropolis test and adaptation, :80-117 ----
        double k1 = 0.0, pvp = 0.0;
#pragma unroll
        for (int b = 0; b < CL; b++) if (cid[b] >= 0) { k1 += rr[b] * rr[b]; pvp += pc - 0.5 * vp[b] * vp[b]; }
        chain_sum3(k0, pv, llp);
        double dummy = 0.0;
        chain_sum3(k1, pvp, dummy);
        k0 *= 0.5; k1 *= 0.5;                                                                       // :66
        llnew = llp;
        const double l1 = llcur + pv, l2 = llnew + pvp;                                            // :82-83
        const double prob = fmin(1.0, exp(-l1 + k0 + l2 - k1));                                    // :84
        double u1, u2;
        dev_rng_uniform2(p.seed, 0u, (uint32_t)it, gchain, 3u, u1, u2);                            // :85
        const bool acc = u1 < prob;                                                                // :86
        lastprob = prob;
        if (acc) {                                                                                 // :102-105
            accept++; llcur = llnew;
#pragma unroll
            for (int b = 0; b < CL; b++) { vc[b] = vp[b]; gc[b] = g[b]; }
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
            for (int b = 0; b < CL; b++) if (cid[b] >= 0) p.dV_out[((size_t)chain * cols + col) * p.ldq + cid[b]] = vc[b];
        }
    }
    if (live && cl == 0) {
        const int C = p.C;
        p.cs_out[SP_EPS * C + chain] = eps; p.cs_out[SP_EBAR * C + chain] = ebar; p.cs_out[SP_H * C + chain] = H;
        p.cs_out[SP_LLCUR * C + chain] = llcur; p.cs_out[SP_K0 * C + chain] = 0.0; p.cs_out[SP_ACCEPT * C + chain] = (double)accept;
        p.cs_out[SP_TOTSTEPS * C + chain] = totsteps; p.cs_out[SP_LASTPROB * C + chain] = lastprob;
    }
}

struct UF {
    std::vector<int> p;
    explicit UF(int n) : p(n) { std::iota(p.begin(), p.end(), 0); }
    int find(int x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
    void unite(int a, int b) { a = find(a); b = find(b); if (a != b) p[std::max(a, b)] = std::min(a, b); }
};

template <int FL, int L, int RL, int CL, bool TRI = false>
int launch_lane(gmb_ctx* ctx, const LaneParams& p) {
    const int chains_per_cta = LN_WPB * p.cpw;
    hmc_lane_kernel<FL, L, RL, CL, TRI><<<(p.C + chains_per_cta - 1) / chains_per_cta, 32 * LN_WPB, 0, ctx->stream>>>(p);
    ctx->launches++;
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}

// (L, slots): lanes per component and rows = columns per lane
template <int FL>
int launch_lane_shape(gmb_ctx* ctx, const LaneParams& p, int L, int slots, bool tri) {
    if (L == 1 && tri) {
        if (slots == 4) return launch_lane<FL, 1, 4, 4, true>(ctx, p);
        if (slots == 5) return launch_lane<FL, 1, 5, 5, true>(ctx, p);
        if (slots == 6) return launch_lane<FL, 1, 6, 6, true>(ctx, p);
    }
    if (L == 1 && slots == 4) return launch_lane<FL, 1, 4, 4>(ctx, p);
    if (L == 1 && slots == 5) return launch_lane<FL, 1, 5, 5>(ctx, p);
    if (L == 1 && slots == 6) return launch_lane<FL, 1, 6, 6>(ctx, p);
    if (L == 2 && slots == 2) return launch_lane<FL, 2, 2, 2>(ctx, p);
    if (L == 2 && slots == 3) return launch_lane<FL, 2, 3, 3>(ctx, p);
    if (L == 3 && slots == 1) return launch_lane<FL, 3, 1, 1>(ctx, p);
    if (L == 3 && slots == 2) return launch_lane<FL, 3, 2, 2>(ctx, p);
    return gmb_set_error(GMB_EINVAL, "lane-per-component sampler: no kernel for components of this size");
}

}  // namespace

void gmb_lane_free(gmb_model* mdl) {
    gmb_dfree(mdl->ctx, mdl->lane.dint); gmb_dfree(mdl->ctx, mdl->lane.dval);
    mdl->lane = gmb_lane();
}

// Connected components of the view's Z L (from its ELL form); the lane kernel applies when there are at most 32 of them with at most 6 rows and
// 6 columns each.  Rebuilt whenever the factor changes (gmb_sparse_invalidate).
int gmb_lane_ensure(gmb_model* mdl) {
    gmb_lane& k = mdl->lane;
    if (k.checked) return GMB_OK;
    k.valid = false; k.checked = true;
    const gmb_ell& e = mdl->ell;
    if (!e.valid || e.ng > 32 * 6 || e.Q > 32 * 6) return GMB_OK;
    gmb_ctx* ctx = mdl->ctx;
    const int ng = e.ng, Q = e.Q, ngp = e.ngp, wr = e.wr;
    std::vector<double> rv((size_t)std::max(wr, 1) * ngp); std::vector<int> rc((size_t)std::max(wr, 1) * ngp);
    if (wr > 0) {
        GMB_CUDA(cudaMemcpyAsync(rv.data(), e.rv, sizeof(double) * (size_t)wr * ngp, cudaMemcpyDeviceToHost, ctx->stream));
        GMB_CUDA(cudaMemcpyAsync(rc.data(), e.rc, sizeof(int) * (size_t)wr * ngp, cudaMemcpyDeviceToHost, ctx->stream));
    }
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    UF uf(ng + Q);
    for (int w = 0; w < wr; w++)
        for (int i = 0; i < ng; i++) if (rv[(size_t)w * ngp + i] != 0.0) uf.unite(i, ng + rc[(size_t)w * ngp + i]);
    std::vector<int> id(ng + Q, -1), comp_of(ng + Q), nr, nc;
    for (int x = 0; x < ng + Q; x++) {
        const int root = uf.find(x);
        if (id[root] < 0) { id[root] = (int)nr.size(); nr.push_back(0); nc.push_back(0); }
        comp_of[x] = id[root];
        if (x < ng) nr[comp_of[x]]++; else nc[comp_of[x]]++;
    }
    const int ncomp = (int)nr.size();
    if (ncomp < 1 || ncomp > 32) return GMB_OK;
    const int mr = *std::max_element(nr.begin(), nr.end()), mc = *std::max_element(nc.begin(), nc.end());
    const int big = std::max(mr, mc);
    if (big > 6) return GMB_OK;
    // lanes per component: as many as fit 32 lanes per chain (the FP64 unit issues per warp, not per lane: spread the work of a chain wide)
    // L = 1 measured best on B200 (C2, 1000 chains: 14.3 ms per draw against 19.8 ms for L = 2 and 17.2 ms for L = 3 — the all-gather shuffles put the
    // dependent chain of a step back on the critical path with fewer than two warps per scheduler); GMB_LANE_L selects the split variants
    int L = 1;
    { static const int forced = [] { const char* e = getenv("GMB_LANE_L"); return e ? atoi(e) : 0; }();      // measurement override
      if (forced >= 1 && forced <= 3 && ncomp * forced <= 32 && big >= forced) L = forced; }
    int slots = (big + L - 1) / L;
    if (L == 1 && slots < 4) slots = 4;
    if (L == 2 && slots < 2) slots = 2;
    const int RL = slots, CL = slots, Rp = RL * L, Cp = CL * L, lpc = ncomp * L;
    // component-local numbering in ascending row / column order; local row r belongs to lane r % L of the group, slot r / L
    std::vector<int> fr(ncomp, 0), fc(ncomp, 0), lrow(ng), lcol(Q);
    std::vector<int> rowid((size_t)lpc * RL, -1), colid((size_t)lpc * CL, -1);
    for (int i = 0; i < ng; i++) { const int q = comp_of[i], r = fr[q]++; lrow[i] = r; rowid[((size_t)q * L + r % L) * RL + r / L] = i; }
    for (int j = 0; j < Q; j++) { const int q = comp_of[ng + j], c = fc[q]++; lcol[j] = c; colid[((size_t)q * L + c % L) * CL + c / L] = j; }
    std::vector<double> brow((size_t)lpc * RL * Cp, 0.0), bcol((size_t)lpc * Rp * CL, 0.0);
    for (int w = 0; w < wr; w++)
        for (int i = 0; i < ng; i++) {
            const double v = rv[(size_t)w * ngp + i];
            if (v == 0.0) continue;
            const int q = comp_of[i], r = lrow[i], c = lcol[rc[(size_t)w * ngp + i]];
            brow[(((size_t)q * L + r % L) * RL + r / L) * Cp + c] = v;           // row r of the block, all columns
            bcol[(((size_t)q * L + c % L) * Rp + r) * CL + c / L] = v;           // column c of the block, all rows
        }
    const size_t ni = rowid.size() + colid.size(), nv = brow.size() + bcol.size();
    if (ni > k.int_cap) { if (k.dint) { gmb_dfree(ctx, k.dint); k.dint = nullptr; } GMB_CUDA(gmb_dmalloc(ctx, &k.dint, sizeof(int) * ni)); k.int_cap = ni; }
    if (nv > k.val_cap) { if (k.dval) { gmb_dfree(ctx, k.dval); k.dval = nullptr; } GMB_CUDA(gmb_dmalloc(ctx, &k.dval, sizeof(double) * nv)); k.val_cap = nv; }
    std::vector<int> hi(ni);
    std::copy(rowid.begin(), rowid.end(), hi.begin());
    std::copy(colid.begin(), colid.end(), hi.begin() + rowid.size());
    std::vector<double> hv(nv);
    std::copy(brow.begin(), brow.end(), hv.begin());
    std::copy(bcol.begin(), bcol.end(), hv.begin() + brow.size());
    GMB_CUDA(cudaMemcpyAsync(k.dint, hi.data(), sizeof(int) * ni, cudaMemcpyHostToDevice, ctx->stream));
    GMB_CUDA(cudaMemcpyAsync(k.dval, hv.data(), sizeof(double) * nv, cudaMemcpyHostToDevice, ctx->stream));
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    bool tri = true;                                                 // lower triangular in the local (ascending) numbering?
    for (int w = 0; w < wr && tri; w++)
        for (int i = 0; i < ng; i++) if (rv[(size_t)w * ngp + i] != 0.0 && lcol[rc[(size_t)w * ngp + i]] > lrow[i]) { tri = false; break; }
    k.ncomp = ncomp; k.L = L; k.slots = slots; k.tri = tri; k.valid = true;
    return GMB_OK;
}

bool gmb_hmc_lane_applicable(const gmb_model* mdl) { return g_hmc_lane && mdl->lane.checked && mdl->lane.valid; }

// Same contract as gmb_hmc_run_sparse: dV_out is ldq x (C * (nsamp + 1)) chain-major; d_cs holds the SP_COUNT x C chain statistics.
int gmb_hmc_run_lane(gmb_model* mdl, double var_par, int warmup, int nsamp, double lambda, int max_steps, double target_accept,
                     int adapt, int C, uint32_t chain_offset, uint64_t seed, double* dV_out, double* d_cs) {
    gmb_ctx* ctx = mdl->ctx;
    const gmb_lane& k = mdl->lane;
    const gmb_agg& a = mdl->agg;
    if (!k.valid) return gmb_set_error(GMB_ESTATE, "lane-per-component sampler: not applicable to this model");
    LaneParams p;
    const int lpc = k.ncomp * k.L;
    p.lpc = lpc; p.cpw = 32 / lpc;
    p.rowid = k.dint; p.colid = k.dint + (size_t)lpc * k.slots;
    p.brow = k.dval; p.bcol = k.dval + (size_t)lpc * k.slots * (k.slots * k.L);
    p.xb = a.active ? a.dxb : mdl->dxb;
    p.cnt = a.dcnt; p.ys = a.dys; p.lcnt = a.dlcnt; p.lys = a.dlys; p.lsq = a.dlsq; p.lrc = a.dlrc;
    p.var_par = var_par; p.lambda = lambda; p.target_accept = target_accept;
    p.warmup = warmup; p.nsamp = nsamp; p.max_steps = max_steps; p.adapt = adapt; p.C = C; p.ldq = mdl->ldq;
    p.chain_offset = chain_offset; p.seed = seed; p.dV_out = dV_out; p.cs_out = d_cs;
    switch (mdl->flink) {
    case 1: return launch_lane_shape<1>(ctx, p, k.L, k.slots, k.tri);
    case 3: return launch_lane_shape<3>(ctx, p, k.L, k.slots, k.tri);
    case 7: return launch_lane_shape<7>(ctx, p, k.L, k.slots, k.tri);
    }
    return gmb_set_error(GMB_EFAMILY, "family/link code %d has no device kernel", mdl->flink);
}
