// hmc_comp.cu — structure-aware sampler for LARGE sparse models: the trajectory decomposes over the connected components of Z L.
//
// The gradient of the log-density couples a random effect only with the observations it enters and, through them, with the other random
// effects of the same connected component of the bipartite graph (rows, columns) of Z L.  With a block-diagonal D and an indicator Z those
// components are small (config C4: n = Q = 10^4, 1000 components of 10 rows and 10 columns), and a leapfrog trajectory — whose step size and
// length are per-chain scalars — can be integrated for every component on its own; only the Metropolis test (mhmcmc.h:80-86) needs sums
// over all of them.  The CTA-per-chain kernel of hmc_sparse.cu streams the whole ELL form of Z L from L2 on every leapfrog step (C4: 2.4 MB
// per step and chain, L2-bandwidth bound).  Here components are packed into groups of <= 32 rows and columns, and per proposal
//
//   comp_traj_kernel    one WARP per (chain, group): its ELL entries in registers, momentum draw, the whole trajectory with v' and r(eta)
//                       exchanged through shared memory, candidate state + gradient and five partial sums written out
//   comp_decide_kernel  one CTA per chain: fixed-order sums over the groups, Metropolis test, dual-averaging step size (:107-117), number
//                       of steps of the next proposal, sample column
//
// so Z L is read from memory once per launch instead of once per leapfrog step.  Chain arithmetic and random streams are those of the other
// sampler kernels (same Philox counters per column, iteration and chain); sums over rows / columns run per group and then over groups.
#include "hmc_sparse_common.cuh"
#include <algorithm>
#include <numeric>
#include <type_traits>

namespace {

constexpr int CP_CAP = 32;       // rows and columns per group (one lane each)
constexpr int CP_WPB = 4;        // warps per CTA of the trajectory kernel
constexpr int CP_MAXW = 12;      // ELL width kept in registers

struct CompParams {
    int G, C, Qp, ldq, W;
    const int* grow; const int* gcol;            // [G][32] view row / column of a lane, -1 for none
    const double* lrv; const int* lrc;           // [G][W][32]: by rows, local column as byte offset
    const double* lcv; const int* lcr;           // [G][W][32]: by columns, local row as byte offset
    const double* xb; const double* cnt; const double* ys;
    const double* lcnt; const double* lys; const double* lsq; const double* lrc2;
    double var_par, lambda, target_accept;
    int warmup, nsamp, max_steps, adapt;
    uint32_t chain_offset; unsigned long long seed;
    double* V; double* Gd;                       // [2][C][Qp]: state / gradient, parity cur[c] = current, the other = candidate
    double* part;                                // [C][G][5]: k0, pv, k1, pvp, ll
    double* cs;                                  // [SP_COUNT][C]
    int* cur; int* steps;                        // [C]
    double* dV_out;
};

// mode 0: initial state (mhmcmc.h:48-49), its gradient and log-likelihood; mode 1: proposal `it` (:61-78)
template <int FL, int W>
__global__ void __launch_bounds__(32 * CP_WPB) comp_traj_kernel(const CompParams p, int mode, int it) {
    __shared__ double sTab[16];                  // 2^(j/16): the conflict-free table of dev_family_resid_w_vec16
    __shared__ double sX[CP_WPB][2][CP_CAP];     // v' and r(eta) of the warp's group
    const int tid = threadIdx.x, t = tid & 31, w = tid >> 5;
    if (tid < 16) sTab[tid] = GMB_EXP2_TAB[4 * tid];
    __syncthreads();
    const long long wg = (long long)blockIdx.x * CP_WPB + w;
    if (wg >= (long long)p.C * p.G) return;
    const int c = (int)(wg / p.G), g = (int)(wg % p.G);
    const uint32_t gchain = p.chain_offset + (uint32_t)c;
    double* s_vp = sX[w][0]; double* s_res = sX[w][1];
    const char* s_vp_b = reinterpret_cast<const char*>(s_vp); const char* s_res_b = reinterpret_cast<const char*>(s_res);

    const int i = p.grow[g * CP_CAP + t], j = p.gcol[g * CP_CAP + t];
    const bool hasr = i >= 0, hasc = j >= 0;
    const double xb = hasr ? p.xb[i] : 0.0, cn = hasr ? p.cnt[i] : 0.0, ysv = hasr ? p.ys[i] : 0.0;
    double erv[W], ecv[W]; int erc[W], ecr[W];
#pragma unroll
    for (int k = 0; k < W; k++) {
        const bool ok = k < p.W;
        const size_t o = ((size_t)g * p.W + k) * CP_CAP + t;
        erv[k] = ok ? p.lrv[o] : 0.0; erc[k] = ok ? p.lrc[o] : 0;
        ecv[k] = ok ? p.lcv[o] : 0.0; ecr[k] = ok ? p.lcr[o] : 0;
    }
    const double sigma = p.var_par;
    const double sc = (FL == 7) ? 1.0 / (sigma * sigma) : 1.0;
    const double c0 = (FL == 7) ? (-1.0 * log(sigma) - 0.5 * log(2 * GMB_PI_FAMILY)) : 0.0;
    const double pc = -1.0 * log(1.0) - 0.5 * log(2 * GMB_PI_FAMILY);   // log_likelihood(v, 0, 1, 7), mcmlmodel.h:149

    double vp = 0.0, r = 0.0, gr = 0.0, ll = 0.0;
    auto grad_eval = [&](auto ll_tag) {
        constexpr bool with_ll = decltype(ll_tag)::value;
        double eta[1] = {xb}, eo = 0.0;
#pragma unroll
        for (int k = 0; k < W; k += 2) {
            eta[0] = fma(erv[k], *reinterpret_cast<const double*>(s_vp_b + erc[k]), eta[0]);
            eo = fma(erv[k + 1], *reinterpret_cast<const double*>(s_vp_b + erc[k + 1]), eo);
        }
        eta[0] += eo;
        const double cn1[1] = {cn}, ys1[1] = {ysv};
        double res[1];
        dev_family_resid_w_vec16<FL, 1>(cn1, ys1, eta, sTab, res);
        s_res[t] = res[0];
        if (with_ll) {
            ll = 0.0;
            if (hasr) {
                const double lq = (FL == 7) ? p.lsq[i] : 0.0, lr = (FL == 1) ? p.lrc2[i] : 0.0;
                ll = dev_family_ll_w<FL>(p.lcnt[i], p.lys[i], lq, lr, eta[0], c0, sigma);
            }
        }
        __syncwarp();
        double gs = 0.0, go = 0.0;
#pragma unroll
        for (int k = 0; k < W; k += 2) {
            gs = fma(ecv[k], *reinterpret_cast<const double*>(s_res_b + ecr[k]), gs);
            go = fma(ecv[k + 1], *reinterpret_cast<const double*>(s_res_b + ecr[k + 1]), go);
        }
        gr = -1.0 * vp + sc * (gs + go);                           // mcmlmodel.h:163 + :173/:191/:235
    };

    double* prt = p.part + ((size_t)c * p.G + g) * 5;
    if (mode == 0) {
        if (hasc) {
            double z0, z1;
            dev_rng_normal2(p.seed, (uint32_t)(j >> 1), 0u, gchain, 0u, z0, z1);
            vp = (j & 1) ? z1 : z0;
        }
        s_vp[t] = vp;
        __syncwarp();
        grad_eval(std::true_type{});
        if (hasc) { p.V[(size_t)c * p.Qp + j] = vp; p.Gd[(size_t)c * p.Qp + j] = gr; }
        ll = warp_sum(ll);
        if (t == 0) { prt[0] = prt[1] = prt[2] = prt[3] = 0.0; prt[4] = ll; }
        return;
    }
    const int par = p.cur[c], steps = p.steps[c];
    const double eps = p.cs[SP_EPS * p.C + c];
    const size_t o_cur = ((size_t)par * p.C + c) * p.Qp, o_cand = ((size_t)(par ^ 1) * p.C + c) * p.Qp;
    double k0 = 0.0, pv = 0.0;
    if (hasc) {                                                    // new_proposal, mhmcmc.h:61-75
        double z0, z1;
        dev_rng_normal2(p.seed, (uint32_t)(j >> 1), (uint32_t)it, gchain, 2u, z0, z1);
        const double z = (j & 1) ? z1 : z0;
        const double vq = p.V[o_cur + j], gq = p.Gd[o_cur + j];
        k0 = z * z; pv = pc - 0.5 * vq * vq;
        r = z + (eps / 2) * gq;                                    // :74 (first step)
        vp = vq + eps * r;                                         // :67, :75
    }
    s_vp[t] = vp;
    __syncwarp();
    for (int s = 0; s + 1 < steps; s++) {                          // leapfrog, :73-78 (padding lanes: g = r = v' = 0 throughout)
        grad_eval(std::false_type{});
        double rr = r + (eps / 2) * gr;                            // :77
        rr = rr + (eps / 2) * gr;                                  // :74 of the next step
        vp = vp + eps * rr;                                        // :75
        s_vp[t] = vp; r = rr;
        __syncwarp();
    }
    grad_eval(std::true_type{});
    r = r + (eps / 2) * gr;                                        // :77 (last step)
    double k1 = 0.0, pvp = 0.0;
    if (hasc) {
        k1 = r * r; pvp = pc - 0.5 * vp * vp;
        p.V[o_cand + j] = vp; p.Gd[o_cand + j] = gr;               // candidate, adopted by comp_decide_kernel if accepted
    }
    k0 = warp_sum(k0); pv = warp_sum(pv); k1 = warp_sum(k1); pvp = warp_sum(pvp); ll = warp_sum(ll);
    if (t == 0) { prt[0] = k0; prt[1] = pv; prt[2] = k1; prt[3] = pvp; prt[4] = ll; }
}

// one CTA per chain.  it = -1: after the initial evaluation (initialise_u, mhmcmc.h:47-59); otherwise the end of proposal `it` (:80-117, :142-147)
__global__ void __launch_bounds__(128) comp_decide_kernel(const CompParams p, int it) {
    __shared__ double red[5][4];
    __shared__ int s_cur;
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, C = p.C;
    double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int g = tid; g < p.G; g += 128) {
        const double* prt = p.part + ((size_t)c * p.G + g) * 5;
#pragma unroll
        for (int k = 0; k < 5; k++) a[k] += prt[k];
    }
#pragma unroll
    for (int k = 0; k < 5; k++) { a[k] = warp_sum(a[k]); if (lane == 0) red[k][w] = a[k]; }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 5; k++) a[k] = ((red[k][0] + red[k][1]) + red[k][2]) + red[k][3];
        double eps, ebar, H;
        int cur = 0;
        if (it < 0) {
            eps = 0.001; ebar = 1.0; H = 0.0;
            p.cs[SP_LLCUR * C + c] = a[4]; p.cs[SP_ACCEPT * C + c] = 0.0; p.cs[SP_TOTSTEPS * C + c] = 0.0; p.cs[SP_LASTPROB * C + c] = 0.0;
            p.cs[SP_K0 * C + c] = 0.0;
        } else {
            eps = p.cs[SP_EPS * C + c]; ebar = p.cs[SP_EBAR * C + c]; H = p.cs[SP_H * C + c];
            cur = p.cur[c];
            const double k0 = 0.5 * a[0], k1 = 0.5 * a[2];                                          // :66
            const double l1 = p.cs[SP_LLCUR * C + c] + a[1], l2 = a[4] + a[3];                      // :82-83
            const double prob = fmin(1.0, exp(-l1 + k0 + l2 - k1));                                 // :84
            double u1, u2;
            dev_rng_uniform2(p.seed, 0u, (uint32_t)it, p.chain_offset + (uint32_t)c, 3u, u1, u2);   // :85
            if (u1 < prob) { cur ^= 1; p.cs[SP_LLCUR * C + c] = a[4]; p.cs[SP_ACCEPT * C + c] += 1.0; }   // :86, :102-105
            p.cs[SP_LASTPROB * C + c] = prob;
            if (it < p.warmup && it < p.adapt) {                                                    // :107-114, :131-136
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
                eps = ebar;                                                                         // :115-117
            }
        }
        p.cs[SP_EPS * C + c] = eps; p.cs[SP_EBAR * C + c] = ebar; p.cs[SP_H * C + c] = H;
        p.cur[c] = cur;
        if (it + 1 < p.warmup + p.nsamp) {                                                          // the next proposal, :69-70
            const double sd = round(p.lambda / eps);
            int st = sd >= (double)p.max_steps ? p.max_steps : (sd < 1.0 ? 1 : (int)sd);
            if (!(sd == sd)) st = p.max_steps;
            p.steps[c] = st;
            p.cs[SP_TOTSTEPS * C + c] += st;
        }
        s_cur = cur;
    }
    __syncthreads();
    const int col = it - p.warmup + 1;                                                              // :142 (col 0), :147
    if (col >= 0) {
        const double* V = p.V + ((size_t)s_cur * C + c) * p.Qp;
        const int cols = p.nsamp + 1;
        for (int j = tid; j < p.Qp && j < p.ldq; j += 128) p.dV_out[((size_t)c * cols + col) * p.ldq + j] = V[j];
    }
}

template <int FL>
int launch_traj(gmb_ctx* ctx, const CompParams& p, int mode, int it) {
    const long long warps = (long long)p.C * p.G;
    const unsigned grid = (unsigned)((warps + CP_WPB - 1) / CP_WPB);
    if (p.W <= 2) comp_traj_kernel<FL, 2><<<grid, 32 * CP_WPB, 0, ctx->stream>>>(p, mode, it);
    else if (p.W <= 4) comp_traj_kernel<FL, 4><<<grid, 32 * CP_WPB, 0, ctx->stream>>>(p, mode, it);
    else if (p.W <= 6) comp_traj_kernel<FL, 6><<<grid, 32 * CP_WPB, 0, ctx->stream>>>(p, mode, it);
    else if (p.W <= 8) comp_traj_kernel<FL, 8><<<grid, 32 * CP_WPB, 0, ctx->stream>>>(p, mode, it);
    else if (p.W <= 10) comp_traj_kernel<FL, 10><<<grid, 32 * CP_WPB, 0, ctx->stream>>>(p, mode, it);
    else comp_traj_kernel<FL, 12><<<grid, 32 * CP_WPB, 0, ctx->stream>>>(p, mode, it);
    ctx->launches++;
    return GMB_OK;
}

struct UF {
    std::vector<int> p;
    explicit UF(int n) : p(n) { std::iota(p.begin(), p.end(), 0); }
    int find(int x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
    void unite(int a, int b) { a = find(a); b = find(b); if (a != b) p[std::max(a, b)] = std::min(a, b); }
};

}  // namespace

void gmb_comp_free(gmb_model* mdl) {
    gmb_comp& k = mdl->comp;
    gmb_ctx* ctx = mdl->ctx;
    gmb_dfree(ctx, k.dint); gmb_dfree(ctx, k.dval);
    k = gmb_comp();
}

// Connected components of the view's Z L and their packing into groups (rebuilt with the ELL form whenever the factor changes).
// Applies to models too large for the warp-per-chain kernels whose components all fit 32 rows and 32 columns with ELL widths <= 12.
int gmb_comp_ensure(gmb_model* mdl) {
    gmb_comp& k = mdl->comp;
    if (k.checked) return GMB_OK;
    k.valid = false; k.checked = true;
    const gmb_ell& e = mdl->ell;
    if (!e.valid || std::max(e.ng, e.Q) <= 128 || e.wr > CP_MAXW || e.wc > CP_MAXW) return GMB_OK;
    gmb_ctx* ctx = mdl->ctx;
    const int ng = e.ng, Q = e.Q, ngp = e.ngp, wr = e.wr;
    std::vector<double> rv((size_t)wr * ngp); std::vector<int> rc((size_t)wr * ngp);
    if (wr > 0) {
        GMB_CUDA(cudaMemcpyAsync(rv.data(), e.rv, sizeof(double) * rv.size(), cudaMemcpyDeviceToHost, ctx->stream));
        GMB_CUDA(cudaMemcpyAsync(rc.data(), e.rc, sizeof(int) * rc.size(), cudaMemcpyDeviceToHost, ctx->stream));
    }
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    UF uf(ng + Q);
    for (int w = 0; w < wr; w++)
        for (int i = 0; i < ng; i++) if (rv[(size_t)w * ngp + i] != 0.0) uf.unite(i, ng + rc[(size_t)w * ngp + i]);
    // components in order of their smallest node (rows first, then column-only components)
    std::vector<int> comp_of(ng + Q), nr, nc;
    std::vector<int> id(ng + Q, -1);
    for (int x = 0; x < ng + Q; x++) {
        const int root = uf.find(x);
        if (id[root] < 0) { id[root] = (int)nr.size(); nr.push_back(0); nc.push_back(0); }
        comp_of[x] = id[root];
        if (x < ng) nr[comp_of[x]]++; else nc[comp_of[x]]++;
    }
    const int ncomp = (int)nr.size();
    for (int q = 0; q < ncomp; q++) if (nr[q] > CP_CAP || nc[q] > CP_CAP) return GMB_OK;
    // greedy packing in component order
    std::vector<int> group_of(ncomp);
    int G = 0, fr = 0, fc = 0;
    for (int q = 0; q < ncomp; q++) {
        if (q == 0 || fr + nr[q] > CP_CAP || fc + nc[q] > CP_CAP) { G++; fr = fc = 0; }
        group_of[q] = G - 1; fr += nr[q]; fc += nc[q];
    }
    if (G < 2) return GMB_OK;
    const int W = std::max(2, (std::max(e.wr, e.wc) + 1) / 2 * 2);
    // local numbering: rows / columns of a group in ascending order
    std::vector<int> grow((size_t)G * CP_CAP, -1), gcol((size_t)G * CP_CAP, -1), lrow(ng), lcol(Q), fillr(G, 0), fillc(G, 0);
    for (int i = 0; i < ng; i++) { const int g = group_of[comp_of[i]]; lrow[i] = fillr[g]; grow[(size_t)g * CP_CAP + fillr[g]++] = i; }
    for (int j = 0; j < Q; j++) { const int g = group_of[comp_of[ng + j]]; lcol[j] = fillc[g]; gcol[(size_t)g * CP_CAP + fillc[g]++] = j; }
    const size_t ne = (size_t)G * W * CP_CAP;
    std::vector<double> lrv(ne, 0.0), lcv(ne, 0.0); std::vector<int> lrc(ne, 0), lcr(ne, 0);
    std::vector<int> cfill(Q, 0);
    for (int i = 0; i < ng; i++) {                                  // ascending rows => the column lists come out in ascending row order
        const int g = group_of[comp_of[i]];
        int wl = 0;
        for (int w = 0; w < wr; w++) {
            const double v = rv[(size_t)w * ngp + i];
            if (v == 0.0) continue;
            const int j = rc[(size_t)w * ngp + i];
            lrv[((size_t)g * W + wl) * CP_CAP + lrow[i]] = v; lrc[((size_t)g * W + wl) * CP_CAP + lrow[i]] = 8 * lcol[j]; wl++;
            const int wc2 = cfill[j]++;
            if (wc2 >= W) return gmb_set_error(GMB_ESTATE, "component sampler: column width exceeds the ELL width");
            lcv[((size_t)g * W + wc2) * CP_CAP + lcol[j]] = v; lcr[((size_t)g * W + wc2) * CP_CAP + lcol[j]] = 8 * lrow[i];
        }
    }
    // one device allocation for the integer arrays, one for the values
    const size_t ni = 2 * (size_t)G * CP_CAP + 2 * ne, nv = 2 * ne;
    if (ni > k.int_cap) { if (k.dint) { GMB_CUDA(cudaStreamSynchronize(ctx->stream)); gmb_dfree(ctx, k.dint); k.dint = nullptr; } GMB_CUDA(gmb_dmalloc(ctx, &k.dint, sizeof(int) * ni)); k.int_cap = ni; }
    if (nv > k.val_cap) { if (k.dval) { GMB_CUDA(cudaStreamSynchronize(ctx->stream)); gmb_dfree(ctx, k.dval); k.dval = nullptr; } GMB_CUDA(gmb_dmalloc(ctx, &k.dval, sizeof(double) * nv)); k.val_cap = nv; }
    std::vector<int> hi(ni);
    std::copy(grow.begin(), grow.end(), hi.begin());
    std::copy(gcol.begin(), gcol.end(), hi.begin() + (size_t)G * CP_CAP);
    std::copy(lrc.begin(), lrc.end(), hi.begin() + 2 * (size_t)G * CP_CAP);
    std::copy(lcr.begin(), lcr.end(), hi.begin() + 2 * (size_t)G * CP_CAP + ne);
    std::vector<double> hv(nv);
    std::copy(lrv.begin(), lrv.end(), hv.begin());
    std::copy(lcv.begin(), lcv.end(), hv.begin() + ne);
    GMB_CUDA(cudaMemcpyAsync(k.dint, hi.data(), sizeof(int) * ni, cudaMemcpyHostToDevice, ctx->stream));
    GMB_CUDA(cudaMemcpyAsync(k.dval, hv.data(), sizeof(double) * nv, cudaMemcpyHostToDevice, ctx->stream));
    GMB_CUDA(cudaStreamSynchronize(ctx->stream));
    k.G = G; k.W = W; k.ncomp = ncomp; k.valid = true;
    return GMB_OK;
}

bool gmb_hmc_comp_applicable(const gmb_model* mdl) { return mdl->comp.checked && mdl->comp.valid; }

size_t gmb_hmc_comp_work_doubles(const gmb_model* mdl, int C) {
    const size_t Qp = mdl->ell.qp;
    return round_up_sz((size_t)SP_COUNT * C + 16, 16) + 4 * (size_t)C * Qp + (size_t)C * mdl->comp.G * 5 + (size_t)C + 16;
}

// Same contract as gmb_hmc_run_sparse: dV_out is ldq x (C * (nsamp + 1)) chain-major; work starts with the SP_COUNT x C chain statistics.
int gmb_hmc_run_comp(gmb_model* mdl, double var_par, int warmup, int nsamp, double lambda, int max_steps, double target_accept,
                     int adapt, int C, uint32_t chain_offset, uint64_t seed, double* dV_out, double* work) {
    gmb_ctx* ctx = mdl->ctx;
    const gmb_comp& k = mdl->comp;
    const gmb_agg& a = mdl->agg;
    if (!k.valid) return gmb_set_error(GMB_ESTATE, "component sampler: not applicable to this model");
    CompParams p;
    p.G = k.G; p.C = C; p.Qp = mdl->ell.qp; p.ldq = mdl->ldq; p.W = k.W;
    const size_t ne = (size_t)k.G * k.W * CP_CAP;
    p.grow = k.dint; p.gcol = k.dint + (size_t)k.G * CP_CAP; p.lrc = k.dint + 2 * (size_t)k.G * CP_CAP; p.lcr = p.lrc + ne;
    p.lrv = k.dval; p.lcv = k.dval + ne;
    p.xb = a.active ? a.dxb : mdl->dxb;
    p.cnt = a.dcnt; p.ys = a.dys; p.lcnt = a.dlcnt; p.lys = a.dlys; p.lsq = a.dlsq; p.lrc2 = a.dlrc;
    p.var_par = var_par; p.lambda = lambda; p.target_accept = target_accept;
    p.warmup = warmup; p.nsamp = nsamp; p.max_steps = max_steps; p.adapt = adapt;
    p.chain_offset = chain_offset; p.seed = seed; p.dV_out = dV_out;
    double* w = work;
    p.cs = w; w += round_up_sz((size_t)SP_COUNT * C + 16, 16);
    p.V = w; w += 2 * (size_t)C * p.Qp; p.Gd = w; w += 2 * (size_t)C * p.Qp;
    p.part = w; w += (size_t)C * k.G * 5;
    p.cur = reinterpret_cast<int*>(w); p.steps = p.cur + C;
    GMB_CUDA(cudaMemsetAsync(p.V, 0, sizeof(double) * 4 * (size_t)C * p.Qp, ctx->stream));
    GMB_CUDA(cudaMemsetAsync(p.cur, 0, sizeof(int) * 2 * (size_t)C, ctx->stream));
    const int total = warmup + nsamp;
    for (int it = -1; it < total; it++) {
        const int mode = it < 0 ? 0 : 1;
        switch (mdl->flink) {
        case 1: GMB_TRY(launch_traj<1>(ctx, p, mode, it)); break;
        case 3: GMB_TRY(launch_traj<3>(ctx, p, mode, it)); break;
        case 7: GMB_TRY(launch_traj<7>(ctx, p, mode, it)); break;
        default: return gmb_set_error(GMB_EFAMILY, "family/link code %d has no device kernel", mdl->flink);
        }
        comp_decide_kernel<<<C, 128, 0, ctx->stream>>>(p, it);
        ctx->launches++;
    }
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}
