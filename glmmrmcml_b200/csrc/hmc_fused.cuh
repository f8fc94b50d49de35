// hmc_fused.cuh — on-chip variant of the batched random-effect sampler (mhmcmc.h:16-160): ONE launch runs the whole
// sample(warmup, nsamp) call, 8 chains per group, with Z L resident in shared memory.
//
// Per leapfrog step a group makes a SINGLE pass over its Z L rows for all 8 chains (mcmlmodel.h:156-279):
//   for every tile of 8 observations (tiles interleaved over the 8 warps, up to 4 tiles in flight per warp)
//     eta  = xb + (Z L) v'            DMMA m8n8k4, A = Z L rows from shared memory, B = v' fragments in registers
//     res  = r(eta)                   family residual in the accumulator registers (table-driven exp + Newton reciprocal,
//                                     branch free so that the 8 residuals of 4 tiles interleave on the FP64 pipe; interleaving
//                                     them with the tensor instructions of other tiles by hand was measured and gave nothing)
//     G   += (Z L)^T res              DMMA m8n8k4 again: the SAME shared-memory rows read transposed, res moved from
//                                     accumulator to B-fragment layout with warp shuffles
// so eta and res never exist in memory, Z L is read from shared memory only, and the two contractions of the
// two-GEMM variant (hmc.cu) plus their epilogues become one loop.  The per-warp partial gradients are summed in a
// fixed order through shared memory (deterministic).
//
// Chain state.  v' and the momentum r of all 8 chains live in registers in the MMA B-fragment layout (lane (fr, fk) holds
// elements q = 4 j + fk of chain fr), REPLICATED in every warp: after the gradient sum every warp reads the gradient in
// that layout and applies the leapfrog update (mhmcmc.h:73-78) itself, so v' never goes through shared memory and no
// barrier separates one step's update from the next step's products.  Metropolis test (:80-105) and dual-averaging step
// size (:107-117) are likewise evaluated redundantly from replicated per-chain scalars (counter-based RNG: every replica
// draws the same numbers).  The current state v and its gradient, needed once per proposal, sit in a small global
// scratch area written by the "owner" lanes of each chain (warp w, fr == w).
//
// Shared-memory budget (227 KB): Z L is stored row-major with a row stride ld = 4 (mod 16) doubles, which makes both
// the A-fragment reads of the forward product (8 rows x 4 k) and of the transposed product (4 rows x 8 q) conflict
// free.
//
// Thread-block clusters (CS = 2 or 4 CTAs per group of 8 chains): the observations are split over the CTAs of a
// cluster, each keeps only its rows of Z L (and of xb and the row weights) in shared memory and computes a partial gradient; the
// partials are exchanged through distributed shared memory (one remote store per gradient element, double buffered)
// and one cluster barrier per leapfrog step, after which every CTA of the cluster sums them in rank order and advances
// an identical replica of the chain state.  That (i) lets a group of 8 chains use CS SMs, so that a sampling run with
// fewer than 148 groups still fills the GPU, and (ii) extends the on-chip variant to models CS times larger than one
// SM's shared memory.  Models that do not fit even with CS = 4 take the two-GEMM path (hmc.cu).
#pragma once
#include "common.cuh"
#include <cstdlib>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

struct FusedParams {
    int n, Q, ld, ks, qt8, ldn, ldq, n8;          // n8: rows of Z L one CTA holds (tiles_per_cta * 8)
    int tiles_per_cta;
    int slots8;                                    // 1: one gradient slot per warp also without a cluster (the model is small enough)
    const double* ZL; const double* xb;            // the sampler's view of the model: n rows (distinct rows of [X | Z], aggregate.cu)
    const double* cnt; const double* ys;           // residual weights of the view's rows
    const double* lcnt; const double* lys; const double* lsq; const double* lrc;   // log-likelihood weights
    double var_par, lambda, target_accept;
    int warmup, nsamp, max_steps, adapt, C;
    uint32_t chain_offset; unsigned long long seed;
    double* dV_out; double* cs_out;
    double* scratch;                               // [grid][8 chains][2 parities][v | grad][ld] : state at the last accepted proposal
    long long* timing;                             // GMB_FUSED_TIMING builds only
};

// Optional per-phase cycle counters (build with EXTRA=-DGMB_FUSED_TIMING): thread 0 of every CTA accumulates clock64()
// differences per phase of a leapfrog step into FusedParams::timing[blockIdx.x * 8 + phase] (tools/hmc_phase_timing.py).
#ifdef GMB_FUSED_TIMING
struct FusedTim { long long acc[12]; long long last; bool rec; };
#define GMB_TICK(i) do { if (tim.rec) { const long long now__ = clock64(); tim.acc[i] += now__ - tim.last; tim.last = now__; } } while (0)
#else
struct FusedTim { };
#define GMB_TICK(i) do { } while (0)
#endif

namespace {

constexpr int CB = 8;          // chains per group = one MMA n-tile
constexpr int NWARP = 8;
constexpr int THREADS = NWARP * 32;

enum { FS_EPS = 0, FS_EBAR, FS_H, FS_LLCUR, FS_K0, FS_ACCEPT, FS_TOTSTEPS, FS_LASTPROB, FS_COUNT };   // = CS_* of hmc.cu

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- cluster exchange primitives: mbarrier with transaction count + st.async (the remote store itself signals the receiver) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, int rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// (.acquire.cluster is required: with the default CTA scope the st.async data of a peer CTA were observed stale — the chain
// parity test fails; the price is one CCTL.IVALL, an L1 invalidation, per wait: profiles/r01_ncu_source_hmc_fused_final.txt)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// 8-byte store into the shared memory of a CTA of the cluster that also signals 8 bytes on that CTA's mbarrier
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                 :: "r"(remote_addr), "l"(__double_as_longlong(v)), "r"(remote_bar) : "memory");
}

// sum over the 4 lanes that share fr (the 4 k-positions of a fragment)
__device__ __forceinline__ double quad_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// Gradient contributions of `NT` (1, 2 or 4) tiles of 8 observations, local rows r0[t] .. r0[t]+7, for the 8 chains of the group.
// MASK: rows >= nloc exist in the tile (only the last tile of a CTA whose row count is not a multiple of 8).
// LL: also accumulate the family log-likelihood of the chains that are on their last leapfrog step.
// SMROW: xb / cnt / ys point to shared memory (cluster variant) instead of global memory.  lw: the log-likelihood weights of this
// CTA's rows in global memory (read on the last leapfrog step of a chain only).
struct LLRows { const double* cnt; const double* ys; const double* sq; const double* rc; };
template <int FL, int KS, int NT, bool MASK, bool LL, bool SMROW>
__device__ __forceinline__ void fused_tiles(const double* __restrict__ sZL, const double* __restrict__ sTab, const double (&bf)[KS],
                                            const int (&r0)[NT], int nloc,
                                            const double* __restrict__ xb, const double* __restrict__ cnt, const double* __restrict__ ys,
                                            const LLRows lw, double c0, double sigma, bool want0, bool want1, int fr, int fk,
                                            double (&gacc)[(KS + 1) / 2][2], double& ll0, double& ll1, FusedTim& tim) {
    constexpr int LD = 4 * KS, QT8 = (KS + 1) / 2;
    double a[NT][2], xbv[NT], cv[NT], yv[NT];
#pragma unroll
    for (int t = 0; t < NT; t++) {
        const int row = r0[t] + fr;
        const bool ok = !MASK || row < nloc;
        if (SMROW) { xbv[t] = ok ? xb[row] : 0.0; cv[t] = ok ? cnt[row] : 0.0; yv[t] = ok ? ys[row] : 0.0; }
        else { xbv[t] = ok ? __ldg(xb + row) : 0.0; cv[t] = ok ? __ldg(cnt + row) : 0.0; yv[t] = ok ? __ldg(ys + row) : 0.0; }
        a[t][0] = a[t][1] = 0.0;
    }
    // eta tiles: rows x 8 chains; NT independent accumulator chains (splitting a tile's k range over several partial accumulators when
    // NT < 4 was measured: slower)
#pragma unroll
    for (int j = 0; j < KS; j++)
#pragma unroll
        for (int t = 0; t < NT; t++) dmma884(a[t][0], a[t][1], sZL[(r0[t] + fr) * LD + 4 * j + fk], bf[j]);
    GMB_TICK(1);                                       // eta products
    double res[NT][2];
#pragma unroll
    for (int t = 0; t < NT; t++) {
        const double eta0 = xbv[t] + a[t][0], eta1 = xbv[t] + a[t][1];
        res[t][0] = dev_family_resid_w<FL>(cv[t], yv[t], eta0, sTab);
        res[t][1] = dev_family_resid_w<FL>(cv[t], yv[t], eta1, sTab);
        if (MASK) { const bool ok = r0[t] + fr < nloc; if (!ok) { res[t][0] = 0.0; res[t][1] = 0.0; } }
        if (LL) {
            const bool ok = !MASK || r0[t] + fr < nloc;
            if (ok) {
                const int row = r0[t] + fr;
                const double lc = __ldg(lw.cnt + row), ly = __ldg(lw.ys + row);
                const double lq = (FL == 7) ? __ldg(lw.sq + row) : 0.0, lr = (FL == 1) ? __ldg(lw.rc + row) : 0.0;
                const double l0 = dev_family_ll_w<FL>(lc, ly, lq, lr, eta0, c0, sigma), l1 = dev_family_ll_w<FL>(lc, ly, lq, lr, eta1, c0, sigma);
                if (want0) ll0 += l0;
                if (want1) ll1 += l1;
            }
        }
    }
    GMB_TICK(6);                                       // residuals
#pragma unroll
    for (int t = 0; t < NT; t++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            // res is held as C fragment [row = lane/4][chain = 2(lane%4) + {0,1}]; the transposed product needs it as
            // B fragment [k = row = 4h + lane%4][n = chain = lane/4]
            const int src = 4 * (4 * h + fk) + (fr >> 1);
            const double t0 = __shfl_sync(0xffffffffu, res[t][0], src);
            const double t1 = __shfl_sync(0xffffffffu, res[t][1], src);
            const double b = (fr & 1) ? t1 : t0;
            const double* zt = sZL + (r0[t] + 4 * h + fk) * LD + fr;
#pragma unroll
            for (int i = 0; i < QT8; i++) dmma884(gacc[i][0], gacc[i][1], zt[8 * i], b);
        }
    GMB_TICK(7);                                       // gradient products
}

// Shared-memory carve-up (doubles), shared by the kernel and the host-side size computation.
struct FusedLayout {
    int zl, tab, rowv, slot, ll, xch, xll, mbar, total;   // offsets in doubles
};
__host__ __device__ inline FusedLayout fused_layout(int n8, int ld, int cs, int fl, int slots8) {
    const int qp8 = ((ld / 4 + 1) / 2) * 8;
    const int npar = cs > 1 ? 2 : 1;
    FusedLayout L;
    int o = 0;
    L.zl = o;    o += n8 * ld + 8;                        // Z L rows of this CTA (+8 spill-over doubles for the padded q-tile reads)
    L.tab = o;   o += 64;                                 // 2^(j/64)
    L.rowv = o;  o += (cs > 1) ? 3 * n8 : 0;               // xb, cnt, ys of this CTA's rows (cluster variant)
    L.slot = o;  o += ((cs > 1 || slots8) ? NWARP : NWARP / 2) * qp8 * 9;   // per-warp partial gradients
    L.ll = o;    o += NWARP * CB;                         // per-warp partial log-likelihoods
    L.xch = o;   o += npar * cs * CB * ld;                // [parity][rank][chain][q] partial gradients of every CTA of the cluster
    L.xll = o;   o += npar * cs * CB;                     // [parity][rank][chain] partial log-likelihoods
    L.mbar = o;  o += 2;                                  // one mbarrier per parity (cluster variant)
    L.total = o;
    return L;
}

template <int FL, int KS, int CS>
__global__ void __launch_bounds__(THREADS, 1) hmc_fused_kernel(const FusedParams p) {
    constexpr int LD = 4 * KS;                       // row stride of the Z L tile; KS = 1 (mod 4) makes it 4 (mod 16)
    constexpr int QT8 = (KS + 1) / 2;                // 8-row tiles of the gradient
    constexpr int QP8 = QT8 * 8;
    constexpr int QT32 = (LD + 31) / 32;             // gradient elements per lane in the [chain = warp][q = lane + 32 k] layout
    constexpr bool CL = CS > 1;
    extern __shared__ __align__(16) double sm[];
    const int Q = p.Q;
    const FusedLayout lay = fused_layout(p.n8, LD, CS, FL, p.slots8);
    const bool slots8 = CL || p.slots8;              // one slot per warp (one barrier) when shared memory allows, else 4 slots in two phases
    double* sZL = sm + lay.zl;
    double* sTab = sm + lay.tab;
    double* sXB = sm + lay.rowv;                      // cluster variant only
    double* sCN = sXB + p.n8;
    double* sYS = sCN + p.n8;
    double* sSlot = sm + lay.slot;                    // [slots][QP8][9]
    double* sLL = sm + lay.ll;                        // [NWARP][CB]
    double* sXch = sm + lay.xch;                      // [parities][CS][CB][LD]
    double* sXll = sm + lay.xll;                      // [parities][CS][CB]
    unsigned long long* sBar = reinterpret_cast<unsigned long long*>(sm + lay.mbar);   // [2]
    constexpr uint32_t TX_BYTES = (uint32_t)CS * NWARP * (LD + 1) * 8;   // bytes every CTA receives per gradient exchange

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    const int crank = CL ? (int)(blockIdx.x % CS) : 0;            // rank of this CTA inside its cluster (cluster dims = (CS,1,1))
    const int group = CL ? (int)(blockIdx.x / CS) : (int)blockIdx.x;
    const int chain = group * CB + fr;                // the chain this lane holds a fragment of (local index, < C if live)
    const bool live = chain < p.C;
    const bool owner = (fr == warp);                  // warp w's lanes with fr == w store chain w's state
    const bool writer = owner && live && crank == 0;  // every CTA of a cluster holds the same state; rank 0 stores the samples
    const uint32_t gchain = p.chain_offset + (uint32_t)chain;
    double* scr = p.scratch + ((size_t)blockIdx.x * CB + fr) * 4 * LD;   // this CTA's copy: [parity][v | grad][LD] of chain fr

    // rows of this CTA: global rows [row0, row0 + nloc)
    const int row0 = crank * p.tiles_per_cta * 8;
    const int nloc = max(0, min(p.n - row0, p.tiles_per_cta * 8));

    // ---- stage this CTA's rows of Z L (global: n x Q column-major) into shared memory, row-major, zero padded ----
    for (int idx = tid; idx < p.n8 * LD + 8; idx += THREADS) sZL[idx] = 0.0;
    if (tid < 64) sTab[tid] = GMB_EXP2_TAB[tid];
    if (CL && tid == 0) {
        mbar_init(smem_u32(&sBar[0]), 1); mbar_init(smem_u32(&sBar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (int idx = tid; idx < nloc * Q; idx += THREADS) {
        const int row = idx % nloc, q = idx / nloc;
        sZL[(size_t)row * LD + q] = p.ZL[row0 + row + (size_t)q * p.ldn];
    }
    if (CL) {
        for (int idx = tid; idx < p.n8; idx += THREADS) {
            const bool ok = idx < nloc;
            sXB[idx] = ok ? p.xb[row0 + idx] : 0.0;
            sCN[idx] = ok ? p.cnt[row0 + idx] : 0.0;
            sYS[idx] = ok ? p.ys[row0 + idx] : 0.0;
        }
    }
    const double* rxb = CL ? sXB : p.xb;
    const double* rcn = CL ? sCN : p.cnt;
    const double* rys = CL ? sYS : p.ys;
    const LLRows lw = {p.lcnt + row0, p.lys + row0, p.lsq + row0, p.lrc + row0};
    // shared::cluster addresses of the exchange buffers and mbarriers of every CTA of the cluster (distributed shared memory)
    uint32_t xch_of[CS], xll_of[CS], bar_of[CS];
    if (CL) {
#pragma unroll
        for (int rk = 0; rk < CS; rk++) {
            xch_of[rk] = mapa_u32(smem_u32(sXch), rk);
            xll_of[rk] = mapa_u32(smem_u32(sXll), rk);
            bar_of[rk] = mapa_u32(smem_u32(&sBar[0]), rk);
        }
    }
    uint32_t phase = 0;                               // bit p: phase parity to wait for on mbarrier p
    int par = 0;                                      // parity of the exchange buffer in use (cluster variant)

    const double sigma = p.var_par;
    const double sc = (FL == 7) ? 1.0 / (sigma * sigma) : 1.0;
    const double c0 = (FL == 7) ? (-1.0 * log(sigma) - 0.5 * log(2 * GMB_PI_FAMILY)) : 0.0;
    const double pc = -1.0 * log(1.0) - 0.5 * log(2 * GMB_PI_FAMILY);   // log_likelihood(v, 0, 1, 7), mcmlmodel.h:149
    const int nfull = nloc / 8;                       // tiles without padding rows
    const bool has_tail = (nloc % 8) != 0;

    // ---- chain state: fragments of v' and r, per-chain scalars (all replicated) ----
    double vp[KS], r[KS], g[KS];
    double eps = 0.001, ebar = 1.0, H = 0.0, llcur = 0.0, llnew = 0.0;   // initialise_u, mhmcmc.h:47-59
    int accept = 0, steps = 1, cur = 0;
    double totsteps = 0.0, lastprob = 0.0;
#pragma unroll
    for (int j = 0; j < KS; j++) {
        const int q = 4 * j + fk;
        double z0, z1;
        dev_rng_normal2(p.seed, (uint32_t)(q >> 1), 0u, gchain, 0u, z0, z1);
        vp[j] = (q < Q) ? ((q & 1) ? z1 : z0) : 0.0;
        r[j] = 0.0;
    }
    if (CL) cg::this_cluster().sync();                // every CTA's shared memory is initialised before any remote store
    else __syncthreads();
    FusedTim tim;
#ifdef GMB_FUSED_TIMING
    for (int i = 0; i < 12; i++) tim.acc[i] = 0;
    tim.last = clock64(); tim.rec = (tid == 0);
#endif
    // One evaluation of the gradient at the v' held in vp[], for the 8 chains of the group.
    // s = leapfrog step index; st0 / st1 = step counts of chains 2 fk and 2 fk + 1 (the accumulator columns of this lane):
    // the log-likelihood of a chain is accumulated on its last step (with_ll: some chain of the group is on its last step).
    // On return g[] holds grad(v') in fragment layout and llnew the family log-likelihood of chain fr (if this was its
    // last step).
    auto grad_eval = [&](int s, bool with_ll, int st0, int st1) {
        double gacc[QT8][2];
#pragma unroll
        for (int i = 0; i < QT8; i++) gacc[i][0] = gacc[i][1] = 0.0;
        const bool want0 = (s == st0 - 1), want1 = (s == st1 - 1);
        double ll0 = 0.0, ll1 = 0.0;
        int tile = warp;
        if (!with_ll) {
            for (; tile + 3 * NWARP < nfull; tile += 4 * NWARP) {
                const int r0[4] = {tile * 8, (tile + NWARP) * 8, (tile + 2 * NWARP) * 8, (tile + 3 * NWARP) * 8};
                fused_tiles<FL, KS, 4, false, false, CL>(sZL, sTab, vp, r0, nloc, rxb, rcn, rys, lw, c0, sigma, want0, want1, fr, fk, gacc, ll0, ll1, tim);
            }
            for (; tile + NWARP < nfull; tile += 2 * NWARP) {
                const int r0[2] = {tile * 8, (tile + NWARP) * 8};
                fused_tiles<FL, KS, 2, false, false, CL>(sZL, sTab, vp, r0, nloc, rxb, rcn, rys, lw, c0, sigma, want0, want1, fr, fk, gacc, ll0, ll1, tim);
            }
            for (; tile < nfull; tile += NWARP) {
                const int r0[1] = {tile * 8};
                fused_tiles<FL, KS, 1, false, false, CL>(sZL, sTab, vp, r0, nloc, rxb, rcn, rys, lw, c0, sigma, want0, want1, fr, fk, gacc, ll0, ll1, tim);
            }
            if (has_tail && tile == nfull) {
                const int r0[1] = {tile * 8};
                fused_tiles<FL, KS, 1, true, false, CL>(sZL, sTab, vp, r0, nloc, rxb, rcn, rys, lw, c0, sigma, want0, want1, fr, fk, gacc, ll0, ll1, tim);
            }
        } else {
            for (; tile < nfull; tile += NWARP) {
                const int r0[1] = {tile * 8};
                fused_tiles<FL, KS, 1, false, true, CL>(sZL, sTab, vp, r0, nloc, rxb, rcn, rys, lw, c0, sigma, want0, want1, fr, fk, gacc, ll0, ll1, tim);
            }
            if (has_tail && tile == nfull) {
                const int r0[1] = {tile * 8};
                fused_tiles<FL, KS, 1, true, true, CL>(sZL, sTab, vp, r0, nloc, rxb, rcn, rys, lw, c0, sigma, want0, want1, fr, fk, gacc, ll0, ll1, tim);
            }
            ll0 += __shfl_xor_sync(0xffffffffu, ll0, 4);  ll1 += __shfl_xor_sync(0xffffffffu, ll1, 4);
            ll0 += __shfl_xor_sync(0xffffffffu, ll0, 8);  ll1 += __shfl_xor_sync(0xffffffffu, ll1, 8);
            ll0 += __shfl_xor_sync(0xffffffffu, ll0, 16); ll1 += __shfl_xor_sync(0xffffffffu, ll1, 16);
            if (fr == 0) { sLL[warp * CB + 2 * fk] = ll0; sLL[warp * CB + 2 * fk + 1] = ll1; }
        }
        // ---- deterministic cross-warp sum of the partial gradients (C-fragment layout -> [q][chain] slots) ----
        double gsum[QT32];                            // this CTA's partial for chain `warp`, q = lane + 32 k
#pragma unroll
        for (int k = 0; k < QT32; k++) gsum[k] = 0.0;
        if (!slots8) {
            // warps 4-7 -> 4 slots, warps 0-3 add, then warp c sums the 4 slots of chain c
            double* slot = sSlot + (warp & 3) * QP8 * 9;
            if (warp >= 4) {
#pragma unroll
                for (int i = 0; i < QT8; i++) {
                    slot[(8 * i + fr) * 9 + 2 * fk] = gacc[i][0];
                    slot[(8 * i + fr) * 9 + 2 * fk + 1] = gacc[i][1];
                }
            }
            __syncthreads();
            if (warp < 4) {
#pragma unroll
                for (int i = 0; i < QT8; i++) {
                    slot[(8 * i + fr) * 9 + 2 * fk] += gacc[i][0];
                    slot[(8 * i + fr) * 9 + 2 * fk + 1] += gacc[i][1];
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < QT32; k++) {
                const int q = lane + 32 * k;
                if (q < LD) gsum[k] = ((sSlot[q * 9 + warp] + sSlot[(QP8 + q) * 9 + warp]) + sSlot[(2 * QP8 + q) * 9 + warp]) + sSlot[(3 * QP8 + q) * 9 + warp];
            }
        } else {
            // every warp stores its partial; one barrier; warp c sums the 8 partials of chain c in warp order
            double* slot = sSlot + warp * QP8 * 9;
#pragma unroll
            for (int i = 0; i < QT8; i++) {
                slot[(8 * i + fr) * 9 + 2 * fk] = gacc[i][0];
                slot[(8 * i + fr) * 9 + 2 * fk + 1] = gacc[i][1];
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < QT32; k++) {
                const int q = lane + 32 * k;
                if (q < LD) {
                    double t = sSlot[q * 9 + warp];
#pragma unroll
                    for (int w = 1; w < NWARP; w++) t += sSlot[(w * QP8 + q) * 9 + warp];
                    gsum[k] = t;
                }
            }
        }
        GMB_TICK(2);                                   // slot stores, barrier(s), local sums
        // ---- hand this CTA's partial (chain `warp`) to every CTA of the cluster and read the sum back in fragment layout.
        //      Cluster variant: one st.async per element and destination; the store signals the destination's mbarrier
        //      (complete_tx), so a step needs neither a cluster-wide barrier nor a memory fence ----
        if (CL && tid == 0) mbar_arrive_expect_tx(smem_u32(&sBar[par]), TX_BYTES);
#pragma unroll
        for (int k = 0; k < QT32; k++) {
            const int q = lane + 32 * k;
            if (q < LD) {
                const int off = ((par * CS + crank) * CB + warp) * LD + q;
                if (CL) {
#pragma unroll
                    for (int rk = 0; rk < CS; rk++) st_async_f64(xch_of[rk] + 8u * off, gsum[k], bar_of[rk] + 8u * par);
                } else {
                    sXch[off] = gsum[k];
                }
            }
        }
        if (lane == 0 && (CL || with_ll)) {
            double l = 0.0;
            if (with_ll) {
#pragma unroll
                for (int w = 0; w < NWARP; w++) l += sLL[w * CB + warp];
            }
            const int off = (par * CS + crank) * CB + warp;
            if (CL) {
#pragma unroll
                for (int rk = 0; rk < CS; rk++) st_async_f64(xll_of[rk] + 8u * off, l, bar_of[rk] + 8u * par);
            } else {
                sXll[off] = l;
            }
        }
        GMB_TICK(3);                                   // exchange stores
        if (CL) { mbar_wait(smem_u32(&sBar[par]), (phase >> par) & 1u); phase ^= 1u << par; }
        else __syncthreads();
        GMB_TICK(4);                                   // exchange completion
#pragma unroll
        for (int j = 0; j < KS; j++) {
            double gs = sXch[((par * CS + 0) * CB + fr) * LD + 4 * j + fk];
#pragma unroll
            for (int rk = 1; rk < CS; rk++) gs += sXch[((par * CS + rk) * CB + fr) * LD + 4 * j + fk];
            g[j] = -1.0 * vp[j] + sc * gs;                                          // mcmlmodel.h:163 + :173/:191/:235
        }
        if (with_ll && s == steps - 1) {
            double l = sXll[(par * CS + 0) * CB + fr];
#pragma unroll
            for (int rk = 1; rk < CS; rk++) l += sXll[(par * CS + rk) * CB + fr];
            llnew = l;
        }
        if (CL) par ^= 1;
        GMB_TICK(5);                                   // fragment read
    };

    // gradient and log-likelihood at the initial state (carried over between proposals instead of recomputed, mhmcmc.h:64,82)
    grad_eval(0, true, 1, 1);
    llcur = llnew;
    if (owner) {
#pragma unroll
        for (int j = 0; j < KS; j++) { scr[4 * j + fk] = vp[j]; scr[LD + 4 * j + fk] = g[j]; }
    }

    const int total = p.warmup + p.nsamp;
    const int cols = p.nsamp + 1;
    if (p.warmup == 0 && writer) {                                                 // samples.col(0) = u_, mhmcmc.h:142
#pragma unroll
        for (int j = 0; j < KS; j++) { const int q = 4 * j + fk; if (q < Q) p.dV_out[((size_t)chain * cols) * p.ldq + q] = vp[j]; }
    }

    for (int t = 0; t < total; t++) {
        GMB_TICK(9);                                   // Metropolis test, adaptation, sample store (previous proposal)
        __syncthreads();                               // the owners' scratch stores of the previous proposal are visible
        // ---- new_proposal, mhmcmc.h:61-75 ----
        double k0 = 0.0, pv = 0.0;
        {
            const double* V = scr + cur * 2 * LD;
            const double* G = V + LD;
#pragma unroll
            for (int j = 0; j < KS; j++) {
                const int q = 4 * j + fk;
                double z0, z1;
                dev_rng_normal2(p.seed, (uint32_t)(q >> 1), (uint32_t)t, gchain, 2u, z0, z1);      // :62-63
                const double z = (q & 1) ? z1 : z0;
                if (q < Q) {
                    const double vq = V[q], gq = G[q];
                    k0 += z * z;
                    pv += pc - 0.5 * vq * vq;
                    r[j] = z + (eps / 2) * gq;                                                      // :74 (first step)
                    vp[j] = vq + eps * r[j];                                                        // :67, :75
                } else { r[j] = 0.0; vp[j] = 0.0; }
            }
        }
        k0 = 0.5 * quad_sum(k0);                                                                    // :66
        pv = quad_sum(pv);
        {
            const double sd = round(p.lambda / eps);                                                // :69
            steps = sd >= (double)p.max_steps ? p.max_steps : (sd < 1.0 ? 1 : (int)sd);             // :69-70
            if (!(sd == sd)) steps = p.max_steps;
            if (!live) steps = 1;
            totsteps += steps;
        }
        int smax = steps;
        smax = max(smax, __shfl_xor_sync(0xffffffffu, smax, 4));
        smax = max(smax, __shfl_xor_sync(0xffffffffu, smax, 8));
        smax = max(smax, __shfl_xor_sync(0xffffffffu, smax, 16));
        const int st0 = __shfl_sync(0xffffffffu, steps, 8 * fk), st1 = __shfl_sync(0xffffffffu, steps, 8 * fk + 4);
        GMB_TICK(8);                                   // proposal: momentum draw, first half step
        // ---- leapfrog integrator, :73-78 ----
        for (int s = 0; s < smax; s++) {
            const bool any_last = __any_sync(0xffffffffu, s == steps - 1);   // some chain needs its log-likelihood on this step
            GMB_TICK(0);                               // leapfrog update (and, once per proposal, everything outside the step loop)
            grad_eval(s, any_last, st0, st1);
            if (s < steps) {
                const bool more = s < steps - 1;
#pragma unroll
                for (int j = 0; j < KS; j++) {
                    double rr = r[j] + (eps / 2) * g[j];                                            // :77
                    if (more) {
                        rr = rr + (eps / 2) * g[j];                                                 // :74 of the next step
                        vp[j] = vp[j] + eps * rr;                                                   // :75
                    }
                    r[j] = rr;
                }
                if (!more && owner) {                  // candidate state and its gradient, adopted if the proposal is accepted
                    double* Vc = scr + (cur ^ 1) * 2 * LD;
#pragma unroll
                    for (int j = 0; j < KS; j++) { Vc[4 * j + fk] = vp[j]; Vc[LD + 4 * j + fk] = g[j]; }
                }
            }
        }
        // ---- Metropolis test and adaptation, :80-117 ----
        double k1 = 0.0, pvp = 0.0;
#pragma unroll
        for (int j = 0; j < KS; j++) {
            if (4 * j + fk < Q) { k1 += r[j] * r[j]; pvp += pc - 0.5 * vp[j] * vp[j]; }
        }
        k1 = 0.5 * quad_sum(k1); pvp = quad_sum(pvp);
        const double l1 = llcur + pv, l2 = llnew + pvp;                                            // :82-83
        const double prob = fmin(1.0, exp(-l1 + k0 + l2 - k1));                                    // :84
        double u1, u2;
        dev_rng_uniform2(p.seed, 0u, (uint32_t)t, gchain, 3u, u1, u2);                             // :85
        const bool acc = u1 < prob;                                                                // :86
        lastprob = prob;
        if (acc) { accept++; llcur = llnew; cur ^= 1; }                                            // :102-105
        if (t < p.warmup && t < p.adapt) {                                                         // :107-114, :131-136
            const int iter = t + 1;
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
        const int col = t - p.warmup + 1;                                                          // :142 (col 0), :147
        if (col >= 0 && writer) {
            const double* V = scr + cur * 2 * LD;      // own stores (same thread): visible without a barrier
#pragma unroll
            for (int j = 0; j < KS; j++) { const int q = 4 * j + fk; if (q < Q) p.dV_out[((size_t)chain * cols + col) * p.ldq + q] = V[q]; }
        }
    }
    if (writer && fk == 0) {
        const int C = p.C;
        p.cs_out[FS_EPS * C + chain] = eps; p.cs_out[FS_EBAR * C + chain] = ebar; p.cs_out[FS_H * C + chain] = H;
        p.cs_out[FS_LLCUR * C + chain] = llcur; p.cs_out[FS_K0 * C + chain] = 0.0; p.cs_out[FS_ACCEPT * C + chain] = (double)accept;
        p.cs_out[FS_TOTSTEPS * C + chain] = totsteps; p.cs_out[FS_LASTPROB * C + chain] = lastprob;
    }
#ifdef GMB_FUSED_TIMING
    if (tid == 0 && p.timing) for (int i = 0; i < 12; i++) p.timing[blockIdx.x * 12 + i] = tim.acc[i];
#endif
    if (CL) cg::this_cluster().sync();                // no CTA may exit while a peer can still store into its shared memory
}

template <int FL, int KS, int CS>
int launch_fused(gmb_ctx* ctx, const FusedParams& p, size_t smem) {
    auto kern = hmc_fused_kernel<FL, KS, CS>;
    GMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int groups = (p.C + CB - 1) / CB;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(groups * CS); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = CS > 1 ? 1 : 0;
    GMB_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    ctx->launches++;
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}

template <int FL, int KS>
int launch_fused_cs(gmb_ctx* ctx, const FusedParams& p, size_t smem, int cs) {
    switch (cs) {
    case 1: return launch_fused<FL, KS, 1>(ctx, p, smem);
    case 2: return launch_fused<FL, KS, 2>(ctx, p, smem);
    case 4: return launch_fused<FL, KS, 4>(ctx, p, smem);
    }
    return gmb_set_error(GMB_EINVAL, "bad cluster size %d", cs);
}

template <int FL>
int launch_fused_ks(gmb_ctx* ctx, const FusedParams& p, size_t smem, int cs) {
    switch (p.ks) {      // ld / 4; ld = 4 (mod 16)
    case 13: return launch_fused_cs<FL, 13>(ctx, p, smem, cs);
#ifndef GMB_FUSED_DEV    /* development builds (EXTRA=-DGMB_FUSED_DEV) instantiate Q <= 52 only: seconds instead of minutes */
    case 1: return launch_fused_cs<FL, 1>(ctx, p, smem, cs);
    case 5: return launch_fused_cs<FL, 5>(ctx, p, smem, cs);
    case 9: return launch_fused_cs<FL, 9>(ctx, p, smem, cs);
    case 17: return launch_fused_cs<FL, 17>(ctx, p, smem, cs);
    case 21: return launch_fused_cs<FL, 21>(ctx, p, smem, cs);
    case 25: return launch_fused_cs<FL, 25>(ctx, p, smem, cs);
    case 29: return launch_fused_cs<FL, 29>(ctx, p, smem, cs);
    case 33: return launch_fused_cs<FL, 33>(ctx, p, smem, cs);
#endif
    }
    return gmb_set_error(GMB_EINVAL, "no on-chip sampler instantiation for Q = %d", p.Q);
}

}  // namespace
