// hmc_fused.cu — host side of the on-chip sampler (kernel: hmc_fused.cuh; instantiated per family in hmc_fused_fl{1,3,7}.cu
// so that the three families compile in parallel).
#include "hmc_fused.cuh"

int gmb_fused_launch_fl1(gmb_ctx* ctx, const FusedParams& p, size_t smem, int cs);
int gmb_fused_launch_fl3(gmb_ctx* ctx, const FusedParams& p, size_t smem, int cs);
int gmb_fused_launch_fl7(gmb_ctx* ctx, const FusedParams& p, size_t smem, int cs);

namespace {

int fused_ld(int Q) {
    int l = (Q + 3) / 4 * 4;
    while (l % 16 != 4) l += 4;
    return l;
}

constexpr size_t SMEM_LIMIT = (size_t)227 * 1024;

size_t fused_smem_bytes(int n8, int ld, int cs, int fl, int slots8) {
    return sizeof(double) * (size_t)fused_layout(n8, ld, cs, fl, slots8).total + 16;
}

int tiles_per_cta(int n, int cs) {
    const int tiles = (n + 7) / 8;
    return (tiles + cs - 1) / cs;
}

// rows of the sampler's view of a model (after row aggregation, aggregate.cu); the model's own n before the view is built
int view_rows(const gmb_model* mdl) { return mdl->agg.built ? mdl->agg.ng : mdl->n; }

bool fused_fits(int n, int Q, int cs, int fl) {
    const int ld = fused_ld(Q);
    if (ld / 4 > 33) return false;
    return fused_smem_bytes(tiles_per_cta(n, cs) * 8, ld, cs, fl, 0) <= SMEM_LIMIT;
}

// forced cluster size (0 = choose): gmb_hmc_set_cluster_size or the environment variable GMB_FUSED_CS
int g_forced_cs = -1;
int forced_cs() {
    if (g_forced_cs < 0) { const char* e = getenv("GMB_FUSED_CS"); g_forced_cs = e ? atoi(e) : 0; }
    return g_forced_cs;
}

// Cluster size for a run of C chains: the one with the shortest estimated leapfrog step among those that fit.
// Cost model per step (in units of one 8-row tile per warp): waves * (tiles per warp + fixed overhead of the reduction,
// state update and barriers); a cluster launch keeps 148 (CS = 2) or 132 (CS = 4) SMs busy (B300_MICROARCH.md).
int choose_cs(const gmb_model* mdl, int C) {
    const int f = forced_cs();
    const int n = view_rows(mdl);
    if (f == 1 || f == 2 || f == 4) return fused_fits(n, mdl->Q, f, mdl->flink) ? f : 0;
    const int groups = (C + CB - 1) / CB, sms = mdl->ctx->sms;
    int best = 0; double best_cost = 0.0;
    for (int cs = 1; cs <= 4; cs *= 2) {
        if (!fused_fits(n, mdl->Q, cs, mdl->flink)) continue;
        const int usable = cs == 4 ? (sms / 4 * 4 * 132) / 148 : sms / cs * cs;
        const int waves = (groups * cs + usable - 1) / (usable > 0 ? usable : 1);
        const int tpw = (tiles_per_cta(n, cs) + NWARP - 1) / NWARP;
        const double cost = waves * (tpw + 2.0 + (cs > 1 ? 0.75 : 0.0));
        if (!best || cost < best_cost) { best = cs; best_cost = cost; }
    }
    return best;
}

}  // namespace

// 0 = choose per run, 1 = one CTA per group of 8 chains, 2 / 4 = split the observations over a cluster of 2 / 4 CTAs
extern "C" int gmb_hmc_set_cluster_size(int cs) {
    if (cs != 0 && cs != 1 && cs != 2 && cs != 4) return gmb_set_error(GMB_EINVAL, "cluster size must be 0 (auto), 1, 2 or 4");
    g_forced_cs = cs;
    return GMB_OK;
}

// true when the on-chip variant can run this model with C chains (its share of Z L + work buffers fit the 227 KB of one SM)
bool gmb_hmc_fused_applicable(const gmb_model* mdl, int C) {
    return choose_cs(mdl, C) != 0;
}

// doubles of global scratch the kernel needs behind the FS_COUNT x C chain statistics (d_cs + gmb_hmc_fused_cs_doubles(C))
size_t gmb_hmc_fused_cs_doubles(int C) { return round_up_sz((size_t)FS_COUNT * C + 16, 16); }
size_t gmb_hmc_fused_scratch_doubles(const gmb_model* mdl, int C) {
    const int groups = (C + CB - 1) / CB;
    return (size_t)groups * 4 /* max cluster size */ * CB * 4 * fused_ld(mdl->Q);
}

// Same contract as the two-GEMM hmc_run of hmc.cu: dV_out is ldq x (C * (nsamp + 1)) chain-major, d_cs is FS_COUNT x C.
int gmb_hmc_run_fused(gmb_model* mdl, double var_par, int warmup, int nsamp, double lambda, int max_steps, double target_accept,
                      int adapt, int C, uint32_t chain_offset, uint64_t seed, double* dV_out, double* d_cs) {
    gmb_ctx* ctx = mdl->ctx;
    const int cs = choose_cs(mdl, C);
    if (!cs) return gmb_set_error(GMB_EINVAL, "Z L (%d x %d) does not fit the on-chip sampler", view_rows(mdl), mdl->Q);
    const gmb_agg& a = mdl->agg;
    if (!a.built) return gmb_set_error(GMB_ESTATE, "on-chip sampler: the row view of the model has not been built");
    FusedParams p;
    p.n = a.ng; p.Q = mdl->Q; p.ld = fused_ld(mdl->Q); p.ks = p.ld / 4; p.qt8 = (mdl->Q + 7) / 8;
    p.ldn = a.ldn; p.ldq = mdl->ldq;
    p.tiles_per_cta = tiles_per_cta(a.ng, cs); p.n8 = p.tiles_per_cta * 8;
    p.ZL = a.active ? a.dZL : mdl->dZL; p.xb = a.active ? a.dxb : mdl->dxb;
    p.cnt = a.dcnt; p.ys = a.dys; p.lcnt = a.dlcnt; p.lys = a.dlys; p.lsq = a.dlsq; p.lrc = a.dlrc;
    p.var_par = var_par; p.lambda = lambda; p.target_accept = target_accept;
    p.warmup = warmup; p.nsamp = nsamp; p.max_steps = max_steps; p.adapt = adapt; p.C = C;
    p.chain_offset = chain_offset; p.seed = seed; p.dV_out = dV_out; p.cs_out = d_cs;
    p.scratch = d_cs + gmb_hmc_fused_cs_doubles(C);
    p.timing = nullptr;
#ifdef GMB_FUSED_TIMING
    {
        static long long* d_tim = nullptr;
        const int ctas = (C + CB - 1) / CB * cs;
        if (!d_tim) GMB_CUDA(cudaMalloc(&d_tim, sizeof(long long) * 12 * 4096));
        GMB_CUDA(cudaMemsetAsync(d_tim, 0, sizeof(long long) * 12 * 4096, ctx->stream));
        p.timing = d_tim;
        struct Dump { long long* d; int ctas; cudaStream_t st; };
        static Dump last; last = {d_tim, ctas, ctx->stream};
        // the caller synchronises the stream; print the previous launch's counters on the next call or at exit
        static bool reg = false;
        if (!reg) { reg = true; atexit([] {
            std::vector<long long> h(12 * 4096); cudaMemcpy(h.data(), last.d, sizeof(long long) * 12 * last.ctas, cudaMemcpyDeviceToHost);
            long long tot[12] = {0}; for (int b = 0; b < last.ctas; b++) for (int i = 0; i < 12; i++) tot[i] += h[b * 12 + i];
            long long all = 0; for (int i = 0; i < 12; i++) all += tot[i];
            const char* nm[12] = {"leapfrog update", "tiles: eta", "slots+local sum", "exchange stores", "barrier", "fragment read", "tiles: residual", "tiles: gradient",
                                  "proposal setup", "metropolis+store", "-", "-"};
            fprintf(stderr, "[GMB_FUSED_TIMING] last launch, %d CTAs, mean cycles per CTA:\n", last.ctas);
            for (int i = 0; i < 10; i++) fprintf(stderr, "  %-18s %12.0f  %5.1f%%\n", nm[i], (double)tot[i] / last.ctas, 100.0 * tot[i] / (all ? all : 1));
        }); }
    }
#endif
    p.slots8 = (cs == 1 && fused_smem_bytes(p.n8, p.ld, cs, mdl->flink, 1) <= SMEM_LIMIT) ? 1 : 0;
    const size_t smem = fused_smem_bytes(p.n8, p.ld, cs, mdl->flink, p.slots8);
    switch (mdl->flink) {
    case 1: return gmb_fused_launch_fl1(ctx, p, smem, cs);
    case 3: return gmb_fused_launch_fl3(ctx, p, smem, cs);
    case 7: return gmb_fused_launch_fl7(ctx, p, smem, cs);
    }
    return gmb_set_error(GMB_EFAMILY, "family/link code %d has no device kernel", mdl->flink);
}
