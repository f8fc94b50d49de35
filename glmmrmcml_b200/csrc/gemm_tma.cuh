// gemm_tma.cuh — FP64 GEMM mainloop with TMA operand staging (cp.async.bulk.tensor + mbarrier expect-tx), a dedicated producer warp and
// 8 DMMA consumer warps; same epilogue concept as gemm_f64.cuh (which stays for small problems and the narrow in-place panel products).
//
// tcgen05 has no f64 kind: the FP64 tensor instruction on sm_100a is mma.sync.m8n8k4 (DMMA), fed from shared memory.  What changes here against
// gemm_f64.cuh is how the operands get there:
//   * CTA tile 128 x 128, k tile BK = 32, 3 stages of 64 KB: one elected thread (lane 0 of warp 0, two k tiles ahead of the arithmetic) arms the
//     stage's `full` mbarrier with the byte count and issues the TMA boxes; the 8 warps wait on `full`, run 8 k4 steps (32 DMMA each per warp) and
//     release the stage through its `empty` mbarrier (one arrive per warp).  There is no block-wide barrier and no per-thread address arithmetic in the main loop
//     (the cp.async version spends 8 LDGSTS + their predicates and address computations per thread and k tile and a __syncthreads per k tile).
//   * Out-of-range rows / columns / k are zero-filled by the TMA unit: no edge predicates.
//   * Bank conflicts: a TMA box cannot be padded, so the boxes are 128 bytes wide (16 doubles) with the 128-byte swizzle, and the m8n8k4
//     fragments pick their rows so that every half-warp access touches 16 distinct 8-byte banks:
//       operand contiguous in m / n:  box [16 (m) x BK (k)], element (mi, k) at 16-byte chunk (mi >> 1) ^ (k & 7) of row k; the 8 rows of fragment
//                                     tile h of a 16-row group are {0,1,8,9,2,3,10,11} + 4h (lane>>2 -> that order);
//       operand contiguous in k:      box [16 (k) x 128 (m)], element (m, kk) at chunk (kk >> 1) ^ (m & 7) of row m; fragment row lane>>2 -> 2 (fr & 3) + (fr >> 2).
//     The epilogue maps the accumulator fragments back through the same permutations.
#pragma once
#include "gemm_f64.cuh"
#include <cuda.h>

namespace gmbtma {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3;
constexpr int CONSUMER_WARPS = 8, THREADS = CONSUMER_WARPS * 32;   // (a ninth, producer-only warp would round the CTA up to 12 warps of registers: 168 per thread, spills)
constexpr int TILE_DOUBLES = BM * BK;                      // one operand, one stage (32 KB)
constexpr size_t SMEM_BYTES = (size_t)STAGES * 2 * TILE_DOUBLES * sizeof(double) + 1024 /* alignment slack */ + 64 /* barriers */;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// row of a 16-row group read by fragment row fr (0..7) of fragment tile h (0, 1): operand contiguous in m / n
__device__ __forceinline__ int rho_mn(int h, int fr) { return (fr & 1) + 8 * ((fr >> 1) & 1) + 2 * (fr >> 2) + 4 * h; }
// row of an 8-row tile read by fragment row fr: operand contiguous in k
__device__ __forceinline__ int rho_k(int fr) { return 2 * (fr & 3) + (fr >> 2); }

// offset (doubles) inside one operand stage of element (mn, k), mn in [0, 128), k in [0, BK)
template <bool KCONT>
__device__ __forceinline__ int tile_off(int mn, int k) {
    if (KCONT) {
        const int kh = k >> 4, kk = k & 15;
        return kh * (128 * 16) + mn * 16 + ((((kk >> 1) ^ (mn & 7))) << 1) + (kk & 1);
    } else {
        const int g = mn >> 4, mi = mn & 15;
        return g * (BK * 16) + k * 16 + (((mi >> 1) ^ (k & 7)) << 1) + (mi & 1);
    }
}
// Per-lane fragment addressing with everything that depends on the lane folded into one offset and four swizzle terms (the straightforward
// tile_off(row, k) per load kept ~40 index registers live and spilled): offset of (fragment tile t, k step ks) = lane_off + sw[.] + constant(t, ks).
template <bool KCONT>
struct FragAddr {
    int lane_off, sw[4];
    __device__ __forceinline__ void init(int row0, int fr, int fk) {           // row0: first row of the warp's range inside the CTA tile
        if (KCONT) {
            const int rho = rho_k(fr);
            lane_off = (row0 + rho) * 16 + (fk & 1);
#pragma unroll
            for (int q = 0; q < 4; q++) sw[q] = (((2 * q + (fk >> 1)) ^ rho) << 1);       // k = 4 q + fk within a 16-wide half
        } else {
            lane_off = (row0 >> 4) * (BK * 16) + fk * 16;
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int e = 0; e < 2; e++) { const int mi = rho_mn(h, fr); sw[2 * h + e] = (((mi >> 1) ^ (4 * e + fk)) << 1) + (mi & 1); }
        }
    }
    // t: fragment tile of the warp (8 rows each), ks: k offset of the k4 step inside the stage (multiple of 4)
    __device__ __forceinline__ int off(int t, int ks) const {
        if (KCONT) return lane_off + (ks >> 4) * (128 * 16) + 8 * t * 16 + sw[(ks >> 2) & 3];
        return lane_off + (t >> 1) * (BK * 16) + ks * 16 + sw[2 * (t & 1) + ((ks >> 2) & 1)];
    }
};

// row (0..63 for A, 0..31 for B, relative to the warp's range) that fragment tile t, fragment row fr reads
template <bool KCONT>
__device__ __forceinline__ int frag_row(int t, int fr) { return KCONT ? 8 * t + rho_k(fr) : 16 * (t >> 1) + rho_mn(t & 1, fr); }

constexpr int RASTER_GROUP = 16;

template <bool A_KCONT, bool B_KCONT, class Epi>
__global__ void __launch_bounds__(THREADS, 1) dgemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                               int M, int N, int K, Epi epi, int tri_arg, int ptiles, unsigned int* queue) {
    constexpr int MT = 8, NT = 4;                          // warp tile 64 x 32: m8n8 fragment tiles per warp
    extern __shared__ unsigned char smraw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~(uintptr_t)1023);
    double* sA = reinterpret_cast<double*>(base);
    double* sB = sA + STAGES * TILE_DOUBLES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * TILE_DOUBLES);
    uint64_t* empty = full + STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // ptiles > 0: PERSISTENT — gridDim.x CTAs walk ptiles tiles, the operand ring and its barrier phases running on across tiles (it_base).
    //  * symmetric rank-k update (tri & 0xff == 3): static stride gridDim.x over the lower tile pairs.  The blocked Cholesky launches its trailing
    //    update on fewer CTAs than SMs so that the panel chain of the next outer block (high-priority side stream, tiny kernels) finds a free SM
    //    at once instead of waiting for a 128 x 128 x 512 tile (~70 us) to retire.
    //  * triangular operand (tri = 1 / 2) with queue != NULL: tiles are numbered HEAVIEST FIRST (longest k range) and handed out through an atomic
    //    counter — a CTA takes the next tile when it is done (longest-processing-time-first): 320 unequal tiles on 148 SMs (the two contractions per
    //    leapfrog step of config C5) finish in ~1.1 x the average load instead of 3 rounds of the longest tile.
    __shared__ int s_next;
    int it_base = 0;
    const int TM = (M + BM - 1) / BM, TN = (N + BN - 1) / BN;
    for (int tile = blockIdx.x, first_tile = 1;; first_tile = 0) {
    if (!first_tile) {
        if (ptiles == 0) break;
        if (queue) {
            __syncthreads();
            if (tid == 0) s_next = (int)gridDim.x + (int)atomicAdd(queue, 1u);
            __syncthreads();
            tile = s_next;
        } else tile += gridDim.x;
    }
    if (ptiles > 0 && tile >= ptiles) break;
    int tri = tri_arg;
    int bx = (tri == 1) ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x, by = (tri == 4) ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
    if (ptiles > 0 && (tri & 0xff) != 3) {                 // persistent, plain tile grid: heaviest tiles first
        if (tri == 1) { bx = TM - 1 - tile / TN; by = tile % TN; }
        else if (tri == 2) { bx = tile / TN; by = tile % TN; }
        else if (tri == 4) { by = TN - 1 - tile / TM; bx = tile % TM; }
        else { bx = tile % TM; by = tile / TM; }
    } else if (tri == 0 && gridDim.x > RASTER_GROUP) {
        // grouped rasterisation: the CTAs resident at one time (one per SM, launched in linear order) cover RASTER_GROUP tile rows x ~9 tile
        // columns instead of a full column of tiles x ~2, so a wave pulls (16 + 9) operand panels through L2 instead of (M/128 + 2): at
        // 8192 x 16384 x 4096 the DRAM reads fall from 19x the algorithmic bytes (ncu, profiles/r02_ncu_gemm_dgemm_tma_ke.txt) to ~3x
        const int pid = blockIdx.x + gridDim.x * blockIdx.y;
        const int per_group = RASTER_GROUP * gridDim.y;
        const int g = pid / per_group, first = g * RASTER_GROUP;
        const int gm = min((int)gridDim.x - first, RASTER_GROUP);
        const int r = pid - g * per_group;
        bx = first + r % gm; by = r / gm;
    }
    if ((tri & 0xff) == 3) {                               // symmetric rank-k update, lower tile pairs only (see gemm_f64.cuh); tile column tj
        const int T = (M + BM - 1) / BM;                   // holds the tile rows max(tj, rmin) .. T - 1
        int rem = tile, tj = (tri >> 8) & 0xfff;
        const int rmin = (tri >> 20) & 0xfff;
        while (rem >= T - max(tj, rmin)) { rem -= T - max(tj, rmin); tj++; }
        bx = max(tj, rmin) + rem; by = tj;
        tri = 0;
    }
    const int m0 = bx * BM, n0 = by * BN;
    const int KT = (K + BK - 1) / BK;
    const int kt0 = (tri == 2) ? min(m0 / BK, KT) : 0;
    // tri = 4: B is triangular, B(n, k) = 0 for k > n (the transposed inverse of a lower factor: a triangular solve as a product)
    const int kt1 = (tri == 1) ? min(KT, (m0 + BM + BK - 1) / BK) : (tri == 4) ? min(KT, (n0 + BN + BK - 1) / BK) : KT;

    if ((tri_arg & 0xff) != 3) {   // skip tiles whose columns are all inactive (chains that finished their trajectory)
        int act = 0;
        for (int c = tid; c < BN; c += THREADS) if (n0 + c < N && epi.column_active(n0 + c)) act = 1;
        if (!__syncthreads_or(act)) { if (ptiles == 0) return; else continue; }
    }

    // ---- producer: lane 0 of warp 0 feeds the ring, STAGES - 1 k tiles ahead of the arithmetic ----
    auto produce = [&](int kt) {
        constexpr uint32_t STAGE_BYTES = 2 * TILE_DOUBLES * sizeof(double);
        const int it = it_base + kt - kt0, s = it % STAGES;
        if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) - 1) & 1);      // all 8 warps have released the stage's previous tile
        mbar_expect_tx(&full[s], STAGE_BYTES);
        double* a = sA + s * TILE_DOUBLES;
        double* b = sB + s * TILE_DOUBLES;
        const int k0 = kt * BK;
        if (A_KCONT) { tma_load_2d(a, &tmA, &full[s], k0, m0); tma_load_2d(a + 128 * 16, &tmA, &full[s], k0 + 16, m0); }
        else {
#pragma unroll
            for (int g = 0; g < 8; g++) tma_load_2d(a + g * (BK * 16), &tmA, &full[s], m0 + 16 * g, k0);
        }
        if (B_KCONT) { tma_load_2d(b, &tmB, &full[s], k0, n0); tma_load_2d(b + 128 * 16, &tmB, &full[s], k0 + 16, n0); }
        else {
#pragma unroll
            for (int g = 0; g < 8; g++) tma_load_2d(b + g * (BK * 16), &tmB, &full[s], n0 + 16 * g, k0);
        }
    };
    if (tid == 0) for (int p = 0; p < STAGES - 1 && kt0 + p < kt1; p++) produce(kt0 + p);

    // ---- consumer warps ----
    const int wm = warp & 1, wn = warp >> 1;               // 2 x 4 warps: 64 rows x 32 columns each
    const int fr = lane >> 2, fk = lane & 3;
    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    FragAddr<A_KCONT> fa; fa.init(wm * 64, fr, fk);
    FragAddr<B_KCONT> fb; fb.init(wn * 32, fr, fk);

    for (int kt = kt0; kt < kt1; kt++) {
        const int it = it_base + kt - kt0, s = it % STAGES;
        if (tid == 0 && kt + STAGES - 1 < kt1) produce(kt + STAGES - 1);
        __syncwarp();
        mbar_wait(&full[s], (it / STAGES) & 1);
        const double* a = sA + s * TILE_DOUBLES;
        const double* b = sB + s * TILE_DOUBLES;
#pragma unroll
        for (int ks = 0; ks < BK; ks += 4) {
            double af[MT], bf[NT];
#pragma unroll
            for (int i = 0; i < MT; i++) af[i] = a[fa.off(i, ks)];
#pragma unroll
            for (int j = 0; j < NT; j++) bf[j] = b[fb.off(j, ks)];
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++) gmbgemm::dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    // epilogue: the thread holds C[fragment row fr of tile i][fragment columns 2 fk + {0, 1} of tile j], mapped back through the row permutations
    double colacc[NT][2];
#pragma unroll
    for (int j = 0; j < NT; j++) colacc[j][0] = colacc[j][1] = 0.0;
#pragma unroll
    for (int i = 0; i < MT; i++) {
        const int m = m0 + wm * 64 + frag_row<A_KCONT>(i, fr);
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < NT; j++) {
#pragma unroll
            for (int v = 0; v < 2; v++) {
                const int n = n0 + wn * 32 + frag_row<B_KCONT>(j, 2 * fk + v);
                if (n < N) {
                    epi.store(m, n, acc[i][j][v]);
                    if (Epi::COLSUM) colacc[j][v] += epi.colterm(m, n, acc[i][j][v]);
                }
            }
        }
    }
    if (Epi::COLSUM) {
        // deterministic column sums: lanes with equal lane%4 -> warp ; the two warps stacked in m -> shared memory, fixed order
        __syncthreads();                                     // every warp is done with the operand stages, nothing is in flight
        double* scol = sA;                                   // [2][BN]
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
            for (int v = 0; v < 2; v++) {
                double x = colacc[j][v];
                x += __shfl_xor_sync(0xffffffffu, x, 4);
                x += __shfl_xor_sync(0xffffffffu, x, 8);
                x += __shfl_xor_sync(0xffffffffu, x, 16);
                if (fr == 0) scol[wm * BN + wn * 32 + frag_row<B_KCONT>(j, 2 * fk + v)] = x;
            }
        __syncthreads();
        for (int c = tid; c < BN; c += THREADS) {
            const int n = n0 + c;
            if (n < N) epi.colsum_out(bx, n, scol[c] + scol[BN + c]);
        }
    }
    it_base += max(kt1 - kt0, 0);
    }   // tile loop
    if (queue && tid == 0) {                                 // the last CTA to leave re-arms the queue for the next launch
        __threadfence();
        if (atomicAdd(queue + 1, 1u) == gridDim.x - 1) { queue[0] = 0u; queue[1] = 0u; }
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------
typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline encode_fn get_encode() {
    static encode_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<encode_fn>(p);
    }();
    return fn;
}

// tensor map of a column-major fp64 operand: rows = extent of the contiguous dimension, cols = the other, ld in doubles
inline int make_map(CUtensorMap* tm, const void* ptr, int rows, int cols, int ld, int box_rows, int box_cols, bool swizzle128 = true, bool f32 = false) {
    encode_fn enc = get_encode();
    if (!enc) return gmb_set_error(GMB_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
    cuuint64_t strides[1] = {(cuuint64_t)ld * (f32 ? sizeof(float) : sizeof(double))};
    cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)box_cols};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return gmb_set_error(GMB_ECUDA, "cuTensorMapEncodeTiled failed (%d) for a %d x %d operand, ld %d", (int)r, rows, cols, ld);
    return GMB_OK;
}

// lower tile pairs of the tile columns [c0, c1) of a T x T tile grid, tile column tj holding the tile rows max(tj, rmin) .. T - 1
inline int syrk_tiles_rmin(int T, int c0, int c1, int rmin) { int c = 0; for (int tj = c0; tj < c1 && tj < T; tj++) c += T - (tj > rmin ? tj : rmin); return c; }

template <bool AK, bool BKC, class Epi>
int launch(gmb_ctx* ctx, int M, int N, int K, const double* A, int lda, const double* B, int ldb, const Epi& epi, int tri, int max_ctas = 0, bool dynamic = false) {
    CUtensorMap tmA, tmB;
    // contiguous dimension first: k for a K-contiguous operand (box 16 k x 128 rows), m / n otherwise (box 16 rows x BK k)
    if (AK) GMB_TRY(make_map(&tmA, A, K, M, lda, 16, 128)); else GMB_TRY(make_map(&tmA, A, M, K, lda, 16, BK));
    if (BKC) GMB_TRY(make_map(&tmB, B, K, N, ldb, 16, 128)); else GMB_TRY(make_map(&tmB, B, N, K, ldb, 16, BK));
    auto kern = dgemm_tma_kernel<AK, BKC, Epi>;
    static bool configured = false;   // per instantiation
    if (!configured) {
        GMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        configured = true;
    }
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
    int ptiles = 0;
    if ((tri & 0xff) == 3) {
        const int tiles = syrk_tiles_rmin((M + BM - 1) / BM, (tri >> 8) & 0xfff, (N + BN - 1) / BN, (tri >> 20) & 0xfff);
        if (tiles <= 0) return GMB_OK;
        grid = dim3(tiles, 1);
        if (max_ctas > 0) { ptiles = tiles; grid = dim3(tiles < max_ctas ? tiles : max_ctas, 1); }
    } else if (max_ctas > 0) {
        ptiles = (int)(grid.x * grid.y);
        grid = dim3(ptiles < max_ctas ? ptiles : max_ctas, 1);
    }
    // tile queue of a dynamic launch: one pair of counters per stream of the context (zero between launches)
    unsigned int* queue = nullptr;
    if (dynamic && ptiles > 0) queue = ctx->d_counter + 58 + 2 * (ctx->stream == ctx->stream2 ? 1 : ctx->stream == ctx->stream3 ? 2 : 0);
    kern<<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(tmA, tmB, M, N, K, epi, tri, ptiles, queue);
    ctx->launches++;
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}

// Which kernel a product runs on: the TMA kernel (128 x 128 tiles, one CTA per SM) when its tiles fill the machine — at least 6 waves, or waves
// that are at least 90 % full — and the k loop is long enough to amortise the 3-stage ring; otherwise the cp.async kernel with 64 x 64
// tiles (several CTAs per SM, finer wave granularity).  GMB_GEMM_TMA=0 disables the TMA kernel, =2 forces it (tests, profiling).
inline int gemm_tma_mode() { static const int mode = [] { const char* e = getenv("GMB_GEMM_TMA"); return e ? atoi(e) : 1; }(); return mode; }
inline bool gemm_tri_queue_enabled() { static const int on = [] { const char* e = getenv("GMB_GEMM_TRI_QUEUE"); return e ? atoi(e) : 1; }(); return on != 0; }
inline bool use_tma(gmb_ctx* ctx, int M, int N, int K) {
    const int mode = gemm_tma_mode();
    if (mode == 0 || !get_encode()) return false;
    if (mode == 2) return true;
    if (K < 64) return false;
    const long tiles = (long)((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const long rounds = (tiles + ctx->sms - 1) / ctx->sms;
    return tiles >= 6L * ctx->sms || (double)tiles >= 0.9 * (double)(rounds * ctx->sms);
}
// number of rows per row tile the dispatcher below will use (callers size their column-sum buffers with it)
inline int row_tile(gmb_ctx* ctx, int M, int N, int K) { return use_tma(ctx, M, N, K) ? BM : 64; }

template <bool AK, bool BKC, class Epi>
int dispatch(gmb_ctx* ctx, int M, int N, int K, const double* A, int lda, const double* B, int ldb, const Epi& epi, int tri = 0) {
    if (M <= 0 || N <= 0) return GMB_OK;
    if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (lda & 1) || (ldb & 1))
        return gmb_set_error(GMB_EINVAL, "dgemm: operands must be 16-byte aligned with even leading dimensions");
    // tri = 4 (a triangular solve as a product, the Cholesky's panel step): unequal tiles, heaviest first — worth the big tiles from half a wave on
    const bool tri4_tma = tri == 4 && gemm_tma_mode() != 0 && get_encode() && K >= 64 && 2L * ((M + BM - 1) / BM) * ((N + BN - 1) / BN) >= ctx->sms;
    // triangular A with unequal tiles: persistent launch with a heaviest-first tile queue as soon as there is a tile per SM
    const long tiles_all = (long)((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    if ((tri == 1 || tri == 2) && !Epi::COLSUM && gemm_tma_mode() != 0 && get_encode() && K >= 256 && tiles_all >= ctx->sms && gemm_tri_queue_enabled())
        return launch<AK, BKC, Epi>(ctx, M, N, K, A, lda, B, ldb, epi, tri, ctx->sms, true);
    if (use_tma(ctx, M, N, K) || tri4_tma) return launch<AK, BKC, Epi>(ctx, M, N, K, A, lda, B, ldb, epi, tri);
    return gmbgemm::launch<64, 64, 2, 4, AK, BKC, Epi>(ctx, M, N, K, A, lda, B, ldb, epi, tri);
}

// C[:, c0:c1) (lower tiles) = epi(P P^T) for the M x K matrix P (m contiguous): column range in elements, multiples of 128
template <class Epi>
int dispatch_syrk_lower(gmb_ctx* ctx, int M, int K, const double* Pm, int ldp, const Epi& epi, int c0, int c1, int rmin_tiles = 0, int max_ctas = 0) {
    if (M <= 0 || c1 <= c0) return GMB_OK;
    if (((uintptr_t)Pm & 15) || (ldp & 1)) return gmb_set_error(GMB_EINVAL, "dsyrk: operand must be 16-byte aligned with an even leading dimension");
    if ((c0 % 128) || (c1 % 128 && c1 < M)) return gmb_set_error(GMB_EINVAL, "dsyrk: the column range must be made of whole 128-column tiles");
    const int N = c1 < M ? c1 : M;
    const long tiles = gmbgemm::syrk_tiles((M + 127) / 128, c0 / 128, (N + 127) / 128);
    const long rounds = (tiles + ctx->sms - 1) / ctx->sms;
    const bool tma = gemm_tma_mode() != 0 && get_encode() && K >= 64 && (gemm_tma_mode() == 2 || tiles >= 6L * ctx->sms || (double)tiles >= 0.9 * (double)(rounds * ctx->sms));
    if (tma || rmin_tiles > 0 || max_ctas > 0) {
        if (!get_encode()) return gmb_set_error(GMB_ECUDA, "dsyrk: the persistent update needs the TMA kernel");
        return launch<false, false, Epi>(ctx, M, N, K, Pm, ldp, Pm, ldp, epi, 3 | ((c0 / 128) << 8) | (rmin_tiles << 20), max_ctas);
    }
    return gmbgemm::launch<64, 64, 2, 4, false, false, Epi>(ctx, M, N, K, Pm, ldp, Pm, ldp, epi, 3 | ((c0 / 64) << 8));
}

}  // namespace gmbtma
