// gemm_f64.cuh — FP64 GEMM mainloop on the DMMA path (mma.sync.m8n8k4.f64) with a cp.async 3-stage pipeline and a
// pluggable epilogue.  Included by gemm_f64.cu (plain alpha/beta epilogue) and hmc.cu (family-residual and leapfrog
// epilogues fused into the Z L v / (Z L)^T r contractions of the sampler).
//
// tcgen05 has no f64 kind, so the FP64 tensor path on sm_100a is mma.sync; measured peak on B200 is 37.1 TFLOP/s for
// every DMMA shape (profiles/r01_microbench_fp64.txt), the same as the DFMA peak — the MMA form is used because it needs
// one shared-memory operand load per 256 FMAs instead of one per 32.
//
// C (M x N) = op(A) * op(B), all column-major:
//   A_KCONT = false: A is M x K (m contiguous) ; A_KCONT = true: A is stored K x M (k contiguous), i.e. op(A) = A^T
//   B_KCONT = true : B is K x N (k contiguous) ; B_KCONT = false: B is stored N x K (n contiguous), i.e. op(B) = B^T
// Requirements: pointers 16-byte aligned, lda/ldb even.
//
// Epilogue concept:
//   static constexpr bool COLSUM;                       // also reduce a per-column term over the rows of the CTA tile
//   __device__ bool column_active(int n) const;         // CTA exits early when none of its columns is active
//   __device__ void store(int m, int n, double acc) const;
//   __device__ double colterm(int m, int n, double acc) const;       (COLSUM only)
//   __device__ void colsum_out(int row_tile, int n, double v) const; (COLSUM only; one call per (row tile, column))
#pragma once
#include "common.cuh"
#include <cstdlib>

namespace gmbgemm {

constexpr int BK = 16;
constexpr int STAGES = 3;
constexpr int THREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// shared-memory tile of one operand: BMN x BK.
//   KCONT = false: stored [k][mn], ld = BMN + 4   (source is mn-contiguous)
//   KCONT = true : stored [mn][k], ld = BK + 4    (source is k-contiguous)
// Both leading dimensions are = 4 (mod 16) doubles, which makes the m8n8k4 fragment reads
// (8 values of mn x 4 values of k per warp) hit 32 distinct 8-byte banks in each half-warp.
template <int BMN, bool KCONT>
struct OperandTile {
    static constexpr int LD = KCONT ? (BK + 4) : (BMN + 4);
    static constexpr int DOUBLES = KCONT ? BMN * LD : BK * LD;
    __device__ static __forceinline__ int idx(int mn, int k) { return KCONT ? mn * LD + k : k * LD + mn; }

    // cooperative global -> shared copy of the tile whose origin is (mn0, k0); zero-fills out of range.
    __device__ static __forceinline__ void load(double* s, const double* __restrict__ g, int ld, int mn0, int k0,
                                                int MN, int K, int tid) {
        if (KCONT) {
            constexpr int CH_PER_ROW = BK / 2;
            constexpr int CHUNKS = BMN * CH_PER_ROW;
#pragma unroll
            for (int c = tid; c < CHUNKS; c += THREADS) {
                int mn = c / CH_PER_ROW, k = (c % CH_PER_ROW) * 2;
                int gmn = mn0 + mn, gk = k0 + k;
                int valid = (gmn < MN) ? min(max(K - gk, 0), 2) : 0;
                const double* src = valid ? g + (size_t)gmn * ld + gk : g;
                cp_async16(s + mn * LD + k, src, valid * 8);
            }
        } else {
            constexpr int CH_PER_COL = BMN / 2;
            constexpr int CHUNKS = BK * CH_PER_COL;
#pragma unroll
            for (int c = tid; c < CHUNKS; c += THREADS) {
                int k = c / CH_PER_COL, mn = (c % CH_PER_COL) * 2;
                int gmn = mn0 + mn, gk = k0 + k;
                int valid = (gk < K) ? min(max(MN - gmn, 0), 2) : 0;
                const double* src = valid ? g + (size_t)gk * ld + gmn : g;
                cp_async16(s + k * LD + mn, src, valid * 8);
            }
        }
    }
};

template <int BM, int BN, int WM, int WN, bool A_KCONT, bool B_KCONT, class Epi>
__global__ void __launch_bounds__(THREADS) dgemm_kernel(int M, int N, int K, const double* __restrict__ A, int lda,
                                                        const double* __restrict__ B, int ldb, Epi epi, int tri) {
    using TA = OperandTile<BM, A_KCONT>;
    using TB = OperandTile<BN, B_KCONT>;
    constexpr int WTM = BM / WM, WTN = BN / WN;       // warp tile
    constexpr int MT = WTM / 8, NT = WTN / 8;         // m8n8 tiles per warp
    static_assert(WM * WN * 32 == THREADS, "warp layout");
    extern __shared__ __align__(16) double smem[];
    double* sA = smem;
    double* sB = smem + STAGES * TA::DOUBLES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;
    // tri = 1: the last row tiles carry the longest k ranges — start them first (CTAs are scheduled in index order)
    // tri & 0xff = 3: symmetric rank-k update, lower tiles only (BM == BN): blockIdx.x enumerates the tile pairs (ti >= tj) of the tile columns
    // tj >= tri >> 8, column by column — every CTA has the same amount of work and no tile of the upper triangle is computed
    int bx = (tri == 1) ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x, by = blockIdx.y;
    if ((tri & 0xff) == 3) {
        const int T = (M + BM - 1) / BM;
        int rem = blockIdx.x, tj = tri >> 8;
        while (rem >= T - tj) { rem -= T - tj; tj++; }
        bx = tj + rem; by = tj;
        tri = 0;
    }
    const int m0 = bx * BM, n0 = by * BN;
    const int KT = (K + BK - 1) / BK;
    // triangular operand (a Cholesky factor): tri = 1: op(A)(m, k) = 0 for k > m ; tri = 2: op(A)(m, k) = 0 for k < m (the factor read transposed).
    // The k tiles that hold only zeros for this row tile are skipped — the same sums without their zero terms.
    const int kt0 = (tri == 2) ? min(m0 / BK, KT) : 0;
    const int kt1 = (tri == 1) ? min(KT, (m0 + BM + BK - 1) / BK) : (tri == 4) ? min(KT, (n0 + BN + BK - 1) / BK) : KT;   // tri = 4: B(n, k) = 0 for k > n

    {   // skip tiles whose columns are all inactive (chains that finished their trajectory)
        int act = 0;
        for (int c = tid; c < BN; c += THREADS) if (n0 + c < N && epi.column_active(n0 + c)) act = 1;
        if (!__syncthreads_or(act)) return;
    }

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (kt0 + s < kt1) {
            TA::load(sA + s * TA::DOUBLES, A, lda, m0, (kt0 + s) * BK, M, K, tid);
            TB::load(sB + s * TB::DOUBLES, B, ldb, n0, (kt0 + s) * BK, N, K, tid);
        }
        cp_async_commit();
    }

    const int fr = lane >> 2, fk = lane & 3;   // fragment row (m or n) and k within the k4 step
    for (int kt = kt0; kt < kt1; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < kt1) {
                int s = (nk - kt0) % STAGES;
                TA::load(sA + s * TA::DOUBLES, A, lda, m0, nk * BK, M, K, tid);
                TB::load(sB + s * TB::DOUBLES, B, ldb, n0, nk * BK, N, K, tid);
            }
            cp_async_commit();
        }
        const double* a = sA + ((kt - kt0) % STAGES) * TA::DOUBLES;
        const double* b = sB + ((kt - kt0) % STAGES) * TB::DOUBLES;
#pragma unroll
        for (int ks = 0; ks < BK; ks += 4) {
            double af[MT], bf[NT];
#pragma unroll
            for (int i = 0; i < MT; i++) af[i] = a[TA::idx(wm * WTM + i * 8 + fr, ks + fk)];
#pragma unroll
            for (int j = 0; j < NT; j++) bf[j] = b[TB::idx(wn * WTN + j * 8 + fr, ks + fk)];
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: thread holds C[m = 8i + lane/4][n = 8j + 2*(lane%4) + {0,1}]
    double colacc[NT][2];
#pragma unroll
    for (int j = 0; j < NT; j++) colacc[j][0] = colacc[j][1] = 0.0;
#pragma unroll
    for (int i = 0; i < MT; i++) {
        int m = m0 + wm * WTM + i * 8 + fr;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < NT; j++) {
#pragma unroll
            for (int v = 0; v < 2; v++) {
                int n = n0 + wn * WTN + j * 8 + 2 * fk + v;
                if (n < N) {
                    epi.store(m, n, acc[i][j][v]);
                    if (Epi::COLSUM) colacc[j][v] += epi.colterm(m, n, acc[i][j][v]);
                }
            }
        }
    }
    if (Epi::COLSUM) {
        // deterministic column sums: lanes with equal lane%4 -> warp ; warps stacked in m -> shared memory, fixed order
        __syncthreads();                       // the pipeline buffers are free now
        double* scol = smem;                   // [WM][BN]
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
            for (int v = 0; v < 2; v++) {
                double x = colacc[j][v];
                x += __shfl_xor_sync(0xffffffffu, x, 4);
                x += __shfl_xor_sync(0xffffffffu, x, 8);
                x += __shfl_xor_sync(0xffffffffu, x, 16);
                if (fr == 0) scol[wm * BN + wn * WTN + j * 8 + 2 * fk + v] = x;
            }
        __syncthreads();
        for (int c = tid; c < BN; c += THREADS) {
            int n = n0 + c;
            if (n < N) {
                double x = 0.0;
#pragma unroll
                for (int w = 0; w < WM; w++) x += scol[w * BN + c];
                epi.colsum_out(bx, n, x);
            }
        }
    }
}

// lower tile pairs (ti >= tj) with tj in [c0, c1) of a T x T tile grid
inline int syrk_tiles(int T, int c0, int c1) { int c = 0; for (int tj = c0; tj < c1 && tj < T; tj++) c += T - tj; return c; }

template <int BM, int BN, int WM, int WN, bool AK, bool BKC, class Epi>
int launch(gmb_ctx* ctx, int M, int N, int K, const double* A, int lda, const double* B, int ldb, const Epi& epi, int tri = 0) {
    using TA = OperandTile<BM, AK>;
    using TB = OperandTile<BN, BKC>;
    size_t smem = (size_t)STAGES * (TA::DOUBLES + TB::DOUBLES) * sizeof(double);
    auto kern = dgemm_kernel<BM, BN, WM, WN, AK, BKC, Epi>;
    static bool configured = false;   // per instantiation
    if (!configured) {
        GMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
    if ((tri & 0xff) == 3) grid = dim3(syrk_tiles((M + BM - 1) / BM, tri >> 8, (N + BN - 1) / BN), 1);
    kern<<<grid, THREADS, smem, ctx->stream>>>(M, N, K, A, lda, B, ldb, epi, tri);
    ctx->launches++;
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}

}  // namespace gmbgemm
