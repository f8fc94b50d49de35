// cov_large.cu — covariance blocks larger than a warp: two-level blocked right-looking Cholesky and two-level blocked
// forward substitution with many right-hand sides, both built on the DMMA GEMM (gemm_f64.cu).
//
// Replaces, for one dense block (e.g. the fexp Gaussian-process block of configs C3/C5), glmmrBase
// gen_block_mat(b, true, false) (unblocked Cholesky–Banachiewicz) and the per-sample
// algo::forward_sub of mcmldmatrix.h:67-75 / moremaths.h:166-179.
#include "common.cuh"

namespace {

constexpr int NB = 128;   // panel width

// fill the lower triangle of the block with D_b(i,j) (upper triangle zero)
__global__ void build_block_kernel(CovBlock b, const CovFn* __restrict__ fns, const double* __restrict__ data,
                                   const double* __restrict__ theta, double* __restrict__ A, int ld) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= b.n || j >= b.n) return;
    double v = 0.0;
    if (j <= i) {
        v = 1.0;
        const double* dat = data + b.data0;
        for (int f = 0; f < b.nfn; f++) {
            const CovFn fn = fns[b.fn0 + f];
            double d2 = 0.0;
            for (int k = 0; k < fn.nvar; k++) {
                double di = dat[i + (size_t)(fn.col0 + k) * b.n] - dat[j + (size_t)(fn.col0 + k) * b.n];
                d2 += di * di;
            }
            v *= dev_cov_fn(fn.id, sqrt(d2), theta + fn.par0, fn.eff);
        }
    }
    A[i + (size_t)j * ld] = v;
}

// Cholesky factor AND its inverse of one kb x kb (kb <= 128) diagonal block at (k0, k0): one CTA of 256 threads, both matrices in registers.
// Thread (tx, ty) = (tid % 16, tid / 16) owns the elements (r, c) = (tx + 16 i, ty + 16 jj) of the LOWER triangles (i >= jj: 36 of the 64
// index pairs) of A and of B = L^-1 (B starts as the identity) — the interleaving keeps every thread busy while the active part of A shrinks
// and the active part of B grows.  One right-looking elimination produces both: at step j the owners of column j of A and of row j of B publish
// them through (double-buffered) shared memory — one block barrier per step —, every thread scales by rsqrt(pivot), column j of L is final,
// row j of L^-1 is final, and the rank-1 updates A[r, c] -= l_r l_c (r, c > j) and B[r, c] -= l_r x_c (r > j, c <= j) are applied to its tiles.
// The panel solve of the factorisation and the diagonal solves of the forward substitution then run as DMMA GEMMs with the inverse instead of
// one serial recurrence per row.  Rows / columns >= kb are padded with the identity.  A: the lower triangle is read; L is written to the lower
// triangle, the strict upper triangle of the block is zeroed.  Linv: 128 x 128 column-major, zero outside the kb x kb lower triangle.
__device__ __forceinline__ unsigned pd_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pd_bar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pd_smem(bar)) : "memory"); }
__device__ __forceinline__ void pd_bar_wait(uint64_t* bar, unsigned parity) {
    unsigned ok = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(pd_smem(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// One elimination step j = 16 JQ + js.  JQN = slot of column / row j + 1 (JQ, or JQ + 1 when js = 15).  The entries of column j + 1 of A and
// of row j + 1 of B are updated FIRST and published for step j + 1 (one mbarrier arrival per warp); the rest of the rank-1 update follows, off
// the critical path: while it runs, the next pivot's rsqrt is already under way in the warps that are done.
template <int JQ, int JQN>
__device__ __forceinline__ bool potrf_step(int js, int tx, int ty, int lane, double (&a)[8][8], double (&bm)[8][8],
                                           double (*colbuf)[NB], double (*rowbuf)[NB], uint64_t* bars, int* status, int pivot_id) {
    const int j = 16 * JQ + js;
    pd_bar_wait(&bars[j & 1], (j >> 1) & 1);
    const double* cb = colbuf[j & 1];
    const double* rb = rowbuf[j & 1];
    const double d = cb[j];
    if (!(d > 0.0)) { if (threadIdx.x == 0) atomicCAS(status, 0, pivot_id + j + 1); return false; }    // uniform: every thread reads the same pivot
    const double inv = rsqrt(d);
    double lr[8], lc[8], xc[8];
#pragma unroll
    for (int i = JQ; i < 8; i++) { lr[i] = cb[tx + 16 * i] * inv; lc[i] = cb[ty + 16 * i] * inv; }
#pragma unroll
    for (int jj = 0; jj <= JQ; jj++) xc[jj] = rb[ty + 16 * jj] * inv;
    // rows r > j of slot i, columns c > j of slot jj
#define PD_ROK(i) (((i) > JQ) || (tx > js))
#define PD_COK(jj) (((jj) > JQ) || (ty > js))
    if (JQN < 8) {
        // ---- priority: column slot JQN of A (all its rows), row slot JQN of B ----
#pragma unroll
        for (int i = JQN; i < 8; i++)
            if (PD_ROK(i) && PD_COK(JQN)) a[i][JQN] = fma(-lr[i], lc[JQN], a[i][JQN]);
        if (PD_ROK(JQN)) {
#pragma unroll
            for (int jj = 0; jj <= JQ; jj++) bm[JQN][jj] = fma(-lr[JQN], xc[jj], bm[JQN][jj]);
        }
        const int jn = j + 1, jsn = jn & 15;
        double* cbn = colbuf[jn & 1];
        double* rbn = rowbuf[jn & 1];
        if (ty == jsn) {                          // owners of column j + 1 of A: rows r >= j + 1 live in slots i >= JQN
#pragma unroll
            for (int i = JQN; i < 8; i++) cbn[tx + 16 * i] = a[i][JQN];
        }
        if (tx == jsn) {                          // owners of row j + 1 of B: columns c <= j + 1 live in slots jj <= JQN
#pragma unroll
            for (int jj = 0; jj <= JQN; jj++) rbn[ty + 16 * jj] = bm[JQN][jj];
        }
        __syncwarp();
        if (lane == 0) pd_bar_arrive(&bars[jn & 1]);
    }
    // ---- the rest of the rank-1 update ----
#pragma unroll
    for (int i = JQ; i < 8; i++) {
        if (PD_ROK(i)) {
#pragma unroll
            for (int jj = JQ; jj <= i; jj++)               // A: columns c > j of the lower triangle
                if (jj != JQN && PD_COK(jj)) a[i][jj] = fma(-lr[i], lc[jj], a[i][jj]);
            if (i != JQN) {
#pragma unroll
                for (int jj = 0; jj <= JQ; jj++)           // B: columns c <= j (entries right of the diagonal of row j are zero)
                    bm[i][jj] = fma(-lr[i], xc[jj], bm[i][jj]);
            }
        }
    }
#undef PD_ROK
#undef PD_COK
    if (ty == js) {                               // column j of L
#pragma unroll
        for (int i = JQ; i < 8; i++) {
            if (i > JQ || tx > js) a[i][JQ] = lr[i];
            else if (tx == js) a[i][JQ] = d * inv;
        }
    }
    if (tx == js) {                               // row j of L^-1
#pragma unroll
        for (int jj = 0; jj <= JQ; jj++) bm[JQ][jj] = xc[jj];
    }
    return true;
}

template <int JQ>
__device__ __forceinline__ bool potrf_phase(int tx, int ty, int lane, double (&a)[8][8], double (&bm)[8][8], double (*colbuf)[NB], double (*rowbuf)[NB],
                                            uint64_t* bars, int* status, int pivot_id) {
#pragma unroll 1
    for (int js = 0; js < 15; js++)
        if (!potrf_step<JQ, JQ>(js, tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id)) return false;
    return potrf_step<JQ, JQ + 1>(15, tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
}

__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ A, int ld, int k0, int kb, int row_offset, int* __restrict__ status,
                                                         double* __restrict__ Linv) {
    __shared__ double colbuf[2][NB], rowbuf[2][NB];
    __shared__ uint64_t bars[2];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pd_smem(&bars[0])), "r"(8));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pd_smem(&bars[1])), "r"(8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    double a[8][8], bm[8][8];                     // only the entries with i >= jj are ever touched (the others are never materialised)
#pragma unroll
    for (int jj = 0; jj < 8; jj++)
#pragma unroll
        for (int i = jj; i < 8; i++) {
            const int r = tx + 16 * i, c = ty + 16 * jj;
            double v = (r == c) ? 1.0 : 0.0;
            if (r < kb && c < kb && r >= c) v = A[(size_t)(k0 + r) + (size_t)(k0 + c) * ld];
            a[i][jj] = v;
            bm[i][jj] = (r == c) ? 1.0 : 0.0;
        }
    __syncthreads();                              // barriers initialised
    if (ty == 0) {                                // column 0 of A, row 0 of B for step 0
#pragma unroll
        for (int i = 0; i < 8; i++) colbuf[0][tx + 16 * i] = a[i][0];
    }
    if (tx == 0) rowbuf[0][ty] = bm[0][0];
    __syncwarp();
    if (lane == 0) pd_bar_arrive(&bars[0]);
    const int pivot_id = row_offset + k0;
    bool ok = potrf_phase<0>(tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
    ok = ok && potrf_phase<1>(tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
    ok = ok && potrf_phase<2>(tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
    ok = ok && potrf_phase<3>(tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
    ok = ok && potrf_phase<4>(tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
    ok = ok && potrf_phase<5>(tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
    ok = ok && potrf_phase<6>(tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
    ok = ok && potrf_phase<7>(tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
    if (!ok) return;
#pragma unroll
    for (int jj = 0; jj < 8; jj++)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int r = tx + 16 * i, c = ty + 16 * jj;
            const bool low = (i > jj) || (i == jj && r >= c);
            if (r < kb && c < kb) A[(size_t)(k0 + r) + (size_t)(k0 + c) * ld] = (i >= jj && low) ? a[i >= jj ? i : jj][jj] : 0.0;
            Linv[r + (size_t)c * NB] = (i >= jj && low && r < kb && c < kb) ? bm[i >= jj ? i : jj][jj] : 0.0;
        }
}

// 1 / sqrt(d) for a positive, normal d: hardware seed (rsqrt.approx.ftz.f64, relative error < 2^-22) and two Newton steps y += (y/2)(1 - d y^2)
// — six dependent FP64 operations, ~1 ulp, without the special-case handling of the library function (the callers have checked d > 0; a
// pivot so small or large that the seed leaves the normal range is a breakdown of the factorisation anyway)
__device__ __forceinline__ double pd_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const double e = fma(-d * y, y, 1.0);
        y = fma(0.5 * y, e, y);
    }
    return y;
}

// ---- two columns per hand-off ------------------------------------------------------------------------------------------------------
// The elimination above pays one hand-off (publish -> mbarrier -> shared-memory read) per column, and the hand-off, not the arithmetic, is
// what a step costs.  Here a step eliminates the PAIR of columns j, j + 1: their owners publish the raw columns c0, c1 (updates through
// column j - 1 applied) and the raw rows b0, b1 of B; every thread derives both factor columns itself —
//     l0 = c0 / sqrt(c0_j),   l10 = l0_{j+1},   c1' = c1 - l10 l0,   l1 = c1' / sqrt(c1'_{j+1}),      x0 = b0 / sqrt(c0_j),   x1 = (b1 - l10 x0) / sqrt(c1'_{j+1})
// — and applies the rank-2 update A -= l0 l0' + l1 l1', B -= l0 x0' + l1 x1': two reciprocal square roots in sequence, but half the hand-offs.
template <int JQ, int JQN>
__device__ __forceinline__ bool potrf_step2(int js /* even */, int tx, int ty, int lane, double (&a)[8][8], double (&bm)[8][8],
                                            double (*colbuf)[2][NB], double (*rowbuf)[2][NB], uint64_t* bars, int* status, int pivot_id) {
    const int j = 16 * JQ + js, pr = j >> 1;                      // first column of the pair, pair index
    pd_bar_wait(&bars[pr & 1], (pr >> 1) & 1);
    const double* c0 = colbuf[pr & 1][0];
    const double* c1 = colbuf[pr & 1][1];
    const double* b0 = rowbuf[pr & 1][0];
    const double* b1 = rowbuf[pr & 1][1];
    const double d0 = c0[j];
    if (!(d0 > 0.0)) { if (threadIdx.x == 0) atomicCAS(status, 0, pivot_id + j + 1); return false; }
    const double inv0 = pd_rsqrt(d0);
    const double l10 = c0[j + 1] * inv0;
    const double d1 = fma(-l10, l10, c1[j + 1]);
    if (!(d1 > 0.0)) { if (threadIdx.x == 0) atomicCAS(status, 0, pivot_id + j + 2); return false; }
    const double inv1 = pd_rsqrt(d1);
    double l0r[8], l1r[8], l0c[8], l1c[8], x0[8], x1[8];
#pragma unroll
    for (int i = JQ; i < 8; i++) {
        l0r[i] = c0[tx + 16 * i] * inv0; l1r[i] = fma(-l0r[i], l10, c1[tx + 16 * i]) * inv1;
        l0c[i] = c0[ty + 16 * i] * inv0; l1c[i] = fma(-l0c[i], l10, c1[ty + 16 * i]) * inv1;
    }
#pragma unroll
    for (int jj = 0; jj <= JQ; jj++) { x0[jj] = b0[ty + 16 * jj] * inv0; x1[jj] = fma(-l10, x0[jj], b1[ty + 16 * jj]) * inv1; }
    // rows r > j + 1 of slot i, columns c > j + 1 of slot jj
#define PD_ROK(i) (((i) > JQ) || (tx > js + 1))
#define PD_COK(jj) (((jj) > JQ) || (ty > js + 1))
    if (JQN < 8) {
        // ---- priority: column slot JQN of A (it holds the next pair's columns), row slot JQN of B ----
#pragma unroll
        for (int i = JQN; i < 8; i++)
            if (PD_ROK(i) && PD_COK(JQN)) a[i][JQN] = fma(-l1r[i], l1c[JQN], fma(-l0r[i], l0c[JQN], a[i][JQN]));
        if (PD_ROK(JQN)) {
#pragma unroll
            for (int jj = 0; jj <= JQ; jj++) bm[JQN][jj] = fma(-l1r[JQN], x1[jj], fma(-l0r[JQN], x0[jj], bm[JQN][jj]));
        }
        const int jn = j + 2, jsn = jn & 15, pn = jn >> 1;
        if (ty == jsn || ty == jsn + 1) {         // owners of columns j + 2, j + 3 of A: rows r >= column live in slots i >= JQN
            double* dst = colbuf[pn & 1][ty - jsn];
#pragma unroll
            for (int i = JQN; i < 8; i++) dst[tx + 16 * i] = a[i][JQN];
        }
        if (tx == jsn || tx == jsn + 1) {         // owners of rows j + 2, j + 3 of B: columns c <= row live in slots jj <= JQN
            double* dst = rowbuf[pn & 1][tx - jsn];
#pragma unroll
            for (int jj = 0; jj <= JQN; jj++) dst[ty + 16 * jj] = bm[JQN][jj];
        }
        __syncwarp();
        if (lane == 0) pd_bar_arrive(&bars[pn & 1]);
    }
    // ---- the rest of the rank-2 update ----
#pragma unroll
    for (int i = JQ; i < 8; i++) {
        if (PD_ROK(i)) {
#pragma unroll
            for (int jj = JQ; jj <= i; jj++)
                if (jj != JQN && PD_COK(jj)) a[i][jj] = fma(-l1r[i], l1c[jj], fma(-l0r[i], l0c[jj], a[i][jj]));
            if (i != JQN) {
#pragma unroll
                for (int jj = 0; jj <= JQ; jj++) bm[i][jj] = fma(-l1r[i], x1[jj], fma(-l0r[i], x0[jj], bm[i][jj]));
            }
        }
    }
#undef PD_ROK
#undef PD_COK
    if (ty == js) {                               // column j of L (row j + 1 of it is l10)
#pragma unroll
        for (int i = JQ; i < 8; i++) {
            if (i > JQ || tx > js) a[i][JQ] = l0r[i];
            else if (tx == js) a[i][JQ] = d0 * inv0;
        }
    }
    if (ty == js + 1) {                           // column j + 1 of L
#pragma unroll
        for (int i = JQ; i < 8; i++) {
            if (i > JQ || tx > js + 1) a[i][JQ] = l1r[i];
            else if (tx == js + 1) a[i][JQ] = d1 * inv1;
        }
    }
    if (tx == js) {                               // row j of L^-1
#pragma unroll
        for (int jj = 0; jj <= JQ; jj++) bm[JQ][jj] = x0[jj];
    }
    if (tx == js + 1) {                           // row j + 1 of L^-1
#pragma unroll
        for (int jj = 0; jj <= JQ; jj++) bm[JQ][jj] = x1[jj];
    }
    return true;
}

template <int JQ>
__device__ __forceinline__ bool potrf_phase2(int tx, int ty, int lane, double (&a)[8][8], double (&bm)[8][8], double (*colbuf)[2][NB], double (*rowbuf)[2][NB],
                                             uint64_t* bars, int* status, int pivot_id) {
#pragma unroll 1
    for (int js = 0; js < 14; js += 2)
        if (!potrf_step2<JQ, JQ>(js, tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id)) return false;
    return potrf_step2<JQ, JQ + 1>(14, tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
}

__global__ void __launch_bounds__(256) potrf_diag2_kernel(double* __restrict__ A, int ld, int k0, int kb, int row_offset, int* __restrict__ status,
                                                          double* __restrict__ Linv) {
    __shared__ double colbuf[2][2][NB], rowbuf[2][2][NB];
    __shared__ uint64_t bars[2];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pd_smem(&bars[0])), "r"(8));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pd_smem(&bars[1])), "r"(8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    double a[8][8], bm[8][8];
#pragma unroll
    for (int jj = 0; jj < 8; jj++)
#pragma unroll
        for (int i = jj; i < 8; i++) {
            const int r = tx + 16 * i, c = ty + 16 * jj;
            double v = (r == c) ? 1.0 : 0.0;
            if (r < kb && c < kb && r >= c) v = A[(size_t)(k0 + r) + (size_t)(k0 + c) * ld];
            a[i][jj] = v;
            bm[i][jj] = (r == c) ? 1.0 : 0.0;
        }
    __syncthreads();                              // barriers initialised
    if (ty < 2) {                                 // columns 0, 1 of A
#pragma unroll
        for (int i = 0; i < 8; i++) colbuf[0][ty][tx + 16 * i] = a[i][0];
    }
    if (tx < 2) rowbuf[0][tx][ty] = bm[0][0];     // rows 0, 1 of B (columns c <= 1 live in slot 0; c > row holds 0)
    __syncwarp();
    if (lane == 0) pd_bar_arrive(&bars[0]);
    const int pivot_id = row_offset + k0;
    bool ok = potrf_phase2<0>(tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
    ok = ok && potrf_phase2<1>(tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
    ok = ok && potrf_phase2<2>(tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
    ok = ok && potrf_phase2<3>(tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
    ok = ok && potrf_phase2<4>(tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
    ok = ok && potrf_phase2<5>(tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
    ok = ok && potrf_phase2<6>(tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
    ok = ok && potrf_phase2<7>(tx, ty, lane, a, bm, colbuf, rowbuf, bars, status, pivot_id);
    if (!ok) return;
#pragma unroll
    for (int jj = 0; jj < 8; jj++)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int r = tx + 16 * i, c = ty + 16 * jj;
            const bool low = (i > jj) || (i == jj && r >= c);
            if (r < kb && c < kb) A[(size_t)(k0 + r) + (size_t)(k0 + c) * ld] = (i >= jj && low) ? a[i >= jj ? i : jj][jj] : 0.0;
            Linv[r + (size_t)c * NB] = (i >= jj && low && r < kb && c < kb) ? bm[i >= jj ? i : jj][jj] : 0.0;
        }
}

__global__ void logdet_diag_kernel(const double* __restrict__ A, int ld, int n, double* __restrict__ out) {
    __shared__ double red[32];
    double c = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) c += 2.0 * log(A[i + (size_t)i * ld]);
    c = block_sum(c, red);
    if (threadIdx.x == 0) out[0] = c;
}

__global__ void __launch_bounds__(256) sumsq_kernel(const double* __restrict__ W, int ldw, int n, int ncols, double* __restrict__ partials) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int j = blockIdx.x; j < ncols; j += gridDim.x) {
        const double* w = W + (size_t)j * ldw;
        for (int i = threadIdx.x; i < n; i += blockDim.x) acc += w[i] * w[i];
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] += acc;      // slots are zeroed by the caller; launches that share a slot are ordered on one stream
}

// lower != 0: the source is a lower-triangular factor whose strict upper triangle holds scratch: column c0 + j is zero above row c0 + j
__global__ void copy_rows_kernel(const double* __restrict__ U, int ldu, int start, int n, int ncols, double* __restrict__ W, int ldw, int lower, int c0) {
    int i = blockIdx.y * blockDim.x + threadIdx.x;      // grid.x runs over the columns (no 65535 limit)
    int j = blockIdx.x;
    if (i < ldw && j < ncols) W[(size_t)j * ldw + i] = (i < n && (!lower || i >= c0 + j)) ? U[(size_t)j * ldu + start + i] : 0.0;
}

}  // namespace

// Two-level blocking with look-ahead.  Panels of NB = 128 columns: potrf_diag_kernel factorises and inverts the diagonal block, the rows
// below are solved as one GEMM with the inverse, and the panel's rank-128 update is applied right away only INSIDE the current outer block of
// NBO = 512 columns.  The trailing matrix receives one rank-512 update per outer block (long k loops for the DMMA GEMM), restricted to the
// lower trapezoid by block columns — split in two: the NARROW part (the next outer block's columns) first, then the WIDE rest.  The panel
// chain of the next outer block — a sequence of small dependent kernels — runs on the context's high-priority side stream as soon as the
// narrow part is done, underneath the wide update on the main stream: the latency of the chain is hidden instead of being paid per panel
// (the round-1 version ran 79 panels x 4 serial launches at n = 5000: 10 ms at 4 TFLOP/s).
constexpr int NBO = 512;   // outer block

struct StreamSwap {       // run the library's launchers on another stream for the lifetime of this object (host code is single threaded)
    gmb_ctx* c; cudaStream_t saved;
    StreamSwap(gmb_ctx* ctx, cudaStream_t s) : c(ctx), saved(ctx->stream) { c->stream = s; }
    ~StreamSwap() { c->stream = saved; }
};

// storage for the inverted diagonal blocks of large block `bi` (allocated for all large blocks at first use)
static int linv_buffer(gmb_cov* cv, int bi, double** out) {
    gmb_ctx* ctx = cv->ctx;
    if (cv->linv_off.empty()) {
        long long off = 0;
        cv->linv_off.assign(cv->blocks.size(), -1);
        for (size_t k = 0; k < cv->blocks.size(); k++)
            if (cv->blocks[k].n > GMB_COV_SMALL_MAX) { cv->linv_off[k] = off; off += (long long)((cv->blocks[k].n + NB - 1) / NB) * NB * NB; }
        cv->linv_doubles = (size_t)off;
        GMB_CUDA(gmb_dmalloc(ctx, &cv->d_linv, sizeof(double) * (off > 0 ? off : 1)));
    }
    if (cv->linv_off[bi] < 0) return gmb_set_error(GMB_ESTATE, "block %d has no inverse-diagonal storage", bi);
    *out = cv->d_linv + cv->linv_off[bi];
    return GMB_OK;
}
static int x512_buffer(gmb_cov* cv, int bi, double** out) {
    gmb_ctx* ctx = cv->ctx;
    if (cv->x512_off.empty()) {
        long long off = 0;
        cv->x512_off.assign(cv->blocks.size(), -1);
        for (size_t k = 0; k < cv->blocks.size(); k++)
            if (cv->blocks[k].n > GMB_COV_SMALL_MAX) { cv->x512_off[k] = off; off += (long long)((cv->blocks[k].n + NBO - 1) / NBO) * NBO * NBO; }
        GMB_CUDA(gmb_dmalloc(ctx, &cv->d_x512, sizeof(double) * (off > 0 ? off : 1)));
    }
    if (cv->x512_off[bi] < 0) return gmb_set_error(GMB_ESTATE, "block %d has no inverse-diagonal storage", bi);
    *out = cv->d_x512 + cv->x512_off[bi];
    return GMB_OK;
}

// 1 (default) = the diagonal kernel eliminates two columns per hand-off (potrf_diag2_kernel), 0 = one (GMB_POTRF_PAIRS)
static int g_potrf_pairs = [] { const char* e = getenv("GMB_POTRF_PAIRS"); return e ? atoi(e) : 1; }();

// the panel chain of the outer block [K0, Kend), confined to the block's own KB x KB diagonal part: per 128-column panel the diagonal factor +
// inverse, the solve of the (at most 384) rows of the block below it and their rank-128 update (on the current stream).  With X != NULL also
// the inverse of the whole KB x KB factor, block row by block row: X[j, 0:j] = -L_jj^-1 (L[j, 0:j] X[0:j, 0:j]), X[j, j] = L_jj^-1 (T: 128 x
// KB scratch) — off the chain's critical path, on ctx->stream3 behind the event of panel j; ctx->evx is recorded when X is complete.
static int chol_diag_chain(gmb_ctx* ctx, double* A, int ld, int K0, int Kend, int row_offset, int* d_status, double* linv, double* X, double* T) {
    cudaStream_t S = ctx->stream, S3 = ctx->stream3;
    for (int k0 = K0, j = 0; k0 < Kend; k0 += NB, j++) {
        const int kb = Kend - k0 < NB ? Kend - k0 : NB;
        double* Li = linv + (size_t)(k0 / NB) * NB * NB;
        if (g_potrf_pairs) potrf_diag2_kernel<<<1, 256, 0, S>>>(A, ld, k0, kb, row_offset, d_status, Li);
        else potrf_diag_kernel<<<1, 256, 0, S>>>(A, ld, k0, kb, row_offset, d_status, Li);
        ctx->launches++;
        if (X) {
            GMB_CUDA(cudaEventRecord(ctx->evd[j], S));
            StreamSwap sw(ctx, S3);
            GMB_CUDA(cudaStreamWaitEvent(S3, ctx->evd[j], 0));
            const int j0 = k0 - K0;
            GMB_CUDA(cudaMemcpy2DAsync(X + j0 + (size_t)j0 * NBO, sizeof(double) * NBO, Li, sizeof(double) * NB, sizeof(double) * NB, NB,
                                       cudaMemcpyDeviceToDevice, S3));
            if (j0 > 0) {
                GMB_TRY(gmb_dgemm(ctx, 0, 0, kb, j0, j0, 1.0, A + k0 + (size_t)K0 * ld, ld, X, NBO, 0.0, T, NB));
                GMB_TRY(gmb_dgemm(ctx, 0, 0, kb, j0, kb, -1.0, Li, NB, T, NB, 0.0, X + j0, NBO));
            }
        }
        const int rows_in = Kend - (k0 + kb);
        if (rows_in > 0) {
            double* Pp = A + (k0 + kb) + (size_t)k0 * ld;                       // panel solve P <- P L_kk^-T, in place
            GMB_TRY(gmb_dgemm_rowpanel_small(ctx, rows_in, kb, kb, 1.0, Pp, ld, Li, NB, Pp, ld));
            GMB_TRY(gmb_dsyrk_lower_small(ctx, rows_in, kb, Pp, ld, A + (k0 + kb) + (size_t)(k0 + kb) * ld, ld));
        }
    }
    if (X) GMB_CUDA(cudaEventRecord(ctx->evx, S3));
    return GMB_OK;
}

size_t gmb_chol_linv_doubles(int n) { return (size_t)((n + NB - 1) / NB) * NB * NB; }

static int g_chol_reserve = [] { const char* e = getenv("GMB_CHOL_RESERVE_SMS"); return e ? atoi(e) : 4; }();

// in-place lower Cholesky of the n x n device matrix A (lower triangle read; the strict upper triangle of the 128 x 128 diagonal blocks is
// zeroed, the rest of the upper triangle is left as scratch).  linv: ceil(n/128) * 128 * 128 doubles (inverted diagonal blocks), status:
// first non-PD pivot + 1 + row_offset (0 if fine), d_logdet: sum of 2 log L_ii.  Work is issued on ctx->stream and ctx->stream2 and joined
// back on ctx->stream before returning.
//
// Schedule per outer block K (512 columns), main stream M and high-priority side stream P:
//   P: chain(K)   the block's 512 x 512 diagonal part only: 4 x (potrf_diag, solve and update of <= 384 rows) and the inverse X_K of the
//                 512 x 512 factor — a chain of ~20 small dependent launches, but none of them touches more than 512 rows
//   S3: the inverse X_K, block row j behind potrf_diag j of chain(K) (off the chain's critical path)
//   M: solve(K)   W = A[below, K] X_K^T as a product over all rows below (tri = 4 skips the zero k tiles), out of place — the next outer
//                 block's 512 rows first
//   S3: copy W back into A (the factor's panel), under the update
//   M: A[next diagonal block] -= W_top W_top^T (10 tiles), which releases chain(K + 1) on P
//   M: rest of the trailing update, one PERSISTENT launch over all remaining lower tiles on (SMs - reserve) CTAs, so that chain(K + 1)
//      always finds a free SM: the chain is hidden under the update as long as the update is the longer of the two.
int gmb_chol_blocked(gmb_ctx* ctx, double* A, int ld, int n, int row_offset, int* d_status, double* linv, double* d_logdet, double* x512) {
    cudaStream_t M = ctx->stream, P = ctx->stream2;
    const int nouter = (n + NBO - 1) / NBO;
    const bool tma = gmb_gemm_tma_available();
    double *Wp = nullptr, *Xall = nullptr, *T = nullptr;
    const int ldw = ld;
    // x512 != NULL: the caller keeps the inverses of ALL 512 x 512 diagonal blocks (the forward substitution multiplies with them);
    // otherwise only those the factorisation itself needs (all but the last) live in the scratch area
    const int nx = x512 ? nouter : nouter - 1;
    if (nouter > 1 || x512) {
        if (nouter > 1 && !tma) return gmb_set_error(GMB_ECUDA, "the blocked Cholesky needs cuTensorMapEncodeTiled (TMA GEMM) for n > %d", NBO);
        const size_t need = (size_t)ldw * NBO + (x512 ? 0 : (size_t)nx * NBO * NBO) + (size_t)NB * NBO;
        GMB_TRY(gmb_ctx_scratch(ctx, need));
        Wp = ctx->d_scratch; Xall = x512 ? x512 : Wp + (size_t)ldw * NBO; T = Wp + (size_t)ldw * NBO + (x512 ? 0 : (size_t)nx * NBO * NBO);
        GMB_CUDA(cudaMemsetAsync(Xall, 0, sizeof(double) * (size_t)nx * NBO * NBO, M));
    }
    int max_ctas = g_chol_reserve < 0 ? 0 : ctx->sms - g_chol_reserve;          // < 0: one CTA per tile (not persistent)
    if (g_chol_reserve >= 0 && max_ctas < 1) max_ctas = 1;
    GMB_CUDA(cudaEventRecord(ctx->evn, M));                                     // everything issued so far (the block build) precedes the first panel
    for (int K0 = 0, ko = 0; K0 < n; K0 += NBO, ko++) {
        const int Kend = K0 + NBO < n ? K0 + NBO : n, KB = Kend - K0;
        double* X = (ko < nx) ? Xall + (size_t)ko * NBO * NBO : nullptr;
        {
            StreamSwap sw(ctx, P);
            GMB_CUDA(cudaStreamWaitEvent(P, ctx->evn, 0));
            GMB_TRY(chol_diag_chain(ctx, A, ld, K0, Kend, row_offset, d_status, linv, X, T));
            GMB_CUDA(cudaEventRecord(ctx->evp, P));
        }
        GMB_CUDA(cudaStreamWaitEvent(M, ctx->evp, 0));
        if (Kend >= n) break;
        const int Mt = n - Kend;
        double* Pn = A + Kend + (size_t)K0 * ld;
        double* Ct = A + Kend + (size_t)Kend * ld;
        GMB_CUDA(cudaStreamWaitEvent(M, ctx->evx, 0));
        const int nd = Mt < NBO ? Mt : NBO;
        GMB_TRY(gmb_dgemm_rtri(ctx, nd, KB, Pn, ld, X, NBO, Wp, ldw));                          // solve(K), the next outer block's rows first
        GMB_TRY(gmb_dsyrk_lower_small(ctx, nd, KB, Wp, ldw, Ct, ld));                           // the next outer block's diagonal part
        GMB_CUDA(cudaEventRecord(ctx->evn, M));
        if (Mt > nd) GMB_TRY(gmb_dgemm_rtri(ctx, Mt - nd, KB, Pn + nd, ld, X, NBO, Wp + nd, ldw));
        GMB_CUDA(cudaEventRecord(ctx->evj, M));
        GMB_CUDA(cudaStreamWaitEvent(ctx->stream3, ctx->evj, 0));                               // the factor's panel goes back into A under the update
        GMB_CUDA(cudaMemcpy2DAsync(Pn, sizeof(double) * ld, Wp, sizeof(double) * ldw, sizeof(double) * Mt, KB, cudaMemcpyDeviceToDevice, ctx->stream3));
        if (Mt > nd) GMB_TRY(gmb_dsyrk_lower_rest(ctx, Mt, KB, Wp, ldw, Ct, ld, nd, max_ctas)); // everything else, under chain(K + 1)
    }
    if (nouter > 1 || x512) {                                                    // the last panel copy / the last block's inverse
        GMB_CUDA(cudaEventRecord(ctx->evx, ctx->stream3));
        GMB_CUDA(cudaStreamWaitEvent(M, ctx->evx, 0));
    }
    if (d_logdet) {
        logdet_diag_kernel<<<1, 256, 0, ctx->stream>>>(A, ld, n, d_logdet);
        ctx->launches++;
    }
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}

int gmb_cov_factor_large(gmb_cov* cv, int bi) {
    gmb_ctx* ctx = cv->ctx;
    const CovBlock& b = cv->blocks[bi];
    const int n = b.n, ld = gmb_cov_ld(n);
    double* A = cv->d_Lblk + b.l0;
    double* linv = nullptr;
    GMB_TRY(linv_buffer(cv, bi, &linv));
    double* x512 = nullptr;
    GMB_TRY(x512_buffer(cv, bi, &x512));
    dim3 blk(32, 8), grd((n + 31) / 32, (n + 7) / 8);
    build_block_kernel<<<grd, blk, 0, ctx->stream>>>(b, cv->d_fns, cv->d_data, cv->d_theta, A, ld);
    ctx->launches++;
    return gmb_chol_blocked(ctx, A, ld, n, b.start, cv->d_status, linv, cv->d_logdet + bi, x512);
}

// d_partial: 64 doubles (zeroed by the caller); receives partial sums of ||L^{-1} u_j||^2 over the columns
// lower_rhs != 0: the right-hand sides are the columns of an n x n lower-triangular matrix (its own rows: start offset 0) — column j is zero above
// row j and so is its solution, hence outer block K only touches the columns < Kend
int gmb_cov_quad_large(gmb_cov* cv, int bi, const double* dU, int ldu, int ncols, double* d_partial, int lower_rhs) {
    gmb_ctx* ctx = cv->ctx;
    const CovBlock& b = cv->blocks[bi];
    const int n = b.n, ld = gmb_cov_ld(n), ldw = round_up(n, 4);
    const double* A = cv->d_Lblk + b.l0;
    // workspace: a copy of the block's rows of U for a chunk of columns (bounded to ~2 GiB)
    size_t max_cols = ((size_t)1 << 28) / (size_t)ldw;
    if (max_cols < 128) max_cols = 128;
    int chunk = ncols < (int)max_cols ? ncols : (int)max_cols;
    size_t need = (size_t)(ldw + 2 * NBO) * chunk;
    if (need > cv->work_doubles) {
        if (cv->d_work) { GMB_CUDA(cudaStreamSynchronize(ctx->stream)); gmb_dfree(ctx, cv->d_work); cv->d_work = nullptr; }
        GMB_CUDA(gmb_dmalloc(ctx, &cv->d_work, need * sizeof(double)));
        cv->work_doubles = need;
    }
    double* W = cv->d_work;
    double* Y[2] = {W + (size_t)ldw * chunk, W + (size_t)(ldw + NBO) * chunk};       // solved rows of the current outer block, double buffered (ld NBO)
    double* x512 = nullptr;
    GMB_TRY(x512_buffer(cv, bi, &x512));
    int nchunks = (ncols + chunk - 1) / chunk;
    if (nchunks > 64) return gmb_set_error(GMB_EINVAL, "too many column chunks for a large covariance block");
    int slots = 64 / nchunks; if (slots < 1) slots = 1;
    cudaStream_t M = ctx->stream, P = ctx->stream2;
    for (int c = 0; c < nchunks; c++) {
        int c0 = c * chunk, nc = ncols - c0 < chunk ? ncols - c0 : chunk;
        const int s0 = (c * slots) % 64;
        copy_rows_kernel<<<dim3(nc, (ldw + 255) / 256), 256, 0, ctx->stream>>>(dU + (size_t)c0 * ldu, ldu, lower_rhs ? 0 : b.start, n, nc, W, ldw, lower_rhs, c0);
        ctx->launches++;
        // Blocked forward substitution on outer blocks of 512 rows.  The factorisation left the inverse X_K of every 512 x 512 diagonal block of L
        // (gmb_chol_blocked), so a block's rows are solved by ONE triangular product Y_K = X_K W[K, :] (tri = 1 skips the zero k tiles) on the side
        // stream, their sum of squares is taken there, and the rows below receive one rank-512 update W[below, :] -= L[below, K] Y_K on the main
        // stream — the next block's 512 rows first, which releases Y_{K+1} underneath the rest of the update.  Round 1-2a ran four diagonal solves and
        // three small updates per outer block instead of the one product.
        GMB_CUDA(cudaEventRecord(ctx->evn, M));
        for (int K0 = 0, ko = 0; K0 < n; K0 += NBO, ko++) {
            const int Kend = K0 + NBO < n ? K0 + NBO : n, KB = Kend - K0;
            double* Yk = Y[ko & 1];
            // columns of this chunk that are not identically zero in the rows of block K (all of them unless the right-hand sides are triangular)
            const int nck = lower_rhs ? std::max(0, std::min(nc, Kend - c0)) : nc;
            {
                StreamSwap sw(ctx, P);
                GMB_CUDA(cudaStreamWaitEvent(P, ctx->evn, 0));
                if (nck > 0) {
                    GMB_TRY(gmb_dgemm_tri(ctx, 0, 0, KB, nck, KB, 1.0, x512 + (size_t)ko * NBO * NBO, NBO, W + K0, ldw, 0.0, Yk, NBO, 1));
                    sumsq_kernel<<<slots, 256, 0, ctx->stream>>>(Yk, NBO, KB, nck, d_partial + s0);
                    ctx->launches++;
                }
                GMB_CUDA(cudaEventRecord(ctx->evp, P));
            }
            GMB_CUDA(cudaStreamWaitEvent(M, ctx->evp, 0));
            if (Kend >= n) break;
            const int nar = n - Kend < NBO ? n - Kend : NBO;          // W[Kend:Kend+nar, :] -= L[Kend:Kend+nar, K0:Kend] Y_K
            if (nck > 0) GMB_TRY(gmb_dgemm(ctx, 0, 0, nar, nck, KB, -1.0, A + Kend + (size_t)K0 * ld, ld, Yk, NBO, 1.0, W + Kend, ldw));
            GMB_CUDA(cudaEventRecord(ctx->evn, M));
            const int rest = n - Kend - nar;
            if (rest > 0 && nck > 0)
                GMB_TRY(gmb_dgemm(ctx, 0, 0, rest, nck, KB, -1.0, A + Kend + nar + (size_t)K0 * ld, ld, Yk, NBO, 1.0, W + Kend + nar, ldw));
        }
    }
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}

// C_b = chol(U_b U_b^T) of the model's local sample columns for large block bi, cached per (model, sample version); state -1 when the Gram matrix
// is not (numerically) positive definite — fewer samples than rows, degenerate samples — and the caller streams the samples instead.
int gmb_cov_gram_large(gmb_cov* cv, int bi, gmb_model* mdl, const double** C_out, int* ldc_out) {
    gmb_ctx* ctx = cv->ctx;
    const CovBlock& b = cv->blocks[bi];
    const int n = b.n, ld = gmb_cov_ld(n), ncols = mdl->m_local;
    if (cv->gram_large.size() != cv->blocks.size()) cv->gram_large.assign(cv->blocks.size(), gmb_cov::GramLarge());
    if (!(cv->gramL_model == mdl && cv->gramL_version == mdl->u_version && cv->gramL_cols == ncols)) {
        for (auto& g : cv->gram_large) g.state = 0;
        cv->gramL_model = mdl; cv->gramL_version = mdl->u_version; cv->gramL_cols = ncols;
    }
    gmb_cov::GramLarge& g = cv->gram_large[bi];
    *C_out = nullptr; *ldc_out = ld;
    if (g.state < 0) return GMB_OK;
    if (g.state == 0) {
        if ((b.start & 1) || (mdl->ldq & 1)) { g.state = -1; return GMB_OK; }
        if (!g.C) {
            GMB_CUDA(gmb_dmalloc(ctx, &g.C, sizeof(double) * (size_t)ld * n));
            GMB_CUDA(gmb_dmalloc(ctx, &g.linv, sizeof(double) * gmb_chol_linv_doubles(n)));
        }
        GMB_TRY(gmb_dsyrk_lower_set(ctx, n, ncols, mdl->dU + b.start, mdl->ldq, g.C, ld));
        int* d_stat = cv->d_status + 1;                      // second status word: the Gram factorisation's
        GMB_CUDA(cudaMemsetAsync(d_stat, 0, sizeof(int), ctx->stream));
        GMB_TRY(gmb_chol_blocked(ctx, g.C, ld, n, 0, d_stat, g.linv, nullptr, nullptr));
        int* hstat = reinterpret_cast<int*>(ctx->h_pinned + 66);
        GMB_CUDA(cudaMemcpyAsync(hstat, d_stat, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GMB_CUDA(cudaStreamSynchronize(ctx->stream));
        g.state = (*hstat == 0) ? 1 : -1;
        if (g.state < 0) return GMB_OK;
    }
    *C_out = g.C;
    return GMB_OK;
}

void gmb_cov_gram_large_free(gmb_cov* cv) {
    for (auto& g : cv->gram_large) { gmb_dfree(cv->ctx, g.C); gmb_dfree(cv->ctx, g.linv); g = gmb_cov::GramLarge(); }
}
