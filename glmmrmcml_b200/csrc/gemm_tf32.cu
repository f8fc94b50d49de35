// gemm_tf32.cu — fp32-mode contraction zd = Z u on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with TMA operand staging and the
// accumulator in tensor memory (TMEM), error-compensated by the 3xTF32 split.
//
// A TF32 product keeps 10 mantissa bits of each operand (relative error 2^-11): far from the 1e-5 the fp32 mode promises.  Each fp32 operand
// is therefore split ONCE into x_hi = tf32(x) and x_lo = tf32(x - x_hi) (split kernels below; Z per model, u per sample matrix), and the
// product is accumulated in fp32 as  A_hi B_hi + A_hi B_lo + A_lo B_hi  — the dropped term A_lo B_lo is ~2^-22 relative.  Three MMAs per k step
// instead of one, still several times the rate of the FP64 DMMA path (tcgen05 has no f64 kind; the fp64 mode runs gemm_tma.cuh).
//
// Kernel (one 128 x 128 output tile per CTA, 192 threads):
//   warp 0, one thread   TMA producer: per k tile of 32 floats four boxes (A_hi, A_lo, B_hi, B_lo; 128 rows x 128 bytes each, 128-byte swizzle)
//                        into a 3-stage ring, completion on the stage's `full` mbarrier
//   warp 1               allocates 256 TMEM columns (two accumulators); one thread issues, per k tile, 4 k-steps x 3 tcgen05.mma (M 128, N 128, K 8) from
//                        shared-memory descriptors and commits the stage's `empty` mbarrier (tcgen05.commit); after every 16 k tiles it commits the
//                        accumulator's `full` mbarrier and switches to the other accumulator
//   warps 2-5            epilogue: per chunk tcgen05.ld their 32-lane quadrant of the accumulator (32 columns at a time), add it to the thread's
//                        row of fp32 sums, release the accumulator; finally store the row as float into the column-major result
// Operands are K-major (k contiguous): B = u is stored that way (Q x m column-major); Z (n x Q column-major) is transposed by its split kernel.
#include "common.cuh"
#include "gemm_tma.cuh"

namespace {

constexpr int TM = 128, TN = 128, TK = 32, TST = 3;
constexpr int OP_BYTES = 128 * 128;                              // one operand tile: 128 rows x 32 floats
constexpr int STAGE_BYTES = 4 * OP_BYTES;                        // A_hi, A_lo, B_hi, B_lo
constexpr size_t TF32_SMEM = (size_t)TST * STAGE_BYTES + 1024 + 128;

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// hi / lo split of a column-major fp64 matrix (rows x cols, ld) into K-major float arrays
//   transpose = 0: out[c * ldo + r]   (the matrix is already k-contiguous: u, Q x m)
//   transpose = 1: out[r * ldo + c]   (Z, n x Q: k = column index becomes contiguous)
__global__ void split_tf32_kernel(int rows, int cols, int ld, const double* __restrict__ src, int transpose, int ldo,
                                  float* __restrict__ hi, float* __restrict__ lo) {
    __shared__ float th[32][33], tl[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
        const int r = r0 + threadIdx.x, c = c0 + cc;
        float h = 0.f, l = 0.f;
        if (r < rows && c < cols) {
            const float x = (float)src[r + (size_t)c * ld];
            h = to_tf32(x); l = to_tf32(x - h);
        }
        if (!transpose) { if (r < ldo && c < cols) { hi[(size_t)c * ldo + r] = h; lo[(size_t)c * ldo + r] = l; } }
        else { th[cc][threadIdx.x] = h; tl[cc][threadIdx.x] = l; }
    }
    if (transpose) {
        __syncthreads();
        for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
            const int r = r0 + rr, c = c0 + threadIdx.x;
            if (r < rows && c < ldo) { hi[(size_t)r * ldo + c] = th[threadIdx.x][rr]; lo[(size_t)r * ldo + c] = tl[threadIdx.x][rr]; }
        }
    }
}

// shared-memory matrix descriptor of a K-major, 128-byte-swizzled operand tile (8-row groups 1024 bytes apart)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(gmbtma::smem_u32(bar)) : "memory");
}

// k tiles accumulated in TMEM before the partial sum is handed to the epilogue.  The tensor core accumulates in fp32 with truncation: the bias grows
// linearly with the number of accumulation steps (measured: relative error of a log-likelihood 1.55e-9 K with one accumulator over the whole of K,
// tools/tf32_error_sweep.py — 1.3e-5 at K = 8192, beyond the fp32 mode's tolerance).  Chunks of 512 are summed by the epilogue with rounded fp32
// additions instead; two TMEM accumulators alternate so that the MMAs of the next chunk run while the epilogue drains the previous one.
constexpr int KCHUNK_TILES = 16;

__global__ void __launch_bounds__(192, 1) sgemm3_tf32_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                                                            const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                                                            int M, int N, int K, float* __restrict__ Cm, int ldc) {
    extern __shared__ unsigned char smraw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)TST * STAGE_BYTES);
    uint64_t* empty = full + TST;
    uint64_t* acc_full = empty + TST;                           // [2]: the chunk's MMAs are done
    uint64_t* acc_empty = acc_full + 2;                         // [2]: the epilogue has drained the accumulator
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const int KT = (K + TK - 1) / TK;
    const int NC = (KT + KCHUNK_TILES - 1) / KCHUNK_TILES;
    if (tid == 0) {
        for (int s = 0; s < TST; s++) { gmbtma::mbar_init(&full[s], 1); gmbtma::mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { gmbtma::mbar_init(&acc_full[b], 1); gmbtma::mbar_init(&acc_empty[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                            // TMEM: two accumulators of 128 columns x 128 lanes of fp32
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gmbtma::smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kt = 0; kt < KT; kt++) {
                const int s = kt % TST;
                if (kt >= TST) gmbtma::mbar_wait(&empty[s], ((kt / TST) - 1) & 1);
                gmbtma::mbar_expect_tx(&full[s], STAGE_BYTES);
                unsigned char* st = base + (size_t)s * STAGE_BYTES;
                gmbtma::tma_load_2d(st, &tmAh, &full[s], kt * TK, m0);
                gmbtma::tma_load_2d(st + OP_BYTES, &tmAl, &full[s], kt * TK, m0);
                gmbtma::tma_load_2d(st + 2 * OP_BYTES, &tmBh, &full[s], kt * TK, n0);
                gmbtma::tma_load_2d(st + 3 * OP_BYTES, &tmBl, &full[s], kt * TK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D fp32, A and B tf32, both K-major, N = 128, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            for (int c = 0; c < NC; c++) {
                const int buf = c & 1;
                if (c >= 2) { gmbtma::mbar_wait(&acc_empty[buf], ((c >> 1) - 1) & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
                const uint32_t tacc = tmem + (uint32_t)(buf * TN);
                const int kt_end = min(KT, (c + 1) * KCHUNK_TILES);
                for (int kt = c * KCHUNK_TILES; kt < kt_end; kt++) {
                    const int s = kt % TST;
                    gmbtma::mbar_wait(&full[s], (kt / TST) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = gmbtma::smem_u32(base + (size_t)s * STAGE_BYTES);
#pragma unroll
                    for (int k8 = 0; k8 < TK / 8; k8++) {       // one instruction covers K = 8 tf32 = 32 bytes along the swizzled row
                        const uint64_t ah = smem_desc(sa + k8 * 32), al = smem_desc(sa + OP_BYTES + k8 * 32);
                        const uint64_t bh = smem_desc(sa + 2 * OP_BYTES + k8 * 32), bl = smem_desc(sa + 3 * OP_BYTES + k8 * 32);
                        mma_tf32(tacc, ah, bh, idesc, (kt > c * KCHUNK_TILES || k8 > 0) ? 1u : 0u);
                        mma_tf32(tacc, ah, bl, idesc, 1u);
                        mma_tf32(tacc, al, bh, idesc, 1u);
                    }
                    mma_commit(&empty[s]);                      // arrives when the MMAs that read this stage are done
                }
                mma_commit(&acc_full[buf]);
            }
        }
    } else {
        // epilogue warps 2..5: TMEM lane quadrant = warp % 4; the thread owns one row of the tile and sums the chunk results
        const int q = warp & 3;
        const int m = m0 + q * 32 + lane;
        float acc[TN];
#pragma unroll
        for (int j = 0; j < TN; j++) acc[j] = 0.f;
        for (int c = 0; c < NC; c++) {
            const int buf = c & 1;
            gmbtma::mbar_wait(&acc_full[buf], (c >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c0 = 0; c0 < TN; c0 += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TN + c0);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                               "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                               "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; j++) acc[c0 + j] += __uint_as_float(v[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) gmbtma::mbar_arrive(&acc_empty[buf]);
        }
        if (m < M) {
#pragma unroll
            for (int j = 0; j < TN; j++) { const int n = n0 + j; if (n < N) Cm[(size_t)n * ldc + m] = acc[j]; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

int make_map_f32(CUtensorMap* tm, const float* ptr, int k_extent, int rows, int ldk) {
    return gmbtma::make_map(tm, ptr, k_extent, rows, ldk, TK, 128, true, true);
}

}  // namespace

// 1 (default) = the fp32 mode forms a dense Z u on the tensor cores (3xTF32); 0 = fp64 DMMA product narrowed to float (GMB_TF32=0)
static int g_tf32 = [] { const char* e = getenv("GMB_TF32"); return e ? atoi(e) : 1; }();
extern "C" int gmb_estep_set_tf32(int on) { g_tf32 = on ? 1 : 0; return GMB_OK; }
bool gmb_tf32_enabled() { return g_tf32 != 0 && gmbtma::get_encode() != nullptr; }

// K-major hi / lo float copies of a column-major fp64 matrix (see split_tf32_kernel); ldo floats per row of the K-major arrays (multiple of 4)
int gmb_split_tf32(gmb_ctx* ctx, int rows, int cols, int ld, const double* src, int transpose, int ldo, float* hi, float* lo) {
    if (rows <= 0 || cols <= 0) return GMB_OK;
    dim3 grid((std::max(rows, transpose ? rows : ldo) + 31) / 32, (std::max(cols, transpose ? ldo : cols) + 31) / 32), block(32, 8);
    split_tf32_kernel<<<grid, block, 0, ctx->stream>>>(rows, cols, ld, src, transpose, ldo, hi, lo);
    ctx->launches++;
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}

// C (M x N float, column-major, ldc) = A B^T with A given as K-major hi / lo (M rows of lda floats), B likewise (N rows of ldb floats)
int gmb_sgemm3_tf32(gmb_ctx* ctx, int M, int N, int K, const float* Ahi, const float* Alo, int lda, const float* Bhi, const float* Blo, int ldb,
                    float* Cm, int ldc) {
    if (M <= 0 || N <= 0) return GMB_OK;
    if ((lda & 3) || (ldb & 3)) return gmb_set_error(GMB_EINVAL, "sgemm3_tf32: leading dimensions must be multiples of 4 floats");
    CUtensorMap a_h, a_l, b_h, b_l;
    GMB_TRY(make_map_f32(&a_h, Ahi, K, M, lda)); GMB_TRY(make_map_f32(&a_l, Alo, K, M, lda));
    GMB_TRY(make_map_f32(&b_h, Bhi, K, N, ldb)); GMB_TRY(make_map_f32(&b_l, Blo, K, N, ldb));
    static bool configured = false;
    if (!configured) { GMB_CUDA(cudaFuncSetAttribute(sgemm3_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TF32_SMEM)); configured = true; }
    dim3 grid((M + TM - 1) / TM, (N + TN - 1) / TN);
    sgemm3_tf32_kernel<<<grid, 192, TF32_SMEM, ctx->stream>>>(a_h, a_l, b_h, b_l, M, N, K, Cm, ldc);
    ctx->launches++;
    GMB_CUDA(cudaGetLastError());
    return GMB_OK;
}
