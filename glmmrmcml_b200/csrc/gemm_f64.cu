// gemm_f64.cu — plain FP64 GEMM entry (alpha/beta epilogue) on the DMMA mainloop of gemm_f64.cuh.
//
// K1 of SURVEY.md §2.2: zd = Z U (mcmlmodel.h:286, hoisted to once per E-step), ZL = Z L (mcmlmodel.h:67,105),
// u = L v (mhmcmc.h:155) and the trailing updates of the blocked Cholesky / TRSM.
//
// C (M x N, ldc) = alpha * op(A) * op(B) + beta * C, all column-major:
//   transA = 0: A is M x K (m contiguous) ; transA = 1: A is K x M (k contiguous)
//   transB = 0: B is K x N (k contiguous) ; transB = 1: B is N x K (n contiguous)
#include "gemm_f64.cuh"

namespace {

struct EpiAxpby {
    static constexpr bool COLSUM = false;
    double alpha, beta;
    double* C;
    int ldc;
    __device__ __forceinline__ bool column_active(int) const { return true; }
    __device__ __forceinline__ void store(int m, int n, double acc) const {
        size_t o = (size_t)n * ldc + m;
        double r = alpha * acc;
        if (beta != 0.0) r += beta * C[o];
        C[o] = r;
    }
    __device__ __forceinline__ double colterm(int, int, double) const { return 0.0; }
    __device__ __forceinline__ void colsum_out(int, int, double) const {}
};

}  // namespace

int gmb_dgemm(gmb_ctx* ctx, int transA, int transB, int M, int N, int K, double alpha, const double* A, int lda,
              const double* B, int ldb, double beta, double* C, int ldc) {
    return gmb_dgemm_tri(ctx, transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, 0);
}

// lower_a != 0: A (before op) is a lower-triangular M x M factor (K = M); the k tiles above (transA = 0) or below (transA = 1) the
// diagonal of each row tile are skipped
int gmb_dgemm_tri(gmb_ctx* ctx, int transA, int transB, int M, int N, int K, double alpha, const double* A, int lda,
                  const double* B, int ldb, double beta, double* C, int ldc, int lower_a) {
    EpiAxpby epi{alpha, beta, C, ldc};
    const int tri = (lower_a && M == K) ? (transA ? 2 : 1) : 0;
    if (tri) {
        if (!transA && !transB) return gmbgemm::dispatch<false, true>(ctx, M, N, K, A, lda, B, ldb, epi, tri);
        if (!transA && transB) return gmbgemm::dispatch<false, false>(ctx, M, N, K, A, lda, B, ldb, epi, tri);
        if (transA && !transB) return gmbgemm::dispatch<true, true>(ctx, M, N, K, A, lda, B, ldb, epi, tri);
        return gmbgemm::dispatch<true, false>(ctx, M, N, K, A, lda, B, ldb, epi, tri);
    }
    // A is k-contiguous when transposed; B is k-contiguous when NOT transposed
    if (!transA && !transB) return gmbgemm::dispatch<false, true>(ctx, M, N, K, A, lda, B, ldb, epi);
    if (!transA && transB) return gmbgemm::dispatch<false, false>(ctx, M, N, K, A, lda, B, ldb, epi);
    if (transA && !transB) return gmbgemm::dispatch<true, true>(ctx, M, N, K, A, lda, B, ldb, epi);
    return gmbgemm::dispatch<true, false>(ctx, M, N, K, A, lda, B, ldb, epi);
}
