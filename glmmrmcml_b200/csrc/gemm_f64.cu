// gemm_f64.cu — plain FP64 GEMM entry (alpha/beta epilogue) on the DMMA mainloop of gemm_f64.cuh.
//
// K1 of SURVEY.md §2.2: zd = Z U (mcmlmodel.h:286, hoisted to once per E-step), ZL = Z L (mcmlmodel.h:67,105),
// u = L v (mhmcmc.h:155) and the trailing updates of the blocked Cholesky / TRSM.
//
// C (M x N, ldc) = alpha * op(A) * op(B) + beta * C, all column-major:
//   transA = 0: A is M x K (m contiguous) ; transA = 1: A is K x M (k contiguous)
//   transB = 0: B is K x N (k contiguous) ; transB = 1: B is N x K (n contiguous)
#include "gemm_tma.cuh"

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
        if (beta == 1.0) { atomicAdd(C + o, r); return; }      // result unused: a RED (no load latency in the epilogue; one CTA owns the element,
        if (beta != 0.0) r += beta * C[o];                     // so the sum is still a single deterministic addition)
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
        if (!transA && !transB) return gmbtma::dispatch<false, true>(ctx, M, N, K, A, lda, B, ldb, epi, tri);
        if (!transA && transB) return gmbtma::dispatch<false, false>(ctx, M, N, K, A, lda, B, ldb, epi, tri);
        if (transA && !transB) return gmbtma::dispatch<true, true>(ctx, M, N, K, A, lda, B, ldb, epi, tri);
        return gmbtma::dispatch<true, false>(ctx, M, N, K, A, lda, B, ldb, epi, tri);
    }
    // A is k-contiguous when transposed; B is k-contiguous when NOT transposed
    if (!transA && !transB) return gmbtma::dispatch<false, true>(ctx, M, N, K, A, lda, B, ldb, epi);
    if (!transA && transB) return gmbtma::dispatch<false, false>(ctx, M, N, K, A, lda, B, ldb, epi);
    if (transA && !transB) return gmbtma::dispatch<true, true>(ctx, M, N, K, A, lda, B, ldb, epi);
    return gmbtma::dispatch<true, false>(ctx, M, N, K, A, lda, B, ldb, epi);
}

// In-place products with a 128 x 128 operand (the inverted diagonal blocks of cov_large.cu): the CTA tile spans the whole 128-wide side, so a
// CTA reads only locations it alone writes, and it writes them after its last read.
//   gmb_dgemm_rowpanel: C (M x 128) = alpha * A (M x 128) * B^T, B stored 128 x 128 (n contiguous), C may alias A
//   gmb_dgemm_colpanel: C (128 x N) = alpha * A (128 x 128) * B (128 x N), C may alias B
int gmb_dgemm_rowpanel(gmb_ctx* ctx, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
    if (N > 128 || M <= 0) return M <= 0 ? GMB_OK : gmb_set_error(GMB_EINVAL, "gmb_dgemm_rowpanel: N must be <= 128");
    EpiAxpby epi{alpha, 0.0, C, ldc};
    return gmbgemm::launch<64, 128, 2, 4, false, false, EpiAxpby>(ctx, M, N, K, A, lda, B, ldb, epi, 0);
}
int gmb_dgemm_colpanel(gmb_ctx* ctx, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
    if (M > 128 || N <= 0) return N <= 0 ? GMB_OK : gmb_set_error(GMB_EINVAL, "gmb_dgemm_colpanel: M must be <= 128");
    EpiAxpby epi{alpha, 0.0, C, ldc};
    return gmbgemm::launch<128, 64, 4, 2, false, true, EpiAxpby>(ctx, M, N, K, A, lda, B, ldb, epi, 0);
}

// C[:, c0:c1) -= P P^T on the lower tiles of the M x M matrix C (P: M x K, m contiguous): the trailing update of the blocked Cholesky
int gmb_dsyrk_lower_sub(gmb_ctx* ctx, int M, int K, const double* Pm, int ldp, double* C, int ldc, int c0, int c1) {
    EpiAxpby epi{-1.0, 1.0, C, ldc};
    return gmbtma::dispatch_syrk_lower(ctx, M, K, Pm, ldp, epi, c0, c1);
}
// lower tiles of C = P P^T (P: M x K, m contiguous): the Gram matrix of a block's samples
int gmb_dsyrk_lower_set(gmb_ctx* ctx, int M, int K, const double* Pm, int ldp, double* C, int ldc) {
    EpiAxpby epi{1.0, 0.0, C, ldc};
    return gmbtma::dispatch_syrk_lower(ctx, M, K, Pm, ldp, epi, 0, M);
}
// the same over ALL lower tiles except those of the leading skip x skip block (skip a multiple of 128), as one persistent launch on at most
// max_ctas CTAs (0: one CTA per tile)
int gmb_dsyrk_lower_rest(gmb_ctx* ctx, int M, int K, const double* Pm, int ldp, double* C, int ldc, int skip, int max_ctas) {
    EpiAxpby epi{-1.0, 1.0, C, ldc};
    return gmbtma::dispatch_syrk_lower(ctx, M, K, Pm, ldp, epi, 0, M, skip / 128, max_ctas);
}
// C (M x N) = A (M x K) * T^T for a lower-triangular N x N matrix T (K = N, column-major, zeros above the diagonal are NOT read):
// the triangular solve X L^T = A as a product with T = L^-1
int gmb_dgemm_rtri(gmb_ctx* ctx, int M, int N, const double* A, int lda, const double* T, int ldt, double* C, int ldc) {
    EpiAxpby epi{1.0, 0.0, C, ldc};
    return gmbtma::dispatch<false, false>(ctx, M, N, N, A, lda, T, ldt, epi, 4);
}
// Small members of the Cholesky panel chain (at most 512 rows): tiles of 16 x 128 / 32 x 32 instead of 64 x 128 / 64 x 64, so that a
// 384 x 128 x 128 product runs on 24-144 SMs for ~3 us instead of on 6-36 SMs for ~15 us — these launches are the chain's critical path.
int gmb_dgemm_rowpanel_small(gmb_ctx* ctx, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
    if (N > 128 || M <= 0) return M <= 0 ? GMB_OK : gmb_set_error(GMB_EINVAL, "gmb_dgemm_rowpanel_small: N must be <= 128");
    EpiAxpby epi{alpha, 0.0, C, ldc};
    return gmbgemm::launch<16, 128, 1, 8, false, false, EpiAxpby>(ctx, M, N, K, A, lda, B, ldb, epi, 0);
}
// C (M x N) = alpha A B^T + beta C, A: M x K and B: N x K, both m / n contiguous
int gmb_dgemm_nt_small(gmb_ctx* ctx, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C, int ldc) {
    if (M <= 0 || N <= 0) return GMB_OK;
    EpiAxpby epi{alpha, beta, C, ldc};
    return gmbgemm::launch<32, 32, 2, 4, false, false, EpiAxpby>(ctx, M, N, K, A, lda, B, ldb, epi, 0);
}
// C -= P P^T on the lower 32 x 32 tiles of the M x M matrix C (M <= 512 or so)
int gmb_dsyrk_lower_small(gmb_ctx* ctx, int M, int K, const double* Pm, int ldp, double* C, int ldc) {
    if (M <= 0) return GMB_OK;
    EpiAxpby epi{-1.0, 1.0, C, ldc};
    return gmbgemm::launch<32, 32, 2, 4, false, false, EpiAxpby>(ctx, M, M, K, Pm, ldp, Pm, ldp, epi, 3);
}
bool gmb_gemm_tma_available() { return gmbtma::gemm_tma_mode() != 0 && gmbtma::get_encode() != nullptr; }
