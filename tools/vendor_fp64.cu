// vendor_fp64.cu — comparator timings of the vendor FP64 libraries on the same B200 (the on-box competitors of the
// dense kernels K1 / K4 / K5): cuBLAS Dgemm 8192^3, cuSOLVER Dpotrf (n = 5000, 10000), cuBLAS Dtrsm (n = 5000, 10^4 right-hand sides).
// TOOLS ONLY: never linked into, or called by, libglmmrmcml_b200.so.  Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/vendor_fp64 tools/vendor_fp64.cu -lcublas -lcusolver
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

#define CK(x) do { auto e_ = (x); if ((int)e_ != 0) { fprintf(stderr, "%s failed (%d) at line %d\n", #x, (int)e_, __LINE__); return 1; } } while (0)

__global__ void fill_spd(double* A, int n) {          // diagonally dominant symmetric matrix: A_ij = 1/(1 + |i - j|) (+ n on the diagonal)
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n * n) return;
    int i = (int)(e % n), j = (int)(e / n);
    int d = i > j ? i - j : j - i;
    A[e] = 1.0 / (1.0 + d) + (i == j ? (double)n : 0.0);
}
__global__ void fill_val(double* A, size_t cnt) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < cnt) A[e] = 1.0 + 1e-3 * (double)(e % 977);
}

static float median(std::vector<float> v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; }

int main() {
    cublasHandle_t bl; cusolverDnHandle_t so;
    CK(cublasCreate(&bl)); CK(cusolverDnCreate(&so));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    printf("{");
    {   // Dgemm 8192^3
        const int n = 8192; const size_t cnt = (size_t)n * n;
        double *A, *B, *C; CK(cudaMalloc(&A, cnt * 8)); CK(cudaMalloc(&B, cnt * 8)); CK(cudaMalloc(&C, cnt * 8));
        fill_val<<<(unsigned)((cnt + 255) / 256), 256>>>(A, cnt); fill_val<<<(unsigned)((cnt + 255) / 256), 256>>>(B, cnt);
        const double one = 1.0, zero = 0.0;
        std::vector<float> ts;
        for (int r = 0; r < 6; r++) {
            CK(cudaEventRecord(e0));
            CK(cublasDgemm(bl, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n));
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r) ts.push_back(ms);
        }
        const float ms = median(ts);
        printf("\"dgemm_8192\": {\"ms\": %.3f, \"tflops\": %.2f}", ms, 2.0 * n * (double)n * n / ms / 1e9);
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    for (int n : {5000, 10000}) {   // Dpotrf (lower), fresh matrix per repetition
        const size_t cnt = (size_t)n * n;
        double* A; int* info; CK(cudaMalloc(&A, cnt * 8)); CK(cudaMalloc(&info, 4));
        int lwork = 0; CK(cusolverDnDpotrf_bufferSize(so, CUBLAS_FILL_MODE_LOWER, n, A, n, &lwork));
        double* work; CK(cudaMalloc(&work, (size_t)lwork * 8));
        std::vector<float> ts;
        for (int r = 0; r < 4; r++) {
            fill_spd<<<(unsigned)((cnt + 255) / 256), 256>>>(A, n);
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            CK(cusolverDnDpotrf(so, CUBLAS_FILL_MODE_LOWER, n, A, n, work, lwork, info));
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r) ts.push_back(ms);
        }
        int hinfo = -1; CK(cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost));
        const float ms = median(ts);
        printf(", \"dpotrf_%d\": {\"ms\": %.3f, \"tflops\": %.2f, \"info\": %d}", n, ms, (double)n * n * n / 3.0 / ms / 1e9, hinfo);
        if (n == 5000) {          // Dtrsm with the factor: L X = B, 10^4 right-hand sides (the mvn_ll solve of config C5)
            const int m = 10000; double* B; CK(cudaMalloc(&B, (size_t)n * m * 8));
            const double one = 1.0;
            std::vector<float> t2;
            for (int r = 0; r < 4; r++) {
                fill_val<<<(unsigned)(((size_t)n * m + 255) / 256), 256>>>(B, (size_t)n * m);
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0));
                CK(cublasDtrsm(bl, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n, m, &one, A, n, B, n));
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms2; CK(cudaEventElapsedTime(&ms2, e0, e1)); if (r) t2.push_back(ms2);
            }
            const float m2 = median(t2);
            printf(", \"dtrsm_5000x10000\": {\"ms\": %.3f, \"tflops\": %.2f}", m2, (double)n * n * m / m2 / 1e9);
            cudaFree(B);
        }
        cudaFree(A); cudaFree(info); cudaFree(work);
    }
    printf("}\n");
    return 0;
}
