// Does a DFMA of a half-masked warp (lanes 16-31 exited) issue faster than a full one?  One warp per scheduler (592 warps), 8 independent
// chains per thread.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_halfwarp tools/microbench_halfwarp.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int active_lanes, int iters, double* out, long long* cyc) {
    const int lane = threadIdx.x & 31;
    if (lane >= active_lanes) return;
    double a[8];
    for (int i = 0; i < 8; i++) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double m = 1.0000001, c = 1e-12;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], m, c);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * 32 + lane] = s;
    if (lane == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 32 * 2048); cudaMalloc(&cyc, 8 * 2048);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wps : {1, 2, 3, 4, 8})
        for (int lanes : {32, 16}) {
            const int blocks = 148 * 4 * wps;
            k<<<blocks, 32>>>(lanes, 100, out, cyc); cudaDeviceSynchronize();
            cudaEventRecord(e0); k<<<blocks, 32>>>(lanes, iters, out, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long h[8]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            printf("warps per scheduler %d, active lanes %2d: %.2f cycles per warp-DFMA in block 0; kernel %.3f ms = %.1f TFLOP/s over all warps (32 lanes counted)\n", wps, lanes,
                   (double)h[0] / (8.0 * iters), ms, (double)blocks * 32 * 8.0 * iters * 2 / (ms * 1e-3) / 1e12);
        }
    return 0;
}
