// Microbenchmark: FP64 pipe throughput on B200 (sm_100a) for the instruction mixes the
// glmmrMCML hot path can use: DFMA, DMMA (mma.sync f64) in its four shapes, and the
// libdevice exp/log/div that the family log-likelihood terms need.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_fp64 microbench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

__global__ void k_dfma(double* out, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < ITERS; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void k_dmma884(double* out, double a, double b) {
    double c[4][2];
    for (int j = 0; j < 4; j++) { c[j][0] = threadIdx.x; c[j][1] = j; }
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    double s = 0; for (int j = 0; j < 4; j++) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma1684(double* out, double a, double b) {
    double c[4][4];
    for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) c[j][k] = threadIdx.x + k;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                         : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3]) : "d"(a), "d"(b), "d"(b));
    }
    double s = 0; for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) s += c[j][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma1688(double* out, double a, double b) {
    double c[4][4];
    for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) c[j][k] = threadIdx.x + k;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
    }
    double s = 0; for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) s += c[j][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16816(double* out, double a, double b) {
    double c[4][4];
    for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) c[j][k] = threadIdx.x + k;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(b), "d"(a));
    }
    double s = 0; for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) s += c[j][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// transcendental mixes: elements/s for exp, log, and the naive logistic term
template <int MODE>
__global__ void k_trans(double* out, double a) {
    double x0 = -3.0 + 1e-3 * threadIdx.x, x1 = x0 + 0.5, x2 = x0 + 1.0, x3 = x0 + 1.5;
    double s = 0;
    for (int i = 0; i < ITERS / 8; i++) {
        if (MODE == 0) { s += exp(x0) + exp(x1) + exp(x2) + exp(x3); }
        if (MODE == 1) { s += log(x0 * x0 + 1.5) + log(x1 * x1 + 1.5) + log(x2 * x2 + 1.5) + log(x3 * x3 + 1.5); }
        if (MODE == 2) { s += log(1.0 / (1.0 + exp(-x0))) + log(1.0 / (1.0 + exp(-x1))) + log(1.0 / (1.0 + exp(-x2))) + log(1.0 / (1.0 + exp(-x3))); }
        if (MODE == 3) { s += 1.0 / (x0 * x0 + 1.5) + 1.0 / (x1 * x1 + 1.5) + 1.0 / (x2 * x2 + 1.5) + 1.0 / (x3 * x3 + 1.5); }
        if (MODE == 4) { s += (double)__expf((float)x0) + (double)__expf((float)x1) + (double)__expf((float)x2) + (double)__expf((float)x3); }
        x0 += a; x1 += a; x2 += a; x3 += a;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    const int threads = 256;
    for (int bps : {1, 2, 4, 8}) {
        int blocks = sms * bps;
        double* out; CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
        double nthr = (double)blocks * threads, nwarp = nthr / 32;
        float ms;
        ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        printf("bps %d DFMA        %8.2f TFLOP/s\n", bps, nthr * ITERS * 8 * 2 / ms / 1e9);
        ms = timeit([&] { k_dmma884<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        printf("bps %d DMMA m8n8k4   %8.2f TFLOP/s\n", bps, nwarp * ITERS * 4 * (8.0 * 8 * 4 * 2) / ms / 1e9);
        ms = timeit([&] { k_dmma1684<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        printf("bps %d DMMA m16n8k4  %8.2f TFLOP/s\n", bps, nwarp * ITERS * 4 * (16.0 * 8 * 4 * 2) / ms / 1e9);
        ms = timeit([&] { k_dmma1688<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        printf("bps %d DMMA m16n8k8  %8.2f TFLOP/s\n", bps, nwarp * ITERS * 4 * (16.0 * 8 * 8 * 2) / ms / 1e9);
        ms = timeit([&] { k_dmma16816<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        printf("bps %d DMMA m16n8k16 %8.2f TFLOP/s\n", bps, nwarp * ITERS * 4 * (16.0 * 8 * 16 * 2) / ms / 1e9);
        ms = timeit([&] { k_trans<0><<<blocks, threads>>>(out, 1e-4); });
        printf("bps %d exp(double)   %8.2f Gelem/s\n", bps, nthr * (ITERS / 8) * 4 / ms / 1e6);
        ms = timeit([&] { k_trans<1><<<blocks, threads>>>(out, 1e-4); });
        printf("bps %d log(double)   %8.2f Gelem/s\n", bps, nthr * (ITERS / 8) * 4 / ms / 1e6);
        ms = timeit([&] { k_trans<2><<<blocks, threads>>>(out, 1e-4); });
        printf("bps %d logistic-ll   %8.2f Gelem/s\n", bps, nthr * (ITERS / 8) * 4 / ms / 1e6);
        ms = timeit([&] { k_trans<3><<<blocks, threads>>>(out, 1e-4); });
        printf("bps %d div(double)   %8.2f Gelem/s\n", bps, nthr * (ITERS / 8) * 4 / ms / 1e6);
        ms = timeit([&] { k_trans<4><<<blocks, threads>>>(out, 1e-4); });
        printf("bps %d __expf        %8.2f Gelem/s\n", bps, nthr * (ITERS / 8) * 4 / ms / 1e6);
        cudaFree(out);
    }
    return 0;
}
