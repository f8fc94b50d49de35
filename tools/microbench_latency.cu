// Microbenchmark: latency / ILP requirements of the FP64 pipe on B200 at LOW occupancy (1 CTA of 8 warps per SM, the
// situation of the on-chip sampler kernel whose Z L tile fills shared memory).
//   dmma chains : NCH independent accumulators per warp, each a dependent chain of DMMA m8n8k4
//   dfma chains : NCH independent dependent DFMA chains per thread
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_latency microbench_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;

template <int NCH>
__global__ void k_dmma(double* out, double a, double b, long long* cyc) {
    double c[NCH][2];
#pragma unroll
    for (int j = 0; j < NCH; j++) { c[j][0] = threadIdx.x; c[j][1] = j; }
    long long t0 = clock64();
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < NCH; j++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < NCH; j++) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NCH>
__global__ void k_dfma(double* out, double a, double b, long long* cyc) {
    double c[NCH];
#pragma unroll
    for (int j = 0; j < NCH; j++) c[j] = threadIdx.x + j;
    long long t0 = clock64();
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < NCH; j++) c[j] = fma(c[j], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < NCH; j++) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <typename K>
void run(const char* name, K kern, int nch, int warps, int sms, double* out, long long* dcyc, bool mma) {
    kern<<<sms, warps * 32>>>(out, 1.0000001, 1e-9, dcyc);
    cudaDeviceSynchronize();
    kern<<<sms, warps * 32>>>(out, 1.0000001, 1e-9, dcyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dcyc, sizeof c, cudaMemcpyDeviceToHost);
    double per = (double)c / ((double)ITERS * nch);
    // per SM: warps * nch * ITERS instructions in c cycles
    double flop_per_clk_sm = (mma ? 512.0 : 64.0) * warps * nch * ITERS / (double)c;
    printf("%-6s warps/SM %2d chains %2d : %7.2f cycles per instr per warp ; chain latency <= %6.1f cycles ; %6.1f flop/clk/SM (peak 128)\n",
           name, warps, nch, per, per * nch, flop_per_clk_sm);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
    long long* dcyc; cudaMalloc(&dcyc, sizeof(long long));
    for (int warps : {4, 8, 16}) {
        run("DMMA", k_dmma<1>, 1, warps, sms, out, dcyc, true);
        run("DMMA", k_dmma<2>, 2, warps, sms, out, dcyc, true);
        run("DMMA", k_dmma<4>, 4, warps, sms, out, dcyc, true);
        run("DMMA", k_dmma<7>, 7, warps, sms, out, dcyc, true);
        run("DMMA", k_dmma<14>, 14, warps, sms, out, dcyc, true);
        run("DFMA", k_dfma<1>, 1, warps, sms, out, dcyc, false);
        run("DFMA", k_dfma<2>, 2, warps, sms, out, dcyc, false);
        run("DFMA", k_dfma<4>, 4, warps, sms, out, dcyc, false);
        run("DFMA", k_dfma<8>, 8, warps, sms, out, dcyc, false);
    }
    return 0;
}
