// hmc_sparse_common.cuh — shared by the structure-aware sampler kernels (hmc_sparse.cu, hmc_comp.cu).
#pragma once
#include "common.cuh"

namespace {

enum { SP_EPS = 0, SP_EBAR, SP_H, SP_LLCUR, SP_K0, SP_ACCEPT, SP_TOTSTEPS, SP_LASTPROB, SP_COUNT };   // = CS_* of hmc.cu

// dev_family_resid_w (common.cuh) for N rows at once, written stage by stage across the rows so that their dependent chains overlap (the sparse sampler runs one or
// two warps per scheduler: nothing else hides the ~20 dependent FP64 operations of exp + reciprocal).  Per row: the operations of
// dev_family_resid_w in the same order, hence the same bits.
// Constants come from the constant bank as direct DFMA operands (literals cost two UMOV each per use inside the step loop: 8 % of the
// issued instructions in the first profile of this kernel).
static __constant__ double SP_EXPC[10] = {
    92.33248261689366,              // 0: 64 / ln2
    6755399441055744.0,             // 1: 1.5 * 2^52 (round to nearest integer by addition)
    -0.010830424667801708,          // 2: -ln2/64 high part
    -2.8447437476627285e-11,        // 3: -ln2/64 low part
    1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0,   // 4..8
    0.0};
template <int FL, int N>
__device__ __forceinline__ void dev_family_resid_w_vec(const double (&c)[N], const double (&ys)[N], const double (&eta)[N],
                                                       const double* __restrict__ tab, double (&out)[N]) {
    if (FL == 7) {
#pragma unroll
        for (int k = 0; k < N; k++) out[k] = fma(-c[k], eta[k], ys[k]);
        return;
    }
    const double* cc = SP_EXPC;
    double t[N], r[N], T[N], q[N], e[N];
    int kk[N];
#pragma unroll
    for (int k = 0; k < N; k++) t[k] = fma(eta[k], cc[0], cc[1]);
#pragma unroll
    for (int k = 0; k < N; k++) { kk[k] = __double2loint(t[k]); t[k] -= cc[1]; }
#pragma unroll
    for (int k = 0; k < N; k++) { T[k] = tab[kk[k] & 63]; r[k] = fma(t[k], cc[2], eta[k]); }
#pragma unroll
    for (int k = 0; k < N; k++) r[k] = fma(t[k], cc[3], r[k]);
#pragma unroll
    for (int k = 0; k < N; k++) q[k] = fma(r[k], cc[4], cc[5]);
#pragma unroll
    for (int k = 0; k < N; k++) q[k] = fma(q[k], r[k], cc[6]);
#pragma unroll
    for (int k = 0; k < N; k++) q[k] = fma(q[k], r[k], cc[7]);
#pragma unroll
    for (int k = 0; k < N; k++) q[k] = fma(q[k], r[k], cc[8]);
#pragma unroll
    for (int k = 0; k < N; k++) q[k] = q[k] * r[k];
#pragma unroll
    for (int k = 0; k < N; k++) {
        const double m = fma(T[k], q[k], T[k]);
        const int ks = min(max(kk[k], -64512), 64512);
        e[k] = __hiloint2double(__double2hiint(m) + ((ks >> 6) << 20), __double2loint(m));
    }
    if (FL == 1) {
#pragma unroll
        for (int k = 0; k < N; k++) out[k] = fma(-c[k], e[k], ys[k]);
        return;
    }
    double d[N], y[N], f[N];
#pragma unroll
    for (int k = 0; k < N; k++) { d[k] = e[k] + cc[8]; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y[k]) : "d"(d[k])); }
#pragma unroll
    for (int k = 0; k < N; k++) f[k] = fma(-d[k], y[k], cc[8]);
#pragma unroll
    for (int k = 0; k < N; k++) y[k] = fma(y[k], f[k], y[k]);
#pragma unroll
    for (int k = 0; k < N; k++) f[k] = fma(-d[k], y[k], cc[8]);
#pragma unroll
    for (int k = 0; k < N; k++) y[k] = fma(y[k], f[k], y[k]);
#pragma unroll
    for (int k = 0; k < N; k++) out[k] = fma(c[k], y[k], ys[k]);
}


// The same residual with a 16-entry table of 2^(j/16) (128 bytes: every lookup pattern is bank-conflict free; random lookups in the 64-entry
// table cost ~5 wavefronts each and were a quarter of the register kernel's shared-memory traffic) and a degree-6 polynomial for
// (exp(r) - 1) / r on |r| <= ln2/32 (truncation r^8/40320 < 1.3e-18), evaluated in Estrin form: the dependent depth stays that of the
// 64-entry variant's degree-4 Horner chain.  Error within 1.1 ulp of the exact value; rounding differs from dev_family_resid_w in the last bit.
static __constant__ double SP_EXPC16[12] = {
    23.083120654223414,             // 0: 16 / ln2
    6755399441055744.0,             // 1: 1.5 * 2^52
    -0.04332169867120683,           // 2: -ln2/16 high part (24 trailing zero bits)
    -1.1378974990650914e-10,        // 3: -ln2/16 low part
    1.0 / 5040.0, 1.0 / 720.0,      // 4, 5
    1.0 / 120.0, 1.0 / 24.0,        // 6, 7
    1.0 / 6.0, 0.5,                 // 8, 9
    1.0, 0.0};
template <int FL, int N>
__device__ __forceinline__ void dev_family_resid_w_vec16(const double (&c)[N], const double (&ys)[N], const double (&eta)[N],
                                                         const double* __restrict__ tab16, double (&out)[N]) {
    if (FL == 7) {
#pragma unroll
        for (int k = 0; k < N; k++) out[k] = fma(-c[k], eta[k], ys[k]);
        return;
    }
    const double* cc = SP_EXPC16;
    double t[N], r[N], r2[N], T[N], a[N], b[N], d3[N], q[N], e[N];
    int kk[N];
#pragma unroll
    for (int k = 0; k < N; k++) t[k] = fma(eta[k], cc[0], cc[1]);
#pragma unroll
    for (int k = 0; k < N; k++) { kk[k] = __double2loint(t[k]); t[k] -= cc[1]; }
#pragma unroll
    for (int k = 0; k < N; k++) { T[k] = tab16[kk[k] & 15]; r[k] = fma(t[k], cc[2], eta[k]); }
#pragma unroll
    for (int k = 0; k < N; k++) r[k] = fma(t[k], cc[3], r[k]);
#pragma unroll
    for (int k = 0; k < N; k++) { r2[k] = r[k] * r[k]; a[k] = fma(r[k], cc[4], cc[5]); b[k] = fma(r[k], cc[6], cc[7]); d3[k] = fma(r[k], cc[8], cc[9]); }
#pragma unroll
    for (int k = 0; k < N; k++) q[k] = fma(a[k], r2[k], b[k]);
#pragma unroll
    for (int k = 0; k < N; k++) q[k] = fma(q[k], r2[k], d3[k]);
#pragma unroll
    for (int k = 0; k < N; k++) q[k] = fma(q[k], r[k], cc[10]);
#pragma unroll
    for (int k = 0; k < N; k++) q[k] = q[k] * r[k];                                                // exp(r) - 1
#pragma unroll
    for (int k = 0; k < N; k++) {
        const double m = fma(T[k], q[k], T[k]);
        const int ks = min(max(kk[k], -16128), 16128);
        e[k] = __hiloint2double(__double2hiint(m) + ((ks >> 4) << 20), __double2loint(m));
    }
    if (FL == 1) {
#pragma unroll
        for (int k = 0; k < N; k++) out[k] = fma(-c[k], e[k], ys[k]);
        return;
    }
    // reciprocal of d = 1 + e^eta >= 1: hardware seed (~20 bits) and ONE third-order step, y (1 + f + f^2) with f = 1 - d y — error ~ f^3, below
    // one ulp, in three dependent FMAs instead of the four of two Newton steps
    double d[N], y[N], f[N];
#pragma unroll
    for (int k = 0; k < N; k++) { d[k] = e[k] + cc[10]; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y[k]) : "d"(d[k])); }
#pragma unroll
    for (int k = 0; k < N; k++) f[k] = fma(-d[k], y[k], cc[10]);
#pragma unroll
    for (int k = 0; k < N; k++) f[k] = fma(f[k], f[k], f[k]);
#pragma unroll
    for (int k = 0; k < N; k++) y[k] = fma(y[k], f[k], y[k]);
#pragma unroll
    for (int k = 0; k < N; k++) out[k] = fma(c[k], y[k], ys[k]);
}

}  // namespace
