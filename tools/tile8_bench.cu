// Latency of the 8x8 diagonal-tile factorisation variants (one warp alone on an SM).
#include <cstdio>
#include "../gprotation_b200/csrc/gp_kernel.cuh"
using namespace gprot;

// variant B: every lane holds the whole lower triangle (no shuffles), rsqrt based
__device__ __forceinline__ void tile8_factor_redundant(Smem& sm, int p, int buf, int lane) {
    double* Mp = sm.regionA + (8 * p) * LDM + 8 * p;
    double l[8][8];
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = 0; b <= a; b++) l[a][b] = Mp[a * LDM + b];
    double rinv[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const double d = l[j][j];
        const double ri = rsqrt_pos(d);
        double ljj = d * ri;
        ljj = fma(fma(-ljj, ljj, d), 0.5 * ri, ljj);
        l[j][j] = ljj; rinv[j] = ri;
#pragma unroll
        for (int a = j + 1; a < 8; a++) l[a][j] *= ri;
#pragma unroll
        for (int a = j + 1; a < 8; a++)
#pragma unroll
            for (int b = j + 1; b <= a; b++) l[a][b] = fma(-l[a][j], l[b][j], l[a][b]);
    }
    const int b = lane & 7;
    double w[8];
#pragma unroll
    for (int a = 0; a < 8; a++) {
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < a; m++) s = fma(l[a][m], w[m], s);
        w[a] = (a == b) ? rinv[a] : ((a > b) ? -rinv[a] * s : 0.0);
    }
    if (lane < 8) {
        double* wt = sm.wt + 64 * buf;
#pragma unroll
        for (int a = 0; a < 8; a++) { wt[a * 8 + b] = w[a]; if (a > b) Mp[b * LDM + a] = w[a]; }
        sm.dinv[8 * p + b] = w[b];
#pragma unroll
        for (int a = 0; a < 8; a++) if (a == b) {
#pragma unroll
            for (int c = 0; c <= a; c++) Mp[a * LDM + c] = l[a][c];
        }
    }
}

// pure chains for calibration
__global__ void k_chain(long long* out) {
    double x = 1.0 + threadIdx.x * 1e-9, y = 1.000001;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; i++) { x = fma(x, y, 1e-9); }
    long long t1 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; i++) { x = __shfl_sync(0xffffffffu, x, (i & 7)); }
    long long t2 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; i++) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + 1.0; }
    long long t3 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; i++) { x = fma(x, y, 1e-9); x = (threadIdx.x & 7) == (i & 7) ? x : y; }   // DFMA -> FSEL -> DFMA
    long long t4 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; i++) { x = rsqrt_pos(x) + 1.0; }
    long long t5 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4; out[5] = (long long)x; }
}

template <int V>
__global__ void __launch_bounds__(32, 1) k_tile8(long long* out, double* sink) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& sm = *reinterpret_cast<Smem*>(raw);
    const int lane = threadIdx.x;
    // SPD 8x8 tile at p = 0: A = I*8 + 0.5^|i-j|
    for (int i = lane; i < 64; i += 32) { int a = i / 8, b = i % 8; sm.regionA[a * LDM + b] = (a == b ? 8.0 : 0.0) + pow(0.5, fabs((double)(a - b))); }
    __syncwarp();
    long long tot = 0;
    for (int rep = 0; rep < 200; rep++) {
        for (int i = lane; i < 64; i += 32) { int a = i / 8, b = i % 8; sm.regionA[a * LDM + b] = (a == b ? 8.0 : 0.0) + pow(0.5, fabs((double)(a - b))) + rep * 1e-3; }
        __syncwarp();
        long long t0 = clock64();
        if (V == 0) tile8_factor(sm, 0, 0, lane); else tile8_factor_redundant(sm, 0, 0, lane);
        __syncwarp();
        long long t1 = clock64();
        tot += t1 - t0;
    }
    if (lane == 0) { out[V] = tot / 200; sink[V] = sm.wt[9] + sm.dinv[3] + sm.regionA[LDM + 1]; }
}

// same, but inside a 512-thread CTA whose other 15 warps wait at the barrier (the situation in gp_lnlike_kernel)
template <int V>
__global__ void __launch_bounds__(512, 1) k_tile8_cta(long long* out, double* sink) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& sm = *reinterpret_cast<Smem*>(raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long tot = 0;
    for (int rep = 0; rep < 200; rep++) {
        for (int i = threadIdx.x; i < 64; i += 512) { int a = i / 8, b = i % 8; sm.regionA[a * LDM + b] = (a == b ? 8.0 : 0.0) + pow(0.5, fabs((double)(a - b))) + rep * 1e-3; }
        __syncthreads();
        long long t0 = clock64();
        if (warp == 0) { if (V == 0) tile8_factor(sm, 0, 0, lane); else tile8_factor_redundant(sm, 0, 0, lane); }
        __syncthreads();
        long long t1 = clock64();
        tot += t1 - t0;
    }
    if (threadIdx.x == 0) { out[V] = tot / 200; sink[V] = sm.wt[9] + sm.dinv[3] + sm.regionA[LDM + 1]; }
}

// the whole diagonal-block sweep (16 micro steps, 512 threads) on an SPD 128x128 block, hot instruction cache
__global__ void __launch_bounds__(512, 1) k_diag(long long* out, double* sink, int reps) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& sm = *reinterpret_cast<Smem*>(raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long tot = 0, first = 0;
    for (int rep = 0; rep < reps; rep++) {
        for (int i = threadIdx.x; i < 128 * 128; i += 512) {
            int a = i / 128, b = i % 128;
            sm.regionA[a * LDM + b] = (a >= b) ? ((a == b ? 4.0 : 0.0) + exp(-0.01 * (a - b) * (a - b)) + rep * 1e-3 * (a == b)) * 1e-6 : 0.0;
        }
        __syncthreads();
        long long t0 = clock64();
        long long dummy = 0;
        diag_factor(sm, warp, lane, nullptr, dummy);
        long long t1 = clock64();
        if (rep == 0) first = t1 - t0; else tot += t1 - t0;
    }
    if (threadIdx.x == 0) { out[0] = first; out[1] = tot / (reps - 1); sink[0] = sm.dinv[77] + sm.regionA[5 * LDM + 100] + sm.fail; }
}

int main() {
    long long* out; double* sink;
    cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 64);
    cudaFuncSetAttribute(k_tile8<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    cudaFuncSetAttribute(k_tile8<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    k_chain<<<1, 32>>>(out + 8);
    k_tile8<0><<<1, 32, sizeof(Smem)>>>(out, sink);
    k_tile8<1><<<1, 32, sizeof(Smem)>>>(out, sink);
    cudaFuncSetAttribute(k_tile8_cta<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    cudaFuncSetAttribute(k_tile8_cta<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    k_tile8_cta<0><<<1, 512, sizeof(Smem)>>>(out + 16, sink + 2);
    k_tile8_cta<1><<<1, 512, sizeof(Smem)>>>(out + 16, sink + 2);
    k_tile8_cta<0><<<148, 512, sizeof(Smem)>>>(out + 24, sink + 4);
    cudaFuncSetAttribute(k_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    k_diag<<<1, 512, sizeof(Smem)>>>(out + 32, sink + 6, 20);
    long long h[64]; double hs[8];
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, 64 * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hs, sink, 64, cudaMemcpyDeviceToHost);
    printf("status %s\n", cudaGetErrorString(e));
    printf("chain per op: DFMA %.1f  SHFL(double) %.1f  MUFU.RSQ64H+DADD %.1f  DFMA+FSEL %.1f  rsqrt_pos+DADD %.1f cycles\n",
           h[8] / 1000.0, h[9] / 1000.0, h[10] / 1000.0, h[11] / 1000.0, h[12] / 1000.0);
    printf("tile8 lane-distributed: %lld cycles   redundant: %lld cycles   (sinks %g %g)\n", h[0], h[1], hs[0], hs[1]);
    printf("inside a 512-thread CTA + barrier: lane-distributed %lld, redundant %lld cycles; 148 CTAs: %lld\n", h[16], h[17], h[24]);
    printf("diag_factor (128x128, 16 steps): first call %lld cycles, warm average %lld cycles (sink %g)\n", h[32], h[33], hs[6]);
    return 0;
}
