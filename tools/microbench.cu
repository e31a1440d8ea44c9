// Microbenchmarks that fix the fp64 roofline denominators on the box (B200, sm_100a).
//   dmma   : mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) issue rate vs warps/SM and ILP
//   dfma   : plain DFMA rate
//   mixed  : DMMA and DFMA warps co-resident (do they share a pipe?)
//   synth  : cost of one covariance entry (sin + exp + div) in entries/s
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <math.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k_dmma(double* out, int iters, long long* cyc) {
    double acc[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; i++) { acc[i][0] = threadIdx.x * 1e-9; acc[i][1] = i; }
    double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) dmma(acc[i][0], acc[i][1], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP>
__global__ void k_dfma(double* out, int iters, long long* cyc) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-9 + i;
    double a = 1.0 + threadIdx.x * 1e-12, b = 1e-13 * threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// warps 0-3 DMMA, warps 4-7 DFMA (one of each per SMSP); each does `iters` x 8
__global__ void k_mixed(double* out, int iters, int mode /*0: dmma only in even warps (odd idle), 1: both, 2: dfma only in odd*/) {
    int w = threadIdx.x >> 5;
    double s = 0;
    if (((w >> 2) & 1) == 0) {   // warps 0-3 DMMA, 4-7 DFMA: every SMSP (w%4) hosts one of each
        if (mode == 0 || mode == 1) {
            double acc[8][2];
#pragma unroll
            for (int i = 0; i < 8; i++) { acc[i][0] = threadIdx.x * 1e-9; acc[i][1] = i; }
            double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
            for (int it = 0; it < iters; it++) {
#pragma unroll
                for (int i = 0; i < 8; i++) dmma(acc[i][0], acc[i][1], a, b);
            }
#pragma unroll
            for (int i = 0; i < 8; i++) s += acc[i][0] + acc[i][1];
        }
    } else {
        if (mode == 1 || mode == 2) {
            double acc[8];
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = threadIdx.x * 1e-9 + i;
            double a = 1.0 + threadIdx.x * 1e-12, b = 1e-13 * threadIdx.x;
            for (int it = 0; it < iters * 8; it++) {   // 8x more DFMA instrs: same nominal pipe time if DMMA=16cyc, DFMA=2cyc
#pragma unroll
                for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], a, b);
            }
#pragma unroll
            for (int i = 0; i < 8; i++) s += acc[i];
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// same warp issues 8 DMMA then nd DFMA per iteration
template <int ND>
__global__ void k_interleave(double* out, int iters) {
    double acc[8][2], f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i][0] = threadIdx.x * 1e-9; acc[i][1] = i; f[i] = i + threadIdx.x * 1e-9; }
    double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12, c = 1e-13 * threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma(acc[i][0], acc[i][1], a, b);
#pragma unroll
        for (int i = 0; i < ND; i++) f[i & 7] = fma(f[i & 7], a, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i][0] + acc[i][1] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// variant: 0 = mirror of george order (div + sin + 2 exp), 1 = div + sin + 1 exp, 2 = reciprocal-mul + sin + 1 exp,
// 3 = angle-addition (no sin) + 1 exp
template <int V>
__global__ void k_synth(const double* __restrict__ t, double* out, int n, double A, double L, double G, double P) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double ti = t[i % n];
    double si, ci;
    sincospi(ti / P, &si, &ci);
    double hl = 0.5 / L, ip = 1.0 / P;
    double s = 0;
    for (int j = 0; j < n; j++) {
        double tj = t[j];
        double d = ti - tj;
        double k;
        if (V == 0) {
            double r2 = d * d / L;
            double sn = sin(M_PI * d / P);
            k = (A * exp(-0.5 * r2)) * exp(-G * sn * sn);
        } else if (V == 1) {
            double sn = sin(M_PI * d / P);
            k = A * exp(-(d * d) * hl - G * sn * sn);
        } else if (V == 2) {
            double sn = sin((M_PI * d) * ip);
            k = A * exp(-(d * d) * hl - G * sn * sn);
        } else {
            double sj, cj;
            sj = t[j + n]; cj = t[j + 2 * n];
            double sn = si * cj - ci * sj;
            k = A * exp(-(d * d) * hl - G * sn * sn);
        }
        s += k;
    }
    out[i] = s;
}

static float timeit(void (*launch)(void*), void* arg, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(arg); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) launch(arg);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

struct Cfg { int grid, block, iters; double* out; long long* cyc; int mode; };

template <int ILP> void run_dmma(Cfg c) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_dmma<ILP><<<c.grid, c.block>>>(c.out, c.iters, c.cyc); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_dmma<ILP><<<c.grid, c.block>>>(c.out, c.iters, c.cyc);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long cyc; CK(cudaMemcpy(&cyc, c.cyc, 8, cudaMemcpyDeviceToHost));
    double n_mma = (double)c.grid * (c.block / 32) * (double)c.iters * ILP;
    double tf = n_mma * 512.0 / (ms * 1e-3) / 1e12;
    int wps = c.block / 32 * (c.grid / 148);
    double cyc_per = (double)cyc / ((double)c.iters * ILP) ;  // per warp
    printf("dmma  grid=%4d block=%4d ILP=%2d  warps/SM=%2d : %8.3f ms  %7.2f TFLOP/s  warp-cycles/DMMA=%.2f  (=> SMSP cycles/DMMA=%.2f) clk=%.0f MHz\n",
           c.grid, c.block, ILP, wps, ms, tf, cyc_per, cyc_per / (wps / 4.0 > 1 ? wps / 4.0 : 1.0), cyc / (ms * 1e-3) / 1e6);
}
template <int ILP> void run_dfma(Cfg c) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_dfma<ILP><<<c.grid, c.block>>>(c.out, c.iters, c.cyc); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_dfma<ILP><<<c.grid, c.block>>>(c.out, c.iters, c.cyc);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long cyc; CK(cudaMemcpy(&cyc, c.cyc, 8, cudaMemcpyDeviceToHost));
    double n = (double)c.grid * c.block * (double)c.iters * ILP;
    double tf = n * 2.0 / (ms * 1e-3) / 1e12;
    int wps = c.block / 32 * (c.grid / 148);
    double cyc_per = (double)cyc / ((double)c.iters * ILP);
    printf("dfma  grid=%4d block=%4d ILP=%2d  warps/SM=%2d : %8.3f ms  %7.2f TFLOP/s  warp-cycles/DFMA=%.2f clk=%.0f MHz\n",
           c.grid, c.block, ILP, wps, ms, tf, cyc_per, cyc / (ms * 1e-3) / 1e6);
}

int main(int argc, char** argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device: %s  SMs=%d  smem/block optin=%zu  L2=%d MB  clock=%d kHz\n", p.name, p.multiProcessorCount,
           p.sharedMemPerBlockOptin, p.l2CacheSize >> 20, p.clockRate);
    double* out; long long* cyc;
    CK(cudaMalloc(&out, 148 * 16 * 1024 * 8)); CK(cudaMalloc(&cyc, 64));
    int it = 20000;
    // DMMA: warps/SM sweep at ILP 8, then ILP sweep at 4 warps/SM (1 per SMSP) and 16 warps/SM
    Cfg c{148, 128, it, out, cyc, 0};
    run_dmma<1>(c); run_dmma<2>(c); run_dmma<4>(c); run_dmma<8>(c); run_dmma<16>(c);
    c.block = 256; run_dmma<8>(c); run_dmma<16>(c);
    c.block = 512; run_dmma<1>(c); run_dmma<4>(c); run_dmma<8>(c); run_dmma<16>(c);
    c.block = 1024; run_dmma<8>(c);
    c.block = 32; run_dmma<1>(c); run_dmma<16>(c);   // single warp: latency & single-warp issue rate
    // long sustained run (~1 s) to see the clock under power cap
    c.block = 512; c.iters = 400000; run_dmma<16>(c); c.iters = it;
    // DFMA
    c.block = 128; run_dfma<1>(c); run_dfma<8>(c);
    c.block = 512; run_dfma<8>(c); c.block = 1024; run_dfma<8>(c);
    c.block = 32; run_dfma<1>(c); run_dfma<16>(c);
    c.block = 512; c.iters = 2000000; run_dfma<8>(c); c.iters = it;
    // mixed
    for (int mode = 0; mode < 3; mode++) {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        k_mixed<<<148, 256>>>(out, it, mode); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); k_mixed<<<148, 256>>>(out, it, mode); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("mixed mode=%d (0 dmma-only, 1 both, 2 dfma-only; 4 dmma warps + 4 dfma warps per SM): %.3f ms\n", mode, ms);
    }
    {
        auto run = [&](int nd) {
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            auto launch = [&]() {
                if (nd == 0) k_interleave<0><<<148, 512>>>(out, it);
                if (nd == 16) k_interleave<16><<<148, 512>>>(out, it);
                if (nd == 32) k_interleave<32><<<148, 512>>>(out, it);
                if (nd == 64) k_interleave<64><<<148, 512>>>(out, it);
            };
            launch(); CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("interleave: 8 DMMA + %2d DFMA per iteration per warp, 16 warps/SM: %.3f ms (DMMA-only pipe time = 8*16 cyc, DFMA-only = %d*2.17 cyc per iteration per warp)\n", nd, ms, nd);
        };
        run(0); run(16); run(32); run(64);
    }
    // synth
    {
        int n = 2048; double* t; CK(cudaMalloc(&t, 3 * n * 8));
        double* h = (double*)malloc(3 * n * 8);
        for (int i = 0; i < n; i++) { h[i] = 0.02043365 * (i * 33 + (i * 7919) % 31); h[i + n] = sinpi(h[i] / 12.3); h[i + 2 * n] = cospi(h[i] / 12.3); }
        CK(cudaMemcpy(t, h, 3 * n * 8, cudaMemcpyHostToDevice));
        int threads = 148 * 1024;
        for (int v = 0; v < 4; v++) {
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            auto launch = [&]() {
                if (v == 0) k_synth<0><<<threads / 256, 256>>>(t, out, n, 2e-6, 1339.0, 0.1, 12.3);
                if (v == 1) k_synth<1><<<threads / 256, 256>>>(t, out, n, 2e-6, 1339.0, 0.1, 12.3);
                if (v == 2) k_synth<2><<<threads / 256, 256>>>(t, out, n, 2e-6, 1339.0, 0.1, 12.3);
                if (v == 3) k_synth<3><<<threads / 256, 256>>>(t, out, n, 2e-6, 1339.0, 0.1, 12.3);
            };
            launch(); CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            double entries = (double)threads * n;
            printf("synth variant %d: %.3f ms  %.1f G entries/s  (N=2048 lower triangle 2.1M entries => %.1f us/eval on the whole GPU)\n",
                   v, ms, entries / ms / 1e6, 2.098e6 / (entries / ms / 1e3) * 1e3 / 1e3);
        }
    }
    return 0;
}
