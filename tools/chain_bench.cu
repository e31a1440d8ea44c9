// Dependent-chain latencies behind the 8 x 8 pivot-tile factorisation (one warp alone on an SM), and the routine itself.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/chain_bench.cu -o build/chain_bench
#include <cstdio>
#include "../gprotation_b200/csrc/gp_kernel.cuh"
using namespace gprot;

#define CHAIN(name, body)                                                                 \
    __global__ void name(long long* out, double* sink, double x0, double c) {            \
        double x = x0 + threadIdx.x * 1e-9;                                              \
        const long long t0 = clock64();                                                  \
        _Pragma("unroll 1") for (int i = 0; i < 256; i++) {                              \
            _Pragma("unroll") for (int k = 0; k < 8; k++) { body; }                      \
        }                                                                                \
        const long long t1 = clock64();                                                  \
        sink[threadIdx.x] = x;                                                           \
        if (threadIdx.x == 0) out[0] = t1 - t0;                                          \
    }

__device__ __forceinline__ double mufu_rsq64h(double d) {
    double y;
    asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    return y;
}
__device__ __forceinline__ double mufu_rcp64h(double d) {
    double y;
    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    return y;
}
// rsqrt seed through the FP32 unit: exponent handled with integer ops, mantissa in [1, 4) -> float -> MUFU.RSQ -> double
__device__ __forceinline__ double rsqrt_seed_f32(double d) {
    const int hi = __double2hiint(d);
    const int e = ((hi >> 20) - 1023) & ~1;                       // even exponent (d > 0, finite, normal)
    const double m = __hiloint2double(hi - (e << 20), __double2loint(d));      // m in [1, 4)
    const float r = rsqrtf((float)m);                             // in (0.5, 1]
    const double rd = (double)r;
    return __hiloint2double(__double2hiint(rd) - ((e >> 1) << 20), __double2loint(rd));
}
__device__ __forceinline__ double rsqrt_pos_f32(double d) {
    const double y = rsqrt_seed_f32(d);
    const double e = fma(-(d * y), y, 1.0);
    return fma(y, e * fma(0.375, e, 0.5), y);
}

CHAIN(k_dfma, x = fma(x, c, 1e-3))
CHAIN(k_dmul_dfma, x = fma(x * c, c, 1e-3))
CHAIN(k_dadd, x = x + c)
CHAIN(k_mufu_rsq, x = mufu_rsq64h(x) + c)
CHAIN(k_mufu_rcp, x = mufu_rcp64h(x) + c)
CHAIN(k_rsqrt_pos, x = rsqrt_pos(x) + c)
CHAIN(k_rsqrt_f32seed, x = rsqrt_seed_f32(x) + c)
CHAIN(k_rsqrt_pos_f32, x = rsqrt_pos_f32(x) + c)
CHAIN(k_f2f, x = (double)((float)x) + c)
CHAIN(k_sqrt, x = sqrt(x) + c)
CHAIN(k_div, x = c / x + c)

// DMMA dependency through the accumulator (C -> D) and through the A operand (the result of one product is the left
// factor of the next, as in the refined tile solves of the TRSM), ILP independent chains per warp
template <int ILP, bool VIA_A>
__global__ void k_dmma_chain(long long* out, double* sink, double b0) {
    double c0[ILP], c1[ILP], a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c0[i] = 1e-3 * threadIdx.x; c1[i] = 1e-3 * i; a[i] = 1.0 + 1e-6 * i; }
    const double b = b0 + 1e-9 * threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; it++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (VIA_A) { double d0 = 0.0, d1 = 0.0; dmma(d0, d1, c0[i], b); c0[i] = d0; c1[i] = d1; }
                else dmma(c0[i], c1[i], a[i], b);
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
    sink[threadIdx.x] = s;
    if (threadIdx.x == 0) out[0] = t1 - t0;
}

__global__ void __launch_bounds__(32, 1) k_tile8(long long* out, double* dump, int reps) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& sm = *reinterpret_cast<Smem*>(raw);
    const int lane = threadIdx.x;
    long long tot = 0;
    for (int i = lane; i < 64; i += 32) sm.lpp[i] = 0.0;
    for (int rep = 0; rep < reps; rep++) {
        for (int i = lane; i < 64; i += 32) {
            const int a = i / 8, b = i % 8;
            const double d = (double)(a - b);
            sm.regionA[a * LDM + b] = (a >= b) ? exp(-0.5 * d * d / 36.0) + (a == b ? 1e-4 : 0.0) : 0.0;
        }
        __syncwarp();
        const long long t0 = clock64();
        tile8_factor_all<Geo128>(sm, 0, lane);
        __syncwarp();
        const long long t1 = clock64();
        if (rep > 0) tot += t1 - t0;
        else if (lane == 0) out[1] = t1 - t0;
    }
    if (lane == 0) out[0] = tot / (reps - 1);
    for (int i = lane; i < 64; i += 32) { dump[i] = sm.W[w_off(0, 0) + i]; dump[64 + i] = sm.lpp[i]; }
}

int main() {
    long long* out; double* sink;
    cudaMalloc(&out, 64); cudaMalloc(&sink, 4096);
    long long h;
#define RUN(k, x0, c, label) k<<<1, 32>>>(out, sink, x0, c); cudaDeviceSynchronize(); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost); \
    printf("%-34s %7.1f cycles per link\n", label, h / 2048.0);
    RUN(k_dfma, 1.0, 0.999, "DFMA");
    RUN(k_dmul_dfma, 1.0, 0.999, "DMUL + DFMA");
    RUN(k_dadd, 1.0, 1e-3, "DADD");
    RUN(k_mufu_rsq, 2.0, 1.0, "MUFU.RSQ64H + DADD");
    RUN(k_mufu_rcp, 2.0, 1.0, "MUFU.RCP64H + DADD");
    RUN(k_rsqrt_pos, 2.0, 1.0, "rsqrt_pos + DADD");
    RUN(k_rsqrt_f32seed, 2.0, 1.0, "rsqrt seed via FP32 MUFU + DADD");
    RUN(k_rsqrt_pos_f32, 2.0, 1.0, "rsqrt_pos (FP32 seed) + DADD");
    RUN(k_f2f, 2.0, 1.0, "F2F f64->f32->f64 + DADD");
    RUN(k_sqrt, 2.0, 1.0, "sqrt() + DADD");
    RUN(k_div, 2.0, 1.0, "c / x + DADD");
#define RUND(ILP, VIA, label) k_dmma_chain<ILP, VIA><<<1, 32>>>(out, sink, 0.25); cudaDeviceSynchronize(); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost); \
    printf("%-46s %7.1f cycles per DMMA (%d chains per warp)\n", label, h / (1024.0 * ILP), ILP);
    RUND(1, false, "DMMA -> DMMA through the accumulator");
    RUND(4, false, "DMMA -> DMMA through the accumulator");
    RUND(1, true, "DMMA -> DMMA through the A operand");
    RUND(2, true, "DMMA -> DMMA through the A operand");
    RUND(4, true, "DMMA -> DMMA through the A operand");
    RUND(8, true, "DMMA -> DMMA through the A operand");
    cudaFuncSetAttribute(k_tile8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    double* dump; cudaMalloc(&dump, 1024);
    k_tile8<<<1, 32, sizeof(Smem)>>>(out, dump, 20);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    long long hcold; cudaMemcpy(&hcold, out + 1, 8, cudaMemcpyDeviceToHost);
    double hd[128];
    cudaMemcpy(hd, dump, 1024, cudaMemcpyDeviceToHost);
    double err = 0;                                            // |W L - I|
    for (int a = 0; a < 8; a++)
        for (int b = 0; b < 8; b++) {
            double s = 0;
            for (int k = 0; k < 8; k++) s += hd[a * 8 + k] * hd[64 + k * 8 + b];
            err = fmax(err, fabs(s - (a == b)));
        }
    printf("tile8_factor_all alone: %lld cycles warm, %lld the first (cold instruction cache) call (%s), |W L - I| = %.2e\n", h, hcold, cudaGetErrorString(e), err);
    return 0;
}
