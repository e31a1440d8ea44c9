// numerics_probe.cu -- what do the fp64 building blocks of gp_kernel.cuh round like on B200?
//   1. mma.sync.m8n8k4.f64 (DMMA): is D = C + sum_k a_k b_k the chain fma(a3,b3, fma(a2,b2, fma(a1,b1, fma(a0,b0,C))))
//      bit for bit?  If not, how far is it from the exactly rounded result (host __float128), and is the error biased?
//   2. rsqrt_pos (MUFU seed + third-order correction): error against 1/sqrt in __float128.
// Build (GPU box or here): nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/numerics_probe tools/numerics_probe.cu -lquadmath
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include <cuda_runtime.h>
#include <quadmath.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double rsqrt_pos(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-(d * y), y, 1.0);
    return fma(y, e * fma(0.375, e, 0.5), y);
}

// one warp per tile: A [8][4] row-major, B [4][8] (k x n) row-major, C/D [8][8]
__global__ void k_dmma(const double* A, const double* B, const double* C, double* D, int ntiles) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= ntiles) return;
    const int g = lane >> 2, q = lane & 3;
    const double a = A[w * 32 + g * 4 + q], b = B[w * 32 + q * 8 + g];
    double d0 = C[w * 64 + g * 8 + 2 * q], d1 = C[w * 64 + g * 8 + 2 * q + 1];
    dmma(d0, d1, a, b);
    D[w * 64 + g * 8 + 2 * q] = d0;
    D[w * 64 + g * 8 + 2 * q + 1] = d1;
}
__global__ void k_rsqrt(const double* x, double* y, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = rsqrt_pos(x[i]);
}

static double ulp_of(double x) { int e; frexp(x, &e); return ldexp(1.0, e - 53); }

int main() {
    const int NT = 1 << 15;
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    std::vector<double> A(NT * 32), B(NT * 32), C(NT * 64), D(NT * 64);
    for (auto& v : A) v = U(rng);
    for (auto& v : B) v = U(rng);
    for (int w = 0; w < NT; w++)
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 8; j++) {
                double s = 0;
                for (int k = 0; k < 4; k++) s += A[w * 32 + i * 4 + k] * B[w * 32 + k * 8 + j];
                // three regimes: generic, cancelling (C ~ -A.B: the Schur complement regime), dominated by C
                const int mode = w % 3;
                C[w * 64 + i * 8 + j] = mode == 0 ? U(rng) : (mode == 1 ? -s * (1.0 + 1e-6 * U(rng)) : 1e3 * U(rng));
            }
    double *dA, *dB, *dC, *dD;
    CK(cudaMalloc(&dA, A.size() * 8)); CK(cudaMalloc(&dB, B.size() * 8)); CK(cudaMalloc(&dC, C.size() * 8)); CK(cudaMalloc(&dD, D.size() * 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice));
    k_dmma<<<NT / 8, 256>>>(dA, dB, dC, dD, NT);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 8, cudaMemcpyDeviceToHost));
    for (int mode = 0; mode < 3; mode++) {
        long n = 0, same_fwd = 0, same_rev = 0, same_exact = 0;
        double sum_err = 0, sum_abs = 0, max_abs = 0, sum_chain_abs = 0;
        for (int w = mode; w < NT; w += 3)
            for (int i = 0; i < 8; i++)
                for (int j = 0; j < 8; j++) {
                    const double c = C[w * 64 + i * 8 + j], d = D[w * 64 + i * 8 + j];
                    double f = c, r = c;
                    __float128 ex = c;
                    for (int k = 0; k < 4; k++) {
                        f = fma(A[w * 32 + i * 4 + k], B[w * 32 + k * 8 + j], f);
                        r = fma(A[w * 32 + i * 4 + 3 - k], B[w * 32 + (3 - k) * 8 + j], r);
                        ex += (__float128)A[w * 32 + i * 4 + k] * (__float128)B[w * 32 + k * 8 + j];
                    }
                    const double exd = (double)ex;
                    // error scale: one ulp of the largest partial magnitude (|c| or the result)
                    const double u = ulp_of(fmax(fabs(c), fabs(exd)));
                    const double err = (double)((__float128)d - ex) / u, cerr = (double)((__float128)f - ex) / u;
                    n++; same_fwd += (d == f); same_rev += (d == r); same_exact += (d == exd);
                    sum_err += err; sum_abs += fabs(err); sum_chain_abs += fabs(cerr);
                    if (fabs(err) > max_abs) max_abs = fabs(err);
                }
        printf("dmma mode %d (%s): %ld entries; == fma chain k=0..3: %.4f  == reversed chain: %.4f  == exactly rounded: %.4f;"
               " error vs exact in ulps of max(|C|,|D|): mean %+.4f  mean|.| %.4f (fma chain %.4f)  max %.3f\n",
               mode, mode == 0 ? "generic" : (mode == 1 ? "cancelling" : "C dominates"), n, (double)same_fwd / n, (double)same_rev / n,
               (double)same_exact / n, sum_err / n, sum_abs / n, sum_chain_abs / n, max_abs);
    }
    // rsqrt_pos
    const int NR = 1 << 20;
    std::vector<double> x(NR), y(NR);
    std::uniform_real_distribution<double> E(-40.0, 10.0);
    for (auto& v : x) v = exp(E(rng));
    double *dx, *dy;
    CK(cudaMalloc(&dx, NR * 8)); CK(cudaMalloc(&dy, NR * 8));
    CK(cudaMemcpy(dx, x.data(), NR * 8, cudaMemcpyHostToDevice));
    k_rsqrt<<<NR / 256, 256>>>(dx, dy, NR);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(y.data(), dy, NR * 8, cudaMemcpyDeviceToHost));
    double mx = 0, sm = 0, sg = 0;
    for (int i = 0; i < NR; i++) {
        const __float128 t = 1.0Q / sqrtq((__float128)x[i]);
        const double e = (double)(((__float128)y[i] - t) / t) / 1.1102230246251565e-16;      // in units of 2^-53
        sm += fabs(e); sg += e;
        if (fabs(e) > mx) mx = fabs(e);
    }
    printf("rsqrt_pos: %d samples, relative error in units of 2^-53: mean %+.4f  mean|.| %.4f  max %.3f\n", NR, sg / NR, sm / NR, mx);
    return 0;
}
