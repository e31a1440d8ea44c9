// Latency of the diagonal-block sweep (diag_factor_rows) alone on one SM, with per-step time stamps of the pivot warp and
// of the owner of tile row p+1, and a check of L_JJ, the pivot-tile inverses and z_J against a long-double Cholesky.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DDIAG_TS [-DGPROT_...] tools/sweep_bench.cu -o build/sweep_bench
//   ./build/sweep_bench [nsteps128] [nsteps64]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../gprotation_b200/csrc/gp_kernel64.cuh"
using namespace gprot;

// S = lower triangle of a smooth-kernel covariance with a small nugget (strongly correlated neighbours, cond ~ 1e5)
__host__ __device__ inline double s_entry(int a, int b, int rep) {
    const double d = (double)(a - b);
    return (exp(-0.5 * d * d / 36.0) + ((a == b) ? 1e-4 + 1e-6 * rep : 0.0)) * 3.0;
}
__host__ __device__ inline double y_entry(int a) { return sin(0.37 * a) + 0.01 * a; }

template <class G, class SM, int NTH>
__global__ void __launch_bounds__(NTH, 1) k_sweep(long long* out, double* dump, int reps, int nsteps, int other) {
    extern __shared__ __align__(128) unsigned char raw[];
    SM& sm = *reinterpret_cast<SM*>(raw);
    constexpr int NB = 8 * G::NR;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (blockIdx.x > 0 && !other) return;
    for (int i = tid; i < (int)(sizeof(sm.W) / 8); i += NTH) sm.W[i] = 0.0;
    for (int i = tid; i < 64 * G::NR; i += NTH) sm.lpp[i] = 0.0;
    long long tot = 0;
    for (int rep = 0; rep < reps; rep++) {
        __syncthreads();
        for (int i = tid; i < NB * NB; i += NTH) {
            const int a = i / NB, b = i % NB;
            sm.regionA[a * G::LD + b] = (a >= b) ? s_entry(a, b, 0) : 0.0;
        }
        if (tid < NB) sm.yj[tid] = y_entry(tid);
        if (tid == 0) sm.fail = 0;
        __syncthreads();
        const long long t0 = clock64();
        diag_factor_rows<G>(sm, warp, lane, nsteps);
        const long long t1 = clock64();
        if (rep > 0) tot += t1 - t0;
    }
    __syncthreads();
    if (blockIdx.x == 0) {
        const int nw = (int)(sizeof(sm.W) / 8);
        for (int i = tid; i < NB * NB; i += NTH) {           // L below the diagonal tiles from the operator, W_pp on them
            const int c = i / NB, m = i % NB;
            dump[i] = ((c >> 3) >= (m >> 3)) ? sm.W[G::woff(c, m)] : 0.0;
        }
        for (int i = tid; i < 64 * G::NR; i += NTH) dump[NB * NB + i] = sm.lpp[i];
        if (tid < NB) { dump[NB * NB + 64 * G::NR + tid] = sm.zs[tid]; dump[NB * NB + 64 * G::NR + NB + tid] = sm.dinv[tid]; }
        if (tid == 0) { out[0] = tot / (reps - 1); out[1] = sm.fail; out[2] = nw; }
    }
}

template <class G>
static void check(const std::vector<double>& d, int nsteps) {
    constexpr int NB = 8 * G::NR;
    const int n = 8 * nsteps;
    std::vector<long double> L((size_t)n * n, 0.0L), z(n);
    for (int a = 0; a < n; a++)
        for (int b = 0; b <= a; b++) {
            long double s = s_entry(a, b, 0);
            for (int k = 0; k < b; k++) s -= L[a * n + k] * L[b * n + k];
            L[a * n + b] = (a == b) ? sqrtl(s) : s / L[b * n + b];
        }
    for (int a = 0; a < n; a++) {
        long double s = y_entry(a);
        for (int k = 0; k < a; k++) s -= L[a * n + k] * z[k];
        z[a] = s / L[a * n + a];
    }
    double eL = 0, eD = 0, eZ = 0, eW = 0, mL = 0, mZ = 0;
    for (int a = 0; a < n; a++)
        for (int b = 0; b <= a; b++) {
            const double ref = (double)L[a * n + b];
            mL = fmax(mL, fabs(ref));
            if ((a >> 3) > (b >> 3)) eL = fmax(eL, fabs(d[a * NB + b] - ref));
            else eD = fmax(eD, fabs(d[NB * NB + 64 * (a >> 3) + (a & 7) * 8 + (b & 7)] - ref));
        }
    for (int p = 0; p < nsteps; p++)                            // W_pp L_pp = I
        for (int a = 0; a < 8; a++)
            for (int b = 0; b < 8; b++) {
                long double s = 0;
                for (int k = 0; k < 8; k++) s += (long double)d[(8 * p + a) * NB + 8 * p + k] * L[(8 * p + k) * n + 8 * p + b] * (k >= b ? 1 : 0);
                eW = fmax(eW, fabs((double)(s - (a == b ? 1.0L : 0.0L))));
            }
    for (int a = 0; a < n; a++) { eZ = fmax(eZ, fabs(d[NB * NB + 64 * G::NR + a] - (double)z[a])); mZ = fmax(mZ, fabs((double)z[a])); }
    printf("   against long double: L below the diagonal tiles %.2e, L_pp tiles %.2e (max |L| %.2e), |W_pp L_pp - I| %.2e, z %.2e (max |z| %.2e)\n",
           eL, eD, mL, eW, eZ, mZ);
}

template <class G, class SM, int NTH>
static void run(const char* name, int nsteps, int other, int reps = 12) {
    constexpr int NB = 8 * G::NR;
    const int ND = NB * NB + 64 * G::NR + 2 * NB;
    long long* out; double* dump;
    cudaMalloc(&out, 64); cudaMalloc(&dump, ND * 8);
    cudaFuncSetAttribute(k_sweep<G, SM, NTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM));
    k_sweep<G, SM, NTH><<<other ? 148 * (NTH == 256 ? 2 : 1) : 1, NTH, sizeof(SM)>>>(out, dump, reps, nsteps, other);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8]; std::vector<double> d(ND);
    cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost); cudaMemcpy(d.data(), dump, ND * 8, cudaMemcpyDeviceToHost);
    printf("%s nsteps=%d%s: %s, %lld cycles per sweep = %lld per step, fail=%lld\n", name, nsteps, other ? " (every SM busy with the same sweep)" : "",
           cudaGetErrorString(e), h[0], h[0] / nsteps, h[1]);
    check<G>(d, nsteps);
#ifdef DIAG_TS
    if (!other) {
        long long ts[16][16];
        cudaMemcpyFromSymbol(ts, gprot::g_ts, sizeof(ts));
        printf("   step: pivot warp {z done, S1 passed, tile handed over, factor done, next step} | row p+1 {panel start, panel done, S1 passed, arrive}  (cycles from the pivot warp's step start)\n");
        for (int p = 0; p + 1 < nsteps; p++) {
            const long long b = ts[p][0];
            printf("   %2d: %6lld %6lld %6lld %6lld %6lld | %6lld %6lld %6lld %6lld\n", p, ts[p][1] - b, ts[p][2] - b, ts[p][3] - b, ts[p][4] - b,
                   ts[p + 1][0] - b, ts[p][8] - b, ts[p][9] - b, ts[p][10] - b, ts[p][11] - b);
        }
    }
#endif
    cudaFree(out); cudaFree(dump);
}

int main(int argc, char** argv) {
    const int n128 = argc > 1 ? atoi(argv[1]) : 16, n64 = argc > 2 ? atoi(argv[2]) : 8, reps = argc > 3 ? atoi(argv[3]) : 12;
    if (reps != 12) {                                          // long single-CTA runs for ncu's stall sampling
        run<Geo128, Smem, 512>("Geo128 (512 threads)", n128, 0, reps);
        run<k64::Geo64, k64::Smem2, 256>("Geo64 (256 threads)", n64, 0, reps);
        return 0;
    }
    run<Geo128, Smem, 512>("Geo128 (512 threads)", n128, 0);
    run<k64::Geo64, k64::Smem2, 256>("Geo64 (256 threads)", n64, 0);
    run<Geo128, Smem, 512>("Geo128 (512 threads)", n128, 1);
    run<k64::Geo64, k64::Smem2, 256>("Geo64 (256 threads)", n64, 1);
    return 0;
}
