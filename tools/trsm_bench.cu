// The off-diagonal block epilogue of gp_lnlike_kernel alone on one SM: trsm_inplace (substitution over 8-wide tile columns,
// refined tile solves) + L z partial sums + store of L_IJ, with the time at which every warp finishes each part, against
// the DMMA pipe's bound (1344 DMMAs per scheduler and block = 21.5 k cycles).  Checks X L_JJ^T = S in long double.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/trsm_bench.cu -o build/trsm_bench
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../gprotation_b200/csrc/gp_kernel.cuh"
using namespace gprot;

__host__ __device__ inline double s_entry(int a, int b) {
    const double d = (double)(a - b);
    return (exp(-0.5 * d * d / 36.0) + ((a == b) ? 1e-4 : 0.0)) * 3.0;
}
__host__ __device__ inline double t_entry(int r, int c) { return (double)((r * 37 + c * 11) % 97) / 97.0 - 0.5 + 0.25 * (double)((r + 3 * c) % 13) / 13.0; }

__global__ void __launch_bounds__(NT, 1) k_trsm(long long* out, double* blk, int reps, int what) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& sm = *reinterpret_cast<Smem*>(raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp >> 2, wc = ((warp & 3) - (warp >> 2)) & 3;
    const int g = lane >> 2, q = lane & 3;
    for (int i = tid; i < W_DOUBLES; i += NT) sm.W[i] = 0.0;
    for (int i = tid; i < 16 * 64; i += NT) sm.lpp[i] = 0.0;
    __syncthreads();
    for (int i = tid; i < 128 * 128; i += NT) {
        const int a = i / 128, b = i % 128;
        sm.regionA[a * LDM + b] = (a >= b) ? s_entry(a, b) : 0.0;
    }
    if (tid < 128) sm.yj[tid] = 0.01 * tid;
    if (tid == 0) sm.fail = 0;
    __syncthreads();
    diag_factor_rows<Geo128>(sm, warp, lane, 16);              // the operator of the solves: sm.W, sm.lpp, sm.zs
    __syncthreads();
    double acc[4][4][2];
    long long tot = 0, t_trsm = 0, t_lz = 0, t_st = 0;
    for (int rep = 0; rep < reps; rep++) {
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
            for (int ni = 0; ni < 4; ni++)
#pragma unroll
                for (int e = 0; e < 2; e++) acc[mi][ni][e] = -t_entry(32 * wr + 8 * mi + g, 32 * wc + 8 * ni + 2 * q + e);
        __syncthreads();
        const long long t0 = clock64();
        if (what & 1) trsm_inplace<Geo128, 4>(acc, sm.regionA, sm.W, sm.lpp, wr, wc, lane);
        const long long t1 = clock64();
        if (what & 2) lz_partials(acc, sm.zs, sm.red, BS, wr, wc, lane);
        const long long t2 = clock64();
        if (what & 4) store_tile_frag<false>(blk + (size_t)blockIdx.x * BLKD, acc, 32 * wr, 32 * wc, lane);
        const long long t3 = clock64();
        fence_async_smem();
        __syncthreads();
        const long long t4 = clock64();
        if (rep > 0) { tot += t4 - t0; t_trsm += t1 - t0; t_lz += t2 - t1; t_st += t3 - t2; }
        if (rep == reps - 1 && blockIdx.x == 0 && lane == 0) { out[64 + 2 * warp] = t0; out[65 + 2 * warp] = t4; }
#ifdef TRSM_TS
        if (rep == reps - 1 && blockIdx.x == 0 && lane == 0 && what == 1)
            printf("warp %2d (wr %d wc %d): own solve starts %6lld, solved +%5lld, staged +%4lld, fenced +%4lld\n", warp, wr, wc, g_trsm_ts[warp][0] - t0,
                   g_trsm_ts[warp][1] - g_trsm_ts[warp][0], g_trsm_ts[warp][2] - g_trsm_ts[warp][1], g_trsm_ts[warp][3] - g_trsm_ts[warp][2]);
#endif
    }
    if (blockIdx.x == 0 && lane == 0) {
        out[4 * warp + 0] = tot / (reps - 1); out[4 * warp + 1] = t_trsm / (reps - 1);
        out[4 * warp + 2] = t_lz / (reps - 1); out[4 * warp + 3] = t_st / (reps - 1);
    }
    if (blockIdx.x == 0 && !(what & 4)) store_tile_frag<false>(blk, acc, 32 * wr, 32 * wc, lane);
    if (blockIdx.x == 0)                                       // L_JJ for the check: below the diagonal tiles from W, the tiles from lpp
        for (int i = tid; i < 128 * 128; i += NT) {
            const int c = i / 128, m = i % 128;
            blk[(size_t)gridDim.x * BLKD + i] = ((c >> 3) > (m >> 3)) ? sm.W[w_off(c, m)] : (((c >> 3) == (m >> 3)) ? sm.lpp[64 * (c >> 3) + (c & 7) * 8 + (m & 7)] : 0.0);
        }
}

int main(int argc, char** argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 20;
    long long* out; double* blk;
    cudaMalloc(&out, 128 * 8);
    cudaMalloc(&blk, (size_t)(148 + 1) * BLKD * 8);
    cudaFuncSetAttribute(k_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    for (int what : {7, 1, 3}) {
        for (int grid : {1, 148}) {
            k_trsm<<<grid, NT, sizeof(Smem)>>>(out, blk, reps, what);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[128];
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            long long worst = 0;                                   // a clock read after a barrier is issued before the barrier completes
            for (int w = 0; w < 16; w++) worst = h[4 * w + 1] + h[4 * w + 2] + h[4 * w + 3] > worst ? h[4 * w + 1] + h[4 * w + 2] + h[4 * w + 3] : worst;
            printf("%s%s%s, %d CTA(s): %s; the last warp is done after %lld cycles (DMMA pipe bound 21504); per warp (wr wc): trsm | lz | store\n", (what & 1) ? "trsm" : "", (what & 2) ? "+lz" : "",
                   (what & 4) ? "+store" : "", grid, cudaGetErrorString(e), worst);
            if (grid == 1 && what != 3)
                for (int w = 0; w < 16; w++) printf("   warp %2d (%d %d): %6lld %5lld %5lld\n", w, w >> 2, ((w & 3) - (w >> 2)) & 3, h[4 * w + 1], h[4 * w + 2], h[4 * w + 3]);
        }
    }
    // check of the last run (what = 3 stores after the loop): X L^T = T
    std::vector<double> h((size_t)2 * BLKD);
    cudaMemcpy(h.data(), blk, BLKD * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(h.data() + BLKD, blk + (size_t)148 * BLKD, BLKD * 8, cudaMemcpyDeviceToHost);
    double err = 0, mx = 0;
    for (int r = 0; r < 128; r++)
        for (int c = 0; c < 128; c++) {
            long double s = 0;
            for (int k = 0; k <= c; k++) s += (long double)h[(k >> 4) * SLAB + (r >> 3) * 128 + ((k >> 3) & 1) * 64 + (r & 7) * 8 + (k & 7)] * h[BLKD + c * 128 + k];
            err = fmax(err, fabs((double)(s - (long double)t_entry(r, c))));
        }
    for (int i = 0; i < BLKD; i++) mx = fmax(mx, fabs(h[i]));
    printf("residual |X L^T - T| = %.2e (max |X| %.2e)\n", err, mx);
    return 0;
}
