// gp_kernel64.cuh -- the throughput variant: TWO problems per SM.
//
// gp_lnlike_kernel (gp_kernel.cuh) gives one problem all 64 K registers and 224 KB of shared memory of an SM, so
// nothing runs under its latency phases (pivot chain of the diagonal block, tile synthesis, stores, barriers): the DMMA
// pipe idles ~12 % of the time.  This kernel halves the footprint -- 256 threads, 64 x 64 blocks, a 128 x 64 super
// tile (two block rows at a time) per step, 104 KB of shared memory -- so that two CTAs are resident per SM and one
// problem's latency phases run under the other's DMMAs.  Same algorithm, same operand layouts (fragment-major 16-k
// slabs, now 64 rows high), same synthesis, same warp-specialised pivot sweep (8 tile rows instead of 16).
// Used for launches with at least two problems per SM; clusters and small batches stay on gp_lnlike_kernel.
#pragma once
#include "gp_kernel.cuh"

namespace gprot {
namespace k64 {

constexpr int BC = 64;            // block edge
constexpr int NT2 = 256;          // 8 warps: 4 (rows of 32) x 2 (columns of 32) over the 128 x 64 super tile
constexpr int SLAB2 = 1024;       // doubles in one 64 x 16 operand slab (8 KB)
constexpr int BLK2 = 4096;        // doubles in one 64 x 64 block (4 slabs)
constexpr int STG = 3;            // TMA pipeline depth (2 measures the same; a second ring of eight one-slab stages for the
                                  // diagonal block, whose stages use 8 of their 24 KB, was measured too: 300-point chunks
                                  // the same, N = 300 +3 % -- the slabs wait for the producer thread, which shares its
                                  // scheduler's issue slots with the synthesis, not for the memory system)
constexpr int STAGE_D = 3 * SLAB2;        // one stage: A slab of block row I, A slab of block row I+1, B slab of block row J
constexpr int LDM2 = 68;          // pitch of the diagonal workspace (== 4 mod 16)
constexpr int REGION2 = STG * STAGE_D;    // 72 KB: stages | 128 x 64 staging block (64 KB) | diagonal workspace (34 KB)
constexpr int W2_DOUBLES = 2560;  // packed lower triangle of the 64 x 64 inverse in B-operand layout
constexpr int SR = 128;           // rows of the super tile

struct Smem2 {
    double regionA[REGION2];
    double W[W2_DOUBLES];
    double red[2 * SR];           // partial sums of L z_J per warp column
    double zs[BC];
    double yj[BC];
    double dinv[BC];
    double lpp[8 * 64];           // 8 x 8 diagonal tiles of L_JJ (refinement of the tile solves)
    double lpart[8];
    double qpart[8];
    double etab[64];
    double colv[3 * BC];          // block column J: t, sin, cos
    double rowv[3 * SR];          // block rows I, I+1: the same
    Hyp hyp;
    unsigned long long full[STG];
    unsigned long long empty[STG];
    int prob;
    int fail;
};

static_assert(2 * (sizeof(Smem2) + 1024) <= 233472, "two CTAs of gp_lnlike64_kernel must fit one SM's 228 KB of shared memory");

__device__ __forceinline__ int w2_slab_off(int s) { return 128 * s * (9 - s); }      // slab s keeps row groups 2s..7
__device__ __forceinline__ int w2_off(int c, int m) {
    const int s = m >> 4;
    return w2_slab_off(s) + ((c >> 3) - 2 * s) * 128 + ((m >> 3) & 1) * 64 + (c & 7) * 8 + (m & 7);
}

// per-point values (t, sin, cos) of `cnt` points starting at point p0
__device__ __forceinline__ void stage_points2(double* dst, int stride, const double* vt, int npad_max, int p0, int cnt, int tid) {
    for (int k = tid; k < 3 * cnt; k += NT2) {
        const int which = k / cnt, i = k - which * cnt;
        dst[which * stride + i] = vt[(size_t)which * npad_max + p0 + i];
    }
}

// ---------------------------------------------------------------- streamed accumulation
template <bool DIAG>
__device__ __forceinline__ void issue_stage2(Smem2& sm, uint32_t it, const double* a0, const double* a1, const double* b, int i, bool two) {
    const uint32_t st = it % STG;
    mbar_wait(&sm.empty[st], ((it / STG) & 1) ^ 1);
    double* dst = sm.regionA + st * STAGE_D;
    mbar_expect_tx(&sm.full[st], (DIAG ? 1 : (two ? 3 : 2)) * SLAB2 * 8);
    tma_load(dst, a0 + (size_t)i * SLAB2, SLAB2 * 8, &sm.full[st]);
    if (!DIAG) {
        if (two) tma_load(dst + SLAB2, a1 + (size_t)i * SLAB2, SLAB2 * 8, &sm.full[st]);
        tma_load(dst + 2 * SLAB2, b + (size_t)i * SLAB2, SLAB2 * 8, &sm.full[st]);
    }
}

template <bool DIAG>
__device__ __forceinline__ void prologue2(Smem2& sm, uint32_t it, const double* a0, const double* a1, const double* b, int niter, bool two) {
    const int np = niter < STG - 1 ? niter : STG - 1;
    for (int i = 0; i < np; i++) issue_stage2<DIAG>(sm, it + i, a0, a1, b, i, two);
}

// warp (wr, wc): rows 32 wr .. 32 wr + 31 of the super tile (wr < 2: block row I, else I+1), columns 32 wc ..
template <bool DIAG>
__device__ __forceinline__ void accumulate2(Smem2& sm, double (&acc)[4][4][2], uint32_t& it, const double* a0, const double* a1,
                                            const double* b, int niter, bool two, int wr, int wc, int lane, bool active) {
    const int tid = threadIdx.x;
    for (int i = 0; i < niter; i++, it++) {
        if (tid == 0 && i + STG - 1 < niter) issue_stage2<DIAG>(sm, it + STG - 1, a0, a1, b, i + STG - 1, two);
        const uint32_t st = it % STG;
        mbar_wait(&sm.full[st], (it / STG) & 1);
        if (active) {
            const double* stage = sm.regionA + st * STAGE_D;
            const double* As = DIAG ? stage : stage + (wr >> 1) * SLAB2;
            const double* Bs = DIAG ? stage : stage + 2 * SLAB2;
            const int rg0 = DIAG ? 4 * wr : 4 * (wr & 1);
#pragma unroll
            for (int kp = 0; kp < 2; kp++) {
                double2 a[4], bb[4];
#pragma unroll
                for (int mi = 0; mi < 4; mi++) a[mi] = *reinterpret_cast<const double2*>(As + (rg0 + mi) * 128 + kp * 64 + lane * 2);
#pragma unroll
                for (int ni = 0; ni < 4; ni++) bb[ni] = *reinterpret_cast<const double2*>(Bs + (4 * wc + ni) * 128 + kp * 64 + lane * 2);
                if (DIAG && wr == wc) {
#pragma unroll
                    for (int mi = 0; mi < 4; mi++)
#pragma unroll
                        for (int ni = 0; ni <= mi; ni++) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi].x, bb[ni].x);
#pragma unroll
                    for (int mi = 0; mi < 4; mi++)
#pragma unroll
                        for (int ni = 0; ni <= mi; ni++) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi].y, bb[ni].y);
                } else {
#pragma unroll
                    for (int mi = 0; mi < 4; mi++)
#pragma unroll
                        for (int ni = 0; ni < 4; ni++) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi].x, bb[ni].x);
#pragma unroll
                    for (int mi = 0; mi < 4; mi++)
#pragma unroll
                        for (int ni = 0; ni < 4; ni++) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi].y, bb[ni].y);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[st]);
    }
}

// ---------------------------------------------------------------- diagonal block (64 x 64)
// the warp-specialised sweep of gp_kernel.cuh (diag_factor_rows / diag_pad_identity / diag_solve) with 8 tile rows
struct Geo64 {
    static constexpr int LD = LDM2;
    static constexpr int NR = 8;
    static constexpr int NTH = NT2;
    __device__ static __forceinline__ int wslab(int s) { return w2_slab_off(s); }
    __device__ static __forceinline__ int woff(int c, int m) { return w2_off(c, m); }
    // tile row of the sweep owned by warp w (with two CTAs per SM a permutation like Geo128's measures no gain)
    __device__ static __forceinline__ int row_of(int w) { return w; }
};

// ---------------------------------------------------------------- the kernel
template <bool EXACT>
__global__ void __launch_bounds__(NT2, 2) gp_lnlike64_kernel(const Params prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem2& sm = *reinterpret_cast<Smem2*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp >> 1, wc = (warp & 1) ^ ((warp >> 2) & 1);   // each scheduler sees one warp of either column
    const int g = lane >> 2, q = lane & 3;

    if (tid < 64) {
        sm.etab[tid] = EXP2_64TH[tid];
        for (int sl = 0; sl < 4; sl++) sm.W[w2_slab_off(sl) + 64 + tid] = 0.0;    // tiles above the diagonal: never written again
    }
    for (int i = tid; i < 8 * 64; i += NT2) sm.lpp[i] = 0.0;          // L_pp tiles: only the lower triangles are ever written
    if (tid == 0) {
        for (int s = 0; s < STG; s++) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t it = 0;
#ifdef GPROT_TIMING
    long long tim_[16] = {0}, tlast_ = clock64();
#endif
    const int slot = (int)blockIdx.x;
    double* const lbase = prm.lscratch + (size_t)slot * prm.lslot_stride;
    double* const vt = prm.vscratch + (size_t)slot * 5 * prm.npad_max;
    double* const vs = vt + prm.npad_max;
    double* const vc = vs + prm.npad_max;
    double* const vd = vc + prm.npad_max;
    double* const yw = vd + prm.npad_max;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            sm.prob = atomicAdd(prm.counter, 1);
            sm.fail = 0;
        }
        __syncthreads();
        if (sm.prob >= prm.nprob) break;
        const int prob = prm.prob_order ? prm.prob_order[sm.prob] : sm.prob;
        TICK(15);

        const int chunk = prm.prob_chunk[prob];
        const int n = prm.chunk_n[chunk];
        const long long off = prm.chunk_off[chunk];
        const double* __restrict__ t = prm.t + off;
        const double* __restrict__ y = prm.y + off;
        const double* __restrict__ yerr = prm.yerr + off;
        const double* th = prm.theta + 5 * (size_t)prm.prob_theta[prob];
        const int T = (n + BC - 1) / BC;
        const int npad = T * BC;

        if (tid == 0) sm.hyp = make_hyper(th);
        __syncthreads();
        const Hyp& h = sm.hyp;
        const double mingap2 = prm.chunk_mingap2[chunk];
        if (!hyper_ok(h, mingap2)) {
            if (tid == 0) prm.prob_out[prob] = -INFINITY;
            continue;
        }
        const bool white_off = mingap2 < h.thr;
        for (int i = tid; i < npad + BC; i += NT2)               // one block of slack: the last super tile may reach past npad
            point_vectors(i, n, t, y, yerr, h, vt, vs, vc, vd, yw);
        __syncthreads();
        TICK(0);

        double logdet = 0.0, quad = 0.0;
        double acc[4][4][2];
        bool failed = false;

        for (int J = 0; J < T; J++) {
            const double* jrow = lbase + (size_t)(J * (J - 1) / 2) * BLK2;      // blocks (J, 0..J-1), contiguous
            const int niter = 4 * J;
            stage_points2(sm.colv, BC, vt, prm.npad_max, J * BC, BC, tid);
            stage_points2(sm.rowv, SR, vt, prm.npad_max, J * BC, BC, tid);
            __syncthreads();
            TICK(12);
            // ================= diagonal block (J, J): warps (0,0), (1,0), (1,1) of the upper half of the warp grid
            {
                const int nreal = min(BC, n - J * BC);
                const int ractJ = (nreal + 31) >> 5;
                const bool active = wr < 2 && wr >= wc && wr < ractJ;
                if (tid == 0) prologue2<true>(sm, it, jrow, jrow, jrow, niter, false);
                if (active) synth_tile<EXACT, true, SR, BC>(acc, sm, 32 * wr, 32 * wc, J * BC + 32 * wr, J * BC + 32 * wc, n, vd, white_off, lane);
                TICK(1);
                accumulate2<true>(sm, acc, it, jrow, jrow, jrow, niter, false, wr, wc, lane, active);
                __syncthreads();
                TICK(2);
                if (wr < 2) {
                    double* M = sm.regionA;
#pragma unroll
                    for (int mi = 0; mi < 4; mi++)
#pragma unroll
                        for (int ni = 0; ni < 4; ni++) {
                            const int r = 32 * wr + 8 * mi + g, c = 32 * wc + 8 * ni + 2 * q;
                            double2 v;
                            v.x = (active && r >= c) ? -acc[mi][ni][0] : ((wr >= ractJ && r == c) ? 1.0 : 0.0);
                            v.y = (active && r >= c + 1) ? -acc[mi][ni][1] : ((wr >= ractJ && r == c + 1) ? 1.0 : 0.0);
                            *reinterpret_cast<double2*>(M + r * LDM2 + c) = v;
                        }
                }
                if (tid < BC) sm.yj[tid] = yw[J * BC + tid];
                __syncthreads();
                TICK(3);
                const int nsteps = (nreal + 7) >> 3;
                diag_factor_rows<Geo64>(sm, warp, lane, nsteps);
                if (nsteps < 8) {
                    diag_pad_identity<Geo64>(sm, 8 * nsteps, tid);
                    __syncthreads();
                }
                TICK(4);
                if (sm.fail) { failed = true; break; }
                diag_sums<Geo64>(sm, warp, lane);
                if (J + 1 < T) stage_points2(sm.rowv, SR, vt, prm.npad_max, (J + 1) * BC, SR, tid);
                fence_async_smem();
                __syncthreads();
                if (tid == 0) {
                    double qs = 0.0, ls = 0.0;
                    for (int w = 0; w < 8; w++) { qs += sm.qpart[w]; ls += sm.lpart[w]; }
                    quad += qs;
                    logdet += ls;
                }
                TICK(5);
            }
            if (tid == 0 && J + 1 < T) {
                const int I = J + 1;
                prologue2<false>(sm, it, lbase + (size_t)(I * (I - 1) / 2) * BLK2, lbase + (size_t)((I + 1) * I / 2) * BLK2, jrow, niter, I + 1 < T);
            }
            // ================= super tiles below the diagonal: block rows (I, I+1)
            for (int I = J + 1; I < T; I += 2) {
                const bool two = I + 1 < T;
                const double* a0 = lbase + (size_t)(I * (I - 1) / 2) * BLK2;
                const double* a1 = lbase + (size_t)((I + 1) * I / 2) * BLK2;
                const int ract = min(4, (n - I * BC + 31) >> 5);            // warp rows that hold real points
                const bool rowact = wr < ract;
                // per-point values (t, sin, cos) of the next super tile and this thread's entry of ytilde: loaded now, used
                // at the end of the super tile (no global-load latency in front of a barrier)
                const int In = I + 2;
                double pv0 = 0.0, pv1 = 0.0, ycur = 0.0;
                if (In < T) {
                    pv0 = vt[(size_t)(tid >> 7) * prm.npad_max + In * BC + (tid & (SR - 1))];                       // t | sin
                    if (tid < SR) pv1 = vt[(size_t)2 * prm.npad_max + In * BC + tid];                              // cos
                }
                if (tid < 32 * ract) ycur = yw[I * BC + tid];
                if (rowact) {
                    if (I * BC + SR <= n && !white_off) synth_tile<EXACT, false, SR, BC>(acc, sm, 32 * wr, 32 * wc, I * BC + 32 * wr, J * BC + 32 * wc, n, vd, white_off, lane);
                    else                                synth_tile<EXACT, true, SR, BC>(acc, sm, 32 * wr, 32 * wc, I * BC + 32 * wr, J * BC + 32 * wc, n, vd, white_off, lane);
                }
                TICK(6);
                accumulate2<false>(sm, acc, it, a0, a1, jrow, niter, two, wr, wc, lane, rowact);
                __syncthreads();                                   // stages drained: regionA becomes the X staging block; sm.rowv is free
                TICK(7);
                if (In < T) {
                    sm.rowv[tid] = pv0;                            // tid < 128: t, else sin (stride SR = 128)
                    if (tid < SR) sm.rowv[2 * SR + tid] = pv1;
                }
                if (rowact) {
                    trsm_inplace<Geo64, 2>(acc, sm.regionA, sm.W, sm.lpp, wr, wc, lane);     // L L_JJ^T = S by substitution, acc: -S -> L
                    lz_partials(acc, sm.zs, sm.red, SR, wr, wc, lane);
                    const int Ib = I + (wr >> 1);                  // this warp's block row
                    store_tile_frag<false, SLAB2>(lbase + ((size_t)(Ib * (Ib - 1) / 2) + J) * BLK2, acc, 32 * (wr & 1), 32 * wc, lane);
                }
                TICK(8);
                fence_async_smem();
                __syncthreads();
                TICK(9);
                if (tid == 0 && In < T)
                    prologue2<false>(sm, it, lbase + (size_t)(In * (In - 1) / 2) * BLK2, lbase + (size_t)((In + 1) * In / 2) * BLK2, jrow, niter, In + 1 < T);
                if (tid < 32 * ract) yw[I * BC + tid] = ycur - (sm.red[tid] + sm.red[SR + tid]);   // ytilde -= L z_J
                TICK(10);
            }
            // blocks of column J are first read (by TMA, the async proxy) in column J+1: one fence per column (fencing per
            // super tile, each warp behind its own stores, was measured: the same on 300-point problems, 3-5 % slower at
            // N = 600 ... 1000)
            __threadfence();
            fence_async();
            __syncthreads();
            TICK(11);
        }
        if (tid == 0) prm.prob_out[prob] = finish_lnlike(quad, logdet, n, failed);
    }
#ifdef GPROT_TIMING
    if (tid == 0 && blockIdx.x == 0 && prm.timing)
        for (int k = 0; k < 16; k++) prm.timing[k] = tim_[k];
#endif
}

}  // namespace k64
}  // namespace gprot
