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
constexpr int STG = 3;            // TMA pipeline depth (2 measures the same: the stream is not latency bound)
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
    double wt[128];
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

__device__ __forceinline__ int w2_slab_off(int s) { return 128 * s * (9 - s); }      // slab s keeps row groups 2s..7
__device__ __forceinline__ int w2_off(int c, int m) {
    const int s = m >> 4;
    return w2_slab_off(s) + ((c >> 3) - 2 * s) * 128 + ((m >> 3) & 1) * 64 + (((c & 7) << 2) + (m & 3)) * 2 + ((m >> 2) & 1);
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

// ---------------------------------------------------------------- diagonal block (64 x 64): the sweep of gp_kernel.cuh with 8 tile rows
__device__ __forceinline__ void tile8_factor_all2(Smem2& sm, int p, int buf, int lane) {
    const double* Mp = sm.regionA + (8 * p) * LDM2 + 8 * p;
    double l[8][8];
#pragma unroll
    for (int a = 0; a < 8; a++) {
#pragma unroll
        for (int b2 = 0; b2 <= a / 2; b2++) {
            const double2 v = *reinterpret_cast<const double2*>(Mp + a * LDM2 + 2 * b2);
            l[a][2 * b2] = v.x;
            if (2 * b2 + 1 <= a) l[a][2 * b2 + 1] = v.y;
        }
    }
    double rinv[8];
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const double d = l[j][j];
        if (!(d > 0.0) || !(d < DBL_MAX)) bad = true;
        const double ri = rsqrt_pos(d);
        rinv[j] = ri;
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
        for (int a = 0; a < 8; a++) {
            wt[a * 8 + b] = w[a];
            sm.W[w2_off(8 * p + a, 8 * p + b)] = w[a];
        }
        sm.dinv[8 * p + b] = w[b];
    }
    if (bad && lane == 0) sm.fail = 1;
}

template <int P>
__device__ __forceinline__ void rows_update2(double (&rt)[8][2], const double* __restrict__ bcol, double* myrow, double a0, double a1,
                                             int chi, int q) {
#pragma unroll
    for (int C = P + 2; C < 8; C++) {
        if (C <= chi) {
            const double b0 = bcol[(8 * C) * LDM2 + q], b1 = bcol[(8 * C) * LDM2 + q + 4];
            dmma(rt[C][0], rt[C][1], a0, b0);
            dmma(rt[C][0], rt[C][1], a1, b1);
            if (C == P + 2) *reinterpret_cast<double2*>(myrow + 8 * C + 2 * q) = make_double2(rt[C][0], rt[C][1]);
        }
    }
}

__device__ __forceinline__ void panel_scale2(Smem2& sm, double* row, const double* wt, int R, int p, int g, int q, double& a0, double& a1) {
    const double x0 = row[q], x1 = row[q + 4];
    const double b0 = wt[g * 8 + q], b1 = wt[g * 8 + q + 4];
    double c0 = 0.0, c1 = 0.0;
    dmma(c0, c1, x0, b0);
    dmma(c0, c1, x1, b1);
    __syncwarp();
    *reinterpret_cast<double2*>(row + 2 * q) = make_double2(c0, c1);
    if (R < p) {
        sm.W[w2_off(8 * p + 2 * q, 8 * R + g)] = c0;
        sm.W[w2_off(8 * p + 2 * q + 1, 8 * R + g)] = c1;
    }
    __syncwarp();
    a0 = -row[q]; a1 = -row[q + 4];
}

// warp 0: pivot warp (and tile row 0's scaling); warps 1..7 own tile rows 1..7; ends with a block barrier
__device__ __forceinline__ void diag_factor_rows2(Smem2& sm, int warp, int lane, int nsteps) {
    const int g = lane >> 2, q = lane & 3;
    double* M = sm.regionA;
    if (warp == 0) {
        tile8_factor_all2(sm, 0, 0, lane);
        __syncthreads();
        for (int p = 0; p < nsteps; p++) {
            if (p >= 1) {
                double a0, a1;
                panel_scale2(sm, M + g * LDM2 + 8 * p, sm.wt + 64 * (p & 1), 0, p, g, q, a0, a1);
            }
            __syncthreads();
            if (p == nsteps - 1) break;
            named_sync(1, 64);
            tile8_factor_all2(sm, p + 1, (p + 1) & 1, lane);
            __syncthreads();
        }
        return;
    }
    const int R = warp;
    double* myrow = M + (8 * R + g) * LDM2;
    double rt[8][2];
#pragma unroll
    for (int C = 2; C < 8; C++) {
        if (C <= R) {
            const double2 v = *reinterpret_cast<const double2*>(myrow + 8 * C + 2 * q);
            rt[C][0] = v.x; rt[C][1] = v.y;
        } else { rt[C][0] = 0.0; rt[C][1] = 0.0; }
    }
    __syncthreads();
    for (int p = 0; p < nsteps; p++) {
        const int P0 = 8 * p;
        const double* wt = sm.wt + 64 * (p & 1);
        double a0, a1;
        if (R != p) panel_scale2(sm, myrow + P0, wt, R, p, g, q, a0, a1);
        else { a0 = -wt[q * 8 + g]; a1 = -wt[(q + 4) * 8 + g]; }
        __syncthreads();
        if (p == nsteps - 1) break;
        const int chi = (R <= p) ? 7 : R;
        const double* bcol = M + g * LDM2 + P0;
        {
            double2* cp = reinterpret_cast<double2*>(myrow + P0 + 8 + 2 * q);
            double2 c = *cp;
            const double b0 = bcol[8 * (p + 1) * LDM2 + q], b1 = bcol[8 * (p + 1) * LDM2 + q + 4];
            dmma(c.x, c.y, a0, b0);
            dmma(c.x, c.y, a1, b1);
            *cp = c;
        }
        if (R == p + 1) {
            __syncwarp();
            __threadfence_block();
            named_arrive(1, 64);
            double r0, r1;
            if (p == 0) { r0 = -wt[q * 8 + g]; r1 = -wt[(q + 4) * 8 + g]; }
            else { r0 = -M[g * LDM2 + P0 + q]; r1 = -M[g * LDM2 + P0 + q + 4]; }
            const double* bp = bcol + 8 * (p + 1) * LDM2 + q;
            double* cp = M + g * LDM2 + P0 + 8 + 2 * q;
            for (int C = p + 1; C <= 7; C++, bp += 8 * LDM2, cp += 8) {
                const double b00 = bp[0], b01 = bp[4];
                double2 x = *reinterpret_cast<double2*>(cp);
                dmma(x.x, x.y, r0, b00);
                dmma(x.x, x.y, r1, b01);
                *reinterpret_cast<double2*>(cp) = x;
            }
        } else {
            switch (p) {
                case 0: rows_update2<0>(rt, bcol, myrow, a0, a1, chi, q); break;
                case 1: rows_update2<1>(rt, bcol, myrow, a0, a1, chi, q); break;
                case 2: rows_update2<2>(rt, bcol, myrow, a0, a1, chi, q); break;
                case 3: rows_update2<3>(rt, bcol, myrow, a0, a1, chi, q); break;
                case 4: rows_update2<4>(rt, bcol, myrow, a0, a1, chi, q); break;
                case 5: rows_update2<5>(rt, bcol, myrow, a0, a1, chi, q); break;
                default: break;
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void diag_pad_identity2(Smem2& sm, int c0, int tid) {
    for (int s = 0; s < 4; s++) {
        const int cnt = (8 - 2 * s) * 128;
        double* dst = sm.W + w2_slab_off(s);
        for (int d = tid; d < cnt; d += NT2) {
            const int rgrel = d >> 7, kp = (d >> 6) & 1, ln = (d >> 1) & 31, hh = d & 1;
            const int c = 8 * (2 * s + rgrel) + (ln >> 2);
            const int m = 16 * s + 4 * (2 * kp + hh) + (ln & 3);
            if (c >= c0) dst[d] = (m == c) ? 1.0 : 0.0;
        }
    }
    if (tid >= c0 && tid < BC) sm.dinv[tid] = 1.0;
}

// z_J = W ytilde_J (DMMA matrix-vector product), |z_J|^2, sum log L_rr; warp w owns rows 8w .. 8w+7
__device__ __forceinline__ void diag_solve2(Smem2& sm, int warp, int lane) {
    const int g = lane >> 2, q = lane & 3;
    double za0 = 0.0, za1 = 0.0, zb0 = 0.0, zb1 = 0.0;
    for (int s = 0; s <= (warp >> 1); s++) {
        const double* base = sm.W + w2_slab_off(s) + (warp - 2 * s) * 128 + lane * 2;
        const double2 f0 = *reinterpret_cast<const double2*>(base);
        const double2 f1 = *reinterpret_cast<const double2*>(base + 64);
        const double* yk = sm.yj + 16 * s + q;
        const double y0 = g == 0 ? yk[0] : 0.0, y1 = g == 0 ? yk[4] : 0.0, y2 = g == 0 ? yk[8] : 0.0, y3 = g == 0 ? yk[12] : 0.0;
        dmma(za0, za1, f0.x, y0);
        dmma(zb0, zb1, f0.y, y1);
        dmma(za0, za1, f1.x, y2);
        dmma(zb0, zb1, f1.y, y3);
    }
    const double z = za0 + zb0;
    double z2 = 0.0, lg = 0.0;
    if (q == 0) {
        sm.zs[8 * warp + g] = z;
        z2 = z * z;
        lg = -log(sm.dinv[8 * warp + g]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        z2 += __shfl_xor_sync(0xffffffffu, z2, o);
        lg += __shfl_xor_sync(0xffffffffu, lg, o);
    }
    if (lane == 0) { sm.qpart[warp] = z2; sm.lpart[warp] = lg; }
}

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
    if (tid == 0) {
        for (int s = 0; s < STG; s++) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t it = 0;
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

        const int chunk = prm.prob_chunk[prob];
        const int n = prm.chunk_n[chunk];
        const long long off = prm.chunk_off[chunk];
        const double* __restrict__ t = prm.t + off;
        const double* __restrict__ y = prm.y + off;
        const double* __restrict__ yerr = prm.yerr + off;
        const double* th = prm.theta + 5 * (size_t)prm.prob_theta[prob];
        const int T = (n + BC - 1) / BC;
        const int npad = T * BC;

        if (tid == 0) {
            Hyp h;
            h.A = exp(th[0]); h.L = exp(th[1]); h.G = exp(th[2]); h.S = exp(th[3]); h.P = exp(th[4]);
            h.hl = fmin(0.5 / h.L, DBL_MAX);
            h.thr = DBL_EPSILON * h.L;
            sm.hyp = h;
        }
        __syncthreads();
        const Hyp& h = sm.hyp;
        const double mingap2 = prm.chunk_mingap2[chunk];
        if (!(h.A >= 0.0 && h.A <= DBL_MAX && h.L > 0.0 && h.L <= DBL_MAX && h.G >= 0.0 && h.G <= DBL_MAX &&
              h.S >= 0.0 && h.S <= DBL_MAX && h.P > 0.0 && h.P <= DBL_MAX && mingap2 >= 0.0)) {
            if (tid == 0) prm.prob_out[prob] = -INFINITY;
            continue;
        }
        const bool white_off = mingap2 < h.thr;
        for (int i = tid; i < npad + BC; i += NT2) {            // one block of slack: the last super tile may reach past npad
            double s = 0.0, c = 1.0, dg = 0.0, yy = 0.0;
            const double ti = t[min(i, n - 1)];
            if (i < n) {
                const double hi = ti / h.P;
                const double lo = fma(-hi, h.P, ti) / h.P;
                const double ph = (hi - rint(hi)) + lo;
                sincospi(ph, &s, &c);
                const double e = sqrt(h.S + yerr[i] * yerr[i]);
                dg = e * e;
                yy = y[i];
            }
            vt[i] = ti; vs[i] = s; vc[i] = c; vd[i] = dg; yw[i] = yy;
        }
        __syncthreads();

        double logdet = 0.0, quad = 0.0;
        double acc[4][4][2];
        bool failed = false;

        for (int J = 0; J < T; J++) {
            const double* jrow = lbase + (size_t)(J * (J - 1) / 2) * BLK2;      // blocks (J, 0..J-1), contiguous
            const int niter = 4 * J;
            stage_points2(sm.colv, BC, vt, prm.npad_max, J * BC, BC, tid);
            stage_points2(sm.rowv, SR, vt, prm.npad_max, J * BC, BC, tid);
            __syncthreads();
            // ================= diagonal block (J, J): warps (0,0), (1,0), (1,1) of the upper half of the warp grid
            {
                const int nreal = min(BC, n - J * BC);
                const int ractJ = (nreal + 31) >> 5;
                const bool active = wr < 2 && wr >= wc && wr < ractJ;
                if (tid == 0) prologue2<true>(sm, it, jrow, jrow, jrow, niter, false);
                if (active) synth_tile<EXACT, true, SR, BC>(acc, sm, 32 * wr, 32 * wc, J * BC + 32 * wr, J * BC + 32 * wc, n, vd, white_off, lane);
                accumulate2<true>(sm, acc, it, jrow, jrow, jrow, niter, false, wr, wc, lane, active);
                __syncthreads();
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
                const int nsteps = (nreal + 7) >> 3;
                diag_factor_rows2(sm, warp, lane, nsteps);
                if (nsteps < 8) {
                    diag_pad_identity2(sm, 8 * nsteps, tid);
                    __syncthreads();
                }
                if (sm.fail) { failed = true; break; }
                diag_solve2(sm, warp, lane);
                if (J + 1 < T) stage_points2(sm.rowv, SR, vt, prm.npad_max, (J + 1) * BC, SR, tid);
                fence_async();
                __syncthreads();
                if (tid == 0) {
                    double qs = 0.0, ls = 0.0;
                    for (int w = 0; w < 8; w++) { qs += sm.qpart[w]; ls += sm.lpart[w]; }
                    quad += qs;
                    logdet += ls;
                }
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
                if (rowact) {
                    if (I * BC + SR <= n && !white_off) synth_tile<EXACT, false, SR, BC>(acc, sm, 32 * wr, 32 * wc, I * BC + 32 * wr, J * BC + 32 * wc, n, vd, white_off, lane);
                    else                                synth_tile<EXACT, true, SR, BC>(acc, sm, 32 * wr, 32 * wc, I * BC + 32 * wr, J * BC + 32 * wc, n, vd, white_off, lane);
                }
                accumulate2<false>(sm, acc, it, a0, a1, jrow, niter, two, wr, wc, lane, rowact);
                __syncthreads();                                   // stages drained: regionA becomes the staging block
                if (rowact) store_tile_frag<true>(sm.regionA, acc, 32 * wr, 32 * wc, lane);   // S = -acc, 128-row A slabs
                __syncthreads();
                if (rowact) {
#pragma unroll
                    for (int mi = 0; mi < 4; mi++)
#pragma unroll
                        for (int ni = 0; ni < 4; ni++) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
                    for (int s = 0; s < 2 * wc; s++) {
                        const double* As = sm.regionA + s * SLAB;
                        const double* Bs = sm.W + w2_slab_off(s) - (2 * s) * 128;
#pragma unroll
                        for (int kp = 0; kp < 2; kp++) {
                            double2 a[4], b[4];
#pragma unroll
                            for (int mi = 0; mi < 4; mi++) a[mi] = *reinterpret_cast<const double2*>(As + (4 * wr + mi) * 128 + kp * 64 + lane * 2);
#pragma unroll
                            for (int ni = 0; ni < 4; ni++) b[ni] = *reinterpret_cast<const double2*>(Bs + (4 * wc + ni) * 128 + kp * 64 + lane * 2);
#pragma unroll
                            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                                for (int ni = 0; ni < 4; ni++) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi].x, b[ni].x);
#pragma unroll
                            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                                for (int ni = 0; ni < 4; ni++) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi].y, b[ni].y);
                        }
                    }
#pragma unroll
                    for (int ds = 0; ds < 2; ds++) {               // the two slabs that cross the diagonal of W
                        const int s = 2 * wc + ds;
                        const double* As = sm.regionA + s * SLAB;
                        const double* Bs = sm.W + w2_slab_off(s) - (2 * s) * 128;
#pragma unroll
                        for (int kp = 0; kp < 2; kp++) {
                            double2 a[4], b[4];
#pragma unroll
                            for (int mi = 0; mi < 4; mi++) a[mi] = *reinterpret_cast<const double2*>(As + (4 * wr + mi) * 128 + kp * 64 + lane * 2);
#pragma unroll
                            for (int ni = 0; ni < 4; ni++)
                                if (4 * ds + 2 * kp <= 2 * ni + 1) b[ni] = *reinterpret_cast<const double2*>(Bs + (4 * wc + ni) * 128 + kp * 64 + lane * 2);
#pragma unroll
                            for (int hh = 0; hh < 2; hh++)
#pragma unroll
                                for (int ni = 0; ni < 4; ni++)
                                    if (4 * ds + 2 * kp + hh <= 2 * ni + 1) {
#pragma unroll
                                        for (int mi = 0; mi < 4; mi++)
                                            dmma(acc[mi][ni][0], acc[mi][ni][1], hh ? a[mi].y : a[mi].x, hh ? b[ni].y : b[ni].x);
                                    }
                        }
                    }
                    {
                        double part[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                        for (int ni = 0; ni < 4; ni++) {
                            const double z0 = sm.zs[32 * wc + 8 * ni + 2 * q], z1 = sm.zs[32 * wc + 8 * ni + 2 * q + 1];
#pragma unroll
                            for (int mi = 0; mi < 4; mi++) part[mi] = fma(acc[mi][ni][0], z0, fma(acc[mi][ni][1], z1, part[mi]));
                        }
#pragma unroll
                        for (int mi = 0; mi < 4; mi++) {
                            part[mi] += __shfl_xor_sync(0xffffffffu, part[mi], 1);
                            part[mi] += __shfl_xor_sync(0xffffffffu, part[mi], 2);
                            if (q == 0) sm.red[wc * SR + 32 * wr + 8 * mi + g] = part[mi];
                        }
                    }
                    {
                        const int Ib = I + (wr >> 1);                // this warp's block row
                        store_tile_frag<false, SLAB2>(lbase + ((size_t)(Ib * (Ib - 1) / 2) + J) * BLK2, acc, 32 * (wr & 1), 32 * wc, lane);
                    }
                }
                const int In = I + 2;
                if (In < T) stage_points2(sm.rowv, SR, vt, prm.npad_max, In * BC, SR, tid);
                fence_async();
                __syncthreads();
                if (tid == 0 && In < T)
                    prologue2<false>(sm, it, lbase + (size_t)(In * (In - 1) / 2) * BLK2, lbase + (size_t)((In + 1) * In / 2) * BLK2, jrow, niter, In + 1 < T);
                if (tid < 32 * ract) yw[I * BC + tid] -= sm.red[tid] + sm.red[SR + tid];
            }
            __threadfence();
            fence_async();
            __syncthreads();
        }
        if (tid == 0) {
            double lnl;
            if (failed) lnl = -INFINITY;
            else {
                lnl = -0.5 * quad - 0.5 * (2.0 * logdet + (double)n * 1.8378770664093453);
                if (!(fabs(lnl) <= DBL_MAX)) lnl = -INFINITY;
            }
            prm.prob_out[prob] = lnl;
        }
    }
}

}  // namespace k64
}  // namespace gprot
