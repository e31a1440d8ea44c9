// gp_kernel.cuh -- fused covariance synthesis + blocked fp64 Cholesky + forward solve, sm_100a.
//
// One persistent CTA (512 threads, 16 warps) per SM -- or a thread-block cluster of 2 / 4 / 8 CTAs per problem when a
// launch has fewer problems than SMs -- pulls (walker, chunk) problems from a global queue.  For a problem of n points
// (padded with identity to Npad = T*128):
//
//   for block column J = 0..T-1                         (left-looking, 128 x 128 blocks)
//     S_JJ = K_JJ - sum_{k<J} L_Jk L_Jk^T               DMMA, operands streamed by TMA bulk copies (mbarrier ring)
//     L_JJ by an 8-wide right-looking sweep             pivot tiles factored and inverted by one warp, panels and trailing
//                                                       tiles on the DMMA pipe; z_J = L_JJ^{-1} ytilde_J rides along
//     quad += |z_J|^2 ; logdet += sum log diag(L_JJ)
//     for I > J:
//        S_IJ = K_IJ - sum_{k<J} L_Ik L_Jk^T            DMMA (the N^3/3 term)
//        L_IJ L_JJ^T = S_IJ                             substitution over 8-wide tile columns in the accumulators; only the
//                                                       8 x 8 pivot tiles are inverted, every tile solve is refined once
//        ytilde_I -= L_IJ z_J                           right-looking forward solve, fused
//        store L_IJ (tile-major) to the slot's scratch
//   lnlike = -1/2 (quad + 2 logdet + n log 2 pi)
//
// K is never materialised: each 128 x 128 tile of it is synthesised from the time stamps and the five hyper-parameters
// directly into the DMMA accumulator registers (C-fragment layout) at the first (and only) touch of that tile.
// Reference arithmetic being replaced: gprot/model.py:328-356 (george kernel build + HODLR factor + lnlikelihood).
// DESIGN.md section 3 has the measured critical paths (tools/sweep_bench.cu, trsm_bench.cu, chain_bench.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>

namespace gprot {

constexpr int BS = 128;           // block edge
constexpr int NT = 512;           // threads per CTA
constexpr int SLAB = 2048;        // doubles in one 128 x 16 operand slab (16 KB)
constexpr int BLKD = 16384;       // doubles in one 128 x 128 block (8 slabs)
constexpr int STAGES = 4;         // TMA pipeline depth; one stage = A slab + B slab = 32 KB
constexpr int LDM = 132;          // row pitch of the diagonal-block workspace (== 4 mod 16: conflict-free fragments)
constexpr int REGION_A = 128 * LDM;   // doubles; >= STAGES * 2 * SLAB
constexpr int W_DOUBLES = 9216;   // packed tile-major lower triangle of the solve operator (72 KB)
constexpr int WBUFS = 4;          // cluster kernels: exchange buffers for W_J, z_J, L_pp (a CTA may run up to WBUFS - 1 columns ahead)
constexpr int WX_DOUBLES = W_DOUBLES + BS + 16 * 64;   // one exchange buffer: W | z | L_pp tiles
__host__ __device__ constexpr int cctr_ints(int tcap) { return 3 * tcap + 8; }   // [0] fail, [1] diagdone, [2..] row counters, rowdone, coldone

struct Hyp {
    double A, L, G, S, P;   // exp(theta)
    double hl;              // 0.5 / L
    double thr;             // DBL_EPSILON * L : WhiteKernel fires where d*d/L < DBL_EPSILON
};

struct Smem {
    double regionA[REGION_A];     // TMA stages | X staging of the block solve (tile-major) | diagonal workspace M (row-major)
    double W[W_DOUBLES];          // operator of the solves with L_JJ (tile-major B operands, rows c >= 16*slab only): W_pp = L_pp^{-1}
                                  // on the diagonal 8 x 8 tiles, L_JJ itself below them
    double red[4 * BS];           // cross-warp partial sums of L_IJ z_J
    double zs[BS];                // z_J
    double yj[BS];                // ytilde_J
    double dinv[BS];              // 1 / diag(L_JJ)
    double lpp[16 * 64];          // the 8 x 8 diagonal tiles L_pp of L_JJ, row-major (= tile-major operand layout), zero above the diagonal
    double lpart[16];             // per-warp partial sums of log diag(L_JJ)
    double etab[64];              // 2^(j/64)
    double colv[3 * BS];          // block column J: t, sin, cos of the reduced phase
    double rowv[3 * BS];          // block row I: the same
    Hyp hyp;
    double qpart[16];             // per-warp partial sums of z_J^2 (fixed-order reduction: deterministic)
    unsigned long long full[STAGES];
    unsigned long long empty[STAGES];
    int prob;
    int fail;
    int prob2;                    // block row handed out by the per-column counter (cluster kernels)
};

static_assert(sizeof(Smem) <= 232448, "Smem exceeds the 227 KB a CTA can opt into on sm_100a");

struct Params {
    // packed light-curve arena (all registered chunks back to back)
    const double* t;
    const double* y;
    const double* yerr;
    const long long* chunk_off;   // [nchunks] offset of the chunk in the arena
    const int* chunk_n;           // [nchunks]
    const double* chunk_mingap2;  // [nchunks] min squared gap between distinct stamps (white-kernel test)
    const double* theta;          // [B,5]
    const int* prob_theta;        // [nprob] row of theta
    const int* prob_chunk;        // [nprob] chunk index
    double* prob_out;             // [nprob]
    const int* prob_order;        // [nprob] queue order (largest problems first) or nullptr = as listed
    int nprob;
    int* counter;                 // work queue head
    double* lscratch;             // per-slot factor scratch
    size_t lslot_stride;          // doubles
    double* vscratch;             // per-slot vectors: t (clamped), sin, cos, diag, ytilde (5 * npad_max)
    int npad_max;
    long long* timing;            // [16] per-phase cycle totals of CTA 0 (only written when built with -DGPROT_TIMING)
    double* wg;                   // cluster kernels, per slot: WBUFS x [W | z] exchange buffers, then per-column sums (2 x tcap)
    size_t wg_stride;             // doubles
    int* cctr;                    // cluster kernels, per slot: CCTR_INTS(tcap) ints of flags and counters (see the cluster path)
    int tcap;                     // capacity in block columns of the per-column arrays
};

#ifdef GPROT_TIMING
#define TICK(k) do { if (tid == 0) { const long long now_ = clock64(); tim_[k] += now_ - tlast_; tlast_ = now_; } } while (0)
#else
#define TICK(k) do { } while (0)
#endif

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    const uint32_t a = smem_u32(bar);
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// order generic-proxy accesses (st.global / ld.shared / st.shared) before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// the same for shared memory only (re-use of a shared-memory region by TMA after generic reads / writes): does not
// wait for this thread's global stores to drain
__device__ __forceinline__ void fence_async_smem() {
#ifdef GPROT_FULL_FENCE
    asm volatile("fence.proxy.async;" ::: "memory");
#else
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
// thread-block cluster: rank, barrier with release/acquire at cluster scope (orders global and shared::cluster
// accesses of all CTAs of the cluster), store into a peer CTA's shared memory (DSMEM)
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// flags in global memory between the CTAs that share a problem (release / acquire at gpu scope)
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void dsmem_store_int(int* local, uint32_t peer, int v) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local)), "r"(peer));
    asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
}

// ---------------------------------------------------------------- layouts
// Tile-major operand layout.  A 128 x 16 slab is stored as [row group 16][k tile 2][8][8]: 8 x 8 tiles, row-major inside.
// Lane (g, q) = lane / 4, lane % 4 of a warp reads the 16 bytes at offset 2 * lane of a tile, i.e. elements [g][2q] and
// [g][2q+1].  As an A (or B) operand of mma.sync.m8n8k4.f64 the first feeds k-slot q of one DMMA, the second k-slot q of
// the next: the two DMMAs of a tile product cover columns {0,2,4,6} and {1,3,5,7} -- any split of k works as long as A
// and B agree, and they always do because both come from this layout.  The same 16 bytes are ALSO the lane's C-fragment
// of that tile ([g][2q], [g][2q+1]): accumulators go to memory, and come back as operands, with no shuffle in between
// (staging of S_IJ, the TRSM strips, the stored blocks of L).  One LDS.128 per lane feeds two DMMAs, conflict-free.
__device__ __forceinline__ int frag_off(int r, int kk) {   // r in [0,128), kk in [0,16)
    return (r >> 3) * 128 + ((kk >> 3) & 1) * 64 + (r & 7) * 8 + (kk & 7);
}
__device__ __forceinline__ int w_slab_off(int s) { return 128 * s * (17 - s); }   // slab s keeps row groups 2s..15
// position of W[c][m] (c >= 16 * (m / 16)) in the packed lower triangle sm.W: k-slab m/16, row group c/8
__device__ __forceinline__ int w_off(int c, int m) {
    const int s = m >> 4;
    return w_slab_off(s) + ((c >> 3) - 2 * s) * 128 + ((m >> 3) & 1) * 64 + (c & 7) * 8 + (m & 7);
}

// Geometry of the diagonal-block sweep, shared by the two kernels (gp_kernel64.cuh has its own for 64 x 64 blocks)
struct Geo128 {
    static constexpr int LD = 132;        // LDM: pitch of the diagonal workspace
    static constexpr int NR = 16;         // 8 x 8 tile rows of a diagonal block
    static constexpr int NTH = 512;       // threads of the CTA
    __device__ static __forceinline__ int wslab(int s) { return w_slab_off(s); }
    __device__ static __forceinline__ int woff(int c, int m) { return w_off(c, m); }
    // Tile row of the sweep owned by warp w (1..15).  Warps 4, 8, 12 share their scheduler (and its FP64 pipe) with the
    // pivot warp 0, whose latency chain is the critical path of the sweep: they get the rows that finish first (1, 2, 3).
    // The other schedulers see rows {4,7,10,13}, {5,8,11,14}, {6,9,12,15}.
    __device__ static __forceinline__ int row_of(int w) {
#ifdef GPROT_NO_ROWPERM
        return w;
#else
        return (w & 3) == 0 ? (w >> 2) : 3 + w - (w >> 2);
#endif
    }
};

// Store a warp's 32 x 32 C-fragment tile (acc[mi][ni][e] at row 8mi+g, col 8ni+2q+e) into a tile-major block at tile
// origin (row0, col0), optionally negated: the lane's two accumulators of an 8 x 8 tile are 16 contiguous bytes.
template <bool NEG, int SLABSTRIDE = SLAB>
__device__ __forceinline__ void store_tile_frag(double* blk, const double (&acc)[4][4][2], int row0, int col0, int lane) {
#pragma unroll
    for (int ni = 0; ni < 4; ni++) {
        const int c = col0 + 8 * ni;                       // multiple of 8
        double* base = blk + (c >> 4) * SLABSTRIDE + ((c >> 3) & 1) * 64 + lane * 2;
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
            const int rg = (row0 >> 3) + mi;
            *reinterpret_cast<double2*>(base + rg * 128) = NEG ? make_double2(-acc[mi][ni][0], -acc[mi][ni][1])
                                                               : make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        }
    }
}

// ---------------------------------------------------------------- exp for the covariance synthesis
// 2^(j/64), correctly rounded (generated with mpmath); copied to shared memory at kernel start.
__constant__ double EXP2_64TH[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0,
};

// exp(x) for finite x <= 0, <= 1 ulp: x = (64 k + j) ln2/64 + r, |r| <= ln2/128, exp = 2^k * T[j] * (1 + p(r)),
// p of degree 5 in Estrin form (truncation r^6/720 < 3.5e-17).  Arguments below -700 give exactly 0 (exp(-700) ~ 1e-304
// is zero against any diagonal this model can have); that compare and the select are integer work.  19 FP64
// instructions per covariance entry in all; the tile synthesis waits on the FP64 pipe (ncu: math-pipe throttle is its
// first stall reason), so each of them counts.  (Measured: folding A and G into the exponent and reducing the argument
// with one FMA saves five more but costs 7x in entry accuracy.)
__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ tab) {
    const double MAGIC = 0x1.8p+52;
    const double t = fma(x, 0x1.71547652b82fep+6, MAGIC);          // 64/ln2
    const int ki = __double2loint(t);                               // 64 k + j (two's complement, <= 0)
    const double kf = t - MAGIC;
    double r = fma(kf, -0x1.62e42fee00000p-7, x);                   // ln2/64 split hi (32 bits) + lo
    r = fma(kf, -0x1.a39ef35793c76p-39, r);
    const double r2 = r * r;
    const double a = fma(r, 0x1.5555555555555p-3, 0.5);             // 1/2 + r/6
    const double b = fma(r, 0x1.1111111111111p-7, 0x1.5555555555555p-5);   // 1/24 + r/120
    const double pp = fma(r2, fma(r2, b, a), r);
    const double tj = tab[ki & 63];
    const double res = fma(tj, pp, tj);
    int hi = __double2hiint(res) + ((ki >> 6) << 20);
    int lo = __double2loint(res);
    // x <= -700 (x is never positive or NaN here: both terms of the exponent are <= 0 and finite or -inf): for negative
    // doubles the order of the magnitudes is the order of the high words as unsigned integers -- an integer compare
    // instead of one more instruction on the FP64 pipe, which the synthesis is bound by
    if ((unsigned)__double2hiint(x) >= 0xC085E000u) { hi = 0; lo = 0; }
    return __hiloint2double(hi, lo);
}

// ---------------------------------------------------------------- covariance synthesis
// EXACT: george's operation order, entry by entry (oracle/gp_oracle.py kernel_matrix).
// fast : sin(pi (ti-tj)/P) = si*cj - ci*sj from per-point phases reduced in double-double, one exp.
// Returns -K_ij (the accumulators hold -S).
template <bool EXACT>
__device__ __forceinline__ double neg_kentry(double ti, double tj, double si, double ci, double sj, double cj, const Hyp& h,
                                             const double* __restrict__ etab) {
    const double d = ti - tj;
    if (EXACT) {
        const double r2 = __ddiv_rn(__dmul_rn(d, d), h.L);
        const double s = sin(__ddiv_rn(__dmul_rn(M_PI, d), h.P));
        const double e1 = exp(__dmul_rn(-0.5, r2));
        const double e2 = exp(__dmul_rn(__dmul_rn(-h.G, s), s));
        return -__dmul_rn(__dmul_rn(h.A, e1), e2);
    } else {
        const double sn = si * cj - ci * sj;
        return -h.A * exp_neg(fma(-h.G * sn, sn, -(d * d) * h.hl), etab);
    }
}

// acc <- -K(I,J) for this warp's 32 x 32 tile, from the per-point values staged in shared memory (sm.rowv: block row,
// sm.colv: block column).  rl0 / cl0 = first row / column of the warp tile inside the block, row0 / col0 the same in
// the problem.  GENERAL handles the diagonal (white noise + yerr), identity padding beyond n, and coincident time
// stamps off the diagonal.
template <bool EXACT, bool GENERAL, int RSTRIDE = BS, int CSTRIDE = BS, class SM = Smem>
__device__ __forceinline__ void synth_tile(double (&acc)[4][4][2], const SM& sm, int rl0, int cl0, int row0, int col0, int n,
                                           const double* vd, bool white_off, int lane) {
    const int g = lane >> 2, q = lane & 3;
    const Hyp& h = sm.hyp;
    double ti[4], si[4], ci[4];
    double vdr[4] = {0.0, 0.0, 0.0, 0.0};                      // diagonal noise of this lane's rows: a global load, issued before
    if (GENERAL && row0 == col0) {                             // the tile's arithmetic instead of inside it (diagonal tiles only)
#pragma unroll
        for (int mi = 0; mi < 4; mi++) vdr[mi] = vd[row0 + 8 * mi + g];
    }
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const int r = rl0 + 8 * mi + g;
        ti[mi] = sm.rowv[r];
        si[mi] = sm.rowv[RSTRIDE + r];
        ci[mi] = sm.rowv[2 * RSTRIDE + r];
    }
#pragma unroll
    for (int ni = 0; ni < 4; ni++) {
        const int cl = cl0 + 8 * ni + 2 * q;
        const double2 tj2 = *reinterpret_cast<const double2*>(sm.colv + cl);
        const double2 sj2 = *reinterpret_cast<const double2*>(sm.colv + CSTRIDE + cl);
        const double2 cj2 = *reinterpret_cast<const double2*>(sm.colv + 2 * CSTRIDE + cl);
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const double tj = e ? tj2.y : tj2.x, sj = e ? sj2.y : sj2.x, cj = e ? cj2.y : cj2.x;
#pragma unroll
            for (int mi = 0; mi < 4; mi++) {
                double nk = neg_kentry<EXACT>(ti[mi], tj, si[mi], ci[mi], sj, cj, h, sm.etab);
                if (GENERAL) {
                    const int r = row0 + 8 * mi + g, c = col0 + 8 * ni + 2 * q + e;
                    const double d = ti[mi] - tj;
                    if (r == c) nk = (nk - h.S) - vdr[mi];
                    else if (white_off && d * d < h.thr) nk -= h.S;
                    if (r >= n || c >= n) nk = (r == c) ? -1.0 : 0.0;
                }
                acc[mi][ni][e] = nk;
            }
        }
    }
}

// stage the per-point values (t, sin, cos) of one 128-point block into shared memory (threads 0..383, one value each)
__device__ __forceinline__ void stage_points(double* dst, const double* vt, int npad_max, int blk, int tid) {
    if (tid < 3 * BS) {
        const int which = tid >> 7, i = tid & (BS - 1);
        dst[tid] = vt[(size_t)which * npad_max + blk * BS + i];       // vt | vs | vc
    }
}

// ---------------------------------------------------------------- streamed accumulation  acc += A * B^T
// The producer (thread 0) keeps STAGES-1 slabs in flight.  `it` is the CTA-wide pipeline iteration counter:
// stage = it % STAGES, parity = (it / STAGES) & 1, identical in every thread.
template <bool DIAG>
__device__ __forceinline__ void issue_stage(Smem& sm, uint32_t it, const double* arow, const double* brow, int i) {
    const uint32_t st = it % STAGES;
    mbar_wait(&sm.empty[st], ((it / STAGES) & 1) ^ 1);
    double* dst = sm.regionA + st * (2 * SLAB);
    mbar_expect_tx(&sm.full[st], (DIAG ? 1 : 2) * SLAB * 8);
    tma_load(dst, arow + (size_t)i * SLAB, SLAB * 8, &sm.full[st]);
    if (!DIAG) tma_load(dst + SLAB, brow + (size_t)i * SLAB, SLAB * 8, &sm.full[st]);
}

template <bool DIAG>
__device__ __forceinline__ void prologue(Smem& sm, uint32_t it, const double* arow, const double* brow, int niter) {
    const int np = niter < STAGES - 1 ? niter : STAGES - 1;
    for (int i = 0; i < np; i++) issue_stage<DIAG>(sm, it + i, arow, brow, i);
}

template <bool DIAG>
__device__ __forceinline__ void accumulate(Smem& sm, double (&acc)[4][4][2], uint32_t& it, const double* arow, const double* brow,
                                           int niter, int wr, int wc, int lane, bool active) {
    const int tid = threadIdx.x;
    for (int i = 0; i < niter; i++, it++) {
        if (tid == 0 && i + STAGES - 1 < niter) issue_stage<DIAG>(sm, it + STAGES - 1, arow, brow, i + STAGES - 1);
        const uint32_t st = it % STAGES;
        mbar_wait(&sm.full[st], (it / STAGES) & 1);
        if (active) {
            const double* As = sm.regionA + st * (2 * SLAB);
            const double* Bs = DIAG ? As : As + SLAB;
#pragma unroll
            for (int kp = 0; kp < 2; kp++) {
                double2 a[4], b[4];
#pragma unroll
                for (int mi = 0; mi < 4; mi++) a[mi] = *reinterpret_cast<const double2*>(As + (4 * wr + mi) * 128 + kp * 64 + lane * 2);
#pragma unroll
                for (int ni = 0; ni < 4; ni++) b[ni] = *reinterpret_cast<const double2*>(Bs + (4 * wc + ni) * 128 + kp * 64 + lane * 2);
                if (DIAG && wr == wc) {                            // symmetric tile: 10 of the 16 8x8 sub-tiles
#pragma unroll
                    for (int mi = 0; mi < 4; mi++)
#pragma unroll
                        for (int ni = 0; ni <= mi; ni++) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi].x, b[ni].x);
#pragma unroll
                    for (int mi = 0; mi < 4; mi++)
#pragma unroll
                        for (int ni = 0; ni <= mi; ni++) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi].y, b[ni].y);
                } else {
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
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[st]);
    }
}

// ---------------------------------------------------------------- diagonal block: L_JJ by an 8-wide right-looking sweep
// Workspace M (row-major, pitch LDM) holds S_JJ in its lower triangle and zeros above.  Eliminating 8 columns at a time:
//
// (i)   warp 0 factors the 8 x 8 pivot tile and inverts it in registers (tile8_factor_all);
// (ii)  every tile of the panel below it is solved against the pivot tile: product with the tile inverse, one step of
//       iterative refinement (panel_scale), and stored both in M and in the operator of the block's later solves (sm.W);
// (iii) all tiles right of the panel receive the rank-8 update (2 DMMAs per tile), held in registers by the warp that
//       owns the tile row (diag_factor_rows).
// Only the 8 x 8 pivot tiles are ever inverted; 1/diag(L) goes to sm.dinv, the pivot tiles themselves to sm.lpp.
// 1/sqrt(d) for a positive, finite, normal d: MUFU.RSQ64H seed (2^-22.9) and one third-order correction
// (error ~ (5/16) e^3 ~ 2^-70), branch free -- the 8 x 8 tile factorisation below is a pure latency chain.
__device__ __forceinline__ double rsqrt_pos(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-(d * y), y, 1.0);
    return fma(y, e * fma(0.375, e, 0.5), y);
}

// Pivot-tile factorisation without shuffles: every lane of the pivot warp holds the whole lower triangle and runs the
// same elimination; lane b < 8 also forms column b of the inverse by forward substitution.  The dependent chain is
// pivot -> rsqrt (MUFU + 5 FP64) -> scale -> one DFMA -> next pivot, 72 cycles per column (tools/chain_bench.cu); a warp
// issues in order, so everything else is kept small and free of that chain:
//   * inverse row j is w[j] = ri_j (delta_jb - sum_{m<j} l[j][m] w[m]): the sum only needs row j of L, which is final once
//     column j-1 is done, so it is formed while rsqrt(d_j) is in flight and one multiplication per row is left behind it;
//   * L_pp, sqrt(d_j) and 1/l_jj are values every lane has: they are stored with lane-independent addresses (all lanes
//     write the same word with the same value) instead of being picked per lane with selects (the first round-2 version
//     spent 105 of its 464 instructions on FSEL, formed the inverse after the elimination and indexed it by lane through
//     local memory: 1413 cycles alone on an SM against ~580 for the chain).
// Needs ~100 registers, which only the pivot warp's own loop can afford.  Outputs: W_pp (diagonal tile p of sm.W: the
// operator of the panel and of the TRSM), L_pp (sm.lpp, zero above the diagonal: set once per kernel), 1/diag (sm.dinv).
template <class G, class SM>
__device__ __forceinline__ void tile8_factor_all(SM& sm, int p, int lane) {
    const double* Mp = sm.regionA + (8 * p) * G::LD + 8 * p;
    double l[8][8];
#pragma unroll
    for (int a = 0; a < 8; a++) {
#pragma unroll
        for (int b2 = 0; b2 <= a / 2; b2++) {
            const double2 v = *reinterpret_cast<const double2*>(Mp + a * G::LD + 2 * b2);
            l[a][2 * b2] = v.x;
            if (2 * b2 + 1 <= a) l[a][2 * b2 + 1] = v.y;
        }
    }
    const int b = lane & 7;
    double w[8];                                               // column b of W_pp
    double rinv[8];
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const double d = l[j][j];
        if (!(d > 0.0) || !(d < DBL_MAX)) bad = true;
        const double ri = rsqrt_pos(d);                        // 1 / l_jj to ~1 ulp
        rinv[j] = ri;
        if (j < 7) {                                           // the chain first: next pivot
            l[j + 1][j] *= ri;
            l[j + 1][j + 1] = fma(-l[j + 1][j], l[j + 1][j], l[j + 1][j + 1]);
        }
        double sj = (j == b) ? 1.0 : 0.0;                      // in the shadow of the rsqrt (operands final before it)
#pragma unroll
        for (int m = 0; m < j; m++) sj = fma(-l[j][m], w[m], sj);
        w[j] = ri * sj;
        double lj = d * ri;                                    // l_jj = sqrt(d) = d / sqrt(d), one Newton step: <= 1 ulp
        l[j][j] = fma(fma(-lj, lj, d), 0.5 * ri, lj);
#pragma unroll
        for (int a = j + 2; a < 8; a++) l[a][j] *= ri;
#pragma unroll
        for (int a = j + 1; a < 8; a++)
#pragma unroll
            for (int c = j + 1; c <= a; c++)
                if (!(a == j + 1 && c == j + 1)) l[a][c] = fma(-l[a][j], l[c][j], l[a][c]);
    }
    {
        double* lp = sm.lpp + 64 * p;
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int k = 0; k <= a / 2; k++)
                *reinterpret_cast<double2*>(lp + a * 8 + 2 * k) = make_double2(l[a][2 * k], (2 * k + 1 <= a) ? l[a][2 * k + 1] : 0.0);
#pragma unroll
        for (int k = 0; k < 4; k++) *reinterpret_cast<double2*>(sm.dinv + 8 * p + 2 * k) = make_double2(rinv[2 * k], rinv[2 * k + 1]);
    }
    if (lane < 8) {
        double* wp = sm.W + G::woff(8 * p, 8 * p);             // diagonal tile p: row-major inside the tile
#pragma unroll
        for (int a = 0; a < 8; a++) wp[a * 8 + b] = w[a];
    }
    if (bad && lane == 0) sm.fail = 1;
}

// ---------------------------------------------------------------- diagonal block, trailing tiles in registers
// The rank-8 updates never touch shared memory: warp R owns tile row R of the workspace (8 rows x 128 columns) and keeps
// its not-yet-eliminated tiles (R, C), C > p, as DMMA accumulators.  Shared memory only carries the current panel column
// p (M(R, p), rewritten in place with X(R) = M(R, p) L_pp^-T), the next panel column (each warp stores tile (R, p+1)
// right after updating it) and the 8 x 8 pivot tile.  Per step the block-wide work is two 8-byte fragment loads per tile
// instead of a 16-byte load + store of every trailing tile (an in-place version was bound by shared-memory bandwidth,
// not by the DMMA pipe); the critical path is the owner of tile row p+1: one tile update, then the pivot-tile
// factorisation (DESIGN.md section 3 has the measured time line).
// rank-8 update of the register tiles C = P+2 .. chi of one tile row with panel column P (a0, a1 = A fragment of -X(R));
// the tile that becomes the next-but-one panel column (C = P+2) is parked in shared memory right after its update
template <class G, int P>
__device__ __forceinline__ void rows_update(double (&rt)[G::NR][2], const double* __restrict__ bcol, double* myrow, double a0,
                                            double a1, int chi, int q) {
#pragma unroll
    for (int C = P + 2; C < G::NR; C++) {
        if (C <= chi) {
            const double b0 = bcol[(8 * C) * G::LD + q], b1 = bcol[(8 * C) * G::LD + q + 4];
            dmma(rt[C][0], rt[C][1], a0, b0);
            dmma(rt[C][0], rt[C][1], a1, b1);
            if (C == P + 2) *reinterpret_cast<double2*>(myrow + 8 * C + 2 * q) = make_double2(rt[C][0], rt[C][1]);
        }
    }
}

#ifdef DIAG_TS   // tools/sweep_bench.cu: time stamps of the pivot warp (TS) and of the owner of tile row p+1 (TSR) per step
__device__ long long g_ts[16][16];
#define TS(p, k) do { if (lane == 0) g_ts[p][k] = clock64(); } while (0)
#define TSR(p, k) do { if (lane == 0 && R == (p) + 1) g_ts[p][8 + (k)] = clock64(); } while (0)
#else
#define TS(p, k) do { } while (0)
#define TSR(p, k) do { } while (0)
#endif
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// barrier 2 + wr for the warps of warp row wr (immediate ids: a register id would make ptxas reserve all 16 barriers)
__device__ __forceinline__ void row_sync(int wr, int count) {
    switch (wr) {
        case 0: asm volatile("bar.sync 2, %0;" ::"r"(count) : "memory"); break;
        case 1: asm volatile("bar.sync 3, %0;" ::"r"(count) : "memory"); break;
        case 2: asm volatile("bar.sync 4, %0;" ::"r"(count) : "memory"); break;
        default: asm volatile("bar.sync 5, %0;" ::"r"(count) : "memory"); break;
    }
}
// the producer's side of the same barrier: counted, not waited for
__device__ __forceinline__ void row_arrive(int wr, int count) {
    switch (wr) {
        case 0: asm volatile("bar.arrive 2, %0;" ::"r"(count) : "memory"); break;
        case 1: asm volatile("bar.arrive 3, %0;" ::"r"(count) : "memory"); break;
        case 2: asm volatile("bar.arrive 4, %0;" ::"r"(count) : "memory"); break;
        default: asm volatile("bar.arrive 5, %0;" ::"r"(count) : "memory"); break;
    }
}

// X = M(R, p) L_pp^-T for one tile row below the pivot (R > p), rewritten in place and stored in sm.W (the operator of
// trsm_inplace: W_pp on the diagonal tiles, L_JJ below).  Returns the A fragment of -X.
// The product with the explicit inverse W_pp is only a first guess X0: its backward error is eps * cond(L_pp), which
// costs a digit where neighbouring points are strongly correlated (cond of an 8 x 8 tile reaches 1e7).  One step of
// iterative refinement on the tensor pipe, X = X0 + (M - X0 L_pp^T) W_pp^T, brings the tile solve to the accuracy of a
// substitution (tools/emulate_trsm.py; profiles/r02_trsm_emulation.log).
template <class G, class SM>
__device__ __forceinline__ void panel_scale(SM& sm, double* row, const double* wt, int R, int p, int g, int q, double& a0, double& a1) {
    const double x0 = row[q], x1 = row[q + 4];
    const double2 m = *reinterpret_cast<const double2*>(row + 2 * q);          // M[g][2q], M[g][2q+1]: C-fragment of the tile
    const double b0 = wt[g * 8 + q], b1 = wt[g * 8 + q + 4];
    double c0 = 0.0, c1 = 0.0;
    dmma(c0, c1, x0, b0);
    dmma(c0, c1, x1, b1);
    const double2 lp = *reinterpret_cast<const double2*>(sm.lpp + 64 * p + g * 8 + 2 * q);   // L_pp[g][2q], [2q+1] as a B operand
    const double2 wv = *reinterpret_cast<const double2*>(wt + g * 8 + 2 * q);                // W_pp[g][2q], [2q+1]
    double r0 = m.x, r1 = m.y;
    dmma(r0, r1, -c0, lp.x);                                                   // R = M - X0 L_pp^T  (a C fragment is an A operand)
    dmma(r0, r1, -c1, lp.y);
    dmma(c0, c1, r0, wv.x);                                                    // X = X0 + R W_pp^T
    dmma(c0, c1, r1, wv.y);
    __syncwarp();
    *reinterpret_cast<double2*>(row + 2 * q) = make_double2(c0, c1);
    // the finished tile L(R, p) also goes to its place in the operator of the column's solves (tile-major: the C fragment as is)
    *reinterpret_cast<double2*>(sm.W + G::woff(8 * R + g, 8 * p + 2 * q)) = make_double2(c0, c1);
    __syncwarp();
    a0 = -row[q]; a1 = -row[q + 4];
}

// Warp 0 is the pivot warp: it only factors and inverts the 8 x 8 pivot tiles.  Warps 1..NR-1 own tile rows 1..NR-1.
// The two roles are separate loops, so the pivot-tile registers and the row accumulators never coexist; they meet at
// the two block barriers of every step, and the owner of tile row p+1 hands the finished pivot tile to warp 0 through
// named barrier 1.  Only L_JJ and the pivot-tile inverses W_pp are formed -- every solve with L_JJ is a substitution
// over the tiles, no inverse of the block exists.  The forward solve z_J = L_JJ^-1 ytilde_J rides along: warp 0 forms
// z_p = W_pp ytilde_p right after the pivot tile (sm.zs), every row warp subtracts L(R, p) z_p from its rows of ytilde.
// nsteps < NR: rows 8 nsteps .. of the block are identity padding; the sweep stops early and diag_pad_identity fills
// their part of the operator.  Ends with a block barrier.
template <class G, class SM>
__device__ __forceinline__ void diag_factor_rows(SM& sm, int warp, int lane, int nsteps) {
    constexpr int LD = G::LD, NR = G::NR;
    const int g = lane >> 2, q = lane & 3;
    double* M = sm.regionA;
    if (warp == 0) {
        tile8_factor_all<G>(sm, 0, lane);
        __syncthreads();
        for (int p = 0; p < nsteps; p++) {
            TS(p, 0);
            {                                                  // z_p = W_pp ytilde_p while the row warps scale their panel tiles
                const double* wrow = sm.W + G::woff(8 * p, 8 * p) + 8 * (lane & 7);    // row of W_pp, zero above the diagonal
                double z = 0.0;
#pragma unroll
                for (int k = 0; k < 8; k++) z = fma(wrow[k], sm.yj[8 * p + k], z);
                if (lane < 8) sm.zs[8 * p + lane] = z;
            }
            TS(p, 1);
            __syncthreads();
            TS(p, 2);
            if (p == nsteps - 1) break;
            named_sync(1, 64);                                 // pivot tile (p+1, p+1) is final
            TS(p, 3);
            tile8_factor_all<G>(sm, p + 1, lane);
            TS(p, 4);
            __syncthreads();
        }
        return;
    }
    const int R = G::row_of(warp);
    double* myrow = M + (8 * R + g) * LD;
    double rt[NR][2];                                          // rt[C] = C fragment of tile (R, C), 2 <= C <= R
#pragma unroll
    for (int C = 2; C < NR; C++) {
        if (C <= R) {
            const double2 v = *reinterpret_cast<const double2*>(myrow + 8 * C + 2 * q);
            rt[C][0] = v.x; rt[C][1] = v.y;
        } else { rt[C][0] = 0.0; rt[C][1] = 0.0; }
    }
    double yR = sm.yj[8 * R + g];                              // ytilde of this lane's row, updated as the columns finish
    __syncthreads();
    for (int p = 0; p < nsteps; p++) {
        const int P0 = 8 * p;
        const double* wt = sm.W + G::woff(8 * p, 8 * p);             // W_pp
        double a0 = 0.0, a1 = 0.0;                             // A fragment of -X(R)
        TSR(p, 0);
        if (R > p) panel_scale<G>(sm, myrow + P0, wt, R, p, g, q, a0, a1);
        TSR(p, 1);
        __syncthreads();                                       // X(C), C > p, and z_p visible to every warp
        TSR(p, 2);
        if (p == nsteps - 1) break;
        if (R > p) {                                           // rows at or above the pivot are finished
            {                                                  // forward solve, fused: ytilde_R -= L(R, p) z_p
                double part = fma(a0, sm.zs[P0 + q], a1 * sm.zs[P0 + q + 4]);
                part += __shfl_xor_sync(0xffffffffu, part, 1);
                part += __shfl_xor_sync(0xffffffffu, part, 2);
                yR += part;
            }
            const double* bcol = M + g * LD + P0;              // B fragments: X(C) at bcol + 8 C LDM
            {                                                  // the next panel column lives in shared memory
                double2* cp = reinterpret_cast<double2*>(myrow + P0 + 8 + 2 * q);
                double2 c = *cp;
                const double b0 = bcol[8 * (p + 1) * LD + q], b1 = bcol[8 * (p + 1) * LD + q + 4];
                dmma(c.x, c.y, a0, b0);
                dmma(c.x, c.y, a1, b1);
                *cp = c;
            }
            if (R == p + 1) {
                if (q == 0) sm.yj[8 * R + g] = yR;             // ytilde of the next pivot rows is final
                __syncwarp();
                __threadfence_block();
                TSR(p, 3);
                named_arrive(1, 64);                           // hand the pivot tile to warp 0
            } else {
                switch (p) {                                   // tiles p+2 .. R of this row take the rank-8 update
                    case 0: if constexpr (0 <= NR - 3) rows_update<G, 0>(rt, bcol, myrow, a0, a1, R, q); break;
                    case 1: if constexpr (1 <= NR - 3) rows_update<G, 1>(rt, bcol, myrow, a0, a1, R, q); break;
                    case 2: if constexpr (2 <= NR - 3) rows_update<G, 2>(rt, bcol, myrow, a0, a1, R, q); break;
                    case 3: if constexpr (3 <= NR - 3) rows_update<G, 3>(rt, bcol, myrow, a0, a1, R, q); break;
                    case 4: if constexpr (4 <= NR - 3) rows_update<G, 4>(rt, bcol, myrow, a0, a1, R, q); break;
                    case 5: if constexpr (5 <= NR - 3) rows_update<G, 5>(rt, bcol, myrow, a0, a1, R, q); break;
                    case 6: if constexpr (6 <= NR - 3) rows_update<G, 6>(rt, bcol, myrow, a0, a1, R, q); break;
                    case 7: if constexpr (7 <= NR - 3) rows_update<G, 7>(rt, bcol, myrow, a0, a1, R, q); break;
                    case 8: if constexpr (8 <= NR - 3) rows_update<G, 8>(rt, bcol, myrow, a0, a1, R, q); break;
                    case 9: if constexpr (9 <= NR - 3) rows_update<G, 9>(rt, bcol, myrow, a0, a1, R, q); break;
                    case 10: if constexpr (10 <= NR - 3) rows_update<G, 10>(rt, bcol, myrow, a0, a1, R, q); break;
                    case 11: if constexpr (11 <= NR - 3) rows_update<G, 11>(rt, bcol, myrow, a0, a1, R, q); break;
                    case 12: if constexpr (12 <= NR - 3) rows_update<G, 12>(rt, bcol, myrow, a0, a1, R, q); break;
                    case 13: if constexpr (13 <= NR - 3) rows_update<G, 13>(rt, bcol, myrow, a0, a1, R, q); break;
                    default: break;
                }
            }
        }
        __syncthreads();                                       // next panel column and its pivot tile are in place
    }
}

// rows c >= c0 of W (and of 1/diag) are those of the identity
template <class G, class SM>
__device__ __forceinline__ void diag_pad_identity(SM& sm, int c0, int tid) {
    for (int s = 0; s < G::NR / 2; s++) {
        const int cnt = (G::NR - 2 * s) * 128;
        double* dst = sm.W + G::wslab(s);
        for (int d = tid; d < cnt; d += G::NTH) {
            const int rgrel = d >> 7, kp = (d >> 6) & 1;
            const int c = 8 * (2 * s + rgrel) + ((d >> 3) & 7);
            const int m = 16 * s + 8 * kp + (d & 7);
            if (c >= c0) dst[d] = (m == c) ? 1.0 : 0.0;
        }
    }
    if (tid >= c0 && tid < 8 * G::NR) { sm.dinv[tid] = 1.0; sm.zs[tid] = 0.0; }      // padded rows: L = I, ytilde = 0
    for (int d = 8 * c0 + tid; d < 64 * G::NR; d += G::NTH)   // L_pp of the padded tiles (c0 is a multiple of 8)
        sm.lpp[d] = ((d >> 3) & 7) == (d & 7) ? 1.0 : 0.0;
}

// |z_J|^2 and sum log L_rr of the block, as fixed-order per-warp partial sums (sm.qpart / sm.lpart)
template <class G, class SM>
__device__ __forceinline__ void diag_sums(SM& sm, int warp, int lane) {
    constexpr int NB = 8 * G::NR;
    const int i = 32 * warp + lane;
    double z2 = 0.0, lg = 0.0;
    if (i < NB) {
        const double z = sm.zs[i];
        z2 = z * z;
        lg = -log(sm.dinv[i]);                                 // log L_rr (identity padding gives exactly 0)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        z2 += __shfl_xor_sync(0xffffffffu, z2, o);
        lg += __shfl_xor_sync(0xffffffffu, lg, o);
    }
    if (lane == 0) { sm.qpart[warp] = z2; sm.lpart[warp] = lg; }
}

// ---------------------------------------------------------------- TRSM by substitution over 8-wide tile columns
// X L_JJ^T = S solved the way LAPACK's dtrsm does, at the granularity of the DMMA tile:
//     X_c = (S_c - sum_{m<c} X_m L_cm^T) W_cc^T ,   c = 0 .. NR-1,   W_cc = L_cc^{-1} (8 x 8 pivot-tile inverses only).
// Round 1 multiplied by the explicit inverse of the whole diagonal block, whose backward error is eps * cond(L_JJ): up to
// 40x further from the 80-bit truth than LAPACK at cond(K) ~ 1e10; 32-wide groups still lose a digit, 16- and 8-wide
// ones are at LAPACK's level (tools/emulate_trsm.py, profiles/r02_trsm_emulation.log).
//
// The block stays in the accumulators of the 4 x NG warp grid (warp (wr, wc): rows 32 wr.., columns 32 wc..), sixteen
// independent 8 x 8 tiles per warp, so the DMMA pipe never waits on a dependent chain:
//   * columns of OTHER warp columns (m < 4 wc): the finished X of warp (wr, c), c < wc, is read from the staging block
//     as an A operand -- one 32-deep product per earlier warp column, after a named barrier of the warp row on which
//     the producer only arrives: a warp leaves as soon as its own columns are solved and does its L z sums and its
//     stores while the warps to its right still work (tools/trsm_bench.cu, profiles/r02_trsm_bench.log: the block
//     epilogue against 21.5 k cycles of DMMA pipe time; giving the next solver the new X before the others was
//     measured too and changes nothing -- the phases in which one warp per scheduler works alone run at 21 cycles per
//     DMMA instead of 16.7 whatever their order);
//   * the warp's OWN four tile columns: substitution inside the registers.  In the tile-major layout a C fragment IS an
//     A operand, so T_c feeds W_cc^T and X_c feeds the updates of the tiles to its right without leaving the registers;
//     the four tile rows (mi) are independent chains.  Each 8 x 8 tile solve X_c L_cc^T = T_c is the product with the
//     explicit inverse followed by one step of iterative refinement against L_cc (sm.lpp).
// sm.W holds the operator in B-operand layout (written by the diagonal sweep): W_cc in the diagonal tiles, L_JJ itself in the tiles
// below.  acc enters as -S and leaves as X = L_IJ.  stag: 128-row staging block, written with X for the warps to the right.
#ifdef TRSM_TS   // tools/trsm_bench.cu -DTRSM_TS: per warp {own solve starts, solved, staged, fenced}
__device__ long long g_trsm_ts[16][4];
#define TRSM_STAMP(k) do { if (lane == 0) g_trsm_ts[threadIdx.x >> 5][k] = clock64(); } while (0)
#else
#define TRSM_STAMP(k) do { } while (0)
#endif
template <class G, int NG>
__device__ __forceinline__ void trsm_inplace(double (&acc)[4][4][2], double* __restrict__ stag, const double* __restrict__ Wm,
                                             const double* __restrict__ Lpp, int wr, int wc, int lane) {
#pragma unroll 1
    for (int c = 0; c <= wc; c++) {
        if (wc == c) {
            TRSM_STAMP(0);
            // ---- own columns: tiles 4c .. 4c+3 of the block, acc = -T
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
                const int C = 4 * c + ni;                                                 // tile column of the block
                const double* bcol = Wm + G::wslab(C >> 1) - (C >> 1) * 256 + (C & 1) * 64 + lane * 2;   // tile (R, C) at bcol + 128 R
                const double2 w = *reinterpret_cast<const double2*>(bcol + 128 * C);       // W_CC
                const double2 lp = *reinterpret_cast<const double2*>(Lpp + 64 * C + lane * 2);   // L_CC
                const double nwx = -w.x, nwy = -w.y;
                // acc holds -T.  X0 = T W_CC^T = (-T)(-W_CC)^T; then one step of iterative refinement (see panel_scale) with
                // the residual kept negated in place of -T: -R = -T + X0 L_CC^T, X = X0 + R W_CC^T = X0 + (-R)(-W_CC)^T.
                // No negated copies and no second set of accumulators: bitwise the same as forming R = T - X0 L_CC^T.
                double xd[4][2];
#pragma unroll
                for (int mi = 0; mi < 4; mi++) {
                    xd[mi][0] = 0.0; xd[mi][1] = 0.0;
                    dmma(xd[mi][0], xd[mi][1], acc[mi][ni][0], nwx);
                }
#pragma unroll
                for (int mi = 0; mi < 4; mi++) dmma(xd[mi][0], xd[mi][1], acc[mi][ni][1], nwy);
#pragma unroll
                for (int mi = 0; mi < 4; mi++) dmma(acc[mi][ni][0], acc[mi][ni][1], xd[mi][0], lp.x);
#pragma unroll
                for (int mi = 0; mi < 4; mi++) dmma(acc[mi][ni][0], acc[mi][ni][1], xd[mi][1], lp.y);
#pragma unroll
                for (int mi = 0; mi < 4; mi++) dmma(xd[mi][0], xd[mi][1], acc[mi][ni][0], nwx);
#pragma unroll
                for (int mi = 0; mi < 4; mi++) dmma(xd[mi][0], xd[mi][1], acc[mi][ni][1], nwy);
#pragma unroll
                for (int mi = 0; mi < 4; mi++) { acc[mi][ni][0] = xd[mi][0]; acc[mi][ni][1] = xd[mi][1]; }
#pragma unroll
                for (int n2 = ni + 1; n2 < 4; n2++) {                                     // -T_n2 += X_C L_{n2,C}^T
                    const double2 b = *reinterpret_cast<const double2*>(bcol + 128 * (4 * c + n2));
#pragma unroll
                    for (int mi = 0; mi < 4; mi++) dmma(acc[mi][n2][0], acc[mi][n2][1], xd[mi][0], b.x);
#pragma unroll
                    for (int mi = 0; mi < 4; mi++) dmma(acc[mi][n2][0], acc[mi][n2][1], xd[mi][1], b.y);
                }
            }
            TRSM_STAMP(1);
            if (c + 1 < NG) {                                                            // X for the warps to the right: counted on their
                store_tile_frag<false>(stag, acc, 32 * wr, 32 * wc, lane);                // barrier, not waited for -- this warp is done
                TRSM_STAMP(2);
                __threadfence_block();                                                   // and goes on to its L z sums and its stores
                TRSM_STAMP(3);
                row_arrive(wr, 32 * (NG - c));                                           // while the others still solve
            }
        } else {
            row_sync(wr, 32 * (NG - c));                                                 // X_c is in the staging block (the producer and
            {                                                                            // the warps right of it: NG - c warps)
                // -T_wc += X_c L_{wc,c}^T, 32 deep
#pragma unroll
                for (int ds = 0; ds < 2; ds++) {
                    const int s = 2 * c + ds;
                    const double* As = stag + s * SLAB;
                    const double* Bs = Wm + G::wslab(s) - (2 * s) * 128;
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
            }
        }
    }
}

// sum over this warp's 32 columns of L_IJ z_J for its 32 rows -> red[wc * stride + row]
__device__ __forceinline__ void lz_partials(const double (&acc)[4][4][2], const double* __restrict__ zs, double* red, int stride,
                                            int wr, int wc, int lane) {
    const int g = lane >> 2, q = lane & 3;
    double part[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int ni = 0; ni < 4; ni++) {
        const double z0 = zs[32 * wc + 8 * ni + 2 * q], z1 = zs[32 * wc + 8 * ni + 2 * q + 1];
#pragma unroll
        for (int mi = 0; mi < 4; mi++) part[mi] = fma(acc[mi][ni][0], z0, fma(acc[mi][ni][1], z1, part[mi]));
    }
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        part[mi] += __shfl_xor_sync(0xffffffffu, part[mi], 1);
        part[mi] += __shfl_xor_sync(0xffffffffu, part[mi], 2);
        if (q == 0) red[wc * stride + 32 * wr + 8 * mi + g] = part[mi];
    }
}

// ---------------------------------------------------------------- per-problem set-up shared by the kernels
// exp(theta) and the derived constants (thread 0 of the CTA)
__device__ __forceinline__ Hyp make_hyper(const double* th) {
    Hyp h;
    h.A = exp(th[0]); h.L = exp(th[1]); h.G = exp(th[2]); h.S = exp(th[3]); h.P = exp(th[4]);
    h.hl = fmin(0.5 / h.L, DBL_MAX);                 // denormal L: d = 0 keeps exp(0), d != 0 underflows, as in the reference
    h.thr = DBL_EPSILON * h.L;
    return h;
}
// a non-finite or non-positive scale makes K non-finite or singular by construction: the reference returns -inf
// (mingap2 < 0 flags a chunk with non-finite t or yerr)
__device__ __forceinline__ bool hyper_ok(const Hyp& h, double mingap2) {
    return h.A >= 0.0 && h.A <= DBL_MAX && h.L > 0.0 && h.L <= DBL_MAX && h.G >= 0.0 && h.G <= DBL_MAX &&
           h.S >= 0.0 && h.S <= DBL_MAX && h.P > 0.0 && h.P <= DBL_MAX && mingap2 >= 0.0;
}
// per-point vectors of point i (i >= n: identity padding): clamped time stamp, sin / cos of the phase of t_i / P reduced in
// double-double, diagonal noise, working copy of y
__device__ __forceinline__ void point_vectors(int i, int n, const double* __restrict__ t, const double* __restrict__ y,
                                              const double* __restrict__ yerr, const Hyp& h, double* vt, double* vs, double* vc,
                                              double* vd, double* yw) {
    double s = 0.0, c = 1.0, dg = 0.0, yy = 0.0;
    const double ti = t[min(i, n - 1)];
    if (i < n) {
        const double hi = ti / h.P;
        const double lo = fma(-hi, h.P, ti) / h.P;      // t/P = hi + lo to ~106 bits
        const double ph = (hi - rint(hi)) + lo;          // |ph| <= 0.5 (+ tiny): sin^2 is invariant to the dropped integer
        sincospi(ph, &s, &c);
        const double e = sqrt(h.S + yerr[i] * yerr[i]);  // gprot/model.py:345-346
        dg = e * e;
        yy = y[i];
    }
    vt[i] = ti; vs[i] = s; vc[i] = c; vd[i] = dg; yw[i] = yy;
}
// lnlike from the two sums; non-finite -> -inf (gprot/model.py:356)
__device__ __forceinline__ double finish_lnlike(double quad, double logdet, int n, bool failed) {
    if (failed) return -INFINITY;
    const double lnl = -0.5 * quad - 0.5 * (2.0 * logdet + (double)n * 1.8378770664093453);   // log(2 pi)
    return (fabs(lnl) <= DBL_MAX) ? lnl : -INFINITY;
}

// ---------------------------------------------------------------- the kernel
// CL = CTAs per problem.  CL > 1 launches thread-block clusters of CL CTAs that share one problem: every CTA forms the
// diagonal block of a column (redundantly: same operands, same order, bitwise the same W and z_J -- the CTAs would
// otherwise wait for it), block row I below the diagonal belongs to CTA I mod CL, and one cluster barrier per column
// publishes the new blocks of L and the updated right-hand side.  Results are bitwise identical for every CL.
template <bool EXACT, int CL>
__global__ void __launch_bounds__(NT, 1) gp_lnlike_kernel(const Params prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp >> 2, wc = ((warp & 3) - (warp >> 2)) & 3;   // skewed so each scheduler sees every wc
    const int g = lane >> 2, q = lane & 3;

    if (tid < 64) {
        sm.etab[tid] = EXP2_64TH[tid];
        for (int sl = 0; sl < 8; sl++) sm.W[w_slab_off(sl) + 64 + tid] = 0.0;   // tiles above the diagonal: never written again
    }
    for (int i = tid; i < 16 * 64; i += NT) sm.lpp[i] = 0.0;         // L_pp tiles: only the lower triangles are ever written
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], 16);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t it = 0;   // pipeline iteration counter (uniform across the CTA)
#ifdef GPROT_TIMING
    long long tim_[16] = {0}, tlast_ = clock64();
#endif
    const int crank = CL > 1 ? (int)cluster_rank() : 0;
    const int slot = CL > 1 ? (int)(blockIdx.x / CL) : (int)blockIdx.x;
    double* const lbase = prm.lscratch + (size_t)slot * prm.lslot_stride;
    double* const vt = prm.vscratch + (size_t)slot * 5 * prm.npad_max;      // t (clamped to the last real point)
    double* const vs = vt + prm.npad_max;                                   // sin, cos of the reduced phase
    double* const vc = vs + prm.npad_max;
    double* const vd = vc + prm.npad_max;                                   // diagonal noise
    double* const yw = vd + prm.npad_max;                                   // running right-hand side

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            if (CL == 1) sm.prob = atomicAdd(prm.counter, 1);
            else if (crank == 0) {
                const int p = atomicAdd(prm.counter, 1);
                for (int c = 0; c < CL; c++) dsmem_store_int(&sm.prob, c, p);
            }
            sm.fail = 0;
        }
        if (CL > 1) cluster_sync(); else __syncthreads();
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
        const int T = (n + BS - 1) / BS;
        const int npad = T * BS;

        if (tid == 0) sm.hyp = make_hyper(th);
        __syncthreads();
        const Hyp& h = sm.hyp;
        const double mingap2 = prm.chunk_mingap2[chunk];
        if (!hyper_ok(h, mingap2)) {
            if (tid == 0 && crank == 0) prm.prob_out[prob] = -INFINITY;
            // every CTA of the cluster takes this branch (same theta, same chunk): meet before rank 0 pops the next problem
            // and overwrites the peers' sm.prob, which their slower threads may not have read yet
            if (CL > 1) cluster_sync();
            continue;
        }
        const bool white_off = mingap2 < h.thr;
        for (int i = crank * NT + tid; i < npad; i += CL * NT) point_vectors(i, n, t, y, yerr, h, vt, vs, vc, vd, yw);
        if (CL > 1 && crank == 0) {                              // flags and counters of this problem (layout: the cluster path below)
            int* const ctr0 = prm.cctr + (size_t)slot * cctr_ints(prm.tcap);
            for (int i = tid; i < 2 + T; i += NT) ctr0[i] = 0;                               // fail, diagdone, row counters
            for (int i = tid; i < T; i += NT) { ctr0[2 + prm.tcap + i] = 0; ctr0[2 + 2 * prm.tcap + i] = 0; }   // rowdone, coldone
        }
        if (CL > 1) cluster_sync(); else __syncthreads();
        TICK(0);

        double logdet = 0.0, quad = 0.0;     // sum log diag(L) and |z|^2, meaningful in thread 0
        double acc[4][4][2];
        bool failed = false;

        // ---- diagonal block (J, J): S_JJ, its factor, W = L_JJ^{-1} (sm.W), z_J (sm.zs), partial sums (sm.qpart / lpart).
        // Expects sm.colv / sm.rowv staged for block J.  nextI >= 0: stage the row side of block row nextI and start its
        // first TMA slabs on the way out.  Returns false if a pivot failed.
        auto diag_block = [&](const int J, const int nextI) -> bool {
            const double* jrow = lbase + (size_t)(J * (J - 1) / 2) * BLKD;   // blocks (J, 0..J-1), contiguous
            const int niter = 8 * J;
            const bool pf = nextI >= 0 && tid < 3 * BS;              // per-point values of block row nextI, parked on the way out
            double pv = 0.0;
            if (pf) pv = vt[(size_t)(tid >> 7) * prm.npad_max + nextI * BS + (tid & (BS - 1))];
            {
                // the last block may be mostly identity padding: warp rows beyond the real points do no work at all
                const int ractJ = (J == T - 1) ? ((n - J * BS + 31) >> 5) : 4;
                const bool active = wr >= wc && wr < ractJ;
                if (tid == 0) prologue<true>(sm, it, jrow, jrow, niter);
                if (active) synth_tile<EXACT, true>(acc, sm, 32 * wr, 32 * wc, J * BS + 32 * wr, J * BS + 32 * wc, n, vd, white_off, lane);
                TICK(1);
                accumulate<true>(sm, acc, it, jrow, jrow, niter, wr, wc, lane, active);
                __syncthreads();                                   // all stages consumed: regionA becomes M
                TICK(2);
                {
                    double* M = sm.regionA;
#pragma unroll
                    for (int mi = 0; mi < 4; mi++)
#pragma unroll
                        for (int ni = 0; ni < 4; ni++) {
                            const int r = 32 * wr + 8 * mi + g, c = 32 * wc + 8 * ni + 2 * q;
                            double2 v;
                            v.x = (active && r >= c) ? -acc[mi][ni][0] : ((wr >= ractJ && r == c) ? 1.0 : 0.0);
                            v.y = (active && r >= c + 1) ? -acc[mi][ni][1] : ((wr >= ractJ && r == c + 1) ? 1.0 : 0.0);
                            *reinterpret_cast<double2*>(M + r * LDM + c) = v;
                        }
                    if (tid < BS) sm.yj[tid] = CL > 1 ? __ldcg(yw + J * BS + tid) : yw[J * BS + tid];
                }
                __syncthreads();
                TICK(3);
                {
                    const int nreal = min(BS, n - J * BS);          // the rest of the block is identity padding
                    const int nsteps = (nreal + 7) >> 3;
                    diag_factor_rows<Geo128>(sm, warp, lane, nsteps);       // ends with __syncthreads
                    if (nsteps < 16) {
                        diag_pad_identity<Geo128>(sm, 8 * nsteps, tid);
                        __syncthreads();
                    }
                }
                TICK(4);
            }
            if (sm.fail) return false;
            diag_sums<Geo128>(sm, warp, lane);                          // |z_J|^2 (z_J came out of the sweep), sum log diag
            if (pf) sm.rowv[tid] = pv;                                  // row side of the first block below
            fence_async_smem();
            __syncthreads();
            TICK(5);
            if (tid == 0 && nextI >= 0) prologue<false>(sm, it, lbase + (size_t)(nextI * (nextI - 1) / 2) * BLKD, jrow, niter);
            return true;
        };

        // ---- block (I, J) below the diagonal: S_IJ, L_IJ = S_IJ W^T, ytilde_I -= L_IJ z_J, store L_IJ.
        // staged: sm.rowv holds block row I and its first slabs are in flight (set up by the previous call through nextI).
        auto offdiag_block = [&](const int I, const int J, const int nextI, const bool staged) {
            const double* jrow = lbase + (size_t)(J * (J - 1) / 2) * BLKD;
            const double* irow = lbase + (size_t)(I * (I - 1) / 2) * BLKD;
            const int niter = 8 * J;
            if (!staged) {
                stage_points(sm.rowv, vt, prm.npad_max, I, tid);
                __syncthreads();
                if (tid == 0) prologue<false>(sm, it, irow, jrow, niter);
            }
            // Warp rows of the last block row that hold only identity padding skip everything but the barriers:
            // their rows of L are never read by a warp that works (operands are consumed by the same warp row).
            const int ract = (I == T - 1) ? ((n - I * BS + 31) >> 5) : 4;
            const bool rowact = wr < ract;
            // per-point values of the next block row and this thread's entry of ytilde_I: loaded now, used at the end of the
            // block (a global load in front of a barrier would put its latency on every warp's critical path)
            const bool pf = nextI >= 0 && tid < 3 * BS;
            double pv = 0.0, ycur = 0.0;
            if (pf) pv = vt[(size_t)(tid >> 7) * prm.npad_max + nextI * BS + (tid & (BS - 1))];
            if (tid < 32 * ract) ycur = CL > 1 ? __ldcg(yw + I * BS + tid) : yw[I * BS + tid];   // CL > 1: written by another SM
            if (rowact) {
                if (I * BS + BS <= n && !white_off) synth_tile<EXACT, false>(acc, sm, 32 * wr, 32 * wc, I * BS + 32 * wr, J * BS + 32 * wc, n, vd, white_off, lane);
                else                                synth_tile<EXACT, true>(acc, sm, 32 * wr, 32 * wc, I * BS + 32 * wr, J * BS + 32 * wc, n, vd, white_off, lane);
            }
            TICK(6);
            accumulate<false>(sm, acc, it, irow, jrow, niter, wr, wc, lane, rowact);
            __syncthreads();                                       // stages drained: regionA becomes the X staging block; sm.rowv is free
            TICK(7);
            // (Measured and rejected: letting the warps of the two left warp columns, which leave the block solve 12 - 20 k
            // cycles before the last one, synthesise their tile of the next block row in that time.  Bitwise the same
            // results, the same kernel time: the FP64 pipe is what the whole block is bound by, its order does not matter.)
            if (pf) sm.rowv[tid] = pv;                             // row side of the next block
            if (rowact) {
                trsm_inplace<Geo128, 4>(acc, sm.regionA, sm.W, sm.lpp, wr, wc, lane);      // L_IJ L_JJ^T = S by substitution, acc: -S -> L_IJ
                lz_partials(acc, sm.zs, sm.red, BS, wr, wc, lane);                 // L_IJ z_J, per warp column
                // stored as each warp finishes: warps with little to do overlap their stores with the others' DMMAs
                store_tile_frag<false>(lbase + ((size_t)(I * (I - 1) / 2) + J) * BLKD, acc, 32 * wr, 32 * wc, lane);
            }
            TICK(8);
            fence_async_smem();
            __syncthreads();                                       // staging reads done (regionA free again), partials visible
            TICK(9);
            if (tid == 0 && nextI >= 0)                            // next block's first slabs fly during its synthesis
                prologue<false>(sm, it, lbase + (size_t)(nextI * (nextI - 1) / 2) * BLKD, jrow, niter);
            if (tid < 32 * ract)                                   // ytilde_I -= L_IJ z_J (right-looking forward solve)
                yw[I * BS + tid] = ycur - ((sm.red[tid] + sm.red[BS + tid]) + (sm.red[2 * BS + tid] + sm.red[3 * BS + tid]));
            TICK(10);
        };

        auto column_sums = [&](double& qs, double& ls) {       // fixed order: deterministic
            qs = 0.0; ls = 0.0;
            for (int w = 0; w < 16; w++) { qs += sm.qpart[w]; ls += sm.lpart[w]; }
        };

        if (CL == 1) {
            for (int J = 0; J < T; J++) {
                stage_points(sm.colv, vt, prm.npad_max, J, tid);       // per-point values of block J: column side ...
                stage_points(sm.rowv, vt, prm.npad_max, J, tid);       // ... and row side of the diagonal tile
                __syncthreads();
                if (!diag_block(J, J + 1 < T ? J + 1 : -1)) { failed = true; break; }
                if (tid == 0) {
                    double qs, ls;
                    column_sums(qs, ls);
                    quad += qs;
                    logdet += ls;
                }
                for (int I = J + 1; I < T; I++) offdiag_block(I, J, I + 1 < T ? I + 1 : -1, true);
                // blocks of column J are first read (by TMA, the async proxy) in column J+1: one fence per column
                __threadfence();
                fence_async();
                __syncthreads();
                TICK(11);
            }
        } else {
            // ---- CL CTAs share the problem.  Block rows of a column are handed out by a per-column counter (whoever is
            // free takes the next one); the CTA that takes row J+1 goes on to form the NEXT diagonal block while the
            // others work on column J (look-ahead), publishes W_{J+1}, z_{J+1} and the column's partial sums through
            // global scratch, and comes back for more rows.  Every block is computed by exactly one CTA with the same
            // arithmetic as in the one-CTA kernel, and the sums are added in column order: bitwise the same result.
            //
            // There is no barrier between columns.  Block (I, J) needs the blocks (I, k < J) of its row, the blocks of
            // row J and W_J, z_J: each block row I carries a counter rowdone[I] = number of its columns that are complete
            // and visible, diagdone = the last diagonal block that has been published, both written with release and
            // polled with acquire by thread 0.  A CTA that finds no row left in column J moves on to column J+1 at once
            // and starts on the rows whose column-J blocks are already there.  Waits only ever point to earlier columns
            // (or to a CTA that is computing, not waiting), so they cannot cycle; the cluster launch guarantees that all
            // CTAs of the group are resident.  A failed pivot raises `fail`, which every wait also watches.
            double* const wg = prm.wg + (size_t)slot * prm.wg_stride;             // WBUFS x [W | z], then qcol[], lcol[]
            double* const qcol = wg + WBUFS * WX_DOUBLES;
            double* const lcol = qcol + prm.tcap;
            int* const ctr = prm.cctr + (size_t)slot * cctr_ints(prm.tcap);
            int* const failflag = ctr;                                             // a pivot failed somewhere
            int* const diagdone = ctr + 1;                                         // number of diagonal blocks whose W_J, z_J, sums are published
            int* const rowctr = ctr + 2;                                           // [J] next block row of column J to hand out
            int* const rowdone = ctr + 2 + prm.tcap;                               // [I] columns of block row I complete (incl. ytilde_I)
            int* const coldone = ctr + 2 + 2 * prm.tcap;                           // [J] CTAs that have left column J
            auto wait_for = [&](const int* flag, const int want) -> bool {         // false: the problem failed meanwhile
                if (tid == 0) {
                    int bad = 0;
                    while (ld_acquire(flag) < want && !(bad = ld_acquire(failflag))) __nanosleep(32);
                    sm.prob2 = bad ? 1 : 0;
                }
                __syncthreads();
                const bool ok = sm.prob2 == 0;
                __syncthreads();                                                   // sm.prob2 is reused right away
                return ok;
            };
            auto publish_w = [&](const int buf) {
                double* dst = wg + (size_t)buf * WX_DOUBLES;
                for (int i = tid; i < W_DOUBLES / 2; i += NT) reinterpret_cast<double2*>(dst)[i] = reinterpret_cast<const double2*>(sm.W)[i];
                if (tid < BS) dst[W_DOUBLES + tid] = sm.zs[tid];
                for (int i = tid; i < 16 * 64; i += NT) dst[W_DOUBLES + BS + i] = sm.lpp[i];
            };
            auto fetch_w = [&](const int buf) {
                const double* src = wg + (size_t)buf * WX_DOUBLES;
                for (int i = tid; i < W_DOUBLES / 2; i += NT) reinterpret_cast<double2*>(sm.W)[i] = __ldcg(reinterpret_cast<const double2*>(src) + i);
                if (tid < BS) sm.zs[tid] = __ldcg(src + W_DOUBLES + tid);
                for (int i = tid; i < 16 * 64; i += NT) sm.lpp[i] = __ldcg(src + W_DOUBLES + BS + i);
            };
            // column 0: every CTA forms the first diagonal block itself (no accumulate, nothing to wait for)
            stage_points(sm.colv, vt, prm.npad_max, 0, tid);
            stage_points(sm.rowv, vt, prm.npad_max, 0, tid);
            __syncthreads();
            if (!diag_block(0, -1)) {
                failed = true;
                if (tid == 0) { atomicExch(failflag, 1); sm.fail = 0; }
            }
            if (!failed && crank == 0) {                                       // W_0 for the CTA that comes back to column 0 after
                if (tid == 0) column_sums(qcol[0], lcol[0]);                    // forming the second diagonal block
                publish_w(0);
                __threadfence();
                __syncthreads();
                if (tid == 0) st_release(diagdone, 1);
            }
            __syncthreads();
            for (int J = 0; J < T && !failed; J++) {
                if (J > 0) {
                    if (!wait_for(diagdone, J + 1)) { failed = true; break; }  // W_J, z_J are there (and row J is complete)
                    fetch_w(J % WBUFS);
                    stage_points(sm.colv, vt, prm.npad_max, J, tid);
                    __syncthreads();
                }
                for (;;) {
                    if (tid == 0) sm.prob2 = atomicAdd(rowctr + J, 1);
                    __syncthreads();
                    const int I = J + 1 + sm.prob2;
                    __syncthreads();
                    if (I >= T) break;
                    if (!wait_for(rowdone + I, J)) { failed = true; break; }   // blocks (I, 0 .. J-1) and ytilde_I through column J-1
                    if (tid == 0) fence_async();                               // ... before this CTA's TMA reads them
                    offdiag_block(I, J, -1, false);
                    __threadfence();                                           // L_IJ and ytilde_I, then the row's counter
                    fence_async();
                    __syncthreads();
                    if (tid == 0) st_release(rowdone + I, J + 1);
                    if (I == J + 1) {
                        stage_points(sm.colv, vt, prm.npad_max, J + 1, tid);
                        stage_points(sm.rowv, vt, prm.npad_max, J + 1, tid);
                        __syncthreads();
                        const bool ok = diag_block(J + 1, -1);
                        if (ok) {
                            // buffer (J+1) % WBUFS last held W_{J+1-WBUFS}: every CTA must have left that column
                            if (J + 1 >= WBUFS && !wait_for(coldone + (J + 1 - WBUFS), CL)) { failed = true; break; }
                            if (tid == 0) column_sums(qcol[J + 1], lcol[J + 1]);
                            publish_w((J + 1) % WBUFS);
                            __threadfence();
                            __syncthreads();
                            if (tid == 0) st_release(diagdone, J + 2);
                        } else {
                            if (tid == 0) { atomicExch(failflag, 1); sm.fail = 0; }
                            failed = true;
                            break;
                        }
                        if (J + 2 < T) {                                       // back to column J: its W, z and column points
                            if (!wait_for(diagdone, J + 1)) { failed = true; break; }   // (J = 0: rank 0 publishes W_0)
                            fetch_w(J % WBUFS);
                            stage_points(sm.colv, vt, prm.npad_max, J, tid);
                        }
                        __syncthreads();
                    }
                }
                if (failed) break;
                if (tid == 0) atomicAdd(coldone + J, 1);                        // this CTA no longer needs W_J
                TICK(11);
            }
            if (crank == 0) {                                                   // rank 0 reports: wait for the last diagonal block
                if (!failed && !wait_for(diagdone, T)) failed = true;
                if (!failed && ld_acquire(failflag)) failed = true;
                if (tid == 0 && !failed)
                    for (int J = 0; J < T; J++) { quad += __ldcg(qcol + J); logdet += __ldcg(lcol + J); }
            }
        }
        if (tid == 0 && crank == 0) prm.prob_out[prob] = finish_lnlike(quad, logdet, n, failed);
    }
#ifdef GPROT_TIMING
    if (tid == 0 && blockIdx.x == 0 && prm.timing)
        for (int k = 0; k < 16; k++) prm.timing[k] = tim_[k];
#endif
}

#ifdef GPROT_API_TU   // small non-template kernels: compiled once, in the translation unit of the C ABI
// out[b] = sum over the row's problems, in a fixed order (gprot/model.py:362-365 np.sum over chunks)
__global__ void reduce_rows_kernel(const double* __restrict__ prob_out, const int* __restrict__ row_first,
                                   const int* __restrict__ row_count, double* __restrict__ out, long long B) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int f = row_first[b], c = row_count[b];
    double s;
    if (c <= 0) s = -INFINITY;
    else {
        s = 0.0;
        for (int i = 0; i < c; i++) s += prob_out[f + i];
        if (!(fabs(s) <= DBL_MAX)) s = -INFINITY;
    }
    out[b] = s;
}

// register-resident DMMA loop: the roofline denominator (same loop as tools/microbench.cu k_dmma<8>)
__global__ void dmma_peak_kernel(double* out, int iters, long long* cyc) {
    double acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i][0] = threadIdx.x * 1e-9; acc[i][1] = i; }
    const double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
    const long long t0 = clock64();
    for (int k = 0; k < iters; k++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma(acc[i][0], acc[i][1], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

#endif

}  // namespace gprot
