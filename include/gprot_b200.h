/* gprot_b200.h -- C ABI of the B200-native quasi-periodic GP log-likelihood engine.
 *
 * This is the drop-in boundary for ONE path of RuthAngus/GProtation: the likelihood that
 * emcee3 evaluates for every walker proposal.  Each entry point cites the reference
 * interface it replaces (paths relative to the reference repository root).  The reference
 * is pure Python on top of george; the binding a maintainer adds is the ctypes stub shown
 * in INTEGRATION.md (gprotation_b200/_lib.py is that stub, shipped).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross this boundary;
 *   - every function returns 0 on success, non-zero on argument/CUDA errors
 *     (text via gprot_last_error); NUMERICAL failure is data: the affected output is -inf,
 *     exactly like gprot/model.py:349-356 maps LinAlgError / non-finite values to -inf;
 *   - theta rows are (ln A, ln l, ln G, ln sigma, ln P), natural logs (gprot/model.py:42);
 *   - a context is bound to one CUDA device, calls on one context are stream-ordered and
 *     not re-entrant (SURVEY.md section 8b "Threading").
 *   - there is NO CPU fallback: without a CUDA device gprot_create fails.
 */
#ifndef GPROT_B200_H
#define GPROT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gprot_ctx gprot_ctx;

/* flags for gprot_lnlike_batch */
#define GPROT_THETA_ON_DEVICE 1u /* theta points to device memory                         */
#define GPROT_OUT_ON_DEVICE   2u /* out points to device memory (no D2H, no sync)         */
#define GPROT_EXACT_KERNEL    4u /* evaluate every covariance entry in george's operation
                                    order (sin(M_PI*d/P), two exps, divisions) instead of
                                    the fast per-point-phase form; ~3x slower synthesis   */

/* ABI / build identification. */
int gprot_abi_version(void);

/* Replaces the per-process george/solver state that each schwimmbad worker holds
 * (scripts/gprot-fit:195, gprot/fit.py:94-98).  device = CUDA ordinal. */
int gprot_create(int device, gprot_ctx **ctx);
int gprot_destroy(gprot_ctx *ctx);
const char *gprot_last_error(const gprot_ctx *ctx);

/* Registers one light curve made of m chunks (m = 1: the --nochunks single N x N problem;
 * m > 1: lc.x_list / y_list / yerr_list of gprot/lc.py:329-342, 375-391).  n[m] are the chunk
 * lengths; t, y, yerr hold the chunks back to back (sum(n) doubles each, host memory).
 * The arrays are copied to the device once (the reference re-sends them to george.GP.compute on
 * every call, gprot/model.py:343-346).  Returns an id through lc_id. */
int gprot_register_lc(gprot_ctx *ctx, int m, const int *n, const double *t, const double *y,
                      const double *yerr, int *lc_id);
/* Forget ONE light curve (its id is not reused; the arena space is reclaimed by compaction once dead points outweigh
 * live ones) or ALL of them (ids restart at 0).  gprot_generation counts the gprot_unregister_all calls so that a caller
 * that caches ids -- GPRotModel caches the id of its LightCurve -- can tell that they are void. */
int gprot_unregister_lc(gprot_ctx *ctx, int lc_id);
int gprot_unregister_all(gprot_ctx *ctx);
int64_t gprot_generation(const gprot_ctx *ctx);

/* GPRotModel.lnlike(theta) for B proposals in ONE launch (gprot/model.py:358-365 called
 * per walker from Emcee3Model.compute_log_likelihood, gprot/fit.py:20-22):
 *   out[b] = sum over chunks c of lc_id[b] of lnlike_function(theta[b], x_c, y_c, yerr_c)
 * theta is [B,5] row-major; lc_id is [B] (NULL: every row uses light curve 0);
 * stream is a cudaStream_t cast to void* (NULL: the legacy default stream).
 * Without GPROT_OUT_ON_DEVICE the call copies out to the host and synchronises. */
int gprot_lnlike_batch(gprot_ctx *ctx, int64_t B, const double *theta, const int32_t *lc_id,
                       double *out, unsigned flags, void *stream);

/* GPRotModel.lnprior(theta) (gprot/model.py:132-163) for B rows, evaluated on the host side of
 * the boundary (it is 20 flops): bounds[10] = (lo,hi) x 5, mu[4], sigma[4], mixture = k rows of
 * (w, mu, sigma) or NULL when acf_prior is off. */
int gprot_lnprior_batch(int64_t B, const double *theta, const double *bounds, const double *mu,
                        const double *sigma, double pmin, double pmax, const double *mixture, int k,
                        double *out);

/* GPRotModel.lnpost(theta) = lnlike + lnprior (gprot/model.py:367-369) for B rows.  Rows whose
 * prior is -inf are not factorised (emcee3 never asks for their likelihood); they return -inf. */
int gprot_lnpost_batch(gprot_ctx *ctx, int64_t B, const double *theta, const int32_t *lc_id,
                       const double *bounds, const double *mu, const double *sigma, double pmin,
                       double pmax, const double *mixture, int k, double *out_lnprior,
                       double *out_lnlike, unsigned flags, void *stream);

int gprot_sync(gprot_ctx *ctx);

/* Kernel choice.  Launches with more than one problem per SM and N < 1920 run two problems per SM (64 x 64 blocks,
 * gp_kernel64.cuh); larger problems run one per SM (128 x 128 blocks).  A launch with fewer problems than SMs
 * (32 problems per GPU for the N=8192 stress at 8 GPUs; one emcee3 half-ensemble of a single star, gprot/fit.py:98) spreads each problem over a
 * thread-block cluster of 2, 4 or 8 CTAs.  0 (default) chooses per launch from the batch size; results are bitwise
 * identical for 1, 2, 4 and 8.  64 forces the two-problems-per-SM kernel, 1 the one-problem-per-SM kernel (the two
 * blockings agree to fp64 rounding, ~eps * cond(K)).  gprot_last_cluster reports what the last launch used. */
int gprot_set_cluster(gprot_ctx *ctx, int ctas_per_problem);
int gprot_last_cluster(const gprot_ctx *ctx);

/* Several GPUs from ONE process (replaces the reference's one-process-per-core MultiPool /
 * MPIPool fan-out of scripts/gprot-fit:127-131,195 when a single driver process owns the box).  A group holds one
 * context per listed device; light curves are replicated to every device; gprot_group_lnlike_batch gives device r
 * the r-th contiguous block of rows (star-major row lists keep a star's walkers together), drives the devices
 * concurrently and gathers into the caller's host buffer -- no device-to-device traffic.  One process per GPU with
 * a distributed launcher (gprotation_b200/dist.py, NCCL all-gather of the log-probabilities) is the multi-process form. */
typedef struct gprot_group gprot_group;
int gprot_group_create(int ndev, const int *devices, gprot_group **grp);
int gprot_group_destroy(gprot_group *grp);
int gprot_group_size(const gprot_group *grp);
const char *gprot_group_last_error(const gprot_group *grp);
int gprot_group_register_lc(gprot_group *grp, int m, const int *n, const double *t, const double *y,
                            const double *yerr, int *lc_id);
int gprot_group_unregister_all(gprot_group *grp);
int gprot_group_unregister_lc(gprot_group *grp, int lc_id);
int gprot_group_lnlike_batch(gprot_group *grp, int64_t B, const double *theta, const int32_t *lc_id,
                             double *out, unsigned flags);
/* gprot_lnpost_batch over the devices of the group (prior on the host, rows outside its support not factorised): what a
 * sampler step calls when GPRotModel runs on a GPGroup (gprot/fit.py:16-22 through pool.map, gprot/model.py:367-369). */
int gprot_group_lnpost_batch(gprot_group *grp, int64_t B, const double *theta, const int32_t *lc_id,
                             const double *bounds, const double *mu, const double *sigma, double pmin,
                             double pmax, const double *mixture, int k, double *out_lnprior,
                             double *out_lnlike, unsigned flags);

/* The ensemble sampler of the reference's fit driver on the host side of the boundary (gprot/fit.py:82-98,123-135: an
 * emcee3 Ensemble over Emcee3Model, Sampler(moves).run(ensemble, n) with KDEMove 0.4 / DEMove(1.0) 0.4 / DESnookerMove 0.2).
 * Same red-blue split, moves and acceptance rule as gprotation_b200/sampler.py, without the interpreter between the two
 * likelihood launches of a step.  Weights <= 0 switch a move off; parameters <= 0 take the defaults of sampler.py
 * (stretch a = 2, DE gamma0 = 2.38 / sqrt(2 ndim), snooker gamma = 1.7, KDE bandwidth = Scott's factor).
 * coords [nwalkers, ndim], lnprior / lnlike [nwalkers] hold the ensemble on entry (every walker finite) and on return;
 * chain [nsteps / thin, nwalkers, ndim], chain_lnprior / chain_lnlike [nsteps / thin, nwalkers] (any of them NULL: not
 * stored) receive the ensemble after every thin-th step, naccepted [nwalkers] is incremented on those steps.
 * The random stream is the library's own (xoshiro256++): a native chain and a Python-driven one are different
 * realisations of the same Markov chain.  Errors: non-zero status, text from gprot_chain_last_error (thread local). */
typedef void (*gprot_lnpost_fn)(void *user, int64_t n, const double *theta, double *lnprior, double *lnlike);
typedef struct gprot_chain_moves {
    double w_stretch, w_de, w_snooker, w_kde;                     /* mixture weights (normalised by the library) */
    double stretch_a, de_sigma, de_gamma0, snooker_gamma, kde_bw; /* move parameters */
} gprot_chain_moves;
/* any log-posterior: fn evaluates n proposals at a time (tests drive it with a Gaussian) */
int gprot_chain_run_fn(gprot_lnpost_fn fn, void *user, int nwalkers, int ndim, int64_t nsteps, int thin,
                       const gprot_chain_moves *moves, uint64_t seed, double *coords, double *lnprior, double *lnlike,
                       double *chain, double *chain_lnprior, double *chain_lnlike, int64_t *naccepted);
/* GPRotModel.lnpost of light curve lc_id through gprot_lnpost_batch (prior arguments as there), ndim = 5 */
int gprot_chain_run(gprot_ctx *ctx, int lc_id, const double *bounds, const double *mu, const double *sigma, double pmin,
                    double pmax, const double *mixture, int k, int nwalkers, int64_t nsteps, int thin,
                    const gprot_chain_moves *moves, uint64_t seed, double *coords, double *lnprior, double *lnlike,
                    double *chain, double *chain_lnprior, double *chain_lnlike, int64_t *naccepted);
const char *gprot_chain_last_error(void);

/* Measurement hooks (bench.py): device time of the last likelihood launch, CUDA events on the
 * launching stream; number of kernels launched by this context so far; FP64 tensor (DMMA) peak
 * of this device measured by a register-resident mma.sync.m8n8k4.f64 loop. */
int gprot_last_kernel_ms(gprot_ctx *ctx, float *ms);
int64_t gprot_launch_count(const gprot_ctx *ctx);
int gprot_measure_dmma_peak(gprot_ctx *ctx, double *tflops, double *sm_clock_mhz);

/* Test hook: copy `count` doubles of work-slot `slot`'s factor scratch to the host
 * (fragment-major 128 x 128 blocks, strictly-lower block rows). */
int gprot_debug_read_scratch(gprot_ctx *ctx, int slot, size_t offset, size_t count, double *host_out);

/* Test hook: per-phase SM-cycle totals of CTA 0 for the last launch; all zero unless the library was built with
 * -DGPROT_TIMING (tools/dev_timing.py builds that variant next to the product library). */
int gprot_debug_read_timing(gprot_ctx *ctx, long long *host_out16);

#ifdef __cplusplus
}
#endif
#endif /* GPROT_B200_H */
