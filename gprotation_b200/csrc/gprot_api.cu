// gprot_api.cu -- C ABI (include/gprot_b200.h) over the sm_100a kernels in gp_kernel.cuh.
// Host side of the boundary: light-curve arena, problem queue, work-slot scratch, launches.
// No CPU fallback: every numeric result comes from gp_lnlike_kernel.
#include "../../include/gprot_b200.h"
#define GPROT_API_TU
#include "gp_kernel.cuh"
#include "gp_kernel64.cuh"

// the likelihood kernels are instantiated in gp_kernel_inst.cu (one object per variant, built in parallel)
namespace gprot {
#define GPROT_EXTERN_KERNEL(E, C) extern template __global__ void gp_lnlike_kernel<E, C>(const Params);
GPROT_EXTERN_KERNEL(false, 1) GPROT_EXTERN_KERNEL(false, 2) GPROT_EXTERN_KERNEL(false, 4) GPROT_EXTERN_KERNEL(false, 8)
GPROT_EXTERN_KERNEL(true, 1) GPROT_EXTERN_KERNEL(true, 2) GPROT_EXTERN_KERNEL(true, 4) GPROT_EXTERN_KERNEL(true, 8)
#undef GPROT_EXTERN_KERNEL
namespace k64 {
extern template __global__ void gp_lnlike64_kernel<false>(const Params);
extern template __global__ void gp_lnlike64_kernel<true>(const Params);
}
}  // namespace gprot

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace gprot;

struct gprot_ctx;

// one persistent host thread per device of a group: gprot_group_* calls hand it a closure and wait (no thread is
// created or joined per call)
struct group_worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = false, quit = false;
    int rc = 0;
    void loop() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return has_job || quit; });
            if (quit) return;
            std::function<int()> f = std::move(job);
            has_job = false;
            lk.unlock();
            const int r = f();
            lk.lock();
            rc = r;
            done = true;
            cv.notify_all();
        }
    }
    void submit(std::function<int()> f) {
        std::lock_guard<std::mutex> lk(mu);
        job = std::move(f);
        has_job = true;
        done = false;
        cv.notify_all();
    }
    int wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done; });
        return rc;
    }
    void stop() {
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
            cv.notify_all();
        }
        if (th.joinable()) th.join();
    }
};

struct gprot_group {
    std::vector<gprot_ctx*> ctxs;
    std::vector<group_worker*> workers;
    std::string err;
};

struct gprot_ctx {
    int device = 0;
    int num_sms = 0;
    std::string err;
    // host copy of the arena
    std::vector<double> h_t, h_y, h_yerr;
    std::vector<long long> h_chunk_off;
    std::vector<int> h_chunk_n;
    std::vector<double> h_chunk_mingap2;
    std::vector<int> lc_first_chunk, lc_nchunks;        // lc_nchunks[id] == 0: unregistered (ids are never reused until unregister_all)
    size_t dead_points = 0;                             // arena points that belong to unregistered light curves
    int64_t generation = 0;                             // bumped whenever previously returned ids stop being valid
    // device copies: grown geometrically, only the new tail is uploaded after a registration
    double *d_t = nullptr, *d_y = nullptr, *d_yerr = nullptr, *d_chunk_mingap2 = nullptr;
    long long* d_chunk_off = nullptr;
    int* d_chunk_n = nullptr;
    size_t cap_points = 0, up_points = 0, cap_chunks = 0, up_chunks = 0;
    // pinned host staging for the problem list of a call (reused; ev_stage marks the end of its last upload)
    int* h_stage = nullptr; size_t cap_stage = 0;
    cudaEvent_t ev_stage = nullptr; bool stage_pending = false;
    int max_lslots = 0;                                 // > 0: the factor scratch was capped at this many slots by device memory
    // per-call buffers (grown on demand)
    double* d_theta = nullptr; size_t cap_theta = 0;
    double* d_out = nullptr; size_t cap_out = 0;
    int *d_prob_theta = nullptr, *d_prob_chunk = nullptr, *d_prob_order = nullptr; size_t cap_prob = 0;
    bool use_order = false;           // the last problem list has mixed sizes and is queued largest first
    double* d_prob_out = nullptr;
    int *d_row_first = nullptr, *d_row_count = nullptr; size_t cap_rows = 0;
    int* d_counter = nullptr;
    long long* d_timing = nullptr;
    // work-slot scratch
    double* d_lscratch = nullptr; size_t lslot_stride = 0; int lslots = 0;
    double* d_vscratch = nullptr; int v_npad = 0; int vslots = 0;
    double* d_wg = nullptr; int* d_cctr = nullptr; int cslots = 0;      // cluster kernels: W exchange + row counters
    // problem-list cache: the device list of the last call without a mask whose rows all use ONE light curve
    long long cache_B = -1; int cache_lc = -1; int cache_nprob = 0; int cache_tmax = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
    int64_t launches = 0;
    bool attr_set = false;
    int force_cluster = 0;            // 0 = choose per launch; 1, 2, 4, 8 = CTAs per problem; 64 = two problems per SM
    bool attr64_set = false;
    int last_nmax = 0;                // largest chunk of the cached problem list
    int max_clusters[9] = {0};        // co-resident clusters of each size (cudaOccupancyMaxActiveClusters), 0 = unknown
    int last_cluster = 1;
};

namespace {

thread_local std::string g_err;

int fail(gprot_ctx* ctx, const std::string& msg) {
    if (ctx) ctx->err = msg;
    g_err = msg;
    return 1;
}
#define CU(ctx, call)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(ctx, std::string(#call) + ": " + cudaGetErrorString(e_));             \
    } while (0)

template <class T>
int grow(gprot_ctx* ctx, T** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) CU(ctx, cudaFree(*p));
    *p = nullptr;
    size_t c = std::max(need, *cap * 2);
    CU(ctx, cudaMalloc((void**)p, c * sizeof(T)));
    *cap = c;
    return 0;
}

// device arena = host arena: grow geometrically (device-to-device copy of what is already there), upload only the tail
// appended since the last call.  Kernels in flight only read below up_points / up_chunks.
template <class T>
int grow_keep(gprot_ctx* ctx, T** p, size_t old_cap, size_t new_cap, size_t used) {
    T* q = nullptr;
    CU(ctx, cudaMalloc((void**)&q, std::max<size_t>(new_cap, 2) * sizeof(T)));
    if (*p && used) CU(ctx, cudaMemcpy(q, *p, used * sizeof(T), cudaMemcpyDeviceToDevice));
    if (*p) CU(ctx, cudaFree(*p));
    (void)old_cap;
    *p = q;
    return 0;
}

int upload_arena(gprot_ctx* ctx) {
    const size_t n = ctx->h_t.size(), m = ctx->h_chunk_n.size();
    if (n == ctx->up_points && m == ctx->up_chunks) return 0;
    if (n > ctx->cap_points) {
        CU(ctx, cudaDeviceSynchronize());                       // nothing may still read the old buffers when they are freed
        const size_t c = std::max<size_t>(std::max(n, 2 * ctx->cap_points), 1 << 14);
        if (grow_keep(ctx, &ctx->d_t, ctx->cap_points, c, ctx->up_points) || grow_keep(ctx, &ctx->d_y, ctx->cap_points, c, ctx->up_points) ||
            grow_keep(ctx, &ctx->d_yerr, ctx->cap_points, c, ctx->up_points)) return 1;
        ctx->cap_points = c;
    }
    if (m > ctx->cap_chunks) {
        CU(ctx, cudaDeviceSynchronize());
        const size_t c = std::max<size_t>(std::max(m, 2 * ctx->cap_chunks), 256);
        if (grow_keep(ctx, &ctx->d_chunk_off, ctx->cap_chunks, c, ctx->up_chunks) || grow_keep(ctx, &ctx->d_chunk_n, ctx->cap_chunks, c, ctx->up_chunks) ||
            grow_keep(ctx, &ctx->d_chunk_mingap2, ctx->cap_chunks, c, ctx->up_chunks)) return 1;
        ctx->cap_chunks = c;
    }
    const size_t p0 = ctx->up_points, c0 = ctx->up_chunks;
    if (n > p0) {
        CU(ctx, cudaMemcpy(ctx->d_t + p0, ctx->h_t.data() + p0, (n - p0) * 8, cudaMemcpyHostToDevice));
        CU(ctx, cudaMemcpy(ctx->d_y + p0, ctx->h_y.data() + p0, (n - p0) * 8, cudaMemcpyHostToDevice));
        CU(ctx, cudaMemcpy(ctx->d_yerr + p0, ctx->h_yerr.data() + p0, (n - p0) * 8, cudaMemcpyHostToDevice));
    }
    if (m > c0) {
        CU(ctx, cudaMemcpy(ctx->d_chunk_off + c0, ctx->h_chunk_off.data() + c0, (m - c0) * 8, cudaMemcpyHostToDevice));
        CU(ctx, cudaMemcpy(ctx->d_chunk_n + c0, ctx->h_chunk_n.data() + c0, (m - c0) * 4, cudaMemcpyHostToDevice));
        CU(ctx, cudaMemcpy(ctx->d_chunk_mingap2 + c0, ctx->h_chunk_mingap2.data() + c0, (m - c0) * 8, cudaMemcpyHostToDevice));
    }
    ctx->up_points = n;
    ctx->up_chunks = m;
    return 0;
}

// drop the arena points of unregistered light curves once they outweigh the live ones: ids keep their meaning, the
// chunk tables are rewritten and everything is uploaded again (rare: amortised over the registrations that caused it)
int compact_arena(gprot_ctx* ctx) {
    CU(ctx, cudaDeviceSynchronize());
    std::vector<double> t, y, e, g2;
    std::vector<long long> off;
    std::vector<int> cn;
    for (size_t id = 0; id < ctx->lc_first_chunk.size(); id++) {
        const int f = ctx->lc_first_chunk[id], m = ctx->lc_nchunks[id];
        if (m == 0) continue;
        ctx->lc_first_chunk[id] = (int)cn.size();
        for (int c = f; c < f + m; c++) {
            const size_t o0 = (size_t)ctx->h_chunk_off[c], len = (size_t)ctx->h_chunk_n[c] + ((size_t)ctx->h_chunk_n[c] & 1);
            off.push_back((long long)t.size());
            cn.push_back(ctx->h_chunk_n[c]);
            g2.push_back(ctx->h_chunk_mingap2[c]);
            t.insert(t.end(), ctx->h_t.begin() + o0, ctx->h_t.begin() + o0 + len);
            y.insert(y.end(), ctx->h_y.begin() + o0, ctx->h_y.begin() + o0 + len);
            e.insert(e.end(), ctx->h_yerr.begin() + o0, ctx->h_yerr.begin() + o0 + len);
        }
    }
    ctx->h_t.swap(t); ctx->h_y.swap(y); ctx->h_yerr.swap(e);
    ctx->h_chunk_off.swap(off); ctx->h_chunk_n.swap(cn); ctx->h_chunk_mingap2.swap(g2);
    ctx->up_points = 0; ctx->up_chunks = 0; ctx->dead_points = 0;
    ctx->cache_B = -1;
    return 0;
}

// gprot/model.py:22-23
inline double ln_gauss(double x, double mu, double sigma) {
    return -0.5 * ((x - mu) * (x - mu) / (sigma * sigma)) + std::log(1.0 / std::sqrt(2.0 * M_PI * sigma * sigma));
}

// gprot/model.py:132-163
double lnprior_one(const double* th, const double* bounds, const double* mu, const double* sigma, double pmin,
                   double pmax, const double* mix, int k) {
    const double ninf = -std::numeric_limits<double>::infinity();
    for (int i = 0; i < 5; i++)
        if (th[i] < bounds[2 * i] || th[i] > bounds[2 * i + 1] || std::isnan(th[i])) return ninf;
    if (th[4] < pmin || th[4] > pmax) return ninf;
    if (th[1] < th[4]) return ninf;
    double lnpr = 0.0;
    for (int i = 0; i < 4; i++) lnpr += ln_gauss(th[i], mu[i], sigma[i]);
    if (mix && k > 0) {   // logsumexp(lnGauss(p, mu_j, sigma_j), b=w_j), gprot/model.py:25-29
        double mx = ninf;
        std::vector<double> a(k);
        for (int j = 0; j < k; j++) {
            a[j] = ln_gauss(th[4], mix[3 * j + 1], mix[3 * j + 2]);
            mx = std::max(mx, a[j]);
        }
        if (!std::isfinite(mx)) mx = 0.0;
        double s = 0.0;
        for (int j = 0; j < k; j++) s += mix[3 * j] * std::exp(a[j] - mx);
        lnpr += std::log(s) + mx;
    }
    return lnpr;
}

typedef void (*kernel_fn)(const Params);
kernel_fn kernel_for(bool exact, int cl) {
    switch (cl) {
        case 2: return exact ? gp_lnlike_kernel<true, 2> : gp_lnlike_kernel<false, 2>;
        case 4: return exact ? gp_lnlike_kernel<true, 4> : gp_lnlike_kernel<false, 4>;
        case 8: return exact ? gp_lnlike_kernel<true, 8> : gp_lnlike_kernel<false, 8>;
        default: return exact ? gp_lnlike_kernel<true, 1> : gp_lnlike_kernel<false, 1>;
    }
}

// Relative time of one problem of T block columns on a cluster of cl CTAs, in units of one 128^3 DMMA block product
// (weights from the phase budget in DESIGN.md section 4, fitted to profiles/r01_cluster_sweep.log): the work is shared,
// each column loses about half a block row to the granularity of the row queue, and pays a cluster barrier plus the
// fetch of W.
double problem_time_model(int T, int cl) {
    double total = 0.0, gran = 0.0;
    for (int J = 0; J < T; J++) {
        total += 0.6 * J + 4.0 + (double)(T - 1 - J) * (J + 2.0);
        if (J < T - 1) gran += J + 2.0;
    }
    if (cl == 1) return total;
    return total / cl + gran * (cl - 1) / (2.0 * cl) + T * 0.75 * cl;
}

int query_clusters(gprot_ctx* ctx) {
    if (ctx->max_clusters[1]) return 0;
    ctx->max_clusters[1] = ctx->num_sms;
    for (int cl : {2, 4, 8}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(ctx->num_sms / cl * cl));
        cfg.blockDim = dim3(NT);
        cfg.dynamicSmemBytes = sizeof(Smem);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, (void*)kernel_for(false, cl), &cfg);
        if (e != cudaSuccess) { cudaGetLastError(); nc = 0; }
        ctx->max_clusters[cl] = std::min(nc, ctx->num_sms / cl);
    }
    return 0;
}

// CTAs per problem for a launch of nprob problems whose largest has tmax block columns: the size that minimises
// rounds x modelled time per problem.  Batches that fill the GPU (nprob >= SMs) with short tails stay at 1.
int choose_cluster(gprot_ctx* ctx, int nprob, int tmax) {
    if (ctx->force_cluster == 64) return 1;
    if (ctx->force_cluster) return ctx->max_clusters[ctx->force_cluster] > 0 ? ctx->force_cluster : 1;
    if (tmax < 4) return 1;                                  // measured: three block columns do not pay for the cluster barriers
    int best = 1;
    double best_t = 0.0;
    for (int cl : {1, 2, 4}) {                               // 8 never wins (too few block rows per column); it can be forced
        const int nc = ctx->max_clusters[cl];
        if (nc <= 0) continue;
        const double rounds = (double)((nprob + nc - 1) / nc);
        const double t = rounds * problem_time_model(tmax, cl);
        if (cl == 1 || t < 0.97 * best_t) { best = cl; best_t = t; }
    }
    return best;
}

// (theta row, chunk) problems of a launch, their per-row ranges for the chunk sum, and the queue order.  The list lives on
// the device; it is rebuilt only when the call differs from the cached one (same B, every row on the same light curve,
// no mask -- what a sampler step or a bench loop repeats).  A rebuild fills pinned host staging owned by the context
// and uploads it asynchronously on the caller's stream: no stream synchronisation, no pageable copies.
struct ProblemList { int nprob = 0, tmax = 0, nmax = 0; };

int build_problem_list(gprot_ctx* ctx, int64_t B, const int32_t* lc_id, const unsigned char* mask, cudaStream_t st, ProblemList& pl) {
    int& nprob = pl.nprob; int& tmax = pl.tmax; int& nmax = pl.nmax;
    const int nlc = (int)ctx->lc_first_chunk.size();
    int uniform = lc_id ? lc_id[0] : 0;                      // every row on one light curve?
    if (lc_id)
        for (int64_t b = 1; b < B && uniform >= 0; b++)
            if (lc_id[b] != uniform) uniform = -1;
    const bool cacheable = (uniform >= 0 && mask == nullptr);
    if (cacheable && ctx->cache_B == B && ctx->cache_lc == uniform) {
        nprob = ctx->cache_nprob;
        tmax = ctx->cache_tmax;
        nmax = ctx->last_nmax;
        return 0;
    }
    // ---- size of the list
    size_t np = 0;
    for (int64_t b = 0; b < B; b++) {
        const int lc = lc_id ? lc_id[b] : 0;
        if (lc < 0 || lc >= nlc) return fail(ctx, "lc_id out of range");
        if (ctx->lc_nchunks[lc] == 0) return fail(ctx, "lc_id refers to an unregistered light curve");
        if (!mask || mask[b]) np += (size_t)ctx->lc_nchunks[lc];
    }
    if (np > (size_t)std::numeric_limits<int>::max()) return fail(ctx, "too many (row, chunk) problems in one call");
    // ---- pinned staging: [row_first B][row_count B][prob_theta np][prob_chunk np][order np]
    const size_t need = 2 * (size_t)B + 3 * np;
    if (ctx->stage_pending) {                                  // the previous list may still be on its way to the device
        CU(ctx, cudaEventSynchronize(ctx->ev_stage));
        ctx->stage_pending = false;
    }
    if (need > ctx->cap_stage) {
        if (ctx->h_stage) CU(ctx, cudaFreeHost(ctx->h_stage));
        ctx->h_stage = nullptr;
        const size_t c = std::max(need, 2 * ctx->cap_stage);
        CU(ctx, cudaMallocHost((void**)&ctx->h_stage, c * sizeof(int)));
        ctx->cap_stage = c;
    }
    int* rfirst = ctx->h_stage; int* rcount = rfirst + B; int* ptheta = rcount + B; int* pchunk = ptheta + np; int* order = pchunk + np;
    int k = 0, nmin = std::numeric_limits<int>::max();
    for (int64_t b = 0; b < B; b++) {
        const int lc = lc_id ? lc_id[b] : 0;
        rfirst[b] = k;
        if (mask && !mask[b]) { rcount[b] = 0; continue; }
        const int f = ctx->lc_first_chunk[lc], m = ctx->lc_nchunks[lc];
        rcount[b] = m;
        for (int c = 0; c < m; c++, k++) {
            ptheta[k] = (int)b;
            pchunk[k] = f + c;
            const int v = ctx->h_chunk_n[f + c];
            nmax = std::max(nmax, v);
            nmin = std::min(nmin, v);
        }
    }
    nprob = k;
    tmax = (nmax + BS - 1) / BS;
    if ((size_t)B > ctx->cap_rows) {
        CU(ctx, cudaStreamSynchronize(st));                     // a queued launch may still read the old arrays
        if (ctx->d_row_first) CU(ctx, cudaFree(ctx->d_row_first));
        if (ctx->d_row_count) CU(ctx, cudaFree(ctx->d_row_count));
        ctx->d_row_first = ctx->d_row_count = nullptr;
        const size_t c = std::max<size_t>((size_t)B, ctx->cap_rows * 2);
        CU(ctx, cudaMalloc((void**)&ctx->d_row_first, c * 4));
        CU(ctx, cudaMalloc((void**)&ctx->d_row_count, c * 4));
        ctx->cap_rows = c;
    }
    if ((size_t)nprob > ctx->cap_prob) {
        CU(ctx, cudaStreamSynchronize(st));
        if (ctx->d_prob_theta) CU(ctx, cudaFree(ctx->d_prob_theta));
        if (ctx->d_prob_chunk) CU(ctx, cudaFree(ctx->d_prob_chunk));
        if (ctx->d_prob_out) CU(ctx, cudaFree(ctx->d_prob_out));
        if (ctx->d_prob_order) CU(ctx, cudaFree(ctx->d_prob_order));
        ctx->d_prob_theta = ctx->d_prob_chunk = ctx->d_prob_order = nullptr; ctx->d_prob_out = nullptr;
        const size_t c = std::max<size_t>(nprob, ctx->cap_prob * 2);
        CU(ctx, cudaMalloc((void**)&ctx->d_prob_theta, c * 4));
        CU(ctx, cudaMalloc((void**)&ctx->d_prob_chunk, c * 4));
        CU(ctx, cudaMalloc((void**)&ctx->d_prob_out, c * 8));
        CU(ctx, cudaMalloc((void**)&ctx->d_prob_order, c * 4));
        ctx->cap_prob = c;
    }
    CU(ctx, cudaMemcpyAsync(ctx->d_row_first, rfirst, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(ctx->d_row_count, rcount, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    ctx->use_order = false;
    if (nprob > 0) {
        CU(ctx, cudaMemcpyAsync(ctx->d_prob_theta, ptheta, (size_t)nprob * 4, cudaMemcpyHostToDevice, st));
        CU(ctx, cudaMemcpyAsync(ctx->d_prob_chunk, pchunk, (size_t)nprob * 4, cudaMemcpyHostToDevice, st));
        // Mixed problem sizes: the persistent CTAs take the largest problems first (longest-processing-time-first), so
        // that a big problem is never the last one popped.  Outputs keep their listed positions.
        if (nprob > 1 && (nmax + BS - 1) / BS != (nmin + BS - 1) / BS) {
            for (int i = 0; i < nprob; i++) order[i] = i;
            std::stable_sort(order, order + nprob, [&](int a, int b) { return ctx->h_chunk_n[pchunk[a]] > ctx->h_chunk_n[pchunk[b]]; });
            CU(ctx, cudaMemcpyAsync(ctx->d_prob_order, order, (size_t)nprob * 4, cudaMemcpyHostToDevice, st));
            ctx->use_order = true;
        }
    }
    CU(ctx, cudaEventRecord(ctx->ev_stage, st));
    ctx->stage_pending = true;
    ctx->last_nmax = nmax;
    if (cacheable) { ctx->cache_B = B; ctx->cache_lc = uniform; ctx->cache_nprob = nprob; ctx->cache_tmax = tmax; }
    else ctx->cache_B = -1;
    return 0;
}

// which kernel, how many CTAs per problem, how many work slots and how much scratch each
struct LaunchPlan { bool use64 = false; int cl = 1; int slots = 0; size_t stride = 0; int npad = 0; };

int plan_launch(gprot_ctx* ctx, const ProblemList& pl, LaunchPlan& lp) {
    const int nprob = pl.nprob, tmax = pl.tmax, nmax = pl.nmax;
    if (!ctx->attr_set) {
        for (int cl : {1, 2, 4, 8})
            for (int ex = 0; ex < 2; ex++)
                CU(ctx, cudaFuncSetAttribute((const void*)kernel_for(ex != 0, cl), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        ctx->attr_set = true;
    }
    query_clusters(ctx);
    int cl = choose_cluster(ctx, nprob, tmax);
    // Launches with more than one problem per SM and N < 1920 run the two-problems-per-SM kernel (gp_kernel64.cuh); the others
    // keep one problem per SM (or per cluster).  gprot_set_cluster(64) / (1) force either.
    const bool use64 = ctx->force_cluster == 64 ||
                       (ctx->force_cluster == 0 && cl == 1 && nprob > ctx->num_sms && nmax < 1920);   // measured crossover
    if (use64) cl = 1;
    ctx->last_cluster = use64 ? 64 : cl;
    if (use64 && !ctx->attr64_set) {
        for (int ex = 0; ex < 2; ex++) {
            const void* f = ex ? (const void*)k64::gp_lnlike64_kernel<true> : (const void*)k64::gp_lnlike64_kernel<false>;
            CU(ctx, cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(k64::Smem2)));
            CU(ctx, cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        }
        ctx->attr64_set = true;
        if (getenv("GPROT_DEBUG")) {
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k64::gp_lnlike64_kernel<false>, k64::NT2, sizeof(k64::Smem2));
            fprintf(stderr, "[gprot] gp_lnlike64_kernel: %d CTAs per SM, %zu bytes of shared memory each\n", nb, sizeof(k64::Smem2));
        }
    }
    const int t64 = (nmax + k64::BC - 1) / k64::BC;
    lp.use64 = use64;
    lp.cl = cl;
    lp.slots = use64 ? std::min(nprob, 2 * ctx->num_sms) : std::min(nprob, ctx->max_clusters[cl]);   // one slot per resident CTA (group)
    lp.stride = use64 ? std::max<size_t>((size_t)t64 * (t64 - 1) / 2, 1) * k64::BLK2
                      : std::max<size_t>((size_t)tmax * (tmax - 1) / 2, 1) * BLKD;
    lp.npad = use64 ? (t64 + 1) * k64::BC : tmax * BS;        // the 64-wide kernel reads one block past the end
    return 0;
}

constexpr int TCAP = 1024;                                   // block columns the cluster kernels can track (N <= 131072)
constexpr size_t WG_STRIDE = WBUFS * (size_t)WX_DOUBLES + 2 * TCAP;

// factor scratch, per-point vectors and (cluster kernels) the W exchange buffers, grown on demand; may lower lp.slots to
// what fits in device memory
int ensure_scratch(gprot_ctx* ctx, const ProblemList& pl, LaunchPlan& lp) {
    // a previous call found that device memory holds only max_lslots slots of this stride: do not try again
    const bool capped = ctx->max_lslots > 0 && lp.stride <= ctx->lslot_stride && ctx->lslots >= ctx->max_lslots;
    if (!capped && (lp.stride > ctx->lslot_stride || lp.slots > ctx->lslots)) {
        CU(ctx, cudaDeviceSynchronize());
        if (ctx->d_lscratch) CU(ctx, cudaFree(ctx->d_lscratch));
        ctx->d_lscratch = nullptr;
        const size_t s2 = std::max(lp.stride, ctx->lslot_stride);
        int want = std::max(lp.slots, ctx->lslots);
        size_t freeb = 0, totalb = 0;
        CU(ctx, cudaMemGetInfo(&freeb, &totalb));
        // per slot: the factor scratch, the per-point vectors and (cluster kernels) the W exchange buffers allocated below
        const size_t per_slot = s2 * 8 + (size_t)5 * std::max(lp.npad, ctx->v_npad) * 8 + WG_STRIDE * 8 + (size_t)cctr_ints(TCAP) * 4;
        const size_t budget = (size_t)(0.85 * (double)freeb);
        ctx->max_lslots = 0;
        if ((size_t)want * per_slot > budget) {
            want = (int)(budget / per_slot);
            ctx->max_lslots = want;
        }
        if (want < 1) return fail(ctx, "not enough device memory for one factor scratch slot");
        CU(ctx, cudaMalloc((void**)&ctx->d_lscratch, (size_t)want * s2 * 8));
        ctx->lslot_stride = s2;
        ctx->lslots = want;
    }
    lp.slots = std::min(lp.slots, ctx->lslots);
    if (lp.npad > ctx->v_npad || lp.slots > ctx->vslots) {
        CU(ctx, cudaDeviceSynchronize());
        if (ctx->d_vscratch) CU(ctx, cudaFree(ctx->d_vscratch));
        ctx->d_vscratch = nullptr;
        const int np2 = std::max(lp.npad, ctx->v_npad);
        const int want = std::max(ctx->lslots, ctx->vslots);
        CU(ctx, cudaMalloc((void**)&ctx->d_vscratch, (size_t)want * 5 * np2 * 8));
        ctx->v_npad = np2;
        ctx->vslots = want;
    }
    if (lp.cl > 1) {
        if (pl.tmax > TCAP) return fail(ctx, "problem too large for the cluster kernels");
        if (lp.slots > ctx->cslots) {
            CU(ctx, cudaDeviceSynchronize());
            if (ctx->d_wg) CU(ctx, cudaFree(ctx->d_wg));
            if (ctx->d_cctr) CU(ctx, cudaFree(ctx->d_cctr));
            ctx->d_wg = nullptr; ctx->d_cctr = nullptr;
            const int want = std::max(lp.slots, ctx->num_sms / 2);
            CU(ctx, cudaMalloc((void**)&ctx->d_wg, (size_t)want * WG_STRIDE * 8));
            CU(ctx, cudaMalloc((void**)&ctx->d_cctr, (size_t)want * cctr_ints(TCAP) * 4));
            ctx->cslots = want;
        }
    }
    return 0;
}

// the likelihood kernel of the plan, bracketed by the timing events
int launch_likelihood(gprot_ctx* ctx, const ProblemList& pl, const LaunchPlan& lp, const double* d_theta, unsigned flags, cudaStream_t st) {
    Params prm;
    prm.wg = ctx->d_wg; prm.wg_stride = WG_STRIDE; prm.cctr = ctx->d_cctr; prm.tcap = TCAP;
    prm.t = ctx->d_t; prm.y = ctx->d_y; prm.yerr = ctx->d_yerr;
    prm.chunk_off = ctx->d_chunk_off; prm.chunk_n = ctx->d_chunk_n; prm.chunk_mingap2 = ctx->d_chunk_mingap2;
    prm.theta = d_theta;
    prm.prob_theta = ctx->d_prob_theta; prm.prob_chunk = ctx->d_prob_chunk; prm.prob_out = ctx->d_prob_out;
    prm.prob_order = ctx->use_order ? ctx->d_prob_order : nullptr;
    prm.nprob = pl.nprob;
    prm.counter = ctx->d_counter;
    prm.lscratch = ctx->d_lscratch; prm.lslot_stride = ctx->lslot_stride;
    prm.vscratch = ctx->d_vscratch; prm.npad_max = ctx->v_npad;
    prm.timing = ctx->d_timing;
    CU(ctx, cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), st));
    CU(ctx, cudaEventRecord(ctx->ev0, st));
    if (lp.use64) {
        if (flags & GPROT_EXACT_KERNEL) k64::gp_lnlike64_kernel<true><<<lp.slots, k64::NT2, sizeof(k64::Smem2), st>>>(prm);
        else                            k64::gp_lnlike64_kernel<false><<<lp.slots, k64::NT2, sizeof(k64::Smem2), st>>>(prm);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(lp.slots * lp.cl));
        cfg.blockDim = dim3(NT);
        cfg.dynamicSmemBytes = sizeof(Smem);
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)lp.cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = lp.cl > 1 ? 1 : 0;
        CU(ctx, cudaLaunchKernelEx(&cfg, kernel_for((flags & GPROT_EXACT_KERNEL) != 0, lp.cl), prm));
    }
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaEventRecord(ctx->ev1, st));
    ctx->launches++;
    return 0;
}

// core: likelihood of the rows of theta selected by mask (NULL = all)
int lnlike_core(gprot_ctx* ctx, int64_t B, const double* theta, const int32_t* lc_id, const unsigned char* mask,
                double* out, unsigned flags, cudaStream_t st) {
    if (!ctx) return fail(nullptr, "null context");
    if (B < 0 || (B > 0 && (!theta || !out))) return fail(ctx, "bad arguments");
    if (B == 0) return 0;
    if (ctx->lc_first_chunk.empty()) return fail(ctx, "no light curve registered");
    CU(ctx, cudaSetDevice(ctx->device));
    if (upload_arena(ctx)) return 1;

    ProblemList pl;
    if (build_problem_list(ctx, B, lc_id, mask, st, pl)) return 1;
    const int nprob = pl.nprob;

    // ---- theta / out on the device
    const double* d_theta = theta;
    if (!(flags & GPROT_THETA_ON_DEVICE)) {
        if (grow(ctx, &ctx->d_theta, &ctx->cap_theta, (size_t)B * 5)) return 1;
        CU(ctx, cudaMemcpyAsync(ctx->d_theta, theta, (size_t)B * 40, cudaMemcpyHostToDevice, st));
        d_theta = ctx->d_theta;
    }
    double* d_out = out;
    if (!(flags & GPROT_OUT_ON_DEVICE)) {
        if (grow(ctx, &ctx->d_out, &ctx->cap_out, (size_t)B)) return 1;
        d_out = ctx->d_out;
    }

    if (nprob > 0) {
        LaunchPlan lp;
        if (plan_launch(ctx, pl, lp) || ensure_scratch(ctx, pl, lp) || launch_likelihood(ctx, pl, lp, d_theta, flags, st)) return 1;
    }
    reduce_rows_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(ctx->d_prob_out, ctx->d_row_first, ctx->d_row_count, d_out, (long long)B);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    if (!(flags & GPROT_OUT_ON_DEVICE)) {
        CU(ctx, cudaMemcpyAsync(out, d_out, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
        CU(ctx, cudaStreamSynchronize(st));
        if (nprob > 0) CU(ctx, cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    }
    return 0;
}

}  // namespace

extern "C" {

int gprot_abi_version(void) { return 2; }

const char* gprot_last_error(const gprot_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int gprot_create(int device, gprot_ctx** out) {
    if (!out) return fail(nullptr, "null out pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, std::string("no CUDA device: gprot_b200 has no CPU fallback (") + cudaGetErrorString(e) + ")");
    if (device < 0 || device >= ndev) return fail(nullptr, "device ordinal out of range");
    gprot_ctx* ctx = new (std::nothrow) gprot_ctx;
    if (!ctx) return fail(nullptr, "out of host memory");
    ctx->device = device;
    cudaDeviceProp p;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&p, device) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, "cudaSetDevice / cudaGetDeviceProperties failed");
    }
    if (p.major != 10) {
        delete ctx;
        return fail(nullptr, "gprot_b200 is built for sm_100a (B200) only; found sm_" + std::to_string(p.major) + std::to_string(p.minor));
    }
    ctx->num_sms = p.multiProcessorCount;
    if (cudaMalloc((void**)&ctx->d_counter, 64) != cudaSuccess || cudaMalloc((void**)&ctx->d_timing, 16 * 8) != cudaSuccess || cudaEventCreate(&ctx->ev0) != cudaSuccess ||
        cudaEventCreate(&ctx->ev1) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_stage, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, "context allocation failed");
    }
    *out = ctx;
    return 0;
}

int gprot_destroy(gprot_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    void* ptrs[] = {ctx->d_t, ctx->d_y, ctx->d_yerr, ctx->d_chunk_mingap2, ctx->d_chunk_off, ctx->d_chunk_n, ctx->d_theta,
                    ctx->d_out, ctx->d_prob_theta, ctx->d_prob_chunk, ctx->d_prob_order, ctx->d_prob_out, ctx->d_row_first, ctx->d_row_count,
                    ctx->d_counter, ctx->d_lscratch, ctx->d_vscratch, ctx->d_timing, ctx->d_wg, ctx->d_cctr};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_stage) cudaEventDestroy(ctx->ev_stage);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    delete ctx;
    return 0;
}

int gprot_register_lc(gprot_ctx* ctx, int m, const int* n, const double* t, const double* y, const double* yerr, int* lc_id) {
    if (!ctx) return fail(nullptr, "null context");
    if (m <= 0 || !n || !t || !y || !yerr) return fail(ctx, "bad arguments");
    size_t pos = 0;
    for (int c = 0; c < m; c++) {
        if (n[c] <= 0) return fail(ctx, "empty chunk");
        pos += (size_t)n[c];
    }
    const int first = (int)ctx->h_chunk_n.size();
    pos = 0;
    for (int c = 0; c < m; c++) {
        ctx->h_chunk_off.push_back((long long)ctx->h_t.size());
        ctx->h_chunk_n.push_back(n[c]);
        std::vector<double> s(t + pos, t + pos + n[c]);
        std::sort(s.begin(), s.end());
        double g2 = std::numeric_limits<double>::infinity();
        for (int i = 1; i < n[c]; i++) {
            const double d = s[i] - s[i - 1];
            g2 = std::min(g2, d * d);
        }
        for (int i = 0; i < n[c]; i++)                       // non-finite stamps or errors: the chunk evaluates to -inf
            if (!std::isfinite(t[pos + i]) || !std::isfinite(yerr[pos + i])) g2 = -1.0;
        ctx->h_chunk_mingap2.push_back(g2);
        ctx->h_t.insert(ctx->h_t.end(), t + pos, t + pos + n[c]);
        ctx->h_y.insert(ctx->h_y.end(), y + pos, y + pos + n[c]);
        ctx->h_yerr.insert(ctx->h_yerr.end(), yerr + pos, yerr + pos + n[c]);
        // keep every chunk 16-byte aligned in the arena
        if (ctx->h_t.size() & 1) { ctx->h_t.push_back(0.0); ctx->h_y.push_back(0.0); ctx->h_yerr.push_back(0.0); }
        pos += (size_t)n[c];
    }
    ctx->lc_first_chunk.push_back(first);
    ctx->lc_nchunks.push_back(m);
    if (lc_id) *lc_id = (int)ctx->lc_first_chunk.size() - 1;
    return 0;                                                // the new tail of the arena goes to the device with the next call
}

int gprot_unregister_lc(gprot_ctx* ctx, int lc_id) {
    if (!ctx) return fail(nullptr, "null context");
    if (lc_id < 0 || lc_id >= (int)ctx->lc_first_chunk.size() || ctx->lc_nchunks[lc_id] == 0) return fail(ctx, "lc_id is not a registered light curve");
    const int f = ctx->lc_first_chunk[lc_id], m = ctx->lc_nchunks[lc_id];
    size_t live = 0;
    for (int c = f; c < f + m; c++) ctx->dead_points += (size_t)ctx->h_chunk_n[c] + ((size_t)ctx->h_chunk_n[c] & 1);
    ctx->lc_nchunks[lc_id] = 0;
    if (ctx->cache_lc == lc_id) ctx->cache_B = -1;
    live = ctx->h_t.size() - ctx->dead_points;
    if (ctx->dead_points > live && ctx->dead_points > (size_t)1 << 16) return compact_arena(ctx);
    return 0;
}

int gprot_unregister_all(gprot_ctx* ctx) {
    if (!ctx) return fail(nullptr, "null context");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaDeviceSynchronize());                        // a queued launch may still read the arena that is about to be rewritten
    ctx->h_t.clear(); ctx->h_y.clear(); ctx->h_yerr.clear();
    ctx->h_chunk_off.clear(); ctx->h_chunk_n.clear(); ctx->h_chunk_mingap2.clear();
    ctx->lc_first_chunk.clear(); ctx->lc_nchunks.clear();
    ctx->up_points = 0; ctx->up_chunks = 0; ctx->dead_points = 0;
    ctx->cache_B = -1;
    ctx->generation++;                                       // ids handed out before this call are void
    return 0;
}

int64_t gprot_generation(const gprot_ctx* ctx) { return ctx ? ctx->generation : -1; }

int gprot_lnlike_batch(gprot_ctx* ctx, int64_t B, const double* theta, const int32_t* lc_id, double* out, unsigned flags,
                       void* stream) {
    return lnlike_core(ctx, B, theta, lc_id, nullptr, out, flags, (cudaStream_t)stream);
}

int gprot_lnprior_batch(int64_t B, const double* theta, const double* bounds, const double* mu, const double* sigma,
                        double pmin, double pmax, const double* mixture, int k, double* out) {
    if (B < 0 || (B > 0 && (!theta || !bounds || !mu || !sigma || !out))) return fail(nullptr, "bad arguments");
    for (int64_t b = 0; b < B; b++) out[b] = lnprior_one(theta + 5 * b, bounds, mu, sigma, pmin, pmax, mixture, k);
    return 0;
}

int gprot_lnpost_batch(gprot_ctx* ctx, int64_t B, const double* theta, const int32_t* lc_id, const double* bounds,
                       const double* mu, const double* sigma, double pmin, double pmax, const double* mixture, int k,
                       double* out_lnprior, double* out_lnlike, unsigned flags, void* stream) {
    if (!ctx) return fail(nullptr, "null context");
    if (flags & (GPROT_THETA_ON_DEVICE | GPROT_OUT_ON_DEVICE)) return fail(ctx, "gprot_lnpost_batch takes host buffers");
    if (B > 0 && (!out_lnprior || !out_lnlike)) return fail(ctx, "bad arguments");
    if (gprot_lnprior_batch(B, theta, bounds, mu, sigma, pmin, pmax, mixture, k, out_lnprior)) return fail(ctx, g_err);
    std::vector<unsigned char> mask((size_t)B);
    for (int64_t b = 0; b < B; b++) mask[b] = std::isfinite(out_lnprior[b]) ? 1 : 0;
    return lnlike_core(ctx, B, theta, lc_id, mask.data(), out_lnlike, flags, (cudaStream_t)stream);
}

int gprot_sync(gprot_ctx* ctx) {
    if (!ctx) return fail(nullptr, "null context");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaDeviceSynchronize());
    return 0;
}

int gprot_last_kernel_ms(gprot_ctx* ctx, float* ms) {
    if (!ctx || !ms) return fail(ctx, "bad arguments");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaEventSynchronize(ctx->ev1));
    CU(ctx, cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    *ms = ctx->last_ms;
    return 0;
}

int64_t gprot_launch_count(const gprot_ctx* ctx) { return ctx ? ctx->launches : 0; }

int gprot_set_cluster(gprot_ctx* ctx, int ctas_per_problem) {
    if (!ctx) return fail(nullptr, "null context");
    if (ctas_per_problem != 0 && ctas_per_problem != 1 && ctas_per_problem != 2 && ctas_per_problem != 4 && ctas_per_problem != 8 &&
        ctas_per_problem != 64)
        return fail(ctx, "ctas_per_problem must be 0 (auto), 1, 2, 4, 8, or 64 (the two-problems-per-SM kernel)");
    ctx->force_cluster = ctas_per_problem;
    return 0;
}

int gprot_last_cluster(const gprot_ctx* ctx) { return ctx ? ctx->last_cluster : 0; }

int gprot_measure_dmma_peak(gprot_ctx* ctx, double* tflops, double* sm_clock_mhz) {
    if (!ctx || !tflops) return fail(ctx, "bad arguments");
    CU(ctx, cudaSetDevice(ctx->device));
    double* out = nullptr;
    long long* cyc = nullptr;
    const int grid = ctx->num_sms, block = 512, iters = 40000;
    CU(ctx, cudaMalloc((void**)&out, (size_t)grid * block * 8));
    CU(ctx, cudaMalloc((void**)&cyc, 8));
    cudaEvent_t e0, e1;
    CU(ctx, cudaEventCreate(&e0));
    CU(ctx, cudaEventCreate(&e1));
    dmma_peak_kernel<<<grid, block>>>(out, iters, cyc);
    CU(ctx, cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        CU(ctx, cudaEventRecord(e0));
        dmma_peak_kernel<<<grid, block>>>(out, iters, cyc);
        CU(ctx, cudaEventRecord(e1));
        CU(ctx, cudaEventSynchronize(e1));
        float ms;
        CU(ctx, cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    long long c = 0;
    CU(ctx, cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
    const double nmma = (double)grid * (block / 32) * (double)iters * 8.0;
    *tflops = nmma * 512.0 / (best * 1e-3) / 1e12;
    if (sm_clock_mhz) *sm_clock_mhz = (double)c / (best * 1e-3) / 1e6;
    cudaFree(out); cudaFree(cyc); cudaEventDestroy(e0); cudaEventDestroy(e1);
    ctx->launches += 4;
    return 0;
}

// ---------------------------------------------------------------- several GPUs from one process
// (star, walker) rows are independent: contiguous blocks of rows go to the devices of the group, one host thread per
// device drives its context, the host buffer is the gather.  No device-to-device traffic.
static int gfail(gprot_group* g, const std::string& msg) {
    if (g) g->err = msg;
    g_err = msg;
    return 1;
}

int gprot_group_create(int ndev, const int* devices, gprot_group** out) {
    if (!out) return fail(nullptr, "null out pointer");
    *out = nullptr;
    if (ndev <= 0 || !devices) return fail(nullptr, "bad arguments");
    gprot_group* g = new (std::nothrow) gprot_group;
    if (!g) return fail(nullptr, "out of host memory");
    for (int i = 0; i < ndev; i++) {
        gprot_ctx* c = nullptr;
        if (gprot_create(devices[i], &c)) {
            const std::string why = g_err;
            for (gprot_ctx* x : g->ctxs) gprot_destroy(x);
            delete g;
            return fail(nullptr, why);
        }
        g->ctxs.push_back(c);
    }
    for (size_t i = 0; i < g->ctxs.size(); i++) {
        group_worker* w = new group_worker;
        w->th = std::thread([w] { w->loop(); });
        g->workers.push_back(w);
    }
    *out = g;
    return 0;
}

int gprot_group_destroy(gprot_group* g) {
    if (!g) return 0;
    for (group_worker* w : g->workers) { w->stop(); delete w; }
    for (gprot_ctx* c : g->ctxs) gprot_destroy(c);
    delete g;
    return 0;
}

int gprot_group_size(const gprot_group* g) { return g ? (int)g->ctxs.size() : 0; }

const char* gprot_group_last_error(const gprot_group* g) { return g ? g->err.c_str() : g_err.c_str(); }

int gprot_group_register_lc(gprot_group* g, int m, const int* n, const double* t, const double* y, const double* yerr, int* lc_id) {
    if (!g) return fail(nullptr, "null group");
    int id0 = -1;
    for (size_t i = 0; i < g->ctxs.size(); i++) {          // light curves are replicated: 24 bytes per point
        int id = -1;
        if (gprot_register_lc(g->ctxs[i], m, n, t, y, yerr, &id)) return gfail(g, g->ctxs[i]->err);
        if (i == 0) id0 = id;
        else if (id != id0) return gfail(g, "light-curve ids diverged between the devices of the group");
    }
    if (lc_id) *lc_id = id0;
    return 0;
}

int gprot_group_unregister_all(gprot_group* g) {
    if (!g) return fail(nullptr, "null group");
    for (gprot_ctx* c : g->ctxs)
        if (gprot_unregister_all(c)) return gfail(g, c->err);
    return 0;
}

int gprot_group_unregister_lc(gprot_group* g, int lc_id) {
    if (!g) return fail(nullptr, "null group");
    for (gprot_ctx* c : g->ctxs)
        if (gprot_unregister_lc(c, lc_id)) return gfail(g, c->err);
    return 0;
}

// rows [lo, lo + cnt) of the batch go to device r; the persistent workers run the devices concurrently
static int group_run(gprot_group* g, int64_t B, const double* theta, const int32_t* lc_id, const unsigned char* mask, double* out, unsigned flags) {
    const int64_t G = (int64_t)g->ctxs.size();
    const int64_t base = B / G, rem = B % G;
    std::vector<char> used((size_t)G, 0);
    for (int64_t r = 0; r < G; r++) {
        const int64_t lo = r * base + std::min<int64_t>(r, rem), cnt = base + (r < rem ? 1 : 0);
        if (cnt == 0) continue;
        gprot_ctx* c = g->ctxs[(size_t)r];
        used[(size_t)r] = 1;
        g->workers[(size_t)r]->submit([=]() {
            return lnlike_core(c, cnt, theta + 5 * lo, lc_id ? lc_id + lo : nullptr, mask ? mask + lo : nullptr, out + lo, flags, nullptr);
        });
    }
    int bad = -1;
    for (int64_t r = 0; r < G; r++)
        if (used[(size_t)r] && g->workers[(size_t)r]->wait() != 0 && bad < 0) bad = (int)r;
    if (bad >= 0) return gfail(g, "device " + std::to_string(g->ctxs[(size_t)bad]->device) + ": " + g->ctxs[(size_t)bad]->err);
    return 0;
}

int gprot_group_lnlike_batch(gprot_group* g, int64_t B, const double* theta, const int32_t* lc_id, double* out, unsigned flags) {
    if (!g) return fail(nullptr, "null group");
    if (flags & (GPROT_THETA_ON_DEVICE | GPROT_OUT_ON_DEVICE)) return gfail(g, "gprot_group_lnlike_batch takes host buffers");
    if (B < 0 || (B > 0 && (!theta || !out))) return gfail(g, "bad arguments");
    if (B == 0) return 0;
    return group_run(g, B, theta, lc_id, nullptr, out, flags);
}

int gprot_group_lnpost_batch(gprot_group* g, int64_t B, const double* theta, const int32_t* lc_id, const double* bounds,
                             const double* mu, const double* sigma, double pmin, double pmax, const double* mixture, int k,
                             double* out_lnprior, double* out_lnlike, unsigned flags) {
    if (!g) return fail(nullptr, "null group");
    if (flags & (GPROT_THETA_ON_DEVICE | GPROT_OUT_ON_DEVICE)) return gfail(g, "gprot_group_lnpost_batch takes host buffers");
    if (B < 0 || (B > 0 && (!theta || !out_lnprior || !out_lnlike))) return gfail(g, "bad arguments");
    if (B == 0) return 0;
    if (gprot_lnprior_batch(B, theta, bounds, mu, sigma, pmin, pmax, mixture, k, out_lnprior)) return gfail(g, g_err);
    std::vector<unsigned char> mask((size_t)B);
    for (int64_t b = 0; b < B; b++) mask[(size_t)b] = std::isfinite(out_lnprior[b]) ? 1 : 0;
    return group_run(g, B, theta, lc_id, mask.data(), out_lnlike, flags);
}

int gprot_debug_read_timing(gprot_ctx* ctx, long long* host_out16) {
    if (!ctx || !host_out16) return fail(ctx, "bad arguments");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaDeviceSynchronize());
    CU(ctx, cudaMemcpy(host_out16, ctx->d_timing, 16 * 8, cudaMemcpyDeviceToHost));
    return 0;
}

int gprot_debug_read_scratch(gprot_ctx* ctx, int slot, size_t offset, size_t count, double* host_out) {
    if (!ctx || !host_out) return fail(ctx, "bad arguments");
    if (!ctx->d_lscratch || slot < 0 || slot >= ctx->lslots || offset + count > ctx->lslot_stride) return fail(ctx, "scratch range");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaDeviceSynchronize());
    CU(ctx, cudaMemcpy(host_out, ctx->d_lscratch + (size_t)slot * ctx->lslot_stride + offset, count * 8, cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
