// gprot_chain.cpp -- the ensemble sampler of the reference's fit driver on the host side of the C ABI.
//
// gprot/fit.py:82-98,123-135 builds an emcee3 Ensemble over Emcee3Model and runs Sampler(moves).run(ensemble, n) with the
// move mix KDEMove 0.4 / DEMove(1.0) 0.4 / DESnookerMove 0.2; every half-ensemble proposal set is one batched likelihood
// launch (gprot_lnpost_batch).  Driven from Python (gprotation_b200/sampler.py) a step of a 64-walker chain costs two
// launches (~0.24 ms each on 300-point chunks) plus ~0.26 ms of interpreter time; this file is the same loop -- the same
// moves, the same red-blue split, the same acceptance rule, statement by statement after sampler.py -- without the
// interpreter.  Host code only: the likelihood is reached through a callback, which is gprot_lnpost_batch for
// gprot_chain_run and anything the caller likes for gprot_chain_run_fn (tests use a Gaussian).
//
// The random stream is this file's own (xoshiro256++ seeded by splitmix64, Box-Muller normals), so a native chain and a
// Python-driven chain started from the same seed are two different, equally valid realisations of the same Markov chain.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/gprot_b200.h"

namespace {

struct Rng {
    uint64_t s[4];
    bool have_spare = false;
    double spare = 0.0;
    explicit Rng(uint64_t seed) {
        uint64_t x = seed;
        for (int i = 0; i < 4; i++) {                      // splitmix64
            x += 0x9E3779B97F4A7C15ull;
            uint64_t z = x;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            s[i] = z ^ (z >> 31);
        }
    }
    static uint64_t rotl(uint64_t v, int k) { return (v << k) | (v >> (64 - k)); }
    uint64_t next() {                                      // xoshiro256++
        const uint64_t r = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }         // [0, 1)
    int below(int n) { return (int)(uniform() * n); }
    double normal() {
        if (have_spare) { have_spare = false; return spare; }
        double u, v, r2;
        do { u = 2.0 * uniform() - 1.0; v = 2.0 * uniform() - 1.0; r2 = u * u + v * v; } while (r2 >= 1.0 || r2 == 0.0);
        const double f = std::sqrt(-2.0 * std::log(r2) / r2);
        spare = v * f; have_spare = true;
        return u * f;
    }
    // k distinct indices out of n, uniform over ordered k-tuples (sampler.py _distinct)
    void distinct(int n, int k, int* out) {
        for (int i = 0; i < k; i++) {
            int v;
            bool clash;
            do {
                v = below(n);
                clash = false;
                for (int j = 0; j < i; j++) clash |= out[j] == v;
            } while (clash);
            out[i] = v;
        }
    }
};

const double NEG_INF = -std::numeric_limits<double>::infinity();
const double LOG_2PI = 1.8378770664093453;

inline double logprob(double lp, double ll) {              // Proposal.log_probability: non-finite -> -inf
    const double p = lp + ll;
    return std::isfinite(p) ? p : NEG_INF;
}

// lower Cholesky factor of a small symmetric positive definite matrix (row-major d x d); false if not positive definite
bool cholesky(std::vector<double>& a, int d) {
    for (int j = 0; j < d; j++) {
        double s = a[j * d + j];
        for (int k = 0; k < j; k++) s -= a[j * d + k] * a[j * d + k];
        if (!(s > 0.0) || !std::isfinite(s)) return false;
        const double ljj = std::sqrt(s);
        a[j * d + j] = ljj;
        for (int i = j + 1; i < d; i++) {
            double t = a[i * d + j];
            for (int k = 0; k < j; k++) t -= a[i * d + k] * a[j * d + k];
            a[i * d + j] = t / ljj;
        }
        for (int k = j + 1; k < d; k++) a[j * d + k] = 0.0;
    }
    return true;
}

struct Chain {
    int nw, nd;
    const gprot_chain_moves* mv;
    Rng rng;
    std::vector<double> q, fac, s, c, lp_new, ll_new, tmp;
    std::vector<int> sidx, cidx;
    std::string err;
    Chain(int nw_, int nd_, const gprot_chain_moves* m, uint64_t seed) : nw(nw_), nd(nd_), mv(m), rng(seed) {}

    // log (1/n) sum_j N(x - c_j; C), C = L L^T: the KDE density of sampler.py KDEMove._logpdf.  As there, the points are
    // whitened once (ds = L^-1 c_j for the whole complement, xs = L^-1 x), so a pair costs five subtractions and an exp.
    std::vector<double> linv, ds, xs;
    void kde_prepare(int nc, const std::vector<double>& L) {
        linv.assign((size_t)nd * nd, 0.0);
        for (int col = 0; col < nd; col++)                   // L linv = I by forward substitution
            for (int a = col; a < nd; a++) {
                double t = (a == col) ? 1.0 : 0.0;
                for (int k = col; k < a; k++) t -= L[a * nd + k] * linv[k * nd + col];
                linv[a * nd + col] = t / L[a * nd + a];
            }
        ds.resize((size_t)nc * nd);
        for (int j = 0; j < nc; j++)
            for (int a = 0; a < nd; a++) {
                double t = 0.0;
                for (int k = 0; k <= a; k++) t += linv[a * nd + k] * c[(size_t)j * nd + k];
                ds[(size_t)j * nd + a] = t;
            }
        xs.resize(nd);
    }
    double kde_logpdf(const double* x, int nc, double logdet) {
        std::vector<double>& e = tmp;
        e.resize(nc);
        for (int a = 0; a < nd; a++) {
            double t = 0.0;
            for (int k = 0; k <= a; k++) t += linv[a * nd + k] * x[k];
            xs[a] = t;
        }
        double mx = NEG_INF;
        for (int j = 0; j < nc; j++) {
            double r2 = 0.0;
            for (int a = 0; a < nd; a++) { const double d = xs[a] - ds[(size_t)j * nd + a]; r2 += d * d; }
            e[j] = -0.5 * r2;
            mx = std::max(mx, e[j]);
        }
        double sum = 0.0;
        for (int j = 0; j < nc; j++) sum += std::exp(e[j] - mx);
        return mx + std::log(sum) - std::log((double)nc) - 0.5 * nd * LOG_2PI - logdet;
    }

    // proposals q [ns, nd] and Hastings factors for the walkers s given the complementary ensemble c (sampler.py moves)
    bool propose(int move, int ns, int nc) {
        q.assign((size_t)ns * nd, 0.0);
        fac.assign(ns, 0.0);
        if (move == 0) {                                    // StretchMove
            const double a = mv->stretch_a > 1.0 ? mv->stretch_a : 2.0;
            for (int i = 0; i < ns; i++) {
                const double zz = std::pow((a - 1.0) * rng.uniform() + 1.0, 2.0) / a;
                const int r = rng.below(nc);
                for (int k = 0; k < nd; k++) q[(size_t)i * nd + k] = c[(size_t)r * nd + k] - (c[(size_t)r * nd + k] - s[(size_t)i * nd + k]) * zz;
                fac[i] = (nd - 1.0) * std::log(zz);
            }
        } else if (move == 1) {                             // DEMove
            const double g0 = mv->de_gamma0 > 0.0 ? mv->de_gamma0 : 2.38 / std::sqrt(2.0 * nd);
            int ab[2];
            for (int i = 0; i < ns; i++) {
                const double f = mv->de_sigma * rng.normal();
                rng.distinct(nc, 2, ab);
                for (int k = 0; k < nd; k++)
                    q[(size_t)i * nd + k] = s[(size_t)i * nd + k] + (c[(size_t)ab[1] * nd + k] - c[(size_t)ab[0] * nd + k]) * (g0 * (1.0 + f));
            }
        } else if (move == 2) {                             // DESnookerMove
            const double gs = mv->snooker_gamma > 0.0 ? mv->snooker_gamma : 1.7;
            int w[3];
            for (int i = 0; i < ns; i++) {
                rng.distinct(nc, 3, w);
                const double *z = &c[(size_t)w[0] * nd], *z1 = &c[(size_t)w[1] * nd], *z2 = &c[(size_t)w[2] * nd], *si = &s[(size_t)i * nd];
                double norm = 0.0;
                for (int k = 0; k < nd; k++) norm += (si[k] - z[k]) * (si[k] - z[k]);
                norm = std::sqrt(norm);
                double proj = 0.0;
                for (int k = 0; k < nd; k++) proj += (si[k] - z[k]) / norm * (z1[k] - z2[k]);
                double qn = 0.0;
                for (int k = 0; k < nd; k++) {
                    const double v = si[k] + (si[k] - z[k]) / norm * (gs * proj);
                    q[(size_t)i * nd + k] = v;
                    qn += (v - z[k]) * (v - z[k]);
                }
                fac[i] = (nd - 1.0) * (std::log(std::sqrt(qn)) - std::log(norm));
            }
        } else {                                            // KDEMove
            const double factor = mv->kde_bw > 0.0 ? mv->kde_bw : std::pow((double)nc, -1.0 / (nd + 4));
            std::vector<double> mean(nd, 0.0), L((size_t)nd * nd, 0.0);
            for (int j = 0; j < nc; j++)
                for (int k = 0; k < nd; k++) mean[k] += c[(size_t)j * nd + k];
            for (int k = 0; k < nd; k++) mean[k] /= nc;
            for (int j = 0; j < nc; j++)
                for (int a = 0; a < nd; a++)
                    for (int b = 0; b <= a; b++) L[a * nd + b] += (c[(size_t)j * nd + a] - mean[a]) * (c[(size_t)j * nd + b] - mean[b]);
            for (int a = 0; a < nd; a++)
                for (int b = 0; b <= a; b++) { L[a * nd + b] *= factor * factor / (nc - 1.0); L[b * nd + a] = L[a * nd + b]; }
            if (!cholesky(L, nd)) { err = "KDE move: the covariance of the complementary ensemble is not positive definite"; return false; }
            double logdet = 0.0;
            for (int a = 0; a < nd; a++) logdet += std::log(L[a * nd + a]);
            kde_prepare(nc, L);
            std::vector<double> g(nd);
            for (int i = 0; i < ns; i++) {
                const int pick = rng.below(nc);
                for (int k = 0; k < nd; k++) g[k] = rng.normal();
                for (int a = 0; a < nd; a++) {
                    double v = c[(size_t)pick * nd + a];
                    for (int k = 0; k <= a; k++) v += L[a * nd + k] * g[k];
                    q[(size_t)i * nd + a] = v;
                }
                fac[i] = kde_logpdf(&s[(size_t)i * nd], nc, logdet) - kde_logpdf(&q[(size_t)i * nd], nc, logdet);
            }
        }
        return true;
    }
};

struct CtxTarget {                                          // gprot_chain_run: the callback is the likelihood ABI itself
    gprot_ctx* ctx;
    int lc_id;
    const double *bounds, *mu, *sigma, *mixture;
    double pmin, pmax;
    int k;
    std::vector<int32_t> ids;
    int status = 0;
};

void ctx_lnpost(void* user, int64_t n, const double* theta, double* lp, double* ll) {
    CtxTarget* t = static_cast<CtxTarget*>(user);
    t->ids.assign((size_t)n, t->lc_id);
    const int rc = gprot_lnpost_batch(t->ctx, n, theta, t->ids.data(), t->bounds, t->mu, t->sigma, t->pmin, t->pmax, t->mixture, t->k, lp, ll, 0, nullptr);
    if (rc != 0) {
        t->status = rc;
        for (int64_t i = 0; i < n; i++) { lp[i] = NEG_INF; ll[i] = NEG_INF; }
    }
}

thread_local std::string g_chain_err;

}  // namespace

extern "C" {

const char* gprot_chain_last_error(void) { return g_chain_err.c_str(); }

int gprot_chain_run_fn(gprot_lnpost_fn fn, void* user, int nwalkers, int ndim, int64_t nsteps, int thin, const gprot_chain_moves* moves,
                       uint64_t seed, double* coords, double* lnprior, double* lnlike, double* chain, double* chain_lnprior,
                       double* chain_lnlike, int64_t* naccepted) {
    g_chain_err.clear();
    if (!fn || !moves || !coords || !lnprior || !lnlike || nwalkers <= 0 || ndim <= 0 || nsteps < 0 || thin <= 0) {
        g_chain_err = "gprot_chain_run: bad arguments";
        return 1;
    }
    if (nwalkers < 2 * ndim) {                              // sampler.py RedBlueMove.update
        g_chain_err = "It is unadvisable to use a red-blue move with fewer walkers than twice the number of dimensions.";
        return 2;
    }
    const double w[4] = {std::max(0.0, moves->w_stretch), std::max(0.0, moves->w_de), std::max(0.0, moves->w_snooker), std::max(0.0, moves->w_kde)};
    const double wsum = w[0] + w[1] + w[2] + w[3];
    if (!(wsum > 0.0)) { g_chain_err = "gprot_chain_run: all move weights are zero"; return 1; }
    Chain ch(nwalkers, ndim, moves, seed);
    const int nd = ndim;
    std::vector<char> accepted(nwalkers);
    int64_t stored = 0;
    for (int64_t it = 0; it < nsteps; it++) {
        int move = 3;                                       // Sampler.sample: cumulative weights, searchsorted(side="right")
        {
            const double u = ch.rng.uniform() * wsum;
            double cum = 0.0;
            for (int m = 0; m < 4; m++) { cum += w[m]; if (u < cum) { move = m; break; } }
            while (w[move] == 0.0 && move > 0) move--;
        }
        std::fill(accepted.begin(), accepted.end(), 0);
        for (int half = 0; half < 2; half++) {              // inds = arange(nwalkers) % 2; S1 = inds == half
            ch.sidx.clear(); ch.cidx.clear();
            for (int i = 0; i < nwalkers; i++) ((i % 2 == half) ? ch.sidx : ch.cidx).push_back(i);
            const int ns = (int)ch.sidx.size(), nc = (int)ch.cidx.size();
            ch.s.resize((size_t)ns * nd); ch.c.resize((size_t)nc * nd);
            for (int i = 0; i < ns; i++) std::memcpy(&ch.s[(size_t)i * nd], coords + (size_t)ch.sidx[i] * nd, nd * sizeof(double));
            for (int j = 0; j < nc; j++) std::memcpy(&ch.c[(size_t)j * nd], coords + (size_t)ch.cidx[j] * nd, nd * sizeof(double));
            if (!ch.propose(move, ns, nc)) { g_chain_err = ch.err; return 3; }
            ch.lp_new.assign(ns, NEG_INF); ch.ll_new.assign(ns, NEG_INF);
            fn(user, ns, ch.q.data(), ch.lp_new.data(), ch.ll_new.data());       // ONE batched evaluation
            for (int i = 0; i < ns; i++) {
                const int wk = ch.sidx[i];
                const double lpn = std::isfinite(ch.lp_new[i]) ? ch.lp_new[i] : NEG_INF;             // Proposal.__init__
                const double lln = (std::isfinite(ch.lp_new[i]) && std::isfinite(ch.ll_new[i])) ? ch.ll_new[i] : NEG_INF;
                const double newp = logprob(lpn, lln);
                const double diff = ch.fac[i] + newp - (lnprior[wk] + lnlike[wk]);
                const double lu = std::log(ch.rng.uniform());
                if (std::isfinite(newp) && diff > lu) {
                    std::memcpy(coords + (size_t)wk * nd, &ch.q[(size_t)i * nd], nd * sizeof(double));
                    lnprior[wk] = lpn; lnlike[wk] = lln;
                    accepted[wk] = 1;
                }
            }
        }
        if ((it + 1) % thin == 0) {                         // backend.update(ensemble)
            if (chain) std::memcpy(chain + (size_t)stored * nwalkers * nd, coords, (size_t)nwalkers * nd * sizeof(double));
            if (chain_lnprior) std::memcpy(chain_lnprior + (size_t)stored * nwalkers, lnprior, nwalkers * sizeof(double));
            if (chain_lnlike) std::memcpy(chain_lnlike + (size_t)stored * nwalkers, lnlike, nwalkers * sizeof(double));
            if (naccepted) for (int i = 0; i < nwalkers; i++) naccepted[i] += accepted[i];
            stored++;
        }
    }
    return 0;
}

int gprot_chain_run(gprot_ctx* ctx, int lc_id, const double* bounds, const double* mu, const double* sigma, double pmin, double pmax,
                    const double* mixture, int k, int nwalkers, int64_t nsteps, int thin, const gprot_chain_moves* moves, uint64_t seed,
                    double* coords, double* lnprior, double* lnlike, double* chain, double* chain_lnprior, double* chain_lnlike,
                    int64_t* naccepted) {
    if (!ctx) { g_chain_err = "gprot_chain_run: no context"; return 1; }
    CtxTarget t{ctx, lc_id, bounds, mu, sigma, mixture, pmin, pmax, k, {}, 0};
    const int rc = gprot_chain_run_fn(ctx_lnpost, &t, nwalkers, 5, nsteps, thin, moves, seed, coords, lnprior, lnlike, chain, chain_lnprior,
                                      chain_lnlike, naccepted);
    if (t.status != 0) {
        g_chain_err = std::string("gprot_chain_run: the likelihood call failed: ") + gprot_last_error(ctx);
        return t.status;
    }
    return rc;
}

}  // extern "C"
