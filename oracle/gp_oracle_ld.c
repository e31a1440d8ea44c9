/* CPU oracle, extended precision -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the quasi-periodic GP log-likelihood of the reference
 * (gprot/model.py:328-369 + george 0.2.x kernel/GP semantics, see oracle/gp_oracle.py
 * for the full derivation and the PARITY UNPINNED note).  Two entry points:
 *
 *   gp_oracle_lnlike_f64 : same operation order as george in IEEE double, dense unblocked
 *                          Cholesky (the arithmetic of george's BasicSolver).
 *   gp_oracle_lnlike_ld  : every operation in x87 80-bit long double (64-bit mantissa);
 *                          an independent "truth" that tells rounding differences from
 *                          formula differences when the fp64 paths disagree at high cond(K).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 * Build: make -C oracle   (gcc -O2 -shared -fPIC -o oracle/_build/libgp_oracle.so)
 */
#include <math.h>
#include <stdlib.h>
#include <float.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* theta = (ln A, ln l, ln G, ln sigma, ln P), gprot/model.py:328-333 */

double gp_oracle_lnlike_f64(const double *theta, int n, const double *t, const double *y, const double *yerr)
{
    double A = exp(theta[0]), L = exp(theta[1]), G = exp(theta[2]), S = exp(theta[3]), P = exp(theta[4]);
    double *K = (double *)malloc((size_t)n * n * sizeof(double));
    double *z = (double *)malloc((size_t)n * sizeof(double));
    if (!K || !z) { free(K); free(z); return -INFINITY; }
    for (int i = 0; i < n; i++) {
        for (int j = 0; j <= i; j++) {
            double d = t[i] - t[j];
            double r2 = d * d / L;                    /* ExpSquaredKernel(metric=L) */
            double s = sin(M_PI * d / P);             /* ExpSine2Kernel(gamma=G, period=P) */
            double k = (A * exp(-0.5 * r2)) * exp(-G * s * s);
            if (r2 < DBL_EPSILON) k += S;             /* WhiteKernel(S) */
            if (i == j) { double e = sqrt(S + yerr[i] * yerr[i]); k += e * e; } /* GP.compute(x, e) */
            K[(size_t)i * n + j] = k;
        }
    }
    double logdet = 0.0, quad = 0.0;
    int ok = 1;
    for (int j = 0; j < n && ok; j++) {               /* row-oriented Cholesky-Banachiewicz */
        for (int c = 0; c <= j; c++) {
            double s = K[(size_t)j * n + c];
            for (int k = 0; k < c; k++) s -= K[(size_t)j * n + k] * K[(size_t)c * n + k];
            if (c == j) {
                if (!(s > 0.0) || !isfinite(s)) { ok = 0; break; }
                K[(size_t)j * n + j] = sqrt(s);
            } else {
                K[(size_t)j * n + c] = s / K[(size_t)c * n + c];
            }
        }
        if (!ok) break;
        double s = y[j];
        for (int k = 0; k < j; k++) s -= K[(size_t)j * n + k] * z[k];
        z[j] = s / K[(size_t)j * n + j];
        quad += z[j] * z[j];
        logdet += 2.0 * log(K[(size_t)j * n + j]);
    }
    free(K); free(z);
    if (!ok) return -INFINITY;
    double lnl = -0.5 * quad - 0.5 * (logdet + n * log(2.0 * M_PI));
    return isfinite(lnl) ? lnl : -INFINITY;
}

double gp_oracle_lnlike_ld(const double *theta, int n, const double *t, const double *y, const double *yerr)
{
    const long double PI = 3.14159265358979323846264338327950288L;
    /* hyper-parameters are the *double* values the reference forms with np.exp, then promoted */
    long double A = exp(theta[0]), L = exp(theta[1]), G = exp(theta[2]), S = exp(theta[3]), P = exp(theta[4]);
    long double *K = (long double *)malloc((size_t)n * n * sizeof(long double));
    long double *z = (long double *)malloc((size_t)n * sizeof(long double));
    if (!K || !z) { free(K); free(z); return -INFINITY; }
    for (int i = 0; i < n; i++) {
        for (int j = 0; j <= i; j++) {
            long double d = (long double)t[i] - (long double)t[j];
            long double r2 = d * d / L;
            long double s = sinl(PI * d / P);
            long double k = A * expl(-0.5L * r2) * expl(-G * s * s);
            if (r2 < DBL_EPSILON) k += S;
            if (i == j) k += S + (long double)yerr[i] * (long double)yerr[i];
            K[(size_t)i * n + j] = k;
        }
    }
    long double logdet = 0.0L, quad = 0.0L;
    int ok = 1;
    for (int j = 0; j < n && ok; j++) {
        for (int c = 0; c <= j; c++) {
            long double s = K[(size_t)j * n + c];
            for (int k = 0; k < c; k++) s -= K[(size_t)j * n + k] * K[(size_t)c * n + k];
            if (c == j) {
                if (!(s > 0.0L)) { ok = 0; break; }
                K[(size_t)j * n + j] = sqrtl(s);
            } else {
                K[(size_t)j * n + c] = s / K[(size_t)c * n + c];
            }
        }
        if (!ok) break;
        long double s = y[j];
        for (int k = 0; k < j; k++) s -= K[(size_t)j * n + k] * z[k];
        z[j] = s / K[(size_t)j * n + j];
        quad += z[j] * z[j];
        logdet += 2.0L * logl(K[(size_t)j * n + j]);
    }
    free(K); free(z);
    if (!ok) return -INFINITY;
    long double lnl = -0.5L * quad - 0.5L * (logdet + n * logl(2.0L * PI));
    return (double)lnl;
}
