"""CPU oracle for the GProtation hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A dense fp64 NumPy/SciPy restatement of the arithmetic that the reference performs for
every emcee3 walker proposal (reference: /root/reference/gprot/model.py).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module; the product path (``gprotation_b200``) never does.

PARITY UNPINNED.  The arithmetic lives in the third-party package **george** (0.2.x API:
``WhiteKernel``, ``solver=george.HODLRSolver``, ``gp.lnlikelihood(y, quiet=True)``;
call sites gprot/model.py:9-10,334,343,346,354), which is neither vendored under
/root/reference nor installable offline, and the reference ships no tests, golden vectors
or recorded lnlike values.  This file therefore restates george 0.2.x's published
semantics (kernels.h: ConstantKernel, ExpSquaredKernel(metric), ExpSine2Kernel(gamma,
period), WhiteKernel; GP.compute adds yerr**2 to the diagonal; BasicSolver = dense
Cholesky) and is anchored on the reference's own call sites.  It is cross-checked by
analytic known answers, an mpmath evaluation and an 80-bit long-double C evaluation
(oracle/gp_oracle_ld.c), see tests/test_oracle.py.  The reference's HODLR solver
(nleaf=100, tol=1e-12) approximates the same matrix to O(tol*cond).

Exact matrix (gprot/model.py:328-347):
    K_ij = ((A * exp(-0.5 * (d*d / L))) * exp(-G * s * s)) + delta_ij * (S + (S + yerr_i**2))
    d = t_i - t_j ;  s = sin(M_PI * d / P) ;  (A, L, G, S, P) = exp(theta)
Likelihood (george GP.lnlikelihood, zero mean):
    lnlike = -0.5 * y^T K^-1 y - 0.5 * (log det K + N log 2 pi)
"""
from __future__ import annotations

import math

import numpy as np
from scipy.linalg import cho_factor, cho_solve, solve_triangular
from scipy.special import logsumexp

# gprot/model.py:36-45 (class constants of GPRotModel)
DEFAULT_BOUNDS = ((-20.0, 0.0), (2.0, 20.0), (-10.0, 3.0), (-20.0, 0.0), (-0.69, 4.61))
DEFAULT_GP_PRIOR_MU = (-13.0, 7.2, -2.3, -17.0)
DEFAULT_GP_PRIOR_SIGMA = (5.7, 1.2, 1.4, 5.0)
# gprot/kepler.py:27-34 (KeplerGPRotModel overrides)
KEPLER_BOUNDS = ((-20.0, 0.0), (2.0, 8.0), (0.0, 3.0), (-20.0, 0.0), (-0.69, 4.61))
KEPLER_GP_PRIOR_MU = (-13.0, 5.0, 1.9, -17.0)
KEPLER_GP_PRIOR_SIGMA = (5.7, 1.2, 1.4, 5.0)


def hyper(theta):
    """gprot/model.py:328-333 -- five np.exp calls."""
    return tuple(float(np.exp(v)) for v in theta)


def kernel_matrix(theta, t, yerr):
    """The N x N matrix george factorises for gp(theta, x, yerr) (gprot/model.py:336-347).

    george 0.2.x order of operations:
      ExpSquaredKernel(L):    r2 = d*d / L ; exp(-0.5 * r2)
      ExpSine2Kernel(G, P):   s = sin(M_PI * d / P) ; exp(-G * s * s)
      A * k1 * k2          -> Product(Product(Constant(A), k1), k2) = (A * k1) * k2
      + WhiteKernel(S)     -> S where the two inputs coincide (r2 < DBL_EPSILON)
      GP.compute(x, e)     -> + e_i**2 on the diagonal, e = sqrt(S + yerr**2)  (model.py:345-346)
    """
    A, L, G, S, P = hyper(theta)
    t = np.asarray(t, dtype=np.float64)
    yerr = np.asarray(yerr, dtype=np.float64)
    d = t[:, None] - t[None, :]
    r2 = d * d / L
    s = np.sin(np.pi * d / P)
    K = (A * np.exp(-0.5 * r2)) * np.exp(-G * s * s)
    K = K + np.where(r2 < np.finfo(np.float64).eps, S, 0.0)
    e = np.sqrt(S + yerr ** 2)
    K[np.diag_indices_from(K)] += e ** 2
    return K


def lnlike_function(theta, t, y, yerr):
    """gprot/model.py:349-356: failure or non-finite -> -inf, never an exception."""
    try:
        with np.errstate(all="ignore"):
            K = kernel_matrix(theta, t, yerr)
            if not np.all(np.isfinite(K)):
                return -np.inf
            c, low = cho_factor(K, lower=True, overwrite_a=True, check_finite=False)
            y = np.asarray(y, dtype=np.float64)
            z = solve_triangular(c, y, lower=True, check_finite=False)
            logdet = 2.0 * np.sum(np.log(np.diag(c)))
            lnl = -0.5 * float(z @ z) - 0.5 * (logdet + len(y) * math.log(2.0 * math.pi))
    except (ValueError, np.linalg.LinAlgError):
        return -np.inf
    return lnl if np.isfinite(lnl) else -np.inf


def lnlike(theta, t, y, yerr, chunks=None):
    """gprot/model.py:358-365: one problem, or the sum over the chunk list.

    ``chunks`` is a list of index arrays / slices (lc.x_list semantics, gprot/lc.py:329-342)."""
    if chunks is None:
        return lnlike_function(theta, t, y, yerr)
    return float(np.sum([lnlike_function(theta, t[c], y[c], yerr[c]) for c in chunks]))


def split_chunks(n, chunksize):
    """gprot/lc.py:329-342: np.array_split into ceil(n/chunksize) near-equal chunks."""
    if chunksize is None:
        return None
    m = n // chunksize
    if n % chunksize != 0:
        m += 1
    return np.array_split(np.arange(n), m)


def lnGauss(x, mu, sigma):
    """gprot/model.py:22-23."""
    return -0.5 * ((x - mu) ** 2 / (sigma ** 2)) + np.log(1.0 / np.sqrt(2 * np.pi * sigma ** 2))


def lnGauss_mixture(x, mix):
    """gprot/model.py:25-29 (scipy.misc.logsumexp moved to scipy.special)."""
    mix = np.asarray(mix, dtype=np.float64)
    return float(logsumexp(lnGauss(x, mix[:, 1], mix[:, 2]), b=mix[:, 0]))


def lnprior(theta, bounds=DEFAULT_BOUNDS, mu=DEFAULT_GP_PRIOR_MU, sigma=DEFAULT_GP_PRIOR_SIGMA,
            pmin=-np.inf, pmax=np.inf, mixture=None):
    """gprot/model.py:132-163."""
    for i, (lo, hi) in enumerate(bounds):
        if theta[i] < lo or theta[i] > hi:
            return -np.inf
    if theta[-1] < pmin or theta[-1] > pmax:
        return -np.inf
    if theta[1] < theta[-1]:
        return -np.inf
    lnpr = float(np.sum(lnGauss(np.array(theta[:-1], dtype=np.float64), np.asarray(mu, float), np.asarray(sigma, float))))
    if mixture is not None:
        lnpr += lnGauss_mixture(theta[-1], mixture)
    return lnpr


def lnpost(theta, t, y, yerr, chunks=None, **prior_kw):
    """gprot/model.py:367-369 -- evaluates the likelihood even where the prior is -inf."""
    return lnlike(theta, t, y, yerr, chunks) + lnprior(theta, **prior_kw)


def sample_from_prior(n, seed=None, bounds=DEFAULT_BOUNDS, mu=DEFAULT_GP_PRIOR_MU, sigma=DEFAULT_GP_PRIOR_SIGMA,
                      pmin=-np.inf, pmax=np.inf):
    """gprot/model.py:111-130 + 227-235 (no ACF prior): Gaussian draws for theta[:4], uniform ln P,
    redrawn until lnprior is finite.  Same legacy-RNG call order as the reference."""
    ndim = 5
    samples = np.inf * np.ones((n, ndim))
    rs = np.random.RandomState(seed)
    m = np.ones(n, dtype=bool)
    nbad = int(m.sum())
    mu = np.asarray(mu, float)
    sigma = np.asarray(sigma, float)
    loP, hiP = bounds[-1]
    loP = max(loP, pmin)
    hiP = min(hiP, pmax)
    while nbad > 0:
        rn = rs.randn(n * (ndim - 1)).reshape((n, ndim - 1))
        for i in range(ndim - 1):
            samples[m, i] = rn[m, i] * sigma[i] + mu[i]
        samples[m, -1] = rs.random_sample(nbad) * (hiP - loP) + loP
        lnp = np.array([lnprior(th, bounds, mu, sigma, pmin, pmax) for th in samples])
        m = ~np.isfinite(lnp)
        nbad = int(m.sum())
    return samples


def flops_per_eval(n):
    """Cholesky (LAPACK count) + one forward solve."""
    return n ** 3 / 3.0 + n ** 2 / 2.0 + n / 6.0 + n ** 2


def lnlike_mpmath(theta, t, y, yerr, dps=40):
    """Extended-precision evaluation of the same formula (small N only; pure Python loops)."""
    import mpmath as mp

    mp.mp.dps = dps
    A, L, G, S, P = [mp.e ** mp.mpf(float(v)) for v in theta]
    n = len(t)
    tt = [mp.mpf(float(v)) for v in t]
    K = mp.zeros(n, n)
    for i in range(n):
        for j in range(n):
            d = tt[i] - tt[j]
            s = mp.sin(mp.pi * d / P)
            K[i, j] = A * mp.e ** (-d * d / (2 * L)) * mp.e ** (-G * s * s)
        K[i, i] += 2 * S + mp.mpf(float(yerr[i])) ** 2
    Lc = mp.cholesky(K)
    yv = mp.matrix([mp.mpf(float(v)) for v in y])
    z = mp.lu_solve(Lc, yv)  # Lc is triangular: exact forward substitution up to dps
    quad = sum(z[i] ** 2 for i in range(n))
    logdet = 2 * sum(mp.log(Lc[i, i]) for i in range(n))
    return float(-(quad + logdet + n * mp.log(2 * mp.pi)) / 2)
