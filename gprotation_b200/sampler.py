"""Ensemble sampler with the emcee3 surface that gprot/fit.py uses -- proposals evaluated as ONE batched launch.

The reference drives the likelihood through dfm/emcee3 (unreleased, unpinned; not installable offline --
SURVEY.md section 8c).  Its call sites (gprot/fit.py:6-7, 11-22, 83-89, 98, 103, 119, 128, 139) use:

    emcee3.Model / State          walker -> log_prior, log_likelihood
    emcee3.Ensemble(walker, coords, pool=pool)
    emcee3.moves.KDEMove(), DEMove(1.0), DESnookerMove()   mixed with weights 0.4 / 0.4 / 0.2
    emcee3.Sampler(moves, backend=...).run(ensemble, n, progress=...)
    sampler.get_integrated_autocorr_time(c=1), sampler.get_coords(flat=True, discard=...), sampler.reset()
    emcee3.backends.Backend / HDFBackend (current_coords, niter), emcee3.autocorr.AutocorrError
    emcee3.pools.DefaultPool

This module restates that surface (same names, argument meaning and error behaviour) following emcee3's published
algorithms: red-blue ensemble moves (each half of the walkers is updated from the complementary half, Foreman-Mackey
et al. 2013), differential evolution (ter Braak 2006), the snooker update (ter Braak & Vrugt 2008), a KDE
independence proposal, and the windowed integrated autocorrelation time.  The one structural difference is the point
of the whole repository: every proposal set of a half-ensemble is known before any likelihood is needed
(SURVEY.md section 8a, a11), so ``Ensemble.propose`` hands the *list* of states to ``pool.map`` once and
``fit.EnsemblePool`` turns it into a single device launch instead of one process call per walker.
"""
from __future__ import annotations

import os

import numpy as np

from .fit import Emcee3Model, EnsemblePool, State

__all__ = ["AutocorrError", "integrated_time", "Backend", "NPZBackend", "HDFBackend", "Ensemble", "Sampler",
           "Proposal", "StretchMove", "DEMove", "DESnookerMove", "KDEMove", "DefaultPool", "Model", "State"]

Model = Emcee3Model


class DefaultPool(object):
    """emcee3.pools.DefaultPool: serial ``map``.  With an ``Emcee3Model`` walker the states are still evaluated as one
    batch (the model's ``evaluate_states``) -- per-walker Python calls are what this repository removes."""

    def map(self, func, iterable):
        if isinstance(func, Emcee3Model) and hasattr(func.mod, "lnpost_batch"):
            return func.evaluate_states(iterable)
        return [func(x) for x in iterable]


# ------------------------------------------------------------------------------------------ autocorrelation
class AutocorrError(Exception):
    """Raised when the chain is too short to estimate the autocorrelation time (emcee3.autocorr.AutocorrError)."""


def autocorr_function(x, axis=0):
    """Normalised autocorrelation function of a (multi-dimensional) time series along ``axis`` (FFT, zero padded)."""
    x = np.atleast_1d(x)
    m = [slice(None)] * x.ndim
    n = int(2 ** np.ceil(np.log2(x.shape[axis]))) * 2
    f = np.fft.fft(x - np.mean(x, axis=axis, keepdims=True), n=n, axis=axis)
    m[axis] = slice(0, x.shape[axis])
    acf = np.fft.ifft(f * np.conjugate(f), axis=axis)[tuple(m)].real
    m[axis] = 0
    with np.errstate(invalid="ignore", divide="ignore"):
        return acf / np.expand_dims(acf[tuple(m)], axis)


def integrated_time(x, low=10, high=None, step=1, c=10, full_output=False, axis=0):
    """Windowed estimate tau = 1 + 2 sum_{t<M} rho(t) with the smallest window M > c * tau
    (the estimator gprot/fit.py:103 calls through ``get_integrated_autocorr_time(c=1)``)."""
    size = 0.5 * x.shape[axis]
    if int(size) <= low:
        raise AutocorrError("The chain is too short")
    if high is None:
        high = int(size / c)
    f = autocorr_function(x, axis=axis)
    if f.ndim == 1:
        f = f[:, None]
        axis = 0
    f = np.moveaxis(f, axis, 0)
    csum = np.cumsum(f[1:], axis=0)                 # csum[M-2] = sum_{t=1}^{M-1} rho(t)
    for M in np.arange(low, max(high, low + 1), step).astype(int):
        tau = 1.0 + 2.0 * csum[M - 2]
        if np.all(tau > 1.0) and M > c * tau.max():
            if full_output:
                return tau, M
            return tau
        if c * tau.max() >= size:
            break
    raise AutocorrError("The chain is too short to reliably estimate the autocorrelation time")


# ------------------------------------------------------------------------------------------ backends
class Backend(object):
    """In-memory chain store (emcee3.backends.Backend)."""

    def __init__(self):
        self._clear()

    def reset(self):
        self._clear()

    def _clear(self):
        self.niter = 0
        self._coords = None
        self._log_prior = None
        self._log_likelihood = None
        self._accepted = None
        self._size = 0

    def extend(self, n, nwalkers, ndim):
        if self._coords is None:
            self._coords = np.empty((n, nwalkers, ndim))
            self._log_prior = np.empty((n, nwalkers))
            self._log_likelihood = np.empty((n, nwalkers))
            self._accepted = np.zeros(nwalkers, dtype=np.int64)
            self._size = n
            return
        if self._coords.shape[1:] != (nwalkers, ndim):
            raise ValueError("the ensemble shape does not match the stored chain")
        need = self.niter + n
        if need > self._size:
            grow = max(need, 2 * self._size) - self._size
            self._coords = np.concatenate([self._coords, np.empty((grow, nwalkers, ndim))])
            self._log_prior = np.concatenate([self._log_prior, np.empty((grow, nwalkers))])
            self._log_likelihood = np.concatenate([self._log_likelihood, np.empty((grow, nwalkers))])
            self._size += grow

    def update(self, ensemble):
        i = self.niter
        if self._coords is None or i >= self._size:
            self.extend(max(1, self._size), ensemble.nwalkers, ensemble.ndim)
        self._coords[i] = ensemble.coords
        self._log_prior[i] = ensemble.log_prior
        self._log_likelihood[i] = ensemble.log_likelihood
        self._accepted += ensemble.acceptance
        self.niter += 1

    # ---- views
    def _check(self):
        if self.niter <= 0 or self._coords is None:
            raise AttributeError("You need to run the chain first or store the chain using the store keyword argument")

    @property
    def coords(self):
        self._check()
        return self._coords[: self.niter]

    @property
    def log_prior(self):
        self._check()
        return self._log_prior[: self.niter]

    @property
    def log_likelihood(self):
        self._check()
        return self._log_likelihood[: self.niter]

    @property
    def log_probability(self):
        return self.log_prior + self.log_likelihood

    @property
    def acceptance_fraction(self):
        self._check()
        return self._accepted / float(self.niter)

    @property
    def current_coords(self):
        self._check()
        return self._coords[self.niter - 1]

    def get_coords(self, flat=False, thin=1, discard=0):
        v = self.coords[discard + thin - 1 :: thin]
        return v.reshape(-1, v.shape[-1]) if flat else v

    def get_log_probability(self, flat=False, thin=1, discard=0):
        v = self.log_probability[discard + thin - 1 :: thin]
        return v.reshape(-1) if flat else v

    def get_integrated_autocorr_time(self, **kwargs):
        return integrated_time(np.mean(self.get_coords(), axis=1), **kwargs)


class NPZBackend(Backend):
    """Chain store persisted to ``filename`` (.npz) so a fit can resume from ``current_coords``
    (gprot/fit.py:69-80 does this with emcee3's HDFBackend; h5py is not installed here, the file layout is the
    same set of arrays).  ``save()`` is called by the sampler after every ``run``."""

    def __init__(self, filename):
        self.filename = filename
        super(NPZBackend, self).__init__()
        if os.path.exists(filename):
            with np.load(filename) as z:
                self._coords = z["coords"].copy()
                self._log_prior = z["log_prior"].copy()
                self._log_likelihood = z["log_likelihood"].copy()
                self._accepted = z["accepted"].copy()
                self.niter = int(self._coords.shape[0])
                self._size = self.niter

    def reset(self):
        self._clear()
        if os.path.exists(self.filename):
            os.remove(self.filename)

    def save(self):
        if self.niter <= 0:
            return
        d = os.path.dirname(os.path.abspath(self.filename))
        if not os.path.exists(d):
            os.makedirs(d)
        tmp = self.filename + ".tmp.npz"
        np.savez(tmp, coords=self.coords, log_prior=self.log_prior, log_likelihood=self.log_likelihood,
                 accepted=self._accepted)
        os.replace(tmp, self.filename)


def HDFBackend(filename, **kwargs):
    """emcee3.backends.HDFBackend stand-in: an HDF5-backed store needs h5py, which this image does not have; the
    chain is kept in the sibling ``.npz`` instead (same arrays, same resume behaviour)."""
    base, ext = os.path.splitext(filename)
    return NPZBackend(base + ".npz" if ext in (".h5", ".hdf5", ".hdf") else filename)


# ------------------------------------------------------------------------------------------ ensemble
class Proposal(object):
    """The evaluated states of one proposal set, as arrays (what ``Ensemble.propose`` returns).  Iterating yields
    ``State`` objects for code written against the per-walker emcee3 interface."""

    def __init__(self, coords, log_prior, log_likelihood):
        self.coords = np.asarray(coords, dtype=np.float64)
        lp = np.asarray(log_prior, dtype=np.float64)
        ll = np.asarray(log_likelihood, dtype=np.float64)
        ok = np.isfinite(lp)
        self.log_prior = np.where(ok, lp, -np.inf)
        self.log_likelihood = np.where(ok & np.isfinite(ll), ll, -np.inf)
        self.accepted = np.zeros(len(self.coords), dtype=bool)

    @property
    def log_probability(self):
        with np.errstate(invalid="ignore"):
            p = self.log_prior + self.log_likelihood
        return np.where(np.isfinite(p), p, -np.inf)

    def __len__(self):
        return len(self.coords)

    def __iter__(self):
        for c, a, b, acc in zip(self.coords, self.log_prior, self.log_likelihood, self.accepted):
            yield State(c, float(a), float(b), bool(acc))


class Ensemble(object):
    """The walkers (emcee3.Ensemble): ``model`` maps a State to a State with log_prior / log_likelihood filled in;
    all states of a proposal set go through the pool in ONE call.  State is kept as arrays; ``walkers`` builds the
    emcee3-style list of State objects on demand."""

    def __init__(self, model, coords, pool=None, random=None):
        self.model = model
        self.pool = DefaultPool() if pool is None else pool
        self.random = np.random.RandomState() if random is None else random
        first = self.propose(np.atleast_2d(np.asarray(coords, dtype=np.float64)))
        self._coords = first.coords.copy()
        self._log_prior = first.log_prior.copy()
        self._log_likelihood = first.log_likelihood.copy()
        self.acceptance = np.ones(len(first), dtype=bool)
        if not np.all(np.isfinite(self.log_probability)):
            raise ValueError("invalid or zero-probability coordinates for at least one walker")

    def propose(self, coords):
        coords = np.atleast_2d(np.asarray(coords, dtype=np.float64))
        model = self.model
        if isinstance(model, Emcee3Model) and hasattr(model.mod, "lnpost_batch") and isinstance(self.pool, (DefaultPool, EnsemblePool)):
            lp, ll = model.mod.lnpost_batch(coords, return_parts=True)        # the whole proposal set: one launch
            return Proposal(coords, lp, ll)
        states = list(self.pool.map(model, [State(c) for c in coords]))
        return Proposal(coords, [w.log_prior for w in states], [w.log_likelihood for w in states])

    def update(self, proposal, subset=None):
        """Replace the walkers of ``subset`` (boolean mask) by the accepted proposals."""
        idx = np.arange(self.nwalkers) if subset is None else np.flatnonzero(subset)
        if len(idx) != len(proposal):
            raise ValueError("subset and proposal sizes differ")
        acc = np.asarray(proposal.accepted, dtype=bool)
        self.acceptance[idx] = acc
        j = idx[acc]
        self._coords[j] = proposal.coords[acc]
        self._log_prior[j] = proposal.log_prior[acc]
        self._log_likelihood[j] = proposal.log_likelihood[acc]

    def __len__(self):
        return self.nwalkers

    @property
    def walkers(self):
        return [State(c, float(a), float(b), bool(acc))
                for c, a, b, acc in zip(self._coords, self._log_prior, self._log_likelihood, self.acceptance)]

    @property
    def nwalkers(self):
        return self._coords.shape[0]

    @property
    def ndim(self):
        return self._coords.shape[1]

    @property
    def coords(self):
        return self._coords

    @property
    def log_prior(self):
        return self._log_prior

    @property
    def log_likelihood(self):
        return self._log_likelihood

    @property
    def log_probability(self):
        return self._log_prior + self._log_likelihood


# ------------------------------------------------------------------------------------------ moves
def _distinct(random, n, k, rows):
    """``rows`` independent draws of k distinct indices out of n, in random order (uniform over ordered k-tuples)."""
    return np.argsort(random.rand(rows, n), axis=1)[:, :k]


class RedBlueMove(object):
    """Split the walkers into ``nsplits`` groups; each group is moved using the others as the complementary
    ensemble, which keeps detailed balance while the whole group is proposed -- and evaluated -- at once."""

    def __init__(self, nsplits=2, randomize_split=False, live_dangerously=False):
        self.nsplits = int(nsplits)
        self.randomize_split = randomize_split
        self.live_dangerously = live_dangerously

    def get_proposal(self, ensemble, s, c):
        raise NotImplementedError

    def update(self, ensemble):
        nwalkers, ndim = ensemble.nwalkers, ensemble.ndim
        if nwalkers < 2 * ndim and not self.live_dangerously:
            raise RuntimeError("It is unadvisable to use a red-blue move with fewer walkers than twice the number of dimensions.")
        ensemble.acceptance[:] = False
        inds = np.arange(nwalkers) % self.nsplits
        if self.randomize_split:
            ensemble.random.shuffle(inds)
        for i in range(self.nsplits):
            S1 = inds == i
            coords = ensemble.coords
            q, factors = self.get_proposal(ensemble, coords[S1], coords[~S1])
            new = ensemble.propose(q)                                   # ONE batched evaluation
            new_lp = new.log_probability
            with np.errstate(invalid="ignore"):
                lnpdiff = factors + new_lp - ensemble.log_probability[S1]
            new.accepted = np.isfinite(new_lp) & (lnpdiff > np.log(ensemble.random.rand(len(new))))
            ensemble.update(new, S1)
        return ensemble


class StretchMove(RedBlueMove):
    """Goodman & Weare (2010) stretch move with scale ``a``."""

    def __init__(self, a=2.0, **kwargs):
        self.a = a
        super(StretchMove, self).__init__(**kwargs)

    def get_proposal(self, ensemble, s, c):
        Ns, Nc, ndim = len(s), len(c), s.shape[1]
        zz = ((self.a - 1.0) * ensemble.random.rand(Ns) + 1) ** 2.0 / self.a
        rint = ensemble.random.randint(Nc, size=Ns)
        return c[rint] - (c[rint] - s) * zz[:, None], (ndim - 1.0) * np.log(zz)


class DEMove(RedBlueMove):
    """Differential evolution (ter Braak 2006): s + gamma (c_a - c_b), gamma = gamma0 (1 + sigma N(0,1)),
    gamma0 = 2.38 / sqrt(2 ndim) by default.  gprot/fit.py:84 uses ``DEMove(1.0)`` (sigma = 1)."""

    def __init__(self, sigma=1.0e-5, gamma0=None, **kwargs):
        self.sigma = sigma
        self.gamma0 = gamma0
        super(DEMove, self).__init__(**kwargs)

    def get_proposal(self, ensemble, s, c):
        Ns, Nc, ndim = len(s), len(c), s.shape[1]
        g0 = 2.38 / np.sqrt(2 * ndim) if self.gamma0 is None else self.gamma0
        f = self.sigma * ensemble.random.randn(Ns)
        ab = _distinct(ensemble.random, Nc, 2, Ns)
        q = s + (c[ab[:, 1]] - c[ab[:, 0]]) * (g0 * (1 + f))[:, None]
        return q, np.zeros(Ns)


class DESnookerMove(RedBlueMove):
    """Snooker update (ter Braak & Vrugt 2008): move along the line through a third walker z by the projection of
    the difference of two others.  The Hastings factor is (|q - z| / |s - z|)^(d-1), the Jacobian of a move along a line
    whose direction depends on s.  (With half that exponent the kernel under-disperses: variance ratio 0.60 on a 5-d
    Gaussian, tests/test_sampler.py -- the invariance test is the authority here, emcee3 itself cannot be run offline.)"""

    def __init__(self, gammas=1.7, **kwargs):
        self.gammas = gammas
        super(DESnookerMove, self).__init__(**kwargs)

    def get_proposal(self, ensemble, s, c):
        Ns, Nc, ndim = len(s), len(c), s.shape[1]
        w = _distinct(ensemble.random, Nc, 3, Ns)
        z, z1, z2 = c[w[:, 0]], c[w[:, 1]], c[w[:, 2]]
        delta = s - z
        norm = np.sqrt(np.sum(delta * delta, axis=1))
        u = delta / norm[:, None]
        q = s + u * (self.gammas * (np.sum(u * z1, axis=1) - np.sum(u * z2, axis=1)))[:, None]
        metropolis = np.log(np.sqrt(np.sum((q - z) ** 2, axis=1))) - np.log(norm)
        return q, (ndim - 1.0) * metropolis


class KDEMove(RedBlueMove):
    """Independence proposal drawn from a Gaussian KDE of the complementary ensemble (bandwidth: Scott's rule on the
    sample covariance, as scipy.stats.gaussian_kde, or ``bw_method`` = a scalar factor); the Hastings factor is the
    ratio of the KDE densities."""

    def __init__(self, bw_method=None, **kwargs):
        self.bw_method = bw_method
        super(KDEMove, self).__init__(**kwargs)

    @staticmethod
    def _logpdf(x, data, chol):
        # log (1/n) sum_j N(x - data_j; C),  C = chol chol^T.  |L^-1 (x - d)|^2 = |L^-1 x|^2 + |L^-1 d|^2 - 2 (L^-1 x).(L^-1 d):
        # two small matrix products instead of a [m, n, d] triangular solve
        n, d = data.shape
        linv = np.linalg.inv(chol)
        xs, ds = x @ linv.T, data @ linv.T
        e = -0.5 * (np.einsum("ij,ij->i", xs, xs)[:, None] + np.einsum("ij,ij->i", ds, ds)[None, :] - 2.0 * (xs @ ds.T))
        mx = np.max(e, axis=1, keepdims=True)
        lse = mx[:, 0] + np.log(np.sum(np.exp(e - mx), axis=1))
        return lse - np.log(n) - 0.5 * d * np.log(2 * np.pi) - np.sum(np.log(np.diag(chol)))

    def get_proposal(self, ensemble, s, c):
        n, d = c.shape
        factor = n ** (-1.0 / (d + 4)) if self.bw_method is None else float(self.bw_method)
        dc = c - c.mean(axis=0)
        chol = np.linalg.cholesky((dc.T @ dc) * (factor ** 2 / (n - 1.0)))          # = cov(c) * factor^2
        pick = ensemble.random.randint(n, size=len(s))
        q = c[pick] + ensemble.random.randn(len(s), d) @ chol.T
        lp = self._logpdf(np.concatenate([s, q]), c, chol)
        return q, lp[: len(s)] - lp[len(s):]


# ------------------------------------------------------------------------------------------ sampler
class Sampler(object):
    """emcee3.Sampler: a (weighted mixture of) move(s) applied to an Ensemble, results stored in a backend."""

    def __init__(self, moves=None, backend=None):
        if moves is None:
            self._moves, self._weights = [StretchMove()], [1.0]
        elif isinstance(moves, (list, tuple)) and len(moves) and isinstance(moves[0], (list, tuple)):
            self._moves = [m for m, _ in moves]
            self._weights = [float(w) for _, w in moves]
        elif isinstance(moves, (list, tuple)):
            self._moves, self._weights = list(moves), [1.0] * len(moves)
        else:
            self._moves, self._weights = [moves], [1.0]
        self._weights = np.asarray(self._weights) / np.sum(self._weights)
        self._cumw = np.cumsum(self._weights)
        self.backend = Backend() if backend is None else backend

    def reset(self):
        self.backend.reset()

    def sample(self, ensemble, niter=None, store=True, thin=1):
        if store and niter is not None:
            self.backend.extend(niter // thin, ensemble.nwalkers, ensemble.ndim)
        i = 0
        while niter is None or i < niter:
            move = self._moves[min(int(np.searchsorted(self._cumw, ensemble.random.rand(), side="right")), len(self._moves) - 1)]
            ensemble = move.update(ensemble)
            if store and (i + 1) % thin == 0:
                self.backend.update(ensemble)
            i += 1
            yield ensemble

    # ---- the same chain without the interpreter between the launches (csrc/gprot_chain.cpp through the C ABI)
    def _native_moves(self):
        """gprot_chain_moves for this sampler's move mix, or None when a move has no native counterpart."""
        from . import _lib

        mv = _lib.ChainMoves()
        for move, w in zip(self._moves, self._weights):
            if not isinstance(move, RedBlueMove) or move.nsplits != 2 or move.randomize_split or move.live_dangerously:
                return None
            if type(move) is StretchMove:
                mv.w_stretch += w
                mv.stretch_a = float(move.a)
            elif type(move) is DEMove:
                mv.w_de += w
                mv.de_sigma = float(move.sigma)
                mv.de_gamma0 = -1.0 if move.gamma0 is None else float(move.gamma0)
            elif type(move) is DESnookerMove:
                mv.w_snooker += w
                mv.snooker_gamma = float(move.gammas)
            elif type(move) is KDEMove:
                mv.w_kde += w
                mv.kde_bw = -1.0 if move.bw_method is None else float(move.bw_method)
            else:
                return None
        if len(set(type(m) for m in self._moves)) != len(self._moves):       # two moves of one kind with different parameters
            return None
        return mv

    def run_native(self, ensemble, niter, store=True, thin=1):
        """``run`` through gprot_chain_run: red-blue split, moves, acceptance rule and storage as in ``sample``, executed on
        the host side of the C ABI -- a step is two likelihood launches and no interpreter time.  With a GPRotModel on a
        single-device engine the library calls gprot_lnpost_batch itself; any other model with ``lnpost_batch`` is reached
        through a callback.  The random stream is the library's (seeded from ``ensemble.random``): a different realisation
        of the same chain than ``run`` gives.  Raises if the move mix has no native counterpart."""
        from ctypes import POINTER, c_double, cast

        from . import _lib

        mv = self._native_moves()
        if mv is None:
            raise ValueError("this move mix has no native counterpart (use Sampler.run)")
        lib = _lib.load()
        nw, nd = ensemble.nwalkers, ensemble.ndim
        nstore = (niter // thin) if store else 0
        be = self.backend
        if nstore:
            be.extend(nstore, nw, nd)
        coords = np.ascontiguousarray(ensemble._coords, dtype=np.float64)
        lp = np.ascontiguousarray(ensemble._log_prior, dtype=np.float64)
        ll = np.ascontiguousarray(ensemble._log_likelihood, dtype=np.float64)
        nacc = np.zeros(nw, dtype=np.int64)
        i0 = be.niter
        ptr = lambda a: a.ctypes.data + i0 * a.strides[0] if nstore else None      # noqa: E731
        chain_args = (coords.ctypes.data, lp.ctypes.data, ll.ctypes.data, ptr(be._coords), ptr(be._log_prior), ptr(be._log_likelihood),
                      nacc.ctypes.data)
        seed = int(ensemble.random.randint(0, 2 ** 31 - 1)) * 2 ** 31 + int(ensemble.random.randint(0, 2 ** 31 - 1))
        model = ensemble.model
        mod = getattr(model, "mod", None)
        eng = None
        if mod is not None and hasattr(mod, "_prior_args"):                # a GPRotModel: its engine (created on first use)
            eng = mod.engine
        import ctypes as _ct
        if nd == 5 and eng is not None and hasattr(eng, "_ctx") and hasattr(mod, "_prior_args"):
            b, mu, sig, mix = mod._prior_args()
            mu, sig = np.ascontiguousarray(mu, dtype=np.float64), np.ascontiguousarray(sig, dtype=np.float64)
            mixc = None if mix is None or len(mix) == 0 else np.ascontiguousarray(mix, dtype=np.float64).reshape(-1, 3)
            rc = lib.gprot_chain_run(eng._ctx, int(mod._registered_id()), b.ctypes.data, mu.ctypes.data, sig.ctypes.data, float(mod.pmin),
                                     float(mod.pmax), mixc.ctypes.data if mixc is not None else None, 0 if mixc is None else mixc.shape[0],
                                     nw, int(niter), int(thin), _ct.byref(mv), seed, *chain_args)
        else:
            target = mod if mod is not None and hasattr(mod, "lnpost_batch") else None
            if target is None:
                raise ValueError("run_native needs a model with lnpost_batch")

            failure = []

            def fn(user, n, theta, out_lp, out_ll):
                if failure:                                   # an earlier call raised: the outputs stay -inf, nothing is accepted
                    return
                try:
                    th = np.ctypeslib.as_array(theta, shape=(n, nd))
                    a, c = target.lnpost_batch(th.copy(), return_parts=True)
                    np.ctypeslib.as_array(out_lp, shape=(n,))[:] = a
                    np.ctypeslib.as_array(out_ll, shape=(n,))[:] = c
                except BaseException as exc:                  # an exception cannot cross the C frames: keep it, raise it below
                    failure.append(exc)

            cb = _lib.LNPOST_FN(fn)
            rc = lib.gprot_chain_run_fn(cast(cb, _ct.c_void_p), None, nw, nd, int(niter), int(thin), _ct.byref(mv), seed, *chain_args)
            if failure:
                raise failure[0]
        if rc != 0:
            raise RuntimeError(lib.gprot_chain_last_error().decode())
        ensemble._coords[...] = coords
        ensemble._log_prior[...] = lp
        ensemble._log_likelihood[...] = ll
        ensemble.acceptance[:] = False
        if nstore:
            be._accepted += nacc
            be.niter += nstore
        if hasattr(be, "save"):
            be.save()
        return ensemble

    def run(self, ensemble, niter, progress=False, native=False, **kwargs):
        if native:
            return self.run_native(ensemble, niter, **kwargs)
        it = self.sample(ensemble, niter, **kwargs)
        if progress:
            try:
                from tqdm import tqdm

                it = tqdm(it, total=niter)
            except Exception:
                pass
        for ensemble in it:
            pass
        if hasattr(self.backend, "save"):
            self.backend.save()
        return ensemble

    # ---- chain access (delegated to the backend, as emcee3 does)
    def get_coords(self, **kwargs):
        return self.backend.get_coords(**kwargs)

    def get_log_probability(self, **kwargs):
        return self.backend.get_log_probability(**kwargs)

    def get_integrated_autocorr_time(self, **kwargs):
        return self.backend.get_integrated_autocorr_time(**kwargs)

    @property
    def acceptance_fraction(self):
        return self.backend.acceptance_fraction
