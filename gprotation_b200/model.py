"""GPRotModel: the reference's model surface (gprot/model.py:32-396) on the B200 engine.

Same constructor, attributes and method meanings as the reference, so ``gprot/fit.py`` and the
nested-sampling shims can be pointed at this class unchanged:

    lnprior(theta) / lnlike(theta) / lnpost(theta) / lnlike_function(theta, x, y, yerr)
    sample_from_prior(N, seed) / sample_period_prior(N) / bounds / param_names / ndim / x / y / yerr / name
    polychord_prior / polychord_lnpost / mnest_prior / mnest_loglike

plus the batched siblings the reference does not have (a whole emcee3 ensemble per launch):

    lnprior_batch(thetas) / lnlike_batch(thetas) / lnpost_batch(thetas)

theta = (ln A, ln l, ln G, ln sigma, ln P).  The likelihood is evaluated by the CUDA engine only
(no george, no CPU fallback); the covariance is the one the reference's code builds, including the
white-noise term counted twice (gprot/model.py:334 and :345-346).
"""
from __future__ import annotations

import numpy as np

from . import engine as _engine

_ENGINES = {}


def get_engine(device=0):
    """Process-wide engine for a device (one context per (process, GPU))."""
    eng = _ENGINES.get(device)
    if eng is None:
        eng = _engine.GPEngine(device)
        _ENGINES[device] = eng
    return eng


def lnGauss(x, mu, sigma):
    """gprot/model.py:22-23."""
    return -0.5 * ((x - mu) ** 2 / (sigma ** 2)) + np.log(1.0 / np.sqrt(2 * np.pi * sigma ** 2))


def lnGauss_mixture(x, mix):
    """gprot/model.py:25-29: log sum_j w_j N(x; mu_j, sigma_j)."""
    mix = np.asarray(mix, dtype=np.float64)
    a = lnGauss(x, mix[:, 1], mix[:, 2])
    m = np.max(a)
    if not np.isfinite(m):
        m = 0.0
    return float(np.log(np.sum(mix[:, 0] * np.exp(a - m))) + m)


class GPRotModel(object):
    """Parameters are A, l, G, sigma, period (natural logs)."""

    # gprot/model.py:36-45
    _default_bounds = ((-20., 0.),
                       (2, 20.),
                       (-10., 3.),
                       (-20., 0.),
                       (-0.69, 4.61))  # 0.5 - 100 d

    param_names = ('ln_A', 'ln_l', 'ln_G', 'ln_sigma', 'ln_period')

    _default_gp_prior_mu = (-13, 7.2, -2.3, -17)
    _default_gp_prior_sigma = (5.7, 1.2, 1.4, 5)

    _acf_pmax = (1, 2, 4, 8, 16, 32, 64, 128)
    _acf_prior_width = 0.2

    def __init__(self, lc, name=None, pmin=None, pmax=None, acf_prior=False,
                 gp_prior_mu=None, gp_prior_sigma=None, bounds=None,
                 engine=None, device=0, period_mixture=None):
        self.lc = lc
        self._name = name
        self.pmin = -np.inf if pmin is None else pmin
        self.pmax = np.inf if pmax is None else pmax
        self.acf_prior = acf_prior
        self.gp_prior_mu = np.array(self._default_gp_prior_mu if gp_prior_mu is None else gp_prior_mu, dtype=np.float64)
        self.gp_prior_sigma = np.array(self._default_gp_prior_sigma if gp_prior_sigma is None else gp_prior_sigma,
                                       dtype=np.float64)
        self._bounds = self._default_bounds if bounds is None else bounds
        self._engine = engine
        self._device = device
        self._lc_id = None
        self._lc_key = None
        self._adhoc = {}
        self._adhoc_generation = -1
        self._adhoc_max = 32
        if period_mixture is not None:
            self._period_mixture = np.asarray(period_mixture, dtype=np.float64).reshape(-1, 3)

    # ---- reference properties (gprot/model.py:84-109)
    @property
    def ndim(self):
        return len(self.param_names)

    @property
    def x(self):
        return self.lc.x

    @property
    def y(self):
        return self.lc.y

    @property
    def yerr(self):
        return self.lc.yerr

    @property
    def name(self):
        return self.lc.name if self._name is None else self._name

    @property
    def bounds(self):
        return self._bounds

    @property
    def acf_pmax(self):
        return self._acf_pmax

    @property
    def acf_prior_width(self):
        return self._acf_prior_width

    @property
    def period_mixture(self):
        """Gaussian mixture (w, mu, sigma) on ln P.  The reference builds it from eight filtered ACFs of the
        light curve (gprot/model.py:288-326 -> gprot/lc.py:127-224, out of scope here); pass it as
        ``period_mixture=`` or assign it before using ``acf_prior=True``."""
        if not hasattr(self, '_period_mixture'):
            raise NotImplementedError(
                "acf_prior=True needs a period mixture; the ACF pre-processing that derives it is outside the "
                "hot path (SURVEY.md section 2, rows 5, 8, 9). Supply period_mixture=[(w, mu, sigma), ...].")
        return self._period_mixture

    @period_mixture.setter
    def period_mixture(self, value):
        self._period_mixture = np.asarray(value, dtype=np.float64).reshape(-1, 3)

    # ---- engine plumbing
    @property
    def engine(self):
        if self._engine is None:
            self._engine = get_engine(self._device)
        return self._engine

    @staticmethod
    def _digest(*arrays):
        """Content key of a set of arrays (id() is reused after garbage collection and blind to in-place edits)."""
        import hashlib

        h = hashlib.blake2b(digest_size=16)
        for a in arrays:
            a = np.ascontiguousarray(a, dtype=np.float64)
            h.update(np.int64(a.size).tobytes())
            h.update(a.tobytes())
        return h.digest()

    def _registered_id(self):
        """Upload the light curve once (the reference hands x, yerr to george on every call).  The registration is
        keyed on the light curve's change counter (or, for foreign objects, its content) and on the engine's generation
        (engine.unregister_all voids every id)."""
        lc = self.lc
        xl = lc.x_list
        version = getattr(lc, "version", None)
        if version is not None:
            # gprotation_b200.lc.LightCurve counts every change of its arrays / chunking; the model holds the object, so
            # its id cannot be recycled (unlike id(lc.x), which a new array may inherit after garbage collection)
            ident = (id(lc), version, len(lc.x))
        else:                                                # a foreign light-curve object: key on the content
            arrays = (lc.x, lc.y, lc.yerr) if xl is None else tuple(xl) + tuple(lc.y_list) + tuple(lc.yerr_list)
            ident = self._digest(*arrays)
        key = (ident, None if xl is None else len(xl), self.engine.generation)
        if self._lc_id is None or key != self._lc_key:
            if self._lc_id is not None and self._lc_key is not None and self._lc_key[2] == self.engine.generation:
                self.engine.unregister_lc(self._lc_id)            # the light curve changed: drop the stale copy
            if xl is None:
                self._lc_id = self.engine.register_lc(lc.x, lc.y, lc.yerr)
            else:
                self._lc_id = self.engine.register_lc_arrays(xl, lc.y_list, lc.yerr_list)
            self._lc_key = key
        return self._lc_id

    # ---- prior (gprot/model.py:111-163, 227-257)
    def lnprior(self, theta):
        for i, (lo, hi) in enumerate(self.bounds):
            if theta[i] < lo or theta[i] > hi:
                return -np.inf
        if theta[-1] < self.pmin or theta[-1] > self.pmax:
            return -np.inf
        if theta[1] < theta[-1]:          # SE correlation length no shorter than P
            return -np.inf
        lnpr = np.sum(lnGauss(np.array(theta[:-1], dtype=np.float64), self.gp_prior_mu, self.gp_prior_sigma))
        lnpr += self.lnprior_period(theta[-1])
        return lnpr

    def lnprior_period(self, p):
        if not self.acf_prior:
            return 0
        return lnGauss_mixture(p, self.period_mixture)

    def sample_period_prior(self, N):
        loP, hiP = self.bounds[-1]
        if self.pmin > loP:
            loP = self.pmin
        if self.pmax < hiP:
            hiP = self.pmax
        if not self.acf_prior:
            return np.random.random(N) * (hiP - loP) + loP
        vals = np.empty(N)
        u = np.random.random(N)
        lo = 0
        for w, mu, sig in self.period_mixture:
            hi = lo + w
            ok = (u > lo) & (u < hi)
            n = ok.sum()
            newvals = np.inf * np.ones(n)
            m = (newvals < loP) | (newvals > hiP)
            nbad = m.sum()
            while nbad > 0:
                newvals[m] = np.random.randn(nbad) * sig + mu
                m = (newvals < loP) | (newvals > hiP)
                nbad = m.sum()
            vals[ok] = newvals
            lo += w
        return vals

    def sample_from_prior(self, N, seed=None):
        """N x ndim prior samples inside the support (same global-RNG draw order as the reference)."""
        samples = np.inf * np.ones((N, self.ndim))
        np.random.seed(seed)
        m = np.ones(N, dtype=bool)
        nbad = m.sum()
        while nbad > 0:
            rn = np.random.randn(N * (self.ndim - 1)).reshape((N, self.ndim - 1))
            for i in range(self.ndim - 1):
                samples[m, i] = rn[m, i] * self.gp_prior_sigma[i] + self.gp_prior_mu[i]
            samples[m, -1] = self.sample_period_prior(nbad)
            lnp = self.lnprior_batch(samples)
            m = ~np.isfinite(lnp)
            nbad = m.sum()
        return samples

    # ---- likelihood (gprot/model.py:328-369)
    def gp_kernel(self, theta):
        """The reference returns a george kernel object; here the five linear-space hyper-parameters of
        A * ExpSquaredKernel(l) * ExpSine2Kernel(G, P) + WhiteKernel(sigma) that the device kernel consumes."""
        with np.errstate(over='ignore'):
            A, l, G, sigma, P = (float(np.exp(v)) for v in theta)
        return {'A': A, 'metric': l, 'gamma': G, 'white': sigma, 'period': P}

    def gp(self, theta, x=None, yerr=None):
        """gprot/model.py:336-347 returns ``george.GP(kernel, solver=HODLRSolver)`` after ``compute(x, sqrt(sigma + yerr**2))``.
        Here: a handle with the two george.GP methods the reference calls on it -- ``lnlikelihood(y, quiet=True)``
        (gprot/model.py:354) and ``predict(y, t)`` (code/koi_demo.py:66) -- evaluated by the CUDA engine."""
        return GPHandle(self, theta, self.x if x is None else x, self.yerr if yerr is None else yerr)

    def lnlike_function(self, theta, x, y, yerr):
        # ad-hoc arrays are registered once per content and kept in a small LRU (the arena must not grow without bound);
        # an engine.unregister_all() elsewhere voids the cached ids (generation)
        if self._adhoc_generation != self.engine.generation:
            self._adhoc.clear()
            self._adhoc_generation = self.engine.generation
        key = self._digest(x, y, yerr)
        lc_id = self._adhoc.pop(key, None)
        if lc_id is None:
            lc_id = self.engine.register_lc(x, y, yerr)
            while len(self._adhoc) >= self._adhoc_max:
                old_key = next(iter(self._adhoc))
                self.engine.unregister_lc(self._adhoc.pop(old_key))
        self._adhoc[key] = lc_id                                 # most recently used last
        lnl = float(self.engine.lnlike_batch(np.asarray(theta, dtype=np.float64).reshape(1, 5), lc_id=lc_id)[0])
        return lnl if np.isfinite(lnl) else -np.inf

    def lnlike(self, theta):
        """One problem, or the sum over the chunk list -- the device sums the chunks of a registered light curve."""
        lnl = float(self.engine.lnlike_batch(np.asarray(theta, dtype=np.float64).reshape(1, 5),
                                             lc_id=self._registered_id())[0])
        return lnl if np.isfinite(lnl) else -np.inf

    def lnpost(self, theta):
        return self.lnlike(theta) + self.lnprior(theta)

    # ---- batched siblings: a whole ensemble per launch
    def _prior_args(self):
        mix = self.period_mixture if self.acf_prior else None
        b = getattr(self, "_bounds_flat", None)
        if b is None or self._bounds_src is not self._bounds:          # flattened once per bounds object, not per call
            b = self._bounds_flat = np.ascontiguousarray(np.asarray(self.bounds, dtype=np.float64).reshape(10))
            self._bounds_src = self._bounds
        return b, self.gp_prior_mu, self.gp_prior_sigma, mix

    def lnprior_batch(self, thetas):
        thetas = np.ascontiguousarray(thetas, dtype=np.float64).reshape(-1, 5)
        b, mu, sig, mix = self._prior_args()
        out = np.empty(thetas.shape[0])
        lib = self.engine._lib if self._engine is not None else _engine._lib.load()
        mixc = None if mix is None else np.ascontiguousarray(mix, dtype=np.float64)
        rc = lib.gprot_lnprior_batch(thetas.shape[0], thetas.ctypes.data, b.ctypes.data,
                                     np.ascontiguousarray(mu).ctypes.data, np.ascontiguousarray(sig).ctypes.data,
                                     float(self.pmin), float(self.pmax), None if mixc is None else mixc.ctypes.data,
                                     0 if mixc is None else mixc.shape[0], out.ctypes.data)
        if rc != 0:
            raise _engine.GProtError("gprot_lnprior_batch failed")
        return out

    def lnlike_batch(self, thetas, exact=False):
        return self.engine.lnlike_batch(thetas, lc_id=self._registered_id(), exact=exact)

    def lnpost_batch(self, thetas, return_parts=False):
        """lnprior + lnlike per row; rows outside the prior support are not factorised (emcee3 never evaluates
        their likelihood) and come back as -inf."""
        b, mu, sig, mix = self._prior_args()
        lp, ll = self.engine.lnpost_batch(thetas, b, mu, sig, self.pmin, self.pmax, mix, lc_id=self._registered_id())
        if return_parts:
            return lp, ll
        with np.errstate(invalid='ignore'):
            post = lp + ll
        post[~np.isfinite(lp)] = -np.inf
        return post

    # ---- predictive distribution (what the reference's plotting scripts get from george's gp.predict, code/koi_demo.py:66)
    def predict(self, theta, t_pred, return_var=True, device=None):
        """GP predictive mean (and variance) at ``t_pred`` for hyper-parameters ``theta``, computed with the batched
        likelihood itself: the log-likelihood of the data extended by one point (t*, y*) is exactly quadratic in y*,
        ``lnL(y*) = const - (y* - mu*)^2 / (2 s*^2)``, so three evaluations at y* = -s, 0, +s give mu* and s*^2.  Every test
        point is three (N+1)-point problems; all of them go to the device in ONE launch.  The variance returned is
        george's: k(t*, t*) minus the explained part, i.e. without the measurement noise of a new observation.  With a
        chunked light curve the prediction uses the chunk whose time span is nearest to t* (chunks are independent in
        the reference's likelihood)."""
        theta = np.asarray(theta, dtype=np.float64).reshape(5)
        t_pred = np.atleast_1d(np.asarray(t_pred, dtype=np.float64))
        A, L, G, S, P = np.exp(theta)
        s = float(np.sqrt(A + S))
        lc = self.lc
        if lc.x_list is None:
            groups = [(np.asarray(lc.x), np.asarray(lc.y), np.asarray(lc.yerr))]
        else:
            groups = [(np.asarray(a), np.asarray(b), np.asarray(c)) for a, b, c in zip(lc.x_list, lc.y_list, lc.yerr_list)]
        eng = self.engine if device is None else get_engine(device)     # the model's own context; the probes are dropped again
        ids = []
        try:
            for ts in t_pred:
                gap = [max(g[0].min() - ts, ts - g[0].max(), 0.0) for g in groups]
                x, y, e = groups[int(np.argmin(gap))]
                xa, ea = np.append(x, ts), np.append(e, 0.0)
                for ystar in (-s, 0.0, s):
                    ids.append(eng.register_lc(xa, np.append(y, ystar), ea))
            f = eng.lnlike_batch(np.tile(theta, (len(ids), 1)), lc_id=np.asarray(ids, dtype=np.int32)).reshape(-1, 3)
        finally:
            for i in ids:
                eng.unregister_lc(i)
        curv = (f[:, 2] + f[:, 0] - 2.0 * f[:, 1]) / (s * s)        # = -1 / s*^2
        slope = (f[:, 2] - f[:, 0]) / (2.0 * s)                     # = mu* / s*^2
        with np.errstate(divide='ignore', invalid='ignore'):
            var_obs = -1.0 / curv
            mu = slope * var_obs
        if not return_var:
            return mu
        return mu, var_obs - S                                       # the extended point carried S + 0^2 of extra noise

    # ---- nested-sampling shims (gprot/model.py:371-388)
    def polychord_prior(self, cube):
        return [lo + (hi - lo) * cube[i] for i, (lo, hi) in enumerate(self.bounds)]

    def polychord_lnpost(self, theta):
        phi = [0.0] * 0
        return self.lnpost(theta), phi

    def mnest_prior(self, cube, ndim, nparams):
        for i in range(5):
            lo, hi = self.bounds[i]
            cube[i] = (hi - lo) * cube[i] + lo

    def mnest_loglike(self, cube, ndim, nparams):
        return self.lnpost(cube)


class GPHandle(object):
    """The part of the george.GP surface GProtation touches, for one (theta, x, yerr): see ``GPRotModel.gp``."""

    def __init__(self, model, theta, x, yerr):
        self.model = model
        self.theta = np.asarray(theta, dtype=np.float64).reshape(5)
        self.x = np.asarray(x, dtype=np.float64)
        self.yerr = np.asarray(yerr, dtype=np.float64)
        self.kernel = model.gp_kernel(self.theta)
        self.computed = True

    def lnlikelihood(self, y, quiet=False):
        """george.GP.lnlikelihood: -inf on failure when ``quiet``, else numpy.linalg.LinAlgError (as george raises)."""
        lnl = self.model.lnlike_function(self.theta, self.x, y, self.yerr)
        if not np.isfinite(lnl) and not quiet:
            raise np.linalg.LinAlgError("covariance matrix is not positive definite (or not finite)")
        return lnl

    def predict(self, y, t, return_var=True):
        """Predictive mean (and variance) at ``t`` given ``y`` at the handle's x.  george returns the full predictive
        covariance; the reference only plots mean and standard deviation, so only the diagonal is provided."""
        from .lc import LightCurve

        probe = GPRotModel(LightCurve(self.x, y, self.yerr, chunksize=None), engine=self.model.engine)
        try:
            return probe.predict(self.theta, t, return_var=return_var)
        finally:
            if probe._lc_id is not None:
                probe.engine.unregister_lc(probe._lc_id)


class KeplerGPRotModel(GPRotModel):
    """Bounds and prior means adjusted to the Kepler population (gprot/kepler.py:21-34)."""
    _default_bounds = ((-20., 0.),
                       (2, 8.),
                       (0., 3.),
                       (-20., 0.),
                       (-0.69, 4.61))
    _default_gp_prior_mu = (-13, 5.0, 1.9, -17)
    _default_gp_prior_sigma = (5.7, 1.2, 1.4, 5)
