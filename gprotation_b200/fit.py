"""Walker adapter: the seam where the reference hands proposals to the likelihood.

Reference: ``Emcee3Model`` (gprot/fit.py:11-22) maps ONE walker state to ``mod.lnprior`` /
``mod.lnlike``; emcee3 fans the walkers out with ``pool.map`` (gprot/fit.py:98,128, pool from
``schwimmbad.choose_pool``, scripts/gprot-fit:195).  Here the same two methods exist with the same
meaning, and ``EnsemblePool.map`` turns the list of walker states of one proposal round into ONE
batched launch -- the device boundary replaces the process boundary.

emcee3 itself is not installable offline, so ``State`` below is the minimal
stand-in for ``emcee3.State`` (coords, log_prior, log_likelihood, log_probability).
"""
from __future__ import annotations

import numpy as np


class State(object):
    """Minimal walker state with the attribute names emcee3 uses."""
    __slots__ = ("coords", "log_prior", "log_likelihood", "accepted")

    def __init__(self, coords, log_prior=-np.inf, log_likelihood=-np.inf, accepted=False):
        self.coords = np.asarray(coords, dtype=np.float64)
        self.log_prior = log_prior
        self.log_likelihood = log_likelihood
        self.accepted = accepted

    @property
    def log_probability(self):
        lp = self.log_prior + self.log_likelihood
        return lp if np.isfinite(lp) else -np.inf


class Emcee3Model(object):
    """Same surface as gprot/fit.py:11-22; ``__call__`` follows emcee3.Model: prior first, likelihood only
    where the prior is finite."""

    def __init__(self, mod):
        self.mod = mod

    def compute_log_prior(self, state):
        state.log_prior = self.mod.lnprior(state.coords)
        return state

    def compute_log_likelihood(self, state):
        state.log_likelihood = self.mod.lnlike(state.coords)
        return state

    def __call__(self, state):
        state = self.compute_log_prior(state)
        if not np.isfinite(state.log_prior):
            state.log_prior = -np.inf
            state.log_likelihood = -np.inf
            return state
        state = self.compute_log_likelihood(state)
        if not np.isfinite(state.log_likelihood):
            state.log_likelihood = -np.inf
        return state

    # batched form: every state of a proposal round in one launch
    def evaluate_states(self, states):
        states = list(states)
        if not states:
            return states
        thetas = np.stack([s.coords for s in states])
        lp, ll = self.mod.lnpost_batch(thetas, return_parts=True)
        for s, a, b in zip(states, lp, ll):
            ok = np.isfinite(a)
            s.log_prior = float(a) if ok else -np.inf
            s.log_likelihood = float(b) if ok and np.isfinite(b) else -np.inf
        return states


class EnsemblePool(object):
    """Pool-shaped adapter (schwimmbad / emcee3.pools interface: ``map(func, iterable)``).  When ``func`` is an
    ``Emcee3Model`` the whole iterable becomes one device launch; anything else is mapped serially."""

    def map(self, func, iterable):
        if isinstance(func, Emcee3Model):
            return func.evaluate_states(iterable)
        return [func(x) for x in iterable]

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
