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


def write_samples(mod, df, resultsdir='results', true_period=None):
    """gprot/fit.py:24-51.  The reference writes ``results/{name}.h5`` with pandas-HDF keys ``samples``, ``lc``,
    ``gp_prior`` and ``period_prior``; pandas' HDF writer needs PyTables, which this image lacks, so the same four
    tables go to HDF when PyTables is importable and otherwise to ``results/{name}.npz`` (arrays under the same
    keys, column names under ``<key>_columns``).  The corner plot (matplotlib) is not part of this path."""
    import os

    import pandas as pd

    if not os.path.exists(resultsdir):
        os.makedirs(resultsdir)
    prior_df = pd.DataFrame({'mu': mod.gp_prior_mu, 'sigma': mod.gp_prior_sigma})
    tables = {'samples': df, 'lc': mod.lc.df, 'gp_prior': prior_df}
    if mod.acf_prior:
        tables['period_prior'] = pd.DataFrame(mod.period_mixture, columns=['w', 'mu', 'sigma'])
    try:
        import tables as _pytables  # noqa: F401

        samplefile = os.path.join(resultsdir, '{}.h5'.format(mod.name))
        for key, t in tables.items():
            t.to_hdf(samplefile, key=key)
    except ImportError:
        samplefile = os.path.join(resultsdir, '{}.npz'.format(mod.name))
        out = {}
        for key, t in tables.items():
            out[key] = t.values
            out[key + '_columns'] = np.array(list(t.columns))
        np.savez(samplefile, **out)
    print('Samples, light curve, and prior saved to {}.'.format(samplefile))
    return samplefile


def read_samples(samplefile, key='samples'):
    """Counterpart of ``pd.read_hdf(file, key)`` in the reference's downstream scripts (gprot/summary.py:38-56)."""
    import pandas as pd

    if samplefile.endswith('.npz'):
        with np.load(samplefile, allow_pickle=False) as z:
            return pd.DataFrame(z[key], columns=[str(c) for c in z[key + '_columns']])
    return pd.read_hdf(samplefile, key)


def _chain_store(mod, sample_directory):
    """Where the chain lives between calls: ``<sample_directory>/<name>.h5`` in the reference (emcee3 HDFBackend,
    gprot/fit.py:69-72); the sibling .npz here.  ``None`` keeps the chain in memory only."""
    import os

    from . import sampler as em

    if sample_directory is None:
        return em.Backend()
    os.makedirs(sample_directory, exist_ok=True)
    return em.HDFBackend(os.path.join(sample_directory, '{}.h5'.format(mod.name)))


def _starting_coords(mod, store, nwalkers, fresh, seed):
    """Resume from the last stored ensemble when there is one (gprot/fit.py:73-76), else draw from the prior."""
    if not fresh:
        try:
            return store.current_coords
        except (AttributeError, KeyError):
            pass
    return mod.sample_from_prior(nwalkers, seed=seed)


def _chain_length_in_autocorr_times(sampler, nburn, nwalkers, verbose):
    """(tau_max, chain length in units of tau_max minus the burn-in allowance): the convergence statistic of
    gprot/fit.py:100-110.  Raises AutocorrError while the chain is too short to tell."""
    tau_max = sampler.get_integrated_autocorr_time(c=1).max()
    neff = sampler.backend.niter / tau_max - nburn
    if verbose:
        print("Maximum autocorrelation time: {0}".format(tau_max))
        print("N_eff: {0}\n".format(neff * nwalkers))
    return tau_max, neff


def fit_emcee3(mod, nwalkers=500, verbose=False, nsamples=5000, targetn=6,
               iter_chunksize=10, pool=None, overwrite=False,
               maxiter=100, sample_directory='mcmc_chains',
               nburn=3, mixedmoves=True, resultsdir='results', seed=None, native=None, **kwargs):
    """The reference's fit driver (gprot/fit.py:53-152), same arguments and behaviour: resume from the stored chain
    unless ``overwrite``; a 0.4 / 0.4 / 0.2 mixture of KDE, DE(1.0) and DE-snooker moves (KDE alone without
    ``mixedmoves``); blocks of ``iter_chunksize`` steps, at most ``maxiter`` of them, until the chain is ``targetn``
    autocorrelation times long after ``nburn`` of them are set aside; ``nsamples`` random draws from what is left,
    written with ``write_samples`` and returned as a DataFrame.  ``pool`` defaults to ``EnsemblePool``: every
    half-ensemble proposal set is ONE device launch (the reference's DefaultPool / schwimmbad pool evaluates one walker
    per call).  ``seed`` (not in the reference) makes a run reproducible.  ``native`` (not in the reference; default: the
    environment variable GPROT_NATIVE_CHAIN=1) runs the blocks through ``Sampler.run_native`` -- the same chain driven from
    the host side of the C ABI (csrc/gprot_chain.cpp), with the library's own random stream."""
    import pandas as pd

    from . import sampler as em

    import os

    if native is None:
        native = os.environ.get("GPROT_NATIVE_CHAIN", "0") not in ("", "0")
    native = bool(native) and pool is None
    rs = np.random.RandomState(seed)
    store = _chain_store(mod, sample_directory)
    sampler = em.Sampler([(em.KDEMove(), 0.4), (em.DEMove(1.0), 0.4), (em.DESnookerMove(), 0.2)] if mixedmoves else em.KDEMove(),
                         backend=store)
    if overwrite:
        sampler.reset()
    ensemble = em.Ensemble(Emcee3Model(mod), _starting_coords(mod, store, nwalkers, overwrite, seed),
                           pool=EnsemblePool() if pool is None else pool, random=rs)

    tau_max, converged = 0, False
    if not overwrite:                                   # a chain left by an earlier call may already be long enough
        if verbose:
            print('Status from previous run:')
        try:
            tau_max, neff = _chain_length_in_autocorr_times(sampler, nburn, nwalkers, verbose)
            converged = neff > targetn
        except (em.AutocorrError, KeyError, AttributeError):
            pass

    blocks = 0
    while not converged and blocks < maxiter:
        blocks += 1
        if verbose:
            print("Iteration {0}...".format(blocks))
        sampler.run(ensemble, iter_chunksize, progress=verbose, native=native)
        try:
            tau_max, neff = _chain_length_in_autocorr_times(sampler, nburn, nwalkers, verbose)
        except em.AutocorrError:
            tau_max = 0
            continue
        converged = neff > targetn

    burnin = int(nburn * tau_max)
    flat = sampler.get_coords(flat=True, discard=burnin)
    keep = min(nsamples, len(flat))
    if verbose:
        print("Discarding {0} samples for burn-in".format(burnin))
        print("Randomly choosing {0} samples".format(keep))
    df = pd.DataFrame(flat[rs.choice(len(flat), size=keep, replace=False)], columns=mod.param_names)
    write_samples(mod, df, resultsdir=resultsdir)
    return df
