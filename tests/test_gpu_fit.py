"""End to end on the GPU: GPRotModel -> Emcee3Model -> Ensemble / Sampler -> fit_emcee3 -> results, and the gprot-fit
command line on an Aigrain-layout data set.  Every log-likelihood the chain stores must be the oracle's."""
import os

import numpy as np
import pytest

from oracle import gp_oracle as o
from gprotation_b200 import sampler as em
from gprotation_b200.fit import Emcee3Model, EnsemblePool, fit_emcee3, read_samples
from gprotation_b200.lc import LightCurve
from gprotation_b200.model import GPRotModel
from gprotation_b200.synth import synthetic_lightcurve, write_aigrain_dataset

pytestmark = pytest.mark.gpu


def _model(engine, n=450, chunksize=150, seed=21):
    t, y, yerr, truth = synthetic_lightcurve(seed, n, baseline=200.0, kind="harmonic")
    lc = LightCurve(t, y, yerr, name="star%d" % seed, chunksize=chunksize)
    return GPRotModel(lc, engine=engine), (t, y, yerr, truth)


def test_chain_log_likelihoods_match_oracle(engine):
    engine.unregister_all()
    mod, (t, y, yerr, truth) = _model(engine)
    rs = np.random.RandomState(3)
    ens = em.Ensemble(Emcee3Model(mod), mod.sample_from_prior(32, seed=5), pool=EnsemblePool(), random=rs)
    l0 = engine.launch_count()
    sam = em.Sampler([(em.KDEMove(), 0.4), (em.DEMove(1.0), 0.4), (em.DESnookerMove(), 0.2)])
    sam.run(ens, 15)
    # two proposal sets per step, each ONE likelihood launch (+ its row reduction)
    assert engine.launch_count() - l0 == 15 * 2 * 2
    chunks = o.split_chunks(len(t), 150)
    coords = sam.get_coords()
    ll = sam.backend.log_likelihood
    lp = sam.backend.log_prior
    for it in (0, 7, 14):
        for w in (0, 13, 31):
            ref = o.lnlike(coords[it, w], t, y, yerr, chunks)
            assert abs(ll[it, w] - ref) <= 1e-9 * abs(ref)
            assert lp[it, w] == pytest.approx(o.lnprior(coords[it, w]), rel=1e-12)
    assert np.all(np.isfinite(sam.get_log_probability()))


def test_fit_emcee3_on_gpu_model_and_resume(engine, tmp_path):
    engine.unregister_all()
    mod, _ = _model(engine, n=300, chunksize=None, seed=22)      # --nochunks: one N x N problem per walker
    kw = dict(nwalkers=24, nsamples=200, targetn=3, iter_chunksize=40, maxiter=2, nburn=1,
              sample_directory=str(tmp_path / "mcmc_chains"), resultsdir=str(tmp_path / "results"), seed=9)
    df = fit_emcee3(mod, **kw)
    assert list(df.columns) == list(mod.param_names) and 0 < len(df) <= 200
    lo, hi = np.array(mod.bounds).T
    assert np.all(df.values >= lo) and np.all(df.values <= hi) and np.all(df['ln_l'] >= df['ln_period'])
    chain = em.HDFBackend(os.path.join(str(tmp_path / "mcmc_chains"), mod.name + ".h5"))
    assert chain.niter in (40, 80)
    files = os.listdir(str(tmp_path / "results"))
    assert len(files) == 1
    assert len(read_samples(os.path.join(str(tmp_path / "results"), files[0]), "lc")) == 300


def test_gprot_fit_cli_on_aigrain_layout(engine, tmp_path, monkeypatch):
    engine.unregister_all()
    d = write_aigrain_dataset(str(tmp_path / "sims"), 1, baseline=120.0)
    monkeypatch.setenv("AIGRAIN_ROTATION", d)
    monkeypatch.chdir(tmp_path)
    from gprotation_b200 import cli, model

    monkeypatch.setitem(model._ENGINES, 0, engine)
    rc = cli.main("0 --aigrain --subsample 12 --chunksize 200 --nwalkers 20 --iter_chunksize 10 --maxiter 1 --seed 2".split())
    assert rc == 0
    assert any(f.startswith("0.") for f in os.listdir(str(tmp_path / "results")))
    assert os.path.exists(str(tmp_path / "mcmc_chains" / "0.npz"))
    # a star that does not exist is logged and skipped, not raised (scripts/gprot-fit:210-213)
    assert cli.main("9 --aigrain --nwalkers 20 --maxiter 1".split()) == 1


def test_predict_through_the_likelihood(engine):
    """GPRotModel.predict (three batched likelihoods per test point) against the dense textbook formulas with the
    matrix the reference's code builds: mu* = k*^T K^-1 y, var* = k(t*,t*) - k*^T K^-1 k*."""
    t, y, yerr, truth = synthetic_lightcurve(31, 260, baseline=120.0, kind="gp")
    mod = GPRotModel(LightCurve(t, y, yerr, name="p", chunksize=None), engine=engine)
    tp = np.array([t[0] - 3.0, 0.5 * (t[100] + t[101]), t[57], t[-1] + 1.5, 60.123])
    mu, var = mod.predict(truth, tp)
    A, L, G, S, P = np.exp(truth)
    K = o.kernel_matrix(truth, t, yerr)

    def k(a, b):
        d = a[:, None] - b[None, :]
        out = A * np.exp(-0.5 * d * d / L) * np.exp(-G * np.sin(np.pi * d / P) ** 2)
        return out + S * (d * d / L < np.finfo(float).eps)            # WhiteKernel at coincident inputs
    ks = k(tp, t)
    ref_mu = ks @ np.linalg.solve(K, y)
    ref_var = (A + S) - np.einsum('ij,ji->i', ks, np.linalg.solve(K, ks.T))
    scale = np.sqrt(A + S)
    assert np.all(np.abs(mu - ref_mu) <= 1e-6 * scale)
    assert np.all(np.abs(var - ref_var) <= 1e-6 * (A + S))
    assert var[2] < 0.1 * (A + S) and var[0] > var[1]                 # on a data point: small; outside the data: larger
