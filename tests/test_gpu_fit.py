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


def test_native_chain_on_the_gpu_model(engine, tmp_path):
    """Sampler.run(native=True) on a GPRotModel: gprot_chain_run drives gprot_lnpost_batch from the host side of the ABI.
    Two launches per step, every stored log-likelihood / log-prior is the oracle's at the stored coordinates, the ensemble
    moves, and fit_emcee3(native=True) produces a posterior sample like the Python-driven fit."""
    engine.unregister_all()
    mod, (t, y, yerr, truth) = _model(engine)
    rs = np.random.RandomState(3)
    ens = em.Ensemble(Emcee3Model(mod), mod.sample_from_prior(32, seed=5), pool=EnsemblePool(), random=rs)
    start = ens.coords.copy()
    l0 = engine.launch_count()
    sam = em.Sampler([(em.KDEMove(), 0.4), (em.DEMove(1.0), 0.4), (em.DESnookerMove(), 0.2)])
    sam.run(ens, 15, native=True)
    assert engine.launch_count() - l0 == 15 * 2 * 2
    chunks = o.split_chunks(len(t), 150)
    coords = sam.get_coords()
    ll, lp = sam.backend.log_likelihood, sam.backend.log_prior
    for it in (0, 7, 14):
        for w in (0, 13, 31):
            ref = o.lnlike(coords[it, w], t, y, yerr, chunks)
            assert abs(ll[it, w] - ref) <= 1e-9 * abs(ref)
            assert lp[it, w] == pytest.approx(o.lnprior(coords[it, w]), rel=1e-12)
    assert np.array_equal(coords[-1], ens.coords) and not np.array_equal(coords[-1], start)
    assert 0.0 < sam.acceptance_fraction.mean() < 1.0
    df = fit_emcee3(mod, nwalkers=24, nsamples=100, targetn=2, iter_chunksize=20, maxiter=3, nburn=1, overwrite=True,
                    sample_directory=str(tmp_path / "mcmc_chains"), resultsdir=str(tmp_path / "results"), seed=2, native=True)
    assert list(df.columns) == list(mod.param_names) and len(df) > 0 and np.all(np.isfinite(df.values))


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


def test_sampler_steps_through_a_group_engine(engine):
    """GPRotModel(engine=GPGroup([...])): Ensemble.propose -> lnpost_batch -> gprot_group_lnpost_batch.  On a one-GPU box
    the group holds two contexts on device 0; the chain must be the one a single engine produces, value for value."""
    import torch

    from gprotation_b200 import GPGroup

    ndev = torch.cuda.device_count()
    grp = GPGroup([0, 1 % ndev])
    try:
        engine.unregister_all()
        chains = []
        for eng in (engine, grp):
            mod, (t, y, yerr, truth) = _model(eng, n=360, chunksize=120, seed=23)
            rs = np.random.RandomState(4)
            ens = em.Ensemble(Emcee3Model(mod), mod.sample_from_prior(20, seed=6), pool=EnsemblePool(), random=rs)
            sam = em.Sampler([(em.KDEMove(), 0.4), (em.DEMove(1.0), 0.4), (em.DESnookerMove(), 0.2)])
            sam.run(ens, 6)
            chains.append((sam.get_coords().copy(), sam.backend.log_likelihood.copy(), sam.backend.log_prior.copy()))
        for a, b in zip(chains[0], chains[1]):
            assert np.array_equal(a, b)
        # and the posterior parts agree with the oracle, rows outside the prior masked
        th = mod.sample_from_prior(7, seed=8)
        th[3, 0] = 1.0
        lp, ll = mod.lnpost_batch(th, return_parts=True)
        assert lp[3] == -np.inf and ll[3] == -np.inf
        chunks = o.split_chunks(len(t), 120)
        for i in (0, 5):
            ref = o.lnlike(th[i], t, y, yerr, chunks)
            assert abs(ll[i] - ref) <= 1e-9 * abs(ref)
            assert lp[i] == pytest.approx(o.lnprior(th[i]), rel=1e-12)
    finally:
        grp.close()


def test_gp_handle_is_the_george_surface_the_reference_uses(engine):
    """GPRotModel.gp(theta, x, yerr).lnlikelihood(y, quiet=True) (gprot/model.py:336-356) and .predict(y, t)."""
    engine.unregister_all()
    t, y, yerr, truth = synthetic_lightcurve(33, 240, baseline=100.0, kind="gp")
    mod = GPRotModel(LightCurve(t, y, yerr, name="g", chunksize=None), engine=engine)
    gp = mod.gp(truth)
    ref = o.lnlike_function(truth, t, y, yerr)
    assert abs(gp.lnlikelihood(y, quiet=True) - ref) <= 1e-9 * abs(ref)
    sub = mod.gp(truth, x=t[:100], yerr=yerr[:100])
    ref = o.lnlike_function(truth, t[:100], y[:100], yerr[:100])
    assert abs(sub.lnlikelihood(y[:100]) - ref) <= 1e-9 * abs(ref)
    bad = truth.copy()
    bad[0] = 800.0
    assert mod.gp(bad).lnlikelihood(y, quiet=True) == -np.inf
    with pytest.raises(np.linalg.LinAlgError):
        mod.gp(bad).lnlikelihood(y)
    mu, var = gp.predict(y, np.array([t[10], t[-1] + 2.0]))
    mu2, var2 = mod.predict(truth, np.array([t[10], t[-1] + 2.0]))
    assert np.allclose(mu, mu2, rtol=1e-9, atol=0) and np.allclose(var, var2, rtol=1e-7, atol=0)


def test_registration_hygiene(engine):
    """Per-light-curve unregister, id validity across unregister_all (generation), the ad-hoc LRU of lnlike_function, the
    re-registration of a light curve that changed, and the cached problem list with an explicit lc id."""
    engine.unregister_all()
    t, y, yerr, truth = synthetic_lightcurve(34, 200, baseline=100.0, kind="harmonic")
    a = engine.register_lc(t, y, yerr)
    b = engine.register_lc(t[:150], y[:150], yerr[:150])
    th = o.sample_from_prior(5, seed=1)
    want_b = engine.lnlike_batch(th, lc_id=b)
    engine.unregister_lc(a)
    assert np.array_equal(engine.lnlike_batch(th, lc_id=b), want_b)          # b keeps its id and its data
    with pytest.raises(Exception):
        engine.lnlike_batch(th, lc_id=a)
    with pytest.raises(Exception):
        engine.unregister_lc(a)
    c = engine.register_lc(t, 2 * y, yerr)
    assert c not in (a, b)
    # repeated identical calls hit the cached device list: one kernel + one reduction each, same bits
    l0 = engine.launch_count()
    for _ in range(3):
        assert np.array_equal(engine.lnlike_batch(th, lc_id=b), want_b)
    assert engine.launch_count() - l0 == 6
    # a model notices that every id became void
    mod = GPRotModel(LightCurve(t, y, yerr, name="h", chunksize=None), engine=engine)
    first = mod.lnlike(truth)
    g0 = engine.generation
    engine.unregister_all()
    assert engine.generation == g0 + 1
    other = engine.register_lc(t[:50], y[:50], yerr[:50])                    # takes id 0: the model must not use it
    assert mod.lnlike(truth) == first
    assert mod._lc_id != other
    # a light curve that changes is registered again, the stale copy dropped
    mod.lc.y = 3.0 * np.asarray(mod.lc.y)
    ref = o.lnlike_function(truth, t, 3.0 * y, yerr)
    assert abs(mod.lnlike(truth) - ref) <= 1e-9 * abs(ref)
    # ad-hoc arrays: bounded
    mod._adhoc_max = 4
    for k in range(9):
        sl = slice(k, k + 60)
        got = mod.lnlike_function(truth, t[sl], y[sl], yerr[sl])
        ref = o.lnlike_function(truth, t[sl], y[sl], yerr[sl])
        assert abs(got - ref) <= 1e-9 * abs(ref)
    assert len(mod._adhoc) == 4
    assert mod.lnlike_function(truth, t[8:68], y[8:68], yerr[8:68]) == got          # a hit, same bits
    # many registrations with interleaved unregisters: the arena is compacted, ids keep their meaning
    keep = engine.register_lc(t, y, yerr)
    base = engine.lnlike_batch(th, lc_id=keep)
    rng = np.random.default_rng(0)
    for k in range(40):
        tt = np.sort(rng.uniform(0, 100, 4000))
        tmp = engine.register_lc(tt, rng.standard_normal(4000) * 1e-3, np.full(4000, 1e-4))
        engine.unregister_lc(tmp)
    assert np.array_equal(engine.lnlike_batch(th, lc_id=keep), base)


@pytest.mark.parametrize("cl", [2, 4])
def test_cluster_kernels_with_invalid_problems_mixed_in(engine, cl):
    """A chunk with a NaN time stamp (mingap2 < 0) and rows with non-finite theta take the early exit of the cluster
    kernels; mixed with valid problems in one launch they must neither hang nor disturb the valid rows."""
    engine.unregister_all()
    t, y, yerr, truth = synthetic_lightcurve(35, 700, baseline=300.0, kind="harmonic")
    tb = t.copy()
    tb[300] = np.nan
    good = engine.register_lc(t, y, yerr)
    badlc = engine.register_lc(tb, y, yerr, chunks=o.split_chunks(700, 250))
    th = o.sample_from_prior(24, seed=2)
    th[5, 0] = np.nan
    th[11, 4] = np.inf
    rows = np.array([badlc if i % 3 == 1 else good for i in range(24)], dtype=np.int32)
    try:
        engine.set_cluster(1)
        ref = engine.lnlike_batch(th, lc_id=rows)
        engine.set_cluster(cl)
        for _ in range(5):
            got = engine.lnlike_batch(th, lc_id=rows)
            assert engine.last_cluster() == cl
            assert np.array_equal(got, ref)
    finally:
        engine.set_cluster(0)
    assert np.all(ref[rows == badlc] == -np.inf) and ref[5] == -np.inf and ref[11] == -np.inf
    ok = [i for i in range(24) if rows[i] == good and i not in (5, 11)]

    def cond(x):
        ev = np.linalg.eigvalsh(o.kernel_matrix(x, t, yerr))
        return ev[-1] / ev[0]
    ok = [i for i in ok if cond(th[i]) <= 1e7][:4]               # the flat bound of tests/test_gpu_parity.py
    assert len(ok) >= 2
    want = np.array([o.lnlike_function(th[i], t, y, yerr) for i in ok])
    assert np.all(np.abs(ref[ok] - want) <= 1e-9 * np.abs(want))
