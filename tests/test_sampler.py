"""Host-side sampler shim (gprotation_b200/sampler.py, fit.fit_emcee3, lc front end, CLI): CPU tests with an analytic model.

emcee3 is not installable offline, so these tests pin the restated algorithms to what they must do: leave a known
target invariant, evaluate each half-ensemble with ONE pool.map call, store / resume chains, and follow the
control flow of gprot/fit.py:53-152."""
import os

import numpy as np
import pytest

from gprotation_b200 import sampler as em
from gprotation_b200.fit import Emcee3Model, EnsemblePool, State, fit_emcee3, read_samples
from gprotation_b200.lc import AigrainLightCurve, LightCurve, sigma_clip
from gprotation_b200.synth import write_aigrain_dataset


class GaussModel(object):
    """Correlated 5-d Gaussian with the GPRotModel surface the sampler touches."""
    param_names = ('ln_A', 'ln_l', 'ln_G', 'ln_sigma', 'ln_period')
    acf_prior = False

    def __init__(self, seed=3, name='gauss'):
        rng = np.random.RandomState(seed)
        a = rng.randn(5, 5)
        self.C = a @ a.T + np.eye(5)
        self.Ci = np.linalg.inv(self.C)
        self.name = name
        self.ndim = 5
        self.gp_prior_mu = np.zeros(4)
        self.gp_prior_sigma = np.ones(4)
        self.lc = LightCurve(np.arange(10.0), np.zeros(10), np.ones(10), name=name)
        self.batch_calls = 0
        self.single_calls = 0

    def lnprior(self, th):
        return 0.0 if np.all(np.abs(th) < 40) else -np.inf

    def lnlike(self, th):
        self.single_calls += 1
        return float(-0.5 * th @ self.Ci @ th)

    def lnpost_batch(self, thetas, return_parts=False):
        self.batch_calls += 1
        thetas = np.atleast_2d(thetas)
        lp = np.where(np.all(np.abs(thetas) < 40, axis=1), 0.0, -np.inf)
        ll = -0.5 * np.einsum('ij,jk,ik->i', thetas, self.Ci, thetas)
        ll = np.where(np.isfinite(lp), ll, -np.inf)
        return (lp, ll) if return_parts else lp + ll

    def sample_from_prior(self, n, seed=None):
        return np.random.RandomState(seed).randn(n, 5)


def test_autocorr_time_of_ar1():
    rng = np.random.RandomState(0)
    phi, n = 0.9, 100000
    x = np.empty((n, 2))
    x[0] = 0
    e = rng.randn(n, 2)
    for i in range(1, n):
        x[i] = phi * x[i - 1] + e[i]
    tau = em.integrated_time(x, c=5)
    assert np.all(np.abs(tau - 19.0) < 2.5)
    with pytest.raises(em.AutocorrError):
        em.integrated_time(x[:15], c=1)


@pytest.mark.parametrize("move", [em.StretchMove(), em.DEMove(1.0), em.DESnookerMove(), em.KDEMove(), "mixed"])
def test_moves_leave_gaussian_invariant(move):
    g = GaussModel()
    rs = np.random.RandomState(11)
    if move == "mixed":
        move = [(em.KDEMove(), 0.4), (em.DEMove(1.0), 0.4), (em.DESnookerMove(), 0.2)]     # gprot/fit.py:82-87
    ens = em.Ensemble(em.Model(g), rs.multivariate_normal(np.zeros(5), g.C, size=64), pool=EnsemblePool(), random=rs)
    sam = em.Sampler(move)
    sam.run(ens, 600)
    c = sam.get_coords(flat=True, discard=100)
    assert c.shape == (500 * 64, 5)
    assert np.abs(c.mean(0)).max() < 0.35 * np.sqrt(np.diag(g.C)).max()
    assert np.abs(np.cov(c.T) - g.C).max() < 0.3 * np.abs(g.C).max()
    assert 0.02 < sam.acceptance_fraction.mean() < 0.98
    assert g.single_calls == 0                      # never one walker at a time
    assert g.batch_calls == 1 + 2 * 600             # initial ensemble + one launch per half-ensemble per step


@pytest.mark.parametrize("move", ["stretch", "de", "snooker", "kde", "mixed"])
def test_native_chain_leaves_gaussian_invariant(move):
    """Sampler.run(native=True): the same chain driven by csrc/gprot_chain.cpp through gprot_chain_run_fn (here with a
    callback into the Gaussian; on the GPU gprot_chain_run calls gprot_lnpost_batch itself).  Same invariance bar as the
    Python-driven moves, same launch accounting, stored log-probabilities consistent with the stored coordinates."""
    g = GaussModel()
    rs = np.random.RandomState(11)
    moves = {"stretch": em.StretchMove(), "de": em.DEMove(1.0), "snooker": em.DESnookerMove(), "kde": em.KDEMove(),
             "mixed": [(em.KDEMove(), 0.4), (em.DEMove(1.0), 0.4), (em.DESnookerMove(), 0.2)]}[move]
    ens = em.Ensemble(Emcee3Model(g), rs.multivariate_normal(np.zeros(5), g.C, size=64), pool=EnsemblePool(), random=rs)
    sam = em.Sampler(moves)
    sam.run(ens, 600, native=True)
    c = sam.get_coords(flat=True, discard=100)
    assert c.shape == (500 * 64, 5)
    assert np.abs(c.mean(0)).max() < 0.35 * np.sqrt(np.diag(g.C)).max()
    assert np.abs(np.cov(c.T) - g.C).max() < 0.3 * np.abs(g.C).max()
    assert 0.02 < sam.acceptance_fraction.mean() < 0.98
    assert g.single_calls == 0
    assert g.batch_calls == 1 + 2 * 600             # initial ensemble + one launch per half-ensemble per step
    last = sam.get_coords()[-1]
    assert np.array_equal(last, ens.coords)
    assert np.allclose(sam.get_log_probability()[-1], g.lnpost_batch(last), rtol=0, atol=1e-12)
    # thinning and a second run appended to the same store
    sam.run(ens, 40, native=True, thin=4)
    assert sam.get_coords().shape == (610, 64, 5) and np.array_equal(sam.get_coords()[-1], ens.coords)


def test_native_chain_refuses_what_it_cannot_mirror():
    g = GaussModel()
    rs = np.random.RandomState(3)
    few = em.Ensemble(Emcee3Model(g), rs.randn(8, 5), pool=EnsemblePool(), random=rs)
    with pytest.raises(RuntimeError):                # red-blue moves need nwalkers >= 2 ndim (as the Python path)
        em.Sampler(em.DEMove(1.0)).run(few, 5, native=True)
    ens = em.Ensemble(Emcee3Model(g), rs.randn(32, 5), pool=EnsemblePool(), random=rs)
    with pytest.raises(ValueError):                  # a move mix without a native counterpart
        em.Sampler(em.StretchMove(randomize_split=True)).run(ens, 5, native=True)
    # an exception raised by the model inside the callback surfaces after the call, it is not swallowed

    class Boom(GaussModel):
        def lnpost_batch(self, thetas, return_parts=False):
            if self.batch_calls >= 3:
                raise KeyError("boom")
            return GaussModel.lnpost_batch(self, thetas, return_parts)

    b = Boom()
    with pytest.raises(KeyError):
        em.Sampler(em.DEMove(1.0)).run(em.Ensemble(Emcee3Model(b), rs.randn(32, 5), pool=EnsemblePool(), random=rs), 20, native=True)
    # proposals outside the support come back with lnprior = -inf and are never accepted
    start = rs.uniform(38.0, 39.9, size=(32, 5))
    edge = em.Ensemble(Emcee3Model(g), start, pool=EnsemblePool(), random=rs)
    sam = em.Sampler(em.DEMove(1.0))
    sam.run(edge, 50, native=True)
    assert np.all(np.abs(sam.get_coords()) < 40) and np.all(np.isfinite(sam.get_log_probability()))


def test_ensemble_rejects_zero_probability_start_and_out_of_support_moves():
    g = GaussModel()
    bad = np.zeros((16, 5))
    bad[3, 0] = 100.0
    with pytest.raises(ValueError):
        em.Ensemble(em.Model(g), bad, pool=EnsemblePool())
    # walkers hugging the support boundary: proposals outside get log_probability -inf and are never accepted
    rs = np.random.RandomState(2)
    start = rs.uniform(38.0, 39.9, size=(32, 5))
    ens = em.Ensemble(em.Model(g), start, pool=EnsemblePool(), random=rs)
    sam = em.Sampler(em.DEMove(1.0))
    sam.run(ens, 50)
    assert np.all(np.abs(sam.get_coords()) < 40)
    assert np.all(np.isfinite(sam.get_log_probability()))


def test_emcee3model_single_and_batched_agree():
    g = GaussModel()
    m = Emcee3Model(g)
    th = np.random.RandomState(5).randn(7, 5)
    th[2, 1] = 55.0
    single = [m(State(t)) for t in th]
    batched = EnsemblePool().map(m, [State(t) for t in th])
    for a, b in zip(single, batched):
        assert a.log_prior == b.log_prior
        assert a.log_likelihood == pytest.approx(b.log_likelihood, rel=1e-14) or (a.log_likelihood == b.log_likelihood)
    assert single[2].log_probability == -np.inf and batched[2].log_probability == -np.inf
    # the emcee3 default pool also batches an Emcee3Model
    assert [s.log_likelihood for s in em.DefaultPool().map(m, [State(t) for t in th])] == [s.log_likelihood for s in batched]


def test_backend_persistence_and_resume(tmp_path):
    g = GaussModel()
    f = str(tmp_path / "chains" / "gauss.h5")
    be = em.HDFBackend(f)
    assert isinstance(be, em.NPZBackend) and be.niter == 0
    with pytest.raises(AttributeError):
        be.current_coords
    rs = np.random.RandomState(1)
    ens = em.Ensemble(em.Model(g), rs.randn(20, 5), pool=EnsemblePool(), random=rs)
    sam = em.Sampler(em.StretchMove(), backend=be)
    sam.run(ens, 30)
    assert be.niter == 30 and os.path.exists(be.filename)
    be2 = em.HDFBackend(f)
    assert be2.niter == 30
    assert np.array_equal(be2.current_coords, ens.coords)
    sam2 = em.Sampler(em.StretchMove(), backend=be2)
    sam2.run(em.Ensemble(em.Model(g), be2.current_coords, pool=EnsemblePool(), random=rs), 10)
    assert be2.niter == 40 and sam2.get_coords().shape == (40, 20, 5)
    sam2.reset()
    assert be2.niter == 0 and not os.path.exists(be2.filename)


def test_fit_emcee3_control_flow(tmp_path):
    g = GaussModel(name='star7')
    df = fit_emcee3(g, nwalkers=40, nsamples=500, targetn=6, iter_chunksize=100, maxiter=30, nburn=2,
                    sample_directory=str(tmp_path / 'mcmc_chains'), resultsdir=str(tmp_path / 'results'), seed=4)
    assert list(df.columns) == list(g.param_names) and len(df) == 500
    assert np.abs(np.cov(df.values.T) - g.C).max() < 0.45 * np.abs(g.C).max()
    out = [p for p in os.listdir(str(tmp_path / 'results')) if p.startswith('star7.')]
    assert len(out) == 1
    f = os.path.join(str(tmp_path / 'results'), out[0])
    assert np.allclose(read_samples(f, 'samples').values, df.values)
    assert list(read_samples(f, 'gp_prior').columns) == ['mu', 'sigma']
    assert set(read_samples(f, 'lc').columns) == {'x', 'y', 'yerr'}
    # second call resumes from the stored chain: converged already, so no new steps are taken
    calls = g.batch_calls
    fit_emcee3(g, nwalkers=40, nsamples=100, targetn=6, iter_chunksize=100, maxiter=30, nburn=2,
               sample_directory=str(tmp_path / 'mcmc_chains'), resultsdir=str(tmp_path / 'results'), seed=5)
    assert g.batch_calls == calls + 1               # only the initial ensemble evaluation


def test_results_file_is_the_reference_hdf_layout_where_pytables_exists(tmp_path):
    """gprot/fit.py:24-41 writes results/{name}.h5 with pandas-HDF keys samples / lc / gp_prior (/ period_prior); downstream
    scripts read them with pd.read_hdf (gprot/summary.py:38-56).  Runs wherever PyTables is installed (not in this image:
    there the same tables go to .npz, covered by test_fit_emcee3_control_flow)."""
    pytest.importorskip("tables")
    import pandas as pd

    from gprotation_b200.fit import write_samples

    g = GaussModel(name='star9')
    g.acf_prior = True
    g.period_mixture = np.array([[0.7, 2.0, 0.1], [0.3, 2.7, 0.2]])
    df = pd.DataFrame(np.random.RandomState(3).normal(size=(50, 5)), columns=list(g.param_names))
    f = write_samples(g, df, resultsdir=str(tmp_path / 'results'))
    assert f.endswith('star9.h5')
    assert np.allclose(pd.read_hdf(f, 'samples').values, df.values) and list(pd.read_hdf(f, 'samples').columns) == list(g.param_names)
    assert list(pd.read_hdf(f, 'gp_prior').columns) == ['mu', 'sigma']
    assert set(pd.read_hdf(f, 'lc').columns) == {'x', 'y', 'yerr'}
    assert list(pd.read_hdf(f, 'period_prior').columns) == ['w', 'mu', 'sigma']


def test_lightcurve_front_end(tmp_path):
    d = write_aigrain_dataset(str(tmp_path / 'aigrain'), 2, baseline=300.0)
    lc = AigrainLightCurve(1, directory=d, sub=10, chunksize=300)
    n_full = len(np.arange(0.0, 300.0, 0.02043365))
    assert lc.name == '1' and len(lc.x_full) == n_full
    assert len(lc.x) <= n_full // 10 and len(lc.x) > 0.98 * (n_full // 10)     # 5-sigma clip drops at most a few
    assert np.all(np.diff(lc.x) > 0) and np.all(lc.yerr == 1e-5)
    assert abs(np.median(lc.y_full)) < 0.1                                     # y - 1
    assert max(len(c) for c in lc.x_list) <= 300 and sum(len(c) for c in lc.x_list) == len(lc.x)
    # same star, same seed -> same subsample (gprot/aigrain.py:101-104)
    assert np.array_equal(AigrainLightCurve(1, directory=d, sub=10).x, lc.x)
    assert lc.sim_params.P_MIN > 0
    lc.make_best_chunks(ndays=[200, 50], npoints=600, chunksize=300)
    assert len(lc.x_list) >= 3 and max(len(c) for c in lc.x_list) <= 300
    x, y, e = sigma_clip(np.arange(100.0), np.r_[np.zeros(99), 50.0], np.ones(100), 5)
    assert len(x) == 99
    nochunks = AigrainLightCurve(0, directory=d, sub=20, chunksize=None)
    assert nochunks.x_list is None


def test_cli_flags_match_reference():
    """Every flag of scripts/gprot-fit:119-191 parses with the same defaults."""
    from gprotation_b200.cli import build_parser

    a = vars(build_parser().parse_args(['3', '7', '--aigrain']))
    ref_defaults = dict(stars=[3, 7], aigrain=True, kepler=False, n_cores=1, mpi=False, verbose=False, clever=False, npoints=600,
                        acf_prior=False, sap=False, filter=False, bestchunk=None, tag=None, quarters=None, sampler='emcee3',
                        resultsdir='results', ndays=None, subsample=30, chunksize=300, nwalkers=500, iter_chunksize=50,
                        targetn=6, nburn=2, maxiter=50, overwrite=False, daterange=None, mixedmoves=True, nlive=1000,
                        test=False, offline=False, altmodel=False, nochunks=False)
    for k, v in ref_defaults.items():
        assert a[k] == v, k
    b = vars(build_parser().parse_args('5 --aigrain --nochunks --nwalkers 64 -n 200 --subsample 10 -o -v --tag t --bestchunk 50 100'.split()))
    assert b['nochunks'] and b['nwalkers'] == 64 and b['ndays'] == 200 and b['overwrite'] and b['bestchunk'] == [50, 100]
