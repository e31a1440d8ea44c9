"""lnprior truth table (gprot/model.py:132-163): oracle vs the model mirror vs the C ABI (CPU only)."""
import json
import os

import numpy as np
import pytest

from oracle import gp_oracle as o
from gprotation_b200.lc import LightCurve
from gprotation_b200.model import GPRotModel, KeplerGPRotModel

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def table():
    return json.load(open(os.path.join(HERE, "golden", "lnprior_golden.json")))


@pytest.fixture(scope="module")
def lc():
    t = np.linspace(0, 10, 50)
    return LightCurve(t, np.zeros(50), np.full(50, 1e-5), chunksize=None)


def _same(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert np.array_equal(np.isfinite(a), np.isfinite(b))
    m = np.isfinite(a)
    assert np.allclose(a[m], b[m], rtol=1e-13, atol=0)


def test_oracle_reproduces_table(table):
    th = table["theta"]
    _same([o.lnprior(r) for r in th], table["default"])
    _same([o.lnprior(r, o.KEPLER_BOUNDS, o.KEPLER_GP_PRIOR_MU, o.KEPLER_GP_PRIOR_SIGMA) for r in th], table["kepler"])
    _same([o.lnprior(r, pmax=3.0) for r in th], table["pmax_3"])
    _same([o.lnprior(r, mixture=table["mixture"]) for r in th], table["default_mixture"])


def test_support_rules(table):
    d = table["default"]
    assert np.isfinite(d[0]) and np.isfinite(d[1]) and np.isfinite(d[2])   # bounds are inclusive
    assert d[3] == -np.inf and d[4] == -np.inf
    assert d[5] == -np.inf and np.isfinite(d[6])                           # ln l >= ln P required
    assert d[7] == d[8] == d[9] == -np.inf
    assert table["pmax_3"][6] == -np.inf                                   # ln P = 3.5 > pmax


def test_model_scalar_and_batch_match_table(table, lc, built):
    th = np.array(table["theta"])
    mod = GPRotModel(lc)
    _same([mod.lnprior(r) for r in th], table["default"])
    _same(mod.lnprior_batch(th), table["default"])
    kep = KeplerGPRotModel(lc)
    _same([kep.lnprior(r) for r in th], table["kepler"])
    _same(kep.lnprior_batch(th), table["kepler"])
    _same(GPRotModel(lc, pmax=3.0).lnprior_batch(th), table["pmax_3"])
    mm = GPRotModel(lc, acf_prior=True, period_mixture=table["mixture"])
    _same([mm.lnprior(r) for r in th], table["default_mixture"])
    _same(mm.lnprior_batch(th), table["default_mixture"])


def test_acf_prior_without_mixture_is_explicit(lc):
    with pytest.raises(NotImplementedError):
        GPRotModel(lc, acf_prior=True).lnprior([-13.0, 7.2, -2.3, -17.0, 2.0])


def test_model_constants_match_reference():
    assert GPRotModel._default_bounds == ((-20., 0.), (2, 20.), (-10., 3.), (-20., 0.), (-0.69, 4.61))
    assert GPRotModel._default_gp_prior_mu == (-13, 7.2, -2.3, -17)
    assert GPRotModel._default_gp_prior_sigma == (5.7, 1.2, 1.4, 5)
    assert GPRotModel.param_names == ('ln_A', 'ln_l', 'ln_G', 'ln_sigma', 'ln_period')
    assert KeplerGPRotModel._default_bounds[1] == (2, 8.) and KeplerGPRotModel._default_bounds[2] == (0., 3.)
    assert KeplerGPRotModel._default_gp_prior_mu == (-13, 5.0, 1.9, -17)


def test_sample_from_prior_matches_oracle_stream(lc, built):
    mod = GPRotModel(lc)
    s = mod.sample_from_prior(64, seed=9)
    assert s.shape == (64, 5)
    assert np.array_equal(s, o.sample_from_prior(64, seed=9))
    assert np.all(np.isfinite(mod.lnprior_batch(s)))
    lo, hi = np.array(mod.bounds).T
    assert np.all(s >= lo) and np.all(s <= hi) and np.all(s[:, 1] >= s[:, 4])
    s3 = GPRotModel(lc, pmax=3.0).sample_from_prior(32, seed=1)
    assert np.all(s3[:, 4] <= 3.0)
