"""Deterministic synthetic Kepler-cadence light curves.

Per star ``s`` (seed = s, as gprot/aigrain.py:101-104 seeds the subsample with the star index):
a 30-minute-cadence grid over ``baseline`` days (0.02043365 d, gprotation/simulate_gps.py:28),
N distinct indices drawn without replacement and sorted (gprot/lc.py:411-416), a quasi-periodic
signal, 1e-5 white noise and a constant yerr = 1e-5 (gprot/aigrain.py:26).  Truth hyper-parameters
follow the simulator distribution of gprotation/simulate_gps.py:14-22 clipped to the model bounds.

``kind='gp'`` draws y from the GP itself (dense NumPy Cholesky on the host: data generation, not the
measured path); ``kind='harmonic'`` is an O(N) quasi-periodic signal (a spot-like sum of damped
harmonics) for sweeps over many stars, where a 2048^3 host factorisation per star would dominate.
"""
from __future__ import annotations

import numpy as np

KEPLER_CADENCE = 0.02043365  # days


def truth_theta(rng):
    th = np.empty(5)
    th[0] = np.clip(rng.normal(-13.0, 2.7), -19.0, -1.0)
    th[4] = rng.uniform(np.log(0.5), np.log(100.0))
    th[1] = np.clip(rng.normal(6.2, 1.5), max(2.0, th[4]) + 0.1, 19.0)
    th[2] = np.clip(rng.normal(-1.4, 1.5), -9.0, 2.9)
    th[3] = -17.0
    return th


def _cov(theta, t, yerr):
    A, L, G, S, P = np.exp(theta)
    d = t[:, None] - t[None, :]
    s = np.sin(np.pi * d / P)
    K = A * np.exp(-0.5 * d * d / L - G * s * s)
    K[np.diag_indices_from(K)] += 2.0 * S + yerr ** 2
    return K


def synthetic_lightcurve(seed, n, baseline=1400.0, cadence=KEPLER_CADENCE, theta_true=None, kind="gp"):
    """Returns (t, y, yerr, theta_true), all float64, t sorted and distinct."""
    rng = np.random.default_rng(seed)
    grid = np.arange(0.0, baseline, cadence)
    if n > grid.size:
        raise ValueError("n exceeds the number of cadences in the baseline")
    t = grid[np.sort(rng.choice(grid.size, n, replace=False))]
    yerr = np.full(n, 1e-5)
    if theta_true is None:
        theta_true = truth_theta(rng)
    theta_true = np.asarray(theta_true, dtype=np.float64)
    if kind == "gp":
        K = _cov(theta_true, t, np.zeros(n)) + 1e-12 * np.exp(theta_true[0]) * np.eye(n)
        y = np.linalg.cholesky(K) @ rng.standard_normal(n)
    elif kind == "harmonic":
        A, L, G, S, P = np.exp(theta_true)
        y = np.zeros(n)
        for h in (1, 2):
            nseg = max(2, int(np.ceil(baseline / np.sqrt(L))))
            knots = np.linspace(0.0, baseline, nseg + 1)
            amp = np.interp(t, knots, rng.standard_normal(nseg + 1))
            ph = np.interp(t, knots, np.cumsum(0.3 * rng.standard_normal(nseg + 1)))
            y += amp * np.sin(2 * np.pi * h * t / P + ph) / h
        y *= np.sqrt(A) / max(np.std(y), 1e-30)
    else:
        raise ValueError("kind must be 'gp' or 'harmonic'")
    y = y + 1e-5 * rng.standard_normal(n)
    return t, y, yerr, theta_true


def write_aigrain_dataset(directory, nstars, baseline=1400.0, cadence=KEPLER_CADENCE, seed=0):
    """Write a synthetic data set in the on-disk layout of the Aigrain et al. (2015) simulations that
    gprot/aigrain.py reads: ``final/lightcurve_%04d.txt`` (time, 1 + relative flux) and ``par/final_table.txt``
    (whitespace table with P_MIN / P_MAX columns).  O(N) harmonic signals: this is input plumbing, not the GP."""
    import os

    os.makedirs(os.path.join(directory, 'final'), exist_ok=True)
    os.makedirs(os.path.join(directory, 'par'), exist_ok=True)
    grid = np.arange(0.0, baseline, cadence)
    rows = []
    for s in range(nstars):
        rng = np.random.default_rng(seed + s)
        th = truth_theta(rng)
        _, y, _, _ = synthetic_lightcurve(seed + s, grid.size, baseline=baseline, cadence=cadence, theta_true=th, kind='harmonic')
        np.savetxt(os.path.join(directory, 'final', 'lightcurve_%04d.txt' % s), np.column_stack([grid, 1.0 + y]), fmt='%.10f')
        P = float(np.exp(th[4]))
        rows.append((s, P, P))
    with open(os.path.join(directory, 'par', 'final_table.txt'), 'w') as f:
        f.write('N P_MIN P_MAX\n')
        for r in rows:
            f.write('%d %.6f %.6f\n' % r)
    return directory
