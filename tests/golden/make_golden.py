"""Regenerates tests/golden/lnlike_golden.npz and lnprior_golden.json.

The reference (gprot/model.py on george 0.2.x + emcee3) cannot be imported in the build container
(SURVEY.md section 8c), and it ships no golden values: PARITY UNPINNED at the george boundary.
These fixtures therefore freeze the *oracle's* answers (oracle/gp_oracle.py, NumPy/LAPACK) together with
two independent evaluations of the same formula -- 80-bit long double C (oracle/gp_oracle_ld.c) and, for the
smallest case, 40-digit mpmath -- so that a later change to the oracle, to NumPy/LAPACK, or to the platform
cannot move the target silently.   Run:  python tests/golden/make_golden.py
"""
import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import gp_oracle as o  # noqa: E402
from gprotation_b200.synth import synthetic_lightcurve  # noqa: E402

lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libgp_oracle.so"))
for f in (lib.gp_oracle_lnlike_f64, lib.gp_oracle_lnlike_ld):
    f.restype = ctypes.c_double
    f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]


def c_eval(fn, th, t, y, e):
    th, t, y, e = (np.ascontiguousarray(a, dtype=np.float64) for a in (th, t, y, e))
    return fn(th.ctypes.data, len(t), t.ctypes.data, y.ctypes.data, e.ctypes.data)


out = {}
cases = [("n24", 24, None, 101), ("n64", 64, None, 102), ("n150", 150, None, 103), ("n300c100", 300, 100, 104),
         ("n257", 257, None, 105)]
for name, n, chunksize, seed in cases:
    t, y, yerr, truth = synthetic_lightcurve(seed, n, baseline=200.0 if n < 100 else 1400.0)
    thetas = o.sample_from_prior(6, seed=seed)
    thetas[0] = truth                      # the generating hyper-parameters
    thetas[1] = [-19.9, 2.1, -9.9, -19.9, -0.6]   # a corner of the bounds box
    chunks = o.split_chunks(n, chunksize)
    ref = np.array([o.lnlike(th, t, y, yerr, chunks) for th in thetas])
    parts = chunks if chunks is not None else [np.arange(n)]
    cf64 = np.array([sum(c_eval(lib.gp_oracle_lnlike_f64, th, t[c], y[c], yerr[c]) for c in parts) for th in thetas])
    cld = np.array([sum(c_eval(lib.gp_oracle_lnlike_ld, th, t[c], y[c], yerr[c]) for c in parts) for th in thetas])
    conds = []
    for th in thetas:
        cmax = 0.0
        for c in parts:
            ev = np.linalg.eigvalsh(o.kernel_matrix(th, t[c], yerr[c]))
            cmax = max(cmax, ev[-1] / ev[0])
        conds.append(cmax)
    out[name + "_t"], out[name + "_y"], out[name + "_yerr"] = t, y, yerr
    out[name + "_theta"], out[name + "_lnlike"] = thetas, ref
    out[name + "_lnlike_c_f64"], out[name + "_lnlike_c_ld"] = cf64, cld
    out[name + "_cond"] = np.array(conds)
    out[name + "_chunksize"] = np.array(-1 if chunksize is None else chunksize)
    if n <= 24:
        out[name + "_lnlike_mpmath"] = np.array([o.lnlike_mpmath(th, t, y, yerr) for th in thetas])
    print(name, "max |numpy - long double| / |.| =", np.max(np.abs(ref - cld) / np.abs(cld)), "max cond %.2e" % max(conds))
np.savez_compressed(os.path.join(HERE, "lnlike_golden.npz"), **out)

# lnprior truth table (gprot/model.py:132-157 with the class constants of :36-45 and gprot/kepler.py:27-34)
rows = [
    [-13.0, 7.2, -2.3, -17.0, 2.0],      # prior means
    [-20.0, 2.0, -10.0, -20.0, -0.69],   # lower corner (on the bounds: allowed)
    [0.0, 20.0, 3.0, 0.0, 4.61],         # upper corner
    [-20.0001, 7.2, -2.3, -17.0, 2.0],   # below lnA bound
    [-13.0, 20.5, -2.3, -17.0, 2.0],     # above ln l bound
    [-13.0, 3.0, -2.3, -17.0, 3.5],      # ln l < ln P
    [-13.0, 3.5, -2.3, -17.0, 3.5],      # ln l == ln P: allowed
    [-13.0, 7.2, 3.01, -17.0, 2.0],
    [-13.0, 7.2, -2.3, 0.5, 2.0],
    [-13.0, 7.2, -2.3, -17.0, 4.7],
    [-5.0, 5.0, 1.9, -10.0, 1.0],
]
mix = [[0.6, 2.0, 0.2], [0.3, 2.0 - np.log(2), 0.2], [0.1, 2.0 + np.log(2), 0.2]]
table = {"theta": rows,
         "default": [o.lnprior(r) for r in rows],
         "kepler": [o.lnprior(r, o.KEPLER_BOUNDS, o.KEPLER_GP_PRIOR_MU, o.KEPLER_GP_PRIOR_SIGMA) for r in rows],
         "pmax_3": [o.lnprior(r, pmax=3.0) for r in rows],
         "mixture": mix,
         "default_mixture": [o.lnprior(r, mixture=mix) for r in rows]}
json.dump(table, open(os.path.join(HERE, "lnprior_golden.json"), "w"), indent=1)
print("wrote fixtures")
