import sys, numpy as np
sys.path.insert(0, '.')
from oracle import gp_oracle as o
from gprotation_b200 import GPEngine
from gprotation_b200.synth import synthetic_lightcurve
eng = GPEngine(0)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tails = [1, 7, 8, 9, 16, 31, 32, 33, 40, 63, 64, 65, 95, 96, 97, 120, 127]
th = o.sample_from_prior(3, seed=5 + k)
for r in tails:
    n = 128 * k + r
    t, y, yerr, _ = synthetic_lightcurve(2000 + n, n, baseline=max(80.0, 0.5 * n), kind="harmonic")
    eng.unregister_all()
    lc = eng.register_lc(t, y, yerr)
    ref = np.array([o.lnlike_function(x, t, y, yerr) for x in th])
    for exact in (False, True):
        got = eng.lnlike_batch(th, lc_id=lc, exact=exact)
        rel = np.abs(got - ref) / np.abs(ref)
        ev = [np.linalg.eigvalsh(o.kernel_matrix(x, t, yerr)) for x in th]
        cond = np.array([e[-1] / e[0] for e in ev])
        flag = "  <-- FAIL" if np.any(rel > 1e-9 * np.maximum(1, cond / 1e7)) else ""
        print(n, r, int(exact), rel, cond, flag)
