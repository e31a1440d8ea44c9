import sys
import numpy as np
sys.path.insert(0, '.')
from oracle import gp_oracle as o
from gprotation_b200 import GPEngine
from gprotation_b200.synth import synthetic_lightcurve
eng = GPEngine(0)
for N, B in ((100, 100), (286, 100), (286, 148), (450, 100), (600, 100), (900, 100), (900, 32), (286, 32)):
    eng.unregister_all()
    t, y, yerr, _ = synthetic_lightcurve(1, N, kind="harmonic")
    lc = eng.register_lc(t, y, yerr)
    th = o.sample_from_prior(B, seed=7)
    res = []
    for mode in (0, 1, 64):
        eng.set_cluster(mode)
        eng.lnlike_batch(th, lc_id=lc)
        ms = []
        for _ in range(3):
            eng.lnlike_batch(th, lc_id=lc); ms.append(eng.last_kernel_ms())
        res.append((mode, eng.last_cluster(), min(ms)))
    print(N, B, " ".join("mode %d (used %d): %.3f ms" % r for r in res), flush=True)
