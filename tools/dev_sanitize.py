"""Small invocation for compute-sanitizer (memcheck / racecheck / synccheck): a chunked and an unchunked light curve,
cluster sizes 1 and 2, a handful of proposals."""
import sys
import numpy as np
sys.path.insert(0, '.')
from oracle import gp_oracle as o
from gprotation_b200 import GPEngine
from gprotation_b200.synth import synthetic_lightcurve

eng = GPEngine(0)
t, y, yerr, _ = synthetic_lightcurve(5, 420, baseline=300.0, kind="harmonic")
a = eng.register_lc(t, y, yerr)
b = eng.register_lc(t, y, yerr, chunks=o.split_chunks(420, 150))
th = o.sample_from_prior(3, seed=1)
for cl in (1, 2):
    eng.set_cluster(cl)
    for lc in (a, b):
        got = eng.lnlike_batch(th, lc_id=lc)
        ref = np.array([o.lnlike(x, t, y, yerr, None if lc == a else o.split_chunks(420, 150)) for x in th])
        print(cl, lc, np.abs(got - ref) / np.abs(ref))
eng.close()
