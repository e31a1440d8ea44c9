"""Small invocation for compute-sanitizer (memcheck / racecheck / synccheck): N in {129, 300, 1000} x {1, 2, 4 CTAs per
problem, two problems per SM}, chunked and unchunked, a handful of proposals, each checked against the oracle.
usage: compute-sanitizer --tool racecheck python tools/dev_sanitize.py [N ...]"""
import sys
import numpy as np
sys.path.insert(0, '.')
from oracle import gp_oracle as o
from gprotation_b200 import GPEngine
from gprotation_b200.synth import synthetic_lightcurve

sizes = [int(a) for a in sys.argv[1:]] or [129, 300, 1000]
eng = GPEngine(0)
worst = 0.0
for n in sizes:
    t, y, yerr, truth = synthetic_lightcurve(5 + n, n, baseline=max(100.0, 0.5 * n), kind="harmonic")
    chunks = o.split_chunks(n, max(64, n // 3))
    eng.unregister_all()
    a = eng.register_lc(t, y, yerr)
    b = eng.register_lc(t, y, yerr, chunks=chunks)
    th = o.sample_from_prior(3, seed=n)
    th[0] = truth
    ref = {a: np.array([o.lnlike(x, t, y, yerr) for x in th]), b: np.array([o.lnlike(x, t, y, yerr, chunks) for x in th])}
    for cl in (1, 2, 4, 64):
        eng.set_cluster(cl)
        for lc in (a, b):
            got = eng.lnlike_batch(th, lc_id=lc)
            rel = np.abs(got - ref[lc]) / np.abs(ref[lc])
            worst = max(worst, float(rel[0]))
            print("n=%d kernel=%d %s: rel err %s" % (n, eng.last_cluster(), "chunked" if lc == b else "single", rel), flush=True)
eng.set_cluster(0)
eng.close()
print("done; worst rel err at the generating hyper-parameters: %.2e" % worst)
