import sys
import numpy as np
sys.path.insert(0, '.')
from oracle import gp_oracle as o
from gprotation_b200 import GPEngine
from gprotation_b200.synth import synthetic_lightcurve
eng = GPEngine(0)
tf, _ = eng.measure_dmma_peak()
def timeit(th, lc, mode):
    eng.set_cluster(mode)
    eng.lnlike_batch(th, lc_id=lc)
    ms = []
    for _ in range(3):
        eng.lnlike_batch(th, lc_id=lc); ms.append(eng.last_kernel_ms())
    return min(ms)
for N, B in ((64, 8192), (128, 8192), (200, 4096), (300, 4096), (450, 4096), (600, 2048), (800, 2048), (1000, 1024), (1300, 1024), (1600, 1024), (2048, 1024)):
    eng.unregister_all()
    t, y, yerr, _ = synthetic_lightcurve(1, N, kind="harmonic")
    lc = eng.register_lc(t, y, yerr)
    th = o.sample_from_prior(B, seed=7)
    a, b = timeit(th, lc, 1), timeit(th, lc, 64)
    print(f"N={N:5d} B={B}: one/SM {a:8.3f} ms   two/SM {b:8.3f} ms   ratio {a/b:.3f}   ({B*o.flops_per_eval(N)/b/1e9/tf*100:.1f}% of peak)", flush=True)
eng.unregister_all()
t, y, yerr, _ = synthetic_lightcurve(3, 2000, kind="harmonic")
lc = eng.register_lc(t, y, yerr, chunks=o.split_chunks(2000, 300))
th = o.sample_from_prior(500, seed=5)
a, b = timeit(th, lc, 1), timeit(th, lc, 64)
print(f"500 walkers x 7 chunks of 286: one/SM {a:.3f} ms  two/SM {b:.3f} ms  ratio {a/b:.3f}")
eng.set_cluster(0)
