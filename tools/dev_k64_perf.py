"""One shape, one kernel choice (for ncu): python tools/dev_k64_perf.py N B mode   (mode 0 auto, 1, 2, 4, 64)"""
import sys
sys.path.insert(0, '.')
from oracle import gp_oracle as o
from gprotation_b200 import GPEngine
from gprotation_b200.synth import synthetic_lightcurve
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
eng = GPEngine(0)
t, y, yerr, _ = synthetic_lightcurve(1, N, kind="harmonic")
lc = eng.register_lc(t, y, yerr)
th = o.sample_from_prior(B, seed=7)
eng.set_cluster(mode)
for _ in range(2):
    eng.lnlike_batch(th, lc_id=lc)
    print(eng.last_kernel_ms(), eng.last_cluster())
