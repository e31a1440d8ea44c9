"""Two-problems-per-SM kernel (gp_kernel64.cuh) against the one-problem-per-SM kernel and the oracle; timing."""
import sys
import numpy as np
sys.path.insert(0, '.')
from oracle import gp_oracle as o
from gprotation_b200 import GPEngine
from gprotation_b200.synth import synthetic_lightcurve

eng = GPEngine(0)
tf, mhz = eng.measure_dmma_peak()
sizes = [int(a) for a in sys.argv[1:]] or [1, 9, 63, 64, 65, 127, 128, 129, 191, 192, 200, 286, 320, 500, 1000]
for n in sizes:
    eng.unregister_all()
    t, y, yerr, truth = synthetic_lightcurve(3000 + n, n, baseline=max(60.0, 0.6 * n), kind="harmonic")
    lc = eng.register_lc(t, y, yerr)
    th = o.sample_from_prior(6, seed=n)
    th[0] = truth
    ref = np.array([o.lnlike_function(x, t, y, yerr) for x in th])
    eng.set_cluster(1)
    a = eng.lnlike_batch(th, lc_id=lc)
    for exact in (False, True):
        eng.set_cluster(64)
        b = eng.lnlike_batch(th, lc_id=lc, exact=exact)
        assert eng.last_cluster() == 64
        rel = np.abs(b - ref) / np.abs(ref)
        print(f"n={n:5d} exact={int(exact)} k64 vs oracle {rel.max():.2e}   k64 vs k128 {np.max(np.abs(a - b) / np.abs(a)):.2e}", flush=True)
for N, B in ((2048, 1024), (2048, 296), (1000, 1024), (8192, 296)):
    if len(sys.argv) > 1:
        break
    eng.unregister_all()
    t, y, yerr, _ = synthetic_lightcurve(1, N, kind="harmonic")
    lc = eng.register_lc(t, y, yerr)
    th = o.sample_from_prior(B, seed=7)
    F = o.flops_per_eval(N)
    for mode in (1, 64):
        eng.set_cluster(mode)
        eng.lnlike_batch(th, lc_id=lc)
        ms = min(eng.lnlike_batch(th, lc_id=lc) is not None and eng.last_kernel_ms() for _ in range(2))
        print(f"N={N} B={B} mode={mode}: {ms:.2f} ms  {B/ms*1e3:.0f} evals/s  {B*F/ms/1e9/tf*100:.1f}% of DMMA peak", flush=True)
eng.set_cluster(0)
