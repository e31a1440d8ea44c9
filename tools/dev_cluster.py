import sys, numpy as np
sys.path.insert(0, '.')
from oracle import gp_oracle as o
from gprotation_b200 import GPEngine
from gprotation_b200.synth import synthetic_lightcurve
eng = GPEngine(0); tf,_ = eng.measure_dmma_peak()
for N, Bs in ((8192, (32,)), (2048, (32, 64)), (1000, (32,))):
    eng.unregister_all()
    t, y, yerr, _ = synthetic_lightcurve(1, N, kind="harmonic")
    lc = eng.register_lc(t, y, yerr)
    F = o.flops_per_eval(N)
    for B in Bs:
        th = o.sample_from_prior(B, seed=9)
        base = None
        for cl in (1, 2, 4, 8):
            if N == 8192 and cl == 1: continue
            eng.set_cluster(cl)
            eng.lnlike_batch(th, lc_id=lc)
            ms = []
            for _ in range(2):
                got = eng.lnlike_batch(th, lc_id=lc); ms.append(eng.last_kernel_ms())
            m = min(ms)
            print(f"N={N} B={B} cl={cl}: {m:.2f} ms {B*F/m/1e9:.2f} TFLOP/s = {B*F/m/1e9/tf*100:.1f}% of peak", flush=True)
            if base is None: base = got
            assert np.array_equal(got, base), (cl, got, base)
eng.set_cluster(0)
