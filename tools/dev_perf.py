"""Development timing run on the GPU box: N=2048 batched lnlike, device time from CUDA events."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from oracle import gp_oracle as o
from gprotation_b200 import GPEngine
from tools.dev_check import make_lc  # noqa

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
Bs = [int(a) for a in sys.argv[2:]] or [148, 1024]
eng = GPEngine(0)
tf, mhz = eng.measure_dmma_peak()
print(f"DMMA peak {tf:.2f} TFLOP/s at {mhz:.0f} MHz")
t, y, yerr = make_lc(N, seed=1)
eng.register_lc(t, y, yerr)
F = o.flops_per_eval(N)
for B in Bs:
    th = o.sample_from_prior(B, seed=7)
    for exact in (False, True):
        eng.lnlike_batch(th, exact=exact)
        ms = []
        for _ in range(3):
            got = eng.lnlike_batch(th, exact=exact); ms.append(eng.last_kernel_ms())
        m = min(ms)
        print(f"N={N} B={B} exact={int(exact)}: {m:.2f} ms  {B/m*1e3:.0f} evals/s  {B*F/m/1e9:.2f} TFLOP/s  = {B*F/m/1e9/tf*100:.1f}% of DMMA peak", flush=True)
    ref = np.array([o.lnlike_function(x, t, y, yerr) for x in th[:6]])
    print("   parity (first 6):", np.abs(got[:6] - ref) / np.abs(ref))
