"""Phase timeline of gp_lnlike_kernel (CTA 0) from the -DGPROT_TIMING build.  GPU box only.
usage: GPROT_B200_LIB=libgprot_b200_timing.so python tools/dev_timing.py [N] [B]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, '.')
from oracle import gp_oracle as o
from gprotation_b200 import GPEngine
from gprotation_b200.synth import synthetic_lightcurve

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
MODE = int(sys.argv[3]) if len(sys.argv) > 3 else 0        # gprot_set_cluster: 0 automatic, 1 one CTA per problem, 64 two problems per SM
names = ["setup", "diag synth", "diag accumulate", "diag M write", "diag factor+inverse", "W convert + z", "offdiag synth",
         "offdiag accumulate", "TRSM + L z partials + store L", "barrier", "TMA prologue issue + y update", "column fence",
         "stage column points (two-per-SM kernel)", "-", "-", "queue pop"]
eng = GPEngine(0)
t, y, yerr, _ = synthetic_lightcurve(1, N, kind="harmonic")
eng.register_lc(t, y, yerr)
eng.set_cluster(MODE)
th = o.sample_from_prior(B, seed=7)
eng.lnlike_batch(th)
eng.lnlike_batch(th)
ms = eng.last_kernel_ms()
tim = np.zeros(16, dtype=np.int64)
eng._check(eng._lib.gprot_debug_read_timing(eng._ctx, tim.ctypes.data))
tot = tim.sum()
print(f"N={N} B={B} kernel choice {eng.last_cluster()}: kernel {ms:.3f} ms; CTA 0 cycles {tot} ({tot/1.965e6:.3f} ms at 1965 MHz)")
for k, v in zip(names, tim):
    if v:
        print(f"  {k:26s} {v:12d} cycles  {100*v/tot:5.1f}%")
