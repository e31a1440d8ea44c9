"""Bitwise fingerprint of the likelihood for a fixed set of inputs and every kernel variant (to check that a refactor
of the kernels changes no bit): python tools/dev_fingerprint.py out.npy   /   ... out.npy --compare"""
import sys
import numpy as np
sys.path.insert(0, '.')
from oracle import gp_oracle as o
from gprotation_b200 import GPEngine
from gprotation_b200.synth import synthetic_lightcurve

eng = GPEngine(0)
outs = []
for n, cs in ((700, None), (1100, 300), (129, None), (64, None), (2048, None)):
    t, y, yerr, _ = synthetic_lightcurve(500 + n, n, kind="harmonic")
    eng.unregister_all()
    lc = eng.register_lc(t, y, yerr, chunks=o.split_chunks(n, cs))
    th = o.sample_from_prior(12, seed=n)
    for mode in (1, 2, 4, 64):
        for exact in (False, True):
            eng.set_cluster(mode)
            outs.append(eng.lnlike_batch(th, lc_id=lc, exact=exact))
a = np.concatenate(outs)
if "--compare" in sys.argv:
    b = np.load(sys.argv[1])
    same = np.array_equal(a, b, equal_nan=True)
    print("bitwise identical:", same, "" if same else "max rel diff %.3e" % np.nanmax(np.abs(a - b) / np.abs(b)))
    sys.exit(0 if same else 1)
np.save(sys.argv[1], a)
print("saved", a.shape)
