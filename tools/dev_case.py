"""Replays chosen cases of tools/parity_campaign.py (seed 41) on the GPU with every kernel variant, against LAPACK and the
80-bit evaluation.  usage: python tools/dev_case.py 0 88 279"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import gp_oracle as o
from gprotation_b200 import GPEngine
from gprotation_b200.synth import synthetic_lightcurve
import emulate_trsm as E

want = set(int(a) for a in sys.argv[1:]) or {0, 88}
rng = np.random.default_rng(41)
eng = GPEngine(0)
for case in range(max(want) + 1):
    n = int(rng.choice([rng.integers(1, 130), rng.integers(130, 700), rng.integers(700, 1600)]))
    kind = "gp" if (n <= 600 and rng.random() < 0.5) else "harmonic"
    s1 = int(rng.integers(1 << 30)); bl = float(rng.uniform(0.3, 3.0)) * max(n, 40)
    if case in want:
        t, y, yerr, truth = synthetic_lightcurve(s1, n, baseline=bl, kind=kind)
    if rng.random() < 0.3:
        f = rng.uniform(0.5, 20.0, size=n)
        if case in want: yerr = yerr * f
    cs = None if (n < 4 or rng.random() < 0.5) else int(rng.integers(max(2, n // 8), n + 1))
    s2 = int(rng.integers(1 << 30)); cl0 = int(rng.choice([0, 1, 2, 4, 64]))
    if case not in want: continue
    th = o.sample_from_prior(5, seed=s2); th[0] = truth
    chunks = o.split_chunks(n, cs)
    eng.unregister_all()
    lc = eng.register_lc(t, y, yerr, chunks=chunks)
    parts = [np.arange(n)] if chunks is None else chunks
    ld = np.array([sum(E.ld(x, t[c], y[c], yerr[c]) for c in parts) for x in th])
    lap = np.array([o.lnlike(x, t, y, yerr, chunks) for x in th])
    print("case %d n=%d cs=%s kind=%s: lapack vs ld %s" % (case, n, cs, kind, np.array2string(np.abs(lap - ld) / np.abs(ld), precision=1)))
    for cl in (1, 2, 64):
        for exact in (False, True):
            eng.set_cluster(cl)
            got = eng.lnlike_batch(th, lc_id=lc, exact=exact)
            print("   cl=%2d exact=%d: gpu vs ld %s" % (cl, exact, np.array2string(np.abs(got - ld) / np.abs(ld), precision=1)))
eng.set_cluster(0)
