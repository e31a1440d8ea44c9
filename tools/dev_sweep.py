"""Shape sweep on the GPU box: the default gprot-fit layout (300-point chunks, 500 walkers), C1, C3-like and C4 shapes.
usage: python tools/dev_sweep.py [case ...]   cases: chunks small c1 c3 c4"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from oracle import gp_oracle as o
from gprotation_b200 import GPEngine
from gprotation_b200.synth import synthetic_lightcurve

cases = sys.argv[1:] or ["chunks", "small", "c1", "c3", "c4"]
eng = GPEngine(0)
tf, mhz = eng.measure_dmma_peak()
print(f"DMMA peak {tf:.2f} TFLOP/s at {mhz:.0f} MHz", flush=True)


def run(label, th, lc_id, nprob_per_row, flops_per_row, check=None, reps=3):
    eng.lnlike_batch(th, lc_id=lc_id)
    ms = []
    for _ in range(reps):
        got = eng.lnlike_batch(th, lc_id=lc_id)
        ms.append(eng.last_kernel_ms())
    m = min(ms)
    B = len(th)
    print(f"{label}: {m:.3f} ms  {B/m*1e3:.0f} rows/s  {B*nprob_per_row/m*1e3:.0f} problems/s  "
          f"{B*flops_per_row/m/1e9:.2f} TFLOP/s = {B*flops_per_row/m/1e9/tf*100:.1f}% of DMMA peak", flush=True)
    if check is not None:
        ref = np.array([check(x) for x in th[:4]])
        print("   parity (first 4):", np.abs(got[:4] - ref) / np.abs(ref), flush=True)
    return got


if "chunks" in cases:
    # gprot-fit defaults: ~2000 points after subsampling, chunksize 300, 500 walkers
    for ntot, cs, B in ((2000, 300, 500), (2000, 300, 64), (4000, 300, 500)):
        eng.unregister_all()
        t, y, yerr, _ = synthetic_lightcurve(3, ntot, kind="harmonic")
        chunks = o.split_chunks(ntot, cs)
        lc = eng.register_lc(t, y, yerr, chunks=chunks)
        th = o.sample_from_prior(B, seed=5)
        F = sum(o.flops_per_eval(len(c)) for c in chunks)
        run(f"chunks ntot={ntot} cs={cs} ({len(chunks)} chunks of ~{len(chunks[0])}) B={B}", th, lc, len(chunks), F,
            check=lambda x: o.lnlike(x, t, y, yerr, chunks))

if "small" in cases:
    for N, B in ((128, 4096), (300, 4096), (600, 2048)):
        eng.unregister_all()
        t, y, yerr, _ = synthetic_lightcurve(4, N, kind="harmonic")
        lc = eng.register_lc(t, y, yerr)
        th = o.sample_from_prior(B, seed=6)
        run(f"single N={N} B={B}", th, lc, 1, o.flops_per_eval(N), check=lambda x: o.lnlike_function(x, t, y, yerr))

if "c1" in cases:
    eng.unregister_all()
    t, y, yerr, _ = synthetic_lightcurve(0, 1000, kind="gp")
    lc = eng.register_lc(t, y, yerr)
    for B in (32, 148, 1024):
        th = o.sample_from_prior(B, seed=7)
        run(f"C1 N=1000 B={B}", th, lc, 1, o.flops_per_eval(1000), check=lambda x: o.lnlike_function(x, t, y, yerr))

if "c3" in cases:
    eng.unregister_all()
    nstar, W = 32, 64
    lcs = []
    for s in range(nstar):
        t, y, yerr, _ = synthetic_lightcurve(s, 2048, kind="harmonic")
        lcs.append((eng.register_lc(t, y, yerr), t, y, yerr))
    th = o.sample_from_prior(nstar * W, seed=8)
    ids = np.repeat(np.array([l[0] for l in lcs], dtype=np.int32), W)
    got = run(f"C3-like {nstar} stars x {W} walkers N=2048", th, ids, 1, o.flops_per_eval(2048))
    for k in (0, W * 5 + 3, nstar * W - 1):
        _, t, y, yerr = lcs[ids[k]]
        ref = o.lnlike_function(th[k], t, y, yerr)
        print(f"   row {k} star {ids[k]}: rel err {abs(got[k]-ref)/abs(ref):.2e}", flush=True)

if "c4" in cases:
    eng.unregister_all()
    t, y, yerr, _ = synthetic_lightcurve(1, 8192, kind="harmonic")
    lc = eng.register_lc(t, y, yerr)
    for B in (32, 148, 256):
        th = o.sample_from_prior(B, seed=9)
        base = None
        for cl in (1, 2, 4, 8, 0):
            eng.set_cluster(cl)
            got = run(f"C4 N=8192 B={B} cluster={cl}", th, lc, 1, o.flops_per_eval(8192), reps=1)
            print(f"   used cluster {eng.last_cluster()}", flush=True)
            if base is None:
                base = got
            assert np.array_equal(got, base)
        if B == 32:
            ref = o.lnlike_function(th[0], t, y, yerr)
            print(f"   parity row 0: {abs(got[0]-ref)/abs(ref):.2e}", flush=True)
    eng.set_cluster(0)

if "cl" in cases:
    for N in (1000, 2048):
        eng.unregister_all()
        t, y, yerr, _ = synthetic_lightcurve(2, N, kind="harmonic")
        lc = eng.register_lc(t, y, yerr)
        for B in (16, 32, 64, 148, 200):
            th = o.sample_from_prior(B, seed=10)
            base = None
            for cl in (1, 2, 4, 8, 0):
                eng.set_cluster(cl)
                got = run(f"N={N} B={B} cluster={cl}", th, lc, 1, o.flops_per_eval(N))
                if cl == 0:
                    print(f"   auto used cluster {eng.last_cluster()}", flush=True)
                if base is None:
                    base = got
                assert np.array_equal(got, base)
    eng.set_cluster(0)

if "mixed" in cases:
    # adverse order: many small problems first, a few large ones last (largest-first queueing should hide the tail)
    eng.unregister_all()
    t, y, yerr, _ = synthetic_lightcurve(6, 300, kind="harmonic")
    small = eng.register_lc(t, y, yerr)
    t2, y2, yerr2, _ = synthetic_lightcurve(7, 2048, kind="harmonic")
    big = eng.register_lc(t2, y2, yerr2)
    th = o.sample_from_prior(3000 + 40, seed=11)
    ids = np.array([small] * 3000 + [big] * 40, dtype=np.int32)
    F = (3000 * o.flops_per_eval(300) + 40 * o.flops_per_eval(2048)) / len(th)
    run("mixed 3000 x N=300 then 40 x N=2048", th, ids, 1, F)
    run("same, large ones listed first", th[::-1].copy(), ids[::-1].copy(), 1, F)
