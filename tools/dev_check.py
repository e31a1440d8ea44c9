"""Development check run on the GPU box: CUDA path vs oracle at several sizes, with block-level diagnostics."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from oracle import gp_oracle as o
from gprotation_b200 import GPEngine

def make_lc(N, seed=0):
    rng = np.random.default_rng(seed)
    full = np.arange(0, 1400, 0.02043365)
    t = full[np.sort(rng.choice(len(full), N, replace=False))]
    yerr = np.full(N, 1e-5)
    truth = np.array([-12., 7., -1., -17., np.log(11.)])
    y = np.linalg.cholesky(o.kernel_matrix(truth, t, yerr)) @ rng.standard_normal(N)
    return t, y, yerr

def unpack_block(buf):
    """tile-major 128x128 block ([slab 8][row group 16][k tile 2][8][8]) -> row-major"""
    B = np.zeros((128, 128))
    d = np.arange(16384)
    slab, r = d // 2048, d % 2048
    rg, r2 = r // 128, r % 128
    kp, r3 = r2 // 64, r2 % 64
    row = rg * 8 + r3 // 8
    col = slab * 16 + 8 * kp + r3 % 8
    B[row, col] = buf
    return B

if __name__ != "__main__": sizes = []
else: sizes = [int(a) for a in sys.argv[1:]] or [64, 128, 200, 256, 300, 384, 1000]
eng = GPEngine(0) if sizes else None
for N in sizes:
    eng.unregister_all()
    t, y, yerr = make_lc(N, seed=N)
    eng.register_lc(t, y, yerr)
    th = o.sample_from_prior(8, seed=N)
    t0 = time.time()
    ref = np.array([o.lnlike_function(x, t, y, yerr) for x in th])
    for exact in (True, False):
        got = eng.lnlike_batch(th, exact=exact)
        rel = np.abs(got - ref) / np.abs(ref)
        print(f"N={N:5d} exact={int(exact)} max rel err {rel.max():.3e}  kernel {eng.last_kernel_ms():.3f} ms", flush=True)
        if not (rel.max() < 1e-6):
            print("   ref", ref[:4], "\n   got", got[:4])
    if N > 128 and N <= 512:
        # block-level check of L for one theta (single problem -> slot 0)
        got = eng.lnlike_batch(th[:1], exact=True)
        K = o.kernel_matrix(th[0], t, yerr)
        T = (N + 127) // 128
        Kp = np.eye(T * 128); Kp[:N, :N] = K
        L = np.linalg.cholesky(Kp)
        for I in range(1, T):
            for J in range(I):
                blk = unpack_block(eng.debug_read_scratch(0, (I * (I - 1) // 2 + J) * 16384, 16384))
                refb = L[I*128:(I+1)*128, J*128:(J+1)*128]
                err = np.abs(blk - refb).max() / max(np.abs(refb).max(), 1e-300)
                print(f"   L block ({I},{J}) max rel-to-max err {err:.2e}")
