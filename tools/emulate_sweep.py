"""NumPy emulation of the diagonal sweep and the tile-level TRSM of gp_kernel.cuh on chosen cases of tools/parity_campaign.py:
panel / off-diagonal tile solves through the explicit 8 x 8 inverse ("inv"), with one step of iterative refinement ("refine") or by
substitution ("sub"), each against the 80-bit evaluation.  usage: python tools/emulate_sweep.py <campaign seed> <n> [<n> ...]
Test infrastructure (imports oracle/); not product code."""
import sys, os, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import gp_oracle as o
from gprotation_b200.synth import synthetic_lightcurve
import emulate_trsm as E
from scipy.linalg import solve_triangular

def diag_gpu(S, ytil, zmode, panelmode):
    """8-wide right-looking sweep on the (nb x nb) block S like diag_factor_rows; returns L, z"""
    nb = S.shape[0]
    M = np.tril(S).copy()
    Wt = {}  # tile inverses
    nt = (nb + 7)//8
    for p in range(nt):
        a, b = 8*p, min(nb, 8*p+8)
        Lpp = np.linalg.cholesky(M[a:b, a:b])
        Wpp = solve_triangular(Lpp, np.eye(b-a), lower=True)
        M[a:b, a:b] = Lpp
        Wt[p] = Wpp
        if b < nb:
            if panelmode == "inv":
                M[b:, a:b] = M[b:, a:b] @ Wpp.T
            elif panelmode == "refine":
                T0 = M[b:, a:b].copy()
                X0 = T0 @ Wpp.T
                R = T0 - X0 @ Lpp.T
                M[b:, a:b] = X0 + R @ Wpp.T
            else:
                M[b:, a:b] = solve_triangular(Lpp, M[b:, a:b].T, lower=True).T
            X = M[b:, a:b]
            M[b:, b:] -= np.tril(X @ X.T)
    L = np.tril(M)
    if zmode == "fullinv":
        # W via tile-level column substitution on identity (what the sweep does)
        W = np.zeros((nb, nb))
        for p in range(nt):
            a, b = 8*p, min(nb, 8*p+8)
            W[a:b, a:b] = Wt[p]
        # W rows: solve W L = I by blocks: W[c, m] for c>m : -W_cc * sum_{m<=k<c} L[c,k] W[k,m]
        for c in range(nt):
            ca, cb = 8*c, min(nb, 8*c+8)
            for m in range(c):
                ma, mb = 8*m, min(nb, 8*m+8)
                acc = np.zeros((cb-ca, mb-ma))
                for k in range(m, c):
                    ka, kb = 8*k, min(nb, 8*k+8)
                    acc += L[ca:cb, ka:kb] @ W[ka:kb, ma:mb]
                W[ca:cb, ma:mb] = -Wt[c] @ acc
        z = W @ ytil
    elif zmode == "sub8":
        z = np.zeros(nb); y = ytil.copy()
        for c in range(nt):
            a, b = 8*c, min(nb, 8*c+8)
            z[a:b] = Wt[c] @ y[a:b]
            y[b:] -= L[b:, a:b] @ z[a:b]
    else:
        z = solve_triangular(L, ytil, lower=True)
    return L, Wt, z

REFINE_OFF = [False]
def lnlike_gpu(K, y, bs, zmode="fullinv", panelmode="inv", offmode="grp8"):
    n = K.shape[0]; T = (n+bs-1)//bs
    L = np.zeros_like(K); yt = y.copy(); quad = logdet = 0.0
    for J in range(T):
        j0, j1 = J*bs, min(n, (J+1)*bs)
        # left-looking accumulate in k order starting from -K (emulate order: acc = -K; acc += L L^T per 8-slab...)
        S = K[j0:j1, j0:j1].copy()
        for k0 in range(0, j0, 16):
            S -= L[j0:j1, k0:k0+16] @ L[j0:j1, k0:k0+16].T
        Ljj, Wt, z = diag_gpu(S, yt[j0:j1], zmode, panelmode)
        L[j0:j1, j0:j1] = Ljj
        quad += float(z @ z); logdet += float(np.sum(np.log(np.diag(Ljj))))
        for I in range(J+1, T):
            i0, i1 = I*bs, min(n, (I+1)*bs)
            Sij = K[i0:i1, j0:j1].copy()
            for k0 in range(0, j0, 16):
                Sij -= L[i0:i1, k0:k0+16] @ L[j0:j1, k0:k0+16].T
            if offmode == "grp8":
                X = np.zeros_like(Sij); Tm = Sij.copy(); nb = j1-j0
                for c in range((nb+7)//8):
                    a, b = 8*c, min(nb, 8*c+8)
                    X[:, a:b] = Tm[:, a:b] @ Wt[c].T
                    if REFINE_OFF[0]:
                        R = Tm[:, a:b] - X[:, a:b] @ Ljj[a:b, a:b].T
                        X[:, a:b] += R @ Wt[c].T
                    if b < nb: Tm[:, b:] -= X[:, a:b] @ Ljj[b:, a:b].T
            else:
                X = solve_triangular(Ljj, Sij.T, lower=True).T
            L[i0:i1, j0:j1] = X
            yt[i0:i1] -= X @ z
    return -0.5*quad - 0.5*(2*logdet + n*np.log(2*np.pi))

seed = int(sys.argv[1]); want = set(int(a) for a in sys.argv[2:])
rng = np.random.default_rng(seed)
for case in range(400):
    n = int(rng.choice([rng.integers(1, 130), rng.integers(130, 700), rng.integers(700, 1600)]))
    kind = "gp" if (n <= 600 and rng.random() < 0.5) else "harmonic"
    s1 = int(rng.integers(1 << 30)); bl = float(rng.uniform(0.3, 3.0)) * max(n, 40)
    if n in want: t, y, yerr, truth = synthetic_lightcurve(s1, n, baseline=bl, kind=kind)
    het = rng.random() < 0.3
    if het:
        f = rng.uniform(0.5, 20.0, size=n)
        if n in want: yerr = yerr * f
    cs = None if (n < 4 or rng.random() < 0.5) else int(rng.integers(max(2, n // 8), n + 1))
    s2 = int(rng.integers(1 << 30)); cl = int(rng.choice([0, 1, 2, 4, 64]))
    if n not in want: continue
    th = o.sample_from_prior(5, seed=s2); th[0] = truth
    chunks = o.split_chunks(n, cs)
    for k, x in enumerate(th):
        for c in ([np.arange(n)] if chunks is None else chunks):
            K = o.kernel_matrix(x, t[c], yerr[c]); ev = np.linalg.eigvalsh(K); cond = ev[-1]/ev[0]
            if cond < 5e7: continue
            tl = E.ld(x, t[c], y[c], yerr[c]); lap = o.lnlike_function(x, t[c], y[c], yerr[c])
            res = {}
            for bs in (128,):
                for pm in ("inv", "refine", "sub"):
                    for ro in (False, True):
                        REFINE_OFF[0] = ro
                        res[(bs, pm, "offdiag refine" if ro else "offdiag inv")] = abs(lnlike_gpu(K, y[c], bs, "fullinv", pm) - tl)/abs(tl)
            print("n", n, "cl", cl, "cond %.1e" % cond, "lapack %.1e" % (abs(lap-tl)/abs(tl)))
            for kk, v in res.items(): print("   ", kk, "%.1e" % v)
