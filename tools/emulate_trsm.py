"""NumPy emulation of the blocked left-looking Cholesky of gp_kernel.cuh with different TRSM forms, against the
80-bit evaluation (oracle/gp_oracle_ld.c).  Decides the group width of the two-level solve.

  python tools/emulate_trsm.py [ncases] [seed]

Forms of L_IJ = S_IJ L_JJ^-T and z_J = L_JJ^-1 ytilde_J:
  inv128    multiply by the explicit inverse of the 128 x 128 diagonal block (round-1 kernel)
  grp<g>    substitution across g-wide column groups, explicit inverses of the g x g diagonal groups only
  subst     scalar substitution (LAPACK dtrsm)
Test infrastructure (imports oracle/); not product code.
"""
import ctypes
import os
import sys

import numpy as np
from scipy.linalg import solve_triangular

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import gp_oracle as o  # noqa: E402
from gprotation_b200.synth import synthetic_lightcurve  # noqa: E402

_c = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_build", "libgp_oracle.so"))
_c.gp_oracle_lnlike_ld.restype = ctypes.c_double
_c.gp_oracle_lnlike_ld.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]


def ld(x, t, y, yerr):
    x, t, y, yerr = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, t, y, yerr))
    return _c.gp_oracle_lnlike_ld(x.ctypes.data, len(t), t.ctypes.data, y.ctypes.data, yerr.ctypes.data)


def tri_inv(L):
    """explicit inverse of a lower-triangular block by 8-wide Gauss-Jordan steps (what diag_factor_rows does)"""
    return solve_triangular(L, np.eye(L.shape[0]), lower=True)


def solve_right(S, Ljj, form):
    """X with X Ljj^T = S"""
    n = Ljj.shape[0]
    if form == "subst":
        return solve_triangular(Ljj, S.T, lower=True).T
    g = n if form == "inv128" else int(form[3:])
    if g >= n:
        return S @ tri_inv(Ljj).T
    X = np.empty_like(S)
    T = S.copy()
    for c0 in range(0, n, g):
        c1 = min(n, c0 + g)
        X[:, c0:c1] = T[:, c0:c1] @ tri_inv(Ljj[c0:c1, c0:c1]).T
        if c1 < n:
            T[:, c1:] -= X[:, c0:c1] @ Ljj[c1:, c0:c1].T
    return X


def blocked_lnlike(K, y, form, bs=128, zform=None):
    zform = zform or form
    n = K.shape[0]
    T = (n + bs - 1) // bs
    L = np.zeros_like(K)
    yt = y.copy()
    quad = logdet = 0.0
    for J in range(T):
        j0, j1 = J * bs, min(n, (J + 1) * bs)
        S = K[j0:j1, j0:j1] - L[j0:j1, :j0] @ L[j0:j1, :j0].T
        try:
            Ljj = np.linalg.cholesky(S)
        except np.linalg.LinAlgError:
            return -np.inf
        L[j0:j1, j0:j1] = Ljj
        z = solve_right(yt[None, j0:j1], Ljj, zform)[0]
        quad += float(z @ z)
        logdet += float(np.sum(np.log(np.diag(Ljj))))
        for I in range(J + 1, T):
            i0, i1 = I * bs, min(n, (I + 1) * bs)
            Sij = K[i0:i1, j0:j1] - L[i0:i1, :j0] @ L[j0:j1, :j0].T
            Lij = solve_right(Sij, Ljj, form)
            L[i0:i1, j0:j1] = Lij
            yt[i0:i1] -= Lij @ z
    return -0.5 * quad - 0.5 * (2.0 * logdet + n * np.log(2 * np.pi))


if __name__ == "__main__":
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 41
    forms = ["inv128", "grp64", "grp32", "grp16", "grp8", "subst"]
    rng = np.random.default_rng(seed)
    rows = []
    for case in range(ncases):
        n = int(rng.integers(200, 1400))
        t, y, yerr, truth = synthetic_lightcurve(int(rng.integers(1 << 30)), n, baseline=float(rng.uniform(0.3, 3.0)) * max(n, 40), kind="harmonic")
        th = o.sample_from_prior(4, seed=int(rng.integers(1 << 30)))
        th[0] = truth
        for x in th:
            K = o.kernel_matrix(x, t, yerr)
            ev = np.linalg.eigvalsh(K)
            cond = ev[-1] / ev[0] if ev[0] > 0 else np.inf
            truth_ld = ld(x, t, y, yerr)
            if not np.isfinite(truth_ld):
                continue
            lap = o.lnlike_function(x, t, y, yerr)
            errs = {f: abs(blocked_lnlike(K, y, f) - truth_ld) / abs(truth_ld) for f in forms}
            errs["lapack"] = abs(lap - truth_ld) / abs(truth_ld)
            rows.append((cond, errs))
    rows.sort(key=lambda r: r[0])
    names = forms + ["lapack"]
    print("%10s " % "cond" + " ".join("%9s" % f for f in names))
    for cond, e in rows:
        if cond > 1e5:
            print("%10.2e " % cond + " ".join("%9.1e" % e[f] for f in names))
    for lo, hi in ((0, 1e5), (1e5, 1e7), (1e7, 1e9), (1e9, 1e30)):
        sel = [e for c, e in rows if lo <= c < hi]
        if sel:
            print("cond in [%.0e, %.0e): %d evals; max err  " % (lo, hi, len(sel)) + " ".join("%s=%.1e" % (f, max(e[f] for e in sel)) for f in names))
            print("                                 median   " + " ".join("%s=%.1e" % (f, np.median([e[f] for e in sel])) for f in names))
