"""Reference-held goldens: run the REFERENCE's own arithmetic (george 0.2.x) on the committed golden inputs.

The build container has no george and the reference ships no golden values, so parity at the
george boundary is unpinned until somebody runs this script where george 0.2.x (the API gprot/model.py uses:
WhiteKernel, solver=george.HODLRSolver, gp.lnlikelihood(y, quiet=True)) is installed:

    pip install "george<0.3"            # needs Eigen headers; any machine with network access
    python tests/golden/make_george_golden.py

It reads the inputs of tests/golden/lnlike_golden.npz (time stamps, fluxes, errors, theta rows, chunk sizes) and writes
tests/golden/george_golden.npz with, per case, the value of the call sequence of gprot/model.py:328-365
(gp_kernel -> george.GP(k, solver=...) -> gp.compute(x, sqrt(sigma + yerr**2)) -> gp.lnlikelihood(y, quiet=True),
summed over the chunk list) for BOTH solvers: george.BasicSolver (dense Cholesky) and george.HODLRSolver (what the
reference uses).  tests/test_george.py compares the oracle (CPU) and the CUDA path (GPU box) against that file when
it exists, and against a live george when one is importable.

The statements between the two marker comments are the reference's, call for call; nothing else in this repository
imports george.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def george_lnlike(theta, x_list, y_list, yerr_list, solver_name):
    """gprot/model.py:327-365 with the solver as a parameter (the reference hard-codes HODLRSolver)."""
    import george
    from george.kernels import ExpSine2Kernel, ExpSquaredKernel, WhiteKernel

    solver = getattr(george, solver_name)
    total = 0.0
    for x, y, yerr in zip(x_list, y_list, yerr_list):
        # ---- reference call sequence begins (gprot/model.py:328-356)
        A = np.exp(theta[0])
        l = np.exp(theta[1])
        G = np.exp(theta[2])
        sigma = np.exp(theta[3])
        P = np.exp(theta[4])
        k = A * ExpSquaredKernel(l) * ExpSine2Kernel(G, P) + WhiteKernel(sigma)
        try:
            gp = george.GP(k, solver=solver)
            gp.compute(x, np.sqrt(sigma + yerr ** 2))
        except (ValueError, np.linalg.LinAlgError):
            return -np.inf
        lnl = gp.lnlikelihood(y, quiet=True)
        lnl = lnl if np.isfinite(lnl) else -np.inf
        # ---- reference call sequence ends
        total += lnl
    return total


def chunk_lists(t, y, yerr, chunksize):
    """gprot/lc.py:329-342."""
    if chunksize is None or chunksize < 0:
        return [t], [y], [yerr]
    n = len(t)
    m = n // chunksize + (1 if n % chunksize else 0)
    return np.array_split(t, m), np.array_split(y, m), np.array_split(yerr, m)


def main():
    try:
        import george
    except ImportError:
        sys.exit("george is not installed here: parity stays unpinned (see the docstring for how to produce the file)")
    version = getattr(george, "__version__", "unknown")
    if not hasattr(george.kernels, "WhiteKernel") or not hasattr(george, "HODLRSolver"):
        sys.exit("george %s does not have the 0.2.x API the reference uses (WhiteKernel, HODLRSolver)" % version)
    gold = np.load(os.path.join(HERE, "lnlike_golden.npz"))
    names = sorted({k[: -len("_theta")] for k in gold.files if k.endswith("_theta")})
    out = {"george_version": np.array(version)}
    for name in names:
        t, y, yerr = gold[name + "_t"], gold[name + "_y"], gold[name + "_yerr"]
        lists = chunk_lists(t, y, yerr, int(gold[name + "_chunksize"]))
        for solver in ("BasicSolver", "HODLRSolver"):
            out["%s_%s" % (name, solver)] = np.array([george_lnlike(th, *lists, solver_name=solver) for th in gold[name + "_theta"]])
        ref = gold[name + "_lnlike"]
        for solver in ("BasicSolver", "HODLRSolver"):
            got = out["%s_%s" % (name, solver)]
            m = np.isfinite(ref)
            print("%-10s %-12s max |george - oracle| / |oracle| = %.2e" % (name, solver, np.max(np.abs(got[m] - ref[m]) / np.abs(ref[m]))))
    np.savez_compressed(os.path.join(HERE, "george_golden.npz"), **out)
    print("wrote tests/golden/george_golden.npz (george %s)" % version)


if __name__ == "__main__":
    main()
