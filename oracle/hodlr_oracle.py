"""HODLR restatement of the reference's ACTUAL solver -- TEST INFRASTRUCTURE / CPU BASELINE, NOT PRODUCT CODE.

gprot/model.py:343 builds ``george.GP(k, solver=george.HODLRSolver)``: george 0.2.x hands the kernel to Ambikasaran's
HODLR library (hierarchically off-diagonal low-rank; Ambikasaran & Darve 2013, J. Sci. Comput. 57, 477;
Ambikasaran et al. 2016, IEEE TPAMI 38, 252), defaults ``nleaf=100, tol=1e-12``.  george is not installable offline
and its sources are not under /root/reference, so this file restates the PUBLISHED algorithm in NumPy:

  * binary tree over the (sorted) points, split in halves down to leaves of <= nleaf points;
  * every off-diagonal block K[left, right] ~= U V^T by partially pivoted adaptive cross approximation (ACA), stopped
    when the new rank-one term is below tol times the running Frobenius norm of the approximation;
  * one-pass factorisation from the leaves up: K_node = diag(K_l, K_r) (I + [[0, K_l^-1 U V^T], [K_r^-1 V U^T, 0]]),
    solved through the 2r x 2r Woodbury system; log det K = log det K_l + log det K_r + log det(2r x 2r system);
  * lnlike = -1/2 (y^T K^-1 y + log det K + N log 2 pi)  (george GP.lnlikelihood, zero mean).

It is NOT bit-compatible with george (pivot choices differ); it answers two questions the dense oracle cannot:
how far HODLR(tol=1e-12) is from the dense factorisation of the same matrix (O(tol * cond), tests/test_oracle.py), and
what an O(N log^2 N) solver costs on the host cores next to the dense port (bench.py ``cpu_baseline.hodlr``) -- the
dense port overstates the CPU cost of the reference at N = 2048.  Only tests/ and bench.py's CPU legs import this.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.linalg import cho_factor, cho_solve, lu_factor, lu_solve

from . import gp_oracle as o


class _Kernel(object):
    """Entry-wise evaluation of the matrix of oracle.gp_oracle.kernel_matrix (same operation order), on demand."""

    def __init__(self, theta, t, yerr):
        self.A, self.L, self.G, self.S, self.P = o.hyper(theta)
        self.t = np.asarray(t, dtype=np.float64)
        e = np.sqrt(self.S + np.asarray(yerr, dtype=np.float64) ** 2)
        self.diag_noise = e * e

    def block(self, rows, cols):
        d = self.t[rows][:, None] - self.t[cols][None, :]
        r2 = d * d / self.L
        s = np.sin(np.pi * d / self.P)
        K = (self.A * np.exp(-0.5 * r2)) * np.exp(-self.G * s * s)
        return K + np.where(r2 < np.finfo(np.float64).eps, self.S, 0.0)

    def dense(self, lo, hi):
        idx = np.arange(lo, hi)
        K = self.block(idx, idx)
        K[np.diag_indices_from(K)] += self.diag_noise[lo:hi]
        return K


def aca(kern, rows, cols, tol, max_rank=None):
    """Partially pivoted ACA of K[rows, cols]: returns U (len(rows) x r), V (len(cols) x r) with K ~= U V^T."""
    m, n = len(rows), len(cols)
    max_rank = min(m, n) if max_rank is None else max_rank
    U, V = np.zeros((m, 0)), np.zeros((n, 0))
    used_r = np.zeros(m, dtype=bool)
    norm2 = 0.0
    i = 0
    for _ in range(max_rank):
        used_r[i] = True
        row = kern.block(rows[i:i + 1], cols)[0] - U[i] @ V.T
        j = int(np.argmax(np.abs(row)))
        piv = row[j]
        if piv == 0.0:
            rest = np.nonzero(~used_r)[0]
            if rest.size == 0:
                break
            i = int(rest[0])
            continue
        v = row / piv
        col = kern.block(rows, cols[j:j + 1])[:, 0] - U @ V[j]
        # running Frobenius norm of the approximation (Bebendorf's update)
        norm2 += 2.0 * float((U.T @ col) @ (V.T @ v)) + float(col @ col) * float(v @ v)
        U = np.column_stack([U, col])
        V = np.column_stack([V, v])
        if math.sqrt(float(col @ col) * float(v @ v)) <= tol * math.sqrt(max(norm2, 0.0)):
            break
        c = np.abs(col)
        c[used_r] = -1.0
        i = int(np.argmax(c))
        if c[i] < 0.0:
            break
    return U, V


class _Node(object):
    __slots__ = ("lo", "hi", "left", "right", "U", "V", "chol", "ZU", "ZV", "lu", "logdet")


def _build(kern, lo, hi, nleaf, tol):
    nd = _Node()
    nd.lo, nd.hi = lo, hi
    if hi - lo <= nleaf:
        nd.left = nd.right = None
        nd.chol = cho_factor(kern.dense(lo, hi), lower=True, check_finite=False)
        nd.logdet = 2.0 * float(np.sum(np.log(np.diag(nd.chol[0]))))
        return nd
    mid = lo + (hi - lo) // 2
    nd.left, nd.right = _build(kern, lo, mid, nleaf, tol), _build(kern, mid, hi, nleaf, tol)
    nd.U, nd.V = aca(kern, np.arange(lo, mid), np.arange(mid, hi), tol)
    # K_l^-1 U and K_r^-1 V, then the 2r x 2r system [[I, U^T Z_U], [V^T Z_V, I]]
    nd.ZU, nd.ZV = _solve(nd.left, nd.U), _solve(nd.right, nd.V)
    r = nd.U.shape[1]
    S = np.eye(2 * r)
    S[:r, r:] = nd.U.T @ nd.ZU
    S[r:, :r] = nd.V.T @ nd.ZV
    nd.lu = lu_factor(S, check_finite=False)
    sign_logdet = np.linalg.slogdet(S)
    if sign_logdet[0] <= 0:
        raise np.linalg.LinAlgError("HODLR: non-positive determinant")
    nd.logdet = nd.left.logdet + nd.right.logdet + float(sign_logdet[1])
    return nd


def _solve(nd, b):
    """K_node^-1 b for b of shape (n,) or (n, k)."""
    if nd.left is None:
        return cho_solve(nd.chol, b, check_finite=False)
    n1 = nd.left.hi - nd.left.lo
    y1, y2 = _solve(nd.left, b[:n1]), _solve(nd.right, b[n1:])
    r = nd.U.shape[1]
    rhs = np.concatenate([nd.U.T @ y1, nd.V.T @ y2], axis=0)
    ca = lu_solve(nd.lu, rhs, check_finite=False)
    c, a = ca[:r], ca[r:]                                   # c = U^T x1, a = V^T x2
    return np.concatenate([y1 - nd.ZU @ a, y2 - nd.ZV @ c], axis=0)


def lnlike_function(theta, t, y, yerr, nleaf=100, tol=1e-12):
    """gprot/model.py:349-356 with the HODLR solver: failure or non-finite -> -inf."""
    try:
        with np.errstate(all="ignore"):
            t = np.asarray(t, dtype=np.float64)
            order = np.argsort(t, kind="stable")            # george.GP.compute(sort=True)
            t, yerr, y = t[order], np.asarray(yerr, dtype=np.float64)[order], np.asarray(y, dtype=np.float64)[order]
            kern = _Kernel(theta, t, yerr)
            if not all(np.isfinite(v) for v in (kern.A, kern.L, kern.G, kern.S, kern.P)):
                return -np.inf
            root = _build(kern, 0, len(t), nleaf, tol)
            alpha = _solve(root, y)
            lnl = -0.5 * float(y @ alpha) - 0.5 * (root.logdet + len(y) * math.log(2.0 * math.pi))
    except (ValueError, np.linalg.LinAlgError):
        return -np.inf
    return lnl if np.isfinite(lnl) else -np.inf


def lnlike(theta, t, y, yerr, chunks=None, **kw):
    """gprot/model.py:358-365."""
    if chunks is None:
        return lnlike_function(theta, t, y, yerr, **kw)
    return float(np.sum([lnlike_function(theta, t[c], y[c], yerr[c], **kw) for c in chunks]))


def ranks(theta, t, yerr, nleaf=100, tol=1e-12):
    """Off-diagonal ranks per tree level (diagnostic)."""
    t = np.sort(np.asarray(t, dtype=np.float64))
    kern = _Kernel(theta, t, np.asarray(yerr, dtype=np.float64))
    out = {}

    def walk(lo, hi, lev):
        if hi - lo <= nleaf:
            return
        mid = lo + (hi - lo) // 2
        U, _ = aca(kern, np.arange(lo, mid), np.arange(mid, hi), tol)
        out.setdefault(lev, []).append(U.shape[1])
        walk(lo, mid, lev + 1)
        walk(mid, hi, lev + 1)
    walk(0, len(t), 0)
    return out
