"""Light-curve container: the input-layout provider of the hot path.

Mirrors the part of the reference's LightCurve that the likelihood reads (gprot/lc.py:23-42,
329-342, 375-459): ``x / y / yerr`` arrays, optional random subsampling, and the lazily built
chunk lists ``x_list / y_list / yerr_list`` (``np.array_split`` into chunks no larger than
``chunksize``; ``None`` when ``chunksize is None`` -- the ``--nochunks`` single N x N problem).
ACF, filtering, quarter splitting, plotting and data download are host pre-processing that runs
once per star and are out of scope.
"""
from __future__ import annotations

import numpy as np


class LightCurve(object):
    def __init__(self, x, y, yerr, name=None, chunksize=200, sub=None):
        self._x = np.array(x, dtype=np.float64)
        self._y = np.array(y, dtype=np.float64)
        self._yerr = np.array(yerr, dtype=np.float64)
        self.chunksize = chunksize
        self._name = name
        self._x_list = self._y_list = self._yerr_list = None
        self._x_full, self._y_full, self._yerr_full = self._x.copy(), self._y.copy(), self._yerr.copy()
        self.sub = sub
        self.subsample(sub)

    # ---- identity
    @property
    def name(self):
        if self._name is None:
            self._name = "lc"
        return self._name

    @name.setter
    def name(self, value):
        self._name = value

    # ---- arrays
    @property
    def x(self):
        return self._x

    @x.setter
    def x(self, v):
        self._x = np.asarray(v, dtype=np.float64)
        self._invalidate()

    @property
    def y(self):
        return self._y

    @y.setter
    def y(self, v):
        self._y = np.asarray(v, dtype=np.float64)
        self._invalidate()

    @property
    def yerr(self):
        return self._yerr

    @yerr.setter
    def yerr(self, v):
        self._yerr = np.asarray(v, dtype=np.float64)
        self._invalidate()

    def _invalidate(self):
        self._x_list = self._y_list = self._yerr_list = None

    # ---- sampling / chunking
    def subsample(self, sub, seed=None):
        """Keep a random 1/sub of the points, sorted (gprot/lc.py:406-416; legacy NumPy RNG, same draw order)."""
        if sub is None:
            return
        n = len(self._x)
        rs = np.random.RandomState(seed)
        inds = np.sort(rs.choice(n, n // sub, replace=False))
        self._x, self._y, self._yerr = self._x[inds], self._y[inds], self._yerr[inds]
        self._invalidate()

    def restrict_range(self, rng):
        m = (self._x > rng[0]) & (self._x < rng[1])
        self._x, self._y, self._yerr = self._x[m], self._y[m], self._yerr[m]
        self._invalidate()

    def _make_chunks(self, chunksize=None):
        if chunksize is None:
            if self.chunksize is None:
                return
            chunksize = self.chunksize
        nall = len(self._x)
        m = nall // chunksize
        if nall % chunksize != 0:      # chunks no larger than chunksize
            m += 1
        self.chunksize = chunksize
        self._x_list = np.array_split(self._x, m)
        self._y_list = np.array_split(self._y, m)
        self._yerr_list = np.array_split(self._yerr, m)

    @property
    def x_list(self):
        if self._x_list is None:
            self._make_chunks()
        return self._x_list

    @property
    def y_list(self):
        if self._y_list is None:
            self._make_chunks()
        return self._y_list

    @property
    def yerr_list(self):
        if self._yerr_list is None:
            self._make_chunks()
        return self._yerr_list

    @property
    def is_split(self):
        return self.x_list is not None

    @property
    def df(self):
        import pandas as pd

        return pd.DataFrame({"x": self.x, "y": self.y, "yerr": self.yerr})
