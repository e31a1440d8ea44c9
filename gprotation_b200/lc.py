"""Light-curve container: the input-layout provider of the hot path.

Mirrors the part of the reference's LightCurve that the likelihood reads (gprot/lc.py:23-42,
329-342, 375-459): ``x / y / yerr`` arrays, optional random subsampling, and the lazily built
chunk lists ``x_list / y_list / yerr_list`` (``np.array_split`` into chunks no larger than
``chunksize``; ``None`` when ``chunksize is None`` -- the ``--nochunks`` single N x N problem).
Also the front end gprot-fit needs around it: sigma clipping, ``best_sublc`` /
``make_best_chunks`` (the ``--clever`` tiered chunks, gprot/lc.py:226-281) and the Aigrain hares-and-hounds text
loader (gprot/aigrain.py:17-56, 117-127).  ACF, bandpass filtering, quarter tables, plotting and data download are
host pre-processing that runs once per star and stay out of scope.
"""
from __future__ import annotations

import os

import numpy as np


def sigma_clip(x, y, yerr, nsigma):
    """gprot/filter.py:5-10: drop points further than nsigma * (rms about the median) from the median."""
    med = np.median(y)
    std = (np.sum((med - y) ** 2) / float(len(y))) ** 0.5
    m = np.abs(y - med) > (nsigma * std)
    return x[~m], y[~m], yerr[~m]


class LightCurve(object):
    def __init__(self, x, y, yerr, name=None, chunksize=200, sub=None):
        self._x = np.array(x, dtype=np.float64)
        self._y = np.array(y, dtype=np.float64)
        self._yerr = np.array(yerr, dtype=np.float64)
        self.chunksize = chunksize
        self._name = name
        self._x_list = self._y_list = self._yerr_list = None
        self.version = 0                      # counts every change of the arrays / chunk lists (see _invalidate)
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
        self.version = getattr(self, "version", 0) + 1      # GPRotModel re-registers the light curve when this moves

    # ---- sampling / chunking
    def subsample(self, sub, seed=None):
        """Keep a random 1/sub of the points, sorted (gprot/lc.py:406-416; legacy NumPy RNG, same draw order)."""
        if sub is None:
            return
        n = len(self._x)
        if seed is not None:                     # np.random.seed(seed); np.random.choice(...) in the reference
            np.random.seed(seed)
        inds = np.sort(np.random.choice(n, n // sub, replace=False))
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

    # ---- full-resolution arrays and the tiered chunks of gprot-fit --clever
    @property
    def x_full(self):
        return self._x_full

    @property
    def y_full(self):
        return self._y_full

    @property
    def yerr_full(self):
        return self._yerr_full

    def sigma_clip(self, nsigma):
        """gprot/lc.py:397-398."""
        self._x, self._y, self._yerr = sigma_clip(self._x, self._y, self._yerr, nsigma)
        self._invalidate()

    def best_sublc(self, ndays, npoints=600, chunksize=300, flat_order=3, **kwargs):
        """New sub-LightCurve over the ``ndays``-long window of the full-resolution data with the largest rms after
        5-sigma clipping and removal of a polynomial of order ``flat_order``; windows advance by 1/50 of their length;
        the result is subsampled to about ``npoints`` points (gprot/lc.py:226-265)."""
        xf, yf, ef = self.x_full, self.y_full, self.yerr_full
        width = int(ndays / np.median(np.diff(xf)))
        hop = max(width // 50, 1)

        def variability(lo):
            xs, ys, _ = sigma_clip(xf[lo:lo + width].copy(), yf[lo:lo + width].copy(), yf[lo:lo + width], 5)
            return np.std(ys - np.polyval(np.polyfit(xs, ys, flat_order), xs))

        starts = range(0, max(len(xf) - width, 0), hop)        # windows that end before the last point, as the reference scans
        scores = [variability(lo) for lo in starts]
        if scores and max(scores) > 0:
            lo = starts[int(np.argmax(scores))]                # first maximum wins, like a strict '>' scan
            hi = lo + width
        else:                                                   # window longer than the light curve: use all of it
            lo, hi = 0, len(xf)
        kwargs.setdefault('sub', max(width // npoints, 1))
        return LightCurve(xf[lo:hi], yf[lo:hi], ef[lo:hi], chunksize=chunksize,
                          name=self.name + '_{:.0f}d'.format(ndays), **kwargs)

    def make_best_chunks(self, ndays=(800, 200, 50), seed=None, **kwargs):
        """Chunk lists built from the most variable 800 / 200 / 50 day windows (gprot/lc.py:267-281); the
        sub-light-curves draw their subsamples from the global NumPy RNG seeded here, as the reference does."""
        if not hasattr(ndays, '__iter__'):
            ndays = [ndays]
        xl, yl, el = [], [], []
        np.random.seed(seed)
        for nd in ndays:
            lc = self.best_sublc(nd, **kwargs)
            xl += list(lc.x_list)
            yl += list(lc.y_list)
            el += list(lc.yerr_list)
        self._x_list, self._y_list, self._yerr_list = xl, yl, el
        self.version = getattr(self, "version", 0) + 1

    @property
    def is_split(self):
        return self.x_list is not None

    @property
    def df(self):
        import pandas as pd

        return pd.DataFrame({"x": self.x, "y": self.y, "yerr": self.yerr})


class AigrainLightCurve(LightCurve):
    """Aigrain et al. (2015) hares-and-hounds simulation ``i`` (gprot/aigrain.py:17-56): two-column text file
    ``<dir>/<subdir>/lightcurve_%04d.txt`` (time, relative flux), ``y - 1``, constant ``yerr = 1e-5``, subsample
    seeded with the star index, optional range restriction, 5-sigma clip.  ``directory`` defaults to the
    AIGRAIN_ROTATION environment variable (gprot/config.py:3-4)."""
    subdir = 'final'

    def __init__(self, i, ndays=None, sub=40, rng=None, nsigma=5, directory=None, **kwargs):
        self.i = i
        self.quarters = None
        if directory is None:
            directory = os.getenv('AIGRAIN_ROTATION', '../code/simulations/kepler_diffrot_full')
        self.directory = directory
        sid = str(int(i)).zfill(4)
        x, y = np.genfromtxt(os.path.join(directory, self.subdir, "lightcurve_{0}.txt".format(sid))).T
        yerr = np.ones(len(y)) * 1e-5
        super(AigrainLightCurve, self).__init__(x, y - 1, yerr, sub=sub, **kwargs)
        if rng is not None:
            self.restrict_range(rng)
        elif ndays is not None:
            self.restrict_range((0, ndays))
        if nsigma is not None:
            self.sigma_clip(nsigma)
        self._sim_params = None

    @property
    def name(self):
        if getattr(self, '_name', None) is None:
            self._name = str(self.i)
        return self._name

    @name.setter
    def name(self, value):
        self._name = value

    def subsample(self, *args, **kwargs):
        if 'seed' not in kwargs:
            kwargs['seed'] = self.i
        super(AigrainLightCurve, self).subsample(*args, **kwargs)

    def make_best_chunks(self, *args, **kwargs):
        if 'seed' not in kwargs:
            kwargs['seed'] = self.i
        super(AigrainLightCurve, self).make_best_chunks(*args, **kwargs)

    @property
    def sim_params(self):
        """Row ``i`` of ``<dir>/par/final_table.txt`` (gprot/aigrain.py:91-95, 117-127)."""
        if self._sim_params is None:
            import pandas as pd

            df = pd.read_csv(os.path.join(self.directory, 'par', 'final_table.txt'), sep=r'\s+')
            self._sim_params = df.iloc[int(self.i)]
        return self._sim_params


class NoiseFreeAigrainLightCurve(AigrainLightCurve):
    subdir = 'noise_free'
