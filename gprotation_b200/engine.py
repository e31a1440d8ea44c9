"""Host-side handle on one B200: owns a gprot_ctx, registers light curves, evaluates batches.

Replaces the per-worker george.GP objects of the reference (gprot/model.py:336-347) and the
schwimmbad pool that fans walkers out to processes (scripts/gprot-fit:195, gprot/fit.py:94-98):
a whole ensemble is one launch.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_float, c_int, c_void_p

import numpy as np

from . import _lib


class GProtError(RuntimeError):
    pass


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class GPEngine(object):
    """One context per (process, GPU).  Not re-entrant; calls are stream-ordered."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        self._ctx = c_void_p()
        rc = self._lib.gprot_create(int(device), byref(self._ctx))
        if rc != 0:
            raise GProtError(self._lib.gprot_last_error(None).decode())
        self.device = int(device)
        self._keep = []

    # -- lifetime
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.gprot_destroy(self._ctx)
            self._ctx = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise GProtError(self._lib.gprot_last_error(self._ctx).decode())

    # -- data
    def register_lc(self, t, y, yerr, chunks=None):
        """Register one light curve.  chunks: None (single N x N problem, the --nochunks path) or a list of
        index arrays (lc.x_list semantics, gprot/lc.py:329-342).  Returns the light-curve id."""
        t, y, yerr = _f64(t), _f64(y), _f64(yerr)
        if not (t.shape == y.shape == yerr.shape and t.ndim == 1 and t.size > 0):
            raise ValueError("t, y, yerr must be equal-length non-empty 1-D arrays")
        if chunks is None:
            n = np.array([t.size], dtype=np.int32)
            tt, yy, ee = t, y, yerr
        else:
            chunks = [np.asarray(c) for c in chunks]
            n = np.array([len(c) for c in chunks], dtype=np.int32)
            tt = _f64(np.concatenate([t[c] for c in chunks]))
            yy = _f64(np.concatenate([y[c] for c in chunks]))
            ee = _f64(np.concatenate([yerr[c] for c in chunks]))
        lc_id = c_int(-1)
        self._check(self._lib.gprot_register_lc(self._ctx, int(n.size), n.ctypes.data, tt.ctypes.data, yy.ctypes.data,
                                                ee.ctypes.data, byref(lc_id)))
        return lc_id.value

    def register_lc_arrays(self, x_list, y_list, yerr_list):
        """Register from the reference's chunk lists directly (lc.x_list, lc.y_list, lc.yerr_list)."""
        n = np.array([len(x) for x in x_list], dtype=np.int32)
        tt, yy, ee = _f64(np.concatenate(x_list)), _f64(np.concatenate(y_list)), _f64(np.concatenate(yerr_list))
        lc_id = c_int(-1)
        self._check(self._lib.gprot_register_lc(self._ctx, int(n.size), n.ctypes.data, tt.ctypes.data, yy.ctypes.data,
                                                ee.ctypes.data, byref(lc_id)))
        return lc_id.value

    def unregister_all(self):
        """Forget every light curve; ids restart at 0 and `generation` moves on, which voids the ids models cached."""
        self._check(self._lib.gprot_unregister_all(self._ctx))

    def unregister_lc(self, lc_id):
        """Forget one light curve (its id is not reused)."""
        self._check(self._lib.gprot_unregister_lc(self._ctx, int(lc_id)))

    @property
    def generation(self):
        return int(self._lib.gprot_generation(self._ctx))

    # -- evaluation (host buffers)
    def lnlike_batch(self, theta, lc_id=None, exact=False):
        """lnlike for every row of theta [B,5] -> ndarray [B].  lc_id: None (all rows on light curve 0),
        an int, or an int array [B]."""
        theta = _f64(theta).reshape(-1, 5)
        B = theta.shape[0]
        out = np.empty(B, dtype=np.float64)
        ids = self._ids(lc_id, B)
        flags = _lib.EXACT_KERNEL if exact else 0
        self._check(self._lib.gprot_lnlike_batch(self._ctx, B, theta.ctypes.data, ids.ctypes.data if ids is not None else None,
                                                 out.ctypes.data, flags, None))
        return out

    def lnpost_batch(self, theta, bounds, mu, sigma, pmin=-np.inf, pmax=np.inf, mixture=None, lc_id=None, exact=False):
        """(lnprior[B], lnlike[B]); rows outside the prior support are not factorised and get lnlike = -inf."""
        theta = _f64(theta).reshape(-1, 5)
        B = theta.shape[0]
        lp = np.empty(B, dtype=np.float64)
        ll = np.empty(B, dtype=np.float64)
        b, m, s = _f64(bounds).reshape(10), _f64(mu).reshape(4), _f64(sigma).reshape(4)
        mix = None if mixture is None or len(mixture) == 0 else _f64(mixture).reshape(-1, 3)
        ids = self._ids(lc_id, B)
        flags = _lib.EXACT_KERNEL if exact else 0
        self._check(self._lib.gprot_lnpost_batch(
            self._ctx, B, theta.ctypes.data, ids.ctypes.data if ids is not None else None, b.ctypes.data, m.ctypes.data,
            s.ctypes.data, float(pmin), float(pmax), mix.ctypes.data if mix is not None else None,
            0 if mix is None else mix.shape[0], lp.ctypes.data, ll.ctypes.data, flags, None))
        return lp, ll

    # -- evaluation (device buffers, e.g. torch tensors' data_ptr()); asynchronous on `stream`
    def lnlike_batch_device(self, theta_ptr, B, out_ptr, lc_id=None, stream=0, exact=False):
        ids = self._ids(lc_id, B)
        flags = _lib.THETA_ON_DEVICE | _lib.OUT_ON_DEVICE | (_lib.EXACT_KERNEL if exact else 0)
        self._check(self._lib.gprot_lnlike_batch(self._ctx, int(B), c_void_p(int(theta_ptr)),
                                                 ids.ctypes.data if ids is not None else None, c_void_p(int(out_ptr)),
                                                 flags, c_void_p(int(stream)) if stream else None))

    @staticmethod
    def _ids(lc_id, B):
        if lc_id is None:
            return None
        if np.isscalar(lc_id):
            return np.full(B, int(lc_id), dtype=np.int32)
        ids = np.ascontiguousarray(lc_id, dtype=np.int32)
        if ids.shape != (B,):
            raise ValueError("lc_id must have one entry per theta row")
        return ids

    # -- measurement
    def sync(self):
        self._check(self._lib.gprot_sync(self._ctx))

    def last_kernel_ms(self):
        ms = c_float(0)
        self._check(self._lib.gprot_last_kernel_ms(self._ctx, byref(ms)))
        return ms.value

    def set_cluster(self, ctas_per_problem=0):
        """Kernel choice: 0 = per launch (default); 1 / 2 / 4 / 8 = CTAs per problem of the one-problem-per-SM kernel
        (bitwise identical results among themselves); 64 = the two-problems-per-SM kernel (64 x 64 blocks; agrees with
        the others to fp64 rounding, ~eps * cond)."""
        self._check(self._lib.gprot_set_cluster(self._ctx, int(ctas_per_problem)))

    def last_cluster(self):
        """What the last launch used: 1, 2, 4, 8 CTAs per problem, or 64 for the two-problems-per-SM kernel."""
        return int(self._lib.gprot_last_cluster(self._ctx))

    def launch_count(self):
        return int(self._lib.gprot_launch_count(self._ctx))

    def measure_dmma_peak(self):
        tf, mhz = c_double(0), c_double(0)
        self._check(self._lib.gprot_measure_dmma_peak(self._ctx, byref(tf), byref(mhz)))
        return tf.value, mhz.value

    def debug_read_scratch(self, slot, offset, count):
        out = np.empty(count, dtype=np.float64)
        self._check(self._lib.gprot_debug_read_scratch(self._ctx, slot, offset, count, out.ctypes.data))
        return out


class GPGroup(object):
    """Several GPUs driven from one process (gprot_group_* in include/gprot_b200.h): rows of a batch are block-sharded
    over the devices, light curves are replicated, the host buffer is the gather.  Same register / lnlike surface as
    GPEngine, so GPRotModel(engine=GPGroup([0, 1, ...])) uses the whole box."""

    def __init__(self, devices):
        self._lib = _lib.load()
        self._grp = c_void_p()
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        rc = self._lib.gprot_group_create(int(dev.size), dev.ctypes.data, byref(self._grp))
        if rc != 0:
            raise GProtError(self._lib.gprot_group_last_error(None).decode())
        self.devices = [int(d) for d in dev]
        self._generation = 0

    def close(self):
        if getattr(self, "_grp", None) is not None and self._grp.value:
            self._lib.gprot_group_destroy(self._grp)
            self._grp = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise GProtError(self._lib.gprot_group_last_error(self._grp).decode())

    def __len__(self):
        return int(self._lib.gprot_group_size(self._grp))

    def register_lc(self, t, y, yerr, chunks=None):
        t, y, yerr = _f64(t), _f64(y), _f64(yerr)
        if chunks is None:
            return self.register_lc_arrays([t], [y], [yerr])
        chunks = [np.asarray(c) for c in chunks]
        return self.register_lc_arrays([t[c] for c in chunks], [y[c] for c in chunks], [yerr[c] for c in chunks])

    def register_lc_arrays(self, x_list, y_list, yerr_list):
        n = np.array([len(x) for x in x_list], dtype=np.int32)
        tt, yy, ee = _f64(np.concatenate(x_list)), _f64(np.concatenate(y_list)), _f64(np.concatenate(yerr_list))
        lc_id = c_int(-1)
        self._check(self._lib.gprot_group_register_lc(self._grp, int(n.size), n.ctypes.data, tt.ctypes.data, yy.ctypes.data,
                                                      ee.ctypes.data, byref(lc_id)))
        return lc_id.value

    def unregister_all(self):
        self._check(self._lib.gprot_group_unregister_all(self._grp))
        self._generation += 1

    def unregister_lc(self, lc_id):
        self._check(self._lib.gprot_group_unregister_lc(self._grp, int(lc_id)))

    @property
    def generation(self):
        return self._generation

    def lnlike_batch(self, theta, lc_id=None, exact=False):
        theta = _f64(theta).reshape(-1, 5)
        B = theta.shape[0]
        out = np.empty(B, dtype=np.float64)
        ids = GPEngine._ids(lc_id, B)
        self._check(self._lib.gprot_group_lnlike_batch(self._grp, B, theta.ctypes.data, ids.ctypes.data if ids is not None else None,
                                                       out.ctypes.data, _lib.EXACT_KERNEL if exact else 0))
        return out

    def lnpost_batch(self, theta, bounds, mu, sigma, pmin=-np.inf, pmax=np.inf, mixture=None, lc_id=None, exact=False):
        """(lnprior[B], lnlike[B]) over the devices of the group; rows outside the prior support are not factorised."""
        theta = _f64(theta).reshape(-1, 5)
        B = theta.shape[0]
        lp = np.empty(B, dtype=np.float64)
        ll = np.empty(B, dtype=np.float64)
        b, m, s = _f64(bounds).reshape(10), _f64(mu).reshape(4), _f64(sigma).reshape(4)
        mix = None if mixture is None or len(mixture) == 0 else _f64(mixture).reshape(-1, 3)
        ids = GPEngine._ids(lc_id, B)
        self._check(self._lib.gprot_group_lnpost_batch(
            self._grp, B, theta.ctypes.data, ids.ctypes.data if ids is not None else None, b.ctypes.data, m.ctypes.data,
            s.ctypes.data, float(pmin), float(pmax), mix.ctypes.data if mix is not None else None,
            0 if mix is None else mix.shape[0], lp.ctypes.data, ll.ctypes.data, _lib.EXACT_KERNEL if exact else 0))
        return lp, ll
