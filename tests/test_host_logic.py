"""Host-side logic: light-curve chunking, sharding, the walker adapter, the synthetic generator (CPU only)."""
import os
import sys

import numpy as np
import pytest

from oracle import gp_oracle as o
from gprotation_b200.dist import shard_bounds, sharded_lnlike
from gprotation_b200.fit import Emcee3Model, EnsemblePool, State
from gprotation_b200.lc import LightCurve
from gprotation_b200.synth import KEPLER_CADENCE, synthetic_lightcurve

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lightcurve_chunks_follow_array_split():
    t = np.arange(1000) * KEPLER_CADENCE
    lc = LightCurve(t, np.sin(t), np.full(1000, 1e-5), chunksize=300)
    assert [len(c) for c in lc.x_list] == [250, 250, 250, 250]       # ceil(1000/300) = 4 near-equal chunks
    assert [len(c) for c in lc.x_list] == [len(c) for c in o.split_chunks(1000, 300)]
    assert np.array_equal(np.concatenate(lc.x_list), lc.x) and lc.is_split
    lc2 = LightCurve(t, np.sin(t), np.full(1000, 1e-5), chunksize=None)
    assert lc2.x_list is None and lc2.y_list is None and not lc2.is_split
    lc3 = LightCurve(t[:7], t[:7], t[:7], chunksize=3)
    assert [len(c) for c in lc3.x_list] == [3, 2, 2]


def test_lightcurve_subsample_is_sorted_and_seeded():
    t = np.arange(3000) * KEPLER_CADENCE
    a = LightCurve(t, t, t, sub=None)
    a.subsample(30, seed=4)
    b = LightCurve(t, t, t, sub=None)
    b.subsample(30, seed=4)
    assert len(a.x) == 100 and np.all(np.diff(a.x) > 0) and np.array_equal(a.x, b.x)
    rs = np.random.RandomState(4)                                     # gprot/lc.py:411-416 draw order
    assert np.array_equal(a.x, t[np.sort(rs.choice(3000, 100, replace=False))])


def test_synthetic_lightcurve_shape_and_determinism():
    t, y, e, th = synthetic_lightcurve(5, 512)
    t2, y2, _, _ = synthetic_lightcurve(5, 512)
    assert np.array_equal(t, t2) and np.array_equal(y, y2)
    assert len(t) == 512 and np.all(np.diff(t) > 0) and np.all(e == 1e-5)
    k = np.rint(t / KEPLER_CADENCE)
    assert np.allclose(k * KEPLER_CADENCE, t, rtol=0, atol=1e-9)      # on the Kepler long-cadence grid
    assert np.isfinite(o.lnprior(th)) and np.isfinite(o.lnlike_function(th, t, y, e))
    th_, yh = synthetic_lightcurve(6, 300, kind="harmonic")[1:3], None
    assert np.all(np.isfinite(th_[0]))


def test_shard_bounds_partition():
    for n in (0, 1, 7, 64, 1000, 1024):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


class _StubModel(object):
    """lnprior from the oracle; 'likelihood' = a cheap deterministic function, to test plumbing without a GPU."""

    def lnprior(self, th):
        return o.lnprior(th)

    def lnlike(self, th):
        return -float(np.sum(np.asarray(th) ** 2))

    def lnpost_batch(self, thetas, return_parts=False):
        lp = np.array([o.lnprior(t) for t in thetas])
        ll = np.array([-float(np.sum(t ** 2)) if np.isfinite(p) else -np.inf for t, p in zip(thetas, lp)])
        return (lp, ll) if return_parts else lp + ll


def test_walker_adapter_scalar_and_pooled_agree():
    walker = Emcee3Model(_StubModel())
    th = o.sample_from_prior(10, seed=2)
    th[3, 0] = 5.0                                                    # outside the support
    one = [walker(State(t)) for t in th]
    many = EnsemblePool().map(walker, [State(t) for t in th])
    for a, b in zip(one, many):
        assert a.log_prior == b.log_prior and a.log_likelihood == b.log_likelihood
    assert one[3].log_prior == -np.inf and one[3].log_likelihood == -np.inf and one[3].log_probability == -np.inf
    assert EnsemblePool().map(lambda v: v + 1, [1, 2]) == [2, 3]


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    th = o.sample_from_prior(13, seed=3)                              # ragged: 7 + 6
    ids = np.arange(13) % 3

    def ev(block, idb):
        return block.sum(axis=1) + 100.0 * idb                        # stand-in for the device launch

    full = sharded_lnlike(ev, th, lc_id=ids, rank=rank, world=world, gather=True)
    local = sharded_lnlike(ev, th, lc_id=ids, rank=rank, world=world, gather=False)
    q.put((rank, full, local))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_lnlike_two_ranks_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    th = o.sample_from_prior(13, seed=3)
    want = th.sum(axis=1) + 100.0 * (np.arange(13) % 3)
    for rank, full, local in res:
        assert np.allclose(full, want, rtol=0, atol=0)
        lo, hi = shard_bounds(13, rank, 2)
        assert np.array_equal(local, want[lo:hi])


def _gather_worker(rank, world, port, q, nitems):
    import torch
    import torch.distributed as dist

    from gprotation_b200.dist import all_gather_rows, gather_width, ordered_rows

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(nitems, rank, world)
    local = torch.full((gather_width(nitems, world),), float("nan"), dtype=torch.float64)
    local[: hi - lo] = torch.arange(lo, hi, dtype=torch.float64) * 1.5 - 7.0          # this rank's block of "log-likelihoods"
    out = all_gather_rows(local, nitems, world)
    q.put((rank, ordered_rows(out, nitems, world)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nitems", [256, 13, 2])
def test_strong_scaling_gather_two_ranks_gloo(nitems):
    """bench.py extra.config3 (one star's ensemble split over the ranks): the block split, the padded all-gather and the
    re-ordering, world_size 2 over gloo -- uneven blocks and fewer rows than a full block included."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + nitems % 7
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q, nitems)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(nitems) * 1.5 - 7.0
    for _, full in res:
        assert np.array_equal(full, want)
