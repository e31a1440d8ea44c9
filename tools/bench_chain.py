"""BASELINE.json configs[4]: end-to-end chain (5000 steps x 64 walkers, KDE/DE/snooker mix) on a synthetic KOI-like light
curve through GPRotModel -> fit.Emcee3Model -> sampler.Ensemble / Sampler on the GPU, wall clock; beside it the same
sampler driving the CPU oracle (dense restatement of the george path) for a bounded number of steps, extrapolated.

usage: python tools/bench_chain.py [--steps 5000] [--walkers 64] [--n 2048] [--chunksize 300 | --nochunks] [--cpu-steps 3]
Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gprotation_b200 import sampler as em  # noqa: E402
from gprotation_b200.fit import Emcee3Model, EnsemblePool  # noqa: E402
from gprotation_b200.lc import LightCurve  # noqa: E402
from gprotation_b200.model import GPRotModel  # noqa: E402
from gprotation_b200.synth import synthetic_lightcurve  # noqa: E402


class OracleModel(object):
    """The reference's per-walker CPU path: lnprior, then lnlike = sum over chunks (gprot/model.py:358-369)."""

    def __init__(self, t, y, yerr, chunks):
        self.t, self.y, self.yerr, self.chunks = t, y, yerr, chunks

    def lnprior(self, th):
        from oracle import gp_oracle as o

        return o.lnprior(th)

    def lnlike(self, th):
        from oracle import gp_oracle as o

        return o.lnlike(th, self.t, self.y, self.yerr, self.chunks)


def _one_blas_thread():
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(1)
    except Exception:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--walkers", type=int, default=64)
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--chunksize", type=int, default=300)
    ap.add_argument("--nochunks", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=6)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--native", action="store_true", help="drive the chain from the host side of the C ABI (Sampler.run_native)")
    a = ap.parse_args()
    cs = None if a.nochunks else a.chunksize
    t, y, yerr, truth = synthetic_lightcurve(a.seed, a.n, kind="harmonic")
    lc = LightCurve(t, y, yerr, name="koi_like", chunksize=cs)
    mod = GPRotModel(lc)
    moves = [(em.KDEMove(), 0.4), (em.DEMove(1.0), 0.4), (em.DESnookerMove(), 0.2)]
    start = mod.sample_from_prior(a.walkers, seed=a.seed)

    ens = em.Ensemble(Emcee3Model(mod), start, pool=EnsemblePool(), random=np.random.RandomState(a.seed))
    sam = em.Sampler(moves)
    sam.run(ens, 5, native=a.native)                             # warm-up: allocations
    sam.reset()
    l0 = mod.engine.launch_count()
    t0 = time.perf_counter()
    sam.run(ens, a.steps, native=a.native)
    mod.engine.sync()
    gpu_s = time.perf_counter() - t0
    launches = mod.engine.launch_count() - l0
    nchunk = 1 if lc.x_list is None else len(lc.x_list)

    # CPU: identical sampler; the walkers of every proposal set are mapped over a process pool with one single-threaded
    # worker per host core -- the reference's --ncores layout (schwimmbad MultiPool, scripts/gprot-fit:127-129,195)
    import multiprocessing as mp

    from oracle import gp_oracle as o

    chunks = o.split_chunks(a.n, cs)
    cpu_model = OracleModel(t, y, yerr, chunks)
    cores = len(os.sched_getaffinity(0))

    class ProcPool(object):
        def __init__(self, n):
            self.pool = mp.get_context("fork").Pool(n, initializer=_one_blas_thread)

        def map(self, func, iterable):
            return self.pool.map(func, list(iterable), chunksize=1)

    pool = ProcPool(cores)
    ens_c = em.Ensemble(Emcee3Model(cpu_model), start, pool=pool, random=np.random.RandomState(a.seed))
    sam_c = em.Sampler(moves)
    t0 = time.perf_counter()
    sam_c.run(ens_c, a.cpu_steps)
    cpu_s_per_step = (time.perf_counter() - t0) / a.cpu_steps
    pool.pool.terminate()
    line = {
        "metric": "end-to-end emcee3-style chain wall clock", "unit": "s", "higher_is_better": False,
        "config": {"workload": "configs[4]: %d steps x %d walkers, N=%d, %s" % (a.steps, a.walkers, a.n,
                   "one N x N problem (--nochunks)" if cs is None else "%d chunks of <= %d points" % (nchunk, cs)),
                   "driver": "gprot_chain_run (host C++ above the likelihood ABI)" if a.native else "gprotation_b200/sampler.py (Python)"},
        "value": gpu_s, "ms_per_step": 1e3 * gpu_s / a.steps, "gpu_launches": int(launches),
        "likelihood_evals": a.steps * a.walkers, "acceptance": float(sam.acceptance_fraction.mean()),
        "cpu_baseline": {"kind": "port", "cores": cores, "value": cpu_s_per_step * a.steps, "unit": "s",
                         "sample": "%d steps timed with the same sampler mapping the walkers of each proposal set over %d single-threaded "
                                   "oracle processes (%.3f s per step), extrapolated to %d steps" % (a.cpu_steps, cores, cpu_s_per_step, a.steps)},
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
