"""BASELINE.json configs[2] and configs[3] across the GPUs of one box (one process per GPU under torchrun).

  configs[3]  long-baseline stress: 1 star, N=8192, 256 walker proposals; the ensemble of ONE star spans the GPUs
              (strong scaling: 256 / world rows per GPU, light curve replicated, NCCL all-gather of the log-likelihoods)
  configs[2]  multi-star sweep: S stars x 64 walkers, N=2048, star-major rows block-sharded over the GPUs
              (--stars 1000 is the full configuration; the default is smaller so that the run takes seconds)

usage: python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/bench_configs.py [--stars 64] [--reps 3]
Times on the device (CUDA events around the local evaluation), max over ranks; rank 0 prints one JSON line per config."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gprotation_b200 import GPEngine  # noqa: E402
from gprotation_b200.dist import shard_bounds, sharded_lnlike  # noqa: E402
from gprotation_b200.synth import synthetic_lightcurve  # noqa: E402
from oracle import gp_oracle as o  # noqa: E402  (prior sampler + spot check only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stars", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = GPEngine(local)
    peak, _ = eng.measure_dmma_peak()

    def timed(fn):
        best = None
        out = None
        for _ in range(a.reps + 1):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            best = float(dt) if best is None else min(best, float(dt))
        return best, out

    # ---- configs[3]: one star's ensemble spans the GPUs
    n, B = 8192, 256
    t, y, yerr, _ = synthetic_lightcurve(1, n, kind="harmonic")
    lc = eng.register_lc(t, y, yerr)
    th = o.sample_from_prior(B, seed=9)
    sec, full = timed(lambda: sharded_lnlike(lambda tb, ids: eng.lnlike_batch(tb, lc_id=lc), th, rank=rank, world=world,
                                             gather=True, device="cuda"))
    if rank == 0:
        F = o.flops_per_eval(n)
        ref = o.lnlike_function(th[0], t, y, yerr)
        print(json.dumps({"config": "configs[3]: 1 star, N=8192, 256 walkers, ensemble split over the GPUs + NCCL all-gather",
                          "n_gpus": world, "scaling": "strong", "value": B / sec, "unit": "evals/s", "seconds": sec,
                          "rows_per_gpu": shard_bounds(B, 0, world)[1], "ctas_per_problem": eng.last_cluster(),
                          "tflops": B * F / sec / 1e12, "frac_of_dmma_peak_all_gpus": B * F / sec / 1e12 / (peak * world),
                          "rel_err_row0_vs_oracle": abs(full[0] - ref) / abs(ref)}), flush=True)

    # ---- configs[2]: many stars x 64 walkers, rows star-major
    eng.unregister_all()
    S, W, n = a.stars, 64, 2048
    lo, hi = shard_bounds(S * W, rank, world)
    mine = range(lo // W, (hi - 1) // W + 1)                       # only the stars whose rows live here are uploaded
    ids = {}
    for s in mine:
        t, y, yerr, _ = synthetic_lightcurve(s, n, kind="harmonic")
        ids[s] = eng.register_lc(t, y, yerr)
    th = o.sample_from_prior(S * W, seed=8)
    rows = np.array([ids.get(i // W, 0) for i in range(S * W)], dtype=np.int32)
    sec, _ = timed(lambda: sharded_lnlike(lambda tb, idb: eng.lnlike_batch(tb, lc_id=idb), th, lc_id=rows, rank=rank, world=world,
                                          gather=False))
    if rank == 0:
        F = o.flops_per_eval(n)
        print(json.dumps({"config": "configs[2]: %d stars x 64 walkers, N=2048, star-major rows sharded over the GPUs, no gather" % S,
                          "n_gpus": world, "scaling": "strong", "value": S * W / sec, "unit": "evals/s", "seconds": sec,
                          "tflops": S * W * F / sec / 1e12, "frac_of_dmma_peak_all_gpus": S * W * F / sec / 1e12 / (peak * world)}),
              flush=True)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
