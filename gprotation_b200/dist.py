"""Sharding of (star, walker) items over the GPUs of one box.

Items are independent, so the data path has no collective: every rank evaluates a contiguous block of
the flattened item list on its own GPU (light curves are replicated, <= 3*N*8 bytes per star).  Only when the
caller needs every rank to see all log-probabilities -- one star's ensemble spanning GPUs -- is there a
small all-gather of B doubles (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(nitems, rank, world):
    """Contiguous block partition; the first (nitems % world) ranks hold one extra item."""
    base, rem = divmod(int(nitems), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_lnlike(evaluate, thetas, lc_id=None, rank=0, world=1, gather=True, device=None):
    """Evaluate this rank's block with ``evaluate(thetas_block, lc_id_block) -> ndarray`` and, if ``gather``,
    all-gather the blocks so every rank returns the full [B] vector (else only the local block)."""
    thetas = np.ascontiguousarray(thetas, dtype=np.float64).reshape(-1, 5)
    B = thetas.shape[0]
    lo, hi = shard_bounds(B, rank, world)
    ids = None
    if lc_id is not None:
        ids = np.full(B, int(lc_id), dtype=np.int32) if np.isscalar(lc_id) else np.asarray(lc_id, dtype=np.int32)
        ids = ids[lo:hi]
    local = np.asarray(evaluate(thetas[lo:hi], ids), dtype=np.float64) if hi > lo else np.empty(0)
    if world == 1 or not gather:
        return local
    import torch
    import torch.distributed as dist

    dev = torch.device(device) if device is not None else torch.device("cpu")
    width = -(-B // world)
    send = torch.full((width,), float("nan"), dtype=torch.float64, device=dev)
    send[: hi - lo] = torch.from_numpy(local).to(dev)
    recv = [torch.empty(width, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(recv, send)
    out = np.empty(B)
    for r in range(world):
        a, b = shard_bounds(B, r, world)
        out[a:b] = recv[r][: b - a].cpu().numpy()
    return out
