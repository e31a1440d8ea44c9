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


def gather_width(nitems, world):
    """Rows per rank of the padded all-gather buffer (the largest block of shard_bounds)."""
    return -(-int(nitems) // int(world))


def all_gather_rows(local, nitems, world, out=None):
    """All-gather of the per-rank blocks of a block-sharded vector of ``nitems`` rows (torch tensors on the device of
    the process group: NCCL on the GPU box, gloo on CPU).  ``local`` holds this rank's block in its first rows, padded to
    ``gather_width``; returns the padded [world * width] buffer (``ordered_rows`` puts the rows back in order).  This
    is the one collective of the path: the log-probabilities of a single star's ensemble that spans GPUs
    (scripts/gprot-fit:127-131,195: the reference gathers them through pool.map)."""
    import torch
    import torch.distributed as dist

    width = gather_width(nitems, world)
    if out is None:
        out = torch.empty(world * width, dtype=local.dtype, device=local.device)
    if world == 1:
        out.copy_(local)
        return out
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(out, local)
    else:
        dist.all_gather(list(out.view(world, width).unbind(0)), local)
    return out


def ordered_rows(gathered, nitems, world):
    """The ``nitems`` rows of an ``all_gather_rows`` buffer, in their original order (NumPy array)."""
    width = gather_width(nitems, world)
    g = gathered.detach().cpu().numpy().reshape(world, width)
    return np.concatenate([g[r][: shard_bounds(nitems, r, world)[1] - shard_bounds(nitems, r, world)[0]] for r in range(world)])
