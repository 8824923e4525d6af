"""Utterance sharding for the multi-GPU path (one process per GPU, torch.distributed plumbing).

Utterances are independent units: rank r analyses a contiguous range of the batch with
no data-path collective, then ONE all-gather per output array reassembles the batch on every rank.
Ranges are balanced by frame count (equal utterance counts when lengths are uniform)."""
from __future__ import annotations


def shard_ranges(frame_counts, world: int):
    """Contiguous [begin, end) utterance ranges, one per rank, balancing sum(frame_counts)."""
    n = len(frame_counts)
    total = sum(frame_counts)
    bounds = [0]
    acc = 0
    u = 0
    for r in range(1, world):
        target = total * r / world
        while u < n and acc + frame_counts[u] / 2 < target:
            acc += frame_counts[u]
            u += 1
        bounds.append(u)
    bounds.append(n)
    for r in range(1, len(bounds)):  # never hand a rank a negative range
        bounds[r] = max(bounds[r], bounds[r - 1])
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def all_gather_rows(dist, local, counts, out=None):
    """All-gather per-rank row blocks `local` ([counts[rank], ...]) into one [sum(counts), ...] array.
    Equal counts use all_gather_into_tensor (in place when `local` is a slice of `out`); ragged counts
    fall back to one broadcast per rank."""
    import torch
    world = dist.get_world_size()
    rank = dist.get_rank()
    total = sum(counts)
    if out is None:
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if len(set(counts)) == 1:
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    off = 0
    for r in range(world):
        blk = out[off:off + counts[r]]
        if r == rank:
            blk.copy_(local)
        dist.broadcast(blk, src=r)
        off += counts[r]
    return out
