"""Multi-process plumbing (one process per GPU, torch.distributed).

Round 1 runs independent row partitions per rank (no data-path collective); the helpers here are the
host-side pieces the sharded build will also use: the reference's own ownership rule for vertex
ranges (utils.py:259-264: block = N // T + 1) and device-timed max-over-ranks reductions.
"""
import numpy as np


def owner_ranges(n_vertices, world_size):
    """Contiguous vertex ranges exactly as the reference partitions writers (utils.py:259-264)."""
    block = n_vertices // world_size + 1
    return [(r * block, min((r + 1) * block, n_vertices)) for r in range(world_size)]


def owner_of(vertex, n_vertices, world_size):
    return np.asarray(vertex) // (n_vertices // world_size + 1)


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (timing contract: the slowest rank defines the step)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])
