"""Frame-parallel multi-GPU plumbing: one process per GPU, Gaussian parameters
replicated, the keyframes of a mapping iteration sharded over the ranks, and ONE all-reduce of a flat
fp32 gradient buffer per optimiser step.  The reference has no multi-GPU code; the data-parallel
axis is the loop over window keyframes of utils/slam_backend.py:510-670, whose per-view gradients are
summed by autograd before a single backward.

The flat buffer is laid out structure-of-arrays so every parameter group is one contiguous slice
(a valid output tensor of ols_lang_backward) and the whole thing is a single NCCL call:

    [ xyz 3P | f_dc 3P | f_rest 3(M-1)P | opacity P | scaling 3P | rotation 4P | language F*P ]
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment of view indices to ranks (every view exactly once, balanced to +-1)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_views, world))


def balanced_views(costs: Sequence[float], world: int) -> List[List[int]]:
    """Assign views to ranks so that the per-rank sums of ``costs`` are as even as possible, every rank getting the
    same number of views (``len(costs)`` must be a multiple of ``world``).  Greedy longest-processing-time-first:
    views in order of decreasing cost, each to the least loaded rank that still has room.  Views differ in cost by
    up to ~1.7x (how much of a tile's list is traversed before the pixels saturate), and a step ends when the
    slowest rank does, so contiguous or round-robin blocks leave the other GPUs waiting."""
    n = len(costs)
    if world <= 0 or n % world != 0:
        raise ValueError(f"{n} views cannot be split evenly over {world} ranks")
    per = n // world
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for v in sorted(range(n), key=lambda i: (-float(costs[i]), i)):
        r = min((r for r in range(world) if len(out[r]) < per), key=lambda r: (load[r], r))
        out[r].append(v)
        load[r] += float(costs[v])
    return [sorted(x) for x in out]


class FlatGradBuffer:
    """One contiguous fp32 buffer holding every per-Gaussian parameter gradient of the rasterizer."""

    def __init__(self, P: int, F: int, M: int = 1, device="cuda"):
        self.P, self.F, self.M = P, F, M
        groups: Sequence[Tuple[str, Tuple[int, ...]]] = (
            ("means3D", (P, 3)), ("sh", (P, M, 3)), ("opacity", (P, 1)), ("scales", (P, 3)), ("rotations", (P, 4)),
            ("language", (P, F)))
        self.floats_per_gaussian = 3 + 3 * M + 1 + 3 + 4 + F
        self.flat = torch.zeros(self.floats_per_gaussian * P, dtype=torch.float32, device=device)
        self.views: Dict[str, torch.Tensor] = {}
        o = 0
        for name, shape in groups:
            n = math.prod(shape)
            self.views[name] = self.flat[o:o + n].view(shape)
            o += n
        assert o == self.flat.numel()

    def zero_(self):
        self.flat.zero_()
        return self

    def backward_outputs(self, scratch: Dict[str, torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """`out=` dictionary for diff_gaussian_rasterization._backward_native(..., accumulate=True)."""
        out = dict(self.views)
        if scratch:
            out.update(scratch)
        return out

    def all_reduce(self):
        """Sum over ranks (no-op without an initialised process group)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat)
        return self

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4
