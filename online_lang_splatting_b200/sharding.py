"""Frame-parallel multi-GPU plumbing: one process per GPU, Gaussian parameters
replicated, the keyframes of a mapping iteration sharded over the ranks, and ONE all-reduce of a flat
fp32 gradient buffer per optimiser step.  The reference has no multi-GPU code; the data-parallel
axis is the loop over window keyframes of utils/slam_backend.py:510-670, whose per-view gradients are
summed by autograd before a single backward.

The flat buffer is laid out structure-of-arrays so every parameter group is one contiguous slice
(a valid output tensor of ols_lang_backward) and the whole thing is a single NCCL call:

    [ rotation 4P | xyz 3P | sh 3M*P (rows [f_dc 3 | f_rest 3(M-1)] per Gaussian) | opacity P | scaling 3P | language F*P ]

(rotations first so that their rows of 4 are 16-byte aligned for every P).  The buffer holds gradients with respect to
the ACTIVATED rasterizer inputs -- sigmoid(opacity), exp(scaling), normalised rotation -- exactly as
``ols_lang_backward`` produces them.  ``FlatParams`` keeps the raw parameters in the same layout and
``FlatGradBuffer.adam_groups`` describes the buffer to ``optim.FlatAdam`` with the activation of every slice, so the
optimiser applies the activations' Jacobians itself (the reference leaves that to autograd) and f_dc / f_rest keep
their separate learning rates (gaussian_model.py:404-413).
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

    def __init__(self, P: int, F: int, M: int = 1, device="cuda", extra: int = 0):
        """``extra`` floats are appended to the buffer (``self.extra``) so that other per-step sums -- the densification
        statistics -- travel in the same all-reduce; ``self.grads`` is the gradient part alone (what FlatAdam consumes)."""
        self.P, self.F, self.M = P, F, M
        groups: Sequence[Tuple[str, Tuple[int, ...]]] = (
            ("rotations", (P, 4)), ("means3D", (P, 3)), ("sh", (P, M, 3)), ("opacity", (P, 1)), ("scales", (P, 3)),
            ("language", (P, F)))
        self.groups = groups
        self.floats_per_gaussian = 3 + 3 * M + 1 + 3 + 4 + F
        n_grad = self.floats_per_gaussian * P
        self.flat = torch.zeros(n_grad + int(extra), dtype=torch.float32, device=device)
        self.grads = self.flat[:n_grad]
        self.extra = self.flat[n_grad:]
        self.views: Dict[str, torch.Tensor] = {}
        o = 0
        for name, shape in groups:
            n = math.prod(shape)
            self.views[name] = self.flat[o:o + n].view(shape)
            o += n
        assert o == n_grad

    def zero_(self):
        self.flat.zero_()
        return self

    def backward_outputs(self, scratch: Dict[str, torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """`out=` dictionary for diff_gaussian_rasterization._backward_native(..., accumulate=True)."""
        out = dict(self.views)
        if scratch:
            out.update(scratch)
        return out

    def adam_groups(self, lr: Dict[str, float]):
        """Group list for ``optim.FlatAdam`` over this layout.  ``lr`` maps the reference's group names (xyz, f_dc,
        f_rest, opacity, scaling, rotation, f_language; gaussian_model.py:393-437) to learning rates."""
        from . import _native as N
        P, F, M = self.P, self.F, self.M
        return [("rotation", 4 * P, lr["rotation"], N.ACT_NORMALIZE4),
                ("xyz", 3 * P, lr["xyz"], N.ACT_NONE),
                ("f_dc+f_rest", 3 * M * P, lr["f_dc"], N.ACT_NONE, 3 * M if M > 1 else 0, 3, lr.get("f_rest", lr["f_dc"] / 20.0)),
                ("opacity", P, lr["opacity"], N.ACT_SIGMOID),
                ("scaling", 3 * P, lr["scaling"], N.ACT_EXP),
                ("f_language", F * P, lr["f_language"], N.ACT_NONE)]

    def all_reduce(self):
        """Sum over ranks (no-op without an initialised process group)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat)
        return self

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4


class FlatParams(FlatGradBuffer):
    """The raw (pre-activation) Gaussian parameters in the layout of ``FlatGradBuffer`` plus a second buffer with the
    activated values the rasterizer reads; ``activate()`` refreshes the latter after an optimiser step (one kernel)."""

    def __init__(self, raw: Dict[str, torch.Tensor], F: int, M: int = 1, device="cuda"):
        P = raw["means3D"].shape[0]
        super().__init__(P, F, M, device=device)
        for name, _ in self.groups:
            self.views[name].copy_(raw[name].reshape(self.views[name].shape))
        self.act_flat = self.flat.clone()
        self.act: Dict[str, torch.Tensor] = {}
        o = 0
        for name, shape in self.groups:
            n = math.prod(shape)
            self.act[name] = self.act_flat[o:o + n].view(shape)
            o += n

    def activate(self):
        from .optim import activate_params
        for name in ("means3D", "sh", "language"):      # identity activations: the rasterizer reads the raw slices
            self.act[name] = self.views[name]
        activate_params(self.views["opacity"], self.views["scales"], self.views["rotations"], self.act["opacity"],
                        self.act["scales"], self.act["rotations"])
        return self.act


class SideStats:
    """Densification statistics of a sharded mapping iteration: every rank folds its own views into per-step deltas
    (``densification.update_stats``), the deltas are reduced over the ranks -- SUM for ``xyz_gradient_accum`` and
    ``denom``, MAX for ``max_radii2D`` (gaussian_model.py:965-969, utils/slam_backend.py:676-680) -- and added to the
    replicated running state, so all ranks take identical densification decisions."""

    def __init__(self, P: int, device="cuda", delta: torch.Tensor = None):
        """``delta``: optional [2 P] view (e.g. ``FlatGradBuffer.extra``) so that the per-step sums are reduced together with
        the gradients; the owner of that buffer then does the all-reduce."""
        self.P = P
        self.shared_delta = delta is not None
        self.delta = delta if delta is not None else torch.zeros(2 * P, dtype=torch.float32, device=device)  # [accum P | denom P], one SUM
        self.delta_max = torch.zeros(P, dtype=torch.float32, device=device)        # one MAX
        self.xyz_gradient_accum = torch.zeros(P, 1, dtype=torch.float32, device=device)
        self.denom = torch.zeros(P, 1, dtype=torch.float32, device=device)
        self.max_radii2D = torch.zeros(P, dtype=torch.float32, device=device)

    def begin_step(self):
        self.delta.zero_()
        self.delta_max.zero_()

    def add_view(self, radii: torch.Tensor, viewspace_grad: torch.Tensor):
        from .densification import update_stats
        P = self.P
        update_stats(radii, viewspace_grad, self.delta_max, self.delta[:P].view(P, 1), self.delta[P:].view(P, 1))

    def fused_outputs(self):
        """``out["stats"]`` for ``_backward_native_batch``: the backward's geometry kernel folds all views of the call into
        the per-step deltas itself (no separate launches)."""
        P = self.P
        return (self.delta_max, self.delta[:P], self.delta[P:])

    def all_reduce(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            if not self.shared_delta:
                dist.all_reduce(self.delta)
            dist.all_reduce(self.delta_max, op=dist.ReduceOp.MAX)
        return self

    def reduce_max_radii(self):
        """Cross-rank MAX of the running ``max_radii2D``.  A running maximum commutes with the reduction, so this is only
        needed when densification reads the statistic (every ``gaussian_update_every`` iterations), not every step."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.max_radii2D, op=dist.ReduceOp.MAX)
        return self

    def apply(self):
        P = self.P
        self.xyz_gradient_accum += self.delta[:P].view(P, 1)
        self.denom += self.delta[P:].view(P, 1)
        torch.maximum(self.max_radii2D, self.delta_max, out=self.max_radii2D)
