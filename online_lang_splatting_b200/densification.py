"""Densification bookkeeping of the mapping loop over the flat per-Gaussian arrays.

* ``update_stats``   -- per view, right after ``loss.backward()``: ``max_radii2D`` and ``add_densification_stats``
  (utils/slam_backend.py:417-428,719-728; gaussian_splatting/scene/gaussian_model.py:965-969) in one kernel with no
  boolean-mask indexing (each ``x[mask] = ...`` of the reference is a ``nonzero()`` + host synchronisation).
* ``densify_flags``  -- the clone / split / prune selection masks of ``densify_and_prune`` (gaussian_model.py:948-963)
  evaluated in one pass on the current state, with their counts.

The gather / concatenate / optimizer-state surgery that follows (``densification_postfix``, ``prune_points``) is
control-plane code and stays with the caller.  CUDA tensors only; there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _native as N
import ctypes as C

CLONE, SPLIT, PRUNE = 1, 2, 4


def _check(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous CUDA tensor of dtype {dtype} (updated in place / read raw)")
    return t


def update_stats(radii: torch.Tensor, viewspace_grad: Optional[torch.Tensor], max_radii2D: torch.Tensor,
                 xyz_gradient_accum: Optional[torch.Tensor] = None, denom: Optional[torch.Tensor] = None) -> None:
    """In place, for the Gaussians with ``radii > 0``: ``max_radii2D = max(max_radii2D, radii)`` and, when
    ``viewspace_grad`` ([P,3], ``viewspace_point_tensor.grad``) is given, ``xyz_gradient_accum += |grad[:, :2]|``,
    ``denom += 1``."""
    N.require_cuda()
    P = int(radii.shape[0])
    _check(radii, torch.int32, "radii"); _check(max_radii2D, torch.float32, "max_radii2D")
    if viewspace_grad is not None:
        _check(viewspace_grad, torch.float32, "viewspace_grad")
        _check(xyz_gradient_accum, torch.float32, "xyz_gradient_accum"); _check(denom, torch.float32, "denom")
        if viewspace_grad.shape != (P, 3) or xyz_gradient_accum.numel() != P or denom.numel() != P:
            raise RuntimeError("shape mismatch")
    if max_radii2D.numel() != P:
        raise RuntimeError("shape mismatch")
    dev = radii.device
    with torch.cuda.device(dev):
        N.check(N.lib().ols_densify_stats(P, radii.data_ptr(), N.ptr(viewspace_grad), max_radii2D.data_ptr(),
                                          N.ptr(xyz_gradient_accum), N.ptr(denom), torch.cuda.current_stream(dev).cuda_stream))


def densify_flags(xyz_gradient_accum: torch.Tensor, denom: torch.Tensor, scaling_raw: torch.Tensor, opacity_raw: torch.Tensor,
                  max_radii2D: torch.Tensor, *, max_grad: float, min_opacity: float, extent: float,
                  max_screen_size: Optional[float], percent_dense: float = 0.01) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (flags uint8 [P] with bits CLONE | SPLIT | PRUNE, counts int32 [3]).  ``scaling_raw`` / ``opacity_raw`` are the
    stored parameters (log-scales [P,1|3], opacity logits [P,1]).

    All three masks are evaluated on the state BEFORE densification, in one pass.  The reference evaluates them in
    sequence (gaussian_model.py:948-963): clone, then split on the set that already holds the clones (whose padded
    gradients are zero, so no clone is split), then prune on the set after clone + split (so the new points are
    tested and the removed split parents are not).  A caller that reproduces ``densify_and_prune`` must therefore apply
    CLONE first, SPLIT second (removing the parents), and evaluate the prune criterion again for the appended points;
    PRUNE as returned here is exact for the points that survive unchanged."""
    N.require_cuda()
    P = int(opacity_raw.shape[0])
    for t, n in ((xyz_gradient_accum, "xyz_gradient_accum"), (denom, "denom"), (scaling_raw, "scaling"),
                 (opacity_raw, "opacity"), (max_radii2D, "max_radii2D")):
        _check(t.detach(), torch.float32, n)
    cols = int(scaling_raw.shape[1])
    dev = opacity_raw.device
    flags = torch.empty(P, dtype=torch.uint8, device=dev)
    counts = torch.empty(3, dtype=torch.int32, device=dev)
    prm = N.DensifyParams(max_grad=max_grad, min_opacity=min_opacity, extent=extent,
                          max_screen_size=float(max_screen_size) if max_screen_size else 0.0, percent_dense=percent_dense)
    with torch.cuda.device(dev):
        N.check(N.lib().ols_densify_flags(P, cols, xyz_gradient_accum.data_ptr(), denom.data_ptr(), scaling_raw.data_ptr(),
                                          opacity_raw.data_ptr(), max_radii2D.data_ptr(), C.byref(prm), flags.data_ptr(),
                                          counts.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    return flags, counts


# ---- plain-torch restatements of the reference lines (test references) ----------------------------------------------
