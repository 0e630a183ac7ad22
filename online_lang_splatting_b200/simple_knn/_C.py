"""``distCUDA2(points) -> mean squared distance to the 3 nearest neighbours`` (submodules/simple-knn/spatial.cu:15-27),
through ``ols_knn_mean_dist2`` of the C ABI.  No CPU path: a CPU tensor raises like the reference's CUDA-only op."""
from __future__ import annotations

import torch

from .. import _native as N


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    N.require_cuda()
    if not points.is_cuda:
        raise RuntimeError("distCUDA2 needs a CUDA tensor")
    if points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    pts = points.detach().to(torch.float32).contiguous()
    P = pts.shape[0]
    out = torch.zeros((P,), dtype=torch.float32, device=pts.device)   # reference: torch::full({P}, 0.0)
    if P == 0:
        return out
    lib = N.lib()
    nbytes = lib.ols_knn_workspace_size(P)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=pts.device)
    with torch.cuda.device(pts.device):
        stream = torch.cuda.current_stream(pts.device).cuda_stream
        N.check(lib.ols_knn_mean_dist2(P, pts.data_ptr(), out.data_ptr(), ws.data_ptr(), nbytes, stream))
    return out
