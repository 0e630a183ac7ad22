"""Test/debug helpers: typed views into the opaque rasterizer workspace (never on the hot path)."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _native as N


def workspace_arrays(state) -> Dict[str, torch.Tensor]:
    """Decode the workspace kept by a forward call (``ctx.state`` of the autograd function).
    Counterpart of reading geomBuffer/binningBuffer/imgBuffer of the reference."""
    a = state.args
    ws = state.keep["workspace"]
    v = N.WsView()
    N.check(N.lib().ols_lang_workspace_view(a.P, a.F, a.W, a.H, a.tile, a.R_cap, ws.data_ptr(), C.byref(v)))
    return _decode(ws, v, a.P, a.W, a.H, max(state.R, 0), has_rgb=True)


def workspace_arrays_dis(state):
    """(colour list, language list) views of a disentangled forward's workspace."""
    d = state.args
    a = d.base
    ws = state.keep["workspace"]
    vc, vl = N.WsView(), N.WsView()
    N.check(N.lib().ols_dis_workspace_view(C.byref(d), C.byref(vc), C.byref(vl)))
    return (_decode(ws, vc, a.P, a.W, a.H, max(state.R, 0), has_rgb=True),
            _decode(ws, vl, a.P, a.W, a.H, max(state.R_lang, 0), has_rgb=False))


def _decode(ws, v, P, W, H, R, has_rgb):
    base = ws.data_ptr()

    def view(p, nbytes, dtype, shape):
        off = p - base
        return ws[off:off + nbytes].view(dtype).view(*shape)

    HW, rec, T = W * H, v.rec_floats, v.n_tiles
    out = {
        "records": view(v.d_records, 4 * rec * P, torch.float32, (P, rec)),
        "cov3D": view(v.d_cov3D, 24 * P, torch.float32, (P, 6)),
        "clamped": view(v.d_clamped, 4 * P, torch.uint8, (P, 4))[:, :3],
        "tiles_touched": view(v.d_tiles_touched, 4 * P, torch.int32, (P,)),
        "ranges": view(v.d_ranges, 8 * T, torch.int32, (T, 2)),
        "point_list": view(v.d_point_list, 4 * R, torch.int32, (R,)) if R else torch.empty(0, dtype=torch.int32),
        "keys": view(v.d_keys, 8 * R, torch.int64, (R,)) if R else torch.empty(0, dtype=torch.int64),
        "final_T": view(v.d_final_T, 4 * HW, torch.float32, (H, W)),
        "n_contrib": view(v.d_n_contrib, 4 * HW, torch.int32, (H, W)),
    }
    r = out["records"]
    out["means2D"] = r[:, 0:2]
    out["conic_opacity"] = torch.stack([r[:, 2], r[:, 3], r[:, 4], r[:, 5]], 1)
    out["depths"] = r[:, 7]
    if has_rgb:
        out["rgb"] = r[:, 8:11]
    return out
