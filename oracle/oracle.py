"""ctypes front-end of the CPU restatement (oracle/ols_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under online_lang_splatting_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

fp = C.POINTER(C.c_float)


class Scene(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("P", "sh_degree", "M", "F", "W", "H", "tile", "prefiltered")] + \
               [(n, C.c_float) for n in ("tanfovx", "tanfovy", "scale_modifier", "_pad")] + \
               [(n, C.c_void_p) for n in ("bg", "means3D", "shs", "colors_precomp", "language", "opacities", "scales",
                                          "rotations", "cov3D_precomp", "viewmatrix", "projmatrix", "projmatrix_raw",
                                          "campos")]


class Geom(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("depths", "radii", "means2D", "cov3D", "conic_opacity", "rgb", "clamped",
                                          "tiles_touched", "point_offsets")]


class Bin(C.Structure):
    _fields_ = [("R", C.c_int64)] + [(n, C.c_void_p) for n in ("keys_sorted", "point_list", "ranges", "final_T",
                                                                "n_contrib")]


class Image(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("color", "language", "depth", "opacity", "n_touched")]


class Grads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dL_dcolor", "dL_dlanguage", "dL_ddepth")] + \
               [("compat", C.c_int32), ("_pad", C.c_int32)] + \
               [(n, C.c_void_p) for n in ("dL_dmeans2D", "dL_dconic", "dL_dopacity", "dL_dcolors", "dL_dlang",
                                          "dL_ddepths", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drots",
                                          "dL_dtau")]


class DisExtra(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("opacities_lang", "scales_lang", "rotations_lang", "cov3D_precomp_lang")]


class DisGeom(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("radii_lang", "cov3D_lang", "conic_opacity_lang", "tiles_touched_lang",
                                          "point_offsets_lang")]


class DisGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dL_dcolor", "dL_dlanguage", "dL_ddepth")] + \
               [("compat", C.c_int32), ("_pad", C.c_int32)] + \
               [(n, C.c_void_p) for n in ("dL_dmeans2D", "dL_dconic", "dL_dconic_lang", "dL_dopacity", "dL_dopacity_lang",
                                          "dL_dcolors", "dL_dlang", "dL_ddepths", "dL_dmeans3D", "dL_dcov3D",
                                          "dL_dcov3D_lang", "dL_dsh", "dL_dscales", "dL_dscales_lang", "dL_drots",
                                          "dL_drots_lang", "dL_dtau")]


def build(force: bool = False) -> str:
    so = os.path.join(HERE, "libols_oracle.so")
    src = os.path.join(HERE, "ols_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "libols_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ols_oracle_preprocess.restype = C.c_int64
        _LIB.ols_oracle_preprocess.argtypes = [C.POINTER(Scene), C.POINTER(Geom)]
        _LIB.ols_oracle_render.argtypes = [C.POINTER(Scene), C.POINTER(Geom), C.POINTER(Bin), C.POINTER(Image)]
        _LIB.ols_oracle_backward.argtypes = [C.POINTER(Scene), C.POINTER(Geom), C.POINTER(Bin), C.POINTER(Grads)]
        _LIB.ols_oracle_dis_preprocess.restype = C.c_int64
        _LIB.ols_oracle_dis_preprocess.argtypes = [C.POINTER(Scene), C.POINTER(DisExtra), C.POINTER(Geom),
                                                   C.POINTER(DisGeom), C.POINTER(C.c_int64)]
        _LIB.ols_oracle_dis_render.argtypes = [C.POINTER(Scene), C.POINTER(Geom), C.POINTER(DisGeom), C.POINTER(Bin),
                                               C.POINTER(Bin), C.POINTER(Image), C.POINTER(Image)]
        _LIB.ols_oracle_dis_backward.argtypes = [C.POINTER(Scene), C.POINTER(DisExtra), C.POINTER(Geom), C.POINTER(DisGeom),
                                                 C.POINTER(Bin), C.POINTER(Bin), C.POINTER(DisGrads)]
    return _LIB


def _np(x, dtype=np.float32):
    if x is None:
        return None
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=dtype)


def _ptr(a):
    return None if a is None else a.ctypes.data


def set_num_threads(n: int):
    lib().ols_oracle_set_num_threads(int(n))


def num_threads() -> int:
    return int(lib().ols_oracle_num_threads())


def reduce_lane_mask(n: int) -> np.ndarray:
    out = np.zeros(n, np.uint8)
    lib().ols_oracle_reduce_lane_mask(int(n), C.c_void_p(out.ctypes.data))
    return out


class OracleRasterizer:
    """Holds one scene + camera; forward() and backward() mirror the reference's two C++ entry points
    (rasterize_points.cu:125-241, :333-455)."""

    def __init__(self, *, means3D, opacities, language, W, H, tanfovx, tanfovy, viewmatrix, projmatrix,
                 projmatrix_raw=None, campos=None, bg=None, shs=None, colors_precomp=None, scales=None,
                 rotations=None, cov3D_precomp=None, sh_degree=0, scale_modifier=1.0, tile=15):
        self.keep = {}
        k = self.keep
        k["means3D"] = _np(means3D).reshape(-1, 3)
        P = k["means3D"].shape[0]
        k["opacities"] = _np(opacities).reshape(-1)
        k["language"] = _np(language).reshape(P, -1)
        F = k["language"].shape[1]
        k["shs"] = None if shs is None else _np(shs).reshape(P, -1, 3)
        M = 0 if k["shs"] is None else k["shs"].shape[1]
        k["colors_precomp"] = None if colors_precomp is None else _np(colors_precomp).reshape(P, 3)
        k["scales"] = None if scales is None else _np(scales).reshape(P, 3)
        k["rotations"] = None if rotations is None else _np(rotations).reshape(P, 4)
        k["cov3D_precomp"] = None if cov3D_precomp is None else _np(cov3D_precomp).reshape(P, 6)
        k["viewmatrix"] = _np(viewmatrix).reshape(16)
        k["projmatrix"] = _np(projmatrix).reshape(16)
        k["projmatrix_raw"] = _np(projmatrix_raw if projmatrix_raw is not None else projmatrix).reshape(16)
        k["campos"] = _np(campos if campos is not None else np.zeros(3)).reshape(3)
        k["bg"] = _np(bg if bg is not None else np.zeros(3)).reshape(3)
        self.P, self.F, self.M, self.W, self.H, self.tile = P, F, M, int(W), int(H), int(tile)
        self.scene = Scene(P=P, sh_degree=int(sh_degree), M=M, F=F, W=self.W, H=self.H, tile=self.tile, prefiltered=0,
                           tanfovx=float(tanfovx), tanfovy=float(tanfovy), scale_modifier=float(scale_modifier),
                           **{n: _ptr(k[n]) for n in ("bg", "means3D", "shs", "colors_precomp", "language", "opacities",
                                                      "scales", "rotations", "cov3D_precomp", "viewmatrix",
                                                      "projmatrix", "projmatrix_raw", "campos")})
        self.geom_np: Dict[str, np.ndarray] = {}
        self.bin_np: Dict[str, np.ndarray] = {}
        self.out: Dict[str, np.ndarray] = {}

    @property
    def grid(self):
        return ((self.W + self.tile - 1) // self.tile, (self.H + self.tile - 1) // self.tile)

    def forward(self) -> Dict[str, np.ndarray]:
        P, F, W, H = self.P, self.F, self.W, self.H
        g = self.geom_np = {
            "depths": np.zeros(P, np.float32), "radii": np.zeros(P, np.int32), "means2D": np.zeros((P, 2), np.float32),
            "cov3D": np.zeros((P, 6), np.float32), "conic_opacity": np.zeros((P, 4), np.float32),
            "rgb": np.zeros((P, 3), np.float32), "clamped": np.zeros((P, 3), np.uint8),
            "tiles_touched": np.zeros(P, np.uint32), "point_offsets": np.zeros(P, np.uint32)}
        self.geom = Geom(**{n: _ptr(a) for n, a in g.items()})
        R = int(lib().ols_oracle_preprocess(C.byref(self.scene), C.byref(self.geom)))
        gx, gy = self.grid
        b = self.bin_np = {"keys_sorted": np.zeros(max(R, 1), np.uint64), "point_list": np.zeros(max(R, 1), np.uint32),
                           "ranges": np.zeros((gx * gy, 2), np.uint32), "final_T": np.zeros(H * W, np.float32),
                           "n_contrib": np.zeros(H * W, np.uint32)}
        self.bin = Bin(R=R, **{n: _ptr(a) for n, a in b.items()})
        o = self.out = {"color": np.zeros((3, H, W), np.float32), "language": np.zeros((F, H, W), np.float32),
                        "depth": np.zeros((1, H, W), np.float32), "opacity": np.zeros((1, H, W), np.float32),
                        "n_touched": np.zeros(P, np.int32)}
        self.img = Image(**{n: _ptr(a) for n, a in o.items()})
        lib().ols_oracle_render(C.byref(self.scene), C.byref(self.geom), C.byref(self.bin), C.byref(self.img))
        res = dict(o)
        res["radii"] = g["radii"]
        res["R"] = R
        res["keys_sorted"] = b["keys_sorted"][:R]
        res["point_list"] = b["point_list"][:R]
        res["ranges"] = b["ranges"]
        res["final_T"] = b["final_T"].reshape(H, W)
        res["n_contrib"] = b["n_contrib"].reshape(H, W)
        for n in ("depths", "means2D", "cov3D", "conic_opacity", "rgb", "clamped", "tiles_touched", "point_offsets"):
            res[n] = g[n]
        return res

    def backward(self, dL_dcolor, dL_dlanguage, dL_ddepth, compat: bool) -> Dict[str, np.ndarray]:
        P, F, M = self.P, self.F, self.M
        gin = {"dL_dcolor": _np(dL_dcolor).reshape(3, self.H, self.W), "dL_dlanguage": _np(dL_dlanguage).reshape(F, self.H, self.W),
               "dL_ddepth": _np(dL_ddepth).reshape(self.H, self.W)}
        out = {"dL_dmeans2D": np.zeros((P, 3), np.float32), "dL_dconic": np.zeros((P, 4), np.float32),
               "dL_dopacity": np.zeros((P, 1), np.float32), "dL_dcolors": np.zeros((P, 3), np.float32),
               "dL_dlang": np.zeros((P, F), np.float32), "dL_ddepths": np.zeros((P, 1), np.float32),
               "dL_dmeans3D": np.zeros((P, 3), np.float32), "dL_dcov3D": np.zeros((P, 6), np.float32),
               "dL_dsh": np.zeros((P, max(M, 1), 3), np.float32), "dL_dscales": np.zeros((P, 3), np.float32),
               "dL_drots": np.zeros((P, 4), np.float32), "dL_dtau": np.zeros((P, 6), np.float32)}
        self._gkeep = (gin, out)
        gr = Grads(compat=int(bool(compat)), **{n: _ptr(a) for n, a in gin.items()}, **{n: _ptr(a) for n, a in out.items()})
        lib().ols_oracle_backward(C.byref(self.scene), C.byref(self.geom), C.byref(self.bin), C.byref(gr))
        if M == 0:
            out["dL_dsh"] = np.zeros((P, 0, 3), np.float32)
        return out


class OracleDisRasterizer(OracleRasterizer):
    """The disentangled variant (D/): forward() / backward() mirror D/rasterize_points.cu:135-265 and :393-515."""

    def __init__(self, *, opacities_lang, scales_lang=None, rotations_lang=None, cov3D_precomp_lang=None, tile=16, **kw):
        super().__init__(tile=tile, **kw)
        k, P = self.keep, self.P
        k["opacities_lang"] = _np(opacities_lang).reshape(-1)
        k["scales_lang"] = None if scales_lang is None else _np(scales_lang).reshape(P, 3)
        k["rotations_lang"] = None if rotations_lang is None else _np(rotations_lang).reshape(P, 4)
        k["cov3D_precomp_lang"] = None if cov3D_precomp_lang is None else _np(cov3D_precomp_lang).reshape(P, 6)
        self.extra = DisExtra(**{n: _ptr(k[n]) for n in ("opacities_lang", "scales_lang", "rotations_lang", "cov3D_precomp_lang")})

    def forward(self) -> Dict[str, np.ndarray]:
        P, F, W, H = self.P, self.F, self.W, self.H
        g = self.geom_np = {
            "depths": np.zeros(P, np.float32), "radii": np.zeros(P, np.int32), "means2D": np.zeros((P, 2), np.float32),
            "cov3D": np.zeros((P, 6), np.float32), "conic_opacity": np.zeros((P, 4), np.float32),
            "rgb": np.zeros((P, 3), np.float32), "clamped": np.zeros((P, 3), np.uint8),
            "tiles_touched": np.zeros(P, np.uint32), "point_offsets": np.zeros(P, np.uint32)}
        gl = self.geoml_np = {
            "radii_lang": np.zeros(P, np.int32), "cov3D_lang": np.zeros((P, 6), np.float32),
            "conic_opacity_lang": np.zeros((P, 4), np.float32), "tiles_touched_lang": np.zeros(P, np.uint32),
            "point_offsets_lang": np.zeros(P, np.uint32)}
        self.geom = Geom(**{n: _ptr(a) for n, a in g.items()})
        self.geoml = DisGeom(**{n: _ptr(a) for n, a in gl.items()})
        Rl = C.c_int64(0)
        R = int(lib().ols_oracle_dis_preprocess(C.byref(self.scene), C.byref(self.extra), C.byref(self.geom),
                                                C.byref(self.geoml), C.byref(Rl)))
        Rl = int(Rl.value)
        gx, gy = self.grid

        def mkbin(n):
            return {"keys_sorted": np.zeros(max(n, 1), np.uint64), "point_list": np.zeros(max(n, 1), np.uint32),
                    "ranges": np.zeros((gx * gy, 2), np.uint32), "final_T": np.zeros(H * W, np.float32),
                    "n_contrib": np.zeros(H * W, np.uint32)}
        bc, bl = mkbin(R), mkbin(Rl)
        self.bin_np, self.binl_np = bc, bl
        self.bin = Bin(R=R, **{n: _ptr(a) for n, a in bc.items()})
        self.binl = Bin(R=Rl, **{n: _ptr(a) for n, a in bl.items()})
        oc = {"color": np.zeros((3, H, W), np.float32), "language": None, "depth": np.zeros((1, H, W), np.float32),
              "opacity": np.zeros((1, H, W), np.float32), "n_touched": np.zeros(P, np.int32)}
        ol = {"color": None, "language": np.zeros((F, H, W), np.float32), "depth": None,
              "opacity": np.zeros((1, H, W), np.float32), "n_touched": np.zeros(P, np.int32)}
        self.out, self.outl = oc, ol
        ic = Image(**{n: _ptr(a) for n, a in oc.items()})
        il = Image(**{n: _ptr(a) for n, a in ol.items()})
        lib().ols_oracle_dis_render(C.byref(self.scene), C.byref(self.geom), C.byref(self.geoml), C.byref(self.bin),
                                    C.byref(self.binl), C.byref(ic), C.byref(il))
        res = {"color": oc["color"], "depth": oc["depth"], "opacity": oc["opacity"], "n_touched": oc["n_touched"],
               "language": ol["language"], "opacity_lang": ol["opacity"], "n_touched_lang": ol["n_touched"],
               "radii": g["radii"], "radii_lang": gl["radii_lang"], "R": R, "R_lang": Rl}
        for suffix, b, n in (("", bc, R), ("_lang", bl, Rl)):
            res["keys_sorted" + suffix] = b["keys_sorted"][:n]
            res["point_list" + suffix] = b["point_list"][:n]
            res["ranges" + suffix] = b["ranges"]
            res["final_T" + suffix] = b["final_T"].reshape(H, W)
            res["n_contrib" + suffix] = b["n_contrib"].reshape(H, W)
        for n in ("depths", "means2D", "cov3D", "conic_opacity", "rgb", "clamped", "tiles_touched", "point_offsets"):
            res[n] = g[n]
        for n in ("cov3D_lang", "conic_opacity_lang", "tiles_touched_lang", "point_offsets_lang"):
            res[n] = gl[n]
        return res

    def backward(self, dL_dcolor, dL_dlanguage, dL_ddepth, compat: bool) -> Dict[str, np.ndarray]:
        P, F, M = self.P, self.F, self.M
        gin = {"dL_dcolor": _np(dL_dcolor).reshape(3, self.H, self.W), "dL_dlanguage": _np(dL_dlanguage).reshape(F, self.H, self.W),
               "dL_ddepth": _np(dL_ddepth).reshape(self.H, self.W)}
        z = lambda *shape: np.zeros(shape, np.float32)
        out = {"dL_dmeans2D": z(P, 3), "dL_dconic": z(P, 4), "dL_dconic_lang": z(P, 4), "dL_dopacity": z(P, 1),
               "dL_dopacity_lang": z(P, 1), "dL_dcolors": z(P, 3), "dL_dlang": z(P, F), "dL_ddepths": z(P, 1),
               "dL_dmeans3D": z(P, 3), "dL_dcov3D": z(P, 6), "dL_dcov3D_lang": z(P, 6), "dL_dsh": z(P, max(M, 1), 3),
               "dL_dscales": z(P, 3), "dL_dscales_lang": z(P, 3), "dL_drots": z(P, 4), "dL_drots_lang": z(P, 4),
               "dL_dtau": z(P, 6)}
        self._gkeep = (gin, out)
        gr = DisGrads(compat=int(bool(compat)), **{n: _ptr(a) for n, a in gin.items()}, **{n: _ptr(a) for n, a in out.items()})
        lib().ols_oracle_dis_backward(C.byref(self.scene), C.byref(self.extra), C.byref(self.geom), C.byref(self.geoml),
                                      C.byref(self.bin), C.byref(self.binl), C.byref(gr))
        if M == 0:
            out["dL_dsh"] = np.zeros((P, 0, 3), np.float32)
        return out
