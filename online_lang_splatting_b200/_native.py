"""ctypes binding of the C ABI in include/ols_b200.h (libols_b200.so).

PyTorch is used here only as the owner of device memory and streams: every call passes raw
``data_ptr()`` values and the current CUDA stream handle.  There is no CPU fallback -- if the
library is missing or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libols_b200.so")

OLS_OK = 0
OLS_ERR_OVERFLOW = -4
FLAG_PREFILTERED = 1 << 0
FLAG_DEBUG = 1 << 1
FLAG_BITEXACT_BLEND = 1 << 2
FLAG_BWD_EXACT = 1 << 3
FLAG_BWD_ACCUMULATE = 1 << 4
TIMING_TAGS = ("preprocess", "binning", "sort", "blend_fwd", "blend_bwd", "geometry_bwd", "ae", "other")
AE_MAX_LAYERS = 8
MAX_BATCH_VIEWS = 16

# every symbol include/ols_b200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = (
    "ols_abi_version", "ols_last_error", "ols_cuda_available", "ols_lang_workspace_size", "ols_lang_forward",
    "ols_lang_read_info", "ols_lang_backward", "ols_mark_visible", "ols_lang_workspace_view",
    "ols_lang_forward_host", "ols_timing_begin", "ols_timing_end", "ols_ae_plan_create", "ols_ae_plan_destroy", "ols_ae_forward",
    "ols_mapping_loss_forward", "ols_mapping_loss_backward", "ols_adam_step", "ols_knn_workspace_size", "ols_knn_mean_dist2",
    "ols_dis_workspace_size", "ols_dis_forward", "ols_dis_read_info", "ols_dis_backward", "ols_dis_workspace_view",
    "ols_hr_plan_create", "ols_hr_plan_destroy", "ols_hr_forward", "ols_hr_read_activation",
    "ols_ssim_loss_forward", "ols_ssim_loss_backward", "ols_densify_stats", "ols_densify_flags",
    "ols_ae_forward_bf16", "ols_hr_forward_features",
    "ols_lang_forward_batch", "ols_lang_backward_batch", "ols_lang_read_info_async", "ols_activate_params", "ols_adam_step_dev",
    "ols_pose_adam_step", "ols_online_ae_scratch_bytes", "ols_online_ae_param_count", "ols_online_ae_train_step",
)


class RasterArgs(C.Structure):
    _fields_ = [("P", C.c_int32), ("F", C.c_int32), ("sh_degree", C.c_int32), ("M", C.c_int32), ("W", C.c_int32),
                ("H", C.c_int32), ("tile", C.c_int32), ("flags", C.c_uint32), ("tanfovx", C.c_float),
                ("tanfovy", C.c_float), ("scale_modifier", C.c_float), ("_pad0", C.c_float)] + \
               [(n, C.c_void_p) for n in ("d_bg", "d_means3D", "d_shs", "d_colors_precomp", "d_language",
                                          "d_opacities", "d_scales", "d_rotations", "d_cov3D_precomp", "d_viewmatrix",
                                          "d_projmatrix", "d_projmatrix_raw", "d_campos", "d_workspace")] + \
               [("workspace_bytes", C.c_size_t), ("R_cap", C.c_int64)]


class FwdOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_color", "d_language", "d_depth", "d_opacity", "d_radii", "d_n_touched")]


class FwdInfo(C.Structure):
    _fields_ = [("R", C.c_int64), ("overflow", C.c_int32), ("max_tile_len", C.c_int32), ("n_visible", C.c_int32),
                ("_pad", C.c_int32 * 3)]


class BwdArgs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_dL_dout_color", "d_dL_dout_language", "d_dL_dout_depth", "d_radii",
                                          "d_dL_dmeans2D", "d_dL_dcolors", "d_dL_dlanguage", "d_dL_dopacity",
                                          "d_dL_dmeans3D", "d_dL_dcov3D", "d_dL_dsh", "d_dL_dscales",
                                          "d_dL_drotations", "d_dL_dtau", "d_dL_dtau_sum", "d_stat_max_radii2D",
                                          "d_stat_xyz_gradient_accum", "d_stat_denom")]


class DisArgs(C.Structure):
    """ols_dis_args: the disentangled (D/) rasterizer's arguments = ols_raster_args + the language footprint."""
    _fields_ = [("base", RasterArgs)] + \
               [(n, C.c_void_p) for n in ("d_opacities_lang", "d_scales_lang", "d_rotations_lang", "d_cov3D_precomp_lang")] + \
               [("R_cap_lang", C.c_int64)]


class DisFwdOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_color", "d_language", "d_depth", "d_opacity", "d_opacity_lang", "d_radii",
                                          "d_radii_lang", "d_n_touched", "d_n_touched_lang")]


class DisBwdArgs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_dL_dout_color", "d_dL_dout_language", "d_dL_dout_depth", "d_radii",
                                          "d_radii_lang", "d_dL_dmeans2D", "d_dL_dcolors", "d_dL_dlanguage",
                                          "d_dL_dopacity", "d_dL_dopacity_lang", "d_dL_dmeans3D", "d_dL_dcov3D",
                                          "d_dL_dcov3D_lang", "d_dL_dsh", "d_dL_dscales", "d_dL_dscales_lang",
                                          "d_dL_drotations", "d_dL_drotations_lang", "d_dL_dtau")]


class LossArgs(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("F", C.c_int32), ("lang_w", C.c_int32), ("lang_h", C.c_int32),
                ("alpha", C.c_float), ("rgb_boundary_threshold", C.c_float), ("exposure_a", C.c_float),
                ("exposure_b", C.c_float), ("lambda_lang", C.c_float)] + \
               [(n, C.c_void_p) for n in ("d_image", "d_depth", "d_language", "d_gt_image", "d_gt_depth", "d_gt_lang",
                                          "d_opacity", "d_grad_mask", "d_exposure_a", "d_exposure_b")]


class AdamGroup(C.Structure):
    _fields_ = [("offset", C.c_int64), ("count", C.c_int64), ("lr", C.c_float), ("activation", C.c_int32),
                ("period", C.c_int32), ("head", C.c_int32), ("lr_tail", C.c_float), ("_pad", C.c_float)]


ACT_NONE, ACT_EXP, ACT_SIGMOID, ACT_NORMALIZE4 = 0, 1, 2, 3


class PoseStep(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_grad_tau", "d_grad_exposure", "d_exposure", "d_exp_avg", "d_exp_avg_sq", "d_step",
                                          "d_R", "d_T", "d_projection", "d_viewmatrix", "d_projmatrix", "d_campos",
                                          "d_converged")] + \
               [(n, C.c_float) for n in ("lr_rot", "lr_trans", "lr_exposure", "beta1", "beta2", "eps", "converged_threshold")] + \
               [("zero_grads", C.c_int32)]


class WsView(C.Structure):
    _fields_ = [("d_records", C.c_void_p), ("rec_floats", C.c_int32), ("n_tiles", C.c_int32)] + \
               [(n, C.c_void_p) for n in ("d_cov3D", "d_clamped", "d_tiles_touched", "d_ranges", "d_point_list",
                                          "d_keys", "d_final_T", "d_n_contrib")]


class HostOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("h_color", "h_language", "h_depth", "h_opacity", "h_radii", "h_n_touched")]


class AEChain(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dims", C.c_int32 * (AE_MAX_LAYERS + 1)), ("normalize", C.c_int32),
                ("input_bf16", C.c_int32), ("precision", C.c_int32), ("_pad", C.c_int32),
                ("d_weight", C.c_void_p * AE_MAX_LAYERS), ("d_bias", C.c_void_p * AE_MAX_LAYERS)]


class SsimArgs(C.Structure):
    _fields_ = [("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("w_l1", C.c_float), ("w_ssim", C.c_float),
                ("d_image", C.c_void_p), ("d_gt", C.c_void_p)]


class DensifyParams(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("max_grad", "min_opacity", "extent", "max_screen_size", "percent_dense")]


HR_N_CONV = 13


class HRWeights(C.Structure):
    _fields_ = [("d_weight", C.c_void_p * HR_N_CONV), ("d_bias", C.c_void_p * HR_N_CONV)]


_LIB: Optional[C.CDLL] = None


class OlsError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"ols_b200 error {status}: {msg}")
        self.status = status


def lib() -> C.CDLL:
    """Loads libols_b200.so (building it first if the sources are newer and nvcc is present)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    L = C.CDLL(LIB_PATH)
    L.ols_abi_version.restype = C.c_int
    L.ols_last_error.restype = C.c_char_p
    L.ols_cuda_available.restype = C.c_int
    L.ols_lang_workspace_size.restype = C.c_size_t
    L.ols_lang_workspace_size.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64]
    L.ols_lang_forward.argtypes = [C.POINTER(RasterArgs), C.POINTER(FwdOut), C.c_void_p]
    L.ols_lang_read_info.argtypes = [C.c_void_p, C.POINTER(FwdInfo), C.c_void_p]
    L.ols_lang_backward.argtypes = [C.POINTER(RasterArgs), C.POINTER(BwdArgs), C.c_void_p]
    L.ols_lang_forward_batch.argtypes = [C.POINTER(RasterArgs), C.POINTER(FwdOut), C.c_int32, C.c_void_p]
    L.ols_lang_backward_batch.argtypes = [C.POINTER(RasterArgs), C.POINTER(BwdArgs), C.c_int32, C.c_void_p]
    L.ols_lang_read_info_async.argtypes = [C.POINTER(RasterArgs), C.c_int32, C.c_void_p, C.c_void_p]
    L.ols_mark_visible.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ols_lang_workspace_view.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                          C.c_void_p, C.POINTER(WsView)]
    L.ols_lang_forward_host.argtypes = [C.POINTER(RasterArgs), C.POINTER(HostOut), C.POINTER(C.c_int64)]
    L.ols_dis_workspace_size.restype = C.c_size_t
    L.ols_dis_workspace_size.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64]
    L.ols_dis_forward.argtypes = [C.POINTER(DisArgs), C.POINTER(DisFwdOut), C.c_void_p]
    L.ols_dis_read_info.argtypes = [C.POINTER(DisArgs), C.POINTER(FwdInfo), C.POINTER(FwdInfo), C.c_void_p]
    L.ols_dis_backward.argtypes = [C.POINTER(DisArgs), C.POINTER(DisBwdArgs), C.c_void_p]
    L.ols_dis_workspace_view.argtypes = [C.POINTER(DisArgs), C.POINTER(WsView), C.POINTER(WsView)]
    L.ols_mapping_loss_forward.argtypes = [C.POINTER(LossArgs), C.c_void_p, C.c_void_p, C.c_void_p]
    L.ols_mapping_loss_backward.argtypes = [C.POINTER(LossArgs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p]
    L.ols_adam_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(AdamGroup), C.c_int32,
                                C.c_double, C.c_double, C.c_double, C.c_int64, C.c_void_p]
    L.ols_adam_step_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(AdamGroup), C.c_int32,
                                    C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    L.ols_pose_adam_step.argtypes = [C.POINTER(PoseStep), C.c_void_p]
    L.ols_online_ae_scratch_bytes.restype = C.c_size_t
    L.ols_online_ae_param_count.restype = C.c_int32
    L.ols_online_ae_train_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float,
                                           C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                           C.c_void_p]
    L.ols_activate_params.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 7
    L.ols_knn_workspace_size.restype = C.c_size_t
    L.ols_knn_workspace_size.argtypes = [C.c_int32]
    L.ols_knn_mean_dist2.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.ols_timing_begin.argtypes = [C.c_int32]
    L.ols_timing_end.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    L.ols_ae_plan_create.argtypes = [C.POINTER(AEChain), C.POINTER(C.c_void_p), C.c_void_p]
    L.ols_ae_plan_destroy.argtypes = [C.c_void_p]
    L.ols_ae_plan_destroy.restype = None
    L.ols_ae_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.ols_ssim_loss_forward.argtypes = [C.POINTER(SsimArgs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ols_ssim_loss_backward.argtypes = [C.POINTER(SsimArgs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ols_densify_stats.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ols_densify_flags.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(DensifyParams), C.c_void_p, C.c_void_p, C.c_void_p]
    L.ols_hr_plan_create.argtypes = [C.POINTER(HRWeights), C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_void_p]
    L.ols_hr_plan_destroy.argtypes = [C.c_void_p]
    L.ols_hr_plan_destroy.restype = None
    L.ols_hr_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                 C.c_void_p, C.c_void_p]
    L.ols_ae_forward_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.ols_hr_forward_features.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                          C.c_int32, C.c_void_p, C.c_void_p]
    L.ols_hr_read_activation.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
    _LIB = L
    return L


def check(status: int) -> None:
    if status != OLS_OK:
        raise OlsError(status, lib().ols_last_error().decode("utf-8", "replace"))


def require_cuda() -> None:
    if not lib().ols_cuda_available():
        raise OlsError(-2, "no CUDA device: online_lang_splatting_b200 has no CPU path")


def ptr(t) -> Optional[int]:
    """data_ptr of a tensor; None / empty tensors map to NULL exactly like the reference's
    `torch.Tensor([])` == "not provided" convention (diff_gaussian_rasterization/__init__.py:530-548)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def timing_begin(max_marks: int = 8192) -> None:
    check(lib().ols_timing_begin(int(max_marks)))


def timing_end():
    """-> {tag: (total_ms, intervals)} for the calls made since timing_begin() on this thread."""
    ms = (C.c_float * len(TIMING_TAGS))()
    cnt = (C.c_int32 * len(TIMING_TAGS))()
    check(lib().ols_timing_end(ms, cnt))
    return {t: (float(ms[i]), int(cnt[i])) for i, t in enumerate(TIMING_TAGS)}
