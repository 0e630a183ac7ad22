/* ols_b200.h -- C ABI of the B200-native language-feature Gaussian rasterizer + autoencoder.
 *
 * This is the drop-in boundary for the hot path of rpng/online_lang_splatting.  Every entry
 * point replaces one function the reference binds through pybind11 in
 *   submodules/diff-gaussian-rasterization/ext.cpp:15-21
 * (the "P/" variant that gaussian_renderer/__init__.py:316-335 really calls), or one torch.nn
 * call of language/autoencoder/model.py.  Conventions:
 *
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - every pointer named d_* is DEVICE memory of the current CUDA device, h_* is HOST memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - all functions return OLS_OK (0) or a negative ols_status; the message of the last error on
 *     the calling thread is returned by ols_last_error();  nothing throws across the ABI;
 *   - no global mutable state besides the thread-local error string, so several processes,
 *     devices and streams can use the library concurrently;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     OLS_ERR_CUDA.
 *
 * Matrices follow the reference's row-vector convention (transposed world->view, see
 * utils/camera_utils.py:103-113 and cuda_rasterizer/auxiliary.h:58-77): element (r,c) of the
 * mathematical matrix is m[c*4+r].  Image tensors are planar CHW float32.
 */
#ifndef OLS_B200_H_
#define OLS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OLS_ABI_VERSION 2

typedef enum ols_status {
    OLS_OK = 0,
    OLS_ERR_INVALID = -1,      /* bad argument combination (reference: Python Exception / AT_ERROR)      */
    OLS_ERR_CUDA = -2,         /* CUDA runtime error or no device                                        */
    OLS_ERR_WORKSPACE = -3,    /* workspace too small for the fixed part                                 */
    OLS_ERR_OVERFLOW = -4,     /* instance capacity (R_cap) exceeded; nothing was rendered, see ols_fwd_info */
    OLS_ERR_UNSUPPORTED = -5   /* feature dimension / tile size not compiled in                          */
} ols_status;

/* flags for ols_raster_args.flags */
#define OLS_FLAG_PREFILTERED     (1u << 0)  /* reference: raster_settings.prefiltered                      */
#define OLS_FLAG_DEBUG           (1u << 1)  /* reference: raster_settings.debug -> sync + check after each kernel */
#define OLS_FLAG_BITEXACT_BLEND  (1u << 2)  /* accumulate c*alpha*T in the reference's operation order      */
#define OLS_FLAG_BWD_EXACT       (1u << 3)  /* backward: mathematically exact gradients instead of the
                                               reference's Q1-Q3 behaviour          */
#define OLS_FLAG_BWD_NO_PACK     (1u << 5)  /* compat backward: keep one thread per pixel even where the reference's
                                               lossy reduction drops the pixel (debugging / A-B comparison)  */
#define OLS_FLAG_BWD_ACCUMULATE  (1u << 4)  /* backward: add into dL_dmeans3D/sh/opacity/scales/rotations/
                                               language/cov3D instead of overwriting (sums the views of a
                                               mapping iteration, utils/slam_backend.py:510-670)            */

/* ---------------------------------------------------------------------------------------------
 * Arguments shared by forward and backward.  Mirrors the positional argument list of
 * RasterizeLanguageGaussiansCUDA (rasterize_points.h:51-73 / rasterize_points.cu:135-157) plus
 * GaussianRasterizationSettings (diff_gaussian_rasterization/__init__.py:405-419).
 * Nullable inputs follow the reference's "empty tensor == not provided" rule.
 * ------------------------------------------------------------------------------------------- */
typedef struct ols_raster_args {
    int32_t P;              /* number of Gaussians                                                     */
    int32_t F;              /* language feature channels (config.h NUM_LANGUAGE_CHANNELS: 15 or 3)       */
    int32_t sh_degree;      /* active SH degree D                                                      */
    int32_t M;              /* SH coefficients per Gaussian ((max_degree+1)^2), 0 if d_shs == NULL       */
    int32_t W, H;           /* image size                                                              */
    int32_t tile;           /* tile edge in pixels: 15 (config.h of P/) or 16 (D/ and perf mode)         */
    uint32_t flags;         /* OLS_FLAG_*                                                              */
    float tanfovx, tanfovy;
    float scale_modifier;
    float _pad0;
    const float* d_bg;              /* [3]                                                             */
    const float* d_means3D;         /* [P,3]                                                           */
    const float* d_shs;             /* [P,M,3] or NULL                                                 */
    const float* d_colors_precomp;  /* [P,3]   or NULL  (exactly one of shs / colors_precomp)          */
    const float* d_language;        /* [P,F]   (language_precomp)                                      */
    const float* d_opacities;       /* [P]                                                             */
    const float* d_scales;          /* [P,3]   or NULL                                                 */
    const float* d_rotations;       /* [P,4]   or NULL  (r,x,y,z), used as given (not normalised)      */
    const float* d_cov3D_precomp;   /* [P,6]   or NULL  (exactly one of scale+rot / cov3D_precomp)     */
    const float* d_viewmatrix;      /* [16]                                                            */
    const float* d_projmatrix;      /* [16]  full projection                                           */
    const float* d_projmatrix_raw;  /* [16]  projection only (used by the pose Jacobian in backward)   */
    const float* d_campos;          /* [3]                                                             */
    void* d_workspace;              /* ols_lang_workspace_size() bytes, 256-B aligned, kept by the caller
                                       from forward to backward (reference: geomBuffer+binningBuffer+imgBuffer) */
    size_t workspace_bytes;
    int64_t R_cap;                  /* capacity for Gaussian/tile instances the workspace was sized for  */
} ols_raster_args;

/* Forward outputs.  Reference: the tuple returned at rasterize_points.cu:230-240.  All buffers are
 * written completely by the call (no pre-zeroing required). */
typedef struct ols_fwd_out {
    float* d_color;       /* [3,H,W]                                                                  */
    float* d_language;    /* [F,H,W]                                                                  */
    float* d_depth;       /* [1,H,W]                                                                  */
    float* d_opacity;     /* [1,H,W]                                                                  */
    int32_t* d_radii;     /* [P]                                                                      */
    int32_t* d_n_touched; /* [P]                                                                      */
} ols_fwd_out;

/* Written asynchronously into the first bytes of the workspace; copy it back with
 * ols_lang_read_info() (one 32-byte D2H) when the host needs R or the overflow flag. */
typedef struct ols_fwd_info {
    int64_t R;            /* num_rendered (reference: return value of LanguageRasterizer::forward)      */
    int32_t overflow;     /* 1 if R > R_cap: binning/sort/blend were skipped, outputs are zero          */
    int32_t max_tile_len; /* longest per-tile list                                                     */
    int32_t n_visible;    /* Gaussians with radii > 0                                                  */
    int32_t _pad[3];
} ols_fwd_info;

/* Backward.  Reference: RasterizeLanguageGaussiansBackwardCUDA (rasterize_points.h:121-148,
 * rasterize_points.cu:333-455).  All d_dL_d* outputs are fully written by the call (the library
 * zero-fills what it accumulates into). */
typedef struct ols_bwd_args {
    const float* d_dL_dout_color;     /* [3,H,W]                                                       */
    const float* d_dL_dout_language;  /* [F,H,W]                                                       */
    const float* d_dL_dout_depth;     /* [1,H,W]                                                       */
    const int32_t* d_radii;           /* [P] from forward                                              */
    float* d_dL_dmeans2D;    /* [P,3]  (NDC units, z = 0)                                              */
    float* d_dL_dcolors;     /* [P,3]                                                                  */
    float* d_dL_dlanguage;   /* [P,F]                                                                  */
    float* d_dL_dopacity;    /* [P,1]                                                                  */
    float* d_dL_dmeans3D;    /* [P,3]                                                                  */
    float* d_dL_dcov3D;      /* [P,6]                                                                  */
    float* d_dL_dsh;         /* [P,M,3] or NULL when M == 0                                            */
    float* d_dL_dscales;     /* [P,3]                                                                  */
    float* d_dL_drotations;  /* [P,4]                                                                  */
    float* d_dL_dtau;        /* [P,6]  per-Gaussian pose gradient (rho | theta) as the reference returns it
                                (rasterize_points.cu:452); may be NULL when d_dL_dtau_sum is given        */
    float* d_dL_dtau_sum;    /* [6]    the same summed over the Gaussians -- what the reference's Python computes
                                right after the call (diff_gaussian_rasterization/__init__.py:383-385) -- or NULL */
    /* Optional densification statistics, updated in place for the Gaussians with radii > 0 of every view of the call
     * (all three or none; taken from grads[0] in a batch).  Same arithmetic as ols_densify_stats, i.e. what the mapping
     * loop does per view right after backward (utils/slam_backend.py:719-728, gaussian_model.py:965-969):
     *   max_radii2D = max(max_radii2D, radii);  xyz_gradient_accum += |dL_dmeans2D.xy|;  denom += 1 */
    float* d_stat_max_radii2D;         /* [P] or NULL */
    float* d_stat_xyz_gradient_accum;  /* [P] or NULL */
    float* d_stat_denom;               /* [P] or NULL */
} ols_bwd_args;

int ols_abi_version(void);
const char* ols_last_error(void);

/* 1 if a CUDA device is usable from this process, else 0 (never fails). */
int ols_cuda_available(void);

/* Bytes of workspace needed for a (P, W, H, tile, F) problem with room for R_cap instances.
 * Replaces the three growable byte tensors + obtain()/required<T>() of
 * cuda_rasterizer/rasterizer_impl.cu:155-212.  Returns 0 on invalid arguments. */
size_t ols_lang_workspace_size(int32_t P, int32_t F, int32_t W, int32_t H, int32_t tile, int64_t R_cap);

/* replaces _C.rasterize_language_gaussians (ext.cpp:18).  Asynchronous on `stream`; no host sync. */
int ols_lang_forward(const ols_raster_args* args, const ols_fwd_out* out, void* stream);

/* D2H copy of the ols_fwd_info header of a workspace (synchronises `stream`). */
int ols_lang_read_info(const void* d_workspace, ols_fwd_info* h_info, void* stream);

/* replaces _C.rasterize_language_gaussians_backward (ext.cpp:19). */
int ols_lang_backward(const ols_raster_args* args, const ols_bwd_args* grads, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-view batch: V views of the SAME Gaussians in one set of launches (grid.y = view).  This is the loop over the
 * window keyframes of a mapping iteration (utils/slam_backend.py:510-662 calls render() 8-12 times on the same
 * Gaussians, sums the losses and back-propagates once), and the loop over rendered views of any evaluation pass.
 * views[0..V) must agree in P, F, M, sh_degree, W, H, tile, flags, scale_modifier, R_cap and in every Gaussian
 * parameter pointer; viewmatrix / projmatrix / projmatrix_raw / campos / tanfov / bg / workspace are per view
 * (V <= OLS_MAX_BATCH_VIEWS, one workspace of ols_lang_workspace_size() bytes per view).  Each Gaussian is read and
 * its 3D covariance computed once for all views; the view-independent covariances are kept in views[0]'s workspace.
 * Backward: grads[v] holds view v's image-space gradients, radii, d_dL_dmeans2D and d_dL_dtau / d_dL_dtau_sum; the
 * parameter gradients (means3D, colors, language, opacity, cov3D, sh, scales, rotations) are taken from grads[0]
 * and receive the SUM over the views -- what autograd accumulates across the reference's V render() calls.
 * ols_lang_forward / ols_lang_backward are the V = 1 case of the same kernels.
 * ------------------------------------------------------------------------------------------- */
#define OLS_MAX_BATCH_VIEWS 16
int ols_lang_forward_batch(const ols_raster_args* views, const ols_fwd_out* outs, int32_t V, void* stream);
int ols_lang_backward_batch(const ols_raster_args* views, const ols_bwd_args* grads, int32_t V, void* stream);
/* asynchronous copy of the V info headers into h_info[V] (pinned host memory for a truly asynchronous copy); the
 * caller orders its read with an event / stream synchronisation of its own */
int ols_lang_read_info_async(const ols_raster_args* views, int32_t V, ols_fwd_info* h_info, void* stream);

/* replaces _C.mark_visible (ext.cpp:20; checkFrustum, rasterizer_impl.cu:54-66): present[i] = view.z > 0.2 */
int ols_mark_visible(int32_t P, const float* d_means3D, const float* d_viewmatrix, const float* d_projmatrix,
                     uint8_t* d_present, void* stream);

/* Views into the workspace for tests / debugging (device pointers; valid after forward).
 * Reference counterpart: the fromChunk() carving of rasterizer_impl.cu:173-212. */
typedef struct ols_ws_view {
    const float* d_records;        /* [P, rec_floats] packed per-Gaussian blend record:
                                      x, y, conicA, conicB, conicC, opacity, pth, depth, r, g, b, lang[F], 0-pad,
                                      ex, ey (last two floats: conservative half-extents of alpha >= 1/255) */
    int32_t rec_floats;
    int32_t n_tiles;
    const float* d_cov3D;          /* [P,6]                                                            */
    const uint8_t* d_clamped;      /* [P,3]                                                            */
    const uint32_t* d_tiles_touched; /* [P]                                                            */
    const uint32_t* d_ranges;      /* [n_tiles,2]                                                      */
    const uint32_t* d_point_list;  /* [R] sorted Gaussian ids                                          */
    const uint64_t* d_keys;        /* [R] per tile segment: (depth_bits << 32 | gaussian id), sorted -- written only
                                      with OLS_FLAG_DEBUG (the blend passes read d_point_list; binning writes bare ids) */
    const float* d_final_T;        /* [H*W]                                                            */
    const uint32_t* d_n_contrib;   /* [H*W]                                                            */
} ols_ws_view;
int ols_lang_workspace_view(int32_t P, int32_t F, int32_t W, int32_t H, int32_t tile, int64_t R_cap,
                            const void* d_workspace, ols_ws_view* view);

/* Host-buffer entry point: the same forward, but every input and output lives in HOST memory.
 * Copies inputs H2D, renders, copies the six outputs D2H, synchronises.  Device scratch is owned
 * by the library per call.  This is the form a non-torch binding (ctypes / cgo / JNI) would use. */
typedef struct ols_host_out {
    float* h_color; float* h_language; float* h_depth; float* h_opacity; int32_t* h_radii; int32_t* h_n_touched;
} ols_host_out;
int ols_lang_forward_host(const ols_raster_args* host_args /* all d_* fields hold HOST pointers; workspace ignored */,
                          const ols_host_out* out, int64_t* num_rendered);

/* ---------------------------------------------------------------------------------------------
 * Disentangled variant ("D/": submodules/diff-gaussian-rasterization-disentangle-optim).  Every
 * Gaussian carries a second footprint (opacities_lang, scales_lang, rotations_lang) used only by
 * the language pass; colour + depth and language are binned, sorted and blended independently.
 * Entry points replace the pybind functions of D/ext.cpp:15-21 whose signatures are
 * D/rasterize_points.h:51-80 (forward, 25 positional arguments, 15 returns) and :128-162
 * (backward, 32 arguments, 14 returns).
 * ------------------------------------------------------------------------------------------- */
typedef struct ols_dis_args {
    ols_raster_args base;               /* colour footprint + everything shared; base.d_workspace holds
                                           ols_dis_workspace_size() bytes; base.R_cap = colour-list capacity */
    const float* d_opacities_lang;      /* [P]                                                           */
    const float* d_scales_lang;         /* [P,3] or NULL                                                 */
    const float* d_rotations_lang;      /* [P,4] or NULL                                                 */
    const float* d_cov3D_precomp_lang;  /* [P,6] or NULL (exactly one of scale+rot / cov3D, as for colour) */
    int64_t R_cap_lang;                 /* language-list capacity                                        */
} ols_dis_args;

/* Forward outputs: the tensors returned at D/rasterize_points.cu:251-265. */
typedef struct ols_dis_fwd_out {
    float* d_color;            /* [3,H,W] */
    float* d_language;         /* [F,H,W] */
    float* d_depth;            /* [1,H,W] */
    float* d_opacity;          /* [1,H,W] */
    float* d_opacity_lang;     /* [1,H,W] */
    int32_t* d_radii;          /* [P] */
    int32_t* d_radii_lang;     /* [P] */
    int32_t* d_n_touched;      /* [P] */
    int32_t* d_n_touched_lang; /* [P] */
} ols_dis_fwd_out;

/* Backward: the 14 tensors returned at D/rasterize_points.cu:499-515.  Language gradients reach
 * language, opacity_lang, scales_lang and rotations_lang only (D/backward.cu:1074,1117). */
typedef struct ols_dis_bwd_args {
    const float* d_dL_dout_color;     /* [3,H,W] */
    const float* d_dL_dout_language;  /* [F,H,W] */
    const float* d_dL_dout_depth;     /* [1,H,W] */
    const int32_t* d_radii;           /* [P] from forward */
    const int32_t* d_radii_lang;      /* [P] from forward */
    float* d_dL_dmeans2D;         /* [P,3] */
    float* d_dL_dcolors;          /* [P,3] */
    float* d_dL_dlanguage;        /* [P,F] */
    float* d_dL_dopacity;         /* [P,1] */
    float* d_dL_dopacity_lang;    /* [P,1] */
    float* d_dL_dmeans3D;         /* [P,3] */
    float* d_dL_dcov3D;           /* [P,6] */
    float* d_dL_dcov3D_lang;      /* [P,6] */
    float* d_dL_dsh;              /* [P,M,3] or NULL when M == 0 */
    float* d_dL_dscales;          /* [P,3] */
    float* d_dL_dscales_lang;     /* [P,3] */
    float* d_dL_drotations;       /* [P,4] */
    float* d_dL_drotations_lang;  /* [P,4] */
    float* d_dL_dtau;             /* [P,6] */
} ols_dis_bwd_args;

size_t ols_dis_workspace_size(int32_t P, int32_t F, int32_t W, int32_t H, int32_t tile, int64_t R_cap, int64_t R_cap_lang);
/* replaces D/'s _C.rasterize_language_gaussians; asynchronous, no host sync */
int ols_dis_forward(const ols_dis_args* args, const ols_dis_fwd_out* out, void* stream);
/* both info headers (colour list, language list) in one call; synchronises `stream` */
int ols_dis_read_info(const ols_dis_args* args, ols_fwd_info* h_info_color, ols_fwd_info* h_info_lang, void* stream);
/* replaces D/'s _C.rasterize_language_gaussians_backward */
int ols_dis_backward(const ols_dis_args* args, const ols_dis_bwd_args* grads, void* stream);
/* workspace views of the two lists (tests / debugging) */
int ols_dis_workspace_view(const ols_dis_args* args, ols_ws_view* view_color, ols_ws_view* view_lang);

/* ---------------------------------------------------------------------------------------------
 * Mapping loss on the rasterizer's outputs (the caller either side of render() in the mapping loop):
 *   get_loss_mapping / get_loss_mapping_rgbd  (utils/slam_utils.py:121-165)  +
 *   F.interpolate(gt_lang_feat, bilinear, align_corners=False) + l1_loss      (utils/slam_backend.py:576-592)
 *   loss = alpha * mean|(exp(a) image + b - gt) m_rgb| + (1 - alpha) * mean|(depth - gt_depth) m_d|
 *          + lambda_lang * mean|language - upsample(gt_lang)|
 * The low-resolution language target stays on the device; no 15 x H x W copy per iteration.
 * ------------------------------------------------------------------------------------------- */
typedef struct ols_loss_args {
    int32_t W, H;              /* rendered size                                                        */
    int32_t F;                 /* language channels, 0 = no language term                              */
    int32_t lang_w, lang_h;    /* size of the low-resolution language target (192 x 192 in the reference) */
    float alpha;               /* config["Training"]["alpha"], default 0.95                            */
    float rgb_boundary_threshold;
    float exposure_a, exposure_b;  /* viewpoint.exposure_a / _b; pass 0, 0 for initialization=True      */
    float lambda_lang;         /* self.lamda_lang                                                      */
    const float* d_image;      /* [3,H,W] render()["render"]                                           */
    const float* d_depth;      /* [1,H,W] render()["depth"]                                            */
    const float* d_language;   /* [F,H,W] render()["language"] or NULL                                 */
    const float* d_gt_image;   /* [3,H,W] viewpoint.original_image                                     */
    const float* d_gt_depth;   /* [1,H,W] viewpoint.depth                                              */
    const float* d_gt_lang;    /* [F,lang_h,lang_w] viewpoint.gt_lang_feat or NULL                     */
    /* tracking form, get_loss_tracking_rgbd (utils/slam_utils.py:96-118); both NULL for the mapping form */
    const float* d_opacity;    /* [1,H,W] render()["opacity"]: weights the colour residual, gates depth at > 0.95 */
    const float* d_grad_mask;  /* [1,H,W] viewpoint.grad_mask (0/1 floats) or NULL                     */
    /* device-resident exposure parameters (viewpoint.exposure_a / _b are nn.Parameters, utils/camera_utils.py:59-64):
     * when non-NULL they replace the two floats above, so the caller never reads them back to the host */
    const float* d_exposure_a; /* [1] or NULL                                                          */
    const float* d_exposure_b; /* [1] or NULL                                                          */
} ols_loss_args;
/* d_out6 = [l1_rgb, l1_depth, l1_lang, dloss/dexposure_a, dloss/dexposure_b, loss]; d_scratch8: 8 floats */
int ols_mapping_loss_forward(const ols_loss_args* args, float* d_out6, float* d_scratch8, void* stream);
/* gradients w.r.t. the three rendered tensors, scaled by the device scalar *d_upstream (dL/dloss) */
int ols_mapping_loss_backward(const ols_loss_args* args, const float* d_upstream, float* d_dL_dimage, float* d_dL_ddepth,
                              float* d_dL_dlanguage, float* d_dL_dopacity /* tracking form only, may be NULL */, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Adam step over the flat parameter / gradient buffers (the step right after backward + all-reduce).
 * Reference: torch.optim.Adam(param_groups, lr=0.0, eps=1e-15) of GaussianModel.training_setup
 * (gaussian_splatting/scene/gaussian_model.py:393-437), one group per parameter tensor.  Groups are
 * consecutive segments of the flat buffer (sharding.FlatGradBuffer layout), each with its own lr.
 * ------------------------------------------------------------------------------------------- */
#define OLS_ADAM_MAX_GROUPS 8
/* The rasterizer's backward produces gradients w.r.t. the ACTIVATED parameters it was given (get_scaling = exp,
 * get_opacity = sigmoid, get_rotation = F.normalize; gaussian_model.py:67-72,93-130) while Adam updates the raw ones;
 * in the reference autograd applies the activations' Jacobians in between.  With `activation` set the kernel applies
 * that chain rule to the flat gradient on the fly (d_param holds the RAW parameters). */
#define OLS_ACT_NONE       0
#define OLS_ACT_EXP        1   /* s = exp(p):            dL/dp = dL/ds * s                                  */
#define OLS_ACT_SIGMOID    2   /* o = sigmoid(p):        dL/dp = dL/do * o (1 - o)                          */
#define OLS_ACT_NORMALIZE4 3   /* q^ = q / max(|q|,1e-12), rows of 4: dL/dq = (g - q^ (q^ . g)) / |q|          */
typedef struct ols_adam_group {
    int64_t offset;   /* first element of the group in the flat buffers */
    int64_t count;    /* elements                                        */
    float lr;
    int32_t activation;   /* OLS_ACT_*                                                                     */
    int32_t period, head; /* period > 0: elements whose index within the group modulo `period` is >= `head` use
                             lr_tail -- f_rest at feature_lr / 20 inside the interleaved [P,M,3] SH block
                             (gaussian_model.py:404-413: f_dc and f_rest are separate groups in the reference) */
    float lr_tail;
    float _pad;
} ols_adam_group;
/* `step` is the 1-based step count after this update (torch's state["step"]). */
int ols_adam_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n,
                  const ols_adam_group* groups, int32_t n_groups, double beta1, double beta2, double eps, int64_t step,
                  void* stream);

/* The same step with the step count on the device (torch.optim.Adam(capturable=True)): *d_step is the number of steps
 * taken so far; the kernel uses *d_step + 1 for the bias corrections and a follow-up kernel increments it, so a
 * captured CUDA graph advances the optimiser state correctly on every replay. */
int ols_adam_step_dev(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n,
                      const ols_adam_group* groups, int32_t n_groups, double beta1, double beta2, double eps,
                      int64_t* d_step, void* stream);

/* Activated parameters from the raw ones (gaussian_model.py:93-130: get_opacity, get_scaling, get_rotation) in one
 * kernel -- what follows an optimiser step before the next render().  scale_cols = 1 (isotropic) or 3. */
int ols_activate_params(int32_t P, int32_t scale_cols, const float* d_opacity_raw, const float* d_scaling_raw,
                        const float* d_rotation_raw, float* d_opacity, float* d_scaling, float* d_rotation, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Pose step of the tracking loop (utils/slam_frontend.py:216-262, utils/pose_utils.py:60-95 update_pose): Adam over
 * (cam_rot_delta, cam_trans_delta, exposure_a, exposure_b), T_w2c <- SE3_exp(tau) T_w2c, deltas back to zero, and the
 * camera tensors render() reads rebuilt in place -- one kernel, no host synchronisation (the reference needs two per
 * iteration).  Matrices in the reference's transposed storage; d_R row-major 3x3.
 * ------------------------------------------------------------------------------------------- */
typedef struct ols_pose_step {
    const float* d_grad_tau;       /* [6] (rho | theta) = ols_bwd_args.d_dL_dtau_sum                              */
    const float* d_grad_exposure;  /* [2] dL/d(exposure_a, exposure_b), or NULL                                   */
    float* d_exposure;             /* [2] exposure_a, exposure_b, updated in place, or NULL                       */
    float* d_exp_avg;              /* [8] Adam moments, order: rot 3 | trans 3 | exposure 2                       */
    float* d_exp_avg_sq;           /* [8]                                                                         */
    int64_t* d_step;               /* Adam step count, incremented                                                */
    float* d_R;                    /* [9] world->camera rotation, updated                                         */
    float* d_T;                    /* [3] world->camera translation, updated                                      */
    const float* d_projection;     /* [16] projection_matrix                                                      */
    float* d_viewmatrix;           /* [16] out: world_view_transform                                              */
    float* d_projmatrix;           /* [16] out: full_proj_transform                                               */
    float* d_campos;               /* [3]  out: camera_center                                                     */
    int32_t* d_converged;          /* set to 1 when |tau| < converged_threshold (never cleared), or NULL          */
    float lr_rot, lr_trans, lr_exposure;   /* config Training.lr.cam_rot_delta / cam_trans_delta, 0.01            */
    float beta1, beta2, eps;               /* torch.optim.Adam defaults: 0.9, 0.999, 1e-8                         */
    float converged_threshold;             /* 1e-4                                                                */
    int32_t zero_grads;                    /* 1: clear d_grad_tau / d_grad_exposure after use (optimizer.zero_grad())  */
} ols_pose_step;
int ols_pose_adam_step(const ols_pose_step* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * distCUDA2 (submodules/simple-knn/spatial.cu + simple_knn.cu:120-220): mean squared distance of every
 * point to its 3 nearest neighbours; GaussianModel uses it to size new Gaussians
 * (gaussian_splatting/scene/gaussian_model.py:256-262).  Asynchronous, no host round trips.
 * ------------------------------------------------------------------------------------------- */
size_t ols_knn_workspace_size(int32_t P);
int ols_knn_mean_dist2(int32_t P, const float* d_points /* [P,3] */, float* d_mean_dist2 /* [P] */, void* d_workspace,
                       size_t workspace_bytes, void* stream);

/* Per-kernel device timing (CUDA events recorded between the kernels of every call while enabled).
 * ols_timing_begin() allocates `max_marks` events and enables recording on the calling thread;
 * ols_timing_end() synchronises, sums the elapsed milliseconds per tag, reports how many intervals
 * each tag saw, and disables recording.  The reference has no counterpart. */
#define OLS_TIMING_TAGS 8
enum { OLS_T_PREPROCESS = 0, OLS_T_BINNING = 1, OLS_T_SORT = 2, OLS_T_BLEND_FWD = 3, OLS_T_BLEND_BWD = 4,
       OLS_T_GEOMETRY_BWD = 5, OLS_T_AE = 6, OLS_T_OTHER = 7 };
int ols_timing_begin(int32_t max_marks);
int ols_timing_end(float* ms_per_tag /* [OLS_TIMING_TAGS] */, int32_t* count_per_tag /* [OLS_TIMING_TAGS] */);

/* ---------------------------------------------------------------------------------------------
 * Autoencoder (language/autoencoder/model.py:15-62 AutoencoderMLP, :314-354 EncoderDecoderOnline).
 * A "chain" is Linear -> [ReLU -> Linear]* with eval-mode BatchNorm already folded into the
 * preceding Linear by the caller, followed by an optional row-wise L2 normalisation.
 * Weights are row-major [out,in] float32 as torch.nn.Linear stores them.
 * ------------------------------------------------------------------------------------------- */
#define OLS_AE_MAX_LAYERS 8
#define OLS_AE_FAST 0
#define OLS_AE_FP32 1
typedef struct ols_ae_chain {
    int32_t n_layers;
    int32_t dims[OLS_AE_MAX_LAYERS + 1];   /* dims[0] = input width, dims[i+1] = output width of layer i */
    int32_t normalize;                     /* 1: y /= ||y||_2 per row after the last layer               */
    int32_t input_bf16;                    /* 1: x is bfloat16 [M, dims[0]] (dims[0] % 64 == 0): ols_ae_forward_bf16      */
    int32_t precision;                     /* OLS_AE_FAST (0): tensor cores, tf32 first layer + bf16 inner layers, fp32 accumulate;
                                              OLS_AE_FP32 (1): parity mode -- every layer in fp32 FMAs on the CUDA cores, the
                                              arithmetic of the reference's fp32 nn.Linear (model.py:52-62); ~10x slower  */
    int32_t _pad;
    const float* d_weight[OLS_AE_MAX_LAYERS];  /* [dims[i+1], dims[i]]                                  */
    const float* d_bias[OLS_AE_MAX_LAYERS];    /* [dims[i+1]]                                           */
} ols_ae_chain;

/* Opaque prepared chain: weights re-laid out for the tensor-core kernel (done once per weight update). */
typedef struct ols_ae_plan ols_ae_plan;
int ols_ae_plan_create(const ols_ae_chain* chain, ols_ae_plan** plan, void* stream);
void ols_ae_plan_destroy(ols_ae_plan* plan);
/* y[M, dims[n]] = chain(x[M, dims[0]]); replaces AutoencoderMLP.encode / .decode. */
int ols_ae_forward(const ols_ae_plan* plan, const float* d_x, float* d_y, int64_t M, void* stream);
/* the same chain on a bfloat16 input matrix (plan created with input_bf16 = 1): used to run the encoder directly on
 * the HR module's last 128-channel activation with final_conv folded into the first Linear (ols_hr_forward_features) */
int ols_ae_forward_bf16(const ols_ae_plan* plan, const void* d_x_bf16, float* d_y, int64_t M, void* stream);

/* ---------------------------------------------------------------------------------------------
 * One training step of the online autoencoder, fused (utils/slam_backend.py:266-323 train_online_autoencoder on
 * language/autoencoder/model.py:314-354 EncoderDecoderOnline, 32 -> 24 -> 15 -> 24 -> 32):
 *     comp = encode(x); recon = decode(comp);
 *     loss = l1_loss(recon, x) + 0.6 * (1 - cosine_similarity(recon, x, dim=1).mean());  loss.backward();  Adam.step()
 * d_params: the 2351 parameters as one flat fp32 vector in nn.Module.parameters() order (encoder.0.weight [24,32],
 * encoder.0.bias, encoder.2.weight [15,24], encoder.2.bias, decoder.0.weight [24,15], decoder.0.bias, decoder.2.weight
 * [32,24], decoder.2.bias), updated in place; d_exp_avg / d_exp_avg_sq: Adam moments (same layout, zero at the start);
 * d_step: device step counter (steps taken so far, incremented by the call).  d_code receives comp [M,15] computed with
 * the parameters BEFORE the update (what the reference returns, :323); d_loss the scalar loss.  d_scratch:
 * ols_online_ae_scratch_bytes() bytes, 256-byte aligned, zero-filled before the first call and then left alone.
 * Forward and backward are fp32 FMAs; the slab reduction order is fixed, so the step is deterministic.
 * ------------------------------------------------------------------------------------------- */
size_t ols_online_ae_scratch_bytes(void);
int32_t ols_online_ae_param_count(void);
int ols_online_ae_train_step(float* d_params, float* d_exp_avg, float* d_exp_avg_sq, int64_t* d_step, const float* d_x /* [M,32] */,
                             int64_t M, float lr, float beta1, float beta2, float eps, float* d_code /* [M,15] or NULL */,
                             float* d_loss /* [1] */, void* d_scratch, size_t scratch_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * HR module (language/supervisedNet.py:45-109 HighResLanguageFeatureNet, called with torch.no_grad() in eval
 * mode at utils/slam_backend.py:381-386,547-552): fv [768,S,S] + res3 [384,h3,w3] + res2 [192,h2,w2] ->
 * [768,8S,8S] dense CLIP map, the autoencoder's input.  13 convolutions in the order
 *   0 initial_conv.0        Conv2d 768->512 3x3        7 attention_fusion2.low_res_align  Conv2d 192->256 1x1
 *   1 upsample1.0           ConvT  512->512 4/2/1      8 attention_fusion2.fusion.0       Conv2d 512->256 3x3
 *   2 af1.low_res_align     Conv2d 384->512 1x1        9 attention_fusion2.attention.0    Conv2d 256->256 3x3
 *   3 af1.fusion.0          Conv2d 1024->512 3x3      10 attention_fusion2.attention.3    Conv2d 256->256 1x1
 *   4 af1.attention.0       Conv2d 512->512 3x3       11 upsample3.0                      ConvT  256->128 4/2/1
 *   5 af1.attention.3       Conv2d 512->512 1x1       12 final_conv                       Conv2d 128->768 1x1
 *   6 upsample2.0           ConvT  512->256 4/2/1
 * Weights are float32 in torch layout (Conv2d [Cout,Cin,kh,kw], ConvTranspose2d [Cin,Cout,4,4]) with the
 * eval-mode BatchNorm that follows a convolution already folded in by the caller (as for the autoencoder).
 * The plan owns bf16 re-laid-out weights, the bf16 NHWC activations between the layers, and a side stream on which the
 * two low_res_align convolutions run next to layers 0-1 (forked from / joined to the caller's stream with events, so
 * the call is still ordered on `stream` and can be captured into a CUDA graph).  A plan is not re-entrant: calls that
 * use the same plan must be ordered on one stream; use one plan per concurrent stream.
 * ------------------------------------------------------------------------------------------- */
#define OLS_HR_N_CONV 13
typedef struct ols_hr_weights {
    const float* d_weight[OLS_HR_N_CONV];
    const float* d_bias[OLS_HR_N_CONV];
} ols_hr_weights;
typedef struct ols_hr_plan ols_hr_plan;
/* S_h x S_w = spatial size of fv (24 x 24 in the reference) */
int ols_hr_plan_create(const ols_hr_weights* w, int32_t S_h, int32_t S_w, ols_hr_plan** plan, void* stream);
void ols_hr_plan_destroy(ols_hr_plan* plan);
/* d_out: float32 [8*S_h, 8*S_w, 768] (channels last: the layout `permute(0,2,3,1).view(-1,768)` at
 * slam_backend.py:392-394 produces, i.e. the autoencoder's [M,768] input).  f3 / f2 are resized to 2S / 4S with
 * bilinear interpolation, align_corners = False (supervisedNet.py:88,97). */
int ols_hr_forward(const ols_hr_plan* plan, const float* d_fv, const float* d_f3, int32_t h3, int32_t w3,
                   const float* d_f2, int32_t h2, int32_t w2, float* d_out, void* stream);
/* Layers 0..11 only: writes the last hidden activation (upsample3's output, bfloat16 [8*S_h * 8*S_w, 128], pixel-major)
 * into d_feat_bf16.  final_conv (a 1x1 convolution = a Linear over channels) and the autoencoder's first Linear have
 * nothing between them (utils/slam_backend.py:392-395), so the caller may fold them (W' = W_ae0 W_final,
 * b' = W_ae0 b_final + b_ae0) and feed d_feat_bf16 to ols_ae_forward_bf16: the 768-channel fp32 map is never written. */
int ols_hr_forward_features(const ols_hr_plan* plan, const float* d_fv, const float* d_f3, int32_t h3, int32_t w3,
                            const float* d_f2, int32_t h2, int32_t w2, void* d_feat_bf16, void* stream);
/* development aid: copy of an intermediate activation (bf16 NHWC) as float32; which = conv index 0..11 */
int ols_hr_read_activation(const ols_hr_plan* plan, int32_t which, float* d_out, int64_t capacity_floats, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SSIM and the colour-refinement loss (gaussian_splatting/utils/loss_utils.py:41-101 `ssim`, window 11, sigma 1.5,
 * zero padding, size_average=True; caller utils/slam_backend.py:797-801):
 *     value = w_l1 * mean|image - gt| + w_ssim * mean(ssim_map(image, gt))
 * (the caller's loss is  (1 - lambda) * l1 + lambda * (1 - ssim)  =  value + lambda  with w_l1 = 1 - lambda,
 * w_ssim = -lambda).  Forward stores three per-pixel partial-derivative maps that backward filters once.
 * ------------------------------------------------------------------------------------------- */
typedef struct ols_ssim_args {
    int32_t C, H, W;
    float w_l1, w_ssim;
    const float* d_image;   /* [C,H,W] render()["render"]        */
    const float* d_gt;      /* [C,H,W] viewpoint.original_image  */
} ols_ssim_args;
/* d_out4 = [l1, ssim, value, 0]; d_partial: [3,C,H,W] floats (may be NULL when no backward follows); d_scratch2: 2 floats */
int ols_ssim_loss_forward(const ols_ssim_args* args, float* d_out4, float* d_partial, float* d_scratch2, void* stream);
/* dL/dimage = *d_upstream * d value / d image */
int ols_ssim_loss_backward(const ols_ssim_args* args, const float* d_partial, const float* d_upstream, float* d_dL_dimage,
                           void* stream);

/* ---------------------------------------------------------------------------------------------
 * Densification bookkeeping over the flat per-Gaussian arrays.
 * ols_densify_stats: per view, right after backward (utils/slam_backend.py:417-428,719-728 +
 *   gaussian_model.py:965-969): for radii > 0:  max_radii2D = max(max_radii2D, radii);
 *   xyz_gradient_accum += |viewspace_grad[:, :2]|;  denom += 1.   d_viewspace_grad may be NULL (colour refinement,
 *   slam_backend.py:807-812, only updates max_radii2D).  No host synchronisation.
 * ols_densify_flags: the selection masks of densify_and_prune (gaussian_model.py:948-963, :855-866, :912-921) on the
 *   current state: bit 0 clone, bit 1 split, bit 2 prune; d_counts3 = how many of each.
 * ------------------------------------------------------------------------------------------- */
typedef struct ols_densify_params {
    float max_grad;         /* grad_threshold                                              */
    float min_opacity;
    float extent;           /* scene extent                                                */
    float max_screen_size;  /* <= 0: None (no screen-size / world-size pruning)            */
    float percent_dense;    /* GaussianModel.percent_dense                                 */
} ols_densify_params;
int ols_densify_stats(int32_t P, const int32_t* d_radii, const float* d_viewspace_grad /* [P,3] or NULL */,
                      float* d_max_radii2D /* [P] */, float* d_xyz_gradient_accum /* [P,1] */, float* d_denom /* [P,1] */,
                      void* stream);
int ols_densify_flags(int32_t P, int32_t scale_cols /* 1 or 3 */, const float* d_xyz_gradient_accum, const float* d_denom,
                      const float* d_scaling_raw /* [P,scale_cols] log-scales */, const float* d_opacity_raw /* [P,1] logits */,
                      const float* d_max_radii2D, const ols_densify_params* params, uint8_t* d_flags /* [P] */,
                      int32_t* d_counts3, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OLS_B200_H_ */
