// ols_loss.cu -- the mapping loss that sits directly on the rasterizer's outputs, fused (SURVEY 8f N2 + N3).
//
// Reference (per view of a mapping iteration, utils/slam_backend.py:576-592 and utils/slam_utils.py:121-165):
//   image_ab  = exp(exposure_a) * image + exposure_b                      (get_loss_mapping, :121-126)
//   l1_rgb    = mean | image_ab * m_rgb - gt_image * m_rgb |,  m_rgb = (sum_c gt_image > rgb_boundary_threshold)
//   l1_depth  = mean | depth * m_d - gt_depth * m_d |,         m_d   = (gt_depth > 0.01)
//   l1_lang   = mean | language - bilinear(gt_lang_feat [F,h,w] -> [F,H,W], align_corners=False) |
//   loss      = alpha * l1_rgb + (1 - alpha) * l1_depth + lambda_lang * l1_lang
// The reference runs this as ~25 elementwise / reduction / interpolate kernels through autograd and copies
// the up-sampled 15 x H x W feature map over PCIe every iteration.  Here the low
// resolution feature map stays on the device and two kernels do everything:
//   k_mapping_loss_fwd   one pass over the 3 + 1 + F rendered channels -> the three sums (+ exposure sums)
//   k_mapping_loss_bwd   one pass -> dL/dimage, dL/ddepth, dL/dlanguage scaled by the upstream gradient
// Both are pure HBM streams: (3+1+F) * 4 B/px rendered + 16 B/px ground truth read, (3+1+F) * 4 B/px written.
#include "ols_common.cuh"

namespace ols {

struct LossArgs {
    int W, H, F, lw, lh;
    float alpha, thr, ea, eb, lambda_lang;  // ea = exp(exposure_a)
    float sx, sy;                           // lw / W, lh / H as PyTorch computes them (float division)
    const float *image, *depth, *language, *gt_image, *gt_depth, *gt_lang;
    const float *opacity, *grad_mask;       // tracking form only (get_loss_tracking_rgbd), else NULL
    float* dopacity;                        // optional dL/dopacity of the tracking form
    const float* upstream;                  // device scalar dL/dloss (backward)
    float *dimage, *ddepth, *dlanguage;
    float* sums;                            // [8]: |rgb|, |depth|, |lang|, d/d exposure_a, d/d exposure_b
    const float *d_ea, *d_eb;               // device-resident exposure_a / exposure_b (override ea / eb when non-NULL)
};

// PyTorch upsample_bilinear2d, align_corners = false (ATen/native/UpSample.h: area_pixel_compute_source_index)
__device__ __forceinline__ void bilinear_tap(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.0f ? 0.0f : src;
    i0 = (int)src;
    i0 = i0 > in_size - 1 ? in_size - 1 : i0;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = src - (float)i0;
}

__device__ __forceinline__ float sgn(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }

constexpr int LOSS_THREADS = 256;

// blockIdx.y splits the channels: part 0 = colour + depth (+ exposure / opacity terms), part p >= 1 = language channels
// [(p-1) * LOSS_LCH, p * LOSS_LCH).  Four times as many, lighter threads: the kernel is latency bound (19 read + 19 written
// planes, 60 gathered taps per pixel), not bandwidth bound, at one thread per pixel.
constexpr int LOSS_LCH = 5;
template <bool BACKWARD>
__global__ void __launch_bounds__(LOSS_THREADS) k_mapping_loss(const LossArgs a) {
    const int part = blockIdx.y;
    const int c_lo = part == 0 ? 0 : (part - 1) * LOSS_LCH, c_hi = part == 0 ? 0 : min(a.F, part * LOSS_LCH);
    const size_t HW = (size_t)a.W * a.H;
    float s_rgb = 0.0f, s_d = 0.0f, s_l = 0.0f, s_ea = 0.0f, s_eb = 0.0f;
    const float up = BACKWARD ? a.upstream[0] : 0.0f;
    const float w_rgb = a.alpha / (3.0f * (float)HW), w_d = (1.0f - a.alpha) / (float)HW;
    const float w_l = a.F > 0 ? a.lambda_lang / ((float)a.F * (float)HW) : 0.0f;
    const float ea = a.d_ea ? expf(a.d_ea[0]) : a.ea, eb = a.d_eb ? a.d_eb[0] : a.eb;
    for (size_t pix = (size_t)blockIdx.x * LOSS_THREADS + threadIdx.x; pix < HW; pix += (size_t)gridDim.x * LOSS_THREADS) {
        const int y = (int)((unsigned)pix / (unsigned)a.W), x = (int)pix - y * a.W;
        if (part == 0) {
        // colour: | (ea * image + eb) * m - gt * m |
        const float g0 = a.gt_image[pix], g1 = a.gt_image[HW + pix], g2 = a.gt_image[2 * HW + pix];
        float m = (g0 + g1 + g2) > a.thr ? 1.0f : 0.0f;
        // tracking form (utils/slam_utils.py:96-118): the colour mask also carries viewpoint.grad_mask, every colour
        // residual is weighted by the rendered opacity, and depth only counts where opacity > 0.95
        const float op = a.opacity ? a.opacity[pix] : 1.0f;
        if (a.grad_mask) m *= a.grad_mask[pix];
        const float gts[3] = {g0, g1, g2};
        float abs_sum = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float im = a.image[c * HW + pix];
            const float diff = (ea * im + eb) * m - gts[c] * m;
            abs_sum += fabsf(diff);
            if (BACKWARD) {
                a.dimage[c * HW + pix] = up * w_rgb * op * sgn(diff) * m * ea;
            } else {
                s_rgb += op * fabsf(diff);
                const float sg = op * sgn(diff) * m;
                s_ea += sg * ea * im;  // d|diff| / d exposure_a
                s_eb += sg;              // d|diff| / d exposure_b
            }
        }
        // depth
        {
            const float gd = a.gt_depth[pix];
            float md = gd > 0.01f ? 1.0f : 0.0f;
            if (a.opacity && !(op > 0.95f)) md = 0.0f;
            if (BACKWARD && a.dopacity) a.dopacity[pix] = up * w_rgb * abs_sum;
            const float diff = a.depth[pix] * md - gd * md;
            if (BACKWARD) a.ddepth[pix] = up * w_d * sgn(diff) * md;
            else s_d += fabsf(diff);
        }
        }
        // language: target = bilinear up-sampling of the low-resolution code map
        if (part > 0) {
            int x0, x1, y0, y1;
            float lx, ly;
            bilinear_tap(x, a.sx, a.lw, x0, x1, lx);
            bilinear_tap(y, a.sy, a.lh, y0, y1, ly);
            const float hx = 1.0f - lx, hy = 1.0f - ly;
            const size_t lhw = (size_t)a.lw * a.lh;
            for (int c = c_lo; c < c_hi; c++) {
                const float* src = a.gt_lang + c * lhw;
                const float t = hy * (hx * src[y0 * a.lw + x0] + lx * src[y0 * a.lw + x1]) +
                                ly * (hx * src[y1 * a.lw + x0] + lx * src[y1 * a.lw + x1]);
                const float diff = a.language[c * HW + pix] - t;
                if (BACKWARD) a.dlanguage[c * HW + pix] = up * w_l * sgn(diff);
                else s_l += fabsf(diff);
            }
        }
    }
    if (!BACKWARD) {
        __shared__ float red[5][LOSS_THREADS / 32];
        float v[5] = {s_rgb, s_d, s_l, s_ea, s_eb};
#pragma unroll
        for (int k = 0; k < 5; k++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
            if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
        }
        __syncthreads();
        if (threadIdx.x < 5) {
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < LOSS_THREADS / 32; w++) t += red[threadIdx.x][w];
            if (t != 0.0f) atomicAdd(&a.sums[threadIdx.x], t);
        }
    }
}

// sums -> [l1_rgb, l1_depth, l1_lang, dloss/dexposure_a, dloss/dexposure_b, loss]
__global__ void k_mapping_loss_finish(LossArgs a, float* out) {
    const float HW = (float)a.W * (float)a.H;
    const float l_rgb = a.sums[0] / (3.0f * HW), l_d = a.sums[1] / HW;
    const float l_l = a.F > 0 ? a.sums[2] / ((float)a.F * HW) : 0.0f;
    out[0] = l_rgb;
    out[1] = l_d;
    out[2] = l_l;
    out[3] = a.alpha * a.sums[3] / (3.0f * HW);
    out[4] = a.alpha * a.sums[4] / (3.0f * HW);
    out[5] = a.alpha * l_rgb + (1.0f - a.alpha) * l_d + a.lambda_lang * l_l;
}

}  // namespace ols

using namespace ols;

static int fill(const ols_loss_args* p, LossArgs* a) {
    if (!p || p->W <= 0 || p->H <= 0 || p->F < 0) { ols_set_error("bad loss arguments"); return OLS_ERR_INVALID; }
    if (!p->d_image || !p->d_depth || !p->d_gt_image || !p->d_gt_depth) { ols_set_error("null image pointer"); return OLS_ERR_INVALID; }
    if (p->F > 0 && (!p->d_language || !p->d_gt_lang || p->lang_w <= 0 || p->lang_h <= 0)) {
        ols_set_error("language term needs language, gt_lang and its size");
        return OLS_ERR_INVALID;
    }
    a->W = p->W; a->H = p->H; a->F = p->F; a->lw = p->lang_w; a->lh = p->lang_h;
    a->alpha = p->alpha; a->thr = p->rgb_boundary_threshold; a->ea = expf(p->exposure_a); a->eb = p->exposure_b;
    a->lambda_lang = p->lambda_lang;
    a->sx = p->F > 0 ? (float)p->lang_w / (float)p->W : 0.0f;
    a->sy = p->F > 0 ? (float)p->lang_h / (float)p->H : 0.0f;
    a->image = p->d_image; a->depth = p->d_depth; a->language = p->d_language;
    a->gt_image = p->d_gt_image; a->gt_depth = p->d_gt_depth; a->gt_lang = p->d_gt_lang;
    a->opacity = p->d_opacity; a->grad_mask = p->d_grad_mask; a->dopacity = nullptr;
    a->d_ea = p->d_exposure_a; a->d_eb = p->d_exposure_b;
    a->upstream = nullptr; a->dimage = nullptr; a->ddepth = nullptr; a->dlanguage = nullptr; a->sums = nullptr;
    return OLS_OK;
}

static int loss_grid(int W, int H) {
    const long long blocks = ((long long)W * H + LOSS_THREADS - 1) / LOSS_THREADS;
    return (int)(blocks < 148 * 8 ? blocks : 148 * 8);  // 8 CTAs of 256 threads per SM
}

extern "C" {

int ols_mapping_loss_forward(const ols_loss_args* p, float* d_out6, float* d_scratch8, void* stream) {
    LossArgs a;
    int rc = fill(p, &a);
    if (rc != OLS_OK) return rc;
    if (!d_out6 || !d_scratch8) { ols_set_error("null output pointer"); return OLS_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    a.sums = d_scratch8;
    OLS_CUDA_TRY(cudaMemsetAsync(d_scratch8, 0, 8 * sizeof(float), st));
    k_mapping_loss<false><<<dim3(loss_grid(p->W, p->H), 1 + (p->F + LOSS_LCH - 1) / LOSS_LCH), LOSS_THREADS, 0, st>>>(a);
    k_mapping_loss_finish<<<1, 1, 0, st>>>(a, d_out6);
    OLS_CUDA_TRY(cudaGetLastError());
    return OLS_OK;
}

int ols_mapping_loss_backward(const ols_loss_args* p, const float* d_upstream, float* d_dL_dimage, float* d_dL_ddepth,
                              float* d_dL_dlanguage, float* d_dL_dopacity, void* stream) {
    LossArgs a;
    int rc = fill(p, &a);
    if (rc != OLS_OK) return rc;
    if (!d_upstream || !d_dL_dimage || !d_dL_ddepth || (p->F > 0 && !d_dL_dlanguage)) {
        ols_set_error("null gradient pointer");
        return OLS_ERR_INVALID;
    }
    a.upstream = d_upstream; a.dimage = d_dL_dimage; a.ddepth = d_dL_ddepth; a.dlanguage = d_dL_dlanguage;
    a.dopacity = p->d_opacity ? d_dL_dopacity : nullptr;
    k_mapping_loss<true><<<dim3(loss_grid(p->W, p->H), 1 + (p->F + LOSS_LCH - 1) / LOSS_LCH), LOSS_THREADS, 0, (cudaStream_t)stream>>>(a);
    OLS_CUDA_TRY(cudaGetLastError());
    return OLS_OK;
}

}  // extern "C"
