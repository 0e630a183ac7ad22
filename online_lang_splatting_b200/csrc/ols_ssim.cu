// ols_ssim.cu -- SSIM and the colour-refinement loss, forward and backward, fused ("SSIM for colour
// refinement").
//
// Reference: gaussian_splatting/utils/loss_utils.py:41-101 (gaussian / create_window / ssim / _ssim) and its one hot
// caller, the 26,000 iterations of utils/slam_backend.py:777-817:
//     loss = (1 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))
// The reference runs five depthwise 11x11 cuDNN convolutions (mu1, mu2, E[x^2], E[y^2], E[xy]), ~15 elementwise
// kernels and the whole thing again, transposed, in autograd.  Here:
//   k_ssim_fwd   one pass: halo tile of both images in shared memory, separable 11-tap Gaussian (rows then columns)
//                for the five moments, SSIM per pixel, the |x - y| term, both sums; it also stores the three partial
//                derivatives of the SSIM map w.r.t. (mu1, E[x^2], E[xy]) that the backward needs;
//   k_ssim_bwd   one pass: the same separable filter over those three maps (the window is symmetric, so the
//                transposed convolution is the convolution),  dL/dx = w_ssim * (G*d_mu1 + 2x G*d_ex2 + y G*d_exy)
//                + w_l1 * sign(x - y).
// Zero padding exactly as F.conv2d(padding=5): out-of-image taps contribute 0 (to the images and to the partial maps).
// Pure streaming kernels: 2 images read, 3 maps written (forward); 3 maps + 2 images read, 1 written (backward).
#include "ols_common.cuh"

namespace ols {

constexpr int SS_T = 16;            // output tile
constexpr int SS_R = 5;             // window radius (window_size 11)
constexpr int SS_H = SS_T + 2 * SS_R;  // 26: halo tile edge
constexpr float SS_C1 = 0.01f * 0.01f, SS_C2 = 0.03f * 0.03f;

struct SsimArgs {
    int C, H, W;
    float win[11];      // gaussian(11, 1.5) normalised, computed on the host exactly like loss_utils.py:41-48
    const float *x, *y; // image, ground truth [C,H,W]
    float* partial;     // [3,C,H,W]: d ssim / d mu1, d E[x^2], d E[xy]
    float* sums;        // [0] sum ssim_map, [1] sum |x - y|
    const float* upstream;
    float w_l1, w_ssim; // backward: dL/dx = upstream * (w_l1 * sign(x - y) + w_ssim * d(sum ssim)/dx)
    float* dx;
};

__global__ void __launch_bounds__(SS_T * SS_T) k_ssim_fwd(const SsimArgs a) {
    __shared__ float sx[SS_H][SS_H + 1], sy[SS_H][SS_H + 1];
    __shared__ float hz[5][SS_H][SS_T + 1];
    __shared__ float red[2][SS_T * SS_T / 32];
    const int c = blockIdx.z, x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    const size_t plane = (size_t)a.H * a.W;
    const float* px = a.x + c * plane;
    const float* py = a.y + c * plane;
    const int tid = threadIdx.y * SS_T + threadIdx.x;
    for (int i = tid; i < SS_H * SS_H; i += SS_T * SS_T) {
        const int r = i / SS_H, q = i - r * SS_H;
        const int gy = y0 + r - SS_R, gx = x0 + q - SS_R;
        const bool in = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
        sx[r][q] = in ? px[(size_t)gy * a.W + gx] : 0.0f;
        sy[r][q] = in ? py[(size_t)gy * a.W + gx] : 0.0f;
    }
    __syncthreads();
    // rows: 26 x 16 positions, five moments each
    for (int i = tid; i < SS_H * SS_T; i += SS_T * SS_T) {
        const int r = i / SS_T, q = i - r * SS_T;
        float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float w = a.win[k], u = sx[r][q + k], v = sy[r][q + k];
            m1 = fmaf(w, u, m1); m2 = fmaf(w, v, m2);
            e11 = fmaf(w, u * u, e11); e22 = fmaf(w, v * v, e22); e12 = fmaf(w, u * v, e12);
        }
        hz[0][r][q] = m1; hz[1][r][q] = m2; hz[2][r][q] = e11; hz[3][r][q] = e22; hz[4][r][q] = e12;
    }
    __syncthreads();
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int gx = x0 + tx, gy = y0 + ty;
    float ssim = 0.0f, l1 = 0.0f;
    if (gx < a.W && gy < a.H) {
        float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float w = a.win[k];
            mu1 = fmaf(w, hz[0][ty + k][tx], mu1); mu2 = fmaf(w, hz[1][ty + k][tx], mu2);
            e11 = fmaf(w, hz[2][ty + k][tx], e11); e22 = fmaf(w, hz[3][ty + k][tx], e22);
            e12 = fmaf(w, hz[4][ty + k][tx], e12);
        }
        // loss_utils.py:76-96
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = e11 - mu1_sq, s2 = e22 - mu2_sq, s12 = e12 - mu12;
        const float A = 2.0f * mu12 + SS_C1, B = 2.0f * s12 + SS_C2, Cc = mu1_sq + mu2_sq + SS_C1, D = s1 + s2 + SS_C2;
        const float inv_cd = 1.0f / (Cc * D);
        ssim = A * B * inv_cd;
        l1 = fabsf(sx[ty + SS_R][tx + SS_R] - sy[ty + SS_R][tx + SS_R]);
        if (a.partial) {
            // partial derivatives with (mu1, E[x^2], E[xy]) as the independent filter outputs:
            // s1 = E[x^2] - mu1^2, s12 = E[xy] - mu1 mu2
            const size_t o = c * plane + (size_t)gy * a.W + gx, cs = (size_t)a.C * plane;
            const float d_mu1 = (2.0f * mu2 * B - 2.0f * mu2 * A) * inv_cd - ssim * (2.0f * mu1 / Cc) + ssim * (2.0f * mu1 / D);
            a.partial[o] = d_mu1;
            a.partial[cs + o] = -ssim / D;
            a.partial[2 * cs + o] = 2.0f * A * inv_cd;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ssim += __shfl_xor_sync(0xffffffffu, ssim, o);
        l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = ssim; red[1][tid >> 5] = l1; }
    __syncthreads();
    if (tid < 2) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < SS_T * SS_T / 32; w++) t += red[tid][w];
        atomicAdd(&a.sums[tid], t);
    }
}

// out4 = [l1 mean, ssim mean, w_l1 * l1 + w_ssim * ssim, 0]
__global__ void k_ssim_finish(const SsimArgs a, float* out) {
    const float n = (float)a.C * (float)a.H * (float)a.W;
    const float ssim = a.sums[0] / n, l1 = a.sums[1] / n;
    out[0] = l1;
    out[1] = ssim;
    out[2] = a.w_l1 * l1 + a.w_ssim * ssim;
    out[3] = 0.0f;
}

__global__ void __launch_bounds__(SS_T * SS_T) k_ssim_bwd(const SsimArgs a) {
    __shared__ float sp[3][SS_H][SS_H + 1];
    __shared__ float hz[3][SS_H][SS_T + 1];
    const int c = blockIdx.z, x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    const size_t plane = (size_t)a.H * a.W, cs = (size_t)a.C * plane;
    const int tid = threadIdx.y * SS_T + threadIdx.x;
    for (int i = tid; i < SS_H * SS_H; i += SS_T * SS_T) {
        const int r = i / SS_H, q = i - r * SS_H;
        const int gy = y0 + r - SS_R, gx = x0 + q - SS_R;
        const bool in = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
        const size_t o = c * plane + (size_t)gy * a.W + gx;
#pragma unroll
        for (int m = 0; m < 3; m++) sp[m][r][q] = in ? a.partial[m * cs + o] : 0.0f;
    }
    __syncthreads();
    for (int i = tid; i < SS_H * SS_T; i += SS_T * SS_T) {
        const int r = i / SS_T, q = i - r * SS_T;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float w = a.win[k];
            s0 = fmaf(w, sp[0][r][q + k], s0); s1 = fmaf(w, sp[1][r][q + k], s1); s2 = fmaf(w, sp[2][r][q + k], s2);
        }
        hz[0][r][q] = s0; hz[1][r][q] = s1; hz[2][r][q] = s2;
    }
    __syncthreads();
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx < a.W && gy < a.H) {
        float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float w = a.win[k];
            g0 = fmaf(w, hz[0][ty + k][tx], g0); g1 = fmaf(w, hz[1][ty + k][tx], g1); g2 = fmaf(w, hz[2][ty + k][tx], g2);
        }
        const size_t o = c * plane + (size_t)gy * a.W + gx;
        const float x = a.x[o], y = a.y[o];
        const float n = (float)a.C * (float)a.H * (float)a.W;
        const float d = x - y;
        const float sg = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f);
        a.dx[o] = a.upstream[0] * (a.w_l1 * sg + a.w_ssim * (g0 + 2.0f * x * g1 + y * g2)) / n;
    }
}

}  // namespace ols

using namespace ols;

static int ssim_fill(const ols_ssim_args* p, SsimArgs* a) {
    if (!p || p->C <= 0 || p->H <= 0 || p->W <= 0 || !p->d_image || !p->d_gt) { ols_set_error("bad SSIM arguments"); return OLS_ERR_INVALID; }
    a->C = p->C; a->H = p->H; a->W = p->W;
    // gaussian(window_size = 11, sigma = 1.5), loss_utils.py:41-48: float32 tensor of exp(...), divided by its float32 sum
    float g[11], sum = 0.0f;
    for (int i = 0; i < 11; i++) { g[i] = (float)exp(-(double)((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5)); sum += g[i]; }
    for (int i = 0; i < 11; i++) a->win[i] = g[i] / sum;
    a->x = p->d_image; a->y = p->d_gt;
    a->w_l1 = p->w_l1; a->w_ssim = p->w_ssim;
    a->partial = nullptr; a->sums = nullptr; a->upstream = nullptr; a->dx = nullptr;
    return OLS_OK;
}

extern "C" {

int ols_ssim_loss_forward(const ols_ssim_args* p, float* d_out4, float* d_partial, float* d_scratch2, void* stream) {
    SsimArgs a;
    int rc = ssim_fill(p, &a);
    if (rc != OLS_OK) return rc;
    if (!d_out4 || !d_scratch2) { ols_set_error("null output pointer"); return OLS_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    a.partial = d_partial; a.sums = d_scratch2;
    OLS_CUDA_TRY(cudaMemsetAsync(d_scratch2, 0, 2 * sizeof(float), st));
    const dim3 grid((p->W + SS_T - 1) / SS_T, (p->H + SS_T - 1) / SS_T, p->C);
    k_ssim_fwd<<<grid, dim3(SS_T, SS_T), 0, st>>>(a);
    k_ssim_finish<<<1, 1, 0, st>>>(a, d_out4);
    OLS_CUDA_TRY(cudaGetLastError());
    return OLS_OK;
}

int ols_ssim_loss_backward(const ols_ssim_args* p, const float* d_partial, const float* d_upstream, float* d_dL_dimage,
                           void* stream) {
    SsimArgs a;
    int rc = ssim_fill(p, &a);
    if (rc != OLS_OK) return rc;
    if (!d_partial || !d_upstream || !d_dL_dimage) { ols_set_error("null gradient pointer"); return OLS_ERR_INVALID; }
    a.partial = const_cast<float*>(d_partial); a.upstream = d_upstream; a.dx = d_dL_dimage;
    const dim3 grid((p->W + SS_T - 1) / SS_T, (p->H + SS_T - 1) / SS_T, p->C);
    k_ssim_bwd<<<grid, dim3(SS_T, SS_T), 0, (cudaStream_t)stream>>>(a);
    OLS_CUDA_TRY(cudaGetLastError());
    return OLS_OK;
}

}  // extern "C"
