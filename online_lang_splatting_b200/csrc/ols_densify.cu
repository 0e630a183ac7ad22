// ols_densify.cu -- per-view densification statistics and the densify / prune decision flags over the flat per-Gaussian
// arrays ("densification stats").
//
// Reference, per view of every mapping iteration (utils/slam_backend.py:417-428,719-728):
//     max_radii2D[visibility_filter] = max(max_radii2D[visibility_filter], radii[visibility_filter])
//     gaussians.add_densification_stats(viewspace_point_tensor, visibility_filter)        (gaussian_model.py:965-969)
//         xyz_gradient_accum[filter] += norm(viewspace_point_tensor.grad[filter, :2], dim=-1, keepdim=True)
//         denom[filter] += 1
// with visibility_filter = radii > 0: eight boolean-mask indexing kernels, each with a nonzero() host round trip.
// Here one streaming kernel, no host synchronisation (k_densify_stats).
//
// And the selection masks of densify_and_prune (gaussian_model.py:948-963 with :855-866, :912-921) evaluated on the
// current state in one pass (k_densify_flags): bit 0 clone, bit 1 split, bit 2 prune, plus their three counts, so
// that the host sizes the gather / concatenation (which stays in torch: control plane) without a device scan.
#include "ols_common.cuh"

namespace ols {

__global__ void __launch_bounds__(256) k_densify_stats(int P, const int* __restrict__ radii, const float* __restrict__ vgrad,
                                                       float* __restrict__ max_radii, float* __restrict__ accum,
                                                       float* __restrict__ denom) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const int r = radii[i];
    if (r <= 0) return;  // visibility_filter = radii > 0
    max_radii[i] = fmaxf(max_radii[i], (float)r);
    if (vgrad) {
        const float gx = vgrad[3 * (size_t)i], gy = vgrad[3 * (size_t)i + 1];
        accum[i] += sqrtf(gx * gx + gy * gy);
        denom[i] += 1.0f;
    }
}

struct FlagArgs {
    int P, scale_cols;
    const float *accum, *denom, *scaling, *opacity, *max_radii;
    float grad_threshold, dense_extent /* percent_dense * scene_extent */, min_opacity, max_screen_size, big_ws /* 0.1 * extent */;
    unsigned char* flags;
    int* counts;
};

__global__ void __launch_bounds__(256) k_densify_flags(const FlagArgs a) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    unsigned f = 0;
    if (i < a.P) {
        // grads = xyz_gradient_accum / denom; grads[isnan] = 0
        float g = a.accum[i] / a.denom[i];
        if (isnan(g)) g = 0.0f;
        // torch.max(get_scaling, dim=1): scaling_activation = exp
        float smax = expf(a.scaling[(size_t)i * a.scale_cols]);
        for (int k = 1; k < a.scale_cols; k++) smax = fmaxf(smax, expf(a.scaling[(size_t)i * a.scale_cols + k]));
        const bool hot = g >= a.grad_threshold;
        if (hot && smax <= a.dense_extent) f |= 1u;   // densify_and_clone
        if (hot && smax > a.dense_extent) f |= 2u;    // densify_and_split
        const float op = 1.0f / (1.0f + expf(-a.opacity[i]));  // opacity_activation = sigmoid
        bool prune = op < a.min_opacity;
        if (a.max_screen_size > 0.0f) prune = prune || a.max_radii[i] > a.max_screen_size || smax > a.big_ws;
        if (prune) f |= 4u;
        a.flags[i] = (unsigned char)f;
    }
    // three counts: ballot per warp, one atomic per warp and flag
#pragma unroll
    for (int b = 0; b < 3; b++) {
        const unsigned m = __ballot_sync(0xffffffffu, (f >> b) & 1u);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&a.counts[b], __popc(m));
    }
}

}  // namespace ols

using namespace ols;

extern "C" {

int ols_densify_stats(int32_t P, const int32_t* d_radii, const float* d_viewspace_grad, float* d_max_radii2D,
                      float* d_xyz_gradient_accum, float* d_denom, void* stream) {
    if (P < 0 || (P > 0 && (!d_radii || !d_max_radii2D))) { ols_set_error("bad densification-stat arguments"); return OLS_ERR_INVALID; }
    if (d_viewspace_grad && (!d_xyz_gradient_accum || !d_denom)) { ols_set_error("gradient given without accumulators"); return OLS_ERR_INVALID; }
    if (P == 0) return OLS_OK;
    k_densify_stats<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, d_radii, d_viewspace_grad, d_max_radii2D,
                                                                      d_xyz_gradient_accum, d_denom);
    OLS_CUDA_TRY(cudaGetLastError());
    return OLS_OK;
}

int ols_densify_flags(int32_t P, int32_t scale_cols, const float* d_xyz_gradient_accum, const float* d_denom,
                      const float* d_scaling_raw, const float* d_opacity_raw, const float* d_max_radii2D,
                      const ols_densify_params* prm, uint8_t* d_flags, int32_t* d_counts3, void* stream) {
    if (P < 0 || !prm || (scale_cols != 1 && scale_cols != 3) ||
        (P > 0 && (!d_xyz_gradient_accum || !d_denom || !d_scaling_raw || !d_opacity_raw || !d_max_radii2D || !d_flags)) || !d_counts3) {
        ols_set_error("bad densification-flag arguments");
        return OLS_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    OLS_CUDA_TRY(cudaMemsetAsync(d_counts3, 0, 3 * sizeof(int32_t), st));
    if (P == 0) return OLS_OK;
    FlagArgs a;
    a.P = P; a.scale_cols = scale_cols;
    a.accum = d_xyz_gradient_accum; a.denom = d_denom; a.scaling = d_scaling_raw; a.opacity = d_opacity_raw; a.max_radii = d_max_radii2D;
    a.grad_threshold = prm->max_grad; a.dense_extent = prm->percent_dense * prm->extent; a.min_opacity = prm->min_opacity;
    a.max_screen_size = prm->max_screen_size; a.big_ws = 0.1f * prm->extent;
    a.flags = d_flags; a.counts = d_counts3;
    k_densify_flags<<<(P + 255) / 256, 256, 0, st>>>(a);
    OLS_CUDA_TRY(cudaGetLastError());
    return OLS_OK;
}

}  // extern "C"
