// ols_optim.cu -- the optimiser step that follows the backward pass, fused over the flat buffers.
//
// Reference: GaussianModel.training_setup builds torch.optim.Adam(param_groups, lr=0.0, eps=1e-15) with one group
// per parameter tensor (xyz, f_dc, f_rest, opacity, scaling, rotation, f_language), each with its own learning rate
// (gaussian_splatting/scene/gaussian_model.py:393-437); optimizer.step() then runs torch's multi-tensor Adam.
// Here parameters, gradients and both moments live in flat fp32 buffers with the layout of
// sharding.FlatGradBuffer (the buffer the backward writes and NCCL all-reduces), so the whole step is ONE
// streaming kernel: 16 B read + 12 B written per element, no per-group launches.
// Update rule = torch.optim.Adam defaults (amsgrad=False, weight_decay=0, maximize=False):
//   m <- m + (g - m) (1 - b1);  v <- b2 v + (1 - b2) g g;  p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "ols_common.cuh"

namespace ols {

struct AdamArgs {
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;
    int n_groups;
    long long begin[OLS_ADAM_MAX_GROUPS + 1];  // group k covers [begin[k], begin[k+1])
    float step_size[OLS_ADAM_MAX_GROUPS];      // lr_k / (1 - b1^t)
    float step_tail[OLS_ADAM_MAX_GROUPS];      // lr_tail_k / (1 - b1^t)
    int activation[OLS_ADAM_MAX_GROUPS], period[OLS_ADAM_MAX_GROUPS], head[OLS_ADAM_MAX_GROUPS];
    float b1, b2, one_m_b1, one_m_b2, eps, inv_sqrt_bc2;  // 1 - beta computed in double on the host, like torch
    // device-resident step count (capturable form): step_size[] / step_tail[] then hold the bare learning rates and the
    // bias corrections are computed from *d_step + 1 by every thread (two double pows per thread, once)
    const long long* d_step;
    double beta1, beta2;
};

__global__ void k_adam_inc(long long* d_step) { *d_step += 1; }

__global__ void __launch_bounds__(256) k_adam(const AdamArgs a) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    float inv_bc1 = 1.0f, inv_sqrt_bc2 = a.inv_sqrt_bc2;
    if (a.d_step) {
        const double t = (double)(*a.d_step + 1);
        inv_bc1 = (float)(1.0 / (1.0 - pow(a.beta1, t)));
        inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow(a.beta2, t)));
    }
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < a.n; i += stride) {
        float p[4], g[4], m[4], v[4];
        const bool full = i + 4 <= a.n;
        if (full) {
            const float4 p4 = *reinterpret_cast<const float4*>(a.p + i), g4 = *reinterpret_cast<const float4*>(a.g + i);
            const float4 m4 = *reinterpret_cast<const float4*>(a.m + i), v4 = *reinterpret_cast<const float4*>(a.v + i);
            p[0] = p4.x; p[1] = p4.y; p[2] = p4.z; p[3] = p4.w; g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w;
            m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w; v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w;
        } else {
            for (int k = 0; k < 4; k++) {
                const bool ok = i + k < a.n;
                p[k] = ok ? a.p[i + k] : 0.f; g[k] = ok ? a.g[i + k] : 0.f; m[k] = ok ? a.m[i + k] : 0.f; v[k] = ok ? a.v[i + k] : 0.f;
            }
        }
        // F.normalize row of 4 (only meaningful inside an OLS_ACT_NORMALIZE4 group): taken before p is updated
        const float q_old[4] = {p[0], p[1], p[2], p[3]};
        const float q_inv = 1.0f / fmaxf(sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + p[3] * p[3]), 1e-12f);
        const float q_dot = (p[0] * g[0] + p[1] * g[1] + p[2] * g[2] + p[3] * g[3]) * q_inv;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const long long e = i + k;
            int grp = 0;
#pragma unroll
            for (int q = 1; q < OLS_ADAM_MAX_GROUPS; q++) grp += (q < a.n_groups && e >= a.begin[q]) ? 1 : 0;
            float step = a.step_size[grp];
            if (a.period[grp] > 0 && (int)((e - a.begin[grp]) % a.period[grp]) >= a.head[grp]) step = a.step_tail[grp];
            step *= inv_bc1;
            float gk = g[k];
            const int act = a.activation[grp];
            if (act == OLS_ACT_EXP) {
                gk = gk * expf(p[k]);
            } else if (act == OLS_ACT_SIGMOID) {
                const float o = 1.0f / (1.0f + expf(-p[k]));
                gk = gk * o * (1.0f - o);
            } else if (act == OLS_ACT_NORMALIZE4) {
                // rows of 4 coincide with this thread's chunk (group offset and count are multiples of 4, validated)
                gk = (gk - q_old[k] * q_inv * q_dot) * q_inv;
            }
            m[k] = m[k] + (gk - m[k]) * a.one_m_b1;
            v[k] = v[k] * a.b2 + a.one_m_b2 * gk * gk;
            const float denom = sqrtf(v[k]) * inv_sqrt_bc2 + a.eps;
            p[k] = p[k] - step * (m[k] / denom);
        }
        if (full) {
            *reinterpret_cast<float4*>(a.p + i) = make_float4(p[0], p[1], p[2], p[3]);
            *reinterpret_cast<float4*>(a.m + i) = make_float4(m[0], m[1], m[2], m[3]);
            *reinterpret_cast<float4*>(a.v + i) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            for (int k = 0; k < 4 && i + k < a.n; k++) { a.p[i + k] = p[k]; a.m[i + k] = m[k]; a.v[i + k] = v[k]; }
        }
    }
}

}  // namespace ols

using namespace ols;

static int adam_launch(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n,
                       const ols_adam_group* groups, int32_t n_groups, double beta1, double beta2, double eps,
                       int64_t step, int64_t* d_step, void* stream) {
    if (!d_param || !d_grad || !d_exp_avg || !d_exp_avg_sq || n < 0 || !groups || n_groups < 1 || n_groups > OLS_ADAM_MAX_GROUPS ||
        (step < 1 && !d_step)) {
        ols_set_error("bad Adam arguments");
        return OLS_ERR_INVALID;
    }
    if ((((uintptr_t)d_param | (uintptr_t)d_grad | (uintptr_t)d_exp_avg | (uintptr_t)d_exp_avg_sq) & 15) != 0) {
        ols_set_error("Adam buffers must be 16-byte aligned");
        return OLS_ERR_INVALID;
    }
    if (n == 0) return OLS_OK;
    AdamArgs a;
    a.p = d_param; a.g = d_grad; a.m = d_exp_avg; a.v = d_exp_avg_sq; a.n = n; a.n_groups = n_groups;
    const double bc1 = d_step ? 1.0 : 1.0 - pow(beta1, (double)step), bc2 = d_step ? 1.0 : 1.0 - pow(beta2, (double)step);
    a.d_step = (const long long*)d_step; a.beta1 = beta1; a.beta2 = beta2;
    long long expect = 0;
    for (int k = 0; k < n_groups; k++) {
        if (groups[k].offset != expect || groups[k].count < 0) { ols_set_error("Adam groups must tile [0, n) in order"); return OLS_ERR_INVALID; }
        a.begin[k] = groups[k].offset;
        a.step_size[k] = (float)((double)groups[k].lr / bc1);
        a.step_tail[k] = (float)((double)groups[k].lr_tail / bc1);
        a.activation[k] = groups[k].activation; a.period[k] = groups[k].period; a.head[k] = groups[k].head;
        if (groups[k].activation < OLS_ACT_NONE || groups[k].activation > OLS_ACT_NORMALIZE4 || groups[k].period < 0 ||
            (groups[k].activation == OLS_ACT_NORMALIZE4 && (groups[k].count % 4 != 0 || groups[k].offset % 4 != 0))) {
            ols_set_error("bad Adam group %d (activation / period)", k);
            return OLS_ERR_INVALID;
        }
        expect += groups[k].count;
    }
    if (expect != n) { ols_set_error("Adam groups cover %lld of %lld elements", expect, (long long)n); return OLS_ERR_INVALID; }
    for (int k = n_groups; k <= OLS_ADAM_MAX_GROUPS; k++) a.begin[k] = n;
    for (int k = n_groups; k < OLS_ADAM_MAX_GROUPS; k++) { a.step_size[k] = a.step_tail[k] = 0.0f; a.activation[k] = a.period[k] = a.head[k] = 0; }
    a.b1 = (float)beta1; a.b2 = (float)beta2; a.one_m_b1 = (float)(1.0 - beta1); a.one_m_b2 = (float)(1.0 - beta2); a.eps = (float)eps; a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    const long long blocks = (n / 4 + 255) / 256 + 1;
    k_adam<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, (cudaStream_t)stream>>>(a);
    if (d_step) k_adam_inc<<<1, 1, 0, (cudaStream_t)stream>>>((long long*)d_step);
    OLS_CUDA_TRY(cudaGetLastError());
    return OLS_OK;
}

extern "C" int ols_adam_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n,
                             const ols_adam_group* groups, int32_t n_groups, double beta1, double beta2, double eps,
                             int64_t step, void* stream) {
    return adam_launch(d_param, d_grad, d_exp_avg, d_exp_avg_sq, n, groups, n_groups, beta1, beta2, eps, step, nullptr, stream);
}

extern "C" int ols_adam_step_dev(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n,
                                 const ols_adam_group* groups, int32_t n_groups, double beta1, double beta2, double eps,
                                 int64_t* d_step, void* stream) {
    if (!d_step) { ols_set_error("null step counter"); return OLS_ERR_INVALID; }
    return adam_launch(d_param, d_grad, d_exp_avg, d_exp_avg_sq, n, groups, n_groups, beta1, beta2, eps, 0, d_step, stream);
}

namespace ols {
// get_opacity / get_scaling / get_rotation (gaussian_model.py:93-130) in one pass over the Gaussians
__global__ void __launch_bounds__(256) k_activate(int P, int scale_cols, const float* __restrict__ o_raw, const float* __restrict__ s_raw,
                                                  const float* __restrict__ q_raw, float* __restrict__ o, float* __restrict__ s,
                                                  float* __restrict__ q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    o[i] = 1.0f / (1.0f + expf(-o_raw[i]));
    for (int c = 0; c < scale_cols; c++) s[(size_t)scale_cols * i + c] = expf(s_raw[(size_t)scale_cols * i + c]);
    const float4 r = reinterpret_cast<const float4*>(q_raw)[i];
    const float inv = 1.0f / fmaxf(sqrtf(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w), 1e-12f);  // F.normalize
    reinterpret_cast<float4*>(q)[i] = make_float4(r.x * inv, r.y * inv, r.z * inv, r.w * inv);
}
}  // namespace ols

extern "C" int ols_activate_params(int32_t P, int32_t scale_cols, const float* d_opacity_raw, const float* d_scaling_raw,
                                   const float* d_rotation_raw, float* d_opacity, float* d_scaling, float* d_rotation, void* stream) {
    if (P < 0 || (scale_cols != 1 && scale_cols != 3) || (P > 0 && (!d_opacity_raw || !d_scaling_raw || !d_rotation_raw || !d_opacity ||
                                                                      !d_scaling || !d_rotation))) {
        ols_set_error("bad activation arguments");
        return OLS_ERR_INVALID;
    }
    if ((((uintptr_t)d_rotation_raw | (uintptr_t)d_rotation) & 15) != 0) { ols_set_error("rotation buffers must be 16-byte aligned"); return OLS_ERR_INVALID; }
    if (P == 0) return OLS_OK;
    k_activate<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, scale_cols, d_opacity_raw, d_scaling_raw, d_rotation_raw, d_opacity,
                                                               d_scaling, d_rotation);
    OLS_CUDA_TRY(cudaGetLastError());
    return OLS_OK;
}

// ---------------------------------------------------------------------------------------------------
// Pose step of the tracking loop, fused (utils/slam_frontend.py:216-262 + utils/pose_utils.py:update_pose):
//   Adam over (cam_rot_delta, cam_trans_delta, exposure_a, exposure_b) with their own learning rates, then
//   T_w2c <- SE3_exp([cam_trans_delta, cam_rot_delta]) * T_w2c, the deltas return to zero, and the three camera tensors
//   render() reads (world_view_transform, full_proj_transform, camera_center) are rebuilt in place.
// The reference does this with ~40 tiny torch kernels and two host synchronisations per tracking iteration (the
// `angle < 1e-5` tests of SO3_exp / V and the `converged` test); here it is one single-thread kernel and the
// convergence flag stays on the device until the caller wants it.
// ---------------------------------------------------------------------------------------------------
namespace ols {
struct PoseStepArgs {
    const float* grad_tau;      // [6] dL/d(rho | theta) = (cam_trans_delta | cam_rot_delta) gradients (ols_bwd_args.d_dL_dtau_sum)
    const float* grad_exposure; // [2] dL/d(exposure_a, exposure_b) or NULL
    float* exposure;            // [2] exposure_a, exposure_b (updated in place) or NULL
    float* m;                   // [8] Adam first moments: rot 3 | trans 3 | exposure 2
    float* v;                   // [8]
    long long* step;            // Adam step count (incremented)
    float* R;                   // [9] row-major world->camera rotation (updated)
    float* T;                   // [3] world->camera translation (updated)
    const float* proj;          // [16] projection_matrix as render() receives it (transposed)
    float *viewmatrix, *projmatrix, *campos;  // [16], [16], [3] outputs
    int* converged;             // OR-accumulated: |tau| < threshold
    float lr_rot, lr_trans, lr_exposure, beta1, beta2, eps, threshold;
    int zero_grads;
};

__global__ void k_pose_step(const PoseStepArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double t = (double)(*a.step + 1);
    const float inv_bc1 = (float)(1.0 / (1.0 - pow((double)a.beta1, t)));
    const float inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)a.beta2, t)));
    float delta[8];
    // parameter order of the reference's optimiser: cam_rot_delta (theta), cam_trans_delta (rho), exposure_a, exposure_b
    for (int k = 0; k < 8; k++) {
        float g;
        if (k < 3) g = a.grad_tau[3 + k];
        else if (k < 6) g = a.grad_tau[k - 3];
        else g = a.grad_exposure ? a.grad_exposure[k - 6] : 0.0f;
        const float lr = k < 3 ? a.lr_rot : (k < 6 ? a.lr_trans : a.lr_exposure);
        float m = a.m[k], v = a.v[k];
        m = m + (g - m) * (1.0f - a.beta1);
        v = v * a.beta2 + (1.0f - a.beta2) * g * g;
        a.m[k] = m; a.v[k] = v;
        delta[k] = -lr * inv_bc1 * (m / (sqrtf(v) * inv_sqrt_bc2 + a.eps));   // the deltas start every iteration at zero
    }
    if (a.exposure) { a.exposure[0] += delta[6]; a.exposure[1] += delta[7]; }
    *a.step += 1;
    if (a.zero_grads) {   // optimizer.zero_grad() of the next iteration (the gradient buffers are accumulated into by autograd)
        for (int k = 0; k < 6; k++) const_cast<float*>(a.grad_tau)[k] = 0.0f;
        if (a.grad_exposure) { const_cast<float*>(a.grad_exposure)[0] = 0.0f; const_cast<float*>(a.grad_exposure)[1] = 0.0f; }
    }
    const float th[3] = {delta[0], delta[1], delta[2]}, rho[3] = {delta[3], delta[4], delta[5]};
    // SO3_exp / V (pose_utils.py:24-57)
    const float W[3][3] = {{0, -th[2], th[1]}, {th[2], 0, -th[0]}, {-th[1], th[0], 0}};
    float W2[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) W2[i][j] = W[i][0] * W[0][j] + W[i][1] * W[1][j] + W[i][2] * W[2][j];
    const float angle = sqrtf(th[0] * th[0] + th[1] * th[1] + th[2] * th[2]);
    float ca, cb, va, vb;   // R = I + ca W + cb W2;  V = I + va W + vb W2
    if (angle < 1e-5f) { ca = 1.0f; cb = 0.5f; va = 0.5f; vb = 1.0f / 6.0f; }
    else {
        ca = sinf(angle) / angle; cb = (1.0f - cosf(angle)) / (angle * angle);
        va = cb; vb = (angle - sinf(angle)) / (angle * angle * angle);
    }
    float dR[3][3], Vm[3][3], dt[3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const float I = i == j ? 1.0f : 0.0f;
            dR[i][j] = I + ca * W[i][j] + cb * W2[i][j];
            Vm[i][j] = I + va * W[i][j] + vb * W2[i][j];
        }
    for (int i = 0; i < 3; i++) dt[i] = Vm[i][0] * rho[0] + Vm[i][1] * rho[1] + Vm[i][2] * rho[2];
    // new_w2c = SE3_exp(tau) @ T_w2c
    float R[3][3], T[3], nR[3][3], nT[3];
    for (int i = 0; i < 3; i++) { T[i] = a.T[i]; for (int j = 0; j < 3; j++) R[i][j] = a.R[3 * i + j]; }
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) nR[i][j] = dR[i][0] * R[0][j] + dR[i][1] * R[1][j] + dR[i][2] * R[2][j];
        nT[i] = dR[i][0] * T[0] + dR[i][1] * T[1] + dR[i][2] * T[2] + dt[i];
    }
    for (int i = 0; i < 3; i++) { a.T[i] = nT[i]; for (int j = 0; j < 3; j++) a.R[3 * i + j] = nR[i][j]; }
    const float nrm = sqrtf(rho[0] * rho[0] + rho[1] * rho[1] + rho[2] * rho[2] + angle * angle);
    if (a.converged && nrm < a.threshold) *a.converged = 1;
    // world_view_transform = getWorld2View2(R, T).T ; full_proj_transform = world_view_transform @ projection_matrix ;
    // camera_center = world_view_transform.inverse()[3, :3] = -R^T T      (utils/camera_utils.py:103-117)
    float Vt[4][4];
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) Vt[j][i] = nR[i][j]; Vt[3][i] = nT[i]; Vt[i][3] = 0.0f; }
    Vt[3][3] = 1.0f;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            a.viewmatrix[4 * i + j] = Vt[i][j];
            float s = 0.0f;
            for (int k = 0; k < 4; k++) s += Vt[i][k] * a.proj[4 * k + j];
            a.projmatrix[4 * i + j] = s;
        }
    for (int j = 0; j < 3; j++) a.campos[j] = -(nR[0][j] * nT[0] + nR[1][j] * nT[1] + nR[2][j] * nT[2]);
}
}  // namespace ols

extern "C" int ols_pose_adam_step(const ols_pose_step* p, void* stream) {
    if (!p || !p->d_grad_tau || !p->d_exp_avg || !p->d_exp_avg_sq || !p->d_step || !p->d_R || !p->d_T || !p->d_projection ||
        !p->d_viewmatrix || !p->d_projmatrix || !p->d_campos) {
        ols_set_error("bad pose-step arguments");
        return OLS_ERR_INVALID;
    }
    PoseStepArgs a;
    a.grad_tau = p->d_grad_tau; a.grad_exposure = p->d_grad_exposure; a.exposure = p->d_exposure; a.m = p->d_exp_avg; a.v = p->d_exp_avg_sq;
    a.step = (long long*)p->d_step; a.R = p->d_R; a.T = p->d_T; a.proj = p->d_projection; a.viewmatrix = p->d_viewmatrix;
    a.projmatrix = p->d_projmatrix; a.campos = p->d_campos; a.converged = p->d_converged;
    a.lr_rot = p->lr_rot; a.lr_trans = p->lr_trans; a.lr_exposure = p->lr_exposure; a.beta1 = p->beta1; a.beta2 = p->beta2;
    a.eps = p->eps; a.threshold = p->converged_threshold; a.zero_grads = p->zero_grads;
    k_pose_step<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    OLS_CUDA_TRY(cudaGetLastError());
    return OLS_OK;
}
