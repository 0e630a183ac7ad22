// ols_forward.cu -- forward pass of the language-feature Gaussian rasterizer for sm_100a.
//
// Pipeline (all on the caller's stream, no host synchronisation, fixed launch geometry -> graph capturable).  Every
// kernel covers the V views of a batch (grid.y = view; V = 1 is the reference's call):
//   k_preprocess        one thread per Gaussian, ALL views of its group: the parameters are read and the 3D covariance is
//                       computed once, then per view cull / project / cov2D / conic / radius / tile rect, the 16-float
//                       colour record, and per-CTA shared-memory tile histograms      (reference: forward.cu:262-371)
//   k_preprocess_dis    the same for the two footprints of the disentangled variant   (D/forward.cu:262-430)
//   k_tile_offsets      column scan of the per-CTA histograms -> per-CTA write cursors and per-tile counts
//   k_tile_scan         one CTA: exclusive scan of per-tile counts -> ranges, R        (replaces the InclusiveSum over
//                       Gaussians + D2H read of rasterizer_impl.cu:451-455 and identifyTileRanges :116-138)
//   k_scatter           writes the Gaussian id (4 bytes) of every (Gaussian, tile) instance into the tile's bucket; a warp
//                       spreads its instances evenly over its lanes                (reference: duplicateWithKeys :70-111)
//   k_sort_tiles_bucket one CTA per tile: gathers the depths, distribution into depth buckets + exact in-bucket ranking
//   k_sort_tiles_radix  tiles the bucket path declines (clumped / tied depths): LSD radix sort in shared memory
//   k_sort_tiles        tiles longer than 4096 entries: bitonic sort, wide strides in global memory
//                       (together they replace the global 44-bit cub::DeviceRadixSort of :478-483; the result order
//                       is identical: (tile, depth bits, id) ascending)
//   k_blend2            one CTA of 4 warps per tile, a warp per 8x8 pixel block, two pixels per lane: front-to-back alpha
//                       blend of RGB + depth + F language channels with per-warp culling  (reference: forward.cu:377-513)
//
// Bit-exactness: the per-Gaussian arithmetic mirrors the compiled reference operation by operation
// (oracle/REF_ARITHMETIC.md), written with explicit-rounding intrinsics so that radii, tile rects,
// depths and every alpha/T threshold decision are identical to the reference build.
#include "ols_common.cuh"

#include <cooperative_groups.h>
#include <math_constants.h>
#include <cstdio>
#include <cstdlib>

namespace ols {

__constant__ float SH_C0 = 0.28209479177387814f;
__constant__ float SH_C1 = 0.4886025119029199f;
__constant__ float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
__constant__ float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

struct PreView {  // what differs between the views of a batch in the preprocess
    const float *viewmatrix, *projmatrix, *campos;
    float tanfovx, tanfovy, focal_x, focal_y;
    char* ws;
    int32_t* radii;
};
struct PreBatch { PreView v[OLS_MAX_VIEWS]; };

struct PreArgs {
    int P, F, sh_degree, M, W, H, tile, gx, gy, rec, V, VG, chunk;
    unsigned flags;
    float scale_modifier;
    const float *means3D, *shs, *colors_precomp, *opacities, *scales, *rotations, *cov3D_precomp;
    size_t o_records, o_depths, o_cov3D, o_clamped, o_tiles_touched, o_rect, o_cta_hist, o_info;  // workspace offsets
};

// m[i]*x + m[4+i]*y + m[8+i]*z + m[12+i] in the compiled reference's order
__device__ __forceinline__ float xform_row(const float* m, int i, float x, float y, float z) {
    return fadd(ffma(z, m[8 + i], ffma(x, m[i], fmul(y, m[4 + i]))), m[12 + i]);
}
__device__ __forceinline__ float dot3c(float a0, float a1, float a2, float b0, float b1, float b2) {
    return ffma(a2, b2, ffma(a0, b0, fmul(a1, b1)));
}

// forward.cu:121-155 (computeCov3D)
__device__ __forceinline__ void cov3d_from_scale_rot(const float sx0, const float sy0, const float sz0, float mod,
                                                     const float4 q, float* out) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    const float sx = fmul(sx0, mod), sy = fmul(sy0, mod), sz = fmul(sz0, mod);
    const float yy = fmul(y, y), zz = fmul(z, z);
    const float xx_yy = ffma(x, x, yy);
    const float R22 = fsub(1.0f, fadd(xx_yy, xx_yy));
    const float xz = fmul(x, z);
    const float xz_m_ry = ffma(-r, y, xz), xz_p_ry = ffma(r, y, xz);
    const float rx = fmul(r, x);
    const float yz_p_rx = ffma(y, z, rx), yz_m_rx = ffma(y, z, -rx);
    const float rz = fmul(r, z);
    const float xy_m_rz = ffma(x, y, -rz), xy_p_rz = ffma(x, y, rz);
    const float yy_zz = fadd(yy, zz);
    const float R00 = fsub(1.0f, fadd(yy_zz, yy_zz));
    const float xx_zz = ffma(x, x, zz);
    const float R11 = fsub(1.0f, fadd(xx_zz, xx_zz));
    const float A2 = fadd(xz_m_ry, xz_m_ry), B2 = fadd(yz_p_rx, yz_p_rx), C2 = fadd(xz_p_ry, xz_p_ry);
    const float D2 = fadd(xy_m_rz, xy_m_rz), E2 = fadd(yz_m_rx, yz_m_rx), G2 = fadd(xy_p_rz, xy_p_rz);
    // M = S * R as GLM evaluates it: the zero entries of S stay in the products
    const float zB = fmul(0.0f, B2);
    const float m22 = ffma(sz, R22, ffma(0.0f, A2, zB));
    const float m02 = ffma(0.0f, R22, ffma(sx, A2, zB));
    const float m12 = ffma(0.0f, R22, ffma(0.0f, A2, fmul(sy, B2)));
    const float z00 = fmul(0.0f, R00);
    const float m20 = ffma(sz, C2, ffma(0.0f, D2, z00));
    const float m00 = ffma(0.0f, C2, ffma(0.0f, D2, fmul(sx, R00)));
    const float m10 = ffma(0.0f, C2, ffma(sy, D2, z00));
    const float z11 = fmul(0.0f, R11);
    const float m21 = ffma(sz, E2, ffma(0.0f, G2, z11));
    const float m01 = ffma(0.0f, E2, ffma(sx, G2, z11));
    const float m11 = ffma(0.0f, E2, ffma(0.0f, G2, fmul(sy, R11)));
    out[0] = dot3c(m00, m10, m20, m00, m10, m20);
    out[1] = dot3c(m00, m10, m20, m01, m11, m21);
    out[2] = dot3c(m00, m10, m20, m02, m12, m22);
    out[3] = dot3c(m01, m11, m21, m01, m11, m21);
    out[4] = dot3c(m01, m11, m21, m02, m12, m22);
    out[5] = dot3c(m02, m12, m22, m02, m12, m22);
}

// auxiliary.h:41-44 -- evaluated in double by the reference
__device__ __forceinline__ float ndc2pix(float v, int S) {
    return __double2float_rn(__dmul_rn(__fma_rn(__dadd_rn((double)v, 1.0), (double)S, -1.0), 0.5));
}

// auxiliary.h:46-56
__device__ __forceinline__ void get_rect(float px, float py, int r, int tile, int gx, int gy, int* mn, int* mx) {
    const float rf = (float)r, tf = (float)tile;
    mn[0] = min(gx, max(0, __float2int_rz(fdiv(fsub(px, rf), tf))));
    mn[1] = min(gy, max(0, __float2int_rz(fdiv(fsub(py, rf), tf))));
    mx[0] = min(gx, max(0, __float2int_rz(fdiv(fadd(fadd(fadd(px, rf), tf), -1.0f), tf))));
    mx[1] = min(gy, max(0, __float2int_rz(fdiv(fadd(fadd(fadd(py, rf), tf), -1.0f), tf))));
}

// forward.cu:23-74 (computeColorFromSH).  Degree 0 (the SLAM configuration) is bit-pinned; higher
// degrees follow the reference's expression order under nvcc's default contraction.
__device__ void sh_to_rgb(int deg, int M, const float* sh, float dx, float dy, float dz, float* res) {
#pragma unroll
    for (int c = 0; c < 3; c++) res[c] = fmul(SH_C0, sh[c]);
    if (deg > 0) {
        const float len = sqrtf(ffma(dz, dz, ffma(dx, dx, fmul(dy, dy))));
        const float x = dx / len, y = dy / len, z = dz / len;
#pragma unroll
        for (int c = 0; c < 3; c++) res[c] = res[c] - SH_C1 * y * sh[3 + c] + SH_C1 * z * sh[6 + c] - SH_C1 * x * sh[9 + c];
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
#pragma unroll
            for (int c = 0; c < 3; c++)
                res[c] = res[c] + SH_C2[0] * xy * sh[12 + c] + SH_C2[1] * yz * sh[15 + c] +
                         SH_C2[2] * (2.0f * zz - xx - yy) * sh[18 + c] + SH_C2[3] * xz * sh[21 + c] +
                         SH_C2[4] * (xx - yy) * sh[24 + c];
            if (deg > 2) {
#pragma unroll
                for (int c = 0; c < 3; c++)
                    res[c] = res[c] + SH_C3[0] * y * (3.0f * xx - yy) * sh[27 + c] + SH_C3[1] * xy * z * sh[30 + c] +
                             SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[33 + c] +
                             SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + c] +
                             SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[39 + c] + SH_C3[5] * z * (xx - yy) * sh[42 + c] +
                             SH_C3[6] * x * (xx - 3.0f * yy) * sh[45 + c];
            }
        }
    }
}

constexpr int PRE_THREADS = 256;
constexpr int PRE_CAM = 36;  // floats of per-view camera data staged in shared memory: V[16] Pm[16] campos[3] pad

// ---- pieces of the per-Gaussian preprocess shared by the joint (P/) and the disentangled (D/) kernels ----

// computeCov2D (forward.cu:77-116): cov2D (a, b, c) with the 0.3 dilation already added to a and c
__device__ __forceinline__ void cov2d_from_cov3d(const float* V, float px3, float py3, float pz3, float vz,
                                                 float tanfovx, float tanfovy, float focal_x, float focal_y,
                                                 const float* c3, float& ca, float& cb, float& cc) {
    const float tz = vz;
    const float txr = xform_row(V, 0, px3, py3, pz3);
    const float tyr = xform_row(V, 1, px3, py3, pz3);
    const float limx = fmul(tanfovx, 1.3f), limy = fmul(tanfovy, 1.3f);
    const float cx = fminf(fmaxf(fdiv(txr, tz), -limx), limx);
    const float cy = fminf(fmaxf(fdiv(tyr, tz), -limy), limy);
    const float tz2 = fmul(tz, tz);
    const float J00 = fdiv(focal_x, tz);
    const float J02 = fdiv(fmul(fmul(tz, -cx), focal_x), tz2);
    const float J11 = fdiv(focal_y, tz);
    const float J12 = fdiv(fmul(fmul(tz, -cy), focal_y), tz2);
    float ta[3], tb[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        ta[k] = ffma(V[4 * k + 2], J02, ffma(V[4 * k], J00, fmul(0.0f, V[4 * k + 1])));
        tb[k] = ffma(V[4 * k + 2], J12, ffma(0.0f, V[4 * k], fmul(V[4 * k + 1], J11)));
    }
    const float ux0 = dot3c(ta[0], ta[1], ta[2], c3[0], c3[1], c3[2]);
    const float ux1 = dot3c(ta[0], ta[1], ta[2], c3[1], c3[3], c3[4]);
    const float ux2 = dot3c(ta[0], ta[1], ta[2], c3[2], c3[4], c3[5]);
    const float uy0 = dot3c(tb[0], tb[1], tb[2], c3[0], c3[1], c3[2]);
    const float uy1 = dot3c(tb[0], tb[1], tb[2], c3[1], c3[3], c3[4]);
    const float uy2 = dot3c(tb[0], tb[1], tb[2], c3[2], c3[4], c3[5]);
    ca = fadd(dot3c(ta[0], ta[1], ta[2], ux0, ux1, ux2), 0.3f);
    cb = dot3c(ta[0], ta[1], ta[2], uy0, uy1, uy2);
    cc = fadd(dot3c(tb[0], tb[1], tb[2], uy0, uy1, uy2), 0.3f);
}

// forward.cu:344-347: radius = ceil(3 * sqrt(larger eigenvalue))
__device__ __forceinline__ float splat_radius(float ca, float cc, float det) {
    const float mid = fmul(fadd(ca, cc), 0.5f);
    const float sq = __fsqrt_rn(fmaxf(ffma(mid, mid, -det), 0.1f));
    const float lam = fmaxf(fadd(mid, sq), fsub(mid, sq));
    return ceilf(fmul(__fsqrt_rn(lam), 3.0f));
}

// SH -> rgb with the clamp flags (forward.cu:23-74,  +0.5 and max(.,0) at :66-73)
__device__ __forceinline__ uint32_t rgb_from_sh(int deg, int M, const float* sh, float dx, float dy, float dz, float* rgb) {
    float res[3];
    sh_to_rgb(deg, M, sh, dx, dy, dz, res);
    uint32_t cl = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        cl |= (uint32_t)(!(res[c] >= -0.5f)) << (8 * c);  // compiled form of "result < 0 after +0.5"
        rgb[c] = fmaxf(fadd(res[c], 0.5f), 0.0f);
    }
    return cl;
}

// Header (x y A B | C op pth depth) and the trailing half-extents (ex, ey) of a blend record.
__device__ __forceinline__ void write_record_frame(float* rf, int rec, float pix_x, float pix_y, float conA, float conB,
                                                   float conC, float op, float depth) {
    // conservative cut: power < pth  =>  op*exp(power) < 1/255 with a wide safety margin
    const float pth = (op == op) ? (op > 0.0f ? logf(1.0f / (255.0f * op)) - 0.01f : CUDART_INF_F) : -CUDART_INF_F;
    float4* r4 = reinterpret_cast<float4*>(rf);
    r4[0] = make_float4(pix_x, pix_y, conA, conB);
    r4[1] = make_float4(conC, op, pth, depth);
    // half-extents of {power >= pth_c} for the conic as stored: |dx| <= sqrt(-2 pth_c C / (AC - B^2)), same for y.
    // pth_c widens pth by more than the rounding error of the float evaluation of `power` when the form is
    // reasonably conditioned (trace^2 < 1000 det, so the fp32 determinant below is good to ~1e-4 relative -- the 1.001
    // factor covers it); otherwise (or for NaNs) the extents are infinite = never rejected.
    const float A = conA, B = conB, Cc = conC;
    const float detc = fmaf(A, Cc, -B * B), tr = A + Cc;
    float ex = CUDART_INF_F, ey = CUDART_INF_F;
    if (pth > 0.0f) {
        ex = ey = -CUDART_INF_F;  // opacity below 1/255: no pixel can pass
    } else if (pth > -CUDART_INF_F && A > 0.0f && Cc > 0.0f && detc > 0.0f && tr * tr < 1000.0f * detc) {
        const float k = -2.0f * (1.002f * pth - 0.02f) / detc;
        ex = sqrtf(k * Cc) * 1.001f + 0.01f;
        ey = sqrtf(k * A) * 1.001f + 0.01f;
    }
    rf[rec - 2] = ex;
    rf[rec - 1] = ey;
}

// One view of one Gaussian: cull / project / cov2D / conic / radius / tile rect / colour, the 16-float colour record,
// per-tile instance counts (forward.cu:262-371 after computeCov3D).  Returns whether the Gaussian is visible.
__device__ __forceinline__ bool preprocess_view(const PreArgs& a, const PreView& pv, int i, float px3, float py3, float pz3,
                                                const float* c3, float op, const float* rgb_pre, const float* cam,
                                                uint32_t* s_hist) {
    char* ws = pv.ws;
    uint32_t* tiles_touched = (uint32_t*)(ws + a.o_tiles_touched);
    pv.radii[i] = 0;
    tiles_touched[i] = 0;
    const float* V = cam;
    const float* Pm = cam + 16;
    // in_frustum (auxiliary.h:139-164)
    const float vz = xform_row(V, 2, px3, py3, pz3);
    if (!(vz > 0.2f)) {
        if (a.flags & OLS_FLAG_PREFILTERED) {
            printf("Point is filtered although prefiltered is set. This shouldn't happen!");
            __trap();
        }
        return false;
    }
    const float hx = xform_row(Pm, 0, px3, py3, pz3);
    const float hy = xform_row(Pm, 1, px3, py3, pz3);
    const float hw = xform_row(Pm, 3, px3, py3, pz3);
    const float pw = __frcp_rn(fadd(hw, 0.0000001f));
    const float projx = fmul(hx, pw), projy = fmul(hy, pw);
    float ca, cb, cc;
    cov2d_from_cov3d(V, px3, py3, pz3, vz, pv.tanfovx, pv.tanfovy, pv.focal_x, pv.focal_y, c3, ca, cb, cc);
    const float det = ffma(ca, cc, -fmul(cb, cb));
    if (det == 0.0f) return false;
    const float det_inv = __frcp_rn(det);
    const float my_radius = splat_radius(ca, cc, det);
    const float pix_x = ndc2pix(projx, a.W), pix_y = ndc2pix(projy, a.H);
    int mn[2], mx[2];
    const int ri = __float2int_rz(my_radius);
    get_rect(pix_x, pix_y, ri, a.tile, a.gx, a.gy, mn, mx);
    const uint32_t tiles = (uint32_t)(mx[0] - mn[0]) * (uint32_t)(mx[1] - mn[1]);
    if (tiles == 0) return false;

    float rgb[3];
    if (rgb_pre) {
#pragma unroll
        for (int c = 0; c < 3; c++) rgb[c] = rgb_pre[c];
    } else {
        ((uint32_t*)(ws + a.o_clamped))[i] = rgb_from_sh(a.sh_degree, a.M, a.shs + (size_t)i * a.M * 3, px3 - cam[32],
                                                          py3 - cam[33], pz3 - cam[34], rgb);
    }
    float* rf = (float*)(ws + a.o_records) + (size_t)i * a.rec;
    float4* r4 = reinterpret_cast<float4*>(rf);
    r4[2] = make_float4(rgb[0], rgb[1], rgb[2], 0.0f);
    for (int c = REC_CH + 4; c < a.rec - 2; c++) rf[c] = 0.0f;  // the language channels are not part of the global record
    write_record_frame(rf, a.rec, pix_x, pix_y, fmul(cc, det_inv), fmul(cb, -det_inv), fmul(ca, det_inv), op, vz);
    ((float*)(ws + a.o_depths))[i] = vz;
    pv.radii[i] = ri;
    ((uint2*)(ws + a.o_rect))[i] = make_uint2((uint32_t)mn[0] | ((uint32_t)mn[1] << 16), (uint32_t)mx[0] | ((uint32_t)mx[1] << 16));
    tiles_touched[i] = tiles;
    // one flat loop over the rect (a warp runs max(tiles) rounds; the nested form ran max(height) * max(width))
    int x = mn[0], t_idx = mn[1] * a.gx + mn[0];
    for (uint32_t t = 0; t < tiles; t++) {
        atomicAdd(&s_hist[t_idx], 1u);
        t_idx++;
        if (++x == mx[0]) { x = mn[0]; t_idx += a.gx - (mx[0] - mn[0]); }
    }
    return true;
}

// Each CTA owns a contiguous chunk of Gaussians (one thread per Gaussian, 256 at a time) and a group of up to VG
// views (blockIdx.y).  A Gaussian's parameters are read and its 3D covariance is computed ONCE for all the views of
// the group -- the reference runs the whole preprocess once per rendered view (slam_backend.py:510-662: 8-12 views per
// mapping iteration over the same Gaussians).  Inputs with 3-float rows are staged through shared memory so the
// global loads are contiguous; the instances a Gaussian contributes to every tile of every view are counted in per-CTA
// shared-memory histograms written out once per CTA (cta_hist[cta][tile]) -- no global atomics on hot tile counters.
__global__ void __launch_bounds__(PRE_THREADS) k_preprocess(const PreArgs a, const __grid_constant__ PreBatch vb) {
    extern __shared__ uint32_t s_dyn[];  // [VG][n_tiles] histograms, then [VG][PRE_CAM] camera floats
    __shared__ float s_mean[PRE_THREADS * 3];
    __shared__ float s_scale[PRE_THREADS * 3];
    __shared__ int s_nvis[OLS_MAX_VIEWS];
    const int tid = threadIdx.x;
    const int n_tiles = a.gx * a.gy;
    const int v0 = blockIdx.y * a.VG, nv = min(a.VG, a.V - v0);
    uint32_t* s_hist = s_dyn;
    float* s_cam = reinterpret_cast<float*>(s_dyn + (size_t)a.VG * n_tiles);
    for (int t = tid; t < nv * n_tiles; t += PRE_THREADS) s_hist[t] = 0;
    for (int e = tid; e < nv * PRE_CAM; e += PRE_THREADS) {
        const int v = e / PRE_CAM, k = e - v * PRE_CAM;
        const PreView& pv = vb.v[v0 + v];
        s_cam[e] = k < 16 ? pv.viewmatrix[k] : (k < 32 ? pv.projmatrix[k - 16] : (k < 35 ? pv.campos[k - 32] : 0.0f));
    }
    if (tid < OLS_MAX_VIEWS) s_nvis[tid] = 0;
    const int chunk_begin = min(a.P, (int)blockIdx.x * a.chunk), chunk_end = min(a.P, chunk_begin + a.chunk);
    for (int base = chunk_begin; base < chunk_end; base += PRE_THREADS) {
        const int nloc = min(PRE_THREADS, chunk_end - base);
        __syncthreads();  // previous iteration's readers of s_mean/s_scale are done (and the histograms are zeroed)
        for (int e = tid; e < nloc * 3; e += PRE_THREADS) {
            s_mean[e] = a.means3D[(size_t)base * 3 + e];
            if (a.scales) s_scale[e] = a.scales[(size_t)base * 3 + e];
        }
        __syncthreads();
        const bool valid = tid < nloc;
        const int i = base + (valid ? tid : 0);
        const float px3 = s_mean[3 * (valid ? tid : 0)], py3 = s_mean[3 * (valid ? tid : 0) + 1], pz3 = s_mean[3 * (valid ? tid : 0) + 2];
        float c3[6], rgb_pre[3] = {0.0f, 0.0f, 0.0f};
        float op = 0.0f;
        if (valid) {
            op = a.opacities[i];
            if (a.cov3D_precomp) {
#pragma unroll
                for (int k = 0; k < 6; k++) c3[k] = a.cov3D_precomp[(size_t)6 * i + k];
            } else {
                const float4 q = reinterpret_cast<const float4*>(a.rotations)[i];
                cov3d_from_scale_rot(s_scale[3 * tid], s_scale[3 * tid + 1], s_scale[3 * tid + 2], a.scale_modifier, q, c3);
                if (blockIdx.y == 0) {  // view-independent: kept once, in the first view's workspace
                    float2* dst = reinterpret_cast<float2*>((float*)(vb.v[0].ws + a.o_cov3D) + (size_t)6 * i);
                    dst[0] = make_float2(c3[0], c3[1]); dst[1] = make_float2(c3[2], c3[3]); dst[2] = make_float2(c3[4], c3[5]);
                }
            }
            if (a.colors_precomp) {
#pragma unroll
                for (int c = 0; c < 3; c++) rgb_pre[c] = a.colors_precomp[(size_t)3 * i + c];
            }
        }
        for (int v = 0; v < nv; v++) {
            bool vis = false;
            if (valid)
                vis = preprocess_view(a, vb.v[v0 + v], i, px3, py3, pz3, c3, op, a.colors_precomp ? rgb_pre : nullptr,
                                      s_cam + v * PRE_CAM, s_hist + (size_t)v * n_tiles);
            const unsigned m = __ballot_sync(0xffffffffu, vis);
            if ((tid & 31) == 0 && m) atomicAdd(&s_nvis[v], __popc(m));
        }
    }
    __syncthreads();  // orders the histogram updates before the flush
    for (int v = 0; v < nv; v++) {
        uint32_t* out = (uint32_t*)(vb.v[v0 + v].ws + a.o_cta_hist) + (size_t)blockIdx.x * n_tiles;
        const uint32_t* h = s_hist + (size_t)v * n_tiles;
        for (int t = tid; t < n_tiles; t += PRE_THREADS) out[t] = h[t];
        if (tid == 0 && s_nvis[v]) atomicAdd(&((DeviceInfo*)(vb.v[v0 + v].ws + a.o_info))->n_visible, s_nvis[v]);
    }
}

// ---------------------------------------------------------------------------------------------------
// Disentangled variant (D/: submodules/diff-gaussian-rasterization-disentangle-optim).  One Gaussian has
// two screen-space footprints -- (scales, rotations, opacities) for colour + depth and (scales_lang,
// rotations_lang, opacities_lang) for the language features -- sharing the projected mean and the depth
// (D/cuda_rasterizer/forward.cu:262-430).  The kernel fills one blend record, rect and tile histogram
// per footprint; binning, sorting and blending then run once per footprint with the kernels below.
// ---------------------------------------------------------------------------------------------------
struct PreDisArgs {
    int P, F, sh_degree, M, W, H, tile, gx, gy, rec_c, rec_l, chunk;
    unsigned flags;
    float tanfovx, tanfovy, focal_x, focal_y, scale_modifier;
    const float *means3D, *shs, *colors_precomp, *language;
    const float *opacities, *scales, *rotations, *cov3D_precomp;
    const float *opacities_lang, *scales_lang, *rotations_lang, *cov3D_precomp_lang;
    const float *viewmatrix, *projmatrix, *campos;
    float *records_c, *records_l, *depths, *cov3D, *cov3D_lang;
    uint32_t *clamped, *tiles_touched_c, *tiles_touched_l, *cta_hist_c, *cta_hist_l;
    uint2 *rect_c, *rect_l;
    int32_t *radii, *radii_lang;
    DeviceInfo *info_c, *info_l;
};

__global__ void __launch_bounds__(PRE_THREADS) k_preprocess_dis(const PreDisArgs a) {
    extern __shared__ uint32_t s_hist2[];  // [2][n_tiles]
    __shared__ float s_V[16], s_Pm[16];
    const int tid = threadIdx.x;
    const int n_tiles = a.gx * a.gy;
    uint32_t* hist_c = s_hist2;
    uint32_t* hist_l = s_hist2 + n_tiles;
    for (int t = tid; t < 2 * n_tiles; t += PRE_THREADS) s_hist2[t] = 0;
    if (tid < 16) { s_V[tid] = a.viewmatrix[tid]; s_Pm[tid] = a.projmatrix[tid]; }
    __syncthreads();
    const float* V = s_V;
    const float* Pm = s_Pm;
    const int chunk_begin = min(a.P, (int)blockIdx.x * a.chunk), chunk_end = min(a.P, chunk_begin + a.chunk);
    int vis_c = 0, vis_l = 0;
    for (int i = chunk_begin + tid; i < chunk_end; i += PRE_THREADS) {
        a.radii[i] = 0;
        a.radii_lang[i] = 0;
        a.tiles_touched_c[i] = 0;
        a.tiles_touched_l[i] = 0;
        float* rc = a.records_c + (size_t)i * a.rec_c;
        float* rl = a.records_l + (size_t)i * a.rec_l;
        // language channels of the record are written for every Gaussian (only listed ones are read)
        for (int c = 0; c < a.F; c++) rl[REC_CH + c] = a.language[(size_t)i * a.F + c];
        for (int c = REC_CH + a.F; c < a.rec_l - 2; c++) rl[c] = 0.0f;
        const float px3 = a.means3D[3 * (size_t)i], py3 = a.means3D[3 * (size_t)i + 1], pz3 = a.means3D[3 * (size_t)i + 2];
        const float vz = xform_row(V, 2, px3, py3, pz3);
        if (!(vz > 0.2f)) {
            if (a.flags & OLS_FLAG_PREFILTERED) {
                printf("Point is filtered although prefiltered is set. This shouldn't happen!");
                __trap();
            }
            continue;
        }
        const float hx = xform_row(Pm, 0, px3, py3, pz3);
        const float hy = xform_row(Pm, 1, px3, py3, pz3);
        const float hw = xform_row(Pm, 3, px3, py3, pz3);
        const float pw = __frcp_rn(fadd(hw, 0.0000001f));
        const float projx = fmul(hx, pw), projy = fmul(hy, pw);
        float c3[6], c3l[6];
        if (a.cov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; k++) c3[k] = a.cov3D_precomp[(size_t)6 * i + k];
        } else {
            const float4 q = reinterpret_cast<const float4*>(a.rotations)[i];
            cov3d_from_scale_rot(a.scales[3 * (size_t)i], a.scales[3 * (size_t)i + 1], a.scales[3 * (size_t)i + 2],
                                 a.scale_modifier, q, c3);
#pragma unroll
            for (int k = 0; k < 6; k++) a.cov3D[(size_t)6 * i + k] = c3[k];
        }
        if (a.cov3D_precomp_lang) {
#pragma unroll
            for (int k = 0; k < 6; k++) c3l[k] = a.cov3D_precomp_lang[(size_t)6 * i + k];
        } else {
            const float4 q = reinterpret_cast<const float4*>(a.rotations_lang)[i];
            cov3d_from_scale_rot(a.scales_lang[3 * (size_t)i], a.scales_lang[3 * (size_t)i + 1],
                                 a.scales_lang[3 * (size_t)i + 2], a.scale_modifier, q, c3l);
#pragma unroll
            for (int k = 0; k < 6; k++) a.cov3D_lang[(size_t)6 * i + k] = c3l[k];
        }
        float ca, cb, cc, la, lb, lc;
        cov2d_from_cov3d(V, px3, py3, pz3, vz, a.tanfovx, a.tanfovy, a.focal_x, a.focal_y, c3, ca, cb, cc);
        cov2d_from_cov3d(V, px3, py3, pz3, vz, a.tanfovx, a.tanfovy, a.focal_x, a.focal_y, c3l, la, lb, lc);
        const float det = ffma(ca, cc, -fmul(cb, cb));
        const float det_l = ffma(la, lc, -fmul(lb, lb));
        if (det == 0.0f && det_l == 0.0f) continue;  // D/forward.cu:357-366: only when BOTH are singular
        const float det_inv = __frcp_rn(det), det_inv_l = __frcp_rn(det_l);
        const float pix_x = ndc2pix(projx, a.W), pix_y = ndc2pix(projy, a.H);
        const int ri = __float2int_rz(splat_radius(ca, cc, det));
        const int ril = __float2int_rz(splat_radius(la, lc, det_l));
        int mn[2], mx[2], mnl[2], mxl[2];
        get_rect(pix_x, pix_y, ri, a.tile, a.gx, a.gy, mn, mx);
        get_rect(pix_x, pix_y, ril, a.tile, a.gx, a.gy, mnl, mxl);
        const uint32_t tiles = (uint32_t)(mx[0] - mn[0]) * (uint32_t)(mx[1] - mn[1]);
        const uint32_t tiles_l = (uint32_t)(mxl[0] - mnl[0]) * (uint32_t)(mxl[1] - mnl[1]);
        if (tiles == 0 && tiles_l == 0) continue;  // D/forward.cu:391-394
        float rgb[3] = {0.0f, 0.0f, 0.0f};
        if (a.colors_precomp) {
#pragma unroll
            for (int c = 0; c < 3; c++) rgb[c] = a.colors_precomp[(size_t)3 * i + c];
        } else if (tiles != 0) {  // D/forward.cu:398-411: zero colour (and no clamp flags) when the colour rect is empty
            a.clamped[i] = rgb_from_sh(a.sh_degree, a.M, a.shs + (size_t)i * a.M * 3, px3 - a.campos[0],
                                       py3 - a.campos[1], pz3 - a.campos[2], rgb);
        }
        rc[REC_CH] = rgb[0];
        rc[REC_CH + 1] = rgb[1];
        rc[REC_CH + 2] = rgb[2];
        for (int c = REC_CH + 3; c < a.rec_c - 2; c++) rc[c] = 0.0f;
        write_record_frame(rc, a.rec_c, pix_x, pix_y, fmul(cc, det_inv), fmul(cb, -det_inv), fmul(ca, det_inv),
                           a.opacities[i], vz);
        write_record_frame(rl, a.rec_l, pix_x, pix_y, fmul(lc, det_inv_l), fmul(lb, -det_inv_l), fmul(la, det_inv_l),
                           a.opacities_lang[i], vz);
        a.depths[i] = vz;
        a.radii[i] = ri;            // both radii are stored even when one rect is empty (D/forward.cu:416-428)
        a.radii_lang[i] = ril;
        a.rect_c[i] = make_uint2((uint32_t)mn[0] | ((uint32_t)mn[1] << 16), (uint32_t)mx[0] | ((uint32_t)mx[1] << 16));
        a.rect_l[i] = make_uint2((uint32_t)mnl[0] | ((uint32_t)mnl[1] << 16), (uint32_t)mxl[0] | ((uint32_t)mxl[1] << 16));
        a.tiles_touched_c[i] = tiles;
        a.tiles_touched_l[i] = tiles_l;
        vis_c += ri > 0;
        vis_l += ril > 0;
        for (int y = mn[1]; y < mx[1]; y++)
            for (int x = mn[0]; x < mx[0]; x++) atomicAdd(&hist_c[y * a.gx + x], 1u);
        for (int y = mnl[1]; y < mxl[1]; y++)
            for (int x = mnl[0]; x < mxl[0]; x++) atomicAdd(&hist_l[y * a.gx + x], 1u);
    }
    __syncthreads();
    uint32_t* out_c = a.cta_hist_c + (size_t)blockIdx.x * n_tiles;
    uint32_t* out_l = a.cta_hist_l + (size_t)blockIdx.x * n_tiles;
    for (int t = tid; t < n_tiles; t += PRE_THREADS) { out_c[t] = hist_c[t]; out_l[t] = hist_l[t]; }
    for (int o = 16; o > 0; o >>= 1) {
        vis_c += __shfl_xor_sync(0xffffffffu, vis_c, o);
        vis_l += __shfl_xor_sync(0xffffffffu, vis_l, o);
    }
    if ((tid & 31) == 0) {
        if (vis_c) atomicAdd(&a.info_c->n_visible, vis_c);
        if (vis_l) atomicAdd(&a.info_l->n_visible, vis_l);
    }
}

// Column scan of cta_hist: afterwards cta_hist[c][t] = instances of tile t emitted by CTAs < c, and
// tile_count[t] = instances of tile t.  A CTA handles 32 tiles; its 8 warps split the CTA axis, each lane
// owns one tile (128-byte coalesced rows), partial sums are combined through shared memory.
constexpr int TO_WARPS = 8;
__global__ void __launch_bounds__(32 * TO_WARPS) k_tile_offsets(const __grid_constant__ PassBatch pb, const WsLayout L) {
    __shared__ uint32_t s_part[TO_WARPS][32];
    char* ws = pb.v[blockIdx.y].ws;
    uint32_t* __restrict__ cta_hist = (uint32_t*)(ws + L.cta_hist);
    uint32_t* __restrict__ tile_count = (uint32_t*)(ws + L.tile_count);
    const int n_tiles = L.n_tiles, n_ctas = L.n_ctas;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + lane;
    const bool ok = t < n_tiles;
    const int seg = (n_ctas + TO_WARPS - 1) / TO_WARPS;
    const int c0 = min(n_ctas, w * seg), c1 = min(n_ctas, c0 + seg);
    uint32_t sum = 0;
    if (ok) {
        int c = c0;
        for (; c + 8 <= c1; c += 8) {
            uint32_t v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = cta_hist[(size_t)(c + k) * n_tiles + t];
#pragma unroll
            for (int k = 0; k < 8; k++) sum += v[k];
        }
        for (; c < c1; c++) sum += cta_hist[(size_t)c * n_tiles + t];
    }
    s_part[w][lane] = sum;
    __syncthreads();
    uint32_t run = 0, total = 0;
#pragma unroll
    for (int k = 0; k < TO_WARPS; k++) {
        const uint32_t v = s_part[k][lane];
        if (k < w) run += v;
        total += v;
    }
    if (!ok) return;
    int c = c0;
    for (; c + 8 <= c1; c += 8) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = cta_hist[(size_t)(c + k) * n_tiles + t];
#pragma unroll
        for (int k = 0; k < 8; k++) { cta_hist[(size_t)(c + k) * n_tiles + t] = run; run += v[k]; }
    }
    for (; c < c1; c++) {
        const uint32_t v = cta_hist[(size_t)c * n_tiles + t];
        cta_hist[(size_t)c * n_tiles + t] = run;
        run += v;
    }
    if (w == 0) tile_count[t] = total;
}

// One CTA: exclusive scan over tiles.  ranges of empty tiles stay (0,0) like the reference's memset
// (rasterizer_impl.cu:485).
constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS) k_tile_scan(const __grid_constant__ PassBatch pb, const WsLayout L,
                                                          unsigned long long R_cap) {
    __shared__ uint32_t s_warp[32];
    char* ws = pb.v[blockIdx.y].ws;
    const uint32_t* __restrict__ tile_count = (const uint32_t*)(ws + L.tile_count);
    uint32_t* __restrict__ tile_cursor = (uint32_t*)(ws + L.tile_cursor);
    uint2* __restrict__ ranges = (uint2*)(ws + L.ranges);
    DeviceInfo* info = (DeviceInfo*)(ws + L.info);
    const int n_tiles = L.n_tiles;
    __shared__ uint32_t s_carry;
    __shared__ uint32_t s_max;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { s_carry = 0; s_max = 0; }
    __syncthreads();
    uint32_t local_max = 0;
    for (int base = 0; base < n_tiles; base += SCAN_THREADS) {
        const int t = base + tid;
        const uint32_t c = t < n_tiles ? tile_count[t] : 0u;
        local_max = max(local_max, c);
        uint32_t v = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        if (lane == 31) s_warp[wid] = v;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += n;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const uint32_t incl = v + (wid ? s_warp[wid - 1] : 0u) + s_carry;
        if (t < n_tiles) {
            ranges[t] = c ? make_uint2(incl - c, incl) : make_uint2(0u, 0u);
            tile_cursor[t] = incl - c;
        }
        __syncthreads();
        if (tid == SCAN_THREADS - 1) s_carry = incl;
        __syncthreads();
    }
    atomicMax(&s_max, local_max);
    __syncthreads();
    if (tid == 0) {
        info->R = s_carry;
        info->overflow = (unsigned long long)s_carry > R_cap ? 1 : 0;
        info->max_tile_len = (int)s_max;
    }
}

// duplicateWithKeys (rasterizer_impl.cu:70-111), bucketed: the tile id is implicit in the bucket.
// CTA c handles the same chunk of Gaussians as in k_preprocess; its write cursors (tile start +
// instances emitted by earlier CTAs) live in shared memory, so slots are handed out by shared-memory
// atomics only.
__global__ void __launch_bounds__(PRE_THREADS) k_scatter(const __grid_constant__ PassBatch pb, const WsLayout L, int P) {
    extern __shared__ uint32_t s_cur[];  // [n_tiles]
    char* ws = pb.v[blockIdx.y].ws;
    const int gx = L.gx, n_tiles = L.n_tiles, chunk = L.chunk;
    const uint32_t* __restrict__ tiles_touched = (const uint32_t*)(ws + L.tiles_touched);
    const uint2* __restrict__ rect = (const uint2*)(ws + L.rect);
    const uint32_t* __restrict__ cta_hist = (const uint32_t*)(ws + L.cta_hist);
    const uint32_t* __restrict__ tile_start = (const uint32_t*)(ws + L.tile_cursor);
    uint32_t* __restrict__ bucket = (uint32_t*)(ws + L.point_list);  // unsorted per-tile lists of Gaussian ids (4 B per instance)
    const DeviceInfo* __restrict__ info = (const DeviceInfo*)(ws + L.info);
    if (info->overflow) return;
    const int tid = threadIdx.x;
    const uint32_t* mine = cta_hist + (size_t)blockIdx.x * n_tiles;
    for (int t = tid; t < n_tiles; t += PRE_THREADS) s_cur[t] = tile_start[t] + mine[t];
    __syncthreads();
    const int chunk_begin = min(P, (int)blockIdx.x * chunk), chunk_end = min(P, chunk_begin + chunk);
    const int lane = tid & 31;
    // A warp takes 32 Gaussians at a time and spreads their (Gaussian, tile) instances evenly over its lanes:
    // lane L emits instances L, L + 32, ... of the warp's flattened list, so a Gaussian covering many tiles does
    // not serialise one thread, and 32 independent slot requests are in flight per step.
    // the three per-Gaussian loads of the next round are issued before the current round is emitted
    auto fetch = [&](int i, uint32_t& cnt, uint2& r) {
        cnt = 0; r = make_uint2(0u, 0u);
        if (i < chunk_end) { cnt = tiles_touched[i]; r = rect[i]; }
    };
    uint32_t n_cnt; uint2 n_r;
    fetch(chunk_begin + (tid & ~31) + lane, n_cnt, n_r);
    for (int i0 = chunk_begin + (tid & ~31); i0 < chunk_end; i0 += PRE_THREADS) {
        const int i = i0 + lane;
        const uint32_t cnt = n_cnt;
        const uint2 r = n_r;
        fetch(i + PRE_THREADS, n_cnt, n_r);
        uint32_t incl = cnt;  // inclusive prefix over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        for (uint32_t k = lane; k < ((total + 31u) & ~31u); k += 32) {
            // owner = first lane whose inclusive prefix exceeds k (binary search over the lanes)
            int owner = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const uint32_t v = __shfl_sync(0xffffffffu, incl, owner + step - 1);
                if (v <= k) owner += step;
            }
            owner = min(owner, 31);
            const uint32_t o_incl = __shfl_sync(0xffffffffu, incl, owner);
            const uint32_t o_cnt = __shfl_sync(0xffffffffu, cnt, owner);
            const uint32_t rx = __shfl_sync(0xffffffffu, r.x, owner), ry = __shfl_sync(0xffffffffu, r.y, owner);
            if (k < total) {
                const uint32_t local = k - (o_incl - o_cnt);
                const int x0 = rx & 0xffff, y0 = rx >> 16, x1 = ry & 0xffff;
                const int w = x1 - x0;
                const int yy = y0 + (int)(local / (uint32_t)w), xx = x0 + (int)(local % (uint32_t)w);
                const uint32_t slot = atomicAdd(&s_cur[yy * gx + xx], 1u);
                bucket[slot] = (uint32_t)(i0 + owner);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Per-tile sort, fast path: one distribution pass + exact ranking inside tiny buckets.
// The depths of a tile's bucket are spread over [dmin, dmax]; a key goes to bucket
// floor((depth - dmin) * NB / (dmax - dmin)), which is monotone in the depth bits (positive floats), so the
// buckets are already in final order and only the few keys sharing a bucket have to be ordered -- by
// counting, for each key, the keys of its bucket that compare lower on the full 64-bit (depth, id) key.
// The result is the same permutation the reference's stable 44-bit radix sort produces.  Tiles whose depth
// distribution is too clumped for this (a bucket with more than BS_LIMIT keys) or that hold non-finite depths
// are left to the radix kernel below; tile_count[tile] = BS_DONE marks the tiles finished here.
// ---------------------------------------------------------------------------------------------------
constexpr int BS_THREADS = 256;
constexpr int BS_CAP = 4096;                     // keys per tile handled in shared memory
constexpr int BS_ITEMS = BS_CAP / BS_THREADS;
constexpr int BS_NB = 2048;                      // buckets
constexpr int BS_LIMIT = 48;                     // largest bucket this path accepts
constexpr uint32_t BS_DONE = 0xffffffffu;

__global__ void __launch_bounds__(BS_THREADS) k_sort_tiles_bucket(const __grid_constant__ PassBatch pb, const WsLayout L,
                                                                 const int write_keys) {
    __shared__ unsigned long long s_key[BS_CAP];  // 32 KB
    char* ws = pb.v[blockIdx.y].ws;
    unsigned long long* __restrict__ keys = (unsigned long long*)(ws + L.keys);
    uint32_t* __restrict__ point_list = (uint32_t*)(ws + L.point_list);
    const float* __restrict__ depths = pb.v[blockIdx.y].depths;
    const uint2* __restrict__ ranges = (const uint2*)(ws + L.ranges);
    uint32_t* __restrict__ tile_count = (uint32_t*)(ws + L.tile_count);
    const DeviceInfo* __restrict__ info = (const DeviceInfo*)(ws + L.info);
    __shared__ uint32_t s_cnt[BS_NB];             // 8 KB: counts -> exclusive starts -> ends
    __shared__ uint32_t s_warp[BS_THREADS / 32];
    __shared__ float s_lohi[2];
    if (info->overflow) return;
    const uint2 rg = ranges[blockIdx.x];
    const int n = (int)(rg.y - rg.x);
    if (n <= 0 || n > BS_CAP) return;
    unsigned long long* g = keys + rg.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    unsigned long long k[BS_ITEMS];
    float lo = CUDART_INF_F, hi = 0.0f;  // depths are non-negative here, so 0 is the identity of max
    bool finite = true;
#pragma unroll
    for (int r = 0; r < BS_ITEMS; r++) {
        if (r * BS_THREADS >= n) break;  // uniform: rounds past the tile's length are skipped, not predicated off
        const int idx = r * BS_THREADS + tid;
        if (idx < n) {
            const uint32_t id = point_list[rg.x + idx];   // the scatter wrote bare ids: the depth half of the key is gathered here
            k[r] = ((unsigned long long)__float_as_uint(depths[id]) << 32) | id;
            const float d = __uint_as_float((uint32_t)(k[r] >> 32));
            finite = finite && (d >= 0.0f) && (d < CUDART_INF_F);
            lo = fminf(lo, d);
            hi = fmaxf(hi, d);
        }
    }
    for (int e = tid; e < BS_NB; e += BS_THREADS) s_cnt[e] = 0;
    if (tid == 0) { s_lohi[0] = CUDART_INF_F; s_lohi[1] = 0.0f; }
    // a tile this path cannot take (non-finite / negative depths) is declined by every thread together
    if (__syncthreads_or(!finite)) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) {  // non-negative floats order like their bit patterns
        atomicMin(reinterpret_cast<unsigned*>(&s_lohi[0]), __float_as_uint(lo));
        atomicMax(reinterpret_cast<unsigned*>(&s_lohi[1]), __float_as_uint(hi));
    }
    __syncthreads();
    const float dmin = s_lohi[0], dmax = s_lohi[1];
    const float inv = dmax > dmin ? (float)BS_NB / (dmax - dmin) : 0.0f;
    auto bucket_of = [&](unsigned long long key) -> int {
        const float d = __uint_as_float((uint32_t)(key >> 32));
        return min(BS_NB - 1, __float2int_rz(fmul(fsub(d, dmin), inv)));
    };
#pragma unroll
    for (int r = 0; r < BS_ITEMS; r++)
        if (r * BS_THREADS >= n) break; else if (r * BS_THREADS + tid < n) atomicAdd(&s_cnt[bucket_of(k[r])], 1u);
    __syncthreads();
    // exclusive scan of the counters (thread t owns BS_NB / BS_THREADS consecutive ones) + largest bucket
    constexpr int PER = BS_NB / BS_THREADS;
    uint32_t c[PER], sum = 0, big = 0;
#pragma unroll
    for (int q = 0; q < PER; q++) { c[q] = s_cnt[tid * PER + q]; sum += c[q]; big = max(big, c[q]); }
    if (__syncthreads_or(big > (uint32_t)BS_LIMIT)) return;  // clumped depths: the radix kernel sorts this tile
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t base = incl - sum;
    for (int w = 0; w < wid; w++) base += s_warp[w];
#pragma unroll
    for (int q = 0; q < PER; q++) { s_cnt[tid * PER + q] = base; base += c[q]; }
    __syncthreads();
    // distribute: afterwards s_cnt[b] is the END of bucket b (its start is the end of bucket b - 1)
#pragma unroll
    for (int r = 0; r < BS_ITEMS; r++)
        if (r * BS_THREADS >= n) break; else if (r * BS_THREADS + tid < n) s_key[atomicAdd(&s_cnt[bucket_of(k[r])], 1u)] = k[r];
    __syncthreads();
    // exact position inside the bucket = number of its keys that compare lower
    uint32_t pos[BS_ITEMS];
#pragma unroll
    for (int r = 0; r < BS_ITEMS; r++) {
        if (r * BS_THREADS >= n) break;  // uniform: rounds past the tile's length are skipped, not predicated off
        if (r * BS_THREADS + tid < n) {
            const int b = bucket_of(k[r]);
            const uint32_t s0 = b ? s_cnt[b - 1] : 0u, s1 = s_cnt[b];
            uint32_t p = s0;
            for (uint32_t q = s0; q < s1; q++) p += s_key[q] < k[r] ? 1u : 0u;
            pos[r] = p;
        }
    }
    __syncthreads();  // every thread has finished reading the bucketed array: overwrite it in final order
#pragma unroll
    for (int r = 0; r < BS_ITEMS; r++)
        if (r * BS_THREADS >= n) break; else if (r * BS_THREADS + tid < n) s_key[pos[r]] = k[r];
    __syncthreads();
    for (int i = tid; i < n; i += BS_THREADS) {  // coalesced write-out
        const unsigned long long v = s_key[i];
        if (write_keys) g[i] = v;  // the sorted keys are only read back by tests / debugging (ols_ws_view.d_keys)
        point_list[rg.x + i] = (uint32_t)v;
    }
    if (tid == 0) tile_count[blockIdx.x] = BS_DONE;
}

// ---------------------------------------------------------------------------------------------------
// Per-tile sort, main path: least-significant-digit radix sort of the bucket in shared memory.
// Keys are (depth_bits << 32 | gaussian id); four stable 8-bit passes over the depth bits, ranks from
// warp match/ballot primitives and per-warp digit counters, then a neighbour fix-up that orders equal
// depths by id (the order cub's stable sort gives the reference, since duplicateWithKeys emits ids in
// ascending order).  Buckets longer than RS_CAP fall through to the bitonic kernel below.
// ---------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_CAP = 4096;
constexpr int RS_ITEMS = RS_CAP / RS_THREADS;  // 16 keys per thread at most

__global__ void __launch_bounds__(RS_THREADS) k_sort_tiles_radix(const __grid_constant__ PassBatch pb, const WsLayout L,
                                                                const int write_keys) {
    __shared__ unsigned long long s_key[RS_CAP];       // 32 KB
    char* ws = pb.v[blockIdx.y].ws;
    unsigned long long* __restrict__ keys = (unsigned long long*)(ws + L.keys);
    uint32_t* __restrict__ point_list = (uint32_t*)(ws + L.point_list);
    const float* __restrict__ depths = pb.v[blockIdx.y].depths;
    const uint2* __restrict__ ranges = (const uint2*)(ws + L.ranges);
    const uint32_t* __restrict__ tile_count = (const uint32_t*)(ws + L.tile_count);
    const DeviceInfo* __restrict__ info = (const DeviceInfo*)(ws + L.info);
    __shared__ uint32_t s_cnt[RS_WARPS][256];          // 8 KB  per-warp digit counters / bases
    __shared__ uint32_t s_scan[RS_WARPS];
    if (info->overflow) return;
    if (tile_count[blockIdx.x] == BS_DONE) return;  // sorted by the bucket kernel
    const uint2 rg = ranges[blockIdx.x];
    const int n = (int)(rg.y - rg.x);
    if (n <= 0 || n > RS_CAP) return;
    unsigned long long* g = keys + rg.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int items = (n + RS_THREADS - 1) / RS_THREADS;  // rounds per warp (uniform)
    const int seg = items * 32;                           // keys per warp, contiguous -> stable
    const unsigned lt_mask = (1u << lane) - 1u;

    unsigned long long k[RS_ITEMS];
    uint32_t dmin = 0xffffffffu, dmax = 0u;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const int idx = wid * seg + r * 32 + lane;
        const bool real = r < items && idx < n;
        if (real) {
            const uint32_t id = point_list[rg.x + idx];
            k[r] = ((unsigned long long)__float_as_uint(depths[id]) << 32) | id;
        } else {
            k[r] = ~0ull;  // +inf padding keeps the tail in place
        }
        if (real) { dmin = min(dmin, (uint32_t)(k[r] >> 32)); dmax = max(dmax, (uint32_t)(k[r] >> 32)); }
    }
    // digits are taken from (depth_bits - tile minimum): only the bytes that differ inside the tile are sorted
    dmin = __reduce_min_sync(0xffffffffu, dmin);
    dmax = __reduce_max_sync(0xffffffffu, dmax);
    if (tid == 0) { s_scan[0] = 0xffffffffu; s_scan[1] = 0u; }
    __syncthreads();
    if (lane == 0) { atomicMin(&s_scan[0], dmin); atomicMax(&s_scan[1], dmax); }
    __syncthreads();
    dmin = s_scan[0];
    const uint32_t spread = s_scan[1] - dmin;
    const int n_pass = spread == 0 ? 0 : (32 - __clz(spread) + 7) / 8;
    __syncthreads();
    for (int pass = 0; pass < n_pass; pass++) {
        const int shift = 8 * pass;
        for (int e = tid; e < RS_WARPS * 256; e += RS_THREADS) (&s_cnt[0][0])[e] = 0;
        __syncthreads();
        uint32_t rk[RS_ITEMS];  // sweep 1: (digit | rank among warp peers << 8 | leader << 16); sweep 2: final rank
#pragma unroll
        for (int r = 0; r < RS_ITEMS; r++) {
            if (r < items) {
                const bool pad = k[r] == ~0ull;
                const unsigned d = pad ? 255u : ((((uint32_t)(k[r] >> 32) - dmin) >> shift) & 255u);
                unsigned peers = 0xffffffffu;  // lanes holding the same digit (8 ballots; no MATCH instruction)
#pragma unroll
                for (int bit = 0; bit < 8; bit++) {
                    const bool on = (d >> bit) & 1u;
                    const unsigned m = __ballot_sync(0xffffffffu, on);
                    peers &= on ? m : ~m;
                }
                const unsigned below = __popc(peers & lt_mask);
                const bool leader = below == 0;
                if (leader) s_cnt[wid][d] += __popc(peers);
                rk[r] = d | (below << 8) | ((leader ? __popc(peers) : 0u) << 16);
                __syncwarp();
            }
        }
        __syncthreads();
        // exclusive scan in (digit major, warp minor) order; thread d owns digit d
        uint32_t tot = 0;
        {
            uint32_t c[RS_WARPS];
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) { c[w] = s_cnt[w][tid]; }
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) { const uint32_t v = c[w]; c[w] = tot; tot += v; }
            // all real keys share this digit -> the pass is the identity (uniform decision)
            const uint32_t n_pad = (uint32_t)(items * RS_THREADS - n);
            const int skip = __syncthreads_or(tot == (uint32_t)n + (tid == 255 ? n_pad : 0u));
            if (skip) continue;
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_scan[wid] = incl;
            __syncthreads();
            uint32_t base = incl - tot;
            for (int w = 0; w < wid; w++) base += s_scan[w];
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) s_cnt[w][tid] = base + c[w];
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RS_ITEMS; r++) {
            if (r < items) {
                const unsigned d = rk[r] & 255u, below = (rk[r] >> 8) & 255u, cnt = rk[r] >> 16;
                rk[r] = s_cnt[wid][d] + below;
                __syncwarp();
                if (cnt) s_cnt[wid][d] += cnt;
                __syncwarp();
            }
        }
#pragma unroll
        for (int r = 0; r < RS_ITEMS; r++)
            if (r < items) s_key[rk[r]] = k[r];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RS_ITEMS; r++)
            if (r < items) k[r] = s_key[wid * seg + r * 32 + lane];
        __syncthreads();
    }
    // final order into shared memory, then order runs of equal depth by id (odd-even transposition)
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++)
        if (r < items) s_key[wid * seg + r * 32 + lane] = k[r];
    __syncthreads();
    for (;;) {
        int swapped = 0;
        for (int phase = 0; phase < 2; phase++) {
            for (int i = 2 * tid + phase; i + 1 < n; i += 2 * RS_THREADS) {
                const unsigned long long x = s_key[i], y = s_key[i + 1];
                if (x > y) { s_key[i] = y; s_key[i + 1] = x; swapped = 1; }
            }
            __syncthreads();
        }
        if (!__syncthreads_or(swapped)) break;
    }
    for (int i = tid; i < n; i += RS_THREADS) {
        const unsigned long long v = s_key[i];
        if (write_keys) g[i] = v;
        point_list[rg.x + i] = (uint32_t)v;
    }
}

// Per-tile sort.  All comparators are ascending ("mirror" first step per stage), so positions >= n
// behave as +inf padding without being stored; tiles longer than the shared-memory chunk run the
// wide strides in global memory and the narrow ones chunk by chunk in shared memory.
constexpr int SORT_THREADS = 256;
constexpr int SORT_CHUNK = 4096;  // keys per shared-memory chunk (32 KB)

__device__ __forceinline__ void cmpswap(unsigned long long* s, int a, int b, int n) {
    if (b < n) {
        const unsigned long long va = s[a], vb = s[b];
        if (va > vb) { s[a] = vb; s[b] = va; }
    }
}
__device__ __forceinline__ int pow2_cover(int n) {
    int span = 1;
    while (span < n) span <<= 1;
    return span;
}
// mirror step of stage k over `half_pairs` comparator slots: (blk*k + off, blk*k + k-1-off)
__device__ __forceinline__ void bitonic_mirror(unsigned long long* s, int k, int half_pairs, int n) {
    const int sh = __ffs(k) - 2;  // log2(k/2)
    for (int i = threadIdx.x; i < half_pairs; i += SORT_THREADS) {
        const int blk = i >> sh, off = i & ((k >> 1) - 1);
        cmpswap(s, (blk << (sh + 1)) + off, (blk << (sh + 1)) + (k - 1 - off), n);
    }
    __syncthreads();
}
// half-cleaner step with stride j: (a, a|j) for every a with bit j clear
__device__ __forceinline__ void bitonic_step(unsigned long long* s, int j, int half_pairs, int n) {
    for (int i = threadIdx.x; i < half_pairs; i += SORT_THREADS) {
        const int a = ((i & ~(j - 1)) << 1) | (i & (j - 1));
        cmpswap(s, a, a | j, n);
    }
    __syncthreads();
}
// complete sort of n <= SORT_CHUNK keys held in shared memory
__device__ void smem_sort_full(unsigned long long* s, int n) {
    const int span = pow2_cover(n);
    for (int k = 2; k <= span; k <<= 1) {
        bitonic_mirror(s, k, span >> 1, n);
        for (int j = k >> 2; j > 0; j >>= 1) bitonic_step(s, j, span >> 1, n);
    }
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_tiles(const __grid_constant__ PassBatch pb, const WsLayout L, int min_len) {
    __shared__ unsigned long long s[SORT_CHUNK];
    char* ws = pb.v[blockIdx.y].ws;
    unsigned long long* __restrict__ keys = (unsigned long long*)(ws + L.keys);
    uint32_t* __restrict__ point_list = (uint32_t*)(ws + L.point_list);
    const float* __restrict__ depths = pb.v[blockIdx.y].depths;
    const uint2* __restrict__ ranges = (const uint2*)(ws + L.ranges);
    const DeviceInfo* __restrict__ info = (const DeviceInfo*)(ws + L.info);
    if (info->overflow) return;
    const uint2 rg = ranges[blockIdx.x];
    const int n = (int)(rg.y - rg.x);
    if (n <= min_len) return;
    unsigned long long* g = keys + rg.x;
    const int tid = threadIdx.x;
    // materialise the 64-bit keys of this (long) tile from the ids the scatter wrote
    for (int i = tid; i < n; i += SORT_THREADS) {
        const uint32_t id = point_list[rg.x + i];
        g[i] = ((unsigned long long)__float_as_uint(depths[id]) << 32) | id;
    }
    __syncthreads();
    if (n <= SORT_CHUNK) {
        for (int i = tid; i < n; i += SORT_THREADS) s[i] = g[i];
        __syncthreads();
        smem_sort_full(s, n);
        for (int i = tid; i < n; i += SORT_THREADS) {
            const unsigned long long v = s[i];
            g[i] = v;
            point_list[rg.x + i] = (uint32_t)v;
        }
        return;
    }
    // long tile: strides >= SORT_CHUNK run in global memory, the rest chunk by chunk in shared memory
    const int span = pow2_cover(n);
    const int n_chunks = (n + SORT_CHUNK - 1) / SORT_CHUNK;
    for (int c = 0; c < n_chunks; c++) {
        const int cb = c * SORT_CHUNK, nl = min(SORT_CHUNK, n - cb);
        for (int i = tid; i < nl; i += SORT_THREADS) s[i] = g[cb + i];
        __syncthreads();
        smem_sort_full(s, nl);
        for (int i = tid; i < nl; i += SORT_THREADS) g[cb + i] = s[i];
        __syncthreads();
    }
    for (int k = SORT_CHUNK * 2; k <= span; k <<= 1) {
        bitonic_mirror(g, k, span >> 1, n);
        for (int j = k >> 2; j >= SORT_CHUNK; j >>= 1) bitonic_step(g, j, span >> 1, n);
        for (int c = 0; c < n_chunks; c++) {
            const int cb = c * SORT_CHUNK, nl = min(SORT_CHUNK, n - cb);
            for (int i = tid; i < nl; i += SORT_THREADS) s[i] = g[cb + i];
            __syncthreads();
            for (int j = SORT_CHUNK >> 1; j > 0; j >>= 1) bitonic_step(s, j, SORT_CHUNK >> 1, nl);
            for (int i = tid; i < nl; i += SORT_THREADS) g[cb + i] = s[i];
            __syncthreads();
        }
    }
    for (int i = tid; i < n; i += SORT_THREADS) point_list[rg.x + i] = (uint32_t)g[i];
}

// ---------------------------------------------------------------------------------------------------
// Forward blend (forward.cu:377-513).  The tile's Gaussian records are streamed global -> shared with cp.async in
// double-buffered batches.  Every warp first tests 32 records at a time -- one per lane -- against its own pixel block
// using the record's conservative half-extents (a Gaussian that cannot reach alpha >= 1/255 anywhere in the block is
// skipped by every pixel of the reference as well), ballots, and then runs the per-pixel evaluation only for the
// surviving records.  The channel accumulators are updated with packed FFMA2.
// ---------------------------------------------------------------------------------------------------
constexpr int BLEND_THREADS = 256;
constexpr int BLEND_BATCH = 64;

struct BlendArgs {
    int W, H, gx;
    const uint2* ranges;
    const uint32_t* point_list;
    const float* records;
    const float* language;  // [P,F] caller's language rows (joint pass), else unused
    const float* bg;
    const DeviceInfo* info;
    float* final_T;
    uint32_t* n_contrib;
    float* out_color;
    float* out_language;
    float* out_depth;
    float* out_opacity;
    int32_t* n_touched;
    uint8_t* warp_hits;  // [R] per list entry: which of the tile's 8 pixel blocks blended it (read by the backward)
};

struct BlendBatch { BlendArgs v[OLS_MAX_VIEWS]; };

typedef unsigned long long f32x2;  // two floats in one 64-bit register pair (lo = first)
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {  // per-component fma.rn
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) {  // per-component mul.rn
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// NCOL = 3: the pass blends rgb (+ background) and depth; NCOL = 0: language channels only (second pass of D/).
// ---------------------------------------------------------------------------------------------------
// Forward blend, two pixels per lane.  One CTA of 128 threads per tile; each of the 4 warps owns an
// 8x8 pixel block -- lane l: column (l & 7), rows (l >> 3) and (l >> 3) + 4 -- so everything of a (block, Gaussian)
// step that does not depend on the pixel row (list walk, the two header loads and the five channel loads from
// shared memory, dx, dx*A, dx*B) is paid once per TWO pixels, and each lane carries two independent dependency
// chains.  Per-pixel arithmetic, thresholds and accumulation order are, in BITEXACT mode, those of the compiled
// reference (forward.cu:437-483), so every output bit is equal to the reference's.  The two 8x4 halves of a warp's block
// are the eight 8x4 pixel blocks the hit byte is defined over: bit w = block w = ((y / 4) * 2 + x / 8) blended the entry.  Without BITEXACT alpha uses ex2.approx(power * log2 e) (2 instructions instead of expf's 8; relative
// error of alpha <= 4e-7, tests/test_forward_gpu.py states the tolerance); the backward then uses the same function
// so that its blend / skip decisions agree with the forward's.
// ---------------------------------------------------------------------------------------------------
constexpr int B2_THREADS = 128;
#ifndef B2_MIN_BLOCKS
#define B2_MIN_BLOCKS 5
#endif
__device__ __forceinline__ float fast_exp(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}

template <int TILE, int NCOL, int F, bool BITEXACT, int BB = BLEND_BATCH>
__global__ void __launch_bounds__(B2_THREADS, B2_MIN_BLOCKS) k_blend2(const __grid_constant__ BlendBatch bb) {
    const BlendArgs& a = bb.v[blockIdx.y];
    static_assert(TILE <= 16, "4 warps of 8x8 pixels cover at most 16x16");
    static_assert(NCOL == 0 || NCOL == 3, "colour channels");
    constexpr int NCH = NCOL + F;
    constexpr int REC = rec_floats_nch(NCH);
    using Stage = RecordStage<NCOL, F>;
    constexpr int OPS = Stage::OPS;
    constexpr int NPAIR = (NCH + 1) / 2;
    constexpr int EXT = REC - 2;
    constexpr int NHALF = BB / 32;
    __shared__ __align__(16) float s_rec[2][BB * REC];
    __shared__ uint32_t s_id[2][BB];
    __shared__ uint32_t s_hit[2][8][NHALF];  // per 8x4 pixel block (k_blend's warp index): bit j = the block blended entry j

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile_x = blockIdx.x % a.gx, tile_y = blockIdx.x / a.gx;
    const int bx0 = (wid & 1) * 8, by0 = (wid >> 1) * 8;
    const int lx = bx0 + (lane & 7);
    const int pxi = tile_x * TILE + lx;
    int ly[2], pyi[2];
    bool inside[2], done[2];
    float pfy[2];
#pragma unroll
    for (int p = 0; p < 2; p++) {
        ly[p] = by0 + (lane >> 3) + 4 * p;
        pyi[p] = tile_y * TILE + ly[p];
        inside[p] = lx < TILE && ly[p] < TILE && pxi < a.W && pyi[p] < a.H;
        done[p] = !inside[p];
        pfy[p] = (float)pyi[p];
    }
    const float pfx = (float)pxi;
    // block ids of the two halves in k_blend's numbering: ((y / 4) * 2 + x / 8)
    const int blk0 = ((wid >> 1) * 2) * 2 + (wid & 1), blk1 = blk0 + 2;
    if (tid < 2 * 8 * NHALF) (&s_hit[0][0][0])[tid] = 0u;
    // this warp's pixel rectangle, clipped to the tile and the image
    const float fx0 = (float)(tile_x * TILE + bx0), fy0 = (float)(tile_y * TILE + by0);
    const float fx1 = (float)min(min(tile_x * TILE + bx0 + 7, tile_x * TILE + TILE - 1), a.W - 1);
    const float fy1 = (float)min(min(tile_y * TILE + by0 + 7, tile_y * TILE + TILE - 1), a.H - 1);

    uint2 rg = a.ranges[blockIdx.x];
    const bool overflow = a.info->overflow != 0;  // instance capacity exceeded: nothing was binned; the images become NaN
    if (overflow) rg = make_uint2(0u, 0u);
    const int total = (int)(rg.y - rg.x);
    const int n_batches = (total + BB - 1) / BB;

    auto flush_hits = [&](int b) {
        if (tid < BB) {
            uint32_t byte = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) byte |= ((s_hit[b & 1][w][tid >> 5] >> (tid & 31)) & 1u) << w;
            const int e = b * BB + tid;
            if (e < total) a.warp_hits[rg.x + e] = (uint8_t)byte;
        }
    };
    auto issue = [&](int b) {
        const int cnt = min(BB, total - b * BB);
        const int buf = b & 1;
        constexpr int TPE = B2_THREADS / BB;
        const int g = tid / TPE;
        if (g < cnt) {
            const uint32_t id = a.point_list[rg.x + b * BB + g];
            if ((tid % TPE) == 0) s_id[buf][g] = id;
#pragma unroll
            for (int q = tid % TPE; q < OPS; q += TPE) Stage::copy(&s_rec[buf][g * REC], a.records, a.language, id, q);
        }
        cp_async_commit();
    };

    float T[2] = {1.0f, 1.0f};
    f32x2 acc2[2][NPAIR];
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
        for (int c = 0; c < NPAIR; c++) acc2[p][c] = 0ull;
    float acc_d[2] = {0.0f, 0.0f};
    uint32_t last_contributor[2] = {0u, 0u};

    if (n_batches > 0) issue(0);
    int b = 0;
    for (; b < n_batches; b++) {
        cp_async_wait<0>();
        if (__syncthreads_count(done[0] && done[1]) == B2_THREADS) break;
        if (b > 0) flush_hits(b - 1);
        if (b + 1 < n_batches) issue(b + 1);
        const int cnt = min(BB, total - b * BB);
        const float* rec = s_rec[b & 1];
        const uint32_t* ids = s_id[b & 1];
        const uint32_t cbase = (uint32_t)b * BB;
#pragma unroll 1
        for (int half = 0; half < NHALF; half++) {
            uint32_t myhits[2] = {0u, 0u}, mytouch[2] = {0u, 0u};
            if (__all_sync(0xffffffffu, done[0] && done[1])) {
                if (lane == 0) { s_hit[b & 1][blk0][half] = 0u; s_hit[b & 1][blk1][half] = 0u; }
                continue;
            }
            const int e = half * 32 + lane;
            bool hit = false;
            if (e < cnt) {
                const float2 c = *reinterpret_cast<const float2*>(rec + e * REC + REC_X);
                const float2 h = *reinterpret_cast<const float2*>(rec + e * REC + EXT);
                hit = (c.x + h.x >= fx0) && (c.x - h.x <= fx1) && (c.y + h.y >= fy0) && (c.y - h.y <= fy1);
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {
                const int jl = __ffs(m) - 1;
                const int j = half * 32 + jl;
                m &= m - 1;
                const float* rj = rec + j * REC;
                const uint32_t jbit = 1u << jl;
                const float4 g0 = *reinterpret_cast<const float4*>(rj);      // x y A B
                const float4 g1 = *reinterpret_cast<const float4*>(rj + 4);  // C op pth depth
                const float dx = fsub(g0.x, pfx);
                const float dxa = fmul(dx, g0.z), dxb = fmul(dx, g0.w);
                float power[2];
                bool pass[2];
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    const float dy = fsub(g0.y, pfy[p]);
                    power[p] = ffma(ffma(dx, dxa, fmul(dy, fmul(dy, g1.x))), -0.5f, -fmul(dy, dxb));
                    pass[p] = !done[p] && !(power[p] > 0.0f) && !(power[p] < g1.z);
                }
                if (pass[0] || pass[1]) {
                    f32x2 v[NPAIR];
#pragma unroll
                    for (int q = 0; q + 1 < NPAIR; q += 2) {
                        const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(rj + REC_CH + 2 * q);
                        v[q] = t.x;
                        v[q + 1] = t.y;
                    }
                    if (NPAIR & 1) v[NPAIR - 1] = *reinterpret_cast<const f32x2*>(rj + REC_CH + 2 * (NPAIR - 1));
#pragma unroll
                    for (int p = 0; p < 2; p++) {
                        if (pass[p]) {
                            const float G = BITEXACT ? expf(power[p]) : fast_exp(power[p]);
                            const float alpha = fminf(fmul(g1.y, G), 0.99f);
                            if (!(alpha < 1.0f / 255.0f)) {
                                const float test_T = fmul(T[p], fsub(1.0f, alpha));
                                if (test_T < 0.0001f) {
                                    done[p] = true;
                                } else {
                                    if (BITEXACT) {  // acc = fma(T, alpha * c, acc) like the compiled reference
                                        const f32x2 a2 = pack2(alpha, alpha), T2 = pack2(T[p], T[p]);
#pragma unroll
                                        for (int q = 0; q < NPAIR; q++) acc2[p][q] = ffma2(T2, fmul2(a2, v[q]), acc2[p][q]);
                                        if (NCOL) acc_d[p] = ffma(T[p], fmul(alpha, g1.w), acc_d[p]);
                                    } else {
                                        const float w = fmul(alpha, T[p]);
                                        const f32x2 w2 = pack2(w, w);
#pragma unroll
                                        for (int q = 0; q < NPAIR; q++) acc2[p][q] = ffma2(w2, v[q], acc2[p][q]);
                                        if (NCOL) acc_d[p] = ffma(w, g1.w, acc_d[p]);
                                    }
                                    myhits[p] |= jbit;
                                    if (test_T > 0.5f) mytouch[p] |= jbit;
                                    T[p] = test_T;
                                }
                            }
                        }
                    }
                }
            }
            // once per 32 entries: the pixel's last contributor so far (its highest hit bit), which entries each 8x4 half
            // blended (for the backward), and the n_touched counts
#pragma unroll
            for (int p = 0; p < 2; p++)
                if (myhits[p]) last_contributor[p] = cbase + (uint32_t)(half * 32) + (32u - (uint32_t)__clz(myhits[p]));
            const uint32_t h0 = __reduce_or_sync(0xffffffffu, myhits[0]), h1 = __reduce_or_sync(0xffffffffu, myhits[1]);
            if (lane == 0) { s_hit[b & 1][blk0][half] = h0; s_hit[b & 1][blk1][half] = h1; }
            for (uint32_t tmask = __reduce_or_sync(0xffffffffu, mytouch[0] | mytouch[1]); tmask; tmask &= tmask - 1) {
                const int jl = __ffs(tmask) - 1;
                const int n = (int)__reduce_add_sync(0xffffffffu, ((mytouch[0] >> jl) & 1u) + ((mytouch[1] >> jl) & 1u));
                if (lane == 0) atomicAdd(&a.n_touched[ids[half * 32 + jl]], n);
            }
        }
    }
    cp_async_wait<0>();
    if (b > 0) {  // the last batch that was worked on (a `break` leaves before flushing it)
        __syncthreads();
        flush_hits(b - 1);
    }
    const size_t HW = (size_t)a.W * a.H;
#pragma unroll
    for (int p = 0; p < 2; p++) {
        if (inside[p]) {
            const size_t pix = (size_t)pyi[p] * a.W + pxi;
            float acc[2 * NPAIR];
#pragma unroll
            for (int q = 0; q < NPAIR; q++) unpack2(acc2[p][q], acc[2 * q], acc[2 * q + 1]);
            if (overflow) {  // never hand back a plausible-looking empty image (the wrapper re-renders with the exact capacity)
#pragma unroll
                for (int q = 0; q < 2 * NPAIR; q++) acc[q] = CUDART_NAN_F;
                acc_d[p] = CUDART_NAN_F;
                T[p] = CUDART_NAN_F;
            }
            a.final_T[pix] = T[p];
            a.n_contrib[pix] = last_contributor[p];
            if (NCOL) {
#pragma unroll
                for (int c = 0; c < NCOL; c++) a.out_color[c * HW + pix] = ffma(a.bg[c], T[p], acc[c]);
                a.out_depth[pix] = acc_d[p];
            }
#pragma unroll
            for (int c = 0; c < F; c++) a.out_language[c * HW + pix] = acc[NCOL + c];
            a.out_opacity[pix] = fsub(1.0f, T[p]);
        }
    }
}

template <int TILE, int NCOL, int F>
static int launch_blend(const BlendBatch& ba, int n_tiles, int V, bool bitexact, cudaStream_t st) {
    const dim3 grid(n_tiles, V);
    if (bitexact)
        k_blend2<TILE, NCOL, F, true><<<grid, B2_THREADS, 0, st>>>(ba);
    else
        k_blend2<TILE, NCOL, F, false><<<grid, B2_THREADS, 0, st>>>(ba);
    return 0;
}

}  // namespace ols

using namespace ols;

#define OLS_DEBUG_SYNC(name)                                                              \
    do {                                                                                  \
        OLS_CUDA_TRY(cudaGetLastError());                                                 \
        if (debug) {                                                                      \
            cudaError_t _e = cudaStreamSynchronize(st);                                   \
            if (_e != cudaSuccess) {                                                      \
                ols_set_error("kernel %s failed: %s", name, cudaGetErrorString(_e));      \
                return OLS_ERR_CUDA;                                                      \
            }                                                                             \
        }                                                                                 \
    } while (0)

static int set_hist_smem(size_t hist_smem, int n_tiles) {
    if (hist_smem > 40 * 1024) {
        if (hist_smem > 200 * 1024) {
            ols_set_error("image too large: %d tiles exceed the shared-memory tile histogram", n_tiles);
            return OLS_ERR_UNSUPPORTED;
        }
        OLS_CUDA_TRY(cudaFuncSetAttribute(k_preprocess, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem));
        OLS_CUDA_TRY(cudaFuncSetAttribute(k_preprocess_dis, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem));
        OLS_CUDA_TRY(cudaFuncSetAttribute(k_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem));
    }
    return OLS_OK;
}

// One footprint's binning + per-tile sort + blend (everything after preprocess) for the V views of a batch
// (grid.y = view).  `L` is the layout every view's workspace shares; `depths` of a view may live in another pass'
// workspace (D/ shares them between its two lists).
static int run_pass(int P, int W, int H, int tile, int ncol, int F, unsigned flags, int64_t R_cap, const PassBatch& pb, int V,
                    const WsLayout& L, const float* language, cudaStream_t st) {
    const bool debug = (flags & OLS_FLAG_DEBUG) != 0;
    const size_t hist_smem = sizeof(uint32_t) * (size_t)L.n_tiles;
    k_tile_offsets<<<dim3((L.n_tiles + 31) / 32, V), 32 * TO_WARPS, 0, st>>>(pb, L);
    OLS_DEBUG_SYNC("tile_offsets");
    k_tile_scan<<<dim3(1, V), SCAN_THREADS, 0, st>>>(pb, L, (unsigned long long)R_cap);
    OLS_DEBUG_SYNC("tile_scan");
    k_scatter<<<dim3(L.n_ctas, V), PRE_THREADS, hist_smem, st>>>(pb, L, P);
    OLS_DEBUG_SYNC("scatter");
    ols_timing_mark(OLS_T_BINNING, st);
    const int write_keys = debug ? 1 : 0;  // sorted 64-bit keys are a debugging view; the blend passes read point_list
    k_sort_tiles_bucket<<<dim3(L.n_tiles, V), BS_THREADS, 0, st>>>(pb, L, write_keys);
    OLS_DEBUG_SYNC("sort_tiles_bucket");
    k_sort_tiles_radix<<<dim3(L.n_tiles, V), RS_THREADS, 0, st>>>(pb, L, write_keys);
    OLS_DEBUG_SYNC("sort_tiles_radix");
    // buckets longer than the radix kernel's shared-memory capacity (rare): bitonic fallback
    k_sort_tiles<<<dim3(L.n_tiles, V), SORT_THREADS, 0, st>>>(pb, L, RS_CAP);
    OLS_DEBUG_SYNC("sort_tiles");
    ols_timing_mark(OLS_T_SORT, st);

    BlendBatch bb;
    for (int v = 0; v < V; v++) {
        BlendArgs& ba = bb.v[v];
        char* ws = pb.v[v].ws;
        ba.W = W; ba.H = H; ba.gx = L.gx;
        ba.ranges = (const uint2*)(ws + L.ranges); ba.point_list = (const uint32_t*)(ws + L.point_list);
        ba.records = (const float*)(ws + L.records); ba.language = language; ba.bg = pb.v[v].bg;
        ba.info = (const DeviceInfo*)(ws + L.info);
        ba.final_T = (float*)(ws + L.final_T); ba.n_contrib = (uint32_t*)(ws + L.n_contrib);
        ba.out_color = pb.v[v].color; ba.out_language = pb.v[v].language; ba.out_depth = pb.v[v].depth;
        ba.out_opacity = pb.v[v].opacity; ba.n_touched = pb.v[v].n_touched;
        ba.warp_hits = (uint8_t*)(ws + L.warp_hits);
    }
    const bool bitexact = (flags & OLS_FLAG_BITEXACT_BLEND) != 0;
    const int key = tile * 10000 + ncol * 100 + F;
    switch (key) {
        case 150315: launch_blend<15, 3, 15>(bb, L.n_tiles, V, bitexact, st); break;
        case 160315: launch_blend<16, 3, 15>(bb, L.n_tiles, V, bitexact, st); break;
        case 150303: launch_blend<15, 3, 3>(bb, L.n_tiles, V, bitexact, st); break;
        case 160303: launch_blend<16, 3, 3>(bb, L.n_tiles, V, bitexact, st); break;
        case 150300: launch_blend<15, 3, 0>(bb, L.n_tiles, V, bitexact, st); break;   // D/ colour pass
        case 160300: launch_blend<16, 3, 0>(bb, L.n_tiles, V, bitexact, st); break;
        case 150003: launch_blend<15, 0, 3>(bb, L.n_tiles, V, bitexact, st); break;   // D/ language pass
        case 160003: launch_blend<16, 0, 3>(bb, L.n_tiles, V, bitexact, st); break;
        case 150015: launch_blend<15, 0, 15>(bb, L.n_tiles, V, bitexact, st); break;
        case 160015: launch_blend<16, 0, 15>(bb, L.n_tiles, V, bitexact, st); break;
        default:
            ols_set_error("unsupported (tile=%d, F=%d): compiled variants are tile in {15,16} x F in {3,15}", tile, F);
            return OLS_ERR_UNSUPPORTED;
    }
    OLS_DEBUG_SYNC("blend");
    ols_timing_mark(OLS_T_BLEND_FWD, st);
    return OLS_OK;
}

// Views per CTA of k_preprocess: as many as fit a 64 KB budget of tile histograms, spread evenly over the groups.
static int preprocess_view_group(int V, int n_tiles) {
    const size_t per_view = sizeof(uint32_t) * (size_t)n_tiles + sizeof(float) * PRE_CAM;
    int vg_max = (int)((64 * 1024) / per_view);
    if (vg_max < 1) vg_max = 1;
    const int groups = (V + vg_max - 1) / vg_max;
    return (V + groups - 1) / groups;
}

// Forward of V views of the same Gaussians (V = 1: the reference's call).  The views share every size, the flags and
// the Gaussian parameter pointers (validated by the caller); matrices, field of view, background, outputs and the
// workspace are per view.
int ols_launch_forward(const ols_raster_args* views, const ols_fwd_out* outs, int V, const WsLayout& L, cudaStream_t st) {
    const ols_raster_args* a = &views[0];
    const bool debug = (a->flags & OLS_FLAG_DEBUG) != 0;
    if (a->F != 3 && a->F != 15) {
        ols_set_error("unsupported (tile=%d, F=%d): compiled variants are tile in {15,16} x F in {3,15}", a->tile, a->F);
        return OLS_ERR_UNSUPPORTED;
    }
    PreBatch vb;
    PassBatch pb;
    for (int v = 0; v < V; v++) {
        char* ws = (char*)views[v].d_workspace;
        OLS_CUDA_TRY(cudaMemsetAsync(ws + L.info, 0, 256, st));
        OLS_CUDA_TRY(cudaMemsetAsync(outs[v].d_n_touched, 0, sizeof(int32_t) * (size_t)a->P, st));
        PreView& pv = vb.v[v];
        pv.viewmatrix = views[v].d_viewmatrix; pv.projmatrix = views[v].d_projmatrix; pv.campos = views[v].d_campos;
        pv.tanfovx = views[v].tanfovx; pv.tanfovy = views[v].tanfovy;
        pv.focal_y = a->H / (2.0f * views[v].tanfovy);  // rasterizer_impl.cu:394-395
        pv.focal_x = a->W / (2.0f * views[v].tanfovx);
        pv.ws = ws; pv.radii = outs[v].d_radii;
        PassView& q = pb.v[v];
        q.ws = ws; q.depths = (const float*)(ws + L.depths); q.bg = views[v].d_bg;
        q.color = outs[v].d_color; q.language = outs[v].d_language; q.depth = outs[v].d_depth; q.opacity = outs[v].d_opacity;
        q.n_touched = outs[v].d_n_touched;
    }
    PreArgs p;
    p.P = a->P; p.F = a->F; p.sh_degree = a->sh_degree; p.M = a->M; p.W = a->W; p.H = a->H; p.tile = a->tile;
    p.gx = L.gx; p.gy = L.gy; p.rec = L.rec; p.V = V; p.chunk = L.chunk; p.flags = a->flags;
    p.scale_modifier = a->scale_modifier;
    p.means3D = a->d_means3D; p.shs = a->d_shs; p.colors_precomp = a->d_colors_precomp;
    p.opacities = a->d_opacities; p.scales = a->d_scales; p.rotations = a->d_rotations;
    p.cov3D_precomp = a->d_cov3D_precomp;
    p.o_records = L.records; p.o_depths = L.depths; p.o_cov3D = L.cov3D; p.o_clamped = L.clamped;
    p.o_tiles_touched = L.tiles_touched; p.o_rect = L.rect; p.o_cta_hist = L.cta_hist; p.o_info = L.info;
    p.VG = preprocess_view_group(V, L.n_tiles);
    const size_t hist_smem = sizeof(uint32_t) * (size_t)L.n_tiles;
    const size_t pre_smem = (size_t)p.VG * (hist_smem + sizeof(float) * PRE_CAM);
    int rc = set_hist_smem(pre_smem, L.n_tiles);
    if (rc != OLS_OK) return rc;
    ols_timing_mark(-1, st);
    k_preprocess<<<dim3(L.n_ctas, (V + p.VG - 1) / p.VG), PRE_THREADS, pre_smem, st>>>(p, vb);
    OLS_DEBUG_SYNC("preprocess");
    ols_timing_mark(OLS_T_PREPROCESS, st);
    return run_pass(a->P, a->W, a->H, a->tile, 3, a->F, a->flags, a->R_cap, pb, V, L, a->d_language, st);
}

// Disentangled forward (D/rasterizer_impl.cu:364-620): one preprocess, then the colour footprint's and
// the language footprint's binning / sort / blend.  Lc / Ll are the two workspace layouts, Ll placed
// right after Lc in the same buffer.
int ols_launch_forward_dis(const ols_dis_args* d, const ols_dis_fwd_out* o, const WsLayout& Lc, const WsLayout& Ll,
                           size_t lang_base, cudaStream_t st) {
    const ols_raster_args* a = &d->base;
    char* wc = (char*)a->d_workspace;
    char* wl = wc + lang_base;
    const bool debug = (a->flags & OLS_FLAG_DEBUG) != 0;
    if (a->F != 3 && a->F != 15) {
        ols_set_error("unsupported (tile=%d, F=%d): compiled variants are tile in {15,16} x F in {3,15}", a->tile, a->F);
        return OLS_ERR_UNSUPPORTED;
    }
    OLS_CUDA_TRY(cudaMemsetAsync(wc + Lc.info, 0, 256, st));
    OLS_CUDA_TRY(cudaMemsetAsync(wl + Ll.info, 0, 256, st));
    OLS_CUDA_TRY(cudaMemsetAsync(o->d_n_touched, 0, sizeof(int32_t) * (size_t)a->P, st));
    OLS_CUDA_TRY(cudaMemsetAsync(o->d_n_touched_lang, 0, sizeof(int32_t) * (size_t)a->P, st));
    PreDisArgs p;
    p.P = a->P; p.F = a->F; p.sh_degree = a->sh_degree; p.M = a->M; p.W = a->W; p.H = a->H; p.tile = a->tile;
    p.gx = Lc.gx; p.gy = Lc.gy; p.rec_c = Lc.rec; p.rec_l = Ll.rec; p.chunk = Lc.chunk; p.flags = a->flags;
    p.tanfovx = a->tanfovx; p.tanfovy = a->tanfovy;
    p.focal_y = a->H / (2.0f * a->tanfovy);
    p.focal_x = a->W / (2.0f * a->tanfovx);
    p.scale_modifier = a->scale_modifier;
    p.means3D = a->d_means3D; p.shs = a->d_shs; p.colors_precomp = a->d_colors_precomp; p.language = a->d_language;
    p.opacities = a->d_opacities; p.scales = a->d_scales; p.rotations = a->d_rotations; p.cov3D_precomp = a->d_cov3D_precomp;
    p.opacities_lang = d->d_opacities_lang; p.scales_lang = d->d_scales_lang; p.rotations_lang = d->d_rotations_lang;
    p.cov3D_precomp_lang = d->d_cov3D_precomp_lang;
    p.viewmatrix = a->d_viewmatrix; p.projmatrix = a->d_projmatrix; p.campos = a->d_campos;
    p.records_c = (float*)(wc + Lc.records); p.records_l = (float*)(wl + Ll.records);
    p.depths = (float*)(wc + Lc.depths); p.cov3D = (float*)(wc + Lc.cov3D); p.cov3D_lang = (float*)(wl + Ll.cov3D);
    p.clamped = (uint32_t*)(wc + Lc.clamped);
    p.tiles_touched_c = (uint32_t*)(wc + Lc.tiles_touched); p.tiles_touched_l = (uint32_t*)(wl + Ll.tiles_touched);
    p.cta_hist_c = (uint32_t*)(wc + Lc.cta_hist); p.cta_hist_l = (uint32_t*)(wl + Ll.cta_hist);
    p.rect_c = (uint2*)(wc + Lc.rect); p.rect_l = (uint2*)(wl + Ll.rect);
    p.radii = o->d_radii; p.radii_lang = o->d_radii_lang;
    p.info_c = (DeviceInfo*)(wc + Lc.info); p.info_l = (DeviceInfo*)(wl + Ll.info);
    const size_t hist_smem = sizeof(uint32_t) * (size_t)Lc.n_tiles;
    int rc = set_hist_smem(2 * hist_smem, Lc.n_tiles);
    if (rc != OLS_OK) return rc;
    ols_timing_mark(-1, st);
    k_preprocess_dis<<<Lc.n_ctas, PRE_THREADS, 2 * hist_smem, st>>>(p);
    OLS_DEBUG_SYNC("preprocess_dis");
    ols_timing_mark(OLS_T_PREPROCESS, st);
    PassBatch pc, pl;
    pc.v[0] = PassView{wc, p.depths, a->d_bg, o->d_color, nullptr, o->d_depth, o->d_opacity, o->d_n_touched};
    rc = run_pass(a->P, a->W, a->H, a->tile, 3, 0, a->flags, a->R_cap, pc, 1, Lc, nullptr, st);
    if (rc != OLS_OK) return rc;
    ols_timing_mark(-1, st);
    pl.v[0] = PassView{wl, p.depths, a->d_bg, nullptr, o->d_language, nullptr, o->d_opacity_lang, o->d_n_touched_lang};
    return run_pass(a->P, a->W, a->H, a->tile, 0, a->F, a->flags, d->R_cap_lang, pl, 1, Ll, nullptr, st);
}
