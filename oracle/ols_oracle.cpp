// ols_oracle.cpp -- CPU restatement of the reference language rasterizer.
//
// *** TEST INFRASTRUCTURE ONLY ***  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load this library.  The
// product path (online_lang_splatting_b200/) never links, imports or calls it.
//
// Parity status: PINNED.  This restatement is checked (tests/test_oracle.py)
// against tests/golden/*.npz, which hold inputs and outputs of the *real*
// reference CUDA (submodules/diff-gaussian-rasterization compiled unmodified into
// oracle/_ref/ref_P_C.so by oracle/build_ref.py and executed on a B200 by
// tests/golden/make_golden.py).
//
// What is restated (reference = /root/reference/submodules/diff-gaussian-rasterization):
//   preprocess  : cuda_rasterizer/forward.cu:262-371  (+ auxiliary.h:41-56,139-164, forward.cu:23-155)
//   binning     : cuda_rasterizer/rasterizer_impl.cu:70-138,451-492 (scan, duplicateWithKeys, radix sort, ranges)
//   fwd blend   : cuda_rasterizer/forward.cu:377-513
//   bwd blend   : cuda_rasterizer/backward.cu:932-1201 (+ reduction helper :684-702)
//   bwd geometry: cuda_rasterizer/backward.cu:150-346 (cov2D), :350-413 (cov3D), :21-145 (SH), :541-682
//
// Floating point: the reference is compiled with -fmad=true, so the exact
// placement of fused multiply-adds decides the low bit of radii/rects/depths.
// The forward arithmetic below follows oracle/REF_ARITHMETIC.md, which is the
// dataflow extracted from the compiled reference's SASS (oracle/sass_dataflow.py).
// Build with -ffp-contract=off so that only the fmaf() calls written here fuse.
// div/sqrt/rcp are IEEE-correct on both sides.  expf() is NOT bit-identical to
// CUDA's (MUFU.EX2 based) expf, so alpha-threshold decisions may differ for a
// handful of (pixel, Gaussian) pairs; tests state that tolerance explicitly.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

struct OracleScene {
    int32_t P, sh_degree, M, F, W, H, tile, prefiltered;
    float tanfovx, tanfovy, scale_modifier, _pad;
    const float* bg;              // [3]
    const float* means3D;         // [P,3]
    const float* shs;             // [P,M,3] or null
    const float* colors_precomp;  // [P,3] or null
    const float* language;        // [P,F]
    const float* opacities;       // [P]
    const float* scales;          // [P,3] or null
    const float* rotations;       // [P,4] or null
    const float* cov3D_precomp;   // [P,6] or null
    const float* viewmatrix;      // [16]
    const float* projmatrix;      // [16]
    const float* projmatrix_raw;  // [16]
    const float* campos;          // [3]
};

struct OracleGeom {  // per-Gaussian state (mirrors LanguageGeometryState, rasterizer_impl.cu:173-188)
    float* depths;           // [P]
    int32_t* radii;          // [P]
    float* means2D;          // [P,2]
    float* cov3D;            // [P,6]
    float* conic_opacity;    // [P,4]
    float* rgb;              // [P,3]
    uint8_t* clamped;        // [P,3]
    uint32_t* tiles_touched; // [P]
    uint32_t* point_offsets; // [P] inclusive scan
};

struct OracleBin {  // BinningState + ImageState (rasterizer_impl.cu:190-212)
    int64_t R;
    uint64_t* keys_sorted;   // [R]
    uint32_t* point_list;    // [R]
    uint32_t* ranges;        // [tiles,2]
    float* final_T;          // [H*W]
    uint32_t* n_contrib;     // [H*W]
};

struct OracleImage {
    float* color;      // [3,H,W]
    float* language;   // [F,H,W]
    float* depth;      // [H,W]
    float* opacity;    // [H,W]
    int32_t* n_touched;// [P]
};

struct OracleGrads {
    const float* dL_dcolor;    // [3,H,W]
    const float* dL_dlanguage; // [F,H,W]
    const float* dL_ddepth;    // [H,W]
    int32_t compat;            // 1: reproduce reference quirks Q1-Q3 ; 0: mathematically exact
    int32_t _pad;
    float* dL_dmeans2D;   // [P,3]
    float* dL_dconic;     // [P,4]
    float* dL_dopacity;   // [P]
    float* dL_dcolors;    // [P,3]
    float* dL_dlang;      // [P,F]
    float* dL_ddepths;    // [P]
    float* dL_dmeans3D;   // [P,3]
    float* dL_dcov3D;     // [P,6]
    float* dL_dsh;        // [P,M,3]
    float* dL_dscales;    // [P,3]
    float* dL_drots;      // [P,4]
    float* dL_dtau;       // [P,6]
};

// ---- disentangled variant (D/ = submodules/diff-gaussian-rasterization-disentangle-optim) ----
struct OracleDisExtra {   // the language footprint (D/rasterize_points.h:51-80)
    const float* opacities_lang;      // [P]
    const float* scales_lang;         // [P,3] or null
    const float* rotations_lang;      // [P,4] or null
    const float* cov3D_precomp_lang;  // [P,6] or null
};

struct OracleDisGeom {    // the *_lang members of D/'s LanguageGeometryState (D/rasterizer_impl.cu:173-200)
    int32_t* radii_lang;           // [P]
    float* cov3D_lang;             // [P,6]
    float* conic_opacity_lang;     // [P,4]
    uint32_t* tiles_touched_lang;  // [P]
    uint32_t* point_offsets_lang;  // [P]
};

struct OracleDisGrads {
    const float* dL_dcolor;    // [3,H,W]
    const float* dL_dlanguage; // [F,H,W]
    const float* dL_ddepth;    // [H,W]
    int32_t compat;            // 1: reference behaviour incl. Q1/Q2 and the joint-visibility early return; 0: exact
    int32_t _pad;
    float* dL_dmeans2D;      // [P,3]
    float* dL_dconic;        // [P,4]
    float* dL_dconic_lang;   // [P,4]
    float* dL_dopacity;      // [P]
    float* dL_dopacity_lang; // [P]
    float* dL_dcolors;       // [P,3]
    float* dL_dlang;         // [P,F]
    float* dL_ddepths;       // [P]
    float* dL_dmeans3D;      // [P,3]
    float* dL_dcov3D;        // [P,6]
    float* dL_dcov3D_lang;   // [P,6]
    float* dL_dsh;           // [P,M,3]
    float* dL_dscales;       // [P,3]
    float* dL_dscales_lang;  // [P,3]
    float* dL_drots;         // [P,4]
    float* dL_drots_lang;    // [P,4]
    float* dL_dtau;          // [P,6]
};

}  // extern "C"

namespace {

const float SH_C0 = 0.28209479177387814f;
const float SH_C1 = 0.4886025119029199f;
const float SH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                       -1.0925484305920792f, 0.5462742152960396f};
const float SH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                       -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

// cvt.rzi.s32.f32 semantics (saturating, NaN -> 0)
inline int f2i_rz(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return std::numeric_limits<int>::max();
    if (v <= -2147483648.0f) return std::numeric_limits<int>::min();
    return (int)v;
}
inline float fmaxf_(float a, float b) { return std::fmax(a, b); }
inline float fminf_(float a, float b) { return std::fmin(a, b); }

// m[i]*x + m[4+i]*y + m[8+i]*z (+ m[12+i]) as compiled: fadd(ffma(z,m8,ffma(x,m0,fmul(y,m4))),m12)
inline float xform_row(const float* m, int i, float x, float y, float z) {
    return fmaf(z, m[8 + i], fmaf(x, m[i], y * m[4 + i])) + m[12 + i];
}
// a0*b0 + a1*b1 + a2*b2 as compiled: ffma(a2,b2,ffma(a0,b0,fmul(a1,b1)))
inline float dot3c(float a0, float a1, float a2, float b0, float b1, float b2) {
    return fmaf(a2, b2, fmaf(a0, b0, a1 * b1));
}

// forward.cu:121-155 (computeCov3D); arithmetic per REF_ARITHMETIC.md section 1.
void cov3d_from_scale_rot(const float* s, float mod, const float* q, float* out) {
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    const float sx = s[0] * mod, sy = s[1] * mod, sz = s[2] * mod;
    const float yy = y * y, zz = z * z;
    const float xx_yy = fmaf(x, x, yy);
    const float R22 = 1.0f - (xx_yy + xx_yy);
    const float xz = x * z;
    const float xz_m_ry = fmaf(-r, y, xz), xz_p_ry = fmaf(r, y, xz);
    const float rx = r * x;
    const float yz_p_rx = fmaf(y, z, rx), yz_m_rx = fmaf(y, z, -rx);
    const float rz = r * z;
    const float xy_m_rz = fmaf(x, y, -rz), xy_p_rz = fmaf(x, y, rz);
    const float yy_zz = yy + zz;
    const float R00 = 1.0f - (yy_zz + yy_zz);
    const float xx_zz = fmaf(x, x, zz);
    const float R11 = 1.0f - (xx_zz + xx_zz);
    const float A2 = xz_m_ry + xz_m_ry;  // 2(xz-ry)
    const float B2 = yz_p_rx + yz_p_rx;  // 2(yz+rx)
    const float C2 = xz_p_ry + xz_p_ry;  // 2(xz+ry)
    const float D2 = xy_m_rz + xy_m_rz;  // 2(xy-rz)
    const float E2 = yz_m_rx + yz_m_rx;  // 2(yz-rx)
    const float G2 = xy_p_rz + xy_p_rz;  // 2(xy+rz)
    // M = S*R evaluated by GLM with the zero entries of S kept (0*x terms survive compilation)
    const float zB = 0.0f * B2;
    const float m22 = fmaf(sz, R22, fmaf(0.0f, A2, zB));
    const float m02 = fmaf(0.0f, R22, fmaf(sx, A2, zB));
    const float m12 = fmaf(0.0f, R22, fmaf(0.0f, A2, sy * B2));
    const float z00 = 0.0f * R00;
    const float m20 = fmaf(sz, C2, fmaf(0.0f, D2, z00));
    const float m00 = fmaf(0.0f, C2, fmaf(0.0f, D2, sx * R00));
    const float m10 = fmaf(0.0f, C2, fmaf(sy, D2, z00));
    const float z11 = 0.0f * R11;
    const float m21 = fmaf(sz, E2, fmaf(0.0f, G2, z11));
    const float m01 = fmaf(0.0f, E2, fmaf(sx, G2, z11));
    const float m11 = fmaf(0.0f, E2, fmaf(0.0f, G2, sy * R11));
    out[0] = dot3c(m00, m10, m20, m00, m10, m20);
    out[1] = dot3c(m00, m10, m20, m01, m11, m21);
    out[2] = dot3c(m00, m10, m20, m02, m12, m22);
    out[3] = dot3c(m01, m11, m21, m01, m11, m21);
    out[4] = dot3c(m01, m11, m21, m02, m12, m22);
    out[5] = dot3c(m02, m12, m22, m02, m12, m22);
}

struct Cov2D { float a, b, c; };  // cov00+0.3, cov01, cov11+0.3

// forward.cu:77-116 (computeCov2D); arithmetic per REF_ARITHMETIC.md section 2.
Cov2D cov2d(const float* p, float fx, float fy, float tanx, float tany, const float* c, const float* V) {
    const float tz = xform_row(V, 2, p[0], p[1], p[2]);
    const float txr = xform_row(V, 0, p[0], p[1], p[2]);
    const float tyr = xform_row(V, 1, p[0], p[1], p[2]);
    const float limx = tanx * 1.3f, limy = tany * 1.3f;
    const float cx = fminf_(fmaxf_(txr / tz, -limx), limx);
    const float cy = fminf_(fmaxf_(tyr / tz, -limy), limy);
    const float tz2 = tz * tz;
    const float J00 = fx / tz;
    const float J02 = ((tz * -cx) * fx) / tz2;
    const float J11 = fy / tz;
    const float J12 = ((tz * -cy) * fy) / tz2;
    float a[3], b[3];
    for (int k = 0; k < 3; k++) {
        a[k] = fmaf(V[4 * k + 2], J02, fmaf(V[4 * k], J00, 0.0f * V[4 * k + 1]));
        b[k] = fmaf(V[4 * k + 2], J12, fmaf(0.0f, V[4 * k], V[4 * k + 1] * J11));
    }
    const float ux0 = dot3c(a[0], a[1], a[2], c[0], c[1], c[2]);
    const float ux1 = dot3c(a[0], a[1], a[2], c[1], c[3], c[4]);
    const float ux2 = dot3c(a[0], a[1], a[2], c[2], c[4], c[5]);
    const float uy0 = dot3c(b[0], b[1], b[2], c[0], c[1], c[2]);
    const float uy1 = dot3c(b[0], b[1], b[2], c[1], c[3], c[4]);
    const float uy2 = dot3c(b[0], b[1], b[2], c[2], c[4], c[5]);
    Cov2D r;
    r.a = dot3c(a[0], a[1], a[2], ux0, ux1, ux2) + 0.3f;
    r.b = dot3c(a[0], a[1], a[2], uy0, uy1, uy2);
    r.c = dot3c(b[0], b[1], b[2], uy0, uy1, uy2) + 0.3f;
    return r;
}

// auxiliary.h:41-44 (double arithmetic; compiled as dfma)
inline float ndc2pix(float v, int S) { return (float)(std::fma((double)v + 1.0, (double)S, -1.0) * 0.5); }

// auxiliary.h:46-56
inline void get_rect(float px, float py, int r, int tile, int gx, int gy, int* mn, int* mx) {
    const float rf = (float)r, tf = (float)tile;
    mn[0] = std::min(gx, std::max(0, f2i_rz((px - rf) / tf)));
    mn[1] = std::min(gy, std::max(0, f2i_rz((py - rf) / tf)));
    mx[0] = std::min(gx, std::max(0, f2i_rz((((px + rf) + tf) + -1.0f) / tf)));
    mx[1] = std::min(gy, std::max(0, f2i_rz((((py + rf) + tf) + -1.0f) / tf)));
}

// forward.cu:23-74
void sh_to_rgb(int idx, int deg, int M, const float* means, const float* campos, const float* shs, uint8_t* clamped,
               float* rgb) {
    const float* sh = shs + (size_t)idx * M * 3;
    float res[3];
    for (int c = 0; c < 3; c++) res[c] = SH_C0 * sh[c];
    if (deg > 0) {
        float dx = means[3 * idx] - campos[0], dy = means[3 * idx + 1] - campos[1], dz = means[3 * idx + 2] - campos[2];
        float len = std::sqrt(fmaf(dz, dz, fmaf(dx, dx, dy * dy)));
        float x = dx / len, y = dy / len, z = dz / len;
        for (int c = 0; c < 3; c++)
            res[c] = res[c] - SH_C1 * y * sh[3 + c] + SH_C1 * z * sh[6 + c] - SH_C1 * x * sh[9 + c];
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            for (int c = 0; c < 3; c++)
                res[c] = res[c] + SH_C2[0] * xy * sh[12 + c] + SH_C2[1] * yz * sh[15 + c] +
                         SH_C2[2] * (2.0f * zz - xx - yy) * sh[18 + c] + SH_C2[3] * xz * sh[21 + c] +
                         SH_C2[4] * (xx - yy) * sh[24 + c];
            if (deg > 2) {
                for (int c = 0; c < 3; c++)
                    res[c] = res[c] + SH_C3[0] * y * (3.0f * xx - yy) * sh[27 + c] + SH_C3[1] * xy * z * sh[30 + c] +
                             SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[33 + c] +
                             SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + c] +
                             SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[39 + c] + SH_C3[5] * z * (xx - yy) * sh[42 + c] +
                             SH_C3[6] * x * (xx - 3.0f * yy) * sh[45 + c];
            }
        }
    }
    for (int c = 0; c < 3; c++) {
        // compiled form: clamped = !(res >= -0.5) ; value = max(res + 0.5, 0)
        clamped[3 * idx + c] = !(res[c] >= -0.5f);
        rgb[3 * idx + c] = fmaxf_(res[c] + 0.5f, 0.0f);
    }
}

inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

// rasterizer_impl.cu:35-50
uint32_t higher_msb(uint32_t n) {
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

// Q3: which lanes of an n-thread block reach data[0] in render_cuda_reduce_sum (backward.cu:684-702)
std::vector<uint8_t> reduce_lane_mask(int n) {
    std::vector<std::vector<int>> sets(n);
    for (int i = 0; i < n; i++) sets[i] = {i};
    for (int i = n / 2; i > 0; i /= 2)
        for (int lane = 0; lane < i; lane++)
            sets[lane].insert(sets[lane].end(), sets[lane + i].begin(), sets[lane + i].end());
    std::vector<uint8_t> m(n, 0);
    for (int l : sets[0]) m[l] = 1;
    return m;
}

}  // namespace

extern "C" {

int ols_oracle_abi_version() { return 1; }

void ols_oracle_reduce_lane_mask(int n, uint8_t* out) {
    auto m = reduce_lane_mask(n);
    std::memcpy(out, m.data(), n);
}

// Phase 1: per-Gaussian preprocess + inclusive scan.  Returns R (number of Gaussian/tile instances).
int64_t ols_oracle_preprocess(const OracleScene* s, OracleGeom* g) {
    const int P = s->P, W = s->W, H = s->H, tile = s->tile;
    const int gx = (W + tile - 1) / tile, gy = (H + tile - 1) / tile;
    const float fy = H / (2.0f * s->tanfovy), fx = W / (2.0f * s->tanfovx);  // rasterizer_impl.cu:394-395
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        g->radii[i] = 0;
        g->tiles_touched[i] = 0;
        const float* p = s->means3D + 3 * (size_t)i;
        const float* V = s->viewmatrix;
        const float* Pm = s->projmatrix;
        const float vz = xform_row(V, 2, p[0], p[1], p[2]);  // in_frustum, auxiliary.h:139-164
        if (!(vz > 0.2f)) continue;
        const float hx = xform_row(Pm, 0, p[0], p[1], p[2]);
        const float hy = xform_row(Pm, 1, p[0], p[1], p[2]);
        const float hw = xform_row(Pm, 3, p[0], p[1], p[2]);
        const float pw = 1.0f / (hw + 0.0000001f);
        const float projx = hx * pw, projy = hy * pw;
        const float* c3;
        if (s->cov3D_precomp) {
            c3 = s->cov3D_precomp + 6 * (size_t)i;
        } else {
            cov3d_from_scale_rot(s->scales + 3 * (size_t)i, s->scale_modifier, s->rotations + 4 * (size_t)i,
                                 g->cov3D + 6 * (size_t)i);
            c3 = g->cov3D + 6 * (size_t)i;
        }
        const Cov2D cv = cov2d(p, fx, fy, s->tanfovx, s->tanfovy, c3, V);
        const float det = fmaf(cv.a, cv.c, -(cv.b * cv.b));
        if (det == 0.0f) continue;
        const float det_inv = 1.0f / det;
        const float mid = (cv.a + cv.c) * 0.5f;
        const float sq = std::sqrt(fmaxf_(fmaf(mid, mid, -det), 0.1f));
        const float lam = fmaxf_(mid + sq, mid - sq);
        const float my_radius = std::ceil(std::sqrt(lam) * 3.0f);
        const float px = ndc2pix(projx, W), py = ndc2pix(projy, H);
        int mn[2], mx[2];
        const int ri = f2i_rz(my_radius);
        get_rect(px, py, ri, tile, gx, gy, mn, mx);
        const uint32_t tiles = (uint32_t)(mx[0] - mn[0]) * (uint32_t)(mx[1] - mn[1]);
        if (tiles == 0) continue;
        if (!s->colors_precomp) sh_to_rgb(i, s->sh_degree, s->M, s->means3D, s->campos, s->shs, g->clamped, g->rgb);
        g->depths[i] = vz;
        g->radii[i] = ri;
        g->means2D[2 * (size_t)i] = px;
        g->means2D[2 * (size_t)i + 1] = py;
        float* co = g->conic_opacity + 4 * (size_t)i;
        co[0] = cv.c * det_inv;
        co[1] = cv.b * -det_inv;
        co[2] = cv.a * det_inv;
        co[3] = s->opacities[i];
        g->tiles_touched[i] = tiles;
    }
    uint64_t acc = 0;  // cub::DeviceScan::InclusiveSum over uint32 (rasterizer_impl.cu:451)
    for (int i = 0; i < P; i++) {
        acc += g->tiles_touched[i];
        g->point_offsets[i] = (uint32_t)acc;
    }
    return (int64_t)acc;
}

// Phase 2 of one blending pass: duplicateWithKeys + stable sort + identifyTileRanges + forward blend.
// The joint pass of P/ blends rgb + depth + F language channels over one list; D/ runs it twice
// (colour + depth with F = 0, then language only with has_color = false; D/forward.cu:437-655).
static int render_pass(const OracleScene* s, int F, bool has_color, const int32_t* radii, const float* conic_opacity,
                       const uint32_t* point_offsets, const float* means2D, const float* depths, const float* feat,
                       OracleBin* b, OracleImage* o) {
    const int P = s->P, W = s->W, H = s->H, tile = s->tile;
    const int gx = (W + tile - 1) / tile, gy = (H + tile - 1) / tile;
    const int64_t R = b->R;
    std::vector<uint64_t> keys((size_t)R);
    std::vector<uint32_t> vals((size_t)R);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < P; i++) {  // rasterizer_impl.cu:70-111
        if (radii[i] > 0) {
            uint32_t off = (i == 0) ? 0 : point_offsets[i - 1];
            int mn[2], mx[2];
            get_rect(means2D[2 * (size_t)i], means2D[2 * (size_t)i + 1], radii[i], tile, gx, gy, mn, mx);
            for (int y = mn[1]; y < mx[1]; y++)
                for (int x = mn[0]; x < mx[0]; x++) {
                    uint64_t key = (uint64_t)(uint32_t)(y * gx + x);
                    key <<= 32;
                    key |= f2u(depths[i]);
                    keys[off] = key;
                    vals[off] = (uint32_t)i;
                    off++;
                }
        }
    }
    // cub::DeviceRadixSort::SortPairs over bits [0, 32+bit): stable LSD radix == stable sort on masked key
    const uint32_t bit = higher_msb((uint32_t)(gx * gy));
    const uint64_t mask = (bit + 32 >= 64) ? ~0ull : ((1ull << (32 + bit)) - 1ull);
    std::vector<uint32_t> perm((size_t)R);
    for (int64_t i = 0; i < R; i++) perm[i] = (uint32_t)i;
    std::stable_sort(perm.begin(), perm.end(),
                     [&](uint32_t a, uint32_t c) { return (keys[a] & mask) < (keys[c] & mask); });
    for (int64_t i = 0; i < R; i++) {
        b->keys_sorted[i] = keys[perm[i]];
        b->point_list[i] = vals[perm[i]];
    }
    std::memset(b->ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy);  // rasterizer_impl.cu:485
    for (int64_t i = 0; i < R; i++) {                                   // rasterizer_impl.cu:116-138
        uint32_t cur = (uint32_t)(b->keys_sorted[i] >> 32);
        if (i == 0) b->ranges[2 * cur] = 0;
        else {
            uint32_t prev = (uint32_t)(b->keys_sorted[i - 1] >> 32);
            if (cur != prev) { b->ranges[2 * prev + 1] = (uint32_t)i; b->ranges[2 * cur] = (uint32_t)i; }
        }
        if (i == R - 1) b->ranges[2 * cur + 1] = (uint32_t)R;
    }
    const size_t HW = (size_t)H * W;
    std::memset(o->n_touched, 0, sizeof(int32_t) * (size_t)P);
    // forward.cu:377-513, one tile at a time; pixels within a tile are independent
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int ty = 0; ty < gy; ty++)
        for (int tx = 0; tx < gx; tx++) {
            const uint32_t r0 = b->ranges[2 * (ty * gx + tx)], r1 = b->ranges[2 * (ty * gx + tx) + 1];
            std::vector<float> L((size_t)std::max(F, 1));
            for (int ly = 0; ly < tile; ly++)
                for (int lx = 0; lx < tile; lx++) {
                    const int pxi = tx * tile + lx, pyi = ty * tile + ly;
                    if (pxi >= W || pyi >= H) continue;
                    const float pfx = (float)pxi, pfy = (float)pyi;
                    float T = 1.0f, C[3] = {0, 0, 0}, D = 0.0f;
                    std::fill(L.begin(), L.end(), 0.0f);
                    uint32_t contributor = 0, last = 0;
                    for (uint32_t k = r0; k < r1; k++) {
                        contributor++;
                        const uint32_t id = b->point_list[k];
                        const float* co = conic_opacity + 4 * (size_t)id;
                        const float dx = means2D[2 * (size_t)id] - pfx, dy = means2D[2 * (size_t)id + 1] - pfy;
                        const float q = fmaf(dx, dx * co[0], dy * (dy * co[2]));
                        const float power = fmaf(q, -0.5f, -(dy * (dx * co[1])));
                        if (power > 0.0f) continue;
                        const float alpha = fminf_(co[3] * expf(power), 0.99f);
                        if (alpha < 1.0f / 255.0f) continue;
                        const float test_T = T * (1.0f - alpha);
                        if (test_T < 0.0001f) break;  // done: entry not blended
                        if (has_color)
                            for (int ch = 0; ch < 3; ch++) C[ch] = fmaf(T, alpha * feat[3 * (size_t)id + ch], C[ch]);
                        const float* lf = s->language + (size_t)F * id;
                        for (int ch = 0; ch < F; ch++) L[ch] = fmaf(T, alpha * lf[ch], L[ch]);
                        if (has_color) D = fmaf(T, alpha * depths[id], D);
                        if (test_T > 0.5f) {
#pragma omp atomic
                            o->n_touched[id] += 1;
                        }
                        T = test_T;
                        last = contributor;
                    }
                    const size_t pix = (size_t)pyi * W + pxi;
                    b->final_T[pix] = T;
                    b->n_contrib[pix] = last;
                    if (has_color) {
                        for (int ch = 0; ch < 3; ch++) o->color[ch * HW + pix] = fmaf(s->bg[ch], T, C[ch]);
                        o->depth[pix] = D;
                    }
                    for (int ch = 0; ch < F; ch++) o->language[ch * HW + pix] = L[ch];
                    o->opacity[pix] = 1.0f - T;
                }
        }
    return 0;
}

int ols_oracle_render(const OracleScene* s, const OracleGeom* g, OracleBin* b, OracleImage* o) {
    const float* feat = s->colors_precomp ? s->colors_precomp : g->rgb;
    return render_pass(s, s->F, true, g->radii, g->conic_opacity, g->point_offsets, g->means2D, g->depths, feat, b, o);
}

// ---- per-Gaussian backward pieces (shared by the P/ and D/ entry points) ---------------------------------
// computeCov2DCUDA (backward.cu:150-346).  D/'s computeCov2DCUDA_no_tau (D/backward.cu:354-446) is the same
// arithmetic keeping only dL/dcov3D; callers pass scratch for dmean / dtau in that case.
static void cov2d_bwd_cpu(const OracleScene* s, float fx, float fy, const float* V, const float* mp, const float* c3,
                          float dcx, float dcy, float dcz, float* dcov, float* dmean, float* dtau) {
    // ---- computeCov2DCUDA (backward.cu:150-346)
    float t[3] = {V[0] * mp[0] + V[4] * mp[1] + V[8] * mp[2] + V[12], V[1] * mp[0] + V[5] * mp[1] + V[9] * mp[2] + V[13],
                  V[2] * mp[0] + V[6] * mp[1] + V[10] * mp[2] + V[14]};
    const float limx = 1.3f * s->tanfovx, limy = 1.3f * s->tanfovy;
    const float txtz = t[0] / t[2], tytz = t[1] / t[2];
    t[0] = fminf_(limx, fmaxf_(-limx, txtz)) * t[2];
    t[1] = fminf_(limy, fmaxf_(-limy, tytz)) * t[2];
    const float xgm = (txtz < -limx || txtz > limx) ? 0.0f : 1.0f;
    const float ygm = (tytz < -limy || tytz > limy) ? 0.0f : 1.0f;
    // GLM column-major: J[c][r]
    float J[3][3] = {{fx / t[2], 0, -(fx * t[0]) / (t[2] * t[2])}, {0, fy / t[2], -(fy * t[1]) / (t[2] * t[2])}, {0, 0, 0}};
    float Wm[3][3] = {{V[0], V[4], V[8]}, {V[1], V[5], V[9]}, {V[2], V[6], V[10]}};
    float Vrk[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    float Tm[3][3];  // T = W * J : T[c][r] = sum_k W[k][r] * J[c][k]
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) Tm[c][r] = Wm[0][r] * J[c][0] + Wm[1][r] * J[c][1] + Wm[2][r] * J[c][2];
    // cov2D = T^T * Vrk^T * T ; only [0][0],[0][1],[1][1] needed
    auto quad = [&](int i0, int i1) {
        float r = 0;
        for (int p_ = 0; p_ < 3; p_++)
            for (int q_ = 0; q_ < 3; q_++) r += Tm[i0][p_] * Vrk[p_][q_] * Tm[i1][q_];
        return r;
    };
    const float a = quad(0, 0) + 0.3f, bb = quad(0, 1), c = quad(1, 1) + 0.3f;
    const float denom = a * c - bb * bb;
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    if (denom2inv != 0) {
        dL_da = denom2inv * (-c * c * dcx + 2 * bb * c * dcy + (denom - a * c) * dcz);
        dL_dc = denom2inv * (-a * a * dcz + 2 * a * bb * dcy + (denom - a * c) * dcx);
        dL_db = denom2inv * 2 * (bb * c * dcx - (denom + 2 * bb * bb) * dcy + a * bb * dcz);
        dcov[0] = (Tm[0][0] * Tm[0][0] * dL_da + Tm[0][0] * Tm[1][0] * dL_db + Tm[1][0] * Tm[1][0] * dL_dc);
        dcov[3] = (Tm[0][1] * Tm[0][1] * dL_da + Tm[0][1] * Tm[1][1] * dL_db + Tm[1][1] * Tm[1][1] * dL_dc);
        dcov[5] = (Tm[0][2] * Tm[0][2] * dL_da + Tm[0][2] * Tm[1][2] * dL_db + Tm[1][2] * Tm[1][2] * dL_dc);
        dcov[1] = 2 * Tm[0][0] * Tm[0][1] * dL_da + (Tm[0][0] * Tm[1][1] + Tm[0][1] * Tm[1][0]) * dL_db + 2 * Tm[1][0] * Tm[1][1] * dL_dc;
        dcov[2] = 2 * Tm[0][0] * Tm[0][2] * dL_da + (Tm[0][0] * Tm[1][2] + Tm[0][2] * Tm[1][0]) * dL_db + 2 * Tm[1][0] * Tm[1][2] * dL_dc;
        dcov[4] = 2 * Tm[0][2] * Tm[0][1] * dL_da + (Tm[0][1] * Tm[1][2] + Tm[0][2] * Tm[1][1]) * dL_db + 2 * Tm[1][1] * Tm[1][2] * dL_dc;
    }
    auto tv = [&](int r_, int k) { return Tm[r_][0] * Vrk[k][0] + Tm[r_][1] * Vrk[k][1] + Tm[r_][2] * Vrk[k][2]; };
    const float dT00 = 2 * tv(0, 0) * dL_da + tv(1, 0) * dL_db, dT01 = 2 * tv(0, 1) * dL_da + tv(1, 1) * dL_db,
                dT02 = 2 * tv(0, 2) * dL_da + tv(1, 2) * dL_db;
    const float dT10 = 2 * tv(1, 0) * dL_dc + tv(0, 0) * dL_db, dT11 = 2 * tv(1, 1) * dL_dc + tv(0, 1) * dL_db,
                dT12 = 2 * tv(1, 2) * dL_dc + tv(0, 2) * dL_db;
    const float dJ00 = Wm[0][0] * dT00 + Wm[0][1] * dT01 + Wm[0][2] * dT02;
    const float dJ02 = Wm[2][0] * dT00 + Wm[2][1] * dT01 + Wm[2][2] * dT02;
    const float dJ11 = Wm[1][0] * dT10 + Wm[1][1] * dT11 + Wm[1][2] * dT12;
    const float dJ12 = Wm[2][0] * dT10 + Wm[2][1] * dT11 + Wm[2][2] * dT12;
    const float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    const float dtx = xgm * -fx * tz2 * dJ02;
    const float dty = ygm * -fy * tz2 * dJ12;
    const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * t[0]) * tz3 * dJ02 + (2 * fy * t[1]) * tz3 * dJ12;
    // pose part: dpC_drho = I ; dpC_dtheta = -skew(t) with columns (0,-tz,ty),(tz,0,-tx),(-ty,tx,0)
    {
        const float th[3][3] = {{0, -t[2], t[1]}, {t[2], 0, -t[0]}, {-t[1], t[0], 0}};
        const float d3[3] = {dtx, dty, dtz};
        for (int k = 0; k < 3; k++) {
            dtau[k] += d3[k];
            dtau[k + 3] += dtx * th[k][0] + dty * th[k][1] + dtz * th[k][2];
        }
    }
    dmean[0] = V[0] * dtx + V[1] * dty + V[2] * dtz;
    dmean[1] = V[4] * dtx + V[5] * dty + V[6] * dtz;
    dmean[2] = V[8] * dtx + V[9] * dty + V[10] * dtz;
    {
        const float dW00 = J[0][0] * dT00, dW01 = J[0][0] * dT01, dW02 = J[0][0] * dT02;
        const float dW10 = J[1][1] * dT10, dW11 = J[1][1] * dT11, dW12 = J[1][1] * dT12;
        const float dW20 = J[0][2] * dT00 + J[1][2] * dT10, dW21 = J[0][2] * dT01 + J[1][2] * dT11,
                    dW22 = J[0][2] * dT02 + J[1][2] * dT12;
        // R columns (W2C rotation): c_k = (V[4k], V[4k+1], V[4k+2]); dL_dW columns likewise
        const float c1[3] = {V[0], V[1], V[2]}, c2[3] = {V[4], V[5], V[6]}, c3_[3] = {V[8], V[9], V[10]};
        const float w1[3] = {dW00, dW10, dW20}, w2[3] = {dW01, dW11, dW21}, w3[3] = {dW02, dW12, dW22};
        auto nskew_col = [](const float* v, int k, float* o3) {  // column k of -skew(v)
            const float S[3][3] = {{0, -v[2], v[1]}, {v[2], 0, -v[0]}, {-v[1], v[0], 0}};
            o3[0] = S[k][0]; o3[1] = S[k][1]; o3[2] = S[k][2];
        };
        for (int k = 0; k < 3; k++) {
            float n1[3], n2[3], n3[3];
            nskew_col(c1, k, n1); nskew_col(c2, k, n2); nskew_col(c3_, k, n3);
            dtau[3 + k] += (w1[0] * n1[0] + w1[1] * n1[1] + w1[2] * n1[2]) + (w2[0] * n2[0] + w2[1] * n2[1] + w2[2] * n2[2]) +
                           (w3[0] * n3[0] + w3[1] * n3[1] + w3[2] * n3[2]);
        }
    }
}

// language_preprocessCUDA (backward.cu:541-682): projection + depth paths
static void proj_bwd_cpu(const float* V, const float* Pm, const float* Praw, const float* mp, float g2x, float g2y, float dzv,
                         float* dmean, float* dtau) {
    // ---- language_preprocessCUDA (backward.cu:541-682)
    const float hxw = Pm[0] * mp[0] + Pm[4] * mp[1] + Pm[8] * mp[2] + Pm[12];
    const float hyw = Pm[1] * mp[0] + Pm[5] * mp[1] + Pm[9] * mp[2] + Pm[13];
    const float hww = Pm[3] * mp[0] + Pm[7] * mp[1] + Pm[11] * mp[2] + Pm[15];
    const float m_w = 1.0f / (hww + 0.0000001f);
    const float mul1 = hxw * m_w * m_w, mul2 = hyw * m_w * m_w;
    dmean[0] += (Pm[0] * m_w - Pm[3] * mul1) * g2x + (Pm[1] * m_w - Pm[3] * mul2) * g2y;
    dmean[1] += (Pm[4] * m_w - Pm[7] * mul1) * g2x + (Pm[5] * m_w - Pm[7] * mul2) * g2y;
    dmean[2] += (Pm[8] * m_w - Pm[11] * mul1) * g2x + (Pm[9] * m_w - Pm[11] * mul2) * g2y;
    {
        const float alpha = 1.0f * m_w, beta = -hxw * m_w * m_w, gamma = -hyw * m_w * m_w;
        const float pa = Praw[0], pb = Praw[5], pe = Praw[11];
        const float pC[3] = {V[0] * mp[0] + V[4] * mp[1] + V[8] * mp[2] + V[12], V[1] * mp[0] + V[5] * mp[1] + V[9] * mp[2] + V[13],
                             V[2] * mp[0] + V[6] * mp[1] + V[10] * mp[2] + V[14]};
        const float d1[3] = {alpha * pa, 0.f, beta * pe}, d2[3] = {0.f, alpha * pb, gamma * pe};
        // dp_C_d_theta = -skew(p_C); (A^T x)_k = dot(column k of A, x)
        const float th[3][3] = {{0, -pC[2], pC[1]}, {pC[2], 0, -pC[0]}, {-pC[1], pC[0], 0}};
        for (int k = 0; k < 3; k++) {
            dtau[k] += g2x * d1[k] + g2y * d2[k];
            const float t1 = th[k][0] * d1[0] + th[k][1] * d1[1] + th[k][2] * d1[2];
            const float t2 = th[k][0] * d2[0] + th[k][1] * d2[1] + th[k][2] * d2[2];
            dtau[3 + k] += g2x * t1 + g2y * t2;
        }
        const float dz = dzv;
        dmean[0] += dz * V[2]; dmean[1] += dz * V[6]; dmean[2] += dz * V[10];
        for (int k = 0; k < 3; k++) {
            dtau[k] += dz * (k == 2 ? 1.0f : 0.0f);
            dtau[3 + k] += dz * th[k][2];
        }
    }
}

static void sh_bwd_cpu(const OracleScene* s, int i, const float* mp, const uint8_t* clamped, const float* dL_dcolors,
                       float* dL_dsh, float* dmean, float* dtau) {
    const int M = s->M;
    if (s->shs) {  // backward.cu:21-145
        const float* sh = s->shs + (size_t)i * M * 3;
        float* dsh = dL_dsh + (size_t)i * M * 3;
        const int deg = s->sh_degree;
        float dir0[3] = {mp[0] - s->campos[0], mp[1] - s->campos[1], mp[2] - s->campos[2]};
        const float len = std::sqrt(dir0[0] * dir0[0] + dir0[1] * dir0[1] + dir0[2] * dir0[2]);
        const float x = dir0[0] / len, y = dir0[1] / len, z = dir0[2] / len;
        float dRGB[3];
        for (int c_ = 0; c_ < 3; c_++) dRGB[c_] = dL_dcolors[3 * (size_t)i + c_] * (clamped[3 * (size_t)i + c_] ? 0.f : 1.f);
        float dx_[3] = {0, 0, 0}, dy_[3] = {0, 0, 0}, dz_[3] = {0, 0, 0};
        for (int c_ = 0; c_ < 3; c_++) dsh[c_] = SH_C0 * dRGB[c_];
        if (deg > 0) {
            for (int c_ = 0; c_ < 3; c_++) {
                dsh[3 + c_] = -SH_C1 * y * dRGB[c_]; dsh[6 + c_] = SH_C1 * z * dRGB[c_]; dsh[9 + c_] = -SH_C1 * x * dRGB[c_];
                dx_[c_] = -SH_C1 * sh[9 + c_]; dy_[c_] = -SH_C1 * sh[3 + c_]; dz_[c_] = SH_C1 * sh[6 + c_];
            }
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                for (int c_ = 0; c_ < 3; c_++) {
                    dsh[12 + c_] = SH_C2[0] * xy * dRGB[c_]; dsh[15 + c_] = SH_C2[1] * yz * dRGB[c_];
                    dsh[18 + c_] = SH_C2[2] * (2.f * zz - xx - yy) * dRGB[c_]; dsh[21 + c_] = SH_C2[3] * xz * dRGB[c_];
                    dsh[24 + c_] = SH_C2[4] * (xx - yy) * dRGB[c_];
                    dx_[c_] += SH_C2[0] * y * sh[12 + c_] + SH_C2[2] * 2.f * -x * sh[18 + c_] + SH_C2[3] * z * sh[21 + c_] + SH_C2[4] * 2.f * x * sh[24 + c_];
                    dy_[c_] += SH_C2[0] * x * sh[12 + c_] + SH_C2[1] * z * sh[15 + c_] + SH_C2[2] * 2.f * -y * sh[18 + c_] + SH_C2[4] * 2.f * -y * sh[24 + c_];
                    dz_[c_] += SH_C2[1] * y * sh[15 + c_] + SH_C2[2] * 2.f * 2.f * z * sh[18 + c_] + SH_C2[3] * x * sh[21 + c_];
                }
                if (deg > 2) {
                    for (int c_ = 0; c_ < 3; c_++) {
                        dsh[27 + c_] = SH_C3[0] * y * (3.f * xx - yy) * dRGB[c_]; dsh[30 + c_] = SH_C3[1] * xy * z * dRGB[c_];
                        dsh[33 + c_] = SH_C3[2] * y * (4.f * zz - xx - yy) * dRGB[c_];
                        dsh[36 + c_] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy) * dRGB[c_];
                        dsh[39 + c_] = SH_C3[4] * x * (4.f * zz - xx - yy) * dRGB[c_]; dsh[42 + c_] = SH_C3[5] * z * (xx - yy) * dRGB[c_];
                        dsh[45 + c_] = SH_C3[6] * x * (xx - 3.f * yy) * dRGB[c_];
                        dx_[c_] += (SH_C3[0] * sh[27 + c_] * 3.f * 2.f * xy + SH_C3[1] * sh[30 + c_] * yz + SH_C3[2] * sh[33 + c_] * -2.f * xy +
                                    SH_C3[3] * sh[36 + c_] * -3.f * 2.f * xz + SH_C3[4] * sh[39 + c_] * (-3.f * xx + 4.f * zz - yy) +
                                    SH_C3[5] * sh[42 + c_] * 2.f * xz + SH_C3[6] * sh[45 + c_] * 3.f * (xx - yy));
                        dy_[c_] += (SH_C3[0] * sh[27 + c_] * 3.f * (xx - yy) + SH_C3[1] * sh[30 + c_] * xz + SH_C3[2] * sh[33 + c_] * (-3.f * yy + 4.f * zz - xx) +
                                    SH_C3[3] * sh[36 + c_] * -3.f * 2.f * yz + SH_C3[4] * sh[39 + c_] * -2.f * xy + SH_C3[5] * sh[42 + c_] * -2.f * yz +
                                    SH_C3[6] * sh[45 + c_] * -3.f * 2.f * xy);
                        dz_[c_] += (SH_C3[1] * sh[30 + c_] * xy + SH_C3[2] * sh[33 + c_] * 4.f * 2.f * yz + SH_C3[3] * sh[36 + c_] * 3.f * (2.f * zz - xx - yy) +
                                    SH_C3[4] * sh[39 + c_] * 4.f * 2.f * xz + SH_C3[5] * sh[42 + c_] * (xx - yy));
                    }
                }
            }
        }
        const float ddir[3] = {dx_[0] * dRGB[0] + dx_[1] * dRGB[1] + dx_[2] * dRGB[2], dy_[0] * dRGB[0] + dy_[1] * dRGB[1] + dy_[2] * dRGB[2],
                               dz_[0] * dRGB[0] + dz_[1] * dRGB[1] + dz_[2] * dRGB[2]};
        const float sum2 = dir0[0] * dir0[0] + dir0[1] * dir0[1] + dir0[2] * dir0[2];
        const float inv32 = 1.0f / std::sqrt(sum2 * sum2 * sum2);
        const float dm[3] = {((+sum2 - dir0[0] * dir0[0]) * ddir[0] - dir0[1] * dir0[0] * ddir[1] - dir0[2] * dir0[0] * ddir[2]) * inv32,
                             (-dir0[0] * dir0[1] * ddir[0] + (sum2 - dir0[1] * dir0[1]) * ddir[1] - dir0[2] * dir0[1] * ddir[2]) * inv32,
                             (-dir0[0] * dir0[2] * ddir[0] - dir0[1] * dir0[2] * ddir[1] + (sum2 - dir0[2] * dir0[2]) * ddir[2]) * inv32};
        for (int k = 0; k < 3; k++) { dmean[k] += dm[k]; dtau[k] += -dm[k]; }
    }
}

static void cov3d_bwd_cpu(const OracleScene* s, const float* scales, const float* rotations, int i, const float* dcov,
                          float* dL_dscales, float* dL_drots) {
    if (scales) {  // backward.cu:350-413
        const float* q = rotations + 4 * (size_t)i; const float* sc = scales + 3 * (size_t)i;
        const float r = q[0], x = q[1], y = q[2], z = q[3];
        // GLM column-major R[c][r] built from the 9 literals (columns)
        const float Rm[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                                {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                                {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
        const float sv[3] = {s->scale_modifier * sc[0], s->scale_modifier * sc[1], s->scale_modifier * sc[2]};
        float Mm[3][3];  // M = S * R : M[c][r] = s_r * R[c][r]
        for (int c_ = 0; c_ < 3; c_++) for (int r_ = 0; r_ < 3; r_++) Mm[c_][r_] = sv[r_] * Rm[c_][r_];
        const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]}, {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]}, {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
        float dM[3][3];  // dL_dM = 2 * M * dL_dSigma : [c][r] = 2 * sum_k M[k][r] * dS[c][k]
        for (int c_ = 0; c_ < 3; c_++) for (int r_ = 0; r_ < 3; r_++) dM[c_][r_] = 2.0f * (Mm[0][r_] * dS[c_][0] + Mm[1][r_] * dS[c_][1] + Mm[2][r_] * dS[c_][2]);
        float Rt[3][3], dMt[3][3];
        for (int c_ = 0; c_ < 3; c_++) for (int r_ = 0; r_ < 3; r_++) { Rt[c_][r_] = Rm[r_][c_]; dMt[c_][r_] = dM[r_][c_]; }
        float* dsc = dL_dscales + 3 * (size_t)i;
        for (int k = 0; k < 3; k++) dsc[k] = Rt[k][0] * dMt[k][0] + Rt[k][1] * dMt[k][1] + Rt[k][2] * dMt[k][2];
        for (int k = 0; k < 3; k++) for (int r_ = 0; r_ < 3; r_++) dMt[k][r_] *= sv[k];
        float* dq = dL_drots + 4 * (size_t)i;
        dq[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
        dq[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
        dq[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
        dq[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
    }
}

// Backward blend of one pass (backward.cu:932-1201; D/backward.cu:1052-1427 runs it once per footprint:
// colour + depth with F = 0, then language only with has_color = false, where the mean2D gradient and the
// background term do not exist).  Per-pixel arithmetic in fp32 as the reference; per-Gaussian sums in fp64
// (order independent referee).  acc is [P, 10 + F]: mean2D.xy, conic.xyw, opacity, color.xyz, depth, lang[F].
static void blend_bwd_pass(const OracleScene* s, int F, bool has_color, const float* conic_opacity, const float* means2D,
                           const float* depths, const float* feat, const OracleBin* b, const float* dL_dcolor,
                           const float* dL_dlanguage, const float* dL_ddepth, bool compat, std::vector<double>& acc) {
    const int W = s->W, H = s->H, tile = s->tile;
    const int gx = (W + tile - 1) / tile, gy = (H + tile - 1) / tile;
    const int BS = tile * tile;
    const size_t HW = (size_t)H * W;
    const int NV = 10 + F;
    std::vector<uint8_t> lane_ok = compat ? reduce_lane_mask(BS) : std::vector<uint8_t>(BS, 1);
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int ty = 0; ty < gy; ty++)
        for (int tx = 0; tx < gx; tx++) {
            const uint32_t r0 = b->ranges[2 * (ty * gx + tx)], r1 = b->ranges[2 * (ty * gx + tx) + 1];
            if (r1 <= r0) continue;
            struct Px {
                bool inside; float T, T_final, last_alpha, accd, lastd, gd; uint32_t contributor, last_contributor;
                float acc[3], lastc[3], gc[3];
            };
            std::vector<Px> px(BS);
            std::vector<float> accF((size_t)BS * F + 1, 0.0f), lastF((size_t)BS * F + 1, 0.0f), gF((size_t)BS * F + 1, 0.0f);
            for (int t = 0; t < BS; t++) {
                Px& q = px[t];
                const int pxi = tx * tile + t % tile, pyi = ty * tile + t / tile;
                q.inside = pxi < W && pyi < H;
                const size_t pix = q.inside ? (size_t)pyi * W + pxi : 0;
                q.T_final = q.inside ? b->final_T[pix] : 0.0f;
                q.T = q.T_final;
                q.contributor = r1 - r0;
                q.last_contributor = q.inside ? b->n_contrib[pix] : 0;
                q.last_alpha = 0; q.accd = 0; q.lastd = 0;
                q.gd = (q.inside && has_color) ? dL_ddepth[pix] : 0.0f;
                for (int c = 0; c < 3; c++) { q.acc[c] = 0; q.lastc[c] = 0; q.gc[c] = (q.inside && has_color) ? dL_dcolor[c * HW + pix] : 0.0f; }
                for (int c = 0; c < F; c++) gF[(size_t)t * F + c] = q.inside ? dL_dlanguage[c * HW + pix] : 0.0f;
            }
            std::vector<uint8_t> skipv(BS);
            std::vector<float> alphav(BS), Gv(BS), dxv(BS), dyv(BS);
            std::vector<double> sum(NV);
            for (uint32_t k = r1; k-- > r0;) {
                const uint32_t id = b->point_list[k];
                const float* co = conic_opacity + 4 * (size_t)id;
                int nskip = 0;
                for (int t = 0; t < BS; t++) {
                    Px& q = px[t];
                    const bool done = !q.inside;
                    bool skip = done;
                    q.contributor = done ? q.contributor : q.contributor - 1;
                    skip |= q.contributor >= q.last_contributor;
                    const float pfx = (float)(tx * tile + t % tile), pfy = (float)(ty * tile + t / tile);
                    const float dx = means2D[2 * (size_t)id] - pfx, dy = means2D[2 * (size_t)id + 1] - pfy;
                    const float qd = fmaf(dx, dx * co[0], dy * (dy * co[2]));
                    const float power = fmaf(qd, -0.5f, -(dy * (dx * co[1])));
                    skip |= power > 0.0f;
                    const float G = expf(power);
                    const float alpha = fminf_(0.99f, co[3] * G);
                    skip |= alpha < 1.0f / 255.0f;
                    skipv[t] = skip; alphav[t] = alpha; Gv[t] = G; dxv[t] = dx; dyv[t] = dy;
                    nskip += skip;
                }
                if (nskip == BS) continue;  // backward.cu:1091-1093 (whole block skips)
                std::fill(sum.begin(), sum.end(), 0.0);
                const float* lf = s->language + (size_t)F * id;
                const float depth = depths[id];
                for (int t = 0; t < BS; t++) {
                    Px& q = px[t];
                    const bool skip = skipv[t];
                    if (!compat && skip) continue;  // exact mode: skipped pixels have no side effects
                    const float alpha = alphav[t], G = Gv[t], dx = dxv[t], dy = dyv[t];
                    q.T = skip ? q.T : q.T / (1.0f - alpha);
                    const float dch = alpha * q.T;
                    float dL_dalpha = 0.0f;
                    float lc[3];
                    for (int c = 0; c < (has_color ? 3 : 0); c++) {
                        const float col = feat[3 * (size_t)id + c];
                        q.acc[c] = skip ? q.acc[c] : q.last_alpha * q.lastc[c] + (1.0f - q.last_alpha) * q.acc[c];
                        q.lastc[c] = skip ? q.lastc[c] : col;
                        dL_dalpha += (col - q.acc[c]) * q.gc[c];
                        lc[c] = skip ? 0.0f : dch * q.gc[c];
                    }
                    float ld = 0.0f;
                    if (has_color) {
                        q.accd = skip ? q.accd : q.last_alpha * q.lastd + (1.0f - q.last_alpha) * q.accd;
                        q.lastd = skip ? q.lastd : depth;
                        dL_dalpha += (depth - q.accd) * q.gd;
                        ld = skip ? 0.0f : dch * q.gd;
                    } else {
                        lc[0] = lc[1] = lc[2] = 0.0f;
                    }
                    float* aF = &accF[(size_t)t * F]; float* lF = &lastF[(size_t)t * F]; const float* gFp = &gF[(size_t)t * F];
                    const bool lang_lane = compat ? (t == 0) : true;  // Q1
                    for (int c = 0; c < F; c++) {
                        // Q2: compat updates the recurrence even when skip (backward.cu:1132-1133)
                        aF[c] = q.last_alpha * lF[c] + (1.0f - q.last_alpha) * aF[c];
                        lF[c] = lf[c];
                        dL_dalpha += (lf[c] - aF[c]) * gFp[c];
                        const float v = skip ? 0.0f : dch * gFp[c];
                        if (lang_lane) sum[10 + c] += v;
                    }
                    dL_dalpha *= q.T;
                    q.last_alpha = skip ? q.last_alpha : alpha;
                    if (has_color) {
                        float bgdot = 0.0f;
                        for (int c = 0; c < 3; c++) bgdot += s->bg[c] * q.gc[c];
                        dL_dalpha += (-q.T_final / (1.0f - alpha)) * bgdot;
                    }
                    const float dL_dG = co[3] * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * co[0] - gdy * co[1];
                    const float dG_ddely = -gdy * co[2] - gdx * co[1];
                    if (skip || !lane_ok[t]) continue;  // Q3 lane mask (all ones unless compat && 225 threads)
                    if (has_color) {  // D/ drops dL_dmean2D of the language footprint (D/backward.cu:1074,1117)
                        sum[0] += dL_dG * dG_ddelx * ddelx_dx;
                        sum[1] += dL_dG * dG_ddely * ddely_dy;
                    }
                    sum[2] += -0.5f * gdx * dx * dL_dG;
                    sum[3] += -0.5f * gdx * dy * dL_dG;
                    sum[4] += -0.5f * gdy * dy * dL_dG;
                    sum[5] += G * dL_dalpha;
                    sum[6] += lc[0]; sum[7] += lc[1]; sum[8] += lc[2];
                    sum[9] += ld;
                }
                double* dst = &acc[(size_t)id * NV];
                for (int v = 0; v < NV; v++) {
                    if (sum[v] != 0.0) {
#pragma omp atomic
                        dst[v] += sum[v];
                    }
                }
            }
        }
}

// Backward: blend then per-Gaussian (backward.cu:150-346, 541-682).
int ols_oracle_backward(const OracleScene* s, const OracleGeom* g, const OracleBin* b, OracleGrads* gr) {
    const int P = s->P, W = s->W, H = s->H, F = s->F, M = s->M;
    const bool compat = gr->compat != 0;
    const float* feat = s->colors_precomp ? s->colors_precomp : g->rgb;
    const int NV = 10 + F;
    std::vector<double> acc((size_t)P * NV, 0.0);
    blend_bwd_pass(s, F, true, g->conic_opacity, g->means2D, g->depths, feat, b, gr->dL_dcolor, gr->dL_dlanguage,
                   gr->dL_ddepth, compat, acc);
    // scatter to the reference's gradient tensors
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        const double* a = &acc[(size_t)i * NV];
        gr->dL_dmeans2D[3 * (size_t)i] = (float)a[0];
        gr->dL_dmeans2D[3 * (size_t)i + 1] = (float)a[1];
        gr->dL_dmeans2D[3 * (size_t)i + 2] = 0.0f;
        gr->dL_dconic[4 * (size_t)i] = (float)a[2];
        gr->dL_dconic[4 * (size_t)i + 1] = (float)a[3];
        gr->dL_dconic[4 * (size_t)i + 2] = 0.0f;
        gr->dL_dconic[4 * (size_t)i + 3] = (float)a[4];
        gr->dL_dopacity[i] = (float)a[5];
        for (int c = 0; c < 3; c++) gr->dL_dcolors[3 * (size_t)i + c] = (float)a[6 + c];
        gr->dL_ddepths[i] = (float)a[9];
        for (int c = 0; c < F; c++) gr->dL_dlang[(size_t)F * i + c] = (float)a[10 + c];
    }

    const float fy = H / (2.0f * s->tanfovy), fx = W / (2.0f * s->tanfovx);
    const float* V = s->viewmatrix; const float* Pm = s->projmatrix; const float* Praw = s->projmatrix_raw;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        float* dmean = gr->dL_dmeans3D + 3 * (size_t)i;
        float* dcov = gr->dL_dcov3D + 6 * (size_t)i;
        float* dtau = gr->dL_dtau + 6 * (size_t)i;
        for (int k = 0; k < 3; k++) dmean[k] = 0;
        for (int k = 0; k < 6; k++) { dcov[k] = 0; dtau[k] = 0; }
        if (gr->dL_dscales) for (int k = 0; k < 3; k++) gr->dL_dscales[3 * (size_t)i + k] = 0;
        if (gr->dL_drots) for (int k = 0; k < 4; k++) gr->dL_drots[4 * (size_t)i + k] = 0;
        if (gr->dL_dsh) for (int k = 0; k < 3 * M; k++) gr->dL_dsh[(size_t)3 * M * i + k] = 0;
        if (!(g->radii[i] > 0)) continue;
        const float* c3 = s->cov3D_precomp ? s->cov3D_precomp + 6 * (size_t)i : g->cov3D + 6 * (size_t)i;
        const float* mp = s->means3D + 3 * (size_t)i;
        cov2d_bwd_cpu(s, fx, fy, V, mp, c3, gr->dL_dconic[4 * (size_t)i], gr->dL_dconic[4 * (size_t)i + 1],
                      gr->dL_dconic[4 * (size_t)i + 3], dcov, dmean, dtau);
        proj_bwd_cpu(V, Pm, Praw, mp, gr->dL_dmeans2D[3 * (size_t)i], gr->dL_dmeans2D[3 * (size_t)i + 1], gr->dL_ddepths[i],
                     dmean, dtau);
        sh_bwd_cpu(s, i, mp, g->clamped, gr->dL_dcolors, gr->dL_dsh, dmean, dtau);
        cov3d_bwd_cpu(s, s->scales, s->rotations, i, dcov, gr->dL_dscales, gr->dL_drots);
    }
    return 0;
}

// =====================================================================================================
// Disentangled variant.  Restates D/cuda_rasterizer/forward.cu:262-430 (preprocess), :437-655 (two-pass
// render), D/cuda_rasterizer/rasterizer_impl.cu:494-567 (doubled scan / duplicate / sort / ranges),
// D/cuda_rasterizer/backward.cu:1052-1427 (two-loop backward render), :354-446 (no_tau), :641-803 and
// :1504-1618 (per-Gaussian backward and its launch order).
// =====================================================================================================
int64_t ols_oracle_dis_preprocess(const OracleScene* s, const OracleDisExtra* x, OracleGeom* g, OracleDisGeom* gl,
                                  int64_t* R_lang) {
    const int P = s->P, W = s->W, H = s->H, tile = s->tile;
    const int gx = (W + tile - 1) / tile, gy = (H + tile - 1) / tile;
    const float fy = H / (2.0f * s->tanfovy), fx = W / (2.0f * s->tanfovx);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        g->radii[i] = 0; gl->radii_lang[i] = 0;
        g->tiles_touched[i] = 0; gl->tiles_touched_lang[i] = 0;
        const float* p = s->means3D + 3 * (size_t)i;
        const float* V = s->viewmatrix;
        const float* Pm = s->projmatrix;
        const float vz = xform_row(V, 2, p[0], p[1], p[2]);
        if (!(vz > 0.2f)) continue;
        const float hx = xform_row(Pm, 0, p[0], p[1], p[2]);
        const float hy = xform_row(Pm, 1, p[0], p[1], p[2]);
        const float hw = xform_row(Pm, 3, p[0], p[1], p[2]);
        const float pw = 1.0f / (hw + 0.0000001f);
        const float projx = hx * pw, projy = hy * pw;
        const float *c3, *c3l;
        if (s->cov3D_precomp) c3 = s->cov3D_precomp + 6 * (size_t)i;
        else {
            cov3d_from_scale_rot(s->scales + 3 * (size_t)i, s->scale_modifier, s->rotations + 4 * (size_t)i, g->cov3D + 6 * (size_t)i);
            c3 = g->cov3D + 6 * (size_t)i;
        }
        if (x->cov3D_precomp_lang) c3l = x->cov3D_precomp_lang + 6 * (size_t)i;
        else {
            cov3d_from_scale_rot(x->scales_lang + 3 * (size_t)i, s->scale_modifier, x->rotations_lang + 4 * (size_t)i,
                                 gl->cov3D_lang + 6 * (size_t)i);
            c3l = gl->cov3D_lang + 6 * (size_t)i;
        }
        const Cov2D cv = cov2d(p, fx, fy, s->tanfovx, s->tanfovy, c3, V);
        const Cov2D cl = cov2d(p, fx, fy, s->tanfovx, s->tanfovy, c3l, V);
        const float det = fmaf(cv.a, cv.c, -(cv.b * cv.b));
        const float det_l = fmaf(cl.a, cl.c, -(cl.b * cl.b));
        if (det == 0.0f && det_l == 0.0f) continue;  // D/forward.cu:357-366
        const float det_inv = 1.0f / det, det_inv_l = 1.0f / det_l;
        auto radius_of = [](const Cov2D& c, float d) {
            const float mid = (c.a + c.c) * 0.5f;
            const float sq = std::sqrt(fmaxf_(fmaf(mid, mid, -d), 0.1f));
            return std::ceil(std::sqrt(fmaxf_(mid + sq, mid - sq)) * 3.0f);
        };
        const int ri = f2i_rz(radius_of(cv, det)), ril = f2i_rz(radius_of(cl, det_l));
        const float px = ndc2pix(projx, W), py = ndc2pix(projy, H);
        int mn[2], mx[2], mnl[2], mxl[2];
        get_rect(px, py, ri, tile, gx, gy, mn, mx);
        get_rect(px, py, ril, tile, gx, gy, mnl, mxl);
        const uint32_t tiles = (uint32_t)(mx[0] - mn[0]) * (uint32_t)(mx[1] - mn[1]);
        const uint32_t tiles_l = (uint32_t)(mxl[0] - mnl[0]) * (uint32_t)(mxl[1] - mnl[1]);
        if (tiles == 0 && tiles_l == 0) continue;  // D/forward.cu:391-394
        if (!s->colors_precomp) {
            if (tiles == 0) { g->rgb[3 * (size_t)i] = 0; g->rgb[3 * (size_t)i + 1] = 0; g->rgb[3 * (size_t)i + 2] = 0; }
            else sh_to_rgb(i, s->sh_degree, s->M, s->means3D, s->campos, s->shs, g->clamped, g->rgb);
        }
        g->depths[i] = vz;
        g->radii[i] = ri;
        g->means2D[2 * (size_t)i] = px;
        g->means2D[2 * (size_t)i + 1] = py;
        float* co = g->conic_opacity + 4 * (size_t)i;
        co[0] = cv.c * det_inv; co[1] = cv.b * -det_inv; co[2] = cv.a * det_inv; co[3] = s->opacities[i];
        g->tiles_touched[i] = tiles;
        gl->radii_lang[i] = ril;
        float* col = gl->conic_opacity_lang + 4 * (size_t)i;
        col[0] = cl.c * det_inv_l; col[1] = cl.b * -det_inv_l; col[2] = cl.a * det_inv_l; col[3] = x->opacities_lang[i];
        gl->tiles_touched_lang[i] = tiles_l;
    }
    uint64_t acc = 0, accl = 0;
    for (int i = 0; i < P; i++) {
        acc += g->tiles_touched[i]; g->point_offsets[i] = (uint32_t)acc;
        accl += gl->tiles_touched_lang[i]; gl->point_offsets_lang[i] = (uint32_t)accl;
    }
    *R_lang = (int64_t)accl;
    return (int64_t)acc;
}

// oc: color / depth / opacity / n_touched of the colour pass; ol: language / opacity (= opacity_lang) / n_touched (= n_touched_lang)
int ols_oracle_dis_render(const OracleScene* s, const OracleGeom* g, const OracleDisGeom* gl, OracleBin* bc, OracleBin* bl,
                          OracleImage* oc, OracleImage* ol) {
    const float* feat = s->colors_precomp ? s->colors_precomp : g->rgb;
    render_pass(s, 0, true, g->radii, g->conic_opacity, g->point_offsets, g->means2D, g->depths, feat, bc, oc);
    render_pass(s, s->F, false, gl->radii_lang, gl->conic_opacity_lang, gl->point_offsets_lang, g->means2D, g->depths, nullptr,
                bl, ol);
    return 0;
}

int ols_oracle_dis_backward(const OracleScene* s, const OracleDisExtra* x, const OracleGeom* g, const OracleDisGeom* gl,
                            const OracleBin* bc, const OracleBin* bl, OracleDisGrads* gr) {
    const int P = s->P, W = s->W, H = s->H, F = s->F, M = s->M;
    const bool compat = gr->compat != 0;
    const float* feat = s->colors_precomp ? s->colors_precomp : g->rgb;
    std::vector<double> accc((size_t)P * 10, 0.0), accl((size_t)P * (10 + F), 0.0);
    blend_bwd_pass(s, 0, true, g->conic_opacity, g->means2D, g->depths, feat, bc, gr->dL_dcolor, nullptr, gr->dL_ddepth,
                   compat, accc);
    blend_bwd_pass(s, F, false, gl->conic_opacity_lang, g->means2D, g->depths, nullptr, bl, nullptr, gr->dL_dlanguage, nullptr,
                   compat, accl);
    const float fy = H / (2.0f * s->tanfovy), fx = W / (2.0f * s->tanfovx);
    const float* V = s->viewmatrix; const float* Pm = s->projmatrix; const float* Praw = s->projmatrix_raw;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        const double* a = &accc[(size_t)i * 10];
        const double* al = &accl[(size_t)i * (10 + F)];
        gr->dL_dmeans2D[3 * (size_t)i] = (float)a[0];
        gr->dL_dmeans2D[3 * (size_t)i + 1] = (float)a[1];
        gr->dL_dmeans2D[3 * (size_t)i + 2] = 0.0f;
        float* dc = gr->dL_dconic + 4 * (size_t)i; float* dcl = gr->dL_dconic_lang + 4 * (size_t)i;
        dc[0] = (float)a[2]; dc[1] = (float)a[3]; dc[2] = 0.0f; dc[3] = (float)a[4];
        dcl[0] = (float)al[2]; dcl[1] = (float)al[3]; dcl[2] = 0.0f; dcl[3] = (float)al[4];
        gr->dL_dopacity[i] = (float)a[5];
        gr->dL_dopacity_lang[i] = (float)al[5];
        for (int c = 0; c < 3; c++) gr->dL_dcolors[3 * (size_t)i + c] = (float)a[6 + c];
        gr->dL_ddepths[i] = (float)a[9];
        for (int c = 0; c < F; c++) gr->dL_dlang[(size_t)F * i + c] = (float)al[10 + c];

        float* dmean = gr->dL_dmeans3D + 3 * (size_t)i;
        float* dcov = gr->dL_dcov3D + 6 * (size_t)i;
        float* dcovl = gr->dL_dcov3D_lang + 6 * (size_t)i;
        float* dtau = gr->dL_dtau + 6 * (size_t)i;
        for (int k = 0; k < 3; k++) dmean[k] = 0;
        for (int k = 0; k < 6; k++) { dcov[k] = 0; dcovl[k] = 0; dtau[k] = 0; }
        for (int k = 0; k < 3; k++) { gr->dL_dscales[3 * (size_t)i + k] = 0; gr->dL_dscales_lang[3 * (size_t)i + k] = 0; }
        for (int k = 0; k < 4; k++) { gr->dL_drots[4 * (size_t)i + k] = 0; gr->dL_drots_lang[4 * (size_t)i + k] = 0; }
        if (gr->dL_dsh) for (int k = 0; k < 3 * M; k++) gr->dL_dsh[(size_t)3 * M * i + k] = 0;
        const bool vis_c = g->radii[i] > 0, vis_l = gl->radii_lang[i] > 0;
        const float* mp = s->means3D + 3 * (size_t)i;
        const float* c3 = s->cov3D_precomp ? s->cov3D_precomp + 6 * (size_t)i : g->cov3D + 6 * (size_t)i;
        const float* c3l = x->cov3D_precomp_lang ? x->cov3D_precomp_lang + 6 * (size_t)i : gl->cov3D_lang + 6 * (size_t)i;
        if (vis_c) cov2d_bwd_cpu(s, fx, fy, V, mp, c3, dc[0], dc[1], dc[3], dcov, dmean, dtau);  // D/backward.cu:1552
        if (vis_l) {                                                                               // D/backward.cu:1566 (no_tau)
            float dm_unused[3] = {0, 0, 0}, dt_unused[6] = {0, 0, 0, 0, 0, 0};
            cov2d_bwd_cpu(s, fx, fy, V, mp, c3l, dcl[0], dcl[1], dcl[3], dcovl, dm_unused, dt_unused);
        }
        // D/backward.cu:676-677: language_preprocessCUDA returns unless BOTH radii are positive
        const bool both = vis_c && vis_l;
        if (compat ? both : vis_c) {
            proj_bwd_cpu(V, Pm, Praw, mp, gr->dL_dmeans2D[3 * (size_t)i], gr->dL_dmeans2D[3 * (size_t)i + 1], gr->dL_ddepths[i],
                         dmean, dtau);
            sh_bwd_cpu(s, i, mp, g->clamped, gr->dL_dcolors, gr->dL_dsh, dmean, dtau);
            cov3d_bwd_cpu(s, s->scales, s->rotations, i, dcov, gr->dL_dscales, gr->dL_drots);
        }
        if (compat ? both : vis_l) cov3d_bwd_cpu(s, x->scales_lang, x->rotations_lang, i, dcovl, gr->dL_dscales_lang, gr->dL_drots_lang);
    }
    return 0;
}

int ols_oracle_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void ols_oracle_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

}  // extern "C"
