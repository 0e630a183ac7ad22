// ols_api.cu -- the C ABI declared in include/ols_b200.h (argument validation, workspace carving,
// error strings, host-buffer convenience entry).  No torch types, no exceptions across the boundary.
#include "ols_common.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

static thread_local char g_err[512] = "";

void ols_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

using namespace ols;

// ---- per-kernel timing -----------------------------------------------------------------------------
struct TimingState {
    bool on = false;
    std::vector<cudaEvent_t> ev;
    std::vector<int> tag;
    int used = 0;
};
static thread_local TimingState g_timing;

void ols_timing_mark(int tag, cudaStream_t st) {
    TimingState& t = g_timing;
    if (!t.on || t.used >= (int)t.ev.size()) return;
    if (cudaEventRecord(t.ev[t.used], st) != cudaSuccess) { cudaGetLastError(); return; }
    t.tag[t.used++] = tag;
}

namespace ols {
__global__ void k_check_frustum(int P, const float* __restrict__ means, const float* __restrict__ V,
                                uint8_t* __restrict__ present) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float x = means[3 * (size_t)i], y = means[3 * (size_t)i + 1], z = means[3 * (size_t)i + 2];
    // transformPoint4x3(...).z in the compiled reference's order (auxiliary.h:58-66,139-164)
    const float vz = fadd(ffma(z, V[10], ffma(x, V[2], fmul(y, V[6]))), V[14]);
    present[i] = vz > 0.2f ? 1 : 0;
}
}  // namespace ols

static int validate(const ols_raster_args* a, WsLayout* L, bool check_workspace = true) {
    if (!a) { ols_set_error("null args"); return OLS_ERR_INVALID; }
    if (a->P < 0 || a->W <= 0 || a->H <= 0) { ols_set_error("bad sizes P=%d W=%d H=%d", a->P, a->W, a->H); return OLS_ERR_INVALID; }
    if (!((a->tile == 15 || a->tile == 16) && (a->F == 3 || a->F == 15))) {
        ols_set_error("unsupported (tile=%d, F=%d): compiled variants are tile in {15,16} x F in {3,15}", a->tile, a->F);
        return OLS_ERR_UNSUPPORTED;
    }
    // reference: diff_gaussian_rasterization/__init__.py:514-527
    if ((a->d_shs == nullptr) == (a->d_colors_precomp == nullptr)) {
        ols_set_error("Please provide excatly one of either SHs or precomputed colors!");
        return OLS_ERR_INVALID;
    }
    const bool sr = a->d_scales != nullptr && a->d_rotations != nullptr;
    const bool any_sr = a->d_scales != nullptr || a->d_rotations != nullptr;
    if ((!sr && a->d_cov3D_precomp == nullptr) || (any_sr && a->d_cov3D_precomp != nullptr)) {
        ols_set_error("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
        return OLS_ERR_INVALID;
    }
    if (a->d_shs && (a->M <= 0 || a->sh_degree < 0 || a->sh_degree > 3 || (a->sh_degree + 1) * (a->sh_degree + 1) > a->M)) {
        ols_set_error("bad SH configuration degree=%d M=%d", a->sh_degree, a->M);
        return OLS_ERR_INVALID;
    }
    if (!a->d_means3D || !a->d_language || !a->d_opacities || !a->d_viewmatrix || !a->d_projmatrix || !a->d_bg ||
        !a->d_campos) {
        ols_set_error("a required pointer is null");
        return OLS_ERR_INVALID;
    }
    if (((a->W + a->tile - 1) / a->tile) > 65535 || ((a->H + a->tile - 1) / a->tile) > 65535) {
        ols_set_error("image too large");
        return OLS_ERR_INVALID;
    }
    *L = ws_layout(a->P, a->F, a->W, a->H, a->tile, a->R_cap);
    if (!check_workspace) return OLS_OK;
    if (!a->d_workspace || a->workspace_bytes < L->total) {
        ols_set_error("workspace too small: have %zu need %zu bytes", a->workspace_bytes, L->total);
        return OLS_ERR_WORKSPACE;
    }
    if (((uintptr_t)a->d_workspace & 255) != 0) {
        ols_set_error("workspace must be 256-byte aligned");
        return OLS_ERR_INVALID;
    }
    return OLS_OK;
}

extern "C" {

int ols_abi_version(void) { return OLS_ABI_VERSION; }
const char* ols_last_error(void) { return g_err; }

int ols_cuda_available(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n > 0 ? 1 : 0;
}

size_t ols_lang_workspace_size(int32_t P, int32_t F, int32_t W, int32_t H, int32_t tile, int64_t R_cap) {
    if (P < 0 || F <= 0 || W <= 0 || H <= 0 || tile <= 0 || R_cap < 0) return 0;
    return ws_layout(P, F, W, H, tile, R_cap).total;
}

// the views of a batch share everything but the camera, the background, the outputs and the workspace
static int validate_batch(const ols_raster_args* views, int V, WsLayout* L) {
    if (!views || V < 1 || V > OLS_MAX_BATCH_VIEWS) { ols_set_error("bad view count %d (1..%d)", V, OLS_MAX_BATCH_VIEWS); return OLS_ERR_INVALID; }
    static_assert(OLS_MAX_BATCH_VIEWS == OLS_MAX_VIEWS, "header / kernel batch limits");
    for (int v = 0; v < V; v++) {
        WsLayout Lv;
        int rc = validate(&views[v], &Lv);
        if (rc != OLS_OK) return rc;
        if (v == 0) { *L = Lv; continue; }
        const ols_raster_args &a = views[0], &b = views[v];
        if (a.P != b.P || a.F != b.F || a.sh_degree != b.sh_degree || a.M != b.M || a.W != b.W || a.H != b.H || a.tile != b.tile ||
            a.flags != b.flags || a.scale_modifier != b.scale_modifier || a.R_cap != b.R_cap || a.d_means3D != b.d_means3D ||
            a.d_shs != b.d_shs || a.d_colors_precomp != b.d_colors_precomp || a.d_language != b.d_language ||
            a.d_opacities != b.d_opacities || a.d_scales != b.d_scales || a.d_rotations != b.d_rotations ||
            a.d_cov3D_precomp != b.d_cov3D_precomp) {
            ols_set_error("view %d of the batch differs from view 0 in a size, a flag, R_cap or a Gaussian parameter pointer", v);
            return OLS_ERR_INVALID;
        }
        if (a.d_workspace == b.d_workspace) { ols_set_error("view %d shares its workspace with view 0", v); return OLS_ERR_INVALID; }
    }
    return OLS_OK;
}

int ols_lang_forward_batch(const ols_raster_args* views, const ols_fwd_out* outs, int32_t V, void* stream) {
    WsLayout L;
    int rc = validate_batch(views, V, &L);
    if (rc != OLS_OK) return rc;
    for (int v = 0; v < V; v++) {
        const ols_fwd_out* o = outs ? &outs[v] : nullptr;
        if (!o || !o->d_color || !o->d_language || !o->d_depth || !o->d_opacity || !o->d_radii || !o->d_n_touched) {
            ols_set_error("null output pointer");
            return OLS_ERR_INVALID;
        }
    }
    if (views[0].P == 0) { ols_set_error("P == 0: nothing to render (reference returns None)"); return OLS_ERR_INVALID; }
    return ols_launch_forward(views, outs, V, L, (cudaStream_t)stream);
}

int ols_lang_forward(const ols_raster_args* a, const ols_fwd_out* o, void* stream) {
    return ols_lang_forward_batch(a, o, 1, stream);
}

int ols_lang_read_info_async(const ols_raster_args* views, int32_t V, ols_fwd_info* h_info, void* stream) {
    if (!views || !h_info || V < 1 || V > OLS_MAX_BATCH_VIEWS) { ols_set_error("bad arguments"); return OLS_ERR_INVALID; }
    for (int v = 0; v < V; v++) {
        if (!views[v].d_workspace) { ols_set_error("null workspace"); return OLS_ERR_INVALID; }
        OLS_CUDA_TRY(cudaMemcpyAsync(&h_info[v], views[v].d_workspace, sizeof(ols_fwd_info), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    }
    return OLS_OK;
}

int ols_lang_read_info(const void* d_workspace, ols_fwd_info* h_info, void* stream) {
    if (!d_workspace || !h_info) { ols_set_error("null pointer"); return OLS_ERR_INVALID; }
    static_assert(sizeof(ols_fwd_info) == sizeof(DeviceInfo), "info header layout");
    OLS_CUDA_TRY(cudaMemcpyAsync(h_info, d_workspace, sizeof(ols_fwd_info), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    OLS_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return OLS_OK;
}

int ols_lang_backward_batch(const ols_raster_args* views, const ols_bwd_args* grads, int32_t V, void* stream) {
    WsLayout L;
    int rc = validate_batch(views, V, &L);
    if (rc != OLS_OK) return rc;
    if (!grads) { ols_set_error("null gradient pointer"); return OLS_ERR_INVALID; }
    const ols_raster_args* a = &views[0];
    const ols_bwd_args* g = &grads[0];
    if (!g->d_dL_dcolors || !g->d_dL_dlanguage || !g->d_dL_dopacity || !g->d_dL_dmeans3D || !g->d_dL_dcov3D ||
        !g->d_dL_dscales || !g->d_dL_drotations) {
        ols_set_error("null gradient pointer");
        return OLS_ERR_INVALID;
    }
    for (int v = 0; v < V; v++) {
        const ols_bwd_args* q = &grads[v];
        if (!q->d_dL_dout_color || !q->d_dL_dout_language || !q->d_dL_dout_depth || !q->d_radii || !q->d_dL_dmeans2D ||
            (!q->d_dL_dtau && !q->d_dL_dtau_sum)) {
            ols_set_error("null per-view gradient pointer (view %d)", v);
            return OLS_ERR_INVALID;
        }
        if (!views[v].d_projmatrix_raw) { ols_set_error("projmatrix_raw is required by backward"); return OLS_ERR_INVALID; }
    }
    if (a->M > 0 && a->d_shs && !g->d_dL_dsh) { ols_set_error("d_dL_dsh is null but SHs were given"); return OLS_ERR_INVALID; }
    return ols_launch_backward(views, grads, V, L, (cudaStream_t)stream);
}

int ols_lang_backward(const ols_raster_args* a, const ols_bwd_args* g, void* stream) {
    return ols_lang_backward_batch(a, g, 1, stream);
}

int ols_timing_begin(int32_t max_marks) {
    TimingState& t = g_timing;
    if (max_marks <= 0) { ols_set_error("max_marks must be positive"); return OLS_ERR_INVALID; }
    while ((int)t.ev.size() < max_marks) {
        cudaEvent_t e;
        OLS_CUDA_TRY(cudaEventCreate(&e));
        t.ev.push_back(e);
    }
    t.tag.assign(t.ev.size(), -1);
    t.used = 0;
    t.on = true;
    return OLS_OK;
}

int ols_timing_end(float* ms_per_tag, int32_t* count_per_tag) {
    TimingState& t = g_timing;
    t.on = false;
    if (!ms_per_tag || !count_per_tag) { ols_set_error("null pointer"); return OLS_ERR_INVALID; }
    for (int i = 0; i < OLS_TIMING_TAGS; i++) { ms_per_tag[i] = 0.0f; count_per_tag[i] = 0; }
    if (t.used > 0) OLS_CUDA_TRY(cudaEventSynchronize(t.ev[t.used - 1]));
    for (int i = 1; i < t.used; i++) {
        const int tag = t.tag[i];
        if (tag < 0 || tag >= OLS_TIMING_TAGS) continue;
        float ms = 0.0f;
        OLS_CUDA_TRY(cudaEventElapsedTime(&ms, t.ev[i - 1], t.ev[i]));
        ms_per_tag[tag] += ms;
        count_per_tag[tag] += 1;
    }
    t.used = 0;
    return OLS_OK;
}

// ---- disentangled variant ---------------------------------------------------------------------------
struct DisLayout { WsLayout c, l; size_t lang_base, total; };
static DisLayout dis_layout(int P, int F, int W, int H, int tile, int64_t R_cap, int64_t R_cap_lang) {
    DisLayout D;
    D.c = ws_layout(P, 0, W, H, tile, R_cap, 3);        // colour + depth list
    D.l = ws_layout(P, F, W, H, tile, R_cap_lang, 0);   // language list
    D.lang_base = align_up(D.c.total, 256);
    D.total = D.lang_base + D.l.total;
    return D;
}

static int validate_dis(const ols_dis_args* d, DisLayout* D) {
    if (!d) { ols_set_error("null args"); return OLS_ERR_INVALID; }
    WsLayout tmp;
    int rc = validate(&d->base, &tmp, false);
    if (rc != OLS_OK) return rc;
    const ols_raster_args* a = &d->base;
    // reference: D/diff_gaussian_rasterization/__init__.py:600-605
    const bool sr = d->d_scales_lang != nullptr && d->d_rotations_lang != nullptr;
    const bool any_sr = d->d_scales_lang != nullptr || d->d_rotations_lang != nullptr;
    if ((!sr && d->d_cov3D_precomp_lang == nullptr) || (any_sr && d->d_cov3D_precomp_lang != nullptr)) {
        ols_set_error("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance for language!");
        return OLS_ERR_INVALID;
    }
    if (!d->d_opacities_lang) { ols_set_error("opacities_lang is required"); return OLS_ERR_INVALID; }
    if (d->R_cap_lang < 0) { ols_set_error("bad R_cap_lang"); return OLS_ERR_INVALID; }
    *D = dis_layout(a->P, a->F, a->W, a->H, a->tile, a->R_cap, d->R_cap_lang);
    if (!a->d_workspace || a->workspace_bytes < D->total) {
        ols_set_error("workspace too small: have %zu need %zu bytes", a->workspace_bytes, D->total);
        return OLS_ERR_WORKSPACE;
    }
    if (((uintptr_t)a->d_workspace & 255) != 0) { ols_set_error("workspace must be 256-byte aligned"); return OLS_ERR_INVALID; }
    return OLS_OK;
}

static void fill_view(ols_ws_view* v, const char* ws, const WsLayout& L) {
    v->d_records = (const float*)(ws + L.records);
    v->rec_floats = L.rec;
    v->n_tiles = L.n_tiles;
    v->d_cov3D = (const float*)(ws + L.cov3D);
    v->d_clamped = (const uint8_t*)(ws + L.clamped);
    v->d_tiles_touched = (const uint32_t*)(ws + L.tiles_touched);
    v->d_ranges = (const uint32_t*)(ws + L.ranges);
    v->d_point_list = (const uint32_t*)(ws + L.point_list);
    v->d_keys = (const uint64_t*)(ws + L.keys);
    v->d_final_T = (const float*)(ws + L.final_T);
    v->d_n_contrib = (const uint32_t*)(ws + L.n_contrib);
}

size_t ols_dis_workspace_size(int32_t P, int32_t F, int32_t W, int32_t H, int32_t tile, int64_t R_cap, int64_t R_cap_lang) {
    if (P < 0 || F <= 0 || W <= 0 || H <= 0 || tile <= 0 || R_cap < 0 || R_cap_lang < 0) return 0;
    return dis_layout(P, F, W, H, tile, R_cap, R_cap_lang).total;
}

int ols_dis_forward(const ols_dis_args* d, const ols_dis_fwd_out* o, void* stream) {
    DisLayout D;
    int rc = validate_dis(d, &D);
    if (rc != OLS_OK) return rc;
    if (!o || !o->d_color || !o->d_language || !o->d_depth || !o->d_opacity || !o->d_opacity_lang || !o->d_radii ||
        !o->d_radii_lang || !o->d_n_touched || !o->d_n_touched_lang) {
        ols_set_error("null output pointer");
        return OLS_ERR_INVALID;
    }
    if (d->base.P == 0) { ols_set_error("P == 0: nothing to render"); return OLS_ERR_INVALID; }
    return ols_launch_forward_dis(d, o, D.c, D.l, D.lang_base, (cudaStream_t)stream);
}

int ols_dis_read_info(const ols_dis_args* d, ols_fwd_info* h_c, ols_fwd_info* h_l, void* stream) {
    if (!d || !h_c || !h_l || !d->base.d_workspace) { ols_set_error("null pointer"); return OLS_ERR_INVALID; }
    const ols_raster_args* a = &d->base;
    const DisLayout D = dis_layout(a->P, a->F, a->W, a->H, a->tile, a->R_cap, d->R_cap_lang);
    const char* ws = (const char*)a->d_workspace;
    OLS_CUDA_TRY(cudaMemcpyAsync(h_c, ws + D.c.info, sizeof(ols_fwd_info), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    OLS_CUDA_TRY(cudaMemcpyAsync(h_l, ws + D.lang_base + D.l.info, sizeof(ols_fwd_info), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    OLS_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return OLS_OK;
}

int ols_dis_backward(const ols_dis_args* d, const ols_dis_bwd_args* g, void* stream) {
    DisLayout D;
    int rc = validate_dis(d, &D);
    if (rc != OLS_OK) return rc;
    if (!g || !g->d_dL_dout_color || !g->d_dL_dout_language || !g->d_dL_dout_depth || !g->d_radii || !g->d_radii_lang ||
        !g->d_dL_dmeans2D || !g->d_dL_dcolors || !g->d_dL_dlanguage || !g->d_dL_dopacity || !g->d_dL_dopacity_lang ||
        !g->d_dL_dmeans3D || !g->d_dL_dcov3D || !g->d_dL_dcov3D_lang || !g->d_dL_dscales || !g->d_dL_dscales_lang ||
        !g->d_dL_drotations || !g->d_dL_drotations_lang || !g->d_dL_dtau) {
        ols_set_error("null gradient pointer");
        return OLS_ERR_INVALID;
    }
    const ols_raster_args* a = &d->base;
    if (a->M > 0 && a->d_shs && !g->d_dL_dsh) { ols_set_error("d_dL_dsh is null but SHs were given"); return OLS_ERR_INVALID; }
    if (!a->d_projmatrix_raw) { ols_set_error("projmatrix_raw is required by backward"); return OLS_ERR_INVALID; }
    if (a->flags & OLS_FLAG_BWD_ACCUMULATE) { ols_set_error("accumulate is not supported by the disentangled backward"); return OLS_ERR_UNSUPPORTED; }
    return ols_launch_backward_dis(d, g, D.c, D.l, D.lang_base, (cudaStream_t)stream);
}

int ols_dis_workspace_view(const ols_dis_args* d, ols_ws_view* vc, ols_ws_view* vl) {
    if (!d || !vc || !vl || !d->base.d_workspace) { ols_set_error("bad arguments"); return OLS_ERR_INVALID; }
    const ols_raster_args* a = &d->base;
    const DisLayout D = dis_layout(a->P, a->F, a->W, a->H, a->tile, a->R_cap, d->R_cap_lang);
    fill_view(vc, (const char*)a->d_workspace, D.c);
    fill_view(vl, (const char*)a->d_workspace + D.lang_base, D.l);
    return OLS_OK;
}

int ols_mark_visible(int32_t P, const float* d_means3D, const float* d_viewmatrix, const float* d_projmatrix,
                     uint8_t* d_present, void* stream) {
    (void)d_projmatrix;  // the reference's x/y NDC test is commented out (auxiliary.h:152)
    if (P < 0 || (P > 0 && (!d_means3D || !d_viewmatrix || !d_present))) { ols_set_error("bad arguments"); return OLS_ERR_INVALID; }
    if (P == 0) return OLS_OK;
    k_check_frustum<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, d_means3D, d_viewmatrix, d_present);
    OLS_CUDA_TRY(cudaGetLastError());
    return OLS_OK;
}

int ols_lang_workspace_view(int32_t P, int32_t F, int32_t W, int32_t H, int32_t tile, int64_t R_cap,
                            const void* d_workspace, ols_ws_view* v) {
    if (!d_workspace || !v || P < 0 || F <= 0 || W <= 0 || H <= 0 || tile <= 0) { ols_set_error("bad arguments"); return OLS_ERR_INVALID; }
    const WsLayout L = ws_layout(P, F, W, H, tile, R_cap);
    const char* ws = (const char*)d_workspace;
    v->d_records = (const float*)(ws + L.records);
    v->rec_floats = L.rec;
    v->n_tiles = L.n_tiles;
    v->d_cov3D = (const float*)(ws + L.cov3D);
    v->d_clamped = (const uint8_t*)(ws + L.clamped);
    v->d_tiles_touched = (const uint32_t*)(ws + L.tiles_touched);
    v->d_ranges = (const uint32_t*)(ws + L.ranges);
    v->d_point_list = (const uint32_t*)(ws + L.point_list);
    v->d_keys = (const uint64_t*)(ws + L.keys);
    v->d_final_T = (const float*)(ws + L.final_T);
    v->d_n_contrib = (const uint32_t*)(ws + L.n_contrib);
    return OLS_OK;
}

// Host-buffer forward: H2D, render (growing the instance capacity if needed), D2H.
int ols_lang_forward_host(const ols_raster_args* h, const ols_host_out* out, int64_t* num_rendered) {
    if (!h || !out) { ols_set_error("null args"); return OLS_ERR_INVALID; }
    if (h->P <= 0) { ols_set_error("P must be > 0"); return OLS_ERR_INVALID; }
    const size_t P = (size_t)h->P, HW = (size_t)h->W * h->H;
    struct Buf { const float* src; size_t n; const float** dst; };
    ols_raster_args d = *h;
    std::vector<void*> owned;
    auto cleanup = [&]() { for (void* p : owned) cudaFree(p); };
    Buf bufs[] = {{h->d_bg, 3, &d.d_bg}, {h->d_means3D, 3 * P, &d.d_means3D}, {h->d_shs, 3 * P * (size_t)h->M, &d.d_shs},
                  {h->d_colors_precomp, 3 * P, &d.d_colors_precomp}, {h->d_language, P * (size_t)h->F, &d.d_language},
                  {h->d_opacities, P, &d.d_opacities}, {h->d_scales, 3 * P, &d.d_scales},
                  {h->d_rotations, 4 * P, &d.d_rotations}, {h->d_cov3D_precomp, 6 * P, &d.d_cov3D_precomp},
                  {h->d_viewmatrix, 16, &d.d_viewmatrix}, {h->d_projmatrix, 16, &d.d_projmatrix},
                  {h->d_projmatrix_raw, 16, &d.d_projmatrix_raw}, {h->d_campos, 3, &d.d_campos}};
    for (auto& b : bufs) {
        *b.dst = nullptr;
        if (!b.src || b.n == 0) continue;
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, b.n * sizeof(float));
        if (e == cudaSuccess) { owned.push_back(p); e = cudaMemcpy(p, b.src, b.n * sizeof(float), cudaMemcpyHostToDevice); }
        if (e != cudaSuccess) { ols_set_error("H2D failed: %s", cudaGetErrorString(e)); cleanup(); return OLS_ERR_CUDA; }
        *b.dst = (const float*)p;
    }
    ols_fwd_out o;
    const size_t F = (size_t)h->F;
    void* img = nullptr; void* ints = nullptr;
    if (cudaMalloc(&img, sizeof(float) * HW * (3 + F + 2)) != cudaSuccess || cudaMalloc(&ints, 8 * P) != cudaSuccess) {
        ols_set_error("out of device memory"); if (img) cudaFree(img); cleanup(); return OLS_ERR_CUDA;
    }
    owned.push_back(img); owned.push_back(ints);
    o.d_color = (float*)img; o.d_language = o.d_color + 3 * HW; o.d_depth = o.d_language + F * HW;
    o.d_opacity = o.d_depth + HW; o.d_radii = (int32_t*)ints; o.d_n_touched = o.d_radii + P;
    int64_t cap = (int64_t)P * 4 + 1024;
    ols_fwd_info info;
    for (int attempt = 0; attempt < 4; attempt++) {
        d.R_cap = cap;
        d.workspace_bytes = ols_lang_workspace_size(h->P, h->F, h->W, h->H, h->tile, cap);
        void* ws = nullptr;
        if (cudaMalloc(&ws, d.workspace_bytes) != cudaSuccess) { ols_set_error("out of device memory (workspace)"); cleanup(); return OLS_ERR_CUDA; }
        d.d_workspace = ws;
        int rc = ols_lang_forward(&d, &o, nullptr);
        if (rc == OLS_OK) rc = ols_lang_read_info(ws, &info, nullptr);
        cudaFree(ws);
        if (rc != OLS_OK) { cleanup(); return rc; }
        if (!info.overflow) break;
        cap = info.R + 1024;
    }
    if (info.overflow) { ols_set_error("instance capacity overflow"); cleanup(); return OLS_ERR_OVERFLOW; }
    if (num_rendered) *num_rendered = info.R;
    cudaError_t e = cudaSuccess;
    auto d2h = [&](void* dst, const void* src, size_t bytes) { if (dst && e == cudaSuccess) e = cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost); };
    d2h(out->h_color, o.d_color, 12 * HW); d2h(out->h_language, o.d_language, 4 * F * HW);
    d2h(out->h_depth, o.d_depth, 4 * HW); d2h(out->h_opacity, o.d_opacity, 4 * HW);
    d2h(out->h_radii, o.d_radii, 4 * P); d2h(out->h_n_touched, o.d_n_touched, 4 * P);
    cleanup();
    if (e != cudaSuccess) { ols_set_error("D2H failed: %s", cudaGetErrorString(e)); return OLS_ERR_CUDA; }
    return OLS_OK;
}

}  // extern "C"
