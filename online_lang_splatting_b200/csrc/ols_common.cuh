// ols_common.cuh -- shared definitions of the B200 rasterizer kernels (workspace layout, helpers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/ols_b200.h"

namespace ols {

// ---- packed per-Gaussian blend record ------------------------------------------------------------
// floats: 0 x | 1 y | 2 conicA | 3 conicB | 4 conicC | 5 opacity | 6 pth | 7 depth | 8.. channels[NCH] | 0-pad | ex | ey
// rounded up to a multiple of 4 floats so a record is a whole number of 16-byte chunks.  The channels
// are what one blending pass accumulates: rgb + lang[F] for the joint pass of P/ (NCH = 3 + F; F=15 ->
// 28 floats = 112 B, F=3 -> 16 floats), rgb only (NCH = 3) or lang[F] only (NCH = F) for the two passes
// of the disentangled variant D/.  `pth` is a conservative lower bound on the exponent below which
// alpha < 1/255 is certain, so the blend kernels can skip expf() for far pixels without changing any
// decision of the reference (forward.cu:446-457).  The last two floats hold conservative half-extents
// (ex, ey) of the region where alpha >= 1/255 is possible; the blend kernels use them to reject a
// Gaussian for a whole warp's pixel block at once.
constexpr int REC_X = 0, REC_Y = 1, REC_A = 2, REC_B = 3, REC_C = 4, REC_OP = 5, REC_PTH = 6, REC_DEPTH = 7,
              REC_CH = 8;
__host__ __device__ constexpr int rec_floats_nch(int nch) { return ((10 + nch) + 3) / 4 * 4; }
__host__ __device__ constexpr int rec_floats(int F) { return rec_floats_nch(3 + F); }  // joint pass, as staged in shared memory
// In global memory the joint pass keeps only the 16-float colour record (x .. depth | r g b 0 | 0 0 ex ey); the
// language channels are staged from the caller's language[P,F] rows when a tile's batch is fetched, so the
// preprocess neither reads nor copies them (only the few entries a tile really traverses are ever fetched).
__host__ __device__ constexpr int rec_floats_global(int ncol, int F) { return (ncol && F) ? rec_floats_nch(ncol) : rec_floats_nch(ncol + F); }

struct WsLayout {
    size_t info, tile_count, tile_cursor, ranges, cta_hist, records, depths, cov3D, clamped, tiles_touched, rect, final_T,
        n_contrib, keys, point_list, warp_hits, gacc, gtouched, total;
    int n_tiles, gx, gy, rec;
    int n_ctas, chunk;  // per-Gaussian kernels: n_ctas CTAs, each owning `chunk` consecutive Gaussians
};

inline __host__ size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Single workspace replacing geomBuffer / binningBuffer / imgBuffer (rasterizer_impl.cu:155-212).
// `ncol` = 3 for passes that blend colour (joint / colour-only), 0 for the language-only pass of D/.
inline __host__ WsLayout ws_layout(int P, int F, int W, int H, int tile, int64_t R_cap, int ncol = 3) {
    WsLayout L;
    L.gx = (W + tile - 1) / tile;
    L.gy = (H + tile - 1) / tile;
    L.n_tiles = L.gx * L.gy;
    L.rec = rec_floats_global(ncol, F);
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 256); return at; };
    const size_t Pz = (size_t)(P > 0 ? P : 1), HW = (size_t)W * H, Rz = (size_t)(R_cap > 0 ? R_cap : 1);
    L.info = take(256);
    L.tile_count = take(4 * (size_t)L.n_tiles);
    L.tile_cursor = take(4 * (size_t)L.n_tiles);
    L.ranges = take(8 * (size_t)L.n_tiles);
    {
        const int blocks = (int)((Pz + 255) / 256);
        L.n_ctas = blocks < 592 ? blocks : 592;  // 4 CTAs per SM on 148 SMs
        L.chunk = (int)(((Pz + L.n_ctas - 1) / L.n_ctas + 255) / 256 * 256);
    }
    L.cta_hist = take(4 * (size_t)L.n_tiles * L.n_ctas);
    L.records = take(4 * (size_t)L.rec * Pz);
    L.depths = take(4 * Pz);
    L.cov3D = take(24 * Pz);
    L.clamped = take(4 * Pz);  // 3 flags packed in the low bytes of a u32
    L.tiles_touched = take(4 * Pz);
    L.rect = take(8 * Pz);
    L.final_T = take(4 * HW);
    L.n_contrib = take(4 * HW);
    L.keys = take(8 * Rz);
    L.point_list = take(4 * Rz);
    L.warp_hits = take(Rz);  // per list entry: bit w set iff some pixel of the tile's 8x4 block w blended it (forward -> backward)
    L.gacc = take(4 * (size_t)(((10 + F) + 3) / 4 * 4) * Pz);  // packed per-Gaussian gradient records (backward)
    L.gtouched = take(Pz);  // one byte per Gaussian: the backward blend flushed something into its gradient record
    L.total = o;
    return L;
}

// ---- multi-view batches ----------------------------------------------------------------------------
// The views of one mapping iteration (utils/slam_backend.py:510-662 renders the 8-12 window keyframes one after the
// other) are rendered by ONE set of launches: blockIdx.y selects the view, every view owns a workspace with the same
// layout, and the kernels receive the per-view pointers as a __grid_constant__ array (constant-bank indexed loads).
constexpr int OLS_MAX_VIEWS = 16;

struct PassView {           // one view of one blending pass (binning + sort + blend)
    char* ws;               // the view's workspace
    const float* depths;    // [P] view-space depths (D/ shares them between its two passes)
    const float* bg;
    float *color, *language, *depth, *opacity;
    int32_t* n_touched;
};
struct PassBatch { PassView v[OLS_MAX_VIEWS]; };

struct DeviceInfo {  // lives at workspace offset 0; first 32 bytes == ols_fwd_info
    unsigned long long R;
    int overflow;
    int max_tile_len;
    int n_visible;
    int pad[3];
};

// ---- arithmetic helpers: explicit rounding so nvcc cannot re-associate / contract ----------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}

// Stages piece `op` of one Gaussian's blend record into shared memory (layout: x..depth | channels | pad | ex ey).
// Without external language the record is copied as REC/4 16-byte chunks; with it, the 16-float colour record is
// split (header 2 x 16 B, r g 8 B, b 4 B, extents 8 B) and the F language floats come from language[id*F..].
template <int NCOL, int F>
struct RecordStage {
    static constexpr bool EXT_LANG = NCOL > 0 && F > 0;
    static constexpr int REC_S = rec_floats_nch(NCOL + F);      // shared-memory record
    static constexpr int REC_G = rec_floats_global(NCOL, F);    // global-memory record
    static constexpr int OPS = EXT_LANG ? 5 + F : REC_S / 4;
    static_assert(!EXT_LANG || REC_CH + NCOL + F == REC_S - 2, "joint layout: language ends where the extents start");
    __device__ static __forceinline__ void copy(float* s, const float* __restrict__ records,
                                                const float* __restrict__ language, uint32_t id, int op) {
        const float* g = records + (size_t)id * REC_G;
        if (!EXT_LANG) {
            cp_async16(s + 4 * op, g + 4 * op);
        } else if (op < 2) {
            cp_async16(s + 4 * op, g + 4 * op);
        } else if (op == 2) {
            cp_async8(s + REC_CH, g + REC_CH);
        } else if (op == 3) {
            cp_async4(s + REC_CH + 2, g + REC_CH + 2);
        } else if (op == 4) {
            cp_async8(s + REC_S - 2, g + REC_G - 2);
        } else {
            cp_async4(s + REC_CH + NCOL + (op - 5), language + (size_t)id * F + (op - 5));
        }
    }
};

}  // namespace ols

// error plumbing (ols_api.cu)
void ols_set_error(const char* fmt, ...);
#define OLS_CUDA_TRY(expr)                                                                     \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ols_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return OLS_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

// timing marks (ols_api.cu): tag < 0 starts a sequence, tag >= 0 closes the interval since the previous mark
void ols_timing_mark(int tag, cudaStream_t st);

// kernel launchers implemented in ols_forward.cu / ols_backward.cu
int ols_launch_forward(const ols_raster_args* views, const ols_fwd_out* outs, int V, const ols::WsLayout& L, cudaStream_t st);
int ols_launch_backward(const ols_raster_args* views, const ols_bwd_args* grads, int V, const ols::WsLayout& L, cudaStream_t st);
int ols_launch_forward_dis(const ols_dis_args* d, const ols_dis_fwd_out* o, const ols::WsLayout& Lc, const ols::WsLayout& Ll,
                           size_t lang_base, cudaStream_t st);
int ols_launch_backward_dis(const ols_dis_args* d, const ols_dis_bwd_args* g, const ols::WsLayout& Lc, const ols::WsLayout& Ll,
                            size_t lang_base, cudaStream_t st);
