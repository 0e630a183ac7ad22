// ols_backward.cu -- backward pass of the language-feature Gaussian rasterizer for sm_100a.
//
//   k_blend_bwd       one CTA per tile, one pixel per thread, back-to-front over the tile's sorted list
//                     (reference: language_render_cuda, backward.cu:932-1201).  Per-pixel gradients of a
//                     Gaussian are summed over the warp with a register butterfly (31 shuffles for all
//                     25 values at F=15), accumulated per batch in shared memory and flushed with one
//                     global atomic per (tile, Gaussian, value) into a packed gradient record.
//   k_geometry_bwd    one thread per Gaussian: conic -> cov2D -> cov3D/mean/pose gradients, projection,
//                     depth, SH, scale/rotation (reference: computeCov2DCUDA backward.cu:150-346 and
//                     language_preprocessCUDA :541-682, fused; the packed record is unpacked here).
//
// Two gradient modes ("parity policy"):
//   compat  reproduces the reference build's behaviour: Q1 language gradient from the tile's first pixel
//           only, Q2 language recurrence also advanced by non-contributing Gaussians the block visits,
//           Q3 the lossy 225-thread tree reduction (lane mask) at 15x15 tiles;
//   exact   the mathematically correct gradient (validated against finite differences and the oracle).
//
// The per-pixel recurrence of the reference (accum_rec[ch], last_color[ch]) only enters dL/dalpha through
// sum_ch (c_ch - accum_rec_ch) * dL/dpix_ch.  Both modes carry that sum as two scalars
//   A <- last_alpha * D_last + (1 - last_alpha) * A,   D = sum_ch c_ch * dL/dpix_ch
// which is algebraically identical and keeps the per-thread state at ~25 registers instead of ~60.
#include "ols_common.cuh"

#include <cstdlib>

namespace ols {

constexpr int BWD_THREADS = 256;
constexpr int BWD_BATCH = 32;

__host__ __device__ constexpr int grad_floats(int F) { return ((10 + F) + 3) / 4 * 4; }
// packed gradient record: 0 dmean2D.x | 1 dmean2D.y | 2 dconic.x | 3 dconic.y | 4 dconic.w | 5 dopacity |
//                         6 ddepth | 7..9 dcolor | 10.. dlanguage[F]
constexpr int GR_MX = 0, GR_MY = 1, GR_CX = 2, GR_CY = 3, GR_CW = 4, GR_OP = 5, GR_DEPTH = 6, GR_RGB = 7, GR_LANG = 10;

struct BwdView {  // one view of a batch (blockIdx.y)
    const uint2* ranges;
    const uint32_t* point_list;
    const float* records;
    const float* bg;
    const DeviceInfo* info;
    const float* final_T;
    const uint32_t* n_contrib;
    const float* dL_dcolor;
    const float* dL_dlanguage;
    const float* dL_ddepth;
    const uint8_t* warp_hits;  // [R] from the forward: which pixel blocks of the tile blended each list entry
    float* gacc;             // [P, grad_floats(F)] zero-initialised
    uint8_t* gtouched;       // [P] zero-initialised: set to 1 for every Gaussian whose record receives a flush
};
struct BwdBlendArgs {
    int W, H, gx;
    int fast_exp;            // the forward blended with ex2.approx (no OLS_FLAG_BITEXACT_BLEND): use the same alpha here
    const float* language;   // [P,F] caller's language rows (joint pass), else unused
    uint32_t lane_ok[8];     // Q3 lane mask (compat); all ones otherwise
    uint8_t packed_rank[128];  // packed compat variant: reference thread rank handled by thread t (255 = none)
    uint8_t packed_fmask[4];   // packed variant: forward 8x4 pixel blocks (bit w) that contain a pixel of packed warp k
    BwdView v[OLS_MAX_VIEWS];
};

typedef unsigned long long f32x2;  // two floats in one 64-bit register pair
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float fast_exp2(float x) {  // same instruction sequence as the forward's fast_exp
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}
__device__ __forceinline__ float hsum2(f32x2 v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo + hi;
}

// Sum N per-lane values over the 32 lanes of a warp with N/2 + N/4 + ... shuffles instead of 5 * N: at the
// level that exchanges lane bit OFF, the lanes with the bit clear keep the lower half of the values and the
// others the upper half (an odd middle value is kept by both).  Afterwards v[0] of lane L holds the total
// of value multi_reduce_owner<N>(L) (or a duplicate when that returns -1).
template <int N, int OFF>
__device__ __forceinline__ void multi_reduce_level(float* v, int lane) {
    constexpr int H = (N + 1) / 2;
    if (N > 1) {
        const bool hi = (lane & OFF) != 0;
#pragma unroll
        for (int i = 0; i < N / 2; i++) {
            const float send = hi ? v[i] : v[i + H];
            const float keep = hi ? v[i + H] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        if (N & 1) v[H - 1] += __shfl_xor_sync(0xffffffffu, v[H - 1], OFF);
    } else {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], OFF);
    }
    if constexpr (OFF > 1) multi_reduce_level<H, OFF / 2>(v, lane);
}
template <int N>
__device__ __forceinline__ void warp_multi_reduce(float* v, int lane) {
    static_assert(N >= 1 && N <= 32, "one value per lane at the end");
    multi_reduce_level<N, 16>(v, lane);
}
template <int N>
__device__ __forceinline__ int multi_reduce_owner(int lane) {
    int ns[6];
    ns[0] = N;
#pragma unroll
    for (int k = 0; k < 5; k++) ns[k + 1] = (ns[k] + 1) / 2;
    int s = 0;
    bool owner = true;
#pragma unroll
    for (int k = 4; k >= 0; k--) {
        const int n = ns[k], h = (n + 1) / 2;
        const bool b = (lane & (16 >> k)) != 0;
        if (n == 1) owner = owner && !b;
        else if (s < n / 2) s += b ? h : 0;
        else owner = owner && !b;  // the odd middle value lives in both halves
    }
    return owner ? s : -1;
}

// NCOL = 3: colour + depth channels take part (joint pass of P/, colour pass of D/); NCOL = 0: the
// language-only pass of D/ (no mean2D gradient, no background term: D/backward.cu:1318-1427).
// Thread -> pixel mapping as in the forward: warp w owns the 8x4 pixel block ((w & 1) * 8, (w >> 1) * 4).
// The forward left one byte per list entry saying which pixel blocks blended it (warp_hits); an entry is
// "visited" by the reference's block-wide loop iff that byte is non-zero (a pixel blends an entry in the
// forward exactly when the backward's skip tests pass for it), so no separate scan of the list is needed.
// PACKED (compat, 15x15 tiles): the reference's lossy block reduction (Q3) lets only 128 of the 225 pixels of a
// tile reach any reduced gradient, and nothing else a dropped pixel computes is ever used.  The packed variant
// therefore runs 128 threads per tile, one per surviving pixel (in reference rank order), instead of carrying 97
// dead lanes through every step.  Its warps are 15-pixel-wide strips, so the forward's per-8x4-block hit bits do
// not apply; a warp decides by a vote after evaluating an entry.
template <int TILE, int NCOL, int F, bool COMPAT, bool PACKED>
__global__ void __launch_bounds__(PACKED ? 128 : BWD_THREADS, PACKED ? 8 : 4) k_blend_bwd(const __grid_constant__ BwdBlendArgs a) {
    const BwdView& vw = a.v[blockIdx.y];
    static_assert(TILE <= 16 && BWD_BATCH == 32, "8 warps of 8x4 pixels; one ballot per batch");
    static_assert(!PACKED || COMPAT, "packing follows the compat lane mask");
    constexpr int NT = PACKED ? 128 : BWD_THREADS;
    static_assert(NCOL == 0 || NCOL == 3, "colour channels");
    constexpr int NCH = NCOL + F;
    constexpr int REC = rec_floats_nch(NCH);
    using Stage = RecordStage<NCOL, F>;
    constexpr int OPS = Stage::OPS;
    constexpr int GR = grad_floats(F);
    constexpr int NPAIR = (NCH + 1) / 2;          // channel pairs as stored from REC_CH on
    constexpr int LP0 = (NCOL + 1) / 2;           // first pair made of language channels only
    constexpr int NGEO = NCOL ? 10 : 4;           // reduced values before the language block
    constexpr int NV = COMPAT ? NGEO : NGEO + F;  // values reduced per (warp, Gaussian)
    __shared__ __align__(16) float s_rec[BWD_BATCH * REC];
    __shared__ uint32_t s_id[BWD_BATCH];
    __shared__ float s_acc[BWD_BATCH * GR];
    __shared__ uint32_t s_maxc;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile_x = blockIdx.x % a.gx, tile_y = blockIdx.x / a.gx;
    int lx, ly;
    if (PACKED) {
        const int rank = a.packed_rank[tid];
        lx = rank == 255 ? TILE : rank % TILE;
        ly = rank == 255 ? TILE : rank / TILE;
    } else {
        lx = (wid & 1) * 8 + (lane & 7);
        ly = (wid >> 1) * 4 + (lane >> 3);
    }
    const int pxi = tile_x * TILE + lx, pyi = tile_y * TILE + ly;
    const bool inside = lx < TILE && ly < TILE && pxi < a.W && pyi < a.H;
    const float pfx = (float)pxi, pfy = (float)pyi;
    const size_t HW = (size_t)a.W * a.H;
    const size_t pix = inside ? (size_t)pyi * a.W + pxi : 0;

    uint2 rg = vw.ranges[blockIdx.x];
    if (vw.info->overflow) rg = make_uint2(0u, 0u);
    // the list is walked as far as ANY pixel of the tile got (also the pixels the packed variant does not carry:
    // the reference's block visits those entries and advances the language recurrence of every pixel, Q2)
    uint32_t last_contributor = inside ? vw.n_contrib[pix] : 0u;
    uint32_t warp_maxc = __reduce_max_sync(0xffffffffu, last_contributor);
    if (PACKED) {
        uint32_t other = 0;
        for (int r = tid; r < TILE * TILE; r += NT) {
            const int ox = tile_x * TILE + r % TILE, oy = tile_y * TILE + r / TILE;
            if (ox < a.W && oy < a.H) other = max(other, vw.n_contrib[(size_t)oy * a.W + ox]);
        }
        warp_maxc = max(warp_maxc, __reduce_max_sync(0xffffffffu, other));
    }
    if (tid == 0) s_maxc = 0;
    for (int e = tid; e < BWD_BATCH * GR; e += NT) s_acc[e] = 0.0f;
    __syncthreads();
    if (lane == 0 && warp_maxc) atomicMax(&s_maxc, warp_maxc);
    __syncthreads();
    // entries at positions >= max n_contrib are skipped by every pixel of the tile in both modes
    const int total = min((int)s_maxc, (int)(rg.y - rg.x));
    if (total == 0) return;

    const float T_final = inside ? vw.final_T[pix] : 0.0f;
    float T = T_final;
    float g[NCH > 0 ? NCH + 1 : 1];  // dL/dpixel per channel (+1: the zero partner of an odd last channel)
    float gd = 0.0f;
#pragma unroll
    for (int c = 0; c < NCH + 1; c++) g[c] = 0.0f;
    if (inside) {
#pragma unroll
        for (int c = 0; c < NCOL; c++) g[c] = vw.dL_dcolor[c * HW + pix];
#pragma unroll
        for (int c = 0; c < F; c++) g[NCOL + c] = vw.dL_dlanguage[c * HW + pix];
        if (NCOL) gd = vw.dL_ddepth[pix];
    }
    f32x2 g2[NPAIR > LP0 ? NPAIR - LP0 : 1];  // language-only pairs of g, packed for FFMA2
#pragma unroll
    for (int p = LP0; p < NPAIR; p++) asm("mov.b64 %0, {%1, %2};" : "=l"(g2[p - LP0]) : "f"(g[2 * p]), "f"(g[2 * p + 1]));
    float bg_dot = 0.0f;
    if (NCOL) bg_dot = vw.bg[0] * g[0] + vw.bg[1] * g[1] + vw.bg[2] * g[2];
    const float ddelx_dx = 0.5f * a.W, ddely_dy = 0.5f * a.H;
    // Q3: the lane mask is indexed by the reference's thread rank ly * TILE + lx
    const int ref_rank = ly * TILE + lx;
    const bool lane_ok = PACKED ? inside : (COMPAT ? (inside && ((a.lane_ok[(ref_rank >> 5) & 7] >> (ref_rank & 31)) & 1u) != 0) : true);
    const int own = multi_reduce_owner<NV>(lane);
    const int own_dst = own < 0 ? -1 : (NCOL ? own : (own < NGEO ? GR_CX + own : GR_LANG + (own - NGEO)));

    float last_alpha = 0.0f;
    float A_c = 0.0f, Dl_c = 0.0f;  // rgb + depth part of sum_ch accum_rec*g and of last_color*g
    float A_f = 0.0f, Dl_f = 0.0f;  // language part (separate because of Q2 in compat mode)
    // Q2 bookkeeping (compat): until a pixel of this warp has blended anything, last_alpha is 0 for all of
    // its lanes and the unguarded recurrence only replaces Dl_f, so only the latest visited entry matters.
    bool fresh = true;

    // language part of sum_ch c_ch * dL/dpix_ch for the record at rj
    auto lang_dot = [&](const float* rj) -> float {
        float d = 0.0f;
        if (F > 0) {
            if (NCOL & 1) d = rj[REC_CH + NCOL] * g[NCOL];  // the language channel sharing a pair with blue
            f32x2 acc = 0ull;
#pragma unroll
            for (int p = LP0; p < NPAIR; p++)
                acc = ffma2(*reinterpret_cast<const f32x2*>(rj + REC_CH + 2 * p), g2[p - LP0], acc);
            d += hsum2(acc);
        }
        return d;
    };

    const int n_batches = (total + BWD_BATCH - 1) / BWD_BATCH;
    for (int b = n_batches - 1; b >= 0; b--) {
        const int base = b * BWD_BATCH;
        const int cnt = min(BWD_BATCH, total - base);
        __syncthreads();  // previous batch fully consumed / flushed
        {   // BWD_THREADS / BWD_BATCH threads share one record: one list lookup each, pieces dealt round-robin
            constexpr int TPE = NT / BWD_BATCH;
            const int gi = tid / TPE;
            if (gi < cnt) {
                const uint32_t id = vw.point_list[rg.x + base + gi];
                if ((tid % TPE) == 0) s_id[gi] = id;
#pragma unroll
                for (int q = tid % TPE; q < OPS; q += TPE) Stage::copy(&s_rec[gi * REC], vw.records, a.language, id, q);
            }
        }
        cp_async_commit();
        const uint32_t hit = lane < cnt ? (uint32_t)vw.warp_hits[rg.x + base + lane] : 0u;
        // entries a pixel of this warp blended (packed variant: not known in advance, decided by a vote below)
        // (packed: entries some forward block overlapping this warp's pixels blended -- a superset, the vote decides)
        const uint32_t mine = PACKED ? __ballot_sync(0xffffffffu, (hit & a.packed_fmask[wid & 3]) != 0u)
                                     : __ballot_sync(0xffffffffu, (hit >> wid) & 1u);
        uint32_t visit = COMPAT ? __ballot_sync(0xffffffffu, hit != 0u) : mine;
        cp_async_wait<0>();
        __syncthreads();
        int pend = -1;  // compat: visited entry whose Dl_f update is still owed (warp still fresh)

        while (visit) {
            const int j = 31 - __clz(visit);  // back to front
            visit &= ~(1u << j);
            const float* rj = s_rec + j * REC;
            float4 g0, g1;
            float dx = 0.0f, dy = 0.0f, G = 0.0f, alpha = 0.0f;
            bool contrib = false;
            auto evaluate = [&]() {
                g0 = *reinterpret_cast<const float4*>(rj);      // x y A B
                g1 = *reinterpret_cast<const float4*>(rj + 4);  // C op pth depth
                dx = fsub(g0.x, pfx);
                dy = fsub(g0.y, pfy);
                const float power = ffma(ffma(dx, fmul(dx, g0.z), fmul(dy, fmul(dy, g1.x))), -0.5f, -fmul(dy, fmul(dx, g0.w)));
                if (inside && (uint32_t)(base + j) < last_contributor && !(power > 0.0f) && !(power < g1.z)) {
                    G = a.fast_exp ? fast_exp2(power) : expf(power);
                    alpha = fminf(0.99f, fmul(g1.y, G));
                    contrib = !(alpha < 1.0f / 255.0f);
                }
            };
            bool warp_blends = (mine >> j) & 1u;
            if (PACKED && warp_blends) {
                evaluate();
                warp_blends = __any_sync(0xffffffffu, contrib);
            }
            if (!warp_blends) {  // compat only: another pixel block of the tile blends this entry
                if (F > 0) {
                    if (fresh) { pend = j; continue; }
                    if (inside) {  // Q2: the language recurrence advances for every pixel of a visited Gaussian
                        A_f = last_alpha * Dl_f + (1.0f - last_alpha) * A_f;
                        Dl_f = lang_dot(rj);
                    }
                }
                continue;
            }
            if (COMPAT && F > 0) {
                if (pend >= 0) { Dl_f = lang_dot(s_rec + pend * REC); pend = -1; }
                fresh = false;
            }
            if (!PACKED) evaluate();
            float v[NV];
#pragma unroll
            for (int i = 0; i < NV; i++) v[i] = 0.0f;
            float D_f = 0.0f;
            if (inside && ((COMPAT && F > 0) || contrib)) {
                D_f = lang_dot(rj);
                if (COMPAT && F > 0) {
                    A_f = last_alpha * Dl_f + (1.0f - last_alpha) * A_f;
                    Dl_f = D_f;
                }
            }
            float w = 0.0f;
            if (contrib) {
                // 1 - alpha >= 0.01: the approximate reciprocal (2 ulp) is far inside the gradient tolerance and
                // is shared by the transmittance update and the background term
                const float inv_1ma = __fdividef(1.0f, 1.0f - alpha);
                T = T * inv_1ma;
                w = alpha * T;
                float D_c = 0.0f;
                if (NCOL) {
                    D_c = rj[REC_CH] * g[0] + rj[REC_CH + 1] * g[1] + rj[REC_CH + 2] * g[2] + g1.w * gd;
                    A_c = last_alpha * Dl_c + (1.0f - last_alpha) * A_c;
                    Dl_c = D_c;
                }
                if (!COMPAT && F > 0) {
                    A_f = last_alpha * Dl_f + (1.0f - last_alpha) * A_f;
                    Dl_f = D_f;
                }
                float dL_dalpha = ((D_c - A_c) + (D_f - A_f)) * T;
                last_alpha = alpha;
                if (NCOL) dL_dalpha += (-T_final * inv_1ma) * bg_dot;
                const float dL_dG = g1.y * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                if (lane_ok) {
                    constexpr int C0 = NCOL ? GR_CX : 0;  // position of conic.x in v[]
                    if (NCOL) {  // D/ drops the mean gradient of the language footprint (D/backward.cu:1074,1117)
                        const float dG_ddelx = -gdx * g0.z - gdy * g0.w;
                        const float dG_ddely = -gdy * g1.x - gdx * g0.w;
                        v[GR_MX] = dL_dG * dG_ddelx * ddelx_dx;
                        v[GR_MY] = dL_dG * dG_ddely * ddely_dy;
                        v[GR_DEPTH] = w * gd;
                        v[GR_RGB + 0] = w * g[0];
                        v[GR_RGB + 1] = w * g[1];
                        v[GR_RGB + 2] = w * g[2];
                    }
                    v[C0 + 0] = -0.5f * gdx * dx * dL_dG;
                    v[C0 + 1] = -0.5f * gdx * dy * dL_dG;
                    v[C0 + 2] = -0.5f * gdy * dy * dL_dG;
                    v[C0 + 3] = G * dL_dalpha;
                    if (!COMPAT) {
#pragma unroll
                        for (int c = 0; c < F; c++) v[NGEO + c] = w * g[NCOL + c];
                    }
                }
            }
            // Q1 (compat): only the tile's first thread contributes its own pixel's language gradient
            if (COMPAT && F > 0 && tid == 0 && contrib) {
#pragma unroll
                for (int c = 0; c < F; c++) s_acc[j * GR + GR_LANG + c] += w * g[NCOL + c];
            }
            if (__any_sync(0xffffffffu, contrib && lane_ok)) {
                warp_multi_reduce<NV>(v, lane);
                if (own_dst >= 0 && v[0] != 0.0f) atomicAdd(&s_acc[j * GR + own_dst], v[0]);
            }
        }
        if (COMPAT && F > 0 && pend >= 0) Dl_f = lang_dot(s_rec + pend * REC);
        __syncthreads();
        // flush the batch: one global atomic per (Gaussian, value) that received something
        for (int e = tid; e < cnt * GR; e += NT) {
            const float val = s_acc[e];
            if (val != 0.0f) {
                const int gi = e / GR, vi = e - gi * GR;
                atomicAdd(&vw.gacc[(size_t)s_id[gi] * GR + vi], val);
                vw.gtouched[s_id[gi]] = 1;   // idempotent byte store: the geometry pass skips untouched records unread
                s_acc[e] = 0.0f;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Backward blend, two pixels per lane (the default).  Same arithmetic, same gradient modes and the same per-entry
// bookkeeping as k_blend_bwd, but a warp carries 64 pixels, so everything that is paid per (warp, entry) -- the walk
// over the list, the record loads from shared memory, the warp-wide butterfly reduction of the 10 (compat) or 25 (exact)
// per-Gaussian sums and the shared-memory atomics -- is paid once per TWO pixels, and every lane runs two independent
// dependency chains.  Thread -> pixel mapping:
//   unpacked  128 threads = 4 warps of 8x8 pixels (lane l: column l & 7, rows l >> 3 and (l >> 3) + 4), the forward's
//             layout, so a warp's entries are those its two 8x4 halves blended (bits 2*(w>>1)*2 + (w&1) and +2);
//   PACKED    (compat, 15x15) 64 threads = 2 warps; lane l of warp w carries the surviving pixels with packed index
//             64 w + l and 64 w + 32 + l (reference rank order); a warp decides by a vote after evaluating an entry.
// ---------------------------------------------------------------------------------------------------
template <int TILE, int NCOL, int F, bool COMPAT, bool PACKED>
__global__ void __launch_bounds__(PACKED ? 64 : 128, PACKED ? 10 : 5) k_blend_bwd2(const __grid_constant__ BwdBlendArgs a) {
    const BwdView& vw = a.v[blockIdx.y];
    static_assert(TILE <= 16 && BWD_BATCH == 32, "4 warps of 8x8 pixels; one ballot per batch");
    static_assert(!PACKED || COMPAT, "packing follows the compat lane mask");
    constexpr int NT = PACKED ? 64 : 128;
    static_assert(NCOL == 0 || NCOL == 3, "colour channels");
    constexpr int NCH = NCOL + F;
    constexpr int REC = rec_floats_nch(NCH);
    using Stage = RecordStage<NCOL, F>;
    constexpr int OPS = Stage::OPS;
    constexpr int GR = grad_floats(F);
    constexpr int NPAIR = (NCH + 1) / 2;
    constexpr int LP0 = (NCOL + 1) / 2;
    constexpr int NLP = NPAIR > LP0 ? NPAIR - LP0 : 1;
    constexpr int NGEO = NCOL ? 10 : 4;
    constexpr int NV = COMPAT ? NGEO : NGEO + F;
    __shared__ __align__(16) float s_rec[BWD_BATCH * REC];
    __shared__ uint32_t s_id[BWD_BATCH];
    __shared__ float s_acc[BWD_BATCH * GR];
    __shared__ uint32_t s_maxc;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile_x = blockIdx.x % a.gx, tile_y = blockIdx.x / a.gx;
    int lx[2], ly[2];
#pragma unroll
    for (int p = 0; p < 2; p++) {
        if (PACKED) {
            const int rank = a.packed_rank[wid * 64 + 32 * p + lane];
            lx[p] = rank == 255 ? TILE : rank % TILE;
            ly[p] = rank == 255 ? TILE : rank / TILE;
        } else {
            lx[p] = (wid & 1) * 8 + (lane & 7);
            ly[p] = (wid >> 1) * 8 + (lane >> 3) + 4 * p;
        }
    }
    bool inside[2];
    float pfx[2], pfy[2];
    size_t pix[2];
    const size_t HW = (size_t)a.W * a.H;
#pragma unroll
    for (int p = 0; p < 2; p++) {
        const int pxi = tile_x * TILE + lx[p], pyi = tile_y * TILE + ly[p];
        inside[p] = lx[p] < TILE && ly[p] < TILE && pxi < a.W && pyi < a.H;
        pfx[p] = (float)pxi; pfy[p] = (float)pyi;
        pix[p] = inside[p] ? (size_t)pyi * a.W + pxi : 0;
    }

    uint2 rg = vw.ranges[blockIdx.x];
    if (vw.info->overflow) rg = make_uint2(0u, 0u);
    uint32_t last_contributor[2];
#pragma unroll
    for (int p = 0; p < 2; p++) last_contributor[p] = inside[p] ? vw.n_contrib[pix[p]] : 0u;
    uint32_t warp_maxc = __reduce_max_sync(0xffffffffu, max(last_contributor[0], last_contributor[1]));
    if (PACKED) {
        uint32_t other = 0;
        for (int r = tid; r < TILE * TILE; r += NT) {
            const int ox = tile_x * TILE + r % TILE, oy = tile_y * TILE + r / TILE;
            if (ox < a.W && oy < a.H) other = max(other, vw.n_contrib[(size_t)oy * a.W + ox]);
        }
        warp_maxc = max(warp_maxc, __reduce_max_sync(0xffffffffu, other));
    }
    if (tid == 0) s_maxc = 0;
    for (int e = tid; e < BWD_BATCH * GR; e += NT) s_acc[e] = 0.0f;
    __syncthreads();
    if (lane == 0 && warp_maxc) atomicMax(&s_maxc, warp_maxc);
    __syncthreads();
    const int total = min((int)s_maxc, (int)(rg.y - rg.x));
    if (total == 0) return;

    float T_final[2], T[2], g[2][NCH > 0 ? NCH + 1 : 1], gd[2], bg_dot[2];
    f32x2 g2[2][NLP];
    bool lane_ok[2];
#pragma unroll
    for (int p = 0; p < 2; p++) {
        T_final[p] = inside[p] ? vw.final_T[pix[p]] : 0.0f;
        T[p] = T_final[p];
        gd[p] = 0.0f;
#pragma unroll
        for (int c = 0; c < NCH + 1; c++) g[p][c] = 0.0f;
        if (inside[p]) {
#pragma unroll
            for (int c = 0; c < NCOL; c++) g[p][c] = vw.dL_dcolor[c * HW + pix[p]];
#pragma unroll
            for (int c = 0; c < F; c++) g[p][NCOL + c] = vw.dL_dlanguage[c * HW + pix[p]];
            if (NCOL) gd[p] = vw.dL_ddepth[pix[p]];
        }
#pragma unroll
        for (int q = LP0; q < NPAIR; q++)
            asm("mov.b64 %0, {%1, %2};" : "=l"(g2[p][q - LP0]) : "f"(g[p][2 * q]), "f"(g[p][2 * q + 1]));
        bg_dot[p] = NCOL ? vw.bg[0] * g[p][0] + vw.bg[1] * g[p][1] + vw.bg[2] * g[p][2] : 0.0f;
        const int ref_rank = ly[p] * TILE + lx[p];
        lane_ok[p] = PACKED ? inside[p]
                            : (COMPAT ? (inside[p] && ((a.lane_ok[(ref_rank >> 5) & 7] >> (ref_rank & 31)) & 1u) != 0) : true);
    }
    const float ddelx_dx = 0.5f * a.W, ddely_dy = 0.5f * a.H;
    const int own = multi_reduce_owner<NV>(lane);
    const int own_dst = own < 0 ? -1 : (NCOL ? own : (own < NGEO ? GR_CX + own : GR_LANG + (own - NGEO)));
    // forward 8x4 blocks whose pixels this warp carries
    uint32_t my_blocks;
    if (PACKED) my_blocks = (uint32_t)a.packed_fmask[(2 * wid) & 3] | (uint32_t)a.packed_fmask[(2 * wid + 1) & 3];
    else { const int b0 = ((wid >> 1) * 2) * 2 + (wid & 1); my_blocks = (1u << b0) | (1u << (b0 + 2)); }

    float last_alpha[2] = {0.0f, 0.0f};
    float A_c[2] = {0.0f, 0.0f}, Dl_c[2] = {0.0f, 0.0f}, A_f[2] = {0.0f, 0.0f}, Dl_f[2] = {0.0f, 0.0f};
    bool fresh = true;  // Q2 bookkeeping (compat): no pixel of this warp has blended anything yet

    // language part of sum_ch c_ch * dL/dpix_ch of the record at rj, for both pixels (the record is loaded once)
    auto lang_dot2 = [&](const float* rj, float& d0, float& d1) {
        d0 = d1 = 0.0f;
        if (F > 0) {
            if (NCOL & 1) { const float c = rj[REC_CH + NCOL]; d0 = c * g[0][NCOL]; d1 = c * g[1][NCOL]; }
            f32x2 acc0 = 0ull, acc1 = 0ull;
#pragma unroll
            for (int q = LP0; q < NPAIR; q++) {
                const f32x2 c2 = *reinterpret_cast<const f32x2*>(rj + REC_CH + 2 * q);
                acc0 = ffma2(c2, g2[0][q - LP0], acc0);
                acc1 = ffma2(c2, g2[1][q - LP0], acc1);
            }
            d0 += hsum2(acc0);
            d1 += hsum2(acc1);
        }
    };

    const int n_batches = (total + BWD_BATCH - 1) / BWD_BATCH;
    for (int b = n_batches - 1; b >= 0; b--) {
        const int base = b * BWD_BATCH;
        const int cnt = min(BWD_BATCH, total - base);
        __syncthreads();  // previous batch fully consumed / flushed
        {
            constexpr int TPE = NT / BWD_BATCH;
            const int gi = tid / TPE;
            if (gi < cnt) {
                const uint32_t id = vw.point_list[rg.x + base + gi];
                if ((tid % TPE) == 0) s_id[gi] = id;
#pragma unroll
                for (int q = tid % TPE; q < OPS; q += TPE) Stage::copy(&s_rec[gi * REC], vw.records, a.language, id, q);
            }
        }
        cp_async_commit();
        const uint32_t hit = lane < cnt ? (uint32_t)vw.warp_hits[rg.x + base + lane] : 0u;
        const uint32_t mine = __ballot_sync(0xffffffffu, (hit & my_blocks) != 0u);
        uint32_t visit = COMPAT ? __ballot_sync(0xffffffffu, hit != 0u) : mine;
        cp_async_wait<0>();
        __syncthreads();
        int pend = -1;

        while (visit) {
            const int j = 31 - __clz(visit);  // back to front
            visit &= ~(1u << j);
            const float* rj = s_rec + j * REC;
            float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
            float dx[2] = {0.0f, 0.0f}, dy[2] = {0.0f, 0.0f}, G[2] = {0.0f, 0.0f}, alpha[2] = {0.0f, 0.0f};
            bool contrib[2] = {false, false};
            auto evaluate = [&]() {
                g0 = *reinterpret_cast<const float4*>(rj);      // x y A B
                g1 = *reinterpret_cast<const float4*>(rj + 4);  // C op pth depth
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    dx[p] = fsub(g0.x, pfx[p]);
                    dy[p] = fsub(g0.y, pfy[p]);
                    const float power = ffma(ffma(dx[p], fmul(dx[p], g0.z), fmul(dy[p], fmul(dy[p], g1.x))), -0.5f,
                                             -fmul(dy[p], fmul(dx[p], g0.w)));
                    if (inside[p] && (uint32_t)(base + j) < last_contributor[p] && !(power > 0.0f) && !(power < g1.z)) {
                        G[p] = a.fast_exp ? fast_exp2(power) : expf(power);
                        alpha[p] = fminf(0.99f, fmul(g1.y, G[p]));
                        contrib[p] = !(alpha[p] < 1.0f / 255.0f);
                    }
                }
            };
            bool warp_blends = (mine >> j) & 1u;
            if (PACKED && warp_blends) {
                evaluate();
                warp_blends = __any_sync(0xffffffffu, contrib[0] || contrib[1]);
            }
            if (!warp_blends) {  // compat only: another pixel block of the tile blends this entry
                if (F > 0) {
                    if (fresh) { pend = j; continue; }
                    float d0, d1;
                    lang_dot2(rj, d0, d1);
                    if (inside[0]) { A_f[0] = last_alpha[0] * Dl_f[0] + (1.0f - last_alpha[0]) * A_f[0]; Dl_f[0] = d0; }
                    if (inside[1]) { A_f[1] = last_alpha[1] * Dl_f[1] + (1.0f - last_alpha[1]) * A_f[1]; Dl_f[1] = d1; }
                }
                continue;
            }
            if (COMPAT && F > 0) {
                if (pend >= 0) {
                    float d0, d1;
                    lang_dot2(s_rec + pend * REC, d0, d1);
                    Dl_f[0] = d0; Dl_f[1] = d1;
                    pend = -1;
                }
                fresh = false;
            }
            if (!PACKED) evaluate();
            float v[NV];
#pragma unroll
            for (int i = 0; i < NV; i++) v[i] = 0.0f;
            float D_f[2] = {0.0f, 0.0f};
            if ((COMPAT && F > 0) || contrib[0] || contrib[1]) {
                float d0, d1;
                lang_dot2(rj, d0, d1);
                const float dd[2] = {d0, d1};
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    if (inside[p] && ((COMPAT && F > 0) || contrib[p])) {
                        D_f[p] = dd[p];
                        if (COMPAT && F > 0) {
                            A_f[p] = last_alpha[p] * Dl_f[p] + (1.0f - last_alpha[p]) * A_f[p];
                            Dl_f[p] = dd[p];
                        }
                    }
                }
            }
            float w[2] = {0.0f, 0.0f};
#pragma unroll
            for (int p = 0; p < 2; p++) {
                if (contrib[p]) {
                    const float inv_1ma = __fdividef(1.0f, 1.0f - alpha[p]);
                    T[p] = T[p] * inv_1ma;
                    w[p] = alpha[p] * T[p];
                    float D_c = 0.0f;
                    if (NCOL) {
                        D_c = rj[REC_CH] * g[p][0] + rj[REC_CH + 1] * g[p][1] + rj[REC_CH + 2] * g[p][2] + g1.w * gd[p];
                        A_c[p] = last_alpha[p] * Dl_c[p] + (1.0f - last_alpha[p]) * A_c[p];
                        Dl_c[p] = D_c;
                    }
                    if (!COMPAT && F > 0) {
                        A_f[p] = last_alpha[p] * Dl_f[p] + (1.0f - last_alpha[p]) * A_f[p];
                        Dl_f[p] = D_f[p];
                    }
                    float dL_dalpha = ((D_c - A_c[p]) + (D_f[p] - A_f[p])) * T[p];
                    last_alpha[p] = alpha[p];
                    if (NCOL) dL_dalpha += (-T_final[p] * inv_1ma) * bg_dot[p];
                    const float dL_dG = g1.y * dL_dalpha;
                    const float gdx = G[p] * dx[p], gdy = G[p] * dy[p];
                    if (lane_ok[p]) {
                        constexpr int C0 = NCOL ? GR_CX : 0;
                        if (NCOL) {
                            const float dG_ddelx = -gdx * g0.z - gdy * g0.w;
                            const float dG_ddely = -gdy * g1.x - gdx * g0.w;
                            v[GR_MX] += dL_dG * dG_ddelx * ddelx_dx;
                            v[GR_MY] += dL_dG * dG_ddely * ddely_dy;
                            v[GR_DEPTH] += w[p] * gd[p];
                            v[GR_RGB + 0] += w[p] * g[p][0];
                            v[GR_RGB + 1] += w[p] * g[p][1];
                            v[GR_RGB + 2] += w[p] * g[p][2];
                        }
                        v[C0 + 0] += -0.5f * gdx * dx[p] * dL_dG;
                        v[C0 + 1] += -0.5f * gdx * dy[p] * dL_dG;
                        v[C0 + 2] += -0.5f * gdy * dy[p] * dL_dG;
                        v[C0 + 3] += G[p] * dL_dalpha;
                        if (!COMPAT) {
#pragma unroll
                            for (int c = 0; c < F; c++) v[NGEO + c] += w[p] * g[p][NCOL + c];
                        }
                    }
                }
            }
            // Q1 (compat): only the tile's first thread contributes its own (first) pixel's language gradient
            if (COMPAT && F > 0 && tid == 0 && contrib[0]) {
#pragma unroll
                for (int c = 0; c < F; c++) s_acc[j * GR + GR_LANG + c] += w[0] * g[0][NCOL + c];
            }
            if (__any_sync(0xffffffffu, (contrib[0] && lane_ok[0]) || (contrib[1] && lane_ok[1]))) {
                warp_multi_reduce<NV>(v, lane);
                if (own_dst >= 0 && v[0] != 0.0f) atomicAdd(&s_acc[j * GR + own_dst], v[0]);
            }
        }
        if (COMPAT && F > 0 && pend >= 0) {
            float d0, d1;
            lang_dot2(s_rec + pend * REC, d0, d1);
            Dl_f[0] = d0; Dl_f[1] = d1;
        }
        __syncthreads();
        for (int e = tid; e < cnt * GR; e += NT) {
            const float val = s_acc[e];
            if (val != 0.0f) {
                const int gi = e / GR, vi = e - gi * GR;
                atomicAdd(&vw.gacc[(size_t)s_id[gi] * GR + vi], val);
                vw.gtouched[s_id[gi]] = 1;   // idempotent byte store: the geometry pass skips untouched records unread
                s_acc[e] = 0.0f;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
struct GeomView {  // camera + per-view buffers of one view
    const float *viewmatrix, *projmatrix, *projmatrix_raw, *campos;
    const uint32_t* clamped;
    const float* gacc;
    const uint8_t* gtouched;  // [P] or NULL: 0 = the gradient record is all zero (not even read)
    const int32_t* radii;
    float* dL_dmeans2D;   // [P,3] this view's screen-space gradient (viewspace_points.grad)
    float* dL_dtau;       // [P,6] per-Gaussian pose gradient as the reference returns it, or NULL
    float* dL_dtau_sum;   // [6]   pose gradient summed over the Gaussians (zeroed by the launcher), or NULL
    float tanfovx, tanfovy, focal_x, focal_y;
};

struct GeomCommon {
    int P, F, sh_degree, M, W, H, gr;
    float scale_modifier;
    const float *means3D, *shs;
    bool colors_precomp;
    bool accumulate;  // += into the parameter gradients (means3D, sh, opacity, scales, rotations, language, cov3D)
    float* dL_dsh;
};

struct GeomBwdArgs : GeomCommon {
    int V;
    const float *scales, *rotations, *cov3D;
    float *dL_dcolors, *dL_dlanguage, *dL_dopacity, *dL_dmeans3D, *dL_dcov3D, *dL_dscales, *dL_drots;
    float *stat_max_radii, *stat_accum, *stat_denom;  // optional densification statistics (all or none)
    GeomView v[OLS_MAX_VIEWS];
};

__constant__ float B_SH_C0 = 0.28209479177387814f;
__constant__ float B_SH_C1 = 0.4886025119029199f;
__constant__ float B_SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                 -1.0925484305920792f, 0.5462742152960396f};
__constant__ float B_SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                                 -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

// computeCov2DCUDA (backward.cu:150-346): conic gradient -> dL/dcov3D (written), dL/dmean3D (assigned) and the
// pose gradient (accumulated).  D/'s computeCov2DCUDA_no_tau (D/backward.cu:354-446) is the same arithmetic
// keeping only dL/dcov3D: callers simply drop the other two results.
__device__ __forceinline__ void geom_cov2d_bwd(const GeomView& a, const float* V, const float* mp, const float* c3,
                                       float dcx, float dcy, float dcz, float* dcov, float* dmean, float* dtau) {
    const float fx = a.focal_x, fy = a.focal_y;
    // ---- computeCov2DCUDA (backward.cu:150-346)
    float t[3] = {V[0] * mp[0] + V[4] * mp[1] + V[8] * mp[2] + V[12], V[1] * mp[0] + V[5] * mp[1] + V[9] * mp[2] + V[13],
                  V[2] * mp[0] + V[6] * mp[1] + V[10] * mp[2] + V[14]};
    const float limx = 1.3f * a.tanfovx, limy = 1.3f * a.tanfovy;
    const float txtz = t[0] / t[2], tytz = t[1] / t[2];
    t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
    t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
    const float xgm = (txtz < -limx || txtz > limx) ? 0.0f : 1.0f;
    const float ygm = (tytz < -limy || tytz > limy) ? 0.0f : 1.0f;
    // column-major like GLM: X[c][r]
    const float J[3][3] = {{fx / t[2], 0, -(fx * t[0]) / (t[2] * t[2])}, {0, fy / t[2], -(fy * t[1]) / (t[2] * t[2])}, {0, 0, 0}};
    const float Wm[3][3] = {{V[0], V[4], V[8]}, {V[1], V[5], V[9]}, {V[2], V[6], V[10]}};
    const float Vrk[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    float Tm[3][3];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++) Tm[c][r] = Wm[0][r] * J[c][0] + Wm[1][r] * J[c][1] + Wm[2][r] * J[c][2];
    float TV[2][3];  // TV[i][q] = sum_p Tm[i][p] * Vrk[p][q]
#pragma unroll
    for (int ii = 0; ii < 2; ii++)
#pragma unroll
        for (int q = 0; q < 3; q++) TV[ii][q] = Tm[ii][0] * Vrk[0][q] + Tm[ii][1] * Vrk[1][q] + Tm[ii][2] * Vrk[2][q];
    const float ca = TV[0][0] * Tm[0][0] + TV[0][1] * Tm[0][1] + TV[0][2] * Tm[0][2] + 0.3f;
    const float cb = TV[0][0] * Tm[1][0] + TV[0][1] * Tm[1][1] + TV[0][2] * Tm[1][2];
    const float cc = TV[1][0] * Tm[1][0] + TV[1][1] * Tm[1][1] + TV[1][2] * Tm[1][2] + 0.3f;
    const float denom = ca * cc - cb * cb;
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    if (denom2inv != 0) {
        dL_da = denom2inv * (-cc * cc * dcx + 2 * cb * cc * dcy + (denom - ca * cc) * dcz);
        dL_dc = denom2inv * (-ca * ca * dcz + 2 * ca * cb * dcy + (denom - ca * cc) * dcx);
        dL_db = denom2inv * 2 * (cb * cc * dcx - (denom + 2 * cb * cb) * dcy + ca * cb * dcz);
        dcov[0] = (Tm[0][0] * Tm[0][0] * dL_da + Tm[0][0] * Tm[1][0] * dL_db + Tm[1][0] * Tm[1][0] * dL_dc);
        dcov[3] = (Tm[0][1] * Tm[0][1] * dL_da + Tm[0][1] * Tm[1][1] * dL_db + Tm[1][1] * Tm[1][1] * dL_dc);
        dcov[5] = (Tm[0][2] * Tm[0][2] * dL_da + Tm[0][2] * Tm[1][2] * dL_db + Tm[1][2] * Tm[1][2] * dL_dc);
        dcov[1] = 2 * Tm[0][0] * Tm[0][1] * dL_da + (Tm[0][0] * Tm[1][1] + Tm[0][1] * Tm[1][0]) * dL_db + 2 * Tm[1][0] * Tm[1][1] * dL_dc;
        dcov[2] = 2 * Tm[0][0] * Tm[0][2] * dL_da + (Tm[0][0] * Tm[1][2] + Tm[0][2] * Tm[1][0]) * dL_db + 2 * Tm[1][0] * Tm[1][2] * dL_dc;
        dcov[4] = 2 * Tm[0][2] * Tm[0][1] * dL_da + (Tm[0][1] * Tm[1][2] + Tm[0][2] * Tm[1][1]) * dL_db + 2 * Tm[1][1] * Tm[1][2] * dL_dc;
    }
    // TV[r][k] equals the reference's (T[r] . Vrk[k]) because Vrk is symmetric
    const float dT00 = 2 * TV[0][0] * dL_da + TV[1][0] * dL_db, dT01 = 2 * TV[0][1] * dL_da + TV[1][1] * dL_db,
                dT02 = 2 * TV[0][2] * dL_da + TV[1][2] * dL_db;
    const float dT10 = 2 * TV[1][0] * dL_dc + TV[0][0] * dL_db, dT11 = 2 * TV[1][1] * dL_dc + TV[0][1] * dL_db,
                dT12 = 2 * TV[1][2] * dL_dc + TV[0][2] * dL_db;
    const float dJ00 = Wm[0][0] * dT00 + Wm[0][1] * dT01 + Wm[0][2] * dT02;
    const float dJ02 = Wm[2][0] * dT00 + Wm[2][1] * dT01 + Wm[2][2] * dT02;
    const float dJ11 = Wm[1][0] * dT10 + Wm[1][1] * dT11 + Wm[1][2] * dT12;
    const float dJ12 = Wm[2][0] * dT10 + Wm[2][1] * dT11 + Wm[2][2] * dT12;
    const float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    const float dtx = xgm * -fx * tz2 * dJ02;
    const float dty = ygm * -fy * tz2 * dJ12;
    const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * t[0]) * tz3 * dJ02 + (2 * fy * t[1]) * tz3 * dJ12;
    {   // pose: dpC/drho = I, dpC/dtheta = -skew(t)
        const float th[3][3] = {{0, -t[2], t[1]}, {t[2], 0, -t[0]}, {-t[1], t[0], 0}};
        const float d3[3] = {dtx, dty, dtz};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            dtau[k] += d3[k];
            dtau[k + 3] += dtx * th[k][0] + dty * th[k][1] + dtz * th[k][2];
        }
    }
    dmean[0] = V[0] * dtx + V[1] * dty + V[2] * dtz;
    dmean[1] = V[4] * dtx + V[5] * dty + V[6] * dtz;
    dmean[2] = V[8] * dtx + V[9] * dty + V[10] * dtz;
    {   // dL/dW through T = W * J, folded onto the rotation's so(3) tangent
        const float dW00 = J[0][0] * dT00, dW01 = J[0][0] * dT01, dW02 = J[0][0] * dT02;
        const float dW10 = J[1][1] * dT10, dW11 = J[1][1] * dT11, dW12 = J[1][1] * dT12;
        const float dW20 = J[0][2] * dT00 + J[1][2] * dT10, dW21 = J[0][2] * dT01 + J[1][2] * dT11,
                    dW22 = J[0][2] * dT02 + J[1][2] * dT12;
        const float c1[3] = {V[0], V[1], V[2]}, c2[3] = {V[4], V[5], V[6]}, c3_[3] = {V[8], V[9], V[10]};
        const float w1[3] = {dW00, dW10, dW20}, w2[3] = {dW01, dW11, dW21}, w3[3] = {dW02, dW12, dW22};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float acc = 0.0f;
            const float* cs[3] = {c1, c2, c3_};
            const float* ws[3] = {w1, w2, w3};
#pragma unroll
            for (int m = 0; m < 3; m++) {
                const float* vv = cs[m];
                const float S[3][3] = {{0, -vv[2], vv[1]}, {vv[2], 0, -vv[0]}, {-vv[1], vv[0], 0}};
                acc += ws[m][0] * S[k][0] + ws[m][1] * S[k][1] + ws[m][2] * S[k][2];
            }
            dtau[3 + k] += acc;
        }
    }
}

// language_preprocessCUDA (backward.cu:541-682): projection and depth paths of the mean / pose gradient
__device__ __forceinline__ void geom_proj_bwd(const float* V, const float* Pm, const float* Praw, const float* mp,
                                      float g2x, float g2y, float dzv, float* dmean, float* dtau) {
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
        const float th[3][3] = {{0, -pC[2], pC[1]}, {pC[2], 0, -pC[0]}, {-pC[1], pC[0], 0}};
        const float dz = dzv;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            dtau[k] += g2x * d1[k] + g2y * d2[k];
            const float t1 = th[k][0] * d1[0] + th[k][1] * d1[1] + th[k][2] * d1[2];
            const float t2 = th[k][0] * d2[0] + th[k][1] * d2[1] + th[k][2] * d2[2];
            dtau[3 + k] += g2x * t1 + g2y * t2;
        }
        dmean[0] += dz * V[2]; dmean[1] += dz * V[6]; dmean[2] += dz * V[10];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            dtau[k] += dz * (k == 2 ? 1.0f : 0.0f);
            dtau[3 + k] += dz * th[k][2];
        }
    }
}

// computeColorFromSH backward (backward.cu:21-145)
__device__ __forceinline__ void geom_sh_bwd(const GeomCommon& a, const GeomView& vw, int i, const float* mp, const float* dcol,
                                    float* dsh0, float* dmean, float* dtau) {
    const int M = a.M;
    if (a.shs && !a.colors_precomp) {  // computeColorFromSH backward (backward.cu:21-145)
        const float* sh = a.shs + (size_t)i * M * 3;
        float* dsh = a.dL_dsh + (size_t)i * M * 3;  // zeroed above unless accumulating: always add
        const int deg = a.sh_degree;
        const float dir0[3] = {mp[0] - vw.campos[0], mp[1] - vw.campos[1], mp[2] - vw.campos[2]};
        const float len = sqrtf(dir0[0] * dir0[0] + dir0[1] * dir0[1] + dir0[2] * dir0[2]);
        const float x = dir0[0] / len, y = dir0[1] / len, z = dir0[2] / len;
        const uint32_t cl = vw.clamped[i];
        float dRGB[3];
#pragma unroll
        for (int c = 0; c < 3; c++) dRGB[c] = dcol[c] * (((cl >> (8 * c)) & 0xffu) ? 0.f : 1.f);
        float dx_[3] = {0, 0, 0}, dy_[3] = {0, 0, 0}, dz_[3] = {0, 0, 0};
        for (int c = 0; c < 3; c++) dsh0[c] = B_SH_C0 * dRGB[c];
        if (deg > 0) {
            for (int c = 0; c < 3; c++) {
                dsh[3 + c] += -B_SH_C1 * y * dRGB[c]; dsh[6 + c] += B_SH_C1 * z * dRGB[c]; dsh[9 + c] += -B_SH_C1 * x * dRGB[c];
                dx_[c] = -B_SH_C1 * sh[9 + c]; dy_[c] = -B_SH_C1 * sh[3 + c]; dz_[c] = B_SH_C1 * sh[6 + c];
            }
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                for (int c = 0; c < 3; c++) {
                    dsh[12 + c] += B_SH_C2[0] * xy * dRGB[c]; dsh[15 + c] += B_SH_C2[1] * yz * dRGB[c];
                    dsh[18 + c] += B_SH_C2[2] * (2.f * zz - xx - yy) * dRGB[c]; dsh[21 + c] += B_SH_C2[3] * xz * dRGB[c];
                    dsh[24 + c] += B_SH_C2[4] * (xx - yy) * dRGB[c];
                    dx_[c] += B_SH_C2[0] * y * sh[12 + c] + B_SH_C2[2] * 2.f * -x * sh[18 + c] + B_SH_C2[3] * z * sh[21 + c] + B_SH_C2[4] * 2.f * x * sh[24 + c];
                    dy_[c] += B_SH_C2[0] * x * sh[12 + c] + B_SH_C2[1] * z * sh[15 + c] + B_SH_C2[2] * 2.f * -y * sh[18 + c] + B_SH_C2[4] * 2.f * -y * sh[24 + c];
                    dz_[c] += B_SH_C2[1] * y * sh[15 + c] + B_SH_C2[2] * 2.f * 2.f * z * sh[18 + c] + B_SH_C2[3] * x * sh[21 + c];
                }
                if (deg > 2) {
                    for (int c = 0; c < 3; c++) {
                        dsh[27 + c] += B_SH_C3[0] * y * (3.f * xx - yy) * dRGB[c]; dsh[30 + c] += B_SH_C3[1] * xy * z * dRGB[c];
                        dsh[33 + c] += B_SH_C3[2] * y * (4.f * zz - xx - yy) * dRGB[c];
                        dsh[36 + c] += B_SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy) * dRGB[c];
                        dsh[39 + c] += B_SH_C3[4] * x * (4.f * zz - xx - yy) * dRGB[c]; dsh[42 + c] += B_SH_C3[5] * z * (xx - yy) * dRGB[c];
                        dsh[45 + c] += B_SH_C3[6] * x * (xx - 3.f * yy) * dRGB[c];
                        dx_[c] += (B_SH_C3[0] * sh[27 + c] * 3.f * 2.f * xy + B_SH_C3[1] * sh[30 + c] * yz + B_SH_C3[2] * sh[33 + c] * -2.f * xy +
                                   B_SH_C3[3] * sh[36 + c] * -3.f * 2.f * xz + B_SH_C3[4] * sh[39 + c] * (-3.f * xx + 4.f * zz - yy) +
                                   B_SH_C3[5] * sh[42 + c] * 2.f * xz + B_SH_C3[6] * sh[45 + c] * 3.f * (xx - yy));
                        dy_[c] += (B_SH_C3[0] * sh[27 + c] * 3.f * (xx - yy) + B_SH_C3[1] * sh[30 + c] * xz + B_SH_C3[2] * sh[33 + c] * (-3.f * yy + 4.f * zz - xx) +
                                   B_SH_C3[3] * sh[36 + c] * -3.f * 2.f * yz + B_SH_C3[4] * sh[39 + c] * -2.f * xy + B_SH_C3[5] * sh[42 + c] * -2.f * yz +
                                   B_SH_C3[6] * sh[45 + c] * -3.f * 2.f * xy);
                        dz_[c] += (B_SH_C3[1] * sh[30 + c] * xy + B_SH_C3[2] * sh[33 + c] * 4.f * 2.f * yz + B_SH_C3[3] * sh[36 + c] * 3.f * (2.f * zz - xx - yy) +
                                   B_SH_C3[4] * sh[39 + c] * 4.f * 2.f * xz + B_SH_C3[5] * sh[42 + c] * (xx - yy));
                    }
                }
            }
        }
        const float ddir[3] = {dx_[0] * dRGB[0] + dx_[1] * dRGB[1] + dx_[2] * dRGB[2], dy_[0] * dRGB[0] + dy_[1] * dRGB[1] + dy_[2] * dRGB[2],
                               dz_[0] * dRGB[0] + dz_[1] * dRGB[1] + dz_[2] * dRGB[2]};
        const float sum2 = dir0[0] * dir0[0] + dir0[1] * dir0[1] + dir0[2] * dir0[2];
        const float inv32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
        const float dm[3] = {((+sum2 - dir0[0] * dir0[0]) * ddir[0] - dir0[1] * dir0[0] * ddir[1] - dir0[2] * dir0[0] * ddir[2]) * inv32,
                             (-dir0[0] * dir0[1] * ddir[0] + (sum2 - dir0[1] * dir0[1]) * ddir[1] - dir0[2] * dir0[1] * ddir[2]) * inv32,
                             (-dir0[0] * dir0[2] * ddir[0] - dir0[1] * dir0[2] * ddir[1] + (sum2 - dir0[2] * dir0[2]) * ddir[2]) * inv32};
#pragma unroll
        for (int k = 0; k < 3; k++) { dmean[k] += dm[k]; dtau[k] += -dm[k]; }
    }
}

// computeCov3D backward (backward.cu:350-413); no quaternion-normalisation Jacobian (:412)
__device__ __forceinline__ void geom_cov3d_bwd(const float* scales, const float* rotations, float scale_modifier_, int i,
                                       const float* dcov, float* dsc, float* dq) {
    if (scales) {  // computeCov3D backward (backward.cu:350-413); no quaternion-normalisation Jacobian (:412)
        const float4 q = reinterpret_cast<const float4*>(rotations)[i];
        const float* sc = scales + 3 * (size_t)i;
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        const float Rm[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                                {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                                {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
        const float sv[3] = {scale_modifier_ * sc[0], scale_modifier_ * sc[1], scale_modifier_ * sc[2]};
        float Mm[3][3];
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int rr = 0; rr < 3; rr++) Mm[c][rr] = sv[rr] * Rm[c][rr];
        const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]}, {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]}, {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
        float dMt[3][3];  // transpose of dL_dM = 2 * M * dL_dSigma
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int rr = 0; rr < 3; rr++) dMt[rr][c] = 2.0f * (Mm[0][rr] * dS[c][0] + Mm[1][rr] * dS[c][1] + Mm[2][rr] * dS[c][2]);
#pragma unroll
        for (int k = 0; k < 3; k++) dsc[k] = Rm[0][k] * dMt[k][0] + Rm[1][k] * dMt[k][1] + Rm[2][k] * dMt[k][2];
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int rr = 0; rr < 3; rr++) dMt[k][rr] *= sv[k];
        dq[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
        dq[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
        dq[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
        dq[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
    }
}

// A warp's 32 Gaussians own 32 * WD consecutive floats of a [P, WD] output.  Letting every thread store its own WD
// floats issues WD store instructions that each touch 32 different sectors; staged through shared memory the same
// WD instructions write 128 contiguous bytes each (8x fewer L2 transactions for the 15-float language rows).
template <int WD>
__device__ __forceinline__ void warp_store_rows(float* s, const float (&v)[WD], float* gbase, int n_rows, int lane) {
#pragma unroll
    for (int k = 0; k < WD; k++) s[lane * WD + k] = v[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < WD; k++) {
        const int t = k * 32 + lane;
        if (t < n_rows * WD) gbase[t] = s[t];
    }
    __syncwarp();
}

// Geometry backward for all V views of the batch.  A warp owns 32 Gaussians.  Only ~15 % of the (Gaussian, view) pairs
// carry a gradient (the others lie behind the saturation depth of every pixel they cover), so running the per-view
// chain rule one view at a time would execute it 8x per warp with ~5 live lanes.  Instead:
//   pass 1  one lane per Gaussian: which views touched its record (one byte per view) and the view-count / radius part
//           of the densification statistics; the warp compacts the touched (Gaussian, view) pairs into a queue;
//   pass 2  the queue is worked off 32 pairs at a time, one pair per lane: the view's packed gradient record is turned
//           into that view's mean / covariance / pose gradients with the view's camera and added to the Gaussian's
//           accumulator row in shared memory (the reference runs its two kernels once per view and lets autograd add
//           the V results); per-view outputs (screen-space gradient, pose gradient) are written from here;
//   pass 3  one lane per Gaussian again: scale / rotation gradient from the SUMMED covariance gradient (it is linear
//           in dL/dcov3D) and coalesced stores of the parameter gradients through a staging block.
constexpr int GA_MEAN = 0, GA_COV = 3, GA_SH0 = 9, GA_COL = 12, GA_OP = 15, GA_NORM = 16, GA_LANG = 17;  // accumulator row
template <int F>
__global__ void __launch_bounds__(256, 2) k_geometry_bwd(const __grid_constant__ GeomBwdArgs a) {
    constexpr int GA = GA_LANG + F;          // floats per accumulator row
    constexpr int GA_LD = GA | 1;            // odd leading dimension: conflict-free when lane = row
    __shared__ float s_tau[OLS_MAX_VIEWS][6];
    __shared__ float s_accum[8][32 * GA_LD];             // per warp: per-Gaussian sums over the views (pass 3 reuses it as the
                                                         // staging block of the coalesced stores)
    __shared__ uint16_t s_queue[8][32 * OLS_MAX_VIEWS];  // per warp: touched pairs (lane | view << 8)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* accw = s_accum[wid];
    float* stg = accw;
    static_assert(GA_LD >= F && GA_LD >= 6, "staging block fits the accumulator block");
    uint16_t* queue = s_queue[wid];
    const int warp_first = i - lane;                                  // first Gaussian of this warp
    const int n_rows = min(32, a.P - warp_first);                     // <= 0 for warps past the end
    const bool valid = i < a.P;
    const int M = a.M;
    constexpr int GRF = grad_floats(F);
    if (threadIdx.x < OLS_MAX_VIEWS * 6) (&s_tau[0][0])[threadIdx.x] = 0.0f;
    for (int e = lane; e < 32 * GA_LD; e += 32) accw[e] = 0.0f;
    __syncthreads();
    const bool acc = a.accumulate;
    const bool stats = a.stat_denom != nullptr;
    if (valid && a.dL_dsh && !acc && M > 1)
        for (int k = 3; k < 3 * M; k++) a.dL_dsh[(size_t)3 * M * i + k] = 0.0f;
    // ---- pass 1 ----
    float st_cnt = 0.0f;
    int st_rad = 0;
    uint32_t tmask = 0;
    if (valid) {
#pragma unroll 4
        for (int v = 0; v < a.V; v++) {
            const GeomView& vw = a.v[v];
            if (vw.gtouched == nullptr || vw.gtouched[i] != 0) tmask |= 1u << v;
            if (stats) {
                const int rad = vw.radii[i];
                if (rad > 0) { st_cnt += 1.0f; st_rad = max(st_rad, rad); }
            }
        }
    }
    int n_queue = 0;
    for (int v = 0; v < a.V; v++) {
        const GeomView& vw = a.v[v];
        const bool t = (tmask >> v) & 1u;
        const unsigned m = __ballot_sync(0xffffffffu, t);
        if (t) queue[n_queue + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(lane | (v << 8));
        n_queue += __popc(m);
        if (n_rows > 0) {  // per-view outputs exist for every Gaussian: zeros now, the touched pairs overwrite theirs in pass 2
            float* m2 = vw.dL_dmeans2D + 3 * (size_t)warp_first;
            for (int t_ = lane; t_ < 3 * n_rows; t_ += 32) m2[t_] = 0.0f;
            if (vw.dL_dtau) {
                float* dt = vw.dL_dtau + 6 * (size_t)warp_first;
                for (int t_ = lane; t_ < 6 * n_rows; t_ += 32) dt[t_] = 0.0f;
            }
        }
    }
    __syncwarp();
    // ---- pass 2: one touched (Gaussian, view) pair per lane ----
    for (int q0 = 0; q0 < n_queue; q0 += 32) {
        const int q = q0 + lane;
        if (q < n_queue) {
            const int l = queue[q] & 0xff, v = queue[q] >> 8;
            const int gi = warp_first + l;
            const GeomView& vw = a.v[v];
            float gr[GRF];
            const float4* src = reinterpret_cast<const float4*>(vw.gacc + (size_t)gi * GRF);
#pragma unroll
            for (int k4 = 0; k4 < GRF / 4; k4++) {
                const float4 t = src[k4];
                gr[4 * k4] = t.x; gr[4 * k4 + 1] = t.y; gr[4 * k4 + 2] = t.z; gr[4 * k4 + 3] = t.w;
            }
            bool any = false;
#pragma unroll
            for (int k = 0; k < GRF; k++) any = any || (gr[k] != 0.0f);
            if (any) {
                float* row = accw + l * GA_LD;
                const float g2x = gr[GR_MX], g2y = gr[GR_MY];
                const float dcol_v[3] = {gr[GR_RGB], gr[GR_RGB + 1], gr[GR_RGB + 2]};
                if (vw.radii[gi] > 0) {
                    const float mp[3] = {a.means3D[3 * (size_t)gi], a.means3D[3 * (size_t)gi + 1], a.means3D[3 * (size_t)gi + 2]};
                    const float* V = vw.viewmatrix;
                    const float* c3 = a.cov3D + 6 * (size_t)gi;
                    float dmean_v[3] = {0, 0, 0}, dcov_v[6] = {0, 0, 0, 0, 0, 0}, dsh0_v[3] = {0, 0, 0}, dtau[6] = {0, 0, 0, 0, 0, 0};
                    geom_cov2d_bwd(vw, V, mp, c3, gr[GR_CX], gr[GR_CY], gr[GR_CW], dcov_v, dmean_v, dtau);
                    geom_proj_bwd(V, vw.projmatrix, vw.projmatrix_raw, mp, g2x, g2y, gr[GR_DEPTH], dmean_v, dtau);
                    geom_sh_bwd(a, vw, gi, mp, dcol_v, dsh0_v, dmean_v, dtau);
#pragma unroll
                    for (int k = 0; k < 3; k++) { atomicAdd(row + GA_MEAN + k, dmean_v[k]); atomicAdd(row + GA_SH0 + k, dsh0_v[k]); }
#pragma unroll
                    for (int k = 0; k < 6; k++) atomicAdd(row + GA_COV + k, dcov_v[k]);
                    if (stats) atomicAdd(row + GA_NORM, sqrtf(g2x * g2x + g2y * g2y));   // gaussian_model.py:965-969
                    if (vw.dL_dtau) {
#pragma unroll
                        for (int k = 0; k < 6; k++) vw.dL_dtau[6 * (size_t)gi + k] = dtau[k];
                    }
                    if (vw.dL_dtau_sum) {  // reference: torch.sum(dL_dtau, dim=0) (__init__.py:383-385)
#pragma unroll
                        for (int k = 0; k < 6; k++)
                            if (dtau[k] != 0.0f) atomicAdd(&s_tau[v][k], dtau[k]);
                    }
                }
#pragma unroll
                for (int k = 0; k < 3; k++) atomicAdd(row + GA_COL + k, dcol_v[k]);
                atomicAdd(row + GA_OP, gr[GR_OP]);
#pragma unroll
                for (int k = 0; k < F; k++) atomicAdd(row + GA_LANG + k, gr[GR_LANG + k]);
                float* m2 = vw.dL_dmeans2D + 3 * (size_t)gi;
                m2[0] = g2x; m2[1] = g2y;
                queue[q] = (uint16_t)(queue[q] | 0x8000u);   // this pair really carried a gradient
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < a.V * 6) {
        const int v = threadIdx.x / 6, k = threadIdx.x - 6 * v;
        const float t = s_tau[v][k];
        if (a.v[v].dL_dtau_sum && t != 0.0f) atomicAdd(&a.v[v].dL_dtau_sum[k], t);
    }
    // ---- pass 3: one lane per Gaussian ----
    bool any_total = false;
    for (int q = 0; q < n_queue; q++) any_total = any_total || ((queue[q] & 0x80ffu) == (0x8000u | (unsigned)lane));
    const float* row = accw + lane * GA_LD;
    float dmean[3], dcov[6], dsh0[3], dcol[3], dlang[F];
#pragma unroll
    for (int k = 0; k < 3; k++) { dmean[k] = row[GA_MEAN + k]; dsh0[k] = row[GA_SH0 + k]; dcol[k] = row[GA_COL + k]; }
#pragma unroll
    for (int k = 0; k < 6; k++) dcov[k] = row[GA_COV + k];
#pragma unroll
    for (int k = 0; k < F; k++) dlang[k] = row[GA_LANG + k];
    const float dop = row[GA_OP];
    const float st_norm = row[GA_NORM];
    __syncwarp();   // every lane has read its accumulator row: the block now serves as the staging area
    if (stats && valid && st_cnt > 0.0f) {
        a.stat_accum[i] += st_norm;
        a.stat_denom[i] += st_cnt;
        a.stat_max_radii[i] = fmaxf(a.stat_max_radii[i], (float)st_rad);
    }
    float dsc[3] = {0, 0, 0}, dq[4] = {0, 0, 0, 0};
    if (valid && any_total) geom_cov3d_bwd(a.scales, a.rotations, a.scale_modifier, i, dcov, dsc, dq);
    if (!acc) {
        // overwrite mode: every Gaussian's gradients are written (zeros included), warp-cooperatively
        if (n_rows <= 0) return;
        const size_t w0 = (size_t)warp_first;
        warp_store_rows<3>(stg, dmean, a.dL_dmeans3D + 3 * w0, n_rows, lane);
        warp_store_rows<3>(stg, dsc, a.dL_dscales + 3 * w0, n_rows, lane);
        warp_store_rows<3>(stg, dcol, a.dL_dcolors + 3 * w0, n_rows, lane);
        warp_store_rows<6>(stg, dcov, a.dL_dcov3D + 6 * w0, n_rows, lane);
        warp_store_rows<4>(stg, dq, a.dL_drots + 4 * w0, n_rows, lane);
        warp_store_rows<F>(stg, dlang, a.dL_dlanguage + (size_t)F * w0, n_rows, lane);
        if (valid) a.dL_dopacity[i] = dop;
        if (a.dL_dsh) {
            if (M == 1) warp_store_rows<3>(stg, dsh0, a.dL_dsh + 3 * w0, n_rows, lane);
            else if (valid) { float* p_sh = a.dL_dsh + (size_t)3 * M * i; p_sh[0] = dsh0[0]; p_sh[1] = dsh0[1]; p_sh[2] = dsh0[2]; }
        }
        return;
    }
    if (!valid || !any_total) return;  // accumulating a zero gradient: nothing to read or rewrite
    a.dL_dopacity[i] += dop;
    float* p_mean = a.dL_dmeans3D + 3 * (size_t)i;
    float* p_cov = a.dL_dcov3D + 6 * (size_t)i;
    float* p_sc = a.dL_dscales + 3 * (size_t)i;
    float* p_q = a.dL_drots + 4 * (size_t)i;
    float* p_col = a.dL_dcolors + 3 * (size_t)i;
    float* p_lang = a.dL_dlanguage + (size_t)F * i;
    float* p_sh = a.dL_dsh ? a.dL_dsh + (size_t)3 * M * i : nullptr;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        p_mean[k] += dmean[k];
        p_sc[k] += dsc[k];
        p_col[k] += dcol[k];
        if (p_sh) p_sh[k] += dsh0[k];
    }
#pragma unroll
    for (int k = 0; k < 6; k++) p_cov[k] += dcov[k];
#pragma unroll
    for (int k = 0; k < 4; k++) p_q[k] += dq[k];
#pragma unroll
    for (int k = 0; k < F; k++) p_lang[k] += dlang[k];
}

// ---------------------------------------------------------------------------------------------------
// Disentangled variant: one thread per Gaussian, both footprints.  Reference order of operations
// (D/backward.cu:1504-1618): computeCov2DCUDA for radii > 0, computeCov2DCUDA_no_tau for radii_lang > 0,
// then language_preprocessCUDA -- which returns early unless BOTH radii are positive (:676-677), so in
// compat mode the projection / depth / SH / scale / rotation gradients of either footprint only exist for
// Gaussians visible in both lists.  Exact mode gates every term by the footprint it belongs to.
struct GeomDisArgs : GeomCommon {
    GeomView vw;  // the single view (its gacc / radii / dL_dmeans2D / dL_dtau belong to the colour footprint)
    const float *scales, *rotations, *cov3D, *scales_lang, *rotations_lang, *cov3D_lang;
    const int32_t* radii_lang;
    const float* gacc_lang;
    bool exact;
    float *dL_dcolors, *dL_dlanguage, *dL_dopacity, *dL_dopacity_lang, *dL_dmeans3D, *dL_dcov3D,
        *dL_dcov3D_lang, *dL_dscales, *dL_dscales_lang, *dL_drots, *dL_drots_lang;
};

template <int F>
__global__ void __launch_bounds__(256) k_geometry_bwd_dis(const GeomDisArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.P) return;
    const int M = a.M;
    constexpr int GRC = grad_floats(0), GRL = grad_floats(F);
    float gc[GRC], gl[GRL];
    {
        const float4* src = reinterpret_cast<const float4*>(a.vw.gacc + (size_t)i * GRC);
#pragma unroll
        for (int q = 0; q < GRC / 4; q++) {
            const float4 v = src[q];
            gc[4 * q] = v.x; gc[4 * q + 1] = v.y; gc[4 * q + 2] = v.z; gc[4 * q + 3] = v.w;
        }
        const float4* srl = reinterpret_cast<const float4*>(a.gacc_lang + (size_t)i * GRL);
#pragma unroll
        for (int q = 0; q < GRL / 4; q++) {
            const float4 v = srl[q];
            gl[4 * q] = v.x; gl[4 * q + 1] = v.y; gl[4 * q + 2] = v.z; gl[4 * q + 3] = v.w;
        }
    }
    float dmean[3] = {0, 0, 0}, dcov[6] = {0, 0, 0, 0, 0, 0}, dcovl[6] = {0, 0, 0, 0, 0, 0}, dtau[6] = {0, 0, 0, 0, 0, 0};
    float dsc[3] = {0, 0, 0}, dq[4] = {0, 0, 0, 0}, dscl[3] = {0, 0, 0}, dql[4] = {0, 0, 0, 0};
    float dcol[3] = {gc[GR_RGB], gc[GR_RGB + 1], gc[GR_RGB + 2]};
    float dsh0[3] = {0, 0, 0};
    const bool vis_c = a.vw.radii[i] > 0, vis_l = a.radii_lang[i] > 0;
    const float g2x = gc[GR_MX], g2y = gc[GR_MY];
    if (a.dL_dsh && M > 1)
        for (int k = 3; k < 3 * M; k++) a.dL_dsh[(size_t)3 * M * i + k] = 0.0f;
    const float* V = a.vw.viewmatrix;
    const float mp[3] = {a.means3D[3 * (size_t)i], a.means3D[3 * (size_t)i + 1], a.means3D[3 * (size_t)i + 2]};
    if (vis_c) geom_cov2d_bwd(a.vw, V, mp, a.cov3D + 6 * (size_t)i, gc[GR_CX], gc[GR_CY], gc[GR_CW], dcov, dmean, dtau);
    if (vis_l) {
        float dm_unused[3] = {0, 0, 0}, dt_unused[6] = {0, 0, 0, 0, 0, 0};
        geom_cov2d_bwd(a.vw, V, mp, a.cov3D_lang + 6 * (size_t)i, gl[GR_CX], gl[GR_CY], gl[GR_CW], dcovl, dm_unused, dt_unused);
    }
    const bool both = vis_c && vis_l;
    if (a.exact ? vis_c : both) {
        geom_proj_bwd(V, a.vw.projmatrix, a.vw.projmatrix_raw, mp, g2x, g2y, gc[GR_DEPTH], dmean, dtau);
        geom_sh_bwd(a, a.vw, i, mp, dcol, dsh0, dmean, dtau);
        geom_cov3d_bwd(a.scales, a.rotations, a.scale_modifier, i, dcov, dsc, dq);
    }
    if (a.exact ? vis_l : both) geom_cov3d_bwd(a.scales_lang, a.rotations_lang, a.scale_modifier, i, dcovl, dscl, dql);

    a.vw.dL_dmeans2D[3 * (size_t)i] = g2x;
    a.vw.dL_dmeans2D[3 * (size_t)i + 1] = g2y;
    a.vw.dL_dmeans2D[3 * (size_t)i + 2] = 0.0f;
    a.dL_dopacity[i] = gc[GR_OP];
    a.dL_dopacity_lang[i] = gl[GR_OP];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        a.dL_dmeans3D[3 * (size_t)i + k] = dmean[k];
        a.dL_dscales[3 * (size_t)i + k] = dsc[k];
        a.dL_dscales_lang[3 * (size_t)i + k] = dscl[k];
        a.dL_dcolors[3 * (size_t)i + k] = dcol[k];
        if (a.dL_dsh) a.dL_dsh[(size_t)3 * M * i + k] = dsh0[k];
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
        a.dL_dcov3D[6 * (size_t)i + k] = dcov[k];
        a.dL_dcov3D_lang[6 * (size_t)i + k] = dcovl[k];
        a.vw.dL_dtau[6 * (size_t)i + k] = dtau[k];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        a.dL_drots[4 * (size_t)i + k] = dq[k];
        a.dL_drots_lang[4 * (size_t)i + k] = dql[k];
    }
#pragma unroll
    for (int k = 0; k < F; k++) a.dL_dlanguage[(size_t)F * i + k] = gl[GR_LANG + k];
}

template <int TILE, int NCOL, int F>
static void launch_blend_bwd(const BwdBlendArgs& ba, int n_tiles, int V, bool exact, bool packed, cudaStream_t st) {
    const dim3 grid(n_tiles, V);
    static const bool v1 = getenv("OLS_BWD_V1") != nullptr;  // A/B aid: the one-pixel-per-lane kernel
    if (!v1) {
        if (exact)
            k_blend_bwd2<TILE, NCOL, F, false, false><<<grid, 128, 0, st>>>(ba);
        else if (packed)
            k_blend_bwd2<TILE, NCOL, F, true, true><<<grid, 64, 0, st>>>(ba);
        else
            k_blend_bwd2<TILE, NCOL, F, true, false><<<grid, 128, 0, st>>>(ba);
        return;
    }
    if (exact)
        k_blend_bwd<TILE, NCOL, F, false, false><<<grid, BWD_THREADS, 0, st>>>(ba);
    else if (packed)
        k_blend_bwd<TILE, NCOL, F, true, true><<<grid, 128, 0, st>>>(ba);
    else
        k_blend_bwd<TILE, NCOL, F, true, false><<<grid, BWD_THREADS, 0, st>>>(ba);
}

}  // namespace ols

using namespace ols;

size_t ols_bwd_scratch_bytes(int P, int F) { return sizeof(float) * (size_t)grad_floats(F) * (size_t)(P > 0 ? P : 1); }

// Q3: lanes of an n-thread block that reach data[0] in the reference's tree reduction
// (render_cuda_reduce_sum, backward.cu:684-702: i = n/2, n/4, ... with integer division)
static void reduce_lane_mask(int n, bool exact, uint32_t* mask8) {
    bool reach[256];
    for (int k = 0; k < 256; k++) reach[k] = exact;
    if (!exact) {
        // walk the reduction backwards: lane l reaches 0 iff repeatedly folding (l -> l - i when i <= l < 2i) ends at 0
        for (int l = 0; l < n; l++) {
            int pos = l;
            bool ok = true;
            for (int i = n / 2; i > 0; i /= 2) {
                if (pos >= i) {
                    if (pos < 2 * i) pos -= i; else { ok = false; break; }
                }
            }
            reach[l] = ok && pos == 0;
        }
    }
    for (int w = 0; w < 8; w++) {
        uint32_t m = 0;
        for (int b = 0; b < 32; b++) m |= (uint32_t)reach[w * 32 + b] << b;
        mask8[w] = m;
    }
}

// One view of a backward blend pass: its workspace (sorted list, records, final_T, n_contrib, gradient scratch) and the
// image-space gradients it starts from.
struct BwdPassView { char* ws; const float* bg; const float *dL_dcolor, *dL_dlanguage, *dL_ddepth; };

// backward blend of one pass for the V views of a batch (grid.y = view)
static int run_blend_bwd(int W, int H, int tile, int ncol, int F, unsigned flags, const BwdPassView* pv, int V, const WsLayout& L,
                         const float* language, cudaStream_t st) {
    const bool exact = (flags & OLS_FLAG_BWD_EXACT) != 0;
    BwdBlendArgs ba;
    ba.W = W; ba.H = H; ba.gx = L.gx;
    ba.language = language;
    ba.fast_exp = (flags & OLS_FLAG_BITEXACT_BLEND) ? 0 : 1;
    for (int v = 0; v < V; v++) {
        char* ws = pv[v].ws;
        BwdView& q = ba.v[v];
        q.ranges = (const uint2*)(ws + L.ranges); q.point_list = (const uint32_t*)(ws + L.point_list);
        q.records = (const float*)(ws + L.records); q.bg = pv[v].bg; q.info = (const DeviceInfo*)(ws + L.info);
        q.final_T = (const float*)(ws + L.final_T); q.n_contrib = (const uint32_t*)(ws + L.n_contrib);
        q.dL_dcolor = pv[v].dL_dcolor; q.dL_dlanguage = pv[v].dL_dlanguage; q.dL_ddepth = pv[v].dL_ddepth;
        q.gacc = (float*)(ws + L.gacc);
        q.gtouched = (uint8_t*)(ws + L.gtouched);
        q.warp_hits = (const uint8_t*)(ws + L.warp_hits);
    }
    reduce_lane_mask(tile * tile, exact, ba.lane_ok);
    // packed compat variant: possible when at most 128 pixels of a tile survive the reference's reduction (15x15: 128)
    bool packed = false;
    if (!exact && !(flags & OLS_FLAG_BWD_NO_PACK)) {
        int n_ok = 0;
        for (int r = 0; r < tile * tile; r++) {
            if ((ba.lane_ok[r >> 5] >> (r & 31)) & 1u) {
                if (n_ok < 128) ba.packed_rank[n_ok] = (uint8_t)r;
                n_ok++;
            }
        }
        packed = n_ok <= 128;
        for (int t = n_ok; t < 128; t++) ba.packed_rank[t] = 255;
        for (int k = 0; k < 4; k++) {
            ba.packed_fmask[k] = 0;
            for (int t = 32 * k; t < 32 * k + 32 && packed; t++) {
                if (ba.packed_rank[t] == 255) continue;
                const int lx = ba.packed_rank[t] % tile, ly = ba.packed_rank[t] / tile;
                ba.packed_fmask[k] |= (uint8_t)(1u << ((ly / 4) * 2 + lx / 8));  // forward: warp = (y / 4) * 2 + x / 8
            }
        }
    }
    const int key = tile * 10000 + ncol * 100 + F;
    switch (key) {
        case 150315: launch_blend_bwd<15, 3, 15>(ba, L.n_tiles, V, exact, packed, st); break;
        case 160315: launch_blend_bwd<16, 3, 15>(ba, L.n_tiles, V, exact, packed, st); break;
        case 150303: launch_blend_bwd<15, 3, 3>(ba, L.n_tiles, V, exact, packed, st); break;
        case 160303: launch_blend_bwd<16, 3, 3>(ba, L.n_tiles, V, exact, packed, st); break;
        case 150300: launch_blend_bwd<15, 3, 0>(ba, L.n_tiles, V, exact, packed, st); break;
        case 160300: launch_blend_bwd<16, 3, 0>(ba, L.n_tiles, V, exact, packed, st); break;
        case 150003: launch_blend_bwd<15, 0, 3>(ba, L.n_tiles, V, exact, packed, st); break;
        case 160003: launch_blend_bwd<16, 0, 3>(ba, L.n_tiles, V, exact, packed, st); break;
        case 150015: launch_blend_bwd<15, 0, 15>(ba, L.n_tiles, V, exact, packed, st); break;
        case 160015: launch_blend_bwd<16, 0, 15>(ba, L.n_tiles, V, exact, packed, st); break;
        default: ols_set_error("unsupported (tile=%d, F=%d)", tile, F); return OLS_ERR_UNSUPPORTED;
    }
    OLS_CUDA_TRY(cudaGetLastError());
    if (flags & OLS_FLAG_DEBUG) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { ols_set_error("kernel blend_bwd failed: %s", cudaGetErrorString(e)); return OLS_ERR_CUDA; }
    }
    return OLS_OK;
}

static void fill_geom_common(GeomCommon& ga, const ols_raster_args* a, float* dL_dsh) {
    ga.accumulate = (a->flags & OLS_FLAG_BWD_ACCUMULATE) != 0;
    ga.P = a->P; ga.F = a->F; ga.sh_degree = a->sh_degree; ga.M = a->M; ga.W = a->W; ga.H = a->H; ga.gr = grad_floats(a->F);
    ga.scale_modifier = a->scale_modifier;
    ga.means3D = a->d_means3D; ga.shs = a->d_shs;
    ga.colors_precomp = a->d_colors_precomp != nullptr;
    ga.dL_dsh = (a->d_shs && a->M > 0) ? dL_dsh : nullptr;
}

static void fill_geom_view(GeomView& q, const ols_raster_args* a, const char* ws, const WsLayout& L, const int32_t* radii,
                           float* dL_dmeans2D, float* dL_dtau, float* dL_dtau_sum) {
    q.viewmatrix = a->d_viewmatrix; q.projmatrix = a->d_projmatrix; q.projmatrix_raw = a->d_projmatrix_raw;
    q.campos = a->d_campos; q.clamped = (const uint32_t*)(ws + L.clamped); q.gacc = (const float*)(ws + L.gacc);
    q.gtouched = (const uint8_t*)(ws + L.gtouched);
    q.radii = radii; q.dL_dmeans2D = dL_dmeans2D; q.dL_dtau = dL_dtau; q.dL_dtau_sum = dL_dtau_sum;
    q.tanfovx = a->tanfovx; q.tanfovy = a->tanfovy;
    q.focal_y = a->H / (2.0f * a->tanfovy); q.focal_x = a->W / (2.0f * a->tanfovx);
}

// Backward of V views of the same Gaussians.  grads[v] carries the view's image-space gradients, its radii and its
// per-view outputs (dL_dmeans2D, dL_dtau / dL_dtau_sum); the parameter-gradient pointers are taken from grads[0] and
// receive the SUM over the views (added to their previous content with OLS_FLAG_BWD_ACCUMULATE).
int ols_launch_backward(const ols_raster_args* views, const ols_bwd_args* grads, int V, const WsLayout& L, cudaStream_t st) {
    const ols_raster_args* a = &views[0];
    const ols_bwd_args* g = &grads[0];
    const bool debug = (a->flags & OLS_FLAG_DEBUG) != 0;
    BwdPassView pv[OLS_MAX_VIEWS];
    for (int v = 0; v < V; v++) {
        char* ws = (char*)views[v].d_workspace;
        // gradient records and their touched bytes are adjacent in the workspace: one memset
        OLS_CUDA_TRY(cudaMemsetAsync(ws + L.gacc, 0, (L.gtouched - L.gacc) + (size_t)(a->P > 0 ? a->P : 1), st));
        if (grads[v].d_dL_dtau_sum) OLS_CUDA_TRY(cudaMemsetAsync(grads[v].d_dL_dtau_sum, 0, 6 * sizeof(float), st));
        pv[v] = BwdPassView{ws, views[v].d_bg, grads[v].d_dL_dout_color, grads[v].d_dL_dout_language, grads[v].d_dL_dout_depth};
    }
    ols_timing_mark(-1, st);
    int rc = run_blend_bwd(a->W, a->H, a->tile, 3, a->F, a->flags, pv, V, L, a->d_language, st);
    if (rc != OLS_OK) return rc;
    ols_timing_mark(OLS_T_BLEND_BWD, st);
    GeomBwdArgs ga;
    fill_geom_common(ga, a, g->d_dL_dsh);
    ga.V = V;
    ga.scales = a->d_scales; ga.rotations = a->d_rotations;
    // the view-independent 3D covariances live in the first view's workspace (k_preprocess)
    ga.cov3D = a->d_cov3D_precomp ? a->d_cov3D_precomp : (const float*)((char*)views[0].d_workspace + L.cov3D);
    ga.dL_dcolors = g->d_dL_dcolors; ga.dL_dlanguage = g->d_dL_dlanguage;
    ga.dL_dopacity = g->d_dL_dopacity; ga.dL_dmeans3D = g->d_dL_dmeans3D; ga.dL_dcov3D = g->d_dL_dcov3D;
    ga.dL_dscales = g->d_dL_dscales; ga.dL_drots = g->d_dL_drotations;
    const bool stats = g->d_stat_max_radii2D && g->d_stat_xyz_gradient_accum && g->d_stat_denom;
    ga.stat_max_radii = stats ? g->d_stat_max_radii2D : nullptr;
    ga.stat_accum = stats ? g->d_stat_xyz_gradient_accum : nullptr;
    ga.stat_denom = stats ? g->d_stat_denom : nullptr;
    for (int v = 0; v < V; v++)
        fill_geom_view(ga.v[v], &views[v], (const char*)views[v].d_workspace, L, grads[v].d_radii, grads[v].d_dL_dmeans2D,
                       grads[v].d_dL_dtau, grads[v].d_dL_dtau_sum);
    if (a->F == 15) k_geometry_bwd<15><<<(a->P + 255) / 256, 256, 0, st>>>(ga);
    else k_geometry_bwd<3><<<(a->P + 255) / 256, 256, 0, st>>>(ga);
    OLS_CUDA_TRY(cudaGetLastError());
    if (debug) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { ols_set_error("kernel geometry_bwd failed: %s", cudaGetErrorString(e)); return OLS_ERR_CUDA; }
    }
    ols_timing_mark(OLS_T_GEOMETRY_BWD, st);
    return OLS_OK;
}

int ols_launch_backward_dis(const ols_dis_args* d, const ols_dis_bwd_args* g, const WsLayout& Lc, const WsLayout& Ll,
                            size_t lang_base, cudaStream_t st) {
    const ols_raster_args* a = &d->base;
    char* wc = (char*)a->d_workspace;
    char* wl = wc + lang_base;
    const bool debug = (a->flags & OLS_FLAG_DEBUG) != 0;
    OLS_CUDA_TRY(cudaMemsetAsync(wc + Lc.gacc, 0, (Lc.gtouched - Lc.gacc) + (size_t)(a->P > 0 ? a->P : 1), st));
    OLS_CUDA_TRY(cudaMemsetAsync(wl + Ll.gacc, 0, (Ll.gtouched - Ll.gacc) + (size_t)(a->P > 0 ? a->P : 1), st));
    ols_timing_mark(-1, st);
    const BwdPassView pc{wc, a->d_bg, g->d_dL_dout_color, nullptr, g->d_dL_dout_depth};
    int rc = run_blend_bwd(a->W, a->H, a->tile, 3, 0, a->flags, &pc, 1, Lc, nullptr, st);
    if (rc != OLS_OK) return rc;
    const BwdPassView pl{wl, a->d_bg, nullptr, g->d_dL_dout_language, nullptr};
    rc = run_blend_bwd(a->W, a->H, a->tile, 0, a->F, a->flags, &pl, 1, Ll, nullptr, st);
    if (rc != OLS_OK) return rc;
    ols_timing_mark(OLS_T_BLEND_BWD, st);
    GeomDisArgs ga;
    fill_geom_common(ga, a, g->d_dL_dsh);
    fill_geom_view(ga.vw, a, wc, Lc, g->d_radii, g->d_dL_dmeans2D, g->d_dL_dtau, nullptr);
    ga.exact = (a->flags & OLS_FLAG_BWD_EXACT) != 0;
    ga.scales = a->d_scales; ga.rotations = a->d_rotations;
    ga.cov3D = a->d_cov3D_precomp ? a->d_cov3D_precomp : (const float*)(wc + Lc.cov3D);
    ga.scales_lang = d->d_scales_lang; ga.rotations_lang = d->d_rotations_lang;
    ga.cov3D_lang = d->d_cov3D_precomp_lang ? d->d_cov3D_precomp_lang : (const float*)(wl + Ll.cov3D);
    ga.radii_lang = g->d_radii_lang;
    ga.gacc_lang = (const float*)(wl + Ll.gacc);
    ga.dL_dcolors = g->d_dL_dcolors; ga.dL_dlanguage = g->d_dL_dlanguage;
    ga.dL_dopacity = g->d_dL_dopacity; ga.dL_dopacity_lang = g->d_dL_dopacity_lang; ga.dL_dmeans3D = g->d_dL_dmeans3D;
    ga.dL_dcov3D = g->d_dL_dcov3D; ga.dL_dcov3D_lang = g->d_dL_dcov3D_lang;
    ga.dL_dscales = g->d_dL_dscales; ga.dL_dscales_lang = g->d_dL_dscales_lang;
    ga.dL_drots = g->d_dL_drotations; ga.dL_drots_lang = g->d_dL_drotations_lang;
    if (a->F == 15) k_geometry_bwd_dis<15><<<(a->P + 255) / 256, 256, 0, st>>>(ga);
    else k_geometry_bwd_dis<3><<<(a->P + 255) / 256, 256, 0, st>>>(ga);
    OLS_CUDA_TRY(cudaGetLastError());
    if (debug) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { ols_set_error("kernel geometry_bwd_dis failed: %s", cudaGetErrorString(e)); return OLS_ERR_CUDA; }
    }
    ols_timing_mark(OLS_T_GEOMETRY_BWD, st);
    return OLS_OK;
}
