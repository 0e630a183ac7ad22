// ols_hr.cu -- the HR module (dense CLIP map up-sampler) as implicit-GEMM convolutions on tcgen05.
//
// Reference: language/supervisedNet.py:6-43 (AttentionFusion) and :45-109 (HighResLanguageFeatureNet.forward), run
// in eval mode under torch.no_grad() at utils/slam_backend.py:381-386,547-552.  The reference executes 13 cuDNN
// convolutions in fp32 NCHW plus BatchNorm / ReLU / sigmoid / cat / interpolate kernels between them.
//
// Design (B200-first):
//   * activations live in HBM as bf16 NHWC; a convolution is the GEMM  out[pixel, cout] = sum_{tap, cin}
//     in[pixel + (dy,dx)_tap, cin] * W[cout, tap, cin]  with M = 128 pixels (an 8 x 16 patch), fp32 accumulation in TMEM;
//   * the A operand of one (tap, 64-channel slab) step is ONE 3-D TMA box {64 ch, 8 px, 16 px} of the NHWC tensor,
//     shifted by the tap offset -- the zero padding of the convolution is TMA's out-of-bounds fill, and the box lands
//     in shared memory in exactly the 128-byte-swizzled K-major layout tcgen05.mma reads (no im2col anywhere);
//   * torch.cat([x, low_res]) is never materialised: the K loop walks two tensor maps;
//   * ConvTranspose2d(4, 2, 1) = four 2x2 convolutions, one per output parity class (blockIdx.z), each writing a
//     strided quarter of the output;
//   * BatchNorm is folded into the weights, bias + ReLU / the sigmoid gate `fused * att + fused` / the fp32 store of
//     the last layer run in the epilogue straight out of TMEM; the finished tile is staged in the (now idle) ring in
//     the swizzled layout and leaves by TMA store, which also clips partial tiles and scatters the parity classes of a
//     transposed convolution through a strided tensor map; the gate's `fused` tile arrives by TMA during the main loop;
//   * warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue; mbarrier ring of 4-8 stages;
//   * the 13 launches are chained with programmatic dependent launch: a layer's prologue (barrier init, TMEM
//     allocation, descriptor prefetch) overlaps the tail of the previous layer (griddepcontrol).
#include "ols_common.cuh"
#include "ols_tc.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace ols {

constexpr int HR_THREADS = 192;
constexpr int HR_BOX_W = 8, HR_BOX_H = 16;  // 128 pixels = one UMMA M tile
constexpr int HR_MAX_TAPS = 9;
constexpr int HR_A_BYTES = 128 * 128;       // one A slab: 128 pixels x 64 bf16

enum { HR_MODE_BF16 = 0, HR_MODE_GATE = 1, HR_MODE_F32 = 2 };

struct HrConv {
    CUtensorMap tmap_a[2];  // NHWC bf16 sources (second one: the low-resolution branch of a concatenation)
    CUtensorMap tmap_b;     // packed weights [n_classes * cout, n_taps * cin_total] bf16, K-major
    int n_src, slabs[2];    // 64-channel slabs per source
    int n_taps, n_classes;
    signed char dy[4][HR_MAX_TAPS], dx[4][HR_MAX_TAPS];
    int grid_w, grid_h, tiles_x;  // the GEMM's pixel grid (= output grid; input grid for a transposed convolution)
    int cout, bn;                 // output channels, channels per CTA
    int out_w, out_h, scale;      // output pixel = grid pixel * scale + class offset
    int mode, relu;
    const float* bias;
    CUtensorMap tmap_gate;        // HR_MODE_GATE: the fused feature the attention map multiplies (same layout as the output)
    CUtensorMap tmap_out[4];      // output tile store, one per parity class
    int n_stages, stage_bytes, gate_bytes;
    int split;                    // split-K: a cluster of `split` CTAs shares one output tile, rank 0 reduces
    unsigned long long* trace;    // development aid (OLS_HR_TRACE): globaltimer stamps of CTA (0,0,0)
};

__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
                 "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// split-K exchange inside a thread-block cluster: partial accumulators travel through distributed shared memory
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// Cluster barrier, used twice per thread in a split-K kernel:
//   #1 "everybody has started": arrive right after the prologue, wait just before the exchange -- distributed shared
//      memory of a CTA may only be written once that CTA is running (co-scheduling alone does not say it has begun);
//   #2 "partials have landed": arrive.release after the remote stores / wait.acquire before rank 0 reads them.
__device__ __forceinline__ void cluster_arrive_started() { asm volatile("barrier.cluster.arrive.relaxed;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_started() { asm volatile("barrier.cluster.wait;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_relacq() {
    asm volatile("barrier.cluster.arrive.release;\nbarrier.cluster.wait.acquire;" ::: "memory");
}
// programmatic dependent launch: wait for the producer grid's memory / allow the consumer grid to start its prologue
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__global__ void __launch_bounds__(HR_THREADS, 1) k_hr_conv(const __grid_constant__ HrConv p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* ring = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* gate_smem = ring + (size_t)p.n_stages * p.stage_bytes;  // [bn/64][128 px][128 B], swizzled
    uint8_t* part_smem = gate_smem + p.gate_bytes;                   // rank 0: [split-1][bn/32][128 px][128 B] fp32 partials
    uint64_t* bars = (uint64_t*)(part_smem + (size_t)(p.split - 1) * p.bn * 512);
    uint64_t* full = bars;
    uint64_t* empty = bars + 8;
    uint64_t* mma_done = bars + 16;
    uint64_t* gate_full = bars + 17;
    uint32_t* tmem_slot = (uint32_t*)(bars + 18);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool tracer = p.trace != nullptr && threadIdx.x == 64 && (blockIdx.x | blockIdx.y | blockIdx.z) == 0;
    if (tracer) p.trace[0] = gtimer();
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.n_stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(mma_done, 1);
        mbar_init(gate_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        // bn (64 / 128 / 256) accumulator columns: several CTAs of a narrow layer can share one SM's TMEM
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)p.bn));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();  // the next layer may start its prologue; it blocks in pdl_wait() until this grid is done
    if (p.split > 1) cluster_arrive_started();

    // split-K: the cluster's CTAs take consecutive ranges of the (tap, source, slab) steps of the same tile
    const uint32_t rank = p.split > 1 ? cluster_ctarank() : 0u;
    const int tile = blockIdx.x / p.split;
    const int x0 = (tile % p.tiles_x) * HR_BOX_W, y0 = (tile / p.tiles_x) * HR_BOX_H;
    const int n0 = blockIdx.y * p.bn;
    const int cls = blockIdx.z;
    const int slabs_total = p.slabs[0] + (p.n_src > 1 ? p.slabs[1] : 0);
    const int ksteps = p.n_taps * slabs_total;
    const int k_begin = ksteps * (int)rank / p.split, k_end = ksteps * ((int)rank + 1) / p.split;

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmap_a[0]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmap_b) : "memory");
            pdl_wait();  // everything below reads what the previous layers wrote
            if (p.mode == HR_MODE_GATE && rank == 0) {
                mbar_expect_tx(gate_full, (uint32_t)p.gate_bytes);
                for (int j = 0; j < p.bn / 64; j++)
                    tma_load_3d(gate_smem + (size_t)j * HR_A_BYTES, &p.tmap_gate, gate_full, n0 + j * 64, x0, y0);
            }
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t bytes = (uint32_t)(HR_A_BYTES + p.bn * 128);
            for (int ks = k_begin; ks < k_end; ks++) {
                const int t = ks / slabs_total, rem = ks - t * slabs_total;
                const int s = rem >= p.slabs[0] ? 1 : 0, sl = rem - (s ? p.slabs[0] : 0);
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* st = ring + (size_t)stage * p.stage_bytes;
                mbar_expect_tx(&full[stage], bytes);
                tma_load_3d(st, &p.tmap_a[s], &full[stage], sl * 64, x0 + p.dx[cls][t], y0 + p.dy[cls][t]);
                tma_load_2d(st + HR_A_BYTES, &p.tmap_b, &full[stage], ks * 64, cls * p.cout + n0);
                if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
            }
        }
        if (p.split > 1) { cluster_wait_started(); cluster_sync_relacq(); }
    } else if (warp == 1) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t idesc = make_idesc(false, p.bn);
        for (int ks = k_begin; ks < k_end; ks++) {
            mbar_wait(&full[stage], phase);
            tcgen05_fence_after();
            if (lane == 0) {
                const uint32_t a_addr = smem_u32(ring + (size_t)stage * p.stage_bytes);
                const uint32_t b_addr = a_addr + HR_A_BYTES;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    umma<false>(tmem_base, make_sdesc(a_addr + k * 32), make_sdesc(b_addr + k * 32), idesc, (ks > k_begin || k) ? 1u : 0u);
                umma_commit(&empty[stage]);
            }
            __syncwarp();
            if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
        }
        if (lane == 0) umma_commit(mma_done);
        __syncwarp();
        if (p.split > 1) { cluster_wait_started(); cluster_sync_relacq(); }
    } else {
        // epilogue: thread = one pixel of the patch = one TMEM lane
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        if (tracer) p.trace[1] = gtimer();
        mbar_wait(mma_done, 0);  // all MMAs retired: the accumulator is complete and the ring is idle
        tcgen05_fence_after();
        if (rank != 0) {
            // ship the partial accumulator into rank 0's shared memory, then meet at the cluster barrier
            cluster_wait_started();
            const uint32_t remote = map_to_rank(smem_u32(part_smem), 0) + (rank - 1) * (uint32_t)p.bn * 512u;
            for (int c = 0; c < p.bn; c += 32) {
                uint32_t r[32];
                tmem_ld32(t_lane + (uint32_t)c, r);
                tmem_ld_wait();
                const uint32_t slab = remote + (uint32_t)(c >> 5) * HR_A_BYTES;
#pragma unroll
                for (int q = 0; q < 8; q++)
                    st_cluster_f4(slab + sw128(row, q), __uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                  __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
            }
            cluster_sync_relacq();
        } else {
        if (p.split > 1) { cluster_wait_started(); cluster_sync_relacq(); }  // the other ranks' partials have landed in part_smem
        if (p.mode == HR_MODE_GATE) mbar_wait(gate_full, 0);
        if (tracer) p.trace[2] = gtimer();
        for (int c = 0; c < p.bn; c += 32) {
            uint32_t r[32];
            tmem_ld32(t_lane + (uint32_t)c, r);
            tmem_ld_wait();
            for (int pr = 0; pr < p.split - 1; pr++) {
                const uint8_t* ps = part_smem + (size_t)pr * p.bn * 512 + (size_t)(c >> 5) * HR_A_BYTES;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const float4 a = *reinterpret_cast<const float4*>(ps + sw128(row, q));
                    r[4 * q + 0] = __float_as_uint(__uint_as_float(r[4 * q + 0]) + a.x);
                    r[4 * q + 1] = __float_as_uint(__uint_as_float(r[4 * q + 1]) + a.y);
                    r[4 * q + 2] = __float_as_uint(__uint_as_float(r[4 * q + 2]) + a.z);
                    r[4 * q + 3] = __float_as_uint(__uint_as_float(r[4 * q + 3]) + a.w);
                }
            }
            float v[32];
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + c);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const float4 bq = __ldg(b4 + q);
                v[4 * q + 0] = __uint_as_float(r[4 * q + 0]) + bq.x;
                v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bq.y;
                v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bq.z;
                v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bq.w;
            }
            if (p.relu) {
#pragma unroll
                for (int i = 0; i < 32; i++) v[i] = fmaxf(v[i], 0.0f);
            }
            if (p.mode == HR_MODE_F32) {
                // staging slab = 32 fp32 channels x 128 pixels
                uint8_t* slab = ring + (size_t)(c >> 5) * HR_A_BYTES;
#pragma unroll
                for (int q = 0; q < 8; q++)
                    *reinterpret_cast<float4*>(slab + sw128(row, q)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
                // staging slab = 64 bf16 channels x 128 pixels; this pass fills 16-byte chunks j0 .. j0+3 of the row
                const int j0 = (c & 63) >> 3;
                if (p.mode == HR_MODE_GATE) {
                    // out = fused * sigmoid(att) + fused  (supervisedNet.py:40-41)
                    const uint8_t* gslab = gate_smem + (size_t)(c >> 6) * HR_A_BYTES;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint4 g = *reinterpret_cast<const uint4*>(gslab + sw128(row, j0 + q));
                        const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                        for (int h = 0; h < 4; h++) {
                            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[h]));
                            const int i = q * 8 + h * 2;
                            v[i] = fmaf(f.x, __fdividef(1.0f, 1.0f + __expf(-v[i])), f.x);
                            v[i + 1] = fmaf(f.y, __fdividef(1.0f, 1.0f + __expf(-v[i + 1])), f.y);
                        }
                    }
                }
                uint8_t* slab = ring + (size_t)(c >> 6) * HR_A_BYTES;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint4 pk;
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[q * 8 + 0], v[q * 8 + 1]);
                    __nv_bfloat162 h1 = __floats2bfloat162_rn(v[q * 8 + 2], v[q * 8 + 3]);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[q * 8 + 4], v[q * 8 + 5]);
                    __nv_bfloat162 h3 = __floats2bfloat162_rn(v[q * 8 + 6], v[q * 8 + 7]);
                    pk.x = *(uint32_t*)&h0; pk.y = *(uint32_t*)&h1; pk.z = *(uint32_t*)&h2; pk.w = *(uint32_t*)&h3;
                    *reinterpret_cast<uint4*>(slab + sw128(row, j0 + q)) = pk;
                }
            }
        }
        // the tile leaves by TMA store (clips partial tiles; strided map for the parity classes of a transposed conv)
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 2 && lane == 0) {
            const int per = p.mode == HR_MODE_F32 ? 32 : 64;
            for (int j = 0; j < p.bn / per; j++)
                tma_store_3d(&p.tmap_out[cls], ring + (size_t)j * HR_A_BYTES, n0 + j * per, x0, y0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        }  // rank 0
    }
    if (tracer) p.trace[3] = gtimer();
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.bn));
    }
    if (tracer) p.trace[4] = gtimer();
}

// ---- weight re-layout --------------------------------------------------------------------------------------------
struct HrPack {
    int transposed;  // 0: Conv2d [Cout,Cin,kh,kw]; 1: ConvTranspose2d [Cin,Cout,4,4]
    int cin, cout, kh, kw, n_taps, n_classes;
    signed char ky[4][HR_MAX_TAPS], kx[4][HR_MAX_TAPS];
};
// out[cls * cout + co][tap * cin + ci] (bf16)
__global__ void k_hr_pack(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, const HrPack a) {
    const size_t K = (size_t)a.n_taps * a.cin;
    const size_t total = (size_t)a.n_classes * a.cout * K;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % K);
        const int rowi = (int)(i / K);
        const int cls = rowi / a.cout, co = rowi % a.cout;
        const int tap = k / a.cin, ci = k % a.cin;
        const int ky = a.ky[cls][tap], kx = a.kx[cls][tap];
        const size_t src = a.transposed ? (((size_t)ci * a.cout + co) * a.kh + ky) * a.kw + kx
                                        : (((size_t)co * a.cin + ci) * a.kh + ky) * a.kw + kx;
        out[i] = __float2bfloat16_rn(w[src]);
    }
}

// ---- input conversion: NCHW fp32 -> NHWC bf16 with bilinear resizing (align_corners = False) ------------------------
// F.interpolate(mode='bilinear', align_corners=False), supervisedNet.py:88,97; identical sizes give an exact copy.
__device__ __forceinline__ void hr_tap(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.0f ? 0.0f : src;
    i0 = (int)src;
    i0 = i0 > in_size - 1 ? in_size - 1 : i0;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = src - (float)i0;
}
// block = 32 output pixels of one row x 64 channels
__global__ void __launch_bounds__(256) k_hr_resize(const float* __restrict__ src, int C, int hin, int win,
                                                  __nv_bfloat16* __restrict__ dst, int hout, int wout) {
    __shared__ float tile[64][33];
    const int xb = blockIdx.x * 32, y = blockIdx.y, cb = blockIdx.z * 64;
    const float sy = (float)hin / (float)hout, sx = (float)win / (float)wout;
    int y0, y1;
    float ly;
    hr_tap(y, sy, hin, y0, y1, ly);
    const int tx = threadIdx.x & 31, tc = threadIdx.x >> 5;
    const int x = xb + tx;
    if (x < wout) {
        int x0i, x1i;
        float lx;
        hr_tap(x, sx, win, x0i, x1i, lx);
        const float hx = 1.0f - lx, hy = 1.0f - ly;
        for (int c = tc; c < 64; c += 8) {
            const float* s = src + (size_t)(cb + c) * hin * win;
            tile[c][tx] = hy * (hx * s[(size_t)y0 * win + x0i] + lx * s[(size_t)y0 * win + x1i]) +
                          ly * (hx * s[(size_t)y1 * win + x0i] + lx * s[(size_t)y1 * win + x1i]);
        }
    }
    __syncthreads();
    // 32 pixels x 64 channels -> 32 x 128 bytes, 2 channels per thread
    for (int i = threadIdx.x; i < 32 * 32; i += 256) {
        const int px = i >> 5, c2 = (i & 31) * 2;
        if (xb + px < wout) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(tile[c2][px], tile[c2 + 1][px]);
            *reinterpret_cast<__nv_bfloat162*>(dst + ((size_t)y * wout + xb + px) * C + cb + c2) = h;
        }
    }
}

__global__ void k_hr_bf16_to_f32(const __nv_bfloat16* __restrict__ s, float* __restrict__ d, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        d[i] = __bfloat162float(s[i]);
}

}  // namespace ols

using namespace ols;

struct ols_hr_plan {
    HrConv conv[OLS_HR_N_CONV];
    dim3 grid[OLS_HR_N_CONV];
    size_t smem[OLS_HR_N_CONV];
    __nv_bfloat16* act_in[3];          // fv, f3 resized, f2 resized (NHWC bf16)
    __nv_bfloat16* act[OLS_HR_N_CONV]; // output of conv i (NHWC bf16), i < 12
    size_t act_elems[OLS_HR_N_CONV];
    std::vector<void*> owned;
    int S_h, S_w;
    cudaStream_t side;           // the two low_res_align convolutions only depend on the inputs: they run beside layers 0-1
    cudaEvent_t ev_fork, ev_join;
};

typedef CUresult (*PFN_encodeTiledHr)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiledHr hr_encode() {
    static PFN_encodeTiledHr fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiledHr)ptr;
    }
    return fn;
}

// pixel-major tensor [H, W, C] seen through arbitrary pixel strides: box {128 bytes of channels, 8, 16}
static int hr_map_px(CUtensorMap* map, const void* base, bool f32, int H, int W, int C, size_t stride_w_bytes,
                     size_t stride_h_bytes) {
    PFN_encodeTiledHr enc = hr_encode();
    if (!enc) { ols_set_error("cuTensorMapEncodeTiled not available"); return OLS_ERR_CUDA; }
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H};
    cuuint64_t strides[2] = {(cuuint64_t)stride_w_bytes, (cuuint64_t)stride_h_bytes};
    cuuint32_t box[3] = {f32 ? 32u : 64u, HR_BOX_W, HR_BOX_H};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ols_set_error("cuTensorMapEncodeTiled (activation) failed (%d)", (int)r); return OLS_ERR_CUDA; }
    return OLS_OK;
}
// NHWC bf16 activation [H, W, C]
static int hr_map_act(CUtensorMap* map, const void* base, int H, int W, int C) {
    return hr_map_px(map, base, false, H, W, C, (size_t)C * 2, (size_t)W * C * 2);
}
// output of a layer: class (py, px) of a scale-2 layer owns the pixels (2y + py, 2x + px)
static int hr_map_out(CUtensorMap* map, void* base, bool f32, const ols::HrConv& c, int cls) {
    const size_t esz = f32 ? 4 : 2;
    const size_t px = (size_t)c.cout * esz;
    uint8_t* b = (uint8_t*)base + ((size_t)(cls >> 1) * c.out_w + (cls & 1)) * px;
    return hr_map_px(map, b, f32, c.grid_h, c.grid_w, c.cout, px * c.scale, px * c.out_w * c.scale);
}
// packed weights [rows, K] bf16: box {64, bn}
static int hr_map_w(CUtensorMap* map, const void* base, int rows, int K, int bn) {
    PFN_encodeTiledHr enc = hr_encode();
    if (!enc) { ols_set_error("cuTensorMapEncodeTiled not available"); return OLS_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ols_set_error("cuTensorMapEncodeTiled (weights) failed (%d)", (int)r); return OLS_ERR_CUDA; }
    return OLS_OK;
}

namespace {
enum Kind { K_CONV3, K_CONV1, K_CONVT };
struct Spec {
    Kind kind;
    int cin0, cin1, cout;
    int level;     // pixel grid of the GEMM = S << level
    int src0, src1;  // activation ids: >= 0 conv output, -1 fv, -2 f3 resized, -3 f2 resized; src1 = -100: none
    int mode, relu, gate;
    int bn;        // output channels per CTA (64, 128 or 256 = TMEM columns)
    int max_stages;  // ring depth cap (a short K loop needs no deep ring: several CTAs then fit one SM)
    int split;       // split-K cluster size: the low-resolution layers have too few tiles to fill 148 SMs otherwise
};
// supervisedNet.py:83-109, one row per convolution
const Spec SPECS[OLS_HR_N_CONV] = {
    {K_CONV3, 768, 0, 512, 0, -1, -100, HR_MODE_BF16, 1, -1, 64, 8, 2},     // initial_conv (+BN+ReLU)
    {K_CONVT, 512, 0, 512, 0, 0, -100, HR_MODE_BF16, 1, -1, 128, 8, 1},     // upsample1 (+BN+ReLU) -> 2S
    {K_CONV1, 384, 0, 512, 1, -2, -100, HR_MODE_BF16, 0, -1, 128, 8, 1},    // af1.low_res_align
    {K_CONV3, 512, 512, 512, 1, 1, 2, HR_MODE_BF16, 1, -1, 128, 8, 2},       // af1.fusion (+BN+ReLU) on cat[x, low]
    {K_CONV3, 512, 0, 512, 1, 3, -100, HR_MODE_BF16, 1, -1, 128, 8, 2},      // af1.attention.0 (+BN+ReLU)
    {K_CONV1, 512, 0, 512, 1, 4, -100, HR_MODE_GATE, 0, 3, 128, 8, 1},      // af1.attention.3 + sigmoid, gate on fused
    {K_CONVT, 512, 0, 256, 1, 5, -100, HR_MODE_BF16, 1, -1, 128, 8, 1},     // upsample2 -> 4S
    {K_CONV1, 192, 0, 256, 2, -3, -100, HR_MODE_BF16, 0, -1, 128, 8, 1},    // af2.low_res_align
    {K_CONV3, 256, 256, 256, 2, 6, 7, HR_MODE_BF16, 1, -1, 128, 8, 1},      // af2.fusion
    {K_CONV3, 256, 0, 256, 2, 8, -100, HR_MODE_BF16, 1, -1, 128, 8, 1},     // af2.attention.0
    {K_CONV1, 256, 0, 256, 2, 9, -100, HR_MODE_GATE, 0, 8, 128, 8, 1},      // af2.attention.3 + gate
    {K_CONVT, 256, 0, 128, 2, 10, -100, HR_MODE_BF16, 1, -1, 128, 8, 1},    // upsample3 -> 8S
    {K_CONV1, 128, 0, 768, 3, 11, -100, HR_MODE_F32, 0, -1, 128, 2, 1},     // final_conv -> fp32
};
}  // namespace

extern "C" {

void ols_hr_plan_destroy(ols_hr_plan* plan) {
    if (!plan) return;
    for (void* q : plan->owned) cudaFree(q);
    if (plan->side) cudaStreamDestroy(plan->side);
    if (plan->ev_fork) cudaEventDestroy(plan->ev_fork);
    if (plan->ev_join) cudaEventDestroy(plan->ev_join);
    delete plan;
}

int ols_hr_plan_create(const ols_hr_weights* w, int32_t S_h, int32_t S_w, ols_hr_plan** out_plan, void* stream) {
    if (!w || !out_plan || S_h <= 0 || S_w <= 0) { ols_set_error("bad HR plan arguments"); return OLS_ERR_INVALID; }
    for (int i = 0; i < OLS_HR_N_CONV; i++)
        if (!w->d_weight[i] || !w->d_bias[i]) { ols_set_error("HR conv %d: null weight or bias", i); return OLS_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    ols_hr_plan* plan = new ols_hr_plan();
    memset(plan->conv, 0, sizeof(plan->conv));
    plan->S_h = S_h; plan->S_w = S_w;
    plan->side = nullptr; plan->ev_fork = nullptr; plan->ev_join = nullptr;
    if (cudaStreamCreateWithFlags(&plan->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&plan->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&plan->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        ols_set_error("cannot create the HR side stream"); ols_hr_plan_destroy(plan); return OLS_ERR_CUDA;
    }
    auto fail = [&](int rc) { ols_hr_plan_destroy(plan); return rc; };
    auto alloc = [&](size_t bytes) -> void* {
        void* q = nullptr;
        if (cudaMalloc(&q, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        plan->owned.push_back(q);
        return q;
    };
    const int in_c[3] = {768, 384, 192};
    for (int i = 0; i < 3; i++) {
        const size_t n = (size_t)(S_h << i) * (S_w << i) * in_c[i];
        plan->act_in[i] = (__nv_bfloat16*)alloc(n * 2);
        if (!plan->act_in[i]) { ols_set_error("out of device memory"); return fail(OLS_ERR_CUDA); }
    }
    for (int i = 0; i < OLS_HR_N_CONV; i++) {
        const Spec& s = SPECS[i];
        HrConv& c = plan->conv[i];
        const int gh = S_h << s.level, gw = S_w << s.level;
        c.grid_w = gw; c.grid_h = gh;
        c.tiles_x = (gw + HR_BOX_W - 1) / HR_BOX_W;
        const int tiles_y = (gh + HR_BOX_H - 1) / HR_BOX_H;
        c.cout = s.cout; c.bn = s.bn;
        c.mode = s.mode; c.relu = s.relu;
        c.n_src = s.cin1 > 0 ? 2 : 1;
        c.slabs[0] = s.cin0 / 64; c.slabs[1] = s.cin1 / 64;
        HrPack pk;
        memset(&pk, 0, sizeof(pk));
        pk.cin = s.cin0 + s.cin1; pk.cout = s.cout;
        if (s.kind == K_CONV3) {
            c.n_taps = 9; c.n_classes = 1; c.scale = 1; pk.kh = pk.kw = 3;
            for (int t = 0; t < 9; t++) { pk.ky[0][t] = t / 3; pk.kx[0][t] = t % 3; c.dy[0][t] = t / 3 - 1; c.dx[0][t] = t % 3 - 1; }
        } else if (s.kind == K_CONV1) {
            c.n_taps = 1; c.n_classes = 1; c.scale = 1; pk.kh = pk.kw = 1;
        } else {
            // out[2m + py] = sum_iy in[iy] w[2m + py + 1 - 2 iy]: py = 0 -> (iy, ky) in {(m,1), (m-1,3)}; py = 1 -> {(m,2), (m+1,0)}
            c.n_taps = 4; c.n_classes = 4; c.scale = 2; pk.kh = pk.kw = 4; pk.transposed = 1;
            const int d[2][2] = {{0, -1}, {0, 1}}, k[2][2] = {{1, 3}, {2, 0}};
            for (int cls = 0; cls < 4; cls++)
                for (int t = 0; t < 4; t++) {
                    const int py = cls >> 1, px = cls & 1, a = t >> 1, b = t & 1;
                    c.dy[cls][t] = d[py][a]; pk.ky[cls][t] = k[py][a];
                    c.dx[cls][t] = d[px][b]; pk.kx[cls][t] = k[px][b];
                }
        }
        pk.n_taps = c.n_taps; pk.n_classes = c.n_classes;
        c.out_w = gw * c.scale; c.out_h = gh * c.scale;
        if (s.cout % c.bn != 0 || (c.bn != 64 && c.bn != 128 && c.bn != 256)) { ols_set_error("HR conv %d: bad channel block", i); return fail(OLS_ERR_INVALID); }
        // packed weights + bias copy
        const int K = c.n_taps * pk.cin, rows = c.n_classes * s.cout;
        __nv_bfloat16* wbuf = (__nv_bfloat16*)alloc((size_t)rows * K * 2);
        float* bbuf = (float*)alloc(sizeof(float) * s.cout);
        if (!wbuf || !bbuf) { ols_set_error("out of device memory"); return fail(OLS_ERR_CUDA); }
        k_hr_pack<<<1024, 256, 0, st>>>(w->d_weight[i], wbuf, pk);
        if (cudaMemcpyAsync(bbuf, w->d_bias[i], sizeof(float) * s.cout, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            ols_set_error("bias copy failed"); return fail(OLS_ERR_CUDA);
        }
        c.bias = bbuf;
        int rc = hr_map_w(&c.tmap_b, wbuf, rows, K, c.bn);
        if (rc != OLS_OK) return fail(rc);
        // output activation
        plan->act_elems[i] = (size_t)c.out_w * c.out_h * s.cout;
        if (s.mode != HR_MODE_F32) {
            plan->act[i] = (__nv_bfloat16*)alloc(plan->act_elems[i] * 2);
            if (!plan->act[i]) { ols_set_error("out of device memory"); return fail(OLS_ERR_CUDA); }
            for (int cls = 0; cls < c.n_classes; cls++) {
                rc = hr_map_out(&c.tmap_out[cls], plan->act[i], false, c, cls);
                if (rc != OLS_OK) return fail(rc);
            }
        }
        // sources
        const int srcs[2] = {s.src0, s.src1};
        const int cins[2] = {s.cin0, s.cin1};
        for (int k = 0; k < c.n_src; k++) {
            const __nv_bfloat16* base = srcs[k] >= 0 ? plan->act[srcs[k]] : plan->act_in[-srcs[k] - 1];
            rc = hr_map_act(&c.tmap_a[k], base, gh, gw, cins[k]);
            if (rc != OLS_OK) return fail(rc);
        }
        c.gate_bytes = 0;
        if (s.gate >= 0) {
            rc = hr_map_act(&c.tmap_gate, plan->act[s.gate], gh, gw, s.cout);
            if (rc != OLS_OK) return fail(rc);
            c.gate_bytes = c.bn / 64 * HR_A_BYTES;
        }
        c.stage_bytes = HR_A_BYTES + c.bn * 128;
        static const bool no_split = getenv("OLS_HR_NO_SPLIT") != nullptr;
        c.split = no_split ? 1 : s.split;
        const int part_bytes = (c.split - 1) * c.bn * 512;
        int ns = (227 * 1024 - 1024 - 256 - c.gate_bytes - part_bytes) / c.stage_bytes;
        c.n_stages = ns > s.max_stages ? s.max_stages : ns;
        // the finished tile is staged in the ring: 128 px x bn channels (bf16, or fp32 for the last layer)
        const int staging = c.bn * 128 * (s.mode == HR_MODE_F32 ? 4 : 2);
        if (c.n_stages < 2 || staging > c.n_stages * c.stage_bytes || c.bn % 64 != 0) {
            ols_set_error("HR conv %d: tile does not fit shared memory", i); return fail(OLS_ERR_UNSUPPORTED);
        }
        plan->smem[i] = (size_t)c.n_stages * c.stage_bytes + c.gate_bytes + part_bytes + 256 + 1024;
        plan->grid[i] = dim3((unsigned)(c.tiles_x * tiles_y * c.split), (unsigned)(s.cout / c.bn), (unsigned)c.n_classes);
    }
    if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) { ols_set_error("HR weight packing failed"); return fail(OLS_ERR_CUDA); }
    if (cudaFuncSetAttribute(k_hr_conv, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
        ols_set_error("cannot reserve shared memory for the HR kernel"); return fail(OLS_ERR_CUDA);
    }
    *out_plan = plan;
    return OLS_OK;
}

static int hr_run(const ols_hr_plan* plan, const float* d_fv, const float* d_f3, int32_t h3, int32_t w3, const float* d_f2,
                  int32_t h2, int32_t w2, float* d_out, void* d_feat, void* stream);

int ols_hr_forward(const ols_hr_plan* plan, const float* d_fv, const float* d_f3, int32_t h3, int32_t w3,
                   const float* d_f2, int32_t h2, int32_t w2, float* d_out, void* stream) {
    if (!d_out) { ols_set_error("bad HR forward arguments"); return OLS_ERR_INVALID; }
    return hr_run(plan, d_fv, d_f3, h3, w3, d_f2, h2, w2, d_out, nullptr, stream);
}

int ols_hr_forward_features(const ols_hr_plan* plan, const float* d_fv, const float* d_f3, int32_t h3, int32_t w3,
                            const float* d_f2, int32_t h2, int32_t w2, void* d_feat_bf16, void* stream) {
    if (!d_feat_bf16) { ols_set_error("bad HR forward arguments"); return OLS_ERR_INVALID; }
    return hr_run(plan, d_fv, d_f3, h3, w3, d_f2, h2, w2, nullptr, d_feat_bf16, stream);
}

}  // extern "C"

// d_out != NULL: all 13 convolutions; d_feat != NULL: convolutions 0..11, the last one writing into the caller's buffer
static int hr_run(const ols_hr_plan* plan, const float* d_fv, const float* d_f3, int32_t h3, int32_t w3, const float* d_f2,
                  int32_t h2, int32_t w2, float* d_out, void* d_feat, void* stream) {
    if (!plan || !d_fv || !d_f3 || !d_f2 || h3 <= 0 || w3 <= 0 || h2 <= 0 || w2 <= 0) {
        ols_set_error("bad HR forward arguments"); return OLS_ERR_INVALID;
    }
    if ((((uintptr_t)d_out | (uintptr_t)d_feat) & 15) != 0) { ols_set_error("HR output must be 16-byte aligned"); return OLS_ERR_INVALID; }
    const int n_conv = d_out ? OLS_HR_N_CONV : OLS_HR_N_CONV - 1;
    cudaStream_t st = (cudaStream_t)stream;
    const float* src[3] = {d_fv, d_f3, d_f2};
    const int hin[3] = {plan->S_h, h3, h2}, win[3] = {plan->S_w, w3, w2}, C[3] = {768, 384, 192};
    ols_timing_mark(-1, st);
    for (int i = 0; i < 3; i++) {
        const int ho = plan->S_h << i, wo = plan->S_w << i;
        k_hr_resize<<<dim3((wo + 31) / 32, ho, C[i] / 64), 256, 0, st>>>(src[i], C[i], hin[i], win[i], plan->act_in[i], ho, wo);
    }
    static const bool trace_on = getenv("OLS_HR_TRACE") != nullptr;  // development aid: per-layer times on stderr
    cudaEvent_t ev[OLS_HR_N_CONV + 1];
    static unsigned long long* d_trace = nullptr;
    if (trace_on && !d_trace) cudaMalloc(&d_trace, OLS_HR_N_CONV * 8 * sizeof(unsigned long long));
    if (trace_on) {
        for (auto& e : ev) cudaEventCreate(&e);
        cudaEventRecord(ev[0], st);
    }
    static const bool no_pdl = getenv("OLS_HR_NO_PDL") != nullptr;
    static const bool no_fork = getenv("OLS_HR_NO_FORK") != nullptr;
    const bool fork = !no_fork && !trace_on;
    auto launch = [&](int i, cudaStream_t stream_i, bool pdl) -> int {
        HrConv c = plan->conv[i];
        if (c.mode == HR_MODE_F32) {
            int rc = hr_map_out(&c.tmap_out[0], d_out, true, c, 0);
            if (rc != OLS_OK) return rc;
        }
        if (d_feat && i == OLS_HR_N_CONV - 2) {
            for (int cls = 0; cls < c.n_classes; cls++) {
                int rc = hr_map_out(&c.tmap_out[cls], d_feat, false, c, cls);
                if (rc != OLS_OK) return rc;
            }
        }
        c.trace = trace_on ? d_trace + i * 8 : nullptr;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = plan->grid[i];
        cfg.blockDim = dim3(HR_THREADS, 1, 1);
        cfg.dynamicSmemBytes = plan->smem[i];
        cfg.stream = stream_i;
        cudaLaunchAttribute attr[2];
        int na = 0;
        if (pdl && !no_pdl && !trace_on) {
            // programmatic dependent launch: this layer may begin its prologue while the previous one drains
            attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[na].val.programmaticStreamSerializationAllowed = 1;
            na++;
        }
        if (c.split > 1) {
            attr[na].id = cudaLaunchAttributeClusterDimension;
            attr[na].val.clusterDim.x = (unsigned)c.split; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
            na++;
        }
        cfg.attrs = attr;
        cfg.numAttrs = na;
        OLS_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_hr_conv, c));
        if (trace_on) cudaEventRecord(ev[i + 1], stream_i);
        return OLS_OK;
    };
    if (fork) {
        // conv 2 and conv 7 (low_res_align of f3 / f2) beside conv 0 and conv 1; joined before conv 3
        OLS_CUDA_TRY(cudaEventRecord(plan->ev_fork, st));
        OLS_CUDA_TRY(cudaStreamWaitEvent(plan->side, plan->ev_fork, 0));
        int rc = launch(2, plan->side, false);
        if (rc == OLS_OK) rc = launch(7, plan->side, true);
        if (rc != OLS_OK) return rc;
        OLS_CUDA_TRY(cudaEventRecord(plan->ev_join, plan->side));
    }
    for (int i = 0; i < n_conv; i++) {
        if (fork && (i == 2 || i == 7)) continue;
        bool pdl = true;
        if (fork && i == 3) { OLS_CUDA_TRY(cudaStreamWaitEvent(st, plan->ev_join, 0)); pdl = false; }
        int rc = launch(i, st, pdl);
        if (rc != OLS_OK) return rc;
    }
    OLS_CUDA_TRY(cudaGetLastError());
    ols_timing_mark(OLS_T_OTHER, st);
    if (trace_on && d_out) {
        cudaStreamSynchronize(st);
        fprintf(stderr, "[hr trace]");
        for (int i = 0; i < OLS_HR_N_CONV; i++) {
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            const HrConv& c = plan->conv[i];
            const double flop = 2.0 * c.grid_w * c.grid_h * c.n_classes * c.cout * 64.0 * c.n_taps *
                                (c.slabs[0] + (c.n_src > 1 ? c.slabs[1] : 0));
            fprintf(stderr, " L%d[%ux%ux%u bn%d] %.1fus %.0fTF |", i, plan->grid[i].x, plan->grid[i].y, plan->grid[i].z, c.bn,
                    ms * 1e3, flop / (ms * 1e-3) * 1e-12);
        }
        fprintf(stderr, "\n[hr trace] CTA 0 (setup, main loop, epilogue, exit) us:");
        unsigned long long h[OLS_HR_N_CONV * 8];
        cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
        for (int i = 0; i < OLS_HR_N_CONV; i++) {
            const unsigned long long* t = h + i * 8;
            fprintf(stderr, " L%d %.1f %.1f %.1f %.1f (gap to next start %.1f) |", i, (t[1] - t[0]) * 1e-3, (t[2] - t[1]) * 1e-3,
                    (t[3] - t[2]) * 1e-3, (t[4] - t[3]) * 1e-3, i + 1 < OLS_HR_N_CONV ? (double)(h[(i + 1) * 8] - t[4]) * 1e-3 : 0.0);
        }
        fprintf(stderr, "\n");
        for (auto& e : ev) cudaEventDestroy(e);
    }
    return OLS_OK;
}

extern "C" {

int ols_hr_read_activation(const ols_hr_plan* plan, int32_t which, float* d_out, int64_t capacity_floats, void* stream) {
    if (!plan || which < 0 || which >= OLS_HR_N_CONV - 1 || !d_out) { ols_set_error("bad activation index"); return OLS_ERR_INVALID; }
    const size_t n = plan->act_elems[which];
    if ((int64_t)n > capacity_floats) { ols_set_error("activation buffer too small (%zu floats needed)", n); return OLS_ERR_INVALID; }
    k_hr_bf16_to_f32<<<592, 256, 0, (cudaStream_t)stream>>>(plan->act[which], d_out, n);
    OLS_CUDA_TRY(cudaGetLastError());
    return OLS_OK;
}

}  // extern "C"
