// ols_ae.cu -- the per-frame language autoencoder as ONE fused tensor-core kernel for sm_100a.
//
// Reference: language/autoencoder/model.py:15-62 (AutoencoderMLP.encode / .decode) and :314-354
// (EncoderDecoderOnline): y = normalize( L_n( relu( ... relu( L_1(x) ) ... ) ) ), applied to the
// [H*W, 768] flattened CLIP map (utils/slam_backend.py:392-395,557-559).  Eval-mode BatchNorm is an
// affine map and is folded into the preceding Linear by the caller.
//
// Design (B200-first, not a translation of the torch module):
//   * persistent kernel, one CTA per SM, a CTA owns 128-row tiles of the activation matrix;
//   * the whole layer chain runs inside the kernel: the 128 x N_l fp32 accumulator of a layer lives in
//     TMEM (tcgen05.mma, M = 128), the epilogue warps read it back (tcgen05.ld), add bias, apply ReLU,
//     convert to bf16 and write it into shared memory in the 128-byte-swizzled K-major layout the next
//     layer's tcgen05.mma reads as its A operand -- activations never touch HBM between layers, so the
//     algorithmic HBM traffic is x in, y out (+ the weights once, they stay L2 resident);
//   * layer 0 multiplies the fp32 input directly (kind::tf32, x tiles arrive by TMA, no conversion
//     pass); the inner layers run kind::f16 on bf16 with fp32 accumulation;
//   * weights are streamed through a TMA ring of K-slabs (128 bytes of K per row), one elected thread
//     issues TMA, one elected thread issues MMA, four warps run the epilogue; mbarriers only.
//   * the final row-wise L2 normalisation is fused into the last epilogue (one thread owns one row).
#include "ols_common.cuh"

#include "ols_tc.cuh"

#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace ols {

constexpr int AE_M = 128;             // rows per tile (UMMA M)
constexpr int AE_THREADS = 192;       // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int AE_SLAB_BYTES = 128;    // bytes of K per row in one slab (SWIZZLE_128B)
constexpr int AE_BLOCK_MAX = 384;     // max rows of a weight slab resident in one ring stage
constexpr int AE_CHUNK_MAX = 256;     // max N of one tcgen05.mma
constexpr int AE_TMEM_COLS = 512;
constexpr int AE_MAX_BLOCKS = 4;
constexpr int AE_STAGING_BYTES = 4 * 32 * 33 * 4 + 128;  // per epilogue warp: a 32 x 32 fp32 chunk, rows padded to 33 (+ pad to 256 B)

struct AeLayer {
    int K, N;            // padded: K multiple of the slab width, N multiple of 16
    int n_slabs;         // K / (elements per 128-byte slab)
    int block_n;         // rows of the weight slab streamed per stage (<= AE_BLOCK_MAX)
    int chunk_n;         // N of one MMA instruction (block_n or block_n / 2)
    int n_blocks;        // N / block_n
    int blocks_per_pass; // N-blocks accumulated in TMEM before the epilogue runs (pass width <= 512 columns)
    int sps;             // K-slabs of this layer's weights that travel in ONE ring stage (inner layers' slabs are small: fewer
                         // load -> MMA -> release round trips per layer)
    int relu;
    int tf32;            // 1: A and B are fp32 (kind::tf32, 32 elements per slab); 0: 16-bit operands (64 per slab)
    int fmt;             // operand format of the instruction descriptor: 0 = fp16, 1 = bf16, 2 = tf32
    const float* bias;   // [N] zero padded
};

struct AeParams {
    CUtensorMap tmap_x;
    CUtensorMap tmap_w[OLS_AE_MAX_LAYERS];
    AeLayer layer[OLS_AE_MAX_LAYERS];
    int n_layers;
    int manual_x;    // 1: layer-0 input is loaded by the epilogue warps (row stride not TMA compatible)
    int l0_half;     // 1: layer 0 runs kind::f16 on fp16 operands: W0 is stored in fp16 (half the bytes re-streamed per tile)
                     //    and the fp32 x tile, staged by TMA in the idle activation buffer, is converted by the epilogue warps
    int x_bf16;      // 1: the input matrix is bf16 (the HR module's last activation); layer 0 then runs kind::f16
    int K0_real;     // real input width
    int out_real;    // real output width
    int normalize;
    int n_stages, stage_bytes, act_bytes;
    int stage_out;   // 1: the last layer's output is transposed through shared memory (wide outputs; needs AE_STAGING_BYTES)
    const float* x;
    float* y;
    long long M;
    int n_tiles;
    unsigned long long* trace;  // development aid (OLS_AE_TRACE=1): globaltimer stamps of CTA 0's epilogue thread
};

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AE_THREADS, 1) k_ae_chain(const __grid_constant__ AeParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* ring = smem;                                          // n_stages * stage_bytes
    uint8_t* act = smem + (size_t)p.n_stages * p.stage_bytes;      // act_bytes: [K-slab][128 rows][128 B]
    float* staging = (float*)(act + p.act_bytes);                  // (stage_out) 4 warps x 32 x 33 floats: output transpose buffer
    uint64_t* bars = (uint64_t*)(act + p.act_bytes + (p.stage_out ? AE_STAGING_BYTES : 0));
    uint64_t* full = bars;                   // [n_stages] TMA -> MMA
    uint64_t* empty = bars + 8;              // [n_stages] MMA -> TMA
    uint64_t* mma_done = bars + 16;          // MMA -> epilogue (accumulator pass complete)
    uint64_t* epi_done = bars + 17;          // epilogue -> MMA (TMEM drained, next A operand in smem)
    uint32_t* tmem_slot = (uint32_t*)(bars + 18);
    uint64_t* xfull = bars + 20;             // [2] TMA -> converter (fp32 x staging slot landed)          (l0_half)
    uint64_t* xfree = bars + 22;             // [2] converter -> TMA (staging slot consumed)
    uint64_t* afull = bars + 24;             // [n_stages <= 4] converter -> MMA (fp16 A slab written into the ring stage)
    uint64_t* act_free = bars + 28;          // MMA -> TMA: the tile's last MMAs have read the activation buffer

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.n_stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(mma_done, 1);
        mbar_init(epi_done, 128);
        for (int s = 0; s < 2; s++) { mbar_init(&xfull[s], 1); mbar_init(&xfree[s], 128); }
        for (int s = 0; s < 4; s++) mbar_init(&afull[s], 128);
        mbar_init(act_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(AE_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmap_x) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            uint32_t xcount = 0, tiles_p = 0;   // l0_half: fp32 x staging slots issued so far; tiles started
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, tiles_p++) {
                for (int l = 0; l < p.n_layers; l++) {
                    const AeLayer& L = p.layer[l];
                    const int slab_elems = L.tf32 ? 32 : 64;
                    const bool half0 = (l == 0) && p.l0_half;
                    const bool load_x = (l == 0) && !p.manual_x && !p.l0_half;
                    for (int b = 0; b < L.n_blocks; b++) {
                        for (int s = 0; s < L.n_slabs; s += L.sps) {
                            const int cnt = min(L.sps, L.n_slabs - s);
                            mbar_wait(&empty[stage], phase ^ 1);
                            uint8_t* st = ring + (size_t)stage * p.stage_bytes;
                            const uint32_t bytes = (uint32_t)cnt * L.block_n * AE_SLAB_BYTES + (load_x ? AE_M * AE_SLAB_BYTES : 0);
                            mbar_expect_tx(&full[stage], bytes);
                            uint8_t* bdst = st;
                            if (load_x) {
                                tma_load_2d(st, &p.tmap_x, &full[stage], s * slab_elems, tile * AE_M);
                                bdst = st + AE_M * AE_SLAB_BYTES;
                            }
                            if (half0) {
                                // fp32 x slab (64 columns = two 128-byte boxes) into staging slot xs of the activation buffer;
                                // the buffer is free once the previous tile's last MMAs have read it
                                bdst = st + AE_M * AE_SLAB_BYTES;   // the stage's first 16 KB take the converted fp16 A slab
                                {
                                    const uint32_t xs = xcount & 1u;
                                    if (b == 0 && s == 0 && tiles_p > 0) mbar_wait(act_free, (tiles_p - 1) & 1u);
                                    mbar_wait(&xfree[xs], ((xcount >> 1) & 1u) ^ 1u);
                                    mbar_expect_tx(&xfull[xs], 2 * AE_M * AE_SLAB_BYTES);
                                    uint8_t* xdst = act + (size_t)xs * (2 * AE_M * AE_SLAB_BYTES);
                                    tma_load_2d(xdst, &p.tmap_x, &xfull[xs], s * 64, tile * AE_M);
                                    tma_load_2d(xdst + AE_M * AE_SLAB_BYTES, &p.tmap_x, &xfull[xs], s * 64 + 32, tile * AE_M);
                                    xcount++;
                                }
                            }
                            for (int q = 0; q < cnt; q++)
                                for (int c = 0; c < L.block_n; c += L.chunk_n)
                                    tma_load_2d(bdst + ((size_t)q * L.block_n + c) * AE_SLAB_BYTES, &p.tmap_w[l], &full[stage],
                                                (s + q) * slab_elems, b * L.block_n + c);
                            if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        int stage = 0;
        uint32_t phase = 0, epi_phase = 0;
        int tiles_done = 0;
        uint32_t a_uses[4] = {0u, 0u, 0u, 0u};   // l0_half: how often each ring stage has carried a converted A slab
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, tiles_done++) {
            for (int l = 0; l < p.n_layers; l++) {
                const AeLayer& L = p.layer[l];
                const bool a_from_ring = (l == 0) && !p.manual_x;
                const bool half0 = (l == 0) && p.l0_half;
                const uint32_t idesc = make_idesc_fmt((uint32_t)L.fmt, L.chunk_n);
                for (int b = 0; b < L.n_blocks; b++) {
                    const int in_pass = b % L.blocks_per_pass;
                    if (in_pass == 0) {
                        // the previous accumulator pass must be drained (and, for l > 0, the A operand written)
                        if (l == 0 && b == 0) {
                            if (tiles_done > 0) { mbar_wait(epi_done, epi_phase); epi_phase ^= 1; }  // last pass of the previous tile
                            if (p.manual_x) { mbar_wait(epi_done, epi_phase); epi_phase ^= 1; }       // hand-written A operand
                        } else {
                            mbar_wait(epi_done, epi_phase);
                            epi_phase ^= 1;
                        }
                        tcgen05_fence_after();
                    }
                    for (int s = 0; s < L.n_slabs; s += L.sps) {
                        const int cnt = min(L.sps, L.n_slabs - s);
                        mbar_wait(&full[stage], phase);
                        if (half0) {   // the fp16 A slab of this stage has been written by the converter warps
                            mbar_wait(&afull[stage], a_uses[stage] & 1u);
                            a_uses[stage]++;
                        }
                        tcgen05_fence_after();
                        if (lane == 0) {
                            uint8_t* st = ring + (size_t)stage * p.stage_bytes;
                            for (int q = 0; q < cnt; q++) {   // the K-slabs this stage carries
                                const uint32_t a_addr = a_from_ring ? smem_u32(st) : smem_u32(act + (size_t)(s + q) * AE_M * AE_SLAB_BYTES);
                                const uint32_t b_addr = smem_u32(st) + (a_from_ring ? AE_M * AE_SLAB_BYTES : 0) +
                                                        (uint32_t)q * L.block_n * AE_SLAB_BYTES;
#pragma unroll
                                for (int k = 0; k < 4; k++) {  // 4 x 32 bytes of K per slab
                                    const uint64_t ad = make_sdesc(a_addr + k * 32);
                                    for (int c = 0; c < L.block_n; c += L.chunk_n) {
                                        const uint64_t bd = make_sdesc(b_addr + (uint32_t)c * AE_SLAB_BYTES + k * 32);
                                        const uint32_t d = tmem_base + (uint32_t)(in_pass * L.block_n + c);
                                        if (L.tf32) umma<true>(d, ad, bd, idesc, ((s + q) | k) ? 1u : 0u);
                                        else umma<false>(d, ad, bd, idesc, ((s + q) | k) ? 1u : 0u);
                                    }
                                }
                            }
                            umma_commit(&empty[stage]);  // frees the ring slot when these MMAs have read it
                        }
                        __syncwarp();
                        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
                    }
                    if (in_pass == L.blocks_per_pass - 1 || b == L.n_blocks - 1) {
                        if (lane == 0) {
                            umma_commit(mma_done);
                            // the tile's last MMAs: once they complete, the activation buffer may stage the next tile's x
                            if (p.l0_half && l == p.n_layers - 1 && b == L.n_blocks - 1) umma_commit(act_free);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else {
        // ===================== epilogue warps (one thread = one row = one TMEM lane) =====================
        const int quad = warp & 3;                 // TMEM lanes [32*quad, 32*quad + 32) are accessible to this warp
        const int row = quad * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        uint32_t done_phase = 0;
        const bool tracer = p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 64;
        int tr_n = 1;
        if (tracer) p.trace[tr_n++] = gtimer();
        uint32_t xcount = 0, ring_slabs = 0;     // l0_half: staging slots consumed; ring slabs seen (to follow the producer's stage / phase)
        uint32_t a_uses_c[4] = {0u, 0u, 0u, 0u};
        (void)a_uses_c;
        int cstage = 0;
        uint32_t cphase = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const long long grow = (long long)tile * AE_M + row;
            if (p.l0_half) {
                // layer-0 A operand: the fp32 x slab staged by TMA (two SWIZZLE_128B boxes of 32 columns) -> fp16, K-major,
                // SWIZZLE_128B, into the first 16 KB of the ring stage whose W slab the producer loads in parallel.
                // The ring stage sequence of layer 0 is followed with the same (stage, phase) arithmetic as the producer's.
                const AeLayer& L0 = p.layer[0];
                for (int b = 0; b < L0.n_blocks; b++) {
                    for (int s = 0; s < L0.n_slabs; s++) {
                        {
                            const uint32_t xs = xcount & 1u;
                            mbar_wait(&xfull[xs], (xcount >> 1) & 1u);           // fp32 slab landed
                            mbar_wait(&empty[cstage], cphase ^ 1u);              // ring stage free (MMAs of its previous use done)
                            const uint8_t* xsrc = act + (size_t)xs * (2 * AE_M * AE_SLAB_BYTES);
                            uint8_t* adst = ring + (size_t)cstage * p.stage_bytes;
#pragma unroll
                            for (int c = 0; c < 8; c++) {      // 16-byte chunk c of the fp16 row = columns 8c .. 8c+7
                                const uint8_t* box = xsrc + (size_t)(c >> 2) * (AE_M * AE_SLAB_BYTES);
                                const float4 lo = *(const float4*)(box + sw128(row, 2 * (c & 3)));
                                const float4 hi = *(const float4*)(box + sw128(row, 2 * (c & 3) + 1));
                                const __half2 h0 = __floats2half2_rn(lo.x, lo.y), h1 = __floats2half2_rn(lo.z, lo.w);
                                const __half2 h2 = __floats2half2_rn(hi.x, hi.y), h3 = __floats2half2_rn(hi.z, hi.w);
                                uint4 pk;
                                pk.x = *(const uint32_t*)&h0; pk.y = *(const uint32_t*)&h1; pk.z = *(const uint32_t*)&h2; pk.w = *(const uint32_t*)&h3;
                                *(uint4*)(adst + sw128(row, c)) = pk;
                            }
                            fence_proxy_async_smem();
                            mbar_arrive(&afull[cstage]);
                            mbar_arrive(&xfree[xs]);
                            xcount++;
                        }
                        if (++cstage == p.n_stages) { cstage = 0; cphase ^= 1u; }
                    }
                }
                // the remaining layers' slabs advance the ring too
                for (int l = 1; l < p.n_layers; l++) {
                    const int adv = p.layer[l].n_blocks * ((p.layer[l].n_slabs + p.layer[l].sps - 1) / p.layer[l].sps);
                    for (int q = 0; q < adv; q++)
                        if (++cstage == p.n_stages) { cstage = 0; cphase ^= 1u; }
                }
                (void)ring_slabs;
            }
            if (p.manual_x) {
                // layer-0 A operand written by hand: fp32, zero padded to the slab width
                const AeLayer& L0 = p.layer[0];
                for (int s = 0; s < L0.n_slabs; s++) {
                    for (int j = 0; j < 8; j++) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        float* vv = &v.x;
                        for (int e = 0; e < 4; e++) {
                            const int col = s * 32 + j * 4 + e;
                            if (col < p.K0_real && grow < p.M) vv[e] = p.x[grow * p.K0_real + col];
                        }
                        *(float4*)(act + (size_t)s * AE_M * AE_SLAB_BYTES + sw128(row, j)) = v;
                    }
                }
                fence_proxy_async_smem();
                mbar_arrive(epi_done);
            }
            for (int l = 0; l < p.n_layers; l++) {
                const AeLayer& L = p.layer[l];
                const bool last = l == p.n_layers - 1;
                const int pass_w = L.block_n * L.blocks_per_pass;
                float sumsq = 0.0f;
                for (int n0 = 0; n0 < L.N; n0 += pass_w) {
                    const int w = min(pass_w, L.N - n0);
                    mbar_wait(mma_done, done_phase);
                    done_phase ^= 1;
                    tcgen05_fence_after();
                    if (tracer && tr_n < 250) p.trace[tr_n++] = gtimer();  // accumulator of (layer l, pass) ready
                    // TMEM -> registers is double buffered: the load of chunk c + 1 is in flight while chunk c gets its
                    // bias / ReLU / bf16 packing (tcgen05.wait::ld waits for every outstanding load of the thread)
                    auto issue_ld = [&](int c, uint32_t (&r)[32]) {
                        if (w - c >= 32) tmem_ld32(t_lane + (uint32_t)c, r);
                        else {
                            tmem_ld16(t_lane + (uint32_t)c, r);
#pragma unroll
                            for (int i = 16; i < 32; i++) r[i] = 0u;
                        }
                    };
                    // Row normalisation (model.py:55,61) without a read-modify-write of the output: in the LAST pass of the last
                    // layer the accumulator is swept twice -- first only for the sum of squares (TMEM reads are cheap), then for
                    // the scaled, final values.  Earlier passes of a wide last layer (the decoder's 768 columns = 2 passes) have
                    // to leave before the norm is known: they are written raw and rescaled once at the end.
                    const bool final_pass = last && (n0 + pass_w >= L.N);
                    float scale = 1.0f;
                    if (final_pass && p.normalize) {
                        uint32_t rs[32];
                        for (int c = 0; c < w; c += 32) {
                            const int nc = min(32, w - c);
                            if (w - c >= 32) tmem_ld32(t_lane + (uint32_t)c, rs);
                            else tmem_ld16(t_lane + (uint32_t)c, rs);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; i++) {
                                const int col = n0 + c + i;
                                if (i < nc && col < p.out_real) {
                                    const float t = __uint_as_float(rs[i]) + __ldg(L.bias + col);
                                    sumsq = fmaf(t, t, sumsq);
                                }
                            }
                        }
                        scale = 1.0f / sqrtf(sumsq);
                    }
                    const bool acc_sq = last && !final_pass;   // raw passes contribute to the norm as they go
                    auto process = [&](int c, uint32_t (&r)[32]) {
                        const int nc = min(32, w - c);
                        float v[32];
                        const float4* b4 = reinterpret_cast<const float4*>(L.bias + n0 + c);  // N is a multiple of 16
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            const float4 bq = (q * 4 < nc) ? __ldg(b4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                            v[4 * q + 0] = __uint_as_float(r[4 * q + 0]) + bq.x;
                            v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bq.y;
                            v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bq.z;
                            v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bq.w;
                        }
                        if (L.relu) {
#pragma unroll
                            for (int i = 0; i < 32; i++) v[i] = fmaxf(v[i], 0.0f);
                        }
                        if (last) {
#pragma unroll
                            for (int i = 0; i < 32; i++) {
                                if (acc_sq && i < nc && n0 + c + i < p.out_real) sumsq = fmaf(v[i], v[i], sumsq);
                                v[i] *= scale;
                            }
                        }
                        if (!last) {
                            // next layer's A operand: bf16, K-major, 128-byte swizzle; column col -> slab col/64
                            const int col0 = n0 + c;
                            uint8_t* slab = act + (size_t)(col0 >> 6) * AE_M * AE_SLAB_BYTES;
                            const int j0 = (col0 & 63) >> 3;  // first 16-byte chunk (8 bf16) inside the slab row
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                uint4 pk;
                                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[q * 8 + 0], v[q * 8 + 1]);
                                __nv_bfloat162 h1 = __floats2bfloat162_rn(v[q * 8 + 2], v[q * 8 + 3]);
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[q * 8 + 4], v[q * 8 + 5]);
                                __nv_bfloat162 h3 = __floats2bfloat162_rn(v[q * 8 + 6], v[q * 8 + 7]);
                                pk.x = *(uint32_t*)&h0; pk.y = *(uint32_t*)&h1; pk.z = *(uint32_t*)&h2; pk.w = *(uint32_t*)&h3;
                                if (q * 8 < nc) *(uint4*)(slab + sw128(row, j0 + q)) = pk;
                            }
                        } else if (!p.stage_out) {
                            // narrow outputs (the 15-dim code): a warp's 32 rows are one contiguous block of global memory
#pragma unroll
                            for (int i = 0; i < 32; i++) {
                                const int col = n0 + c + i;
                                if (i < nc && col < p.out_real && grow < p.M) p.y[grow * p.out_real + col] = v[i];
                            }
                        } else {
                            // A thread owns a row, so writing its values directly would scatter 4-byte stores one row stride
                            // apart.  The warp's 32 x 32 chunk is transposed through shared memory instead and leaves as
                            // one contiguous run of up to 128 bytes per row.
                            float* stg = staging + quad * (32 * 33);
#pragma unroll
                            for (int i = 0; i < 32; i++) stg[lane * 33 + i] = v[i];
                            __syncwarp();
                            const long long row0 = (long long)tile * AE_M + quad * 32;
                            const int col = n0 + c + lane;
                            if (lane < nc && col < p.out_real) {
#pragma unroll 4
                                for (int r = 0; r < 32; r++)
                                    if (row0 + r < p.M) p.y[(row0 + r) * p.out_real + col] = stg[r * 33 + lane];
                            }
                            __syncwarp();
                        }
                    };
                    {
                        uint32_t ra[32], rb[32];
                        issue_ld(0, ra);
                        for (int c = 0; c < w; c += 64) {
                            tmem_ld_wait();
                            if (c + 32 < w) issue_ld(c + 32, rb);
                            process(c, ra);
                            if (c + 32 < w) {
                                tmem_ld_wait();
                                if (c + 64 < w) issue_ld(c + 64, ra);
                                process(c + 32, rb);
                            }
                        }
                    }
                    if (!last) {
                        // zero the K padding of the next layer's A operand (its K may exceed this layer's N)
                        const AeLayer& Ln = p.layer[l + 1];
                        for (int col0 = L.N; col0 < Ln.K; col0 += 8)
                            *(uint4*)(act + (size_t)(col0 >> 6) * AE_M * AE_SLAB_BYTES + sw128(row, (col0 & 63) >> 3)) =
                                make_uint4(0u, 0u, 0u, 0u);
                    }
                    tcgen05_fence_before();
                    if (!last) fence_proxy_async_smem();
                    mbar_arrive(epi_done);
                    if (tracer && tr_n < 250) p.trace[tr_n++] = gtimer();  // epilogue of (layer l, pass) done
                }
                const int raw_cols = min(p.out_real, ((L.N - 1) / pass_w) * pass_w);   // columns written before the norm was known
                if (last && p.normalize && raw_cols > 0) {
                    // the warp's 32 rows x raw_cols raw values were just written by its own lanes (still in L2): rescale them
                    // with coalesced accesses, four rows of independent loads in flight, lane r supplying row r's 1 / norm
                    __syncwarp();
                    const float inv = 1.0f / sqrtf(sumsq);
                    const long long row0 = (long long)tile * AE_M + quad * 32;
                    const bool vec4 = (p.out_real & 3) == 0 && (raw_cols & 3) == 0 && (((uintptr_t)p.y) & 15) == 0;
                    if (vec4) {
                        const int nq = raw_cols >> 2;
                        for (int r0 = 0; r0 < 32; r0 += 4) {
                            float invs[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) invs[u] = __shfl_sync(0xffffffffu, inv, r0 + u);
                            for (int q = lane; q < nq; q += 32) {
                                float4 t[4];
#pragma unroll
                                for (int u = 0; u < 4; u++)
                                    if (row0 + r0 + u < p.M) t[u] = __ldcg(reinterpret_cast<const float4*>(p.y + (row0 + r0 + u) * p.out_real) + q);
#pragma unroll
                                for (int u = 0; u < 4; u++)
                                    if (row0 + r0 + u < p.M) {
                                        t[u].x *= invs[u]; t[u].y *= invs[u]; t[u].z *= invs[u]; t[u].w *= invs[u];
                                        reinterpret_cast<float4*>(p.y + (row0 + r0 + u) * p.out_real)[q] = t[u];
                                    }
                            }
                        }
                    } else {
                        for (int r = 0; r < 32; r++) {
                            const float inv_r = __shfl_sync(0xffffffffu, inv, r);
                            if (row0 + r >= p.M) break;
                            float* yr = p.y + (row0 + r) * p.out_real;
                            for (int col = lane; col < raw_cols; col += 32) yr[col] *= inv_r;
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }
    if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 64) p.trace[0] = 0xffffffffull;  // marks a finished trace
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(AE_TMEM_COLS));
    }
}

// ---------------------------------------------------------------------------------------------------
// Parity mode (OLS_AE_FP32): the same chain in plain fp32 FMAs -- the arithmetic of the reference's fp32 nn.Linear
// (language/autoencoder/model.py:52-62) up to summation order.  A CTA owns 32 rows; the activations of the current and
// the next layer live in shared memory ([row][k], padded); thread t produces output column n = t, t + 128, ... for all
// 32 rows (32 accumulators), reading the transposed weights Wt[k][n] coalesced and the activations as broadcasts.
// Not a fast path (about 10x the tensor-core kernel): it exists so that a caller can ask for reference-grade numbers.
// ---------------------------------------------------------------------------------------------------
constexpr int AF_ROWS = 32, AF_THREADS = 128;
struct AeFp32Params {
    int n_layers, normalize;
    int dims[OLS_AE_MAX_LAYERS + 1];
    const float* wt[OLS_AE_MAX_LAYERS];    // [K][N] transposed weights
    const float* bias[OLS_AE_MAX_LAYERS];  // [N] or NULL
    const float* x;
    float* y;
    long long M;
    int max_width;                          // widest activation (floats per row of the two shared-memory buffers)
};

__global__ void __launch_bounds__(AF_THREADS) k_ae_chain_fp32(const AeFp32Params p) {
    extern __shared__ float af_smem[];
    const int ld = p.max_width + 1;
    float* bufA = af_smem;
    float* bufB = af_smem + (size_t)AF_ROWS * ld;
    __shared__ float s_inv[AF_ROWS];
    const int tid = threadIdx.x;
    const long long n_tiles = (p.M + AF_ROWS - 1) / AF_ROWS;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = tile * AF_ROWS;
        const int rows = (int)min((long long)AF_ROWS, p.M - row0);
        const int K0 = p.dims[0];
        for (int e = tid; e < AF_ROWS * K0; e += AF_THREADS) {
            const int r = e / K0, k = e - r * K0;
            bufA[r * ld + k] = r < rows ? p.x[(row0 + r) * K0 + k] : 0.0f;
        }
        __syncthreads();
        float* in = bufA;
        float* out = bufB;
        for (int l = 0; l < p.n_layers; l++) {
            const int K = p.dims[l], N = p.dims[l + 1];
            const bool last = l == p.n_layers - 1;
            const float* __restrict__ wt = p.wt[l];
            for (int n = tid; n < N; n += AF_THREADS) {
                float acc[AF_ROWS];
                const float b = p.bias[l] ? p.bias[l][n] : 0.0f;
#pragma unroll
                for (int r = 0; r < AF_ROWS; r++) acc[r] = b;
                for (int k = 0; k < K; k++) {
                    const float w = __ldg(wt + (size_t)k * N + n);
#pragma unroll
                    for (int r = 0; r < AF_ROWS; r++) acc[r] = fmaf(in[r * ld + k], w, acc[r]);
                }
#pragma unroll
                for (int r = 0; r < AF_ROWS; r++) out[r * ld + n] = last ? acc[r] : fmaxf(acc[r], 0.0f);
            }
            __syncthreads();
            float* t = in; in = out; out = t;
        }
        const int N = p.dims[p.n_layers];
        if (tid < AF_ROWS) {
            float ss = 0.0f;
            for (int n = 0; n < N; n++) ss = fmaf(in[tid * ld + n], in[tid * ld + n], ss);
            s_inv[tid] = p.normalize ? 1.0f / sqrtf(ss) : 1.0f;
        }
        __syncthreads();
        for (int e = tid; e < rows * N; e += AF_THREADS) {
            const int r = e / N, n = e - r * N;
            p.y[(row0 + r) * N + n] = in[r * ld + n] * s_inv[r];
        }
        __syncthreads();
    }
}

__global__ void k_ae_transpose_weight(const float* __restrict__ w, int N, int K, float* __restrict__ wt) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)N * K) return;
    const int n = (int)(i / K), k = (int)(i % K);
    wt[(size_t)k * N + n] = w[i];
}

// weight re-layout: pad [N,K] fp32 -> [N_pad,K_pad] fp32 (layer 0) or bf16 (inner layers), zero filled
__global__ void k_ae_pack_weight(const float* __restrict__ w, int N, int K, void* __restrict__ out, int N_pad, int K_pad,
                                 int fmt /* 2 = fp32 (tf32 operand), 1 = bf16, 0 = fp16 */) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)N_pad * K_pad) return;
    const int n = (int)(i / K_pad), k = (int)(i % K_pad);
    const float v = (n < N && k < K) ? w[(size_t)n * K + k] : 0.0f;
    if (fmt == 1) ((__nv_bfloat16*)out)[i] = __float2bfloat16_rn(v);
    else if (fmt == 0) ((__half*)out)[i] = __float2half_rn(v);
    else ((float*)out)[i] = v;
}
__global__ void k_ae_pack_bias(const float* __restrict__ b, int N, float* __restrict__ out, int N_pad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N_pad) out[i] = (i < N && b) ? b[i] : 0.0f;
}

}  // namespace ols

using namespace ols;

struct ols_ae_plan {
    AeParams p;
    int fp32_mode = 0;
    AeFp32Params pf;
    std::vector<void*> owned;
    int device;
    int sm_count;
    size_t smem_bytes;
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)ptr;
    }
    return fn;
}

// 2-D row-major tensor [rows, cols] with a [box_rows, 128 bytes] box, 128-byte swizzle
static int make_map(CUtensorMap* map, const void* base, bool bf16, uint64_t rows, uint64_t cols, uint32_t box_rows, bool fp16 = false) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) { ols_set_error("cuTensorMapEncodeTiled not available"); return OLS_ERR_CUDA; }
    const uint32_t esz = (bf16 || fp16) ? 2 : 4;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * esz};
    cuuint32_t box[2] = {AE_SLAB_BYTES / esz, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : (bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32), 2, (void*)base, dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ols_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return OLS_ERR_CUDA; }
    return OLS_OK;
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

extern "C" {

int ols_ae_plan_create(const ols_ae_chain* chain, ols_ae_plan** out_plan, void* stream) {
    if (!chain || !out_plan) { ols_set_error("null argument"); return OLS_ERR_INVALID; }
    if (chain->n_layers < 1 || chain->n_layers > OLS_AE_MAX_LAYERS) { ols_set_error("n_layers out of range"); return OLS_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    ols_ae_plan* plan = new ols_ae_plan();
    AeParams& p = plan->p;
    memset(&p, 0, sizeof(p));
    auto fail = [&](int rc) { ols_ae_plan_destroy(plan); return rc; };
    if (cudaGetDevice(&plan->device) != cudaSuccess) { ols_set_error("no CUDA device"); return fail(OLS_ERR_CUDA); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, plan->device) != cudaSuccess) { ols_set_error("cudaGetDeviceProperties failed"); return fail(OLS_ERR_CUDA); }
    plan->sm_count = prop.multiProcessorCount;
    p.n_layers = chain->n_layers;
    p.K0_real = chain->dims[0];
    p.out_real = chain->dims[chain->n_layers];
    p.normalize = chain->normalize;
    p.x_bf16 = chain->input_bf16 ? 1 : 0;
    if (chain->precision == OLS_AE_FP32) {
        // parity mode: transposed fp32 weights, no tensor-core plan
        if (chain->input_bf16) { ols_set_error("the fp32 parity mode takes a float32 input"); return fail(OLS_ERR_UNSUPPORTED); }
        plan->fp32_mode = 1;
        AeFp32Params& f = plan->pf;
        memset(&f, 0, sizeof(f));
        f.n_layers = chain->n_layers; f.normalize = chain->normalize;
        for (int l = 0; l <= chain->n_layers; l++) { f.dims[l] = chain->dims[l]; if (chain->dims[l] > f.max_width) f.max_width = chain->dims[l]; }
        for (int l = 0; l < chain->n_layers; l++) {
            const int K = chain->dims[l], N = chain->dims[l + 1];
            if (K <= 0 || N <= 0 || !chain->d_weight[l]) { ols_set_error("bad layer %d", l); return fail(OLS_ERR_INVALID); }
            float* wt = nullptr; float* bb = nullptr;
            if (cudaMalloc((void**)&wt, sizeof(float) * (size_t)N * K) != cudaSuccess) { ols_set_error("out of device memory"); return fail(OLS_ERR_CUDA); }
            plan->owned.push_back(wt);
            k_ae_transpose_weight<<<(unsigned)(((size_t)N * K + 255) / 256), 256, 0, st>>>(chain->d_weight[l], N, K, wt);
            f.wt[l] = wt;
            if (chain->d_bias[l]) {
                if (cudaMalloc((void**)&bb, sizeof(float) * N) != cudaSuccess) { ols_set_error("out of device memory"); return fail(OLS_ERR_CUDA); }
                plan->owned.push_back(bb);
                cudaMemcpyAsync(bb, chain->d_bias[l], sizeof(float) * N, cudaMemcpyDeviceToDevice, st);
            }
            f.bias[l] = bb;
        }
        if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) { ols_set_error("weight packing failed"); return fail(OLS_ERR_CUDA); }
        plan->smem_bytes = sizeof(float) * 2 * (size_t)AF_ROWS * (f.max_width + 1);
        // (static + dynamic shared memory share the 227 KB opt-in limit: reserve what this kernel can ever need, once)
        if (plan->smem_bytes > 226 * 1024 ||
            cudaFuncSetAttribute(k_ae_chain_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) != cudaSuccess) {
            cudaGetLastError();
            ols_set_error("layer too wide for the fp32 parity kernel"); return fail(OLS_ERR_UNSUPPORTED);
        }
        *out_plan = plan;
        return OLS_OK;
    }
    p.manual_x = (chain->dims[0] % 32 != 0) ? 1 : 0;  // TMA needs 16-byte row strides; keep whole slabs too
    if (p.x_bf16 && chain->dims[0] % 64 != 0) { ols_set_error("bf16 input needs a width that is a multiple of 64"); return fail(OLS_ERR_UNSUPPORTED); }
    // fp16 first layer (same 10-bit mantissa as tf32, half the weight bytes re-streamed per tile, twice the MMA rate): needs
    // a TMA-loadable fp32 input of whole 64-column slabs and an activation buffer that can stage two fp32 x slabs (64 KB),
    // i.e. a second layer of at least 256 inputs.  OLS_AE_L0_TF32=1 keeps the tf32 first layer (A/B aid).
    static const bool force_tf32 = getenv("OLS_AE_L0_TF32") != nullptr;
    p.l0_half = (!force_tf32 && !p.manual_x && !p.x_bf16 && chain->dims[0] % 64 == 0 && chain->n_layers >= 2 &&
                 round_up(chain->dims[1], 64) * AE_M * 2 >= 2 * 2 * AE_M * AE_SLAB_BYTES) ? 1 : 0;
    int stage_bytes = 0, act_cols_bytes = 0;
    for (int l = 0; l < chain->n_layers; l++) {
        const int K = chain->dims[l], N = chain->dims[l + 1];
        if (K <= 0 || N <= 0 || !chain->d_weight[l]) { ols_set_error("bad layer %d", l); return fail(OLS_ERR_INVALID); }
        AeLayer& L = p.layer[l];
        L.tf32 = (l == 0 && !p.x_bf16 && !p.l0_half) ? 1 : 0;
        L.fmt = L.tf32 ? 2 : ((l == 0 && p.l0_half) ? 0 : 1);
        const int slab = L.tf32 ? 32 : 64;
        // K of layer l must equal the padded N of layer l-1 (the activation the epilogue wrote)
        L.K = round_up(K, slab);
        L.N = round_up(N, 16);
        L.n_slabs = L.K / slab;
        // split N into equal blocks of at most AE_BLOCK_MAX rows, each a multiple of 16
        int nb = (L.N + AE_BLOCK_MAX - 1) / AE_BLOCK_MAX;
        while ((L.N % nb) != 0 || ((L.N / nb) % 16) != 0) nb++;
        L.n_blocks = nb;
        L.block_n = L.N / nb;
        L.chunk_n = L.block_n;
        if (L.chunk_n > AE_CHUNK_MAX) {
            if ((L.block_n / 2) % 16 != 0) { ols_set_error("layer %d: cannot split N=%d", l, L.N); return fail(OLS_ERR_UNSUPPORTED); }
            L.chunk_n = L.block_n / 2;
        }
        if (L.n_blocks > AE_MAX_BLOCKS) { ols_set_error("layer %d too wide (N=%d)", l, N); return fail(OLS_ERR_UNSUPPORTED); }
        L.blocks_per_pass = AE_TMEM_COLS / L.block_n;
        if (L.blocks_per_pass > L.n_blocks) L.blocks_per_pass = L.n_blocks;
        const bool last = l == chain->n_layers - 1;
        if (!last && L.blocks_per_pass != L.n_blocks) { ols_set_error("hidden layer %d wider than 512", l); return fail(OLS_ERR_UNSUPPORTED); }
        L.relu = last ? 0 : 1;
        const int sb = L.block_n * AE_SLAB_BYTES + ((l == 0 && !p.manual_x) ? AE_M * AE_SLAB_BYTES : 0);
        if (sb > stage_bytes) stage_bytes = sb;
        if (l > 0 || p.manual_x) { const int ab = L.n_slabs * AE_M * AE_SLAB_BYTES; if (ab > act_cols_bytes) act_cols_bytes = ab; }
        // packed weights and bias
        void* wbuf = nullptr; float* bbuf = nullptr;
        const size_t wbytes = (size_t)L.N * L.K * (L.tf32 ? 4 : 2);
        if (cudaMalloc(&wbuf, wbytes) != cudaSuccess || cudaMalloc((void**)&bbuf, sizeof(float) * L.N) != cudaSuccess) {
            if (wbuf) cudaFree(wbuf);
            ols_set_error("out of device memory"); return fail(OLS_ERR_CUDA);
        }
        plan->owned.push_back(wbuf); plan->owned.push_back(bbuf);
        const size_t tot = (size_t)L.N * L.K;
        k_ae_pack_weight<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(chain->d_weight[l], N, K, wbuf, L.N, L.K, L.fmt);
        k_ae_pack_bias<<<(L.N + 255) / 256, 256, 0, st>>>(chain->d_bias[l], N, bbuf, L.N);
        L.bias = bbuf;
        int rc = make_map(&p.tmap_w[l], wbuf, L.fmt == 1, (uint64_t)L.N, (uint64_t)L.K, (uint32_t)L.chunk_n, L.fmt == 0);
        if (rc != OLS_OK) return fail(rc);
    }
    for (int l = 1; l < chain->n_layers; l++) {
        if (p.layer[l].K < p.layer[l - 1].N) { ols_set_error("internal: K/N padding mismatch at layer %d", l); return fail(OLS_ERR_INVALID); }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) { ols_set_error("weight packing failed"); return fail(OLS_ERR_CUDA); }
    stage_bytes = round_up(stage_bytes, 1024);
    for (int l = 0; l < chain->n_layers; l++) {
        AeLayer& L = p.layer[l];
        const int slab_bytes = L.block_n * AE_SLAB_BYTES;
        L.sps = (l == 0 && !p.manual_x) ? 1 : stage_bytes / slab_bytes;   // layer 0's stage also holds the A slab
        if (L.sps < 1) L.sps = 1;
        if (L.sps > L.n_slabs) L.sps = L.n_slabs;
    }
    const int act_bytes = round_up(act_cols_bytes > 0 ? act_cols_bytes : 1024, 1024);
    const int budget = 227 * 1024 - 1024 /*alignment*/ - 256 /*barriers*/ - act_bytes;
    int n_stages = budget / stage_bytes;
    if (n_stages > 8) n_stages = 8;
    if (p.l0_half && n_stages > 4) n_stages = 4;   // afull[] has four barriers
    // wide outputs leave through a shared-memory transpose when it fits beside the ring (it does for every decoder of the
    // reference; the encoders write 15- / 32-float rows, which are contiguous per warp anyway)
    p.stage_out = (p.out_real > 32 && n_stages >= 2 && budget - n_stages * stage_bytes >= AE_STAGING_BYTES) ? 1 : 0;
    if (p.out_real > 32 && !p.stage_out && n_stages > 2) { n_stages--; p.stage_out = 1; }
    if (n_stages < 2) { ols_set_error("layer chain does not fit shared memory (stage %d B, act %d B)", stage_bytes, act_bytes); return fail(OLS_ERR_UNSUPPORTED); }
    p.n_stages = n_stages; p.stage_bytes = stage_bytes; p.act_bytes = act_bytes;
    plan->smem_bytes = (size_t)n_stages * stage_bytes + act_bytes + (p.stage_out ? AE_STAGING_BYTES : 0) + 256 + 1024;
    // the attribute belongs to the kernel, not to the plan: always reserve the opt-in maximum so that plans with
    // different shared-memory needs (encoder / decoder) can be launched in any order
    if (plan->smem_bytes > 227 * 1024 ||
        cudaFuncSetAttribute(k_ae_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
        ols_set_error("cannot reserve %zu bytes of shared memory", plan->smem_bytes); return fail(OLS_ERR_CUDA);
    }
    *out_plan = plan;
    return OLS_OK;
}

void ols_ae_plan_destroy(ols_ae_plan* plan) {
    if (!plan) return;
    for (void* q : plan->owned) cudaFree(q);
    delete plan;
}

static int ae_forward(const ols_ae_plan* plan, const void* d_x, float* d_y, int64_t M, void* stream);

int ols_ae_forward(const ols_ae_plan* plan, const float* d_x, float* d_y, int64_t M, void* stream) {
    if (plan && !plan->fp32_mode && plan->p.x_bf16) { ols_set_error("this plan takes a bf16 input: call ols_ae_forward_bf16"); return OLS_ERR_INVALID; }
    return ae_forward(plan, d_x, d_y, M, stream);
}

int ols_ae_forward_bf16(const ols_ae_plan* plan, const void* d_x_bf16, float* d_y, int64_t M, void* stream) {
    if (plan && (plan->fp32_mode || !plan->p.x_bf16)) { ols_set_error("this plan takes a float32 input: call ols_ae_forward"); return OLS_ERR_INVALID; }
    return ae_forward(plan, d_x_bf16, d_y, M, stream);
}

}  // extern "C"

static int ae_forward(const ols_ae_plan* plan, const void* d_x, float* d_y, int64_t M, void* stream) {
    if (!plan || !d_x || !d_y || M < 0) { ols_set_error("bad arguments"); return OLS_ERR_INVALID; }
    if (M == 0) return OLS_OK;
    if (plan->fp32_mode) {
        AeFp32Params f = plan->pf;
        f.x = (const float*)d_x; f.y = d_y; f.M = M;
        const long long tiles = (M + AF_ROWS - 1) / AF_ROWS;
        const int grid = (int)(tiles < 2LL * plan->sm_count ? tiles : 2LL * plan->sm_count);
        ols_timing_mark(-1, (cudaStream_t)stream);
        k_ae_chain_fp32<<<grid, AF_THREADS, plan->smem_bytes, (cudaStream_t)stream>>>(f);
        OLS_CUDA_TRY(cudaGetLastError());
        ols_timing_mark(OLS_T_AE, (cudaStream_t)stream);
        return OLS_OK;
    }
    AeParams p = plan->p;
    p.x = (const float*)d_x; p.y = d_y; p.M = M;
    p.n_tiles = (int)((M + AE_M - 1) / AE_M);
    if (!p.manual_x) {
        if (((uintptr_t)d_x & 15) != 0) { ols_set_error("x must be 16-byte aligned"); return OLS_ERR_INVALID; }
        int rc = make_map(&p.tmap_x, d_x, p.x_bf16 != 0, (uint64_t)M, (uint64_t)p.K0_real, AE_M);
        if (rc != OLS_OK) return rc;
    }
    // OLS_AE_MAX_CTAS: cap on the persistent grid, e.g. to leave SMs to a collective that overlaps the encodes
    static const int max_ctas = getenv("OLS_AE_MAX_CTAS") ? atoi(getenv("OLS_AE_MAX_CTAS")) : 0;
    int grid = p.n_tiles < plan->sm_count ? p.n_tiles : plan->sm_count;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    static const bool trace_on = getenv("OLS_AE_TRACE") != nullptr;  // development aid: phase timeline of CTA 0 on stderr
    p.trace = nullptr;
    if (trace_on) {
        static unsigned long long* d_trace = nullptr;
        if (!d_trace) cudaMalloc(&d_trace, 256 * sizeof(unsigned long long));
        cudaMemsetAsync(d_trace, 0, 256 * sizeof(unsigned long long), (cudaStream_t)stream);
        p.trace = d_trace;
    }
    ols_timing_mark(-1, (cudaStream_t)stream);
    k_ae_chain<<<grid, AE_THREADS, plan->smem_bytes, (cudaStream_t)stream>>>(p);
    OLS_CUDA_TRY(cudaGetLastError());
    ols_timing_mark(OLS_T_AE, (cudaStream_t)stream);
    if (trace_on) {
        unsigned long long h[256];
        cudaStreamSynchronize((cudaStream_t)stream);
        cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[ae trace] M=%lld tiles=%d grid=%d stages=%d:", (long long)M, p.n_tiles, grid, p.n_stages);
        for (int i = 2; i < 256 && h[i]; i++) fprintf(stderr, " %.2f", (double)(h[i] - h[1]) * 1e-3);
        fprintf(stderr, " us\n");
    }
    return OLS_OK;
}


