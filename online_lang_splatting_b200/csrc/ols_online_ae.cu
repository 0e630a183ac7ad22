// ols_online_ae.cu -- one training step of the online language autoencoder as ONE kernel.
//
// Reference: BackEnd.train_online_autoencoder (utils/slam_backend.py:266-323) on EncoderDecoderOnline
// (language/autoencoder/model.py:314-354):
//     comp  = normalize(W2 relu(W1 x + b1) + b2)                       32 -> 24 -> 15
//     recon = normalize(W4 relu(W3 comp + b3) + b4)                    15 -> 24 -> 32
//     loss  = l1_loss(recon, x) + 0.6 * (1 - cosine_similarity(recon, x, dim=1).mean())
//     loss.backward();  torch.optim.Adam(lr).step()
// over the [36864, 32] code map of a keyframe.  In torch that is ~45 launch-latency-bound kernels per step (and it runs
// once per keyframe plus twice per mapping iteration, :557-567,:640-648).  Here:
//   phase A  one thread per row: forward, loss terms and the four delta vectors, fp32 FMAs, weights in shared memory,
//            row vectors kept in shared memory (element-major, padded: conflict-free);
//   phase B  the CTA's partial weight gradients = delta^T * activation over its 128 rows, one (i, j) entry per thread and
//            pass, written to a per-CTA slab of global memory;
//   phase C  the last CTA to finish (ticket counter) sums the slabs in CTA order -- deterministic -- and applies Adam to
//            the 2351 parameters in place, bumps the device step counter and writes the loss.
// No tensor cores: 2 x 2256 MACs per row is far too small a contraction per byte, and the reference arithmetic is fp32.
#include "ols_common.cuh"

namespace ols {

constexpr int OA_D0 = 32, OA_D1 = 24, OA_D2 = 15;
constexpr int OA_W1 = 0, OA_B1 = OA_W1 + OA_D1 * OA_D0, OA_W2 = OA_B1 + OA_D1, OA_B2 = OA_W2 + OA_D2 * OA_D1,
              OA_W3 = OA_B2 + OA_D2, OA_B3 = OA_W3 + OA_D1 * OA_D2, OA_W4 = OA_B3 + OA_D1, OA_B4 = OA_W4 + OA_D0 * OA_D1,
              OA_NPARAM = OA_B4 + OA_D0;  // 2351, in nn.Module.parameters() order
constexpr int OA_ROWS = 128, OA_THREADS = 128, OA_LD = OA_ROWS + 1;  // row vectors: [element][row], leading dimension 129
constexpr int OA_D2P = 16;  // 15 padded to whole float4 groups
// shared-memory row vectors (floats per row): x 32 | h1 24 | c 16 | h3 24 | dr 32 | dh3 24 | dz 16 | dh1 24
constexpr int OA_X = 0, OA_H1 = OA_X + OA_D0, OA_C = OA_H1 + OA_D1, OA_H3 = OA_C + OA_D2P, OA_DR = OA_H3 + OA_D1,
              OA_DH3 = OA_DR + OA_D0, OA_DZ = OA_DH3 + OA_D1, OA_DH1 = OA_DZ + OA_D2P, OA_VEC = OA_DH1 + OA_D1;  // 192
// shared-memory weights: transposed copies ([k][n], n contiguous) for the forward products, the original row-major
// matrices ([n][k], k contiguous) for the transposed products of the backward, rows padded to whole float4 groups
constexpr int OA_SW1T = 0, OA_SW2T = OA_SW1T + OA_D0 * OA_D1, OA_SW3T = OA_SW2T + OA_D1 * OA_D2P, OA_SW4T = OA_SW3T + OA_D2P * OA_D1,
              OA_SW2 = OA_SW4T + OA_D1 * OA_D0, OA_SW3 = OA_SW2 + OA_D2P * OA_D1, OA_SW4 = OA_SW3 + OA_D1 * OA_D2P,
              OA_SB1 = OA_SW4 + OA_D0 * OA_D1, OA_SB2 = OA_SB1 + OA_D1, OA_SB3 = OA_SB2 + OA_D2P, OA_SB4 = OA_SB3 + OA_D1,
              OA_SWTOT = OA_SB4 + OA_D0;
constexpr size_t OA_SMEM = sizeof(float) * ((size_t)OA_SWTOT + (size_t)OA_NPARAM + 1 + (size_t)OA_VEC * OA_LD + 8);

struct OnlineAeArgs {
    float *params, *m, *v;
    long long* step;
    const float* x;
    long long M;
    float lr, beta1, beta2, eps;
    float* code;        // [M, 15] or NULL
    float* loss;        // [1]
    float* partial;     // [grid, OA_NPARAM + 2] per-CTA gradient slabs (+ l1 sum, cos sum)
    unsigned int* ticket;
};

// out[n] = (RELU ? max(.,0) : .)( bias[n] + sum_k Wt[k][n] * in[k] ) for this thread's row; Wt is [K][N] with N a multiple
// of 4 (16-byte aligned rows: one LDS.128 of weights + one LDS of the input feed four FMAs); in / out are row vectors
// in shared memory ([element][row], stride OA_LD).
template <int K, int N, bool RELU>
__device__ __forceinline__ void oa_layer(const float* __restrict__ Wt, const float* __restrict__ bias, const float* in, float* out) {
#pragma unroll 1
    for (int n0 = 0; n0 < N; n0 += 4) {
        float s0 = bias ? bias[n0] : 0.0f, s1 = bias ? bias[n0 + 1] : 0.0f, s2 = bias ? bias[n0 + 2] : 0.0f, s3 = bias ? bias[n0 + 3] : 0.0f;
#pragma unroll 8
        for (int k = 0; k < K; k++) {
            const float xv = in[(size_t)k * OA_LD];
            const float4 w = *reinterpret_cast<const float4*>(Wt + k * N + n0);
            s0 = fmaf(w.x, xv, s0); s1 = fmaf(w.y, xv, s1); s2 = fmaf(w.z, xv, s2); s3 = fmaf(w.w, xv, s3);
        }
        if (RELU) { s0 = fmaxf(s0, 0.0f); s1 = fmaxf(s1, 0.0f); s2 = fmaxf(s2, 0.0f); s3 = fmaxf(s3, 0.0f); }
        out[(size_t)(n0 + 0) * OA_LD] = s0; out[(size_t)(n0 + 1) * OA_LD] = s1;
        out[(size_t)(n0 + 2) * OA_LD] = s2; out[(size_t)(n0 + 3) * OA_LD] = s3;
    }
}

__global__ void __launch_bounds__(OA_THREADS) k_online_ae_step(const OnlineAeArgs a) {
    extern __shared__ __align__(16) float oa_smem[];
    float* SW = oa_smem;                          // weights (see OA_S*)
    float* gacc = oa_smem + OA_SWTOT;             // this CTA's partial gradient (entry e is owned by thread e % 128)
    float* vec = gacc + OA_NPARAM + 1;            // row vectors
    float* red = vec + (size_t)OA_VEC * OA_LD;
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    for (int e = tid; e < OA_SWTOT; e += OA_THREADS) SW[e] = 0.0f;
    for (int e = tid; e < OA_NPARAM; e += OA_THREADS) gacc[e] = 0.0f;
    __syncthreads();
    for (int e = tid; e < OA_NPARAM; e += OA_THREADS) {
        const float w = a.params[e];
        if (e < OA_B1) { const int j = e / OA_D0, i = e % OA_D0; SW[OA_SW1T + i * OA_D1 + j] = w; }
        else if (e < OA_W2) SW[OA_SB1 + (e - OA_B1)] = w;
        else if (e < OA_B2) { const int k = (e - OA_W2) / OA_D1, j = (e - OA_W2) % OA_D1; SW[OA_SW2T + j * OA_D2P + k] = w; SW[OA_SW2 + k * OA_D1 + j] = w; }
        else if (e < OA_W3) SW[OA_SB2 + (e - OA_B2)] = w;
        else if (e < OA_B3) { const int j = (e - OA_W3) / OA_D2, k = (e - OA_W3) % OA_D2; SW[OA_SW3T + k * OA_D1 + j] = w; SW[OA_SW3 + j * OA_D2P + k] = w; }
        else if (e < OA_W4) SW[OA_SB3 + (e - OA_B3)] = w;
        else if (e < OA_B4) { const int i = (e - OA_W4) / OA_D1, j = (e - OA_W4) % OA_D1; SW[OA_SW4T + j * OA_D0 + i] = w; SW[OA_SW4 + i * OA_D1 + j] = w; }
        else SW[OA_SB4 + (e - OA_B4)] = w;
    }
    float l1_sum = 0.0f, cos_sum = 0.0f;
    const float inv_M = 1.0f / (float)a.M, inv_MD = 1.0f / ((float)a.M * (float)OA_D0);
    __syncthreads();

    const long long n_tiles = (a.M + OA_ROWS - 1) / OA_ROWS;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row = tile * OA_ROWS + tid;
        const bool live = row < a.M;
        float* my = vec + tid;   // element e of this thread's row vectors: my[(base + e) * OA_LD]
        // ---- phase A: one row per thread, every vector in shared memory ----
        {
            if (live) {
                const float4* src = reinterpret_cast<const float4*>(a.x + row * OA_D0);
#pragma unroll
                for (int q = 0; q < OA_D0 / 4; q++) {
                    const float4 t = src[q];
                    my[(OA_X + 4 * q) * OA_LD] = t.x; my[(OA_X + 4 * q + 1) * OA_LD] = t.y;
                    my[(OA_X + 4 * q + 2) * OA_LD] = t.z; my[(OA_X + 4 * q + 3) * OA_LD] = t.w;
                }
            } else {
                for (int i = 0; i < OA_D0; i++) my[(OA_X + i) * OA_LD] = 0.0f;
            }
            oa_layer<OA_D0, OA_D1, true>(SW + OA_SW1T, SW + OA_SB1, my + OA_X * OA_LD, my + OA_H1 * OA_LD);
            oa_layer<OA_D1, OA_D2P, false>(SW + OA_SW2T, SW + OA_SB2, my + OA_H1 * OA_LD, my + OA_C * OA_LD);   // z (padding column: 0)
            float zz = 0.0f;
#pragma unroll
            for (int k = 0; k < OA_D2; k++) { const float z = my[(OA_C + k) * OA_LD]; zz = fmaf(z, z, zz); }
            const float inv_z = 1.0f / sqrtf(zz);
#pragma unroll
            for (int k = 0; k < OA_D2; k++) {
                const float c = my[(OA_C + k) * OA_LD] * inv_z;
                my[(OA_C + k) * OA_LD] = c;
                if (live && a.code) a.code[row * OA_D2 + k] = c;
            }
            oa_layer<OA_D2P, OA_D1, true>(SW + OA_SW3T, SW + OA_SB3, my + OA_C * OA_LD, my + OA_H3 * OA_LD);
            oa_layer<OA_D1, OA_D0, false>(SW + OA_SW4T, SW + OA_SB4, my + OA_H3 * OA_LD, my + OA_DR * OA_LD);   // r
            float rr = 0.0f;
#pragma unroll
            for (int i = 0; i < OA_D0; i++) { const float r = my[(OA_DR + i) * OA_LD]; rr = fmaf(r, r, rr); }
            const float inv_r = 1.0f / sqrtf(rr);
            // loss terms (y = r / |r|)
            float xy = 0.0f, xx = 0.0f, yy = 0.0f, l1 = 0.0f;
#pragma unroll
            for (int i = 0; i < OA_D0; i++) {
                const float y = my[(OA_DR + i) * OA_LD] * inv_r, x = my[(OA_X + i) * OA_LD];
                my[(OA_DR + i) * OA_LD] = y;
                xy = fmaf(x, y, xy); xx = fmaf(x, x, xx); yy = fmaf(y, y, yy);
                l1 += fabsf(y - x);
            }
            // F.cosine_similarity: w12 / sqrt(clamp(w1 * w2, eps^2)), eps = 1e-8
            const float n12 = sqrtf(fmaxf(xx * yy, 1e-16f));
            const float cosv = xy / n12;
            if (live) { l1_sum += l1; cos_sum += cosv; }
            const float wc = live ? -0.6f * inv_M : 0.0f, wl = live ? inv_MD : 0.0f;
            const float inv_n12 = 1.0f / n12, c_yy = cosv / yy;
            float ydy = 0.0f;
#pragma unroll
            for (int i = 0; i < OA_D0; i++) {   // dL/dy_i = sign(y - x) / (M * 32) - 0.6 / M * (x_i / n12 - cos * y_i / yy)
                const float y = my[(OA_DR + i) * OA_LD], x = my[(OA_X + i) * OA_LD];
                const float d = y - x;
                const float sg = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f);
                ydy = fmaf(y, wl * sg + wc * (x * inv_n12 - c_yy * y), ydy);
            }
#pragma unroll
            for (int i = 0; i < OA_D0; i++) {   // through y = r / |r|: dr = (dy - y (y . dy)) / |r|
                const float y = my[(OA_DR + i) * OA_LD], x = my[(OA_X + i) * OA_LD];
                const float d = y - x;
                const float sg = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f);
                const float dy = wl * sg + wc * (x * inv_n12 - c_yy * y);
                my[(OA_DR + i) * OA_LD] = (dy - y * ydy) * inv_r;
            }
            // dh3 = relu'(h3) * W4^T dr;  dc = W3^T dh3;  dz = (dc - c (c . dc)) / |z|;  dh1 = relu'(h1) * W2^T dz
            oa_layer<OA_D0, OA_D1, false>(SW + OA_SW4, nullptr, my + OA_DR * OA_LD, my + OA_DH3 * OA_LD);
#pragma unroll
            for (int j = 0; j < OA_D1; j++)
                if (!(my[(OA_H3 + j) * OA_LD] > 0.0f)) my[(OA_DH3 + j) * OA_LD] = 0.0f;
            oa_layer<OA_D1, OA_D2P, false>(SW + OA_SW3, nullptr, my + OA_DH3 * OA_LD, my + OA_DZ * OA_LD);    // dc
            float cdc = 0.0f;
#pragma unroll
            for (int k = 0; k < OA_D2; k++) cdc = fmaf(my[(OA_C + k) * OA_LD], my[(OA_DZ + k) * OA_LD], cdc);
#pragma unroll
            for (int k = 0; k < OA_D2; k++) my[(OA_DZ + k) * OA_LD] = (my[(OA_DZ + k) * OA_LD] - my[(OA_C + k) * OA_LD] * cdc) * inv_z;
            my[(OA_DZ + OA_D2) * OA_LD] = 0.0f;
            oa_layer<OA_D2P, OA_D1, false>(SW + OA_SW2, nullptr, my + OA_DZ * OA_LD, my + OA_DH1 * OA_LD);
#pragma unroll
            for (int j = 0; j < OA_D1; j++)
                if (!(my[(OA_H1 + j) * OA_LD] > 0.0f)) my[(OA_DH1 + j) * OA_LD] = 0.0f;
        }
        __syncthreads();
        // ---- phase B: partial weight gradients of this tile; entry e = tid + 128 k of the flat parameter vector ----
#pragma unroll 1
        for (int e = tid; e < OA_NPARAM; e += OA_THREADS) {
            int dbase, abase, i, j;  // gradient entry = sum_rows delta[dbase + i][row] * act[abase + j][row]   (bias: act = 1)
            bool bias = false;
            if (e < OA_B1) { dbase = OA_DH1; abase = OA_X; i = e / OA_D0; j = e % OA_D0; }
            else if (e < OA_W2) { dbase = OA_DH1; abase = 0; i = e - OA_B1; j = 0; bias = true; }
            else if (e < OA_B2) { dbase = OA_DZ; abase = OA_H1; i = (e - OA_W2) / OA_D1; j = (e - OA_W2) % OA_D1; }
            else if (e < OA_W3) { dbase = OA_DZ; abase = 0; i = e - OA_B2; j = 0; bias = true; }
            else if (e < OA_B3) { dbase = OA_DH3; abase = OA_C; i = (e - OA_W3) / OA_D2; j = (e - OA_W3) % OA_D2; }
            else if (e < OA_W4) { dbase = OA_DH3; abase = 0; i = e - OA_B3; j = 0; bias = true; }
            else if (e < OA_B4) { dbase = OA_DR; abase = OA_H3; i = (e - OA_W4) / OA_D1; j = (e - OA_W4) % OA_D1; }
            else { dbase = OA_DR; abase = 0; i = e - OA_B4; j = 0; bias = true; }
            const float* dl = vec + (size_t)(dbase + i) * OA_LD;
            const float* ac = vec + (size_t)(abase + j) * OA_LD;
            float s0 = 0.0f, s1 = 0.0f;
            if (bias) {
#pragma unroll 8
                for (int rr_ = 0; rr_ < OA_ROWS; rr_ += 2) { s0 += dl[rr_]; s1 += dl[rr_ + 1]; }
            } else {
#pragma unroll 8
                for (int rr_ = 0; rr_ < OA_ROWS; rr_ += 2) { s0 = fmaf(dl[rr_], ac[rr_], s0); s1 = fmaf(dl[rr_ + 1], ac[rr_ + 1], s1); }
            }
            gacc[e] += s0 + s1;
        }
        __syncthreads();
    }
    // ---- per-CTA slab: gradients + the two loss sums ----
    float* slab = a.partial + (size_t)blockIdx.x * (OA_NPARAM + 2);
    for (int e = tid; e < OA_NPARAM; e += OA_THREADS) slab[e] = gacc[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l1_sum += __shfl_xor_sync(0xffffffffu, l1_sum, o);
        cos_sum += __shfl_xor_sync(0xffffffffu, cos_sum, o);
    }
    if ((tid & 31) == 0) { red[(tid >> 5) * 2] = l1_sum; red[(tid >> 5) * 2 + 1] = cos_sum; }
    __syncthreads();
    if (tid == 0) {
        float s0 = 0.0f, s1 = 0.0f;
        for (int w = 0; w < OA_THREADS / 32; w++) { s0 += red[2 * w]; s1 += red[2 * w + 1]; }
        slab[OA_NPARAM] = s0;
        slab[OA_NPARAM + 1] = s1;
    }
    // ---- phase C: the last CTA reduces the slabs (in CTA order) and applies Adam ----
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const double t = (double)(*a.step + 1);
    const float inv_bc1 = (float)(1.0 / (1.0 - pow((double)a.beta1, t)));
    const float inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)a.beta2, t)));
    // thread t owns entries e = t + 128 k.  The slabs are summed in CTA order (deterministic); the loop nest keeps
    // 4 x 19 independent loads in flight per thread instead of one dependent chain per entry.
    constexpr int NK = (OA_NPARAM + 2 + OA_THREADS - 1) / OA_THREADS;
    float gsum[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) gsum[k] = 0.0f;
    for (unsigned c0 = 0; c0 < gridDim.x; c0 += 4) {
        float t4[4][NK];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const bool ok = c0 + u < gridDim.x;
            const float* sl = a.partial + (size_t)(c0 + u) * (OA_NPARAM + 2);
#pragma unroll
            for (int k = 0; k < NK; k++) {
                const int e = tid + k * OA_THREADS;
                t4[u][k] = (ok && e < OA_NPARAM + 2) ? __ldcg(sl + e) : 0.0f;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int k = 0; k < NK; k++) gsum[k] += t4[u][k];
    }
#pragma unroll
    for (int k = 0; k < NK; k++) {
        const int e = tid + k * OA_THREADS;
        if (e >= OA_NPARAM + 2) continue;
        const float g = gsum[k];
        if (e < OA_NPARAM) {
            // torch.optim.Adam defaults (amsgrad=False, weight_decay=0): same update form as ols_optim.cu
            float m = a.m[e], v = a.v[e];
            m = m + (g - m) * (1.0f - a.beta1);
            v = v * a.beta2 + (1.0f - a.beta2) * g * g;
            a.m[e] = m; a.v[e] = v;
            a.params[e] = a.params[e] - a.lr * inv_bc1 * (m / (sqrtf(v) * inv_sqrt_bc2 + a.eps));
        } else {
            red[e - OA_NPARAM] = g;
        }
    }
    __syncthreads();
    if (tid == 0) {
        a.loss[0] = red[0] * inv_MD + 0.6f * (1.0f - red[1] * inv_M);
        *a.step += 1;
        *a.ticket = 0u;  // ready for the next launch (graph replays included)
    }
}

}  // namespace ols

using namespace ols;

extern "C" {

size_t ols_online_ae_scratch_bytes(void) { return sizeof(float) * (size_t)(OA_NPARAM + 2) * 296 + 256; }
int32_t ols_online_ae_param_count(void) { return OA_NPARAM; }

int ols_online_ae_train_step(float* d_params, float* d_exp_avg, float* d_exp_avg_sq, int64_t* d_step, const float* d_x, int64_t M,
                             float lr, float beta1, float beta2, float eps, float* d_code, float* d_loss, void* d_scratch,
                             size_t scratch_bytes, void* stream) {
    if (!d_params || !d_exp_avg || !d_exp_avg_sq || !d_step || !d_x || !d_loss || !d_scratch || M <= 0) {
        ols_set_error("bad online-autoencoder arguments");
        return OLS_ERR_INVALID;
    }
    if (scratch_bytes < ols_online_ae_scratch_bytes() || ((uintptr_t)d_scratch & 255) != 0 || ((uintptr_t)d_x & 15) != 0) {
        ols_set_error("online-autoencoder scratch too small / misaligned buffers");
        return OLS_ERR_WORKSPACE;
    }
    static bool attr_set = false;
    if (!attr_set) {
        OLS_CUDA_TRY(cudaFuncSetAttribute(k_online_ae_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OA_SMEM));
        attr_set = true;
    }
    OnlineAeArgs a;
    a.params = d_params; a.m = d_exp_avg; a.v = d_exp_avg_sq; a.step = (long long*)d_step; a.x = d_x; a.M = M;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.code = d_code; a.loss = d_loss;
    a.ticket = (unsigned int*)d_scratch;                 // first 256 bytes: the ticket counter (zero before the first use)
    a.partial = (float*)((char*)d_scratch + 256);
    const long long n_tiles = (M + OA_ROWS - 1) / OA_ROWS;
    const int grid = (int)(n_tiles < 148 ? n_tiles : 148);  // one CTA (124 KB of shared memory) per SM
    k_online_ae_step<<<grid, OA_THREADS, OA_SMEM, (cudaStream_t)stream>>>(a);
    OLS_CUDA_TRY(cudaGetLastError());
    return OLS_OK;
}

}  // extern "C"
