// ols_knn.cu -- distCUDA2: mean squared distance of every point to its 3 nearest neighbours.
//
// Reference: submodules/simple-knn/simple_knn.cu:120-220 (SimpleKNN::knn) behind spatial.cu's distCUDA2, called by
// GaussianModel.create_pcd_from_image_and_depth to size new Gaussians (gaussian_splatting/scene/gaussian_model.py:256-262).
// The reference sorts the points along a Morton curve (thrust + cub, three blocking D2H copies), cuts the order into
// boxes of 1024 and lets every point test ALL boxes -- O(P * P / 1024) box tests, exact because the pruning is
// conservative.  Any exact 3-NN search returns the same three distances, so this file uses a uniform grid instead:
//   k_knn_bounds   bounding box of the cloud (order-preserving integer atomics, no host round trip)
//   k_knn_count    points per cell;  k_knn_scan  exclusive scan of the cell counts (one CTA);
//   k_knn_scatter  points copied into cell order (counting sort);
//   k_knn_query    one thread per point: cells are visited in growing cubic shells around the point's cell until the
//                  third-best distance is provably smaller than anything an unvisited shell can hold.
// Distances and the final mean use the reference's expressions (updateKBest<3>, (b0 + b1 + b2) / 3.0f), so the
// output is the same float for every point; fewer than 3 other points leave FLT_MAX entries exactly like the reference.
#include "ols_common.cuh"

#include <cfloat>

namespace ols {

struct KnnGrid {     // lives in the workspace, filled on the device
    unsigned lo[3], hi[3];  // order-preserving encodings of the bounding box
    float min[3], cell, inv_cell, extent;
};

__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void k_knn_init(KnnGrid* g) {
    if (threadIdx.x < 3) { g->lo[threadIdx.x] = 0xffffffffu; g->hi[threadIdx.x] = 0u; }
}

__global__ void __launch_bounds__(256) k_knn_bounds(int P, const float* __restrict__ pts, KnnGrid* g) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float v = pts[3 * (size_t)i + k];
            lo[k] = fminf(lo[k], v);
            hi[k] = fmaxf(hi[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&g->lo[k], f2ord(lo[k]));
            atomicMax(&g->hi[k], f2ord(hi[k]));
        }
    }
}

__global__ void k_knn_setup(KnnGrid* g, int n) {
    float ext = 0.0f;
    for (int k = 0; k < 3; k++) {
        const float lo = ord2f(g->lo[k]), hi = ord2f(g->hi[k]);
        g->min[k] = lo;
        ext = fmaxf(ext, hi - lo);
    }
    if (!(ext > 0.0f) || !(ext < FLT_MAX)) ext = 1.0f;  // all points coincide (or non-finite input): one cell row
    g->extent = ext;
    g->cell = ext / (float)n * 1.0001f;                  // the largest coordinate still falls into cell n - 1
    g->inv_cell = 1.0f / g->cell;
}

__device__ __forceinline__ int3 cell_of(const KnnGrid* g, float x, float y, float z, int n) {
    int3 c;
    c.x = min(n - 1, max(0, __float2int_rd((x - g->min[0]) * g->inv_cell)));
    c.y = min(n - 1, max(0, __float2int_rd((y - g->min[1]) * g->inv_cell)));
    c.z = min(n - 1, max(0, __float2int_rd((z - g->min[2]) * g->inv_cell)));
    return c;
}

__global__ void __launch_bounds__(256) k_knn_count(int P, const float* __restrict__ pts, const KnnGrid* g, int n,
                                                   uint32_t* __restrict__ cell_id, uint32_t* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int3 c = cell_of(g, pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2], n);
    const uint32_t id = ((uint32_t)c.z * n + c.y) * n + c.x;
    cell_id[i] = id;
    atomicAdd(&count[id], 1u);
}

// exclusive scan of `count` (n_cells entries) by one CTA; count[] becomes the running cursor of the scatter
__global__ void __launch_bounds__(1024) k_knn_scan(int n_cells, uint32_t* __restrict__ count, uint32_t* __restrict__ start) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_cells; base += 4096) {
        const int i0 = base + tid * 4;
        uint32_t c[4], sum = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { c[k] = (i0 + k < n_cells) ? count[i0 + k] : 0u; sum += c[k]; }
        uint32_t v = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) s_warp[wid] = v;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        uint32_t run = v - sum + (wid ? s_warp[wid - 1] : 0u) + s_carry;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (i0 + k < n_cells) { start[i0 + k] = run; count[i0 + k] = run; run += c[k]; }
        __syncthreads();
        if (tid == 1023) s_carry = run;
        __syncthreads();
    }
    if (tid == 0) start[n_cells] = s_carry;
}

__global__ void __launch_bounds__(256) k_knn_scatter(int P, const float* __restrict__ pts, const uint32_t* __restrict__ cell_id,
                                                     uint32_t* __restrict__ cursor, float4* __restrict__ sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint32_t pos = atomicAdd(&cursor[cell_id[i]], 1u);
    sorted[pos] = make_float4(pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2], __int_as_float(i));
}

// simple_knn.cu:132-146
__device__ __forceinline__ void update_k_best(const float4& ref, const float4& point, float* knn) {
    const float dx = point.x - ref.x, dy = point.y - ref.y, dz = point.z - ref.z;
    float dist = dx * dx + dy * dy + dz * dz;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (knn[j] > dist) {
            const float t = knn[j];
            knn[j] = dist;
            dist = t;
        }
    }
}

__global__ void __launch_bounds__(128) k_knn_query(int P, const float4* __restrict__ sorted, const uint32_t* __restrict__ start,
                                                   const KnnGrid* __restrict__ g, int n, float* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P) return;
    const float4 me = sorted[t];
    const int3 c = cell_of(g, me.x, me.y, me.z, n);
    const float cell = g->cell, slack = 1e-5f * g->extent;
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    for (int r = 0; r <= n; r++) {
        if (r > 0) {
            // everything not yet visited is at least (r - 1) * cell away (minus rounding slack of the cell assignment)
            const float lim = (float)(r - 1) * cell * 0.9999f - slack;
            if (lim > 0.0f && best[2] <= lim * lim) break;
        }
        const int z0 = max(0, c.z - r), z1 = min(n - 1, c.z + r), y0 = max(0, c.y - r), y1 = min(n - 1, c.y + r);
        for (int z = z0; z <= z1; z++)
            for (int y = y0; y <= y1; y++) {
                const bool face = (z == c.z - r) || (z == c.z + r) || (y == c.y - r) || (y == c.y + r);
                const int xs = face ? 1 : max(1, 2 * r);  // inside rows only touch the two x faces of the shell
                for (int x = c.x - r; x <= c.x + r; x += xs) {
                    if (x < 0 || x >= n) continue;
                    const uint32_t id = ((uint32_t)z * n + y) * n + x;
                    const uint32_t s0 = start[id], s1 = start[id + 1];
                    for (uint32_t q = s0; q < s1; q++)
                        if ((int)q != t) update_k_best(me, sorted[q], best);
                }
            }
    }
    out[__float_as_int(me.w)] = (best[0] + best[1] + best[2]) / 3.0f;  // simple_knn.cu:178
}

static int knn_cells_per_axis(int P) {
    // depth-frame clouds are surfaces: ~P / n^2 points per occupied cell
    int n = (int)sqrtf((float)P / 4.0f);
    if (n < 1) n = 1;
    if (n > 160) n = 160;
    return n;
}

struct KnnLayout { size_t grid, count, start, cell_id, sorted, total; int n; size_t n_cells; };
static KnnLayout knn_layout(int P) {
    KnnLayout L;
    L.n = knn_cells_per_axis(P);
    L.n_cells = (size_t)L.n * L.n * L.n;
    size_t o = 0;
    auto take = [&](size_t b) { size_t at = o; o = align_up(o + b, 256); return at; };
    L.grid = take(sizeof(KnnGrid));
    L.count = take(4 * (L.n_cells + 1));
    L.start = take(4 * (L.n_cells + 1));
    L.cell_id = take(4 * (size_t)(P > 0 ? P : 1));
    L.sorted = take(16 * (size_t)(P > 0 ? P : 1));
    L.total = o;
    return L;
}

}  // namespace ols

using namespace ols;

extern "C" size_t ols_knn_workspace_size(int32_t P) { return P < 0 ? 0 : knn_layout(P).total; }

extern "C" int ols_knn_mean_dist2(int32_t P, const float* d_points, float* d_mean_dist2, void* d_workspace,
                                  size_t workspace_bytes, void* stream) {
    if (P < 0 || (P > 0 && (!d_points || !d_mean_dist2))) { ols_set_error("bad arguments"); return OLS_ERR_INVALID; }
    if (P == 0) return OLS_OK;
    const KnnLayout L = knn_layout(P);
    if (!d_workspace || workspace_bytes < L.total || ((uintptr_t)d_workspace & 255) != 0) {
        ols_set_error("knn workspace: need %zu bytes, 256-byte aligned", L.total);
        return OLS_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)d_workspace;
    KnnGrid* g = (KnnGrid*)(ws + L.grid);
    uint32_t* count = (uint32_t*)(ws + L.count);
    uint32_t* start = (uint32_t*)(ws + L.start);
    uint32_t* cell_id = (uint32_t*)(ws + L.cell_id);
    float4* sorted = (float4*)(ws + L.sorted);
    OLS_CUDA_TRY(cudaMemsetAsync(count, 0, 4 * (L.n_cells + 1), st));
    k_knn_init<<<1, 32, 0, st>>>(g);
    const int blocks = (P + 255) / 256;
    k_knn_bounds<<<blocks < 592 ? blocks : 592, 256, 0, st>>>(P, d_points, g);
    k_knn_setup<<<1, 1, 0, st>>>(g, L.n);
    k_knn_count<<<blocks, 256, 0, st>>>(P, d_points, g, L.n, cell_id, count);
    k_knn_scan<<<1, 1024, 0, st>>>((int)L.n_cells, count, start);
    k_knn_scatter<<<blocks, 256, 0, st>>>(P, d_points, cell_id, count, sorted);
    k_knn_query<<<(P + 127) / 128, 128, 0, st>>>(P, sorted, start, g, L.n, d_mean_dist2);
    OLS_CUDA_TRY(cudaGetLastError());
    return OLS_OK;
}
