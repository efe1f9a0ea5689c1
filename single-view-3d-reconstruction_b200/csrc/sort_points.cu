// Spatial ordering of query points: a counting sort on (scene, Morton code of a 16^3 cell).
// Consecutive rows of the fused query kernel then touch neighbouring voxels, so the corner fetches
// of the coarse feature levels (87 % of the gathered bytes) hit in L1 instead of going to L2.
// The order inside a cell is irrelevant (every row is independent); only integer counters are used.
#include "common.cuh"

namespace svr {

constexpr int SORT_CELLS = 16;                                  // per axis
constexpr int SORT_KEYS = SORT_CELLS * SORT_CELLS * SORT_CELLS;   // per scene

__device__ __forceinline__ uint32_t spread3(uint32_t v) {   // up to 5 bits -> every third bit
    v &= 0x1F;
    v = (v | (v << 8)) & 0x100F;
    v = (v | (v << 4)) & 0x10C3;
    v = (v | (v << 2)) & 0x1249;
    return v;
}

__device__ __forceinline__ int point_key(const float *p) {
    uint32_t c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float t = (p[k] + 0.5f) * (float)SORT_CELLS;
        t = fminf(fmaxf(t, 0.f), (float)(SORT_CELLS - 1));   // NaN -> 0
        c[k] = (uint32_t)t;
    }
    return (int)(spread3(c[0]) << 2 | spread3(c[1]) << 1 | spread3(c[2]));
}

__global__ void sort_count_kernel(const float *__restrict__ pts, int N, int64_t total, int *__restrict__ key_of, int *__restrict__ count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int key = (int)(i / N) * SORT_KEYS + point_key(pts + i * 3);
    key_of[i] = key;
    atomicAdd(count + key, 1);
}

// exclusive scan of `n` ints by ONE block of 1024 threads (n = B * 4096, a few thousand entries)
__global__ void __launch_bounds__(1024) sort_scan_kernel(const int *__restrict__ in, int *__restrict__ out, int n) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s, chunk_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < n ? in[i] : 0, incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sums[lane] = wi - w;
            if (lane == 31) chunk_s = wi;
        }
        __syncthreads();
        if (i < n) out[i] = carry_s + warp_sums[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s += chunk_s;
        __syncthreads();
    }
}

__global__ void sort_fill_kernel(const int *__restrict__ key_of, int64_t total, const int *__restrict__ start, int *__restrict__ cursor,
                                 int *__restrict__ perm) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int key = key_of[i];
    perm[start[key] + atomicAdd(cursor + key, 1)] = (int)i;
}

}  // namespace svr

using namespace svr;

extern "C" {

size_t svr_sort_points_workspace_bytes(int B, int N) {
    return ((size_t)B * N + 3 * (size_t)B * SORT_KEYS) * sizeof(int) + 1024;
}

int svr_sort_points(const float *points, int B, int N, int *perm, void *workspace, size_t workspace_bytes, void *stream) {
    SVR_REQUIRE(points && perm && workspace, "sort_points: null pointer");
    SVR_REQUIRE((int64_t)B * N < ((int64_t)1 << 31), "sort_points: too many points");
    SVR_REQUIRE(workspace_bytes >= svr_sort_points_workspace_bytes(B, N), "sort_points: workspace too small");
    const int64_t total = (int64_t)B * N;
    if (total == 0) return 0;
    cudaStream_t st = as_stream(stream);
    const int nkeys = B * SORT_KEYS;
    int *key_of = (int *)workspace;
    int *count = key_of + (((size_t)total + 63) / 64) * 64;
    int *cursor = count + nkeys;
    int *start = cursor + nkeys;
    SVR_CUDA(cudaMemsetAsync(count, 0, 2 * (size_t)nkeys * sizeof(int), st));
    const unsigned g = (unsigned)ceil_div<int64_t>(total, 256);
    sort_count_kernel<<<g, 256, 0, st>>>(points, N, total, key_of, count);
    sort_scan_kernel<<<1, 1024, 0, st>>>(count, start, nkeys);
    sort_fill_kernel<<<g, 256, 0, st>>>(key_of, total, start, cursor, perm);
    SVR_LAUNCH_CHECK();
    return 0;
}
}
