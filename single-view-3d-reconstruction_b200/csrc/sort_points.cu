// Spatial ordering of query points: a counting sort on (scene, Morton code of a 16^3 cell).
// Consecutive rows of the fused query kernel then touch neighbouring voxels, so the corner fetches
// of the coarse feature levels (87 % of the gathered bytes) hit in L1 instead of going to L2.
// The sort is STABLE (rows of a cell keep their original order) and uses integer counters only, so the processing
// order is reproducible run to run.
#include "common.cuh"

namespace svr {

constexpr int SORT_CELLS = 16;                                  // per axis
constexpr int SORT_KEYS = SORT_CELLS * SORT_CELLS * SORT_CELLS;   // per scene
constexpr float SORT_SHIFT = 0.125f;

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
        // The cell lattice is shifted by SORT_SHIFT cells against the voxel lattices of the coarse feature levels: a group
        // of 2x2x2 cells plus the 7-point stencil then touches 5 (not 6) voxels per axis of a 16^3 level and 3 of an 8^3
        // level (align_corners=False puts voxel centres at half-integer multiples of the cell size), 8 of a 32^3 level --
        // the voxel boxes of the tensor-core interpolation / scatter (fused_query.cu, scatter_tc.cu) shrink 216 -> 125.
        float t = (p[k] + 0.5f) * (float)SORT_CELLS - SORT_SHIFT;
        t = fminf(fmaxf(t, 0.f), (float)(SORT_CELLS - 1));   // NaN -> 0
        c[k] = (uint32_t)t;
    }
    return (int)(spread3(c[0]) << 2 | spread3(c[1]) << 1 | spread3(c[2]));
}

// STABLE counting sort (rows of one cell keep their original order), integer counters only: the processing order
// -- and with it every fp32 reduction over rows downstream (split-K weight gradients, column sums, scatter) -- is the
// same in every run.  A block owns SORT_BLOCK consecutive points of ONE scene:
//   1. sort_count_kernel : key per point; rank of the point among the EARLIER points of its block with the same key
//      (warps take turns in warp order, __match_any_sync inside a warp); per-(key, block) counts -> hist, per-key
//      totals -> count (integer atomics: the sum is order-independent);
//   2. sort_scan_kernel  : exclusive scan of count over (scene, key) -> start;
//   3. sort_block_prefix_kernel : per key, exclusive prefix of hist over the scene's blocks;
//   4. sort_fill_kernel  : perm[start[key] + hist[key][block] + rank] = point.
constexpr int SORT_BLOCK = 1024;

__global__ void __launch_bounds__(SORT_BLOCK) sort_count_kernel(const float *__restrict__ pts, int N, int nblk, int *__restrict__ key_rank,
                                                                int *__restrict__ count, int *__restrict__ hist) {
    __shared__ int cnt[SORT_KEYS];
    const int scene = blockIdx.x / nblk, blk = blockIdx.x - scene * nblk;
    const int local = blk * SORT_BLOCK + threadIdx.x;
    const bool ok = local < N;
    const int64_t i = (int64_t)scene * N + local;
    for (int k = threadIdx.x; k < SORT_KEYS; k += SORT_BLOCK) cnt[k] = 0;
    const int key = ok ? point_key(pts + i * 3) : -1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int before = __popc(peers & ((1u << lane) - 1));      // earlier lanes of this warp with the same key
    int rank = 0;
    __syncthreads();
    for (int w = 0; w < SORT_BLOCK / 32; ++w) {                 // warps in order: deterministic ranks
        if (w == warp && ok) {
            rank = cnt[key] + before;
            __syncwarp(peers);
            if (before == 0) cnt[key] += __popc(peers);
        }
        __syncthreads();
    }
    if (ok) key_rank[i] = key | (rank << 12);                   // SORT_KEYS = 4096 keys, rank < 1024
    for (int k = threadIdx.x; k < SORT_KEYS; k += SORT_BLOCK) {
        const int c = cnt[k];
        hist[((int64_t)scene * SORT_KEYS + k) * nblk + blk] = c;
        if (c) atomicAdd(count + scene * SORT_KEYS + k, c);
    }
}

// exclusive scan of `n` ints by ONE block of 1024 threads (n = B * 4096, a few thousand entries)
__global__ void __launch_bounds__(1024) sort_scan_kernel(const int *__restrict__ in, int *__restrict__ out, int n, int *__restrict__ total_out) {
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
    if (total_out && threadIdx.x == 0) *total_out = carry_s;
}

// one thread per (scene, key): exclusive prefix of the key's per-block counts, in place
__global__ void sort_block_prefix_kernel(int *__restrict__ hist, int nkeys, int nblk) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkeys) return;
    int *h = hist + (int64_t)k * nblk;
    int run = 0;
    for (int b = 0; b < nblk; ++b) {
        const int c = h[b];
        h[b] = run;
        run += c;
    }
}

__global__ void __launch_bounds__(SORT_BLOCK) sort_fill_kernel(const int *__restrict__ key_rank, int N, int nblk, const int *__restrict__ start,
                                                               const int *__restrict__ hist, int *__restrict__ perm) {
    const int scene = blockIdx.x / nblk, blk = blockIdx.x - scene * nblk;
    const int local = blk * SORT_BLOCK + threadIdx.x;
    if (local >= N) return;
    const int64_t i = (int64_t)scene * N + local;
    const int kr = key_rank[i], key = scene * SORT_KEYS + (kr & (SORT_KEYS - 1)), rank = kr >> 12;
    perm[start[key] + hist[(int64_t)key * nblk + blk] + rank] = (int)i;
}

}  // namespace svr

using namespace svr;

extern "C" {

static inline int sort_nblk(int N) { return (N + SORT_BLOCK - 1) / SORT_BLOCK; }

size_t svr_sort_points_workspace_bytes(int B, int N) {
    const size_t nkeys = (size_t)B * SORT_KEYS;
    return ((((size_t)B * N + 63) / 64) * 64 + 2 * nkeys + nkeys * (size_t)sort_nblk(N)) * sizeof(int) + 1024;
}

int svr_sort_cells_per_scene(void) { return SORT_KEYS; }

int svr_sort_points(const float *points, int B, int N, int *perm, int *cell_start, void *workspace, size_t workspace_bytes, void *stream) {
    SVR_REQUIRE(points && perm && workspace, "sort_points: null pointer");
    SVR_REQUIRE((int64_t)B * N < ((int64_t)1 << 31), "sort_points: too many points");
    SVR_REQUIRE(workspace_bytes >= svr_sort_points_workspace_bytes(B, N), "sort_points: workspace too small");
    const int64_t total = (int64_t)B * N;
    cudaStream_t st = as_stream(stream);
    const int nkeys = B * SORT_KEYS, nblk = sort_nblk(N);
    if (total == 0) {
        if (cell_start && B > 0) SVR_CUDA(cudaMemsetAsync(cell_start, 0, ((size_t)nkeys + 1) * sizeof(int), st));
        return 0;
    }
    SVR_REQUIRE((int64_t)B * nblk < ((int64_t)1 << 31), "sort_points: too many blocks");
    int *key_rank = (int *)workspace;
    int *count = key_rank + (((size_t)total + 63) / 64) * 64;
    int *start = count + nkeys;
    int *hist = start + nkeys;
    SVR_CUDA(cudaMemsetAsync(count, 0, (size_t)nkeys * sizeof(int), st));
    sort_count_kernel<<<(unsigned)(B * nblk), SORT_BLOCK, 0, st>>>(points, N, nblk, key_rank, count, hist);
    sort_scan_kernel<<<1, 1024, 0, st>>>(count, start, nkeys, cell_start ? cell_start + nkeys : nullptr);
    sort_block_prefix_kernel<<<(unsigned)ceil_div(nkeys, 256), 256, 0, st>>>(hist, nkeys, nblk);
    sort_fill_kernel<<<(unsigned)(B * nblk), SORT_BLOCK, 0, st>>>(key_rank, N, nblk, start, hist, perm);
    if (cell_start) {      // first sorted row of every (scene, cell) + the total: the row ranges of spatial groups
        SVR_CUDA(cudaMemcpyAsync(cell_start, start, (size_t)nkeys * sizeof(int), cudaMemcpyDeviceToDevice, st));
    }
    SVR_LAUNCH_CHECK();
    return 0;
}
}
