// Projection path: unproject -> normalise -> bit-exact trilinear voxelisation -> separable blur.
// Replaces the device work of model/projection.py (reference root) -- see include/svr_b200.h.
//
// Voxelisation design (B200-first, not a port of the 8x index_put_):
//   the dense grid is written once (memset) and only touched voxels are revisited.  Points are
//   bucketed by their floor cell through a 1-bit-per-cell bitmap + popcount rank (a compact cell
//   id without sorting 8x the points), counted and placed with INTEGER counters only, ranked by
//   point index inside each cell, and every touched voxel is then summed by exactly one thread in
//   the reference's serial order (pass (k,j,i)-major, then point index) with non-contracted fp32
//   adds.  No floating-point atomics anywhere; results are bit-exact and run-to-run deterministic.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace svr {

// ------------------------------------------------------------------------------------------------
// unproject / normalise
// ------------------------------------------------------------------------------------------------
struct UnprojParams {
    float f, cx, cy;
    float scale[3], offset[3], half[3], size[3];
    int normalise;
};

__global__ void unproject_fwd_kernel(const float *__restrict__ depth, int H, int W, int64_t total, UnprojParams p,
                                     float *__restrict__ pts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int u = (int)(i % W);
    int v = (int)((i / W) % H);
    float d = depth[i];
    float cam[3];
    cam[0] = __fdiv_rn(__fsub_rn(__fmul_rn((float)u, d), __fmul_rn(p.cx, d)), p.f);
    cam[1] = -__fdiv_rn(__fsub_rn(__fmul_rn((float)v, d), __fmul_rn(p.cy, d)), p.f);
    cam[2] = d;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float g = __fadd_rn(__fmul_rn(p.scale[k], cam[k]), p.offset[k]);
        if (p.normalise) g = __fdiv_rn(__fsub_rn(g, p.half[k]), p.size[k]);
        pts[i * 3 + k] = g;
    }
}

// four consecutive pixels of a row per thread (W % 4 == 0): one 16-byte depth load, three 16-byte point stores; the
// arithmetic per pixel is the scalar kernel's, operation for operation
__global__ void unproject_fwd4_kernel(const float4 *__restrict__ depth, int H, int W, int64_t quads, UnprojParams p,
                                      float4 *__restrict__ pts) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= quads) return;
    const int64_t i = q * 4;
    const int u0 = (int)(i % W);
    const int v = (int)((i / W) % H);
    const float4 d4 = __ldg(depth + q);
    const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
    float o[12];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float d = dd[j];
        float cam[3];
        cam[0] = __fdiv_rn(__fsub_rn(__fmul_rn((float)(u0 + j), d), __fmul_rn(p.cx, d)), p.f);
        cam[1] = -__fdiv_rn(__fsub_rn(__fmul_rn((float)v, d), __fmul_rn(p.cy, d)), p.f);
        cam[2] = d;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float g = __fadd_rn(__fmul_rn(p.scale[k], cam[k]), p.offset[k]);
            if (p.normalise) g = __fdiv_rn(__fsub_rn(g, p.half[k]), p.size[k]);
            o[j * 3 + k] = g;
        }
    }
    pts[q * 3 + 0] = make_float4(o[0], o[1], o[2], o[3]);
    pts[q * 3 + 1] = make_float4(o[4], o[5], o[6], o[7]);
    pts[q * 3 + 2] = make_float4(o[8], o[9], o[10], o[11]);
}

__global__ void unproject_bwd_kernel(const float *__restrict__ gpts, int H, int W, int64_t total, UnprojParams p,
                                     float *__restrict__ gdepth) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int u = (int)(i % W);
    int v = (int)((i / W) % H);
    // cam_k = c_k * d  with c = ((u-cx)/f, -(v-cy)/f, 1);  pts_k = scale_k*cam_k (+const) [/size_k]
    float c[3] = {((float)u - p.cx) / p.f, -((float)v - p.cy) / p.f, 1.0f};
    float g = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float s = p.scale[k] * c[k];
        if (p.normalise) s = s / p.size[k];
        g += gpts[i * 3 + k] * s;
    }
    gdepth[i] = g;
}

__global__ void norm_grid_space_kernel(float *__restrict__ pts, int64_t n, UnprojParams p) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 3) return;
    int k = (int)(i % 3);
    pts[i] = __fdiv_rn(__fsub_rn(pts[i], p.half[k]), p.size[k]);
}

// ------------------------------------------------------------------------------------------------
// voxelisation
// ------------------------------------------------------------------------------------------------
struct VoxParams {
    int B, N;
    int S[3];
    int64_t V;       // voxels per map
    int Wd;          // bitmap words per map
    float lo, hi;    // validity bounds, fp32(-0.5+eps), fp32(0.5-eps)
    float Sm1[3];    // fp32(S-1)
};

// bit-exact per-point quantities (projection.py:44-58); returns validity
__device__ __forceinline__ bool point_frac(const float *__restrict__ p, const VoxParams &vp, int f[3], float r[3],
                                           float m[3]) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v = p[k];
        ok = ok && (v < vp.hi) && (v > vp.lo);
        float g = __fmul_rn(__fadd_rn(v, 0.5f), vp.Sm1[k]);
        float f0 = floorf(g);
        f[k] = (int)f0;
        r[k] = __fsub_rn(g, f0);
        m[k] = __fsub_rn(1.0f, r[k]);
    }
    return ok;
}

__device__ __forceinline__ float corner_weight(const float r[3], const float m[3], int pass) {
    float a = (pass & 4) ? r[0] : m[0];
    float b = (pass & 2) ? r[1] : m[1];
    float c = (pass & 1) ? r[2] : m[2];
    return __fmul_rn(__fmul_rn(a, b), c);
}

// K1: mark occupied cells
__global__ void vox_mark_kernel(const float *__restrict__ pts, VoxParams vp, int *__restrict__ cell_of_point,
                                uint32_t *__restrict__ bitmap) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)vp.B * vp.N) return;
    int b = (int)(i / vp.N);
    int f[3];
    float r[3], m[3];
    bool ok = point_frac(pts + i * 3, vp, f, r, m);
    // the reference would raise on an out-of-range corner; such points cannot pass the validity
    // test, but guard anyway so a malformed input can never write out of bounds
    ok = ok && f[0] >= 0 && f[1] >= 0 && f[2] >= 0 && f[0] + 1 < vp.S[0] && f[1] + 1 < vp.S[1] && f[2] + 1 < vp.S[2];
    int cell = -1;
    if (ok) {
        cell = (f[0] * vp.S[1] + f[1]) * vp.S[2] + f[2];
        atomicOr(bitmap + (size_t)b * vp.Wd + (cell >> 5), 1u << (cell & 31));
    }
    cell_of_point[i] = cell;
}

// Exclusive scan of every map's array, two launches, both grid-wide (round 1 used ONE CTA per map: 64 CTAs on 148 SMs
// walking up to 512 chunks each).  MODE 0: popcount of bitmap words, MODE 1: ints.  A block owns SCAN_BLOCK consecutive
// entries of one map; n_dynamic (nullable) holds the per-map number of valid entries.
constexpr int SCAN_BLOCK = 4096, SCAN_THREADS = 256, SCAN_PER_THREAD = SCAN_BLOCK / SCAN_THREADS;

template <int MODE>
__device__ __forceinline__ uint32_t scan_load(const uint32_t *__restrict__ src, int i, int n) {
    if (i >= n) return 0u;
    return MODE == 0 ? (uint32_t)__popc(src[i]) : src[i];
}

// launch 1: per-block totals
template <int MODE>
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(const uint32_t *__restrict__ in, int n_static,
                                                                        const int *__restrict__ n_dynamic, int stride, int nblk,
                                                                        uint32_t *__restrict__ bsum) {
    __shared__ uint32_t red[SCAN_THREADS / 32];
    const int b = blockIdx.x / nblk, j = blockIdx.x - b * nblk;
    const int n = n_dynamic ? n_dynamic[b] : n_static;
    const uint32_t *src = in + (size_t)b * stride;
    uint32_t v = 0;
    const int i0 = j * SCAN_BLOCK + threadIdx.x * SCAN_PER_THREAD;
    if (i0 + SCAN_PER_THREAD <= n && (stride & 3) == 0) {
        // 16-byte loads: a thread's 16 entries are one 64-byte line segment (scalar loads issued 16 requests of 32 sectors
        // each per warp: the two popcount scans of the 256^3 bitmaps took 232 us for 2 x 134 MB)
#pragma unroll
        for (int k = 0; k < SCAN_PER_THREAD; k += 4) {
            const uint4 q = *reinterpret_cast<const uint4 *>(src + i0 + k);
            v += MODE == 0 ? (uint32_t)(__popc(q.x) + __popc(q.y) + __popc(q.z) + __popc(q.w)) : q.x + q.y + q.z + q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_PER_THREAD; ++k) v += scan_load<MODE>(src, i0 + k, n);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += red[w];
        bsum[blockIdx.x] = t;
    }
}

// launch 2: block prefix (sum of the map's earlier block totals) + local exclusive scan; the map total goes to total_out
template <int MODE>
__global__ void __launch_bounds__(SCAN_THREADS) scan_blocks_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int n_static,
                                                                    const int *__restrict__ n_dynamic, int stride, int nblk,
                                                                    const uint32_t *__restrict__ bsum, int *__restrict__ total_out) {
    __shared__ uint32_t red[SCAN_THREADS / 32];
    __shared__ uint32_t base_s;
    const int b = blockIdx.x / nblk, j = blockIdx.x - b * nblk;
    const int n = n_dynamic ? n_dynamic[b] : n_static;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // prefix of the earlier blocks of this map (nblk is small: <= 128 for a 256^3 bitmap)
    uint32_t pre = 0;
    for (int k = threadIdx.x; k < j; k += SCAN_THREADS) pre += bsum[b * nblk + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
    if (lane == 0) red[warp] = pre;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += red[w];
        base_s = t;
    }
    __syncthreads();
    const uint32_t *src = in + (size_t)b * stride;
    uint32_t *dst = out + (size_t)b * stride;
    const int i0 = j * SCAN_BLOCK + threadIdx.x * SCAN_PER_THREAD;
    uint32_t v[SCAN_PER_THREAD], tsum = 0;
    const bool vec = i0 + SCAN_PER_THREAD <= n && (stride & 3) == 0;
    if (vec) {
#pragma unroll
        for (int k = 0; k < SCAN_PER_THREAD; k += 4) {
            const uint4 q = *reinterpret_cast<const uint4 *>(src + i0 + k);
            v[k] = MODE == 0 ? (uint32_t)__popc(q.x) : q.x;
            v[k + 1] = MODE == 0 ? (uint32_t)__popc(q.y) : q.y;
            v[k + 2] = MODE == 0 ? (uint32_t)__popc(q.z) : q.z;
            v[k + 3] = MODE == 0 ? (uint32_t)__popc(q.w) : q.w;
        }
#pragma unroll
        for (int k = 0; k < SCAN_PER_THREAD; ++k) tsum += v[k];
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_PER_THREAD; ++k) {
            v[k] = scan_load<MODE>(src, i0 + k, n);
            tsum += v[k];
        }
    }
    uint32_t incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();          // red is reused
    if (lane == 31) red[warp] = incl;
    __syncthreads();
    uint32_t wpre = 0;
    for (int w = 0; w < warp; ++w) wpre += red[w];
    uint32_t run = base_s + wpre + incl - tsum;
    if (vec) {
#pragma unroll
        for (int k = 0; k < SCAN_PER_THREAD; k += 4) {
            uint4 q;
            q.x = run;
            q.y = q.x + v[k];
            q.z = q.y + v[k + 1];
            q.w = q.z + v[k + 2];
            run = q.w + v[k + 3];
            *reinterpret_cast<uint4 *>(dst + i0 + k) = q;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_PER_THREAD; ++k) {
            if (i0 + k < n) dst[i0 + k] = run;
            run += v[k];
        }
    }
    if (total_out && j == nblk - 1 && threadIdx.x == SCAN_THREADS - 1) total_out[b] = (int)run;
}

__device__ __forceinline__ int cell_rank(const uint32_t *__restrict__ bitmap, const uint32_t *__restrict__ wprefix,
                                         int cell) {
    uint32_t w = bitmap[cell >> 5];
    return (int)(wprefix[cell >> 5] + __popc(w & ((1u << (cell & 31)) - 1u)));
}

// K3: compact id per point (the rank of its cell among the map's occupied cells), per-cell counts, linear index per id
__global__ void vox_count_kernel(VoxParams vp, const int *__restrict__ cell_of_point, const uint32_t *__restrict__ bitmap,
                                 const uint32_t *__restrict__ wprefix, int *__restrict__ cid_of_point,
                                 int *__restrict__ count, int *__restrict__ cell_lin) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)vp.B * vp.N) return;
    int b = (int)(i / vp.N);
    int cell = cell_of_point[i];
    int id = -1;
    if (cell >= 0) {
        id = cell_rank(bitmap + (size_t)b * vp.Wd, wprefix + (size_t)b * vp.Wd, cell);
        atomicAdd(count + (size_t)b * vp.N + id, 1);
        cell_lin[(size_t)b * vp.N + id] = cell;  // every writer stores the same value
    }
    cid_of_point[i] = id;
}

// K5: place points into their cell's run (arbitrary order inside the run, fixed by K6).  (Writing the record of a point that
// is alone in its cell here, with K6 skipping it, was slower: fill 48 -> 75 us, rank 69 -> 65 us at 128^3 -- the scattered
// 16-byte record stores, not the cursor atomics, are what both kernels wait on.)
__global__ void vox_fill_kernel(VoxParams vp, const int *__restrict__ cid_of_point, const int *__restrict__ start,
                                const int *__restrict__ count, int *__restrict__ cursor, int *__restrict__ order) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)vp.B * vp.N) return;
    int b = (int)(i / vp.N), n = (int)(i % vp.N);
    int id = cid_of_point[i];
    if (id < 0) return;
    size_t o = (size_t)b * vp.N;
    if (count[o + id] == 1) return;          // alone in its cell: rank 0, K6 needs no list (one load instead of an atomic + a store)
    int slot = start[o + id] + atomicAdd(cursor + o + id, 1);
    order[o + slot] = n;
}

// K6: rank every point inside its cell by point index (counting smaller indices) -> deterministic; the point's
// fractional coordinates (the only per-point data the accumulation needs) are stored at its final slot, so that a cell's
// contributions are one contiguous run of 16-byte records in reference (point index) order
__global__ void vox_rank_kernel(const float *__restrict__ pts, VoxParams vp, const int *__restrict__ cid_of_point,
                                const int *__restrict__ start, const int *__restrict__ count, const int *__restrict__ order,
                                float4 *__restrict__ rec) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)vp.B * vp.N) return;
    int b = (int)(i / vp.N), n = (int)(i % vp.N);
    int id = cid_of_point[i];
    if (id < 0) return;
    size_t o = (size_t)b * vp.N;
    int s = start[o + id], c = count[o + id];
    int rank = 0;
    if (c > 1)
        for (int t = 0; t < c; ++t) rank += (order[o + s + t] < n);
    int f[3];
    float r[3], m[3];
    point_frac(pts + i * 3, vp, f, r, m);
    rec[o + s + rank] = make_float4(r[0], r[1], r[2], 0.f);
}

// The low NEED bits of a map's bitmap starting at bit index bit0 (may be negative or run past the end: zeros there); the
// bits above them are unspecified.  The second word is only loaded when the NEED bits straddle a word boundary (3 of 32
// offsets for the 3-bit windows of the accumulate pass, which loaded 18 words per cell before).
template <int NEED>
__device__ __forceinline__ uint32_t bitmap_window(const uint32_t *__restrict__ bm, int64_t bit0, int64_t nbits) {
    if (bit0 <= -32 || bit0 >= nbits) return 0u;
    const int64_t w0 = bit0 >> 5;          // floor division (arithmetic shift)
    const int sh = (int)(bit0 & 31);
    const int64_t nw = (nbits + 31) >> 5;
    const uint32_t lo = (w0 >= 0 && w0 < nw) ? bm[w0] : 0u;
    const uint32_t hi = (sh > 32 - NEED && w0 + 1 >= 0 && w0 + 1 < nw) ? bm[w0 + 1] : 0u;
    return __funnelshift_r(lo, hi, sh);
}

__device__ __forceinline__ float vox_finish(float acc, int64_t flat, int64_t tail_start, uint32_t *__restrict__ sat_mask) {
    float s8;
    if (flat < tail_start) {  // torch.stack(8 aliases).sum(0): row-after-row accumulation
        s8 = acc;
#pragma unroll
        for (int q = 0; q < 7; ++q) s8 = __fadd_rn(s8, acc);
    } else {                  // ATen row_sum remainder path: 4 interleaved partial sums
        const float p2 = __fadd_rn(acc, acc);
        s8 = __fadd_rn(__fadd_rn(__fadd_rn(p2, p2), p2), p2);
    }
    if (sat_mask && s8 > 1.0f) atomicOr(sat_mask + (flat >> 5), 1u << (flat & 31));
    return s8 < 0.f ? 0.f : (s8 > 1.f ? 1.f : s8);
}

// Neighbourhood bit of the source cell of pass ps for the voxel at corner shift sh: ((sz-pz+1)*3 + (sy-py+1))*3 + (sx-px+1).
// c_vox_lower[sh]: the bits of the passes ps < sh (an occupied one owns the voxel); c_vox_src[sh]: those of ps >= sh.
__constant__ uint32_t c_vox_lower[8] = {0x0u, 0x4000u, 0x18000u, 0x34000u, 0x6c0000u, 0xd84000u, 0x3618000u, 0x6c34000u};
__constant__ uint32_t c_vox_src[8] = {0x361bu, 0x2c36u, 0x30d8u, 0x21b0u, 0x3600u, 0x2c00u, 0x3000u, 0x2000u};

// K7: one thread per OCCUPIED CELL (round 1: eight threads per cell, each decoding the cell and probing eight source cells:
// 745 warp instructions per warp, issue bound).  The cell's 3x3x3 neighbourhood occupancy is read once (nine 3-bit windows
// of the bitmap).  Voxel cell+s (s = corner shift = pass index, projection.py:75-78) receives pass s' from the cell
// cell + s - s'; the voxel is OWNED by the occupied source cell with the smallest pass index, and its owner sums every
// contribution in the reference's serial order (pass-major, point index inside a cell: the cell's records are one
// contiguous run in that order) with non-contracted fp32 adds.  An isolated cell (nothing else in the neighbourhood: 45 %
// of the cells of an iid-random depth map at 128^3, 90 % at 256^3) owns all 8 of its voxels and needs only its own run.
// Linear cell offsets may wrap at row / slab ends, but only onto cells with a coordinate S-1, which are never occupied
// (a cell needs f+1 < S); negative / out-of-map indices read as empty.
__global__ void __launch_bounds__(256) vox_accumulate_kernel(VoxParams vp, const int *__restrict__ ucount, const int *__restrict__ cell_lin,
                                                             const uint32_t *__restrict__ bitmap, const uint32_t *__restrict__ wprefix,
                                                             const int *__restrict__ start, const int *__restrict__ count,
                                                             const float4 *__restrict__ rec, int64_t tail_start, float *__restrict__ grid,
                                                             uint32_t *__restrict__ sat_mask) {
    // Phase A: every thread classifies ITS cell (occupied-cell index t): neighbourhood occupancy, isolated or not.  The two
    // kinds are then compacted into two block-wide lists, and phases B / C walk them with consecutive threads: a warp runs
    // ONE of the two code paths on 32 cells of its kind.  (One thread per cell in index order made every warp that held a
    // single non-isolated cell pay for the general path -- 10 % of the cells at 256^3, a third of the warps' time.)
    __shared__ int s_cell[256], s_map[256], s_start[256], s_cnt[256];
    __shared__ uint32_t s_occ[256];
    __shared__ float4 s_rec0[256];                 // the cell's first record
    __shared__ uint16_t s_list[256 + 2048];        // isolated cells (thread index), then owned voxels (thread index << 3 | sh)
    __shared__ int s_wcnt[9][8], s_n[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + tid;
    bool valid = t < (int64_t)vp.B * vp.N;
    int b = 0, u = 0;
    if (valid) {
        b = (int)(t / vp.N);
        u = (int)(t - (int64_t)b * vp.N);
        valid = u < ucount[b];
    }
    const int S2 = vp.S[2], s12 = vp.S[1] * vp.S[2];
    uint32_t occ = 0;
    if (valid) {
        const uint32_t *bm = bitmap + (size_t)b * vp.Wd;
        // The cell's run (start, count: coalesced) and its first record are fetched HERE, next to the bitmap rows: phase B
        // used to start with the dependent chain start/count -> record (ncu at 256^3: 44 % issue active, 16 stalled warps
        // per issue on the long scoreboard, a third of them on these two loads).
        const int cell = cell_lin[(size_t)b * vp.N + u];
        const int st0 = start[(size_t)b * vp.N + u], cn0 = count[(size_t)b * vp.N + u];
        const float4 r0 = __ldg(rec + (size_t)b * vp.N + st0);
        // occupancy of the neighbourhood: bit ((dz+1)*3 + (dy+1))*3 + (dx+1)
        const int first = cell - s12 - S2 - 1;
        if (first >= 0 && (int64_t)cell + s12 + S2 + 34 <= (int64_t)vp.Wd * 32) {
            // interior cell (all but the first and last two slabs' worth): no range checks, 32-bit indices, the nine
            // first words loaded back to back
            uint32_t lo[9], hi[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int bit0 = first + (k / 3) * s12 + (k % 3) * S2;
                lo[k] = bm[bit0 >> 5];
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int bit0 = first + (k / 3) * s12 + (k % 3) * S2;
                hi[k] = (bit0 & 31) > 29 ? bm[(bit0 >> 5) + 1] : 0u;
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int bit0 = first + (k / 3) * s12 + (k % 3) * S2;
                occ |= (__funnelshift_r(lo[k], hi[k], bit0 & 31) & 7u) << (k * 3);
            }
        } else {
#pragma unroll
            for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    const uint32_t w3 = bitmap_window<3>(bm, (int64_t)cell + dz * s12 + dy * S2 - 1, vp.V) & 7u;
                    occ |= w3 << (((dz + 1) * 3 + (dy + 1)) * 3);
                }
        }
        s_cell[tid] = cell;
        s_map[tid] = b;
        s_start[tid] = st0;
        s_cnt[tid] = cn0;
        s_rec0[tid] = r0;
        s_occ[tid] = occ;
    }
    // Compaction.  List 0: the isolated cells.  Lists 1..8: the voxels (cell, corner shift sh) that a general cell OWNS, shift
    // by shift -- one entry per voxel that phase C has to sum, so that every lane of a phase-C warp has a voxel (walking
    // the eight shifts of a cell in one thread left the lanes whose cell does not own the current shift idle: about half
    // of them), and lanes of a warp mostly share the shift, which bounds the number of sources (8, 4, 4, 2, 4, 2, 2, 1).
    const bool iso = valid && occ == (1u << 13), gen = valid && !iso;
    uint32_t own = 0;                                  // bit sh: no occupied source with a smaller pass index
    if (gen) {
#pragma unroll
        for (int sh = 0; sh < 8; ++sh) own |= (occ & c_vox_lower[sh]) ? 0u : (1u << sh);
    }
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t rank_lo = 0, rank_hi = 0;                 // 9 ranks inside the warp, 6 bits each (<= 31)
    {
        const uint32_t m = __ballot_sync(0xffffffffu, iso);
        if (lane == 0) s_wcnt[0][warp] = __popc(m);
        rank_lo = __popc(m & lt);
    }
#pragma unroll
    for (int sh = 0; sh < 8; ++sh) {
        const uint32_t m = __ballot_sync(0xffffffffu, (own >> sh) & 1u);
        if (lane == 0) s_wcnt[sh + 1][warp] = __popc(m);
        const uint32_t r = __popc(m & lt);
        if (sh < 4) rank_lo |= r << (6 * (sh + 1));
        else rank_hi |= r << (6 * (sh - 4));
    }
    __syncthreads();
    if (warp == 0) {
        // exclusive scan of the 72 (list, warp) counts in list-major order; the voxel lists share one index space
        int c[3], tot = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int idx = lane * 3 + k;
            c[k] = idx < 72 ? s_wcnt[idx >> 3][idx & 7] : 0;
            tot += c[k];
        }
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t2 = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t2;
        }
        int run = incl - tot;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int idx = lane * 3 + k;
            if (idx < 72) s_wcnt[idx >> 3][idx & 7] = run;
            run += c[k];
        }
        if (lane == 31) s_n[1] = incl;                              // everything: isolated cells + owned voxels
    }
    __syncthreads();
    // list 0 occupies [0, n_iso) of the shared index space, the voxel lists follow: n_iso = offset of (list 1, warp 0)
    if (tid == 0) s_n[0] = s_wcnt[1][0];
    if (iso) s_list[s_wcnt[0][warp] + (rank_lo & 63u)] = (uint16_t)tid;
    if (gen) {
#pragma unroll
        for (int sh = 0; sh < 8; ++sh)
            if ((own >> sh) & 1u) {
                const uint32_t r = sh < 4 ? (rank_lo >> (6 * (sh + 1))) & 63u : (rank_hi >> (6 * (sh - 4))) & 63u;
                s_list[s_wcnt[sh + 1][warp] + r] = (uint16_t)((tid << 3) | sh);
            }
    }
    __syncthreads();
    const int n_iso = s_n[0], n_all = s_n[1];
    // Phase B: isolated cells -- every one of the 8 voxels has the cell's own run as its only contribution
    for (int i = tid; i < n_iso; i += blockDim.x) {
        const int e = s_list[i], cell = s_cell[e], bb = s_map[e];
        const size_t o = (size_t)bb * vp.N;
        float *gmap = grid + (int64_t)bb * vp.V;
        const int s_own = s_start[e], c_own = s_cnt[e];
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int q = 0; q < c_own; ++q) {
            const float4 rr = q == 0 ? s_rec0[e] : __ldg(rec + o + s_own + q);
            const float r[3] = {rr.x, rr.y, rr.z};
            const float m[3] = {__fsub_rn(1.0f, rr.x), __fsub_rn(1.0f, rr.y), __fsub_rn(1.0f, rr.z)};
#pragma unroll
            for (int ps = 0; ps < 8; ++ps) acc[ps] = __fadd_rn(acc[ps], corner_weight(r, m, ps));
        }
#pragma unroll
        for (int ps = 0; ps < 8; ++ps) {
            const int64_t v = (int64_t)cell + ((ps >> 2) & 1) * s12 + ((ps >> 1) & 1) * S2 + (ps & 1);
            gmap[v] = vox_finish(acc[ps], (int64_t)bb * vp.V + v, tail_start, sat_mask);
        }
    }
    // Phase C: the voxels owned by general cells, one per thread.  The owner sums every contribution in the reference's
    // order: the sources are walked as the SET BITS of occ & mask from the highest bit down, which is ascending pass order
    // (bit index and pass index run opposite ways in every coordinate).  (Round-2 history: a rolled (sh, ps) double loop
    // per cell spent 87 % of the kernel's instructions on index arithmetic and bit tests (ncu source view); unrolling it
    // made 36 separate bodies that the lanes of a warp no longer share: 520 -> 870 us.)
    for (int i = n_iso + tid; i < n_all; i += blockDim.x) {
        const int ent = s_list[i], e = ent >> 3, sh = ent & 7;
        const int cell = s_cell[e], bb = s_map[e];
        const uint32_t occ_e = s_occ[e];
        const size_t o = (size_t)bb * vp.N;
        const uint32_t *bm = bitmap + (size_t)bb * vp.Wd;
        const uint32_t *wp = wprefix + (size_t)bb * vp.Wd;
        const int sz = (sh >> 2) & 1, sy = (sh >> 1) & 1, sx = sh & 1;
        uint32_t srcs = occ_e & c_vox_src[sh];               // includes this cell (bit 13, pass sh)
        float acc = 0.f;
        while (srcs) {
            const int bit = 31 - __clz(srcs);
            srcs &= ~(1u << bit);
            const int q9 = (bit * 57) >> 9, r9 = bit - q9 * 9, q3 = (r9 * 11) >> 5;      // bit / 9, bit % 9, (bit % 9) / 3
            const int nz = q9 - 1, ny = q3 - 1, nx = r9 - q3 * 3 - 1;
            const int ps = ((sz - nz) << 2) | ((sy - ny) << 1) | (sx - nx);
            int s0 = s_start[e], cnt = s_cnt[e];
            if (bit != 13) {
                const int id = cell_rank(bm, wp, cell + nz * s12 + ny * S2 + nx);
                s0 = start[o + id];
                cnt = count[o + id];
            }
            for (int q = 0; q < cnt; ++q) {
                const float4 rr = __ldg(rec + o + s0 + q);
                const float r[3] = {rr.x, rr.y, rr.z};
                const float m[3] = {__fsub_rn(1.0f, rr.x), __fsub_rn(1.0f, rr.y), __fsub_rn(1.0f, rr.z)};
                acc = __fadd_rn(acc, corner_weight(r, m, ps));
            }
        }
        const int64_t v = (int64_t)cell + sz * s12 + sy * S2 + sx;
        grid[(int64_t)bb * vp.V + v] = vox_finish(acc, (int64_t)bb * vp.V + v, tail_start, sat_mask);
    }
}

__global__ void vox_bwd_kernel(const float *__restrict__ pts, const float *__restrict__ ggrid,
                               const uint32_t *__restrict__ sat_mask, VoxParams vp, float *__restrict__ gpts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)vp.B * vp.N) return;
    int b = (int)(i / vp.N);
    int f[3];
    float r[3], m[3];
    bool ok = point_frac(pts + i * 3, vp, f, r, m);
    ok = ok && f[0] >= 0 && f[1] >= 0 && f[2] >= 0 && f[0] + 1 < vp.S[0] && f[1] + 1 < vp.S[1] && f[2] + 1 < vp.S[2];
    float g[3] = {0.f, 0.f, 0.f};
    if (ok) {
#pragma unroll
        for (int ps = 0; ps < 8; ++ps) {
            int k = (ps >> 2) & 1, j = (ps >> 1) & 1, ii = ps & 1;
            int64_t flat = (int64_t)b * vp.V + ((int64_t)(f[0] + k) * vp.S[1] + (f[1] + j)) * vp.S[2] + (f[2] + ii);
            bool sat = sat_mask ? ((sat_mask[flat >> 5] >> (flat & 31)) & 1u) : false;
            float G = sat ? 0.f : 8.0f * ggrid[flat];
            float w0 = k ? r[0] : m[0], w1 = j ? r[1] : m[1], w2 = ii ? r[2] : m[2];
            g[0] += G * (k ? 1.f : -1.f) * w1 * w2;
            g[1] += G * (j ? 1.f : -1.f) * w0 * w2;
            g[2] += G * (ii ? 1.f : -1.f) * w0 * w1;
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) gpts[i * 3 + k] = g[k] * vp.Sm1[k];
}

// ------------------------------------------------------------------------------------------------
// separable blur
// ------------------------------------------------------------------------------------------------
// out[pos] = sum_t taps[t] * in[pos + (t - r) along axis] (zero outside), optional clamp(0,1)
__global__ void conv_axis_kernel(const float *__restrict__ in, float *__restrict__ out, int D, int H, int W,
                                 int64_t total, int axis, const float *__restrict__ taps, int k, int flip, int clamp01) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int x = (int)(i % W), y = (int)((i / W) % H), z = (int)((i / ((int64_t)W * H)) % D);
    int pos = axis == 0 ? z : (axis == 1 ? y : x);
    int len = axis == 0 ? D : (axis == 1 ? H : W);
    int64_t stride = axis == 0 ? (int64_t)H * W : (axis == 1 ? W : 1);
    int r = k / 2;
    float acc = 0.f;
    for (int t = 0; t < k; ++t) {
        int q = pos + t - r;
        if (q >= 0 && q < len) acc += __ldg(taps + (flip ? k - 1 - t : t)) * in[i + (int64_t)(t - r) * stride];
    }
    if (clamp01) acc = fminf(fmaxf(acc, 0.f), 1.f);
    out[i] = acc;
}

// Fused separable blur: ONE pass over the grid.  A block owns a (16 x 64) column of (y, x) and marches
// along z: every input plane (plus its y/x halo) is loaded once, the W and H correlations run in shared
// memory, the result enters a ring of kd planes, and the D correlation + clamp is emitted as soon as the
// ring holds the needed planes: 4 B read + 4 B written per voxel instead of three read+write passes.
// Out-of-grid input is zero, which reproduces the zero padding of every stage (the reference's order
// W, then H, then D -- projection.py:110-114).
constexpr int BL_TY = 16, BL_TX = 64, BL_MAXR = 3, BL_MAXK = 2 * BL_MAXR + 1;

__global__ void __launch_bounds__(256) blur_fused_kernel(const float *__restrict__ in, float *__restrict__ out, int D, int H, int W,
                                                         const float *__restrict__ taps_w, int kw, const float *__restrict__ taps_h, int kh,
                                                         const float *__restrict__ taps_d, int kd) {
    constexpr int EY = BL_TY + 2 * BL_MAXR, EX = BL_TX + 2 * BL_MAXR;
    constexpr int NLOAD = (EY * EX + 255) / 256;       // halo-plane elements per thread (upper bound)
    constexpr int NW = (EY * BL_TX + 255) / 256;       // W-pass elements per thread
    constexpr int NO = (BL_TY * BL_TX) / 256;          // output elements per thread
    __shared__ float s_in[EY * EX];
    __shared__ float s_w[EY * BL_TX];
    __shared__ float ring[BL_MAXK][BL_TY * BL_TX];
    __shared__ float tw[BL_MAXK], th[BL_MAXK], td[BL_MAXK];
    const int rw = kw / 2, rh = kh / 2, rd = kd / 2;
    const int ey = BL_TY + 2 * rh, ex = BL_TX + 2 * rw;
    if (threadIdx.x < kw) tw[threadIdx.x] = taps_w[threadIdx.x];
    if (threadIdx.x < kh) th[threadIdx.x] = taps_h[threadIdx.x];
    if (threadIdx.x < kd) td[threadIdx.x] = taps_d[threadIdx.x];
    const int tiles_x = (W + BL_TX - 1) / BL_TX, tiles_y = (H + BL_TY - 1) / BL_TY;
    int t = blockIdx.x;
    const int tx = t % tiles_x;
    t /= tiles_x;
    const int ty = t % tiles_y;
    const int64_t b = t / tiles_y;
    const int x0 = tx * BL_TX, y0 = ty * BL_TY;
    const float *src = in + b * (int64_t)D * H * W;
    float *dst = out + b * (int64_t)D * H * W;
    const int64_t plane_stride = (int64_t)H * W;
    // per-thread constant addressing (computed once, reused for every plane)
    int ld_off[NLOAD];       // offset inside a plane, or -1 (outside the grid / unused slot)
#pragma unroll
    for (int j = 0; j < NLOAD; ++j) {
        const int i = threadIdx.x + j * 256;
        ld_off[j] = -1;
        if (i < ey * ex) {
            const int lx = i % ex, ly = i / ex;
            const int x = x0 + lx - rw, y = y0 + ly - rh;
            if (x >= 0 && x < W && y >= 0 && y < H) ld_off[j] = y * W + x;
        }
    }
    int st_off[NO];
#pragma unroll
    for (int j = 0; j < NO; ++j) {
        const int i = threadIdx.x + j * 256;
        const int lx = i % BL_TX, ly = i / BL_TX;
        st_off[j] = (x0 + lx < W && y0 + ly < H) ? (y0 + ly) * W + x0 + lx : -1;
    }
    float nxt[NLOAD];
#pragma unroll
    for (int j = 0; j < NLOAD; ++j) nxt[j] = ld_off[j] >= 0 ? __ldg(src + ld_off[j]) : 0.f;      // plane 0
    __syncthreads();
    for (int zi = 0; zi < D + rd; ++zi) {
        float *plane = ring[zi % kd];
        if (zi < D) {
#pragma unroll
            for (int j = 0; j < NLOAD; ++j) {
                const int i = threadIdx.x + j * 256;
                if (i < ey * ex) s_in[i] = nxt[j];
            }
            if (zi + 1 < D) {                          // prefetch the next plane while this one is processed
                const float *pn = src + (int64_t)(zi + 1) * plane_stride;
#pragma unroll
                for (int j = 0; j < NLOAD; ++j) nxt[j] = ld_off[j] >= 0 ? __ldg(pn + ld_off[j]) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < NW; ++j) {             // W pass
                const int i = threadIdx.x + j * 256;
                if (i < ey * BL_TX) {
                    const int lx = i % BL_TX, ly = i / BL_TX;
                    float acc = 0.f;
                    for (int k = 0; k < kw; ++k) acc += tw[k] * s_in[ly * ex + lx + k];
                    s_w[i] = acc;
                }
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < NO; ++j) {             // H pass -> ring
                const int i = threadIdx.x + j * 256;
                const int lx = i % BL_TX, ly = i / BL_TX;
                float acc = 0.f;
                for (int k = 0; k < kh; ++k) acc += th[k] * s_w[(ly + k) * BL_TX + lx];
                plane[i] = acc;
            }
        }
        __syncthreads();
        const int zo = zi - rd;                        // D pass + clamp for plane zo (each thread reads its own ring slots)
        if (zo >= 0) {
            float *po = dst + (int64_t)zo * plane_stride;
#pragma unroll
            for (int j = 0; j < NO; ++j) {
                if (st_off[j] < 0) continue;
                const int i = threadIdx.x + j * 256;
                float acc = 0.f;
                for (int k = 0; k < kd; ++k) {
                    const int z = zo + k - rd;
                    if (z >= 0 && z < D) acc += td[k] * ring[z % kd][i];
                }
                po[st_off[j]] = fminf(fmaxf(acc, 0.f), 1.f);
            }
        }
        // no barrier needed here: the next iteration only overwrites ring[(zi+1) % kd] after two more barriers,
        // and every thread reads exactly the ring slots it wrote itself
    }
}

// Specialisation for the reference's default 3x3x3 kernel (util/arguments.py:25): compile-time tap counts,
// the D-pass ring lives in registers (each thread only ever reads the planes it produced itself).
__global__ void __launch_bounds__(256) blur_fused333_kernel(const float *__restrict__ in, float *__restrict__ out, int D, int H, int W,
                                                            const float *__restrict__ taps_w, const float *__restrict__ taps_h,
                                                            const float *__restrict__ taps_d) {
    constexpr int EY = BL_TY + 2, EX = BL_TX + 2;
    constexpr int NLOAD = (EY * EX + 255) / 256, NW = (EY * BL_TX + 255) / 256, NO = (BL_TY * BL_TX) / 256;
    __shared__ float s_in[EY * EX];
    __shared__ float s_w[EY * BL_TX];
    const float w0 = __ldg(taps_w), w1 = __ldg(taps_w + 1), w2 = __ldg(taps_w + 2);
    const float h0 = __ldg(taps_h), h1 = __ldg(taps_h + 1), h2 = __ldg(taps_h + 2);
    const float d0 = __ldg(taps_d), d1 = __ldg(taps_d + 1), d2 = __ldg(taps_d + 2);
    const int tiles_x = (W + BL_TX - 1) / BL_TX, tiles_y = (H + BL_TY - 1) / BL_TY;
    int t = blockIdx.x;
    const int tx = t % tiles_x;
    t /= tiles_x;
    const int ty = t % tiles_y;
    const int64_t b = t / tiles_y;
    const int x0 = tx * BL_TX, y0 = ty * BL_TY;
    const float *src = in + b * (int64_t)D * H * W;
    float *dst = out + b * (int64_t)D * H * W;
    const int64_t plane_stride = (int64_t)H * W;
    int ld_off[NLOAD];
#pragma unroll
    for (int j = 0; j < NLOAD; ++j) {
        const int i = threadIdx.x + j * 256;
        ld_off[j] = -1;
        if (i < EY * EX) {
            const int lx = i % EX, ly = i / EX;
            const int x = x0 + lx - 1, y = y0 + ly - 1;
            if (x >= 0 && x < W && y >= 0 && y < H) ld_off[j] = y * W + x;
        }
    }
    int st_off[NO];
#pragma unroll
    for (int j = 0; j < NO; ++j) {
        const int i = threadIdx.x + j * 256;
        const int lx = i % BL_TX, ly = i / BL_TX;
        st_off[j] = (x0 + lx < W && y0 + ly < H) ? (y0 + ly) * W + x0 + lx : -1;
    }
    float nxt[NLOAD], prev2[NO], prev1[NO];
#pragma unroll
    for (int j = 0; j < NLOAD; ++j) nxt[j] = ld_off[j] >= 0 ? __ldg(src + ld_off[j]) : 0.f;
#pragma unroll
    for (int j = 0; j < NO; ++j) prev2[j] = prev1[j] = 0.f;
    for (int zi = 0; zi <= D; ++zi) {
        float cur[NO];
#pragma unroll
        for (int j = 0; j < NO; ++j) cur[j] = 0.f;
        if (zi < D) {
#pragma unroll
            for (int j = 0; j < NLOAD; ++j) {
                const int i = threadIdx.x + j * 256;
                if (i < EY * EX) s_in[i] = nxt[j];
            }
            if (zi + 1 < D) {
                const float *pn = src + (int64_t)(zi + 1) * plane_stride;
#pragma unroll
                for (int j = 0; j < NLOAD; ++j) nxt[j] = ld_off[j] >= 0 ? __ldg(pn + ld_off[j]) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < NW; ++j) {
                const int i = threadIdx.x + j * 256;
                if (i < EY * BL_TX) {
                    const int lx = i % BL_TX, ly = i / BL_TX;
                    const float *r = s_in + ly * EX + lx;
                    s_w[i] = w0 * r[0] + w1 * r[1] + w2 * r[2];
                }
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < NO; ++j) {
                const int i = threadIdx.x + j * 256;
                const int lx = i % BL_TX, ly = i / BL_TX;
                const float *r = s_w + ly * BL_TX + lx;
                cur[j] = h0 * r[0] + h1 * r[BL_TX] + h2 * r[2 * BL_TX];
            }
        }
        if (zi >= 1) {       // output plane zi-1 = d0 * plane(zi-2) + d1 * plane(zi-1) + d2 * plane(zi)
            float *po = dst + (int64_t)(zi - 1) * plane_stride;
#pragma unroll
            for (int j = 0; j < NO; ++j)
                if (st_off[j] >= 0) po[st_off[j]] = fminf(fmaxf(d0 * prev2[j] + d1 * prev1[j] + d2 * cur[j], 0.f), 1.f);
        }
#pragma unroll
        for (int j = 0; j < NO; ++j) {
            prev2[j] = prev1[j];
            prev1[j] = cur[j];
        }
        __syncthreads();     // s_in / s_w are rewritten by the next iteration
    }
}

// 3x3x3 blur, round 2c: ONE WARP marches a (BR rows x 128 x) column along z with everything in registers -- no shared
// memory, no block barrier.  A lane holds 4 consecutive x of every row as one 16-byte load; the x neighbours of the W
// correlation come from the adjacent lanes by shuffle (the two tile-edge lanes load one halo element per row, none when the
// row is a single 128-wide tile), the H correlation runs over the lane's BR + 2 rows, the D correlation over a two-plane
// register ring.  The block version above stages every plane through shared memory with three barriers per plane and
// scalar loads / stores (ncu: 49 thread instructions per voxel, 61 % issue active, 43-48 % of the warps resident: 3.6-3.8
// TB/s of compulsory traffic); this one issues ~14.  Needs W % 4 == 0 and 16-byte aligned grids.  A column is cut into
// `zsplit` z segments (two extra planes of halo each), see svr_blur_fwd.
constexpr int BR = 4;

__device__ __forceinline__ float4 blur_w3(const float4 v, const float left, const float right, const float w0, const float w1,
                                          const float w2) {
    float4 r;
    r.x = w0 * left + w1 * v.x + w2 * v.y;
    r.y = w0 * v.x + w1 * v.y + w2 * v.z;
    r.z = w0 * v.y + w1 * v.z + w2 * v.w;
    r.w = w0 * v.z + w1 * v.w + w2 * right;
    return r;
}

__global__ void __launch_bounds__(64) blur_rows333_kernel(const float *__restrict__ in, float *__restrict__ out, int D, int H, int W,
                                                          int zsplit, int64_t n_columns, const float *__restrict__ taps_w,
                                                          const float *__restrict__ taps_h, const float *__restrict__ taps_d) {
    const int lane = threadIdx.x & 31;
    int64_t id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (id >= n_columns) return;                                  // warp-uniform
    const float w0 = __ldg(taps_w), w1 = __ldg(taps_w + 1), w2 = __ldg(taps_w + 2);
    const float h0 = __ldg(taps_h), h1 = __ldg(taps_h + 1), h2 = __ldg(taps_h + 2);
    const float d0 = __ldg(taps_d), d1 = __ldg(taps_d + 1), d2 = __ldg(taps_d + 2);
    const int tiles_x = (W + 127) >> 7, tiles_y = (H + BR - 1) / BR;
    const int xt = (int)(id % tiles_x);
    id /= tiles_x;
    const int yt = (int)(id % tiles_y);
    id /= tiles_y;
    const int zs = (int)(id % zsplit);
    const int64_t b = id / zsplit;
    const int zseg = (D + zsplit - 1) / zsplit;
    const int z_lo = zs * zseg, z_hi = min(D, z_lo + zseg);
    if (z_lo >= z_hi) return;
    const int x = (xt << 7) + lane * 4, y0 = yt * BR;
    const bool x_in = x < W;
    const bool edge_l = lane == 0 && x > 0, edge_r = lane == 31 && x + 4 < W;
    const float *src = in + b * (int64_t)D * H * W;
    float *dst = out + b * (int64_t)D * H * W;
    const int64_t plane_stride = (int64_t)H * W;
    // row offsets inside a plane (-1: the row is outside the grid -> zeros)
    int roff[BR + 2];
#pragma unroll
    for (int r = 0; r < BR + 2; ++r) {
        const int y = y0 - 1 + r;
        roff[r] = (y >= 0 && y < H && x_in) ? y * W + x : -1;
    }
    float4 raw[BR + 2];
    float edge[BR + 2];
    auto load_plane = [&](int z) {
        const float *pl = src + (int64_t)z * plane_stride;
#pragma unroll
        for (int r = 0; r < BR + 2; ++r) {
            raw[r] = roff[r] >= 0 ? __ldg(reinterpret_cast<const float4 *>(pl + roff[r])) : make_float4(0.f, 0.f, 0.f, 0.f);
            edge[r] = 0.f;
            if (tiles_x > 1 && roff[r] >= 0) {
                if (edge_l) edge[r] = __ldg(pl + roff[r] - 1);
                if (edge_r) edge[r] = __ldg(pl + roff[r] + 4);
            }
        }
    };
    float4 p2[BR], p1[BR];
#pragma unroll
    for (int j = 0; j < BR; ++j) p2[j] = p1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int z_first = z_lo - 1;
    load_plane(z_first >= 0 ? z_first : 0);       // z_lo == 0: there is no plane -1, the loop's first round is the zero plane
    for (int zi = z_first; zi <= z_hi; ++zi) {
        float4 cur[BR];
        if (zi >= 0 && zi < D) {
            // W correlation of the BR + 2 rows (raw holds plane zi)
            float4 wp[BR + 2];
#pragma unroll
            for (int r = 0; r < BR + 2; ++r) {
                float left = __shfl_up_sync(0xffffffffu, raw[r].w, 1), right = __shfl_down_sync(0xffffffffu, raw[r].x, 1);
                if (lane == 0) left = edge[r];
                if (lane == 31) right = edge[r];
                wp[r] = blur_w3(raw[r], left, right, w0, w1, w2);
            }
            if (zi + 1 < D && zi + 1 <= z_hi) load_plane(zi + 1);        // in flight during the H / D correlations and the stores
#pragma unroll
            for (int j = 0; j < BR; ++j) {
                cur[j].x = h0 * wp[j].x + h1 * wp[j + 1].x + h2 * wp[j + 2].x;
                cur[j].y = h0 * wp[j].y + h1 * wp[j + 1].y + h2 * wp[j + 2].y;
                cur[j].z = h0 * wp[j].z + h1 * wp[j + 1].z + h2 * wp[j + 2].z;
                cur[j].w = h0 * wp[j].w + h1 * wp[j + 1].w + h2 * wp[j + 2].w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < BR; ++j) cur[j] = make_float4(0.f, 0.f, 0.f, 0.f);      // plane -1 / plane D: zero padding
        }
        const int zo = zi - 1;                   // output plane zo = d0 * plane(zo-1) + d1 * plane(zo) + d2 * plane(zo+1)
        if (zo >= z_lo) {
            float *po = dst + (int64_t)zo * plane_stride;
#pragma unroll
            for (int j = 0; j < BR; ++j) {
                if (roff[j + 1] < 0) continue;
                float4 o;
                o.x = fminf(fmaxf(d0 * p2[j].x + d1 * p1[j].x + d2 * cur[j].x, 0.f), 1.f);
                o.y = fminf(fmaxf(d0 * p2[j].y + d1 * p1[j].y + d2 * cur[j].y, 0.f), 1.f);
                o.z = fminf(fmaxf(d0 * p2[j].z + d1 * p1[j].z + d2 * cur[j].z, 0.f), 1.f);
                o.w = fminf(fmaxf(d0 * p2[j].w + d1 * p1[j].w + d2 * cur[j].w, 0.f), 1.f);
                *reinterpret_cast<float4 *>(po + roff[j + 1]) = o;
            }
        }
#pragma unroll
        for (int j = 0; j < BR; ++j) {
            p2[j] = p1[j];
            p1[j] = cur[j];
        }
    }
}

// g3 = gout * [conv_d(t1) <= 1]   (clamp backward with inclusive bounds; values are never < 0)
__global__ void clamp_mask_kernel(const float *__restrict__ t1, const float *__restrict__ gout, float *__restrict__ g3,
                                  int D, int H, int W, int64_t total, const float *__restrict__ taps, int k) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int z = (int)((i / ((int64_t)W * H)) % D);
    int64_t stride = (int64_t)H * W;
    int r = k / 2;
    float acc = 0.f;
    for (int t = 0; t < k; ++t) {
        int q = z + t - r;
        if (q >= 0 && q < D) acc += __ldg(taps + t) * t1[i + (int64_t)(t - r) * stride];
    }
    g3[i] = (acc >= 0.f && acc <= 1.f) ? gout[i] : 0.f;
}

// partial[block][t] = sum over the block's positions of g[pos] * src[pos + (t-r) along axis]
__global__ void __launch_bounds__(256) tap_grad_kernel(const float *__restrict__ g, const float *__restrict__ src,
                                                       int D, int H, int W, int64_t total, int axis, int k,
                                                       float *__restrict__ partial) {
    __shared__ float red[8][SVR_MAX_TAPS];
    float loc[SVR_MAX_TAPS];
#pragma unroll
    for (int t = 0; t < SVR_MAX_TAPS; ++t) loc[t] = 0.f;
    int r = k / 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int x = (int)(i % W), y = (int)((i / W) % H), z = (int)((i / ((int64_t)W * H)) % D);
        int pos = axis == 0 ? z : (axis == 1 ? y : x);
        int len = axis == 0 ? D : (axis == 1 ? H : W);
        int64_t stride = axis == 0 ? (int64_t)H * W : (axis == 1 ? W : 1);
        float gv = g[i];
        if (gv != 0.f) {
#pragma unroll
            for (int t = 0; t < SVR_MAX_TAPS; ++t) {
                if (t < k) {
                    int q = pos + t - r;
                    if (q >= 0 && q < len) loc[t] += gv * src[i + (int64_t)(t - r) * stride];
                }
            }
        }
    }
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < SVR_MAX_TAPS; ++t) {
        float v = loc[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < k) {
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        partial[(size_t)blockIdx.x * SVR_MAX_TAPS + threadIdx.x] = v;
    }
}

__global__ void tap_grad_reduce_kernel(const float *__restrict__ partial, int nblocks, int k, float *__restrict__ out) {
    // one warp per tap, fixed summation order -> deterministic
    int t = blockIdx.x;
    if (t >= k) return;
    float v = 0.f;
    for (int b = threadIdx.x; b < nblocks; b += 32) v += partial[(size_t)b * SVR_MAX_TAPS + t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) out[t] = v;
}

static int fill_unproj(UnprojParams &p, float f, float cx, float cy, const float *scale3, const float *offset3,
                       const int64_t *dims3, int normalise) {
    p.f = f;
    p.cx = cx;
    p.cy = cy;
    for (int k = 0; k < 3; ++k) {
        p.scale[k] = scale3 ? scale3[k] : 1.f;
        p.offset[k] = offset3 ? offset3[k] : 0.f;
        p.half[k] = dims3 ? (float)((double)dims3[k] / 2.0) : 0.f;
        p.size[k] = dims3 ? (float)dims3[k] : 1.f;
    }
    p.normalise = normalise;
    return 0;
}

static int fill_vox(VoxParams &vp, int B, int N, const int64_t *dims3, double eps) {
    SVR_REQUIRE(B >= 0 && N >= 0, "voxelize: negative sizes");
    SVR_REQUIRE(dims3 && dims3[0] >= 2 && dims3[1] >= 2 && dims3[2] >= 2, "voxelize: every grid axis must be >= 2");
    int64_t V = dims3[0] * dims3[1] * dims3[2];
    SVR_REQUIRE(V < (int64_t)1 << 31, "voxelize: more than 2^31 voxels per map");
    SVR_REQUIRE((int64_t)B * N * 8 < ((int64_t)1 << 40), "voxelize: too many points");
    vp.B = B;
    vp.N = N;
    vp.V = V;
    vp.Wd = (int)((V + 31) / 32);
    for (int k = 0; k < 3; ++k) {
        vp.S[k] = (int)dims3[k];
        vp.Sm1[k] = (float)(dims3[k] - 1);
    }
    vp.hi = (float)(0.5 - eps);    // python double arithmetic, then cast to the tensor dtype
    vp.lo = (float)(-0.5 + eps);
    return 0;
}

struct VoxWorkspace {
    int *cell_of_point, *cid_of_point, *cell_lin, *count, *start, *cursor, *order, *ucount;
    float4 *rec;
    uint32_t *bitmap, *wprefix, *bsum;
    size_t bytes, zero_bytes;   // [bitmap | count | cursor] are contiguous and zero-filled per call
};

static void layout_ws(VoxWorkspace &w, char *base, int B, int N, int Wd) {
    size_t bn = (size_t)B * N, bw = (size_t)B * Wd;
    auto take = [&](size_t n_elems) {
        char *p = base;
        base += ((n_elems * 4 + 255) / 256) * 256;
        return p;
    };
    char *origin = base;
    w.bitmap = (uint32_t *)take(bw);
    w.count = (int *)take(bn);
    w.cursor = (int *)take(bn);
    w.zero_bytes = (size_t)(base - origin);
    w.wprefix = (uint32_t *)take(bw);
    w.cell_of_point = (int *)take(bn);
    w.cid_of_point = (int *)take(bn);
    w.cell_lin = (int *)take(bn);
    w.start = (int *)take(bn);
    w.order = (int *)take(bn);
    w.rec = (float4 *)take(bn * 4);
    w.ucount = (int *)take((size_t)B);
    const size_t nblk = (size_t)(((Wd > N ? Wd : N) + SCAN_BLOCK - 1) / SCAN_BLOCK);
    w.bsum = (uint32_t *)take((size_t)B * nblk);
    w.bytes = (size_t)(base - origin);
}

}  // namespace svr

using namespace svr;

extern "C" {

int svr_unproject_fwd(const float *depth, int B, int H, int W, float f, float cx, float cy, const float *scale3_host,
                      const float *offset3_host, const int64_t *dims3_host, int normalise, float *pts, void *stream) {
    SVR_REQUIRE(depth && pts && scale3_host && offset3_host, "unproject: null pointer");
    SVR_REQUIRE(!normalise || dims3_host, "unproject: dims required when normalise is set");
    UnprojParams p;
    fill_unproj(p, f, cx, cy, scale3_host, offset3_host, dims3_host, normalise);
    int64_t total = (int64_t)B * H * W;
    if (total == 0) return 0;
    if ((W & 3) == 0 && (((uintptr_t)depth | (uintptr_t)pts) & 15) == 0)
        unproject_fwd4_kernel<<<(unsigned)ceil_div<int64_t>(total / 4, 256), 256, 0, as_stream(stream)>>>(
            (const float4 *)depth, H, W, total / 4, p, (float4 *)pts);
    else
        unproject_fwd_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, as_stream(stream)>>>(depth, H, W, total, p, pts);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_unproject_bwd(const float *grad_pts, int B, int H, int W, float f, float cx, float cy,
                      const float *scale3_host, const int64_t *dims3_host, int normalise, float *grad_depth,
                      void *stream) {
    SVR_REQUIRE(grad_pts && grad_depth && scale3_host, "unproject_bwd: null pointer");
    UnprojParams p;
    fill_unproj(p, f, cx, cy, scale3_host, nullptr, dims3_host, normalise);
    int64_t total = (int64_t)B * H * W;
    if (total == 0) return 0;
    unproject_bwd_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, as_stream(stream)>>>(grad_pts, H, W, total, p,
                                                                                               grad_depth);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_norm_grid_space(float *pts, int64_t n_points, const int64_t *dims3_host, void *stream) {
    SVR_REQUIRE(pts && dims3_host, "norm_grid_space: null pointer");
    UnprojParams p;
    fill_unproj(p, 1.f, 0.f, 0.f, nullptr, nullptr, dims3_host, 1);
    if (n_points == 0) return 0;
    norm_grid_space_kernel<<<(unsigned)ceil_div<int64_t>(n_points * 3, 256), 256, 0, as_stream(stream)>>>(pts, n_points, p);
    SVR_LAUNCH_CHECK();
    return 0;
}

size_t svr_voxelize_workspace_bytes(int B, int N, const int64_t *dims3_host) {
    if (!dims3_host || B <= 0 || N <= 0) return 256;
    int64_t V = dims3_host[0] * dims3_host[1] * dims3_host[2];
    VoxWorkspace w;
    layout_ws(w, nullptr, B, N, (int)((V + 31) / 32));
    return w.bytes + 256;
}

int svr_voxelize_fwd(const float *pts, int B, int N, const int64_t *dims3_host, double eps, int64_t tail_start,
                     float *grid, uint32_t *sat_mask, void *workspace, size_t workspace_bytes, void *stream) {
    VoxParams vp;
    if (int rc = fill_vox(vp, B, N, dims3_host, eps)) return rc;
    SVR_REQUIRE(grid, "voxelize: null grid");
    cudaStream_t st = as_stream(stream);
    int64_t total_vox = (int64_t)B * vp.V;
    if (tail_start < 0 || tail_start > total_vox) tail_start = total_vox;
    if (total_vox == 0) return 0;
    SVR_CUDA(cudaMemsetAsync(grid, 0, (size_t)total_vox * sizeof(float), st));
    if (sat_mask) SVR_CUDA(cudaMemsetAsync(sat_mask, 0, (size_t)((total_vox + 31) / 32) * 4, st));
    if (N == 0) return 0;
    SVR_REQUIRE(pts && workspace, "voxelize: null pointer");
    SVR_REQUIRE(((uintptr_t)workspace & 255) == 0, "voxelize: workspace must be 256-byte aligned");
    SVR_REQUIRE(vp.V < ((int64_t)1 << 31), "voxelize: more than 2^31 voxels per map");
    VoxWorkspace w;
    layout_ws(w, (char *)workspace, B, N, vp.Wd);
    SVR_REQUIRE(workspace_bytes >= w.bytes, "voxelize: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    SVR_CUDA(cudaMemsetAsync(w.bitmap, 0, w.zero_bytes, st));
    const int64_t bn = (int64_t)B * N;
    const unsigned gp = (unsigned)ceil_div<int64_t>(bn, 256);
    vox_mark_kernel<<<gp, 256, 0, st>>>(pts, vp, w.cell_of_point, w.bitmap);
    SVR_LAUNCH_CHECK();
    const int nblk_w = ceil_div(vp.Wd, SCAN_BLOCK), nblk_n = ceil_div(N, SCAN_BLOCK);
    scan_block_sums_kernel<0><<<(unsigned)(B * nblk_w), SCAN_THREADS, 0, st>>>(w.bitmap, vp.Wd, nullptr, vp.Wd, nblk_w, w.bsum);
    scan_blocks_kernel<0><<<(unsigned)(B * nblk_w), SCAN_THREADS, 0, st>>>(w.bitmap, w.wprefix, vp.Wd, nullptr, vp.Wd, nblk_w, w.bsum, w.ucount);
    SVR_LAUNCH_CHECK();
    vox_count_kernel<<<gp, 256, 0, st>>>(vp, w.cell_of_point, w.bitmap, w.wprefix, w.cid_of_point, w.count, w.cell_lin);
    SVR_LAUNCH_CHECK();
    scan_block_sums_kernel<1><<<(unsigned)(B * nblk_n), SCAN_THREADS, 0, st>>>((const uint32_t *)w.count, N, w.ucount, N, nblk_n, w.bsum);
    scan_blocks_kernel<1><<<(unsigned)(B * nblk_n), SCAN_THREADS, 0, st>>>((const uint32_t *)w.count, (uint32_t *)w.start, N, w.ucount, N, nblk_n,
                                                                          w.bsum, nullptr);
    SVR_LAUNCH_CHECK();
    vox_fill_kernel<<<gp, 256, 0, st>>>(vp, w.cid_of_point, w.start, w.count, w.cursor, w.order);
    SVR_LAUNCH_CHECK();
    vox_rank_kernel<<<gp, 256, 0, st>>>(pts, vp, w.cid_of_point, w.start, w.count, w.order, w.rec);
    SVR_LAUNCH_CHECK();
    vox_accumulate_kernel<<<gp, 256, 0, st>>>(vp, w.ucount, w.cell_lin, w.bitmap, w.wprefix, w.start, w.count, w.rec, tail_start, grid, sat_mask);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_voxelize_bwd(const float *pts, const float *grad_grid, const uint32_t *sat_mask, int B, int N,
                     const int64_t *dims3_host, double eps, float *grad_pts, void *stream) {
    VoxParams vp;
    if (int rc = fill_vox(vp, B, N, dims3_host, eps)) return rc;
    if ((int64_t)B * N == 0) return 0;
    SVR_REQUIRE(pts && grad_grid && grad_pts, "voxelize_bwd: null pointer");
    vox_bwd_kernel<<<(unsigned)ceil_div<int64_t>((int64_t)B * N, 256), 256, 0, as_stream(stream)>>>(pts, grad_grid, sat_mask,
                                                                                                 vp, grad_pts);
    SVR_LAUNCH_CHECK();
    return 0;
}

static int check_taps(const float *t, int k) {
    SVR_REQUIRE(t && k >= 1 && k <= SVR_MAX_TAPS && (k & 1), "blur: tap count must be odd and <= %d (got %d)", SVR_MAX_TAPS, k);
    return 0;
}

int svr_blur_fwd(const float *in, int B, int D, int H, int W, const float *taps_w, int kw, const float *taps_h, int kh,
                 const float *taps_d, int kd, float *out, float *tmp0, float *tmp1, void *stream) {
    SVR_REQUIRE(in && out && tmp0 && tmp1, "blur: null pointer");
    SVR_REQUIRE(in != out, "blur: in-place operation is not supported (every output voxel reads its input neighbours)");
    if (int rc = check_taps(taps_w, kw)) return rc;
    if (int rc = check_taps(taps_h, kh)) return rc;
    if (int rc = check_taps(taps_d, kd)) return rc;
    int64_t total = (int64_t)B * D * H * W;
    if (total == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (kw <= 2 * BL_MAXR + 1 && kh <= 2 * BL_MAXR + 1 && kd <= 2 * BL_MAXR + 1) {
        // fused single pass (tmp0/tmp1 unused)
        const int64_t tiles = (int64_t)B * ceil_div(H, BL_TY) * ceil_div(W, BL_TX);
        SVR_REQUIRE(tiles < ((int64_t)1 << 31), "blur: too many tiles");
        if (kw == 3 && kh == 3 && kd == 3 && (W & 3) == 0 && (((uintptr_t)in | (uintptr_t)out) & 15) == 0) {
            const int64_t cols = (int64_t)B * ceil_div(H, BR) * ceil_div(W, 128);
            // z segments of 8..15 planes (two halo planes each, absorbed by L2).  Measured, 64 maps: 256^3 3.38 ms with whole
            // columns, 2.93 / 2.53 / 2.16 / 1.96 / 1.92 ms with 2 / 4 / 8 / 16 / 32 segments (2.11 with 64) -- the resident warps
            // then work inside one or two maps instead of eighteen, and start staggered instead of marching in lockstep;
            // 128^3 0.203 ms +- 2 % for every split from 1 to 16; 64^3 0.072 -> 0.053 ms with 8 (more warps).
            const int zsplit = std::max(1, D / 8);
            const int64_t n_columns = cols * zsplit;
            SVR_REQUIRE(ceil_div<int64_t>(n_columns, 2) < ((int64_t)1 << 31), "blur: too many columns");
            blur_rows333_kernel<<<(unsigned)ceil_div<int64_t>(n_columns, 2), 64, 0, st>>>(in, out, D, H, W, zsplit, n_columns, taps_w,
                                                                                          taps_h, taps_d);
        } else if (kw == 3 && kh == 3 && kd == 3)
            blur_fused333_kernel<<<(unsigned)tiles, 256, 0, st>>>(in, out, D, H, W, taps_w, taps_h, taps_d);
        else
            blur_fused_kernel<<<(unsigned)tiles, 256, 0, st>>>(in, out, D, H, W, taps_w, kw, taps_h, kh, taps_d, kd);
        SVR_LAUNCH_CHECK();
        return 0;
    }
    unsigned g = (unsigned)ceil_div<int64_t>(total, 256);
    conv_axis_kernel<<<g, 256, 0, st>>>(in, tmp0, D, H, W, total, 2, taps_w, kw, 0, 0);
    conv_axis_kernel<<<g, 256, 0, st>>>(tmp0, tmp1, D, H, W, total, 1, taps_h, kh, 0, 0);
    conv_axis_kernel<<<g, 256, 0, st>>>(tmp1, out, D, H, W, total, 0, taps_d, kd, 0, 1);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_blur_bwd(const float *in, const float *grad_out, int B, int D, int H, int W, const float *taps_w, int kw,
                 const float *taps_h, int kh, const float *taps_d, int kd, float *grad_in, float *grad_taps, float *tmp,
                 void *stream) {
    SVR_REQUIRE(in && grad_out && grad_in && grad_taps && tmp, "blur_bwd: null pointer");
    if (int rc = check_taps(taps_w, kw)) return rc;
    if (int rc = check_taps(taps_h, kh)) return rc;
    if (int rc = check_taps(taps_d, kd)) return rc;
    int64_t total = (int64_t)B * D * H * W;
    if (total == 0) return 0;
    cudaStream_t st = as_stream(stream);
    unsigned g = (unsigned)ceil_div<int64_t>(total, 256);
    float *T0 = tmp, *T1 = tmp + total, *T2 = tmp + 2 * total, *T3 = tmp + 3 * total;
    // the tap-gradient partial sums use the head of grad_in (written last) as scratch
    int nblocks = (int)((total / 256 < 1184) ? (total / 256 > 0 ? total / 256 : 1) : 1184);
    SVR_REQUIRE((int64_t)nblocks * SVR_MAX_TAPS <= total, "blur_bwd: grid too small for the tap-gradient scratch");
    float *partial = grad_in;
    conv_axis_kernel<<<g, 256, 0, st>>>(in, T0, D, H, W, total, 2, taps_w, kw, 0, 0);        // t0 = conv_w(x)
    conv_axis_kernel<<<g, 256, 0, st>>>(T0, T1, D, H, W, total, 1, taps_h, kh, 0, 0);        // t1 = conv_h(t0)
    clamp_mask_kernel<<<g, 256, 0, st>>>(T1, grad_out, T2, D, H, W, total, taps_d, kd);      // g3
    tap_grad_kernel<<<nblocks, 256, 0, st>>>(T2, T1, D, H, W, total, 0, kd, partial);
    tap_grad_reduce_kernel<<<kd, 32, 0, st>>>(partial, nblocks, kd, grad_taps + kw + kh);
    conv_axis_kernel<<<g, 256, 0, st>>>(T2, T3, D, H, W, total, 0, taps_d, kd, 1, 0);        // g2
    tap_grad_kernel<<<nblocks, 256, 0, st>>>(T3, T0, D, H, W, total, 1, kh, partial);
    tap_grad_reduce_kernel<<<kh, 32, 0, st>>>(partial, nblocks, kh, grad_taps + kw);
    conv_axis_kernel<<<g, 256, 0, st>>>(T3, T2, D, H, W, total, 1, taps_h, kh, 1, 0);        // g1
    tap_grad_kernel<<<nblocks, 256, 0, st>>>(T2, in, D, H, W, total, 2, kw, partial);
    tap_grad_reduce_kernel<<<kw, 32, 0, st>>>(partial, nblocks, kw, grad_taps);
    conv_axis_kernel<<<g, 256, 0, st>>>(T2, grad_in, D, H, W, total, 2, taps_w, kw, 1, 0);   // grad_in
    SVR_LAUNCH_CHECK();
    return 0;
}
}
