// Tensor-core scatter-add for the coarse feature levels of the IF-Net backward
// (grid_sampler_3d_backward of model/ifnet.py:181-193's F.grid_sample calls, reference root).
//
// Rows arrive spatially sorted (svr_sort_points), so the 128 rows x 7 stencil samples of a tile touch
// a small box of voxels.  For one level and one stencil point d the scatter is a (sparse) matrix
// product          dV_box[voxel, c] += sum_row  S_d[voxel, row] * dF_d[row, c]
// with 8 trilinear weights per row in S_d.  S_d is written DENSE (bf16, zeros elsewhere) into a
// 128 x 128 K-major UMMA tile with plain stores (the 8 corners of a sample are distinct voxels, so
// there are no collisions and no atomics), the d-slice of the tile's d-feature rows is staged as the
// MN-major B operand, and tcgen05.mma accumulates the 7 stencil points in TMEM (fp32).  The box is
// then added to the gradient volume with ONE 16-byte reduction per (voxel, 4 channels) and tile.
// 94 % of S_d is zeros, but the tensor pipe has ~100x the throughput of the list-walking CUDA-core
// formulation it replaces.
#include "common.cuh"
#include "sampling.cuh"
#include "tc05.cuh"

namespace svr {
using namespace tc;

constexpr int ST_TILE = 128, ST_THREADS = 256, ST_TMEM_COLS = 256;
constexpr int ST_A_BYTES = 128 * 128 * 2;   // S_d^T tile: 128 voxels x 128 rows bf16 (two 64-wide K chunks)
constexpr int ST_B_BYTES = 128 * 128 * 2;   // dF slice: 128 rows x up to 128 channels bf16 (double buffered)
constexpr int ST_MAX_VOX = 1024;            // larger boxes fall back to direct reductions
constexpr int ST_SMEM = 1024 + ST_A_BYTES + 2 * ST_B_BYTES + ST_TILE * 16 + 256;

struct StGrad {
    float *g[SVR_MAX_LEVELS];
};

__device__ __forceinline__ void st_red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Pipeline per level:  for d in 0..6 { prefetch slice d+1 (cp.async) | for each voxel tile mt of the
// current group { un-fill the previous weights, fill S_d^T for (d, mt), 8 x tcgen05.mma into the mt-th
// TMEM accumulator } }, then one epilogue per voxel tile.  The A tile is zeroed once per kernel and kept
// clean by un-filling exactly the entries that were written.
__global__ void __launch_bounds__(ST_THREADS, 2) scatter_tc_kernel(const float *__restrict__ points, const int *__restrict__ perm,
                                                                   int N, int64_t total_rows, Pyr P,
                                                                   const __nv_bfloat16 *__restrict__ dfeat, StGrad gv, int level_mask) {
    extern __shared__ uint8_t st_raw[];
    uint8_t *base = (uint8_t *)(((uintptr_t)st_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *a_tile = base;
    uint8_t *b_tile = base + ST_A_BYTES;           // two buffers
    float4 *pts = (float4 *)(b_tile + 2 * ST_B_BYTES);
    int *box = (int *)(pts + ST_TILE);             // [6] + flag
    uint64_t *mma_done = (uint64_t *)(box + 8);
    uint32_t *tmem_ptr = (uint32_t *)(mma_done + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * ST_TILE;
    if (tid == 0) {
        mbar_init(mma_done, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, ST_TMEM_COLS);
    if (tid < ST_TILE) {
        int64_t row = row0 + tid;
        float4 q = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        if (row < total_rows) {
            int64_t pt = perm ? (int64_t)perm[row] : row;
            q = make_float4(points[pt * 3], points[pt * 3 + 1], points[pt * 3 + 2], __int_as_float((int)(pt / N)));
        }
        pts[tid] = q;
    }
    for (int i = tid; i < ST_A_BYTES / 16; i += ST_THREADS) reinterpret_cast<uint4 *>(a_tile)[i] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const int scene0 = __float_as_int(pts[0].w);
    const int pr = tid & (ST_TILE - 1), pk = (tid >> 7) * 4;     // thread -> (row, 4 of the 8 corners)
    const float4 q = pts[pr];
    const int my_scene = __float_as_int(q.w);
    uint32_t n_commits = 0;
    uint32_t filled[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};   // byte offsets written into A (to un-fill)

    for (int level = 1; level < P.n_levels; ++level) {
        if (!((level_mask >> level) & 1) || !gv.g[level]) continue;   // uniform
        const int C = P.C[level], W = P.W[level], H = P.H[level], D = P.D[level], ncg = C / 8;
        // ---- bounding box of the in-bounds corners + single-scene test
        if (tid < 6) box[tid] = (tid < 3) ? 0x7fffffff : -0x7fffffff;
        if (tid == 6) box[6] = 1;
        __syncthreads();
        if (tid < ST_TILE && my_scene >= 0) {
            if (my_scene != scene0) box[6] = 0;
            int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-0x7fffffff, -0x7fffffff, -0x7fffffff};
#pragma unroll
            for (int d = 0; d < 7; ++d) {
                Corners c;
                stencil_corners(P, level, d, q.x, q.y, q.z, c);
                const int xa = max(c.x0, 0), xb = min(c.x0 + 1, W - 1), ya = max(c.y0, 0), yb = min(c.y0 + 1, H - 1);
                const int za = max(c.z0, 0), zb = min(c.z0 + 1, D - 1);
                if (xa <= xb && ya <= yb && za <= zb) {
                    lo[0] = min(lo[0], xa); lo[1] = min(lo[1], ya); lo[2] = min(lo[2], za);
                    hi[0] = max(hi[0], xb); hi[1] = max(hi[1], yb); hi[2] = max(hi[2], zb);
                }
            }
#pragma unroll
            for (int a = 0; a < 3; ++a)
                if (lo[a] <= hi[a]) {
                    atomicMin(&box[a], lo[a]);
                    atomicMax(&box[3 + a], hi[a]);
                }
        }
        __syncthreads();
        const int bx0 = box[0], by0 = box[1], bz0 = box[2];
        const int nx = box[3] - bx0 + 1, ny = box[4] - by0 + 1, nz = box[5] - bz0 + 1;
        const bool one_scene = box[6] != 0 && scene0 >= 0;
        __syncthreads();   // box is re-initialised by the next level
        if (nx <= 0 || ny <= 0 || nz <= 0) continue;   // nothing in bounds (uniform)
        const int64_t nvox64 = (int64_t)nx * ny * nz;
        if (!one_scene || nvox64 > ST_MAX_VOX) {
            // box too large (unsorted rows / scene boundary): direct per-contribution reductions for this
            // level and tile, one thread per (row, unit)
            const int units = 7 * ncg;
            for (int task = tid; task < ST_TILE * units; task += ST_THREADS) {
                const int r = task / units, uu = task - r * units;
                const int d = uu / ncg, cg = uu - d * ncg;
                const float4 qq = pts[r];
                const int scene = __float_as_int(qq.w);
                if (scene < 0) continue;
                float g[8];
                bf16x8_to_float(__ldg(reinterpret_cast<const uint4 *>(dfeat + (row0 + r) * P.kp + (int64_t)(P.ubase[level] + uu) * 8)), g);
                Corners c;
                stencil_corners(P, level, d, qq.x, qq.y, qq.z, c);
                const int64_t vb = (int64_t)scene * D * H * W * C + cg * 8;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int aa = k & 1, bb = (k >> 1) & 1, e = k >> 2;
                    const int x = c.x0 + aa, y = c.y0 + bb, z = c.z0 + e;
                    if (!corner_in(P, level, x, y, z)) continue;
                    const float w = c.wx[aa] * c.wy[bb] * c.wz[e];
                    float *dst = gv.g[level] + vb + (((int64_t)z * H + y) * W + x) * C;
                    st_red_add_v4(dst, g[0] * w, g[1] * w, g[2] * w, g[3] * w);
                    st_red_add_v4(dst + 4, g[4] * w, g[5] * w, g[6] * w, g[7] * w);
                }
            }
            continue;
        }
        const int nvox = (int)nvox64;
        const int n_mt = (nvox + 127) / 128;
        const int group = ST_TMEM_COLS / C;                    // voxel tiles that fit in TMEM at once
        const int64_t vol_base = (int64_t)scene0 * D * H * W * C;
        const uint32_t idesc = make_idesc_bf16(128, C, 0, 1);  // A K-major, B MN-major
        const __nv_bfloat16 *tile_rows = dfeat + row0 * P.kp;
        const int slice_chunks = ST_TILE * ncg;

        auto prefetch_slice = [&](int d, int buf) {            // cp.async the d-slice into B[buf] (MN-major, swizzled)
            const int ubase_d = (P.ubase[level] + d * P.upd[level]) * 8;
            uint8_t *bt = b_tile + buf * ST_B_BYTES;
            for (int i = tid; i < slice_chunks; i += ST_THREADS) {
                const int r = i / ncg, ch = i - r * ncg;
                const bool ok = row0 + r < total_rows;
                cp_async16(smem_u32(bt + (ch >> 3) * (ST_TILE * 128) + swz128(r, ch & 7)),
                           tile_rows + (ok ? (int64_t)r * P.kp + ubase_d + ch * 8 : 0), ok);
            }
            cp_async_commit();
        };

        for (int mt0 = 0; mt0 < n_mt; mt0 += group) {
            const int mt1 = min(mt0 + group, n_mt);
            // all MMAs issued so far must be done before B[0] is refilled
            if (n_commits > 0) mbar_wait(mma_done, (n_commits - 1) & 1);
            tc_fence_after();
            prefetch_slice(0, 0);
#pragma unroll 1
            for (int d = 0; d < 7; ++d) {
                // MMAs of stencil point d-1 read B[(d-1)&1]: they must be done before slice d+1 lands there
                if (d > 0 && n_commits > 0) mbar_wait(mma_done, (n_commits - 1) & 1);
                if (d < 6) {
                    prefetch_slice(d + 1, (d + 1) & 1);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                Corners c;
                if (my_scene >= 0) stencil_corners(P, level, d, q.x, q.y, q.z, c);
                for (int mt = mt0; mt < mt1; ++mt) {
                    if (n_commits > 0) mbar_wait(mma_done, (n_commits - 1) & 1);   // A is free again
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        if (filled[kk] != 0xffffffffu) *reinterpret_cast<__nv_bfloat16 *>(a_tile + filled[kk]) = __float2bfloat16(0.f);
                        filled[kk] = 0xffffffffu;
                    }
                    if (my_scene >= 0) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const int k = pk + kk;
                            const int aa = k & 1, bb = (k >> 1) & 1, e = k >> 2;
                            const int x = c.x0 + aa, y = c.y0 + bb, z = c.z0 + e;
                            if (!corner_in(P, level, x, y, z)) continue;
                            const int lid = ((z - bz0) * ny + (y - by0)) * nx + (x - bx0) - mt * 128;
                            if (lid < 0 || lid >= 128) continue;
                            // element (m = lid, k = pr) of a K-major 128B-swizzled tile with two 64-wide K chunks
                            const uint32_t off = (pr >> 6) * (128 * 128) + swz128(lid, (pr & 63) >> 3) + (pr & 7) * 2;
                            *reinterpret_cast<__nv_bfloat16 *>(a_tile + off) = __float2bfloat16(c.wx[aa] * c.wy[bb] * c.wz[e]);
                            filled[kk] = off;
                        }
                    }
                    fence_proxy_async();     // also covers the cp.async'ed slice (waited above by every thread)
                    __syncthreads();
                    if (tid == 0) {
                        tc_fence_after();
                        const uint32_t a_s = smem_u32(a_tile), b_s = smem_u32(b_tile + (d & 1) * ST_B_BYTES);
                        const uint32_t acc = tmem + (uint32_t)((mt - mt0) * C);
#pragma unroll
                        for (int k16 = 0; k16 < 8; ++k16) {
                            const uint64_t ad = make_smem_desc(a_s + (k16 >> 2) * (128 * 128) + (k16 & 3) * 32, 16, 1024, kSwizzle128B);
                            const uint64_t bd = make_smem_desc(b_s + k16 * 2048, ST_TILE * 128, 1024, kSwizzle128B);
                            umma_bf16(acc, ad, bd, idesc, (d | k16) != 0);
                        }
                        umma_commit(mma_done);
                    }
                    ++n_commits;
                }
            }
            // ---- epilogue: TMEM (lane = voxel, column = channel) -> vector reductions into the volume
            mbar_wait(mma_done, (n_commits - 1) & 1);
            tc_fence_after();
            if (warp < 4) {
                for (int mt = mt0; mt < mt1; ++mt) {
                    const int vox = mt * 128 + warp * 32 + lane;
                    const bool ok = vox < nvox;
                    int lx = 0, ly = 0, lz = 0;
                    if (ok) {
                        lx = vox % nx;
                        ly = (vox / nx) % ny;
                        lz = vox / (nx * ny);
                    }
                    float *dst = gv.g[level] + vol_base + (((int64_t)(bz0 + lz) * H + (by0 + ly)) * W + (bx0 + lx)) * C;
                    for (int c0 = 0; c0 < C; c0 += 32) {
                        uint32_t v[32];
                        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)((mt - mt0) * C + c0), v);
                        tmem_ld_wait();
                        if (ok) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float a0 = __uint_as_float(v[4 * j]), a1 = __uint_as_float(v[4 * j + 1]);
                                const float a2 = __uint_as_float(v[4 * j + 2]), a3 = __uint_as_float(v[4 * j + 3]);
                                if (a0 != 0.f || a1 != 0.f || a2 != 0.f || a3 != 0.f) st_red_add_v4(dst + c0 + 4 * j, a0, a1, a2, a3);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncthreads();   // TMEM is reused by the next group / level
        }
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem, ST_TMEM_COLS);
    }
}

int launch_scatter_tc(const float *points, const int *perm, int N, int64_t total_rows, const Pyr &P, const __nv_bfloat16 *dfeat,
                      float *const *gvols, int level_mask, cudaStream_t st) {
    static DeviceOnce once;
    int dev;
    if (once.needed(dev)) {
        SVR_CUDA(cudaFuncSetAttribute(scatter_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
        once.done(dev);
    }
    StGrad g;
    for (int l = 0; l < SVR_MAX_LEVELS; ++l) g.g[l] = gvols[l];
    const int64_t n_tiles = ceil_div<int64_t>(total_rows, ST_TILE);
    scatter_tc_kernel<<<(unsigned)n_tiles, ST_THREADS, ST_SMEM, st>>>(points, perm, N, total_rows, P, dfeat, g, level_mask);
    SVR_LAUNCH_CHECK();
    return 0;
}

}  // namespace svr
