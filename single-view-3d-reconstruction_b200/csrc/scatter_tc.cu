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
#include "tiles.cuh"

#include <type_traits>

namespace svr {
using namespace tc;

constexpr int ST_THREADS = 256, ST_TMEM_COLS = 256;
constexpr int ST_A_BYTES = 128 * 128 * 2;   // S_d^T tile: 128 voxels x 128 rows bf16 (two 64-row K chunks, filled and consumed alternately)
constexpr int ST_B_BYTES = 128 * 128 * 2;   // dF slice: 128 rows x up to 128 channels bf16 (double buffered)
constexpr int ST_MAX_VOX = 1024;            // larger boxes fall back to direct reductions
constexpr int ST_SMEM = 1024 + ST_A_BYTES + 2 * ST_B_BYTES + ST_TILE * 16 + 256;

struct StGrad {
    float *g[SVR_MAX_LEVELS];
};

__device__ __forceinline__ void st_red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Tile table: the rows of one group of ST_SUPER consecutive sort cells (a 2x2x2 block of the 16^3 grid: rows that are
// spatially close) form ceil(n / 128) tiles of equal size.  A tile never straddles two groups, so its voxel box is
// bounded by the group's extent plus the stencil reach: 8^3 voxels on a 32^3 level instead of the ~800-voxel boxes (30 %
// above ST_MAX_VOX) of fixed 128-row tiles that cut across group boundaries.  One block; entries = scenes x groups.
__global__ void __launch_bounds__(1024) st_tiles_kernel(const int *__restrict__ cell_start, int n_groups, StTile *__restrict__ tiles,
                                                        int *__restrict__ n_tiles) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s, chunk_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_groups; base += 1024) {
        const int i = base + threadIdx.x;
        int r0 = 0, n = 0;
        if (i < n_groups) {
            r0 = cell_start[i * ST_SUPER];
            n = cell_start[(i + 1) * ST_SUPER] - r0;
        }
        const int t = (n + ST_TILE - 1) / ST_TILE;
        int incl = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = warp_sums[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += v;
            }
            warp_sums[lane] = wi - w;
            if (lane == 31) chunk_s = wi;
        }
        __syncthreads();
        if (t > 0) {
            const int first = carry_s + warp_sums[warp] + incl - t;
            const int per = (n + t - 1) / t;                 // equal shares (130 rows -> 65 + 65, not 128 + 2)
            for (int j = 0; j < t; ++j) tiles[first + j] = StTile{r0 + j * per, min(per, n - j * per)};
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s += chunk_s;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_tiles = carry_s;
}

// Pipeline per level.  A "unit" is (stencil point d, voxel tile mt, K half h): the 64-row half h of the A tile is
// un-filled / filled with the trilinear weights of rows 64h .. 64h+63 for (d, mt) while the tensor core still works on
// the other half (one mbarrier per half), then 4 x tcgen05.mma accumulate into the mt-th TMEM accumulator.  The d+1
// slice of the tile's d-feature rows streams in by cp.async behind the units of d; one epilogue per group of voxel tiles.
// The A tile is zeroed once per kernel and kept clean by un-filling exactly the entries that were written.
__global__ void __launch_bounds__(ST_THREADS, 2) scatter_tc_kernel(const float *__restrict__ points, const int *__restrict__ perm,
                                                                   int N, int64_t total_rows, Pyr P,
                                                                   const __nv_bfloat16 *__restrict__ dfeat, StGrad gv, int level_mask,
                                                                   const StTile *__restrict__ tiles, const int *__restrict__ n_tiles) {
    extern __shared__ uint8_t st_raw[];
    int64_t row0;
    int rows;
    if (tiles) {
        if ((int)blockIdx.x >= *n_tiles) return;           // the grid is an upper bound (uniform exit, nothing set up yet)
        const StTile t = tiles[blockIdx.x];
        row0 = t.row0;
        rows = t.rows;
    } else {
        row0 = (int64_t)blockIdx.x * ST_TILE;
        rows = (int)min((int64_t)ST_TILE, total_rows - row0);
    }
    uint8_t *base = (uint8_t *)(((uintptr_t)st_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *a_tile = base;
    uint8_t *b_tile = base + ST_A_BYTES;           // two buffers
    float4 *pts = (float4 *)(b_tile + 2 * ST_B_BYTES);
    int *box = (int *)(pts + ST_TILE);             // [6] + flag
    uint64_t *mma_done = (uint64_t *)(box + 8);    // [2]: one per K half of the A tile
    uint32_t *tmem_ptr = (uint32_t *)(mma_done + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(mma_done, 1);
        mbar_init(mma_done + 1, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, ST_TMEM_COLS);
    if (tid < ST_TILE) {
        float4 q = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        if (tid < rows) {
            const int64_t row = row0 + tid;
            const int64_t pt = perm ? (int64_t)perm[row] : row;
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
    // fill role: thread -> (row hr of each K half, corner pair): the two x-neighbours of corner (y + bb, z + e).  The four
    // threads of a row are neighbouring lanes of ONE warp: between two units a voxel of the row can pass from one corner
    // pair to another (the cell base moves with the stencil point), and the old owner's un-fill must not overtake the new
    // owner's fill -- a __syncwarp between the two phases orders them.
    const int hr = tid >> 2, bb = tid & 1, e = (tid >> 1) & 1;
    const float4 qh[2] = {pts[hr], pts[64 + hr]};
    const bool vh[2] = {__float_as_int(qh[0].w) >= 0, __float_as_int(qh[1].w) >= 0};
    const int nh = rows > 64 ? 2 : 1;
    uint32_t nc[2] = {0u, 0u};                     // commits so far on each half's barrier
    uint32_t filled[2][2] = {{0xffffffffu, 0xffffffffu}, {0xffffffffu, 0xffffffffu}};   // byte offsets written into A (to un-fill)
    auto wait_half = [&](int h) {
        if (nc[h] > 0) mbar_wait(mma_done + h, (nc[h] - 1) & 1);
    };

    for (int level = 1; level < P.n_levels; ++level) {
        if (!((level_mask >> level) & 1) || !gv.g[level]) continue;   // uniform
        const int C = P.C[level], W = P.W[level], H = P.H[level], D = P.D[level], ncg = C / 8, ncg_sh = 31 - __clz(ncg);
        // ---- bounding box of the in-bounds corners + single-scene test
        if (tid < 6) box[tid] = (tid < 3) ? 0x7fffffff : -0x7fffffff;
        if (tid == 6) box[6] = 1;
        __syncthreads();
        if (tid < rows) {
            const float4 q = pts[tid];
            if (__float_as_int(q.w) != scene0) box[6] = 0;
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
            // box too large (rows that are not spatially grouped / scene boundary): direct per-contribution reductions
            // for this level and tile, one thread per (row, unit)
            const int units = 7 * ncg;
            for (int task = tid; task < rows * units; task += ST_THREADS) {
                const int r = task / units, uu = task - r * units;
                const int d = uu >> ncg_sh, cg = uu & (ncg - 1);
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
                    const int aa = k & 1, b2 = (k >> 1) & 1, e2 = k >> 2;
                    const int x = c.x0 + aa, y = c.y0 + b2, z = c.z0 + e2;
                    if (!corner_in(P, level, x, y, z)) continue;
                    const float w = c.wx[aa] * c.wy[b2] * c.wz[e2];
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
        const int slice_chunks = (nh * 64) << ncg_sh;          // 16-byte pieces of one slice (rows of unused halves are not staged)

        auto prefetch_slice = [&](int d, int buf) {            // cp.async the d-slice into B[buf] (MN-major, swizzled)
            const __nv_bfloat16 *src0 = tile_rows + (P.ubase[level] + d * P.upd[level]) * 8;
            const uint32_t bt = smem_u32(b_tile + buf * ST_B_BYTES);
            for (int i = tid; i < slice_chunks; i += ST_THREADS) {
                const int r = i >> ncg_sh, ch = i & (ncg - 1);
                const bool ok = r < rows;                     // rows past the tile are zero-filled (their weights are zero too)
                cp_async16(bt + (ch >> 3) * (ST_TILE * 128) + swz128(r, ch & 7), src0 + (ok ? (int64_t)r * P.kp + ch * 8 : 0), ok);
            }
            cp_async_commit();
        };

        for (int mt0 = 0; mt0 < n_mt; mt0 += group) {
            const int mt1 = min(mt0 + group, n_mt);
            // every MMA issued so far must be done before B[0] is refilled (and the accumulators reused)
            wait_half(0);
            wait_half(1);
            tc_fence_after();
            prefetch_slice(0, 0);
#pragma unroll 1
            for (int d = 0; d < 7; ++d) {
                cp_async_wait<0>();                            // this thread's pieces of slice d have landed
                Corners ch2[2];
                stencil_corners(P, level, d, qh[0].x, qh[0].y, qh[0].z, ch2[0]);
                stencil_corners(P, level, d, qh[1].x, qh[1].y, qh[1].z, ch2[1]);
                // one unit; `h` is a literal at both call sites, so the per-half state stays in registers
                auto do_unit = [&](auto hc, const int mt) {
                    constexpr int h = decltype(hc)::value;
                    wait_half(h);                          // the MMAs that read this half are done: it may be rewritten
                    tc_fence_after();
                    // after the first unit on every half of this d, every MMA of d-1 is known to be complete:
                    // B[(d+1)&1] may be refilled
                    if (mt == mt0 && h == nh - 1 && d < 6) prefetch_slice(d + 1, (d + 1) & 1);
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        if (filled[h][kk] != 0xffffffffu) *reinterpret_cast<__nv_bfloat16 *>(a_tile + filled[h][kk]) = __float2bfloat16(0.f);
                        filled[h][kk] = 0xffffffffu;
                    }
                    __syncwarp();
                    if (vh[h]) {
                        const Corners &c = ch2[h];
                        const int y = c.y0 + bb, z = c.z0 + e;
                        if (y >= 0 && y < H && z >= 0 && z < D) {
                            const int lrow = ((z - bz0) * ny + (y - by0)) * nx - bx0 - mt * 128;
                            const float wyz = (bb ? c.wy[1] : c.wy[0]) * (e ? c.wz[1] : c.wz[0]);   // selects: no local-memory indexing
#pragma unroll
                            for (int aa = 0; aa < 2; ++aa) {
                                const int x = c.x0 + aa, lid = lrow + x;
                                if (x < 0 || x >= W || lid < 0 || lid >= 128) continue;
                                // element (m = lid, k = row) of a K-major 128B-swizzled tile with two 64-row K chunks
                                const uint32_t off = h * (128 * 128) + swz128(lid, hr >> 3) + (hr & 7) * 2;
                                *reinterpret_cast<__nv_bfloat16 *>(a_tile + off) = __float2bfloat16(c.wx[aa] * wyz);
                                filled[h][aa] = off;
                            }
                        }
                    }
                    fence_proxy_async();     // also covers the cp.async'ed slice (waited above by every thread)
                    __syncthreads();
                    if (tid == 0) {
                        tc_fence_after();
                        const uint32_t a_s = smem_u32(a_tile) + h * (128 * 128), b_s = smem_u32(b_tile + (d & 1) * ST_B_BYTES) + h * (4 * 2048);
                        const uint32_t acc = tmem + (uint32_t)((mt - mt0) * C);
                        const int nk = min(4, (rows - 64 * h + 15) >> 4);      // 16-row K steps that hold rows
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            if (kk < nk) {
                                const uint64_t ad = make_smem_desc(a_s + kk * 32, 16, 1024, kSwizzle128B);
                                const uint64_t bd = make_smem_desc(b_s + kk * 2048, ST_TILE * 128, 1024, kSwizzle128B);
                                umma_bf16(acc, ad, bd, idesc, (d | h | kk) != 0);
                            }
                        }
                        umma_commit(mma_done + h);
                    }
                    ++nc[h];
                };
#pragma unroll 1
                for (int mt = mt0; mt < mt1; ++mt) {
                    do_unit(std::integral_constant<int, 0>{}, mt);
                    if (nh == 2) do_unit(std::integral_constant<int, 1>{}, mt);
                }
            }
            // ---- epilogue: TMEM (lane = voxel, column = channel) -> vector reductions into the volume; warp w reads
            // the 32 lanes of its TMEM quarter (w & 3) and one half of the channel blocks (w >> 2)
            wait_half(0);
            wait_half(1);
            tc_fence_after();
            for (int mt = mt0; mt < mt1; ++mt) {
                const int vox = mt * 128 + (warp & 3) * 32 + lane;
                const bool ok = vox < nvox;
                int lx = 0, ly = 0, lz = 0;
                if (ok) {
                    lx = vox % nx;
                    ly = (vox / nx) % ny;
                    lz = vox / (nx * ny);
                }
                float *dst = gv.g[level] + vol_base + (((int64_t)(bz0 + lz) * H + (by0 + ly)) * W + (bx0 + lx)) * C;
                for (int c0 = (warp >> 2) * 32; c0 < C; c0 += 64) {
                    uint32_t v[32];
                    tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((mt - mt0) * C + c0), v);
                    tmem_ld_wait();
                    if (ok) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float a0 = __uint_as_float(v[4 * j]), a1 = __uint_as_float(v[4 * j + 1]);
                            const float a2 = __uint_as_float(v[4 * j + 2]), a3 = __uint_as_float(v[4 * j + 3]);
                            if (c0 + 4 * j < C && (a0 != 0.f || a1 != 0.f || a2 != 0.f || a3 != 0.f)) st_red_add_v4(dst + c0 + 4 * j, a0, a1, a2, a3);
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

int launch_st_tiles(const int *cell_start, int n_groups, StTile *tiles, int *n_tiles, cudaStream_t st) {
    st_tiles_kernel<<<1, 1024, 0, st>>>(cell_start, n_groups, tiles, n_tiles);
    SVR_LAUNCH_CHECK();
    return 0;
}

// cell_start: first sorted row of every (scene, sort cell), n_cells + 1 entries (svr_sort_points), or nullptr: fixed
// tiles of 128 consecutive rows
int launch_scatter_tc(const float *points, const int *perm, const int *cell_start, int n_cells, int N, int64_t total_rows, const Pyr &P,
                      const __nv_bfloat16 *dfeat, float *const *gvols, int level_mask, cudaStream_t st) {
    static DeviceOnce once;
    int dev;
    if (once.needed(dev)) {
        SVR_CUDA(cudaFuncSetAttribute(scatter_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
        once.done(dev);
    }
    StGrad g;
    for (int l = 0; l < SVR_MAX_LEVELS; ++l) g.g[l] = gvols[l];
    int64_t n_tiles = ceil_div<int64_t>(total_rows, ST_TILE);
    StTile *tiles = nullptr;
    int *count = nullptr;
    void *scratch = nullptr;
    if (cell_start) {
        SVR_REQUIRE(n_cells % ST_SUPER == 0, "scatter: the number of sort cells must be a multiple of %d", ST_SUPER);
        const int n_groups = n_cells / ST_SUPER;
        n_tiles += n_groups;                                   // sum of ceil(n_k / 128) <= total / 128 + groups
        if (int rc = scratch_alloc(&scratch, (size_t)n_tiles * sizeof(StTile) + 16, st)) return rc;
        count = (int *)scratch;
        tiles = (StTile *)((uint8_t *)scratch + 16);
        if (int rc = launch_st_tiles(cell_start, n_groups, tiles, count, st)) return rc;
    }
    scatter_tc_kernel<<<(unsigned)n_tiles, ST_THREADS, ST_SMEM, st>>>(points, perm, N, total_rows, P, dfeat, g, level_mask, tiles, count);
    SVR_LAUNCH_CHECK();
    if (scratch) SVR_CUDA(cudaFreeAsync(scratch, st));
    return 0;
}

}  // namespace svr
