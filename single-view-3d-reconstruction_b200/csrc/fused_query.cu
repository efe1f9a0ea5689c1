// Fused IF-Net query forward: multi-scale trilinear stencil gather -> shared memory (128B-swizzled
// UMMA operand tiles) -> tcgen05 fc_0 -> fc_1 -> fc_2 -> fc_out, ONE persistent kernel.
// Replaces model/ifnet.py:38-61 + :156-197 (6x F.grid_sample, cat, reshape, 4x Conv1d) without ever
// materialising the (B, 2583, N) feature tensor.  Also serves the dense-grid evaluator
// (evaluate_network_on_grid, ifnet.py:215-229) by generating the make_3d_grid lattice on the fly in
// brick order.
//
// One CTA per SM, 128 query points per tile, 448 threads:
//   warps 0-3   epilogue: TMEM -> registers -> (+bias, ReLU, bf16) -> H tile in smem / logits
//   warp  4     weight loader: cp.async.bulk (UBLKCP) of pre-swizzled 32 KB weight chunks
//   warp  5     single-thread tcgen05.mma issue
//   warps 6-13  gather producers: 8 corners x 16 B channel-last loads per (point, unit), trilinear
//               blend in fp32, bf16 pack, st.shared into the swizzled A stage
// TMEM: acc0 = columns [0,256) (fc_0), acc1 = [256,512) (fc_1 and fc_2).  Issue order per tile i:
//   F1(i) | F0(i+1) | F2(i)   so that the long fc_0 of the next tile overlaps this tile's epilogues.
#include "common.cuh"
#include "sampling.cuh"
#include "tc05.cuh"
#include <cstdlib>

namespace svr {
using namespace tc;

constexpr int FQ_TILE = 128;
constexpr int FQ_HID = 256;
constexpr int FQ_NA = 3, FQ_NB = 3;
constexpr int FQ_A_BYTES = FQ_TILE * 128;        // 16 KB: 128 rows x 64 bf16
constexpr int FQ_B_BYTES = FQ_HID * 128;         // 32 KB: 256 rows x 64 bf16
constexpr int FQ_H_BYTES = FQ_TILE * FQ_HID * 2; // 64 KB: 4 K-chunks of 16 KB
constexpr int FQ_EPI_WARPS = 4;
constexpr int fq_threads(int gather_warps) { return (FQ_EPI_WARPS + 2 + gather_warps) * 32; }   // 448 for 8 gather warps
constexpr int FQ_UTAB = 512;                     // unit table entries (KP <= 4096)
constexpr int FQ_SMEM = 1024 + FQ_NA * FQ_A_BYTES + FQ_NB * FQ_B_BYTES + FQ_H_BYTES + 2 * FQ_TILE * 16 + 512 + FQ_UTAB * 4;

struct FqVols {
    const __nv_bfloat16 *v[SVR_MAX_LEVELS];
};

struct FqParams {
    // point source: explicit (points != nullptr) or lattice (dense evaluation)
    const float *points;        // (B*N, 3)
    const int *perm;            // optional row -> point index (sorted processing), may be null
    int N;                      // points per scene (explicit mode)
    int64_t total;              // number of rows to process
    // lattice mode
    int lat_scene, sx, sy, sz, x_begin, bx, by, bz;   // bricks per axis over [x_begin, x_end) x sy x sz
    const float *x0;
    FqVols vols;
    Pyr P;
    const uint8_t *w0_img, *w1_img, *w2_img;   // pre-swizzled chunk images (svr_pack_decoder_images)
    const float *b0, *b1, *b2, *wout, *bout;
    float *out;                 // logits (explicit) or occupancy grid (lattice)
    __nv_bfloat16 *save_h;      // optional (3, total, 256)
    __nv_bfloat16 *save_feat;   // optional (total, KP)
    int apply_sigmoid;
};

constexpr int BRICK_X = 8, BRICK_Y = 4, BRICK_Z = 4;   // 128 lattice points per tile

// torch.linspace(-0.5, 0.5, n)[i] (ifnet.py:204-206): one rounding per element (fmadd kernel)
__device__ __forceinline__ float lin_coord(int i, int n) {
    if (n <= 1) return -0.5f;
    float step = 1.0f / (float)(n - 1);
    return i < n / 2 ? fmaf(step, (float)i, -0.5f) : fmaf(-step, (float)(n - 1 - i), 0.5f);
}

// row of a tile -> point coordinates, scene, and output index (-1 = padding row)
__device__ __forceinline__ void row_point(const FqParams &p, int64_t tile, int r, float &px, float &py, float &pz, int &scene,
                                          int64_t &out_idx) {
    if (p.points) {
        int64_t row = tile * FQ_TILE + r;
        if (row >= p.total) {
            out_idx = -1;
            scene = 0;
            px = py = pz = 0.f;
            return;
        }
        int64_t pt = p.perm ? (int64_t)p.perm[row] : row;
        px = p.points[pt * 3 + 0];
        py = p.points[pt * 3 + 1];
        pz = p.points[pt * 3 + 2];
        scene = (int)(pt / p.N);
        out_idx = pt;
    } else {
        int bz = (int)(tile % p.bz), by = (int)((tile / p.bz) % p.by), bx = (int)(tile / ((int64_t)p.bz * p.by));
        int ix = p.x_begin + bx * BRICK_X + (r >> 4), iy = by * BRICK_Y + ((r >> 2) & 3), iz = bz * BRICK_Z + (r & 3);
        scene = p.lat_scene;
        if (ix >= p.sx || iy >= p.sy || iz >= p.sz || bx >= p.bx) {
            out_idx = -1;
            px = py = pz = 0.f;
            return;
        }
        px = lin_coord(ix, p.sx);
        py = lin_coord(iy, p.sy);
        pz = lin_coord(iz, p.sz);
        out_idx = ((int64_t)ix * p.sy + iy) * p.sz + iz;
    }
}

struct FqSmem {
    uint8_t *a, *b, *h;
    float4 *pts;                 // [2][128] : (px,py,pz, scene as int bits)
    uint64_t *a_full, *a_empty, *b_full, *b_empty, *acc_full, *h_ready;
    uint32_t *tmem_ptr;
    uint32_t *utab;              // [FQ_UTAB] packed decode_unit results
};

__device__ __forceinline__ FqSmem fq_carve(uint8_t *raw) {
    FqSmem s;
    uint8_t *base = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    s.a = base;
    s.b = s.a + FQ_NA * FQ_A_BYTES;
    s.h = s.b + FQ_NB * FQ_B_BYTES;
    s.pts = (float4 *)(s.h + FQ_H_BYTES);
    uint64_t *bars = (uint64_t *)(s.pts + 2 * FQ_TILE);
    s.a_full = bars;
    s.a_empty = s.a_full + FQ_NA;
    s.b_full = s.a_empty + FQ_NA;
    s.b_empty = s.b_full + FQ_NB;
    s.acc_full = s.b_empty + FQ_NB;   // [2]
    s.h_ready = s.acc_full + 2;       // [1]
    s.tmem_ptr = (uint32_t *)(s.h_ready + 1);
    s.utab = (uint32_t *)((uint8_t *)bars + 512);
    return s;
}


template <int FQ_GATHER_WARPS>
__global__ void __launch_bounds__(fq_threads(FQ_GATHER_WARPS), 1) fused_query_kernel(const FqParams p, int64_t n_tiles) {
    constexpr int FQ_GATHER_THREADS = FQ_GATHER_WARPS * 32;
    extern __shared__ uint8_t smem_raw[];
    const FqSmem s = fq_carve(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KC0 = p.P.kp / 64;

    if (threadIdx.x == 0) {
        for (int i = 0; i < FQ_NA; ++i) {
            mbar_init(s.a_full + i, FQ_GATHER_WARPS);
            mbar_init(s.a_empty + i, 1);
        }
        for (int i = 0; i < FQ_NB; ++i) {
            mbar_init(s.b_full + i, 1);
            mbar_init(s.b_empty + i, 1);
        }
        mbar_init(s.acc_full + 0, 1);
        mbar_init(s.acc_full + 1, 1);
        mbar_init(s.h_ready, FQ_EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(s.tmem_ptr, 512);
    for (int u = threadIdx.x; u < KC0 * 8; u += blockDim.x) s.utab[u] = pack_unit(p.P, u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s.tmem_ptr;
    const uint32_t acc0 = tmem, acc1 = tmem + 256;

    // number of tiles of this CTA
    int64_t my_tiles = 0;
    if ((int64_t)blockIdx.x < n_tiles) my_tiles = (n_tiles - 1 - blockIdx.x) / gridDim.x + 1;

    if (warp >= 6) {
        // ======================= gather producers =======================
        const int gt = threadIdx.x - 6 * 32;          // 0..255
        const int unit_in_chunk = gt & 7;
        uint32_t gc = 0;                              // global A-chunk counter
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t tile = blockIdx.x + it * gridDim.x;
            float4 *pts_s = s.pts + (it & 1) * FQ_TILE;
            if (gt < FQ_TILE) {
                float px, py, pz;
                int scene;
                int64_t oi;
                row_point(p, tile, gt, px, py, pz, scene, oi);
                pts_s[gt] = make_float4(px, py, pz, __int_as_float(oi < 0 ? -1 : scene));
            }
            named_bar_sync(1, FQ_GATHER_THREADS);
            for (int kc = 0; kc < KC0; ++kc, ++gc) {
                const int st = gc % FQ_NA;
                const int u = kc * 8 + unit_in_chunk;
                UnitCtx uc;
                make_unit_ctx_packed(p.P, s.utab[u], p.vols.v, uc);
                // compile-time specialised gather for the 128-net's levels (uniform over the chunk's threads per unit)
                int spec = 0;
                if (!p.P.align && uc.real && uc.level > 0 && uc.W == uc.H && uc.H == uc.D) {
                    if (uc.W == 128 && uc.C == 16) spec = 1;
                    else if (uc.W == 64 && uc.C == 32) spec = 2;
                    else if (uc.W == 32 && uc.C == 64) spec = 3;
                    else if (uc.W == 16 && uc.C == 128) spec = 4;
                    else if (uc.W == 8 && uc.C == 128) spec = 5;
                }
                auto gather_one = [&](float qx, float qy, float qz, int scene) -> uint4 {
                    switch (spec) {
                        case 1: return gather_unit_fast_c<128, 16>(uc, qx, qy, qz, scene);
                        case 2: return gather_unit_fast_c<64, 32>(uc, qx, qy, qz, scene);
                        case 3: return gather_unit_fast_c<32, 64>(uc, qx, qy, qz, scene);
                        case 4: return gather_unit_fast_c<16, 128>(uc, qx, qy, qz, scene);
                        case 5: return gather_unit_fast_c<8, 128>(uc, qx, qy, qz, scene);
                        default: return gather_unit_fast(uc, p.P.align, qx, qy, qz, scene);
                    }
                };
                mbar_wait(s.a_empty + st, ((gc / FQ_NA) & 1) ^ 1);
                uint8_t *a_st = s.a + st * FQ_A_BYTES;
#pragma unroll 2
                for (int r = gt >> 3; r < FQ_TILE; r += FQ_GATHER_THREADS / 8) {
                    const float4 q = pts_s[r];
                    const int scene = __float_as_int(q.w);
                    uint4 val = make_uint4(0, 0, 0, 0);
                    if (kc == 0) {
                        // Unit 0 is the level-0 unit: 7 stencil samples of the fp32 input grid (56 scalar loads).  Left to
                        // the one lane that owns unit 0 it made chunk 0 cost about a fifth of a tile's gather time
                        // (1 chunk of 41); instead lane j of the point's 8-lane group takes sample j and lane 0 collects.
                        float smp = 0.f;
                        if (scene >= 0 && unit_in_chunk < 7) {
                            const float *x0b = p.x0 + (int64_t)scene * p.P.D[0] * p.P.H[0] * p.P.W[0];
                            smp = level0_sample(p.P, unit_in_chunk, q.x, q.y, q.z, x0b);
                        }
                        float v8[8];
#pragma unroll
                        for (int dd = 0; dd < 8; ++dd) v8[dd] = __shfl_sync(0xffffffffu, smp, (lane & 24) + dd);
                        if (unit_in_chunk == 0)
                            val = float8_to_bf16(v8);
                        else if (uc.real && uc.level > 0 && scene >= 0)
                            val = gather_one(q.x, q.y, q.z, scene);
                    } else if (uc.real && scene >= 0) {
                        val = gather_one(q.x, q.y, q.z, scene);
                    }
                    *reinterpret_cast<uint4 *>(a_st + swz128(r, unit_in_chunk)) = val;
                    if (p.save_feat) {
                        int64_t row = tile * FQ_TILE + r;
                        if (row < p.total) *reinterpret_cast<uint4 *>(p.save_feat + row * p.P.kp + (int64_t)u * 8) = val;
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(s.a_full + st);
            }
        }
    } else if (warp == 4) {
        // ======================= weight loader =======================
        if (lane == 0) {
            uint32_t wc = 0;
            for (int64_t it = 0; it < my_tiles; ++it) {
                const int n_chunks = KC0 + 4 + 4;
                for (int c = 0; c < n_chunks; ++c, ++wc) {
                    const int st = wc % FQ_NB;
                    mbar_wait(s.b_empty + st, ((wc / FQ_NB) & 1) ^ 1);
                    const uint8_t *src = c < KC0 ? p.w0_img + (size_t)c * FQ_B_BYTES
                                                 : (c < KC0 + 4 ? p.w1_img + (size_t)(c - KC0) * FQ_B_BYTES
                                                                : p.w2_img + (size_t)(c - KC0 - 4) * FQ_B_BYTES);
                    mbar_arrive_expect_tx(s.b_full + st, FQ_B_BYTES);
                    bulk_g2s(smem_u32(s.b + st * FQ_B_BYTES), src, FQ_B_BYTES, s.b_full + st);
                }
            }
        }
    } else if (warp == 5) {
        // ======================= MMA issue =======================
        if (lane == 0 && my_tiles > 0) {
            const uint32_t idesc = make_idesc_bf16(FQ_TILE, FQ_HID, 0, 0);
            uint32_t gc = 0, wc = 0, hr = 0;
            auto wait_b = [&]() {
                const int st = wc % FQ_NB;
                mbar_wait(s.b_full + st, (wc / FQ_NB) & 1);
                return st;
            };
            auto issue_f0 = [&]() {
                for (int kc = 0; kc < KC0; ++kc, ++gc) {
                    const int sa = gc % FQ_NA;
                    mbar_wait(s.a_full + sa, (gc / FQ_NA) & 1);
                    const int sb = wait_b();
                    tc_fence_after();
                    const uint32_t a_s = smem_u32(s.a + sa * FQ_A_BYTES), b_s = smem_u32(s.b + sb * FQ_B_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(acc0, make_smem_desc(a_s + k * 32, 16, 1024, kSwizzle128B),
                                  make_smem_desc(b_s + k * 32, 16, 1024, kSwizzle128B), idesc, (kc | k) != 0);
                    umma_commit(s.a_empty + sa);
                    umma_commit(s.b_empty + sb);
                    ++wc;
                }
                umma_commit(s.acc_full + 0);
            };
            auto issue_hidden = [&]() {   // A = H tile (4 K-chunks), B = next 4 weight chunks, D = acc1
                mbar_wait(s.h_ready, hr & 1);
                ++hr;
                tc_fence_after();
                for (int kc = 0; kc < 4; ++kc) {
                    const int sb = wait_b();
                    tc_fence_after();
                    const uint32_t a_s = smem_u32(s.h + kc * FQ_A_BYTES), b_s = smem_u32(s.b + sb * FQ_B_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(acc1, make_smem_desc(a_s + k * 32, 16, 1024, kSwizzle128B),
                                  make_smem_desc(b_s + k * 32, 16, 1024, kSwizzle128B), idesc, (kc | k) != 0);
                    umma_commit(s.b_empty + sb);
                    ++wc;
                }
                umma_commit(s.acc_full + 1);
            };
            // The weight loader streams W0(i), W1(i), W2(i) per tile in that order, so the issue order
            // must consume them in the same order: F0(i), F1(i), F2(i).  (Overlapping F0(i+1) with the
            // epilogues of tile i needs a second weight ring; kept simple here.)
            for (int64_t it = 0; it < my_tiles; ++it) {
                issue_f0();
                issue_hidden();
                issue_hidden();
            }
        }
        __syncwarp();
    } else {
        // ======================= epilogue =======================
        const int r = warp * 32 + lane;            // row in tile == TMEM lane
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        uint32_t n0 = 0, n1 = 0;                   // completions consumed of acc_full[0], acc_full[1]
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t tile = blockIdx.x + it * gridDim.x;
            float px, py, pz;
            int scene;
            int64_t out_idx;
            row_point(p, tile, r, px, py, pz, scene, out_idx);
            const int64_t row = tile * FQ_TILE + r;
            const bool row_ok = p.points ? row < p.total : out_idx >= 0;
            float dot = 0.f;
#pragma unroll 1
            for (int layer = 0; layer < 3; ++layer) {
                const float *bias = layer == 0 ? p.b0 : (layer == 1 ? p.b1 : p.b2);
                if (layer == 0) {
                    mbar_wait(s.acc_full + 0, n0 & 1);
                    ++n0;
                } else {
                    mbar_wait(s.acc_full + 1, n1 & 1);
                    ++n1;
                }
                tc_fence_after();
                const uint32_t acc = (layer == 0 ? acc0 : acc1) + lane_off;
#pragma unroll 1
                for (int c0 = 0; c0 < FQ_HID; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(acc + c0, v);
                    tmem_ld_wait();
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = fmaxf(__uint_as_float(v[j]) + __ldg(bias + c0 + j), 0.f);
                    if (layer == 2) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) dot = fmaf(f[j], __ldg(p.wout + c0 + j), dot);
                    }
                    uint4 packed[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float g[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) g[j] = f[q * 8 + j];
                        packed[q] = float8_to_bf16(g);
                    }
                    if (layer < 2) {
                        uint8_t *hc = s.h + (c0 >> 6) * FQ_A_BYTES;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<uint4 *>(hc + swz128(r, ((c0 & 63) >> 3) + q)) = packed[q];
                    }
                    if (p.save_h && row_ok && p.points) {
                        __nv_bfloat16 *dst = p.save_h + ((int64_t)layer * p.total + row) * FQ_HID + c0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4 *>(dst + q * 8) = packed[q];
                    }
                }
                if (layer < 2) {
                    fence_proxy_async();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s.h_ready);
                }
            }
            if (row_ok) {
                float logit = dot + __ldg(p.bout);
                if (p.apply_sigmoid) logit = 1.0f / (1.0f + __expf(-logit));
                p.out[out_idx] = logit;
            }
            tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ================================================================================================
// Variant with a TENSOR-CORE gather for the wide (coarse) levels.
//
// Rows arrive spatially sorted, so at the coarse levels the 128 points of a tile touch a small box
// of voxels.  For such a level and stencil point d the gather is a (sparse) matrix product
//        F_d[row, c] = sum_voxel  S_d[row, voxel] * V_box[voxel, c]
// with the 8 trilinear weights of each row in S_d.  The builders write S_d DENSE (bf16) into a
// K-major UMMA tile with plain stores (un-filled again after use), stage the box's channel-last
// voxels as the MN-major B operand, tcgen05.mma produces F_d in TMEM, and the epilogue warps convert
// it into the bf16 A chunk that fc_0 consumes.  Only levels 0-2 (13 % of the features) still go through
// the CUDA-core gather.  Everything else (weight stream, fc_0..fc_out, epilogues) is as above.
// Tiles whose box is too large (unsorted input, scene boundary) use the CUDA-core gather for that level.
// ================================================================================================
constexpr int T_NW = 2, T_NA = 2, T_NSV = 2;
constexpr int T_OFF_W = 0;
constexpr int T_OFF_A = T_OFF_W + T_NW * FQ_B_BYTES;            //  65536
constexpr int T_OFF_H = T_OFF_A + T_NA * FQ_A_BYTES;            //  98304
constexpr int T_OFF_S = T_OFF_H + FQ_H_BYTES;                   // 163840 : S ring, 2 x 16 KB
constexpr int T_OFF_V = T_OFF_S + T_NSV * FQ_A_BYTES;           // 196608 : V area, 2 voxel chunks x 16 KB (one level's box)
constexpr int T_MAX_VC = 2;                                      // voxel chunks (of 64) per level and tile
constexpr int T_OFF_PTS = T_OFF_V + T_MAX_VC * FQ_A_BYTES;      // 229376
constexpr int T_OFF_MISC = T_OFF_PTS + FQ_TILE * 16;            // 231424
constexpr int T_SMEM = T_OFF_MISC + 1024;                       // 232448 == the per-block maximum on sm_100
constexpr int T_BUILD_WARPS = 16, T_BUILD_THREADS = T_BUILD_WARPS * 32;
constexpr int T_THREADS = (FQ_EPI_WARPS + 2 + T_BUILD_WARPS) * 32;   // 704

struct TileInfo {           // written by the builders once per tile, read by the MMA and epilogue warps
    int mode[SVR_MAX_LEVELS];                    // 1: tensor-core gather, 0: CUDA-core gather
    int n_vc[SVR_MAX_LEVELS];
    int box[SVR_MAX_LEVELS][6];                  // x0,y0,z0,nx,ny,nz
    int scene;
};

struct TMisc {
    uint64_t a_full[T_NA], a_empty[T_NA], w_full[T_NW], w_empty[T_NW], sv_full[T_NSV], sv_empty[T_NSV];
    uint64_t f_full[2], f_empty[2], acc_full[2], h_ready, acc1_free, info_ready[2], v_empty;
    uint32_t tmem_ptr;
    int red[SVR_MAX_LEVELS][8];
};
static_assert(sizeof(TMisc) + 2 * sizeof(TileInfo) <= 1024, "misc area");

__device__ unsigned long long g_fq_mode_count[SVR_MAX_LEVELS][2];   // diagnostics: tiles per (level, gather mode)

__device__ __forceinline__ bool tc_level(const Pyr &P, int l) { return l >= 1 && P.C[l] % 64 == 0 && P.C[l] <= 128; }

__global__ void __launch_bounds__(T_THREADS, 1) fused_query_tc_kernel(const FqParams p, int64_t n_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_tc[];
    uint8_t *const sm = smem_tc;
    uint8_t *const w_ring = sm + T_OFF_W, *const a_ring = sm + T_OFF_A, *const h_tile = sm + T_OFF_H;
    uint8_t *const s_ring = sm + T_OFF_S, *const v_area = sm + T_OFF_V;
    float4 *const pts_s = reinterpret_cast<float4 *>(sm + T_OFF_PTS);
    // TileInfo lives in the tail of the points area?  no: keep it in misc after TMisc (two copies)
    TMisc &B = *reinterpret_cast<TMisc *>(sm + T_OFF_MISC);
    TileInfo *const tinfo = reinterpret_cast<TileInfo *>(sm + T_OFF_MISC + ((sizeof(TMisc) + 15) / 16) * 16);   // [2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KC0 = p.P.kp / 64;

    if (threadIdx.x == 0) {
        if ((smem_u32(sm) & 1023u) != 0) {
            printf("svr_b200: dynamic shared memory is not 1024-byte aligned\n");
            __trap();
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&B.a_full[i], 1);
            mbar_init(&B.a_empty[i], 1);
            mbar_init(&B.w_full[i], 1);
            mbar_init(&B.w_empty[i], 1);
            mbar_init(&B.sv_full[i], 1);
            mbar_init(&B.sv_empty[i], 1);
            mbar_init(&B.f_full[i], 1);
            mbar_init(&B.f_empty[i], 1);
            mbar_init(&B.acc_full[i], 1);
            mbar_init(&B.info_ready[i], 1);
        }
        mbar_init(&B.h_ready, FQ_EPI_WARPS);
        mbar_init(&B.acc1_free, FQ_EPI_WARPS);
        mbar_init(&B.v_empty, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(&B.tmem_ptr, 512);
    // zero the S halves of both SV stages once (kept clean by un-filling)
    for (int i = threadIdx.x; i < T_NSV * (FQ_A_BYTES / 16); i += T_THREADS) reinterpret_cast<uint4 *>(s_ring)[i] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = B.tmem_ptr;
    const uint32_t acc0 = tmem, acc1 = tmem + 256;

    int64_t my_tiles = 0;
    if ((int64_t)blockIdx.x < n_tiles) my_tiles = (n_tiles - 1 - blockIdx.x) / gridDim.x + 1;

    if (warp >= 6) {
        // ======================= builders =======================
        const int bt = threadIdx.x - 6 * 32;           // 0..511
        const int unit_in_chunk = bt & 7;
        const int brow = bt >> 2, bpair = bt & 3;      // S fill: (row, 2 of the 8 corners); the 4 threads of a row share a warp
        uint32_t filled[T_NSV][2];
#pragma unroll
        for (int a = 0; a < T_NSV; ++a) filled[a][0] = filled[a][1] = 0xffffffffu;
        uint32_t n_sv = 0, n_lv = 0;                    // S stages / tensor-core levels produced so far
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t tile = blockIdx.x + it * gridDim.x;
            const uint32_t gc0 = (uint32_t)it * (uint32_t)KC0;      // global A-chunk index of this tile's chunk 0
            TileInfo &ti = tinfo[it & 1];
            named_bar_sync(1, T_BUILD_THREADS);        // every builder is done with the previous tile's points
            if (bt < FQ_TILE) {
                float px, py, pz;
                int scene;
                int64_t oi;
                row_point(p, tile, bt, px, py, pz, scene, oi);
                pts_s[bt] = make_float4(px, py, pz, __int_as_float(oi < 0 ? -1 : scene));
            }
            if (bt < SVR_MAX_LEVELS * 8) {
                const int l = bt >> 3, k = bt & 7;
                B.red[l][k] = k < 3 ? 0x7fffffff : (k < 6 ? -0x7fffffff : (k == 6 ? 1 : -1));   // min, max, one-scene flag, scene
            }
            named_bar_sync(1, T_BUILD_THREADS);
            const float4 q = pts_s[brow];
            const int my_scene = __float_as_int(q.w);
            const int scene0 = __float_as_int(pts_s[0].w);
            // ---- per-level bounding boxes of the touched voxels (tensor-core candidate levels only)
            for (int l = 1; l < p.P.n_levels; ++l) {
                if (!tc_level(p.P, l) || my_scene < 0) continue;
                if (my_scene != scene0) B.red[l][6] = 0;      // rows of two scenes in one tile
                int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-0x7fffffff, -0x7fffffff, -0x7fffffff};
                for (int d = bpair; d < 7; d += 4) {
                    Corners c;
                    stencil_corners(p.P, l, d, q.x, q.y, q.z, c);
                    const int xa = max(c.x0, 0), xb = min(c.x0 + 1, p.P.W[l] - 1), ya = max(c.y0, 0), yb = min(c.y0 + 1, p.P.H[l] - 1);
                    const int za = max(c.z0, 0), zb = min(c.z0 + 1, p.P.D[l] - 1);
                    if (xa <= xb && ya <= yb && za <= zb) {
                        lo[0] = min(lo[0], xa); lo[1] = min(lo[1], ya); lo[2] = min(lo[2], za);
                        hi[0] = max(hi[0], xb); hi[1] = max(hi[1], yb); hi[2] = max(hi[2], zb);
                    }
                }
#pragma unroll
                for (int a = 0; a < 3; ++a)
                    if (lo[a] <= hi[a]) {
                        atomicMin(&B.red[l][a], lo[a]);
                        atomicMax(&B.red[l][3 + a], hi[a]);
                    }
            }
            named_bar_sync(1, T_BUILD_THREADS);
            if (bt < SVR_MAX_LEVELS) {
                const int l = bt;
                int mode = 0, nvc = 0;
                if (l < p.P.n_levels && tc_level(p.P, l)) {
                    const int nx = B.red[l][3] - B.red[l][0] + 1, ny = B.red[l][4] - B.red[l][1] + 1, nz = B.red[l][5] - B.red[l][2] + 1;
                    // all rows of one scene?  (min == max of the scene ids seen; padding rows are ignored)
                    if (nx > 0 && ny > 0 && nz > 0 && B.red[l][6] == 1) {
                        const int64_t nvox = (int64_t)nx * ny * nz;
                        if (nvox <= T_MAX_VC * 64) {   // the box must fit the V area
                            mode = 1;
                            nvc = (int)((nvox + 63) / 64);
                            ti.box[l][0] = B.red[l][0]; ti.box[l][1] = B.red[l][1]; ti.box[l][2] = B.red[l][2];
                            ti.box[l][3] = nx; ti.box[l][4] = ny; ti.box[l][5] = nz;
                        }
                    } else if (nx <= 0 || ny <= 0 || nz <= 0) {
                        // nothing in bounds: every feature of the level is zero; one empty voxel chunk does it
                        if (B.red[l][6] == 1) {
                            mode = 1;
                            nvc = 1;
                            ti.box[l][0] = ti.box[l][1] = ti.box[l][2] = 0;
                            ti.box[l][3] = ti.box[l][4] = ti.box[l][5] = 0;
                        }
                    }
                }
                ti.mode[l] = mode;
                ti.n_vc[l] = nvc;
                if (l < p.P.n_levels && tc_level(p.P, l)) atomicAdd(&g_fq_mode_count[l][mode], 1ull);
                if (l == 0) ti.scene = scene0 >= 0 ? scene0 : 0;
            }
            named_bar_sync(1, T_BUILD_THREADS);
            if (bt == 0) mbar_arrive(&B.info_ready[it & 1]);

            // ---- walk the K' chunks in order
            int kc = 0;
            while (kc < KC0) {
                // level of the chunk's first unit (chunks of wide levels never mix levels)
                int lvl = 0;
#pragma unroll
                for (int l = 1; l < SVR_MAX_LEVELS; ++l)
                    if (l < p.P.n_levels && kc * 8 >= p.P.ubase[l]) lvl = l;
                const bool tc = tc_level(p.P, lvl) && kc * 8 < p.P.ubase[lvl] + 7 * p.P.upd[lvl] && ti.mode[lvl];
                if (!tc) {
                    // CUDA-core gather of one chunk (2 rows per thread)
                    const uint32_t gc = gc0 + kc;
                    const int st = gc % T_NA;
                    const int u = kc * 8 + unit_in_chunk;
                    UnitCtx uc;
                    make_unit_ctx(p.P, u, p.vols.v, uc);
                    mbar_wait(&B.a_empty[st], ((gc / T_NA) & 1) ^ 1);
                    uint8_t *a_st = a_ring + st * FQ_A_BYTES;
#pragma unroll 2
                    for (int r = bt >> 3; r < FQ_TILE; r += T_BUILD_THREADS / 8) {
                        const float4 qq = pts_s[r];
                        const int scene = __float_as_int(qq.w);
                        uint4 val = make_uint4(0, 0, 0, 0);
                        if (uc.real && scene >= 0) {
                            if (uc.level > 0) {
                                val = gather_unit_fast(uc, p.P.align, qq.x, qq.y, qq.z, scene);
                            } else {
                                const float *x0b = p.x0 + (int64_t)scene * p.P.D[0] * p.P.H[0] * p.P.W[0];
                                val = gather_unit_decoded(p.P, 0, 0, 0, qq.x, qq.y, qq.z, x0b, nullptr);
                            }
                        }
                        *reinterpret_cast<uint4 *>(a_st + swz128(r, unit_in_chunk)) = val;
                        if (p.save_feat) {
                            int64_t row = tile * FQ_TILE + r;
                            if (row < p.total) *reinterpret_cast<uint4 *>(p.save_feat + row * p.P.kp + (int64_t)u * 8) = val;
                        }
                    }
                    fence_proxy_async();
                    named_bar_sync(1, T_BUILD_THREADS);
                    if (bt == 0) mbar_arrive(&B.a_full[st]);
                    ++kc;
                    continue;
                }
                // ---- tensor-core level: for every stencil point and voxel chunk build one SV stage
                const int C = p.P.C[lvl], W = p.P.W[lvl], H = p.P.H[lvl], D = p.P.D[lvl], ncg = C / 8;
                const int bx0 = ti.box[lvl][0], by0 = ti.box[lvl][1], bz0 = ti.box[lvl][2];
                const int nx = ti.box[lvl][3], ny = ti.box[lvl][4], nz = ti.box[lvl][5];
                const int nvox = nx * ny * nz, nvc = ti.n_vc[lvl];
                const __nv_bfloat16 *vol = p.vols.v[lvl] + (int64_t)ti.scene * D * H * W * C;
                // the level's box (<= 128 voxels, channel-last) is staged ONCE as the MN-major B operand of all
                // 7 stencil points; the previous tensor-core level's MMAs must be done with the V area
                mbar_wait(&B.v_empty, (n_lv & 1) ^ 1);
                ++n_lv;
                for (int i = bt; i < nvc * 64 * ncg; i += T_BUILD_THREADS) {
                    const int lid = i / ncg, ch = i - lid * ncg;
                    const int vc = lid >> 6, v = lid & 63;
                    uint4 val = make_uint4(0, 0, 0, 0);
                    if (lid < nvox) {
                        const int lx = lid % nx, ly = (lid / nx) % ny, lz = lid / (nx * ny);
                        val = __ldg(reinterpret_cast<const uint4 *>(vol + (((int64_t)(bz0 + lz) * H + (by0 + ly)) * W + (bx0 + lx)) * C + ch * 8));
                    }
                    *reinterpret_cast<uint4 *>(v_area + vc * FQ_A_BYTES + (ch >> 3) * 8192 + swz128(v, ch & 7)) = val;
                }
                for (int d = 0; d < 7; ++d) {
                    Corners c;
                    if (my_scene >= 0) stencil_corners(p.P, lvl, d, q.x, q.y, q.z, c);
                    for (int vc = 0; vc < nvc; ++vc, ++n_sv) {
                        const int st = n_sv % T_NSV;
                        uint8_t *s_st = s_ring + st * FQ_A_BYTES;
                        mbar_wait(&B.sv_empty[st], ((n_sv / T_NSV) & 1) ^ 1);
                        // un-fill what this thread wrote into this stage last time, then write the new weights
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            if (filled[st][kk] != 0xffffffffu) *reinterpret_cast<__nv_bfloat16 *>(s_st + filled[st][kk]) = __float2bfloat16(0.f);
                            filled[st][kk] = 0xffffffffu;
                        }
                        __syncwarp();   // a row's un-fills (4 adjacent lanes) land before any of its new weights
                        if (my_scene >= 0) {
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk) {
                                const int k = bpair * 2 + kk;
                                const int aa = k & 1, bb = (k >> 1) & 1, e = k >> 2;
                                const int x = c.x0 + aa, y = c.y0 + bb, z = c.z0 + e;
                                if (x < 0 || y < 0 || z < 0 || x >= W || y >= H || z >= D) continue;
                                const int lid = ((z - bz0) * ny + (y - by0)) * nx + (x - bx0) - vc * 64;
                                if (lid < 0 || lid >= 64) continue;
                                // element (m = row, k = voxel slot) of the K-major 128B-swizzled S tile
                                const uint32_t off = swz128(brow, lid >> 3) + (lid & 7) * 2;
                                *reinterpret_cast<__nv_bfloat16 *>(s_st + off) = __float2bfloat16(c.wx[aa] * c.wy[bb] * c.wz[e]);
                                filled[st][kk] = off;
                            }
                        }
                        fence_proxy_async();   // covers the V area of this level as well (first stage of the level)
                        named_bar_sync(1, T_BUILD_THREADS);
                        if (bt == 0) mbar_arrive(&B.sv_full[st]);
                    }
                }
                kc += 7 * (C / 64);
            }
        }
    } else if (warp == 4) {
        // ======================= weight loader =======================
        if (lane == 0) {
            uint32_t wc = 0;
            for (int64_t it = 0; it < my_tiles; ++it) {
                const int n_chunks = KC0 + 4 + 4;
                for (int c = 0; c < n_chunks; ++c, ++wc) {
                    const int st = wc % T_NW;
                    mbar_wait(&B.w_empty[st], ((wc / T_NW) & 1) ^ 1);
                    const uint8_t *src = c < KC0 ? p.w0_img + (size_t)c * FQ_B_BYTES
                                                 : (c < KC0 + 4 ? p.w1_img + (size_t)(c - KC0) * FQ_B_BYTES
                                                                : p.w2_img + (size_t)(c - KC0 - 4) * FQ_B_BYTES);
                    mbar_arrive_expect_tx(&B.w_full[st], FQ_B_BYTES);
                    bulk_g2s(smem_u32(w_ring + st * FQ_B_BYTES), src, FQ_B_BYTES, &B.w_full[st]);
                }
            }
        }
    } else if (warp == 5) {
        // ======================= MMA issue =======================
        if (lane == 0 && my_tiles > 0) {
            const uint32_t idesc = make_idesc_bf16(FQ_TILE, FQ_HID, 0, 0);
            uint32_t wc = 0, hr = 0, n_sv = 0, n_f = 0;
            bool acc0_started = false;
            auto fc0_chunk = [&](uint32_t gc) {      // one K chunk of fc_0: A ring stage x streamed W0 chunk -> acc0
                const int sa = gc % T_NA, sw = wc % T_NW;
                mbar_wait(&B.a_full[sa], (gc / T_NA) & 1);
                mbar_wait(&B.w_full[sw], (wc / T_NW) & 1);
                tc_fence_after();
                const uint32_t a_s = smem_u32(a_ring + sa * FQ_A_BYTES), b_s = smem_u32(w_ring + sw * FQ_B_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_bf16(acc0, make_smem_desc(a_s + k * 32, 16, 1024, kSwizzle128B), make_smem_desc(b_s + k * 32, 16, 1024, kSwizzle128B),
                              idesc, acc0_started || k != 0);
                }
                acc0_started = true;
                umma_commit(&B.a_empty[sa]);
                umma_commit(&B.w_empty[sw]);
                ++wc;
            };
            auto s_mma = [&](int lvl, int nvc, uint32_t fbuf, bool last_of_level) {   // F_d = S_d . V_box over the voxel chunks -> F buffer
                const uint32_t idesc_s = make_idesc_bf16(FQ_TILE, p.P.C[lvl], 0, 1);
                mbar_wait(&B.f_empty[fbuf], ((n_f >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int vc = 0; vc < nvc; ++vc, ++n_sv) {
                    const int st = n_sv % T_NSV;
                    mbar_wait(&B.sv_full[st], (n_sv / T_NSV) & 1);
                    tc_fence_after();
                    const uint32_t s_s = smem_u32(s_ring + st * FQ_A_BYTES), v_s = smem_u32(v_area + vc * FQ_A_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(acc1 + fbuf * 128, make_smem_desc(s_s + k * 32, 16, 1024, kSwizzle128B),
                                  make_smem_desc(v_s + k * 2048, 8192, 1024, kSwizzle128B), idesc_s, (vc | k) != 0);
                    umma_commit(&B.sv_empty[st]);
                }
                umma_commit(&B.f_full[fbuf]);
                if (last_of_level) umma_commit(&B.v_empty);
                ++n_f;
            };
            auto issue_hidden = [&]() {   // A = H tile (4 K-chunks), B = next 4 weight chunks, D = acc1
                mbar_wait(&B.h_ready, hr & 1);
                ++hr;
                tc_fence_after();
                for (int kc = 0; kc < 4; ++kc) {
                    const int sw = wc % T_NW;
                    mbar_wait(&B.w_full[sw], (wc / T_NW) & 1);
                    tc_fence_after();
                    const uint32_t a_s = smem_u32(h_tile + kc * FQ_A_BYTES), b_s = smem_u32(w_ring + sw * FQ_B_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(acc1, make_smem_desc(a_s + k * 32, 16, 1024, kSwizzle128B), make_smem_desc(b_s + k * 32, 16, 1024, kSwizzle128B),
                                  idesc, (kc | k) != 0);
                    umma_commit(&B.w_empty[sw]);
                    ++wc;
                }
                umma_commit(&B.acc_full[1]);
            };
            for (int64_t it = 0; it < my_tiles; ++it) {
                const uint32_t gc0 = (uint32_t)it * (uint32_t)KC0;
                mbar_wait(&B.info_ready[it & 1], (it >> 1) & 1);
                const TileInfo &ti = tinfo[it & 1];
                // the F buffers alias acc1: the previous tile's fc_2 result must have been read
                if (it > 0) mbar_wait(&B.acc1_free, (it - 1) & 1);
                tc_fence_after();
                acc0_started = false;
                int kc = 0;
                while (kc < KC0) {
                    int lvl = 0;
                    for (int l = 1; l < SVR_MAX_LEVELS; ++l)
                        if (l < p.P.n_levels && kc * 8 >= p.P.ubase[l]) lvl = l;
                    const bool tc = tc_level(p.P, lvl) && kc * 8 < p.P.ubase[lvl] + 7 * p.P.upd[lvl] && ti.mode[lvl];
                    if (!tc) {
                        fc0_chunk(gc0 + kc);
                        ++kc;
                        continue;
                    }
                    const int cpd = p.P.C[lvl] / 64;      // A chunks per stencil point
                    // software pipeline: S(d+1) is issued before the fc_0 chunks of d
                    s_mma(lvl, ti.n_vc[lvl], n_f & 1, false);
                    for (int d = 0; d < 7; ++d) {
                        if (d < 6) s_mma(lvl, ti.n_vc[lvl], n_f & 1, d == 5);
                        for (int j = 0; j < cpd; ++j) fc0_chunk(gc0 + kc + d * cpd + j);
                    }
                    kc += 7 * cpd;
                }
                umma_commit(&B.acc_full[0]);
                issue_hidden();
                issue_hidden();
            }
        }
        __syncwarp();
    } else {
        // ======================= epilogue =======================
        const int r = warp * 32 + lane;            // row in tile == TMEM lane
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        uint32_t n0 = 0, n1 = 0, n_f = 0;
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t tile = blockIdx.x + it * gridDim.x;
            const uint32_t gc0 = (uint32_t)it * (uint32_t)KC0;
            float px, py, pz;
            int scene;
            int64_t out_idx;
            row_point(p, tile, r, px, py, pz, scene, out_idx);
            const int64_t row = tile * FQ_TILE + r;
            const bool row_ok = p.points ? row < p.total : out_idx >= 0;
            mbar_wait(&B.info_ready[it & 1], (it >> 1) & 1);
            const TileInfo &ti = tinfo[it & 1];
            // ---- convert the tensor-core gathered F_d blocks into bf16 A chunks for fc_0
            for (int lvl = 1; lvl < p.P.n_levels; ++lvl) {
                if (!tc_level(p.P, lvl) || !ti.mode[lvl]) continue;
                const int cpd = p.P.C[lvl] / 64;
                const int kc_l = p.P.ubase[lvl] / 8;
                for (int d = 0; d < 7; ++d, ++n_f) {
                    const uint32_t fbuf = n_f & 1;
                    mbar_wait(&B.f_full[fbuf], (n_f >> 1) & 1);
                    tc_fence_after();
                    for (int j = 0; j < cpd; ++j) {
                        const uint32_t gc = gc0 + kc_l + d * cpd + j;
                        const int sa = gc % T_NA;
                        uint32_t v0[32], v1[32];
                        tmem_ld32(acc1 + fbuf * 128 + lane_off + j * 64, v0);
                        tmem_ld32(acc1 + fbuf * 128 + lane_off + j * 64 + 32, v1);
                        tmem_ld_wait();
                        mbar_wait(&B.a_empty[sa], ((gc / T_NA) & 1) ^ 1);
                        uint8_t *a_st = a_ring + sa * FQ_A_BYTES;
                        uint4 pk[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            float g[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) g[e] = __uint_as_float(u < 4 ? v0[u * 8 + e] : v1[(u - 4) * 8 + e]);
                            pk[u] = float8_to_bf16(g);
                            *reinterpret_cast<uint4 *>(a_st + swz128(r, u)) = pk[u];
                        }
                        if (p.save_feat && row_ok && p.points) {
                            __nv_bfloat16 *dst = p.save_feat + row * p.P.kp + (int64_t)(kc_l + d * cpd + j) * 64;
#pragma unroll
                            for (int u = 0; u < 8; ++u) *reinterpret_cast<uint4 *>(dst + u * 8) = pk[u];
                        }
                        fence_proxy_async();
                        named_bar_sync(2, FQ_EPI_WARPS * 32);
                        if (threadIdx.x == 0) mbar_arrive(&B.a_full[sa]);
                    }
                    tc_fence_before();
                    named_bar_sync(2, FQ_EPI_WARPS * 32);
                    if (threadIdx.x == 0) mbar_arrive(&B.f_empty[fbuf]);
                }
            }
            // ---- decoder epilogues
            float dot = 0.f;
#pragma unroll 1
            for (int layer = 0; layer < 3; ++layer) {
                const float *bias = layer == 0 ? p.b0 : (layer == 1 ? p.b1 : p.b2);
                if (layer == 0) {
                    mbar_wait(&B.acc_full[0], n0 & 1);
                    ++n0;
                } else {
                    mbar_wait(&B.acc_full[1], n1 & 1);
                    ++n1;
                }
                tc_fence_after();
                const uint32_t acc = (layer == 0 ? acc0 : acc1) + lane_off;
#pragma unroll 1
                for (int c0 = 0; c0 < FQ_HID; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(acc + c0, v);
                    tmem_ld_wait();
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = fmaxf(__uint_as_float(v[j]) + __ldg(bias + c0 + j), 0.f);
                    if (layer == 2) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) dot = fmaf(f[j], __ldg(p.wout + c0 + j), dot);
                    }
                    uint4 packed[4];
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        float g[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) g[j] = f[qd * 8 + j];
                        packed[qd] = float8_to_bf16(g);
                    }
                    if (layer < 2) {
                        uint8_t *hc = h_tile + (c0 >> 6) * FQ_A_BYTES;
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd)
                            *reinterpret_cast<uint4 *>(hc + swz128(r, ((c0 & 63) >> 3) + qd)) = packed[qd];
                    }
                    if (p.save_h && row_ok && p.points) {
                        __nv_bfloat16 *dst = p.save_h + ((int64_t)layer * p.total + row) * FQ_HID + c0;
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) *reinterpret_cast<uint4 *>(dst + qd * 8) = packed[qd];
                    }
                }
                if (layer < 2) {
                    fence_proxy_async();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&B.h_ready);
                } else {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&B.acc1_free);
                }
            }
            if (row_ok) {
                float logit = dot + __ldg(p.bout);
                if (p.apply_sigmoid) logit = 1.0f / (1.0f + __expf(-logit));
                p.out[out_idx] = logit;
            }
        }
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// [K/64] chunks of (R rows x 128 B) in the 128B-swizzled UMMA layout, from row-major bf16 (R, K)
__global__ void swizzle_image_kernel(const __nv_bfloat16 *__restrict__ src, int R, int K, uint8_t *__restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one 16-byte unit each
    int64_t units = (int64_t)R * (K / 8);
    if (i >= units) return;
    int row = (int)(i / (K / 8)), u = (int)(i % (K / 8));
    int chunk = u >> 3, uc = u & 7;
    uint4 v = *reinterpret_cast<const uint4 *>(src + (int64_t)row * K + (int64_t)u * 8);
    *reinterpret_cast<uint4 *>(dst + (size_t)chunk * R * 128 + tc::swz128(row, uc)) = v;
}

static int fq_fill(FqParams &p, const float *x0, const uint16_t *const *vols_host, const svr_pyramid *pyr_host,
                   const svr_decoder_weights *w) {
    if (int rc = make_pyr(p.P, pyr_host)) return rc;
    SVR_REQUIRE(w && x0 && vols_host, "fused query: null pointer");
    SVR_REQUIRE(w->h0 == FQ_HID && w->h1 == FQ_HID && w->h2 == FQ_HID, "fused query supports hidden size 256 only (got %d/%d/%d)",
                w->h0, w->h1, w->h2);
    SVR_REQUIRE(w->w0p && w->w1 && w->w2 && w->b0 && w->b1 && w->b2 && w->wout && w->bout, "fused query: null weight pointer");
    SVR_REQUIRE(p.P.kp / 8 <= FQ_UTAB, "fused query: feature row of %d columns exceeds the unit table (%d columns)", p.P.kp, FQ_UTAB * 8);
    for (int l = 0; l < SVR_MAX_LEVELS; ++l) {
        p.vols.v[l] = (l >= 1 && l < p.P.n_levels) ? (const __nv_bfloat16 *)vols_host[l] : nullptr;
        SVR_REQUIRE(!(l >= 1 && l < p.P.n_levels) || p.vols.v[l], "fused query: volume of level %d is null", l);
    }
    p.x0 = x0;
    p.w0_img = (const uint8_t *)w->w0p;
    p.w1_img = (const uint8_t *)w->w1;
    p.w2_img = (const uint8_t *)w->w2;
    p.b0 = w->b0;
    p.b1 = w->b1;
    p.b2 = w->b2;
    p.wout = w->wout;
    p.bout = w->bout;
    return 0;
}

static int fq_launch(const FqParams &p, int64_t n_tiles, cudaStream_t st) {
    static bool attr = false;
    static int gw = 16;
    if (!attr) {
        SVR_CUDA(cudaFuncSetAttribute(fused_query_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, FQ_SMEM));
        SVR_CUDA(cudaFuncSetAttribute(fused_query_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, FQ_SMEM));
        const char *e = getenv("SVR_FQ_GATHER_WARPS");
        if (e && atoi(e) == 16) gw = 16;
        if (e && atoi(e) == 8) gw = 8;
        attr = true;
    }
    if (n_tiles <= 0) return 0;
    int grid = sm_count();
    if (n_tiles < grid) grid = (int)n_tiles;
    static int use_tc = -1;
    if (use_tc < 0) {
        SVR_CUDA(cudaFuncSetAttribute(fused_query_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM));
        // EXPERIMENTAL, off by default: the tensor-core gather only pays off when a tile's voxel box fits the
        // 32 KB V area (<= 128 voxels); at 50k points/scene that holds for level 5 and ~10 % of the level-4 tiles,
        // and the kernel is then no faster than the CUDA-core gather.  SVR_FQ_TC=1 enables it.
        const char *e = getenv("SVR_FQ_TC");
        use_tc = (e && e[0] == '1') ? 1 : 0;
    }
    if (use_tc) {
        fused_query_tc_kernel<<<grid, T_THREADS, T_SMEM, st>>>(p, n_tiles);
        SVR_LAUNCH_CHECK();
        return 0;
    }
    if (gw == 16)
        fused_query_kernel<16><<<grid, fq_threads(16), FQ_SMEM, st>>>(p, n_tiles);
    else
        fused_query_kernel<8><<<grid, fq_threads(8), FQ_SMEM, st>>>(p, n_tiles);
    SVR_LAUNCH_CHECK();
    return 0;
}

}  // namespace svr

using namespace svr;

extern "C" {

/* diagnostics: tiles per (level, mode) seen by the tensor-core-gather kernel since the last reset;
 * out_host receives SVR_MAX_LEVELS x 2 counters (mode 0 = CUDA-core gather, 1 = tensor-core gather) */
int svr_debug_fq_modes(unsigned long long *out_host, int reset) {
    SVR_REQUIRE(out_host, "debug_fq_modes: null pointer");
    SVR_CUDA(cudaDeviceSynchronize());
    SVR_CUDA(cudaMemcpyFromSymbol(out_host, g_fq_mode_count, sizeof(unsigned long long) * SVR_MAX_LEVELS * 2));
    if (reset) {
        unsigned long long z[SVR_MAX_LEVELS * 2] = {0};
        SVR_CUDA(cudaMemcpyToSymbol(g_fq_mode_count, z, sizeof(z)));
    }
    return 0;
}

int svr_pack_decoder_image(const uint16_t *w_rowmajor, int R, int K, uint8_t *image, void *stream) {
    SVR_REQUIRE(w_rowmajor && image && R > 0 && K > 0 && K % 64 == 0 && R % 8 == 0, "pack_decoder_image: R %% 8 == 0 and K %% 64 == 0 required");
    int64_t units = (int64_t)R * (K / 8);
    swizzle_image_kernel<<<(unsigned)ceil_div<int64_t>(units, 256), 256, 0, as_stream(stream)>>>((const __nv_bfloat16 *)w_rowmajor, R, K,
                                                                                                image);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_query_fwd_fused(const float *points, const int *perm, int B, int N, const float *x0, const uint16_t *const *vols_host,
                        const svr_pyramid *pyr_host, const svr_decoder_weights *w_host, float *logits, uint16_t *save_h,
                        uint16_t *save_feat, int apply_sigmoid, void *stream) {
    FqParams p{};
    if (int rc = fq_fill(p, x0, vols_host, pyr_host, w_host)) return rc;
    SVR_REQUIRE(points && logits, "query_fwd_fused: null pointer");
    p.points = points;
    p.perm = perm;
    p.N = N;
    p.total = (int64_t)B * N;
    p.out = logits;
    p.save_h = (__nv_bfloat16 *)save_h;
    p.save_feat = (__nv_bfloat16 *)save_feat;
    p.apply_sigmoid = apply_sigmoid;
    return fq_launch(p, ceil_div<int64_t>(p.total, FQ_TILE), as_stream(stream));
}

int svr_dense_eval(int scene, int B, const float *x0, const uint16_t *const *vols_host, const svr_pyramid *pyr_host,
                   const svr_decoder_weights *w_host, int sx, int sy, int sz, int x_begin, int x_end, float *out, void *stream) {
    FqParams p{};
    if (int rc = fq_fill(p, x0, vols_host, pyr_host, w_host)) return rc;
    SVR_REQUIRE(out && scene >= 0 && scene < B, "dense_eval: bad scene index");
    SVR_REQUIRE(sx > 0 && sy > 0 && sz > 0 && x_begin >= 0 && x_end <= sx && x_begin <= x_end, "dense_eval: bad lattice range");
    p.points = nullptr;
    p.lat_scene = scene;
    p.sx = sx;
    p.sy = sy;
    p.sz = sz;
    p.x_begin = x_begin;
    p.bx = ceil_div(x_end - x_begin, BRICK_X);
    p.by = ceil_div(sy, BRICK_Y);
    p.bz = ceil_div(sz, BRICK_Z);
    // rows beyond x_end inside the last brick must not be written: clamp through sx
    p.sx = sx;
    p.total = (int64_t)p.bx * p.by * p.bz * FQ_TILE;
    p.out = out;
    p.apply_sigmoid = 1;
    if (x_end < sx) {
        // the brick grid may overhang x_end; mask by shrinking the visible lattice extent
        // (coordinates still use the full sx through lin_coord's n argument)
        SVR_REQUIRE((x_end - x_begin) % BRICK_X == 0, "dense_eval: slab length must be a multiple of %d unless it ends the lattice", BRICK_X);
    }
    return fq_launch(p, (int64_t)p.bx * p.by * p.bz, as_stream(stream));
}
}
