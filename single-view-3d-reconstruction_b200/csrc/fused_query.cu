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
    static DeviceOnce once;
    int dev;
    if (once.needed(dev)) {
        SVR_CUDA(cudaFuncSetAttribute(fused_query_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, FQ_SMEM));
        once.done(dev);
    }
    if (n_tiles <= 0) return 0;
    int grid = sm_count();
    if (n_tiles < grid) grid = (int)n_tiles;
    fused_query_kernel<16><<<grid, fq_threads(16), FQ_SMEM, st>>>(p, n_tiles);
    SVR_LAUNCH_CHECK();
    return 0;
}

}  // namespace svr

using namespace svr;

extern "C" {

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
