// Fused IF-Net query forward: multi-scale trilinear stencil gather -> shared memory (128B-swizzled
// UMMA operand tiles) -> tcgen05 fc_0 -> fc_1 -> fc_2 -> fc_out, ONE persistent kernel.
// Replaces model/ifnet.py:38-61 + :156-197 (6x F.grid_sample, cat, reshape, 4x Conv1d) without ever
// materialising the (B, 2583, N) feature tensor.  Also serves the dense-grid evaluator
// (evaluate_network_on_grid, ifnet.py:215-229) by generating the make_3d_grid lattice on the fly in
// brick order.
//
// One CTA per SM, 128 query points per tile, 704 threads:
//   warps 0-3   epilogue: TMEM -> registers -> (+bias, ReLU, bf16) -> H tile in smem / logits
//   warp  4     weight loader: cp.async.bulk (UBLKCP) of pre-swizzled 32 KB weight chunks
//   warp  5     single-thread tcgen05.mma issue
//   warps 6-21  gather producers: 8 corners x 16 B channel-last loads per (point, unit), trilinear
//               blend in fp32, bf16 pack, st.shared into the swizzled A stage.
//               Two code paths: the generic one (any level, bounds-checked corners) and the WIDE path for levels
//               with C % 64 == 0 whose volumes come with a one-voxel zero halo (svr_pack_volume_halo): a sample
//               descriptor (corner base pointer + 8 trilinear weights, no bounds logic: the halo supplies the
//               zeros of grid_sample's zero padding) is computed once per (row, stencil point) and reused for all
//               the level's channel groups of the thread (C/64 consecutive K chunks).
// TMEM (512 columns): [0,256) the fp32 accumulator of whichever layer is running, [256,384) H0 = relu(fc_0) and
// [384,512) H1 = relu(fc_1) as packed bf16 pairs: the hidden activations never touch shared memory -- the epilogue writes
// them with tcgen05.st and fc_1 / fc_2 read them as the TMEM A operand of tcgen05.mma (the 64 KB H tile of round 1 is
// gone: shared memory 221 -> 157 KB, which the hardware hands to the L1 data cache the gather loads go through).
#include "common.cuh"
#include "sampling.cuh"
#include "tc05.cuh"

namespace svr {
using namespace tc;

constexpr int FQ_TILE = 128;
constexpr int FQ_HID = 256;
constexpr int FQ_NA = 3, FQ_NB = 2, FQ_NB_MAX = 4;      // FQ_NB: default weight-ring depth (p.nb at run time; 2: 1.03 ms, 3: 1.04, 4: 1.10 -- L1 capacity)
constexpr int FQ_A_BYTES = FQ_TILE * 128;        // 16 KB: 128 rows x 64 bf16
constexpr int FQ_B_BYTES = FQ_HID * 128;         // 32 KB: 256 rows x 64 bf16
constexpr int FQ_BIAS_BYTES = 4 * FQ_HID * 4;    // b0, b1, b2, wout as fp32 in shared memory (epilogue operands)
constexpr int FQ_EPI_WARPS = 4;
constexpr int fq_threads(int gather_warps) { return (FQ_EPI_WARPS + 2 + gather_warps) * 32; }   // 448 for 8 gather warps
constexpr int FQ_UTAB = 512;                     // unit table entries (KP <= 4096)
constexpr int FQ_WGEO_BYTES = 512;               // per-level geometry of the wide path
// Shared memory is kept SMALL on purpose: what a CTA does not request stays L1 data cache (228 KB - shared memory per SM),
// and the gather lives on L1 hits -- neighbouring (spatially sorted) rows and the 7 stencil points of a row read the same
// voxels.  Staging the corner loads through shared memory (cp.async, one 128-byte slot per thread: latency fully hidden, no
// registers in flight) was measured SLOWER (1.39 vs 1.02 ms at config 2) because its 64 KB shrink L1 to a few KB.  A warp
// that prefetches the next tile's fine-level sectors into L2 (prefetch.global.L2) was also slower (1.08 vs 0.98 ms), and so
// were fp32 halo'd copies of the wide levels (the blend then needs no bf16 unpacking -- 46 instead of 110 instructions per
// unit -- but reads twice the bytes through L1: 1.54 vs 1.08 ms).
constexpr int fq_smem(int nb) { return 1024 + FQ_NA * FQ_A_BYTES + nb * FQ_B_BYTES + FQ_BIAS_BYTES + 2 * FQ_TILE * 16 + 512 + FQ_UTAB * 4 + FQ_WGEO_BYTES; }
constexpr int FQ_SMEM = fq_smem(FQ_NB_MAX);

// geometry of one WIDE level (C % 64 == 0) sampled from its halo'd copy (B, D+2, H+2, W+2, C), see the header comment
struct WideGeo {
    float fw, fh, fd;              // unpadded sizes
    int C, sy, sz;                 // element strides of a y / z step in the halo'd volume: (W+2)*C, (H+2)*(W+2)*C
    int cpd;                       // K chunks per stencil point (C / 64)
    int chunk0;                    // first K chunk of the level
    int spec;                      // compile-time specialisation id of (C, sy, sz); 0 = runtime strides
    long long scene;               // halo'd elements per scene
    const __nv_bfloat16 *base;     // halo'd volume
};
static_assert(sizeof(WideGeo) * SVR_MAX_LEVELS <= FQ_WGEO_BYTES, "wide geometry table");

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
    FqVols halo;                // optional halo'd copies of the wide levels (null entries: generic path)
    int wide_level0;            // first level taken by the wide path (== P.n_levels: none); every level from here on is wide
    Pyr P;
    const uint8_t *w0_img, *w1_img, *w2_img;   // pre-swizzled chunk images (svr_pack_decoder_images)
    const float *b0, *b1, *b2, *wout, *bout;
    float *out;                 // logits (explicit) or occupancy grid (lattice)
    __nv_bfloat16 *save_h;      // optional (3, total, 256)
    __nv_bfloat16 *save_feat;   // optional (total, KP)
    int apply_sigmoid;
    int nb;                     // weight-ring depth (2..FQ_NB_MAX)
    long long *trace;           // debug: per-role (tag, SM clock) records of block 0 (svr_debug_fq_trace), else null
};

// role 0: first gather warp, 1: MMA thread, 2: first epilogue warp, 3: weight loader
__device__ __forceinline__ void fq_trace(const FqParams &p, int role, int &n, int tag) {
    if (p.trace && blockIdx.x == 0 && n < 1024) {
        p.trace[(role * 1024 + n) * 2] = tag;
        p.trace[(role * 1024 + n) * 2 + 1] = clock64();
        ++n;
    }
}

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
    uint8_t *a, *b;
    float *bias;                 // [4][256]: b0, b1, b2, wout
    float4 *pts;                 // [2][128] : (px,py,pz, scene as int bits)
    uint64_t *a_full, *a_empty, *b_full, *b_empty, *acc_full, *h_ready, *acc_free;
    uint32_t *tmem_ptr;
    uint32_t *utab;              // [FQ_UTAB] packed decode_unit results
    WideGeo *wgeo;               // [SVR_MAX_LEVELS]
};

__device__ __forceinline__ FqSmem fq_carve(uint8_t *raw, int nb) {
    FqSmem s;
    uint8_t *base = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    s.a = base;
    s.b = s.a + FQ_NA * FQ_A_BYTES;
    s.bias = (float *)(s.b + nb * FQ_B_BYTES);
    s.pts = (float4 *)((uint8_t *)s.bias + FQ_BIAS_BYTES);
    uint64_t *bars = (uint64_t *)(s.pts + 2 * FQ_TILE);
    s.a_full = bars;
    s.a_empty = s.a_full + FQ_NA;
    s.b_full = s.a_empty + FQ_NA;
    s.b_empty = s.b_full + FQ_NB_MAX;
    s.acc_full = s.b_empty + FQ_NB_MAX;   // [1] accumulator complete (every layer)
    s.h_ready = s.acc_full + 2;       // [1] hidden activations written to TMEM
    s.acc_free = s.h_ready + 1;       // [1] accumulator drained by the last epilogue of a tile
    s.tmem_ptr = (uint32_t *)(s.acc_free + 1);
    s.utab = (uint32_t *)((uint8_t *)bars + 512);
    s.wgeo = (WideGeo *)((uint8_t *)s.utab + FQ_UTAB * 4);
    return s;
}

// ---- wide path ------------------------------------------------------------------------------------------------
// Sample descriptor of one (row, stencil point) on a halo'd level: pointer to corner (z0, y0, x0) of channel 0 and the 8
// trilinear weights.  Same index arithmetic and the same weights as gather_unit_fast; a corner outside the volume reads
// a zero from the halo (fmaf(0, w, acc) == acc: the bits of the bounds-checked path), a sample whose cell lies entirely
// outside (or a padding row) gets zero weights and a clamped in-range pointer.
__device__ __forceinline__ void wide_desc(const WideGeo &G, int align, float dx, float dy, float dz, const float4 q,
                                          const __nv_bfloat16 *&ptr, float (&w)[8]) {
    const int scene = __float_as_int(q.w);
    const float ix = unnorm(__fadd_rn(__fmul_rn(2.0f, q.z), dx), G.fw, align);
    const float iy = unnorm(__fadd_rn(__fmul_rn(2.0f, q.y), dy), G.fh, align);
    const float iz = unnorm(__fadd_rn(__fmul_rn(2.0f, q.x), dz), G.fd, align);
    const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    const bool valid = scene >= 0 && fx >= -1.0f && fx <= G.fw - 1.0f && fy >= -1.0f && fy <= G.fh - 1.0f && fz >= -1.0f &&
                       fz <= G.fd - 1.0f;                                        // NaN coordinates compare false
    const int x0 = (int)fminf(fmaxf(fx, -1.0f), G.fw - 1.0f) + 1;               // halo coordinates; fmaxf(NaN, -1) = -1
    const int y0 = (int)fminf(fmaxf(fy, -1.0f), G.fh - 1.0f) + 1;
    const int z0 = (int)fminf(fmaxf(fz, -1.0f), G.fd - 1.0f) + 1;
    const float wx1 = ix - fx, wx0 = (fx + 1.0f) - ix;
    const float wy1 = iy - fy, wy0 = (fy + 1.0f) - iy;
    const float wz1 = valid ? iz - fz : 0.f, wz0 = valid ? (fz + 1.0f) - iz : 0.f;
    const float wxy[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = wxy[k & 3] * ((k & 4) ? wz1 : wz0);
    ptr = G.base + (long long)(scene < 0 ? 0 : scene) * G.scene + (z0 * G.sz + y0 * G.sy + x0 * G.C);
}

// 8 channels (one 16-byte unit) of one sample: 8 corner loads at compile-time (OC != 0) or run-time offsets, FFMA2 blend
// in corner order from a zero accumulator (the arithmetic of gather_unit_fast), bf16 pack
template <int OC, int OY, int OZ>
__device__ __forceinline__ uint4 wide_blend(const __nv_bfloat16 *ptr, const float (&w)[8], int oc, int oy, int oz) {
    const int ex = OC ? OC : oc, ey = OC ? OY : oy, ez = OC ? OZ : oz;
    uint4 raw[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
        raw[k] = __ldg(reinterpret_cast<const uint4 *>(ptr + (((k & 1) ? ex : 0) + ((k & 2) ? ey : 0) + ((k & 4) ? ez : 0))));
    unsigned long long acc[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        ffma2(acc[0], raw[k].x, w[k]);
        ffma2(acc[1], raw[k].y, w[k]);
        ffma2(acc[2], raw[k].z, w[k]);
        ffma2(acc[3], raw[k].w, w[k]);
    }
    uint32_t out[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
        __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
        out[i] = *reinterpret_cast<uint32_t *>(&h);
    }
    return make_uint4(out[0], out[1], out[2], out[3]);
}

// halo'd strides of the 128-net's wide levels on a 128^3 scene: 64 ch @ 32^3, 128 ch @ 16^3, 128 ch @ 8^3
constexpr int WS1_C = 64, WS1_Y = 34 * 64, WS1_Z = 34 * 34 * 64;
constexpr int WS2_C = 128, WS2_Y = 18 * 128, WS2_Z = 18 * 18 * 128;
constexpr int WS3_C = 128, WS3_Y = 10 * 128, WS3_Z = 10 * 10 * 128;

template <int FQ_GATHER_WARPS>
__global__ void __launch_bounds__(fq_threads(FQ_GATHER_WARPS), 1) fused_query_kernel(const FqParams p, int64_t n_tiles) {
    constexpr int FQ_GATHER_THREADS = FQ_GATHER_WARPS * 32;
    extern __shared__ uint8_t smem_raw[];
    const FqSmem s = fq_carve(smem_raw, p.nb);
    const int NB = p.nb;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KC0 = p.P.kp / 64;

    if (threadIdx.x == 0) {
        for (int i = 0; i < FQ_NA; ++i) {
            mbar_init(s.a_full + i, FQ_GATHER_WARPS);
            mbar_init(s.a_empty + i, 1);
        }
        for (int i = 0; i < FQ_NB_MAX; ++i) {
            mbar_init(s.b_full + i, 1);
            mbar_init(s.b_empty + i, 1);
        }
        mbar_init(s.acc_full + 0, 1);
        mbar_init(s.acc_full + 1, 1);
        mbar_init(s.h_ready, FQ_EPI_WARPS);
        mbar_init(s.acc_free, FQ_EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(s.tmem_ptr, 512);
    for (int u = threadIdx.x; u < KC0 * 8; u += blockDim.x) s.utab[u] = pack_unit(p.P, u);
    for (int i = threadIdx.x; i < 4 * FQ_HID; i += blockDim.x) {
        const float *src = i < FQ_HID ? p.b0 : (i < 2 * FQ_HID ? p.b1 : (i < 3 * FQ_HID ? p.b2 : p.wout));
        s.bias[i] = src[i & (FQ_HID - 1)];
    }
    if (threadIdx.x >= 32 && threadIdx.x < 32 + SVR_MAX_LEVELS) {
        const int l = threadIdx.x - 32;
        WideGeo g{};
        if (l >= p.wide_level0 && l < p.P.n_levels) {
            g.fw = (float)p.P.W[l];
            g.fh = (float)p.P.H[l];
            g.fd = (float)p.P.D[l];
            g.C = p.P.C[l];
            g.sy = (p.P.W[l] + 2) * g.C;
            g.sz = (p.P.H[l] + 2) * g.sy;
            g.cpd = g.C / 64;
            g.chunk0 = p.P.ubase[l] / 8;
            g.scene = (long long)(p.P.D[l] + 2) * g.sz;
            g.base = p.halo.v[l];
            if (g.C == WS1_C && g.sy == WS1_Y && g.sz == WS1_Z) g.spec = 1;
            if (g.C == WS2_C && g.sy == WS2_Y && g.sz == WS2_Z) g.spec = 2;
            if (g.C == WS3_C && g.sy == WS3_Y && g.sz == WS3_Z) g.spec = 3;
        }
        s.wgeo[l] = g;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s.tmem_ptr;
    const uint32_t acc0 = tmem, tm_h0 = tmem + 256, tm_h1 = tmem + 384;

    // number of tiles of this CTA
    int64_t my_tiles = 0;
    if ((int64_t)blockIdx.x < n_tiles) my_tiles = (n_tiles - 1 - blockIdx.x) / gridDim.x + 1;

    if (warp >= 6) {
        // ======================= gather producers =======================
        const int gt = threadIdx.x - 6 * 32;          // 0..511
        const int unit_in_chunk = gt & 7;
        uint32_t gc = 0;                              // global A-chunk counter
        int tn = 0;
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
            // ---- generic chunks (levels without a halo'd copy, alignment padding, the tail) -----------------------
            auto generic_chunk = [&](int kc) {
                const int st = gc % FQ_NA;
                const int u = kc * 8 + unit_in_chunk;
                UnitCtx uc;
                make_unit_ctx_packed(p.P, s.utab[u], p.vols.v, uc);
                auto gather_one = [&](float qx, float qy, float qz, int scene) -> uint4 {
                    return gather_unit_bf(uc, p.P.align, qx, qy, qz, scene);
                };
                mbar_wait(s.a_empty + st, ((gc / FQ_NA) & 1) ^ 1);
                if (gt == 0) fq_trace(p, 0, tn, 100 + kc);
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
                if (gt == 0) fq_trace(p, 0, tn, 200 + kc);
                ++gc;
            };
            const int wide_c0 = p.wide_level0 < p.P.n_levels ? p.P.ubase[p.wide_level0] / 8 : KC0;
            int kc = 0;
            for (; kc < wide_c0; ++kc) generic_chunk(kc);
            // ---- wide levels: descriptor per (row, stencil point), reused over the level's channel groups ----------
            constexpr int ROWS = FQ_TILE / (FQ_GATHER_THREADS / 8);
            for (int l = p.wide_level0; l < p.P.n_levels; ++l) {
                const WideGeo G = s.wgeo[l];
#pragma unroll 1
                for (int d = 0; d < 7; ++d) {
                    const float sgn = (d & 1) ? -p.P.delta : p.P.delta;
                    const float dx = (d == 1 || d == 2) ? sgn : 0.f, dy = (d == 3 || d == 4) ? sgn : 0.f, dz = (d == 5 || d == 6) ? sgn : 0.f;
                    const __nv_bfloat16 *ptr[ROWS];
                    float w[ROWS][8];
#pragma unroll
                    for (int j = 0; j < ROWS; ++j)
                        wide_desc(G, p.P.align, dx, dy, dz, pts_s[(gt >> 3) + j * (FQ_GATHER_THREADS / 8)], ptr[j], w[j]);
#pragma unroll 1
                    for (int h = 0; h < G.cpd; ++h, ++kc, ++gc) {
                        const int st = gc % FQ_NA;
                        const int goff = (h * 8 + unit_in_chunk) * 8;          // first channel of this thread's group
                        mbar_wait(s.a_empty + st, ((gc / FQ_NA) & 1) ^ 1);
                        if (gt == 0) fq_trace(p, 0, tn, 100 + kc);
                        uint8_t *a_st = s.a + st * FQ_A_BYTES;
#pragma unroll
                        for (int j = 0; j < ROWS; ++j) {
                            const int r = (gt >> 3) + j * (FQ_GATHER_THREADS / 8);
                            uint4 val;
                            switch (G.spec) {
                                case 1: val = wide_blend<WS1_C, WS1_Y, WS1_Z>(ptr[j] + goff, w[j], 0, 0, 0); break;
                                case 2: val = wide_blend<WS2_C, WS2_Y, WS2_Z>(ptr[j] + goff, w[j], 0, 0, 0); break;
                                case 3: val = wide_blend<WS3_C, WS3_Y, WS3_Z>(ptr[j] + goff, w[j], 0, 0, 0); break;
                                default: val = wide_blend<0, 0, 0>(ptr[j] + goff, w[j], G.C, G.sy, G.sz); break;
                            }
                            *reinterpret_cast<uint4 *>(a_st + swz128(r, unit_in_chunk)) = val;
                            if (p.save_feat) {
                                const int64_t row = tile * FQ_TILE + r;
                                if (row < p.total)
                                    *reinterpret_cast<uint4 *>(p.save_feat + row * p.P.kp + (int64_t)(kc * 8 + unit_in_chunk) * 8) = val;
                            }
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(s.a_full + st);
                        if (gt == 0) fq_trace(p, 0, tn, 200 + kc);
                    }
                }
            }
            for (; kc < KC0; ++kc) generic_chunk(kc);      // zero padding up to KP
        }
    } else if (warp == 4) {
        // ======================= weight loader =======================
        if (lane == 0) {
            uint32_t wc = 0;
            for (int64_t it = 0; it < my_tiles; ++it) {
                const int n_chunks = KC0 + 4 + 4;
                for (int c = 0; c < n_chunks; ++c, ++wc) {
                    const int st = wc % NB;
                    mbar_wait(s.b_empty + st, ((wc / NB) & 1) ^ 1);
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
            int tn = 0;
            auto wait_b = [&]() {
                const int st = wc % NB;
                mbar_wait(s.b_full + st, (wc / NB) & 1);
                return st;
            };
            auto issue_f0 = [&](int64_t it) {
                for (int kc = 0; kc < KC0; ++kc, ++gc) {
                    const int sa = gc % FQ_NA;
                    mbar_wait(s.a_full + sa, (gc / FQ_NA) & 1);
                    fq_trace(p, 1, tn, 300 + kc);
                    const int sb = wait_b();
                    fq_trace(p, 1, tn, 400 + kc);
                    if (kc == 0 && it > 0) mbar_wait(s.acc_free, (uint32_t)(it - 1) & 1);   // previous tile's logits epilogue drained the accumulator
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
                umma_commit(s.acc_full);
            };
            auto issue_hidden = [&](uint32_t tm_a) {   // A = hidden activations in TMEM (K = 256: 128 columns), B = next 4 weight chunks
                mbar_wait(s.h_ready, hr & 1);          // the epilogue has also finished READING the accumulator these MMAs overwrite
                fq_trace(p, 1, tn, 500 + (int)(hr & 1));
                ++hr;
                tc_fence_after();
                for (int kc = 0; kc < 4; ++kc) {
                    const int sb = wait_b();
                    tc_fence_after();
                    const uint32_t b_s = smem_u32(s.b + sb * FQ_B_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ts(acc0, tm_a + (uint32_t)(kc * 4 + k) * 8, make_smem_desc(b_s + k * 32, 16, 1024, kSwizzle128B), idesc,
                                     (kc | k) != 0);
                    umma_commit(s.b_empty + sb);
                    ++wc;
                }
                umma_commit(s.acc_full);
            };
            // The weight loader streams W0(i), W1(i), W2(i) per tile in that order; the issue order consumes them in the
            // same order: F0(i), F1(i), F2(i).
            for (int64_t it = 0; it < my_tiles; ++it) {
                issue_f0(it);
                issue_hidden(tm_h0);
                issue_hidden(tm_h1);
            }
        }
        __syncwarp();
    } else {
        // ======================= epilogue =======================
        const int r = warp * 32 + lane;            // row in tile == TMEM lane
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        uint32_t n_acc = 0;                        // completions consumed of acc_full
        int tn = 0;
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
                const float4 *bias4 = reinterpret_cast<const float4 *>(s.bias + layer * FQ_HID);
                const float4 *wout4 = reinterpret_cast<const float4 *>(s.bias + 3 * FQ_HID);
                mbar_wait(s.acc_full, n_acc & 1);
                ++n_acc;
                tc_fence_after();
                if (threadIdx.x == 0) fq_trace(p, 2, tn, 600 + layer);
                const uint32_t acc = acc0 + lane_off;
                const uint32_t tm_dst = (layer == 0 ? tm_h0 : tm_h1) + lane_off;
#pragma unroll 1
                for (int c0 = 0; c0 < FQ_HID; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(acc + c0, v);
                    tmem_ld_wait();
                    float f[32];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {          // shared-memory broadcast reads (every lane the same address)
                        const float4 b4 = bias4[(c0 >> 2) + q];
                        f[4 * q + 0] = fmaxf(__uint_as_float(v[4 * q + 0]) + b4.x, 0.f);
                        f[4 * q + 1] = fmaxf(__uint_as_float(v[4 * q + 1]) + b4.y, 0.f);
                        f[4 * q + 2] = fmaxf(__uint_as_float(v[4 * q + 2]) + b4.z, 0.f);
                        f[4 * q + 3] = fmaxf(__uint_as_float(v[4 * q + 3]) + b4.w, 0.f);
                    }
                    if (layer == 2) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 w4 = wout4[(c0 >> 2) + q];
                            dot = fmaf(f[4 * q + 0], w4.x, dot);
                            dot = fmaf(f[4 * q + 1], w4.y, dot);
                            dot = fmaf(f[4 * q + 2], w4.z, dot);
                            dot = fmaf(f[4 * q + 3], w4.w, dot);
                        }
                    }
                    uint32_t packed[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                        packed[j] = *reinterpret_cast<uint32_t *>(&h);
                    }
                    if (layer < 2) tmem_st16(tm_dst + (c0 >> 1), packed);     // K elements (c0 .. c0+31) -> 16 packed columns
                    if (p.save_h && row_ok && p.points) {
                        __nv_bfloat16 *dst = p.save_h + ((int64_t)layer * p.total + row) * FQ_HID + c0;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<uint4 *>(dst + q * 8) = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
                    }
                }
                if (layer < 2) {
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s.h_ready);
                } else {
                    tc_fence_before();                  // accumulator reads complete: the next tile's fc_0 may overwrite it
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s.acc_free);
                }
                if (threadIdx.x == 0) fq_trace(p, 2, tn, 610 + layer);
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

static int fq_fill(FqParams &p, const float *x0, const uint16_t *const *vols_host, const uint16_t *const *halo_host,
                   const svr_pyramid *pyr_host, const svr_decoder_weights *w) {
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
    // wide path: the trailing run of levels with C % 64 == 0 (chunk-aligned by make_pyr) that come with a halo'd copy
    p.wide_level0 = p.P.n_levels;
    for (int l = 0; l < SVR_MAX_LEVELS; ++l) p.halo.v[l] = nullptr;
    if (halo_host) {
        for (int l = p.P.n_levels - 1; l >= 1; --l) {
            if (p.P.C[l] % 64 != 0 || !halo_host[l]) break;
            SVR_REQUIRE(((uintptr_t)halo_host[l] & 15) == 0, "fused query: halo volume of level %d is not 16-byte aligned", l);
            SVR_REQUIRE((int64_t)(p.P.D[l] + 2) * (p.P.H[l] + 2) * (p.P.W[l] + 2) * p.P.C[l] < ((int64_t)1 << 31),
                        "fused query: halo volume of level %d exceeds 2^31 elements per scene", l);
            p.halo.v[l] = (const __nv_bfloat16 *)halo_host[l];
            p.wide_level0 = l;
        }
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

static long long *g_fq_trace = nullptr;
// 0: every level gathered on the CUDA cores (this file's kernel).  1 (default): the dense evaluator runs the box kernel
// (fused_query_box.cu: coarse levels interpolated on the tensor cores from voxel boxes staged in shared memory).
// 2: explicit query points take the box kernel too when they come with the sort-cell table.  3: like 1, voxel boxes staged
// by cp.async instead of TMA tensor copies (ablation).
static int g_fq_interp = 1;

namespace fqb {
void set_interp(int on, int tma);
void set_trace(long long *buf, int block);
int query_fwd(const float *points, const int *perm, const int *cell_start, int B, int N, const float *x0,
              const uint16_t *const *vols_host, const uint16_t *const *halo_vols_host, const svr_pyramid *pyr_host,
              const svr_decoder_weights *w_host, float *logits, uint16_t *save_h, uint16_t *save_feat, int apply_sigmoid, void *stream);
int dense_eval(int scene, int B, const float *x0, const uint16_t *const *vols_host, const uint16_t *const *halo_vols_host,
               const svr_pyramid *pyr_host, const svr_decoder_weights *w_host, int sx, int sy, int sz, int x_begin, int x_end, float *out,
               void *stream);
}  // namespace fqb

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
    const_cast<FqParams &>(p).trace = g_fq_trace;
    const_cast<FqParams &>(p).nb = FQ_NB;
    fused_query_kernel<16><<<grid, fq_threads(16), fq_smem(FQ_NB), st>>>(p, n_tiles);
    SVR_LAUNCH_CHECK();
    return 0;
}

}  // namespace svr

using namespace svr;

extern "C" {

/* debug: per-role (tag, SM clock) records of block 0 of the next fused query launches; buf = 4 x 1024 x 2 int64 (device), null = off */
int svr_debug_fq_trace(void *buf) {
    g_fq_trace = (long long *)buf;
    fqb::set_trace((long long *)buf, 0);
    return 0;
}

int svr_debug_fq_trace_block(int block) {
    fqb::set_trace(g_fq_trace, block);
    return 0;
}

int svr_debug_fq_interp(int mode) {
    g_fq_interp = mode == 3 ? 1 : (mode == 4 ? 2 : mode);     // 4: like 2 with cp.async staging
    fqb::set_interp(mode != 0, mode != 3 && mode != 4);
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

int svr_query_fwd_fused(const float *points, const int *perm, const int *cell_start, int B, int N, const float *x0,
                        const uint16_t *const *vols_host, const uint16_t *const *halo_vols_host, const svr_pyramid *pyr_host,
                        const svr_decoder_weights *w_host, float *logits, uint16_t *save_h, uint16_t *save_feat, int apply_sigmoid,
                        void *stream) {
    SVR_REQUIRE(!cell_start || perm, "query_fwd_fused: cell_start needs the order it belongs to");
    if (g_fq_interp == 2 && cell_start)
        return fqb::query_fwd(points, perm, cell_start, B, N, x0, vols_host, halo_vols_host, pyr_host, w_host, logits, save_h, save_feat,
                              apply_sigmoid, stream);
    FqParams p{};
    if (int rc = fq_fill(p, x0, vols_host, halo_vols_host, pyr_host, w_host)) return rc;
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

int svr_dense_eval(int scene, int B, const float *x0, const uint16_t *const *vols_host, const uint16_t *const *halo_vols_host,
                   const svr_pyramid *pyr_host, const svr_decoder_weights *w_host, int sx, int sy, int sz, int x_begin, int x_end,
                   float *out, void *stream) {
    if (g_fq_interp) return fqb::dense_eval(scene, B, x0, vols_host, halo_vols_host, pyr_host, w_host, sx, sy, sz, x_begin, x_end, out, stream);
    FqParams p{};
    if (int rc = fq_fill(p, x0, vols_host, halo_vols_host, pyr_host, w_host)) return rc;
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
