// Box variant of the fused IF-Net query forward (see fused_query.cu for the all-gather kernel that serves explicit query
// points): multi-scale trilinear stencil gather -> shared memory (128B-swizzled
// UMMA operand tiles) -> tcgen05 fc_0 -> fc_1 -> fc_2 -> fc_out, ONE persistent kernel.
// Replaces model/ifnet.py:38-61 + :156-197 (6x F.grid_sample, cat, reshape, 4x Conv1d) without ever
// materialising the (B, 2583, N) feature tensor.  Also serves the dense-grid evaluator
// (evaluate_network_on_grid, ifnet.py:215-229) by generating the make_3d_grid lattice on the fly in
// brick order.
//
// One CTA per SM, up to 128 query points per tile, 704 threads:
//   warps 0-3   epilogue: TMEM -> registers -> (+bias, ReLU, bf16) -> H in TMEM / logits; bf16 rounding of the
//               tensor-core interpolated features
//   warp  4     weight loader: cp.async.bulk (UBLKCP) of pre-swizzled 32 KB weight chunks
//   warp  5     single-thread tcgen05.mma issue
//   warps 6-21  producers of the A operand of fc_0.  Three ways, chosen per level (and per tile):
//     GENERIC   8 corners x 16 B channel-last loads per (point, unit), trilinear blend in fp32, bf16 pack, st.shared into
//               the swizzled A stage (any level, clamped corners / zeroed weights, branch-free);
//     WIDE      levels with C % 64 == 0 that come with a one-voxel zero halo (svr_pack_volume_halo): corner pointer + 8
//               weights per (row, stencil point), no bounds logic (the halo supplies grid_sample's zero padding);
//     TENSOR-CORE INTERPOLATION of the coarse levels: rows are spatially sorted and a tile never straddles two groups
//               of sort cells (tiles.cuh), so on a 16^3 / 8^3 level (for the bricks of the dense evaluator also the 32^3
//               level) the 128 rows x 7 stencil points of a tile touch a box of <= 256 voxels.  The box (bf16,
//               channel-last) is staged ONCE per tile in shared memory as an MN-major UMMA B operand; per stencil point
//               the producers only write the 8 trilinear weights of every row into a zeroed (128 x box) K-major A tile,
//               one tcgen05.mma chain computes the 128 x C features of that stencil point into TMEM, the epilogue warps
//               round them to bf16 in TMEM (and store them for the backward), and fc_0 consumes them as a TMEM A operand.
//               That takes 224 of the 323 gathered units of a point (69 %) off the CUDA cores and the L1.
// TMEM (512 columns): [0,256) the fp32 accumulator of whichever layer is running.  While fc_0 runs: [256,384) and
// [384,512) two fp32 interpolation accumulators, each overwritten in place by its bf16 rounding (first 64 columns, the
// A operand of fc_0).  Afterwards [256,384) H0 = relu(fc_0) and
// [384,512) H1 = relu(fc_1) as packed bf16 pairs: the hidden activations never touch shared memory -- the epilogue writes
// them with tcgen05.st and fc_1 / fc_2 read them as the TMEM A operand of tcgen05.mma.
//
// Per tile the work is a static schedule of three kinds of steps, written by the prep warp and followed by every role:
//   G(i)  fc_0 += A chunk i (gathered) x W0 chunk            I(r)  interpolation of round r = (level, stencil point)
//   F(r)  fc_0 += bf16(I(r)) x W0 chunks of (level, stencil point)
//   I(0), I(1), F(0), I(2), F(1), ... with the G steps spread evenly in between: round r+1 is interpolated while the
//   epilogue warps round the features of round r to bf16 IN PLACE (two 128-column TMEM buffers), and the gathers hide
//   both.
#include "common.cuh"
#include "sampling.cuh"
#include "tc05.cuh"
#include "tiles.cuh"

namespace svr {
namespace fqb {
using namespace tc;

constexpr int FQ_TILE = 128;
constexpr int FQ_HID = 256;
constexpr int FQ_NA = 5, FQ_NB = 2, FQ_NB_MAX = 4;      // A ring: a round of the 32^3 level takes two stages, and the producers must be able to build the next round meanwhile
constexpr int FQ_A_BYTES = FQ_TILE * 128;        // 16 KB: 128 rows x 64 bf16
constexpr int FQ_B_BYTES = FQ_HID * 128;         // 32 KB: 256 rows x 64 bf16
constexpr int FQ_BIAS_BYTES = 4 * FQ_HID * 4;    // b0, b1, b2, wout as fp32 in shared memory (epilogue operands)
constexpr int FQ_EPI_WARPS = 4;
constexpr int FQ_GATHER_WARPS = 16;
constexpr int FQ_GATHER_THREADS = FQ_GATHER_WARPS * 32;
constexpr int FQ_PREP_WARP = FQ_EPI_WARPS + 2 + FQ_GATHER_WARPS;    // warp 22: per-tile header + voxel-box staging
constexpr int FQ_EPI_B0 = FQ_PREP_WARP + 1;                          // warps 23-26: second half of the epilogue columns
constexpr int FQ_IMMA_WARP = FQ_EPI_B0 + FQ_EPI_WARPS;               // warp 27: issues the interpolation products
constexpr int FQ_THREADS = (FQ_IMMA_WARP + 1) * 32;
constexpr int FQ_UTAB = 512;                     // unit table entries (KP <= 4096)
constexpr int FQ_WGEO_BYTES = 512;               // per-level geometry of the wide path
constexpr int FQ_VBOX_BYTES = 48 * 1024;         // voxel boxes of the tensor-core interpolated levels of one tile
constexpr int FQ_KMAX = 256;                     // largest voxel box (K of the interpolation product)
constexpr int FQ_HDR_BYTES = 1280;                // per-tile header (double buffered)
constexpr int FQ_MAX_ROUNDS = 32, FQ_MAX_CHUNKS = 64;
// TMA staging of the voxel boxes: one tensor copy per (y, z) line of a box = box_x voxels x 64 channels, 128B-swizzled rows
// in box order.  The line length is fixed per tensor map; the smallest of a few pre-encoded ones that covers the box is used
// and becomes the box's x stride.
constexpr int FQ_TMA_LEVELS = 3, FQ_TMA_NBX = 5;
__host__ __device__ constexpr int fq_tma_bx(int i) { return i == 0 ? 3 : (i == 1 ? 4 : (i == 2 ? 5 : (i == 3 ? 6 : 8))); }
// Shared memory is kept SMALL on purpose: what a CTA does not request stays L1 data cache (228 KB - shared memory per SM),
// and the gather lives on L1 hits -- neighbouring (spatially sorted) rows and the 7 stencil points of a row read the same
// voxels.  Staging the corner loads through shared memory (cp.async, one 128-byte slot per thread: latency fully hidden, no
// registers in flight) was measured SLOWER (1.39 vs 1.02 ms at config 2) because its 64 KB shrink L1 to a few KB.  A warp
// that prefetches the next tile's fine-level sectors into L2 (prefetch.global.L2) was also slower (1.08 vs 0.98 ms), and so
// were fp32 halo'd copies of the wide levels (the blend then needs no bf16 unpacking -- 46 instead of 110 instructions per
// unit -- but reads twice the bytes through L1: 1.54 vs 1.08 ms).  The voxel boxes of the tensor-core interpolation are the
// exception: they REPLACE the L1 traffic of the levels they serve.
constexpr int fq_smem(int nb) {
    return 1024 + FQ_NA * FQ_A_BYTES + nb * FQ_B_BYTES + FQ_VBOX_BYTES + FQ_BIAS_BYTES + 2 * FQ_TILE * 16 + 512 + FQ_UTAB * 4 + FQ_WGEO_BYTES +
           2 * FQ_HDR_BYTES + 256 + FQ_TILE * 4 + 2 * FQ_TMA_LEVELS * FQ_TILE * 16;
}
static_assert(fq_smem(FQ_NB) <= 227 * 1024, "fused query shared memory");

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

// Per-tile header, written by the producers, read by every role
struct TileHdr {
    int row0, rows;                // explicit mode: first sorted row and row count of the tile
    int nG, nR;                    // gathered chunks / interpolation rounds of this tile
    int tc_mask;                   // levels interpolated on the tensor cores
    int n_ops;
    int pad_[6];
    int bx0[SVR_MAX_LEVELS], by0[SVR_MAX_LEVELS], bz0[SVR_MAX_LEVELS];   // voxel box origin
    int nx[SVR_MAX_LEVELS], ny[SVR_MAX_LEVELS], nvox[SVR_MAX_LEVELS];    // box extent (x, y) and voxel count
    int voff[SVR_MAX_LEVELS];      // byte offset of the level's box in the box arena
    int nz[SVR_MAX_LEVELS], bxi[SVR_MAX_LEVELS];   // box extent (z); tensor map index of the box's lines (-1: staged by cp.async)
    uint8_t gk[FQ_MAX_CHUNKS];     // K chunks produced by the gather, ascending
    uint8_t rl[FQ_MAX_ROUNDS];     // rounds: level << 4 | stencil point
    uint8_t ops[FQ_MAX_CHUNKS + 2 * FQ_MAX_ROUNDS];   // the tile's schedule: FQ_OP_* << 6 | index into gk / rl
    uint4 rd[FQ_MAX_ROUNDS];       // per round: box address (shared), slab stride, nk16 | nkc << 8 | cpd << 12 | C << 16, first W0 chunk
};
enum : int { FQ_OP_G = 0, FQ_OP_I = 1, FQ_OP_F = 2 };
static_assert(sizeof(TileHdr) <= FQ_HDR_BYTES, "tile header");

struct FqVols {
    const __nv_bfloat16 *v[SVR_MAX_LEVELS];
};

struct FqParams {
    // point source: explicit (points != nullptr) or lattice (dense evaluation)
    const float *points;        // (B*N, 3)
    const int *perm;            // optional row -> point index (sorted processing), may be null
    const StTile *tiles;        // optional tile table (rows cut at sort-cell group boundaries), else 128 consecutive rows
    const int *n_tiles_dev;     // number of entries of `tiles`
    int N;                      // points per scene (explicit mode)
    int64_t total;              // number of rows to process
    // lattice mode
    int lat_scene, sx, sy, sz, x_begin, bx, by, bz;   // bricks per axis over [x_begin, x_end) x sy x sz
    const float *x0;
    FqVols vols;
    FqVols halo;                // optional halo'd copies of the wide levels (null entries: generic path)
    int wide_level0;            // first level taken by the wide path (== P.n_levels: none); every level from here on is wide
    int tc_enable;              // tensor-core interpolation of the wide levels whose voxel box fits
    int use_tma;                // the voxel boxes are staged by TMA tensor copies (else cp.async)
    TensorMap tmap[FQ_TMA_LEVELS][FQ_TMA_NBX];   // per wide level (index l - wide_level0) and x-line length FQ_TMA_BX[i]
    Pyr P;
    const uint8_t *w0_img, *w1_img, *w2_img;   // pre-swizzled chunk images (svr_pack_decoder_images)
    const float *b0, *b1, *b2, *wout, *bout;
    float *out;                 // logits (explicit) or occupancy grid (lattice)
    __nv_bfloat16 *save_h;      // optional (3, total, 256)
    __nv_bfloat16 *save_feat;   // optional (total, KP)
    int apply_sigmoid;
    int nb;                     // weight-ring depth (2..FQ_NB_MAX)
    int trace_block;
    long long *trace;           // debug: per-role (tag, SM clock) records of block 0 (svr_debug_fq_trace), else null
};

// role 0: first gather warp, 1: MMA thread, 2: first epilogue warp, 3: weight loader
__device__ __forceinline__ void fq_trace(const FqParams &p, int role, int &n, int tag) {
    if (p.trace && blockIdx.x == p.trace_block && n < 1024) {
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

// rows of a tile (explicit mode)
__device__ __forceinline__ void tile_rows(const FqParams &p, int64_t tile, int64_t &row0, int &rows) {
    if (!p.points) {
        row0 = tile * FQ_TILE;
        rows = FQ_TILE;
    } else if (p.tiles) {
        const StTile t = p.tiles[tile];
        row0 = t.row0;
        rows = t.rows;
    } else {
        row0 = tile * FQ_TILE;
        const int64_t left = p.total - row0;
        rows = left < FQ_TILE ? (int)left : FQ_TILE;
    }
}

// row of a tile -> point coordinates, scene, and output index (-1 = padding row)
__device__ __forceinline__ void row_point(const FqParams &p, int64_t tile, int64_t row0, int rows, int r, float &px, float &py, float &pz,
                                          int &scene, int64_t &out_idx) {
    if (p.points) {
        if (r >= rows) {
            out_idx = -1;
            scene = 0;
            px = py = pz = 0.f;
            return;
        }
        const int64_t row = row0 + r;
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
    uint8_t *a, *b, *vbox;
    float *bias;                 // [4][256]: b0, b1, b2, wout
    float4 *pts;                 // [2][128] : (px,py,pz, scene as int bits)
    uint64_t *a_full, *a_empty, *b_full, *b_empty, *acc_full, *h_ready, *acc_free;
    uint64_t *hdr_full, *vbox_full, *vbox_free, *i_full, *f_full, *f_free, *hid_done;
    uint32_t *tmem_ptr;
    int *tiles_done;             // tiles completed by the epilogue (guards the double-buffered header / points)
    uint32_t *utab;              // [FQ_UTAB] packed decode_unit results
    WideGeo *wgeo;               // [SVR_MAX_LEVELS]
    TileHdr *hdr;                // [2]
    int *box;                    // spare
    float *dot;                  // [128] partial fc_out dot products of the second epilogue half
    float4 *ctr;                 // [2][FQ_TMA_LEVELS][128] unnormalised sample position (x, y, z) of the stencil centre per wide level
};

__device__ __forceinline__ FqSmem fq_carve(uint8_t *raw, int nb) {
    FqSmem s;
    uint8_t *base = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    s.a = base;
    s.b = s.a + FQ_NA * FQ_A_BYTES;
    s.vbox = s.b + nb * FQ_B_BYTES;
    s.bias = (float *)(s.vbox + FQ_VBOX_BYTES);
    s.pts = (float4 *)((uint8_t *)s.bias + FQ_BIAS_BYTES);
    uint64_t *bars = (uint64_t *)(s.pts + 2 * FQ_TILE);
    s.a_full = bars;
    s.a_empty = s.a_full + FQ_NA;
    s.b_full = s.a_empty + FQ_NA;
    s.b_empty = s.b_full + FQ_NB_MAX;
    s.acc_full = s.b_empty + FQ_NB_MAX;   // [1] accumulator complete (every layer)
    s.h_ready = s.acc_full + 2;       // [1] hidden activations written to TMEM
    s.acc_free = s.h_ready + 1;       // [1] accumulator drained by the last epilogue of a tile
    s.hdr_full = s.acc_free + 1;      // [2] tile header written
    s.vbox_full = s.hdr_full + 2;     // [1] voxel boxes of the tile staged
    s.vbox_free = s.vbox_full + 1;    // [1] interpolation products of the tile complete
    s.i_full = s.vbox_free + 1;       // [2] interpolation accumulator of a round complete (per TMEM buffer)
    s.f_full = s.i_full + 2;          // [2] bf16 features of a round in TMEM
    s.f_free = s.f_full + 2;          // [2] ... consumed by fc_0: the buffer may take the round after next
    s.hid_done = s.f_free + 2;        // [1] hidden-layer products of the tile complete (their TMEM operands are the buffers)
    s.tmem_ptr = (uint32_t *)(s.hid_done + 1);
    s.tiles_done = (int *)(s.tmem_ptr + 1);
    s.utab = (uint32_t *)((uint8_t *)bars + 512);
    s.wgeo = (WideGeo *)((uint8_t *)s.utab + FQ_UTAB * 4);
    s.hdr = (TileHdr *)((uint8_t *)s.wgeo + FQ_WGEO_BYTES);
    s.box = (int *)((uint8_t *)s.hdr + 2 * FQ_HDR_BYTES);
    s.dot = (float *)((uint8_t *)s.box + 256);
    s.ctr = (float4 *)((uint8_t *)s.dot + FQ_TILE * 4);
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


__global__ void __launch_bounds__(FQ_THREADS, 1) fused_query_kernel(const __grid_constant__ FqParams p, int64_t n_tiles_host) {
    extern __shared__ uint8_t smem_raw[];
    const FqSmem s = fq_carve(smem_raw, p.nb);
    const int NB = p.nb;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KC0 = p.P.kp / 64;
    const int64_t n_tiles = p.n_tiles_dev ? (int64_t)*p.n_tiles_dev : n_tiles_host;

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
        mbar_init(s.h_ready, 2 * FQ_EPI_WARPS);
        mbar_init(s.acc_free, 2 * FQ_EPI_WARPS);
        mbar_init(s.hdr_full + 0, 1);
        mbar_init(s.hdr_full + 1, 1);
        mbar_init(s.vbox_full, 1);
        mbar_init(s.vbox_free, 1);
        mbar_init(s.i_full + 0, 1);
        mbar_init(s.i_full + 1, 1);
        mbar_init(s.f_full + 0, 2 * FQ_EPI_WARPS);
        mbar_init(s.f_full + 1, 2 * FQ_EPI_WARPS);
        mbar_init(s.f_free + 0, 1);
        mbar_init(s.f_free + 1, 1);
        mbar_init(s.hid_done, 1);
        *s.tiles_done = 0;
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(s.tmem_ptr, 512);
    for (int i = threadIdx.x; i < FQ_VBOX_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4 *>(s.vbox)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
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
    const uint32_t tm_ibuf = tmem + 256;      // + 128 * (round & 1)

    // number of tiles of this CTA
    int64_t my_tiles = 0;
    if ((int64_t)blockIdx.x < n_tiles) my_tiles = (n_tiles - 1 - blockIdx.x) / gridDim.x + 1;

    if (warp == FQ_PREP_WARP) {
        // ======================= tile prep: points, voxel boxes, schedule of the NEXT tiles; box staging =======================
        int tn = 0;
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t tile = blockIdx.x + it * gridDim.x;
            float4 *pts_s = s.pts + (it & 1) * FQ_TILE;
            TileHdr *hdr = s.hdr + (it & 1);
            // the header / points of tile it-2 (same buffers) must not be in use any more
            if (it >= 2) {
                volatile int *done = s.tiles_done;
                long long t0 = clock64();
                while (*done < (int)it - 1) {
                    if (clock64() - t0 > 4000000000LL) {
                        if (lane == 0) printf("svr_b200: fused query tile prep timed out waiting for the epilogue (block %d)\n", (int)blockIdx.x);
                        __trap();
                    }
                }
            }
            __syncwarp();
            int64_t row0;
            int rows;
            tile_rows(p, tile, row0, rows);
            if (lane == 0) fq_trace(p, 3, tn, 40);
            int smin = 0x7fffffff, smax = -1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = lane + 32 * j;
                float px, py, pz;
                int scene;
                int64_t oi;
                row_point(p, tile, row0, rows, r, px, py, pz, scene, oi);
                pts_s[r] = make_float4(px, py, pz, __int_as_float(oi < 0 ? -1 : scene));
                if (oi >= 0) {
                    smin = min(smin, scene);
                    smax = max(smax, scene);
                }
            }
            smin = __reduce_min_sync(0xffffffffu, smin);
            smax = __reduce_max_sync(0xffffffffu, smax);
            const bool one_scene = smax >= 0 && smin == smax;
            __syncwarp();
            // ---- voxel boxes of the wide levels: the extreme corners of a row come from the -/+ displaced samples (the
            // index arithmetic of stencil_corners), reduced over the warp; lane 0 then writes the schedule
            int tc_mask = 0, voff = 0, nR = 0, nG = 0;
            for (int l = p.wide_level0; l < p.P.n_levels && l - p.wide_level0 < FQ_TMA_LEVELS; ++l) {
                if (!p.tc_enable || !one_scene) break;
                int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-0x7fffffff, -0x7fffffff, -0x7fffffff};
                const int size[3] = {p.P.W[l], p.P.H[l], p.P.D[l]};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 q = pts_s[lane + 32 * j];
                    if (__float_as_int(q.w) < 0) continue;
                    const float q2[3] = {__fmul_rn(2.0f, q.z), __fmul_rn(2.0f, q.y), __fmul_rn(2.0f, q.x)};
                    int rlo[3], rhi[3];
                    bool any = true;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const float sz = (float)size[a];
                        const float f1 = fminf(fmaxf(floorf(unnorm(__fadd_rn(q2[a], -p.P.delta), sz, p.P.align)), -4.0f), sz + 2.0f);
                        const float f2 = fminf(fmaxf(floorf(unnorm(__fadd_rn(q2[a], p.P.delta), sz, p.P.align)), -4.0f), sz + 2.0f);
                        rlo[a] = max((int)fminf(f1, f2), 0);
                        rhi[a] = min((int)fmaxf(f1, f2) + 1, size[a] - 1);
                        any = any && rlo[a] <= rhi[a];
                    }
                    if (any) {
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            lo[a] = min(lo[a], rlo[a]);
                            hi[a] = max(hi[a], rhi[a]);
                        }
                    }
                }
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
                    hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
                }
                int nx = hi[0] - lo[0] + 1;
                const int ny = hi[1] - lo[1] + 1, nz = hi[2] - lo[2] + 1;
                if (lo[0] > hi[0] || nx > FQ_KMAX || ny > FQ_KMAX || nz > FQ_KMAX) continue;
                int bxi = -1;                                   // tensor map of the box's lines (its line length = x stride of the box)
                if (p.use_tma && l - p.wide_level0 < FQ_TMA_LEVELS) {
                    for (int i = FQ_TMA_NBX - 1; i >= 0; --i)
                        if (fq_tma_bx(i) >= nx) bxi = i;
                    if (bxi >= 0) nx = fq_tma_bx(bxi);
                }
                const int nvox = nx * ny * nz, C = p.P.C[l];
                const int bytes = ((nvox + 15) & ~15) * C * 2;
                if (nvox > FQ_KMAX || C > 128 || voff + bytes > FQ_VBOX_BYTES || nR + 7 > FQ_MAX_ROUNDS) continue;
                tc_mask |= 1 << l;
                if (lane == 0) {
                    hdr->bx0[l] = lo[0];
                    hdr->by0[l] = lo[1];
                    hdr->bz0[l] = lo[2];
                    hdr->nx[l] = nx;
                    hdr->ny[l] = ny;
                    hdr->nvox[l] = nvox;
                    hdr->voff[l] = voff;
                    hdr->bxi[l] = bxi;
                    hdr->nz[l] = nz;
                }
                if (lane < 7) {
                    hdr->rl[nR + lane] = (uint8_t)(l << 4 | lane);
                    const int rows16 = (nvox + 15) & ~15, cpd = C >> 6;
                    hdr->rd[nR + lane] = make_uint4(smem_u32(s.vbox + voff), (uint32_t)rows16 * 128u,
                                                    (uint32_t)(rows16 >> 4) | (uint32_t)((nvox + 63) >> 6) << 8 | (uint32_t)cpd << 12 | (uint32_t)C << 16,
                                                    (uint32_t)(s.wgeo[l].chunk0 + lane * cpd));
                }
                nR += 7;
                voff += bytes;
            }
            if (lane == 0) {
                int kc = 0;
                for (int l = p.wide_level0; l < p.P.n_levels; ++l) {
                    const int c0 = s.wgeo[l].chunk0, c1 = c0 + 7 * s.wgeo[l].cpd;
                    for (; kc < c0; ++kc) hdr->gk[nG++] = (uint8_t)kc;
                    if ((tc_mask >> l) & 1) kc = c1;
                    for (; kc < c1; ++kc) hdr->gk[nG++] = (uint8_t)kc;
                }
                for (; kc < KC0; ++kc) hdr->gk[nG++] = (uint8_t)kc;
                // schedule: I(0), I(1), F(0), I(2), F(1), ..., F(nR-1) with the gathered chunks spread evenly in between
                int n = 0, rI = 0;
                const int nS = nG > 0 ? nG : 1;
                for (int i = 0; i < nS; ++i) {
                    if (i < nG) hdr->ops[n++] = (uint8_t)(FQ_OP_G << 6 | i);
                    const int target = ((i + 1) * nR) / nS;
                    for (; rI < target; ++rI) {
                        hdr->ops[n++] = (uint8_t)(FQ_OP_I << 6 | rI);
                        if (rI >= 1) hdr->ops[n++] = (uint8_t)(FQ_OP_F << 6 | (rI - 1));
                    }
                }
                if (nR > 0) hdr->ops[n++] = (uint8_t)(FQ_OP_F << 6 | (nR - 1));
                hdr->n_ops = n;
                hdr->row0 = (int)row0;
                hdr->rows = rows;
                hdr->tc_mask = tc_mask;
                hdr->nG = nG;
                hdr->nR = nR;
                hdr->pad_[0] = smin;
            }
            // unnormalised sample positions of the stencil centre on the tensor-core interpolated levels: the producers need
            // them in every round, and the index arithmetic was half of what a round cost them
            for (int l = p.wide_level0; l < p.P.n_levels && l - p.wide_level0 < FQ_TMA_LEVELS; ++l) {
                if (!((tc_mask >> l) & 1)) continue;
                const float fw = (float)p.P.W[l], fh = (float)p.P.H[l], fd = (float)p.P.D[l];
                float4 *dst = s.ctr + ((it & 1) * FQ_TMA_LEVELS + (l - p.wide_level0)) * FQ_TILE;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 q = pts_s[lane + 32 * j];
                    dst[lane + 32 * j] = make_float4(unnorm(__fmul_rn(2.0f, q.z), fw, p.P.align), unnorm(__fmul_rn(2.0f, q.y), fh, p.P.align),
                                                     unnorm(__fmul_rn(2.0f, q.x), fd, p.P.align), 0.f);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(s.hdr_full + (it & 1));
            if (lane == 0) fq_trace(p, 3, tn, 50);
            // ---- stage the voxel boxes of this tile (cp.async, zero rows up to a multiple of 16)
            if (it > 0) mbar_wait(s.vbox_free, (uint32_t)(it - 1) & 1);     // the previous tile's interpolation products are done
            uint32_t tma_bytes = 0;
            if (tc_mask) {
                const int scene = smin;
                for (int l = p.wide_level0; l < p.P.n_levels; ++l) {
                    if (!((tc_mask >> l) & 1)) continue;
                    const int C = p.P.C[l], W = p.P.W[l], H = p.P.H[l], D = p.P.D[l];
                    const int ncg = C >> 3, nvox = hdr->nvox[l], rows16 = (nvox + 15) & ~15;
                    const int nx = hdr->nx[l], ny = hdr->ny[l], bx0 = hdr->bx0[l], by0 = hdr->by0[l], bz0 = hdr->bz0[l];
                    const uint32_t dst0 = smem_u32(s.vbox + hdr->voff[l]);
                    const int bxi = hdr->bxi[l];
                    if (bxi >= 0) {
                        // one tensor copy per (y, z) line and 64-channel slab: nx voxels x 128 B, rows in box order; voxels beyond
                        // the volume's x extent are zero-filled (their weights are zero too)
                        const int nz = hdr->nz[l], lines = ny * nz, slabs = C >> 6;
                        tma_bytes += (uint32_t)(lines * slabs * nx * 128);
                        const TensorMap *tm = &p.tmap[l - p.wide_level0][bxi];
                        for (int i = lane; i < lines * slabs; i += 32) {
                            const int sl = i / lines, ln = i - sl * lines, lz = ln / ny, ly = ln - lz * ny;
                            tma_load_5d(dst0 + sl * (rows16 * 128) + ln * nx * 128, tm, sl * 64, bx0, by0 + ly, bz0 + lz, scene, s.vbox_full);
                        }
                        continue;
                    }
                    const __nv_bfloat16 *vol = p.vols.v[l] + (int64_t)scene * D * H * W * C;
                    const float inv_ncg = 1.0f / (float)ncg, inv_nx = 1.0f / (float)nx, inv_ny = 1.0f / (float)ny;
                    for (int i = lane; i < rows16 * ncg; i += 32) {
                        // small-integer divisions through exact float reciprocals (operands < 2^13)
                        const int vox = (int)(((float)i + 0.5f) * inv_ncg), cg = i - vox * ncg;
                        const bool ok = vox < nvox;
                        const int t2 = (int)(((float)vox + 0.5f) * inv_nx), lx = vox - t2 * nx;
                        const int lz = (int)(((float)t2 + 0.5f) * inv_ny), ly = t2 - lz * ny;
                        const int64_t off = ok ? ((((int64_t)(bz0 + lz) * H + (by0 + ly)) * W + (bx0 + lx)) * C + cg * 8) : 0;
                        cp_async16(dst0 + (cg >> 3) * (rows16 * 128) + swz128(vox, cg & 7), vol + off, ok);
                    }
                }
            }
            cp_async_commit();
            cp_async_wait<0>();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                if (tma_bytes) mbar_arrive_expect_tx(s.vbox_full, tma_bytes);
                else mbar_arrive(s.vbox_full);
            }
            if (lane == 0) fq_trace(p, 3, tn, 51);
        }
    } else if (warp >= 6 && warp < FQ_PREP_WARP) {
        // ======================= producers =======================
        const int gt = threadIdx.x - 6 * 32;          // 0..511
        const int wg = warp - 6;                      // 0..15
        const int unit_in_chunk = lane & 7;
        uint32_t gc = 0;                              // global A-chunk counter
        int tn = 0;
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t tile = blockIdx.x + it * gridDim.x;
            float4 *pts_s = s.pts + (it & 1) * FQ_TILE;
            TileHdr *hdr = s.hdr + (it & 1);
            mbar_wait(s.hdr_full + (it & 1), (uint32_t)(it >> 1) & 1);      // header, points (and soon the voxel boxes) by the prep warp
            if (gt == 0) fq_trace(p, 0, tn, 50);
            const int64_t row0 = hdr->row0;
            const int rows = hdr->rows;

            // ---- one gathered K chunk: thread -> (rows slot, slot + 64) x unit (lane & 7); both rows of a thread are in
            // flight together (16 independent corner loads), a warp whose second (or first) row quad lies beyond the tile's
            // rows skips it (warp-uniform)
            const int slot = gt >> 3;
            const bool have0 = 4 * wg < rows, have1 = 64 + 4 * wg < rows;
            auto produce_G = [&](int kc) {
                const int st = gc % FQ_NA;
                mbar_wait(s.a_empty + st, ((gc / FQ_NA) & 1) ^ 1);
                if (gt == 0) fq_trace(p, 0, tn, 100 + kc);
                uint8_t *a_st = s.a + st * FQ_A_BYTES;
                if (have0) {
                    int wl = -1;
                    for (int l = p.wide_level0; l < p.P.n_levels; ++l)
                        if (kc >= s.wgeo[l].chunk0 && kc < s.wgeo[l].chunk0 + 7 * s.wgeo[l].cpd) wl = l;
                    if (wl < 0) {
                        // generic chunk (levels without a halo'd copy, alignment padding, the tail)
                        const int u = kc * 8 + unit_in_chunk;
                        UnitCtx uc;
                        make_unit_ctx_packed(p.P, s.utab[u], p.vols.v, uc);
                        auto one_row = [&](int r) {
                            const float4 pq = pts_s[r];
                            const int scene = __float_as_int(pq.w);
                            uint4 val = make_uint4(0, 0, 0, 0);
                            if (kc == 0) {
                                // Unit 0 is the level-0 unit: 7 stencil samples of the fp32 input grid (56 scalar loads): lane j
                                // of the point's 8-lane group takes sample j and lane 0 collects.
                                float smp = 0.f;
                                if (scene >= 0 && unit_in_chunk < 7) {
                                    const float *x0b = p.x0 + (int64_t)scene * p.P.D[0] * p.P.H[0] * p.P.W[0];
                                    smp = level0_sample(p.P, unit_in_chunk, pq.x, pq.y, pq.z, x0b);
                                }
                                float v8[8];
#pragma unroll
                                for (int dd = 0; dd < 8; ++dd) v8[dd] = __shfl_sync(0xffffffffu, smp, (lane & 24) + dd);
                                if (unit_in_chunk == 0)
                                    val = float8_to_bf16(v8);
                                else if (uc.real && uc.level > 0 && scene >= 0)
                                    val = gather_unit_bf(uc, p.P.align, pq.x, pq.y, pq.z, scene);
                            } else if (uc.real && scene >= 0) {
                                val = gather_unit_bf(uc, p.P.align, pq.x, pq.y, pq.z, scene);
                            }
                            *reinterpret_cast<uint4 *>(a_st + swz128(r, unit_in_chunk)) = val;
                            if (p.save_feat && r < rows)
                                *reinterpret_cast<uint4 *>(p.save_feat + (row0 + r) * p.P.kp + (int64_t)u * 8) = val;
                        };
                        if (have1) {
#pragma unroll 2
                            for (int j = 0; j < 2; ++j) one_row(slot + 64 * j);
                        } else {
                            one_row(slot);
                        }
                    } else {
                        // wide chunk (level wl, stencil point d, 64-channel slice h) sampled from the halo'd copy
                        const WideGeo G = s.wgeo[wl];
                        const int rel = kc - G.chunk0, d = rel / G.cpd, h = rel - d * G.cpd;
                        const float sgn = (d & 1) ? -p.P.delta : p.P.delta;
                        const float dx = (d == 1 || d == 2) ? sgn : 0.f, dy = (d == 3 || d == 4) ? sgn : 0.f, dz = (d == 5 || d == 6) ? sgn : 0.f;
                        const int goff = (h * 8 + unit_in_chunk) * 8;          // first channel of this thread's group
                        auto blend = [&](const __nv_bfloat16 *ptr, const float (&w)[8]) -> uint4 {
                            switch (G.spec) {
                                case 1: return wide_blend<WS1_C, WS1_Y, WS1_Z>(ptr + goff, w, 0, 0, 0);
                                case 2: return wide_blend<WS2_C, WS2_Y, WS2_Z>(ptr + goff, w, 0, 0, 0);
                                case 3: return wide_blend<WS3_C, WS3_Y, WS3_Z>(ptr + goff, w, 0, 0, 0);
                                default: return wide_blend<0, 0, 0>(ptr + goff, w, G.C, G.sy, G.sz);
                            }
                        };
                        auto put = [&](int r, const uint4 val) {
                            *reinterpret_cast<uint4 *>(a_st + swz128(r, unit_in_chunk)) = val;
                            if (p.save_feat && r < rows)
                                *reinterpret_cast<uint4 *>(p.save_feat + (row0 + r) * p.P.kp + (int64_t)(kc * 8 + unit_in_chunk) * 8) = val;
                        };
                        if (have1) {
                            const __nv_bfloat16 *ptr[2];
                            float w[2][8];
#pragma unroll
                            for (int j = 0; j < 2; ++j) wide_desc(G, p.P.align, dx, dy, dz, pts_s[slot + 64 * j], ptr[j], w[j]);
#pragma unroll
                            for (int j = 0; j < 2; ++j) put(slot + 64 * j, blend(ptr[j], w[j]));
                        } else {
                            const __nv_bfloat16 *ptr;
                            float w[8];
                            wide_desc(G, p.P.align, dx, dy, dz, pts_s[slot], ptr, w);
                            put(slot, blend(ptr, w));
                        }
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(s.a_full + st);
                if (gt == 0) fq_trace(p, 0, tn, 200 + kc);
                ++gc;
            };
            // ---- the interpolation weights of one round: (128 rows x box) K-major, zeros except 8 weights per row; this
            // warp owns rows 8*wg .. 8*wg+7, four lanes per row: lane pair (bb, e) writes the two x-neighbours of corner
            // (y0 + bb, z0 + e).  A row's lanes zero their quarter of the row first (same warp: __syncwarp orders it).
            auto produce_A = [&](int code) {
                const int l = code >> 4, d = code & 15;
                const int r = wg * 8 + (lane >> 2), qd = lane & 3, bb = qd & 1, e = qd >> 1;
                const float4 pq = pts_s[r];
                const bool rv = __float_as_int(pq.w) >= 0;
                // the corners of stencil_corners(p.P, l, d, ...): centre positions from the prep warp's table, only the displaced
                // axis is re-evaluated (same operations, same bits)
                Corners c;
                {
                    const WideGeo &G = s.wgeo[l];
                    const float4 ct = s.ctr[((it & 1) * FQ_TMA_LEVELS + (l - p.wide_level0)) * FQ_TILE + r];
                    float ix = ct.x, iy = ct.y, iz = ct.z;
                    if (d != 0) {
                        const float sgn = (d & 1) ? -p.P.delta : p.P.delta;
                        if (d <= 2) ix = unnorm(__fadd_rn(__fmul_rn(2.0f, pq.z), sgn), G.fw, p.P.align);
                        else if (d <= 4) iy = unnorm(__fadd_rn(__fmul_rn(2.0f, pq.y), sgn), G.fh, p.P.align);
                        else iz = unnorm(__fadd_rn(__fmul_rn(2.0f, pq.x), sgn), G.fd, p.P.align);
                    }
                    const float fx = fminf(fmaxf(floorf(ix), -4.0f), G.fw + 2.0f), fy = fminf(fmaxf(floorf(iy), -4.0f), G.fh + 2.0f),
                                fz = fminf(fmaxf(floorf(iz), -4.0f), G.fd + 2.0f);
                    c.x0 = (int)fx;
                    c.y0 = (int)fy;
                    c.z0 = (int)fz;
                    c.wx[1] = ix - fx;
                    c.wx[0] = (fx + 1.0f) - ix;
                    c.wy[1] = iy - fy;
                    c.wy[0] = (fy + 1.0f) - iy;
                    c.wz[1] = iz - fz;
                    c.wz[0] = (fz + 1.0f) - iz;
                }
                const int nvox = hdr->nvox[l], nkc = (nvox + 63) >> 6;
                const int y = c.y0 + bb, z = c.z0 + e;
                const bool yz_ok = rv && y >= 0 && y < p.P.H[l] && z >= 0 && z < p.P.D[l];
                const int lrow = ((z - hdr->bz0[l]) * hdr->ny[l] + (y - hdr->by0[l])) * hdr->nx[l] - hdr->bx0[l];
                const float wyz = (bb ? c.wy[1] : c.wy[0]) * (e ? c.wz[1] : c.wz[0]);
                const int W = p.P.W[l];
                if (gt == 0) fq_trace(p, 0, tn, 1000 + code);
                // up to FQ_NA chunks at a time: fill them, ONE proxy fence, then hand them to the MMA thread
                for (int ck0 = 0; ck0 < nkc; ck0 += FQ_NA) {
                    const int n = min(FQ_NA, nkc - ck0);
                    for (int j = 0; j < n; ++j) {
                        const uint32_t g = gc + j;
                        const int st = g % FQ_NA;
                        mbar_wait(s.a_empty + st, ((g / FQ_NA) & 1) ^ 1);
                        uint8_t *a_st = s.a + st * FQ_A_BYTES;
                        *reinterpret_cast<uint4 *>(a_st + swz128(r, 2 * qd)) = make_uint4(0, 0, 0, 0);
                        *reinterpret_cast<uint4 *>(a_st + swz128(r, 2 * qd + 1)) = make_uint4(0, 0, 0, 0);
                    }
                    __syncwarp();
                    if (yz_ok) {
#pragma unroll
                        for (int aa = 0; aa < 2; ++aa) {
                            const int x = c.x0 + aa, lid = lrow + x, j = (lid >> 6) - ck0;
                            if (x < 0 || x >= W || lid < 0 || lid >= nvox || j < 0 || j >= n) continue;
                            uint8_t *a_st = s.a + ((gc + j) % FQ_NA) * FQ_A_BYTES;
                            *reinterpret_cast<__nv_bfloat16 *>(a_st + swz128(r, (lid & 63) >> 3) + (lid & 7) * 2) = __float2bfloat16(c.wx[aa] * wyz);
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0)
                        for (int j = 0; j < n; ++j) mbar_arrive(s.a_full + (gc + j) % FQ_NA);
                    gc += n;
                }
            };
            const int n_ops = hdr->n_ops;
            for (int i = 0; i < n_ops; ++i) {
                const int op = hdr->ops[i];
                if ((op >> 6) == FQ_OP_G) produce_G(hdr->gk[op & 63]);
                if ((op >> 6) == FQ_OP_I) produce_A(hdr->rl[op & 63]);
            }
        }
    } else if (warp == 4) {
        // ======================= weight loader =======================
        if (lane == 0) {
            uint32_t wc = 0;
            auto load = [&](const uint8_t *src) {
                const int st = wc % NB;
                mbar_wait(s.b_empty + st, ((wc / NB) & 1) ^ 1);
                mbar_arrive_expect_tx(s.b_full + st, FQ_B_BYTES);
                bulk_g2s(smem_u32(s.b + st * FQ_B_BYTES), src, FQ_B_BYTES, s.b_full + st);
                ++wc;
            };
            for (int64_t it = 0; it < my_tiles; ++it) {
                const TileHdr *hdr = s.hdr + (it & 1);
                mbar_wait(s.hdr_full + (it & 1), (uint32_t)(it >> 1) & 1);
                const int n_ops = hdr->n_ops;
                for (int i = 0; i < n_ops; ++i) {
                    const int op = hdr->ops[i];
                    if ((op >> 6) == FQ_OP_G) load(p.w0_img + (size_t)hdr->gk[op & 63] * FQ_B_BYTES);
                    if ((op >> 6) == FQ_OP_F) {
                        const int code = hdr->rl[op & 63], l = code >> 4, d = code & 15;
                        const int cpd = s.wgeo[l].cpd, c0 = s.wgeo[l].chunk0 + d * cpd;
                        for (int h = 0; h < cpd; ++h) load(p.w0_img + (size_t)(c0 + h) * FQ_B_BYTES);
                    }
                }
                for (int c = 0; c < 4; ++c) load(p.w1_img + (size_t)c * FQ_B_BYTES);
                for (int c = 0; c < 4; ++c) load(p.w2_img + (size_t)c * FQ_B_BYTES);
            }
        }
    } else if (warp == FQ_IMMA_WARP) {
        // ======================= interpolation MMA issue =======================
        // I(r): features of round r = weights (A ring, K-major) x voxel box (MN-major) -> TMEM buffer r & 1.  A second issuing
        // thread: the ~50 small steps of a tile were bound by ONE thread's issue rate.
        if (lane == 0 && my_tiles > 0) {
            uint32_t gc = 0, rI = 0;
            int tn = 0;
            for (int64_t it = 0; it < my_tiles; ++it) {
                const TileHdr *hdr = s.hdr + (it & 1);
                mbar_wait(s.hdr_full + (it & 1), (uint32_t)(it >> 1) & 1);
                mbar_wait(s.vbox_full, (uint32_t)it & 1);
                if (it > 0) mbar_wait(s.hid_done, (uint32_t)(it - 1) & 1);   // the buffers were H0 / H1 of the previous tile
                tc_fence_after();
                const int n_ops = hdr->n_ops;
                for (int i = 0; i < n_ops; ++i) {
                    const int op = hdr->ops[i], kind = op >> 6, idx = op & 63;
                    if (kind == FQ_OP_G) {
                        ++gc;
                    } else if (kind == FQ_OP_I) {
                        const uint4 rd = hdr->rd[idx];
                        const int nk16 = rd.z & 255, nkc = (rd.z >> 8) & 15, C = rd.z >> 16;
                        const uint32_t idesc_i = make_idesc_bf16(FQ_TILE, C, 0, 1);
                        const uint32_t tm_i = tm_ibuf + 128u * (rI & 1);
                        if (rI >= 2) mbar_wait(s.f_free + (rI & 1), ((rI - 2) >> 1) & 1);   // fc_0 has consumed the round before last
                        tc_fence_after();
                        for (int ck = 0; ck < nkc; ++ck, ++gc) {
                            const int sa = gc % FQ_NA;
                            mbar_wait(s.a_full + sa, (gc / FQ_NA) & 1);
                            tc_fence_after();
                            const uint32_t a_s = smem_u32(s.a + sa * FQ_A_BYTES);
                            const int nk = min(4, nk16 - 4 * ck);
                            for (int kk = 0; kk < nk; ++kk)
                                umma_bf16(tm_i, make_smem_desc(a_s + kk * 32, 16, 1024, kSwizzle128B),
                                          make_smem_desc(rd.x + (uint32_t)(ck * 4 + kk) * 2048, rd.y, 1024, kSwizzle128B), idesc_i, (ck | kk) != 0);
                            umma_commit(s.a_empty + sa);
                        }
                        umma_commit(s.i_full + (rI & 1));
                        ++rI;
                    }
                }
                umma_commit(s.vbox_free);
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ======================= MMA issue =======================
        if (lane == 0 && my_tiles > 0) {
            const uint32_t idesc = make_idesc_bf16(FQ_TILE, FQ_HID, 0, 0);
            uint32_t gc = 0, wc = 0, hr = 0, rF = 0;
            int tn = 0;
            auto wait_b = [&]() {
                const int st = wc % NB;
                mbar_wait(s.b_full + st, (wc / NB) & 1);
                return st;
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
            for (int64_t it = 0; it < my_tiles; ++it) {
                const TileHdr *hdr = s.hdr + (it & 1);
                mbar_wait(s.hdr_full + (it & 1), (uint32_t)(it >> 1) & 1);
                const int n_ops = hdr->n_ops;
                bool acc_started = false;
                auto first_acc = [&]() {     // before the first MMA into the fc accumulator: the previous tile's logits epilogue drained it
                    if (!acc_started && it > 0) {
                        mbar_wait(s.acc_free, (uint32_t)(it - 1) & 1);
                        tc_fence_after();
                    }
                };
                for (int i = 0; i < n_ops; ++i) {
                    const int op = hdr->ops[i], kind = op >> 6, idx = op & 63;
                    if (kind == FQ_OP_G) {
                        // ---- G: gathered chunk x W0 chunk
                        const int sa = gc % FQ_NA;
                        mbar_wait(s.a_full + sa, (gc / FQ_NA) & 1);
                        fq_trace(p, 1, tn, 300 + hdr->gk[idx]);
                        const int sb = wait_b();
                        first_acc();
                        tc_fence_after();
                        const uint32_t a_s = smem_u32(s.a + sa * FQ_A_BYTES), b_s = smem_u32(s.b + sb * FQ_B_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(acc0, make_smem_desc(a_s + k * 32, 16, 1024, kSwizzle128B),
                                      make_smem_desc(b_s + k * 32, 16, 1024, kSwizzle128B), idesc, acc_started || k != 0);
                        umma_commit(s.a_empty + sa);
                        umma_commit(s.b_empty + sb);
                        acc_started = true;
                        ++gc;
                        ++wc;
                    } else if (kind == FQ_OP_I) {
                        gc += (hdr->rd[idx].z >> 8) & 15;      // the interpolation products are issued by warp FQ_IMMA_WARP
                    } else {
                        // ---- F: bf16 features of a round (TMEM, rounded in place by the epilogue) x W0 chunks of (level, stencil point)
                        const int cpd = (hdr->rd[idx].z >> 12) & 15;
                        const uint32_t tm_f = tm_ibuf + 128u * (rF & 1);
                        mbar_wait(s.f_full + (rF & 1), (rF >> 1) & 1);
                        fq_trace(p, 1, tn, 700 + idx);
                        tc_fence_after();
                        for (int h = 0; h < cpd; ++h) {
                            const int sb = wait_b();
                            first_acc();
                            tc_fence_after();
                            const uint32_t b_s = smem_u32(s.b + sb * FQ_B_BYTES);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16_ts(acc0, tm_f + (uint32_t)(h * 4 + k) * 8, make_smem_desc(b_s + k * 32, 16, 1024, kSwizzle128B), idesc,
                                             acc_started || k != 0);
                            umma_commit(s.b_empty + sb);
                            acc_started = true;
                            ++wc;
                        }
                        umma_commit(s.f_free + (rF & 1));
                        ++rF;
                    }
                }
                umma_commit(s.acc_full);
                issue_hidden(tm_h0);
                issue_hidden(tm_h1);
                umma_commit(s.hid_done);
            }
        }
        __syncwarp();
    } else {
        // ======================= epilogue =======================
        // Two warps per TMEM lane quarter (a warp reaches the lanes 32 * (warp % 4) ..): warps 0-3 take the first half of
        // the columns of every pass, warps 23-26 the second half.
        const int half = warp >= FQ_EPI_B0 ? 1 : 0, quarter = warp & 3;
        const int r = quarter * 32 + lane;         // row in tile == TMEM lane
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        uint32_t n_acc = 0;                        // completions consumed of acc_full
        uint32_t rE = 0;                           // interpolation rounds consumed
        int tn = 0;
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t tile = blockIdx.x + it * gridDim.x;
            const TileHdr *hdr = s.hdr + (it & 1);
            mbar_wait(s.hdr_full + (it & 1), (uint32_t)(it >> 1) & 1);
            const int nR = hdr->nR;
            int64_t row0;
            int rows;
            tile_rows(p, tile, row0, rows);
            float px, py, pz;
            int scene;
            int64_t out_idx;
            row_point(p, tile, row0, rows, r, px, py, pz, scene, out_idx);
            const int64_t row = row0 + r;
            const bool row_ok = p.points ? r < rows : out_idx >= 0;
            // ---- interpolation rounds: fp32 features (TMEM) -> bf16 pairs (TMEM, A operand of fc_0) [+ saved features]
#pragma unroll 1
            for (int rr = 0; rr < nR; ++rr) {
                const int code = hdr->rl[rr], l = code >> 4, d = code & 15;
                const int C = s.wgeo[l].C, cpd = s.wgeo[l].cpd, chunk = s.wgeo[l].chunk0 + d * cpd;
                const uint32_t tm_i = tm_ibuf + 128u * (rE & 1) + lane_off;
                mbar_wait(s.i_full + (rE & 1), (rE >> 1) & 1);
                tc_fence_after();
                if (threadIdx.x == 0) fq_trace(p, 2, tn, 900 + rr);
                __nv_bfloat16 *frow = (p.save_feat && row_ok && p.points) ? p.save_feat + row * p.P.kp + (int64_t)chunk * 64 : nullptr;
                // In place: the packed pairs of columns c0 .. c0+31 go to columns c0/2 .. c0/2+15.  This half takes the
                // 32-column groups half, half + 2: its first store lands on columns the OTHER half reads in its first
                // step (and vice versa), so the two warps of a lane quarter meet once between their first load and store.
#pragma unroll 1
                for (int c0 = 32 * half; c0 < C; c0 += 64) {
                    uint32_t v[32];
                    tmem_ld32(tm_i + c0, v);
                    tmem_ld_wait();
                    if (c0 < 64) named_bar_sync(2 + quarter, 64);
                    uint32_t packed[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                        packed[j] = *reinterpret_cast<uint32_t *>(&h);
                    }
                    tmem_st16(tm_i + (c0 >> 1), packed);
                    if (frow) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<uint4 *>(frow + c0 + q * 8) = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(s.f_full + (rE & 1));
                ++rE;
            }
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
                for (int c0 = half * (FQ_HID / 2); c0 < (half + 1) * (FQ_HID / 2); c0 += 32) {
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
            // fc_out: the second half hands its partial dot product to the first
            if (half) s.dot[r] = dot;
            named_bar_sync(2 + quarter, 64);
            if (!half) dot += s.dot[r];
            if (row_ok && !half) {
                float logit = dot + __ldg(p.bout);
                if (p.apply_sigmoid) logit = 1.0f / (1.0f + __expf(-logit));
                p.out[out_idx] = logit;
            }
            if (threadIdx.x == 0) {
                __threadfence_block();
                *(volatile int *)s.tiles_done = (int)it + 1;
            }
        }
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

static int fq_fill(FqParams &p, const float *x0, const uint16_t *const *vols_host, const uint16_t *const *halo_host,
                   const svr_pyramid *pyr_host, const svr_decoder_weights *w) {
    if (int rc = make_pyr(p.P, pyr_host)) return rc;
    SVR_REQUIRE(w && x0 && vols_host, "fused query: null pointer");
    SVR_REQUIRE(w->h0 == FQ_HID && w->h1 == FQ_HID && w->h2 == FQ_HID, "fused query supports hidden size 256 only (got %d/%d/%d)",
                w->h0, w->h1, w->h2);
    SVR_REQUIRE(w->w0p && w->w1 && w->w2 && w->b0 && w->b1 && w->b2 && w->wout && w->bout, "fused query: null weight pointer");
    SVR_REQUIRE(p.P.kp / 8 <= FQ_UTAB, "fused query: feature row of %d columns exceeds the unit table (%d columns)", p.P.kp, FQ_UTAB * 8);
    SVR_REQUIRE(p.P.kp / 64 <= FQ_MAX_CHUNKS, "fused query: feature row of %d columns exceeds %d K chunks", p.P.kp, FQ_MAX_CHUNKS);
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
            SVR_REQUIRE(((uintptr_t)p.vols.v[l] & 15) == 0, "fused query: volume of level %d is not 16-byte aligned", l);
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
static int g_fq_trace_block = 0;
static int g_fq_interp = 1;     // 0: the coarse levels are gathered on the CUDA cores like the others (ablation)
static int g_fq_tma = 1;        // 0: voxel boxes staged by cp.async instead of TMA tensor copies (ablation)
void set_interp(int on, int tma) {
    g_fq_interp = on;
    g_fq_tma = tma;
}

static int fq_launch(FqParams &p, int64_t n_tiles, int n_scenes, cudaStream_t st) {
    static DeviceOnce once;
    int dev;
    if (once.needed(dev)) {
        SVR_CUDA(cudaFuncSetAttribute(fused_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fq_smem(FQ_NB)));
        once.done(dev);
    }
    if (n_tiles <= 0) return 0;
    int grid = sm_count();
    if (n_tiles < grid) grid = (int)n_tiles;
    p.trace = g_fq_trace;
    p.trace_block = g_fq_trace_block;
    p.nb = FQ_NB;
    p.tc_enable = (g_fq_interp && p.wide_level0 + 4 >= p.P.n_levels) ? 1 : 0;    // the box pass covers up to four wide levels
    p.use_tma = 0;
    if (p.tc_enable && g_fq_tma) {
        p.use_tma = 1;
        for (int l = p.wide_level0; l < p.P.n_levels && l - p.wide_level0 < FQ_TMA_LEVELS && p.use_tma; ++l)
            for (int i = 0; i < FQ_TMA_NBX && p.use_tma; ++i)
                if (make_tmap_vol_bf16(&p.tmap[l - p.wide_level0][i], p.vols.v[l], n_scenes, p.P.D[l], p.P.H[l], p.P.W[l], p.P.C[l], fq_tma_bx(i)))
                    p.use_tma = 0;      // (the cp.async path needs no descriptors)
    }
    fused_query_kernel<<<grid, FQ_THREADS, fq_smem(FQ_NB), st>>>(p, n_tiles);
    SVR_LAUNCH_CHECK();
    return 0;
}

void set_trace(long long *buf, int block) {
    g_fq_trace = buf;
    g_fq_trace_block = block;
}

int query_fwd(const float *points, const int *perm, const int *cell_start, int B, int N, const float *x0,
                        const uint16_t *const *vols_host, const uint16_t *const *halo_vols_host, const svr_pyramid *pyr_host,
                        const svr_decoder_weights *w_host, float *logits, uint16_t *save_h, uint16_t *save_feat, int apply_sigmoid,
                        void *stream) {
    FqParams p{};
    if (int rc = fq_fill(p, x0, vols_host, halo_vols_host, pyr_host, w_host)) return rc;
    SVR_REQUIRE(points && logits, "query_fwd_fused: null pointer");
    SVR_REQUIRE(!cell_start || perm, "query_fwd_fused: cell_start needs the order it belongs to");
    p.points = points;
    p.perm = perm;
    p.N = N;
    p.total = (int64_t)B * N;
    p.out = logits;
    p.save_h = (__nv_bfloat16 *)save_h;
    p.save_feat = (__nv_bfloat16 *)save_feat;
    p.apply_sigmoid = apply_sigmoid;
    if (p.total <= 0) return 0;
    cudaStream_t st = as_stream(stream);
    int64_t n_tiles = ceil_div<int64_t>(p.total, FQ_TILE);
    void *scratch = nullptr;
    if (cell_start && g_fq_interp && p.wide_level0 < p.P.n_levels) {
        // rows cut at the boundaries of sort-cell groups: a tile's samples stay inside a small voxel box on the coarse levels
        const int n_cells = B * svr_sort_cells_per_scene();
        SVR_REQUIRE(n_cells % ST_SUPER == 0, "query_fwd_fused: the number of sort cells must be a multiple of %d", ST_SUPER);
        const int n_groups = n_cells / ST_SUPER;
        n_tiles += n_groups;                                   // sum of ceil(n_k / 128) <= total / 128 + groups
        if (int rc = scratch_alloc(&scratch, (size_t)n_tiles * sizeof(StTile) + 16, st)) return rc;
        if (int rc = launch_st_tiles(cell_start, n_groups, (StTile *)((uint8_t *)scratch + 16), (int *)scratch, st)) return rc;
        p.tiles = (const StTile *)((uint8_t *)scratch + 16);
        p.n_tiles_dev = (const int *)scratch;
    }
    const int rc = fq_launch(p, n_tiles, B, st);
    if (scratch) SVR_CUDA(cudaFreeAsync(scratch, st));
    return rc;
}

int dense_eval(int scene, int B, const float *x0, const uint16_t *const *vols_host, const uint16_t *const *halo_vols_host,
                   const svr_pyramid *pyr_host, const svr_decoder_weights *w_host, int sx, int sy, int sz, int x_begin, int x_end,
                   float *out, void *stream) {
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
    return fq_launch(p, (int64_t)p.bx * p.by * p.bz, B, as_stream(stream));
}

}  // namespace fqb
}  // namespace svr
