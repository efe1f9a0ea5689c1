// Device-side description of the sampled feature pyramid and the trilinear stencil math shared by
// the gather / scatter kernels and the fused query kernel.
//
// Feature axis layout ("K' order").  The reference concatenates levels on channels and flattens
// (C,7) -> k = c*7 + d (model/ifnet.py:43-45,197).  The kernels use 16-byte UNITS of 8 bf16
// channels so that one unit == one contiguous channel-last load per corner:
//   unit 0            : level 0 (1 channel): the 7 stencil samples d=0..6, then one zero
//   units of level l  : for d in 0..6, for g in 0..C_l/8-1 : channels g*8..g*8+7 of stencil point d
//   padding units     : zeros (a) before every level whose channel count is a multiple of 64, up to the
//                       next multiple of 8 units, so that each (level, stencil point) block of such a level
//                       is a whole number of 64-wide K chunks (the tensor-core gather produces whole chunks),
//                       and (b) at the end up to KP = 64 * ceil(units/8)
// fc_0's weight is permuted to the same order once per weight update (svr_pack_w0).
#pragma once
#include "common.cuh"

namespace svr {

struct Pyr {
    int n_levels;
    int C[SVR_MAX_LEVELS], D[SVR_MAX_LEVELS], H[SVR_MAX_LEVELS], W[SVR_MAX_LEVELS];
    int upd[SVR_MAX_LEVELS];        // units per stencil point (C/8), level >= 1
    int ubase[SVR_MAX_LEVELS + 1];  // first unit of each level; ubase[n_levels] = number of real units
    int coff[SVR_MAX_LEVELS];       // channel offset in the reference's concat order
    int n_units, kp, ctot;
    int align;
    float delta;
};

// host: validate + derive
static inline int make_pyr(Pyr &P, const svr_pyramid *h) {
    SVR_REQUIRE(h, "pyramid: null");
    SVR_REQUIRE(h->n_levels >= 2 && h->n_levels <= SVR_MAX_LEVELS, "pyramid: n_levels must be in [2,%d]", SVR_MAX_LEVELS);
    SVR_REQUIRE(h->channels[0] == 1, "pyramid: level 0 must have exactly one channel");
    P.n_levels = h->n_levels;
    int u = 1, c = 0;
    for (int l = 0; l < h->n_levels; ++l) {
        SVR_REQUIRE(h->dims[l][0] > 0 && h->dims[l][1] > 0 && h->dims[l][2] > 0, "pyramid: empty level %d", l);
        SVR_REQUIRE(l == 0 || (h->channels[l] > 0 && h->channels[l] % 8 == 0), "pyramid: level %d channels must be a multiple of 8", l);
        P.C[l] = h->channels[l];
        P.D[l] = h->dims[l][0];
        P.H[l] = h->dims[l][1];
        P.W[l] = h->dims[l][2];
        P.coff[l] = c;
        c += h->channels[l];
        if (l == 0) {
            P.upd[0] = 1;
            P.ubase[0] = 0;
        } else {
            P.upd[l] = h->channels[l] / 8;
            if (h->channels[l] % 64 == 0) u = ((u + 7) / 8) * 8;   // chunk-align coarse (wide) levels
            P.ubase[l] = u;
            u += 7 * P.upd[l];
        }
    }
    for (int l = h->n_levels; l <= SVR_MAX_LEVELS; ++l) P.ubase[l] = u;
    for (int l = h->n_levels; l < SVR_MAX_LEVELS; ++l) P.C[l] = P.D[l] = P.H[l] = P.W[l] = P.upd[l] = P.coff[l] = 0;
    P.n_units = u;
    P.kp = ((u + 7) / 8) * 64;
    P.ctot = c;
    P.align = h->align_corners ? 1 : 0;
    P.delta = h->displacement;
    return 0;
}

#ifdef __CUDACC__
// unit -> (level, stencil index, first channel); returns false for padding units
__device__ __forceinline__ bool decode_unit(const Pyr &P, int u, int &level, int &d, int &c0) {
    if (u >= P.n_units) return false;
    level = 0;
#pragma unroll
    for (int l = 1; l < SVR_MAX_LEVELS; ++l)
        if (l < P.n_levels && u >= P.ubase[l]) level = l;
    int t = u - P.ubase[level];
    if (t >= (level == 0 ? 1 : 7 * P.upd[level])) return false;   // alignment padding between levels
    d = t / P.upd[level];
    c0 = (t - d * P.upd[level]) * 8;
    return true;
}

// 8 trilinear corners of one stencil sample in one level
struct Corners {
    int x0, y0, z0;
    float wx[2], wy[2], wz[2];
};

// model/ifnet.py:156-159 + F.grid_sample unnormalisation (ATen GridSampler.h:27-35)
__device__ __forceinline__ void stencil_corners(const Pyr &P, int level, int d, float px, float py, float pz, Corners &c) {
    // point coords are (D,H,W)-ordered; grid_sample's x indexes W, y indexes H, z indexes D
    // displacement order (ifnet.py:144-153): d=1,2 -> x -/+, d=3,4 -> y -/+, d=5,6 -> z -/+
    // (selects instead of a dynamically indexed local array: everything stays in registers)
    const float sgn = (d & 1) ? -P.delta : P.delta;
    float q[3];
    q[0] = __fmul_rn(2.0f, pz);
    q[1] = __fmul_rn(2.0f, py);
    q[2] = __fmul_rn(2.0f, px);
    if (d == 1 || d == 2) q[0] = __fadd_rn(q[0], sgn);
    if (d == 3 || d == 4) q[1] = __fadd_rn(q[1], sgn);
    if (d == 5 || d == 6) q[2] = __fadd_rn(q[2], sgn);
    const int size[3] = {P.W[level], P.H[level], P.D[level]};
    float idx[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float sz = (float)size[a];
        idx[a] = P.align ? __fmul_rn(__fmul_rn(__fadd_rn(q[a], 1.0f), 0.5f), sz - 1.0f)
                         : __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(q[a], 1.0f), sz), 1.0f), 0.5f);
    }
    float fx = floorf(idx[0]), fy = floorf(idx[1]), fz = floorf(idx[2]);
    // clamp the integer base so that wildly out-of-range points cannot overflow int arithmetic;
    // every corner of such a sample is out of bounds and contributes zero either way
    fx = fminf(fmaxf(fx, -4.0f), (float)size[0] + 2.0f);
    fy = fminf(fmaxf(fy, -4.0f), (float)size[1] + 2.0f);
    fz = fminf(fmaxf(fz, -4.0f), (float)size[2] + 2.0f);
    c.x0 = (int)fx;
    c.y0 = (int)fy;
    c.z0 = (int)fz;
    c.wx[1] = idx[0] - fx;
    c.wx[0] = (fx + 1.0f) - idx[0];
    c.wy[1] = idx[1] - fy;
    c.wy[0] = (fy + 1.0f) - idx[1];
    c.wz[1] = idx[2] - fz;
    c.wz[0] = (fz + 1.0f) - idx[2];
}

__device__ __forceinline__ bool corner_in(const Pyr &P, int level, int x, int y, int z) {
    return x >= 0 && y >= 0 && z >= 0 && x < P.W[level] && y < P.H[level] && z < P.D[level];
}

__device__ __forceinline__ void bf16x8_to_float(const uint4 &raw, float (&f)[8]) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

__device__ __forceinline__ uint4 float8_to_bf16(const float (&f)[8]) {
    __nv_bfloat162 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return *reinterpret_cast<uint4 *>(h);
}

// Stencil sample dd of the level-0 (input occupancy) grid at one point: zero-padded trilinear interpolation with the
// reference's corner order.
__device__ __forceinline__ float level0_sample(const Pyr &P, int dd, float px, float py, float pz, const float *__restrict__ x0_b) {
    Corners c;
    stencil_corners(P, 0, dd, px, py, pz, c);
    // branch-free: all 8 (clamped, always valid) loads are issued back to back -- with a branch per corner the
    // loads of this 33 MB/scene fp32 grid were serialised into 8 dependent DRAM round trips per sample
    float v[8];
    bool in[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int x = c.x0 + (k & 1), y = c.y0 + ((k >> 1) & 1), z = c.z0 + (k >> 2);
        in[k] = corner_in(P, 0, x, y, z);
        const int xc = min(max(x, 0), P.W[0] - 1), yc = min(max(y, 0), P.H[0] - 1), zc = min(max(z, 0), P.D[0] - 1);
        v[k] = __ldg(x0_b + ((int64_t)zc * P.H[0] + yc) * P.W[0] + xc);
    }
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {     // same order and the same fused multiply-adds as the per-corner form
        const float w = c.wx[k & 1] * c.wy[(k >> 1) & 1] * c.wz[k >> 2];
        a = in[k] ? fmaf(v[k], w, a) : a;
    }
    return a;
}

// One unit of the feature row of a point: 8 bf16 values packed in a uint4.
//   x0_b  : level-0 grid of the point's scene (fp32, D*H*W)
//   vol_l : bf16 NDHWC volume base of the point's scene for `level` (unused for level 0)
__device__ __forceinline__ uint4 gather_unit_decoded(const Pyr &P, int level, int d, int c0, float px, float py, float pz,
                                                     const float *__restrict__ x0_b, const __nv_bfloat16 *__restrict__ vol_l) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (level == 0) {
#pragma unroll
        for (int dd = 0; dd < 7; ++dd) acc[dd] = level0_sample(P, dd, px, py, pz, x0_b);
        return float8_to_bf16(acc);
    }
    Corners c;
    stencil_corners(P, level, d, px, py, pz, c);
    const __nv_bfloat16 *vb = vol_l + c0;
    const int C = P.C[level];
    uint4 raw[8];
    float w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int aa = k & 1, b = (k >> 1) & 1, e = k >> 2;
        int x = c.x0 + aa, y = c.y0 + b, z = c.z0 + e;
        bool in = corner_in(P, level, x, y, z);
        w[k] = in ? c.wx[aa] * c.wy[b] * c.wz[e] : 0.f;
        raw[k] = in ? __ldg(reinterpret_cast<const uint4 *>(vb + (((int64_t)z * P.H[level] + y) * P.W[level] + x) * C))
                    : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float f[8];
        bf16x8_to_float(raw[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[j], w[k], acc[j]);
    }
    return float8_to_bf16(acc);
}

__device__ __forceinline__ uint4 gather_unit(const Pyr &P, int u, float px, float py, float pz, const float *__restrict__ x0_b,
                                             const __nv_bfloat16 *const *vol_b) {
    int level, d, c0;
    if (!decode_unit(P, u, level, d, c0)) return make_uint4(0, 0, 0, 0);
    return gather_unit_decoded(P, level, d, c0, px, py, pz, x0_b, level > 0 ? vol_b[level] : nullptr);
}
// ------------------------------------------------------------------------------------------------
// Fast path used by the fused kernel: everything that depends only on the UNIT (level geometry,
// displacement, channel offset, strides) is hoisted into a UnitCtx built once per K-chunk; per point
// only the index math, 8 vector loads and the blend remain.  The blend uses packed FFMA2
// (fma.rn.f32x2, sm_100+): two fp32 channels per instruction, fp32 accumulation.
// ------------------------------------------------------------------------------------------------
struct UnitCtx {
    int level, d, c0;
    bool real;
    bool coarse;                 // small grid: many border samples, take the predicated path without branching
    int W, H, D, C;
    int sy, sz;                  // element strides of y and z steps (W*C, H*W*C)
    int64_t scene_stride;        // elements per scene
    float fw, fh, fd;            // sizes as float
    float dx, dy, dz;            // stencil displacement of this unit (grid_sample x,y,z order)
    const __nv_bfloat16 *base;   // volume of the level + c0
};

// decode_unit result packed in 32 bits (level | d << 4 | c0 << 8 | real << 31): kernels that walk the units many
// times decode them once into a shared-memory table (the decode has an integer division and a level search)
__device__ __forceinline__ uint32_t pack_unit(const Pyr &P, int u) {
    int level, d, c0;
    if (!decode_unit(P, u, level, d, c0)) return 0u;
    return (uint32_t)level | ((uint32_t)d << 4) | ((uint32_t)c0 << 8) | 0x80000000u;
}

__device__ __forceinline__ void make_unit_ctx_packed(const Pyr &P, uint32_t packed, const __nv_bfloat16 *const *vols, UnitCtx &c);

__device__ __forceinline__ void make_unit_ctx(const Pyr &P, int u, const __nv_bfloat16 *const *vols, UnitCtx &c) {
    make_unit_ctx_packed(P, pack_unit(P, u), vols, c);
}

__device__ __forceinline__ void make_unit_ctx_packed(const Pyr &P, uint32_t packed, const __nv_bfloat16 *const *vols, UnitCtx &c) {
    c.real = (packed >> 31) != 0;
    c.level = (int)(packed & 15u);
    c.d = (int)((packed >> 4) & 15u);
    c.c0 = (int)((packed >> 8) & 0x7fffffu);
    const int l = c.level;
    c.W = P.W[l];
    c.H = P.H[l];
    c.D = P.D[l];
    c.C = P.C[l];
    c.sy = c.W * c.C;
    c.sz = c.H * c.sy;
    c.scene_stride = (int64_t)c.D * c.sz;
    c.fw = (float)c.W;
    c.fh = (float)c.H;
    c.fd = (float)c.D;
    const float sgn = (c.d & 1) ? -P.delta : P.delta;
    c.dx = (c.d == 1 || c.d == 2) ? sgn : 0.f;
    c.dy = (c.d == 3 || c.d == 4) ? sgn : 0.f;
    c.dz = (c.d == 5 || c.d == 6) ? sgn : 0.f;
    c.base = l > 0 ? vols[l] + c.c0 : nullptr;
    // On a 16^3 (8^3) grid 18 % (33 %) of the samples touch the zero padding, so a warp of 4 points almost always
    // holds both kinds and a branch would execute the interior AND the border code; on the fine grids the border is rare.
    c.coarse = c.W <= 16 || c.H <= 16 || c.D <= 16;
}

__device__ __forceinline__ float unnorm(float q, float size, int align) {
    return align ? __fmul_rn(__fmul_rn(__fadd_rn(q, 1.0f), 0.5f), size - 1.0f)
                 : __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(q, 1.0f), size), 1.0f), 0.5f);
}

__device__ __forceinline__ void ffma2(unsigned long long &acc, uint32_t packed_bf16x2, float w) {
    // (lo, hi) bf16 pair -> two fp32 lanes of a 64-bit register; acc += pair * w
    unsigned long long v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(packed_bf16x2 << 16), "r"(packed_bf16x2 & 0xffff0000u));
    unsigned long long ww;
    asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(v), "l"(ww));
}

// one unit (level >= 1) of one point; returns 8 bf16
__device__ __forceinline__ uint4 gather_unit_fast(const UnitCtx &c, int align, float px, float py, float pz, int scene) {
    // q = 2*p + displacement: 2*p is exact, so the fused form rounds exactly like the reference's mul + add
    const float ix = unnorm(__fadd_rn(__fmul_rn(2.0f, pz), c.dx), c.fw, align);
    const float iy = unnorm(__fadd_rn(__fmul_rn(2.0f, py), c.dy), c.fh, align);
    const float iz = unnorm(__fadd_rn(__fmul_rn(2.0f, px), c.dz), c.fd, align);
    float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    fx = fminf(fmaxf(fx, -4.0f), c.fw + 2.0f);
    fy = fminf(fmaxf(fy, -4.0f), c.fh + 2.0f);
    fz = fminf(fmaxf(fz, -4.0f), c.fd + 2.0f);
    const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
    const float wx1 = ix - fx, wx0 = (fx + 1.0f) - ix;
    const float wy1 = iy - fy, wy0 = (fy + 1.0f) - iy;
    const float wz1 = iz - fz, wz0 = (fz + 1.0f) - iz;
    const __nv_bfloat16 *ptr = c.base + (int64_t)scene * c.scene_stride + ((z0 * c.H + y0) * c.W + x0) * c.C;
    const bool interior = x0 >= 0 && y0 >= 0 && z0 >= 0 && x0 + 1 < c.W && y0 + 1 < c.H && z0 + 1 < c.D;
    uint4 raw[8];
    float w[8];
    const float wxy[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
    if (interior && !c.coarse) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int off = ((k & 1) ? c.C : 0) + ((k & 2) ? c.sy : 0) + ((k & 4) ? c.sz : 0);
            raw[k] = __ldg(reinterpret_cast<const uint4 *>(ptr + off));
            w[k] = wxy[k & 3] * ((k & 4) ? wz1 : wz0);
        }
    } else {
        const bool vx[2] = {(unsigned)x0 < (unsigned)c.W, (unsigned)(x0 + 1) < (unsigned)c.W};
        const bool vy[2] = {(unsigned)y0 < (unsigned)c.H, (unsigned)(y0 + 1) < (unsigned)c.H};
        const bool vz[2] = {(unsigned)z0 < (unsigned)c.D, (unsigned)(z0 + 1) < (unsigned)c.D};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const bool in = vx[k & 1] && vy[(k >> 1) & 1] && vz[k >> 2];
            const int off = ((k & 1) ? c.C : 0) + ((k & 2) ? c.sy : 0) + ((k & 4) ? c.sz : 0);
            raw[k] = in ? __ldg(reinterpret_cast<const uint4 *>(ptr + off)) : make_uint4(0, 0, 0, 0);
            w[k] = in ? wxy[k & 3] * ((k & 4) ? wz1 : wz0) : 0.f;
        }
    }
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
// Branch-free variant used by the fused kernel for the levels that have no halo'd copy: every corner coordinate is clamped
// into the volume and all 8 loads are issued unconditionally; the weight of an out-of-range corner is forced to zero
// (fmaf(v, 0, acc) == acc for finite v: the bits of the bounds-checked form), a cell that lies entirely outside gets zero
// weights.  No divergence between interior and border rows of a warp, and the loads are not serialised behind predicates.
__device__ __forceinline__ uint4 gather_unit_bf(const UnitCtx &c, int align, float px, float py, float pz, int scene) {
    const float ix = unnorm(__fadd_rn(__fmul_rn(2.0f, pz), c.dx), c.fw, align);
    const float iy = unnorm(__fadd_rn(__fmul_rn(2.0f, py), c.dy), c.fh, align);
    const float iz = unnorm(__fadd_rn(__fmul_rn(2.0f, px), c.dz), c.fd, align);
    const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    const bool cell_in = fx >= -1.0f && fx <= c.fw - 1.0f && fy >= -1.0f && fy <= c.fh - 1.0f && fz >= -1.0f && fz <= c.fd - 1.0f;
    const int x0 = (int)fminf(fmaxf(fx, -1.0f), c.fw - 1.0f), y0 = (int)fminf(fmaxf(fy, -1.0f), c.fh - 1.0f),
              z0 = (int)fminf(fmaxf(fz, -1.0f), c.fd - 1.0f);
    const float wx[2] = {x0 >= 0 ? (fx + 1.0f) - ix : 0.f, x0 + 1 < c.W ? ix - fx : 0.f};
    const float wy[2] = {y0 >= 0 ? (fy + 1.0f) - iy : 0.f, y0 + 1 < c.H ? iy - fy : 0.f};
    const float wz[2] = {(cell_in && z0 >= 0) ? (fz + 1.0f) - iz : 0.f, (cell_in && z0 + 1 < c.D) ? iz - fz : 0.f};
    const int ox[2] = {max(x0, 0) * c.C, min(x0 + 1, c.W - 1) * c.C};
    const int oy[2] = {max(y0, 0) * c.sy, min(y0 + 1, c.H - 1) * c.sy};
    const int oz[2] = {max(z0, 0) * c.sz, min(z0 + 1, c.D - 1) * c.sz};
    const __nv_bfloat16 *base = c.base + (int64_t)scene * c.scene_stride;
    const float wxy[4] = {wx[0] * wy[0], wx[1] * wy[0], wx[0] * wy[1], wx[1] * wy[1]};
    uint4 raw[8];
    float w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        raw[k] = __ldg(reinterpret_cast<const uint4 *>(base + (oz[k >> 2] + oy[(k >> 1) & 1] + ox[k & 1])));
        w[k] = wxy[k & 3] * wz[k >> 2];
    }
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

#endif  // __CUDACC__

}  // namespace svr
