// Multi-scale trilinear stencil gather (forward) and scatter-add (backward), volume / weight
// packing.  Replaces the 6x F.grid_sample + torch.cat of model/ifnet.py:156-197 (reference root)
// and their autograd backward (grid_sampler_3d_backward).
#include "common.cuh"
#include "sampling.cuh"

namespace svr {

int launch_scatter_tc(const float *points, const int *perm, const int *cell_start, int n_cells, int N, int64_t total_rows, const Pyr &P,
                      const __nv_bfloat16 *dfeat, float *const *gvols, int level_mask, cudaStream_t st);   // scatter_tc.cu

struct VolPtrs {
    const __nv_bfloat16 *v[SVR_MAX_LEVELS];
};
struct GradPtrs {
    float *g[SVR_MAX_LEVELS];
};

// ------------------------------------------------------------------------------------------------
// forward gather: one thread per (point, unit); consecutive threads = consecutive units of a point
// so that the 16-byte loads of a corner and the 16-byte stores of the feature row coalesce.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_fwd_kernel(const float *__restrict__ points, int N, int64_t total_pts,
                                                         const float *__restrict__ x0, VolPtrs vols, Pyr P,
                                                         __nv_bfloat16 *__restrict__ feat) {
    const int UP = P.kp / 8;
    const int64_t total = total_pts * UP;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pt = t / UP;
        const int u = (int)(t - pt * UP);
        const int b = (int)(pt / N);
        const float px = points[pt * 3 + 0], py = points[pt * 3 + 1], pz = points[pt * 3 + 2];
        const __nv_bfloat16 *vb[SVR_MAX_LEVELS];
#pragma unroll
        for (int l = 1; l < SVR_MAX_LEVELS; ++l)
            vb[l] = l < P.n_levels ? vols.v[l] + (int64_t)b * P.D[l] * P.H[l] * P.W[l] * P.C[l] : nullptr;
        vb[0] = nullptr;
        const float *x0b = x0 + (int64_t)b * P.D[0] * P.H[0] * P.W[0];
        uint4 r = gather_unit(P, u, px, py, pz, x0b, vb);
        *reinterpret_cast<uint4 *>(feat + pt * P.kp + (int64_t)u * 8) = r;
    }
}

// ------------------------------------------------------------------------------------------------
// backward scatter-add.  dfeat unit (8 channels) x 8 corners -> 16-byte vector reductions
// (red.global.add.v4.f32) into the fp32 channel-last gradient volumes.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool NEED_DPTS>
__global__ void __launch_bounds__(256) gather_bwd_kernel(const float *__restrict__ points, const int *__restrict__ perm, int N,
                                                         int64_t total_pts,
                                                         const float *__restrict__ x0, VolPtrs vols, Pyr P,
                                                         const __nv_bfloat16 *__restrict__ dfeat, float *__restrict__ gx0,
                                                         GradPtrs gv, float *__restrict__ gpoints, int red_level_mask,
                                                         int u_lo, int u_cnt) {
    // threads cover the unit window [u_lo, u_lo + u_cnt) of every row (all units when d(points) is needed)
    const int UP = u_cnt;
    const int64_t total = total_pts * UP;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = t < total;
    const int64_t row = active ? t / UP : total_pts - 1;
    const int u = active ? u_lo + (int)(t - row * UP) : P.n_units;  // inactive -> padding unit
    const int64_t pt = perm ? (int64_t)perm[row] : row;       // dfeat rows are in processing order
    const int b = (int)(pt / N);
    const float px = points[pt * 3 + 0], py = points[pt * 3 + 1], pz = points[pt * 3 + 2];
    float dq[3] = {0.f, 0.f, 0.f};  // d loss / d (x,y,z) sample coordinate, already scaled to normalised units
    int level, d, c0;
    bool work = decode_unit(P, u, level, d, c0);
    if (work && !NEED_DPTS && level > 0 && !((red_level_mask >> level) & 1)) work = false;
    if (work && !NEED_DPTS && level == 0 && !gx0) work = false;
    if (work) {
        float g[8];
        bf16x8_to_float(*reinterpret_cast<const uint4 *>(dfeat + row * P.kp + (int64_t)u * 8), g);
        if (level == 0) {
            const int64_t base = (int64_t)b * P.D[0] * P.H[0] * P.W[0];
#pragma unroll
            for (int dd = 0; dd < 7; ++dd) {
                Corners c;
                stencil_corners(P, 0, dd, px, py, pz, c);
                float gi[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    int aa = k & 1, bb = (k >> 1) & 1, e = k >> 2;
                    int x = c.x0 + aa, y = c.y0 + bb, z = c.z0 + e;
                    if (!corner_in(P, 0, x, y, z)) continue;
                    int64_t off = base + ((int64_t)z * P.H[0] + y) * P.W[0] + x;
                    if (gx0) atomicAdd(gx0 + off, g[dd] * (c.wx[aa] * c.wy[bb] * c.wz[e]));
                    if (NEED_DPTS) {
                        float v = __ldg(x0 + off) * g[dd];
                        gi[0] += v * (aa ? 1.f : -1.f) * c.wy[bb] * c.wz[e];
                        gi[1] += v * (bb ? 1.f : -1.f) * c.wx[aa] * c.wz[e];
                        gi[2] += v * (e ? 1.f : -1.f) * c.wx[aa] * c.wy[bb];
                    }
                }
                if (NEED_DPTS) {
                    const float sc[3] = {P.align ? 0.5f * (P.W[0] - 1) : 0.5f * P.W[0], P.align ? 0.5f * (P.H[0] - 1) : 0.5f * P.H[0],
                                         P.align ? 0.5f * (P.D[0] - 1) : 0.5f * P.D[0]};
#pragma unroll
                    for (int a = 0; a < 3; ++a) dq[a] += gi[a] * sc[a];
                }
            }
        } else {
            Corners c;
            stencil_corners(P, level, d, px, py, pz, c);
            const int C = P.C[level];
            const int64_t vbase = (int64_t)b * P.D[level] * P.H[level] * P.W[level] * C + c0;
            float gi[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                int aa = k & 1, bb = (k >> 1) & 1, e = k >> 2;
                int x = c.x0 + aa, y = c.y0 + bb, z = c.z0 + e;
                if (!corner_in(P, level, x, y, z)) continue;
                int64_t off = vbase + (((int64_t)z * P.H[level] + y) * P.W[level] + x) * C;
                float w = c.wx[aa] * c.wy[bb] * c.wz[e];
                if (gv.g[level] && ((red_level_mask >> level) & 1)) {
                    float *dst = gv.g[level] + off;
                    red_add_v4(dst, g[0] * w, g[1] * w, g[2] * w, g[3] * w);
                    red_add_v4(dst + 4, g[4] * w, g[5] * w, g[6] * w, g[7] * w);
                }
                if (NEED_DPTS) {
                    float f[8];
                    bf16x8_to_float(__ldg(reinterpret_cast<const uint4 *>(vols.v[level] + off)), f);
                    float dot = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) dot = fmaf(f[j], g[j], dot);
                    gi[0] += dot * (aa ? 1.f : -1.f) * c.wy[bb] * c.wz[e];
                    gi[1] += dot * (bb ? 1.f : -1.f) * c.wx[aa] * c.wz[e];
                    gi[2] += dot * (e ? 1.f : -1.f) * c.wx[aa] * c.wy[bb];
                }
            }
            if (NEED_DPTS) {
                dq[0] = gi[0] * (P.align ? 0.5f * (P.W[level] - 1) : 0.5f * P.W[level]);
                dq[1] = gi[1] * (P.align ? 0.5f * (P.H[level] - 1) : 0.5f * P.H[level]);
                dq[2] = gi[2] * (P.align ? 0.5f * (P.D[level] - 1) : 0.5f * P.D[level]);
            }
        }
    }
    if (NEED_DPTS) {
        // d pt[0] = 2*dq.z, d pt[1] = 2*dq.y, d pt[2] = 2*dq.x.  A warp spans at most two points
        // (a point has >= 32 units): reduce both halves with shuffles, one atomic per point & axis.
        const int64_t first = __shfl_sync(0xffffffffu, pt, 0);
        const int lane = threadIdx.x & 31;
        float out[3] = {2.f * dq[2], 2.f * dq[1], 2.f * dq[0]};
        if (UP < 32) {   // tiny windows: a warp may span more than two points
#pragma unroll
            for (int a = 0; a < 3; ++a)
                if (active && out[a] != 0.f) atomicAdd(gpoints + pt * 3 + a, out[a]);
            return;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float va = (pt == first) ? out[a] : 0.f;
            float vb = (pt == first) ? 0.f : out[a];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                va += __shfl_xor_sync(0xffffffffu, va, o);
                vb += __shfl_xor_sync(0xffffffffu, vb, o);
            }
            const int64_t last = __shfl_sync(0xffffffffu, pt, 31);
            if (lane == 0 && va != 0.f) atomicAdd(gpoints + first * 3 + a, va);
            if (lane == 31 && last != first && vb != 0.f) atomicAdd(gpoints + last * 3 + a, vb);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// packing
// ------------------------------------------------------------------------------------------------
// fp32 (B,C,D,H,W) with arbitrary strides -> bf16 NDHWC; tile = 32 spatial positions x C channels
__global__ void __launch_bounds__(256) pack_volume_kernel(const float *__restrict__ src, int C, int D, int H, int W,
                                                          int64_t spatial_total, int64_t sB, int64_t sC, int64_t sD,
                                                          int64_t sH, int64_t sW, __nv_bfloat16 *__restrict__ dst) {
    extern __shared__ float tile[];   // [C][33]
    const int64_t s0 = (int64_t)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t s = s0 + tx;
    int64_t off = 0;
    const bool ok = s < spatial_total;
    if (ok) {
        int w = (int)(s % W), h = (int)((s / W) % H), d = (int)((s / ((int64_t)W * H)) % D);
        int64_t b = s / ((int64_t)W * H * D);
        off = b * sB + d * sD + h * sH + w * sW;
    }
    for (int c = ty; c < C; c += 8) tile[c * 33 + tx] = ok ? src[off + c * sC] : 0.f;
    __syncthreads();
    const int half_c = C / 2;
    for (int i = threadIdx.x; i < 32 * half_c; i += 256) {
        int pos = i / half_c, cp = i - pos * half_c;
        if (s0 + pos < spatial_total) {
            __nv_bfloat162 v = __floats2bfloat162_rn(tile[(2 * cp) * 33 + pos], tile[(2 * cp + 1) * 33 + pos]);
            *reinterpret_cast<__nv_bfloat162 *>(dst + (s0 + pos) * C + 2 * cp) = v;
        }
    }
}

// fast path: source already channel-last contiguous -> plain vectorised convert (8 elements/thread)
__global__ void __launch_bounds__(256) convert_bf16_kernel(const float *__restrict__ src, int64_t n8, __nv_bfloat16 *__restrict__ dst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(src) + 2 * i);
        const float4 b = __ldg(reinterpret_cast<const float4 *>(src) + 2 * i + 1);
        const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        reinterpret_cast<uint4 *>(dst)[i] = float8_to_bf16(f);
    }
}

// fp32 (B,C,D,H,W) with arbitrary strides -> bf16 (B, D+2, H+2, W+2, C) with a one-voxel ZERO halo: the wide path of the
// fused query kernel reads the zeros of grid_sample's zero padding from the halo instead of testing every corner.
// thread = (halo'd voxel, 8-channel group); only used for the coarse levels (a few MB per scene)
__global__ void __launch_bounds__(256) pack_volume_halo_kernel(const float *__restrict__ src, int C, int D, int H, int W, int64_t total,
                                                               int64_t sB, int64_t sC, int64_t sD, int64_t sH, int64_t sW,
                                                               __nv_bfloat16 *__restrict__ dst) {
    const int cg = C / 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int g = (int)(i % cg);
        int64_t v = i / cg;
        const int x = (int)(v % (W + 2)) - 1;
        v /= (W + 2);
        const int y = (int)(v % (H + 2)) - 1;
        v /= (H + 2);
        const int z = (int)(v % (D + 2)) - 1;
        const int64_t b = v / (D + 2);
        float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (x >= 0 && y >= 0 && z >= 0 && x < W && y < H && z < D) {
            const float *sp = src + b * sB + z * sD + y * sH + x * sW + (int64_t)g * 8 * sC;
            if (sC == 1 && (((uintptr_t)sp) & 15) == 0) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(sp)), c = __ldg(reinterpret_cast<const float4 *>(sp) + 1);
                f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = c.x; f[5] = c.y; f[6] = c.z; f[7] = c.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = __ldg(sp + j * sC);
            }
        }
        reinterpret_cast<uint4 *>(dst)[i] = float8_to_bf16(f);
    }
}

// fp32 NDHWC -> (+)= strided fp32 (B,C,D,H,W)
__global__ void __launch_bounds__(256) unpack_volume_grad_kernel(const float *__restrict__ src, int C, int D, int H, int W,
                                                                 int64_t spatial_total, int64_t sB, int64_t sC, int64_t sD,
                                                                 int64_t sH, int64_t sW, float *__restrict__ dst,
                                                                 int accumulate) {
    extern __shared__ float tile[];   // [C][33]
    const int64_t s0 = (int64_t)blockIdx.x * 32;
    for (int i = threadIdx.x; i < 32 * C; i += 256) {
        int pos = i / C, c = i - pos * C;
        tile[c * 33 + pos] = (s0 + pos < spatial_total) ? src[(s0 + pos) * C + c] : 0.f;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t s = s0 + tx;
    if (s >= spatial_total) return;
    int w = (int)(s % W), h = (int)((s / W) % H), d = (int)((s / ((int64_t)W * H)) % D);
    int64_t b = s / ((int64_t)W * H * D);
    int64_t off = b * sB + d * sD + h * sH + w * sW;
    for (int c = ty; c < C; c += 8) {
        float v = tile[c * 33 + tx];
        float *p = dst + off + c * sC;
        *p = accumulate ? *p + v : v;
    }
}

__global__ void pack_w0_kernel(const float *__restrict__ w0, int H0, Pyr P, __nv_bfloat16 *__restrict__ w0p,
                               __nv_bfloat16 *__restrict__ w0pT) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)H0 * P.kp) return;
    int o = (int)(i / P.kp), kprime = (int)(i % P.kp);
    int u = kprime >> 3, j = kprime & 7;
    const int K = P.ctot * 7;
    float v = 0.f;
    int level, d, c0;
    if (decode_unit(P, u, level, d, c0)) {
        if (level == 0) {
            if (j < 7) v = w0[(int64_t)o * K + j];   // channel 0, stencil j
        } else {
            v = w0[(int64_t)o * K + (int64_t)(P.coff[level] + c0 + j) * 7 + d];
        }
    }
    __nv_bfloat16 h = __float2bfloat16(v);
    if (w0p) w0p[i] = h;
    if (w0pT) w0pT[(int64_t)kprime * H0 + o] = h;
}

__global__ void unpack_w0_grad_kernel(const float *__restrict__ gw0p, int H0, Pyr P, float *__restrict__ gw0) {
    const int K = P.ctot * 7;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)H0 * K) return;
    int o = (int)(i / K), k = (int)(i % K);
    int c = k / 7, d = k - c * 7;
    int level = 0;
#pragma unroll
    for (int l = 1; l < SVR_MAX_LEVELS; ++l)
        if (l < P.n_levels && c >= P.coff[l]) level = l;
    int kprime;
    if (level == 0) {
        kprime = d;
    } else {
        int cc = c - P.coff[level];
        kprime = (P.ubase[level] + d * P.upd[level] + (cc >> 3)) * 8 + (cc & 7);
    }
    gw0[i] = gw0p[(int64_t)o * P.kp + kprime];
}

}  // namespace svr

using namespace svr;

extern "C" {

int svr_feature_kp(const svr_pyramid *pyr_host) {
    Pyr P;
    if (make_pyr(P, pyr_host)) return -1;
    return P.kp;
}

int svr_pack_volume(const float *src, int B, int C, int D, int H, int W, int64_t sB, int64_t sC, int64_t sD, int64_t sH,
                    int64_t sW, uint16_t *dst, void *stream) {
    SVR_REQUIRE(src && dst, "pack_volume: null pointer");
    SVR_REQUIRE(C > 0 && C % 2 == 0 && C <= 256, "pack_volume: C must be even and <= 256");
    int64_t spatial = (int64_t)B * D * H * W;
    if (spatial == 0) return 0;
    const bool ndhwc = sC == 1 && sW == C && sH == (int64_t)W * C && sD == (int64_t)H * W * C && sB == (int64_t)D * H * W * C;
    if (ndhwc && C % 8 == 0 && ((uintptr_t)src & 15) == 0) {
        int64_t n8 = spatial * C / 8;
        int64_t blocks = ceil_div<int64_t>(n8, 256);
        int64_t cap = (int64_t)sm_count() * 16;
        convert_bf16_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(src, n8, (__nv_bfloat16 *)dst);
        SVR_LAUNCH_CHECK();
        return 0;
    }
    size_t smem = (size_t)C * 33 * sizeof(float);
    pack_volume_kernel<<<(unsigned)ceil_div<int64_t>(spatial, 32), 256, smem, as_stream(stream)>>>(
        src, C, D, H, W, spatial, sB, sC, sD, sH, sW, (__nv_bfloat16 *)dst);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_pack_volume_halo(const float *src, int B, int C, int D, int H, int W, int64_t sB, int64_t sC, int64_t sD, int64_t sH,
                         int64_t sW, uint16_t *dst, void *stream) {
    SVR_REQUIRE(src && dst, "pack_volume_halo: null pointer");
    SVR_REQUIRE(C > 0 && C % 8 == 0, "pack_volume_halo: C must be a positive multiple of 8");
    SVR_REQUIRE(((uintptr_t)dst & 15) == 0, "pack_volume_halo: dst must be 16-byte aligned");
    const int64_t total = (int64_t)B * (D + 2) * (H + 2) * (W + 2) * (C / 8);
    if (total == 0) return 0;
    int64_t blocks = ceil_div<int64_t>(total, 256), cap = (int64_t)sm_count() * 16;
    pack_volume_halo_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(src, C, D, H, W, total, sB, sC, sD, sH, sW,
                                                                                                 (__nv_bfloat16 *)dst);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_unpack_volume_grad(const float *src_ndhwc, int B, int C, int D, int H, int W, int64_t sB, int64_t sC, int64_t sD,
                           int64_t sH, int64_t sW, float *dst, int accumulate, void *stream) {
    SVR_REQUIRE(src_ndhwc && dst, "unpack_volume_grad: null pointer");
    SVR_REQUIRE(C > 0 && C <= 256, "unpack_volume_grad: C must be <= 256");
    int64_t spatial = (int64_t)B * D * H * W;
    if (spatial == 0) return 0;
    size_t smem = (size_t)C * 33 * sizeof(float);
    unpack_volume_grad_kernel<<<(unsigned)ceil_div<int64_t>(spatial, 32), 256, smem, as_stream(stream)>>>(
        src_ndhwc, C, D, H, W, spatial, sB, sC, sD, sH, sW, dst, accumulate);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_pack_w0(const float *w0, int H0, const svr_pyramid *pyr_host, uint16_t *w0p, uint16_t *w0pT, void *stream) {
    Pyr P;
    if (int rc = make_pyr(P, pyr_host)) return rc;
    SVR_REQUIRE(w0 && (w0p || w0pT) && H0 > 0, "pack_w0: bad arguments");
    int64_t n = (int64_t)H0 * P.kp;
    pack_w0_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(w0, H0, P, (__nv_bfloat16 *)w0p,
                                                                                      (__nv_bfloat16 *)w0pT);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_unpack_w0_grad(const float *gw0p, int H0, const svr_pyramid *pyr_host, float *gw0, void *stream) {
    Pyr P;
    if (int rc = make_pyr(P, pyr_host)) return rc;
    SVR_REQUIRE(gw0p && gw0 && H0 > 0, "unpack_w0_grad: bad arguments");
    int64_t n = (int64_t)H0 * P.ctot * 7;
    unpack_w0_grad_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(gw0p, H0, P, gw0);
    SVR_LAUNCH_CHECK();
    return 0;
}

static int fill_vols(VolPtrs &vp, const uint16_t *const *vols_host, const Pyr &P) {
    SVR_REQUIRE(vols_host, "null volume pointer table");
    for (int l = 0; l < SVR_MAX_LEVELS; ++l) {
        vp.v[l] = (l >= 1 && l < P.n_levels) ? (const __nv_bfloat16 *)vols_host[l] : nullptr;
        SVR_REQUIRE(!(l >= 1 && l < P.n_levels) || vp.v[l], "volume of level %d is null", l);
        SVR_REQUIRE(((uintptr_t)vp.v[l] & 15) == 0, "volume of level %d is not 16-byte aligned", l);
    }
    return 0;
}

int svr_gather_fwd(const float *points, int B, int N, const float *x0, const uint16_t *const *vols_host,
                   const svr_pyramid *pyr_host, uint16_t *feat, void *stream) {
    Pyr P;
    if (int rc = make_pyr(P, pyr_host)) return rc;
    VolPtrs vp;
    if (int rc = fill_vols(vp, vols_host, P)) return rc;
    SVR_REQUIRE(points && x0 && feat, "gather_fwd: null pointer");
    int64_t total_pts = (int64_t)B * N;
    if (total_pts == 0) return 0;
    int64_t total = total_pts * (P.kp / 8);
    int64_t blocks = ceil_div<int64_t>(total, 256);
    int64_t cap = (int64_t)sm_count() * 64;
    if (blocks > cap) blocks = cap;
    gather_fwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(points, N, total_pts, x0, vp, P, (__nv_bfloat16 *)feat);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_gather_bwd(const float *points, const int *perm, const int *cell_start, int B, int N, const float *x0, const uint16_t *const *vols_host,
                   const svr_pyramid *pyr_host, const uint16_t *dfeat, float *gx0, float *const *gvols_host, float *gpoints,
                   void *stream) {
    Pyr P;
    if (int rc = make_pyr(P, pyr_host)) return rc;
    VolPtrs vp;
    if (int rc = fill_vols(vp, vols_host, P)) return rc;
    SVR_REQUIRE(points && x0 && dfeat, "gather_bwd: null pointer");
    GradPtrs gp;
    for (int l = 0; l < SVR_MAX_LEVELS; ++l) {
        gp.g[l] = (gvols_host && l >= 1 && l < P.n_levels) ? gvols_host[l] : nullptr;
        SVR_REQUIRE(((uintptr_t)gp.g[l] & 15) == 0, "gradient volume of level %d is not 16-byte aligned", l);
    }
    int64_t total_pts = (int64_t)B * N;
    if (total_pts == 0) return 0;
    // Coarse levels (<= 32^3-class grids: a sorted tile touches a few hundred voxels at most) go through the
    // tensor-core scatter (scatter_tc.cu) when the rows are spatially sorted (perm given); the fine levels (few
    // contributions per voxel and tile), level 0 and d(points) stay on the direct kernel.
    int agg_mask = 0;
    if (perm && gvols_host)
        for (int l = 1; l < P.n_levels; ++l)
            if (gp.g[l] && (int64_t)P.D[l] * P.H[l] * P.W[l] <= 40 * 40 * 40 && P.C[l] >= 16 && P.C[l] <= 128 && (P.C[l] & (P.C[l] - 1)) == 0)
                agg_mask |= 1 << l;
    const int direct_mask = ~agg_mask;
    bool direct_needed = gpoints || gx0;
    for (int l = 1; l < P.n_levels; ++l)
        if (gp.g[l] && ((direct_mask >> l) & 1)) direct_needed = true;
    if (direct_needed) {
        // unit window of the direct kernel: everything for d(points), else only the units with work
        int u_lo = 0, u_hi = P.kp / 8;
        if (!gpoints) {
            u_lo = P.n_units;
            u_hi = 0;
            if (gx0) {
                u_lo = 0;
                u_hi = 1;
            }
            for (int l = 1; l < P.n_levels; ++l)
                if (gp.g[l] && ((direct_mask >> l) & 1)) {
                    u_lo = u_lo < P.ubase[l] ? u_lo : P.ubase[l];
                    u_hi = u_hi > P.ubase[l + 1] ? u_hi : P.ubase[l + 1];
                }
        }
        const int u_cnt = u_hi - u_lo;
        const int64_t total = total_pts * u_cnt;
        const unsigned blocks = (unsigned)ceil_div<int64_t>(total, 256);
        if (gpoints)
            gather_bwd_kernel<true><<<blocks, 256, 0, as_stream(stream)>>>(points, perm, N, total_pts, x0, vp, P,
                                                                           (const __nv_bfloat16 *)dfeat, gx0, gp, gpoints, direct_mask,
                                                                           u_lo, u_cnt);
        else
            gather_bwd_kernel<false><<<blocks, 256, 0, as_stream(stream)>>>(points, perm, N, total_pts, x0, vp, P,
                                                                            (const __nv_bfloat16 *)dfeat, gx0, gp, gpoints, direct_mask,
                                                                            u_lo, u_cnt);
    }
    if (agg_mask)
        if (int rc = launch_scatter_tc(points, perm, cell_start, B * svr_sort_cells_per_scene(), N, total_pts, P, (const __nv_bfloat16 *)dfeat,
                                       gp.g, agg_mask, as_stream(stream)))
            return rc;
    SVR_LAUNCH_CHECK();
    return 0;
}
}
