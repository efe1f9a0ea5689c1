// First encoder stage of the 128-net fused end to end:  y = BatchNorm3d(relu(Conv3d(1 -> 16, 3x3x3, pad 1)(x)))
// (model/ifnet.py:126,137,164 of the reference: `net = self.actvn(self.conv_in(x)); net = self.conv_in_bn(net)`),
// training-mode batch statistics included, channels-last output.
//
// The pre-BN activation a = relu(conv(x)) is a 0.54 GB tensor at batch 4 x 128^3, yet it is a 27-tap stencil
// of a ONE-channel input (33 MB): recomputing it is cheaper than storing and re-reading it.  So instead of
//   conv+ReLU (write a) -> BN statistics (read a) -> BN apply (read a, write y)            forward
//   BN backward (read a, gy twice, write da) -> ReLU mask + wgrad (read a, da)              backward
// the stage becomes four streaming passes that only ever touch x, y and gy:
//   stats   : a recomputed, per-channel sum / sum of squares            (reads x)
//   apply   : a recomputed, y = a * s + t                               (reads x, writes y)
//   bwd 1   : a recomputed, sum gy and sum gy * xhat                    (reads x, gy)
//   bwd 2   : a recomputed, da -> ReLU mask -> weight/bias gradient     (reads x, gy)
// `a` is recomputed by the same inlined routine everywhere, so the ReLU mask of the backward pass is
// bit-identical to the forward activation.  When the caller keeps y and the ReLU bit mask the apply pass can
// emit (16 bits per voxel), the backward passes take xhat = (y - beta) / gamma and the mask from memory instead
// (0.56 GB more traffic, 0.2 ms less arithmetic per pass at batch 4 x 128^3); they fall back to the recomputation
// if any |gamma| is too small to invert.
//
// Work unit: a warp owns a 64-voxel segment of an x-row; the 3x3 neighbouring rows (with halo) are staged in
// a warp-private shared tile, lane v computes the 16 channels of voxels v and v+32 (f32x2 FMAs, one broadcast
// weight load serves both voxels).  The weight gradient dW[ch][tap] = sum_vox dz[vox][ch] * x[vox + tap] is a
// (16 x 64) . (64 x 32) product per segment: the masked gradients are parked in shared memory and multiplied
// on the tensor cores with mma.sync m16n8k8 TF32, dz split into a TF32 head and tail (two MMAs) so that only
// x is rounded to TF32 (exact for occupancy grids; cuDNN's default wgrad rounds both operands).
#include "common.cuh"

namespace svr {

constexpr int CB_CO = 16;
constexpr int CB_WARPS = 8;
constexpr int CB_SEG = 64;                // voxels per segment
constexpr int CB_RS = 72;                 // tile row stride in floats (66 used); 72 = 8 mod 32 keeps the taps of an
                                          // MMA operand load on distinct banks
constexpr int CB_TILE = 11 * CB_RS;       // 9 stencil rows + a row of ones (bias "tap") + a row of zeros
constexpr int CB_DS = 24;                 // stride of a parked gradient row (16 used): 4 voxel rows x 8 channels conflict-free

struct CbGeom {
    int B, D, H, W, segs_per_row, n_segs;
    int64_t n_vox;
    int pow2, sh_spr, sh_h, sh_d;     // segs_per_row, H and D powers of two: segment decode by shifts instead of divisions
};

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ void fma2(f32x2 &acc, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
__device__ __forceinline__ uint32_t to_tf32(float f) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(f));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct CbShared {
    ulonglong2 wt[27][CB_CO / 4];         // [tap][channel quad] as two f32x2 pairs
    float4 bs[CB_CO / 4];
    float ch[6][CB_CO];                   // per-channel constants of the pass
    float tile[CB_WARPS][2][CB_TILE];     // double-buffered: the next segment's rows arrive by cp.async during the math
};

__device__ __forceinline__ void cb_load_weights(CbShared &s, const float *__restrict__ w, const float *__restrict__ bias) {
    for (int i = threadIdx.x; i < 27 * CB_CO; i += blockDim.x) {
        const int co = i / 27, tap = i - co * 27;
        reinterpret_cast<float *>(&s.wt[tap][0])[co] = w[i];
    }
    if (threadIdx.x < CB_CO) reinterpret_cast<float *>(s.bs)[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int e = lane; e < CB_RS; e += 32)
        for (int b = 0; b < 2; ++b) {
            s.tile[warp][b][9 * CB_RS + e] = 1.f;
            s.tile[warp][b][10 * CB_RS + e] = 0.f;
        }
}

__device__ __forceinline__ void cb_cp4(uint32_t dst_smem, const float *src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cb_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cb_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct CbSeg {
    int64_t p0;   // voxel index of the segment's first voxel
    int x0, yy, zz, b;
};

// Issues the asynchronous copies of the 3x3 rows x0-1 .. x0+64 of x into `tile` (tile[r][e] = x at x0 - 1 + e, zero
// outside the volume) and commits them as one cp.async group.  The caller must have passed a __syncwarp since the
// last reader of `tile`.
__device__ __forceinline__ CbSeg cb_decode(const CbGeom &g, int seg) {
    CbSeg sg;
    if (g.pow2) {
        const int row = seg >> g.sh_spr;
        sg.x0 = (seg & (g.segs_per_row - 1)) * CB_SEG;
        sg.yy = row & (g.H - 1);
        const int r2 = row >> g.sh_h;
        sg.zz = r2 & (g.D - 1);
        sg.b = r2 >> g.sh_d;
        sg.p0 = (((int64_t)sg.b * g.D + sg.zz) * g.H + sg.yy) * g.W + sg.x0;
        return sg;
    }
    const int row = seg / g.segs_per_row;
    sg.x0 = (seg - row * g.segs_per_row) * CB_SEG;
    sg.yy = row % g.H;
    const int r2 = row / g.H;
    sg.zz = r2 % g.D;
    sg.b = r2 / g.D;
    sg.p0 = (((int64_t)sg.b * g.D + sg.zz) * g.H + sg.yy) * g.W + sg.x0;
    return sg;
}

__device__ __forceinline__ CbSeg cb_issue(const CbGeom &g, const float *__restrict__ x, int seg, int lane, float *tile) {
    const CbSeg sg = cb_decode(g, seg);
    const int yy = sg.yy, zz = sg.zz;
    const int64_t base = sg.p0 - sg.x0;
    const int xa = sg.x0 - 1 + lane;
    const uint32_t sz_a = (xa >= 0 && xa < g.W) ? 4u : 0u, sz_b = (xa + 32 < g.W) ? 4u : 0u, sz_c = (xa + 64 < g.W) ? 4u : 0u;
    const float *src0 = x + base + xa;
    const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(tile + lane);
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        const int z = zz + r / 3 - 1, yq = yy + r % 3 - 1;
        if (z >= 0 && z < g.D && yq >= 0 && yq < g.H) {          // warp-uniform
            const float *src = src0 + ((r / 3 - 1) * g.H + (r % 3 - 1)) * g.W;
            cb_cp4(dst0 + r * CB_RS * 4, src, sz_a);
            cb_cp4(dst0 + r * CB_RS * 4 + 128, src + 32, sz_b);
            if (lane < 2) cb_cp4(dst0 + r * CB_RS * 4 + 256, src + 64, sz_c);
        } else {
            tile[r * CB_RS + lane] = 0.f;
            tile[r * CB_RS + 32 + lane] = 0.f;
            if (lane < 2) tile[r * CB_RS + 64 + lane] = 0.f;
        }
    }
    cb_commit();
    return sg;
}

// Segment loop of every pass: the rows of segment i+1 are in flight while `body(tile, segment)` works on segment i.
template <typename Body>
__device__ __forceinline__ void cb_for_each_segment(const CbGeom &g, const float *__restrict__ x, CbShared &s, Body body) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stride = gridDim.x * CB_WARPS;
    int seg = blockIdx.x * CB_WARPS + warp, buf = 0;
    if (seg >= g.n_segs) return;
    CbSeg cur = cb_issue(g, x, seg, lane, s.tile[warp][0]);
    while (true) {
        const int next = seg + stride;
        CbSeg nxt = cur;
        if (next < g.n_segs) {
            nxt = cb_issue(g, x, next, lane, s.tile[warp][buf ^ 1]);
            cb_wait<1>();
        } else {
            cb_wait<0>();
        }
        __syncwarp();                       // every lane's copies of the current tile have landed
        body(s.tile[warp][buf], cur);
        __syncwarp();                       // readers done before the tile is refilled (two iterations ahead)
        if (next >= g.n_segs) break;
        seg = next;
        cur = nxt;
        buf ^= 1;
    }
}

// a[j][c] = relu(bias[c] + sum_tap w[c][tap] * x[voxel_j + tap]) for the lane's voxels lane and lane+32 (fixed
// accumulation order: the same bits in every pass)
__device__ __forceinline__ void cb_conv_relu(const CbShared &s, const float *tile, int lane, float (&a)[2][CB_CO]) {
    f32x2 acc[2][CB_CO / 2];
#pragma unroll
    for (int c = 0; c < CB_CO / 4; ++c) {
        acc[0][2 * c] = acc[1][2 * c] = pack2(s.bs[c].x, s.bs[c].y);
        acc[0][2 * c + 1] = acc[1][2 * c + 1] = pack2(s.bs[c].z, s.bs[c].w);
    }
#pragma unroll
    for (int tap = 0; tap < 27; ++tap) {
        const float v0 = tile[(tap / 3) * CB_RS + lane + tap % 3], v1 = tile[(tap / 3) * CB_RS + 32 + lane + tap % 3];
        const f32x2 vv0 = pack2(v0, v0), vv1 = pack2(v1, v1);
#pragma unroll
        for (int c = 0; c < CB_CO / 4; ++c) {
            const ulonglong2 w = s.wt[tap][c];
            fma2(acc[0][2 * c], vv0, w.x);
            fma2(acc[0][2 * c + 1], vv0, w.y);
            fma2(acc[1][2 * c], vv1, w.x);
            fma2(acc[1][2 * c + 1], vv1, w.y);
        }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < CB_CO / 2; ++c) {
            float lo, hi;
            unpack2(acc[j][c], lo, hi);
            a[j][2 * c] = fmaxf(lo, 0.f);
            a[j][2 * c + 1] = fmaxf(hi, 0.f);
        }
}

// Incoming gradient of y at voxel x0 + lane + 32 j of the segment, channel quad c: the gradient that reaches y directly
// (gy, nullable) plus the share routed through the fused nn.MaxPool3d(2) (g_pool: gradient of the pooled tensor,
// pool_idx: the winner codes svr_maxpool2_cl_fwd wrote, both nullable) -- the pooling backward and the sum of the
// two gradient branches never touch memory.
struct CbGrad {
    const float *gy;
    const float *g_pool;
    const uint32_t *pool_idx;
    const float *y;          // stage output (nullable): with relu_mask, replaces the recomputation of the activation
    const uint16_t *relu_mask;
    const float *gamma, *beta;
};

// true when xhat can be taken from y: y and the mask are given and every gamma is safely invertible
// (s.ch[4] = 1/gamma, s.ch[5] = beta must be loaded; uniform over the grid)
__device__ __forceinline__ bool cb_from_y(const CbGrad &gr) {
    if (!gr.y || !gr.relu_mask) return false;
    bool ok = true;
    for (int c = 0; c < CB_CO; ++c) ok = ok && fabsf(gr.gamma ? gr.gamma[c] : 1.f) >= 1e-6f;
    return ok;
}
__device__ __forceinline__ float4 cb_load_grad(const CbGrad &gr, const CbGeom &g, const CbSeg &sg, int xq, int c) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (xq >= g.W) return v;
    if (gr.gy) v = __ldg(reinterpret_cast<const float4 *>(gr.gy + (sg.p0 + (xq - sg.x0)) * CB_CO) + c);
    const int Do = g.D >> 1, Ho = g.H >> 1, Wo = g.W >> 1;
    const int zo = sg.zz >> 1, yo = sg.yy >> 1, xo = xq >> 1;
    if (gr.g_pool && zo < Do && yo < Ho && xo < Wo) {
        const int64_t i = ((((int64_t)sg.b * Do + zo) * Ho + yo) * Wo + xo) * (CB_CO / 4) + c;
        const uint32_t arg = __ldg(gr.pool_idx + i), k = (uint32_t)((sg.zz & 1) * 4 + (sg.yy & 1) * 2 + (xq & 1));
        const float4 gp = __ldg(reinterpret_cast<const float4 *>(gr.g_pool) + i);
        if ((arg & 0xffu) == k) v.x += gp.x;
        if (((arg >> 8) & 0xffu) == k) v.y += gp.y;
        if (((arg >> 16) & 0xffu) == k) v.z += gp.z;
        if ((arg >> 24) == k) v.w += gp.w;
    }
    return v;
}

// (voxel, quad) item of the from-y passes: all global loads of a batch are issued before any is consumed (a warp keeps
// 4 x 4 independent 16-byte loads in flight; one load-use round trip per item made the passes latency bound).
struct CbItem {
    float4 gy, y, gp;
    uint32_t arg, mask;
    bool ok, pooled;
    uint32_t k;
};
__device__ __forceinline__ void cb_item_load(const CbGrad &gr, const CbGeom &g, const CbSeg &sg, int v, int q, CbItem &it) {
    const int xq = sg.x0 + v;
    it.ok = xq < g.W;
    const int64_t vox = sg.p0 + v;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    it.gy = (it.ok && gr.gy) ? __ldg(reinterpret_cast<const float4 *>(gr.gy + vox * CB_CO) + q) : z4;
    it.y = it.ok ? __ldg(reinterpret_cast<const float4 *>(gr.y + vox * CB_CO) + q) : z4;
    it.mask = it.ok ? (uint32_t)__ldg(gr.relu_mask + vox) >> (4 * q) : 0u;
    const int Do = g.D >> 1, Ho = g.H >> 1, Wo = g.W >> 1;
    const int zo = sg.zz >> 1, yo = sg.yy >> 1, xo = xq >> 1;
    it.pooled = it.ok && gr.g_pool && zo < Do && yo < Ho && xo < Wo;
    const int64_t i = ((((int64_t)sg.b * Do + zo) * Ho + yo) * Wo + xo) * (CB_CO / 4) + q;
    it.arg = it.pooled ? __ldg(gr.pool_idx + i) : 0u;
    it.gp = it.pooled ? __ldg(reinterpret_cast<const float4 *>(gr.g_pool) + i) : z4;
    it.k = (uint32_t)((sg.zz & 1) * 4 + (sg.yy & 1) * 2 + (xq & 1));
}
__device__ __forceinline__ void cb_item_grad(const CbItem &it, float (&gv)[4]) {
    gv[0] = it.gy.x; gv[1] = it.gy.y; gv[2] = it.gy.z; gv[3] = it.gy.w;
    if (it.pooled) {
        if ((it.arg & 0xffu) == it.k) gv[0] += it.gp.x;
        if (((it.arg >> 8) & 0xffu) == it.k) gv[1] += it.gp.y;
        if (((it.arg >> 16) & 0xffu) == it.k) gv[2] += it.gp.z;
        if ((it.arg >> 24) == it.k) gv[3] += it.gp.w;
    }
}

// sum the 2*CB_CO per-lane accumulators over the block, one row of `partial` per block
__device__ __forceinline__ void cb_block_reduce32(float (&s1)[CB_CO], float (&s2)[CB_CO], float *red /* [CB_WARPS][32] */, float *__restrict__ partial) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < CB_CO; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], o);
            s2[c] += __shfl_xor_sync(0xffffffffu, s2[c], o);
        }
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < CB_CO; ++c) {
            red[warp * 32 + c] = s1[c];
            red[warp * 32 + CB_CO + c] = s2[c];
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = 0.f;
        for (int w = 0; w < CB_WARPS; ++w) v += red[w * 32 + threadIdx.x];
        partial[(int64_t)blockIdx.x * 32 + threadIdx.x] = v;
    }
}

// ---- forward pass 1: batch statistics of a -------------------------------------------------------------------
__global__ void __launch_bounds__(CB_WARPS * 32) cb_stats_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                                 const float *__restrict__ bias, const CbGeom g, float *__restrict__ partial) {
    extern __shared__ __align__(16) uint8_t cb_dyn[];
    CbShared &s = *reinterpret_cast<CbShared *>(cb_dyn);
    cb_load_weights(s, w, bias);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float s1[CB_CO], s2[CB_CO];
#pragma unroll
    for (int c = 0; c < CB_CO; ++c) s1[c] = s2[c] = 0.f;
    cb_for_each_segment(g, x, s, [&](const float *tile, const CbSeg &sg) {
        float a[2][CB_CO];
        cb_conv_relu(s, tile, lane, a);
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (sg.x0 + lane + 32 * j < g.W) {
#pragma unroll
                for (int c = 0; c < CB_CO; ++c) {
                    s1[c] += a[j][c];
                    s2[c] = fmaf(a[j][c], a[j][c], s2[c]);
                }
            }
    });
    cb_block_reduce32(s1, s2, &s.tile[0][0][0], partial);   // the tiles are free after the loop (barrier inside)
}

// mean / inverse standard deviation from the block partials (double), running statistics updated like
// torch.nn.BatchNorm3d (momentum update with the unbiased variance)
__global__ void cb_stats_finalize_kernel(const float *__restrict__ partial, int nblocks, double n, float eps, float momentum,
                                         float *__restrict__ running_mean, float *__restrict__ running_var, float *__restrict__ mean,
                                         float *__restrict__ invstd) {
    __shared__ double acc[32][32];
    const int col = threadIdx.x & 31, part = threadIdx.x >> 5;   // 1024 threads
    double v = 0.0;
    for (int b = part; b < nblocks; b += 32) v += (double)partial[(int64_t)b * 32 + col];
    acc[part][col] = v;
    __syncthreads();
    if (threadIdx.x < CB_CO) {
        double s1 = 0.0, s2 = 0.0;
        for (int p = 0; p < 32; ++p) {
            s1 += acc[p][threadIdx.x];
            s2 += acc[p][CB_CO + threadIdx.x];
        }
        const double m = s1 / n;
        double var = s2 / n - m * m;
        if (var < 0.0) var = 0.0;
        mean[threadIdx.x] = (float)m;
        invstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
        if (running_mean) running_mean[threadIdx.x] = (1.f - momentum) * running_mean[threadIdx.x] + momentum * (float)m;
        if (running_var) {
            const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
            running_var[threadIdx.x] = (1.f - momentum) * running_var[threadIdx.x] + momentum * (float)unbiased;
        }
    }
}

// ---- forward pass 2: y = (a - mean) * invstd * gamma + beta ------------------------------------------------------
__global__ void __launch_bounds__(CB_WARPS * 32) cb_apply_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                                 const float *__restrict__ bias, const float *__restrict__ mean,
                                                                 const float *__restrict__ invstd, const float *__restrict__ gamma,
                                                                 const float *__restrict__ beta, const CbGeom g, float *__restrict__ y,
                                                                 uint4 *__restrict__ y_bf16, uint16_t *__restrict__ relu_mask) {
    extern __shared__ __align__(16) uint8_t cb_dyn[];
    CbShared &s = *reinterpret_cast<CbShared *>(cb_dyn);
    cb_load_weights(s, w, bias);
    if (threadIdx.x < CB_CO) {
        const float sc = (gamma ? gamma[threadIdx.x] : 1.f) * invstd[threadIdx.x];
        s.ch[0][threadIdx.x] = sc;
        s.ch[1][threadIdx.x] = (beta ? beta[threadIdx.x] : 0.f) - mean[threadIdx.x] * sc;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    cb_for_each_segment(g, x, s, [&](const float *tile, const CbSeg &sg) {
        float a[2][CB_CO];
        cb_conv_relu(s, tile, lane, a);
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (sg.x0 + lane + 32 * j < g.W) {
                float4 *dst = reinterpret_cast<float4 *>(y + (sg.p0 + lane + 32 * j) * CB_CO);
                float o[CB_CO];
#pragma unroll
                for (int c = 0; c < CB_CO; ++c) o[c] = fmaf(a[j][c], s.ch[0][c], s.ch[1][c]);
#pragma unroll
                for (int c = 0; c < CB_CO / 4; ++c) dst[c] = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                if (relu_mask) {
                    uint32_t m = 0;
#pragma unroll
                    for (int c = 0; c < CB_CO; ++c) m |= (a[j][c] > 0.f ? 1u : 0u) << c;
                    relu_mask[sg.p0 + lane + 32 * j] = (uint16_t)m;
                }
                if (y_bf16) {   // the gather's bf16 NDHWC copy of the volume, saving a separate pack pass
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        __nv_bfloat162 q[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) q[e] = __floats2bfloat162_rn(o[8 * h + 2 * e], o[8 * h + 2 * e + 1]);
                        y_bf16[(sg.p0 + lane + 32 * j) * 2 + h] = *reinterpret_cast<uint4 *>(q);
                    }
                }
            }
    });
}

// ---- backward pass 1: sum gy, sum gy * xhat ---------------------------------------------------------------------
__global__ void __launch_bounds__(CB_WARPS * 32) cb_bwd_reduce_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                                      const float *__restrict__ bias, const float *__restrict__ mean,
                                                                      const float *__restrict__ invstd, const CbGrad gr, const CbGeom g,
                                                                      float *__restrict__ partial) {
    extern __shared__ __align__(16) uint8_t cb_dyn[];
    CbShared &s = *reinterpret_cast<CbShared *>(cb_dyn);
    cb_load_weights(s, w, bias);
    if (threadIdx.x < CB_CO) {
        s.ch[0][threadIdx.x] = mean[threadIdx.x];
        s.ch[1][threadIdx.x] = invstd[threadIdx.x];
        s.ch[4][threadIdx.x] = 1.f / (gr.gamma ? gr.gamma[threadIdx.x] : 1.f);
        s.ch[5][threadIdx.x] = gr.beta ? gr.beta[threadIdx.x] : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float s1[CB_CO], s2[CB_CO];
#pragma unroll
    for (int c = 0; c < CB_CO; ++c) s1[c] = s2[c] = 0.f;
    if (cb_from_y(gr)) {
        // pure stream of gy (+ pooled branch) and y, no stencil.  Lane = (voxel, channel quad): consecutive lanes read
        // consecutive 16-byte pieces, every load instruction is one fully used 512-byte run.
        const int warp = threadIdx.x >> 5, q = lane & 3;
        float bq[4], iq[4], t1[4] = {0.f, 0.f, 0.f, 0.f}, t2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            bq[e] = s.ch[5][4 * q + e];
            iq[e] = s.ch[4][4 * q + e];
        }
        for (int seg = blockIdx.x * CB_WARPS + warp; seg < g.n_segs; seg += gridDim.x * CB_WARPS) {
            const CbSeg sg = cb_decode(g, seg);
#pragma unroll
            for (int i0 = 0; i0 < CB_SEG / 8; i0 += 4) {
                CbItem it[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) cb_item_load(gr, g, sg, (i0 + i) * 8 + (lane >> 2), q, it[i]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float gv[4];
                    cb_item_grad(it[i], gv);
                    const float yy4[4] = {it[i].y.x, it[i].y.y, it[i].y.z, it[i].y.w};
                    if (it[i].ok) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            t1[e] += gv[e];
                            t2[e] = fmaf(gv[e], (yy4[e] - bq[e]) * iq[e], t2[e]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int o = 16; o >= 4; o >>= 1) {
                t1[e] += __shfl_xor_sync(0xffffffffu, t1[e], o);
                t2[e] += __shfl_xor_sync(0xffffffffu, t2[e], o);
            }
        float *red = &s.tile[0][0][0];
        __syncthreads();
        if (lane < 4) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                red[warp * 32 + 4 * lane + e] = t1[e];
                red[warp * 32 + CB_CO + 4 * lane + e] = t2[e];
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            float vsum = 0.f;
            for (int ww = 0; ww < CB_WARPS; ++ww) vsum += red[ww * 32 + threadIdx.x];
            partial[(int64_t)blockIdx.x * 32 + threadIdx.x] = vsum;
        }
        return;
    }
    cb_for_each_segment(g, x, s, [&](const float *tile, const CbSeg &sg) {
        float4 gq[2][CB_CO / 4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int c = 0; c < CB_CO / 4; ++c) gq[j][c] = cb_load_grad(gr, g, sg, sg.x0 + lane + 32 * j, c);
        float a[2][CB_CO];
        cb_conv_relu(s, tile, lane, a);
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int c = 0; c < CB_CO / 4; ++c) {
                const float gv[4] = {gq[j][c].x, gq[j][c].y, gq[j][c].z, gq[j][c].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int ch = 4 * c + e;
                    const float xh = (a[j][ch] - s.ch[0][ch]) * s.ch[1][ch];
                    s1[ch] += gv[e];
                    s2[ch] = fmaf(gv[e], xh, s2[ch]);
                }
            }
    });
    cb_block_reduce32(s1, s2, &s.tile[0][0][0], partial);
}

// ggamma = sum gy * xhat, gbeta = sum gy (double reduce of the block partials)
__global__ void cb_bwd_finalize_kernel(const float *__restrict__ partial, int nblocks, float *__restrict__ gbeta, float *__restrict__ ggamma) {
    __shared__ double acc[32][32];
    const int col = threadIdx.x & 31, part = threadIdx.x >> 5;   // 1024 threads
    double v = 0.0;
    for (int b = part; b < nblocks; b += 32) v += (double)partial[(int64_t)b * 32 + col];
    acc[part][col] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = 0.0;
        for (int p = 0; p < 32; ++p) t += acc[p][threadIdx.x];
        if (threadIdx.x < CB_CO)
            gbeta[threadIdx.x] = (float)t;
        else
            ggamma[threadIdx.x - CB_CO] = (float)t;
    }
}

// ---- backward pass 2: da = gamma * invstd * (gy - mean(gy) - xhat * mean(gy * xhat)), ReLU mask, weight gradient -----
__global__ void __launch_bounds__(CB_WARPS * 32, 2) cb_bwd_wgrad_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                                     const float *__restrict__ bias, const float *__restrict__ mean,
                                                                     const float *__restrict__ invstd, const float *__restrict__ gamma,
                                                                     const float *__restrict__ gbeta, const float *__restrict__ ggamma,
                                                                     const CbGrad gr, const CbGeom g, float inv_n,
                                                                     float *__restrict__ partial /* [grid][28][16] */) {
    extern __shared__ __align__(16) uint8_t cb_dyn[];
    CbShared &s = *reinterpret_cast<CbShared *>(cb_dyn);
    float *park = reinterpret_cast<float *>(cb_dyn + sizeof(CbShared));   // [CB_WARPS][CB_SEG][CB_DS]
    __shared__ float k1[CB_CO];
    cb_load_weights(s, w, bias);
    if (threadIdx.x < CB_CO) {
        s.ch[0][threadIdx.x] = mean[threadIdx.x];
        s.ch[1][threadIdx.x] = invstd[threadIdx.x];
        s.ch[2][threadIdx.x] = (gamma ? gamma[threadIdx.x] : 1.f) * invstd[threadIdx.x];
        s.ch[3][threadIdx.x] = ggamma[threadIdx.x] * inv_n;
        s.ch[4][threadIdx.x] = 1.f / (gr.gamma ? gr.gamma[threadIdx.x] : 1.f);
        s.ch[5][threadIdx.x] = gr.beta ? gr.beta[threadIdx.x] : 0.f;
        k1[threadIdx.x] = gbeta[threadIdx.x] * inv_n;
    }
    __syncthreads();
    const bool from_y = cb_from_y(gr);
    // per-lane constants of the lane's channel quad (from-y path): gamma*invstd, mean(gy), mean(gy*xhat), 1/gamma, beta
    const int q = threadIdx.x & 3;
    float cq[5][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        cq[0][e] = s.ch[2][4 * q + e];
        cq[1][e] = k1[4 * q + e];
        cq[2][e] = s.ch[3][4 * q + e];
        cq[3][e] = s.ch[4][4 * q + e];
        cq[4][e] = s.ch[5][4 * q + e];
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *dz = park + warp * (CB_SEG * CB_DS);
    // MMA roles (m16n8k8: M = channel, N = tap, K = voxel): gq = lane / 4, t = lane % 4.
    //   A (dz^T)  a0 (ch gq, vox t)  a1 (ch gq+8, vox t)  a2 (ch gq, vox t+4)  a3 (ch gq+8, vox t+4)
    //   B (xcol)  b0 (vox t, tap gq + 8 nb)  b1 (vox t+4, tap gq + 8 nb);   tap 27 = row of ones (bias), > 27 = zeros
    //   C         c0 (ch gq, tap 8nb+2t)  c1 (ch gq, tap 8nb+2t+1)  c2 (ch gq+8, tap 8nb+2t)  c3 (ch gq+8, tap 8nb+2t+1)
    const int gq = lane >> 2, t = lane & 3;
    int toff[4];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
        const int tap = gq + 8 * nb;
        toff[nb] = (tap < 27 ? (tap / 3) * CB_RS + tap % 3 : (tap == 27 ? 9 * CB_RS : 10 * CB_RS)) + t;
    }
    float acc[4][4];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nb][e] = 0.f;
    cb_for_each_segment(g, x, s, [&](const float *tile, const CbSeg &sg) {   // __syncwarp after each body: parked rows consumed
        const int x0 = sg.x0;
        if (from_y) {
            // lane = (voxel, channel quad): coalesced reads of gy, y and the mask; the masked gradient goes straight to its
            // parked row
#pragma unroll
            for (int i0 = 0; i0 < CB_SEG / 8; i0 += 4) {
                CbItem it[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) cb_item_load(gr, g, sg, (i0 + i) * 8 + (lane >> 2), q, it[i]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float gv[4], r[4];
                    cb_item_grad(it[i], gv);
                    const float yy4[4] = {it[i].y.x, it[i].y.y, it[i].y.z, it[i].y.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float xhat = (yy4[e] - cq[4][e]) * cq[3][e];
                        const float da = cq[0][e] * (gv[e] - cq[1][e] - xhat * cq[2][e]);
                        r[e] = ((it[i].mask >> e) & 1u) ? da : 0.f;      // mask is 0 outside the volume
                    }
                    *reinterpret_cast<float4 *>(dz + ((i0 + i) * 8 + (lane >> 2)) * CB_DS + 4 * q) = make_float4(r[0], r[1], r[2], r[3]);
                }
            }
        } else {
            float4 gv4[2][CB_CO / 4];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int c = 0; c < CB_CO / 4; ++c) gv4[j][c] = cb_load_grad(gr, g, sg, x0 + lane + 32 * j, c);
            float a[2][CB_CO];
            cb_conv_relu(s, tile, lane, a);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const bool ok = x0 + lane + 32 * j < g.W;
#pragma unroll
                for (int c = 0; c < CB_CO / 4; ++c) {
                    const float gv[4] = {gv4[j][c].x, gv4[j][c].y, gv4[j][c].z, gv4[j][c].w};
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int ch = 4 * c + e;
                        const float xhat = (a[j][ch] - s.ch[0][ch]) * s.ch[1][ch];
                        const float da = s.ch[2][ch] * (gv[e] - k1[ch] - xhat * s.ch[3][ch]);
                        o[e] = (ok && a[j][ch] > 0.f) ? da : 0.f;
                    }
                    *reinterpret_cast<float4 *>(dz + (lane + 32 * j) * CB_DS + 4 * c) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (int kb = 0; kb < CB_SEG / 8; ++kb) {
            const float *dr = dz + (8 * kb + t) * CB_DS + gq;
            const float af[4] = {dr[0], dr[8], dr[4 * CB_DS], dr[4 * CB_DS + 8]};
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                hi[e] = to_tf32(af[e]);
                lo[e] = to_tf32(af[e] - __uint_as_float(hi[e]));
            }
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                const uint32_t b0 = to_tf32(tile[toff[nb] + 8 * kb]), b1 = to_tf32(tile[toff[nb] + 8 * kb + 4]);
                mma_tf32(acc[nb], hi, b0, b1);
                mma_tf32(acc[nb], lo, b0, b1);
            }
        }
    });
    // block reduction over the warps through the (now free) tile area, in two halves of 4 warps
    // (4 warps x 28 taps x 16 channels = 1792 floats of the tile area)
    __syncthreads();
    float *red = &s.tile[0][0][0];
    for (int half = 0; half < 2; ++half) {
        if ((warp >> 2) == half) {
#pragma unroll
            for (int nb = 0; nb < 4; ++nb)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int tap = 8 * nb + 2 * t + (e & 1), chn = gq + 8 * (e >> 1);
                    if (tap < 28) red[((warp & 3) * 28 + tap) * CB_CO + chn] = acc[nb][e];
                }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 28 * CB_CO; i += blockDim.x) {
            float tsum = 0.f;
            for (int ww = 0; ww < 4; ++ww) tsum += red[ww * 28 * CB_CO + i];
            float *dst = partial + (int64_t)blockIdx.x * 28 * CB_CO + i;   // [tap][co]
            *dst = half == 0 ? tsum : *dst + tsum;
        }
        __syncthreads();
    }
}

// gw[co][tap] / gb[co] = sum over blocks of partial[block][tap][co] (fixed order)
__global__ void cb_wgrad_reduce_kernel(const float *__restrict__ partial, int nblocks, float *__restrict__ gw, float *__restrict__ gb) {
    // block = 32 outputs x 8 block-slices (coalesced 128-byte rows), then a fixed-order sum of the 8 slices
    __shared__ float sl[8][32];
    const int col = threadIdx.x & 31, part = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + col;
    float v = 0.f;
    if (i < 28 * CB_CO)
        for (int b = part; b < nblocks; b += 8) v += partial[(int64_t)b * 28 * CB_CO + i];
    sl[part][col] = v;
    __syncthreads();
    if (part != 0 || i >= 28 * CB_CO) return;
    v = 0.f;
    for (int p2 = 0; p2 < 8; ++p2) v += sl[p2][col];
    const int tap = i / CB_CO, co = i - tap * CB_CO;
    if (tap < 27)
        gw[co * 27 + tap] = v;
    else
        gb[co] = v;
}

static bool cb_geom(CbGeom &g, int B, int D, int H, int W) {
    g.B = B; g.D = D; g.H = H; g.W = W;
    g.segs_per_row = ceil_div(W, CB_SEG);
    const int64_t n = (int64_t)B * D * H * g.segs_per_row;
    g.n_vox = (int64_t)B * D * H * W;
    if (n >= ((int64_t)1 << 31)) return false;
    g.n_segs = (int)n;
    auto log2i = [](int v) {
        int l = 0;
        while ((1 << l) < v) ++l;
        return l;
    };
    auto is_pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    g.pow2 = is_pow2(g.segs_per_row) && is_pow2(H) && is_pow2(D) ? 1 : 0;
    g.sh_spr = log2i(g.segs_per_row);
    g.sh_h = log2i(H);
    g.sh_d = log2i(D);
    return true;
}

constexpr size_t CB_SMEM = sizeof(CbShared);
constexpr size_t CB_SMEM_WGRAD = sizeof(CbShared) + (size_t)CB_WARPS * CB_SEG * CB_DS * sizeof(float);

static int cb_attrs() {
    static DeviceOnce once;
    int dev;
    if (once.needed(dev)) {
        SVR_CUDA(cudaFuncSetAttribute(cb_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CB_SMEM));
        SVR_CUDA(cudaFuncSetAttribute(cb_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CB_SMEM));
        SVR_CUDA(cudaFuncSetAttribute(cb_bwd_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CB_SMEM));
        SVR_CUDA(cudaFuncSetAttribute(cb_bwd_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CB_SMEM_WGRAD));
        once.done(dev);
    }
    return 0;
}

static int cb_grid(const CbGeom &g) {
    const int64_t want = ceil_div<int64_t>(g.n_segs, CB_WARPS);
    const int64_t cap = (int64_t)sm_count() * 8;
    return (int)(want < cap ? want : cap);
}

}  // namespace svr

using namespace svr;

extern "C" {

size_t svr_conv1_bn_workspace_bytes(void) { return (size_t)sm_count() * 8 * (28 * CB_CO + 32) * sizeof(float) + 256; }

int svr_conv1_relu_bn_stats(const float *x, const float *w, const float *bias, int B, int D, int H, int W, int Co, float eps, float momentum,
                            float *running_mean, float *running_var, float *mean, float *invstd, void *workspace, size_t workspace_bytes,
                            void *stream) {
    SVR_REQUIRE(x && w && mean && invstd && workspace, "conv1_relu_bn_stats: null pointer");
    SVR_REQUIRE(Co == CB_CO, "conv1_relu_bn: 16 output channels supported (got %d)", Co);
    SVR_REQUIRE(workspace_bytes >= svr_conv1_bn_workspace_bytes(), "conv1_relu_bn_stats: workspace too small");
    CbGeom g;
    SVR_REQUIRE(cb_geom(g, B, D, H, W) && g.n_vox > 0, "conv1_relu_bn_stats: empty or too large grid");
    const int grid = cb_grid(g);
    if (int rc = cb_attrs()) return rc;
    cb_stats_kernel<<<grid, CB_WARPS * 32, CB_SMEM, as_stream(stream)>>>(x, w, bias, g, (float *)workspace);
    cb_stats_finalize_kernel<<<1, 1024, 0, as_stream(stream)>>>((const float *)workspace, grid, (double)g.n_vox, eps, momentum, running_mean,
                                                              running_var, mean, invstd);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_conv1_relu_bn_apply(const float *x, const float *w, const float *bias, const float *mean, const float *invstd, const float *gamma,
                            const float *beta, int B, int D, int H, int W, int Co, float *y, uint16_t *y_bf16, uint16_t *relu_mask,
                            void *stream) {
    SVR_REQUIRE(x && w && mean && invstd && y, "conv1_relu_bn_apply: null pointer");
    SVR_REQUIRE(Co == CB_CO, "conv1_relu_bn: 16 output channels supported (got %d)", Co);
    CbGeom g;
    SVR_REQUIRE(cb_geom(g, B, D, H, W), "conv1_relu_bn_apply: grid too large");
    if (g.n_vox == 0) return 0;
    if (int rc = cb_attrs()) return rc;
    cb_apply_kernel<<<cb_grid(g), CB_WARPS * 32, CB_SMEM, as_stream(stream)>>>(x, w, bias, mean, invstd, gamma, beta, g, y, reinterpret_cast<uint4 *>(y_bf16), relu_mask);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_conv1_relu_bn_bwd(const float *x, const float *w, const float *bias, const float *mean, const float *invstd, const float *gamma,
                          const float *beta, const float *y, const uint16_t *relu_mask, const float *gy, const float *g_pooled,
                          const uint32_t *pool_idx, int B, int D, int H, int W, int Co, float *gw, float *gb, float *ggamma, float *gbeta,
                          void *workspace, size_t workspace_bytes, void *stream) {
    SVR_REQUIRE(x && w && mean && invstd && gw && gb && ggamma && gbeta && workspace, "conv1_relu_bn_bwd: null pointer");
    SVR_REQUIRE((g_pooled == nullptr) == (pool_idx == nullptr), "conv1_relu_bn_bwd: g_pooled and pool_idx go together");
    SVR_REQUIRE((y == nullptr) == (relu_mask == nullptr), "conv1_relu_bn_bwd: y and relu_mask go together");
    const CbGrad gr{gy, g_pooled, pool_idx, y, relu_mask, gamma, beta};
    SVR_REQUIRE(Co == CB_CO, "conv1_relu_bn: 16 output channels supported (got %d)", Co);
    SVR_REQUIRE(workspace_bytes >= svr_conv1_bn_workspace_bytes(), "conv1_relu_bn_bwd: workspace too small");
    CbGeom g;
    SVR_REQUIRE(cb_geom(g, B, D, H, W) && g.n_vox > 0, "conv1_relu_bn_bwd: empty or too large grid");
    const int grid = cb_grid(g);
    float *p32 = (float *)workspace, *p448 = p32 + (size_t)sm_count() * 8 * 32;
    cudaStream_t st = as_stream(stream);
    if (int rc = cb_attrs()) return rc;
    cb_bwd_reduce_kernel<<<grid, CB_WARPS * 32, CB_SMEM, st>>>(x, w, bias, mean, invstd, gr, g, p32);
    cb_bwd_finalize_kernel<<<1, 1024, 0, st>>>(p32, grid, gbeta, ggamma);
    cb_bwd_wgrad_kernel<<<grid, CB_WARPS * 32, CB_SMEM_WGRAD, st>>>(x, w, bias, mean, invstd, gamma, gbeta, ggamma, gr, g, (float)(1.0 / (double)g.n_vox), p448);
    cb_wgrad_reduce_kernel<<<ceil_div(28 * CB_CO, 32), 256, 0, st>>>(p448, grid, gw, gb);
    SVR_LAUNCH_CHECK();
    return 0;
}
}
