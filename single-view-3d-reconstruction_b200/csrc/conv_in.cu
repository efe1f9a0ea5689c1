// First encoder convolution of IF-Net fused with its ReLU: Conv3d(1 -> Co, 3x3x3, padding 1) + ReLU
// (model/ifnet.py:126,164 `self.actvn(self.conv_in(x))`; :68,100 for the 32-net), channels-last output.
//
// With a single input channel the layout of x is ambiguous and cuDNN runs this layer on its NCDHW path:
// 1.7 ms forward plus a 537 MB NDHWC->NCDHW copy of the gradient and an NCDHW wgrad (3.2 ms) per
// step at batch 4 -- 29 % of the whole training step for 0.2 % of its FLOPs.  The layer is a 27-tap
// stencil with 16 (32) outputs and purely bandwidth bound: this kernel reads x once (L1 reuse of the
// 27 neighbours) and writes the NDHWC activation once.  fp32 FMA arithmetic (cuDNN would use TF32).
#include "common.cuh"

namespace svr {

template <int CO>
__global__ void __launch_bounds__(256) conv1_relu_fwd_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                             const float *__restrict__ bias, int D, int H, int W, int64_t n_vox,
                                                             float *__restrict__ y) {
    __shared__ float4 wt[27][CO / 4];   // [tap][co]
    __shared__ float4 bs[CO / 4];
    for (int i = threadIdx.x; i < 27 * CO; i += 256) {
        int co = i / 27, tap = i - co * 27;
        reinterpret_cast<float *>(&wt[tap][0])[co] = w[i];
    }
    if (threadIdx.x < CO) reinterpret_cast<float *>(bs)[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n_vox; p += (int64_t)gridDim.x * 256) {
        const int xx = (int)(p % W), yy = (int)((p / W) % H), zz = (int)((p / ((int64_t)W * H)) % D);
        float4 acc[CO / 4];
#pragma unroll
        for (int c = 0; c < CO / 4; ++c) acc[c] = bs[c];
#pragma unroll
        for (int tap = 0; tap < 27; ++tap) {
            const int dz = tap / 9 - 1, dy = (tap / 3) % 3 - 1, dx = tap % 3 - 1;
            const int z = zz + dz, yq = yy + dy, xq = xx + dx;
            float v = 0.f;
            if (z >= 0 && z < D && yq >= 0 && yq < H && xq >= 0 && xq < W) v = __ldg(x + p + ((int64_t)dz * H + dy) * W + dx);
#pragma unroll
            for (int c = 0; c < CO / 4; ++c) {
                const float4 ww = wt[tap][c];
                acc[c].x = fmaf(v, ww.x, acc[c].x);
                acc[c].y = fmaf(v, ww.y, acc[c].y);
                acc[c].z = fmaf(v, ww.z, acc[c].z);
                acc[c].w = fmaf(v, ww.w, acc[c].w);
            }
        }
        float4 *dst = reinterpret_cast<float4 *>(y + p * CO);
#pragma unroll
        for (int c = 0; c < CO / 4; ++c)
            dst[c] = make_float4(fmaxf(acc[c].x, 0.f), fmaxf(acc[c].y, 0.f), fmaxf(acc[c].z, 0.f), fmaxf(acc[c].w, 0.f));
    }
}

// weight/bias gradient partials.  Work unit = one x-row segment of up to 128 voxels.  Per segment the
// masked gradient gz = gy * [y > 0] (128 x CO) and the 3x3 neighbouring x-rows (with a 1-voxel halo)
// are staged in shared memory with coalesced loads; thread (tap, channel quad) then accumulates its 4
// weights over the segment from shared memory only.  Tap 27 is the bias row (x == 1).
constexpr int C1_TV = 128;

template <int CO>
__global__ void __launch_bounds__(256) conv1_relu_wgrad_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                                               const float *__restrict__ gy, int B, int D, int H, int W,
                                                               int64_t n_seg, int64_t seg_per_block, float *__restrict__ partial) {
    constexpr int CQ = CO / 4;
    constexpr int NT = 28 * CQ;                 // (tap, channel quad) workers: 112 (CO=16) / 224 (CO=32)
    constexpr int SETS = 256 / NT;              // voxel sub-ranges processed concurrently: 2 / 1
    __shared__ float4 gz_s[C1_TV][CQ];
    __shared__ float halo[9][C1_TV + 2];        // rows (dz,dy) in {-1,0,1}^2, columns x0-1 .. x0+nv
    __shared__ float4 red[256];
    const int tid = threadIdx.x;
    const int set = tid / NT, t = tid - set * NT;
    const bool worker = set < SETS;
    const int tap = worker ? t / CQ : 0, cq = worker ? t - (t / CQ) * CQ : 0;
    const int trow = tap < 27 ? tap / 3 : 0, tdx = tap < 27 ? tap % 3 : 0;
    const int segs_per_row = (W + C1_TV - 1) / C1_TV;
    const int64_t s0 = (int64_t)blockIdx.x * seg_per_block;
    int64_t s1 = s0 + seg_per_block;
    if (s1 > n_seg) s1 = n_seg;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t sg = s0; sg < s1; ++sg) {
        const int xs0 = (int)(sg % segs_per_row) * C1_TV;
        int64_t r = sg / segs_per_row;                       // row index over (b, z, y)
        const int yy = (int)(r % H);
        r /= H;
        const int zz = (int)(r % D);
        const int64_t b = r / D;
        const int nv = (W - xs0) < C1_TV ? (W - xs0) : C1_TV;
        const int64_t base = ((b * D + zz) * H + yy) * (int64_t)W + xs0;
        for (int i = tid; i < C1_TV * CQ; i += 256) {        // stage gz
            const int v = i / CQ, c = i - v * CQ;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < nv) {
                g = __ldg(reinterpret_cast<const float4 *>(gy + (base + v) * CO) + c);
                const float4 yv = __ldg(reinterpret_cast<const float4 *>(y + (base + v) * CO) + c);
                g.x = yv.x > 0.f ? g.x : 0.f;
                g.y = yv.y > 0.f ? g.y : 0.f;
                g.z = yv.z > 0.f ? g.z : 0.f;
                g.w = yv.w > 0.f ? g.w : 0.f;
            }
            gz_s[v][c] = g;
        }
        for (int i = tid; i < 9 * (C1_TV + 2); i += 256) {   // stage the 3x3 rows with halo
            const int row = i / (C1_TV + 2), col = i - row * (C1_TV + 2);
            const int z = zz + row / 3 - 1, yq = yy + row % 3 - 1, xq = xs0 + col - 1;
            float val = 0.f;
            if (col < nv + 2 && z >= 0 && z < D && yq >= 0 && yq < H && xq >= 0 && xq < W)
                val = __ldg(x + ((b * D + z) * H + yq) * (int64_t)W + xq);
            halo[row][col] = val;
        }
        __syncthreads();
        if (worker) {
            const int v0 = set * (C1_TV / SETS);
            int v1 = v0 + C1_TV / SETS;
            if (v1 > nv) v1 = nv;
#pragma unroll 4
            for (int v = v0; v < v1; ++v) {
                const float4 g = gz_s[v][cq];
                const float xv = tap < 27 ? halo[trow][v + tdx] : 1.f;
                acc.x = fmaf(g.x, xv, acc.x);
                acc.y = fmaf(g.y, xv, acc.y);
                acc.z = fmaf(g.z, xv, acc.z);
                acc.w = fmaf(g.w, xv, acc.w);
            }
        }
        __syncthreads();
    }
    red[tid] = acc;
    __syncthreads();
    if (worker && set == 0) {
        float4 o = acc;
#pragma unroll
        for (int s2 = 1; s2 < SETS; ++s2) {
            const float4 rr = red[s2 * NT + t];
            o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
        }
        reinterpret_cast<float4 *>(partial + (int64_t)blockIdx.x * 28 * CO + tap * CO)[cq] = o;   // partial[block][tap][co]
    }
}

// gw[co][tap] / gb[co] = sum over blocks (fixed order)
__global__ void conv1_wgrad_reduce_kernel(const float *__restrict__ partial, int nblocks, int CO, float *__restrict__ gw,
                                          float *__restrict__ gb) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;   // over 28*CO, layout [tap][co]
    if (i >= 28 * CO) return;
    float v = 0.f;
    for (int b = 0; b < nblocks; ++b) v += partial[(int64_t)b * 28 * CO + i];
    int tap = i / CO, co = i - tap * CO;
    if (tap < 27)
        gw[co * 27 + tap] = v;
    else
        gb[co] = v;
}

// d x (needed when the input grid requires grad: scene-net trainer)
template <int CO>
__global__ void __launch_bounds__(256) conv1_relu_dgrad_kernel(const float *__restrict__ y, const float *__restrict__ gy,
                                                               const float *__restrict__ w, int D, int H, int W, int64_t n_vox,
                                                               float *__restrict__ gx) {
    __shared__ float4 wt[27][CO / 4];
    for (int i = threadIdx.x; i < 27 * CO; i += 256) {
        int co = i / 27, tap = i - co * 27;
        reinterpret_cast<float *>(&wt[tap][0])[co] = w[i];
    }
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n_vox; p += (int64_t)gridDim.x * 256) {
        const int xx = (int)(p % W), yy = (int)((p / W) % H), zz = (int)((p / ((int64_t)W * H)) % D);
        float acc = 0.f;
#pragma unroll 1
        for (int tap = 0; tap < 27; ++tap) {
            const int dz = tap / 9 - 1, dy = (tap / 3) % 3 - 1, dx = tap % 3 - 1;
            const int z = zz - dz, yq = yy - dy, xq = xx - dx;   // output voxel that used x[p] with this tap
            if (z < 0 || z >= D || yq < 0 || yq >= H || xq < 0 || xq >= W) continue;
            const int64_t po = p - (((int64_t)dz * H + dy) * W + dx);
#pragma unroll
            for (int c = 0; c < CO / 4; ++c) {
                const float4 g = __ldg(reinterpret_cast<const float4 *>(gy + po * CO) + c);
                const float4 yv = __ldg(reinterpret_cast<const float4 *>(y + po * CO) + c);
                const float4 ww = wt[tap][c];
                acc = fmaf(yv.x > 0.f ? g.x : 0.f, ww.x, acc);
                acc = fmaf(yv.y > 0.f ? g.y : 0.f, ww.y, acc);
                acc = fmaf(yv.z > 0.f ? g.z : 0.f, ww.z, acc);
                acc = fmaf(yv.w > 0.f ? g.w : 0.f, ww.w, acc);
            }
        }
        gx[p] = acc;
    }
}

template <int CO>
static int conv1_bwd_impl(const float *x, const float *y, const float *gy, const float *w, int B, int D, int H, int W, float *gw,
                          float *gb, float *gx, float *partial, int nblocks, cudaStream_t st) {
    const int64_t n_vox = (int64_t)B * D * H * W;
    const int64_t n_seg = (int64_t)B * D * H * ceil_div(W, C1_TV);
    const int64_t per_block = ceil_div<int64_t>(n_seg, nblocks);
    conv1_relu_wgrad_kernel<CO><<<nblocks, 256, 0, st>>>(x, y, gy, B, D, H, W, n_seg, per_block, partial);
    conv1_wgrad_reduce_kernel<<<ceil_div(28 * CO, 128), 128, 0, st>>>(partial, nblocks, CO, gw, gb);
    if (gx) {
        int64_t blocks = ceil_div<int64_t>(n_vox, 256);
        int64_t cap = (int64_t)sm_count() * 16;
        conv1_relu_dgrad_kernel<CO><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(y, gy, w, D, H, W, n_vox, gx);
    }
    SVR_LAUNCH_CHECK();
    return 0;
}

}  // namespace svr

using namespace svr;

extern "C" {

int svr_conv1_relu_fwd(const float *x, const float *w, const float *bias, int B, int D, int H, int W, int Co, float *y, void *stream) {
    SVR_REQUIRE(x && w && y, "conv1_relu_fwd: null pointer");
    SVR_REQUIRE(Co == 16 || Co == 32, "conv1_relu_fwd: 16 or 32 output channels supported (got %d)", Co);
    const int64_t n_vox = (int64_t)B * D * H * W;
    if (n_vox == 0) return 0;
    int64_t blocks = ceil_div<int64_t>(n_vox, 256);
    int64_t cap = (int64_t)sm_count() * 16;
    const unsigned g = (unsigned)(blocks < cap ? blocks : cap);
    if (Co == 16)
        conv1_relu_fwd_kernel<16><<<g, 256, 0, as_stream(stream)>>>(x, w, bias, D, H, W, n_vox, y);
    else
        conv1_relu_fwd_kernel<32><<<g, 256, 0, as_stream(stream)>>>(x, w, bias, D, H, W, n_vox, y);
    SVR_LAUNCH_CHECK();
    return 0;
}

size_t svr_conv1_relu_bwd_workspace_bytes(int Co) { return (size_t)sm_count() * 8 * 28 * Co * sizeof(float) + 256; }

int svr_conv1_relu_bwd(const float *x, const float *y, const float *gy, const float *w, int B, int D, int H, int W, int Co, float *gw,
                       float *gb, float *gx, void *workspace, size_t workspace_bytes, void *stream) {
    SVR_REQUIRE(x && y && gy && w && gw && gb && workspace, "conv1_relu_bwd: null pointer");
    SVR_REQUIRE(Co == 16 || Co == 32, "conv1_relu_bwd: 16 or 32 output channels supported (got %d)", Co);
    SVR_REQUIRE(workspace_bytes >= svr_conv1_relu_bwd_workspace_bytes(Co), "conv1_relu_bwd: workspace too small");
    const int nblocks = sm_count() * 8;
    if ((int64_t)B * D * H * W == 0) return 0;
    if (Co == 16) return conv1_bwd_impl<16>(x, y, gy, w, B, D, H, W, gw, gb, gx, (float *)workspace, nblocks, as_stream(stream));
    return conv1_bwd_impl<32>(x, y, gy, w, B, D, H, W, gw, gb, gx, (float *)workspace, nblocks, as_stream(stream));
}
}
