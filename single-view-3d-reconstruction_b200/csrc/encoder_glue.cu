// Elementwise glue of the encoder's Conv3d -> ReLU layers (model/ifnet.py:127-135,165-183: `self.actvn(self.conv_x(net))`),
// channels-last fp32.  torch runs this as separate passes over the activation (bias add, clamp, threshold_backward,
// a strided bias-gradient reduction, dtype casts); each is a pure HBM stream, so they are fused into one forward and
// one backward pass:
//   forward : y <- max(y + bias, 0) in place on the bias-free convolution output
//   backward: g = gy * [y > 0], written as bf16 and/or fp32 for the convolution backward kernels, and the bias
//             gradient sum_v g[v][c] from the same read (deterministic two-level sum)
#include "common.cuh"

namespace svr {

// torch's ReLU (clamp_min / threshold) propagates NaN; fmaxf would turn it into 0 and hide a divergence
__device__ __forceinline__ float relu_nan(float v) { return v > 0.f ? v : (v != v ? v : 0.f); }

__global__ void __launch_bounds__(256) bias_relu_kernel(float4 *__restrict__ y, const float *__restrict__ bias, int c4, int64_t n4) {
    extern __shared__ float4 b_s[];
    for (int i = threadIdx.x; i < c4; i += blockDim.x) b_s[i] = bias ? reinterpret_cast<const float4 *>(bias)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // stride is a multiple of c4 (host guarantees blockDim % c4 == 0), so a thread keeps its channel quad
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const float4 b = b_s[i0 % c4];
    for (int64_t i = i0; i < n4; i += stride) {
        float4 v = y[i];
        v.x = relu_nan(v.x + b.x);
        v.y = relu_nan(v.y + b.y);
        v.z = relu_nan(v.z + b.z);
        v.w = relu_nan(v.w + b.w);
        y[i] = v;
    }
}

// thread = (row lane, channel quad); rows strided over the grid; per-block column sums to partial[block][C]
__global__ void __launch_bounds__(256) relu_bwd_kernel(const float4 *__restrict__ gy, const float4 *__restrict__ y, int c4, int64_t rows,
                                                       float4 *__restrict__ g_f32, uint2 *__restrict__ g_bf16, float4 *__restrict__ partial) {
    __shared__ float4 red[256];
    const int cq = threadIdx.x % c4, rl = threadIdx.x / c4, rpb = blockDim.x / c4;   // rows per block step
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = (int64_t)blockIdx.x * rpb + rl; r < rows; r += (int64_t)gridDim.x * rpb) {
        const int64_t i = r * c4 + cq;
        const float4 gv = __ldg(gy + i), yv = __ldg(y + i);
        float4 o;
        o.x = yv.x > 0.f ? gv.x : 0.f;
        o.y = yv.y > 0.f ? gv.y : 0.f;
        o.z = yv.z > 0.f ? gv.z : 0.f;
        o.w = yv.w > 0.f ? gv.w : 0.f;
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        if (g_f32) g_f32[i] = o;
        if (g_bf16) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
            g_bf16[i] = make_uint2(*reinterpret_cast<uint32_t *>(&lo), *reinterpret_cast<uint32_t *>(&hi));
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (rl == 0) {
        for (int k = 1; k < rpb; ++k) {
            const float4 t = red[k * c4 + cq];
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        partial[(int64_t)blockIdx.x * c4 + cq] = acc;
    }
}

// gb[c] = sum over blocks of partial[block][c]; one warp per 32 channels x 8 block slices
__global__ void relu_bwd_reduce_kernel(const float *__restrict__ partial, int nblocks, int C, float *__restrict__ gb) {
    __shared__ float sl[32][33];
    const int col = threadIdx.x & 31, part = threadIdx.x >> 5;   // 1024 threads: 32 block slices
    const int c = blockIdx.x * 32 + col;
    float v = 0.f;
    if (c < C)
        for (int b = part; b < nblocks; b += 32) v += partial[(int64_t)b * C + c];
    sl[part][col] = v;
    __syncthreads();
    if (part == 0 && c < C) {
        v = 0.f;
        for (int k = 0; k < 32; ++k) v += sl[k][col];
        gb[c] = v;
    }
}

// bf16 -> fp32 of a dense buffer (cuDNN's bf16 backward-data result feeding an fp32 gradient chain)
__global__ void __launch_bounds__(256) widen_bf16_kernel(const uint4 *__restrict__ src, int64_t n8, float4 *__restrict__ dst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(src + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        float f[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            f[2 * k] = __uint_as_float(w[k] << 16);
            f[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
        }
        dst[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
        dst[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
    }
}

}  // namespace svr

using namespace svr;

extern "C" {

int svr_bias_relu_cl(float *y, const float *bias, int64_t rows, int C, void *stream) {
    SVR_REQUIRE(C > 0 && C % 4 == 0 && 256 % (C / 4) == 0, "bias_relu_cl: channel count must be 4, 8, ..., 1024 with 256 %% (C/4) == 0 (got %d)", C);
    SVR_REQUIRE(rows >= 0, "bias_relu_cl: negative row count");
    const int64_t n4 = rows * (C / 4);
    if (n4 == 0) return 0;
    SVR_REQUIRE(y, "bias_relu_cl: null pointer");
    int64_t blocks = ceil_div<int64_t>(n4, 256 * 4);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    bias_relu_kernel<<<(unsigned)blocks, 256, (C / 4) * sizeof(float4), as_stream(stream)>>>(reinterpret_cast<float4 *>(y), bias, C / 4, n4);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_widen_bf16(const uint16_t *src, int64_t n, float *dst, void *stream) {
    SVR_REQUIRE(n >= 0 && n % 8 == 0, "widen_bf16: element count must be a non-negative multiple of 8");
    if (n == 0) return 0;
    SVR_REQUIRE(src && dst, "widen_bf16: null pointer");
    SVR_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, "widen_bf16: buffers must be 16-byte aligned");
    int64_t blocks = ceil_div<int64_t>(n / 8, 256 * 2);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    widen_bf16_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4 *>(src), n / 8, reinterpret_cast<float4 *>(dst));
    SVR_LAUNCH_CHECK();
    return 0;
}

size_t svr_relu_bwd_cl_workspace_bytes(int C) { return (size_t)sm_count() * 8 * C * sizeof(float) + 256; }

int svr_relu_bwd_cl(const float *gy, const float *y, int64_t rows, int C, float *g_f32, uint16_t *g_bf16, float *gbias, void *workspace,
                    size_t workspace_bytes, void *stream) {
    SVR_REQUIRE(C > 0 && C % 4 == 0 && 256 % (C / 4) == 0, "relu_bwd_cl: channel count must be 4, 8, ..., 1024 with 256 %% (C/4) == 0 (got %d)", C);
    SVR_REQUIRE(rows >= 0, "relu_bwd_cl: negative row count");
    if (rows == 0) {
        if (gbias) SVR_CUDA(cudaMemsetAsync(gbias, 0, C * sizeof(float), as_stream(stream)));
        return 0;
    }
    SVR_REQUIRE(gy && y && workspace && (g_f32 || g_bf16), "relu_bwd_cl: null pointer");
    SVR_REQUIRE(workspace_bytes >= svr_relu_bwd_cl_workspace_bytes(C), "relu_bwd_cl: workspace too small");
    const int c4 = C / 4, rpb = 256 / c4;
    int64_t blocks = ceil_div<int64_t>(rows, rpb * 4);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    relu_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4 *>(gy), reinterpret_cast<const float4 *>(y), c4, rows,
                                                                       reinterpret_cast<float4 *>(g_f32), reinterpret_cast<uint2 *>(g_bf16),
                                                                       reinterpret_cast<float4 *>(workspace));
    if (gbias) relu_bwd_reduce_kernel<<<ceil_div(C, 32), 1024, 0, as_stream(stream)>>>((const float *)workspace, (int)blocks, C, gbias);
    SVR_LAUNCH_CHECK();
    return 0;
}
}
