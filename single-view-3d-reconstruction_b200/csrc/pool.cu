// Channels-last (NDHWC) 2x2x2 max pooling, forward and backward, for the IF-Net encoder
// (nn.MaxPool3d(2) of model/ifnet.py:133,169-190, reference root).  torch's CUDA max_pool3d has no
// channels-last path: on an NDHWC activation it first makes an NCDHW copy (a 537 MB strided copy for
// the level-1 volume at batch 4) and hands NCDHW to the next cuDNN conv, which transposes it back.
// This kernel keeps the whole encoder channels-last.  Semantics are torch's: floor output size, the
// FIRST maximum in (d,h,w) scan order wins, NaN propagates.
#include "common.cuh"

namespace svr {

__global__ void __launch_bounds__(256) maxpool_cl_fwd_kernel(const float4 *__restrict__ in, int D, int H, int W, int C4, int64_t n_out,
                                                             float4 *__restrict__ out, uint32_t *__restrict__ idx) {
    const int Do = D / 2, Ho = H / 2, Wo = W / 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        int64_t v = i / C4;
        const int wo = (int)(v % Wo);
        v /= Wo;
        const int ho = (int)(v % Ho);
        v /= Ho;
        const int dq = (int)(v % Do);
        const int64_t b = v / Do;
        float4 best = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t arg = 0;   // 4 x 3-bit winners packed in bytes
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int dz = k >> 2, dy = (k >> 1) & 1, dx = k & 1;
            const float4 x = __ldg(in + ((((b * D + (2 * dq + dz)) * H + (2 * ho + dy)) * W + (2 * wo + dx)) * C4 + c));
            if (k == 0) {
                best = x;
            } else {
                if (x.x > best.x || x.x != x.x) { if (!(best.x != best.x)) { best.x = x.x; arg = (arg & ~0xffu) | (uint32_t)k; } }
                if (x.y > best.y || x.y != x.y) { if (!(best.y != best.y)) { best.y = x.y; arg = (arg & ~0xff00u) | ((uint32_t)k << 8); } }
                if (x.z > best.z || x.z != x.z) { if (!(best.z != best.z)) { best.z = x.z; arg = (arg & ~0xff0000u) | ((uint32_t)k << 16); } }
                if (x.w > best.w || x.w != x.w) { if (!(best.w != best.w)) { best.w = x.w; arg = (arg & ~0xff000000u) | ((uint32_t)k << 24); } }
            }
        }
        out[i] = best;
        idx[i] = arg;
    }
}

__global__ void __launch_bounds__(256) maxpool_cl_bwd_kernel(const float4 *__restrict__ gout, const uint32_t *__restrict__ idx, int D, int H,
                                                             int W, int C4, int64_t n_out, float4 *__restrict__ gin) {
    const int Do = D / 2, Ho = H / 2, Wo = W / 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        int64_t v = i / C4;
        const int wo = (int)(v % Wo);
        v /= Wo;
        const int ho = (int)(v % Ho);
        v /= Ho;
        const int dq = (int)(v % Do);
        const int64_t b = v / Do;
        const float4 g = __ldg(gout + i);
        const uint32_t arg = idx[i];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int dz = k >> 2, dy = (k >> 1) & 1, dx = k & 1;
            float4 o;
            o.x = ((arg & 0xff) == (uint32_t)k) ? g.x : 0.f;
            o.y = (((arg >> 8) & 0xff) == (uint32_t)k) ? g.y : 0.f;
            o.z = (((arg >> 16) & 0xff) == (uint32_t)k) ? g.z : 0.f;
            o.w = ((arg >> 24) == (uint32_t)k) ? g.w : 0.f;
            gin[(((b * D + (2 * dq + dz)) * H + (2 * ho + dy)) * W + (2 * wo + dx)) * C4 + c] = o;
        }
    }
}

}  // namespace svr

using namespace svr;

extern "C" {

int svr_maxpool2_cl_fwd(const float *in, int B, int D, int H, int W, int C, float *out, uint32_t *idx, void *stream) {
    SVR_REQUIRE(in && out && idx, "maxpool: null pointer");
    SVR_REQUIRE(C % 4 == 0 && D >= 2 && H >= 2 && W >= 2, "maxpool: C %% 4 == 0 and every spatial dim >= 2 required");
    int64_t n_out = (int64_t)B * (D / 2) * (H / 2) * (W / 2) * (C / 4);
    if (n_out == 0) return 0;
    int64_t blocks = ceil_div<int64_t>(n_out, 256);
    int64_t cap = (int64_t)sm_count() * 32;
    maxpool_cl_fwd_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>((const float4 *)in, D, H, W, C / 4, n_out,
                                                                                               (float4 *)out, idx);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_maxpool2_cl_bwd(const float *gout, const uint32_t *idx, int B, int D, int H, int W, int C, float *gin, void *stream) {
    SVR_REQUIRE(gout && gin && idx, "maxpool_bwd: null pointer");
    SVR_REQUIRE(C % 4 == 0 && D >= 2 && H >= 2 && W >= 2, "maxpool_bwd: C %% 4 == 0 and every spatial dim >= 2 required");
    cudaStream_t st = as_stream(stream);
    if ((D | H | W) & 1)   // odd extents: the last plane/row/column receives no gradient
        SVR_CUDA(cudaMemsetAsync(gin, 0, (size_t)B * D * H * W * C * sizeof(float), st));
    int64_t n_out = (int64_t)B * (D / 2) * (H / 2) * (W / 2) * (C / 4);
    if (n_out == 0) return 0;
    int64_t blocks = ceil_div<int64_t>(n_out, 256);
    int64_t cap = (int64_t)sm_count() * 32;
    maxpool_cl_bwd_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>((const float4 *)gout, idx, D, H, W, C / 4, n_out,
                                                                                 (float4 *)gin);
    SVR_LAUNCH_CHECK();
    return 0;
}
}
