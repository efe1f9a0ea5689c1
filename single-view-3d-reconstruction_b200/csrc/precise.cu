// fp32-accurate tier of the IF-Net query path (`configure(precision=32)`): logits within 1e-3 of the fp32 reference
// (model/ifnet.py:38-61,155-199 run in fp32), gradients within 1e-3 relative L2.
//
// The reference's arithmetic is fp32 end to end.  This tier keeps every tensor fp32 in HBM (volumes are read as the
// encoder wrote them: fp32 NDHWC, no bf16 copy; features, hidden activations and all gradients fp32) and still runs
// the contractions on the tensor cores: every fp32 operand X is split into two bf16 terms X = hi + lo
// (hi = bf16(X), lo = bf16(X - hi), |X - hi - lo| <= 2^-17 |X|) and a product is accumulated as
//        A.B  ~=  A_hi.B_hi + A_lo.B_hi + A_hi.B_lo        (fp32 accumulation in TMEM, dropped term ~2^-16)
// by three passes of the bf16 tcgen05 GEMMs of gemm.cu over the same fp32 output (`flags & 32`: accumulate).
// The kernels here are the CUDA-core pieces around those GEMMs: the stencil gather / scatter on fp32 volumes,
// the hi/lo split, and the small reductions of the decoder backward in fp32.
#include "common.cuh"
#include "sampling.cuh"

namespace svr {

struct VolPtrsF {
    const float *v[SVR_MAX_LEVELS];
};
struct GradPtrsF {
    float *g[SVR_MAX_LEVELS];
};

// X -> (hi, lo) bf16, 8 elements per thread
__global__ void __launch_bounds__(256) split_bf16_kernel(const float *__restrict__ x, int64_t n8, __nv_bfloat16 *__restrict__ hi,
                                                         __nv_bfloat16 *__restrict__ lo) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(x) + 2 * i);
        const float4 b = __ldg(reinterpret_cast<const float4 *>(x) + 2 * i + 1);
        const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        float h[8], l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            h[j] = __bfloat162float(__float2bfloat16_rn(f[j]));
            l[j] = f[j] - h[j];      // exact in fp32
        }
        reinterpret_cast<uint4 *>(hi)[i] = float8_to_bf16(h);
        reinterpret_cast<uint4 *>(lo)[i] = float8_to_bf16(l);
    }
}

// one feature unit (8 fp32 values) of one point from fp32 channel-last volumes; same index arithmetic as the bf16 path
__device__ __forceinline__ void gather_unit_f32(const Pyr &P, int u, float px, float py, float pz, const float *__restrict__ x0_b,
                                                const float *const *vol_b, float (&acc)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    int level, d, c0;
    if (!decode_unit(P, u, level, d, c0)) return;
    if (level == 0) {
#pragma unroll
        for (int dd = 0; dd < 7; ++dd) acc[dd] = level0_sample(P, dd, px, py, pz, x0_b);
        return;
    }
    Corners c;
    stencil_corners(P, level, d, px, py, pz, c);
    const float *vb = vol_b[level] + c0;
    const int C = P.C[level];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int aa = k & 1, b = (k >> 1) & 1, e = k >> 2;
        const int x = c.x0 + aa, y = c.y0 + b, z = c.z0 + e;
        if (!corner_in(P, level, x, y, z)) continue;
        const float w = c.wx[aa] * c.wy[b] * c.wz[e];
        const float4 *src = reinterpret_cast<const float4 *>(vb + (((int64_t)z * P.H[level] + y) * P.W[level] + x) * C);
        const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
        acc[0] = fmaf(v0.x, w, acc[0]);
        acc[1] = fmaf(v0.y, w, acc[1]);
        acc[2] = fmaf(v0.z, w, acc[2]);
        acc[3] = fmaf(v0.w, w, acc[3]);
        acc[4] = fmaf(v1.x, w, acc[4]);
        acc[5] = fmaf(v1.y, w, acc[5]);
        acc[6] = fmaf(v1.z, w, acc[6]);
        acc[7] = fmaf(v1.w, w, acc[7]);
    }
}

__global__ void __launch_bounds__(256) gather_fwd_f32_kernel(const float *__restrict__ points, int N, int64_t total_pts,
                                                             const float *__restrict__ x0, VolPtrsF vols, Pyr P, float *__restrict__ feat) {
    const int UP = P.kp / 8;
    const int64_t total = total_pts * UP;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pt = t / UP;
        const int u = (int)(t - pt * UP);
        const int b = (int)(pt / N);
        const float px = points[pt * 3 + 0], py = points[pt * 3 + 1], pz = points[pt * 3 + 2];
        const float *vb[SVR_MAX_LEVELS];
#pragma unroll
        for (int l = 1; l < SVR_MAX_LEVELS; ++l)
            vb[l] = l < P.n_levels ? vols.v[l] + (int64_t)b * P.D[l] * P.H[l] * P.W[l] * P.C[l] : nullptr;
        vb[0] = nullptr;
        float acc[8];
        gather_unit_f32(P, u, px, py, pz, x0 + (int64_t)b * P.D[0] * P.H[0] * P.W[0], vb, acc);
        float4 *dst = reinterpret_cast<float4 *>(feat + pt * P.kp + (int64_t)u * 8);
        dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

__device__ __forceinline__ void red_add_v4f(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// scatter-add of fp32 d-features with fp32 trilinear weights (ATen grid_sampler_3d_backward semantics): one thread per
// (point, unit); 16-byte vector reductions into the fp32 channel-last gradient volumes; optional d(points).
template <bool NEED_DPTS>
__global__ void __launch_bounds__(256) gather_bwd_f32_kernel(const float *__restrict__ points, int N, int64_t total_pts,
                                                             const float *__restrict__ x0, VolPtrsF vols, Pyr P,
                                                             const float *__restrict__ dfeat, float *__restrict__ gx0, GradPtrsF gv,
                                                             float *__restrict__ gpoints) {
    const int UP = P.kp / 8;
    const int64_t total = total_pts * UP;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int64_t pt = t / UP;
    const int u = (int)(t - pt * UP);
    const int b = (int)(pt / N);
    int level, d, c0;
    if (!decode_unit(P, u, level, d, c0)) return;
    if (!NEED_DPTS && (level == 0 ? !gx0 : !gv.g[level])) return;
    const float px = points[pt * 3 + 0], py = points[pt * 3 + 1], pz = points[pt * 3 + 2];
    const float4 g0 = *reinterpret_cast<const float4 *>(dfeat + pt * P.kp + (int64_t)u * 8);
    const float4 g1 = *reinterpret_cast<const float4 *>(dfeat + pt * P.kp + (int64_t)u * 8 + 4);
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float dq[3] = {0.f, 0.f, 0.f};
    if (level == 0) {
        const int64_t base = (int64_t)b * P.D[0] * P.H[0] * P.W[0];
#pragma unroll
        for (int dd = 0; dd < 7; ++dd) {
            Corners c;
            stencil_corners(P, 0, dd, px, py, pz, c);
            float gi[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int aa = k & 1, bb = (k >> 1) & 1, e = k >> 2;
                const int x = c.x0 + aa, y = c.y0 + bb, z = c.z0 + e;
                if (!corner_in(P, 0, x, y, z)) continue;
                const int64_t off = base + ((int64_t)z * P.H[0] + y) * P.W[0] + x;
                if (gx0) atomicAdd(gx0 + off, g[dd] * (c.wx[aa] * c.wy[bb] * c.wz[e]));
                if (NEED_DPTS) {
                    const float v = __ldg(x0 + off) * g[dd];
                    gi[0] += v * (aa ? 1.f : -1.f) * c.wy[bb] * c.wz[e];
                    gi[1] += v * (bb ? 1.f : -1.f) * c.wx[aa] * c.wz[e];
                    gi[2] += v * (e ? 1.f : -1.f) * c.wx[aa] * c.wy[bb];
                }
            }
            if (NEED_DPTS) {
                dq[0] += gi[0] * (P.align ? 0.5f * (P.W[0] - 1) : 0.5f * P.W[0]);
                dq[1] += gi[1] * (P.align ? 0.5f * (P.H[0] - 1) : 0.5f * P.H[0]);
                dq[2] += gi[2] * (P.align ? 0.5f * (P.D[0] - 1) : 0.5f * P.D[0]);
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
            const int aa = k & 1, bb = (k >> 1) & 1, e = k >> 2;
            const int x = c.x0 + aa, y = c.y0 + bb, z = c.z0 + e;
            if (!corner_in(P, level, x, y, z)) continue;
            const int64_t off = vbase + (((int64_t)z * P.H[level] + y) * P.W[level] + x) * C;
            const float w = c.wx[aa] * c.wy[bb] * c.wz[e];
            if (gv.g[level]) {
                float *dst = gv.g[level] + off;
                red_add_v4f(dst, g[0] * w, g[1] * w, g[2] * w, g[3] * w);
                red_add_v4f(dst + 4, g[4] * w, g[5] * w, g[6] * w, g[7] * w);
            }
            if (NEED_DPTS) {
                const float4 v0 = __ldg(reinterpret_cast<const float4 *>(vols.v[level] + off));
                const float4 v1 = __ldg(reinterpret_cast<const float4 *>(vols.v[level] + off) + 1);
                float dot = v0.x * g[0];
                dot = fmaf(v0.y, g[1], dot);
                dot = fmaf(v0.z, g[2], dot);
                dot = fmaf(v0.w, g[3], dot);
                dot = fmaf(v1.x, g[4], dot);
                dot = fmaf(v1.y, g[5], dot);
                dot = fmaf(v1.z, g[6], dot);
                dot = fmaf(v1.w, g[7], dot);
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
    if (NEED_DPTS) {   // d pt[0] = 2*dq.z, d pt[1] = 2*dq.y, d pt[2] = 2*dq.x (ifnet.py:156-157 swaps and doubles the coordinates)
        if (dq[2] != 0.f) atomicAdd(gpoints + pt * 3 + 0, 2.f * dq[2]);
        if (dq[1] != 0.f) atomicAdd(gpoints + pt * 3 + 1, 2.f * dq[1]);
        if (dq[0] != 0.f) atomicAdd(gpoints + pt * 3 + 2, 2.f * dq[0]);
    }
}

// dz2 = dlogit (x) wout masked by h2 > 0 (fp32); per-block partials of gwout / gbout
__global__ void __launch_bounds__(256) head_bwd_f32_kernel(const float *__restrict__ dlogit, const float *__restrict__ h2,
                                                           const float *__restrict__ wout, int M, int Hd, float *__restrict__ dz2,
                                                           float *__restrict__ part, int rows_per_block) {
    // thread = column (Hd <= 256); rows of the block walked serially: fixed summation order
    const int col = threadIdx.x;
    const int r0 = blockIdx.x * rows_per_block;
    int r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    const float w = col < Hd ? wout[col] : 0.f;
    float gw = 0.f, gb = 0.f;
    for (int r = r0; r < r1; ++r) {
        const float dl = dlogit[r];
        gb += dl;
        if (col < Hd) {
            const float h = h2[(int64_t)r * Hd + col];
            gw = fmaf(dl, h, gw);
            dz2[(int64_t)r * Hd + col] = h > 0.f ? dl * w : 0.f;
        }
    }
    if (col < Hd) part[(int64_t)blockIdx.x * (Hd + 1) + col] = gw;
    if (col == 0) part[(int64_t)blockIdx.x * (Hd + 1) + Hd] = gb;
}

// column sums of an fp32 (M, N) matrix, two deterministic levels (per-block partials, then a fixed-order sum)
__global__ void __launch_bounds__(256) colsum_f32_partial_kernel(const float *__restrict__ a, int M, int N, int64_t lda, int rows_per_block,
                                                                 float *__restrict__ part) {
    const int col = blockIdx.x * 256 + threadIdx.x;
    if (col >= N) return;
    const int r0 = blockIdx.y * rows_per_block;
    int r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    float v = 0.f;
    for (int r = r0; r < r1; ++r) v += a[(int64_t)r * lda + col];
    part[(int64_t)blockIdx.y * N + col] = v;
}

__global__ void colsum_f32_reduce_kernel(const float *__restrict__ part, int nblocks, int N, int64_t ldp, float *__restrict__ out) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= N) return;
    float v = 0.f;
    for (int b = 0; b < nblocks; ++b) v += part[(int64_t)b * ldp + col];
    out[col] = v;
}

// fc_0.weight (H0, C*7) fp32, reference order k = c*7+d  ->  kernel K' order (H0, KP) fp32 with zero padding
__global__ void pack_w0_f32_kernel(const float *__restrict__ w0, int H0, Pyr P, float *__restrict__ w0p) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)H0 * P.kp) return;
    const int o = (int)(i / P.kp), kprime = (int)(i % P.kp);
    const int u = kprime >> 3, j = kprime & 7;
    const int K = P.ctot * 7;
    float v = 0.f;
    int level, d, c0;
    if (decode_unit(P, u, level, d, c0)) {
        if (level == 0) {
            if (j < 7) v = w0[(int64_t)o * K + j];
        } else {
            v = w0[(int64_t)o * K + (int64_t)(P.coff[level] + c0 + j) * 7 + d];
        }
    }
    w0p[i] = v;
}

static int fill_vols_f32(VolPtrsF &vp, const float *const *vols_host, const Pyr &P) {
    SVR_REQUIRE(vols_host, "null volume pointer table");
    for (int l = 0; l < SVR_MAX_LEVELS; ++l) {
        vp.v[l] = (l >= 1 && l < P.n_levels) ? vols_host[l] : nullptr;
        SVR_REQUIRE(!(l >= 1 && l < P.n_levels) || vp.v[l], "volume of level %d is null", l);
        SVR_REQUIRE(((uintptr_t)vp.v[l] & 15) == 0, "volume of level %d is not 16-byte aligned", l);
    }
    return 0;
}

}  // namespace svr

using namespace svr;

extern "C" {

int svr_split_bf16(const float *x, int64_t n, uint16_t *hi, uint16_t *lo, void *stream) {
    SVR_REQUIRE(x && hi && lo, "split_bf16: null pointer");
    SVR_REQUIRE(n >= 0 && n % 8 == 0 && (((uintptr_t)x | (uintptr_t)hi | (uintptr_t)lo) & 15) == 0,
                "split_bf16: n must be a multiple of 8 and the pointers 16-byte aligned");
    if (n == 0) return 0;
    const int64_t n8 = n / 8;
    int64_t blocks = ceil_div<int64_t>(n8, 256), cap = (int64_t)sm_count() * 16;
    split_bf16_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(x, n8, (__nv_bfloat16 *)hi, (__nv_bfloat16 *)lo);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_pack_w0_f32(const float *w0, int H0, const svr_pyramid *pyr_host, float *w0p, void *stream) {
    Pyr P;
    if (int rc = make_pyr(P, pyr_host)) return rc;
    SVR_REQUIRE(w0 && w0p && H0 > 0, "pack_w0_f32: bad arguments");
    const int64_t n = (int64_t)H0 * P.kp;
    pack_w0_f32_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(w0, H0, P, w0p);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_gather_fwd_f32(const float *points, int B, int N, const float *x0, const float *const *vols_host, const svr_pyramid *pyr_host,
                       float *feat, void *stream) {
    Pyr P;
    if (int rc = make_pyr(P, pyr_host)) return rc;
    VolPtrsF vp;
    if (int rc = fill_vols_f32(vp, vols_host, P)) return rc;
    SVR_REQUIRE(points && x0 && feat, "gather_fwd_f32: null pointer");
    const int64_t total_pts = (int64_t)B * N;
    if (total_pts == 0) return 0;
    const int64_t total = total_pts * (P.kp / 8);
    int64_t blocks = ceil_div<int64_t>(total, 256), cap = (int64_t)sm_count() * 64;
    gather_fwd_f32_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(points, N, total_pts, x0, vp, P, feat);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_gather_bwd_f32(const float *points, int B, int N, const float *x0, const float *const *vols_host, const svr_pyramid *pyr_host,
                       const float *dfeat, float *gx0, float *const *gvols_host, float *gpoints, void *stream) {
    Pyr P;
    if (int rc = make_pyr(P, pyr_host)) return rc;
    VolPtrsF vp;
    if (int rc = fill_vols_f32(vp, vols_host, P)) return rc;
    SVR_REQUIRE(points && x0 && dfeat, "gather_bwd_f32: null pointer");
    GradPtrsF gp;
    for (int l = 0; l < SVR_MAX_LEVELS; ++l) {
        gp.g[l] = (gvols_host && l >= 1 && l < P.n_levels) ? gvols_host[l] : nullptr;
        SVR_REQUIRE(((uintptr_t)gp.g[l] & 15) == 0, "gradient volume of level %d is not 16-byte aligned", l);
    }
    const int64_t total_pts = (int64_t)B * N;
    if (total_pts == 0) return 0;
    const int64_t total = total_pts * (P.kp / 8);
    SVR_REQUIRE(ceil_div<int64_t>(total, 256) < ((int64_t)1 << 31), "gather_bwd_f32: too many points");
    const unsigned blocks = (unsigned)ceil_div<int64_t>(total, 256);
    if (gpoints)
        gather_bwd_f32_kernel<true><<<blocks, 256, 0, as_stream(stream)>>>(points, N, total_pts, x0, vp, P, dfeat, gx0, gp, gpoints);
    else
        gather_bwd_f32_kernel<false><<<blocks, 256, 0, as_stream(stream)>>>(points, N, total_pts, x0, vp, P, dfeat, gx0, gp, gpoints);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_decoder_head_bwd_f32(const float *dlogit, const float *h2, const float *wout, int M, int Hd, float *dz2, float *gwout,
                             float *gbout, void *stream) {
    SVR_REQUIRE(dlogit && h2 && wout && dz2 && gwout && gbout, "decoder_head_bwd_f32: null pointer");
    SVR_REQUIRE(Hd > 0 && Hd <= 256, "decoder_head_bwd_f32: hidden size must be <= 256");
    cudaStream_t st = as_stream(stream);
    if (M == 0) {
        SVR_CUDA(cudaMemsetAsync(gwout, 0, sizeof(float) * Hd, st));
        SVR_CUDA(cudaMemsetAsync(gbout, 0, sizeof(float), st));
        return 0;
    }
    const int rows_per_block = ceil_div(M, 4 * sm_count()) > 8 ? ceil_div(M, 4 * sm_count()) : 8;
    const int nblocks = ceil_div(M, rows_per_block);
    float *scratch = nullptr;
    if (int rc = scratch_alloc((void **)&scratch, (size_t)nblocks * (Hd + 1) * sizeof(float), st)) return rc;
    head_bwd_f32_kernel<<<nblocks, 256, 0, st>>>(dlogit, h2, wout, M, Hd, dz2, scratch, rows_per_block);
    colsum_f32_reduce_kernel<<<ceil_div(Hd, 256), 256, 0, st>>>(scratch, nblocks, Hd, Hd + 1, gwout);
    colsum_f32_reduce_kernel<<<1, 32, 0, st>>>(scratch + Hd, nblocks, 1, Hd + 1, gbout);
    SVR_LAUNCH_CHECK();
    SVR_CUDA(cudaFreeAsync(scratch, st));
    return 0;
}

int svr_colsum_f32(const float *a, int M, int N, int64_t lda, float *out, void *stream) {
    SVR_REQUIRE(a && out && N > 0, "colsum_f32: bad arguments");
    cudaStream_t st = as_stream(stream);
    if (M == 0) {
        SVR_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, st));
        return 0;
    }
    const int rows_per_block = ceil_div(M, 4 * sm_count()) > 8 ? ceil_div(M, 4 * sm_count()) : 8;
    const int nblocks = ceil_div(M, rows_per_block);
    float *scratch = nullptr;
    if (int rc = scratch_alloc((void **)&scratch, (size_t)nblocks * N * sizeof(float), st)) return rc;
    colsum_f32_partial_kernel<<<dim3(ceil_div(N, 256), nblocks), 256, 0, st>>>(a, M, N, lda, rows_per_block, scratch);
    colsum_f32_reduce_kernel<<<ceil_div(N, 256), 256, 0, st>>>(scratch, nblocks, N, N, out);
    SVR_LAUNCH_CHECK();
    SVR_CUDA(cudaFreeAsync(scratch, st));
    return 0;
}
}
