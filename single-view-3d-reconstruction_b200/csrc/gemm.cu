// tcgen05/TMEM bf16 GEMMs for the IF-Net decoder (the Conv1d(k=1) layers of model/ifnet.py:55-59
// and their backward), plus the small fused helpers around them.
//
//   NT : C[M,N] = epi(A[M,K] . B[N,K]^T)        -- forward layers and backward-data
//   TN : C[M,N] = A[P,M]^T . B[P,N]              -- weight gradients (contraction over points)
//
// One CTA computes a 128 x 256 fp32 accumulator tile in TMEM (256 columns).  Warp roles:
//   warps 0-3  epilogue (TMEM lanes 32w..32w+31 -> registers -> global)
//   warp  4    TMEM allocation + single-thread UMMA issue
//   warps 5-8  producers: cp.async 16-byte chunks into 128B-swizzled stages
// Stages are handed over with mbarriers (full: 128 producer arrivals after a writer-side
// fence.proxy.async; empty: tcgen05.commit).
#include "common.cuh"
#include "tc05.cuh"

namespace svr {
using namespace tc;

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KB
constexpr int GEMM_THREADS = 288;
constexpr int NUM_PRODUCERS = 128;
constexpr size_t gemm_smem(int stages) { return 1024 + (size_t)stages * (A_STAGE_BYTES + B_STAGE_BYTES) + 256; }

struct GemmNT {
    const __nv_bfloat16 *A, *B;
    int64_t lda, ldb, ldc;
    const float *bias;
    int M, N, K, flags;
    __nv_bfloat16 *c_bf16;
    float *c_f32;
    const __nv_bfloat16 *mask;
    const float *dot_w;
    const float *dot_b;
    float *out_dot;
};

struct GemmTN {
    const __nv_bfloat16 *A, *B;
    int64_t lda, ldb;
    int M, N, P, chunks_per_split;
    float *partial;   // [splits][M][N]
};

struct SmemLayout {
    uint8_t *a, *b;
    uint64_t *full, *empty, *tmem_full;
    uint32_t *tmem_ptr;
};

template <int STAGES>
__device__ __forceinline__ SmemLayout carve(uint8_t *raw) {
    SmemLayout s;
    uint8_t *base = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    s.a = base;
    s.b = base + STAGES * A_STAGE_BYTES;
    uint8_t *tail = s.b + STAGES * B_STAGE_BYTES;
    s.full = (uint64_t *)tail;
    s.empty = s.full + STAGES;
    s.tmem_full = s.empty + STAGES;
    s.tmem_ptr = (uint32_t *)(s.tmem_full + 1);
    return s;
}

template <int STAGES>
__device__ __forceinline__ uint32_t gemm_prologue(const SmemLayout &s, int warp) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(s.full + i, NUM_PRODUCERS);
            mbar_init(s.empty + i, 1);
        }
        mbar_init(s.tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(s.tmem_ptr, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *s.tmem_ptr;
}

// producer bookkeeping shared by both kernels: signal stage (kc-LAG) once its copies have landed
template <int STAGES>
__device__ __forceinline__ void producer_signal(const SmemLayout &s, int kc_done) {
    fence_proxy_async();
    mbar_arrive(s.full + (kc_done % STAGES));
}

// ------------------------------------------------------------------------------------------------
// NT kernel
// ------------------------------------------------------------------------------------------------
template <int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, (STAGES <= 2 ? 2 : 1)) gemm_nt_kernel(GemmNT p) {
    constexpr int LAG = STAGES <= 2 ? 1 : 2;
    extern __shared__ uint8_t smem_raw[];
    const SmemLayout s = carve<STAGES>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tmem = gemm_prologue<STAGES>(s, warp);
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int KC = p.K / BK;

    if (warp >= 5) {
        const int pt = threadIdx.x - 160;
        for (int kc = 0; kc < KC; ++kc) {
            const int st = kc % STAGES;
            mbar_wait(s.empty + st, ((kc / STAGES) & 1) ^ 1);
            const uint32_t a_s = smem_u32(s.a + st * A_STAGE_BYTES), b_s = smem_u32(s.b + st * B_STAGE_BYTES);
#pragma unroll
            for (int i = 0; i < (BM * 8) / NUM_PRODUCERS; ++i) {
                int c = pt + i * NUM_PRODUCERS, row = c >> 3, ch = c & 7;
                bool ok = (m0 + row) < p.M;
                const __nv_bfloat16 *src = p.A + (int64_t)(ok ? m0 + row : 0) * p.lda + (int64_t)kc * BK + ch * 8;
                cp_async16(a_s + swz128(row, ch), src, ok);
            }
#pragma unroll
            for (int i = 0; i < (BN * 8) / NUM_PRODUCERS; ++i) {
                int c = pt + i * NUM_PRODUCERS, row = c >> 3, ch = c & 7;
                bool ok = (n0 + row) < p.N;
                const __nv_bfloat16 *src = p.B + (int64_t)(ok ? n0 + row : 0) * p.ldb + (int64_t)kc * BK + ch * 8;
                cp_async16(b_s + swz128(row, ch), src, ok);
            }
            cp_async_commit();
            if (kc >= LAG) {
                cp_async_wait<LAG>();
                producer_signal<STAGES>(s, kc - LAG);
            }
        }
        cp_async_wait<0>();
        for (int kc = (KC > LAG ? KC - LAG : 0); kc < KC; ++kc) producer_signal<STAGES>(s, kc);
    } else if (warp == 4) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
            for (int kc = 0; kc < KC; ++kc) {
                const int st = kc % STAGES;
                mbar_wait(s.full + st, (kc / STAGES) & 1);
                tc_fence_after();
                const uint32_t a_s = smem_u32(s.a + st * A_STAGE_BYTES), b_s = smem_u32(s.b + st * B_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    uint64_t ad = make_smem_desc(a_s + k * 32, 16, 1024, kSwizzle128B);
                    uint64_t bd = make_smem_desc(b_s + k * 32, 16, 1024, kSwizzle128B);
                    umma_bf16(tmem, ad, bd, idesc, (kc | k) != 0);
                }
                umma_commit(s.empty + st);
            }
            umma_commit(s.tmem_full);
        }
        __syncwarp();
    } else {
        mbar_wait(s.tmem_full, 0);
        tc_fence_after();
        const int row = m0 + warp * 32 + lane;
        const bool row_ok = row < p.M;
        const bool relu = p.flags & 1, st_bf = p.flags & 2, st_f = p.flags & 4, use_mask = p.flags & 8, dot = p.flags & 16;
        const bool accum = p.flags & 32;   // C_f32 += A.B^T : the running sum of a split-operand (hi/lo) product, see precise.cu
        float dot_acc = 0.f;
        for (int c0 = 0; c0 < BN; c0 += 32) {
            if (n0 + c0 >= p.N) break;   // warp-uniform
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            tmem_ld_wait();
            if (accum && row_ok) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    int n = n0 + c0 + q * 4;
                    if (n < p.N) {
                        const float4 prev = *reinterpret_cast<const float4 *>(p.c_f32 + (int64_t)row * p.ldc + n);
                        v[q * 4] = __float_as_uint(__uint_as_float(v[q * 4]) + prev.x);
                        v[q * 4 + 1] = __float_as_uint(__uint_as_float(v[q * 4 + 1]) + prev.y);
                        v[q * 4 + 2] = __float_as_uint(__uint_as_float(v[q * 4 + 2]) + prev.z);
                        v[q * 4 + 3] = __float_as_uint(__uint_as_float(v[q * 4 + 3]) + prev.w);
                    }
                }
            }
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                int n = n0 + c0 + j;
                float x = __uint_as_float(v[j]);
                if (p.bias && n < p.N) x += p.bias[n];
                if (relu) x = fmaxf(x, 0.f);
                f[j] = x;
            }
            if (use_mask && row_ok) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    int n = n0 + c0 + q * 8;
                    if (n < p.N) {
                        uint4 mk = *reinterpret_cast<const uint4 *>(p.mask + (int64_t)row * p.ldc + n);
                        const __nv_bfloat16 *mb = reinterpret_cast<const __nv_bfloat16 *>(&mk);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (!(__bfloat162float(mb[j]) > 0.f)) f[q * 8 + j] = 0.f;
                    }
                }
            }
            if (dot) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    int n = n0 + c0 + j;
                    if (n < p.N) dot_acc += f[j] * p.dot_w[n];
                }
            }
            if (row_ok) {
                if (st_bf) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        int n = n0 + c0 + q * 8;
                        if (n < p.N) {
                            __nv_bfloat162 h[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[q * 8 + 2 * j], f[q * 8 + 2 * j + 1]);
                            *reinterpret_cast<uint4 *>(p.c_bf16 + (int64_t)row * p.ldc + n) = *reinterpret_cast<uint4 *>(h);
                        }
                    }
                }
                if (st_f) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        int n = n0 + c0 + q * 4;
                        if (n < p.N)
                            *reinterpret_cast<float4 *>(p.c_f32 + (int64_t)row * p.ldc + n) =
                                make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
                    }
                }
            }
        }
        if (dot && row_ok) p.out_dot[row] = dot_acc + (p.dot_b ? __ldg(p.dot_b) : 0.f);
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

// ------------------------------------------------------------------------------------------------
// TN kernel: both operands MN-major.  Stage image: 64-element (128 B) column blocks, each holding
// 64 k-rows of 128 B (8-row swizzle atoms of 1024 B); column blocks are 8192 B apart.
// ------------------------------------------------------------------------------------------------
template <int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, (STAGES <= 2 ? 2 : 1)) gemm_tn_kernel(GemmTN p) {
    constexpr int LAG = STAGES <= 2 ? 1 : 2;
    extern __shared__ uint8_t smem_raw[];
    const SmemLayout s = carve<STAGES>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tmem = gemm_prologue<STAGES>(s, warp);
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN, split = blockIdx.z;
    const int total_chunks = (p.P + BK - 1) / BK;
    const int kc_begin = split * p.chunks_per_split;
    int KC = total_chunks - kc_begin;
    if (KC > p.chunks_per_split) KC = p.chunks_per_split;
    if (KC < 0) KC = 0;

    if (warp >= 5) {
        const int pt = threadIdx.x - 160;
        for (int kc = 0; kc < KC; ++kc) {
            const int st = kc % STAGES;
            mbar_wait(s.empty + st, ((kc / STAGES) & 1) ^ 1);
            const uint32_t a_s = smem_u32(s.a + st * A_STAGE_BYTES), b_s = smem_u32(s.b + st * B_STAGE_BYTES);
            const int64_t p0 = (int64_t)(kc_begin + kc) * BK;
#pragma unroll
            for (int i = 0; i < (BK * 16) / NUM_PRODUCERS; ++i) {   // A: 64 k-rows x 16 chunks (128 cols)
                int c = pt + i * NUM_PRODUCERS, kr = c >> 4, ch = c & 15;
                bool ok = (p0 + kr) < p.P && (m0 + ch * 8) < p.M;
                const __nv_bfloat16 *src = p.A + (ok ? (p0 + kr) * p.lda + m0 + ch * 8 : 0);
                cp_async16(a_s + (ch >> 3) * 8192 + swz128(kr, ch & 7), src, ok);
            }
#pragma unroll
            for (int i = 0; i < (BK * 32) / NUM_PRODUCERS; ++i) {   // B: 64 k-rows x 32 chunks (256 cols)
                int c = pt + i * NUM_PRODUCERS, kr = c >> 5, ch = c & 31;
                bool ok = (p0 + kr) < p.P && (n0 + ch * 8) < p.N;
                const __nv_bfloat16 *src = p.B + (ok ? (p0 + kr) * p.ldb + n0 + ch * 8 : 0);
                cp_async16(b_s + (ch >> 3) * 8192 + swz128(kr, ch & 7), src, ok);
            }
            cp_async_commit();
            if (kc >= LAG) {
                cp_async_wait<LAG>();
                producer_signal<STAGES>(s, kc - LAG);
            }
        }
        cp_async_wait<0>();
        for (int kc = (KC > LAG ? KC - LAG : 0); kc < KC; ++kc) producer_signal<STAGES>(s, kc);
    } else if (warp == 4) {
        if (lane == 0 && KC > 0) {
            const uint32_t idesc = make_idesc_bf16(BM, BN, 1, 1);
            for (int kc = 0; kc < KC; ++kc) {
                const int st = kc % STAGES;
                mbar_wait(s.full + st, (kc / STAGES) & 1);
                tc_fence_after();
                const uint32_t a_s = smem_u32(s.a + st * A_STAGE_BYTES), b_s = smem_u32(s.b + st * B_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    uint64_t ad = make_smem_desc(a_s + k * 2048, 8192, 1024, kSwizzle128B);
                    uint64_t bd = make_smem_desc(b_s + k * 2048, 8192, 1024, kSwizzle128B);
                    umma_bf16(tmem, ad, bd, idesc, (kc | k) != 0);
                }
                umma_commit(s.empty + st);
            }
            umma_commit(s.tmem_full);
        }
        __syncwarp();
    } else {
        const int row = m0 + warp * 32 + lane;
        float *dst = p.partial + ((int64_t)split * p.M + row) * p.N;
        if (KC > 0) {
            mbar_wait(s.tmem_full, 0);
            tc_fence_after();
        }
        for (int c0 = 0; c0 < BN; c0 += 32) {
            if (n0 + c0 >= p.N) break;
            uint32_t v[32];
            if (KC > 0) {
                tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            if (row < p.M) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    int n = n0 + c0 + q * 4;
                    if (n < p.N)
                        *reinterpret_cast<float4 *>(dst + n) =
                            make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]),
                                        __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

__global__ void splitk_reduce_kernel(const float *__restrict__ partial, int splits, int64_t MN, int N, float *__restrict__ C,
                                     int64_t ldc, int accumulate) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= MN) return;
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += partial[(int64_t)s * MN + i];   // fixed order: deterministic
    int64_t r = i / N, c = i % N;
    float *d = C + r * ldc + c;
    *d = accumulate ? *d + v : v;
}

// ------------------------------------------------------------------------------------------------
// small helpers of the decoder backward
// ------------------------------------------------------------------------------------------------
// dz2 = dlogit (x) wout masked by h2 > 0;  per-block partials of gwout / gbout.
// thread = (row lane, 8-column group): 16-byte loads/stores, 512 B per row and warp.
__global__ void __launch_bounds__(256) head_bwd_kernel(const float *__restrict__ dlogit, const int *__restrict__ perm,
                                                       const __nv_bfloat16 *__restrict__ h2, const float *__restrict__ wout, int M, int Hd,
                                                       __nv_bfloat16 *__restrict__ dz2, float *__restrict__ part_w,
                                                       float *__restrict__ part_b, int rows_per_block) {
    __shared__ float red_w[8][256];
    __shared__ float red_b[8];
    const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;        // 32 column groups (256 columns) x 8 row lanes
    const int c0 = cg * 8;
    const bool col_ok = c0 < Hd;
    const int r0 = blockIdx.x * rows_per_block;
    int r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    float w[8], gw[8], gb = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        w[j] = (col_ok && c0 + j < Hd) ? wout[c0 + j] : 0.f;
        gw[j] = 0.f;
    }
    for (int r = r0 + rl; r < r1; r += 8) {
        const float dl = dlogit[perm ? perm[r] : r];
        if (cg == 0) gb += dl;
        if (col_ok) {
            const uint4 raw = *reinterpret_cast<const uint4 *>(h2 + (int64_t)r * Hd + c0);
            const uint32_t ww[4] = {raw.x, raw.y, raw.z, raw.w};
            float h[8], o[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                h[2 * j] = __uint_as_float(ww[j] << 16);
                h[2 * j + 1] = __uint_as_float(ww[j] & 0xffff0000u);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                gw[j] = fmaf(dl, h[j], gw[j]);
                o[j] = h[j] > 0.f ? dl * w[j] : 0.f;
            }
            __nv_bfloat162 pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) pk[j] = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
            *reinterpret_cast<uint4 *>(dz2 + (int64_t)r * Hd + c0) = *reinterpret_cast<uint4 *>(pk);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red_w[rl][c0 + j] = gw[j];
    if (cg == 0) red_b[rl] = gb;
    __syncthreads();
    const int col = threadIdx.x;
    if (col < Hd) {
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) v += red_w[q][col];
        part_w[(int64_t)blockIdx.x * Hd + col] = v;
    }
    if (col == 0) {
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) v += red_b[q];
        part_b[blockIdx.x] = v;
    }
}

__global__ void head_bwd_reduce_kernel(const float *__restrict__ part_w, const float *__restrict__ part_b, int nblocks,
                                       int Hd, float *__restrict__ gwout, float *__restrict__ gbout) {
    // block = 32 columns x 32 block slices (a serial walk over the blocks is a chain of ~300 dependent L2 round trips);
    // column Hd is the bias partial.  Fixed summation order: deterministic.
    __shared__ float sl[32][33];
    const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + lane;
    float v = 0.f;
    if (col < Hd)
        for (int b = part; b < nblocks; b += 32) v += part_w[(int64_t)b * Hd + col];
    else if (col == Hd)
        for (int b = part; b < nblocks; b += 32) v += part_b[b];
    sl[part][lane] = v;
    __syncthreads();
    if (part == 0 && col <= Hd) {
        v = 0.f;
        for (int k = 0; k < 32; ++k) v += sl[k][lane];
        if (col < Hd)
            gwout[col] += v;
        else
            gbout[0] += v;
    }
}

// column sums of a bf16 (M, N) matrix: thread = (row lane, 8-column group), 16-byte loads
__global__ void __launch_bounds__(256) colsum_partial_kernel(const __nv_bfloat16 *__restrict__ a, int M, int N, int64_t lda,
                                                             int rows_per_block, float *__restrict__ part) {
    __shared__ float red[8][256];
    const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 256 + cg * 8;
    const int r0 = blockIdx.y * rows_per_block;
    int r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (c0 < N) {
        for (int r = r0 + rl; r < r1; r += 8) {
            const uint4 raw = *reinterpret_cast<const uint4 *>(a + (int64_t)r * lda + c0);
            const uint32_t ww[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[2 * j] += __uint_as_float(ww[j] << 16);
                acc[2 * j + 1] += __uint_as_float(ww[j] & 0xffff0000u);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[rl][cg * 8 + j] = acc[j];
    __syncthreads();
    const int col = blockIdx.x * 256 + threadIdx.x;
    if (col < N) {
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) v += red[q][threadIdx.x];
        part[(int64_t)blockIdx.y * N + col] = v;
    }
}

__global__ void colsum_reduce_kernel(const float *__restrict__ part, int nblocks, int N, float *__restrict__ out,
                                     int accumulate) {
    __shared__ float sl[32][33];     // 32 columns x 32 block slices, fixed summation order
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + lane;
    float v = 0.f;
    if (col < N)
        for (int b = slice; b < nblocks; b += 32) v += part[(int64_t)b * N + col];
    sl[slice][lane] = v;
    __syncthreads();
    if (slice == 0 && col < N) {
        v = 0.f;
        for (int k = 0; k < 32; ++k) v += sl[k][lane];
        out[col] = accumulate ? out[col] + v : v;
    }
}

__global__ void pack_matrix_kernel(const float *__restrict__ w, int R, int C, __nv_bfloat16 *__restrict__ wp,
                                   __nv_bfloat16 *__restrict__ wpT) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)R * C) return;
    int r = (int)(i / C), c = (int)(i % C);
    __nv_bfloat16 v = __float2bfloat16(w[i]);
    if (wp) wp[i] = v;
    if (wpT) wpT[(int64_t)c * R + r] = v;
}

static int tn_splits(int M, int N, int P) {
    int tiles = ceil_div(M, BM) * ceil_div(N, BN);
    int chunks = ceil_div(P, BK);
    int s = sm_count() / (tiles > 0 ? tiles : 1);
    if (s < 1) s = 1;
    if (s > chunks) s = chunks > 0 ? chunks : 1;
    if (s > 64) s = 64;
    return s;
}

static int set_gemm_attrs() {
    static DeviceOnce once;
    int dev;
    if (!once.needed(dev)) return 0;
    SVR_CUDA(cudaFuncSetAttribute(gemm_nt_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem(2)));
    SVR_CUDA(cudaFuncSetAttribute(gemm_nt_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem(4)));
    SVR_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem(4)));
    once.done(dev);
    return 0;
}

// scratch for the head/colsum partials: stream-ordered allocation from the library's private pool (common.cu)
static int scratch_alloc(float **ptr, size_t n_floats, cudaStream_t st) {
    return ::svr::scratch_alloc((void **)ptr, n_floats * sizeof(float), st);
}

}  // namespace svr

using namespace svr;

extern "C" {

int svr_gemm_nt(const uint16_t *A, int64_t lda, const uint16_t *B, int64_t ldb, const float *bias, int M, int N, int K,
                int flags, uint16_t *c_bf16, float *c_f32, int64_t ldc, const uint16_t *mask, const float *dot_w,
                const float *dot_b, float *out_dot, void *stream) {
    SVR_REQUIRE(A && B, "gemm_nt: null operand");
    SVR_REQUIRE(M >= 0 && N > 0 && K > 0 && K % BK == 0, "gemm_nt: K (%d) must be a positive multiple of %d", K, BK);
    SVR_REQUIRE(N % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, "gemm_nt: N, lda, ldb must be multiples of 8");
    SVR_REQUIRE(!(flags & 2) || (c_bf16 && ldc % 8 == 0), "gemm_nt: bf16 output needs c_bf16 and ldc %% 8 == 0");
    SVR_REQUIRE(!(flags & 4) || (c_f32 && ldc % 4 == 0), "gemm_nt: fp32 output needs c_f32 and ldc %% 4 == 0");
    SVR_REQUIRE(!(flags & 8) || (mask && ldc % 8 == 0), "gemm_nt: mask needs a pointer and ldc %% 8 == 0");
    SVR_REQUIRE(!(flags & 16) || (dot_w && out_dot && N <= BN), "gemm_nt: row-dot needs dot_w/out_dot and N <= %d", BN);
    SVR_REQUIRE(!(flags & 32) || (c_f32 && ldc % 4 == 0), "gemm_nt: accumulation needs the fp32 output");
    if (M == 0) return 0;
    GemmNT p{(const __nv_bfloat16 *)A, (const __nv_bfloat16 *)B, lda, ldb, ldc, bias, M, N, K, flags,
             (__nv_bfloat16 *)c_bf16, c_f32, (const __nv_bfloat16 *)mask, dot_w, dot_b, out_dot};
    if (int rc = set_gemm_attrs()) return rc;
    dim3 grid(ceil_div(M, BM), ceil_div(N, BN));
    if (K <= 4 * BK)   // short-K (backward-data) GEMMs: 2 stages -> 2 CTAs/SM, epilogues overlap the next tile's loads
        gemm_nt_kernel<2><<<grid, GEMM_THREADS, gemm_smem(2), as_stream(stream)>>>(p);
    else
        gemm_nt_kernel<4><<<grid, GEMM_THREADS, gemm_smem(4), as_stream(stream)>>>(p);
    SVR_LAUNCH_CHECK();
    return 0;
}

size_t svr_gemm_tn_workspace_bytes(int M, int N, int P) {
    return (size_t)tn_splits(M, N, P) * (size_t)M * (size_t)N * sizeof(float) + 256;
}

int svr_gemm_tn(const uint16_t *A, int64_t lda, const uint16_t *B, int64_t ldb, int M, int N, int P, float *C, int64_t ldc,
                int accumulate, void *workspace, size_t workspace_bytes, void *stream) {
    SVR_REQUIRE(A && B && C && workspace, "gemm_tn: null pointer");
    SVR_REQUIRE(M > 0 && N > 0 && P >= 0, "gemm_tn: bad sizes");
    SVR_REQUIRE(M % 8 == 0 && N % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, "gemm_tn: M, N, lda, ldb must be multiples of 8");
    const int splits = tn_splits(M, N, P);
    SVR_REQUIRE(workspace_bytes >= (size_t)splits * M * N * sizeof(float), "gemm_tn: workspace too small");
    if (int rc = set_gemm_attrs()) return rc;
    const int chunks = ceil_div(P, BK);
    GemmTN p{(const __nv_bfloat16 *)A, (const __nv_bfloat16 *)B, lda, ldb, M, N, P, ceil_div(chunks > 0 ? chunks : 1, splits),
             (float *)workspace};
    dim3 grid(ceil_div(M, BM), ceil_div(N, BN), splits);
    cudaStream_t st = as_stream(stream);
    gemm_tn_kernel<4><<<grid, GEMM_THREADS, gemm_smem(4), st>>>(p);
    SVR_LAUNCH_CHECK();
    int64_t MN = (int64_t)M * N;
    splitk_reduce_kernel<<<(unsigned)ceil_div<int64_t>(MN, 256), 256, 0, st>>>((const float *)workspace, splits, MN, N, C, ldc,
                                                                              accumulate);
    SVR_LAUNCH_CHECK();
    return 0;
}

int svr_decoder_head_bwd(const float *dlogit, const int *perm, const uint16_t *h2, const float *wout, int M, int Hd,
                         uint16_t *dz2, float *gwout, float *gbout, void *stream) {
    SVR_REQUIRE(dlogit && h2 && wout && dz2 && gwout && gbout, "decoder_head_bwd: null pointer");
    SVR_REQUIRE(Hd > 0 && Hd <= 256 && Hd % 8 == 0, "decoder_head_bwd: hidden size must be a multiple of 8 and <= 256");
    if (M == 0) return 0;
    cudaStream_t st = as_stream(stream);
    const int rows_per_block = ceil_div(M, 4 * sm_count()) > 16 ? ceil_div(M, 4 * sm_count()) : 16;
    const int nblocks = ceil_div(M, rows_per_block);
    float *scratch = nullptr;
    if (int rc = scratch_alloc(&scratch, (size_t)nblocks * (Hd + 1), st)) return rc;
    head_bwd_kernel<<<nblocks, 256, 0, st>>>(dlogit, perm, (const __nv_bfloat16 *)h2, wout, M, Hd, (__nv_bfloat16 *)dz2, scratch,
                                             scratch + (size_t)nblocks * Hd, rows_per_block);
    head_bwd_reduce_kernel<<<ceil_div(Hd + 1, 32), 1024, 0, st>>>(scratch, scratch + (size_t)nblocks * Hd, nblocks, Hd, gwout, gbout);
    SVR_LAUNCH_CHECK();
    SVR_CUDA(cudaFreeAsync(scratch, st));
    return 0;
}

int svr_colsum_bf16(const uint16_t *a, int M, int N, int64_t lda, float *out, int accumulate, void *stream) {
    SVR_REQUIRE(a && out && N > 0 && N % 8 == 0 && lda % 8 == 0, "colsum: N and lda must be positive multiples of 8");
    cudaStream_t st = as_stream(stream);
    if (M == 0) {
        if (!accumulate) SVR_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, st));
        return 0;
    }
    const int rows_per_block = ceil_div(M, 2 * sm_count()) > 16 ? ceil_div(M, 2 * sm_count()) : 16;
    const int nblocks = ceil_div(M, rows_per_block);
    float *scratch = nullptr;
    if (int rc = scratch_alloc(&scratch, (size_t)nblocks * N, st)) return rc;
    colsum_partial_kernel<<<dim3(ceil_div(N, 256), nblocks), 256, 0, st>>>((const __nv_bfloat16 *)a, M, N, lda, rows_per_block,
                                                                          scratch);
    colsum_reduce_kernel<<<ceil_div(N, 32), 1024, 0, st>>>(scratch, nblocks, N, out, accumulate);
    SVR_LAUNCH_CHECK();
    SVR_CUDA(cudaFreeAsync(scratch, st));
    return 0;
}

int svr_pack_matrix(const float *w, int R, int C, uint16_t *wp, uint16_t *wpT, void *stream) {
    SVR_REQUIRE(w && (wp || wpT) && R > 0 && C > 0, "pack_matrix: bad arguments");
    int64_t n = (int64_t)R * C;
    pack_matrix_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(w, R, C, (__nv_bfloat16 *)wp,
                                                                                          (__nv_bfloat16 *)wpT);
    SVR_LAUNCH_CHECK();
    return 0;
}
}
