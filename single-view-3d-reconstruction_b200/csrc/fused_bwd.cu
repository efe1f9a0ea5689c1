// Fused decoder backward-data chain (the Conv1d(k=1) backward of model/ifnet.py:55-58, reference root):
//     dz1 = (dz2 . W2) * [h1 > 0]      dz0 = (dz1 . W1) * [h0 > 0]      dfeat = dz0 . W0'
// in ONE persistent kernel.  The 128-row dz tile stays in shared memory as the A operand of all three
// GEMMs (rewritten in place by the epilogues), the pre-swizzled weight chunks are streamed with
// cp.async.bulk, the 11 N-tiles of dfeat alternate between two TMEM accumulators so that the epilogue of
// one overlaps the MMAs of the next.  Replaces three separate NT GEMM launches (11 + 1 + 1 CTAs per row
// tile, each re-loading the dz tile).  All global traffic of the tiles goes through TMA: the dz2 tile is a
// tensor load, dz1 / dz0 are tensor stores straight out of the (swizzled) A operand tile, and dfeat is
// staged in 64-column swizzled boxes and written with tensor stores (a row-per-thread epilogue would issue
// 32 partial sectors per store instruction and be LSU-bound).
//   warps 0-7  row workers (two per TMEM lane quarter, splitting the columns): ReLU-mask epilogues from
//              prefetched mask bits, staging + TMA stores of dz1 / dz0 / dfeat
//   warp  8    loader: dz2 tile (UTMALDG) and weight chunks (UBLKCP)
//   warp  9    single-thread tcgen05.mma issue
#include "common.cuh"
#include "sampling.cuh"
#include "tc05.cuh"

namespace svr {
using namespace tc;

constexpr int FB_TILE = 128, FB_HID = 256, FB_NB = 3;
constexpr int FB_A_BYTES = FB_TILE * FB_HID * 2;     // 64 KB: 4 K-chunks of 16 KB
constexpr int FB_B_BYTES = FB_HID * 128;             // 32 KB: 256 rows x 64 bf16
constexpr int FB_NS = 4;                             // dfeat staging slots (128 rows x 64 columns, 16 KB)
constexpr int FB_S_BYTES = FB_TILE * 128;
constexpr int FB_WORKERS = 256;                      // row-worker threads
constexpr int FB_THREADS = FB_WORKERS + 64;
constexpr int FB_SMEM = 1024 + FB_A_BYTES + FB_NB * FB_B_BYTES + FB_NS * FB_S_BYTES + 256;

struct FbParams {
    TensorMap tm_dz2, tm_dz1, tm_dz0, tm_dfeat;    // boxes of 128 rows x 64 columns, 128-byte swizzle
    const __nv_bfloat16 *h1, *h0;
    const uint8_t *w2t_img, *w1t_img, *w0pt_img;   // chunk images of W2^T (256x256), W1^T (256x256), W0'^T (KP x 256)
    int64_t M;
    int kp;
    long long *trace;   // debug: per-role (tag, clock) records of block 0 (svr_debug_fb_trace), else null
};

// role 0 = row worker thread 0, 1 = MMA thread, 2 = loader thread; 1024 records per role
__device__ __forceinline__ void fb_trace(const FbParams &p, int role, int &n, int tag) {
    if (p.trace && blockIdx.x == 0 && n < 1024) {
        p.trace[(role * 1024 + n) * 2] = tag;
        p.trace[(role * 1024 + n) * 2 + 1] = clock64();
        ++n;
    }
}

__global__ void __launch_bounds__(FB_THREADS, 1) fused_bwd_kernel(const __grid_constant__ FbParams p, int64_t n_tiles) {
    extern __shared__ __align__(1024) uint8_t fb_smem[];
    uint8_t *base = (uint8_t *)(((uintptr_t)fb_smem + 1023) & ~(uintptr_t)1023);
    uint8_t *a_tile = base, *b_ring = base + FB_A_BYTES;
    uint8_t *stage = b_ring + FB_NB * FB_B_BYTES;
    uint64_t *bars = (uint64_t *)(stage + FB_NS * FB_S_BYTES);
    uint64_t *b_full = bars, *b_empty = bars + FB_NB, *acc_full = bars + 2 * FB_NB, *acc_free = acc_full + 2, *a_ready = acc_free + 2;
    uint64_t *a_loaded = a_ready + 1, *a_free = a_ready + 2;
    uint32_t *tmem_ptr = (uint32_t *)(a_ready + 3);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_nt = (p.kp + 255) / 256;                 // N tiles of dfeat
    const int last_n = p.kp - (n_nt - 1) * 256;          // columns of the last N tile (multiple of 64)

    if (threadIdx.x == 0) {
        for (int i = 0; i < FB_NB; ++i) {
            mbar_init(b_full + i, 1);
            mbar_init(b_empty + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(acc_full + i, 1);
            mbar_init(acc_free + i, FB_WORKERS / 32);
        }
        mbar_init(a_ready, 1);    // row worker 0, after the tile-wide named barrier
        mbar_init(a_loaded, 1);   // loader arrive + dz2 transaction bytes
        mbar_init(a_free, 2);     // MMA commit (operand reads done) + row worker 0 (dz0 tensor stores have read the tile)
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    int64_t my_tiles = 0;
    if ((int64_t)blockIdx.x < n_tiles) my_tiles = (n_tiles - 1 - blockIdx.x) / gridDim.x + 1;

    if (warp == 8) {
        // ======================= weight loader =======================
        if (lane == 0) {
            uint32_t wc = 0;
            int tn = 0;
            auto load = [&](const uint8_t *src, uint32_t bytes) {
                const int st = wc % FB_NB;
                mbar_wait(b_empty + st, ((wc / FB_NB) & 1) ^ 1);
                mbar_arrive_expect_tx(b_full + st, bytes);
                bulk_g2s(smem_u32(b_ring + st * FB_B_BYTES), src, bytes, b_full + st);
                ++wc;
            };
            for (int64_t it = 0; it < my_tiles; ++it) {
                const int64_t tile = blockIdx.x + it * gridDim.x;
                mbar_wait(a_free, (it & 1) ^ 1);
                fb_trace(p, 2, tn, 60);
                mbar_arrive_expect_tx(a_loaded, FB_A_BYTES);
                for (int c = 0; c < 4; ++c) tma_load_2d(smem_u32(a_tile + c * (FB_TILE * 128)), &p.tm_dz2, c * 64, (int)(tile * FB_TILE), a_loaded);
                for (int c = 0; c < 4; ++c) load(p.w2t_img + (size_t)c * FB_B_BYTES, FB_B_BYTES);
                for (int c = 0; c < 4; ++c) load(p.w1t_img + (size_t)c * FB_B_BYTES, FB_B_BYTES);
                for (int j = 0; j < n_nt; ++j) {
                    const uint32_t rows = j + 1 < n_nt ? 256u : (uint32_t)last_n;
                    for (int c = 0; c < 4; ++c) load(p.w0pt_img + (size_t)c * p.kp * 128 + (size_t)j * FB_B_BYTES, rows * 128u);
                }
            }
        }
    } else if (warp == 9) {
        // ======================= MMA issue =======================
        if (lane == 0 && my_tiles > 0) {
            uint32_t wc = 0, n_ready = 0, use[2] = {0, 0};
            int tn = 0;
            auto gemm = [&](int acc_i, int n_cols) {    // acc[acc_i] = A tile (K = 256) . next 4 weight chunks
                mbar_wait(acc_free + acc_i, ((use[acc_i] & 1) ^ 1));
                ++use[acc_i];
                tc_fence_after();
                const uint32_t idesc = make_idesc_bf16(FB_TILE, n_cols, 0, 0);
                for (int c = 0; c < 4; ++c, ++wc) {
                    const int st = wc % FB_NB;
                    mbar_wait(b_full + st, (wc / FB_NB) & 1);
                    tc_fence_after();
                    const uint32_t a_s = smem_u32(a_tile + c * (FB_TILE * 128)), b_s = smem_u32(b_ring + st * FB_B_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem + acc_i * 256, make_smem_desc(a_s + k * 32, 16, 1024, kSwizzle128B),
                                  make_smem_desc(b_s + k * 32, 16, 1024, kSwizzle128B), idesc, (c | k) != 0);
                    umma_commit(b_empty + st);
                }
                umma_commit(acc_full + acc_i);
            };
            for (int64_t it = 0; it < my_tiles; ++it) {
                mbar_wait(a_loaded, it & 1);            // dz2 tile in smem
                fb_trace(p, 1, tn, 1);
                tc_fence_after();
                gemm(0, 256);                           // dz1 pre-mask
                fb_trace(p, 1, tn, 2);
                mbar_wait(a_ready, n_ready & 1);        // dz1 tile in smem
                fb_trace(p, 1, tn, 3);
                ++n_ready;
                tc_fence_after();
                gemm(1, 256);                           // dz0 pre-mask
                fb_trace(p, 1, tn, 4);
                mbar_wait(a_ready, n_ready & 1);        // dz0 tile in smem
                fb_trace(p, 1, tn, 5);
                ++n_ready;
                tc_fence_after();
                for (int j = 0; j < n_nt; ++j) {
                    gemm(j & 1, j + 1 < n_nt ? 256 : last_n);
                    fb_trace(p, 1, tn, 10 + j);
                }
                umma_commit(a_free);
            }
        }
        __syncwarp();
    } else {
        // ======================= row workers =======================
        const int wq = warp & 3, half = warp >> 2;      // TMEM lane quarter, column half
        const int r = wq * 32 + lane;
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        uint32_t nf[2] = {0, 0}, gs = 0;
        int tn = 0;
        auto wait_acc = [&](int acc_i) {
            mbar_wait(acc_full + acc_i, nf[acc_i] & 1);
            ++nf[acc_i];
            tc_fence_after();
        };
        auto release_acc = [&](int acc_i) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_free + acc_i);
        };
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t tile = blockIdx.x + it * gridDim.x;
            const int row0 = (int)(tile * FB_TILE);
            const int64_t row = tile * FB_TILE + r;
            const bool row_ok = row < p.M;
            // ---- ReLU masks of this thread's 128 columns as bits: all 16 loads of a layer in flight at once (one
            //      DRAM round trip), layer 0 while the dz2 tile loads, layer 1 behind the second GEMM
            uint32_t mb[2][4];
            auto fetch_mask_bits = [&](int layer) {
                const uint4 *src = reinterpret_cast<const uint4 *>((layer == 0 ? p.h1 : p.h0) + row * FB_HID + half * 128);
                uint4 mk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) mk[i] = row_ok ? __ldg(src + i) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    uint32_t bits = 0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t mw[4] = {mk[w * 4 + q].x, mk[w * 4 + q].y, mk[w * 4 + q].z, mk[w * 4 + q].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {   // bf16 "> 0": 1 <= bits <= 0x7fff per half word
                            bits |= (((mw[j] & 0xffffu) - 1u) < 0x7fffu ? 1u : 0u) << (q * 8 + 2 * j);
                            bits |= (((mw[j] >> 16) - 1u) < 0x7fffu ? 1u : 0u) << (q * 8 + 2 * j + 1);
                        }
                    }
                    mb[layer][w] = bits;
                }
            };
            fetch_mask_bits(0);
            // ---- two masked layers: acc -> (* [h > 0]) -> bf16 -> A tile (next operand), tensor-stored from there
            if (threadIdx.x == 0) fb_trace(p, 0, tn, 19);
#pragma unroll
            for (int layer = 0; layer < 2; ++layer) {
                wait_acc(layer);
                if (threadIdx.x == 0) fb_trace(p, 0, tn, 20 + 2 * layer);
                if (layer == 1) {   // the dz1 tensor stores must have read the tile before it is overwritten
                    if (threadIdx.x == 0) bulk_wait_read<0>();
                    named_bar_sync(1, FB_WORKERS);
                }
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int c0 = half * 128 + w * 32;
                    uint32_t v[32];
                    tmem_ld32(tmem + layer * 256 + lane_off + c0, v);
                    tmem_ld_wait();
                    const uint32_t bits = mb[layer][w];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float g[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) g[e] = (bits >> (q * 8 + e)) & 1u ? __uint_as_float(v[q * 8 + e]) : 0.f;
                        *reinterpret_cast<uint4 *>(a_tile + (c0 >> 6) * (FB_TILE * 128) + swz128(r, ((c0 & 63) >> 3) + q)) = float8_to_bf16(g);
                    }
                }
                release_acc(layer);
                fence_proxy_async();
                named_bar_sync(1, FB_WORKERS);
                if (threadIdx.x == 0) {
                    fb_trace(p, 0, tn, 21 + 2 * layer);
                    mbar_arrive(a_ready);
                    const TensorMap *tm = layer == 0 ? &p.tm_dz1 : &p.tm_dz0;
                    for (int c = 0; c < 4; ++c) tma_store_2d(tm, c * 64, row0, smem_u32(a_tile + c * (FB_TILE * 128)));
                    bulk_commit();
                }
                if (layer == 0) fetch_mask_bits(1);
            }
            // ---- dfeat: 64-column boxes staged in shared memory, written with tensor stores
#pragma unroll 1
            for (int j = 0; j < n_nt; ++j) {
                const int acc_i = j & 1;
                const int ncols = j + 1 < n_nt ? 256 : last_n;
                wait_acc(acc_i);
                if (threadIdx.x == 0) fb_trace(p, 0, tn, 30 + j);
#pragma unroll 1
                for (int c0 = 0; c0 < ncols; c0 += 64, ++gs) {
                    uint8_t *slot = stage + (gs % FB_NS) * FB_S_BYTES;
                    uint32_t v[32];
                    tmem_ld32(tmem + acc_i * 256 + lane_off + c0 + half * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float g[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) g[e] = __uint_as_float(v[q * 8 + e]);
                        *reinterpret_cast<uint4 *>(slot + swz128(r, half * 4 + q)) = float8_to_bf16(g);
                    }
                    if (c0 + 64 >= ncols) release_acc(acc_i);
                    fence_proxy_async();
                    named_bar_sync(1, FB_WORKERS);
                    if (threadIdx.x == 0) {
                        tma_store_2d(&p.tm_dfeat, j * 256 + c0, row0, smem_u32(slot));
                        bulk_commit();
                        bulk_wait_read<FB_NS - 2>();   // slot of box gs+2 (= gs-2) is free once thread 0 reaches the next barrier
                    }
                }
                if (threadIdx.x == 0) fb_trace(p, 0, tn, 50 + j);
            }
            if (threadIdx.x == 0) {
                if (p.kp < 64 * (FB_NS - 1)) bulk_wait_read<0>();   // else the dz0 stores are older than FB_NS-2 committed groups
                mbar_arrive(a_free);
            }
        }
        if (threadIdx.x == 0) bulk_wait<0>();
    }
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace svr

using namespace svr;

static long long *g_fb_trace = nullptr;
// debug: device buffer of 3 * 1024 * 2 int64 that the next fused backward launches fill (block 0), or null to stop
extern "C" int svr_debug_fb_trace(void *buf) {
    g_fb_trace = (long long *)buf;
    return 0;
}

extern "C" int svr_decoder_bwd_fused(const uint16_t *dz2, const uint16_t *h1, const uint16_t *h0, const void *w2t_img, const void *w1t_img,
                                     const void *w0pt_img, int64_t M, int kp, uint16_t *dz1, uint16_t *dz0, uint16_t *dfeat, void *stream) {
    SVR_REQUIRE(dz2 && h1 && h0 && w2t_img && w1t_img && w0pt_img && dz1 && dz0 && dfeat, "decoder_bwd_fused: null pointer");
    SVR_REQUIRE(kp > 0 && kp % 64 == 0, "decoder_bwd_fused: KP must be a positive multiple of 64");
    if (M == 0) return 0;
    static DeviceOnce once;
    int dev;
    if (once.needed(dev)) {
        SVR_CUDA(cudaFuncSetAttribute(fused_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM));
        once.done(dev);
    }
    FbParams p;
    if (make_tmap_bf16_sw128(&p.tm_dz2, dz2, M, FB_HID, FB_HID, FB_TILE)) return -1;
    if (make_tmap_bf16_sw128(&p.tm_dz1, dz1, M, FB_HID, FB_HID, FB_TILE)) return -1;
    if (make_tmap_bf16_sw128(&p.tm_dz0, dz0, M, FB_HID, FB_HID, FB_TILE)) return -1;
    if (make_tmap_bf16_sw128(&p.tm_dfeat, dfeat, M, kp, kp, FB_TILE)) return -1;
    p.h1 = (const __nv_bfloat16 *)h1;
    p.h0 = (const __nv_bfloat16 *)h0;
    p.w2t_img = (const uint8_t *)w2t_img;
    p.w1t_img = (const uint8_t *)w1t_img;
    p.w0pt_img = (const uint8_t *)w0pt_img;
    p.M = M;
    p.kp = kp;
    p.trace = g_fb_trace;
    SVR_REQUIRE(M < (int64_t)1 << 31, "decoder_bwd_fused: M must fit 31 bits (TMA coordinates)");
    const int64_t n_tiles = ceil_div<int64_t>(M, FB_TILE);
    int grid = sm_count();
    if (n_tiles < grid) grid = (int)n_tiles;
    fused_bwd_kernel<<<grid, FB_THREADS, FB_SMEM, as_stream(stream)>>>(p, n_tiles);
    SVR_LAUNCH_CHECK();
    return 0;
}
