// Shared host/device helpers for the svr_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#include "../../include/svr_b200.h"

namespace svr {

void set_error(const char *fmt, ...);

// status helpers --------------------------------------------------------------------------------
#define SVR_REQUIRE(cond, ...)                      \
    do {                                            \
        if (!(cond)) {                              \
            ::svr::set_error(__VA_ARGS__);          \
            return -1;                              \
        }                                           \
    } while (0)

#define SVR_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            ::svr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return (int)e__;                                                                    \
        }                                                                                       \
    } while (0)

#define SVR_LAUNCH_CHECK()                                                                      \
    do {                                                                                        \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess) {                                                               \
            ::svr::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return (int)e__;                                                                    \
        }                                                                                       \
    } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
static inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

int sm_count();   // cached per device

// One-time per-DEVICE initialisation (cudaFuncSetAttribute is a per-device setting, a process may drive several
// GPUs): `if (once.needed(dev)) { ...set attributes...; once.done(dev); }`.  Thread-safe; setting an attribute
// twice from two racing threads is harmless.
struct DeviceOnce {
    unsigned long long bits[4] = {0, 0, 0, 0};   // up to 256 devices
    bool needed(int &dev);
    void done(int dev);
};

// stream-ordered scratch from a PRIVATE per-device memory pool (the host application's default pool is left alone)
int scratch_alloc(void **ptr, size_t bytes, cudaStream_t st);

// 128-byte opaque TMA descriptor (same layout and alignment as the driver's CUtensorMap)
struct alignas(64) TensorMap {
    uint64_t opaque[16];
};
// row-major bf16 matrix (rows x cols, leading dimension ld elements) tiled in boxes of box_rows x 64 columns
// with the 128-byte swizzle the UMMA K-major operand layout uses.  0 on success.
int make_tmap_bf16_sw128(TensorMap *out, const void *base, int64_t rows, int64_t cols, int64_t ld, int box_rows);
// channel-last bf16 volume (B, D, H, W, C), C % 64 == 0, as a 5-D tensor (C, W, H, D, B) tiled in boxes of 64 channels x box_x
// voxels of one x-line, 128-byte swizzle (rows of an MN-major UMMA operand); out-of-volume voxels read as zeros.  0 on success.
int make_tmap_vol_bf16(TensorMap *out, const void *base, int B, int D, int H, int W, int C, int box_x);

// small POD passed by value to kernels
struct Dims3 {
    int d[3];
};

struct Taps {
    float t[SVR_MAX_TAPS];
    int k;
};

}  // namespace svr
