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

// small POD passed by value to kernels
struct Dims3 {
    int d[3];
};

struct Taps {
    float t[SVR_MAX_TAPS];
    int k;
};

}  // namespace svr
