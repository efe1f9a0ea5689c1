// Error plumbing and device queries of the svr_b200 C-ABI library.
#include "common.cuh"
#include <cuda.h>
#include <cstring>
#include <mutex>

namespace svr {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

bool DeviceOnce::needed(int &dev) {
    dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 256) return true;   // unknown device: always (re)apply
    return ((__atomic_load_n(&bits[dev >> 6], __ATOMIC_ACQUIRE) >> (dev & 63)) & 1ull) == 0;
}

void DeviceOnce::done(int dev) {
    if (dev >= 0 && dev < 256) __atomic_fetch_or(&bits[dev >> 6], 1ull << (dev & 63), __ATOMIC_RELEASE);
}

int scratch_alloc(void **ptr, size_t bytes, cudaStream_t st) {
    // Freed scratch stays in the pool (release threshold = max): with the default threshold a pool returns its memory
    // to the OS at every synchronisation and the next allocation re-maps it (milliseconds).  The pool is ours, one per
    // device, so the host application's default pool keeps its own settings.
    static cudaMemPool_t pools[256] = {};
    static std::mutex mu;
    int dev = 0;
    SVR_CUDA(cudaGetDevice(&dev));
    SVR_REQUIRE(dev >= 0 && dev < 256, "scratch_alloc: device index %d out of range", dev);
    cudaMemPool_t pool;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!pools[dev]) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            SVR_CUDA(cudaMemPoolCreate(&pools[dev], &props));
            uint64_t keep = ~0ull;
            SVR_CUDA(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
        }
        pool = pools[dev];
    }
    SVR_CUDA(cudaMallocFromPoolAsync(ptr, bytes, pool, st));
    return 0;
}

typedef CUresult (*TmapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);

static TmapEncodeFn tmap_encoder() {   // resolved through the runtime: no link-time libcuda dependency
    static TmapEncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled is not available from the driver (%s)", cudaGetErrorString(e));
            return nullptr;
        }
        encode = (TmapEncodeFn)fn;
    }
    return encode;
}

int make_tmap_vol_bf16(TensorMap *out, const void *base, int B, int D, int H, int W, int C, int box_x) {
    static_assert(sizeof(TensorMap) == sizeof(CUtensorMap) && alignof(TensorMap) == alignof(CUtensorMap), "TensorMap layout");
    TmapEncodeFn encode = tmap_encoder();
    if (!encode) return -1;
    if (((uintptr_t)base & 15) || C % 64 || box_x < 1 || box_x > 256) {
        set_error("volume tensor map: base must be 16-byte aligned and C a multiple of 64 (C=%d)", C);
        return -1;
    }
    const cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
    const cuuint64_t gstride[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
    const cuuint32_t box[5] = {64, (cuuint32_t)box_x, 1, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode((CUtensorMap *)out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void *>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (5-D volume) failed with CUresult %d (B=%d D=%d H=%d W=%d C=%d box_x=%d)", (int)r, B, D, H, W, C, box_x);
        return -1;
    }
    return 0;
}

int make_tmap_bf16_sw128(TensorMap *out, const void *base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    static_assert(sizeof(TensorMap) == sizeof(CUtensorMap) && alignof(TensorMap) == alignof(CUtensorMap), "TensorMap layout");
    TmapEncodeFn encode = tmap_encoder();
    if (!encode) return -1;
    if (((uintptr_t)base & 15) || (ld * 2) % 16 || cols % 64 || box_rows < 1 || box_rows > 256) {
        set_error("tensor map: base must be 16-byte aligned, ld*2 a multiple of 16, cols a multiple of 64 (ld=%lld cols=%lld)", (long long)ld,
                  (long long)cols);
        return -1;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode((CUtensorMap *)out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r, (long long)rows, (long long)cols,
                  (long long)ld);
        return -1;
    }
    return 0;
}

}  // namespace svr

extern "C" {

int svr_abi_version(void) { return SVR_ABI_VERSION; }

const char *svr_last_error(void) { return svr::g_err; }

int svr_device_info(char *name_host, int *sm_count_host, int *cc_major_host, int *cc_minor_host) {
    int dev = 0;
    SVR_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    SVR_CUDA(cudaGetDeviceProperties(&p, dev));
    if (name_host) {
        strncpy(name_host, p.name, 255);
        name_host[255] = 0;
    }
    if (sm_count_host) *sm_count_host = p.multiProcessorCount;
    if (cc_major_host) *cc_major_host = p.major;
    if (cc_minor_host) *cc_minor_host = p.minor;
    SVR_REQUIRE(p.major == 10, "svr_b200 needs an sm_100-class device, found compute capability %d.%d (%s)", p.major,
                p.minor, p.name);
    return 0;
}
}
