// Error plumbing and device queries of the svr_b200 C-ABI library.
#include "common.cuh"
#include <cstring>

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
