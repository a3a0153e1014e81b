// C-ABI plumbing shared by all entry points: thread-local error string, device queries.
#include "spp_common.cuh"

#include <cstring>

namespace spp {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        set_error("cudaGetDevice failed: no CUDA device (libspp has no CPU fallback)");
        return -1;
    }
    if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        set_error("cudaDeviceGetAttribute(MultiProcessorCount) failed");
        return -1;
    }
    if (dev >= 0 && dev < 64) cached[dev] = n;
    return n;
}

}  // namespace spp

extern "C" int spp_abi_version(void) { return 1; }
extern "C" const char *spp_last_error(void) { return spp::g_error; }
extern "C" int spp_device_sm_count(void) { return spp::sm_count(); }
