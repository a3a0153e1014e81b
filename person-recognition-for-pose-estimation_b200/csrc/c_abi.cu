// C-ABI plumbing shared by all entry points: thread-local error string, device queries.
#include "spp_common.cuh"

#include <atomic>
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

namespace spp {
static std::atomic<int> g_limits[3] = {{0}, {0}, {0}};
int launch_limit(int which) { return (which >= 0 && which < 3) ? g_limits[which].load(std::memory_order_relaxed) : 0; }
}  // namespace spp

extern "C" int spp_set_launch_limit(int which, int max_ctas) {
    if (which < 0 || which >= 3) return -1;
    if (max_ctas < 0) return spp::g_limits[which].load();
    return spp::g_limits[which].exchange(max_ctas);
}

extern "C" int spp_abi_version(void) { return 2; }
extern "C" const char *spp_last_error(void) { return spp::g_error; }
extern "C" int spp_device_sm_count(void) { return spp::sm_count(); }

// ---- exchange buffers shared between the per-GPU processes of one box (CUDA IPC) ----------------
extern "C" int spp_peer_alloc(size_t bytes, void **dev_ptr, unsigned char *handle_out) {
    SPP_CHECK_ARG(dev_ptr && handle_out && bytes > 0, "peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == SPP_IPC_HANDLE_BYTES, "IPC handle size");
    void *p = nullptr;
    SPP_CHECK_CUDA(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        spp::set_error("peer_alloc: %s", cudaGetErrorString(e));
        return SPP_ERR_CUDA;
    }
    memcpy(handle_out, &h, sizeof(h));
    *dev_ptr = p;
    return SPP_OK;
}

extern "C" int spp_peer_open(const unsigned char *handle, void **dev_ptr) {
    SPP_CHECK_ARG(handle && dev_ptr, "peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    SPP_CHECK_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SPP_OK;
}

extern "C" int spp_peer_close(void *dev_ptr) {
    if (dev_ptr) SPP_CHECK_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return SPP_OK;
}

extern "C" int spp_peer_free(void *dev_ptr) {
    if (dev_ptr) SPP_CHECK_CUDA(cudaFree(dev_ptr));
    return SPP_OK;
}

extern "C" int spp_peer_can_access(int other_device) {
    int dev = 0, ok = 0;
    SPP_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev == other_device) return 1;
    SPP_CHECK_CUDA(cudaDeviceCanAccessPeer(&ok, dev, other_device));
    return ok;
}
