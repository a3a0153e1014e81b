// Shared helpers for the libspp kernels: error reporting behind the C ABI, PTX wrappers for
// mbarrier / bulk-TMA, warp reductions.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/spp.h"

namespace spp {

void set_error(const char *fmt, ...);

#define SPP_CHECK_ARG(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            ::spp::set_error(__VA_ARGS__);       \
            return SPP_ERR_INVALID;              \
        }                                        \
    } while (0)

#define SPP_CHECK_CUDA(expr)                                                                   \
    do {                                                                                       \
        cudaError_t err__ = (expr);                                                            \
        if (err__ != cudaSuccess) {                                                            \
            ::spp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
            return SPP_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define SPP_CHECK_LAUNCH() SPP_CHECK_CUDA(cudaGetLastError())

int sm_count();
int launch_limit(int which);      // spp_set_launch_limit: 0 = heatmap decode CTAs, 1 = match GEMM CTAs (0 = no limit), 2 = CTA slots the crop leaves free

#ifdef __CUDACC__
#define SPP_HD __host__ __device__
#else
#define SPP_HD
#endif

SPP_HD static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

#ifdef __CUDACC__

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug turns into a trap (reported as a CUDA error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}

// ---- bulk TMA (1-D): global -> shared, completion on an mbarrier ------------------------------
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- warp reductions ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
// arg-max with "first maximum wins" (lowest flat index on ties) — np.argmax / torch.max semantics
__device__ __forceinline__ void warp_argmax(float &v, int &idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(FULL, v, o);
        int oi = __shfl_xor_sync(FULL, idx, o);
        if (ov > v || (ov == v && oi < idx)) {
            v = ov;
            idx = oi;
        }
    }
}

// monotone float -> signed int key (a < b  <=>  key(a) < key(b)), and back
__device__ __host__ __forceinline__ int32_t float_to_ordered(float f) {
    int32_t i;
#ifdef __CUDA_ARCH__
    i = __float_as_int(f);
#else
    union { float f; int32_t i; } u; u.f = f; i = u.i;
#endif
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __host__ __forceinline__ float ordered_to_float(int32_t i) {
    i = i ^ ((i >> 31) & 0x7fffffff);
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    union { float f; int32_t i; } u; u.i = i; return u.f;
#endif
}

#endif  // __CUDACC__

}  // namespace spp
