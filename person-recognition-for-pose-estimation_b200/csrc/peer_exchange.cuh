// Peer-memory exchange for the gallery-sharded match: one process per GPU, every rank owns one
// exchange buffer that all ranks of the box map through CUDA IPC (NVLink / NVSwitch peer access).  Kernels of the
// match chain write straight into their peers' buffers (plain P2P stores), publish a step number with a
// system-scope release store, and the consumer spins on its own copy of the flag: no NCCL call, no host round trip,
// CUDA-graph capturable.
//
//   [header 1 KB]  step counter + CTA arrival counters (local use only)
//   [flags  1 KB]  probe_flag[2][16], key_flag[2][16]      written by the peers (their step number)
//   probes_f32  [2][G*M][512] fp32   normalised probes of every rank (exact re-score operand)
//   probes_bf16 [2][G*M][512] bf16   the same rounded to bf16 (tcgen05 A operand, read by TMA)
//   keys_in     [2][G][M]     u64    packed (similarity, id) winners of MY probes, one row per source rank
// Everything indexed [2] is double-buffered by the parity of the step counter.  Why that is enough: a rank can
// start step k+2 only after its step k+1 finished, which needed every peer's keys of step k+1, which a peer
// sends after ITS step k is completely over — so nobody still reads parity (k & 1) when it is overwritten.
#pragma once

#include "spp_common.cuh"

namespace spp {

constexpr int kPeerDim = 512;
constexpr size_t kPeerHeaderBytes = 1024;
constexpr size_t kPeerFlagBytes = 1024;

struct PeerHeader {
    unsigned step;         // completed steps of this rank (all ranks advance in lock step)
    unsigned push_count;   // arrival counter: CTAs of the probe push
    unsigned key_count;    // arrival counter: CTAs of the finalize / key push
    unsigned done_count;   // arrival counter: CTAs of the reduce
};

struct PeerLayout {
    size_t off_flags, off_f32, off_bf16, off_keys, bytes;
    size_t f32_stride, bf16_stride, keys_stride;     // bytes per parity
};

SPP_HD static inline PeerLayout peer_layout(int world, int m_local) {
    PeerLayout l;
    const size_t rows = (size_t)world * (size_t)m_local;
    l.off_flags = kPeerHeaderBytes;
    l.off_f32 = l.off_flags + kPeerFlagBytes;
    l.f32_stride = align_up(rows * kPeerDim * 4, 1024);
    l.off_bf16 = l.off_f32 + 2 * l.f32_stride;
    l.bf16_stride = align_up(rows * kPeerDim * 2, 1024);
    l.off_keys = l.off_bf16 + 2 * l.bf16_stride;
    l.keys_stride = align_up(rows * 8, 1024);
    l.bytes = l.off_keys + 2 * l.keys_stride;
    return l;
}

struct PeerPtrs {
    unsigned char *buf[SPP_MAX_PEERS];
    int world, rank, m_local;
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned *peer_probe_flag(unsigned char *buf, int parity, int src) {
    return reinterpret_cast<unsigned *>(buf + kPeerHeaderBytes) + parity * SPP_MAX_PEERS + src;
}
__device__ __forceinline__ unsigned *peer_key_flag(unsigned char *buf, int parity, int src) {
    return reinterpret_cast<unsigned *>(buf + kPeerHeaderBytes) + (2 + parity) * SPP_MAX_PEERS + src;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Spin until *flag reaches `want` (step numbers only grow).  Bounded: a peer that never arrives (crashed rank, ranks
// that disagree on the number of steps) turns into a trap — a CUDA error on this rank — after kPeerTimeoutNs.
constexpr unsigned long long kPeerTimeoutNs = 60ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ void peer_wait_flag(const unsigned *flag, unsigned want) {
    if ((int)(ld_acquire_sys(flag) - want) >= 0) return;
    const unsigned long long t0 = global_timer_ns();
    while ((int)(ld_acquire_sys(flag) - want) < 0) {
        __nanosleep(64);
        if (global_timer_ns() - t0 > kPeerTimeoutNs) __trap();
    }
}
#endif

}  // namespace spp
