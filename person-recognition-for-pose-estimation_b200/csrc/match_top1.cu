// Embedding L2-normalisation + cosine gallery match with top-1 and threshold gate.
//
// Replaces:
//   Backbone.forward tail   libs/net_adaface.py:334-337            (spp_l2_normalize mode 0)
//   l2_norm / cosine        libs/head_adaface.py:39-42, 79-81
//   match + top-1           training/lightning/face_recognition/module.py:136-145
//
// The one dense contraction of the path, [M,512] x [512,N], runs on the 5th-gen tensor cores:
//   * probes are normalised in fp32 and rounded to bf16 (A operand), the gallery is bf16 (B operand),
//     both K-major; operands reach shared memory by 2-D TMA with the 128-byte swizzle that the UMMA
//     shared-memory descriptors expect;
//   * each CTA keeps its 128-probe A tile (128 x 512 bf16 = 128 KB) resident for its whole life and
//     streams 256-identity B tiles through a 3-stage mbarrier ring (32 KB per k-block);
//   * one elected thread issues tcgen05.mma (M=128, N=256, K=16) into one of two 256-column TMEM
//     accumulators, so the epilogue of tile t overlaps the MMAs of tile t+1;
//   * the epilogue reads TMEM with tcgen05.ld (thread = probe row, registers = gallery columns) and
//     keeps a running per-row top-2 (value, index); the [M,N] score matrix never exists in memory.
// bf16 products cannot reproduce an fp32 arg-max bit for bit, so the (few) surviving candidates of
// every row are re-scored with an exact fp32 dot product in match_finalize_kernel, which also applies
// the gate and packs the (value, index) key for the multi-GPU top-1 reduction.
#include "spp_common.cuh"

#include <cstdlib>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cmath>
#include <mutex>

namespace spp {
namespace {

constexpr int kDim = 512;
constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kKBlocks = kDim / BK;              // 8
constexpr int kStages = 3;
constexpr int kABytes = BM * BK * 2;             // 16 KB per k-block
constexpr int kBBytes = BN * BK * 2;             // 32 KB per stage
constexpr int kGemmThreads = 192;                // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int kTmemCols = 512;
constexpr size_t kGemmSmem = 1024 /*align slack*/ + (size_t)kKBlocks * kABytes + (size_t)kStages * kBBytes + 256;

struct Cand {
    float v;
    int i;
};

// ------------------------------------------------------------------------------------------------
// normalisation / conversion kernels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2_normalize_kernel(const float *__restrict__ x, int m, int dim, int mode, float eps,
                                                           float *__restrict__ out, float *__restrict__ norm,
                                                           __nv_bfloat16 *__restrict__ out_bf16) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= m) return;
    const float *xr = x + (size_t)row * dim;
    float ss = 0.f;
    for (int i = lane; i < dim; i += 32) {
        const float v = xr[i];
        ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    const float den = mode == 1 ? fmaxf(nrm, eps) : nrm;
    if (lane == 0 && norm) norm[row] = nrm;
    for (int i = lane; i < dim; i += 32) {
        const float v = __fdiv_rn(xr[i], den);
        if (out) out[(size_t)row * dim + i] = v;
        if (out_bf16) out_bf16[(size_t)row * dim + i] = __float2bfloat16_rn(v);
    }
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float *__restrict__ x, size_t count, __nv_bfloat16 *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) out[i] = __float2bfloat16_rn(x[i]);
}

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMA PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, 128-byte swizzle, 8-row atoms 1024 B apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);            // start address
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset
    d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------------
// GEMM + fused row top-2
// ------------------------------------------------------------------------------------------------
struct GemmParams {
    int m, n;
    int tiles_n, nsplit;   // every M-tile is cut into nsplit chunks of gallery tiles; work item = (M-tile, chunk)
    int items;             // m_tiles * nsplit, distributed round-robin over the persistent CTAs
    Cand *part;            // [m_tiles*BM, nsplit, 2]
};

__global__ void __launch_bounds__(kGemmThreads, 1)
match_gemm_top2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams prm) {
    extern __shared__ unsigned char gemm_smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *smem_a = smem;                                         // [8][128 x 64] bf16, SW128
    unsigned char *smem_b = smem + (size_t)kKBlocks * kABytes;            // [stages][256 x 64] bf16, SW128
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_b + (size_t)kStages * kBBytes);
    uint64_t *full = bars;                  // [kStages]  B stage landed
    uint64_t *empty = bars + kStages;       // [kStages]  B stage consumed by the MMAs
    uint64_t *a_full = bars + 2 * kStages;  // [1]        probe tile landed
    uint64_t *a_empty = a_full + 1;         // [1]        probe tile no longer read (item finished)
    uint64_t *t_full = a_empty + 1;         // [2]        accumulator complete
    uint64_t *t_empty = t_full + 2;         // [2]        accumulator drained by the epilogue
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = prm.tiles_n / prm.nsplit, rem = prm.tiles_n - per * prm.nsplit;
    // chunk sp of an M-tile covers gallery tiles [nt0, nt0 + ntiles)
    auto chunk_range = [&](int sp, int &nt0, int &ntiles) {
        nt0 = sp * per + (sp < rem ? sp : rem);
        ntiles = per + (sp < rem ? 1 : 0);
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(&t_full[a], 1);
            mbar_init(&t_empty[a], 4);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, it = 0;
            for (int item = blockIdx.x; item < prm.items; item += gridDim.x, ++it) {
                const int mt = item / prm.nsplit, sp = item - mt * prm.nsplit;
                int nt0, ntiles;
                chunk_range(sp, nt0, ntiles);
                mbar_wait(a_empty, (it & 1) ^ 1);            // previous item's MMAs are done with the probe tile
                mbar_arrive_expect_tx(a_full, kKBlocks * kABytes);
                for (int kb = 0; kb < kKBlocks; ++kb) tma_load_2d(smem_a + (size_t)kb * kABytes, &tmap_a, a_full, kb * BK, mt * BM);
                for (int t = 0; t < ntiles; ++t) {
                    const int n0 = (nt0 + t) * BN;
                    for (int kb = 0; kb < kKBlocks; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&full[stage], kBBytes);
                        tma_load_2d(smem_b + (size_t)stage * kBBytes, &tmap_b, &full[stage], kb * BK, n0);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor: fp32 accumulate, bf16 x bf16, both K-major, M=128, N=256
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint32_t a_addr = smem_u32(smem_a), b_addr = smem_u32(smem_b);
            int stage = 0;
            uint32_t phase = 0, it = 0, tcount = 0;          // tcount: tiles issued by this CTA (accumulator ring)
            for (int item = blockIdx.x; item < prm.items; item += gridDim.x, ++it) {
                const int sp = item % prm.nsplit;
                int nt0, ntiles;
                chunk_range(sp, nt0, ntiles);
                mbar_wait(a_full, it & 1);
                for (int t = 0; t < ntiles; ++t, ++tcount) {
                    const int acc = tcount & 1;
                    const uint32_t acc_phase = (tcount >> 1) & 1;
                    mbar_wait(&t_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_addr = tmem_base + (uint32_t)(acc * BN);
                    for (int kb = 0; kb < kKBlocks; ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint64_t ad = umma_desc_sw128(a_addr + kb * kABytes + k * 32);
                            const uint64_t bd = umma_desc_sw128(b_addr + stage * kBBytes + k * 32);
                            umma_bf16(d_addr, ad, bd, idesc, (kb | k) ? 1u : 0u);
                        }
                        umma_commit(&empty[stage]);          // frees the B stage once these MMAs have read it
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&t_full[acc]);               // accumulator complete
                }
                umma_commit(a_empty);                        // all MMAs of this item have read the probe tile
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> running top-2 per probe row =====
        const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
        const int row_in_tile = quarter * 32 + lane;
        uint32_t tcount = 0;
        for (int item = blockIdx.x; item < prm.items; item += gridDim.x) {
            const int mt = item / prm.nsplit, sp = item - mt * prm.nsplit;
            int nt0, ntiles;
            chunk_range(sp, nt0, ntiles);
            float v1 = -INFINITY, v2 = -INFINITY;
            int i1 = 0x7fffffff, i2 = 0x7fffffff;
            for (int t = 0; t < ntiles; ++t, ++tcount) {
                const int acc = tcount & 1;
                const uint32_t acc_phase = (tcount >> 1) & 1;
                mbar_wait(&t_full[acc], acc_phase);
                tc_fence_after();
                const int n0 = (nt0 + t) * BN;
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    float v[32];
                    tmem_ld32(taddr + c * 32, v);
                    float mx = v[0];
#pragma unroll
                    for (int j = 1; j < 32; ++j) mx = fmaxf(mx, v[j]);
                    if (mx > v2) {                          // rare after the first tiles
                        const int nb = n0 + c * 32;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float x = v[j];
                            const int n = nb + j;
                            if (n < prm.n) {
                                if (x > v1) { v2 = v1; i2 = i1; v1 = x; i1 = n; }
                                else if (x > v2) { v2 = x; i2 = n; }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_empty[acc]);
            }
            Cand *o = prm.part + ((size_t)(mt * BM + row_in_tile) * prm.nsplit + sp) * 2;
            o[0] = Cand{v1, i1};
            o[1] = Cand{v2, i2};
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// CUDA-core fp32 version of the same candidate search — device-side cross-check for the tests
// (spp_match_top1 never dispatches to it).
__global__ void __launch_bounds__(128) match_simt_top2_kernel(const float *__restrict__ qn, const __nv_bfloat16 *__restrict__ gal,
                                                              int m, int n, int nsplit, Cand *part) {
    __shared__ float q[kDim];
    const int row = blockIdx.x, sp = blockIdx.y;
    for (int i = threadIdx.x; i < kDim; i += blockDim.x) q[i] = qn[(size_t)row * kDim + i];
    __syncthreads();
    const int per = (n + nsplit - 1) / nsplit;
    const int lo = sp * per, hi = min(n, lo + per);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float v1 = -INFINITY, v2 = -INFINITY;
    int i1 = 0x7fffffff, i2 = 0x7fffffff;
    for (int g = lo + warp; g < hi; g += 4) {
        const __nv_bfloat16 *gr = gal + (size_t)g * kDim;
        float s = 0.f;
        for (int i = lane; i < kDim; i += 32) s = fmaf(q[i], __bfloat162float(gr[i]), s);
        s = warp_sum(s);
        if (s > v1 || (s == v1 && g < i1)) { v2 = v1; i2 = i1; v1 = s; i1 = g; }
        else if (s > v2 || (s == v2 && g < i2)) { v2 = s; i2 = g; }
    }
    __shared__ Cand sc[4][2];
    if (lane == 0) { sc[warp][0] = Cand{v1, i1}; sc[warp][1] = Cand{v2, i2}; }
    __syncthreads();
    if (threadIdx.x == 0) {
        Cand b1{-INFINITY, 0x7fffffff}, b2{-INFINITY, 0x7fffffff};
        for (int w = 0; w < 4; ++w)
            for (int k = 0; k < 2; ++k) {
                const Cand c = sc[w][k];
                if (c.v > b1.v || (c.v == b1.v && c.i < b1.i)) { b2 = b1; b1 = c; }
                else if (c.v > b2.v || (c.v == b2.v && c.i < b2.i)) { b2 = c; }
            }
        Cand *o = part + ((size_t)row * nsplit + sp) * 2;
        o[0] = b1;
        o[1] = b2;
    }
}

// Exact fp32 re-score of the surviving candidates, gate, key packing.  One warp per probe.
// A candidate whose bf16 score is more than kPrune below the row's best bf16 score cannot be the fp32
// arg-max: the probe is rounded to bf16 (unit roundoff 2^-8), the gallery values are exact, so
// |s_bf16 - s_fp32| <= 2^-8 * sum|q_i g_i| <= 2^-8 for unit vectors, and kPrune = 0.01 > 2 * 2^-8.
constexpr float kPrune = 0.01f;

__global__ void __launch_bounds__(256) match_finalize_kernel(const float *__restrict__ qn, const __nv_bfloat16 *__restrict__ gal,
                                                             const Cand *__restrict__ part, int m, int n, int ncand,
                                                             float threshold, int id_offset, int *out_id, float *out_sim,
                                                             unsigned long long *out_key) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= m) return;
    // lane owns 16 contiguous dimensions: the probe is 4 x 128-bit loads, a gallery row 2 x 128-bit loads per lane
    constexpr int kPer = kDim / 32;                  // 16
    float q[kPer];
    {
        const float4 *qp = reinterpret_cast<const float4 *>(qn + (size_t)row * kDim + lane * kPer);
#pragma unroll
        for (int i = 0; i < kPer / 4; ++i) {
            const float4 v = __ldg(qp + i);
            q[4 * i] = v.x; q[4 * i + 1] = v.y; q[4 * i + 2] = v.z; q[4 * i + 3] = v.w;
        }
    }
    const Cand *c = part + (size_t)row * ncand;
    float vmax = -INFINITY;
    for (int k = lane; k < ncand; k += 32) {
        const Cand x = c[k];
        if (x.i >= 0 && x.i < n) vmax = fmaxf(vmax, x.v);
    }
    vmax = warp_max(vmax);
    float best = -INFINITY;
    int bidx = 0x7fffffff;
    auto dot16 = [&](const uint4 &lo, const uint4 &hi) {
        const unsigned w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {                // bf16 -> fp32 is a 16-bit shift
            s = fmaf(q[2 * i], __uint_as_float(w[i] << 16), s);
            s = fmaf(q[2 * i + 1], __uint_as_float(w[i] & 0xffff0000u), s);
        }
        return s;
    };
    for (int k0 = 0; k0 < ncand; k0 += 32) {
        const int k = k0 + lane;
        Cand x{-INFINITY, -1};
        if (k < ncand) x = c[k];
        const bool need = x.i >= 0 && x.i < n && x.v >= vmax - kPrune;
        unsigned todo = __ballot_sync(FULL, need);
        while (todo) {
            // up to four candidates per round: their rows are all requested before the first dot product
            int g[4];
            uint4 lo[4], hi[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                g[u] = -1;
                if (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    g[u] = __shfl_sync(FULL, x.i, src);
                    const uint4 *gr = reinterpret_cast<const uint4 *>(gal + (size_t)g[u] * kDim + lane * kPer);
                    lo[u] = __ldg(gr);
                    hi[u] = __ldg(gr + 1);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (g[u] < 0) continue;              // warp-uniform
                const float s = warp_sum(dot16(lo[u], hi[u]));
                if (s > best || (s == best && g[u] < bidx)) { best = s; bidx = g[u]; }
            }
        }
    }
    if (lane == 0) {
        const bool found = bidx != 0x7fffffff;
        const int gid = found ? bidx + id_offset : -1;
        const bool pass = found && !(best < threshold);      // NaN threshold: no gate
        if (out_id) out_id[row] = pass ? gid : -1;
        if (out_sim) out_sim[row] = found ? best : -INFINITY;
        if (out_key) {
            const unsigned long long hi = (unsigned long long)(unsigned)float_to_ordered(found ? best : -INFINITY);
            out_key[row] = (hi << 32) | (unsigned long long)(0xffffffffu - (unsigned)(found ? gid : 0x7fffffff));
        }
    }
}

__global__ void match_unpack_kernel(const unsigned long long *keys, int m, float threshold, int *out_id, float *out_sim) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const unsigned long long k = keys[i];
    const float sim = ordered_to_float((int32_t)(unsigned)(k >> 32));
    const unsigned gid = 0xffffffffu - (unsigned)(k & 0xffffffffu);
    const bool found = gid != 0x7fffffffu;
    if (out_sim) out_sim[i] = sim;
    if (out_id) out_id[i] = (found && !(sim < threshold)) ? (int)gid : -1;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// [rows, 512] bf16 row-major -> tiles of {64 cols, box_rows rows}, 128-byte swizzle, zero fill out of bounds
int make_tmap(CUtensorMap *map, const void *base, int rows, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("match: cuTensorMapEncodeTiled is not available from the driver");
        return SPP_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)kDim, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)kDim * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("match: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return SPP_ERR_CUDA;
    }
    return SPP_OK;
}

struct MatchPlan {
    int m_tiles, tiles_n, nsplit, sms;
    size_t off_qn, off_qb, off_part, bytes;
};

MatchPlan plan_match(int m, int n) {
    MatchPlan p{};
    p.m_tiles = (m + BM - 1) / BM;
    p.tiles_n = (n + BN - 1) / BN;
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    // Work items = (M-tile, chunk of gallery tiles).  Small problems (fewer items than SMs even at one gallery
    // tile per item): one item per CTA and as many CTAs as there are SMs — every CTA pays a 128 KB probe-tile
    // load, but the kernel is over in ~1/6 of the time 8-tile chunks take (cfg2: 145 CTAs x 1-2 tiles instead of
    // 25 CTAs x 8; measured 72 us -> see profiles/README.md), and a short kernel is what the graph's other
    // branches need from it.  SPP_MATCH_CHUNK_TILES=<t> restores chunks of >= t tiles (profiling knob).
    // Large problems: the grid is one persistent CTA per SM and the chunk count is chosen so that
    // items ~ r * SMs (balanced rounds, <= 8 of them).
    static const int chunk_tiles = [] { const char *e = getenv("SPP_MATCH_CHUNK_TILES"); const int v = e ? atoi(e) : 1; return v < 1 ? 1 : v; }();
    const int want = (p.tiles_n + chunk_tiles - 1) / chunk_tiles > 0 ? (p.tiles_n + chunk_tiles - 1) / chunk_tiles : 1;
    int ns;
    if ((long long)p.m_tiles * want <= sms) {
        ns = want;
    } else if ((long long)p.m_tiles * p.tiles_n <= 2LL * sms || p.tiles_n <= 8 * (sms / p.m_tiles > 0 ? sms / p.m_tiles : 1)) {
        // at most a couple of tiles per SM: one round, chunks as even as the tile count allows
        ns = sms / p.m_tiles > 0 ? sms / p.m_tiles : 1;
        if (ns > want) ns = want;
    } else {
        ns = 1;
        double best = 0.0;
        const int cap = (p.tiles_n + 7) / 8 > 0 ? (p.tiles_n + 7) / 8 : 1;      // >= 8 tiles per chunk: the probe-tile reload stays < 7 %
        for (int r = 1; r <= 8; ++r) {
            int c = (int)((long long)sms * r / p.m_tiles);
            if (c < 1) continue;
            if (c > cap) c = cap;
            const long long items = (long long)p.m_tiles * c;
            const long long rounds = (items + sms - 1) / sms;
            const double util = (double)items / (double)(rounds * sms);
            if (util > best + 1e-9) { best = util; ns = c; }
            if (util >= 0.94) break;      // fewest rounds that balance well: every extra round reloads the probe tile
        }
    }
    p.nsplit = ns;
    p.sms = sms;
    size_t off = 0;
    p.off_qn = off;   off += align_up((size_t)m * kDim * 4, 1024);
    p.off_qb = off;   off += align_up((size_t)m * kDim * 2, 1024);
    p.off_part = off; off += align_up((size_t)p.m_tiles * BM * ns * 2 * sizeof(Cand), 1024);
    p.bytes = off;
    return p;
}

}  // namespace
}  // namespace spp

using namespace spp;

extern "C" int spp_l2_normalize(const float *x, int m, int dim, int mode, float eps, float *out, float *norm,
                                uint16_t *out_bf16, spp_stream_t stream) {
    if (m == 0) return SPP_OK;
    SPP_CHECK_ARG(x && m >= 0 && dim > 0, "l2_normalize: bad arguments");
    SPP_CHECK_ARG(mode == 0 || mode == 1, "l2_normalize: mode must be 0 (x/||x||) or 1 (F.normalize)");
    if (m == 0) return SPP_OK;
    l2_normalize_kernel<<<(m + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, m, dim, mode, eps, out, norm,
                                                                                    reinterpret_cast<__nv_bfloat16 *>(out_bf16));
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}

extern "C" int spp_f32_to_bf16(const float *x, size_t count, uint16_t *out, spp_stream_t stream) {
    SPP_CHECK_ARG(x && out, "f32_to_bf16: null pointer");
    if (count == 0) return SPP_OK;
    size_t blocks = (count + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    f32_to_bf16_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, count, reinterpret_cast<__nv_bfloat16 *>(out));
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}

extern "C" size_t spp_match_workspace_bytes(int m, int n, int dim) {
    if (m < 0 || n < 1 || dim != kDim) return 0;
    return plan_match(m > 0 ? m : 1, n).bytes;
}

static int match_common(const float *emb, const uint16_t *gallery, int m, int n, int dim, float threshold, int id_offset,
                        int *out_id, float *out_sim, unsigned long long *out_key, void *workspace, size_t workspace_bytes,
                        spp_stream_t stream, bool simt) {
    if (m == 0) return SPP_OK;
    SPP_CHECK_ARG(emb && gallery && workspace, "match_top1: null pointer");
    SPP_CHECK_ARG(dim == kDim, "match_top1: embedding dimension must be %d (got %d)", kDim, dim);
    SPP_CHECK_ARG(m >= 0 && n >= 1, "match_top1: bad m=%d n=%d", m, n);
    SPP_CHECK_ARG((reinterpret_cast<uintptr_t>(gallery) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 1023) == 0,
                  "match_top1: gallery must be 16-byte and workspace 1024-byte aligned");
    if (m == 0) return SPP_OK;
    const MatchPlan p = plan_match(m, n);
    if (workspace_bytes < p.bytes) {
        set_error("match_top1: workspace %zu < required %zu bytes", workspace_bytes, p.bytes);
        return SPP_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    float *qn = reinterpret_cast<float *>(ws + p.off_qn);
    __nv_bfloat16 *qb = reinterpret_cast<__nv_bfloat16 *>(ws + p.off_qb);
    Cand *part = reinterpret_cast<Cand *>(ws + p.off_part);
    const __nv_bfloat16 *gal = reinterpret_cast<const __nv_bfloat16 *>(gallery);

    l2_normalize_kernel<<<(m + 7) / 8, 256, 0, st>>>(emb, m, kDim, 1, 1e-12f, qn, nullptr, qb);
    SPP_CHECK_LAUNCH();

    if (simt) {
        dim3 grid(m, p.nsplit);
        match_simt_top2_kernel<<<grid, 128, 0, st>>>(qn, gal, m, n, p.nsplit, part);
        SPP_CHECK_LAUNCH();
    } else {
        CUtensorMap ta, tb;
        int rc = make_tmap(&ta, qb, m, BM);
        if (rc) return rc;
        rc = make_tmap(&tb, gal, n, BN);
        if (rc) return rc;
        static bool configured = false;
        if (!configured) {
            SPP_CHECK_CUDA(cudaFuncSetAttribute(match_gemm_top2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
            configured = true;
        }
        const int items = p.m_tiles * p.nsplit;
        GemmParams gp{m, n, p.tiles_n, p.nsplit, items, part};
        match_gemm_top2_kernel<<<items < p.sms ? items : p.sms, kGemmThreads, kGemmSmem, st>>>(ta, tb, gp);
        SPP_CHECK_LAUNCH();
    }
    match_finalize_kernel<<<(m + 7) / 8, 256, 0, st>>>(qn, gal, part, m, n, p.nsplit * 2, threshold, id_offset, out_id, out_sim,
                                                      out_key);
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}

extern "C" int spp_match_top1(const float *emb, const uint16_t *gallery, int m, int n, int dim, float threshold, int id_offset,
                              int *out_id, float *out_sim, unsigned long long *out_key, void *workspace,
                              size_t workspace_bytes, spp_stream_t stream) {
    return match_common(emb, gallery, m, n, dim, threshold, id_offset, out_id, out_sim, out_key, workspace, workspace_bytes,
                        stream, false);
}

// Test hook (declared in spp_internal.h, not part of the drop-in surface): CUDA-core candidate search.
extern "C" int spp_debug_match_top1_simt(const float *emb, const uint16_t *gallery, int m, int n, int dim, float threshold,
                                         int id_offset, int *out_id, float *out_sim, unsigned long long *out_key,
                                         void *workspace, size_t workspace_bytes, spp_stream_t stream) {
    return match_common(emb, gallery, m, n, dim, threshold, id_offset, out_id, out_sim, out_key, workspace, workspace_bytes,
                        stream, true);
}

extern "C" int spp_match_unpack_keys(const unsigned long long *keys, int m, float threshold, int *out_id, float *out_sim,
                                     spp_stream_t stream) {
    SPP_CHECK_ARG(keys && m >= 0, "match_unpack_keys: bad arguments");
    if (m == 0) return SPP_OK;
    match_unpack_kernel<<<(m + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(keys, m, threshold, out_id, out_sim);
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}
