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
#include "peer_exchange.cuh"

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
// tcgen05.ld 32 lanes x 32 columns: thread = accumulator row, registers = 32 consecutive columns.  Issue and wait are
// separate so that the load of the next column block overlaps the arithmetic on the current one; the wait names the
// registers as in/out operands, which keeps every use of them behind it.
#define SPP_R32(r) \
    r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12], r[13], r[14], r[15], r[16], r[17], r[18], \
        r[19], r[20], r[21], r[22], r[23], r[24], r[25], r[26], r[27], r[28], r[29], r[30], r[31]
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// ---- candidate keys ---------------------------------------------------------------------------------------------
// The epilogue ranks scores through 32-bit keys: the fp32 bits of (score + off) — positive, hence ordered as integers —
// with the low 8 mantissa bits replaced by (255 - column inside the 256-wide gallery tile).  Inserting a key into a
// sorted (best, second, dropped-max) triple is four integer min / max operations, branch-free, and the winner's column
// travels inside the key.  Cost: scores are resolved to 2^-14 (off + score < 4), which the re-score band absorbs.
__device__ __forceinline__ void key_insert(uint32_t k, uint32_t &k1, uint32_t &k2, uint32_t &k3) {
    k3 = max(k3, min(k, k2));
    const uint32_t t = max(k, k2);
    k2 = min(t, k1);
    k1 = max(k, k1);
}
__device__ __forceinline__ float key_score_lo(uint32_t k, float off) { return k ? __uint_as_float(k & 0xffffff00u) - off : -INFINITY; }
__device__ __forceinline__ float key_score_hi(uint32_t k, float off) { return k ? __uint_as_float(k | 0xffu) - off : -INFINITY; }

// ------------------------------------------------------------------------------------------------
// GEMM + fused row top-2
// ------------------------------------------------------------------------------------------------
struct GemmParams {
    int m, n;
    int tiles_n, nsplit;   // every M-tile is cut into nsplit chunks of gallery tiles; work item = (M-tile, chunk)
    int items;             // m_tiles * nsplit, distributed round-robin over the persistent CTAs
    int m_tiles;           // items are numbered probe tile fastest (mt = item % m_tiles, chunk = item / m_tiles): the CTAs of one round
                           // work on the same few gallery chunks, whose tiles then cross HBM once and come from L2 for the other
                           // probe tiles (chunk fastest, 640 x 1 M: each chunk was streamed from HBM in both rounds)
    int chunk_fastest;     // SPP_MATCH_ORDER=chunk: the old numbering (profiling knob)
    const unsigned *step;  // peer exchange: device step counter, parity selects the probe buffer (tmap_a0 / tmap_a1); NULL: tmap_a0
    Cand *part;            // [m_tiles*BM, nsplit, 2]  best two (score, id) of every (probe row, chunk)
    float *dropped;        // [tiles_n, m_tiles*BM]    per (gallery tile, probe row): upper bound of the scores of that tile's
                           //                          rows that are NOT one of the chunk's best two
    float key_off;         // score -> key offset (> the largest |score|)
};

__global__ void __launch_bounds__(kGemmThreads, 1)
match_gemm_top2_kernel(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1,
                       const __grid_constant__ CUtensorMap tmap_b, const GemmParams prm) {
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
    const CUtensorMap *tmap_a = (prm.step && (__ldcg(prm.step) & 1u)) ? &tmap_a1 : &tmap_a0;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(tmap_a);
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
                const int mt = prm.chunk_fastest ? item / prm.nsplit : item % prm.m_tiles, sp = prm.chunk_fastest ? item % prm.nsplit : item / prm.m_tiles;
                int nt0, ntiles;
                chunk_range(sp, nt0, ntiles);
                mbar_wait(a_empty, (it & 1) ^ 1);            // previous item's MMAs are done with the probe tile
                mbar_arrive_expect_tx(a_full, kKBlocks * kABytes);
                for (int kb = 0; kb < kKBlocks; ++kb) tma_load_2d(smem_a + (size_t)kb * kABytes, tmap_a, a_full, kb * BK, mt * BM);
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
                const int sp = prm.chunk_fastest ? item % prm.nsplit : item / prm.m_tiles;
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
        // ===== epilogue: TMEM -> registers -> per probe row: best two of the chunk + per-tile dropped maximum =====
        const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
        const int row_in_tile = quarter * 32 + lane;
        const float off = prm.key_off;
        const int m_pad = ((prm.m + BM - 1) / BM) * BM;
        uint32_t tcount = 0;
        for (int item = blockIdx.x; item < prm.items; item += gridDim.x) {
            const int mt = prm.chunk_fastest ? item / prm.nsplit : item % prm.m_tiles, sp = prm.chunk_fastest ? item % prm.nsplit : item / prm.m_tiles;
            int nt0, ntiles;
            chunk_range(sp, nt0, ntiles);
            uint32_t K1 = 0, K2 = 0;                     // running best two keys of the chunk and the tiles they came from
            int T1 = 0, T2 = 0;
            float *drow = prm.dropped + (size_t)(mt * BM + row_in_tile);
            for (int t = 0; t < ntiles; ++t, ++tcount) {
                const int acc = tcount & 1;
                const uint32_t acc_phase = (tcount >> 1) & 1;
                mbar_wait(&t_full[acc], acc_phase);
                tc_fence_after();
                const int tile = nt0 + t, n0 = tile * BN;
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
                const bool partial = n0 + BN > prm.n;    // last tile: rows past the gallery's end are zero-filled, not candidates
                uint32_t L1 = 0, L2 = 0, L3 = 0;         // tile-local best, second, dropped maximum (keys)
                // One 32-column block.  Skipped whole when even its maximum cannot enter the top two (the usual case after
                // the first tiles); otherwise 32 branch-free insertions.  Dropped scores only ever raise L3.
                auto block = [&](const uint32_t (&r)[32], int c) {
                    float mx = __uint_as_float(r[0]);
#pragma unroll
                    for (int j = 1; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
                    const uint32_t kmx = __float_as_uint(mx + off) | 0xffu;          // >= every key of the block
                    if (partial || kmx > max(K2, L2)) {
                        const uint32_t cb = 255u - (uint32_t)(c * 32);               // low five bits all ones: cb - j == cb ^ j
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            uint32_t k = (__float_as_uint(__uint_as_float(r[j]) + off) & 0xffffff00u) | (cb ^ (uint32_t)j);
                            if (partial && n0 + c * 32 + j >= prm.n) k = 0;
                            key_insert(k, L1, L2, L3);
                        }
                    } else {
                        L3 = max(L3, kmx);
                    }
                };
                uint32_t ra[32], rb[32];
                tmem_ld32_issue(taddr, ra);
                tmem_ld32_wait(ra);
#pragma unroll 1
                for (int c = 0; c < BN / 32; c += 2) {
                    tmem_ld32_issue(taddr + (c + 1) * 32, rb);
                    block(ra, c);
                    tmem_ld32_wait(rb);
                    if (c + 2 < BN / 32) tmem_ld32_issue(taddr + (c + 2) * 32, ra);
                    block(rb, c + 1);
                    if (c + 2 < BN / 32) tmem_ld32_wait(ra);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_empty[acc]);
                // merge the tile's best two into the chunk's; whatever falls out is recorded against ITS tile
                uint32_t D = L3;
                auto drop = [&](uint32_t k, int tl) {
                    if (!k) return;
                    if (tl == tile) { D = max(D, k); return; }
                    float *p = drow + (size_t)tl * m_pad;                       // an earlier tile of this chunk: this thread wrote it
                    *p = fmaxf(*p, key_score_hi(k, off));
                };
                auto offer = [&](uint32_t k) {
                    if (k > K1) { drop(K2, T2); K2 = K1; T2 = T1; K1 = k; T1 = tile; }
                    else if (k > K2) { drop(K2, T2); K2 = k; T2 = tile; }
                    else D = max(D, k);
                };
                offer(L1);
                offer(L2);
                drow[(size_t)tile * m_pad] = key_score_hi(D, off);
            }
            Cand *o = prm.part + ((size_t)(mt * BM + row_in_tile) * prm.nsplit + sp) * 2;
            o[0] = Cand{key_score_lo(K1, off), K1 ? T1 * BN + 255 - (int)(K1 & 0xffu) : 0x7fffffff};
            o[1] = Cand{key_score_lo(K2, off), K2 ? T2 * BN + 255 - (int)(K2 & 0xffu) : 0x7fffffff};
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// chunk sp of an M-tile covers gallery tiles [nt0, nt0 + ntiles) (same split as the GEMM's work items)
__host__ __device__ __forceinline__ void chunk_tiles(int tiles_n, int nsplit, int sp, int &nt0, int &ntiles) {
    const int per = tiles_n / nsplit, rem = tiles_n - per * nsplit;
    nt0 = sp * per + (sp < rem ? sp : rem);
    ntiles = per + (sp < rem ? 1 : 0);
}

// CUDA-core fp32 version of the same candidate search — device-side cross-check for the tests
// (spp_match_top1 never dispatches to it).  One CTA per (probe row, chunk); pass 1 finds the chunk's best two, pass 2
// the per-tile maximum over all other rows (the `dropped` table).
__global__ void __launch_bounds__(128) match_simt_top2_kernel(const float *__restrict__ qn, const __nv_bfloat16 *__restrict__ gal,
                                                              int m, int n, int tiles_n, int nsplit, int m_pad, Cand *part, float *dropped) {
    __shared__ float q[kDim];
    const int row = blockIdx.x, sp = blockIdx.y;
    for (int i = threadIdx.x; i < kDim; i += blockDim.x) q[i] = qn[(size_t)row * kDim + i];
    __syncthreads();
    int nt0, ntiles;
    chunk_tiles(tiles_n, nsplit, sp, nt0, ntiles);
    const int lo = nt0 * BN, hi = min(n, (nt0 + ntiles) * BN);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto score = [&](int g) {
        const __nv_bfloat16 *gr = gal + (size_t)g * kDim;
        float s = 0.f;
        for (int i = lane; i < kDim; i += 32) s = fmaf(q[i], __bfloat162float(gr[i]), s);
        return warp_sum(s);
    };
    float v1 = -INFINITY, v2 = -INFINITY;
    int i1 = 0x7fffffff, i2 = 0x7fffffff;
    for (int g = lo + warp; g < hi; g += 4) {
        const float s = score(g);
        if (s > v1 || (s == v1 && g < i1)) { v2 = v1; i2 = i1; v1 = s; i1 = g; }
        else if (s > v2 || (s == v2 && g < i2)) { v2 = s; i2 = g; }
    }
    __shared__ Cand sc[4][2];
    __shared__ Cand top[2];
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
        top[0] = b1;
        top[1] = b2;
    }
    __syncthreads();
    const int k1 = top[0].i, k2 = top[1].i;
    for (int t = warp; t < ntiles; t += 4) {                     // a warp per tile
        const int t_lo = (nt0 + t) * BN, t_hi = min(n, t_lo + BN);
        float d = -INFINITY;
        for (int g = t_lo; g < t_hi; ++g)
            if (g != k1 && g != k2) d = fmaxf(d, score(g));
        if (lane == 0) dropped[(size_t)(nt0 + t) * m_pad + row] = d;
    }
}

// ------------------------------------------------------------------------------------------------
// exact fp32 re-score, gate, key packing
// ------------------------------------------------------------------------------------------------
// A gallery row whose candidate-search score is more than `prune` below the row's best candidate-search score cannot be
// the fp32 arg-max.  Candidate search = bf16 probe x bf16 gallery with fp32 accumulation; re-score = fp32 probe x the
// re-score gallery (the same bf16 rows, or the caller's fp32 rows).  With R = the largest gallery row norm and unit-norm
// probes, |s_search - s_rescore| <= u * R (+ u * R when the re-score gallery is fp32, whose rows the bf16 copy rounds),
// u = 2^-8 the bf16 unit roundoff (Cauchy-Schwarz on sum |q_i g_i|), plus the key quantisation and < 1e-4 of fp32
// accumulation error.  prune = 2 * that bound; the host computes it (match_prune_margin).
//
// The GEMM epilogue keeps the best TWO rows per (probe, gallery chunk) and, per (probe, 256-row gallery tile), an upper
// bound of the scores of all OTHER rows of that tile (`dropped`).  A tile whose bound reaches vmax - prune may hide the
// fp32 arg-max (three or more near-duplicate enrolments of one person, the crowded top of a 1M-id gallery): its 256 rows
// are then re-scored in exact fp32 by all warps of the CTA.  Rare on realistic data — and the returned id is the fp32
// arg-max unconditionally.
struct FinParams {
    const float *qn;              // [m, 512] normalised probes (peer exchange: parity-0 buffer)
    size_t qn_parity_stride;      // peer exchange: elements between the two parity buffers
    const unsigned *step;         // peer exchange: device step counter (parity), else NULL
    const __nv_bfloat16 *gal;
    const float *gal_f32;         // optional fp32 rows for the re-score (NULL: the bf16 rows)
    const Cand *part;
    const float *dropped;         // [tiles_n, m_pad]
    int m, n, nsplit, tiles_n, m_pad;
    float prune, threshold;
    int id_offset;
    int *out_id;
    float *out_sim;
    unsigned long long *out_key;
    PeerPtrs peer;                // world > 0: keys go to the owners' exchange buffers instead of out_key
    size_t off_keys, keys_stride;
};

constexpr int kFinWarps = 8;      // probe rows per CTA (m_pad is a multiple of 128, so a CTA's 8 rows share a 32-byte sector of `dropped`)
constexpr int kFinScan = 4;       // gallery tiles per thread and scan round (independent loads in flight)
constexpr int kFinQueue = 32 * kFinWarps * kFinWarps * kFinScan;   // worst case of one scan round: every (tile, row) pair hits
constexpr int kPer = kDim / 32;   // 16 dimensions per lane

template <bool F32G>
struct RowFrag {
    uint4 v[F32G ? 4 : 2];
};
template <bool F32G>
__device__ __forceinline__ RowFrag<F32G> load_frag(const FinParams &p, int g, int lane) {
    RowFrag<F32G> f;
    if constexpr (F32G) {
        const uint4 *r = reinterpret_cast<const uint4 *>(p.gal_f32 + (size_t)g * kDim + lane * kPer);
#pragma unroll
        for (int i = 0; i < 4; ++i) f.v[i] = __ldg(r + i);
    } else {
        const uint4 *r = reinterpret_cast<const uint4 *>(p.gal + (size_t)g * kDim + lane * kPer);
        f.v[0] = __ldg(r);
        f.v[1] = __ldg(r + 1);
    }
    return f;
}
template <bool F32G>
__device__ __forceinline__ float dot_frag(const float (&q)[kPer], const RowFrag<F32G> &f) {
    float s = 0.f;
    if constexpr (F32G) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            s = fmaf(q[4 * i], __uint_as_float(f.v[i].x), s);
            s = fmaf(q[4 * i + 1], __uint_as_float(f.v[i].y), s);
            s = fmaf(q[4 * i + 2], __uint_as_float(f.v[i].z), s);
            s = fmaf(q[4 * i + 3], __uint_as_float(f.v[i].w), s);
        }
    } else {
        const unsigned w[8] = {f.v[0].x, f.v[0].y, f.v[0].z, f.v[0].w, f.v[1].x, f.v[1].y, f.v[1].z, f.v[1].w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {                // bf16 -> fp32 is a 16-bit shift
            s = fmaf(q[2 * i], __uint_as_float(w[i] << 16), s);
            s = fmaf(q[2 * i + 1], __uint_as_float(w[i] & 0xffff0000u), s);
        }
    }
    return s;
}
__device__ __forceinline__ void load_probe(const float *qn, int row, int lane, float (&q)[kPer]) {
    // lane owns 16 contiguous dimensions: the probe is 4 x 128-bit loads, a bf16 gallery row 2 x 128-bit loads per lane
    const float4 *qp = reinterpret_cast<const float4 *>(qn + (size_t)row * kDim + lane * kPer);
#pragma unroll
    for (int i = 0; i < kPer / 4; ++i) {
        const float4 v = __ldg(qp + i);
        q[4 * i] = v.x; q[4 * i + 1] = v.y; q[4 * i + 2] = v.z; q[4 * i + 3] = v.w;
    }
}

template <bool F32G>
__global__ void __launch_bounds__(32 * kFinWarps) match_finalize_kernel(const FinParams prm) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * kFinWarps, row = row0 + warp;
    const bool active = row < prm.m;
    const int n = prm.n, nsplit = prm.nsplit;
    unsigned step = 0;
    const float *qn = prm.qn;
    if (prm.step) {
        step = __ldcg(prm.step);
        qn += (size_t)(step & 1u) * prm.qn_parity_stride;
    }
    __shared__ float s_thr[kFinWarps];
    __shared__ int s_q[kFinQueue];
    __shared__ int s_nq;
    __shared__ float s_pv[kFinWarps];
    __shared__ int s_pi[kFinWarps];

    // first round of the `dropped` scan (below): thread = gallery tile, one 32-byte sector holds the bounds of this CTA's
    // 8 probe rows.  Requested here so that the loads are in flight during the candidate re-score.
    constexpr int kRound = 32 * kFinWarps * kFinScan;
    float4 da[kFinScan], db[kFinScan];
    auto load_bounds = [&](int t0) {
#pragma unroll
        for (int u = 0; u < kFinScan; ++u) {
            const int tile = t0 + u * 32 * kFinWarps + (int)threadIdx.x;
            if (tile < prm.tiles_n) {
                const float4 *d = reinterpret_cast<const float4 *>(prm.dropped + (size_t)tile * prm.m_pad + row0);
                da[u] = __ldg(d);
                db[u] = __ldg(d + 1);
            }
        }
    };
    load_bounds(0);

    float q[kPer];
    float vmax = -INFINITY, best = -INFINITY;
    int bidx = 0x7fffffff;
    if (active) {
        load_probe(qn, row, lane, q);
        const int ncand = nsplit * 2;
        const Cand *c = prm.part + (size_t)row * ncand;
        for (int k = lane; k < ncand; k += 32) {
            const Cand x = c[k];
            if (x.i >= 0 && x.i < n) vmax = fmaxf(vmax, x.v);
        }
        vmax = warp_max(vmax);
        for (int k0 = 0; k0 < ncand; k0 += 32) {
            const int k = k0 + lane;
            Cand x{-INFINITY, -1};
            if (k < ncand) x = c[k];
            const bool need = x.i >= 0 && x.i < n && x.v >= vmax - prm.prune;
            unsigned todo = __ballot_sync(FULL, need);
            while (todo) {
                // up to four candidates per round: their rows are all requested before the first dot product
                int g[4];
                RowFrag<F32G> fr[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    g[u] = -1;
                    if (todo) {
                        const int src = __ffs(todo) - 1;
                        todo &= todo - 1;
                        g[u] = __shfl_sync(FULL, x.i, src);
                        fr[u] = load_frag<F32G>(prm, g[u], lane);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (g[u] < 0) continue;              // warp-uniform
                    const float s = warp_sum(dot_frag<F32G>(q, fr[u]));
                    if (s > best || (s == best && g[u] < bidx)) { best = s; bidx = g[u]; }
                }
            }
        }
    }

    // ---- tiles whose dropped rows reach the re-score band: exact fp32 re-scan of the tile by the whole CTA ----
    if (lane == 0) s_thr[warp] = active ? vmax - prm.prune : INFINITY;
    if (threadIdx.x == 0) s_nq = 0;
    __syncthreads();
    constexpr int kInFlight = F32G ? 4 : 8;
    for (int t0 = 0; t0 < prm.tiles_n; t0 += kRound) {
        if (t0) load_bounds(t0);
#pragma unroll
        for (int u = 0; u < kFinScan; ++u) {
            const int tile = t0 + u * 32 * kFinWarps + (int)threadIdx.x;
            if (tile < prm.tiles_n) {
                const float bound[kFinWarps] = {da[u].x, da[u].y, da[u].z, da[u].w, db[u].x, db[u].y, db[u].z, db[u].w};
#pragma unroll
                for (int r = 0; r < kFinWarps; ++r)
                    if (row0 + r < prm.m && bound[r] >= s_thr[r]) s_q[atomicAdd(&s_nq, 1)] = (r << 24) | tile;
            }
        }
        __syncthreads();
        const int nq = s_nq;
        for (int qi = 0; qi < nq; ++qi) {
            const int e = s_q[qi];
            const int owner = e >> 24, tl = e & 0xffffff;
            float oq[kPer];
            load_probe(qn, row0 + owner, lane, oq);
            const int lo = tl * BN + warp * (BN / kFinWarps), hi = min(n, lo + BN / kFinWarps);
            float pb = -INFINITY;
            int pi = 0x7fffffff;
            for (int g0 = lo; g0 < hi; g0 += kInFlight) {
                RowFrag<F32G> fr[kInFlight];
#pragma unroll
                for (int u = 0; u < kInFlight; ++u)
                    if (g0 + u < hi) fr[u] = load_frag<F32G>(prm, g0 + u, lane);
#pragma unroll
                for (int u = 0; u < kInFlight; ++u) {
                    if (g0 + u >= hi) continue;
                    const float sc = warp_sum(dot_frag<F32G>(oq, fr[u]));
                    if (sc > pb) { pb = sc; pi = g0 + u; }          // ascending ids: the first maximum stays
                }
            }
            if (lane == 0) { s_pv[warp] = pb; s_pi[warp] = pi; }
            __syncthreads();
            if (warp == owner) {
#pragma unroll
                for (int w = 0; w < kFinWarps; ++w) {
                    const float v = s_pv[w];
                    const int i = s_pi[w];
                    if (v > best || (v == best && i < bidx)) { best = v; bidx = i; }
                }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) s_nq = 0;
        __syncthreads();
    }

    if (active && lane == 0) {
        const bool found = bidx != 0x7fffffff;
        const int gid = found ? bidx + prm.id_offset : -1;
        const bool pass = found && !(best < prm.threshold);      // NaN threshold: no gate
        if (prm.out_id) prm.out_id[row] = pass ? gid : -1;
        if (prm.out_sim) prm.out_sim[row] = found ? best : -INFINITY;
        const unsigned long long hi = (unsigned long long)(unsigned)float_to_ordered(found ? best : -INFINITY);
        const unsigned long long key = (hi << 32) | (unsigned long long)(0xffffffffu - (unsigned)(found ? gid : 0x7fffffff));
        if (prm.peer.world > 0) {
            // the key of probe `row` = (owner rank, local row) goes straight into the owner's exchange buffer (P2P store)
            const int owner = row / prm.peer.m_local, j = row - owner * prm.peer.m_local;
            unsigned long long *dst = reinterpret_cast<unsigned long long *>(prm.peer.buf[owner] + prm.off_keys + (step & 1u) * prm.keys_stride) +
                                      (size_t)prm.peer.rank * prm.peer.m_local + j;
            *dst = key;
        } else if (prm.out_key) {
            prm.out_key[row] = key;
        }
    }
    if (prm.peer.world > 0) {
        // last CTA: every key of this rank has been written -> publish the step number to all owners
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            PeerHeader *h = reinterpret_cast<PeerHeader *>(prm.peer.buf[prm.peer.rank]);
            if (atomicAdd(&h->key_count, 1u) == gridDim.x - 1) {
                h->key_count = 0;
                __threadfence_system();
                for (int r = 0; r < prm.peer.world; ++r) st_release_sys(peer_key_flag(prm.peer.buf[r], step & 1u, prm.peer.rank), step + 1);
            }
        }
    }
}

__device__ __forceinline__ void unpack_key(unsigned long long k, float threshold, int &id, float &sim) {
    sim = ordered_to_float((int32_t)(unsigned)(k >> 32));
    const unsigned gid = 0xffffffffu - (unsigned)(k & 0xffffffffu);
    const bool found = gid != 0x7fffffffu;
    id = (found && !(sim < threshold)) ? (int)gid : -1;
}

__global__ void match_unpack_kernel(const unsigned long long *keys, int m, float threshold, int *out_id, float *out_sim) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int id;
    float sim;
    unpack_key(keys[i], threshold, id, sim);
    if (out_sim) out_sim[i] = sim;
    if (out_id) out_id[i] = id;
}

// ------------------------------------------------------------------------------------------------
// peer exchange kernels (gallery sharded over the GPUs of one box; peer_exchange.cuh)
// ------------------------------------------------------------------------------------------------
// F.normalize of this rank's probes, written as fp32 + bf16 into EVERY rank's exchange buffer (the all-gather, as P2P
// stores over NVLink), then the step number is published to every peer by the last CTA.
__global__ void __launch_bounds__(256) peer_normalize_push_kernel(const float *__restrict__ emb, float eps, const PeerPtrs peer,
                                                                  const PeerLayout lay) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    PeerHeader *h = reinterpret_cast<PeerHeader *>(peer.buf[peer.rank]);
    const unsigned step = __ldcg(&h->step), parity = step & 1u;
    if (row < peer.m_local) {
        const float *xr = emb + (size_t)row * kDim;
        float v[kDim / 32];
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < kDim / 32; ++k) {           // same element order as l2_normalize_kernel: bit-identical probes
            v[k] = xr[lane + 32 * k];
            ss = fmaf(v[k], v[k], ss);
        }
        ss = warp_sum(ss);
        const float den = fmaxf(sqrtf(ss), eps);
#pragma unroll
        for (int k = 0; k < kDim / 32; ++k) v[k] = __fdiv_rn(v[k], den);
        const size_t grow = (size_t)peer.rank * peer.m_local + row;
        for (int r = 0; r < peer.world; ++r) {
            float *df = reinterpret_cast<float *>(peer.buf[r] + lay.off_f32 + parity * lay.f32_stride) + grow * kDim;
            __nv_bfloat16 *db = reinterpret_cast<__nv_bfloat16 *>(peer.buf[r] + lay.off_bf16 + parity * lay.bf16_stride) + grow * kDim;
#pragma unroll
            for (int k = 0; k < kDim / 32; ++k) {
                df[lane + 32 * k] = v[k];
                db[lane + 32 * k] = __float2bfloat16_rn(v[k]);
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&h->push_count, 1u) == gridDim.x - 1) {
            h->push_count = 0;
            __threadfence_system();
            for (int r = 0; r < peer.world; ++r) st_release_sys(peer_probe_flag(peer.buf[r], parity, peer.rank), step + 1);
        }
    }
}

// Wait until the probes of every rank have landed in this rank's buffer.  A kernel of its own: the kernel boundary makes
// the peers' (generic-proxy) stores visible to the TMA loads of the GEMM that follows.
__global__ void peer_wait_probes_kernel(const PeerPtrs peer) {
    unsigned char *own = peer.buf[peer.rank];
    const unsigned step = __ldcg(&reinterpret_cast<PeerHeader *>(own)->step);
    if ((int)threadIdx.x < peer.world) peer_wait_flag(peer_probe_flag(own, step & 1u, threadIdx.x), step + 1);
}

// The top-1 (value, index) reduction: wait for every rank's keys of MY probes, take the integer maximum (higher
// similarity first, lower id on ties), unpack + gate.  The last CTA advances the step counter.
__global__ void __launch_bounds__(256) peer_reduce_unpack_kernel(const PeerPtrs peer, const PeerLayout lay, float threshold,
                                                                 int *out_id, float *out_sim, unsigned long long *out_key) {
    unsigned char *own = peer.buf[peer.rank];
    PeerHeader *h = reinterpret_cast<PeerHeader *>(own);
    const unsigned step = __ldcg(&h->step), parity = step & 1u;
    if ((int)threadIdx.x < peer.world) peer_wait_flag(peer_key_flag(own, parity, threadIdx.x), step + 1);
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < peer.m_local) {
        const unsigned long long *keys = reinterpret_cast<const unsigned long long *>(own + lay.off_keys + parity * lay.keys_stride);
        long long best = (long long)__ldcg(keys + j);
        for (int r = 1; r < peer.world; ++r) {
            const long long k = (long long)__ldcg(keys + (size_t)r * peer.m_local + j);
            best = k > best ? k : best;
        }
        int id;
        float sim;
        unpack_key((unsigned long long)best, threshold, id, sim);
        if (out_id) out_id[j] = id;
        if (out_sim) out_sim[j] = sim;
        if (out_key) out_key[j] = (unsigned long long)best;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&h->done_count, 1u) == gridDim.x - 1) {
            h->done_count = 0;
            __threadfence();
            h->step = step + 1;
        }
    }
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
    int m_tiles, tiles_n, nsplit, sms, m_pad;
    size_t off_qn, off_qb, off_part, off_drop, bytes;
};

// own_probes: the workspace also holds the normalised probes (fp32 + bf16); false when they live in a peer exchange buffer
MatchPlan plan_match(int m, int n, bool own_probes = true) {
    MatchPlan p{};
    p.m_tiles = (m + BM - 1) / BM;
    p.tiles_n = (n + BN - 1) / BN;
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    {   // spp_set_launch_limit(SPP_LIMIT_MATCH_CTAS): run the GEMM on a few SMs beside a kernel that holds the others
        const int lim = launch_limit(1);
        if (lim > 0 && sms > lim) sms = lim;
    }
    // Work items = (M-tile, chunk of gallery tiles).  Small problems (fewer items than SMs even at one gallery
    // tile per item): one item per CTA and as many CTAs as there are SMs — every CTA pays a 128 KB probe-tile
    // load, but the kernel is over in ~1/6 of the time 8-tile chunks take (cfg2: 145 CTAs x 1-2 tiles instead of
    // 25 CTAs x 8; measured 72 us -> see profiles/README.md), and a short kernel is what the graph's other
    // branches need from it.  SPP_MATCH_CHUNK_TILES=<t> restores chunks of >= t tiles (profiling knob).
    // Large problems: the grid is one persistent CTA per SM and the chunk count is chosen so that
    // items ~ r * SMs (balanced rounds, <= 8 of them).
    static const int chunk_tiles_env = [] { const char *e = getenv("SPP_MATCH_CHUNK_TILES"); const int v = e ? atoi(e) : 1; return v < 1 ? 1 : v; }();
    const int want = (p.tiles_n + chunk_tiles_env - 1) / chunk_tiles_env > 0 ? (p.tiles_n + chunk_tiles_env - 1) / chunk_tiles_env : 1;
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
    p.off_qn = off;    off += own_probes ? align_up((size_t)m * kDim * 4, 1024) : 0;
    p.off_qb = off;    off += own_probes ? align_up((size_t)m * kDim * 2, 1024) : 0;
    p.m_pad = p.m_tiles * BM;
    p.off_part = off;  off += align_up((size_t)p.m_pad * ns * 2 * sizeof(Cand), 1024);
    p.off_drop = off;  off += align_up((size_t)p.m_pad * p.tiles_n * sizeof(float), 1024);
    p.bytes = off;
    return p;
}

// score -> key offset: keys are the fp32 bits of (score + off), which must stay positive; |score| <= R * 1.01
float match_key_offset(float max_row_norm) { return (max_row_norm > 0.f ? max_row_norm : 1.0f) * 1.01f + 1.0f; }

float match_prune_margin(float max_row_norm, bool f32_rescore) {
    const float r = max_row_norm > 0.f ? max_row_norm : 1.0f;
    const float u = 0.00390625f;                         // bf16 unit roundoff 2^-8
    // key quantisation: the low 8 mantissa bits of (score + off) < 2 * off are dropped
    int e = 0;
    std::frexp(2.0f * match_key_offset(r), &e);          // 2 * off = f * 2^e, f in [0.5, 1)
    const float quant = std::ldexp(1.0f, e - 1 - 15);
    return 2.0f * ((f32_rescore ? 2.0f : 1.0f) * u * r * 1.02f + quant + 1e-4f);
}

// candidate search (tcgen05 GEMM or the SIMT cross-check) over `m` normalised probes
int launch_search(const __nv_bfloat16 *qb0, const __nv_bfloat16 *qb1, const unsigned *step, const float *qn_simt,
                  const __nv_bfloat16 *gal, int m, int n, float max_row_norm, const MatchPlan &p, Cand *part, float *dropped,
                  cudaStream_t st, bool simt) {
    if (simt) {
        dim3 grid(m, p.nsplit);
        match_simt_top2_kernel<<<grid, 128, 0, st>>>(qn_simt, gal, m, n, p.tiles_n, p.nsplit, p.m_pad, part, dropped);
        SPP_CHECK_LAUNCH();
        return SPP_OK;
    }
    CUtensorMap ta0, ta1, tb;
    int rc = make_tmap(&ta0, qb0, m, BM);
    if (rc) return rc;
    rc = make_tmap(&ta1, qb1 ? qb1 : qb0, m, BM);
    if (rc) return rc;
    rc = make_tmap(&tb, gal, n, BN);
    if (rc) return rc;
    // per device and per context: set on every launch (about a microsecond; legal during stream capture)
    SPP_CHECK_CUDA(cudaFuncSetAttribute(match_gemm_top2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    const int items = p.m_tiles * p.nsplit;
    static const int chunk_fastest = [] { const char *e = getenv("SPP_MATCH_ORDER"); return (e && e[0] == 'c') ? 1 : 0; }();
    GemmParams gp{m, n, p.tiles_n, p.nsplit, items, p.m_tiles, chunk_fastest, step, part, dropped, match_key_offset(max_row_norm)};
    match_gemm_top2_kernel<<<items < p.sms ? items : p.sms, kGemmThreads, kGemmSmem, st>>>(ta0, ta1, tb, gp);
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}

int launch_finalize(const FinParams &fp, cudaStream_t st) {
    const int grid = (fp.m + kFinWarps - 1) / kFinWarps;
    if (fp.gal_f32) match_finalize_kernel<true><<<grid, 32 * kFinWarps, 0, st>>>(fp);
    else match_finalize_kernel<false><<<grid, 32 * kFinWarps, 0, st>>>(fp);
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}

}  // namespace
}  // namespace spp

using namespace spp;

extern "C" int spp_l2_normalize(const float *x, int m, int dim, int mode, float eps, float *out, float *norm,
                                uint16_t *out_bf16, spp_stream_t stream) {
    if (m == 0) return SPP_OK;
    SPP_CHECK_ARG(x && m >= 0 && dim > 0, "l2_normalize: bad arguments");
    SPP_CHECK_ARG(mode == 0 || mode == 1, "l2_normalize: mode must be 0 (x/||x||) or 1 (F.normalize)");
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

static int match_common(const float *emb, const uint16_t *gallery, const float *gallery_f32, float max_row_norm, int m, int n, int dim,
                        float threshold, int id_offset, int *out_id, float *out_sim, unsigned long long *out_key, void *workspace,
                        size_t workspace_bytes, spp_stream_t stream, bool simt) {
    if (m == 0) return SPP_OK;
    SPP_CHECK_ARG(emb && gallery && workspace, "match_top1: null pointer");
    SPP_CHECK_ARG(dim == kDim, "match_top1: embedding dimension must be %d (got %d)", kDim, dim);
    SPP_CHECK_ARG(m >= 0 && n >= 1, "match_top1: bad m=%d n=%d", m, n);
    SPP_CHECK_ARG(max_row_norm > 0.0f && max_row_norm < 1e4f, "match_top1: max_row_norm must be the largest gallery row norm (got %g)",
                  (double)max_row_norm);
    SPP_CHECK_ARG((reinterpret_cast<uintptr_t>(gallery) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 1023) == 0 &&
                      (reinterpret_cast<uintptr_t>(gallery_f32) & 15) == 0,
                  "match_top1: gallery must be 16-byte and workspace 1024-byte aligned");
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
    float *dropped = reinterpret_cast<float *>(ws + p.off_drop);
    const __nv_bfloat16 *gal = reinterpret_cast<const __nv_bfloat16 *>(gallery);

    l2_normalize_kernel<<<(m + 7) / 8, 256, 0, st>>>(emb, m, kDim, 1, 1e-12f, qn, nullptr, qb);
    SPP_CHECK_LAUNCH();
    int rc = launch_search(qb, nullptr, nullptr, qn, gal, m, n, max_row_norm, p, part, dropped, st, simt);
    if (rc) return rc;
    FinParams fp{};
    fp.qn = qn; fp.gal = gal; fp.gal_f32 = gallery_f32; fp.part = part; fp.dropped = dropped;
    fp.m = m; fp.n = n; fp.nsplit = p.nsplit; fp.tiles_n = p.tiles_n; fp.m_pad = p.m_pad;
    fp.prune = match_prune_margin(max_row_norm, gallery_f32 != nullptr);
    fp.threshold = threshold; fp.id_offset = id_offset;
    fp.out_id = out_id; fp.out_sim = out_sim; fp.out_key = out_key;
    return launch_finalize(fp, st);
}

extern "C" int spp_match_top1(const float *emb, const uint16_t *gallery, int m, int n, int dim, float threshold, int id_offset,
                              int *out_id, float *out_sim, unsigned long long *out_key, void *workspace,
                              size_t workspace_bytes, spp_stream_t stream) {
    return match_common(emb, gallery, nullptr, 1.0f, m, n, dim, threshold, id_offset, out_id, out_sim, out_key, workspace,
                        workspace_bytes, stream, false);
}

extern "C" int spp_match_top1_ex(const float *emb, const uint16_t *gallery, const float *gallery_f32, float max_row_norm, int m, int n,
                                 int dim, float threshold, int id_offset, int *out_id, float *out_sim, unsigned long long *out_key,
                                 void *workspace, size_t workspace_bytes, spp_stream_t stream) {
    return match_common(emb, gallery, gallery_f32, max_row_norm, m, n, dim, threshold, id_offset, out_id, out_sim, out_key, workspace,
                        workspace_bytes, stream, false);
}

// Test hook (declared in spp_internal.h, not part of the drop-in surface): CUDA-core candidate search.
extern "C" int spp_debug_match_top1_simt(const float *emb, const uint16_t *gallery, int m, int n, int dim, float threshold,
                                         int id_offset, int *out_id, float *out_sim, unsigned long long *out_key,
                                         void *workspace, size_t workspace_bytes, spp_stream_t stream) {
    return match_common(emb, gallery, nullptr, 1.0f, m, n, dim, threshold, id_offset, out_id, out_sim, out_key, workspace,
                        workspace_bytes, stream, true);
}

extern "C" int spp_match_unpack_keys(const unsigned long long *keys, int m, float threshold, int *out_id, float *out_sim,
                                     spp_stream_t stream) {
    SPP_CHECK_ARG(keys && m >= 0, "match_unpack_keys: bad arguments");
    if (m == 0) return SPP_OK;
    match_unpack_kernel<<<(m + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(keys, m, threshold, out_id, out_sim);
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}

// ------------------------------------------------------------------------------------------------
// gallery sharded over the GPUs of one box: peer-memory exchange (include/spp.h, peer_exchange.cuh)
// ------------------------------------------------------------------------------------------------
extern "C" size_t spp_peer_buffer_bytes(int world, int m_local, int dim) {
    if (world < 1 || world > SPP_MAX_PEERS || m_local < 1 || dim != kDim) return 0;
    return peer_layout(world, m_local).bytes;
}

extern "C" size_t spp_sharded_match_workspace_bytes(int world, int m_local, int n_shard, int dim) {
    if (world < 1 || world > SPP_MAX_PEERS || m_local < 1 || n_shard < 1 || dim != kDim) return 0;
    return plan_match(world * m_local, n_shard, false).bytes;
}

extern "C" int spp_sharded_match_top1(const spp_peer_group *group, const float *emb, const uint16_t *shard, const float *shard_f32,
                                      float max_row_norm, int n_shard, int dim, int id_offset, float threshold, int stages, int *out_id,
                                      float *out_sim, unsigned long long *out_key, void *workspace, size_t workspace_bytes,
                                      spp_stream_t stream) {
    SPP_CHECK_ARG(group && emb && shard && workspace, "sharded_match_top1: null pointer");
    SPP_CHECK_ARG(group->world >= 1 && group->world <= SPP_MAX_PEERS && group->rank >= 0 && group->rank < group->world && group->m_local >= 1,
                  "sharded_match_top1: bad peer group (world %d rank %d m_local %d)", group->world, group->rank, group->m_local);
    SPP_CHECK_ARG(dim == kDim && n_shard >= 1, "sharded_match_top1: dim must be %d and the shard non-empty", kDim);
    SPP_CHECK_ARG(max_row_norm > 0.0f && max_row_norm < 1e4f, "sharded_match_top1: bad max_row_norm");
    SPP_CHECK_ARG((reinterpret_cast<uintptr_t>(shard) & 15) == 0 && (reinterpret_cast<uintptr_t>(shard_f32) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(workspace) & 1023) == 0,
                  "sharded_match_top1: shard must be 16-byte and workspace 1024-byte aligned");
    PeerPtrs peer{};
    peer.world = group->world; peer.rank = group->rank; peer.m_local = group->m_local;
    for (int r = 0; r < group->world; ++r) {
        SPP_CHECK_ARG(group->buffers[r] && (reinterpret_cast<uintptr_t>(group->buffers[r]) & 1023) == 0,
                      "sharded_match_top1: exchange buffer of rank %d is null or not 1024-byte aligned", r);
        peer.buf[r] = static_cast<unsigned char *>(group->buffers[r]);
    }
    const PeerLayout lay = peer_layout(peer.world, peer.m_local);
    const int m = peer.world * peer.m_local;
    const MatchPlan p = plan_match(m, n_shard, false);
    if (workspace_bytes < p.bytes) {
        set_error("sharded_match_top1: workspace %zu < required %zu bytes", workspace_bytes, p.bytes);
        return SPP_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    Cand *part = reinterpret_cast<Cand *>(ws + p.off_part);
    float *dropped = reinterpret_cast<float *>(ws + p.off_drop);
    unsigned char *own = peer.buf[peer.rank];
    const unsigned *step = &reinterpret_cast<const PeerHeader *>(own)->step;
    const __nv_bfloat16 *gal = reinterpret_cast<const __nv_bfloat16 *>(shard);
    if (stages & SPP_SHARDED_STAGE_PUSH) {
        peer_normalize_push_kernel<<<(peer.m_local + 7) / 8, 256, 0, st>>>(emb, 1e-12f, peer, lay);
        SPP_CHECK_LAUNCH();
    }
    if (stages & SPP_SHARDED_STAGE_WAIT) {
        peer_wait_probes_kernel<<<1, 32, 0, st>>>(peer);
        SPP_CHECK_LAUNCH();
    }
    if (stages & SPP_SHARDED_STAGE_SEARCH) {
        const __nv_bfloat16 *qb0 = reinterpret_cast<const __nv_bfloat16 *>(own + lay.off_bf16);
        const __nv_bfloat16 *qb1 = reinterpret_cast<const __nv_bfloat16 *>(own + lay.off_bf16 + lay.bf16_stride);
        int rc = launch_search(qb0, qb1, step, nullptr, gal, m, n_shard, max_row_norm, p, part, dropped, st, false);
        if (rc) return rc;
    }
    if (stages & SPP_SHARDED_STAGE_FINALIZE) {
        FinParams fp{};
        fp.qn = reinterpret_cast<const float *>(own + lay.off_f32);
        fp.qn_parity_stride = lay.f32_stride / 4;
        fp.step = step;
        fp.gal = gal; fp.gal_f32 = shard_f32; fp.part = part; fp.dropped = dropped;
        fp.m = m; fp.n = n_shard; fp.nsplit = p.nsplit; fp.tiles_n = p.tiles_n; fp.m_pad = p.m_pad;
        fp.prune = match_prune_margin(max_row_norm, shard_f32 != nullptr);
        fp.threshold = NAN; fp.id_offset = id_offset;           // the gate is applied after the reduction
        fp.peer = peer; fp.off_keys = lay.off_keys; fp.keys_stride = lay.keys_stride;
        int rc = launch_finalize(fp, st);
        if (rc) return rc;
    }
    if (stages & SPP_SHARDED_STAGE_REDUCE) {
        peer_reduce_unpack_kernel<<<(peer.m_local + 255) / 256, 256, 0, st>>>(peer, lay, threshold, out_id, out_sim, out_key);
        SPP_CHECK_LAUNCH();
    }
    return SPP_OK;
}
