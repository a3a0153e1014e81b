// Heatmap decode: flip-test averaging + per-joint arg-max + {DARK/UDP | soft-argmax | quarter-offset}
// refinement + back-projection, ONE pass over the heatmaps.
//
// Replaces:
//   flip-back + average   training/lightning/pose_estimation/module.py:473-484 (channel swap as in
//                         "module copy.py":465-472 / HF modeling_vitpose.py:80-117)
//   DARK + UDP            HF image_processing_vitpose.py:175-313, 450-463
//   soft-argmax           training/lightning/pose_estimation/module.py:237-296, 534-546
//   quarter offset        gluoncv get_final_preds as called at pose_estimation/module_v2.py:214-222
//
// Design (HBM-bound: K*H*W*4 bytes per crop, x2 with the flip test, 16 B out per joint):
//   * persistent grid, one CTA per SM; every warp owns a private ring of smem stages and is its own
//     producer: lane 0 issues 1-D bulk-TMA copies (cp.async.bulk, one 12 KB joint map per copy, two
//     with the flip test) that complete on a per-stage mbarrier; no block-wide barrier anywhere.
//     ~190 KB of loads are in flight per SM, far above the ~45 KB Little's-law requirement.
//   * the warp scans the map from shared memory with 128-bit conflict-free reads, forming the
//     flip-average on the fly (mirrored column, pair-swapped channel) and keeping a running
//     (max, first index) per lane, then a shuffle arg-max with lowest-index tie-break.
//   * DARK touches only the (2r+3)^2 window around the arg-max: the separable Gaussian is evaluated
//     for the 3x3 tap block exactly as scipy does it (axis 0 first, fp64 accumulation in scipy's
//     summation order, fp32 intermediate, 'reflect' borders), then clip/log in fp32 and the 2x2
//     Newton step in fp64.  All of it lives in registers + 64 floats of per-warp scratch.
#include "spp_common.cuh"

#include <cuda_bf16.h>
#include <cfloat>
#include <type_traits>
#include <cstdlib>
#include <climits>
#include <cmath>

namespace spp {

namespace {

constexpr int kMaxRadius = 8;
constexpr int kModeCopyOnly = 99;   // undocumented: stream the maps through the pipeline and do nothing
constexpr int kScratch = 64 + 368;  // floats per warp: 3*(2r+3) <= 57 filter rows, then the (2r+3)^2 <= 361 window

struct DecodeParams {
    const void *hm;    // [P, K, H, W] fp32 or bf16 (template parameter T of the kernel)
    const void *hmf;
    const int *perm;
    const float *boxes;
    float *kpts;
    float *scores;
    int *amax;
    int P, K, H, W;
    int mode, flags, radius, crop_h, crop_w;
    int warps, stages;
    unsigned w4_magic;
    double gw[kMaxRadius + 1];  // Gaussian weights, gw[0] = centre (scipy _gaussian_kernel1d, normalised)
};

// scipy 'reflect' (d c b a | a b c d | d c b a) for -n <= i < 2n (the host checks radius + 1 <= n)
__device__ __forceinline__ int reflect_idx(int i, int n) {
    if (i < 0) i = -i - 1;
    if (i >= n) i = 2 * n - 1 - i;
    return i;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// element load / 4-element load widened to fp32 (bf16 -> fp32 is exact)
__device__ __forceinline__ float ld1(const float *p) { return *p; }
__device__ __forceinline__ float ld1(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
__device__ __forceinline__ float4 ld4(const float *base, int q4) { return reinterpret_cast<const float4 *>(base)[q4]; }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16 *base, int q4) {
    const uint2 w = reinterpret_cast<const uint2 *>(base)[q4];
    return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u), __uint_as_float(w.y << 16),
                       __uint_as_float(w.y & 0xffff0000u));
}

// One (possibly flip-averaged) heatmap; A/B may point to shared or global memory.
template <bool FLIP, typename T>
struct MapView {
    const T *A;
    const T *B;  // raw map of the mirrored crop, channel already pair-swapped
    int W;
    __device__ __forceinline__ float at(int r, int c) const {
        float a = ld1(A + r * W + c);
        if (FLIP) a = (a + ld1(B + r * W + (W - 1 - c))) * 0.5f;
        return a;
    }
};

// scipy.ndimage.correlate1d, symmetric-kernel branch: centre tap first, then the pairs from the far
// end inwards, accumulated in fp64 without contraction.  R > 0: compile-time radius (all taps are
// loaded before the dependent fp64 chain starts); R == 0: run-time radius.
template <int R, typename F>
__device__ __forceinline__ double sym_filter(F sample, const double *gw, int radius) {
    if constexpr (R > 0) {
        float lo[R > 0 ? R : 1], hi[R > 0 ? R : 1];
        const float c = sample(0);
#pragma unroll
        for (int j = 0; j < R; ++j) {
            lo[j] = sample(-(j + 1));
            hi[j] = sample(j + 1);
        }
        double acc = __dmul_rn((double)c, gw[0]);
#pragma unroll
        for (int j = R - 1; j >= 0; --j)
            acc = __dadd_rn(acc, __dmul_rn(__dadd_rn((double)lo[j], (double)hi[j]), gw[j + 1]));
        return acc;
    } else {
        double acc = __dmul_rn((double)sample(0), gw[0]);
        for (int jj = radius; jj >= 1; --jj)
            acc = __dadd_rn(acc, __dmul_rn(__dadd_rn((double)sample(-jj), (double)sample(jj)), gw[jj]));
        return acc;
    }
}

__device__ __forceinline__ float clip_log(float v) {
    v = fminf(fmaxf(v, 0.001f), 50.0f);
    return logf(v);
}

// log(clip(blur(map)))[y, x] for ONE position, whole warp cooperating (rare path: score <= 0 quirk).
template <bool FLIP, typename T>
__device__ float warp_blurred_log_single(const MapView<FLIP, T> &mv, int H, int y, int x, const double *gw, int radius,
                                         float *scratch, int lane) {
    const int n = 2 * radius + 1;
    __syncwarp();
    if (lane < n) {
        const int xx = reflect_idx(x - radius + lane, mv.W);
        double acc = sym_filter<0>([&](int d) { return mv.at(reflect_idx(y + d, H), xx); }, gw, radius);
        scratch[lane] = (float)acc;
    }
    __syncwarp();
    double acc = sym_filter<0>([&](int d) { return scratch[radius + d]; }, gw, radius);
    return clip_log((float)acc);
}

// Quirk Q6 helper: log(clip(blur(.))) at flat position t of HF's flattened, edge-padded batch ([P*K, H+2, W+2]); the
// staged map (A, B in shared memory) is used when t falls into map q, global memory otherwise.  Deliberately not inlined:
// seven call sites, opt-in path.
template <bool FLIP, typename T>
__device__ __noinline__ float hf_tap_value(const DecodeParams prm, const T *A, const T *B, long long q, long long t,
                                           float *scratch, int lane) {
    const int H = prm.H, W = prm.W, K = prm.K;
    const long long total = (long long)prm.P * K, stride = (long long)(W + 2) * (H + 2);
    const size_t map_elems = (size_t)H * W;
    if (t < 0) t += total * stride;                        // numpy negative index
    if (t >= total * stride) t = total * stride - 1;       // (numpy would raise; cannot happen for in-range arg-maxes)
    const long long qt = t / stride;
    const int r = (int)(t - qt * stride);
    const int py = r / (W + 2), px = r - py * (W + 2);
    const int ty = clampi(py - 1, 0, H - 1), tx = clampi(px - 1, 0, W - 1);
    MapView<FLIP, T> mv{A, B, W};
    if (qt != q) {
        const long long pt = qt / K;
        const int kt = (int)(qt - pt * K);
        mv.A = static_cast<const T *>(prm.hm) + qt * map_elems;
        if (FLIP) mv.B = static_cast<const T *>(prm.hmf) + (pt * K + (prm.perm ? __ldg(prm.perm + kt) : kt)) * map_elems;
    }
    return warp_blurred_log_single(mv, H, ty, tx, prm.gw, prm.radius, scratch, lane);
}

// running "first maximum" of one lane-private chain, tracked per float4 quad
struct Best {
    float v;
    int q;
    __device__ __forceinline__ void take(const float4 &s, int q4) {
        const float m = fmaxf(fmaxf(s.x, s.y), fmaxf(s.z, s.w));
        if (m > v) { v = m; q = q4; }
    }
};

template <bool FLIP, int RADIUS, bool HFQ, typename T>
__global__ void __launch_bounds__((FLIP && sizeof(T) == 4) ? 256 : 512, 1) heatmap_decode_kernel(const DecodeParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = prm.H, W = prm.W, K = prm.K;
    const int map_elems = H * W;
    const uint32_t map_bytes = (uint32_t)map_elems * (uint32_t)sizeof(T);
    const T *hm = static_cast<const T *>(prm.hm);
    const T *hmf = static_cast<const T *>(prm.hmf);
    constexpr int NB = FLIP ? 2 : 1;
    const int stages = prm.stages, warps = prm.warps;

    T *wbase = reinterpret_cast<T *>(smem_raw) + (size_t)warp * stages * NB * map_elems;
    unsigned char *after = smem_raw + (size_t)warps * stages * NB * map_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(after) + warp * stages;
    float *scratch = reinterpret_cast<float *>(after + (size_t)warps * stages * 8) + warp * kScratch;
    float *win = scratch + 64;       // (2r+3)^2 window of the averaged map around the arg-max

    const long long total = (long long)prm.P * K;
    const long long gwarp = (long long)blockIdx.x * warps + warp;
    const long long nwarps = (long long)gridDim.x * warps;

    if (lane == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
        fence_proxy_async();
    }
    __syncwarp();

    auto issue = [&](long long q, int s) {  // lane 0 only
        mbar_arrive_expect_tx(&bars[s], NB * map_bytes);
        T *dst = wbase + (size_t)s * NB * map_elems;
        bulk_g2s(dst, hm + q * map_elems, map_bytes, &bars[s]);
        if (FLIP) {
            const long long p = q / K;
            const int k = (int)(q - p * K);
            const int kk = prm.perm ? __ldg(prm.perm + k) : k;
            bulk_g2s(dst + map_elems, hmf + (p * K + kk) * map_elems, map_bytes, &bars[s]);
        }
    };

    if (lane == 0) {
        for (int s = 0; s < stages; ++s) {
            const long long q = gwarp + (long long)s * nwarps;
            if (q < total) issue(q, s);
        }
    }

    const int W4 = W >> 2;
    const int n4 = map_elems >> 2;
    const unsigned magic = prm.w4_magic;  // (q * magic) >> 16 == q / W4 for q < n4 (checked on the host)
    const int radius = RADIUS > 0 ? RADIUS : prm.radius;
    const int wn = 2 * radius + 3;
    const double *gw = prm.gw;

    int it = 0;
    for (long long q = gwarp; q < total; q += nwarps, ++it) {
        const int s = it % stages;
        const uint32_t parity = (uint32_t)((it / stages) & 1);
        mbar_wait(&bars[s], parity);

        if (prm.mode == kModeCopyOnly) {     // profiling aid: the bare bulk-TMA pipeline, no arithmetic
            __syncwarp();
            const long long qn = q + (long long)stages * nwarps;
            if (lane == 0 && qn < total) issue(qn, s);
            if (lane == 0) prm.scores[q] = ld1(wbase + (size_t)s * NB * map_elems);
            continue;
        }
        const T *A = wbase + (size_t)s * NB * map_elems;
        const T *B = A + map_elems;

        // One float4 of 2 x the flip-average (or of the map itself): element e of quad q4 is flat index
        // 4*q4 + e.  The * 0.5 is an exact power-of-two scaling, so the arg-max can be taken on the sums.
        auto quad = [&](int q4) -> float4 {
            float4 v = ld4(A, q4);
            if (FLIP) {
                const int r = (int)(((unsigned)q4 * magic) >> 16);
                const float4 m = ld4(B, 2 * r * W4 + W4 - 1 - q4);     // same row, mirrored quad, reversed lanes
                v.x += m.w;
                v.y += m.z;
                v.z += m.y;
                v.w += m.x;
            }
            return v;
        };

        // ---- phase 1: arg-max; 4 independent chains per lane, one compare per quad -----------------
        Best ch[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) ch[c] = Best{-INFINITY, INT_MAX};
        int q4 = lane;
        for (; q4 + 96 < n4; q4 += 128) {
            float4 v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = quad(q4 + 32 * c);
#pragma unroll
            for (int c = 0; c < 4; ++c) ch[c].take(v[c], q4 + 32 * c);
        }
        for (; q4 < n4; q4 += 32) ch[0].take(quad(q4), q4);
        float best = ch[0].v;
        int bq = ch[0].q;
#pragma unroll
        for (int c = 1; c < 4; ++c)
            if (ch[c].v > best || (ch[c].v == best && ch[c].q < bq)) { best = ch[c].v; bq = ch[c].q; }
        warp_argmax(best, bq);       // lowest quad holding the maximum == quad of the first maximum
        int bidx = 0;                // nothing compared greater than -inf: np.argmax gives 0
        if (bq != INT_MAX) {
            const float4 v = quad(bq);
            bidx = 4 * bq + (v.x == best ? 0 : (v.y == best ? 1 : (v.z == best ? 2 : 3)));
        } else {
            best = quad(0).x;        // all -inf (or NaN): the reference's max is element 0
        }
        if (FLIP) best *= 0.5f;
        const int ax = bidx % W, ay = bidx / W;
        const long long p = q / K;
        MapView<FLIP, T> mv{A, B, W};

        // ---- phase 2: everything that still needs the staged map ---------------------------------------
        const bool valid = best > 0.0f;
        float c00 = 0.f, cbr = 0.f, cbl = 0.f;          // DARK, score <= 0 path
        float se = 0.f, sxe = 0.f, sye = 0.f;           // soft-argmax sums
        float qdx = 0.f, qdy = 0.f;                     // quarter-offset differences
        float hfL[7];                                    // quirk Q6 taps: L11, L12, L21, L22, L00, L10, L01
        const bool hf_index = HFQ && prm.mode == SPP_DECODE_DARK;      // HFQ: separate instantiation, the default kernels carry none of it
        if (hf_index) {
            // HF post_dark_unbiased_data_processing (image_processing_vitpose.py:248-257) addresses its 7 taps in the
            // flattened, edge-padded batch through a FLOAT32 index (`index += stride * arange(...)` adds in place into
            // a float32 array): exact below 2^24, i.e. for the first 2^24 / ((W+2)(H+2)) maps of a call (5 084 maps =
            // 299 crops of 17 joints), rounded to even / multiples of 4 beyond — the taps then sit 1-2 padded columns
            // off.  Reproduced literally: same float32 index, same 7 offsets, each tap located by integer division
            // in the padded batch (so shifted taps wrap rows / maps exactly as numpy's flat indexing does).
            const float cxh = valid ? (float)ax : -1.0f, cyh = valid ? (float)ay : -1.0f;
            const float c32 = __fadd_rn(__fadd_rn(cxh, 1.0f), __fmul_rn(__fadd_rn(cyh, 1.0f), (float)(W + 2)));
            const long long stride = (long long)(W + 2) * (H + 2);
            const long long idx = (long long)(float)((double)c32 + (double)(stride * q));
            const int offs[7] = {0, 1, W + 2, W + 3, -(W + 3), -1, -(W + 2)};
#pragma unroll
            for (int t7 = 0; t7 < 7; ++t7) hfL[t7] = hf_tap_value<FLIP, T>(prm, A, B, q, idx + offs[t7], scratch, lane);
        } else if (prm.mode == SPP_DECODE_DARK) {
            if (valid) {
                // (2r+3)^2 window of the averaged map, 'reflect'-indexed like scipy's line extension
                for (int t = lane; t < wn * wn; t += 32) {
                    const int wr = t / wn, wc = t - wr * wn;
                    win[t] = mv.at(reflect_idx(ay - (radius + 1) + wr, H), reflect_idx(ax - (radius + 1) + wc, W));
                }
            } else {
                // HF indexes the flattened, edge-padded batch with coordinate -1: the centre taps land on
                // padded[0,0] of this map, the "minus" taps on the tail of the PREVIOUS map (numpy
                // negative indices wrap to the last map for q == 0).  Reproduced for parity.
                const long long qp = (q + total - 1) % total;
                const long long pp = qp / K;
                const int kp = (int)(qp - pp * K);
                MapView<FLIP, T> prev{hm + qp * map_elems, nullptr, W};
                if (FLIP) prev.B = hmf + (pp * K + (prm.perm ? __ldg(prm.perm + kp) : kp)) * map_elems;
                c00 = warp_blurred_log_single(mv, H, 0, 0, gw, radius, scratch, lane);
                cbr = warp_blurred_log_single(prev, H, H - 1, W - 1, gw, radius, scratch, lane);
                cbl = warp_blurred_log_single(prev, H, H - 1, 0, gw, radius, scratch, lane);
            }
        } else if (prm.mode == SPP_DECODE_SOFTARGMAX) {
            // softmax over the flattened map (max-subtracted), expected column / row, max probability
            const float half = FLIP ? 0.5f : 1.0f;
            for (int q4 = lane; q4 < n4; q4 += 32) {
                const float4 v = quad(q4);
                const int r = (int)(((unsigned)q4 * magic) >> 16);
                const int c4 = q4 - r * W4;
                // arguments are in [-range, 0]: the fast exponential (ex2.approx) is accurate to ~1e-6 relative here
                const float e0 = __expf(v.x * half - best), e1 = __expf(v.y * half - best), e2 = __expf(v.z * half - best),
                            e3 = __expf(v.w * half - best);
                const float c0 = (float)(c4 << 2);
                const float es = (e0 + e1) + (e2 + e3);
                se += es;
                sxe += e0 * c0 + e1 * (c0 + 1.f) + e2 * (c0 + 2.f) + e3 * (c0 + 3.f);
                sye += es * (float)r;
            }
        } else {
            const int px = valid ? ax : 0, py = valid ? ay : 0;
            if (px > 1 && px < W - 1 && py > 1 && py < H - 1) {
                qdx = mv.at(py, px + 1) - mv.at(py, px - 1);
                qdy = mv.at(py + 1, px) - mv.at(py - 1, px);
            }
        }

        // ---- release the stage: the next map streams in while this one is refined ------------------------
        __syncwarp();
        {
            const long long qn = q + (long long)stages * nwarps;
            if (lane == 0 && qn < total) issue(qn, s);
        }

        // ---- phase 3: refinement + back-projection (registers and per-warp scratch only) ------------------
        float out_x = 0.f, out_y = 0.f, out_s = best;

        if (prm.mode == SPP_DECODE_DARK) {
            // HF get_keypoint_predictions: coordinates -1 where score <= 0
            const float cx = valid ? (float)ax : -1.0f, cy = valid ? (float)ay : -1.0f;
            float L00, L01, L10, L11, L12, L21, L22;
            if (hf_index) {
                L11 = hfL[0]; L12 = hfL[1]; L21 = hfL[2]; L22 = hfL[3]; L00 = hfL[4]; L10 = hfL[5]; L01 = hfL[6];
            } else if (valid) {
                // vertical pass (scipy filters axis 0 first): T[r3][j] for the 3 tap rows x (2r+3) columns,
                // two outputs per lane issued together; window row of map row u is u - (ay - r - 1)
                const int t0 = lane, t1 = lane + 32;
                const bool has1 = t1 < 3 * wn;
                const int r30 = t0 / wn, j0 = t0 - r30 * wn;
                const int r31 = has1 ? t1 / wn : 0, j1 = has1 ? t1 - r31 * wn : 0;
                // np.pad(mode="edge") on the tap rows
                const int wy0 = clampi(ay + r30 - 1, 0, H - 1) - (ay - (radius + 1));
                const int wy1 = clampi(ay + r31 - 1, 0, H - 1) - (ay - (radius + 1));
                const double a0 = sym_filter<RADIUS>([&](int d) { return win[(wy0 + d) * wn + j0]; }, gw, radius);
                const double a1 = sym_filter<RADIUS>([&](int d) { return win[(wy1 + d) * wn + j1]; }, gw, radius);
                if (t0 < 3 * wn) scratch[t0] = (float)a0;    // fp32 intermediate between the axes
                if (has1) scratch[t1] = (float)a1;
                __syncwarp();
                float Lv = 0.f;
                if (lane < 9) {
                    const int r3 = lane / 3, c3 = lane - r3 * 3;
                    const int xc = clampi(ax + c3 - 1, 0, W - 1);
                    const float *row = scratch + r3 * wn + (xc - ax) + radius + 1;
                    const double acc = sym_filter<RADIUS>([&](int d) { return row[d]; }, gw, radius);
                    Lv = clip_log((float)acc);
                }
                L00 = __shfl_sync(FULL, Lv, 0);
                L01 = __shfl_sync(FULL, Lv, 1);
                L10 = __shfl_sync(FULL, Lv, 3);
                L11 = __shfl_sync(FULL, Lv, 4);
                L12 = __shfl_sync(FULL, Lv, 5);
                L21 = __shfl_sync(FULL, Lv, 7);
                L22 = __shfl_sync(FULL, Lv, 8);
                __syncwarp();                                // scratch / win are rewritten for the next map
            } else {
                L11 = L12 = L21 = L22 = c00;
                L10 = cbr;  // padded[index - 1]
                L01 = cbl;  // padded[index - (W + 2)]
                L00 = cbr;  // padded[index - (W + 3)]
            }
            const float i_ = L11, ix1 = L12, iy1 = L21, ix1y1 = L22, ix1_y1_ = L00, ix1_ = L10, iy1_ = L01;
            const float dx = 0.5f * (ix1 - ix1_);
            const float dy = 0.5f * (iy1 - iy1_);
            const float dxx = (ix1 - 2.0f * i_) + ix1_;
            const float dyy = (iy1 - 2.0f * i_) + iy1_;
            const float dxy = 0.5f * (((((((ix1y1 - ix1) - iy1) + i_) + i_) - ix1_) - iy1_) + ix1_y1_);
            const double eps = (double)FLT_EPSILON;
            const double ha = (double)dxx + eps, hb = (double)dxy, hd = (double)dyy + eps;
            const double inv_det = 1.0 / (ha * hd - hb * hb);
            const double sx = (hd * (double)dx - hb * (double)dy) * inv_det;
            const double sy = (ha * (double)dy - hb * (double)dx) * inv_det;
            out_x = (float)((double)cx - sx);
            out_y = (float)((double)cy - sy);
            if (prm.boxes) {
                // box_to_center_and_scale (HF:68-109) in double, as for Python-float boxes; then
                // transform_preds (HF:268-313) in fp32, left to right.
                const float4 bx = __ldg(reinterpret_cast<const float4 *>(prm.boxes) + p);
                float cxf, cyf, s0, s1;
                if (prm.flags & SPP_DECODE_FLAG_CENTER_SCALE) {
                    // HF keypoints_from_heatmaps(heatmaps, center, scale): (cx, cy, scale_x, scale_y) given
                    cxf = bx.x; cyf = bx.y;
                    s0 = __fmul_rn(bx.z, 200.0f);
                    s1 = __fmul_rn(bx.w, 200.0f);
                } else {
                    double bw = bx.z, bh = bx.w;
                    const double aspect = (double)prm.crop_w / (double)prm.crop_h;
                    cxf = (float)((double)bx.x + bw * 0.5);
                    cyf = (float)((double)bx.y + bh * 0.5);
                    if (bw > aspect * bh) bh = bw * 1.0 / aspect;
                    else if (bw < aspect * bh) bw = bh * aspect;
                    s0 = __fmul_rn(__fmul_rn((float)(bw / 200.0), 1.25f), 200.0f);
                    s1 = __fmul_rn(__fmul_rn((float)(bh / 200.0), 1.25f), 200.0f);
                }
                const float scale_x = __fdiv_rn(s0, (float)(W - 1));
                const float scale_y = __fdiv_rn(s1, (float)(H - 1));
                out_x = __fsub_rn(__fadd_rn(__fmul_rn(out_x, scale_x), cxf), __fmul_rn(s0, 0.5f));
                out_y = __fsub_rn(__fadd_rn(__fmul_rn(out_y, scale_y), cyf), __fmul_rn(s1, 0.5f));
            }
        } else if (prm.mode == SPP_DECODE_SOFTARGMAX) {
            se = warp_sum(se);
            sxe = warp_sum(sxe);
            sye = warp_sum(sye);
            const float inv = 1.0f / se;
            out_s = inv;  // exp(max - max) / sum
            out_x = (sxe * inv + 0.5f) / (float)W;
            out_y = (sye * inv + 0.5f) / (float)H;
            if (prm.boxes) {
                const float4 bx = __ldg(reinterpret_cast<const float4 *>(prm.boxes) + p);  // x1 y1 x2 y2
                const float bw = bx.z - bx.x, bh = bx.w - bx.y;
                if (prm.flags & SPP_DECODE_FLAG_SCALE_SCORE) {
                    const float wgt = fminf(fmaxf(sqrtf(bw * bh) / 96.0f, 0.5f), 2.0f);
                    out_s = out_s * wgt;
                }
                if (prm.flags & SPP_DECODE_FLAG_BACKPROJECT) {
                    out_x = __fadd_rn(__fmul_rn(out_x, bw), bx.x);
                    out_y = __fadd_rn(__fmul_rn(out_y, bh), bx.y);
                }
            }
        } else {  // SPP_DECODE_QUARTER
            float cx = valid ? (float)ax : 0.f, cy = valid ? (float)ay : 0.f;
            cx += 0.25f * (float)((qdx > 0.f) - (qdx < 0.f));
            cy += 0.25f * (float)((qdy > 0.f) - (qdy < 0.f));
            out_x = cx;
            out_y = cy;
            if (prm.boxes) {
                // centre/scale of datamodule_v2.py:119-129 (double), inverse similarity in fp32
                const float4 bx = __ldg(reinterpret_cast<const float4 *>(prm.boxes) + p);  // x y w h
                const double aspect = (double)prm.crop_w / (double)prm.crop_h;
                double ccx = (double)bx.x + (double)bx.z * 0.5, ccy = (double)bx.y + (double)bx.w * 0.5;
                if (prm.flags & SPP_DECODE_FLAG_CENTER_SCALE) {   // get_final_preds(heatmaps, center, scale)
                    ccx = bx.x; ccy = bx.y;
                } else if (aspect > 1.0) ccx += (double)bx.z * 0.5 * (aspect - 1.0);
                else ccy += (double)bx.w * 0.5 * (1.0 / aspect - 1.0);
                const float sw = bx.z;
                const float rr = __fdiv_rn(sw, (float)W);
                out_x = __fsub_rn(__fadd_rn(__fmul_rn(cx, rr), (float)ccx), __fmul_rn(sw, 0.5f));
                out_y = __fsub_rn(__fadd_rn(__fmul_rn(cy, rr), (float)ccy), __fmul_rn(__fmul_rn(rr, (float)H), 0.5f));
            }
        }

        if (lane == 0) {
            reinterpret_cast<float2 *>(prm.kpts)[q] = make_float2(out_x, out_y);
            prm.scores[q] = out_s;
            if (prm.amax) prm.amax[q] = bidx;
        }
    }
}

template <typename T, bool FLIP, int RADIUS, bool HFQ = false>
int launch_decode(const DecodeParams &prm, unsigned grid, size_t smem, cudaStream_t st) {
    // per device and per context: set on every launch (about a microsecond; legal during stream capture)
    SPP_CHECK_CUDA(cudaFuncSetAttribute(heatmap_decode_kernel<FLIP, RADIUS, HFQ, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    heatmap_decode_kernel<FLIP, RADIUS, HFQ, T><<<grid, prm.warps * 32, smem, st>>>(prm);
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}

}  // namespace

}  // namespace spp

namespace spp {
namespace {
template <typename T>
int heatmap_decode_impl(const T *hm, const T *hm_flipped, const int *perm, int p, int k, int h, int w,
                        const float *boxes, int mode, int flags, int kernel, int crop_h, int crop_w,
                        float *keypoints, float *scores, int *argmax, spp_stream_t stream) {
    SPP_CHECK_ARG(p >= 0 && k > 0 && h > 0 && w > 0, "heatmap_decode: bad shape p=%d k=%d h=%d w=%d", p, k, h, w);
    SPP_CHECK_ARG(p == 0 || (hm && keypoints && scores), "heatmap_decode: hm, keypoints and scores must be non-null");
    SPP_CHECK_ARG(w % 4 == 0, "heatmap_decode: heatmap width must be a multiple of 4 (got %d)", w);
    SPP_CHECK_ARG(((size_t)h * w * sizeof(T)) % 16 == 0, "heatmap_decode: a %dx%d map of %zu-byte elements is not a multiple of 16 bytes", h, w, sizeof(T));
    SPP_CHECK_ARG((mode >= SPP_DECODE_DARK && mode <= SPP_DECODE_QUARTER) || mode == kModeCopyOnly, "heatmap_decode: unknown mode %d", mode);
    SPP_CHECK_ARG(kernel >= 3 && kernel <= 2 * kMaxRadius + 1 && (kernel & 1), "heatmap_decode: kernel must be odd in 3..%d (got %d)",
                  2 * kMaxRadius + 1, kernel);
    SPP_CHECK_ARG(crop_h > 0 && crop_w > 0, "heatmap_decode: bad crop size");
    SPP_CHECK_ARG((reinterpret_cast<uintptr_t>(hm) & 15) == 0 && (reinterpret_cast<uintptr_t>(hm_flipped) & 15) == 0,
                  "heatmap_decode: heatmaps must be 16-byte aligned");
    SPP_CHECK_ARG(!boxes || (reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "heatmap_decode: boxes must be 16-byte aligned");
    SPP_CHECK_ARG((reinterpret_cast<uintptr_t>(keypoints) & 7) == 0, "heatmap_decode: keypoints must be 8-byte aligned");
    if (p == 0) return SPP_OK;

    DecodeParams prm{};
    prm.hm = hm; prm.hmf = hm_flipped; prm.perm = perm; prm.boxes = boxes;
    prm.kpts = keypoints; prm.scores = scores; prm.amax = argmax;
    prm.P = p; prm.K = k; prm.H = h; prm.W = w;
    prm.mode = mode; prm.flags = flags; prm.radius = (kernel - 1) / 2; prm.crop_h = crop_h; prm.crop_w = crop_w;
    {   // scipy.ndimage._filters._gaussian_kernel1d(sigma=0.8, order=0, radius)
        const double sigma = 0.8;
        double phi[kMaxRadius + 1], sum = 0.0;
        for (int x = 0; x <= prm.radius; ++x) phi[x] = std::exp(-0.5 / (sigma * sigma) * (double)(x * x));
        for (int x = -prm.radius; x <= prm.radius; ++x) sum += phi[x < 0 ? -x : x];
        for (int x = 0; x <= prm.radius; ++x) prm.gw[x] = phi[x] / sum;
    }

    SPP_CHECK_ARG(h >= prm.radius + 2 && w >= prm.radius + 2, "heatmap_decode: map %dx%d too small for kernel %d", h, w, kernel);
    const bool flip = hm_flipped != nullptr;
    const size_t stage_bytes = (size_t)(flip ? 2 : 1) * h * w * sizeof(T);
    const size_t budget = 200 * 1024;
    const int slots = (int)(budget / stage_bytes);
    SPP_CHECK_ARG(slots >= 1, "heatmap_decode: a %dx%d map does not fit the shared-memory pipeline", h, w);
    // 8 warps per SM hide the ALU latency of the scan; whatever shared memory is left deepens each
    // warp's private ring (flip test, 64x48: 8 warps x 1 stage x 24 KB).
    // (without the flip test a map is half the bytes for the same refinement work: 16 warps x 1 stage)
    // (bf16 maps with the flip test are the same 12 KB per stage as fp32 maps without it: 16 warps as well)
    const int max_warps = (flip && sizeof(T) == 4) ? 8 : 16;
    int warps = slots < max_warps ? slots : max_warps;
    int stages = slots / warps;
    if (stages > 4) stages = 4;
    {   // tuning knobs (profiling only): SPP_HM_WARPS / SPP_HM_STAGES override the split of the slots
        static int env_w = -1, env_s = -1;
        if (env_w < 0) {
            const char *ew = getenv("SPP_HM_WARPS"), *es = getenv("SPP_HM_STAGES");
            env_w = ew ? atoi(ew) : 0;
            env_s = es ? atoi(es) : 0;
        }
        if (env_w > 0 && env_s > 0 && env_w * env_s <= slots && env_w <= max_warps) { warps = env_w; stages = env_s; }
    }
    prm.warps = warps; prm.stages = stages;
    {   // q / W4 by multiply-shift, verified for every quad index of a map
        const unsigned w4 = (unsigned)(w / 4), n4 = (unsigned)(h * w / 4);
        unsigned magic = (65536u + w4 - 1) / w4;
        bool ok = n4 < 65536u;
        for (unsigned q = 0; ok && q < n4; ++q) ok = ((q * magic) >> 16) == q / w4;
        SPP_CHECK_ARG(ok, "heatmap_decode: unsupported map shape %dx%d", h, w);
        prm.w4_magic = magic;
    }
    const size_t smem = (size_t)warps * stages * stage_bytes + (size_t)warps * stages * 8 + (size_t)warps * kScratch * 4;

    const long long total = (long long)p * k;
    int sms = sm_count();
    if (sms <= 0) return SPP_ERR_CUDA;
    long long grid = (total + warps - 1) / warps;
    if (grid > sms) grid = sms;
    {   // spp_set_launch_limit(SPP_LIMIT_HEATMAP_CTAS): leave SMs free for a kernel that needs whole SMs beside this one
        const int lim = launch_limit(0);
        if (lim > 0 && grid > lim) grid = lim;
    }

    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mode == SPP_DECODE_DARK && (flags & SPP_DECODE_FLAG_HF_F32_INDEX))      // quirk Q6: run-time radius, own instantiation
        return flip ? launch_decode<T, true, 0, true>(prm, (unsigned)grid, smem, st) : launch_decode<T, false, 0, true>(prm, (unsigned)grid, smem, st);
    if (flip) return prm.radius == 5 ? launch_decode<T, true, 5>(prm, (unsigned)grid, smem, st) : launch_decode<T, true, 0>(prm, (unsigned)grid, smem, st);
    return prm.radius == 5 ? launch_decode<T, false, 5>(prm, (unsigned)grid, smem, st) : launch_decode<T, false, 0>(prm, (unsigned)grid, smem, st);
}
}  // namespace
}  // namespace spp

extern "C" int spp_heatmap_decode(const float *hm, const float *hm_flipped, const int *perm, int p, int k, int h, int w,
                                  const float *boxes, int mode, int flags, int kernel, int crop_h, int crop_w,
                                  float *keypoints, float *scores, int *argmax, spp_stream_t stream) {
    return spp::heatmap_decode_impl<float>(hm, hm_flipped, perm, p, k, h, w, boxes, mode, flags, kernel, crop_h, crop_w, keypoints, scores,
                                           argmax, stream);
}

extern "C" int spp_heatmap_decode_bf16(const uint16_t *hm, const uint16_t *hm_flipped, const int *perm, int p, int k, int h, int w,
                                       const float *boxes, int mode, int flags, int kernel, int crop_h, int crop_w,
                                       float *keypoints, float *scores, int *argmax, spp_stream_t stream) {
    return spp::heatmap_decode_impl<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16 *>(hm), reinterpret_cast<const __nv_bfloat16 *>(hm_flipped),
                                                   perm, p, k, h, w, boxes, mode, flags, kernel, crop_h, crop_w, keypoints, scores, argmax, stream);
}

