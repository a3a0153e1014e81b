// Detection-head decode + class-aware NMS.
//
// Replaces:
//   make_anchors          training/yolopt/util.py:85-96
//   DFL / Head.forward    training/yolopt/nets/nn.py:222-225, 255-270 (eval branch)
//   non_max_suppression   training/yolopt/util.py:123-169 (+ wh2xy :76-82; torchvision.ops.nms :162)
//
// Kernels
//   head_decode_kernel     full decode to [B, 4+nc, A] (drop-in for Head.forward); one thread per
//                          anchor, every channel plane read as coalesced 128 B lines.
//   nms_kernel<2>          THE fused path, one CTA per image, one launch per head:
//                          (0) candidate scan over the image's class planes only (logit pre-filter, exact fp32
//                              sigmoid > conf), warp-aggregated append of 64-bit sort keys (~score | anchor*nc+cls);
//                              the 64 DFL planes are touched just for the candidates (half-warp per candidate,
//                              lane = DFL bin), whose decoded boxes are parked in a per-image [A] float4 table;
//                          (1) bitonic sort of the keys (score descending, candidate key ascending — a stable
//                              order);
//                          (2) greedy suppression in that order against the boxes kept so far.  A box survives
//                              iff no earlier KEPT box has IoU > thr, so only n * kept (<= n * max_det) IoUs are
//                              needed instead of the n^2/2 of a full bitmask; every warp clears, with one ballot
//                              per 32-candidate bitmask word, the later candidates the current box suppresses.
//                              The loop stops after max_det kept boxes.
//   cand_scan_kernel + cand_decode_kernel + nms_kernel<1>   the same work as three launches (SPP_DET_SPLIT=1; kept
//                          as the profiling baseline: two full-machine latency-bound launches in front of the NMS).
//   cand_decoded_kernel + nms_kernel<0>   candidate pass + NMS over an already decoded [B, 4+nc, A] tensor
//                          (drop-in for non_max_suppression).
// IoU arithmetic is fp32 with round-to-nearest intrinsics (never contracted to FMA): it must agree bit
// for bit with torchvision's CPU loop, which is what the reference's keep indices come from.
#include "spp_common.cuh"

#include <atomic>
#include <climits>
#include <cstdlib>
#include <cmath>

namespace spp {

namespace {

constexpr int kDfl = 16;
constexpr int kNmsThreads = 1024;
constexpr int kSortSmemMax = 8192;  // keys sorted in shared memory up to this (padded) count
constexpr int kBoxSmemMax = 2048;   // sorted boxes kept in shared memory (the rest is re-gathered)
constexpr int kAliveWords = 1024;   // one alive bit per candidate: max_nms <= 32768

// Per level: the 64 DFL planes and the nc class planes of image b start at box[l] + b * bs_box[l] and
// cls[l] + b * bs_cls[l] (elements).  One concatenated map [B, 64+nc, H, W] (nn.py:257) gives
// cls = box + 64*H*W, bs_box = bs_cls = (64+nc)*H*W; the un-concatenated conv outputs of nn.py:256-257
// ([B, 64, H, W] and [B, nc, H, W]) give bs_box = 64*H*W, bs_cls = nc*H*W.
struct Levels {
    const float *box[SPP_MAX_LEVELS];
    const float *cls[SPP_MAX_LEVELS];
    long long bs_box[SPP_MAX_LEVELS], bs_cls[SPP_MAX_LEVELS];
    int h[SPP_MAX_LEVELS], w[SPP_MAX_LEVELS], off[SPP_MAX_LEVELS + 1];
    float stride[SPP_MAX_LEVELS];
    int n, A;
};

// Resolved pyramid level of one anchor.  Selected with compile-time indices only, so the by-value
// kernel parameter stays in the constant bank (dynamic indexing would copy it to local memory per thread).
struct LevelRef {
    const float *box;
    const float *cls;
    long long bs_box, bs_cls;
    int w, hw, i;
    float stride;
};
__device__ __forceinline__ LevelRef find_level(const Levels &lv, int a) {
    LevelRef r{lv.box[0], lv.cls[0], lv.bs_box[0], lv.bs_cls[0], lv.w[0], lv.h[0] * lv.w[0], a, lv.stride[0]};
#pragma unroll
    for (int i = 1; i < SPP_MAX_LEVELS; ++i)
        if (i < lv.n && a >= lv.off[i])
            r = LevelRef{lv.box[i], lv.cls[i], lv.bs_box[i], lv.bs_cls[i], lv.w[i], lv.h[i] * lv.w[i], a - lv.off[i], lv.stride[i]};
    return r;
}

__device__ __forceinline__ float sigmoidf_ref(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// DFL expectation for the four sides + dist2bbox; returns (cx, cy, w, h) in pixels.  `base` points at
// channel 0 of this (image, level) for anchor i; `hw` is the plane stride.
// softmax expectation as one ratio, sum_j j*e_j / sum_j e_j with e_j = exp(v_j - max): 16 loads, 16 fast
// exponentials and one division per side (the reference normalises every bin first; the two agree to
// ~1e-7 relative, far inside the 1e-3 tolerance on box coordinates).
__device__ __forceinline__ float4 decode_box(const float *base, int hw, float ax, float ay, float stride) {
    float d[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        float v[kDfl];
#pragma unroll
        for (int j = 0; j < kDfl; ++j) v[j] = __ldg(base + (size_t)(s * kDfl + j) * hw);
        float m = v[0];
#pragma unroll
        for (int j = 1; j < kDfl; ++j) m = fmaxf(m, v[j]);
        float sum = 0.f, num = 0.f;
#pragma unroll
        for (int j = 0; j < kDfl; ++j) {
            const float e = __expf(v[j] - m);
            sum += e;
            num = fmaf((float)j, e, num);
        }
        d[s] = __fdiv_rn(num, sum);
    }
    const float x1 = __fsub_rn(ax, d[0]), y1 = __fsub_rn(ay, d[1]);
    const float x2 = __fadd_rn(ax, d[2]), y2 = __fadd_rn(ay, d[3]);
    float4 r;
    r.x = __fmul_rn(__fmul_rn(__fadd_rn(x1, x2), 0.5f), stride);
    r.y = __fmul_rn(__fmul_rn(__fadd_rn(y1, y2), 0.5f), stride);
    r.z = __fmul_rn(__fsub_rn(x2, x1), stride);
    r.w = __fmul_rn(__fsub_rn(y2, y1), stride);
    return r;
}

// wh2xy, util.py:76-82
__device__ __forceinline__ float4 wh2xy(float4 c) {
    float4 r;
    const float hw = __fmul_rn(c.z, 0.5f), hh = __fmul_rn(c.w, 0.5f);
    r.x = __fsub_rn(c.x, hw);
    r.y = __fsub_rn(c.y, hh);
    r.z = __fadd_rn(c.x, hw);
    r.w = __fadd_rn(c.y, hh);
    return r;
}

__global__ void __launch_bounds__(256) head_decode_kernel(const Levels lv, int nc, float *__restrict__ out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (a >= lv.A) return;
    const LevelRef lr = find_level(lv, a);
    const int i = lr.i, w = lr.w, hw = lr.hw;
    const int y = i / w, x = i - y * w;
    const float *base = lr.box + (size_t)b * lr.bs_box + i;
    const float *cbase = lr.cls + (size_t)b * lr.bs_cls + i;
    const float4 box = decode_box(base, hw, (float)x + 0.5f, (float)y + 0.5f, lr.stride);
    float *o = out + (size_t)b * (4 + nc) * lv.A + a;
    o[0] = box.x;
    o[(size_t)lv.A] = box.y;
    o[(size_t)2 * lv.A] = box.z;
    o[(size_t)3 * lv.A] = box.w;
    for (int j = 0; j < nc; ++j) o[(size_t)(4 + j) * lv.A] = sigmoidf_ref(__ldg(cbase + (size_t)j * hw));
}

__device__ __forceinline__ unsigned long long make_sort_key(float score, unsigned cand) {
    // ascending sort on this key == score descending, candidate key ascending (scores are > 0)
    return ((unsigned long long)(~__float_as_uint(score)) << 32) | cand;
}

// warp-aggregated append of one candidate per participating lane
__device__ __forceinline__ void append_candidate(bool is_cand, unsigned long long key, int *count, unsigned long long *keys,
                                                 int cap) {
    const unsigned m = __ballot_sync(FULL, is_cand);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(FULL, base, leader);
    const int slot = base + __popc(m & ((1u << lane) - 1u));
    if (is_cand && slot < cap) keys[slot] = key;
}

__global__ void __launch_bounds__(256) cand_decoded_kernel(const float *__restrict__ pred, int nc, int A, float conf, int cap,
                                                           int cap_pad, int *counts, unsigned long long *keys) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    const float *cls = pred + ((size_t)b * (4 + nc) + 4) * A;
    unsigned long long *k = keys + (size_t)b * cap_pad;
    for (int j = 0; j < nc; ++j) {
        float s = 0.f;
        bool c = false;
        if (a < A) {
            s = __ldg(cls + (size_t)j * A + a);
            c = s > conf;
        }
        append_candidate(c, make_sort_key(s, (unsigned)(a * nc + j)), counts + b, k, cap);
    }
}

// Candidate scan over the raw class planes only.  sigmoid is monotone, so anchors whose logit is below
// logit(conf) - 0.05 are rejected without evaluating it; the exact fp32 test sigmoid(x) > conf decides
// the rest.  Candidates are appended as sort keys; their boxes are decoded by cand_decode_kernel.
constexpr int kScanPerThread = 8;        // anchors per thread: 8 independent loads in flight, 1/8 of the CTAs
__global__ void __launch_bounds__(256) cand_scan_kernel(const Levels lv, int nc, float conf, float logit_lo, int cap, int cap_pad,
                                                        int *counts, unsigned long long *keys) {
    const int a0 = blockIdx.x * (256 * kScanPerThread) + threadIdx.x;
    const int b = blockIdx.y;
    unsigned long long *k = keys + (size_t)b * cap_pad;
    for (int j = 0; j < nc; ++j) {
        float x[kScanPerThread];
#pragma unroll
        for (int u = 0; u < kScanPerThread; ++u) {
            const int a = a0 + u * 256;
            x[u] = -INFINITY;
            if (a < lv.A) {
                const LevelRef lr = find_level(lv, a);
                x[u] = __ldg(lr.cls + (size_t)b * lr.bs_cls + (size_t)j * lr.hw + lr.i);
            }
        }
        // one warp-aggregated append for the 8 anchors of every lane: a single atomicAdd round trip per warp
        float sc[kScanPerThread];
        unsigned ball[kScanPerThread];
        int total = 0;
#pragma unroll
        for (int u = 0; u < kScanPerThread; ++u) {
            sc[u] = 0.f;
            bool c = false;
            if (x[u] > logit_lo) {               // -inf past the last anchor
                sc[u] = sigmoidf_ref(x[u]);
                c = sc[u] > conf;
            }
            ball[u] = __ballot_sync(FULL, c);
            total += __popc(ball[u]);
        }
        if (total) {                             // warp-uniform
            const int lane = threadIdx.x & 31;
            int base = 0;
            if (lane == 0) base = atomicAdd(counts + b, total);
            base = __shfl_sync(FULL, base, 0);
#pragma unroll
            for (int u = 0; u < kScanPerThread; ++u) {
                if ((ball[u] >> lane) & 1u) {
                    const int slot = base + __popc(ball[u] & ((1u << lane) - 1u));
                    if (slot < cap) k[slot] = make_sort_key(sc[u], (unsigned)((a0 + u * 256) * nc + j));
                }
                base += __popc(ball[u]);
            }
        }
    }
}

// DFL decode of ONE candidate anchor by one half-warp (both halves of a warp call this together: the shuffles
// use the full mask with xor offsets < 16).  Lane `sub` owns DFL bin `sub` of all four sides (4 loads in flight
// per lane, 64 per candidate); softmax max / sum and the expectation are 16-lane xor-shuffle reductions.
// The xyxy box is written to img_boxes[anchor] by lane 0 of the half-warp.
struct DflLoad {
    float v[4];          // this lane's DFL bin of the four sides
    int i, w;            // anchor index inside its level, level width
    float stride;
};
__device__ __forceinline__ DflLoad dfl_load(const Levels &lv, int b, int anchor, int sub) {
    DflLoad d;
    const LevelRef lr = find_level(lv, anchor);
    const float *basep = lr.box + (size_t)b * lr.bs_box + lr.i;
#pragma unroll
    for (int sd = 0; sd < 4; ++sd) d.v[sd] = __ldg(basep + (size_t)(sd * kDfl + sub) * lr.hw);
    d.i = lr.i; d.w = lr.w; d.stride = lr.stride;
    return d;
}
__device__ __forceinline__ void dfl_finish(const DflLoad &ld, int anchor, bool valid, int sub, float4 *img_boxes) {
    const DflLoad &lr = ld;
    float m[4], sum[4], d[4];
#pragma unroll
    for (int sd = 0; sd < 4; ++sd) m[sd] = ld.v[sd];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1)
#pragma unroll
        for (int sd = 0; sd < 4; ++sd) m[sd] = fmaxf(m[sd], __shfl_xor_sync(FULL, m[sd], o));
#pragma unroll
    for (int sd = 0; sd < 4; ++sd) {
        sum[sd] = __expf(ld.v[sd] - m[sd]);
        d[sd] = (float)sub * sum[sd];
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1)
#pragma unroll
        for (int sd = 0; sd < 4; ++sd) {
            sum[sd] += __shfl_xor_sync(FULL, sum[sd], o);
            d[sd] += __shfl_xor_sync(FULL, d[sd], o);
        }
#pragma unroll
    for (int sd = 0; sd < 4; ++sd) d[sd] = __fdiv_rn(d[sd], sum[sd]);      // sum_j j*e_j / sum_j e_j
    if (valid && sub == 0) {
        const int y = lr.i / lr.w, x = lr.i - y * lr.w;
        const float ax = (float)x + 0.5f, ay = (float)y + 0.5f;
        const float x1 = __fsub_rn(ax, d[0]), y1 = __fsub_rn(ay, d[1]);
        const float x2 = __fadd_rn(ax, d[2]), y2 = __fadd_rn(ay, d[3]);
        float4 r;
        r.x = __fmul_rn(__fmul_rn(__fadd_rn(x1, x2), 0.5f), lr.stride);
        r.y = __fmul_rn(__fmul_rn(__fadd_rn(y1, y2), 0.5f), lr.stride);
        r.z = __fmul_rn(__fsub_rn(x2, x1), lr.stride);
        r.w = __fmul_rn(__fsub_rn(y2, y1), lr.stride);
        img_boxes[anchor] = wh2xy(r);
    }
}
__device__ __forceinline__ void decode_candidate(const Levels &lv, int b, int anchor, bool valid, int sub, float4 *img_boxes) {
    const DflLoad ld = dfl_load(lv, b, anchor, sub);
    dfl_finish(ld, anchor, valid, sub, img_boxes);
}

// Box decode for the candidates only: one half-warp per (image, candidate).  Lane j of the half-warp
// owns DFL bin j of all four sides (4 loads in flight per lane, 64 per candidate); the softmax
// max / sum and the expectation are 16-lane xor-shuffle reductions, four sides at a time.
__global__ void __launch_bounds__(256) cand_decode_kernel(const Levels lv, int nc, int batch, int cap, int cap_pad,
                                                          const int *__restrict__ counts, const unsigned long long *__restrict__ keys,
                                                          float4 *__restrict__ boxes) {
    extern __shared__ int pre[];                                  // [batch + 1] exclusive prefix of the counts
    const int lane = threadIdx.x & 31, sub = lane & 15;
    const int half_id = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 4);
    const int nhalf = (int)((gridDim.x * blockDim.x) >> 4);
    for (int b = threadIdx.x; b < batch; b += blockDim.x) pre[b + 1] = min(__ldg(counts + b), cap);
    __syncthreads();
    if (threadIdx.x == 0) {
        pre[0] = 0;
        for (int b = 0; b < batch; ++b) pre[b + 1] += pre[b];
    }
    __syncthreads();
    const int total = pre[batch];                                 // items = candidates of all images, back to back
    // both halves of a warp iterate together (full-mask shuffles); an idle half works on a dummy item
    for (int base = (half_id & ~1); base < total; base += nhalf) {
        const int t = base + (half_id & 1);
        const bool valid = t < total;
        int lo = 0, hi = batch;                                   // largest b with pre[b] <= t
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (pre[mid] <= (valid ? t : 0)) lo = mid; else hi = mid;
        }
        const int b = lo, slot = (valid ? t : 0) - pre[lo];
        const unsigned cand = valid ? (unsigned)(keys[(size_t)b * cap_pad + slot] & 0xffffffffu) : 0u;
        const int anchor = (int)(cand / (unsigned)nc);
        decode_candidate(lv, valid ? b : 0, anchor, valid, sub, boxes + (size_t)b * lv.A);
    }
}

struct NmsParams {
    const float *pred;          // decoded path
    float4 *boxes;              // raw path: [B, A] xyxy (written by cand_decode_kernel, or by the fused kernel itself)
    Levels lv;                  // fused path: the raw head maps
    float conf, logit_lo;       // fused path: candidate threshold and its logit pre-filter
    const int *counts;
    unsigned long long *keys;   // [B, cap_pad]
    int nc, A, cap, cap_pad;
    float iou, max_wh;
    int max_det, max_nms;
    int compact_pct;            // compaction when the alive share of the not-yet-visited candidates drops to this percentage
    float *out_dets;
    int *out_count, *out_keys;
};

__device__ __forceinline__ float box_inter(const float4 &a, const float4 &b) {
    const float w = fmaxf(0.0f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
    const float h = fmaxf(0.0f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
    return __fmul_rn(w, h);
}
// inter / (area_a + area_b - inter) > thr, with the result of the IEEE division torchvision's CPU loop performs — but the
// division itself is only executed when the ratio lies within 2^-20 (relative) of the threshold:
//   u = (area_a + area_b) - inter exactly as the reference forms it, t = thr * u (one rounding, 2^-24).
//   inter > t * (1 + 2^-20)  =>  inter / u > thr * (1 + 2^-21)  =>  RN(inter / u) > thr;   symmetric for "<".
// Valid for u > 0 and thr > 0; anything else (empty or degenerate boxes: u <= 0, NaN) takes the division.  The division
// (a ~20-instruction sequence with a range check) was the bulk of the apply phase at crowd density, where most kept boxes
// intersect some lane of every later word.
__device__ __forceinline__ bool ratio_gt(float inter, float area_a, float area_b, float thr, bool live) {
    const float u = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    const float t = __fmul_rn(thr, u);
    const bool fast = live && u > 0.0f && thr > 0.0f;
    const bool surely = fast && inter > __fmul_rn(t, 1.00000095367431640625f);         // 1 + 2^-20
    const bool surely_not = fast && inter < __fmul_rn(t, 0.99999904632568359375f);     // 1 - 2^-20
    bool r = surely;
    if (live && !surely && !surely_not) r = __fdiv_rn(inter, u) > thr;                 // rare: the exact IEEE quotient decides
    return r;
}
// Warp-level "which alive lanes does box `a` suppress": the division is skipped when no alive lane intersects the box at
// all (an empty intersection gives IoU 0, or NaN for two empty boxes — never > thr for thr >= 0)
__device__ __forceinline__ unsigned suppress_ballot(const float4 &a, float area_a, const float4 &bj, float aj, bool alive_lane, float thr) {
    const float inter = box_inter(a, bj);
    const bool touch = alive_lane && (inter > 0.0f || thr < 0.0f);      // thr < 0: IoU 0 suppresses too, no shortcut
    if (!__any_sync(FULL, touch)) return 0u;
    return __ballot_sync(FULL, ratio_gt(inter, area_a, aj, thr, touch));
}

__device__ void bitonic_sort(unsigned long long *d, int npad) {
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (npad >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool up = (i & k) == 0;
                const unsigned long long x = d[i], y = d[p];
                if ((x > y) == up) {
                    d[i] = y;
                    d[p] = x;
                }
            }
            __syncthreads();
        }
    }
}

// Bitonic sort of up to NT keys, one per thread: compare-exchange partners closer than 32 are
// reached with warp shuffles (no barrier, no shared memory), the 10 longer strides go through `xch`.
template <int NT>
__device__ unsigned long long block_bitonic_reg(unsigned long long key, unsigned long long *xch) {
    const int tid = threadIdx.x;
    for (int k = 2; k <= NT; k <<= 1) {
        const bool up = (tid & k) == 0;
        for (int j = k >> 1; j > 0; j >>= 1) {
            unsigned long long other;
            if (j >= 32) {
                xch[tid] = key;
                __syncthreads();
                other = xch[tid ^ j];
                __syncthreads();
            } else {
                other = __shfl_xor_sync(FULL, key, j);
            }
            const bool lower = (tid & j) == 0;
            // the lower index of a pair keeps the smaller key in an ascending run, the larger otherwise
            const bool take_min = lower == up;
            key = take_min ? (key < other ? key : other) : (key > other ? key : other);
        }
    }
    return key;
}

// Bitonic sort of NT * E keys, E consecutive keys per thread: compare-exchange distances below E stay in registers,
// distances up to 16 threads go through warp shuffles, only the longer ones (15 of the 78 steps at 4 096 keys) through
// `xch` (NT * E keys, conflict-free [e][thread] layout) with two block barriers each.
template <int NT, int E>
__device__ void block_bitonic_multi(unsigned long long (&key)[E], unsigned long long *xch) {
    const int tid = threadIdx.x;
    auto cmpx = [](unsigned long long &a, unsigned long long &b, bool up) {      // a at the lower index
        const unsigned long long lo = a < b ? a : b, hi = a < b ? b : a;
        a = up ? lo : hi;
        b = up ? hi : lo;
    };
    for (int K = 2; K <= NT * E; K <<= 1) {
        for (int J = K >> 1; J > 0; J >>= 1) {
            if (J < E) {                                       // compile-time register indices only (no local memory)
#pragma unroll
                for (int JJ = 1; JJ < E; JJ <<= 1) {
                    if (JJ != J) continue;
#pragma unroll
                    for (int e = 0; e < E; ++e)
                        if ((e & JJ) == 0) cmpx(key[e], key[e | JJ], ((tid * E + e) & K) == 0);
                }
            } else {
                const int dt = J / E;                          // partner thread distance
                unsigned long long other[E];
                if (dt < 32) {
#pragma unroll
                    for (int e = 0; e < E; ++e) other[e] = __shfl_xor_sync(FULL, key[e], dt);
                } else {
#pragma unroll
                    for (int e = 0; e < E; ++e) xch[e * NT + tid] = key[e];
                    __syncthreads();
#pragma unroll
                    for (int e = 0; e < E; ++e) other[e] = xch[e * NT + (tid ^ dt)];
                    __syncthreads();
                }
                const bool lower = (tid & dt) == 0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const bool up = ((tid * E + e) & K) == 0;
                    const bool take_min = lower == up;
                    key[e] = take_min ? (key[e] < other[e] ? key[e] : other[e]) : (key[e] > other[e] ? key[e] : other[e]);
                }
            }
        }
    }
}

// One CTA per image.
//   1. bitonic sort of the 64-bit keys: in registers + shuffles up to one key per thread, in shared
//      memory up to kSortSmemMax, in the workspace beyond;
//   2. sorted, class-offset boxes + areas are staged in shared memory (first kBoxSmemMax; beyond that they
//      are re-gathered on the fly), one "alive" bit per candidate;
//   3. greedy suppression one 32-candidate word at a time: every warp computes one suppression row of the word (ballot),
//      warp 0 resolves the word with bit operations and publishes the kept lanes, every warp applies those kept boxes to
//      the later words it owns (four boxes in flight).  Three block barriers per word; stops after max_det.
//   MODE 0: candidates from cand_decoded_kernel, boxes from the decoded [B, 4+nc, A] tensor;
//   MODE 1: candidates from cand_scan_kernel, boxes from cand_decode_kernel's table;
//   MODE 2: FUSED — step 0 of the CTA is the candidate scan over its image's class planes and the DFL decode of
//           the candidates (half-warp per candidate, two in flight per half-warp): one launch per head instead
//           of three, and one CTA per image instead of two full-machine latency-bound launches in front of it.
//   CFG: launch configuration.  NmsBig (1 024 threads, 112 KB of shared memory) handles any candidate count;
//   NmsSmall (512 threads, ~15 KB) is chosen by the host when the caller bounds the candidate list to <= 512 per image
//   (max_candidates): same code, same results, but a CTA that fits beside the resident CTAs of the bandwidth-bound
//   kernels (heatmap decode: 206 KB of shared memory per SM), so the detection chain overlaps them instead of queueing.
struct NmsBig {
    static constexpr int kThreads = kNmsThreads, kSortMax = kSortSmemMax, kBoxMax = kBoxSmemMax, kWords = kAliveWords, kGroup = 4;
    static constexpr bool kCompact = true;
    static constexpr int kMaxRegs = 64;
};
struct NmsSmall {
    static constexpr int kThreads = 512, kSortMax = 512, kBoxMax = 512, kWords = 16, kGroup = 1;
    static constexpr bool kCompact = false;
    static constexpr int kMaxRegs = 40;       // 512 x 40 registers fit beside the heatmap decode (256 x 160) in one SM register file
};

template <int MODE, typename CFG>
__global__ void __launch_bounds__(CFG::kThreads) __maxnreg__(CFG::kMaxRegs) nms_kernel(const NmsParams prm) {
    constexpr bool RAW = MODE != 0;
    constexpr int kNmsThreads = CFG::kThreads, kSortSmemMax = CFG::kSortMax, kBoxSmemMax = CFG::kBoxMax, kAliveWords = CFG::kWords;
    extern __shared__ __align__(16) unsigned char nms_smem[];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kNmsThreads / 32;

    unsigned long long *skeys = reinterpret_cast<unsigned long long *>(nms_smem);             // [kSortSmemMax]
    float4 *sbox = reinterpret_cast<float4 *>(skeys + kSortSmemMax);                           // [kBoxSmemMax]
    float *sarea = reinterpret_cast<float *>(sbox + kBoxSmemMax);                              // [kBoxSmemMax]
    unsigned *alive = reinterpret_cast<unsigned *>(sarea + kBoxSmemMax);                       // [kAliveWords]
    int *kept_idx = reinterpret_cast<int *>(alive + kAliveWords);                              // [max_det]
    int *sorig = kept_idx + prm.max_det;                                                       // [kBoxSmemMax] (compaction only)

    unsigned long long *gkeys = prm.keys + (size_t)b * prm.cap_pad;
    int raw_count;
#ifdef SPP_NMS_PROF
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long pc = clock64();
#define SPP_PROF_MARK(i) do { const long long now_ = clock64(); pt[i] += now_ - pc; pc = now_; } while (0)
#else
#define SPP_PROF_MARK(i) do { } while (0)
#endif
    if (MODE == 2) {
        __shared__ int s_count;
        if (tid == 0) s_count = 0;
        __syncthreads();
        const Levels &lv = prm.lv;
        const int nc = prm.nc;
        // ---- candidate scan, level by level (plain pointer arithmetic per anchor instead of a level search), U anchors per
        //      thread in flight, class planes read as coalesced lines ----
        constexpr int U = 8;
        for (int j = 0; j < nc; ++j) {
#pragma unroll
            for (int l = 0; l < SPP_MAX_LEVELS; ++l) {
                if (l >= lv.n) continue;
                const int hw = lv.h[l] * lv.w[l], off = lv.off[l];
                const float *plane = lv.cls[l] + (size_t)b * lv.bs_cls[l] + (size_t)j * hw;
                for (int i0 = 0; i0 < hw; i0 += U * kNmsThreads) {
                    float x[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int i = i0 + u * kNmsThreads + tid;
                        x[u] = i < hw ? __ldg(plane + i) : -INFINITY;
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int a = off + i0 + u * kNmsThreads + tid;
                        float sc = 0.f;
                        bool c = false;
                        if (x[u] > prm.logit_lo) {          // -inf for anchors past the end of the level
                            sc = sigmoidf_ref(x[u]);
                            c = sc > prm.conf;
                        }
                        append_candidate(c, make_sort_key(sc, (unsigned)(a * nc + j)), &s_count, gkeys, prm.cap);
                    }
                }
            }
        }
        __syncthreads();
        raw_count = s_count;
        SPP_PROF_MARK(6);  // fused: candidate scan
        // ---- DFL decode of the candidates into this image's box table ----
        const int nd = raw_count < prm.cap ? raw_count : prm.cap;
        const int sub = lane & 15, half = tid >> 4;
        constexpr int NH = kNmsThreads / 16;
        float4 *img_boxes = prm.boxes + (size_t)b * lv.A;
        constexpr int DU = 2;                            // candidates per half-warp in flight (3 / 4 spill at 40 registers and measured slower)
        for (int base = 0; base < nd; base += DU * NH) {
            int anchor[DU];
            bool valid[DU];
#pragma unroll
            for (int u = 0; u < DU; ++u) {
                const int t = base + u * NH + half;
                valid[u] = t < nd;
                const unsigned cand = valid[u] ? (unsigned)(gkeys[t] & 0xffffffffu) : 0u;
                anchor[u] = (int)(cand / (unsigned)nc);
            }
            DflLoad ld[DU];
#pragma unroll
            for (int u = 0; u < DU; ++u) ld[u] = dfl_load(lv, b, anchor[u], sub);
#pragma unroll
            for (int u = 0; u < DU; ++u) dfl_finish(ld[u], anchor[u], valid[u], sub, img_boxes);
        }
        __syncthreads();                                 // keys and boxes written above are read below
        SPP_PROF_MARK(7);  // fused: candidate decode
    } else {
        raw_count = prm.counts[b];
    }
    int n = raw_count < prm.cap ? raw_count : prm.cap;
    unsigned long long *keys;
    if (n <= kNmsThreads) {
        const unsigned long long mine = block_bitonic_reg<kNmsThreads>(tid < n ? gkeys[tid] : ~0ull, reinterpret_cast<unsigned long long *>(sbox));
        skeys[tid] = mine;
        __syncthreads();
        keys = skeys;
    } else {
        int npad = 1;
        while (npad < n) npad <<= 1;
        if (npad <= 4 * kNmsThreads && 4 * kNmsThreads <= kSortSmemMax) {
            // up to 4 096 candidates (crowd scenes): four keys per thread, sorted in registers / shuffles
            unsigned long long k4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) k4[e] = tid * 4 + e < n ? gkeys[tid * 4 + e] : ~0ull;
            block_bitonic_multi<kNmsThreads, 4>(k4, skeys);
            __syncthreads();
#pragma unroll
            for (int e = 0; e < 4; ++e) skeys[tid * 4 + e] = k4[e];
            __syncthreads();
            keys = skeys;
            npad = 0;                                       // sorted: skip the generic path below
        }
        const bool in_smem = npad <= kSortSmemMax;
        if (npad) {
            keys = in_smem ? skeys : gkeys;
            if (in_smem) {
                for (int i = tid; i < npad; i += kNmsThreads) skeys[i] = i < n ? gkeys[i] : ~0ull;
            } else {
                for (int i = n + tid; i < npad; i += kNmsThreads) gkeys[i] = ~0ull;
            }
            __syncthreads();
            bitonic_sort(keys, npad);
        }
    }
    if (n > prm.max_nms) n = prm.max_nms;                                   // util.py:157 [:max_nms]
    SPP_PROF_MARK(0);      // sort

    // un-offset box of sorted candidate j
    auto raw_box = [&](unsigned cand, int &cls) -> float4 {
        const int anchor = cand / prm.nc;
        cls = cand - anchor * prm.nc;
        if (RAW) return prm.boxes[(size_t)b * prm.A + anchor];
        const float *pb = prm.pred + (size_t)b * (4 + prm.nc) * prm.A + anchor;
        return wh2xy(make_float4(__ldg(pb), __ldg(pb + prm.A), __ldg(pb + 2 * (size_t)prm.A), __ldg(pb + 3 * (size_t)prm.A)));
    };
    // class-offset box (util.py:160-161) and its area
    auto load_box = [&](int j, float4 &obox, float &area) {
        int cls;
        const float4 box = raw_box((unsigned)(keys[j] & 0xffffffffu), cls);
        const float off = __fmul_rn((float)cls, prm.max_wh);
        obox = make_float4(__fadd_rn(box.x, off), __fadd_rn(box.y, off), __fadd_rn(box.z, off), __fadd_rn(box.w, off));
        area = __fmul_rn(__fsub_rn(obox.z, obox.x), __fsub_rn(obox.w, obox.y));
    };
    auto get_box = [&](int j, float4 &obox, float &area) {
        if (j < kBoxSmemMax) {
            obox = sbox[j];
            area = sarea[j];
        } else {
            load_box(j, obox, area);
        }
    };

    int nwords = (n + 31) >> 5;
    for (int j = tid; j < n && j < kBoxSmemMax; j += kNmsThreads) {
        float4 ob;
        float ar;
        load_box(j, ob, ar);
        sbox[j] = ob;
        sarea[j] = ar;
    }
    for (int w = tid; w < nwords; w += kNmsThreads) alive[w] = (w * 32 + 32 <= n) ? 0xffffffffu : ((1u << (n - w * 32)) - 1u);
    __syncthreads();

    SPP_PROF_MARK(1);      // staging
    const float thr = prm.iou;
    const int max_det = prm.max_det;
    // Greedy suppression, one GROUP of G 32-candidate bitmask words at a time (three block barriers per 32*G candidates):
    //   (r) every warp computes, for the alive candidates c of the group it owns, the row "which candidates of this group
    //       after c does c suppress" (IoU > thr, one ballot per word) into shared memory — geometry, independent of who
    //       survives;
    //   (a) warp 0 resolves the group in sorted order with bit operations only: the first alive candidate is kept and its
    //       row cleared from the group's alive bits; it publishes the kept bits;
    //   (b) every warp applies those kept boxes to the later words it owns, four boxes in flight (intersections first,
    //       the IEEE division only when some alive lane intersects one of them).
    // A candidate is kept iff no earlier KEPT candidate suppresses it: identical to the serial loop of torchvision's CPU
    // nms.  G = 4 (crowd scenes: ~2 000 candidates, the barrier chain per word was the cost) or 1 (small configuration).
    constexpr int G = CFG::kGroup;
    int nk = 0;
    __shared__ unsigned s_rows[32 * G][G];
    __shared__ float4 s_kb[32 * G];          // boxes kept in the current round (class-offset) and their areas: what (b) applies
    __shared__ float s_ka[32 * G];
    __shared__ int s_kpos[32 * G];
    __shared__ int s_nkr;
    // Compaction (crowd scenes): once more than half of the not-yet-visited candidates are dead, the alive ones are moved
    // (order preserved) to the front of the shared-memory box table and the loop restarts on the dense list: the apply
    // phase and the row phase cost per WORD, dead lanes included.  `sorig` maps a position back to the sorted index the
    // emit phase needs.
    __shared__ int s_alive;                  // alive candidates in words not yet visited
    __shared__ int s_pref[kAliveWords > 256 ? 256 : kAliveWords];
    bool compacted = false;
    if (tid == 0) s_alive = n;
    __syncthreads();
    for (int w0 = 0; w0 < nwords && nk < max_det; w0 += G) {
        unsigned aw[G];
        unsigned any_alive = 0u;
        int group_alive = 0;
#pragma unroll
        for (int q = 0; q < G; ++q) {
            aw[q] = w0 + q < nwords ? alive[w0 + q] : 0u;       // final: every earlier group has been applied
            any_alive |= aw[q];
            group_alive += __popc(aw[q]);
        }
        if (any_alive == 0u) continue;                          // uniform: every thread reads the same words
        // (r)  this lane's own candidate of each word of the group, loaded once per round
        float4 bjq[G];
        float ajq[G];
#pragma unroll
        for (int q = 0; q < G; ++q) {
            bjq[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            ajq[q] = 0.f;
            if ((aw[q] >> lane) & 1u) get_box((w0 + q) * 32 + lane, bjq[q], ajq[q]);
        }
        for (int c = warp; c < 32 * G; c += NW) {
            const int cq = c >> 5, cl = c & 31;
            if (!((aw[cq] >> cl) & 1u)) continue;               // warp-uniform
            float4 bl;
            float al;
            get_box((w0 + cq) * 32 + cl, bl, al);
#pragma unroll
            for (int q = 0; q < G; ++q) {
                if (q < cq) continue;
                const bool mine = (aw[q] >> lane) & 1u;
                const unsigned row = suppress_ballot(bl, al, bjq[q], ajq[q], mine && (q > cq || lane > cl), thr);
                if (lane == 0) s_rows[c][q] = row;
            }
        }
        __syncthreads();
        SPP_PROF_MARK(2);  // (r) + barrier
        if (warp == 0) {   // (a)
            int cnt = nk;
#pragma unroll
            for (int q = 0; q < G; ++q) {
                unsigned word = aw[q];
                while (word && cnt < max_det) {
                    const int l = __ffs(word) - 1;
                    if (lane == 0) {
                        const int pos = (w0 + q) * 32 + l;
                        s_kpos[cnt - nk] = pos;
                        kept_idx[cnt] = compacted ? sorig[pos] : pos;
                    }
                    ++cnt;
                    word &= ~(1u << l) & ~s_rows[q * 32 + l][q];
#pragma unroll
                    for (int q2 = 0; q2 < G; ++q2)
                        if (q2 > q) aw[q2] &= ~s_rows[q * 32 + l][q2];
                }
            }
            __syncwarp();
            for (int i = lane; i < cnt - nk; i += 32) {          // stage this round's kept boxes for the apply phase
                float4 kb;
                float ka;
                get_box(s_kpos[i], kb, ka);
                s_kb[i] = kb;
                s_ka[i] = ka;
            }
            if (lane == 0) {
                s_nkr = cnt - nk;
                s_alive -= group_alive;                          // the whole group is now decided
#pragma unroll
                for (int q = 0; q < G; ++q)
                    if (w0 + q < nwords) alive[w0 + q] = 0u;
            }
        }
        __syncthreads();
        SPP_PROF_MARK(3);  // (a) + barrier
        const int nkr = s_nkr;
        nk += nkr;
        if (nk < max_det) {   // (b)
            for (int wi = w0 + G + warp; wi < nwords; wi += NW) {
                unsigned word = alive[wi];
                if (!word) continue;
                const int j = wi * 32 + lane;
                float4 bj = make_float4(0.f, 0.f, 0.f, 0.f);
                float aj = 0.f;
                const bool alive_lane = (word >> lane) & 1u;
                if (alive_lane) get_box(j, bj, aj);
                bool sup = false;
                for (int i0 = 0; i0 < nkr; i0 += 4) {
                    // four kept boxes in flight: straight-line intersections; the (banded) IoU test only for touching lanes
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (i0 + u < nkr) {                  // warp-uniform
                            const float4 bl = s_kb[i0 + u];
                            const float inter = box_inter(bl, bj);
                            const bool touch = alive_lane && (inter > 0.0f || thr < 0.0f);
                            sup |= ratio_gt(inter, s_ka[i0 + u], aj, thr, touch);
                        }
                    }
                }
                const unsigned gone = word & __ballot_sync(FULL, sup);   // all of them are KEPT boxes: any one suppresses the lane
                word &= ~gone;
                if (lane == 0) {
                    alive[wi] = word;
                    if (CFG::kCompact && gone) atomicSub(&s_alive, __popc(gone));
                }
            }
        }
        SPP_PROF_MARK(4);  // (b) own work
        __syncthreads();
        SPP_PROF_MARK(5);  // (b) waiting for the slowest warp
        if (CFG::kCompact) {
            const int first = w0 + G;                            // first word not yet visited
            const int span = n - first * 32, left = s_alive;     // candidates / alive candidates from there on
            if (nk < max_det && span >= 256 && left * 100 <= span * prm.compact_pct && left <= kBoxSmemMax && nwords - first <= 256) {
                // exclusive prefix of the alive counts of the remaining words (warp 0, up to 8 words per lane)
                if (warp == 0) {
                    int run = 0;
                    for (int base = first; base < nwords; base += 32) {
                        const int w = base + lane;
                        const int c = w < nwords ? __popc(alive[w]) : 0;
                        int inc = c;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int t = __shfl_up_sync(FULL, inc, o);
                            if (lane >= o) inc += t;
                        }
                        if (w < nwords) s_pref[w - first] = run + inc - c;
                        run += __shfl_sync(FULL, inc, 31);
                    }
                }
                __syncthreads();
                // move the alive candidates to the front, a chunk of kNmsThreads old positions at a time (new <= old)
                for (int base = first * 32; base < n; base += kNmsThreads) {
                    const int j = base + tid;
                    float4 ob = make_float4(0.f, 0.f, 0.f, 0.f);
                    float ar = 0.f;
                    int np = -1, og = 0;
                    if (j < n) {
                        const unsigned wd = alive[j >> 5];
                        if ((wd >> (j & 31)) & 1u) {
                            np = s_pref[(j >> 5) - first] + __popc(wd & ((1u << (j & 31)) - 1u));
                            get_box(j, ob, ar);
                            og = compacted ? sorig[j] : j;
                        }
                    }
                    __syncthreads();
                    if (np >= 0) {
                        sbox[np] = ob;
                        sarea[np] = ar;
                        sorig[np] = og;
                    }
                    __syncthreads();
                }
                n = left;
                nwords = (n + 31) >> 5;
                for (int w = tid; w < nwords; w += kNmsThreads) alive[w] = (w * 32 + 32 <= n) ? 0xffffffffu : ((1u << (n - w * 32)) - 1u);
                compacted = true;
                w0 = -G;                                         // restart on the dense list
                __syncthreads();
            }
        }
    }
    __syncthreads();
#ifdef SPP_NMS_PROF
    if (b == 0 && (tid == 0 || tid == 32 * 17)) printf("nms prof tid %d n %d nk %d: scan %lld decode %lld sort %lld staging %lld r %lld a %lld b %lld bwait %lld cycles\n", tid, n, nk, pt[6], pt[7], pt[0], pt[1], pt[2], pt[3], pt[4], pt[5]);
#endif

    // emit the kept rows (un-offset box, score, class) in parallel
    float *dets = prm.out_dets + (size_t)b * max_det * 6;
    int *okeys = prm.out_keys ? prm.out_keys + (size_t)b * max_det : nullptr;
    for (int r = tid; r < max_det; r += kNmsThreads) {
        float *row = dets + (size_t)r * 6;
        if (r < nk) {
            const unsigned long long key = keys[kept_idx[r]];
            const unsigned cand = (unsigned)(key & 0xffffffffu);
            const float score = __uint_as_float(~(unsigned)(key >> 32));
            int cls;
            const float4 box = raw_box(cand, cls);
            row[0] = box.x; row[1] = box.y; row[2] = box.z; row[3] = box.w; row[4] = score; row[5] = (float)cls;
            if (okeys) okeys[r] = (int)cand;
        } else {
            row[0] = row[1] = row[2] = row[3] = row[4] = row[5] = 0.f;
            if (okeys) okeys[r] = -1;
        }
    }
    if (tid == 0) prm.out_count[b] = raw_count > prm.cap ? ~nk : nk;        // overflow: -(nk + 1), distinguishable at nk == 0
}

// cls_levels == nullptr: levels[l] is the concatenated [B, 64+nc, H, W] map; else levels[l] = [B, 64, H, W] box
// conv output and cls_levels[l] = [B, nc, H, W] class conv output.
int fill_levels(Levels &lv, const float *const *levels, const float *const *cls_levels, const int *level_h, const int *level_w,
                const float *strides, int num_levels, int nc) {
    SPP_CHECK_ARG(levels && level_h && level_w && strides, "detection: null level description");
    SPP_CHECK_ARG(num_levels >= 1 && num_levels <= SPP_MAX_LEVELS, "detection: num_levels must be 1..%d", SPP_MAX_LEVELS);
    lv.n = num_levels;
    lv.off[0] = 0;
    for (int l = 0; l < num_levels; ++l) {
        SPP_CHECK_ARG(levels[l] && level_h[l] > 0 && level_w[l] > 0, "detection: bad level %d", l);
        const long long hw = (long long)level_h[l] * level_w[l];
        lv.box[l] = levels[l];
        if (cls_levels) {
            SPP_CHECK_ARG(cls_levels[l], "detection: bad class level %d", l);
            lv.cls[l] = cls_levels[l];
            lv.bs_box[l] = 4 * kDfl * hw;
            lv.bs_cls[l] = (long long)nc * hw;
        } else {
            lv.cls[l] = levels[l] + 4 * kDfl * hw;
            lv.bs_box[l] = lv.bs_cls[l] = (4 * kDfl + nc) * hw;
        }
        lv.h[l] = level_h[l];
        lv.w[l] = level_w[l];
        lv.stride[l] = strides[l];
        lv.off[l + 1] = lv.off[l] + level_h[l] * level_w[l];
    }
    lv.A = lv.off[num_levels];
    return SPP_OK;
}

struct Workspace {
    int *counts;
    unsigned long long *keys;
    float4 *boxes;
    int cap, cap_pad;
    size_t bytes;
};

Workspace carve(void *ws, int batch, int num_anchors, int nc, int max_candidates) {
    Workspace w{};
    long long cap = (long long)num_anchors * nc;
    if (max_candidates > 0 && cap > max_candidates) cap = max_candidates;
    if (cap < 1) cap = 1;
    long long pad = 1;
    while (pad < cap) pad <<= 1;
    w.cap = (int)cap;
    w.cap_pad = (int)pad;
    size_t off = 0;
    unsigned char *p = static_cast<unsigned char *>(ws);
    w.counts = reinterpret_cast<int *>(p + off);
    off += align_up((size_t)batch * sizeof(int), 256);
    w.keys = reinterpret_cast<unsigned long long *>(p + off);
    off += align_up((size_t)batch * pad * sizeof(unsigned long long), 256);
    w.boxes = reinterpret_cast<float4 *>(p + off);
    off += align_up((size_t)batch * num_anchors * sizeof(float4), 256);
    w.bytes = off;
    return w;
}

template <int MODE, typename CFG>
int launch_nms_cfg(const NmsParams &prm, int batch, cudaStream_t st) {
    const size_t smem = (size_t)CFG::kSortMax * 8 + (size_t)CFG::kBoxMax * 20 + (size_t)CFG::kWords * 4 + (size_t)prm.max_det * 4 +
                        (CFG::kCompact ? (size_t)CFG::kBoxMax * 4 : 0);
    // per device and per context: set on every launch (about a microsecond; legal during stream capture)
    SPP_CHECK_CUDA(cudaFuncSetAttribute(nms_kernel<MODE, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    SPP_CHECK_ARG(smem <= 160 * 1024, "nms: max_det %d too large", prm.max_det);
    nms_kernel<MODE, CFG><<<batch, CFG::kThreads, smem, st>>>(prm);
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}

template <int MODE>
int launch_nms(const NmsParams &prm_in, int batch, cudaStream_t st) {
    NmsParams prm = prm_in;
    static const int compact_pct = [] { const char *e = getenv("SPP_NMS_COMPACT_PCT"); const int v = e ? atoi(e) : 50; return v < 0 ? 0 : (v > 100 ? 100 : v); }();
    prm.compact_pct = compact_pct;
    // a caller-bounded candidate list of <= 512 per image: the small-footprint configuration
    if (prm.cap <= NmsSmall::kSortMax && prm.max_det <= 1024)
        return launch_nms_cfg<MODE, NmsSmall>(prm, batch, st);
    return launch_nms_cfg<MODE, NmsBig>(prm, batch, st);
}

int check_nms_args(int batch, int nc, float iou, int max_det, int max_nms, const float *out_dets, const int *out_count,
                   const void *ws) {
    SPP_CHECK_ARG(batch >= 0 && batch <= 8192 && nc >= 1, "nms: need 0 <= batch <= 8192 and nc >= 1 (got %d / %d)", batch, nc);
    if (batch == 0) return SPP_OK;
    SPP_CHECK_ARG(max_det >= 1 && max_det <= 4096 && max_nms >= 1 && max_nms <= kAliveWords * 32,
                  "nms: need 1 <= max_det <= 4096 and 1 <= max_nms <= %d (got %d / %d)", kAliveWords * 32, max_det, max_nms);
    SPP_CHECK_ARG(out_dets && out_count && ws, "nms: null output / workspace");
    (void)iou;
    return SPP_OK;
}

}  // namespace
}  // namespace spp

using namespace spp;

static std::atomic<int> g_det_mode{[] { const char *e = getenv("SPP_DET_FUSED"); return (e && atoi(e) != 0) ? 1 : 0; }()};

extern "C" int spp_decode_nms_mode(int mode) {
    if (mode != 0 && mode != 1) return g_det_mode.load();
    return g_det_mode.exchange(mode);
}

static int head_decode_impl(const float *const *levels, const float *const *cls_levels, const int *level_h, const int *level_w,
                            const float *strides, int num_levels, int batch, int nc, float *out, spp_stream_t stream) {
    if (batch == 0) return SPP_OK;
    SPP_CHECK_ARG(out && batch >= 0 && nc >= 1, "head_decode: bad arguments");
    Levels lv{};
    int rc = fill_levels(lv, levels, cls_levels, level_h, level_w, strides, num_levels, nc);
    if (rc) return rc;
    if (batch == 0) return SPP_OK;
    dim3 grid((lv.A + 255) / 256, batch);
    head_decode_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(lv, nc, out);
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}

extern "C" int spp_head_decode(const float *const *levels, const int *level_h, const int *level_w, const float *strides,
                               int num_levels, int batch, int nc, float *out, spp_stream_t stream) {
    return head_decode_impl(levels, nullptr, level_h, level_w, strides, num_levels, batch, nc, out, stream);
}

extern "C" int spp_head_decode_split(const float *const *box_levels, const float *const *cls_levels, const int *level_h,
                                     const int *level_w, const float *strides, int num_levels, int batch, int nc, float *out,
                                     spp_stream_t stream) {
    SPP_CHECK_ARG(batch == 0 || cls_levels, "head_decode_split: null class levels");
    return head_decode_impl(box_levels, cls_levels, level_h, level_w, strides, num_levels, batch, nc, out, stream);
}

extern "C" size_t spp_nms_workspace_bytes(int batch, int num_anchors, int nc, int max_candidates) {
    if (batch < 0 || num_anchors < 1 || nc < 1) return 0;
    return carve(nullptr, batch, num_anchors, nc, max_candidates).bytes;
}

extern "C" int spp_nms_decoded(const float *pred, int batch, int nc, int num_anchors, float conf_thres, float iou_thres,
                               int max_det, int max_nms, float max_wh, int max_candidates, float *out_dets, int *out_count,
                               int *out_keys, void *workspace, size_t workspace_bytes, spp_stream_t stream) {
    int rc = check_nms_args(batch, nc, iou_thres, max_det, max_nms, out_dets, out_count, workspace);
    if (rc) return rc;
    if (batch == 0) return SPP_OK;
    SPP_CHECK_ARG(pred && num_anchors >= 1, "nms_decoded: bad pred / num_anchors");
    Workspace w = carve(workspace, batch, num_anchors, nc, max_candidates);
    if (workspace_bytes < w.bytes) {
        set_error("nms_decoded: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return SPP_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SPP_CHECK_CUDA(cudaMemsetAsync(w.counts, 0, (size_t)batch * sizeof(int), st));
    dim3 grid((num_anchors + 255) / 256, batch);
    cand_decoded_kernel<<<grid, 256, 0, st>>>(pred, nc, num_anchors, conf_thres, w.cap, w.cap_pad, w.counts, w.keys);
    SPP_CHECK_LAUNCH();
    NmsParams prm{};
    prm.pred = pred; prm.boxes = nullptr; prm.counts = w.counts; prm.keys = w.keys;
    prm.nc = nc; prm.A = num_anchors; prm.cap = w.cap; prm.cap_pad = w.cap_pad;
    prm.iou = iou_thres; prm.max_wh = max_wh; prm.max_det = max_det; prm.max_nms = max_nms;
    prm.out_dets = out_dets; prm.out_count = out_count; prm.out_keys = out_keys;
    return launch_nms<0>(prm, batch, st);
}

static int decode_nms_impl(const float *const *levels, const float *const *cls_levels, const int *level_h, const int *level_w,
                           const float *strides, int num_levels, int batch, int nc, float conf_thres, float iou_thres, int max_det,
                           int max_nms, float max_wh, int max_candidates, float *out_dets, int *out_count, int *out_keys,
                           void *workspace, size_t workspace_bytes, spp_stream_t stream) {
    if (batch == 0) return SPP_OK;
    SPP_CHECK_ARG(nc >= 1, "decode_nms: nc must be >= 1");
    Levels lv{};
    int rc = fill_levels(lv, levels, cls_levels, level_h, level_w, strides, num_levels, nc);
    if (rc) return rc;
    rc = check_nms_args(batch, nc, iou_thres, max_det, max_nms, out_dets, out_count, workspace);
    if (rc) return rc;
    if (batch == 0) return SPP_OK;
    Workspace w = carve(workspace, batch, lv.A, nc, max_candidates);
    if (workspace_bytes < w.bytes) {
        set_error("decode_nms: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return SPP_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // sigmoid(x) > conf can only hold for x > logit(conf); 0.05 of slack covers fp32 rounding
    float logit_lo = -INFINITY;
    if (conf_thres >= 1.0f) logit_lo = INFINITY;
    else if (conf_thres > 0.0f) logit_lo = (float)(std::log((double)conf_thres / (1.0 - (double)conf_thres)) - 0.05);
    NmsParams prm{};
    prm.pred = nullptr; prm.boxes = w.boxes; prm.counts = w.counts; prm.keys = w.keys;
    prm.nc = nc; prm.A = lv.A; prm.cap = w.cap; prm.cap_pad = w.cap_pad;
    prm.iou = iou_thres; prm.max_wh = max_wh; prm.max_det = max_det; prm.max_nms = max_nms;
    prm.out_dets = out_dets; prm.out_count = out_count; prm.out_keys = out_keys;
    // Default: three launches per head (candidate scan, candidate decode, sort + NMS).  spp_decode_nms_mode(1) / SPP_DET_FUSED=1
    // selects ONE fused kernel per head (a CTA per image does all of it): same results bit for bit, one launch instead of
    // three + a memset; measured on B200 at cfg2 it is slower alone (57 vs 45 us per head: one CTA serialises what two
    // full-machine launches do in parallel) and equal inside the step, so it is the option, not the default.
    const bool split3 = g_det_mode.load(std::memory_order_relaxed) == 0;
    if (!split3) {
        prm.lv = lv; prm.conf = conf_thres; prm.logit_lo = logit_lo;
        return launch_nms<2>(prm, batch, st);
    }
    SPP_CHECK_CUDA(cudaMemsetAsync(w.counts, 0, (size_t)batch * sizeof(int), st));
    dim3 grid((lv.A + 256 * kScanPerThread - 1) / (256 * kScanPerThread), batch);
    cand_scan_kernel<<<grid, 256, 0, st>>>(lv, nc, conf_thres, logit_lo, w.cap, w.cap_pad, w.counts, w.keys);
    SPP_CHECK_LAUNCH();
    {
        int sms = sm_count();
        if (sms <= 0) return SPP_ERR_CUDA;
        cand_decode_kernel<<<sms * 8, 256, (size_t)(batch + 1) * sizeof(int), st>>>(lv, nc, batch, w.cap, w.cap_pad, w.counts, w.keys, w.boxes);
        SPP_CHECK_LAUNCH();
    }
    return launch_nms<1>(prm, batch, st);
}

extern "C" int spp_decode_nms(const float *const *levels, const int *level_h, const int *level_w, const float *strides,
                              int num_levels, int batch, int nc, float conf_thres, float iou_thres, int max_det,
                              int max_nms, float max_wh, int max_candidates, float *out_dets, int *out_count,
                              int *out_keys, void *workspace, size_t workspace_bytes, spp_stream_t stream) {
    return decode_nms_impl(levels, nullptr, level_h, level_w, strides, num_levels, batch, nc, conf_thres, iou_thres, max_det, max_nms,
                           max_wh, max_candidates, out_dets, out_count, out_keys, workspace, workspace_bytes, stream);
}

extern "C" int spp_decode_nms_split(const float *const *box_levels, const float *const *cls_levels, const int *level_h,
                                    const int *level_w, const float *strides, int num_levels, int batch, int nc, float conf_thres,
                                    float iou_thres, int max_det, int max_nms, float max_wh, int max_candidates, float *out_dets,
                                    int *out_count, int *out_keys, void *workspace, size_t workspace_bytes, spp_stream_t stream) {
    SPP_CHECK_ARG(batch == 0 || cls_levels, "decode_nms_split: null class levels");
    return decode_nms_impl(box_levels, cls_levels, level_h, level_w, strides, num_levels, batch, nc, conf_thres, iou_thres, max_det,
                           max_nms, max_wh, max_candidates, out_dets, out_count, out_keys, workspace, workspace_bytes, stream);
}
