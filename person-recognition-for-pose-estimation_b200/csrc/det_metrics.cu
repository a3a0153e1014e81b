// Detection evaluation on the device: the true-positive matrix and the AP / precision / recall summary
// of the reference's test loop.
//
// Replaces
//   compute_metric   training/yolopt/util.py:99-120    (spp_det_match_targets: one CTA per image)
//   compute_ap       training/yolopt/util.py:225-300   (spp_det_average_precision; `smooth` :172-177)
// as driven by training/yolopt/main.py:199-234.  Plotting is not reproduced.
//
// compute_metric is integer / fp32 work and is reproduced bit for bit (IoU with round-to-nearest intrinsics in the
// reference's operation order).  compute_ap is float64 numpy: the same formulas in fp64, numpy's own summation orders
// where they are defined (pairwise add.reduce, sequential axis-0 reduce), numpy.interp's exact branch structure; the one
// order numpy does not define (the BLAS dot inside numpy.convolve, used only to pick the max-F1 operating point) is summed
// left to right.  Latency-bound bookkeeping: nothing here is on the per-frame hot path.
#include "spp_common.cuh"

#include <cmath>

namespace spp {
namespace {

constexpr int kMaxIou = 16;
constexpr int kPx = 1000;       // util.py:245 px = linspace(0, 1, 1000)
constexpr int kApPts = 101;     // util.py:273 101-point interpolation (COCO)

struct IouV {
    float v[kMaxIou];
    int n;
};

// ------------------------------------------------------------------------------------------------ compute_metric
// One CTA per image.  A detection's best label is the class-matching label of highest IoU (independent of the threshold);
// per threshold the detection is a candidate when that IoU reaches it, and a label goes to the LOWEST-index candidate
// detection (util.py:115-117: sort by IoU, unique per detection, then unique per label on an array that the first unique
// left ordered by detection index).
__global__ void __launch_bounds__(256) det_match_targets_kernel(const float *__restrict__ dets, const int *__restrict__ dcount, int dcap,
                                                                const float *__restrict__ targets, const int *__restrict__ tcount,
                                                                int tcap, const IouV iv, unsigned char *__restrict__ correct) {
    extern __shared__ __align__(16) unsigned char dm_smem[];
    float *tg = reinterpret_cast<float *>(dm_smem);              // [tcap, 5]
    float *biou = tg + (size_t)tcap * 5;                         // [dcap]
    int *blab = reinterpret_cast<int *>(biou + dcap);            // [dcap]
    int *winner = blab + dcap;                                   // [tcap]
    const int b = blockIdx.x, tid = threadIdx.x;
    int n = dcount[b];
    n = n < 0 ? ~n : n;                                          // ~kept flags an upstream candidate overflow
    n = n < dcap ? n : dcap;
    int m = tcount[b];
    m = m < 0 ? 0 : (m < tcap ? m : tcap);
    const float *d0 = dets + (size_t)b * dcap * 6;
    for (int i = tid; i < m * 5; i += blockDim.x) tg[i] = targets[(size_t)b * tcap * 5 + i];
    __syncthreads();
    for (int d = tid; d < dcap; d += blockDim.x) {
        float best = -INFINITY;
        int bl = -1;
        if (d < n) {
            const float bx1 = d0[d * 6], by1 = d0[d * 6 + 1], bx2 = d0[d * 6 + 2], by2 = d0[d * 6 + 3], bc = d0[d * 6 + 5];
            const float area_b = __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(by2, by1));
            for (int l = 0; l < m; ++l) {
                const float *t = tg + l * 5;
                if (t[0] != bc) continue;
                const float w = fmaxf(__fsub_rn(fminf(t[3], bx2), fmaxf(t[1], bx1)), 0.0f);
                const float h = fmaxf(__fsub_rn(fminf(t[4], by2), fmaxf(t[2], by1)), 0.0f);
                const float inter = __fmul_rn(w, h);
                const float area_a = __fmul_rn(__fsub_rn(t[3], t[1]), __fsub_rn(t[4], t[2]));
                const float iou = __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(area_a, area_b), inter), 1e-7f));
                if (iou > best) { best = iou; bl = l; }
            }
        }
        biou[d] = best;
        blab[d] = bl;
    }
    for (int i = 0; i < iv.n; ++i) {
        for (int l = tid; l < m; l += blockDim.x) winner[l] = 0x7fffffff;
        __syncthreads();
        const float thr = iv.v[i];
        for (int d = tid; d < n; d += blockDim.x)
            if (blab[d] >= 0 && biou[d] >= thr) atomicMin(&winner[blab[d]], d);
        __syncthreads();
        for (int d = tid; d < dcap; d += blockDim.x)
            correct[((size_t)b * dcap + d) * iv.n + i] = (d < n && blab[d] >= 0 && biou[d] >= thr && winner[blab[d]] == d) ? 1 : 0;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ sort by confidence
// keys ascending == confidence descending (confidences are >= 0, so their bit patterns order like the values), row
// index ascending on ties.
__global__ void ap_make_keys_kernel(const float *__restrict__ conf, int n, int n_pad, unsigned long long *keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    keys[i] = i < n ? (((unsigned long long)(~__float_as_uint(conf[i]))) << 32) | (unsigned)i : ~0ull;
}

constexpr int kSortBlock = 2048;
// Bitonic network restricted to one 2048-key block in shared memory: `full` = the complete sort of the block
// (k = 2 .. 2048); otherwise the last 11 steps (j = 1024 .. 1) of merge size k > 2048.
__global__ void __launch_bounds__(1024) ap_bitonic_local_kernel(unsigned long long *keys, int k_merge, bool full) {
    __shared__ unsigned long long s[kSortBlock];
    const int base = blockIdx.x * kSortBlock, tid = threadIdx.x;
    s[tid] = keys[base + tid];
    s[tid + 1024] = keys[base + tid + 1024];
    __syncthreads();
    for (int k = full ? 2 : k_merge; k <= (full ? kSortBlock : k_merge); k <<= 1) {
        for (int j = (k >> 1) < 1024 ? (k >> 1) : 1024; j > 0; j >>= 1) {
            const int i = ((tid & ~(j - 1)) << 1) | (tid & (j - 1));
            const int p = i | j;
            const bool up = ((base + i) & k) == 0;
            const unsigned long long x = s[i], y = s[p];
            if ((x > y) == up) { s[i] = y; s[p] = x; }
            __syncthreads();
        }
    }
    keys[base + tid] = s[tid];
    keys[base + tid + 1024] = s[tid + 1024];
}
__global__ void ap_bitonic_global_kernel(unsigned long long *keys, int n_pad, int j, int k) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (n_pad >> 1)) return;
    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const int p = i | j;
    const bool up = (i & k) == 0;
    const unsigned long long x = keys[i], y = keys[p];
    if ((x > y) == up) { keys[i] = y; keys[p] = x; }
}

// ------------------------------------------------------------------------------------------------ class table
struct ApTable {          // device-side bookkeeping, [nc_max] each
    int *cls, *nt, *no, *off, *num;
};

// numpy.unique(target, return_counts=True) over small non-negative integer class ids + detections per class.
__global__ void __launch_bounds__(1024) ap_class_table_kernel(const float *__restrict__ target_cls, int nt_total, const float *__restrict__ pred_cls,
                                                              int n, int nc_max, ApTable tb) {
    extern __shared__ int ct_smem[];
    int *lab = ct_smem, *det = ct_smem + nc_max;
    for (int c = threadIdx.x; c < 2 * nc_max; c += blockDim.x) ct_smem[c] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < nt_total; i += blockDim.x) {
        const float c = target_cls[i];
        if (c >= 0.0f && c < (float)nc_max && c == floorf(c)) atomicAdd(&lab[(int)c], 1);
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float c = pred_cls[i];
        if (c >= 0.0f && c < (float)nc_max && c == floorf(c)) atomicAdd(&det[(int)c], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int nu = 0, off = 0;
        for (int c = 0; c < nc_max; ++c)
            if (lab[c] > 0) {
                tb.cls[nu] = c; tb.nt[nu] = lab[c]; tb.no[nu] = det[c]; tb.off[nu] = off;
                off += det[c];
                ++nu;
            }
        for (int c = nu; c < nc_max; ++c) { tb.cls[c] = -1; tb.nt[c] = 0; tb.no[c] = 0; tb.off[c] = off; }
        *tb.num = nu;
    }
}

// ------------------------------------------------------------------------------------------------ per-class curves
__device__ __forceinline__ double px_at(int q) { return q == kPx - 1 ? 1.0 : (double)q * (1.0 / 999.0); }        // numpy.linspace(0, 1, 1000)
__device__ __forceinline__ double x101_at(int q) { return q == kApPts - 1 ? 1.0 : (double)q * (1.0 / 100.0); }   // numpy.linspace(0, 1, 101)

// numpy.interp (compiled_base.c arr_interp): j = last index with xp[j] <= x; left / right / exact-hit branches as numpy.
template <typename XP, typename FP>
__device__ __forceinline__ double np_interp(double x, int len, XP xp, FP fp, double left) {
    if (x > xp(len - 1)) return fp(len - 1);
    if (x < xp(0)) return left;
    int lo = 0, hi = len;                       // first index with xp > x
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (x >= xp(mid)) lo = mid + 1; else hi = mid;
    }
    const int j = lo - 1;
    if (j == len - 1) return fp(j);
    if (xp(j) == x) return fp(j);
    const double slope = __ddiv_rn(__dsub_rn(fp(j + 1), fp(j)), __dsub_rn(xp(j + 1), xp(j)));
    double r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xp(j))), fp(j));
    if (isnan(r)) {
        r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xp(j + 1))), fp(j + 1));
        if (isnan(r) && fp(j) == fp(j + 1)) r = fp(j);
    }
    return r;
}

// numpy's pairwise summation (loops_utils.h.src pairwise_sum_DOUBLE) over a strided array
__device__ double np_pairwise_sum(const double *a, int n, int stride) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[(size_t)i * stride]);
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = a[(size_t)k * stride];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], a[(size_t)(i + k) * stride]);
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __dadd_rn(res, a[(size_t)i * stride]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(np_pairwise_sum(a, n2, stride), np_pairwise_sum(a + (size_t)n2 * stride, n - n2, stride));
}

struct ApParams {
    const unsigned char *tp;      // [n, T]
    const float *conf, *pred_cls;
    int n, T, nc_max;
    double eps;
    const unsigned long long *keys;
    ApTable tb;
    int *order;                   // [n]  rows of each class in confidence order, class segments back to back
    double *xconf;                // [n]  -conf of those rows (interp abscissa)
    double *rec, *prec;           // [T, n]
    double *p, *r;                // [nc_max, 1000]
    double *ap;                   // [nc_max, T]
};

constexpr int kApThreads = 256;

// block-wide exclusive prefix sum of one int per thread (blockDim = 256); returns the prefix, writes the block total
__device__ __forceinline__ int block_exclusive_sum(int v, int *warp_tot, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();                     // warp_tot is reused by back-to-back calls
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int before = 0;
    total = 0;
#pragma unroll
    for (int w = 0; w < kApThreads / 32; ++w) {
        const int t = warp_tot[w];
        if (w < warp) before += t;
        total += t;
    }
    return before + inc - v;
}

__global__ void __launch_bounds__(kApThreads) ap_class_curves_kernel(const ApParams prm) {
    __shared__ int warp_tot[kApThreads / 32];
    __shared__ double sd[kApThreads];
    __shared__ double y101[kApPts];
    const int ci = blockIdx.x, tid = threadIdx.x;
    if (ci >= *prm.tb.num) return;
    const int c = prm.tb.cls[ci], nl = prm.tb.nt[ci], no = prm.tb.no[ci], off = prm.tb.off[ci];
    if (no == 0 || nl == 0) return;              // util.py:252-253: the class keeps its zero rows
    const int n = prm.n, T = prm.T;
    int *order = prm.order + off;
    double *xc = prm.xconf + off;

    // ---- rows of this class, in confidence order (ordered compaction of the globally sorted list) ----
    int running = 0;
    for (int base = 0; base < n; base += kApThreads) {
        const int i = base + tid;
        int idx = 0, flag = 0;
        if (i < n) {
            idx = (int)(prm.keys[i] & 0xffffffffull);
            flag = prm.pred_cls[idx] == (float)c;
        }
        int total;
        const int pos = running + block_exclusive_sum(flag, warp_tot, total);
        if (flag) {
            order[pos] = idx;
            xc[pos] = -(double)prm.conf[idx];
        }
        running += total;
    }
    __syncthreads();

    // ---- cumulative TP / FP -> recall and precision curves, one IoU threshold at a time (util.py:256-265) ----
    for (int j = 0; j < T; ++j) {
        double *rec = prm.rec + (size_t)j * n + off, *prec = prm.prec + (size_t)j * n + off;
        int carry = 0;
        for (int base = 0; base < no; base += kApThreads) {
            const int k = base + tid;
            const int v = k < no ? (int)prm.tp[(size_t)order[k] * T + j] : 0;
            int total;
            const int tpc = carry + block_exclusive_sum(v, warp_tot, total) + v;
            if (k < no) {
                rec[k] = __ddiv_rn((double)tpc, __dadd_rn((double)nl, prm.eps));
                prec[k] = __ddiv_rn((double)tpc, (double)(k + 1));          // tpc + fpc == k + 1
            }
            carry += total;
        }
    }
    __syncthreads();

    // ---- precision / recall at 1000 confidence levels, threshold 0 (util.py:262, 266) ----
    {
        const double *rec0 = prm.rec + off, *prec0 = prm.prec + off;
        auto xp = [&](int k) { return xc[k]; };
        for (int q = tid; q < kPx; q += kApThreads) {
            const double x = -px_at(q);
            prm.r[(size_t)ci * kPx + q] = np_interp(x, no, xp, [&](int k) { return rec0[k]; }, 0.0);
            prm.p[(size_t)ci * kPx + q] = np_interp(x, no, xp, [&](int k) { return prec0[k]; }, 1.0);
        }
    }
    __syncthreads();

    // ---- AP per threshold: precision envelope, 101-point interpolation, trapezoid (util.py:269-275) ----
    for (int j = 0; j < T; ++j) {
        double *rec = prm.rec + (size_t)j * n + off, *prec = prm.prec + (size_t)j * n + off;
        // envelope = reverse running maximum, chunk by chunk from the end (suffix max inside the chunk, then the carry)
        double carry = 0.0;                                     // the appended m_pre[-1] = 0
        for (int base = ((no - 1) / kApThreads) * kApThreads; base >= 0; base -= kApThreads) {
            const int k = base + tid;
            sd[tid] = k < no ? prec[k] : 0.0;
            __syncthreads();
            for (int o = 1; o < kApThreads; o <<= 1) {
                const double other = tid + o < kApThreads ? sd[tid + o] : 0.0;
                __syncthreads();
                sd[tid] = fmax(sd[tid], other);
                __syncthreads();
            }
            const double chunk_max = sd[0];
            if (k < no) prec[k] = fmax(sd[tid], carry);
            __syncthreads();
            carry = fmax(carry, chunk_max);
        }
        __syncthreads();
        // m_rec = [0, rec..., 1], m_pre = [1, envelope..., 0]
        const int len = no + 2;
        auto xr = [&](int k) { return k == 0 ? 0.0 : (k == len - 1 ? 1.0 : rec[k - 1]); };
        auto yp = [&](int k) { return k == 0 ? 1.0 : (k == len - 1 ? 0.0 : prec[k - 1]); };
        for (int q = tid; q < kApPts; q += kApThreads) y101[q] = np_interp(x101_at(q), len, xr, yp, 0.0);
        __syncthreads();
        if (tid == 0) {
            double term[kApPts - 1];
            for (int k = 0; k < kApPts - 1; ++k)
                term[k] = __ddiv_rn(__dmul_rn(__dsub_rn(x101_at(k + 1), x101_at(k)), __dadd_rn(y101[k + 1], y101[k])), 2.0);
            prm.ap[(size_t)ci * T + j] = np_pairwise_sum(term, kApPts - 1, 1);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ summary
// f1, the max-F1 operating point (util.py:278-293) and the means.  One CTA.
__global__ void __launch_bounds__(1024) ap_summary_kernel(const ApParams prm, double *f1mean, double *out_class_stats, double *out_summary) {
    __shared__ double best_v[1024];
    __shared__ int best_i[1024];
    __shared__ int s_idx;
    const int nu = *prm.tb.num, tid = threadIdx.x, T = prm.T;
    // f1.mean(0): rows added one after the other (numpy's axis-0 reduce), then / nc
    for (int q = tid; q < kPx; q += blockDim.x) {
        double acc = 0.0;
        for (int c = 0; c < nu; ++c) {
            const double p = prm.p[(size_t)c * kPx + q], r = prm.r[(size_t)c * kPx + q];
            const double f1 = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, p), r), __dadd_rn(__dadd_rn(p, r), prm.eps));
            acc = c == 0 ? f1 : __dadd_rn(acc, f1);
        }
        f1mean[q] = nu > 0 ? __ddiv_rn(acc, (double)nu) : 0.0;
    }
    __syncthreads();
    // smooth(y, 0.1): nf = 101 taps of 1/101 over the edge-padded curve (util.py:172-177), then argmax (first maximum)
    const int nf = 101, half = nf / 2;
    const double w = 1.0 / (double)nf;
    double bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int q = tid; q < kPx; q += blockDim.x) {
        double acc = 0.0;
        for (int k = 0; k < nf; ++k) {
            int src = q + k - half;
            src = src < 0 ? 0 : (src > kPx - 1 ? kPx - 1 : src);
            acc = __dadd_rn(acc, __dmul_rn(f1mean[src], w));
        }
        if (acc > bv) { bv = acc; bi = q; }
    }
    best_v[tid] = bv;
    best_i[tid] = bi;
    __syncthreads();
    if (tid == 0) {
        double v = -INFINITY;
        int i = 0;
        for (int t = 0; t < (int)blockDim.x; ++t)
            if (best_v[t] > v || (best_v[t] == v && best_i[t] < i)) { v = best_v[t]; i = best_i[t]; }
        s_idx = nu > 0 ? i : 0;
    }
    __syncthreads();
    const int idx = s_idx;
    // per class at that confidence level: p, r, tp = round(r * nt), fp = round(tp / (p + eps) - tp); AP@0.5 and mean AP
    double *pi = f1mean + kPx, *ri = pi + prm.nc_max, *ap50 = ri + prm.nc_max, *apm = ap50 + prm.nc_max;      // scratch after the curve
    for (int c = tid; c < nu; c += blockDim.x) {
        const double p = prm.p[(size_t)c * kPx + idx], r = prm.r[(size_t)c * kPx + idx];
        pi[c] = p;
        ri[c] = r;
        const double tpn = rint(__dmul_rn(r, (double)prm.tb.nt[c]));
        const double fpn = rint(__dsub_rn(__ddiv_rn(tpn, __dadd_rn(p, prm.eps)), tpn));
        out_class_stats[c * 4 + 0] = tpn;
        out_class_stats[c * 4 + 1] = fpn;
        out_class_stats[c * 4 + 2] = p;
        out_class_stats[c * 4 + 3] = r;
        ap50[c] = prm.ap[(size_t)c * T];
        apm[c] = __ddiv_rn(np_pairwise_sum(prm.ap + (size_t)c * T, T, 1), (double)T);
    }
    __syncthreads();
    if (tid == 0) {
        const double dn = (double)(nu > 0 ? nu : 1);
        out_summary[0] = __ddiv_rn(np_pairwise_sum(pi, nu, 1), dn);       // m_pre
        out_summary[1] = __ddiv_rn(np_pairwise_sum(ri, nu, 1), dn);       // m_rec
        out_summary[2] = __ddiv_rn(np_pairwise_sum(ap50, nu, 1), dn);     // map50
        out_summary[3] = __ddiv_rn(np_pairwise_sum(apm, nu, 1), dn);      // mean_ap
        out_summary[4] = (double)idx;
        out_summary[5] = (double)nu;
    }
}

struct ApPlan {
    int n_pad;
    size_t off_keys, off_order, off_xconf, off_rec, off_prec, off_p, off_r, off_tab, off_f1, bytes;
};

ApPlan plan_ap(int n, int T, int nc_max) {
    ApPlan p{};
    int np2 = kSortBlock;
    while (np2 < n) np2 <<= 1;
    p.n_pad = np2;
    const size_t nn = n > 0 ? (size_t)n : 1;
    size_t off = 0;
    p.off_keys = off;  off += align_up((size_t)np2 * 8, 256);
    p.off_order = off; off += align_up(nn * 4, 256);
    p.off_xconf = off; off += align_up(nn * 8, 256);
    p.off_rec = off;   off += align_up(nn * T * 8, 256);
    p.off_prec = off;  off += align_up(nn * T * 8, 256);
    p.off_p = off;     off += align_up((size_t)nc_max * kPx * 8, 256);
    p.off_r = off;     off += align_up((size_t)nc_max * kPx * 8, 256);
    p.off_tab = off;   off += align_up((size_t)(4 * nc_max + 1) * 4, 256);
    p.off_f1 = off;    off += align_up((size_t)(kPx + 4 * nc_max) * 8, 256);
    p.bytes = off;
    return p;
}

}  // namespace
}  // namespace spp

using namespace spp;

extern "C" int spp_det_match_targets(const float *dets, const int *det_count, int det_cap, const float *targets, const int *target_count,
                                     int target_cap, const float *iou_v, int n_iou, int batch, unsigned char *correct, spp_stream_t stream) {
    if (batch == 0) return SPP_OK;
    SPP_CHECK_ARG(dets && det_count && targets && target_count && iou_v && correct, "det_match_targets: null pointer");
    SPP_CHECK_ARG(batch > 0 && det_cap >= 1 && target_cap >= 1 && n_iou >= 1 && n_iou <= kMaxIou,
                  "det_match_targets: need det_cap, target_cap >= 1 and 1 <= n_iou <= %d", kMaxIou);
    IouV iv{};
    iv.n = n_iou;
    for (int i = 0; i < n_iou; ++i) iv.v[i] = iou_v[i];
    const size_t smem = (size_t)target_cap * 5 * 4 + (size_t)det_cap * 8 + (size_t)target_cap * 4;
    SPP_CHECK_ARG(smem <= 200 * 1024, "det_match_targets: det_cap / target_cap too large for one CTA per image");
    SPP_CHECK_CUDA(cudaFuncSetAttribute(det_match_targets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    det_match_targets_kernel<<<batch, 256, smem, static_cast<cudaStream_t>(stream)>>>(dets, det_count, det_cap, targets, target_count,
                                                                                    target_cap, iv, correct);
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}

extern "C" size_t spp_det_ap_workspace_bytes(int n, int n_iou, int nc_max) {
    if (n < 0 || n_iou < 1 || n_iou > kMaxIou || nc_max < 1 || nc_max > 4096) return 0;
    return plan_ap(n, n_iou, nc_max).bytes;
}

extern "C" int spp_det_average_precision(const unsigned char *tp, const float *conf, const float *pred_cls, int n, const float *target_cls,
                                         int nt, int n_iou, int nc_max, double eps, int *out_classes, int *out_num_classes, double *out_ap,
                                         double *out_class_stats, double *out_summary, void *workspace, size_t workspace_bytes,
                                         spp_stream_t stream) {
    SPP_CHECK_ARG(n >= 0 && nt >= 0 && n_iou >= 1 && n_iou <= kMaxIou && nc_max >= 1 && nc_max <= 4096, "det_average_precision: bad sizes");
    SPP_CHECK_ARG(out_classes && out_num_classes && out_ap && out_class_stats && out_summary && workspace, "det_average_precision: null output");
    SPP_CHECK_ARG(n == 0 || (tp && conf && pred_cls), "det_average_precision: null detections");
    SPP_CHECK_ARG(nt == 0 || target_cls, "det_average_precision: null labels");
    SPP_CHECK_ARG(n < (1 << 30), "det_average_precision: too many detections");
    const ApPlan pl = plan_ap(n, n_iou, nc_max);
    if (workspace_bytes < pl.bytes) {
        set_error("det_average_precision: workspace %zu < required %zu bytes", workspace_bytes, pl.bytes);
        return SPP_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    ApParams prm{};
    prm.tp = tp; prm.conf = conf; prm.pred_cls = pred_cls; prm.n = n; prm.T = n_iou; prm.nc_max = nc_max; prm.eps = eps;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(ws + pl.off_keys);
    prm.keys = keys;
    prm.order = reinterpret_cast<int *>(ws + pl.off_order);
    prm.xconf = reinterpret_cast<double *>(ws + pl.off_xconf);
    prm.rec = reinterpret_cast<double *>(ws + pl.off_rec);
    prm.prec = reinterpret_cast<double *>(ws + pl.off_prec);
    prm.p = reinterpret_cast<double *>(ws + pl.off_p);
    prm.r = reinterpret_cast<double *>(ws + pl.off_r);
    prm.ap = out_ap;
    int *tab = reinterpret_cast<int *>(ws + pl.off_tab);
    prm.tb = ApTable{out_classes, tab, tab + nc_max, tab + 2 * nc_max, out_num_classes};
    double *f1 = reinterpret_cast<double *>(ws + pl.off_f1);

    // numpy.zeros for p, r, ap (util.py:242-244)
    SPP_CHECK_CUDA(cudaMemsetAsync(prm.p, 0, (size_t)nc_max * kPx * 8, st));
    SPP_CHECK_CUDA(cudaMemsetAsync(prm.r, 0, (size_t)nc_max * kPx * 8, st));
    SPP_CHECK_CUDA(cudaMemsetAsync(out_ap, 0, (size_t)nc_max * n_iou * 8, st));
    SPP_CHECK_CUDA(cudaMemsetAsync(out_class_stats, 0, (size_t)nc_max * 4 * 8, st));
    if (n > 0) {
        ap_make_keys_kernel<<<(pl.n_pad + 255) / 256, 256, 0, st>>>(conf, n, pl.n_pad, keys);
        SPP_CHECK_LAUNCH();
        ap_bitonic_local_kernel<<<pl.n_pad / kSortBlock, 1024, 0, st>>>(keys, 0, true);
        SPP_CHECK_LAUNCH();
        for (int k = 2 * kSortBlock; k <= pl.n_pad; k <<= 1) {
            for (int j = k >> 1; j >= kSortBlock; j >>= 1) {
                ap_bitonic_global_kernel<<<(pl.n_pad / 2 + 255) / 256, 256, 0, st>>>(keys, pl.n_pad, j, k);
                SPP_CHECK_LAUNCH();
            }
            ap_bitonic_local_kernel<<<pl.n_pad / kSortBlock, 1024, 0, st>>>(keys, k, false);
            SPP_CHECK_LAUNCH();
        }
    }
    ap_class_table_kernel<<<1, 1024, (size_t)2 * nc_max * sizeof(int), st>>>(target_cls, nt, pred_cls, n, nc_max, prm.tb);
    SPP_CHECK_LAUNCH();
    if (n > 0) {
        ap_class_curves_kernel<<<nc_max, kApThreads, 0, st>>>(prm);
        SPP_CHECK_LAUNCH();
    }
    ap_summary_kernel<<<1, 1024, 0, st>>>(prm, f1, out_class_stats, out_summary);
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}
