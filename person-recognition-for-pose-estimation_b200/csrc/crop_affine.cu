// Person-box affine crop to the ViTPose input (bilinear warp + normalisation).
//
// Replaces:
//   variant 0  HF VitPoseImageProcessor.preprocess — image_processing_vitpose.py:68-109
//              (box_to_center_and_scale), :112-146 (get_warp_matrix, UDP), :149-172 (scipy order-1
//              warp, zero outside the frame), :386-448; normalisation per
//              image_processing_backends.py:292-331
//   variant 1  training/lightning/pose_estimation/datamodule_v2.py:119-129, 213-226 (gluoncv
//              get_affine_transform + warpAffine; exact bilinear here, no 1/32-px quantisation)
//
// The warp has no rotation, so source coordinates are separable: x_s depends only on the output
// column, y_s only on the output row.  Both coordinate tables are computed in fp64 (the oracle inverts the
// fp32 matrix in fp64 and samples in fp64; an fp32 coordinate at x ~ 1000 px would already be off by 6e-5 px)
// and kept as (index, weight) pairs in shared memory.  A producer warp streams the needed source rows into
// shared-memory band buffers with bulk TMA (full / empty mbarriers, no block barrier in the main loop); six
// consumer warps sample them in fp32.  HBM-bound: 4 B (2 B with bf16 output) written per output element + the
// source ROI read once.
//
// Two implementations of the same work item (crop, channel, slab of output rows), bit-identical results:
//   crop_affine_kernel    one CTA per item, tables built by the CTA itself (no workspace; spp_crop_affine[_u8]);
//   crop_plan_kernel + crop_stream_kernel    tables planned once per crop into a workspace, resident CTAs pulling
//                         items from a ticket counter and fetching their tables by bulk TMA (spp_crop_affine*_ws /
//                         _run / _ex with a workspace; chosen up to ~8 waves of items, see launch_crop).
//   crop_affine_direct_kernel    staging-free fallback (unaligned frame rows, windows too wide for a band buffer).
#include "spp_common.cuh"

#include <cuda_bf16.h>

#include <atomic>
#include <climits>
#include <cstdlib>
#include <type_traits>

namespace spp {
namespace {

struct CropParams {
    const void *frames;         // [num_frames, 3, fh, fw] fp32 or uint8
    const float *boxes;
    const int *frame_idx;
    void *out;                  // [P, 3, oh, ow] fp32 or bf16 (template parameter O of the kernels)
    int num_frames, fh, fw, P, oh, ow, variant;
    int stage_bytes;            // size of one shared-memory band buffer
    int stages;                 // band buffers per CTA (a power of two)
    int stages_log2;
    int split;                  // persistent kernels: row slabs per (crop, channel)
    unsigned char *ws;          // persistent kernels: workspace (ticket counter, planned tables and item descriptors)
#ifdef SPP_CROP_TRACE
    unsigned long long *trace;  // debug build (make EXTRA=-DSPP_CROP_TRACE): event log of one CTA, see tools/crop_trace.py
#endif
    int ncc, rg;                // staged kernel: column chunks (of 32 C columns) x row groups = warps per CTA
    float mean[3], stdv[3];
};

// Source coordinate map for one crop: src = a * dst + b (per axis), fp64.
struct AxisMap {
    double ax, bx, ay, by;
};

__device__ __forceinline__ AxisMap crop_axis_map(const float4 box, int ow, int oh, int variant) {
    AxisMap m;
    double w = box.z, h = box.w;
    const double aspect = (double)ow / (double)oh;
    if (variant == SPP_CROP_HF_UDP) {
        // box_to_center_and_scale (Python-float arithmetic, fp32 storage)
        const float cx = (float)((double)box.x + w * 0.5), cy = (float)((double)box.y + h * 0.5);
        if (w > aspect * h) h = w * 1.0 / aspect;
        else if (w < aspect * h) w = h * aspect;
        const float sx = __fmul_rn((float)(w / 200.0), 1.25f), sy = __fmul_rn((float)(h / 200.0), 1.25f);
        // get_warp_matrix(0, center*2, (W-1, H-1), scale*200): fp32 inputs, fp64 ratio, fp32 storage
        const float in_x = __fmul_rn(cx, 2.0f), in_y = __fmul_rn(cy, 2.0f);
        const float tg_x = __fmul_rn(sx, 200.0f), tg_y = __fmul_rn(sy, 200.0f);
        const double rx = (double)(ow - 1) / (double)tg_x, ry = (double)(oh - 1) / (double)tg_y;
        const float m00 = (float)rx, m11 = (float)ry;
        const float m02 = (float)(rx * (double)__fadd_rn(__fmul_rn(-0.5f, in_x), __fmul_rn(0.5f, tg_x)));
        const float m12 = (float)(ry * (double)__fadd_rn(__fmul_rn(-0.5f, in_y), __fmul_rn(0.5f, tg_y)));
        // scipy_warp_affine inverts the fp32 matrix in fp64
        m.ax = 1.0 / (double)m00;
        m.bx = -(double)m02 / (double)m00;
        m.ay = 1.0 / (double)m11;
        m.by = -(double)m12 / (double)m11;
    } else {
        double cx = (double)box.x + w * 0.5, cy = (double)box.y + h * 0.5;
        if (aspect > 1.0) cx += w * 0.5 * (aspect - 1.0);
        else cy += h * 0.5 * (1.0 / aspect - 1.0);
        const double r = (double)ow / w;
        m.ax = 1.0 / r;
        m.bx = cx - (double)ow * 0.5 / r;
        m.ay = 1.0 / r;
        m.by = cy - (double)oh * 0.5 / r;
    }
    return m;
}

// (i0, t): sample = v[i0]*(1-t) + v[i0+1]*t with i0 + 1 always inside the axis (at the far edge the pair
// (n-2, t=1) stands for (n-1, t=0)); i0 < 0 marks "outside the frame -> 0".  n >= 2.
// fp32 frames keep an fp32 weight.  uint8 frames (whose result is rounded to an integer, so the decisive
// interpolation has to be as exact as scipy's fp64) keep the fp64 weight next to its fp32 rounding.
template <typename T>
struct AxisEntry;
template <>
struct __align__(8) AxisEntry<float> {      // one 8-byte shared-memory load per entry
    int i0;
    float t;
    __device__ __forceinline__ void set(int i, double w) { i0 = i; t = (float)w; }
};
template <>
struct __align__(16) AxisEntry<unsigned char> {
    int i0;
    float t;
    double td;
    __device__ __forceinline__ void set(int i, double w) { i0 = i; t = (float)w; td = w; }
};
template <typename T>
__device__ __forceinline__ AxisEntry<T> axis_entry(double s, int n) {
    AxisEntry<T> e;
    if (!(s >= 0.0 && s <= (double)(n - 1))) {
        e.set(-1, 0.0);
        return e;
    }
    const double f = floor(s);
    int i0 = (int)f;
    double t = s - f;
    if (i0 >= n - 1) {
        i0 = n - 2;
        t = 1.0;
    }
    e.set(i0, t);
    return e;
}

// ---- per-pixel arithmetic ---------------------------------------------------------------------------
// float frames: vertical lerp + (v - mean)/std as one FMA
__device__ __forceinline__ float finish_px(float top, float bot, float wy, float inv_sd, float nmean) {
    return fmaf(fmaf(bot - top, wy, top), inv_sd, nmean);
}
// uint8 frames, exact path: fp64 lerps, then scipy's integer output conversion (round half up, clamp to
// 0..255) before (q - mean) / std.
__device__ __forceinline__ float exact_px_u8(unsigned p00, unsigned p01, unsigned p10, unsigned p11, double wx, double wy,
                                             float inv_sd, float nmean) {
    const double a = (double)p00, b = (double)p01, c = (double)p10, d = (double)p11;
    const double top = fma(b - a, wx, a);
    const double bot = fma(d - c, wx, c);
    double v = fma(bot - top, wy, top) + 0.5;
    v = v > 255.0 ? 255.0 : v;
    return fmaf((float)(int)v, inv_sd, nmean);        // v >= 0.5: truncation == floor
}
// uint8 frames, fast path.  The fp32 estimate v of the interpolated value is within 1e-4 of the fp64 one
// (|v| <= 255: four fp32 roundings of <= 1.6e-5 each plus the fp32 weights).  r = nearest integer (2^23
// trick, two FADDs); unless v is within kU8Guard of a tie (x.5), r == floor(exact + 0.5) and the fp32 result
// IS scipy's.  Near a tie (2 * kU8Guard of the pixels on noise; every pixel of e.g. an exact 2x upscale)
// the caller recomputes in fp64.
constexpr float kU8Guard = 1.0f / 1024.0f;
// No clamp to 255 on this path: both lerps are convex combinations of values <= 255 with fp32 weights in [0, 1], so
// v <= 255 + 2^-16 and its nearest integer is at most 255.
__device__ __forceinline__ bool fast_px_u8(float top, float bot, float wy, float inv_sd, float nmean, float &out) {
    const float v = fmaf(bot - top, wy, top);
    const float r = __fsub_rn(__fadd_rn(v, 8388608.0f), 8388608.0f);
    const float fr = fabsf(v - r);
    out = fmaf(r, inv_sd, nmean);
    return fr < 0.5f - kU8Guard;
}

// shared-memory loads by 32-bit address (volatile: never moved across the mbarrier waits / block barriers)
__device__ __forceinline__ float lds_px(uint32_t addr, float) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_px1(uint32_t addr, float) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds_u8(uint32_t addr) {
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <typename T>
__device__ __forceinline__ float hlerp_s(uint32_t addr, float wx) {
    if constexpr (std::is_same<T, unsigned char>::value) {
        // 2^23 + a and 2^23 + b as floats (one logic op each); their difference is exactly b - a, so only a is de-biased
        const float A = __int_as_float(0x4B000000 | (int)lds_u8(addr)), B = __int_as_float(0x4B000000 | (int)lds_u8(addr + 1));
        return fmaf(B - A, wx, A - 8388608.0f);
    } else {
        const float a = lds_px(addr, T()), b = lds_px1(addr, T());
        return fmaf(b - a, wx, a);
    }
}

constexpr int kStageBytesDefault = 20 * 1024;   // one source band buffer of a CTA, fp32 frames
constexpr int kStageBytesDefaultU8 = 12 * 1024; // uint8 frames (rows are a quarter of the bytes)
constexpr int kCropWarps = 6;                   // warps per CTA = column chunks x row groups

// Coordinate tables of one CTA: xt[0..ow), yt[ry0..ry1] (indexed by output row), and the valid ranges
// s_v = {first, last valid column, first, last valid row}.  Ends with a block barrier.
template <typename T>
__device__ __forceinline__ void build_tables(const AxisMap &m, const CropParams &prm, AxisEntry<T> *xt, AxisEntry<T> *yt, int ry0, int ry1,
                                             int *s_v) {
    const int tid = threadIdx.x, nthreads = blockDim.x, ow = prm.ow;
    int x_lo = INT_MAX, x_hi = -1, y_lo = INT_MAX, y_hi = -1;       // this thread's valid entries
    for (int i = tid; i < ow + (ry1 - ry0 + 1); i += nthreads) {
        if (i < ow) {
            const AxisEntry<T> e = axis_entry<T>(m.ax * (double)i + m.bx, prm.fw);
            xt[i] = e;
            if (e.i0 >= 0) { x_lo = min(x_lo, i); x_hi = max(x_hi, i); }
        } else {
            const int y = ry0 + i - ow;
            const AxisEntry<T> e = axis_entry<T>(m.ay * (double)y + m.by, prm.fh);
            yt[y] = e;
            if (e.i0 >= 0) { y_lo = min(y_lo, y); y_hi = max(y_hi, y); }
        }
    }
    // warp reductions, then one shared-memory atomic per warp and quantity (not two per table entry)
    x_lo = __reduce_min_sync(FULL, x_lo); x_hi = __reduce_max_sync(FULL, x_hi);
    y_lo = __reduce_min_sync(FULL, y_lo); y_hi = __reduce_max_sync(FULL, y_hi);
    if ((tid & 31) == 0) {
        if (x_hi >= 0) { atomicMin(&s_v[0], x_lo); atomicMax(&s_v[1], x_hi); }
        if (y_hi >= 0) { atomicMin(&s_v[2], y_lo); atomicMax(&s_v[3], y_hi); }
    }
    __syncthreads();
}

// Streaming store of one output element: fp32, or rounded to bf16 (round to nearest even — what `.to(torch.bfloat16)` does to
// the fp32 result; for a backbone that runs under bf16 autocast it halves the bytes this HBM-bound op writes).
__device__ __forceinline__ void store_px(float *p, float v) { __stcs(p, v); }
__device__ __forceinline__ void store_px(__nv_bfloat16 *p, float v) {
    __stcs(reinterpret_cast<unsigned short *>(p), __bfloat16_as_ushort(__float2bfloat16_rn(v)));
}

// Direct global gathers for output rows [ry0, ry1] of one (crop, channel): no monotonicity assumption on the source
// map.  Used by the staging-free fallback kernel and, inside the staged kernel, for mirrored crops (a box with negative
// width AND height gives a negative scale on both axes; HF / scipy then produce the mirrored crop).
template <typename T, typename O>
__device__ __forceinline__ void direct_gather_rows(const CropParams &prm, const T *src, O *dst, const AxisEntry<T> *xt,
                                                   const AxisEntry<T> *yt, int ry0, int ry1, float inv_sd, float nmean, int tid,
                                                   int nthreads) {
    constexpr bool kU8 = std::is_same<T, unsigned char>::value;
    const int ow = prm.ow;
    for (int x = tid; x < ow; x += nthreads) {
        const AxisEntry<T> ex = xt[x];
        O *o = dst + (size_t)ry0 * ow + x;
        for (int y = ry0; y <= ry1; ++y, o += ow) {
            const AxisEntry<T> ey = yt[y];
            float v = nmean;
            if (ex.i0 >= 0 && ey.i0 >= 0) {
                const T *ra = src + (size_t)ey.i0 * prm.fw + ex.i0;
                const T *rb = ra + prm.fw;
                const T p00 = __ldg(ra), p01 = __ldg(ra + 1), p10 = __ldg(rb), p11 = __ldg(rb + 1);
                if constexpr (kU8) {
                    v = exact_px_u8(p00, p01, p10, p11, ex.td, ey.td, inv_sd, nmean);
                } else {
                    const float top = fmaf(p01 - p00, ex.t, p00), bot = fmaf(p11 - p10, ex.t, p10);
                    v = finish_px(top, bot, ey.t, inv_sd, nmean);
                }
            }
            store_px(o, v);
        }
    }
}

// One CTA per (crop, channel, slab of output rows).
//   * coordinate tables (fp64 -> index + weight) for the out_w columns and the slab's rows in smem;
//   * the valid output rows are processed in bands; for each band the needed source rows, restricted to
//     the needed (16 B aligned) column range, are copied into a shared-memory band buffer by bulk-TMA row
//     copies (cp.async.bulk) issued by a dedicated PRODUCER warp — coalesced full-line HBM reads instead of
//     four scattered gathers per output pixel.  `stages` band buffers form a ring: full[s] (transaction
//     mbarrier) tells the consumers the band has landed, empty[s] (one arrival per consumer warp) tells the
//     producer the buffer can be refilled; there is no block-wide barrier after the tables are built;
//   * a CONSUMER warp owns a chunk of 32*C output columns (lane + 32 j, j < C) and one of the row groups of
//     every band: the per-row work (table read, row address, loop) is paid once per C pixels.  The source
//     row index is warp-uniform, so the horizontal lerps of a source row are carried in registers from one
//     output row to the next whenever consecutive output rows share it;
//   * one coalesced 4-byte streaming store per pixel.
// Measured on B200 (cfg2, 640 crops): more resident CTAs beat deeper rings (per-CTA start-up = box load +
// fp64 tables + first band), hence 2 stages of 20 KB (fp32) / 12 KB (uint8) -> 5-7 CTAs per SM.
// Needs 16-byte aligned source rows (bulk TMA) and room for 2 source rows of the widest window in a band
// buffer; anything else goes to crop_affine_direct_kernel.
// FULL: out_w is a multiple of the 32*C columns a warp covers, so no per-pixel column bound is needed.
template <typename T, typename O, int C, bool FULL>
__global__ void __launch_bounds__(32 * (kCropWarps + 1)) crop_affine_kernel(const CropParams prm) {
    using Entry = AxisEntry<T>;
    constexpr bool kU8 = std::is_same<T, unsigned char>::value;
    constexpr int kAlign = 16 / (int)sizeof(T);                                          // elements per 16 bytes
    extern __shared__ __align__(128) unsigned char crop_smem[];
    const int ow = prm.ow, oh = prm.oh;
    const int kStageBytes = prm.stage_bytes;
    // 1-D grid of items in (slab, crop, channel) order, slab fastest: the slabs of a crop and the crops of a frame are
    // resident together, so the source rows that overlapping boxes share come from L2 (see crop_stream_kernel).
    const int nsl = prm.split;
    const int slab = (oh + nsl - 1) / nsl;
    const int item_z = (int)(blockIdx.x % (unsigned)nsl), item_pc = (int)(blockIdx.x / (unsigned)nsl);
    const int ry0 = item_z * slab;
    const int ry1 = (ry0 + slab < oh ? ry0 + slab : oh) - 1;                             // inclusive
    if (ry0 > ry1) return;
    Entry *xt = reinterpret_cast<Entry *>(crop_smem + (size_t)prm.stages * kStageBytes); // [ow]
    Entry *yt = xt + ow - ry0;                                                           // [slab], indexed by output row
    uint64_t *full = reinterpret_cast<uint64_t *>(xt + ow + slab);                       // [stages] band landed
    uint64_t *empty = full + prm.stages;                                                 // [stages] band released by the consumers
    __shared__ int s_v[4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
    const int p = item_pc % prm.P, c = item_pc / prm.P;
    __shared__ AxisMap s_map;
    int f = __ldg(prm.frame_idx + p);
    if (warp == 0) {          // the fp64 map (about ten fp64 divisions) is built by one warp, not by all seven
        const float4 box = __ldg(reinterpret_cast<const float4 *>(prm.boxes) + p);
        const AxisMap mm = crop_axis_map(box, ow, oh, prm.variant);
        if (lane == 0) s_map = mm;
    }
    if (tid == 32) {
        s_v[0] = ow; s_v[1] = -1; s_v[2] = oh; s_v[3] = -1;
        for (int i = 0; i < prm.stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], prm.ncc * prm.rg);
        }
        mbar_fence_init();
        fence_proxy_async();
    }
    __syncthreads();
    const AxisMap m = s_map;
    build_tables<T>(m, prm, xt, yt, ry0, ry1, s_v);
    const int vx0 = s_v[0], vx1 = s_v[1], vy0 = s_v[2], vy1 = s_v[3];

    f = f < 0 ? 0 : (f >= prm.num_frames ? prm.num_frames - 1 : f);
    const T *src = static_cast<const T *>(prm.frames) + ((size_t)f * 3 + c) * prm.fh * prm.fw;
    O *dst = static_cast<O *>(prm.out) + ((size_t)p * 3 + c) * oh * ow;
    const float mean = c == 0 ? prm.mean[0] : (c == 1 ? prm.mean[1] : prm.mean[2]);     // constant-bank selects
    const float inv_sd = 1.0f / (c == 0 ? prm.stdv[0] : (c == 1 ? prm.stdv[1] : prm.stdv[2]));
    const float nmean = -mean * inv_sd;              // out = v * inv_sd + nmean
    const float zero_out = nmean;

    const bool any = vx1 >= vx0 && vy1 >= vy0;
    if (any && (m.ax < 0.0 || m.ay < 0.0)) {         // mirrored crop: the band staging below assumes a non-decreasing map
        direct_gather_rows<T, O>(prm, src, dst, xt, yt, ry0, ry1, inv_sd, nmean, threadIdx.x, blockDim.x);
        return;
    }
    // rows with no valid source: constant
    for (int y = ry0 + warp; y <= ry1; y += nthreads / 32) {
        if (any && y >= vy0 && y <= vy1) continue;
        for (int x = lane; x < ow; x += 32) store_px(dst + (size_t)y * ow + x, zero_out);
    }
    if (!any) return;

    // source column window, 16-byte aligned; xt[].i0 + 1 is always a valid column
    const int cx0 = xt[vx0].i0 & ~(kAlign - 1);
    int cx1 = (xt[vx1].i0 + 2 + kAlign - 1) & ~(kAlign - 1);      // exclusive
    if (cx1 > prm.fw) cx1 = prm.fw;
    const int row_elems = cx1 - cx0;
    const int pitch = row_elems * (int)sizeof(T);                 // bytes
    const int rows_cap = kStageBytes / pitch;                     // >= 2 (checked on the host for the widest window)
    const int ncc = prm.ncc, rg = prm.rg;                         // column chunks x row groups = warps of the CTA
    int band;
    {
        const double per = m.ay > 0.0 ? m.ay : 1.0;
        const double br = floor((double)(rows_cap - 3) / per) + 1.0;
        band = br > (double)(ry1 - ry0 + 1) ? ry1 - ry0 + 1 : (int)br;
        if (band > rg) band = band / rg * rg;                     // whole row groups
        if (band < 1) band = 1;
    }
    const int nbands = (vy1 - vy0 + band) / band;
    const int rpg = (band + rg - 1) / rg;                         // rows per group

    // ---- producer warp: bulk-copy the source rows of band b into buffer b % S once its previous tenant
    //      (band b - S) has been released by every consumer warp ---------------------------------------------
    const int S = prm.stages, lgS = prm.stages_log2;         // S = 1 << lgS: stage and phase of band b by mask and shift
    if (warp == ncc * rg) {
        for (int b = 0; b < nbands; ++b) {
            const int s = b & (S - 1);
            if (b >= S) mbar_wait(&empty[s], (uint32_t)(((b >> lgS) - 1) & 1));
            const int r0 = vy0 + b * band;
            const int r1 = (r0 + band - 1) < vy1 ? (r0 + band - 1) : vy1;
            const int sy_lo = yt[r0].i0;
            const int nrows = yt[r1].i0 + 1 - sy_lo + 1;
            unsigned char *buf = crop_smem + (size_t)s * kStageBytes;
            if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)(nrows * pitch));
            __syncwarp();
            for (int r = lane; r < nrows; r += 32)
                bulk_g2s(buf + (size_t)r * pitch, src + (size_t)(sy_lo + r) * prm.fw + cx0, (uint32_t)pitch, &full[s]);
        }
        return;
    }

    // ---- consumer warps ----------------------------------------------------------------------------------------
    // this warp's columns (fixed for the whole slab): byte offset inside a staged row, weight, liveness
    const int cc = warp % ncc, grp = warp / ncc;
    uint32_t coff[C];
    float wx[C];
    float isd[C];                       // 1/std of a live column, 0 of a dead one: (finite sample) * 0 + nmean = the constant
    const int xbase = cc * 32 * C + lane;
    const int safe_col = xt[vx0].i0;
#pragma unroll
    for (int j = 0; j < C; ++j) {
        const int x = xbase + 32 * j;
        const Entry ex = xt[x < ow ? x : ow - 1];
        const bool live = ex.i0 >= 0;
        coff[j] = (uint32_t)(((live ? ex.i0 : safe_col) - cx0) * (int)sizeof(T));   // dead columns read a staged address
        wx[j] = ex.t;
        isd[j] = live ? inv_sd : 0.0f;
    }

    const uint32_t smem_base = smem_u32(crop_smem);
    for (int b = 0; b < nbands; ++b) {
        const int s = b & (S - 1);
        const int r0 = vy0 + b * band;
        const int r1 = (r0 + band - 1) < vy1 ? (r0 + band - 1) : vy1;
        mbar_wait(&full[s], (uint32_t)((b >> lgS) & 1));

        const int ya = r0 + grp * rpg;
        const int yb = (ya + rpg - 1) < r1 ? (ya + rpg - 1) : r1;
        // byte address of source row 0 of the plane as if the whole plane were staged
        const uint32_t tile = smem_base + (uint32_t)(s * kStageBytes) - (uint32_t)(yt[r0].i0 * pitch);
        float top[C], bot[C];
        int prev = INT_MIN;
        O *o = dst + (size_t)ya * ow + xbase;
        for (int y = ya; y <= yb; ++y, o += ow) {
            // (index, fp32 weight) only: the fp64 weight of a uint8 entry is read on the rare exact path
            const int2 ey2 = *reinterpret_cast<const int2 *>(&yt[y]);
            const int ey_i0 = ey2.x;
            const float ey_t = __int_as_float(ey2.y);
            const uint32_t ra = tile + (uint32_t)(ey_i0 * pitch);
            const uint32_t rb = ra + (uint32_t)pitch;
            if (ey_i0 == prev + 1) {
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    top[j] = bot[j];
                    bot[j] = hlerp_s<T>(rb + coff[j], wx[j]);
                }
            } else if (ey_i0 != prev) {
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    top[j] = hlerp_s<T>(ra + coff[j], wx[j]);
                    bot[j] = hlerp_s<T>(rb + coff[j], wx[j]);
                }
            }
            prev = ey_i0;
            if constexpr (kU8) {
                float v[C];
                unsigned bad = 0u;
#pragma unroll
                for (int j = 0; j < C; ++j)
                    if (!fast_px_u8(top[j], bot[j], ey_t, isd[j], nmean, v[j])) bad |= 1u << j;
                if (bad) {                           // near a rounding tie (0.2 % of the pixels on noise): fp64, as scipy computes it
                    const double wyd = yt[y].td;
#pragma unroll
                    for (int j = 0; j < C; ++j) {
                        if ((bad >> j) & 1u) {
                            const int x = xbase + 32 * j;
                            const uint32_t a0 = ra + coff[j], a1 = rb + coff[j];
                            v[j] = exact_px_u8(lds_u8(a0), lds_u8(a0 + 1), lds_u8(a1), lds_u8(a1 + 1), xt[x < ow ? x : ow - 1].td, wyd,
                                               isd[j], nmean);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < C; ++j)
                    if (FULL || xbase + 32 * j < ow) store_px(o + 32 * j, v[j]);
            } else {
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    const float v = finish_px(top[j], bot[j], ey_t, isd[j], nmean);
                    if (FULL || xbase + 32 * j < ow) store_px(o + 32 * j, v);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);       // this warp is done with the buffer
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent variant (spp_crop_affine*_ws, the default of ops.crop_affine): a PLAN kernel + a STREAM kernel.
//
// Why.  Per-CTA traces of the one-CTA-per-item kernel above (cfg2, 3 840 CTAs of 12-51 us) show where its time goes beyond
// the bytes: every CTA spends ~6 us (box -> fp64 map -> tables -> first band over a loaded HBM queue) before its first
// pixel, with both of its band buffers empty; the last CTA starts at 128 us and the machine drains until 157 us; time vs
// crop count is 18 us + 145 us per 640 crops.  Shorter items would shrink the drain but multiply the set-up.
//   * crop_plan_kernel, one CTA per crop: the fp64 map and both coordinate tables ONCE per crop (not once per channel and
//     slab), plus one 64-byte descriptor per row slab (valid ranges, source window, band layout), into the workspace;
//   * crop_stream_kernel, resident CTAs pulling (slab, crop, channel) tickets from a counter in the workspace: the set-up of
//     an item is three bulk-TMA copies (x table, the slab's y entries, descriptor) into the spare half of a double-buffered
//     table area, issued by the producer warp one item ahead and completing on an mbarrier — no arithmetic, no block
//     barrier — so the producer goes straight from the last band of item k to the first band of item k + 1 while the
//     consumers are still on item k, and items can be 64-row slabs (12 per crop).
// Barriers: full[2] / empty[2] as above, indexed by a band counter that runs across items; ready[2] (tables + descriptor of
// the item in half h have landed; a transaction barrier) and done[2] (one arrival per consumer warp: half h may be reused).
// The ticket for item k + 2 is requested while item k + 1 is fetched: a global atomic takes microseconds under a saturated
// HBM queue and must never sit between two band copies.
struct __align__(16) CropItem {
    int kind;                 // kItemStaged / kItemConstant / kItemMirrored / kItemStop
    int p, f, ry0, ry1;       // crop, frame, output rows [ry0, ry1]
    int vx0, vy0, vy1;        // first valid column, valid output rows of the slab
    int cx0, pitch;           // source window start (elements), staged row pitch (bytes)
    int band, nbands, rpg;    // output rows per band, number of bands, rows per row group
    int safe_off;             // staged byte offset that dead output columns read
    int pad[2];
};
static_assert(sizeof(CropItem) == 64, "CropItem is copied by bulk TMA");
enum { kItemStaged = 0, kItemConstant = 1, kItemMirrored = 2, kItemStop = 3 };
constexpr int kPlanHeader = 256;              // workspace: [header: ticket counter][tables P x (ow + oh)][items P x nslabs]
constexpr int kPlanMaxSlabs = 32;
// header word 1: written by the plan kernel, checked by the stream kernel — a run on a workspace that was never planned for
// this (crop count, slab count) traps instead of producing crops from stale tables
__host__ __device__ __forceinline__ unsigned plan_tag(int P, int nslabs) { return 0x53505043u ^ ((unsigned)P * 2654435761u) ^ ((unsigned)nslabs << 24); }

template <typename T>
__global__ void __launch_bounds__(256) crop_plan_kernel(const CropParams prm) {
    using Entry = AxisEntry<T>;
    constexpr int kAlign = 16 / (int)sizeof(T);
    const int ow = prm.ow, oh = prm.oh, p = blockIdx.x, tid = threadIdx.x;
    const int slab = (oh + prm.split - 1) / prm.split, nslabs = (oh + slab - 1) / slab;
    Entry *tab = reinterpret_cast<Entry *>(prm.ws + kPlanHeader) + (size_t)p * (ow + oh);
    CropItem *items = reinterpret_cast<CropItem *>(prm.ws + kPlanHeader + (size_t)prm.P * (ow + oh) * sizeof(Entry)) + (size_t)p * nslabs;
    __shared__ AxisMap s_map;
    __shared__ int s_x[2], s_y[kPlanMaxSlabs][2];
    if (p == 0 && tid == 0) {
        reinterpret_cast<unsigned int *>(prm.ws)[0] = 0u;                            // the stream kernel's ticket counter
        reinterpret_cast<unsigned int *>(prm.ws)[1] = plan_tag(prm.P, nslabs);
    }
    if (tid < 32) {
        const float4 box = __ldg(reinterpret_cast<const float4 *>(prm.boxes) + p);
        const AxisMap mm = crop_axis_map(box, ow, oh, prm.variant);
        if (tid == 0) { s_map = mm; s_x[0] = INT_MAX; s_x[1] = -1; }
    } else if (tid - 32 < nslabs) {
        s_y[tid - 32][0] = INT_MAX; s_y[tid - 32][1] = -1;
    }
    __syncthreads();
    const AxisMap m = s_map;
    for (int i = tid; i < ow + oh; i += blockDim.x) {
        if (i < ow) {
            const Entry e = axis_entry<T>(m.ax * (double)i + m.bx, prm.fw);
            tab[i] = e;
            if (e.i0 >= 0) { atomicMin(&s_x[0], i); atomicMax(&s_x[1], i); }
        } else {
            const int y = i - ow;
            const Entry e = axis_entry<T>(m.ay * (double)y + m.by, prm.fh);
            tab[i] = e;
            if (e.i0 >= 0) { atomicMin(&s_y[y / slab][0], y); atomicMax(&s_y[y / slab][1], y); }
        }
    }
    __syncthreads();                                   // also orders this CTA's table stores before the reads below
    if (tid < nslabs) {
        const int z = tid;
        int f = __ldg(prm.frame_idx + p);
        f = f < 0 ? 0 : (f >= prm.num_frames ? prm.num_frames - 1 : f);
        CropItem it;
        it.p = p; it.f = f; it.ry0 = z * slab; it.ry1 = (it.ry0 + slab < oh ? it.ry0 + slab : oh) - 1;
        const int vx0 = s_x[0], vx1 = s_x[1], vy0 = s_y[z][0], vy1 = s_y[z][1];
        it.vx0 = vx0; it.vy0 = vy0; it.vy1 = vy1;
        it.cx0 = 0; it.pitch = 0; it.band = 1; it.nbands = 0; it.rpg = 1; it.safe_off = 0; it.pad[0] = it.pad[1] = 0;
        if (!(vx1 >= vx0 && vy1 >= vy0)) {
            it.kind = kItemConstant;
        } else if (m.ax < 0.0 || m.ay < 0.0) {         // mirrored crop: the band staging assumes a non-decreasing source map
            it.kind = kItemMirrored;
        } else {
            it.kind = kItemStaged;
            const int i_lo = tab[vx0].i0, i_hi = tab[vx1].i0;
            const int cx0 = i_lo & ~(kAlign - 1);
            int cx1 = (i_hi + 2 + kAlign - 1) & ~(kAlign - 1);
            if (cx1 > prm.fw) cx1 = prm.fw;
            const int pitch = (cx1 - cx0) * (int)sizeof(T);
            const int rows_cap = prm.stage_bytes / pitch;
            const int rg = prm.rg, nrows = it.ry1 - it.ry0 + 1;
            const double per = m.ay > 0.0 ? m.ay : 1.0;
            const double br = floor((double)(rows_cap - 3) / per) + 1.0;
            int band = br > (double)nrows ? nrows : (int)br;
            if (band > rg) band = band / rg * rg;
            if (band < 1) band = 1;
            it.cx0 = cx0; it.pitch = pitch; it.band = band;
            it.nbands = (vy1 - vy0 + band) / band;
            it.rpg = (band + rg - 1) / rg;
            it.safe_off = (i_lo - cx0) * (int)sizeof(T);
        }
        items[z] = it;
    }
}

// Event log of one CTA of the stream kernel (trace builds only): (type, item, band) + %globaltimer per event.
#ifdef SPP_CROP_TRACE
__device__ __forceinline__ void crop_trace(const CropParams &prm, int type, int k, int b) {
    if (prm.trace && blockIdx.x == 7 && (threadIdx.x & 31) == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const unsigned long long i = atomicAdd(prm.trace, 1ull);
        if (i < 4000) {
            prm.trace[1 + 2 * i] = ((unsigned long long)type << 48) | ((unsigned long long)k << 24) | (unsigned)b;
            prm.trace[2 + 2 * i] = t;
        }
    }
}
#define SPP_CROP_EVENT(type, k, b) crop_trace(prm, type, k, b)
#else
#define SPP_CROP_EVENT(type, k, b) do { } while (0)
#endif

template <typename T, typename O, int C, bool FULL>
__global__ void __launch_bounds__(32 * (kCropWarps + 1), 5) crop_stream_kernel(const CropParams prm) {
    using Entry = AxisEntry<T>;
    constexpr bool kU8 = std::is_same<T, unsigned char>::value;
    extern __shared__ __align__(128) unsigned char crop_smem[];
    const int ow = prm.ow, oh = prm.oh, P = prm.P;
    const int kStageBytes = prm.stage_bytes;
    const int slab = (oh + prm.split - 1) / prm.split, nslabs = (oh + slab - 1) / slab;
    const int total = P * 3 * nslabs;
    const int tab_n = ow + slab;
    Entry *tab = reinterpret_cast<Entry *>(crop_smem + 2 * (size_t)kStageBytes);        // [2][ow + slab]
    CropItem *items = reinterpret_cast<CropItem *>(tab + 2 * (size_t)tab_n);            // [2]
    uint64_t *full = reinterpret_cast<uint64_t *>(items + 2);                           // [2] band landed
    uint64_t *empty = full + 2;                                                         // [2] band released by the consumers
    uint64_t *ready = full + 4;                                                         // [2] tables + descriptor landed
    uint64_t *done = full + 6;                                                          // [2] item finished by the consumers
    int *s_chan = reinterpret_cast<int *>(full + 8);                                    // [2] channel of the item in each half

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncc = prm.ncc, rg = prm.rg, nwarps = ncc * rg;                            // consumer warps
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], nwarps);
            mbar_init(&ready[i], 1);
            mbar_init(&done[i], nwarps);
        }
        mbar_fence_init();
        fence_proxy_async();
    }
    __syncthreads();

    // ---- producer warp ---------------------------------------------------------------------------------------------
    if (warp == nwarps) {
        unsigned int *tk = reinterpret_cast<unsigned int *>(prm.ws);
        const Entry *gtab = reinterpret_cast<const Entry *>(prm.ws + kPlanHeader);
        const CropItem *gitems = reinterpret_cast<const CropItem *>(prm.ws + kPlanHeader + (size_t)P * (ow + oh) * sizeof(Entry));
        // fetch item t into half h: descriptor + tables by bulk TMA (or the stop mark), completing on ready[h]
        auto fetch = [&](int t, int h) {
            if (lane != 0) return;
            if (t >= total) {
                items[h].kind = kItemStop;
                mbar_arrive(&ready[h]);
                return;
            }
            // Tickets: slab fastest, then crop, then channel — the slabs of a crop and the crops of a frame are in flight at
            // the same time, so the source rows that overlapping boxes share are read from HBM once and from L2 afterwards
            // (crop fastest, then slab: 2.72 GB of DRAM reads at 2 560 crops against 2.23 GB for the per-item kernel).
            const int pz = t % (P * nslabs), c = t / (P * nslabs);
            const int z = pz % nslabs, p = pz / nslabs;
            const int ry0 = z * slab, nrows = (ry0 + slab < oh ? ry0 + slab : oh) - ry0;
            s_chan[h] = c;
            Entry *xt = tab + (size_t)h * tab_n;
            const uint32_t bx = (uint32_t)(ow * sizeof(Entry)), by = (uint32_t)(nrows * sizeof(Entry));
            mbar_arrive_expect_tx(&ready[h], bx + by + (uint32_t)sizeof(CropItem));   // release: s_chan
            bulk_g2s(xt, gtab + (size_t)p * (ow + oh), bx, &ready[h]);
            bulk_g2s(xt + ow, gtab + (size_t)p * (ow + oh) + ow + ry0, by, &ready[h]);
            bulk_g2s(&items[h], gitems + (size_t)p * nslabs + z, (uint32_t)sizeof(CropItem), &ready[h]);
        };
        const unsigned tag = lane == 0 ? __ldcg(reinterpret_cast<const unsigned int *>(prm.ws) + 1) : 0u;   // checked below, off the critical path
        fetch((int)blockIdx.x, 0);
        int t_next = 0;                                    // lane 0: ticket of item k + 1, requested one item early
        if (lane == 0) t_next = (int)gridDim.x + (int)atomicAdd(tk, 1u);
        if (lane == 0 && tag != plan_tag(P, nslabs)) {
            if (blockIdx.x == 0) printf("crop_stream_kernel: the workspace holds no plan for %d crops x %d slabs (spp_crop_plan first)\n", P, nslabs);
            __trap();
        }
        uint32_t gb = 0;                                   // bands issued so far, over all items
        for (int k = 0;; ++k) {
            mbar_wait(&ready[k & 1], (uint32_t)((k >> 1) & 1));
            const CropItem it = items[k & 1];
            if (it.kind == kItemStop) break;
            // Item k + 1 goes into the other half as soon as the consumers have left it (item k - 1) — but its three copies and
            // the ticket request for item k + 2 are issued AFTER item k's first band copy, never in front of it.
            // (The wait is mbar_wait, i.e. try_wait: polling this barrier with mbarrier.test_wait between the band copies and
            // fetching on success reproducibly killed the launch with an illegal-instruction error on B200.)
            bool pending = true;
            auto next_item = [&]() {
                if (k >= 1) mbar_wait(&done[(k + 1) & 1], (uint32_t)(((k - 1) >> 1) & 1));
                fetch(t_next, (k + 1) & 1);
                if (lane == 0) t_next = (int)gridDim.x + (int)atomicAdd(tk, 1u);
                pending = false;
            };
            if (it.kind == kItemStaged) {
                const int c = s_chan[k & 1];
                const T *src = static_cast<const T *>(prm.frames) + ((size_t)it.f * 3 + c) * prm.fh * prm.fw;
                const Entry *yt = tab + (size_t)(k & 1) * tab_n + ow - it.ry0;
                for (int b = 0; b < it.nbands; ++b, ++gb) {
                    const int s = (int)(gb & 1u);
                    if (gb >= 2u) mbar_wait(&empty[s], ((gb >> 1) - 1u) & 1u);
                    const int r0 = it.vy0 + b * it.band;
                    const int r1 = (r0 + it.band - 1) < it.vy1 ? (r0 + it.band - 1) : it.vy1;
                    const int sy_lo = yt[r0].i0;
                    const int nrows = yt[r1].i0 + 1 - sy_lo + 1;
                    unsigned char *buf = crop_smem + (size_t)s * kStageBytes;
                    SPP_CROP_EVENT(1, k, b);
                    if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)(nrows * it.pitch));
                    __syncwarp();
                    for (int r = lane; r < nrows; r += 32)
                        bulk_g2s(buf + (size_t)r * it.pitch, src + (size_t)(sy_lo + r) * prm.fw + it.cx0, (uint32_t)it.pitch, &full[s]);
                    if (pending) next_item();
                }
            }
            SPP_CROP_EVENT(2, k, 0);
            if (pending) next_item();
        }
        return;
    }

    // ---- consumer warps --------------------------------------------------------------------------------------------
    const int cc = warp % ncc, grp = warp / ncc;
    const int xbase = cc * 32 * C + lane;
    const uint32_t smem_base = smem_u32(crop_smem);
    uint32_t gb = 0;
    for (int k = 0;; ++k) {
        mbar_wait(&ready[k & 1], (uint32_t)((k >> 1) & 1));
        const CropItem it = items[k & 1];
        if (warp == 0) SPP_CROP_EVENT(10, k, it.kind);
        if (it.kind == kItemStop) break;
        const Entry *xt = tab + (size_t)(k & 1) * tab_n;
        const Entry *yt = xt + ow - it.ry0;
        const int c = s_chan[k & 1];
        O *dst = static_cast<O *>(prm.out) + ((size_t)it.p * 3 + c) * oh * ow;
        const float mean = c == 0 ? prm.mean[0] : (c == 1 ? prm.mean[1] : prm.mean[2]);
        const float inv_sd = 1.0f / (c == 0 ? prm.stdv[0] : (c == 1 ? prm.stdv[1] : prm.stdv[2]));
        const float nmean = -mean * inv_sd;
        if (it.kind == kItemMirrored) {
            const T *src = static_cast<const T *>(prm.frames) + ((size_t)it.f * 3 + c) * prm.fh * prm.fw;
            direct_gather_rows<T, O>(prm, src, dst, xt, yt, it.ry0, it.ry1, inv_sd, nmean, tid, 32 * nwarps);
        } else {
            const bool staged = it.kind == kItemStaged;
            for (int y = it.ry0 + warp; y <= it.ry1; y += nwarps) {          // rows with no valid source: constant
                if (staged && y >= it.vy0 && y <= it.vy1) continue;
                for (int x = lane; x < ow; x += 32) store_px(dst + (size_t)y * ow + x, nmean);
            }
            if (staged) {
                const int pitch = it.pitch;
                uint32_t coff[C];
                float wx[C], isd[C];
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    const int x = xbase + 32 * j;
                    const int2 e2 = *reinterpret_cast<const int2 *>(&xt[x < ow ? x : ow - 1]);
                    const bool live = e2.x >= 0;
                    coff[j] = live ? (uint32_t)((e2.x - it.cx0) * (int)sizeof(T)) : (uint32_t)it.safe_off;
                    wx[j] = __int_as_float(e2.y);
                    isd[j] = live ? inv_sd : 0.0f;
                }
                for (int b = 0; b < it.nbands; ++b, ++gb) {
                    const int s = (int)(gb & 1u);
                    const int r0 = it.vy0 + b * it.band;
                    const int r1 = (r0 + it.band - 1) < it.vy1 ? (r0 + it.band - 1) : it.vy1;
                    if (warp == 0) SPP_CROP_EVENT(11, k, b);
                    mbar_wait(&full[s], (gb >> 1) & 1u);
                    if (warp == 0) SPP_CROP_EVENT(12, k, b);
                    const int ya = r0 + grp * it.rpg;
                    const int yb = (ya + it.rpg - 1) < r1 ? (ya + it.rpg - 1) : r1;
                    const uint32_t tile = smem_base + (uint32_t)(s * kStageBytes) - (uint32_t)(yt[r0].i0 * pitch);
                    float top[C], bot[C];
                    int prev = INT_MIN;
                    O *o = dst + (size_t)ya * ow + xbase;
                    for (int y = ya; y <= yb; ++y, o += ow) {
                        const int2 ey2 = *reinterpret_cast<const int2 *>(&yt[y]);
                        const int ey_i0 = ey2.x;
                        const float ey_t = __int_as_float(ey2.y);
                        const uint32_t ra = tile + (uint32_t)(ey_i0 * pitch);
                        const uint32_t rb = ra + (uint32_t)pitch;
                        if (ey_i0 == prev + 1) {
#pragma unroll
                            for (int j = 0; j < C; ++j) {
                                top[j] = bot[j];
                                bot[j] = hlerp_s<T>(rb + coff[j], wx[j]);
                            }
                        } else if (ey_i0 != prev) {
#pragma unroll
                            for (int j = 0; j < C; ++j) {
                                top[j] = hlerp_s<T>(ra + coff[j], wx[j]);
                                bot[j] = hlerp_s<T>(rb + coff[j], wx[j]);
                            }
                        }
                        prev = ey_i0;
                        if constexpr (kU8) {
                            float v[C];
                            unsigned bad = 0u;
#pragma unroll
                            for (int j = 0; j < C; ++j)
                                if (!fast_px_u8(top[j], bot[j], ey_t, isd[j], nmean, v[j])) bad |= 1u << j;
                            if (bad) {
                                const double wyd = yt[y].td;
#pragma unroll
                                for (int j = 0; j < C; ++j) {
                                    if ((bad >> j) & 1u) {
                                        const int x = xbase + 32 * j;
                                        const uint32_t a0 = ra + coff[j], a1 = rb + coff[j];
                                        v[j] = exact_px_u8(lds_u8(a0), lds_u8(a0 + 1), lds_u8(a1), lds_u8(a1 + 1), xt[x < ow ? x : ow - 1].td,
                                                           wyd, isd[j], nmean);
                                    }
                                }
                            }
#pragma unroll
                            for (int j = 0; j < C; ++j)
                                if (FULL || xbase + 32 * j < ow) store_px(o + 32 * j, v[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < C; ++j) {
                                const float v = finish_px(top[j], bot[j], ey_t, isd[j], nmean);
                                if (FULL || xbase + 32 * j < ow) store_px(o + 32 * j, v);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s]);
                }
            }
        }
        __syncwarp();
        if (warp == 0) SPP_CROP_EVENT(13, k, 0);
        if (lane == 0) mbar_arrive(&done[k & 1]);            // this warp no longer reads half k & 1
    }
}

// Fallback without staging (frame rows not 16-byte aligned, or a source window too wide for a band buffer):
// the same tables, direct global gathers, one thread per output column.
template <typename T, typename O>
__global__ void __launch_bounds__(256) crop_affine_direct_kernel(const CropParams prm) {
    using Entry = AxisEntry<T>;
    extern __shared__ __align__(128) unsigned char crop_smem[];
    const int ow = prm.ow, oh = prm.oh;
    Entry *xt = reinterpret_cast<Entry *>(crop_smem);
    Entry *yt = xt + ow;
    __shared__ int s_v[4];
    const int tid = threadIdx.x;
    const int p = blockIdx.x, c = blockIdx.y;
    const float4 box = __ldg(reinterpret_cast<const float4 *>(prm.boxes) + p);
    const AxisMap m = crop_axis_map(box, ow, oh, prm.variant);
    if (tid == 0) { s_v[0] = ow; s_v[1] = -1; s_v[2] = oh; s_v[3] = -1; }
    __syncthreads();
    build_tables<T>(m, prm, xt, yt, 0, oh - 1, s_v);
    int f = __ldg(prm.frame_idx + p);
    f = f < 0 ? 0 : (f >= prm.num_frames ? prm.num_frames - 1 : f);
    const T *src = static_cast<const T *>(prm.frames) + ((size_t)f * 3 + c) * prm.fh * prm.fw;
    O *dst = static_cast<O *>(prm.out) + ((size_t)p * 3 + c) * oh * ow;
    const float mean = c == 0 ? prm.mean[0] : (c == 1 ? prm.mean[1] : prm.mean[2]);
    const float inv_sd = 1.0f / (c == 0 ? prm.stdv[0] : (c == 1 ? prm.stdv[1] : prm.stdv[2]));
    const float nmean = -mean * inv_sd;
    direct_gather_rows<T, O>(prm, src, dst, xt, yt, 0, oh - 1, inv_sd, nmean, threadIdx.x, blockDim.x);
}

}  // namespace
}  // namespace spp

namespace spp {
namespace {
int env_int(const char *name, int dflt, int lo, int hi) {
    const char *e = getenv(name);
    if (!e) return dflt;
    const int v = atoi(e);
    return (v < lo || v > hi) ? dflt : v;
}

template <typename T, typename O, int C, bool FULL>
int launch_staged_full(const CropParams &prm, dim3 grid, size_t smem, cudaStream_t st) {
    // per device and per context: set on every launch (about a microsecond; legal during stream capture)
    SPP_CHECK_CUDA(cudaFuncSetAttribute(crop_affine_kernel<T, O, C, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    crop_affine_kernel<T, O, C, FULL><<<grid, 32 * (prm.ncc * prm.rg + 1), smem, st>>>(prm);
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}

enum { kCropPlanAndRun = 0, kCropPlanOnly = 1, kCropRunOnly = 2 };
template <typename T, typename O, int C, bool FULL>
int launch_stream_full(const CropParams &prm, int total, size_t smem, cudaStream_t st, int mode) {
    SPP_CHECK_CUDA(cudaFuncSetAttribute(crop_stream_kernel<T, O, C, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int occ = 0;
    SPP_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, crop_stream_kernel<T, O, C, FULL>, 32 * (kCropWarps + 1), smem));
    if (occ < 1) occ = 1;
    const int sms = sm_count() > 0 ? sm_count() : 148;
    int slots = occ * sms - launch_limit(2);               // SPP_LIMIT_CROP_FREE_CTAS: slots left to kernels enqueued beside this one
    if (slots < sms) slots = sms;
    const int grid = total < slots ? total : slots;
    if (mode != kCropRunOnly) {
        crop_plan_kernel<T><<<prm.P, 256, 0, st>>>(prm);
        SPP_CHECK_LAUNCH();
    }
    if (mode != kCropPlanOnly) {
        crop_stream_kernel<T, O, C, FULL><<<grid, 32 * (kCropWarps + 1), smem, st>>>(prm);
        SPP_CHECK_LAUNCH();
    }
    return SPP_OK;
}
template <typename T, typename O, int C>
int launch_stream(const CropParams &prm, int total, size_t smem, cudaStream_t st, int mode) {
    return prm.ow % (32 * C) == 0 ? launch_stream_full<T, O, C, true>(prm, total, smem, st, mode) : launch_stream_full<T, O, C, false>(prm, total, smem, st, mode);
}

// Row slabs of the persistent kernel (SPP_CROP_SPLIT overrides), at most kPlanMaxSlabs.  Measured at cfg2 with `rows` = 64 / 32 /
// 16: fp32 frames 156 / 145 / 160 us, uint8 frames 154 / 163 / 191 us; at cfg4 (6 400 crops) fp32 1 029 / 1 116 / 1 327 us.
std::atomic<int> &crop_policy() {
    static std::atomic<int> policy{env_int("SPP_CROP_PERSIST", 1, 0, 2)};
    return policy;
}
int stream_split(int out_h, int split_env, int rows) {
    int split = split_env ? split_env : (out_h + rows - 1) / rows;
    if (split > kPlanMaxSlabs) split = kPlanMaxSlabs;
    if (split > out_h) split = out_h;
    return split < 1 ? 1 : split;
}
template <typename T>
size_t stream_workspace_bytes(int p, int out_h, int out_w, int split) {
    const int slab = (out_h + split - 1) / split, nslabs = (out_h + slab - 1) / slab;
    return (size_t)kPlanHeader + (size_t)p * (out_w + out_h) * sizeof(AxisEntry<T>) + (size_t)p * nslabs * sizeof(CropItem);
}

template <typename T, typename O, int C>
int launch_staged(const CropParams &prm, dim3 grid, size_t smem, cudaStream_t st) {
    return prm.ow % (32 * C) == 0 ? launch_staged_full<T, O, C, true>(prm, grid, smem, st) : launch_staged_full<T, O, C, false>(prm, grid, smem, st);
}

template <typename T, typename O = float>
int launch_crop(const void *frames, int num_frames, int frame_h, int frame_w, const float *boxes, const int *frame_idx, int p,
                int out_h, int out_w, const float *mean, const float *std, int variant, void *out, spp_stream_t stream,
                void *workspace = nullptr, size_t workspace_bytes = 0, int mode = kCropPlanAndRun) {
    if (p == 0) return SPP_OK;
    SPP_CHECK_ARG(boxes && frame_idx && mean && std && (mode == kCropPlanOnly || (frames && out)), "crop_affine: null pointer");
    SPP_CHECK_ARG(num_frames > 0 && frame_h >= 2 && frame_w >= 2 && p >= 0, "crop_affine: frames must be at least 2x2");
    SPP_CHECK_ARG(out_h > 0 && out_w > 0 && out_w <= 2048 && out_h <= 2048, "crop_affine: output size must be within 2048 x 2048");
    SPP_CHECK_ARG(variant == SPP_CROP_HF_UDP || variant == SPP_CROP_GLUONCV, "crop_affine: unknown variant %d", variant);
    SPP_CHECK_ARG((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "crop_affine: boxes must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CropParams prm{};
    prm.frames = frames; prm.boxes = boxes; prm.frame_idx = frame_idx; prm.out = out;
    prm.num_frames = num_frames; prm.fh = frame_h; prm.fw = frame_w; prm.P = p; prm.oh = out_h; prm.ow = out_w;
    prm.variant = variant;
    for (int c = 0; c < 3; ++c) { prm.mean[c] = mean[c]; prm.stdv[c] = std[c]; }
    // tuning knobs: SPP_CROP_STAGE_KB (one band buffer; a CTA has two), SPP_CROP_SPLIT (row slabs per crop channel),
    // SPP_CROP_COLS (output columns per lane)
    static const int stage_kb = env_int("SPP_CROP_STAGE_KB", (sizeof(T) == 1 ? kStageBytesDefaultU8 : kStageBytesDefault) / 1024, 2, 80);
    static const int split_env = env_int("SPP_CROP_SPLIT", 0, 1, 64);
    static const int cols = env_int("SPP_CROP_COLS", 3, 3, 6) == 6 ? 6 : 3;
    static const int stages_log2 = [] { const int v = env_int("SPP_CROP_STAGES", 2, 2, 8); return v >= 8 ? 3 : (v >= 4 ? 2 : 1); }();
    prm.stage_bytes = stage_kb * 1024;
    prm.stages_log2 = stages_log2;
    prm.stages = 1 << stages_log2;
    // Bulk-TMA row copies need 16-byte aligned row segments (base and row pitch multiples of 16 bytes), and the widest
    // possible source window (the whole frame width) has to leave room for the 2 source rows of a 1-row band.
    prm.ncc = (out_w + 32 * cols - 1) / (32 * cols);
    prm.rg = prm.ncc <= kCropWarps ? kCropWarps / prm.ncc : 1;
    const bool staged = ((size_t)frame_w * sizeof(T) % 16 == 0) && ((reinterpret_cast<uintptr_t>(frames) & 15) == 0) &&
                        ((size_t)prm.stage_bytes / ((size_t)frame_w * sizeof(T)) >= 2) && prm.ncc <= kCropWarps;
    if (!staged) {
        if (mode == kCropPlanOnly) return SPP_OK;      // nothing to plan: the run call takes the staging-free kernel
        const size_t smem = (size_t)(out_w + out_h) * sizeof(AxisEntry<T>);
        SPP_CHECK_ARG(smem <= 160 * 1024, "crop_affine: output size too large");
        SPP_CHECK_CUDA(cudaFuncSetAttribute(crop_affine_direct_kernel<T, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        int threads = (out_w + 31) / 32 * 32;
        if (threads > 256) threads = 256;
        crop_affine_direct_kernel<T, O><<<dim3(p, 3), threads, smem, st>>>(prm);
        SPP_CHECK_LAUNCH();
        return SPP_OK;
    }
    // Persistent plan + stream kernels when the caller brings a workspace (spp_crop_affine*_ws).  They need exactly
    // kCropWarps consumer warps and table slices that bulk TMA can copy (16-byte multiples: even sizes for 8-byte entries).
    // Policy (spp_crop_policy: 0 never / 1 automatic / 2 always; SPP_CROP_PERSIST sets the initial value).  Automatic = fp32
    // frames, up to 4 096 crops: measured 145 / 276 / 538 us against 152 / 308 / 599 for the per-item kernel at 640 / 1 280 / 2 560
    // crops of ten per frame, but 1 029-1 037 against 1 013 us at 6 400 crops of a hundred per frame (cfg4: most source rows are
    // L2 hits either way and the per-item kernel's single 256-row item per crop channel has the least overhead); with uint8 frames
    // the kernel is bound by its instruction stream and the per-item kernel is faster (149 against 154 us at cfg2, 1 190 against
    // 1 281 at cfg4).
    const int persist = crop_policy().load(std::memory_order_relaxed);
    if (workspace && (persist == 2 || (persist == 1 && sizeof(T) == 4 && p <= 4096)) && prm.ncc * prm.rg == kCropWarps) {
        // 32-row slabs (short items, short drain) up to 4 096 crops (2 560 crops of ten per frame: 538 us against 579 with 64-row
        // slabs), 64-row slabs beyond (6 400 crops of a hundred per frame: 1 029 against 1 116 us) and for uint8 frames
        const int psplit = stream_split(out_h, split_env, (sizeof(T) == 1 || p > 4096) ? 64 : 32);
        const int pslab = (out_h + psplit - 1) / psplit, nslabs = (out_h + pslab - 1) / pslab;
        const bool even = sizeof(AxisEntry<T>) % 16 == 0 || (out_w % 2 == 0 && out_h % 2 == 0 && pslab % 2 == 0);
        if (even && nslabs <= kPlanMaxSlabs) {
            SPP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "crop_affine: workspace must be 16-byte aligned");
            SPP_CHECK_ARG(workspace_bytes >= stream_workspace_bytes<T>(p, out_h, out_w, psplit), "crop_affine: workspace too small (%zu bytes)",
                          workspace_bytes);
            prm.split = psplit;
            prm.ws = static_cast<unsigned char *>(workspace);
#ifdef SPP_CROP_TRACE
            { const char *e = getenv("SPP_CROP_TRACE_PTR"); prm.trace = e ? reinterpret_cast<unsigned long long *>(strtoull(e, nullptr, 10)) : nullptr; }
#endif
            prm.stages = 2; prm.stages_log2 = 1;
            const size_t psmem = 2 * (size_t)prm.stage_bytes + 2 * (size_t)(out_w + pslab) * sizeof(AxisEntry<T>) + 2 * sizeof(CropItem) +
                                 8 * sizeof(uint64_t) + 16;
            SPP_CHECK_ARG(psmem <= 200 * 1024, "crop_affine: output size too large");
            const long long total = (long long)p * 3 * nslabs;
            SPP_CHECK_ARG(total < (1LL << 30), "crop_affine: too many crops");
            return cols == 3 ? launch_stream<T, O, 3>(prm, (int)total, psmem, st, mode) : launch_stream<T, O, 6>(prm, (int)total, psmem, st, mode);
        }
    }
    if (mode == kCropPlanOnly) return SPP_OK;          // shapes the persistent kernels do not take: the run call does all the work
    // Slabs of output rows: enough CTAs that the last wave is a small part of the launch, but never slabs shorter
    // than 64 rows (the tables and the first band are per-CTA overhead).
    int split = split_env;
    if (split == 0) {
        const int sms = sm_count() > 0 ? sm_count() : 148;
        split = 1;
        while (split < 8 && (long long)p * 3 * split < 16LL * sms && out_h / (split * 2) >= 64) split *= 2;
    }
    if (split > out_h) split = out_h;
    const int slab = (out_h + split - 1) / split;
    const size_t smem = (size_t)prm.stages * prm.stage_bytes + (size_t)(out_w + slab) * sizeof(AxisEntry<T>) + 16 * (size_t)prm.stages;
    SPP_CHECK_ARG(smem <= 200 * 1024, "crop_affine: output size too large");
    prm.split = split;
    SPP_CHECK_ARG((long long)p * 3 * split < (1LL << 31), "crop_affine: too many crops");
    dim3 grid((unsigned)((long long)p * 3 * split));
    return cols == 3 ? launch_staged<T, O, 3>(prm, grid, smem, st) : launch_staged<T, O, 6>(prm, grid, smem, st);
}
}  // namespace
}  // namespace spp

extern "C" int spp_crop_affine(const float *frames, int num_frames, int frame_h, int frame_w, const float *boxes,
                               const int *frame_idx, int p, int out_h, int out_w, const float *mean, const float *std,
                               int variant, float *out, spp_stream_t stream) {
    return spp::launch_crop<float>(frames, num_frames, frame_h, frame_w, boxes, frame_idx, p, out_h, out_w, mean, std, variant, out, stream);
}

// Everything in one call: frames fp32 or uint8, output fp32 or bf16, with or without a workspace, planned ahead or not.
extern "C" int spp_crop_affine_ex(const void *frames, int frames_u8, int num_frames, int frame_h, int frame_w, const float *boxes,
                                  const int *frame_idx, int p, int out_h, int out_w, const float *mean, const float *std, int variant,
                                  void *out, int out_bf16, void *workspace, size_t workspace_bytes, int planned, spp_stream_t stream) {
    SPP_CHECK_ARG(!planned || workspace, "crop_affine_ex: planned = 1 needs the workspace spp_crop_plan wrote");
    const int mode = planned ? spp::kCropRunOnly : spp::kCropPlanAndRun;
    if (frames_u8)
        return out_bf16 ? spp::launch_crop<unsigned char, __nv_bfloat16>(frames, num_frames, frame_h, frame_w, boxes, frame_idx, p, out_h, out_w, mean,
                                                                        std, variant, out, stream, workspace, workspace_bytes, mode)
                        : spp::launch_crop<unsigned char, float>(frames, num_frames, frame_h, frame_w, boxes, frame_idx, p, out_h, out_w, mean, std,
                                                                 variant, out, stream, workspace, workspace_bytes, mode);
    return out_bf16 ? spp::launch_crop<float, __nv_bfloat16>(frames, num_frames, frame_h, frame_w, boxes, frame_idx, p, out_h, out_w, mean, std, variant,
                                                            out, stream, workspace, workspace_bytes, mode)
                    : spp::launch_crop<float, float>(frames, num_frames, frame_h, frame_w, boxes, frame_idx, p, out_h, out_w, mean, std, variant, out,
                                                     stream, workspace, workspace_bytes, mode);
}

extern "C" int spp_crop_policy(int mode) {
    if (mode < 0) return spp::crop_policy().load();
    if (mode > 2) return -1;
    return spp::crop_policy().exchange(mode);
}

extern "C" size_t spp_crop_workspace_bytes(int p, int out_h, int out_w, int frames_u8) {
    if (p <= 0 || out_h <= 0 || out_w <= 0) return 0;
    // sized for the largest slab count the launcher may choose (SPP_CROP_SPLIT is read there)
    const int split = spp::kPlanMaxSlabs < out_h ? spp::kPlanMaxSlabs : out_h;
    return frames_u8 ? spp::stream_workspace_bytes<unsigned char>(p, out_h, out_w, split) : spp::stream_workspace_bytes<float>(p, out_h, out_w, split);
}

extern "C" int spp_crop_affine_ws(const float *frames, int num_frames, int frame_h, int frame_w, const float *boxes,
                                  const int *frame_idx, int p, int out_h, int out_w, const float *mean, const float *std,
                                  int variant, float *out, void *workspace, size_t workspace_bytes, spp_stream_t stream) {
    return spp::launch_crop<float>(frames, num_frames, frame_h, frame_w, boxes, frame_idx, p, out_h, out_w, mean, std, variant, out, stream,
                                   workspace, workspace_bytes);
}

// Plan and run as two calls (same arguments, same workspace; the run must follow its plan on the device): the plan only
// reads the boxes, so a caller can enqueue it early — SelectivePosePipeline runs it beside the heatmap decode.
extern "C" int spp_crop_plan(int frames_u8, int num_frames, int frame_h, int frame_w, const float *boxes, const int *frame_idx, int p,
                             int out_h, int out_w, int variant, void *workspace, size_t workspace_bytes, spp_stream_t stream) {
    const float unit[3] = {0.f, 0.f, 0.f}, one[3] = {1.f, 1.f, 1.f};
    SPP_CHECK_ARG(workspace, "crop_plan: workspace required");
    return frames_u8 ? spp::launch_crop<unsigned char>(nullptr, num_frames, frame_h, frame_w, boxes, frame_idx, p, out_h, out_w, unit, one, variant,
                                                       nullptr, stream, workspace, workspace_bytes, spp::kCropPlanOnly)
                     : spp::launch_crop<float>(nullptr, num_frames, frame_h, frame_w, boxes, frame_idx, p, out_h, out_w, unit, one, variant, nullptr,
                                               stream, workspace, workspace_bytes, spp::kCropPlanOnly);
}
extern "C" int spp_crop_affine_run(const float *frames, int num_frames, int frame_h, int frame_w, const float *boxes,
                                   const int *frame_idx, int p, int out_h, int out_w, const float *mean, const float *std,
                                   int variant, float *out, void *workspace, size_t workspace_bytes, spp_stream_t stream) {
    SPP_CHECK_ARG(workspace, "crop_affine_run: workspace required");
    return spp::launch_crop<float>(frames, num_frames, frame_h, frame_w, boxes, frame_idx, p, out_h, out_w, mean, std, variant, out, stream,
                                   workspace, workspace_bytes, spp::kCropRunOnly);
}
extern "C" int spp_crop_affine_u8_run(const uint8_t *frames, int num_frames, int frame_h, int frame_w, const float *boxes,
                                      const int *frame_idx, int p, int out_h, int out_w, const float *mean, const float *std,
                                      int variant, float *out, void *workspace, size_t workspace_bytes, spp_stream_t stream) {
    SPP_CHECK_ARG(workspace, "crop_affine_run: workspace required");
    return spp::launch_crop<unsigned char>(frames, num_frames, frame_h, frame_w, boxes, frame_idx, p, out_h, out_w, mean, std, variant, out,
                                           stream, workspace, workspace_bytes, spp::kCropRunOnly);
}

extern "C" int spp_crop_affine_u8_ws(const uint8_t *frames, int num_frames, int frame_h, int frame_w, const float *boxes,
                                     const int *frame_idx, int p, int out_h, int out_w, const float *mean, const float *std,
                                     int variant, float *out, void *workspace, size_t workspace_bytes, spp_stream_t stream) {
    return spp::launch_crop<unsigned char>(frames, num_frames, frame_h, frame_w, boxes, frame_idx, p, out_h, out_w, mean, std, variant, out,
                                           stream, workspace, workspace_bytes);
}

extern "C" int spp_crop_affine_u8(const uint8_t *frames, int num_frames, int frame_h, int frame_w, const float *boxes,
                                  const int *frame_idx, int p, int out_h, int out_w, const float *mean, const float *std,
                                  int variant, float *out, spp_stream_t stream) {
    return spp::launch_crop<unsigned char>(frames, num_frames, frame_h, frame_w, boxes, frame_idx, p, out_h, out_w, mean, std, variant, out, stream);
}
