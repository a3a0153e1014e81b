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
// column, y_s only on the output row.  One CTA per (crop, channel) computes both coordinate tables
// ONCE in fp64 (the oracle inverts the fp32 matrix in fp64 and samples in fp64; an fp32 coordinate
// at x ~ 1000 px would already be off by 6e-5 px) and keeps (index, fp32 weight) pairs in shared
// memory; the sampling loop is fp32, four output pixels per thread, one 128-bit streaming store per
// thread per row.  HBM-bound: 4 B written per output element + the source ROI read once.
#include "spp_common.cuh"

#include <cstdlib>
#include <type_traits>

namespace spp {
namespace {

struct CropParams {
    const void *frames;         // [num_frames, 3, fh, fw] fp32 or uint8
    const float *boxes;
    const int *frame_idx;
    float *out;
    int num_frames, fh, fw, P, oh, ow, variant;
    int stage_bytes;            // size of the shared-memory band buffer
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
// Weight type: fp32 for float frames, fp64 for uint8 frames (whose result is rounded to an integer, so
// the interpolation itself has to be as exact as scipy's).
template <typename W>
struct AxisEntry {
    int i0;
    W t;
};
template <typename W>
__device__ __forceinline__ AxisEntry<W> axis_entry(double s, int n) {
    AxisEntry<W> e;
    if (!(s >= 0.0 && s <= (double)(n - 1))) {
        e.i0 = -1;
        e.t = (W)0;
        return e;
    }
    const double f = floor(s);
    e.i0 = (int)f;
    e.t = (W)(s - f);
    if (e.i0 >= n - 1) {
        e.i0 = n - 2;
        e.t = (W)1;
    }
    return e;
}

// bilinear sample + normalisation.  float frames: fp32 lerps.  uint8 frames: fp64 lerps, then scipy's
// integer output conversion (round half up, clamp to 0..255) before (q - mean) / std.
__device__ __forceinline__ float sample_px(float p00, float p01, float p10, float p11, float wx, float wy, float inv_sd, float nmean) {
    const float top = fmaf(p01 - p00, wx, p00);
    const float bot = fmaf(p11 - p10, wx, p10);
    return fmaf(fmaf(bot - top, wy, top), inv_sd, nmean);
}
__device__ __forceinline__ float sample_px(unsigned char p00, unsigned char p01, unsigned char p10, unsigned char p11, double wx,
                                           double wy, float inv_sd, float nmean) {
    const double a = (double)p00, b = (double)p01, c = (double)p10, d = (double)p11;
    const double top = fma(b - a, wx, a);
    const double bot = fma(d - c, wx, c);
    double v = fma(bot - top, wy, top) + 0.5;
    v = v > 255.0 ? 255.0 : v;
    return fmaf((float)(int)v, inv_sd, nmean);        // v >= 0.5: truncation == floor
}

constexpr int kStageBytesDefault = 40 * 1024;   // source band buffer per CTA -> 4-5 CTAs per SM

// One CTA per (crop, channel), one thread per output column.
//   * coordinate tables (fp64 -> index + fp32 weight) for the out_w columns and out_h rows in smem;
//   * the valid output rows are processed in bands; for each band the needed source rows, restricted to
//     the needed (16 B aligned) column range, are copied into shared memory by bulk-TMA row copies
//     (cp.async.bulk, issued by warp 0, all completing on one mbarrier) — coalesced full-line HBM reads
//     instead of four scattered 4-byte gathers per output pixel;
//   * per pixel: one 8-byte table read, 4 shared-memory reads at fixed offsets from one address, 7 FP
//     ops (bilinear + (v - mean)/std as one FMA) and one coalesced 4-byte streaming store.
// `staged == 0` (frame width not a multiple of 4, unaligned base, or a band that cannot fit): the same
// loop gathers straight from global memory.
template <typename T>
__global__ void __launch_bounds__(256) crop_affine_kernel(const CropParams prm, int staged) {
    using W = typename std::conditional<std::is_same<T, float>::value, float, double>::type;
    using Entry = AxisEntry<W>;
    constexpr int kAlign = 16 / (int)sizeof(T);                                          // elements per 16 bytes
    extern __shared__ __align__(128) unsigned char crop_smem[];
    const int ow = prm.ow, oh = prm.oh;
    const int kStageBytes = prm.stage_bytes;
    T *buf = reinterpret_cast<T *>(crop_smem);                                           // [stage_bytes]
    Entry *xt = reinterpret_cast<Entry *>(crop_smem + kStageBytes);                      // [ow]
    Entry *yt = xt + ow;                                                                 // [oh]
    uint64_t *bar = reinterpret_cast<uint64_t *>(yt + oh);
    __shared__ int s_v[4];   // first/last valid column, first/last valid row

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
    const int p = blockIdx.x, c = blockIdx.y;
    const float4 box = __ldg(reinterpret_cast<const float4 *>(prm.boxes) + p);
    const AxisMap m = crop_axis_map(box, ow, oh, prm.variant);
    if (tid == 0) {
        s_v[0] = ow; s_v[1] = -1; s_v[2] = oh; s_v[3] = -1;
        mbar_init(bar, 1);
        mbar_fence_init();
        fence_proxy_async();
    }
    __syncthreads();
    for (int i = tid; i < ow + oh; i += nthreads) {
        if (i < ow) {
            const Entry e = axis_entry<W>(m.ax * (double)i + m.bx, prm.fw);
            xt[i] = e;
            if (e.i0 >= 0) { atomicMin(&s_v[0], i); atomicMax(&s_v[1], i); }
        } else {
            const int y = i - ow;
            const Entry e = axis_entry<W>(m.ay * (double)y + m.by, prm.fh);
            yt[y] = e;
            if (e.i0 >= 0) { atomicMin(&s_v[2], y); atomicMax(&s_v[3], y); }
        }
    }
    __syncthreads();
    const int vx0 = s_v[0], vx1 = s_v[1], vy0 = s_v[2], vy1 = s_v[3];

    int f = __ldg(prm.frame_idx + p);
    f = f < 0 ? 0 : (f >= prm.num_frames ? prm.num_frames - 1 : f);
    const T *src = static_cast<const T *>(prm.frames) + ((size_t)f * 3 + c) * prm.fh * prm.fw;
    float *dst = prm.out + ((size_t)p * 3 + c) * oh * ow;
    const float mean = c == 0 ? prm.mean[0] : (c == 1 ? prm.mean[1] : prm.mean[2]);     // constant-bank selects
    const float inv_sd = 1.0f / (c == 0 ? prm.stdv[0] : (c == 1 ? prm.stdv[1] : prm.stdv[2]));
    const float nmean = -mean * inv_sd;              // out = v * inv_sd + nmean
    const float zero_out = nmean;

    const bool any = vx1 >= vx0 && vy1 >= vy0;
    // rows with no valid source: constant
    for (int y = warp; y < oh; y += nthreads / 32) {
        if (any && y >= vy0 && y <= vy1) continue;
        for (int x = lane; x < ow; x += 32) __stcs(dst + (size_t)y * ow + x, zero_out);
    }
    if (!any) return;

    // source column window, 16-byte aligned; xt[].i0 + 1 is always a valid column
    const int cx0 = xt[vx0].i0 & ~(kAlign - 1);
    int cx1 = (xt[vx1].i0 + 2 + kAlign - 1) & ~(kAlign - 1);      // exclusive
    if (cx1 > prm.fw) cx1 = prm.fw;
    const int row_elems = cx1 - cx0;
    const int rows_cap = kStageBytes / (row_elems * (int)sizeof(T));
    const bool use_stage = staged && rows_cap >= 3;
    int band = oh;
    if (use_stage) {
        const double per = m.ay > 0.0 ? m.ay : 1.0;
        double br = floor((double)(rows_cap - 3) / per) + 1.0;
        band = br > (double)oh ? oh : (int)br;
        if (band < 1) band = 1;
    }

    uint32_t parity = 0;
    for (int r0 = vy0; r0 <= vy1; r0 += band) {
        const int r1 = (r0 + band - 1) < vy1 ? (r0 + band - 1) : vy1;
        const int sy_lo = yt[r0].i0;
        if (use_stage) {
            const int nrows = yt[r1].i0 + 1 - sy_lo + 1;
            if (warp == 0) {
                if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(nrows * row_elems * (int)sizeof(T)));
                for (int r = lane; r < nrows; r += 32)
                    bulk_g2s(buf + (size_t)r * row_elems, src + (size_t)(sy_lo + r) * prm.fw + cx0,
                             (uint32_t)(row_elems * (int)sizeof(T)), bar);
            }
            mbar_wait(bar, parity);
            parity ^= 1;
        }
        for (int x = tid; x < ow; x += nthreads) {
            const Entry ex = xt[x];
            float *o = dst + (size_t)r0 * ow + x;
            if (ex.i0 < 0) {
                for (int y = r0; y <= r1; ++y, o += ow) __stcs(o, zero_out);
                continue;
            }
            const W wx = ex.t;
            if (use_stage) {
                // shared-memory tile: element (iy, ix) at buf[(iy - sy_lo) * row_elems + (ix - cx0)], 32-bit indices
                const int colbase = ex.i0 - cx0 - sy_lo * row_elems;
#pragma unroll 4
                for (int y = r0; y <= r1; ++y, o += ow) {
                    const Entry ey = yt[y];
                    const int ia = ey.i0 * row_elems + colbase;
                    const int ib = ia + row_elems;
                    __stcs(o, sample_px(buf[ia], buf[ia + 1], buf[ib], buf[ib + 1], wx, ey.t, inv_sd, nmean));
                }
            } else {
                const T *col = src + ex.i0;
#pragma unroll 4
                for (int y = r0; y <= r1; ++y, o += ow) {
                    const Entry ey = yt[y];
                    const T *ra = col + (size_t)ey.i0 * prm.fw;
                    const T *rb = ra + prm.fw;
                    __stcs(o, sample_px(__ldg(ra), __ldg(ra + 1), __ldg(rb), __ldg(rb + 1), wx, ey.t, inv_sd, nmean));
                }
            }
        }
        if (use_stage) __syncthreads();              // the band buffer is refilled next
    }
}

}  // namespace
}  // namespace spp

namespace spp {
namespace {
template <typename T>
int launch_crop(const void *frames, int num_frames, int frame_h, int frame_w, const float *boxes, const int *frame_idx, int p,
                int out_h, int out_w, const float *mean, const float *std, int variant, float *out, spp_stream_t stream) {
    if (p == 0) return SPP_OK;
    SPP_CHECK_ARG(frames && boxes && frame_idx && out && mean && std, "crop_affine: null pointer");
    SPP_CHECK_ARG(num_frames > 0 && frame_h >= 2 && frame_w >= 2 && p >= 0, "crop_affine: frames must be at least 2x2");
    SPP_CHECK_ARG(out_h > 0 && out_w > 0 && out_w <= 2048 && out_h <= 2048, "crop_affine: output size must be within 2048 x 2048");
    SPP_CHECK_ARG(variant == SPP_CROP_HF_UDP || variant == SPP_CROP_GLUONCV, "crop_affine: unknown variant %d", variant);
    SPP_CHECK_ARG((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "crop_affine: boxes must be 16-byte aligned");
    if (p == 0) return SPP_OK;
    CropParams prm{};
    prm.frames = frames; prm.boxes = boxes; prm.frame_idx = frame_idx; prm.out = out;
    prm.num_frames = num_frames; prm.fh = frame_h; prm.fw = frame_w; prm.P = p; prm.oh = out_h; prm.ow = out_w;
    prm.variant = variant;
    for (int c = 0; c < 3; ++c) { prm.mean[c] = mean[c]; prm.stdv[c] = std[c]; }
    // bulk-TMA row copies need 16-byte aligned row segments: base and row pitch multiples of 16 bytes
    const int staged = ((size_t)frame_w * sizeof(T) % 16 == 0) && ((reinterpret_cast<uintptr_t>(frames) & 15) == 0);
    using W = typename std::conditional<std::is_same<T, float>::value, float, double>::type;
    static int stage_kb = 0;
    if (stage_kb == 0) {                         // tuning knob; default 40 KB
        const char *e = getenv("SPP_CROP_STAGE_KB");
        stage_kb = e ? atoi(e) : kStageBytesDefault / 1024;
        if (stage_kb < 4 || stage_kb > 96) stage_kb = kStageBytesDefault / 1024;
    }
    prm.stage_bytes = stage_kb * 1024;
    const size_t smem = (size_t)prm.stage_bytes + (size_t)(out_w + out_h) * sizeof(AxisEntry<W>) + 16;
    int threads = (out_w + 31) / 32 * 32;
    if (threads > 256) threads = 256;
    static bool configured = false;
    if (!configured) {
        SPP_CHECK_CUDA(cudaFuncSetAttribute(crop_affine_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        configured = true;
    }
    SPP_CHECK_ARG(smem <= 100 * 1024, "crop_affine: output size too large");
    dim3 grid(p, 3);
    crop_affine_kernel<T><<<grid, threads, smem, static_cast<cudaStream_t>(stream)>>>(prm, staged);
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}
}  // namespace
}  // namespace spp

extern "C" int spp_crop_affine(const float *frames, int num_frames, int frame_h, int frame_w, const float *boxes,
                               const int *frame_idx, int p, int out_h, int out_w, const float *mean, const float *std,
                               int variant, float *out, spp_stream_t stream) {
    return spp::launch_crop<float>(frames, num_frames, frame_h, frame_w, boxes, frame_idx, p, out_h, out_w, mean, std, variant, out, stream);
}

extern "C" int spp_crop_affine_u8(const uint8_t *frames, int num_frames, int frame_h, int frame_w, const float *boxes,
                                  const int *frame_idx, int p, int out_h, int out_w, const float *mean, const float *std,
                                  int variant, float *out, spp_stream_t stream) {
    return spp::launch_crop<unsigned char>(frames, num_frames, frame_h, frame_w, boxes, frame_idx, p, out_h, out_w, mean, std, variant, out, stream);
}
