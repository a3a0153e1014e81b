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

namespace spp {
namespace {

struct CropParams {
    const float *frames;
    const float *boxes;
    const int *frame_idx;
    float *out;
    int num_frames, fh, fw, P, oh, ow, variant;
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

// (i0, i1, t): sample = v[i0]*(1-t) + v[i1]*t; i0 < 0 marks "outside the frame -> 0"
__device__ __forceinline__ void axis_entry(double s, int n, int &i0, int &i1, float &t) {
    if (!(s >= 0.0 && s <= (double)(n - 1))) {
        i0 = -1; i1 = 0; t = 0.f;
        return;
    }
    const double f = floor(s);
    i0 = (int)f;
    i1 = i0 + 1 < n ? i0 + 1 : n - 1;
    t = (float)(s - f);
}

__global__ void __launch_bounds__(256) crop_affine_kernel(const CropParams prm) {
    extern __shared__ __align__(16) unsigned char crop_smem[];
    const int ow = prm.ow, oh = prm.oh;
    int *x0 = reinterpret_cast<int *>(crop_smem);
    int *x1 = x0 + ow;
    float *tx = reinterpret_cast<float *>(x1 + ow);
    int *y0 = reinterpret_cast<int *>(tx + ow);
    int *y1 = y0 + oh;
    float *ty = reinterpret_cast<float *>(y1 + oh);

    const int p = blockIdx.x, c = blockIdx.y;
    const float4 box = __ldg(reinterpret_cast<const float4 *>(prm.boxes) + p);
    const AxisMap m = crop_axis_map(box, ow, oh, prm.variant);
    for (int i = threadIdx.x; i < ow + oh; i += blockDim.x) {
        if (i < ow) axis_entry(m.ax * (double)i + m.bx, prm.fw, x0[i], x1[i], tx[i]);
        else axis_entry(m.ay * (double)(i - ow) + m.by, prm.fh, y0[i - ow], y1[i - ow], ty[i - ow]);
    }
    __syncthreads();

    int f = __ldg(prm.frame_idx + p);
    f = f < 0 ? 0 : (f >= prm.num_frames ? prm.num_frames - 1 : f);
    const float *src = prm.frames + ((size_t)f * 3 + c) * prm.fh * prm.fw;
    float *dst = prm.out + ((size_t)p * 3 + c) * oh * ow;
    const float mean = prm.mean[c], sd = prm.stdv[c];
    const float zero_out = __fdiv_rn(__fsub_rn(0.f, mean), sd);

    const int ow4 = ow >> 2;
    const int rows_per_pass = blockDim.x / ow4;
    const int xq = threadIdx.x % ow4, yr = threadIdx.x / ow4;
    if (yr >= rows_per_pass) return;
    int ix0[4], ix1[4];
    float wx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ix0[k] = x0[xq * 4 + k];
        ix1[k] = x1[xq * 4 + k];
        wx[k] = tx[xq * 4 + k];
    }
    for (int y = yr; y < oh; y += rows_per_pass) {
        const int iy0 = y0[y], iy1 = y1[y];
        const float wy = ty[y];
        float o[4];
        if (iy0 < 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = zero_out;
        } else {
            const float *r0 = src + (size_t)iy0 * prm.fw, *r1 = src + (size_t)iy1 * prm.fw;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (ix0[k] < 0) {
                    o[k] = zero_out;
                } else {
                    const float p00 = __ldg(r0 + ix0[k]), p01 = __ldg(r0 + ix1[k]);
                    const float p10 = __ldg(r1 + ix0[k]), p11 = __ldg(r1 + ix1[k]);
                    const float top = fmaf(p01 - p00, wx[k], p00);
                    const float bot = fmaf(p11 - p10, wx[k], p10);
                    const float v = fmaf(bot - top, wy, top);
                    o[k] = __fdiv_rn(__fsub_rn(v, mean), sd);
                }
            }
        }
        __stcs(reinterpret_cast<float4 *>(dst + (size_t)y * ow) + xq, make_float4(o[0], o[1], o[2], o[3]));
    }
}

}  // namespace
}  // namespace spp

extern "C" int spp_crop_affine(const float *frames, int num_frames, int frame_h, int frame_w, const float *boxes,
                               const int *frame_idx, int p, int out_h, int out_w, const float *mean, const float *std,
                               int variant, float *out, spp_stream_t stream) {
    using namespace spp;
    SPP_CHECK_ARG(frames && boxes && frame_idx && out && mean && std, "crop_affine: null pointer");
    SPP_CHECK_ARG(num_frames > 0 && frame_h > 0 && frame_w > 0 && p >= 0, "crop_affine: bad frame shape");
    SPP_CHECK_ARG(out_h > 0 && out_w > 0 && out_w % 4 == 0 && out_w <= 1024, "crop_affine: out_w must be a multiple of 4, <= 1024");
    SPP_CHECK_ARG(variant == SPP_CROP_HF_UDP || variant == SPP_CROP_GLUONCV, "crop_affine: unknown variant %d", variant);
    SPP_CHECK_ARG((reinterpret_cast<uintptr_t>(boxes) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                  "crop_affine: boxes and out must be 16-byte aligned");
    SPP_CHECK_ARG(p <= 2147483647 / 1 && 3 <= 65535, "crop_affine: too many crops");
    if (p == 0) return SPP_OK;
    CropParams prm{};
    prm.frames = frames; prm.boxes = boxes; prm.frame_idx = frame_idx; prm.out = out;
    prm.num_frames = num_frames; prm.fh = frame_h; prm.fw = frame_w; prm.P = p; prm.oh = out_h; prm.ow = out_w;
    prm.variant = variant;
    for (int c = 0; c < 3; ++c) { prm.mean[c] = mean[c]; prm.stdv[c] = std[c]; }
    const int ow4 = out_w / 4;
    int threads = ow4 * (256 / ow4 > 0 ? 256 / ow4 : 1);
    if (threads < 32) threads = 32;
    const size_t smem = (size_t)(out_w + out_h) * 12;
    dim3 grid(p, 3);
    crop_affine_kernel<<<grid, threads, smem, static_cast<cudaStream_t>(stream)>>>(prm);
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}
