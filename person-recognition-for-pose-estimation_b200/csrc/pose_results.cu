// COCO keypoint result rows + OKS on the device.
//
// Replaces
//   the result loop of PoseEstimationModule.validation_step, training/lightning/pose_estimation/
//   module.py:534-549: x = kx * (x2 - x1) + x1, y = ky * (y2 - y1) + y1, v = 2 if score > thresh else 1,
//   instance score = mean of the keypoint scores (a triple Python loop with .item() per value);
//   the per-pair OKS of pycocotools COCOeval.computeOks as driven from module.py:598-615 (pycocotools is a
//   third-party dependency, not vendored and not installed here: the published formula is restated in
//   oracle/results.py — PARITY UNPINNED for the OKS part).
//
// Tiny, latency-bound kernels (16 B in / 12 B out per joint): one warp per person, lanes over joints.
#include "spp_common.cuh"

namespace spp {
namespace {

__global__ void __launch_bounds__(256) pose_results_kernel(const float *__restrict__ kpts, const float *__restrict__ scores,
                                                           const float *__restrict__ boxes, int p, int k, float thresh,
                                                           float *__restrict__ out, float *__restrict__ inst) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= p) return;
    float bx = 0.f, by = 0.f, bw = 1.f, bh = 1.f;
    if (boxes) {
        const float4 b = __ldg(reinterpret_cast<const float4 *>(boxes) + warp);     // x1 y1 x2 y2
        bx = b.x; by = b.y; bw = __fsub_rn(b.z, b.x); bh = __fsub_rn(b.w, b.y);
    }
    float sum = 0.f;
    for (int j = lane; j < k; j += 32) {
        const float2 c = __ldg(reinterpret_cast<const float2 *>(kpts) + (size_t)warp * k + j);
        const float s = __ldg(scores + (size_t)warp * k + j);
        float x = c.x, y = c.y;
        if (boxes) {
            x = __fadd_rn(__fmul_rn(x, bw), bx);
            y = __fadd_rn(__fmul_rn(y, bh), by);
        }
        float *o = out + ((size_t)warp * k + j) * 3;
        o[0] = x;
        o[1] = y;
        o[2] = s > thresh ? 2.0f : 1.0f;
        sum += s;
    }
    sum = warp_sum(sum);
    if (lane == 0 && inst) inst[warp] = sum / (float)k;
}

// COCOeval.computeOks for explicit (detection, ground truth) pairs, fp64 like numpy.
__global__ void __launch_bounds__(256) pose_oks_kernel(const float *__restrict__ pred, int pred_stride, const float *__restrict__ gt,
                                                       const float *__restrict__ gt_box, const float *__restrict__ area,
                                                       const float *__restrict__ sigmas, int p, int k, float *__restrict__ oks) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= p) return;
    int vis = 0;
    for (int j = lane; j < k; j += 32) vis += gt[((size_t)warp * k + j) * 3 + 2] > 0.f;
    vis = __reduce_add_sync(FULL, vis);
    const double a = (double)area[warp] + 2.220446049250313e-16;      // np.spacing(1)
    double x0 = 0, x1 = 0, y0 = 0, y1 = 0;
    if (vis == 0 && gt_box) {     // no labelled joint: distance to the doubled box (COCOeval)
        const float4 b = __ldg(reinterpret_cast<const float4 *>(gt_box) + warp);    // x y w h
        x0 = (double)b.x - (double)b.z; x1 = (double)b.x + (double)b.z * 2.0;
        y0 = (double)b.y - (double)b.w; y1 = (double)b.y + (double)b.w * 2.0;
    }
    double acc = 0.0;
    for (int j = lane; j < k; j += 32) {
        const float *g = gt + ((size_t)warp * k + j) * 3;
        const float *d = pred + ((size_t)warp * k + j) * pred_stride;
        double dx, dy;
        if (vis > 0) {
            if (!(g[2] > 0.f)) continue;
            dx = (double)d[0] - (double)g[0];
            dy = (double)d[1] - (double)g[1];
        } else {
            dx = fmax(0.0, x0 - (double)d[0]) + fmax(0.0, (double)d[0] - x1);
            dy = fmax(0.0, y0 - (double)d[1]) + fmax(0.0, (double)d[1] - y1);
        }
        const double s2 = (double)sigmas[j] * 2.0;
        const double e = (dx * dx + dy * dy) / (s2 * s2) / a / 2.0;
        acc += exp(-e);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    if (lane == 0) oks[warp] = (float)(acc / (double)(vis > 0 ? vis : k));
}

}  // namespace
}  // namespace spp

extern "C" int spp_pose_results(const float *keypoints, const float *scores, const float *boxes_xyxy, int p, int k,
                                float keypoint_thresh, float *out_keypoints, float *out_instance_score, spp_stream_t stream) {
    using namespace spp;
    SPP_CHECK_ARG(p >= 0 && k > 0, "pose_results: bad shape p=%d k=%d", p, k);
    if (p == 0) return SPP_OK;
    SPP_CHECK_ARG(keypoints && scores && out_keypoints, "pose_results: null pointer");
    SPP_CHECK_ARG((reinterpret_cast<uintptr_t>(keypoints) & 7) == 0 && (!boxes_xyxy || (reinterpret_cast<uintptr_t>(boxes_xyxy) & 15) == 0),
                  "pose_results: keypoints must be 8-byte and boxes 16-byte aligned");
    pose_results_kernel<<<(p + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(keypoints, scores, boxes_xyxy, p, k, keypoint_thresh,
                                                                                  out_keypoints, out_instance_score);
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}

extern "C" int spp_pose_oks(const float *pred, int pred_stride, const float *gt, const float *gt_boxes_xywh, const float *gt_area,
                            const float *sigmas, int p, int k, float *out_oks, spp_stream_t stream) {
    using namespace spp;
    SPP_CHECK_ARG(p >= 0 && k > 0, "pose_oks: bad shape p=%d k=%d", p, k);
    SPP_CHECK_ARG(pred_stride == 2 || pred_stride == 3, "pose_oks: predictions are [p, k, 2] or [p, k, 3] (got stride %d)", pred_stride);
    if (p == 0) return SPP_OK;
    SPP_CHECK_ARG(pred && gt && gt_area && sigmas && out_oks, "pose_oks: null pointer");
    SPP_CHECK_ARG(!gt_boxes_xywh || (reinterpret_cast<uintptr_t>(gt_boxes_xywh) & 15) == 0, "pose_oks: boxes must be 16-byte aligned");
    pose_oks_kernel<<<(p + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(pred, pred_stride, gt, gt_boxes_xywh, gt_area, sigmas, p, k,
                                                                              out_oks);
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}
