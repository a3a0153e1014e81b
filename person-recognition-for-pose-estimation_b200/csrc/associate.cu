// Face -> person association: the "selective" step between gallery match and crop.
//
// NOT in the reference (its inference glue is a TODO, scripts/modify_models.py:71-76):
// builder-defined, parity unpinned, restated in oracle/assoc.py.  Rule, per frame:
//   1. every face detection with a matched identity (id >= 0) picks the person detection that contains the
//      face centre (inclusive) and covers most of the face (largest intersection / face area; ties: the
//      person with the lower row = higher detection score);
//   2. a person is selected iff some face picked it; its identity is that of the lowest-row such face;
//   3. selected persons are emitted in row order as COCO (x, y, w, h) boxes, at most `cap` per frame.
// One CTA per frame; everything (<= 300 x 300 pairs) lives in shared memory.  Outputs are fixed-capacity
// (zero boxes / identity -1 padding) so the crop and heatmap-decode launches that follow need no host sync.
#include "spp_common.cuh"

namespace spp {
namespace {

constexpr int kAssocThreads = 256;

__global__ void __launch_bounds__(kAssocThreads) associate_kernel(const float *__restrict__ face_dets, const int *__restrict__ face_count,
                                                                  const int *__restrict__ face_ids, int face_cap,
                                                                  const float *__restrict__ person_dets, const int *__restrict__ person_count,
                                                                  int person_cap, int cap, float *__restrict__ out_boxes,
                                                                  int *__restrict__ out_ident, int *__restrict__ out_row, int *__restrict__ out_count) {
    extern __shared__ int assoc_smem[];
    int *pick = assoc_smem;                 // [face_cap]  person row picked by each face, -1 = none
    int *ident = pick + face_cap;           // [person_cap] identity per person, -1 = not selected
    __shared__ int s_n;
    const int b = blockIdx.x, tid = threadIdx.x;
    int nf = face_count[b], np = person_count[b];
    nf = nf < 0 ? ~nf : nf;                 // a negative count (~kept) flags a candidate overflow upstream; rows are still valid
    np = np < 0 ? ~np : np;
    nf = nf < face_cap ? nf : face_cap;
    np = np < person_cap ? np : person_cap;
    const float *fd = face_dets + (size_t)b * face_cap * 6;
    const float *pd = person_dets + (size_t)b * person_cap * 6;

    for (int f = tid; f < nf; f += kAssocThreads) {
        int best = -1;
        if (face_ids[(size_t)b * face_cap + f] >= 0) {
            const float fx1 = fd[f * 6], fy1 = fd[f * 6 + 1], fx2 = fd[f * 6 + 2], fy2 = fd[f * 6 + 3];
            const float cx = __fmul_rn(__fadd_rn(fx1, fx2), 0.5f), cy = __fmul_rn(__fadd_rn(fy1, fy2), 0.5f);
            const float farea = __fmul_rn(__fsub_rn(fx2, fx1), __fsub_rn(fy2, fy1));
            float bestv = -1.0f;
            for (int r = 0; r < np; ++r) {
                const float px1 = pd[r * 6], py1 = pd[r * 6 + 1], px2 = pd[r * 6 + 2], py2 = pd[r * 6 + 3];
                if (cx < px1 || cx > px2 || cy < py1 || cy > py2) continue;
                const float w = fmaxf(0.f, __fsub_rn(fminf(fx2, px2), fmaxf(fx1, px1)));
                const float h = fmaxf(0.f, __fsub_rn(fminf(fy2, py2), fmaxf(fy1, py1)));
                const float v = farea > 0.f ? __fdiv_rn(__fmul_rn(w, h), farea) : 0.f;
                if (v > bestv) { bestv = v; best = r; }
            }
        }
        pick[f] = best;
    }
    for (int r = tid; r < np; r += kAssocThreads) ident[r] = -1;
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int r = tid; r < np; r += kAssocThreads) {
        for (int f = 0; f < nf; ++f)
            if (pick[f] == r) { ident[r] = face_ids[(size_t)b * face_cap + f]; break; }
    }
    __syncthreads();
    // ordered compaction of the selected rows (np <= person_cap is small: one warp scans with ballots)
    if (tid < 32) {
        int n = 0;
        for (int base = 0; base < np; base += 32) {
            const int r = base + tid;
            const bool sel = r < np && ident[r] >= 0;
            const unsigned m = __ballot_sync(FULL, sel);
            const int pos = n + __popc(m & ((1u << tid) - 1u));
            if (sel && pos < cap) {
                float *o = out_boxes + ((size_t)b * cap + pos) * 4;
                const float x1 = pd[r * 6], y1 = pd[r * 6 + 1];
                o[0] = x1; o[1] = y1; o[2] = __fsub_rn(pd[r * 6 + 2], x1); o[3] = __fsub_rn(pd[r * 6 + 3], y1);
                out_ident[(size_t)b * cap + pos] = ident[r];
                if (out_row) out_row[(size_t)b * cap + pos] = r;
            }
            n += __popc(m);
        }
        if (tid == 0) s_n = n < cap ? n : cap;
    }
    __syncthreads();
    const int n = s_n;
    if (tid == 0) out_count[b] = n;
    for (int i = n + tid; i < cap; i += kAssocThreads) {
        float *o = out_boxes + ((size_t)b * cap + i) * 4;
        o[0] = o[1] = o[2] = o[3] = 0.f;
        out_ident[(size_t)b * cap + i] = -1;
        if (out_row) out_row[(size_t)b * cap + i] = -1;
    }
}

}  // namespace
}  // namespace spp

extern "C" int spp_associate(const float *face_dets, const int *face_count, const int *face_ids, int face_cap,
                             const float *person_dets, const int *person_count, int person_cap, int batch, int cap,
                             float *out_boxes, int *out_ident, int *out_row, int *out_count, spp_stream_t stream) {
    using namespace spp;
    SPP_CHECK_ARG(face_dets && face_count && face_ids && person_dets && person_count && out_boxes && out_ident && out_count,
                  "associate: null pointer");
    SPP_CHECK_ARG(batch >= 0 && face_cap >= 1 && person_cap >= 1 && cap >= 1 && face_cap <= 4096 && person_cap <= 4096,
                  "associate: capacities must be in 1..4096");
    if (batch == 0) return SPP_OK;
    const size_t smem = (size_t)(face_cap + person_cap) * sizeof(int);
    associate_kernel<<<batch, kAssocThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        face_dets, face_count, face_ids, face_cap, person_dets, person_count, person_cap, cap, out_boxes, out_ident, out_row, out_count);
    SPP_CHECK_LAUNCH();
    return SPP_OK;
}
