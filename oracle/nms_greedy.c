/* Oracle (TEST INFRASTRUCTURE, not product code): plain-C restatement of the greedy suppression
 * loop the reference reaches through torchvision.ops.nms at training/yolopt/util.py:162
 * (torchvision 0.26 csrc/ops/cpu/nms_kernel.cpp — third-party, restated, not copied).
 *
 * boxes  [n,4] fp32 (x1,y1,x2,y2), already class-offset (util.py:160-161)
 * order  [n]   int32, candidate indices sorted by score descending (stable)
 * keep   [n]   int32 out; returns the number of kept boxes.
 * Suppress j iff inter/(area_i+area_j-inter) > thr (strict), fp32, no FMA contraction
 * (build with -ffp-contract=off).
 */
#include <stdlib.h>

static float fmaxf_(float a, float b) { return a > b ? a : b; }
static float fminf_(float a, float b) { return a < b ? a : b; }

int oracle_nms_greedy(const float *boxes, const int *order, int n, float thr, int *keep)
{
    unsigned char *dead = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
    float *area = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    int kept = 0;
    for (int i = 0; i < n; ++i) {
        const float *b = boxes + 4 * (size_t)i;
        area[i] = (b[2] - b[0]) * (b[3] - b[1]);
    }
    for (int a = 0; a < n; ++a) {
        int i = order[a];
        if (dead[i]) continue;
        keep[kept++] = i;
        const float *bi = boxes + 4 * (size_t)i;
        for (int c = a + 1; c < n; ++c) {
            int j = order[c];
            if (dead[j]) continue;
            const float *bj = boxes + 4 * (size_t)j;
            float w = fmaxf_(0.0f, fminf_(bi[2], bj[2]) - fmaxf_(bi[0], bj[0]));
            float h = fmaxf_(0.0f, fminf_(bi[3], bj[3]) - fmaxf_(bi[1], bj[1]));
            float inter = w * h;
            float ovr = inter / (area[i] + area[j] - inter);
            if (ovr > thr) dead[j] = 1;
        }
    }
    free(dead);
    free(area);
    return kept;
}
