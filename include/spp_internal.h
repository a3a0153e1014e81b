/* spp_internal.h — test hooks exported by libspp.so that are NOT part of the drop-in surface. */
#ifndef SPP_INTERNAL_H_
#define SPP_INTERNAL_H_
#include "spp.h"
#ifdef __cplusplus
extern "C" {
#endif
/* Same contract as spp_match_top1, but the candidate search runs as a CUDA-core fp32 kernel instead of
 * the tcgen05 GEMM.  Used by the GPU tests as a device-side cross-check; the product never calls it. */
int spp_debug_match_top1_simt(const float *emb, const uint16_t *gallery, int m, int n, int dim, float threshold,
                              int id_offset, int *out_id, float *out_sim, unsigned long long *out_key, void *workspace,
                              size_t workspace_bytes, spp_stream_t stream);
#ifdef __cplusplus
}
#endif
#endif
