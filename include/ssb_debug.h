/* ssb_debug.h -- A/B baselines and diagnostics of the B200 StrongSORT path.  NOT part of the product
 * boundary (include/ssb.h): these entry points exist only in libssb_dbg.so (built with -DSSB_BASELINES),
 * which the parity tests and tools/ load to compare the product kernels against the fp32 SIMT OSNet, the
 * 9-tap and the round-1 pointwise/depthwise OSBlock kernels, and to read intermediate results. */
#ifndef SSB_DEBUG_H_
#define SSB_DEBUG_H_

#include "ssb.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- fp32 weights of the SIMT baseline network ------------------------------------------------- */
int ssb_reid_num_tensors(void);
/* fills sizes[ssb_reid_num_tensors()] with the element count of each folded
 * tensor in canonical order */
int ssb_reid_tensor_sizes(int64_t *sizes);
int ssb_reid_set_weights(ssb_tracker *t, const float *blob_dev, const int64_t *sizes, int n);

/* ReID implementation used by ssb_reid / ssb_update: 0 = fp32 SIMT network (csrc/reid.cu), 1 = tcgen05
 * OSBlocks with the LightConvs as 9 shifted GEMMs (csrc/reid_tc.cu), 2 = round-1 pointwise/depthwise
 * OSBlocks on float32 NHWC activations with recomputed halos (csrc/reid_tc3.cu), 3 = the product path
 * (operand planes + DSMEM halo exchange, csrc/reid_tc4.cu; the only mode of libssb.so). */
int ssb_reid_use_tc(ssb_tracker *t, int mode);
/* one OSBlock on caller arrays: x float32 NHWC [n][H][W][cin] -> y [n][H][W][cout]; use_tc = mode as above
 * (mode 3 converts to operand planes and back) */
int ssb_reid_block(ssb_tracker *t, int block, const float *x_dev, float *y_dev, int n, int use_tc,
                   ssb_stream_t stream);
/* CTA 0 of every tensor-core OSBlock / stem launch writes clock64() phase stamps into
 * buf_dev (int64[64], [0] = count); NULL switches it off (default) */
int ssb_reid_tc_debug(void *buf_dev);

/* per-frame association trace of the LAST ssb_update (device buffers, valid
 * until the next update): appearance cost [nA_rows, n] float32 is NOT kept;
 * the gated/clamped stage-A cost matrix float64 [rows_a, n] and the stage-B
 * matrix [rows_b, cols_b] are.  dims_out int32 [4] = rows_a, cols_a, rows_b,
 * cols_b (device).                                                          */
int ssb_debug_cost_ptrs(ssb_tracker *t, const double **cost_a_dev, const double **cost_b_dev,
                        const int32_t **dims_dev);

/* ---- diagnostic: one tcgen05 GEMM tile in the operand layout of the ReID kernels
 * D[128][n] (f32) = A[shift:shift+128][k] (f16) * B[n][k]^T (f16); status!=0: timeout */
int ssb_tc_probe(const void *a_dev, int a_rows, int shift, const void *b_dev, int n, int k,
                 float *d_dev, int32_t *status_dev, ssb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SSB_DEBUG_H_ */
