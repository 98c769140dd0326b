// OSNet-x0.25 ReID forward of the product path: orchestration, workspace and tensor-core weights.
//
//   stem (crop + resize + normalise + conv7x7/2 + ReLU + maxpool, reid_tc.cu) -> 6 OSBlocks (reid_tc4.cu)
//   with a transition (conv1x1 + ReLU + avgpool) after blocks 1 and 3 -> tail (conv5 + GAP + fc + ReLU):
//   11 launches, every activation between them a pair of fp16 hi/lo operand planes
//   [crop][hl][C/8][H*W][8] (4 bytes per element: the ping-pong buffers hold 4 * 64 * 2048 bytes per crop).
//
// The fp32 SIMT network (reid.cu), the 9-tap and the round-1 pointwise/depthwise OSBlocks (reid_tc.cu,
// reid_tc3.cu) are A/B baselines: compiled only with -DSSB_BASELINES into libssb_dbg.so (include/ssb_debug.h).
#include "ssb_common.cuh"

#define REID_BIG 131072          // floats (= 4-byte words) per crop of the largest activation: 64 x 32 x 64
#define REID_MID 32768
int64_t ssb_reid_ws_floats(int max_dets) {
#ifdef SSB_BASELINES             // + the SIMT network's intermediates (DS, X1, T0, T1, PW, X2, gate)
    return (int64_t)max_dets * (3 * REID_BIG + 5 * REID_MID + 1024) + 1024;
#else
    return (int64_t)max_dets * (2 * REID_BIG) + 1024;
#endif
}

#ifdef SSB_BASELINES
int ssb_reid_forward_baseline(ssb_tracker *t, int slot, const uint8_t *img, int h, int w, int pitch, const int *boxes,
                              int n, float *feats_out, cudaStream_t st);
#endif

// mode 3 (default): every activation between kernels is a pair of fp16 operand planes (reid_tc4.cu):
// stem -> K0 -> K1 -> transition -> K2 -> K3 -> transition -> K4 -> K5 -> tail, 11 launches
// ws: 2 * n * REID_BIG 4-byte words of activation workspace (ping-pong)
static int reid_forward_planes(ssb_tracker *t, float *ws, const uint8_t *img, int h, int w, int pitch, const int *boxes,
                               int n, float *feats_out, cudaStream_t st) {
    float *A = ws;
    float *Bf = A + (size_t)n * REID_BIG;
    const unsigned char *W = t->w_tc;
    int rc = ssb_reid_tc_stem(img, h, w, pitch, boxes, W + t->w_tc_off[9], A, n, t->tc_status, st, 1);
    if (rc) return rc;
    float *cur = A, *nxt = Bf;
    const bool fused = ssb_pw_fused();
    for (int b = 0; b < 6; b++) {
        if (fused && (b == 1 || b == 3)) {          // OSBlock + its stage's transition layer in one launch (9 launches in all)
            rc = ssb_reid_tc4_block_pw(b, cur, nxt, W + t->w_tc_off[10 + b], W + t->w_tc_off[6 + (b == 1 ? 0 : 1)], n,
                                       t->tc_status, st);
            if (rc) return rc;
            { float *tmp = cur; cur = nxt; nxt = tmp; }
            continue;
        }
        rc = ssb_reid_tc4_block(b, cur, nxt, W + t->w_tc_off[10 + b], n, t->tc_status, st);
        if (rc) return rc;
        { float *tmp = cur; cur = nxt; nxt = tmp; }
        if (b == 1 || b == 3) {
            const int a = b == 1 ? 0 : 1;
            rc = ssb_reid_tc_aux(a, cur, nxt, W + t->w_tc_off[6 + a], n, t->tc_status, st, 1);
            if (rc) return rc;
            { float *tmp = cur; cur = nxt; nxt = tmp; }
        }
    }
    return ssb_reid_tc_aux(2, cur, feats_out, W + t->w_tc_off[8], n, t->tc_status, st, 1);
}

int ssb_reid_forward(ssb_tracker *t, int slot, const uint8_t *img, int h, int w, int pitch, const int *boxes,
                     int n, float *feats_out, cudaStream_t st) {
    if (n <= 0) return 0;
    if (!t->w_tc || !t->have_tc3) { ssb_set_error("ReID weights not set (ssb_reid_set_weights_tc)"); return -1; }
#ifdef SSB_BASELINES
    if (t->use_tc != 3) return ssb_reid_forward_baseline(t, slot, img, h, w, pitch, boxes, n, feats_out, st);
#endif
    return reid_forward_planes(t, (slot & 1) ? t->reid_ws1 : t->reid_ws, img, h, w, pitch, boxes, n, feats_out, st);
}

// The product path on the two halves of a slot's workspace: the crops [0, n0) on `st`, [n0, n) on `side` (the caller
// forks / joins the streams).  Every ReID kernel ends in a partial wave; two independent half-frames in flight fill
// each other's tails (measured: 489 vs 558 us per 100 crops).  Returns 1 when the split does not apply (baseline
// modes of the debug library: the caller falls back to ssb_reid_forward).
int ssb_reid_forward_parts(ssb_tracker *t, int slot, const uint8_t *img, int h, int w, int pitch, const int *boxes,
                           int n, float *feats_out, int parts, const cudaStream_t *streams) {
    if (!t->w_tc || !t->have_tc3) { ssb_set_error("ReID weights not set (ssb_reid_set_weights_tc)"); return -1; }
    if (t->use_tc != 3) return 1;
    float *ws = (slot & 1) ? t->reid_ws1 : t->reid_ws;
    for (int p = 0, off = 0; p < parts; p++) {
        const int np = (n - off + (parts - p) - 1) / (parts - p);
        int rc = reid_forward_planes(t, ws + (size_t)2 * off * REID_BIG, img, h, w, pitch, boxes + 4 * off, np,
                                     feats_out + (size_t)off * t->dims.D, streams[p]);
        if (rc) return rc;
        off += np;
    }
    return 0;
}

// ---- tensor-core weights -----------------------------------------------------
extern "C" int64_t ssb_reid_tc_weight_bytes(int section) {
    if (section >= 10) return ssb_reid_tc4_block_bytes(section - 10);
    return section < 6 ? ssb_reid_tc_block_bytes(section) : ssb_reid_tc_aux_bytes(section - 6);
}

extern "C" int ssb_reid_set_weights_tc(ssb_tracker *t, const void *blob_dev, const int64_t *block_offsets,
                                       int n_blocks) {
    if (!t || !blob_dev || !block_offsets) { ssb_set_error("null argument"); return -1; }
    if (n_blocks != 10 && n_blocks != 16) {
        ssb_set_error("expected 10 sections (6 OSBlocks, 2 transitions, tail, stem) or 16 (+ 6 pointwise/depthwise OSBlocks), got %d", n_blocks);
        return -1;
    }
    for (int b = 0; b < n_blocks; b++) {
        if (block_offsets[b] % 128 != 0) { ssb_set_error("section %d offset not 128-byte aligned", b); return -1; }
        const int64_t need = ssb_reid_tc_weight_bytes(b);
        if (b < n_blocks - 1 && block_offsets[b + 1] - block_offsets[b] < need) {
            ssb_set_error("section %d too small", b);
            return -1;
        }
        t->w_tc_off[b] = block_offsets[b];
    }
    if (((uintptr_t)blob_dev & 127) != 0) { ssb_set_error("tc blob must be 128-byte aligned"); return -1; }
    t->w_tc = (const unsigned char *)blob_dev;
    SSB_CHECK_CUDA(cudaMemset(t->tc_status, 0, 64 * sizeof(int)));      // the workspace arrives uninitialised
    t->have_tc3 = n_blocks == 16;
    t->use_tc = t->have_tc3 ? 3 : 1;
    return 0;
}

extern "C" int ssb_reid_tc_status(ssb_tracker *t, int32_t *status_host, ssb_stream_t stream) {
    if (!t || !status_host) { ssb_set_error("null argument"); return -1; }
    SSB_CHECK_CUDA(cudaMemcpyAsync(status_host, t->tc_status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    SSB_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

