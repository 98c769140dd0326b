/* ssb.h -- C-ABI of the B200-native StrongSORT per-frame tracking path.
 *
 * Boundary: the reference's tracker seam is the single call
 *     results = model.track(image, ..., persist=True, tracker="botsort.yaml")
 * at /root/reference/yolo_multi_model.py:41, behind which a StrongSORT plug-in
 * exposes  StrongSORT.update(dets[N,6], img[H,W,3]) -> rows[M,7]
 * (the upstream strong_sort/ package is absent from the
 * snapshot).  Every entry point below is what a Python/ctypes (or any FFI)
 * binding of that seam would bind; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain C types only; all *_dev pointers are device pointers into buffers
 *     the CALLER owns (torch CUDA tensors in the Python host); the library
 *     never allocates device memory: ssb_create() carves the caller-provided
 *     workspace of ssb_workspace_bytes() bytes.
 *   - all work is enqueued on the caller's stream (cudaStream_t passed as
 *     void*); no entry point synchronises; graph-capturable.  (ssb_update and ssb_reid fork
 *     half of a frame's embedding onto an internal non-blocking stream and joins
 *     it back with events before the association -- ordering on the caller's
 *     stream is unchanged.)
 *   - return 0 on success, negative on error; ssb_last_error() returns a
 *     thread-local message.  No exceptions cross the ABI.
 *   - one handle per video stream; a handle is not thread-safe, distinct
 *     handles are independent.
 */
#ifndef SSB_H_
#define SSB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssb_tracker ssb_tracker;
typedef void *ssb_stream_t; /* cudaStream_t */

/* Constructor knobs == strong_sort.yaml of upstream. */
typedef struct ssb_config {
    int32_t max_tracks;      /* track slots in the device table (>= live tracks) */
    int32_t max_dets;        /* max detections per frame                        */
    int32_t nn_budget;       /* NN_BUDGET   (100)                               */
    int32_t feat_dim;        /* 512                                             */
    int32_t n_init;          /* N_INIT      (3)                                 */
    int32_t max_age;         /* MAX_AGE     (30)                                */
    double max_dist;         /* MAX_DIST    (0.2)  appearance threshold         */
    double max_iou_distance; /* MAX_IOU_DISTANCE (0.7)                          */
    double mc_lambda;        /* MC_LAMBDA   (0.995)                             */
    double ema_alpha;        /* EMA_ALPHA   (0.9)                               */
} ssb_config;

/* counts written by ssb_update() to counts_dev[8] */
enum {
    SSB_CNT_OUT_ROWS = 0,   /* M rows written to out_dev                */
    SSB_CNT_TRACKS = 1,     /* live tracks after the update             */
    SSB_CNT_CONFIRMED = 2,  /* confirmed tracks after the update        */
    SSB_CNT_NEXT_ID = 3,    /* next track id to be issued               */
    SSB_CNT_MATCHES_A = 4,  /* matches of the appearance stage          */
    SSB_CNT_MATCHES_B = 5,  /* matches of the IoU stage                 */
    SSB_CNT_NEW = 6,        /* tracks initiated this frame              */
    SSB_CNT_ERROR = 7,      /* bit 0: table overflow (tracks/dets dropped);
                               bit 1: a tensor-core barrier wait of the ReID
                               kernels timed out (embeddings invalid)   */
    SSB_CNT_N = 8
};
#define SSB_OUT_COLS 8 /* x1,y1,x2,y2,track_id,cls,conf,det_index(-1 if none) */

int ssb_version(void);
/* kernels launched by this library since load (bench.py reports the delta) */
int64_t ssb_launch_count(void);
const char *ssb_last_error(void);
void ssb_default_config(ssb_config *cfg);

/* ---- lifecycle ---------------------------------------------------------- */
int64_t ssb_workspace_bytes(const ssb_config *cfg);
int ssb_create(const ssb_config *cfg, void *workspace_dev, int64_t workspace_bytes,
               ssb_tracker **out);
int ssb_destroy(ssb_tracker *t);
/* forget all tracks, ids restart at 1 */
int ssb_reset(ssb_tracker *t, ssb_stream_t stream);

/* StrongSORT.increment_ages() of upstream (SURVEY.md A.2 caller side: its stream loop calls it on
 * frames WITHOUT detections instead of update()): every track  age += 1, time_since_update += 1,
 * mark_missed(); no Kalman predict.  Tracks deleted here leave the list at the next ssb_update. */
int ssb_increment_ages(ssb_tracker *t, ssb_stream_t stream);

/* The reference's --count reduction (yolo_multi_model.py:284-300: per track id the most frequent class of
 * its label lines -- smallest class on ties --, then the number of ids per class) over the device track
 * table: out_dev int32 [80]; ids of deleted tracks keep counting.  Classes >= 80 share the last bin. */
int ssb_class_counts(ssb_tracker *t, int32_t *out_dev, ssb_stream_t stream);

/* ---- ReID weights (OSNet-x0.25, BN folded by the host; see weights.py) ---
 * fp16 hi/lo operand blob built by weights.pack_tc(); block_offsets[16] are byte offsets of the sections,
 * each >= ssb_reid_tc_weight_bytes(i) bytes and 128-byte aligned: 0..5 the OSBlocks in the 9-tap layout
 * (used by the A/B baseline of libssb_dbg.so only), 6..7 transition layers, 8 tail, 9 stem, 10..15 the
 * OSBlocks with pointwise-GEMM + fp32-depthwise LightConvs (csrc/reid_tc4.cu). */
int64_t ssb_reid_tc_weight_bytes(int section);
int ssb_reid_set_weights_tc(ssb_tracker *t, const void *blob_dev, const int64_t *block_offsets,
                            int n_blocks);
/* copies the tensor-core path's device status word (0 = ok) to the host; synchronises */
int ssb_reid_tc_status(ssb_tracker *t, int32_t *status_host, ssb_stream_t stream);

/* ---- the per-frame hot path: StrongSORT.update(dets, img) ---------------- */
/* dets_dev  : float32 [n,6] x1,y1,x2,y2,conf,cls
 * img_dev   : uint8 BGR, h rows of `pitch` bytes (pitch >= 3*w)
 * feats_dev : NULL -> embeddings are computed by the built-in OSNet from img;
 *             else float32 [n,feat_dim] caller-supplied embeddings (tests)
 * out_dev   : float64 [max_tracks, SSB_OUT_COLS]
 * counts_dev: int32 [SSB_CNT_N]
 * track_hint: upper bound on live tracks entering this frame (the
 *             SSB_CNT_TRACKS of the previous frame), or -1 for max_tracks;
 *             only sizes launch grids, never changes results.            */
int ssb_update(ssb_tracker *t, const float *dets_dev, int n, const uint8_t *img_dev,
               int h, int w, int pitch, const float *feats_dev, double *out_dev,
               int32_t *counts_dev, int track_hint, ssb_stream_t stream);

/* The same step split in its two stages, for callers that overlap them on two streams:
 * ssb_embed (detection prep + OSNet embeddings into slot 0/1; independent of the track
 * table) for frame t+1 may run while ssb_associate (everything else) runs for frame t.
 * ssb_update(...) == ssb_embed(slot 0) + ssb_associate(slot 0).  (Frames of >= 48 crops are embedded in up to
 * three parts -- the others on internal streams of the slot, of the caller's stream priority, joined back before the
 * call returns control of `stream` -- inside the slot's own workspace.)  The caller orders the stages with
 * events; img_dev == NULL in ssb_embed skips the OSNet (the caller then passes feats_dev to ssb_associate). */
int ssb_embed(ssb_tracker *t, int slot, const float *dets_dev, int n, const uint8_t *img_dev,
              int h, int w, int pitch, ssb_stream_t stream);
int ssb_associate(ssb_tracker *t, int slot, int n, int h, int w, const float *feats_dev,
                  double *out_dev, int32_t *counts_dev, int track_hint, ssb_stream_t stream);

/* ---- stage entry points (parity tests call these one by one) ------------- */
/* OSNet embeddings of n crops: boxes_dev int32 [n,4] x1,y1,x2,y2 with crop =
 * img[y1:y2, x1:x2]; out float32 [n,512].                                   */
int ssb_reid(ssb_tracker *t, const uint8_t *img_dev, int h, int w, int pitch,
             const int32_t *boxes_dev, int n, float *feats_out_dev, ssb_stream_t stream);
/* crop boxes exactly as _get_features/_xywh_to_xyxy derive them from dets    */
int ssb_crop_boxes(const float *dets_dev, int n, int h, int w, int32_t *boxes_out_dev,
                   ssb_stream_t stream);
/* batched Kalman filter on caller arrays: mean [n,8], cov [n,8,8] float64    */
int ssb_kf_predict(double *mean_dev, double *cov_dev, int n, ssb_stream_t stream);
int ssb_kf_update(double *mean_dev, double *cov_dev, const float *xyah_dev,
                  const float *conf_dev, int n, ssb_stream_t stream);
/* squared Mahalanobis (4 dof): maha_out [n_tracks, n_meas]                   */
int ssb_kf_gating(const double *mean_dev, const double *cov_dev, int n_tracks,
                  const float *xyah_dev, int n_meas, double *maha_out_dev,
                  ssb_stream_t stream);
/* appearance cost: gallery float32 [n_tracks, budget, dim] with counts[n_tracks]
 * valid samples each; feats [n_dets, dim]; cost_out float32 [n_tracks, n_dets] */
int ssb_appearance_cost(const float *gallery_dev, const int32_t *counts_dev,
                        int n_tracks, int budget, const float *feats_dev, int n_dets,
                        int dim, float *cost_out_dev, ssb_stream_t stream);
/* the same cost on the tensor cores (the tracker's default path, csrc/appearance.cu): the arrays are first
 * re-laid as fp16 hi/lo operand planes into scratch_dev (>= ssb_appearance_tc_scratch_bytes(n_tracks) bytes);
 * dim == 512, budget <= 128, n_dets <= 512; status_dev int32[1] receives a non-zero code on a barrier timeout */
int64_t ssb_appearance_tc_scratch_bytes(int n_tracks);
int ssb_appearance_cost_tc(const float *gallery_dev, const int32_t *counts_dev, int n_tracks, int budget,
                           const float *feats_dev, int n_dets, int dim, float *cost_out_dev,
                           void *scratch_dev, int32_t *status_dev, ssb_stream_t stream);
/* A/B switch of the tracker's appearance stage: 1 (default) tensor-core kernel on the operand planes the
 * tracker maintains, 0 the fp32 SIMT kernel on the float32 gallery */
int ssb_appearance_use_tc(ssb_tracker *t, int enable);
/* 1 - IoU: track tlwh float64 [n_tracks,4], det tlwh float32 [n_dets,4]       */
int ssb_iou_cost(const double *track_tlwh_dev, int n_tracks, const float *det_tlwh_dev,
                 int n_dets, double *cost_out_dev, ssb_stream_t stream);
/* rectangular LSAP with scipy's tie-breaks: cost float64 [nr,nc] row-major;
 * col4row_out int32 [nr] (-1 = unassigned), row4col_out int32 [nc]          */
int ssb_lsap(const double *cost_dev, int nr, int nc, int32_t *col4row_out_dev,
             int32_t *row4col_out_dev, ssb_stream_t stream);
/* YOLO post-process (thresholds of yolo_multi_model.py:18-21):
 * pred float32 [4+nc(+extra), A] (xywh, class scores, extra channels);
 * out float32 [max_det, 6+extra] x1,y1,x2,y2,conf,cls,extra...; count int32[1];
 * scratch_dev >= ssb_nms_scratch_bytes(A) bytes                              */
int64_t ssb_nms_scratch_bytes(int num_anchors);
int ssb_yolo_nms(const float *pred_dev, int num_classes, int num_extra, int num_anchors,
                 float conf_thres, float iou_thres, int max_det, int agnostic,
                 float *out_dev, int32_t *count_dev, void *scratch_dev,
                 ssb_stream_t stream);

/* ultralytics scale_boxes: the first count_dev[0] rows of an ssb_yolo_nms output (cols floats per row, box in
 * columns 0..3, network-input pixels) back to the original w0 x h0 frame: subtract the letterbox padding,
 * divide by the gain, clip.  In place. */
int ssb_yolo_scale_boxes(float *rows_dev, int cols, const int32_t *count_dev, int max_det, float gain,
                         float pad_x, float pad_y, int w0, int h0, ssb_stream_t stream);

/* The whole YOLOv8 detector post-process in one call (5 launches): decode with the best class per anchor fused,
 * confidence filter + class-aware NMS, and -- when gain > 0 -- scale_boxes fused into the gather.  Same results as
 * ssb_yolo_decode_v8 -> ssb_yolo_nms -> ssb_yolo_scale_boxes.  pred_scratch_dev: float32 [4 + nc + 3*kpts, A]. */
int ssb_yolo_postprocess_v8(const float *raw_dev, int num_classes, int num_kpts, int in_h, int in_w,
                            float conf_thres, float iou_thres, int max_det, int agnostic,
                            float gain, float pad_x, float pad_y, int w0, int h0,
                            float *pred_scratch_dev, float *out_dev, int32_t *count_dev, void *scratch_dev,
                            ssb_stream_t stream);

/* YOLOv8 detect / pose head decode: raw float32 [4*16 + nc + 3*kpts, A] (DFL bins,
 * class logits, keypoint x,y,vis) for a network input of in_h x in_w (multiples of 32; A =
 * ssb_yolo_num_anchors, stride-8/16/32 levels concatenated) -> pred float32 [4 + nc + 3*kpts, A],
 * the layout ssb_yolo_nms consumes (extra channels = 3*kpts).                                    */
int ssb_yolo_num_anchors(int in_h, int in_w);
int ssb_yolo_decode_v8(const float *raw_dev, int num_classes, int num_kpts, int in_h, int in_w,
                       float *pred_out_dev, ssb_stream_t stream);
/* YOLOv5 / v7 head (BASELINE config C1, the detectors upstream StrongSORT-YOLO pairs with the tracker):
 * raw float32 [3 * ssb_yolo_num_anchors(in_h, in_w), 5 + nc] logits (level by level, [na][gy][gx] inside a
 * level), anchors_px_dev float32 [3 levels][3][w, h] in pixels -> pred [4 + nc, A] with score = obj * cls,
 * zero where obj <= conf_thres (yolov5 non_max_suppression's candidate rule), ready for ssb_yolo_nms.   */
int ssb_yolo_decode_v5(const float *raw_dev, int num_classes, int in_h, int in_w, float conf_thres,
                       const float *anchors_px_dev, float *pred_out_dev, ssb_stream_t stream);

/* camera-motion compensation, tracker side (upstream Track.camera_update after its ECC call,
 * SURVEY.md A.9): warp2x3_host = row-major 2x3 matrix (host doubles, e.g. cv2.findTransformECC's
 * result with the translation scaled back to full resolution), applied to the tl/br corners of
 * every live track; rewrites mean[:4].  Called between frames, outside ssb_update.             */
int ssb_camera_update(ssb_tracker *t, const double *warp2x3_host, ssb_stream_t stream);

/* ---- optional cross-stream ReID gallery (config C5; read-only, see csrc/gallery.cu) ----------
 * export: the confirmed tracks of this stream in list order -> feat float32 [t_max, feat_dim] (unit
 * EMA vectors, zero rows past the count), ids int32 [2 * t_max] (first t_max: track ids, -1 past the
 * count; second half: scratch), count int32 [1].  The host all-gathers feat / ids over NCCL.
 * cross_match: local [t_max, dim] against the gathered [n_ranks, t_max, dim]: per local track the
 * nearest track of another rank in cosine distance -> rank / id (-1 when farther than max_dist
 * or no foreign track) and the distance.                                                        */
int ssb_gallery_export(ssb_tracker *t, int t_max, float *feat_out_dev, int32_t *ids_out_dev,
                       int32_t *count_out_dev, ssb_stream_t stream);
int ssb_gallery_cross_match(const float *local_feat_dev, const int32_t *local_ids_dev,
                            const float *all_feat_dev, const int32_t *all_ids_dev, int n_ranks,
                            int self_rank, int t_max, int dim, float max_dist,
                            int32_t *match_rank_out_dev, int32_t *match_id_out_dev,
                            float *match_dist_out_dev, ssb_stream_t stream);

/* per-stage timing of ssb_associate (bench.py's stage split): CUDA events between its kernels.
 * ms_out9: prep (norms, KF predict, lists) | appearance | gate | LSAP A + lists | IoU | LSAP B + lists |
 * KF / EMA update | bookkeeping | gallery append.  ssb_profile_read synchronises on the last event.   */
int ssb_profile_enable(ssb_tracker *t, int on);
int ssb_profile_read(ssb_tracker *t, float *ms_out9);

/* the same match on the packed exchange layout: per rank [t_max * dim float32 | t_max int32 ids], so one
 * collective moves a stream's whole export (dist.SharedGallery) */
int ssb_gallery_cross_match_packed(const void *all_packed_dev, int n_ranks, int self_rank, int t_max, int dim,
                                   float max_dist, int32_t *match_rank_out_dev, int32_t *match_id_out_dev,
                                   float *match_dist_out_dev, ssb_stream_t stream);

/* the same match WITHOUT a collective: peer_ptrs_dev is a device array of n_ranks pointers, entry r = rank r's
 * packed export in peer-accessible memory (torch symmetric-memory buffers: the kernel pulls the other streams'
 * rows over NVLink itself, each foreign row crossing once); best_scratch_dev t_max uint64; dim == 512 */
int ssb_gallery_peer_match(const void *peer_ptrs_dev, int n_ranks, int self_rank, int t_max, int dim, float max_dist,
                           void *best_scratch_dev, int32_t *match_rank_out_dev, int32_t *match_id_out_dev,
                           float *match_dist_out_dev, ssb_stream_t stream);

/* ---- introspection for tests: copy the live track table (list order) ----- */
/* any pointer may be NULL.  ids/state/hits/age/tsu/gallery_len int32 [T];
 * mean float64 [T,8]; cov [T,8,8]; feat float32 [T,dim]                     */
int ssb_export_tracks(ssb_tracker *t, int32_t *ids, int32_t *state, int32_t *hits,
                      int32_t *age, int32_t *tsu, int32_t *gallery_len, double *mean,
                      double *cov, float *feat, ssb_stream_t stream);
#ifdef __cplusplus
}
#endif
#endif /* SSB_H_ */
