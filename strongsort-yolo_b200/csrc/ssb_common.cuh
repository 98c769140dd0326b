// Shared declarations for the StrongSORT-on-B200 kernels (sm_100a).
// Data layout in HBM (DESIGN.md "layout"): a slot-indexed struct-of-arrays
// track table, a per-slot appearance gallery ring, and per-frame scratch, all
// carved from ONE caller-owned workspace (include/ssb.h: ssb_create).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ssb.h"

#define SSB_TENTATIVE 1
#define SSB_CONFIRMED 2
#define SSB_DELETED 3
#define SSB_INFTY_COST 1e5
#define SSB_CHI2INV95_4 9.4877

// device-side scalar slots (TrackTable::scalars)
enum { SC_N_TRACKS = 0, SC_NEXT_ID, SC_N_FREE, SC_ERROR, SC_FRAME, SC_COUNT = 16 };
// per-frame counters (FrameScratch::cnt)
enum {
    FC_N_CONF = 0, FC_N_UNCONF, FC_N_UNDET_A, FC_N_CAND_B, FC_N_UNTRK_A_KEEP,
    FC_N_MATCH, FC_N_MATCH_A, FC_N_UNTRK, FC_N_UNDET, FC_N_NEW, FC_ROWS_A, FC_COLS_A,
    FC_ROWS_B, FC_COLS_B, FC_COUNT = 32
};

struct TrackTable {
    double *mean;      // [S][8]
    double *cov;       // [S][64]
    int *track_id, *state, *hits, *age, *tsu, *cls, *last_det;  // [S]
    float *conf;       // [S]
    float *feat;       // [S][D]   EMA-smoothed unit feature (Track.features[-1])
    float *gallery;    // [S][B][D] ring of appended features (metric.samples[id])
    unsigned char *gal_planes;  // [S][hl][D/8][SSB_GAL_ROWS][8] fp16: the same ring as unit vectors * 2^6 split into
                                // hi/lo tensor-core operand planes at append time (appearance.cu); B <= SSB_GAL_ROWS only
    int *gal_count, *gal_head;  // [S]
    int *cls_hist;     // [S][SSB_NCLS] classes of the rows this track has been REPORTED with (the reference's label lines)
    int *dead_count;   // [SSB_NCLS] majority class of every reported track that has since been deleted
    int *order;        // [S] list position -> slot  (== Tracker.tracks order)
    int *order_tmp;    // [S]
    int *free_stack;   // [S]
    int *scalars;      // [SC_COUNT]
};

struct FrameScratch {
    float *det_tlwh;   // [N][4] f32
    float *det_xyah;   // [N][4] f32
    int *det_box;      // [N][4] crop x1,y1,x2,y2
    float *det_conf;   // [N]
    float *det_cls;    // [N]
    float *feats;      // [N][D] raw embeddings
    float *det_norm;      // [N] L2 norm of each raw embedding
    unsigned char *det_planes;  // [hl][D/8][npad][8] fp16: unit embeddings * 2^6 as hi/lo operand planes, npad = 128/256/512
    float *app_cost;   // [S][N] f32 nearest-neighbour cosine distance
    double *cost_a;    // [S][N] gated + clamped stage-A cost
    double *cost_b;    // [S][N] clamped IoU cost
    int *conf_list, *unconf_list;       // [S] list positions
    int *cand_b;       // [S]
    int *untrk_a_keep; // [S] unmatched stage-A tracks with tsu != 1
    int *undet_a;      // [N] unmatched dets after stage A (ordered)
    int *undet;        // [N] final unmatched dets (ordered) -> new ids
    int *untrk;        // [S] final unmatched tracks (positions)
    int *match_trk, *match_det;         // [min(S,N)] positions / det index
    int *col4row, *row4col;             // [max(S,N)] LSAP scratch results
    int *cnt;          // [FC_COUNT]
    double *lsap_ws;   // LSAP global workspace (u, v, spc when not in smem)
};

// the per-detection buffers of one frame; two slots so that the embedding stage of
// frame t+1 can run (on another stream) while frame t is being associated
struct DetSlot {
    float *det_tlwh, *det_xyah;
    int *det_box;
    float *det_conf, *det_cls, *feats, *det_norm;
    unsigned char *det_planes;
};
#define SSB_NCLS 80               // class histogram width of the --count reduction (COCO; larger ids share the last bin)
#define SSB_GAL_ROWS 128          // gallery rows per slot in the operand planes (tensor-core path: nn_budget <= 128)
#define SSB_DET_PLANES_MAX 512    // detections per frame the tensor-core appearance kernel handles (TMEM: 4 x 128 columns)

struct SsbDims {
    int S, N, B, D;
    int n_init, max_age;
    double max_dist, max_iou, mc_lambda, one_minus_lambda;
    float ema_alpha, one_minus_alpha;
};

// OSNet execution (reid.cu)
struct ReidNet;
struct ssb_tracker {
    ssb_config cfg;
    SsbDims dims;
    TrackTable tt;
    FrameScratch fs;
    char *ws_base;
    int64_t ws_bytes;
    // reid
    const float *w_blob;      // folded weights (device, caller-owned)
    int64_t *w_off;           // host: offsets per tensor
    int n_w;
    float *reid_ws;           // activation workspace of slot 0 (device)
    float *reid_ws1;          // second activation workspace: embeddings of two frames may be in flight
    int64_t reid_ws_floats;
    int *boxes_tmp;           // [N][4]
    // tensor-core OSBlocks (reid_tc.cu): hi/lo fp16 operand blob, per-block offsets
    const unsigned char *w_tc;
    int64_t w_tc_off[16];     // 6 OSBlocks (9-tap), 2 transitions, tail, stem, 6 OSBlocks (pointwise + SIMT depthwise)
    int have_tc3;             // sections 10..15 present
    int use_tc;               // 0: fp32 SIMT baseline, 1: tcgen05 OSBlocks with 9 shifted GEMMs per LightConv (reid_tc.cu),
                              // 2: tcgen05 pointwise + fp32 CUDA-core depthwise (reid_tc3.cu)
    // optional per-stage timing of ssb_associate (ssb_profile_enable): events recorded between the kernels
    cudaEvent_t prof_ev[12];
    int prof_on, prof_have;
    int app_simt;             // 1: force the fp32 SIMT appearance kernel (A/B baseline; ssb_appearance_use_tc)
    int *tc_status;           // device int: !=0 -> an mbarrier wait timed out
    DetSlot slot[2];          // slot 0 aliases the buffers in `fs`
    // ssb_update embeds the two halves of a frame's crops on two streams (fork / join by events):
    // the partial last waves of one half's kernels are filled by the other half's
    cudaStream_t side_stream[2][3];   // per detection slot (two frames' embeddings may be in flight) x extra parts
    cudaEvent_t ev_fork[2], ev_join[2][3];
    int side_prio[2];
};

void ssb_set_error(const char *fmt, ...);
// cudaFuncSetAttribute opt-ins (and cached device properties) are PER DEVICE: a guard must fire once per
// (device, kernel), not once per process.  ssb_new_key() hands out a key per call site (use a function-local
// static so each template instantiation gets its own); ssb_first_on_device(key) is true exactly once per
// (current device, key).  Both are thread-safe.
int ssb_new_key();
bool ssb_first_on_device(int key);
int ssb_num_sms();      // SM count of the current device (cached per device)
bool ssb_pdl_enabled(); // programmatic dependent launch between the ReID kernels (SSB_PDL=0 switches it off: A/B)
int ssb_split_parts();      // ssb_update / ssb_reid embed a frame's crops as this many parts on as many streams (SSB_SPLIT=1: unsplit, A/B)
#define SSB_CHECK_CUDA(expr)                                                        \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            ssb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                          __FILE__, __LINE__);                                      \
            return -2;                                                              \
        }                                                                           \
    } while (0)
extern long long g_ssb_launches;   // kernels launched by this library (bench.py: gpu_launches)
#define SSB_CHECK_LAUNCH()                         \
    do {                                           \
        g_ssb_launches++;                          \
        SSB_CHECK_CUDA(cudaGetLastError());        \
    } while (0)

// launchers implemented across the .cu files --------------------------------
int ssb_launch_prep(const SsbDims &d, const float *dets, int n, int h, int w,
                    FrameScratch fs, cudaStream_t st);
int ssb_launch_track_frame(ssb_tracker *t, int slot, int n, int h, int w, const float *feats,
                           double *out, int *counts, int track_hint, cudaStream_t st);
inline FrameScratch ssb_slot_view(const ssb_tracker *t, int slot) {
    FrameScratch f = t->fs;
    const DetSlot &d = t->slot[slot & 1];
    f.det_tlwh = d.det_tlwh; f.det_xyah = d.det_xyah; f.det_box = d.det_box;
    f.det_conf = d.det_conf; f.det_cls = d.det_cls; f.feats = d.feats; f.det_norm = d.det_norm;
    f.det_planes = d.det_planes;
    return f;
}
int ssb_launch_appearance(const float *gallery, const int *gal_count, const int *gal_head,
                          const int *row_slot_list, const int *order, const int *n_rows_dev,
                          int max_rows, int budget, const float *feats, int n_dets, int dim,
                          float *cost, int ld, cudaStream_t st);
// tensor-core appearance cost on the operand planes (appearance.cu); n_dets <= SSB_DET_PLANES_MAX, budget <= SSB_GAL_ROWS
int ssb_launch_appearance_tc(const unsigned char *gal_planes, const int *gal_count, const int *row_pos_list,
                             const int *order, const int *n_rows_dev, int max_rows, int budget,
                             const unsigned char *det_planes, int n_dets, float *cost, int ld, int *status,
                             cudaStream_t st);
inline int ssb_det_npad(int n) { return n <= 128 ? 128 : (n <= 256 ? 256 : 512); }
int ssb_reid_forward(ssb_tracker *t, int slot, const uint8_t *img, int h, int w, int pitch,
                     const int *boxes, int n, float *feats_out, cudaStream_t st);
int ssb_reid_forward_parts(ssb_tracker *t, int slot, const uint8_t *img, int h, int w, int pitch, const int *boxes,
                           int n, float *feats_out, int parts, const cudaStream_t *streams);
int64_t ssb_reid_ws_floats(int max_dets);
int64_t ssb_reid_tc_block_bytes(int b);
int ssb_reid_tc_block(int b, const float *x, float *y, const unsigned char *w, int n, int *status,
                      cudaStream_t st);
int64_t ssb_reid_tc3_block_bytes(int b);
int ssb_reid_tc3_block(int b, const float *x, float *y, const unsigned char *w, int n, int *status,
                       cudaStream_t st);
int64_t ssb_reid_tc_aux_bytes(int which);
// planes != 0: activations are hi/lo fp16 operand planes [crop][hl][C/8][H*W][8] (reid_tc4.cu) instead of float32 NHWC
int ssb_reid_tc_aux(int which, const float *x, float *y, const unsigned char *w, int n, int *status,
                    cudaStream_t st, int planes = 0);
int ssb_reid_tc_stem(const uint8_t *img, int h, int w, int pitch, const int *boxes, const unsigned char *wsec,
                     float *out, int n, int *status, cudaStream_t st, int planes = 0);
int64_t ssb_reid_tc4_block_bytes(int b);
int ssb_reid_tc4_block(int b, const void *x, void *y, const unsigned char *w, int n, int *status, cudaStream_t st);
int ssb_reid_tc4_block_pw(int b, const void *x, void *y, const unsigned char *w, const unsigned char *pw, int n, int *status,
                          cudaStream_t st);
bool ssb_pw_fused();    // transitions fused behind OSBlocks 1 and 3 (SSB_PW_FUSED=0: separate pw_tc launches, A/B)
int ssb_reid_nhwc_to_planes(const float *x, void *y, int n, int hw, int c, cudaStream_t st);
int ssb_reid_planes_to_nhwc(const void *x, float *y, int n, int hw, int c, cudaStream_t st);
