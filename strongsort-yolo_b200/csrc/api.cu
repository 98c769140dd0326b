// extern "C" boundary of libssb.so (include/ssb.h): handle lifecycle, workspace
// carving and the per-frame orchestration of StrongSORT.update(dets, img)
// (seam: /root/reference/yolo_multi_model.py:41).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ssb_common.cuh"

int ssb_launch_reset(ssb_tracker *t, cudaStream_t st);
int ssb_launch_export(ssb_tracker *t, int *ids, int *state, int *hits, int *age, int *tsu, int *gal,
                      double *mean, double *cov, float *feat, cudaStream_t st);

static thread_local char g_err[512] = "";

void ssb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *ssb_last_error(void) { return g_err; }

// ---- per-device one-time guards (see ssb_common.cuh) ---------------------------------------
#include <mutex>
static std::mutex g_dev_mu;
static int g_next_key = 0;
static unsigned char g_dev_done[64][256];
static int g_dev_sms[64];
int ssb_new_key() {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    return g_next_key < 255 ? g_next_key++ : 255;
}
bool ssb_first_on_device(int key) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || key < 0 || key >= 255) return true;      // out of table: always (re)apply
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (g_dev_done[dev][key]) return false;
    g_dev_done[dev][key] = 1;
    return true;
}
static int env_flag(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) != 0 : dflt;
}
bool ssb_pdl_enabled() { static const int on = env_flag("SSB_PDL", 1); return on != 0; }
bool ssb_pw_fused() { static const int on = env_flag("SSB_PW_FUSED", 1); return on != 0; }
int ssb_split_parts() {        // SSB_SPLIT = parts a frame's crops are embedded in (0 / 1: unsplit), default 3
    static const int p = [] { const char *v = getenv("SSB_SPLIT"); return v && *v ? atoi(v) : 3; }();
    return p < 1 ? 1 : p > 4 ? 4 : p;
}
int ssb_num_sms() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && g_dev_sms[dev] > 0) return g_dev_sms[dev];
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    if (dev >= 0 && dev < 64) g_dev_sms[dev] = n;
    return n;
}
long long g_ssb_launches = 0;
extern "C" int ssb_version(void) { return 100; }
extern "C" int64_t ssb_launch_count(void) { return (int64_t)g_ssb_launches; }

extern "C" void ssb_default_config(ssb_config *c) {
    c->max_tracks = 1024;
    c->max_dets = 512;
    c->nn_budget = 100;
    c->feat_dim = 512;
    c->n_init = 3;
    c->max_age = 30;
    c->max_dist = 0.2;
    c->max_iou_distance = 0.7;
    c->mc_lambda = 0.995;
    c->ema_alpha = 0.9;
}

static int check_cfg(const ssb_config *c) {
    if (!c) { ssb_set_error("null config"); return -1; }
    if (c->max_tracks < 1 || c->max_tracks > 4096) { ssb_set_error("max_tracks must be in [1,4096]"); return -1; }
    if (c->max_dets < 1 || c->max_dets > 4096) { ssb_set_error("max_dets must be in [1,4096]"); return -1; }
    if (c->nn_budget < 1 || c->nn_budget > 1024) { ssb_set_error("nn_budget must be in [1,1024]"); return -1; }
    if (c->feat_dim != 512) { ssb_set_error("feat_dim must be 512 (OSNet)"); return -1; }
    if (c->n_init < 1 || c->max_age < 1) { ssb_set_error("n_init/max_age must be >= 1"); return -1; }
    return 0;
}

// bump allocator over the caller's workspace; pass base==nullptr to size it
struct Carver {
    char *base;
    size_t off;
    template <typename T>
    T *take(size_t n) {
        off = (off + 255) & ~(size_t)255;
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

static size_t carve(const ssb_config *c, char *base, TrackTable *tt, FrameScratch *fs,
                    float **reid_ws, int64_t *reid_floats, int **boxes_tmp, int **tc_status = nullptr,
                    DetSlot *slot1 = nullptr, float **reid_ws1 = nullptr) {
    const size_t S = c->max_tracks, N = c->max_dets, B = c->nn_budget, D = c->feat_dim;
    const size_t L = S > N ? S : N;
    Carver k{base, 0};
    TrackTable t;
    t.mean = k.take<double>(S * 8);
    t.cov = k.take<double>(S * 64);
    t.track_id = k.take<int>(S); t.state = k.take<int>(S); t.hits = k.take<int>(S);
    t.age = k.take<int>(S); t.tsu = k.take<int>(S); t.cls = k.take<int>(S);
    t.last_det = k.take<int>(S);
    t.conf = k.take<float>(S);
    t.feat = k.take<float>(S * D);
    t.gallery = k.take<float>(S * B * D);
    t.gal_planes = k.take<unsigned char>(S * (size_t)(2 * (D / 8) * SSB_GAL_ROWS * 16));
    t.gal_count = k.take<int>(S); t.gal_head = k.take<int>(S);
    t.cls_hist = k.take<int>(S * SSB_NCLS); t.dead_count = k.take<int>(SSB_NCLS);
    t.order = k.take<int>(S); t.order_tmp = k.take<int>(S); t.free_stack = k.take<int>(S);
    t.scalars = k.take<int>(SC_COUNT);
    FrameScratch f;
    f.det_tlwh = k.take<float>(N * 4); f.det_xyah = k.take<float>(N * 4);
    f.det_box = k.take<int>(N * 4);
    f.det_conf = k.take<float>(N); f.det_cls = k.take<float>(N);
    f.feats = k.take<float>(N * D);
    f.det_norm = k.take<float>(N);
    f.det_planes = k.take<unsigned char>((size_t)2 * (D / 8) * SSB_DET_PLANES_MAX * 16);
    f.app_cost = k.take<float>(S * N);
    f.cost_a = k.take<double>(S * N);
    f.cost_b = k.take<double>(S * N);
    f.conf_list = k.take<int>(S); f.unconf_list = k.take<int>(S); f.cand_b = k.take<int>(S);
    f.untrk_a_keep = k.take<int>(S);
    f.undet_a = k.take<int>(N); f.undet = k.take<int>(N);
    f.untrk = k.take<int>(S);
    f.match_trk = k.take<int>(L); f.match_det = k.take<int>(L);
    f.col4row = k.take<int>(L); f.row4col = k.take<int>(L);
    f.cnt = k.take<int>(FC_COUNT);
    f.lsap_ws = k.take<double>(16);
    const int64_t rf = ssb_reid_ws_floats((int)N);
    float *rw = k.take<float>((size_t)rf);
    float *rw1 = k.take<float>((size_t)rf);
    if (reid_ws1) *reid_ws1 = rw1;
    int *bt = k.take<int>(N * 4);
    int *tcs = k.take<int>(64);
    if (tc_status) *tc_status = tcs;
    DetSlot s1;
    s1.det_tlwh = k.take<float>(N * 4); s1.det_xyah = k.take<float>(N * 4);
    s1.det_box = k.take<int>(N * 4);
    s1.det_conf = k.take<float>(N); s1.det_cls = k.take<float>(N);
    s1.feats = k.take<float>(N * D); s1.det_norm = k.take<float>(N);
    s1.det_planes = k.take<unsigned char>((size_t)2 * (D / 8) * SSB_DET_PLANES_MAX * 16);
    if (slot1) *slot1 = s1;
    if (tt) *tt = t;
    if (fs) *fs = f;
    if (reid_ws) *reid_ws = rw;
    if (reid_floats) *reid_floats = rf;
    if (boxes_tmp) *boxes_tmp = bt;
    return (k.off + 255) & ~(size_t)255;
}

extern "C" int64_t ssb_workspace_bytes(const ssb_config *cfg) {
    if (check_cfg(cfg)) return -1;
    return (int64_t)carve(cfg, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int ssb_create(const ssb_config *cfg, void *workspace_dev, int64_t workspace_bytes,
                          ssb_tracker **out) {
    if (check_cfg(cfg)) return -1;
    if (!out || !workspace_dev) { ssb_set_error("null argument"); return -1; }
    const int64_t need = ssb_workspace_bytes(cfg);
    if (workspace_bytes < need) {
        ssb_set_error("workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)need);
        return -1;
    }
    if (((uintptr_t)workspace_dev & 255) != 0) { ssb_set_error("workspace must be 256-byte aligned"); return -1; }
    ssb_tracker *t = (ssb_tracker *)calloc(1, sizeof(ssb_tracker));
    if (!t) { ssb_set_error("out of host memory"); return -1; }
    t->cfg = *cfg;
    t->ws_base = (char *)workspace_dev;
    t->ws_bytes = workspace_bytes;
    carve(cfg, t->ws_base, &t->tt, &t->fs, &t->reid_ws, &t->reid_ws_floats, &t->boxes_tmp, &t->tc_status,
          &t->slot[1], &t->reid_ws1);
    t->slot[0] = DetSlot{t->fs.det_tlwh, t->fs.det_xyah, t->fs.det_box, t->fs.det_conf, t->fs.det_cls,
                         t->fs.feats, t->fs.det_norm, t->fs.det_planes};
    SsbDims &d = t->dims;
    d.S = cfg->max_tracks; d.N = cfg->max_dets; d.B = cfg->nn_budget; d.D = cfg->feat_dim;
    d.n_init = cfg->n_init; d.max_age = cfg->max_age;
    d.max_dist = cfg->max_dist; d.max_iou = cfg->max_iou_distance;
    d.mc_lambda = cfg->mc_lambda; d.one_minus_lambda = 1 - cfg->mc_lambda;
    // python-float weights applied to float32 arrays: rounded once to float32
    d.ema_alpha = (float)cfg->ema_alpha;
    d.one_minus_alpha = (float)(1.0 - cfg->ema_alpha);
    *out = t;
    return 0;
}

extern "C" int ssb_destroy(ssb_tracker *t) {
    if (!t) return 0;
    for (int i = 0; i < 2; i++) {
        for (int p = 0; p < 3; p++)
            if (t->side_stream[i][p]) { cudaStreamDestroy(t->side_stream[i][p]); cudaEventDestroy(t->ev_join[i][p]); }
        if (t->ev_fork[i]) cudaEventDestroy(t->ev_fork[i]);
    }
    if (t->prof_ev[0])
        for (int i = 0; i < 12; i++) cudaEventDestroy(t->prof_ev[i]);
    free(t->w_off);
    free(t);
    return 0;
}

extern "C" int ssb_appearance_use_tc(ssb_tracker *t, int enable) {
    if (!t) { ssb_set_error("null handle"); return -1; }
    t->app_simt = enable ? 0 : 1;
    return 0;
}

extern "C" int ssb_reset(ssb_tracker *t, ssb_stream_t stream) {
    if (!t) { ssb_set_error("null handle"); return -1; }
    return ssb_launch_reset(t, (cudaStream_t)stream);
}

static int reid_forward_split(ssb_tracker *t, int slot, const uint8_t *img_dev, int h, int w, int pitch, const int *boxes,
                              int n, float *feats, cudaStream_t st);

// stage 1 of a frame: detection prep + OSNet embeddings into slot (0/1).  Independent of
// the track table, so frame t+1 may be embedded while frame t is still being associated.
extern "C" int ssb_embed(ssb_tracker *t, int slot, const float *dets_dev, int n, const uint8_t *img_dev,
                         int h, int w, int pitch, ssb_stream_t stream) {
    if (!t) { ssb_set_error("null handle"); return -1; }
    if (n < 0 || n > t->dims.N) { ssb_set_error("n=%d outside [0,%d]", n, t->dims.N); return -1; }
    if (n > 0 && !dets_dev) { ssb_set_error("null dets"); return -1; }
    if (h <= 0 || w <= 0) { ssb_set_error("bad image size %dx%d", w, h); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    FrameScratch fs = ssb_slot_view(t, slot);
    int rc = ssb_launch_prep(t->dims, dets_dev, n, h, w, fs, st);
    if (rc || n == 0) return rc;
    if (!img_dev) return 0;                         // caller supplies embeddings to ssb_associate
    if (pitch < 3 * w) { ssb_set_error("bad pitch"); return -1; }
    if (n >= 48 && ssb_split_parts() > 1)
        return reid_forward_split(t, slot, img_dev, h, w, pitch, fs.det_box, n, fs.feats, st);
    return ssb_reid_forward(t, slot, img_dev, h, w, pitch, fs.det_box, n, fs.feats, st);
}

// stage 2: association + track-table update from the detections/embeddings of `slot`
extern "C" int ssb_associate(ssb_tracker *t, int slot, int n, int h, int w, const float *feats_dev,
                             double *out_dev, int32_t *counts_dev, int track_hint, ssb_stream_t stream) {
    if (!t || !out_dev || !counts_dev) { ssb_set_error("null argument"); return -1; }
    if (n < 0 || n > t->dims.N) { ssb_set_error("n=%d outside [0,%d]", n, t->dims.N); return -1; }
    return ssb_launch_track_frame(t, slot, n, h, w, feats_dev, out_dev, counts_dev, track_hint,
                                  (cudaStream_t)stream);
}

// Embedding of a frame's crops in `parts` parts on as many streams (fork / join by events, still graph-capturable),
// inside ONE detection slot's workspace, so it serves the synchronous calls (ssb_update, ssb_reid) and the two-stage
// pipeline (ssb_embed of frame k+1 while frame k is associated) alike.  Every ReID kernel ends in a partial wave; with
// independent parts in flight the block scheduler fills one part's tail with another part's CTAs.
static int reid_forward_split(ssb_tracker *t, int slot, const uint8_t *img_dev, int h, int w, int pitch, const int *boxes,
                              int n, float *feats, cudaStream_t st) {
    slot &= 1;
    int parts = ssb_split_parts();
    if (parts > n / 24) parts = n / 24;               // a part below ~24 crops no longer fills the GPU's tail, it is one
    if (parts < 2) return ssb_reid_forward(t, slot, img_dev, h, w, pitch, boxes, n, feats, st);
    // same priority as the caller's stream: the parts must interleave CTA by CTA (a high-priority part simply
    // runs first and the tail-filling effect is gone: 591 instead of 468 us per 99 crops, measured) -- so the side
    // streams are re-made when a caller arrives on a stream of another priority
    int prio = 0;
    SSB_CHECK_CUDA(cudaStreamGetPriority(st, &prio));
    if (!t->ev_fork[slot]) SSB_CHECK_CUDA(cudaEventCreateWithFlags(&t->ev_fork[slot], cudaEventDisableTiming));
    cudaStream_t streams[4] = {st, nullptr, nullptr, nullptr};
    for (int p = 0; p < parts - 1; p++) {
        cudaStream_t &ss = t->side_stream[slot][p];
        if (ss && t->side_prio[slot] != prio) {
            SSB_CHECK_CUDA(cudaStreamSynchronize(ss));
            SSB_CHECK_CUDA(cudaStreamDestroy(ss));
            ss = nullptr;
        }
        if (!ss) {
            SSB_CHECK_CUDA(cudaStreamCreateWithPriority(&ss, cudaStreamNonBlocking, prio));
            if (!t->ev_join[slot][p]) SSB_CHECK_CUDA(cudaEventCreateWithFlags(&t->ev_join[slot][p], cudaEventDisableTiming));
        }
        streams[p + 1] = ss;
    }
    t->side_prio[slot] = prio;
    SSB_CHECK_CUDA(cudaEventRecord(t->ev_fork[slot], st));
    for (int p = 1; p < parts; p++) SSB_CHECK_CUDA(cudaStreamWaitEvent(streams[p], t->ev_fork[slot], 0));
    int rc = ssb_reid_forward_parts(t, slot, img_dev, h, w, pitch, boxes, n, feats, parts, streams);
    if (rc == 1) rc = ssb_reid_forward(t, slot, img_dev, h, w, pitch, boxes, n, feats, st);   // baseline modes: unsplit
    if (rc) return rc;
    for (int p = 1; p < parts; p++) {
        SSB_CHECK_CUDA(cudaEventRecord(t->ev_join[slot][p - 1], streams[p]));
        SSB_CHECK_CUDA(cudaStreamWaitEvent(st, t->ev_join[slot][p - 1], 0));
    }
    return 0;
}

extern "C" int ssb_update(ssb_tracker *t, const float *dets_dev, int n, const uint8_t *img_dev,
                          int h, int w, int pitch, const float *feats_dev, double *out_dev,
                          int32_t *counts_dev, int track_hint, ssb_stream_t stream) {
    if (!t || !out_dev || !counts_dev) { ssb_set_error("null argument"); return -1; }
    if (!feats_dev && n > 0 && !img_dev) { ssb_set_error("null image"); return -1; }
    int rc = ssb_embed(t, 0, dets_dev, n, feats_dev ? nullptr : img_dev, h, w, pitch, stream);
    if (rc) return rc;
    return ssb_associate(t, 0, n, h, w, feats_dev, out_dev, counts_dev, track_hint, stream);
}

extern "C" int ssb_reid(ssb_tracker *t, const uint8_t *img_dev, int h, int w, int pitch,
                        const int32_t *boxes_dev, int n, float *feats_out_dev, ssb_stream_t stream) {
    if (!t || !img_dev || !feats_out_dev) { ssb_set_error("null argument"); return -1; }
    if (n < 0 || n > t->dims.N) { ssb_set_error("n=%d outside [0,%d]", n, t->dims.N); return -1; }
    if (n == 0) return 0;
    if (n >= 48 && pitch >= 3 * w && ssb_split_parts() > 1)
        return reid_forward_split(t, 0, img_dev, h, w, pitch, boxes_dev, n, feats_out_dev, (cudaStream_t)stream);
    return ssb_reid_forward(t, 0, img_dev, h, w, pitch, boxes_dev, n, feats_out_dev, (cudaStream_t)stream);
}

extern "C" int ssb_export_tracks(ssb_tracker *t, int32_t *ids, int32_t *state, int32_t *hits,
                                 int32_t *age, int32_t *tsu, int32_t *gallery_len, double *mean,
                                 double *cov, float *feat, ssb_stream_t stream) {
    if (!t) { ssb_set_error("null handle"); return -1; }
    return ssb_launch_export(t, ids, state, hits, age, tsu, gallery_len, mean, cov, feat, (cudaStream_t)stream);
}

#ifdef SSB_BASELINES
extern "C" int ssb_debug_cost_ptrs(ssb_tracker *t, const double **cost_a_dev,
                                   const double **cost_b_dev, const int32_t **dims_dev) {
    if (!t) { ssb_set_error("null handle"); return -1; }
    if (cost_a_dev) *cost_a_dev = t->fs.cost_a;
    if (cost_b_dev) *cost_b_dev = t->fs.cost_b;
    if (dims_dev) *dims_dev = t->fs.cnt + FC_ROWS_A;
    return 0;
}
#endif
