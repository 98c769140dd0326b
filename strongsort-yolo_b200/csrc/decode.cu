// YOLOv8 detection / pose head decode on the GPU: the raw head tensor
//   raw [4*16 + nc + 3*K][A]   (DFL box bins, class logits, K keypoints x (x, y, vis)), A = anchors of
//   the stride-8/16/32 levels of an (in_h x in_w) network input, level after level, row-major
// becomes the decoded tensor ssb_yolo_nms() consumes
//   pred [4 + nc + 3*K][A]     (cx, cy, w, h in input pixels, class probabilities, keypoints).
// Arithmetic follows ultralytics' Detect/Pose inference path (third-party, not vendored): softmax over
// the 16 DFL bins -> expectation -> dist2bbox(xywh) around the anchor centre (+0.5) -> x stride;
// sigmoid class scores; keypoints (v*2 + anchor - 0.5) * stride, visibility sigmoid.  float32.
//
// One thread per anchor; every channel row is read/written coalesced over anchors.  HBM-bound:
// (64 + nc + 3K + 4 + nc + 3K) * A * 4 bytes = 1.9 MB for nc = 80, A = 8400 -- a few microseconds.
#include "ssb_common.cuh"

namespace {

// conf_out / cls_out (nullable): the anchor's best class probability and index, the first maximum winning like
// torch.max -- what ssb_yolo_nms's score pass would recompute from `out` (fused here: one pass less over the head).
// CTA = 32 consecutive anchors x 4 roles (warp r: DFL side r, a quarter of the classes and of the keypoints), so
// every channel row is still read / written as one coalesced 128-byte run per warp, with 4x the threads and a
// quarter of the serial expf chain per thread of the one-thread-per-anchor version (48.6 -> ~12 us at A = 5040).
__global__ void __launch_bounds__(128)
yolo_decode_v8_kernel(const float *__restrict__ raw, int nc, int nk, int in_h, int in_w,
                      int A, float *__restrict__ out, float *__restrict__ conf_out,
                      int *__restrict__ cls_out) {
    __shared__ float s_d[4][32];
    __shared__ float s_best[4][32];
    __shared__ int s_bi[4][32];
    const int al = threadIdx.x & 31, role = threadIdx.x >> 5;
    const int a = blockIdx.x * 32 + al;
    const bool live = a < A;
    // level / grid position of anchor a
    int rem = live ? a : 0, stride = 8, gw = in_w / 8, gh = in_h / 8;
    for (int l = 0; l < 3; l++) {
        if (rem < gw * gh) break;
        rem -= gw * gh;
        stride *= 2; gw = in_w / stride; gh = in_h / stride;
    }
    const float ax = (float)(rem % gw) + 0.5f, ay = (float)(rem / gw) + 0.5f, fs = (float)stride;
    if (live) {                                   // DFL expectation of side `role`
        float v[16], m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 16; j++) { v[j] = raw[(size_t)(role * 16 + j) * A + a]; m = fmaxf(m, v[j]); }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 16; j++) { v[j] = expf(v[j] - m); sum += v[j]; }
        float e = 0.f;
#pragma unroll
        for (int j = 0; j < 16; j++) e += (v[j] / sum) * (float)j;
        s_d[role][al] = e;
    }
    float best = -INFINITY;
    int bi = 0;
    const int c0 = (nc * role) / 4, c1 = (nc * (role + 1)) / 4;
    if (live) {
        for (int c = c0; c < c1; c++) {
            const float p = 1.f / (1.f + expf(-raw[(size_t)(64 + c) * A + a]));
            out[(size_t)(4 + c) * A + a] = p;
            if (c == c0 || p > best) { best = p; bi = c; }
        }
        for (int k = role; k < nk; k += 4) {
            const size_t ri = (size_t)(64 + nc + 3 * k) * A + a, oi = (size_t)(4 + nc + 3 * k) * A + a;
            out[oi] = (raw[ri] * 2.f + (ax - 0.5f)) * fs;
            out[oi + A] = (raw[ri + A] * 2.f + (ay - 0.5f)) * fs;
            out[oi + 2 * (size_t)A] = 1.f / (1.f + expf(-raw[ri + 2 * (size_t)A]));
        }
    }
    s_best[role][al] = c1 > c0 ? best : -INFINITY;
    s_bi[role][al] = bi;
    __syncthreads();
    if (role == 0 && live) {
        const float d0 = s_d[0][al], d1 = s_d[1][al], d2 = s_d[2][al], d3 = s_d[3][al];
        const float x1 = ax - d0, y1 = ay - d1, x2 = ax + d2, y2 = ay + d3;
        out[(size_t)0 * A + a] = ((x1 + x2) / 2.f) * fs;
        out[(size_t)1 * A + a] = ((y1 + y2) / 2.f) * fs;
        out[(size_t)2 * A + a] = (x2 - x1) * fs;
        out[(size_t)3 * A + a] = (y2 - y1) * fs;
        if (conf_out) {                           // roles hold ascending class ranges: strict > keeps the first maximum
            float bb = s_best[0][al];
            int bc = s_bi[0][al];
#pragma unroll
            for (int r = 1; r < 4; r++)
                if (s_best[r][al] > bb) { bb = s_best[r][al]; bc = s_bi[r][al]; }
            conf_out[a] = bb; cls_out[a] = bc;
        }
    }
}

// YOLOv5 / v7 head (the detectors upstream StrongSORT-YOLO wires to the tracker): raw [A][5 + nc] logits,
// A = 3 anchors x the stride-8/16/32 grids, level by level, anchor-major inside a level ([na][gy][gx]).
//   xy = (2 s(x) - 0.5 + grid) * stride ; wh = (2 s(w))^2 * anchor ; obj, cls = sigmoid
// and non_max_suppression's scoring folded in: score_c = obj * cls_c, zeroed when obj <= conf_thres
// (yolov5 utils/general.py: `xc = prediction[..., 4] > conf_thres`, `x[:, 5:] *= x[:, 4:5]`), so the result
// is the channel-major [4 + nc][A] tensor ssb_yolo_nms consumes.  anchors_px: 3 levels x 3 anchors x (w, h).
__global__ void yolo_decode_v5_kernel(const float *__restrict__ raw, int nc, int in_h, int in_w, int A,
                                      float conf_thres, const float *__restrict__ anchors_px,
                                      float *__restrict__ out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= A) return;
    int rem = a, stride = 8, gw = in_w / 8, gh = in_h / 8, lvl = 0;
    for (; lvl < 2; lvl++) {
        if (rem < 3 * gw * gh) break;
        rem -= 3 * gw * gh;
        stride *= 2; gw = in_w / stride; gh = in_h / stride;
    }
    const int ia = rem / (gw * gh), cell = rem - ia * gw * gh;
    const float gx = (float)(cell % gw), gy = (float)(cell / gw), fs = (float)stride;
    const float *p = raw + (size_t)a * (5 + nc);
    auto sg = [](float v) { return 1.f / (1.f + expf(-v)); };
    const float sx = sg(p[0]), sy = sg(p[1]), sw = sg(p[2]), sh = sg(p[3]), obj = sg(p[4]);
    out[(size_t)0 * A + a] = (sx * 2.f - 0.5f + gx) * fs;
    out[(size_t)1 * A + a] = (sy * 2.f - 0.5f + gy) * fs;
    out[(size_t)2 * A + a] = (sw * 2.f) * (sw * 2.f) * anchors_px[(lvl * 3 + ia) * 2 + 0];
    out[(size_t)3 * A + a] = (sh * 2.f) * (sh * 2.f) * anchors_px[(lvl * 3 + ia) * 2 + 1];
    const bool cand = obj > conf_thres;
    for (int c = 0; c < nc; c++) out[(size_t)(4 + c) * A + a] = cand ? sg(p[5 + c]) * obj : 0.f;
}

}  // namespace

extern "C" int ssb_yolo_decode_v5(const float *raw_dev, int num_classes, int in_h, int in_w, float conf_thres,
                                  const float *anchors_px_dev, float *pred_out_dev, ssb_stream_t stream) {
    if (!raw_dev || !pred_out_dev || !anchors_px_dev) { ssb_set_error("null argument"); return -1; }
    if (in_h <= 0 || in_w <= 0 || in_h % 32 || in_w % 32 || num_classes < 1) { ssb_set_error("bad head geometry"); return -1; }
    const int A = 3 * ((in_h / 8) * (in_w / 8) + (in_h / 16) * (in_w / 16) + (in_h / 32) * (in_w / 32));
    yolo_decode_v5_kernel<<<(A + 127) / 128, 128, 0, (cudaStream_t)stream>>>(raw_dev, num_classes, in_h, in_w, A, conf_thres,
                                                                             anchors_px_dev, pred_out_dev);
    SSB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ssb_yolo_num_anchors(int in_h, int in_w) {
    if (in_h <= 0 || in_w <= 0 || in_h % 32 || in_w % 32) return -1;
    return (in_h / 8) * (in_w / 8) + (in_h / 16) * (in_w / 16) + (in_h / 32) * (in_w / 32);
}

extern "C" int ssb_yolo_decode_v8(const float *raw_dev, int num_classes, int num_kpts, int in_h, int in_w,
                                  float *pred_out_dev, ssb_stream_t stream) {
    if (!raw_dev || !pred_out_dev) { ssb_set_error("null argument"); return -1; }
    const int A = ssb_yolo_num_anchors(in_h, in_w);
    if (A <= 0 || num_classes < 1 || num_kpts < 0) { ssb_set_error("bad head geometry %dx%d nc=%d kpts=%d", in_h, in_w, num_classes, num_kpts); return -1; }
    yolo_decode_v8_kernel<<<(A + 31) / 32, 128, 0, (cudaStream_t)stream>>>(raw_dev, num_classes, num_kpts, in_h, in_w, A,
                                                                            pred_out_dev, nullptr, nullptr);
    SSB_CHECK_LAUNCH();
    return 0;
}

// decode + best class per anchor in one pass (ssb_yolo_postprocess_v8, nms.cu)
int ssb_launch_decode_v8_scored(const float *raw, int nc, int nk, int in_h, int in_w, float *pred, float *conf, int *cls,
                                cudaStream_t st) {
    const int A = ssb_yolo_num_anchors(in_h, in_w);
    if (A <= 0 || nc < 1 || nk < 0) { ssb_set_error("bad head geometry %dx%d nc=%d kpts=%d", in_h, in_w, nc, nk); return -1; }
    yolo_decode_v8_kernel<<<(A + 31) / 32, 128, 0, st>>>(raw, nc, nk, in_h, in_w, A, pred, conf, cls);
    SSB_CHECK_LAUNCH();
    return 0;
}
