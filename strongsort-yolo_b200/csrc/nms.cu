// YOLO post-process: confidence filter + class-aware NMS.
// In-tree facts: conf 0.3, iou 0.4, agnostic_nms False, max_det 1000
// (/root/reference/yolo_multi_model.py:18-21).  The arithmetic mirrors
// ultralytics' non_max_suppression -> torchvision.ops.nms (third-party, not
// vendored): best class only, xywh->xyxy in float32, class offset 7680,
// stable descending score order, greedy suppression on IoU > thr, first
// max_det survivors.  Bit-exact vs torchvision's CPU kernel (same float32
// expression order, IEEE division).
//
// HBM-bound on the score read ((4+nc)*A*4 bytes); everything after the filter
// touches only the few hundred candidates.
#include "ssb_common.cuh"

#define NMS_MAX_WH 7680.0f
// ultralytics keeps at most max_nms = 30000 candidates (highest scores) before torchvision.ops.nms;
// the suppression matrix and its grids are sized by min(A, NMS_MAX_CAND)
#define NMS_MAX_CAND 30000

struct NmsScratch {
    float *conf;        // [A]
    int *cls;           // [A]
    int *cand;          // [A] anchors passing the filter, anchor order
    int *order;         // [A] candidate index sorted by score (stable, desc)
    float *sbox;        // [A][4] offset boxes in sorted order
    float *sarea;       // [A]
    unsigned long long *mask;   // [M][ceil(M/64)]
    int *count;         // [4]: M
};

__host__ __device__ inline size_t nms_align(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t nms_carve(char *base, int A, NmsScratch *s) {
    size_t off = 0;
    auto take = [&](size_t bytes) { char *p = base ? base + off : nullptr; off = nms_align(off + bytes); return p; };
    NmsScratch t;
    t.conf = (float *)take((size_t)A * 4);
    t.cls = (int *)take((size_t)A * 4);
    t.cand = (int *)take((size_t)A * 4);
    t.order = (int *)take((size_t)A * 4);
    t.sbox = (float *)take((size_t)A * 16);
    t.sarea = (float *)take((size_t)A * 4);
    t.count = (int *)take(64);
    const size_t Mc = A < NMS_MAX_CAND ? A : NMS_MAX_CAND;
    const size_t words = (Mc + 63) / 64;
    t.mask = (unsigned long long *)take(Mc * words * 8);
    if (s) *s = t;
    return off;
}

extern "C" int64_t ssb_nms_scratch_bytes(int num_anchors) {
    if (num_anchors <= 0) return 0;
    return (int64_t)nms_carve(nullptr, num_anchors, nullptr);
}

// best class per anchor: pred is channel-major [4+nc+extra][A] -> coalesced over anchors
__global__ void nms_score_kernel(const float *__restrict__ pred, int nc, int A, NmsScratch s) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= A) return;
    float best = pred[(size_t)4 * A + a];
    int bi = 0;
    for (int c = 1; c < nc; c++) {
        const float v = pred[(size_t)(4 + c) * A + a];
        if (v > best) { best = v; bi = c; }     // first maximum wins, like torch.max
    }
    s.conf[a] = best;
    s.cls[a] = bi;
}

__device__ __forceinline__ int warp_incl_scan_i(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// ordered compaction of anchors with conf > thr (single CTA, 1024 threads)
__global__ void __launch_bounds__(1024)
nms_compact_kernel(int A, float conf_thres, NmsScratch s) {
    __shared__ int s_w[33];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (A + 1023) / 1024;
    const int b = tid * per, e = min(A, b + per);
    int c = 0;
    for (int a = b; a < e; a++) c += s.conf[a] > conf_thres;
    int inc = warp_incl_scan_i(c, lane);
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = s_w[lane];
        int wi = warp_incl_scan_i(w, lane);
        s_w[lane] = wi - w;
        if (lane == 31) s_w[32] = wi;
    }
    __syncthreads();
    int off = s_w[wid] + inc - c;
    for (int a = b; a < e; a++)
        if (s.conf[a] > conf_thres) s.cand[off++] = a;
    if (tid == 0) s.count[0] = s_w[32];
}

// stable descending rank by counting; writes order[], offset boxes and areas
__global__ void nms_rank_kernel(const float *__restrict__ pred, int A, int agnostic, NmsScratch s) {
    const int M = s.count[0];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    const int a = s.cand[k];
    const float sc = s.conf[a];
    int rank = 0;
    for (int j = 0; j < M; j++) {
        const float sj = s.conf[s.cand[j]];
        rank += (sj > sc) || (sj == sc && j < k);
    }
    if (rank >= NMS_MAX_CAND) return;        // beyond max_nms: never a candidate
    s.order[rank] = k;
    // xywh -> xyxy (float32), then the class offset
    const float x = pred[a], y = pred[(size_t)A + a];
    const float dw = pred[(size_t)2 * A + a] / 2.f, dh = pred[(size_t)3 * A + a] / 2.f;
    const float c = agnostic ? 0.f : (float)s.cls[a] * NMS_MAX_WH;
    const float x1 = (x - dw) + c, y1 = (y - dh) + c, x2 = (x + dw) + c, y2 = (y + dh) + c;
    s.sbox[rank * 4 + 0] = x1; s.sbox[rank * 4 + 1] = y1;
    s.sbox[rank * 4 + 2] = x2; s.sbox[rank * 4 + 3] = y2;
    s.sarea[rank] = (x2 - x1) * (y2 - y1);
}

// suppression bit matrix over score-sorted boxes: bit (i, j) = IoU(i, j) > thr, j > i
__global__ void nms_mask_kernel(float iou_thres, NmsScratch s) {
    const int M = min(s.count[0], NMS_MAX_CAND);
    const int words = (M + 63) / 64;
    const int wj = blockIdx.x * blockDim.x + threadIdx.x;
    if (wj >= words) return;
    // the candidate count lives on the device: a bounded grid strides over the rows (the worst-case grid of
    // A / 8 row blocks cost 22 us of empty CTAs per frame)
    for (int i = blockIdx.y * blockDim.y + threadIdx.y; i < M; i += gridDim.y * blockDim.y) {
    const float ix1 = s.sbox[i * 4], iy1 = s.sbox[i * 4 + 1];
    const float ix2 = s.sbox[i * 4 + 2], iy2 = s.sbox[i * 4 + 3];
    const float iarea = s.sarea[i];
    unsigned long long bits = 0;
    const int j0 = wj * 64;
    for (int b = 0; b < 64; b++) {
        const int j = j0 + b;
        if (j >= M) break;
        if (j <= i) continue;
        const float xx1 = fmaxf(ix1, s.sbox[j * 4]), yy1 = fmaxf(iy1, s.sbox[j * 4 + 1]);
        const float xx2 = fminf(ix2, s.sbox[j * 4 + 2]), yy2 = fminf(iy2, s.sbox[j * 4 + 3]);
        const float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
        const float inter = w * h;
        const float ovr = inter / (iarea + s.sarea[j] - inter);
        if (ovr > iou_thres) bits |= 1ull << b;
    }
    s.mask[(size_t)i * words + wj] = bits;
    }
}

// greedy scan + gather (single CTA).  The suppression matrix of the M candidates is first copied into shared
// memory when it fits (M <= ~700: 48 KB); ONE warp then walks the sorted candidates with the "removed" bit set
// spread over its lanes (lane l owns words l, l + 32: M <= 4096) -- no block barrier per kept box (the round-1
// kernel paid two __syncthreads and a dependent global read per box: 37 us at M = 100) -- and the whole CTA
// gathers the kept rows.  gain > 0 folds ultralytics' scale_boxes (letterbox -> frame) into the gather.
#define NMS_SMEM_MASK_BYTES 49152
__global__ void __launch_bounds__(256)
nms_scan_kernel(const float *__restrict__ pred, int nc, int n_extra, int A, int max_det,
                NmsScratch s, float *__restrict__ out, int *__restrict__ count_out,
                float gain, float pad_x, float pad_y, float w0, float h0) {
    extern __shared__ unsigned long long s_mask[];
    __shared__ int s_keep_n;
    const int M = min(s.count[0], NMS_MAX_CAND);
    const int words = (M + 63) / 64;
    const bool cached = (size_t)M * words * 8 <= NMS_SMEM_MASK_BYTES;
    if (cached)
        for (int e = threadIdx.x; e < M * words; e += blockDim.x) s_mask[e] = s.mask[e];
    int *keep = reinterpret_cast<int *>(s.sarea);  // sarea[] (float[A]) is dead after the mask kernel: the kept ranks
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int nkeep = 0;
        if (words <= 64) {
            unsigned long long rem0 = 0ull, rem1 = 0ull;
            for (int i = 0; i < M && nkeep < max_det; i++) {
                const int w = i >> 6;
                const unsigned long long rw = __shfl_sync(0xffffffffu, (w >> 5) ? rem1 : rem0, w & 31);
                if ((rw >> (i & 63)) & 1ull) continue;                 // warp-uniform
                if (lane == 0) keep[nkeep] = i;
                nkeep++;
                const unsigned long long *row = cached ? s_mask + (size_t)i * words : s.mask + (size_t)i * words;
                if (lane < words) rem0 |= row[lane];
                if (lane + 32 < words) rem1 |= row[lane + 32];
            }
        } else {                              // > 4096 candidates (rare): test each box against the kept ones' mask rows
            for (int i = 0; i < M && nkeep < max_det; i++) {
                bool removed = false;
                for (int k0 = 0; k0 < nkeep && !removed; k0 += 32) {
                    const int k = k0 + lane;
                    const bool hit = k < nkeep && ((s.mask[(size_t)keep[k] * words + (i >> 6)] >> (i & 63)) & 1ull);
                    removed = __any_sync(0xffffffffu, hit);
                }
                if (removed) continue;
                if (lane == 0) keep[nkeep] = i;
                __syncwarp();
                nkeep++;
            }
        }
        if (lane == 0) s_keep_n = nkeep;
    }
    __syncthreads();
    const int nkeep = s_keep_n;
    const int cols = 6 + n_extra;
    for (int r = threadIdx.x; r < nkeep; r += blockDim.x) {            // gather (original, un-offset boxes)
        const int a = s.cand[s.order[keep[r]]];
        const float x = pred[a], y = pred[(size_t)A + a];
        const float dw = pred[(size_t)2 * A + a] / 2.f, dh = pred[(size_t)3 * A + a] / 2.f;
        float x1 = x - dw, y1 = y - dh, x2 = x + dw, y2 = y + dh;
        if (gain > 0.f) {
            x1 = fminf(fmaxf((x1 - pad_x) / gain, 0.f), w0); y1 = fminf(fmaxf((y1 - pad_y) / gain, 0.f), h0);
            x2 = fminf(fmaxf((x2 - pad_x) / gain, 0.f), w0); y2 = fminf(fmaxf((y2 - pad_y) / gain, 0.f), h0);
        }
        float *o = out + (size_t)r * cols;
        o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2;
        o[4] = s.conf[a]; o[5] = (float)s.cls[a];
        for (int e = 0; e < n_extra; e++) o[6 + e] = pred[(size_t)(4 + nc + e) * A + a];
    }
    if (threadIdx.x == 0) count_out[0] = nkeep;
}

static int nms_after_scores(const float *pred_dev, int num_classes, int num_extra, int A, float conf_thres, float iou_thres,
                            int max_det, int agnostic, float *out_dev, int32_t *count_dev, const NmsScratch &s,
                            float gain, float pad_x, float pad_y, float w0, float h0, cudaStream_t st) {
    nms_compact_kernel<<<1, 1024, 0, st>>>(A, conf_thres, s);
    SSB_CHECK_LAUNCH();
    // the candidate count lives on the device: grids are sized for the worst case
    nms_rank_kernel<<<(A + 127) / 128, 128, 0, st>>>(pred_dev, A, agnostic, s);
    SSB_CHECK_LAUNCH();
    const int Mc = A < NMS_MAX_CAND ? A : NMS_MAX_CAND;
    const int words = (Mc + 63) / 64;
    dim3 mb(32, 8), mg((words + 31) / 32, ((Mc + 7) / 8) < 64 ? ((Mc + 7) / 8) : 64);
    nms_mask_kernel<<<mg, mb, 0, st>>>(iou_thres, s);
    SSB_CHECK_LAUNCH();
    static const int key = ssb_new_key();
    if (ssb_first_on_device(key))
        SSB_CHECK_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NMS_SMEM_MASK_BYTES));
    nms_scan_kernel<<<1, 256, NMS_SMEM_MASK_BYTES, st>>>(pred_dev, num_classes, num_extra, A, max_det, s, out_dev, count_dev,
                                                        gain, pad_x, pad_y, w0, h0);
    SSB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ssb_yolo_nms(const float *pred_dev, int num_classes, int num_extra, int num_anchors,
                            float conf_thres, float iou_thres, int max_det, int agnostic,
                            float *out_dev, int32_t *count_dev, void *scratch_dev,
                            ssb_stream_t stream) {
    if (!pred_dev || !out_dev || !count_dev || !scratch_dev) { ssb_set_error("null argument"); return -1; }
    if (num_classes < 1 || num_anchors < 1 || num_extra < 0 || max_det < 1) { ssb_set_error("bad NMS dims"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    NmsScratch s;
    nms_carve((char *)scratch_dev, num_anchors, &s);
    const int A = num_anchors;
    nms_score_kernel<<<(A + 255) / 256, 256, 0, st>>>(pred_dev, num_classes, A, s);
    SSB_CHECK_LAUNCH();
    return nms_after_scores(pred_dev, num_classes, num_extra, A, conf_thres, iou_thres, max_det, agnostic, out_dev, count_dev,
                            s, 0.f, 0.f, 0.f, 0.f, 0.f, st);
}

int ssb_launch_decode_v8_scored(const float *raw, int nc, int nk, int in_h, int in_w, float *pred, float *conf, int *cls,
                                cudaStream_t st);

// The whole detector post-process of a YOLOv8 head in ONE call, 5 launches: decode (+ best class per anchor fused),
// ordered compaction, stable rank, suppression matrix, scan + gather (+ scale_boxes fused when gain > 0).
// Same results as ssb_yolo_decode_v8 -> ssb_yolo_nms -> ssb_yolo_scale_boxes.
extern "C" int ssb_yolo_postprocess_v8(const float *raw_dev, int num_classes, int num_kpts, int in_h, int in_w,
                                       float conf_thres, float iou_thres, int max_det, int agnostic,
                                       float gain, float pad_x, float pad_y, int w0, int h0,
                                       float *pred_scratch_dev, float *out_dev, int32_t *count_dev, void *scratch_dev,
                                       ssb_stream_t stream) {
    if (!raw_dev || !pred_scratch_dev || !out_dev || !count_dev || !scratch_dev) { ssb_set_error("null argument"); return -1; }
    if (max_det < 1) { ssb_set_error("bad NMS dims"); return -1; }
    const int A = ssb_yolo_num_anchors(in_h, in_w);
    if (A <= 0) { ssb_set_error("bad head geometry"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    NmsScratch s;
    nms_carve((char *)scratch_dev, A, &s);
    int rc = ssb_launch_decode_v8_scored(raw_dev, num_classes, num_kpts, in_h, in_w, pred_scratch_dev, s.conf, s.cls, st);
    if (rc) return rc;
    return nms_after_scores(pred_scratch_dev, num_classes, 3 * num_kpts, A, conf_thres, iou_thres, max_det, agnostic, out_dev,
                            count_dev, s, gain, pad_x, pad_y, (float)w0, (float)h0, st);
}

// ultralytics scale_boxes (ops.py; SURVEY.md C.2 "boxes are scaled back from letterbox to the original
// frame"): x -= pad_x, y -= pad_y, / gain (float32, IEEE division), clip to [0, w0] x [0, h0]; in place
// on the first count[0] rows of the NMS output.
__global__ void scale_boxes_kernel(float *__restrict__ rows, int cols, const int *__restrict__ count, int max_det,
                                   float gain, float pad_x, float pad_y, float w0, float h0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = min(count[0], max_det);
    if (i >= m) return;
    float *b = rows + (size_t)i * cols;
    b[0] = fminf(fmaxf((b[0] - pad_x) / gain, 0.f), w0);
    b[1] = fminf(fmaxf((b[1] - pad_y) / gain, 0.f), h0);
    b[2] = fminf(fmaxf((b[2] - pad_x) / gain, 0.f), w0);
    b[3] = fminf(fmaxf((b[3] - pad_y) / gain, 0.f), h0);
}

extern "C" int ssb_yolo_scale_boxes(float *rows_dev, int cols, const int32_t *count_dev, int max_det, float gain,
                                    float pad_x, float pad_y, int w0, int h0, ssb_stream_t stream) {
    if (!rows_dev || !count_dev || cols < 4 || max_det < 1 || !(gain > 0.f)) { ssb_set_error("bad scale_boxes argument"); return -1; }
    scale_boxes_kernel<<<(max_det + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rows_dev, cols, count_dev, max_det, gain,
                                                                              pad_x, pad_y, (float)w0, (float)h0);
    SSB_CHECK_LAUNCH();
    return 0;
}
