// Association-side kernels of StrongSORT.update(): detection prep, batched
// Kalman filter, Mahalanobis gating, IoU cost, linear assignment with scipy's
// tie-breaks, and the device-resident track table (state machine, id
// allocation, gallery ring).  float64 throughout, compiled with -fmad=false so
// that expression order matches the NumPy restatement (oracle/strongsort_np.py).
//
// Reference semantics: SURVEY.md Appendix A.4-A.8 (the upstream strong_sort/
// package is absent from /root/reference; the seam is yolo_multi_model.py:41).
// These kernels are latency-bound (working set <= 1-3 MB): the
// design goal is few launches, no host round trip, coalesced SoA access.
#include <math.h>
#include <stdio.h>

#include <cuda_fp16.h>
#include <stdlib.h>

#include "ssb_common.cuh"

#define KF_WP (1.0 / 20)
#define KF_WV (1.0 / 160)

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// exclusive scan of one int per thread over the block; returns exclusive
// prefix, *total gets the block sum.  s_w: >= 33 ints of shared memory.
__device__ int block_excl_scan(int v, int *s_w, int *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    int inc = warp_incl_scan(v, lane);
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < nw ? s_w[lane] : 0;
        int wi = warp_incl_scan(w, lane);
        s_w[lane] = wi - w;
        if (lane == 31) s_w[32] = wi;
    }
    __syncthreads();
    int ex = s_w[wid] + inc - v;
    *total = s_w[32];
    __syncthreads();
    return ex;
}

// lower Cholesky of a symmetric 4x4 (only the lower triangle of S is read)
__device__ __forceinline__ void chol4(const double S[4][4], double L[4][4]) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        double d = S[j][j];
#pragma unroll
        for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k];
        d = sqrt(d);
        L[j][j] = d;
#pragma unroll
        for (int i = j + 1; i < 4; i++) {
            double s = S[i][j];
#pragma unroll
            for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
            L[i][j] = s / d;
        }
    }
}

// KalmanFilter.project (A.4): mu = mean[:4], S = cov[:4,:4] + diag(std^2)
__device__ __forceinline__ void kf_project(const double *mean, const double *cov,
                                           double conf, double mu[4], double S[4][4]) {
    const double h = mean[3];
    double std[4] = {KF_WP * h, KF_WP * h, 1e-1, KF_WP * h};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        mu[i] = mean[i];
        std[i] = (1.0 - conf) * std[i];
#pragma unroll
        for (int j = 0; j < 4; j++) S[i][j] = cov[i * 8 + j];
        S[i][i] = S[i][i] + std[i] * std[i];
    }
}

// ---------------------------------------------------------------------------
// detection prep (A.2 / A.3): xyxy -> xywh -> tlwh, xyah, int crop boxes
// ---------------------------------------------------------------------------
__global__ void prep_dets_kernel(const float *__restrict__ dets, int n, int H, int W,
                                 FrameScratch fs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x1 = dets[i * 6 + 0], y1 = dets[i * 6 + 1];
    const float x2 = dets[i * 6 + 2], y2 = dets[i * 6 + 3];
    const float cx = (x1 + x2) / 2.f, cy = (y1 + y2) / 2.f;
    const float w = x2 - x1, h = y2 - y1;
    // _xywh_to_xyxy: python int() truncates toward zero
    int bx1 = max((int)(cx - w / 2.f), 0);
    int bx2 = min((int)(cx + w / 2.f), W - 1);
    int by1 = max((int)(cy - h / 2.f), 0);
    int by2 = min((int)(cy + h / 2.f), H - 1);
    fs.det_box[i * 4 + 0] = bx1; fs.det_box[i * 4 + 1] = by1;
    fs.det_box[i * 4 + 2] = bx2; fs.det_box[i * 4 + 3] = by2;
    const float tx = cx - w / 2.0f, ty = cy - h / 2.0f;   // _xywh_to_tlwh
    fs.det_tlwh[i * 4 + 0] = tx; fs.det_tlwh[i * 4 + 1] = ty;
    fs.det_tlwh[i * 4 + 2] = w;  fs.det_tlwh[i * 4 + 3] = h;
    // Detection.to_xyah in float32
    fs.det_xyah[i * 4 + 0] = tx + w / 2.f;
    fs.det_xyah[i * 4 + 1] = ty + h / 2.f;
    fs.det_xyah[i * 4 + 2] = w / h;
    fs.det_xyah[i * 4 + 3] = h;
    fs.det_conf[i] = dets[i * 6 + 4];
    fs.det_cls[i] = dets[i * 6 + 5];
}

__global__ void crop_boxes_kernel(const float *__restrict__ dets, int n, int H, int W,
                                  int *__restrict__ boxes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x1 = dets[i * 6 + 0], y1 = dets[i * 6 + 1];
    const float x2 = dets[i * 6 + 2], y2 = dets[i * 6 + 3];
    const float cx = (x1 + x2) / 2.f, cy = (y1 + y2) / 2.f;
    const float w = x2 - x1, h = y2 - y1;
    boxes[i * 4 + 0] = max((int)(cx - w / 2.f), 0);
    boxes[i * 4 + 1] = max((int)(cy - h / 2.f), 0);
    boxes[i * 4 + 2] = min((int)(cx + w / 2.f), W - 1);
    boxes[i * 4 + 3] = min((int)(cy + h / 2.f), H - 1);
}

// L2 norm of every raw embedding (float64 accumulation, rounded to float32)
__global__ void det_norm_kernel(const float *__restrict__ feats, int n, int D,
                                float *__restrict__ norm_out, unsigned char *__restrict__ planes, int npad) {
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    double s = 0;
    for (int k = lane; k < D; k += 32) {
        double v = feats[(size_t)w * D + k];
        s += v * v;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) norm_out[w] = (float)sqrt(s);
    // tensor-core appearance cost (appearance.cu): the unit embedding * 2^6 as hi/lo fp16 operand planes
    // [hl][D/8][npad][8]; lane l writes chunks 2l and 2l+1
    if (planes) {
        const float nrm = (float)sqrt(__shfl_sync(0xffffffffu, s, 0));
        for (int c = lane; c < D / 8; c += 32) {
            const float4 a = *reinterpret_cast<const float4 *>(feats + (size_t)w * D + c * 8);
            const float4 b = *reinterpret_cast<const float4 *>(feats + (size_t)w * D + c * 8 + 4);
            const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            __align__(16) __half h[8];
            __align__(16) __half l[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float u = (v[j] / nrm) * 64.0f;
                h[j] = __float2half_rn(u);
                l[j] = __float2half_rn(u - __half2float(h[j]));
            }
            unsigned char *dst = planes + ((size_t)c * npad + w) * 16;
            *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<uint4 *>(h);
            *reinterpret_cast<uint4 *>(dst + (size_t)(D / 8) * npad * 16) = *reinterpret_cast<uint4 *>(l);
        }
    }
}

// ---------------------------------------------------------------------------
// batched Kalman predict: one warp per track, lanes own cov elements
//   mean <- F mean ; cov <- F (cov F^T) + Q      (multi_dot order of the oracle)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void kf_predict_warp(double *mean, double *cov, int lane) {
    double m_old = lane < 8 ? mean[lane] : 0.0;
    double m_hi = lane < 4 ? mean[lane + 4] : 0.0;
    double out[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int e = lane + 32 * r, i = e >> 3, j = e & 7;
        double x = cov[i * 8 + j];
        if (j < 4) x = x + cov[i * 8 + j + 4];
        if (i < 4) {
            double y = cov[(i + 4) * 8 + j];
            if (j < 4) y = y + cov[(i + 4) * 8 + j + 4];
            x = x + y;
        }
        out[r] = x;
    }
    // motion noise from the OLD mean
    const double m0 = __shfl_sync(0xffffffffu, m_old, 0), m1 = __shfl_sync(0xffffffffu, m_old, 1);
    const double m2 = __shfl_sync(0xffffffffu, m_old, 2), m3 = __shfl_sync(0xffffffffu, m_old, 3);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int e = lane + 32 * r, i = e >> 3, j = e & 7;
        if (i == j) {
            double s;
            switch (i) {
                case 0: s = KF_WP * m0; break;
                case 1: s = KF_WP * m1; break;
                case 2: s = 1 * m2; break;
                case 3: s = KF_WP * m3; break;
                case 4: s = KF_WV * m0; break;
                case 5: s = KF_WV * m1; break;
                case 6: s = 0.1 * m2; break;
                default: s = KF_WV * m3; break;
            }
            out[r] = out[r] + s * s;
        }
        cov[e] = out[r];
    }
    if (lane < 4) mean[lane] = m_old + m_hi;
}

__global__ void kf_predict_tracks_kernel(TrackTable tt) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= tt.scalars[SC_N_TRACKS]) return;
    const int s = tt.order[w];
    kf_predict_warp(tt.mean + (size_t)s * 8, tt.cov + (size_t)s * 64, lane);
    if (lane == 0) { tt.age[s] += 1; tt.tsu[s] += 1; }
}

__global__ void kf_predict_arrays_kernel(double *mean, double *cov, int n) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    kf_predict_warp(mean + (size_t)w * 8, cov + (size_t)w * 64, lane);
}

// KalmanFilter.update for one track, executed by a single thread
__device__ void kf_update_thread(double *mean, double *cov, const float *xyah, double conf) {
    double mu[4], S[4][4], L[4][4];
    double m[8], P[64];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = mean[i];
    for (int i = 0; i < 64; i++) P[i] = cov[i];
    kf_project(m, P, conf, mu, S);
    chol4(S, L);
    // K = (cho_solve(S, (P H^T)^T))^T : row i of K solves S k = P[i, :4]
    double K[8][4];
    for (int i = 0; i < 8; i++) {
        double y[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {          // forward  L y = b
            double s = P[i * 8 + r];
#pragma unroll
            for (int k = 0; k < r; k++) s -= L[r][k] * y[k];
            y[r] = s / L[r][r];
        }
#pragma unroll
        for (int r = 3; r >= 0; r--) {         // backward L^T x = y
            double s = y[r];
#pragma unroll
            for (int k = r + 1; k < 4; k++) s -= L[k][r] * K[i][k];
            K[i][r] = s / L[r][r];
        }
    }
    double innov[4];
#pragma unroll
    for (int j = 0; j < 4; j++) innov[j] = (double)xyah[j] - mu[j];
    for (int i = 0; i < 8; i++) {
        double s = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) s += innov[j] * K[i][j];
        mean[i] = m[i] + s;
    }
    // cov - K (S K^T)
    double SKt[4][8];
#pragma unroll
    for (int a = 0; a < 4; a++)
        for (int j = 0; j < 8; j++) {
            double s = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) s += S[a][b] * K[j][b];
            SKt[a][j] = s;
        }
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
            double s = 0;
#pragma unroll
            for (int a = 0; a < 4; a++) s += K[i][a] * SKt[a][j];
            cov[i * 8 + j] = P[i * 8 + j] - s;
        }
}

// KalmanFilter.update, one warp per track: lanes 0..7 own the rows of the gain K,
// every lane owns two covariance entries.  Same expression order as the
// single-thread version (and as the oracle): K = P H^T S^-1 by Cholesky solves,
// cov - K (S K^T).
__device__ void kf_update_warp(double *mean, double *cov, const float *xyah, double conf, int lane,
                               double *s_K /* [32] shared */) {
    double mu[4], S[4][4], L[4][4];
    kf_project(mean, cov, conf, mu, S);
    chol4(S, L);
    const double m_old = lane < 8 ? mean[lane] : 0.0;
    double pold[2];
    pold[0] = cov[lane];
    pold[1] = cov[lane + 32];
    if (lane < 8) {
        double y[4], kr[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            double t = cov[lane * 8 + r];
#pragma unroll
            for (int k = 0; k < r; k++) t -= L[r][k] * y[k];
            y[r] = t / L[r][r];
        }
#pragma unroll
        for (int r = 3; r >= 0; r--) {
            double t = y[r];
#pragma unroll
            for (int k = r + 1; k < 4; k++) t -= L[k][r] * kr[k];
            kr[r] = t / L[r][r];
        }
#pragma unroll
        for (int r = 0; r < 4; r++) s_K[lane * 4 + r] = kr[r];
    }
    __syncwarp();
    if (lane < 8) {
        double t = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) t += ((double)xyah[j] - mu[j]) * s_K[lane * 4 + j];
        mean[lane] = m_old + t;
    }
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int e = lane + 32 * r, i = e >> 3, j = e & 7;
        double acc = 0;
#pragma unroll
        for (int a = 0; a < 4; a++) {
            double skt = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) skt += S[a][b] * s_K[j * 4 + b];
            acc += s_K[i * 4 + a] * skt;
        }
        cov[e] = pold[r] - acc;
    }
}

__global__ void kf_update_arrays_kernel(double *mean, double *cov, const float *xyah,
                                        const float *conf, int n) {
    __shared__ double s_K[4][32];
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    kf_update_warp(mean + (size_t)w * 8, cov + (size_t)w * 64, xyah + w * 4, (double)conf[w], lane,
                   s_K[(threadIdx.x >> 5) & 3]);
}

// squared Mahalanobis distance of all measurements to one projected track
__device__ __forceinline__ double maha4(const double L[4][4], const double mu[4],
                                        const float *z) {
    double d0 = (double)z[0] - mu[0], d1 = (double)z[1] - mu[1];
    double d2 = (double)z[2] - mu[2], d3 = (double)z[3] - mu[3];
    double z0 = d0 / L[0][0];
    double z1 = (d1 - L[1][0] * z0) / L[1][1];
    double z2 = (d2 - L[2][0] * z0 - L[2][1] * z1) / L[2][2];
    double z3 = (d3 - L[3][0] * z0 - L[3][1] * z1 - L[3][2] * z2) / L[3][3];
    return z0 * z0 + z1 * z1 + z2 * z2 + z3 * z3;
}

__global__ void kf_gating_arrays_kernel(const double *mean, const double *cov, int n_tracks,
                                        const float *xyah, int n_meas, double *out) {
    const int r = blockIdx.x;
    if (r >= n_tracks) return;
    double mu[4], S[4][4], L[4][4];
    kf_project(mean + (size_t)r * 8, cov + (size_t)r * 64, 0.0, mu, S);
    chol4(S, L);
    for (int j = threadIdx.x; j < n_meas; j += blockDim.x)
        out[(size_t)r * n_meas + j] = maha4(L, mu, xyah + j * 4);
}

// ---------------------------------------------------------------------------
// list building: confirmed / unconfirmed positions in track-list order (A.6)
// ---------------------------------------------------------------------------
__global__ void build_lists_kernel(TrackTable tt, FrameScratch fs) {
    __shared__ int s_w[33];
    const int T = tt.scalars[SC_N_TRACKS];
    const int per = (T + blockDim.x - 1) / blockDim.x;
    const int b = threadIdx.x * per, e = min(T, b + per);
    int c = 0;
    for (int p = b; p < e; p++) c += tt.state[tt.order[p]] == SSB_CONFIRMED;
    int tot;
    int off = block_excl_scan(c, s_w, &tot);
    int offu = b - off;   // unconfirmed before b
    for (int p = b; p < e; p++) {
        if (tt.state[tt.order[p]] == SSB_CONFIRMED) fs.conf_list[off++] = p;
        else fs.unconf_list[offu++] = p;
    }
    if (threadIdx.x == 0) {
        fs.cnt[FC_N_CONF] = tot;
        fs.cnt[FC_N_UNCONF] = T - tot;
    }
}

// ---------------------------------------------------------------------------
// stage-A cost: gate_cost_matrix + clamp (A.6).  One block per confirmed row.
//   g = maha^2 ; cost = app ; cost[g > chi2] = 1e5 ; cost = l*cost + (1-l)*g ;
//   cost[cost > max_dist] = max_dist + 1e-5
// ---------------------------------------------------------------------------
__global__ void gate_cost_kernel(TrackTable tt, FrameScratch fs, SsbDims d, int n) {
    const int r = blockIdx.x;
    if (r >= fs.cnt[FC_N_CONF]) return;
    const int s = tt.order[fs.conf_list[r]];
    double mu[4], S[4][4], L[4][4];
    kf_project(tt.mean + (size_t)s * 8, tt.cov + (size_t)s * 64, 0.0, mu, S);
    chol4(S, L);
    const double clampv = d.max_dist + 1e-5;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double g = maha4(L, mu, fs.det_xyah + j * 4);
        double c = (double)fs.app_cost[(size_t)r * n + j];
        if (g > SSB_CHI2INV95_4) c = SSB_INFTY_COST;
        c = d.mc_lambda * c + d.one_minus_lambda * g;
        if (c > d.max_dist) c = clampv;
        fs.cost_a[(size_t)r * n + j] = c;
    }
}

// ---------------------------------------------------------------------------
// stage-B cost: iou_cost + clamp (A.7).  One block per candidate row.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double iou_cost_one(const double tl[4], const float *c) {
    const double bx = tl[0], by = tl[1], bw = tl[2], bh = tl[3];
    const double bbr_x = bx + bw, bbr_y = by + bh;
    const float cbx = c[0] + c[2], cby = c[1] + c[3];        // float32 adds
    const float carea = c[2] * c[3];                           // float32 product
    const double tlx = fmax(bx, (double)c[0]), tly = fmax(by, (double)c[1]);
    const double brx = fmin(bbr_x, (double)cbx), bry = fmin(bbr_y, (double)cby);
    const double w = fmax(0.0, brx - tlx), h = fmax(0.0, bry - tly);
    const double inter = w * h;
    const double area_b = bw * bh;
    return 1.0 - inter / (area_b + (double)carea - inter);
}

__global__ void iou_cost_kernel(TrackTable tt, FrameScratch fs, SsbDims d) {
    const int r = blockIdx.x;
    const int rows = fs.cnt[FC_N_CAND_B], cols = fs.cnt[FC_N_UNDET_A];
    if (r >= rows) return;
    const int s = tt.order[fs.cand_b[r]];
    const double clampv = d.max_iou + 1e-5;
    const bool stale = tt.tsu[s] > 1;
    double tl[4];
    {   // Track.to_tlwh
        const double *m = tt.mean + (size_t)s * 8;
        double w = m[2] * m[3];
        tl[2] = w; tl[3] = m[3];
        tl[0] = m[0] - w / 2; tl[1] = m[1] - m[3] / 2;
    }
    for (int j = threadIdx.x; j < cols; j += blockDim.x) {
        double c = stale ? SSB_INFTY_COST : iou_cost_one(tl, fs.det_tlwh + fs.undet_a[j] * 4);
        if (c > d.max_iou) c = clampv;
        fs.cost_b[(size_t)r * cols + j] = c;
    }
}

__global__ void iou_cost_arrays_kernel(const double *tlwh, int n_tracks, const float *det_tlwh,
                                       int n_dets, double *out) {
    const int r = blockIdx.x;
    if (r >= n_tracks) return;
    double tl[4] = {tlwh[r * 4], tlwh[r * 4 + 1], tlwh[r * 4 + 2], tlwh[r * 4 + 3]};
    for (int j = threadIdx.x; j < n_dets; j += blockDim.x)
        out[(size_t)r * n_dets + j] = iou_cost_one(tl, det_tlwh + j * 4);
}

// ---------------------------------------------------------------------------
// rectangular LSAP, one warp, scipy tie-breaks (oracle/lsap.c is the C twin)
// ---------------------------------------------------------------------------
struct LsapSmem {
    long long lsap_t_stage, lsap_t_solve;   // clock64 deltas of the last lsap_block (thread 0; debug builds read them)
    long long stats[3];                     // register solver: search steps, cycles in the searches, cycles elsewhere
    double *u, *v, *spc, *cost;   // cost == nullptr -> read global
    int *path, *row4col, *col4row, *remaining;
    unsigned char *SR, *SC;
};

__device__ __forceinline__ size_t lsap_smem_fixed_bytes(int L) {
    // u, v, spc: 3*L doubles; path,row4col,col4row,remaining: 4*L ints; SR,SC: 2*L bytes
    return (size_t)L * (3 * 8 + 4 * 4 + 2) + 64;
}

__device__ void lsap_carve(unsigned char *base, int L, LsapSmem &m) {
    m.u = (double *)base;            base += (size_t)L * 8;
    m.v = (double *)base;            base += (size_t)L * 8;
    m.spc = (double *)base;          base += (size_t)L * 8;
    m.path = (int *)base;            base += (size_t)L * 4;
    m.row4col = (int *)base;         base += (size_t)L * 4;
    m.col4row = (int *)base;         base += (size_t)L * 4;
    m.remaining = (int *)base;       base += (size_t)L * 4;
    m.SR = base;                     base += L;
    m.SC = base;                     base += L;
    base = (unsigned char *)(((uintptr_t)base + 15) & ~(uintptr_t)15);
    m.cost = (double *)base;
}

// Register-resident shortest-augmenting-path core for nc <= 32*CPL columns: lane l
// owns columns l, l+32, ... (v, shortest-path cost, path, row4col and the column's
// position in scipy's `remaining[]` all live in registers), so a search step is CPL
// independent cost loads + three warp REDUX ops instead of a shared-memory walk.
// scipy's scan-order tie-break is reproduced through `pos` (the column's index in
// remaining[]): among equal costs an unassigned column with the LARGEST position
// wins, else the SMALLEST position; swap-removal moves the last column into the
// freed position.  (tests emulate this against scipy; oracle/lsap.c is the serial twin.)
__device__ __forceinline__ unsigned long long f64_order_key(double x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// SPARSE: the matrix is `bg` everywhere except the entries listed per working row in shared memory
// (rowptr / rowlen / ecol / eval) -- the clamped cost matrices of min_cost_matching are >99 % one value
// (max_d + 1e-5), so a search step needs no global loads at all; cij is the same double either way,
// hence bit-identical results.
struct LsapSparse {
    double bg;
    const int *rowptr, *rowlen, *ecol;
    const double *eval;
};

// One search step's per-lane work, branch-free: the CPL columns of a lane are independent, and written as guarded
// blocks the compiler serialises them (four BSSY regions of LDS -> 3 DADD -> DSETP chains: ~110 cycles each, the
// step's critical path).  Here every column's reduced cost is computed unconditionally (in-range address for the
// columns past nc), the updates are selects, and the lane's arg-min (value ascending, tie key descending -- keys are
// unique per column, so the order of comparison does not matter) is a tree.
template <int CPL>
__device__ __forceinline__ void lsap_lane_best(const double (&cand)[CPL], const unsigned (&ckey)[CPL],
                                               double &bv, unsigned &bkey, int &bk) {
    if (CPL == 4) {
        const bool t01 = cand[1] < cand[0] || (cand[1] == cand[0] && ckey[1] > ckey[0]);
        const bool t23 = cand[3] < cand[2] || (cand[3] == cand[2] && ckey[3] > ckey[2]);
        const double v01 = t01 ? cand[1] : cand[0], v23 = t23 ? cand[3] : cand[2];
        const unsigned k01 = t01 ? ckey[1] : ckey[0], k23 = t23 ? ckey[3] : ckey[2];
        const bool t = v23 < v01 || (v23 == v01 && k23 > k01);
        bv = t ? v23 : v01;
        bkey = t ? k23 : k01;
        bk = t ? (t23 ? 3 : 2) : (t01 ? 1 : 0);
    } else {
        bv = cand[0]; bkey = ckey[0]; bk = 0;
#pragma unroll
        for (int k = 1; k < CPL; k++) {
            const bool t = cand[k] < bv || (cand[k] == bv && ckey[k] > bkey);
            bv = t ? cand[k] : bv; bkey = t ? ckey[k] : bkey; bk = t ? k : bk;
        }
    }
    if (bkey == 0u) bk = -1;                        // no active column on this lane (active keys are > 0)
}

template <int CPL, bool SPARSE>
__device__ bool lsap_warp_reg(const double *__restrict__ C, bool staged, bool tr, int nc0, int nr,
                              int nc, double *u, int *col4row, int *row4col_out, int lane,
                              LsapSparse sp = LsapSparse(), long long *stats = nullptr) {
    long long st_steps = 0, st_search = 0, st_rest = 0, st_t = stats ? clock64() : 0;
    double v[CPL], spc[CPL];
    int pos[CPL], r4c[CPL], path[CPL];
    unsigned sc = 0;
#pragma unroll
    for (int k = 0; k < CPL; k++) { v[k] = 0.0; r4c[k] = -1; path[k] = -1; }
    for (int curRow = 0; curRow < nr; curRow++) {
        sc = 0;
#pragma unroll
        for (int k = 0; k < CPL; k++) {
            const int j = lane + 32 * k;
            spc[k] = INFINITY;
            pos[k] = j < nc ? nc - 1 - j : -1;
        }
        double minVal = 0.0;
        int i = curRow, num_remaining = nc, sink = -1;
        if (stats) { const long long t = clock64(); st_rest += t - st_t; st_t = t; }
        while (sink == -1) {
            st_steps++;
            const double ui = u[i];
            double bv;
            unsigned bkey;
            int bk;
            double cs[SPARSE ? CPL : 1];
            if (SPARSE) {
#pragma unroll
                for (int k = 0; k < CPL; k++) cs[k] = sp.bg;
                const int e0 = sp.rowptr[i], e1 = e0 + sp.rowlen[i];
                for (int e = e0; e < e1; e++) {                  // warp-uniform trip count
                    const int cj = sp.ecol[e];
                    const double cv = sp.eval[e];
                    const bool mine = lane == (cj & 31);
#pragma unroll
                    for (int k = 0; k < CPL; k++) cs[k] = (mine && k == (cj >> 5)) ? cv : cs[k];     // selects, not a dynamic index
                }
            }
            double cand[CPL];
            unsigned ckey[CPL];
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                const int j = lane + 32 * k;
                const bool act = j < nc && !((sc >> k) & 1u);
                const int jj = j < nc ? j : 0;
                const double cij = SPARSE ? cs[SPARSE ? k : 0]
                                          : (staged ? C[(size_t)i * nc + jj]
                                                    : (tr ? C[(size_t)jj * nc0 + i] : C[(size_t)i * nc0 + jj]));
                const double r = minVal + cij - ui - v[k];
                const bool upd = act && r < spc[k];
                path[k] = upd ? i : path[k];
                spc[k] = upd ? r : spc[k];
                const unsigned key = r4c[k] == -1 ? (0x40000000u | (unsigned)pos[k]) : (unsigned)(nc - pos[k]);
                cand[k] = act ? spc[k] : INFINITY;
                ckey[k] = act ? key : 0u;
            }
            lsap_lane_best<CPL>(cand, ckey, bv, bkey, bk);
            // warp arg-min: value (two 32-bit REDUX over an order-preserving key), then tie key
            const unsigned long long ok = f64_order_key(bv);
            const unsigned hi = (unsigned)(ok >> 32), lo = (unsigned)ok;
            const unsigned hmin = __reduce_min_sync(0xffffffffu, bk >= 0 ? hi : 0xffffffffu);
            bool vwin = bk >= 0 && hi == hmin;
            unsigned wmask = __ballot_sync(0xffffffffu, vwin);
            if (wmask == 0) return false;
            if (wmask & (wmask - 1)) {              // several lanes share the top word: compare the low word ...
                const unsigned lmin = __reduce_min_sync(0xffffffffu, vwin ? lo : 0xffffffffu);
                vwin = vwin && lo == lmin;
                wmask = __ballot_sync(0xffffffffu, vwin);
                if (wmask & (wmask - 1)) {          // ... and on equal values scipy's scan-order rule
                    const unsigned kmax = __reduce_max_sync(0xffffffffu, vwin ? bkey : 0u);
                    wmask = __ballot_sync(0xffffffffu, vwin && bkey == kmax);
                }
            }
            const int wl = __ffs(wmask) - 1;
            // every lane prepares the (position, row) of ITS best column, so the four broadcasts from the winning lane
            // are independent (one shuffle round instead of two dependent ones)
            int my_pos = -1, my_r = -1;
#pragma unroll
            for (int k = 0; k < CPL; k++) { my_pos = k == bk ? pos[k] : my_pos; my_r = k == bk ? r4c[k] : my_r; }
            const int wk = __shfl_sync(0xffffffffu, bk, wl);
            const int index = __shfl_sync(0xffffffffu, my_pos, wl);
            const int rj = __shfl_sync(0xffffffffu, my_r, wl);
            minVal = __shfl_sync(0xffffffffu, bv, wl);
            if (!(minVal < INFINITY)) return false;                 // NaN / inf costs: infeasible
            const int jwin = wl + 32 * wk;
            if (lane == wl) sc |= 1u << wk;
#pragma unroll
            for (int k = 0; k < CPL; k++)
                if (!((sc >> k) & 1u) && pos[k] == num_remaining - 1) pos[k] = index;
            num_remaining--;
            if (rj == -1) sink = jwin; else i = rj;
        }
        if (stats) { const long long t = clock64(); st_search += t - st_t; st_t = t; }
        // dual updates (column-wise: visited column j with row r = row4col[j] gives u[r])
        if (lane == 0) u[curRow] += minVal;
        {   // the CPL columns' updates interleaved (loads first, then the stores); rows of distinct columns are distinct
            double nu[CPL];
            bool wr[CPL];
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                const bool vis = (sc >> k) & 1u;
                const double dlt = minVal - spc[k];
                wr[k] = vis && r4c[k] >= 0;
                nu[k] = u[wr[k] ? r4c[k] : curRow] + dlt;
                v[k] = vis ? v[k] - dlt : v[k];
            }
#pragma unroll
            for (int k = 0; k < CPL; k++)
                if (wr[k]) u[r4c[k]] = nu[k];
        }
        __syncwarp();
        // augment along the path
        int j = sink;
        while (true) {
            const int ow = j & 31, kk = j >> 5;
            int mine = -1;
#pragma unroll
            for (int k = 0; k < CPL; k++)
                if (k == kk) mine = path[k];
            const int r = __shfl_sync(0xffffffffu, mine, ow);
            if (lane == ow) {
#pragma unroll
                for (int k = 0; k < CPL; k++)
                    if (k == kk) r4c[k] = r;
            }
            const int t = col4row[r];
            __syncwarp();
            if (lane == 0) col4row[r] = j;
            __syncwarp();
            j = t;
            if (r == curRow) break;
        }
    }
#pragma unroll
    for (int k = 0; k < CPL; k++) {
        const int j = lane + 32 * k;
        if (j < nc) row4col_out[j] = r4c[k];
    }
    __syncwarp();
    if (stats && lane == 0) { stats[0] = st_steps; stats[1] = st_search; stats[2] = st_rest + (clock64() - st_t); }
    return true;
}

// The same solver spread over NW warps for wide problems (config C4 once more than 512 tracks are alive): 4 columns
// per lane, 128 per warp, up to 8 warps = 1024 columns.  A search step is the single-warp step on each warp's own
// 128 columns, one shared-memory exchange of the NW warp winners (double-buffered, one named barrier of the NW
// warps) and the same "lowest value, then largest tie key" rule applied to them -- the key is a function of the
// column's global position in scipy's remaining[], so the combined choice is the serial algorithm's choice.
// Control flow is replicated in every warp (they all see the same winner); u[] and col4row[] live in shared memory.
struct LsapXch {                 // one warp's winner of a search step
    double bv;
    unsigned bkey;
    int j, pos, r4c;
};
__device__ __forceinline__ void lsap_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

template <int NW, bool SPARSE, int CPL = 4>
__device__ bool lsap_block_reg(const double *__restrict__ C, bool tr, int nc0, int nr, int nc, double *u, int *col4row,
                               int *row4col_out, LsapXch *xch /* [2][NW] shared */, int *s_aug /* [2] shared */,
                               LsapSparse sp = LsapSparse()) {
    constexpr int WC = 32 * CPL;                 // columns per warp (CPL = 8: 1025 .. 2048 working columns on 8 warps)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, base = wid * WC;
    double v[CPL], spc[CPL];
    int pos[CPL], r4c[CPL], path[CPL];
    unsigned sc = 0;
    int xb = 0;                                   // exchange buffer parity
#pragma unroll
    for (int k = 0; k < CPL; k++) { v[k] = 0.0; r4c[k] = -1; path[k] = -1; }
    for (int curRow = 0; curRow < nr; curRow++) {
        sc = 0;
#pragma unroll
        for (int k = 0; k < CPL; k++) {
            const int j = base + lane + 32 * k;
            spc[k] = INFINITY;
            pos[k] = j < nc ? nc - 1 - j : -1;
        }
        double minVal = 0.0;
        int i = curRow, num_remaining = nc, sink = -1;
        while (sink == -1) {
            const double ui = u[i];
            double bv;
            unsigned bkey;
            int bk;
            double cs[SPARSE ? CPL : 1];
            if (SPARSE) {
#pragma unroll
                for (int k = 0; k < CPL; k++) cs[k] = sp.bg;
                const int e0 = sp.rowptr[i], e1 = e0 + sp.rowlen[i];
                for (int e = e0; e < e1; e++) {                  // block-uniform trip count
                    const int cj = sp.ecol[e] - base;
                    const double cv = sp.eval[e];
                    const bool mine = cj >= 0 && cj < WC && lane == (cj & 31);
#pragma unroll
                    for (int k = 0; k < CPL; k++) cs[k] = (mine && k == (cj >> 5)) ? cv : cs[k];     // selects, not a dynamic index
                }
            }
            double cand[CPL];
            unsigned ckey[CPL];
#pragma unroll
            for (int k = 0; k < CPL; k++) {                      // branch-free, as in lsap_warp_reg
                const int j = base + lane + 32 * k;
                const bool act = j < nc && !((sc >> k) & 1u);
                const int jj = j < nc ? j : 0;
                const double cij = SPARSE ? cs[SPARSE ? k : 0]
                                          : (tr ? C[(size_t)jj * nc0 + i] : C[(size_t)i * nc0 + jj]);
                const double r = minVal + cij - ui - v[k];
                const bool upd = act && r < spc[k];
                path[k] = upd ? i : path[k];
                spc[k] = upd ? r : spc[k];
                const unsigned key = r4c[k] == -1 ? (0x40000000u | (unsigned)pos[k]) : (unsigned)(nc - pos[k]);
                cand[k] = act ? spc[k] : INFINITY;
                ckey[k] = act ? key : 0u;
            }
            lsap_lane_best<CPL>(cand, ckey, bv, bkey, bk);
            // this warp's winner (value, then tie key), as in lsap_warp_reg
            const unsigned long long ok = f64_order_key(bv);
            const unsigned hi = (unsigned)(ok >> 32), lo = (unsigned)ok;
            const unsigned hmin = __reduce_min_sync(0xffffffffu, bk >= 0 ? hi : 0xffffffffu);
            bool vwin = bk >= 0 && hi == hmin;
            unsigned wmask = __ballot_sync(0xffffffffu, vwin);
            if (wmask & (wmask - 1)) {              // several lanes share the top word: low word, then the tie key
                const unsigned lmin = __reduce_min_sync(0xffffffffu, vwin ? lo : 0xffffffffu);
                vwin = vwin && lo == lmin;
                wmask = __ballot_sync(0xffffffffu, vwin);
                if (wmask & (wmask - 1)) {
                    const unsigned kmax = __reduce_max_sync(0xffffffffu, vwin ? bkey : 0u);
                    wmask = __ballot_sync(0xffffffffu, vwin && bkey == kmax);
                }
            }
            const int wl = wmask ? __ffs(wmask) - 1 : -1;
            int my_pos = -1, my_r = -1;
#pragma unroll
            for (int k = 0; k < CPL; k++)
                if (k == bk) { my_pos = pos[k]; my_r = r4c[k]; }
            if (lane == (wl >= 0 ? wl : 0)) {
                LsapXch x;
                x.bv = wl >= 0 ? bv : INFINITY;
                x.bkey = wl >= 0 ? bkey : 0u;
                x.j = wl >= 0 ? base + lane + 32 * bk : -1;
                x.pos = my_pos; x.r4c = my_r;
                xch[xb * NW + wid] = x;
            }
            lsap_bar(32 * NW);
            // every warp combines the NW winners identically
            LsapXch g = xch[xb * NW];
#pragma unroll
            for (int w = 1; w < NW; w++) {
                const LsapXch o = xch[xb * NW + w];
                const bool take = o.j >= 0 && (g.j < 0 || o.bv < g.bv || (o.bv == g.bv && o.bkey > g.bkey));
                if (take) g = o;
            }
            xb ^= 1;
            if (g.j < 0) return false;
            minVal = g.bv;
            if (!(minVal < INFINITY)) return false;                 // NaN / inf costs: infeasible
            const int jwin = g.j, index = g.pos, rj = g.r4c;
            if (jwin - base == lane + 32 * ((jwin - base) >> 5) && jwin >= base && jwin < base + WC)
                sc |= 1u << ((jwin - base) >> 5);
#pragma unroll
            for (int k = 0; k < CPL; k++)
                if (!((sc >> k) & 1u) && pos[k] == num_remaining - 1) pos[k] = index;
            num_remaining--;
            if (rj == -1) sink = jwin; else i = rj;
        }
        // dual updates: every lane for its own visited columns (distinct rows of u)
        if (threadIdx.x == 0) u[curRow] += minVal;
#pragma unroll
        for (int k = 0; k < CPL; k++) {
            if ((sc >> k) & 1u) {
                const double dlt = minVal - spc[k];
                if (r4c[k] >= 0) u[r4c[k]] += dlt;
                v[k] -= dlt;
            }
        }
        lsap_bar(32 * NW);
        // augment along the path: the owner of column j publishes path[j]; col4row is read by all, then written by one
        int j = sink;
        while (true) {
            const int lj = j - base;
            const bool mine = lj >= 0 && lj < WC && (lj & 31) == lane;
            if (mine) {
                int r = -1;
#pragma unroll
                for (int k = 0; k < CPL; k++)
                    if (k == (lj >> 5)) { r = path[k]; r4c[k] = r; }
                s_aug[0] = r;
            }
            lsap_bar(32 * NW);
            const int r = s_aug[0];
            const int t = col4row[r];
            lsap_bar(32 * NW);
            if (threadIdx.x == 0) col4row[r] = j;
            j = t;
            if (r == curRow) break;
        }
        lsap_bar(32 * NW);
    }
#pragma unroll
    for (int k = 0; k < CPL; k++) {
        const int j = base + lane + 32 * k;
        if (j < nc) row4col_out[j] = r4c[k];
    }
    return true;
}

// Solve with the whole block staging, warp 0 iterating.  C is [nr0][nc0]
// row-major (ld = nc0).  Results in ORIGINAL orientation: col4row_out[nr0],
// row4col_out[nc0] (-1 = unassigned).  Must be called by all threads.
__device__ void lsap_block(const double *__restrict__ C, int nr0, int nc0, LsapSmem &m,
                           size_t cost_smem_bytes, int *col4row_out, int *row4col_out) {
    const int tid = threadIdx.x, lane = tid & 31;
    const long long t_l0 = clock64();
    m.lsap_t_stage = m.lsap_t_solve = 0;
    if (nr0 == 0 || nc0 == 0) {
        for (int i = tid; i < nr0; i += blockDim.x) col4row_out[i] = -1;
        for (int j = tid; j < nc0; j += blockDim.x) row4col_out[j] = -1;
        __syncthreads();
        return;
    }
    const bool tr = nc0 < nr0;
    const int nr = tr ? nc0 : nr0, nc = tr ? nr0 : nc0;
    const bool staged = (size_t)nr * nc * 8 <= cost_smem_bytes;
    if (staged) {
        // eight loads in flight per thread: the matrix was just written by the cost kernels (L2), and a one-load-per-
        // iteration loop spent 20 K cycles (11 us) on round trips for config C2's 78 KB
        const int total = nr0 * nc0, step = blockDim.x;
        for (int e0 = tid; e0 < total; e0 += 8 * step) {
            double c[8];
#pragma unroll
            for (int q = 0; q < 8; q++) { const int e = e0 + q * step; c[q] = e < total ? C[e] : 0.0; }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int e = e0 + q * step;
                if (e < total) {
                    const int i0 = e / nc0, j0 = e - i0 * nc0;
                    if (tr) m.cost[(size_t)j0 * nc + i0] = c[q]; else m.cost[e] = c[q];
                }
            }
        }
    }
    for (int i = tid; i < nr; i += blockDim.x) { m.u[i] = 0.0; m.col4row[i] = -1; }
    for (int j = tid; j < nc; j += blockDim.x) { m.v[j] = 0.0; m.row4col[j] = -1; m.path[j] = -1; }
    __syncthreads();
    m.lsap_t_stage = clock64() - t_l0;

    // Too big for shared memory (C4: 344 x 498): every search step would wait for an L2 round trip.
    // The clamped matrices are one background value (their maximum) plus a few real entries per row,
    // so keep only those in shared memory (the register solver's SPARSE mode).  m.remaining / m.path
    // (unused by the register path) hold the per-row offsets / lengths; falls back to the dense
    // global-memory path when the real entries do not fit.
    LsapSparse sp;
    bool sparse = false;
    if (!staged && nc <= 2048) {
        __shared__ double s_bg[8];
        __shared__ int s_total;
        const int wid = tid >> 5, nwarp = blockDim.x >> 5;
        // All three passes keep EIGHT loads in flight per thread: the matrix sits in L2 (the cost kernels just wrote
        // it), and with one load per iteration each pass cost (elements / 256) x one L2 round trip -- ~0.3 ms per pass
        // at C4's 600 x 496, three passes, half of the stage.
        double mx = -INFINITY;
        {
            const int total = nr0 * nc0, step = blockDim.x;
            for (int e0 = tid; e0 < total; e0 += 8 * step) {
                double c[8];
#pragma unroll
                for (int q = 0; q < 8; q++) { const int e = e0 + q * step; c[q] = e < total ? C[e] : -INFINITY; }
#pragma unroll
                for (int q = 0; q < 8; q++) if (c[q] > mx) mx = c[q];
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, mx, o); if (t > mx) mx = t; }
        if (lane == 0) s_bg[wid] = mx;
        __syncthreads();
        mx = s_bg[0];
        for (int w = 1; w < nwarp; w++) if (s_bg[w] > mx) mx = s_bg[w];
        const size_t cap = cost_smem_bytes / 12;
        double *evals = m.cost;
        int *ecols = (int *)(m.cost + cap);
        int *rowptr = m.remaining, *rowlen = m.path;
        // pass 1: entries != bg per working row.  Not transposed: warp per row, lanes across the (contiguous) columns.
        // Transposed (working row = original column): lane per working row, so that a warp's 32 loads of one original
        // row are contiguous -- lanes across columns would touch 32 cache lines per load.
        if (!tr) {
            for (int i = wid; i < nr; i += nwarp) {
                int cnt = 0;
                for (int j0 = 0; j0 < nc; j0 += 256) {
                    double c[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) { const int j = j0 + 32 * q + lane; c[q] = j < nc ? C[(size_t)i * nc0 + j] : mx; }
#pragma unroll
                    for (int q = 0; q < 8; q++) cnt += __popc(__ballot_sync(0xffffffffu, !(c[q] == mx)));
                }
                if (lane == 0) rowlen[i] = cnt;
            }
        } else {
            for (int i0 = wid * 32; i0 < nr; i0 += nwarp * 32) {
                const int i = i0 + lane;
                int cnt = 0;
                for (int j0 = 0; j0 < nc; j0 += 8) {
                    double c[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) c[q] = (i < nr && j0 + q < nc) ? C[(size_t)(j0 + q) * nc0 + i] : mx;
#pragma unroll
                    for (int q = 0; q < 8; q++) cnt += !(c[q] == mx);
                }
                if (i < nr) rowlen[i] = cnt;
            }
        }
        __syncthreads();
        if (tid == 0) {                                    // serial exclusive scan (nr <= 1024)
            int acc = 0;
            for (int i = 0; i < nr; i++) { rowptr[i] = acc; acc += rowlen[i]; }
            s_total = acc;
        }
        __syncthreads();
        if ((size_t)s_total <= cap) {
            if (!tr) {
                for (int i = wid; i < nr; i += nwarp) {        // pass 2: fill, column order
                    int off = rowptr[i];
                    for (int j0 = 0; j0 < nc; j0 += 256) {
                        double c[8];
#pragma unroll
                        for (int q = 0; q < 8; q++) { const int j = j0 + 32 * q + lane; c[q] = j < nc ? C[(size_t)i * nc0 + j] : mx; }
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            const bool nz = !(c[q] == mx);
                            const unsigned bal = __ballot_sync(0xffffffffu, nz);
                            if (nz) { const int w = off + __popc(bal & ((1u << lane) - 1)); evals[w] = c[q]; ecols[w] = j0 + 32 * q + lane; }
                            off += __popc(bal);
                        }
                    }
                }
            } else {
                for (int i0 = wid * 32; i0 < nr; i0 += nwarp * 32) {
                    const int i = i0 + lane;
                    int off = i < nr ? rowptr[i] : 0;
                    for (int j0 = 0; j0 < nc; j0 += 8) {
                        double c[8];
#pragma unroll
                        for (int q = 0; q < 8; q++) c[q] = (i < nr && j0 + q < nc) ? C[(size_t)(j0 + q) * nc0 + i] : mx;
#pragma unroll
                        for (int q = 0; q < 8; q++)
                            if (!(c[q] == mx)) { evals[off] = c[q]; ecols[off] = j0 + q; off++; }
                    }
                }
            }
            sparse = true;
            sp.bg = mx; sp.rowptr = rowptr; sp.rowlen = rowlen; sp.ecol = ecols; sp.eval = evals;
        }
        __syncthreads();
    }

    __shared__ LsapXch s_xch[2 * 8];
    __shared__ int s_aug[2];
    if (nc <= 128) {
        if (tid < 32) {                       // one warp, everything in registers (config C2)
            const double *Cw = staged ? m.cost : C;
            bool okr;
            if (sparse) okr = lsap_warp_reg<4, true>(Cw, staged, tr, nc0, nr, nc, m.u, m.col4row, m.row4col, lane, sp);
            else {
#ifdef SSB_BASELINES
                okr = lsap_warp_reg<4, false>(Cw, staged, tr, nc0, nr, nc, m.u, m.col4row, m.row4col, lane, LsapSparse(), m.stats);
#else
                okr = lsap_warp_reg<4, false>(Cw, staged, tr, nc0, nr, nc, m.u, m.col4row, m.row4col, lane);
#endif
            }
            if (!okr) {
                for (int i = lane; i < nr; i += 32) m.col4row[i] = -1;
                for (int j = lane; j < nc; j += 32) m.row4col[j] = -1;
            }
        }
    } else if (nc <= 2048) {
        // 129 .. 1024 working columns: 4 columns per lane on 2 / 4 / 8 warps (lsap_block_reg); up to 2048: 8 per lane
        // (a crowded stream confirms short-lived tracks by chance overlaps and keeps them max_age frames: C4 passes
        // 1024 confirmed tracks after ~25 frames, and the dense global path behind this branch costs 10 ms a frame)
        const int nwarp_s = nc <= 256 ? 2 : (nc <= 512 ? 4 : 8);
        if (tid < 32 * nwarp_s) {
            const double *Cw = staged ? m.cost : C;
            const bool trw = staged ? false : tr;
            const int ldw = staged ? nc : nc0;
            bool okr;
            if (nwarp_s == 2)
                okr = sparse ? lsap_block_reg<2, true>(Cw, trw, ldw, nr, nc, m.u, m.col4row, m.row4col, s_xch, s_aug, sp)
                             : lsap_block_reg<2, false>(Cw, trw, ldw, nr, nc, m.u, m.col4row, m.row4col, s_xch, s_aug);
            else if (nwarp_s == 4)
                okr = sparse ? lsap_block_reg<4, true>(Cw, trw, ldw, nr, nc, m.u, m.col4row, m.row4col, s_xch, s_aug, sp)
                             : lsap_block_reg<4, false>(Cw, trw, ldw, nr, nc, m.u, m.col4row, m.row4col, s_xch, s_aug);
            else if (nc <= 1024)
                okr = sparse ? lsap_block_reg<8, true>(Cw, trw, ldw, nr, nc, m.u, m.col4row, m.row4col, s_xch, s_aug, sp)
                             : lsap_block_reg<8, false>(Cw, trw, ldw, nr, nc, m.u, m.col4row, m.row4col, s_xch, s_aug);
            else
                okr = sparse ? lsap_block_reg<8, true, 8>(Cw, trw, ldw, nr, nc, m.u, m.col4row, m.row4col, s_xch, s_aug, sp)
                             : lsap_block_reg<8, false, 8>(Cw, trw, ldw, nr, nc, m.u, m.col4row, m.row4col, s_xch, s_aug);
            if (!okr) {
                for (int i = tid; i < nr; i += 32 * nwarp_s) m.col4row[i] = -1;
                for (int j = tid; j < nc; j += 32 * nwarp_s) m.row4col[j] = -1;
            }
        }
    } else if (tid < 32) {
        bool failed = false;
        for (int curRow = 0; curRow < nr && !failed; curRow++) {
            for (int j = lane; j < nc; j += 32) {
                m.remaining[j] = nc - j - 1;
                m.spc[j] = INFINITY;
                m.SC[j] = 0;
            }
            for (int i = lane; i < nr; i += 32) m.SR[i] = 0;
            __syncwarp();
            double minVal = 0.0;
            int i = curRow, num_remaining = nc, sink = -1;
            while (sink == -1) {
                if (lane == 0) m.SR[i] = 1;
                const double ui = m.u[i];
                double best_v = INFINITY;
                int best_it = -1, best_un = 0;
                for (int it = lane; it < num_remaining; it += 32) {
                    const int j = m.remaining[it];
                    double cij;
                    if (staged) cij = m.cost[(size_t)i * nc + j];
                    else cij = tr ? C[(size_t)j * nc0 + i] : C[(size_t)i * nc0 + j];
                    const double r = minVal + cij - ui - m.v[j];
                    double sj = m.spc[j];
                    if (r < sj) { m.path[j] = i; m.spc[j] = r; sj = r; }
                    const int un = m.row4col[j] == -1;
                    if (sj < best_v || (sj == best_v && un)) { best_v = sj; best_it = it; best_un = un; }
                }
                // warp arg-min with scipy's rule: lower value; on ties an unassigned
                // column wins (the LAST such in scan order), else the FIRST scanned.
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best_v, o);
                    const int oit = __shfl_xor_sync(0xffffffffu, best_it, o);
                    const int oun = __shfl_xor_sync(0xffffffffu, best_un, o);
                    bool take;
                    if (oit < 0) take = false;
                    else if (best_it < 0) take = true;
                    else if (ov != best_v) take = ov < best_v;
                    else if (oun != best_un) take = oun > best_un;
                    else take = oun ? (oit > best_it) : (oit < best_it);
                    if (take) { best_v = ov; best_it = oit; best_un = oun; }
                }
                if (best_it < 0 || !(best_v < INFINITY)) {   // NaN/inf costs: infeasible
                    failed = true;
                    break;
                }
                minVal = best_v;
                const int j = m.remaining[best_it];
                const int rj = m.row4col[j];
                __syncwarp();
                if (lane == 0) {
                    m.SC[j] = 1;
                    m.remaining[best_it] = m.remaining[num_remaining - 1];
                }
                num_remaining--;
                if (rj == -1) sink = j; else i = rj;
                __syncwarp();
            }
            if (failed) break;
            // dual updates
            if (lane == 0) m.u[curRow] += minVal;
            for (int r = lane; r < nr; r += 32)
                if (m.SR[r] && r != curRow) m.u[r] += minVal - m.spc[m.col4row[r]];
            for (int j = lane; j < nc; j += 32)
                if (m.SC[j]) m.v[j] -= minVal - m.spc[j];
            __syncwarp();
            if (lane == 0) {           // augment
                int j = sink;
                while (true) {
                    const int r = m.path[j];
                    m.row4col[j] = r;
                    const int t = m.col4row[r];
                    m.col4row[r] = j;
                    j = t;
                    if (r == curRow) break;
                }
            }
            __syncwarp();
        }
        if (failed) {   // leave everything unassigned; callers treat rows as unmatched
            for (int i = lane; i < nr; i += 32) m.col4row[i] = -1;
            for (int j = lane; j < nc; j += 32) m.row4col[j] = -1;
        }
    }
    __syncthreads();
    m.lsap_t_solve = clock64() - t_l0;
    if (!tr) {
        for (int i = tid; i < nr0; i += blockDim.x) col4row_out[i] = m.col4row[i];
        for (int j = tid; j < nc0; j += blockDim.x) row4col_out[j] = m.row4col[j];
    } else {   // working rows == original columns
        for (int i = tid; i < nr0; i += blockDim.x) col4row_out[i] = m.row4col[i];
        for (int j = tid; j < nc0; j += blockDim.x) row4col_out[j] = m.col4row[j];
    }
    __syncthreads();
}

__global__ void lsap_kernel(const double *C, int nr, int nc, int L, size_t cost_smem_bytes,
                            int *col4row_out, int *row4col_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    LsapSmem m;
    lsap_carve(smem, L, m);
    lsap_block(C, nr, nc, m, cost_smem_bytes, col4row_out, row4col_out);
}

// Ordered multi-way compaction helper: every thread owns a contiguous chunk of
// [0, n) and classifies each index into one of NCLS lists (or none, cls < 0);
// emit(cls, position_in_that_list, index) is called in index order per list.
template <int NCLS, typename Classify, typename Emit>
__device__ void block_partition(int n, int *s_w, int *totals, Classify classify, Emit emit) {
    const int per = (n + blockDim.x - 1) / blockDim.x;
    const int b = threadIdx.x * per, e = min(n, b + per);
    int cnt[NCLS];
#pragma unroll
    for (int k = 0; k < NCLS; k++) cnt[k] = 0;
    for (int i = b; i < e; i++) {
        const int c = classify(i);
#pragma unroll
        for (int k = 0; k < NCLS; k++) cnt[k] += (c == k);
    }
    int off[NCLS];
#pragma unroll
    for (int k = 0; k < NCLS; k++) off[k] = block_excl_scan(cnt[k], s_w, &totals[k]);
    for (int i = b; i < e; i++) {
        const int c = classify(i);
#pragma unroll
        for (int k = 0; k < NCLS; k++)
            if (c == k) emit(k, off[k]++, i);
    }
}

// Stage A: LSAP over confirmed x dets, then min_cost_matching's list building
// and the stage-B candidate list (A.6), all lists by ordered block compaction.
__global__ void assign_stage_a_kernel(TrackTable tt, FrameScratch fs, SsbDims d, int n, int L,
                                      size_t cost_smem_bytes) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_w[33];
    __shared__ int s_nu0;
    LsapSmem m;
    lsap_carve(smem, L, m);
    const int rows = fs.cnt[FC_N_CONF], cols = n;
#ifdef SSB_BASELINES            // phase stamps (libssb_dbg.so, tools/time_stages.py): cnt[20..23] = start, staged, solved, lists
    const long long t_a0 = clock64();
#endif
    lsap_block(fs.cost_a, rows, cols, m, cost_smem_bytes, fs.col4row, fs.row4col);
#ifdef SSB_BASELINES
    const long long t_a1 = clock64();
#endif
    const int n_unconf = fs.cnt[FC_N_UNCONF];
    // unmatched detections, part 1: unassigned columns ascending
    {
        int tot[1];
        block_partition<1>(cols, s_w, tot,
                           [&](int c) { return fs.row4col[c] < 0 ? 0 : -1; },
                           [&](int, int pos, int c) { fs.undet_a[pos] = c; });
        if (threadIdx.x == 0) s_nu0 = tot[0];
    }
    for (int k = threadIdx.x; k < n_unconf; k += blockDim.x) fs.cand_b[k] = fs.unconf_list[k];
    __syncthreads();
    const int nu0 = s_nu0;
    // rows in order: 0 = match, 1 = rejected pair (cost > max_dist), then the
    // unmatched track goes to 2 = stage-B candidate (tsu == 1) or 3 = kept unmatched
    auto row_class = [&](int r) -> int {          // 0 match, 1 reject+cand, 2 reject+keep, 3 un+cand, 4 un+keep
        const int c = fs.col4row[r];
        const bool one = tt.tsu[tt.order[fs.conf_list[r]]] == 1;
        if (c >= 0) {
            if (fs.cost_a[(size_t)r * cols + c] > d.max_dist) return one ? 1 : 2;
            return 0;
        }
        return one ? 3 : 4;
    };
    int tot3[3];
    block_partition<3>(rows, s_w, tot3,
                       [&](int r) { const int k = row_class(r); return k == 0 ? 0 : ((k == 1 || k == 3) ? 1 : 2); },
                       [&](int cls, int pos, int r) {
                           const int p = fs.conf_list[r];
                           if (cls == 0) { fs.match_trk[pos] = p; fs.match_det[pos] = fs.col4row[r]; }
                           else if (cls == 1) fs.cand_b[n_unconf + pos] = p;
                           else fs.untrk_a_keep[pos] = p;
                       });
    int totr[1];
    block_partition<1>(rows, s_w, totr,
                       [&](int r) { const int k = row_class(r); return (k == 1 || k == 2) ? 0 : -1; },
                       [&](int, int pos, int r) { fs.undet_a[nu0 + pos] = fs.col4row[r]; });
    if (threadIdx.x == 0) {
        fs.cnt[FC_N_UNDET_A] = nu0 + totr[0];
        fs.cnt[FC_N_CAND_B] = n_unconf + tot3[1];
        fs.cnt[FC_N_UNTRK_A_KEEP] = tot3[2];
        fs.cnt[FC_N_MATCH] = tot3[0];
        fs.cnt[FC_N_MATCH_A] = tot3[0];
        fs.cnt[FC_ROWS_A] = rows; fs.cnt[FC_COLS_A] = cols;
#ifdef SSB_BASELINES
        fs.cnt[20] = (int)(t_a1 - t_a0);
        fs.cnt[21] = (int)(clock64() - t_a1);
        fs.cnt[22] = (int)m.lsap_t_stage;
        fs.cnt[23] = (int)m.lsap_t_solve;
        fs.cnt[24] = (int)m.stats[0]; fs.cnt[25] = (int)m.stats[1]; fs.cnt[26] = (int)m.stats[2];
#endif
    }
}

// Stage B: LSAP over candidates x leftover dets (IoU), final lists.
__global__ void assign_stage_b_kernel(TrackTable tt, FrameScratch fs, SsbDims d, int L,
                                      size_t cost_smem_bytes) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_w[33];
    __shared__ int s_nu0;
    LsapSmem m;
    lsap_carve(smem, L, m);
    const int rows = fs.cnt[FC_N_CAND_B], cols = fs.cnt[FC_N_UNDET_A];
    const int nm0 = fs.cnt[FC_N_MATCH], nk = fs.cnt[FC_N_UNTRK_A_KEEP];
    lsap_block(fs.cost_b, rows, cols, m, cost_smem_bytes, fs.col4row, fs.row4col);
    for (int k = threadIdx.x; k < nk; k += blockDim.x) fs.untrk[k] = fs.untrk_a_keep[k];
    {
        int tot[1];
        block_partition<1>(cols, s_w, tot,
                           [&](int c) { return fs.row4col[c] < 0 ? 0 : -1; },
                           [&](int, int pos, int c) { fs.undet[pos] = fs.undet_a[c]; });
        if (threadIdx.x == 0) s_nu0 = tot[0];
    }
    __syncthreads();
    const int nu0 = s_nu0;
    auto row_class = [&](int r) -> int {          // 0 match, 1 rejected pair, 2 unassigned row
        const int c = fs.col4row[r];
        if (c < 0) return 2;
        return fs.cost_b[(size_t)r * cols + c] > d.max_iou ? 1 : 0;
    };
    int tot2[2];
    block_partition<2>(rows, s_w, tot2,
                       [&](int r) { return row_class(r) == 0 ? 0 : 1; },
                       [&](int cls, int pos, int r) {
                           const int p = fs.cand_b[r];
                           if (cls == 0) { fs.match_trk[nm0 + pos] = p; fs.match_det[nm0 + pos] = fs.undet_a[fs.col4row[r]]; }
                           else fs.untrk[nk + pos] = p;
                       });
    int totr[1];
    block_partition<1>(rows, s_w, totr,
                       [&](int r) { return row_class(r) == 1 ? 0 : -1; },
                       [&](int, int pos, int r) { fs.undet[nu0 + pos] = fs.undet_a[fs.col4row[r]]; });
    if (threadIdx.x == 0) {
        fs.cnt[FC_N_MATCH] = nm0 + tot2[0];
        fs.cnt[FC_N_UNTRK] = nk + tot2[1];
        fs.cnt[FC_N_UNDET] = nu0 + totr[0];
        fs.cnt[FC_ROWS_B] = rows; fs.cnt[FC_COLS_B] = cols;
    }
}

// ---------------------------------------------------------------------------
// Track.update for every match: KF update (thread 0) + EMA feature (block)
// ---------------------------------------------------------------------------
__device__ double block_sum_f64(double v, double *s_red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[wid] = v;
    __syncthreads();
    double t = 0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int k = 0; k < nw; k++) t += s_red[k];
    __syncthreads();
    return t;
}

__global__ void update_matched_kernel(TrackTable tt, FrameScratch fs, SsbDims d) {
    __shared__ double s_red[32];
    __shared__ double s_K[32];
    const int k = blockIdx.x;
    if (k >= fs.cnt[FC_N_MATCH]) return;
    const int pos = fs.match_trk[k], det = fs.match_det[k];
    const int s = tt.order[pos];
    const int D = d.D;
    const float *f = fs.feats + (size_t)det * D;
    float *tf = tt.feat + (size_t)s * D;
    // feature = det.feature / ||det.feature||
    const float nrm = fs.det_norm[det];
    double acc = 0;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const float fn = f[i] / nrm;
        const float sm = d.ema_alpha * tf[i] + d.one_minus_alpha * fn;
        tf[i] = sm;       // own elements only: safe in place
        acc += (double)sm * (double)sm;
    }
    const double tot = block_sum_f64(acc, s_red);
    const float n2 = (float)sqrt(tot);
    for (int i = threadIdx.x; i < D; i += blockDim.x) tf[i] = tf[i] / n2;
    if (threadIdx.x < 32)
        kf_update_warp(tt.mean + (size_t)s * 8, tt.cov + (size_t)s * 64, fs.det_xyah + det * 4,
                       (double)fs.det_conf[det], threadIdx.x, s_K);
    if (threadIdx.x == 0) {
        tt.conf[s] = fs.det_conf[det];
        tt.cls[s] = (int)fs.det_cls[det];
        tt.last_det[s] = det;
        const int h = tt.hits[s] + 1;
        tt.hits[s] = h;
        tt.tsu[s] = 0;
        if (tt.state[s] == SSB_TENTATIVE && h >= d.n_init) tt.state[s] = SSB_CONFIRMED;
    }
}

// ---------------------------------------------------------------------------
// bookkeeping: mark_missed, _initiate_track (ids in unmatched-det order),
// drop deleted tracks (order preserved), output rows (A.2), counters.
// Single block; the per-new-track feature copies use all warps.
// ---------------------------------------------------------------------------
// the reference's --count reduction for one track (yolo_multi_model.py:284-300): the most frequent class of
// its label lines, the SMALLEST class on ties (Counter(sorted(list)).most_common(1)); -1 when never reported
__device__ int hist_majority(int *h, bool clear) {
    int best = -1, bn = 0;
    for (int c = 0; c < SSB_NCLS; c++) {
        const int v = h[c];
        if (v > bn) { bn = v; best = c; }
        if (clear) h[c] = 0;
    }
    return best;
}

__global__ void class_counts_kernel(TrackTable tt, int *out) {
    __shared__ int s_cnt[SSB_NCLS];
    for (int c = threadIdx.x; c < SSB_NCLS; c += blockDim.x) s_cnt[c] = tt.dead_count[c];
    __syncthreads();
    const int T = tt.scalars[SC_N_TRACKS];
    for (int p = threadIdx.x; p < T; p += blockDim.x) {
        const int maj = hist_majority(tt.cls_hist + (size_t)tt.order[p] * SSB_NCLS, false);
        if (maj >= 0) atomicAdd(&s_cnt[maj], 1);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < SSB_NCLS; c += blockDim.x) out[c] = s_cnt[c];
}

__global__ void bookkeep_kernel(TrackTable tt, FrameScratch fs, SsbDims d, int H, int W,
                                double *out, int *counts, const int *tc_status) {
    __shared__ int s_w[33];
    __shared__ int s_nnew;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    const int T0 = tt.scalars[SC_N_TRACKS];
    const int n_free = tt.scalars[SC_N_FREE];
    const int next_id = tt.scalars[SC_NEXT_ID];
    // unmatched tracks did not see a detection this frame
    for (int p = tid; p < T0; p += blockDim.x) tt.last_det[tt.order[p]] = -1;
    __syncthreads();
    for (int k = tid; k < fs.cnt[FC_N_MATCH]; k += blockDim.x)
        tt.last_det[tt.order[fs.match_trk[k]]] = fs.match_det[k];
    // 1. mark_missed
    for (int k = tid; k < fs.cnt[FC_N_UNTRK]; k += blockDim.x) {
        const int s = tt.order[fs.untrk[k]];
        if (tt.state[s] == SSB_TENTATIVE) tt.state[s] = SSB_DELETED;
        else if (tt.tsu[s] > d.max_age) tt.state[s] = SSB_DELETED;
    }
    // 2. new tracks
    if (tid == 0) {
        int nn = fs.cnt[FC_N_UNDET];
        int err = 0;
        if (nn > n_free) { nn = n_free; err = 1; }
        if (nn > d.S - T0) { nn = d.S - T0; err = 1; }
        s_nnew = nn;
        if (err) tt.scalars[SC_ERROR] = 1;
    }
    __syncthreads();
    const int n_new = s_nnew;
    // (C4 starts ~240 tracks per frame: as one warp per track with lane 0 writing the scalars and the 64 covariance
    // entries and 16 dependent load -> divide -> store rounds for the feature, this step was 140 of the kernel's
    // 162 us; now three flat loops over all threads, the feature copy with eight loads in flight)
    for (int k = tid; k < n_new; k += blockDim.x) {            // 2a. scalars + mean
        const int det = fs.undet[k];
        const int s = tt.free_stack[n_free - 1 - k];
        tt.order[T0 + k] = s;
        const float *z = fs.det_xyah + det * 4;
        double *m = tt.mean + (size_t)s * 8;
        m[0] = (double)z[0]; m[1] = (double)z[1]; m[2] = (double)z[2]; m[3] = (double)z[3];
        m[4] = 0; m[5] = 0; m[6] = 0; m[7] = 0;
        tt.track_id[s] = next_id + k;
        tt.state[s] = SSB_TENTATIVE;
        tt.hits[s] = 1; tt.age[s] = 1; tt.tsu[s] = 0;
        tt.cls[s] = (int)fs.det_cls[det];
        tt.conf[s] = fs.det_conf[det];
        tt.last_det[s] = det;
        tt.gal_count[s] = 0; tt.gal_head[s] = 0;
    }
    for (int e = tid; e < n_new * 64; e += blockDim.x) {       // 2b. covariance: diag(std^2), zeros elsewhere
        const int k = e >> 6, idx = e & 63, i = idx / 9;
        const int det = fs.undet[k];
        const int s = tt.free_stack[n_free - 1 - k];
        double v = 0.0;
        if (idx == i * 9) {
            const float *z = fs.det_xyah + det * 4;
            const double z0 = z[0], z1 = z[1], z2 = z[2], z3 = z[3];
            const double std = i == 0 ? 2 * KF_WP * z0 : i == 1 ? 2 * KF_WP * z1 : i == 2 ? 1 * z2 : i == 3 ? 2 * KF_WP * z3
                             : i == 4 ? 10 * KF_WV * z0 : i == 5 ? 10 * KF_WV * z1 : i == 6 ? 0.1 * z2 : 10 * KF_WV * z3;
            v = std * std;
        }
        tt.cov[(size_t)s * 64 + idx] = v;
    }
    {                                                          // 2c. EMA feature = unit embedding
        const int D = d.D, total = n_new * D, step = blockDim.x;
        for (int e0 = tid; e0 < total; e0 += 8 * step) {
            float f[8], nrm[8];
            int dst[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int e = e0 + q * step;
                dst[q] = -1; f[q] = 0.f; nrm[q] = 1.f;
                if (e < total) {
                    const int k = e / D, i = e - k * D;
                    const int det = fs.undet[k];
                    f[q] = fs.feats[(size_t)det * D + i];
                    nrm[q] = fs.det_norm[det];
                    dst[q] = tt.free_stack[n_free - 1 - k] * D + i;
                }
            }
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (dst[q] >= 0) tt.feat[dst[q]] = f[q] / nrm[q];
        }
    }
    __syncthreads();
    // 3. drop deleted tracks, keep order; recycle their slots
    const int T1 = T0 + n_new;
    const int per = (T1 + blockDim.x - 1) / blockDim.x;
    const int b = tid * per, e = min(T1, b + per);
    int c = 0;
    for (int p = b; p < e; p++) c += tt.state[tt.order[p]] != SSB_DELETED;
    int keep_tot;
    int off = block_excl_scan(c, s_w, &keep_tot);
    int offd = b - off;
    const int free_base = n_free - n_new;
    for (int p = b; p < e; p++) {
        const int s = tt.order[p];
        if (tt.state[s] != SSB_DELETED) tt.order_tmp[off++] = s;
        else {
            tt.free_stack[free_base + offd] = s; offd++; tt.gal_count[s] = 0; tt.gal_head[s] = 0;
            // --count: a deleted id keeps counting.  Only a track that was ever confirmed (hits >= n_init; a tentative
            // track dies on its first miss) has label lines, i.e. a non-zero histogram: the 80-entry scan + clear of the
            // ~240 tentative tracks C4 drops per frame was 100 of this kernel's 160 us
            if (tt.hits[s] >= d.n_init) {
                const int maj = hist_majority(tt.cls_hist + (size_t)s * SSB_NCLS, true);
                if (maj >= 0) atomicAdd(&tt.dead_count[maj], 1);
            }
        }
    }
    __syncthreads();
    for (int p = tid; p < keep_tot; p += blockDim.x) tt.order[p] = tt.order_tmp[p];
    __syncthreads();
    // 4. output rows: confirmed and time_since_update <= 1, list order
    const int per2 = (keep_tot + blockDim.x - 1) / blockDim.x;
    const int b2 = tid * per2, e2 = min(keep_tot, b2 + per2);
    int c2 = 0, cc = 0;
    for (int p = b2; p < e2; p++) {
        const int s = tt.order[p];
        const bool conf = tt.state[s] == SSB_CONFIRMED;
        cc += conf;
        c2 += conf && tt.tsu[s] <= 1;
    }
    int out_tot;
    int off2 = block_excl_scan(c2, s_w, &out_tot);
    int conf_tot;
    (void)block_excl_scan(cc, s_w, &conf_tot);
    for (int p = b2; p < e2; p++) {
        const int s = tt.order[p];
        if (tt.state[s] != SSB_CONFIRMED || tt.tsu[s] > 1) continue;
        const double *m = tt.mean + (size_t)s * 8;
        const double w = m[2] * m[3];
        const double x = m[0] - w / 2, y = m[1] - m[3] / 2, h = m[3];
        double *o = out + (size_t)off2 * SSB_OUT_COLS;
        o[0] = (double)max((int)x, 0);
        o[1] = (double)max((int)y, 0);
        o[2] = (double)min((int)(x + w), W - 1);
        o[3] = (double)min((int)(y + h), H - 1);
        o[4] = (double)tt.track_id[s];
        o[5] = (double)tt.cls[s];
        o[6] = (double)tt.conf[s];
        o[7] = (double)tt.last_det[s];
        off2++;
        const int c = tt.cls[s];
        tt.cls_hist[(size_t)s * SSB_NCLS + (c < 0 ? 0 : (c >= SSB_NCLS ? SSB_NCLS - 1 : c))] += 1;
    }
    if (tid == 0) {
        tt.scalars[SC_N_TRACKS] = keep_tot;
        tt.scalars[SC_NEXT_ID] = next_id + n_new;
        tt.scalars[SC_N_FREE] = free_base + (T1 - keep_tot);
        tt.scalars[SC_FRAME] += 1;
        counts[SSB_CNT_OUT_ROWS] = out_tot;
        counts[SSB_CNT_TRACKS] = keep_tot;
        counts[SSB_CNT_CONFIRMED] = conf_tot;
        counts[SSB_CNT_NEXT_ID] = next_id + n_new;
        counts[SSB_CNT_MATCHES_A] = fs.cnt[FC_N_MATCH_A];
        counts[SSB_CNT_MATCHES_B] = fs.cnt[FC_N_MATCH] - fs.cnt[FC_N_MATCH_A];
        counts[SSB_CNT_NEW] = n_new;
        // bit 0: table overflow; bit 1: a tensor-core barrier wait of the ReID kernels timed out
        // (their device status word) -- the embeddings of this frame cannot be trusted
        counts[SSB_CNT_ERROR] = (tt.scalars[SC_ERROR] ? 1 : 0) | ((tc_status && *tc_status) ? 2 : 0);
    }
}

// StrongSORT.increment_ages() of upstream (called by its stream loop on frames WITHOUT detections):
// for every track  age += 1, time_since_update += 1, mark_missed()  -- no Kalman predict.  Tracks
// deleted here stay in the list until the next update() drops them (upstream filters the list only at
// the end of Tracker.update); until then they sit among the unconfirmed tracks, exactly as upstream.
__global__ void increment_ages_kernel(TrackTable tt, SsbDims d) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= tt.scalars[SC_N_TRACKS]) return;
    const int s = tt.order[p];
    tt.age[s] += 1;
    const int tsu = tt.tsu[s] + 1;
    tt.tsu[s] = tsu;
    if (tt.state[s] == SSB_TENTATIVE) tt.state[s] = SSB_DELETED;
    else if (tsu > d.max_age) tt.state[s] = SSB_DELETED;
}

// gallery partial_fit (A.5/A.6): every confirmed track appends its current
// smoothed feature once per frame; ring of `budget` samples per slot.
__global__ void gallery_append_kernel(TrackTable tt, SsbDims d) {
    const int p = blockIdx.x;
    if (p >= tt.scalars[SC_N_TRACKS]) return;
    const int s = tt.order[p];
    if (tt.state[s] != SSB_CONFIRMED) return;
    const int head = tt.gal_head[s];
    const float *f = tt.feat + (size_t)s * d.D;
    float *g = tt.gallery + ((size_t)s * d.B + head) * d.D;
    for (int i = threadIdx.x; i < d.D; i += blockDim.x) g[i] = f[i];
    if (d.B <= SSB_GAL_ROWS && d.D == 512) {
        // the same sample as a tensor-core operand row: re-normalised like _cosine_distance does
        // (a / ||a||, float32), * 2^6, hi/lo fp16; thread i owns elements 4i .. 4i+3 (128 threads)
        __shared__ float s_ss[4];
        const float4 v = *reinterpret_cast<const float4 *>(f + threadIdx.x * 4);
        float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if ((threadIdx.x & 31) == 0) s_ss[threadIdx.x >> 5] = ss;
        __syncthreads();
        const float nrm = sqrtf((s_ss[0] + s_ss[1]) + (s_ss[2] + s_ss[3]));
        const float u[4] = {(v.x / nrm) * 64.0f, (v.y / nrm) * 64.0f, (v.z / nrm) * 64.0f, (v.w / nrm) * 64.0f};
        __align__(8) __half h[4];
        __align__(8) __half l[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { h[j] = __float2half_rn(u[j]); l[j] = __float2half_rn(u[j] - __half2float(h[j])); }
        const int chunk = threadIdx.x >> 1, half8 = (threadIdx.x & 1) * 8;
        unsigned char *dst = tt.gal_planes + (size_t)s * (2 * 64 * SSB_GAL_ROWS * 16) + ((size_t)chunk * SSB_GAL_ROWS + head) * 16 + half8;
        *reinterpret_cast<uint2 *>(dst) = *reinterpret_cast<uint2 *>(h);
        *reinterpret_cast<uint2 *>(dst + 64 * SSB_GAL_ROWS * 16) = *reinterpret_cast<uint2 *>(l);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        tt.gal_head[s] = (head + 1) % d.B;
        tt.gal_count[s] = min(tt.gal_count[s] + 1, d.B);
    }
}

__global__ void reset_table_kernel(TrackTable tt, SsbDims d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d.S) {
        tt.free_stack[i] = d.S - 1 - i;   // pop order: slot 0 first
        tt.state[i] = SSB_DELETED;
        tt.gal_count[i] = 0; tt.gal_head[i] = 0;
        tt.order[i] = 0;
        for (int c = 0; c < SSB_NCLS; c++) tt.cls_hist[(size_t)i * SSB_NCLS + c] = 0;
    }
    if (i < SSB_NCLS) tt.dead_count[i] = 0;
    if (i == 0) {
        for (int k = 0; k < SC_COUNT; k++) tt.scalars[k] = 0;
        tt.scalars[SC_NEXT_ID] = 1;
        tt.scalars[SC_N_FREE] = d.S;
    }
}

__global__ void export_tracks_kernel(TrackTable tt, SsbDims d, int *ids, int *state, int *hits,
                                     int *age, int *tsu, int *gal, double *mean, double *cov,
                                     float *feat) {
    const int p = blockIdx.x;
    if (p >= tt.scalars[SC_N_TRACKS]) return;
    const int s = tt.order[p];
    if (threadIdx.x == 0) {
        if (ids) ids[p] = tt.track_id[s];
        if (state) state[p] = tt.state[s];
        if (hits) hits[p] = tt.hits[s];
        if (age) age[p] = tt.age[s];
        if (tsu) tsu[p] = tt.tsu[s];
        if (gal) gal[p] = tt.gal_count[s];
    }
    if (mean) for (int i = threadIdx.x; i < 8; i += blockDim.x) mean[(size_t)p * 8 + i] = tt.mean[(size_t)s * 8 + i];
    if (cov) for (int i = threadIdx.x; i < 64; i += blockDim.x) cov[(size_t)p * 64 + i] = tt.cov[(size_t)s * 64 + i];
    if (feat) for (int i = threadIdx.x; i < d.D; i += blockDim.x) feat[(size_t)p * d.D + i] = tt.feat[(size_t)s * d.D + i];
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
static int g_lsap_smem_limit_dev[64];   // bytes of dynamic smem opted in, per device

// L: upper bound of max(rows, cols) of the problem(s) the launch will solve; want_cost_bytes: bytes of the largest
// dense cost matrix it may have to stage.  The launch asks for what THIS frame needs, not for the whole SM: a
// 225 KB request (round 1) can only be placed on an EMPTY SM, so in the two-stage pipeline every assignment kernel
// waited for a gap between the ReID kernels of the next frame; ~90 KB at C2 co-resides with a ReID CTA.
static int lsap_prepare(int L, size_t want_cost_bytes, size_t *dyn_bytes, size_t *cost_bytes) {
    int dev = 0;
    SSB_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { ssb_set_error("device index %d out of range", dev); return -3; }
    static const int key = ssb_new_key();
    if (ssb_first_on_device(key) || !g_lsap_smem_limit_dev[dev]) {
        int maxopt = 0;
        SSB_CHECK_CUDA(cudaDeviceGetAttribute(&maxopt, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        int want = maxopt - 2048;
        SSB_CHECK_CUDA(cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, want));
        SSB_CHECK_CUDA(cudaFuncSetAttribute(assign_stage_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, want));
        SSB_CHECK_CUDA(cudaFuncSetAttribute(assign_stage_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, want));
        g_lsap_smem_limit_dev[dev] = want;
    }
    const int g_lsap_smem_limit = g_lsap_smem_limit_dev[dev];
    size_t fixed = (size_t)L * (3 * 8 + 4 * 4 + 2) + 64;
    if (fixed + 1024 > (size_t)g_lsap_smem_limit) {
        ssb_set_error("LSAP dimension %d exceeds shared memory", L);
        return -3;
    }
    const size_t avail = (size_t)g_lsap_smem_limit - fixed - 16;
    // The WHOLE request (per-row/column state + matrix) is capped at 112 KB so that the kernel fits the slot one
    // stage-2 ReID CTA (108.6 KB with its reserve) leaves when it retires.  It matters in the two-stage pipeline: the
    // block scheduler backfills freed slots with whatever fits, so a high-priority CTA that needs more than any
    // single freed slot is starved until the ReID kernels of the next frame run out of CTAs -- at C4 (13 waves per
    // kernel, three parts on three streams) that was the end of the whole forward: association and embedding ran
    // back to back (6.98 ms per frame = 2.2 + 3.6 + gaps; SSB_LSAP_EXCL-style whole-SM requests changed nothing).
    // Inside the cap: config C2's dense 100 x 100 matrix (staging it densely is 18 % faster than its sparse form:
    // 113 vs 134 us); a bigger matrix is kept as its non-background entries (12 bytes each; C4 has ~2 500) -- beyond
    // that the dense global path runs.
    static const size_t total_cap = [] { const char *v = getenv("SSB_LSAP_SMEM_CAP"); return v && *v ? (size_t)atoll(v) : (size_t)114688; }();
    size_t cap = total_cap > fixed + 16 ? total_cap - fixed - 16 : 0;
    // ... but never at the price of matrix capacity: once the per-row/column state leaves less than 64 KB under the
    // cap (C4 with two frames of detections in the track bound: 80 KB of state), the request takes what the SM has --
    // a stream whose non-background entries outgrow the store falls to the dense global path, 10 ms instead of 2.5
    // per frame (profiles/r02_pipeline_trace.md, frames 27+ of the C4 stream under the 112 KB cap).
    if (cap < 65536) cap = avail;
    size_t cost = want_cost_bytes < cap ? want_cost_bytes : cap;
    if (cost > avail) cost = avail;
    if (cost < 16384) cost = 16384 < avail ? 16384 : avail;
    *cost_bytes = cost;
    *dyn_bytes = (fixed + cost + 16 + 1023) & ~(size_t)1023;
    if (*dyn_bytes > (size_t)g_lsap_smem_limit) *dyn_bytes = g_lsap_smem_limit;
    return 0;
}

int ssb_launch_prep(const SsbDims &d, const float *dets, int n, int h, int w, FrameScratch fs,
                    cudaStream_t st) {
    if (n > 0) {
        prep_dets_kernel<<<(n + 127) / 128, 128, 0, st>>>(dets, n, h, w, fs);
        SSB_CHECK_LAUNCH();
    }
    return 0;
}

int ssb_launch_track_frame(ssb_tracker *t, int slot, int n, int h, int w, const float *feats, double *out,
                           int *counts, int track_hint, cudaStream_t st) {
    const SsbDims &d = t->dims;
    TrackTable tt = t->tt;
    FrameScratch fs = ssb_slot_view(t, slot);
    if (feats) fs.feats = const_cast<float *>(feats);
    const int Tmax = (track_hint >= 0 && track_hint <= d.S) ? track_hint : d.S;
    // rows <= live tracks entering the frame (Tmax), columns <= n, for both assignment stages
    const int L = (Tmax > n ? Tmax : n) > 1 ? (Tmax > n ? Tmax : n) : 1;
    size_t dyn = 0, cost_b = 0;
    int rc = lsap_prepare(L, (size_t)Tmax * (size_t)n * 8, &dyn, &cost_b);
    if (rc) return rc;
#define SSB_PROF(i) do { if (t->prof_on) SSB_CHECK_CUDA(cudaEventRecord(t->prof_ev[i], st)); } while (0)
    SSB_PROF(0);

    // tensor-core appearance cost whenever the operand planes can hold the problem (else the fp32 SIMT kernel)
    const bool app_tc = d.B <= SSB_GAL_ROWS && n <= SSB_DET_PLANES_MAX && d.D == 512 && !t->app_simt;
    if (n > 0) {
        det_norm_kernel<<<(n * 32 + 127) / 128, 128, 0, st>>>(fs.feats, n, d.D, fs.det_norm,
                                                              app_tc ? fs.det_planes : nullptr, ssb_det_npad(n));
        SSB_CHECK_LAUNCH();
    }
    if (Tmax > 0) {
        kf_predict_tracks_kernel<<<(Tmax * 32 + 127) / 128, 128, 0, st>>>(tt);
        SSB_CHECK_LAUNCH();
    }
    build_lists_kernel<<<1, 256, 0, st>>>(tt, fs);
    SSB_CHECK_LAUNCH();
    SSB_PROF(1);
    if (Tmax > 0 && n > 0) {
        if (app_tc)
            rc = ssb_launch_appearance_tc(tt.gal_planes, tt.gal_count, fs.conf_list, tt.order, fs.cnt + FC_N_CONF,
                                          Tmax, d.B, fs.det_planes, n, fs.app_cost, n, t->tc_status, st);
        else
            rc = ssb_launch_appearance(tt.gallery, tt.gal_count, tt.gal_head, fs.conf_list, tt.order,
                                       fs.cnt + FC_N_CONF, Tmax, d.B, fs.feats, n, d.D, fs.app_cost, n, st);
        if (rc) return rc;
        SSB_PROF(2);
        gate_cost_kernel<<<Tmax, 128, 0, st>>>(tt, fs, d, n);
        SSB_CHECK_LAUNCH();
    } else SSB_PROF(2);
    SSB_PROF(3);
    assign_stage_a_kernel<<<1, 256, dyn, st>>>(tt, fs, d, n, L, cost_b);
    SSB_CHECK_LAUNCH();
    SSB_PROF(4);
    if (Tmax > 0 && n > 0) {
        iou_cost_kernel<<<Tmax, 128, 0, st>>>(tt, fs, d);
        SSB_CHECK_LAUNCH();
    }
    SSB_PROF(5);
    assign_stage_b_kernel<<<1, 256, dyn, st>>>(tt, fs, d, L, cost_b);
    SSB_CHECK_LAUNCH();
    SSB_PROF(6);
    const int max_match = Tmax < n ? Tmax : n;
    if (max_match > 0) {
        update_matched_kernel<<<max_match, 128, 0, st>>>(tt, fs, d);
        SSB_CHECK_LAUNCH();
    }
    SSB_PROF(7);
    bookkeep_kernel<<<1, 256, 0, st>>>(tt, fs, d, h, w, out, counts, t->tc_status);
    SSB_CHECK_LAUNCH();
    SSB_PROF(8);
    int Tafter = Tmax + n;
    if (Tafter > d.S) Tafter = d.S;
    if (Tafter > 0) {
        gallery_append_kernel<<<Tafter, 128, 0, st>>>(tt, d);
        SSB_CHECK_LAUNCH();
    }
    SSB_PROF(9);
    if (t->prof_on) t->prof_have = 1;
    return 0;
}

// ---- stage entry points ----------------------------------------------------
extern "C" int ssb_crop_boxes(const float *dets_dev, int n, int h, int w, int32_t *boxes_out_dev,
                              ssb_stream_t stream) {
    if (n <= 0) return 0;
    crop_boxes_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dets_dev, n, h, w, boxes_out_dev);
    SSB_CHECK_LAUNCH();
    return 0;
}

// Track.camera_update for every live track: warp the tl / br corners with
// the 2x3 matrix and rewrite mean[:4].  Same float64 expression order as oracle Track.camera_update.
struct Warp6 { double a, b, tx, c, d, ty; };
__global__ void camera_update_kernel(TrackTable tt, Warp6 m) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= tt.scalars[SC_N_TRACKS]) return;
    double *mean = tt.mean + (size_t)tt.order[i] * 8;
    const double h0 = mean[3], w0 = mean[2] * h0;
    const double x1 = mean[0] - w0 / 2, y1 = mean[1] - h0 / 2;
    const double x2 = x1 + w0, y2 = y1 + h0;
    const double x1_ = m.a * x1 + m.b * y1 + m.tx, y1_ = m.c * x1 + m.d * y1 + m.ty;
    const double x2_ = m.a * x2 + m.b * y2 + m.tx, y2_ = m.c * x2 + m.d * y2 + m.ty;
    const double w = x2_ - x1_, h = y2_ - y1_;
    mean[0] = x1_ + w / 2;
    mean[1] = y1_ + h / 2;
    mean[2] = w / h;
    mean[3] = h;
}

extern "C" int ssb_camera_update(ssb_tracker *t, const double *warp2x3_host, ssb_stream_t stream) {
    if (!t || !warp2x3_host) { ssb_set_error("null argument"); return -1; }
    Warp6 m = {warp2x3_host[0], warp2x3_host[1], warp2x3_host[2], warp2x3_host[3], warp2x3_host[4], warp2x3_host[5]};
    camera_update_kernel<<<(t->dims.S + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t->tt, m);
    SSB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ssb_kf_predict(double *mean_dev, double *cov_dev, int n, ssb_stream_t stream) {
    if (n <= 0) return 0;
    kf_predict_arrays_kernel<<<(n * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mean_dev, cov_dev, n);
    SSB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ssb_kf_update(double *mean_dev, double *cov_dev, const float *xyah_dev,
                             const float *conf_dev, int n, ssb_stream_t stream) {
    if (n <= 0) return 0;
    kf_update_arrays_kernel<<<(n * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mean_dev, cov_dev, xyah_dev, conf_dev, n);
    SSB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ssb_kf_gating(const double *mean_dev, const double *cov_dev, int n_tracks,
                             const float *xyah_dev, int n_meas, double *maha_out_dev,
                             ssb_stream_t stream) {
    if (n_tracks <= 0 || n_meas <= 0) return 0;
    kf_gating_arrays_kernel<<<n_tracks, 128, 0, (cudaStream_t)stream>>>(mean_dev, cov_dev, n_tracks, xyah_dev, n_meas, maha_out_dev);
    SSB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ssb_iou_cost(const double *track_tlwh_dev, int n_tracks, const float *det_tlwh_dev,
                            int n_dets, double *cost_out_dev, ssb_stream_t stream) {
    if (n_tracks <= 0 || n_dets <= 0) return 0;
    iou_cost_arrays_kernel<<<n_tracks, 128, 0, (cudaStream_t)stream>>>(track_tlwh_dev, n_tracks, det_tlwh_dev, n_dets, cost_out_dev);
    SSB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ssb_lsap(const double *cost_dev, int nr, int nc, int32_t *col4row_out_dev,
                        int32_t *row4col_out_dev, ssb_stream_t stream) {
    if (nr < 0 || nc < 0) { ssb_set_error("negative LSAP dims"); return -1; }
    const int L = (nr > nc ? nr : nc) > 1 ? (nr > nc ? nr : nc) : 1;
    size_t dyn = 0, cost_b = 0;
    int rc = lsap_prepare(L, (size_t)nr * (size_t)nc * 8, &dyn, &cost_b);
    if (rc) return rc;
    lsap_kernel<<<1, 256, dyn, (cudaStream_t)stream>>>(cost_dev, nr, nc, L, cost_b, col4row_out_dev, row4col_out_dev);
    SSB_CHECK_LAUNCH();
    return 0;
}

int ssb_launch_reset(ssb_tracker *t, cudaStream_t st) {
    reset_table_kernel<<<(t->dims.S + 127) / 128, 128, 0, st>>>(t->tt, t->dims);
    SSB_CHECK_LAUNCH();
    SSB_CHECK_CUDA(cudaMemsetAsync(t->tc_status, 0, 64 * sizeof(int), st));     // workspace memory is uninitialised
    return 0;
}

// per-stage timing of the association (bench.py's stage split): events between the kernels of ssb_associate.
// ms_out[9]: prep (norms, KF predict, lists) | appearance | gate | LSAP A + lists | IoU | LSAP B + lists |
//            KF/EMA update | bookkeeping | gallery append.  Synchronises on the last event.
extern "C" int ssb_profile_enable(ssb_tracker *t, int on) {
    if (!t) { ssb_set_error("null handle"); return -1; }
    if (on && !t->prof_ev[0])
        for (int i = 0; i < 12; i++) SSB_CHECK_CUDA(cudaEventCreate(&t->prof_ev[i]));
    t->prof_on = on ? 1 : 0;
    t->prof_have = 0;
    return 0;
}
extern "C" int ssb_profile_read(ssb_tracker *t, float *ms_out9) {
    if (!t || !ms_out9) { ssb_set_error("null argument"); return -1; }
    if (!t->prof_have) { ssb_set_error("no profiled frame yet"); return -1; }
    SSB_CHECK_CUDA(cudaEventSynchronize(t->prof_ev[9]));
    for (int i = 0; i < 9; i++) SSB_CHECK_CUDA(cudaEventElapsedTime(&ms_out9[i], t->prof_ev[i], t->prof_ev[i + 1]));
    return 0;
}

// --count of the reference (yolo_multi_model.py:284-300) as a reduction over the track table: out_dev[c] =
// number of track ids whose label lines so far carry class c most often (ties: smallest class)
extern "C" int ssb_class_counts(ssb_tracker *t, int32_t *out_dev, ssb_stream_t stream) {
    if (!t || !out_dev) { ssb_set_error("null argument"); return -1; }
    class_counts_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(t->tt, out_dev);
    SSB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ssb_increment_ages(ssb_tracker *t, ssb_stream_t stream) {
    if (!t) { ssb_set_error("null handle"); return -1; }
    increment_ages_kernel<<<(t->dims.S + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t->tt, t->dims);
    SSB_CHECK_LAUNCH();
    return 0;
}

int ssb_launch_export(ssb_tracker *t, int *ids, int *state, int *hits, int *age, int *tsu, int *gal,
                      double *mean, double *cov, float *feat, cudaStream_t st) {
    export_tracks_kernel<<<t->dims.S, 64, 0, st>>>(t->tt, t->dims, ids, state, hits, age, tsu, gal, mean, cov, feat);
    SSB_CHECK_LAUNCH();
    return 0;
}
