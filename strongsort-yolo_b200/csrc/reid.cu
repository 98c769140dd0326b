// OSNet-x0.25 ReID embedding (SURVEY.md A.3 + Appendix B), fp32 SIMT version.
//
//   crop(img[y1:y2,x1:x2], BGR as is) -> bilinear 256x128 (half-pixel centres)
//   -> /255, ImageNet mean/std -> conv7x7/2 + ReLU -> maxpool3x3/2
//   -> 6 OSBlocks (+2 transition conv/avgpool) -> conv5 -> GAP -> fc+ReLU [512]
//
// BatchNorm is folded into the preceding conv by the host (weights.py); all
// activations are NHWC float32 so that 1x1 convs are [pixels x Cin]x[Cin x Cout]
// products with channels innermost (coalesced).  This file is the bit-careful
// baseline the tensor-core path (reid_tc.cu) is validated against.
#include "ssb_common.cuh"

#ifdef SSB_BASELINES        // A/B baselines: libssb_dbg.so only (build.py)

// ---------------------------------------------------------------------------
// architecture walk shared by host packer (weights.py mirrors it) and runtime
// ---------------------------------------------------------------------------
static const int kBlocks[6][2] = {{16, 64}, {64, 64}, {64, 96}, {96, 96}, {96, 128}, {128, 128}};
#define REID_H 256
#define REID_W 128

static int reid_tensor_sizes_host(int64_t *sizes) {
    int k = 0;
    auto add = [&](int64_t n) { if (sizes) sizes[k] = n; k++; };
    add(7 * 7 * 3 * 16); add(16);
    for (int b = 0; b < 6; b++) {
        const int cin = kBlocks[b][0], cout = kBlocks[b][1], mid = cout / 4;
        const int r = mid / 16 > 0 ? mid / 16 : 1;
        add((int64_t)cin * mid); add(mid);
        for (int l = 0; l < 10; l++) { add((int64_t)mid * mid); add(9 * mid); add(mid); }
        add((int64_t)mid * r); add(r); add((int64_t)r * mid); add(mid);
        add((int64_t)mid * cout); add(cout);
        if (cin != cout) { add((int64_t)cin * cout); add(cout); }
        if (b == 1 || b == 3) { add((int64_t)cout * cout); add(cout); }
    }
    add(128 * 128); add(128);
    add(128 * 512); add(512);
    return k;
}

extern "C" int ssb_reid_num_tensors(void) { return reid_tensor_sizes_host(nullptr); }
extern "C" int ssb_reid_tensor_sizes(int64_t *sizes) { reid_tensor_sizes_host(sizes); return 0; }

extern "C" int ssb_reid_set_weights(ssb_tracker *t, const float *blob_dev, const int64_t *sizes, int n) {
    if (!t || !blob_dev || !sizes) { ssb_set_error("null argument"); return -1; }
    const int want = ssb_reid_num_tensors();
    if (n != want) { ssb_set_error("expected %d weight tensors, got %d", want, n); return -1; }
    int64_t *exp = (int64_t *)malloc(sizeof(int64_t) * want);
    reid_tensor_sizes_host(exp);
    free(t->w_off);
    t->w_off = (int64_t *)malloc(sizeof(int64_t) * (want + 1));
    int64_t off = 0;
    for (int i = 0; i < want; i++) {
        if (sizes[i] != exp[i]) {
            ssb_set_error("weight tensor %d: expected %lld elements, got %lld", i, (long long)exp[i], (long long)sizes[i]);
            free(exp);
            return -1;
        }
        t->w_off[i] = off;
        off += (exp[i] + 3) & ~(int64_t)3;     // every tensor 16-byte aligned in the blob
    }
    t->w_off[want] = off;
    free(exp);
    t->w_blob = blob_dev;
    t->n_w = want;
    return 0;
}

#define REID_BIG 131072
#define REID_MID 32768

// ---------------------------------------------------------------------------
// stem: crop + resize + normalise + conv7x7 s2 + ReLU + maxpool3x3 s2  (fused)
// grid (32 tiles, N): each CTA makes an 8x8 patch of the pooled 64x32x16 map
// ---------------------------------------------------------------------------
#define ST_PATCH 39
#define ST_CONV 17
#define ST_IN_FLOATS (ST_PATCH * ST_PATCH * 3 + 1)   // +1 keeps the next arrays 16-byte aligned
__global__ void __launch_bounds__(256)
reid_stem_kernel(const uint8_t *__restrict__ img, int H, int W, int pitch,
                 const int *__restrict__ boxes, const float *__restrict__ wts,
                 const float *__restrict__ bias, float *__restrict__ out) {
    extern __shared__ float sm[];
    float *s_in = sm;                                   // [39][39][3]
    float *s_conv = s_in + ST_IN_FLOATS;                // [17][17][16]
    float *s_w = s_conv + ST_CONV * ST_CONV * 16;       // [147][16]
    const int n = blockIdx.y;
    const int tyi = blockIdx.x >> 2, txi = blockIdx.x & 3;
    const int py0 = tyi * 8, px0 = txi * 8;
    const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;     // first conv row/col of the tile
    const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;     // first resized-image row/col
    const int bx1 = boxes[n * 4 + 0], by1 = boxes[n * 4 + 1];
    const int cw = boxes[n * 4 + 2] - bx1, ch = boxes[n * 4 + 3] - by1;
    const int tid = threadIdx.x;
    for (int i = tid; i < 147 * 16; i += 256) s_w[i] = wts[i];
    const float mean[3] = {0.485f, 0.456f, 0.406f};
    const float stdv[3] = {0.229f, 0.224f, 0.225f};
    const float sc_y = (float)ch / (float)REID_H, sc_x = (float)cw / (float)REID_W;
    for (int p = tid; p < ST_PATCH * ST_PATCH; p += 256) {
        const int iy = p / ST_PATCH, ix = p - iy * ST_PATCH;
        const int gy = iy0 + iy, gx = ix0 + ix;
        float v[3] = {0.f, 0.f, 0.f};
        if (gy >= 0 && gy < REID_H && gx >= 0 && gx < REID_W && cw > 0 && ch > 0) {
            float sy = sc_y * ((float)gy + 0.5f) - 0.5f;
            float sx = sc_x * ((float)gx + 0.5f) - 0.5f;
            if (sy < 0.f) sy = 0.f;
            if (sx < 0.f) sx = 0.f;
            const int y0 = (int)sy, x0 = (int)sx;
            const int y1 = y0 + (y0 < ch - 1 ? 1 : 0), x1 = x0 + (x0 < cw - 1 ? 1 : 0);
            const float ly = sy - (float)y0, lx = sx - (float)x0;
            const float hy = 1.f - ly, hx = 1.f - lx;
            const uint8_t *r0 = img + (size_t)(by1 + y0) * pitch + (size_t)bx1 * 3;
            const uint8_t *r1 = img + (size_t)(by1 + y1) * pitch + (size_t)bx1 * 3;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float p00 = r0[x0 * 3 + c], p01 = r0[x1 * 3 + c];
                const float p10 = r1[x0 * 3 + c], p11 = r1[x1 * 3 + c];
                const float val = hy * (hx * p00 + lx * p01) + ly * (hx * p10 + lx * p11);
                v[c] = (val / 255.0f - mean[c]) / stdv[c];
            }
        }
        s_in[p * 3 + 0] = v[0]; s_in[p * 3 + 1] = v[1]; s_in[p * 3 + 2] = v[2];
    }
    __syncthreads();
    // conv: items = 289 positions x 2 halves of 8 output channels
    for (int it = tid; it < ST_CONV * ST_CONV * 2; it += 256) {
        const int pos = it >> 1, half = (it & 1) * 8;
        const int cy = pos / ST_CONV, cx = pos - cy * ST_CONV;
        const int gcy = cy0 + cy, gcx = cx0 + cx;
        float acc[8];
        if (gcy < 0 || gcy >= 128 || gcx < 0 || gcx >= 64) {
#pragma unroll
            for (int k = 0; k < 8; k++) s_conv[pos * 16 + half + k] = -INFINITY;
            continue;
        }
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = bias[half + k];
        for (int ky = 0; ky < 7; ky++) {
            const float *ip = s_in + ((2 * cy + ky) * ST_PATCH + 2 * cx) * 3;
            const float *wp = s_w + (ky * 21) * 16 + half;
#pragma unroll
            for (int t = 0; t < 21; t++) {     // kx*3 + ci
                const float a = ip[t];
                const float4 w0 = *reinterpret_cast<const float4 *>(wp + t * 16);
                const float4 w1 = *reinterpret_cast<const float4 *>(wp + t * 16 + 4);
                acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]);
                acc[2] = fmaf(a, w0.z, acc[2]); acc[3] = fmaf(a, w0.w, acc[3]);
                acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]);
                acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) s_conv[pos * 16 + half + k] = fmaxf(acc[k], 0.f);
    }
    __syncthreads();
    // maxpool 3x3 s2 p1 -> 8x8x16
    for (int o = tid; o < 8 * 8 * 16; o += 256) {
        const int c = o & 15, q = o >> 4, qx = q & 7, qy = q >> 3;
        float m = -INFINITY;
#pragma unroll
        for (int dy = 0; dy < 3; dy++)
#pragma unroll
            for (int dx = 0; dx < 3; dx++)
                m = fmaxf(m, s_conv[((2 * qy + dy) * ST_CONV + 2 * qx + dx) * 16 + c]);
        out[(((size_t)n * 64 + py0 + qy) * 32 + px0 + qx) * 16 + c] = m;
    }
}

// ---------------------------------------------------------------------------
// pointwise conv: out[p][co] = act( sum_ci in[p][ci] W[ci][co] + b[co] (+ res) )
// 64 pixels per CTA, 256 threads = 64 px x 4 channel groups of CG outputs
// ---------------------------------------------------------------------------
template <int CG>
__global__ void __launch_bounds__(256)
pw_conv_kernel(const float *__restrict__ in, int P, int cin, const float *__restrict__ wts,
               const float *__restrict__ bias, const float *__restrict__ res,
               float *__restrict__ out, int relu) {
    constexpr int COUT = CG * 4;
    extern __shared__ float sm[];
    float *s_w = sm;                       // [cin][COUT]
    float *s_x = sm + cin * COUT;          // [64][cin+1]
    const int tid = threadIdx.x;
    const int p0 = blockIdx.x * 64;
    for (int i = tid; i < cin * COUT; i += 256) s_w[i] = wts[i];
    const int ldx = cin + 1;
    for (int i = tid; i < 64 * cin; i += 256) {
        const int px = i / cin, ci = i - px * cin;
        s_x[px * ldx + ci] = (p0 + px < P) ? in[(size_t)(p0 + px) * cin + ci] : 0.f;
    }
    __syncthreads();
    const int px = tid & 63, g = tid >> 6;
    float acc[CG];
#pragma unroll
    for (int k = 0; k < CG; k++) acc[k] = bias ? bias[g * CG + k] : 0.f;
    const float *xr = s_x + px * ldx;
    for (int ci = 0; ci < cin; ci++) {
        const float a = xr[ci];
        const float *wr = s_w + ci * COUT + g * CG;
#pragma unroll
        for (int k = 0; k < CG; k++) acc[k] = fmaf(a, wr[k], acc[k]);
    }
    if (p0 + px < P) {
        float *o = out + (size_t)(p0 + px) * COUT + g * CG;
        const float *rr = res ? res + (size_t)(p0 + px) * COUT + g * CG : nullptr;
#pragma unroll
        for (int k = 0; k < CG; k++) {
            float v = acc[k];
            if (rr) v += rr[k];
            if (relu) v = fmaxf(v, 0.f);
            o[k] = v;
        }
    }
}

// depthwise 3x3, pad 1, + bias + ReLU.  thread = (pixel, 4 channels)
__global__ void dw3x3_kernel(const float *__restrict__ in, int N, int H, int W, int C,
                             const float *__restrict__ wts, const float *__restrict__ bias,
                             float *__restrict__ out) {
    const int cg = C >> 2;
    const size_t total = (size_t)N * H * W * cg;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = idx % cg;
    size_t p = idx / cg;
    const int x = p % W; p /= W;
    const int y = p % H;
    const int n = p / H;
    float4 acc = *reinterpret_cast<const float4 *>(bias + g * 4);
#pragma unroll
    for (int dy = -1; dy <= 1; dy++) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
            const int xx = x + dx;
            if (xx < 0 || xx >= W) continue;
            const float4 v = *reinterpret_cast<const float4 *>(in + (((size_t)n * H + yy) * W + xx) * C + g * 4);
            const float4 w = *reinterpret_cast<const float4 *>(wts + ((dy + 1) * 3 + dx + 1) * C + g * 4);
            acc.x = fmaf(v.x, w.x, acc.x); acc.y = fmaf(v.y, w.y, acc.y);
            acc.z = fmaf(v.z, w.z, acc.z); acc.w = fmaf(v.w, w.w, acc.w);
        }
    }
    acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f);
    acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
    *reinterpret_cast<float4 *>(out + (((size_t)n * H + y) * W + x) * C + g * 4) = acc;
}

// global average pool per crop (+ optional ChannelGate MLP -> sigmoid gate)
// one CTA per crop.  gate_out[n][c] = sigmoid(fc2(relu(fc1(mean)))) if w1 != null
// else mean.
__global__ void __launch_bounds__(256)
gap_gate_kernel(const float *__restrict__ in, int px, int C, const float *__restrict__ w1,
                const float *__restrict__ b1, const float *__restrict__ w2,
                const float *__restrict__ b2, int r, float *__restrict__ gate_out) {
    __shared__ float s_part[256];
    __shared__ float s_mean[128];
    __shared__ float s_hid[8];
    const int n = blockIdx.x, tid = threadIdx.x;
    const int lanes = 256 / C;              // pixel phases per channel
    const int c = tid % C, ph = tid / C;
    float acc = 0.f;
    if (ph < lanes) {
        const float *base = in + (size_t)n * px * C + c;
        for (int p = ph; p < px; p += lanes) acc += base[(size_t)p * C];
    }
    s_part[tid] = acc;
    __syncthreads();
    if (tid < C) {
        float s = 0.f;
        for (int k = 0; k < lanes; k++) s += s_part[k * C + tid];
        s_mean[tid] = s / (float)px;
    }
    __syncthreads();
    if (!w1) {
        if (tid < C) gate_out[(size_t)n * C + tid] = s_mean[tid];
        return;
    }
    if (tid < r) {
        float h = b1[tid];
        for (int k = 0; k < C; k++) h = fmaf(s_mean[k], w1[k * r + tid], h);
        s_hid[tid] = fmaxf(h, 0.f);
    }
    __syncthreads();
    if (tid < C) {
        float g = b2[tid];
        for (int k = 0; k < r; k++) g = fmaf(s_hid[k], w2[k * C + tid], g);
        gate_out[(size_t)n * C + tid] = 1.f / (1.f + expf(-g));
    }
}

// x2 (+)= s * gate   thread = (pixel, 4 channels)
__global__ void gate_apply_kernel(const float *__restrict__ s, const float *__restrict__ gate,
                                  int N, int px, int C, float *__restrict__ x2, int first) {
    const int cg = C >> 2;
    const size_t total = (size_t)N * px * cg;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = idx % cg;
    const size_t p = idx / cg;
    const int n = p / px;
    const float4 v = *reinterpret_cast<const float4 *>(s + p * C + g * 4);
    const float4 gt = *reinterpret_cast<const float4 *>(gate + (size_t)n * C + g * 4);
    float4 o = make_float4(v.x * gt.x, v.y * gt.y, v.z * gt.z, v.w * gt.w);
    if (!first) {
        const float4 a = *reinterpret_cast<const float4 *>(x2 + p * C + g * 4);
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    }
    *reinterpret_cast<float4 *>(x2 + p * C + g * 4) = o;
}

__global__ void avgpool2_kernel(const float *__restrict__ in, int N, int H, int W, int C,
                                float *__restrict__ out) {
    const int Ho = H / 2, Wo = W / 2, cg = C >> 2;
    const size_t total = (size_t)N * Ho * Wo * cg;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = idx % cg;
    size_t p = idx / cg;
    const int x = p % Wo; p /= Wo;
    const int y = p % Ho;
    const int n = p / Ho;
    const float *b = in + (((size_t)n * H + 2 * y) * W + 2 * x) * C + g * 4;
    const float4 a0 = *reinterpret_cast<const float4 *>(b);
    const float4 a1 = *reinterpret_cast<const float4 *>(b + C);
    const float4 a2 = *reinterpret_cast<const float4 *>(b + (size_t)W * C);
    const float4 a3 = *reinterpret_cast<const float4 *>(b + (size_t)W * C + C);
    float4 o;
    o.x = (a0.x + a1.x + a2.x + a3.x) * 0.25f; o.y = (a0.y + a1.y + a2.y + a3.y) * 0.25f;
    o.z = (a0.z + a1.z + a2.z + a3.z) * 0.25f; o.w = (a0.w + a1.w + a2.w + a3.w) * 0.25f;
    *reinterpret_cast<float4 *>(out + (((size_t)n * Ho + y) * Wo + x) * C + g * 4) = o;
}

// fc: out[n][co] = relu( sum_k v[n][k] W[k][co] + b[co] ),  128 -> 512
__global__ void __launch_bounds__(256)
fc_kernel(const float *__restrict__ v, const float *__restrict__ wts, const float *__restrict__ bias,
          float *__restrict__ out) {
    __shared__ float s_v[128];
    const int n = blockIdx.x, tid = threadIdx.x;
    if (tid < 128) s_v[tid] = v[(size_t)n * 128 + tid];
    __syncthreads();
    for (int co = tid; co < 512; co += 256) {
        float acc = bias[co];
        for (int k = 0; k < 128; k++) acc = fmaf(s_v[k], wts[k * 512 + co], acc);
        out[(size_t)n * 512 + co] = fmaxf(acc, 0.f);
    }
}

// ---------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------

template <int CG>
static int launch_pw_t(const float *in, int P, int cin, const float *w, const float *b,
                       const float *res, float *out, int relu, cudaStream_t st) {
    const size_t smem = ((size_t)cin * CG * 4 + 64 * (cin + 1)) * sizeof(float);
    pw_conv_kernel<CG><<<(P + 63) / 64, 256, smem, st>>>(in, P, cin, w, b, res, out, relu);
    SSB_CHECK_LAUNCH();
    return 0;
}

static int reid_init_attrs() {
    static const int key = ssb_new_key();
    if (ssb_first_on_device(key)) {
        const int big = 128 * 1024;
        SSB_CHECK_CUDA(cudaFuncSetAttribute(pw_conv_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        SSB_CHECK_CUDA(cudaFuncSetAttribute(pw_conv_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        SSB_CHECK_CUDA(cudaFuncSetAttribute(pw_conv_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        SSB_CHECK_CUDA(cudaFuncSetAttribute(pw_conv_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        SSB_CHECK_CUDA(cudaFuncSetAttribute(pw_conv_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        SSB_CHECK_CUDA(cudaFuncSetAttribute(pw_conv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        SSB_CHECK_CUDA(cudaFuncSetAttribute(reid_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    }
    return 0;
}

static int launch_pw(const float *in, int P, int cin, int cout, const float *w, const float *b,
                     const float *res, float *out, int relu, cudaStream_t st) {
    switch (cout) {
        case 16: return launch_pw_t<4>(in, P, cin, w, b, res, out, relu, st);
        case 24: return launch_pw_t<6>(in, P, cin, w, b, res, out, relu, st);
        case 32: return launch_pw_t<8>(in, P, cin, w, b, res, out, relu, st);
        case 64: return launch_pw_t<16>(in, P, cin, w, b, res, out, relu, st);
        case 96: return launch_pw_t<24>(in, P, cin, w, b, res, out, relu, st);
        case 128: return launch_pw_t<32>(in, P, cin, w, b, res, out, relu, st);
    }
    ssb_set_error("pw conv: unsupported cout %d", cout);
    return -1;
}

// first weight-tensor index of OSBlock b in the canonical walk
static int block_tensor_index(int b) {
    int k = 2;
    for (int i = 0; i < b; i++) {
        const int cin = kBlocks[i][0], cout = kBlocks[i][1];
        k += 2 + 30 + 4 + 2 + (cin != cout ? 2 : 0) + ((i == 1 || i == 3) ? 2 : 0);
    }
    return k;
}

struct ReidBufs { float *DS, *X1, *T0, *T1, *PW, *X2, *GATE; };

static ReidBufs reid_bufs(ssb_tracker *t, int slot, int n, float **A, float **Bf) {
    float *ws = (slot & 1) ? t->reid_ws1 : t->reid_ws;
    const size_t N = (size_t)n;
    ReidBufs r;
    *A = ws;      ws += N * REID_BIG;
    *Bf = ws;     ws += N * REID_BIG;
    r.DS = ws;    ws += N * REID_BIG;
    r.X1 = ws;    ws += N * REID_MID;
    r.T0 = ws;    ws += N * REID_MID;
    r.T1 = ws;    ws += N * REID_MID;
    r.PW = ws;    ws += N * REID_MID;
    r.X2 = ws;    ws += N * REID_MID;
    r.GATE = ws;
    return r;
}

// one OSBlock, fp32 SIMT kernels: cur [n][Hc][Wc][cin] -> nxt [n][Hc][Wc][cout]
static int reid_block_simt(ssb_tracker *t, int b, const float *cur, float *nxt, int n, int Hc, int Wc,
                           const ReidBufs &B, cudaStream_t st) {
    const float *W = t->w_blob;
    const int64_t *off = t->w_off;
    int wi = block_tensor_index(b);
    auto nextw = [&]() { return W + off[wi++]; };
    const int cin = kBlocks[b][0], cout = kBlocks[b][1], mid = cout / 4;
    const int r = mid / 16 > 0 ? mid / 16 : 1;
    const int px = Hc * Wc, P = n * px;
    int rc;
    const float *c1w = nextw(), *c1b = nextw();
    rc = launch_pw(cur, P, cin, mid, c1w, c1b, nullptr, B.X1, 1, st);
    if (rc) return rc;
    const float *lw[10][3];
    for (int l = 0; l < 10; l++) { lw[l][0] = nextw(); lw[l][1] = nextw(); lw[l][2] = nextw(); }
    const float *g1w = nextw(), *g1b = nextw(), *g2w = nextw(), *g2b = nextw();
    const float *c3w = nextw(), *c3b = nextw();
    const float *dw_ = nullptr, *db_ = nullptr;
    if (cin != cout) { dw_ = nextw(); db_ = nextw(); }
    int l = 0;
    const int ew_blocks = (int)(((size_t)P * (mid / 4) + 255) / 256);
    for (int s = 0; s < 4; s++) {
        const float *src = B.X1;
        float *dst = B.T0;
        for (int k = 0; k <= s; k++, l++) {
            rc = launch_pw(src, P, mid, mid, lw[l][0], nullptr, nullptr, B.PW, 0, st);
            if (rc) return rc;
            dw3x3_kernel<<<ew_blocks, 256, 0, st>>>(B.PW, n, Hc, Wc, mid, lw[l][1], lw[l][2], dst);
            SSB_CHECK_LAUNCH();
            src = dst;
            dst = (dst == B.T0) ? B.T1 : B.T0;
        }
        gap_gate_kernel<<<n, 256, 0, st>>>(src, px, mid, g1w, g1b, g2w, g2b, r, B.GATE);
        SSB_CHECK_LAUNCH();
        gate_apply_kernel<<<ew_blocks, 256, 0, st>>>(src, B.GATE, n, px, mid, B.X2, s == 0);
        SSB_CHECK_LAUNCH();
    }
    const float *resid = cur;
    if (cin != cout) {
        rc = launch_pw(cur, P, cin, cout, dw_, db_, nullptr, B.DS, 0, st);
        if (rc) return rc;
        resid = B.DS;
    }
    return launch_pw(B.X2, P, mid, cout, c3w, c3b, resid, nxt, 1, st);
}

static int reid_block(ssb_tracker *t, int b, const float *cur, float *nxt, int n, int Hc, int Wc,
                      const ReidBufs &B, int use_tc, cudaStream_t st) {
    if (use_tc) {
        if (!t->w_tc) { ssb_set_error("tensor-core ReID weights not set"); return -1; }
        if (use_tc == 2) {
            if (!t->have_tc3) { ssb_set_error("pointwise/depthwise tensor-core weights (sections 10..15) not set"); return -1; }
            return ssb_reid_tc3_block(b, cur, nxt, t->w_tc + t->w_tc_off[10 + b], n, t->tc_status, st);
        }
        return ssb_reid_tc_block(b, cur, nxt, t->w_tc + t->w_tc_off[b], n, t->tc_status, st);
    }
    return reid_block_simt(t, b, cur, nxt, n, Hc, Wc, B, st);
}

// modes 0 (fp32 SIMT), 1 (9-tap tcgen05 OSBlocks), 2 (round-1 pointwise/depthwise OSBlocks): float32 NHWC activations
int ssb_reid_forward_baseline(ssb_tracker *t, int slot, const uint8_t *img, int h, int w, int pitch, const int *boxes,
                              int n, float *feats_out, cudaStream_t st) {
    if (n <= 0) return 0;
    if (!t->w_blob) { ssb_set_error("fp32 ReID weights not set (ssb_reid_set_weights)"); return -1; }
    { int rc = reid_init_attrs(); if (rc) return rc; }
    const float *W = t->w_blob;
    const int64_t *off = t->w_off;
    float *A, *Bf;
    const ReidBufs B = reid_bufs(t, slot, n, &A, &Bf);
    if (t->use_tc) {        // stem as 16 shifted GEMMs on the space-to-depth image (reid_tc.cu)
        int rc = ssb_reid_tc_stem(img, h, w, pitch, boxes, t->w_tc + t->w_tc_off[9], A, n, t->tc_status, st);
        if (rc) return rc;
    } else {
        const float *w0 = W + off[0], *b0 = W + off[1];
        const size_t smem = (ST_IN_FLOATS + ST_CONV * ST_CONV * 16 + 147 * 16) * sizeof(float);
        reid_stem_kernel<<<dim3(32, n), 256, smem, st>>>(img, h, w, pitch, boxes, w0, b0, A);
        SSB_CHECK_LAUNCH();
    }
    int Hc = 64, Wc = 32;
    float *cur = A, *nxt = Bf;
    for (int b = 0; b < 6; b++) {
        const int cout = kBlocks[b][1];
        int rc = reid_block(t, b, cur, nxt, n, Hc, Wc, B, t->use_tc, st);
        if (rc) return rc;
        { float *tmp = cur; cur = nxt; nxt = tmp; }
        if ((b == 1 || b == 3) && t->use_tc) {      // transition conv + ReLU + avgpool, one tcgen05 kernel
            const int a = b == 1 ? 0 : 1;
            rc = ssb_reid_tc_aux(a, cur, nxt, t->w_tc + t->w_tc_off[6 + a], n, t->tc_status, st);
            if (rc) return rc;
            { float *tmp = cur; cur = nxt; nxt = tmp; }
            Hc /= 2; Wc /= 2;
        } else if (b == 1 || b == 3) {
            const int wi = block_tensor_index(b + 1) - 2;
            const float *tw = W + off[wi], *tb = W + off[wi + 1];
            rc = launch_pw(cur, n * Hc * Wc, cout, cout, tw, tb, nullptr, nxt, 1, st);
            if (rc) return rc;
            const int blocks = (int)(((size_t)n * (Hc / 2) * (Wc / 2) * (cout / 4) + 255) / 256);
            avgpool2_kernel<<<blocks, 256, 0, st>>>(nxt, n, Hc, Wc, cout, cur);
            SSB_CHECK_LAUNCH();
            Hc /= 2; Wc /= 2;
        }
    }
    if (t->use_tc) {        // conv5 + GAP + fc in one tcgen05 kernel
        int rc = ssb_reid_tc_aux(2, cur, feats_out, t->w_tc + t->w_tc_off[8], n, t->tc_status, st);
        if (rc) return rc;
    } else {
        const int wi = block_tensor_index(6);
        const float *w5 = W + off[wi], *b5 = W + off[wi + 1];
        int rc = launch_pw(cur, n * Hc * Wc, 128, 128, w5, b5, nullptr, nxt, 1, st);
        if (rc) return rc;
        gap_gate_kernel<<<n, 256, 0, st>>>(nxt, Hc * Wc, 128, nullptr, nullptr, nullptr, nullptr, 0, B.GATE);
        SSB_CHECK_LAUNCH();
        const float *fw = W + off[wi + 2], *fb = W + off[wi + 3];
        fc_kernel<<<n, 256, 0, st>>>(B.GATE, fw, fb, feats_out);
        SSB_CHECK_LAUNCH();
        if (wi + 4 != t->n_w) { ssb_set_error("internal: weight walk ends at %d of %d tensors", wi + 4, t->n_w); return -4; }
    }
    return 0;
}

extern "C" int ssb_reid_use_tc(ssb_tracker *t, int enable) {
    if (!t) { ssb_set_error("null handle"); return -1; }
    if (enable && !t->w_tc) { ssb_set_error("tensor-core ReID weights not set"); return -1; }
    if (enable < 0 || enable > 3) { ssb_set_error("ReID mode must be 0 (simt), 1 (tc, 9-tap), 2 (tc, pointwise + depthwise) or 3 (operand planes + halo exchange)"); return -1; }
    if (enable >= 2 && !t->have_tc3) { ssb_set_error("pointwise/depthwise tensor-core weights (sections 10..15) not set"); return -1; }
    t->use_tc = enable;
    return 0;
}

// y = OSBlock_b(x): x [n][H][W][cin] float32 NHWC, y [n][H][W][cout]; use_tc picks the path.
// status_out (device int, may be NULL) receives the tensor-core path's timeout flag.
extern "C" int ssb_reid_block(ssb_tracker *t, int block, const float *x_dev, float *y_dev, int n,
                              int use_tc, ssb_stream_t stream) {
    if (!t || !x_dev || !y_dev) { ssb_set_error("null argument"); return -1; }
    if (block < 0 || block > 5 || n < 1 || n > t->dims.N) { ssb_set_error("bad block / n"); return -1; }
    if (!t->w_blob) { ssb_set_error("ReID weights not set"); return -1; }
    { int rc = reid_init_attrs(); if (rc) return rc; }
    float *A, *Bf;
    const ReidBufs B = reid_bufs(t, 0, n, &A, &Bf);
    const int Hc = block < 2 ? 64 : (block < 4 ? 32 : 16), Wc = Hc / 2;
    if (use_tc == 3) {          // operand-plane kernels: convert in, run, convert out
        if (!t->w_tc || !t->have_tc3) { ssb_set_error("pointwise/depthwise tensor-core weights (sections 10..15) not set"); return -1; }
        cudaStream_t st = (cudaStream_t)stream;
        const int cin = kBlocks[block][0], cout = kBlocks[block][1];
        int rc = ssb_reid_nhwc_to_planes(x_dev, A, n, Hc * Wc, cin, st);
        if (rc) return rc;
        rc = ssb_reid_tc4_block(block, A, Bf, t->w_tc + t->w_tc_off[10 + block], n, t->tc_status, st);
        if (rc) return rc;
        return ssb_reid_planes_to_nhwc(Bf, y_dev, n, Hc * Wc, cout, st);
    }
    return reid_block(t, block, x_dev, y_dev, n, Hc, Wc, B, use_tc, (cudaStream_t)stream);
}

// diagnostic: CTA 0 of every tensor-core OSBlock launch writes clock64() phase stamps to
// buf_dev[0..63] (buf_dev[0] = number of stamps); pass NULL to switch it off.
extern long long *g_ssb_tc_dbg;
extern "C" int ssb_reid_tc_debug(void *buf_dev) {
    g_ssb_tc_dbg = (long long *)buf_dev;
    return 0;
}

#endif  // SSB_BASELINES
