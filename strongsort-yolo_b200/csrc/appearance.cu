// Appearance cost of StrongSORT stage A:
//   cost[t][n] = min_b ( 1 - <g_tb/|g_tb| , f_n/|f_n|> )
// over the <= budget gallery samples of every confirmed track.
//
// Roofline: at N=100 the arithmetic intensity is ~N/2 flop/B in
// fp32, far under the ridge -> HBM-bound on the gallery read
// ((T*B + N) * D * 4 bytes per frame).  fp32 FMA is kept on purpose: the
// result feeds exact threshold tests (> 0.2) and the assignment must equal the
// NumPy oracle's, so no tf32/bf16 rounding is allowed here.
//
// Kernel: one CTA per (track row, 64-det tile): a 128(b) x 64(n) x 512(k)
// SIMT GEMM tile, gallery rows streamed once with float4 loads, row norms
// accumulated on the fly, min over b in the epilogue.
#include "ssb_common.cuh"
#include "tc_common.cuh"

#define AP_BM 128
#define AP_BN 64
#define AP_BK 16
#define AP_THREADS 256

__global__ void __launch_bounds__(AP_THREADS)
appearance_cost_kernel(const float *__restrict__ gallery, const int *__restrict__ gal_count,
                       const int *__restrict__ row_pos_list, const int *__restrict__ order,
                       const int *__restrict__ n_rows_dev, int n_rows_host, int budget,
                       const float *__restrict__ feats, int n_dets, int D,
                       float *__restrict__ cost, int ld) {
    __shared__ float As[AP_BK][AP_BM + 4];
    __shared__ float Bs[AP_BK][AP_BN + 4];
    __shared__ float s_gn[AP_BM];       // gallery row norms^2
    __shared__ float s_fn[AP_BN];       // det norms^2
    __shared__ float s_min[AP_THREADS / 16][AP_BN];

    const int r = blockIdx.x;
    const int n_rows = n_rows_dev ? *n_rows_dev : n_rows_host;
    if (r >= n_rows) return;
    const int slot = row_pos_list ? order[row_pos_list[r]] : r;
    const int cnt = min(gal_count[slot], budget);
    const int n0 = blockIdx.y * AP_BN;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;         // 16 x 16 thread grid
    const float *G = gallery + (size_t)slot * budget * D;

    float best[4];
#pragma unroll
    for (int j = 0; j < 4; j++) best[j] = INFINITY;

    for (int b0 = 0; b0 < cnt; b0 += AP_BM) {
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
        float gsq = 0.f, fsq = 0.f;
        // loader mapping: A: thread -> row tid/2, half (tid&1)*8 ; B: threads < 128 -> row tid/2
        const int arow = tid >> 1, ahalf = (tid & 1) * 8;
        const bool a_ok = (b0 + arow) < cnt;
        const bool b_ok = tid < 2 * AP_BN && (n0 + arow) < n_dets;
        const float *ap = G + (size_t)(b0 + arow) * D + ahalf;
        const float *bp = feats + (size_t)(n0 + arow) * D + ahalf;
        // register double buffering: the loads of K chunk k0+BK are in flight while chunk k0
        // is multiplied (the kernel was latency-bound: 82 us for 20.7 MB before this)
        float4 a0 = make_float4(0, 0, 0, 0), a1 = a0, q0 = a0, q1 = a0;
        if (a_ok) {
            a0 = *reinterpret_cast<const float4 *>(ap);
            a1 = *reinterpret_cast<const float4 *>(ap + 4);
        }
        if (b_ok) {
            q0 = *reinterpret_cast<const float4 *>(bp);
            q1 = *reinterpret_cast<const float4 *>(bp + 4);
        }
        for (int k0 = 0; k0 < D; k0 += AP_BK) {
            __syncthreads();
            As[ahalf + 0][arow] = a0.x; As[ahalf + 1][arow] = a0.y;
            As[ahalf + 2][arow] = a0.z; As[ahalf + 3][arow] = a0.w;
            As[ahalf + 4][arow] = a1.x; As[ahalf + 5][arow] = a1.y;
            As[ahalf + 6][arow] = a1.z; As[ahalf + 7][arow] = a1.w;
            gsq += a0.x * a0.x + a0.y * a0.y + a0.z * a0.z + a0.w * a0.w +
                   a1.x * a1.x + a1.y * a1.y + a1.z * a1.z + a1.w * a1.w;
            if (tid < 2 * AP_BN) {
                Bs[ahalf + 0][arow] = q0.x; Bs[ahalf + 1][arow] = q0.y;
                Bs[ahalf + 2][arow] = q0.z; Bs[ahalf + 3][arow] = q0.w;
                Bs[ahalf + 4][arow] = q1.x; Bs[ahalf + 5][arow] = q1.y;
                Bs[ahalf + 6][arow] = q1.z; Bs[ahalf + 7][arow] = q1.w;
                fsq += q0.x * q0.x + q0.y * q0.y + q0.z * q0.z + q0.w * q0.w +
                       q1.x * q1.x + q1.y * q1.y + q1.z * q1.z + q1.w * q1.w;
            }
            __syncthreads();
            if (k0 + AP_BK < D) {            // prefetch the next chunk
                if (a_ok) {
                    a0 = *reinterpret_cast<const float4 *>(ap + k0 + AP_BK);
                    a1 = *reinterpret_cast<const float4 *>(ap + k0 + AP_BK + 4);
                }
                if (b_ok) {
                    q0 = *reinterpret_cast<const float4 *>(bp + k0 + AP_BK);
                    q1 = *reinterpret_cast<const float4 *>(bp + k0 + AP_BK + 4);
                }
            }
#pragma unroll
            for (int k = 0; k < AP_BK; k++) {
                float a[8], b[4];
                const float4 av0 = *reinterpret_cast<const float4 *>(&As[k][ty * 8]);
                const float4 av1 = *reinterpret_cast<const float4 *>(&As[k][ty * 8 + 4]);
                const float4 bv = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
                a[0] = av0.x; a[1] = av0.y; a[2] = av0.z; a[3] = av0.w;
                a[4] = av1.x; a[5] = av1.y; a[6] = av1.z; a[7] = av1.w;
                b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
        // row norms: the two half-row loaders are adjacent lanes
        gsq += __shfl_xor_sync(0xffffffffu, gsq, 1);
        fsq += __shfl_xor_sync(0xffffffffu, fsq, 1);
        __syncthreads();
        if ((tid & 1) == 0) {
            s_gn[arow] = gsq;
            if (tid < 2 * AP_BN) s_fn[arow] = fsq;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int b = b0 + ty * 8 + i;
            if (b >= cnt) continue;
            const float gn = sqrtf(s_gn[ty * 8 + i]);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float fn = sqrtf(s_fn[tx * 4 + j]);
                const float dist = 1.0f - acc[i][j] / (gn * fn);
                best[j] = fminf(best[j], dist);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; j++) s_min[ty][tx * 4 + j] = best[j];
    __syncthreads();
    if (tid < AP_BN) {
        float m = s_min[0][tid];
#pragma unroll
        for (int y = 1; y < AP_THREADS / 16; y++) m = fminf(m, s_min[y][tid]);
        if (n0 + tid < n_dets) cost[(size_t)r * ld + n0 + tid] = m;
    }
}

int ssb_launch_appearance(const float *gallery, const int *gal_count, const int * /*gal_head*/,
                          const int *row_pos_list, const int *order, const int *n_rows_dev,
                          int max_rows, int budget, const float *feats, int n_dets, int dim,
                          float *cost, int ld, cudaStream_t st) {
    if (max_rows <= 0 || n_dets <= 0) return 0;
    if (dim % AP_BK != 0 || dim % 4 != 0) {
        ssb_set_error("appearance: feature dim %d must be a multiple of %d", dim, AP_BK);
        return -1;
    }
    dim3 grid(max_rows, (n_dets + AP_BN - 1) / AP_BN);
    appearance_cost_kernel<<<grid, AP_THREADS, 0, st>>>(gallery, gal_count, row_pos_list, order,
                                                        n_rows_dev, max_rows, budget, feats,
                                                        n_dets, dim, cost, ld);
    SSB_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Tensor-core path (the tracker's default).  Both sides of the contraction are stored as what the
// tensor core reads: unit vectors * 2^6 split into fp16 (hi, lo) pairs (hi + lo = value to ~2^-22;
// the scale keeps the lo halves out of fp16's subnormal range; the product carries 2^12, removed
// exactly in the epilogue) in K-major no-swizzle operand planes -- the gallery ring at
// gallery_append time ([slot][hl][64][128 rows][8], tracker.cu), the frame's embeddings in
// det_norm_kernel ([hl][64][npad][8]).  One CTA OWNS A TRACK FOR ALL DETECTIONS: its gallery is
// streamed from HBM exactly once (K in stages of 8 * KCH, one cp.async.bulk per (operand, hl) and
// stage, 3-4 stages in flight), the detections come from L2;  D[det][b] = f_det . g_b accumulates in
// TMEM (M = 128 detections per tile, up to 4 tiles = 512 columns, N = budget rounded up to 16, a
// product = Ah.Bh + Al.Bh + Ah.Bl), and the epilogue takes max_b over the valid ring entries in
// registers (lane = detection): cost = 1 - 2^-12 * max_b.
// HBM bound: (T * 128 + npad) * 512 * 4 bytes -> 26.5 MB at C2.
// ---------------------------------------------------------------------------------------------
template <int NTILES>
struct AptCfg {
    // Stage sizes chosen so that the CTA (80-96 KB) fits the slot a retiring stage-2 ReID CTA leaves: in the two-stage
    // pipeline the previous frame's association runs under the next frame's embedding, and a 160-192 KB request
    // (SSB_APT_BIG: 8 / NTILES chunks per stage) would wait for an SM to drain completely -- which a many-wave ReID
    // kernel never lets happen.
#ifdef SSB_APT_BIG
    static constexpr int KCH = 8 / NTILES;                 // 16-byte K chunks per stage
    static constexpr int NSTAGE = NTILES == 1 ? 3 : 4;
#else
    static constexpr int KCH = NTILES == 1 ? 4 : 2;
    static constexpr int NSTAGE = NTILES == 1 ? 3 : NTILES == 2 ? 4 : 2;
#endif
    static constexpr int NPAD = NTILES * 128;
    static constexpr int A_HALF_B = KCH * NPAD * 16, A_B = 2 * A_HALF_B;            // detections
    static constexpr int B_HALF_B = KCH * SSB_GAL_ROWS * 16, B_B = 2 * B_HALF_B;    // gallery
    static constexpr int STAGE_B = A_B + B_B;
    static constexpr int NSTEPS = 64 / KCH;
    static constexpr int SMEM_B = NSTAGE * STAGE_B + 256;
    static constexpr int TM_COLS = NTILES * 128;
};

template <int NTILES>
__global__ void __launch_bounds__(128, 1)
appearance_tc_kernel(const unsigned char *__restrict__ gal_planes, const int *__restrict__ gal_count,
                     const int *__restrict__ row_pos_list, const int *__restrict__ order,
                     const int *__restrict__ n_rows_dev, int budget, const unsigned char *__restrict__ det_planes,
                     int n_dets, float *__restrict__ cost, int ld, int *__restrict__ status) {
    using C = AptCfg<NTILES>;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int r = blockIdx.x;
    if (n_rows_dev && r >= *n_rows_dev) return;            // whole CTA, before any barrier / TMEM allocation
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const int slot = row_pos_list ? order[row_pos_list[r]] : r;
    const int cnt = min(gal_count[slot], budget);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + C::NSTAGE * C::STAGE_B);
    uint64_t *empty = full + C::NSTAGE, *done = empty + C::NSTAGE;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(done + 1);
    if (warp == 0) tc::tmem_alloc(s_tmem, C::TM_COLS);
    if (tid == 0) {
        for (int i = 0; i < C::NSTAGE; i++) { tc::mbar_init(full + i, 1); tc::mbar_init(empty + i, 1); }
        tc::mbar_init(done, 1);
        tc::fence_mbar_init();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *s_tmem;
    bool ok = true;
    const unsigned char *gsrc = gal_planes + (size_t)slot * (2 * 64 * SSB_GAL_ROWS * 16);
    if (warp_u == 0) {                                     // producer: one elected lane drives the TMA engine
        if (tc::elect_one()) {
            for (int it = 0; it < C::NSTEPS; it++) {
                const int sg = it % C::NSTAGE;
                if (it >= C::NSTAGE && !tc::mbar_wait(empty + sg, ((it / C::NSTAGE) - 1) & 1)) ok = false;
                unsigned char *sa = smem + sg * C::STAGE_B, *sb = sa + C::A_B;
                tc::mbar_arrive_expect_tx(full + sg, C::STAGE_B);
                const size_t c0 = (size_t)it * C::KCH;
#pragma unroll
                for (int hl = 0; hl < 2; hl++) {
                    tc::bulk_g2s(sb + hl * C::B_HALF_B, gsrc + (size_t)hl * (64 * SSB_GAL_ROWS * 16) + c0 * SSB_GAL_ROWS * 16,
                                 C::B_HALF_B, full + sg);
                    tc::bulk_g2s(sa + hl * C::A_HALF_B, det_planes + (size_t)hl * (64 * C::NPAD * 16) + c0 * C::NPAD * 16,
                                 C::A_HALF_B, full + sg);
                }
            }
        }
        __syncwarp();
    } else if (warp_u == 1) {                              // MMA issuer
        if (tc::elect_one()) {
            const uint32_t idesc = tc::make_idesc_f16(128, (budget + 15) & ~15);
            for (int it = 0; it < C::NSTEPS; it++) {
                const int sg = it % C::NSTAGE;
                if (!tc::mbar_wait(full + sg, (it / C::NSTAGE) & 1)) ok = false;
                tc::fence_after_sync();
                const uint32_t sa = tc::smem_u32(smem + sg * C::STAGE_B), sb = sa + C::A_B;
#pragma unroll
                for (int t = 0; t < NTILES; t++) {
#pragma unroll
                    for (int ks = 0; ks < C::KCH / 2; ks++) {
                        const uint64_t ah = tc::make_smem_desc(sa + t * 2048 + ks * 2 * C::NPAD * 16, C::NPAD * 16, 128);
                        const uint64_t al = tc::make_smem_desc(sa + C::A_HALF_B + t * 2048 + ks * 2 * C::NPAD * 16, C::NPAD * 16, 128);
                        const uint64_t bh = tc::make_smem_desc(sb + ks * 2 * SSB_GAL_ROWS * 16, SSB_GAL_ROWS * 16, 128);
                        const uint64_t bl = tc::make_smem_desc(sb + C::B_HALF_B + ks * 2 * SSB_GAL_ROWS * 16, SSB_GAL_ROWS * 16, 128);
                        const uint32_t d = tmem + t * 128;
                        tc::mma_f16_ss(d, ah, bh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
                        tc::mma_f16_ss(d, al, bh, idesc, 1u);
                        tc::mma_f16_ss(d, ah, bl, idesc, 1u);
                    }
                }
                tc::mma_commit(empty + sg);                // stage free once these MMAs have read it
            }
            tc::mma_commit(done);
        }
        __syncwarp();
    }
    if (!tc::mbar_wait(done, 0)) ok = false;
    tc::fence_after_sync();
    const int gn = (budget + 15) & ~15;
#pragma unroll
    for (int t = 0; t < NTILES; t++) {
        const int det = t * 128 + warp * 32 + lane;
        if (t * 128 >= n_dets) break;                      // warp-uniform: no detections in this tile
        float best = -INFINITY;
        for (int c0 = 0; c0 < gn; c0 += 16) {
            float v[16];
            tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + t * 128 + c0, v);
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (c0 + j < cnt) best = fmaxf(best, v[j]);
        }
        if (det < n_dets) cost[(size_t)r * ld + det] = 1.0f - best * (1.0f / 4096.0f);
    }
    if (!ok && tid == 0) atomicExch(status, 7);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, C::TM_COLS);
}

template <int NTILES>
static int launch_app_tc(const unsigned char *gal_planes, const int *gal_count, const int *row_pos_list,
                         const int *order, const int *n_rows_dev, int max_rows, int budget,
                         const unsigned char *det_planes, int n_dets, float *cost, int ld, int *status, cudaStream_t st) {
    using C = AptCfg<NTILES>;
    static const int key = ssb_new_key();
    if (ssb_first_on_device(key))
        SSB_CHECK_CUDA(cudaFuncSetAttribute(appearance_tc_kernel<NTILES>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_B));
    appearance_tc_kernel<NTILES><<<max_rows, 128, C::SMEM_B, st>>>(gal_planes, gal_count, row_pos_list, order, n_rows_dev,
                                                                  budget, det_planes, n_dets, cost, ld, status);
    SSB_CHECK_LAUNCH();
    return 0;
}

int ssb_launch_appearance_tc(const unsigned char *gal_planes, const int *gal_count, const int *row_pos_list,
                             const int *order, const int *n_rows_dev, int max_rows, int budget,
                             const unsigned char *det_planes, int n_dets, float *cost, int ld, int *status,
                             cudaStream_t st) {
    if (max_rows <= 0 || n_dets <= 0) return 0;
    if (budget > SSB_GAL_ROWS || n_dets > SSB_DET_PLANES_MAX) { ssb_set_error("appearance_tc: problem exceeds the operand planes"); return -1; }
    const int npad = ssb_det_npad(n_dets);
    if (npad == 128) return launch_app_tc<1>(gal_planes, gal_count, row_pos_list, order, n_rows_dev, max_rows, budget, det_planes, n_dets, cost, ld, status, st);
    if (npad == 256) return launch_app_tc<2>(gal_planes, gal_count, row_pos_list, order, n_rows_dev, max_rows, budget, det_planes, n_dets, cost, ld, status, st);
    return launch_app_tc<4>(gal_planes, gal_count, row_pos_list, order, n_rows_dev, max_rows, budget, det_planes, n_dets, cost, ld, status, st);
}

// ---- stage entry point of the tensor-core path on caller arrays (parity tests): the float32 gallery /
//      embeddings are first re-laid as operand planes into caller scratch (what gallery_append_kernel and
//      det_norm_kernel do inside the tracker), then the same kernel runs
__global__ void rows_to_planes_kernel(const float *__restrict__ x, int rows, int rows_per_group, int group_rows_pad,
                                      size_t group_bytes, unsigned char *__restrict__ planes) {
    // one warp per row of 512 floats; row r belongs to group r / rows_per_group (a track's ring, or the frame)
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= rows) return;
    const float *f = x + (size_t)w * 512;
    float ss = 0.f;
    for (int k = lane; k < 512; k += 32) ss += f[k] * f[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float nrm = sqrtf(ss);
    const int g = w / rows_per_group, rr = w - g * rows_per_group;
    unsigned char *base = planes + (size_t)g * group_bytes;
    for (int c = lane; c < 64; c += 32) {
        __align__(16) __half h[8];
        __align__(16) __half l[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float u = (f[c * 8 + j] / nrm) * 64.0f;
            h[j] = __float2half_rn(u);
            l[j] = __float2half_rn(u - __half2float(h[j]));
        }
        unsigned char *dst = base + ((size_t)c * group_rows_pad + rr) * 16;
        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<uint4 *>(h);
        *reinterpret_cast<uint4 *>(dst + (size_t)64 * group_rows_pad * 16) = *reinterpret_cast<uint4 *>(l);
    }
}

extern "C" int64_t ssb_appearance_tc_scratch_bytes(int n_tracks) {
    if (n_tracks < 0) return -1;
    return (int64_t)n_tracks * (2 * 64 * SSB_GAL_ROWS * 16) + (int64_t)2 * 64 * SSB_DET_PLANES_MAX * 16 + 256;
}

extern "C" int ssb_appearance_cost_tc(const float *gallery_dev, const int32_t *counts_dev, int n_tracks, int budget,
                                      const float *feats_dev, int n_dets, int dim, float *cost_out_dev,
                                      void *scratch_dev, int32_t *status_dev, ssb_stream_t stream) {
    if (!gallery_dev || !counts_dev || !feats_dev || !cost_out_dev || !scratch_dev || !status_dev) { ssb_set_error("null argument"); return -1; }
    if (dim != 512 || budget < 1 || budget > SSB_GAL_ROWS || n_dets > SSB_DET_PLANES_MAX) {
        ssb_set_error("appearance_tc: dim must be 512, budget <= %d, n_dets <= %d", SSB_GAL_ROWS, SSB_DET_PLANES_MAX);
        return -1;
    }
    if (n_tracks <= 0 || n_dets <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char *gal_planes = (unsigned char *)scratch_dev;
    unsigned char *det_planes = gal_planes + (size_t)n_tracks * (2 * 64 * SSB_GAL_ROWS * 16);
    const int npad = ssb_det_npad(n_dets);
    const int grows = n_tracks * budget;
    rows_to_planes_kernel<<<(grows * 32 + 127) / 128, 128, 0, st>>>(gallery_dev, grows, budget, SSB_GAL_ROWS,
                                                                  (size_t)2 * 64 * SSB_GAL_ROWS * 16, gal_planes);
    SSB_CHECK_LAUNCH();
    rows_to_planes_kernel<<<(n_dets * 32 + 127) / 128, 128, 0, st>>>(feats_dev, n_dets, n_dets, npad, 0, det_planes);
    SSB_CHECK_LAUNCH();
    return ssb_launch_appearance_tc(gal_planes, counts_dev, nullptr, nullptr, nullptr, n_tracks, budget, det_planes,
                                    n_dets, cost_out_dev, n_dets, status_dev, st);
}

extern "C" int ssb_appearance_cost(const float *gallery_dev, const int32_t *counts_dev,
                                   int n_tracks, int budget, const float *feats_dev, int n_dets,
                                   int dim, float *cost_out_dev, ssb_stream_t stream) {
    return ssb_launch_appearance(gallery_dev, counts_dev, nullptr, nullptr, nullptr, nullptr,
                                 n_tracks, budget, feats_dev, n_dets, dim, cost_out_dev, n_dets,
                                 (cudaStream_t)stream);
}
