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

extern "C" int ssb_appearance_cost(const float *gallery_dev, const int32_t *counts_dev,
                                   int n_tracks, int budget, const float *feats_dev, int n_dets,
                                   int dim, float *cost_out_dev, ssb_stream_t stream) {
    return ssb_launch_appearance(gallery_dev, counts_dev, nullptr, nullptr, nullptr, nullptr,
                                 n_tracks, budget, feats_dev, n_dets, dim, cost_out_dev, n_dets,
                                 (cudaStream_t)stream);
}
