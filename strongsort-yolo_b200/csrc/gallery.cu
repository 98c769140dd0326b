// Optional cross-stream ReID gallery (BASELINE.json config C5).  The reference
// defines no semantics for it (its workers share nothing, yolo_multi_model.py:351-354), so it
// is READ-ONLY here: every stream exports the EMA appearance vector of its confirmed tracks,
// the per-GPU exports are all-gathered over NCCL / NVLink by the host (dist.SharedGallery),
// and each stream reports, per local track, the nearest track of any OTHER stream in cosine
// distance.  Per-stream track ids and the parity with the oracle are untouched.
//
//   ssb_gallery_export      confirmed tracks (list order) -> feat [t_max][D], ids [t_max], count
//   ssb_gallery_cross_match local [t_max][D] x gathered [G][t_max][D] -> nearest foreign track
//
// HBM/L2-bound: G * t_max * D * 4 bytes (4.2 MB for 8 x 256 x 512) streamed once per local track
// block from L2; 2 * T * G * t_max * D flops (0.5 GFLOP) -- microseconds, off the per-frame path.
#include "ssb_common.cuh"

namespace {

__global__ void gallery_export_kernel(TrackTable tt, int D, int t_max, float *__restrict__ feat,
                                      int *__restrict__ ids, int *__restrict__ count) {
    // one block; ordered compaction of the confirmed tracks by a block-wide ballot scan
    __shared__ int s_base;
    __shared__ int s_warp[32];
    const int n = tt.scalars[SC_N_TRACKS];
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int p0 = 0; p0 < n; p0 += blockDim.x) {
        const int p = p0 + threadIdx.x;
        const int slot = p < n ? tt.order[p] : -1;
        const bool keep = slot >= 0 && tt.state[slot] == SSB_CONFIRMED;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; w++) off += s_warp[w];
        const int dst = off + __popc(m & ((1u << lane) - 1));
        if (keep && dst < t_max) ids[dst] = tt.track_id[slot];
        // the feature rows are copied cooperatively below: remember the slot of every kept row
        if (keep && dst < t_max) ids[t_max + dst] = slot;            // scratch half of ids (see launcher)
        __syncthreads();
        if (threadIdx.x == 0) { int tot = 0; for (int w = 0; w < nw; w++) tot += s_warp[w]; s_base += tot; }
        __syncthreads();
    }
    const int cnt = s_base < t_max ? s_base : t_max;
    if (threadIdx.x == 0) *count = cnt;
    for (int r = threadIdx.x; r < t_max - cnt; r += blockDim.x) ids[cnt + r] = -1;
}

__global__ void gallery_copy_kernel(TrackTable tt, int D, int t_max, const int *__restrict__ ids,
                                    const int *__restrict__ count, float *__restrict__ feat) {
    const int r = blockIdx.x;
    const bool live = r < *count;
    const int slot = live ? ids[t_max + r] : 0;
    for (int k = threadIdx.x; k < D; k += blockDim.x)
        feat[(size_t)r * D + k] = live ? tt.feat[(size_t)slot * D + k] : 0.f;
}

// one block per local track; warp w scans foreign rows w, w + nw, ...; lanes split the D dimension
__global__ void gallery_cross_match_kernel(const float *__restrict__ local, const int *__restrict__ local_ids,
                                           const float *__restrict__ all, const int *__restrict__ all_ids,
                                           int n_ranks, int self_rank, int t_max, int D, float max_dist,
                                           int *__restrict__ m_rank, int *__restrict__ m_id, float *__restrict__ m_dist,
                                           size_t feat_rank_stride, size_t ids_rank_stride) {
    const int i = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __shared__ float s_d[32];
    __shared__ int s_j[32];
    float best = INFINITY;
    int best_j = -1;
    if (local_ids[i] >= 0) {
        const float *f = local + (size_t)i * D;
        for (int j = warp; j < n_ranks * t_max; j += nw) {
            const int jr = j / t_max, jt = j - jr * t_max;
            if (jr == self_rank || all_ids[jr * ids_rank_stride + jt] < 0) continue;     // warp-uniform
            const float *g = all + jr * feat_rank_stride + (size_t)jt * D;
            float acc = 0.f;
            for (int k = lane * 4; k < D; k += 128) {
                const float4 a = *reinterpret_cast<const float4 *>(f + k);
                const float4 b = *reinterpret_cast<const float4 *>(g + k);
                acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
                acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            const float d = 1.f - acc;                                   // both sides are unit vectors
            if (d < best) { best = d; best_j = j; }                      // ascending j: first minimum wins
        }
    }
    if (lane == 0) { s_d[warp] = best; s_j[warp] = best_j; }
    __syncthreads();
    if (threadIdx.x == 0) {
        best = INFINITY;
        best_j = -1;
        for (int w = 0; w < nw; w++)        // minimum by (distance, flat index): lowest rank / row wins ties
            if (s_j[w] >= 0 && (best_j < 0 || s_d[w] < best || (s_d[w] == best && s_j[w] < best_j))) { best = s_d[w]; best_j = s_j[w]; }
        const bool hit = best_j >= 0 && best <= max_dist;
        m_rank[i] = hit ? best_j / t_max : -1;
        m_id[i] = hit ? all_ids[(best_j / t_max) * ids_rank_stride + best_j % t_max] : -1;
        m_dist[i] = best_j >= 0 ? best : INFINITY;
    }
}

}  // namespace

extern "C" int ssb_gallery_export(ssb_tracker *t, int t_max, float *feat_out_dev, int32_t *ids_out_dev,
                                  int32_t *count_out_dev, ssb_stream_t stream) {
    if (!t || !feat_out_dev || !ids_out_dev || !count_out_dev || t_max < 1) { ssb_set_error("bad argument"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    // ids_out_dev must hold 2 * t_max ints: [ids | slot scratch]
    gallery_export_kernel<<<1, 256, 0, st>>>(t->tt, t->dims.D, t_max, feat_out_dev, ids_out_dev, count_out_dev);
    SSB_CHECK_LAUNCH();
    gallery_copy_kernel<<<t_max, 128, 0, st>>>(t->tt, t->dims.D, t_max, ids_out_dev, count_out_dev, feat_out_dev);
    SSB_CHECK_LAUNCH();
    return 0;
}

extern "C" int ssb_gallery_cross_match(const float *local_feat_dev, const int32_t *local_ids_dev,
                                       const float *all_feat_dev, const int32_t *all_ids_dev, int n_ranks,
                                       int self_rank, int t_max, int dim, float max_dist,
                                       int32_t *match_rank_out_dev, int32_t *match_id_out_dev,
                                       float *match_dist_out_dev, ssb_stream_t stream) {
    if (!local_feat_dev || !local_ids_dev || !all_feat_dev || !all_ids_dev || !match_rank_out_dev ||
        !match_id_out_dev || !match_dist_out_dev) { ssb_set_error("null argument"); return -1; }
    if (n_ranks < 1 || self_rank < 0 || self_rank >= n_ranks || t_max < 1 || dim % 128) {
        ssb_set_error("bad gallery geometry (dim must be a multiple of 128)");
        return -1;
    }
    gallery_cross_match_kernel<<<t_max, 256, 0, (cudaStream_t)stream>>>(
        local_feat_dev, local_ids_dev, all_feat_dev, all_ids_dev, n_ranks, self_rank, t_max, dim, max_dist,
        match_rank_out_dev, match_id_out_dev, match_dist_out_dev, (size_t)t_max * dim, (size_t)t_max);
    SSB_CHECK_LAUNCH();
    return 0;
}

// The same match on the PACKED exchange buffer: per rank  [t_max * dim float32 features | t_max int32 ids]
// (t_max * (dim + 1) 4-byte words), so that one collective (or one peer copy) moves a stream's whole export.
extern "C" int ssb_gallery_cross_match_packed(const void *all_packed_dev, int n_ranks, int self_rank, int t_max, int dim,
                                              float max_dist, int32_t *match_rank_out_dev, int32_t *match_id_out_dev,
                                              float *match_dist_out_dev, ssb_stream_t stream) {
    if (!all_packed_dev || !match_rank_out_dev || !match_id_out_dev || !match_dist_out_dev) { ssb_set_error("null argument"); return -1; }
    if (n_ranks < 1 || self_rank < 0 || self_rank >= n_ranks || t_max < 1 || dim % 128) {
        ssb_set_error("bad gallery geometry (dim must be a multiple of 128)");
        return -1;
    }
    const size_t L = (size_t)t_max * (dim + 1);
    const float *base = (const float *)all_packed_dev;
    const int *ids = (const int *)all_packed_dev + (size_t)t_max * dim;
    gallery_cross_match_kernel<<<t_max, 256, 0, (cudaStream_t)stream>>>(
        base + self_rank * L, ids + self_rank * L, base, ids, n_ranks, self_rank, t_max, dim, max_dist,
        match_rank_out_dev, match_id_out_dev, match_dist_out_dev, L, L);
    SSB_CHECK_LAUNCH();
    return 0;
}
