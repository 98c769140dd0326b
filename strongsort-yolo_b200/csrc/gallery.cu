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

// ---------------------------------------------------------------------------------------------------------
// Peer-memory exchange: the match reads every OTHER stream's packed export straight out of that GPU's memory
// over NVLink (symmetric-memory peer pointers) -- no all-gather, no staging copy, no NCCL kernel.  A block owns
// a tile of RT foreign rows: it pulls them across NVLink ONCE (coalesced 16-byte peer loads into shared
// memory; the whole exchange moves (G - 1) * t_max * (D + 1) * 4 bytes per rank per frame = 3.6 MB at G = 8),
// then sweeps all local tracks against the tile (local rows come from the local L2) and folds every distance
// into a per-local-track 64-bit key (order-preserving distance bits << 32 | flat foreign index) with
// atomicMin -- the same "lowest distance, then lowest rank / row" rule as gallery_cross_match_kernel, with the
// same per-pair summation order, so both paths give identical results.
// ---------------------------------------------------------------------------------------------------------
#define GAL_RT 8
__device__ __forceinline__ unsigned long long gal_key(float d, int j) {
    unsigned u = __float_as_uint(d);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned)j;
}
__device__ __forceinline__ float gal_key_dist(unsigned long long k) {
    unsigned u = (unsigned)(k >> 32);
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}

__global__ void gallery_peer_init_kernel(unsigned long long *best, int t_max) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < t_max) best[i] = ~0ull;
}

__global__ void __launch_bounds__(256)
gallery_peer_match_kernel(const unsigned long long *__restrict__ peer_ptrs, int n_ranks, int self_rank, int t_max, int D,
                          unsigned long long *__restrict__ best) {
    extern __shared__ float s_tile[];                  // [GAL_RT][D]
    __shared__ int s_ids[GAL_RT];
    const int tiles = (t_max + GAL_RT - 1) / GAL_RT;
    int r = blockIdx.x / tiles;
    const int j0 = (blockIdx.x - r * tiles) * GAL_RT;
    if (r >= self_rank) r++;                           // the G - 1 foreign ranks
    const float *peer = reinterpret_cast<const float *>(peer_ptrs[r]);
    const int *peer_ids = reinterpret_cast<const int *>(peer) + (size_t)t_max * D;
    const float *local = reinterpret_cast<const float *>(peer_ptrs[self_rank]);
    const int *local_ids = reinterpret_cast<const int *>(local) + (size_t)t_max * D;
    if (threadIdx.x < GAL_RT) s_ids[threadIdx.x] = j0 + threadIdx.x < t_max ? peer_ids[j0 + threadIdx.x] : -1;
    __syncthreads();
    bool any = false;
#pragma unroll
    for (int q = 0; q < GAL_RT; q++) any |= s_ids[q] >= 0;
    if (!any) return;                                  // block-uniform: nothing exported in this tile
    for (int e = threadIdx.x * 4; e < GAL_RT * D; e += blockDim.x * 4) {
        const int q = e / D;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s_ids[q] >= 0) v = *reinterpret_cast<const float4 *>(peer + (size_t)(j0 + q) * D + (e - q * D));     // NVLink
        *reinterpret_cast<float4 *>(s_tile + e) = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = warp; i < t_max; i += nw) {
        if (local_ids[i] < 0) continue;                // warp-uniform
        const float *f = local + (size_t)i * D;
        float4 a[4];                                   // D = 512: lane l owns elements 4l + 128m .. + 3
#pragma unroll
        for (int m = 0; m < 4; m++) a[m] = *reinterpret_cast<const float4 *>(f + lane * 4 + 128 * m);
#pragma unroll
        for (int q = 0; q < GAL_RT; q++) {
            if (s_ids[q] < 0) continue;
            const float *g = s_tile + q * D;
            float acc = 0.f;
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const float4 b = *reinterpret_cast<const float4 *>(g + lane * 4 + 128 * m);
                acc = fmaf(a[m].x, b.x, acc); acc = fmaf(a[m].y, b.y, acc);
                acc = fmaf(a[m].z, b.z, acc); acc = fmaf(a[m].w, b.w, acc);
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (lane == 0) atomicMin(best + i, gal_key(1.f - acc, r * t_max + j0 + q));
        }
    }
}

__global__ void gallery_peer_finalize_kernel(const unsigned long long *__restrict__ peer_ptrs, int self_rank, int t_max, int D,
                                             float max_dist, const unsigned long long *__restrict__ best,
                                             int *__restrict__ m_rank, int *__restrict__ m_id, float *__restrict__ m_dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t_max) return;
    const unsigned long long k = best[i];
    const bool found = k != ~0ull;
    const int j = (int)(unsigned)k;
    const float d = found ? gal_key_dist(k) : INFINITY;
    const bool hit = found && d <= max_dist;
    int id = -1;
    if (hit) id = (reinterpret_cast<const int *>(peer_ptrs[j / t_max]) + (size_t)t_max * D)[j % t_max];
    m_rank[i] = hit ? j / t_max : -1;
    m_id[i] = id;
    m_dist[i] = d;
    (void)self_rank;
}

}  // namespace

// peer_ptrs_dev: device array of n_ranks pointers, entry r = rank r's packed export [t_max * dim f32 | t_max int32 ids]
// (peer-accessible memory: symmetric-memory buffers over NVLink, or plain device pointers in a single-GPU test);
// best_scratch_dev: t_max uint64.  dim must be 512.
extern "C" int ssb_gallery_peer_match(const void *peer_ptrs_dev, int n_ranks, int self_rank, int t_max, int dim, float max_dist,
                                      void *best_scratch_dev, int32_t *match_rank_out_dev, int32_t *match_id_out_dev,
                                      float *match_dist_out_dev, ssb_stream_t stream) {
    if (!peer_ptrs_dev || !best_scratch_dev || !match_rank_out_dev || !match_id_out_dev || !match_dist_out_dev) { ssb_set_error("null argument"); return -1; }
    if (n_ranks < 1 || self_rank < 0 || self_rank >= n_ranks || t_max < 1 || dim != 512) {
        ssb_set_error("bad gallery geometry (dim must be 512)");
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long *pp = (const unsigned long long *)peer_ptrs_dev;
    unsigned long long *best = (unsigned long long *)best_scratch_dev;
    gallery_peer_init_kernel<<<(t_max + 127) / 128, 128, 0, st>>>(best, t_max);
    SSB_CHECK_LAUNCH();
    if (n_ranks > 1) {
        const int tiles = (t_max + GAL_RT - 1) / GAL_RT;
        gallery_peer_match_kernel<<<(n_ranks - 1) * tiles, 256, GAL_RT * dim * sizeof(float), st>>>(pp, n_ranks, self_rank, t_max, dim, best);
        SSB_CHECK_LAUNCH();
    }
    gallery_peer_finalize_kernel<<<(t_max + 127) / 128, 128, 0, st>>>(pp, self_rank, t_max, dim, max_dist, best, match_rank_out_dev,
                                                                     match_id_out_dev, match_dist_out_dev);
    SSB_CHECK_LAUNCH();
    return 0;
}

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
