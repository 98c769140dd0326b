// OSNet OSBlock, fused, on the 5th-gen tensor cores (tcgen05 / TMEM), sm_100a.
//
// One launch computes a whole OSBlock:
//   x1 = relu(conv1x1(x));  s_k = LightConv3x3^k(x1), k = 1..4;
//   y  = relu( conv1x1_lin( sum_k gate(s_k) * s_k ) + residual(x) )
// for a band of R image rows of one crop per CTA; the bands of a crop form a
// thread-block cluster so the ChannelGate's global average pool is reduced over
// distributed shared memory.  All intermediate maps stay in shared memory.
//
// Tensor-core formulation
//   * 1x1 convs are GEMMs  [pixels x Cin] x [Cin x Cout].
//   * LightConv3x3 (1x1 conv, then depthwise 3x3, no nonlinearity in between)
//     is algebraically ONE dense 3x3 conv with W'[c][tap][ci] = dw[tap][c] *
//     pw[ci][c]; it runs as 9 shifted GEMMs: the activation map is stored
//     [channel/8][pixel][8] (K-major, no swizzle, SBO = 128 B) with a one-pixel
//     zero ring, so tap (dy,dx) is the same operand with its descriptor start
//     address moved by (dy*(W+2)+dx)*16 bytes -- no im2col copy, and the
//     depthwise work (no tensor-core mapping on its own) rides the tensor pipe.
//   * Precision: every operand is an fp16 pair (hi, lo) with hi+lo == fp32 value
//     to ~2^-22; a product is 3 MMAs (hi*hi + lo*hi + hi*lo) accumulated in
//     fp32 TMEM.  Plain fp16 moved the embedding by 8.5e-3 rel (weights 8.4e-3,
//     activations 2.7e-3) -- outside the 1e-3 parity bar; the split gives 6e-6.
//   * The gate is a per-channel scale, so  W3 * sum_k g_k.s_k = sum_k (W3 diag(g_k)) s_k:
//     each stream is multiplied by its own gate-scaled copy of W3 and
//     accumulated in the same TMEM tile; no gated map is ever materialised.
//
// One thread issues all MMAs of a layer, one tcgen05.commit per layer signals an
// mbarrier, the 4 warps drain TMEM (bias, ReLU, zero-ring mask, hi/lo split) into
// the next layer's operand map.  Weights arrive by cp.async.bulk (TMA engine).
#include <cooperative_groups.h>

#include "ssb_common.cuh"
#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace {

template <int CIN_, int MID_, int MIDP_, int COUT_, int H_, int W_, int R_, int HALO_, int NB_,
          bool DOWN_, int NSTAGE_>
struct BlkCfg {
    static constexpr int CIN = CIN_, MID = MID_, MIDP = MIDP_, COUT = COUT_, H = H_, W = W_, R = R_;
    static constexpr int HALO = HALO_, NB = NB_, NSTAGE = NSTAGE_;
    static constexpr bool DOWN = DOWN_;
    static constexpr int RH = R + 2 * HALO;           // band rows kept (with halo)
    static constexpr int WP = W + 2;                  // padded row pitch
    static constexpr int NPX = (RH + 2) * WP;         // padded band pixels
    static constexpr int NT = (NPX + 127) / 128;      // M tiles
    static constexpr int GUARD = WP + 2;
    static constexpr int MAP_PX = GUARD + NT * 128 + GUARD;
    static constexpr int MCH = MIDP / 8;              // 16-byte K chunks per pixel
    static constexpr int PLANE_B = MAP_PX * 16;       // LBO of a map operand
    static constexpr int MAP_HALF_B = MCH * PLANE_B;
    static constexpr int MAP_B = 2 * MAP_HALF_B;      // hi planes then lo planes
    static constexpr int IN_P0 = (1 + HALO) * WP, IN_P1 = (1 + HALO + R) * WP;
    static constexpr int IT0 = IN_P0 / 128, IT1 = (IN_P1 + 127) / 128, NIT = IT1 - IT0;
    // "trapezoid": a LightConv with r more LightConvs after it in its stream only has to be
    // right on the rows within r of the band's own rows, so the M tiles wholly outside
    // [own - r, own + r] are skipped (no MMAs, no drain) -- the halo recompute shrinks as the
    // stream gets closer to its end.
    __host__ __device__ static constexpr int tile_lo(int r) {
        int row = 1 + HALO - r; if (row < 0) row = 0;
        return (row * WP) / 128;
    }
    __host__ __device__ static constexpr int tile_hi(int r) {
        int row = 1 + HALO + R + r; if (row > RH + 2) row = RH + 2;
        int t = (row * WP + 127) / 128; return t > NT ? NT : t;
    }
    // CAT: the 3x3 / conv1 MMAs are bound by the shared-memory read of the A tile (4 KB per
    // MMA, ~40 cycles, measured), so where TMEM has room the hi and lo weight rows are
    // concatenated along N:  D[:, 0:M] += Ah*Bh (+ Al*Bh),  D[:, M:2M] += Ah*Bl  -- two A
    // reads per product instead of three; the drain adds the two column halves.
    static constexpr bool CAT = (NT * 2 * MIDP + NIT * COUT) <= 512;
    static constexpr int LCN = CAT ? 2 * MIDP : MIDP;     // TMEM columns per 3x3 tile
    static constexpr int TM_LC = 0, TM_C3 = NT * LCN;
    static constexpr int TM_COLS = TM_C3 + NIT * COUT;
    static constexpr int LCW_HALF_B = 9 * MIDP * MIDP * 2, LCW_B = 2 * LCW_HALF_B;
    static constexpr int C1W_HALF_B = CIN * MIDP * 2, C1W_B = 2 * C1W_HALF_B;
    static constexpr int DNW_HALF_B = DOWN ? CIN * COUT * 2 : 0, DNW_B = 2 * DNW_HALF_B;
    static constexpr int C3W_HALF_B = MIDP * COUT * 2, C3W_B = 2 * C3W_HALF_B;
    static constexpr int W1_B = (C1W_B + DNW_B) > LCW_B ? (C1W_B + DNW_B) : LCW_B;
    static constexpr int STG_HALF_B = (CIN / 8) * 128 * 16, STG_B = 2 * STG_HALF_B;
    static constexpr int NPAR = MIDP + 10 * MIDP + COUT + MIDP * 2 + 2 + 2 * MIDP + MIDP;  // floats
    // shared-memory carve-up (bytes)
    static constexpr int OFF_X1 = 0;
    static constexpr int OFF_P = OFF_X1 + MAP_B;
    static constexpr int OFF_Q = OFF_P + MAP_B;
    static constexpr int OFF_W1 = OFF_Q + MAP_B;          // conv1+down weights, later LC weights
    static constexpr int OFF_C3 = OFF_W1 + W1_B;          // gate-scaled conv3 weights
    static constexpr int OFF_PAR = OFF_C3 + C3W_B;        // fp32 biases + gate params
    static constexpr int OFF_GAP = OFF_PAR + ((NPAR * 4 + 127) / 128) * 128;   // [4][MIDP] floats
    static constexpr int OFF_MISC = OFF_GAP + 4 * MIDP * 4;
    static constexpr int SMEM_B = OFF_MISC + 128 + 18 * 32 * 4 + 64;
    static_assert(NT <= 8, "per-tile barriers");
    static_assert(NSTAGE * STG_B <= 2 * MAP_B, "x staging must fit in the P+Q maps");
    static_assert(TM_COLS <= 512, "TMEM columns");
    static_assert(SMEM_B <= 232448, "shared memory");
    static_assert(MIDP % 16 == 0 && COUT % 16 == 0 && CIN % 16 == 0, "MMA shapes");
    static_assert(H % R == 0 && H / R == NB, "bands");
    static_assert(MID <= MIDP && COUT == 4 * MID, "OSBlock channel plan");
    // global blob sections (bytes): C1W | DNW | LCW[10] | PAR (fp32) | W3 (fp32 [MIDP][COUT])
    static constexpr int G_C1W = 0;
    static constexpr int G_LCW = C1W_B + DNW_B;
    static constexpr int G_PAR = G_LCW + 10 * LCW_B;
    static constexpr int G_W3 = G_PAR + ((NPAR * 4 + 127) / 128) * 128;
    static constexpr int G_TOTAL = G_W3 + MIDP * COUT * 4;
};

// parameter offsets inside PAR (floats)
template <class C> struct Par {
    static constexpr int B1 = 0;
    static constexpr int BLC = C::MIDP;
    static constexpr int B3 = BLC + 10 * C::MIDP;
    static constexpr int GW1 = B3 + C::COUT;       // [MIDP][2]
    static constexpr int GB1 = GW1 + 2 * C::MIDP;  // [2]
    static constexpr int GW2 = GB1 + 2;            // [2][MIDP]
    static constexpr int GB2 = GW2 + 2 * C::MIDP;  // [MIDP]
};

__device__ __forceinline__ void split_hl(float v, __half &h, __half &l) {
    h = __float2half_rn(v);
    l = __float2half_rn(v - __half2float(h));
}

// CTA shape of the OSBlock kernel: 16 warps.  Warp w may touch TMEM lanes
// 32*(w%4)..+31 only, so the 4 warps sharing a lane quadrant split the M tiles
// (group g = w/4 takes tiles g, g+4, ...).  More resident warps is what hides
// the ALU/LDTM latency of the epilogues: with 4 warps (one per scheduler) 45 %
// of the time was dependent-issue stall.
constexpr int OSB_THREADS = 512;
constexpr int OSB_GROUPS = OSB_THREADS / 128;

// two values at once: one packed convert each way (F2FP / HADD2.F32) instead of four scalar ones
__device__ __forceinline__ void split_hl2(float a, float b, __half2 &h, __half2 &l) {
    h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    l = __floats2half2_rn(a - hf.x, b - hf.y);
}

struct TrueT { static constexpr bool value = true; };
struct FalseT { static constexpr bool value = false; };

// The MMA-issuing thread is the critical path (ncu: tensor pipe 8 % busy while every
// other warp waits on the commit barrier), so descriptors are built ONCE per operand
// and advanced by adding the offset in 16-byte units to the low word -- the start
// address field is the low 14 bits and never carries out for valid shared addresses.
__device__ __forceinline__ uint64_t desc_adv(uint64_t base, int units16) {
    return base + (uint64_t)(int64_t)units16;
}
// one hi/lo product: D (+)= Ah*Bh + Al*Bh + Ah*Bl
__device__ __forceinline__ void mma3(uint32_t d, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl,
                                     uint32_t idesc, uint32_t acc) {
    tc::mma_f16_ss(d, ah, bh, idesc, acc);
    tc::mma_f16_ss(d, al, bh, idesc, 1);
    tc::mma_f16_ss(d, ah, bl, idesc, 1);
}

#ifdef SSB_BASELINES        // the 9-tap OSBlock kernel is an A/B baseline: libssb_dbg.so only
struct Pipe {           // one mbarrier, bulk-synchronous use: every thread waits every commit
    uint64_t *bar;
    uint32_t phase;
    bool ok;
    __device__ __forceinline__ void wait() {
        if (!tc::mbar_wait(bar, phase)) ok = false;
        phase ^= 1;
    }
};

template <class C>
__global__ void __launch_bounds__(OSB_THREADS, 1)
osblock_tc_kernel(const float *__restrict__ x, float *__restrict__ y,
                  const unsigned char *__restrict__ wblob, int n_crops, int *__restrict__ status,
                  long long *__restrict__ dbg) {
    extern __shared__ __align__(1024) unsigned char smem[];
    using P = Par<C>;
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, grp = warp >> 2;          // TMEM lane quadrant / tile group
    const int crop = blockIdx.x / C::NB, band = blockIdx.x % C::NB;
    const int row0 = band * C::R - C::HALO;            // image row of local row lr = 1

    unsigned char *sX1 = smem + C::OFF_X1, *sP = smem + C::OFF_P, *sQ = smem + C::OFF_Q;
    unsigned char *sW1 = smem + C::OFF_W1, *sC3 = smem + C::OFF_C3;
    float *sPar = reinterpret_cast<float *>(smem + C::OFF_PAR);
    float *sGap = reinterpret_cast<float *>(smem + C::OFF_GAP);      // [4][MIDP] partial sums
    uint64_t *bar_mma = reinterpret_cast<uint64_t *>(smem + C::OFF_MISC);
    uint64_t *bar_w = bar_mma + 1;
    uint64_t *bar_stg = bar_mma + 2;                   // [2] one per x-staging buffer
    uint64_t *bar_tile = bar_mma + 4;                  // [8] per M tile of the current 3x3 layer
    uint64_t *bar_c3 = bar_mma + 12;                   // conv3 accumulation of the current stream
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bar_mma + 13);
    float *s_scr = reinterpret_cast<float *>(smem + C::OFF_MISC + 128);   // [16][MIDP] warp partials
    float *s_mean = s_scr + 16 * C::MIDP;                                 // [MIDP]
    float *s_gate = s_mean + C::MIDP;                                     // [MIDP]

    if (warp == 0) tc::tmem_alloc(s_tmem, 512);
    if (tid == 0) {
        tc::mbar_init(bar_mma, 1);
        tc::mbar_init(bar_w, 1);
        tc::mbar_init(bar_stg, 1);
        tc::mbar_init(bar_stg + 1, 1);
        for (int i = 0; i < 8; i++) tc::mbar_init(bar_tile + i, 1);
        tc::mbar_init(bar_c3, C::NIT < 4 ? C::NIT : 4);
        tc::fence_mbar_init();
    }
    for (int i = tid; i < C::NPAR; i += OSB_THREADS)
        sPar[i] = reinterpret_cast<const float *>(wblob + C::G_PAR)[i];
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *s_tmem;
    // optional phase timestamps of CTA 0 (tools/time_stages.py); dbg == nullptr in production
    int dbg_n = 0;
    auto stamp = [&]() { if (dbg && blockIdx.x == 0 && tid == 0 && dbg_n < 63) dbg[1 + dbg_n++] = clock64(); };
    stamp();
    Pipe mma{bar_mma, 0, true};
    uint32_t w_phase = 0;
    bool ok = true;

    // phase-1 weights (conv1 [+ downsample]) by bulk copy
    if (tid == 0) {
        tc::mbar_arrive_expect_tx(bar_w, C::C1W_B + C::DNW_B);
        tc::bulk_g2s(sW1, wblob + C::G_C1W, C::C1W_B + C::DNW_B, bar_w);
    }

    auto pixel_valid = [&](int p, int &gr, int &gc) -> bool {
        if (p >= C::NPX) return false;
        const int lr = p / C::WP, lc = p - lr * C::WP;
        gr = row0 + lr - 1;
        gc = lc - 1;
        return lc >= 1 && lc <= C::W && lr >= 1 && lr <= C::RH && gr >= 0 && gr < C::H;
    };

    // this thread drains tile rows p = t*128 + quad*32 + lane for t = grp, grp+4: the same
    // pixels in every layer, so their validity is decided once (bit j: j-th tile of the thread)
    unsigned valid_m = 0, own_m = 0;
    {
        int j = 0;
        for (int t = grp; t < C::NT; t += OSB_GROUPS, j++) {
            const int p = t * 128 + quad * 32 + lane;
            int gr, gc;
            const bool v = pixel_valid(p, gr, gc);
            const int lr = p / C::WP;
            if (v) valid_m |= 1u << j;
            if (v && lr >= 1 + C::HALO && lr < 1 + C::HALO + C::R) own_m |= 1u << j;
        }
    }

    // ------------------------------------------------------------------
    // phase 1: X1 = relu(conv1(x)) on every band tile; downsample on inner tiles
    // ------------------------------------------------------------------
    constexpr uint32_t IDESC_MID = tc::make_idesc_f16(128, C::MIDP);
    constexpr uint32_t IDESC_CAT = tc::make_idesc_f16(128, 2 * C::MIDP);
    constexpr uint32_t IDESC_OUT = tc::make_idesc_f16(128, C::COUT);
    const float *xin = x + (size_t)crop * C::H * C::W * C::CIN;
    // a waiter may lag an mbarrier by at most one phase, so every staging buffer has
    // its own barrier: tile t commits to bar_stg[t % NSTAGE] and is waited before reuse
    uint32_t stg_phase[2] = {0, 0};
    // software pipeline over the x tiles: the global loads of tile t+1 are issued before the
    // fence / barrier / MMA issue of tile t, so their latency is off the critical path
    constexpr int F4 = C::CIN / 4;
    constexpr int PER = (128 * F4) / OSB_THREADS;            // float4 items per thread per tile
    static_assert((128 * F4) % OSB_THREADS == 0, "tile items divide the CTA");
    float4 xr[PER];
    auto load_tile = [&](int t) {
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int idx = tid + q * OSB_THREADS;
            const int px = idx / F4, f4 = idx - px * F4;
            int gr, gc;
            xr[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pixel_valid(t * 128 + px, gr, gc))
                xr[q] = *reinterpret_cast<const float4 *>(xin + ((size_t)gr * C::W + gc) * C::CIN + f4 * 4);
        }
    };
    load_tile(0);
    for (int t = 0; t < C::NT; t++) {
        const int sb = t % C::NSTAGE;
        unsigned char *stg = sP + sb * C::STG_B;
        if (t >= C::NSTAGE) {                                // MMAs that read this buffer are done
            if (!tc::mbar_wait(bar_stg + sb, stg_phase[sb])) ok = false;
            stg_phase[sb] ^= 1;
        }
        // stage tile t of x: [CIN/8][128][8] hi, then lo
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int idx = tid + q * OSB_THREADS;
            const int px = idx / F4, f4 = idx - px * F4;
            const float4 v = xr[q];
            __align__(8) __half2 h[2], l[2];
            split_hl2(v.x, v.y, h[0], l[0]);
            split_hl2(v.z, v.w, h[1], l[1]);
            const int off = (f4 >> 1) * 2048 + px * 16 + (f4 & 1) * 8;
            *reinterpret_cast<uint2 *>(stg + off) = *reinterpret_cast<uint2 *>(h);
            *reinterpret_cast<uint2 *>(stg + C::STG_HALF_B + off) = *reinterpret_cast<uint2 *>(l);
        }
        if (t + 1 < C::NT) load_tile(t + 1);
        tc::fence_async_smem();
        if (t == 0) { if (!tc::mbar_wait(bar_w, w_phase)) ok = false; w_phase ^= 1; }
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        if (tid == 0) {
            const uint64_t ah0 = tc::make_smem_desc(tc::smem_u32(stg), 2048, 128);
            const uint64_t al0 = desc_adv(ah0, C::STG_HALF_B / 16);
            const uint32_t d1 = tmem + C::TM_LC + t * C::LCN;
            if (C::CAT) {
                const uint64_t bc0 = tc::make_smem_desc(tc::smem_u32(sW1), 2 * C::MIDP * 16, 128);
#pragma unroll
                for (int ks = 0; ks < C::CIN / 16; ks++) {
                    tc::mma_f16_ss(d1, desc_adv(ah0, ks * 256), desc_adv(bc0, ks * 4 * C::MIDP), IDESC_CAT, ks > 0);
                    tc::mma_f16_ss(d1, desc_adv(al0, ks * 256), desc_adv(bc0, ks * 4 * C::MIDP), IDESC_MID, 1);
                }
            } else {
                const uint64_t bh0 = tc::make_smem_desc(tc::smem_u32(sW1), C::MIDP * 16, 128);
                const uint64_t bl0 = desc_adv(bh0, C::C1W_HALF_B / 16);
#pragma unroll
                for (int ks = 0; ks < C::CIN / 16; ks++)
                    mma3(d1, desc_adv(ah0, ks * 256), desc_adv(al0, ks * 256), desc_adv(bh0, ks * 2 * C::MIDP),
                         desc_adv(bl0, ks * 2 * C::MIDP), IDESC_MID, ks > 0);
            }
            if (C::DOWN && t >= C::IT0 && t < C::IT1) {
                const uint64_t dh0 = tc::make_smem_desc(tc::smem_u32(sW1) + C::C1W_B, C::COUT * 16, 128);
                const uint64_t dl0 = desc_adv(dh0, C::DNW_HALF_B / 16);
                const uint32_t d2 = tmem + C::TM_C3 + (t - C::IT0) * C::COUT;
#pragma unroll
                for (int ks = 0; ks < C::CIN / 16; ks++)
                    mma3(d2, desc_adv(ah0, ks * 256), desc_adv(al0, ks * 256), desc_adv(dh0, ks * 2 * C::COUT),
                         desc_adv(dl0, ks * 2 * C::COUT), IDESC_OUT, ks > 0);
            }
            tc::mma_commit(bar_stg + sb);
        }
    }
    for (int sb = 0; sb < C::NSTAGE && sb < C::NT; sb++) {    // the last commit of every buffer
        if (!tc::mbar_wait(bar_stg + sb, stg_phase[sb])) ok = false;
        stg_phase[sb] ^= 1;
    }
    tc::fence_after_sync();
    // W1 region is free: fetch the first LightConv's weights while the conv1 epilogue runs
    if (tid == 0) {
        tc::mbar_arrive_expect_tx(bar_w, C::LCW_B);
        tc::bulk_g2s(sW1, wblob + C::G_LCW, C::LCW_B, bar_w);
    }

    // epilogue: TMEM tile -> (+bias, relu, zero-ring mask) -> hi/lo operand map.
    // GAPACC: also accumulate per-channel sums over the band's own inner pixels.
    uint32_t tile_par_bits = 0;       // bit t: parity of bar_tile[t] (a barrier only flips in layers that use it)
    auto tile_par_of = [&](int t) -> uint32_t { return (tile_par_bits >> t) & 1u; };
    auto tile_used_advance = [&](int lo, int hi) { for (int t = lo; t < hi; t++) tile_par_bits ^= 1u << t; };
    // tile_bar != nullptr: tile t may be drained as soon as ITS MMAs have completed
    // (tile_bar[t], parity tile_par), while later tiles are still on the tensor pipe.
    auto drain_to_map = [&](unsigned char *dst, const float *bias, auto gapacc, float *gap,
                            uint64_t *tile_bar, uint32_t tile_par, int t_lo, int t_hi) {
        constexpr bool GAPACC = decltype(gapacc)::value;
        int jt = 0;
        for (int t = grp; t < C::NT; t += OSB_GROUPS, jt++) {
            if (t < t_lo || t >= t_hi) continue;
            if (tile_bar) {
                if (!tc::mbar_wait(tile_bar + t, tile_par_of(t))) ok = false;
                tc::fence_after_sync();
            }
            (void)tile_par;
            const int p = t * 128 + quad * 32 + lane;
            const bool valid = (valid_m >> jt) & 1u;
            const bool own = (own_m >> jt) & 1u;
            unsigned char *d_hi = dst + (C::GUARD + p) * 16, *d_lo = d_hi + C::MAP_HALF_B;
            float v[C::MIDP];
            tc::tmem_ldN<C::MIDP>(tmem + ((uint32_t)(quad * 32) << 16) + C::TM_LC + t * C::LCN, v);
            if (C::CAT) {
                float w[C::MIDP];
                tc::tmem_ldN<C::MIDP>(tmem + ((uint32_t)(quad * 32) << 16) + C::TM_LC + t * C::LCN + C::MIDP, w);
#pragma unroll
                for (int j = 0; j < C::MIDP; j++) v[j] += w[j];
            }
#pragma unroll
            for (int c0 = 0; c0 < C::MIDP; c0 += 16) {
                __align__(16) __half2 h[8];
                __align__(16) __half2 l[8];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    const float f0 = valid ? fmaxf(v[c0 + j] + bias[c0 + j], 0.f) : 0.f;
                    const float f1 = valid ? fmaxf(v[c0 + j + 1] + bias[c0 + j + 1], 0.f) : 0.f;
                    if (GAPACC) { if (own) { gap[c0 + j] += f0; gap[c0 + j + 1] += f1; } }
                    split_hl2(f0, f1, h[j >> 1], l[j >> 1]);
                }
                const int pl = (c0 >> 3) * C::PLANE_B;
                *reinterpret_cast<uint4 *>(d_hi + pl) = *reinterpret_cast<uint4 *>(&h[0]);
                *reinterpret_cast<uint4 *>(d_hi + pl + C::PLANE_B) = *reinterpret_cast<uint4 *>(&h[4]);
                *reinterpret_cast<uint4 *>(d_lo + pl) = *reinterpret_cast<uint4 *>(&l[0]);
                *reinterpret_cast<uint4 *>(d_lo + pl + C::PLANE_B) = *reinterpret_cast<uint4 *>(&l[4]);
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    };

    stamp();                                   // [1] phase 1 (staging + conv1/down MMAs) done
    drain_to_map(sX1, sPar + P::B1, FalseT{}, nullptr, nullptr, 0, 0, C::NT);
    stamp();                                   // [2] X1 drained

    // ------------------------------------------------------------------
    // phase 2: four streams of dense 3x3 convs + gated conv3 accumulation
    // ------------------------------------------------------------------
    const float *w3 = reinterpret_cast<const float *>(wblob + C::G_W3);      // [MIDP][COUT]
    int lc = 0;
    uint32_t c3_par = 0;
    bool c3_pending = false;
    for (int s = 0; s < 4; s++) {
        const unsigned char *src = sX1;
        unsigned char *dst = sP;
        for (int k = 0; k <= s; k++, lc++) {
            const int rem = s - k;                                // LightConvs after this one in the stream
            const int t_lo = rem == 0 ? C::tile_lo(0) : rem == 1 ? C::tile_lo(1) : rem == 2 ? C::tile_lo(2) : C::tile_lo(3);
            const int t_hi = rem == 0 ? C::tile_hi(0) : rem == 1 ? C::tile_hi(1) : rem == 2 ? C::tile_hi(2) : C::tile_hi(3);
            if (!tc::mbar_wait(bar_w, w_phase)) ok = false;       // this LightConv's weights landed
            w_phase ^= 1;
            tc::fence_after_sync();
            stamp();                           // LC start (weights ready)
            // four issuing threads (one per SM sub-partition), tile t -> warp t % 4; every
            // tile commits to its own barrier so its drain overlaps the later tiles' MMAs
            if (warp < 4 && lane == 0) {
                const uint64_t ah0 = tc::make_smem_desc(tc::smem_u32(src) + C::GUARD * 16, C::PLANE_B, 128);
                const uint64_t al0 = desc_adv(ah0, C::MAP_HALF_B / 16);
                const uint64_t bh0 = tc::make_smem_desc(tc::smem_u32(sW1), (C::CAT ? 2 : 1) * C::MIDP * 16, 128);
                const uint64_t bl0 = desc_adv(bh0, C::LCW_HALF_B / 16);       // (unused with CAT)
#pragma unroll 1
                for (int t = t_lo + warp; t < t_hi; t += 4) {
                    const uint32_t d = tmem + C::TM_LC + t * C::LCN;
                    const uint64_t aht = desc_adv(ah0, t * 128), alt = desc_adv(al0, t * 128);
#pragma unroll
                    for (int tap = 0; tap < 9; tap++) {
                        const int po = (tap / 3 - 1) * C::WP + (tap % 3 - 1);        // pixels == 16-byte units
#pragma unroll
                        for (int ks = 0; ks < C::MIDP / 16; ks++) {
                            const int ka = ks * 2 * C::MAP_PX;
                            if (C::CAT) {
                                const int kb = (tap * C::MCH + ks * 2) * 2 * C::MIDP;
                                tc::mma_f16_ss(d, desc_adv(aht, po + ka), desc_adv(bh0, kb), IDESC_CAT, (tap | ks) != 0);
                                tc::mma_f16_ss(d, desc_adv(alt, po + ka), desc_adv(bh0, kb), IDESC_MID, 1);
                            } else {
                                const int kb = (tap * C::MCH + ks * 2) * C::MIDP;
                                mma3(d, desc_adv(aht, po + ka), desc_adv(alt, po + ka), desc_adv(bh0, kb),
                                     desc_adv(bl0, kb), IDESC_MID, (tap | ks) != 0);
                            }
                        }
                    }
                    tc::mma_commit(bar_tile + t);
                }
            }
            // one otherwise lightly loaded thread waits for the whole layer and then refills
            // the weight buffer with the next LightConv's weights
            if (warp == 15 && lane == 0) {
                for (int t = t_lo; t < t_hi; t++)
                    if (!tc::mbar_wait(bar_tile + t, tile_par_of(t))) ok = false;
                if (lc + 1 < 10) {
                    tc::mbar_arrive_expect_tx(bar_w, C::LCW_B);
                    tc::bulk_g2s(sW1, wblob + C::G_LCW + (size_t)(lc + 1) * C::LCW_B, C::LCW_B, bar_w);
                }
            }
            // the previous stream's conv3 MMAs read P/Q: they must be done before this
            // layer's drain overwrites those maps
            if (c3_pending) {
                if (!tc::mbar_wait(bar_c3, c3_par)) ok = false;
                c3_par ^= 1;
                c3_pending = false;
            }
            const bool last = (k == s);
            stamp();                           // LC issued (thread 0: its own tiles)
            if (!last) {
                drain_to_map(dst, sPar + P::BLC + lc * C::MIDP, FalseT{}, nullptr, bar_tile, 0, t_lo, t_hi);
                tile_used_advance(t_lo, t_hi);
                stamp();                       // LC drained
            } else {
                float gap[C::MIDP];
#pragma unroll
                for (int j = 0; j < C::MIDP; j++) gap[j] = 0.f;
                drain_to_map(dst, sPar + P::BLC + lc * C::MIDP, TrueT{}, gap, bar_tile, 0, t_lo, t_hi);
                tile_used_advance(t_lo, t_hi);
                stamp();                       // last LC of the stream drained
                // ---- ChannelGate: band-partial sums -> cluster -> mean -> MLP -> sigmoid
                // transpose-reduce over the warp: 31 (MIDP = 32) / 16 (MIDP = 16) shuffles in
                // total instead of 5 per channel; lane c (c >> 1 for MIDP = 16) ends with channel c
#pragma unroll
                for (int off = 16, nh = C::MIDP / 2; nh >= 1; off >>= 1, nh >>= 1) {
                    const bool up = (lane & off) != 0;
#pragma unroll
                    for (int j = 0; j < nh; j++) {
                        const float send = up ? gap[j] : gap[j + nh];
                        const float recv = __shfl_xor_sync(0xffffffffu, send, off);
                        gap[j] = (up ? gap[j + nh] : gap[j]) + recv;
                    }
                }
                if (C::MIDP == 16) {
                    gap[0] += __shfl_xor_sync(0xffffffffu, gap[0], 1);
                    if (!(lane & 1)) s_scr[warp * C::MIDP + (lane >> 1)] = gap[0];
                } else {
                    s_scr[warp * C::MIDP + lane] = gap[0];
                }
                __syncthreads();
                if (tid < C::MIDP) {
                    float tot = 0.f;
#pragma unroll
                    for (int w = 0; w < OSB_THREADS / 32; w++) tot += s_scr[w * C::MIDP + tid];
                    sGap[s * C::MIDP + tid] = tot;
                }
                if (C::NB > 1) cluster.sync(); else __syncthreads();
                if (tid < C::MIDP) {          // fixed band order: every CTA of the crop gets the same bits
                    float tot = 0.f;
                    for (int b = 0; b < C::NB; b++) {
                        const float *rg = (C::NB > 1) ? cluster.map_shared_rank(sGap, b) : sGap;
                        tot += rg[s * C::MIDP + tid];
                    }
                    s_mean[tid] = tot / (float)(C::H * C::W);
                }
                __syncthreads();
                if (tid < C::MIDP) {
                    float h0 = sPar[P::GB1 + 0], h1 = sPar[P::GB1 + 1];
                    for (int q = 0; q < C::MIDP; q++) {
                        const float m = s_mean[q];
                        h0 = fmaf(m, sPar[P::GW1 + q * 2 + 0], h0);
                        h1 = fmaf(m, sPar[P::GW1 + q * 2 + 1], h1);
                    }
                    h0 = fmaxf(h0, 0.f);
                    h1 = fmaxf(h1, 0.f);
                    float g = sPar[P::GB2 + tid];
                    g = fmaf(h0, sPar[P::GW2 + tid], g);
                    g = fmaf(h1, sPar[P::GW2 + C::MIDP + tid], g);
                    s_gate[tid] = 1.f / (1.f + expf(-g));
                }
                __syncthreads();
                // ---- gate-scaled conv3 weights  B[kc][co][8] = W3[k][co] * g[k]  (hi / lo)
                for (int u = tid; u < C::COUT * C::MCH; u += OSB_THREADS) {
                    const int kc = u / C::COUT, co = u - kc * C::COUT;
                    __align__(16) __half h[8];
                    __align__(16) __half l[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int kk = kc * 8 + j;
                        split_hl(w3[kk * C::COUT + co] * s_gate[kk], h[j], l[j]);
                    }
                    *reinterpret_cast<uint4 *>(sC3 + (size_t)u * 16) = *reinterpret_cast<uint4 *>(h);
                    *reinterpret_cast<uint4 *>(sC3 + C::C3W_HALF_B + (size_t)u * 16) = *reinterpret_cast<uint4 *>(l);
                }
                tc::fence_async_smem();
                tc::fence_before_sync();
                __syncthreads();
                tc::fence_after_sync();
                if (warp < 4 && warp < C::NIT && lane == 0) {
                    const uint64_t ah0 = tc::make_smem_desc(tc::smem_u32(dst) + C::GUARD * 16, C::PLANE_B, 128);
                    const uint64_t al0 = desc_adv(ah0, C::MAP_HALF_B / 16);
                    const uint64_t bh0 = tc::make_smem_desc(tc::smem_u32(sC3), C::COUT * 16, 128);
                    const uint64_t bl0 = desc_adv(bh0, C::C3W_HALF_B / 16);
                    const uint32_t acc0 = (C::DOWN || s > 0) ? 1u : 0u;
#pragma unroll 1
                    for (int i = warp; i < C::NIT; i += 4) {
                        const uint32_t d = tmem + C::TM_C3 + i * C::COUT;
#pragma unroll
                        for (int ks = 0; ks < C::MIDP / 16; ks++) {
                            const int ka = (C::IT0 + i) * 128 + ks * 2 * C::MAP_PX, kb = ks * 2 * C::COUT;
                            mma3(d, desc_adv(ah0, ka), desc_adv(al0, ka), desc_adv(bh0, kb), desc_adv(bl0, kb),
                                 IDESC_OUT, ks > 0 ? 1u : acc0);
                        }
                    }
                    tc::mma_commit(bar_c3);
                }
                c3_pending = true;
                stamp();                       // gate + conv3 issued
            }
            src = dst;
            dst = (dst == sP) ? sQ : sP;
        }
    }
    if (c3_pending) {                 // the last stream's conv3 MMAs
        if (!tc::mbar_wait(bar_c3, c3_par)) ok = false;
        c3_par ^= 1;
    }
    tc::fence_after_sync();

    // ------------------------------------------------------------------
    // final epilogue: y = relu(conv3 + bias (+ downsample already in TMEM) (+ x))
    // ------------------------------------------------------------------
    // Every thread owns one pixel row of 32 channels (128 B); written straight from the
    // registers that is 32 different cache lines per store instruction (the epilogue was
    // LSU-wavefront bound: 3.8 K cycles per unit).  Rows are therefore transposed through a
    // warp-private staging tile in the (now dead) map area so that 8 lanes cover one row:
    // 4 full lines per instruction, for the residual loads as well as for the stores.
    float *yout = y + (size_t)crop * C::H * C::W * C::COUT;
    constexpr int CCH = C::COUT / 32;                         // 32-column chunks per tile
    float *stage = reinterpret_cast<float *>(sX1) + warp * (32 * 36);
    static_assert(16 * 32 * 36 * 4 <= 3 * C::MAP_B, "staging tiles fit in the dead maps");
    for (int u = grp; u < C::NIT * CCH; u += OSB_GROUPS) {    // (tile, chunk) units over the 4 groups
        const int i = u / CCH, c0 = (u - i * CCH) * 32;
        const int p = (C::IT0 + i) * 128 + quad * 32 + lane;
        int gr, gc;
        const bool valid = pixel_valid(p, gr, gc);
        const int lr = p / C::WP;
        const bool own = valid && lr >= 1 + C::HALO && lr < 1 + C::HALO + C::R;
        const int rowoff = own ? gr * C::W + gc : -1;
        if (!C::DOWN) {                                       // identity residual, coalesced
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int row = it * 4 + (lane >> 3);
                const int ro = __shfl_sync(0xffffffffu, rowoff, row);
                float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ro >= 0) x4 = *reinterpret_cast<const float4 *>(xin + (size_t)ro * C::CIN + c0 + (lane & 7) * 4);
                *reinterpret_cast<float4 *>(stage + row * 36 + (lane & 7) * 4) = x4;
            }
        }
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + C::TM_C3 + i * C::COUT + c0, v);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 bb = *reinterpret_cast<const float4 *>(sPar + P::B3 + c0 + j);
            float4 r = make_float4(v[j] + bb.x, v[j + 1] + bb.y, v[j + 2] + bb.z, v[j + 3] + bb.w);
            if (!C::DOWN) {
                const float4 xv = *reinterpret_cast<const float4 *>(stage + lane * 36 + j);
                r.x += xv.x; r.y += xv.y; r.z += xv.z; r.w += xv.w;
            }
            r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f);
            r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
            *reinterpret_cast<float4 *>(stage + lane * 36 + j) = r;      // own row only: no hazard
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int row = it * 4 + (lane >> 3);
            const int ro = __shfl_sync(0xffffffffu, rowoff, row);
            if (ro >= 0)
                *reinterpret_cast<float4 *>(yout + (size_t)ro * C::COUT + c0 + (lane & 7) * 4) =
                    *reinterpret_cast<const float4 *>(stage + row * 36 + (lane & 7) * 4);
        }
        __syncwarp();
    }
    stamp();                                   // final epilogue done
    if (dbg && blockIdx.x == 0 && tid == 0) dbg[0] = dbg_n;
    if (!ok || !mma.ok) { if (tid == 0) atomicExch(status, 1); }
    tc::fence_before_sync();
    if (C::NB > 1) cluster.sync(); else __syncthreads();     // remote sGap reads are done
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
    (void)n_crops;
}

// ---------------------------------------------------------------------------
// the six OSBlocks of osnet_x0_25 (stage 2: 64x32, stage 3: 32x16, stage 4: 16x8)
// ---------------------------------------------------------------------------
//               CIN MID MIDP COUT  H   W   R HALO NB DOWN NSTAGE
#endif  // SSB_BASELINES (the shapes below also size the weight blob sections of the product build)
using Blk0 = BlkCfg<16, 16, 16, 64, 64, 32, 16, 4, 4, true, 2>;
using Blk1 = BlkCfg<64, 16, 16, 64, 64, 32, 16, 4, 4, false, 2>;
using Blk2 = BlkCfg<64, 24, 32, 96, 32, 16, 8, 4, 4, true, 2>;
using Blk3 = BlkCfg<96, 24, 32, 96, 32, 16, 8, 4, 4, false, 2>;
using Blk4 = BlkCfg<96, 32, 32, 128, 16, 8, 16, 0, 1, true, 1>;
using Blk5 = BlkCfg<128, 32, 32, 128, 16, 8, 16, 0, 1, false, 1>;
#ifdef SSB_BASELINES

template <class C>
int launch_block(const float *x, float *y, const unsigned char *w, int n, int *status, long long *dbg,
                 cudaStream_t st) {
    static const int key = ssb_new_key();
    if (ssb_first_on_device(key))
        SSB_CHECK_CUDA(cudaFuncSetAttribute(osblock_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_B));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n * C::NB);
    cfg.blockDim = dim3(OSB_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_B;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C::NB;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SSB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, osblock_tc_kernel<C>, x, y, w, n, status, dbg));
    g_ssb_launches++;
    return 0;
}
#endif  // SSB_BASELINES

// ---------------------------------------------------------------------------
// 1x1 conv + ReLU (+ 2x2 average pool) on the tensor cores: the two transition
// layers between the OSNet stages.  Persistent CTAs, 128-pixel M tiles, x tile
// and TMEM accumulator double-buffered so the MMA of tile i+1 overlaps the
// epilogue of tile i.  With POOL the 128 tile rows are 32 output pixels x their
// 4 window positions, so the pool is a 2-step shuffle over adjacent TMEM lanes.
// ---------------------------------------------------------------------------
template <int CIN_, int COUT_, int H_, int W_, bool POOL_>
struct PwCfg {
    static constexpr int CIN = CIN_, COUT = COUT_, H = H_, W = W_;
    static constexpr bool POOL = POOL_;
    static constexpr int STG_HALF_B = (CIN / 8) * 128 * 16, STG_B = 2 * STG_HALF_B;
    static constexpr int W_HALF_B = CIN * COUT * 2, W_B = 2 * W_HALF_B;
    static constexpr int OFF_STG = 0;
    static constexpr int OFF_W = 2 * STG_B;
    static constexpr int OFF_BIAS = OFF_W + W_B;
    static constexpr int OFF_MISC = OFF_BIAS + COUT * 4;
    static constexpr int SMEM_B = OFF_MISC + 128;
    static constexpr int G_W = 0, G_BIAS = W_B, G_TOTAL = W_B + ((COUT * 4 + 127) / 128) * 128;
    static_assert(SMEM_B <= 232448, "shared memory");
    static_assert(2 * COUT <= 512, "TMEM columns");
};

// PLANES: input and output are hi/lo fp16 operand planes [crop][hl][C/8][H*W][8] (reid_tc4.cu) instead of
// float32 NHWC: staging is a plain 16-byte copy (no split), the epilogue writes operand rows directly.
template <class C, bool PLANES>
__global__ void __launch_bounds__(OSB_THREADS, 1)
pw_tc_kernel(const float *__restrict__ x, float *__restrict__ y,
             const unsigned char *__restrict__ wblob, int n_crops, int *__restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, grp = warp >> 2;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    unsigned char *sStg = smem + C::OFF_STG, *sW = smem + C::OFF_W;
    float *sBias = reinterpret_cast<float *>(smem + C::OFF_BIAS);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + C::OFF_MISC);     // [2] mma, [2]=weights
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bar + 3);
    constexpr int HO = C::POOL ? C::H / 2 : C::H, WO = C::POOL ? C::W / 2 : C::W;
    constexpr int OUT_PER_TILE = C::POOL ? 32 : 128;
    const long long total_out = (long long)n_crops * HO * WO;
    const int ntiles = (int)((total_out + OUT_PER_TILE - 1) / OUT_PER_TILE);

    constexpr int PW_TM = 2 * C::COUT <= 128 ? 128 : 256;     // two accumulator buffers
    if (warp == 0) tc::tmem_alloc(s_tmem, PW_TM);
    if (tid == 0) {
        tc::mbar_init(bar, 1); tc::mbar_init(bar + 1, 1); tc::mbar_init(bar + 2, 1);
        tc::fence_mbar_init();
    }
    for (int i = tid; i < C::COUT; i += OSB_THREADS) sBias[i] = reinterpret_cast<const float *>(wblob + C::G_BIAS)[i];
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *s_tmem;
    tc::pdl_launch_dependents();
    if (tid == 0) {
        tc::mbar_arrive_expect_tx(bar + 2, C::W_B);
        tc::bulk_g2s(sW, wblob + C::G_W, C::W_B, bar + 2);
    }
    tc::pdl_wait();                    // the producer of x has completed (programmatic dependent launch)
    bool ok = true;
    uint32_t ph[2] = {0, 0};
    constexpr uint32_t IDESC = tc::make_idesc_f16(128, C::COUT);

    auto epilogue = [&](int tile, int buf) {
        if (!tc::mbar_wait(bar + buf, ph[buf])) ok = false;
        ph[buf] ^= 1;
        tc::fence_after_sync();
        const int m = quad * 32 + lane;
        long long o = C::POOL ? (long long)tile * 32 + (m >> 2) : (long long)tile * 128 + m;
        const bool wr = (o < total_out) && (!C::POOL || (m & 3) == 0);
#pragma unroll 1
        for (int c0 = grp * 16; c0 < C::COUT; c0 += 16 * OSB_GROUPS) {
            float v[16];
            tc::tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + buf * C::COUT + c0, v);
#pragma unroll
            for (int j = 0; j < 16; j++) {
                float f = fmaxf(v[j] + sBias[c0 + j], 0.f);
                if (C::POOL) {
                    f += __shfl_xor_sync(0xffffffffu, f, 1);
                    f += __shfl_xor_sync(0xffffffffu, f, 2);
                    f *= 0.25f;
                }
                v[j] = f;
            }
            if (wr && !PLANES) {
                float *dst = y + o * C::COUT + c0;
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            if (wr && PLANES) {
                const int n = (int)(o / (HO * WO)), pix = (int)(o - (long long)n * HO * WO);
                unsigned char *yb = reinterpret_cast<unsigned char *>(y) + (size_t)n * (4 * C::COUT * HO * WO);
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    __align__(16) __half2 h[4];
                    __align__(16) __half2 l[4];
#pragma unroll
                    for (int q = 0; q < 8; q += 2) split_hl2(v[j * 8 + q], v[j * 8 + q + 1], h[q >> 1], l[q >> 1]);
                    unsigned char *dst = yb + (size_t)(c0 / 8 + j) * (HO * WO) * 16 + (size_t)pix * 16;
                    *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<uint4 *>(h);
                    *reinterpret_cast<uint4 *>(dst + 2 * C::COUT * HO * WO) = *reinterpret_cast<uint4 *>(l);
                }
            }
        }
        tc::fence_before_sync();
    };

    int it = 0, prev_tile = -1;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        const int buf = it & 1;
        unsigned char *stg = sStg + buf * C::STG_B;
        constexpr int F4 = C::CIN / 4;
        if (PLANES) {
            // one 16-byte operand row per item: (hl, chunk) plane x tile row m; the row's pixel is the same
            // for all planes, so consecutive m of a warp read 32-byte window pairs of one plane
            constexpr int XCH = C::CIN / 8;
            const unsigned char *xb = reinterpret_cast<const unsigned char *>(x);
#pragma unroll 4
            for (int idx = tid; idx < 128 * 2 * XCH; idx += OSB_THREADS) {
                const int m = idx & 127, c2 = idx >> 7;
                const int hl = c2 / XCH, ch = c2 - hl * XCH;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                long long o;
                int ipx;
                if (C::POOL) {
                    o = (long long)tile * 32 + (m >> 2);
                    const int q = m & 3;
                    const int n = (int)(o / (HO * WO)), rem = (int)(o - (long long)n * HO * WO);
                    const int oy = rem / WO, ox = rem - oy * WO;
                    ipx = (2 * oy + (q >> 1)) * C::W + 2 * ox + (q & 1);
                    if (o < total_out)
                        v = *reinterpret_cast<const uint4 *>(xb + (size_t)n * (4 * C::CIN * C::H * C::W) + (size_t)hl * (2 * C::CIN * C::H * C::W) +
                                                             (size_t)ch * (C::H * C::W) * 16 + (size_t)ipx * 16);
                } else {
                    o = (long long)tile * 128 + m;
                    const int n = (int)(o / (C::H * C::W));
                    ipx = (int)(o - (long long)n * C::H * C::W);
                    if (o < total_out)
                        v = *reinterpret_cast<const uint4 *>(xb + (size_t)n * (4 * C::CIN * C::H * C::W) + (size_t)hl * (2 * C::CIN * C::H * C::W) +
                                                             (size_t)ch * (C::H * C::W) * 16 + (size_t)ipx * 16);
                }
                *reinterpret_cast<uint4 *>(stg + hl * C::STG_HALF_B + ch * 2048 + m * 16) = v;
            }
        }
#pragma unroll 2
        for (int idx = tid; idx < (PLANES ? 0 : 128 * F4); idx += OSB_THREADS) {
            const int m = idx / F4, f4 = idx - m * F4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (C::POOL) {
                const long long o = (long long)tile * 32 + (m >> 2);
                if (o < total_out) {
                    const int q = m & 3;
                    const int n = (int)(o / (HO * WO)), rem = (int)(o - (long long)n * HO * WO);
                    const int oy = rem / WO, ox = rem - oy * WO;
                    const int iy = 2 * oy + (q >> 1), ix = 2 * ox + (q & 1);
                    v = *reinterpret_cast<const float4 *>(x + (((size_t)n * C::H + iy) * C::W + ix) * C::CIN + f4 * 4);
                }
            } else {
                const long long o = (long long)tile * 128 + m;
                if (o < total_out) v = *reinterpret_cast<const float4 *>(x + (size_t)o * C::CIN + f4 * 4);
            }
            __align__(8) __half2 h[2], l[2];
            split_hl2(v.x, v.y, h[0], l[0]);
            split_hl2(v.z, v.w, h[1], l[1]);
            const int off = (f4 >> 1) * 2048 + m * 16 + (f4 & 1) * 8;
            *reinterpret_cast<uint2 *>(stg + off) = *reinterpret_cast<uint2 *>(h);
            *reinterpret_cast<uint2 *>(stg + C::STG_HALF_B + off) = *reinterpret_cast<uint2 *>(l);
        }
        tc::fence_async_smem();
        if (it == 0) { if (!tc::mbar_wait(bar + 2, 0)) ok = false; }
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        if (warp_u == 0 && tc::elect_one()) {      // warp-uniform issue: ~3x cheaper than `if (tid == 0)` (tc_common.cuh)
            const uint32_t a_hi = tc::smem_u32(stg), a_lo = a_hi + C::STG_HALF_B;
            const uint32_t b_hi = tc::smem_u32(sW), b_lo = b_hi + C::W_HALF_B;
            constexpr uint32_t LBO_B = C::COUT * 16;
            const uint32_t d = tmem + buf * C::COUT;
#pragma unroll
            for (int ks = 0; ks < C::CIN / 16; ks++) {
                const uint64_t ah = tc::make_smem_desc(a_hi + ks * 2 * 2048, 2048, 128);
                const uint64_t al = tc::make_smem_desc(a_lo + ks * 2 * 2048, 2048, 128);
                const uint64_t bh = tc::make_smem_desc(b_hi + ks * 2 * LBO_B, LBO_B, 128);
                const uint64_t bl = tc::make_smem_desc(b_lo + ks * 2 * LBO_B, LBO_B, 128);
                tc::mma_f16_ss(d, ah, bh, IDESC, ks > 0);
                tc::mma_f16_ss(d, al, bh, IDESC, 1);
                tc::mma_f16_ss(d, ah, bl, IDESC, 1);
            }
            tc::mma_commit(bar + buf);
        }
        if (prev_tile >= 0) epilogue(prev_tile, buf ^ 1);     // overlaps the MMAs just issued
        prev_tile = tile;
    }
    if (prev_tile >= 0) epilogue(prev_tile, (it - 1) & 1);
    if (!ok && tid == 0) atomicExch(status, 2);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, PW_TM);
}

// ---------------------------------------------------------------------------
// tail: conv5 (1x1, 128->128, ReLU) -> global average pool -> fc 128->512 (+BN1d
// folded) -> ReLU.  One CTA per crop: the 16x8 map is exactly one 128-row M tile.
// ---------------------------------------------------------------------------
struct TailCfg {
    static constexpr int CIN = 128, COUT = 128, PX = 128, FEAT = 512;      // FEAT == OSB_THREADS
    static constexpr int STG_HALF_B = (CIN / 8) * 128 * 16, STG_B = 2 * STG_HALF_B;   // 64 KB
    static constexpr int W_HALF_B = CIN * COUT * 2, W_B = 2 * W_HALF_B;               // 64 KB
    static constexpr int OFF_STG = 0, OFF_W = STG_B, OFF_BIAS = OFF_W + W_B;
    static constexpr int OFF_V = OFF_BIAS + COUT * 4, OFF_MISC = OFF_V + COUT * 4;
    static constexpr int SMEM_B = OFF_MISC + 128;
    // blob: W5 hi/lo | b5 [128] | fc W [128][512] fp32 | fc b [512]
    static constexpr int G_W = 0, G_B5 = W_B, G_FCW = G_B5 + 512, G_FCB = G_FCW + 128 * 512 * 4;
    static constexpr int G_TOTAL = G_FCB + 512 * 4;
};

template <bool PLANES>
__global__ void __launch_bounds__(OSB_THREADS, 1)
tail_tc_kernel(const float *__restrict__ x, float *__restrict__ feats,
               const unsigned char *__restrict__ wblob, int *__restrict__ status) {
    using C = TailCfg;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, grp = warp >> 2;
    const int crop = blockIdx.x;
    unsigned char *stg = smem + C::OFF_STG, *sW = smem + C::OFF_W;
    float *sBias = reinterpret_cast<float *>(smem + C::OFF_BIAS);
    float *sV = reinterpret_cast<float *>(smem + C::OFF_V);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + C::OFF_MISC);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bar + 2);
    if (warp == 0) tc::tmem_alloc(s_tmem, 128);
    if (tid == 0) { tc::mbar_init(bar, 1); tc::mbar_init(bar + 1, 1); tc::fence_mbar_init(); }
    if (tid < C::COUT) sBias[tid] = reinterpret_cast<const float *>(wblob + C::G_B5)[tid];
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *s_tmem;
    tc::pdl_launch_dependents();
    if (tid == 0) {
        // PLANES: a crop's operand planes [hl][16 chunks][128 px][8] ARE the staging layout: two 32 KB bulk copies
        tc::mbar_arrive_expect_tx(bar + 1, C::W_B + (PLANES ? C::STG_B : 0));
        tc::bulk_g2s(sW, wblob + C::G_W, C::W_B / 2, bar + 1);
        tc::bulk_g2s(sW + C::W_B / 2, wblob + C::G_W + C::W_B / 2, C::W_B / 2, bar + 1);
        tc::pdl_wait();
        if (PLANES) {
            const unsigned char *xb = reinterpret_cast<const unsigned char *>(x) + (size_t)crop * C::STG_B;
            tc::bulk_g2s(stg, xb, C::STG_HALF_B, bar + 1);
            tc::bulk_g2s(stg + C::STG_HALF_B, xb + C::STG_HALF_B, C::STG_HALF_B, bar + 1);
        }
    }
    tc::pdl_wait();
    const float *xin = x + (size_t)crop * C::PX * C::CIN;
    constexpr int F4 = C::CIN / 4;
#pragma unroll 4
    for (int idx = tid; idx < (PLANES ? 0 : 128 * F4); idx += OSB_THREADS) {
        const int m = idx / F4, f4 = idx - m * F4;
        const float4 v = *reinterpret_cast<const float4 *>(xin + (size_t)m * C::CIN + f4 * 4);
        __align__(8) __half2 h[2], l[2];
        split_hl2(v.x, v.y, h[0], l[0]);
        split_hl2(v.z, v.w, h[1], l[1]);
        const int off = (f4 >> 1) * 2048 + m * 16 + (f4 & 1) * 8;
        *reinterpret_cast<uint2 *>(stg + off) = *reinterpret_cast<uint2 *>(h);
        *reinterpret_cast<uint2 *>(stg + C::STG_HALF_B + off) = *reinterpret_cast<uint2 *>(l);
    }
    tc::fence_async_smem();
    bool ok = tc::mbar_wait(bar + 1, 0);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (__shfl_sync(0xffffffffu, warp, 0) == 0 && tc::elect_one()) {
        constexpr uint32_t IDESC = tc::make_idesc_f16(128, C::COUT);
        const uint32_t a_hi = tc::smem_u32(stg), a_lo = a_hi + C::STG_HALF_B;
        const uint32_t b_hi = tc::smem_u32(sW), b_lo = b_hi + C::W_HALF_B;
        constexpr uint32_t LBO_B = C::COUT * 16;
#pragma unroll
        for (int ks = 0; ks < C::CIN / 16; ks++) {
            const uint64_t ah = tc::make_smem_desc(a_hi + ks * 2 * 2048, 2048, 128);
            const uint64_t al = tc::make_smem_desc(a_lo + ks * 2 * 2048, 2048, 128);
            const uint64_t bh = tc::make_smem_desc(b_hi + ks * 2 * LBO_B, LBO_B, 128);
            const uint64_t bl = tc::make_smem_desc(b_lo + ks * 2 * LBO_B, LBO_B, 128);
            tc::mma_f16_ss(tmem, ah, bh, IDESC, ks > 0);
            tc::mma_f16_ss(tmem, al, bh, IDESC, 1);
            tc::mma_f16_ss(tmem, ah, bl, IDESC, 1);
        }
        tc::mma_commit(bar);
    }
    if (!tc::mbar_wait(bar, 0)) ok = false;
    tc::fence_after_sync();
    // relu(conv5) -> [px][129] floats over the (now dead) staging buffer, then column means
    float *sAct = reinterpret_cast<float *>(stg);
    const int m = quad * 32 + lane;
#pragma unroll 1
    for (int c0 = grp * 16; c0 < C::COUT; c0 += 16 * OSB_GROUPS) {
        float v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + c0, v);
#pragma unroll
        for (int j = 0; j < 16; j++) sAct[m * 129 + c0 + j] = fmaxf(v[j] + sBias[c0 + j], 0.f);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (tid < C::COUT) {
        float s = 0.f;
        for (int p = 0; p < C::PX; p++) s += sAct[p * 129 + tid];
        sV[tid] = s / (float)C::PX;
    }
    __syncthreads();
    const float *fw = reinterpret_cast<const float *>(wblob + C::G_FCW);
    const float *fb = reinterpret_cast<const float *>(wblob + C::G_FCB);
    {   // one output feature per thread (OSB_THREADS == FEAT == 512)
        float acc = fb[tid];
#pragma unroll 8
        for (int k = 0; k < C::CIN; k++) acc = fmaf(sV[k], fw[(size_t)k * C::FEAT + tid], acc);
        feats[(size_t)crop * C::FEAT + tid] = fmaxf(acc, 0.f);
    }
    if (!ok && tid == 0) atomicExch(status, 3);
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------------------
// stem on the tensor cores: crop + bilinear resize + normalise + conv7x7/2 + ReLU
// + maxpool3x3/2, one CTA per (crop, band of 8 pooled rows).
//
// A stride-2 7x7 conv on 3 channels is a stride-1 4x4 conv on the 2x2
// space-to-depth image (12 channels, padded to 16):  with ky = 2a+dy, kx = 2b+dx
//   out[cy][cx] = sum_{a,b<4} sum_{dy,dx,c} W[2a+dy][2b+dx][c] * S[cy+a][cx+b][(dy,dx,c)],
//   S[Y][X][(dy,dx,c)] = Rpad[2Y+dy][2X+dx][c],  Rpad = resized crop, zero-padded by 3.
// S is built once per band in the [chunk][pixel][8] operand layout, so the 16
// taps are 16 shifted GEMMs (K = 16, N = 16) exactly like the 3x3 convs above.
// ---------------------------------------------------------------------------
#ifndef SSB_STEM_PR
#define SSB_STEM_PR 4
#endif
struct StemCfg {
    static constexpr int PR = SSB_STEM_PR;          // pooled rows per CTA (4: two CTAs per SM, one builds its S map
                                                    // while the other's MMAs run; 8: round-1 shape, one CTA per SM)
    static constexpr int THREADS = PR >= 8 ? 512 : 256, MINB = PR >= 8 ? 1 : 2, GROUPS = THREADS / 128;
    static constexpr int NB = 64 / PR;              // bands per crop
    static constexpr int CROWS = 2 * PR + 1;        // conv rows needed
    static constexpr int SROWS = CROWS + 3;         // space-to-depth rows
    static constexpr int WPS = 67;                  // S row pitch (64 conv cols + 3)
    static constexpr int NPX = CROWS * WPS;         // output pixel index space
    static constexpr int NT = (NPX + 127) / 128;    // M tiles (5 / 9)
    static constexpr int MAP_PX = (NT * 128 + 3 * WPS + 3 + 7) / 8 * 8;
    static constexpr int PLANE_B = MAP_PX * 16;
    static constexpr int MAP_HALF_B = 2 * PLANE_B, MAP_B = 2 * MAP_HALF_B;
    static constexpr int W_HALF_B = 16 * 16 * 16 * 2, W_B = 2 * W_HALF_B;     // [tap][2][16][8]
    static constexpr int CP = 20;                   // floats per conv pixel in sConv (16 + 4 pad: the pooling stage's
                                                    // stride-2 column walk then touches distinct banks)
    static constexpr int CONV_B = CROWS * 64 * CP * 4;
    // sConv (the drained conv rows) reuses the operand map: it is first written after the last MMA has read the map
    static constexpr int OFF_MAP = 0, OFF_CONV = 0, OFF_W = MAP_B;
    static constexpr int OFF_BIAS = OFF_W + W_B, OFF_MISC = OFF_BIAS + 64;
    static constexpr int SMEM_B = OFF_MISC + 64;
    static constexpr int TM_COLS = NT * 32, TM_ALLOC = TM_COLS <= 256 ? 256 : 512;
    static constexpr int G_W = 0, G_BIAS = W_B, G_TOTAL = W_B + 128;
    static_assert(64 % PR == 0, "bands");
    static_assert(NT * 128 + 3 * WPS + 3 <= MAP_PX, "map guard");
    static_assert(CONV_B <= MAP_B, "sConv aliases the operand map");
    static_assert(SMEM_B <= 232448, "shared memory");
};

template <bool PLANES>
__global__ void __launch_bounds__(StemCfg::THREADS, StemCfg::MINB)
stem_tc_kernel(const uint8_t *__restrict__ img, int H, int W, int pitch, const int *__restrict__ boxes,
               const unsigned char *__restrict__ wblob, float *__restrict__ out, int *__restrict__ status,
               long long *__restrict__ dbg) {
    using C = StemCfg;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, grp = warp >> 2;
    const int crop = blockIdx.x / C::NB, band = blockIdx.x % C::NB;
    const int py0 = band * C::PR, cy0 = 2 * py0 - 1;      // first conv row of the band
    unsigned char *sMap = smem + C::OFF_MAP, *sW = smem + C::OFF_W;
    float *sConv = reinterpret_cast<float *>(smem + C::OFF_CONV);
    float *sBias = reinterpret_cast<float *>(smem + C::OFF_BIAS);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + C::OFF_MISC);     // [0] mma, [1] weights
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bar + 2);
    if (warp == 0) tc::tmem_alloc(s_tmem, C::TM_ALLOC);
    if (tid == 0) { tc::mbar_init(bar, 1); tc::mbar_init(bar + 1, 1); tc::fence_mbar_init(); }
    if (tid < 16) sBias[tid] = reinterpret_cast<const float *>(wblob + C::G_BIAS)[tid];
    // the operand map is written in full below (every S pixel, all 16 channels); only the guard
    // pixels past the band's S rows (read by the shifted taps of the last tile) are zeroed here
    for (int i = tid; i < (C::MAP_PX - C::SROWS * C::WPS) * 4; i += C::THREADS) {
        const int px = C::SROWS * C::WPS + i / 4, pl = i & 3;       // 4 planes: hi c0, hi c1, lo c0, lo c1
        *reinterpret_cast<uint4 *>(sMap + (pl >> 1) * C::MAP_HALF_B + (pl & 1) * C::PLANE_B + px * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *s_tmem;
    tc::pdl_launch_dependents();
    if (tid == 0) {
        tc::mbar_arrive_expect_tx(bar + 1, C::W_B);
        tc::bulk_g2s(sW, wblob + C::G_W, C::W_B, bar + 1);
    }
    tc::pdl_wait();                    // the crop boxes come from the previous kernel of the stream
    // ---- build S: every resized+normalised pixel of the band lands in exactly one slot
    const int bx1 = boxes[crop * 4 + 0], by1 = boxes[crop * 4 + 1];
    const int cw = boxes[crop * 4 + 2] - bx1, ch = boxes[crop * 4 + 3] - by1;
    const float sc_y = (float)ch / 256.f, sc_x = (float)cw / 128.f;
    const int gy0 = 2 * cy0 - 3;                               // resized row of S row 0, dy = 0
    int dbg_n = 0;
    auto stamp = [&]() { if (dbg && blockIdx.x == 0 && tid == 0 && dbg_n < 15) dbg[41 + dbg_n++] = clock64(); };   // slots 40..: stem
    stamp();
    // One thread per S pixel q = (Y, X): its 2x2 resized pixels x 3 channels = the 12 real channels of the
    // pixel's two 16-byte operand rows (hi and lo), written with four conflict-free 16-byte stores.
    // All 48 byte loads are unconditional (coordinates clamped into the crop, out-of-range samples zeroed
    // afterwards) so they are in flight together; u8 -> f32 is the 2^23 magic-number trick (LOP3 + FADD,
    // full rate; I2F runs at a quarter of it); (v/255 - mean)/std is one FMA.
    const bool crop_ok = cw > 0 && ch > 0;
    const int cwm = crop_ok ? cw - 1 : 0, chm = crop_ok ? ch - 1 : 0;
    const uint8_t *cbase = img + (size_t)(crop_ok ? by1 : 0) * pitch + (size_t)(crop_ok ? bx1 : 0) * 3;
    const float nsc[3] = {1.0f / (255.0f * 0.229f), 1.0f / (255.0f * 0.224f), 1.0f / (255.0f * 0.225f)};
    const float nof[3] = {-0.485f / 0.229f, -0.456f / 0.224f, -0.406f / 0.225f};
    auto u8f = [](uint8_t b) { return __uint_as_float(0x4B000000u | (unsigned)b) - 8388608.0f; };
    for (int q = tid; q < C::SROWS * C::WPS; q += C::THREADS) {
        const int Y = q / C::WPS, X = q - Y * C::WPS;
        int xo0[2], xo1[2];
        float lxv[2];
        bool xin[2];
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
            const int gx = 2 * X + dx - 3;
            float sx = sc_x * ((float)gx + 0.5f) - 0.5f;
            if (sx < 0.f) sx = 0.f;
            int x0 = (int)sx;
            lxv[dx] = sx - (float)x0;
            x0 = x0 < cwm ? x0 : cwm;
            xo0[dx] = x0 * 3;
            xo1[dx] = (x0 + (x0 < cwm ? 1 : 0)) * 3;
            xin[dx] = gx >= 0 && gx < 128;
        }
        uint8_t pa[2][2][2][3], pb[2][2][2][3];       // [dy][source row 0/1][dx][c] at x0 (pa) and x1 (pb)
        float lyv[2];
        bool yin[2];
#pragma unroll
        for (int dy = 0; dy < 2; dy++) {
            const int gy = gy0 + 2 * Y + dy;
            float sy = sc_y * ((float)gy + 0.5f) - 0.5f;
            if (sy < 0.f) sy = 0.f;
            int y0 = (int)sy;
            lyv[dy] = sy - (float)y0;
            y0 = y0 < chm ? y0 : chm;
            const int y1 = y0 + (y0 < chm ? 1 : 0);
            yin[dy] = crop_ok && gy >= 0 && gy < 256;
            const uint8_t *r0 = cbase + (size_t)y0 * pitch, *r1 = cbase + (size_t)y1 * pitch;
#pragma unroll
            for (int dx = 0; dx < 2; dx++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    pa[dy][0][dx][c] = r0[xo0[dx] + c];  pb[dy][0][dx][c] = r0[xo1[dx] + c];
                    pa[dy][1][dx][c] = r1[xo0[dx] + c];  pb[dy][1][dx][c] = r1[xo1[dx] + c];
                }
        }
        __align__(16) __half hv[16];
        __align__(16) __half lv[16];
#pragma unroll
        for (int e = 12; e < 16; e++) { hv[e] = __float2half_rn(0.f); lv[e] = __float2half_rn(0.f); }
#pragma unroll
        for (int dy = 0; dy < 2; dy++)
#pragma unroll
            for (int dx = 0; dx < 2; dx++) {
                const float ly = lyv[dy], hy = 1.f - ly, lx = lxv[dx], hx = 1.f - lx;
                const bool in = yin[dy] && xin[dx];
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float val = hy * (hx * u8f(pa[dy][0][dx][c]) + lx * u8f(pb[dy][0][dx][c])) +
                                      ly * (hx * u8f(pa[dy][1][dx][c]) + lx * u8f(pb[dy][1][dx][c]));
                    const float v = in ? fmaf(val, nsc[c], nof[c]) : 0.f;
                    split_hl(v, hv[(dy * 2 + dx) * 3 + c], lv[(dy * 2 + dx) * 3 + c]);
                }
            }
        unsigned char *d = sMap + q * 16;
        *reinterpret_cast<uint4 *>(d) = *reinterpret_cast<uint4 *>(&hv[0]);
        *reinterpret_cast<uint4 *>(d + C::PLANE_B) = *reinterpret_cast<uint4 *>(&hv[8]);
        *reinterpret_cast<uint4 *>(d + C::MAP_HALF_B) = *reinterpret_cast<uint4 *>(&lv[0]);
        *reinterpret_cast<uint4 *>(d + C::MAP_HALF_B + C::PLANE_B) = *reinterpret_cast<uint4 *>(&lv[8]);
    }
    tc::fence_async_smem();
    bool ok = tc::mbar_wait(bar + 1, 0);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    stamp();                                   // S built
    // warp-uniform issue by an elected lane (~3x cheaper per MMA than `if (tid == 0)`, tc_common.cuh);
    // hi and lo weight rows concatenated along N: D[:, 0:16] += Ah*Bh + Al*Bh, D[:, 16:32] += Ah*Bl --
    // two reads of the 4 KB A tile (what bounds an M=128, K=16 MMA) per product instead of three
    if (__shfl_sync(0xffffffffu, warp, 0) == 0 && tc::elect_one()) {
        constexpr uint32_t IDESC = tc::make_idesc_f16(128, 16);
        constexpr uint32_t IDESC_CAT = tc::make_idesc_f16(128, 32);
        const uint64_t ah0 = tc::make_smem_desc(tc::smem_u32(sMap), C::PLANE_B, 128);
        const uint64_t al0 = desc_adv(ah0, C::MAP_HALF_B / 16);
        const uint64_t bc0 = tc::make_smem_desc(tc::smem_u32(sW), 32 * 16, 128);      // [tap][kc][hi 16 | lo 16][8]
#pragma unroll 1
        for (int t = 0; t < C::NT; t++) {
            const uint32_t d = tmem + t * 32;
            const uint64_t aht = desc_adv(ah0, t * 128), alt = desc_adv(al0, t * 128);
#pragma unroll
            for (int tap = 0; tap < 16; tap++) {
                const int po = (tap >> 2) * C::WPS + (tap & 3);
                tc::mma_f16_ss(d, desc_adv(aht, po), desc_adv(bc0, tap * 64), IDESC_CAT, tap != 0);
                tc::mma_f16_ss(d, desc_adv(alt, po), desc_adv(bc0, tap * 64), IDESC, 1);
            }
        }
        tc::mma_commit(bar);
    }
    if (!tc::mbar_wait(bar, 0)) ok = false;
    tc::fence_after_sync();
    stamp();                                   // MMAs done
    // ---- epilogue: bias + ReLU -> sConv[lcy][cx][16]; rows outside the conv map = -inf
    for (int t = grp; t < C::NT; t += C::GROUPS) {
        const int p = t * 128 + quad * 32 + lane;
        float v[16], w[16];
        tc::tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + t * 32, v);
        tc::tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + t * 32 + 16, w);
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] += w[j];
        const int lcy = p / C::WPS, cx = p - lcy * C::WPS;
        if (lcy < C::CROWS && cx < 64) {
            const int gcy = cy0 + lcy;
            const bool inside = gcy >= 0 && gcy < 128;
            float *o = sConv + ((size_t)lcy * 64 + cx) * C::CP;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                float4 r;
                r.x = inside ? fmaxf(v[j] + sBias[j], 0.f) : -INFINITY;
                r.y = inside ? fmaxf(v[j + 1] + sBias[j + 1], 0.f) : -INFINITY;
                r.z = inside ? fmaxf(v[j + 2] + sBias[j + 2], 0.f) : -INFINITY;
                r.w = inside ? fmaxf(v[j + 3] + sBias[j + 3], 0.f) : -INFINITY;
                *reinterpret_cast<float4 *>(o + j) = r;
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    stamp();                                   // conv drained
    // ---- maxpool 3x3 s2 p1 -> operand planes [crop][hl][2 chunks][64*32 px][8] (PLANES) ...
    if (PLANES) {
        unsigned char *ob = reinterpret_cast<unsigned char *>(out) + (size_t)crop * (4 * 16 * 2048);
        for (int o = tid; o < C::PR * 32 * 2; o += C::THREADS) {
            const int px = o & 31, ch = (o >> 5) & 1, pr = o >> 6;
            float4 m0 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), m1 = m0;
#pragma unroll
            for (int dy = 0; dy < 3; dy++) {
                const int lcy = 2 * pr + dy;
#pragma unroll
                for (int dx = 0; dx < 3; dx++) {
                    const int cx = 2 * px - 1 + dx;
                    if (cx < 0 || cx >= 64) continue;
                    const float *src = sConv + ((size_t)lcy * 64 + cx) * C::CP + ch * 8;
                    const float4 v0 = *reinterpret_cast<const float4 *>(src), v1 = *reinterpret_cast<const float4 *>(src + 4);
                    m0.x = fmaxf(m0.x, v0.x); m0.y = fmaxf(m0.y, v0.y); m0.z = fmaxf(m0.z, v0.z); m0.w = fmaxf(m0.w, v0.w);
                    m1.x = fmaxf(m1.x, v1.x); m1.y = fmaxf(m1.y, v1.y); m1.z = fmaxf(m1.z, v1.z); m1.w = fmaxf(m1.w, v1.w);
                }
            }
            __align__(16) __half2 h[4];
            __align__(16) __half2 l[4];
            split_hl2(m0.x, m0.y, h[0], l[0]); split_hl2(m0.z, m0.w, h[1], l[1]);
            split_hl2(m1.x, m1.y, h[2], l[2]); split_hl2(m1.z, m1.w, h[3], l[3]);
            unsigned char *dst = ob + (size_t)ch * 2048 * 16 + (size_t)((py0 + pr) * 32 + px) * 16;
            *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<uint4 *>(h);
            *reinterpret_cast<uint4 *>(dst + 2 * 16 * 2048) = *reinterpret_cast<uint4 *>(l);
        }
    }
    // ... or float32 out[crop][py][px][16]
    for (int o = tid; o < (PLANES ? 0 : C::PR * 32 * 4); o += C::THREADS) {
        const int c4 = o & 3, px = (o >> 2) & 31, pr = o >> 7;
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int dy = 0; dy < 3; dy++) {
            const int lcy = 2 * pr + dy;
#pragma unroll
            for (int dx = 0; dx < 3; dx++) {
                const int cx = 2 * px - 1 + dx;
                if (cx < 0 || cx >= 64) continue;
                const float4 v = *reinterpret_cast<const float4 *>(sConv + ((size_t)lcy * 64 + cx) * C::CP + c4 * 4);
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
        *reinterpret_cast<float4 *>(out + (((size_t)crop * 64 + py0 + pr) * 32 + px) * 16 + c4 * 4) = m;
    }
    stamp();                                   // pooled + stored
    if (dbg && blockIdx.x == 0 && tid == 0) dbg[40] = dbg_n;
    if (!ok && tid == 0) atomicExch(status, 4);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, C::TM_ALLOC);
}

// launch with the programmatic-dependent-launch attribute (tc_common.cuh: pdl_wait / pdl_launch_dependents)
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl_t(void (*kern)(KArgs...), int grid, int threads, int smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = ssb_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int smem, cudaStream_t st, Args... args) {
    return launch_pdl_t(kern, grid, OSB_THREADS, smem, st, args...);
}

using PwT1 = PwCfg<64, 64, 64, 32, true>;     // after conv2: 64x32x64 -> 32x16x64
using PwT2 = PwCfg<96, 96, 32, 16, true>;     // after conv3: 32x16x96 -> 16x8x96

template <class C, bool PLANES>
int launch_pw_tc(const float *x, float *y, const unsigned char *w, int n, int *status, int sms, cudaStream_t st) {
    static const int key = ssb_new_key();
    if (ssb_first_on_device(key))
        SSB_CHECK_CUDA(cudaFuncSetAttribute(pw_tc_kernel<C, PLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_B));
    constexpr int HO = C::POOL ? C::H / 2 : C::H, WO = C::POOL ? C::W / 2 : C::W;
    const long long total = (long long)n * HO * WO;
    const int per = C::POOL ? 32 : 128;
    int tiles = (int)((total + per - 1) / per);
    int grid = tiles < sms ? tiles : sms;
    SSB_CHECK_CUDA(launch_pdl(pw_tc_kernel<C, PLANES>, grid, C::SMEM_B, st, x, y, w, n, status));
    g_ssb_launches++;
    return 0;
}

}  // namespace

int64_t ssb_reid_tc_block_bytes(int b) {
    switch (b) {
        case 0: return Blk0::G_TOTAL;
        case 1: return Blk1::G_TOTAL;
        case 2: return Blk2::G_TOTAL;
        case 3: return Blk3::G_TOTAL;
        case 4: return Blk4::G_TOTAL;
        case 5: return Blk5::G_TOTAL;
    }
    return -1;
}

long long *g_ssb_tc_dbg = nullptr;      // device buffer of 64 int64 (ssb_reid_tc_debug), usually null

#ifdef SSB_BASELINES
int ssb_reid_tc_block(int b, const float *x, float *y, const unsigned char *w, int n, int *status,
                      cudaStream_t st) {
    long long *dbg = g_ssb_tc_dbg;
    switch (b) {
        case 0: return launch_block<Blk0>(x, y, w, n, status, dbg, st);
        case 1: return launch_block<Blk1>(x, y, w, n, status, dbg, st);
        case 2: return launch_block<Blk2>(x, y, w, n, status, dbg, st);
        case 3: return launch_block<Blk3>(x, y, w, n, status, dbg, st);
        case 4: return launch_block<Blk4>(x, y, w, n, status, dbg, st);
        case 5: return launch_block<Blk5>(x, y, w, n, status, dbg, st);
    }
    ssb_set_error("bad OSBlock index %d", b);
    return -1;
}
#endif

static int num_sms() { return ssb_num_sms(); }

// which: 0 = transition after conv2, 1 = transition after conv3, 2 = tail (conv5+GAP+fc)
int64_t ssb_reid_tc_aux_bytes(int which) {
    switch (which) {
        case 0: return PwT1::G_TOTAL;
        case 1: return PwT2::G_TOTAL;
        case 2: return TailCfg::G_TOTAL;
        case 3: return StemCfg::G_TOTAL;
    }
    return -1;
}

int ssb_reid_tc_stem(const uint8_t *img, int h, int w, int pitch, const int *boxes, const unsigned char *wsec,
                     float *out, int n, int *status, cudaStream_t st, int planes) {
    static const int key = ssb_new_key();
    if (ssb_first_on_device(key)) {
        SSB_CHECK_CUDA(cudaFuncSetAttribute(stem_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, StemCfg::SMEM_B));
        SSB_CHECK_CUDA(cudaFuncSetAttribute(stem_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, StemCfg::SMEM_B));
    }
    long long *dbg = g_ssb_tc_dbg;
    if (planes)
        SSB_CHECK_CUDA(launch_pdl_t(stem_tc_kernel<true>, n * StemCfg::NB, StemCfg::THREADS, StemCfg::SMEM_B, st, img, h, w, pitch, boxes, wsec, out, status, dbg));
    else
        SSB_CHECK_CUDA(launch_pdl_t(stem_tc_kernel<false>, n * StemCfg::NB, StemCfg::THREADS, StemCfg::SMEM_B, st, img, h, w, pitch, boxes, wsec, out, status, dbg));
    g_ssb_launches++;
    return 0;
}

int ssb_reid_tc_aux(int which, const float *x, float *y, const unsigned char *w, int n, int *status,
                    cudaStream_t st, int planes) {
    switch (which) {
        case 0: return planes ? launch_pw_tc<PwT1, true>(x, y, w, n, status, num_sms(), st)
                              : launch_pw_tc<PwT1, false>(x, y, w, n, status, num_sms(), st);
        case 1: return planes ? launch_pw_tc<PwT2, true>(x, y, w, n, status, num_sms(), st)
                              : launch_pw_tc<PwT2, false>(x, y, w, n, status, num_sms(), st);
        case 2: {
            static const int key = ssb_new_key();
            if (ssb_first_on_device(key)) {
                SSB_CHECK_CUDA(cudaFuncSetAttribute(tail_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TailCfg::SMEM_B));
                SSB_CHECK_CUDA(cudaFuncSetAttribute(tail_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TailCfg::SMEM_B));
            }
            if (planes) SSB_CHECK_CUDA(launch_pdl(tail_tc_kernel<true>, n, TailCfg::SMEM_B, st, x, y, w, status));
            else SSB_CHECK_CUDA(launch_pdl(tail_tc_kernel<false>, n, TailCfg::SMEM_B, st, x, y, w, status));
            g_ssb_launches++;
            return 0;
        }
    }
    ssb_set_error("bad aux kernel index %d", which);
    return -1;
}
