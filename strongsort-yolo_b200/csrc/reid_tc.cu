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
    static constexpr int TM_LC = 0, TM_C3 = NT * MIDP;
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
    static constexpr int SMEM_B = OFF_MISC + 1024;
    static_assert(NSTAGE * STG_B <= 2 * MAP_B, "x staging must fit in the P+Q maps");
    static_assert(TM_COLS <= 512, "TMEM columns");
    static_assert(SMEM_B <= 232448, "shared memory");
    static_assert(MIDP % 16 == 0 && COUT % 16 == 0 && CIN % 16 == 0, "MMA shapes");
    static_assert(H % R == 0 && H / R == NB, "bands");
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

struct TrueT { static constexpr bool value = true; };
struct FalseT { static constexpr bool value = false; };

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
__global__ void __launch_bounds__(128, 1)
osblock_tc_kernel(const float *__restrict__ x, float *__restrict__ y,
                  const unsigned char *__restrict__ wblob, int n_crops, int *__restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    using P = Par<C>;
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int crop = blockIdx.x / C::NB, band = blockIdx.x % C::NB;
    const int row0 = band * C::R - C::HALO;            // image row of local row lr = 1

    unsigned char *sX1 = smem + C::OFF_X1, *sP = smem + C::OFF_P, *sQ = smem + C::OFF_Q;
    unsigned char *sW1 = smem + C::OFF_W1, *sC3 = smem + C::OFF_C3;
    float *sPar = reinterpret_cast<float *>(smem + C::OFF_PAR);
    float *sGap = reinterpret_cast<float *>(smem + C::OFF_GAP);      // [4][MIDP] partial sums
    uint64_t *bar_mma = reinterpret_cast<uint64_t *>(smem + C::OFF_MISC);
    uint64_t *bar_w = bar_mma + 1;
    uint64_t *bar_stg = bar_mma + 2;                   // [2] one per x-staging buffer
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bar_mma + 4);
    float *s_scr = reinterpret_cast<float *>(smem + C::OFF_MISC + 64);    // [4][MIDP] warp partials
    float *s_mean = s_scr + 4 * C::MIDP;                                  // [MIDP]
    float *s_gate = s_mean + C::MIDP;                                     // [MIDP]

    if (warp == 0) tc::tmem_alloc(s_tmem, 512);
    if (tid == 0) {
        tc::mbar_init(bar_mma, 1);
        tc::mbar_init(bar_w, 1);
        tc::mbar_init(bar_stg, 1);
        tc::mbar_init(bar_stg + 1, 1);
        tc::fence_mbar_init();
    }
    for (int i = tid; i < C::NPAR; i += 128)
        sPar[i] = reinterpret_cast<const float *>(wblob + C::G_PAR)[i];
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *s_tmem;
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

    // ------------------------------------------------------------------
    // phase 1: X1 = relu(conv1(x)) on every band tile; downsample on inner tiles
    // ------------------------------------------------------------------
    constexpr uint32_t IDESC_MID = tc::make_idesc_f16(128, C::MIDP);
    constexpr uint32_t IDESC_OUT = tc::make_idesc_f16(128, C::COUT);
    const float *xin = x + (size_t)crop * C::H * C::W * C::CIN;
    // a waiter may lag an mbarrier by at most one phase, so every staging buffer has
    // its own barrier: tile t commits to bar_stg[t % NSTAGE] and is waited before reuse
    uint32_t stg_phase[2] = {0, 0};
    for (int t = 0; t < C::NT; t++) {
        const int sb = t % C::NSTAGE;
        unsigned char *stg = sP + sb * C::STG_B;
        if (t >= C::NSTAGE) {                                // MMAs that read this buffer are done
            if (!tc::mbar_wait(bar_stg + sb, stg_phase[sb])) ok = false;
            stg_phase[sb] ^= 1;
        }
        // stage tile t of x: [CIN/8][128][8] hi, then lo
        constexpr int F4 = C::CIN / 4;
        for (int idx = tid; idx < 128 * F4; idx += 128) {
            const int px = idx / F4, f4 = idx - px * F4;
            int gr, gc;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pixel_valid(t * 128 + px, gr, gc))
                v = *reinterpret_cast<const float4 *>(xin + ((size_t)gr * C::W + gc) * C::CIN + f4 * 4);
            __half h[4], l[4];
            split_hl(v.x, h[0], l[0]); split_hl(v.y, h[1], l[1]);
            split_hl(v.z, h[2], l[2]); split_hl(v.w, h[3], l[3]);
            const int off = (f4 >> 1) * 2048 + px * 16 + (f4 & 1) * 8;
            *reinterpret_cast<uint2 *>(stg + off) = *reinterpret_cast<uint2 *>(h);
            *reinterpret_cast<uint2 *>(stg + C::STG_HALF_B + off) = *reinterpret_cast<uint2 *>(l);
        }
        tc::fence_async_smem();
        if (t == 0) { if (!tc::mbar_wait(bar_w, w_phase)) ok = false; w_phase ^= 1; }
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        if (tid == 0) {
            const uint32_t a_hi = tc::smem_u32(stg), a_lo = a_hi + C::STG_HALF_B;
            const uint32_t b_hi = tc::smem_u32(sW1), b_lo = b_hi + C::C1W_HALF_B;
            constexpr uint32_t LBO_B1 = C::MIDP * 16;
            for (int ks = 0; ks < C::CIN / 16; ks++) {
                const uint64_t ah = tc::make_smem_desc(a_hi + ks * 2 * 2048, 2048, 128);
                const uint64_t al = tc::make_smem_desc(a_lo + ks * 2 * 2048, 2048, 128);
                const uint64_t bh = tc::make_smem_desc(b_hi + ks * 2 * LBO_B1, LBO_B1, 128);
                const uint64_t bl = tc::make_smem_desc(b_lo + ks * 2 * LBO_B1, LBO_B1, 128);
                const uint32_t d = tmem + C::TM_LC + t * C::MIDP;
                tc::mma_f16_ss(d, ah, bh, IDESC_MID, ks > 0);
                tc::mma_f16_ss(d, al, bh, IDESC_MID, 1);
                tc::mma_f16_ss(d, ah, bl, IDESC_MID, 1);
            }
            if (C::DOWN && t >= C::IT0 && t < C::IT1) {
                const uint32_t d_hi = tc::smem_u32(sW1) + C::C1W_B, d_lo = d_hi + C::DNW_HALF_B;
                constexpr uint32_t LBO_BD = C::COUT * 16;
                for (int ks = 0; ks < C::CIN / 16; ks++) {
                    const uint64_t ah = tc::make_smem_desc(a_hi + ks * 2 * 2048, 2048, 128);
                    const uint64_t al = tc::make_smem_desc(a_lo + ks * 2 * 2048, 2048, 128);
                    const uint64_t bh = tc::make_smem_desc(d_hi + ks * 2 * LBO_BD, LBO_BD, 128);
                    const uint64_t bl = tc::make_smem_desc(d_lo + ks * 2 * LBO_BD, LBO_BD, 128);
                    const uint32_t d = tmem + C::TM_C3 + (t - C::IT0) * C::COUT;
                    tc::mma_f16_ss(d, ah, bh, IDESC_OUT, ks > 0);
                    tc::mma_f16_ss(d, al, bh, IDESC_OUT, 1);
                    tc::mma_f16_ss(d, ah, bl, IDESC_OUT, 1);
                }
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
    auto drain_to_map = [&](unsigned char *dst, const float *bias, auto gapacc, float *gap) {
        constexpr bool GAPACC = decltype(gapacc)::value;
        for (int t = 0; t < C::NT; t++) {
            const int p = t * 128 + warp * 32 + lane;
            int gr, gc;
            const bool valid = pixel_valid(p, gr, gc);
            const int lr = p / C::WP;
            const bool own = valid && lr >= 1 + C::HALO && lr < 1 + C::HALO + C::R;
            unsigned char *d_hi = dst + (C::GUARD + p) * 16, *d_lo = d_hi + C::MAP_HALF_B;
#pragma unroll
            for (int c0 = 0; c0 < C::MIDP; c0 += 16) {
                float v[16];
                tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + C::TM_LC + t * C::MIDP + c0, v);
                __align__(16) __half h[16];
                __align__(16) __half l[16];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const float f = valid ? fmaxf(v[j] + bias[c0 + j], 0.f) : 0.f;
                    if (GAPACC) { if (own) gap[c0 + j] += f; }
                    split_hl(f, h[j], l[j]);
                }
                const int pl = (c0 >> 3) * C::PLANE_B;
                *reinterpret_cast<uint4 *>(d_hi + pl) = *reinterpret_cast<uint4 *>(&h[0]);
                *reinterpret_cast<uint4 *>(d_hi + pl + C::PLANE_B) = *reinterpret_cast<uint4 *>(&h[8]);
                *reinterpret_cast<uint4 *>(d_lo + pl) = *reinterpret_cast<uint4 *>(&l[0]);
                *reinterpret_cast<uint4 *>(d_lo + pl + C::PLANE_B) = *reinterpret_cast<uint4 *>(&l[8]);
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    };

    drain_to_map(sX1, sPar + P::B1, FalseT{}, nullptr);

    // ------------------------------------------------------------------
    // phase 2: four streams of dense 3x3 convs + gated conv3 accumulation
    // ------------------------------------------------------------------
    const float *w3 = reinterpret_cast<const float *>(wblob + C::G_W3);      // [MIDP][COUT]
    int lc = 0;
    for (int s = 0; s < 4; s++) {
        const unsigned char *src = sX1;
        unsigned char *dst = sP;
        for (int k = 0; k <= s; k++, lc++) {
            if (!tc::mbar_wait(bar_w, w_phase)) ok = false;       // this LightConv's weights landed
            w_phase ^= 1;
            tc::fence_after_sync();
            if (tid == 0) {
                const uint32_t a_hi = tc::smem_u32(src) + C::GUARD * 16, a_lo = a_hi + C::MAP_HALF_B;
                const uint32_t b_hi = tc::smem_u32(sW1), b_lo = b_hi + C::LCW_HALF_B;
                constexpr uint32_t LBO_B = C::MIDP * 16;
                for (int t = 0; t < C::NT; t++) {
                    const uint32_t d = tmem + C::TM_LC + t * C::MIDP;
                    uint32_t acc = 0;
#pragma unroll 1
                    for (int tap = 0; tap < 9; tap++) {
                        const int off = ((tap / 3 - 1) * C::WP + (tap % 3 - 1) + t * 128) * 16;
#pragma unroll
                        for (int ks = 0; ks < C::MIDP / 16; ks++) {
                            const uint32_t ka = ks * 2 * C::PLANE_B;
                            const uint32_t kb = (tap * C::MCH + ks * 2) * LBO_B;
                            const uint64_t ah = tc::make_smem_desc(a_hi + off + ka, C::PLANE_B, 128);
                            const uint64_t al = tc::make_smem_desc(a_lo + off + ka, C::PLANE_B, 128);
                            const uint64_t bh = tc::make_smem_desc(b_hi + kb, LBO_B, 128);
                            const uint64_t bl = tc::make_smem_desc(b_lo + kb, LBO_B, 128);
                            tc::mma_f16_ss(d, ah, bh, IDESC_MID, acc);
                            tc::mma_f16_ss(d, al, bh, IDESC_MID, 1);
                            tc::mma_f16_ss(d, ah, bl, IDESC_MID, 1);
                            acc = 1;
                        }
                    }
                }
                tc::mma_commit(bar_mma);
            }
            mma.wait();
            tc::fence_after_sync();
            // the weight buffer is free again: prefetch the next LightConv's weights
            if (tid == 0 && lc + 1 < 10) {
                tc::mbar_arrive_expect_tx(bar_w, C::LCW_B);
                tc::bulk_g2s(sW1, wblob + C::G_LCW + (size_t)(lc + 1) * C::LCW_B, C::LCW_B, bar_w);
            }
            const bool last = (k == s);
            if (!last) {
                drain_to_map(dst, sPar + P::BLC + lc * C::MIDP, FalseT{}, nullptr);
            } else {
                float gap[C::MIDP];
#pragma unroll
                for (int j = 0; j < C::MIDP; j++) gap[j] = 0.f;
                drain_to_map(dst, sPar + P::BLC + lc * C::MIDP, TrueT{}, gap);
                // ---- ChannelGate: band-partial sums -> cluster -> mean -> MLP -> sigmoid
#pragma unroll
                for (int j = 0; j < C::MIDP; j++) {
                    float vs = gap[j];
#pragma unroll
                    for (int o = 16; o; o >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, o);
                    if (lane == 0) s_scr[warp * C::MIDP + j] = vs;
                }
                __syncthreads();
                if (tid < C::MIDP)
                    sGap[s * C::MIDP + tid] = (s_scr[tid] + s_scr[C::MIDP + tid]) +
                                              (s_scr[2 * C::MIDP + tid] + s_scr[3 * C::MIDP + tid]);
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
                for (int u = tid; u < C::COUT * C::MCH; u += 128) {
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
                if (tid == 0) {
                    const uint32_t a_hi = tc::smem_u32(dst) + C::GUARD * 16, a_lo = a_hi + C::MAP_HALF_B;
                    const uint32_t b_hi = tc::smem_u32(sC3), b_lo = b_hi + C::C3W_HALF_B;
                    constexpr uint32_t LBO_B3 = C::COUT * 16;
                    for (int i = 0; i < C::NIT; i++) {
                        const uint32_t d = tmem + C::TM_C3 + i * C::COUT;
                        const int off = (C::IT0 + i) * 128 * 16;
#pragma unroll
                        for (int ks = 0; ks < C::MIDP / 16; ks++) {
                            const uint32_t ka = ks * 2 * C::PLANE_B, kb = ks * 2 * LBO_B3;
                            const uint64_t ah = tc::make_smem_desc(a_hi + off + ka, C::PLANE_B, 128);
                            const uint64_t al = tc::make_smem_desc(a_lo + off + ka, C::PLANE_B, 128);
                            const uint64_t bh = tc::make_smem_desc(b_hi + kb, LBO_B3, 128);
                            const uint64_t bl = tc::make_smem_desc(b_lo + kb, LBO_B3, 128);
                            tc::mma_f16_ss(d, ah, bh, IDESC_OUT, (C::DOWN || s > 0 || ks > 0) ? 1 : 0);
                            tc::mma_f16_ss(d, al, bh, IDESC_OUT, 1);
                            tc::mma_f16_ss(d, ah, bl, IDESC_OUT, 1);
                        }
                    }
                    if (s == 3) tc::mma_commit(bar_mma);
                }
            }
            src = dst;
            dst = (dst == sP) ? sQ : sP;
        }
    }
    mma.wait();                       // the last stream's conv3 MMAs
    tc::fence_after_sync();

    // ------------------------------------------------------------------
    // final epilogue: y = relu(conv3 + bias (+ downsample already in TMEM) (+ x))
    // ------------------------------------------------------------------
    float *yout = y + (size_t)crop * C::H * C::W * C::COUT;
    for (int i = 0; i < C::NIT; i++) {
        const int p = (C::IT0 + i) * 128 + warp * 32 + lane;
        int gr, gc;
        const bool valid = pixel_valid(p, gr, gc);
        const int lr = p / C::WP;
        const bool own = valid && lr >= 1 + C::HALO && lr < 1 + C::HALO + C::R;
#pragma unroll 1
        for (int c0 = 0; c0 < C::COUT; c0 += 16) {
            float v[16];
            tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + C::TM_C3 + i * C::COUT + c0, v);
            if (own) {
                float *o = yout + ((size_t)gr * C::W + gc) * C::COUT + c0;
                const float *xr = xin + ((size_t)gr * C::W + gc) * C::CIN + c0;
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 r;
                    r.x = v[j] + sPar[P::B3 + c0 + j];
                    r.y = v[j + 1] + sPar[P::B3 + c0 + j + 1];
                    r.z = v[j + 2] + sPar[P::B3 + c0 + j + 2];
                    r.w = v[j + 3] + sPar[P::B3 + c0 + j + 3];
                    if (!C::DOWN) {
                        const float4 xv = *reinterpret_cast<const float4 *>(xr + j);
                        r.x += xv.x; r.y += xv.y; r.z += xv.z; r.w += xv.w;
                    }
                    r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f);
                    r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
                    *reinterpret_cast<float4 *>(o + j) = r;
                }
            }
        }
    }
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
using Blk0 = BlkCfg<16, 16, 16, 64, 64, 32, 16, 4, 4, true, 2>;
using Blk1 = BlkCfg<64, 16, 16, 64, 64, 32, 16, 4, 4, false, 2>;
using Blk2 = BlkCfg<64, 24, 32, 96, 32, 16, 8, 4, 4, true, 2>;
using Blk3 = BlkCfg<96, 24, 32, 96, 32, 16, 8, 4, 4, false, 2>;
using Blk4 = BlkCfg<96, 32, 32, 128, 16, 8, 16, 0, 1, true, 1>;
using Blk5 = BlkCfg<128, 32, 32, 128, 16, 8, 16, 0, 1, false, 1>;

template <class C>
int launch_block(const float *x, float *y, const unsigned char *w, int n, int *status, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        SSB_CHECK_CUDA(cudaFuncSetAttribute(osblock_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_B));
        attr = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n * C::NB);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = C::SMEM_B;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C::NB;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SSB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, osblock_tc_kernel<C>, x, y, w, n, status));
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

int ssb_reid_tc_block(int b, const float *x, float *y, const unsigned char *w, int n, int *status,
                      cudaStream_t st) {
    switch (b) {
        case 0: return launch_block<Blk0>(x, y, w, n, status, st);
        case 1: return launch_block<Blk1>(x, y, w, n, status, st);
        case 2: return launch_block<Blk2>(x, y, w, n, status, st);
        case 3: return launch_block<Blk3>(x, y, w, n, status, st);
        case 4: return launch_block<Blk4>(x, y, w, n, status, st);
        case 5: return launch_block<Blk5>(x, y, w, n, status, st);
    }
    ssb_set_error("bad OSBlock index %d", b);
    return -1;
}
