// OSNet OSBlock, fused, third tensor-core formulation ("planes + halo exchange"), sm_100a.
//
// Same math as reid_tc3.cu (pointwise convs on tcgen05 with hi/lo fp16 operand pairs, LightConv
// depthwise 3x3 on the CUDA cores in exact fp32, ChannelGate folded into conv3's weights), but the
// three things round 1's profile showed to cost the most are restructured:
//
//  1. Activations travel BETWEEN kernels as the tensor core's own operand: two fp16 planes (hi, lo with
//     hi + lo = value to ~2^-22) in the K-major no-swizzle layout  [crop][hl][C/8][H*W pixels][8 halves].
//     A band of rows is a contiguous slice of every (hl, chunk) plane, so the consumer pulls its whole
//     input with one cp.async.bulk per plane (TMA engine, mbarrier complete_tx) straight into the MMA
//     operand buffer -- no ld.global -> register -> split -> st.shared staging pass (18 K of an identity
//     block's 84 K cycles in reid_tc3), and the producer's epilogue stores are 512-byte coalesced runs
//     (lane = pixel) without the transposing shared-memory stage.
//  2. No recomputed halo.  A band owns R rows and nothing else; what the depthwise 3x3 needs from the
//     neighbouring bands -- one row of the POINTWISE result T on each side -- is pushed into their T
//     rings over distributed shared memory (st.shared::cluster) while the TMEM accumulator is drained,
//     and made visible by the hardware cluster barrier.  Stage 2 ran 24 rows per 16 useful (x1.5 of every
//     MMA, drain and staging byte), stage 3 16 per 8 (x2).
//  3. Small bands + 256-thread CTAs so that TWO CTAs (two independent dependency chains: MMA -> drain ->
//     barrier -> depthwise -> barrier -> MMA ...) are resident per SM where shared memory and TMEM
//     (<= 256 columns each) allow: the chain is latency-bound (round 1: warps active 25 %, tensor pipe
//     4-9 %), and a second CTA fills the bubbles of the first.
//
// Cluster protocol per LightConv layer (all threads of all bands execute every barrier instruction):
//     [wait A]   the neighbours have finished reading their T rings of the previous layer
//     drain      TMEM -> own rows of T; first / last own row also -> neighbour's bottom / top ring row
//     arrive B, wait B    pushes visible (also the CTA-wide barrier for T)
//     depthwise  reads T incl. ring rows, writes the next operand map P
//     arrive A   (non-blocking; paired with the next layer's [wait A])
// The last layer of a stream has no "arrive A": the ChannelGate's cluster.sync() plays that role.
#include <cooperative_groups.h>

#include "ssb_common.cuh"
#include "tc_common.cuh"

namespace cg = cooperative_groups;

#ifndef SSB_DW_CHAINS
#define SSB_DW_CHAINS 3          // depthwise FMA chains per output (3: one per window row; 1: a single 9-term chain)
#endif

namespace {

__host__ __device__ constexpr int cmax4(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int rup128_4(int a) { return (a + 127) / 128 * 128; }
__host__ __device__ constexpr int pow2cols(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

template <int CIN_, int MID_, int MIDP_, int COUT_, int H_, int W_, int R_, int NB_, bool DOWN_, int SEG_,
          int THREADS_, int MINB_, bool SPLIT_, bool PW_ = false>
struct B4 {
    // PW: the stage's transition layer (conv1x1 COUT -> COUT + ReLU + avgpool 2x2, the former pw_tc_kernel launch) is
    // fused behind the block: the block's output never goes to HBM -- the final epilogue writes it as an operand map
    // in shared memory with the tile rows in pooling-window order, one more tcgen05 GEMM runs on it, and the 2x2
    // average is a 2-step shuffle over adjacent TMEM lanes; what is stored is the pooled map's operand planes.
    static constexpr bool PW = PW_;
    static constexpr int CIN = CIN_, MID = MID_, MIDP = MIDP_, COUT = COUT_, H = H_, W = W_, R = R_, NB = NB_;
    static constexpr int SEG = SEG_, THREADS = THREADS_, MINB = MINB_;
    static constexpr bool DOWN = DOWN_, SPLIT = SPLIT_, EXCH = NB_ > 1;
    static constexpr int NWARPS = THREADS / 32, GROUPS = THREADS / 128;
    static constexpr int HW = H * W;
    static constexpr int NPX = R * W;                     // band pixels, p = lr * W + col
    static constexpr int NT = NPX / 128;                  // M tiles
    static constexpr int MCH = MIDP / 8;                  // 16-byte K chunks per pixel of a mid map
    static constexpr int CG = MID / 4;                    // real float4 channel groups
    static constexpr int XL = W < 16 ? W : 16;            // depthwise: column lanes
    static constexpr int CPW = 16 / XL;                   // channel-group pairs per warp
    static constexpr int NCB = W / XL;                    // column blocks
    static constexpr int NCGW = (CG / 2) / CPW;           // channel-group-pair groups
    static constexpr int TPS = NCB * NCGW;                // warp tasks per row segment
    static constexpr int PLANE_B = NPX * 16;              // one (hl, chunk) plane of the band
    static constexpr int MAP_HALF_B = MCH * PLANE_B, MAP_B = 2 * MAP_HALF_B;
    static constexpr int XCH = CIN / 8;
    static constexpr int XS_HALF_B = XCH * PLANE_B, XS_B = 2 * XS_HALF_B;       // input operand [hl][CIN/8][NPX][8]
    static constexpr int TW = W + 2, TH = R + 2, TPX = TW * TH;
    static constexpr int T_B = CG * TPX * 16;             // fp32 pointwise result with a one-pixel ring
    static constexpr int A_B = rup128_4(cmax4(MAP_B + T_B, XS_B));              // P map + T | input operand
    static constexpr int LCN = 2 * MIDP;                  // TMEM columns of a pointwise tile (hi | lo weights)
    static constexpr int TM_C3 = NT * LCN, TM_COLS = TM_C3 + NT * COUT, TM_ALLOC = pow2cols(TM_COLS);
    static constexpr int C1W_B = CIN * MIDP * 4;
    static constexpr int DNW_HALF_B = DOWN ? CIN * COUT * 2 : 0, DNW_B = 2 * DNW_HALF_B;
    static constexpr int LCW_B = MIDP * MIDP * 4;
    static constexpr int WALL_B = C1W_B + DNW_B + 10 * LCW_B;
    static constexpr int PWW_HALF_B = COUT * COUT * 2, PWW_B = 2 * PWW_HALF_B;   // transition weights [COUT/8][COUT][8] hi | lo
    static constexpr int C3W_HALF_B = MIDP * COUT * 2, C3W_B = 2 * C3W_HALF_B;
    // PAR (floats): B1[MIDP] | 10 x { DW[9][MIDP], B[MIDP] } | B3[COUT] | GW1[MIDP][2] | GB1[2] |
    //               GW2[2][MIDP] | GB2[MIDP]          (identical to reid_tc3's B3: same weight blob)
    static constexpr int P_B1 = 0, P_LC = MIDP, P_B3 = P_LC + 100 * MIDP, P_GW1 = P_B3 + COUT;
    static constexpr int P_GB1 = P_GW1 + 2 * MIDP, P_GW2 = P_GB1 + 2, P_GB2 = P_GW2 + 2 * MIDP;
    static constexpr int NPAR = P_GB2 + MIDP;
    static constexpr int OFF_X1 = 0;
    static constexpr int OFF_A = OFF_X1 + MAP_B;
    static constexpr int OFF_W = OFF_A + A_B;
    static constexpr int OFF_C3 = OFF_W + rup128_4(WALL_B);
    static constexpr int OFF_PAR = OFF_C3 + C3W_B;
    static constexpr int OFF_GAP = OFF_PAR + rup128_4(NPAR * 4);     // [4 streams][NB bands][MIDP] floats
    static constexpr int OFF_MISC = OFF_GAP + 4 * NB * MIDP * 4;
    static constexpr int SCR_FLOATS = SEG * NCB * CG * 4;
    static constexpr int SMEM_B = OFF_MISC + 192 + (SCR_FLOATS + 2 * MIDP) * 4 + 64;
    static constexpr int SMEM_LIMIT = MINB >= 2 ? (232448 / MINB - 1024) : 232448;
    static_assert(NPX % 128 == 0, "band = whole M tiles");
    static_assert((W & (W - 1)) == 0, "W power of two");
    static_assert(NT <= 8, "per-tile barriers");
    static_assert(TM_ALLOC * MINB <= 512, "TMEM columns of the co-resident CTAs");
    static_assert(SMEM_B <= SMEM_LIMIT, "shared memory");
    static_assert(MIDP % 16 == 0 && COUT % 32 == 0 && CIN % 16 == 0 && MID % 4 == 0, "MMA shapes");
    static_assert(H % R == 0 && H / R == NB && NB <= 8, "bands / portable cluster size");
    static_assert(MID <= MIDP && COUT == 4 * MID && MIDP <= 32 && COUT * (MIDP / 8) <= THREADS, "OSBlock channel plan");
    static_assert(CG % 2 == 0 && (CG / 2) % CPW == 0 && W % XL == 0, "depthwise warp tasks");
    static_assert(SEG * TPS <= NWARPS, "depthwise warp tasks fit the CTA");
    static_assert(WALL_B + XS_B < (1 << 20), "mbarrier tx count");
    static_assert(!PW || (!DOWN && CIN == COUT && R % 2 == 0 && (PWW_B <= MAP_B || PWW_B <= WALL_B)), "fused transition");
    // global weight blob sections (bytes): C1W | DNW | LCW[10] | PAR (fp32) | W3 (fp32 [MIDP][COUT])
    static constexpr int G_PAR = rup128_4(WALL_B);
    static constexpr int G_W3 = G_PAR + rup128_4(NPAR * 4);
    static constexpr int G_TOTAL = G_W3 + MIDP * COUT * 4;
};

__device__ __forceinline__ void split2(float a, float b, __half2 &h, __half2 &l) {
    h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    l = __floats2half2_rn(a - hf.x, b - hf.y);
}
// same split, the residual as ONE packed fp32x2 subtraction
__device__ __forceinline__ void split2p(float a, float b, __half2 &h, __half2 &l) {
    h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    unsigned long long v = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(a);
    const unsigned long long hv = ((unsigned long long)__float_as_uint(hf.y) << 32) | __float_as_uint(hf.x);
    asm("sub.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(hv));
    l = __floats2half2_rn(__uint_as_float((uint32_t)v), __uint_as_float((uint32_t)(v >> 32)));
}
__device__ __forceinline__ void split1(float v, __half &h, __half &l) {
    h = __float2half_rn(v);
    l = __float2half_rn(v - __half2float(h));
}
__device__ __forceinline__ uint64_t dadv(uint64_t base, int units16) {
    return base + (uint64_t)(int64_t)units16;
}
__device__ __forceinline__ void mma3(uint32_t d, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl,
                                     uint32_t idesc, uint32_t acc) {
    tc::mma_f16_ss(d, ah, bh, idesc, acc);
    tc::mma_f16_ss(d, al, bh, idesc, 1);
    tc::mma_f16_ss(d, ah, bl, idesc, 1);
}
__device__ __forceinline__ void tmem_ld4_nw(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8_nw(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// hardware cluster barrier, split phase (release / acquire at cluster scope)
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// phase "A" of the ring protocol publishes nothing (it only says "I have finished READING my ring", and those loads
// have been consumed by the arithmetic before it): no release fence -- the MEMBAR / ERRBAR pair behind a releasing
// arrive was ~14 % of the kernel's stall samples
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_f4(uint32_t cluster_addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}
// depthwise arithmetic on 4 channels: two packed fp32 pairs, Blackwell's FFMA2 (two IEEE fp32 FMAs per instruction)
struct P4 { unsigned long long a, b; };          // channels (0,1), (2,3)
__device__ __forceinline__ P4 ldp4(const float4 *p) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p);
    return P4{v.x, v.y};
}
__device__ __forceinline__ void fma4(P4 &o, const P4 &w, const P4 &v) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(o.a) : "l"(w.a), "l"(v.a));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(o.b) : "l"(w.b), "l"(v.b));
}
__device__ __forceinline__ void mul4(P4 &o, const P4 &w, const P4 &v) {
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(o.a) : "l"(w.a), "l"(v.a));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(o.b) : "l"(w.b), "l"(v.b));
}
__device__ __forceinline__ void add4(P4 &o, const P4 &v) {
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(o.a) : "l"(v.a));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(o.b) : "l"(v.b));
}
__device__ __forceinline__ float p4x(const P4 &v) { return __uint_as_float((uint32_t)v.a); }
__device__ __forceinline__ float p4y(const P4 &v) { return __uint_as_float((uint32_t)(v.a >> 32)); }
__device__ __forceinline__ float p4z(const P4 &v) { return __uint_as_float((uint32_t)v.b); }
__device__ __forceinline__ float p4w(const P4 &v) { return __uint_as_float((uint32_t)(v.b >> 32)); }

// value = hi + lo of a packed operand pair
__device__ __forceinline__ void unsplit8(const uint4 &h, const uint4 &l, float *out) {
    const __half2 *hh = reinterpret_cast<const __half2 *>(&h);
    const __half2 *ll = reinterpret_cast<const __half2 *>(&l);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float2 a = __half22float2(hh[j]), b = __half22float2(ll[j]);
        out[2 * j] = a.x + b.x;
        out[2 * j + 1] = a.y + b.y;
    }
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
osblock4_kernel(const unsigned char *__restrict__ x, unsigned char *__restrict__ y,
                const unsigned char *__restrict__ wblob, int n_crops, int *__restrict__ status,
                long long *__restrict__ dbg, const unsigned char *__restrict__ pwblob) {
    extern __shared__ __align__(1024) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, grp = warp >> 2;          // TMEM lane quadrant / work group
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);            // provably warp-uniform copy of warp
    const int crop = blockIdx.x / C::NB, band = blockIdx.x % C::NB;

    unsigned char *sX1 = smem + C::OFF_X1, *sP = smem + C::OFF_A;
    float4 *sT = reinterpret_cast<float4 *>(sP + C::MAP_B);
    unsigned char *sW = smem + C::OFF_W, *sC3 = smem + C::OFF_C3;
    float *sPar = reinterpret_cast<float *>(smem + C::OFF_PAR);
    float *sGap = reinterpret_cast<float *>(smem + C::OFF_GAP);      // [4][NB][MIDP] band-partial sums of every band
    uint64_t *bar_w = reinterpret_cast<uint64_t *>(smem + C::OFF_MISC);   // weights + input operand landed
    uint64_t *bar_c1 = bar_w + 1;                      // conv1 (+ downsample) MMAs
    uint64_t *bar_tile = bar_w + 3;                    // [8] per M tile of the current pointwise conv
    uint64_t *bar_c3 = bar_w + 11;                     // conv3 accumulation of the current stream
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bar_w + 16);
    float *s_scr = reinterpret_cast<float *>(smem + C::OFF_MISC + 192);   // [SEG][NCB][CG][4] partial sums

    if (warp == 0) tc::tmem_alloc(s_tmem, C::TM_ALLOC);
    if (tid == 0) {
        tc::mbar_init(bar_w, 1);
        tc::mbar_init(bar_c1, 1);
        for (int i = 0; i < 8; i++) tc::mbar_init(bar_tile + i, 1);
        tc::mbar_init(bar_c3, 1);
        tc::fence_mbar_init();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *s_tmem;
    int dbg_n = 0;
    auto stamp = [&]() { if (dbg && blockIdx.x == 0 && tid == 0 && dbg_n < 63) dbg[1 + dbg_n++] = clock64(); };
    // finer stamps of two layers (lc 1: a layer followed by another LightConv, lc 2: the last layer of a stream)
    // into dbg[64..]: [64] = count
    // (compiled in with -DSSB_FINE_STAMPS only: the extra branches cost the forward ~9 %)
#ifdef SSB_FINE_STAMPS
    int dbg_f = 0;
    auto fstamp = [&](int layer) {
        if (dbg && blockIdx.x == 0 && tid == 0 && (layer == 1 || layer == 2) && dbg_f < 40) { dbg[65 + dbg_f++] = clock64(); dbg[64] = dbg_f; }
    };
#else
    auto fstamp = [](int) {};
#endif
    stamp();
    bool ok = true;

    // ------------------------------------------------------------------
    // phase 1: every weight operand of the block and the band's whole input operand (one bulk copy per
    // (hl, 8-channel chunk) plane: the band is a contiguous NPX * 16 byte slice of each) arrive on one
    // mbarrier; then X1 = relu(conv1(x)) and the downsample conv, all tiles issued back to back
    // ------------------------------------------------------------------
    const unsigned char *xin = x + (size_t)crop * (4 * C::CIN * C::HW);          // [hl][CIN/8][HW][8] halves
    constexpr int X_LO = 2 * C::CIN * C::HW;                                       // byte offset of the lo planes
    // programmatic dependent launch: everything above (TMEM, barriers, parameters) and the weight copies below
    // run while the previous kernel of the stream is still draining; its output is only touched after the wait
    tc::pdl_launch_dependents();
    if (tid == 0) {
        // the fp32 parameters (biases, depthwise taps, gate MLP) ride the same mbarrier: a per-thread ld.global ->
        // st.shared copy of these 7-14 KB was 4.5 % of the kernel's stall samples (ncu source page, round 2)
        constexpr int PAR_B = (C::NPAR * 4 + 15) / 16 * 16;
        tc::mbar_arrive_expect_tx(bar_w, C::WALL_B + C::XS_B + PAR_B);
        for (int o = 0; o < C::WALL_B; o += 32768) {
            const int nb = C::WALL_B - o < 32768 ? C::WALL_B - o : 32768;
            tc::bulk_g2s(sW + o, wblob + o, nb, bar_w);
        }
        tc::bulk_g2s(sPar, wblob + C::G_PAR, PAR_B, bar_w);
        tc::pdl_wait();
        for (int hl = 0; hl < 2; hl++)
            for (int ch = 0; ch < C::XCH; ch++)
                tc::bulk_g2s(sP + hl * C::XS_HALF_B + ch * C::PLANE_B,
                             xin + hl * X_LO + (size_t)ch * C::HW * 16 + (size_t)band * C::PLANE_B, C::PLANE_B, bar_w);
    }

    constexpr uint32_t IDESC_MID = tc::make_idesc_f16(128, C::MIDP);
    constexpr uint32_t IDESC_CAT = tc::make_idesc_f16(128, 2 * C::MIDP);
    constexpr uint32_t IDESC_OUT = tc::make_idesc_f16(128, C::COUT);
    // MMAs are issued from warp-uniform code by an elected lane (~28 cycles per MMA + commit instead of ~90 from a
    // divergent single-thread branch, tc_common.cuh) of a warp without depthwise rows when there is one
    const bool issuer = warp_u == (C::SEG * C::TPS < C::NWARPS ? C::SEG * C::TPS : 0);

    if (!tc::mbar_wait(bar_w, 0)) ok = false;
    tc::fence_after_sync();
    if (issuer) {
        if (tc::elect_one()) {
            const uint64_t ah0 = tc::make_smem_desc(tc::smem_u32(sP), C::PLANE_B, 128);
            const uint64_t al0 = dadv(ah0, C::XS_HALF_B / 16);
            const uint64_t bc0 = tc::make_smem_desc(tc::smem_u32(sW), 2 * C::MIDP * 16, 128);
#pragma unroll
            for (int t = 0; t < C::NT; t++) {
                const uint32_t d1 = tmem + t * C::LCN;
#pragma unroll
                for (int ks = 0; ks < C::CIN / 16; ks++) {
                    const int ka = t * 128 + ks * 2 * C::NPX;
                    tc::mma_f16_ss(d1, dadv(ah0, ka), dadv(bc0, ks * 4 * C::MIDP), IDESC_CAT, ks > 0);
                    tc::mma_f16_ss(d1, dadv(al0, ka), dadv(bc0, ks * 4 * C::MIDP), IDESC_MID, 1);
                }
                if (C::DOWN) {
                    const uint64_t dh0 = tc::make_smem_desc(tc::smem_u32(sW) + C::C1W_B, C::COUT * 16, 128);
                    const uint64_t dl0 = dadv(dh0, C::DNW_HALF_B / 16);
                    const uint32_t d2 = tmem + C::TM_C3 + t * C::COUT;
#pragma unroll
                    for (int ks = 0; ks < C::CIN / 16; ks++) {
                        const int ka = t * 128 + ks * 2 * C::NPX;
                        mma3(d2, dadv(ah0, ka), dadv(al0, ka), dadv(dh0, ks * 2 * C::COUT),
                             dadv(dl0, ks * 2 * C::COUT), IDESC_OUT, ks > 0);
                    }
                }
            }
            tc::mma_commit(bar_c1);
        }
        __syncwarp();
    }
    if (!tc::mbar_wait(bar_c1, 0)) ok = false;
    tc::fence_after_sync();
    stamp();                                   // [1] phase 1 (input + weights landed, conv1/down MMAs) done

    // the input operand buffer becomes the P map (its pad-channel planes must read as zero) and T (zero ring)
    for (int i = tid; i < (C::MAP_B + C::T_B) / 16; i += C::THREADS)
        reinterpret_cast<uint4 *>(sP)[i] = make_uint4(0u, 0u, 0u, 0u);
    // X1 epilogue: TMEM tile -> (+bias, relu) -> hi/lo operand map; units (tile, 8-channel K chunk) spread
    // over the warp groups, all TMEM loads of a thread in flight before one wait
    {
        constexpr int NU1 = C::NT * C::MCH, UPT1 = (NU1 + C::GROUPS - 1) / C::GROUPS;
        uint32_t va[UPT1][8], vb[UPT1][8];
#pragma unroll
        for (int e = 0; e < UPT1; e++) {
            const int u = grp + e * C::GROUPS;
            if (u < NU1) {
                const int t = u / C::MCH, kc = u - t * C::MCH;
                const uint32_t ta = tmem + ((uint32_t)(quad * 32) << 16) + t * C::LCN + kc * 8;
                tmem_ld8_nw(ta, va[e]);
                tmem_ld8_nw(ta + C::MIDP, vb[e]);
            }
        }
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < UPT1; e++) {
            const int u = grp + e * C::GROUPS;
            if (u < NU1) {
                const int t = u / C::MCH, kc = u - t * C::MCH;
                const int p = t * 128 + quad * 32 + lane;
                const float *bias = sPar + C::P_B1 + kc * 8;
                __align__(16) __half2 h[4];
                __align__(16) __half2 l[4];
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const float f0 = fmaxf(__uint_as_float(va[e][j]) + __uint_as_float(vb[e][j]) + bias[j], 0.f);
                    const float f1 = fmaxf(__uint_as_float(va[e][j + 1]) + __uint_as_float(vb[e][j + 1]) + bias[j + 1], 0.f);
                    split2(f0, f1, h[j >> 1], l[j >> 1]);
                }
                unsigned char *d_hi = sX1 + kc * C::PLANE_B + p * 16;
                *reinterpret_cast<uint4 *>(d_hi) = *reinterpret_cast<uint4 *>(h);
                *reinterpret_cast<uint4 *>(d_hi + C::MAP_HALF_B) = *reinterpret_cast<uint4 *>(l);
            }
        }
    }
    // cluster barrier phase "A": my T ring is zeroed -- the neighbours may push into it
    bool a_pending = false;
    if (C::EXCH) { cluster_arrive(); a_pending = true; }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    stamp();                                   // [2] X1 drained

    // ------------------------------------------------------------------
    // phase 2: four streams of LightConvs (pointwise on tcgen05, depthwise on the CUDA cores)
    //          + gated conv3 accumulation
    // ------------------------------------------------------------------
    auto issue_pw = [&](const unsigned char *src, int layer, int t0, int t1) {      // issuer warp; tiles [t0, t1)
        const uint64_t ah0 = tc::make_smem_desc(tc::smem_u32(src), C::PLANE_B, 128);
        const uint64_t al0 = dadv(ah0, C::MAP_HALF_B / 16);
        const uint64_t b0 = tc::make_smem_desc(tc::smem_u32(sW) + C::C1W_B + C::DNW_B + layer * C::LCW_B,
                                               2 * C::MIDP * 16, 128);
#pragma unroll
        for (int t = 0; t < C::NT; t++) {
            if (t < t0 || t >= t1) continue;
            const uint32_t d = tmem + t * C::LCN;
            if (tc::elect_one()) {
#pragma unroll
                for (int ks = 0; ks < C::MIDP / 16; ks++) {
                    tc::mma_f16_ss(d, dadv(ah0, t * 128 + ks * 2 * C::NPX), dadv(b0, ks * 4 * C::MIDP), IDESC_CAT, ks > 0);
                    tc::mma_f16_ss(d, dadv(al0, t * 128 + ks * 2 * C::NPX), dadv(b0, ks * 4 * C::MIDP), IDESC_MID, 1);
                }
                tc::mma_commit(bar_tile + t);
            }
            __syncwarp();
        }
    };
    tc::pdl_wait();                           // every thread: before its own global reads / writes of activations
    float w3r[8];                             // conv3 weights of this thread's operand row (kc, co)
    {
        const float *w3 = reinterpret_cast<const float *>(wblob + C::G_W3);  // [MIDP][COUT]
        const int kc = tid / C::COUT, co = tid - kc * C::COUT;
#pragma unroll
        for (int j = 0; j < 8; j++) w3r[j] = tid < C::COUT * C::MCH ? w3[(kc * 8 + j) * C::COUT + co] : 0.f;
    }
    // depthwise task of this thread: lane = 2 * column + parity (+ 2 * XL * chunk pair): the two lanes of a pixel
    // are adjacent, so their 8-byte hi (and lo) halves of a 16-byte operand row are one contiguous store
    const int dw_par = lane & 1;
    const int dw_seg = warp / C::TPS, dw_cb = (warp % C::TPS) % C::NCB;
    const int dw_cg = (((warp % C::TPS) / C::NCB) * C::CPW + lane / (2 * C::XL)) * 2 + dw_par;
    const int dw_col = dw_cb * C::XL + (lane >> 1) % C::XL;
    // cluster-window addresses of the neighbours' T rings (identical shared-memory layout in every band)
    const uint32_t sT_up = (C::EXCH && band > 0) ? tc::mapa_u32(sT, band - 1) : 0u;
    const uint32_t sT_dn = (C::EXCH && band < C::NB - 1) ? tc::mapa_u32(sT, band + 1) : 0u;
    int lc = 0;
    uint32_t tile_par = 0, c3_par = 0;
    bool c3_pending = false;
    if (issuer) issue_pw(sX1, 0, 0, C::NT);
    for (int s = 0; s < 4; s++) {
        for (int k = 0; k <= s; k++, lc++) {
            const bool last = (k == s);
            // ---- this layer's depthwise taps into registers (rides under the wait for the pointwise MMAs)
            P4 wd[9], bs;
            {
                const float *wl = sPar + C::P_LC + lc * (10 * C::MIDP) + dw_cg * 4;
#pragma unroll
                for (int tap = 0; tap < 9; tap++) wd[tap] = ldp4(reinterpret_cast<const float4 *>(wl + tap * C::MIDP));
                bs = ldp4(reinterpret_cast<const float4 *>(wl + 9 * C::MIDP));
            }
            fstamp(lc);                        // f0: layer start (taps loaded)
            // ---- pointwise result: TMEM -> fp32 T own rows (+ edge rows into the neighbours' rings)
            {
                constexpr int NU = C::NT * C::CG, UPT = (NU + C::GROUPS - 1) / C::GROUPS;
                uint32_t ra_[UPT][4], rb_[UPT][4];
#pragma unroll
                for (int i = 0; i < UPT; i++) {
                    const int u = grp + i * C::GROUPS;
                    if (u >= NU) continue;                        // warp-uniform
                    const int t = u / C::CG, cgi = u - t * C::CG;
                    if (!tc::mbar_wait(bar_tile + t, tile_par)) ok = false;
                    tc::fence_after_sync();
                    const uint32_t ta = tmem + ((uint32_t)(quad * 32) << 16) + t * C::LCN + cgi * 4;
                    tmem_ld4_nw(ta, ra_[i]);
                    tmem_ld4_nw(ta + C::MIDP, rb_[i]);
                }
                tmem_wait_ld();
                fstamp(lc);                    // f1: MMAs complete, TMEM loaded
                if (a_pending) { cluster_wait(); a_pending = false; }       // [wait A] rings are free
                fstamp(lc);                    // f2: wait A done
#pragma unroll
                for (int i = 0; i < UPT; i++) {
                    const int u = grp + i * C::GROUPS;
                    if (u >= NU) continue;
                    const int t = u / C::CG, cgi = u - t * C::CG;
                    const int p = t * 128 + quad * 32 + lane;
                    const int lr = p / C::W, col = p % C::W;
                    const float v0 = __uint_as_float(ra_[i][0]) + __uint_as_float(rb_[i][0]);
                    const float v1 = __uint_as_float(ra_[i][1]) + __uint_as_float(rb_[i][1]);
                    const float v2 = __uint_as_float(ra_[i][2]) + __uint_as_float(rb_[i][2]);
                    const float v3 = __uint_as_float(ra_[i][3]) + __uint_as_float(rb_[i][3]);
                    sT[cgi * C::TPX + (lr + 1) * C::TW + col + 1] = make_float4(v0, v1, v2, v3);
                    if (C::EXCH) {
                        if (lr == 0 && sT_up)          // my first row = bottom ring row of the band above
                            st_cluster_f4(sT_up + (uint32_t)(cgi * C::TPX + (C::R + 1) * C::TW + col + 1) * 16u, v0, v1, v2, v3);
                        if (lr == C::R - 1 && sT_dn)   // my last row = top ring row of the band below
                            st_cluster_f4(sT_dn + (uint32_t)(cgi * C::TPX + col + 1) * 16u, v0, v1, v2, v3);
                    }
                }
            }
            tile_par ^= 1;
            fstamp(lc);                        // f3: T written, edges pushed
            tc::fence_before_sync();
            if (C::EXCH) { cluster_arrive(); cluster_wait(); }             // [B] pushes visible; CTA-wide barrier too
            else __syncthreads();
            tc::fence_after_sync();
            fstamp(lc);                        // f4: barrier B done
            stamp();                           // T ready
            if (C::PW && lc == 9 && tid == 0) {
                // every pointwise MMA has completed: X1 and the LightConv weights are dead -- the transition weights
                // land in whichever of the two regions holds them, long before the final epilogue needs them
                unsigned char *dstw = C::PWW_B <= C::MAP_B ? sX1 : sW;
                tc::mbar_arrive_expect_tx(bar_w, C::PWW_B);
                for (int o = 0; o < C::PWW_B; o += 32768) {
                    const int nb = C::PWW_B - o < 32768 ? C::PWW_B - o : 32768;
                    tc::bulk_g2s(dstw + o, pwblob + o, nb, bar_w);
                }
            }
            // the next stream starts from X1: its pointwise conv runs under this depthwise pass
            if (last && s < 3 && issuer) issue_pw(sX1, lc + 1, 0, C::NT);
            // the previous stream's conv3 MMAs read P: done before this stream overwrites it
            if (k == 0 && c3_pending) {
                if (!tc::mbar_wait(bar_c3, c3_par)) ok = false;
                c3_par ^= 1;
                c3_pending = false;
                tc::fence_after_sync();
            }
            // ---- depthwise 3x3 + bias + ReLU -> hi/lo operand map P (in place: the pointwise MMAs that
            //      read P have completed)
            float4 gacc = make_float4(0.f, 0.f, 0.f, 0.f);
            auto dw_rows = [&](int ra, int rb) {
                if (rb <= ra) return;
                if (dw_seg >= C::SEG) return;                        // warp-uniform
                const int r0 = ra + ((rb - ra) * dw_seg) / C::SEG, r1 = ra + ((rb - ra) * (dw_seg + 1)) / C::SEG;
                if (r0 >= r1) return;
                const float4 *Tp = sT + dw_cg * C::TPX + dw_col;         // window columns col-1 .. col+1 (ring offset 1)
                unsigned char *dbase = sP + (dw_cg >> 1) * C::PLANE_B + dw_col * 16 + dw_par * 8;
                auto ldrow = [&](P4 *w, int trow) {
#pragma unroll
                    for (int j = 0; j < 3; j++) w[j] = ldp4(Tp + trow * C::TW + j);
                };
                auto dwrow = [&](int lr, const P4 *wa, const P4 *wb, const P4 *wc) {
                    // three independent 3-term chains (one per window row): measured on B200 the pass is bound by
                    // dependent-issue latency, not by instruction count -- one 9-term chain per packed pair (4 fewer
                    // instructions per row) ran the whole forward 4.5 % SLOWER (profiles/r02_reid_variants.md)
#if SSB_DW_CHAINS == 3
                    P4 o = bs, o1, o2;
                    fma4(o, wd[0], wa[0]);
                    mul4(o1, wd[3], wb[0]);
                    mul4(o2, wd[6], wc[0]);
                    fma4(o, wd[1], wa[1]);
                    fma4(o1, wd[4], wb[1]);
                    fma4(o2, wd[7], wc[1]);
                    fma4(o, wd[2], wa[2]);
                    fma4(o1, wd[5], wb[2]);
                    fma4(o2, wd[8], wc[2]);
                    add4(o1, o2);
                    add4(o, o1);
#else
                    P4 o = bs;
#pragma unroll
                    for (int dx = 0; dx < 3; dx++) {
                        fma4(o, wd[dx], wa[dx]);
                        fma4(o, wd[3 + dx], wb[dx]);
                        fma4(o, wd[6 + dx], wc[dx]);
                    }
#endif
                    const float ox = fmaxf(p4x(o), 0.f), oy = fmaxf(p4y(o), 0.f);
                    const float oz = fmaxf(p4z(o), 0.f), ow = fmaxf(p4w(o), 0.f);
                    if (last) { gacc.x += ox; gacc.y += oy; gacc.z += oz; gacc.w += ow; }
                    __half2 h[2], l[2];
                    split2p(ox, oy, h[0], l[0]);
                    split2p(oz, ow, h[1], l[1]);
                    unsigned char *d = dbase + lr * (C::W * 16);
                    *reinterpret_cast<uint2 *>(d) = make_uint2(*reinterpret_cast<uint32_t *>(&h[0]), *reinterpret_cast<uint32_t *>(&h[1]));
                    *reinterpret_cast<uint2 *>(d + C::MAP_HALF_B) =
                        make_uint2(*reinterpret_cast<uint32_t *>(&l[0]), *reinterpret_cast<uint32_t *>(&l[1]));
                };
                P4 w0[3], w1[3], w2[3];                            // rotating 3-row window (no register moves)
                ldrow(w0, r0);
                ldrow(w1, r0 + 1);
                for (int lr = r0; lr < r1; lr += 3) {
                    ldrow(w2, lr + 2);
                    dwrow(lr, w0, w1, w2);
                    if (lr + 1 < r1) {
                        ldrow(w0, lr + 3);
                        dwrow(lr + 1, w1, w2, w0);
                    }
                    if (lr + 2 < r1) {
                        ldrow(w1, lr + 4);
                        dwrow(lr + 2, w2, w0, w1);
                    }
                }
            };
            auto publish = [&]() {             // operand map writes -> visible to the tensor core, CTA-wide
                tc::fence_async_smem();
                tc::fence_before_sync();
                __syncthreads();
                tc::fence_after_sync();
            };
            if (!last) {
                if (C::SPLIT && C::NT >= 2) {
                    // the next pointwise conv is per pixel: its MMAs on the upper half of the tiles are issued as
                    // soon as the upper rows are written and run under the lower half's depthwise pass
                    constexpr int TH_ = C::NT / 2, HR = TH_ * 128 / C::W;
                    dw_rows(0, HR);
                    publish();
                    if (issuer) issue_pw(sP, lc + 1, 0, TH_);
                    dw_rows(HR, C::R);
                    if (C::EXCH) { cluster_arrive_relaxed(); a_pending = true; }    // [arrive A] done reading my ring
                    publish();
                    if (issuer) issue_pw(sP, lc + 1, TH_, C::NT);
                } else {
                    dw_rows(0, C::R);
                    fstamp(lc);                // f5: depthwise rows done (this warp)
                    if (C::EXCH) { cluster_arrive_relaxed(); a_pending = true; }    // [arrive A]
                    publish();
                    fstamp(lc);                // f6: published
                    if (issuer) issue_pw(sP, lc + 1, 0, C::NT);
                    fstamp(lc);                // f7: next pointwise issued
                }
                stamp();                       // depthwise done, next pointwise issued
                continue;
            }
            dw_rows(0, C::R);
            fstamp(lc);                        // f5: depthwise rows done (this warp)
            {                                  // the column lanes of a (segment, column block, channel group) are adjacent
#pragma unroll
                for (int off = C::XL; off >= 2; off >>= 1) {
                    gacc.x += __shfl_xor_sync(0xffffffffu, gacc.x, off);
                    gacc.y += __shfl_xor_sync(0xffffffffu, gacc.y, off);
                    gacc.z += __shfl_xor_sync(0xffffffffu, gacc.z, off);
                    gacc.w += __shfl_xor_sync(0xffffffffu, gacc.w, off);
                }
                if (dw_seg < C::SEG && ((lane >> 1) % C::XL) == 0)
                    *reinterpret_cast<float4 *>(s_scr + ((dw_seg * C::NCB + dw_cb) * C::CG + dw_cg) * 4) = gacc;
            }
            publish();
            fstamp(lc);                        // f6: published
            stamp();                           // depthwise done
            // ---- ChannelGate: band-partial sums -> every band's shared memory (pushed), cluster barrier,
            //      mean -> MLP -> sigmoid -> gate-scaled conv3 weights (computed redundantly, identically, per band)
            if (warp == 0 && lane < C::MIDP) {
                float tot = 0.f;
                if (lane < C::MID)
                    for (int g = 0; g < C::SEG * C::NCB; g++) tot += s_scr[(g * C::CG + (lane >> 2)) * 4 + (lane & 3)];
                float *slot = sGap + (s * C::NB + band) * C::MIDP + lane;
                if (C::NB > 1) {
#pragma unroll
                    for (int b = 0; b < C::NB; b++) tc::st_cluster_f32(tc::mapa_u32(slot, b), tot);
                } else {
                    *slot = tot;
                }
            }
            fstamp(lc);                        // f7: gate sums pushed (warp 0)
            if (C::NB > 1) cluster.sync(); else __syncthreads();      // also phase "A" of the ring protocol
            fstamp(lc);                        // f8: gate barrier
            if (warp * 32 < C::COUT * C::MCH) {      // warps that own conv3 weight rows; lane c = channel c
                float tot = 0.f;                   // fixed band order: every CTA of the crop gets the same bits
                if (lane < C::MIDP)
                    for (int b = 0; b < C::NB; b++) tot += sGap[(s * C::NB + b) * C::MIDP + lane];
                const float m = tot / (float)(C::H * C::W);
                float h0 = lane < C::MIDP ? m * sPar[C::P_GW1 + lane * 2 + 0] : 0.f;
                float h1 = lane < C::MIDP ? m * sPar[C::P_GW1 + lane * 2 + 1] : 0.f;
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    h0 += __shfl_xor_sync(0xffffffffu, h0, off);
                    h1 += __shfl_xor_sync(0xffffffffu, h1, off);
                }
                h0 = fmaxf(h0 + sPar[C::P_GB1 + 0], 0.f);
                h1 = fmaxf(h1 + sPar[C::P_GB1 + 1], 0.f);
                float gate = 0.f;
                if (lane < C::MIDP) {
                    float g = sPar[C::P_GB2 + lane];
                    g = fmaf(h0, sPar[C::P_GW2 + lane], g);
                    g = fmaf(h1, sPar[C::P_GW2 + C::MIDP + lane], g);
                    gate = 1.f / (1.f + expf(-g));
                }
                const int u = tid;
                const int kc = u / C::COUT;
                __align__(16) __half h[8];
                __align__(16) __half l[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float gk = __shfl_sync(0xffffffffu, gate, (kc * 8 + j) & 31);
                    split1(w3r[j] * gk, h[j], l[j]);
                }
                if (u < C::COUT * C::MCH) {
                    *reinterpret_cast<uint4 *>(sC3 + (size_t)u * 16) = *reinterpret_cast<uint4 *>(h);
                    *reinterpret_cast<uint4 *>(sC3 + C::C3W_HALF_B + (size_t)u * 16) = *reinterpret_cast<uint4 *>(l);
                }
            }
            fstamp(lc);                        // f9: gate MLP + scaled conv3 weights written
            tc::fence_async_smem();
            tc::fence_before_sync();
            __syncthreads();
            tc::fence_after_sync();
            if (issuer) {
                const uint64_t ah0 = tc::make_smem_desc(tc::smem_u32(sP), C::PLANE_B, 128);
                const uint64_t al0 = dadv(ah0, C::MAP_HALF_B / 16);
                const uint64_t bh0 = tc::make_smem_desc(tc::smem_u32(sC3), C::COUT * 16, 128);
                const uint64_t bl0 = dadv(bh0, C::C3W_HALF_B / 16);
                const uint32_t acc0 = (C::DOWN || s > 0) ? 1u : 0u;
                if (tc::elect_one()) {
#pragma unroll
                    for (int i = 0; i < C::NT; i++) {
                        const uint32_t d = tmem + C::TM_C3 + i * C::COUT;
#pragma unroll
                        for (int ks = 0; ks < C::MIDP / 16; ks++) {
                            const int ka = i * 128 + ks * 2 * C::NPX, kb = ks * 2 * C::COUT;
                            mma3(d, dadv(ah0, ka), dadv(al0, ka), dadv(bh0, kb), dadv(bl0, kb), IDESC_OUT,
                                 ks > 0 ? 1u : acc0);
                        }
                    }
                    tc::mma_commit(bar_c3);
                }
                __syncwarp();
            }
            c3_pending = true;
            stamp();                           // gate + conv3 issued
        }
    }

    // ------------------------------------------------------------------
    // final epilogue: y = relu(conv3 + bias (+ downsample already in TMEM) (+ x)) written as hi/lo operand
    // planes: lane = pixel, so every (hl, chunk) store of a warp is one 512-byte run.  The residual of the
    // identity blocks is re-read from the input planes (L2) before the wait for the last conv3 MMAs.
    // ------------------------------------------------------------------
    unsigned char *yout = y + (size_t)crop * (4 * C::COUT * C::HW);
    constexpr int Y_LO = 2 * C::COUT * C::HW;
    constexpr int CCH = C::COUT / 32;                         // 32-column chunks per tile
    constexpr int NUE = C::NT * CCH, UE = (NUE + C::GROUPS - 1) / C::GROUPS;
    constexpr bool PREF = !C::DOWN && UE <= 2;                 // residual of every unit prefetched up front
    uint4 xh[PREF ? UE : 1][4], xl[PREF ? UE : 1][4];
    auto load_res = [&](int u, uint4 *h, uint4 *l) {
        const int i = u / CCH, c0 = (u - i * CCH) * 32;
        const size_t gp = (size_t)band * C::NPX + i * 128 + quad * 32 + lane;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const unsigned char *src = xin + (size_t)(c0 / 8 + j) * C::HW * 16 + gp * 16;
            h[j] = *reinterpret_cast<const uint4 *>(src);
            l[j] = *reinterpret_cast<const uint4 *>(src + X_LO);
        }
    };
    if (PREF) {
#pragma unroll
        for (int e = 0; e < UE; e++) {
            const int u = grp + e * C::GROUPS;
            if (u < NUE) load_res(u, xh[PREF ? e : 0], xl[PREF ? e : 0]);
        }
    }
    if (c3_pending) {                 // the last stream's conv3 MMAs
        if (!tc::mbar_wait(bar_c3, c3_par)) ok = false;
        c3_par ^= 1;
    }
    tc::fence_after_sync();
#pragma unroll
    for (int e = 0; e < UE; e++) {
        const int u = grp + e * C::GROUPS;
        if (u >= NUE) continue;                               // warp-uniform
        const int i = u / CCH, c0 = (u - i * CCH) * 32;
        const size_t gp = (size_t)band * C::NPX + i * 128 + quad * 32 + lane;
        if (!C::DOWN && !PREF) load_res(u, xh[0], xl[0]);
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + C::TM_C3 + i * C::COUT + c0, v);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float r[8];
            if (!C::DOWN) unsplit8(xh[PREF ? e : 0][j], xl[PREF ? e : 0][j], r);
            __align__(16) __half2 h[4];
            __align__(16) __half2 l[4];
#pragma unroll
            for (int q = 0; q < 8; q += 2) {
                float f0 = v[j * 8 + q] + sPar[C::P_B3 + c0 + j * 8 + q];
                float f1 = v[j * 8 + q + 1] + sPar[C::P_B3 + c0 + j * 8 + q + 1];
                if (!C::DOWN) { f0 += r[q]; f1 += r[q + 1]; }
                split2(fmaxf(f0, 0.f), fmaxf(f1, 0.f), h[q >> 1], l[q >> 1]);
            }
            if (!C::PW) {
                unsigned char *dst = yout + (size_t)(c0 / 8 + j) * C::HW * 16 + gp * 16;
                *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<uint4 *>(h);
                *reinterpret_cast<uint4 *>(dst + Y_LO) = *reinterpret_cast<uint4 *>(l);
            } else {
                // operand map of the transition GEMM (the dead input-operand buffer): pixel (lr, col) goes to tile row
                // (pooled pixel) * 4 + (window position), so the four members of a 2x2 window are adjacent TMEM lanes
                const int p = i * 128 + quad * 32 + lane, lr = p / C::W, col = p % C::W;
                const int mrow = (((lr >> 1) * (C::W / 2) + (col >> 1)) << 2) + ((lr & 1) << 1) + (col & 1);
                unsigned char *dst = sP + (c0 / 8 + j) * C::PLANE_B + mrow * 16;
                *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<uint4 *>(h);
                *reinterpret_cast<uint4 *>(dst + C::XS_HALF_B) = *reinterpret_cast<uint4 *>(l);
            }
        }
    }
    if (C::PW) {
        // ---- fused transition: relu(conv1x1(y) + b), 2x2 average, stored as the pooled map's operand planes
        tc::fence_async_smem();
        if (!tc::mbar_wait(bar_w, 1)) ok = false;              // transition weights landed (issued after layer 9's drain)
        tc::fence_before_sync();
        __syncthreads();                                       // y map complete; conv3 accumulators fully drained
        tc::fence_after_sync();
        if (issuer) {
            if (tc::elect_one()) {
                const unsigned char *wsm = C::PWW_B <= C::MAP_B ? sX1 : sW;
                const uint64_t ah0 = tc::make_smem_desc(tc::smem_u32(sP), C::PLANE_B, 128);
                const uint64_t al0 = dadv(ah0, C::XS_HALF_B / 16);
                const uint64_t bh0 = tc::make_smem_desc(tc::smem_u32(wsm), C::COUT * 16, 128);
                const uint64_t bl0 = dadv(bh0, C::PWW_HALF_B / 16);
#pragma unroll
                for (int t = 0; t < C::NT; t++) {
                    const uint32_t d = tmem + C::TM_C3 + t * C::COUT;          // the drained conv3 accumulator columns
#pragma unroll
                    for (int ks = 0; ks < C::COUT / 16; ks++) {
                        const int ka = t * 128 + ks * 2 * C::NPX, kb = ks * 2 * C::COUT;
                        mma3(d, dadv(ah0, ka), dadv(al0, ka), dadv(bh0, kb), dadv(bl0, kb), IDESC_OUT, ks > 0);
                    }
                }
                tc::mma_commit(bar_c1);
            }
            __syncwarp();
        }
        if (!tc::mbar_wait(bar_c1, 1)) ok = false;
        tc::fence_after_sync();
        constexpr int HWO = C::HW / 4;                          // pooled map
        unsigned char *pout = y + (size_t)crop * (4 * C::COUT * HWO);
        const float *pwb = reinterpret_cast<const float *>(pwblob + C::PWW_B);
        constexpr int NU2 = C::NT * (C::COUT / 16), U2 = (NU2 + C::GROUPS - 1) / C::GROUPS;
#pragma unroll 1
        for (int e = 0; e < U2; e++) {
            const int u = grp + e * C::GROUPS;
            if (u >= NU2) continue;                             // warp-uniform
            const int t = u / (C::COUT / 16), c0 = (u - t * (C::COUT / 16)) * 16;
            float v[16];
            tc::tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + C::TM_C3 + t * C::COUT + c0, v);
            const int m = quad * 32 + lane;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                float f = fmaxf(v[j] + pwb[c0 + j], 0.f);
                f += __shfl_xor_sync(0xffffffffu, f, 1);
                f += __shfl_xor_sync(0xffffffffu, f, 2);
                v[j] = f * 0.25f;
            }
            if ((m & 3) == 0) {
                const size_t gpo = (size_t)band * (C::NPX / 4) + t * 32 + (m >> 2);
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    __align__(16) __half2 h[4];
                    __align__(16) __half2 l[4];
#pragma unroll
                    for (int q = 0; q < 8; q += 2) split2(v[j * 8 + q], v[j * 8 + q + 1], h[q >> 1], l[q >> 1]);
                    unsigned char *dst = pout + (size_t)(c0 / 8 + j) * HWO * 16 + gpo * 16;
                    *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<uint4 *>(h);
                    *reinterpret_cast<uint4 *>(dst + 2 * C::COUT * HWO) = *reinterpret_cast<uint4 *>(l);
                }
            }
        }
    }
    stamp();                                   // final epilogue done
    if (dbg && blockIdx.x == 0 && tid == 0) dbg[0] = dbg_n;
    if (!ok) { if (tid == 0) atomicExch(status, 6); }
    tc::fence_before_sync();
    if (C::NB > 1) cluster.sync(); else __syncthreads();     // no band exits while its peers may still push into it
    if (warp == 0) tc::tmem_dealloc(tmem, C::TM_ALLOC);
    (void)n_crops;
}

// ---------------------------------------------------------------------------
// the six OSBlocks of osnet_x0_25 (stage 2: 64x32, stage 3: 32x16, stage 4: 16x8)
// ---------------------------------------------------------------------------
// Band geometry is a compile-time choice (tools/build_variants.py A/Bs them on the GPU):
//   stage 2: SSB_S2_R = 8  -> 8 bands x 256 px, 256 threads, two CTAs per SM (default)
//            SSB_S2_R = 16 -> 4 bands x 512 px, 512 threads, one CTA per SM
//   stage 3: SSB_S3_R = 16 -> 2 bands x 256 px (default);  8 -> 4 bands x 128 px
#ifndef SSB_S2_R
#define SSB_S2_R 8
#endif
#ifndef SSB_DW_CHAINS
#define SSB_DW_CHAINS 3
#endif
#ifndef SSB_S2_SPLIT
#define SSB_S2_SPLIT false
#endif
#ifndef SSB_S3_R
#define SSB_S3_R 16
#endif
#ifndef SSB_S3_SPLIT
#define SSB_S3_SPLIT false
#endif
constexpr int S2_THREADS = SSB_S2_R == 8 ? 256 : 512, S2_MINB = SSB_S2_R == 8 ? 2 : 1, S2_SEG = SSB_S2_R == 8 ? 2 : 3;
//            CIN MID MIDP COUT  H   W   R  NB  DOWN SEG THREADS MINB SPLIT
using K0 = B4<16, 16, 16, 64, 64, 32, SSB_S2_R, 64 / SSB_S2_R, true, S2_SEG, S2_THREADS, S2_MINB, SSB_S2_SPLIT>;
using K1 = B4<64, 16, 16, 64, 64, 32, SSB_S2_R, 64 / SSB_S2_R, false, S2_SEG, S2_THREADS, S2_MINB, SSB_S2_SPLIT>;
using K2 = B4<64, 24, 32, 96, 32, 16, SSB_S3_R, 32 / SSB_S3_R, true, 4, 512, 1, SSB_S3_SPLIT>;
using K3 = B4<96, 24, 32, 96, 32, 16, SSB_S3_R, 32 / SSB_S3_R, false, 4, 512, 1, SSB_S3_SPLIT>;
// blocks 1 and 3 with their stage's transition layer fused behind them (the forward's default)
using K1F = B4<64, 16, 16, 64, 64, 32, SSB_S2_R, 64 / SSB_S2_R, false, S2_SEG, S2_THREADS, S2_MINB, SSB_S2_SPLIT, true>;
using K3F = B4<96, 24, 32, 96, 32, 16, SSB_S3_R, 32 / SSB_S3_R, false, 4, 512, 1, SSB_S3_SPLIT, true>;
using K4 = B4<96, 32, 32, 128, 16, 8, 16, 1, true, 7, 512, 1, false>;
using K5 = B4<128, 32, 32, 128, 16, 8, 16, 1, false, 7, 512, 1, false>;

template <class C>
int launch4(const unsigned char *x, unsigned char *y, const unsigned char *w, int n, int *status, long long *dbg,
            cudaStream_t st, const unsigned char *pw = nullptr) {
    static const int key = ssb_new_key();
    if (ssb_first_on_device(key))
        SSB_CHECK_CUDA(cudaFuncSetAttribute(osblock4_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_B));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n * C::NB);
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_B;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C::NB;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = ssb_pdl_enabled() ? 2 : 1;
    SSB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, osblock4_kernel<C>, x, y, w, n, status, dbg, pw));
    g_ssb_launches++;
    return 0;
}

// ---------------------------------------------------------------------------
// layout converters (parity tests and the fp32 entry points): float32 NHWC <-> hi/lo operand planes
// ---------------------------------------------------------------------------
__global__ void nhwc_to_planes_kernel(const float *__restrict__ x, unsigned char *__restrict__ y, int n, int HW, int Cc) {
    const long long total = (long long)n * HW * (Cc / 8);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int px = (int)(i % HW);
        const int ch = (int)((i / HW) % (Cc / 8));
        const int crop = (int)(i / ((long long)HW * (Cc / 8)));
        const float *src = x + ((size_t)crop * HW + px) * Cc + ch * 8;
        __align__(16) __half2 h[4];
        __align__(16) __half2 l[4];
#pragma unroll
        for (int j = 0; j < 4; j++) split2(src[2 * j], src[2 * j + 1], h[j], l[j]);
        unsigned char *dst = y + (size_t)crop * (4 * Cc * HW) + (size_t)ch * HW * 16 + (size_t)px * 16;
        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<uint4 *>(h);
        *reinterpret_cast<uint4 *>(dst + (size_t)2 * Cc * HW) = *reinterpret_cast<uint4 *>(l);
    }
}
__global__ void planes_to_nhwc_kernel(const unsigned char *__restrict__ x, float *__restrict__ y, int n, int HW, int Cc) {
    const long long total = (long long)n * HW * (Cc / 8);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int px = (int)(i % HW);
        const int ch = (int)((i / HW) % (Cc / 8));
        const int crop = (int)(i / ((long long)HW * (Cc / 8)));
        const unsigned char *src = x + (size_t)crop * (4 * Cc * HW) + (size_t)ch * HW * 16 + (size_t)px * 16;
        const uint4 h = *reinterpret_cast<const uint4 *>(src);
        const uint4 l = *reinterpret_cast<const uint4 *>(src + (size_t)2 * Cc * HW);
        float r[8];
        unsplit8(h, l, r);
        float *dst = y + ((size_t)crop * HW + px) * Cc + ch * 8;
#pragma unroll
        for (int j = 0; j < 8; j++) dst[j] = r[j];
    }
}

}  // namespace

int64_t ssb_reid_tc4_block_bytes(int b) {
    switch (b) {
        case 0: return K0::G_TOTAL;
        case 1: return K1::G_TOTAL;
        case 2: return K2::G_TOTAL;
        case 3: return K3::G_TOTAL;
        case 4: return K4::G_TOTAL;
        case 5: return K5::G_TOTAL;
    }
    return -1;
}

extern long long *g_ssb_tc_dbg;

// x, y: hi/lo operand planes [n][hl][C/8][H*W][8 halves] (4 * C * H * W bytes per crop)
int ssb_reid_tc4_block(int b, const void *x, void *y, const unsigned char *w, int n, int *status, cudaStream_t st) {
    long long *dbg = g_ssb_tc_dbg;
    const unsigned char *xi = (const unsigned char *)x;
    unsigned char *yo = (unsigned char *)y;
    switch (b) {
        case 0: return launch4<K0>(xi, yo, w, n, status, dbg, st);
        case 1: return launch4<K1>(xi, yo, w, n, status, dbg, st);
        case 2: return launch4<K2>(xi, yo, w, n, status, dbg, st);
        case 3: return launch4<K3>(xi, yo, w, n, status, dbg, st);
        case 4: return launch4<K4>(xi, yo, w, n, status, dbg, st);
        case 5: return launch4<K5>(xi, yo, w, n, status, dbg, st);
    }
    ssb_set_error("bad OSBlock index %d", b);
    return -1;
}

// block b (1 or 3) + the transition layer behind it in one launch: y = the POOLED map's operand planes;
// pw = the transition's weight section (reid_tc.cu PwCfg blob: hi | lo weights, then the bias)
int ssb_reid_tc4_block_pw(int b, const void *x, void *y, const unsigned char *w, const unsigned char *pw, int n, int *status,
                          cudaStream_t st) {
    long long *dbg = g_ssb_tc_dbg;
    const unsigned char *xi = (const unsigned char *)x;
    unsigned char *yo = (unsigned char *)y;
    if (b == 1) return launch4<K1F>(xi, yo, w, n, status, dbg, st, pw);
    if (b == 3) return launch4<K3F>(xi, yo, w, n, status, dbg, st, pw);
    ssb_set_error("only OSBlocks 1 and 3 are followed by a transition layer (got %d)", b);
    return -1;
}

int ssb_reid_nhwc_to_planes(const float *x, void *y, int n, int hw, int c, cudaStream_t st) {
    if (n <= 0) return 0;
    const long long total = (long long)n * hw * (c / 8);
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    nhwc_to_planes_kernel<<<blocks, 256, 0, st>>>(x, (unsigned char *)y, n, hw, c);
    SSB_CHECK_LAUNCH();
    return 0;
}
int ssb_reid_planes_to_nhwc(const void *x, float *y, int n, int hw, int c, cudaStream_t st) {
    if (n <= 0) return 0;
    const long long total = (long long)n * hw * (c / 8);
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    planes_to_nhwc_kernel<<<blocks, 256, 0, st>>>((const unsigned char *)x, y, n, hw, c);
    SSB_CHECK_LAUNCH();
    return 0;
}
