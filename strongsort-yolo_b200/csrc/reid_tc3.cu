// OSNet OSBlock, fused, second tensor-core formulation ("pwdw"), sm_100a.
//
// reid_tc.cu runs a LightConv3x3 (1x1 conv, then depthwise 3x3) as ONE dense 3x3 conv = 9
// shifted tcgen05 GEMMs.  Measured (profiles/r01_tc_phases.md): every M=128, K=16 MMA costs
// ~40 cycles whatever N is, because it is bound by the 4 KB shared-memory read of its A tile
// -- so the 9 taps x (hi, lo) operand reads made the LightConv layers 55 % of the kernel.
// Here the LightConv is split the way it is defined:
//   * the pointwise 1x1 (the only real contraction) stays on the tensor cores: ONE tap,
//     2 A-tile reads per tile and K step (hi/lo operands, N-concatenated weights);
//   * the depthwise 3x3 runs on the CUDA cores in exact fp32 on the GEMM result: the TMEM
//     accumulator is drained to a zero-ringed fp32 map T[c/4][row][col][4] in shared memory,
//     a sliding 3-row register window per thread (1 column x 4 channels; lanes = adjacent
//     columns, so every shared-memory access is conflict-free) makes 3 float4 loads per output
//     float4, and the result (bias, ReLU, hi/lo split) is written straight into the next
//     layer's K-major operand map -- the two lanes holding the halves of an 8-channel K chunk
//     swap hi/lo words by shuffle so that each writes one full 16-byte operand row.
// Because no operand is shifted any more the maps need no zero ring / guard pixels:
// pixel p = row * W + col, a band is a whole number of 128-row M tiles, the band's own
// rows start on a tile boundary in stage 2, and all 10 LightConv weight sets (1-4 KB each
// now) stay resident in shared memory after one bulk copy.
//
// Everything else follows reid_tc.cu: band of R rows (+HALO recomputed rows each side) of one
// crop per CTA, the bands of a crop form a cluster (DSMEM reduction of the ChannelGate's
// global average pool), gate folded into per-stream scaled copies of conv3's weights,
// conv3 + downsample accumulated in TMEM, residual + ReLU in the final epilogue.
#ifdef SSB_BASELINES        // round-1 OSBlock kernel, kept as an A/B baseline: libssb_dbg.so only
#include <cooperative_groups.h>

#include "ssb_common.cuh"
#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int K3_THREADS = 512;
constexpr int K3_GROUPS = K3_THREADS / 128;

__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int rup128(int a) { return (a + 127) / 128 * 128; }

template <int CIN_, int MID_, int MIDP_, int COUT_, int H_, int W_, int R_, int HALO_, int NB_,
          bool DOWN_, int NSTAGE_, int SEG_>
struct B3 {
    static constexpr int CIN = CIN_, MID = MID_, MIDP = MIDP_, COUT = COUT_, H = H_, W = W_, R = R_;
    static constexpr int HALO = HALO_, NB = NB_, NSTAGE = NSTAGE_, SEG = SEG_;
    static constexpr bool DOWN = DOWN_;
    static constexpr int RH = R + 2 * HALO;               // band rows kept (with halo)
    static constexpr int NPX = RH * W;                    // band pixels, p = lr * W + col
    static constexpr int NT = NPX / 128;                  // M tiles
    static constexpr int OWN_P0 = HALO * W, OWN_P1 = (HALO + R) * W;
    static constexpr int IT0 = OWN_P0 / 128, IT1 = (OWN_P1 + 127) / 128, NIT = IT1 - IT0;
    static constexpr int MCH = MIDP / 8;                  // 16-byte K chunks per pixel
    static constexpr int CG = MID / 4;                    // real float4 channel groups
    // depthwise pass: a warp = XL adjacent columns x 2 channel groups of one 8-channel K chunk
    // (x CPW chunk pairs when the map is narrower than 16), walking down a row segment
    static constexpr int XL = W < 16 ? W : 16;            // column lanes
    static constexpr int CPW = 16 / XL;                   // channel-group pairs per warp
    static constexpr int NCB = W / XL;                    // column blocks
    static constexpr int NCGW = (CG / 2) / CPW;           // channel-group-pair groups
    static constexpr int TPS = NCB * NCGW;                // warp tasks per row segment
    static constexpr int PLANE_B = NPX * 16;              // LBO of a map operand
    static constexpr int MAP_HALF_B = MCH * PLANE_B, MAP_B = 2 * MAP_HALF_B;   // hi planes, lo planes
    static constexpr int TW = W + 2, TH = RH + 2, TPX = TW * TH;
    static constexpr int T_B = CG * TPX * 16;             // fp32 pointwise result with zero ring
    static constexpr int STG_HALF_B = (CIN / 8) * 128 * 16, STG_B = 2 * STG_HALF_B;
    static constexpr int A_B = rup128(cmax(MAP_B + T_B, NSTAGE * STG_B));      // P map + T | x staging
    static constexpr int LCN = 2 * MIDP;                  // TMEM columns of a pointwise tile (hi | lo weights)
    static constexpr int TM_C3 = NT * LCN, TM_COLS = TM_C3 + NIT * COUT;
    static constexpr int C1W_B = CIN * MIDP * 4;
    static constexpr int DNW_HALF_B = DOWN ? CIN * COUT * 2 : 0, DNW_B = 2 * DNW_HALF_B;
    static constexpr int LCW_B = MIDP * MIDP * 4;
    static constexpr int WALL_B = C1W_B + DNW_B + 10 * LCW_B;
    static constexpr int C3W_HALF_B = MIDP * COUT * 2, C3W_B = 2 * C3W_HALF_B;
    // PAR (floats): B1[MIDP] | 10 x { DW[9][MIDP], B[MIDP] } | B3[COUT] | GW1[MIDP][2] | GB1[2] |
    //               GW2[2][MIDP] | GB2[MIDP]
    static constexpr int P_B1 = 0, P_LC = MIDP, P_B3 = P_LC + 100 * MIDP, P_GW1 = P_B3 + COUT;
    static constexpr int P_GB1 = P_GW1 + 2 * MIDP, P_GW2 = P_GB1 + 2, P_GB2 = P_GW2 + 2 * MIDP;
    static constexpr int NPAR = P_GB2 + MIDP;
    // shared-memory carve-up (bytes)
    static constexpr int OFF_X1 = 0;
    static constexpr int OFF_A = OFF_X1 + MAP_B;
    static constexpr int OFF_W = OFF_A + A_B;
    static constexpr int OFF_C3 = OFF_W + rup128(WALL_B);
    static constexpr int OFF_PAR = OFF_C3 + C3W_B;
    static constexpr int OFF_GAP = OFF_PAR + rup128(NPAR * 4);       // [4 streams][NB bands][MIDP] floats
    static constexpr int OFF_MISC = OFF_GAP + 4 * NB * MIDP * 4;
    static constexpr int SCR_FLOATS = SEG * NCB * CG * 4;
    static constexpr int SMEM_B = OFF_MISC + 192 + (SCR_FLOATS + 2 * MIDP) * 4 + 64;
    static_assert(NPX % 128 == 0, "band = whole M tiles");
    static_assert((W & (W - 1)) == 0, "W power of two");
    static_assert(NT <= 8, "per-tile barriers");
    static_assert(TM_COLS <= 512, "TMEM columns");
    static_assert(SMEM_B <= 232448, "shared memory");
    static_assert(MIDP % 16 == 0 && COUT % 16 == 0 && CIN % 16 == 0 && MID % 4 == 0, "MMA shapes");
    static_assert(H % R == 0 && H / R == NB, "bands");
    static_assert(MID <= MIDP && COUT == 4 * MID && MIDP <= 32 && COUT * (MIDP / 8) <= K3_THREADS, "OSBlock channel plan");
    static_assert(CG % 2 == 0 && (CG / 2) % CPW == 0 && W % XL == 0, "depthwise warp tasks");
    static_assert(SEG * TPS < K3_THREADS / 32, "depthwise warps + one MMA-issuing warp");
    static_assert((NT * CG) % K3_GROUPS == 0, "pointwise drain units divide the groups");
    static_assert(16 * 32 * 36 * 4 <= OFF_C3, "final-epilogue staging fits in the dead maps/weights");
    // global blob sections (bytes): C1W | DNW | LCW[10] | PAR (fp32) | W3 (fp32 [MIDP][COUT])
    static constexpr int G_PAR = rup128(WALL_B);
    static constexpr int G_W3 = G_PAR + rup128(NPAR * 4);
    static constexpr int G_TOTAL = G_W3 + MIDP * COUT * 4;
};

__device__ __forceinline__ void split2(float a, float b, __half2 &h, __half2 &l) {
    h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    l = __floats2half2_rn(a - hf.x, b - hf.y);
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
// 4 consecutive TMEM columns of this warp's 32 lanes, no wait (pair with tmem_wait_ld)
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
// four fp32 channels of the depthwise pass.  SSB_DW_FFMA2 = 1: two packed pairs and Blackwell's FFMA2
// (fma.rn.f32x2, two IEEE fp32 FMAs per instruction); 0: four scalar FFMAs.  Same results either way.
#ifndef SSB_DW_FFMA2
#define SSB_DW_FFMA2 1
#endif
#if SSB_DW_FFMA2
struct P4 { unsigned long long a, b; };          // channels (0,1), (2,3)
__device__ __forceinline__ P4 ldp4(const float4 *p) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p);
    return P4{v.x, v.y};
}
__device__ __forceinline__ void fma4(P4 &o, const P4 &w, const P4 &v) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(o.a) : "l"(w.a), "l"(v.a));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(o.b) : "l"(w.b), "l"(v.b));
}
__device__ __forceinline__ float p4x(const P4 &v) { return __uint_as_float((uint32_t)v.a); }
__device__ __forceinline__ float p4y(const P4 &v) { return __uint_as_float((uint32_t)(v.a >> 32)); }
__device__ __forceinline__ float p4z(const P4 &v) { return __uint_as_float((uint32_t)v.b); }
__device__ __forceinline__ float p4w(const P4 &v) { return __uint_as_float((uint32_t)(v.b >> 32)); }
#else
typedef float4 P4;
__device__ __forceinline__ P4 ldp4(const float4 *p) { return *p; }
__device__ __forceinline__ void fma4(P4 &o, const P4 &w, const P4 &v) {
    o.x = fmaf(w.x, v.x, o.x); o.y = fmaf(w.y, v.y, o.y);
    o.z = fmaf(w.z, v.z, o.z); o.w = fmaf(w.w, v.w, o.w);
}
__device__ __forceinline__ float p4x(const P4 &v) { return v.x; }
__device__ __forceinline__ float p4y(const P4 &v) { return v.y; }
__device__ __forceinline__ float p4z(const P4 &v) { return v.z; }
__device__ __forceinline__ float p4w(const P4 &v) { return v.w; }
#endif

template <class C>
__global__ void __launch_bounds__(K3_THREADS, 1)
osblock3_kernel(const float *__restrict__ x, float *__restrict__ y,
                const unsigned char *__restrict__ wblob, int n_crops, int *__restrict__ status,
                long long *__restrict__ dbg) {
    extern __shared__ __align__(1024) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, grp = warp >> 2;          // TMEM lane quadrant / work group
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);            // provably warp-uniform copy of warp
    const int crop = blockIdx.x / C::NB, band = blockIdx.x % C::NB;
    const int row0 = band * C::R - C::HALO;              // image row of band-local row 0

    unsigned char *sX1 = smem + C::OFF_X1, *sP = smem + C::OFF_A;
    float4 *sT = reinterpret_cast<float4 *>(sP + C::MAP_B);
    unsigned char *sW = smem + C::OFF_W, *sC3 = smem + C::OFF_C3;
    float *sPar = reinterpret_cast<float *>(smem + C::OFF_PAR);
    float *sGap = reinterpret_cast<float *>(smem + C::OFF_GAP);      // [4][NB][MIDP] band-partial sums of every band
    uint64_t *bar_w = reinterpret_cast<uint64_t *>(smem + C::OFF_MISC);
    uint64_t *bar_stg = bar_w + 1;                     // [2] one per x-staging buffer
    uint64_t *bar_tile = bar_w + 3;                    // [8] per M tile of the current pointwise conv
    uint64_t *bar_c3 = bar_w + 11;                     // conv3 accumulation of the current stream
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bar_w + 16);
    float *s_scr = reinterpret_cast<float *>(smem + C::OFF_MISC + 192);   // [SEG][CG][4] partial sums

    if (warp == 0) tc::tmem_alloc(s_tmem, 512);
    if (tid == 0) {
        tc::mbar_init(bar_w, 1);
        tc::mbar_init(bar_stg, 1);
        tc::mbar_init(bar_stg + 1, 1);
        for (int i = 0; i < 8; i++) tc::mbar_init(bar_tile + i, 1);
        tc::mbar_init(bar_c3, 1);
        tc::fence_mbar_init();
    }
    for (int i = tid; i < C::NPAR; i += K3_THREADS)
        sPar[i] = reinterpret_cast<const float *>(wblob + C::G_PAR)[i];
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *s_tmem;
    int dbg_n = 0;
    auto stamp = [&]() { if (dbg && blockIdx.x == 0 && tid == 0 && dbg_n < 63) dbg[1 + dbg_n++] = clock64(); };
    stamp();
    bool ok = true;

    // every weight operand of the block (conv1, downsample, 10 pointwise sets) in one go
    if (tid == 0) {
        tc::mbar_arrive_expect_tx(bar_w, C::WALL_B);
        for (int o = 0; o < C::WALL_B; o += 32768) {
            const int nb = C::WALL_B - o < 32768 ? C::WALL_B - o : 32768;
            tc::bulk_g2s(sW + o, wblob + o, nb, bar_w);
        }
    }

    constexpr uint32_t IDESC_MID = tc::make_idesc_f16(128, C::MIDP);
    constexpr uint32_t IDESC_CAT = tc::make_idesc_f16(128, 2 * C::MIDP);
    constexpr uint32_t IDESC_OUT = tc::make_idesc_f16(128, C::COUT);

    // ------------------------------------------------------------------
    // phase 1: X1 = relu(conv1(x)) on every band tile; downsample on the own tiles
    // ------------------------------------------------------------------
    const float *xin = x + (size_t)crop * C::H * C::W * C::CIN;
    uint32_t stg_phase[2] = {0, 0};
    constexpr int F4 = C::CIN / 4;
    constexpr int PER = (128 * F4) / K3_THREADS;             // float4 items per thread per tile
    static_assert((128 * F4) % K3_THREADS == 0, "tile items divide the CTA");
    // two tiles of x are in flight while a third is staged (the phase is bound by L2 latency x bytes in flight)
    float4 xr[2][PER];
    auto load_tile = [&](int t, float4 *dst) {
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int idx = tid + q * K3_THREADS;
            const int px = idx / F4, f4 = idx - px * F4;
            const int p = t * 128 + px;
            const int gr = row0 + p / C::W, gc = p % C::W;
            dst[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gr >= 0 && gr < C::H)
                dst[q] = *reinterpret_cast<const float4 *>(xin + ((size_t)gr * C::W + gc) * C::CIN + f4 * 4);
        }
    };
    load_tile(0, xr[0]);
    if (C::NT > 1) load_tile(1, xr[1]);
#pragma unroll
    for (int t = 0; t < C::NT; t++) {
        const int sb = t % C::NSTAGE;
        unsigned char *stg = sP + sb * C::STG_B;
        if (t >= C::NSTAGE) {                                // MMAs that read this buffer are done
            if (!tc::mbar_wait(bar_stg + sb, stg_phase[sb])) ok = false;
            stg_phase[sb] ^= 1;
        }
#pragma unroll
        for (int q = 0; q < PER; q++) {                      // stage tile t of x: [CIN/8][128][8] hi, then lo
            const int idx = tid + q * K3_THREADS;
            const int px = idx / F4, f4 = idx - px * F4;
            const float4 v = xr[t & 1][q];
            __align__(8) __half2 h[2], l[2];
            split2(v.x, v.y, h[0], l[0]);
            split2(v.z, v.w, h[1], l[1]);
            const int off = (f4 >> 1) * 2048 + px * 16 + (f4 & 1) * 8;
            *reinterpret_cast<uint2 *>(stg + off) = *reinterpret_cast<uint2 *>(h);
            *reinterpret_cast<uint2 *>(stg + C::STG_HALF_B + off) = *reinterpret_cast<uint2 *>(l);
        }
        if (t + 2 < C::NT) load_tile(t + 2, xr[t & 1]);
        tc::fence_async_smem();
        if (t == 0) { if (!tc::mbar_wait(bar_w, 0)) ok = false; }
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        if (warp_u == 0 && tc::elect_one()) {
            const uint64_t ah0 = tc::make_smem_desc(tc::smem_u32(stg), 2048, 128);
            const uint64_t al0 = dadv(ah0, C::STG_HALF_B / 16);
            const uint64_t bc0 = tc::make_smem_desc(tc::smem_u32(sW), 2 * C::MIDP * 16, 128);
            const uint32_t d1 = tmem + t * C::LCN;
#pragma unroll
            for (int ks = 0; ks < C::CIN / 16; ks++) {
                tc::mma_f16_ss(d1, dadv(ah0, ks * 256), dadv(bc0, ks * 4 * C::MIDP), IDESC_CAT, ks > 0);
                tc::mma_f16_ss(d1, dadv(al0, ks * 256), dadv(bc0, ks * 4 * C::MIDP), IDESC_MID, 1);
            }
            if (C::DOWN && t >= C::IT0 && t < C::IT1) {
                const uint64_t dh0 = tc::make_smem_desc(tc::smem_u32(sW) + C::C1W_B, C::COUT * 16, 128);
                const uint64_t dl0 = dadv(dh0, C::DNW_HALF_B / 16);
                const uint32_t d2 = tmem + C::TM_C3 + (t - C::IT0) * C::COUT;
#pragma unroll
                for (int ks = 0; ks < C::CIN / 16; ks++)
                    mma3(d2, dadv(ah0, ks * 256), dadv(al0, ks * 256), dadv(dh0, ks * 2 * C::COUT),
                         dadv(dl0, ks * 2 * C::COUT), IDESC_OUT, ks > 0);
            }
            tc::mma_commit(bar_stg + sb);
        }
    }
    for (int sb = 0; sb < C::NSTAGE && sb < C::NT; sb++) {    // the last commit of every buffer
        if (!tc::mbar_wait(bar_stg + sb, stg_phase[sb])) ok = false;
        stg_phase[sb] ^= 1;
    }
    tc::fence_after_sync();
    stamp();                                   // [1] phase 1 (staging + conv1/down MMAs) done

    // the staging area becomes the P map (its pad-channel planes must read as zero) and T (zero ring)
    for (int i = tid; i < (C::MAP_B + C::T_B) / 16; i += K3_THREADS)
        reinterpret_cast<uint4 *>(sP)[i] = make_uint4(0u, 0u, 0u, 0u);
    // X1 epilogue: TMEM tile -> (+bias, relu, out-of-image mask) -> hi/lo operand map; units (tile, 8-channel
    // K chunk) spread evenly over the 4 warp groups, all TMEM loads of a thread in flight before one wait
    {
        constexpr int NU1 = C::NT * C::MCH, UPT1 = (NU1 + K3_GROUPS - 1) / K3_GROUPS;
        uint32_t va[UPT1][8], vb[UPT1][8];
#pragma unroll
        for (int e = 0; e < UPT1; e++) {
            const int u = grp + e * K3_GROUPS;
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
            const int u = grp + e * K3_GROUPS;
            if (u < NU1) {
                const int t = u / C::MCH, kc = u - t * C::MCH;
                const int p = t * 128 + quad * 32 + lane;
                const int gr = row0 + p / C::W;
                const bool valid = gr >= 0 && gr < C::H;
                const float *bias = sPar + C::P_B1 + kc * 8;
                __align__(16) __half2 h[4];
                __align__(16) __half2 l[4];
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const float f0 = valid ? fmaxf(__uint_as_float(va[e][j]) + __uint_as_float(vb[e][j]) + bias[j], 0.f) : 0.f;
                    const float f1 = valid ? fmaxf(__uint_as_float(va[e][j + 1]) + __uint_as_float(vb[e][j + 1]) + bias[j + 1], 0.f) : 0.f;
                    split2(f0, f1, h[j >> 1], l[j >> 1]);
                }
                unsigned char *d_hi = sX1 + kc * C::PLANE_B + p * 16;
                *reinterpret_cast<uint4 *>(d_hi) = *reinterpret_cast<uint4 *>(h);
                *reinterpret_cast<uint4 *>(d_hi + C::MAP_HALF_B) = *reinterpret_cast<uint4 *>(l);
            }
        }
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    stamp();                                   // [2] X1 drained

    // ------------------------------------------------------------------
    // phase 2: four streams of LightConvs (pointwise on tcgen05, depthwise on the CUDA cores)
    //          + gated conv3 accumulation
    // ------------------------------------------------------------------
    // MMAs are issued from warp-uniform code by an elected lane (~28 cycles per MMA + commit instead of
    // ~90 from a divergent single-thread branch, measured) of a warp that has no depthwise rows
    // (SEG * TPS < 16 warps), so nobody waits for the issue
    const bool issuer = warp_u == C::SEG * C::TPS;
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
    float w3r[8];                             // conv3 weights of this thread's operand row (kc, co)
    {
        const float *w3 = reinterpret_cast<const float *>(wblob + C::G_W3);  // [MIDP][COUT]
        const int kc = tid / C::COUT, co = tid - kc * C::COUT;
#pragma unroll
        for (int j = 0; j < 8; j++) w3r[j] = tid < C::COUT * C::MCH ? w3[(kc * 8 + j) * C::COUT + co] : 0.f;
    }
    // depthwise task of this thread (see B3::XL): column, channel group (parity = lane bit XL), row segment
    // lane = 2 * column + parity (+ 2 * XL * chunk pair): the two lanes of a pixel are adjacent, so their
    // 8-byte hi (and lo) halves of a 16-byte operand row are one contiguous, conflict-free store
    const int dw_par = lane & 1;
    const int dw_seg = warp / C::TPS, dw_cb = (warp % C::TPS) % C::NCB;
    const int dw_cg = (((warp % C::TPS) / C::NCB) * C::CPW + lane / (2 * C::XL)) * 2 + dw_par;
    const int dw_col = dw_cb * C::XL + (lane >> 1) % C::XL;
    int lc = 0;
    uint32_t tile_par = 0, c3_par = 0;
    bool c3_pending = false;
    if (issuer) issue_pw(sX1, 0, 0, C::NT);
    for (int s = 0; s < 4; s++) {
        for (int k = 0; k <= s; k++, lc++) {
            const int rem = s - k;                                // LightConvs after this one in the stream
            const bool last = (k == s);
            // ---- this layer's depthwise taps into registers: shared-memory traffic that rides under
            //      the wait for the pointwise MMAs instead of the LSU-bound depthwise pass
            P4 wd[9], bs;
            {
                const float *wl = sPar + C::P_LC + lc * (10 * C::MIDP) + dw_cg * 4;
#pragma unroll
                for (int tap = 0; tap < 9; tap++) wd[tap] = ldp4(reinterpret_cast<const float4 *>(wl + tap * C::MIDP));
                bs = ldp4(reinterpret_cast<const float4 *>(wl + 9 * C::MIDP));
            }
            // ---- pointwise result: TMEM -> fp32 T (zero ring untouched), units (tile, channel group);
            //      all TMEM loads of the thread are in flight before the single wait
            {
                constexpr int UPT = (C::NT * C::CG) / K3_GROUPS;
                uint32_t ra_[UPT][4], rb_[UPT][4];
#pragma unroll
                for (int i = 0; i < UPT; i++) {
                    const int u = grp + i * K3_GROUPS;
                    const int t = u / C::CG, cgi = u - t * C::CG;
                    if (!tc::mbar_wait(bar_tile + t, tile_par)) ok = false;
                    tc::fence_after_sync();
                    const uint32_t ta = tmem + ((uint32_t)(quad * 32) << 16) + t * C::LCN + cgi * 4;
                    tmem_ld4_nw(ta, ra_[i]);
                    tmem_ld4_nw(ta + C::MIDP, rb_[i]);
                }
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < UPT; i++) {
                    const int u = grp + i * K3_GROUPS;
                    const int t = u / C::CG, cgi = u - t * C::CG;
                    const int p = t * 128 + quad * 32 + lane;
                    const int lr = p / C::W, col = p % C::W;
                    sT[cgi * C::TPX + (lr + 1) * C::TW + col + 1] =
                        make_float4(__uint_as_float(ra_[i][0]) + __uint_as_float(rb_[i][0]),
                                    __uint_as_float(ra_[i][1]) + __uint_as_float(rb_[i][1]),
                                    __uint_as_float(ra_[i][2]) + __uint_as_float(rb_[i][2]),
                                    __uint_as_float(ra_[i][3]) + __uint_as_float(rb_[i][3]));
                }
            }
            tile_par ^= 1;
            tc::fence_before_sync();
            __syncthreads();
            tc::fence_after_sync();
            stamp();                           // T ready
            // the next stream starts from X1: its pointwise conv runs under this depthwise pass
            if (last && s < 3 && issuer) issue_pw(sX1, lc + 1, 0, C::NT);
            // the previous stream's conv3 MMAs read P: done before this stream overwrites it
            if (k == 0 && c3_pending) {
                if (!tc::mbar_wait(bar_c3, c3_par)) ok = false;
                c3_par ^= 1;
                c3_pending = false;
                tc::fence_after_sync();
            }
            // ---- depthwise 3x3 + bias + ReLU -> hi/lo operand map P (in place: the pointwise
            //      MMAs that read P have completed); rows [ra, rb) are the "trapezoid" this
            //      layer has to get right for the rem layers after it
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
                    P4 o = bs;
#pragma unroll
                    for (int dx = 0; dx < 3; dx++) {
                        fma4(o, wd[dx], wa[dx]);
                        fma4(o, wd[3 + dx], wb[dx]);
                        fma4(o, wd[6 + dx], wc[dx]);
                    }
                    float ox = fmaxf(p4x(o), 0.f), oy = fmaxf(p4y(o), 0.f);
                    float oz = fmaxf(p4z(o), 0.f), ow = fmaxf(p4w(o), 0.f);
                    const int gr = row0 + lr;
                    if (gr < 0 || gr >= C::H) { ox = 0.f; oy = 0.f; oz = 0.f; ow = 0.f; }     // warp-uniform
                    if (last) {                // rows [ra, rb) == the band's own rows when rem == 0
                        gacc.x += ox; gacc.y += oy; gacc.z += oz; gacc.w += ow;
                    }
                    __half2 h[2], l[2];
                    split2(ox, oy, h[0], l[0]);
                    split2(oz, ow, h[1], l[1]);
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
            const int ra = C::HALO - rem > 0 ? C::HALO - rem : 0;
            const int rb = C::HALO + C::R + rem < C::RH ? C::HALO + C::R + rem : C::RH;
            if (!last) {
                // the next LightConv's pointwise conv is per pixel: its MMAs on the upper half of the
                // tiles are issued as soon as the upper rows are written and run under the lower
                // half's depthwise pass; the lower half's MMAs run under the upper half's drain
                if (C::NT >= 2) {
                    constexpr int TH_ = C::NT / 2, HR = TH_ * 128 / C::W;
                    dw_rows(ra, rb < HR ? rb : HR);
                    publish();
                    if (issuer) issue_pw(sP, lc + 1, 0, TH_);
                    dw_rows(ra > HR ? ra : HR, rb);
                    publish();
                    if (issuer) issue_pw(sP, lc + 1, TH_, C::NT);
                } else {
                    dw_rows(ra, rb);
                    publish();
                    if (issuer) issue_pw(sP, lc + 1, 0, C::NT);
                }
                stamp();                       // depthwise done, next pointwise issued
                continue;
            }
            dw_rows(ra, rb);
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
            stamp();                           // depthwise done
            // ---- ChannelGate: band-partial sums -> cluster -> mean -> MLP -> sigmoid -> scaled conv3 weights
            // every band PUSHES its partial sums into all bands' shared memory (st.shared::cluster); the
            // cluster barrier (release / acquire) then makes them visible, and the sums are read locally --
            // no remote loads after the barrier.  (Per-lane remote mbarrier arrives instead of the hardware
            // cluster barrier were measured slower: 2.65 K vs 2.2 K cycles per gate.)
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
            if (C::NB > 1) cluster.sync(); else __syncthreads();
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
                // gate-scaled conv3 weights  B[kc][co][8] = W3[k][co] * g[k]  (hi / lo); the thread's
                // 8 fp32 W3 values live in registers for the whole kernel
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
                    for (int i = 0; i < C::NIT; i++) {
                        const uint32_t d = tmem + C::TM_C3 + i * C::COUT;
#pragma unroll
                        for (int ks = 0; ks < C::MIDP / 16; ks++) {
                            const int ka = (C::IT0 + i) * 128 + ks * 2 * C::NPX, kb = ks * 2 * C::COUT;
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
    // identity blocks: the residual x rows of this warp's (tile, 32-column) units are fetched (coalesced:
    // 8 lanes per 128-byte pixel row) before the wait for the last conv3 MMAs, so their latency is hidden
    float *yout = y + (size_t)crop * C::H * C::W * C::COUT;
    constexpr int CCH = C::COUT / 32;                         // 32-column chunks per tile
    constexpr int NUE = C::NIT * CCH, UE = (NUE + K3_GROUPS - 1) / K3_GROUPS;
    float4 xres[C::DOWN ? 1 : UE][8];
    int rowoffs[UE];
#pragma unroll
    for (int e = 0; e < UE; e++) {
        const int u = grp + e * K3_GROUPS;
        rowoffs[e] = -1;
        if (u < NUE) {                                        // warp-uniform
            const int i = u / CCH, c0 = (u - i * CCH) * 32;
            const int p = (C::IT0 + i) * 128 + quad * 32 + lane;
            const bool own = p >= C::OWN_P0 && p < C::OWN_P1;
            rowoffs[e] = own ? band * C::R * C::W + (p - C::OWN_P0) : -1;     // pixel index inside the crop
            if (!C::DOWN) {
#pragma unroll
                for (int it = 0; it < 8; it++) {
                    const int ro = __shfl_sync(0xffffffffu, rowoffs[e], it * 4 + (lane >> 3));
                    xres[C::DOWN ? 0 : e][it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ro >= 0)
                        xres[C::DOWN ? 0 : e][it] = *reinterpret_cast<const float4 *>(xin + (size_t)ro * C::CIN + c0 + (lane & 7) * 4);
                }
            }
        }
    }
    if (c3_pending) {                 // the last stream's conv3 MMAs
        if (!tc::mbar_wait(bar_c3, c3_par)) ok = false;
        c3_par ^= 1;
    }
    tc::fence_after_sync();
    __syncthreads();                  // every thread is past its last read of the maps / weights

    // ------------------------------------------------------------------
    // final epilogue: y = relu(conv3 + bias (+ downsample already in TMEM) (+ x)), rows transposed
    // through a warp-private staging tile so that 8 lanes cover one 128-byte pixel row
    // ------------------------------------------------------------------
    float *stage = reinterpret_cast<float *>(smem) + warp * (32 * 36);
#pragma unroll
    for (int e = 0; e < UE; e++) {
        const int u = grp + e * K3_GROUPS;
        if (u >= NUE) continue;                               // warp-uniform
        const int i = u / CCH, c0 = (u - i * CCH) * 32;
        const int rowoff = rowoffs[e];
        if (!C::DOWN) {
#pragma unroll
            for (int it = 0; it < 8; it++)
                *reinterpret_cast<float4 *>(stage + (it * 4 + (lane >> 3)) * 36 + (lane & 7) * 4) = xres[C::DOWN ? 0 : e][it];
        }
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + C::TM_C3 + i * C::COUT + c0, v);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 bb = *reinterpret_cast<const float4 *>(sPar + C::P_B3 + c0 + j);
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
    if (!ok) { if (tid == 0) atomicExch(status, 5); }
    tc::fence_before_sync();
    if (C::NB > 1) cluster.sync(); else __syncthreads();     // no band exits while its peers may still push into it
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
    (void)n_crops;
}

// ---------------------------------------------------------------------------
// the six OSBlocks of osnet_x0_25 (stage 2: 64x32, stage 3: 32x16, stage 4: 16x8)
// ---------------------------------------------------------------------------
//            CIN MID MIDP COUT  H   W   R HALO NB DOWN NSTAGE SEG
using K0 = B3<16, 16, 16, 64, 64, 32, 16, 4, 4, true, 2, 3>;
using K1 = B3<64, 16, 16, 64, 64, 32, 16, 4, 4, false, 2, 3>;
using K2 = B3<64, 24, 32, 96, 32, 16, 8, 4, 4, true, 2, 5>;
using K3 = B3<96, 24, 32, 96, 32, 16, 8, 4, 4, false, 2, 5>;
using K4 = B3<96, 32, 32, 128, 16, 8, 16, 0, 1, true, 1, 7>;
using K5 = B3<128, 32, 32, 128, 16, 8, 16, 0, 1, false, 1, 7>;

template <class C>
int launch3(const float *x, float *y, const unsigned char *w, int n, int *status, long long *dbg,
            cudaStream_t st) {
    static const int key = ssb_new_key();
    if (ssb_first_on_device(key))
        SSB_CHECK_CUDA(cudaFuncSetAttribute(osblock3_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_B));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n * C::NB);
    cfg.blockDim = dim3(K3_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_B;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C::NB;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SSB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, osblock3_kernel<C>, x, y, w, n, status, dbg));
    g_ssb_launches++;
    return 0;
}

}  // namespace

int64_t ssb_reid_tc3_block_bytes(int b) {
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

int ssb_reid_tc3_block(int b, const float *x, float *y, const unsigned char *w, int n, int *status,
                       cudaStream_t st) {
    long long *dbg = g_ssb_tc_dbg;
    switch (b) {
        case 0: return launch3<K0>(x, y, w, n, status, dbg, st);
        case 1: return launch3<K1>(x, y, w, n, status, dbg, st);
        case 2: return launch3<K2>(x, y, w, n, status, dbg, st);
        case 3: return launch3<K3>(x, y, w, n, status, dbg, st);
        case 4: return launch3<K4>(x, y, w, n, status, dbg, st);
        case 5: return launch3<K5>(x, y, w, n, status, dbg, st);
    }
    ssb_set_error("bad OSBlock index %d", b);
    return -1;
}

#endif  // SSB_BASELINES
