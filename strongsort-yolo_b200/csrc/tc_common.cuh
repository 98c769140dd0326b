// tcgen05 / TMEM / mbarrier primitives for sm_100a, written as inline PTX.
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" /
// "instruction descriptor" tables (also mirrored by CUTLASS's
// cute/arch/mma_sm100_desc.hpp, which was read for the field positions only).
//
// Operand layout used throughout this repo: K-major, NO swizzle ("interleave"):
//   element (row r, k) of an operand tile lives at
//       start + (r % 8) * 16 + (r / 8) * SBO + (k / 8) * LBO + (k % 8) * 2   bytes (fp16)
// With SBO = 128 the rows are simply 16 bytes apart, i.e. a K-chunk plane is
// [rows][8 halves] -- which makes "shift the A tile by p rows" a plain +16*p on
// the descriptor's start address.  That is how the 3x3 convolutions are fed to
// the tensor core as 9 shifted GEMMs without an im2col copy.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- shared-memory matrix descriptor (K-major, no swizzle) -----------------
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version 1 (Blackwell)
    return d;                        // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// ---- instruction descriptor, kind::f16, fp16 x fp16 -> fp32, both K-major ---
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4)                 // c_format  = F32
           | (0u << 7)               // a_format  = F16
           | (0u << 10)              // b_format  = F16
           | (0u << 15) | (0u << 16) // a_major, b_major = K
           | ((uint32_t)(N >> 3) << 17)
           | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc,
                                           uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// One lane of a CONVERGED warp.  Issue tcgen05.mma / commit from warp-uniform code under this
// predicate: measured on B200 (tools/probes/mma_issue_probe.cu) an MMA + commit costs the issuing
// thread ~28 cycles this way and ~90 cycles from a divergent `if (tid == 0)` branch, where the
// compiler wraps every UTCHMMA in an ELECT / BRA.U.ANY loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

// all previously issued MMAs of this thread arrive on the mbarrier when done
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}

// ---- TMEM management (one full warp executes alloc / dealloc) ---------------
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// generic-proxy smem writes -> visible to the async proxy (UMMA / TMA reads)
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
// 32 consecutive columns in one instruction (one wait per 32 values)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
template <int N>
__device__ __forceinline__ void tmem_ldN(uint32_t taddr, float *v) {
    static_assert(N == 16 || N == 32, "tmem_ldN");
    if (N == 16) tmem_ld16(taddr, v); else tmem_ld32(taddr, v);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}

// ---- mbarrier ----------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// bounded wait so that a protocol bug cannot hang the GPU: returns false after
// ~0.25 s.  Plain try_wait blocks in hardware until the phase flips or an
// implementation time limit passes (SYNCS.PHASECHK.TRYWAIT); a suspend-time
// hint must NOT be given -- ptxas turns it into a NANOSLEEP of that length and
// 45 % of the OSBlock kernel's samples sat in it (profiles/r01_tc_v1).
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    long long t0 = 0;
    for (uint32_t it = 0;; it++) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return true;
        // back off: a hot spin steals issue slots from the MMA-issuing warps that share
        // this SM sub-partition (ncu counted 5.4 M polls per launch before this)
        __nanosleep(it < 8 ? 32 : 128);
        if (it == 64) t0 = clock64();
        if (it > 64 && clock64() - t0 > 500000000LL) return false;
    }
}

// ---- distributed shared memory (thread-block cluster): push + remote mbarrier arrive ----------
// address of `local_smem_ptr`'s counterpart in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_u32(const void *local_smem_ptr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local_smem_ptr)), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t cluster_addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
// arrive (release at cluster scope: this thread's earlier remote stores are visible to the waiter)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
// bounded wait with acquire at cluster scope (pairs with mbar_arrive_cluster from peer CTAs)
__device__ __forceinline__ bool mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    long long t0 = 0;
    for (uint32_t it = 0;; it++) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return true;
        __nanosleep(it < 8 ? 32 : 128);
        if (it == 64) t0 = clock64();
        if (it > 64 && clock64() - t0 > 500000000LL) return false;
    }
}

// programmatic dependent launch (launch attribute programmaticStreamSerializationAllowed): a kernel may start
// while its predecessor on the stream is still running; griddepcontrol.wait blocks until the predecessor has
// completed and its memory is visible, launch_dependents lets the NEXT kernel begin its own prologue.  Both are
// no-ops when the kernel was launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// 1-D bulk async copy global -> shared (TMA engine, no tensor map needed),
// completion counted in bytes on an mbarrier.  size % 16 == 0, 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace tc
