// Diagnostic: one tcgen05 GEMM tile through the exact operand layout the ReID
// kernels use (K-major, no swizzle, SBO = 128 so that A can be row-shifted by
// moving the descriptor's start address).  tests/test_gpu_tc.py checks it
// against a float32 matmul; it exists to pin the descriptor encodings of
// tc_common.cuh on real hardware before the fused kernels rely on them.
#ifdef SSB_BASELINES        // diagnostic: libssb_dbg.so only
#include "ssb_common.cuh"
#include "tc_common.cuh"

// D[128][N] = A[shift : shift+128][K] * B[N][K]^T     (fp16 in, fp32 out)
__global__ void __launch_bounds__(128)
tc_probe_kernel(const __half *__restrict__ A, int a_rows, int shift, const __half *__restrict__ B,
                int N, int K, float *__restrict__ D, int *__restrict__ status) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kc = K / 8;                               // 16-byte chunks along K
    __half *sA = reinterpret_cast<__half *>(smem);                       // [kc][a_rows][8]
    __half *sB = sA + (size_t)kc * a_rows * 8;                           // [kc][N][8]
    for (int i = tid; i < a_rows * K; i += 128) {
        const int r = i / K, k = i - r * K;
        sA[((size_t)(k >> 3) * a_rows + r) * 8 + (k & 7)] = A[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i - r * K;
        sB[((size_t)(k >> 3) * N + r) * 8 + (k & 7)] = B[i];
    }
    uint32_t ncols = 32;
    while ((int)ncols < N) ncols <<= 1;
    if (warp == 0) tc::tmem_alloc(&s_tmem, ncols);
    if (tid == 0) { tc::mbar_init(&s_bar, 1); tc::fence_mbar_init(); }
    tc::fence_async_smem();          // make the st.shared operand writes visible to the tensor core
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_f16(128, N);
        const uint32_t lboA = (uint32_t)a_rows * 16, lboB = (uint32_t)N * 16;
        for (int ks = 0; ks < K / 16; ks++) {
            const uint64_t da = tc::make_smem_desc(tc::smem_u32(sA) + shift * 16 + 2 * ks * lboA, lboA, 128);
            const uint64_t db = tc::make_smem_desc(tc::smem_u32(sB) + 2 * ks * lboB, lboB, 128);
            tc::mma_f16_ss(tmem, da, db, idesc, ks > 0);
        }
        tc::mma_commit(&s_bar);
    }
    const bool ok = tc::mbar_wait(&s_bar, 0);
    tc::fence_after_sync();
    if (!ok) { if (tid == 0) status[0] = 1; }
    else {
        for (int c0 = 0; c0 < N; c0 += 16) {
            float v[16];
            tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            const int row = warp * 32 + lane;
#pragma unroll
            for (int j = 0; j < 16; j++) D[(size_t)row * N + c0 + j] = v[j];
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

extern "C" int ssb_tc_probe(const void *a_dev, int a_rows, int shift, const void *b_dev, int n, int k,
                            float *d_dev, int32_t *status_dev, ssb_stream_t stream) {
    if (n % 16 || n < 16 || n > 256 || k % 16 || a_rows < shift + 128) {
        ssb_set_error("tc_probe: bad shape");
        return -1;
    }
    const size_t smem = (size_t)(k / 8) * (a_rows + n) * 16 + 256;
    static const int key = ssb_new_key();
    if (ssb_first_on_device(key))
        SSB_CHECK_CUDA(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (smem > 200 * 1024) { ssb_set_error("tc_probe: operands exceed shared memory"); return -1; }
    tc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __half *)a_dev, a_rows, shift,
                                                            (const __half *)b_dev, n, k, d_dev, status_dev);
    SSB_CHECK_LAUNCH();
    return 0;
}

#endif  // SSB_BASELINES
