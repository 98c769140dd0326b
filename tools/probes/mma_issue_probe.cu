// Micro-benchmark: how long does the issuing thread spend per tcgen05.mma (M=128, K=16, N given),
// in the situations the OSBlock kernel has: divergent single-thread issue vs warp-uniform
// elect.sync issue, idle SM vs 15 busy warps, right after fence.proxy.async + barrier.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I strongsort-yolo_b200/csrc -o /tmp/mma_probe tools/probes/mma_issue_probe.cu && /tmp/mma_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"

constexpr int NMMA = 12;

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

template <int MODE, int BUSY, int N>
__global__ void __launch_bounds__(512, 1) probe(long long *out, float *sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar[NMMA + 1];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 64 * 1024 / 16; i += 512) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 0) tc::tmem_alloc(&s_tmem, 512);
    if (tid == 0) { for (int i = 0; i <= NMMA; i++) tc::mbar_init(bar + i, 1); tc::fence_mbar_init(); }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    constexpr uint32_t IDESC = tc::make_idesc_f16(128, N);
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    long long t[NMMA + 2];
    if (MODE == 0) {                    // divergent: one thread
        if (tid == 480) {
            const uint64_t a0 = tc::make_smem_desc(tc::smem_u32(smem), 12288, 128);
            const uint64_t b0 = tc::make_smem_desc(tc::smem_u32(smem) + 49152, 512, 128);
            t[0] = clock64();
#pragma unroll
            for (int i = 0; i < NMMA; i++) {
                tc::mma_f16_ss(tmem + (i & 3) * N, a0 + i * 128, b0, IDESC, 0);
                tc::mma_commit(bar + i);
                t[i + 1] = clock64();
            }
            tc::mbar_wait(bar + NMMA - 1, 0);
            t[NMMA + 1] = clock64();
            for (int i = 0; i < NMMA + 2; i++) out[i] = t[i] - t[0];
        }
    } else {                            // warp-uniform + elect
        if (warp_u == 15) {
            const uint64_t a0 = tc::make_smem_desc(tc::smem_u32(smem), 12288, 128);
            const uint64_t b0 = tc::make_smem_desc(tc::smem_u32(smem) + 49152, 512, 128);
            t[0] = clock64();
#pragma unroll
            for (int i = 0; i < NMMA; i++) {
                if (elect_one()) {
                    tc::mma_f16_ss(tmem + (i & 3) * N, a0 + i * 128, b0, IDESC, 0);
                    tc::mma_commit(bar + i);
                }
                __syncwarp();
                t[i + 1] = clock64();
            }
            tc::mbar_wait(bar + NMMA - 1, 0);
            t[NMMA + 1] = clock64();
            if (lane == 0) for (int i = 0; i < NMMA + 2; i++) out[i] = t[i] - t[0];
        }
    }
    if (BUSY && warp < 15) {            // the other warps grind FMAs + shared loads like the depthwise pass
        float a = tid, b = 1.0001f, c = 0.f;
        const float4 *p = reinterpret_cast<const float4 *>(smem) + tid;
        for (int i = 0; i < 600; i++) {
            float4 v = p[(i * 512) & 2047];
            a = fmaf(a, b, v.x); c = fmaf(c, b, v.y); a = fmaf(a, b, v.z); c = fmaf(c, b, v.w);
        }
        sink[tid] = a + c;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

template <int MODE, int BUSY, int N>
void run(const char *name) {
    long long *out; float *sink;
    cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 512 * 4);
    cudaFuncSetAttribute(probe<MODE, BUSY, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int r = 0; r < 3; r++) probe<MODE, BUSY, N><<<1, 512, 64 * 1024>>>(out, sink);
    long long h[64];
    cudaMemcpy(h, out, 64 * 8, cudaMemcpyDeviceToHost);
    printf("%-44s:", name);
    for (int i = 1; i <= NMMA; i++) printf(" %lld", h[i] - h[i - 1]);
    printf(" | all issued %lld, last done %lld  (%s)\n", h[NMMA], h[NMMA + 1], cudaGetErrorString(cudaGetLastError()));
}

int main() {
    run<0, 0, 32>("divergent thread, idle SM, N=32");
    run<0, 1, 32>("divergent thread, 15 busy warps, N=32");
    run<1, 0, 32>("uniform warp + elect, idle SM, N=32");
    run<1, 1, 32>("uniform warp + elect, 15 busy warps, N=32");
    run<0, 0, 64>("divergent thread, idle SM, N=64");
    run<1, 1, 64>("uniform warp + elect, 15 busy warps, N=64");
    return 0;
}
