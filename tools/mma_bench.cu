// Microbenchmark: cost of tcgen05.mma (kind::f16, bf16, M = 128, K = 16, cta_group::1) per instruction as a function of N, for
// the SS (A in shared memory) and TS (A in TMEM) forms: cycles the issuing thread spends per MMA, and cycles until the batch has
// completed (tcgen05.commit -> mbarrier).  Operands are zeros; 64 MMAs per batch accumulate into one TMEM region.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I anomaly-detection-super-resolution_b200/csrc -o tools/mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"

using namespace adsr;

__global__ void __launch_bounds__(128, 1) mma_kernel(int n, int ts, int count, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem));            // [128 rows x 64 bf16]
        const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem + 16384));    // [<= 256 rows x 64 bf16]
        const uint32_t idesc = umma_idesc_bf16_m128(static_cast<uint32_t>(n));
        long long t_issue = 0, t_done = 0;
        for (int r = 0; r < reps; ++r) {
            __syncwarp();
            const long long t0 = clock64();
            if (elect_one_sync()) {
                for (int i = 0; i < count; ++i) {
                    const uint32_t k = static_cast<uint32_t>(i & 3) * 2;      // K step inside the 64-wide panel
                    if (ts) umma_bf16_ts(tmem, tmem + 256 + 16 * (i & 3), bdesc + k, idesc, i ? 1u : 0u);
                    else umma_bf16(tmem, adesc + k, bdesc + k, idesc, i ? 1u : 0u);
                }
                umma_commit(&bar);
            }
            __syncwarp();
            const long long t1 = clock64();
            mbar_wait(&bar, static_cast<uint32_t>(r) & 1);
            const long long t2 = clock64();
            if (r > 0) { t_issue += t1 - t0; t_done += t2 - t0; }
        }
        if (threadIdx.x == 0) { out[0] = t_issue / (reps - 1); out[1] = t_done / (reps - 1); }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem);
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768);
    printf("tcgen05.mma kind::f16 bf16 M=128 K=16, 64 MMAs per batch (one CTA, zeros)\n");
    for (int ts = 0; ts < 2; ++ts)
        for (int n : {16, 32, 64, 96, 128, 192, 256}) {
            mma_kernel<<<1, 128, 16384 + 32768>>>(n, ts, 64, 9, d);
            long long h[2];
            if (cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            printf("%s N=%3d: issue %6.1f clk/MMA   complete %6.1f clk/MMA   (ideal %5.1f at 8192 flop/clk)\n", ts ? "TS" : "SS", n, h[0] / 64.0,
                   h[1] / 64.0, 128.0 * n * 16 * 2 / 8192.0);
        }
    // short batches: 4 MMAs + commit, the pattern of one K slab
    for (int n : {32, 128}) {
        mma_kernel<<<1, 128, 16384 + 32768>>>(n, 0, 4, 9, d);
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("SS N=%3d, batch of 4: issue %6.1f clk total   complete %6.1f clk total\n", n, (double)h[0], (double)h[1]);
    }
    return 0;
}
