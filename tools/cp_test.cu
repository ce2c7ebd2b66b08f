// Probe of tcgen05.cp (shared memory -> TMEM) as a way to stage an MMA A operand: copies the K = 16 slices of a [128 rows x 64 bf16]
// K-major, 128-byte-swizzled panel (the layout of every A tile in this repo) with the .128x256b shape, reads them back with
// tcgen05.ld and compares with the source; then runs the SAME GEMM once with A from shared memory (SS) and once with A from TMEM
// (TS) and compares the accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I anomaly-detection-super-resolution_b200/csrc -o tools/cp_test tools/cp_test.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "ptx.cuh"

using namespace adsr;

__global__ void __launch_bounds__(128, 1) cp_kernel(int* mismatches, float* max_diff) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* a_s = smem;              // [128 x 64] bf16, SW128
    uint8_t* b_s = smem + 16384;      // [128 x 64] bf16, SW128 (N = 128 rows of B)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto val_a = [](int r, int k) { return static_cast<float>(((r * 7 + k * 3) % 31) - 15); };
    auto val_b = [](int n, int k) { return static_cast<float>(((n * 5 + k * 11) % 29) - 14) * 0.125f; };
    for (int i = tid; i < 128 * 64; i += 128) {
        const int r = i >> 6, k = i & 63;
        const int off = r * 128 + ((((k >> 3) ^ (r & 7)) << 4) | ((k & 7) << 1));
        *reinterpret_cast<__nv_bfloat16*>(a_s + off) = __float2bfloat16(val_a(r, k));
        *reinterpret_cast<__nv_bfloat16*>(b_s + off) = __float2bfloat16(val_b(r, k));
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint64_t adesc = umma_desc_k_sw128(smem_u32(a_s));
    const uint64_t bdesc = umma_desc_k_sw128(smem_u32(b_s));
    const uint32_t idesc = umma_idesc_bf16_m128(128u);
    uint32_t phase = 0;
    // ---- 1. copy the four K16 slices to TMEM columns 256 + 8 s
    if (warp == 0) {
        if (elect_one_sync()) {
            for (int s = 0; s < 4; ++s) tmem_cp_128x256b(tmem + 256 + 8 * s, adesc + 2 * s);
            umma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, phase); phase ^= 1;
    tc_fence_after_sync();
    int bad = 0;
    for (int s = 0; s < 4; ++s) {
        uint32_t v[8];
        tmem_ld8(tmem + ((static_cast<uint32_t>(warp * 32)) << 16) + 256 + 8 * s, v);
        tmem_ld_wait();
        const int r = warp * 32 + lane;
        for (int e = 0; e < 16; ++e) {
            const uint32_t w = v[e >> 1];
            const float got = __bfloat162float(__ushort_as_bfloat16(static_cast<unsigned short>((e & 1) ? (w >> 16) : (w & 0xffffu))));
            if (got != val_a(r, 16 * s + e)) ++bad;
        }
    }
    atomicAdd(mismatches, bad);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    // ---- 2. D_ss (columns 0..127) = A B^T with A from shared memory; D_ts (columns 128..255) with A from TMEM
    if (warp == 0) {
        if (elect_one_sync()) {
            for (int s = 0; s < 4; ++s) umma_bf16(tmem, adesc + 2 * s, bdesc + 2 * s, idesc, s ? 1u : 0u);
            for (int s = 0; s < 4; ++s) umma_bf16_ts(tmem + 128, tmem + 256 + 8 * s, bdesc + 2 * s, idesc, s ? 1u : 0u);
            umma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, phase); phase ^= 1;
    tc_fence_after_sync();
    float md = 0.f;
    for (int c = 0; c < 128; c += 8) {
        uint32_t a[8], b[8];
        const uint32_t base = tmem + ((static_cast<uint32_t>(warp * 32)) << 16);
        tmem_ld8(base + c, a);
        tmem_ld8(base + 128 + c, b);
        tmem_ld_wait();
        const int r = warp * 32 + lane;
        for (int e = 0; e < 8; ++e) {
            float ref = 0.f;
            for (int k = 0; k < 64; ++k) ref += val_a(r, k) * val_b(c + e, k);
            md = fmaxf(md, fabsf(__uint_as_float(a[e]) - ref));
            md = fmaxf(md, fabsf(__uint_as_float(b[e]) - ref));
        }
    }
    atomicMax(reinterpret_cast<int*>(max_diff), __float_as_int(md));
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) { tc_fence_after_sync(); tmem_dealloc<512>(tmem); }
}

int main() {
    int* d_bad; float* d_md;
    cudaMalloc(&d_bad, 4); cudaMalloc(&d_md, 4);
    cudaMemset(d_bad, 0, 4); cudaMemset(d_md, 0, 4);
    cudaFuncSetAttribute(cp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    cp_kernel<<<1, 128, 32768>>>(d_bad, d_md);
    int bad = -1; float md = -1.f;
    const cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&md, d_md, 4, cudaMemcpyDeviceToHost);
    printf("tcgen05.cp 128x256b from a K-major SW128 panel: %s, %d mismatching elements of 8192; SS / TS GEMM max |diff| vs exact = %g\n",
           cudaGetErrorString(e), bad, md);
    return 0;
}
