// Microbenchmark: TMEM read / write throughput of tcgen05.ld / tcgen05.st (32x32b shapes) with 4, 8 and 16 warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_bench tools/tmem_bench.cu && tools/tmem_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int MODE>   // 0 = ld x32, 1 = ld x16, 2 = st x32, 3 = ld x32 two in flight
__global__ void __launch_bounds__(512, 1) tmem_kernel(int iters, long long* cycles, float* sink) {
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t = tmem_base + ((static_cast<uint32_t>(warp & 3) * 32u) << 16) + static_cast<uint32_t>((warp >> 2) * 64);
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t a = t + static_cast<uint32_t>((it & 1) * 32);
        if (MODE == 0 || MODE == 3) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                  "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                  "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(a)
                : "memory");
            if (MODE == 0 || (it & 1)) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += __uint_as_float(v[0]) + __uint_as_float(v[31]);
        } else if (MODE == 1) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                           "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(a)
                         : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += __uint_as_float(v[0]) + __uint_as_float(v[15]);
        } else {
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(a),
                "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
                "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
                "r"(v[31])
                : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

template <int MODE>
void run(const char* name, int warps, int bytes_per_warp_iter) {
    long long* d_cyc; float* sink;
    cudaMalloc(&d_cyc, 8 * 148); cudaMalloc(&sink, 4 * 512 * 148);
    const int iters = 4096;
    tmem_kernel<MODE><<<148, warps * 32>>>(iters, d_cyc, sink);
    cudaDeviceSynchronize();
    tmem_kernel<MODE><<<148, warps * 32>>>(iters, d_cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps=%2d: %8.1f cycles/iter/warp  -> %7.1f B/clk/SM (%s)\n", name, warps, double(c) / iters,
           double(bytes_per_warp_iter) * warps * iters / double(c), cudaGetErrorString(e));
    cudaFree(d_cyc); cudaFree(sink);
}

int main() {
    for (int w : {4, 8, 16}) {
        run<0>("tcgen05.ld 32x32b.x32 (+wait)", w, 4096);
        run<3>("tcgen05.ld x32, 2 in flight", w, 4096);
        run<1>("tcgen05.ld 32x32b.x16 (+wait)", w, 2048);
        run<2>("tcgen05.st 32x32b.x32 (+wait)", w, 4096);
    }
    return 0;
}
