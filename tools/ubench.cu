// Stand-alone sm_100a microbenchmarks that size the fused kernels (not part of the product library):
//   T1  tcgen05.mma SS (A, B in shared memory) issue rate vs N  -> is the shared-memory operand read the limiter?
//   T2  tcgen05.mma TS (A in TMEM) issue rate vs N
//   T3  L2 -> shared memory streaming rate of a weight-sized buffer by all SMs (cp.async.bulk)
//   T4  the same with 2-CTA clusters and .multicast::cluster (each CTA fetches half, both receive all)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../anomaly-detection-super-resolution_b200/csrc/ptx.cuh"

using namespace adsr;

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, long long limit = 200000000LL) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > limit) return false;
    }
    return true;
}

__device__ __forceinline__ uint32_t idesc_m128(uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

// ------------------------------------------------------------------------------------------------ T1 / T2
template <bool TS>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int iters, long long* cycles, int* err) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    // 4 stages of (A 16 KB + B 32 KB)
    for (int i = threadIdx.x; i < 4 * 48 * 1024 / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + ((i * 2654435761u) >> 28) * 0x00010001u;   // small bf16 values
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_slot);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_m128(N);
        // warm-up
        for (int s = 0; s < 4; ++s) {
            const uint64_t ad = umma_desc_k_sw128(smem_u32(smem + s * 49152));
            const uint64_t bd = umma_desc_k_sw128(smem_u32(smem + s * 49152 + 16384));
            for (int k = 0; k < 4; ++k) {
                if (TS) umma_ts(tmem, tmem + 256 + 8 * k, bd + 2 * k, idesc, 1);
                else umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, 1);
            }
        }
        umma_commit(&bar);
        if (!mbar_wait_bounded(&bar, 0)) { *err = 1; }
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int s = it & 3;
            const uint64_t ad = umma_desc_k_sw128(smem_u32(smem + s * 49152));
            const uint64_t bd = umma_desc_k_sw128(smem_u32(smem + s * 49152 + 16384));
            const uint32_t d = TS ? tmem : tmem + ((it >> 2) & 1) * 256;   // SS: alternate accumulators every 4 stages
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (TS) umma_ts(d, tmem + 256 + 8 * k, bd + 2 * k, idesc, 1);
                else umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, 1);
            }
        }
        umma_commit(&bar);
        if (!mbar_wait_bounded(&bar, 1)) { *err = 2; }
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem);
    }
}

// ------------------------------------------------------------------------------------------------ T3
// every CTA streams `buf` (buf_bytes, L2 resident after the first pass) `rounds` times into a ring of kSlots slots
template <int kSlots>
__global__ void __launch_bounds__(64, 1) l2_stream_kernel(const uint8_t* buf, int buf_bytes, int chunk, int rounds, int* err) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[kSlots], empty[kSlots];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kSlots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        fence_barrier_init();
    }
    __syncthreads();
    const int n_chunks = buf_bytes / chunk;
    const int total = n_chunks * rounds;
    if (threadIdx.x == 0) {          // producer
        int slot = 0; uint32_t ph = 0;
        // start each CTA at a different chunk so that the L2 slices are hit evenly
        int c = (blockIdx.x * 7) % n_chunks;
        for (int i = 0; i < total; ++i) {
            if (!mbar_wait_bounded(&empty[slot], ph ^ 1)) { *err = 3; return; }
            mbar_arrive_expect_tx(&full[slot], chunk);
            bulk_g2s(smem + slot * chunk, buf + static_cast<size_t>(c) * chunk, chunk, &full[slot]);
            if (++c == n_chunks) c = 0;
            if (++slot == kSlots) { slot = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {  // consumer: just releases the slot
        int slot = 0; uint32_t ph = 0;
        for (int i = 0; i < total; ++i) {
            if (!mbar_wait_bounded(&full[slot], ph)) { *err = 4; return; }
            mbar_arrive(&empty[slot]);
            if (++slot == kSlots) { slot = 0; ph ^= 1; }
        }
    }
}

// ------------------------------------------------------------------------------------------------ T4
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}

template <int kSlots>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1)
    l2_stream_mc_kernel(const uint8_t* buf, int buf_bytes, int chunk, int rounds, int* err) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[kSlots], empty[kSlots];
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) {
        for (int s = 0; s < kSlots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2); }   // empty: both CTAs' consumers
        fence_barrier_init();
    }
    cluster_sync_all();
    const int n_chunks = buf_bytes / chunk;
    const int total = n_chunks * rounds;
    const int half = chunk / 2;
    if (threadIdx.x == 0) {
        int slot = 0; uint32_t ph = 0;
        int c = ((blockIdx.x >> 1) * 7) % n_chunks;
        for (int i = 0; i < total; ++i) {
            if (!mbar_wait_bounded(&empty[slot], ph ^ 1)) { *err = 5; break; }
            mbar_arrive_expect_tx(&full[slot], chunk);            // my barrier receives both halves
            bulk_g2s_mc(smem + slot * chunk + rank * half, buf + static_cast<size_t>(c) * chunk + rank * half, half, &full[slot],
                        static_cast<uint16_t>(3));
            if (++c == n_chunks) c = 0;
            if (++slot == kSlots) { slot = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        int slot = 0; uint32_t ph = 0;
        for (int i = 0; i < total; ++i) {
            if (!mbar_wait_bounded(&full[slot], ph)) { *err = 6; break; }
            mbar_arrive_remote(&empty[slot], 0);
            mbar_arrive_remote(&empty[slot], 1);
            if (++slot == kSlots) { slot = 0; ph ^= 1; }
        }
    }
    cluster_sync_all();      // nobody exits while the peer may still write into its shared memory / barriers
}

int main(int argc, char** argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
    printf("device %s, %d SMs, clock %d MHz\n", prop.name, sms, clk_khz / 1000);

    long long* d_cycles;
    int* d_err;
    CK(cudaMalloc(&d_cycles, sizeof(long long) * 256));
    CK(cudaMalloc(&d_err, sizeof(int)));
    CK(cudaMemset(d_err, 0, sizeof(int)));
    std::vector<long long> h_cycles(256);
    int h_err = 0;

    const int smem_mma = 4 * 49152;
    CK(cudaFuncSetAttribute(mma_rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_mma));
    CK(cudaFuncSetAttribute(mma_rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_mma));
    const int iters = 2048;     // x4 MMAs each
    for (int ts = 0; ts < 2; ++ts) {
        for (int grid : {1, sms}) {
            for (int N : {32, 64, 96, 128, 160, 192, 256}) {
                cudaEvent_t e0, e1;
                CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                CK(cudaEventRecord(e0));
                if (ts) mma_rate_kernel<true><<<grid, 128, smem_mma>>>(N, iters, d_cycles, d_err);
                else mma_rate_kernel<false><<<grid, 128, smem_mma>>>(N, iters, d_cycles, d_err);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                CK(cudaMemcpy(h_cycles.data(), d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
                long long mx = 0, mn = 1LL << 60;
                for (int i = 0; i < grid; ++i) { mx = h_cycles[i] > mx ? h_cycles[i] : mx; mn = h_cycles[i] < mn ? h_cycles[i] : mn; }
                const double per = static_cast<double>(mx) / (iters * 4.0);
                const double flops = 2.0 * 128 * N * 16 * iters * 4.0 * grid;
                printf("T%d %s grid=%3d N=%3d : %7.1f cyc/MMA (min-CTA %7.1f)  floor %5.1f  ratio %.2f  kernel %.3f ms -> %.0f TFLOP/s  err=%d\n",
                       ts + 1, ts ? "TS" : "SS", grid, N, per, static_cast<double>(mn) / (iters * 4.0), N / 2.0, per / (N / 2.0), ms,
                       flops / (ms * 1e-3) / 1e12, h_err);
            }
        }
    }

    // ---- streaming
    const int buf_bytes = 384 * 1024;
    uint8_t* d_buf;
    CK(cudaMalloc(&d_buf, buf_bytes));
    CK(cudaMemset(d_buf, 1, buf_bytes));
    const int rounds = 48;
    for (int chunk : {8192, 16384, 32768}) {
        const int smem_bytes = 6 * chunk;
        CK(cudaFuncSetAttribute(l2_stream_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        CK(cudaFuncSetAttribute(l2_stream_mc_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        for (int mc = 0; mc < 2; ++mc) {
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEvent_t e0, e1;
                CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                CK(cudaEventRecord(e0));
                if (mc) l2_stream_mc_kernel<6><<<sms, 64, smem_bytes>>>(d_buf, buf_bytes, chunk, rounds, d_err);
                else l2_stream_kernel<6><<<sms, 64, smem_bytes>>>(d_buf, buf_bytes, chunk, rounds, d_err);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                best = ms < best ? ms : best;
            }
            CK(cudaMemcpy(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
            const double bytes = static_cast<double>(buf_bytes) * rounds * sms;      // bytes landed in shared memory
            printf("T%d %s chunk=%5d : %.3f ms  %.2f TB/s into smem (%.1f GB/s per SM)  err=%d\n", 3 + mc,
                   mc ? "multicast x2" : "unicast     ", chunk, best, bytes / (best * 1e-3) / 1e12, bytes / sms / (best * 1e-3) / 1e9,
                   h_err);
        }
    }
    return 0;
}
