// tcgen05 / TMEM GEMM and implicit-GEMM 3x3 convolution for sm_100a.
//
// One persistent, warp-specialised kernel.  Per 128-row output tile the K dimension is streamed
// through a 4-deep shared-memory ring; each ring stage holds
//   * an A tile  [128 rows x 64 bf16]  (K-major, 128-byte swizzle) written by 4 producer warps that
//     either copy token rows (GEMM) or gather the 3x3 taps of an NHWC image with zero padding
//     (implicit GEMM: the im2col matrix is never materialised), and
//   * a  B tile  [BN rows  x 64 bf16]  = pre-swizzled weight image fetched with ONE bulk async copy
//     (cp.async.bulk, the TMA engine) -- weights are packed once at load time into exactly the
//     shared-memory image the tensor core wants, so no tensor map is needed.
// A single elected thread issues tcgen05.mma (M=128, N=BN<=256, K=16) into one of two TMEM
// accumulator buffers; 4 epilogue warps drain the other buffer with tcgen05.ld and apply
// bias / activation / residual / pixel-shuffle before storing bf16.
//
// Replaces (reference call sites): nn.Linear qkv/proj/fc1/fc2 (src/drct.py:278,300,185-188), the 1x1
// adjust convs (src/drct.py:334-374, 389-393), conv_after_body / conv_before_upsample / Upsample
// convs + PixelShuffle (src/drct.py:837,844-845,702-705) and every 3x3 conv of DRN (src/drn.py:29-32).
#include "adsr_kernels.h"
#include "ptx.cuh"

namespace adsr {

namespace {

constexpr int kStages = 4;
constexpr int kAStageBytes = 128 * 128;       // 128 rows x 64 bf16
constexpr int kBStageBytes = 256 * 128;       // up to 256 rows x 64 bf16
constexpr int kNumThreads = 384;              // 12 warps
constexpr int kTmemCols = 512;                // 2 accumulator buffers x 256 fp32 columns
constexpr int kSmemBytes = kStages * (kAStageBytes + kBStageBytes) + 256;

struct __align__(8) RingBarriers {
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == ADSR_ACT_LRELU) return v > 0.f ? v : v * slope;
    if (act == ADSR_ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    if (act == ADSR_ACT_RELU) return fmaxf(v, 0.f);
    return v;
}

__global__ void __launch_bounds__(kNumThreads, 1) tc_gemm_kernel(const TcGemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * kAStageBytes;
    RingBarriers* bars = reinterpret_cast<RingBarriers*>(smem + kStages * (kAStageBytes + kBStageBytes));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.m_tiles * p.n_tiles;

    if ((smem_u32(smem) & 1023u) != 0) __trap();   // swizzle-128B operands need 1024 B alignment

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bars->full[s], 128 + 1);    // 128 producer threads + the bulk-copy issuer
            mbar_init(&bars->empty[s], 1);         // one tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->tmem_full[b], 1);
            mbar_init(&bars->tmem_empty[b], 128);  // 128 epilogue threads
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<kTmemCols>(&bars->tmem_base);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ================================ B loader: one bulk copy per ring stage =================
        if (lane == 0) {
            const uint32_t b_bytes = static_cast<uint32_t>(p.BN) * 128u;
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n_tile = tile % p.n_tiles;
                const uint8_t* src = p.Bp + static_cast<size_t>(n_tile) * p.num_k_stages * b_bytes;
                for (int ks = 0; ks < p.num_k_stages; ++ks) {
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&bars->full[stage], b_bytes);
                    bulk_g2s(smem_b + stage * kBStageBytes, src + static_cast<size_t>(ks) * b_bytes, b_bytes,
                             &bars->full[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (single thread) ================================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16_m128(static_cast<uint32_t>(p.BN));
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t use = static_cast<uint32_t>(it >> 1);
                mbar_wait(&bars->tmem_empty[buf], (use & 1) ^ 1);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * 256);
                for (int ks = 0; ks < p.num_k_stages; ++ks) {
                    int steps;
                    if (p.conv) {
                        steps = p.k16_per_tap - 4 * (ks % p.stages_per_tap);
                    } else {
                        steps = p.k16_total - 4 * ks;
                    }
                    steps = steps > 4 ? 4 : steps;
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after_sync();
                    const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem_a + stage * kAStageBytes));
                    const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem_b + stage * kBStageBytes));
                    for (int k = 0; k < steps; ++k) {
                        // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in 16-byte units
                        umma_bf16(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                                  (ks | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&bars->empty[stage]);          // frees the ring slot when the MMAs retire
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&bars->tmem_full[buf]);            // accumulator complete -> epilogue
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ================================ epilogue: TMEM -> registers -> global ======================
        const int quad = warp & 3;                             // TMEM lane quadrant this warp may read
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int m_tile = tile / p.n_tiles;
            const int n_tile = tile % p.n_tiles;
            const int buf = it & 1;
            const uint32_t use = static_cast<uint32_t>(it >> 1);
            mbar_wait(&bars->tmem_full[buf], use & 1);
            tc_fence_after_sync();
            const int row = m_tile * 128 + quad * 32 + lane;
            const bool row_ok = row < p.M;
            const uint32_t taddr = tmem_base + static_cast<uint32_t>(buf * 256) + (static_cast<uint32_t>(quad * 32) << 16);
            const int n_base = n_tile * p.BN;

            // pixel-shuffle destination (out_mode 1): this row is input pixel (b, y, x)
            long long ps_base = 0;
            if (p.out_mode == ADSR_OUT_PIXEL_SHUFFLE2 && row_ok) {
                const int hw = p.Hout * p.Wout;
                const int b = row / hw;
                const int rem = row - b * hw;
                const int y = rem / p.Wout;
                const int x = rem - y * p.Wout;
                ps_base = ((static_cast<long long>(b) * (2 * p.Hout) + 2 * y) * (2 * p.Wout) + 2 * x) * p.ldo;
            }

            for (int c0 = 0; c0 < p.BN; c0 += 16) {
                uint32_t raw[16];
                __syncwarp();                                   // tcgen05.ld is .sync.aligned: reconverge first
                tmem_ld16(taddr + static_cast<uint32_t>(c0), raw);
                tmem_ld_wait();
                const int n0 = n_base + c0;
                if (row_ok && n0 < p.n_store) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float a = __uint_as_float(raw[j]) + __ldg(p.bias + n0 + j);
                    v[j] = apply_act(a, p.act, p.slope) * p.alpha;
                }
                if (p.res != nullptr) {
                    const uint4* rp = reinterpret_cast<const uint4*>(p.res + static_cast<long long>(row) * p.ldres + n0);
                    uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                    const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (n0 + 2 * j < p.N) v[2 * j] += bf16_lo(rr[j]);
                        if (n0 + 2 * j + 1 < p.N) v[2 * j + 1] += bf16_hi(rr[j]);
                    }
                }
                if (p.out_mode == ADSR_OUT_ROWS) {
                    uint32_t o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
                    __nv_bfloat16* dst = p.out + static_cast<long long>(row) * p.ldo + p.ocol0 + n0;
                    if ((p.ocol0 & 7) == 0) {
                        uint4* d4 = reinterpret_cast<uint4*>(dst);
                        d4[0] = make_uint4(o[0], o[1], o[2], o[3]);
                        d4[1] = make_uint4(o[4], o[5], o[6], o[7]);
                    } else {                                    // slab slices start at 8-byte aligned columns
                        uint2* d2 = reinterpret_cast<uint2*>(dst);
                        d2[0] = make_uint2(o[0], o[1]);
                        d2[1] = make_uint2(o[2], o[3]);
                        d2[2] = make_uint2(o[4], o[5]);
                        d2[3] = make_uint2(o[6], o[7]);
                    }
                } else {
                    // PixelShuffle(2): column n = c*4 + i*2 + j  ->  out[b, 2y+i, 2x+j, c]
                    const int ch0 = n0 >> 2;
#pragma unroll
                    for (int sub = 0; sub < 4; ++sub) {
                        const int i = sub >> 1, j = sub & 1;
                        uint2 o;
                        o.x = pack_bf16x2(v[0 * 4 + sub], v[1 * 4 + sub]);
                        o.y = pack_bf16x2(v[2 * 4 + sub], v[3 * 4 + sub]);
                        __nv_bfloat16* dst = p.out + ps_base + (static_cast<long long>(i) * (2 * p.Wout) + j) * p.ldo + ch0;
                        *reinterpret_cast<uint2*>(dst) = o;
                    }
                }
                }
            }
            __syncwarp();
            tc_fence_before_sync();
            mbar_arrive(&bars->tmem_empty[buf]);
        }
    } else if (warp >= 8) {
        // ================================ A producers (4 warps, 128 threads) =========================
        const int pw = warp - 8;
        const int chunk = lane & 7;                            // 16-byte chunk inside the 128 B row
        const int rsub = pw * 4 + (lane >> 3);                 // row inside each 16-row step
        const uint32_t sw_off = static_cast<uint32_t>(rsub * 128 + ((chunk ^ (rsub & 7)) << 4));
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_tile = tile / p.n_tiles;
            const int m0 = m_tile * 128;
            // per-row source bookkeeping (8 rows per thread: r = step*16 + rsub)
            long long row_off[8];
            int oy[8], ox[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const int m = m0 + s * 16 + rsub;
                if (m >= p.M) {
                    row_off[s] = -1; oy[s] = 0; ox[s] = 0;
                } else if (p.conv) {
                    const int hw = p.Hout * p.Wout;
                    const int b = m / hw;
                    const int rem = m - b * hw;
                    oy[s] = rem / p.Wout;
                    ox[s] = rem - oy[s] * p.Wout;
                    row_off[s] = static_cast<long long>(b) * p.Hin * p.Win;   // pixel index of image start
                } else {
                    row_off[s] = static_cast<long long>(m) * p.lda; oy[s] = 0; ox[s] = 0;
                }
            }
            for (int ks = 0; ks < p.num_k_stages; ++ks) {
                uint4 v[8];
                if (p.conv) {
                    const int tap = ks / p.stages_per_tap;
                    const int col = (ks - tap * p.stages_per_tap) * 64 + chunk * 8;
                    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                    const bool col_ok = col < p.K8;
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        const int iy = oy[s] * p.stride + dy, ix = ox[s] * p.stride + dx;
                        const bool ok = col_ok && row_off[s] >= 0 && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
                        if (ok) {
                            const long long pix = row_off[s] + static_cast<long long>(iy) * p.Win + ix;
                            v[s] = __ldg(reinterpret_cast<const uint4*>(p.A + pix * p.lda + col));
                        } else {
                            v[s] = make_uint4(0, 0, 0, 0);
                        }
                    }
                } else {
                    const int col = ks * 64 + chunk * 8;
                    const bool col_ok = col < p.K8;
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        if (col_ok && row_off[s] >= 0) {
                            v[s] = __ldg(reinterpret_cast<const uint4*>(p.A + row_off[s] + col));
                        } else {
                            v[s] = make_uint4(0, 0, 0, 0);
                        }
                    }
                }
                mbar_wait(&bars->empty[stage], phase ^ 1);
                uint8_t* dst = smem_a + stage * kAStageBytes + sw_off;
#pragma unroll
                for (int s = 0; s < 8; ++s) *reinterpret_cast<uint4*>(dst + s * 16 * 128) = v[s];
                fence_proxy_async_smem();
                mbar_arrive(&bars->full[stage]);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

}  // namespace

int launch_tc_gemm(const TcGemmParams& p, int num_sms, cudaStream_t stream) {
    if (p.M <= 0) return ADSR_OK;
    if (p.BN < 16 || p.BN > 256 || (p.BN % 16) != 0) return ADSR_ERR_BAD_SHAPE;
    if ((p.lda % 8) != 0 || (p.ldo % 4) != 0 || (p.ocol0 % 4) != 0) return ADSR_ERR_BAD_SHAPE;
    if ((reinterpret_cast<uintptr_t>(p.A) & 15) || (reinterpret_cast<uintptr_t>(p.Bp) & 15)) return ADSR_ERR_BAD_ALIGN;
    if (p.res != nullptr && ((p.ldres % 8) != 0 || (reinterpret_cast<uintptr_t>(p.res) & 15))) return ADSR_ERR_BAD_ALIGN;
    if (p.out_mode == ADSR_OUT_ROWS && (p.ocol0 % 8) == 0 &&
        ((reinterpret_cast<uintptr_t>(p.out) & 15) || (p.ldo % 8) != 0))
        return ADSR_ERR_BAD_ALIGN;
    if (p.num_k_stages <= 0 || p.n_tiles <= 0) return ADSR_ERR_BAD_SHAPE;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return ADSR_ERR_CUDA;
        attr_set = true;
    }
    const int tiles = p.m_tiles * p.n_tiles;
    const int grid = tiles < num_sms ? tiles : num_sms;
    tc_gemm_kernel<<<grid, kNumThreads, kSmemBytes, stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? ADSR_OK : ADSR_ERR_LAUNCH;
}

}  // namespace adsr
