// 3x3 convolution (stride 1, pad 1) on NHWC bf16 for narrow layers whose weights fit in shared memory -- DRN-L's RCAB convs
// (80 -> 80 channels, src/drn.py:143-158): "halo tile" implicit GEMM.
//
// The TMA-fed implicit GEMM of tc_gemm_manual.cu streams, for every 128-pixel tile, nine shifted copies of its input and the
// whole weight set from L2 (~364 KB per tile at 80 channels): it is bound by L2 -> SM bandwidth.  Here
//   * the weights (K-concatenated compact image: K index = tap * 16 ceil(Cin / 16) + channel) are loaded ONCE per CTA and stay
//     resident in shared memory;
//   * output positions are enumerated in the PADDED raster of the image (row pitch Wh = W + 2: one pad column on each side), a
//     tile = 128 consecutive positions; its input halo is a contiguous range of padded rows, fetched by ONE 4-D TMA box per
//     64-channel panel ([64 ch x Wh x rows], zero fill outside the image) -- each input pixel is read from L2 ~1.3 x instead of 9 x;
//   * tap (dy, dx) of the GEMM is the SAME shared-memory panel read through an A descriptor whose start is shifted by
//     dy * Wh + dx rows (K-major, 128-byte swizzle: the tensor core derives the swizzle from the shared-memory address, so any
//     row offset is consistent with what the TMA unit wrote -- measured: the descriptor's base-offset field must stay 0);
//   * outputs at pad columns are computed and discarded (W / Wh = 94 - 97 % of the MMA rows are useful).
// One MMA warp issues the 9 * ceil(Cin / 16) MMAs of a tile back to back (no per-stage hand-shakes), 4 epilogue warps turn the
// double-buffered TMEM accumulator into bf16 rows (+ bias, ReLU / LeakyReLU).
#include "adsr_kernels.h"
#include "ptx.cuh"

#define HALO_TRACE 0   // 1: CTA 0 records a clock64 timeline per tile (tools/halo_trace.py)
namespace adsr {
#if HALO_TRACE
__device__ long long g_halo_trace[3 * 64 * 2];
#define TR(role, it, k) do { if (blockIdx.x == 0 && (it) < 64 && lane == 0) g_halo_trace[((role) * 64 + (it)) * 2 + (k)] = clock64(); } while (0)
#else
#define TR(role, it, k) do {} while (0)
#endif

namespace {

constexpr int kThreads = 12 * 32;             // warps: 0 loader, 1 MMA, 2 TMEM alloc, 3 spare, 4..11 epilogue (two per TMEM lane quadrant)
constexpr int kGuardBytes = 1024;             // the most negative tap shift of a tile's first position is row -1
constexpr int kSmemLimit = 232448;

struct __align__(8) HaloBarriers {
    uint64_t w_full, a_full[2], a_empty[2];
    uint64_t acc_full[2], acc_free[2];
    uint32_t tmem_base;
};

// K-major operand rows of 16 bf16 = 32 bytes, 32-byte swizzle (8-row groups 256 B apart): the tail panel
__device__ __forceinline__ uint64_t umma_desc_k_sw32(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(256 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(6) << 61;                       // SWIZZLE_32B
    return d;
}

// The 9 * K16 MMAs of one tile as straight-line code: the descriptors differ from (a0, b0) only in the 14-bit start-address field
// (16-byte units), by offsets that are compile-time constants times a few launch constants.  A rolled loop with run-time
// descriptor arithmetic costs ~120 cycles of the issuing thread per MMA (measured), twice the shared-memory time of an N = 80 MMA.
// TAIL: the last K = 16 step of every tap reads the 32-byte-swizzled tail panel (descriptor t0, rows 2 units apart).
template <int K16, bool TAIL>
__device__ __forceinline__ void issue_tile(uint64_t a0, uint64_t t0, int wh, uint32_t pan_u, uint64_t b0, uint32_t slab_u, uint32_t d,
                                           uint32_t idesc) {
    const uint32_t a_hi = static_cast<uint32_t>(a0 >> 32), a_lo = static_cast<uint32_t>(a0);
    const uint32_t t_hi = static_cast<uint32_t>(t0 >> 32), t_lo = static_cast<uint32_t>(t0);
    const uint32_t b_hi = static_cast<uint32_t>(b0 >> 32), b_lo = static_cast<uint32_t>(b0);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        const int shift = dy * wh + dx;                          // rows
#pragma unroll
        for (int c = 0; c < K16; ++c) {
            const int idx = tap * K16 + c;
            const uint32_t bl = b_lo + static_cast<uint32_t>(idx >> 2) * slab_u + static_cast<uint32_t>((idx & 3) * 2);
            const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | bl;
            if (TAIL && c == K16 - 1) {
                umma_bf16(d, (static_cast<uint64_t>(t_hi) << 32) | (t_lo + static_cast<uint32_t>(shift * 2)), bd, idesc, idx != 0 ? 1u : 0u);
            } else {
                const uint32_t al = a_lo + static_cast<uint32_t>(shift * 8) + static_cast<uint32_t>(c >> 2) * pan_u + static_cast<uint32_t>((c & 3) * 2);
                umma_bf16(d, (static_cast<uint64_t>(a_hi) << 32) | al, bd, idesc, idx != 0 ? 1u : 0u);
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads, 1) conv_halo_kernel(const __grid_constant__ ConvHaloParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_s = smem;                                                         // slabs x [BN x 64] resident weights
    uint8_t* a_s = w_s + p.slabs * p.BN * 128 + kGuardBytes;                      // n_abuf x full_panels x panel_bytes (128-byte swizzle)
    const int full_bytes = p.full_panels * p.panel_bytes;
    uint8_t* t_s = a_s + p.n_abuf * full_bytes;                                   // n_abuf x tail_bytes (32-byte swizzle)
    float* s_bias = reinterpret_cast<float*>(t_s + p.n_abuf * p.tail_bytes);      // [128]
    HaloBarriers* bars = reinterpret_cast<HaloBarriers*>(s_bias + 128);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int my_tiles = static_cast<int>(blockIdx.x) < p.n_tiles ? (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int ab_mask = p.n_abuf - 1, ab_shift = p.n_abuf - 1;                     // n_abuf is 1 or 2

    if ((smem_u32(smem) & 1023u) != 0) __trap();
    for (int i = threadIdx.x; i < 128; i += kThreads) s_bias[i] = i < p.BN ? p.bias[i] : 0.f;
    if (warp == 1 && lane == 0) {
        mbar_init(&bars->w_full, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->a_full[b], 1);
            mbar_init(&bars->a_empty[b], 1);
            mbar_init(&bars->acc_full[b], 1);
            mbar_init(&bars->acc_free[b], 256);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<256>(&bars->tmem_base);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = bars->tmem_base;
    pdl_launch_dependents();                                            // the next kernel's prologue may overlap my tail

    if (warp == 0) {
        // ============================================================ loader: weights once, then one halo tile per tile
        if (lane == 0) {
            tma_prefetch_desc(&p.tmap_in);
            if (p.tail) tma_prefetch_desc(&p.tmap_tail);
        }
        if (elect_one_sync()) {
            const uint32_t slab_bytes = static_cast<uint32_t>(p.BN) * 128u;
            mbar_arrive_expect_tx(&bars->w_full, static_cast<uint32_t>(p.slabs) * slab_bytes);
            for (int i = 0; i < p.slabs; ++i) bulk_g2s(w_s + static_cast<size_t>(i) * slab_bytes, p.wp + static_cast<size_t>(i) * slab_bytes, slab_bytes, &bars->w_full);
        }
        __syncwarp();
        pdl_wait();                                                     // the input image is the previous kernel's output
        const uint32_t tx_bytes = static_cast<uint32_t>(p.full_panels * p.box_bytes + p.tail * p.tail_box_bytes);
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = it * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x);
            const int b = tile / p.tiles_per_img, t = tile - b * p.tiles_per_img;
            const int q0 = p.Wh + 128 * t;                              // first output position of the tile (padded raster)
            const int pr0 = q0 / p.Wh - 1;                              // first padded row of the halo box; image row = pr0 - 1
            const int s = it & ab_mask;
            mbar_wait(&bars->a_empty[s], (static_cast<uint32_t>(it >> ab_shift) & 1) ^ 1);
            TR(0, it, 0);
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(&bars->a_full[s], tx_bytes);
                for (int pn = 0; pn < p.full_panels; ++pn)
                    tma_load_4d(a_s + s * full_bytes + pn * p.panel_bytes, &p.tmap_in, pn * 64, -1, pr0 - 1, b, &bars->a_full[s]);
                if (p.tail) tma_load_4d(t_s + s * p.tail_bytes, &p.tmap_tail, p.full_panels * 64, -1, pr0 - 1, b, &bars->a_full[s]);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer: 9 * k16 MMAs per tile, back to back
        const uint32_t idesc = umma_idesc_bf16_m128(static_cast<uint32_t>(p.BN));
        const uint32_t a_base = smem_u32(a_s), t_base = smem_u32(t_s);
        const uint64_t b0 = umma_desc_k_sw128(smem_u32(w_s));
        const uint32_t slab_u = static_cast<uint32_t>(p.BN) * 8u, pan_u = static_cast<uint32_t>(p.panel_bytes) >> 4;   // 16-byte units
        mbar_wait(&bars->w_full, 0);
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = it * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x);
            const int b = tile / p.tiles_per_img, t = tile - b * p.tiles_per_img;
            const int q0 = p.Wh + 128 * t;
            const int pr0 = q0 / p.Wh - 1;
            const int row0 = q0 - pr0 * p.Wh;                           // shared-memory row of the tile's first position (tap 0, 0)
            const int buf = it & 1, s = it & ab_mask;
            (void)b;
            mbar_wait(&bars->acc_free[buf], (static_cast<uint32_t>(it >> 1) & 1) ^ 1);
            TR(1, it, 0);
            mbar_wait(&bars->a_full[s], static_cast<uint32_t>(it >> ab_shift) & 1);
            TR(0, it, 1);
            tc_fence_after_sync();
            if (elect_one_sync()) {
                const uint32_t d = tmem + static_cast<uint32_t>(buf * 128);
                const uint64_t a0 = umma_desc_k_sw128(a_base + static_cast<uint32_t>(s * full_bytes + row0 * 128));
                const uint64_t t0 = umma_desc_k_sw32(t_base + static_cast<uint32_t>(s * p.tail_bytes + row0 * 32));
                if (p.tail) {
                    issue_tile<5, true>(a0, t0, p.Wh, pan_u, b0, slab_u, d, idesc);
                } else {
                    switch (p.k16_per_tap) {
                        case 1: issue_tile<1, false>(a0, t0, p.Wh, pan_u, b0, slab_u, d, idesc); break;
                        case 2: issue_tile<2, false>(a0, t0, p.Wh, pan_u, b0, slab_u, d, idesc); break;
                        case 3: issue_tile<3, false>(a0, t0, p.Wh, pan_u, b0, slab_u, d, idesc); break;
                        case 4: issue_tile<4, false>(a0, t0, p.Wh, pan_u, b0, slab_u, d, idesc); break;
                        case 5: issue_tile<5, false>(a0, t0, p.Wh, pan_u, b0, slab_u, d, idesc); break;
                        case 6: issue_tile<6, false>(a0, t0, p.Wh, pan_u, b0, slab_u, d, idesc); break;
                        case 7: issue_tile<7, false>(a0, t0, p.Wh, pan_u, b0, slab_u, d, idesc); break;
                        default: issue_tile<8, false>(a0, t0, p.Wh, pan_u, b0, slab_u, d, idesc); break;
                    }
                }
                umma_commit(&bars->a_empty[s]);
                umma_commit(&bars->acc_full[buf]);
            }
            __syncwarp();
            TR(1, it, 1);
        }
    } else if (warp >= 4) {
        // ============================================================ epilogue: thread = tile row = padded-raster position
        const int quad = warp & 3;
        const int i = quad * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const int units = p.BN >> 4;
        const int u_begin = warp < 8 ? 0 : (units + 1) >> 1, u_end = warp < 8 ? (units + 1) >> 1 : units;   // column halves
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = it * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x);
            const int b = tile / p.tiles_per_img, t = tile - b * p.tiles_per_img;
            const int q = p.Wh + 128 * t + i;
            const int pr = q / p.Wh, xc = q - pr * p.Wh;                // padded row / column of my position
            const bool valid = xc >= 1 && xc <= p.W && pr >= 1 && pr <= p.H;
            __nv_bfloat16* dst = p.out + (static_cast<long long>(b) * p.H * p.W + static_cast<long long>(pr - 1) * p.W + (xc - 1)) * p.ldo + p.ocol0;
            const int buf = it & 1;
            mbar_wait(&bars->acc_full[buf], static_cast<uint32_t>(it >> 1) & 1);
            if (warp == 4) TR(2, it, 0);
            tc_fence_after_sync();
            const uint32_t taddr = tmem + lane_off + static_cast<uint32_t>(buf * 128);
            for (int u = u_begin; u < u_end; ++u) {
                uint32_t raw[16];
                tmem_ld16(taddr + static_cast<uint32_t>(16 * u), raw);
                tmem_ld_wait();
                float f[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    float v = __uint_as_float(raw[e]) + s_bias[16 * u + e];
                    if (p.act == ADSR_ACT_RELU) v = fmaxf(v, 0.f);
                    else if (p.act == ADSR_ACT_LRELU) v = v > 0.f ? v : v * p.slope;
                    f[e] = v;
                }
                uint32_t pk[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) pk[e] = pack_bf16x2(f[2 * e], f[2 * e + 1]);
                if (p.chan_part != nullptr) {
                    // column sums over the 32 rows of this warp (pad / out-of-image positions count as 0): butterfly that halves the
                    // values per lane at every step (16 shuffles per unit); lane l ends with column 16 u + (l >> 1)
#pragma unroll
                    for (int e = 0; e < 16; ++e) f[e] = valid ? f[e] : 0.f;
#pragma unroll
                    for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
                        const bool up = (lane & bit) != 0;
#pragma unroll
                        for (int j = 0; j < w; ++j) {
                            const float send = up ? f[j] : f[j + w];
                            const float keep = up ? f[j + w] : f[j];
                            f[j] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
                        }
                    }
                    f[0] += __shfl_xor_sync(0xffffffffu, f[0], 1);
                    if ((lane & 1) == 0) p.chan_part[(static_cast<long long>(tile) * 4 + quad) * p.BN + 16 * u + (lane >> 1)] = f[0];
                }
                if (valid && 16 * u < p.n_store) {
                    uint4* d4 = reinterpret_cast<uint4*>(dst + 16 * u);
                    d4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    if (16 * u + 8 < p.n_store) d4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&bars->acc_free[buf]);
            if (warp == 4) TR(2, it, 1);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc<256>(tmem);
    }
}

}  // namespace

// chan_part (optional): [n_tiles * 4, BN] fp32 per-(tile, row quadrant) column sums of the outputs -- a tile never straddles images,
// n_tiles = B * ceil(H (W + 2) / 128): the AdaptiveAvgPool2d(1) of CALayer (src/drn.py:126, 137) without re-reading the output.
// ADSR_ERR_BAD_SHAPE = shape not covered: the caller uses the streaming implicit GEMM (launch_tc_gemm) with the ordinary packing.
int launch_conv_halo(ConvHaloParams& p, const void* in, long long ld_in, int num_sms, cudaStream_t stream) {
    if (p.B <= 0) return ADSR_OK;
    if (p.Cin <= 0 || p.Cin > 128 || p.BN < 16 || p.BN > 128 || (p.BN % 16) || p.N > p.BN || p.W < 8 || p.W > 126 || p.H < 1) return ADSR_ERR_BAD_SHAPE;
    if ((ld_in % 8) || (p.ldo % 8) || (p.ocol0 % 8) || (p.n_store % 8) || p.n_store > p.BN || (reinterpret_cast<uintptr_t>(in) & 15) ||
        (reinterpret_cast<uintptr_t>(p.out) & 15) || (reinterpret_cast<uintptr_t>(p.wp) & 15))
        return ADSR_ERR_BAD_ALIGN;
    p.k16_per_tap = (p.Cin + 15) / 16;
    // channel panels: 64-wide (128-byte swizzle); a remainder of <= 16 channels behind ONE full panel becomes a 16-wide tail panel
    // (32-byte swizzle, a quarter of the shared memory and L2 traffic) -- DRN-L's 80 channels
    const int rem = p.Cin % 64;
    p.full_panels = p.Cin / 64;
    p.tail = 0;
    if (rem != 0) {
        if (p.full_panels == 1 && rem <= 16) p.tail = 1;
        else p.full_panels += 1;
    }
    p.slabs = (9 * p.k16_per_tap + 3) / 4;
    p.Wh = p.W + 2;
    // rows of the halo box: the tile's 128 positions start anywhere inside a padded row, plus one padded row above and below
    p.box_rows = (128 + 4 * p.Wh) / p.Wh;
    p.box_bytes = p.box_rows * p.Wh * 128;
    p.panel_bytes = (p.box_bytes + 1023) / 1024 * 1024;
    p.tail_box_bytes = p.box_rows * p.Wh * 32;
    p.tail_bytes = p.tail ? (p.tail_box_bytes + 255) / 256 * 256 : 0;
    p.tiles_per_img = (p.H * p.Wh + 127) / 128;
    p.n_tiles = p.B * p.tiles_per_img;
    const int fixed = p.slabs * p.BN * 128 + kGuardBytes + 128 * 4 + static_cast<int>(sizeof(HaloBarriers));
    const int per_buf = p.full_panels * p.panel_bytes + p.tail_bytes;
    p.n_abuf = fixed + 2 * per_buf <= kSmemLimit ? 2 : 1;               // the next tile's halo loads while this one's MMAs run
    const int smem_bytes = fixed + p.n_abuf * per_buf;
    if (smem_bytes > kSmemLimit || p.box_rows > 256) return ADSR_ERR_BAD_SHAPE;
    const int k8 = (p.Cin + 7) & ~7;
    if (encode_tmap_nhwc_box3_bf16(&p.tmap_in, in, p.B, p.H, p.W, k8, ld_in, 64, p.Wh, p.box_rows) != ADSR_OK) return ADSR_ERR_CUDA;
    if (p.tail && encode_tmap_nhwc_box3_bf16(&p.tmap_tail, in, p.B, p.H, p.W, k8, ld_in, 16, p.Wh, p.box_rows) != ADSR_OK) return ADSR_ERR_CUDA;
    const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
    if (ensure_dynamic_smem(conv_halo_kernel, smem_bytes) != cudaSuccess) return ADSR_ERR_CUDA;
    // every global access that depends on (or could disturb) the previous kernel sits behind the loader's pdl_wait(): the A tiles,
    // hence the MMAs, hence the epilogue's stores
    return launch_pdl(conv_halo_kernel, dim3(grid), dim3(kThreads), static_cast<size_t>(smem_bytes), stream, p) == cudaSuccess ? ADSR_OK : ADSR_ERR_LAUNCH;
}

#if HALO_TRACE
extern "C" int adsr_debug_halo_trace(long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, g_halo_trace, sizeof(g_halo_trace)) == cudaSuccess ? 0 : 4;
}
#endif
}  // namespace adsr
