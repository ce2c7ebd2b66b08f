// tcgen05 / TMEM GEMM and implicit-GEMM 3x3 convolution for sm_100a -- variant with the MANUAL epilogue
// (ordinary coalesced stores; used for outputs without a residual, PixelShuffle stores and slab slices at 8-byte
// aligned columns).  The mainloop is identical to tc_gemm.cu, which carries the TMA epilogue; see there for the design.
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
constexpr int kNumThreadsConv = 512;          // 16 warps: loader, MMA, TMEM alloc, spare, 8 epilogue, 4 A producers
constexpr int kNumThreadsGemm = 384;          // GEMM mode (A by TMA): no producer warps -> up to 168 registers per thread
enum Mode { MODE_GEMM = 0, MODE_GEMM_LN = 1, MODE_GEMM_STATS = 2, MODE_CONV = 3, MODE_CONV_TMA = 4 };
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kTmemCols = 512;                // 2 accumulator buffers x 256 fp32 columns
constexpr int kStageOutBytes = 32 * 64;       // per epilogue warp: 32 rows x 32 bf16 staging tile
constexpr int kRingBytes = kStages * (kAStageBytes + kBStageBytes);
constexpr int kSmemBytes = kRingBytes + kEpiWarps * kStageOutBytes + 2 * 256 * 4 /*bias*/ + 2 * 256 * 4 /*colsum*/ + 256;

struct __align__(8) RingBarriers {
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};

// exact-erf GELU to 8e-7 absolute: erf(z) = 1 - 2^-p(z) with a degree-5 fit of p (z = |x|/sqrt 2 folded
// into the coefficients), so gelu(x) = 0.5 * (x + |x| - |x| * 2^-q(|x|)): 5 FMA + 1 MUFU + 3.
__device__ __forceinline__ float gelu_erf(float x) {
    const float a = fabsf(x);
    float q = 4.88103149e-4f;                    // c5 / 2^2.5
    q = fmaf(q, a, -7.19872210e-3f);             // c4 / 4
    q = fmaf(q, a, 5.21466284e-2f);              // c3 / 2^1.5
    q = fmaf(q, a, 4.59595859e-1f);              // c2 / 2
    q = fmaf(q, a, 1.15100050e+0f);              // c1 / sqrt 2
    const float e = exp2f(-q * a);
    return 0.5f * (x + fmaf(-a, e, a));
}

template <int ACT>
__device__ __forceinline__ float apply_act(float v, float slope) {
    if constexpr (ACT == ADSR_ACT_LRELU) return v > 0.f ? v : v * slope;
    if constexpr (ACT == ADSR_ACT_GELU) return gelu_erf(v);
    if constexpr (ACT == ADSR_ACT_RELU) return fmaxf(v, 0.f);
    return v;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ uint32_t add_bf16x2_f32(uint32_t a, uint32_t b, bool lo_ok, bool hi_ok) {
    const float l = bf16_lo(a) + (lo_ok ? bf16_lo(b) : 0.f);
    const float h = bf16_hi(a) + (hi_ok ? bf16_hi(b) : 0.f);
    return pack_bf16x2(l, h);
}

template <int ACT, int MODE>
__global__ void __launch_bounds__(MODE == MODE_CONV ? kNumThreadsConv : kNumThreadsGemm, 1)
tc_gemm_manual_kernel(const __grid_constant__ TcGemmParams p) {
    constexpr bool STATS = MODE == MODE_GEMM_STATS;
    constexpr bool LNF = MODE == MODE_GEMM_LN;
    constexpr bool CONV = MODE == MODE_CONV;            // A stages gathered by producer warps
    constexpr bool CONVT = MODE == MODE_CONV_TMA;       // A stages = 4-D TMA boxes of the NHWC image, shifted by the tap
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * kAStageBytes;
    uint8_t* smem_out = smem + kRingBytes;                                   // [8 warps][32 rows][64 B]
    float* smem_bias = reinterpret_cast<float*>(smem_out + kEpiWarps * kStageOutBytes);   // [2][256]
    float* smem_colsum = smem_bias + 2 * 256;                                             // [2][256]
    RingBarriers* bars = reinterpret_cast<RingBarriers*>(smem_colsum + 2 * 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.m_tiles * p.n_tiles;

    if ((smem_u32(smem) & 1023u) != 0) __trap();   // swizzle-128B operands need 1024 B alignment

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bars->full[s], CONV ? 128 + 1 : 1);   // (128 producer threads +) the copy issuer
            mbar_init(&bars->empty[s], 1);         // one tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->tmem_full[b], 1);
            mbar_init(&bars->tmem_empty[b], kEpiThreads);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<kTmemCols>(&bars->tmem_base);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0 || (CONVT && warp == 3)) {
        // ================================ loader(s): one bulk copy (+ one TMA box) per ring stage.  The TMA-fed conv walks 18 - 27
        // stages per tile and is bound by the per-stage overhead of this loop: two warps share it (even / odd ring stages)
        {   // converged warp, one elected lane issues (keeps descriptors in uniform registers)
            const uint32_t b_bytes = static_cast<uint32_t>(p.BN) * 128u;
            const uint32_t tx_bytes = b_bytes + (!CONV ? static_cast<uint32_t>(kAStageBytes) : 0u);
            if (!CONV && lane == 0) tma_prefetch_desc(&p.tmap_a);
            const int hw = p.Hin * p.Win;
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n_tile = tile % p.n_tiles;
                const int m0 = (tile / p.n_tiles) * 128;
                const uint8_t* src = p.Bp + static_cast<size_t>(n_tile) * p.num_k_stages * b_bytes;
                // conv, TMA-fed: tile = 128 / W whole image rows starting at (b, y0).  The stage loop is bound by its own
                // per-pass overhead (18 - 27 stages per tile), so nothing in it divides: tap and panel are counters
                const int cb = CONVT ? m0 / hw : 0, cy0 = CONVT ? (m0 - cb * hw) / p.Win : 0;
                int tap_y = -1, tap_x = -1, cpan = 0;                  // tap (dy, dx) and channel panel of stage ks
                for (int ks = 0; ks < p.num_k_stages; ++ks) {
                    const bool mine = !CONVT || ((stage & 1) == (warp == 0 ? 0 : 1));   // kStages is even: parity of the ring stage
                    if (mine) mbar_wait(&bars->empty[stage], phase ^ 1);
                    if (mine && elect_one_sync()) {
                        mbar_arrive_expect_tx(&bars->full[stage], tx_bytes);
                        if (CONVT) {
                            // tap (dy, dx) reads the tile's box shifted by (dx, dy): columns / rows outside the image (and channels
                            // >= Cin) are zero-filled by the TMA unit
                            tma_load_4d(smem_a + stage * kAStageBytes, &p.tmap_a, cpan * 64, tap_x, cy0 + tap_y, cb, &bars->full[stage]);
                        } else if (!CONV) {
                            tma_load_2d(smem_a + stage * kAStageBytes, &p.tmap_a, ks * 64, m0, &bars->full[stage]);
                        }
                        bulk_g2s(smem_b + stage * kBStageBytes, src + static_cast<size_t>(ks) * b_bytes, b_bytes,
                                 &bars->full[stage]);
                    }
                    __syncwarp();
                    if (CONVT && ++cpan == p.stages_per_tap) {
                        cpan = 0;
                        if (++tap_x == 2) { tap_x = -1; ++tap_y; }
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (single thread) ================================
        {   // converged warp, one elected lane issues
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
                int cpan = 0;                                          // channel panel of stage ks inside its tap (conv)
                for (int ks = 0; ks < p.num_k_stages; ++ks) {
                    int steps;
                    if (CONV || CONVT) {
                        steps = p.k16_per_tap - 4 * cpan;
                        if (++cpan == p.stages_per_tap) cpan = 0;
                    } else {
                        steps = p.k16_total - 4 * ks;
                    }
                    steps = steps > 4 ? 4 : steps;
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after_sync();
                    const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem_a + stage * kAStageBytes));
                    const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem_b + stage * kBStageBytes));
                    if (elect_one_sync()) {
                        // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in 16-byte units
                        umma_bf16(d_tmem, adesc, bdesc, idesc, ks != 0 ? 1u : 0u);
                        if (steps > 1) umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
                        if (steps > 2) umma_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
                        if (steps > 3) umma_bf16(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
                        umma_commit(&bars->empty[stage]);      // frees the ring slot when the MMAs retire
                        if (ks == p.num_k_stages - 1) umma_commit(&bars->tmem_full[buf]);   // accumulator complete -> epilogue
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4 && warp < 4 + kEpiWarps) {
        // ================================ epilogue: TMEM -> regs -> smem staging -> coalesced global ==
        // 8 warps: quadrant q = warp & 3 owns TMEM lanes / tile rows 32q..32q+31, the two warps of a
        // quadrant take alternate 32-column chunks.  Values are staged as bf16 in a per-warp 32x32
        // swizzled tile so that global stores (and residual loads) are row-contiguous 64 B segments.
        const int ew = warp - 4;
        const int quad = warp & 3;
        const int half = ew >> 2;
        const int et = threadIdx.x - 4 * 32;                   // 0..255 inside the epilogue group
        uint8_t* stg = smem_out + ew * kStageOutBytes;
        const uint32_t stg_w = smem_u32(stg) + static_cast<uint32_t>(lane * 64);     // my row when writing
        const int wsw = (lane >> 1) & 3;
        const int n_chunks = (p.BN + 31) >> 5;
        const bool st16 = (p.ocol0 & 7) == 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int m_tile = tile / p.n_tiles;
            const int n_tile = tile % p.n_tiles;
            const int buf = it & 1;
            const uint32_t use = static_cast<uint32_t>(it >> 1);
            const int n_base = n_tile * p.BN;
            const int n_lim = min(p.n_store, n_base + p.BN);   // never write into the next N tile's columns
            float* bias_s = smem_bias + buf * 256;
            float* colsum_s = smem_colsum + buf * 256;
            if (et < p.BN) {
                bias_s[et] = __ldg(p.bias + n_base + et);
                if (LNF) colsum_s[et] = __ldg(p.colsum + n_base + et);
            }
            // LayerNorm statistics of the row this thread owns: partial (sum, sumsq) slots of the producing kernel(s)
            float ln_mean = 0.f, ln_rstd = 1.f;
            const int own_row = m_tile * 128 + quad * 32 + lane;
            if (LNF && own_row < p.M) {
                const float2* sp = p.stats_in + static_cast<long long>(own_row) * p.stats_in_stride;
                float s1 = 0.f, s2 = 0.f;
                for (int k = 0; k < p.stats_in_slots; ++k) {
                    const float2 sv = __ldg(sp + k);
                    s1 += sv.x;
                    s2 += sv.y;
                }
                const float inv_c = 1.0f / static_cast<float>(p.ln_C);
                ln_mean = s1 * inv_c;
                ln_rstd = rsqrtf(fmaxf(s2 * inv_c - ln_mean * ln_mean, 0.f) + p.ln_eps);
            }
            float st_sum = 0.f, st_sq = 0.f;
            named_bar_sync(1, kEpiThreads);
            mbar_wait(&bars->tmem_full[buf], use & 1);
            tc_fence_after_sync();
            const int row0 = m_tile * 128 + quad * 32;
            const uint32_t taddr = tmem_base + static_cast<uint32_t>(buf * 256) + (static_cast<uint32_t>(quad * 32) << 16);

            for (int ch = half; ch < n_chunks; ch += 2) {
                const int c0 = ch * 32;
                const int n0 = n_base + c0;
                const bool wide = c0 + 32 <= p.BN;             // BN % 32 == 16: last chunk is 16 columns
                const bool do_store = n0 < n_lim;              // warp-uniform
                // ---- residual prefetch in the coalesced (write-out) mapping: 4 rows-of-8 x 4 chunks
                uint4 rres[4];
                if (p.res != nullptr && do_store) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = row0 + i * 8 + (lane >> 2);
                        const int col = n0 + (lane & 3) * 8;
                        rres[i] = (r < p.M && col < n_lim)
                                      ? __ldg(reinterpret_cast<const uint4*>(p.res + static_cast<long long>(r) * p.ldres + col))
                                      : make_uint4(0, 0, 0, 0);
                    }
                }
                uint32_t raw[32];
                __syncwarp();                                   // tcgen05.ld is .sync.aligned
                if (wide) {
                    tmem_ld32(taddr + static_cast<uint32_t>(c0), raw);
                } else {
                    uint32_t lo[16];
                    tmem_ld16(taddr + static_cast<uint32_t>(c0), lo);
#pragma unroll
                    for (int j = 0; j < 16; ++j) { raw[j] = lo[j]; raw[16 + j] = 0; }
                }
                tmem_ld_wait();
                if (!do_store) continue;                        // warp-uniform
                // ---- bias + activation + alpha, pack to bf16, stage
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 bb = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * j);
                    float a0 = __uint_as_float(raw[4 * j + 0]), a1 = __uint_as_float(raw[4 * j + 1]);
                    float a2 = __uint_as_float(raw[4 * j + 2]), a3 = __uint_as_float(raw[4 * j + 3]);
                    if (LNF) {                                  // LN(x) W^T = rstd * (x W'^T - mean * colsum)
                        const float4 cs = *reinterpret_cast<const float4*>(colsum_s + c0 + 4 * j);
                        a0 = ln_rstd * fmaf(-ln_mean, cs.x, a0); a1 = ln_rstd * fmaf(-ln_mean, cs.y, a1);
                        a2 = ln_rstd * fmaf(-ln_mean, cs.z, a2); a3 = ln_rstd * fmaf(-ln_mean, cs.w, a3);
                    }
                    const float v0 = apply_act<ACT>(a0 + bb.x, p.slope) * p.alpha;
                    const float v1 = apply_act<ACT>(a1 + bb.y, p.slope) * p.alpha;
                    const float v2 = apply_act<ACT>(a2 + bb.z, p.slope) * p.alpha;
                    const float v3 = apply_act<ACT>(a3 + bb.w, p.slope) * p.alpha;
                    if (STATS) {                                // row statistics of the stored values (no residual here)
                        st_sum += (v0 + v1) + (v2 + v3);
                        st_sq = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, st_sq))));
                    }
                    if (p.out_mode == ADSR_OUT_ROWS) {
                        pk[2 * j] = pack_bf16x2(v0, v1);
                        pk[2 * j + 1] = pack_bf16x2(v2, v3);
                    } else {
                        // PixelShuffle(2): column 4c+sub -> staging position sub*8 + c  (8 channels per chunk)
                        // handled below from the float values: keep them in raw[]
                        raw[4 * j + 0] = __float_as_uint(v0); raw[4 * j + 1] = __float_as_uint(v1);
                        raw[4 * j + 2] = __float_as_uint(v2); raw[4 * j + 3] = __float_as_uint(v3);
                    }
                }
                if (p.out_mode != ADSR_OUT_ROWS) {
#pragma unroll
                    for (int sub = 0; sub < 4; ++sub)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            pk[sub * 4 + c] = pack_bf16x2(__uint_as_float(raw[(2 * c) * 4 + sub]),
                                                          __uint_as_float(raw[(2 * c + 1) * 4 + sub]));
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(stg_w + static_cast<uint32_t>((j ^ wsw) << 4)),
                                 "r"(pk[4 * j]), "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                                 : "memory");
                }
                __syncwarp();
                // ---- coalesced write-out: lane -> (row = i*8 + lane/4, 16-byte chunk = lane%4)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rl = i * 8 + (lane >> 2);
                    const int cc = lane & 3;
                    const uint4 val = *reinterpret_cast<const uint4*>(stg + rl * 64 + ((cc ^ ((rl >> 1) & 3)) << 4));
                    const int r = row0 + rl;
                    if (r >= p.M) continue;
                    if (p.out_mode == ADSR_OUT_ROWS) {
                        const int col = n0 + cc * 8;
                        if (col >= n_lim) continue;
                        uint4 o = val;
                        if (p.res != nullptr) {
                            o.x = add_bf16x2_f32(val.x, rres[i].x, col + 0 < p.N, col + 1 < p.N);
                            o.y = add_bf16x2_f32(val.y, rres[i].y, col + 2 < p.N, col + 3 < p.N);
                            o.z = add_bf16x2_f32(val.z, rres[i].z, col + 4 < p.N, col + 5 < p.N);
                            o.w = add_bf16x2_f32(val.w, rres[i].w, col + 6 < p.N, col + 7 < p.N);
                        }
                        __nv_bfloat16* dst = p.out + static_cast<long long>(r) * p.ldo + p.ocol0 + col;
                        if (col + 8 > n_lim) {                  // n_store % 8 == 4: only the first 4 columns belong to us
                            reinterpret_cast<uint2*>(dst)[0] = make_uint2(o.x, o.y);
                        } else if (st16) {
                            *reinterpret_cast<uint4*>(dst) = o;
                        } else {                                // slab slices start at 8-byte aligned columns
                            reinterpret_cast<uint2*>(dst)[0] = make_uint2(o.x, o.y);
                            reinterpret_cast<uint2*>(dst)[1] = make_uint2(o.z, o.w);
                        }
                    } else {
                        // staging chunk cc = sub-pixel (i2, j2); 8 channels n0/4 .. n0/4+7
                        const int hw = p.Hout * p.Wout;
                        const int b = r / hw;
                        const int rem = r - b * hw;
                        const int y = rem / p.Wout;
                        const int x = rem - y * p.Wout;
                        const int i2 = cc >> 1, j2 = cc & 1;
                        __nv_bfloat16* dst = p.out +
                            ((static_cast<long long>(b) * (2 * p.Hout) + 2 * y + i2) * (2 * p.Wout) + 2 * x + j2) * p.ldo + (n0 >> 2);
                        *reinterpret_cast<uint4*>(dst) = val;
                    }
                }
                __syncwarp();                                   // staging tile is reused by the next chunk
            }
            if (STATS && own_row < p.M)
                p.stats_out[static_cast<long long>(own_row) * p.stats_out_stride + p.stats_out_slot0 + n_tile * 2 + half] =
                    make_float2(st_sum, st_sq);
            __syncwarp();
            tc_fence_before_sync();
            mbar_arrive(&bars->tmem_empty[buf]);
        }
    } else if (CONV && warp >= 4 + kEpiWarps) {
        // ================================ A producers (4 warps, 128 threads) =========================
        const int pw = warp - (4 + kEpiWarps);
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

int launch_tc_gemm_manual(const TcGemmParams& p, int grid, cudaStream_t stream) {
    auto launch = [&](auto kernel, int threads) -> int {
        if (ensure_dynamic_smem(kernel, kSmemBytes) != cudaSuccess) return ADSR_ERR_CUDA;
        kernel<<<grid, threads, kSmemBytes, stream>>>(p);
        return cudaGetLastError() == cudaSuccess ? ADSR_OK : ADSR_ERR_LAUNCH;
    };
    const bool conv = !p.use_tma && !p.conv_tma;
    if ((conv || p.conv_tma) && (p.ln_fold || p.stats_out != nullptr)) return ADSR_ERR_BAD_SHAPE;
    if (p.ln_fold && p.stats_out != nullptr) return ADSR_ERR_BAD_SHAPE;
#define ADSR_CASE(ACT_)                                                                                        \
    case ACT_:                                                                                                 \
        if (p.conv_tma) return launch(tc_gemm_manual_kernel<ACT_, MODE_CONV_TMA>, kNumThreadsGemm);            \
        if (conv) return launch(tc_gemm_manual_kernel<ACT_, MODE_CONV>, kNumThreadsConv);                      \
        if (p.ln_fold) return launch(tc_gemm_manual_kernel<ACT_, MODE_GEMM_LN>, kNumThreadsGemm);              \
        if (p.stats_out != nullptr) return launch(tc_gemm_manual_kernel<ACT_, MODE_GEMM_STATS>, kNumThreadsGemm); \
        return launch(tc_gemm_manual_kernel<ACT_, MODE_GEMM>, kNumThreadsGemm);
    switch (p.act) {
        ADSR_CASE(ADSR_ACT_NONE)
        ADSR_CASE(ADSR_ACT_LRELU)
        ADSR_CASE(ADSR_ACT_GELU)
        ADSR_CASE(ADSR_ACT_RELU)
    }
#undef ADSR_CASE
    return ADSR_ERR_BAD_SHAPE;
}

}  // namespace adsr
