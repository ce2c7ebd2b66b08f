// Fused Swin MLP half for sm_100a:   z = y + fc2( GELU_erf( fc1( LayerNorm(y) ) ) )      (src/drct.py:510, 173-190)
//
// One persistent, warp-specialised kernel (grid = #SMs, 20 warps).  A CTA owns 128 token rows at a time; the hidden
// activations [128 x H] never leave the SM:
//   * the y tile [128 x C] arrives by TMA (128-byte swizzle, K / M tails zero-filled) into one of TWO shared-memory
//     buffers (the next tile is prefetched a whole tile ahead); it is the A operand of fc1, the residual of the last
//     epilogue, and -- overwritten in place with z -- the source of the TMA store of the result;
//   * the weights of fc1 (gamma-folded) and fc2 stream from L2 through a ring of shared-memory slots, one
//     [rows x 64] slab per slot, in exactly the order a host-built static schedule consumes them;
//   * fc1 is computed in hidden chunks of <= 128 columns into a double-buffered TMEM accumulator (tcgen05.mma, SS);
//     16 epilogue warps turn a chunk into bf16 GELU activations (LayerNorm folded: rstd * (acc - mean * colsum) + b)
//     and write them back IN PLACE into the same TMEM columns (tcgen05.st), from where fc2 consumes them as the
//     TMEM A operand (tcgen05.mma, TS) while fc1 of the next chunk is already running;
//   * fc2 accumulates over the chunks into a third TMEM region; the last epilogue adds bias and the residual.
// The 0.5 of GELU is folded into the packed fc2 weights.
#include "adsr_kernels.h"
#include "ptx.cuh"

namespace adsr {

namespace {

constexpr int kThreads = 640;                 // warp 0 loader, 1 MMA issuer, 2 TMEM alloc + TMA store, 3 spare, 4..19 epilogue
constexpr int kEpiWarps = 16;
constexpr int kPanelBytes = 128 * 128;        // 128 rows x 64 bf16
constexpr int kMaxHidden = 512;               // padded hidden columns (sum of chunk strides)
constexpr int kMaxN2 = 320;
constexpr int kConstBytes = (2 * kMaxHidden + kMaxN2) * 4;
constexpr int kSmemLimit = 232448;            // 227 KB

struct __align__(8) MlpBarriers {
    uint64_t w_full[8];
    uint64_t w_empty[8];
    uint64_t a_full[2];
    uint64_t a_empty[2];
    uint64_t z_ready[2];
    uint64_t acc1_full[2];
    uint64_t h_ready[2];
    uint64_t acc2_full;
    uint64_t acc2_free;
    uint32_t tmem_base;
};

// GELU(x) * 2 = x + |x| - |x| * erfc(|x| / sqrt 2), erfc(a / sqrt 2) ~ 2^(-a q(a)), q quadratic: |error| <= 8.6e-5 after
// the 0.5 that lives in the fc2 weights (the activations are rounded to bf16 right after: 2^-9 relative).
__device__ __forceinline__ float gelu2(float x) {
    const float a = fabsf(x);
    float q = fmaf(0.027645503f, a, 0.48822206f);
    q = fmaf(q, a, 1.1409364f);
    const float e = ex2_approx(-q * a);
    return x + fmaf(-a, e, a);
}

__global__ void __launch_bounds__(kThreads, 1) swin_mlp_kernel(const __grid_constant__ SwinMlpParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_buf = smem;                                            // 2 x a_buf_bytes
    uint8_t* ring = smem + 2 * p.a_buf_bytes;                         // n_slots x slot_bytes
    float* s_bias1 = reinterpret_cast<float*>(ring + p.n_slots * p.slot_bytes);
    float* s_colsum1 = s_bias1 + kMaxHidden;
    float* s_bias2 = s_colsum1 + kMaxHidden;
    MlpBarriers* bars = reinterpret_cast<MlpBarriers*>(s_bias2 + kMaxN2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int my_tiles = static_cast<int>(blockIdx.x) < p.m_tiles ? (p.m_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if ((smem_u32(smem) & 1023u) != 0) __trap();

    for (int i = threadIdx.x; i < p.nc * p.hc; i += kThreads) {
        s_bias1[i] = p.bias1[i];
        s_colsum1[i] = p.colsum1[i];
    }
    for (int i = threadIdx.x; i < p.n2; i += kThreads) s_bias2[i] = p.bias2[i];

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.n_slots; ++s) {
            mbar_init(&bars->w_full[s], 1);
            mbar_init(&bars->w_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->a_full[b], 1);
            mbar_init(&bars->a_empty[b], 1);
            mbar_init(&bars->z_ready[b], kEpiWarps);
            mbar_init(&bars->acc1_full[b], 1);
            mbar_init(&bars->h_ready[b], kEpiWarps);
        }
        mbar_init(&bars->acc2_full, 1);
        mbar_init(&bars->acc2_free, kEpiWarps);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(&bars->tmem_base);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = bars->tmem_base;

    if (warp == 0) {
        // ============================================================ loader (one thread)
        if (lane == 0 && my_tiles > 0) {
            tma_prefetch_desc(&p.tmap_y);
            auto load_a = [&](int it) {
                const int ab = it & 1;
                const int m0 = (it * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x)) * 128;
                mbar_arrive_expect_tx(&bars->a_full[ab], static_cast<uint32_t>(p.a_buf_bytes));
                for (int pn = 0; pn < p.ks1; ++pn)
                    tma_load_2d(a_buf + ab * p.a_buf_bytes + pn * kPanelBytes, &p.tmap_y, pn * 64, m0, &bars->a_full[ab]);
            };
            load_a(0);
            int slot = 0;
            uint32_t phase = 0;
            const int t_prefetch = p.n_stages >> 1;
            for (int it = 0; it < my_tiles; ++it) {
                for (int t = 0; t < p.n_stages; ++t) {
                    if (t == t_prefetch && it + 1 < my_tiles) {
                        // the store of tile it-1 (same buffer) must have finished reading it
                        mbar_wait(&bars->a_empty[(it + 1) & 1], (static_cast<uint32_t>((it + 1) >> 1) & 1) ^ 1);
                        load_a(it + 1);
                    }
                    const uint32_t bytes = p.stages[t].bytes;
                    mbar_wait(&bars->w_empty[slot], phase ^ 1);
                    mbar_arrive_expect_tx(&bars->w_full[slot], bytes);
                    bulk_g2s(ring + slot * p.slot_bytes, p.wp + p.stages[t].goff, bytes, &bars->w_full[slot]);
                    if (++slot == p.n_slots) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer (one thread)
        if (lane == 0) {
            int slot = 0;
            uint32_t phase = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int ab = it & 1;
                mbar_wait(&bars->a_full[ab], static_cast<uint32_t>(it >> 1) & 1);
                tc_fence_after_sync();
                const uint32_t a_base = smem_u32(a_buf + ab * p.a_buf_bytes);
                for (int t = 0; t < p.n_stages; ++t) {
                    const MlpStage e = p.stages[t];
                    const int cg = it * p.nc + e.chunk;           // running chunk counter: buffer = cg & 1, use = cg >> 1
                    const int b = cg & 1;
                    if (e.flags & MLP_STAGE_WAIT_H) {
                        mbar_wait(&bars->h_ready[b], static_cast<uint32_t>(cg >> 1) & 1);
                        if (e.chunk == 0) mbar_wait(&bars->acc2_free, (static_cast<uint32_t>(it) & 1) ^ 1);
                        tc_fence_after_sync();
                    }
                    mbar_wait(&bars->w_full[slot], phase);
                    tc_fence_after_sync();
                    const uint64_t bdesc = umma_desc_k_sw128(smem_u32(ring + slot * p.slot_bytes));
                    const uint32_t idesc = umma_idesc_bf16_m128(e.rows);
                    const uint32_t acc1 = tmem + static_cast<uint32_t>(p.acc1_col[b]);
                    if (e.kind == 0) {
                        const uint64_t adesc = umma_desc_k_sw128(a_base + e.kidx * kPanelBytes);
                        for (int k = 0; k < e.ksteps; ++k)
                            umma_bf16(acc1, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                                      ((e.flags & MLP_STAGE_FIRST) && k == 0) ? 0u : 1u);
                    } else {
                        const uint32_t d = tmem + e.dcol;
                        for (int k = 0; k < e.ksteps; ++k)
                            umma_bf16_ts(d, acc1 + static_cast<uint32_t>(8 * (4 * e.kidx + k)), bdesc + static_cast<uint64_t>(2 * k),
                                         idesc, ((e.flags & MLP_STAGE_FIRST) && k == 0) ? 0u : 1u);
                    }
                    umma_commit(&bars->w_empty[slot]);
                    if (e.flags & MLP_STAGE_ACC1_DONE) umma_commit(&bars->acc1_full[b]);
                    if (e.flags & MLP_STAGE_ACC2_DONE) umma_commit(&bars->acc2_full);
                    if (++slot == p.n_slots) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 2) {
        // ============================================================ TMA store of finished tiles (one thread)
        if (lane == 0) {
            for (int it = 0; it < my_tiles; ++it) {
                const int ab = it & 1;
                const int m0 = (it * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x)) * 128;
                mbar_wait(&bars->z_ready[ab], static_cast<uint32_t>(it >> 1) & 1);
                for (int pn = 0; pn < p.ks1; ++pn)
                    tma_store_2d_box(&p.tmap_z, a_buf + ab * p.a_buf_bytes + pn * kPanelBytes, pn * 64, m0);
                bulk_commit_group();
                bulk_wait_group_read0();
                mbar_arrive(&bars->a_empty[ab]);
            }
            bulk_wait_group0();
        }
    } else if (warp >= 4) {
        // ============================================================ epilogue: 4 quadrants (TMEM lanes) x 4 column groups
        const int quad = warp & 3;
        const int grp = (warp - 4) >> 2;
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const int r_in_tile = quad * 32 + lane;
        const uint32_t row_off = static_cast<uint32_t>(r_in_tile * 128);
        const int rsw = r_in_tile & 7;

        for (int it = 0; it < my_tiles; ++it) {
            const int ab = it & 1;
            const int m0 = (it * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x)) * 128;
            const int row = m0 + r_in_tile;
            float rstd = 1.f, nrm = 0.f;                              // nrm = -mean * rstd
            if (row < p.M) {
                const float2* sp = p.stats_in + static_cast<long long>(row) * p.stats_in_stride;
                float s1 = 0.f, s2 = 0.f;
                for (int k = 0; k < p.stats_in_slots; ++k) {
                    const float2 v = __ldg(sp + k);
                    s1 += v.x;
                    s2 += v.y;
                }
                const float inv_c = 1.0f / static_cast<float>(p.C);
                const float mean = s1 * inv_c;
                rstd = rsqrtf(fmaxf(s2 * inv_c - mean * mean, 0.f) + p.ln_eps);
                nrm = -mean * rstd;
            }

            // ---------------- fc1 chunks: TMEM fp32 -> LN fold + bias + GELU -> bf16 back into the same TMEM columns
            for (int j = 0; j < p.nc; ++j) {
                const int cg = it * p.nc + j;
                const int b = cg & 1;
                const int octets = p.hcw[j] >> 3;
                const int c0 = ((grp * octets) >> 2) << 3;
                const int c1 = (((grp + 1) * octets) >> 2) << 3;
                const uint32_t taddr = tmem + static_cast<uint32_t>(p.acc1_col[b]) + lane_off;
                mbar_wait(&bars->acc1_full[b], static_cast<uint32_t>(cg >> 1) & 1);
                tc_fence_after_sync();
                uint32_t raw[32];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (c0 + 8 * u < c1) tmem_ld8(taddr + static_cast<uint32_t>(c0 + 8 * u), &raw[8 * u]);
                tmem_ld_wait();
                uint32_t pk[16];
                const float* bp = s_bias1 + j * p.hc + c0;
                const float* cp = s_colsum1 + j * p.hc + c0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (c0 + 8 * u < c1) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const float4 bb = *reinterpret_cast<const float4*>(bp + 8 * u + 4 * h);
                            const float4 cs = *reinterpret_cast<const float4*>(cp + 8 * u + 4 * h);
                            const float v0 = gelu2(fmaf(rstd, __uint_as_float(raw[8 * u + 4 * h + 0]), fmaf(nrm, cs.x, bb.x)));
                            const float v1 = gelu2(fmaf(rstd, __uint_as_float(raw[8 * u + 4 * h + 1]), fmaf(nrm, cs.y, bb.y)));
                            const float v2 = gelu2(fmaf(rstd, __uint_as_float(raw[8 * u + 4 * h + 2]), fmaf(nrm, cs.z, bb.z)));
                            const float v3 = gelu2(fmaf(rstd, __uint_as_float(raw[8 * u + 4 * h + 3]), fmaf(nrm, cs.w, bb.w)));
                            pk[4 * u + 2 * h] = pack_bf16x2(v0, v1);
                            pk[4 * u + 2 * h + 1] = pack_bf16x2(v2, v3);
                        }
                    }
                }
                // the packed activations of my columns land in columns some OTHER warp of this quadrant may still be reading
                tc_fence_before_sync();
                named_bar_sync(1 + quad, 128);
                tc_fence_after_sync();
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (c0 + 8 * u < c1) tmem_st4(taddr + static_cast<uint32_t>((c0 + 8 * u) >> 1), &pk[4 * u]);
                tmem_st_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->h_ready[b]);
            }

            // ---------------- fc2: + bias + residual (the y tile in shared memory), written back in place as z
            {
                const int octets = p.n2 >> 3;
                const int c0 = ((grp * octets) >> 2) << 3;
                const int c1 = (((grp + 1) * octets) >> 2) << 3;
                const uint32_t a_row = smem_u32(a_buf + ab * p.a_buf_bytes) + row_off;
                mbar_wait(&bars->a_full[ab], static_cast<uint32_t>(it >> 1) & 1);    // TMA-written tile visible to me
                mbar_wait(&bars->acc2_full, static_cast<uint32_t>(it) & 1);
                tc_fence_after_sync();
                const uint32_t taddr = tmem + lane_off;
                for (int c = c0; c < c1; c += 8) {
                    uint32_t raw[8];
                    tmem_ld8(taddr + static_cast<uint32_t>(c), raw);
                    const uint32_t saddr = a_row + static_cast<uint32_t>((c >> 6) * kPanelBytes) +
                                           static_cast<uint32_t>((((c >> 3) & 7) ^ rsw) << 4);
                    uint32_t rw[4];
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(rw[0]), "=r"(rw[1]), "=r"(rw[2]), "=r"(rw[3]) : "r"(saddr));
                    const float4 b0 = *reinterpret_cast<const float4*>(s_bias2 + c);
                    const float4 b1 = *reinterpret_cast<const float4*>(s_bias2 + c + 4);
                    tmem_ld_wait();
                    uint32_t o[4];
                    o[0] = pack_bf16x2(__uint_as_float(raw[0]) + b0.x + bf16_lo(rw[0]), __uint_as_float(raw[1]) + b0.y + bf16_hi(rw[0]));
                    o[1] = pack_bf16x2(__uint_as_float(raw[2]) + b0.z + bf16_lo(rw[1]), __uint_as_float(raw[3]) + b0.w + bf16_hi(rw[1]));
                    o[2] = pack_bf16x2(__uint_as_float(raw[4]) + b1.x + bf16_lo(rw[2]), __uint_as_float(raw[5]) + b1.y + bf16_hi(rw[2]));
                    o[3] = pack_bf16x2(__uint_as_float(raw[6]) + b1.z + bf16_lo(rw[3]), __uint_as_float(raw[7]) + b1.w + bf16_hi(rw[3]));
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
                }
                tc_fence_before_sync();
                fence_proxy_async_smem();                              // my st.shared -> visible to the TMA store
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&bars->acc2_free);
                    mbar_arrive(&bars->z_ready[ab]);
                }
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem);
    }
}

}  // namespace

int launch_swin_mlp(SwinMlpParams& p, const void* y, long long ldy, void* z, long long ldz, int num_sms, cudaStream_t stream) {
    if (p.M <= 0) return ADSR_OK;
    if (p.n_stages <= 0 || p.n_stages > kMlpMaxStages || p.nc <= 0 || p.nc > 8) return ADSR_ERR_BAD_SHAPE;
    if (p.n2 > kMaxN2 || (p.n2 % 16) != 0 || p.nc * p.hc > kMaxHidden || (p.hc % 8) != 0) return ADSR_ERR_BAD_SHAPE;
    if (p.n2 + 2 * p.hc > 512 || p.acc1_col[0] < p.n2 || p.acc1_col[1] < p.acc1_col[0] + p.hc || p.acc1_col[1] + p.hc > 512)
        return ADSR_ERR_BAD_SHAPE;
    if (p.ks1 <= 0 || p.ks1 > 5 || p.n_slots < 2 || p.n_slots > 8 || (p.slot_bytes % 1024) != 0) return ADSR_ERR_BAD_SHAPE;
    uint32_t goff = 0;
    for (int t = 0; t < p.n_stages; ++t) {
        MlpStage& e = p.stages[t];
        if (e.rows < 16 || e.rows > 256 || (e.rows % 16) != 0 || e.ksteps < 1 || e.ksteps > 4 || e.chunk >= p.nc) return ADSR_ERR_BAD_SHAPE;
        if (e.bytes != static_cast<uint32_t>(e.rows) * 128u || static_cast<int>(e.bytes) > p.slot_bytes) return ADSR_ERR_BAD_SHAPE;
        if (e.kind == 0 && (e.kidx >= p.ks1 || e.rows != p.hcw[e.chunk])) return ADSR_ERR_BAD_SHAPE;
        if (e.kind == 1 && (e.dcol + e.rows > p.n2 || 16 * (4 * e.kidx + e.ksteps) > p.hcw[e.chunk])) return ADSR_ERR_BAD_SHAPE;
        e.goff = goff;
        goff += e.bytes;
    }
    for (int j = 0; j < p.nc; ++j)
        if (p.hcw[j] <= 0 || p.hcw[j] > p.hc || p.hcw[j] > 128 || (p.hcw[j] % 16) != 0) return ADSR_ERR_BAD_SHAPE;
    p.a_buf_bytes = p.ks1 * kPanelBytes;
    const int smem_bytes = 2 * p.a_buf_bytes + p.n_slots * p.slot_bytes + kConstBytes + static_cast<int>(sizeof(MlpBarriers));
    if (smem_bytes > kSmemLimit) return ADSR_ERR_BAD_SHAPE;
    if ((reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(z) & 15) || (ldy % 8) || (ldz % 8) ||
        (reinterpret_cast<uintptr_t>(p.wp) & 15))
        return ADSR_ERR_BAD_ALIGN;
    int st = encode_tmap_rows_bf16(&p.tmap_y, y, p.M, p.C, ldy);
    if (st != ADSR_OK) return st;
    st = encode_tmap_rows_bf16(&p.tmap_z, z, p.M, p.C, ldz);
    if (st != ADSR_OK) return st;
    p.m_tiles = (p.M + 127) / 128;
    const int grid = p.m_tiles < num_sms ? p.m_tiles : num_sms;
    if (cudaFuncSetAttribute(swin_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) return ADSR_ERR_CUDA;
    swin_mlp_kernel<<<grid, kThreads, smem_bytes, stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? ADSR_OK : ADSR_ERR_LAUNCH;
}

}  // namespace adsr
