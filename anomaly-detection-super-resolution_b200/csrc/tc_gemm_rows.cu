// tcgen05 / TMEM GEMM for token-row outputs, second generation ("row-tile" kernel):
//     out[:, ocol0 + n] = alpha * act( LNfold( A[M, K] W^T ) + bias ) (+ res)      K <= 320
// Used for the qkv / proj / adjust5 linears of DRCT (src/drct.py:278, 300, 334-374 with norm1 folded, src/drct.py:481).
//
// Differences to tc_gemm.cu (which stays for long K, convolutions, PixelShuffle and 8-byte aligned slab slices):
//   * a CTA keeps the WHOLE K extent of its 128-row A tile in shared memory (one TMA load per tile, double buffered when
//     it fits) and walks over all N tiles of 128 columns, so A is fetched from L2 / HBM once instead of once per N tile;
//   * four 128-column accumulators in TMEM: the MMA warp runs up to four N tiles ahead of the epilogue;
//   * 16 epilogue warps (4 TMEM lane quadrants x 4 column groups) work on 64-column slices: tcgen05.ld -> fp32 math with
//     packed fp32x2 instructions -> bf16 into a [128 x 64] shared-memory staging panel (128-byte swizzle) that a
//     dedicated warp stores with ONE TMA store per slice; the residual slice is TMA-loaded into the same panel beforehand
//     and added in place.  No per-thread global loads or stores anywhere;
//   * control warps stay converged and issue through one elected lane (uniform-register descriptors).
#include "adsr_kernels.h"
#include "ptx.cuh"

namespace adsr {

namespace {

constexpr int kEpiWarps = 16;                 // warps 0..15
constexpr int kBLoaderWarp = 16;
constexpr int kMmaWarp = 17;                  // MMA issuer of the even N tiles (running count)
constexpr int kALoaderWarp = 18;              // + TMEM alloc
constexpr int kStoreWarp = 19;
constexpr int kMmaWarp2 = 20;                 // MMA issuer of the odd N tiles: the bookkeeping of one overlaps the MMAs of the other
constexpr int kBLoaderWarp2 = 21;             // each issuer has its own weight ring and loader (single producer / single consumer)
constexpr int kThreads = 22 * 32;
constexpr int kPanelBytes = 128 * 128;
constexpr int kBN = 128;
constexpr int kAcc = 4;
constexpr int kMaxSlots = 8;
constexpr int kStg = 2;                       // staging panels (64 output columns each); two TMA stores may be in flight
constexpr int kMaxN = 1024;
constexpr int kSmemLimit = 232448;

struct __align__(8) RowBarriers {
    uint64_t b_full[2][kMaxSlots], b_empty[2][kMaxSlots];
    uint64_t a_full[2], a_empty[2];
    uint64_t acc_full[kAcc], acc_free[kAcc];
    uint64_t stg_ready[kStg];                 // 16 epilogue warps have written the staging panel
    uint64_t stg_free[kStg];                   // the store has read it / the residual slice has landed in it
    uint32_t tmem_base;
};

struct RowParams {
    TcGemmParams g;
    int ks1, n_abuf, n_slots;
    int rev;                                  // walk the M tiles from the last one down
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

// exact-erf GELU (pair), 8.6e-5 absolute: see swin_mlp.cu
__device__ __forceinline__ float2 gelu_pair(float2 x) {
    const float2 a = f2(fabsf(x.x), fabsf(x.y));          // folds into |R| operand modifiers of FFMA2 / FMUL2
    float2 q = __ffma2_rn(f2(-0.027645503f, -0.027645503f), a, f2(-0.48822206f, -0.48822206f));
    q = __ffma2_rn(q, a, f2(-1.1409364f, -1.1409364f));
    const float2 m = __fmul2_rn(q, a);
    const float2 e = f2(ex2_approx(m.x), ex2_approx(m.y));
    const float2 g = __ffma2_rn(e, f2(-1.f, -1.f), f2(1.f, 1.f));
    const float2 r = __ffma2_rn(a, g, x);
    return __fmul2_rn(r, f2(0.5f, 0.5f));
}

template <int ACT>
__device__ __forceinline__ float2 act_pair(float2 v, float slope) {
    if constexpr (ACT == ADSR_ACT_GELU) return gelu_pair(v);
    if constexpr (ACT == ADSR_ACT_LRELU) return f2(v.x > 0.f ? v.x : v.x * slope, v.y > 0.f ? v.y : v.y * slope);
    if constexpr (ACT == ADSR_ACT_RELU) return f2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f));
    return v;
}

template <int ACT>
__global__ void __launch_bounds__(kThreads, 1) tc_gemm_rows_kernel(const __grid_constant__ RowParams rp) {
    const TcGemmParams& p = rp.g;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int a_bytes = rp.ks1 * kPanelBytes;
    uint8_t* a_buf = smem;                                             // n_abuf x [ks1 panels]
    uint8_t* ring = a_buf + rp.n_abuf * a_bytes;                       // 2 rings x n_slots x 16 KB
    uint8_t* stg = ring + 2 * rp.n_slots * kPanelBytes;                    // kStg x 16 KB staging panels
    float* s_bias = reinterpret_cast<float*>(stg + kStg * kPanelBytes);   // [n_tiles * 128]
    float* s_colsum = s_bias + kMaxN;
    float2* s_stat = reinterpret_cast<float2*>(s_colsum + kMaxN);      // [2][4][128]
    RowBarriers* bars = reinterpret_cast<RowBarriers*>(s_stat + 2 * 4 * 128);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int my_tiles = static_cast<int>(blockIdx.x) < p.m_tiles ? (p.m_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    // tile order: forward, or from the LAST tile down (rev) when the producer of A wrote it front to back: the rows it wrote last
    // are the ones still in L2
    auto tile_of = [&](int it) -> int {
        const int t = it * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x);
        return rp.rev ? p.m_tiles - 1 - t : t;
    };
    const bool has_res = p.res != nullptr;

    if ((smem_u32(smem) & 1023u) != 0) __trap();
    for (int i = threadIdx.x; i < p.n_tiles * kBN; i += kThreads) {
        s_bias[i] = p.bias[i];
        s_colsum[i] = p.ln_fold ? p.colsum[i] : 0.f;
    }
    if (warp == kMmaWarp && lane == 0) {
        for (int s = 0; s < kMaxSlots; ++s) {
            for (int w = 0; w < 2; ++w) {
                mbar_init(&bars->b_full[w][s], 1);
                mbar_init(&bars->b_empty[w][s], 1);
            }
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->a_full[b], 1);
            mbar_init(&bars->a_empty[b], 2);                 // both MMA issuers are done with the tile
        }
        for (int b = 0; b < kStg; ++b) {
            mbar_init(&bars->stg_ready[b], kEpiWarps);
            mbar_init(&bars->stg_free[b], 1);
        }
        for (int b = 0; b < kAcc; ++b) {
            mbar_init(&bars->acc_full[b], 1);
            mbar_init(&bars->acc_free[b], kEpiWarps);
        }
        fence_barrier_init();
    }
    if (warp == kALoaderWarp) tmem_alloc<512>(&bars->tmem_base);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = bars->tmem_base;
    // programmatic dependent launch: only the weight loaders and the MMA issuers may run ahead of the predecessor's completion
    pdl_launch_dependents();
    if (warp < kEpiWarps || warp == kALoaderWarp || warp == kStoreWarp) pdl_wait();
    const int n_last = ((p.N - (p.n_tiles - 1) * kBN) + 15) & ~15;     // MMA N of the last N tile

    if (warp == kBLoaderWarp || warp == kBLoaderWarp2) {
        // ============================================================ weight slabs of the N tiles of issuer `me`, (N tile, K slab) in order
        const int me = warp == kBLoaderWarp ? 0 : 1;
        uint8_t* my_ring = ring + me * rp.n_slots * kPanelBytes;
        int slot = 0, c = 0;
        uint32_t phase = 0;
        for (int it = 0; it < my_tiles; ++it) {
            for (int nt = 0; nt < p.n_tiles; ++nt, ++c) {
                if ((c & 1) != me) continue;
                const uint32_t bytes = static_cast<uint32_t>(nt == p.n_tiles - 1 ? n_last : kBN) * 128u;
                const uint8_t* src = p.Bp + static_cast<size_t>(nt) * rp.ks1 * kPanelBytes;
                for (int s = 0; s < rp.ks1; ++s) {
                    mbar_wait(&bars->b_empty[me][slot], phase ^ 1);
                    if (elect_one_sync()) {
                        mbar_arrive_expect_tx(&bars->b_full[me][slot], bytes);
                        bulk_g2s(my_ring + slot * kPanelBytes, src + static_cast<size_t>(s) * kPanelBytes, bytes, &bars->b_full[me][slot]);
                    }
                    __syncwarp();
                    if (++slot == rp.n_slots) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == kALoaderWarp) {
        // ============================================================ A tiles (whole K extent), n_abuf deep
        if (lane == 0) tma_prefetch_desc(&p.tmap_a);
        for (int it = 0; it < my_tiles; ++it) {
            const int ab = rp.n_abuf == 2 ? (it & 1) : 0;
            const int m0 = tile_of(it) * 128;
            mbar_wait(&bars->a_empty[ab], (static_cast<uint32_t>(rp.n_abuf == 2 ? (it >> 1) : it) & 1) ^ 1);
            if (lane == 0) {
                mbar_arrive_expect_tx(&bars->a_full[ab], static_cast<uint32_t>(a_bytes));
                for (int pn = 0; pn < rp.ks1; ++pn)
                    tma_load_2d(a_buf + ab * a_bytes + pn * kPanelBytes, &p.tmap_a, pn * 64, m0, &bars->a_full[ab]);
            }
            __syncwarp();
        }
    } else if (warp == kMmaWarp || warp == kMmaWarp2) {
        // ============================================================ MMA issuers: N tile c (running count) goes to issuer c & 1
        const int me = warp == kMmaWarp ? 0 : 1;
        const uint64_t ring_desc = umma_desc_k_sw128(smem_u32(ring + me * rp.n_slots * kPanelBytes));
        int slot = 0;
        uint32_t phase = 0;
        const uint64_t a_desc0 = umma_desc_k_sw128(smem_u32(a_buf));
        int c = 0;                                                     // running N-tile counter: accumulator = c % 4
        for (int it = 0; it < my_tiles; ++it) {
            const int ab = rp.n_abuf == 2 ? (it & 1) : 0;
            mbar_wait(&bars->a_full[ab], static_cast<uint32_t>(rp.n_abuf == 2 ? (it >> 1) : it) & 1);
            const uint64_t a_desc = a_desc0 + static_cast<uint64_t>(static_cast<uint32_t>(ab) * static_cast<uint32_t>(a_bytes >> 4));
            for (int nt = 0; nt < p.n_tiles; ++nt, ++c) {
                if ((c & 1) != me) continue;
                const int buf = c % kAcc;
                const uint32_t idesc = umma_idesc_bf16_m128(static_cast<uint32_t>(nt == p.n_tiles - 1 ? n_last : kBN));
                const uint32_t d = tmem + static_cast<uint32_t>(buf * kBN);
                mbar_wait(&bars->acc_free[buf], (static_cast<uint32_t>(c / kAcc) & 1) ^ 1);
                tc_fence_after_sync();
                for (int s = 0; s < rp.ks1; ++s) {
                    const int ksteps = min(4, p.k16_total - 4 * s);
                    mbar_wait(&bars->b_full[me][slot], phase);
                    tc_fence_after_sync();
                    if (elect_one_sync()) {
                        const uint64_t adesc = a_desc + static_cast<uint64_t>(s * (kPanelBytes >> 4));
                        const uint64_t bdesc = ring_desc + static_cast<uint64_t>(static_cast<uint32_t>(slot) * (kPanelBytes >> 4));
                        umma_bf16(d, adesc, bdesc, idesc, s == 0 ? 0u : 1u);
                        if (ksteps > 1) umma_bf16(d, adesc + 2, bdesc + 2, idesc, 1u);
                        if (ksteps > 2) umma_bf16(d, adesc + 4, bdesc + 4, idesc, 1u);
                        if (ksteps > 3) umma_bf16(d, adesc + 6, bdesc + 6, idesc, 1u);
                        umma_commit(&bars->b_empty[me][slot]);
                        if (s == rp.ks1 - 1) umma_commit(&bars->acc_full[buf]);
                    }
                    __syncwarp();
                    if (++slot == rp.n_slots) { slot = 0; phase ^= 1; }
                }
            }
            // my MMAs on this A tile are all issued: the commit arrives once they have completed (count 2: both issuers)
            if (elect_one_sync()) umma_commit(&bars->a_empty[ab]);
            __syncwarp();
        }
    } else if (warp == kStoreWarp) {
        // ============================================================ staging panels: residual loads in, output stores out
        if (lane == 0) {
            int q = 0;                                                 // running slice counter: panel = q & 1
            // residual slices are fetched kStg slices ahead; all panels are free at the start
            int rq_it = 0, rq_nt = 0, rq_sl = 0, rq = 0;               // cursor of the next residual slice to fetch
            auto advance = [&](int& it, int& nt, int& sl) {
                const int nn = nt == p.n_tiles - 1 ? n_last : kBN;
                if (++sl * 64 >= nn) { sl = 0; if (++nt == p.n_tiles) { nt = 0; ++it; } }
            };
            auto fetch_res = [&]() {
                if (rq_it >= my_tiles) return;
                const int m0 = tile_of(rq_it) * 128;
                mbar_arrive_expect_tx(&bars->stg_free[rq % kStg], kPanelBytes);
                tma_load_2d(stg + (rq % kStg) * kPanelBytes, &p.tmap_res, rq_nt * kBN + rq_sl * 64, m0, &bars->stg_free[rq % kStg]);
                advance(rq_it, rq_nt, rq_sl);
                ++rq;
            };
            if (has_res)
                for (int i = 0; i < kStg; ++i) fetch_res();
            int it = 0, nt = 0, sl = 0;
            while (it < my_tiles) {
                const int m0 = tile_of(it) * 128;
                mbar_wait(&bars->stg_ready[q % kStg], static_cast<uint32_t>(q / kStg) & 1);
                tma_store_2d_box(&p.tmap_out, stg + (q % kStg) * kPanelBytes, p.ocol0 + nt * kBN + sl * 64, m0);
                bulk_commit_group();
                if (q >= kStg - 2) {
                    // at most kStg - 2 stores still reading: the panel of slice q - (kStg - 2) can be refilled
                    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kStg - 2) : "memory");
                    if (has_res) fetch_res();
                    else mbar_arrive(&bars->stg_free[(q - (kStg - 2)) % kStg]);
                }
                advance(it, nt, sl);
                ++q;
            }
            bulk_wait_group0();
        }
    } else if (warp < kEpiWarps) {
        // ============================================================ epilogue: quadrant (TMEM lanes) x 16-column group
        const int quad = warp & 3, grp = warp >> 2;
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const int r = quad * 32 + lane;
        const int rsw = r & 7;
        const uint32_t stg_row = static_cast<uint32_t>(r * 128);
        const float2 alpha2 = f2(p.alpha, p.alpha);
        int q = 0, c = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int row = tile_of(it) * 128 + r;
            float rstd = 1.f, nrm = 0.f;
            if (p.ln_fold && row < p.M) {
                const float2* sp = p.stats_in + static_cast<long long>(row) * p.stats_in_stride;
                float s1 = 0.f, s2 = 0.f;
                for (int k = 0; k < p.stats_in_slots; ++k) {
                    const float2 v = __ldg(sp + k);
                    s1 += v.x;
                    s2 += v.y;
                }
                const float inv_c = 1.0f / static_cast<float>(p.ln_C);
                const float mean = s1 * inv_c;
                rstd = rsqrtf(fmaxf(s2 * inv_c - mean * mean, 0.f) + p.ln_eps);
                nrm = -mean * rstd;
            }
            const float2 rstd2 = f2(rstd, rstd), nrm2 = f2(nrm, nrm);
            for (int nt = 0; nt < p.n_tiles; ++nt, ++c) {
                const int buf = c % kAcc;
                const int nn = nt == p.n_tiles - 1 ? n_last : kBN;
                const uint32_t taddr = tmem + static_cast<uint32_t>(buf * kBN) + lane_off;
                mbar_wait(&bars->acc_full[buf], static_cast<uint32_t>(c / kAcc) & 1);
                tc_fence_after_sync();
                float2 st = f2(0.f, 0.f), sq = f2(0.f, 0.f);
                for (int sl = 0; sl * 64 < nn; ++sl, ++q) {
                    const int cs = sl * 64 + grp * 16;                 // my 16 columns inside the N tile
                    const bool active = cs < nn;
                    uint32_t raw[16];
                    if (active) tmem_ld16(taddr + static_cast<uint32_t>(cs), raw);
                    // the panel: residual landed (res) / previous store has read it (no res)
                    mbar_wait(&bars->stg_free[q % kStg], has_res ? (static_cast<uint32_t>(q / kStg) & 1) : ((static_cast<uint32_t>(q / kStg) & 1) ^ 1));
                    if (active) {
                        tmem_ld_wait();
                        const float* bp = s_bias + nt * kBN + cs;
                        const float* cp = s_colsum + nt * kBN + cs;
                        const uint32_t sbase = smem_u32(stg + (q % kStg) * kPanelBytes) + stg_row;
#pragma unroll
                        for (int o = 0; o < 2; ++o) {                  // 8 columns = one 16-byte chunk of the swizzled row
                            const uint32_t sa = sbase + static_cast<uint32_t>((((grp * 2 + o) ^ rsw) << 4));
                            uint32_t rw[4] = {0u, 0u, 0u, 0u};
                            if (has_res)
                                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(rw[0]), "=r"(rw[1]), "=r"(rw[2]), "=r"(rw[3]) : "r"(sa));
                            uint32_t pk[4];
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                const int e = 8 * o + 2 * h;
                                float2 v = f2(__uint_as_float(raw[e]), __uint_as_float(raw[e + 1]));
                                const float2 bb = f2(bp[e], bp[e + 1]);
                                if (p.ln_fold) v = __ffma2_rn(rstd2, v, __ffma2_rn(nrm2, f2(cp[e], cp[e + 1]), bb));
                                else v = __fadd2_rn(v, bb);
                                v = act_pair<ACT>(v, p.slope);
                                v = __ffma2_rn(v, alpha2, f2(bf16_lo(rw[h]), bf16_hi(rw[h])));
                                st = __fadd2_rn(st, v);
                                sq = __ffma2_rn(v, v, sq);
                                pk[h] = pack_bf16x2(v.x, v.y);
                            }
                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sa), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->stg_ready[q % kStg]);
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->acc_free[buf]);
                if (p.stats_out != nullptr) {
                    // (sum, sumsq) of my row over this N tile: the four column groups meet in shared memory, two slots per tile
                    float2* ss = s_stat + (c & 1) * 512;
                    ss[grp * 128 + r] = f2(st.x + st.y, sq.x + sq.y);
                    named_bar_sync(1 + quad, 128);
                    if ((grp & 1) == 0 && row < p.M) {
                        const float2 a = ss[grp * 128 + r], b = ss[(grp + 1) * 128 + r];
                        p.stats_out[static_cast<long long>(row) * p.stats_out_stride + p.stats_out_slot0 + nt * 2 + (grp >> 1)] =
                            f2(a.x + b.x, a.y + b.y);
                    }
                }
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == kALoaderWarp) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem);
    }
}

}  // namespace

// true when the row-tile kernel covers this GEMM (the packed weight must have BN = 128)
bool tc_gemm_rows_eligible(const TcGemmParams& p) {
    // the last N tile only computes n_last columns: everything stored must lie inside them
    const int n_last = ((p.N - (p.n_tiles - 1) * kBN) + 15) & ~15;
    if (p.N <= (p.n_tiles - 1) * kBN || p.n_store > (p.n_tiles - 1) * kBN + n_last) return false;
    return p.use_tma && !p.conv && p.out_mode == ADSR_OUT_ROWS && p.BN == kBN && p.num_k_stages <= 5 && (p.ocol0 % 8) == 0 &&
           (p.n_store % 8) == 0 && p.n_tiles * kBN <= kMaxN && p.n_store > (p.n_tiles - 1) * kBN && p.n_store <= p.n_tiles * kBN &&
           (p.ldo % 8) == 0 && (p.res == nullptr || (p.ldres % 8) == 0);
}

int launch_tc_gemm_rows(TcGemmParams& g, int num_sms, cudaStream_t stream) {
    RowParams rp{};
    rp.rev = g.rev_tiles;
    rp.ks1 = g.num_k_stages;
    const int fixed = kStg * kPanelBytes + 2 * kMaxN * 4 + 2 * 4 * 128 * 8 + static_cast<int>(sizeof(RowBarriers)) + 64;
    const int a_bytes = rp.ks1 * kPanelBytes;
    rp.n_abuf = (kSmemLimit - fixed - 2 * a_bytes) / kPanelBytes >= 6 ? 2 : 1;     // double-buffer A only if >= 3 slots per ring remain
    rp.n_slots = ((kSmemLimit - fixed - rp.n_abuf * a_bytes) / kPanelBytes) / 2;    // per ring
    if (rp.n_slots > kMaxSlots) rp.n_slots = kMaxSlots;
    if (rp.n_slots < 2) return ADSR_ERR_BAD_SHAPE;
    int st = encode_tmap_rows_bf16(&g.tmap_out, g.out, g.M, g.ocol0 + g.n_store, g.ldo);
    if (st != ADSR_OK) return st;
    if (g.res != nullptr) {
        st = encode_tmap_rows_bf16(&g.tmap_res, g.res, g.M, g.N, g.ldres);
        if (st != ADSR_OK) return st;
    }
    rp.g = g;
    const int smem_bytes = rp.n_abuf * a_bytes + 2 * rp.n_slots * kPanelBytes + fixed;
    const int grid = g.m_tiles < num_sms ? g.m_tiles : num_sms;
    auto launch = [&](auto kernel) -> int {
        if (ensure_dynamic_smem(kernel, smem_bytes) != cudaSuccess) return ADSR_ERR_CUDA;
        return launch_pdl(kernel, dim3(grid), dim3(kThreads), static_cast<size_t>(smem_bytes), stream, rp) == cudaSuccess ? ADSR_OK : ADSR_ERR_LAUNCH;
    };
    switch (g.act) {
        case ADSR_ACT_NONE: return launch(tc_gemm_rows_kernel<ADSR_ACT_NONE>);
        case ADSR_ACT_LRELU: return launch(tc_gemm_rows_kernel<ADSR_ACT_LRELU>);
        case ADSR_ACT_GELU: return launch(tc_gemm_rows_kernel<ADSR_ACT_GELU>);
        case ADSR_ACT_RELU: return launch(tc_gemm_rows_kernel<ADSR_ACT_RELU>);
    }
    return ADSR_ERR_BAD_SHAPE;
}

}  // namespace adsr
