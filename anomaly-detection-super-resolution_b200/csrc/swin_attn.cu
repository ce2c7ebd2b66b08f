// Fused attention half of a Swin block for sm_100a (8 x 8 windows, the DRCT-L shape):
//     y = x + proj( WindowAttention( LayerNorm(x) ) )                               (src/drct.py:478-509, 271-302)
// i.e. norm1 (folded) -> qkv Linear -> cyclic shift + window_partition -> softmax(q k^T * scale + rel-pos bias + mask) v
// -> window_reverse + un-shift -> proj Linear -> + shortcut, in ONE persistent warp-specialised kernel.  Neither the
// q|k|v rows (1.1 - 1.9 KB per token) nor the attention output ever touch HBM.
//
// A CTA owns one TILE = two consecutive windows = 128 token rows and walks over the heads:
//   * 4 producer warps gather the raw x rows of the tile through the closed-form shifted-window index map
//     (src/drct.py:483, 193-204) with zero-filling 16-byte cp.async into K-major 128-byte-swizzled panels (A operand);
//   * per head h the MMA warp computes  [q_h | k_h | v_h] = x W_h^T  (tcgen05.mma SS, N = 3 hdp, weights streamed from
//     L2 through a ring of [3 hdp x 64] slabs); 16 epilogue warps apply the folded LayerNorm + bias and write q back
//     IN PLACE into TMEM as bf16 (A operand of S), k and v as bf16 into shared-memory operand panels;
//   * S = q k^T (tcgen05.mma TS, M = N = 128; the off-diagonal 64 x 64 blocks belong to the other window and are never
//     used) lands on the dead k|v accumulator columns; softmax: four threads per query row, relative-position bias from a
//     shared-memory table, -100 mask from per-row 64-bit same-region words (src/drct.py:449-470); unnormalised bf16
//     probabilities go back in place into TMEM;  O = P v  (TS, v as MN-major B operand);
//   * fuse_proj: O is normalised and written back in place as bf16 and  Y += O_h Wp_h^T  (TS) accumulates the proj
//     Linear over the heads in a persistent TMEM accumulator; the last epilogue adds bias and the shortcut (re-read from
//     L2), stores y at the ORIGINAL token rows (window_reverse + un-shift are the inverse permutation) and leaves the
//     per-row (sum, sumsq) for the norm2 fold of the MLP kernel.
//     Otherwise (heads of 80 / 128 padded channels: TMEM cannot hold the proj accumulator next to q|k|v) the normalised
//     O rows are stored to `out` [M, nH * hdp] and the row-tile GEMM applies proj.
#include "adsr_kernels.h"
#include "ptx.cuh"

namespace adsr {

namespace {

constexpr int kEpiWarps = 16;                 // warps 0..15: quadrant (TMEM lanes) = warp & 3, column group = warp >> 2
constexpr int kProducerWarp0 = 16;            // warps 16..19: x-tile gather
constexpr int kMmaWarp = 20;
constexpr int kW1LoaderWarp = 21;             // qkv weight slabs, TMEM alloc
constexpr int kW2LoaderWarp = 22;             // proj weight slabs
constexpr int kThreads = 23 * 32;
constexpr int kPanelBytes = 128 * 128;
constexpr int kMaxSlots = 8;
constexpr int kSmemLimit = 232448;

struct __align__(8) AttnBlockBarriers {
    uint64_t w1_full[kMaxSlots], w1_empty[kMaxSlots];
    uint64_t w2_full[2], w2_empty[2];
    uint64_t x_full, x_empty;
    uint64_t meta_free[2];
    uint64_t qkv_full, qkv_ready, s_full, p_ready, o_full, o_ready;
    uint64_t proj_full, proj_free;
    uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc_m128(uint32_t n, uint32_t b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all_() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8_(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st8_zero_(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ int region_1d_(int t, int L, int shift) { return (t >= L - 8 ? 1 : 0) + (t >= L - shift ? 1 : 0); }
__device__ __forceinline__ float2 f2_(float a, float b) { return make_float2(a, b); }

// optional timeline of CTA 0 (tools/attn_trace.py): trace[(((role * 4 + tile) * 9 + head) * 8 + k)] = clock64()
// roles: 0 = MMA warp, 1 = epilogue warp 0, 2 = producer warp 0; head slot 8 = per-tile events
template <bool TRACE>
__device__ __forceinline__ void tr_ev(long long* trace, int role, int it, int h, int k) {
    if constexpr (TRACE) {
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && it < 4) trace[((role * 4 + it) * 9 + h) * 8 + k] = clock64();
    }
}

template <bool TRACE>
__global__ void __launch_bounds__(kThreads, 1) swin_attn_kernel(const __grid_constant__ SwinAttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int x_bytes = p.ks * kPanelBytes;
    const int op_bytes = p.pan * kPanelBytes;
    const int nqkv = p.nH * 3 * p.hdp;
    uint8_t* x_buf = smem;                                             // [ks panels] raw x rows of the tile (A operand of qkv)
    uint8_t* k_buf = x_buf + x_bytes;                                  // [pan panels] k of the current head (K-major B operand of S)
    uint8_t* v_buf = k_buf + op_bytes;                                 // [pan panels] v of the current head (MN-major B operand of P V)
    uint8_t* ring1 = v_buf + op_bytes;                                 // w1_slots x w1_slot_bytes
    uint8_t* ring2 = ring1 + p.w1_slots * p.w1_slot_bytes;             // w2_slots x w2_slot_bytes
    float* s_bq = reinterpret_cast<float*>(ring2 + p.w2_slots * p.w2_slot_bytes);   // [nqkv] folded qkv bias, per-head q|k|v order
    float* s_cq = s_bq + nqkv;                                         // [nqkv] column sums of the gamma-folded weights
    float* s_bp = s_cq + nqkv;                                         // [cp] proj bias
    float* s_bias = s_bp + p.cp;                                       // [nH][232] rel-pos table * log2(e)
    int* s_tok = reinterpret_cast<int*>(s_bias + p.nH * 232);          // [2][128] token row of each tile row
    uint32_t* s_msk = reinterpret_cast<uint32_t*>(s_tok + 256);        // [2][128][2] "same mask region" bits of the row's 64 keys
    float* s_max = reinterpret_cast<float*>(s_msk + 512);              // [4][128] partial row maxima
    float* s_sum = s_max + 512;                                        // [4][128] partial row sums
    float2* s_stat = reinterpret_cast<float2*>(s_max);                 // [4][128] (sum, sumsq) partials of y -- aliases s_max | s_sum
    AttnBlockBarriers* bars = reinterpret_cast<AttnBlockBarriers*>(s_sum + 512);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int my_tiles = static_cast<int>(blockIdx.x) < p.n_tiles ? (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if ((smem_u32(smem) & 1023u) != 0) __trap();
    for (int i = threadIdx.x; i < nqkv; i += kThreads) {
        s_bq[i] = p.bias_qkv[i];
        s_cq[i] = p.colsum_qkv[i];
    }
    if (p.fuse_proj)
        for (int i = threadIdx.x; i < p.cp; i += kThreads) s_bp[i] = p.bias_p[i];
    for (int i = threadIdx.x; i < 225 * p.nH; i += kThreads) {
        const int h = i % p.nH, e = i / p.nH;
        s_bias[h * 232 + e] = __ldg(p.table + i) * 1.4426950408889634f;
    }
    if (warp == kMmaWarp && lane == 0) {
        for (int s = 0; s < kMaxSlots; ++s) {
            mbar_init(&bars->w1_full[s], 1);
            mbar_init(&bars->w1_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->w2_full[s], 1);
            mbar_init(&bars->w2_empty[s], 1);
            mbar_init(&bars->meta_free[s], kEpiWarps);
        }
        mbar_init(&bars->x_full, 128);
        mbar_init(&bars->x_empty, 1);
        mbar_init(&bars->qkv_full, 1);
        mbar_init(&bars->qkv_ready, kEpiWarps);
        mbar_init(&bars->s_full, 1);
        mbar_init(&bars->p_ready, kEpiWarps);
        mbar_init(&bars->o_full, 1);
        mbar_init(&bars->o_ready, kEpiWarps);
        mbar_init(&bars->proj_full, 1);
        mbar_init(&bars->proj_free, kEpiWarps);
        fence_barrier_init();
    }
    if (warp == kW1LoaderWarp) tmem_alloc<512>(&bars->tmem_base);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = bars->tmem_base;
    const int uq = p.hdp >> 4;                                         // 16-column units (= K16 steps) per head operand

    if (warp >= kProducerWarp0 && warp < kProducerWarp0 + 4) {
        // ============================================================ producers: thread = tile row for the index math,
        // a warp walks along one row's 16-byte chunks for the copies
        const int pw = warp - kProducerWarp0;
        const int r = pw * 32 + lane;
        const int chunks = p.k16 * 2;                                  // 16-byte chunks per row up to the K16 boundary
        const int nwx = p.W >> 3, nW = (p.H >> 3) * nwx;
        const uint32_t x_base = smem_u32(x_buf);
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = it * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x);
            const int mb = it & 1;
            const int win = tile * 2 + (r >> 6);
            const int b = win / nW, w = win - b * nW;
            const int n = r & 63;
            const int ys = (w / nwx) * 8 + (n >> 3), xs = (w % nwx) * 8 + (n & 7);
            int y = ys + p.shift; if (y >= p.H) y -= p.H;
            int x = xs + p.shift; if (x >= p.W) x -= p.W;
            const int tok = (b * p.H + y) * p.W + x;
            uint32_t m0 = 0xffffffffu, m1 = 0xffffffffu;
            if (p.shift > 0) {
                const int wy = (w / nwx) * 8, wx = (w % nwx) * 8;
                const int my = 3 * region_1d_(ys, p.H, p.shift) + region_1d_(xs, p.W, p.shift);
                m0 = m1 = 0u;
                for (int k = 0; k < 64; ++k) {
                    const int rk = 3 * region_1d_(wy + (k >> 3), p.H, p.shift) + region_1d_(wx + (k & 7), p.W, p.shift);
                    if (rk == my) { if (k < 32) m0 |= 1u << k; else m1 |= 1u << (k - 32); }
                }
            }
            mbar_wait(&bars->meta_free[mb], (static_cast<uint32_t>(it >> 1) & 1) ^ 1);
            s_tok[mb * 128 + r] = tok;
            s_msk[(mb * 128 + r) * 2] = m0;
            s_msk[(mb * 128 + r) * 2 + 1] = m1;
            __syncwarp();
            if (pw == 0) tr_ev<TRACE>(p.trace, 2, it, 8, 0);
            mbar_wait(&bars->x_empty, (static_cast<uint32_t>(it) & 1) ^ 1);
            if (pw == 0) tr_ev<TRACE>(p.trace, 2, it, 8, 1);
#pragma unroll 4
            for (int i = 0; i < 32; ++i) {
                const int row = pw * 32 + i;
                const __nv_bfloat16* src = p.x + static_cast<long long>(s_tok[mb * 128 + row]) * p.ldx;
                for (int c = lane; c < chunks; c += 32) {
                    int valid = p.C * 2 - c * 16;
                    valid = valid < 0 ? 0 : (valid > 16 ? 16 : valid);
                    const uint32_t d = x_base + static_cast<uint32_t>((c >> 3) * kPanelBytes) + static_cast<uint32_t>(row * 128) +
                                       static_cast<uint32_t>((((c & 7) ^ (row & 7)) << 4));
                    cp_async16_zfill(d, src + c * 8, valid);
                }
            }
            cp_async_commit_();
            if (pw == 0) tr_ev<TRACE>(p.trace, 2, it, 8, 2);
            cp_async_wait_all_();
            fence_proxy_async_smem();
            mbar_arrive(&bars->x_full);
            if (pw == 0) tr_ev<TRACE>(p.trace, 2, it, 8, 3);
        }
    } else if (warp == kW1LoaderWarp) {
        // ============================================================ qkv weight slabs: (head, K slab, N piece) in order, every tile
        int slot = 0;
        uint32_t phase = 0;
        const size_t slab_bytes = static_cast<size_t>(3 * p.hdp) * 128u;
        for (int it = 0; it < my_tiles; ++it) {
            for (int hs = 0; hs < p.nH * p.ks; ++hs) {
                for (int pc = 0; pc < p.qkv_pieces; ++pc) {
                    const uint32_t bytes = static_cast<uint32_t>(p.qp_rows[pc]) * 128u;
                    const uint8_t* src = p.w1p + static_cast<size_t>(hs) * slab_bytes + (pc ? static_cast<size_t>(p.qp_rows[0]) * 128u : 0u);
                    mbar_wait(&bars->w1_empty[slot], phase ^ 1);
                    if (elect_one_sync()) {
                        mbar_arrive_expect_tx(&bars->w1_full[slot], bytes);
                        bulk_g2s(ring1 + slot * p.w1_slot_bytes, src, bytes, &bars->w1_full[slot]);
                    }
                    __syncwarp();
                    if (++slot == p.w1_slots) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == kW2LoaderWarp) {
        // ============================================================ proj weight slabs: one [cp x 64] slab per head
        if (p.fuse_proj) {
            int slot = 0;
            uint32_t phase = 0;
            const uint32_t bytes = static_cast<uint32_t>(p.cp) * 128u;
            for (int it = 0; it < my_tiles; ++it) {
                for (int h = 0; h < p.nH; ++h) {
                    mbar_wait(&bars->w2_empty[slot], phase ^ 1);
                    if (elect_one_sync()) {
                        mbar_arrive_expect_tx(&bars->w2_full[slot], bytes);
                        bulk_g2s(ring2 + slot * p.w2_slot_bytes, p.w2p + static_cast<size_t>(h) * bytes, bytes, &bars->w2_full[slot]);
                    }
                    __syncwarp();
                    if (++slot == p.w2_slots) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ============================================================ MMA issuer (converged warp, one elected lane issues).
        // The tensor pipe executes in issue order, which is what lets S overlay the dead k|v accumulator columns and the
        // next head's q|k|v overwrite the region once P V has been issued.
        const uint32_t idesc_s = idesc_m128(128, 0);
        const uint32_t idesc_pv = idesc_m128(static_cast<uint32_t>(p.hdp), 1);
        const uint32_t idesc_q0 = idesc_m128(static_cast<uint32_t>(p.qp_rows[0]), 0);
        const uint32_t idesc_q1 = idesc_m128(static_cast<uint32_t>(p.qp_rows[1] > 0 ? p.qp_rows[1] : 16), 0);
        const uint32_t idesc_p0 = idesc_m128(static_cast<uint32_t>(p.pp_rows[0] > 0 ? p.pp_rows[0] : 16), 0);
        const uint32_t idesc_p1 = idesc_m128(static_cast<uint32_t>(p.pp_rows[1] > 0 ? p.pp_rows[1] : 16), 0);
        const uint64_t x_desc = umma_desc_k_sw128(smem_u32(x_buf));
        const uint32_t k_addr = smem_u32(k_buf), v_addr = smem_u32(v_buf);
        const uint64_t ring1_desc = umma_desc_k_sw128(smem_u32(ring1));
        const uint64_t ring2_desc = umma_desc_k_sw128(smem_u32(ring2));
        const uint32_t slot1_units = static_cast<uint32_t>(p.w1_slot_bytes >> 4);
        const uint32_t slot2_units = static_cast<uint32_t>(p.w2_slot_bytes >> 4);
        const uint32_t t_r = tmem + static_cast<uint32_t>(p.col_r), t_s = tmem + static_cast<uint32_t>(p.col_s);
        const uint32_t t_o = tmem + static_cast<uint32_t>(p.col_o), t_p = tmem + static_cast<uint32_t>(p.col_proj);
        int slot1 = 0, slot2 = 0;
        uint32_t ph1 = 0, ph2 = 0;

        auto issue_qkv = [&](int h) {
            for (int s = 0; s < p.ks; ++s) {
                const int ksteps = min(4, p.k16 - 4 * s);
                for (int pc = 0; pc < p.qkv_pieces; ++pc) {
                    mbar_wait(&bars->w1_full[slot1], ph1);
                    tc_fence_after_sync();
                    if (elect_one_sync()) {
                        const uint64_t adesc = x_desc + static_cast<uint64_t>(s * (kPanelBytes >> 4));
                        const uint64_t bdesc = ring1_desc + static_cast<uint64_t>(static_cast<uint32_t>(slot1) * slot1_units);
                        const uint32_t d = pc ? t_r + static_cast<uint32_t>(p.qp_rows[0]) : t_r;
                        const uint32_t idesc = pc ? idesc_q1 : idesc_q0;
                        for (int j = 0; j < ksteps; ++j) umma_bf16(d, adesc + 2 * j, bdesc + 2 * j, idesc, (s > 0 || j > 0) ? 1u : 0u);
                        umma_commit(&bars->w1_empty[slot1]);
                        if (s == p.ks - 1 && pc == p.qkv_pieces - 1) {
                            umma_commit(&bars->qkv_full);
                            if (h == p.nH - 1) umma_commit(&bars->x_empty);   // the x tile is no longer an operand
                        }
                    }
                    __syncwarp();
                    if (++slot1 == p.w1_slots) { slot1 = 0; ph1 ^= 1; }
                }
            }
        };

        for (int it = 0; it < my_tiles; ++it) {
            tr_ev<TRACE>(p.trace, 0, it, 8, 0);
            mbar_wait(&bars->x_full, static_cast<uint32_t>(it) & 1);
            tc_fence_after_sync();
            tr_ev<TRACE>(p.trace, 0, it, 8, 1);
            issue_qkv(0);
            tr_ev<TRACE>(p.trace, 0, it, 8, 2);
            for (int h = 0; h < p.nH; ++h) {
                const uint32_t par = static_cast<uint32_t>(it * p.nH + h) & 1;
                // ---- S = q k^T : q bf16 in TMEM (unit u at column 16 u of the region), k panel in shared memory
                tr_ev<TRACE>(p.trace, 0, it, h, 0);
                mbar_wait(&bars->qkv_ready, par);
                tc_fence_after_sync();
                tr_ev<TRACE>(p.trace, 0, it, h, 1);
                if (elect_one_sync()) {
                    for (int u = 0; u < uq; ++u) {
                        const uint32_t off = static_cast<uint32_t>((u >> 2) * kPanelBytes + (u & 3) * 32);
                        umma_bf16_ts(t_s, t_r + static_cast<uint32_t>(16 * u), umma_desc_k_sw128(k_addr + off), idesc_s, u == 0 ? 0u : 1u);
                    }
                    umma_commit(&bars->s_full);
                }
                __syncwarp();
                // ---- O = P v : P bf16 in TMEM (16 keys per step = 8 packed columns), v rows as MN-major B operand
                mbar_wait(&bars->p_ready, par);
                tc_fence_after_sync();
                tr_ev<TRACE>(p.trace, 0, it, h, 2);
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        umma_bf16_ts(t_o, t_s + static_cast<uint32_t>(8 * k), desc_mn_sw128(v_addr + static_cast<uint32_t>(k * 2048), kPanelBytes),
                                     idesc_pv, k == 0 ? 0u : 1u);
                    umma_commit(&bars->o_full);
                }
                __syncwarp();
                if (p.fuse_proj) {
                    tr_ev<TRACE>(p.trace, 0, it, h, 3);
                    if (h + 1 < p.nH) issue_qkv(h + 1);                  // runs while the epilogue normalises O
                    tr_ev<TRACE>(p.trace, 0, it, h, 4);
                    mbar_wait(&bars->o_ready, par);
                    tr_ev<TRACE>(p.trace, 0, it, h, 5);
                    if (h == 0) mbar_wait(&bars->proj_free, (static_cast<uint32_t>(it) & 1) ^ 1);
                    mbar_wait(&bars->w2_full[slot2], ph2);
                    tc_fence_after_sync();
                    if (elect_one_sync()) {
                        const uint64_t bdesc = ring2_desc + static_cast<uint64_t>(static_cast<uint32_t>(slot2) * slot2_units);
                        for (int u = 0; u < uq; ++u) {
                            const uint32_t acc = (h > 0 || u > 0) ? 1u : 0u;
                            umma_bf16_ts(t_p, t_o + static_cast<uint32_t>(16 * u), bdesc + 2 * u, idesc_p0, acc);
                            if (p.n_pp > 1)
                                umma_bf16_ts(t_p + static_cast<uint32_t>(p.pp_rows[0]), t_o + static_cast<uint32_t>(16 * u),
                                             bdesc + static_cast<uint64_t>(p.pp_rows[0] * 8) + 2 * u, idesc_p1, acc);
                        }
                        umma_commit(&bars->w2_empty[slot2]);
                        if (h == p.nH - 1) umma_commit(&bars->proj_full);
                    }
                    __syncwarp();
                    tr_ev<TRACE>(p.trace, 0, it, h, 6);
                    if (++slot2 == p.w2_slots) { slot2 = 0; ph2 ^= 1; }
                } else {
                    tr_ev<TRACE>(p.trace, 0, it, h, 3);
                    mbar_wait(&bars->o_ready, par);                      // O overlays the q|k|v region: wait until it has been read
                    tc_fence_after_sync();
                    tr_ev<TRACE>(p.trace, 0, it, h, 5);
                    if (h + 1 < p.nH) issue_qkv(h + 1);
                    tr_ev<TRACE>(p.trace, 0, it, h, 6);
                }
            }
        }
    } else if (warp < kEpiWarps) {
        // ============================================================ epilogue / softmax warps
        const int quad = warp & 3, grp = warp >> 2;
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const int r = quad * 32 + lane;                                // tile row; window = r >> 6, slot = r & 63
        const int rsw = r & 7;
        const int n = r & 63;
        const uint32_t kcol0 = static_cast<uint32_t>((r >> 6) * 64);   // my window's keys inside the 128 S columns
        const int qb = ((n >> 3) + 7) * 15 + (n & 7) + 7 - grp * 30;   // my keys are 16 grp .. 16 grp + 15 = key rows 2 grp, 2 grp + 1
        const uint32_t k_row = smem_u32(k_buf) + static_cast<uint32_t>(r * 128);
        const uint32_t v_row = smem_u32(v_buf) + static_cast<uint32_t>(r * 128);
        const float2 sc2 = f2_(p.scale_log2e, p.scale_log2e);
        const uint32_t t_r = tmem + lane_off + static_cast<uint32_t>(p.col_r), t_s = tmem + lane_off + static_cast<uint32_t>(p.col_s);
        const uint32_t t_o = tmem + lane_off + static_cast<uint32_t>(p.col_o), t_p = tmem + lane_off + static_cast<uint32_t>(p.col_proj);

        for (int it = 0; it < my_tiles; ++it) {
            const int mb = it & 1;
            mbar_wait(&bars->x_full, static_cast<uint32_t>(it) & 1);   // s_tok / s_msk of this tile are visible
            const int tok = s_tok[mb * 128 + r];
            const uint32_t mword = s_msk[(mb * 128 + r) * 2 + (grp >> 1)];
            const uint32_t mbits = (mword >> (16 * (grp & 1))) & 0xffffu;
            float rstd, nrm;
            {
                const float2* sp = p.stats_in + static_cast<long long>(tok) * p.stats_in_stride;
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
            const float2 rstd2 = f2_(rstd, rstd), nrm2 = f2_(nrm, nrm);

            for (int h = 0; h < p.nH; ++h) {
                const uint32_t par = static_cast<uint32_t>(it * p.nH + h) & 1;
                // ---- q | k | v of head h: folded LayerNorm + bias -> bf16; q in place (TMEM), k / v into the operand panels
                if (warp == 0) tr_ev<TRACE>(p.trace, 1, it, h, 0);
                mbar_wait(&bars->qkv_full, par);
                tc_fence_after_sync();
                if (warp == 0) tr_ev<TRACE>(p.trace, 1, it, h, 1);
                for (int u = grp; u < 3 * uq; u += 4) {
                    uint32_t raw[16];
                    tmem_ld16(t_r + static_cast<uint32_t>(16 * u), raw);
                    tmem_ld_wait();
                    const float* bp = s_bq + h * 3 * p.hdp + 16 * u;
                    const float* cp = s_cq + h * 3 * p.hdp + 16 * u;
                    uint32_t pk[8];
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const float4 bb = *reinterpret_cast<const float4*>(bp + 4 * q4);
                        const float4 cs = *reinterpret_cast<const float4*>(cp + 4 * q4);
                        const float2 x0 = __ffma2_rn(rstd2, f2_(__uint_as_float(raw[4 * q4]), __uint_as_float(raw[4 * q4 + 1])),
                                                     __ffma2_rn(nrm2, f2_(cs.x, cs.y), f2_(bb.x, bb.y)));
                        const float2 x1 = __ffma2_rn(rstd2, f2_(__uint_as_float(raw[4 * q4 + 2]), __uint_as_float(raw[4 * q4 + 3])),
                                                     __ffma2_rn(nrm2, f2_(cs.z, cs.w), f2_(bb.z, bb.w)));
                        pk[2 * q4] = pack_bf16x2(x0.x, x0.y);
                        pk[2 * q4 + 1] = pack_bf16x2(x1.x, x1.y);
                    }
                    if (u < uq) {
                        tmem_st8_(t_r + static_cast<uint32_t>(16 * u), pk);
                    } else {
                        const bool is_k = u < 2 * uq;
                        const int cu = u - (is_k ? uq : 2 * uq);       // unit inside the operand: columns 16 cu .. 16 cu + 15
                        const uint32_t rowa = (is_k ? k_row : v_row) + static_cast<uint32_t>((cu >> 2) * kPanelBytes);
                        const int c0 = (2 * cu) & 7;
                        st_shared_v4(rowa + static_cast<uint32_t>(((c0 ^ rsw) << 4)), pk[0], pk[1], pk[2], pk[3]);
                        st_shared_v4(rowa + static_cast<uint32_t>((((c0 + 1) ^ rsw) << 4)), pk[4], pk[5], pk[6], pk[7]);
                    }
                }
                tmem_st_wait();
                fence_proxy_async_smem();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->qkv_ready);
                if (warp == 0) tr_ev<TRACE>(p.trace, 1, it, h, 2);

                // ---- softmax over my window's 64 keys, 16 per thread
                mbar_wait(&bars->s_full, par);
                tc_fence_after_sync();
                if (warp == 0) tr_ev<TRACE>(p.trace, 1, it, h, 3);
                float t[16];
                float mx = -INFINITY;
                {
                    uint32_t raw[16];
                    tmem_ld16(t_s + kcol0 + static_cast<uint32_t>(16 * grp), raw);
                    tmem_ld_wait();
                    const float* bias = s_bias + h * 232 + qb;
#pragma unroll
                    for (int k = 0; k < 16; k += 2) {
                        const float2 bv = f2_(bias[-((k >> 3) * 15 + (k & 7))], bias[-(((k + 1) >> 3) * 15 + ((k + 1) & 7))]);
                        const float2 v = __ffma2_rn(f2_(__uint_as_float(raw[k]), __uint_as_float(raw[k + 1])), sc2, bv);
                        t[k] = v.x;
                        t[k + 1] = v.y;
                        mx = fmaxf(mx, fmaxf(v.x, v.y));
                    }
                }
                s_max[grp * 128 + r] = mx;
                named_bar_sync(1 + quad, 128);
                mx = fmaxf(fmaxf(s_max[r], s_max[128 + r]), fmaxf(s_max[256 + r], s_max[384 + r]));
                {
                    const float2 nmx = f2_(-mx, -mx);
                    float2 acc = f2_(0.f, 0.f);
                    uint32_t pk[8];
#pragma unroll
                    for (int k = 0; k < 16; k += 2) {
                        const float2 d = __fadd2_rn(f2_(t[k], t[k + 1]), nmx);
                        float2 e = f2_(ex2_approx(d.x), ex2_approx(d.y));
                        // keys of another mask region get -100 in the reference: their probability is exp(-100) ~ 0
                        if (!((mbits >> k) & 1u)) e.x = 0.f;
                        if (!((mbits >> (k + 1)) & 1u)) e.y = 0.f;
                        acc = __fadd2_rn(acc, e);
                        pk[k >> 1] = pack_bf16x2(e.x, e.y);
                    }
                    s_sum[grp * 128 + r] = acc.x + acc.y;
                    // P in place: key j of the tile -> packed column j / 2; the same keys' slots of the OTHER window get zeros
                    tmem_st8_(t_s + ((kcol0 + static_cast<uint32_t>(16 * grp)) >> 1), pk);
                    tmem_st8_zero_(t_s + (((64u - kcol0) + static_cast<uint32_t>(16 * grp)) >> 1));
                }
                tmem_st_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->p_ready);
                if (warp == 0) tr_ev<TRACE>(p.trace, 1, it, h, 4);

                // ---- O = P v done: normalise
                mbar_wait(&bars->o_full, par);
                tc_fence_after_sync();
                if (warp == 0) tr_ev<TRACE>(p.trace, 1, it, h, 5);
                {
                    const float inv = 1.0f / ((s_sum[r] + s_sum[128 + r]) + (s_sum[256 + r] + s_sum[384 + r]));
                    const float2 inv2 = f2_(inv, inv);
                    for (int u = grp; u < uq; u += 4) {
                        uint32_t raw[16];
                        tmem_ld16(t_o + static_cast<uint32_t>(16 * u), raw);
                        tmem_ld_wait();
                        uint32_t pk[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float2 v = __fmul2_rn(f2_(__uint_as_float(raw[2 * e]), __uint_as_float(raw[2 * e + 1])), inv2);
                            pk[e] = pack_bf16x2(v.x, v.y);
                        }
                        if (p.fuse_proj) {
                            tmem_st8_(t_o + static_cast<uint32_t>(16 * u), pk);
                        } else {
                            uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<long long>(tok) * p.ldo + h * p.hdp + 16 * u);
                            dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                        }
                    }
                }
                tmem_st_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->o_ready);
                if (warp == 0) tr_ev<TRACE>(p.trace, 1, it, h, 6);
            }

            if (p.fuse_proj) {
                // ---- y = proj + bias + shortcut -> original token rows; (sum, sumsq) of the row for norm2.
                // 64-column chunks are staged in the idle k / v panels so that every global access is a coalesced 128-byte
                // row piece: the shortcut chunk is fetched into the panel by cp.async (two chunks ahead, L2 hits), the epilogue
                // adds bias + accumulator in place, then the same four warps copy the panel rows out.  A quadrant (4 warps,
                // 32 tile rows) only touches its own panel rows, so quadrant-wide named barriers are all the sync needed.
                const int nch = (p.cp + 63) >> 6;
                const uint32_t panel0 = smem_u32(k_buf), panel1 = smem_u32(v_buf);
                const int t128 = grp * 32 + lane;
                auto issue_res = [&](int c) {
                    const uint32_t pbase = (c & 1) ? panel1 : panel0;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int id = t128 + 128 * i;
                        const int rq = quad * 32 + (id >> 3), j = id & 7;
                        const int col = 64 * c + 8 * j;
                        int valid = (p.C - col) * 2;
                        valid = valid < 0 ? 0 : (valid > 16 ? 16 : valid);
                        const __nv_bfloat16* src = p.x + static_cast<long long>(s_tok[mb * 128 + rq]) * p.ldx + (valid > 0 ? col : 0);
                        cp_async16_zfill(pbase + static_cast<uint32_t>(rq * 128) + static_cast<uint32_t>(((j ^ (rq & 7)) << 4)), src, valid);
                    }
                    cp_async_commit_();
                };
                issue_res(0);
                if (nch > 1) issue_res(1);
                if (warp == 0) tr_ev<TRACE>(p.trace, 1, it, 8, 0);
                mbar_wait(&bars->proj_full, static_cast<uint32_t>(it) & 1);
                tc_fence_after_sync();
                if (warp == 0) tr_ev<TRACE>(p.trace, 1, it, 8, 1);
                float2 st = f2_(0.f, 0.f), sq = f2_(0.f, 0.f);
                for (int c = 0; c < nch; ++c) {
                    const uint32_t pbase = (c & 1) ? panel1 : panel0;
                    const int col0 = 64 * c + 16 * grp;
                    const bool active = col0 < p.cp;
                    uint32_t raw[16];
                    if (active) tmem_ld16(t_p + static_cast<uint32_t>(col0), raw);
                    if (c + 1 < nch) asm volatile("cp.async.wait_group 1;" ::: "memory");
                    else asm volatile("cp.async.wait_group 0;" ::: "memory");
                    named_bar_sync(1 + quad, 128);                       // the quadrant's shortcut chunk has landed
                    if (active) {
                        tmem_ld_wait();
                        const uint32_t rowa = pbase + static_cast<uint32_t>(r * 128);
#pragma unroll
                        for (int o = 0; o < 2; ++o) {
                            const uint32_t sa = rowa + static_cast<uint32_t>((((2 * grp + o) ^ rsw) << 4));
                            uint32_t rw[4];
                            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(rw[0]), "=r"(rw[1]), "=r"(rw[2]), "=r"(rw[3]) : "r"(sa));
                            const float4 b0 = *reinterpret_cast<const float4*>(s_bp + col0 + 8 * o);
                            const float4 b1 = *reinterpret_cast<const float4*>(s_bp + col0 + 8 * o + 4);
                            const uint32_t* a8 = &raw[8 * o];
                            const float2 v0 = __fadd2_rn(__fadd2_rn(f2_(__uint_as_float(a8[0]), __uint_as_float(a8[1])), f2_(b0.x, b0.y)),
                                                         f2_(bf16_lo(rw[0]), bf16_hi(rw[0])));
                            const float2 v1 = __fadd2_rn(__fadd2_rn(f2_(__uint_as_float(a8[2]), __uint_as_float(a8[3])), f2_(b0.z, b0.w)),
                                                         f2_(bf16_lo(rw[1]), bf16_hi(rw[1])));
                            const float2 v2 = __fadd2_rn(__fadd2_rn(f2_(__uint_as_float(a8[4]), __uint_as_float(a8[5])), f2_(b1.x, b1.y)),
                                                         f2_(bf16_lo(rw[2]), bf16_hi(rw[2])));
                            const float2 v3 = __fadd2_rn(__fadd2_rn(f2_(__uint_as_float(a8[6]), __uint_as_float(a8[7])), f2_(b1.z, b1.w)),
                                                         f2_(bf16_lo(rw[3]), bf16_hi(rw[3])));
                            // columns >= C: zero weight rows, zero bias, zero-filled shortcut -> exact zeros, no effect on the sums
                            st = __fadd2_rn(st, __fadd2_rn(__fadd2_rn(v0, v1), __fadd2_rn(v2, v3)));
                            sq = __ffma2_rn(v0, v0, sq);
                            sq = __ffma2_rn(v1, v1, sq);
                            sq = __ffma2_rn(v2, v2, sq);
                            sq = __ffma2_rn(v3, v3, sq);
                            st_shared_v4(sa, pack_bf16x2(v0.x, v0.y), pack_bf16x2(v1.x, v1.y), pack_bf16x2(v2.x, v2.y), pack_bf16x2(v3.x, v3.y));
                        }
                    }
                    named_bar_sync(1 + quad, 128);                       // the quadrant's y chunk is complete
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int id = t128 + 128 * i;
                        const int rq = quad * 32 + (id >> 3), j = id & 7;
                        const int col = 64 * c + 8 * j;
                        const int nvalid = p.C - col;
                        if (nvalid > 0) {
                            uint4 val;
                            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                                         : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                                         : "r"(pbase + static_cast<uint32_t>(rq * 128) + static_cast<uint32_t>(((j ^ (rq & 7)) << 4))));
                            __nv_bfloat16* dst = p.out + static_cast<long long>(s_tok[mb * 128 + rq]) * p.ldo + col;
                            if (nvalid >= 8) {
                                *reinterpret_cast<uint4*>(dst) = val;
                            } else {
                                const uint32_t w4[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
                                for (int e = 0; e < 8; ++e)
                                    if (e < nvalid) reinterpret_cast<uint16_t*>(dst)[e] = static_cast<uint16_t>((w4[e >> 1] >> (16 * (e & 1))) & 0xffffu);
                            }
                        }
                    }
                    if (c + 2 < nch) {
                        named_bar_sync(1 + quad, 128);                   // the panel has been copied out: refill it
                        issue_res(c + 2);
                    }
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->proj_free);
                if (warp == 0) tr_ev<TRACE>(p.trace, 1, it, 8, 2);
                if (p.stats_out != nullptr) {
                    named_bar_sync(1 + quad, 128);                       // everybody is past the softmax scratch of the last head
                    s_stat[grp * 128 + r] = f2_(st.x + st.y, sq.x + sq.y);
                    named_bar_sync(1 + quad, 128);
                    if (grp == 0) {
                        const float2 a = s_stat[r], b = s_stat[128 + r], c = s_stat[256 + r], d = s_stat[384 + r];
                        p.stats_out[static_cast<long long>(tok) * p.stats_out_stride + p.stats_out_slot0] =
                            f2_((a.x + b.x) + (c.x + d.x), (a.y + b.y) + (c.y + d.y));
                    }
                    named_bar_sync(1 + quad, 128);                       // s_stat aliases s_max | s_sum of the next tile
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->meta_free[mb]);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == kW1LoaderWarp) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem);
    }
}

int fixed_smem_bytes(int nH, int hdp, int cp) {
    return (2 * nH * 3 * hdp + cp + nH * 232) * 4 + 256 * 4 + 512 * 4 + 2 * 512 * 4 + static_cast<int>(sizeof(AttnBlockBarriers)) + 64;
}

}  // namespace

// Static plan of the fused kernel for one block shape.  Returns 0 = not covered (use the separate kernels),
// 1 = qkv + attention fused (out = attention rows [M, nH * hdp], proj by the row-tile GEMM), 2 = whole attention half.
int swin_attn_plan(SwinAttnParams& p, int C, int nH, int hdp, int allow_proj) {
    if (C <= 0 || C > 320 || nH < 1 || nH > 8 || hdp < 32 || hdp > 128 || (hdp % 16) || nH * 3 * hdp > 1024) return 0;
    p.C = C; p.nH = nH; p.hdp = hdp;
    p.ks = (C + 63) / 64;
    p.k16 = (C + 15) / 16;
    p.pan = (hdp + 63) / 64;
    p.cp = (C + 15) / 16 * 16;
    p.fuse_proj = (allow_proj && hdp <= 64 && p.cp + 2 * hdp + 128 <= 512) ? 1 : 0;
    if (p.fuse_proj) {
        p.col_proj = 0; p.col_r = p.cp; p.col_s = p.cp + hdp; p.col_o = p.cp + hdp + 128;
    } else {
        if (3 * hdp > 384) return 0;
        p.col_proj = 0; p.col_r = 0; p.col_s = 384; p.col_o = hdp;     // O overlays the dead k accumulator
    }
    const int n3 = 3 * hdp;
    if (n3 <= 256) { p.qkv_pieces = 1; p.qp_rows[0] = n3; p.qp_rows[1] = 0; }
    else { p.qkv_pieces = 2; p.qp_rows[0] = (n3 / 2 + 15) / 16 * 16; p.qp_rows[1] = n3 - p.qp_rows[0]; }
    if (p.cp <= 256) { p.n_pp = 1; p.pp_rows[0] = p.cp; p.pp_rows[1] = 0; }
    else { p.n_pp = 2; p.pp_rows[0] = (p.cp / 2 + 15) / 16 * 16; p.pp_rows[1] = p.cp - p.pp_rows[0]; }
    p.w1_slot_bytes = (p.qp_rows[0] * 128 + 1023) / 1024 * 1024;       // qp_rows[0] >= qp_rows[1]
    p.w2_slot_bytes = p.fuse_proj ? (p.cp * 128 + 1023) / 1024 * 1024 : 0;
    p.w2_slots = p.fuse_proj ? 1 : 0;
    const int avail = kSmemLimit - p.ks * kPanelBytes - 2 * p.pan * kPanelBytes - fixed_smem_bytes(nH, hdp, p.cp) - p.w2_slots * p.w2_slot_bytes;
    p.w1_slots = avail / p.w1_slot_bytes;
    if (p.w1_slots > kMaxSlots) p.w1_slots = kMaxSlots;
    if (p.w1_slots < 2) return 0;
    if (p.fuse_proj && avail - p.w1_slots * p.w1_slot_bytes >= p.w2_slot_bytes && p.w1_slots >= 4) p.w2_slots = 2;
    return p.fuse_proj ? 2 : 1;
}

int launch_swin_attn(SwinAttnParams& p, int num_sms, cudaStream_t stream) {
    const int nW = (p.H / 8) * (p.W / 8);
    if ((p.H % 8) || (p.W % 8) || ((p.B * nW) & 1) || p.shift < 0 || p.shift >= 8) return ADSR_ERR_BAD_SHAPE;
    if ((reinterpret_cast<uintptr_t>(p.x) & 15) || (reinterpret_cast<uintptr_t>(p.out) & 15) || (p.ldx % 8) || (p.ldo % 8) ||
        (reinterpret_cast<uintptr_t>(p.w1p) & 15) || (reinterpret_cast<uintptr_t>(p.w2p) & 15))
        return ADSR_ERR_BAD_ALIGN;
    if (p.ldx < p.k16 * 16) return ADSR_ERR_BAD_SHAPE;                 // zero-filled chunk addresses stay inside the row pitch
    if (p.fuse_proj ? p.ldo < (p.C + 7) / 8 * 8 : p.ldo < p.nH * p.hdp) return ADSR_ERR_BAD_SHAPE;
    p.n_tiles = p.B * nW / 2;
    const int smem_bytes = p.ks * kPanelBytes + 2 * p.pan * kPanelBytes + p.w1_slots * p.w1_slot_bytes + p.w2_slots * p.w2_slot_bytes +
                           fixed_smem_bytes(p.nH, p.hdp, p.cp);
    if (smem_bytes > kSmemLimit) return ADSR_ERR_BAD_SHAPE;
    const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
    auto launch = [&](auto kernel) -> int {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) return ADSR_ERR_CUDA;
        kernel<<<grid, kThreads, smem_bytes, stream>>>(p);
        return cudaGetLastError() == cudaSuccess ? ADSR_OK : ADSR_ERR_LAUNCH;
    };
    return p.trace != nullptr ? launch(swin_attn_kernel<true>) : launch(swin_attn_kernel<false>);
}

}  // namespace adsr
