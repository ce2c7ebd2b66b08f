// Fused norm1 + qkv Linear + (shifted-)window attention for 8 x 8 windows, TWO HEADS IN FLIGHT  (src/drct.py:478-505, 271-299):
//     att[token, h * hdp ...] = WindowAttention_h( LayerNorm(x) )          (the proj Linear + shortcut follow as a row-tile GEMM)
// Same data path as the attention-only mode of swin_attn.cu (x tile by 4-D TMA boxes through the closed-form shifted-window map,
// per-head q|k|v = x W_h^T on the tensor core, q written back in place into TMEM as bf16, k / v into shared-memory panels,
// S = q k^T and O = P v with P in place in TMEM), but the per-head dependency chain
//     q|k|v ready -> convert -> S -> softmax -> P v -> O
// which left the tensor pipe idle ~70 % of the time is now run TWICE IN PARALLEL: the 16 epilogue warps form two groups of 8 that
// own the even / odd heads, each with its own TMEM region, its own k / v panel set and its own barriers, so one group's
// conversion / softmax fills the other's MMA waits.  The MMA issue is split over two warps as well: one streams the q|k|v
// weight slabs (ring order = head order), the other only issues the short S and P v batches whenever a group is ready for
// them -- neither ever waits behind the other's barrier.  Nothing on the issue paths divides by a run-time value.
#include "adsr_kernels.h"
#include "ptx.cuh"

namespace adsr {

namespace {

constexpr int kEpiWarps = 16;                 // warps 0..15: group = warp >> 3 (head parity), half = (warp >> 2) & 1, quadrant = warp & 3
constexpr int kXLoaderWarp = 16;              // x-tile TMA loads
constexpr int kMmaWarp = 17;                  // S and P v of both groups
constexpr int kWLoaderWarp = 18;              // qkv weight slabs, TMEM alloc
constexpr int kQkvWarp = 19;                  // q|k|v MMAs, head after head
constexpr int kThreads = 20 * 32;
constexpr int kPanelBytes = 128 * 128;
constexpr int kMaxSlots = 10;
constexpr int kSmemLimit = 232448;

struct __align__(8) Attn2Barriers {
    uint64_t w_full[kMaxSlots], w_empty[kMaxSlots];
    uint64_t x_full, x_empty;
    uint64_t qkv_full[2], qkv_ready[2], s_full[2], p_ready[2], o_full[2], o_ready[2];   // per group / TMEM region
    uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t desc_mn_sw128_2(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc2_m128(uint32_t n, uint32_t b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void tm2_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tm2_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tm2_st16_zero(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z)
                 : "memory");
}
__device__ __forceinline__ void tm2_ld32(uint32_t taddr, uint32_t* v) {
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
__device__ __forceinline__ void st2_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ int region2_1d(int t, int L, int shift) { return (t >= L - 8 ? 1 : 0) + (t >= L - shift ? 1 : 0); }
__device__ __forceinline__ float2 f22(float a, float b) { return make_float2(a, b); }

__global__ void __launch_bounds__(kThreads, 1) swin_attn2_kernel(const __grid_constant__ SwinAttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int x_bytes = p.ks * kPanelBytes;
    const int op_bytes = p.pan * kPanelBytes;
    const int nqkv = p.nH * 3 * p.hdp;
    uint8_t* x_buf = smem;                                             // [ks panels] raw x rows of the tile (A operand of qkv)
    uint8_t* kv_buf = x_buf + x_bytes;                                 // [2 groups][k: pan panels | v: pan panels]
    uint8_t* ring = kv_buf + 4 * op_bytes;                             // qkv slabs: w_slots x w_slot_bytes
    float* s_bq = reinterpret_cast<float*>(ring + p.w_slots * p.w_slot_bytes);   // [nqkv] folded qkv bias, per-head q|k|v order
    float* s_cq = s_bq + nqkv;                                         // [nqkv] column sums of the gamma-folded weights
    float* s_bias = s_cq + nqkv;                                       // [nH][232] rel-pos table * log2(e)
    int* s_tok = reinterpret_cast<int*>(s_bias + p.nH * 232);          // [2 groups][2][128] token row of each tile row
    float* s_max = reinterpret_cast<float*>(s_tok + 512);              // [2 groups][2 halves][128] partial row maxima
    float* s_sum = s_max + 512;                                        // [2 groups][2 halves][128] partial row sums
    Attn2Barriers* bars = reinterpret_cast<Attn2Barriers*>(s_sum + 512);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int my_tiles = static_cast<int>(blockIdx.x) < p.n_tiles ? (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int total_heads = my_tiles * p.nH;

    if ((smem_u32(smem) & 1023u) != 0) __trap();
    for (int i = threadIdx.x; i < nqkv; i += kThreads) {
        s_bq[i] = p.bias_qkv[i];
        s_cq[i] = p.colsum_qkv[i];
    }
    for (int i = threadIdx.x; i < 225 * p.nH; i += kThreads) {
        const int h = i % p.nH, e = i / p.nH;
        s_bias[h * 232 + e] = __ldg(p.table + i) * 1.4426950408889634f;
    }
    if (warp == kMmaWarp && lane == 0) {
        for (int s = 0; s < kMaxSlots; ++s) {
            mbar_init(&bars->w_full[s], 1);
            mbar_init(&bars->w_empty[s], 1);
        }
        mbar_init(&bars->x_full, 1);
        mbar_init(&bars->x_empty, 1);
        for (int g = 0; g < 2; ++g) {
            mbar_init(&bars->qkv_full[g], 1);
            mbar_init(&bars->qkv_ready[g], 8);
            mbar_init(&bars->s_full[g], 1);
            mbar_init(&bars->p_ready[g], 8);
            mbar_init(&bars->o_full[g], 1);
            mbar_init(&bars->o_ready[g], 8);
        }
        fence_barrier_init();
    }
    if (warp == kWLoaderWarp) tmem_alloc<512>(&bars->tmem_base);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = bars->tmem_base;
    // programmatic dependent launch: set-up and weight loads run under the predecessor's tail; warps that touch activations / row
    // statistics wait for it to complete
    pdl_launch_dependents();
    if (warp < kEpiWarps || warp == kXLoaderWarp) pdl_wait();
    const int uq = p.hdp >> 4;                                         // 16-column units (= K16 steps) per head operand
    const int nwx = p.W >> 3, nW = (p.H >> 3) * nwx;

    if (warp == kXLoaderWarp) {
        // ============================================================ x tile: one TMA box per (window, R x R token block, panel);
        // tile row = window * 64 + block * R*R + y' * R + x'  (block-major slot order; R = 8: the usual y * 8 + x)
        if (lane == 0) tma_prefetch_desc(&p.tmap_x);
        const int R = p.box_r, nb = 8 / R;
        const int combos = 2 * nb * nb;
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = it * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x);
            mbar_wait(&bars->x_empty, (static_cast<uint32_t>(it) & 1) ^ 1);
            if (lane == 0) mbar_arrive_expect_tx(&bars->x_full, static_cast<uint32_t>(x_bytes));
            __syncwarp();
            for (int cb = lane; cb < combos; cb += 32) {
                const int w2 = cb / (nb * nb), blk = cb - w2 * nb * nb;
                const int by = blk / nb, bx = blk - by * nb;
                const int win = tile * 2 + w2;
                const int b = win / nW, w = win - b * nW;
                int y = (w / nwx) * 8 + by * R + p.shift; if (y >= p.H) y -= p.H;
                int x = (w % nwx) * 8 + bx * R + p.shift; if (x >= p.W) x -= p.W;
                uint8_t* dst = x_buf + (w2 * 64 + blk * R * R) * 128;
                for (int pn = 0; pn < p.ks; ++pn) tma_load_4d(dst + pn * kPanelBytes, &p.tmap_x, pn * 64, x, y, b, &bars->x_full);
            }
            __syncwarp();
        }
    } else if (warp == kWLoaderWarp) {
        // ============================================================ qkv weight slabs: (head, K slab, N piece) in order, every tile
        int slot = 0;
        uint32_t phase = 0;
        const size_t slab_bytes = static_cast<size_t>(3 * p.hdp) * 128u;
        for (int it = 0; it < my_tiles; ++it) {
            for (int hs = 0; hs < p.nH * p.ks; ++hs) {
                int row0 = 0;
                for (int pc = 0; pc < p.qkv_pieces; ++pc) {
                    const uint32_t bytes = static_cast<uint32_t>(p.qp_rows[pc]) * 128u;
                    mbar_wait(&bars->w_empty[slot], phase ^ 1);
                    if (elect_one_sync()) {
                        mbar_arrive_expect_tx(&bars->w_full[slot], bytes);
                        bulk_g2s(ring + slot * p.w_slot_bytes, p.w1p + static_cast<size_t>(hs) * slab_bytes + static_cast<size_t>(row0) * 128u, bytes,
                                 &bars->w_full[slot]);
                    }
                    __syncwarp();
                    row0 += p.qp_rows[pc];
                    if (++slot == p.w_slots) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == kQkvWarp) {
        // ============================================================ q|k|v MMAs: head g into region g & 1 as soon as the region's
        // previous head (g - 2) has left it; ring slabs are consumed in head order
        const uint64_t x_desc = umma_desc_k_sw128(smem_u32(x_buf));
        const uint64_t ring_desc = umma_desc_k_sw128(smem_u32(ring));
        const uint32_t slot_units = static_cast<uint32_t>(p.w_slot_bytes >> 4);
        const uint32_t idesc_q0 = idesc2_m128(static_cast<uint32_t>(p.qp_rows[0]), 0);
        const uint32_t idesc_q1 = idesc2_m128(static_cast<uint32_t>(p.qp_rows[1] > 0 ? p.qp_rows[1] : 16), 0);
        int slot = 0;
        uint32_t wph = 0;
        int g = 0;
        for (int it = 0; it < my_tiles; ++it) {
            mbar_wait(&bars->x_full, static_cast<uint32_t>(it) & 1);
            for (int h = 0; h < p.nH; ++h, ++g) {
                const int rg = g & 1;
                if (g >= 2) mbar_wait(&bars->o_ready[rg], static_cast<uint32_t>((g - 2) >> 1) & 1);
                const uint32_t dst0 = tmem + static_cast<uint32_t>(rg * p.rsz);
                for (int s = 0; s < p.ks; ++s) {
                    for (int pc = 0; pc < p.qkv_pieces; ++pc) {
                        mbar_wait(&bars->w_full[slot], wph);
                        tc_fence_after_sync();
                        if (elect_one_sync()) {
                            const int ksteps = min(4, p.k16 - 4 * s);
                            const uint32_t d = pc ? dst0 + static_cast<uint32_t>(p.qp_rows[0]) : dst0;
                            const uint64_t adesc = x_desc + static_cast<uint64_t>(s * (kPanelBytes >> 4));
                            const uint64_t bdesc = ring_desc + static_cast<uint64_t>(static_cast<uint32_t>(slot) * slot_units);
                            const uint32_t idesc = pc ? idesc_q1 : idesc_q0;
                            umma_bf16(d, adesc, bdesc, idesc, s > 0 ? 1u : 0u);
                            if (ksteps > 1) umma_bf16(d, adesc + 2, bdesc + 2, idesc, 1u);
                            if (ksteps > 2) umma_bf16(d, adesc + 4, bdesc + 4, idesc, 1u);
                            if (ksteps > 3) umma_bf16(d, adesc + 6, bdesc + 6, idesc, 1u);
                            umma_commit(&bars->w_empty[slot]);
                            if (s == p.ks - 1 && pc == p.qkv_pieces - 1) {
                                umma_commit(&bars->qkv_full[rg]);
                                if (h == p.nH - 1) umma_commit(&bars->x_empty);    // the x tile is no longer an operand
                            }
                        }
                        __syncwarp();
                        if (++slot == p.w_slots) { slot = 0; wph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ============================================================ S = q k^T and O = P v of both groups, whichever is ready
        const uint32_t idesc_s = idesc2_m128(128, 0);
        const uint32_t idesc_pv = idesc2_m128(static_cast<uint32_t>(p.hdp), 1);
        const uint32_t kv0 = smem_u32(kv_buf);
        auto ready = [&](uint64_t* bar, uint32_t parity) -> bool {     // one lane polls, the warp stays converged
            uint32_t ok = 0;
            if (lane == 0) ok = mbar_test_wait(bar, parity) ? 1u : 0u;
            return __shfl_sync(0xffffffffu, ok, 0) != 0;
        };
        int gs = 0, gp = 0;                                            // next head whose S / P v is to be issued
        while (gp < total_heads) {
            if (gs < total_heads && gs < gp + 2) {
                const int rg = gs & 1;
                if (ready(&bars->qkv_ready[rg], static_cast<uint32_t>(gs >> 1) & 1)) {
                    // S = q k^T : q bf16 in TMEM (unit u at column 16 u of the region), k panel of the group in shared memory
                    tc_fence_after_sync();
                    const uint32_t t_r = tmem + static_cast<uint32_t>(rg * p.rsz);
                    const uint64_t k_desc = umma_desc_k_sw128(kv0 + static_cast<uint32_t>(rg * 2 * op_bytes));
                    if (elect_one_sync()) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (u < uq)
                                umma_bf16_ts(t_r + static_cast<uint32_t>(p.hdp), t_r + static_cast<uint32_t>(16 * u),
                                             k_desc + static_cast<uint64_t>(((u >> 2) * kPanelBytes + (u & 3) * 32) >> 4), idesc_s, u == 0 ? 0u : 1u);
                        }
                        umma_commit(&bars->s_full[rg]);
                    }
                    __syncwarp();
                    ++gs;
                    continue;
                }
            }
            if (gp < gs) {
                const int rg = gp & 1;
                if (ready(&bars->p_ready[rg], static_cast<uint32_t>(gp >> 1) & 1)) {
                    // O = P v : P bf16 in TMEM (16 keys per step = 8 packed columns), v rows as MN-major B operand; O lands on the
                    // q columns (q was consumed by S, which has completed: the softmax ran on it)
                    tc_fence_after_sync();
                    const uint32_t t_r = tmem + static_cast<uint32_t>(rg * p.rsz);
                    const uint32_t v_addr = kv0 + static_cast<uint32_t>((rg * 2 + 1) * op_bytes);
                    if (elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            umma_bf16_ts(t_r, t_r + static_cast<uint32_t>(p.hdp + 8 * k), desc_mn_sw128_2(v_addr + static_cast<uint32_t>(k * 2048), kPanelBytes),
                                         idesc_pv, k == 0 ? 0u : 1u);
                        umma_commit(&bars->o_full[rg]);
                    }
                    __syncwarp();
                    ++gp;
                }
            }
        }
    } else if (warp < kEpiWarps) {
        // ============================================================ epilogue / softmax warps: group = head parity
        const int grp = warp >> 3, half = (warp >> 2) & 1, quad = warp & 3;
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const int r = quad * 32 + lane;                                // tile row; window = r >> 6, slot = r & 63
        const int rsw = r & 7;
        const int n = r & 63;
        const uint32_t kcol0 = static_cast<uint32_t>((r >> 6) * 64);   // my window's keys inside the 128 S columns
        // slot -> (y, x) inside the window (block-major for R = 4, see the x loader); my 32 keys are slots 32 half .. 32 half + 31:
        // R = 8: key rows 4 half .. 4 half + 3;  R = 4: the 4 x 4 blocks 2 half (left) and 2 half + 1 (right) = the same rows
        const bool blk4 = p.box_r == 4;
        const int ny = blk4 ? ((n >> 5) << 2) + ((n >> 2) & 3) : (n >> 3);
        const int nx = blk4 ? (((n >> 4) & 1) << 2) + (n & 3) : (n & 7);
        const int ky0 = 4 * half;
        const int qb = (ny - ky0 + 7) * 15 + (nx + 7);                 // bias index of key (ky0, 0); key (ky0 + dy, kx): - (15 dy + kx)
        const uint32_t kv_g = smem_u32(kv_buf) + static_cast<uint32_t>(grp * 2 * op_bytes);
        const uint32_t k_row = kv_g + static_cast<uint32_t>(r * 128);
        const uint32_t v_row = k_row + static_cast<uint32_t>(op_bytes);
        const float2 sc2 = f22(p.scale_log2e, p.scale_log2e);
        const uint32_t t_r = tmem + lane_off + static_cast<uint32_t>(grp * p.rsz);
        const uint32_t t_s = t_r + static_cast<uint32_t>(p.hdp);
        const int pair_bar = 1 + grp * 4 + quad;                       // the two warps (key / column halves) that share my rows
        const int t64 = half * 32 + lane;
        float* my_max = s_max + (grp * 2 + half) * 128 + r;            // other half: +-128 floats; s_sum: +512 floats
        const int other = half ? -128 : 128;
        int* tok_g = s_tok + grp * 256;
        const uint32_t cpr_magic = (65536u + static_cast<uint32_t>(p.hdp >> 3) - 1u) / static_cast<uint32_t>(p.hdp >> 3);   // id / (hdp / 8), id < 512

        // ---- per-tile set-up of my row: slot -> token (closed form of roll + window_partition), mask bits of my 32 keys, LayerNorm
        // statistics.  The (sum, sumsq) slots are a COLD global load, so the set-up of the group's next tile runs while the current
        // tile waits for its first S anyway
        auto row_setup = [&](int tile, int& tok_o, uint32_t& mbits_o, float& rstd_o, float& nrm_o) {
            const int win = tile * 2 + (r >> 6);
            const int b = win / nW, w = win - b * nW;
            const int wy = (w / nwx) * 8, wx = (w % nwx) * 8;
            const int ys = wy + ny, xs = wx + nx;
            int y = ys + p.shift; if (y >= p.H) y -= p.H;
            int x = xs + p.shift; if (x >= p.W) x -= p.W;
            const int tok = (b * p.H + y) * p.W + x;
            uint32_t mb_ = 0xffffffffu;
            if (p.shift > 0) {
                // key kk of my 32: row ky0 + dy(kk), column kx(kk); it counts iff its row AND its column region equal mine
                const int my_ry = region2_1d(ys, p.H, p.shift), my_rx = region2_1d(xs, p.W, p.shift);
                uint32_t xmask = 0, ymask = 0;
                for (int j = 0; j < 8; ++j) xmask |= (region2_1d(wx + j, p.W, p.shift) == my_rx ? 1u : 0u) << j;
                for (int j = 0; j < 4; ++j) ymask |= (region2_1d(wy + ky0 + j, p.H, p.shift) == my_ry ? 1u : 0u) << j;
                mb_ = 0u;
                for (int kk = 0; kk < 32; ++kk) {
                    const int dy = blk4 ? ((kk >> 2) & 3) : (kk >> 3);
                    const int kx = blk4 ? (((kk >> 4) << 2) + (kk & 3)) : (kk & 7);
                    if (((ymask >> dy) & 1u) && ((xmask >> kx) & 1u)) mb_ |= 1u << kk;
                }
            }
            const float2* sp = p.stats_in + static_cast<long long>(tok) * p.stats_in_stride;
            float s1 = 0.f, s2 = 0.f;
            for (int k = 0; k < p.stats_in_slots; ++k) {
                const float2 v = __ldg(sp + k);
                s1 += v.x;
                s2 += v.y;
            }
            const float inv_c = 1.0f / static_cast<float>(p.C);
            const float mean = s1 * inv_c;
            const float rs = rsqrtf(fmaxf(s2 * inv_c - mean * mean, 0.f) + p.ln_eps);
            tok_o = tok; mbits_o = mb_; rstd_o = rs; nrm_o = -mean * rs;
        };
        int it = 0, h = grp;
        while (h >= p.nH) { h -= p.nH; ++it; }
        int cur_it = -1, set_it = -1;                                  // tile of the current values / of the *_n values
        int tok_n = 0;
        float rstd = 0.f, nrm = 0.f, rstd_n = 0.f, nrm_n = 0.f;
        uint32_t mbits = 0xffffffffu, mbits_n = 0xffffffffu;
        for (int g = grp; g < total_heads; g += 2) {
            const uint32_t par = static_cast<uint32_t>(g >> 1) & 1;
            const int mb = it & 1;
            const bool first_of_tile = it != cur_it;
            if (first_of_tile) {
                if (set_it != it) row_setup(it * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x), tok_n, mbits_n, rstd_n, nrm_n);
                cur_it = it;
                mbits = mbits_n; rstd = rstd_n; nrm = nrm_n;
                if (half == 0) tok_g[mb * 128 + r] = tok_n;
            }
            const float2 rstd2 = f22(rstd, rstd), nrm2 = f22(nrm, nrm);

            // ---- q | k | v of head h: folded LayerNorm + bias -> bf16; q in place (TMEM), k / v into my group's operand panels
            mbar_wait(&bars->qkv_full[grp], par);
            tc_fence_after_sync();
            for (int u = half; u < 3 * uq; u += 2) {
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
                    const float2 x0 = __ffma2_rn(rstd2, f22(__uint_as_float(raw[4 * q4]), __uint_as_float(raw[4 * q4 + 1])),
                                                 __ffma2_rn(nrm2, f22(cs.x, cs.y), f22(bb.x, bb.y)));
                    const float2 x1 = __ffma2_rn(rstd2, f22(__uint_as_float(raw[4 * q4 + 2]), __uint_as_float(raw[4 * q4 + 3])),
                                                 __ffma2_rn(nrm2, f22(cs.z, cs.w), f22(bb.z, bb.w)));
                    pk[2 * q4] = pack_bf16x2(x0.x, x0.y);
                    pk[2 * q4 + 1] = pack_bf16x2(x1.x, x1.y);
                }
                if (u < uq) {
                    tm2_st8(t_r + static_cast<uint32_t>(16 * u), pk);
                } else {
                    const bool is_k = u < 2 * uq;
                    const int cu = u - (is_k ? uq : 2 * uq);           // unit inside the operand: columns 16 cu .. 16 cu + 15
                    const uint32_t rowa = (is_k ? k_row : v_row) + static_cast<uint32_t>((cu >> 2) * kPanelBytes);
                    const int c0 = (2 * cu) & 7;
                    st2_shared_v4(rowa + static_cast<uint32_t>(((c0 ^ rsw) << 4)), pk[0], pk[1], pk[2], pk[3]);
                    st2_shared_v4(rowa + static_cast<uint32_t>((((c0 + 1) ^ rsw) << 4)), pk[4], pk[5], pk[6], pk[7]);
                }
            }
            tmem_st_wait();
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->qkv_ready[grp]);

            // ---- softmax over my window's 64 keys, 32 per thread
            if (first_of_tile && it + 1 < my_tiles) {
                set_it = it + 1;
                row_setup((it + 1) * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x), tok_n, mbits_n, rstd_n, nrm_n);
            }
            mbar_wait(&bars->s_full[grp], par);
            tc_fence_after_sync();
            {
                uint32_t raw[32];
                tm2_ld32(t_s + kcol0 + static_cast<uint32_t>(32 * half), raw);
                tmem_ld_wait();
                const float* bias = s_bias + h * 232 + qb;
                float mx = -INFINITY;
                if (blk4) {
#pragma unroll
                    for (int k = 0; k < 32; k += 2) {
                        const float2 bv = f22(bias[-(((k >> 2) & 3) * 15 + ((k >> 4) << 2) + (k & 3))],
                                              bias[-((((k + 1) >> 2) & 3) * 15 + (((k + 1) >> 4) << 2) + ((k + 1) & 3))]);
                        const float2 v = __ffma2_rn(f22(__uint_as_float(raw[k]), __uint_as_float(raw[k + 1])), sc2, bv);
                        raw[k] = __float_as_uint(v.x);
                        raw[k + 1] = __float_as_uint(v.y);
                        mx = fmaxf(mx, fmaxf(v.x, v.y));
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 32; k += 2) {
                        const float2 bv = f22(bias[-((k >> 3) * 15 + (k & 7))], bias[-(((k + 1) >> 3) * 15 + ((k + 1) & 7))]);
                        const float2 v = __ffma2_rn(f22(__uint_as_float(raw[k]), __uint_as_float(raw[k + 1])), sc2, bv);
                        raw[k] = __float_as_uint(v.x);
                        raw[k + 1] = __float_as_uint(v.y);
                        mx = fmaxf(mx, fmaxf(v.x, v.y));
                    }
                }
                *my_max = mx;
                named_bar_sync(pair_bar, 64);
                mx = fmaxf(mx, my_max[other]);
                const float2 nmx = f22(-mx, -mx);
                float2 acc = f22(0.f, 0.f);
                uint32_t pk[16];
                if (__all_sync(0xffffffffu, mbits == 0xffffffffu)) {   // no key of this warp's rows is masked (warp-uniform branch)
#pragma unroll
                    for (int k = 0; k < 32; k += 2) {
                        const float2 d = __fadd2_rn(f22(__uint_as_float(raw[k]), __uint_as_float(raw[k + 1])), nmx);
                        const float2 e = f22(ex2_approx(d.x), ex2_approx(d.y));
                        acc = __fadd2_rn(acc, e);
                        pk[k >> 1] = pack_bf16x2(e.x, e.y);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 32; k += 2) {
                        const float2 d = __fadd2_rn(f22(__uint_as_float(raw[k]), __uint_as_float(raw[k + 1])), nmx);
                        float2 e = f22(ex2_approx(d.x), ex2_approx(d.y));
                        // keys of another mask region get -100 in the reference: their probability is exp(-100) ~ 0
                        if (!((mbits >> k) & 1u)) e.x = 0.f;
                        if (!((mbits >> (k + 1)) & 1u)) e.y = 0.f;
                        acc = __fadd2_rn(acc, e);
                        pk[k >> 1] = pack_bf16x2(e.x, e.y);
                    }
                }
                my_max[512] = acc.x + acc.y;
                // P in place: key j of the tile -> packed column j / 2; the same keys' slots of the OTHER window get zeros
                tm2_st16(t_s + ((kcol0 + static_cast<uint32_t>(32 * half)) >> 1), pk);
                tm2_st16_zero(t_s + (((64u - kcol0) + static_cast<uint32_t>(32 * half)) >> 1));
            }
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->p_ready[grp]);

            // ---- O = P v done: normalise, stage the row pieces in my group's k panel (dead: S has completed), copy them out
            mbar_wait(&bars->o_full[grp], par);
            tc_fence_after_sync();
            {
                const float inv = 1.0f / (my_max[512] + my_max[512 + other]);
                const float2 inv2 = f22(inv, inv);
                for (int u = half; u < uq; u += 2) {
                    uint32_t raw[16];
                    tmem_ld16(t_r + static_cast<uint32_t>(16 * u), raw);
                    tmem_ld_wait();
                    uint32_t pk[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float2 v = __fmul2_rn(f22(__uint_as_float(raw[2 * e]), __uint_as_float(raw[2 * e + 1])), inv2);
                        pk[e] = pack_bf16x2(v.x, v.y);
                    }
                    const uint32_t rowa = k_row + static_cast<uint32_t>((u >> 2) * kPanelBytes);
                    const int c0 = (2 * u) & 7;
                    st2_shared_v4(rowa + static_cast<uint32_t>(((c0 ^ rsw) << 4)), pk[0], pk[1], pk[2], pk[3]);
                    st2_shared_v4(rowa + static_cast<uint32_t>((((c0 + 1) ^ rsw) << 4)), pk[4], pk[5], pk[6], pk[7]);
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->o_ready[grp]);           // O has left TMEM: the region takes the group's next head
            // attention rows -> out[token, h * hdp ...]: the quadrant's 32 staged rows are copied out by its two warps of this
            // group, consecutive lanes along a row (coalesced 16-byte pieces)
            named_bar_sync(pair_bar, 64);                              // the quadrant's rows (and tok_g) are complete
            {
                const int cpr = p.hdp >> 3;                            // 16-byte chunks per row
                for (int id = t64; id < 32 * cpr; id += 64) {
                    const int rl = static_cast<int>((static_cast<uint32_t>(id) * cpr_magic) >> 16), ch = id - rl * cpr;
                    const int rq = quad * 32 + rl;
                    uint4 val;
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                                 : "r"(kv_g + static_cast<uint32_t>((ch >> 3) * kPanelBytes + rq * 128 + (((ch & 7) ^ (rq & 7)) << 4))));
                    *reinterpret_cast<uint4*>(p.out + static_cast<long long>(tok_g[mb * 128 + rq]) * p.ldo + h * p.hdp + ch * 8) = val;
                }
            }
            named_bar_sync(pair_bar, 64);                              // copied out: my group's next head may overwrite the k panel
            h += 2;
            while (h >= p.nH) { h -= p.nH; ++it; }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == kWLoaderWarp) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem);
    }
}

int fixed_smem_bytes2(int nH, int hdp) {
    return (2 * nH * 3 * hdp + nH * 232) * 4 + 512 * 4 + 2 * 512 * 4 + static_cast<int>(sizeof(Attn2Barriers)) + 64;
}

int round16_2(int v) { return (v + 15) / 16 * 16; }

}  // namespace

// Static plan of the two-heads-in-flight kernel for one block shape.  Returns 1 if the shape is covered (two TMEM regions and two
// k / v panel sets fit), else 0 (use swin_attn.cu).
int swin_attn2_plan(SwinAttnParams& p, int C, int nH, int hdp) {
    if (C <= 0 || C > 320 || nH < 2 || nH > 8 || hdp < 32 || hdp > 128 || (hdp % 16) || nH * 3 * hdp > 1024) return 0;
    p.C = C; p.nH = nH; p.hdp = hdp;
    p.ks = (C + 63) / 64;
    p.k16 = (C + 15) / 16;
    p.pan = (hdp + 63) / 64;
    p.cp = round16_2(C);
    p.fuse_proj = 0;
    // a TMEM region hosts a head from start to end: q|k|v accumulators (3 hdp) -> q bf16 in place + S / P over the dead k|v columns
    // (hdp + 128) -> O over the dead q columns; one region per group
    p.rsz = 3 * hdp > hdp + 128 ? 3 * hdp : hdp + 128;
    if (2 * p.rsz > 512) return 0;
    p.nreg = 2;
    p.col_o = 0;
    p.col_acc = -1;
    const int n3 = 3 * hdp;
    p.qkv_pieces = n3 <= 256 ? 1 : 2;
    p.qp_rows[0] = p.qkv_pieces == 1 ? n3 : round16_2(n3 / 2);
    p.qp_rows[1] = n3 - p.qp_rows[0];
    p.qp_rows[2] = p.qp_rows[3] = 0;
    p.w_slot_bytes = (p.qp_rows[0] * 128 + 1023) / 1024 * 1024;
    p.n_pp = 0; p.p_slots = 0; p.p_slot_bytes = 0;
    for (int i = 0; i < 4; ++i) p.pp_rows[i] = 0;
    const int avail = kSmemLimit - p.ks * kPanelBytes - 4 * p.pan * kPanelBytes - fixed_smem_bytes2(nH, hdp);
    p.w_slots = avail / p.w_slot_bytes;
    if (p.w_slots > kMaxSlots) p.w_slots = kMaxSlots;
    if (p.w_slots < 2) return 0;
    return 1;
}

int launch_swin_attn2(SwinAttnParams& p, int num_sms, cudaStream_t stream) {
    const int nW = (p.H / 8) * (p.W / 8);
    if ((p.H % 8) || (p.W % 8) || ((p.B * nW) & 1) || p.shift < 0 || p.shift >= 8) return ADSR_ERR_BAD_SHAPE;
    if ((reinterpret_cast<uintptr_t>(p.x) & 15) || (reinterpret_cast<uintptr_t>(p.out) & 15) || (p.ldx % 8) || (p.ldo % 8) ||
        (reinterpret_cast<uintptr_t>(p.w1p) & 15))
        return ADSR_ERR_BAD_ALIGN;
    if (p.ldx < (p.C + 7) / 8 * 8 || p.ldo < p.nH * p.hdp) return ADSR_ERR_BAD_SHAPE;
    p.n_tiles = p.B * nW / 2;
    if (p.shift != 0 && p.shift != 4) return ADSR_ERR_BAD_SHAPE;      // square token boxes of side 8 / 4 cover these two
    p.box_r = p.shift == 0 ? 8 : 4;
    const int st = encode_tmap_nhwc_box_bf16(&p.tmap_x, p.x, p.B, p.H, p.W, p.C, p.ldx, p.box_r);
    if (st != ADSR_OK) return st;
    const int smem_bytes = p.ks * kPanelBytes + 4 * p.pan * kPanelBytes + p.w_slots * p.w_slot_bytes + fixed_smem_bytes2(p.nH, p.hdp);
    if (smem_bytes > kSmemLimit) return ADSR_ERR_BAD_SHAPE;
    const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
    if (ensure_dynamic_smem(swin_attn2_kernel, smem_bytes) != cudaSuccess) return ADSR_ERR_CUDA;
    return launch_pdl(swin_attn2_kernel, dim3(grid), dim3(kThreads), static_cast<size_t>(smem_bytes), stream, p) == cudaSuccess ? ADSR_OK
                                                                                                                                 : ADSR_ERR_LAUNCH;
}

}  // namespace adsr
