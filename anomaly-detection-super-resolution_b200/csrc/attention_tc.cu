// tcgen05 / TMEM (shifted-)window attention for 8x8 windows (N = 64 tokens), the DRCT-L shape.
//
// A persistent CTA owns one TILE = two consecutive windows = 128 query rows at a time and walks over the heads:
//   * 4 producer warps gather the q | k | v rows of the tile through the closed-form shifted-window index map
//     (src/drct.py:483, 193-204) with 16-byte cp.async into K-major 128-byte-swizzled shared-memory panels
//     (thread = row); heads of 32 padded channels are loaded in pairs so that a panel row is a full 128 bytes;
//   * one MMA warp issues  S = Q K^T  (tcgen05.mma SS, M = 128, N = 128, fp32 in TMEM; the two off-diagonal 64 x 64
//     blocks are the other window's keys and are never used) and  O = P V  (tcgen05.mma TS: P is read from TMEM, V is
//     the MN-major B operand straight from the gathered rows -- no transpose anywhere);
//   * 8 softmax warps (two groups that alternate heads, thread = query row) read their window's 64 scores from TMEM,
//     add scale, relative-position bias (table in shared memory, index (yq-yk+7)*15 + xq-xk+7, src/drct.py:284-287) and
//     the -100 shift mask from region ids (src/drct.py:449-470), take the softmax in the exp2 domain and write the
//     UNNORMALISED probabilities back in place as bf16 (zeros for the other window's keys); after P V they scale the
//     rows by 1 / sum, stage the bf16 tile in the dead Q panel and store it with coalesced 16-byte rows at the query's
//     ORIGINAL token row (window_reverse + un-shift are the inverse permutation, src/drct.py:500-505).
// S and O are double buffered in TMEM (2 x 128 + 2 x 128 columns), the q|k|v panels 2- or 3-deep in shared memory.
#include "adsr_kernels.h"
#include "ptx.cuh"

namespace adsr {

namespace {

constexpr int kSoftmaxWarps = 16;             // warps 0..15: group = warp >> 3 handles heads with (unit & 1) == group; inside a
                                              // group, half = (warp >> 2) & 1 takes 32 of the row's 64 keys, quad = warp & 3
constexpr int kProducerWarp0 = 16;            // warps 16..19: row gather
constexpr int kMmaWarp = 20;
constexpr int kAllocWarp = 21;
constexpr int kThreads = 22 * 32;
constexpr int kPanelBytes = 128 * 128;
constexpr int kMaxBuf = 3;
constexpr int kSmemLimit = 232448;

struct TcAttnParams {
    const __nv_bfloat16* qkv;
    long long ldq;
    __nv_bfloat16* out;
    long long ldo;
    const float* table;   // [225, nH]
    int B, H, W, shift, nH, hdp;
    int hpl;              // heads per load group (2 for hdp == 32, else 1)
    int pan;              // 64-column panels per operand (1 or 2)
    int nbuf;             // shared-memory load-group buffers
    int n_tiles;          // window pairs
    float scale_log2e;
};

struct __align__(8) AttnBarriers {
    uint64_t full[kMaxBuf];
    uint64_t empty[kMaxBuf];
    uint64_t s_full[2];
    uint64_t p_ready[2];
    uint64_t o_full[2];
    uint64_t o_free[2];
    uint32_t tmem_base;
};

// MN-major B operand, 128-byte swizzle: 64 MN elements (128 B) contiguous per K row, 8 K rows per 1024-byte group,
// further 64-element MN blocks `lbo_bytes` apart.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc_bf16_m128(uint32_t n, uint32_t b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
        "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z)
                 : "memory");
}
__device__ __forceinline__ void tmem_st32_zero(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32x(uint32_t taddr, uint32_t* v) {
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

__device__ __forceinline__ int region_1d(int t, int L, int shift) { return (t >= L - 8 ? 1 : 0) + (t >= L - shift ? 1 : 0); }

__global__ void __launch_bounds__(kThreads, 1) window_attn_tc_kernel(const __grid_constant__ TcAttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int op_bytes = p.pan * kPanelBytes;                          // one operand (q, k or v) of one load group
    const int buf_bytes = 3 * op_bytes;
    uint8_t* bufs = smem;                                              // nbuf x [q | k | v]
    int* s_tok = reinterpret_cast<int*>(smem + p.nbuf * buf_bytes);    // [nbuf][128] token row of each tile row
    uint32_t* s_msk = reinterpret_cast<uint32_t*>(s_tok + kMaxBuf * 128);   // [nbuf][128][2] "same mask region" bits of the row's 64 keys
    float* s_max = reinterpret_cast<float*>(s_msk + kMaxBuf * 256);    // [2 groups][2 halves][128] partial row maxima
    float* s_sum = s_max + 512;                                        // [2 parities][2 groups][2 halves][128] partial row sums
    float* s_bias = s_sum + 1024;                                      // [nH][232] table * log2(e)
    AttnBarriers* bars = reinterpret_cast<AttnBarriers*>(s_bias + p.nH * 232);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int my_tiles = static_cast<int>(blockIdx.x) < p.n_tiles ? (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int groups_per_tile = p.nH / p.hpl;
    const int n_groups = my_tiles * groups_per_tile;                   // load groups of this CTA
    const int n_units = my_tiles * p.nH;                               // (tile, head) units of this CTA

    if ((smem_u32(smem) & 1023u) != 0) __trap();
    for (int i = threadIdx.x; i < 225 * p.nH; i += kThreads) {
        const int h = i % p.nH, e = i / p.nH;
        s_bias[h * 232 + e] = __ldg(p.table + i) * 1.4426950408889634f;
    }
    if (warp == kMmaWarp && lane == 0) {
        for (int b = 0; b < kMaxBuf; ++b) {
            mbar_init(&bars->full[b], 128);
            mbar_init(&bars->empty[b], static_cast<uint32_t>(p.hpl * 9));   // per head: the P V commit + 8 epilogue warps
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->s_full[b], 1);
            mbar_init(&bars->p_ready[b], 8);
            mbar_init(&bars->o_full[b], 1);
            mbar_init(&bars->o_free[b], 8);
        }
        fence_barrier_init();
    }
    if (warp == kAllocWarp) tmem_alloc<512>(&bars->tmem_base);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = bars->tmem_base;
    const int nwx = p.W >> 3, nW = (p.H >> 3) * nwx;

    if (warp >= kProducerWarp0 && warp < kProducerWarp0 + 4) {
        // ============================================================ producers: thread = tile row
        const int r = (warp - kProducerWarp0) * 32 + lane;
        const int chunks = p.hpl * p.hdp >> 3;                         // 16-byte chunks per operand row of one load group
        const int cpl = chunks <= 8 ? 8 : 16;                          // lanes per row (power of two >= chunks)
        const int rows_per_iter = 32 / cpl;
        const int wrow0 = (warp - kProducerWarp0) * 32;                // this warp gathers tile rows wrow0 .. wrow0 + 31
        int pending = -1;                                              // load group whose copies are committed but not yet signalled
        uint32_t mk0 = 0xffffffffu, mk1 = 0xffffffffu;                 // mask bits of the current tile (computed with its first load group)
        for (int g = 0; g < n_groups; ++g) {
            const int it = g / groups_per_tile, lg = g - it * groups_per_tile;
            const int tile = it * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x);
            const int buf = g % p.nbuf;
            // slot -> token (closed form of roll + window_partition) and its mask region
            const int win = tile * 2 + (r >> 6);
            const int b = win / nW, w = win - b * nW;
            const int n = r & 63;
            const int ys = (w / nwx) * 8 + (n >> 3), xs = (w % nwx) * 8 + (n & 7);
            int y = ys + p.shift; if (y >= p.H) y -= p.H;
            int x = xs + p.shift; if (x >= p.W) x -= p.W;
            const int tok = (b * p.H + y) * p.W + x;
            // bit k of my 64-bit word: key k of my window lies in the same shift-mask region as I do (src/drct.py:449-470)
            uint32_t m0 = 0xffffffffu, m1 = 0xffffffffu;
            if (p.shift > 0 && lg == 0) {
                const int wy = (w / nwx) * 8, wx = (w % nwx) * 8;
                const int my = 3 * region_1d(ys, p.H, p.shift) + region_1d(xs, p.W, p.shift);
                m0 = m1 = 0u;
                for (int k = 0; k < 64; ++k) {
                    const int rk = 3 * region_1d(wy + (k >> 3), p.H, p.shift) + region_1d(wx + (k & 7), p.W, p.shift);
                    if (rk == my) { if (k < 32) m0 |= 1u << k; else m1 |= 1u << (k - 32); }
                }
            }
            const uint32_t e_par = (static_cast<uint32_t>(g / p.nbuf) & 1) ^ 1;
            if (pending >= 0 && !mbar_try_wait(&bars->empty[buf], e_par)) {
                // the buffer is still in use: hand over the copies in flight first -- the consumers of THIS buffer may be
                // queued behind the load group that is only signalled after the next commit (2-deep rings would deadlock)
                cp_async_wait<0>();
                fence_proxy_async_smem();
                mbar_arrive(&bars->full[pending % p.nbuf]);
                pending = -1;
            }
            mbar_wait(&bars->empty[buf], e_par);
            s_tok[buf * 128 + r] = tok;
            if (lg == 0) { mk0 = m0; mk1 = m1; }
            s_msk[(buf * 128 + r) * 2] = mk0;
            s_msk[(buf * 128 + r) * 2 + 1] = mk1;
            __syncwarp();                                              // my warp's 32 token rows are in s_tok
            // coalesced gather: cpl lanes walk along one row's 16-byte chunks, 32 / cpl rows per instruction
            {
                const int sub = lane / cpl, c = lane - sub * cpl;
                const long long col0 = static_cast<long long>(lg) * p.hpl * p.hdp + c * 8;
                const uint32_t buf_base = smem_u32(bufs + buf * buf_bytes);
                if (c < chunks) {
                    for (int i = 0; i < cpl; ++i) {
                        const int row = wrow0 + i * rows_per_iter + sub;
                        const __nv_bfloat16* src = p.qkv + static_cast<long long>(s_tok[buf * 128 + row]) * p.ldq + col0;
                        const uint32_t d = buf_base + static_cast<uint32_t>(row * 128) + static_cast<uint32_t>((c >> 3) * kPanelBytes) +
                                           static_cast<uint32_t>((((c & 7) ^ (row & 7)) << 4));
                        cp_async16(d, src);                                                        // q
                        cp_async16(d + static_cast<uint32_t>(op_bytes), src + p.nH * p.hdp);       // k
                        cp_async16(d + 2u * static_cast<uint32_t>(op_bytes), src + 2 * p.nH * p.hdp);   // v
                    }
                }
            }
            cp_async_commit();
            if (pending >= 0) {
                cp_async_wait<1>();
                fence_proxy_async_smem();
                mbar_arrive(&bars->full[pending % p.nbuf]);
            }
            pending = g;
        }
        if (pending >= 0) {
            cp_async_wait<0>();
            fence_proxy_async_smem();
            mbar_arrive(&bars->full[pending % p.nbuf]);
        }
    } else if (warp == kMmaWarp) {
        // ============================================================ MMA issuer (converged warp, one elected lane issues)
        const uint32_t idesc_s = idesc_bf16_m128(128, 0);
        const uint32_t n_pv = static_cast<uint32_t>(p.hpl * p.hdp);     // pairs: both heads' V columns in one N = 64 MMA
        const uint32_t idesc_pv = idesc_bf16_m128(n_pv, 1);
        const int ksteps_s = p.hdp >> 4;
        if (p.hpl == 2) {
            // head pairs (3 load groups per tile): the fixed order S(u), P V(u - 1) measured faster than the opportunistic one
            for (int u = 0; u <= n_units; ++u) {
                if (u < n_units) {
                    // ---- S(u) = Q K^T
                    const int it = u / p.nH, h = u - it * p.nH;
                    const int g = it * groups_per_tile + h / p.hpl;
                    const int buf = g % p.nbuf;
                    const int sb = u & 1;
                    if ((h % p.hpl) == 0) mbar_wait(&bars->full[buf], static_cast<uint32_t>(g / p.nbuf) & 1);
                    tc_fence_after_sync();
                    const uint32_t qa = smem_u32(bufs + buf * buf_bytes);
                    const uint32_t k0 = static_cast<uint32_t>((h % p.hpl) * ksteps_s);   // first K=16 step of this head in the panel
                    if (elect_one_sync()) {
                        for (int k = 0; k < ksteps_s; ++k) {
                            const uint32_t ks = k0 + static_cast<uint32_t>(k);
                            const uint32_t off = (ks >> 2) * kPanelBytes + (ks & 3) * 32;
                            umma_bf16(tmem + static_cast<uint32_t>(sb * 128), umma_desc_k_sw128(qa + off),
                                      umma_desc_k_sw128(qa + op_bytes + off), idesc_s, k == 0 ? 0u : 1u);
                        }
                        umma_commit(&bars->s_full[sb]);
                    }
                    __syncwarp();
                }
                if (u >= 1) {
                    // ---- O(v) = P V
                    const int v = u - 1;
                    const int it = v / p.nH, h = v - it * p.nH;
                    const int g = it * groups_per_tile + h / p.hpl;
                    const int buf = g % p.nbuf;
                    const int sb = v & 1;
                    mbar_wait(&bars->p_ready[sb], static_cast<uint32_t>(v >> 1) & 1);
                    mbar_wait(&bars->o_free[sb], (static_cast<uint32_t>(v >> 1) & 1) ^ 1);
                    tc_fence_after_sync();
                    const uint32_t va = smem_u32(bufs + buf * buf_bytes) + 2u * static_cast<uint32_t>(op_bytes);
                    if (elect_one_sync()) {
    #pragma unroll
                        for (int k = 0; k < 8; ++k)                         // 16 keys per step: 16 rows x 128 B further down the panel
                            umma_bf16_ts(tmem + 256u + static_cast<uint32_t>(sb * 128), tmem + static_cast<uint32_t>(sb * 128 + 8 * k),
                                         umma_desc_mn_sw128(va + static_cast<uint32_t>(k * 2048), kPanelBytes), idesc_pv, k == 0 ? 0u : 1u);
                        umma_commit(&bars->o_full[sb]);
                        umma_commit(&bars->empty[buf]);
                    }
                    __syncwarp();
                }
            }
        } else {
            // S(u) needs its load group, P V(u) needs the softmax of unit u: whichever is ready first is issued first, so that a
            // late load never holds back a P V whose probabilities are waiting in TMEM (and the other way round).  The S buffer of
            // unit u is the one of unit u - 2: S(u) may only be issued once P V(u - 2) has been (the tensor pipe runs in order).
            auto ready = [&](uint64_t* bar, uint32_t parity) -> bool {     // one lane polls, the warp stays converged
                uint32_t ok = 0;
                if (lane == 0) ok = mbar_test_wait(bar, parity) ? 1u : 0u;
                return __shfl_sync(0xffffffffu, ok, 0) != 0;
            };
            int us = 0, up = 0;                                            // next S unit / next P V unit to issue
            while (up < n_units) {
                if (us < n_units && us < up + 2) {
                    const int it = us / p.nH, h = us - it * p.nH;
                    const int g = it * groups_per_tile + h / p.hpl;
                    const int buf = g % p.nbuf;
                    if ((h % p.hpl) != 0 || ready(&bars->full[buf], static_cast<uint32_t>(g / p.nbuf) & 1)) {
                        // ---- S(us) = Q K^T
                        const int sb = us & 1;
                        tc_fence_after_sync();
                        const uint32_t qa = smem_u32(bufs + buf * buf_bytes);
                        const uint32_t k0 = static_cast<uint32_t>((h % p.hpl) * ksteps_s);   // first K=16 step of this head in the panel
                        if (elect_one_sync()) {
                            for (int k = 0; k < ksteps_s; ++k) {
                                const uint32_t ks = k0 + static_cast<uint32_t>(k);
                                const uint32_t off = (ks >> 2) * kPanelBytes + (ks & 3) * 32;
                                umma_bf16(tmem + static_cast<uint32_t>(sb * 128), umma_desc_k_sw128(qa + off),
                                          umma_desc_k_sw128(qa + op_bytes + off), idesc_s, k == 0 ? 0u : 1u);
                            }
                            umma_commit(&bars->s_full[sb]);
                        }
                        __syncwarp();
                        ++us;
                    }
                }
                if (up < us) {
                    const int v = up;
                    const int sb = v & 1;
                    if (ready(&bars->p_ready[sb], static_cast<uint32_t>(v >> 1) & 1) &&
                        ready(&bars->o_free[sb], (static_cast<uint32_t>(v >> 1) & 1) ^ 1)) {
                        // ---- O(v) = P V
                        const int it = v / p.nH, h = v - it * p.nH;
                        const int g = it * groups_per_tile + h / p.hpl;
                        const int buf = g % p.nbuf;
                        tc_fence_after_sync();
                        const uint32_t va = smem_u32(bufs + buf * buf_bytes) + 2u * static_cast<uint32_t>(op_bytes);
                        if (elect_one_sync()) {
    #pragma unroll
                            for (int k = 0; k < 8; ++k)                     // 16 keys per step: 16 rows x 128 B further down the panel
                                umma_bf16_ts(tmem + 256u + static_cast<uint32_t>(sb * 128), tmem + static_cast<uint32_t>(sb * 128 + 8 * k),
                                             umma_desc_mn_sw128(va + static_cast<uint32_t>(k * 2048), kPanelBytes), idesc_pv, k == 0 ? 0u : 1u);
                            umma_commit(&bars->o_full[sb]);
                            umma_commit(&bars->empty[buf]);
                        }
                        __syncwarp();
                        ++up;
                    }
                }
            }
        }
    } else if (warp < kSoftmaxWarps) {
        // ============================================================ softmax + output: two threads per query row
        const int grp = warp >> 3, half = (warp >> 2) & 1, quad = warp & 3;
        const int r = quad * 32 + lane;                                // tile row; window = r >> 6, slot = r & 63
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const int n = r & 63;
        const int qb = ((n >> 3) + 7) * 15 + (n & 7) + 7 - half * 60;  // my keys are 32 half .. 32 half + 31: (k >> 3) * 15 starts at 60 half
        const uint32_t kcol0 = static_cast<uint32_t>((r >> 6) * 64);   // my window's keys inside the 128 S columns
        const int bar_id = 1 + grp * 4 + quad;                         // the two warps that share my rows
        float* my_max = s_max + (grp * 2 + half) * 128 + r;
        const float* other_max = s_max + (grp * 2 + (half ^ 1)) * 128 + r;
        int it = 0, h = grp;                                           // unit u = it * nH + h, u = grp, grp + 2, ...
        while (h >= p.nH) { h -= p.nH; ++it; }
        for (int u = grp; u < n_units; u += 2) {
            const int lgi = h / p.hpl;
            const int g = it * groups_per_tile + lgi;
            const int buf = g % p.nbuf;
            const int hin = h - lgi * p.hpl;                           // head inside its load group
            const int sb = grp;                                        // == u & 1
            const uint32_t use = static_cast<uint32_t>(u >> 1) & 1;
            const float* bias = s_bias + h * 232 + qb;
            float* my_sum = s_sum + ((use * 2 + grp) * 2 + half) * 128 + r;
            const float* other_sum = s_sum + ((use * 2 + grp) * 2 + (half ^ 1)) * 128 + r;
            mbar_wait(&bars->s_full[sb], use);
            tc_fence_after_sync();
            const uint32_t mbits = s_msk[(buf * 128 + r) * 2 + half];
            const uint32_t s_addr = tmem + static_cast<uint32_t>(sb * 128) + lane_off;
            uint32_t raw[32];
            tmem_ld32x(s_addr + kcol0 + static_cast<uint32_t>(32 * half), raw);
            tmem_ld_wait();
            float t[32];
            float mx = -INFINITY;
            const float2 sc2 = make_float2(p.scale_log2e, p.scale_log2e);
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
                const float2 bv = make_float2(bias[-((k >> 3) * 15 + (k & 7))], bias[-(((k + 1) >> 3) * 15 + ((k + 1) & 7))]);
                const float2 v = __ffma2_rn(make_float2(__uint_as_float(raw[k]), __uint_as_float(raw[k + 1])), sc2, bv);
                t[k] = v.x;
                t[k + 1] = v.y;
                mx = fmaxf(mx, fmaxf(v.x, v.y));
            }
            *my_max = mx;
            named_bar_sync(bar_id, 64);
            mx = fmaxf(mx, *other_max);
            const float2 nmx = make_float2(-mx, -mx);
            float2 acc = make_float2(0.f, 0.f);
            uint32_t pk[16];
            if (mbits == 0xffffffffu) {
#pragma unroll
                for (int k = 0; k < 32; k += 2) {
                    const float2 d = __fadd2_rn(make_float2(t[k], t[k + 1]), nmx);
                    const float2 e = make_float2(ex2_approx(d.x), ex2_approx(d.y));
                    acc = __fadd2_rn(acc, e);
                    pk[k >> 1] = pack_bf16x2(e.x, e.y);
                }
            } else {
                // keys of another mask region get -100 in the reference: their probability is exp(-100) ~ 0
#pragma unroll
                for (int k = 0; k < 32; k += 2) {
                    const float2 d = __fadd2_rn(make_float2(t[k], t[k + 1]), nmx);
                    float2 e = make_float2(ex2_approx(d.x), ex2_approx(d.y));
                    if (!((mbits >> k) & 1u)) e.x = 0.f;
                    if (!((mbits >> (k + 1)) & 1u)) e.y = 0.f;
                    acc = __fadd2_rn(acc, e);
                    pk[k >> 1] = pack_bf16x2(e.x, e.y);
                }
            }
            *my_sum = acc.x + acc.y;
            // P in place: key j of the tile -> packed column j / 2; my 16 columns get the values, the same keys' slots of the
            // OTHER window (which this row must not see) get zeros
            tmem_st16(s_addr + ((kcol0 + static_cast<uint32_t>(32 * half)) >> 1), pk);
            tmem_st16_zero(s_addr + (((64u - kcol0) + static_cast<uint32_t>(32 * half)) >> 1));
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->p_ready[sb]);

            // ---- O = P V done: normalise my 16-column units, stage them in the dead Q panel, store them as 32-byte row pieces
            mbar_wait(&bars->o_full[sb], use);
            tc_fence_after_sync();
            const float inv = 1.0f / (*my_sum + *other_sum);
            const int col0 = hin * p.hdp;                              // column of this head inside the load group's panel row
            const uint32_t o_addr = tmem + 256u + static_cast<uint32_t>(sb * 128) + lane_off + static_cast<uint32_t>(col0);
            const uint32_t stage_base = smem_u32(bufs + buf * buf_bytes);
            const uint32_t stage_row = stage_base + static_cast<uint32_t>(r * 128);
            const int units = p.hdp >> 4;
            for (int un = half; un < units; un += 2) {
                uint32_t ro[16];
                tmem_ld16(o_addr + static_cast<uint32_t>(16 * un), ro);
                tmem_ld_wait();
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    const int cc = (col0 + 16 * un + 8 * o) >> 3;      // 16-byte chunk index along the row
                    const uint32_t d = stage_row + static_cast<uint32_t>((cc >> 3) * kPanelBytes) + static_cast<uint32_t>((((cc & 7) ^ (r & 7)) << 4));
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(d),
                                 "r"(pack_bf16x2(__uint_as_float(ro[8 * o + 0]) * inv, __uint_as_float(ro[8 * o + 1]) * inv)),
                                 "r"(pack_bf16x2(__uint_as_float(ro[8 * o + 2]) * inv, __uint_as_float(ro[8 * o + 3]) * inv)),
                                 "r"(pack_bf16x2(__uint_as_float(ro[8 * o + 4]) * inv, __uint_as_float(ro[8 * o + 5]) * inv)),
                                 "r"(pack_bf16x2(__uint_as_float(ro[8 * o + 6]) * inv, __uint_as_float(ro[8 * o + 7]) * inv))
                                 : "memory");
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->o_free[sb]);
            // write-out of what this warp staged: per unit 32 rows x 32 bytes, two lanes per row
            {
                const int rr = quad * 32 + (lane >> 1), o = lane & 1;
                for (int pass = 0; pass < 2; ++pass) {
                    const int row = rr + pass * 16;
                    const int tok = s_tok[buf * 128 + row];
                    __nv_bfloat16* dst = p.out + static_cast<long long>(tok) * p.ldo + h * p.hdp;
                    for (int un = half; un < units; un += 2) {
                        const int cc = (col0 + 16 * un + 8 * o) >> 3;
                        const uint32_t sa = stage_base + static_cast<uint32_t>(row * 128) + static_cast<uint32_t>((cc >> 3) * kPanelBytes) +
                                            static_cast<uint32_t>((((cc & 7) ^ (row & 7)) << 4));
                        uint4 val;
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w) : "r"(sa));
                        *reinterpret_cast<uint4*>(dst + 16 * un + 8 * o) = val;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->empty[buf]);
            h += 2;
            while (h >= p.nH) { h -= p.nH; ++it; }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == kAllocWarp) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem);
    }
}

}  // namespace

// 8 x 8 windows, an even number of windows, hdp in {32, 48, 64, 80, ..., 128}; returns ADSR_ERR_BAD_SHAPE otherwise so that
// the caller can fall back to the mma.sync kernel for other window sizes.
int launch_window_attention_tc(const void* qkv, long long ldq, void* out, long long ldo, const float* table, int B, int H, int W,
                               int shift, int nH, int hd, int hdp, int num_sms, cudaStream_t stream) {
    const int nW = (H / 8) * (W / 8);
    if ((H % 8) || (W % 8) || ((B * nW) & 1) || hdp < 32 || hdp > 128 || (hdp % 16) || nH < 1 || nH > 16) return ADSR_ERR_BAD_SHAPE;
    TcAttnParams p{};
    p.qkv = static_cast<const __nv_bfloat16*>(qkv);
    p.ldq = ldq;
    p.out = static_cast<__nv_bfloat16*>(out);
    p.ldo = ldo;
    p.table = table;
    p.B = B; p.H = H; p.W = W; p.shift = shift; p.nH = nH; p.hdp = hdp;
    p.hpl = (hdp == 32 && (nH % 2) == 0) ? 2 : 1;
    p.pan = (p.hpl * hdp + 63) / 64;
    p.n_tiles = B * nW / 2;
    p.scale_log2e = (1.0f / sqrtf(static_cast<float>(hd))) * 1.4426950408889634f;
    const int fixed = 3 * kMaxBuf * 128 * 4 + 1536 * 4 + nH * 232 * 4 + static_cast<int>(sizeof(AttnBarriers)) + 64;
    p.nbuf = (kSmemLimit - fixed) / (3 * p.pan * kPanelBytes);
    if (p.nbuf > kMaxBuf) p.nbuf = kMaxBuf;
    if (p.nbuf < 2) return ADSR_ERR_BAD_SHAPE;
    const int smem_bytes = p.nbuf * 3 * p.pan * kPanelBytes + fixed;
    const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
    if (ensure_dynamic_smem(window_attn_tc_kernel, smem_bytes) != cudaSuccess) return ADSR_ERR_CUDA;
    window_attn_tc_kernel<<<grid, kThreads, smem_bytes, stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? ADSR_OK : ADSR_ERR_LAUNCH;
}

}  // namespace adsr
