// tcgen05 / TMEM (shifted-)window attention for 16 x 16 windows (N = 256 tokens): DRCT-L at 64 px LR, BASELINE configs[3]
// (src/drct.py:271-302 with window_size = img_size // 4 = 16, src/main.py:286; shift 8, mask of src/drct.py:449-470).
//
// A persistent CTA is bound to ONE head (CTA c -> head c % nH, so its relative-position table stays resident) and walks
// over windows.  A unit = (window, head): 256 queries = two M = 128 tiles ("streams") against the window's 256 keys.
//   * 4 producer warps gather the q | k | v rows of the window through the closed-form shifted-window index map
//     (src/drct.py:483, 193-204) with 16-byte cp.async into K-major 128-byte-swizzled panels (K: 256 rows, Q: 128 rows per
//     stream, V: 256 key rows read as the MN-major B operand of P V); K / V / Q rings are released by tcgen05.commit as
//     soon as the MMAs that read them have completed;
//   * one MMA warp issues, whichever is ready first,  S_g = Q_g K^T  (SS, M = 128, N = 256: the stream's whole 256-column
//     TMEM half) and  O_g = P_g V  (TS: P from TMEM, 16 K-steps, N = hdp);
//   * 8 softmax warps per stream (thread = query row x key half): pass A reads S, applies scale + relative-position bias
//     (four shifted copies of the x-reversed table in shared memory so that the 16 biases of a key row are four aligned,
//     conflict-free 128-bit loads) + the -100 shift mask (as -inf), writes the logits back in place and takes the row max;
//     pass B re-reads them, exponentiates (exp2 domain), and writes the UNNORMALISED bf16 probabilities in place -- the
//     lower key half packs upwards into columns [0, 64), the upper half is processed downwards and packs into [192, 256),
//     so neither half ever overwrites a logit that is still unread and the 128 columns in between are free for O;
//   * the same warps normalise O (registers), release the TMEM half, stage 32-column slices in a small swizzled panel and
//     copy them out as coalesced 64-byte row pieces to the query's ORIGINAL token row (window_reverse + un-shift).
// The two streams drift half a period apart, so one stream's MMAs run under the other's softmax.
#include "adsr_kernels.h"
#include "ptx.cuh"

namespace adsr {

namespace {

constexpr int kSoftmaxWarps = 16;             // warps 0..15: stream = warp >> 3, key half = (warp >> 2) & 1, quadrant = warp & 3
constexpr int kProducerWarp0 = 16;            // warps 16, 17: row gather (warp g gathers the 128 window slots of stream g)
constexpr int kProducerWarps = 2;
constexpr int kMmaWarp = 18;
constexpr int kAllocWarp = 19;
constexpr int kThreads = 20 * 32;             // 20 warps: 96 registers per thread (a 21st warp would cost 16 of them)
constexpr int kSmemLimit = 232448;
constexpr int kTabPitch = 28;                 // floats per table row of a shifted copy
constexpr int kTabCopy = 872;                 // floats per shifted copy (31 rows x 28, padded so that copy s starts 2 s bank groups on)
constexpr int kStageBytes = 128 * 64;         // output staging per stream: 128 rows x 32 bf16

struct Tc16Params {
    const __nv_bfloat16* qkv;
    long long ldq;
    __nv_bfloat16* out;
    long long ldo;
    const float* table;   // [961, nH]
    int B, H, W, shift, nH, hdp;
    int pan;              // 64-column panels per operand (1 or 2)
    int nbuf;             // ring depth of the K / V / Q buffers (2 when pan == 1, else 1)
    int n_win;            // windows in the batch
    int n_slots;          // CTAs per head
    float scale_log2e;
};

struct __align__(8) Bars16 {
    uint64_t k_full[2], k_empty[2], v_full[2], v_empty[2];
    uint64_t q_full[2][2], q_empty[2][2];     // [stream][stage]
    uint64_t s_full[2], p_ready[2], o_full[2], o_free[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t desc_mn_sw128_16(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc16_m128(uint32_t n, uint32_t b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tm_ld32(uint32_t taddr, uint32_t* v) {
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
__device__ __forceinline__ void tm_st8x(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tm_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void st_sh_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// region id along one axis of the shifted frame (src/drct.py:455-462: slices [0, -ws), [-ws, -shift), [-shift, L))
__device__ __forceinline__ int region16(int t, int L, int shift) { return (t >= L - 16 ? 1 : 0) + (t >= L - shift ? 1 : 0); }

__global__ void __launch_bounds__(kThreads, 1) window_attn16_tc_kernel(const __grid_constant__ Tc16Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int kv_bytes = p.pan * 256 * 128;                            // one K (or V) stage: 256 rows x 128 B per panel
    const int q_bytes = p.pan * 128 * 128;                             // one Q stage of one stream
    uint8_t* k_ring = smem;                                            // [nbuf]
    uint8_t* v_ring = k_ring + p.nbuf * kv_bytes;                      // [nbuf]
    uint8_t* q_ring = v_ring + p.nbuf * kv_bytes;                      // [stream][nbuf]
    uint8_t* stage = q_ring + 2 * p.nbuf * q_bytes;                    // [stream] 128 rows x 64 B
    float* s_tab = reinterpret_cast<float*>(stage + 2 * kStageBytes);  // [4 copies][kTabCopy]
    float* s_max = s_tab + 4 * kTabCopy;                               // [stream][half][128]
    float* s_bmax = reinterpret_cast<float*>(smem + p.nbuf * (2 * kv_bytes + 2 * q_bytes) + 2 * kStageBytes + 4 * kTabCopy * 4 + 2 * 512 * 4);   // [20 warps]
    float* s_sum = s_max + 512;                                        // [stream][half][128]
    Bars16* bars = reinterpret_cast<Bars16*>(s_sum + 512 + 32);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int head = static_cast<int>(blockIdx.x) % p.nH;
    const int slot = static_cast<int>(blockIdx.x) / p.nH;
    const int n_units = slot < p.n_win ? (p.n_win - slot + p.n_slots - 1) / p.n_slots : 0;

    if ((smem_u32(smem) & 1023u) != 0) __trap();
    // copy s of the x-reversed table: C_s[dy][i] = T[dy][30 - (i + s)] * log2(e)   (dy = yq - yk + 15, dx = xq - xk + 15;
    // the 16 keys kx = 0..15 of a key row are C_s[dy][(15 - xq - s) + kx] with s = (15 - xq) % 4: 16-byte aligned)
    float bmax_part = 0.f;                                             // upper bound of the bias (>= 0: the pad entries are zeros)
    for (int i = threadIdx.x; i < 4 * kTabCopy; i += kThreads) {
        const int s = i / kTabCopy, rem = i - s * kTabCopy;
        const int dy = rem / kTabPitch, j = rem - dy * kTabPitch;
        const int e = j + s;
        const float v = (dy < 31 && e <= 30) ? __ldg(p.table + (dy * 31 + (30 - e)) * p.nH + head) * 1.4426950408889634f : 0.f;
        s_tab[i] = v;
        bmax_part = fmaxf(bmax_part, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bmax_part = fmaxf(bmax_part, __shfl_xor_sync(0xffffffffu, bmax_part, o));
    if (lane == 0) s_bmax[warp] = bmax_part;
    if (warp == kMmaWarp && lane == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->k_full[b], kProducerWarps * 32);
            mbar_init(&bars->k_empty[b], 2);                           // the commits behind S_0 and S_1
            mbar_init(&bars->v_full[b], kProducerWarps * 32);
            mbar_init(&bars->v_empty[b], 2);                           // the commits behind O_0 and O_1
            for (int g = 0; g < 2; ++g) {
                mbar_init(&bars->q_full[g][b], kProducerWarps * 32);
                mbar_init(&bars->q_empty[g][b], 1);
            }
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
    const int nwx = p.W >> 4, nW = (p.H >> 4) * nwx;

    if (warp >= kProducerWarp0 && warp < kProducerWarp0 + kProducerWarps) {
        // ============================================================ producers
        const int pw = warp - kProducerWarp0;
        const int chunks = p.hdp >> 3;                                 // 16-byte chunks per operand row
        const int cpl = chunks <= 4 ? 4 : (chunks <= 8 ? 8 : 16);      // lanes per row
        const int rows_per_it = 32 / cpl;
        const int sub = lane / cpl, c = lane - sub * cpl;
        const uint32_t c_off = static_cast<uint32_t>(c >> 3), c_low = static_cast<uint32_t>(c & 7);
        const long long qcol = static_cast<long long>(head) * p.hdp + c * 8;
        const long long kcol = qcol + static_cast<long long>(p.nH) * p.hdp;
        const long long vcol = kcol + static_cast<long long>(p.nH) * p.hdp;
        // pending = copies committed but not yet handed over: 1 = {K, Q_0, Q_1} of stage pst, 2 = V of stage pst
        int pending = 0, pst = 0;
        auto hand_over = [&]() {
            fence_proxy_async_smem();
            if (pending == 1) {
                mbar_arrive(&bars->k_full[pst]);
                mbar_arrive(&bars->q_full[0][pst]);
                mbar_arrive(&bars->q_full[1][pst]);
            } else if (pending == 2) {
                mbar_arrive(&bars->v_full[pst]);
            }
            pending = 0;
        };
        for (int i = 0; i < n_units; ++i) {
            const int win = slot + i * p.n_slots;
            const int b = win / nW, w = win - b * nW;
            const int wy = (w / nwx) * 16, wx = (w % nwx) * 16;
            const int st = i % p.nbuf;
            const uint32_t e_par = (static_cast<uint32_t>(i / p.nbuf) & 1) ^ 1;
            // window slot n -> token row: closed form of roll + window_partition (src/drct.py:483, 193-204)
            auto src_row = [&](int n) -> const __nv_bfloat16* {
                int y = wy + (n >> 4) + p.shift; if (y >= p.H) y -= p.H;
                int x = wx + (n & 15) + p.shift; if (x >= p.W) x -= p.W;
                return p.qkv + static_cast<long long>((b * p.H + y) * p.W + x) * p.ldq;
            };
            // ---- {K, Q_0, Q_1}: this warp gathers window slots 128 pw .. 128 pw + 127 (= the queries of stream pw)
            if (pending && !(mbar_try_wait(&bars->k_empty[st], e_par) && mbar_try_wait(&bars->q_empty[0][st], e_par) &&
                             mbar_try_wait(&bars->q_empty[1][st], e_par))) {
                // the buffers are still in use: hand over the copies in flight first (their consumers may be what frees these)
                cp_wait<0>();
                hand_over();
            }
            mbar_wait(&bars->k_empty[st], e_par);
            mbar_wait(&bars->q_empty[0][st], e_par);
            mbar_wait(&bars->q_empty[1][st], e_par);
            if (c < chunks) {
                const uint32_t kb = smem_u32(k_ring + st * kv_bytes) + c_off * (256u * 128u);
                const uint32_t qb = smem_u32(q_ring + (pw * p.nbuf + st) * q_bytes) + c_off * (128u * 128u);
                for (int it = 0; it < 128; it += rows_per_it) {
                    const int n = pw * 128 + it + sub;
                    const __nv_bfloat16* src = src_row(n);
                    const uint32_t sw = (c_low ^ static_cast<uint32_t>(n & 7)) << 4;
                    cp16(kb + static_cast<uint32_t>(n * 128) + sw, src + kcol);
                    cp16(qb + static_cast<uint32_t>((n & 127) * 128) + sw, src + qcol);
                }
            }
            cp_commit();
            if (pending) {
                cp_wait<1>();
                hand_over();
            }
            pending = 1; pst = st;
            // ---- V
            if (!mbar_try_wait(&bars->v_empty[st], e_par)) {
                cp_wait<0>();
                hand_over();
            }
            mbar_wait(&bars->v_empty[st], e_par);
            if (c < chunks) {
                const uint32_t vb = smem_u32(v_ring + st * kv_bytes) + c_off * (256u * 128u);
                for (int it = 0; it < 128; it += rows_per_it) {
                    const int n = pw * 128 + it + sub;
                    cp16(vb + static_cast<uint32_t>(n * 128) + ((c_low ^ static_cast<uint32_t>(n & 7)) << 4), src_row(n) + vcol);
                }
            }
            cp_commit();
            if (pending) {
                cp_wait<1>();
                hand_over();
            }
            pending = 2; pst = st;
        }
        if (pending) {
            cp_wait<0>();
            hand_over();
        }
    } else if (warp == kMmaWarp) {
        // ============================================================ MMA issuer (converged warp, one elected lane issues)
        const uint32_t idesc_s = idesc16_m128(256, 0);
        const uint32_t idesc_pv = idesc16_m128(static_cast<uint32_t>(p.hdp), 1);
        const int ksteps_s = p.hdp >> 4;
        auto ready = [&](uint64_t* bar, uint32_t parity) -> bool {     // one lane polls, the warp stays converged
            uint32_t ok = 0;
            if (lane == 0) ok = mbar_test_wait(bar, parity) ? 1u : 0u;
            return __shfl_sync(0xffffffffu, ok, 0) != 0;
        };
        int ns0 = 0, ns1 = 0, np0 = 0, np1 = 0;                        // next S / next P V unit of each stream
        auto try_pv = [&](int g, int& np_, int ns_) {
            if (np_ >= ns_) return;
            const int j = np_, st = j % p.nbuf;
            if (!ready(&bars->p_ready[g], static_cast<uint32_t>(j) & 1) || !ready(&bars->v_full[st], static_cast<uint32_t>(j / p.nbuf) & 1)) return;
            tc_fence_after_sync();
            const uint32_t va = smem_u32(v_ring + st * kv_bytes);
            const uint32_t treg = tmem + static_cast<uint32_t>(g * 256);
            if (elect_one_sync()) {
#pragma unroll
                for (int k = 0; k < 16; ++k)                           // 16 keys per step; P of keys >= 128 sits in columns 192..
                    umma_bf16_ts(treg + 64u, treg + static_cast<uint32_t>(k < 8 ? 8 * k : 192 + 8 * (k - 8)),
                                 desc_mn_sw128_16(va + static_cast<uint32_t>(k * 2048), 256u * 128u), idesc_pv, k == 0 ? 0u : 1u);
                umma_commit(&bars->o_full[g]);
                umma_commit(&bars->v_empty[st]);
            }
            __syncwarp();
            ++np_;
        };
        auto try_s = [&](int g, int& ns_, int np_) {
            if (ns_ >= n_units || ns_ != np_) return;                  // the stream's TMEM half hosts one unit at a time
            const int i = ns_, st = i % p.nbuf;
            const uint32_t ph = static_cast<uint32_t>(i / p.nbuf) & 1;
            if (i > 0 && !ready(&bars->o_free[g], static_cast<uint32_t>(i - 1) & 1)) return;
            if (!ready(&bars->k_full[st], ph) || !ready(&bars->q_full[g][st], ph)) return;
            tc_fence_after_sync();
            const uint32_t qa = smem_u32(q_ring + (g * p.nbuf + st) * q_bytes);
            const uint32_t ka = smem_u32(k_ring + st * kv_bytes);
            if (elect_one_sync()) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (k < ksteps_s) {
                        const uint32_t pn = static_cast<uint32_t>(k >> 2), off = static_cast<uint32_t>(k & 3) * 32u;
                        umma_bf16(tmem + static_cast<uint32_t>(g * 256), umma_desc_k_sw128(qa + pn * (128u * 128u) + off),
                                  umma_desc_k_sw128(ka + pn * (256u * 128u) + off), idesc_s, k == 0 ? 0u : 1u);
                    }
                }
                umma_commit(&bars->s_full[g]);
                umma_commit(&bars->q_empty[g][st]);
                umma_commit(&bars->k_empty[st]);
            }
            __syncwarp();
            ++ns_;
        };
        while (np0 < n_units || np1 < n_units) {
            try_pv(0, np0, ns0);
            try_s(0, ns0, np0);
            try_pv(1, np1, ns1);
            try_s(1, ns1, np1);
        }
    } else if (warp < kSoftmaxWarps) {
        // ============================================================ softmax + output
        const int g = warp >> 3, half = (warp >> 2) & 1, quad = warp & 3;
        const int r = quad * 32 + lane;                                // row of the stream's M tile
        const int qs = g * 128 + r;                                    // query slot in the window
        const int yq = qs >> 4, xq = qs & 15;
        const uint32_t treg = tmem + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(g * 256);
        const int e0 = 15 - xq, sc = e0 & 3;
        const uint32_t tab = smem_u32(s_tab + sc * kTabCopy + (e0 - sc));   // + (dy * kTabPitch + 4 m) * 4 bytes
        const int pair_bar = 3 + g * 4 + quad;                         // the two warps (key halves) that share my rows
        // my row's slot in s_max; the other key half's slot is 128 floats away, the s_sum slots 512 floats further on
        float* my_max = s_max + (g * 2 + half) * 128 + r;
        const int other = half ? -128 : 128;
        const float2 sc2 = make_float2(p.scale_log2e, p.scale_log2e);
        const int units = p.hdp >> 4;                                  // 16-column units of O
        const int t256 = (half * 4 + quad) * 32 + lane;                // thread index inside the stream
        const uint32_t stage_g = smem_u32(stage + g * kStageBytes);
        const float ninf = -INFINITY;
        float bmax = 0.f;
        for (int w = 0; w < kThreads / 32; ++w) bmax = fmaxf(bmax, s_bmax[w]);

        for (int i = 0; i < n_units; ++i) {
            const int win = slot + i * p.n_slots;
            const int b = win / nW, w = win - b * nW;
            const int wy = (w / nwx) * 16, wx = (w % nwx) * 16;
            // shift mask (src/drct.py:449-470): key (ky, kx) counts iff its row region AND its column region equal mine
            uint32_t ymask = 0xffffu, xmask = 0xffffu;
            if (p.shift > 0) {
                const int my_ry = region16(wy + yq, p.H, p.shift), my_rx = region16(wx + xq, p.W, p.shift);
                ymask = xmask = 0u;
                for (int j = 0; j < 16; ++j) {
                    ymask |= (region16(wy + j, p.H, p.shift) == my_ry ? 1u : 0u) << j;
                    xmask |= (region16(wx + j, p.W, p.shift) == my_rx ? 1u : 0u) << j;
                }
            }
            const bool masked = (ymask & xmask) != 0xffffu;
            const uint32_t par = static_cast<uint32_t>(i) & 1;

            // ---- pass A: an UPPER BOUND of the row's largest logit, scale * max_k S + max(bias): softmax is shift-invariant, so any
            // bound within a few units of the true maximum serves (the bias spans a few units; the probabilities stay far from
            // underflow) -- no bias look-ups, no write-back
            mbar_wait(&bars->s_full[g], par);
            tc_fence_after_sync();
            float mx = ninf;
            {
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t ra[32];
                    tm_ld32(treg + static_cast<uint32_t>(128 * half + 32 * c), ra);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 32; k += 4)
                        mx = fmaxf(fmaxf(mx, fmaxf(__uint_as_float(ra[k]), __uint_as_float(ra[k + 1]))),
                                   fmaxf(__uint_as_float(ra[k + 2]), __uint_as_float(ra[k + 3])));
                }
            }
            *my_max = mx;
            named_bar_sync(pair_bar, 64);
            mx = fmaf(fmaxf(mx, my_max[other]), p.scale_log2e, bmax);

            // ---- pass B: logits = S * scale + bias, unnormalised probabilities in place (lower key half ascending -> columns
            // [0, 64), upper half descending -> [192, 256)); masked keys (-100 in the reference) get probability 0
            {
                const float2 nmx = make_float2(-mx, -mx);
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    const int c = half ? 3 - cc : cc;
                    uint32_t raw[32];
                    tm_ld32(treg + static_cast<uint32_t>(128 * half + 32 * c), raw);
#pragma unroll
                    for (int kr = 0; kr < 2; ++kr) {
                        uint32_t pk[8];
                        const int ky = 8 * half + 2 * c + kr;
                        const uint32_t bp = tab + static_cast<uint32_t>((yq - ky + 15) * kTabPitch * 4);
                        float4 bv[4];
#pragma unroll
                        for (int m = 0; m < 4; ++m)
                            asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bv[m].x), "=f"(bv[m].y), "=f"(bv[m].z), "=f"(bv[m].w) : "r"(bp + 16u * m));
                        if (kr == 0) tmem_ld_wait();                   // the first key row's biases are already on their way
                        const uint32_t ybit = (ymask >> ky) & 1u;
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            const int o = 16 * kr + 4 * m;
                            const float2 d0 = __fadd2_rn(__ffma2_rn(make_float2(__uint_as_float(raw[o]), __uint_as_float(raw[o + 1])), sc2,
                                                                    make_float2(bv[m].x, bv[m].y)), nmx);
                            const float2 d1 = __fadd2_rn(__ffma2_rn(make_float2(__uint_as_float(raw[o + 2]), __uint_as_float(raw[o + 3])), sc2,
                                                                    make_float2(bv[m].z, bv[m].w)), nmx);
                            float2 e0 = make_float2(ex2_approx(d0.x), ex2_approx(d0.y));
                            float2 e1 = make_float2(ex2_approx(d1.x), ex2_approx(d1.y));
                            if (masked) {
                                const uint32_t xm = ybit ? (xmask >> (4 * m)) : 0u;
                                if (!(xm & 1u)) e0.x = 0.f;
                                if (!(xm & 2u)) e0.y = 0.f;
                                if (!(xm & 4u)) e1.x = 0.f;
                                if (!(xm & 8u)) e1.y = 0.f;
                            }
                            acc = __fadd2_rn(acc, __fadd2_rn(e0, e1));
                            pk[2 * m] = pack_bf16x2(e0.x, e0.y);
                            pk[2 * m + 1] = pack_bf16x2(e1.x, e1.y);
                        }
                        // the chunk's 32 logits are in registers: its packed columns may be overwritten (8 per key row)
                        tm_st8x(treg + static_cast<uint32_t>((half ? 192 : 0) + 16 * c + 8 * kr), pk);
                    }
                }
                my_max[512] = acc.x + acc.y;
            }
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->p_ready[g]);

            // ---- O = P V: normalise into registers, release the TMEM half, then stage + copy out 32 columns at a time
            mbar_wait(&bars->o_full[g], par);
            tc_fence_after_sync();
            const float inv = 1.0f / (my_max[512] + my_max[512 + other]);
            uint32_t ok[4][8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int un = 2 * j + half;
                if (un < units) {
                    uint32_t ro[16];
                    tmem_ld16(treg + 64u + static_cast<uint32_t>(16 * un), ro);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 8; ++e) ok[j][e] = pack_bf16x2(__uint_as_float(ro[2 * e]) * inv, __uint_as_float(ro[2 * e + 1]) * inv);
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->o_free[g]);
            long long orow[2];                                         // output rows of the two staged rows this thread copies out
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int n = g * 128 + pass * 64 + (t256 >> 2);
                int y = wy + (n >> 4) + p.shift; if (y >= p.H) y -= p.H;
                int x = wx + (n & 15) + p.shift; if (x >= p.W) x -= p.W;
                orow[pass] = static_cast<long long>((b * p.H + y) * p.W + x) * p.ldo + head * p.hdp;
            }
            const int rounds = (units + 1) >> 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < rounds) {
                    if (2 * j + half < units) {
                        // row r of the staging panel: 64 bytes, 16-byte chunk cc stored at cc ^ ((r >> 1) & 3)
                        const uint32_t rowa = stage_g + static_cast<uint32_t>(r * 64);
                        const uint32_t sw = static_cast<uint32_t>((r >> 1) & 3);
                        st_sh_v4(rowa + (((2u * half) ^ sw) << 4), ok[j][0], ok[j][1], ok[j][2], ok[j][3]);
                        st_sh_v4(rowa + (((2u * half + 1u) ^ sw) << 4), ok[j][4], ok[j][5], ok[j][6], ok[j][7]);
                    }
                    named_bar_sync(1 + g, 256);                        // the stream's 128 staged rows are complete
#pragma unroll
                    for (int pass = 0; pass < 2; ++pass) {
                        const int row = pass * 64 + (t256 >> 2), cc = t256 & 3;
                        const int col = 32 * j + 8 * cc;
                        if (col < p.hdp) {
                            uint4 val;
                            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                                         : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                                         : "r"(stage_g + static_cast<uint32_t>(row * 64) + ((static_cast<uint32_t>(cc) ^ static_cast<uint32_t>((row >> 1) & 3)) << 4)));
                            *reinterpret_cast<uint4*>(p.out + orow[pass] + col) = val;
                        }
                    }
                    named_bar_sync(1 + g, 256);                        // copied out: the panel may be overwritten
                }
            }
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

// 16 x 16 windows, hdp in {32, 48, ..., 128}; ADSR_ERR_BAD_SHAPE for anything else (the caller falls back to the mma.sync kernel).
int launch_window_attention_tc16(const void* qkv, long long ldq, void* out, long long ldo, const float* table, int B, int H, int W,
                                 int shift, int nH, int hd, int hdp, int num_sms, cudaStream_t stream) {
    if ((H % 16) || (W % 16) || hdp < 32 || hdp > 128 || (hdp % 16) || nH < 1 || nH > num_sms || shift < 0 || shift >= 16)
        return ADSR_ERR_BAD_SHAPE;
    Tc16Params p{};
    p.qkv = static_cast<const __nv_bfloat16*>(qkv);
    p.ldq = ldq;
    p.out = static_cast<__nv_bfloat16*>(out);
    p.ldo = ldo;
    p.table = table;
    p.B = B; p.H = H; p.W = W; p.shift = shift; p.nH = nH; p.hdp = hdp;
    p.pan = (hdp + 63) / 64;
    p.nbuf = p.pan == 1 ? 2 : 1;
    p.n_win = B * (H / 16) * (W / 16);
    p.n_slots = num_sms / nH;
    if (p.n_slots > p.n_win) p.n_slots = p.n_win;
    p.scale_log2e = (1.0f / sqrtf(static_cast<float>(hd))) * 1.4426950408889634f;
    const int smem_bytes = p.nbuf * (2 * p.pan * 256 * 128 + 2 * p.pan * 128 * 128) + 2 * kStageBytes + 4 * kTabCopy * 4 + 2 * 512 * 4 +
                           static_cast<int>(sizeof(Bars16)) + 128 + 64;
    if (smem_bytes > kSmemLimit) return ADSR_ERR_BAD_SHAPE;
    if (ensure_dynamic_smem(window_attn16_tc_kernel, smem_bytes) != cudaSuccess) return ADSR_ERR_CUDA;
    window_attn16_tc_kernel<<<p.n_slots * nH, kThreads, smem_bytes, stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? ADSR_OK : ADSR_ERR_LAUNCH;
}

}  // namespace adsr
