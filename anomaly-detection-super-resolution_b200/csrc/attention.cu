// Fused (shifted-)window multi-head attention for DRCT Swin blocks.
//
// One CTA = 64 consecutive window slots (queries) of one head.  The cyclic shift, window partition,
// relative-position bias gather, shift mask, softmax, P*V, window reverse and un-shift of the
// reference (src/drct.py:482-505, 193-220, 271-299, 449-470) collapse into:
//   gather q/k/v rows through the closed-form index map -> S = q k^T (tensor cores, fp32 accumulate)
//   -> S*scale + table[rel_index] + (-100 if region ids differ) -> online softmax -> O = P v
//   -> scatter O back to the query's ORIGINAL token row (the inverse permutation is free).
// Head dims in DRCT are 30/53/122/46/77 (padded to 32/64/128/48/80 by the QKV weight packing), windows
// hold N = 64 (ws 8) or 256 (ws 16) tokens: tiles this small do not fill a 128-row tcgen05 MMA, so
// this kernel uses warp-level mma.sync m16n8k16 bf16 tiles with the softmax entirely in registers.
// (The dense contractions - 94 % of the FLOPs - run on tcgen05 in tc_gemm.cu.)
#include "adsr_kernels.h"
#include "ptx.cuh"

namespace adsr {
namespace {

struct AttnParams {
    const __nv_bfloat16* qkv;
    long long ldq;
    __nv_bfloat16* out;
    long long ldo;
    const float* table;   // [(2ws-1)^2, nH]
    int B, H, W, ws, shift, nH, hdp;
    int N, nW, total_slots;
    float scale_log2e;    // hd^-0.5 * log2(e)
};

__device__ __forceinline__ int region_1d(int t, int L, int ws, int shift) {
    return (t >= L - ws ? 1 : 0) + (t >= L - shift ? 1 : 0);
}

// slot g (global, windowed order) -> token row; also window-local coords and mask region id
__device__ __forceinline__ int slot_to_token(const AttnParams& p, int g, int& info, int& win) {
    if (g >= p.total_slots) { info = 0; win = -1; return -1; }
    win = g / p.N;
    const int n = g - win * p.N;
    const int b = win / p.nW;
    const int w = win - b * p.nW;
    const int nwx = p.W / p.ws;
    const int yi = n / p.ws, xi = n - yi * p.ws;
    const int ys = (w / nwx) * p.ws + yi;
    const int xs = (w % nwx) * p.ws + xi;
    int y = ys + p.shift; if (y >= p.H) y -= p.H;
    int x = xs + p.shift; if (x >= p.W) x -= p.W;
    const int id = p.shift > 0 ? 3 * region_1d(ys, p.H, p.ws, p.shift) + region_1d(xs, p.W, p.ws, p.shift) : 0;
    info = xi | (yi << 8) | (id << 16);
    return (b * p.H + y) * p.W + x;
}

template <int KD>
__global__ void __launch_bounds__(128) window_attn_kernel(const AttnParams p) {
    constexpr int HDP = KD * 16;
    constexpr int PITCH = HDP + 8;          // bf16 elements; 16 B pad keeps ldmatrix conflict-free
    constexpr int CH = HDP / 8;             // 16-byte chunks per row
    extern __shared__ __align__(16) uint8_t smem[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);
    __nv_bfloat16* sK = sQ + 64 * PITCH;
    __nv_bfloat16* sV = sK + 64 * PITCH;
    int* sTokQ = reinterpret_cast<int*>(sV + 64 * PITCH);
    int* sInfoQ = sTokQ + 64;
    int* sWinQ = sInfoQ + 64;
    int* sTokK = sWinQ + 64;
    int* sInfoK = sTokK + 64;
    int* sWinK = sInfoK + 64;
    float* sBias = reinterpret_cast<float*>(sWinK + 64);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int h = blockIdx.y;
    const int g0 = blockIdx.x * 64;
    const int nbias = (2 * p.ws - 1) * (2 * p.ws - 1);

    for (int i = tid; i < nbias; i += 128) sBias[i] = __ldg(p.table + static_cast<long long>(i) * p.nH + h);
    if (tid < 64) {
        int info, win;
        sTokQ[tid] = slot_to_token(p, g0 + tid, info, win);
        sInfoQ[tid] = info;
        sWinQ[tid] = win;
    }
    __syncthreads();

    // ---- gather Q rows
    const long long qcol = static_cast<long long>(h) * HDP;
    for (int idx = tid; idx < 64 * CH; idx += 128) {
        const int r = idx / CH, c = idx - r * CH;
        const int tok = sTokQ[r];
        uint4 v = make_uint4(0, 0, 0, 0);
        if (tok >= 0) v = __ldg(reinterpret_cast<const uint4*>(p.qkv + tok * p.ldq + qcol + c * 8));
        *reinterpret_cast<uint4*>(sQ + r * PITCH + c * 8) = v;
    }
    __syncthreads();

    uint32_t qf[KD][4];
#pragma unroll
    for (int kd = 0; kd < KD; ++kd)
        ldmatrix_x4(qf[kd], smem_u32(sQ + (warp * 16 + (lane & 15)) * PITCH + kd * 16 + (lane >> 4) * 8));

    const int gq = lane >> 2, tq = lane & 3;
    const int r0 = warp * 16 + gq, r1 = r0 + 8;
    const int infoq0 = sInfoQ[r0], infoq1 = sInfoQ[r1];
    const int winq0 = sWinQ[r0], winq1 = sWinQ[r1];
    const int bw = 2 * p.ws - 1;
    // rel index = (yq - yk + ws-1)*(2ws-1) + (xq - xk + ws-1)
    const int qb0 = (((infoq0 >> 8) & 255) + p.ws - 1) * bw + (infoq0 & 255) + p.ws - 1;
    const int qb1 = (((infoq1 >> 8) & 255) + p.ws - 1) * bw + (infoq1 & 255) + p.ws - 1;
    const int idq0 = infoq0 >> 16, idq1 = infoq1 >> 16;

    float o[2 * KD][4];
#pragma unroll
    for (int i = 0; i < 2 * KD; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    const bool small_win = p.N < 64;
    const int nkb = small_win ? 1 : p.N / 64;
    const int kbase0 = small_win ? g0 : (g0 / p.N) * p.N;
    const long long kcol = static_cast<long long>(p.nH + h) * HDP;
    const long long vcol = static_cast<long long>(2 * p.nH + h) * HDP;

    for (int kb = 0; kb < nkb; ++kb) {
        __syncthreads();   // previous block's sK/sV reads are done
        if (tid < 64) {
            int info, win;
            sTokK[tid] = slot_to_token(p, kbase0 + kb * 64 + tid, info, win);
            sInfoK[tid] = info;
            sWinK[tid] = win;
        }
        __syncthreads();
        for (int idx = tid; idx < 64 * CH; idx += 128) {
            const int r = idx / CH, c = idx - r * CH;
            const int tok = sTokK[r];
            uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
            if (tok >= 0) {
                kv = __ldg(reinterpret_cast<const uint4*>(p.qkv + tok * p.ldq + kcol + c * 8));
                vv = __ldg(reinterpret_cast<const uint4*>(p.qkv + tok * p.ldq + vcol + c * 8));
            }
            *reinterpret_cast<uint4*>(sK + r * PITCH + c * 8) = kv;
            *reinterpret_cast<uint4*>(sV + r * PITCH + c * 8) = vv;
        }
        __syncthreads();

        // ---- S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int kd = 0; kd < KD; ++kd) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t bfr[4];
                ldmatrix_x4(bfr, smem_u32(sK + (8 * (2 * np + (lane >> 4)) + (lane & 7)) * PITCH + kd * 16 + ((lane >> 3) & 1) * 8));
                mma_bf16_16816(s[2 * np], qf[kd], bfr[0], bfr[1]);
                mma_bf16_16816(s[2 * np + 1], qf[kd], bfr[2], bfr[3]);
            }
        }
        // ---- logits (in log2 units): s*scale + bias + mask
        float mx0 = -INFINITY, mx1 = -INFINITY;
        constexpr float LOG2E = 1.4426950408889634f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int kc = nt * 8 + tq * 2 + e;
                const int ik = sInfoK[kc];
                const int wk = sWinK[kc];
                const int kb_off = ((ik >> 8) & 255) * bw + (ik & 255);
                const int idk = ik >> 16;
                float v0 = s[nt][e] * p.scale_log2e + (sBias[qb0 - kb_off] + (idq0 != idk ? -100.f : 0.f)) * LOG2E;
                float v1 = s[nt][2 + e] * p.scale_log2e + (sBias[qb1 - kb_off] + (idq1 != idk ? -100.f : 0.f)) * LOG2E;
                if (wk < 0 || (small_win && wk != winq0)) v0 = -INFINITY;
                if (wk < 0 || (small_win && wk != winq1)) v1 = -INFINITY;
                s[nt][e] = v0;
                s[nt][2 + e] = v1;
                mx0 = fmaxf(mx0, v0);
                mx1 = fmaxf(mx1, v1);
            }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        // rows of padding slots (no valid key) keep mn = -inf: guard the subtraction
        const float sub0 = mn0 == -INFINITY ? 0.f : mn0, sub1 = mn1 == -INFINITY ? 0.f : mn1;
        const float c0 = exp2f(m0 - sub0), c1 = exp2f(m1 - sub1);
        m0 = mn0; m1 = mn1;
        l0 *= c0; l1 *= c1;
#pragma unroll
        for (int i = 0; i < 2 * KD; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
        uint32_t pf[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p00 = exp2f(s[nt][0] - sub0), p01 = exp2f(s[nt][1] - sub0);
            const float p10 = exp2f(s[nt][2] - sub1), p11 = exp2f(s[nt][3] - sub1);
            l0 += p00 + p01;
            l1 += p10 + p11;
            const int j = nt >> 1;
            if ((nt & 1) == 0) { pf[j][0] = pack_bf16x2(p00, p01); pf[j][1] = pack_bf16x2(p10, p11); }
            else               { pf[j][2] = pack_bf16x2(p00, p01); pf[j][3] = pack_bf16x2(p10, p11); }
        }
        // ---- O += P V
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int nd = 0; nd < 2 * KD; nd += 2) {
                uint32_t bfr[4];
                ldmatrix_x4_trans(bfr, smem_u32(sV + (16 * j + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + 8 * (nd + (lane >> 4))));
                mma_bf16_16816(o[nd], pf[j], bfr[0], bfr[1]);
                mma_bf16_16816(o[nd + 1], pf[j], bfr[2], bfr[3]);
            }
        }
    }

    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;

    // ---- stage this warp's 16 output rows in its own sQ rows, then store 16 B chunks
    __syncwarp();
#pragma unroll
    for (int nd = 0; nd < 2 * KD; ++nd) {
        *reinterpret_cast<uint32_t*>(sQ + r0 * PITCH + nd * 8 + tq * 2) = pack_bf16x2(o[nd][0] * inv0, o[nd][1] * inv0);
        *reinterpret_cast<uint32_t*>(sQ + r1 * PITCH + nd * 8 + tq * 2) = pack_bf16x2(o[nd][2] * inv1, o[nd][3] * inv1);
    }
    __syncwarp();
    for (int idx = lane; idx < 16 * CH; idx += 32) {
        const int r = warp * 16 + idx / CH, c = idx % CH;
        const int tok = sTokQ[r];
        if (tok >= 0)
            *reinterpret_cast<uint4*>(p.out + tok * p.ldo + qcol + c * 8) = *reinterpret_cast<const uint4*>(sQ + r * PITCH + c * 8);
    }
}

// ------------------------------------------------------------------------------------------------
// Fast path for N == 64 (window 8x8, the headline configuration): one CTA = one window x HPC heads.
// All q|k|v rows of the window are pulled into shared memory with cp.async (16 B, L1 bypass) in one
// sweep (deep memory-level parallelism, 3 row segments of HPC*hdp*2 bytes per token), each of the
// 4 warps owns 16 query rows and loops over the CTA's heads, and the output rows are staged over the
// (consumed) q columns so that global stores are contiguous HPC*hdp*2-byte segments.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <int KD, int HPC>
__global__ void __launch_bounds__(128) window_attn64_kernel(const AttnParams p) {
    constexpr int HDP = KD * 16;
    constexpr int SEG = HPC * HDP;            // elements per q / k / v segment of a row
    constexpr int PITCH = 3 * SEG + 8;        // odd number of 16 B chunks: conflict-free ldmatrix
    constexpr int CH = 3 * SEG / 8;           // 16 B chunks per row
    extern __shared__ __align__(16) uint8_t smem[];
    __nv_bfloat16* sX = reinterpret_cast<__nv_bfloat16*>(smem);
    int* sTok = reinterpret_cast<int*>(sX + 64 * PITCH);
    int* sInfo = sTok + 64;
    float* sBias = reinterpret_cast<float*>(sInfo + 64);      // [HPC][nbias]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int hg = blockIdx.y;                                 // head group
    const int g0 = blockIdx.x * 64;
    const int nbias = (2 * p.ws - 1) * (2 * p.ws - 1);

    if (tid < 64) {
        int info, win;
        sTok[tid] = slot_to_token(p, g0 + tid, info, win);
        sInfo[tid] = info;
    }
    for (int i = tid; i < HPC * nbias; i += 128) {
        const int hl = i / nbias, e = i - hl * nbias;
        sBias[i] = __ldg(p.table + static_cast<long long>(e) * p.nH + hg * HPC + hl);
    }
    __syncthreads();
    {
        const long long seg_stride = static_cast<long long>(p.nH) * HDP;          // q -> k -> v column distance
        const long long col0 = static_cast<long long>(hg) * SEG;
        const uint32_t sbase = smem_u32(sX);
        for (int idx = tid; idx < 64 * CH; idx += 128) {
            const int r = idx / CH, c = idx - r * CH;
            const int part = c / (SEG / 8), cc = c - part * (SEG / 8);
            const int tok = sTok[r];
            const uint32_t dst = sbase + static_cast<uint32_t>((r * PITCH + c * 8) * 2);
            if (tok >= 0) {
                cp_async16(dst, p.qkv + tok * p.ldq + part * seg_stride + col0 + cc * 8);
            } else {
                asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0) : "memory");
            }
        }
        cp_async_wait_all();
    }
    __syncthreads();

    const int gq = lane >> 2, tq = lane & 3;
    const int r0 = warp * 16 + gq, r1 = r0 + 8;
    const int infoq0 = sInfo[r0], infoq1 = sInfo[r1];
    const int bw = 2 * p.ws - 1;
    const int qb0 = (((infoq0 >> 8) & 255) + p.ws - 1) * bw + (infoq0 & 255) + p.ws - 1;
    const int qb1 = (((infoq1 >> 8) & 255) + p.ws - 1) * bw + (infoq1 & 255) + p.ws - 1;
    const int idq0 = infoq0 >> 16, idq1 = infoq1 >> 16;
    constexpr float LOG2E = 1.4426950408889634f;

    // per-key mask / bias offsets do not depend on the head: hoist them (16 keys per thread)
    int kboff[16];
    float mk0[16], mk1[16];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int kc = nt * 8 + tq * 2 + e;
            const int ik = sInfo[kc];
            kboff[nt * 2 + e] = ((ik >> 8) & 255) * bw + (ik & 255);
            const int idk = ik >> 16;
            const bool kvalid = sTok[kc] >= 0;
            mk0[nt * 2 + e] = !kvalid ? -INFINITY : (idq0 != idk ? -100.f * LOG2E : 0.f);
            mk1[nt * 2 + e] = !kvalid ? -INFINITY : (idq1 != idk ? -100.f * LOG2E : 0.f);
        }

#pragma unroll 1
    for (int hl = 0; hl < HPC; ++hl) {
        const __nv_bfloat16* sQ = sX + hl * HDP;
        const __nv_bfloat16* sK = sX + SEG + hl * HDP;
        const __nv_bfloat16* sV = sX + 2 * SEG + hl * HDP;
        const float* bias = sBias + hl * nbias;
        uint32_t qf[KD][4];
#pragma unroll
        for (int kd = 0; kd < KD; ++kd)
            ldmatrix_x4(qf[kd], smem_u32(sQ + (warp * 16 + (lane & 15)) * PITCH + kd * 16 + (lane >> 4) * 8));
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int kd = 0; kd < KD; ++kd) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t bfr[4];
                ldmatrix_x4(bfr, smem_u32(sK + (8 * (2 * np + (lane >> 4)) + (lane & 7)) * PITCH + kd * 16 + ((lane >> 3) & 1) * 8));
                mma_bf16_16816(s[2 * np], qf[kd], bfr[0], bfr[1]);
                mma_bf16_16816(s[2 * np + 1], qf[kd], bfr[2], bfr[3]);
            }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ki = nt * 2 + e;
                const float v0 = fmaf(s[nt][e], p.scale_log2e, fmaf(bias[qb0 - kboff[ki]], LOG2E, mk0[ki]));
                const float v1 = fmaf(s[nt][2 + e], p.scale_log2e, fmaf(bias[qb1 - kboff[ki]], LOG2E, mk1[ki]));
                s[nt][e] = v0;
                s[nt][2 + e] = v1;
                mx0 = fmaxf(mx0, v0);
                mx1 = fmaxf(mx1, v1);
            }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float sub0 = mx0 == -INFINITY ? 0.f : mx0, sub1 = mx1 == -INFINITY ? 0.f : mx1;
        float l0 = 0.f, l1 = 0.f;
        uint32_t pf[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p00 = exp2f(s[nt][0] - sub0), p01 = exp2f(s[nt][1] - sub0);
            const float p10 = exp2f(s[nt][2] - sub1), p11 = exp2f(s[nt][3] - sub1);
            l0 += p00 + p01;
            l1 += p10 + p11;
            const int j = nt >> 1;
            if ((nt & 1) == 0) { pf[j][0] = pack_bf16x2(p00, p01); pf[j][1] = pack_bf16x2(p10, p11); }
            else               { pf[j][2] = pack_bf16x2(p00, p01); pf[j][3] = pack_bf16x2(p10, p11); }
        }
        float o[2 * KD][4];
#pragma unroll
        for (int i = 0; i < 2 * KD; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int nd = 0; nd < 2 * KD; nd += 2) {
                uint32_t bfr[4];
                ldmatrix_x4_trans(bfr, smem_u32(sV + (16 * j + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + 8 * (nd + (lane >> 4))));
                mma_bf16_16816(o[nd], pf[j], bfr[0], bfr[1]);
                mma_bf16_16816(o[nd + 1], pf[j], bfr[2], bfr[3]);
            }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
        __syncwarp();                                   // every lane's q fragments of this head are in registers
        __nv_bfloat16* sO = sX + hl * HDP;              // overwrite this warp's own q rows with the output
#pragma unroll
        for (int nd = 0; nd < 2 * KD; ++nd) {
            *reinterpret_cast<uint32_t*>(sO + r0 * PITCH + nd * 8 + tq * 2) = pack_bf16x2(o[nd][0] * inv0, o[nd][1] * inv0);
            *reinterpret_cast<uint32_t*>(sO + r1 * PITCH + nd * 8 + tq * 2) = pack_bf16x2(o[nd][2] * inv1, o[nd][3] * inv1);
        }
    }
    __syncwarp();
    constexpr int OCH = SEG / 8;
    for (int idx = lane; idx < 16 * OCH; idx += 32) {
        const int r = warp * 16 + idx / OCH, c = idx % OCH;
        const int tok = sTok[r];
        if (tok >= 0)
            *reinterpret_cast<uint4*>(p.out + tok * p.ldo + static_cast<long long>(hg) * SEG + c * 8) =
                *reinterpret_cast<const uint4*>(sX + r * PITCH + c * 8);
    }
}

template <int KD, int HPC>
int launch_attn64(const AttnParams& p, cudaStream_t stream) {
    constexpr int HDP = KD * 16;
    const int nbias = (2 * p.ws - 1) * (2 * p.ws - 1);
    const size_t smem = static_cast<size_t>(64) * (3 * HPC * HDP + 8) * 2 + 2 * 64 * 4 + static_cast<size_t>(HPC) * nbias * 4;
    if (smem > 48 * 1024) {
        if (ensure_dynamic_smem(window_attn64_kernel<KD, HPC>, static_cast<int>(smem)) != cudaSuccess)
            return ADSR_ERR_CUDA;
    }
    dim3 grid((p.total_slots + 63) / 64, p.nH / HPC);
    window_attn64_kernel<KD, HPC><<<grid, 128, smem, stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? ADSR_OK : ADSR_ERR_LAUNCH;
}

template <int KD>
int dispatch_attn64(const AttnParams& p, cudaStream_t stream) {
    // heads per CTA: the largest of {3,2,1} that divides nH and keeps the q|k|v tile around 60 KB
    if constexpr (KD <= 3) { if (p.nH % 3 == 0) return launch_attn64<KD, 3>(p, stream); }
    if constexpr (KD <= 5) { if (p.nH % 2 == 0) return launch_attn64<KD, 2>(p, stream); }
    return launch_attn64<KD, 1>(p, stream);
}

template <int KD>
int launch_attn(const AttnParams& p, cudaStream_t stream) {
    constexpr int HDP = KD * 16;
    const int nbias = (2 * p.ws - 1) * (2 * p.ws - 1);
    const size_t smem = 3 * 64 * (HDP + 8) * 2 + 6 * 64 * 4 + static_cast<size_t>(nbias) * 4;
    if (smem > 200 * 1024) return ADSR_ERR_BAD_SHAPE;
    if (smem > 48 * 1024) {
        if (ensure_dynamic_smem(window_attn_kernel<KD>, static_cast<int>(smem)) != cudaSuccess)
            return ADSR_ERR_CUDA;
    }
    dim3 grid((p.total_slots + 63) / 64, p.nH);
    window_attn_kernel<KD><<<grid, 128, smem, stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? ADSR_OK : ADSR_ERR_LAUNCH;
}

}  // namespace
}  // namespace adsr

// tuning / test hook (not part of the public header): 0 forces the mma.sync kernels even for 8x8 windows, 2 forces the tcgen05
// kernel for every head width it supports
static int g_attn_tc_enabled = 1;
extern "C" void adsr_debug_set_attention_tc(int enabled) { g_attn_tc_enabled = enabled; }

extern "C" int adsr_window_attention(const void* qkv, int64_t ldq, void* out, int64_t ldo, const float* bias_table,
                                     int B, int H, int W, int ws, int shift, int nH, int hd, int hdp, void* stream) {
    using namespace adsr;
    if (B <= 0) return ADSR_OK;
    if (ws <= 0 || H % ws || W % ws || shift < 0 || shift >= ws || ws > 128) return ADSR_ERR_BAD_SHAPE;
    const int N = ws * ws;
    if (!((N < 64 && 64 % N == 0) || (N >= 64 && N % 64 == 0))) return ADSR_ERR_BAD_SHAPE;
    if (hdp % 16 || hdp < 16 || hdp > 128 || hd > hdp || hd <= 0) return ADSR_ERR_BAD_SHAPE;
    if ((ldq % 8) || (ldo % 8) || (reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
        return ADSR_ERR_BAD_ALIGN;
    AttnParams p;
    p.qkv = static_cast<const __nv_bfloat16*>(qkv);
    p.ldq = ldq;
    p.out = static_cast<__nv_bfloat16*>(out);
    p.ldo = ldo;
    p.table = bias_table;
    p.B = B; p.H = H; p.W = W; p.ws = ws; p.shift = shift; p.nH = nH; p.hdp = hdp;
    p.N = N;
    p.nW = (H / ws) * (W / ws);
    p.total_slots = B * p.nW * N;
    p.scale_log2e = (1.0f / sqrtf(static_cast<float>(hd))) * 1.4426950408889634f;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (ws == 16 && g_attn_tc_enabled != 0) {
        // DRCT-L at 64 px LR (BASELINE configs[3]): tcgen05 kernel with S / P / O of the 256-key windows in TMEM
        int num_sms = 0, dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            return ADSR_ERR_CUDA;
        const int rc = launch_window_attention_tc16(qkv, ldq, out, ldo, bias_table, B, H, W, shift, nH, hd, hdp, num_sms, st);
        if (rc != ADSR_ERR_BAD_SHAPE) return rc;
    }
    if (ws == 8 && (g_attn_tc_enabled == 2 || (g_attn_tc_enabled == 1 && hdp <= 64))) {
        // (heads wider than 64 channels need two panels per operand and only a 2-deep ring: the mma.sync kernel is still faster there)
        // DRCT-L shape: tcgen05 kernel (S and O in TMEM); shapes it does not cover fall through to the mma.sync kernels
        static int num_sms = 0;
        if (num_sms == 0) {
            int dev = 0;
            if (cudaGetDevice(&dev) != cudaSuccess ||
                cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
                return ADSR_ERR_CUDA;
        }
        const int rc = launch_window_attention_tc(qkv, ldq, out, ldo, bias_table, B, H, W, shift, nH, hd, hdp, num_sms, st);
        if (rc != ADSR_ERR_BAD_SHAPE) return rc;
    }
    if (N == 64) {
        switch (hdp / 16) {
            case 1: return dispatch_attn64<1>(p, st);
            case 2: return dispatch_attn64<2>(p, st);
            case 3: return dispatch_attn64<3>(p, st);
            case 4: return dispatch_attn64<4>(p, st);
            case 5: return dispatch_attn64<5>(p, st);
            case 6: return dispatch_attn64<6>(p, st);
            case 7: return dispatch_attn64<7>(p, st);
            case 8: return dispatch_attn64<8>(p, st);
        }
    }
    switch (hdp / 16) {
        case 1: return launch_attn<1>(p, st);
        case 2: return launch_attn<2>(p, st);
        case 3: return launch_attn<3>(p, st);
        case 4: return launch_attn<4>(p, st);
        case 5: return launch_attn<5>(p, st);
        case 6: return launch_attn<6>(p, st);
        case 7: return launch_attn<7>(p, st);
        case 8: return launch_attn<8>(p, st);
    }
    return ADSR_ERR_BAD_SHAPE;
}
