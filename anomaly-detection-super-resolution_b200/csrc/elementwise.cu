// Bandwidth-bound kernels of the DRCT path: LayerNorm, window index maps / shuffles, the 3-channel
// head convolution fused with patch_embed LayerNorm, the 3-channel tail convolution fused with the
// evaluator's uint8 truncation.  All are coalesced, 16-byte vectorised where the layout allows and
// use warp-shuffle reductions.
#include "adsr_kernels.h"
#include "ptx.cuh"

namespace adsr {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the first C columns of a bf16 row; one warp per row, row held in registers.
// Two-pass statistics (mean, then centred variance) in fp32 like ATen's native_layer_norm.
// MAXV = 16-byte chunks per lane (C <= 256*MAXV).
// `src_row` lets the same body serve the plain and the shift+partition variants.
// ------------------------------------------------------------------------------------------------
template <int MAXV>
__device__ __forceinline__ void ln_row(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                       const float* __restrict__ gamma, const float* __restrict__ beta, int C,
                                       float eps, int lane) {
    const int cpad = (C + 15) & ~15;        // columns written (zeros beyond C)
    float x[MAXV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c0 = (lane + 32 * i) * 8;
        if (c0 < C) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(in + c0));
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                x[i][2 * j] = (c0 + 2 * j < C) ? bf16_lo(w[j]) : 0.f;
                x[i][2 * j + 1] = (c0 + 2 * j + 1 < C) ? bf16_hi(w[j]) : 0.f;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[i][j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += x[i][j];
    }
    const float mean = warp_sum(sum) / static_cast<float>(C);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c0 = (lane + 32 * i) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = (c0 + j < C) ? x[i][j] - mean : 0.f;
            var += d * d;
        }
    }
    const float rstd = rsqrtf(warp_sum(var) / static_cast<float>(C) + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c0 = (lane + 32 * i) * 8;
        if (c0 < cpad) {
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = c0 + j;
                y[j] = (c < C) ? (x[i][j] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c) : 0.f;
            }
            *reinterpret_cast<uint4*>(out + c0) =
                make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
        }
    }
}

// Plain LayerNorm kernel: one warp normalises RPW rows at a time so that RPW independent 16-byte loads per
// lane are in flight (the single-row version was latency bound at ~25 % of HBM bandwidth).
template <int RPW, int CPL>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const __nv_bfloat16* __restrict__ in, long long ldi,
                                                              __nv_bfloat16* __restrict__ out, long long ldo,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int M, int C, float eps) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int cpad = (C + 15) & ~15;
    const float inv_c = 1.0f / static_cast<float>(C);
    for (long long row = (static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5)) * RPW; row < M;
         row += static_cast<long long>(gridDim.x) * warps_per_block * RPW) {
        uint4 u[RPW][CPL];
#pragma unroll
        for (int r = 0; r < RPW; ++r)
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                const int c0 = (lane + 32 * k) * 8;
                u[r][k] = (c0 < C && row + r < M) ? __ldg(reinterpret_cast<const uint4*>(in + (row + r) * ldi + c0))
                                                  : make_uint4(0, 0, 0, 0);
            }
        float x[RPW][CPL][8], mean[RPW], rstd[RPW];
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                const int c0 = (lane + 32 * k) * 8;
                const uint32_t w[4] = {u[r][k].x, u[r][k].y, u[r][k].z, u[r][k].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    x[r][k][2 * j] = (c0 + 2 * j < C) ? bf16_lo(w[j]) : 0.f;
                    x[r][k][2 * j + 1] = (c0 + 2 * j + 1 < C) ? bf16_hi(w[j]) : 0.f;
                    s += x[r][k][2 * j] + x[r][k][2 * j + 1];
                }
            }
            mean[r] = s;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int r = 0; r < RPW; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            mean[r] *= inv_c;
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < CPL; ++k)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float d = ((lane + 32 * k) * 8 + j < C) ? x[r][k][j] - mean[r] : 0.f;
                    v += d * d;
                }
            rstd[r] = v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int r = 0; r < RPW; ++r) rstd[r] += __shfl_xor_sync(0xffffffffu, rstd[r], o);
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const int c0 = (lane + 32 * k) * 8;
            if (c0 >= cpad) continue;
            float g[8], bt[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                g[j] = (c0 + j < C) ? __ldg(gamma + c0 + j) : 0.f;
                bt[j] = (c0 + j < C) ? __ldg(beta + c0 + j) : 0.f;
            }
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                if (row + r >= M) continue;
                const float rs = rsqrtf(rstd[r] * inv_c + eps);
                float y[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = (c0 + j < C) ? (x[r][k][j] - mean[r]) * rs * g[j] + bt[j] : 0.f;
                *reinterpret_cast<uint4*>(out + (row + r) * ldo + c0) =
                    make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
            }
        }
    }
}

// generic fallback (C > 256): one row per warp, MAXV chunks per lane
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_rows_wide_kernel(const __nv_bfloat16* __restrict__ in, long long ldi,
                                                                   __nv_bfloat16* __restrict__ out, long long ldo,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, int M, int C, float eps) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (long long row = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); row < M;
         row += static_cast<long long>(gridDim.x) * warps_per_block)
        ln_row<MAXV>(in + row * ldi, out + row * ldo, gamma, beta, C, eps, lane);
}

// ------------------------------------------------------------------------------------------------
// window index maps (closed form of torch.roll + window_partition and of calculate_mask's regions)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int region_1d(int t, int L, int ws, int shift) {
    return (t >= L - ws ? 1 : 0) + (t >= L - shift ? 1 : 0);
}
__device__ __forceinline__ int window_src_pixel(int w, int n, int H, int W, int ws, int shift, int* region) {
    const int nwx = W / ws;
    const int ys = (w / nwx) * ws + n / ws;
    const int xs = (w % nwx) * ws + n % ws;
    if (region) *region = shift > 0 ? 3 * region_1d(ys, H, ws, shift) + region_1d(xs, W, ws, shift) : 0;
    return ((ys + shift) % H) * W + (xs + shift) % W;
}

__global__ void window_index_map_kernel(int H, int W, int ws, int shift, int32_t* src, int32_t* region) {
    const int N = ws * ws, total = (H / ws) * (W / ws) * N;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int r;
        const int s = window_src_pixel(i / N, i % N, H, W, ws, shift, &r);
        if (src) src[i] = s;
        if (region) region[i] = r;
    }
}

// LayerNorm + cyclic shift + window partition: windows[(b*nW + w)*N + n, :] = LN(x[b, src(w, n), :])
template <int MAXV>
__global__ void __launch_bounds__(256) ln_shift_partition_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                                  __nv_bfloat16* __restrict__ win, long long ldw,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float eps, int B, int H,
                                                                  int W, int C, int ws, int shift) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int N = ws * ws, L = H * W;
    const long long total = static_cast<long long>(B) * L;
    for (long long g = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); g < total;
         g += static_cast<long long>(gridDim.x) * warps_per_block) {
        const int b = static_cast<int>(g / L);
        const int rem = static_cast<int>(g - static_cast<long long>(b) * L);
        const int src = window_src_pixel(rem / N, rem % N, H, W, ws, shift, nullptr);
        ln_row<MAXV>(x + (static_cast<long long>(b) * L + src) * ldx, win + g * ldw, gamma, beta, C, eps, lane);
    }
}

// inverse: x[b, src(w, n), :C] = windows[(b*nW + w)*N + n, :C]   (8-byte vectors: C % 4 == 0)
__global__ void window_reverse_unshift_kernel(const __nv_bfloat16* __restrict__ win, long long ldw,
                                              __nv_bfloat16* __restrict__ x, long long ldx, int B, int H, int W, int C,
                                              int ws, int shift) {
    const int N = ws * ws, L = H * W;
    const int vec_per_row = C / 4;
    const long long total = static_cast<long long>(B) * L * vec_per_row;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long g = i / vec_per_row;
        const int v = static_cast<int>(i - g * vec_per_row);
        const int b = static_cast<int>(g / L);
        const int rem = static_cast<int>(g - static_cast<long long>(b) * L);
        const int src = window_src_pixel(rem / N, rem % N, H, W, ws, shift, nullptr);
        const uint2 val = __ldg(reinterpret_cast<const uint2*>(win + g * ldw + v * 4));
        *reinterpret_cast<uint2*>(x + (static_cast<long long>(b) * L + src) * ldx + v * 4) = val;
    }
}

// ------------------------------------------------------------------------------------------------
// DRCT head: (x - mean) * img_range -> 3x3 conv (nc -> C) + bias -> x0 (bf16) and LayerNorm -> slab.
// One warp per pixel; lane owns channels lane, lane+32, ...  (C <= 256).  Weights live in shared
// memory as [ceil32(C)][kp], kp = nc*9 rounded up to 4 (+4 when that is a multiple of 32 banks... it never is for nc <= 3): a lane
// fetches four taps of a channel with one 16-byte load (row pitch 12 or 28 floats: conflict-free per quarter warp), rows >= C are
// zero so the multiply-add loop carries no channel predicate.  (ncu, round 2: the scalar-load version issued at 85 % of the slots,
// 69 % LSU, 55 % ALU -- address arithmetic and predicates -- for 31 % FMA.)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) drct_head_kernel(const float* __restrict__ x, int B, int nc, int H, int W,
                                                         const float* __restrict__ weight, const float* __restrict__ bias,
                                                         const float* __restrict__ mean, float img_range,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float eps, int C, __nv_bfloat16* __restrict__ x0, long long ld0,
                                                         __nv_bfloat16* __restrict__ slab, long long lds,
                                                         float2* __restrict__ stats_out, int stats_stride) {
    extern __shared__ __align__(16) float sw[];    // [c32][kp] then bias[c32], gamma[C], beta[C]
    const int kk = nc * 9;
    const int kp = (kk + 3) & ~3;                 // 12 / 20 / 28
    const int c32 = (C + 31) & ~31;
    const int nj = c32 >> 5;
    float* sb = sw + c32 * kp;
    float* sg = sb + c32;
    float* sbt = sg + C;
    for (int i = threadIdx.x; i < c32 * kp; i += blockDim.x) {
        const int ch = i / kp, k = i - ch * kp;
        sw[i] = (ch < C && k < kk) ? weight[ch * kk + k] : 0.f;
    }
    for (int i = threadIdx.x; i < c32; i += blockDim.x) sb[i] = i < C ? bias[i] : 0.f;
    for (int i = threadIdx.x; i < C; i += blockDim.x) { sg[i] = gamma[i]; sbt[i] = beta[i]; }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const long long total = static_cast<long long>(B) * H * W;
    const int cpad = (C + 15) & ~15;
    for (long long pix = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); pix < total;
         pix += static_cast<long long>(gridDim.x) * warps_per_block) {
        const int b = static_cast<int>(pix / (H * W));
        const int rem = static_cast<int>(pix - static_cast<long long>(b) * H * W);
        const int y = rem / W, xx = rem - y * W;
        // lanes 0..kk-1 fetch one input tap each, then broadcast by shuffle
        float tapv = 0.f;
        if (lane < kk) {
            const int c = lane / 9, t = lane - c * 9;
            const int iy = y + t / 3 - 1, ix = xx + t % 3 - 1;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W)
                tapv = (__ldg(x + ((static_cast<long long>(b) * nc + c) * H + iy) * W + ix) - __ldg(mean + c)) * img_range;
        }
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = j < nj ? sb[lane + 32 * j] : 0.f;
        // same multiply-add order per channel as ever (k ascending); taps k >= kk meet zero weights
        for (int k4 = 0; k4 < kp; k4 += 4) {
            const float v0 = __shfl_sync(0xffffffffu, tapv, k4), v1 = __shfl_sync(0xffffffffu, tapv, k4 + 1);
            const float v2 = __shfl_sync(0xffffffffu, tapv, k4 + 2), v3 = __shfl_sync(0xffffffffu, tapv, k4 + 3);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < nj) {
                    const float4 w = *reinterpret_cast<const float4*>(sw + (lane + 32 * j) * kp + k4);
                    acc[j] = fmaf(v0, w.x, acc[j]);
                    acc[j] = fmaf(v1, w.y, acc[j]);
                    acc[j] = fmaf(v2, w.z, acc[j]);
                    acc[j] = fmaf(v3, w.w, acc[j]);
                }
            }
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += (lane + 32 * j < C) ? acc[j] : 0.f;
        const float mu = warp_sum(sum) / static_cast<float>(C);
        float var = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = (lane + 32 * j < C) ? acc[j] - mu : 0.f;
            var += d * d;
        }
        const float rstd = rsqrtf(warp_sum(var) / static_cast<float>(C) + eps);
        float osum = 0.f, osq = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int ch = lane + 32 * j;
            if (ch < cpad) {
                const bool ok = ch < C;
                const float o = ok ? (acc[j] - mu) * rstd * sg[ch] + sbt[ch] : 0.f;
                x0[pix * ld0 + ch] = __float2bfloat16(ok ? acc[j] : 0.f);
                slab[pix * lds + ch] = __float2bfloat16(o);
                osum += o;
                osq = fmaf(o, o, osq);
            }
        }
        if (stats_out != nullptr) {                           // row statistics of the slab for the first folded LayerNorm
            osum = warp_sum(osum);
            osq = warp_sum(osq);
            if (lane == 0) {
                stats_out[pix * stats_stride] = make_float2(osum, osq);
                for (int s = 1; s < stats_stride && s < 4; ++s)     // the slots later written by the adjust5 epilogue
                    stats_out[pix * stats_stride + s] = make_float2(0.f, 0.f);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// DRCT tail: conv_last 3x3 (Cin -> nc<=4) on NHWC bf16, + x/img_range + mean, optional fp32 NCHW store
// and the evaluator's uint8 truncation (HWC).  One thread per output pixel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv_last_quant_kernel(const __nv_bfloat16* __restrict__ in, long long ld_in, int B,
                                                               int H, int W, int Cin, const float* __restrict__ weight,
                                                               const float* __restrict__ bias, int nc,
                                                               const float* __restrict__ mean, float inv_img_range,
                                                               float u8_scale, float* __restrict__ out,
                                                               uint8_t* __restrict__ out_u8) {
    extern __shared__ float swl[];                // [9][Cin][4]
    for (int i = threadIdx.x; i < 9 * Cin * 4; i += blockDim.x) {
        const int o = i & 3, c = (i >> 2) % Cin, t = (i >> 2) / Cin;
        swl[i] = o < nc ? weight[(o * Cin + c) * 9 + t] : 0.f;
    }
    __syncthreads();
    const long long total = static_cast<long long>(B) * H * W;
    const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (pix >= total) return;
    const int b = static_cast<int>(pix / (H * W));
    const int rem = static_cast<int>(pix - static_cast<long long>(b) * H * W);
    const int y = rem / W, x = rem - y * W;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t = 0; t < 9; ++t) {
        const int iy = y + t / 3 - 1, ix = x + t % 3 - 1;
        if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
        const __nv_bfloat16* src = in + ((static_cast<long long>(b) * H + iy) * W + ix) * ld_in;
        const float4* wt = reinterpret_cast<const float4*>(swl + t * Cin * 4);
        for (int c0 = 0; c0 < Cin; c0 += 8) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + c0));
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a0 = bf16_lo(w[j]), a1 = bf16_hi(w[j]);
                const float4 w0 = wt[c0 + 2 * j], w1 = wt[c0 + 2 * j + 1];
                acc[0] = fmaf(a0, w0.x, acc[0]); acc[1] = fmaf(a0, w0.y, acc[1]);
                acc[2] = fmaf(a0, w0.z, acc[2]); acc[3] = fmaf(a0, w0.w, acc[3]);
                acc[0] = fmaf(a1, w1.x, acc[0]); acc[1] = fmaf(a1, w1.y, acc[1]);
                acc[2] = fmaf(a1, w1.z, acc[2]); acc[3] = fmaf(a1, w1.w, acc[3]);
            }
        }
    }
    for (int o = 0; o < nc; ++o) {
        const float v = (acc[o] + __ldg(bias + o)) * inv_img_range + __ldg(mean + o);
        if (out) out[((static_cast<long long>(b) * nc + o) * H + y) * W + x] = v;
        if (out_u8) {
            const float q = fminf(fmaxf(v * u8_scale, 0.f), 255.f);
            out_u8[pix * nc + o] = static_cast<uint8_t>(q);   // truncation, as torch .byte()
        }
    }
}

// Tiled variant for Cin = 64 (the DRCT tail) and W % 4 == 0: a block owns an 8 x 64 pixel tile; its 10 x 66 x 64-channel
// halo tile is fetched with coalesced 16-byte cp.async (zero fill outside the image) into shared memory with a pixel pitch
// of 144 bytes (conflict-free 16-byte reads); a thread computes FOUR horizontally adjacent pixels, so every weight vector
// read from shared memory feeds 16 multiply-adds (packed fma.f32x2).  Same fp32 arithmetic as the kernel above.
constexpr int kClTileW = 64, kClTileH = 8, kClPitch = 144, kClCin = 64;
constexpr int kClHaloW = kClTileW + 2, kClHaloH = kClTileH + 2;
constexpr int kClSmemBytes = kClHaloH * kClHaloW * kClPitch + 9 * kClCin * 4 * 4;

__global__ void __launch_bounds__(128) conv_last_quant_tiled_kernel(const __nv_bfloat16* __restrict__ in, long long ld_in, int B,
                                                                     int H, int W, const float* __restrict__ weight,
                                                                     const float* __restrict__ bias, int nc,
                                                                     const float* __restrict__ mean, float inv_img_range,
                                                                     float u8_scale, float* __restrict__ out,
                                                                     uint8_t* __restrict__ out_u8) {
    extern __shared__ __align__(16) uint8_t cl_smem[];
    uint8_t* tile = cl_smem;                                                     // [10][66] pixels x 144 B
    float* swl = reinterpret_cast<float*>(cl_smem + kClHaloH * kClHaloW * kClPitch);   // [9][64][4]
    const int tiles_x = (W + kClTileW - 1) / kClTileW, tiles_y = (H + kClTileH - 1) / kClTileH;
    int bid = blockIdx.x;
    const int tx0 = (bid % tiles_x) * kClTileW;
    bid /= tiles_x;
    const int ty0 = (bid % tiles_y) * kClTileH;
    const int b = bid / tiles_y;

    // halo tile: 660 pixels x 8 chunks of 16 bytes; consecutive threads fetch consecutive chunks (a pixel row is contiguous)
    const uint32_t tile_s = static_cast<uint32_t>(__cvta_generic_to_shared(tile));
    for (int i = threadIdx.x; i < kClHaloH * kClHaloW * 8; i += 128) {
        const int ch = i & 7, px = i >> 3;
        const int hy = px / kClHaloW, hx = px - hy * kClHaloW;
        const int iy = ty0 + hy - 1, ix = tx0 + hx - 1;
        const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
        const __nv_bfloat16* src = in + ((static_cast<long long>(b) * H + (ok ? iy : 0)) * W + (ok ? ix : 0)) * ld_in + ch * 8;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tile_s + static_cast<uint32_t>(px * kClPitch + ch * 16)), "l"(src),
                     "r"(ok ? 16 : 0)
                     : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int i = threadIdx.x; i < 9 * kClCin * 4; i += 128) {
        const int o = i & 3, c = (i >> 2) % kClCin, t = (i >> 2) / kClCin;
        swl[i] = o < nc ? weight[(o * kClCin + c) * 9 + t] : 0.f;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int lx = (threadIdx.x & 15) * 4, ly = threadIdx.x >> 4;               // my 4 pixels: tile row ly, columns lx .. lx + 3
    float2 acc[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int ty = 0; ty < 3; ++ty) {
        const uint8_t* row = tile + ((ly + ty) * kClHaloW + lx) * kClPitch;      // halo column lx = image column lx - 1
#pragma unroll 1
        for (int ch = 0; ch < 8; ++ch) {
            uint4 v[6];                                                          // 8 channels of the 6 input pixels of this row
#pragma unroll
            for (int j = 0; j < 6; ++j) v[j] = *reinterpret_cast<const uint4*>(row + j * kClPitch + ch * 16);
#pragma unroll
            for (int tx = 0; tx < 3; ++tx) {
                const float4* wt = reinterpret_cast<const float4*>(swl) + ((ty * 3 + tx) * kClCin + ch * 8);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float4 w = wt[e];
                    const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint4& u = v[i + tx];
                        const uint32_t word = e < 2 ? u.x : (e < 4 ? u.y : (e < 6 ? u.z : u.w));
                        const float a = (e & 1) ? bf16_hi(word) : bf16_lo(word);
                        const float2 a2 = make_float2(a, a);
                        acc[i][0] = __ffma2_rn(a2, w01, acc[i][0]);
                        acc[i][1] = __ffma2_rn(a2, w23, acc[i][1]);
                    }
                }
            }
        }
    }
    const int y = ty0 + ly, x = tx0 + lx;
    if (y >= H || x >= W) return;                                               // W % 4 == 0: a thread's 4 pixels are all in or all out
    float vout[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a[4] = {acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y};
#pragma unroll
        for (int o = 0; o < 4; ++o) vout[i][o] = o < nc ? (a[o] + __ldg(bias + o)) * inv_img_range + __ldg(mean + o) : 0.f;
    }
    if (out) {
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (o < nc)
                *reinterpret_cast<float4*>(out + ((static_cast<long long>(b) * nc + o) * H + y) * W + x) =
                    make_float4(vout[0][o], vout[1][o], vout[2][o], vout[3][o]);
    }
    if (out_u8) {
        uint8_t* dst = out_u8 + ((static_cast<long long>(b) * H + y) * W + x) * nc;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int o = 0; o < 4; ++o)
                if (o < nc) dst[i * nc + o] = static_cast<uint8_t>(fminf(fmaxf(vout[i][o] * u8_scale, 0.f), 255.f));   // truncation
    }
}

__global__ void quantize_u8_kernel(const float* __restrict__ x, int B, int nc, int H, int W, float u8_scale,
                                   uint8_t* __restrict__ out) {
    const long long total = static_cast<long long>(B) * H * W;
    for (long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pix < total;
         pix += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = pix / (H * W);
        const long long rem = pix - b * H * W;
        for (int c = 0; c < nc; ++c) {
            const float q = fminf(fmaxf(__ldg(x + (b * nc + c) * H * W + rem) * u8_scale, 0.f), 255.f);
            out[pix * nc + c] = static_cast<uint8_t>(q);
        }
    }
}

// uint8 HWC (what the PNG decoder yields) -> fp32 NCHW scaled by rgb_range / 255: the loader's np2Tensor (src/data.py:11-17) on the device
__global__ void u8_to_float_nchw_kernel(const uint8_t* __restrict__ x, int B, int nc, int H, int W, float scale, float* __restrict__ out) {
    const long long total = static_cast<long long>(B) * H * W;
    for (long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pix < total;
         pix += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = pix / (H * W);
        const long long rem = pix - b * H * W;
        for (int c = 0; c < nc; ++c) out[(b * nc + c) * H * W + rem] = static_cast<float>(__ldg(x + pix * nc + c)) * scale;
    }
}

inline int grid_for(long long work_items, int per_block, int cap = 148 * 16) {
    long long g = (work_items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    return static_cast<int>(g < cap ? g : cap);
}
inline int check_launch() { return cudaGetLastError() == cudaSuccess ? ADSR_OK : ADSR_ERR_LAUNCH; }

}  // namespace
}  // namespace adsr

using namespace adsr;

extern "C" int adsr_layernorm_rows(const void* in, int64_t ldi, void* out, int64_t ldo, const float* gamma,
                                   const float* beta, int M, int C, float eps, void* stream) {
    if (M <= 0) return ADSR_OK;
    if (C <= 0 || C > 1024 || (ldi % 8) || (ldo % 8) || ldi < C || ldo < ((C + 15) & ~15)) return ADSR_ERR_BAD_SHAPE;
    if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return ADSR_ERR_BAD_ALIGN;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = grid_for(M, 8);
    auto a = static_cast<const __nv_bfloat16*>(in);
    auto o = static_cast<__nv_bfloat16*>(out);
    if (C <= 256) layernorm_rows_kernel<4, 1><<<grid_for(M, 32), 256, 0, st>>>(a, ldi, o, ldo, gamma, beta, M, C, eps);
    else if (C <= 512) layernorm_rows_kernel<2, 2><<<grid_for(M, 16), 256, 0, st>>>(a, ldi, o, ldo, gamma, beta, M, C, eps);
    else layernorm_rows_wide_kernel<4><<<grid, 256, 0, st>>>(a, ldi, o, ldo, gamma, beta, M, C, eps);
    return check_launch();
}

extern "C" int adsr_window_index_map(int H, int W, int ws, int shift, int32_t* src_index, int32_t* region_id, void* stream) {
    if (ws <= 0 || H % ws || W % ws || shift < 0 || shift >= ws) return ADSR_ERR_BAD_SHAPE;
    window_index_map_kernel<<<grid_for(static_cast<long long>(H) * W, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        H, W, ws, shift, src_index, region_id);
    return check_launch();
}

extern "C" int adsr_ln_shift_partition(const void* x, int64_t ldx, void* windows, int64_t ldw, const float* gamma,
                                       const float* beta, float eps, int B, int H, int W, int C, int ws, int shift,
                                       void* stream) {
    if (B <= 0) return ADSR_OK;
    if (ws <= 0 || H % ws || W % ws || shift < 0 || shift >= ws) return ADSR_ERR_BAD_SHAPE;
    if (C <= 0 || C > 1024 || (ldx % 8) || (ldw % 8) || ldx < C || ldw < ((C + 15) & ~15)) return ADSR_ERR_BAD_SHAPE;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(windows) & 15)) return ADSR_ERR_BAD_ALIGN;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = grid_for(static_cast<long long>(B) * H * W, 8);
    auto a = static_cast<const __nv_bfloat16*>(x);
    auto o = static_cast<__nv_bfloat16*>(windows);
    if (C <= 256) ln_shift_partition_kernel<1><<<grid, 256, 0, st>>>(a, ldx, o, ldw, gamma, beta, eps, B, H, W, C, ws, shift);
    else if (C <= 512) ln_shift_partition_kernel<2><<<grid, 256, 0, st>>>(a, ldx, o, ldw, gamma, beta, eps, B, H, W, C, ws, shift);
    else ln_shift_partition_kernel<4><<<grid, 256, 0, st>>>(a, ldx, o, ldw, gamma, beta, eps, B, H, W, C, ws, shift);
    return check_launch();
}

extern "C" int adsr_window_reverse_unshift(const void* windows, int64_t ldw, void* x, int64_t ldx, int B, int H, int W,
                                           int C, int ws, int shift, void* stream) {
    if (B <= 0) return ADSR_OK;
    if (ws <= 0 || H % ws || W % ws || shift < 0 || shift >= ws || (C % 4) || (ldw % 4) || (ldx % 4)) return ADSR_ERR_BAD_SHAPE;
    if ((reinterpret_cast<uintptr_t>(x) & 7) || (reinterpret_cast<uintptr_t>(windows) & 7)) return ADSR_ERR_BAD_ALIGN;
    const long long items = static_cast<long long>(B) * H * W * (C / 4);
    window_reverse_unshift_kernel<<<grid_for(items, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(windows), ldw, static_cast<__nv_bfloat16*>(x), ldx, B, H, W, C, ws, shift);
    return check_launch();
}

extern "C" int adsr_drct_head(const float* x_nchw, int B, int nc, int H, int W, const float* weight, const float* bias,
                              const float* mean, float img_range, const float* ln_gamma, const float* ln_beta, float eps,
                              int C, void* x0, int64_t ld0, void* slab, int64_t lds, float* stats_out, int stats_out_stride,
                              void* stream) {
    if (B <= 0) return ADSR_OK;
    if (nc < 1 || nc > 3 || C <= 0 || C > 256) return ADSR_ERR_BAD_SHAPE;
    const int cpad = (C + 15) & ~15;
    if (ld0 < cpad || lds < cpad || (stats_out != nullptr && stats_out_stride < 2)) return ADSR_ERR_BAD_SHAPE;
    const int c32 = (C + 31) & ~31, kp = (nc * 9 + 3) & ~3;
    const size_t smem = (static_cast<size_t>(c32) * kp + c32 + 2 * C) * sizeof(float);
    if (smem > 48 * 1024) return ADSR_ERR_BAD_SHAPE;
    drct_head_kernel<<<grid_for(static_cast<long long>(B) * H * W, 8), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        x_nchw, B, nc, H, W, weight, bias, mean, img_range, ln_gamma, ln_beta, eps, C, static_cast<__nv_bfloat16*>(x0), ld0,
        static_cast<__nv_bfloat16*>(slab), lds, reinterpret_cast<float2*>(stats_out), stats_out_stride);
    return check_launch();
}

extern "C" int adsr_conv_last_quant(const void* in, int64_t ld_in, int B, int H, int W, int Cin, const float* weight,
                                    const float* bias, int nc, const float* mean, float img_range, float rgb_range,
                                    float* out_nchw, uint8_t* out_u8_hwc, void* stream) {
    if (B <= 0) return ADSR_OK;
    if (nc < 1 || nc > 4 || Cin <= 0 || (Cin % 8) || (ld_in % 8)) return ADSR_ERR_BAD_SHAPE;
    if (reinterpret_cast<uintptr_t>(in) & 15) return ADSR_ERR_BAD_ALIGN;
    if (Cin == kClCin && (W % 4) == 0) {
        static bool attr_set = false;
        if (!attr_set) {
            if (ensure_dynamic_smem(conv_last_quant_tiled_kernel, kClSmemBytes) != cudaSuccess)
                return ADSR_ERR_CUDA;
            attr_set = true;
        }
        const long long blocks = static_cast<long long>(B) * ((H + kClTileH - 1) / kClTileH) * ((W + kClTileW - 1) / kClTileW);
        if (blocks > 0x7fffffffLL) return ADSR_ERR_BAD_SHAPE;
        conv_last_quant_tiled_kernel<<<static_cast<int>(blocks), 128, kClSmemBytes, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16*>(in), ld_in, B, H, W, weight, bias, nc, mean, 1.0f / img_range,
            static_cast<float>(255.0 / static_cast<double>(rgb_range)), out_nchw, out_u8_hwc);
        return check_launch();
    }
    const size_t smem = static_cast<size_t>(9) * Cin * 4 * sizeof(float);
    if (smem > 48 * 1024) return ADSR_ERR_BAD_SHAPE;
    const long long total = static_cast<long long>(B) * H * W;
    const int grid = static_cast<int>((total + 127) / 128);
    conv_last_quant_kernel<<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(in), ld_in, B, H, W, Cin, weight, bias, nc, mean, 1.0f / img_range,
        static_cast<float>(255.0 / static_cast<double>(rgb_range)), out_nchw, out_u8_hwc);
    return check_launch();
}

extern "C" int adsr_u8_to_float_nchw(const uint8_t* x_u8_hwc, int B, int nc, int H, int W, float rgb_range, float* out_nchw, void* stream) {
    if (B <= 0) return ADSR_OK;
    if (x_u8_hwc == nullptr || out_nchw == nullptr || nc < 1 || H < 1 || W < 1) return ADSR_ERR_BAD_SHAPE;
    u8_to_float_nchw_kernel<<<grid_for(static_cast<long long>(B) * H * W, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x_u8_hwc, B, nc, H, W, static_cast<float>(static_cast<double>(rgb_range) / 255.0), out_nchw);
    return check_launch();
}

extern "C" int adsr_quantize_u8(const float* x_nchw, int B, int nc, int H, int W, float rgb_range, uint8_t* out_u8_hwc,
                                void* stream) {
    if (B <= 0) return ADSR_OK;
    quantize_u8_kernel<<<grid_for(static_cast<long long>(B) * H * W, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x_nchw, B, nc, H, W, static_cast<float>(255.0 / static_cast<double>(rgb_range)), out_u8_hwc);
    return check_launch();
}
