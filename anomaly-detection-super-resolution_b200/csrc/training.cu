// First slice of the training step (reference: Trainer.train src/trainer.py:141-227, Loss "1*L1"
// src/loss.py:83-84, 108-121, Adam src/trainer.py:49-59):
//   * L1 loss (nn.L1Loss, mean) and its gradient on the SR image in one pass,
//   * backward of the last layer, conv_last (3x3, 64 -> n_colors, src/drct.py:847, 895): input gradient (bf16 NHWC, the layout the
//     rest of the backward would consume), weight and bias gradients (deterministic two-stage reduction),
//   * a fused multi-tensor Adam step (torch.optim.Adam semantics, weight_decay folded into the gradient) over a chunk table.
// These are bandwidth / ALU-bound kernels (conv_last has N = 3: no tensor-core shape), coalesced and vectorised; the backward of the
// tcgen05 blocks (attention, MLP, implicit-GEMM convs) is not built yet.
#include "adsr_kernels.h"
#include "ptx.cuh"

namespace adsr {
namespace {

// ---- L1 loss + gradient: grad = sign(sr - hr) * grad_scale / n ; partial[block] = sum |sr - hr| (fp64)
__global__ void __launch_bounds__(256) l1_loss_grad_kernel(const float* __restrict__ sr, const float* __restrict__ hr, long long n,
                                                           float gscale, float* __restrict__ grad, double* __restrict__ partial) {
    __shared__ double red[8];
    double acc = 0.0;
    const long long n4 = n >> 2;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(sr) + i), b = __ldg(reinterpret_cast<const float4*>(hr) + i);
        const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
        acc += static_cast<double>(fabsf(d0)) + static_cast<double>(fabsf(d1)) + static_cast<double>(fabsf(d2)) + static_cast<double>(fabsf(d3));
        float4 g;
        g.x = d0 > 0.f ? gscale : (d0 < 0.f ? -gscale : 0.f);
        g.y = d1 > 0.f ? gscale : (d1 < 0.f ? -gscale : 0.f);
        g.z = d2 > 0.f ? gscale : (d2 < 0.f ? -gscale : 0.f);
        g.w = d3 > 0.f ? gscale : (d3 < 0.f ? -gscale : 0.f);
        reinterpret_cast<float4*>(grad)[i] = g;
    }
    if (blockIdx.x == 0)
        for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
            const float d = sr[i] - hr[i];
            acc += static_cast<double>(fabsf(d));
            grad[i] = d > 0.f ? gscale : (d < 0.f ? -gscale : 0.f);
        }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        partial[blockIdx.x] = t;
    }
}
__global__ void l1_loss_finish_kernel(const double* __restrict__ partial, int nparts, double inv_n, float* __restrict__ loss) {
    double t = 0.0;                                                    // one thread, fixed order: deterministic
    for (int i = 0; i < nparts; ++i) t += partial[i];
    loss[0] = static_cast<float>(t * inv_n);
}

// ---- conv_last input gradient: dx[b, y, x, ci] = sum_{co, ky, kx} g[b, co, y - ky + 1, x - kx + 1] * w[co, ci, ky, kx]
// thread = pixel, all Cin (<= 64) channels in registers; g is fp32 NCHW (what the loss kernel wrote), dx bf16 NHWC rows
template <int CIN>
__global__ void __launch_bounds__(128) conv_last_dgrad_kernel(const float* __restrict__ g, const float* __restrict__ w, int B, int nc, int H, int W,
                                                              __nv_bfloat16* __restrict__ dx, long long ldx) {
    __shared__ float ws[3 * 9 * CIN];                                  // [co][tap][ci]
    for (int i = threadIdx.x; i < nc * 9 * CIN; i += blockDim.x) {
        const int co = i / (9 * CIN), rem = i - co * 9 * CIN, tap = rem / CIN, ci = rem - tap * CIN;
        ws[i] = w[(co * CIN + ci) * 9 + tap];
    }
    __syncthreads();
    const long long total = static_cast<long long>(B) * H * W;
    for (long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pix < total; pix += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(pix / (H * W)), rem = static_cast<int>(pix - static_cast<long long>(b) * H * W);
        const int y = rem / W, x = rem - y * W;
        float acc[CIN];
#pragma unroll
        for (int c = 0; c < CIN; ++c) acc[c] = 0.f;
        for (int co = 0; co < nc; ++co) {
            const float* gp = g + (static_cast<long long>(b) * nc + co) * H * W;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = y - ky + 1;
                if (yy < 0 || yy >= H) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = x - kx + 1;
                    if (xx < 0 || xx >= W) continue;
                    const float gv = __ldg(gp + yy * W + xx);
                    const float* wr = ws + (co * 9 + ky * 3 + kx) * CIN;
#pragma unroll
                    for (int c = 0; c < CIN; ++c) acc[c] = fmaf(gv, wr[c], acc[c]);
                }
            }
        }
        uint4* d4 = reinterpret_cast<uint4*>(dx + pix * ldx);
#pragma unroll
        for (int c = 0; c < CIN; c += 8)
            d4[c >> 3] = make_uint4(pack_bf16x2(acc[c], acc[c + 1]), pack_bf16x2(acc[c + 2], acc[c + 3]), pack_bf16x2(acc[c + 4], acc[c + 5]),
                                    pack_bf16x2(acc[c + 6], acc[c + 7]));
    }
}

// ---- conv_last weight / bias gradient, stage 1: block = one image row; thread = (ci, quarter of the row);
// partial[block][co][ci][tap] and partial[block][nc * CIN * 9 + co] (bias)
template <int CIN>
__global__ void __launch_bounds__(4 * CIN) conv_last_wgrad_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ g, int B,
                                                                  int nc, int H, int W, float* __restrict__ partial) {
    extern __shared__ float sm[];                                      // [4][27 * CIN + 4]
    const int ci = threadIdx.x % CIN, sub = threadIdx.x / CIN;
    const int row = blockIdx.x;                                        // b * H + y
    const int b = row / H, y = row - b * H;
    float acc[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) acc[i] = 0.f;
    float bsum = 0.f;
    const int per = (W + 3) / 4, x0 = sub * per, x1 = min(W, x0 + per);
    for (int xo = x0; xo < x1; ++xo) {
        float gv[3] = {0.f, 0.f, 0.f};
        for (int co = 0; co < nc; ++co) gv[co] = __ldg(g + ((static_cast<long long>(b) * nc + co) * H + y) * W + xo);
        if (ci < nc) bsum += gv[ci];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = y + ky - 1;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = xo + kx - 1;
                if (xx < 0 || xx >= W) continue;
                const float xv = __bfloat162float(x[(static_cast<long long>(b) * H * W + static_cast<long long>(yy) * W + xx) * ldx + ci]);
#pragma unroll
                for (int co = 0; co < 3; ++co) acc[co * 9 + ky * 3 + kx] = fmaf(gv[co], xv, acc[co * 9 + ky * 3 + kx]);
            }
        }
    }
    float* mine = sm + sub * (27 * CIN + 4);
#pragma unroll
    for (int i = 0; i < 27; ++i) mine[i * CIN + ci] = acc[i];
    if (ci < 3) mine[27 * CIN + ci] = bsum;
    __syncthreads();
    float* out = partial + static_cast<long long>(row) * (27 * CIN + 4);
    for (int i = threadIdx.x; i < 27 * CIN + 4; i += blockDim.x) {
        const float t = (sm[i] + sm[27 * CIN + 4 + i]) + (sm[2 * (27 * CIN + 4) + i] + sm[3 * (27 * CIN + 4) + i]);
        out[i] = t;
    }
}
// stage 2: fixed-order sum over the rows -> dw [co][ci][3][3], db [co]
__global__ void __launch_bounds__(256) conv_last_wgrad_finish_kernel(const float* __restrict__ partial, int rows, int cin, int nc, float* __restrict__ dw,
                                                                     float* __restrict__ db) {
    const int stride = 27 * cin + 4;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= stride) return;
    double t = 0.0;
    for (int r = 0; r < rows; ++r) t += static_cast<double>(partial[static_cast<long long>(r) * stride + i]);
    if (i < 27 * cin) {
        const int co_tap = i / cin, ci = i - co_tap * cin, co = co_tap / 9, tap = co_tap - co * 9;
        if (co < nc) dw[(co * cin + ci) * 9 + tap] = static_cast<float>(t);
    } else if (i - 27 * cin < nc) {
        db[i - 27 * cin] = static_cast<float>(t);
    }
}

// ---- fused multi-tensor Adam (torch.optim.Adam: amsgrad off, maximize off); chunk table: {tensor index, first element}
struct AdamTensor { float* p; const float* g; float* m; float* v; long long n; };
__global__ void __launch_bounds__(256) adam_step_kernel(const AdamTensor* __restrict__ tensors, const int2* __restrict__ chunks, int chunk_elems,
                                                        float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt) {
    const int2 ch = chunks[blockIdx.x];
    const AdamTensor t = tensors[ch.x];
    const long long start = static_cast<long long>(ch.y) * chunk_elems;
    const long long end = start + chunk_elems < t.n ? start + chunk_elems : t.n;
    const float step_size = lr / bc1;
    for (long long i = start + threadIdx.x; i < end; i += blockDim.x) {
        float g = t.g[i];
        const float p = t.p[i];
        if (wd != 0.f) g = fmaf(wd, p, g);
        const float m = fmaf(beta1, t.m[i], (1.f - beta1) * g);
        const float v = fmaf(beta2, t.v[i], (1.f - beta2) * g * g);
        t.m[i] = m;
        t.v[i] = v;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        t.p[i] = p - step_size * (m / denom);
    }
}

}  // namespace
}  // namespace adsr

extern "C" int adsr_l1_loss_grad(const float* sr, const float* hr, int64_t n, float grad_scale, float* grad, double* partial_ws, int n_partial,
                                 float* loss, void* stream) {
    using namespace adsr;
    if (n <= 0 || sr == nullptr || hr == nullptr || grad == nullptr || partial_ws == nullptr || loss == nullptr || n_partial < 1) return ADSR_ERR_BAD_SHAPE;
    if ((reinterpret_cast<uintptr_t>(sr) | reinterpret_cast<uintptr_t>(hr) | reinterpret_cast<uintptr_t>(grad)) & 15) return ADSR_ERR_BAD_ALIGN;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    l1_loss_grad_kernel<<<n_partial, 256, 0, st>>>(sr, hr, n, grad_scale / static_cast<float>(n), grad, partial_ws);
    l1_loss_finish_kernel<<<1, 1, 0, st>>>(partial_ws, n_partial, 1.0 / static_cast<double>(n), loss);
    return cudaGetLastError() == cudaSuccess ? ADSR_OK : ADSR_ERR_LAUNCH;
}

extern "C" int64_t adsr_conv_last_bwd_workspace_bytes(int B, int H, int Cin) { return static_cast<int64_t>(B) * H * (27 * Cin + 4) * 4; }

extern "C" int adsr_conv_last_bwd(const void* x, int64_t ldx, const float* grad_out_nchw, const float* weight, int B, int H, int W, int Cin, int nc,
                                  void* dx, int64_t ld_dx, float* dw, float* db, float* workspace, void* stream) {
    using namespace adsr;
    if (B <= 0) return ADSR_OK;
    if (Cin != 64 || nc < 1 || nc > 3 || x == nullptr || grad_out_nchw == nullptr || weight == nullptr || workspace == nullptr || ldx < Cin ||
        (dx != nullptr && ((ld_dx % 8) || ld_dx < Cin)))
        return ADSR_ERR_BAD_SHAPE;
    if (dx != nullptr && (reinterpret_cast<uintptr_t>(dx) & 15)) return ADSR_ERR_BAD_ALIGN;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long total = static_cast<long long>(B) * H * W;
    if (dx != nullptr) {
        const int grid = static_cast<int>(std::min<long long>((total + 127) / 128, 148 * 32));
        conv_last_dgrad_kernel<64><<<grid, 128, 0, st>>>(grad_out_nchw, weight, B, nc, H, W, static_cast<__nv_bfloat16*>(dx), ld_dx);
    }
    if (dw != nullptr && db != nullptr) {
        const int smem = 4 * (27 * 64 + 4) * 4;
        conv_last_wgrad_kernel<64><<<B * H, 256, smem, st>>>(static_cast<const __nv_bfloat16*>(x), ldx, grad_out_nchw, B, nc, H, W, workspace);
        conv_last_wgrad_finish_kernel<<<(27 * 64 + 4 + 255) / 256, 256, 0, st>>>(workspace, B * H, 64, nc, dw, db);
    }
    return cudaGetLastError() == cudaSuccess ? ADSR_OK : ADSR_ERR_LAUNCH;
}

extern "C" int adsr_adam_step(const void* tensor_table, const void* chunk_table, int n_chunks, int chunk_elems, float lr, float beta1, float beta2,
                              float eps, float weight_decay, int step, void* stream) {
    using namespace adsr;
    if (n_chunks <= 0) return ADSR_OK;
    if (tensor_table == nullptr || chunk_table == nullptr || chunk_elems <= 0 || step < 1) return ADSR_ERR_BAD_SHAPE;
    const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, static_cast<float>(step)));
    adam_step_kernel<<<n_chunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const AdamTensor*>(tensor_table),
                                                                              static_cast<const int2*>(chunk_table), chunk_elems, lr, beta1, beta2,
                                                                              eps, weight_decay, bc1, bc2_sqrt);
    return cudaGetLastError() == cudaSuccess ? ADSR_OK : ADSR_ERR_LAUNCH;
}
