// Per-image anomaly scores: box-filter SSIM with reflect padding for every window size of the
// evaluator's sweep, MSE and PSNR, on uint8 HWC image pairs.
//
// Replaces the reference's per-pixel Python double loop `ssim_numpy` (src/metrics.py:26-67; 0.5-1.6 s
// per call, 14 calls per image), the MSE of src/evaluate.py:259-260 and psnr_numpy (src/metrics.py:15-23).
//
// Grid = (n_ws + 1, B): CTA (j, b) computes SSIM of image b for window size ws_j; the extra CTA
// computes MSE/PSNR.  Inside a CTA one thread owns one image column and walks down the rows keeping
// the five vertical window sums (x, y, x^2, y^2, xy) as fp64 running sums (add the entering row,
// subtract the leaving one, reflect indices at the borders); per row the five column sums are turned
// into prefix sums across the row (warp shuffles + one shared-memory hop), so each horizontal window
// sum is a difference of two prefix values (three differences when the window reflects at a border).
// Work per pixel is O(1) in the window size; the images are read straight from L2 as uint8 and no
// workspace is needed.  Sums are fp64, so the result equals the vectorised oracle to ~1e-12 and the
// reference's fp32 loop to ~1e-6.
#include "adsr_kernels.h"
#include "ptx.cuh"

namespace adsr {
namespace {

constexpr int kMaxWs = 64;
struct WsList { int32_t ws[kMaxWs]; };

__device__ __forceinline__ int reflect_idx(int i, int n) {        // np.pad(mode="reflect"), |overhang| < n
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// gray value in [0,1] exactly as the reference builds it: u8 -> fp32 / 255, then dot with the fp32
// (65.738, 129.057, 25.064)/256 coefficients (src/metrics.py:37-39); single channel: value / 255.
__device__ __forceinline__ float gray01(const uint8_t* __restrict__ px, int C) {
    if (C == 1) return static_cast<float>(px[0]) / 255.0f;
    const float c0 = 65.738f / 256.0f, c1 = 129.057f / 256.0f, c2 = 25.064f / 256.0f;
    const float r = static_cast<float>(px[0]) / 255.0f, g = static_cast<float>(px[1]) / 255.0f,
                b = static_cast<float>(px[2]) / 255.0f;
    return fmaf(b, c2, fmaf(g, c1, r * c0));
}

__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }

__global__ void __launch_bounds__(1024) score_images_kernel(const uint8_t* __restrict__ sr, const uint8_t* __restrict__ hr,
                                                             int H, int W, int C, WsList wl, int n_ws,
                                                             double* __restrict__ scores) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int b = blockIdx.y, j = blockIdx.x;
    const uint8_t* srb = sr + static_cast<long long>(b) * H * W * C;
    const uint8_t* hrb = hr + static_cast<long long>(b) * H * W * C;
    double* red = sm;                                   // [32] block reduction scratch

    if (j == n_ws) {
        // ---------------- MSE / PSNR over all pixels and channels of u8/255 images ---------------
        const long long n = static_cast<long long>(H) * W * C;
        unsigned long long acc = 0;
        for (long long i = tid; i < n; i += blockDim.x) {
            const int d = static_cast<int>(srb[i]) - static_cast<int>(hrb[i]);
            acc += static_cast<unsigned long long>(d * d);
        }
        double v = static_cast<double>(acc);
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp] = v;
        __syncthreads();
        if (tid == 0) {
            double t = 0;
            for (int w = 0; w < nwarps; ++w) t += red[w];
            const double mse = t / (65025.0 * static_cast<double>(n));
            scores[static_cast<long long>(b) * (n_ws + 2) + n_ws] = mse;
            scores[static_cast<long long>(b) * (n_ws + 2) + n_ws + 1] =
                mse == 0.0 ? __longlong_as_double(0x7ff0000000000000LL) : 10.0 * log10(1.0 / mse);
        }
        return;
    }

    // ---------------- SSIM for window size ws --------------------------------------------------
    const int ws = wl.ws[j], pad = ws / 2;
    const int col = tid;                                 // one column per thread (blockDim.x >= W)
    const bool active = col < W;
    double* wtot = sm + 32;                              // [2][5][32] per-warp totals (double buffered)
    double* pref = wtot + 2 * 5 * 32;                    // [2][5][W + 1] prefix sums (double buffered)
    const double inv_n = 1.0 / (static_cast<double>(ws) * ws);
    const double C1 = 0.01 * 0.01, C2 = 0.03 * 0.03;

    double v[5] = {0, 0, 0, 0, 0};
    auto add_row = [&](int r, double sign) {
        const long long off = (static_cast<long long>(r) * W + col) * C;
        const float x = gray01(hrb + off, C), y = gray01(srb + off, C);      // x = reference (HR), y = SR
        v[0] += sign * static_cast<double>(x);
        v[1] += sign * static_cast<double>(y);
        v[2] += sign * static_cast<double>(x * x);      // products are formed in fp32 like `ref * ref`
        v[3] += sign * static_cast<double>(y * y);
        v[4] += sign * static_cast<double>(x * y);
    };
    if (active)
        for (int r = -pad; r <= pad; ++r) add_row(reflect_idx(r, H), 1.0);

    double acc = 0.0;
    for (int i = 0; i < H; ++i) {
        const int bufi = i & 1;
        if (i > 0 && active) {
            add_row(reflect_idx(i + pad, H), 1.0);
            add_row(reflect_idx(i - pad - 1, H), -1.0);
        }
        // inclusive scan of the 5 column sums across the row
        double s[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            double t = active ? v[q] : 0.0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double u = shfl_up_d(t, d);
                if (lane >= d) t += u;
            }
            s[q] = t;
            if (lane == 31) wtot[(bufi * 5 + q) * 32 + warp] = t;
        }
        __syncthreads();
        double* P = pref + bufi * 5 * (W + 1);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            double base = 0.0;
            for (int w = 0; w < warp; ++w) base += wtot[(bufi * 5 + q) * 32 + w];
            if (active) P[q * (W + 1) + col + 1] = s[q] + base;
            if (tid == 0) P[q * (W + 1)] = 0.0;
        }
        __syncthreads();
        if (active) {
            const int lo = col - pad, hi = col + pad;
            const int a0 = lo < 0 ? 0 : lo, a1 = hi > W - 1 ? W - 1 : hi;
            double box[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const double* Pq = P + q * (W + 1);
                double t = Pq[a1 + 1] - Pq[a0];
                if (lo < 0) t += Pq[-lo + 1] - Pq[1];                       // reflected columns 1 .. -lo
                if (hi > W - 1) t += Pq[W - 1] - Pq[2 * (W - 1) - hi];      // reflected columns 2(W-1)-hi .. W-2
                box[q] = t * inv_n;
            }
            const double mu1 = box[0], mu2 = box[1];
            const double s1 = box[2] - mu1 * mu1, s2 = box[3] - mu2 * mu2, s12 = box[4] - mu1 * mu2;
            acc += ((2.0 * mu1 * mu2 + C1) * (2.0 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2));
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __syncthreads();
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        double t = 0;
        for (int w = 0; w < nwarps; ++w) t += red[w];
        scores[static_cast<long long>(b) * (n_ws + 2) + j] = t / (static_cast<double>(H) * W);
    }
}

}  // namespace
}  // namespace adsr

extern "C" int adsr_score_images(const uint8_t* sr_u8_hwc, const uint8_t* hr_u8_hwc, int B, int H, int W, int C,
                                 const int32_t* host_ws_list, int n_ws, double* scores, void* stream) {
    using namespace adsr;
    if (B <= 0) return ADSR_OK;
    if (n_ws < 0 || n_ws > kMaxWs || (C != 1 && C != 3) || W > 1024 || W < 2 || H < 2) return ADSR_ERR_BAD_SHAPE;
    WsList wl;
    for (int i = 0; i < kMaxWs; ++i) wl.ws[i] = 1;
    for (int i = 0; i < n_ws; ++i) {
        const int ws = host_ws_list[i];
        if (ws < 1 || (ws % 2) == 0 || ws / 2 >= H || ws / 2 >= W) return ADSR_ERR_BAD_SHAPE;
        wl.ws[i] = ws;
    }
    const int threads = ((W + 31) / 32) * 32;
    const size_t smem = (32 + 2 * 5 * 32 + 2 * 5 * static_cast<size_t>(W + 1)) * sizeof(double);
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(score_images_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
            return ADSR_ERR_CUDA;
    }
    dim3 grid(n_ws + 1, B);
    score_images_kernel<<<grid, threads, smem, static_cast<cudaStream_t>(stream)>>>(sr_u8_hwc, hr_u8_hwc, H, W, C, wl, n_ws, scores);
    return cudaGetLastError() == cudaSuccess ? ADSR_OK : ADSR_ERR_LAUNCH;
}
