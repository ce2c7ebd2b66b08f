// Internal declarations shared by the .cu translation units (not part of the public C ABI).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/adsr_b200.h"

namespace adsr {

struct TcGemmParams {
    CUtensorMap tmap_a;  // GEMM mode: [M rows x K cols] bf16, box 128 x 64, 128-byte swizzle (first: 64 B aligned)
    int use_tma;         // 1 = A tiles arrive by TMA (GEMM), 0 = producer warps gather them (conv)
    // A operand: token rows (GEMM) or NHWC image (implicit-GEMM conv)
    const __nv_bfloat16* A;
    long long lda;       // row / pixel pitch in elements (multiple of 8)
    int M;               // output rows (tokens or output pixels)
    int K8;              // readable K columns per row (GEMM) or channels per pixel (conv), rounded up to 8
    int num_k_stages;    // 64-wide K stages streamed per tile
    int k16_total;       // GEMM: number of K=16 MMA steps
    int conv;            // 0 = GEMM, 1 = 3x3 conv
    int Hin, Win, Hout, Wout, stride;
    int stages_per_tap, k16_per_tap;
    // B operand: packed weight image
    const uint8_t* Bp;
    int n_tiles, BN, m_tiles;
    // epilogue
    const float* bias;
    int N, n_store, act;
    float slope, alpha;
    const __nv_bfloat16* res;
    long long ldres;
    __nv_bfloat16* out;
    long long ldo;
    int ocol0, out_mode;
};

int launch_tc_gemm(const TcGemmParams& p, int num_sms, cudaStream_t stream);
// Encodes the TMA descriptor of a row-major bf16 matrix (driver entry point resolved at run time).
int encode_tmap_rows_bf16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld_elems);

}  // namespace adsr
