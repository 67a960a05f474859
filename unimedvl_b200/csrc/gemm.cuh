// Host-side interface of the linear-layer kernels (gemm.cu).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"

namespace umv {

enum Epilogue : int {
    EPI_BF16 = 0,      // y = bf16(acc + bias)
    EPI_GELU = 1,      // y = bf16(gelu_tanh(bf16(acc + bias)))
    EPI_SWIGLU = 2,    // weight rows interleaved [64 gate | 64 up]; y[:, j] = bf16(bf16(silu(g)) * u), g,u = bf16(acc)
    EPI_RESID = 3,     // y = bf16(bf16(acc + bias) + residual)
    EPI_PARTIAL = 4,   // fp32 split-K partials ws[split][token][feature] (weight-major only)
};

enum GemmImpl : int { GEMM_AUTO = 0, GEMM_TOKEN_MAJOR = 1, GEMM_WEIGHT_MAJOR = 2, GEMM_SIMPLE = 3 };

struct LinearCall {
    const bf16* x = nullptr;        // [M, K] activations, row stride ldx elements
    int ldx = 0;
    const bf16* w = nullptr;        // [N, K] weights (nn.Linear layout), contiguous rows
    const bf16* bias = nullptr;     // [N] or null
    const bf16* residual = nullptr; // [M, N] (EPI_RESID), row stride ldy
    bf16* y = nullptr;              // [M, N] (or [M, N/2] for EPI_SWIGLU), row stride ldy
    int ldy = 0;
    float* ws = nullptr;            // split-K workspace (EPI_PARTIAL)
    int splits = 1;
    int M = 0, N = 0, K = 0;
    int epi = EPI_BF16;
    int impl = GEMM_AUTO;
    int out_head_dim = 0, out_head_pad = 0;   // token-major EPI_BF16 only: output column n lands at (n / dim) * pad + n % dim
                                    //   (heads of `dim` columns padded to `pad`; the pad columns are never written)
    int stages = 0;                 // 0: deepest TMA ring that fits; else ring depth (shared memory left for a co-resident kernel)
    bool w_static = true;           // false: `w` is produced by the preceding kernel (activation x activation product)
    // Weight-major path only: a copy of `w` in TILE-major order -- [N / 128][K / 64] tiles of 128 rows x 64 columns, each tile 16 KB
    // contiguous (N % 128 == 0, K % 64 == 0).  A CTA's K range is then one contiguous run in HBM instead of 128-byte pieces
    // a row pitch (7-37 KB) apart: measured 7.3 TB/s for contiguous 16 KB reads against 6.6 TB/s for the strided boxes
    // (tools/micro/read_bw.cu, profiles/r2_decode_experiments.md).
    const bf16* w_tiled = nullptr;
};
// dst <- tile-major copy of the row-major [N, K] matrix src (see LinearCall::w_tiled)
int tile_weights(const bf16* src, bf16* dst, int N, int K, cudaStream_t stream);

// shared TMA helpers
int make_tmap_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
int make_tmap_3d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                 uint64_t stride2_bytes, uint32_t b1, uint32_t b2);
int gemm_sm_count();

// Token-major linear on CTA pairs (gemm_2cta.cu, tcgen05 cta_group::2): 256-row tiles, each CTA stages half of the weight tile.
bool linear_2cta_supported(const LinearCall& c);
int linear_2cta_forward(const LinearCall& c, cudaStream_t stream, int bn);      // bn: 256 or 128 (SwiGLU: 256)

// Returns 0 or a umv_status.  For EPI_PARTIAL the caller sums ws[0..splits) in a fixed order.
int linear_forward(const LinearCall& c, cudaStream_t stream);
// y = epilogue(sum of `splits` fp32 partials + bias) [+ residual] for epi in {EPI_BF16, EPI_GELU, EPI_RESID}.
// res_rows (optional): the residual of output row i is residual[res_rows[i]] (rows gathered out of a larger matrix)
int splitk_finish(const float* ws, int splits, int M, int N, const bf16* bias, const bf16* residual, bf16* y, int ldy, int epi,
                  cudaStream_t stream, const int* res_rows = nullptr);
// Heuristic split count for a weight-major (decode) GEMM: fills the 148 SMs without starving a split.
int pick_splits(int N, int K, int sm_count);
// tile-order group (row tiles of rows_a rows; column tiles of rows_b weight rows; `slots` concurrent tiles) -- gemm.cu
int pick_tile_group(int a_tiles, int b_tiles, int rows_a, int rows_b, int K, int slots);
// Must be called once before the first tcgen05 launch (resolves cuTensorMapEncodeTiled, sets smem attrs).
int gemm_init();

int trace_begin(int max_slots);       // diagnostic timeline (umv_trace_begin / umv_trace_read)
int trace_read(unsigned long long* out, char* names, int name_len, int max_slots, int* n);

extern long long g_launches;   // kernel launch counter (umv_launch_count)

}  // namespace umv
