// HBM-bound glue of the LLM / ViT forward: residual+RMSNorm, LayerNorm, q/k-norm + RoPE + paged KV
// append, embedding gathers, argmax/sampling and the device-resident decode-loop state.
// Each kernel restates the reference's bf16 rounding points; they
// replace the ~6-20 eager elementwise kernels per op of the reference.
#include "../../include/umv.h"
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace umv {

#define UMV_LAUNCH_CHECK(name)                                                        \
    do {                                                                              \
        ++g_launches;                                                                 \
        cudaError_t _e = cudaGetLastError();                                          \
        if (_e != cudaSuccess) {                                                      \
            set_error("%s launch failed: %s", name, cudaGetErrorString(_e));          \
            return UMV_ERR_CUDA;                                                      \
        }                                                                             \
    } while (0)

// ------------------------------------------------------------------------------------------
// h (+= delta) ; y = w * bf16(h * rsqrt(mean(h^2) + eps))          one CTA per row
constexpr int kNormThreads = 256;
constexpr int kNormMaxChunks = 8;   // 8 chunks * 8 elems * 256 threads = D <= 16384

// NC = 8-element chunks per thread (D <= 2048 NC): keeps the register footprint -- and with it the number of resident
// CTAs per SM on many-row (prefill / flow) calls -- proportional to the row width actually used.
template <int NC>
__global__ void __launch_bounds__(kNormThreads) add_rmsnorm_kernel(AddNormArgs a) {
    pdl_launch_dependents();
    trace_start(a.trace);
    pdl_wait();
    trace_wait(a.trace);
    __shared__ float red[32];
    const int row = blockIdx.x;
    const int nchunk = a.D / 8;
    bf16* hrow = a.h + (size_t)row * a.D;
    const bf16* w = (a.row_sel && a.row_sel[row]) ? a.w1 : a.w0;
    float v[NC][8];
    U4 wv[NC];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int ch = threadIdx.x + c * kNormThreads;
        if (ch < nchunk) {
            const U4 hv = ldg16(hrow + ch * 8);
            if (a.y) wv[c] = ldg16(w + ch * 8);
            const uint32_t* hw = &hv.x;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack2(hw[j]);
                v[c][2 * j] = f.x;
                v[c][2 * j + 1] = f.y;
            }
            if (a.delta || a.partial) {
                float d[8];
                if (a.delta) {
                    const U4 dv = ldg16(a.delta + (size_t)row * a.D + ch * 8);
                    const uint32_t* dw = &dv.x;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = unpack2(dw[j]);
                        d[2 * j] = f.x;
                        d[2 * j + 1] = f.y;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) d[j] = 0.f;
                    for (int s = 0; s < a.splits; ++s) {   // fixed order -> deterministic
                        const float4* pp = reinterpret_cast<const float4*>(a.partial + ((size_t)s * a.M + row) * a.D + ch * 8);
                        const float4 p0 = pp[0], p1 = pp[1];
                        d[0] += p0.x; d[1] += p0.y; d[2] += p0.z; d[3] += p0.w;
                        d[4] += p1.x; d[5] += p1.y; d[6] += p1.z; d[7] += p1.w;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) d[j] = rbf(d[j]);      // the Linear's own bf16 output (R3)
                }
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[c][2 * j] = rbf(v[c][2 * j] + d[2 * j]);         // residual add in bf16 (R7)
                    v[c][2 * j + 1] = rbf(v[c][2 * j + 1] + d[2 * j + 1]);
                    o[j] = pack2(v[c][2 * j], v[c][2 * j + 1]);
                }
                stg16(hrow + ch * 8, U4{o[0], o[1], o[2], o[3]});
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) ss = fmaf(v[c][j], v[c][j], ss);   // order shared with rmsnorm_rows_warp_kernel
        }
    }
    if (a.y == nullptr) return;
    ss = block_sum(ss, red);
    const float inv = 1.0f / sqrtf(ss / (float)a.D + a.eps);
    bf16* yrow = a.y + (size_t)row * a.D;
    const int slot2 = (a.y2 && a.row_slot) ? a.row_slot[row] : -1;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int ch = threadIdx.x + c * kNormThreads;
        if (ch < nchunk) {
            const uint32_t* ww = &wv[c].x;
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)      // R2: bf16(w * bf16(h * inv)); the outer product of two bf16 values is one mul.rn.bf16x2
                o[j] = bmul2(ww[j], pack2(__fmul_rn(v[c][2 * j], inv), __fmul_rn(v[c][2 * j + 1], inv)));
            stg16(yrow + ch * 8, U4{o[0], o[1], o[2], o[3]});
            if (slot2 >= 0) stg16(a.y2 + (size_t)slot2 * a.D + ch * 8, U4{o[0], o[1], o[2], o[3]});
        }
    }
    trace_end<false>(a.trace);
}

// Many-row variant without residual input (prefill / flow: the residual add already happened in the linear's epilogue):
// one WARP per row, the row held as packed bf16 in registers, warp shuffles only -- no block barrier, a quarter of the
// registers per row of the block kernel, so several times more rows in flight per SM (M=8208: 43 -> 34 us, M=16416:
// 77 -> 53 us; slower below ~4k rows, where the block kernel's 256 threads per row hide the load latency better).
// The sum of squares is taken in EXACTLY the order of add_rmsnorm_kernel<NB> (lane l plays threads 32c+l, c = 0..7: per-thread
// fma chain over its NB chunks, xor tree per "warp" c, then the xor tree over the eight warp sums), so which of the two
// kernels runs -- a function of M -- never changes a bit of the result: batch invariance stays bit-exact.
template <int NB>
__global__ void __launch_bounds__(256, NB == 1 ? 4 : 2) rmsnorm_rows_warp_kernel(AddNormArgs a) {
    pdl_launch_dependents();
    trace_start(a.trace);
    pdl_wait();
    trace_wait(a.trace);
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row < a.M) {
        const int nchunk = a.D / 8;
        const bf16* hrow = a.h + (size_t)row * a.D;
        const bf16* w = (a.row_sel && a.row_sel[row]) ? a.w1 : a.w0;
        const int slot2 = (a.y2 && a.row_slot) ? a.row_slot[row] : -1;
        U4 hv[NB][8];
#pragma unroll
        for (int n = 0; n < NB; ++n)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int ch = n * kNormThreads + c * 32 + lane;
                hv[n][c] = U4{0, 0, 0, 0};
                if (ch < nchunk) hv[n][c] = ldg16(hrow + ch * 8);
            }
        float ws[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float ss = 0.f;
#pragma unroll
            for (int n = 0; n < NB; ++n) {           // zero-filled chunks past the row end add +0: same bits as skipping them
                const uint32_t* hw = &hv[n][c].x;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack2(hw[j]);
                    ss = fmaf(f.x, f.x, ss);
                    ss = fmaf(f.y, f.y, ss);
                }
            }
            ws[c] = warp_sum(ss);
        }
#pragma unroll
        for (int n = 0; n < NB; ++n)       // opaque to the optimiser: re-unpack below instead of keeping 8 fp32 per chunk live
#pragma unroll
            for (int c = 0; c < 8; ++c)
                asm volatile("" : "+r"(hv[n][c].x), "+r"(hv[n][c].y), "+r"(hv[n][c].z), "+r"(hv[n][c].w));
        const float ss = ((ws[0] + ws[4]) + (ws[2] + ws[6])) + ((ws[1] + ws[5]) + (ws[3] + ws[7]));   // block_sum's last tree
        const float inv = 1.0f / sqrtf(ss / (float)a.D + a.eps);
        bf16* yrow = a.y + (size_t)row * a.D;
#pragma unroll
        for (int n = 0; n < NB; ++n)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int ch = n * kNormThreads + c * 32 + lane;
                if (ch < nchunk) {
                    const U4 wv = ldg16(w + ch * 8);
                    const uint32_t* hw = &hv[n][c].x;
                    const uint32_t* ww = &wv.x;
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = unpack2(hw[j]);
                        o[j] = bmul2(ww[j], pack2(__fmul_rn(f.x, inv), __fmul_rn(f.y, inv)));   // R2 (see add_rmsnorm_kernel)
                    }
                    stg16(yrow + ch * 8, U4{o[0], o[1], o[2], o[3]});
                    if (slot2 >= 0) stg16(a.y2 + (size_t)slot2 * a.D + ch * 8, U4{o[0], o[1], o[2], o[3]});
                }
            }
    }
    trace_end<false>(a.trace);
}

// Mid-size variant (256 .. 4095 rows, e.g. the 3096 rows of a text-to-image flow step): one CTA of 128 threads per row, packed
// registers.  add_rmsnorm_kernel keeps 62 registers x 256 threads per row -> 4 rows in flight per SM (ncu: occupancy limited by
// registers), 5.2 rounds for 21 rows per SM; here a row costs 128 x ~40 registers, so 12 fit.  Same summation order again: thread u
// plays the block kernel's threads u and u + 128 (virtual warps w and w + 4), then the fixed tree over the eight warp sums.
template <int NB>
__global__ void __launch_bounds__(128, 8) rmsnorm_rows_half_kernel(AddNormArgs a) {
    pdl_launch_dependents();
    trace_start(a.trace);
    pdl_wait();
    trace_wait(a.trace);
    __shared__ float red[8];
    const int row = blockIdx.x, u = threadIdx.x, warp = u >> 5, lane = u & 31;
    const int nchunk = a.D / 8;
    const bf16* hrow = a.h + (size_t)row * a.D;
    const bf16* w = (a.row_sel && a.row_sel[row]) ? a.w1 : a.w0;
    const int slot2 = (a.y2 && a.row_slot) ? a.row_slot[row] : -1;
    U4 hv[2][NB];
#pragma unroll
    for (int v = 0; v < 2; ++v)
#pragma unroll
        for (int n = 0; n < NB; ++n) {
            const int ch = n * kNormThreads + v * 128 + u;
            hv[v][n] = U4{0, 0, 0, 0};
            if (ch < nchunk) hv[v][n] = ldg16(hrow + ch * 8);
        }
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        float ss = 0.f;
#pragma unroll
        for (int n = 0; n < NB; ++n) {
            const uint32_t* hw = &hv[v][n].x;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack2(hw[j]);
                ss = fmaf(f.x, f.x, ss);
                ss = fmaf(f.y, f.y, ss);
            }
        }
        ss = warp_sum(ss);
        if (lane == 0) red[warp + 4 * v] = ss;
    }
    __syncthreads();
    const float ss = ((red[0] + red[4]) + (red[2] + red[6])) + ((red[1] + red[5]) + (red[3] + red[7]));   // block_sum's last tree
    const float inv = 1.0f / sqrtf(ss / (float)a.D + a.eps);
#pragma unroll
    for (int v = 0; v < 2; ++v)
#pragma unroll
        for (int n = 0; n < NB; ++n)
            asm volatile("" : "+r"(hv[v][n].x), "+r"(hv[v][n].y), "+r"(hv[v][n].z), "+r"(hv[v][n].w));   // stay packed
    bf16* yrow = a.y + (size_t)row * a.D;
#pragma unroll
    for (int v = 0; v < 2; ++v)
#pragma unroll
        for (int n = 0; n < NB; ++n) {
            const int ch = n * kNormThreads + v * 128 + u;
            if (ch < nchunk) {
                const U4 wv = ldg16(w + ch * 8);
                const uint32_t* hw = &hv[v][n].x;
                const uint32_t* ww = &wv.x;
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack2(hw[j]);
                    o[j] = bmul2(ww[j], pack2(__fmul_rn(f.x, inv), __fmul_rn(f.y, inv)));   // R2 (see add_rmsnorm_kernel)
                }
                stg16(yrow + ch * 8, U4{o[0], o[1], o[2], o[3]});
                if (slot2 >= 0) stg16(a.y2 + (size_t)slot2 * a.D + ch * 8, U4{o[0], o[1], o[2], o[3]});
            }
        }
    trace_end<false>(a.trace);
}

// Decode variant (split-K partials of the preceding weight-major linear, a handful of rows): the kernel sits on the
// critical path of the decode chain and is pure latency, so one thread owns one 8-element chunk and EVERY global load
// (norm weight before griddepcontrol.wait -- it is constant --, residual row and all split partials right after it) is
// issued before the first use: one memory round trip instead of one per split.  Same arithmetic and summation order as
// add_rmsnorm_kernel.
constexpr int kNormDecThreads = 512;
constexpr int kNormDecMaxSplits = 8;
__global__ void __launch_bounds__(kNormDecThreads) add_rmsnorm_splitk_kernel(AddNormArgs a) {
    pdl_launch_dependents();
    trace_start(a.trace);
    __shared__ float red[32];
    const int row = blockIdx.x, ch = threadIdx.x;
    const bool active = ch < a.D / 8;
    U4 wv = {0, 0, 0, 0};
    if (active && a.y) wv = ldg16(a.w0 + ch * 8);
    pdl_wait();
    trace_wait(a.trace);
    bf16* hrow = a.h + (size_t)row * a.D;
    float ss = 0.f;
    float v[8];
    if (active) {
        const U4 hv = ldg16(hrow + ch * 8);
        float4 p[kNormDecMaxSplits][2];
#pragma unroll
        for (int s = 0; s < kNormDecMaxSplits; ++s) {
            if (s < a.splits) {
                const float4* pp = reinterpret_cast<const float4*>(a.partial + ((size_t)s * a.M + row) * a.D + ch * 8);
                p[s][0] = pp[0];
                p[s][1] = pp[1];
            } else {
                p[s][0] = p[s][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < kNormDecMaxSplits; ++s) {      // fixed order -> deterministic; adding +0.f for s >= splits is exact
            if (s < a.splits) {
                d[0] += p[s][0].x; d[1] += p[s][0].y; d[2] += p[s][0].z; d[3] += p[s][0].w;
                d[4] += p[s][1].x; d[5] += p[s][1].y; d[6] += p[s][1].z; d[7] += p[s][1].w;
            }
        }
        const uint32_t* hw = &hv.x;
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack2(hw[j]);
            v[2 * j] = rbf(f.x + rbf(d[2 * j]));               // Linear output rounded (R3), residual add in bf16 (R7)
            v[2 * j + 1] = rbf(f.y + rbf(d[2 * j + 1]));
            o[j] = pack2(v[2 * j], v[2 * j + 1]);
        }
        stg16(hrow + ch * 8, U4{o[0], o[1], o[2], o[3]});
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += v[j] * v[j];
    }
    trace_dbg(a.trace, 0);
    if (a.y == nullptr) return;
    ss = block_sum(ss, red);
    trace_dbg(a.trace, 1);
    if (!active) return;
    const float inv = 1.0f / sqrtf(ss / (float)a.D + a.eps);
    const uint32_t* ww = &wv.x;
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 wf = unpack2(ww[j]);
        o[j] = pack2(wf.x * rbf(v[2 * j] * inv), wf.y * rbf(v[2 * j + 1] * inv));   // R2
    }
    stg16(a.y + (size_t)row * a.D + ch * 8, U4{o[0], o[1], o[2], o[3]});
    trace_end<false>(a.trace);
}

int add_rmsnorm(const AddNormArgs& a0, cudaStream_t s) {
    if (a0.M <= 0) return UMV_OK;
    AddNormArgs a = a0;
    a.trace = trace_next("add_rmsnorm");
    UMV_REQUIRE(a.D % 8 == 0 && a.D <= 8 * kNormThreads * kNormMaxChunks, UMV_ERR_UNSUPPORTED,
                "add_rmsnorm: D=%d must be a multiple of 8 and <= %d", a.D, 8 * kNormThreads * kNormMaxChunks);
    if (a.partial && !a.delta && !a.row_sel && !a.y2 && a.splits <= kNormDecMaxSplits && a.D <= 8 * kNormDecThreads && a.M <= 64) {
        launch_k(add_rmsnorm_splitk_kernel, dim3(a.M), dim3(kNormDecThreads), 0, s, a);
        UMV_LAUNCH_CHECK("add_rmsnorm_splitk_kernel");
        return UMV_OK;
    }
    // plain RMSNorm of many rows: kernel by row count (0 = block, 1 = warp per row, 2 = 128 threads per row); all three sum in the
    // same order, so the choice never changes a bit.  UMV_NORM_WARP forces one (read per call: tests switch inside one process).
    const char* warp_env = getenv("UMV_NORM_WARP");
    if (!a.delta && !a.partial && a.y && a.D / 8 <= 2 * kNormThreads) {
        const int kind = warp_env ? atoi(warp_env) : (a.M >= 4096 ? 1 : (a.M >= 256 ? 2 : 0));
        if (kind == 1) {
            const dim3 grid((a.M + 7) / 8);
            if (a.D / 8 <= kNormThreads) launch_k(rmsnorm_rows_warp_kernel<1>, grid, dim3(256), 0, s, a);
            else launch_k(rmsnorm_rows_warp_kernel<2>, grid, dim3(256), 0, s, a);
            UMV_LAUNCH_CHECK("rmsnorm_rows_warp_kernel");
            return UMV_OK;
        }
        if (kind == 2) {
            if (a.D / 8 <= kNormThreads) launch_k(rmsnorm_rows_half_kernel<1>, dim3(a.M), dim3(128), 0, s, a);
            else launch_k(rmsnorm_rows_half_kernel<2>, dim3(a.M), dim3(128), 0, s, a);
            UMV_LAUNCH_CHECK("rmsnorm_rows_half_kernel");
            return UMV_OK;
        }
    }
    const int nc = (a.D / 8 + kNormThreads - 1) / kNormThreads;
    if (nc <= 1) launch_k(add_rmsnorm_kernel<1>, dim3(a.M), dim3(kNormThreads), 0, s, a);
    else if (nc <= 2) launch_k(add_rmsnorm_kernel<2>, dim3(a.M), dim3(kNormThreads), 0, s, a);
    else if (nc <= 4) launch_k(add_rmsnorm_kernel<4>, dim3(a.M), dim3(kNormThreads), 0, s, a);
    else launch_k(add_rmsnorm_kernel<kNormMaxChunks>, dim3(a.M), dim3(kNormThreads), 0, s, a);
    UMV_LAUNCH_CHECK("add_rmsnorm_kernel");
    return UMV_OK;
}

// ------------------------------------------------------------------------------------------
// LayerNorm, fp32 statistics (two pass over registers), bf16 out.        one CTA per row
__global__ void __launch_bounds__(kNormThreads) layernorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                                  const bf16* __restrict__ b, bf16* __restrict__ y, int D,
                                                                  float eps) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[32];
    const int row = blockIdx.x;
    const int nchunk = D / 8;
    float v[kNormMaxChunks][8];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < kNormMaxChunks; ++c) {
        const int ch = threadIdx.x + c * kNormThreads;
        if (ch < nchunk) {
            const U4 hv = ldg16(x + (size_t)row * D + ch * 8);
            const uint32_t* hw = &hv.x;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack2(hw[j]);
                v[c][2 * j] = f.x;
                v[c][2 * j + 1] = f.y;
                sum += f.x + f.y;
            }
        }
    }
    const float mean = block_sum(sum, red) / (float)D;
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < kNormMaxChunks; ++c) {
        const int ch = threadIdx.x + c * kNormThreads;
        if (ch < nchunk) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float dlt = v[c][j] - mean;
                sq += dlt * dlt;
            }
        }
    }
    const float var = block_sum(sq, red) / (float)D;
    const float inv = 1.0f / sqrtf(var + eps);
#pragma unroll
    for (int c = 0; c < kNormMaxChunks; ++c) {
        const int ch = threadIdx.x + c * kNormThreads;
        if (ch < nchunk) {
            const U4 wv = ldg16(w + ch * 8), bv = ldg16(b + ch * 8);
            const uint32_t* ww = &wv.x;
            const uint32_t* bw = &bv.x;
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 wf = unpack2(ww[j]), bf = unpack2(bw[j]);
                o[j] = pack2((v[c][2 * j] - mean) * inv * wf.x + bf.x, (v[c][2 * j + 1] - mean) * inv * wf.y + bf.y);
            }
            stg16(y + (size_t)row * D + ch * 8, U4{o[0], o[1], o[2], o[3]});
        }
    }
}

// Rows up to 4096 wide (the ViT's 1152): one WARP per row, the row as packed bf16 in registers, shuffle-only statistics.  The
// block kernel above keeps 64 fp32 per thread for its widest case (107 registers -> 2 rows in flight per SM, two block
// barriers each): ~100 us for the ViT's 8192 x 1152 rows where the traffic is worth 8.  Chosen by D only, never by M.
template <int WC>
__global__ void __launch_bounds__(256) layernorm_rows_warp_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                                   const bf16* __restrict__ b, bf16* __restrict__ y, int M, int D,
                                                                   float eps, TraceSlot* trace) {
    pdl_launch_dependents();
    trace_start(trace);
    pdl_wait();
    trace_wait(trace);
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row < M) {
        const int nchunk = D / 8;
        const bf16* xrow = x + (size_t)row * D;
        U4 hv[WC];
#pragma unroll
        for (int c = 0; c < WC; ++c) {
            const int ch = lane + c * 32;
            hv[c] = U4{0, 0, 0, 0};
            if (ch < nchunk) hv[c] = ldg16(xrow + ch * 8);
        }
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < WC; ++c) {
            const uint32_t* hw = &hv[c].x;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack2(hw[j]);
                sum += f.x + f.y;
            }
        }
        const float mean = warp_sum(sum) / (float)D;
#pragma unroll
        for (int c = 0; c < WC; ++c) asm volatile("" : "+r"(hv[c].x), "+r"(hv[c].y), "+r"(hv[c].z), "+r"(hv[c].w));   // stay packed
        float sq = 0.f;
#pragma unroll
        for (int c = 0; c < WC; ++c) {
            if (lane + c * 32 < nchunk) {
                const uint32_t* hw = &hv[c].x;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack2(hw[j]);
                    const float d0 = f.x - mean, d1 = f.y - mean;
                    sq += d0 * d0;
                    sq += d1 * d1;
                }
            }
        }
        const float var = warp_sum(sq) / (float)D;
        const float inv = 1.0f / sqrtf(var + eps);
#pragma unroll
        for (int c = 0; c < WC; ++c) asm volatile("" : "+r"(hv[c].x), "+r"(hv[c].y), "+r"(hv[c].z), "+r"(hv[c].w));
#pragma unroll
        for (int c = 0; c < WC; ++c) {
            const int ch = lane + c * 32;
            if (ch < nchunk) {
                const U4 wv = ldg16(w + ch * 8), bv = ldg16(b + ch * 8);
                const uint32_t* hw = &hv[c].x;
                const uint32_t* ww = &wv.x;
                const uint32_t* bw = &bv.x;
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack2(hw[j]), wf = unpack2(ww[j]), bf = unpack2(bw[j]);
                    o[j] = pack2((f.x - mean) * inv * wf.x + bf.x, (f.y - mean) * inv * wf.y + bf.y);
                }
                stg16(y + (size_t)row * D + ch * 8, U4{o[0], o[1], o[2], o[3]});
            }
        }
    }
    trace_end<false>(trace);
}

int layernorm_bf16(const bf16* x, const bf16* w, const bf16* b, bf16* y, int M, int D, float eps, cudaStream_t s) {
    if (M <= 0) return UMV_OK;
    if (D % 8 == 0 && D / 8 <= 32 * 16) {
        const int wc = (D / 8 + 31) / 32;
        const dim3 grid((M + 7) / 8);
        TraceSlot* tr = trace_next("layernorm");
        if (wc <= 4) launch_k(layernorm_rows_warp_kernel<4>, grid, dim3(256), 0, s, x, w, b, y, M, D, eps, tr);
        else if (wc <= 8) launch_k(layernorm_rows_warp_kernel<8>, grid, dim3(256), 0, s, x, w, b, y, M, D, eps, tr);
        else launch_k(layernorm_rows_warp_kernel<16>, grid, dim3(256), 0, s, x, w, b, y, M, D, eps, tr);
        UMV_LAUNCH_CHECK("layernorm_rows_warp_kernel");
        return UMV_OK;
    }
    UMV_REQUIRE(D % 8 == 0 && D <= 8 * kNormThreads * kNormMaxChunks, UMV_ERR_UNSUPPORTED, "layernorm: bad D=%d", D);
    launch_k(layernorm_kernel, dim3(M), dim3(kNormThreads), 0, s, x, w, b, y, D, eps);
    UMV_LAUNCH_CHECK("layernorm_kernel");
    return UMV_OK;
}

// ------------------------------------------------------------------------------------------
// One warp per (row, head-slot): slots [0,H) = q heads, [H,H+Hkv) = k heads, [H+Hkv,H+2Hkv) = v heads.
// head_dim 128: lane l owns elements 4l..4l+3; the rotate_half partner (i +- 64) lives in lane l^16.
// One CTA per row: the rope angles (and their bf16-rounded cos / sin), the row's norm-weight choice and its KV slot are
// evaluated once and reused by the H + 2 Hkv head slots the CTA's warps walk over (one warp per slot, lane = 4 columns).
constexpr int kRopeWarps = 4;        // 128 threads per row: more rows resident per SM (the kernel is a latency chain per row)
__global__ void __launch_bounds__(kRopeWarps * 32, 8) rope_append_kernel(RopeAppendArgs a) {
    pdl_launch_dependents();
    trace_start(a.trace);
    pdl_wait();
    trace_wait(a.trace);
    constexpr int kBatch = 5;                        // loads of a warp's next 5 head slots in flight together
    const int row = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slots = a.H + 2 * a.Hkv;
    const int ncols = slots * a.dh;
    const int half = a.dh / 2;
    // every load that needs no other load goes out first (the kernel is a chain of a few memory round trips per row)
    uint2 raw[kBatch];
    if (!a.partial) {
#pragma unroll
        for (int i = 0; i < kBatch; ++i) {
            const int slot = warp + kRopeWarps * i;
            raw[i] = make_uint2(0u, 0u);
            if (slot < slots) raw[i] = *reinterpret_cast<const uint2*>(a.qkv + (size_t)row * ncols + slot * a.dh + lane * 4);
        }
    }
    const bool gen_row = a.row_sel && a.row_sel[row];
    const int pos_kv = a.row_kvpos[row];
    const int seq = a.row_seq[row];
    const float4 c4 = *reinterpret_cast<const float4*>(a.rope_cs + (size_t)row * a.dh + (lane * 4) % half);
    const float4 s4 = *reinterpret_cast<const float4*>(a.rope_cs + (size_t)row * a.dh + half + (lane * 4) % half);
    const float c[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
    const int page = a.page_table[(size_t)seq * a.max_pages + pos_kv / kPageTokens];
    // norm weights of this row (q heads / k heads), loaded once
    const uint2 wq2 = *reinterpret_cast<const uint2*>((gen_row ? a.qn1 : a.qn0) + lane * 4);
    const uint2 wk2 = *reinterpret_cast<const uint2*>((gen_row ? a.kn1 : a.kn0) + lane * 4);
    auto process = [&](int slot, const float (&x)[4]) {
        const int col = slot * a.dh + lane * 4;
        const bool is_q = slot < a.H;
        const bool is_v = slot >= a.H + a.Hkv;
        float o[4];
        if (is_v) {
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = x[j];
        } else {
            float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
            ss = warp_sum(ss);
            const float inv = 1.0f / sqrtf(ss / (float)a.dh + a.eps);
            const uint2 wv = is_q ? wq2 : wk2;
            const float2 w0 = unpack2(wv.x), w1 = unpack2(wv.y);
            const float w[4] = {w0.x, w0.y, w1.x, w1.y};
            float n[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (a.gen_mode) n[j] = __fmul_rn(w[j], __fmul_rn(x[j], inv));          // fp32 throughout
                else n[j] = rbf(w[j] * rbf(x[j] * inv));                               // R4: two bf16 roundings
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float partner = __shfl_xor_sync(0xffffffffu, n[j], 16);
                const float rot = (lane < 16) ? -partner : partner;
                if (a.gen_mode) o[j] = rbf(__fadd_rn(__fmul_rn(n[j], c[j]), __fmul_rn(rot, sn[j])));
                else o[j] = rbf(rbf(n[j] * c[j]) + rbf(rot * sn[j]));                   // R5: three roundings
            }
        }
        const uint2 packed = make_uint2(pack2(o[0], o[1]), pack2(o[2], o[3]));
        if (is_q) {
            *reinterpret_cast<uint2*>(a.q_out + (size_t)(a.q_row_map ? a.q_row_map[row] : row) * a.ldq + col) = packed;
        } else {
            const int kv = is_v ? 1 : 0;
            const int head = slot - a.H - (is_v ? a.Hkv : 0);
            bf16* dst = a.pool.base + a.pool.tile_offset(page, a.layer, kv, head) + (size_t)(pos_kv % kPageTokens) * a.dh + lane * 4;
            *reinterpret_cast<uint2*>(dst) = packed;
        }
    };
    if (a.partial) {
        for (int slot = warp; slot < slots; slot += kRopeWarps) {
            const int col = slot * a.dh + lane * 4;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int s = 0; s < a.splits; ++s) {
                const float4 p = *reinterpret_cast<const float4*>(a.partial + ((size_t)s * a.M + row) * ncols + col);
                acc[0] += p.x; acc[1] += p.y; acc[2] += p.z; acc[3] += p.w;
            }
            const uint2 bv = *reinterpret_cast<const uint2*>(a.bias + col);
            const float2 b0 = unpack2(bv.x), b1 = unpack2(bv.y);
            const float x[4] = {rbf(acc[0] + b0.x), rbf(acc[1] + b0.y), rbf(acc[2] + b1.x), rbf(acc[3] + b1.y)};
            process(slot, x);
        }
    } else {
        for (int s0 = warp; s0 < slots; s0 += kRopeWarps * kBatch) {
            if (s0 != warp) {
#pragma unroll
                for (int i = 0; i < kBatch; ++i) {
                    const int slot = s0 + kRopeWarps * i;
                    if (slot < slots) raw[i] = *reinterpret_cast<const uint2*>(a.qkv + (size_t)row * ncols + slot * a.dh + lane * 4);
                }
            }
#pragma unroll
            for (int i = 0; i < kBatch; ++i) {
                const int slot = s0 + kRopeWarps * i;
                if (slot < slots) {
                    const float2 f0 = unpack2(raw[i].x), f1 = unpack2(raw[i].y);
                    const float x[4] = {f0.x, f0.y, f1.x, f1.y};
                    process(slot, x);
                }
            }
        }
    }
    trace_end<false>(a.trace);
}

// Many-row variant (prefill / flow; input = the q|k|v linear's bf16 output): ncu showed rope_append_kernel issue-bound at
// 8208 rows (73 % issue slots, 185 warp instructions per head slot), so this one spends a quarter of the instructions: a lane
// owns 8 columns (16-byte loads / stores), a warp works on TWO head slots at once (16 lanes each; the rotate_half partner is
// lane ^ 8), and the understanding-mode rounding chain (R4, R5) runs on packed bf16x2: for bf16 operands
// mul.rn.bf16x2 == bf16(fp32 product) and add.rn.bf16x2 == bf16(fp32 sum) (both fp32 results are exact before the
// rounding).  The sum of squares follows rope_append_kernel's order (old lane 2j | 2j+1 = the two halves of new lane j, then
// the same xor tree), so both kernels -- and the fused decode attention -- produce the same bits.
constexpr int kRopePairs = 5;        // slot pairs per warp kept in flight (4 warps x 5 pairs x 2 = 40 >= H + 2 Hkv = 36)
template <bool GEN>
__global__ void __launch_bounds__(kRopeWarps * 32, 8) rope_append_rows_kernel(RopeAppendArgs a) {
    pdl_launch_dependents();
    trace_start(a.trace);
    pdl_wait();
    trace_wait(a.trace);
    const int row = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane & 15, hsel = lane >> 4;
    const int slots = a.H + 2 * a.Hkv;
    const int ncols = slots * 128;
    const bf16* qrow = a.qkv + (size_t)row * ncols + sub * 8;
    for (int p0 = warp; 2 * p0 < slots; p0 += kRopeWarps * kRopePairs) {
        U4 raw[kRopePairs];
#pragma unroll
        for (int i = 0; i < kRopePairs; ++i) {
            const int slot = 2 * (p0 + kRopeWarps * i) + hsel;
            raw[i] = U4{0, 0, 0, 0};
            if (slot < slots) raw[i] = ldg16(qrow + slot * 128);
        }
        const bool gen_row = a.row_sel && a.row_sel[row];
        const int pos_kv = a.row_kvpos[row];
        const int seq = a.row_seq[row];
        const float* csrow = a.rope_cs + (size_t)row * 128 + (sub & 7) * 8;
        const float4 c0 = *reinterpret_cast<const float4*>(csrow), c1 = *reinterpret_cast<const float4*>(csrow + 4);
        const float4 s0 = *reinterpret_cast<const float4*>(csrow + 64), s1 = *reinterpret_cast<const float4*>(csrow + 68);
        const U4 wq = ldg16((gen_row ? a.qn1 : a.qn0) + sub * 8);
        const U4 wk = ldg16((gen_row ? a.kn1 : a.kn0) + sub * 8);
        const int page = a.page_table[(size_t)seq * a.max_pages + pos_kv / kPageTokens];
        const float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        uint32_t c2[4], s2[4];                          // the table holds bf16-rounded values: packing is exact
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            c2[k] = pack2(c[2 * k], c[2 * k + 1]);
            s2[k] = pack2(sn[2 * k], sn[2 * k + 1]);
        }
        const uint32_t sign = (sub < 8) ? 0x80008000u : 0u;        // rotate_half: -x[i + 64] for i < 64, +x[i - 64] above
#pragma unroll
        for (int i = 0; i < kRopePairs; ++i) {
            const int slot = 2 * (p0 + kRopeWarps * i) + hsel;
            if (2 * (p0 + kRopeWarps * i) >= slots) break;          // warp-uniform
            const bool valid = slot < slots;
            const bool is_q = slot < a.H;
            const bool is_v = slot >= a.H + a.Hkv;
            const uint32_t* rw = &raw[i].x;
            float x[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = unpack2(rw[k]);
                x[2 * k] = f.x;
                x[2 * k + 1] = f.y;
            }
            float sa = fmaf(x[3], x[3], fmaf(x[2], x[2], fmaf(x[1], x[1], __fmul_rn(x[0], x[0]))));
            float sb = fmaf(x[7], x[7], fmaf(x[6], x[6], fmaf(x[5], x[5], __fmul_rn(x[4], x[4]))));
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                sa += __shfl_xor_sync(0xffffffffu, sa, o);
                sb += __shfl_xor_sync(0xffffffffu, sb, o);
            }
            const float inv = 1.0f / sqrtf((sa + sb) * (1.0f / 128.0f) + a.eps);
            const U4 wv = is_q ? wq : wk;
            const uint32_t* ww = &wv.x;
            uint32_t o2[4];
            if constexpr (GEN) {                                                         // fp32 throughout, one rounding
                float n[8];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 wf = unpack2(ww[k]);
                    n[2 * k] = __fmul_rn(wf.x, __fmul_rn(x[2 * k], inv));
                    n[2 * k + 1] = __fmul_rn(wf.y, __fmul_rn(x[2 * k + 1], inv));
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float r[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float partner = __shfl_xor_sync(0xffffffffu, n[2 * k + e], 8);
                        const float rot = (sub < 8) ? -partner : partner;
                        r[e] = __fadd_rn(__fmul_rn(n[2 * k + e], c[2 * k + e]), __fmul_rn(rot, sn[2 * k + e]));
                    }
                    o2[k] = pack2(r[0], r[1]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t t2 = pack2(__fmul_rn(x[2 * k], inv), __fmul_rn(x[2 * k + 1], inv));   // R4: bf16(x * inv)
                    const uint32_t n2 = bmul2(ww[k], t2);                                               //     bf16(w * .)
                    const uint32_t rot2 = __shfl_xor_sync(0xffffffffu, n2, 8) ^ sign;
                    o2[k] = badd2(bmul2(n2, c2[k]), bmul2(rot2, s2[k]));                                // R5: three roundings
                }
            }
            if (!valid) continue;
            const U4 outv = is_v ? raw[i] : U4{o2[0], o2[1], o2[2], o2[3]};
            if (is_q) {
                stg16(a.q_out + (size_t)(a.q_row_map ? a.q_row_map[row] : row) * a.ldq + slot * 128 + sub * 8, outv);
            } else {
                const int kv = is_v ? 1 : 0;
                const int head = slot - a.H - (is_v ? a.Hkv : 0);
                stg16(a.pool.base + a.pool.tile_offset(page, a.layer, kv, head) + (size_t)(pos_kv % kPageTokens) * 128 + sub * 8, outv);
            }
        }
    }
    trace_end<false>(a.trace);
}

__global__ void rope_table_kernel(const int* __restrict__ positions, const float* __restrict__ inv_freq, int M, int dh,
                                  float* __restrict__ cs) {
    pdl_launch_dependents();
    pdl_wait();
    const int half = dh / 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * half) return;
    const int m = idx / half, i = idx % half;
    const float ang = __fmul_rn((float)positions[m], inv_freq[i]);
    cs[(size_t)m * dh + i] = rbf(cosf(ang));                       // cos/sin cast to bf16 (R5)
    cs[(size_t)m * dh + half + i] = rbf(sinf(ang));
}
int rope_table(const int* positions, const float* inv_freq, int M, int dh, float* cs, cudaStream_t s) {
    if (M <= 0) return UMV_OK;
    launch_k(rope_table_kernel, dim3((M * (dh / 2) + 255) / 256), dim3(256), 0, s, positions, inv_freq, M, dh, cs);
    UMV_LAUNCH_CHECK("rope_table_kernel");
    return UMV_OK;
}

int rope_append(const RopeAppendArgs& a0, cudaStream_t s) {
    if (a0.M <= 0) return UMV_OK;
    RopeAppendArgs a = a0;
    a.trace = trace_next("rope_append");
    UMV_REQUIRE(a.dh == 128, UMV_ERR_UNSUPPORTED, "rope_append: head_dim %d (only 128 is built)", a.dh);
    UMV_REQUIRE(a.rope_cs != nullptr, UMV_ERR_INVALID, "rope_append: the per-forward cos/sin table (rope_table) is required");
    const int blocks = a.M;
    const char* rows_env = getenv("UMV_ROPE_ROWS");        // read per call: tests switch paths inside one process
    if (!a.partial && !(rows_env && atoi(rows_env) == 0)) {
        if (a.gen_mode) launch_k(rope_append_rows_kernel<true>, dim3(blocks), dim3(kRopeWarps * 32), 0, s, a);
        else launch_k(rope_append_rows_kernel<false>, dim3(blocks), dim3(kRopeWarps * 32), 0, s, a);
        UMV_LAUNCH_CHECK("rope_append_rows_kernel");
        return UMV_OK;
    }
    launch_k(rope_append_kernel, dim3(blocks), dim3(kRopeWarps * 32), 0, s, a);
    UMV_LAUNCH_CHECK("rope_append_kernel");
    return UMV_OK;
}

// ------------------------------------------------------------------------------------------
__global__ void embed_rows_kernel(const bf16* __restrict__ table, const int64_t* __restrict__ ids, int D, int64_t vocab,
                                  bf16* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x;
    int64_t id = ids[row];
    if (id < 0 || id >= vocab) id = 0;      // host validates ids; never read out of bounds
    const U4* src = reinterpret_cast<const U4*>(table + (size_t)id * D);
    U4* dst = reinterpret_cast<U4*>(out + (size_t)row * D);
    for (int c = threadIdx.x; c < D / 8; c += blockDim.x) dst[c] = src[c];
}
int embed_rows(const bf16* table, const int64_t* ids, int n, int D, int64_t vocab, bf16* out, cudaStream_t s) {
    if (n <= 0) return UMV_OK;
    launch_k(embed_rows_kernel, dim3(n), dim3(128), 0, s, table, ids, D, vocab, out);
    UMV_LAUNCH_CHECK("embed_rows_kernel");
    return UMV_OK;
}

// dst[dst_rows[i]] = table[ids[i]]: marker / text-token embeddings written straight into the packed query sequence
// (packed_sequence[packed_text_indexes] = embed_tokens(packed_text_ids), bagel.py:577-579).
__global__ void embed_rows_scatter_kernel(const bf16* __restrict__ table, const int64_t* __restrict__ ids, const int* __restrict__ dst_rows,
                                          int D, int64_t vocab, bf16* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    int64_t id = ids[blockIdx.x];
    if (id < 0 || id >= vocab) id = 0;
    const U4* src = reinterpret_cast<const U4*>(table + (size_t)id * D);
    U4* dst = reinterpret_cast<U4*>(out + (size_t)dst_rows[blockIdx.x] * D);
    for (int c = threadIdx.x; c < D / 8; c += blockDim.x) dst[c] = src[c];
}
int embed_rows_scatter(const bf16* table, const int64_t* ids, const int* dst_rows, int n, int D, int64_t vocab, bf16* out, cudaStream_t s) {
    if (n <= 0) return UMV_OK;
    launch_k(embed_rows_scatter_kernel, dim3(n), dim3(128), 0, s, table, ids, dst_rows, D, vocab, out);
    UMV_LAUNCH_CHECK("embed_rows_scatter_kernel");
    return UMV_OK;
}

// dst[dst_rows[m]] = bf16(src[m] + table[ids[m]]): the frozen 2-D position embedding added to the connector output and the
// result scattered to its rows of the packed query sequence in one pass (bagel.py:590-595).
__global__ void scatter_add_rows_kernel(const bf16* __restrict__ src, const bf16* __restrict__ table, const int64_t* __restrict__ ids,
                                        const int* __restrict__ dst_rows, bf16* __restrict__ dst, int D) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x;
    const bf16* t = table + (size_t)ids[row] * D;
    const bf16* xr = src + (size_t)row * D;
    bf16* yr = dst + (size_t)dst_rows[row] * D;
    for (int c = threadIdx.x; c < D / 8; c += blockDim.x) {
        const U4 a = ldg16(xr + c * 8), b = ldg16(t + c * 8);
        const uint32_t* aw = &a.x;
        const uint32_t* bw = &b.x;
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 fa = unpack2(aw[j]), fb = unpack2(bw[j]);
            o[j] = pack2(fa.x + fb.x, fa.y + fb.y);
        }
        stg16(yr + c * 8, U4{o[0], o[1], o[2], o[3]});
    }
}
int scatter_add_rows(const bf16* src, const bf16* table, const int64_t* ids, const int* dst_rows, bf16* dst, int M, int D,
                     cudaStream_t s) {
    if (M <= 0) return UMV_OK;
    launch_k(scatter_add_rows_kernel, dim3(M), dim3(128), 0, s, src, table, ids, dst_rows, dst, D);
    UMV_LAUNCH_CHECK("scatter_add_rows_kernel");
    return UMV_OK;
}

__global__ void gather_add_rows_kernel(bf16* __restrict__ x, const bf16* __restrict__ table, const int64_t* __restrict__ ids,
                                       int D) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x;
    const bf16* t = table + (size_t)ids[row] * D;
    bf16* xr = x + (size_t)row * D;
    for (int c = threadIdx.x; c < D / 8; c += blockDim.x) {
        const U4 a = ldg16(xr + c * 8), b = ldg16(t + c * 8);
        const uint32_t* aw = &a.x;
        const uint32_t* bw = &b.x;
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 fa = unpack2(aw[j]), fb = unpack2(bw[j]);
            o[j] = pack2(fa.x + fb.x, fa.y + fb.y);
        }
        stg16(xr + c * 8, U4{o[0], o[1], o[2], o[3]});
    }
}
int gather_add_rows(bf16* x, const bf16* table, const int64_t* ids, int M, int D, cudaStream_t s) {
    if (M <= 0) return UMV_OK;
    launch_k(gather_add_rows_kernel, dim3(M), dim3(128), 0, s, x, table, ids, D);
    UMV_LAUNCH_CHECK("gather_add_rows_kernel");
    return UMV_OK;
}

// ToTensor + Normalize(0.5, 0.5) + patchify (data/transforms.py:104-115, data_utils.py:43-58): the same fp32 operations in the
// same order as torch (u / 255, - 0.5, / 0.5), so the patch vectors are bit-identical to the host path.
__global__ void patchify_u8_kernel(const uint8_t* __restrict__ img, int W, int patch, int max_per_side, float* __restrict__ out,
                                   int64_t* __restrict__ pos_ids) {
    pdl_launch_dependents();
    pdl_wait();
    const int pr = blockIdx.y, pc = blockIdx.x, cols = gridDim.x;
    const int token = pr * cols + pc;
    const int n = patch * patch * 3;
    float* dst = out + (size_t)token * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = i % 3, q = (i / 3) % patch, p = i / (3 * patch);          // (p_row, p_col, channel)
        const uint8_t u = img[((size_t)(pr * patch + p) * W + (pc * patch + q)) * 3 + c];
        const float x = __fdiv_rn((float)u, 255.0f);
        dst[i] = __fdiv_rn(__fsub_rn(x, 0.5f), 0.5f);
    }
    if (threadIdx.x == 0) pos_ids[token] = (int64_t)pr * max_per_side + pc;
}
int patchify_u8(const uint8_t* img, int H, int W, int patch, int max_per_side, float* out, int64_t* pos_ids, cudaStream_t s) {
    launch_k(patchify_u8_kernel, dim3(W / patch, H / patch), dim3(128), 0, s, img, W, patch, max_per_side, out, pos_ids);
    UMV_LAUNCH_CHECK("patchify_u8_kernel");
    return UMV_OK;
}

__global__ void f32_to_bf16_padded_kernel(const float* __restrict__ x, bf16* __restrict__ y, int K, int Kpad) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x;
    for (int c = threadIdx.x; c < Kpad; c += blockDim.x)
        y[(size_t)row * Kpad + c] = c < K ? f2b(x[(size_t)row * K + c]) : f2b(0.f);
}
int f32_to_bf16_padded(const float* x, bf16* y, int M, int K, int Kpad, cudaStream_t s) {
    if (M <= 0) return UMV_OK;
    launch_k(f32_to_bf16_padded_kernel, dim3(M), dim3(128), 0, s, x, y, K, Kpad);
    UMV_LAUNCH_CHECK("f32_to_bf16_padded_kernel");
    return UMV_OK;
}

// rows[]: scatter==0: dst[i] = src[rows[i]] ; scatter==1: dst[rows[i]] = src[i]
__global__ void copy_rows_kernel(const bf16* __restrict__ src, int lds, const int* __restrict__ rows, bf16* __restrict__ dst,
                                 int ldd, int D, int scatter, TraceSlot* trace) {
    pdl_launch_dependents();
    trace_start(trace);
    pdl_wait();
    trace_wait(trace);
    const int i = blockIdx.x;
    const int r = rows[i];
    const bf16* s = src + (size_t)(scatter ? i : r) * lds;
    bf16* d = dst + (size_t)(scatter ? r : i) * ldd;
    for (int c = threadIdx.x; c < D / 8; c += blockDim.x) stg16(d + c * 8, ldg16(s + c * 8));
    trace_end<false>(trace);
}
int copy_rows(const bf16* src, int lds, const int* rows, bf16* dst, int ldd, int n, int D, int scatter, cudaStream_t s) {
    if (n <= 0) return UMV_OK;
    launch_k(copy_rows_kernel, dim3(n), dim3(128), 0, s, src, lds, rows, dst, ldd, D, scatter, trace_next("copy_rows"));
    UMV_LAUNCH_CHECK("copy_rows_kernel");
    return UMV_OK;
}

// ------------------------------------------------------------------------------------------
// torch.argmax over bf16 logits: maximum value, ties -> lowest index (R8).  One CTA per row.
// `end` (decode loop only): the row's CTA also advances that sample's loop state -- what decode_end_kernel does -- saving a
// launch on the per-step critical path (nothing later in the step reads positions / kv_len / step).
__global__ void __launch_bounds__(1024) argmax_kernel(const bf16* __restrict__ logits, int vocab, int64_t* __restrict__ out,
                                                       DecodeState end, int advance) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sval[32];
    __shared__ int sidx[32];
    const bf16* row = logits + (size_t)blockIdx.x * vocab;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    const int nchunk = vocab / 8;
    constexpr int kAhead = 5;                       // a thread's loads of 5 strides in flight together (19 chunks per thread at 152k)
    for (int c0 = threadIdx.x; c0 < nchunk; c0 += kAhead * blockDim.x) {
        U4 v[kAhead];
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {
            const int c = c0 + u * blockDim.x;
            if (c < nchunk) v[u] = ldg16_stream(row + c * 8);
        }
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {
            const int c = c0 + u * blockDim.x;
            if (c < nchunk) {
                const uint32_t* w = &v[u].x;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack2(w[j]);
                    const int i0 = c * 8 + 2 * j;
                    if (f.x > best || (f.x == best && i0 < bi)) { best = f.x; bi = i0; }
                    if (f.y > best || (f.y == best && i0 + 1 < bi)) { best = f.y; bi = i0 + 1; }
                }
            }
        }
    }
    for (int i = nchunk * 8 + threadIdx.x; i < vocab; i += blockDim.x) {
        const float f = b2f(row[i]);
        if (f > best || (f == best && i < bi)) { best = f; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sval[wid] = best; sidx[wid] = bi; }
    __syncthreads();
    if (wid == 0) {
        const int nw = blockDim.x >> 5;
        best = lane < nw ? sval[lane] : -INFINITY;
        bi = lane < nw ? sidx[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) {
            out[blockIdx.x] = (bi == 0x7fffffff) ? 0 : bi;   // all-NaN row -> 0
            if (advance) {
                const int b = blockIdx.x;
                end.positions[b] += 1;     // packed_query_position_ids + 1 (bagel.py:1310)
                end.kv_len[b] += 1;        // key_values_lens + 1 (bagel.py:1309)
                end.row_kvpos[b] += 1;
                if (b == 0) *end.step += 1;
            }
        }
    }
}
int argmax_rows(const bf16* logits, int rows, int vocab, int64_t* out, cudaStream_t s, const DecodeState* end) {
    if (rows <= 0) return UMV_OK;
    UMV_REQUIRE((reinterpret_cast<uintptr_t>(logits) & 15) == 0 && vocab % 8 == 0, UMV_ERR_INVALID,
                "argmax: logits must be 16-byte aligned with vocab %% 8 == 0 (vocab=%d)", vocab);
    launch_k(argmax_kernel, dim3(rows), dim3(1024), 0, s, logits, vocab, out, end ? *end : DecodeState{}, end ? 1 : 0);
    UMV_LAUNCH_CHECK("argmax_kernel");
    return UMV_OK;
}

// softmax(logits / T) sampling (bagel.py:1298-1299) with a counter-based hash RNG: u ~ U[0,1) from
// (seed, step, row); inverse-CDF over the fp32 probabilities.  One CTA per row.
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// `u_force` < 0: draw u from the hash; otherwise use it as u (test hook for the u*total == total rounding edge).
__global__ void __launch_bounds__(1024) sample_kernel(const bf16* __restrict__ logits, int vocab, float inv_t, uint64_t seed,
                                                       const int* __restrict__ step, int64_t* __restrict__ out, float u_force) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[32];
    __shared__ float s_prefix[1024];
    const bf16* row = logits + (size_t)blockIdx.x * vocab;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < vocab; i += blockDim.x) mx = fmaxf(mx, rbf(b2f(row[i]) * inv_t));
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    // each thread owns a contiguous span so the CDF is in index order
    const int span = (vocab + blockDim.x - 1) / blockDim.x;
    const int i0 = threadIdx.x * span, i1 = min(vocab, i0 + span);
    float local = 0.f;
    for (int i = i0; i < i1; ++i) local += expf(rbf(b2f(row[i]) * inv_t) - mx);
    s_prefix[threadIdx.x] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float run = 0.f;
        for (int t = 0; t < blockDim.x; ++t) { const float v = s_prefix[t]; s_prefix[t] = run; run += v; }
        red[0] = run;
    }
    __syncthreads();
    const float total = red[0];
    const uint64_t h = splitmix64(seed ^ splitmix64(((uint64_t)(step ? *step : 0) << 32) | blockIdx.x));
    const float u = u_force >= 0.f ? u_force : (float)((h >> 40) * (1.0 / 16777216.0));
    const float target = u * total;
    const float before = s_prefix[threadIdx.x];
    // u < 1, but u * total may round UP to total in fp32: then no span satisfies target < before + local and the LAST NON-EMPTY
    // span must take it (with span = ceil(vocab / 1024) the trailing threads own nothing: 149 x 1021 >= 152064)
    const int last = (vocab - 1) / span;
    if (target >= before && (target < before + local || threadIdx.x == last) && i0 < i1) {
        float run = before;
        int pick = i1 - 1;
        for (int i = i0; i < i1; ++i) {
            run += expf(rbf(b2f(row[i]) * inv_t) - mx);
            if (target < run) { pick = i; break; }
        }
        out[blockIdx.x] = pick;
    }
}
int sample_rows(const bf16* logits, int rows, int vocab, float temperature, uint64_t seed, const int* step, int64_t* out,
                cudaStream_t s, float u_force) {
    if (rows <= 0) return UMV_OK;
    // logits / T on a bf16 CUDA tensor with a Python scalar is x * (1.0f / T) in fp32, rounded to bf16 (ATen's div-by-scalar)
    launch_k(sample_kernel, dim3(rows), dim3(1024), 0, s, logits, vocab, 1.0f / temperature, seed, step, out, u_force);
    UMV_LAUNCH_CHECK("sample_kernel");
    return UMV_OK;
}

// ------------------------------------------------------------------------------------------
// Decode-loop state (device resident so a captured step can be replayed without host work).
__global__ void decode_begin_kernel(const bf16* __restrict__ table, int D, int64_t vocab, DecodeState st,
                                    const int64_t* __restrict__ forced, int64_t* __restrict__ tokens_out, int B,
                                    bf16* __restrict__ x) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.x;
    const int step = *st.step;
    int64_t tok = forced ? forced[(size_t)step * B + b] : st.cur_tokens[b];
    if (threadIdx.x == 0) tokens_out[(size_t)step * B + b] = tok;     // generated_sequence.append(curr_tokens)
    if (tok < 0 || tok >= vocab) tok = 0;
    const U4* src = reinterpret_cast<const U4*>(table + (size_t)tok * D);
    U4* dst = reinterpret_cast<U4*>(x + (size_t)b * D);
    for (int c = threadIdx.x; c < D / 8; c += blockDim.x) dst[c] = src[c];
    // rope angles of this step (modeling_qwen2.py:164-181): the same for every layer, so they are evaluated once here
    if (st.rope_cs && threadIdx.x < st.dh / 2) {
        const float ang = __fmul_rn((float)st.positions[b], st.inv_freq[threadIdx.x]);
        st.rope_cs[(size_t)b * st.dh + threadIdx.x] = rbf(cosf(ang));                  // cos/sin cast to bf16 (R5)
        st.rope_cs[(size_t)b * st.dh + st.dh / 2 + threadIdx.x] = rbf(sinf(ang));
    }
}
int decode_begin_step(const bf16* table, int D, int64_t vocab, DecodeState st, const int64_t* forced, int64_t* tokens_out,
                      int B, bf16* x, cudaStream_t s) {
    launch_k(decode_begin_kernel, dim3(B), dim3(128), 0, s, table, D, vocab, st, forced, tokens_out, B, x);
    UMV_LAUNCH_CHECK("decode_begin_kernel");
    return UMV_OK;
}
__global__ void decode_end_kernel(DecodeState st, int B) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = threadIdx.x;
    if (b < B) {
        st.positions[b] += 1;      // packed_query_position_ids + 1 (bagel.py:1310)
        st.kv_len[b] += 1;         // key_values_lens + 1 (bagel.py:1309)
        st.row_kvpos[b] += 1;
    }
    if (b == 0) *st.step += 1;
}
int decode_end_step(DecodeState st, int B, cudaStream_t s) {
    launch_k(decode_end_kernel, dim3(1), dim3(64), 0, s, st, B);
    UMV_LAUNCH_CHECK("decode_end_kernel");
    return UMV_OK;
}

// ------------------------------------------------------------------------------------------
__global__ void export_kv_kernel(KVPool pool, int layer, const int* __restrict__ pages, int len, bf16* __restrict__ k_out,
                                 bf16* __restrict__ v_out) {
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x, head = blockIdx.y;
    const int page = pages[t / kPageTokens], slot = t % kPageTokens;
    const bf16* ks = pool.base + pool.tile_offset(page, layer, 0, head) + (size_t)slot * pool.head_dim;
    const bf16* vs = pool.base + pool.tile_offset(page, layer, 1, head) + (size_t)slot * pool.head_dim;
    const size_t o = ((size_t)t * pool.kv_heads + head) * pool.head_dim;
    for (int i = threadIdx.x; i < pool.head_dim; i += blockDim.x) {
        k_out[o + i] = ks[i];
        v_out[o + i] = vs[i];
    }
}
int export_kv(KVPool pool, int layer, const int* pages, int len, bf16* k_out, bf16* v_out, cudaStream_t s) {
    if (len <= 0) return UMV_OK;
    launch_k(export_kv_kernel, dim3(dim3(len, pool.kv_heads)), dim3(128), 0, s, pool, layer, pages, len, k_out, v_out);
    UMV_LAUNCH_CHECK("export_kv_kernel");
    return UMV_OK;
}

__global__ void copy_page_kernel(KVPool pool, int src_page, int dst_page) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t per = pool.page_layer_elems() / 8;            // 16-byte chunks of one (layer, page) block
    const size_t n = (size_t)pool.layers * per;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int layer = (int)(i / per);
        const size_t off = (i % per) * 8;
        *reinterpret_cast<U4*>(pool.base + pool.tile_offset(dst_page, layer, 0, 0) + off) =
            *reinterpret_cast<const U4*>(pool.base + pool.tile_offset(src_page, layer, 0, 0) + off);
    }
}
int copy_page(KVPool pool, int src_page, int dst_page, cudaStream_t s) {
    launch_k(copy_page_kernel, dim3(148), dim3(256), 0, s, pool, src_page, dst_page);
    UMV_LAUNCH_CHECK("copy_page_kernel");
    return UMV_OK;
}

// Synthetic weights: uniform(mean - bound, mean + bound) from a counter hash (benchmark random init).
__global__ void fill_uniform_kernel(bf16* __restrict__ p, size_t n, uint64_t seed, float bound, float mean) {
    pdl_launch_dependents();
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t h = splitmix64(seed + i);
        const float u = (float)(h >> 40) * (1.0f / 16777216.0f);
        p[i] = f2b(mean + bound * (2.f * u - 1.f));
    }
}
int fill_uniform_bf16(bf16* p, size_t n, uint64_t seed, float bound, float mean, cudaStream_t s) {
    if (n == 0) return UMV_OK;
    launch_k(fill_uniform_kernel, dim3(148 * 8), dim3(256), 0, s, p, n, seed, bound, mean);
    UMV_LAUNCH_CHECK("fill_uniform_kernel");
    return UMV_OK;
}

}  // namespace umv
