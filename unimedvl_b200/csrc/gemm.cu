// Linear layers (y = x W^T + b with fused epilogues) for sm_100a.
//
// One persistent, warp-specialised kernel: TMA (cp.async.bulk.tensor, 128-byte swizzle) stages
// 128 x 64 / BN x 64 bf16 operand tiles through a multi-stage mbarrier ring, a single elected thread
// issues tcgen05.mma (M=128, N=BN, K=16, kind::f16, fp32 accumulators in TMEM, double buffered) and
// four epilogue warps drain TMEM with tcgen05.ld and apply the reference's rounding points.
//
//   token-major  (prefill, M large): A = activations [M,K], B = weights [N,K]; TMEM lane = token,
//                column = output feature.  Tensor-core bound.
//   weight-major (decode,  M <= 64): A = weights [N,K] streamed once from HBM (evict-first),
//                B = activations; TMEM lane = output feature, column = token.  HBM bound; optional
//                split-K writes fp32 partials that the consumer kernel reduces in a fixed order.
//
// Replaces the reference's F.linear -> cuBLAS call sites:
// qwen2_navit.py:541-543,555-562,617-620; modeling_qwen2.py:234-235; bagel.py:1295;
// siglip_navit.py:190,216-218,243,256-258; modeling_utils.py:108,120-122.
#include <cuda.h>
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/umv.h"
#include "common.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace umv {

long long g_launches = 0;
bool g_pdl = true;

// ---- timeline trace state (one trace per process; slots are handed out in launch order)
static TraceSlot* g_trace_dev = nullptr;
static int g_trace_cap = 0, g_trace_n = 0;
static std::vector<std::string> g_trace_names;
bool trace_active() { return g_trace_dev != nullptr; }
TraceSlot* trace_next(const char* kernel_name) {
    if (!g_trace_dev || g_trace_n >= g_trace_cap) return nullptr;
    g_trace_names.emplace_back(kernel_name);
    return g_trace_dev + g_trace_n++;
}
int trace_begin(int max_slots) {
    if (g_trace_dev) cudaFree(g_trace_dev);
    g_trace_dev = nullptr;
    g_trace_cap = g_trace_n = 0;
    g_trace_names.clear();
    if (max_slots <= 0) return UMV_OK;
    UMV_CUDA_OK(cudaMalloc(&g_trace_dev, sizeof(TraceSlot) * max_slots));
    UMV_CUDA_OK(cudaMemset(g_trace_dev, 0, sizeof(TraceSlot) * max_slots));
    g_trace_cap = max_slots;
    return UMV_OK;
}
int trace_read(unsigned long long* out, char* names, int name_len, int max_slots, int* n) {
    UMV_CUDA_OK(cudaDeviceSynchronize());
    const int k = g_trace_n < max_slots ? g_trace_n : max_slots;
    if (k > 0) UMV_CUDA_OK(cudaMemcpy(out, g_trace_dev, sizeof(TraceSlot) * k, cudaMemcpyDeviceToHost));
    for (int i = 0; i < k; ++i) {
        strncpy(names + (size_t)i * name_len, g_trace_names[i].c_str(), name_len - 1);
        names[(size_t)i * name_len + name_len - 1] = 0;
    }
    *n = k;
    return UMV_OK;
}

constexpr int BM = 128;   // UMMA M (TMEM lanes)
constexpr int BK = 64;    // 64 bf16 = one 128-byte swizzle row
constexpr int kATile = BM * BK * 2;

struct GemmParams {
    int a_rows, b_rows, K;
    int a_tiles, b_tiles, splits, kb_total, kb_per_split;
    int tokens, features;
    bf16* y;
    int ldy;
    const bf16* bias;
    const bf16* residual;
    float* ws;
    int epi;
    int early_a;    // weight-major: A (weights) is constant data and may be fetched before griddepcontrol.wait
    TraceSlot* trace;
    int hd, hp;     // output head padding (LinearCall::out_head_dim / out_head_pad), 0 = off
    int stages;     // ring depth of this launch (<= TcCfg::kStages); a shallow ring leaves shared memory for a neighbour CTA
    int a_tiled;    // weight-major: the A tensor map views a tile-major weight copy as [tiles * 128 rows, 64 columns]
    int group_a;    // token-major tile order: column tiles are swept inside groups of group_a row tiles (see gemm_2cta.cu: tile_coords)
};

template <int BN, bool SWAP>
struct TcCfg {
    static constexpr int kBTile = BN * BK * 2;
    static constexpr int kStageBytes = kATile + kBTile;
    static constexpr int kStagesRaw = (196 * 1024) / kStageBytes;
    static constexpr int kStages = kStagesRaw > 10 ? 10 : kStagesRaw;
    static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int kBarBytes = 256;
    static constexpr int kExchBytes = SWAP ? 64 * BN * 4 : 0;   // swiglu gate/up exchange (weight-major only)
    static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kBarBytes + kExchBytes;
    static constexpr int smem_bytes(int stages) { return 1024 + stages * kStageBytes + kBarBytes + kExchBytes; }
};

__device__ __forceinline__ float epi_value(int epi, float acc, float bias) {
    float v = rbf(acc + bias);
    if (epi == EPI_GELU) v = rbf(gelu_tanh_f(v));
    return v;
}

// MODE 0: bf16 / gelu / +residual (runtime p.epi), 1: fp32 split-K partials, 2: swiglu
// Threads: warp 0 TMA producer, warp 1 MMA issuer, then the epilogue warps: 4 on the weight-major path (its epilogue is
// a few columns), 8 on the token-major path -- two warps per TMEM lane quarter, each draining half of the tile's columns,
// so that short-K tiles (ViT K = 1152, VAE) are not paced by the epilogue.
template <bool SWAP>
constexpr int gemm_threads() { return SWAP ? 192 : 320; }

// (row tile, column tile) of tile index t2.  Weight-major: weight tiles fastest.  Token-major: row tiles fastest inside groups of
// p.group_a row tiles, column tiles next, groups last -- the grouped order of gemm_2cta.cu (tile_coords); results do not depend on it.
template <bool SWAP>
__device__ __forceinline__ void tile_ab(int t2, const GemmParams& p, int& a_tile, int& b_tile) {
    if (SWAP || p.group_a >= p.a_tiles) {
        a_tile = t2 % p.a_tiles;
        b_tile = t2 / p.a_tiles;
        return;
    }
    const int per_group = p.group_a * p.b_tiles;
    const int g = t2 / per_group;
    const int first = g * p.group_a;
    const int gsz = min(p.group_a, p.a_tiles - first);
    const int r = t2 - g * per_group;
    a_tile = first + r % gsz;
    b_tile = r / gsz;
}

template <int BN, int MODE, bool SWAP>
__global__ void __launch_bounds__(gemm_threads<SWAP>(), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    using Cfg = TcCfg<BN, SWAP>;
    constexpr int kMaxStages = Cfg::kStages;
    const int kStages = p.stages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sB = smem + kStages * kATile;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
    uint64_t* empty = full + kMaxStages;
    uint64_t* tfull = empty + kMaxStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* sExch = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes + Cfg::kBarBytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    trace_start(p.trace);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < kStages; ++i) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&tfull[i], 1);
                mbar_init(&tempty[i], SWAP ? 4 : 8);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_tiles = p.a_tiles * p.b_tiles * p.splits;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (elect_one()) {
            const uint64_t hintA = SWAP ? kEvictFirst : kEvictNormal;
            const uint64_t hintB = SWAP ? kEvictLast : kEvictNormal;
            int stage = 0;
            uint32_t phase = 0;
            // Weight-major: the A operand (weights) does not depend on the previous kernel, so the first
            // kStages weight tiles are requested BEFORE griddepcontrol.wait and stream in while the
            // predecessor (a norm / rope / attention kernel on a handful of SMs) is still running; the
            // activation tiles of those stages are requested right after the wait.
            bool waited = !SWAP || !p.early_a;
            int n_deferred = 0;
            int def_kb[kMaxStages], def_bt[kMaxStages];
            if (waited) pdl_wait();
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int split = tile % p.splits;
                const int t2 = tile / p.splits;
                int a_tile, b_tile;
                tile_ab<SWAP>(t2, p, a_tile, b_tile);
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (!waited && n_deferred == kStages) {
                        pdl_wait();
                        waited = true;
                        for (int i = 0; i < n_deferred; ++i)
                            tma_load_2d(sB + i * Cfg::kBTile, &tmB, &full[i], def_kb[i] * BK, def_bt[i] * BN, hintB);
                    }
                    mbar_wait(&empty[stage], phase ^ 1u);
                    mbar_expect_tx(&full[stage], Cfg::kStageBytes);
                    if (SWAP && p.a_tiled) tma_load_2d(sA + stage * kATile, &tmA, &full[stage], 0, (a_tile * p.kb_total + kb) * BM, hintA);
                    else tma_load_2d(sA + stage * kATile, &tmA, &full[stage], kb * BK, a_tile * BM, hintA);
                    if (waited) {
                        tma_load_2d(sB + stage * Cfg::kBTile, &tmB, &full[stage], kb * BK, b_tile * BN, hintB);
                    } else {
                        def_kb[n_deferred] = kb;          // stage index == n_deferred during the first ring pass
                        def_bt[n_deferred] = b_tile;
                        ++n_deferred;
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
            if (p.trace && blockIdx.x == 0) p.trace->t_wait = gtime();     // all tiles requested (short jobs: == after the wait)
            if (!waited) {
                pdl_wait();
                for (int i = 0; i < n_deferred; ++i)
                    tma_load_2d(sB + i * Cfg::kBTile, &tmB, &full[i], def_kb[i] * BK, def_bt[i] * BN, hintB);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(BN);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int split = tile % p.splits;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                mbar_wait(&tempty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * kATile);
                    const uint32_t b_addr = smem_u32(sB + stage * Cfg::kBTile);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        umma_bf16(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                                  (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(&tfull[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (4 or 8 warps)
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may read
        const int chalf = SWAP ? 0 : (warp - 2) >> 2;  // token-major: which half of the tile's columns this warp drains
        const int lrow = quarter * 32 + lane;          // lane within the 128-row tile
        pdl_wait();                                    // outputs / bias / residual belong to the dependency chain
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int split = tile % p.splits;
            const int t2 = tile / p.splits;
            int a_tile, b_tile;
            tile_ab<SWAP>(t2, p, a_tile, b_tile);
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;

            if constexpr (!SWAP && MODE == 0) {
                const int row = a_tile * BM + lrow;            // token
                const bool row_ok = row < p.tokens;
#pragma unroll 1
                for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(taddr + c0, r);
                    tmem_ld_wait();
                    const int n0 = b_tile * BN + c0;
                    if (row_ok && n0 < p.features) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int n = n0 + h * 8;
                            if (n < p.features) {
                                U4 bv = {0, 0, 0, 0}, rv = {0, 0, 0, 0};
                                if (p.bias) bv = ldg16(p.bias + n);
                                if (p.epi == EPI_RESID) rv = ldg16(p.residual + (size_t)row * p.ldy + n);
                                const uint32_t* bw = &bv.x;
                                const uint32_t* rw = &rv.x;
                                uint32_t o[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float2 b2 = unpack2(bw[j]);
                                    float v0 = epi_value(p.epi, __uint_as_float(r[h * 8 + 2 * j]), b2.x);
                                    float v1 = epi_value(p.epi, __uint_as_float(r[h * 8 + 2 * j + 1]), b2.y);
                                    if (p.epi == EPI_RESID) {
                                        const float2 r2 = unpack2(rw[j]);
                                        v0 += r2.x;
                                        v1 += r2.y;
                                    }
                                    o[j] = pack2(v0, v1);
                                }
                                const int nd = p.hd ? (n / p.hd) * p.hp + n % p.hd : n;
                                stg16(p.y + (size_t)row * p.ldy + nd, U4{o[0], o[1], o[2], o[3]});
                            }
                        }
                    }
                }
            } else if constexpr (!SWAP && MODE == 2) {
                // weight rows interleaved [64 gate | 64 up] -> tile columns [blk*128 + c] / [blk*128 + 64 + c]
                const int row = a_tile * BM + lrow;
                const bool row_ok = row < p.tokens;
                constexpr int kUnits = (BN / 128) * 4;             // 16-column gate/up chunk pairs of the tile
#pragma unroll 1
                for (int unit = chalf * (kUnits / 2); unit < (chalf + 1) * (kUnits / 2); ++unit) {
                    const int blk = unit >> 2, c0 = (unit & 3) * 16;
                    {
                        uint32_t g[16], u[16];
                        tmem_ld16(taddr + blk * 128 + c0, g);
                        tmem_ld16(taddr + blk * 128 + 64 + c0, u);
                        tmem_ld_wait();
                        const int j0 = (b_tile * BN) / 2 + blk * 64 + c0;       // activation column
                        if (row_ok && j0 < p.features / 2) {
                            uint32_t o[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float a0 = rbf(silu_f(rbf(__uint_as_float(g[2 * j])))) * rbf(__uint_as_float(u[2 * j]));
                                float a1 = rbf(silu_f(rbf(__uint_as_float(g[2 * j + 1])))) * rbf(__uint_as_float(u[2 * j + 1]));
                                o[j] = pack2(a0, a1);
                            }
                            bf16* dst = p.y + (size_t)row * p.ldy + j0;
                            stg16(dst, U4{o[0], o[1], o[2], o[3]});
                            stg16(dst + 8, U4{o[4], o[5], o[6], o[7]});
                        }
                    }
                }
            } else if constexpr (SWAP && MODE != 2) {
                const int f = a_tile * BM + lrow;              // output feature
                const bool f_ok = f < p.features;
                float bias = 0.f;
                if (MODE == 0 && p.bias && f_ok) bias = b2f(p.bias[f]);
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(taddr + c0, r);
                    tmem_ld_wait();
                    if (f_ok) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int t = b_tile * BN + c0 + j;
                            if (t < p.tokens) {
                                const float a = __uint_as_float(r[j]);
                                if constexpr (MODE == 1) {
                                    p.ws[((size_t)split * p.tokens + t) * p.features + f] = a;
                                } else {
                                    float v = epi_value(p.epi, a, bias);
                                    if (p.epi == EPI_RESID) v += b2f(p.residual[(size_t)t * p.ldy + f]);
                                    p.y[(size_t)t * p.ldy + f] = f2b(v);
                                }
                            }
                        }
                    }
                }
            } else {
                // SWAP + swiglu: lanes 0..63 hold gate rows, 64..127 the matching up rows.
                const bool is_up = lrow >= 64;
                const int jl = lrow & 63;
                const int jglob = a_tile * 64 + jl;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(taddr + c0, r);
                    tmem_ld_wait();
                    if (is_up) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) sExch[(c0 + j) * 64 + jl] = rbf(__uint_as_float(r[j]));
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (!is_up && jglob < p.features / 2) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int t = b_tile * BN + c0 + j;
                            if (t < p.tokens) {
                                const float gv = rbf(silu_f(rbf(__uint_as_float(r[j]))));
                                p.y[(size_t)t * p.ldy + jglob] = f2b(gv * sExch[(c0 + j) * 64 + jl]);
                            }
                        }
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    trace_end(p.trace);
}

// ---------------------------------------------------------------------------------------------
// Plain CUDA-core kernel: same contract, used for shapes TMA cannot address (row stride not a
// multiple of 16 bytes) and as the in-library cross-check of the tcgen05 path (impl = 3).
__global__ void gemm_simple_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ w,
                                   const bf16* __restrict__ bias, const bf16* __restrict__ residual,
                                   bf16* __restrict__ y, int ldy, int M, int N, int K, int epi) {
    const int n_out = (epi == EPI_SWIGLU) ? N / 2 : N;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y * blockDim.y + threadIdx.y;
    pdl_launch_dependents();
    pdl_wait();
    if (m >= M || n >= n_out) return;
    const bf16* xr = x + (size_t)m * ldx;
    if (epi == EPI_SWIGLU) {
        const bf16* wg = w + (size_t)((n / 64) * 128 + (n % 64)) * K;
        const bf16* wu = wg + (size_t)64 * K;
        float g = 0.f, u = 0.f;
        for (int k = 0; k < K; ++k) {
            const float xv = b2f(xr[k]);
            g = fmaf(xv, b2f(wg[k]), g);
            u = fmaf(xv, b2f(wu[k]), u);
        }
        y[(size_t)m * ldy + n] = f2b(rbf(silu_f(rbf(g))) * rbf(u));
        return;
    }
    const bf16* wr = w + (size_t)n * K;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(b2f(xr[k]), b2f(wr[k]), acc);
    float v = epi_value(epi, acc, bias ? b2f(bias[n]) : 0.f);
    if (epi == EPI_RESID) v += b2f(residual[(size_t)m * ldy + n]);
    y[(size_t)m * ldy + n] = f2b(v);
}

// ---------------------------------------------------------------------------------------------
// Finish of a split-K weight-major linear whose consumer is not fused with the reduction (the few understanding-expert
// rows of a generation-mode forward, the time-embedding MLP): y = epilogue(sum_s ws[s] + bias) [+ residual], partials
// summed in split order (deterministic), every partial load of a thread in flight together.  One thread per 8 features.
constexpr int kFinishMaxSplits = 16;
__global__ void __launch_bounds__(256) splitk_finish_kernel(const float* __restrict__ ws, int splits, int M, int N,
                                                            const bf16* __restrict__ bias, const bf16* __restrict__ residual,
                                                            bf16* __restrict__ y, int ldy, int epi, const int* __restrict__ res_rows,
                                                            TraceSlot* trace) {
    pdl_launch_dependents();
    trace_start(trace);
    pdl_wait();
    trace_wait(trace);
    const int nch = N / 8;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < M * nch) {
        const int row = idx / nch, n = (idx % nch) * 8;
        float4 p[kFinishMaxSplits][2];
#pragma unroll
        for (int sp = 0; sp < kFinishMaxSplits; ++sp) {
            if (sp < splits) {
                const float4* pp = reinterpret_cast<const float4*>(ws + ((size_t)sp * M + row) * N + n);
                p[sp][0] = pp[0];
                p[sp][1] = pp[1];
            }
        }
        U4 bv = {0, 0, 0, 0}, rv = {0, 0, 0, 0};
        if (bias) bv = ldg16(bias + n);
        if (epi == EPI_RESID) rv = ldg16(residual + (size_t)(res_rows ? res_rows[row] : row) * ldy + n);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int sp = 0; sp < kFinishMaxSplits; ++sp) {
            if (sp < splits) {
                acc[0] += p[sp][0].x; acc[1] += p[sp][0].y; acc[2] += p[sp][0].z; acc[3] += p[sp][0].w;
                acc[4] += p[sp][1].x; acc[5] += p[sp][1].y; acc[6] += p[sp][1].z; acc[7] += p[sp][1].w;
            }
        }
        const uint32_t* bw = &bv.x;
        const uint32_t* rw = &rv.x;
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 b2 = unpack2(bw[j]);
            float v0 = epi_value(epi, acc[2 * j], b2.x), v1 = epi_value(epi, acc[2 * j + 1], b2.y);
            if (epi == EPI_RESID) {
                const float2 r2 = unpack2(rw[j]);
                v0 += r2.x;
                v1 += r2.y;
            }
            o[j] = pack2(v0, v1);
        }
        stg16(y + (size_t)row * ldy + n, U4{o[0], o[1], o[2], o[3]});
    }
    trace_end<false>(trace);
}

int splitk_finish(const float* ws, int splits, int M, int N, const bf16* bias, const bf16* residual, bf16* y, int ldy, int epi,
                  cudaStream_t stream, const int* res_rows) {
    UMV_REQUIRE(splits >= 1 && splits <= kFinishMaxSplits && N % 8 == 0 && ldy % 8 == 0, UMV_ERR_INVALID,
                "splitk_finish: splits=%d N=%d ldy=%d", splits, N, ldy);
    UMV_REQUIRE(epi == EPI_BF16 || epi == EPI_GELU || epi == EPI_RESID, UMV_ERR_INVALID, "splitk_finish: epilogue %d", epi);
    const int total = M * (N / 8);
    cudaError_t e = launch_k(splitk_finish_kernel, dim3((total + 255) / 256), dim3(256), 0, stream, ws, splits, M, N, bias, residual, y,
                             ldy, epi, res_rows, trace_next("splitk_finish"));
    ++g_launches;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("splitk_finish_kernel launch failed: %s", cudaGetErrorString(e));
        return UMV_ERR_CUDA;
    }
    return UMV_OK;
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_sm_count = 148;
static std::once_flag g_init_once;
static int g_init_status = 0;

template <int BN, int MODE, bool SWAP>
static void set_smem_attr() {
    cudaFuncSetAttribute(gemm_tc_kernel<BN, MODE, SWAP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         TcCfg<BN, SWAP>::kSmemBytes);
}

int gemm_init() {
    std::call_once(g_init_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled unavailable (%s): a CUDA 12 driver and an sm_100a GPU are required",
                      cudaGetErrorString(e));
            g_init_status = UMV_ERR_CUDA;
            return;
        }
        g_encode = reinterpret_cast<EncodeTiledFn>(fn);
        if (const char* v = getenv("UMV_PDL")) g_pdl = atoi(v) != 0;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        set_smem_attr<256, 0, false>(); set_smem_attr<128, 0, false>(); set_smem_attr<64, 0, false>();
        set_smem_attr<256, 2, false>(); set_smem_attr<128, 2, false>();
        set_smem_attr<16, 0, true>(); set_smem_attr<32, 0, true>(); set_smem_attr<64, 0, true>();
        set_smem_attr<16, 1, true>(); set_smem_attr<32, 1, true>(); set_smem_attr<64, 1, true>();
        set_smem_attr<16, 2, true>(); set_smem_attr<32, 2, true>(); set_smem_attr<64, 2, true>();
        cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) {
            set_error("cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString(le));
            g_init_status = UMV_ERR_CUDA;
        }
    });
    return g_init_status;
}

int gemm_sm_count() { return g_sm_count; }
int make_tmap_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
static int make_tmap(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    return make_tmap_2d(m, ptr, rows, cols, ld, box_rows);
}
int make_tmap_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) ptr=%p rows=%llu cols=%llu ld=%llu box=%u", (int)r, ptr,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows);
        return UMV_ERR_CUDA;
    }
    return UMV_OK;
}

// 3-D bf16 map with 128-byte swizzle: dims (innermost first) {d0, d1, d2}, byte strides of d1 / d2, box {64, b1, b2}.
int make_tmap_3d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                 uint64_t stride2_bytes, uint32_t b1, uint32_t b2) {
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {BK, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(3d) failed (%d) ptr=%p dims=%llu,%llu,%llu strides=%llu,%llu box=64,%u,%u", (int)r, ptr,
                  (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)stride1_bytes,
                  (unsigned long long)stride2_bytes, b1, b2);
        return UMV_ERR_CUDA;
    }
    return UMV_OK;
}

int pick_splits(int N, int K, int sm_count) {
    const int a_tiles = (N + BM - 1) / BM;
    const int kb_total = (K + BK - 1) / BK;
    if (a_tiles >= sm_count) return 1;
    int s = (sm_count + a_tiles / 2) / a_tiles;          // ~one wave of CTAs
    const int max_by_k = kb_total / 8 > 0 ? kb_total / 8 : 1;   // keep >= 8 k-blocks (128 KB of weights) per split
    if (s > max_by_k) s = max_by_k;
    if (s > 16) s = 16;
    // 112 CTAs already stream at the full HBM rate (measured), and every extra split is another fp32 partial the consumer
    // has to read on the critical path: do not go beyond 4 splits once 4 of them occupy three quarters of the SMs
    if (s > 4 && 4 * a_tiles * 4 >= 3 * sm_count) s = 4;
    if (s < 1) s = 1;
    const int per = (kb_total + s - 1) / s;
    return (kb_total + per - 1) / per;        // effective count: no empty split
}

// Row tiles per tile-order group of a token-major linear: the candidate with the least estimated DRAM traffic.  Per group the weights
// are read once (every column tile is visited) and the group's activation rows once per wave of the group, unless they are small
// enough to stay in L2 next to the streaming weights and outputs (a quarter of the 126 MB; measured: the prefill's 59 MB of gate/up
// activations are already re-read by a third per wave).  The estimate is crude (the 8,208-row down_proj went from 2.5 to 1.9 GB of DRAM
// reads where it predicts 1.2), so the plain all-rows order stays unless a quarter of the traffic is at stake (at 3,072 rows the
// candidates are within +-1 % of each other in time, profiles/r2_tile_order.md).  UMV_RASTER_G forces a value (>= tiles: plain order).
int pick_tile_group(int a_tiles, int b_tiles, int rows_a, int rows_b, int K, int slots) {
    const char* env = getenv("UMV_RASTER_G");               // read per call: experiments switch inside one process
    if (env && atoi(env) > 0) return std::min(atoi(env), a_tiles);
    const double act_tile = (double)rows_a * K * 2, w_all = (double)b_tiles * rows_b * K * 2, l2_keep = 32e6;
    auto estimate = [&](int g) {
        double traffic = 0;
        for (int first = 0; first < a_tiles; first += g) {
            const int gsz = std::min(g, a_tiles - first);
            const double act = gsz * act_tile;
            const int waves = (gsz * b_tiles + slots - 1) / slots;
            traffic += w_all + act * (act <= l2_keep ? 1 : waves);
        }
        return traffic;
    };
    double best = 1e300;
    int best_g = a_tiles;
    for (int g = 1; g <= a_tiles; ++g) {
        const double t = estimate(g);
        if (t <= best) { best = t; best_g = g; }                 // ties: the larger group (fewer weight passes)
    }
    return best < 0.75 * estimate(a_tiles) ? best_g : a_tiles;
}

template <int BN, int MODE, bool SWAP>
static int launch_tc(const LinearCall& c, cudaStream_t stream) {
    GemmParams p{};
    const bf16* A = SWAP ? c.w : c.x;
    const bf16* B = SWAP ? c.x : c.w;
    p.a_rows = SWAP ? c.N : c.M;
    p.b_rows = SWAP ? c.M : c.N;
    const int lda = SWAP ? c.K : c.ldx, ldb = SWAP ? c.ldx : c.K;
    p.K = c.K;
    p.a_tiles = (p.a_rows + BM - 1) / BM;
    p.b_tiles = (p.b_rows + BN - 1) / BN;
    p.kb_total = (c.K + BK - 1) / BK;
    int splits = (MODE == 1) ? (c.splits < 1 ? 1 : c.splits) : 1;
    p.kb_per_split = (p.kb_total + splits - 1) / splits;
    p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
    if (MODE == 1 && p.splits != splits) {
        set_error("linear: split count %d leaves empty splits for K=%d (use %d)", splits, c.K, p.splits);
        return UMV_ERR_INVALID;
    }
    p.tokens = c.M;
    p.features = c.N;
    p.y = c.y;
    p.ldy = c.ldy;
    p.bias = c.bias;
    p.residual = c.residual;
    p.ws = c.ws;
    p.epi = c.epi;
    p.early_a = c.w_static ? 1 : 0;
    p.group_a = SWAP ? p.a_tiles : pick_tile_group(p.a_tiles, p.b_tiles, BM, BN, c.K, g_sm_count);
    p.hd = c.out_head_dim; p.hp = c.out_head_pad;
    if (p.hd && (SWAP || MODE != 0 || c.epi != EPI_BF16 || p.hd % 8 != 0 || p.hp % 8 != 0)) {
        set_error("linear: output head padding exists only on the token-major bf16 epilogue with 8-column aligned heads");
        return UMV_ERR_UNSUPPORTED;
    }
    {
        char nm[32];
        snprintf(nm, sizeof nm, "gemm<%d,%d,%d> N%d K%d", BN, MODE, (int)SWAP, c.N, c.K);
        p.trace = trace_next(nm);
    }
    p.stages = (c.stages > 0 && c.stages < TcCfg<BN, SWAP>::kStages) ? (c.stages < 2 ? 2 : c.stages) : TcCfg<BN, SWAP>::kStages;
    CUtensorMap tmA, tmB;
    int rc;
    if (SWAP && c.w_tiled && c.N % BM == 0 && c.K % BK == 0) {
        p.a_tiled = 1;
        rc = make_tmap(&tmA, c.w_tiled, (uint64_t)p.a_tiles * p.kb_total * BM, BK, BK, BM);
    } else {
        rc = make_tmap(&tmA, A, p.a_rows, c.K, lda, BM);
    }
    if (rc) return rc;
    rc = make_tmap(&tmB, B, p.b_rows, c.K, ldb, BN);
    if (rc) return rc;
    const int tiles = p.a_tiles * p.b_tiles * p.splits;
    const int grid = tiles < g_sm_count ? tiles : g_sm_count;
    cudaError_t e = launch_k(gemm_tc_kernel<BN, MODE, SWAP>, dim3(grid), dim3(gemm_threads<SWAP>()), TcCfg<BN, SWAP>::smem_bytes(p.stages), stream, tmA, tmB, p);
    ++g_launches;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("gemm_tc_kernel<%d,%d,%d> launch failed: %s", BN, MODE, (int)SWAP, cudaGetErrorString(e));
        return UMV_ERR_CUDA;
    }
    return UMV_OK;
}

// [N, K] row-major -> [N/128][K/64][128][64]: one thread per 16-byte chunk
__global__ void tile_weights_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int N, int K) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;          // destination chunk
    const size_t total = (size_t)N * K / 8;
    if (i >= total) return;
    const int c8 = (int)(i % 8), r = (int)((i / 8) % BM);
    const size_t tile = i / (8 * BM);
    const int KT = K / BK;
    const int kt = (int)(tile % KT), nt = (int)(tile / KT);
    *reinterpret_cast<U4*>(dst + i * 8) = *reinterpret_cast<const U4*>(src + ((size_t)nt * BM + r) * K + (size_t)kt * BK + c8 * 8);
}
int tile_weights(const bf16* src, bf16* dst, int N, int K, cudaStream_t stream) {
    UMV_REQUIRE(N % BM == 0 && K % BK == 0, UMV_ERR_INVALID, "tile_weights: N=%d K=%d", N, K);
    const size_t total = (size_t)N * K / 8;
    tile_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src, dst, N, K);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("tile_weights_kernel launch failed: %s", cudaGetErrorString(e));
        return UMV_ERR_CUDA;
    }
    return UMV_OK;
}

int linear_forward(const LinearCall& c, cudaStream_t stream) {
    if (c.M <= 0 || c.N <= 0 || c.K <= 0) return UMV_OK;
    int impl = c.impl;
    const bool tma_ok = (c.K % 8 == 0) && (c.ldx % 8 == 0) && ((reinterpret_cast<uintptr_t>(c.x) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(c.w) & 15) == 0);
    if (impl == GEMM_AUTO) {
        if (!tma_ok) impl = GEMM_SIMPLE;
        else if (c.M <= 64) impl = GEMM_WEIGHT_MAJOR;
        else impl = GEMM_TOKEN_MAJOR;
    }
    if (c.epi == EPI_SWIGLU && (c.N % 128 != 0)) {
        set_error("linear: swiglu epilogue needs N %% 128 == 0 (got %d)", c.N);
        return UMV_ERR_INVALID;
    }
    if (impl == GEMM_TOKEN_MAJOR && (c.N % 8 != 0 || c.ldy % 8 != 0)) impl = GEMM_SIMPLE;
    if (impl == GEMM_WEIGHT_MAJOR && c.M > 64) impl = GEMM_TOKEN_MAJOR;
    if (impl != GEMM_SIMPLE) {
        if (!tma_ok) {
            set_error("linear: tcgen05 path needs 16-byte aligned rows (K=%d ldx=%d)", c.K, c.ldx);
            return UMV_ERR_INVALID;
        }
        int rc = gemm_init();
        if (rc) return rc;
    }
    if (c.out_head_dim && impl != GEMM_TOKEN_MAJOR) {
        set_error("linear: output head padding needs the token-major path (M=%d N=%d K=%d)", c.M, c.N, c.K);
        return UMV_ERR_UNSUPPORTED;
    }
    if (c.epi == EPI_PARTIAL && impl != GEMM_WEIGHT_MAJOR) {
        set_error("linear: split-K partial epilogue exists only on the weight-major path");
        return UMV_ERR_INVALID;
    }
    if (impl == GEMM_SIMPLE) {
        dim3 block(32, 8);
        const int n_out = c.epi == EPI_SWIGLU ? c.N / 2 : c.N;
        dim3 grid((n_out + 31) / 32, (c.M + 7) / 8);
        cudaError_t e = launch_k(gemm_simple_kernel, grid, block, 0, stream, c.x, c.ldx, c.w, c.bias, c.residual, c.y, c.ldy,
                                 c.M, c.N, c.K, c.epi);
        ++g_launches;
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            set_error("gemm_simple_kernel launch failed: %s", cudaGetErrorString(e));
            return UMV_ERR_CUDA;
        }
        return UMV_OK;
    }
    if (impl == GEMM_TOKEN_MAJOR) {
        // Tile shape: single CTA (128 x BN) or CTA pair (256 x BN, gemm_2cta.cu), BN 256 or 128, by a cost model in units of
        // one 128 x 256 single-CTA tile time: waves of the persistent grid x per-tile factor (measured on B200 at the 14B
        // shapes, tools/gemm_bench.py: pair tiles 0.90; 128-wide tiles 0.54 single / 0.47 pair, plus up to 0.15 when the
        // activation matrix no longer fits L2 and every extra column of tiles re-reads it from HBM).
        const char* pair_env = getenv("UMV_2CTA");              // read per call: tests switch paths inside one process
        const bool pair_ok = !(pair_env && atoi(pair_env) == 0) && linear_2cta_supported(c);
        static const int force_bn = getenv("UMV_BN") ? atoi(getenv("UMV_BN")) : 0;
        const long m1 = (c.M + BM - 1) / BM, m2 = (c.M + 2 * BM - 1) / (2 * BM);
        const long n256 = (c.N + 255) / 256, n128 = (c.N + 127) / 128;
        const long sms = g_sm_count, pairs = g_sm_count / 2;
        auto waves = [](long tiles, long slots) { return (double)((tiles + slots - 1) / slots); };
        const double spill = std::min(1.0, (double)c.M * c.K * 2 / 300e6) * 0.15;
        if (c.epi == EPI_SWIGLU) {
            const bool wide = c.N % 256 == 0;
            if (pair_ok && wide && waves(m2 * n256, pairs) * 0.90 < waves(m1 * n256, sms)) return linear_2cta_forward(c, stream, 256);
            return wide ? launch_tc<256, 2, false>(c, stream) : launch_tc<128, 2, false>(c, stream);
        }
        if (c.N > 128) {
            double best = 1e30;
            int best_pair = 0, best_bn = 256;
            auto consider = [&](double cost, int pair, int bn) {
                if (force_bn && bn != force_bn) return;
                if (cost < best) { best = cost; best_pair = pair; best_bn = bn; }
            };
            consider(waves(m1 * n256, sms) * 1.0, 0, 256);
            consider(waves(m1 * n128, sms) * (0.54 + spill), 0, 128);
            if (pair_ok) {
                consider(waves(m2 * n256, pairs) * 0.90, 1, 256);
                consider(waves(m2 * n128, pairs) * (0.47 + spill), 1, 128);
            }
            if (best_pair) return linear_2cta_forward(c, stream, best_bn);
            return best_bn == 128 ? launch_tc<128, 0, false>(c, stream) : launch_tc<256, 0, false>(c, stream);
        }
        if (c.N > 64) return launch_tc<128, 0, false>(c, stream);
        return launch_tc<64, 0, false>(c, stream);
    }
    // weight-major
    const int bn = c.M <= 16 ? 16 : (c.M <= 32 ? 32 : 64);
    const int mode = c.epi == EPI_PARTIAL ? 1 : (c.epi == EPI_SWIGLU ? 2 : 0);
#define UMV_WM(BN_)                                                             \
    (mode == 0 ? launch_tc<BN_, 0, true>(c, stream)                             \
               : mode == 1 ? launch_tc<BN_, 1, true>(c, stream) : launch_tc<BN_, 2, true>(c, stream))
    if (bn == 16) return UMV_WM(16);
    if (bn == 32) return UMV_WM(32);
    return UMV_WM(64);
#undef UMV_WM
}

}  // namespace umv
