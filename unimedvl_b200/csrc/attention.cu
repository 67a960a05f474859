// Variable-length GQA attention for the packed LLM (head_dim 128, paged KV) and the ViT
// (head_dim 72, packed K/V): flash-style single pass, warp-level mma.sync m16n8k16 bf16 tiles,
// cp.async double-buffered 64-key blocks staged in (swizzled) shared memory, fp32 online softmax with
// warp-shuffle row reductions, probabilities rounded to bf16 before the PV product.
//
// Rows of a CTA tile are (token, head-in-group) pairs of ONE kv head, so a K/V block is read once
// for all `group` query heads that share it; the same kernel therefore serves prefill (many tokens)
// and decode (1 token x group heads) -- decode adds split-KV over key blocks plus a combine pass.
//
// Replaces flash_attn_varlen_func (flash-attn 2.x, external to the reference) at
// qwen2_navit.py:605-614 (causal = bottom-right aligned, or full) and siglip_navit.py:232-241, and the
// per-step full KV re-materialisation of qwen2_navit.py:589-600 (the cache is read in place).
#include <cstdlib>

#include "../../include/umv.h"
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace umv {

constexpr int kTileKeys = 64;   // keys per block == KV page size
// ROWS = query rows per CTA (16 per warp).  64 is what is built: at 128 rows the kernel needs > 128 registers per thread
// (64 O + 32 S + 32 Q fragments), i.e. one CTA per SM and no more resident warps than two 64-row CTAs.

template <int HD, int ROWS = 64>
struct AttnCfg {
    static constexpr int kThreads = ROWS * 2;                       // 16 rows per warp
    static constexpr int kChunks = HD / 8;                          // 16-byte chunks of real data per row
    static constexpr int kKSteps = (HD + 15) / 16;                  // QK^T k-steps (HD padded to 16)
    static constexpr int kDTiles = HD / 8;                          // PV n-tiles
    static constexpr bool kSwizzle = (HD == 128);
    static constexpr int kRowBytes = kSwizzle ? 256 : (kKSteps * 32 + 16);   // 72 -> 176 B (conflict-free ldmatrix)
    static constexpr int kTileBytes = kTileKeys * kRowBytes;        // one K or V block
    static constexpr int kQBytes = ROWS * kRowBytes;
    static constexpr int kStages = 3;                               // K/V blocks in flight (two 112 KB CTAs fit one SM)
    static constexpr int kSmemBytes = kQBytes + 2 * kStages * kTileBytes;     // Q + kStages x (K, V)
};

template <int HD>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
    if (AttnCfg<HD>::kSwizzle) return row * 256 + (((chunk ^ (row & 7)) & 15) << 4);
    return row * AttnCfg<HD>::kRowBytes + (chunk << 4);
}

template <int HD, int ROWS>
__global__ void __launch_bounds__(ROWS * 2) attn_fwd_kernel(AttnArgs a, float scale_log2) {
    using Cfg = AttnCfg<HD, ROWS>;
    constexpr int kTileRows = ROWS, kAttnThreads = ROWS * 2;
    pdl_launch_dependents();
    trace_start(a.trace);
    pdl_wait();
    trace_wait(a.trace);
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = smem + Cfg::kQBytes;
    uint8_t* sV = sK + Cfg::kStages * Cfg::kTileBytes;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y / a.Hkv, kvh = blockIdx.y % a.Hkv;
    const int split = blockIdx.z;
    const int G = a.H / a.Hkv;
    const int qlen = a.q_len[b], kvlen = a.kv_len[b];
    // causal: the last row tiles see the most keys -- schedule them first so the grid does not end on its longest CTAs
    const int r0 = (a.causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * kTileRows;
    const int nrows = qlen * G;
    if (r0 >= nrows) return;
    const int qs = a.q_start[b];
    const int ks0 = a.paged ? 0 : a.k_start[b];

    // visible key range of this tile
    const int last_tok = min(qlen - 1, (r0 + kTileRows - 1) / G);
    const int vis_keys = a.causal ? min(kvlen, kvlen - qlen + last_tok + 1) : kvlen;
    const int blocks_total = (vis_keys + kTileKeys - 1) / kTileKeys;
    const int bps = (blocks_total + a.splits - 1) / a.splits;
    const int kb_begin = split * bps, kb_end = min(blocks_total, kb_begin + bps);

    if (!Cfg::kSwizzle) {
        // zero the padding chunk (columns HD..HD+7) that the last QK^T k-step reads
        for (int i = tid; i < kTileRows + 2 * Cfg::kStages * kTileKeys; i += kAttnThreads)      // rows of Q, then of the K/V stage tiles
            *reinterpret_cast<U4*>(smem + tile_off<HD>(i, Cfg::kChunks)) = U4{0, 0, 0, 0};
    }

    // ---- stage Q
    for (int i = tid; i < kTileRows * Cfg::kChunks; i += kAttnThreads) {
        const int r = i / Cfg::kChunks, ch = i % Cfg::kChunks;
        const int R = r0 + r;
        const bool ok = R < nrows;
        const int tok = ok ? R / G : 0, head = kvh * G + (ok ? R % G : 0);
        cp_async16(sQ + tile_off<HD>(r, ch), a.q + (size_t)(qs + tok) * a.ldq + head * HD + ch * 8, ok);
    }
    auto load_kv = [&](int kb, int stage) {
        const bf16* kbase;
        const bf16* vbase;
        size_t kstride, vstride;
        if (a.paged) {
            const int page = a.page_table[(size_t)b * a.max_pages + kb];
            kbase = a.pool.base + a.pool.tile_offset(page, a.layer, 0, kvh);
            vbase = a.pool.base + a.pool.tile_offset(page, a.layer, 1, kvh);
            kstride = vstride = HD;
        } else {
            kbase = a.k + (size_t)(ks0 + kb * kTileKeys) * a.ldk + kvh * HD;
            vbase = a.v + (size_t)(ks0 + kb * kTileKeys) * a.ldv + kvh * HD;
            kstride = a.ldk;
            vstride = a.ldv;
        }
        uint8_t* dk = sK + stage * Cfg::kTileBytes;
        uint8_t* dv = sV + stage * Cfg::kTileBytes;
        for (int i = tid; i < kTileKeys * Cfg::kChunks; i += kAttnThreads) {
            const int r = i / Cfg::kChunks, ch = i % Cfg::kChunks;
            const bool ok = kb * kTileKeys + r < kvlen;
            cp_async16(dk + tile_off<HD>(r, ch), kbase + (size_t)(ok ? r : 0) * kstride + ch * 8, ok);
            cp_async16(dv + tile_off<HD>(r, ch), vbase + (size_t)(ok ? r : 0) * vstride + ch * 8, ok);
        }
    };
    // ring of kStages blocks, one commit group per slot (group 0 also carries Q); a refill group per finished block
#pragma unroll
    for (int j = 0; j < Cfg::kStages; ++j) {
        if (kb_begin + j < kb_end) load_kv(kb_begin + j, j);
        cp_async_commit();
    }

    const int g = lane >> 2, t = lane & 3;
    const int Ra = r0 + warp * 16 + g, Rb = Ra + 8;
    const int tok_a = min(Ra, nrows - 1) / G, tok_b = min(Rb, nrows - 1) / G;
    const int lim_a = a.causal ? (kvlen - qlen + tok_a) : (kvlen - 1);   // last visible key index per row
    const int lim_b = a.causal ? (kvlen - qlen + tok_b) : (kvlen - 1);

    uint32_t qf[Cfg::kKSteps][4];
    float o[Cfg::kDTiles][4];
#pragma unroll
    for (int i = 0; i < Cfg::kDTiles; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;

    for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int stage = (kb - kb_begin) % Cfg::kStages;
        cp_async_wait<Cfg::kStages - 1>();       // kStages + i groups committed at iteration i: group i has landed
        __syncthreads();
        if (kb == kb_begin) {
#pragma unroll
            for (int ks = 0; ks < Cfg::kKSteps; ++ks)
                ldmatrix_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3],
                            smem_u32(sQ + tile_off<HD>(warp * 16 + (lane & 15), ks * 2 + (lane >> 4))));
        }
        const uint8_t* cK = sK + stage * Cfg::kTileBytes;
        const uint8_t* cV = sV + stage * Cfg::kTileBytes;

        // ---- S = Q K^T  (16 rows x 64 keys per warp)
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < Cfg::kKSteps; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(b0, b1, b2, b3,
                            smem_u32(cK + tile_off<HD>(np * 16 + (lane & 7) + ((lane >> 4) << 3), ks * 2 + ((lane >> 3) & 1))));
                mma_bf16_16816(s[2 * np], qf[ks], b0, b1);
                mma_bf16_16816(s[2 * np + 1], qf[ks], b2, b3);
            }
        }
        // ---- mask, online softmax
        float mx_a = -INFINITY, mx_b = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int key = kb * kTileKeys + nt * 8 + 2 * t + e;
                s[nt][e] = key <= lim_a ? s[nt][e] * scale_log2 : -INFINITY;
                s[nt][2 + e] = key <= lim_b ? s[nt][2 + e] * scale_log2 : -INFINITY;
                mx_a = fmaxf(mx_a, s[nt][e]);
                mx_b = fmaxf(mx_b, s[nt][2 + e]);
            }
        }
        mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1));
        mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
        mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1));
        mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
        const float mn_a = fmaxf(m_a, mx_a), mn_b = fmaxf(m_b, mx_b);
        const float ms_a = mn_a == -INFINITY ? 0.f : mn_a, ms_b = mn_b == -INFINITY ? 0.f : mn_b;
        const float al_a = exp2f(m_a - ms_a), al_b = exp2f(m_b - ms_b);
        m_a = mn_a;
        m_b = mn_b;
        float rs_a = 0.f, rs_b = 0.f;
        uint32_t pf[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p0 = exp2f(s[nt][0] - ms_a), p1 = exp2f(s[nt][1] - ms_a);
            const float p2 = exp2f(s[nt][2] - ms_b), p3 = exp2f(s[nt][3] - ms_b);
            rs_a += p0 + p1;
            rs_b += p2 + p3;
            // accumulator layout of two adjacent n-tiles == A-fragment layout of one 16-key k-step
            pf[nt >> 1][(nt & 1) * 2 + 0] = pack2(p0, p1);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack2(p2, p3);
        }
        l_a = l_a * al_a + rs_a;
        l_b = l_b * al_b + rs_b;
#pragma unroll
        for (int dt = 0; dt < Cfg::kDTiles; ++dt) {
            o[dt][0] *= al_a; o[dt][1] *= al_a;
            o[dt][2] *= al_b; o[dt][3] *= al_b;
        }
        // ---- O += P V
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int dp = 0; dp + 1 < Cfg::kDTiles; dp += 2) {
                uint32_t b0, b1, b2, b3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                             : "r"(smem_u32(cV + tile_off<HD>(kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), dp + (lane >> 4)))));
                mma_bf16_16816(o[dp], pf[kk], b0, b1);
                mma_bf16_16816(o[dp + 1], pf[kk], b2, b3);
            }
            if (Cfg::kDTiles & 1) {
                uint32_t b0, b1;
                ldmatrix_x2_trans(b0, b1, smem_u32(cV + tile_off<HD>(kk * 16 + (lane & 15), Cfg::kDTiles - 1)));
                mma_bf16_16816(o[Cfg::kDTiles - 1], pf[kk], b0, b1);
            }
        }
        __syncthreads();                          // the slot is free: refill it
        if (kb + Cfg::kStages < kb_end) load_kv(kb + Cfg::kStages, stage);
        cp_async_commit();
    }
    cp_async_wait<0>();

    l_a += __shfl_xor_sync(0xffffffffu, l_a, 1);
    l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);
    l_b += __shfl_xor_sync(0xffffffffu, l_b, 1);
    l_b += __shfl_xor_sync(0xffffffffu, l_b, 2);
    const float il_a = l_a > 0.f ? 1.f / l_a : 0.f, il_b = l_b > 0.f ? 1.f / l_b : 0.f;

#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int R = half ? Rb : Ra;
        if (R >= nrows) continue;
        const int tok = R / G, head = kvh * G + R % G;
        const float il = half ? il_b : il_a;
        if (a.splits == 1) {
            bf16* dst = a.out + (size_t)(a.out_row_map ? a.out_row_map[qs + tok] : qs + tok) * a.ldo + head * HD;
#pragma unroll
            for (int dt = 0; dt < Cfg::kDTiles; ++dt)
                *reinterpret_cast<uint32_t*>(dst + dt * 8 + 2 * t) = pack2(o[dt][half * 2] * il, o[dt][half * 2 + 1] * il);
        } else {
            const size_t ridx = (size_t)(qs + tok) * a.H + head;
            float* dst = a.ws + ((size_t)split * a.total_q * a.H + ridx) * HD;
#pragma unroll
            for (int dt = 0; dt < Cfg::kDTiles; ++dt)
                *reinterpret_cast<float2*>(dst + dt * 8 + 2 * t) = make_float2(o[dt][half * 2] * il, o[dt][half * 2 + 1] * il);
            if (t == 0) {
                const float m = half ? m_b : m_a, l = half ? l_b : l_a;
                a.ws[(size_t)a.splits * a.total_q * a.H * HD + (size_t)split * a.total_q * a.H + ridx] =
                    l > 0.f ? m + log2f(l) : -INFINITY;
            }
        }
    }
    trace_end<false>(a.trace);
}

// Split-KV combine: out = sum_s w_s O_s / sum_s w_s, w_s = 2^(lse_s - max lse).   one warp per (token, head)
template <int HD>
__global__ void attn_combine_kernel(AttnArgs a) {
    pdl_launch_dependents();
    trace_start(a.trace_combine);
    pdl_wait();
    trace_wait(a.trace_combine);
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int rows = a.total_q * a.H;
    if (gw >= rows) return;
    const float* lse = a.ws + (size_t)a.splits * rows * HD;
    float mx = -INFINITY;
    for (int s = 0; s < a.splits; ++s) mx = fmaxf(mx, lse[(size_t)s * rows + gw]);
    float acc[(HD + 31) / 32];
#pragma unroll
    for (int i = 0; i < (HD + 31) / 32; ++i) acc[i] = 0.f;
    float wsum = 0.f;
    for (int s = 0; s < a.splits; ++s) {
        const float l = lse[(size_t)s * rows + gw];
        const float w = (l == -INFINITY) ? 0.f : exp2f(l - mx);
        wsum += w;
        const float* src = a.ws + ((size_t)s * rows + gw) * HD;
#pragma unroll
        for (int i = 0; i < (HD + 31) / 32; ++i) {
            const int d = lane + i * 32;
            if (d < HD) acc[i] += w * src[d];
        }
    }
    const float inv = wsum > 0.f ? 1.f / wsum : 0.f;
    const int tok = gw / a.H, head = gw % a.H;
    bf16* dst = a.out + (size_t)(a.out_row_map ? a.out_row_map[tok] : tok) * a.ldo + head * HD;
#pragma unroll
    for (int i = 0; i < (HD + 31) / 32; ++i) {
        const int d = lane + i * 32;
        if (d < HD) dst[d] = f2b(acc[i] * inv);
    }
    trace_end<false>(a.trace_combine);
}


// ---------------------------------------------------------------------------------------------
// Fused decode attention: ONE launch per layer for the chain
//   split-K reduce (+bias) of the q/k/v projection -> q/k RMSNorm -> RoPE -> KV append -> attention over the paged
//   cache -> split-KV combine.
// A thread-block cluster of `S` CTAs (one CTA per SM) owns one (sample, kv head); CTA r covers a contiguous, balanced range
// of 64-key blocks.  The kernel is on the critical path of the decode chain and moves little data (18 MB per layer at B=8,
// ctx 1.1k), so it is organised around the number of DEPENDENT memory round trips, not around bandwidth:
//   1. norm weights and bias (constants) are read before griddepcontrol.wait; after it every load that does not depend on
//      another load is issued at once: kv_len, the sample's page-table row (-> shared memory), the per-step cos/sin table,
//      all split-K partials of this CTA's q/k/v rows;
//   2. as soon as the page row is known, up to six blocks (the CTA's whole range at ctx <= 1.5k) are requested from TMA
//      -- 64 x 64 swizzled boxes straight out of the paged pool, one issuing lane and one mbarrier per ring slot -- while
//      the warps normalise / rotate the G query rows and the new K/V row; the new row goes to the page in global memory and
//      is patched into the staged tile in shared memory (no reload);
//   3. the 8 warps take two blocks at a time (4 key quarters each; only G <= 7 query rows exist); partial (m, l, O) are
//      merged across warps in shared memory and then PUSHED to the CTA that finishes the head through distributed
//      shared memory (remote stores, no remote round trip), one cluster barrier, local combine, store.
// In-graph timeline (us after the dependency wait, B200): loads landed 1.5, Q ready 2.6, first K/V block 4.3, key loop done
// 5.5, merged 6.2, cluster barrier 8.2, stored 8.8 (profiles/r1_decode_timeline.md).
// Numerics identical to rope_append_kernel + attn_fwd_kernel + attn_combine_kernel (und mode, R4-R6).
constexpr int kDecThreads = 256;
constexpr int kDecStages = 6;                                           // 64-key blocks in flight per CTA (one CTA per SM)
constexpr int kDecQBytes = 16 * 256;
constexpr int kDecStageBytes = 2 * kTileKeys * 256;                    // K tile + V tile
constexpr int kDecSmemBytes = 1024 + kDecQBytes + kDecStages * kDecStageBytes;   // + slack to align the TMA tiles to 1 KB
constexpr int kDecMaxPages = 512;                                       // page-table row staged in shared memory
constexpr int kDecMaxSplits = 8;                                        // split-K partials summed with all loads in flight
constexpr int kDecRecvRows = 16;                                        // >= S * ceil(G / S) for G <= 8, S <= 8

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_rank(const void* smem_ptr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(smem_ptr)), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_f1(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// A K or V tile as TMA leaves it: two [64 slots x 64 columns] halves (8 KB each), rows of 128 B with the 128-byte swizzle.
__device__ __forceinline__ uint32_t kv_off(int row, int chunk) {
    return ((chunk >> 3) << 13) + row * 128 + ((((chunk & 7) ^ (row & 7))) << 4);
}

__global__ void __launch_bounds__(kDecThreads, 1)
attn_decode_kernel(const __grid_constant__ CUtensorMap tmKV, DecodeAttnArgs a, float scale_log2) {
    constexpr int HD = 128;
    pdl_launch_dependents();
    trace_start(a.trace);
    trace_dbg_max(a.trace, 3);             // dbg3: the LAST CTA of the grid to start
    cluster_arrive();                      // phase 1: "this CTA runs" (its shared memory may be written by peers later)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem;
    uint8_t* sKV = smem + kDecQBytes;
    __shared__ uint64_t kvbar[kDecStages];
    __shared__ int s_pages[kDecMaxPages];
    __shared__ __align__(16) float recv_o[kDecRecvRows][HD];
    __shared__ float recv_m[kDecRecvRows], recv_l[kDecRecvRows];
    // after the key loop the stage ring is dead: the cross-warp merge buffers live there
    float* red_o = reinterpret_cast<float*>(sKV);                     // [8 warps][8 rows][HD]
    float* red_m = red_o + 8 * 8 * HD;                                // [8][8]
    float* red_l = red_m + 64;                                        // [8][8]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t S, rank;
    asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(S));
    asm volatile("mov.u32 %0, %%cluster_ctaid.x;" : "=r"(rank));
    const int b = blockIdx.y / a.Hkv, kvh = blockIdx.y % a.Hkv;
    const int G = a.H / a.Hkv;
    const int hpc = (G + (int)S - 1) / (int)S;                        // heads finished per CTA
    const int ncols = (a.H + 2 * a.Hkv) * HD;

    // ---- constants (weights): before the dependency wait
    if (tid == 0) {
        tma_prefetch_desc(&tmKV);
        for (int j = 0; j < kDecStages; ++j) mbar_init(&kvbar[j], 1);
        fence_mbar_init();
    }
    for (int i = tid; i < 16 * 16; i += kDecThreads) {                // padding rows of the Q tile
        const int r = i >> 4, ch = i & 15;
        if (r >= G) *reinterpret_cast<U4*>(sQ + tile_off<HD>(r, ch)) = U4{0, 0, 0, 0};
    }
    // row tasks: warps 0..G-1 one query head each, warp G the new K row and the new V row
    const bool is_q = warp < G, is_kv = warp == G;
    const int col_a = (is_q ? kvh * G + warp : a.H + kvh) * HD + lane * 4;          // q head / k row
    const int col_v = (a.H + a.Hkv + kvh) * HD + lane * 4;
    uint2 wv = make_uint2(0u, 0u), ba = make_uint2(0u, 0u), bv = make_uint2(0u, 0u);
    if (is_q || is_kv) {
        wv = *reinterpret_cast<const uint2*>((is_q ? a.qn : a.kn) + lane * 4);
        if (a.partial) {
            ba = *reinterpret_cast<const uint2*>(a.bias + col_a);
            if (is_kv) bv = *reinterpret_cast<const uint2*>(a.bias + col_v);
        }
    }
    // Two groups of loads.  STATE (kv_len, the page-table row, the step's cos / sin) belongs to the decode step, not to the previous
    // kernel: with `a.early` (every layer but the first, whose predecessors are the step's own state kernels) it is read -- and the
    // K/V tiles requested from TMA -- BEFORE griddepcontrol.wait, while the q/k/v projection is still running: the cache rows of earlier
    // tokens were written a whole forward ago.  The projection's split-K partials are the only loads that need the wait.
    int kvlen = 0, blocks_total = 0, kb_begin = 0, kb_end = 0, last_block = 0, new_slot = 0;
    bool owner = false;
    float xa[4] = {0.f, 0.f, 0.f, 0.f}, xv[4] = {0.f, 0.f, 0.f, 0.f};
    float4 rc = make_float4(0.f, 0.f, 0.f, 0.f), rs = rc;            // bf16-rounded cos / sin of this lane's 4 elements
    // ---- round trip 2 (issue): one thread asks TMA for up to kDecStages blocks (the CTA's whole range at B = 8, ctx <= 1.5k): four
    // 8 KB boxes per block (K / V x column halves) completing on the slot's mbarrier.  The block that receives the new
    // token is loaded like the others; its new row is patched in shared memory from the registers of the K/V warp.
    auto issue_kv = [&](int kb, int stage) {
        const int page = s_pages[kb];
        const int row_k = (int)(a.pool.tile_offset(page, a.layer, 0, kvh) / HD);
        const int row_v = (int)(a.pool.tile_offset(page, a.layer, 1, kvh) / HD);
        uint8_t* dk = sKV + stage * kDecStageBytes;
        mbar_expect_tx(&kvbar[stage], kDecStageBytes);
        tma_load_2d(dk, &tmKV, &kvbar[stage], 0, row_k, kEvictNormal);
        tma_load_2d(dk + 8192, &tmKV, &kvbar[stage], 64, row_k, kEvictNormal);
        tma_load_2d(dk + 16384, &tmKV, &kvbar[stage], 0, row_v, kEvictNormal);
        tma_load_2d(dk + 24576, &tmKV, &kvbar[stage], 64, row_v, kEvictNormal);
    };
    auto load_state = [&]() {
        kvlen = a.kv_len[b];
        for (int i = tid; i < a.max_pages; i += kDecThreads) s_pages[i] = a.page_table[(size_t)b * a.max_pages + i];
        if ((is_q || is_kv) && a.rope_cs) {
            rc = *reinterpret_cast<const float4*>(a.rope_cs + (size_t)b * HD + (lane * 4) % (HD / 2));
            rs = *reinterpret_cast<const float4*>(a.rope_cs + (size_t)b * HD + HD / 2 + (lane * 4) % (HD / 2));
        }
    };
    auto request_tiles = [&]() {
        // key-block range of this CTA (balanced: sizes differ by at most one)
        blocks_total = (kvlen + kTileKeys - 1) / kTileKeys;
        kb_begin = (int)(((long long)blocks_total * rank) / S);
        kb_end = (int)(((long long)blocks_total * (rank + 1)) / S);
        last_block = (kvlen - 1) / kTileKeys;
        new_slot = (kvlen - 1) % kTileKeys;
        owner = kb_begin <= last_block && last_block < kb_end;     // this CTA appends the new K/V row
        __syncthreads();                                           // s_pages (and the Q padding, the mbarriers) visible
        // one issuing lane per block (lane 0 of warp j takes ring slot j): the descriptor fetch + 4 requests of a block cost a
        // few hundred cycles of a single thread, and every warp has to pass the barrier below before the key loop starts
        static_assert(kDecStages <= kDecThreads / 32, "one warp per first-pass ring slot");
        if (lane == 0 && warp < kDecStages && kb_begin + warp < kb_end) issue_kv(kb_begin + warp, warp);
    };
    if (a.early) {
        load_state();
        request_tiles();
    }
    pdl_wait();
    trace_wait(a.trace);

    // ---- round trip 1: everything that needs no other load
    if (!a.early) load_state();
    if (is_q || is_kv) {
        if (!a.rope_cs) {
            const float pos = (float)a.positions[b];
            const float4 fr = *reinterpret_cast<const float4*>(a.inv_freq + (lane * 4) % (HD / 2));
            const float f[4] = {fr.x, fr.y, fr.z, fr.w};
            float c[4], sn[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float ang = __fmul_rn(pos, f[j]);
                c[j] = rbf(cosf(ang));
                sn[j] = rbf(sinf(ang));
            }
            rc = make_float4(c[0], c[1], c[2], c[3]);
            rs = make_float4(sn[0], sn[1], sn[2], sn[3]);
        }
        if (a.partial) {
            float4 pa[kDecMaxSplits], pv[kDecMaxSplits];
#pragma unroll
            for (int sp = 0; sp < kDecMaxSplits; ++sp) {
                pa[sp] = pv[sp] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (sp < a.ksplits) {
                    const float* rowp = a.partial + ((size_t)sp * a.M + b) * ncols;
                    pa[sp] = *reinterpret_cast<const float4*>(rowp + col_a);
                    if (is_kv) pv[sp] = *reinterpret_cast<const float4*>(rowp + col_v);
                }
            }
            float sa[4] = {0.f, 0.f, 0.f, 0.f}, sv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int sp = 0; sp < kDecMaxSplits; ++sp) {            // fixed order; +0.f beyond ksplits is exact
                if (sp < a.ksplits) {
                    sa[0] += pa[sp].x; sa[1] += pa[sp].y; sa[2] += pa[sp].z; sa[3] += pa[sp].w;
                    sv[0] += pv[sp].x; sv[1] += pv[sp].y; sv[2] += pv[sp].z; sv[3] += pv[sp].w;
                }
            }
            const float2 a0 = unpack2(ba.x), a1 = unpack2(ba.y), v0 = unpack2(bv.x), v1 = unpack2(bv.y);
            xa[0] = rbf(sa[0] + a0.x); xa[1] = rbf(sa[1] + a0.y); xa[2] = rbf(sa[2] + a1.x); xa[3] = rbf(sa[3] + a1.y);
            xv[0] = rbf(sv[0] + v0.x); xv[1] = rbf(sv[1] + v0.y); xv[2] = rbf(sv[2] + v1.x); xv[3] = rbf(sv[3] + v1.y);
        } else {
            const uint2 qa = *reinterpret_cast<const uint2*>(a.qkv + (size_t)b * ncols + col_a);
            uint2 qv = make_uint2(0u, 0u);
            if (is_kv) qv = *reinterpret_cast<const uint2*>(a.qkv + (size_t)b * ncols + col_v);
            const float2 f0 = unpack2(qa.x), f1 = unpack2(qa.y), g0 = unpack2(qv.x), g1 = unpack2(qv.y);
            xa[0] = f0.x; xa[1] = f0.y; xa[2] = f1.x; xa[3] = f1.y;
            xv[0] = g0.x; xv[1] = g0.y; xv[2] = g1.x; xv[3] = g1.y;
        }
    }

    if (!a.early) request_tiles();
    trace_dbg(a.trace, 0);

    // ---- rows of this step: RMSNorm + RoPE of the query heads (-> sQ) and of the new K row, V row as is (-> page)
    uint2 k_new = make_uint2(0u, 0u), v_new = make_uint2(0u, 0u);
    if (is_q || is_kv) {
        const float2 w0 = unpack2(wv.x), w1 = unpack2(wv.y);
        const float nw[4] = {w0.x, w0.y, w1.x, w1.y};
        const float c[4] = {rc.x, rc.y, rc.z, rc.w}, sn[4] = {rs.x, rs.y, rs.z, rs.w};
        float ss = xa[0] * xa[0] + xa[1] * xa[1] + xa[2] * xa[2] + xa[3] * xa[3];
        ss = warp_sum(ss);
        const float inv = 1.0f / sqrtf(ss / (float)HD + a.eps);
        float n[4], o4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) n[j] = rbf(nw[j] * rbf(xa[j] * inv));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float partner = __shfl_xor_sync(0xffffffffu, n[j], 16);
            const float rot = (lane < 16) ? -partner : partner;
            o4[j] = rbf(rbf(n[j] * c[j]) + rbf(rot * sn[j]));
        }
        const uint2 packed = make_uint2(pack2(o4[0], o4[1]), pack2(o4[2], o4[3]));
        if (is_q) {
            *reinterpret_cast<uint2*>(sQ + tile_off<HD>(warp, lane >> 1) + (lane & 1) * 8) = packed;
        } else if (owner) {
            k_new = packed;
            v_new = make_uint2(pack2(xv[0], xv[1]), pack2(xv[2], xv[3]));
            const int page = s_pages[last_block];
            bf16* kdst = a.pool.base + a.pool.tile_offset(page, a.layer, 0, kvh) + (size_t)new_slot * HD + lane * 4;
            bf16* vdst = a.pool.base + a.pool.tile_offset(page, a.layer, 1, kvh) + (size_t)new_slot * HD + lane * 4;
            *reinterpret_cast<uint2*>(kdst) = k_new;
            *reinterpret_cast<uint2*>(vdst) = v_new;
        }
    }
    __syncthreads();                       // sQ complete
    trace_dbg(a.trace, 1);

    const int g = lane >> 2, t = lane & 3;
    const int quarter = warp & 3, sub = warp >> 2;      // warp = (16-key quarter, which block of the round's pair)
    uint32_t qf[8][4];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
        ldmatrix_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], smem_u32(sQ + tile_off<HD>(lane & 15, ks * 2 + (lane >> 4))));
    float o[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = 0.f;
    float m_a = -INFINITY, l_a = 0.f;

    const int nblk = kb_end - kb_begin;
    for (int rd = 0; 2 * rd < nblk; ++rd) {
        const int j = 2 * rd + sub, kb = kb_begin + j, stage = j % kDecStages;
        if (owner && (last_block == kb_begin + 2 * rd || last_block == kb_begin + 2 * rd + 1)) {      // CTA-uniform
            // the newest block: wait for its tile, clear the V rows of slots that do not exist yet (whatever the pool holds
            // there is multiplied by probability 0), put the new K / V row in place
            const int jl = last_block - kb_begin;
            uint8_t* tk = sKV + (jl % kDecStages) * kDecStageBytes;
            mbar_wait(&kvbar[jl % kDecStages], (jl / kDecStages) & 1);
            for (int i = tid; i < (kTileKeys - 1 - new_slot) * 16; i += kDecThreads)
                *reinterpret_cast<U4*>(tk + 16384 + kv_off(new_slot + 1 + (i >> 4), i & 15)) = U4{0, 0, 0, 0};
            if (is_kv) {
                *reinterpret_cast<uint2*>(tk + kv_off(new_slot, lane >> 1) + (lane & 1) * 8) = k_new;
                *reinterpret_cast<uint2*>(tk + 16384 + kv_off(new_slot, lane >> 1) + (lane & 1) * 8) = v_new;
            }
            __syncthreads();
        }
        if (j < nblk) mbar_wait(&kvbar[stage], (j / kDecStages) & 1);
        if (rd == 0) trace_dbg(a.trace, 2);
        const uint8_t* cK = sKV + stage * kDecStageBytes;
        const uint8_t* cV = cK + 16384;
        if (j < nblk) {
            float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(b0, b1, b2, b3,
                            smem_u32(cK + kv_off(quarter * 16 + (lane & 7) + ((lane >> 4) << 3), ks * 2 + ((lane >> 3) & 1))));
                mma_bf16_16816(s0, qf[ks], b0, b1);
                mma_bf16_16816(s1, qf[ks], b2, b3);
            }
            const int key0 = kb * kTileKeys + quarter * 16 + 2 * t;
            float v00 = key0 < kvlen ? s0[0] * scale_log2 : -INFINITY;
            float v01 = key0 + 1 < kvlen ? s0[1] * scale_log2 : -INFINITY;
            float v10 = key0 + 8 < kvlen ? s1[0] * scale_log2 : -INFINITY;
            float v11 = key0 + 9 < kvlen ? s1[1] * scale_log2 : -INFINITY;
            float mx = fmaxf(fmaxf(v00, v01), fmaxf(v10, v11));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const float mn = fmaxf(m_a, mx);
            const float ms = mn == -INFINITY ? 0.f : mn;
            const float al = exp2f(m_a - ms);
            m_a = mn;
            const float p00 = exp2f(v00 - ms), p01 = exp2f(v01 - ms), p10 = exp2f(v10 - ms), p11 = exp2f(v11 - ms);
            l_a = l_a * al + (p00 + p01 + p10 + p11);
            const uint32_t pf[4] = {pack2(p00, p01), 0u, pack2(p10, p11), 0u};      // rows 8..15 of the tile are padding
#pragma unroll
            for (int dt = 0; dt < 16; ++dt) { o[dt][0] *= al; o[dt][1] *= al; }
#pragma unroll
            for (int dp = 0; dp < 16; dp += 2) {
                uint32_t b0, b1, b2, b3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                             : "r"(smem_u32(cV + kv_off(quarter * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), dp + (lane >> 4)))));
                float acc0[4] = {o[dp][0], o[dp][1], 0.f, 0.f}, acc1[4] = {o[dp + 1][0], o[dp + 1][1], 0.f, 0.f};
                mma_bf16_16816(acc0, pf, b0, b1);
                mma_bf16_16816(acc1, pf, b2, b3);
                o[dp][0] = acc0[0]; o[dp][1] = acc0[1];
                o[dp + 1][0] = acc1[0]; o[dp + 1][1] = acc1[1];
            }
        }
        if (2 * rd + kDecStages < nblk) {       // long contexts only: refill the two slots just consumed
            __syncthreads();
            if (tid == 0) {
                fence_proxy_async_smem();       // the generic-proxy reads of the slots are ordered before the TMA writes
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int jn = 2 * rd + u + kDecStages;
                    if (jn < nblk) issue_kv(kb_begin + jn, jn % kDecStages);
                }
            }
        }
    }
    l_a += __shfl_xor_sync(0xffffffffu, l_a, 1);
    l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);

    // ---- merge the 8 (key-quarter, block parity) partials of this CTA (buffers alias the dead stage ring)
    __syncthreads();
    trace_dbg(a.trace, 5);
    trace_dbg_max(a.trace, 4);             // dbg4: the LAST CTA of the grid to finish its key loop
    if (t == 0) { red_m[warp * 8 + g] = m_a; red_l[warp * 8 + g] = l_a; }
#pragma unroll
    for (int dt = 0; dt < 16; ++dt)
        *reinterpret_cast<float2*>(&red_o[(warp * 8 + g) * HD + dt * 8 + 2 * t]) = make_float2(o[dt][0], o[dt][1]);
    __syncthreads();
    cluster_wait();                        // phase 1 complete: every CTA of the cluster is running
    trace_dbg(a.trace, 6);
    if (tid < 128) {
        const int row = tid >> 4, d0 = (tid & 15) * 8;
        if (row < G) {
            float M = -INFINITY;
#pragma unroll
            for (int w = 0; w < 8; ++w) M = fmaxf(M, red_m[w * 8 + row]);
            float L = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const float mw = red_m[w * 8 + row];
                const float wg = (mw == -INFINITY) ? 0.f : exp2f(mw - M);
                L += wg * red_l[w * 8 + row];
                const float4 x0 = *reinterpret_cast<const float4*>(&red_o[(w * 8 + row) * HD + d0]);
                const float4 x1 = *reinterpret_cast<const float4*>(&red_o[(w * 8 + row) * HD + d0 + 4]);
                acc[0] += wg * x0.x; acc[1] += wg * x0.y; acc[2] += wg * x0.z; acc[3] += wg * x0.w;
                acc[4] += wg * x1.x; acc[5] += wg * x1.y; acc[6] += wg * x1.z; acc[7] += wg * x1.w;
            }
            // push this key range's partial of head `row` to the CTA that finishes the head
            const uint32_t dst = (uint32_t)row % S;
            const int slot = (int)rank * hpc + row / (int)S;
            const uint32_t ro = map_to_rank(&recv_o[slot][d0], dst);
            st_cluster_f4(ro, make_float4(acc[0], acc[1], acc[2], acc[3]));
            st_cluster_f4(ro + 16, make_float4(acc[4], acc[5], acc[6], acc[7]));
            if ((tid & 15) == 0) {
                st_cluster_f1(map_to_rank(&recv_m[slot], dst), M);
                st_cluster_f1(map_to_rank(&recv_l[slot], dst), L);
            }
        }
    }
    cluster_arrive();                      // phase 2: pushes released ...
    cluster_wait();                        // ... and everyone's pushes into this CTA acquired
    trace_dbg(a.trace, 7);
    // ---- finish heads rank, rank + S, ...: merge the S key ranges in a fixed order (deterministic)
    for (int hi = tid >> 7; hi < hpc; hi += kDecThreads / 128) {
        const int head = (int)rank + hi * (int)S, d = tid & 127;
        if (head >= G) break;
        float M = -INFINITY;
        for (int r = 0; r < (int)S; ++r) M = fmaxf(M, recv_m[r * hpc + hi]);
        float L = 0.f, acc = 0.f;
        for (int r = 0; r < (int)S; ++r) {
            const float mr = recv_m[r * hpc + hi];
            const float wg = (mr == -INFINITY) ? 0.f : exp2f(mr - M);
            L += wg * recv_l[r * hpc + hi];
            acc += wg * recv_o[r * hpc + hi][d];
        }
        a.out[(size_t)b * a.ldo + (kvh * G + head) * HD + d] = f2b(L > 0.f ? acc / L : 0.f);
    }
    trace_end(a.trace);
}

int decode_attention(const DecodeAttnArgs& a, cudaStream_t s) {
    const int G = a.Hkv > 0 ? a.H / a.Hkv : 0;
    UMV_REQUIRE(a.H % a.Hkv == 0 && G <= 7, UMV_ERR_UNSUPPORTED, "decode_attention: GQA group %d > 7", G);
    UMV_REQUIRE(a.cluster >= 1 && a.cluster <= 8, UMV_ERR_INVALID, "decode_attention: cluster size %d", a.cluster);
    UMV_REQUIRE(a.cluster * ((G + a.cluster - 1) / a.cluster) <= kDecRecvRows, UMV_ERR_UNSUPPORTED, "decode_attention: G=%d S=%d", G, a.cluster);
    UMV_REQUIRE(a.max_pages <= kDecMaxPages, UMV_ERR_UNSUPPORTED, "decode_attention: %d pages per sample > %d", a.max_pages, kDecMaxPages);
    UMV_REQUIRE(a.partial == nullptr || a.ksplits <= kDecMaxSplits, UMV_ERR_UNSUPPORTED, "decode_attention: %d split-K partials > %d",
                a.ksplits, kDecMaxSplits);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(a.cluster, a.M * a.Hkv);
    cfg.blockDim = dim3(kDecThreads);
    cfg.dynamicSmemBytes = kDecSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = a.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 2 : 1;
    const float scale_log2 = (1.0f / sqrtf(128.0f)) * 1.4426950408889634f;
    UMV_REQUIRE(a.kv_tmap != nullptr, UMV_ERR_INVALID, "decode_attention: the KV pool tensor map is missing");
    DecodeAttnArgs at = a;
    at.trace = trace_next("attn_decode");
    at.kv_tmap = nullptr;
    cudaError_t e = cudaLaunchKernelEx(&cfg, attn_decode_kernel, *a.kv_tmap, at, scale_log2);
    ++g_launches;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("attn_decode_kernel launch failed: %s", cudaGetErrorString(e));
        return UMV_ERR_CUDA;
    }
    return UMV_OK;
}

// Shapes the fused decode kernel covers (else: rope_append + attn_fwd + attn_combine).
bool decode_attention_supported(int H, int Hkv, int dh, int max_pages, int ksplits) {
    return dh == 128 && Hkv > 0 && H % Hkv == 0 && H / Hkv <= 7 && max_pages <= kDecMaxPages && ksplits <= kDecMaxSplits;
}

int attention_init() {
    cudaFuncSetAttribute(attn_fwd_kernel<128, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<128, 64>::kSmemBytes);
    cudaFuncSetAttribute(attn_fwd_kernel<72, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<72, 64>::kSmemBytes);
    cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDecSmemBytes);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("attention_init: %s", cudaGetErrorString(e));
        return UMV_ERR_CUDA;
    }
    return UMV_OK;
}

template <int HD, int ROWS>
static int launch_attn(const AttnArgs& a0, cudaStream_t s) {
    AttnArgs a = a0;
    a.trace = trace_next("attn_fwd");
    if (a.splits > 1) a.trace_combine = trace_next("attn_combine");
    const int G = a.H / a.Hkv;
    const int row_tiles = (a.max_q_len * G + ROWS - 1) / ROWS;
    const float scale_log2 = (1.0f / sqrtf((float)HD)) * 1.4426950408889634f;
    dim3 grid(row_tiles, a.n * a.Hkv, a.splits);
    cudaError_t e = launch_k(attn_fwd_kernel<HD, ROWS>, grid, dim3(ROWS * 2), AttnCfg<HD, ROWS>::kSmemBytes, s, a, scale_log2);
    ++g_launches;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("attn_fwd_kernel<%d,%d> launch failed: %s", HD, ROWS, cudaGetErrorString(e));
        return UMV_ERR_CUDA;
    }
    if (a.splits > 1) {
        const int rows = a.total_q * a.H;
        e = launch_k(attn_combine_kernel<HD>, dim3((rows * 32 + 127) / 128), dim3(128), 0, s, a);
        ++g_launches;
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            set_error("attn_combine_kernel launch failed: %s", cudaGetErrorString(e));
            return UMV_ERR_CUDA;
        }
    }
    return UMV_OK;
}

int attention_forward(const AttnArgs& a, cudaStream_t s) {
    if (a.n <= 0 || a.max_q_len <= 0) return UMV_OK;
    UMV_REQUIRE(a.H % a.Hkv == 0, UMV_ERR_INVALID, "attention: heads %d not a multiple of kv heads %d", a.H, a.Hkv);
    UMV_REQUIRE(a.splits == 1 || a.ws != nullptr, UMV_ERR_INVALID, "attention: split-KV needs a workspace");
    UMV_REQUIRE(a.ldq % 8 == 0 && a.ldo % 2 == 0, UMV_ERR_INVALID, "attention: q/out row strides must keep 16-byte rows");
    const char* tc_env = getenv("UMV_ATTN_TC");            // read per call: tests switch paths inside one process
    const bool tc_on = !(tc_env && atoi(tc_env) == 0);
    if (tc_on && attention_tc_supported(a)) return attention_tc_forward(a, s);
    if (a.dh == 128) return launch_attn<128, 64>(a, s);
    if (a.dh == 72) return launch_attn<72, 64>(a, s);
    set_error("attention: head_dim %d is not built (128 and 72 are)", a.dh);
    return UMV_ERR_UNSUPPORTED;
}

}  // namespace umv
