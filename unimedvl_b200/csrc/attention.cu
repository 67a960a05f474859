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
#include <cooperative_groups.h>

#include "../../include/umv.h"
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace umv {

constexpr int kAttnThreads = 128;
constexpr int kTileRows = 64;   // query rows per CTA (4 warps x 16)
constexpr int kTileKeys = 64;   // keys per block == KV page size

template <int HD>
struct AttnCfg {
    static constexpr int kChunks = HD / 8;                          // 16-byte chunks of real data per row
    static constexpr int kKSteps = (HD + 15) / 16;                  // QK^T k-steps (HD padded to 16)
    static constexpr int kDTiles = HD / 8;                          // PV n-tiles
    static constexpr bool kSwizzle = (HD == 128);
    static constexpr int kRowBytes = kSwizzle ? 256 : (kKSteps * 32 + 16);   // 72 -> 176 B (conflict-free ldmatrix)
    static constexpr int kTileBytes = kTileRows * kRowBytes;
    static constexpr int kSmemBytes = 5 * kTileBytes;               // Q + 2 x (K, V)
};

template <int HD>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
    if (AttnCfg<HD>::kSwizzle) return row * 256 + (((chunk ^ (row & 7)) & 15) << 4);
    return row * AttnCfg<HD>::kRowBytes + (chunk << 4);
}

template <int HD>
__global__ void __launch_bounds__(kAttnThreads) attn_fwd_kernel(AttnArgs a, float scale_log2) {
    using Cfg = AttnCfg<HD>;
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = smem + Cfg::kTileBytes;
    uint8_t* sV = smem + 3 * Cfg::kTileBytes;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y / a.Hkv, kvh = blockIdx.y % a.Hkv;
    const int split = blockIdx.z;
    const int G = a.H / a.Hkv;
    const int qlen = a.q_len[b], kvlen = a.kv_len[b];
    const int r0 = blockIdx.x * kTileRows;
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
        for (int i = tid; i < 5 * kTileRows; i += kAttnThreads)
            *reinterpret_cast<U4*>(smem + (i / kTileRows) * Cfg::kTileBytes + tile_off<HD>(i % kTileRows, Cfg::kChunks)) = U4{0, 0, 0, 0};
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
    if (kb_begin < kb_end) load_kv(kb_begin, 0);
    cp_async_commit();

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
        const int stage = (kb - kb_begin) & 1;
        if (kb + 1 < kb_end) {
            load_kv(kb + 1, stage ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
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
        __syncthreads();
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
            bf16* dst = a.out + (size_t)(qs + tok) * a.ldo + head * HD;
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
}

// Split-KV combine: out = sum_s w_s O_s / sum_s w_s, w_s = 2^(lse_s - max lse).   one warp per (token, head)
template <int HD>
__global__ void attn_combine_kernel(AttnArgs a) {
    pdl_launch_dependents();
    pdl_wait();
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
    bf16* dst = a.out + (size_t)tok * a.ldo + head * HD;
#pragma unroll
    for (int i = 0; i < (HD + 31) / 32; ++i) {
        const int d = lane + i * 32;
        if (d < HD) dst[d] = f2b(acc[i] * inv);
    }
}


// ---------------------------------------------------------------------------------------------
// Fused decode attention: ONE launch per layer for the chain
//   split-K reduce (+bias) of the q/k/v projection -> q/k RMSNorm -> RoPE -> KV append -> attention over the paged
//   cache -> split-KV combine.
// A thread-block cluster of `S` CTAs owns one (sample, kv head): CTA r covers key blocks [r*bps, (r+1)*bps); inside a
// CTA the 4 warps split each 64-key block (16 keys each) since only G <= 8 query rows exist; the CTA whose range holds
// the newest position rotates and stores the new K/V row before its block is staged; partial (m, l, O) are merged
// first across warps in shared memory, then across the cluster through distributed shared memory (no global round
// trip, no second kernel).  Numerics identical to rope_append_kernel + attn_fwd_kernel (und mode, R4-R6).
namespace cg = cooperative_groups;
constexpr int kDecQBytes = 16 * 256;
constexpr int kDecSmemBytes = kDecQBytes + 4 * kTileKeys * 256;

constexpr int kDecThreads = 256;   // warps 0-3: one 16-key quarter of every block each; all 8 warps: rope rows + staging
__global__ void __launch_bounds__(kDecThreads, 2) attn_decode_kernel(DecodeAttnArgs a, float scale_log2) {
    constexpr int HD = 128;
    pdl_launch_dependents();
    pdl_wait();
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = smem + kDecQBytes;
    uint8_t* sV = sK + 2 * kTileKeys * 256;
    __shared__ float red_m[4][8], red_l[4][8];
    __shared__ float red_o[4][8][HD];
    __shared__ float fin_o[8][HD];
    __shared__ float fin_m[8], fin_l[8];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int b = blockIdx.y / a.Hkv, kvh = blockIdx.y % a.Hkv;
    const int G = a.H / a.Hkv;
    const int ncols = (a.H + 2 * a.Hkv) * HD;
    const int kvlen = a.kv_len[b];
    const int blocks_total = (kvlen + kTileKeys - 1) / kTileKeys;
    const int bps = (blocks_total + S - 1) / S;
    const int kb_begin = rank * bps, kb_end = min(blocks_total, kb_begin + bps);
    const int last_block = (kvlen - 1) / kTileKeys;
    const bool owner = kb_begin <= last_block && last_block < kb_end;

    auto load_kv = [&](int kb, int stage) {
        const int page = a.page_table[(size_t)b * a.max_pages + kb];
        const bf16* kbase = a.pool.base + a.pool.tile_offset(page, a.layer, 0, kvh);
        const bf16* vbase = a.pool.base + a.pool.tile_offset(page, a.layer, 1, kvh);
        uint8_t* dk = sK + stage * kTileKeys * 256;
        uint8_t* dv = sV + stage * kTileKeys * 256;
        for (int i = tid; i < kTileKeys * 16; i += kDecThreads) {
            const int r = i >> 4, ch = i & 15;
            const bool ok = kb * kTileKeys + r < kvlen;
            cp_async16(dk + tile_off<HD>(r, ch), kbase + (size_t)(ok ? r : 0) * HD + ch * 8, ok);
            cp_async16(dv + tile_off<HD>(r, ch), vbase + (size_t)(ok ? r : 0) * HD + ch * 8, ok);
        }
    };
    // stage the first block right away unless it is the one that receives the new token
    const bool pre = kb_begin < kb_end && !(owner && kb_begin == last_block);
    if (pre) load_kv(kb_begin, 0);
    cp_async_commit();

    // ---- rows of this step: G query heads (-> sQ) and, on the owner CTA, the new K and V rows (-> page)
    for (int i = tid; i < 16 * 16; i += kDecThreads) {
        const int r = i >> 4, ch = i & 15;
        if (r >= G) *reinterpret_cast<U4*>(sQ + tile_off<HD>(r, ch)) = U4{0, 0, 0, 0};
    }
    const float pos = (float)a.positions[b];
    const int n_tasks = G + (owner ? 2 : 0);
    for (int task = warp; task < n_tasks; task += kDecThreads / 32) {
        const bool is_q = task < G, is_v = task == G + 1;
        const int col = (is_q ? (kvh * G + task) : (is_v ? a.H + a.Hkv + kvh : a.H + kvh)) * HD + lane * 4;
        float x[4];
        if (a.partial) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int sp = 0; sp < a.ksplits; ++sp) {
                const float4 p4 = *reinterpret_cast<const float4*>(a.partial + ((size_t)sp * a.M + b) * ncols + col);
                acc[0] += p4.x; acc[1] += p4.y; acc[2] += p4.z; acc[3] += p4.w;
            }
            const uint2 bv = *reinterpret_cast<const uint2*>(a.bias + col);
            const float2 b0 = unpack2(bv.x), b1 = unpack2(bv.y);
            x[0] = rbf(acc[0] + b0.x); x[1] = rbf(acc[1] + b0.y); x[2] = rbf(acc[2] + b1.x); x[3] = rbf(acc[3] + b1.y);
        } else {
            const uint2 qv = *reinterpret_cast<const uint2*>(a.qkv + (size_t)b * ncols + col);
            const float2 f0 = unpack2(qv.x), f1 = unpack2(qv.y);
            x[0] = f0.x; x[1] = f0.y; x[2] = f1.x; x[3] = f1.y;
        }
        float o4[4];
        if (is_v) {
#pragma unroll
            for (int j = 0; j < 4; ++j) o4[j] = x[j];
        } else {
            const bf16* nw = is_q ? a.qn : a.kn;
            float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
            ss = warp_sum(ss);
            const float inv = 1.0f / sqrtf(ss / (float)HD + a.eps);
            const uint2 wv = *reinterpret_cast<const uint2*>(nw + lane * 4);
            const float2 w0 = unpack2(wv.x), w1 = unpack2(wv.y);
            const float w[4] = {w0.x, w0.y, w1.x, w1.y};
            float n[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) n[j] = rbf(w[j] * rbf(x[j] * inv));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = lane * 4 + j;
                const float ang = __fmul_rn(pos, a.inv_freq[i % (HD / 2)]);
                const float c = rbf(cosf(ang)), sn = rbf(sinf(ang));
                const float partner = __shfl_xor_sync(0xffffffffu, n[j], 16);
                const float rot = (i < HD / 2) ? -partner : partner;
                o4[j] = rbf(rbf(n[j] * c) + rbf(rot * sn));
            }
        }
        const uint2 packed = make_uint2(pack2(o4[0], o4[1]), pack2(o4[2], o4[3]));
        if (is_q) {
            *reinterpret_cast<uint2*>(sQ + tile_off<HD>(task, lane >> 1) + (lane & 1) * 8) = packed;
        } else {
            const int slot = (kvlen - 1) % kTileKeys;
            const int page = a.page_table[(size_t)b * a.max_pages + last_block];
            bf16* dst = a.pool.base + a.pool.tile_offset(page, a.layer, is_v ? 1 : 0, kvh) + (size_t)slot * HD + lane * 4;
            *reinterpret_cast<uint2*>(dst) = packed;
        }
    }
    __syncthreads();                       // sQ complete; the owner's K/V row is visible to the CTA's later cp.async
    if (!pre && kb_begin < kb_end) load_kv(kb_begin, 0);
    cp_async_commit();

    const int g = lane >> 2, t = lane & 3;
    uint32_t qf[8][4];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
        ldmatrix_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], smem_u32(sQ + tile_off<HD>(lane & 15, ks * 2 + (lane >> 4))));
    float o[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = 0.f;
    float m_a = -INFINITY, l_a = 0.f;

    for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int stage = (kb - kb_begin) & 1;
        if (kb + 1 < kb_end) {
            load_kv(kb + 1, stage ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const uint8_t* cK = sK + stage * kTileKeys * 256;
        const uint8_t* cV = sV + stage * kTileKeys * 256;
        if (warp < 4) {
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4(b0, b1, b2, b3,
                        smem_u32(cK + tile_off<HD>(warp * 16 + (lane & 7) + ((lane >> 4) << 3), ks * 2 + ((lane >> 3) & 1))));
            mma_bf16_16816(s0, qf[ks], b0, b1);
            mma_bf16_16816(s1, qf[ks], b2, b3);
        }
        const int key0 = kb * kTileKeys + warp * 16 + 2 * t;
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
                         : "r"(smem_u32(cV + tile_off<HD>(warp * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), dp + (lane >> 4)))));
            float acc0[4] = {o[dp][0], o[dp][1], 0.f, 0.f}, acc1[4] = {o[dp + 1][0], o[dp + 1][1], 0.f, 0.f};
            mma_bf16_16816(acc0, pf, b0, b1);
            mma_bf16_16816(acc1, pf, b2, b3);
            o[dp][0] = acc0[0]; o[dp][1] = acc0[1];
            o[dp + 1][0] = acc1[0]; o[dp + 1][1] = acc1[1];
        }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
    l_a += __shfl_xor_sync(0xffffffffu, l_a, 1);
    l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);

    // ---- merge the 4 key-quarters of this CTA
    if (warp < 4) {
        if (t == 0) { red_m[warp][g] = m_a; red_l[warp][g] = l_a; }
#pragma unroll
        for (int dt = 0; dt < 16; ++dt) *reinterpret_cast<float2*>(&red_o[warp][g][dt * 8 + 2 * t]) = make_float2(o[dt][0], o[dt][1]);
    }
    __syncthreads();
    if (tid < 128) {
        const int row = tid >> 4, d0 = (tid & 15) * 8;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < 4; ++w) M = fmaxf(M, red_m[w][row]);
        float L = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const float mw = red_m[w][row];
            const float wg = (mw == -INFINITY) ? 0.f : exp2f(mw - M);
            L += wg * red_l[w][row];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += wg * red_o[w][row][d0 + j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) fin_o[row][d0 + j] = acc[j];
        if ((tid & 15) == 0) { fin_m[row] = M; fin_l[row] = L; }
    }
    cluster.sync();
    // ---- merge the key ranges of the cluster: CTA `rank` finishes heads rank, rank+S, ...
    for (int head = rank; head < G && tid < HD; head += S) {
        float M = -INFINITY;
        for (int r = 0; r < S; ++r) M = fmaxf(M, cluster.map_shared_rank(fin_m, r)[head]);
        float L = 0.f, acc = 0.f;
        for (int r = 0; r < S; ++r) {                      // fixed order: deterministic
            const float mr = cluster.map_shared_rank(fin_m, r)[head];
            const float wg = (mr == -INFINITY) ? 0.f : exp2f(mr - M);
            L += wg * cluster.map_shared_rank(fin_l, r)[head];
            acc += wg * cluster.map_shared_rank(&fin_o[0][0], r)[head * HD + tid];
        }
        a.out[(size_t)b * a.ldo + (kvh * G + head) * HD + tid] = f2b(L > 0.f ? acc / L : 0.f);
    }
    cluster.sync();                        // peers may still be reading this CTA's shared memory
}

int decode_attention(const DecodeAttnArgs& a, cudaStream_t s) {
    UMV_REQUIRE(a.H % a.Hkv == 0 && a.H / a.Hkv <= 8, UMV_ERR_UNSUPPORTED, "decode_attention: GQA group %d > 8", a.H / a.Hkv);
    UMV_REQUIRE(a.cluster >= 1 && a.cluster <= 8, UMV_ERR_INVALID, "decode_attention: cluster size %d", a.cluster);
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
    cudaError_t e = cudaLaunchKernelEx(&cfg, attn_decode_kernel, a, scale_log2);
    ++g_launches;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("attn_decode_kernel launch failed: %s", cudaGetErrorString(e));
        return UMV_ERR_CUDA;
    }
    return UMV_OK;
}

int attention_init() {
    cudaFuncSetAttribute(attn_fwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<128>::kSmemBytes);
    cudaFuncSetAttribute(attn_fwd_kernel<72>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<72>::kSmemBytes);
    cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDecSmemBytes);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("attention_init: %s", cudaGetErrorString(e));
        return UMV_ERR_CUDA;
    }
    return UMV_OK;
}

template <int HD>
static int launch_attn(const AttnArgs& a, cudaStream_t s) {
    const int G = a.H / a.Hkv;
    const int row_tiles = (a.max_q_len * G + kTileRows - 1) / kTileRows;
    const float scale_log2 = (1.0f / sqrtf((float)HD)) * 1.4426950408889634f;
    dim3 grid(row_tiles, a.n * a.Hkv, a.splits);
    cudaError_t e = launch_k(attn_fwd_kernel<HD>, grid, dim3(kAttnThreads), AttnCfg<HD>::kSmemBytes, s, a, scale_log2);
    ++g_launches;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("attn_fwd_kernel<%d> launch failed: %s", HD, cudaGetErrorString(e));
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
    if (a.dh == 128) return launch_attn<128>(a, s);
    if (a.dh == 72) return launch_attn<72>(a, s);
    set_error("attention: head_dim %d is not built (128 and 72 are)", a.dh);
    return UMV_ERR_UNSUPPORTED;
}

}  // namespace umv
