// Prefill / flow / ViT attention on the 5th-generation tensor cores (head_dim 128; paged KV, or packed K / V matrices whose
// heads are zero-padded to 128 columns -- the ViT's 72-wide heads: the kernel is paced by the softmax pipe, not by the
// MMAs, so the padded contraction is free).
//
// One CTA owns 128 query rows of ONE kv head -- rows are (token, head-in-group) pairs, 18 tokens x 7 heads at the 14B dims,
// so a K/V block is fetched once for the whole GQA group -- and walks the visible keys in blocks of 128:
//
//   warp 0   TMA: the Q tile (3-D box: 64 columns x G heads x TOK tokens lands as [row][128 B], 128-byte swizzle) and a
//            two-stage ring of K / V blocks (two 64-slot KV pages each, 64 x 64 boxes straight out of the paged pool)
//   warp 1   one elected thread issues tcgen05.mma: S = Q K^T (M = N = K-extent 128, both operands K-major) into one of two
//            TMEM score buffers, then O += P V (P K-major from shared memory, V MN-major exactly as it lies in the pool)
//   warps 2-17 softmax, four threads per query row (32 keys each): tcgen05.ld of the scores, scale + mask, running max /
//            sum in registers (no shuffles; the parts exchange one maximum per block through shared memory), P rounded to
//            bf16 into the swizzled A-operand tile, O rescaled in TMEM (tcgen05.ld / st) only when the row maximum moved
//            by more than 2^8 (lazy rescaling); S(j+1) is computed while softmax(j) runs.
//
// Numerics: those of attn_fwd_kernel / flash-attn 2 (fp32 scores and statistics, P rounded to bf16 before P V, fp32 O) with
// 128-key instead of 64-key rescaling steps.  Replaces flash_attn_varlen_func at qwen2_navit.py:605-614.
#include <cuda.h>

#include "../../include/umv.h"
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace umv {

namespace {

constexpr int HD = 128;
constexpr int TQ = 128;                 // query rows per CTA (UMMA M)
constexpr int kHalf = 128 * 64 * 2;     // a [128 x 64] bf16 half tile (Q, P)
constexpr int kTile = 2 * kHalf;        // [128 x 128]
constexpr int kSplit = 2;               // softmax threads per query row (4: measured slower, 290 vs 239 us per 14B prefill layer)
constexpr int kOCols = HD / kSplit;     // O columns per softmax thread
constexpr int kTcThreads = 64 + kSplit * 128;      // TMA warp, MMA warp, 4 * kSplit softmax warps
constexpr int kPairedMaxKeys = 512;     // key ranges up to this may use the 64-key, two-CTAs-per-SM build (attention_tc_forward)

// Two builds of the kernel.  TK = 128 keys per block: one CTA per SM (193 KB of shared memory, 512 TMEM columns), the fewest
// per-key overheads -- long key ranges.  TK = 64: half-size K / V / P stages, a single V stage and 256 TMEM columns, so TWO CTAs
// share an SM (100 KB each, <= 96 registers): a tile over a short key range is a chain of dependent latencies (Q / K load, first
// S, softmax, P V, store: ~13 us for 3 blocks at the flow step, 4 of them arithmetic), and the second CTA fills the other's bubbles.
template <int TK_>
struct TcAttn {
    static constexpr int TK = TK_;                       // keys per block (UMMA N of the score product, K extent of P V)
    static constexpr int kKeys = TK / kSplit;            // keys per softmax thread and block
    static constexpr int kKVHalf = TK * 64 * 2;          // [TK keys x 64 columns]
    static constexpr int kKVTile = 2 * kKVHalf;          // [TK x 128]
    static constexpr int kPTile = (TK / 64) * kHalf;     // [128 rows x TK keys]
    static constexpr int kVStages = TK == 128 ? 2 : 1;
    static constexpr int kSmemBytes = 1024 + kTile /*Q*/ + 2 * kKVTile /*K*/ + kVStages * kKVTile /*V*/ + kPTile + 256;
    static constexpr int kColS = 0, kColO = 2 * TK;      // TMEM columns: S0 [0, TK), S1 [TK, 2 TK), O [2 TK, 2 TK + 128)
    static constexpr int kTmemCols = TK == 128 ? 512 : 256;
    static constexpr int kCtasPerSm = TK == 128 ? 1 : 2;
    static_assert(kKeys % 32 == 0, "whole 32-column TMEM loads per thread");
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int TK>
__global__ void __launch_bounds__(kTcThreads, TcAttn<TK>::kCtasPerSm)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, AttnArgs a, int G, int TOK, float scale_log2) {
    using C = TcAttn<TK>;
    constexpr int kKeys = C::kKeys, kKVHalf = C::kKVHalf, kKVTile = C::kKVTile, kVStages = C::kVStages, kColS = C::kColS, kColO = C::kColO;
    pdl_launch_dependents();
    trace_start(a.trace);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + kTile;
    uint8_t* sV = sK + 2 * kKVTile;
    uint8_t* sP = sV + kVStages * kKVTile;
    uint64_t* q_full = reinterpret_cast<uint64_t*>(sP + C::kPTile);
    uint64_t* k_full = q_full + 1;       // [2]  K and V have separate rings: a K stage is free as soon as its score
    uint64_t* k_empty = k_full + 2;      // [2]  product has completed, one whole block before P V releases the V stage --
    uint64_t* v_full = k_empty + 2;      // [2]  so the next K block is requested a block period ahead of its use
    uint64_t* v_empty = v_full + 2;      // [2]
    uint64_t* s_full = v_empty + 2;      // [2]
    uint64_t* p_ready = s_full + 2;
    uint64_t* pv_done = p_ready + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);
    __shared__ float s_max[2][kSplit][TQ];     // per-block row maxima of the key parts (parity double-buffered)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y / a.Hkv, kvh = blockIdx.y % a.Hkv;
    const int t0 = blockIdx.x * TOK;
    pdl_wait();
    trace_wait(a.trace);
    const int qlen = a.q_len[b], kvlen = a.kv_len[b];
    if (t0 >= qlen) return;                                   // whole CTA, before any barrier / TMEM state exists
    const int qs = a.q_start[b];
    const int ntok = min(TOK, qlen - t0);
    const int vis = a.causal ? min(kvlen, kvlen - qlen + t0 + ntok) : kvlen;      // keys the tile's last token sees
    const int nkb = (vis + TK - 1) / TK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == 1) {
        if (lane == 0) {
            mbar_init(q_full, 1);
            for (int i = 0; i < 2; ++i) {
                mbar_init(&k_full[i], 1);
                mbar_init(&k_empty[i], 1);
                mbar_init(&v_full[i], 1);
                mbar_init(&v_empty[i], 1);
                mbar_init(&s_full[i], 1);
            }
            mbar_init(p_ready, kSplit * 128);
            mbar_init(pv_done, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<C::kTmemCols>(tmem_slot);
    }
    if (warp >= 2) {
        // rows of the Q tile the TMA box does not cover (TOK * G .. 127): keep them finite
        const int e = (warp - 2) * 32 + lane;
        for (int i = e; i < (TQ - TOK * G) * 16; i += kSplit * 128) {
            const int r = TOK * G + (i >> 4), ch = i & 15;
            *reinterpret_cast<U4*>(sQ + (ch >> 3) * kHalf + r * 128 + (((ch & 7) ^ (r & 7)) << 4)) = U4{0, 0, 0, 0};
        }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA loads
        if (elect_one()) {
            mbar_expect_tx(q_full, (uint32_t)(2 * 64 * G * TOK * 2));
            tma_load_3d(sQ, &tmQ, q_full, 0, kvh * G, qs + t0);
            tma_load_3d(sQ + kHalf, &tmQ, q_full, 64, kvh * G, qs + t0);
            constexpr int kPages = TK / 64;                  // 64-slot KV pages per key block
            for (int j = 0; j < nkb; ++j) {
                const int stage = j & 1, vstage = j % kVStages;
                int row_k[kPages], row_v[kPages];
                int col0 = 0;                      // column of the kv head's first element in the K / V matrices
#pragma unroll
                for (int pg = 0; pg < kPages; ++pg) {
                    const int kb64 = kPages * j + pg;
                    if (a.paged) {
                        const int page = kb64 < a.max_pages ? a.page_table[(size_t)b * a.max_pages + kb64] : 0;
                        row_k[pg] = (int)(a.pool.tile_offset(page, a.layer, 0, kvh) / HD);
                        row_v[pg] = (int)(a.pool.tile_offset(page, a.layer, 1, kvh) / HD);
                    } else {                       // packed [Tk, Hkv * 128] matrices: rows past the sample's keys are masked
                        row_k[pg] = row_v[pg] = a.k_start[b] + kb64 * 64;
                        col0 = kvh * HD;
                    }
                }
                mbar_wait(&k_empty[stage], ((j >> 1) & 1) ^ 1u);
                mbar_expect_tx(&k_full[stage], kKVTile);
#pragma unroll
                for (int pg = 0; pg < kPages; ++pg)
#pragma unroll
                    for (int half = 0; half < 2; ++half)
                        tma_load_2d(sK + stage * kKVTile + half * kKVHalf + pg * 8192, &tmK, &k_full[stage], col0 + half * 64, row_k[pg], kEvictNormal);
                mbar_wait(&v_empty[vstage], ((j / kVStages) & 1) ^ 1u);
                mbar_expect_tx(&v_full[vstage], kKVTile);
#pragma unroll
                for (int pg = 0; pg < kPages; ++pg)
#pragma unroll
                    for (int half = 0; half < 2; ++half)
                        tma_load_2d(sV + vstage * kKVTile + half * kKVHalf + pg * 8192, &tmV, &v_full[vstage], col0 + half * 64, row_v[pg], kEvictNormal);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (elect_one()) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(TK);       // M 128, N = TK
            constexpr uint32_t idesc_pv = umma_idesc_bf16(HD) | (1u << 16);      // B (= V) is MN-major
            mbar_wait(q_full, 0);
            auto issue_s = [&](int j) {
                const int stage = j & 1;
                mbar_wait(&k_full[stage], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t d = tmem_base + kColS + (j & 1) * TK;
                const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + stage * kKVTile);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    umma_bf16(d, umma_desc_sw128(qa + (k >> 2) * kHalf + (k & 3) * 32), umma_desc_sw128(ka + (k >> 2) * kKVHalf + (k & 3) * 32),
                              idesc_s, k > 0 ? 1u : 0u);
                umma_commit(&s_full[j & 1]);
                umma_commit(&k_empty[stage]);
            };
            issue_s(0);
            for (int j = 0; j < nkb; ++j) {
                if (j + 1 < nkb) issue_s(j + 1);             // next scores while the softmax warps work on block j
                const int vstage = j % kVStages;
                mbar_wait(&v_full[vstage], (j / kVStages) & 1);
                mbar_wait(p_ready, j & 1);
                tc_fence_after();
                const uint32_t d = tmem_base + kColO;
                const uint32_t pa = smem_u32(sP), va = smem_u32(sV + vstage * kKVTile);
#pragma unroll
                for (int k = 0; k < TK / 16; ++k)
                    umma_bf16(d, umma_desc_sw128(pa + (k >> 2) * kHalf + (k & 3) * 32), umma_desc_sw128_mn(va + k * 2048, kKVHalf), idesc_pv,
                              (j > 0 || k > 0) ? 1u : 0u);
                umma_commit(pv_done);
                umma_commit(&v_empty[vstage]);
            }
        }
    } else {
        // ------------------------------------------------------------ softmax: kSplit threads per query row
        // warps w, w + 4, ... read the same TMEM lane quarter; part p of a row takes keys [p * kKeys, (p + 1) * kKeys) of every
        // block and the same columns of O; the parts exchange one maximum per block.  Measured inside a block (B200, ns):
        // max + exchange 160, 64 exponentials + bf16 packs per thread 990 (the quarter-rate ex2 / cvt pipe is the limit:
        // 16,384 + 8,192 operations per block at 16 per clock), wait for the previous P V 160, P store + fences 220.
        const int quarter = warp & 3, part = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const int tok_l = row / G, head_l = row % G;
        const bool valid = row < TOK * G && tok_l < ntok;
        const int lim = (a.causal && valid) ? (kvlen - qlen + t0 + tok_l) : (kvlen - 1);     // last visible key of the row
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        float m = -INFINITY, l = 0.f;          // l: this thread's part of the row sum (all parts share the maxima)
        uint32_t r[kKeys > kOCols ? kKeys : kOCols];
        for (int j = 0; j < nkb; ++j) {
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t ts = lane_base + kColS + (j & 1) * TK + part * kKeys;
            const int key0 = j * TK + part * kKeys;
            const bool full = key0 + kKeys - 1 <= lim;            // no masking needed for this thread's keys
#pragma unroll
            for (int c = 0; c < kKeys / 32; ++c) tmem_ld32(ts + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
            tmem_ld_wait();
            float mraw = -INFINITY;
            if (full) {
#pragma unroll
                for (int i = 0; i < kKeys; ++i) mraw = fmaxf(mraw, __uint_as_float(r[i]));
            } else {
#pragma unroll
                for (int i = 0; i < kKeys; ++i)
                    if (key0 + i <= lim) mraw = fmaxf(mraw, __uint_as_float(r[i]));
            }
            s_max[j & 1][part][row] = mraw * scale_log2;          // scale > 0: max commutes with it (-inf stays -inf)
            asm volatile("bar.sync 1, %0;" ::"n"(kSplit * 128) : "memory");
            // Lazy rescaling: the reference maximum only follows the running maximum when that moved by more than 2^8.
            // Exponentials, sums and O are then all scaled by the same power of two, which commutes with every rounding on the
            // way, so the result equals the eager scheme bit for bit -- but O is almost never touched and the exponentials of
            // block j do not wait for P V of block j-1.
            float mblk = s_max[j & 1][0][row];
#pragma unroll
            for (int q = 1; q < kSplit; ++q) mblk = fmaxf(mblk, s_max[j & 1][q][row]);
            float alpha = 1.f;
            if (mblk > m + 8.f || m == -INFINITY) {               // (-inf > -inf + 8) is false: the m == -inf test is the first block
                const float m_new = fmaxf(m, mblk);
                alpha = (m == -INFINITY) ? 0.f : ex2_approx(m - m_new);
                m = m_new;
            }
            const float ms = m == -INFINITY ? 0.f : m;
            // probabilities of this thread's keys, packed to bf16 in registers
            float lsum = 0.f;
            uint32_t pk[kKeys / 2];
#pragma unroll
            for (int i = 0; i < kKeys / 2; ++i) {
                float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * i]), scale_log2, -ms));
                float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), scale_log2, -ms));
                if (!full) {
                    if (key0 + 2 * i > lim) p0 = 0.f;
                    if (key0 + 2 * i + 1 > lim) p1 = 0.f;
                }
                lsum += p0 + p1;
                pk[i] = pack2(p0, p1);
            }
            l = l * alpha + lsum;
            // O and the P tile belong to P V of the previous block until it has completed
            if (j > 0) {
                mbar_wait(pv_done, (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
                    for (int c = 0; c < kOCols / 32; ++c) {
                        uint32_t o[32];
                        tmem_ld32(lane_base + kColO + part * kOCols + c * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st32(lane_base + kColO + part * kOCols + c * 32, o);
                    }
                    tmem_st_wait();
                }
            }
            // P -> the swizzled A-operand tile: this thread's keys are chunks part * kKeys / 8 .. of the row
#pragma unroll
            for (int q = 0; q < kKeys / 8; ++q) {
                const int cc = part * (kKeys / 8) + q;
                *reinterpret_cast<U4*>(sP + (cc >> 3) * kHalf + row * 128 + (((cc & 7) ^ (row & 7)) << 4)) =
                    U4{pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]};
            }
            fence_proxy_async_smem();                         // P (generic proxy) -> visible to the tensor core
            tc_fence_before();                                // orders the tcgen05.ld / st above before the arrive
            mbar_arrive(p_ready);
        }
        // ---- epilogue: O / l -> bf16 (each thread its 32 columns)
        s_max[nkb & 1][part][row] = l;
        asm volatile("bar.sync 1, %0;" ::"n"(kSplit * 128) : "memory");
        float lt = 0.f;
#pragma unroll
        for (int q = 0; q < kSplit; ++q) lt += s_max[nkb & 1][q][row];
        mbar_wait(pv_done, (nkb - 1) & 1);
        tc_fence_after();
        const float inv = lt > 0.f ? 1.f / lt : 0.f;
        const int odh = a.out_dh ? a.out_dh : HD;             // real head width of the output (padded-head callers: < 128)
        const int orow = qs + t0 + (valid ? tok_l : 0);
        bf16* dst = a.out + (size_t)(a.out_row_map ? a.out_row_map[orow] : orow) * a.ldo + (kvh * G + (valid ? head_l : 0)) * odh + part * kOCols;
#pragma unroll
        for (int c = 0; c < kOCols / 32; ++c) tmem_ld32(lane_base + kColO + part * kOCols + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
        tmem_ld_wait();
        if (valid) {
#pragma unroll
            for (int q = 0; q < kOCols / 8; ++q) {
                uint32_t o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    o[i] = pack2(__uint_as_float(r[8 * q + 2 * i]) * inv, __uint_as_float(r[8 * q + 2 * i + 1]) * inv);
                if (part * kOCols + q * 8 < odh) stg16(dst + q * 8, U4{o[0], o[1], o[2], o[3]});
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<C::kTmemCols>(tmem_base);
    trace_end(a.trace);
}

}  // namespace

bool attention_tc_supported(const AttnArgs& a) {
    if (!(a.dh == HD && a.splits == 1 && a.Hkv > 0 && a.H % a.Hkv == 0)) return false;
    if (a.paged) {
        if (!a.kv_tmap) return false;
    } else {
        if (!(a.k && a.v && a.k_start && a.total_k > 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 &&
              ((reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v)) & 15) == 0))
            return false;
    }
    const int G = a.H / a.Hkv;
    if (G > 16 || a.max_q_len * G < TQ) return false;            // at least one full tile of rows
    const int odh = a.out_dh ? a.out_dh : HD;
    return (a.ldq % 8 == 0) && (a.ldo % 8 == 0) && (odh % 8 == 0) && odh <= HD && ((reinterpret_cast<uintptr_t>(a.q) & 15) == 0) &&
           ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);
}

template <int TK>
static int launch_attn_tc(const AttnArgs& a, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, int G, int TOK,
                          float scale_log2, cudaStream_t s) {
    using C = TcAttn<TK>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t ae = cudaFuncSetAttribute(attn_tc_kernel<TK>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
        if (ae != cudaSuccess) {
            set_error("attn_tc_kernel<%d>: cudaFuncSetAttribute(max dynamic smem) failed: %s", TK, cudaGetErrorString(ae));
            return UMV_ERR_CUDA;
        }
        attr_set = true;
    }
    dim3 grid((a.max_q_len + TOK - 1) / TOK, a.n * a.Hkv);
    cudaError_t e = launch_k(attn_tc_kernel<TK>, grid, dim3(kTcThreads), C::kSmemBytes, s, tmQ, tmK, tmV, a, G, TOK, scale_log2);
    ++g_launches;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("attn_tc_kernel<%d> launch failed: %s", TK, cudaGetErrorString(e));
        return UMV_ERR_CUDA;
    }
    return UMV_OK;
}

int attention_tc_forward(const AttnArgs& a0, cudaStream_t s) {
    AttnArgs a = a0;
    const int G = a.H / a.Hkv, TOK = TQ / G;
    int rc = gemm_init();
    if (rc) return rc;
    CUtensorMap tmQ, tmK, tmV;
    // q rows viewed as [tokens][heads][128]: a box of 64 columns x G heads x TOK tokens is the tile's [row][128 B] image
    rc = make_tmap_3d(&tmQ, a.q, HD, a.H, a.total_q, (uint64_t)HD * 2, (uint64_t)a.ldq * 2, G, TOK);
    if (rc) return rc;
    if (a.paged) {
        tmK = tmV = *a.kv_tmap;
    } else {
        if ((rc = make_tmap_2d(&tmK, a.k, a.total_k, (uint64_t)a.Hkv * HD, a.ldk, 64))) return rc;
        if ((rc = make_tmap_2d(&tmV, a.v, a.total_k, (uint64_t)a.Hkv * HD, a.ldv, 64))) return rc;
    }
    a.kv_tmap = nullptr;
    const float scale_log2 = (a.scale > 0.f ? a.scale : 1.0f / sqrtf((float)HD)) * 1.4426950408889634f;
    // Key-block size.  Two co-resident CTAs with 64-key blocks when there are tiles enough to double up on every SM and the key
    // range is short -- the flow step: 12 x 258 rows on 292 keys, 49.0 -> 39.6 us for the kernel alone, 65.6 -> 51.5 us per layer inside
    // the step.  One CTA with 128-key blocks otherwise: few tiles (8 x 34 prompt rows on 1,060 keys: 15.9 vs 23.5 us), long ranges
    // (4,096 causal: 191 vs 216 us), and the ~1k-key image / ViT prefill, where the paired build wins alone (227 -> 216 us) but loses
    // inside the job (ViT + image prefill 97.8 -> 102.4 ms: two half-size CTAs per SM delay the next linear's early-launched CTAs,
    // which need the whole SM's shared memory and all 512 TMEM columns).
    // UMV_ATTN_TK = 64 | 128 forces one (read per call: tests switch inside one process).
    const char* tk_env = getenv("UMV_ATTN_TK");
    const long tiles = (long)((a.max_q_len + TOK - 1) / TOK) * a.n * a.Hkv;
    const int tk = tk_env && (atoi(tk_env) == 64 || atoi(tk_env) == 128) ? atoi(tk_env)
                   : (tiles >= 2L * gemm_sm_count() && a.max_kv_len <= kPairedMaxKeys ? 64 : 128);
    a.trace = trace_next(tk == 64 ? "attn_tc64" : "attn_tc");
    return tk == 64 ? launch_attn_tc<64>(a, tmQ, tmK, tmV, G, TOK, scale_log2, s) : launch_attn_tc<128>(a, tmQ, tmK, tmV, G, TOK, scale_log2, s);
}

}  // namespace umv
