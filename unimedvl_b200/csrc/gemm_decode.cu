// Decode linears (M <= 8 tokens) with their neighbours of the dependency chain folded in.
//
// The decode step is a chain of small kernels in which every link costs a few microseconds of pure latency
// (profiles/r1_decode_timeline.md): norm -> q/k/v -> attention -> o_proj -> norm -> gate/up -> down.  This file removes
// the two norm links per layer:
//
//   * NORMB: the consumer linear (q/k/v, gate/up, lm_head) builds its activation operand itself.  Its four epilogue warps
//     read the residual stream h [M, K] (57 KB at K = 3584), apply Qwen2RMSNorm (modeling_qwen2.py:89-94: fp32 statistics,
//     bf16 rounding before and after the weight) and leave the K-major, 128-byte-swizzled UMMA operand for ALL k-blocks
//     resident in shared memory (1 KB per k-block: 8 token rows; the MMA's N = 16 reads a second 8-row group that aliases
//     the next k-block's tile and lands in accumulator columns nobody reads).  Meanwhile the TMA producer streams
//     weights from the first instruction on -- weights are constants, so it never waits for the predecessor.
//   * CLUSTER_RESID: the producer linear (o_proj, down_proj) splits K over a thread-block cluster of 4 CTAs (112 CTAs
//     stream at full HBM rate), reduces the four fp32 partial tiles through distributed shared memory in a fixed order
//     and writes h = bf16(h + bf16(sum)) itself -- no fp32 partials in global memory, no reduce kernel.
//
// Rounding points are those of gemm.cu + add_rmsnorm (R1-R3, R7); only fp32 summation orders differ.
// Replaces, for decode, qwen2_navit.py:541-543,617-620 (q/k/v, o_proj), modeling_qwen2.py:89-94,234-235 (norms, MLP),
// bagel.py:1295 (lm_head).
#include <cuda.h>

#include "../../include/umv.h"
#include "common.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace umv {

namespace {

constexpr int DBM = 128, DBK = 64, DBN = 16;
constexpr int kATile = DBM * DBK * 2;          // 16 KB of weights per stage
constexpr int kBTileTma = DBN * DBK * 2;       // 2 KB activation tile (TMA path)
constexpr int kBTileNorm = 1024;               // 8 token rows x 128 B (resident operand)
constexpr int kMaxStages = 10;
constexpr int kBarBytes = 256;
constexpr int kTailBytes = 4096;               // swiglu exchange (8 x 64 fp32) / cluster partial tile (8 x 128 fp32)
constexpr int kStaticReserve = 1024;           // static shared memory (s_ss) counts against the same 227 KB
constexpr int kSmemLimit = 227 * 1024;
constexpr int kNormWarps = 16;                 // NORMB: warps 2..17 build the operand (4 of them are the epilogue warps)
constexpr int kNormThreadsB = kNormWarps * 32;
constexpr int kNormMaxK = kNormThreadsB * 8;   // one 8-element chunk per builder thread and token

struct DecGemmParams {
    int N, K, a_tiles, splits, kb_total, kb_per_split, tokens;
    bf16* y;
    int ldy;
    const bf16* bias;
    float* ws;
    const bf16* nh;      // NORMB: residual stream [tokens, K]
    const bf16* nw;      // NORMB: norm weight [K]
    float eps;
    int stages;
    TraceSlot* trace;
};

__device__ __forceinline__ void cluster_arrive_rel() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acq() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ float4 ld_cluster_f4(const void* smem_ptr, uint32_t rank) {
    uint32_t addr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(smem_u32(smem_ptr)), "r"(rank));
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// MODE 0: y = bf16(acc + bias)   1: fp32 split-K partials   2: SwiGLU (weights interleaved [64 gate | 64 up])
//      3: cluster split-K, y = bf16(y + bf16(sum))  (y is the residual stream, updated in place)
template <bool NORMB>
constexpr int dec_threads() { return NORMB ? 64 + kNormThreadsB : 192; }

template <bool NORMB, int MODE>
__global__ void __launch_bounds__(dec_threads<NORMB>(), 1)
gemm_decode_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const DecGemmParams p) {
    const int kStages = p.stages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sB = smem + kStages * kATile;
    const int b_bytes = NORMB ? (p.kb_total + 1) * kBTileNorm : kStages * kBTileTma;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + b_bytes);
    uint64_t* empty = full + kMaxStages;
    uint64_t* tfull = empty + kMaxStages;
    uint64_t* tempty = tfull + 2;
    uint64_t* bready = tempty + 2;
    uint64_t* hraw = bready + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hraw + 1);
    float* sTail = reinterpret_cast<float*>(sB + b_bytes + kBarBytes);
    __shared__ float s_ss[kNormWarps][8];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    trace_start(p.trace);
    if (MODE == 3) cluster_arrive_rel();          // phase 1: this CTA runs (peers read its shared memory later)
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
                mbar_init(&tempty[i], 4);
            }
            mbar_init(bready, kNormThreadsB);
            mbar_init(hraw, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<32>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int num_tiles = p.a_tiles * p.splits;
    uint2 res_keep = make_uint2(0u, 0u);           // MODE 3: residual values of this thread's reduce slot

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            if (NORMB) {
                // weights only, and weights are constants: stream from the first instruction, never wait for the predecessor
                for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                    const int split = tile % p.splits, a_tile = tile / p.splits;
                    const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1u);
                        mbar_expect_tx(&full[stage], kATile);
                        tma_load_2d(sA + stage * kATile, &tmA, &full[stage], kb * DBK, a_tile * DBM, kEvictFirst);
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
                }
            } else {
                // first ring of weight tiles before griddepcontrol.wait, their activation tiles right after it
                bool waited = false;
                int n_deferred = 0;
                int def_kb[kMaxStages];
                for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                    const int split = tile % p.splits, a_tile = tile / p.splits;
                    const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        if (!waited && n_deferred == kStages) {
                            pdl_wait();
                            waited = true;
                            for (int i = 0; i < n_deferred; ++i)
                                tma_load_2d(sB + i * kBTileTma, &tmB, &full[i], def_kb[i] * DBK, 0, kEvictLast);
                        }
                        mbar_wait(&empty[stage], phase ^ 1u);
                        mbar_expect_tx(&full[stage], kATile + kBTileTma);
                        tma_load_2d(sA + stage * kATile, &tmA, &full[stage], kb * DBK, a_tile * DBM, kEvictFirst);
                        if (waited) tma_load_2d(sB + stage * kBTileTma, &tmB, &full[stage], kb * DBK, 0, kEvictLast);
                        else def_kb[n_deferred++] = kb;
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
                }
                if (!waited) {
                    pdl_wait();
                    for (int i = 0; i < n_deferred; ++i)
                        tma_load_2d(sB + i * kBTileTma, &tmB, &full[i], def_kb[i] * DBK, 0, kEvictLast);
                }
            }
            if (p.trace && blockIdx.x == 0) p.trace->t_wait = gtime();
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(DBN);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            if (NORMB) {
                mbar_wait(bready, 0);
                tc_fence_after();
            }
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int split = tile % p.splits;
                const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                mbar_wait(&tempty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * DBN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * kATile);
                    const uint32_t b_addr = NORMB ? smem_u32(sB + kb * kBTileNorm) : smem_u32(sB + stage * kBTileTma);
#pragma unroll
                    for (int k = 0; k < DBK / 16; ++k)
                        umma_bf16(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                                  (kb > kb0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(&tfull[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------ operand builder (NORMB) + epilogue, 4 warps
        const int quarter = warp & 3;
        const int lrow = quarter * 32 + lane;
        const int ethr = (warp - 2) * 32 + lane;          // 0..127 on the epilogue warps, up to 511 on the builder warps
        if (NORMB) {
            // 16 warps, one 16-byte chunk (8 columns) of every token per thread: with only the 4 epilogue warps the
            // 28,672 conversions per CTA take ~10 us (one warp per scheduler, dependent bf16 rounding chains)
            const int nch = p.K / 8;                      // 16-byte chunks per row
            const int c = ethr;
            const bool mine = c < nch;
            U4 wreg = {0, 0, 0, 0};                       // norm-weight chunk (a constant: before the wait)
            if (mine) wreg = ldg16(p.nw + c * 8);
            // The residual rows arrive by TMA (box = 64 columns x 8 rows, 128-byte swizzle): every k-block's 1 KB tile lands
            // raw in its final place, all requests in flight together -- one memory round trip even though the weight
            // stream of this very CTA keeps the memory system saturated.
            if (ethr == 0) {
                pdl_wait();
                if (p.trace && blockIdx.x == 0) p.trace->dbg[0] = gtime();
                mbar_expect_tx(hraw, (uint32_t)p.kb_total * kBTileNorm);
                for (int kb = 0; kb < p.kb_total; ++kb) tma_load_2d(sB + kb * kBTileNorm, &tmB, hraw, kb * DBK, 0, kEvictNormal);
            }
            mbar_wait(hraw, 0);
            if (p.trace && blockIdx.x == 0 && ethr == 0) p.trace->dbg[2] = gtime();
            U4 raw[8];
            float ss[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                ss[t] = 0.f;
                raw[t] = U4{0, 0, 0, 0};
                if (mine) raw[t] = *reinterpret_cast<const U4*>(sB + (c >> 3) * kBTileNorm + t * 128 + (((c & 7) ^ t) << 4));
                const uint32_t* hw = &raw[t].x;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f = unpack2(hw[q]);
                    ss[t] += f.x * f.x + f.y * f.y;
                }
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) ss[t] = warp_sum(ss[t]);
            if (lane == 0) {
#pragma unroll
                for (int t = 0; t < 8; ++t) s_ss[warp - 2][t] = ss[t];
            }
            asm volatile("bar.sync 2, %0;" ::"n"(kNormThreadsB) : "memory");
            const uint32_t* ww = &wreg.x;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                float tot = 0.f;
#pragma unroll
                for (int w = 0; w < kNormWarps; ++w) tot += s_ss[w][t];
                const float inv = 1.0f / sqrtf(tot / (float)p.K + p.eps);
                if (mine) {                                 // R1/R2: bf16(w * bf16(h * inv)), in place
                    const uint32_t* hw = &raw[t].x;
                    uint32_t o[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 f = unpack2(hw[q]), wf = unpack2(ww[q]);
                        o[q] = pack2(wf.x * rbf(f.x * inv), wf.y * rbf(f.y * inv));
                    }
                    *reinterpret_cast<U4*>(sB + (c >> 3) * kBTileNorm + t * 128 + (((c & 7) ^ t) << 4)) = U4{o[0], o[1], o[2], o[3]};
                }
            }
            if (ethr < 64) *reinterpret_cast<U4*>(sB + p.kb_total * kBTileNorm + ethr * 16) = U4{0, 0, 0, 0};   // pad tile
            fence_proxy_async_smem();                     // generic-proxy writes -> visible to the tensor core (async proxy)
            mbar_arrive(bready);
            if (p.trace && blockIdx.x == 0 && ethr == 0) p.trace->dbg[1] = gtime();
        } else {
            pdl_wait();
        }
        // MODE 3: this thread's residual values, fetched now -- by the time the reduce needs them the next kernel's weight
        // stream already saturates the memory system and a dependent load would wait behind it
        uint2 res_early = make_uint2(0u, 0u);
        if (MODE == 3) {
            uint32_t S, rank;
            asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(S));
            asm volatile("mov.u32 %0, %%cluster_ctaid.x;" : "=r"(rank));
            const int fpr = DBM / (int)S;
            if (ethr < p.tokens * (fpr / 4)) {
                const int t = ethr / (fpr / 4), f = (blockIdx.x / p.splits) * DBM + (int)rank * fpr + (ethr % (fpr / 4)) * 4;
                if (f < p.N) res_early = *reinterpret_cast<const uint2*>(p.y + (size_t)t * p.ldy + f);
            }
            res_keep = res_early;
        }

        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; warp < 6 && tile < num_tiles; tile += gridDim.x) {
            const int split = tile % p.splits, a_tile = tile / p.splits;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * DBN;
            uint32_t r[16];
            tmem_ld16(taddr, r);
            tmem_ld_wait();
            if (MODE == 0 || MODE == 1) {
                const int f = a_tile * DBM + lrow;
                if (f < p.N) {
                    const float bias = (MODE == 0 && p.bias) ? b2f(p.bias[f]) : 0.f;
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        if (t < p.tokens) {
                            const float a = __uint_as_float(r[t]);
                            if (MODE == 1) p.ws[((size_t)split * p.tokens + t) * p.N + f] = a;
                            else p.y[(size_t)t * p.ldy + f] = f2b(rbf(a + bias));
                        }
                    }
                }
            } else if (MODE == 2) {
                // lanes 0..63 hold gate rows, 64..127 the matching up rows
                const bool is_up = lrow >= 64;
                const int jl = lrow & 63;
                const int jglob = a_tile * 64 + jl;
                if (is_up) {
#pragma unroll
                    for (int t = 0; t < 8; ++t) sTail[t * 64 + jl] = rbf(__uint_as_float(r[t]));
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (!is_up && jglob < p.N / 2) {
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        if (t < p.tokens) {
                            const float gv = rbf(silu_f(rbf(__uint_as_float(r[t]))));
                            p.y[(size_t)t * p.ldy + jglob] = f2b(gv * sTail[t * 64 + jl]);
                        }
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            } else {
                // cluster split-K: park this CTA's fp32 partial tile [token][feature] for the cluster reduce below
#pragma unroll
                for (int t = 0; t < 8; ++t) sTail[t * DBM + lrow] = __uint_as_float(r[t]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }

    if (MODE == 3) {
        // ---- reduce the cluster's partial tiles through distributed shared memory, add the residual, write h
        uint32_t S, rank;
        asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(S));
        asm volatile("mov.u32 %0, %%cluster_ctaid.x;" : "=r"(rank));
        cluster_wait_acq();                        // phase 1
        cluster_arrive_rel();                      // phase 2: partial tiles parked ...
        cluster_wait_acq();                        // ... everywhere
        if (warp >= 2 && warp < 6) {
            const int ethr = (warp - 2) * 32 + lane;
            const int fpr = DBM / (int)S;          // features finished by this CTA
            const int a_tile = blockIdx.x / p.splits;
            const int idx = ethr;                  // tokens (<= 8) x fpr / 4 (<= 16) slots: at most one per thread
            if (idx < p.tokens * (fpr / 4)) {
                const int t = idx / (fpr / 4), fl = (int)rank * fpr + (idx % (fpr / 4)) * 4;
                const int f = a_tile * DBM + fl;
                if (f < p.N) {
                    float4 part[8];
#pragma unroll
                    for (int s = 0; s < 8; ++s)
                        if (s < (int)S) part[s] = ld_cluster_f4(&sTail[t * DBM + fl], s);
                    float acc4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int s = 0; s < 8; ++s) {          // split order: deterministic
                        if (s < (int)S) { acc4[0] += part[s].x; acc4[1] += part[s].y; acc4[2] += part[s].z; acc4[3] += part[s].w; }
                    }
                    const float2 h0 = unpack2(res_keep.x), h1 = unpack2(res_keep.y);
                    // the Linear's own bf16 output (R3), then the residual add in bf16 (R7)
                    *reinterpret_cast<uint2*>(p.y + (size_t)t * p.ldy + f) =
                        make_uint2(pack2(h0.x + rbf(acc4[0]), h0.y + rbf(acc4[1])), pack2(h1.x + rbf(acc4[2]), h1.y + rbf(acc4[3])));
                }
            }
        }
        cluster_arrive_rel();                      // phase 3: peers are done reading this CTA's shared memory
        cluster_wait_acq();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<32>(tmem_base);
    trace_end(p.trace);
}

template <bool NORMB, int MODE>
int launch_decode(const DecodeLinear& c, cudaStream_t stream) {
    DecGemmParams p{};
    p.N = c.N; p.K = c.K; p.tokens = c.M;
    p.a_tiles = (c.N + DBM - 1) / DBM;
    p.kb_total = (c.K + DBK - 1) / DBK;
    const int splits = (MODE == 1 || MODE == 3) ? (c.splits < 1 ? 1 : c.splits) : 1;
    p.kb_per_split = (p.kb_total + splits - 1) / splits;
    p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
    UMV_REQUIRE(p.splits == splits, UMV_ERR_INVALID, "decode linear: %d splits leave an empty split for K=%d", splits, c.K);
    p.y = c.y; p.ldy = c.ldy; p.bias = c.bias; p.ws = c.ws;
    p.nh = c.norm_h; p.nw = c.norm_w; p.eps = c.eps;
    const int b_bytes_fixed = NORMB ? (p.kb_total + 1) * kBTileNorm : 0;
    const int per_stage = kATile + (NORMB ? 0 : kBTileTma);
    int stages = (kSmemLimit - kStaticReserve - 1024 - b_bytes_fixed - kBarBytes - kTailBytes) / per_stage;
    if (stages > kMaxStages) stages = kMaxStages;
    UMV_REQUIRE(stages >= 4, UMV_ERR_UNSUPPORTED, "decode linear: K=%d leaves %d pipeline stages", c.K, stages);
    p.stages = stages;
    const int smem = 1024 + stages * per_stage + b_bytes_fixed + kBarBytes + kTailBytes;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t ae = cudaFuncSetAttribute(gemm_decode_kernel<NORMB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              kSmemLimit - kStaticReserve);
        if (ae != cudaSuccess) {
            set_error("gemm_decode_kernel: cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString(ae));
            return UMV_ERR_CUDA;
        }
        attr_set = true;
    }
    CUtensorMap tmA, tmB;
    int rc = make_tmap_2d(&tmA, c.w, c.N, c.K, c.K, DBM);
    if (rc) return rc;
    if (NORMB) rc = make_tmap_2d(&tmB, c.norm_h, c.M, c.K, c.K, 8);          // residual rows: 64 x 8 boxes (rows >= M: zero fill)
    else rc = make_tmap_2d(&tmB, c.x, c.M, c.K, c.ldx, DBN);
    if (rc) return rc;
    const int tiles = p.a_tiles * p.splits;
    const int sms = gemm_sm_count();
    int grid = tiles < sms ? tiles : sms;
    char nm[32];
    snprintf(nm, sizeof nm, "dec<%d,%d> N%d K%d", (int)NORMB, MODE, c.N, c.K);
    p.trace = trace_next(nm);
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (MODE == 3) {
        UMV_REQUIRE(tiles <= 132 && (p.splits == 2 || p.splits == 4 || p.splits == 8), UMV_ERR_UNSUPPORTED,
                    "cluster split-K: %d tiles x %d splits does not fit one wave of clusters", p.a_tiles, p.splits);
        grid = tiles;
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = p.splits;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (g_pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(dec_threads<NORMB>());
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_decode_kernel<NORMB, MODE>, tmA, tmB, p);
    ++g_launches;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("gemm_decode_kernel<%d,%d> launch failed: %s", (int)NORMB, MODE, cudaGetErrorString(e));
        return UMV_ERR_CUDA;
    }
    return UMV_OK;
}

}  // namespace

bool decode_linear_supported(int M, int D) {
    return M >= 1 && M <= 8 && D % 64 == 0 && D <= kNormMaxK;
}

int decode_linear(const DecodeLinear& c, cudaStream_t stream) {
    UMV_REQUIRE(c.M >= 1 && c.M <= 8, UMV_ERR_UNSUPPORTED, "decode_linear: %d rows (1..8 are built)", c.M);
    UMV_REQUIRE(c.K % 64 == 0 && c.N % 8 == 0 && ((reinterpret_cast<uintptr_t>(c.w) & 15) == 0), UMV_ERR_INVALID,
                "decode_linear: K=%d must be a multiple of 64, N=%d of 8, weights 16-byte aligned", c.K, c.N);
    int rc = gemm_init();
    if (rc) return rc;
    const bool normb = c.norm_h != nullptr;
    if (normb) {
        UMV_REQUIRE(c.K <= kNormMaxK && c.norm_w, UMV_ERR_UNSUPPORTED, "decode_linear: fused norm needs K <= %d", kNormMaxK);
        switch (c.epi) {
            case EPI_BF16: return launch_decode<true, 0>(c, stream);
            case EPI_PARTIAL: return launch_decode<true, 1>(c, stream);
            case EPI_SWIGLU:
                UMV_REQUIRE(c.N % 128 == 0, UMV_ERR_INVALID, "decode_linear: swiglu needs N %% 128 == 0");
                return launch_decode<true, 2>(c, stream);
            default: break;
        }
    } else if (c.epi == EPI_CLUSTER_RESID) {
        UMV_REQUIRE(c.x && c.ldx % 8 == 0 && ((reinterpret_cast<uintptr_t>(c.x) & 15) == 0) && c.ldy % 4 == 0, UMV_ERR_INVALID,
                    "decode_linear: activation rows must be 16-byte aligned");
        return launch_decode<false, 3>(c, stream);
    }
    set_error("decode_linear: epilogue %d with%s fused norm is not built", c.epi, normb ? "" : "out");
    return UMV_ERR_UNSUPPORTED;
}

}  // namespace umv
