// Token-major linear on CTA pairs (tcgen05 cta_group::2): two CTAs of a cluster -- two SMs of one TPC -- share one
// 256 x BN output tile.  Each CTA stages its own 128 activation rows and HALF of the weight tile (BN / 2 rows) per k-block,
// the leader CTA's elected thread issues tcgen05.mma.cta_group::2 (M = 256, N = BN) which reads both halves of the weight
// tile from the two shared memories, and each CTA drains its own 128 x BN accumulator from its own TMEM.  Per CTA and
// k-block that is 16 KB + BN / 2 * 128 B of shared-memory fill for 128 x BN x 64 MACs -- 64 (BN = 256) / 96 (BN = 128)
// bytes per clock instead of 96 / 128 for the single-CTA kernel of gemm.cu, which is what limits its narrow tiles.
//
// Synchronisation: per-stage `full` barrier lives in the leader (both CTAs' TMA loads complete on it, the peer announces
// its stage with a remote arrive); `empty` and `tfull` exist in both CTAs and are signalled by multicast tcgen05.commit;
// `tempty` lives in the leader and collects the epilogue warps of both CTAs (remote arrives from the peer).
//
// Same epilogues, rounding points and call sites as gemm.cu (MODE 0: bias / GELU / residual, MODE 2: SwiGLU).
#include <cuda.h>
#include <algorithm>
#include <cstdlib>

#include "../../include/umv.h"
#include "common.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace umv {

namespace {

constexpr int BM = 128, BK = 64;
constexpr int kATile = BM * BK * 2;
constexpr int kThreads = 320;            // warp 0 TMA, warp 1 MMA (leader only), warps 2-9 epilogue
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;      // shared::cluster address of the same offset in the pair's even CTA

struct Gemm2Params {
    int K, kb_total, m_pairs, n_tiles, tokens, features;
    bf16* y;
    int ldy;
    const bf16* bias;
    const bf16* residual;
    int epi;
    int hd, hp;          // output head padding (LinearCall::out_head_dim / out_head_pad), 0 = off
    int stages;
    int group_m;         // tile order: groups of `group_m` row pairs are swept over all column tiles before the next group starts
    TraceSlot* trace;
};

// Tile order.  The persistent grid takes tiles tile, tile + clusters, ...: the ~74 tiles in flight at any moment are CONSECUTIVE
// indices, and all of them walk K in lock step -- so a k-block of activations or weights is fetched from HBM once if every tile that
// needs it is in flight together, and again for every later wave that needs it.  Row pairs fastest over ALL rows (group_m = m_pairs)
// makes each wave touch every activation row: at K = 18944 the 8,208-row prefill re-read its 311 MB of activations in each of 7 waves
// (ncu: 2.5 GB of DRAM reads for 0.5 GB of operands).  Sweeping the column tiles inside groups of group_m row pairs keeps a wave's
// operand set near-square; the host picks group_m from the operand sizes (pick_tile_group, gemm.cu).  The order changes no result bit.
__device__ __forceinline__ void tile_coords(int tile, const Gemm2Params& p, int& mp, int& b_tile) {
    const int per_group = p.group_m * p.n_tiles;
    const int g = tile / per_group;
    const int first = g * p.group_m;
    const int gsz = min(p.group_m, p.m_pairs - first);
    const int r = tile - g * per_group;
    mp = first + r % gsz;
    b_tile = r / gsz;
}

template <int BN>
struct Cfg2 {
    static constexpr int kBTile = (BN / 2) * BK * 2;            // this CTA's half of the weight tile
    static constexpr int kStageBytes = kATile + kBTile;
    static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
    static constexpr int kTmemCols = 2 * BN <= 256 ? 256 : 512;
    static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 256;
};

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {       // arrive on the barrier at `bar`'s offset in CTA `rank`
    uint32_t addr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
// 2-D tiled load whose completion is counted on the LEADER CTA's barrier (executed by both CTAs of the pair)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_in_smem) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_in_smem)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_pair(int n) {         // kind::f16: D f32, A = B = bf16 K-major, M = 256, N = n
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
}
__device__ __forceinline__ float epi_value2(int epi, float acc, float bias) {
    float v = rbf(acc + bias);
    if (epi == EPI_GELU) v = rbf(gelu_tanh_f(v));
    return v;
}

template <int BN, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
gemm_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Gemm2Params p) {
    using C2 = Cfg2<BN>;
    const int kStages = p.stages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sB = smem + kStages * kATile;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * C2::kStageBytes);
    uint64_t* empty = full + 8;
    uint64_t* tfull = empty + 8;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const bool leader = rank == 0;
    pdl_launch_dependents();
    trace_start(p.trace);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < kStages; ++i) {
                mbar_init(&full[i], 2);          // the leader's expect_tx arrive + the peer's announcement
                mbar_init(&empty[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&tfull[i], 1);
                mbar_init(&tempty[i], 16);       // 8 epilogue warps of each CTA
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc_pair<C2::kTmemCols>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                          // the peer's barriers exist before any remote arrive / TMA completion
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
    const int num_tiles = p.m_pairs * p.n_tiles;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (elect_one()) {
            pdl_wait();
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
                int mp, b_tile;
                tile_coords(tile, p, mp, b_tile);
                const int a_row = (2 * mp + (int)rank) * BM;
                const int b_row = b_tile * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < p.kb_total; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    if (leader) mbar_expect_tx(&full[stage], 2 * C2::kStageBytes);
                    else mbar_arrive_remote(&full[stage], 0);
                    tma_load_2d_pair(sA + stage * kATile, &tmA, &full[stage], kb * BK, a_row);
                    tma_load_2d_pair(sB + stage * C2::kBTile, &tmB, &full[stage], kb * BK, b_row);
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
            if (p.trace && blockIdx.x == 0) p.trace->t_wait = gtime();
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        if (leader && elect_one()) {
            constexpr uint32_t idesc = idesc_pair(BN);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
                mbar_wait(&tempty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < p.kb_total; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * kATile);
                    const uint32_t b_addr = smem_u32(sB + stage * C2::kBTile);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16_pair(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                                       (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit_pair(&empty[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
                umma_commit_pair(&tfull[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: this CTA's 128 rows (8 warps)
        const int quarter = warp & 3;
        const int chalf = (warp - 2) >> 2;
        const int lrow = quarter * 32 + lane;
        pdl_wait();
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
            int mp, b_tile;
            tile_coords(tile, p, mp, b_tile);
            const int a_tile = 2 * mp + (int)rank;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
            const int row = a_tile * BM + lrow;
            const bool row_ok = row < p.tokens;
            if constexpr (MODE == 0) {
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
                                    float v0 = epi_value2(p.epi, __uint_as_float(r[h * 8 + 2 * j]), b2.x);
                                    float v1 = epi_value2(p.epi, __uint_as_float(r[h * 8 + 2 * j + 1]), b2.y);
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
            } else {
                // weight rows interleaved [64 gate | 64 up] -> tile columns [blk*128 + c] / [blk*128 + 64 + c]
                constexpr int kUnits = (BN / 128) * 4;
#pragma unroll 1
                for (int unit = chalf * (kUnits / 2); unit < (chalf + 1) * (kUnits / 2); ++unit) {
                    const int blk = unit >> 2, c0 = (unit & 3) * 16;
                    uint32_t g[16], u[16];
                    tmem_ld16(taddr + blk * 128 + c0, g);
                    tmem_ld16(taddr + blk * 128 + 64 + c0, u);
                    tmem_ld_wait();
                    const int j0 = (b_tile * BN) / 2 + blk * 64 + c0;
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
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(&tempty[acc]);
                else mbar_arrive_remote(&tempty[acc], 0);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                          // the pair's MMAs, remote arrives and TMEM reads are all done
    if (warp == 1) tmem_dealloc_pair<C2::kTmemCols>(tmem_base);
    trace_end(p.trace);
}

template <int BN, int MODE>
int launch_2cta(const LinearCall& c, cudaStream_t stream) {
    using C2 = Cfg2<BN>;
    Gemm2Params p{};
    p.K = c.K;
    p.kb_total = (c.K + BK - 1) / BK;
    p.m_pairs = (c.M + 2 * BM - 1) / (2 * BM);
    p.n_tiles = (c.N + BN - 1) / BN;
    p.tokens = c.M; p.features = c.N;
    p.y = c.y; p.ldy = c.ldy; p.bias = c.bias; p.residual = c.residual; p.epi = c.epi;
    p.hd = c.out_head_dim; p.hp = c.out_head_pad;
    UMV_REQUIRE(!p.hd || (MODE == 0 && c.epi == EPI_BF16 && p.hd % 8 == 0 && p.hp % 8 == 0), UMV_ERR_UNSUPPORTED,
                "linear: output head padding exists only on the bf16 epilogue with 8-column aligned heads");
    p.stages = C2::kStages;
    p.group_m = pick_tile_group(p.m_pairs, p.n_tiles, 2 * BM, BN, c.K, gemm_sm_count() / 2);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t ae = cudaFuncSetAttribute(gemm_2cta_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2::kSmemBytes);
        if (ae != cudaSuccess) {
            set_error("gemm_2cta_kernel: cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString(ae));
            return UMV_ERR_CUDA;
        }
        attr_set = true;
    }
    CUtensorMap tmA, tmB;
    int rc = make_tmap_2d(&tmA, c.x, c.M, c.K, c.ldx, BM);
    if (rc) return rc;
    rc = make_tmap_2d(&tmB, c.w, c.N, c.K, c.K, BN / 2);
    if (rc) return rc;
    const int tiles = p.m_pairs * p.n_tiles;
    const int max_clusters = gemm_sm_count() / 2;
    const int clusters = tiles < max_clusters ? tiles : max_clusters;
    char nm[32];
    snprintf(nm, sizeof nm, "gemm2<%d,%d> N%d K%d", BN, MODE, c.N, c.K);
    p.trace = trace_next(nm);
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C2::kSmemBytes;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_2cta_kernel<BN, MODE>, tmA, tmB, p);
    ++g_launches;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("gemm_2cta_kernel<%d,%d> launch failed: %s", BN, MODE, cudaGetErrorString(e));
        return UMV_ERR_CUDA;
    }
    return UMV_OK;
}

}  // namespace

// Shapes the pair kernel takes over from the single-CTA token-major kernel.
bool linear_2cta_supported(const LinearCall& c) {
    if (c.M < 2 * BM || c.N < 128 || c.N % 8 != 0 || c.ldy % 8 != 0) return false;
    if (c.epi == EPI_SWIGLU) return c.N % 256 == 0;
    return c.epi == EPI_BF16 || c.epi == EPI_GELU || c.epi == EPI_RESID;
}

int linear_2cta_forward(const LinearCall& c, cudaStream_t stream, int bn) {
    if (c.epi == EPI_SWIGLU) return launch_2cta<256, 2>(c, stream);
    return bn == 128 ? launch_2cta<128, 0>(c, stream) : launch_2cta<256, 0>(c, stream);
}

}  // namespace umv
