// Host-side interface of the HBM-bound glue kernels (kernels.cu) and attention (attention.cu).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"

namespace umv {

constexpr int kPageTokens = 64;

// Paged KV pool geometry: pool[layer][page][k|v][kv_head][slot][head_dim] bf16.  LAYER-major: the tiles one attention launch reads
// (one layer, the pages of the running samples) lie within pages * 128 KB of each other instead of one 3.67 MB page-stride apart, so a
// launch touches tens of 2 MB translations rather than one per KV page (the TLB reaches 256 MB; at the 1,000+ page pools of the serving configs the page-major layout put every tile of a launch on its own
// translation).  Measured neutral at B = 8 / 208 pages (profiles/r2_decode_experiments.md).
struct KVPool {
    bf16* base = nullptr;
    int layers = 0, kv_heads = 0, head_dim = 0, pages = 0;
    __host__ __device__ size_t tile_elems() const { return (size_t)kPageTokens * head_dim; }
    __host__ __device__ size_t tile_offset(int page, int layer, int kv, int head) const {
        return ((((size_t)layer * pages + page) * 2 + kv) * kv_heads + head) * tile_elems();
    }
    __host__ __device__ size_t page_layer_elems() const { return (size_t)2 * kv_heads * tile_elems(); }   // one (layer, page): contiguous
};

// ---- residual add + RMSNorm (Qwen2RMSNorm, modeling_qwen2.py:89-94; residual adds qwen2_navit.py:883,901)
struct AddNormArgs {
    bf16* h = nullptr;              // [M, D] residual stream; updated in place when a delta is given
    const bf16* delta = nullptr;    // [M, D] bf16 linear output, or
    const float* partial = nullptr; // [splits][M][D] fp32 split-K partials of that linear output
    int splits = 0;
    const bf16* w0 = nullptr;       // norm weight for rows with row_sel == 0 (understanding expert)
    const bf16* w1 = nullptr;       // norm weight for rows with row_sel == 1 (generation expert)
    const uint8_t* row_sel = nullptr;
    bf16* y = nullptr;              // [M, D] normalised output (null: only the residual add)
    bf16* y2 = nullptr;             // optional compact copy: rows with row_slot[m] >= 0 are also written to y2[row_slot[m]]
    const int* row_slot = nullptr;  //   (the understanding-expert rows of a gen-mode forward, gathered for their own linears)
    int M = 0, D = 0;
    float eps = 1e-6f;
    TraceSlot* trace = nullptr;
};
int add_rmsnorm(const AddNormArgs& a, cudaStream_t s);

// nn.LayerNorm under CUDA autocast (fp32 math) followed by the consumer Linear's cast: bf16 out.
int layernorm_bf16(const bf16* x, const bf16* w, const bf16* b, bf16* y, int M, int D, float eps, cudaStream_t s);

// ---- q/k RMSNorm + RoPE + KV append (qwen2_navit.py:544-545,568-600; modeling_qwen2.py:164-220)
struct RopeAppendArgs {
    const bf16* qkv = nullptr;       // [M, (H+2Hkv)*dh] bf16 (bias already applied), or
    const float* partial = nullptr;  // [splits][M][(H+2Hkv)*dh] fp32 partials (+ bias below)
    int splits = 0;
    const bf16* bias = nullptr;      // used with `partial`
    bf16* q_out = nullptr;           // [M, ldq] rotated queries (may alias qkv)
    int ldq = 0;
    const int* q_row_map = nullptr;  // optional: the queries of row m land in row q_row_map[m] of q_out (which must then NOT alias qkv)
    const int* positions = nullptr;  // [M] rope position per row
    const int* row_seq = nullptr;    // [M] index into the page table rows
    const int* row_kvpos = nullptr;  // [M] absolute KV slot of the row within its sequence
    const int* page_table = nullptr; // [n_seqs][max_pages]
    int max_pages = 0;
    const float* inv_freq = nullptr; // [dh/2] fp32
    const float* rope_cs = nullptr;  // optional [M][dh]: bf16-rounded cos | sin of positions[m] * inv_freq (rope_table)
    const bf16* qn0 = nullptr; const bf16* kn0 = nullptr;   // understanding q_norm / k_norm
    const bf16* qn1 = nullptr; const bf16* kn1 = nullptr;   // *_moe_gen
    const uint8_t* row_sel = nullptr;                        // null in "und" mode
    int gen_mode = 0;                // 1: fp32 norm+rope, single rounding (qwen2_navit.py:568-583)
    KVPool pool; int layer = 0;
    int M = 0, H = 0, Hkv = 0, dh = 0;
    float eps = 1e-6f;
    TraceSlot* trace = nullptr;
};
int rope_append(const RopeAppendArgs& a, cudaStream_t s);
// cs[m][0:dh/2] = bf16(cos(positions[m] * inv_freq)), cs[m][dh/2:dh] = bf16(sin(..)) as fp32 values: the angles are the same
// for every layer of a forward, so they are evaluated once (modeling_qwen2.py:164-181 recomputes them per layer).
int rope_table(const int* positions, const float* inv_freq, int M, int dh, float* cs, cudaStream_t s);

// ---- attention (flash_attn_varlen_func call sites qwen2_navit.py:605-614, siglip_navit.py:232-241)
struct AttnArgs {
    const bf16* q = nullptr; int ldq = 0;     // row (token) stride in elements; head h at column h*dh
    bf16* out = nullptr; int ldo = 0;
    const int* out_row_map = nullptr;         // optional: the output of query row m lands in row out_row_map[m] of out
    // un-paged K/V (ViT, op-level): [Tk, Hkv, dh] with row strides
    const bf16* k = nullptr; const bf16* v = nullptr; int ldk = 0, ldv = 0;
    // paged K/V (LLM)
    int paged = 0; KVPool pool; int layer = 0; const int* page_table = nullptr; int max_pages = 0;
    const int* q_start = nullptr;   // device [n+1] cumulative query offsets
    const int* k_start = nullptr;   // device [n+1] cumulative key offsets (un-paged only)
    const int* q_len = nullptr;     // device [n]
    const int* kv_len = nullptr;    // device [n] keys visible to the sample (past + new)
    int n = 0, H = 0, Hkv = 0, dh = 0, causal = 0;
    int max_q_len = 0;              // host upper bound (grid sizing)
    int max_kv_len = 0;             // host upper bound (split sizing)
    int splits = 1;                 // split-KV factor (>1 needs ws)
    float* ws = nullptr;            // [splits][total_q*H][dh+1] fp32 partial outputs + lse
    int total_q = 0;
    TraceSlot* trace = nullptr; TraceSlot* trace_combine = nullptr;
    const CUtensorMap* kv_tmap = nullptr;   // host pointer: the paged pool as [slot rows, 128] (64 x 64 boxes); enables attention_tc
    // head_dim-padded callers (ViT: 72 real columns in 128-wide, zero-padded heads): softmax scale and output width of the
    // real head_dim; total_k = rows of the un-paged K / V matrices (tensor-map extent)
    float scale = 0.f;              // 0: 1 / sqrt(dh)
    int out_dh = 0;                 // 0: dh; else out holds heads of out_dh columns and only those are stored
    int total_k = 0;
};
// tcgen05 path (attention_tc.cu): head_dim 128, paged KV, at least one 128-row tile of (token, head) rows per sample
bool attention_tc_supported(const AttnArgs& a);
int attention_tc_forward(const AttnArgs& a, cudaStream_t s);
int attention_forward(const AttnArgs& a, cudaStream_t s);
int attention_init();   // once per process, before any capture

// ---- misc
int embed_rows(const bf16* table, const int64_t* ids, int n, int D, int64_t vocab, bf16* out, cudaStream_t s);
int gather_add_rows(bf16* x, const bf16* table, const int64_t* ids, int M, int D, cudaStream_t s);  // x[m] = bf16(x[m] + table[ids[m]])
int embed_rows_scatter(const bf16* table, const int64_t* ids, const int* dst_rows, int n, int D, int64_t vocab, bf16* out, cudaStream_t s);
int scatter_add_rows(const bf16* src, const bf16* table, const int64_t* ids, const int* dst_rows, bf16* dst, int M, int D, cudaStream_t s);
// uint8 HWC image -> normalised fp32 patch vectors + position ids (one CTA per patch)
int patchify_u8(const uint8_t* img, int H, int W, int patch, int max_per_side, float* out, int64_t* pos_ids, cudaStream_t s);
int f32_to_bf16_padded(const float* x, bf16* y, int M, int K, int Kpad, cudaStream_t s);
struct DecodeState;
// `end` != null (decode loop): each row's CTA also advances that sample's loop state (replaces decode_end_step)
int argmax_rows(const bf16* logits, int rows, int vocab, int64_t* out, cudaStream_t s, const DecodeState* end = nullptr);
int copy_rows(const bf16* src, int lds, const int* rows, bf16* dst, int ldd, int n, int D, int scatter, cudaStream_t s);
int fill_uniform_bf16(bf16* p, size_t n, uint64_t seed, float bound, float mean, cudaStream_t s);

// decode-loop state kernels (Bagel.generate_text, bagel.py:1262-1311)
struct DecodeState {
    int64_t* cur_tokens;     // [B]
    int* positions;          // [B] rope position of the next query
    int* kv_len;             // [B] KV length including the token being processed this step
    int* row_kvpos;          // [B] slot the new token's K/V go to (= kv_len - 1)
    int* step;               // [1]
    float* rope_cs;          // [B][dh] cos | sin of the step's rope angles (bf16-rounded), shared by all layers; may be null
    const float* inv_freq;   // [dh/2]
    int dh;
};
int decode_begin_step(const bf16* table, int D, int64_t vocab, DecodeState st, const int64_t* forced, int64_t* tokens_out,
                      int B, bf16* x, cudaStream_t s);
int decode_end_step(DecodeState st, int B, cudaStream_t s);
int sample_rows(const bf16* logits, int rows, int vocab, float temperature, uint64_t seed, const int* step, int64_t* out,
                cudaStream_t s, float u_force = -1.f);
int export_kv(KVPool pool, int layer, const int* pages, int len, bf16* k_out, bf16* v_out, cudaStream_t s);
int copy_page(KVPool pool, int src_page, int dst_page, cudaStream_t s);

// ---- rectified-flow glue (flow.cu)
struct CfgArgs {
    const bf16* v;
    int rows_per_branch, C;
    int text_branch, img_branch;       // branch index or -1
    float text_scale, img_scale, renorm_min;
    int renorm_type;                   // 0 global (per image), 1 channel, 2 text_channel
    const int* img_row0;               // [B] packed row of the image's first latent token
    const int* img_lat0;               // [B] first latent index
    const int* img_n;                  // [B] latent tokens
    float* out;                        // [n_lat, C] fp32
};
int timestep_freq(float t, const float* freqs, int half, bf16* out, cudaStream_t s);
int silu_inplace(bf16* x, int n, cudaStream_t s);
int flow_compose(const bf16* lat, const bf16* temb, const bf16* pos_table, const int64_t* pos_ids, const bf16* embed,
                 int64_t id_start, int64_t id_end, const int* row_src, int rows_per_branch, int branches, int D, bf16* out,
                 cudaStream_t s, const int* row_dst = nullptr);
int cfg_combine(const CfgArgs& a, int n_images, cudaStream_t s);
int euler_step(float* x, const float* v, int64_t n, float dt, int v_is_bf16, cudaStream_t s);

}  // namespace umv

namespace umv {
// ---- fused decode attention (attention.cu): split-K reduce + bias + q/k RMSNorm + RoPE + KV append + split-KV
// attention + combine through distributed shared memory of a thread-block cluster, one launch per layer.
struct DecodeAttnArgs {
    const bf16* qkv = nullptr;        // [M, ncols] bf16 (bias applied), or
    const float* partial = nullptr;   // [ksplits][M][ncols] fp32 split-K partials (+ bias)
    int ksplits = 0;
    const bf16* bias = nullptr;
    bf16* out = nullptr; int ldo = 0; // [M, H*dh]
    const int* positions = nullptr;   // [M] rope position of the (single) query token of each sample
    const int* kv_len = nullptr;      // [M] keys visible, including the token being appended
    const int* page_table = nullptr; int max_pages = 0;
    const float* inv_freq = nullptr;
    const float* rope_cs = nullptr;   // optional [M][dh]: bf16-rounded cos (first dh/2) and sin of position * inv_freq, per sample
    const bf16* qn = nullptr; const bf16* kn = nullptr;
    KVPool pool; int layer = 0;
    int M = 0, H = 0, Hkv = 0;
    const CUtensorMap* kv_tmap = nullptr;   // host pointer: the pool as a 2-D tensor [slot rows, 128], 64 x 64 boxes, 128 B swizzle
    int cluster = 8;                  // CTAs (key ranges) per (sample, kv head)
    int early = 0;                    // 1: read the step state and request the K/V tiles before griddepcontrol.wait (not for layer 0)
    float eps = 1e-6f;
    TraceSlot* trace = nullptr;
};
int decode_attention(const DecodeAttnArgs& a, cudaStream_t s);
bool decode_attention_supported(int H, int Hkv, int dh, int max_pages, int ksplits);
}  // namespace umv
