// Engine state: weights in the engine layout, KV page pool + sequences, workspaces.
#pragma once
#include <cuda.h>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/umv.h"
#include "common.cuh"
#include "kernels.cuh"

namespace umv {

enum SlotKind : int { SLOT_PLAIN = 0, SLOT_GATE = 1, SLOT_UP = 2 };

// One reference state-dict tensor -> where it lives in the engine layout.
struct Slot {
    bf16* dst = nullptr;     // first destination element
    int64_t rows = 1, cols = 1;   // reference shape (1-D tensors: rows = 1)
    int64_t dst_ld = 0;      // destination row stride in elements
    int kind = SLOT_PLAIN;
    int ndim = 2;
    int conv_k = 0, conv_cin = 0;   // >0: reference tensor is a conv weight [rows, conv_cin, k, k]; stored [rows, k, k, conv_cin]
    bool loaded = false;
    float synth_bound = 0.02f * 1.7320508f, synth_mean = 0.f;
};

struct LayerW {                    // index 0: understanding expert, 1: generation expert (*_moe_gen)
    bf16 *wqkv[2], *bqkv[2], *wo[2], *wgu[2], *wdown[2], *ln1[2], *ln2[2], *qn[2], *kn[2];
};
struct VitLayerW {
    bf16 *ln1w, *ln1b, *wqkv, *bqkv, *wo, *bo, *ln2w, *ln2b, *w1, *b1, *w2, *b2;
};

struct Seq {
    std::vector<int> pages;
    int len = 0;
    bool alive = false;
    // Identity of the committed token history: 0 = empty, a fork inherits its parent's, every committed append / truncation draws a
    // fresh one.  Equal ids (and lengths) mean bit-identical K / V -- what lets a flow step evaluate two CFG branches over the same
    // context once (umv_flow_velocity).
    uint64_t content = 0;
};

// Device-side metadata of one packed forward call (one H2D copy).
struct CallMeta {
    int* q_start = nullptr;     // [n+1]
    int* k_start = nullptr;     // [n+1] (un-paged attention)
    int* q_len = nullptr;       // [n]
    int* kv_len = nullptr;      // [n]
    int* positions = nullptr;   // [M]
    int* row_seq = nullptr;     // [M]
    int* row_kvpos = nullptr;   // [M]
    int* page_table = nullptr;  // [n][max_pages]
    int* text_rows = nullptr;   // [T] rows routed to the understanding expert (gen mode)
    int* text_slot = nullptr;   // [M] inverse map: index into text_rows, or -1
    uint8_t* row_sel = nullptr; // [M]
    // a prefill whose samples end in a CAUSAL tail behind a full-mask block (image block + prompt in one forward, llm_run's causal_tail):
    // attention runs twice per layer -- the block groups above with the call's mask, the tail groups below causally
    int* tail_q_start = nullptr; int* tail_q_len = nullptr; int* tail_kv_len = nullptr; int* tail_page_table = nullptr;
    int* rope_page_table = nullptr;   // [samples][max_pages]: the K/V append indexes pages by SAMPLE (row_seq), the attention launches by group
    int n_tail = 0, tail_max_q = 0;
    int* seg_to_packed = nullptr;     // [M] segregated gen-mode rows (llm_run): packed row of segregated row i ...
    int* packed_to_seg = nullptr;     // [M] ... and the inverse
    const float* rope_cs = nullptr;   // decode loop: per-step cos | sin table [n][dh]
    int max_pages = 0, n_text = 0;
};

}  // namespace umv

struct umv_engine;
namespace umv {
// ---- VAE (autoencoder.py:38-257) weights, NHWC / tap-major conv layout
struct VaeConv { bf16* w = nullptr; bf16* b = nullptr; int cin = 0, cout = 0, k = 0; };
struct VaeNorm { bf16* w = nullptr; bf16* b = nullptr; int c = 0; };
struct VaeRes { VaeNorm n1, n2; VaeConv c1, c2, sc; int cin = 0, cout = 0; };
struct VaeAttn { VaeNorm n; VaeConv q, k, v, o; int c = 0; };
struct VaeLevel { std::vector<VaeRes> blocks; VaeConv resample; bool has_resample = false; };
struct VaeHalf { VaeConv conv_in, conv_out; VaeRes mid1, mid2; VaeAttn attn; std::vector<VaeLevel> levels; VaeNorm norm_out; };
struct VaeState {
    VaeHalf enc, dec;
    int ch = 128, z = 16, nlev = 4, nres = 2;
    int mult[4] = {1, 2, 4, 4};
    float scale = 0.3611f, shift = 0.1159f;
    // workspaces (grown on demand)
    bf16 *a0 = nullptr, *a1 = nullptr, *a2 = nullptr, *col = nullptr, *p = nullptr, *vt = nullptr;
    float *s = nullptr, *stats = nullptr;
    size_t act_elems = 0, col_elems = 0, s_elems = 0;
};
int vae_build(umv_engine* e);
int latent_patchify(const bf16* z, bf16* rows, int C, int Hl, int Wl, int h, int w, int p, cudaStream_t st);
// engine.cu internals shared with flow.cu / vae.cu
int engine_alloc(umv_engine* e, void** out, size_t bytes);
void engine_reg(umv_engine* e, const std::string& name, bf16* dst, int64_t rows, int64_t cols, int ndim, int conv_k, int conv_cin,
                float bound, float mean);
int lin(umv_engine* e, const bf16* x, int ldx, const bf16* w, const bf16* bias, const bf16* res, bf16* y, int ldy, int M, int N,
        int K, int epi, cudaStream_t st, int impl = 0, float* ws = nullptr, int splits = 1, int stages = 0,
        const int* res_rows = nullptr, bf16* res_gather_tmp = nullptr);
// Parity hook (umv_op_attention_block): run only the attention block of one layer on caller-provided projection outputs.
struct AttnProbe {
    int layer = 0;
    const bf16* qkv = nullptr;        // [M, (H+2Hkv)*dh] bf16 (bias applied), or
    const float* partial = nullptr;   // [splits][M][(H+2Hkv)*dh] fp32 split-K partials + bias
    int splits = 0;
    const bf16* bias = nullptr;
    bf16* out = nullptr;              // [M, H*dh]
    int path = 0;                     // out: 1 mma.sync, 2 tcgen05, 3 fused decode cluster kernel
};
int llm_run(umv_engine* e, const bf16* x, int n_seqs, const int32_t* seqs, const int32_t* q_lens, const int32_t* positions,
            const uint8_t* row_is_gen, int is_causal, int update_kv, bf16* out, cudaStream_t st, AttnProbe* probe = nullptr,
            int presegregated = 0, const int32_t* causal_tail = nullptr);
// A non-causal gen-mode forward runs on SEGREGATED rows: the generation-expert rows of all samples first (packed order), then the
// understanding-expert (marker) rows.  True when llm_run will use that layout for these arguments (a full-mask forward with rows of both
// kinds; UMV_GEN_SEG=0 switches it off); *n_gen = generation rows.
bool gen_rows_segregate(int n_seqs, const int32_t* q_lens, const uint8_t* row_is_gen, int is_causal, int* n_gen);
}  // namespace umv

struct umv_engine {
    umv_dims d{};
    int dh = 0, qkvn = 0, vit_kpad = 0, sm_count = 148;
    bool finalized = false;
    bool use_splitk = true, use_graph = true;
    int gemm_impl = 0;

    std::vector<void*> allocs;
    // tile-major twins of the weights the decode step streams (LinearCall::w_tiled): row-major pointer -> twin.  Built by umv_finalize
    // (UMV_TILED=0: none).  180 GB of HBM per GPU pays for the second copy (+14.1 GB at the 14B dims) of the understanding expert.
    std::map<const umv::bf16*, umv::bf16*> tiled;
    std::vector<void*> tiled_allocs;
    std::map<std::string, umv::Slot> slots;

    // LLM
    umv::bf16 *embed = nullptr, *lm_head = nullptr, *final_norm[2] = {nullptr, nullptr};
    std::vector<umv::LayerW> layers;
    float* inv_freq = nullptr;
    // ViT + connector
    umv::bf16 *vit_patch_w = nullptr, *vit_patch_b = nullptr, *vit_pos = nullptr, *vit_post_w = nullptr, *vit_post_b = nullptr;
    std::vector<umv::VitLayerW> vit;
    umv::bf16 *conn_w1 = nullptr, *conn_b1 = nullptr, *conn_w2 = nullptr, *conn_b2 = nullptr, *vit_pos_embed = nullptr;
    // generation glue
    umv::bf16 *t_w0 = nullptr, *t_b0 = nullptr, *t_w2 = nullptr, *t_b2 = nullptr, *vae2llm_w = nullptr, *vae2llm_b = nullptr,
              *llm2vae_w = nullptr, *llm2vae_b = nullptr, *latent_pos = nullptr;

    // KV
    umv::KVPool pool;
    CUtensorMap kv_tmap;               // pool as [slot rows, head_dim] for the decode-attention TMA loads
    bool kv_tmap_ok = false;
    std::vector<int> page_ref, free_pages;
    std::vector<umv::Seq> seqs;
    uint64_t content_counter = 0;   // source of Seq::content ids
    int last_flow_branches = 0;     // CFG branches the last umv_flow_velocity ran (umv_flow_branches_last)

    // workspaces
    umv::bf16 *h = nullptr, *xn = nullptr, *qkv = nullptr, *attn = nullptr, *act = nullptr, *logits = nullptr;
    umv::bf16* vit_qkvp = nullptr;   // [max_tokens, 3 * vit_heads * 128]: ViT q|k|v with heads zero-padded to 128 columns (tcgen05 attention)
    umv::bf16 *xt = nullptr, *ht = nullptr, *yt = nullptr, *actt = nullptr;   // text-row (understanding expert) staging, gen mode
    float* ws = nullptr;          // split-K partials
    size_t ws_elems = 0;
    float* attn_ws = nullptr;     // split-KV partials
    size_t attn_ws_elems = 0;
    int w_h = 0, w_qkv = 0, w_act = 0;   // workspace row widths

    // call metadata staging
    static constexpr int kMetaRing = 4;
    uint8_t* meta_host[kMetaRing] = {};
    uint8_t* meta_dev[kMetaRing] = {};
    cudaEvent_t meta_ev[kMetaRing] = {};
    size_t meta_bytes = 0;
    int meta_next = 0;

    // decode state
    int64_t* dec_tokens = nullptr;
    umv::VaeState* vae = nullptr;
    // flow scratch
    umv::bf16 *flow_small = nullptr;   // [4, hidden]: timestep frequencies / hidden / embedding
    float* t_freqs = nullptr;          // [128] exp(-ln(1e4) i / 128)
    int *dec_pos = nullptr, *dec_kvlen = nullptr, *dec_kvpos = nullptr, *dec_step = nullptr, *dec_rowseq = nullptr,
        *dec_qstart = nullptr, *dec_qlen = nullptr, *dec_pages = nullptr;
    float* rope_tab = nullptr;         // [max_tokens][dh] per-forward rope cos | sin
    float* dec_rope = nullptr;         // [64][dh] per-step rope cos | sin
    int dec_pages_cap = 0;
    // decode-step CUDA graphs kept across umv_generate_text calls (capture + instantiate of the 201-launch step costs ~7 ms)
    struct DecodeGraph {
        int B = 0, max_pages = 0, blocks = 0;
        float temperature = 0.f;
        uint64_t seed = 0;
        unsigned flags = 0;
        cudaGraphExec_t exec = nullptr;
        long long launches = 0;      // kernels per replay
        uint64_t used = 0;           // LRU stamp
    };
    std::vector<DecodeGraph> dec_graphs;
    uint64_t dec_graph_clock = 0;
    int64_t* dec_out = nullptr;      // [steps, B] tokens of a cached-graph run (copied to the caller's tokens_out afterwards)
    size_t dec_out_cap = 0;
};
