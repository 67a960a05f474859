// Engine: weights in the engine layout, ref-counted KV pages, packed LLM / ViT forward orchestration,
// device-resident greedy decode loop (CUDA-graph replay) and the C ABI of include/umv.h.
//
// Reference drivers restated here: Bagel.forward_cache_update_{text,vit} (bagel.py:412-458,523-615),
// Qwen2Model.forward_inference (qwen2_navit.py:1115-1176), Qwen2MoTDecoderLayer.forward_inference
// (:843-902), SiglipVisionTransformer.forward (siglip_navit.py:345-371), Bagel.generate_text
// (bagel.py:1236-1317), NaiveCache + deepcopy forks (qwen2_navit.py:207-221, inferencer.py:261,587-607).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "engine.cuh"
#include "gemm.cuh"
#include "kernels.cuh"

namespace umv {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

template <typename T>
static int dev_alloc(umv_engine* e, T** out, size_t n) {
    void* p = nullptr;
    cudaError_t err = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
    if (err != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(err));
        return UMV_ERR_NOMEM;
    }
    e->allocs.push_back(p);
    *out = static_cast<T*>(p);
    return UMV_OK;
}
#define UMV_TRY(expr)            \
    do {                         \
        int _rc = (expr);        \
        if (_rc != UMV_OK) return _rc; \
    } while (0)

static void reg(umv_engine* e, const std::string& name, bf16* dst, int64_t rows, int64_t cols, int64_t dst_ld, int ndim,
                int kind = SLOT_PLAIN, float bound = 0.02f * 1.7320508f, float mean = 0.f) {
    Slot s;
    s.dst = dst; s.rows = rows; s.cols = cols; s.dst_ld = dst_ld; s.ndim = ndim; s.kind = kind;
    s.synth_bound = bound; s.synth_mean = mean;
    e->slots[name] = s;
}

int engine_alloc(umv_engine* e, void** out, size_t bytes) {
    uint8_t* p = nullptr;
    UMV_TRY(dev_alloc(e, &p, bytes));
    *out = p;
    return UMV_OK;
}
void engine_reg(umv_engine* e, const std::string& name, bf16* dst, int64_t rows, int64_t cols, int ndim, int conv_k, int conv_cin,
                float bound, float mean) {
    reg(e, name, dst, rows, cols, cols, ndim, SLOT_PLAIN, bound, mean);
    e->slots[name].conv_k = conv_k;
    e->slots[name].conv_cin = conv_cin;
}

// Allocates one linear layer [rows, cols] (+ optional bias) and registers its reference names.
static int alloc_linear(umv_engine* e, const std::string& name, int64_t rows, int64_t cols, bool bias, bf16** w, bf16** b,
                        int64_t ld = 0) {
    if (ld == 0) ld = cols;
    UMV_TRY(dev_alloc(e, w, (size_t)rows * ld));
    if (ld != cols) cudaMemset(*w, 0, (size_t)rows * ld * sizeof(bf16));
    reg(e, name + ".weight", *w, rows, cols, ld, 2, SLOT_PLAIN, std::sqrt(3.0f / (float)cols));
    if (bias) {
        UMV_TRY(dev_alloc(e, b, (size_t)rows));
        reg(e, name + ".bias", *b, 1, rows, rows, 1);
    }
    return UMV_OK;
}
static int alloc_vec(umv_engine* e, const std::string& name, int64_t n, bf16** p, float mean, float bound) {
    UMV_TRY(dev_alloc(e, p, (size_t)n));
    reg(e, name, *p, 1, n, n, 1, SLOT_PLAIN, bound, mean);
    return UMV_OK;
}

static int build_weights(umv_engine* e) {
    const umv_dims& d = e->d;
    const int D = d.hidden, I = d.inter, dh = e->dh, H = d.heads, Hkv = d.kv_heads, QN = e->qkvn;
    const std::string LM = "language_model.";
    UMV_TRY(dev_alloc(e, &e->embed, (size_t)d.vocab * D));
    reg(e, LM + "model.embed_tokens.weight", e->embed, d.vocab, D, D, 2);
    UMV_TRY(dev_alloc(e, &e->lm_head, (size_t)d.vocab * D));
    reg(e, LM + "lm_head.weight", e->lm_head, d.vocab, D, D, 2);
    const int nexp = d.enable_gen ? 2 : 1;
    const char* sfx[2] = {"", "_moe_gen"};
    for (int x = 0; x < nexp; ++x) UMV_TRY(alloc_vec(e, LM + "model.norm" + sfx[x] + ".weight", D, &e->final_norm[x], 1.f, 0.1f));
    e->layers.resize(d.layers);
    for (int li = 0; li < d.layers; ++li) {
        LayerW& L = e->layers[li];
        memset(&L, 0, sizeof(L));
        const std::string P = LM + "model.layers." + std::to_string(li) + ".";
        for (int x = 0; x < nexp; ++x) {
            const std::string s = sfx[x];
            UMV_TRY(dev_alloc(e, &L.wqkv[x], (size_t)QN * D));
            UMV_TRY(dev_alloc(e, &L.bqkv[x], (size_t)QN));
            reg(e, P + "self_attn.q_proj" + s + ".weight", L.wqkv[x], H * dh, D, D, 2);
            reg(e, P + "self_attn.k_proj" + s + ".weight", L.wqkv[x] + (size_t)H * dh * D, Hkv * dh, D, D, 2);
            reg(e, P + "self_attn.v_proj" + s + ".weight", L.wqkv[x] + (size_t)(H + Hkv) * dh * D, Hkv * dh, D, D, 2);
            reg(e, P + "self_attn.q_proj" + s + ".bias", L.bqkv[x], 1, H * dh, H * dh, 1);
            reg(e, P + "self_attn.k_proj" + s + ".bias", L.bqkv[x] + H * dh, 1, Hkv * dh, Hkv * dh, 1);
            reg(e, P + "self_attn.v_proj" + s + ".bias", L.bqkv[x] + (H + Hkv) * dh, 1, Hkv * dh, Hkv * dh, 1);
            bf16* nob = nullptr;
            UMV_TRY(alloc_linear(e, P + "self_attn.o_proj" + s, D, H * dh, false, &L.wo[x], &nob));
            UMV_TRY(dev_alloc(e, &L.wgu[x], (size_t)2 * I * D));
            reg(e, P + "mlp" + s + ".gate_proj.weight", L.wgu[x], I, D, D, 2, SLOT_GATE);
            reg(e, P + "mlp" + s + ".up_proj.weight", L.wgu[x] + (size_t)64 * D, I, D, D, 2, SLOT_UP);
            UMV_TRY(alloc_linear(e, P + "mlp" + s + ".down_proj", D, I, false, &L.wdown[x], &nob));
            UMV_TRY(alloc_vec(e, P + "input_layernorm" + s + ".weight", D, &L.ln1[x], 1.f, 0.1f));
            UMV_TRY(alloc_vec(e, P + "post_attention_layernorm" + s + ".weight", D, &L.ln2[x], 1.f, 0.1f));
            UMV_TRY(alloc_vec(e, P + "self_attn.q_norm" + s + ".weight", dh, &L.qn[x], 1.f, 0.1f));
            UMV_TRY(alloc_vec(e, P + "self_attn.k_norm" + s + ".weight", dh, &L.kn[x], 1.f, 0.1f));
        }
    }
    if (d.enable_vit) {
        const int Dv = d.vit_hidden, Iv = d.vit_inter;
        const std::string V = "vit_model.vision_model.";
        UMV_TRY(alloc_linear(e, V + "embeddings.patch_embedding", Dv, d.vit_patch_dim, true, &e->vit_patch_w, &e->vit_patch_b,
                             e->vit_kpad));
        UMV_TRY(dev_alloc(e, &e->vit_pos, (size_t)d.vit_positions * Dv));
        reg(e, V + "embeddings.position_embedding.weight", e->vit_pos, d.vit_positions, Dv, Dv, 2);
        e->vit.resize(d.vit_layers);
        for (int li = 0; li < d.vit_layers; ++li) {
            VitLayerW& L = e->vit[li];
            const std::string P = V + "encoder.layers." + std::to_string(li) + ".";
            UMV_TRY(dev_alloc(e, &L.wqkv, (size_t)3 * Dv * Dv));
            UMV_TRY(dev_alloc(e, &L.bqkv, (size_t)3 * Dv));
            const char* nm[3] = {"q_proj", "k_proj", "v_proj"};
            for (int j = 0; j < 3; ++j) {
                reg(e, P + "self_attn." + nm[j] + ".weight", L.wqkv + (size_t)j * Dv * Dv, Dv, Dv, Dv, 2, SLOT_PLAIN,
                    std::sqrt(3.0f / Dv));
                reg(e, P + "self_attn." + nm[j] + ".bias", L.bqkv + j * Dv, 1, Dv, Dv, 1);
            }
            UMV_TRY(alloc_linear(e, P + "self_attn.out_proj", Dv, Dv, true, &L.wo, &L.bo));
            UMV_TRY(alloc_linear(e, P + "mlp.fc1", Iv, Dv, true, &L.w1, &L.b1));
            UMV_TRY(alloc_linear(e, P + "mlp.fc2", Dv, Iv, true, &L.w2, &L.b2));
            UMV_TRY(alloc_vec(e, P + "layer_norm1.weight", Dv, &L.ln1w, 1.f, 0.1f));
            UMV_TRY(alloc_vec(e, P + "layer_norm1.bias", Dv, &L.ln1b, 0.f, 0.05f));
            UMV_TRY(alloc_vec(e, P + "layer_norm2.weight", Dv, &L.ln2w, 1.f, 0.1f));
            UMV_TRY(alloc_vec(e, P + "layer_norm2.bias", Dv, &L.ln2b, 0.f, 0.05f));
        }
        UMV_TRY(alloc_vec(e, V + "post_layernorm.weight", Dv, &e->vit_post_w, 1.f, 0.1f));
        UMV_TRY(alloc_vec(e, V + "post_layernorm.bias", Dv, &e->vit_post_b, 0.f, 0.05f));
        UMV_TRY(alloc_linear(e, "connector.fc1", D, Dv, true, &e->conn_w1, &e->conn_b1));
        UMV_TRY(alloc_linear(e, "connector.fc2", D, D, true, &e->conn_w2, &e->conn_b2));
        UMV_TRY(dev_alloc(e, &e->vit_pos_embed, (size_t)d.vit_pos_table * D));
        reg(e, "vit_pos_embed.pos_embed", e->vit_pos_embed, d.vit_pos_table, D, D, 2, SLOT_PLAIN, 1.0f);
    }
    if (d.enable_gen) {
        UMV_TRY(alloc_linear(e, "time_embedder.mlp.0", D, 256, true, &e->t_w0, &e->t_b0));
        UMV_TRY(alloc_linear(e, "time_embedder.mlp.2", D, D, true, &e->t_w2, &e->t_b2));
        UMV_TRY(alloc_linear(e, "vae2llm", D, d.latent_dim, true, &e->vae2llm_w, &e->vae2llm_b));
        UMV_TRY(alloc_linear(e, "llm2vae", d.latent_dim, D, true, &e->llm2vae_w, &e->llm2vae_b));
        UMV_TRY(dev_alloc(e, &e->latent_pos, (size_t)d.latent_pos_table * D));
        reg(e, "latent_pos_embed.pos_embed", e->latent_pos, d.latent_pos_table, D, D, 2, SLOT_PLAIN, 1.0f);
    }
    return UMV_OK;
}

static int build_runtime(umv_engine* e) {
    const umv_dims& d = e->d;
    const int D = d.hidden, Dv = d.enable_vit ? d.vit_hidden : 0, Iv = d.enable_vit ? d.vit_inter : 0;
    const int Mmax = d.max_tokens;
    e->w_h = std::max(D, Dv);
    e->w_qkv = std::max(e->qkvn, 3 * Dv);
    e->w_act = std::max(std::max(d.inter, Iv), D);
    UMV_TRY(dev_alloc(e, &e->h, (size_t)Mmax * e->w_h));
    UMV_TRY(dev_alloc(e, &e->xn, (size_t)Mmax * e->w_h));
    UMV_TRY(dev_alloc(e, &e->qkv, (size_t)Mmax * e->w_qkv));
    UMV_TRY(dev_alloc(e, &e->attn, (size_t)Mmax * e->w_h));
    UMV_TRY(dev_alloc(e, &e->rope_tab, (size_t)Mmax * e->dh));
    if (d.enable_vit && d.vit_heads > 0) {
        const int dhv = d.vit_hidden / d.vit_heads;
        if (dhv % 8 == 0 && dhv < 128) {           // pad columns are zeroed once and never written again
            const size_t n = (size_t)Mmax * 3 * d.vit_heads * 128;
            UMV_TRY(dev_alloc(e, &e->vit_qkvp, n));
            UMV_CUDA_OK(cudaMemset(e->vit_qkvp, 0, n * sizeof(bf16)));
        }
    }
    UMV_TRY(dev_alloc(e, &e->act, (size_t)Mmax * e->w_act));
    UMV_TRY(dev_alloc(e, &e->logits, (size_t)64 * d.vocab));
    const int Tmax = std::min(Mmax, 8 * d.max_seqs);
    UMV_TRY(dev_alloc(e, &e->xt, (size_t)Tmax * std::max(D, d.inter)));
    UMV_TRY(dev_alloc(e, &e->ht, (size_t)Tmax * D));
    UMV_TRY(dev_alloc(e, &e->yt, (size_t)Tmax * std::max(e->qkvn, D)));
    UMV_TRY(dev_alloc(e, &e->actt, (size_t)Tmax * d.inter));
    e->ws_elems = (size_t)16 * 64 * std::max(e->qkvn, D);
    UMV_TRY(dev_alloc(e, &e->ws, e->ws_elems));
    e->attn_ws_elems = (size_t)32 * 64 * d.heads * (e->dh + 1);
    UMV_TRY(dev_alloc(e, &e->attn_ws, e->attn_ws_elems));
    // RoPE inverse frequencies: ROPE_INIT_FUNCTIONS['default'] -> 1 / theta^(2i/dh), fp32 (never bf16)
    std::vector<float> inv(e->dh / 2);
    for (int i = 0; i < e->dh / 2; ++i) {
        const float ex = (float)(2 * i) / (float)e->dh;          // arange(0,dh,2).float() / dh
        inv[i] = 1.0f / powf(d.rope_theta, ex);                  // fp32 pow, as torch does on the host
    }
    UMV_TRY(dev_alloc(e, &e->inv_freq, inv.size()));
    cudaMemcpy(e->inv_freq, inv.data(), inv.size() * sizeof(float), cudaMemcpyHostToDevice);
    // KV pool
    e->pool.layers = d.layers;
    e->pool.pages = d.kv_pages;
    e->pool.kv_heads = d.kv_heads;
    e->pool.head_dim = e->dh;
    const size_t page_elems = (size_t)d.layers * 2 * d.kv_heads * kPageTokens * e->dh;
    UMV_TRY(dev_alloc(e, &e->pool.base, page_elems * d.kv_pages));
    // Zero the pool once: the decode kernel brings whole 64-slot tiles in by TMA and multiplies not-yet-written V rows by
    // probability 0 -- they must never hold NaN / Inf bit patterns.
    UMV_CUDA_OK(cudaMemset(e->pool.base, 0, page_elems * d.kv_pages * sizeof(bf16)));
    e->kv_tmap_ok = false;
    if (e->dh == 128 && gemm_init() == UMV_OK) {
        // the pool as a 2-D tensor [slot rows, 128]: one 64 x 64 box = half (64 columns) of a K or V tile
        e->kv_tmap_ok = make_tmap_2d(&e->kv_tmap, e->pool.base, (uint64_t)d.kv_pages * d.layers * 2 * d.kv_heads * kPageTokens, 128,
                                     128, kPageTokens) == UMV_OK;
    }
    e->page_ref.assign(d.kv_pages, 0);
    e->free_pages.clear();
    for (int p = d.kv_pages - 1; p >= 0; --p) e->free_pages.push_back(p);
    // metadata staging
    e->meta_bytes = (size_t)64 * 1024 + (size_t)28 * Mmax + (size_t)4 * d.max_seqs * 3 * (std::min(d.kv_pages, 8192) + 8);
    for (int i = 0; i < umv_engine::kMetaRing; ++i) {
        if (cudaMallocHost(reinterpret_cast<void**>(&e->meta_host[i]), e->meta_bytes) != cudaSuccess) {
            set_error("cudaMallocHost(%zu) failed", e->meta_bytes);
            return UMV_ERR_NOMEM;
        }
        UMV_TRY(dev_alloc(e, &e->meta_dev[i], e->meta_bytes));
        cudaEventCreateWithFlags(&e->meta_ev[i], cudaEventDisableTiming);
    }
    // flow scratch: timestep frequencies exp(-ln(10000) * i / 128) in fp32 (modeling_utils.py:96-99)
    UMV_TRY(dev_alloc(e, &e->flow_small, (size_t)4 * std::max(d.hidden, 256)));
    {
        std::vector<float> f(128);
        for (int i = 0; i < 128; ++i) f[i] = expf(-logf(10000.0f) * (float)i / 128.0f);
        UMV_TRY(dev_alloc(e, &e->t_freqs, f.size()));
        cudaMemcpy(e->t_freqs, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice);
    }
    // decode state
    UMV_TRY(dev_alloc(e, &e->dec_tokens, 64));
    UMV_TRY(dev_alloc(e, &e->dec_pos, 64));
    UMV_TRY(dev_alloc(e, &e->dec_kvlen, 64));
    UMV_TRY(dev_alloc(e, &e->dec_kvpos, 64));
    UMV_TRY(dev_alloc(e, &e->dec_rope, (size_t)64 * e->dh));
    UMV_TRY(dev_alloc(e, &e->dec_rowseq, 64));
    UMV_TRY(dev_alloc(e, &e->dec_qstart, 65));
    UMV_TRY(dev_alloc(e, &e->dec_qlen, 64));
    UMV_TRY(dev_alloc(e, &e->dec_step, 1));
    e->dec_pages_cap = 64 * std::min(d.kv_pages, 4096);
    UMV_TRY(dev_alloc(e, &e->dec_pages, (size_t)e->dec_pages_cap));
    return UMV_OK;
}

// ------------------------------------------------------------------------------ KV sequences
static int page_alloc(umv_engine* e, int* page) {
    if (e->free_pages.empty()) {
        set_error("KV page pool exhausted (%d pages of %d tokens)", e->d.kv_pages, kPageTokens);
        return UMV_ERR_NOMEM;
    }
    *page = e->free_pages.back();
    e->free_pages.pop_back();
    e->page_ref[*page] = 1;
    return UMV_OK;
}
static void page_release(umv_engine* e, int page) {
    if (--e->page_ref[page] == 0) e->free_pages.push_back(page);
}
static Seq* get_seq(umv_engine* e, int id) {
    if (id < 0 || id >= (int)e->seqs.size() || !e->seqs[id].alive) {
        set_error("invalid sequence id %d", id);
        return nullptr;
    }
    return &e->seqs[id];
}
// Make slots [len, new_len) of `s` writable: private partial tail page (copy-on-write) + enough pages.
static int seq_reserve(umv_engine* e, Seq* s, int new_len, cudaStream_t st) {
    if (new_len <= s->len) return UMV_OK;
    if (s->len % kPageTokens != 0) {
        const int tail = s->len / kPageTokens;
        if (e->page_ref[s->pages[tail]] > 1) {
            int np;
            UMV_TRY(page_alloc(e, &np));
            UMV_TRY(copy_page(e->pool, s->pages[tail], np, st));
            page_release(e, s->pages[tail]);
            s->pages[tail] = np;
        }
    }
    const int need = (new_len + kPageTokens - 1) / kPageTokens;
    // pages past the committed length may be scratch left by a no-update forward; keep them only if private
    const int committed = (s->len + kPageTokens - 1) / kPageTokens;
    for (int p = committed; p < (int)s->pages.size(); ++p) {
        if (e->page_ref[s->pages[p]] > 1) {
            int np;
            UMV_TRY(page_alloc(e, &np));
            page_release(e, s->pages[p]);
            s->pages[p] = np;
        }
    }
    while ((int)s->pages.size() < need) {
        int np;
        UMV_TRY(page_alloc(e, &np));
        s->pages.push_back(np);
    }
    return UMV_OK;
}

// ------------------------------------------------------------------------------ call metadata
struct MetaBuilder {
    umv_engine* e;
    uint8_t* host;
    uint8_t* dev;
    size_t off = 0;
    int slot;
    template <typename T>
    T* put(const T* src, size_t n, T** dev_ptr) {
        off = (off + 15) & ~size_t(15);
        T* h = reinterpret_cast<T*>(host + off);
        if (src) memcpy(h, src, n * sizeof(T));
        *dev_ptr = reinterpret_cast<T*>(dev + off);
        off += n * sizeof(T);
        return h;
    }
};
static int meta_begin(umv_engine* e, MetaBuilder* mb, size_t need) {
    if (need + 4096 > e->meta_bytes) {
        set_error("call metadata (%zu bytes) exceeds the staging buffer (%zu); raise max_tokens / max_seqs", need, e->meta_bytes);
        return UMV_ERR_NOMEM;
    }
    mb->e = e;
    mb->slot = e->meta_next;
    e->meta_next = (e->meta_next + 1) % umv_engine::kMetaRing;
    cudaEventSynchronize(e->meta_ev[mb->slot]);     // previous use of this staging slot has been consumed
    mb->host = e->meta_host[mb->slot];
    mb->dev = e->meta_dev[mb->slot];
    mb->off = 0;
    return UMV_OK;
}
static int meta_commit(umv_engine* e, MetaBuilder* mb, cudaStream_t st) {
    UMV_CUDA_OK(cudaMemcpyAsync(mb->dev, mb->host, mb->off, cudaMemcpyHostToDevice, st));
    UMV_CUDA_OK(cudaEventRecord(e->meta_ev[mb->slot], st));
    return UMV_OK;
}

// ------------------------------------------------------------------------------ LLM layers
struct LlmRun {
    int M = 0, n_seqs = 0, max_q_len = 0, max_kv_len = 0;
    bool causal = true, gen = false, weight_major = false;
    bool seg = false;           // gen mode, segregated rows: [0, Mg) generation expert, [Mg, M) understanding expert (llm_run)
    int Mg = 0;
    CallMeta m;
    AttnProbe* probe = nullptr;
};

int lin(umv_engine* e, const bf16* x, int ldx, const bf16* w, const bf16* bias, const bf16* res, bf16* y, int ldy, int M,
               int N, int K, int epi, cudaStream_t st, int impl, float* ws, int splits, int stages, const int* res_rows,
               bf16* res_gather_tmp) {
    // res_rows: the residual of row i is res[res_rows[i]] (row stride ldy).  The split-K finish kernel reads it in place;
    // every other path gets the rows gathered into res_gather_tmp first.
    LinearCall c;
    c.stages = stages;
    c.x = x; c.ldx = ldx; c.w = w; c.bias = bias; c.residual = res; c.y = y; c.ldy = ldy;
    if (M <= 64 && !e->tiled.empty()) {
        auto it = e->tiled.find(w);
        if (it != e->tiled.end()) c.w_tiled = it->second;
    }
    c.M = M; c.N = N; c.K = K; c.epi = epi; c.ws = ws; c.splits = splits;
    c.impl = e->gemm_impl ? e->gemm_impl : impl;
    if (c.impl == GEMM_SIMPLE && epi == EPI_PARTIAL) c.impl = GEMM_WEIGHT_MAJOR;
    // A few rows against a weight matrix with far fewer 128-row tiles than SMs (o_proj / down_proj / q,k,v of the
    // understanding-expert rows in a generation-mode forward, the time-embedding MLP): the weights would stream through a
    // fraction of the SMs.  Split K across the machine into fp32 partials and finish with a tiny reduce + epilogue kernel.
    if ((c.impl == GEMM_AUTO || c.impl == GEMM_WEIGHT_MAJOR) && M <= 64 && e->use_splitk && e->ws && ws == nullptr &&
        (epi == EPI_BF16 || epi == EPI_GELU || epi == EPI_RESID) && K % 8 == 0 && ldx % 8 == 0 && N % 8 == 0 && ldy % 8 == 0 &&
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 &&
        (N + 127) / 128 <= e->sm_count / 2) {
        const int s = pick_splits(N, K, e->sm_count);
        if (s > 1 && (size_t)s * M * N <= e->ws_elems) {
            c.impl = GEMM_WEIGHT_MAJOR; c.epi = EPI_PARTIAL; c.ws = e->ws; c.splits = s;
            c.bias = nullptr; c.residual = nullptr; c.y = nullptr;
            UMV_TRY(linear_forward(c, st));
            return splitk_finish(e->ws, s, M, N, bias, res, y, ldy, epi, st, res_rows);
        }
    }
    if (res_rows) {
        UMV_REQUIRE(res_gather_tmp != nullptr, UMV_ERR_INVALID, "lin: residual row map without a gather buffer");
        UMV_TRY(copy_rows(res, ldy, res_rows, res_gather_tmp, ldy, M, N, 0, st));
        c.residual = res_gather_tmp;
    }
    return linear_forward(c, st);
}

static int llm_layers(umv_engine* e, const LlmRun& r, bf16* out, cudaStream_t st) {
    const umv_dims& d = e->d;
    const int D = d.hidden, I = d.inter, H = d.heads, Hkv = d.kv_heads, dh = e->dh, QN = e->qkvn, M = r.M;
    const int E = r.gen ? 1 : 0;
    const bool partial = r.weight_major && !r.gen && e->use_splitk && e->gemm_impl != GEMM_SIMPLE;
    const int T = r.gen ? r.m.n_text : 0;
    // segregated gen-mode rows (llm_run): each expert's linears run on its own contiguous slice -- the generation expert never touches
    // the marker rows (256-row tiles stay full: 12 x 256 latent rows instead of 13 tiles for 3,096 packed rows) and the understanding
    // expert's few rows need no gather / scatter
    const bool seg = r.seg;
    const int Mg = seg ? r.Mg : M;
    int pending_splits = 0;     // >0: e->ws holds split-K partials of the last residual-branch linear

    static const int attn_cluster_max = getenv("UMV_ATTN_CLUSTER") ? atoi(getenv("UMV_ATTN_CLUSTER")) : 8;
    auto norm = [&](const bf16* w0, const bf16* w1, bf16* y) {
        AddNormArgs a;

        a.h = e->h; a.M = M; a.D = D; a.eps = d.rms_eps; a.w0 = w0; a.w1 = w1 ? w1 : w0;
        a.row_sel = r.gen ? r.m.row_sel : nullptr;
        a.y = y;
        if (T > 0 && y == e->xn && !seg) { a.y2 = e->xt; a.row_slot = r.m.text_slot; }     // text rows also land gathered in xt
        if (pending_splits > 0) { a.partial = e->ws; a.splits = pending_splits; }
        pending_splits = 0;
        return add_rmsnorm(a, st);
    };

    int attn_splits = 1;
    if (r.weight_major && r.max_q_len == 1) {
        const int blocks = (r.max_kv_len + kPageTokens - 1) / kPageTokens;
        static const int max_splits = getenv("UMV_ATTN_SPLITS") ? atoi(getenv("UMV_ATTN_SPLITS")) : 16;
        static const int waves = getenv("UMV_ATTN_WAVES") ? atoi(getenv("UMV_ATTN_WAVES")) : 2;
        attn_splits = std::max(1, std::min(std::min(max_splits, blocks), (waves * e->sm_count + r.n_seqs * Hkv - 1) / (r.n_seqs * Hkv)));
    }

    // q/k RMSNorm + RoPE + KV append + attention over the paged cache (past + the rows just appended) of layer li; `ra` names the
    // projection outputs (bf16 rows or split-K partials + bias).  Returns through *path which kernel family ran.
    auto attention_block = [&](int li, RopeAppendArgs& ra, int* path) -> int {
        const LayerW& L = e->layers[li];
        // decode (one query token per sample, understanding expert): the whole rope -> append -> attention -> combine
        // chain is one cluster launch
        const char* fa_env = getenv("UMV_FUSED_ATTN");        // read per call: tests switch paths inside one process
        const bool fused_attn = !(fa_env && atoi(fa_env) == 0);
        const bool fuse = fused_attn && r.weight_major && r.max_q_len == 1 && !r.gen && e->kv_tmap_ok &&
                          decode_attention_supported(H, Hkv, dh, r.m.max_pages, ra.splits);
        if (fuse) {
            DecodeAttnArgs da;
            da.qkv = ra.qkv; da.partial = ra.partial; da.ksplits = ra.splits; da.bias = ra.bias;
            da.out = e->attn; da.ldo = D; da.positions = r.m.positions; da.kv_len = r.m.kv_len;
            da.page_table = r.m.page_table; da.max_pages = r.m.max_pages; da.inv_freq = e->inv_freq; da.rope_cs = r.m.rope_cs;
            da.qn = L.qn[0]; da.kn = L.kn[0]; da.pool = e->pool; da.layer = li; da.M = M; da.H = H; da.Hkv = Hkv;
            da.eps = d.rms_eps; da.kv_tmap = e->kv_tmap_ok ? &e->kv_tmap : nullptr;
            const int blocks = (r.max_kv_len + kPageTokens - 1) / kPageTokens;
            // key ranges per (sample, kv head): as many as fit one wave of 2 CTAs per SM, at most one per 64-key block
            da.cluster = std::max(1, std::min(std::min(attn_cluster_max, blocks), e->sm_count / std::max(1, M * Hkv)));
            // K/V tiles requested before the dependency wait: from layer 1 on (layer 0 follows the step's own state kernels, whose
            // kv_len / cos-sin writes the early reads must not overtake).  UMV_ATTN_EARLY=0 off, 2 = also layer 0 (op-level tests).
            const char* early_env = getenv("UMV_ATTN_EARLY");         // read per call: tests switch inside one process
            const int early_mode = early_env ? atoi(early_env) : 1;
            da.early = early_mode == 2 ? 1 : (early_mode == 1 && li > 0 ? 1 : 0);
            if (path) *path = 3;
            return decode_attention(da, st);
        }
        ra.q_out = e->qkv; ra.ldq = QN;
        if (seg) { ra.q_out = e->act; ra.ldq = H * dh; ra.q_row_map = r.m.seg_to_packed; }    // queries to their packed rows (e->act is idle here)
        ra.positions = r.m.positions; ra.row_seq = r.m.row_seq; ra.row_kvpos = r.m.row_kvpos;
        ra.page_table = r.m.rope_page_table; ra.max_pages = r.m.max_pages; ra.inv_freq = e->inv_freq; ra.rope_cs = r.m.rope_cs;
        ra.qn0 = L.qn[0]; ra.kn0 = L.kn[0]; ra.qn1 = L.qn[E]; ra.kn1 = L.kn[E];
        ra.row_sel = r.gen ? r.m.row_sel : nullptr; ra.gen_mode = r.gen ? 1 : 0;
        ra.pool = e->pool; ra.layer = li; ra.M = M; ra.H = H; ra.Hkv = Hkv; ra.dh = dh; ra.eps = d.rms_eps;
        UMV_TRY(rope_append(ra, st));
        AttnArgs aa;
        aa.q = ra.q_out; aa.ldq = ra.ldq; aa.out = e->attn; aa.ldo = D;
        aa.out_row_map = seg ? r.m.packed_to_seg : nullptr;
        aa.paged = 1; aa.pool = e->pool; aa.layer = li; aa.page_table = r.m.page_table; aa.max_pages = r.m.max_pages;
        aa.q_start = r.m.q_start; aa.q_len = r.m.q_len; aa.kv_len = r.m.kv_len;
        aa.n = r.n_seqs; aa.H = H; aa.Hkv = Hkv; aa.dh = dh; aa.causal = r.causal ? 1 : 0;
        aa.max_q_len = r.max_q_len; aa.max_kv_len = r.max_kv_len; aa.splits = attn_splits; aa.ws = e->attn_ws; aa.total_q = M;
        aa.kv_tmap = e->kv_tmap_ok ? &e->kv_tmap : nullptr;
        if (path) {
            const char* tc_env = getenv("UMV_ATTN_TC");
            *path = (!(tc_env && atoi(tc_env) == 0) && attention_tc_supported(aa)) ? 2 : 1;
        }
        if (r.m.n_tail > 0) {        // causal tails (llm_run): the block groups under the call's mask, then the tail groups causally
            if (aa.n > 0) UMV_TRY(attention_forward(aa, st));
            aa.q_start = r.m.tail_q_start; aa.q_len = r.m.tail_q_len; aa.kv_len = r.m.tail_kv_len; aa.page_table = r.m.tail_page_table;
            aa.n = r.m.n_tail; aa.max_q_len = r.m.tail_max_q; aa.causal = 1;
        }
        return attention_forward(aa, st);
    };

    if (r.probe) {      // umv_op_attention_block: this layer's attention block alone, on the caller's projection outputs
        AttnProbe& p = *r.probe;
        RopeAppendArgs ra;
        if (p.partial) {
            ra.partial = p.partial; ra.splits = p.splits; ra.bias = p.bias;
        } else {
            UMV_CUDA_OK(cudaMemcpyAsync(e->qkv, p.qkv, (size_t)M * QN * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
            ra.qkv = e->qkv;
        }
        UMV_TRY(attention_block(p.layer, ra, &p.path));
        UMV_CUDA_OK(cudaMemcpyAsync(p.out, e->attn, (size_t)M * D * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
        return UMV_OK;
    }

    for (int li = 0; li < d.layers; ++li) {
        const LayerW& L = e->layers[li];
        UMV_TRY(norm(L.ln1[0], L.ln1[1], e->xn));
        // ---- q/k/v projections
        RopeAppendArgs ra;
        if (partial) {
            const int s = pick_splits(QN, D, e->sm_count);
            UMV_TRY(lin(e, e->xn, D, L.wqkv[0], nullptr, nullptr, nullptr, 0, M, QN, D, EPI_PARTIAL, st, GEMM_WEIGHT_MAJOR, e->ws, s));
            ra.partial = e->ws; ra.splits = s; ra.bias = L.bqkv[0];
        } else {
            UMV_TRY(lin(e, e->xn, D, L.wqkv[E], L.bqkv[E], nullptr, e->qkv, QN, Mg, QN, D, EPI_BF16, st));
            if (T > 0 && seg) {
                UMV_TRY(lin(e, e->xn + (size_t)Mg * D, D, L.wqkv[0], L.bqkv[0], nullptr, e->qkv + (size_t)Mg * QN, QN, T, QN, D, EPI_BF16, st));
            } else if (T > 0) {      // xt = text rows of xn (written by the norm kernel)
                UMV_TRY(lin(e, e->xt, D, L.wqkv[0], L.bqkv[0], nullptr, e->yt, QN, T, QN, D, EPI_BF16, st));
                UMV_TRY(copy_rows(e->yt, QN, r.m.text_rows, e->qkv, QN, T, QN, 1, st));
            }
            ra.qkv = e->qkv;
        }
        UMV_TRY(attention_block(li, ra, nullptr));
        // ---- output projection + residual
        if (partial) {
            const int s = pick_splits(D, D, e->sm_count);
            UMV_TRY(lin(e, e->attn, D, L.wo[0], nullptr, nullptr, nullptr, 0, M, D, D, EPI_PARTIAL, st, GEMM_WEIGHT_MAJOR, e->ws, s));
            pending_splits = s;
        } else {
            bf16* ht = e->h + (size_t)Mg * D;       // segregated: the understanding-expert rows of the residual stream
            if (T > 0 && seg) {
                UMV_TRY(lin(e, e->attn + (size_t)Mg * D, D, L.wo[0], nullptr, ht, ht, D, T, D, D, EPI_RESID, st));
            } else if (T > 0) {
                UMV_TRY(copy_rows(e->attn, D, r.m.text_rows, e->xt, D, T, D, 0, st));
                UMV_TRY(lin(e, e->xt, D, L.wo[0], nullptr, e->h, e->yt, D, T, D, D, EPI_RESID, st, 0, nullptr, 1, 0, r.m.text_rows, e->ht));
            }
            UMV_TRY(lin(e, e->attn, D, L.wo[E], nullptr, e->h, e->h, D, Mg, D, D, EPI_RESID, st));
            if (T > 0 && !seg) UMV_TRY(copy_rows(e->yt, D, r.m.text_rows, e->h, D, T, D, 1, st));
        }
        UMV_TRY(norm(L.ln2[0], L.ln2[1], e->xn));
        // ---- SwiGLU MLP + residual
        if (T > 0 && seg) {
            bf16* ht = e->h + (size_t)Mg * D;
            UMV_TRY(lin(e, e->xn + (size_t)Mg * D, D, L.wgu[0], nullptr, nullptr, e->act + (size_t)Mg * I, I, T, 2 * I, D, EPI_SWIGLU, st));
            UMV_TRY(lin(e, e->act + (size_t)Mg * I, I, L.wdown[0], nullptr, ht, ht, D, T, D, I, EPI_RESID, st));
        } else if (T > 0) {
            UMV_TRY(lin(e, e->xt, D, L.wgu[0], nullptr, nullptr, e->actt, I, T, 2 * I, D, EPI_SWIGLU, st));
            UMV_TRY(lin(e, e->actt, I, L.wdown[0], nullptr, e->h, e->yt, D, T, D, I, EPI_RESID, st, 0, nullptr, 1, 0, r.m.text_rows, e->ht));
        }
        UMV_TRY(lin(e, e->xn, D, L.wgu[E], nullptr, nullptr, e->act, I, Mg, 2 * I, D, EPI_SWIGLU, st));
        if (partial) {
            const int s = pick_splits(D, I, e->sm_count);
            UMV_TRY(lin(e, e->act, I, L.wdown[0], nullptr, nullptr, nullptr, 0, M, D, I, EPI_PARTIAL, st, GEMM_WEIGHT_MAJOR, e->ws, s));
            pending_splits = s;
        } else {
            UMV_TRY(lin(e, e->act, I, L.wdown[E], nullptr, e->h, e->h, D, Mg, D, I, EPI_RESID, st));
            if (T > 0 && !seg) UMV_TRY(copy_rows(e->yt, D, r.m.text_rows, e->h, D, T, D, 1, st));
        }
    }
    // final norm (norm / norm_moe_gen, qwen2_navit.py:1162-1169); also folds the last down-proj partials
    return norm(e->final_norm[0], e->final_norm[1], out ? out : e->xn);
}

}  // namespace umv

using namespace umv;

// =============================================================================================
//                                           C ABI
// =============================================================================================
extern "C" {

const char* umv_last_error(void) { return get_error(); }
int umv_abi_version(void) { return UMV_ABI_VERSION; }
int64_t umv_launch_count(void) { return g_launches; }
int umv_trace_begin(int32_t max_slots) { return trace_begin(max_slots); }
int umv_trace_read(uint64_t* stamps, char* names, int32_t name_len, int32_t max_slots, int32_t* n) {
    UMV_REQUIRE(stamps && names && n && name_len > 0, UMV_ERR_INVALID, "umv_trace_read: null argument");
    int k = 0;
    int rc = trace_read(reinterpret_cast<unsigned long long*>(stamps), names, name_len, max_slots, &k);
    *n = k;
    return rc;
}

int umv_create(const umv_dims* dims, umv_engine** out) {
    UMV_REQUIRE(dims && out, UMV_ERR_INVALID, "umv_create: null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("umv_create: no CUDA device (this engine has no CPU fallback)");
        return UMV_ERR_CUDA;
    }
    const umv_dims& d = *dims;
    UMV_REQUIRE(d.hidden > 0 && d.heads > 0 && d.kv_heads > 0 && d.hidden % d.heads == 0 && d.heads % d.kv_heads == 0,
                UMV_ERR_INVALID, "umv_create: bad head geometry");
    UMV_REQUIRE(d.hidden / d.heads == 128, UMV_ERR_UNSUPPORTED, "umv_create: LLM head_dim %d (128 is built)", d.hidden / d.heads);
    UMV_REQUIRE(d.inter % 64 == 0 && d.hidden % 8 == 0 && d.vocab % 8 == 0, UMV_ERR_UNSUPPORTED,
                "umv_create: inter %% 64, hidden %% 8 and vocab %% 8 must be 0");
    UMV_REQUIRE(!d.enable_vit || (d.vit_hidden % d.vit_heads == 0 && d.vit_hidden / d.vit_heads == 72 && d.vit_inter % 8 == 0),
                UMV_ERR_UNSUPPORTED, "umv_create: ViT head_dim must be 72 and vit_inter %% 8 == 0");
    UMV_REQUIRE(d.max_tokens > 0 && d.max_seqs > 0 && d.max_seqs <= 64 && d.kv_pages > 0, UMV_ERR_INVALID,
                "umv_create: max_tokens/max_seqs(<=64)/kv_pages must be positive");
    umv_engine* e = new umv_engine();
    e->d = d;
    e->dh = d.hidden / d.heads;
    e->qkvn = (d.heads + 2 * d.kv_heads) * e->dh;
    e->vit_kpad = (d.vit_patch_dim + 7) / 8 * 8;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (const char* v = getenv("UMV_SPLITK")) e->use_splitk = atoi(v) != 0;
    if (const char* v = getenv("UMV_GRAPH")) e->use_graph = atoi(v) != 0;
    if (const char* v = getenv("UMV_GEMM_IMPL")) e->gemm_impl = atoi(v);
    int rc = gemm_init();
    if (rc == UMV_OK) rc = attention_init();
    if (rc == UMV_OK) rc = build_weights(e);
    if (rc == UMV_OK) rc = build_runtime(e);
    if (rc == UMV_OK && d.enable_vae) rc = vae_build(e);
    if (rc != UMV_OK) {
        umv_destroy(e);
        return rc;
    }
    *out = e;
    return UMV_OK;
}

int umv_destroy(umv_engine* e) {
    if (!e) return UMV_OK;
    cudaDeviceSynchronize();
    for (void* p : e->allocs) cudaFree(p);
    for (void* p : e->tiled_allocs) cudaFree(p);
    for (auto& g : e->dec_graphs) cudaGraphExecDestroy(g.exec);
    if (e->dec_out) cudaFree(e->dec_out);
    delete e->vae;
    for (int i = 0; i < umv_engine::kMetaRing; ++i) {
        if (e->meta_host[i]) cudaFreeHost(e->meta_host[i]);
        if (e->meta_ev[i]) cudaEventDestroy(e->meta_ev[i]);
    }
    delete e;
    return UMV_OK;
}

static int slot_copy(umv_engine* e, Slot& s, void* host, bool to_device) {
    // geometry of the engine-side region, expressed as a 2-D pitched copy
    size_t width, height, dpitch, spitch = 0;
    bf16* dst = s.dst;
    if (s.kind == SLOT_PLAIN) {
        width = (size_t)s.cols * 2; height = (size_t)s.rows; dpitch = (size_t)s.dst_ld * 2; spitch = width;
    } else {   // gate/up: 64-row blocks of the source land every 128 rows
        width = (size_t)64 * s.cols * 2; height = (size_t)s.rows / 64; dpitch = (size_t)128 * s.dst_ld * 2; spitch = width;
    }
    cudaError_t err = to_device ? cudaMemcpy2D(dst, dpitch, host, spitch, width, height, cudaMemcpyDefault)
                                : cudaMemcpy2D(host, spitch, dst, dpitch, width, height, cudaMemcpyDeviceToHost);
    if (err != cudaSuccess) {
        set_error("weight copy failed: %s", cudaGetErrorString(err));
        return UMV_ERR_CUDA;
    }
    return UMV_OK;
}

int umv_load_tensor(umv_engine* e, const char* name, const void* data, int dtype, int ndim, const int64_t* shape,
                    int data_on_device) {
    UMV_REQUIRE(e && name && data && shape, UMV_ERR_INVALID, "umv_load_tensor: null argument");
    auto it = e->slots.find(name);
    UMV_REQUIRE(it != e->slots.end(), UMV_ERR_INVALID, "umv_load_tensor: unknown tensor '%s'", name);
    Slot& s = it->second;
    if (s.conv_k > 0) {
        // conv weight [cout, cin, k, k] -> engine layout [cout, k, k, cin] (tap-major, matches the NHWC im2col)
        UMV_REQUIRE(ndim == 4 && shape[0] == s.rows && shape[1] == s.conv_cin && shape[2] == s.conv_k && shape[3] == s.conv_k,
                    UMV_ERR_INVALID, "umv_load_tensor: '%s' expects a conv weight [%lld,%d,%d,%d]", name, (long long)s.rows,
                    s.conv_cin, s.conv_k, s.conv_k);
        UMV_REQUIRE(dtype == UMV_BF16 || (dtype == UMV_F32 && !data_on_device), UMV_ERR_UNSUPPORTED,
                    "umv_load_tensor: dtype must be bf16 (or host fp32)");
        const size_t n = (size_t)s.rows * s.cols;
        std::vector<bf16> src(n), dst(n);
        if (dtype == UMV_F32) {      // an fp32 ae.safetensors: round to bf16 on the way, as `.to(torch.bfloat16)` after load_ae does
            const float* f = static_cast<const float*>(data);
            for (size_t i = 0; i < n; ++i) src[i] = __float2bfloat16_rn(f[i]);
        } else {
            UMV_CUDA_OK(cudaMemcpy(src.data(), data, n * 2, cudaMemcpyDefault));
        }
        const int kk = s.conv_k * s.conv_k, cin = s.conv_cin;
        for (int64_t o = 0; o < s.rows; ++o)
            for (int c = 0; c < cin; ++c)
                for (int t = 0; t < kk; ++t) dst[(o * kk + t) * cin + c] = src[(o * cin + c) * kk + t];
        int rc2 = slot_copy(e, s, dst.data(), true);
        if (rc2 == UMV_OK) { s.loaded = true; e->finalized = false; }
        return rc2;
    }
    const int64_t rows = ndim == 2 ? shape[0] : 1, cols = ndim == 2 ? shape[1] : shape[0];
    UMV_REQUIRE(ndim == s.ndim && rows == s.rows && cols == s.cols, UMV_ERR_INVALID,
                "umv_load_tensor: '%s' expects shape [%lld,%lld] (ndim %d), got ndim %d [%lld,%lld]", name, (long long)s.rows,
                (long long)s.cols, s.ndim, ndim, (long long)rows, (long long)cols);
    UMV_REQUIRE(dtype == UMV_BF16 || (dtype == UMV_F32 && !data_on_device), UMV_ERR_UNSUPPORTED,
                "umv_load_tensor: dtype must be bf16 (or host fp32)");
    UMV_REQUIRE(s.kind == SLOT_PLAIN || s.rows % 64 == 0, UMV_ERR_UNSUPPORTED, "gate/up rows must be a multiple of 64");
    int rc;
    if (dtype == UMV_F32) {
        const size_t n = (size_t)rows * cols;
        std::vector<bf16> tmp(n);
        const float* f = static_cast<const float*>(data);
        for (size_t i = 0; i < n; ++i) tmp[i] = __float2bfloat16_rn(f[i]);
        rc = slot_copy(e, s, tmp.data(), true);
    } else {
        rc = slot_copy(e, s, const_cast<void*>(data), true);
    }
    if (rc == UMV_OK) { s.loaded = true; e->finalized = false; }      // umv_finalize re-derives the tile-major twins
    return rc;
}

int umv_export_tensor(umv_engine* e, const char* name, void* host_dst, size_t bytes) {
    UMV_REQUIRE(e && name && host_dst, UMV_ERR_INVALID, "umv_export_tensor: null argument");
    auto it = e->slots.find(name);
    UMV_REQUIRE(it != e->slots.end(), UMV_ERR_INVALID, "umv_export_tensor: unknown tensor '%s'", name);
    Slot& s = it->second;
    UMV_REQUIRE(bytes == (size_t)s.rows * s.cols * 2, UMV_ERR_INVALID, "umv_export_tensor: '%s' needs %zu bytes", name,
                (size_t)s.rows * s.cols * 2);
    if (s.conv_k > 0) {
        const size_t n = (size_t)s.rows * s.cols;
        std::vector<bf16> tmp(n);
        UMV_TRY(slot_copy(e, s, tmp.data(), false));
        bf16* out = static_cast<bf16*>(host_dst);
        const int kk = s.conv_k * s.conv_k, cin = s.conv_cin;
        for (int64_t o = 0; o < s.rows; ++o)
            for (int c = 0; c < cin; ++c)
                for (int t = 0; t < kk; ++t) out[(o * cin + c) * kk + t] = tmp[(o * kk + t) * cin + c];
        return UMV_OK;
    }
    return slot_copy(e, s, host_dst, false);
}

int umv_fill_synthetic(umv_engine* e, uint64_t seed) {
    UMV_REQUIRE(e, UMV_ERR_INVALID, "null engine");
    uint64_t k = 0;
    for (auto& kv : e->slots) {
        Slot& s = kv.second;
        const uint64_t sd = seed * 0x9E3779B97F4A7C15ull + (++k) * 0xD1B54A32D192ED03ull;
        if (s.kind == SLOT_PLAIN) {
            for (int64_t r = 0; r < s.rows && s.dst_ld != s.cols; ++r)
                UMV_TRY(fill_uniform_bf16(s.dst + r * s.dst_ld, (size_t)s.cols, sd + r, s.synth_bound, s.synth_mean, 0));
            if (s.dst_ld == s.cols) UMV_TRY(fill_uniform_bf16(s.dst, (size_t)s.rows * s.cols, sd, s.synth_bound, s.synth_mean, 0));
        } else {
            for (int64_t blk = 0; blk < s.rows / 64; ++blk)
                UMV_TRY(fill_uniform_bf16(s.dst + blk * 128 * s.dst_ld, (size_t)64 * s.cols, sd + blk, s.synth_bound, s.synth_mean, 0));
        }
        s.loaded = true;
    }
    UMV_CUDA_OK(cudaDeviceSynchronize());
    return UMV_OK;
}

int umv_finalize(umv_engine* e) {
    UMV_REQUIRE(e, UMV_ERR_INVALID, "null engine");
    for (auto& kv : e->slots)
        UMV_REQUIRE(kv.second.loaded, UMV_ERR_STATE, "umv_finalize: tensor '%s' was never loaded", kv.first.c_str());
    UMV_CUDA_OK(cudaDeviceSynchronize());
    // tile-major twins of what a decode step streams: the understanding expert's four linears per layer and lm_head
    for (void* p : e->tiled_allocs) cudaFree(p);
    e->tiled_allocs.clear();
    e->tiled.clear();
    const char* tl = getenv("UMV_TILED");
    if (!(tl && atoi(tl) == 0) && gemm_init() == UMV_OK) {
        const int D = e->d.hidden, I = e->d.inter, QN = e->qkvn;
        auto twin = [&](const bf16* w, int N, int K) -> int {
            if (!w || N % 128 != 0 || K % 64 != 0) return UMV_OK;
            void* p = nullptr;
            if (cudaMalloc(&p, (size_t)N * K * sizeof(bf16)) != cudaSuccess) {
                cudaGetLastError();
                return UMV_OK;                    // no room for the second copy: the row-major weights serve
            }
            e->tiled_allocs.push_back(p);
            UMV_TRY(tile_weights(w, static_cast<bf16*>(p), N, K, nullptr));
            e->tiled[w] = static_cast<bf16*>(p);
            return UMV_OK;
        };
        for (const LayerW& L : e->layers) {
            UMV_TRY(twin(L.wqkv[0], QN, D));
            UMV_TRY(twin(L.wo[0], D, D));
            UMV_TRY(twin(L.wgu[0], 2 * I, D));
            UMV_TRY(twin(L.wdown[0], D, I));
        }
        UMV_TRY(twin(e->lm_head, e->d.vocab, D));
        UMV_CUDA_OK(cudaDeviceSynchronize());
    }
    e->finalized = true;
    return UMV_OK;
}

int umv_weight_bytes(umv_engine* e, int64_t* bytes) {
    UMV_REQUIRE(e && bytes, UMV_ERR_INVALID, "null argument");
    int64_t n = 0;
    for (auto& kv : e->slots) n += kv.second.rows * kv.second.cols * 2;
    *bytes = n;
    return UMV_OK;
}

// ------------------------------------------------------------------------------ sequences
int umv_seq_new(umv_engine* e, int32_t* seq) {
    UMV_REQUIRE(e && seq, UMV_ERR_INVALID, "null argument");
    for (size_t i = 0; i < e->seqs.size(); ++i)
        if (!e->seqs[i].alive) {
            e->seqs[i] = Seq();
            e->seqs[i].alive = true;
            *seq = (int)i;
            return UMV_OK;
        }
    e->seqs.emplace_back();
    e->seqs.back().alive = true;
    *seq = (int)e->seqs.size() - 1;
    return UMV_OK;
}
int umv_seq_fork(umv_engine* e, int32_t src, int32_t* dst) {
    UMV_REQUIRE(e && dst, UMV_ERR_INVALID, "null argument");
    Seq* s = get_seq(e, src);
    if (!s) return UMV_ERR_INVALID;
    const int committed = (s->len + kPageTokens - 1) / kPageTokens;
    std::vector<int> pages(s->pages.begin(), s->pages.begin() + committed);
    const int len = s->len;
    UMV_TRY(umv_seq_new(e, dst));      // may reallocate e->seqs
    Seq& n = e->seqs[*dst];
    n.pages = pages;
    n.len = len;
    n.content = e->seqs[src].content;          // same committed tokens, same K / V
    for (int p : n.pages) ++e->page_ref[p];
    return UMV_OK;
}
int umv_seq_free(umv_engine* e, int32_t seq) {
    UMV_REQUIRE(e, UMV_ERR_INVALID, "null engine");
    Seq* s = get_seq(e, seq);
    if (!s) return UMV_ERR_INVALID;
    for (int p : s->pages) page_release(e, p);
    *s = Seq();
    return UMV_OK;
}
int umv_seq_len(umv_engine* e, int32_t seq, int32_t* len) {
    UMV_REQUIRE(e && len, UMV_ERR_INVALID, "null argument");
    Seq* s = get_seq(e, seq);
    if (!s) return UMV_ERR_INVALID;
    *len = s->len;
    return UMV_OK;
}
int umv_seq_truncate(umv_engine* e, int32_t seq, int32_t len) {
    UMV_REQUIRE(e, UMV_ERR_INVALID, "null engine");
    Seq* s = get_seq(e, seq);
    if (!s) return UMV_ERR_INVALID;
    UMV_REQUIRE(len >= 0 && len <= s->len, UMV_ERR_INVALID, "umv_seq_truncate: %d not in [0,%d]", len, s->len);
    if (len != s->len) s->content = len ? ++e->content_counter : 0;
    s->len = len;
    const int keep = (len + kPageTokens - 1) / kPageTokens;
    while ((int)s->pages.size() > keep) {
        page_release(e, s->pages.back());
        s->pages.pop_back();
    }
    return UMV_OK;
}
int umv_pages_free(umv_engine* e, int32_t* n) {
    UMV_REQUIRE(e && n, UMV_ERR_INVALID, "null argument");
    *n = (int)e->free_pages.size();
    return UMV_OK;
}
int umv_seq_export(umv_engine* e, int32_t seq, int32_t layer, void* k_dev, void* v_dev, void* stream) {
    UMV_REQUIRE(e && k_dev && v_dev, UMV_ERR_INVALID, "null argument");
    Seq* s = get_seq(e, seq);
    if (!s) return UMV_ERR_INVALID;
    UMV_REQUIRE(layer >= 0 && layer < e->d.layers, UMV_ERR_INVALID, "bad layer %d", layer);
    if (s->len == 0) return UMV_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MetaBuilder mb;
    UMV_TRY(meta_begin(e, &mb, s->pages.size() * 4));
    int* dpages;
    mb.put<int>(s->pages.data(), s->pages.size(), &dpages);
    UMV_TRY(meta_commit(e, &mb, st));
    return export_kv(e->pool, layer, dpages, s->len, static_cast<bf16*>(k_dev), static_cast<bf16*>(v_dev), st);
}

// ------------------------------------------------------------------------------ forward passes
int umv_patchify_u8(const uint8_t* images, const int64_t* offsets, const int32_t* hw, int32_t n_images, int32_t patch,
                    int32_t max_per_side, float* out_pixels, int64_t* out_pos_ids, void* stream) {
    UMV_REQUIRE(images && offsets && hw && out_pixels && out_pos_ids && n_images > 0 && patch > 0, UMV_ERR_INVALID,
                "umv_patchify_u8: null/empty argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    size_t tok = 0;
    for (int i = 0; i < n_images; ++i) {
        const int H = hw[2 * i], W = hw[2 * i + 1];
        UMV_REQUIRE(H > 0 && W > 0 && H % patch == 0 && W % patch == 0, UMV_ERR_INVALID,
                    "umv_patchify_u8: image %d is %dx%d, not a multiple of the %d-pixel patch (resize first)", i, H, W, patch);
        UMV_REQUIRE(H / patch <= max_per_side && W / patch <= max_per_side, UMV_ERR_INVALID,
                    "umv_patchify_u8: image %d has more than %d patches per side", i, max_per_side);
        UMV_TRY(patchify_u8(images + offsets[i], H, W, patch, max_per_side, out_pixels + tok * 3 * patch * patch, out_pos_ids + tok, st));
        tok += (size_t)(H / patch) * (W / patch);
    }
    return UMV_OK;
}

int umv_embed_tokens(umv_engine* e, const int64_t* ids, int32_t n, void* out, void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    return embed_rows(e->embed, ids, n, e->d.hidden, e->d.vocab, static_cast<bf16*>(out), static_cast<cudaStream_t>(stream));
}

int umv_llm_forward(umv_engine* e, const void* x, int32_t n_seqs, const int32_t* seqs, const int32_t* q_lens,
                    const int32_t* positions, const uint8_t* row_is_gen, int32_t is_causal, int32_t update_kv, void* out,
                    void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(x, UMV_ERR_INVALID, "umv_llm_forward: null x");
    return umv::llm_run(e, static_cast<const bf16*>(x), n_seqs, seqs, q_lens, positions, row_is_gen, is_causal, update_kv,
                        static_cast<bf16*>(out), static_cast<cudaStream_t>(stream));
}

}  // extern "C"

namespace umv {
bool gen_rows_segregate(int n_seqs, const int32_t* q_lens, const uint8_t* row_is_gen, int is_causal, int* n_gen) {
    if (n_gen) *n_gen = 0;
    if (!row_is_gen || is_causal) return false;
    const char* env = getenv("UMV_GEN_SEG");              // read per call: tests switch layouts inside one process
    if (env && atoi(env) == 0) return false;
    int M = 0, ng = 0;
    for (int b = 0; b < n_seqs; ++b) M += q_lens[b];
    for (int i = 0; i < M; ++i) ng += row_is_gen[i] ? 1 : 0;
    const int nt = M - ng;
    if (n_gen) *n_gen = ng;
    return ng > 0 && nt > 0;
}

// Packed forward over n_seqs sequences; x == nullptr means the packed query sequence is already in e->h (presegregated: in the
// segregated row order of gen_rows_segregate, and `out` is wanted in that order too -- the flow step composes and consumes it so).
int llm_run(umv_engine* e, const bf16* x, int n_seqs, const int32_t* seqs, const int32_t* q_lens, const int32_t* positions,
            const uint8_t* row_is_gen, int is_causal, int update_kv, bf16* out, cudaStream_t st, AttnProbe* probe, int presegregated,
            const int32_t* causal_tail) {
    UMV_REQUIRE(seqs && q_lens && positions && n_seqs > 0, UMV_ERR_INVALID, "umv_llm_forward: null/empty argument");
    UMV_REQUIRE(n_seqs <= 3 * e->d.max_seqs, UMV_ERR_INVALID, "umv_llm_forward: %d sequences > 3*max_seqs", n_seqs);
    UMV_REQUIRE(!row_is_gen || e->d.enable_gen, UMV_ERR_STATE, "generation expert weights were not enabled");
    const int D = e->d.hidden;
    LlmRun r;
    r.probe = probe;
    r.causal = is_causal != 0;
    r.gen = row_is_gen != nullptr;
    int M = 0;
    for (int b = 0; b < n_seqs; ++b) {
        UMV_REQUIRE(q_lens[b] > 0, UMV_ERR_INVALID, "umv_llm_forward: empty query for sample %d", b);
        M += q_lens[b];
    }
    UMV_REQUIRE(M <= e->d.max_tokens, UMV_ERR_NOMEM, "umv_llm_forward: %d tokens > max_tokens %d", M, e->d.max_tokens);
    r.M = M;
    r.seg = !probe && e->d.heads * e->dh <= e->w_act && gen_rows_segregate(n_seqs, q_lens, row_is_gen, is_causal, &r.Mg);
    UMV_REQUIRE(!presegregated || r.seg, UMV_ERR_STATE, "llm_run: rows were laid out segregated but the forward is not");
    // causal_tail[b] > 0: the last causal_tail[b] rows of sample b are a causal continuation (a prompt) behind a block that runs under
    // the call's mask (an image block, full mask) -- two prefills of the reference (forward_cache_update_vit, then _text) in ONE pass
    // over the weights.  The block's rows do not see the tail; every tail row sees the block and the tail rows before it.
    bool tails = false;
    if (causal_tail)
        for (int b = 0; b < n_seqs; ++b) {
            UMV_REQUIRE(causal_tail[b] >= 0 && causal_tail[b] <= q_lens[b], UMV_ERR_INVALID, "llm_run: causal tail of sample %d out of range", b);
            tails = tails || causal_tail[b] > 0;
        }
    UMV_REQUIRE(!tails || (!r.gen && !probe && !is_causal), UMV_ERR_UNSUPPORTED, "llm_run: causal tails exist for understanding-mode full-mask prefills");
    // reserve pages, gather geometry
    int max_pages = 0;
    std::vector<Seq*> sq(n_seqs);
    for (int b = 0; b < n_seqs; ++b) {
        sq[b] = get_seq(e, seqs[b]);
        if (!sq[b]) return UMV_ERR_INVALID;
        for (int c = 0; c < b; ++c) UMV_REQUIRE(seqs[c] != seqs[b], UMV_ERR_INVALID, "sequence %d appears twice in one call", seqs[b]);
        UMV_TRY(seq_reserve(e, sq[b], sq[b]->len + q_lens[b], st));
        max_pages = std::max(max_pages, (int)sq[b]->pages.size());
        r.max_kv_len = std::max(r.max_kv_len, sq[b]->len + q_lens[b]);
    }
    // Segregated rows: the LINEARS and norms see [generation rows of every sample | marker rows]; ATTENTION keeps the packed order
    // (one contiguous query range per sample, so the 128-row query tiles stay as full as the reference's packing makes them): the
    // q/k-norm + RoPE kernel writes the rotated queries to their packed rows of a separate buffer, the attention kernels store each
    // row's output to its segregated row.  The cache keeps the packed order either way.
    std::vector<int> order, inverse;             // segregated: new row -> packed row, packed row -> new row
    if (r.seg) {
        order.reserve(M);
        for (int pass = 1; pass >= 0; --pass)    // generation rows first
            for (int i = 0; i < M; ++i)
                if ((row_is_gen[i] != 0) == (pass == 1)) order.push_back(i);
        inverse.resize(M);
        for (int i = 0; i < M; ++i) inverse[order[i]] = i;
    }
    r.n_seqs = n_seqs;
    for (int b = 0; b < n_seqs; ++b) r.max_q_len = std::max(r.max_q_len, q_lens[b]);
    MetaBuilder mb;
    UMV_TRY(meta_begin(e, &mb, (size_t)(n_seqs * 48 + 256) + (size_t)M * 32 + (size_t)n_seqs * max_pages * (tails ? 12 : 4)));
    CallMeta& m = r.m;
    m.max_pages = max_pages;
    int* hpos = mb.put<int>(r.seg ? nullptr : positions, M, &m.positions);
    int* hrs = mb.put<int>(nullptr, M, &m.row_seq);
    int* hrp = mb.put<int>(nullptr, M, &m.row_kvpos);
    int* hpt = mb.put<int>(nullptr, (size_t)n_seqs * max_pages, &m.rope_page_table);
    if (r.seg) {
        mb.put<int>(order.data(), M, &m.seg_to_packed);
        mb.put<int>(inverse.data(), M, &m.packed_to_seg);
    }
    // attention groups: (sample, first row, rows, keys).  One per sample -- or, with causal tails, the block groups and the tail groups
    struct Group { int b, row0, n, kv; };
    std::vector<Group> main_groups, tail_groups;
    int row = 0;
    for (int b = 0; b < n_seqs; ++b) {
        const int tail = tails ? causal_tail[b] : 0, head = q_lens[b] - tail;
        if (head > 0 || !tails) main_groups.push_back({b, row, head, sq[b]->len + head});
        if (tail > 0) tail_groups.push_back({b, row + head, tail, sq[b]->len + q_lens[b]});
        for (int j = 0; j < q_lens[b]; ++j, ++row) {
            const int at = r.seg ? inverse[row] : row;     // where the per-row metadata of packed row `row` lives
            hrs[at] = b;
            hrp[at] = sq[b]->len + j;
            if (r.seg) hpos[at] = positions[row];
        }
        for (int p = 0; p < max_pages; ++p) hpt[(size_t)b * max_pages + p] = p < (int)sq[b]->pages.size() ? sq[b]->pages[p] : 0;
    }
    auto put_groups = [&](const std::vector<Group>& gs, int** q_start, int** q_len, int** kv_len, int** page_table, int* max_q) {
        const int ng = (int)gs.size();
        int* hq = mb.put<int>(nullptr, ng + 1, q_start);
        int* hql = mb.put<int>(nullptr, ng, q_len);
        int* hkl = mb.put<int>(nullptr, ng, kv_len);
        int* pt = mb.put<int>(nullptr, (size_t)ng * max_pages, page_table);
        *max_q = 0;
        for (int gi = 0; gi < ng; ++gi) {
            hq[gi] = gs[gi].row0;
            hql[gi] = gs[gi].n;
            hkl[gi] = gs[gi].kv;
            *max_q = std::max(*max_q, gs[gi].n);
            memcpy(pt + (size_t)gi * max_pages, hpt + (size_t)gs[gi].b * max_pages, (size_t)max_pages * sizeof(int));
        }
        hq[ng] = ng ? gs[ng - 1].row0 + gs[ng - 1].n : 0;
    };
    put_groups(main_groups, &m.q_start, &m.q_len, &m.kv_len, &m.page_table, &r.max_q_len);
    r.n_seqs = (int)main_groups.size();
    if (tails) {
        put_groups(tail_groups, &m.tail_q_start, &m.tail_q_len, &m.tail_kv_len, &m.tail_page_table, &m.tail_max_q);
        m.n_tail = (int)tail_groups.size();
    }
    if (r.gen && r.seg) {
        std::vector<uint8_t> sel(M, 0);
        std::fill(sel.begin(), sel.begin() + r.Mg, 1);
        mb.put<uint8_t>(sel.data(), M, &m.row_sel);
        m.n_text = M - r.Mg;
    } else if (r.gen) {
        std::vector<int> text;
        for (int i = 0; i < M; ++i)
            if (!row_is_gen[i]) text.push_back(i);
        UMV_REQUIRE((int)text.size() <= std::min(e->d.max_tokens, 8 * e->d.max_seqs), UMV_ERR_NOMEM,
                    "umv_llm_forward: %zu understanding-expert rows in a gen-mode call exceed the staging size", text.size());
        mb.put<uint8_t>(row_is_gen, M, &m.row_sel);
        mb.put<int>(text.data(), text.size(), &m.text_rows);
        m.n_text = (int)text.size();
        std::vector<int> slot(M, -1);
        for (size_t i = 0; i < text.size(); ++i) slot[text[i]] = (int)i;
        mb.put<int>(slot.data(), M, &m.text_slot);
    }
    UMV_TRY(meta_commit(e, &mb, st));
    if (r.seg && !presegregated) {               // packed rows -> segregated rows, once per forward
        if (x) {
            UMV_TRY(copy_rows(x, D, m.seg_to_packed, e->h, D, M, D, 0, st));
        } else {
            UMV_TRY(copy_rows(e->h, D, m.seg_to_packed, e->act, D, M, D, 0, st));
            UMV_CUDA_OK(cudaMemcpyAsync(e->h, e->act, (size_t)M * D * 2, cudaMemcpyDeviceToDevice, st));
        }
    } else if (x) {
        UMV_CUDA_OK(cudaMemcpy2DAsync(e->h, (size_t)D * 2, x, (size_t)D * 2, (size_t)D * 2, M, cudaMemcpyDeviceToDevice, st));
    }
    r.weight_major = M <= 64;
    UMV_TRY(rope_table(m.positions, e->inv_freq, M, e->dh, e->rope_tab, st));
    r.m.rope_cs = e->rope_tab;
    if (r.seg && !presegregated && out) {        // final norm into a dead workspace, then back to the caller's packed order
        UMV_TRY(llm_layers(e, r, e->act, st));
        UMV_TRY(copy_rows(e->act, D, m.seg_to_packed, out, D, M, D, 1, st));
    } else {
        UMV_TRY(llm_layers(e, r, out, st));
    }
    if (update_kv)
        for (int b = 0; b < n_seqs; ++b) {
            sq[b]->len += q_lens[b];
            sq[b]->content = ++e->content_counter;
        }
    return UMV_OK;
}
}  // namespace umv

extern "C" {

int umv_lm_head(umv_engine* e, const void* hidden, int32_t m, void* logits, void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    return lin(e, static_cast<const bf16*>(hidden), e->d.hidden, e->lm_head, nullptr, nullptr, static_cast<bf16*>(logits),
               e->d.vocab, m, e->d.vocab, e->d.hidden, EPI_BF16, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// SiglipVisionTransformer.forward (siglip_navit.py:345-371): leaves the post-layernorm rows (bf16: the value the consumer Linear
// sees after its autocast cast) in e->xn [M, vit_hidden]; *M_out = number of patch tokens.
static int vit_tower(umv_engine* e, const float* pixels, const int64_t* pos_ids, const int32_t* seqlens, int32_t n_images,
                     cudaStream_t st, int* M_out) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(e->d.enable_vit, UMV_ERR_STATE, "ViT weights were not enabled");
    UMV_REQUIRE(pixels && pos_ids && seqlens && n_images > 0, UMV_ERR_INVALID, "vit: null/empty argument");
    const umv_dims& d = e->d;
    const int Dv = d.vit_hidden, Iv = d.vit_inter, Kp = e->vit_kpad;
    int M = 0, max_len = 0;
    for (int i = 0; i < n_images; ++i) {
        UMV_REQUIRE(seqlens[i] > 0, UMV_ERR_INVALID, "umv_vit_embed: empty image %d", i);
        M += seqlens[i];
        max_len = std::max(max_len, seqlens[i]);
    }
    UMV_REQUIRE(M <= d.max_tokens, UMV_ERR_NOMEM, "umv_vit_embed: %d patches > max_tokens %d", M, d.max_tokens);
    MetaBuilder mb;
    UMV_TRY(meta_begin(e, &mb, (size_t)n_images * 16 + 64));
    int *dq, *dl;
    int* hq = mb.put<int>(nullptr, n_images + 1, &dq);
    mb.put<int>(seqlens, n_images, &dl);
    hq[0] = 0;
    for (int i = 0; i < n_images; ++i) hq[i + 1] = hq[i] + seqlens[i];
    UMV_TRY(meta_commit(e, &mb, st));

    bf16* hv = e->h;        // residual stream [M, Dv]
    bf16* xb = e->act;      // bf16 pixels [M, Kp]
    UMV_TRY(f32_to_bf16_padded(pixels, xb, M, d.vit_patch_dim, Kp, st));
    UMV_TRY(lin(e, xb, Kp, e->vit_patch_w, e->vit_patch_b, nullptr, hv, Dv, M, Dv, Kp, EPI_BF16, st));
    UMV_TRY(gather_add_rows(hv, e->vit_pos, pos_ids, M, Dv, st));
    const int dhv = Dv / d.vit_heads, ldp = 3 * d.vit_heads * 128;
    const char* tc_env = getenv("UMV_ATTN_TC");
    const bool tc_vit = !(tc_env && atoi(tc_env) == 0) && e->vit_qkvp && e->gemm_impl == 0 && M > 64 && max_len >= 128 && Dv % 8 == 0;
    for (int li = 0; li < d.vit_layers; ++li) {
        const VitLayerW& L = e->vit[li];
        UMV_TRY(layernorm_bf16(hv, L.ln1w, L.ln1b, e->xn, M, Dv, d.vit_eps, st));
        AttnArgs aa;
        aa.out = e->attn; aa.ldo = Dv;
        aa.q_start = dq; aa.k_start = dq; aa.q_len = dl; aa.kv_len = dl;
        aa.n = n_images; aa.H = d.vit_heads; aa.Hkv = d.vit_heads; aa.causal = 0;
        aa.max_q_len = max_len; aa.max_kv_len = max_len; aa.total_q = M;
        if (tc_vit) {
            // q|k|v written with every 72-wide head padded to 128 zero-filled columns: the tcgen05 attention kernel then runs
            // as for head_dim 128 (its softmax pipe, not the contraction length, sets the pace) with the real softmax scale
            LinearCall c;
            c.x = e->xn; c.ldx = Dv; c.w = L.wqkv; c.bias = L.bqkv; c.y = e->vit_qkvp; c.ldy = ldp; c.M = M; c.N = 3 * Dv; c.K = Dv;
            c.epi = EPI_BF16; c.impl = GEMM_TOKEN_MAJOR; c.out_head_dim = dhv; c.out_head_pad = 128;
            UMV_TRY(linear_forward(c, st));
            aa.q = e->vit_qkvp; aa.k = e->vit_qkvp + d.vit_heads * 128; aa.v = e->vit_qkvp + 2 * d.vit_heads * 128;
            aa.ldq = aa.ldk = aa.ldv = ldp;
            aa.dh = 128; aa.out_dh = dhv; aa.scale = 1.0f / sqrtf((float)dhv); aa.total_k = M;
        } else {
            UMV_TRY(lin(e, e->xn, Dv, L.wqkv, L.bqkv, nullptr, e->qkv, 3 * Dv, M, 3 * Dv, Dv, EPI_BF16, st));
            aa.q = e->qkv; aa.ldq = 3 * Dv; aa.k = e->qkv + Dv; aa.v = e->qkv + 2 * Dv; aa.ldk = aa.ldv = 3 * Dv;
            aa.dh = dhv;
        }
        UMV_TRY(attention_forward(aa, st));
        UMV_TRY(lin(e, e->attn, Dv, L.wo, L.bo, hv, hv, Dv, M, Dv, Dv, EPI_RESID, st));
        UMV_TRY(layernorm_bf16(hv, L.ln2w, L.ln2b, e->xn, M, Dv, d.vit_eps, st));
        UMV_TRY(lin(e, e->xn, Dv, L.w1, L.b1, nullptr, e->act, Iv, M, Iv, Dv, EPI_GELU, st));
        UMV_TRY(lin(e, e->act, Iv, L.w2, L.b2, hv, hv, Dv, M, Dv, Iv, EPI_RESID, st));
    }
    *M_out = M;
    return layernorm_bf16(hv, e->vit_post_w, e->vit_post_b, e->xn, M, Dv, d.vit_eps, st);
}
// MLPconnector.forward (modeling_utils.py:119-123): x [M, vit_hidden] -> out [M, hidden]; e->act is the fc1 scratch.
static int connector_rows(umv_engine* e, const bf16* x, int M, bf16* out, cudaStream_t st) {
    const int Dv = e->d.vit_hidden, D = e->d.hidden;
    UMV_TRY(lin(e, x, Dv, e->conn_w1, e->conn_b1, nullptr, e->act, D, M, D, Dv, EPI_GELU, st));
    return lin(e, e->act, D, e->conn_w2, e->conn_b2, nullptr, out, D, M, D, D, EPI_BF16, st);
}
// host index array -> checked against [0, M) and that text rows + block rows tile the packed sequence exactly once
static int check_row_cover(const int32_t* a, int na, const int32_t* b, int nb, int M, const char* what) {
    std::vector<uint8_t> seen((size_t)M, 0);
    for (int pass = 0; pass < 2; ++pass) {
        const int32_t* r = pass ? b : a;
        const int n = pass ? nb : na;
        for (int i = 0; i < n; ++i) {
            UMV_REQUIRE(r[i] >= 0 && r[i] < M && !seen[r[i]], UMV_ERR_INVALID, "%s: packed row index %d out of range or repeated", what, r[i]);
            seen[r[i]] = 1;
        }
    }
    UMV_REQUIRE(na + nb == M, UMV_ERR_INVALID, "%s: %d text + %d block rows do not tile the %d packed rows", what, na, nb, M);
    return UMV_OK;
}

extern "C" {

int umv_vit_embed(umv_engine* e, const float* pixels, const int64_t* pos_ids, const int32_t* seqlens, int32_t n_images,
                  void* out, void* stream) {
    UMV_REQUIRE(out, UMV_ERR_INVALID, "umv_vit_embed: null out");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int M = 0;
    UMV_TRY(vit_tower(e, pixels, pos_ids, seqlens, n_images, st, &M));
    UMV_TRY(connector_rows(e, e->xn, M, static_cast<bf16*>(out), st));      // + vit_pos_embed (bagel.py:590-594)
    return gather_add_rows(static_cast<bf16*>(out), e->vit_pos_embed, pos_ids, M, e->d.hidden, st);
}

// ---- the inner module boundary: each sub-module of the reference's Bagel as its own call
int umv_vit_model(umv_engine* e, const float* pixels, const int64_t* pos_ids, const int32_t* seqlens, int32_t n_images, void* out,
                  void* stream) {
    UMV_REQUIRE(out, UMV_ERR_INVALID, "umv_vit_model: null out");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int M = 0;
    UMV_TRY(vit_tower(e, pixels, pos_ids, seqlens, n_images, st, &M));
    UMV_CUDA_OK(cudaMemcpyAsync(out, e->xn, (size_t)M * e->d.vit_hidden * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
    return UMV_OK;
}
int umv_connector(umv_engine* e, const void* x, int32_t n, void* out, void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(e->d.enable_vit, UMV_ERR_STATE, "ViT weights were not enabled");
    UMV_REQUIRE(x && out && n > 0 && n <= e->d.max_tokens, UMV_ERR_INVALID, "umv_connector: bad argument");
    return connector_rows(e, static_cast<const bf16*>(x), n, static_cast<bf16*>(out), static_cast<cudaStream_t>(stream));
}
int umv_pos_embed(umv_engine* e, int32_t which, const int64_t* pos_ids, int32_t n, void* out, void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(pos_ids && out && n > 0 && (which == 0 || which == 1), UMV_ERR_INVALID, "umv_pos_embed: bad argument");
    const bf16* table = which == 0 ? e->vit_pos_embed : e->latent_pos;
    UMV_REQUIRE(table, UMV_ERR_STATE, "umv_pos_embed: that sub-model was not enabled");
    return embed_rows(table, pos_ids, n, e->d.hidden, which == 0 ? e->d.vit_pos_table : e->d.latent_pos_table, static_cast<bf16*>(out),
                      static_cast<cudaStream_t>(stream));
}
int umv_vae2llm(umv_engine* e, const float* x, int32_t n, void* out, void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(e->d.enable_gen, UMV_ERR_STATE, "generation expert weights were not enabled");
    UMV_REQUIRE(x && out && n > 0 && n <= e->d.max_tokens, UMV_ERR_INVALID, "umv_vae2llm: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int C = e->d.latent_dim;
    UMV_TRY(f32_to_bf16_padded(x, e->attn, n, C, C, st));          // the autocast Linear casts its fp32 input to bf16
    return lin(e, e->attn, C, e->vae2llm_w, e->vae2llm_b, nullptr, static_cast<bf16*>(out), e->d.hidden, n, e->d.hidden, C, EPI_BF16, st);
}
int umv_llm2vae(umv_engine* e, const void* h, int32_t n, void* out, void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(e->d.enable_gen, UMV_ERR_STATE, "generation expert weights were not enabled");
    UMV_REQUIRE(h && out && n > 0, UMV_ERR_INVALID, "umv_llm2vae: bad argument");
    const int C = e->d.latent_dim, D = e->d.hidden;
    return lin(e, static_cast<const bf16*>(h), D, e->llm2vae_w, e->llm2vae_b, nullptr, static_cast<bf16*>(out), C, n, C, D, EPI_BF16,
               static_cast<cudaStream_t>(stream));
}

int umv_generate_text(umv_engine* e, int32_t n_seqs, const int32_t* seqs, const int64_t* start_tokens,
                      const int32_t* positions, int32_t n_steps, float temperature, uint64_t seed,
                      const int64_t* forced_tokens, int64_t* tokens_out, void* logits_out, int64_t* next_tokens_out,
                      void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(seqs && start_tokens && positions && tokens_out && n_seqs > 0 && n_steps > 0, UMV_ERR_INVALID,
                "umv_generate_text: null/empty argument");
    UMV_REQUIRE(n_seqs <= e->d.max_seqs && n_seqs <= 64, UMV_ERR_INVALID, "umv_generate_text: %d sequences > max_seqs", n_seqs);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const umv_dims& d = e->d;
    const int B = n_seqs, D = d.hidden, V = d.vocab;
    std::vector<Seq*> sq(B);
    int max_pages = 0;
    LlmRun r;
    r.M = B; r.n_seqs = B; r.max_q_len = 1; r.causal = true; r.gen = false; r.weight_major = true;
    for (int b = 0; b < B; ++b) {
        sq[b] = get_seq(e, seqs[b]);
        if (!sq[b]) return UMV_ERR_INVALID;
        for (int c = 0; c < b; ++c) UMV_REQUIRE(seqs[c] != seqs[b], UMV_ERR_INVALID, "sequence %d appears twice", seqs[b]);
        UMV_TRY(seq_reserve(e, sq[b], sq[b]->len + n_steps, st));
        max_pages = std::max(max_pages, (int)sq[b]->pages.size());
        r.max_kv_len = std::max(r.max_kv_len, sq[b]->len + n_steps);
    }
    UMV_REQUIRE(B * max_pages <= e->dec_pages_cap, UMV_ERR_NOMEM, "decode page table too large");
    // upload the loop state
    std::vector<int> hpos(B), hkl(B), hkp(B), hrs(B), hqs(B + 1), hql(B), hpt((size_t)B * max_pages, 0);
    for (int b = 0; b < B; ++b) {
        hpos[b] = positions[b]; hkl[b] = sq[b]->len + 1; hkp[b] = sq[b]->len; hrs[b] = b; hqs[b] = b; hql[b] = 1;
        for (size_t p = 0; p < sq[b]->pages.size(); ++p) hpt[(size_t)b * max_pages + p] = sq[b]->pages[p];
    }
    hqs[B] = B;
    const int zero = 0;
    UMV_CUDA_OK(cudaMemcpyAsync(e->dec_tokens, start_tokens, B * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    UMV_CUDA_OK(cudaMemcpyAsync(e->dec_pos, hpos.data(), B * 4, cudaMemcpyHostToDevice, st));
    UMV_CUDA_OK(cudaMemcpyAsync(e->dec_kvlen, hkl.data(), B * 4, cudaMemcpyHostToDevice, st));
    UMV_CUDA_OK(cudaMemcpyAsync(e->dec_kvpos, hkp.data(), B * 4, cudaMemcpyHostToDevice, st));
    UMV_CUDA_OK(cudaMemcpyAsync(e->dec_rowseq, hrs.data(), B * 4, cudaMemcpyHostToDevice, st));
    UMV_CUDA_OK(cudaMemcpyAsync(e->dec_qstart, hqs.data(), (B + 1) * 4, cudaMemcpyHostToDevice, st));
    UMV_CUDA_OK(cudaMemcpyAsync(e->dec_qlen, hql.data(), B * 4, cudaMemcpyHostToDevice, st));
    UMV_CUDA_OK(cudaMemcpyAsync(e->dec_pages, hpt.data(), hpt.size() * 4, cudaMemcpyHostToDevice, st));
    UMV_CUDA_OK(cudaMemcpyAsync(e->dec_step, &zero, 4, cudaMemcpyHostToDevice, st));
    UMV_CUDA_OK(cudaStreamSynchronize(st));     // host vectors above go out of scope; also a clean capture start
    r.m.q_start = e->dec_qstart; r.m.q_len = e->dec_qlen; r.m.kv_len = e->dec_kvlen; r.m.positions = e->dec_pos;
    r.m.row_seq = e->dec_rowseq; r.m.row_kvpos = e->dec_kvpos; r.m.page_table = e->dec_pages; r.m.max_pages = max_pages;
    r.m.rope_page_table = e->dec_pages;
    DecodeState ds{e->dec_tokens, e->dec_pos, e->dec_kvlen, e->dec_kvpos, e->dec_step, e->dec_rope, e->inv_freq, e->dh};
    r.m.rope_cs = e->dec_rope;

    auto step = [&](cudaStream_t s, bf16* logits) -> int {
        UMV_TRY(decode_begin_step(e->embed, D, V, ds, forced_tokens, tokens_out, B, e->h, s));
        UMV_TRY(llm_layers(e, r, e->xn, s));
        UMV_TRY(lin(e, e->xn, D, e->lm_head, nullptr, nullptr, logits, V, B, V, D, EPI_BF16, s));
        if (temperature > 0.f) UMV_TRY(sample_rows(logits, B, V, temperature, seed, e->dec_step, e->dec_tokens, s));
        else return argmax_rows(logits, B, V, e->dec_tokens, s, &ds);        // greedy: the argmax launch also ends the step
        return decode_end_step(ds, B, s);
    };

    const bool graph_ok = e->use_graph && logits_out == nullptr && st != nullptr && n_steps > 1;
    // A captured step depends on the batch size, the page-table stride, the 64-key block count the attention split was sized
    // for, the sampling parameters and the per-call kernel switches -- not on the sequences themselves (lengths, pages and
    // tokens live in device arrays the graph only points to).  Same key -> replay the instantiated graph of an earlier call.
    const bool cache_ok = graph_ok && forced_tokens == nullptr && !trace_active() && !(getenv("UMV_GRAPH_CACHE") && atoi(getenv("UMV_GRAPH_CACHE")) == 0);
    if (cache_ok) {
        auto flag = [](const char* name, int bit) { const char* v = getenv(name); return (v && atoi(v) == 0) ? (1u << bit) : 0u; };
        const unsigned flags = flag("UMV_FUSED_ATTN", 0) | flag("UMV_ATTN_TC", 1) | flag("UMV_ROPE_ROWS", 2) | flag("UMV_NORM_WARP", 3);
        const int blocks = (r.max_kv_len + kPageTokens - 1) / kPageTokens;
        const size_t need = (size_t)n_steps * B;
        if (need > e->dec_out_cap) {                   // grown rarely; graphs captured against the old buffer are dropped
            for (auto& g : e->dec_graphs) cudaGraphExecDestroy(g.exec);
            e->dec_graphs.clear();
            if (e->dec_out) cudaFree(e->dec_out);
            e->dec_out = nullptr; e->dec_out_cap = 0;
            UMV_CUDA_OK(cudaMalloc(&e->dec_out, std::max(need, (size_t)4096) * sizeof(int64_t)));
            e->dec_out_cap = std::max(need, (size_t)4096);
        }
        umv_engine::DecodeGraph* hit = nullptr;
        for (auto& g : e->dec_graphs)
            if (g.B == B && g.max_pages == max_pages && g.blocks == blocks && g.temperature == temperature &&
                (temperature <= 0.f || g.seed == seed) && g.flags == flags) hit = &g;
        if (!hit) {
            cudaGraph_t graph = nullptr;
            umv_engine::DecodeGraph g;
            g.B = B; g.max_pages = max_pages; g.blocks = blocks; g.temperature = temperature; g.seed = seed; g.flags = flags;
            UMV_CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            const long long launches_before = g_launches;
            auto cstep = [&]() -> int {
                UMV_TRY(decode_begin_step(e->embed, D, V, ds, nullptr, e->dec_out, B, e->h, st));
                UMV_TRY(llm_layers(e, r, e->xn, st));
                UMV_TRY(lin(e, e->xn, D, e->lm_head, nullptr, nullptr, e->logits, V, B, V, D, EPI_BF16, st));
                if (temperature > 0.f) UMV_TRY(sample_rows(e->logits, B, V, temperature, seed, e->dec_step, e->dec_tokens, st));
                else return argmax_rows(e->logits, B, V, e->dec_tokens, st, &ds);
                return decode_end_step(ds, B, st);
            };
            int rc = cstep();
            g.launches = g_launches - launches_before;
            g_launches = launches_before;
            cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc != UMV_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            UMV_CUDA_OK(ce);
            ce = cudaGraphInstantiate(&g.exec, graph, 0);
            cudaGraphDestroy(graph);
            UMV_CUDA_OK(ce);
            if (e->dec_graphs.size() >= 8) {           // evict the least recently used
                size_t lru = 0;
                for (size_t i = 1; i < e->dec_graphs.size(); ++i) if (e->dec_graphs[i].used < e->dec_graphs[lru].used) lru = i;
                cudaGraphExecDestroy(e->dec_graphs[lru].exec);
                e->dec_graphs.erase(e->dec_graphs.begin() + lru);
            }
            e->dec_graphs.push_back(g);
            hit = &e->dec_graphs.back();
        }
        hit->used = ++e->dec_graph_clock;
        for (int i = 0; i < n_steps; ++i) {
            cudaError_t le = cudaGraphLaunch(hit->exec, st);
            if (le != cudaSuccess) {
                set_error("cudaGraphLaunch failed at step %d: %s", i, cudaGetErrorString(le));
                return UMV_ERR_CUDA;
            }
        }
        g_launches += hit->launches * n_steps;
        UMV_CUDA_OK(cudaMemcpyAsync(tokens_out, e->dec_out, need * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    } else if (graph_ok) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        UMV_CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const long long launches_before = g_launches;
        int rc = step(st, e->logits);
        g_launches += (g_launches - launches_before) * (n_steps - 1);    // every replay launches the captured kernels
        cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc != UMV_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        UMV_CUDA_OK(ce);
        UMV_CUDA_OK(cudaGraphInstantiate(&exec, graph, 0));
        for (int i = 0; i < n_steps; ++i) {
            cudaError_t le = cudaGraphLaunch(exec, st);
            if (le != cudaSuccess) {
                set_error("cudaGraphLaunch failed at step %d: %s", i, cudaGetErrorString(le));
                cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
                return UMV_ERR_CUDA;
            }
        }
        cudaGraphExecDestroy(exec);
        cudaGraphDestroy(graph);
    } else {
        for (int i = 0; i < n_steps; ++i) {
            bf16* lg = logits_out ? static_cast<bf16*>(logits_out) + (size_t)i * B * V : e->logits;
            UMV_TRY(step(st, lg));
        }
    }
    if (next_tokens_out)
        UMV_CUDA_OK(cudaMemcpyAsync(next_tokens_out, e->dec_tokens, B * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    for (int b = 0; b < B; ++b) {
        sq[b]->len += n_steps;
        sq[b]->content = ++e->content_counter;
    }
    return UMV_OK;
}


// ------------------------------------------------------------------------------ rectified flow
int umv_flow_branches_last(umv_engine* e, int32_t* n) {
    UMV_REQUIRE(e && n, UMV_ERR_INVALID, "null argument");
    *n = e->last_flow_branches;
    return UMV_OK;
}
int umv_flow_velocity(umv_engine* e, const umv_flow_args* a, const float* x_t, float* v_out, void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(e->d.enable_gen, UMV_ERR_STATE, "generation expert weights were not enabled");
    UMV_REQUIRE(a && x_t && v_out && a->seqs && a->lat_lens && a->positions && a->marker_ids && a->lat_pos_ids && a->n_seqs > 0,
                UMV_ERR_INVALID, "umv_flow_velocity: null/empty argument");
    UMV_REQUIRE(a->renorm_type >= 0 && a->renorm_type <= 2, UMV_ERR_UNSUPPORTED, "cfg_renorm_type %d is not supported", a->renorm_type);
    const bool has_text = a->cfg_text_scale > 1.0f, has_img = a->cfg_img_scale > 1.0f;
    UMV_REQUIRE(!has_text || (a->cfg_text_seqs && a->cfg_text_positions), UMV_ERR_INVALID, "cfg_text context missing");
    UMV_REQUIRE(!has_img || (a->cfg_img_seqs && a->cfg_img_positions), UMV_ERR_INVALID, "cfg_img context missing");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const umv_dims& d = e->d;
    const int D = d.hidden, C = d.latent_dim, B = a->n_seqs;
    int n_lat = 0;
    for (int b = 0; b < B; ++b) {
        UMV_REQUIRE(a->lat_lens[b] > 0, UMV_ERR_INVALID, "empty image %d", b);
        n_lat += a->lat_lens[b];
    }
    const int Mb = n_lat + 2 * B;
    // Two branches over the SAME context are one forward: in a pure text-to-image request the image-free context (cfg_img) holds
    // exactly the main context's tokens (inferencer.py:578-600 feeds both the same text), so its velocity IS the main branch's, bit for
    // bit (same rows, same keys, batch-invariant kernels) -- evaluate it once and let the CFG mix read it twice.  Detected from the
    // sequences' content ids (a fork of the main context, or the main context itself), same lengths and rope positions; the reference
    // runs the third forward anyway.  UMV_CFG_DEDUP=0 switches it off (read per call).
    bool img_is_main = false;
    if (has_img && has_text) {
        const char* dd = getenv("UMV_CFG_DEDUP");
        img_is_main = !(dd && atoi(dd) == 0);
        for (int b = 0; b < B && img_is_main; ++b) {
            const Seq* sm = get_seq(e, a->seqs[b]);
            const Seq* si = get_seq(e, a->cfg_img_seqs[b]);
            if (!sm || !si) return UMV_ERR_INVALID;
            img_is_main = sm->content == si->content && sm->len == si->len && a->positions[b] == a->cfg_img_positions[b];
        }
    }
    // the reference evaluates cfg_img only inside the cfg_text branch (bagel.py:1173-1207): without cfg_text it is not run at all
    const bool run_img = has_img && has_text && !img_is_main;
    const int text_branch = has_text ? 1 : -1, img_branch = (has_img && has_text) ? (img_is_main ? 0 : 2) : -1;
    const int nb = 1 + (has_text ? 1 : 0) + (run_img ? 1 : 0);
    e->last_flow_branches = nb;
    UMV_REQUIRE(nb * Mb <= d.max_tokens, UMV_ERR_NOMEM, "flow step needs %d rows > max_tokens %d", nb * Mb, d.max_tokens);
    // the reference evaluates cfg_img only inside the cfg_text branch (bagel.py:1173-1207): img without text is ignored
    // branch / sample / row geometry of the one packed gen-mode forward over all branches (rows are independent; contexts and rope
    // positions differ per branch)
    std::vector<int32_t> seqs, qlens, pos;
    std::vector<uint8_t> is_gen;
    const int32_t* bseq[3] = {a->seqs, nullptr, nullptr};
    const int32_t* bpos[3] = {a->positions, nullptr, nullptr};
    if (has_text) { bseq[text_branch] = a->cfg_text_seqs; bpos[text_branch] = a->cfg_text_positions; }
    if (run_img) { bseq[img_branch] = a->cfg_img_seqs; bpos[img_branch] = a->cfg_img_positions; }
    for (int br = 0; br < nb; ++br)
        for (int b = 0; b < B; ++b) {
            seqs.push_back(bseq[br][b]);
            qlens.push_back(a->lat_lens[b] + 2);
            for (int j = 0; j < a->lat_lens[b] + 2; ++j) {
                pos.push_back(bpos[br][b]);
                is_gen.push_back(j > 0 && j < a->lat_lens[b] + 1 ? 1 : 0);
            }
        }
    // the forward's segregated row order (llm_run): latent rows of every branch first, then the marker rows; flow_compose writes it directly
    int n_gen_rows = 0;
    const bool seg = gen_rows_segregate(nb * B, qlens.data(), is_gen.data(), 0, &n_gen_rows);
    MetaBuilder mb;
    UMV_TRY(meta_begin(e, &mb, (size_t)Mb * 4 + (size_t)B * 16 + 64 + (seg ? (size_t)nb * Mb * 4 : 0)));
    int *d_row_src, *d_row0, *d_lat0, *d_n;
    int* h_src = mb.put<int>(nullptr, Mb, &d_row_src);
    int* h_row0 = mb.put<int>(nullptr, B, &d_row0);
    int* h_lat0 = mb.put<int>(nullptr, B, &d_lat0);
    mb.put<int>(a->lat_lens, B, &d_n);
    int r = 0, l = 0;
    for (int b = 0; b < B; ++b) {
        h_src[r++] = -1;
        h_row0[b] = r;
        h_lat0[b] = l;
        for (int j = 0; j < a->lat_lens[b]; ++j) h_src[r++] = l++;
        h_src[r++] = -2;
    }
    int* d_row_dst = nullptr;
    if (seg) {
        int* h_dst = mb.put<int>(nullptr, (size_t)nb * Mb, &d_row_dst);
        int ig = 0, it = n_gen_rows;
        for (int i = 0; i < nb * Mb; ++i) h_dst[i] = is_gen[i] ? ig++ : it++;
    }
    UMV_TRY(meta_commit(e, &mb, st));
    // 1. vae2llm(x_t): fp32 latents enter the autocast Linear as bf16
    bf16* xtb = e->attn;
    bf16* lat = e->act;
    UMV_TRY(f32_to_bf16_padded(x_t, xtb, n_lat, C, C, st));
    UMV_TRY(lin(e, xtb, C, e->vae2llm_w, e->vae2llm_b, nullptr, lat, D, n_lat, D, C, EPI_BF16, st));
    // 2. timestep embedding, once per step (the reference recomputes it per token for an identical t)
    bf16* tf = e->flow_small;
    bf16* th = e->flow_small + std::max(D, 256);
    bf16* temb = e->flow_small + 2 * std::max(D, 256);
    UMV_TRY(timestep_freq(a->timestep, e->t_freqs, 128, tf, st));
    UMV_TRY(lin(e, tf, 256, e->t_w0, e->t_b0, nullptr, th, D, 1, D, 256, EPI_BF16, st));
    UMV_TRY(silu_inplace(th, D, st));
    UMV_TRY(lin(e, th, D, e->t_w2, e->t_b2, nullptr, temb, D, 1, D, D, EPI_BF16, st));
    // 3. packed query sequence for every branch
    UMV_REQUIRE(a->marker_ids[0] >= 0 && a->marker_ids[0] < d.vocab && a->marker_ids[1] >= 0 && a->marker_ids[1] < d.vocab,
                UMV_ERR_INVALID, "marker token id out of range");
    UMV_TRY(flow_compose(lat, temb, e->latent_pos, a->lat_pos_ids, e->embed, a->marker_ids[0], a->marker_ids[1], d_row_src, Mb, nb,
                         D, e->h, st, d_row_dst));
    // 4. one packed gen-mode forward over all branches
    UMV_TRY(llm_run(e, nullptr, nb * B, seqs.data(), qlens.data(), pos.data(), is_gen.data(), 0, 0, e->xn, st, nullptr, seg ? 1 : 0));
    // 5. llm2vae (segregated: on the latent rows only; else on every row), then CFG on the latent rows
    bf16* vall = e->qkv;
    const int rows_per_branch = seg ? n_lat : Mb;
    UMV_TRY(lin(e, e->xn, D, e->llm2vae_w, e->llm2vae_b, nullptr, vall, C, nb * rows_per_branch, C, D, EPI_BF16, st));
    CfgArgs c;
    c.v = vall; c.rows_per_branch = rows_per_branch; c.C = C; c.text_branch = text_branch; c.img_branch = img_branch;
    c.text_scale = a->cfg_text_scale; c.img_scale = a->cfg_img_scale; c.renorm_min = a->cfg_renorm_min;
    c.renorm_type = a->renorm_type; c.img_row0 = seg ? d_lat0 : d_row0; c.img_lat0 = d_lat0; c.img_n = d_n; c.out = v_out;
    return cfg_combine(c, B, st);
}

}  // extern "C"
// TimestepEmbedder.forward (modeling_utils.py:106-109) for one timestep: fp32 sinusoid -> Linear -> SiLU (bf16) -> Linear -> temb [hidden]
static int time_embed_one(umv_engine* e, float t, bf16* temb, cudaStream_t st) {
    const int D = e->d.hidden;
    bf16* tf = e->flow_small;
    bf16* th = e->flow_small + std::max(D, 256);
    UMV_TRY(timestep_freq(t, e->t_freqs, 128, tf, st));
    UMV_TRY(lin(e, tf, 256, e->t_w0, e->t_b0, nullptr, th, D, 1, D, 256, EPI_BF16, st));
    UMV_TRY(silu_inplace(th, D, st));
    return lin(e, th, D, e->t_w2, e->t_b2, nullptr, temb, D, 1, D, D, EPI_BF16, st);
}
extern "C" {
int umv_time_embedder(umv_engine* e, const float* timesteps, int32_t n, void* out, void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(e->d.enable_gen, UMV_ERR_STATE, "generation expert weights were not enabled");
    UMV_REQUIRE(timesteps && out && n > 0, UMV_ERR_INVALID, "umv_time_embedder: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int D = e->d.hidden;
    bf16* temb = e->flow_small + 2 * std::max(D, 256);
    for (int i = 0; i < n; ++i) {        // a row per DISTINCT value would do; callers pass a handful of timesteps
        if (i > 0 && timesteps[i] == timesteps[i - 1]) {
            UMV_CUDA_OK(cudaMemcpyAsync(static_cast<bf16*>(out) + (size_t)i * D, static_cast<bf16*>(out) + (size_t)(i - 1) * D,
                                        (size_t)D * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
            continue;
        }
        UMV_TRY(time_embed_one(e, timesteps[i], temb, st));
        UMV_CUDA_OK(cudaMemcpyAsync(static_cast<bf16*>(out) + (size_t)i * D, temb, (size_t)D * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
    }
    return UMV_OK;
}

// ---- prefill drivers: one call per reference method, the packed query sequence is composed inside the engine
}  // extern "C"

// Decode riders (umv_decode_riders): running requests whose next token is computed by the SAME forward that prefills other requests'
// blocks.  Their rows (one per rider: the embedding of its current token) are appended after the M prefill rows already composed in
// e->h; after the forward their final hidden states go through lm_head and argmax / sampling.  A single query row sees its whole
// context under either mask, so riders join causal (text) and full-attention (image block) prefills alike.
static int run_with_riders(umv_engine* e, int M, int n_seqs, const int32_t* seqs, const int32_t* q_lens, const int32_t* positions,
                           const uint8_t* row_is_gen, int is_causal, const umv_decode_riders* rd, cudaStream_t st,
                           const int32_t* causal_tail = nullptr) {
    const int n = rd ? rd->n : 0;
    if (n <= 0) return umv::llm_run(e, nullptr, n_seqs, seqs, q_lens, positions, row_is_gen, is_causal, 1, nullptr, st, nullptr, 0, causal_tail);
    UMV_REQUIRE(rd->seqs && rd->tokens && rd->positions && rd->next_tokens && n <= 64, UMV_ERR_INVALID, "decode riders: bad argument (n <= 64)");
    UMV_REQUIRE(M + n <= e->d.max_tokens, UMV_ERR_NOMEM, "prefill rows + decode riders (%d) > max_tokens %d", M + n, e->d.max_tokens);
    const int D = e->d.hidden, V = e->d.vocab;
    for (int i = 0; i < n; ++i) UMV_REQUIRE(rd->tokens[i] >= 0 && rd->tokens[i] < V, UMV_ERR_INVALID, "rider token id out of range");
    MetaBuilder mb;
    UMV_TRY(meta_begin(e, &mb, (size_t)n * 8 + 64));
    int64_t* d_ids;
    mb.put<int64_t>(rd->tokens, n, &d_ids);
    UMV_TRY(meta_commit(e, &mb, st));
    UMV_TRY(embed_rows(e->embed, d_ids, n, D, V, e->h + (size_t)M * D, st));
    std::vector<int32_t> sq(seqs, seqs + n_seqs), ql(q_lens, q_lens + n_seqs), pos(positions, positions + M);
    std::vector<uint8_t> gen;
    if (row_is_gen) gen.assign(row_is_gen, row_is_gen + M);
    for (int i = 0; i < n; ++i) {
        sq.push_back(rd->seqs[i]);
        ql.push_back(1);
        pos.push_back(rd->positions[i]);
        if (row_is_gen) gen.push_back(0);                    // a decode row is a text row: understanding expert
    }
    std::vector<int32_t> ct;
    if (causal_tail) {
        ct.assign(causal_tail, causal_tail + n_seqs);
        ct.resize(n_seqs + n, 0);                            // a rider is one row: no tail
    }
    UMV_TRY(umv::llm_run(e, nullptr, n_seqs + n, sq.data(), ql.data(), pos.data(), row_is_gen ? gen.data() : nullptr, is_causal, 1, nullptr, st,
                         nullptr, 0, causal_tail ? ct.data() : nullptr));
    // final-normed hidden states of all rows are in e->xn; the riders' are the last n
    UMV_TRY(lin(e, e->xn + (size_t)M * D, D, e->lm_head, nullptr, nullptr, e->logits, V, n, V, D, EPI_BF16, st));
    if (rd->temperature > 0.f) return sample_rows(e->logits, n, V, rd->temperature, rd->seed, nullptr, rd->next_tokens, st);
    return argmax_rows(e->logits, n, V, rd->next_tokens, st);
}

extern "C" {

int umv_forward_cache_update_text_riders(umv_engine* e, int32_t n_seqs, const int32_t* seqs, const int32_t* text_lens,
                                         const int64_t* text_ids, const int32_t* positions, const umv_decode_riders* riders,
                                         void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(seqs && text_lens && text_ids && positions && n_seqs > 0, UMV_ERR_INVALID, "umv_forward_cache_update_text: null/empty argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int M = 0;
    for (int b = 0; b < n_seqs; ++b) M += text_lens[b];
    UMV_REQUIRE(M > 0 && M <= e->d.max_tokens, UMV_ERR_NOMEM, "umv_forward_cache_update_text: %d tokens > max_tokens %d", M, e->d.max_tokens);
    for (int i = 0; i < M; ++i)
        UMV_REQUIRE(text_ids[i] >= 0 && text_ids[i] < e->d.vocab, UMV_ERR_INVALID, "token id %lld out of range", (long long)text_ids[i]);
    MetaBuilder mb;
    UMV_TRY(meta_begin(e, &mb, (size_t)M * 8 + 64));
    int64_t* d_ids;
    mb.put<int64_t>(text_ids, M, &d_ids);
    UMV_TRY(meta_commit(e, &mb, st));
    UMV_TRY(embed_rows(e->embed, d_ids, M, e->d.hidden, e->d.vocab, e->h, st));
    return run_with_riders(e, M, n_seqs, seqs, text_lens, positions, nullptr, 1, riders, st);
}
int umv_forward_cache_update_text(umv_engine* e, int32_t n_seqs, const int32_t* seqs, const int32_t* text_lens,
                                  const int64_t* text_ids, const int32_t* positions, void* stream) {
    return umv_forward_cache_update_text_riders(e, n_seqs, seqs, text_lens, text_ids, positions, nullptr, stream);
}

int umv_forward_cache_update_vit(umv_engine* e, int32_t n_seqs, const int32_t* seqs, const int32_t* seq_lens, int32_t n_text,
                                 const int64_t* text_ids, const int32_t* text_rows, const float* pixels, const int64_t* vit_pos_ids,
                                 int32_t n_images, const int32_t* vit_seqlens, const int32_t* vit_rows, const int32_t* positions,
                                 void* stream) {
    return umv_forward_cache_update_vit_riders(e, n_seqs, seqs, seq_lens, n_text, text_ids, text_rows, pixels, vit_pos_ids, n_images,
                                               vit_seqlens, vit_rows, positions, nullptr, stream);
}
int umv_forward_cache_update_vit_riders(umv_engine* e, int32_t n_seqs, const int32_t* seqs, const int32_t* seq_lens, int32_t n_text,
                                        const int64_t* text_ids, const int32_t* text_rows, const float* pixels,
                                        const int64_t* vit_pos_ids, int32_t n_images, const int32_t* vit_seqlens,
                                        const int32_t* vit_rows, const int32_t* positions, const umv_decode_riders* riders,
                                        void* stream) {
    return umv_forward_cache_update_vit_prompt(e, n_seqs, seqs, seq_lens, nullptr, n_text, text_ids, text_rows, pixels, vit_pos_ids, n_images,
                                               vit_seqlens, vit_rows, positions, riders, stream);
}
int umv_forward_cache_update_vit_prompt(umv_engine* e, int32_t n_seqs, const int32_t* seqs, const int32_t* seq_lens,
                                        const int32_t* prompt_lens, int32_t n_text, const int64_t* text_ids, const int32_t* text_rows,
                                        const float* pixels, const int64_t* vit_pos_ids, int32_t n_images, const int32_t* vit_seqlens,
                                        const int32_t* vit_rows, const int32_t* positions, const umv_decode_riders* riders,
                                        void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(seqs && seq_lens && text_ids && text_rows && vit_rows && positions && n_seqs > 0 && n_text >= 0, UMV_ERR_INVALID,
                "umv_forward_cache_update_vit: null/empty argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int D = e->d.hidden;
    int M = 0, N = 0;
    for (int b = 0; b < n_seqs; ++b) M += seq_lens[b];
    for (int i = 0; i < n_images; ++i) N += vit_seqlens[i];
    UMV_REQUIRE(M <= e->d.max_tokens, UMV_ERR_NOMEM, "umv_forward_cache_update_vit: %d tokens > max_tokens %d", M, e->d.max_tokens);
    UMV_TRY(check_row_cover(text_rows, n_text, vit_rows, N, M, "umv_forward_cache_update_vit"));
    for (int i = 0; i < n_text; ++i)
        UMV_REQUIRE(text_ids[i] >= 0 && text_ids[i] < e->d.vocab, UMV_ERR_INVALID, "token id %lld out of range", (long long)text_ids[i]);
    int Mv = 0;
    UMV_TRY(vit_tower(e, pixels, vit_pos_ids, vit_seqlens, n_images, st, &Mv));
    UMV_TRY(connector_rows(e, e->xn, Mv, e->attn, st));            // e->h (the tower's residual stream) is free from here on
    MetaBuilder mb;
    UMV_TRY(meta_begin(e, &mb, (size_t)n_text * 12 + (size_t)N * 4 + 64));
    int64_t* d_ids; int *d_trows, *d_vrows;
    mb.put<int64_t>(text_ids, n_text, &d_ids);
    mb.put<int>(text_rows, n_text, &d_trows);
    mb.put<int>(vit_rows, N, &d_vrows);
    UMV_TRY(meta_commit(e, &mb, st));
    UMV_TRY(scatter_add_rows(e->attn, e->vit_pos_embed, vit_pos_ids, d_vrows, e->h, N, D, st));      // + vit_pos_embed, to its rows
    UMV_TRY(embed_rows_scatter(e->embed, d_ids, d_trows, n_text, D, e->d.vocab, e->h, st));          // marker (and prompt) embeddings
    return run_with_riders(e, M, n_seqs, seqs, seq_lens, positions, nullptr, 0, riders, st, prompt_lens);
}

int umv_forward_cache_update_vae(umv_engine* e, int32_t n_seqs, const int32_t* seqs, const int32_t* seq_lens, int32_t n_text,
                                 const int64_t* text_ids, const int32_t* text_rows, const void* latent, int32_t n_images, int32_t Hl,
                                 int32_t Wl, const int32_t* latent_hw, int32_t patch, const int64_t* lat_pos_ids, const int32_t* lat_rows,
                                 float timestep, const int32_t* positions, void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(e->d.enable_gen, UMV_ERR_STATE, "generation expert weights were not enabled");
    UMV_REQUIRE(seqs && seq_lens && text_ids && text_rows && latent && latent_hw && lat_pos_ids && lat_rows && positions && n_seqs > 0 &&
                n_images > 0 && patch > 0 && n_text == 2 * n_images, UMV_ERR_INVALID, "umv_forward_cache_update_vae: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int D = e->d.hidden, C = e->d.latent_dim, Cz = C / (patch * patch);
    UMV_REQUIRE(Cz * patch * patch == C, UMV_ERR_INVALID, "latent patch %d does not divide latent_dim %d", patch, C);
    int M = 0, N = 0;
    for (int b = 0; b < n_seqs; ++b) M += seq_lens[b];
    for (int i = 0; i < n_images; ++i) {
        UMV_REQUIRE(latent_hw[2 * i] > 0 && latent_hw[2 * i + 1] > 0 && latent_hw[2 * i] * patch <= Hl && latent_hw[2 * i + 1] * patch <= Wl,
                    UMV_ERR_INVALID, "latent shape of image %d exceeds the padded plane", i);
        N += latent_hw[2 * i] * latent_hw[2 * i + 1];
    }
    UMV_REQUIRE(M <= e->d.max_tokens, UMV_ERR_NOMEM, "umv_forward_cache_update_vae: %d tokens > max_tokens %d", M, e->d.max_tokens);
    UMV_TRY(check_row_cover(text_rows, n_text, lat_rows, N, M, "umv_forward_cache_update_vae"));
    for (int i = 0; i < n_text; ++i)
        UMV_REQUIRE(text_ids[i] == text_ids[i % 2] && text_ids[i] >= 0 && text_ids[i] < e->d.vocab, UMV_ERR_UNSUPPORTED,
                    "every image must use the same start / end marker ids");
    // latent patchify "chpwq->hwpqc" (bagel.py:760-765) -> bf16 rows; vae2llm; + time embedding + latent_pos_embed; markers
    bf16* xtb = e->attn;
    bf16* lat = e->act;
    int off = 0;
    for (int i = 0; i < n_images; ++i) {
        const int h = latent_hw[2 * i], w = latent_hw[2 * i + 1];
        UMV_TRY(latent_patchify(static_cast<const bf16*>(latent) + (size_t)i * Cz * Hl * Wl, xtb + (size_t)off * C, Cz, Hl, Wl, h, w, patch, st));
        off += h * w;
    }
    UMV_TRY(lin(e, xtb, C, e->vae2llm_w, e->vae2llm_b, nullptr, lat, D, N, D, C, EPI_BF16, st));
    bf16* temb = e->flow_small + 2 * std::max(D, 256);
    UMV_TRY(time_embed_one(e, timestep, temb, st));
    std::vector<int> src((size_t)M, 0);
    std::vector<uint8_t> is_gen((size_t)M, 0);
    for (int i = 0; i < n_text; ++i) src[text_rows[i]] = (i % 2 == 0) ? -1 : -2;
    for (int i = 0; i < N; ++i) { src[lat_rows[i]] = i; is_gen[lat_rows[i]] = 1; }
    MetaBuilder mb;
    UMV_TRY(meta_begin(e, &mb, (size_t)M * 4 + 64));
    int* d_src;
    mb.put<int>(src.data(), M, &d_src);
    UMV_TRY(meta_commit(e, &mb, st));
    UMV_TRY(flow_compose(lat, temb, e->latent_pos, lat_pos_ids, e->embed, text_ids[0], text_ids[1], d_src, M, 1, D, e->h, st));
    return umv::llm_run(e, nullptr, n_seqs, seqs, seq_lens, positions, is_gen.data(), 0, 1, nullptr, st);
}

int umv_latent_embed(umv_engine* e, const float* x, const int64_t* pos_ids, int32_t n, float timestep, void* out, void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(e->d.enable_gen, UMV_ERR_STATE, "generation expert weights were not enabled");
    UMV_REQUIRE(x && pos_ids && out && n > 0 && n <= e->d.max_tokens, UMV_ERR_INVALID, "umv_latent_embed: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int D = e->d.hidden, C = e->d.latent_dim;
    MetaBuilder mb;
    UMV_TRY(meta_begin(e, &mb, (size_t)n * 4 + 64));
    int* d_src;
    int* h_src = mb.put<int>(nullptr, n, &d_src);
    for (int i = 0; i < n; ++i) h_src[i] = i;
    UMV_TRY(meta_commit(e, &mb, st));
    bf16* xtb = e->attn;
    bf16* lat = e->act;
    UMV_TRY(f32_to_bf16_padded(x, xtb, n, C, C, st));
    UMV_TRY(lin(e, xtb, C, e->vae2llm_w, e->vae2llm_b, nullptr, lat, D, n, D, C, EPI_BF16, st));
    bf16* tf = e->flow_small;
    bf16* th = e->flow_small + std::max(D, 256);
    bf16* temb = e->flow_small + 2 * std::max(D, 256);
    UMV_TRY(timestep_freq(timestep, e->t_freqs, 128, tf, st));
    UMV_TRY(lin(e, tf, 256, e->t_w0, e->t_b0, nullptr, th, D, 1, D, 256, EPI_BF16, st));
    UMV_TRY(silu_inplace(th, D, st));
    UMV_TRY(lin(e, th, D, e->t_w2, e->t_b2, nullptr, temb, D, 1, D, D, EPI_BF16, st));
    return flow_compose(lat, temb, e->latent_pos, pos_ids, e->embed, 0, 0, d_src, n, 1, D, static_cast<bf16*>(out), st);
}

int umv_flow_euler(umv_engine* e, float* x_t, const float* v, int64_t n, float dt, int32_t v_is_bf16, void* stream) {
    UMV_REQUIRE(e && x_t && v, UMV_ERR_INVALID, "umv_flow_euler: null argument");
    return euler_step(x_t, v, n, dt, v_is_bf16, static_cast<cudaStream_t>(stream));
}

int umv_bench_decode_linear(umv_engine* e, int32_t which, int32_t layer, int32_t m, int64_t* weight_bytes, void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(layer >= 0 && layer < e->d.layers && m > 0 && m <= 64 && which >= 0 && which <= 4, UMV_ERR_INVALID,
                "umv_bench_decode_linear: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const umv_dims& d = e->d;
    const int D = d.hidden, I = d.inter, QN = e->qkvn;
    const LayerW& L = e->layers[layer];
    int64_t wb = 0;
    int rc = UMV_OK;
    switch (which) {
        case 0: wb = (int64_t)QN * D * 2;
            rc = lin(e, e->xn, D, L.wqkv[0], nullptr, nullptr, nullptr, 0, m, QN, D, EPI_PARTIAL, st, GEMM_WEIGHT_MAJOR, e->ws,
                     pick_splits(QN, D, e->sm_count));
            break;
        case 1: wb = (int64_t)D * D * 2;
            rc = lin(e, e->attn, D, L.wo[0], nullptr, nullptr, nullptr, 0, m, D, D, EPI_PARTIAL, st, GEMM_WEIGHT_MAJOR, e->ws,
                     pick_splits(D, D, e->sm_count));
            break;
        case 2: wb = (int64_t)2 * I * D * 2;
            rc = lin(e, e->xn, D, L.wgu[0], nullptr, nullptr, e->act, I, m, 2 * I, D, EPI_SWIGLU, st, GEMM_WEIGHT_MAJOR);
            break;
        case 3: wb = (int64_t)D * I * 2;
            rc = lin(e, e->act, I, L.wdown[0], nullptr, nullptr, nullptr, 0, m, D, I, EPI_PARTIAL, st, GEMM_WEIGHT_MAJOR, e->ws,
                     pick_splits(D, I, e->sm_count));
            break;
        default: wb = (int64_t)d.vocab * D * 2;
            rc = lin(e, e->xn, D, e->lm_head, nullptr, nullptr, e->logits, d.vocab, m, d.vocab, D, EPI_BF16, st, GEMM_WEIGHT_MAJOR);
    }
    if (weight_bytes) *weight_bytes = wb;
    return rc;
}

// ------------------------------------------------------------------------------ op-level entry points
int umv_op_linear(const void* x, const void* w, const void* bias, const void* residual, void* y, int32_t M, int32_t N,
                  int32_t K, int32_t epi, int32_t impl, void* stream) {
    UMV_REQUIRE(x && w && y, UMV_ERR_INVALID, "umv_op_linear: null argument");
    UMV_REQUIRE(epi >= 0 && epi <= 3, UMV_ERR_INVALID, "umv_op_linear: epi %d", epi);
    LinearCall c;
    c.x = static_cast<const bf16*>(x); c.ldx = K; c.w = static_cast<const bf16*>(w);
    c.bias = static_cast<const bf16*>(bias); c.residual = static_cast<const bf16*>(residual);
    c.y = static_cast<bf16*>(y); c.ldy = epi == EPI_SWIGLU ? N / 2 : N; c.M = M; c.N = N; c.K = K; c.epi = epi; c.impl = impl;
    return linear_forward(c, static_cast<cudaStream_t>(stream));
}
int umv_op_rmsnorm(const void* x, const void* w, void* y, int32_t M, int32_t D, float eps, void* stream) {
    AddNormArgs a;
    a.h = const_cast<bf16*>(static_cast<const bf16*>(x)); a.w0 = a.w1 = static_cast<const bf16*>(w);
    a.y = static_cast<bf16*>(y); a.M = M; a.D = D; a.eps = eps;
    return add_rmsnorm(a, static_cast<cudaStream_t>(stream));
}
int umv_op_layernorm(const void* x, const void* w, const void* b, void* y, int32_t M, int32_t D, float eps, void* stream) {
    return layernorm_bf16(static_cast<const bf16*>(x), static_cast<const bf16*>(w), static_cast<const bf16*>(b),
                          static_cast<bf16*>(y), M, D, eps, static_cast<cudaStream_t>(stream));
}
int umv_op_attention(const void* q, const void* k, const void* v, void* out, int32_t n, const int32_t* q_lens,
                     const int32_t* k_lens, int32_t heads, int32_t kv_heads, int32_t head_dim, int32_t causal, void* stream) {
    UMV_REQUIRE(q && k && v && out && q_lens && k_lens && n > 0, UMV_ERR_INVALID, "umv_op_attention: null/empty argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::vector<int> meta(4 * n + 2);
    int* qs = meta.data(); int* ks = qs + n + 1; int* ql = ks + n + 1; int* kl = ql + n;
    qs[0] = ks[0] = 0;
    int mq = 0, mk = 0;
    for (int i = 0; i < n; ++i) {
        qs[i + 1] = qs[i] + q_lens[i]; ks[i + 1] = ks[i] + k_lens[i]; ql[i] = q_lens[i]; kl[i] = k_lens[i];
        mq = std::max(mq, q_lens[i]); mk = std::max(mk, k_lens[i]);
    }
    int* dmeta = nullptr;
    UMV_CUDA_OK(cudaMalloc(&dmeta, meta.size() * 4));
    cudaMemcpyAsync(dmeta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice, st);
    AttnArgs a;
    a.q = static_cast<const bf16*>(q); a.ldq = heads * head_dim; a.out = static_cast<bf16*>(out); a.ldo = heads * head_dim;
    a.k = static_cast<const bf16*>(k); a.v = static_cast<const bf16*>(v); a.ldk = a.ldv = kv_heads * head_dim;
    a.q_start = dmeta; a.k_start = dmeta + n + 1; a.q_len = dmeta + 2 * n + 2; a.kv_len = dmeta + 3 * n + 2;
    a.n = n; a.H = heads; a.Hkv = kv_heads; a.dh = head_dim; a.causal = causal; a.max_q_len = mq; a.max_kv_len = mk;
    a.total_q = qs[n];
    int rc = attention_init();
    if (rc == UMV_OK) rc = attention_forward(a, st);
    cudaStreamSynchronize(st);
    cudaFree(dmeta);
    return rc;
}
int umv_op_attention_block(umv_engine* e, int32_t layer, const void* qkv, const float* qkv_partial, int32_t n_partials,
                           const void* bias, int32_t n_seqs, const int32_t* seqs, const int32_t* q_lens, const int32_t* positions,
                           const uint8_t* row_is_gen, int32_t is_causal, int32_t update_kv, void* out, int32_t* path_out,
                           void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(layer >= 0 && layer < e->d.layers && out && (qkv || (qkv_partial && bias && n_partials > 0)), UMV_ERR_INVALID,
                "umv_op_attention_block: bad argument");
    int M = 0;
    for (int b = 0; b < n_seqs; ++b) M += q_lens[b];
    UMV_REQUIRE(!qkv_partial || (size_t)n_partials * M * e->qkvn <= e->ws_elems, UMV_ERR_NOMEM,
                "umv_op_attention_block: partials exceed the split-K workspace");
    AttnProbe p;
    p.layer = layer; p.qkv = static_cast<const bf16*>(qkv); p.partial = qkv_partial; p.splits = qkv_partial ? n_partials : 0;
    p.bias = static_cast<const bf16*>(bias); p.out = static_cast<bf16*>(out);
    int rc = umv::llm_run(e, nullptr, n_seqs, seqs, q_lens, positions, row_is_gen, is_causal, update_kv, nullptr,
                          static_cast<cudaStream_t>(stream), &p);
    if (path_out) *path_out = p.path;
    return rc;
}
int umv_op_argmax(const void* logits, int32_t rows, int32_t vocab, int64_t* out, void* stream) {
    return argmax_rows(static_cast<const bf16*>(logits), rows, vocab, out, static_cast<cudaStream_t>(stream));
}
int umv_op_sample(const void* logits, int32_t rows, int32_t vocab, float temperature, uint64_t seed, float u_force, int64_t* out,
                  void* stream) {
    UMV_REQUIRE(logits && out && rows > 0 && vocab > 0 && temperature > 0.f, UMV_ERR_INVALID, "umv_op_sample: bad argument");
    return sample_rows(static_cast<const bf16*>(logits), rows, vocab, temperature, seed, nullptr, out, static_cast<cudaStream_t>(stream),
                       u_force);
}

}  // extern "C"
