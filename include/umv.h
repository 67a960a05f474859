/*
 * umv.h -- C ABI of the B200-native UniMedVL inference engine (libumv.so).
 *
 * The reference (uni-medical/UniMedVL) has no FFI, plugin registry or native code: its boundary for
 * the unified understanding+generation forward path is the Python method surface of
 * codes/modeling/unimedvl/bagel.py:Bagel and its sub-models.  Each entry point
 * below names the reference interface it stands in for; unimedvl_b200/bagel.py re-creates the
 * Python surface on top of these calls (ctypes binding: unimedvl_b200/_lib.py, maintainers' stub
 * in INTEGRATION.md).
 *
 * Conventions: every function returns 0 (UMV_OK) or a negative umv_status; umv_last_error() gives
 * the message of the calling thread's last failure.  Bulk tensors are DEVICE pointers (bf16 unless
 * stated), small index/shape metadata are HOST pointers.  `stream` is a cudaStream_t passed as
 * void* (NULL = legacy default stream).  One engine per GPU per process; a handle is not
 * thread-safe (the reference is single-threaded, codes/interactive_vqa_inferencer.py:19-20).
 * There is no CPU fallback: without a CUDA device every compute call fails with UMV_ERR_CUDA.
 */
#ifndef UMV_H_
#define UMV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UMV_ABI_VERSION 1

typedef enum umv_status {
    UMV_OK = 0,
    UMV_ERR_INVALID = -1,        /* bad argument            -> Python ValueError        */
    UMV_ERR_UNSUPPORTED = -2,    /* not implemented variant -> Python NotImplementedError */
    UMV_ERR_CUDA = -3,           /* CUDA runtime / launch failure -> RuntimeError        */
    UMV_ERR_NOMEM = -4,          /* KV page pool or workspace exhausted -> MemoryError  */
    UMV_ERR_STATE = -5           /* call order violated (e.g. weights not finalized) -> AssertionError */
} umv_status;

typedef enum umv_dtype { UMV_BF16 = 0, UMV_F32 = 1, UMV_I64 = 2, UMV_I32 = 3, UMV_U8 = 4 } umv_dtype;

/* Model geometry.  Reference: Qwen2Config (qwen2_navit.py:46-204), SiglipVisionConfig
 * (siglip_navit.py:21-99), BagelConfig (bagel.py:30-88); values come from the checkpoint's
 * llm_config.json / vit_config.json (interactive_vqa_inferencer.py:206-213). */
typedef struct umv_dims {
    int32_t hidden, heads, kv_heads, inter, layers, vocab;
    float rope_theta, rms_eps;
    int32_t vit_hidden, vit_heads, vit_inter, vit_layers, vit_patch_dim, vit_positions;
    float vit_eps;
    int32_t vit_pos_table;        /* vit_max_num_patch_per_side^2 (bagel.py:143)   */
    int32_t latent_dim;           /* latent_patch_size^2 * z_channels (bagel.py:113) */
    int32_t latent_pos_table;     /* max_latent_size^2 (bagel.py:120)              */
    int32_t max_tokens;           /* workspace rows: largest packed query length of one call */
    int32_t max_seqs;             /* largest number of samples in one call         */
    int32_t kv_pages;             /* KV page pool size (pages of UMV_PAGE_TOKENS tokens) */
    int32_t enable_vit, enable_gen;   /* allocate ViT / generation-expert weights  */
    int32_t enable_vae;               /* allocate the FLUX-style autoencoder (autoencoder.py:338-349 geometry) */
} umv_dims;

#define UMV_PAGE_TOKENS 64

typedef struct umv_engine umv_engine;

const char* umv_last_error(void);
int umv_abi_version(void);

/* ---- lifetime / weights ------------------------------------------------------------------
 * Stands in for Bagel.__init__ + load_checkpoint_and_dispatch(dtype=bf16)
 * (interactive_vqa_inferencer.py:225-229,153-156).  Tensor names are the reference state_dict
 * keys (SURVEY.md section 8b "Weight interface"); the engine re-lays them out (fused QKV, gate/up
 * interleaved in 64-row blocks). */
int umv_create(const umv_dims* dims, umv_engine** out);
int umv_destroy(umv_engine* e);
int umv_load_tensor(umv_engine* e, const char* name, const void* data, int dtype, int ndim,
                    const int64_t* shape, int data_on_device);
int umv_export_tensor(umv_engine* e, const char* name, void* host_dst, size_t bytes);
int umv_fill_synthetic(umv_engine* e, uint64_t seed);   /* random-init weights generated on the device */
int umv_finalize(umv_engine* e);
int umv_weight_bytes(umv_engine* e, int64_t* bytes);

/* ---- KV sequences: NaiveCache (qwen2_navit.py:207-221) as ref-counted pages -----------------
 * One sequence per sample.  fork == copy.deepcopy(gen_context) (inferencer.py:261,587,600,607)
 * without copying pages (copy-on-write of the partial tail page).  export gives the reference's
 * packed layout [len, kv_heads, head_dim] for parity tests. */
int umv_seq_new(umv_engine* e, int32_t* seq);
int umv_seq_fork(umv_engine* e, int32_t src, int32_t* dst);
int umv_seq_free(umv_engine* e, int32_t seq);
int umv_seq_len(umv_engine* e, int32_t seq, int32_t* len);
int umv_seq_truncate(umv_engine* e, int32_t seq, int32_t len);
int umv_seq_export(umv_engine* e, int32_t seq, int32_t layer, void* k_dev, void* v_dev, void* stream);
int umv_pages_free(umv_engine* e, int32_t* n);

/* ---- ViT + connector: SiglipVisionModel.forward (siglip_navit.py:389-402) + MLPconnector +
 * vit_pos_embed, i.e. bagel.py:581-594.  pixels: f32 [n_tokens, vit_patch_dim] (patchify order
 * h,w,p,q,c); pos_ids: i64 [n_tokens] device; seqlens: host i32 [n_images]; out: bf16
 * [n_tokens, hidden]. */
int umv_vit_embed(umv_engine* e, const float* pixels, const int64_t* pos_ids, const int32_t* seqlens,
                  int32_t n_images, void* out, void* stream);

/* ---- image pre-processing on the device: ToTensor + Normalize(0.5, 0.5) + patchify of
 * uint8 HWC images, i.e. data/transforms.py:104-115 (`x / 255`, `(x - 0.5) / 0.5` in fp32) followed by
 * data/data_utils.py:43-50 (patch vectors in (p_row, p_col, channel) order) and the flattened position ids of
 * data_utils.py:53-58.  images: device uint8, image i is [h_i, w_i, 3] at byte offset offsets[i]; hw: host i32
 * [n_images][2] (multiples of `patch`; resizing stays with the caller); out_pixels: f32 [sum_i h_i w_i / patch^2,
 * 3 patch^2]; out_pos_ids: i64 [n_tokens] (row * max_per_side + col).  Bit-identical to the host transforms. */
int umv_patchify_u8(const uint8_t* images, const int64_t* offsets, const int32_t* hw, int32_t n_images, int32_t patch,
                    int32_t max_per_side, float* out_pixels, int64_t* out_pos_ids, void* stream);

/* Bicubic, antialiased resize of 8-bit HWC images on the device: PIL.Image.resize((out_w, out_h), BICUBIC) as
 * MaxLongEdgeMinShortEdgeResize.forward performs it through torchvision F.resize (data/transforms.py:60-87).  Pillow's
 * 8-bit resampler restated (weights in double precision on the host, 22-bit fixed point, horizontal pass into an 8-bit
 * intermediate, vertical pass): bit-identical to Pillow.  src: u8 [in_h, in_w, 3] device; dst: u8 [out_h, out_w, 3]
 * device; workspace: device scratch of umv_resize_workspace_bytes(...) bytes (weight tables + intermediate image).
 * umv_resize_coefficients exposes one axis' table (host only: bounds i32 [out][2] = first tap, tap count; kk i32
 * [out][*ksize]); pass NULL tables to query ksize. */
int umv_resize_bicubic_u8(const uint8_t* src, int32_t in_h, int32_t in_w, uint8_t* dst, int32_t out_h, int32_t out_w,
                          void* workspace, int64_t workspace_bytes, void* stream);
int64_t umv_resize_workspace_bytes(int32_t in_h, int32_t in_w, int32_t out_h, int32_t out_w);
int umv_resize_coefficients(int32_t in_size, int32_t out_size, int32_t* bounds, int32_t* kk, int32_t* ksize);

/* ---- embedding rows: language_model.model.embed_tokens (bagel.py:438,577,1264) ------------- */
int umv_embed_tokens(umv_engine* e, const int64_t* ids, int32_t n, void* out, void* stream);

/* ---- packed LLM forward: Qwen2ForCausalLM.forward_inference (qwen2_navit.py:1243-1274).
 * x: bf16 [M, hidden] packed query sequence (samples contiguous, in seq order); q_lens / seqs:
 * host i32 [n_seqs]; positions: host i32 [M] (packed_query_position_ids); row_is_gen: host u8 [M]
 * or NULL (mode "und"); non-NULL selects mode "gen" with rows==1 routed to the *_moe_gen weights
 * (packed_vae_token_indexes) and rows==0 to the understanding weights (packed_text_indexes).
 * update_kv: update_past_key_values.  out: bf16 [M, hidden] final-normed hidden states (may be
 * NULL when only the cache update is wanted). */
int umv_llm_forward(umv_engine* e, const void* x, int32_t n_seqs, const int32_t* seqs, const int32_t* q_lens,
                    const int32_t* positions, const uint8_t* row_is_gen, int32_t is_causal, int32_t update_kv,
                    void* out, void* stream);

/* lm_head (bagel.py:1295): hidden bf16 [m, hidden] -> logits bf16 [m, vocab]. */
int umv_lm_head(umv_engine* e, const void* hidden, int32_t m, void* logits, void* stream);

/* ---- greedy / sampled decode loop: Bagel.generate_text (bagel.py:1236-1317) ----------------
 * Runs `n_steps` forward passes entirely on the device (no per-step host sync).  tokens_out: i64
 * [n_steps, n_seqs] device, row 0 = start tokens (the reference's layout).  forced_tokens (i64
 * [n_steps, n_seqs] device, optional) teacher-forces the inputs; logits_out (bf16
 * [n_steps, n_seqs, vocab] device, optional) receives every step's logits.  temperature <= 0 ->
 * argmax (ties -> lowest index, torch.argmax); > 0 -> softmax(logits/T) sampling with the engine's
 * own counter-based RNG (seed).  next_tokens_out (i64 [n_seqs] device, optional) receives the token
 * computed by the last step (the reference discards it; a caller that continues the loop in chunks
 * feeds it back as start_tokens).  KV lengths advance by n_steps. */
int umv_generate_text(umv_engine* e, int32_t n_seqs, const int32_t* seqs, const int64_t* start_tokens,
                      const int32_t* positions, int32_t n_steps, float temperature, uint64_t seed,
                      const int64_t* forced_tokens, int64_t* tokens_out, void* logits_out,
                      int64_t* next_tokens_out, void* stream);

/* ---- rectified-flow step: Bagel._forward_flow (bagel.py:989-1211) -------------------------
 * One velocity evaluation incl. up to three LLM forwards (main / cfg_text / cfg_img contexts) and
 * the CFG mix + renorm.  x_t: f32 [n_lat, latent_dim] device; v_out: f32 [n_lat, latent_dim].
 * Markers (packed_text_ids) are given per sequence: 2 ids each.  lat_lens: latent tokens per image.
 * renorm_type: 0 global (per image), 1 channel, 2 text_channel.  *_seqs may be NULL
 * when the corresponding scale <= 1.  v_is_f32_out tells the caller whether the reference would
 * have produced an fp32 (1) or bf16-valued (0) velocity. */
typedef struct umv_flow_args {
    int32_t n_seqs;
    const int32_t* seqs;            /* main context                                  */
    const int32_t* cfg_text_seqs;   /* or NULL                                        */
    const int32_t* cfg_img_seqs;    /* or NULL                                        */
    const int32_t* lat_lens;        /* host [n_seqs] h*w per image                    */
    const int32_t* positions;       /* host [n_seqs] rope position of each image (main) */
    const int32_t* cfg_text_positions;
    const int32_t* cfg_img_positions;
    const int64_t* marker_ids;      /* host [2] start_of_image, end_of_image          */
    const int64_t* lat_pos_ids;     /* device i64 [n_lat] packed_vae_position_ids     */
    float timestep;
    float cfg_text_scale, cfg_img_scale, cfg_renorm_min;
    int32_t renorm_type;
} umv_flow_args;
/* Number of CFG branches the last umv_flow_velocity call evaluated: 1 + (cfg_text) + (cfg_img), minus the cfg_img branch when its context
 * is known to hold the main context's tokens (a fork of it: pure text-to-image, inferencer.py:578-600) -- that velocity is the main
 * branch's, bit for bit, and is computed once.  Measurement / test hook. */
int umv_flow_branches_last(umv_engine* e, int32_t* n);
int umv_flow_velocity(umv_engine* e, const umv_flow_args* a, const float* x_t, float* v_out, void* stream);
/* x_t <- x_t - bf16(v * dt) (bagel.py:983; v*dt rounds to bf16 when v is bf16-valued). */
int umv_flow_euler(umv_engine* e, float* x_t, const float* v, int64_t n, float dt, int32_t v_is_bf16, void* stream);

/* Latent tokens entering the LLM (bagel.py:777-781, 1100-1107): bf16(bf16(vae2llm(x) + time_embedder(t)) +
 * latent_pos_embed[pos]).  x: f32 [n, latent_dim] device; pos_ids: i64 [n] device; out: bf16 [n, hidden]. */
int umv_latent_embed(umv_engine* e, const float* x, const int64_t* pos_ids, int32_t n, float timestep, void* out, void* stream);

/* ---- VAE: AutoEncoder.decode / Encoder (autoencoder.py:300-307, 122-257).  Weight names are the
 * reference's with the prefix "vae_model." (the submodule name inside Bagel, bagel.py:102).
 * decode: z bf16 [n, z_channels, h, w] (NCHW, the scaled latent decode_image builds, inferencer.py:239-241)
 *   -> bf16 image [n, 3, 8h, 8w].  encode_moments: x bf16 [n, 3, H, W] -> bf16 [n, 2*z_channels, H/8, W/8]
 *   (mean | logvar); the caller draws DiagonalGaussian's noise and applies scale/shift (autoencoder.py:266-303). */
int umv_vae_decode(umv_engine* e, const void* z, int32_t n, int32_t h, int32_t w, void* out, void* stream);
int umv_vae_encode_moments(umv_engine* e, const void* x, int32_t n, int32_t H, int32_t W, void* out, void* stream);

/* ---- the reference's prefill drivers, one call each: the packed query sequence (marker / text embeddings, ViT or latent
 * embeddings scattered to their packed rows) is composed inside the engine and run through the LLM with the cache update.
 * Index arguments are the `generation_input` tensors of the matching Bagel.prepare_* method, as HOST arrays:
 *   text:  Bagel.forward_cache_update_text (bagel.py:412-458): text_lens [n_seqs], text_ids [sum] (packed_text_ids), positions [sum]
 *          (packed_text_position_ids); causal.
 *   vit:   Bagel.forward_cache_update_vit (bagel.py:523-615): seq_lens = packed_seqlens, text_ids / text_rows = packed_text_ids /
 *          packed_text_indexes (the <vision_start> / <vision_end> markers), pixels / vit_pos_ids (DEVICE) / vit_seqlens as umv_vit_embed,
 *          vit_rows = packed_vit_token_indexes, positions = packed_position_ids; full attention inside the block.
 *   vae:   Bagel.forward_cache_update_vae (bagel.py:697-806) after vae_model.encode: latent = the encoded, padded batch bf16
 *          [n_images, latent_dim / patch^2, Hl, Wl] (DEVICE); latent_hw = patchified_vae_latent_shapes [n_images][2]; lat_pos_ids =
 *          packed_vae_position_ids (DEVICE); lat_rows = packed_vae_token_indexes; timestep = packed_timesteps (one value);
 *          generation-expert routing for the latent rows, full attention.
 * text_rows and the block rows must tile [0, sum(seq_lens)) exactly (UMV_ERR_INVALID otherwise). */
int umv_forward_cache_update_text(umv_engine* e, int32_t n_seqs, const int32_t* seqs, const int32_t* text_lens,
                                  const int64_t* text_ids, const int32_t* positions, void* stream);
int umv_forward_cache_update_vit(umv_engine* e, int32_t n_seqs, const int32_t* seqs, const int32_t* seq_lens, int32_t n_text,
                                 const int64_t* text_ids, const int32_t* text_rows, const float* pixels, const int64_t* vit_pos_ids,
                                 int32_t n_images, const int32_t* vit_seqlens, const int32_t* vit_rows, const int32_t* positions,
                                 void* stream);
/* Mixed prefill + decode in ONE forward (the reference's loop to replace: inferencer.py:552-638, which runs
 * a request's prefills and its decode strictly one after the other).  `riders` are running requests: rider i feeds token tokens[i] at
 * rope position positions[i] as ONE extra query row of sequence seqs[i], appended after the prefill rows of the same packed forward; its K/V
 * row is appended to its cache and next_tokens[i] (DEVICE i64 [n]) receives argmax (temperature <= 0) or a sample of its logits.  A single
 * query row attends to its whole context under either mask, so riders join the causal text prefill and the full-attention image block
 * alike.  riders == NULL or n == 0: exactly the plain driver.  n <= 64. */
typedef struct umv_decode_riders {
    int32_t n;
    const int32_t* seqs;        /* host [n] */
    const int64_t* tokens;      /* host [n] current input token of each rider */
    const int32_t* positions;   /* host [n] rope position of that token */
    float temperature;
    uint64_t seed;
    int64_t* next_tokens;       /* device [n] */
} umv_decode_riders;
int umv_forward_cache_update_text_riders(umv_engine* e, int32_t n_seqs, const int32_t* seqs, const int32_t* text_lens,
                                         const int64_t* text_ids, const int32_t* positions, const umv_decode_riders* riders,
                                         void* stream);
int umv_forward_cache_update_vit_riders(umv_engine* e, int32_t n_seqs, const int32_t* seqs, const int32_t* seq_lens, int32_t n_text,
                                        const int64_t* text_ids, const int32_t* text_rows, const float* pixels,
                                        const int64_t* vit_pos_ids, int32_t n_images, const int32_t* vit_seqlens,
                                        const int32_t* vit_rows, const int32_t* positions, const umv_decode_riders* riders,
                                        void* stream);
/* Image block AND prompt of every sample in ONE forward -- what the reference does as forward_cache_update_vit (bagel.py:522-615, full
 * mask) followed by forward_cache_update_text (bagel.py:411-458, causal) in the VQA drivers (inferencer.py:640-680,
 * interactive_vqa_inferencer.py:312-321), with one pass over the weights instead of two.  seq_lens[b] counts ALL packed rows of
 * sample b, the last prompt_lens[b] of which are the prompt (text rows: their ids are part of text_ids / text_rows, their rope positions
 * part of positions); those rows attend causally -- to the whole block and to the prompt rows before them -- while the block's rows see
 * the block only.  K / V and hidden states are bit-identical to the two separate calls.  prompt_lens == NULL: umv_forward_cache_update_vit. */
int umv_forward_cache_update_vit_prompt(umv_engine* e, int32_t n_seqs, const int32_t* seqs, const int32_t* seq_lens,
                                        const int32_t* prompt_lens, int32_t n_text, const int64_t* text_ids, const int32_t* text_rows,
                                        const float* pixels, const int64_t* vit_pos_ids, int32_t n_images, const int32_t* vit_seqlens,
                                        const int32_t* vit_rows, const int32_t* positions, const umv_decode_riders* riders,
                                        void* stream);
int umv_forward_cache_update_vae(umv_engine* e, int32_t n_seqs, const int32_t* seqs, const int32_t* seq_lens, int32_t n_text,
                                 const int64_t* text_ids, const int32_t* text_rows, const void* latent, int32_t n_images, int32_t Hl,
                                 int32_t Wl, const int32_t* latent_hw, int32_t patch, const int64_t* lat_pos_ids, const int32_t* lat_rows,
                                 float timestep, const int32_t* positions, void* stream);

/* ---- the inner module boundary, for layer-level parity against the reference's own sub-modules:
 *   umv_vit_model     vit_model(packed_pixel_values, packed_flattened_position_ids, cu_seqlens, max_seqlen) (siglip_navit.py:389-402)
 *                     -> bf16 [n_tokens, vit_hidden]: the post-layernorm rows as the next Linear sees them (the reference's LayerNorm
 *                     returns fp32 under CUDA autocast and the connector's first Linear rounds it to exactly this bf16 value)
 *   umv_connector     connector(x) (modeling_utils.py:119-123): bf16 [n, vit_hidden] -> bf16 [n, hidden]
 *   umv_pos_embed     vit_pos_embed(ids) (which = 0) / latent_pos_embed(ids) (which = 1) (modeling_utils.py:142-143): rows of the frozen table
 *   umv_vae2llm       vae2llm(x) (bagel.py:114): fp32 [n, latent_dim] (cast to bf16 by autocast) -> bf16 [n, hidden]
 *   umv_llm2vae       llm2vae(h) (bagel.py:115): bf16 [n, hidden] -> bf16 [n, latent_dim]
 *   umv_time_embedder time_embedder(t) (modeling_utils.py:73-109): HOST fp32 [n] -> bf16 [n, hidden]
 * language_model.forward_inference is umv_llm_forward, lm_head umv_lm_head, embed_tokens umv_embed_tokens, vae_model.encode / decode
 * umv_vae_encode_moments (+ umv_vae_sample) / umv_vae_decode. */
int umv_vit_model(umv_engine* e, const float* pixels, const int64_t* pos_ids, const int32_t* seqlens, int32_t n_images, void* out,
                  void* stream);
int umv_connector(umv_engine* e, const void* x, int32_t n, void* out, void* stream);
int umv_pos_embed(umv_engine* e, int32_t which, const int64_t* pos_ids, int32_t n, void* out, void* stream);
int umv_vae2llm(umv_engine* e, const float* x, int32_t n, void* out, void* stream);
int umv_llm2vae(umv_engine* e, const void* h, int32_t n, void* out, void* stream);
int umv_time_embedder(umv_engine* e, const float* timesteps, int32_t n, void* out, void* stream);

/* DiagonalGaussian sampling + AutoEncoder.encode's affine (autoencoder.py:266-272,300-303), op by op on bf16 as torch does:
 * out = scale * ((mean + exp(0.5 * logvar) * noise) - shift).  moments bf16 [n, 2z, h, w] (umv_vae_encode_moments), noise bf16
 * [n, z, h, w] drawn by the caller (the reference uses torch.randn_like) or NULL for sample=False; out bf16 [n, z, h, w]. */
int umv_vae_sample(umv_engine* e, const void* moments, const void* noise, int32_t n, int32_t h, int32_t w, void* out, void* stream);

/* InterleaveInferencer.decode_image (inferencer.py:234-256) on the device: packed latent tokens fp32 [n][h*w][patch^2 * z] (what
 * generate_image returns per image) -> unpatchify "nhwpqc->nchpwq" -> AutoEncoder.decode -> (x * 0.5 + 0.5).clamp(0, 1) * 255 -> uint8
 * with torch's bf16 rounding after every op and the truncating cast; out uint8 [n][8 h patch][8 w patch][3] (HWC, what Image.fromarray
 * takes).  n images of one size per call. */
int umv_decode_image_u8(umv_engine* e, const float* latent_tokens, int32_t n, int32_t h, int32_t w, int32_t patch, uint8_t* out,
                        void* stream);

/* ---- op-level entry points (parity tests; each is the kernel the model path uses) ---------- */
/* y[M,N] = x[M,K] @ w[N,K]^T (+bias) with epilogue `epi`: 0 bf16, 1 gelu_tanh, 2 swiglu (w rows
 * interleaved 64 gate | 64 up, y is [M, N/2]), 3 +residual (y = bf16(bf16(acc+bias) + res)).
 * impl: 0 auto, 1 tcgen05 token-major (prefill), 2 tcgen05 weight-major (decode, M<=32, split-K),
 * 3 simple reference kernel. */
int umv_op_linear(const void* x, const void* w, const void* bias, const void* residual, void* y, int32_t M,
                  int32_t N, int32_t K, int32_t epi, int32_t impl, void* stream);
int umv_op_rmsnorm(const void* x, const void* w, void* y, int32_t M, int32_t D, float eps, void* stream);
int umv_op_layernorm(const void* x, const void* w, const void* b, void* y, int32_t M, int32_t D, float eps,
                     void* stream);
/* varlen attention over packed q [Tq,H,dh], k/v [Tk,Hkv,dh] (flash_attn_varlen_func call sites
 * qwen2_navit.py:605-614, siglip_navit.py:232-241); q_lens/k_lens host i32 [n]. */
int umv_op_attention(const void* q, const void* k, const void* v, void* out, int32_t n, const int32_t* q_lens,
                     const int32_t* k_lens, int32_t heads, int32_t kv_heads, int32_t head_dim, int32_t causal,
                     void* stream);
/* The attention block of decoder layer `layer` ALONE, exactly as the model path runs it (PackedAttentionMoT.forward_inference after
 * the q/k/v projections, qwen2_navit.py:544-614): per-head q/k RMSNorm (the layer's q_norm / k_norm weights, *_moe_gen for rows with
 * row_is_gen) + RoPE at `positions` + append of the new K/V rows to the sequences' pages + varlen attention over past + new keys.
 * Projection outputs come either as bf16 rows `qkv` [M, (heads + 2 kv_heads) * head_dim] (bias applied) or -- as the decode step
 * produces them -- as `n_partials` fp32 split-K partials `qkv_partial` [n_partials][M][..] plus `bias` (bf16).  out: bf16
 * [M, heads * head_dim].  *path_out: kernel family that ran (1 mma.sync flash kernel, 2 tcgen05 flash kernel, 3 fused decode cluster
 * kernel); chosen as in the model (query rows per sample, UMV_ATTN_TC / UMV_FUSED_ATTN switches).  With update_kv the sequences
 * keep the appended rows.  Parity hook: lets a test compare attn_tc_kernel / attn_decode_kernel with the oracle directly. */
int umv_op_attention_block(umv_engine* e, int32_t layer, const void* qkv, const float* qkv_partial, int32_t n_partials,
                           const void* bias, int32_t n_seqs, const int32_t* seqs, const int32_t* q_lens, const int32_t* positions,
                           const uint8_t* row_is_gen, int32_t is_causal, int32_t update_kv, void* out, int32_t* path_out,
                           void* stream);
/* ONE block of the autoencoder (autoencoder.py:38-119) on a caller-provided activation, for block-level parity: `path` is the reference's
 * module path inside AutoEncoder -- "decoder.mid.block_1", "decoder.mid.attn_1", "decoder.up.3.block.0", "decoder.up.2.upsample" (nearest x2 +
 * conv), "encoder.down.0.downsample", "decoder.conv_in", "decoder.norm_out" (GroupNorm + swish) ... x: bf16 [C, H, W] (one image); out: bf16
 * [out_chw[0], out_chw[1], out_chw[2]] (the caller sizes it for 4x the input pixels and 512 channels at most). */
int umv_op_vae_block(umv_engine* e, const char* path, const void* x, int32_t C, int32_t H, int32_t W, void* out, int32_t* out_chw,
                     void* stream);
int umv_op_argmax(const void* logits, int32_t rows, int32_t vocab, int64_t* out, void* stream);
/* The sampling branch of generate_text (bagel.py:1297-1301: softmax(pred_logits / temperature) in fp32 + multinomial) as the
 * decode loop runs it: row r draws from softmax(bf16(logits[r] * (1/T))) by inverse CDF with u from the engine's counter-based
 * hash of (seed, step 0, r).  u_force >= 0 replaces u (test hook for the u * total == total rounding edge). */
int umv_op_sample(const void* logits, int32_t rows, int32_t vocab, float temperature, uint64_t seed, float u_force, int64_t* out,
                  void* stream);

/* Measurement hook (bench.py roofline): launch ONE weight-major decode linear of layer `layer` exactly
 * as the decode step does, on m rows of the engine's activation workspace.  which: 0 qkv (split-K
 * partials), 1 o_proj, 2 gate/up + SwiGLU, 3 down_proj, 4 lm_head.  *weight_bytes receives the bytes of
 * weights the launch streams (the algorithmic HBM traffic besides the m activation rows). */
int umv_bench_decode_linear(umv_engine* e, int32_t which, int32_t layer, int32_t m, int64_t* weight_bytes, void* stream);

/* Diagnostic timeline: after umv_trace_begin(max_slots) every traced kernel launch (linears, norms, attention) gets a
 * slot in launch order and stamps %globaltimer (ns) into it: [0] first CTA started, [1] it passed the dependency wait
 * (linears: all weight tiles requested), [2] first CTA ended, [3] last CTA ended.  Launches captured into a CUDA graph
 * keep their slot, so after a replayed decode loop the slots hold the timeline of the last step.  umv_trace_read
 * synchronises the device and copies min(n, max_slots) slots (12 x uint64 each: the 4 stamps + 8 kernel-specific ones) and kernel names (name_len bytes each).
 * umv_trace_begin(0) turns tracing off.  No reference counterpart (tools/decode_trace.py). */
int umv_trace_begin(int32_t max_slots);
int umv_trace_read(uint64_t* stamps, char* names, int32_t name_len, int32_t max_slots, int32_t* n);

/* Kernel launches issued by this library since load (bench.py "gpu_launches"). */
int64_t umv_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* UMV_H_ */
