"""Drop-in stand-in for the reference's ``Bagel`` module (codes/modeling/unimedvl/bagel.py:91).

Same method surface, argument names, dict keys and return types as the reference,
so ``codes/inferencer.py:InterleaveInferencer`` and the two ``interactive_*`` entry scripts can hold this
object instead of the torch module.  All device work goes to the CUDA engine through the C ABI; the
``prepare_*`` methods are host packing (unimedvl_b200/packing.py).  No torch.nn compute, no fallback.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional, Sequence, Tuple

import torch

from . import packing
from .cache import NaiveCache, paged_handle
from .config import BagelDims
from .engine import Engine


def _canonical(indexes: torch.Tensor, expected: torch.Tensor, what: str):
    """The engine addresses the cache by (sample, position); it accepts exactly the packed index layout
    the reference's prepare_* methods emit and refuses anything else instead of mis-addressing."""
    if indexes is None:
        return
    if indexes.numel() != expected.numel() or not torch.equal(indexes.to("cpu", torch.int64), expected):
        raise NotImplementedError(f"{what}: only the index layout produced by Bagel.prepare_* is supported")


class BaseNavitOutputWithPast:
    """qwen2_navit.py:224-227."""

    def __init__(self, packed_query_sequence=None, past_key_values=None):
        self.packed_query_sequence, self.past_key_values = packed_query_sequence, past_key_values

    def __iter__(self):
        return iter((self.packed_query_sequence, self.past_key_values))


class _Module:
    """A sub-module of the reference's Bagel as a callable over the engine (SURVEY.md section 8b "inner module boundary"):
    `module(x)`, `.parameters()` for device discovery, `.eval()`."""

    def __init__(self, engine: Engine, fn, dtype=torch.bfloat16):
        self._engine, self._fn = engine, fn
        self._probe = torch.empty(0, dtype=dtype, device=engine.device)

    def __call__(self, *a, **k):
        return self._fn(*a, **k)

    forward = __call__

    def parameters(self):
        yield self._probe

    def eval(self):
        return self


class _QwenBody:
    """language_model.model: only `embed_tokens` is reached from outside (bagel.py:438,577,1264; inferencer.py:99-101)."""

    def __init__(self, engine: Engine):
        self.embed_tokens = _Module(engine, lambda ids: engine.embed_tokens(ids.reshape(-1)).reshape(*ids.shape, -1))


class _LanguageModel:
    """Qwen2ForCausalLM as the reference's Bagel uses it: `.model.embed_tokens`, `.lm_head`, `.forward_inference(...)`
    (qwen2_navit.py:1243-1274) over umv_embed_tokens / umv_lm_head / umv_llm_forward."""

    def __init__(self, bagel: "Bagel"):
        self._b = bagel
        eng = bagel.engine
        self.model = _QwenBody(eng)
        self.lm_head = _Module(eng, lambda h: eng.lm_head(h.reshape(-1, h.shape[-1])).reshape(*h.shape[:-1], -1))

    def eval(self):
        return self

    @torch.no_grad()
    def forward_inference(self, packed_query_sequence, query_lens, packed_query_position_ids, packed_query_indexes,
                          past_key_values=None, key_values_lens=None, packed_key_value_indexes=None, update_past_key_values=True,
                          is_causal=True, mode="und", packed_vae_token_indexes=None, packed_text_indexes=None):
        """Same arguments and return shape as the reference.  The cache is addressed by (sample, position), so the index tensors
        must be the canonical layout Bagel.prepare_* emits (checked); `past_key_values` None runs against empty contexts."""
        b, eng = self._b, self._b.engine
        lens = b._ints(query_lens)
        if mode not in ("und", "gen"):
            raise ValueError(f"forward_inference: mode {mode!r}")
        cache = past_key_values if past_key_values is not None else NaiveCache(b.config.llm_config.num_hidden_layers)
        h = paged_handle(cache, eng, len(lens))
        kv = b._ints(key_values_lens) if key_values_lens is not None else [0] * len(lens)
        b._check_kv(h, kv, packed_key_value_indexes, lens, packed_query_indexes, "forward_inference")
        M = sum(lens)
        is_gen = None
        if mode == "gen":
            is_gen = [0] * M
            for i in b._ints(packed_vae_token_indexes):
                is_gen[i] = 1
            if packed_text_indexes is not None and sorted(b._ints(packed_text_indexes) + b._ints(packed_vae_token_indexes)) != list(range(M)):
                raise ValueError("forward_inference: packed_text_indexes + packed_vae_token_indexes must cover every packed row once")
        out = eng.llm_forward(packed_query_sequence, h.seqs, lens, b._ints(packed_query_position_ids), row_is_gen=is_gen,
                              is_causal=bool(is_causal), update_kv=bool(update_past_key_values), want_hidden=True)
        return BaseNavitOutputWithPast(packed_query_sequence=out, past_key_values=cache if update_past_key_values else past_key_values)


class Bagel:
    def __init__(self, engine: Engine, dims: BagelDims):
        self.engine = engine
        self.dims = dims
        l = dims.llm
        self.config = SimpleNamespace(
            llm_config=SimpleNamespace(num_hidden_layers=l.layers, hidden_size=l.hidden, layer_module="Qwen2MoTDecoderLayer"),
            vit_config=SimpleNamespace(patch_size=dims.vit.patch, hidden_size=dims.vit.hidden),
            vae_config=SimpleNamespace(downsample=dims.vae.downsample, z_channels=dims.vae.z_channels),
            visual_gen=True, visual_und=True, latent_patch_size=dims.latent_patch_size,
            max_latent_size=dims.max_latent_size, vit_max_num_patch_per_side=dims.vit_max_num_patch_per_side)
        self.hidden_size = l.hidden
        self.use_moe = True
        self.num_heads = l.heads
        self.latent_patch_size = dims.latent_patch_size
        self.latent_downsample = dims.latent_downsample
        self.max_latent_size = dims.max_latent_size
        self.latent_channel = dims.vae.z_channels
        self.patch_latent_dim = dims.patch_latent_dim
        self.vit_patch_size = dims.vit.patch
        self.vit_max_num_patch_per_side = dims.vit_max_num_patch_per_side
        self.vit_hidden_size = dims.vit.hidden
        self.device = engine.device
        # inner module boundary: the reference's sub-modules as callables over the C ABI
        self.language_model = _LanguageModel(self)
        e = engine

        def _vit(packed_pixel_values, packed_flattened_position_ids, cu_seqlens, max_seqlen=None):
            cu = self._ints(cu_seqlens)
            return e.vit_model(packed_pixel_values, packed_flattened_position_ids, [b - a for a, b in zip(cu[:-1], cu[1:])])
        self.vit_model = _Module(e, _vit)                                  # siglip_navit.py:389-402
        self.connector = _Module(e, lambda x: e.connector(x))                           # modeling_utils.py:119-123
        self.vit_pos_embed = _Module(e, lambda ids: e.pos_embed(0, ids))   # modeling_utils.py:142-143
        self.latent_pos_embed = _Module(e, lambda ids: e.pos_embed(1, ids))
        self.vae2llm = _Module(e, lambda x: e.vae2llm(x))                               # bagel.py:114
        self.llm2vae = _Module(e, lambda h: e.llm2vae(h))                               # bagel.py:115
        self.time_embedder = _Module(e, lambda t: e.time_embedder(t))                   # modeling_utils.py:73-109

    def eval(self):
        return self

    def parameters(self):
        return self.language_model.model.embed_tokens.parameters()

    # ------------------------------------------------------------------ host packing (reference signatures)
    def prepare_prompts(self, curr_kvlens, curr_rope, prompts, tokenizer, new_token_ids):
        return packing.prepare_prompts(curr_kvlens, curr_rope, prompts, tokenizer, new_token_ids)

    def prepare_vit_images(self, curr_kvlens, curr_rope, images, transforms, new_token_ids):
        return packing.prepare_vit_images(curr_kvlens, curr_rope, images, transforms, new_token_ids, self.vit_patch_size,
                                          self.vit_max_num_patch_per_side)

    def prepare_vae_images(self, curr_kvlens, curr_rope, images, transforms, new_token_ids, timestep=0):
        return packing.prepare_vae_images(curr_kvlens, curr_rope, images, transforms, new_token_ids, self.latent_downsample,
                                          self.max_latent_size, timestep)

    def prepare_vae_latent(self, curr_kvlens, curr_rope, image_sizes, new_token_ids):
        return packing.prepare_vae_latent(curr_kvlens, curr_rope, image_sizes, new_token_ids, self.latent_downsample,
                                          self.max_latent_size, self.patch_latent_dim)

    def prepare_vae_latent_cfg(self, curr_kvlens, curr_rope, image_sizes):
        return packing.prepare_vae_latent_cfg(curr_kvlens, curr_rope, image_sizes, self.latent_downsample)

    def prepare_start_tokens(self, curr_kvlens, curr_rope, new_token_ids):
        return packing.prepare_start_tokens(curr_kvlens, curr_rope, new_token_ids, device=self.device)

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _ints(t) -> List[int]:
        return [int(v) for v in (t.tolist() if torch.is_tensor(t) else t)]

    def _check_kv(self, handle, key_values_lens, packed_key_value_indexes, new_lens, packed_indexes, what):
        kv = self._ints(key_values_lens)
        have = handle.lens()
        if have != kv:
            raise ValueError(f"{what}: key_values_lens {kv} do not match the cache ({have})")
        exp_kv, starts = packing._layout(kv, new_lens)
        _canonical(packed_key_value_indexes, torch.as_tensor(exp_kv), what + ".packed_key_value_indexes")
        if packed_indexes is not None:
            import numpy as np
            exp_q = np.concatenate([np.arange(s, s + n) for s, n in zip(starts, new_lens)]) if new_lens else np.zeros(0, "int64")
            _canonical(packed_indexes, torch.as_tensor(exp_q), what + ".packed_indexes")

    # ------------------------------------------------------------------ prefill
    # Each driver is ONE C-ABI call: the engine composes the packed query sequence itself (marker / text embeddings and the ViT /
    # latent embeddings written at their packed rows) -- no torch index_put / permute / cat on the path.
    @torch.no_grad()
    def forward_cache_update_text(self, past_key_values, packed_text_ids, packed_text_position_ids, text_token_lens,
                                  packed_text_indexes, packed_key_value_indexes, key_values_lens, decode_riders=None):
        """bagel.py:412-458: embed -> LLM forward (mode und, causal) -> cache update.
        `decode_riders` (extension): (seqs, tokens, positions) of running requests that decode one token
        inside this same forward; their next tokens are left in ``self.rider_tokens`` (device i64)."""
        lens = self._ints(text_token_lens)
        h = paged_handle(past_key_values, self.engine, len(lens))
        self._check_kv(h, key_values_lens, packed_key_value_indexes, lens, packed_text_indexes, "forward_cache_update_text")
        self.rider_tokens = self.engine.forward_cache_update_text(h.seqs, lens, self._ints(packed_text_ids),
                                                                  self._ints(packed_text_position_ids), riders=decode_riders)
        return past_key_values

    @torch.no_grad()
    def forward_cache_update_vit(self, past_key_values, packed_text_ids, packed_text_indexes, packed_vit_tokens,
                                 packed_vit_token_indexes, packed_vit_position_ids, vit_token_seqlens, packed_position_ids,
                                 packed_seqlens, packed_indexes, packed_key_value_indexes, key_values_lens, decode_riders=None):
        """bagel.py:523-615: markers + ViT/connector embeddings -> LLM forward (mode und, full attention).  `decode_riders`: as in
        forward_cache_update_text."""
        lens = self._ints(packed_seqlens)
        h = paged_handle(past_key_values, self.engine, len(lens))
        self._check_kv(h, key_values_lens, packed_key_value_indexes, lens, packed_indexes, "forward_cache_update_vit")
        self.rider_tokens = self.engine.forward_cache_update_vit(
            h.seqs, lens, self._ints(packed_text_ids), self._ints(packed_text_indexes), packed_vit_tokens, packed_vit_position_ids,
            self._ints(vit_token_seqlens), self._ints(packed_vit_token_indexes), self._ints(packed_position_ids), riders=decode_riders)
        return past_key_values

    @torch.no_grad()
    def forward_cache_update_vae(self, vae_model, past_key_values, padded_images, patchified_vae_latent_shapes,
                                 packed_vae_position_ids, packed_timesteps, packed_vae_token_indexes, packed_text_ids,
                                 packed_text_indexes, packed_position_ids, packed_seqlens, packed_indexes, key_values_lens,
                                 packed_key_value_indexes):
        """bagel.py:697-806: VAE-encode the image, patchify the latent (chpwq->hwpqc), embed the latent tokens at
        timestep `packed_timesteps` and prefill them through the GENERATION expert (mode gen, full attention)."""
        lens = self._ints(packed_seqlens)
        h = paged_handle(past_key_values, self.engine, len(lens))
        self._check_kv(h, key_values_lens, packed_key_value_indexes, lens, packed_indexes, "forward_cache_update_vae")
        latent = vae_model.encode(padded_images.to(self.device))                 # [B, C, Hp/8, Wp/8] bf16
        # the reference hands the float tensor to time_embedder (bagel.py:777): no integer truncation of e.g. timestep=0.5
        ts = packed_timesteps.reshape(-1).float().tolist() if torch.is_tensor(packed_timesteps) else [float(v) for v in packed_timesteps]
        if any(v != ts[0] for v in ts):
            raise NotImplementedError("forward_cache_update_vae: one timestep per call (prepare_vae_images emits a constant)")
        self.engine.forward_cache_update_vae(h.seqs, lens, self._ints(packed_text_ids), self._ints(packed_text_indexes), latent,
                                             [(int(a), int(b)) for a, b in patchified_vae_latent_shapes], self.latent_patch_size,
                                             packed_vae_position_ids, self._ints(packed_vae_token_indexes), ts[0],
                                             self._ints(packed_position_ids))
        return past_key_values

    # ------------------------------------------------------------------ image generation (rectified flow)
    _RENORM = {"global": 0, "channel": 1, "text_channel": 2}

    def _flow_geometry(self, packed_seqlens, packed_position_ids):
        lens = self._ints(packed_seqlens)
        lat_lens = [n - 2 for n in lens]
        pos = self._ints(packed_position_ids)
        first, starts = 0, []
        for n in lens:
            starts.append(first)
            first += n
        return lens, lat_lens, [pos[s] for s in starts]

    @torch.no_grad()
    def generate_image(self, packed_text_ids, packed_text_indexes, packed_init_noises, packed_vae_position_ids,
                       packed_vae_token_indexes, packed_seqlens, packed_position_ids, packed_indexes, past_key_values,
                       key_values_lens, packed_key_value_indexes, num_timesteps: int = 24, timestep_shift: float = 1.0,
                       cfg_renorm_min: float = 0.0, cfg_renorm_type: str = "global", cfg_interval=(0, 1),
                       cfg_text_scale: float = 1.0, cfg_text_packed_query_indexes=None, cfg_text_packed_position_ids=None,
                       cfg_text_past_key_values=None, cfg_text_key_values_lens=None, cfg_text_packed_key_value_indexes=None,
                       cfg_img_scale: float = 1.0, cfg_img_packed_query_indexes=None, cfg_img_packed_position_ids=None,
                       cfg_img_past_key_values=None, cfg_img_key_values_lens=None, cfg_img_packed_key_value_indexes=None,
                       cfg_type: str = "parallel"):
        """bagel.py:901-986: shifted-time Euler integration of the guided velocity; returns a tuple of fp32
        [h*w, patch_latent_dim] latents (one per image, on the device).  Batching rule: the "global"
        renorm is taken per image, which equals the reference for its only supported case (batch 1, bagel.py:1196)."""
        if cfg_renorm_type not in self._RENORM:
            raise NotImplementedError(f"{cfg_renorm_type} is not suppoprted")
        lens, lat_lens, pos = self._flow_geometry(packed_seqlens, packed_position_ids)
        B = len(lens)
        ids = self._ints(packed_text_ids)
        if ids[0::2] != [ids[0]] * B or ids[1::2] != [ids[1]] * B:
            raise NotImplementedError("generate_image: every image must use the same start/end marker ids")
        h = paged_handle(past_key_values, self.engine, B)
        self._check_kv(h, key_values_lens, packed_key_value_indexes, lens, packed_indexes, "generate_image")

        used, temps = set(h.seqs), []

        def branch(cache, kv_lens, kv_idx, q_idx, pos_ids, what):
            if cache is None:
                return None
            hb = paged_handle(cache, self.engine, B)
            self._check_kv(hb, kv_lens, kv_idx, lens, q_idx, what)
            if used & set(hb.seqs):          # the same context object passed for two branches: give it its own pages
                hb = hb.fork()
                temps.append(hb)
            used.update(hb.seqs)
            return hb.seqs, self._flow_geometry(packed_seqlens, pos_ids)[2]
        cfg_text = branch(cfg_text_past_key_values, cfg_text_key_values_lens, cfg_text_packed_key_value_indexes,
                          cfg_text_packed_query_indexes, cfg_text_packed_position_ids, "generate_image.cfg_text")
        cfg_img = branch(cfg_img_past_key_values, cfg_img_key_values_lens, cfg_img_packed_key_value_indexes,
                         cfg_img_packed_query_indexes, cfg_img_packed_position_ids, "generate_image.cfg_img")
        dev = self.device
        x_t = packed_init_noises.to(dev, torch.float32).contiguous().clone()
        pos_ids = packed_vae_position_ids.to(dev)
        v = torch.empty_like(x_t)
        # schedule exactly as the reference computes it (fp32 tensors, bagel.py:937-940)
        timesteps = torch.linspace(1, 0, num_timesteps)
        timesteps = timestep_shift * timesteps / (1 + (timestep_shift - 1) * timesteps)
        dts = timesteps[:-1] - timesteps[1:]
        timesteps = timesteps[:-1]
        renorm = self._RENORM[cfg_renorm_type]
        for i, t in enumerate(timesteps):
            on = bool(t > cfg_interval[0] and t <= cfg_interval[1])
            ts_, is_ = (cfg_text_scale, cfg_img_scale) if on else (1.0, 1.0)
            use_text = cfg_text if ts_ > 1.0 else None
            # the reference evaluates the cfg_img branch only inside `if cfg_text_scale > 1.0` (bagel.py:1173-1207): without
            # text guidance it is computed and discarded -- skip the forward
            if ts_ > 1.0 and use_text is None:
                raise ValueError("cfg_text_scale > 1 needs cfg_text_past_key_values")
            if is_ > 1.0 and cfg_img is None:
                raise ValueError("cfg_img_scale > 1 needs cfg_img_past_key_values")
            if ts_ <= 1.0:
                is_ = 1.0
            use_img = cfg_img if is_ > 1.0 else None
            self.engine.flow_velocity(x_t, pos_ids, lat_lens, h.seqs, pos, ids[:2], float(t), use_text, use_img, ts_, is_,
                                      cfg_renorm_min, renorm, out=v)
            # velocity dtype in the reference: bf16 unless a per-token (fp32) norm scaled it
            v_is_bf16 = not (ts_ > 1.0 and renorm in (1, 2))
            self.engine.flow_euler(x_t, v, float(dts[i]), v_is_bf16)
        return x_t.split(lat_lens)

    # ------------------------------------------------------------------ decode
    @torch.no_grad()
    def generate_text(self, past_key_values, packed_key_value_indexes, key_values_lens, packed_start_tokens,
                      packed_query_position_ids, max_length: int, do_sample: bool = False, temperature: float = 1.0,
                      end_token_id: Optional[int] = None, seed: int = 0, chunk: int = 32, stop: str = "first"):
        """bagel.py:1236-1317.  Returns LongTensor [steps, B] whose row 0 is the start tokens.  The loop runs
        on the device; with ``end_token_id`` set it proceeds in chunks and stops -- as the reference does --
        when SAMPLE 0 produces the end token (bagel.py:1313), trimming the cache to the steps the reference
        would have executed.

        ``stop="each"`` (extension): every sample stops at its OWN end token.  The loop ends
        when all samples have produced it (or at ``max_length``), each sequence's cache is trimmed to the steps that sample
        executed, and rows of a finished sample are filled with ``end_token_id`` -- column b then equals what sample b
        alone would return with the reference rule."""
        if stop not in ("first", "each"):
            raise ValueError(f"generate_text: stop={stop!r} (expected 'first' or 'each')")
        B = len(self._ints(key_values_lens))
        h = paged_handle(past_key_values, self.engine, B)
        self._check_kv(h, key_values_lens, packed_key_value_indexes, [0] * B, None, "generate_text")
        start_lens = h.lens()
        tokens = self._ints(packed_start_tokens)
        pos = self._ints(packed_query_position_ids)
        temp = float(temperature) if do_sample else 0.0
        rows, done = [], 0
        executed = [None] * B                   # steps executed by sample b up to and including the one that produced EOS
        while done < max_length:
            n = max_length - done if end_token_id is None else min(chunk, max_length - done)
            toks, nxt = self.engine.generate_text(h.seqs, tokens, pos, n, temperature=temp, seed=seed + done,
                                                  return_next=True)
            rows.append(toks)
            done += n
            if end_token_id is not None:
                fed = torch.cat([toks[1:], nxt[None]], dim=0)          # token computed at each executed step, per sample
                hit = (fed == end_token_id)
                for b in range(B if stop == "each" else 1):
                    if executed[b] is None and bool(hit[:, b].any()):
                        executed[b] = done - n + int(torch.nonzero(hit[:, b])[0]) + 1
                if stop == "first" and executed[0] is not None:
                    out = torch.cat(rows, dim=0)[:executed[0]]
                    for s, l0 in zip(h.seqs, start_lens):
                        self.engine.seq_truncate(s, l0 + executed[0])
                    return out
                if stop == "each" and all(e is not None for e in executed):
                    break
            tokens = nxt.tolist()
            pos = [p + n for p in pos]
        out = torch.cat(rows, dim=0)
        if stop == "each" and end_token_id is not None:
            steps = [e if e is not None else done for e in executed]
            out = out[:max(steps)].clone()
            for b, (s, l0) in enumerate(zip(h.seqs, start_lens)):
                out[steps[b]:, b] = end_token_id
                self.engine.seq_truncate(s, l0 + steps[b])
        return out

    @torch.no_grad()
    def chat(self, tokenizer, new_token_ids, image_transform, images, prompt, max_length: int, do_sample: bool = False,
             temperature: float = 1.0):
        """bagel.py:1321-1392 (single-sample VQA helper)."""
        past_key_values = NaiveCache(self.config.llm_config.num_hidden_layers)
        newlens, new_rope = [0], [0]
        for image in images:
            g, newlens, new_rope = self.prepare_vit_images(newlens, new_rope, [image], image_transform, new_token_ids)
            past_key_values = self.forward_cache_update_vit(past_key_values, **g)
        g, newlens, new_rope = self.prepare_prompts(newlens, new_rope, [prompt], tokenizer, new_token_ids)
        past_key_values = self.forward_cache_update_text(past_key_values, **g)
        g = self.prepare_start_tokens(newlens, new_rope, new_token_ids)
        toks = self.generate_text(past_key_values=past_key_values, max_length=max_length, do_sample=do_sample,
                                  temperature=temperature, end_token_id=new_token_ids["eos_token_id"], **g)
        output = tokenizer.decode(toks[:, 0])
        return output.split("<|im_end|>")[0].split("<|im_start|>")[1]

    # ------------------------------------------------------------------ batched VQA job (bench / serving)
    @torch.no_grad()
    def vqa_generate_images(self, images: Sequence[torch.Tensor], prompt_ids: Sequence[Sequence[int]], new_token_ids: dict,
                            max_length: int, resize_transform=None, fused_prefill: bool = True) -> torch.Tensor:
        """vqa_generate from uint8 [H, W, 3] images in pinned host memory: the images go up as uint8 and are normalised /
        patchified on the device (Engine.patchify_u8).  Without `resize_transform` they must already be at their ViT size
        (sides multiples of the patch); with it (``vit_transform.resize_transform``, data/transforms.py:15-87) each image is
        first resized on the device to the size that rule gives (Engine.resize_u8)."""
        if resize_transform is not None:          # the ViT transform's resize rule, applied on the device (bit-identical to PIL)
            sized = []
            for im in images:
                w, h = resize_transform.target_size(int(im.shape[1]), int(im.shape[0]))
                sized.append(im if (h, w) == tuple(im.shape[:2]) else self.engine.resize_u8(im, h, w))
            images = sized
        pixels, pos, lens = self.engine.patchify_u8(images)
        return self.vqa_generate(pixels, pos, lens, prompt_ids, new_token_ids, max_length, fused_prefill)

    @torch.no_grad()
    def vqa_prefill(self, pixels: torch.Tensor, vit_pos_ids: torch.Tensor, vit_seqlens: Sequence[int],
                    prompt_ids: Sequence[Sequence[int]], new_token_ids: dict, fused_prefill: bool = True):
        """Context of a fresh VQA job for B samples (one image each): image block + prompt.  Returns (cache, kv_lens, rope)."""
        B = len(vit_seqlens)
        dev = self.device
        cache = NaiveCache(self.config.llm_config.num_hidden_layers)
        L = packing.image_prompt_layout([int(n) for n in vit_seqlens], prompt_ids, new_token_ids) if fused_prefill else None
        if L is not None and sum(L["seq_lens"]) <= self.engine.max_tokens:      # else: the engine's workspace only holds the two calls
            # image block and prompt of every sample in one pass over the weights (umv_forward_cache_update_vit_prompt): the prompt rows
            # attend causally behind the full-mask block; K / V are bit-identical to the two calls below
            h = paged_handle(cache, self.engine, B)
            self.engine.forward_cache_update_vit(h.seqs, L["seq_lens"], L["text_ids"], L["text_rows"], pixels, vit_pos_ids,
                                                 [int(n) for n in vit_seqlens], L["vit_rows"], L["positions"], prompt_lens=L["prompt_lens"])
            return cache, L["kv_lens"], L["rope"]
        zeros = [0] * B
        # image block: [<vision_start>, patches, <vision_end>] per sample, one rope position
        L = packing._image_block_layout(zeros, zeros, [int(n) for n in vit_seqlens], new_token_ids)
        g = dict(packed_text_ids=torch.as_tensor(L["text_ids"]), packed_text_indexes=torch.as_tensor(L["text_idx"]),
                 packed_vit_tokens=pixels.to(dev, non_blocking=True), packed_vit_token_indexes=torch.as_tensor(L["img_idx"]),
                 packed_vit_position_ids=vit_pos_ids.to(dev, non_blocking=True), vit_token_seqlens=list(vit_seqlens),
                 packed_position_ids=torch.as_tensor(L["pos"]), packed_seqlens=L["seqlens"],
                 packed_indexes=torch.as_tensor(L["packed_idx"]), packed_key_value_indexes=torch.as_tensor(L["kv_indexes"]),
                 key_values_lens=zeros)
        cache = self.forward_cache_update_vit(cache, **g)
        lens, rope = L["seqlens"], [1] * B

        class _Ids:
            def __init__(self, table): self.t = table
            def encode(self, i): return list(self.t[i])
        g, lens, rope = packing.prepare_prompts(lens, rope, list(range(B)), _Ids(prompt_ids), new_token_ids)
        cache = self.forward_cache_update_text(cache, **g)
        return cache, lens, rope

    @torch.no_grad()
    def vqa_generate(self, pixels: torch.Tensor, vit_pos_ids: torch.Tensor, vit_seqlens: Sequence[int],
                     prompt_ids: Sequence[Sequence[int]], new_token_ids: dict, max_length: int, fused_prefill: bool = True) -> torch.Tensor:
        """One packed VQA job for B samples (one image each): host (pinned) tensors in, host tokens out.
        The same sequence of calls the reference's chat() makes, batched over samples as
        InterleaveInferencer's packed API allows: ViT prefill -> prompt prefill -> greedy decode (fused_prefill: the two prefills in
        one forward)."""
        cache, lens, rope = self.vqa_prefill(pixels, vit_pos_ids, vit_seqlens, prompt_ids, new_token_ids, fused_prefill)
        g = packing.prepare_start_tokens(lens, rope, new_token_ids)
        toks = self.generate_text(past_key_values=cache, max_length=max_length, end_token_id=None, **g)
        out = torch.empty(toks.shape, dtype=toks.dtype, pin_memory=True)
        out.copy_(toks, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out
