"""Oracle: the unified model's inference drivers (TEST INFRASTRUCTURE).

Restates bagel.py:412-458 (forward_cache_update_text), :523-615 (…_vit), :697-806
(…_vae), :901-986 (generate_image), :989-1211 (_forward_flow incl. CFG / renorm),
:1236-1317 (generate_text).  Inputs are the ``generation_input`` dicts the
reference's ``prepare_*`` methods build (same keys and dtypes).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import llm, vit, vae
from . import numerics as nm
from .numerics import BF16, F32, Semantics


@dataclass
class BagelDims:
    llm: llm.LLMDims
    vit: vit.ViTDims | None
    vae: vae.VAEDims
    latent_patch_size: int = 2
    max_latent_size: int = 64
    vit_max_num_patch_per_side: int = 70

    @property
    def latent_downsample(self) -> int:
        return 8 * self.latent_patch_size          # vae downsample (autoencoder.py:341) * patch

    @property
    def patch_latent_dim(self) -> int:
        return self.latent_patch_size ** 2 * self.vae.z_channels


class BagelOracle:
    def __init__(self, sd: dict, dims: BagelDims, sem: Semantics = Semantics.cuda, exact: bool = True,
                 vae_sd: dict | None = None):
        self.sd, self.dims, self.sem, self.exact, self.vae_sd = sd, dims, sem, exact, vae_sd
        self.inv_freq = nm.default_inv_freq(dims.llm.head_dim, dims.llm.rope_theta)

    # ------------------------------------------------------------------ helpers
    def new_cache(self) -> llm.PackedKV:
        return llm.PackedKV(self.dims.llm.layers)

    def _llm(self, seq, qlens, pos, qidx, cache, kvlens, kvidx, update, causal, mode="und",
             vae_idx=None, text_idx=None, taps=None):
        return llm.forward_inference(self.sd, self.dims.llm, seq, qlens, pos, qidx, cache, kvlens, kvidx,
                                     update, causal, mode, vae_idx, text_idx, self.inv_freq, self.exact, taps,
                                     p_bf16=self.sem is Semantics.cuda)

    def time_embed(self, t: torch.Tensor) -> torch.Tensor:
        """TimestepEmbedder.forward, modeling_utils.py:106-109 (fp32 sinusoid -> autocast Linear ->
        SiLU on bf16 -> Linear)."""
        f = nm.timestep_frequencies(t)
        h = nm.linear(f, self.sd["time_embedder.mlp.0.weight"], self.sd["time_embedder.mlp.0.bias"], self.exact)
        h = nm.silu(h)
        return nm.linear(h, self.sd["time_embedder.mlp.2.weight"], self.sd["time_embedder.mlp.2.bias"], self.exact)

    # ------------------------------------------------------------------ prefill
    def forward_cache_update_text(self, cache, packed_text_ids, packed_text_position_ids, text_token_lens,
                                  packed_text_indexes, packed_key_value_indexes, key_values_lens, taps=None):
        """bagel.py:412-458."""
        emb = llm.embed(self.sd, packed_text_ids)
        _, cache = self._llm(emb, text_token_lens, packed_text_position_ids, packed_text_indexes, cache,
                             key_values_lens, packed_key_value_indexes, True, True, taps=taps)
        return cache

    def forward_cache_update_vit(self, cache, packed_text_ids, packed_text_indexes, packed_vit_tokens,
                                 packed_vit_token_indexes, packed_vit_position_ids, vit_token_seqlens,
                                 packed_position_ids, packed_seqlens, packed_indexes,
                                 packed_key_value_indexes, key_values_lens, taps=None):
        """bagel.py:523-615."""
        D = self.dims.llm.hidden
        emb = llm.embed(self.sd, packed_text_ids)
        seq = emb.new_zeros((int(sum(int(s) for s in packed_seqlens)), D))
        seq[packed_text_indexes] = emb
        v = vit.vit_tokens_to_llm(self.sd, self.dims.vit, packed_vit_tokens, packed_vit_position_ids,
                                  vit_token_seqlens, self.sem, self.exact)
        if taps is not None:
            taps["vit_embed"] = v.clone()
        seq[packed_vit_token_indexes] = v
        _, cache = self._llm(seq, packed_seqlens, packed_position_ids, packed_indexes, cache, key_values_lens,
                             packed_key_value_indexes, True, False, taps=taps)
        return cache

    def _latent_tokens(self, x_t, timestep, packed_vae_position_ids):
        """bagel.py:1100-1107 / :777-781: vae2llm(x) + time_embedder(t) + latent_pos_embed[pos] -> bf16."""
        pos = self.sd["latent_pos_embed.pos_embed"][packed_vae_position_ids]
        temb = self.time_embed(timestep)
        h = nm.linear(x_t, self.sd["vae2llm.weight"], self.sd["vae2llm.bias"], self.exact) + temb + pos
        return h.to(BF16)

    def forward_cache_update_vae(self, cache, padded_images, patchified_vae_latent_shapes,
                                 packed_vae_position_ids, packed_timesteps, packed_vae_token_indexes,
                                 packed_text_ids, packed_text_indexes, packed_position_ids, packed_seqlens,
                                 packed_indexes, key_values_lens, packed_key_value_indexes, noise=None,
                                 taps=None):
        """bagel.py:697-806.  ``noise`` injects DiagonalGaussian's sample (a13)."""
        D, p, C = self.dims.llm.hidden, self.dims.latent_patch_size, self.dims.vae.z_channels
        emb = llm.embed(self.sd, packed_text_ids)
        seq = emb.new_zeros((int(sum(int(s) for s in packed_seqlens)), D))
        seq[packed_text_indexes] = emb
        lat = vae.encode(self.vae_sd, padded_images, self.dims.vae, self.sem, noise)
        rows = []
        for z, (h, w) in zip(lat, patchified_vae_latent_shapes):
            z = z[:, :h * p, :w * p].reshape(C, h, p, w, p)
            rows.append(torch.einsum("chpwq->hwpqc", z).reshape(-1, p * p * C))
        packed_latent = torch.cat(rows, 0)
        seq[packed_vae_token_indexes] = self._latent_tokens(packed_latent, packed_timesteps,
                                                            packed_vae_position_ids)
        _, cache = self._llm(seq, packed_seqlens, packed_position_ids, packed_indexes, cache, key_values_lens,
                             packed_key_value_indexes, True, False, "gen", packed_vae_token_indexes,
                             packed_text_indexes, taps=taps)
        return cache

    # ------------------------------------------------------------------- decode
    def generate_text(self, cache, packed_key_value_indexes, key_values_lens, packed_start_tokens,
                      packed_query_position_ids, max_length, end_token_id=None, forced_tokens=None,
                      logits_out=None):
        """bagel.py:1236-1317, greedy branch.  ``forced_tokens`` ([max_length, B]) teacher-forces the
        inputs (row s is fed at step s) so per-step logits can be compared between implementations
        whatever their earlier argmax decisions were; ``logits_out`` collects the bf16 logits."""
        kv_lens = key_values_lens.clone().to(torch.int64)
        kvidx = packed_key_value_indexes.clone()
        pos = packed_query_position_ids.clone()
        curr = packed_start_tokens
        out = []
        step = 0
        B = len(kv_lens)
        while step < max_length:
            if forced_tokens is not None:
                curr = forced_tokens[step]
            out.append(curr)
            emb = llm.embed(self.sd, curr)
            qidx = torch.cumsum(kv_lens, 0) + torch.arange(B)
            parts = list(kvidx.split(kv_lens.tolist()))
            kvidx = torch.cat([p + i for i, p in enumerate(parts)])
            h, cache = self._llm(emb, torch.ones(B, dtype=torch.int64), pos, qidx, cache, kv_lens, kvidx,
                                 True, True)
            logits = llm.lm_head(self.sd, h, self.exact)
            if logits_out is not None:
                logits_out.append(logits)
            curr = torch.argmax(logits, dim=-1)
            parts = list(kvidx.split(kv_lens.tolist()))
            kvidx = torch.cat([torch.cat([p, p[-1:] + 1]) for p in parts])
            kv_lens = kv_lens + 1
            pos = pos + 1
            step += 1
            if end_token_id is not None and int(curr[0]) == end_token_id:
                break
        return torch.stack(out, 0)

    # --------------------------------------------------------------------- flow
    def _velocity(self, seq, g, cache, pos, qidx, kvlens, kvidx):
        h, _ = self._llm(seq, g["packed_seqlens"], pos, qidx, cache, kvlens, kvidx, False, False, "gen",
                         g["packed_vae_token_indexes"], g["packed_text_indexes"])
        v = nm.linear(h, self.sd["llm2vae.weight"], self.sd["llm2vae.bias"], self.exact)
        return v[g["packed_vae_token_indexes"]]

    def _norm(self, x, dim=None):
        """torch.norm under autocast: cuda -> fp32-policy (fp32 result), cpu -> stays bf16."""
        if self.sem is Semantics.cuda:
            x = x.float()
        return torch.norm(x) if dim is None else torch.norm(x, dim=dim, keepdim=True)

    def forward_flow(self, x_t, timestep, g, cache, cfg_text=None, cfg_img=None, cfg_text_scale=1.0,
                     cfg_img_scale=1.0, cfg_renorm_type="global", cfg_renorm_min=0.0, per_image_global=False):
        """bagel.py:989-1211.  ``cfg_text`` / ``cfg_img``: dicts with ``cache`` + the four
        ``cfg_*`` tensors of prepare_vae_latent_cfg.  ``per_image_global`` computes the "global"
        norm per image instead of over the whole pack (the engine's batching rule)."""
        D = self.dims.llm.hidden
        emb = llm.embed(self.sd, g["packed_text_ids"])
        seq = emb.new_zeros((int(sum(int(s) for s in g["packed_seqlens"])), D))
        seq[g["packed_text_indexes"]] = emb
        assert timestep.unique().shape[0] == 1
        seq[g["packed_vae_token_indexes"]] = self._latent_tokens(x_t, timestep, g["packed_vae_position_ids"])

        v_t = self._velocity(seq, g, cache, g["packed_position_ids"], g["packed_indexes"],
                             g["key_values_lens"], g["packed_key_value_indexes"])
        if cfg_text_scale > 1.0:
            c = cfg_text
            v_text = self._velocity(seq, g, c["cache"], c["cfg_packed_position_ids"], c["cfg_packed_query_indexes"],
                                    c["cfg_key_values_lens"], c["cfg_packed_key_value_indexes"])
        if cfg_img_scale > 1.0:
            c = cfg_img
            v_img = self._velocity(seq, g, c["cache"], c["cfg_packed_position_ids"], c["cfg_packed_query_indexes"],
                                   c["cfg_key_values_lens"], c["cfg_packed_key_value_indexes"])
        if cfg_text_scale > 1.0:
            if cfg_renorm_type == "text_channel":
                v_text_ = v_text + cfg_text_scale * (v_t - v_text)
                n0 = self._norm(v_t, -1)
                n1 = self._norm(v_text_, -1)
                scale = (n0 / (n1 + 1e-8)).clamp(min=cfg_renorm_min, max=1.0)
                v_tt = v_text_ * scale
                v_t = v_img + cfg_img_scale * (v_tt - v_img) if cfg_img_scale > 1.0 else v_tt
            else:
                v_text_ = v_text + cfg_text_scale * (v_t - v_text)
                v_ = v_img + cfg_img_scale * (v_text_ - v_img) if cfg_img_scale > 1.0 else v_text_
                if cfg_renorm_type == "global":
                    if per_image_global:
                        lens = [int(s) - 2 for s in g["packed_seqlens"]]
                        n0 = torch.cat([self._norm(a).expand(a.shape[0], 1) for a in v_t.split(lens)])
                        n1 = torch.cat([self._norm(a).expand(a.shape[0], 1) for a in v_.split(lens)])
                        scale = (n0 / (n1 + 1e-8)).clamp(min=cfg_renorm_min, max=1.0)
                        # a 0-dim scale leaves v_ in bf16 (type promotion); keep that per image
                        v_t = torch.cat([a * s[0, 0] for a, s in zip(v_.split(lens), scale.split(lens))])
                        return v_t
                    n0, n1 = self._norm(v_t), self._norm(v_)
                elif cfg_renorm_type == "channel":
                    n0, n1 = self._norm(v_t, -1), self._norm(v_, -1)
                else:
                    raise NotImplementedError(f"{cfg_renorm_type} is not suppoprted")
                scale = (n0 / (n1 + 1e-8)).clamp(min=cfg_renorm_min, max=1.0)
                v_t = v_ * scale
        return v_t

    def generate_image(self, g, cache, cfg_text=None, cfg_img=None, num_timesteps=24, timestep_shift=1.0,
                       cfg_renorm_min=0.0, cfg_renorm_type="global", cfg_interval=(0, 1), cfg_text_scale=1.0,
                       cfg_img_scale=1.0, per_image_global=False, trace=None):
        """bagel.py:901-986 (shifted-time Euler integration; x_t stays fp32, R9)."""
        x_t = g["packed_init_noises"]
        ts = torch.linspace(1, 0, num_timesteps)
        ts = timestep_shift * ts / (1 + (timestep_shift - 1) * ts)
        dts = ts[:-1] - ts[1:]
        ts = ts[:-1]
        for i, t in enumerate(ts):
            timestep = torch.tensor([t] * x_t.shape[0])
            on = bool(t > cfg_interval[0] and t <= cfg_interval[1])
            v = self.forward_flow(x_t, timestep, g, cache, cfg_text, cfg_img,
                                  cfg_text_scale if on else 1.0, cfg_img_scale if on else 1.0,
                                  cfg_renorm_type, cfg_renorm_min, per_image_global)
            if trace is not None:
                trace.append(v.clone())
            x_t = x_t - v * dts[i]
        return x_t.split([int(s) - 2 for s in g["packed_seqlens"]])

    def decode_image(self, latent, image_shape):
        """inferencer.py:234-256 (uint8 HWC)."""
        H, W = image_shape
        d, p, C = self.dims.latent_downsample, self.dims.latent_patch_size, self.dims.vae.z_channels
        h, w = H // d, W // d
        z = latent.reshape(1, h, w, p, p, C)
        z = torch.einsum("nhwpqc->nchpwq", z).reshape(1, C, h * p, w * p).to(BF16)
        return vae.image_to_uint8(vae.decode(self.vae_sd, z, self.dims.vae, self.sem))
