"""Oracle: packed Qwen2 Mixture-of-Transformers decoder, inference path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Functional restatement over a flat
``state_dict`` whose keys are the reference's (prefix ``language_model.``).
Follows qwen2_navit.py:525-626 (PackedAttentionMoT.forward_inference),
:843-902 (Qwen2MoTDecoderLayer.forward_inference), :1115-1176
(Qwen2Model.forward_inference) and modeling_qwen2.py:80-97,164-235.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch

from . import numerics as nm
from .numerics import BF16, F32


@dataclass
class LLMDims:
    hidden: int
    heads: int
    kv_heads: int
    inter: int
    layers: int
    vocab: int
    rope_theta: float = 1e6
    eps: float = 1e-6

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads


@dataclass
class PackedKV:
    """NaiveCache (qwen2_navit.py:207-221): per layer one packed K and one packed V
    tensor [sum_ctx, kv_heads, head_dim], samples contiguous in batch order."""
    num_layers: int
    key: dict = field(default_factory=dict)
    value: dict = field(default_factory=dict)

    def __post_init__(self):
        for i in range(self.num_layers):
            self.key.setdefault(i, None)
            self.value.setdefault(i, None)

    def clone(self) -> "PackedKV":
        c = PackedKV(self.num_layers)
        for i in range(self.num_layers):
            c.key[i] = None if self.key[i] is None else self.key[i].clone()
            c.value[i] = None if self.value[i] is None else self.value[i].clone()
        return c


def _w(sd, name):
    return sd["language_model." + name]


def _route(fn_text, fn_vae, x, text_idx, vae_idx, out_dtype=None, out_cols=None):
    """MoT row routing (gen mode): rows text_idx through the understanding
    weights, rows vae_idx through the ``_moe_gen`` weights, scattered into a
    zero tensor (qwen2_navit.py:548-562,863-866,891-899)."""
    yt = fn_text(x[text_idx])
    yv = fn_vae(x[vae_idx])
    cols = yt.shape[-1] if out_cols is None else out_cols
    out = torch.zeros((x.shape[0], cols), dtype=out_dtype or yt.dtype)
    out[text_idx] = yt.to(out.dtype)
    out[vae_idx] = yv.to(out.dtype)
    return out


def attention_block(sd, dims: LLMDims, li: int, x, query_lens, cos, sin, packed_query_indexes,
                    cache: PackedKV | None, key_values_lens, packed_key_value_indexes,
                    update_cache: bool, is_causal: bool, mode: str, text_idx, vae_idx, exact=True, p_bf16=True):
    """PackedAttentionMoT.forward_inference, qwen2_navit.py:525-626."""
    p = f"model.layers.{li}.self_attn."
    H, Hkv, dh, eps = dims.heads, dims.kv_heads, dims.head_dim, dims.eps
    lin = lambda n, t: nm.linear(t, _w(sd, p + n + ".weight"), sd.get("language_model." + p + n + ".bias"), exact)
    if mode == "und":
        q = lin("q_proj", x).view(-1, H, dh)
        k = lin("k_proj", x).view(-1, Hkv, dh)
        v = lin("v_proj", x).view(-1, Hkv, dh)
        q = nm.rmsnorm(q, _w(sd, p + "q_norm.weight"), eps)            # bf16 in/out (R4)
        k = nm.rmsnorm(k, _w(sd, p + "k_norm.weight"), eps)
    else:
        x = x.to(BF16)
        q = _route(lambda t: lin("q_proj", t), lambda t: lin("q_proj_moe_gen", t), x, text_idx, vae_idx)
        k = _route(lambda t: lin("k_proj", t), lambda t: lin("k_proj_moe_gen", t), x, text_idx, vae_idx)
        v = _route(lambda t: lin("v_proj", t), lambda t: lin("v_proj_moe_gen", t), x, text_idx, vae_idx)
        q = q.view(-1, H, dh).to(F32)                                   # :568 up-cast before the norm
        k = k.view(-1, Hkv, dh).to(F32)
        v = v.view(-1, Hkv, dh)
        qn, kn = q.clone(), k.clone()
        qn[text_idx] = nm.rmsnorm(q[text_idx], _w(sd, p + "q_norm.weight"), eps)
        qn[vae_idx] = nm.rmsnorm(q[vae_idx], _w(sd, p + "q_norm_moe_gen.weight"), eps)
        kn[text_idx] = nm.rmsnorm(k[text_idx], _w(sd, p + "k_norm.weight"), eps)
        kn[vae_idx] = nm.rmsnorm(k[vae_idx], _w(sd, p + "k_norm_moe_gen.weight"), eps)
        q, k = qn, kn
    q, k = nm.apply_rope(q, k, cos, sin)
    q, k, v = q.to(BF16), k.to(BF16), v.to(BF16)                        # :581-583

    qlens = [int(t) for t in query_lens]
    if cache is not None and cache.key[li] is not None:
        past_k, past_v = cache.key[li], cache.value[li]
        total = sum(qlens) + sum(int(t) for t in key_values_lens)
        mk = past_k.new_zeros((total, Hkv, dh))
        mv = past_k.new_zeros((total, Hkv, dh))
        mk[packed_query_indexes] = k
        mk[packed_key_value_indexes] = past_k
        mv[packed_query_indexes] = v
        mv[packed_key_value_indexes] = past_v
        klens = [int(a) + b for a, b in zip(key_values_lens, qlens)]
    else:
        mk, mv, klens = k, v, qlens

    o = nm.attention_varlen(q, mk, mv, qlens, klens, causal=is_causal, p_bf16=p_bf16).reshape(-1, dims.hidden)
    if mode == "und":
        o = nm.linear(o, _w(sd, p + "o_proj.weight"), None, exact)
    else:
        # :619-620 writes the routed projections back IN PLACE into the bf16 attention output.
        o2 = o.clone()
        o2[text_idx] = nm.linear(o[text_idx], _w(sd, p + "o_proj.weight"), None, exact)
        o2[vae_idx] = nm.linear(o[vae_idx], _w(sd, p + "o_proj_moe_gen.weight"), None, exact)
        o = o2
    if update_cache:
        cache.key[li], cache.value[li] = mk, mv
    return o


def mlp(sd, prefix: str, x, exact=True):
    """Qwen2MLP.forward, modeling_qwen2.py:234-235: down(silu(gate(x)) * up(x)); every
    intermediate is a bf16 tensor (R7)."""
    g = nm.linear(x, _w(sd, prefix + "gate_proj.weight"), None, exact)
    u = nm.linear(x, _w(sd, prefix + "up_proj.weight"), None, exact)
    return nm.linear(nm.silu(g) * u, _w(sd, prefix + "down_proj.weight"), None, exact)


def decoder_layer(sd, dims, li, x, query_lens, cos, sin, packed_query_indexes, cache, key_values_lens,
                  packed_key_value_indexes, update_cache, is_causal, mode, text_idx, vae_idx, exact=True, p_bf16=True):
    """Qwen2MoTDecoderLayer.forward_inference, qwen2_navit.py:843-902."""
    p = f"model.layers.{li}."
    eps = dims.eps
    residual = x
    if mode == "und":
        h = nm.rmsnorm(x, _w(sd, p + "input_layernorm.weight"), eps)
    else:
        h = torch.zeros_like(x)
        h[text_idx] = nm.rmsnorm(x[text_idx], _w(sd, p + "input_layernorm.weight"), eps)
        h[vae_idx] = nm.rmsnorm(x[vae_idx], _w(sd, p + "input_layernorm_moe_gen.weight"), eps)
    h = attention_block(sd, dims, li, h, query_lens, cos, sin, packed_query_indexes, cache, key_values_lens,
                        packed_key_value_indexes, update_cache, is_causal, mode, text_idx, vae_idx, exact, p_bf16)
    x = residual + h
    residual = x
    if mode == "und":
        h = mlp(sd, p + "mlp.", nm.rmsnorm(x, _w(sd, p + "post_attention_layernorm.weight"), eps), exact)
    else:
        ht = nm.rmsnorm(x[text_idx], _w(sd, p + "post_attention_layernorm.weight"), eps).to(BF16)
        hv = nm.rmsnorm(x[vae_idx], _w(sd, p + "post_attention_layernorm_moe_gen.weight"), eps).to(BF16)
        h = torch.zeros_like(x).to(BF16)
        h[text_idx] = mlp(sd, p + "mlp.", ht, exact)
        h[vae_idx] = mlp(sd, p + "mlp_moe_gen.", hv, exact)
    return residual + h


def forward_inference(sd, dims: LLMDims, packed_query_sequence, query_lens, packed_query_position_ids,
                      packed_query_indexes, past_key_values: PackedKV | None = None, key_values_lens=None,
                      packed_key_value_indexes=None, update_past_key_values=True, is_causal=True,
                      mode="und", packed_vae_token_indexes=None, packed_text_indexes=None,
                      inv_freq=None, exact=True, taps=None, p_bf16=True):
    """Qwen2Model.forward_inference, qwen2_navit.py:1115-1176.  Returns
    (final-normed hidden [M, D], cache).  ``taps`` (optional dict) receives the
    hidden state after every layer for layer-level parity tests."""
    if inv_freq is None:
        inv_freq = nm.default_inv_freq(dims.head_dim, dims.rope_theta)
    x = packed_query_sequence
    cos, sin = nm.rope_cos_sin(packed_query_position_ids, inv_freq, x.dtype)
    for li in range(dims.layers):
        x = decoder_layer(sd, dims, li, x, query_lens, cos, sin, packed_query_indexes, past_key_values,
                          key_values_lens, packed_key_value_indexes, update_past_key_values, is_causal,
                          mode, packed_text_indexes, packed_vae_token_indexes, exact, p_bf16)
        if taps is not None:
            taps[f"layer{li}"] = x.clone()
    if mode == "und":
        x = nm.rmsnorm(x, _w(sd, "model.norm.weight"), dims.eps)
    else:
        y = torch.zeros_like(x)
        y[packed_text_indexes] = nm.rmsnorm(x[packed_text_indexes], _w(sd, "model.norm.weight"), dims.eps)
        y[packed_vae_token_indexes] = nm.rmsnorm(x[packed_vae_token_indexes],
                                                  _w(sd, "model.norm_moe_gen.weight"), dims.eps)
        x = y
    return x, past_key_values


def embed(sd, ids: torch.Tensor) -> torch.Tensor:
    """language_model.model.embed_tokens (R1: bf16 table row)."""
    return _w(sd, "model.embed_tokens.weight")[ids]


def lm_head(sd, x: torch.Tensor, exact=True) -> torch.Tensor:
    """bagel.py:1295: bf16 logits (R8)."""
    return nm.linear(x, _w(sd, "lm_head.weight"), None, exact)
