"""Oracle: packed SigLIP-NaViT vision tower + connector (TEST INFRASTRUCTURE).

Follows siglip_navit.py:184-195 (embeddings), :202-244 (attention), :255-259 (MLP),
:271-300 (encoder layer), :345-371 (transformer, rope=False branch) and
modeling_utils.py:119-123 (MLPconnector), :142-143 (PositionEmbedding).
State-dict keys are the reference's (``vit_model.vision_model.*``, ``connector.*``,
``vit_pos_embed.pos_embed``).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import numerics as nm
from .numerics import BF16, Semantics


@dataclass
class ViTDims:
    hidden: int
    heads: int
    inter: int
    layers: int
    patch: int = 14
    channels: int = 3
    eps: float = 1e-6

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads


def vit_forward(sd, dims: ViTDims, packed_pixel_values, packed_flattened_position_ids, seqlens,
                sem: Semantics = Semantics.cuda, exact=True, taps=None):
    """SiglipVisionTransformer.forward (siglip_navit.py:345-371).  Returns the
    post-layernorm output: fp32 under cuda semantics, bf16 under cpu semantics."""
    P = "vit_model.vision_model."
    lin = lambda name, t: nm.linear(t, sd[P + name + ".weight"], sd[P + name + ".bias"], exact)
    x = lin("embeddings.patch_embedding", packed_pixel_values)                      # :190 (linearised conv)
    x = x + sd[P + "embeddings.position_embedding.weight"][packed_flattened_position_ids]   # :192 bf16+bf16
    if taps is not None:
        taps["vit_embed"] = x.clone()
    lens = [int(t) for t in seqlens]
    H, dh = dims.heads, dims.head_dim
    for li in range(dims.layers):
        L = f"encoder.layers.{li}."
        h = nm.layernorm(x, sd[P + L + "layer_norm1.weight"], sd[P + L + "layer_norm1.bias"], dims.eps, sem)
        q = lin(L + "self_attn.q_proj", h).view(-1, H, dh)
        k = lin(L + "self_attn.k_proj", h).view(-1, H, dh)
        v = lin(L + "self_attn.v_proj", h).view(-1, H, dh)
        a = nm.attention_varlen(q, k, v, lens, lens, causal=False, p_bf16=sem is Semantics.cuda).reshape(-1, dims.hidden)   # :232-241
        x = x + lin(L + "self_attn.out_proj", a)
        if taps is not None:
            taps[f"vit_attn{li}"] = x.clone()
        h = nm.layernorm(x, sd[P + L + "layer_norm2.weight"], sd[P + L + "layer_norm2.bias"], dims.eps, sem)
        h = lin(L + "mlp.fc1", h)
        h = nm.gelu_tanh(h)
        x = x + lin(L + "mlp.fc2", h)
        if taps is not None:
            taps[f"vit_layer{li}"] = x.clone()
    return nm.layernorm(x, sd[P + "post_layernorm.weight"], sd[P + "post_layernorm.bias"], dims.eps, sem)


def connector(sd, x, exact=True):
    """MLPconnector.forward, modeling_utils.py:119-123 (gelu_pytorch_tanh)."""
    h = nm.linear(x, sd["connector.fc1.weight"], sd["connector.fc1.bias"], exact)
    h = nm.gelu_tanh(h)
    return nm.linear(h, sd["connector.fc2.weight"], sd["connector.fc2.bias"], exact)


def vit_tokens_to_llm(sd, dims: ViTDims, packed_vit_tokens, packed_vit_position_ids, vit_token_seqlens,
                      sem: Semantics = Semantics.cuda, exact=True):
    """bagel.py:584-594: ViT -> connector -> + vit_pos_embed[pos] -> bf16 rows."""
    h = vit_forward(sd, dims, packed_vit_tokens, packed_vit_position_ids, vit_token_seqlens, sem, exact)
    h = connector(sd, h, exact)
    h = h + sd["vit_pos_embed.pos_embed"][packed_vit_position_ids]
    return h.to(BF16)
