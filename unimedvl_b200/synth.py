"""Deterministic synthetic weights / inputs with the reference's state-dict names.

There is no checkpoint offline (README.md:66-67 of the reference: weights are
downloaded), so parity tests and the benchmark run on random-init weights of the
reference architecture.  Every tensor is drawn from numpy's PCG64 (stream stable
across numpy versions and machines) seeded by crc32(name) ^ seed, so the golden
generator (which loads them into the *reference* modules), the oracle and the
engine all see bit-identical bf16 tensors without committing them.

Names/shapes: SURVEY.md section 8b "Weight interface" (qwen2_navit.py:389-407,713-731,
1032-1041,1186; siglip_navit.py:145-182,247-269,330-343; modeling_utils.py:73-143;
bagel.py:114-143; autoencoder.py:38-257).
"""
from __future__ import annotations

import math
import zlib

import numpy as np
import torch

from .config import BagelDims, LLMDims, ViTDims, VAEDims


def _key(name: str, seed: int) -> int:
    return (zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0xFFFFFFFF


def _rng(name: str, seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(_key(name, seed)))


_DEVICE = None      # set by on_device(): draw with torch's generator on that device instead of numpy on the host


class on_device:
    """``with synth.on_device("cuda"): sd = synth.bagel_state_dict(dims)`` draws every tensor with torch's Philox
    generator on the device (seeded per tensor name as on the host path).  For full-width (14B-dims) parity runs, where
    2e9 host-side normals would take minutes; the values differ from the host stream, so fixtures stay on the host path."""

    def __init__(self, device):
        self.device = torch.device(device)

    def __enter__(self):
        global _DEVICE
        self._prev, _DEVICE = _DEVICE, self.device
        return self

    def __exit__(self, *exc):
        global _DEVICE
        _DEVICE = self._prev


def _device_gen(name, seed):
    g = torch.Generator(device=_DEVICE)
    g.manual_seed(_key(name, seed))
    return g


def _normal(name, shape, std, seed, mean=0.0, dtype=torch.bfloat16):
    if _DEVICE is not None:
        a = torch.randn(tuple(shape), generator=_device_gen(name, seed), device=_DEVICE, dtype=torch.float32)
        return a.mul_(std).add_(mean).to(dtype)
    a = _rng(name, seed).standard_normal(size=shape, dtype=np.float32) * np.float32(std) + np.float32(mean)
    return torch.from_numpy(a).to(dtype)


def _uniform(name, shape, bound, seed, dtype=torch.bfloat16):
    if _DEVICE is not None:
        a = torch.rand(tuple(shape), generator=_device_gen(name, seed), device=_DEVICE, dtype=torch.float32)
        return a.mul_(2 * bound).sub_(bound).to(dtype)
    a = _rng(name, seed).random(size=shape, dtype=np.float32) * np.float32(2 * bound) - np.float32(bound)
    return torch.from_numpy(a).to(dtype)


def sincos_2d_table(embed_dim: int, grid: int) -> torch.Tensor:
    """Frozen PositionEmbedding table (modeling_utils.py:23-65,137-140), fp32 [grid*grid, embed_dim]:
    channels [0, D/2) encode the column index, [D/2, D) the row index, each as [sin | cos] over
    float64 angles pos / 10000^(2i/(D/2))."""
    quarter = embed_dim // 4
    omega = 1.0 / (10000.0 ** (np.arange(quarter, dtype=np.float64) / quarter))
    idx = np.arange(grid, dtype=np.float64)
    ang = idx[:, None] * omega[None, :]                                  # [grid, D/4]
    axis = np.concatenate([np.sin(ang), np.cos(ang)], axis=1)           # [grid, D/2]
    cols = np.tile(axis[None, :, :], (grid, 1, 1)).reshape(grid * grid, -1)
    rows = np.repeat(axis[:, None, :], grid, axis=1).reshape(grid * grid, -1)
    return torch.from_numpy(np.concatenate([cols, rows], axis=1)).float()


def llm_state_dict(d: LLMDims, seed: int = 0, std: float = 0.02) -> dict:
    sd = {}
    P = "language_model."
    D, H, Hkv, dh, I = d.hidden, d.heads, d.kv_heads, d.head_dim, d.inter
    sd[P + "model.embed_tokens.weight"] = _normal(P + "embed", (d.vocab, D), std, seed)
    sd[P + "lm_head.weight"] = _normal(P + "lm_head", (d.vocab, D), std, seed)
    for suffix in ("", "_moe_gen"):
        sd[P + f"model.norm{suffix}.weight"] = _normal(P + "norm" + suffix, (D,), 0.1, seed, mean=1.0)
    for i in range(d.layers):
        L = P + f"model.layers.{i}."
        for s in ("", "_moe_gen"):
            for n, rows in (("q", H * dh), ("k", Hkv * dh), ("v", Hkv * dh)):
                sd[L + f"self_attn.{n}_proj{s}.weight"] = _normal(L + n + s + "w", (rows, D), std, seed)
                sd[L + f"self_attn.{n}_proj{s}.bias"] = _normal(L + n + s + "b", (rows,), std, seed)
            sd[L + f"self_attn.o_proj{s}.weight"] = _normal(L + "o" + s, (D, H * dh), std, seed)
            for n in ("q", "k"):
                sd[L + f"self_attn.{n}_norm{s}.weight"] = _normal(L + n + "norm" + s, (dh,), 0.1, seed, mean=1.0)
            sd[L + f"mlp{s}.gate_proj.weight"] = _normal(L + "gate" + s, (I, D), std, seed)
            sd[L + f"mlp{s}.up_proj.weight"] = _normal(L + "up" + s, (I, D), std, seed)
            sd[L + f"mlp{s}.down_proj.weight"] = _normal(L + "down" + s, (D, I), std, seed)
            sd[L + f"input_layernorm{s}.weight"] = _normal(L + "ln1" + s, (D,), 0.1, seed, mean=1.0)
            sd[L + f"post_attention_layernorm{s}.weight"] = _normal(L + "ln2" + s, (D,), 0.1, seed, mean=1.0)
    return sd


def vit_state_dict(d: ViTDims, seed: int = 0) -> dict:
    sd = {}
    P = "vit_model.vision_model."
    Dv, Iv = d.hidden, d.inter
    lin = lambda n, o, i: (_normal(P + n + "w", (o, i), 1.0 / math.sqrt(i), seed), _normal(P + n + "b", (o,), 0.02, seed))
    sd[P + "embeddings.patch_embedding.weight"], sd[P + "embeddings.patch_embedding.bias"] = lin("patch", Dv, d.patch_dim)
    sd[P + "embeddings.position_embedding.weight"] = _normal(P + "pos", (d.num_positions, Dv), 0.02, seed)
    for i in range(d.layers):
        L = P + f"encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[L + f"self_attn.{n}.weight"], sd[L + f"self_attn.{n}.bias"] = lin(f"{i}{n}", Dv, Dv)
        sd[L + "mlp.fc1.weight"], sd[L + "mlp.fc1.bias"] = lin(f"{i}fc1", Iv, Dv)
        sd[L + "mlp.fc2.weight"], sd[L + "mlp.fc2.bias"] = lin(f"{i}fc2", Dv, Iv)
        for n in ("layer_norm1", "layer_norm2"):
            sd[L + n + ".weight"] = _normal(L + n + "w", (Dv,), 0.1, seed, mean=1.0)
            sd[L + n + ".bias"] = _normal(L + n + "b", (Dv,), 0.05, seed)
    sd[P + "post_layernorm.weight"] = _normal(P + "postw", (Dv,), 0.1, seed, mean=1.0)
    sd[P + "post_layernorm.bias"] = _normal(P + "postb", (Dv,), 0.05, seed)
    return sd


def glue_state_dict(d: BagelDims, seed: int = 0, std: float = 0.02) -> dict:
    """connector, vit_pos_embed, time_embedder, vae2llm, llm2vae, latent_pos_embed (bagel.py:114-143).
    llm2vae is zero-initialised in the reference (bagel.py:156-159); it is randomised here or v_t == 0."""
    sd = {}
    D, Dv, Z = d.llm.hidden, d.vit.hidden, d.patch_latent_dim
    sd["connector.fc1.weight"] = _normal("c.fc1w", (D, Dv), 1.0 / math.sqrt(Dv), seed)
    sd["connector.fc1.bias"] = _normal("c.fc1b", (D,), std, seed)
    sd["connector.fc2.weight"] = _normal("c.fc2w", (D, D), 1.0 / math.sqrt(D), seed)
    sd["connector.fc2.bias"] = _normal("c.fc2b", (D,), std, seed)
    sd["vit_pos_embed.pos_embed"] = sincos_2d_table(D, d.vit_max_num_patch_per_side).to(torch.bfloat16).to(_DEVICE or "cpu")
    sd["latent_pos_embed.pos_embed"] = sincos_2d_table(D, d.max_latent_size).to(torch.bfloat16).to(_DEVICE or "cpu")
    sd["time_embedder.mlp.0.weight"] = _normal("t.0w", (D, 256), 1.0 / 16.0, seed)
    sd["time_embedder.mlp.0.bias"] = _normal("t.0b", (D,), std, seed)
    sd["time_embedder.mlp.2.weight"] = _normal("t.2w", (D, D), 1.0 / math.sqrt(D), seed)
    sd["time_embedder.mlp.2.bias"] = _normal("t.2b", (D,), std, seed)
    sd["vae2llm.weight"] = _normal("vae2llm.w", (D, Z), 1.0 / math.sqrt(Z), seed)
    sd["vae2llm.bias"] = _normal("vae2llm.b", (D,), std, seed)
    sd["llm2vae.weight"] = _normal("llm2vae.w", (Z, D), 1.0 / math.sqrt(D), seed)
    sd["llm2vae.bias"] = _normal("llm2vae.b", (Z,), std, seed)
    return sd


def bagel_state_dict(d: BagelDims, seed: int = 0) -> dict:
    sd = llm_state_dict(d.llm, seed)
    sd.update(vit_state_dict(d.vit, seed))
    sd.update(glue_state_dict(d, seed))
    return sd


def vae_state_dict(d: VAEDims = VAEDims(), seed: int = 0, decoder: bool = True, encoder: bool = True) -> dict:
    """AutoEncoder.state_dict() names (autoencoder.py:122-257)."""
    sd = {}

    def conv(name, cout, cin, k):
        b = 1.0 / math.sqrt(cin * k * k)
        sd[name + ".weight"] = _uniform(name + "w", (cout, cin, k, k), b, seed)
        sd[name + ".bias"] = _uniform(name + "b", (cout,), b, seed)

    def gn(name, c):
        sd[name + ".weight"] = _normal(name + "w", (c,), 0.1, seed, mean=1.0)
        sd[name + ".bias"] = _normal(name + "b", (c,), 0.05, seed)

    def res(p, cin, cout):
        gn(p + "norm1", cin); conv(p + "conv1", cout, cin, 3)
        gn(p + "norm2", cout); conv(p + "conv2", cout, cout, 3)
        if cin != cout:
            conv(p + "nin_shortcut", cout, cin, 1)

    def attn(p, c):
        gn(p + "norm", c)
        for n in ("q", "k", "v", "proj_out"):
            conv(p + n, c, c, 1)

    nres = len(d.ch_mult)
    if encoder:
        P = "encoder."
        conv(P + "conv_in", d.ch, d.in_channels, 3)
        in_mult = (1,) + tuple(d.ch_mult)
        block_in = d.ch
        for lvl in range(nres):
            block_in, block_out = d.ch * in_mult[lvl], d.ch * d.ch_mult[lvl]
            for bi in range(d.num_res_blocks):
                res(f"{P}down.{lvl}.block.{bi}.", block_in, block_out)
                block_in = block_out
            if lvl != nres - 1:
                conv(f"{P}down.{lvl}.downsample.conv", block_in, block_in, 3)
        res(P + "mid.block_1.", block_in, block_in); attn(P + "mid.attn_1.", block_in); res(P + "mid.block_2.", block_in, block_in)
        gn(P + "norm_out", block_in); conv(P + "conv_out", 2 * d.z_channels, block_in, 3)
    if decoder:
        P = "decoder."
        block_in = d.ch * d.ch_mult[-1]
        conv(P + "conv_in", block_in, d.z_channels, 3)
        res(P + "mid.block_1.", block_in, block_in); attn(P + "mid.attn_1.", block_in); res(P + "mid.block_2.", block_in, block_in)
        for lvl in reversed(range(nres)):
            block_out = d.ch * d.ch_mult[lvl]
            for bi in range(d.num_res_blocks + 1):
                res(f"{P}up.{lvl}.block.{bi}.", block_in, block_out)
                block_in = block_out
            if lvl != 0:
                conv(f"{P}up.{lvl}.upsample.conv", block_in, block_in, 3)
        gn(P + "norm_out", block_in); conv(P + "conv_out", d.out_ch, block_in, 3)
    return sd


def synthetic_image(i: int, height: int = 448, width: int = 448) -> np.ndarray:
    """SURVEY.md section 8d synthetic input: uint8 HxWx3 noise, seed 1234+i."""
    return np.random.default_rng(1234 + i).integers(0, 256, (height, width, 3), dtype=np.uint8)


def synthetic_prompt_ids(i: int, n: int = 30, vocab_limit: int = 151643) -> list:
    """n random token ids in [0, vocab_limit), seed 4321+i."""
    return np.random.default_rng(4321 + i).integers(0, vocab_limit, n).tolist()
