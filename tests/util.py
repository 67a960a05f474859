"""Shared test helpers: golden fixture access, oracle construction, bf16 ulp metrics."""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import model as omodel, llm as ollm, vit as ovit, vae as ovae
from oracle.numerics import Semantics
from unimedvl_b200 import config as ucfg, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOK = dict(bos_token_id=2040, eos_token_id=2041, start_of_image=2042, end_of_image=2043)

_DT = {"torch.int64": torch.int64, "torch.int32": torch.int32, "torch.float32": torch.float32,
       "torch.bfloat16": torch.bfloat16, "torch.uint8": torch.uint8}


class Golden:
    def __init__(self, name: str):
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))

    def has(self, key):
        return key in self.z.files

    def t(self, key: str) -> torch.Tensor:
        a = self.z[key]
        dt = str(self.z[key + ".dtype"]) if (key + ".dtype") in self.z.files else None
        if a.dtype == np.uint16:
            return torch.from_numpy(a.copy()).view(torch.bfloat16)
        t = torch.from_numpy(a.copy())
        return t.to(_DT[dt]) if dt in _DT else t

    def group(self, prefix: str) -> dict:
        """All tensors stored as ``prefix.key`` -> {key: tensor}, i.e. the generation_input dict."""
        out = {}
        for f in self.z.files:
            if f.startswith(prefix + ".") and not f.endswith(".dtype"):
                k = f[len(prefix) + 1:]
                if "." not in k:
                    out[k] = self.t(f)
        return out


def oracle_dims(d: ucfg.BagelDims) -> omodel.BagelDims:
    l, v, a = d.llm, d.vit, d.vae
    return omodel.BagelDims(
        llm=ollm.LLMDims(l.hidden, l.heads, l.kv_heads, l.inter, l.layers, l.vocab, l.rope_theta, l.eps),
        vit=ovit.ViTDims(v.hidden, v.heads, v.inter, v.layers, v.patch, v.channels, v.eps),
        vae=ovae.VAEDims(a.ch, tuple(a.ch_mult), a.num_res_blocks, a.z_channels, a.in_channels, a.out_ch,
                         a.scale_factor, a.shift_factor),
        latent_patch_size=d.latent_patch_size, max_latent_size=d.max_latent_size,
        vit_max_num_patch_per_side=d.vit_max_num_patch_per_side)


_CACHE = {}


def tiny_weights(seed: int = 0, vae: bool = False):
    key = (seed, vae)
    if key not in _CACHE:
        d = ucfg.tiny()
        _CACHE[key] = (d, synth.bagel_state_dict(d, seed), synth.vae_state_dict(d.vae, seed) if vae else None)
    return _CACHE[key]


def make_oracle(sem: Semantics, seed: int = 0, vae: bool = False, exact: bool = True) -> omodel.BagelOracle:
    d, sd, vsd = tiny_weights(seed, vae)
    return omodel.BagelOracle(sd, oracle_dims(d), sem, exact, vsd)


# ----------------------------------------------------------------------------- bf16 metrics
def bf16_ordinal(t: torch.Tensor) -> torch.Tensor:
    """Map bf16 values to integers that are monotone in the value (so |a-b| counts ulps)."""
    i = t.to(torch.bfloat16).contiguous().view(torch.int16).to(torch.int32)
    return torch.where(i < 0, -(i & 0x7FFF), i)


def ulp_stats(a: torch.Tensor, b: torch.Tensor) -> dict:
    """a, b: same shape; compared as bf16.  Returns mismatch fraction, max ulp distance, rel-L2."""
    a, b = a.detach().cpu(), b.detach().cpu()
    da = (bf16_ordinal(a) - bf16_ordinal(b)).abs()
    af, bf = a.float(), b.float()
    denom = bf.norm().item() or 1.0
    return dict(frac=(da > 0).float().mean().item(), max_ulp=int(da.max().item()) if da.numel() else 0,
                frac_gt1=(da > 1).float().mean().item(), rel_l2=((af - bf).norm().item() / denom),
                max_abs=(af - bf).abs().max().item() if da.numel() else 0.0)


def assert_close_bf16(a, b, max_frac=0.0, max_ulp=0, name="", rel_l2=None):
    s = ulp_stats(a, b)
    ok = s["frac"] <= max_frac and s["max_ulp"] <= max_ulp
    if rel_l2 is not None:
        ok = ok and s["rel_l2"] <= rel_l2
    assert ok, f"{name}: {s} (allowed frac<={max_frac}, ulp<={max_ulp}, rel_l2<={rel_l2})"
    return s
