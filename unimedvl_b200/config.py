"""Plain-struct dimensions of the unified model.

The reference keeps these in HF config objects read from the checkpoint directory
(``llm_config.json`` / ``vit_config.json``, codes/interactive_vqa_inferencer.py:206-213)
plus ``BagelConfig`` (codes/modeling/unimedvl/bagel.py:30-88) and the hard-coded VAE
parameters (codes/modeling/autoencoder.py:338-349).  The engine is dims-driven; the
values below are the benchmark defaults (BAGEL-7B-MoT).
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict


@dataclass(frozen=True)
class LLMDims:
    hidden: int = 3584
    heads: int = 28
    kv_heads: int = 4
    inter: int = 18944
    layers: int = 28
    vocab: int = 152064
    rope_theta: float = 1e6
    eps: float = 1e-6

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads


@dataclass(frozen=True)
class ViTDims:
    hidden: int = 1152
    heads: int = 16
    inter: int = 4304
    layers: int = 26            # vit_config.num_hidden_layers - 1 (interactive_vqa_inferencer.py:213)
    patch: int = 14
    channels: int = 3
    image_size: int = 980       # -> 70x70 learned position table
    eps: float = 1e-6

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def patch_dim(self) -> int:
        return self.channels * self.patch * self.patch

    @property
    def num_positions(self) -> int:
        return (self.image_size // self.patch) ** 2


@dataclass(frozen=True)
class VAEDims:
    ch: int = 128
    ch_mult: tuple = (1, 2, 4, 4)
    num_res_blocks: int = 2
    z_channels: int = 16
    in_channels: int = 3
    out_ch: int = 3
    downsample: int = 8
    scale_factor: float = 0.3611
    shift_factor: float = 0.1159


@dataclass(frozen=True)
class BagelDims:
    llm: LLMDims = field(default_factory=LLMDims)
    vit: ViTDims = field(default_factory=ViTDims)
    vae: VAEDims = field(default_factory=VAEDims)
    latent_patch_size: int = 2
    max_latent_size: int = 64
    vit_max_num_patch_per_side: int = 70

    @property
    def latent_downsample(self) -> int:
        return self.vae.downsample * self.latent_patch_size

    @property
    def patch_latent_dim(self) -> int:
        return self.latent_patch_size ** 2 * self.vae.z_channels

    def to_dict(self) -> dict:
        return asdict(self)


def bagel_7b_mot() -> BagelDims:
    """The "14B" configuration the headline metric is quoted on."""
    return BagelDims()


def tiny(llm_layers: int = 2, vit_layers: int = 2) -> BagelDims:
    """Second-scale parity configuration.  Keeps the head dims (128 / 72), the GQA group of 7
    and the 64-multiple intermediate size so the same kernel specialisations run as at 14B."""
    return BagelDims(
        llm=LLMDims(hidden=896, heads=7, kv_heads=1, inter=1536, layers=llm_layers, vocab=2048),
        vit=ViTDims(hidden=144, heads=2, inter=328, layers=vit_layers),
    )


# Qwen2.5 special-token ids (data_utils.add_special_tokens, codes/data/data_utils.py:140-176).
QWEN25_TOKEN_IDS = dict(bos_token_id=151644, eos_token_id=151645, start_of_image=151652, end_of_image=151653)
