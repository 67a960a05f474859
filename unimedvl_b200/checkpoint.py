"""Checkpoint ingest: a UniMedVL / BAGEL checkpoint directory -> engine dims + resident weights.

Replaces the reference's loading path -- ``Qwen2Config.from_json_file(llm_config.json)`` / ``SiglipVisionConfig`` with
``num_hidden_layers - 1`` (interactive_vqa_inferencer.py:206-213), the ema / model ``*.safetensors`` selection with an
optional on-disk bf16 re-save (interactive_vqa_inferencer.py:93-151), ``load_checkpoint_and_dispatch(dtype=bf16)``
(:153-156) and ``load_ae(ae.safetensors)`` (interactive_image_generator.py:222) -- by streaming the tensors one at a
time through ``umv_load_tensor``: fp32 shards are rounded to bf16 on the way to the GPU (no converted copy is written),
host memory holds one tensor at a time, and the engine packs them into its own layout (fused q|k|v, interleaved gate|up,
tap-major conv kernels).  Tensor names are the reference ``state_dict`` keys.
"""
from __future__ import annotations

import json
import os
from typing import Iterable, Optional

import torch

from .config import BagelDims, LLMDims, ViTDims

# file preference of the reference (interactive_vqa_inferencer.py:133-150): the bf16 copy first, then the original
_CANDIDATES = {False: ("ema_bf16.safetensors", "ema.safetensors"), True: ("model_bf16.safetensors", "model.safetensors")}


def dims_from_checkpoint(model_path: str) -> BagelDims:
    """llm_config.json / vit_config.json -> BagelDims (missing files: the BAGEL-7B-MoT defaults of config.py)."""
    llm, vit = LLMDims(), ViTDims()
    p = os.path.join(model_path, "llm_config.json")
    if os.path.exists(p):
        c = json.load(open(p))
        llm = LLMDims(hidden=c["hidden_size"], heads=c["num_attention_heads"], kv_heads=c.get("num_key_value_heads", c["num_attention_heads"]),
                      inter=c["intermediate_size"], layers=c["num_hidden_layers"], vocab=c["vocab_size"],
                      rope_theta=float(c.get("rope_theta", 1e6)), eps=float(c.get("rms_norm_eps", 1e-6)))
    p = os.path.join(model_path, "vit_config.json")
    if os.path.exists(p):
        c = json.load(open(p))
        vit = ViTDims(hidden=c["hidden_size"], heads=c["num_attention_heads"], inter=c["intermediate_size"],
                      layers=c["num_hidden_layers"] - 1,            # the reference drops the last encoder layer
                      patch=c.get("patch_size", 14), channels=c.get("num_channels", 3), image_size=c.get("image_size", 980),
                      eps=float(c.get("layer_norm_eps", 1e-6)))
    return BagelDims(llm=llm, vit=vit)


def find_checkpoint(model_path: str, use_model_checkpoint: bool = False) -> str:
    for name in _CANDIDATES[bool(use_model_checkpoint)]:
        p = os.path.join(model_path, name)
        if os.path.exists(p):
            return p
    raise FileNotFoundError(f"Checkpoint not found: {model_path}")          # same message as the reference (:150)


def _stream(path: str) -> Iterable[tuple[str, torch.Tensor]]:
    from safetensors import safe_open
    with safe_open(path, framework="pt", device="cpu") as f:
        for name in f.keys():
            yield name, f.get_tensor(name)


def load_checkpoint(engine, model_path: str, use_model_checkpoint: bool = False, load_vae: Optional[bool] = None,
                    strict: bool = False) -> dict:
    """Stream ``ema[_bf16].safetensors`` (and ``ae.safetensors`` when the engine was built with the generation path)
    into ``engine``.  Tensors the engine has no slot for (training-only buffers, the unused last ViT layer) are skipped
    unless ``strict``; ``engine.finalize()`` still fails loudly if a tensor the engine needs never arrived.
    Returns {"file", "tensors", "skipped", "bytes"}."""
    path = find_checkpoint(model_path, use_model_checkpoint)
    stats = {"file": path, "tensors": 0, "skipped": 0, "bytes": 0}

    def put(name: str, t: torch.Tensor):
        if t.dtype in (torch.float16, torch.float64):
            t = t.to(torch.float32)
        if t.dtype not in (torch.bfloat16, torch.float32):      # integer buffers (step counters, position ids): not weights
            if strict:
                raise ValueError(f"{name}: dtype {t.dtype} (bf16 or fp32 expected)")
            stats["skipped"] += 1
            return
        try:
            engine.load_state_dict({name: t}, strict=True)
            stats["tensors"] += 1
            stats["bytes"] += t.numel() * 2
        except Exception as e:          # unknown tensor name: skip unless strict
            if strict or "unknown tensor" not in str(e):
                raise
            stats["skipped"] += 1

    for name, t in _stream(path):
        put(name, t)
    ae = os.path.join(model_path, "ae.safetensors")
    has_vae = bool(getattr(getattr(engine, "_c_dims", None), "enable_vae", 0))
    want_vae = load_vae if load_vae is not None else (has_vae and os.path.exists(ae))
    if want_vae:
        if not os.path.exists(ae):
            raise FileNotFoundError(f"VAE checkpoint not found: {ae}")
        for name, t in _stream(ae):          # load_ae strips a DataParallel "module." prefix (autoencoder.py:352-361)
            put("vae_model." + (name[len("module."):] if name.startswith("module.") else name), t)
    return stats
