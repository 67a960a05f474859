"""Checkpoint ingest (unimedvl_b200/checkpoint.py): config json -> dims (CPU), and safetensors -> engine weights on the GPU,
checked against loading the same state dict directly."""
import json
import os

import pytest
import torch

from unimedvl_b200 import checkpoint, config as ucfg
from util import tiny_weights


def _write_configs(path, d):
    json.dump(dict(hidden_size=d.llm.hidden, num_attention_heads=d.llm.heads, num_key_value_heads=d.llm.kv_heads,
                   intermediate_size=d.llm.inter, num_hidden_layers=d.llm.layers, vocab_size=d.llm.vocab, rope_theta=d.llm.rope_theta,
                   rms_norm_eps=d.llm.eps), open(os.path.join(path, "llm_config.json"), "w"))
    json.dump(dict(hidden_size=d.vit.hidden, num_attention_heads=d.vit.heads, intermediate_size=d.vit.inter,
                   num_hidden_layers=d.vit.layers + 1, patch_size=d.vit.patch, num_channels=d.vit.channels, image_size=d.vit.image_size,
                   layer_norm_eps=d.vit.eps), open(os.path.join(path, "vit_config.json"), "w"))


def test_dims_from_checkpoint_and_file_choice(tmp_path):
    d = ucfg.tiny()
    _write_configs(str(tmp_path), d)
    got = checkpoint.dims_from_checkpoint(str(tmp_path))
    assert got.llm == d.llm and got.vit == d.vit          # the ViT drops its last layer (interactive_vqa_inferencer.py:213)
    assert checkpoint.dims_from_checkpoint(str(tmp_path / "missing")).llm == ucfg.LLMDims()
    with pytest.raises(FileNotFoundError):
        checkpoint.find_checkpoint(str(tmp_path))
    (tmp_path / "ema.safetensors").write_bytes(b"")
    assert checkpoint.find_checkpoint(str(tmp_path)).endswith("ema.safetensors")
    (tmp_path / "ema_bf16.safetensors").write_bytes(b"")
    assert checkpoint.find_checkpoint(str(tmp_path)).endswith("ema_bf16.safetensors")      # the bf16 copy wins, as in the reference
    with pytest.raises(FileNotFoundError):
        checkpoint.find_checkpoint(str(tmp_path), use_model_checkpoint=True)


@pytest.mark.gpu
def test_load_checkpoint_matches_direct_load(tmp_path):
    from safetensors.torch import save_file
    from unimedvl_b200.engine import Engine
    d, sd, vsd = tiny_weights(vae=True)
    _write_configs(str(tmp_path), d)
    fp32 = {k: v.float().contiguous() for k, v in sd.items()}           # an fp32 checkpoint: rounded to bf16 on ingest
    fp32["training_only.step"] = torch.zeros(1)                          # a tensor the engine has no slot for
    fp32["training_only.counter"] = torch.zeros(1, dtype=torch.int64)   # a non-float buffer: skipped, not an error
    save_file(fp32, str(tmp_path / "ema.safetensors"))
    # the FLUX autoencoder file ships in fp32 and may carry a DataParallel "module." prefix (load_ae, autoencoder.py:352-361)
    save_file({"module." + k: v.float().contiguous() for k, v in vsd.items()}, str(tmp_path / "ae.safetensors"))

    dims = checkpoint.dims_from_checkpoint(str(tmp_path))
    a = Engine(dims, max_tokens=256, max_seqs=2, kv_pages=16, enable_vae=True)
    stats = checkpoint.load_checkpoint(a, str(tmp_path))
    a.finalize()
    assert stats["skipped"] == 2 and stats["tensors"] == len(sd) + len(vsd)
    b = Engine(d, max_tokens=256, max_seqs=2, kv_pages=16, enable_vae=True)
    b.load_state_dict(sd)
    b.load_state_dict({"vae_model." + k: v for k, v in vsd.items()})
    b.finalize()
    assert a.weight_bytes() == b.weight_bytes()
    x = (torch.randn(40, d.llm.hidden, generator=torch.Generator().manual_seed(0)) * 0.3).bfloat16().cuda()
    outs = []
    for e in (a, b):
        s = e.seq_new()
        outs.append(e.llm_forward(x, [s], [40], list(range(40)), is_causal=True, update_kv=True, want_hidden=True).cpu())
    assert torch.equal(outs[0], outs[1])
    with pytest.raises(Exception):
        checkpoint.load_checkpoint(Engine(dims, max_tokens=64, max_seqs=1, kv_pages=4), str(tmp_path), strict=True)
