"""Harness that runs the UNMODIFIED reference (uni-medical/UniMedVL `codes/`) as the parity oracle.

Test / measurement infrastructure only (tests/, bench.py's reference legs); the product never imports it.  The reference tree is looked up at, in order, $UMV_REFERENCE,
`baseline/_ref/codes` (the git-ignored copy tools/install_ref.py makes, which travels to the GPU box) and
`/root/reference/codes` (build container only).

Compatibility shims, none of which touches arithmetic:
  S1  Qwen2Config(pad_token_id=None)             transformers >= 5 dropped the default (read at qwen2_navit.py:1028)
  S2  ROPE_INIT_FUNCTIONS['default']             transformers >= 5 dropped the key (modeling_qwen2.py:139-141)
  S3  flash_attn_varlen_func -> per-sample SDPA  CPU ONLY (flash-attn is CUDA-only); on the GPU the reference's real
                                                 flash_attn_varlen_func call sites run unchanged
  S4  parameters cast to bf16 one by one          never model.to(bf16): RoPE inv_freq stays fp32 as under accelerate
Construction happens on the meta device (like the reference's own `init_empty_weights`,
interactive_vqa_inferencer.py:225-229) and the synthetic state dict is attached with assign=True, so a
full-width model costs no fp32 random init.
"""
from __future__ import annotations

import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = [os.environ.get("UMV_REFERENCE"), os.path.join(ROOT, "baseline", "_ref", "codes"), "/root/reference/codes"]


def ref_path():
    for p in _CANDIDATES:
        if p and os.path.isfile(os.path.join(p, "inferencer.py")):
            return p
    return None


def _default_rope(config, device=None, seq_len=None, **kw):
    dh = config.hidden_size // config.num_attention_heads
    inv = 1.0 / (config.rope_theta ** (torch.arange(0, dh, 2, dtype=torch.int64).float().to(device) / dh))
    return inv, 1.0


def sdpa_varlen(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, causal=False, **kw):
    """S3: CPU stand-in for flash_attn_varlen_func: q [Tq,Hq,D], k/v [Tk,Hkv,D]; fp32, bottom-right causal."""
    g, outs = q.shape[1] // k.shape[1], []
    for i in range(len(cu_seqlens_q) - 1):
        qs, qe, ks, ke = map(int, (cu_seqlens_q[i], cu_seqlens_q[i + 1], cu_seqlens_k[i], cu_seqlens_k[i + 1]))
        qi = q[qs:qe].transpose(0, 1).float()
        ki = k[ks:ke].transpose(0, 1).repeat_interleave(g, 0).float()
        vi = v[ks:ke].transpose(0, 1).repeat_interleave(g, 0).float()
        m = torch.ones(qe - qs, ke - ks, dtype=torch.bool).tril((ke - ks) - (qe - qs)) if causal else None
        with torch.autocast("cpu", enabled=False):
            o = torch.nn.functional.scaled_dot_product_attention(qi[None], ki[None], vi[None], attn_mask=m)[0]
        outs.append(o.transpose(0, 1).to(q.dtype))
    return torch.cat(outs, 0)


_MODS = None


def load():
    """Import the reference's modules (once).  Returns a namespace; raises RuntimeError when no copy is present."""
    global _MODS
    if _MODS is not None:
        return _MODS
    p = ref_path()
    if p is None:
        raise RuntimeError("reference not found: run tools/install_ref.py (baseline/_ref/codes)")
    if p not in sys.path:
        sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings("ignore", category=FutureWarning)
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS
    ROPE_INIT_FUNCTIONS.setdefault("default", _default_rope)                       # S2
    import modeling.unimedvl.qwen2_navit as qn
    import modeling.unimedvl.siglip_navit as sn
    from modeling.unimedvl.bagel import Bagel, BagelConfig
    from modeling.autoencoder import load_ae
    from data.transforms import ImageTransform
    import inferencer
    _MODS = types.SimpleNamespace(path=p, qn=qn, sn=sn, Bagel=Bagel, BagelConfig=BagelConfig, load_ae=load_ae,
                                  ImageTransform=ImageTransform, InterleaveInferencer=inferencer.InterleaveInferencer,
                                  NaiveCache=qn.NaiveCache, flash=qn.flash_attn_varlen_func, inferencer=inferencer)
    return _MODS


def random_fill(device, seed: int = 0):
    """A filler for build_reference(sd=None): random-init weights of the reference architecture without a host-side RNG pass over
    14.6e9 parameters.  CUDA: torch.randn on the device.  CPU: one 32 Mi-element normal block tiled into every tensor (distinct
    memory per tensor, so a timed CPU forward streams the real number of bytes; the values repeat, which timing does not see).
    Norm weights are 1, everything else N(0, 0.02) -- activations stay O(1) (no denormals on the CPU path)."""
    device = torch.device(device)
    state = {"base": None, "g": None}

    def fill(name: str, shape):
        n = 1
        for d in shape:
            n *= int(d)
        if name.endswith("norm.weight") or "layernorm" in name or "layer_norm" in name or ".norm" in name and name.endswith("weight") and len(shape) == 1:
            return torch.ones(tuple(shape), dtype=torch.bfloat16, device=device)
        if device.type == "cuda":
            if state["g"] is None:
                state["g"] = torch.Generator(device=device).manual_seed(seed)
            return (torch.randn(tuple(shape), device=device, generator=state["g"], dtype=torch.float32) * 0.02).to(torch.bfloat16)
        if state["base"] is None:
            state["base"] = (torch.randn(1 << 25, generator=torch.Generator().manual_seed(seed)) * 0.02).to(torch.bfloat16)
        base = state["base"]
        reps = (n + base.numel() - 1) // base.numel()
        return (base.repeat(reps)[:n] if reps > 1 else base[:n].clone()).view(tuple(shape))
    return fill


def build_reference(dims, sd: dict | None, vsd: dict | None, device="cpu", fill=None):
    """The reference `Bagel` (+ AutoEncoder) exactly as interactive_image_generator.py:226-238 assembles it, at `dims`,
    holding the tensors of `sd` / `vsd` (bf16, reference state-dict names) on `device` -- or, with `sd=None`, whatever
    `fill(name, shape)` returns for every tensor of the model (random_fill above).  Returns (model, vae)."""
    R = load()
    device = torch.device(device)
    if device.type == "cpu":
        R.qn.flash_attn_varlen_func = R.sn.flash_attn_varlen_func = sdpa_varlen      # S3
    else:
        R.qn.flash_attn_varlen_func = R.sn.flash_attn_varlen_func = R.flash           # the real kernel
    l, v = dims.llm, dims.vit
    llm_cfg = R.qn.Qwen2Config(vocab_size=l.vocab, hidden_size=l.hidden, intermediate_size=l.inter,
                               num_hidden_layers=l.layers, num_attention_heads=l.heads, num_key_value_heads=l.kv_heads,
                               rope_theta=l.rope_theta, rms_norm_eps=l.eps, max_position_embeddings=32768,
                               qk_norm=True, layer_module="Qwen2MoTDecoderLayer", tie_word_embeddings=False,
                               pad_token_id=None)                                   # S1
    vit_cfg = R.sn.SiglipVisionConfig(hidden_size=v.hidden, intermediate_size=v.inter, num_hidden_layers=v.layers,
                                      num_attention_heads=v.heads, image_size=v.image_size, patch_size=v.patch, rope=False)
    with torch.device("meta"):
        vae, vae_cfg = R.load_ae(local_path=None)
        cfg = R.BagelConfig(visual_gen=True, visual_und=True, llm_config=llm_cfg, vit_config=vit_cfg, vae_config=vae_cfg,
                            vit_max_num_patch_per_side=dims.vit_max_num_patch_per_side, connector_act="gelu_pytorch_tanh",
                            latent_patch_size=dims.latent_patch_size, max_latent_size=dims.max_latent_size)
        model = R.Bagel(R.qn.Qwen2ForCausalLM(llm_cfg), R.sn.SiglipVisionModel(vit_cfg), cfg, vae_model=vae)
        model.vit_model.vision_model.embeddings.convert_conv2d_to_linear(vit_cfg, meta=True)
    if sd is None:
        assert fill is not None, "build_reference: give a state dict or a filler"
        full = {k: fill(k, t.shape) for k, t in model.state_dict().items()}
        model.load_state_dict(full, strict=True, assign=True)
        full = None
    own = {k for k in model.state_dict() if not k.startswith("vae_model.")}
    assert sd is None or own == set(sd), (sorted(own - set(sd))[:5], sorted(set(sd) - own)[:5])
    full = {k: t.to(device=device, dtype=torch.bfloat16) for k, t in sd.items()} if sd is not None else {}     # S4: parameters only
    if sd is None:
        pass
    elif vsd is not None:
        assert set(vae.state_dict()) == set(vsd)
        full.update({"vae_model." + k: t.to(device=device, dtype=torch.bfloat16) for k, t in vsd.items()})
    else:       # no VAE weights wanted: give the unused autoencoder zeros so nothing stays on the meta device
        full.update({"vae_model." + k: torch.zeros(t.shape, dtype=torch.bfloat16, device=device) for k, t in vae.state_dict().items()})
    if sd is not None:
        model.load_state_dict(full, strict=True, assign=True)
    rot = model.language_model.model.rotary_emb
    inv, _ = _default_rope(llm_cfg, device)
    rot.register_buffer("inv_freq", inv, persistent=False)                          # fp32, as under accelerate (S4)
    rot.original_inv_freq = rot.inv_freq
    for p in model.parameters():
        p.requires_grad_(False)
        assert p.device.type == device.type and p.dtype == torch.bfloat16, "parameter left on meta / wrong dtype"
    assert rot.inv_freq.dtype == torch.float32
    return model.eval(), vae.eval()


class FakeTokenizer:
    """encode: deterministic ids from characters; decode: space-joined ids (no vocab files offline).  The same class the
    golden generator uses."""

    def __init__(self, vocab_limit: int = 2000, bos: int = 2040, eos: int = 2041):
        self.n, self.m = vocab_limit, {bos: "<|im_start|>", eos: "<|im_end|>"}

    def encode(self, text):
        return [(ord(c) * 7 + i * 13) % self.n for i, c in enumerate(text)]

    def decode(self, ids):
        return " ".join(self.m.get(int(i), str(int(i))) for i in ids)


def exact_reductions(on: bool = True):
    """`torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction = not on`.  PyTorch's default (True) lets cuBLASLt reduce
    split-K partial sums in bf16; whether it does depends on the GEMM shape and the GPU, not on the model.  Measured on the B200
    (tools/exp/redprec.py, profiles/r2_reference_parity.md): with the default, 43 % of the reference's own tiny-dims patch-embedding
    outputs (one Linear, K = 588) sit >= 1 bf16 ulp away from the fp32-accumulated value; with fp32 reductions that drops to 1e-4 (1 ulp).
    The parity tests therefore run the reference with fp32 reductions (a global torch switch, no reference code touched) and record the
    default-switch figures next to them.  Returns the previous setting."""
    prev = torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction
    torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction = not on
    return not prev


def autocast(device):
    """The context the reference's entry points run under: CUDA autocast bf16 (interactive_vqa_inferencer.py:311,
    inferencer.py:651); on CPU the equivalent CPU autocast."""
    return torch.autocast(torch.device(device).type, dtype=torch.bfloat16)
