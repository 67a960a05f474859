"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE.

Run in the build container only (needs /root/reference; the GPU box has no copy):

    python tests/golden/make_golden.py [packing|vqa|t2i|edit|e2e|recon|think ...]

The reference ships no tests or golden vectors, so parity is pinned on the
outputs of the unmodified reference modules imported from /root/reference/codes, executed on CPU
(bf16 parameters, torch.autocast("cpu", bfloat16)) on the deterministic synthetic weights of
unimedvl_b200.synth at the `tiny` dims.  Compatibility shims (none touches arithmetic):
  S1 Qwen2Config(pad_token_id=None)            (transformers>=5 dropped the default)
  S2 ROPE_INIT_FUNCTIONS['default']            (transformers>=5 dropped the key)
  S3 flash_attn_varlen_func -> per-sample fp32 SDPA, autocast off (flash-attn is CUDA-only)
  S4 parameters cast to bf16 one by one, never model.to(bf16) (keeps RoPE inv_freq fp32)
Only inputs and outputs are stored; weights are regenerated from the recipe by the tests.
"""
import os
import sys

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("UMV_REFERENCE", "/root/reference/codes")
sys.path.insert(0, REF)

from unimedvl_b200 import config as ucfg, synth  # noqa: E402

# ---- shims ------------------------------------------------------------------------------------
from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS  # noqa: E402


def _default_rope(config, device=None, seq_len=None, **kw):
    dh = config.hidden_size // config.num_attention_heads
    inv = 1.0 / (config.rope_theta ** (torch.arange(0, dh, 2, dtype=torch.int64).float().to(device) / dh))
    return inv, 1.0


ROPE_INIT_FUNCTIONS.setdefault("default", _default_rope)


def sdpa_varlen(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, causal=False, **kw):
    g, outs = q.shape[1] // k.shape[1], []
    for i in range(len(cu_seqlens_q) - 1):
        qs, qe, ks, ke = map(int, (cu_seqlens_q[i], cu_seqlens_q[i + 1], cu_seqlens_k[i], cu_seqlens_k[i + 1]))
        qi = q[qs:qe].transpose(0, 1).float()
        ki = k[ks:ke].transpose(0, 1).repeat_interleave(g, 0).float()
        vi = v[ks:ke].transpose(0, 1).repeat_interleave(g, 0).float()
        m = torch.ones(qe - qs, ke - ks, dtype=torch.bool).tril((ke - ks) - (qe - qs)) if causal else None
        with torch.autocast("cpu", enabled=False):       # keep the stand-in in fp32 (SDPA is an autocast op)
            o = torch.nn.functional.scaled_dot_product_attention(qi[None], ki[None], vi[None], attn_mask=m)[0]
        outs.append(o.transpose(0, 1).to(q.dtype))
    return torch.cat(outs, 0)


import modeling.unimedvl.qwen2_navit as qn  # noqa: E402
import modeling.unimedvl.siglip_navit as sn  # noqa: E402

qn.flash_attn_varlen_func = sn.flash_attn_varlen_func = sdpa_varlen

from modeling.unimedvl.bagel import Bagel, BagelConfig  # noqa: E402
from modeling.unimedvl.qwen2_navit import Qwen2Config, Qwen2ForCausalLM, NaiveCache  # noqa: E402
from modeling.unimedvl.siglip_navit import SiglipVisionConfig, SiglipVisionModel  # noqa: E402
from modeling.autoencoder import load_ae  # noqa: E402
from data.transforms import ImageTransform  # noqa: E402
from inferencer import InterleaveInferencer  # noqa: E402

TOK = dict(bos_token_id=2040, eos_token_id=2041, start_of_image=2042, end_of_image=2043)


class FakeTokenizer:
    """encode: deterministic ids from characters; decode: space-joined ids (no vocab files offline)."""
    def encode(self, text):
        return [(ord(c) * 7 + i * 13) % 2000 for i, c in enumerate(text)]

    def decode(self, ids):
        m = {2040: "<|im_start|>", 2041: "<|im_end|>"}
        return " ".join(m.get(int(i), str(int(i))) for i in ids)


def build_reference(dims: ucfg.BagelDims, seed=0):
    l, v = dims.llm, dims.vit
    llm_cfg = Qwen2Config(vocab_size=l.vocab, hidden_size=l.hidden, intermediate_size=l.inter,
                          num_hidden_layers=l.layers, num_attention_heads=l.heads, num_key_value_heads=l.kv_heads,
                          rope_theta=l.rope_theta, rms_norm_eps=l.eps, max_position_embeddings=32768,
                          qk_norm=True, layer_module="Qwen2MoTDecoderLayer", tie_word_embeddings=False,
                          pad_token_id=None)
    vit_cfg = SiglipVisionConfig(hidden_size=v.hidden, intermediate_size=v.inter, num_hidden_layers=v.layers,
                                 num_attention_heads=v.heads, image_size=v.image_size, patch_size=v.patch, rope=False)
    vae, vae_cfg = load_ae(local_path=None)
    cfg = BagelConfig(visual_gen=True, visual_und=True, llm_config=llm_cfg, vit_config=vit_cfg, vae_config=vae_cfg,
                      vit_max_num_patch_per_side=dims.vit_max_num_patch_per_side, connector_act="gelu_pytorch_tanh",
                      latent_patch_size=dims.latent_patch_size, max_latent_size=dims.max_latent_size)
    model = Bagel(Qwen2ForCausalLM(llm_cfg), SiglipVisionModel(vit_cfg), cfg, vae_model=vae)
    model.vit_model.vision_model.embeddings.convert_conv2d_to_linear(vit_cfg)
    sd = synth.bagel_state_dict(dims, seed)
    vsd = synth.vae_state_dict(dims.vae, seed)
    own = {k: t for k, t in model.state_dict().items() if not k.startswith("vae_model.")}
    missing = set(own) - set(sd)
    extra = set(sd) - set(own)
    assert not missing and not extra, (sorted(missing)[:5], sorted(extra)[:5])
    for k, t in own.items():
        assert tuple(t.shape) == tuple(sd[k].shape), (k, t.shape, sd[k].shape)
    model.load_state_dict({**sd, **{"vae_model." + k: t for k, t in vsd.items()}}, strict=True)
    assert set(vae.state_dict()) == set(vsd)
    for p in model.parameters():                      # S4
        p.data = p.data.to(torch.bfloat16)
    for p in vae.parameters():
        p.data = p.data.to(torch.bfloat16)
    assert model.language_model.model.rotary_emb.inv_freq.dtype == torch.float32
    return model.eval(), vae.eval()


def npy(t):
    """bf16 tensors are stored as their uint16 bit patterns (exact); everything else as-is."""
    if torch.is_tensor(t):
        # .copy(): the reference mutates some inputs in place (bagel.py:1272-1274 shifts
        # packed_key_value_indexes through views), the fixture must hold the value at call time.
        if t.dtype == torch.bfloat16:
            return t.contiguous().view(torch.uint16).numpy().copy()
        return t.detach().numpy().copy()
    return np.asarray(t).copy()


def dump_dict(prefix, d, out):
    for k, v in d.items():
        if torch.is_tensor(v):
            out[f"{prefix}.{k}"] = npy(v)
            out[f"{prefix}.{k}.dtype"] = str(v.dtype)
        else:
            out[f"{prefix}.{k}"] = np.asarray(v)


def cache_arrays(prefix, cache, out):
    for i in range(cache.num_layers):
        out[f"{prefix}.k{i}"] = npy(cache.key_cache[i])
        out[f"{prefix}.v{i}"] = npy(cache.value_cache[i])


def make_images(sizes, base=0):
    return [Image.fromarray(synth.synthetic_image(base + i, h, w)) for i, (h, w) in enumerate(sizes)]


def golden_packing(model, out):
    """G1: host packing dicts (integer / fp32, bit-exact contract)."""
    tok = FakeTokenizer()
    vit_tf = ImageTransform(980, 28, 14)
    vae_tf = ImageTransform(1024, 32, 16)
    imgs = make_images([(84, 112), (56, 70), (100, 61)])
    g, lens, rope = model.prepare_vit_images([0, 5, 9], [0, 5, 3], imgs, vit_tf, TOK)
    dump_dict("vit", g, out); out["vit.newlens"] = np.asarray(lens); out["vit.new_rope"] = np.asarray(rope)
    g, lens2, rope2 = model.prepare_prompts(lens, rope, ["What is shown?", "Describe.", "x"], tok, TOK)
    dump_dict("prompts", g, out); out["prompts.newlens"] = np.asarray(lens2); out["prompts.new_rope"] = np.asarray(rope2)
    g = model.prepare_start_tokens(lens2, rope2, TOK)
    dump_dict("start", g, out)
    g, lens3, rope3 = model.prepare_vae_images(lens2, rope2, imgs, vae_tf, TOK)
    dump_dict("vaeimg", g, out); out["vaeimg.newlens"] = np.asarray(lens3); out["vaeimg.new_rope"] = np.asarray(rope3)
    torch.manual_seed(42)
    g = model.prepare_vae_latent(lens2, rope2, [(64, 64), (64, 96), (32, 48)], TOK)
    dump_dict("latent", g, out)
    g = model.prepare_vae_latent_cfg([3, 0, 7], [3, 0, 2], [(64, 64), (64, 96), (32, 48)])
    dump_dict("latentcfg", g, out)
    # ImageTransform on a non-trivial resize
    t = ImageTransform(980, 378, 14)(Image.fromarray(synth.synthetic_image(7, 300, 500)))
    out["transform.300x500"] = npy(t)
    t = ImageTransform(1024, 32, 16)(Image.fromarray(synth.synthetic_image(8, 50, 70)))
    out["transform.vae.50x70"] = npy(t)


def golden_vqa(model, out):
    """G2: understanding path, B=2 ragged, greedy decode with per-step logits."""
    tok = FakeTokenizer()
    vit_tf = ImageTransform(980, 28, 14)
    imgs = make_images([(84, 112), (56, 70)], base=10)
    prompts = ["What abnormality is visible in this chest radiograph?", "Is there a fracture?"]
    taps = {}
    h1 = model.connector.register_forward_hook(lambda m, i, o: taps.setdefault("connector_out", o.detach().clone()))
    h2 = model.vit_model.register_forward_hook(lambda m, i, o: taps.setdefault("vit_out", o.detach().clone()))
    logits = []
    h3 = model.language_model.lm_head.register_forward_hook(lambda m, i, o: logits.append(o.detach().clone()))
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        cache = NaiveCache(model.config.llm_config.num_hidden_layers)
        g, lens, rope = model.prepare_vit_images([0, 0], [0, 0], imgs, vit_tf, TOK)
        dump_dict("vqa.vit_in", g, out)
        cache = model.forward_cache_update_vit(cache, **g)
        cache_arrays("vqa.after_vit", cache, out)
        g, lens, rope = model.prepare_prompts(lens, rope, prompts, tok, TOK)
        dump_dict("vqa.text_in", g, out)
        cache = model.forward_cache_update_text(cache, **g)
        cache_arrays("vqa.after_text", cache, out)
        g = model.prepare_start_tokens(lens, rope, TOK)
        dump_dict("vqa.start", g, out)
        toks = model.generate_text(past_key_values=cache, max_length=9, do_sample=False, end_token_id=None, **g)
    for h in (h1, h2, h3):
        h.remove()
    out["vqa.vit_out"] = npy(taps["vit_out"]); out["vqa.vit_out.dtype"] = str(taps["vit_out"].dtype)
    out["vqa.connector_out"] = npy(taps["connector_out"])
    out["vqa.tokens"] = npy(toks)
    out["vqa.logits"] = npy(torch.stack(logits, 0))
    out["vqa.kvlens"] = np.asarray(lens); out["vqa.ropes"] = np.asarray(rope)


def golden_t2i(model, vae, out):
    """G3: generation path, B=2 ragged images, dual CFG, all three renorm types, 4 Euler steps;
    G4: VAE decode to uint8.  Contexts are built once and shared (generate_image never updates them)."""
    tok = FakeTokenizer()
    shapes = [(64, 64), (64, 96)]
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        # main context: two text prompts; cfg_text context: empty; cfg_img context: same text (no image)
        cache = NaiveCache(model.config.llm_config.num_hidden_layers)
        g, lens, rope = model.prepare_prompts([0, 0], [0, 0], ["a chest x-ray with cardiomegaly", "retina"], tok, TOK)
        dump_dict("t2i.text_in", g, out)
        cache = model.forward_cache_update_text(cache, **g)
        cfg_cache = NaiveCache(model.config.llm_config.num_hidden_layers)
        g2, lens_c, rope_c = model.prepare_prompts([0, 0], [0, 0], ["x", "yy"], tok, TOK)
        dump_dict("t2i.cfg_text_in", g2, out)
        cfg_cache = model.forward_cache_update_text(cfg_cache, **g2)
        torch.manual_seed(42)
        gi = model.prepare_vae_latent(lens, rope, shapes, TOK)
        dump_dict("t2i.latent_in", gi, out)
        gc_text = model.prepare_vae_latent_cfg(lens_c, rope_c, shapes)
        gc_img = model.prepare_vae_latent_cfg(lens, rope, shapes)
        dump_dict("t2i.cfg_text", gc_text, out)
        dump_dict("t2i.cfg_img", gc_img, out)
        out["t2i.kvlens"] = np.asarray(lens); out["t2i.ropes"] = np.asarray(rope)
        out["t2i.cfg_kvlens"] = np.asarray(lens_c); out["t2i.cfg_ropes"] = np.asarray(rope_c)
        for renorm in ("global", "channel", "text_channel"):
            trace = []
            orig = model._forward_flow

            def spy(*a, **k):
                v = orig(*a, **k)
                trace.append(v.detach().clone().float())
                return v
            model._forward_flow = spy
            lat = model.generate_image(
                past_key_values=cache, cfg_text_past_key_values=cfg_cache, cfg_img_past_key_values=cache,
                num_timesteps=5, timestep_shift=3.0, cfg_text_scale=4.0, cfg_img_scale=1.5, cfg_interval=(0.4, 1.0),
                cfg_renorm_min=0.0, cfg_renorm_type=renorm, **gi,
                cfg_text_packed_position_ids=gc_text["cfg_packed_position_ids"],
                cfg_text_packed_query_indexes=gc_text["cfg_packed_query_indexes"],
                cfg_text_key_values_lens=gc_text["cfg_key_values_lens"],
                cfg_text_packed_key_value_indexes=gc_text["cfg_packed_key_value_indexes"],
                cfg_img_packed_position_ids=gc_img["cfg_packed_position_ids"],
                cfg_img_packed_query_indexes=gc_img["cfg_packed_query_indexes"],
                cfg_img_key_values_lens=gc_img["cfg_key_values_lens"],
                cfg_img_packed_key_value_indexes=gc_img["cfg_packed_key_value_indexes"])
            model._forward_flow = orig
            out[f"t2i.{renorm}.v_trace"] = npy(torch.stack(trace, 0))
            for i, l in enumerate(lat):
                out[f"t2i.{renorm}.latent{i}"] = npy(l)
        # G4: decode the first "global" latent through InterleaveInferencer.decode_image
        inf = InterleaveInferencer(model, vae, tok, ImageTransform(1024, 32, 16), ImageTransform(980, 28, 14), TOK)
        lat0 = torch.from_numpy(out["t2i.global.latent0"])
        img = inf.decode_image(lat0, shapes[0])
        out["vae.decode_uint8"] = np.asarray(img)
        z = torch.einsum("nhwpqc->nchpwq", lat0.reshape(1, 4, 4, 2, 2, 16)).reshape(1, 16, 8, 8).to(torch.bfloat16)
        out["vae.decode_in"] = npy(z)
        out["vae.decode_out"] = npy(vae.decode(z))


def golden_edit(model, vae, out):
    """G5: image-conditioned context (VAE encode -> gen-mode prefill) with the encoder noise pinned."""
    tok = FakeTokenizer()
    vae_tf = ImageTransform(1024, 32, 16)
    imgs = make_images([(64, 80)], base=20)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        cache = NaiveCache(model.config.llm_config.num_hidden_layers)
        g, lens, rope = model.prepare_vae_images([0], [0], imgs, vae_tf, TOK)
        dump_dict("edit.vae_in", g, out)
        out["edit.vae_in.shapes"] = np.asarray(g["patchified_vae_latent_shapes"])
        moments = vae.encoder(g["padded_images"])
        out["edit.moments"] = npy(moments)
        torch.manual_seed(7)
        noise = torch.randn_like(torch.chunk(moments, 2, dim=1)[0])
        out["edit.noise"] = npy(noise)
        torch.manual_seed(7)
        cache = model.forward_cache_update_vae(vae, cache, **g)
        cache_arrays("edit.after_vae", cache, out)


def golden_e2e(model, vae, out):
    """G6: the caller-level workflows through the reference's own InterleaveInferencer.__call__ (inferencer.py:640-680)."""
    tok = FakeTokenizer()
    inf = InterleaveInferencer(model, vae, tok, ImageTransform(1024, 32, 16), ImageTransform(980, 28, 14), TOK)
    img = make_images([(70, 98)], base=30)[0]
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        r = inf(image=img, text="What is shown in this image?", understanding_output=True, max_think_token_n=9, do_sample=False)
        out["e2e.i2t_text"] = np.asarray(r["text"])
        torch.manual_seed(42)
        r = inf(text="a chest x-ray with cardiomegaly", understanding_output=False, num_timesteps=5, image_shapes=(64, 64),
                cfg_text_scale=4.0, cfg_img_scale=1.5)
        out["e2e.t2i_image"] = np.asarray(r["image"])
        torch.manual_seed(43)
        r = inf(image=img, text="make it brighter", understanding_output=False, num_timesteps=4, image_shapes=(64, 80),
                cfg_text_scale=4.0, cfg_img_scale=2.0, cfg_interval=[0, 1.0], cfg_renorm_type="text_channel")
        out["e2e.edit_image"] = np.asarray(r["image"])


def golden_recon(model, vae, out):
    """G7: the VQA + reconstruction workflows (inferencer.py:282-549) through the reference's own methods; `__call__` with
    inference_ver=1 for ver1.  On CPU the VAE posterior noise (autoencoder.py:270) and the initial latent noise
    (bagel.py:835-837) both come from the default CPU generator, so a seed pins the whole run."""
    tok = FakeTokenizer()
    inf = InterleaveInferencer(model, vae, tok, ImageTransform(1024, 32, 16), ImageTransform(980, 28, 14), TOK)
    imgs = make_images([(70, 98), (64, 64)], base=30)
    sizes = [(98, 70), (64, 64), (2000, 1500), (3000, 500), (20, 400), (1024, 1024), (17, 17), (1500, 1499), (4096, 4096)]
    out["recon.sizes_in"] = np.asarray(sizes)
    out["recon.sizes_out"] = np.asarray([inf._calculate_target_size_with_aspect_ratio(w, h) for w, h in sizes])
    kw = dict(reconstruct_image=True, max_think_token_n=7, do_sample=False, num_timesteps=3, cfg_interval=[0.0, 1.0])
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        torch.manual_seed(51)
        r = inf(image=imgs, text="Describe the findings.", inference_ver=1, **kw)
        out["recon.ver1_text"] = np.asarray(r["text"])
        for i, im in enumerate(r["image"]):
            out[f"recon.ver1_image{i}"] = np.asarray(im)
        torch.manual_seed(52)
        r = inf.interleave_inference_for_vqa_reconstruction_ver0_1(imgs + ["Describe the findings."], **kw)
        out["recon.ver0_1_text"] = np.asarray(r[0])
        for i, im in enumerate(r[1:]):
            out[f"recon.ver0_1_image{i}"] = np.asarray(im)
        torch.manual_seed(52)          # same draws as ver0_1 up to its first image
        r = inf.interleave_inference_for_vqa_reconstruction_ver0(imgs + ["Describe the findings."], **kw)
        out["recon.ver0_text"] = np.asarray(r[0])
        assert len(r) == 2
        out["recon.ver0_image0"] = np.asarray(r[1])
        r = inf(image=imgs[0], text="Describe the findings.", inference_ver=1, max_think_token_n=7, do_sample=False)
        assert r["image"] is None
        out["recon.ver1_noimage_text"] = np.asarray(r["text"])


def golden_think(model, vae, out):
    """G8: think mode of interleave_inference (inferencer.py:574-577,612-620): the system prompt is prefilled first; for
    generation the model first writes its plan (gen_text), the plan joins the context, then the image is generated."""
    tok = FakeTokenizer()
    inf = InterleaveInferencer(model, vae, tok, ImageTransform(1024, 32, 16), ImageTransform(980, 28, 14), TOK)
    img = make_images([(70, 98)], base=30)[0]
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        r = inf(image=img, text="What is shown in this image?", think=True, understanding_output=True, max_think_token_n=8,
                do_sample=False)
        out["think.i2t_text"] = np.asarray(r["text"])
        torch.manual_seed(61)
        r = inf(text="a chest x-ray with cardiomegaly", think=True, understanding_output=False, max_think_token_n=6, do_sample=False,
                num_timesteps=3, image_shapes=(64, 64), cfg_text_scale=4.0, cfg_img_scale=1.5, cfg_interval=[0.0, 1.0])
        out["think.t2i_text"] = np.asarray(r["text"])
        out["think.t2i_image"] = np.asarray(r["image"])


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    dims = ucfg.tiny()
    model, vae = build_reference(dims)
    for name, fn in (("packing", lambda o: golden_packing(model, o)),
                     ("vqa", lambda o: golden_vqa(model, o)),
                     ("t2i", lambda o: golden_t2i(model, vae, o)),
                     ("edit", lambda o: golden_edit(model, vae, o)),
                     ("e2e", lambda o: golden_e2e(model, vae, o)),
                     ("recon", lambda o: golden_recon(model, vae, o)),
                     ("think", lambda o: golden_think(model, vae, o))):
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        out = {}
        fn(out)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(name, len(out), "arrays", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
