"""Pin the oracle (oracle/, CPU restatement) against fixtures produced by RUNNING THE REFERENCE
(tests/golden/make_golden.py).  Semantics.cpu + native bf16 GEMMs reproduce the reference's CPU execution:
stages that use the same torch kernels are bit-identical, the rest agree to bf16 summation-order noise
(any 1-ulp flip is amplified chaotically through the norms, so whole-model agreement is stated as a
relative L2 bound; single ops are bit-exact / <= 1 ulp, see the block-level checks)."""
import torch

from oracle import vae as ovae, vit as ovit
from oracle.llm import PackedKV
from util import Golden, Semantics, make_oracle, ulp_stats


def test_vqa_prefill_and_decode_pinned():
    g = Golden("vqa")
    o = make_oracle(Semantics.cpu, exact=False)
    gi = g.group("vqa.vit_in")
    # ViT output and connector: same kernels -> tiny deviations only from the fp32 SDPA block order
    vit_out = ovit.vit_forward(o.sd, o.dims.vit, gi["packed_vit_tokens"], gi["packed_vit_position_ids"],
                               gi["vit_token_seqlens"], Semantics.cpu, False)
    assert vit_out.dtype == g.t("vqa.vit_out").dtype == torch.bfloat16
    assert ulp_stats(vit_out, g.t("vqa.vit_out"))["rel_l2"] < 5e-3
    conn = ovit.connector(o.sd, g.t("vqa.vit_out"), False)
    assert torch.equal(conn, g.t("vqa.connector_out"))                     # bit-exact given identical input
    cache = o.forward_cache_update_vit(o.new_cache(), **gi)
    cache = o.forward_cache_update_text(cache, **g.group("vqa.text_in"))
    for li in range(o.dims.llm.layers):
        assert ulp_stats(cache.key[li], g.t(f"vqa.after_text.k{li}"))["rel_l2"] < 1e-2
        assert ulp_stats(cache.value[li], g.t(f"vqa.after_text.v{li}"))["rel_l2"] < 1e-2
    # decode from the reference's own cache: the loop logic (index shifting, positions, KV merge) is exact
    ref_cache = PackedKV(o.dims.llm.layers)
    for li in range(o.dims.llm.layers):
        ref_cache.key[li] = g.t(f"vqa.after_text.k{li}")
        ref_cache.value[li] = g.t(f"vqa.after_text.v{li}")
    st = g.group("vqa.start")
    lg = []
    toks = o.generate_text(ref_cache, st["packed_key_value_indexes"], st["key_values_lens"], st["packed_start_tokens"],
                           st["packed_query_position_ids"], 9, logits_out=lg)
    assert torch.equal(toks, g.t("vqa.tokens"))
    gl = g.t("vqa.logits")
    exact_steps = sum(int(torch.equal(lg[s], gl[s])) for s in range(9))
    assert exact_steps >= 4, exact_steps                                    # bit-identical logits on most steps
    for s in range(9):
        assert ulp_stats(lg[s], gl[s])["rel_l2"] < 6e-3


def test_fp32_contraction_mode_is_equivalent():
    """exact=True (fp32 GEMMs, the mode GPU parity tests use) stays within bf16 noise of the fixtures."""
    g = Golden("vqa")
    o = make_oracle(Semantics.cpu, exact=True)
    ref_cache = PackedKV(o.dims.llm.layers)
    for li in range(o.dims.llm.layers):
        ref_cache.key[li] = g.t(f"vqa.after_text.k{li}")
        ref_cache.value[li] = g.t(f"vqa.after_text.v{li}")
    st = g.group("vqa.start")
    lg = []
    toks = o.generate_text(ref_cache, st["packed_key_value_indexes"], st["key_values_lens"], st["packed_start_tokens"],
                           st["packed_query_position_ids"], 9, logits_out=lg, forced_tokens=g.t("vqa.tokens"))
    gl = g.t("vqa.logits")
    for s in range(9):
        assert ulp_stats(lg[s], gl[s])["rel_l2"] < 8e-3
        assert torch.equal(lg[s].float().argmax(-1), gl[s].float().argmax(-1))


def test_t2i_flow_pinned():
    g = Golden("t2i")
    o = make_oracle(Semantics.cpu, exact=False, vae=True)
    cache = o.forward_cache_update_text(o.new_cache(), **g.group("t2i.text_in"))
    cfgc = o.forward_cache_update_text(o.new_cache(), **g.group("t2i.cfg_text_in"))
    gi = g.group("t2i.latent_in")
    ct = dict(g.group("t2i.cfg_text"), cache=cfgc)
    ci = dict(g.group("t2i.cfg_img"), cache=cache)
    for renorm in ("global", "channel", "text_channel"):
        tr = []
        lat = o.generate_image(gi, cache, ct, ci, num_timesteps=5, timestep_shift=3.0, cfg_renorm_type=renorm,
                               cfg_interval=(0.4, 1.0), cfg_text_scale=4.0, cfg_img_scale=1.5, trace=tr)
        gt = g.t(f"t2i.{renorm}.v_trace")
        assert ulp_stats(tr[0], gt[0])["rel_l2"] < 2.5e-2      # CFG (scale 4) amplifies the bf16 noise of three forwards
        for i in range(2):
            ref = g.t(f"t2i.{renorm}.latent{i}")
            assert lat[i].dtype == ref.dtype == torch.float32
            assert ((lat[i] - ref).norm() / ref.norm()).item() < 3e-2


def test_vae_decode_and_edit_pinned():
    g = Golden("t2i")
    o = make_oracle(Semantics.cpu, exact=False, vae=True)
    img = ovae.decode(o.vae_sd, g.t("vae.decode_in"), o.dims.vae, Semantics.cpu)
    assert ulp_stats(img, g.t("vae.decode_out"))["rel_l2"] < 4e-2
    u8 = o.decode_image(g.t("t2i.global.latent0"), (64, 64))
    gu = g.t("vae.decode_uint8")
    assert u8.shape == gu.shape and (u8.int() - gu.int()).abs().max().item() <= 8
    ge = Golden("edit")
    gi = ge.group("edit.vae_in")
    gi["patchified_vae_latent_shapes"] = [tuple(x) for x in ge.t("edit.vae_in.shapes").tolist()]
    gi.pop("shapes", None)
    cache = o.forward_cache_update_vae(o.new_cache(), **gi, noise=ge.t("edit.noise"))
    for li in range(o.dims.llm.layers):
        assert ulp_stats(cache.key[li], ge.t(f"edit.after_vae.k{li}"))["rel_l2"] < 4e-2
