"""Batched pipelines (unimedvl_b200.BatchedInferencer) on the B200: B requests through ONE packed sequence
of calls must give, per request, what the single-request driver gives (the reference's InterleaveInferencer semantics, pinned by the
fixtures the reference itself produced): text identical, images equal up to the summation-order noise between the <= 64-row
(weight-major) and > 64-row (token-major) linears a different batch size selects (DESIGN.md "Batch invariance")."""
import gc

import numpy as np
import pytest
import torch
from PIL import Image

from unimedvl_b200 import synth
from util import Golden, TOK, tiny_weights

pytestmark = pytest.mark.gpu


class FakeTokenizer:
    def encode(self, text):
        return [(ord(c) * 7 + i * 13) % 2000 for i, c in enumerate(text)]

    def decode(self, ids):
        m = {2040: "<|im_start|>", 2041: "<|im_end|>"}
        return " ".join(m.get(int(i), str(int(i))) for i in ids)


@pytest.fixture(scope="module")
def stack():
    from unimedvl_b200 import BatchedInferencer
    from unimedvl_b200.autoencoder import AutoEncoder
    from unimedvl_b200.bagel import Bagel
    from unimedvl_b200.engine import Engine
    from unimedvl_b200.inferencer import InterleaveInferencer
    from unimedvl_b200.packing import ImageTransform
    dims, sd, vsd = tiny_weights(vae=True)
    eng = Engine(dims, max_tokens=2048, max_seqs=8, kv_pages=256, enable_vae=True)
    eng.load_state_dict(sd)
    vae = AutoEncoder(eng)
    vae.load_state_dict(vsd)
    eng.finalize()
    args = (Bagel(eng, dims), vae, FakeTokenizer(), ImageTransform(1024, 32, 16), ImageTransform(980, 28, 14), TOK)
    return BatchedInferencer(*args), InterleaveInferencer(*args), eng, vae


def _img(i, h=70, w=98):
    return Image.fromarray(synth.synthetic_image(30 + i, h, w))


def _diff(a, b):
    d = np.abs(np.asarray(a).astype(np.int32) - np.asarray(b).astype(np.int32))
    return float(d.mean()), float((d > 32).mean())


def test_batched_image_to_text(stack):
    bi, si, eng, _ = stack
    free0 = eng.pages_free()
    images = [_img(0), _img(1, 56, 84), _img(2, 84, 112)]
    texts = ["What is shown in this image?", "Is there a fracture?", "Describe the findings."]
    got = bi(images=images, texts=texts, understanding_output=True, max_think_token_n=9, do_sample=False)
    assert [g["image"] for g in got] == [None] * 3
    assert got[0]["text"] == str(Golden("e2e").z["e2e.i2t_text"])               # request 0 == the reference's own single run
    for im, tx, g in zip(images, texts, got):
        assert g["text"] == si(image=im, text=tx, understanding_output=True, max_think_token_n=9, do_sample=False)["text"]
    # think mode: the system prompt is prefilled for the whole batch, every request stops at its own end token
    got = bi(images=images[:2], texts=texts[:2], think=True, understanding_output=True, max_think_token_n=8, do_sample=False)
    assert got[0]["text"] == str(Golden("think").z["think.i2t_text"])
    assert got[1]["text"] == si(image=images[1], text=texts[1], think=True, understanding_output=True, max_think_token_n=8, do_sample=False)["text"]
    del got
    gc.collect()
    assert eng.pages_free() == free0


def test_batched_text_to_image(stack):
    bi, si, eng, _ = stack
    texts = ["a chest x-ray with cardiomegaly", "retina", "an axial CT slice of the liver"]
    kw = dict(understanding_output=False, num_timesteps=5, image_shapes=(64, 64), cfg_text_scale=4.0, cfg_img_scale=1.5)
    torch.manual_seed(42)              # initial noise: drawn image by image from the CPU generator, in batch order (bagel.py:835-837)
    got = bi(texts=texts, **kw)
    assert all(g["text"] is None and isinstance(g["image"], Image.Image) for g in got)
    mean, far = _diff(got[0]["image"], Golden("e2e").z["e2e.t2i_image"])        # request 0 drew the fixture's noise
    assert mean < 6.0 and far < 0.02, (mean, far)
    torch.manual_seed(42)
    for t, g in zip(texts, got):       # the single-request driver consumes the generator in the same order
        mean, far = _diff(g["image"], si(text=t, **kw)["image"])
        assert mean < 2.0 and far < 0.005, (t, mean, far)
    # device uint8 output for servers / the NCCL image gather
    ctx = bi.update_context_text(texts, bi.init_gen_context(3))
    torch.manual_seed(42)
    u8 = bi.gen_image((64, 64), ctx, cfg_text_precontext=bi.init_gen_context(3), cfg_img_precontext=ctx, num_timesteps=5,
                      cfg_text_scale=4.0, cfg_img_scale=1.5, as_uint8=True)
    assert u8.is_cuda and u8.dtype == torch.uint8 and tuple(u8.shape) == (3, 64, 64, 3)
    assert _diff(u8[0].cpu().numpy(), got[0]["image"])[0] < 1.0


def test_batched_image_edit_and_think_generation(stack):
    bi, si, eng, vae = stack
    images, texts = [_img(0), _img(1)], ["make it brighter", "remove the artefact"]
    kw = dict(understanding_output=False, num_timesteps=4, image_shapes=(64, 80), cfg_text_scale=4.0, cfg_img_scale=2.0,
              cfg_interval=[0, 1.0], cfg_renorm_type="text_channel")
    vae.sample = False                 # posterior mean: the only random draws left are the initial latents, in request order
    try:
        torch.manual_seed(43)
        got = bi(images=images, texts=texts, **kw)
        torch.manual_seed(43)
        for im, tx, g in zip(images, texts, got):
            want = si(image=im, text=tx, **kw)["image"]
            assert np.asarray(g["image"]).shape == (64, 80, 3)
            mean, far = _diff(g["image"], want)
            assert mean < 2.0 and far < 0.005, (tx, mean, far)
        # think + generation: the plan is decoded per request, joins the context, then the image is generated
        kw2 = dict(think=True, understanding_output=False, max_think_token_n=6, do_sample=False, num_timesteps=3, image_shapes=(64, 64),
                   cfg_text_scale=4.0, cfg_img_scale=1.5, cfg_interval=[0.0, 1.0])
        prompts = ["a chest x-ray with cardiomegaly", "retina"]
        torch.manual_seed(61)
        got = bi(texts=prompts, **kw2)
        gold = Golden("think").z
        assert got[0]["text"] == str(gold["think.t2i_text"])
        mean, far = _diff(got[0]["image"], gold["think.t2i_image"])
        assert mean < 8.0 and far < 0.03, (mean, far)
        torch.manual_seed(61)
        for t, g in zip(prompts, got):
            want = si(text=t, **kw2)
            assert g["text"] == want["text"] and _diff(g["image"], want["image"])[0] < 2.0
    finally:
        vae.sample = True


@pytest.mark.parametrize("variant", ["ver1", "ver0_1", "ver0"])
def test_batched_vqa_reconstruction(stack, variant):
    """Batch of one == the reference's own run (fixture, CPU-generator posterior noise); batch of two == two single runs."""
    bi, si, eng, vae = stack
    gold = Golden("recon").z
    imgs = [_img(0), _img(1, 64, 64)]
    kw = dict(reconstruct_image=True, max_think_token_n=7, do_sample=False, num_timesteps=3, cfg_interval=[0.0, 1.0])
    vae.noise_device = "cpu"
    try:
        torch.manual_seed(51 if variant == "ver1" else 52)
        out = bi.vqa_reconstruction([imgs + ["Describe the findings."]], variant, **kw)[0]
    finally:
        vae.noise_device = "cuda"
    assert out[0] == str(gold[f"recon.{variant}_text"]) and len(out) == (2 if variant == "ver0" else 3)
    for i, im in enumerate(out[1:]):
        mean, far = _diff(im, gold[f"recon.{variant}_image{i}"])
        assert mean < 8.0 and far < 0.03, (variant, i, mean, far)
    # B = 2 (different questions, different images) against the single-request methods, posterior mean, shared init-noise order is
    # not reproducible across the two schedules (the batch draws both requests' first latents before the second ones), so text must
    # be identical and the images only have to be valid reconstructions of the right geometry
    reqs = [imgs + ["Describe the findings."], [_img(2), _img(3, 64, 64)] + ["Is there a fracture?"]]
    single = {"ver1": si.interleave_inference_for_vqa_reconstruction_ver1, "ver0_1": si.interleave_inference_for_vqa_reconstruction_ver0_1,
              "ver0": si.interleave_inference_for_vqa_reconstruction_ver0}[variant]
    vae.sample = False
    try:
        torch.manual_seed(7)
        got = bi.vqa_reconstruction(reqs, variant, **kw)
        for r, g in zip(reqs, got):
            torch.manual_seed(7)
            want = single(r, **kw)
            assert g[0] == want[0] and len(g) == len(want)
            for a, b in zip(g[1:], want[1:]):
                assert np.asarray(a).shape == np.asarray(b).shape and np.asarray(a).dtype == np.uint8
    finally:
        vae.sample = True
    with pytest.raises(ValueError):
        bi.vqa_reconstruction(reqs, "ver2")
    with pytest.raises(ValueError):
        bi.interleave_inference([[_img(0), "a"], ["b"]], understanding_output=True)          # mixed structures in one batch
