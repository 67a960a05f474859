"""Caller-level workflows on the B200 through unimedvl_b200.InterleaveInferencer (the reference's
InterleaveInferencer surface) vs fixtures produced by the reference's own InterleaveInferencer.__call__:
image -> text (understanding), text -> image (dual CFG, VAE decode to uint8), image + text -> image
(VAE-encode context, text_channel renorm).  Text must be identical; images agree to bf16 noise (a few grey
levels on average -- the reference on CPU vs the oracle shows the same, tests/test_oracle_golden.py)."""
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
def inferencer():
    from unimedvl_b200.autoencoder import AutoEncoder
    from unimedvl_b200.bagel import Bagel
    from unimedvl_b200.engine import Engine
    from unimedvl_b200.inferencer import InterleaveInferencer
    from unimedvl_b200.packing import ImageTransform
    dims, sd, vsd = tiny_weights(vae=True)
    eng = Engine(dims, max_tokens=1024, max_seqs=4, kv_pages=128, enable_vae=True)
    eng.load_state_dict(sd)
    vae = AutoEncoder(eng)
    vae.load_state_dict(vsd)
    eng.finalize()
    return InterleaveInferencer(Bagel(eng, dims), vae, FakeTokenizer(), ImageTransform(1024, 32, 16),
                                ImageTransform(980, 28, 14), TOK), eng


def _img():
    return Image.fromarray(synth.synthetic_image(30, 70, 98))


def _diff(a, b):
    d = np.abs(np.asarray(a).astype(np.int32) - np.asarray(b).astype(np.int32))
    return d.mean(), (d > 32).mean()


def test_image_to_text(inferencer):
    inf, eng = inferencer
    free0 = eng.pages_free()
    r = inf(image=_img(), text="What is shown in this image?", understanding_output=True, max_think_token_n=9, do_sample=False)
    assert r["image"] is None and r["text"] == str(Golden("e2e").z["e2e.i2t_text"])
    import gc
    gc.collect()
    assert eng.pages_free() == free0


def test_text_to_image(inferencer):
    inf, eng = inferencer
    torch.manual_seed(42)          # prepare_vae_latent draws the initial noise from the CPU generator (bagel.py:835-837)
    r = inf(text="a chest x-ray with cardiomegaly", understanding_output=False, num_timesteps=5, image_shapes=(64, 64),
            cfg_text_scale=4.0, cfg_img_scale=1.5)
    gold = Golden("e2e").z["e2e.t2i_image"]
    assert r["text"] is None and isinstance(r["image"], Image.Image) and np.asarray(r["image"]).shape == gold.shape
    mean, far = _diff(r["image"], gold)
    assert mean < 6.0 and far < 0.02, (mean, far)


def test_image_edit(inferencer):
    """update_context_image(vae=True, vit=True) -> generation with cfg_renorm_type="text_channel" (the shipped
    image-generator workflow, interactive_image_generator.py:365-371).  The VAE encoder samples with the device RNG
    in the reference (not reproducible), so only structure and range are compared with the fixture."""
    inf, eng = inferencer
    torch.manual_seed(43)
    r = inf(image=_img(), text="make it brighter", understanding_output=False, num_timesteps=4, image_shapes=(64, 80),
            cfg_text_scale=4.0, cfg_img_scale=2.0, cfg_interval=[0, 1.0], cfg_renorm_type="text_channel")
    gold = Golden("e2e").z["e2e.edit_image"]
    got = np.asarray(r["image"])
    assert got.shape == gold.shape == (64, 80, 3) and got.dtype == np.uint8
    assert abs(float(got.mean()) - float(gold.mean())) < 25.0


def test_think_mode(inferencer):
    """think=True (inferencer.py:574-577,612-620): system prompt first; for generation the plan is decoded, appended to
    the context and returned next to the image."""
    inf, eng = inferencer
    gold = Golden("think").z
    r = inf(image=_img(), text="What is shown in this image?", think=True, understanding_output=True, max_think_token_n=8,
            do_sample=False)
    assert r["image"] is None and r["text"] == str(gold["think.i2t_text"])
    torch.manual_seed(61)
    r = inf(text="a chest x-ray with cardiomegaly", think=True, understanding_output=False, max_think_token_n=6, do_sample=False,
            num_timesteps=3, image_shapes=(64, 64), cfg_text_scale=4.0, cfg_img_scale=1.5, cfg_interval=[0.0, 1.0])
    assert r["text"] == str(gold["think.t2i_text"])
    mean, far = _diff(r["image"], gold["think.t2i_image"])
    assert mean < 8.0 and far < 0.03, (mean, far)


def _recon_images():
    return [Image.fromarray(synth.synthetic_image(30 + i, h, w)) for i, (h, w) in enumerate([(70, 98), (64, 64)])]


_RECON_KW = dict(reconstruct_image=True, max_think_token_n=7, do_sample=False, num_timesteps=3, cfg_interval=[0.0, 1.0])


@pytest.mark.parametrize("variant", ["ver1", "ver0_1", "ver0"])
def test_vqa_reconstruction_workflows(inferencer, variant):
    """answer the question, then regenerate the input image(s) from [image, answer]
    (inferencer.py:282-549) vs the reference's own methods run on CPU.  With the VAE posterior noise drawn from the CPU
    generator (as the CPU reference run does) one seed pins every draw: same answer text, images to bf16 noise."""
    inf, eng = inferencer
    gold = Golden("recon").z
    imgs = _recon_images()
    inf.vae_model.noise_device = "cpu"
    try:
        if variant == "ver1":
            torch.manual_seed(51)
            r = inf(image=imgs, text="Describe the findings.", inference_ver=1, **_RECON_KW)
            text, images = r["text"], r["image"]
        elif variant == "ver0_1":
            torch.manual_seed(52)
            r = inf.interleave_inference_for_vqa_reconstruction_ver0_1(imgs + ["Describe the findings."], **_RECON_KW)
            text, images = r[0], r[1:]
        else:
            torch.manual_seed(52)          # same draws as ver0_1 up to its first image
            r = inf.interleave_inference_for_vqa_reconstruction_ver0(imgs + ["Describe the findings."], **_RECON_KW)
            text, images = r[0], r[1:]
    finally:
        inf.vae_model.noise_device = "cuda"
    assert text == str(gold[f"recon.{variant}_text"])
    assert len(images) == (1 if variant == "ver0" else 2)
    for i, im in enumerate(images):
        g = gold[f"recon.{variant}_image{i}"]
        assert np.asarray(im).shape == g.shape
        mean, far = _diff(im, g)
        assert mean < 8.0 and far < 0.03, (variant, i, mean, far)


def test_vqa_reconstruction_without_reconstruction(inferencer):
    inf, eng = inferencer
    free0 = eng.pages_free()
    r = inf(image=_recon_images()[0], text="Describe the findings.", inference_ver=1, max_think_token_n=7, do_sample=False)
    assert r["image"] is None and r["text"] == str(Golden("recon").z["recon.ver1_noimage_text"])
    with pytest.raises(ValueError):
        inf(text="x", inference_ver=2)
    import gc
    gc.collect()
    assert eng.pages_free() == free0


def test_unsupported_input_raises(inferencer):
    inf, _ = inferencer
    with pytest.raises(ValueError):
        inf.interleave_inference([3.14], understanding_output=True)
    assert inf() == {"image": None, "text": None}
