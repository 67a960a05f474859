"""The product's HOST orchestration (unimedvl_b200.InterleaveInferencer: contexts, CFG pre-contexts, think mode, the three
VQA-reconstruction workflows, __call__ dispatch; Bagel.prepare_*; packing) run on CPU over the oracle's forwards
(tests/oracle_bagel.py) against the fixtures produced by the reference's own InterleaveInferencer: the answer text must be
identical, images agree to bf16 noise.  No GPU: this is the -m "not gpu" counterpart of tests/test_e2e_gpu.py."""
import numpy as np
import pytest
import torch
from PIL import Image

import unimedvl_b200.inferencer as inferencer_mod
from oracle_bagel import OracleBagel, OracleCache, OracleVAE
from unimedvl_b200 import synth
from unimedvl_b200.packing import ImageTransform
from util import Golden, Semantics, TOK, make_oracle, tiny_weights


class FakeTokenizer:
    def encode(self, text):
        return [(ord(c) * 7 + i * 13) % 2000 for i, c in enumerate(text)]

    def decode(self, ids):
        m = {2040: "<|im_start|>", 2041: "<|im_end|>"}
        return " ".join(m.get(int(i), str(int(i))) for i in ids)


@pytest.fixture(scope="module")
def inf():
    dims, _, _ = tiny_weights(vae=True)
    o = make_oracle(Semantics.cpu, vae=True, exact=False)            # the reference's own CPU kernels / autocast behaviour
    mp = pytest.MonkeyPatch()
    mp.setattr(inferencer_mod, "NaiveCache", OracleCache)
    yield inferencer_mod.InterleaveInferencer(OracleBagel(o, dims), OracleVAE(o), FakeTokenizer(), ImageTransform(1024, 32, 16),
                                              ImageTransform(980, 28, 14), TOK)
    mp.undo()


def _img():
    return Image.fromarray(synth.synthetic_image(30, 70, 98))


def _diff(a, b):
    d = np.abs(np.asarray(a).astype(np.int32) - np.asarray(b).astype(np.int32))
    return d.mean(), (d > 32).mean()


def test_image_to_text_and_text_to_image(inf):
    gold = Golden("e2e").z
    r = inf(image=_img(), text="What is shown in this image?", understanding_output=True, max_think_token_n=9, do_sample=False)
    assert r["image"] is None and r["text"] == str(gold["e2e.i2t_text"])
    torch.manual_seed(42)
    r = inf(text="a chest x-ray with cardiomegaly", understanding_output=False, num_timesteps=5, image_shapes=(64, 64),
            cfg_text_scale=4.0, cfg_img_scale=1.5)
    mean, far = _diff(r["image"], gold["e2e.t2i_image"])
    assert r["text"] is None and mean < 4.0 and far < 0.01, (mean, far)


def test_image_edit_with_the_reference_noise_stream(inf):
    """On CPU the VAE posterior noise comes from the same generator as in the reference run, so even the edit image matches."""
    torch.manual_seed(43)
    r = inf(image=_img(), text="make it brighter", understanding_output=False, num_timesteps=4, image_shapes=(64, 80),
            cfg_text_scale=4.0, cfg_img_scale=2.0, cfg_interval=[0, 1.0], cfg_renorm_type="text_channel")
    mean, far = _diff(r["image"], Golden("e2e").z["e2e.edit_image"])
    assert mean < 4.0 and far < 0.01, (mean, far)


def test_think_mode(inf):
    gold = Golden("think").z
    r = inf(image=_img(), text="What is shown in this image?", think=True, understanding_output=True, max_think_token_n=8,
            do_sample=False)
    assert r["text"] == str(gold["think.i2t_text"])
    torch.manual_seed(61)
    r = inf(text="a chest x-ray with cardiomegaly", think=True, understanding_output=False, max_think_token_n=6, do_sample=False,
            num_timesteps=3, image_shapes=(64, 64), cfg_text_scale=4.0, cfg_img_scale=1.5, cfg_interval=[0.0, 1.0])
    assert r["text"] == str(gold["think.t2i_text"])
    mean, far = _diff(r["image"], gold["think.t2i_image"])
    assert mean < 4.0 and far < 0.01, (mean, far)


@pytest.mark.parametrize("variant", ["ver1", "ver0_1", "ver0"])
def test_vqa_reconstruction_workflows(inf, variant):
    gold = Golden("recon").z
    imgs = [Image.fromarray(synth.synthetic_image(30 + i, h, w)) for i, (h, w) in enumerate([(70, 98), (64, 64)])]
    kw = dict(reconstruct_image=True, max_think_token_n=7, do_sample=False, num_timesteps=3, cfg_interval=[0.0, 1.0])
    if variant == "ver1":
        torch.manual_seed(51)
        r = inf(image=imgs, text="Describe the findings.", inference_ver=1, **kw)
        text, images = r["text"], r["image"]
    else:
        torch.manual_seed(52)
        fn = getattr(inf, f"interleave_inference_for_vqa_reconstruction_{variant}")
        r = fn(imgs + ["Describe the findings."], **kw)
        text, images = r[0], r[1:]
    assert text == str(gold[f"recon.{variant}_text"])
    assert len(images) == (1 if variant == "ver0" else 2)
    for i, im in enumerate(images):
        mean, far = _diff(im, gold[f"recon.{variant}_image{i}"])
        assert np.asarray(im).shape == gold[f"recon.{variant}_image{i}"].shape and mean < 4.0 and far < 0.01, (variant, i, mean, far)


def test_batched_inferencer_host_logic_on_cpu():
    """unimedvl_b200.BatchedInferencer's host orchestration (a batch of ONE, over the oracle's forwards) reproduces the reference's
    fixtures exactly like the single-request driver: interleaved I2T / T2I, think mode, and reconstruction ver0_1."""
    import unimedvl_b200.batched as batched_mod
    dims, _, _ = tiny_weights(vae=True)
    o = make_oracle(Semantics.cpu, vae=True, exact=False)
    mp = pytest.MonkeyPatch()
    mp.setattr(batched_mod, "NaiveCache", OracleCache)
    try:
        bi = batched_mod.BatchedInferencer(OracleBagel(o, dims), OracleVAE(o), FakeTokenizer(), ImageTransform(1024, 32, 16),
                                           ImageTransform(980, 28, 14), TOK)
        gold = Golden("e2e").z
        r = bi(images=[_img()], texts=["What is shown in this image?"], understanding_output=True, max_think_token_n=9, do_sample=False)
        assert len(r) == 1 and r[0]["image"] is None and r[0]["text"] == str(gold["e2e.i2t_text"])
        torch.manual_seed(42)
        r = bi(texts=["a chest x-ray with cardiomegaly"], understanding_output=False, num_timesteps=5, image_shapes=(64, 64),
               cfg_text_scale=4.0, cfg_img_scale=1.5)
        mean, far = _diff(r[0]["image"], gold["e2e.t2i_image"])
        assert r[0]["text"] is None and mean < 4.0 and far < 0.01, (mean, far)
        g = Golden("recon").z
        imgs = [Image.fromarray(synth.synthetic_image(30 + i, h, w)) for i, (h, w) in enumerate([(70, 98), (64, 64)])]
        torch.manual_seed(52)
        out = bi.vqa_reconstruction([imgs + ["Describe the findings."]], "ver0_1", reconstruct_image=True, max_think_token_n=7,
                                    do_sample=False, num_timesteps=3, cfg_interval=[0.0, 1.0])[0]
        assert out[0] == str(g["recon.ver0_1_text"]) and len(out) == 3
        for i, im in enumerate(out[1:]):
            mean, far = _diff(im, g[f"recon.ver0_1_image{i}"])
            assert mean < 4.0 and far < 0.01, (i, mean, far)
        with pytest.raises(ValueError):
            bi.interleave_inference([[_img(), "a"], ["b"]], understanding_output=True)
        assert bi() == []
    finally:
        mp.undo()
