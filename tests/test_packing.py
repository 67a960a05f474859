"""Host packing vs fixtures produced by the reference's Bagel.prepare_* (tests/golden/make_golden.py, G1).
Integer / fp32 host work -> bit-exact."""
import numpy as np
import pytest
import torch
from PIL import Image

from unimedvl_b200 import packing, synth
from util import Golden, TOK


class FakeTokenizer:
    def encode(self, text):
        return [(ord(c) * 7 + i * 13) % 2000 for i, c in enumerate(text)]


def _imgs(sizes, base=0):
    return [Image.fromarray(synth.synthetic_image(base + i, h, w)) for i, (h, w) in enumerate(sizes)]


def _cmp(got: dict, g: Golden, prefix: str):
    want = g.group(prefix)
    for k, w in want.items():
        if k in ("newlens", "new_rope"):
            continue
        v = got[k]
        if not torch.is_tensor(v):
            v = torch.as_tensor(np.asarray(v))
        assert v.dtype == w.dtype, (prefix, k, v.dtype, w.dtype)
        assert v.shape == w.shape, (prefix, k, v.shape, w.shape)
        assert torch.equal(v, w), (prefix, k)
    assert set(got) >= set(k for k in want if k not in ("newlens", "new_rope"))


def test_packing_matches_reference():
    g = Golden("packing")
    tok = FakeTokenizer()
    vit_tf = packing.ImageTransform(980, 28, 14)
    vae_tf = packing.ImageTransform(1024, 32, 16)
    imgs = _imgs([(84, 112), (56, 70), (100, 61)])
    d, lens, rope = packing.prepare_vit_images([0, 5, 9], [0, 5, 3], imgs, vit_tf, TOK)
    _cmp(d, g, "vit")
    assert lens == g.t("vit.newlens").tolist() and rope == g.t("vit.new_rope").tolist()
    d, lens2, rope2 = packing.prepare_prompts(lens, rope, ["What is shown?", "Describe.", "x"], tok, TOK)
    _cmp(d, g, "prompts")
    assert lens2 == g.t("prompts.newlens").tolist() and rope2 == g.t("prompts.new_rope").tolist()
    _cmp(packing.prepare_start_tokens(lens2, rope2, TOK), g, "start")
    d, lens3, rope3 = packing.prepare_vae_images(lens2, rope2, imgs, vae_tf, TOK)
    shapes = d.pop("patchified_vae_latent_shapes")
    assert [list(s) for s in shapes] == g.t("vaeimg.patchified_vae_latent_shapes").tolist()
    _cmp({**d, "patchified_vae_latent_shapes": np.asarray(shapes)}, g, "vaeimg")
    assert lens3 == g.t("vaeimg.newlens").tolist() and rope3 == g.t("vaeimg.new_rope").tolist()
    torch.manual_seed(42)
    _cmp(packing.prepare_vae_latent(lens2, rope2, [(64, 64), (64, 96), (32, 48)], TOK), g, "latent")
    _cmp(packing.prepare_vae_latent_cfg([3, 0, 7], [3, 0, 2], [(64, 64), (64, 96), (32, 48)]), g, "latentcfg")


def test_image_transform_matches_reference():
    g = Golden("packing")
    t = packing.ImageTransform(980, 378, 14)(Image.fromarray(synth.synthetic_image(7, 300, 500)))
    assert torch.equal(t, g.t("transform.300x500"))
    t = packing.ImageTransform(1024, 32, 16)(Image.fromarray(synth.synthetic_image(8, 50, 70)))
    assert torch.equal(t, g.t("transform.vae.50x70"))


def test_empty_and_single():
    d, lens, rope = packing.prepare_prompts([0], [0], [""], FakeTokenizer(), TOK)
    assert d["packed_text_ids"].tolist() == [TOK["bos_token_id"], TOK["eos_token_id"]] and lens == [2] and rope == [2]
    s = packing.prepare_start_tokens([], [], TOK)
    assert s["packed_start_tokens"].numel() == 0 and s["packed_key_value_indexes"].numel() == 0


@pytest.mark.gpu
def test_device_patchify_is_bit_identical_to_the_host_transforms():
    """umv_patchify_u8 (ToTensor + Normalize + patchify + position ids on the device) vs ImageTransform + patchify +
    flattened_position_ids on the host, ragged image sizes."""
    import torch
    from unimedvl_b200 import config as ucfg, packing, synth
    from unimedvl_b200.engine import Engine
    from PIL import Image
    dims = ucfg.tiny()
    eng = Engine(dims, max_tokens=64, max_seqs=1, kv_pages=4, enable_gen=False)
    tf = packing.ImageTransform(980, 14, 14, max_pixels=2_007_040)
    sizes = [(56, 84), (448, 448), (14, 14), (70, 28)]
    imgs = [torch.from_numpy(synth.synthetic_image(i, h, w).copy()) for i, (h, w) in enumerate(sizes)]
    pix, pos, lens = eng.patchify_u8([im.pin_memory() for im in imgs])
    ref_pix, ref_pos = [], []
    for im in imgs:
        t = tf(Image.fromarray(im.numpy()))
        assert tuple(t.shape[1:]) == tuple(im.shape[:2])             # these sizes pass the resize rule unchanged
        ref_pix.append(packing.patchify(t, dims.vit.patch))
        ref_pos.append(packing.flattened_position_ids(t.size(1), t.size(2), dims.vit.patch, dims.vit_max_num_patch_per_side))
    assert lens == [p.shape[0] for p in ref_pix]
    assert torch.equal(pix.cpu(), torch.cat(ref_pix, 0))
    assert torch.equal(pos.cpu(), torch.cat(ref_pos, 0))


def test_reconstruction_target_size_matches_reference():
    """InterleaveInferencer._calculate_target_size_with_aspect_ratio (inferencer.py:42-71) vs the reference's own values."""
    from unimedvl_b200.inferencer import InterleaveInferencer
    from unimedvl_b200.packing import ImageTransform
    z = Golden("recon").z
    inf = InterleaveInferencer(None, None, None, ImageTransform(1024, 32, 16), ImageTransform(980, 28, 14), {})
    got = [inf._calculate_target_size_with_aspect_ratio(int(w), int(h)) for w, h in z["recon.sizes_in"]]
    assert got == [tuple(int(v) for v in r) for r in z["recon.sizes_out"]]


def test_image_prompt_layout_matches_the_two_packers():
    """packing.image_prompt_layout (one fused prefill) lays out exactly the rows prepare_vit_images' block layout followed by
    prepare_prompts would (bagel.py:460-520, 377-409): same ids, same rope positions, same cache state afterwards."""
    n_img, prompts = [6, 20, 1], [[11, 12, 13], [], [7] * 9]

    class _Ids:
        def encode(self, i): return list(prompts[i])
    L = packing.image_prompt_layout(n_img, prompts, TOK)
    blk = packing._image_block_layout([0] * 3, [0] * 3, n_img, TOK)
    gp, lens, rope = packing.prepare_prompts(blk["seqlens"], [1] * 3, [0, 1, 2], _Ids(), TOK)
    assert L["kv_lens"] == lens and L["rope"] == rope
    assert L["seq_lens"] == [a + b for a, b in zip(blk["seqlens"], gp["text_token_lens"].tolist())]
    assert L["prompt_lens"] == gp["text_token_lens"].tolist()
    row, ids, pos, k = 0, [], [], 0
    text = dict(zip(L["text_rows"], L["text_ids"]))
    for b, n in enumerate(n_img):
        pl = L["prompt_lens"][b]
        assert [text[row], text[row + n + 1]] == [TOK["start_of_image"], TOK["end_of_image"]]
        assert L["positions"][row:row + n + 2] == [0] * (n + 2)
        assert [text[r] for r in range(row + n + 2, row + n + 2 + pl)] == gp["packed_text_ids"][k:k + pl].tolist()
        assert L["positions"][row + n + 2:row + n + 2 + pl] == gp["packed_text_position_ids"][k:k + pl].tolist()
        assert L["vit_rows"][sum(n_img[:b]):sum(n_img[:b + 1])] == list(range(row + 1, row + 1 + n))
        row += n + 2 + pl
        k += pl
    assert sorted(L["text_rows"] + L["vit_rows"]) == list(range(row))
