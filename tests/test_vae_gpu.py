"""VAE (FLUX-style autoencoder) on the B200 vs the oracle (CUDA semantics) and the reference-generated
fixtures; image-conditioned context (VAE encode -> generation-expert prefill) vs the oracle.

About 30 convolutions with a GroupNorm between each: a 1-ulp flip anywhere is renormalised into every channel,
so whole-image agreement between two correct bf16 implementations is a few 1e-2 relative (reference on CPU vs
oracle: 1.8e-2, tests/test_oracle_golden.py; reference on CUDA vs engine / oracle: 2.0e-2, tests/test_reference_gpu.py); block-level
agreement is much tighter and is what pins the math: test_decoder_block_by_block runs every block alone (umv_op_vae_block)."""
import pytest
import torch

from util import Golden, Semantics, make_oracle, tiny_weights, ulp_stats

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def stack():
    from unimedvl_b200.autoencoder import AutoEncoder
    from unimedvl_b200.bagel import Bagel
    from unimedvl_b200.engine import Engine
    dims, sd, vsd = tiny_weights(vae=True)
    eng = Engine(dims, max_tokens=512, max_seqs=4, kv_pages=64, enable_vae=True)
    eng.load_state_dict(sd)
    vae = AutoEncoder(eng)
    vae.load_state_dict(vsd)
    eng.finalize()
    return eng, Bagel(eng, dims), vae, make_oracle(Semantics.cuda, vae=True), dims


def test_conv_weight_roundtrip(stack):
    eng, _, _, o, _ = stack
    for k in ("decoder.conv_in.weight", "decoder.up.1.block.0.nin_shortcut.weight", "encoder.down.0.downsample.conv.weight",
              "decoder.mid.attn_1.q.weight", "decoder.norm_out.bias"):
        assert torch.equal(eng.export_tensor("vae_model." + k, o.vae_sd[k].shape), o.vae_sd[k]), k


def test_decode_blocks_and_image(stack):
    from oracle import vae as ovae
    eng, _, vae, o, _ = stack
    g = Golden("t2i")
    z = g.t("vae.decode_in")
    taps = {}
    ref = ovae.decode(o.vae_sd, z, o.dims.vae, Semantics.cuda, taps)
    img = vae.decode(z.cuda())
    assert img.shape == ref.shape == (1, 3, 64, 64) and img.dtype == torch.bfloat16
    s = ulp_stats(img, ref)
    assert s["rel_l2"] < 5e-2, s
    s = ulp_stats(img, g.t("vae.decode_out"))          # fixture from the reference (CPU autocast semantics)
    assert s["rel_l2"] < 6e-2, s
    u8 = ovae.image_to_uint8(img.cpu())
    assert (u8.int() - g.t("vae.decode_uint8").int()).abs().max().item() <= 12


def test_encoder_moments_after_decode(stack):
    """Runs right after a decode of a different shape on the same engine: guards the programmatic-dependent-launch
    ordering of the activation x activation product in the mid attention (its "weight" operand is dynamic)."""
    from oracle import vae as ovae
    eng, _, vae, o, _ = stack
    torch.manual_seed(0)
    x = (torch.rand(1, 3, 32, 48) * 2 - 1).bfloat16()
    ref = ovae.encoder(o.vae_sd, x, o.dims.vae, Semantics.cuda)
    vae.decode(Golden("t2i").t("vae.decode_in").cuda())
    got = vae.encode_moments(x.cuda())
    assert got.shape == ref.shape == (1, 32, 4, 6)
    s = ulp_stats(got, ref)
    assert s["rel_l2"] < 5e-2, s
    assert torch.equal(got, vae.encode_moments(x.cuda()))          # deterministic


def test_edit_context_matches_oracle(stack):
    """update_context_image(vae=True): prepare_vae_images -> forward_cache_update_vae with the sampling noise pinned."""
    from unimedvl_b200.cache import NaiveCache
    eng, model, vae, o, dims = stack
    ge = Golden("edit")
    gi = ge.group("edit.vae_in")
    gi["patchified_vae_latent_shapes"] = [tuple(x) for x in ge.t("edit.vae_in.shapes").tolist()]
    gi.pop("shapes", None)
    noise = ge.t("edit.noise")

    class PinnedVae:
        def encode(self, x):
            return vae.encode(x, noise=noise)
    cache = model.forward_cache_update_vae(PinnedVae(), NaiveCache(dims.llm.layers), **gi)
    oc = o.forward_cache_update_vae(o.new_cache(), **gi, noise=noise)
    for li in range(dims.llm.layers):
        assert ulp_stats(cache.key_cache[li], oc.key[li])["rel_l2"] < 6e-2
        assert ulp_stats(cache.value_cache[li], oc.value[li])["rel_l2"] < 6e-2
    assert cache._umv.lens() == [int(gi["packed_seqlens"][0])]


# ------------------------------------------------------------------------------------------------ block level
def _blocks(d):
    """(reference module path, cin, cout, kind) of every decoder block, in execution order (autoencoder.py:240-257)."""
    nres = len(d.ch_mult)
    cin = d.ch * d.ch_mult[-1]
    out = [("decoder.conv_in", d.z_channels, cin, "conv"), ("decoder.mid.block_1", cin, cin, "res"), ("decoder.mid.attn_1", cin, cin, "attn"),
           ("decoder.mid.block_2", cin, cin, "res")]
    for lvl in reversed(range(nres)):
        cout = d.ch * d.ch_mult[lvl]
        for bi in range(d.num_res_blocks + 1):
            out.append((f"decoder.up.{lvl}.block.{bi}", cin, cout, "res"))
            cin = cout
        if lvl != 0:
            out.append((f"decoder.up.{lvl}.upsample", cin, cin, "up"))
    out += [("decoder.norm_out", cin, cin, "gn"), ("decoder.conv_out", cin, d.out_ch, "conv")]
    return out


def test_decoder_block_by_block(stack):
    """Every block of the decoder ALONE (umv_op_vae_block) on the input the oracle's own forward hands it, against the oracle's block:
    conv_in, the two mid ResnetBlocks, the mid attention, the 12 up ResnetBlocks (incl. the nin_shortcut ones), the 3 nearest-2x upsample
    convolutions, norm_out + swish and conv_out.  One block is one or two GroupNorms and one to four contractions, so agreement is an
    order of magnitude tighter than for the whole image -- this is what pins the VAE math (SURVEY.md R-points: GroupNorm / swish in fp32 under
    CUDA autocast, bf16 conv outputs, residual adds in bf16)."""
    import torch.nn.functional as F
    from oracle import vae as ovae
    eng, _, vae, o, dims = stack
    d, sd, sem = o.dims.vae, o.vae_sd, Semantics.cuda
    z = Golden("t2i").t("vae.decode_in")
    x = (z / d.scale_factor + d.shift_factor)           # AutoEncoder.decode's affine, bf16 op by op
    worst = {}
    for path, cin, cout, kind in _blocks(d):
        p = path[len("decoder."):] + "."
        if kind == "conv":
            want = ovae.conv(sd, path, x)
        elif kind == "res":
            want = ovae.resnet_block(sd, path + ".", x, cin, cout, d, sem)
        elif kind == "attn":
            want = ovae.attn_block(sd, path + ".", x, d, sem)
        elif kind == "up":
            want = ovae.conv(sd, path + ".conv", F.interpolate(x, scale_factor=2.0, mode="nearest"))
        else:
            want = ovae.gn_swish(sd, path, x, d, sem)
        got = eng.vae_block(path, x.to(torch.bfloat16))
        assert got.shape == want.shape, (path, got.shape, want.shape)
        s = ulp_stats(got, want.to(torch.bfloat16))
        worst[kind] = max(worst.get(kind, 0.0), s["rel_l2"])
        bound = {"conv": 2e-3, "gn": 2e-3, "up": 2e-3, "res": 6e-3, "attn": 6e-3}[kind]
        assert s["rel_l2"] < bound, (path, s)
        x = want                                         # the next block sees the ORACLE's activation: errors do not accumulate
    assert x.shape == (1, 3, 64, 64)
    print("worst rel-L2 per block kind:", {k: round(v, 5) for k, v in worst.items()})
    with pytest.raises(ValueError):
        eng.vae_block("decoder.up.9.block.0", z)
    with pytest.raises(ValueError):
        eng.vae_block("decoder.mid.block_1", z)          # 16 channels into a 512-channel block
