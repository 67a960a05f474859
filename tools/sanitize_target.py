#!/usr/bin/env python
"""Workload for compute-sanitizer (tools/sanitize.sh): the tiny-dims hot path end to end -- ViT + connector, image and prompt
prefill (tcgen05 linears + attention), 6 greedy decode steps (weight-major split-K linears, fused cluster decode attention;
eager launches, UMV_GRAPH=0, so every launch is visible to the tool), one guided flow step with three CFG branches (gen-mode
routing) and one VAE decode + encode; "r2": the round-2 paths -- a gen-mode full-mask forward of 2 x (marker + 256 latents + marker)
rows on segregated rows (row maps in the RoPE and attention kernels) through BOTH tcgen05 attention builds (UMV_ATTN_TK = 64 / 128), and
the token-major linears with forced tile-order groups (pair and single-CTA kernels).  Exits non-zero if any output is non-finite."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("UMV_GRAPH", "0")

import torch  # noqa: E402

from unimedvl_b200.autoencoder import AutoEncoder  # noqa: E402
from unimedvl_b200.bagel import Bagel  # noqa: E402
from unimedvl_b200.cache import NaiveCache  # noqa: E402
from unimedvl_b200.engine import Engine  # noqa: E402
from util import Golden, tiny_weights  # noqa: E402


def main():
    which = sys.argv[1:] or ["vqa", "flow", "vae"]
    dims, sd, vsd = tiny_weights(vae=True)
    eng = Engine(dims, max_tokens=1024, max_seqs=4, kv_pages=96, enable_vae=True)
    eng.load_state_dict(sd)
    vae = AutoEncoder(eng)
    vae.load_state_dict(vsd)
    eng.finalize()
    model = Bagel(eng, dims)
    ok = True
    if "vqa" in which:
        g = Golden("vqa")
        c = model.forward_cache_update_vit(NaiveCache(dims.llm.layers), **g.group("vqa.vit_in"))
        c = model.forward_cache_update_text(c, **g.group("vqa.text_in"))
        toks = model.generate_text(past_key_values=c, max_length=7, end_token_id=None, **g.group("vqa.start")).cpu()
        ok &= bool(torch.equal(toks, g.t("vqa.tokens")[:7]))
        print("vqa tokens", toks.T.tolist(), "== golden:", ok)
    if "flow" in which:
        g = Golden("t2i")
        ctx = model.forward_cache_update_text(NaiveCache(dims.llm.layers), **g.group("t2i.text_in"))
        cfg = model.forward_cache_update_text(NaiveCache(dims.llm.layers), **g.group("t2i.cfg_text_in"))
        gi, ct, ci = g.group("t2i.latent_in"), g.group("t2i.cfg_text"), g.group("t2i.cfg_img")
        lat = model.generate_image(
            past_key_values=ctx, cfg_text_past_key_values=cfg, cfg_img_past_key_values=ctx, num_timesteps=3, timestep_shift=3.0,
            cfg_text_scale=4.0, cfg_img_scale=1.5, cfg_interval=(0.0, 1.0), cfg_renorm_type="text_channel", **gi,
            cfg_text_packed_position_ids=ct["cfg_packed_position_ids"], cfg_text_packed_query_indexes=ct["cfg_packed_query_indexes"],
            cfg_text_key_values_lens=ct["cfg_key_values_lens"], cfg_text_packed_key_value_indexes=ct["cfg_packed_key_value_indexes"],
            cfg_img_packed_position_ids=ci["cfg_packed_position_ids"], cfg_img_packed_query_indexes=ci["cfg_packed_query_indexes"],
            cfg_img_key_values_lens=ci["cfg_key_values_lens"], cfg_img_packed_key_value_indexes=ci["cfg_packed_key_value_indexes"])
        fin = all(bool(torch.isfinite(x).all()) for x in lat)
        ok &= fin
        print("flow latents finite:", fin)
    if "vae" in which:
        z = Golden("t2i").t("vae.decode_in").cuda()
        img = vae.decode(z)
        mom = vae.encode_moments((torch.rand(1, 3, 32, 48) * 2 - 1).bfloat16().cuda())
        fin = bool(torch.isfinite(img.float()).all() and torch.isfinite(mom.float()).all())
        ok &= fin
        print("vae finite:", fin)
    if "r2" in which:
        from unimedvl_b200.engine import op_linear
        D = dims.llm.hidden
        gsd = torch.Generator().manual_seed(0)
        outs = []
        for tk in ("64", "128"):
            os.environ["UMV_ATTN_TK"] = tk
            seqs = [eng.seq_new() for _ in range(2)]
            ctx = (torch.randn(40, D, generator=torch.Generator().manual_seed(1)) * 0.5).bfloat16()
            eng.llm_forward(ctx, seqs, [30, 10], list(range(30)) + list(range(10)), want_hidden=False)
            x = (torch.randn(2 * 258, D, generator=torch.Generator().manual_seed(2)) * 0.5).bfloat16()
            is_gen = ([0] + [1] * 256 + [0]) * 2
            h = eng.llm_forward(x, seqs, [258, 258], [30] * 258 + [10] * 258, row_is_gen=is_gen, is_causal=False, update_kv=True)
            outs.append(h.float().cpu())
            for s_ in seqs:
                eng.seq_free(s_)
        os.environ.pop("UMV_ATTN_TK", None)
        rel = ((outs[0] - outs[1]).norm() / outs[1].norm()).item()
        fin = bool(torch.isfinite(outs[0]).all()) and rel < 2e-2
        xs = torch.randn(700, 512, generator=gsd).bfloat16().cuda()
        ws = (torch.randn(1024, 512, generator=gsd) * 0.05).bfloat16().cuda()
        ys = []
        for pair in ("1", "0"):
            os.environ["UMV_2CTA"] = pair
            for g_ in ("1", "2", "100000"):
                os.environ["UMV_RASTER_G"] = g_
                ys.append(op_linear(xs, ws, None, None, epi=0, impl=1).clone())
        os.environ.pop("UMV_2CTA", None)
        os.environ.pop("UMV_RASTER_G", None)
        same = all(bool(torch.equal(y, ys[0])) for y in ys)
        ok &= fin and same
        print("r2: segregated gen forward, attention builds agree to", f"{rel:.2e}", "; tile orders bit-identical:", same)
    torch.cuda.synchronize()
    print("launches", eng.launch_count())
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
