"""Experiment: does torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction explain the reference-CUDA vs oracle gap?"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "baseline"))
import torch
import refharness as rh
from util import tiny_weights, Golden, make_oracle, Semantics, ulp_stats
from oracle import vit as ovit, numerics as nm

dims, sd, vsd = tiny_weights(vae=True)
ref, _ = rh.build_reference(dims, sd, vsd, "cuda")
o = make_oracle(Semantics.cuda, vae=True)
gv = Golden("vqa").group("vqa.vit_in"); lens = gv["vit_token_seqlens"]
ot = {}
ovit.vit_forward(o.sd, o.dims.vit, gv["packed_vit_tokens"], gv["packed_vit_position_ids"], lens, Semantics.cuda, True, ot)
cu = torch.nn.functional.pad(torch.cumsum(lens, 0), (1, 0)).to(torch.int32).cuda()
vm = ref.vit_model.vision_model
for flag in (True, False):
    torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction = flag
    taps = {}
    hs = [vm.embeddings.register_forward_hook(lambda m, i, out: taps.__setitem__("vit_embed", out.detach().cpu()))]
    for li, layer in enumerate(vm.encoder.layers):
        hs.append(layer.register_forward_hook(lambda m, i, out, li=li: taps.__setitem__(f"vit_layer{li}", (out[0] if isinstance(out, tuple) else out).detach().cpu())))
    with torch.no_grad(), rh.autocast("cuda"):
        ref.vit_model(packed_pixel_values=gv["packed_vit_tokens"].cuda(), packed_flattened_position_ids=gv["packed_vit_position_ids"].cuda(), cu_seqlens=cu, max_seqlen=int(lens.max()))
    for h in hs: h.remove()
    for k in ("vit_embed", "vit_layer0", "vit_layer1"):
        s = ulp_stats(ot[k], taps[k])
        print("reduced_precision_reduction", flag, k, {a: round(b, 6) if isinstance(b, float) else b for a, b in s.items()})
# plain op check: one linear, tiny and wide shapes
for (M, N, K) in ((182, 144, 588), (2052, 1152, 588), (8, 3584, 3584), (2052, 3584, 1152), (1058, 4608, 3584)):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(M, K, device="cuda", generator=g).bfloat16(); w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    exact = (x.double() @ w.double().T + b.double()).float().bfloat16()
    for flag in (True, False):
        torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction = flag
        y = torch.nn.functional.linear(x, w, b)
        s = ulp_stats(y, exact)
        print("linear", (M, N, K), "reduced", flag, "frac", round(s["frac"], 5), "frac_gt1", round(s["frac_gt1"], 6), "rel_l2", round(s["rel_l2"], 6))
