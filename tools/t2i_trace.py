"""In-stream timeline of the text-to-image flow forwards at the 14B dims (4 images of 256x256 per GPU, 3 CFG branches):
per-kernel-class share of the critical path of two Euler steps.     python tools/t2i_trace.py [out.md]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimedvl_b200 import config as ucfg, packing, synth, _lib  # noqa: E402
from unimedvl_b200.bagel import Bagel  # noqa: E402
from unimedvl_b200.cache import NaiveCache, paged_handle  # noqa: E402
from unimedvl_b200.engine import Engine  # noqa: E402

B, H, W = int(os.environ.get("B", "4")), 256, 256
dims = ucfg.bagel_7b_mot()
eng = Engine(dims, max_tokens=3 * B * 258 + 64, max_seqs=3 * B + 2, kv_pages=64, enable_vit=False, enable_gen=True)
eng.fill_synthetic(0)
eng.finalize()
model = Bagel(eng, dims)
tok = dict(ucfg.QWEN25_TOKEN_IDS)


class _Ids:
    def encode(self, i): return synth.synthetic_prompt_ids(100 + i, 30)


g, lens, rope = packing.prepare_prompts([0] * B, [0] * B, list(range(B)), _Ids(), tok)
ctx = model.forward_cache_update_text(NaiveCache(dims.llm.layers), **g)
cfg_text = NaiveCache(dims.llm.layers)
paged_handle(cfg_text, eng, B)
cfg_img = model.forward_cache_update_text(NaiveCache(dims.llm.layers), **g)
torch.manual_seed(42)
gi = model.prepare_vae_latent(lens, rope, [(H, W)] * B, tok)
ct = model.prepare_vae_latent_cfg([0] * B, [0] * B, [(H, W)] * B)
ci = model.prepare_vae_latent_cfg(lens, rope, [(H, W)] * B)


def run(steps):
    return model.generate_image(
        past_key_values=ctx, cfg_text_past_key_values=cfg_text, cfg_img_past_key_values=cfg_img, num_timesteps=steps,
        timestep_shift=3.0, cfg_text_scale=4.0, cfg_img_scale=1.5, cfg_interval=(0.4, 1.0), cfg_renorm_min=0.0,
        cfg_renorm_type="global", **gi,
        cfg_text_packed_position_ids=ct["cfg_packed_position_ids"], cfg_text_packed_query_indexes=ct["cfg_packed_query_indexes"],
        cfg_text_key_values_lens=ct["cfg_key_values_lens"], cfg_text_packed_key_value_indexes=ct["cfg_packed_key_value_indexes"],
        cfg_img_packed_position_ids=ci["cfg_packed_position_ids"], cfg_img_packed_query_indexes=ci["cfg_packed_query_indexes"],
        cfg_img_key_values_lens=ci["cfg_key_values_lens"], cfg_img_packed_key_value_indexes=ci["cfg_packed_key_value_indexes"])


run(4)
torch.cuda.synchronize()
CAP, NL = 4096, 32
_lib.check(eng.lib.umv_trace_begin(CAP))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if os.environ.get("UMV_NCU"):           # ncu --profile-from-start off: only these two steps are profiled
    torch.cuda.profiler.start()
ev0.record()
run(3)                      # two Euler steps, both inside the CFG interval
ev1.record()
torch.cuda.synchronize()
if os.environ.get("UMV_NCU"):
    torch.cuda.profiler.stop()
stamps = np.zeros((CAP, 12), dtype=np.uint64)
names = C.create_string_buffer(CAP * NL)
n = C.c_int32()
_lib.check(eng.lib.umv_trace_read(stamps.ctypes.data_as(C.c_void_p), names, NL, CAP, C.byref(n)))
_lib.check(eng.lib.umv_trace_begin(0))
n = n.value
nm = [names.raw[i * NL:(i + 1) * NL].split(b"\0")[0].decode() for i in range(n)]
t = stamps[:n].astype(np.int64)
out = [f"# T2I flow timeline: {B} images x 3 CFG branches, 2 Euler steps, {n} traced launches; "
       f"wall {ev0.elapsed_time(ev1):.2f} ms, first start -> last end {(t[:, 3].max() - t[0, 0]) / 1e6:.2f} ms\n\n"
       "critical-path share = this kernel's last-CTA end minus the previous traced kernel's last-CTA end (untraced glue kernels "
       "in between are charged to the next traced kernel); run = dependency wait passed -> last CTA end\n\n"
       "| kernel | launches | mean us | total ms | share | mean run us |\n|---|---:|---:|---:|---:|---:|\n"]
agg = {}
for i in range(1, n):
    a = agg.setdefault(nm[i], [0, 0.0, 0.0])
    a[0] += 1
    a[1] += (t[i, 3] - t[i - 1, 3]) / 1e3
    a[2] += (t[i, 3] - t[i, 1]) / 1e3
tot = sum(a[1] for a in agg.values())
for k, (c, s, r) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {c} | {s / c:.1f} | {s / 1e3:.3f} | {100 * s / tot:.1f}% | {r / c:.1f} |\n")
text = "".join(out)
print(text)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
