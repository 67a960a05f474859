"""Why is the decode phase of the e2e job slower than the resident-context decode bench.py reports as `value`?
Times Engine-level decode (128 steps, B=8, ctx 1058) (a) from forks of one resident context, back to back (bench `value`),
(b) right after a fresh prefill (the e2e order), (c) after a fresh prefill + an idle pause, with the decode-graph cache on/off.
    python tools/decode_after_prefill.py"""
import os
import sys
import time
from copy import deepcopy

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from unimedvl_b200 import config as ucfg, packing  # noqa: E402
from unimedvl_b200.bagel import Bagel  # noqa: E402
from unimedvl_b200.cache import NaiveCache  # noqa: E402
from unimedvl_b200.engine import Engine  # noqa: E402

B = bench.B_PER_GPU
dims = ucfg.bagel_7b_mot()
ntok_img = (bench.IMG // 14) ** 2 + 2
eng = Engine(dims, max_tokens=B * (ntok_img + 34), max_seqs=B, kv_pages=B * 64, enable_vit=True, enable_gen=False)
eng.fill_synthetic(0)
eng.finalize()
model = Bagel(eng, dims)
tok = dict(ucfg.QWEN25_TOKEN_IDS)
pixels, pos_ids, lens, prompts, images = bench.synthetic_job(0)
pixels_d, pos_d = pixels.cuda(), pos_ids.cuda()


class _Ids:
    def encode(self, i): return list(prompts[i])


def prefill():
    cache = NaiveCache(dims.llm.layers)
    zeros = [0] * B
    L = packing._image_block_layout(zeros, zeros, lens, tok)
    g = dict(packed_text_ids=torch.as_tensor(L["text_ids"]), packed_text_indexes=torch.as_tensor(L["text_idx"]),
             packed_vit_tokens=pixels_d, packed_vit_token_indexes=torch.as_tensor(L["img_idx"]),
             packed_vit_position_ids=pos_d, vit_token_seqlens=lens, packed_position_ids=torch.as_tensor(L["pos"]),
             packed_seqlens=L["seqlens"], packed_indexes=torch.as_tensor(L["packed_idx"]),
             packed_key_value_indexes=torch.as_tensor(L["kv_indexes"]), key_values_lens=zeros)
    cache = model.forward_cache_update_vit(cache, **g)
    gp, kvl, rope = packing.prepare_prompts(L["seqlens"], [1] * B, list(range(B)), _Ids(), tok)
    cache = model.forward_cache_update_text(cache, **gp)
    return cache, packing.prepare_start_tokens(kvl, rope, tok)


def decode(cache, start):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    model.generate_text(past_key_values=cache, max_length=bench.DECODE_STEPS, end_token_id=None, **start)
    e1.record()
    cpu = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), cpu


for cache_flag in ("1", "0"):
    os.environ["UMV_GRAPH_CACHE"] = cache_flag
    cache, start = prefill()
    for _ in range(3):
        decode(deepcopy(cache), start)
    a = [decode(deepcopy(cache), start) for _ in range(4)]
    b = []
    for _ in range(4):
        c, s = prefill()
        b.append(decode(c, s))
    c_ = []
    for _ in range(3):
        c, s = prefill()
        torch.cuda.synchronize()
        time.sleep(0.3)
        c_.append(decode(c, s))
    d = []
    for _ in range(3):                      # prefill, then decode from a fork of the OLD resident context
        prefill()
        d.append(decode(deepcopy(cache), start))
    fmt = lambda xs: ", ".join(f"{g:.1f} (cpu {c:.1f})" for g, c in xs)
    print(f"UMV_GRAPH_CACHE={cache_flag}")
    print(f"  (a) forks of a resident context, back to back : {fmt(a)}")
    print(f"  (b) right after a fresh prefill               : {fmt(b)}")
    print(f"  (c) fresh prefill + 0.3 s idle                : {fmt(c_)}")
    print(f"  (d) fresh prefill, decode the OLD context     : {fmt(d)}")
