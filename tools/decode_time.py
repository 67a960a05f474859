"""Device-timed decode forwards at the 14B dims (B samples, ctx ~1058 -> +steps), CUDA-graph replay as in the product path.
Tuning knobs are process-wide environment variables read once by the engine (UMV_PF_*, UMV_FUSED_ATTN, ...), so each
configuration runs in its own process:

    UMV_PF_GU_KB=20 python tools/decode_time.py            -> prints one JSON line with ms per decode forward
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unimedvl_b200 import config as ucfg  # noqa: E402
from unimedvl_b200.engine import Engine  # noqa: E402

torch.manual_seed(0)
B = int(os.environ.get("B", "8"))
steps = int(os.environ.get("STEPS", "128"))
ctx = int(os.environ.get("CTX", "1058"))
d = ucfg.bagel_7b_mot()
eng = Engine(d, max_tokens=max(1100, ctx), max_seqs=B, kv_pages=B * ((ctx + steps * 4) // 64 + 2), enable_gen=False)
eng.fill_synthetic(0)
eng.finalize()
seqs = [eng.seq_new() for _ in range(B)]
for b in range(B):
    x = (torch.randn(ctx, d.llm.hidden, device="cuda") * 0.05).bfloat16()
    eng.llm_forward(x, [seqs[b]], [ctx], list(range(ctx)), is_causal=True, update_kv=True, want_hidden=False)
pos = ctx
best = None
for rep in range(3):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(eng.stream):
        ev0.record()
    toks = eng.generate_text(seqs, [151644] * B, [pos] * B, steps)
    with torch.cuda.stream(eng.stream):
        ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    if rep > 0:
        best = ms if best is None else min(best, ms)
    for s in seqs:
        eng.seq_truncate(s, ctx)
knobs = {k: v for k, v in os.environ.items() if k.startswith("UMV_")}
print(json.dumps({"ms_per_decode_forward": round(best, 4), "B": B, "ctx": ctx, "steps": steps, "knobs": knobs,
                  "tokens_checksum": int(toks.sum().item())}), flush=True)
