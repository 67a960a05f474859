"""ncu target: a few decode forwards at the 14B dims (B=8, ctx 1058), bracketed by cudaProfilerStart/Stop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_decode.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unimedvl_b200 import config as ucfg  # noqa: E402
from unimedvl_b200.engine import Engine  # noqa: E402

B = int(os.environ.get("B", "8"))
steps = int(os.environ.get("STEPS", "2"))
d = ucfg.bagel_7b_mot()
eng = Engine(d, max_tokens=1100, max_seqs=B, kv_pages=B * 24, enable_gen=False)
eng.fill_synthetic(0)
eng.finalize()
seqs = [eng.seq_new() for _ in range(B)]
for b in range(B):
    x = (torch.randn(1058, d.llm.hidden, device="cuda") * 0.05).bfloat16()
    eng.llm_forward(x, [seqs[b]], [1058], [0] * 1026 + list(range(1, 33)), is_causal=True, update_kv=True, want_hidden=False)
eng.generate_text(seqs, [151644] * B, [33] * B, 2)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.generate_text(seqs, [151644] * B, [35] * B, steps)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches", eng.launch_count())
