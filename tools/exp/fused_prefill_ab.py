"""e2e VQA job (8 x 448x448 + 32-token prompt -> 128 tokens, 14B dims) with the image block and the prompt prefilled in one forward
(umv_forward_cache_update_vit_prompt) vs the two calls of the reference's drivers; alternating on one box."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from unimedvl_b200 import config as ucfg  # noqa: E402
from unimedvl_b200.bagel import Bagel  # noqa: E402
from unimedvl_b200.engine import Engine  # noqa: E402

B = bench.B_PER_GPU
dims = ucfg.bagel_7b_mot()
eng = Engine(dims, max_tokens=B * ((bench.IMG // 14) ** 2 + 36), max_seqs=B, kv_pages=B * 64, enable_vit=True, enable_gen=False)
eng.fill_synthetic(0)
eng.finalize()
model = Bagel(eng, dims)
tok = dict(ucfg.QWEN25_TOKEN_IDS)
_, _, _, prompts, images = bench.synthetic_job(0)
steps = int(os.environ.get("STEPS", "128"))


def run(fused, n=3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        t = model.vqa_generate_images(images, prompts, tok, steps, fused_prefill=fused)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t


for f in (True, False):
    run(f, 2)
a = run(True)[1]
b = run(False)[1]
print("tokens identical:", bool(torch.equal(a, b)))
for rnd in range(3):
    for f in (True, False):
        print(f"fused={f}: {run(f)[0]:.2f} ms per job", flush=True)
