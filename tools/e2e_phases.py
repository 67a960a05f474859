"""Host-side phase timeline of the end-to-end VQA job bench.py times as `e2e` (Bagel.vqa_generate_images, 14B dims,
8 samples): per phase, the CPU time of the call (no synchronisation: how far the host runs ahead) and the time with a
device synchronisation after every phase (what the GPU needs for it).
    python tools/e2e_phases.py [out.md]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from unimedvl_b200 import config as ucfg  # noqa: E402
from unimedvl_b200.bagel import Bagel  # noqa: E402
from unimedvl_b200.engine import Engine  # noqa: E402

B = bench.B_PER_GPU
dims = ucfg.bagel_7b_mot()
ntok_img = (bench.IMG // 14) ** 2 + 2
eng = Engine(dims, max_tokens=B * (ntok_img + 34), max_seqs=B, kv_pages=B * 64, enable_vit=True, enable_gen=False)
eng.fill_synthetic(0)
eng.finalize()
model = Bagel(eng, dims)
tok = dict(ucfg.QWEN25_TOKEN_IDS)
_, _, _, prompts, images = bench.synthetic_job(0)

SYNC = False
log = []


def wrap(obj, name, label=None):
    fn = getattr(obj, name)
    label = label or name

    def timed(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        if SYNC:
            torch.cuda.synchronize()
        log.append((label, (time.perf_counter() - t0) * 1e3))
        return r
    setattr(obj, name, timed)


wrap(eng, "patchify_u8")
wrap(eng, "vit_embed")
wrap(eng, "llm_forward")
wrap(eng, "generate_text", "Engine.generate_text")
wrap(model, "forward_cache_update_vit")
wrap(model, "forward_cache_update_text")
wrap(model, "generate_text", "Bagel.generate_text")

out = ["# e2e phase timeline (Bagel.vqa_generate_images, 8 x 448x448, 32-token prompt, 128-token decode)\n\n"
       "Nested: `forward_cache_update_vit` contains `vit_embed` + one `llm_forward`; `forward_cache_update_text` one "
       "`llm_forward`; `Bagel.generate_text` contains `Engine.generate_text`.\n"]
for sync in (False, True):
    SYNC = sync
    for _ in range(3):
        model.vqa_generate_images(images, prompts, tok, bench.DECODE_STEPS)
    torch.cuda.synchronize()
    log.clear()
    t0 = time.perf_counter()
    model.vqa_generate_images(images, prompts, tok, bench.DECODE_STEPS)
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) * 1e3
    out.append(f"\n## {'synchronised after every phase (GPU time per phase)' if sync else 'asynchronous (CPU time per call)'}: "
               f"total {total:.2f} ms\n\n| call | ms |\n|---|---:|\n")
    for name, ms in log:
        out.append(f"| `{name}` | {ms:.2f} |\n")
text = "".join(out)
print(text)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
