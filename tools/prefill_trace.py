"""In-stream timeline of the VQA prefill at the 14B dims (8 samples of 448x448 + 32-token prompt per GPU): ViT + connector,
image prefill, prompt prefill, then 2 decode steps.  Per-kernel-class share of the critical path.
    python tools/prefill_trace.py [out.md]
    UMV_NCU=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file l.csv \\
        python tools/prefill_trace.py     # launch list of exactly one job"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from unimedvl_b200 import config as ucfg, _lib  # noqa: E402
from unimedvl_b200.bagel import Bagel  # noqa: E402
from unimedvl_b200.engine import Engine  # noqa: E402

B = bench.B_PER_GPU
dims = ucfg.bagel_7b_mot()
ntok_img = (bench.IMG // 14) ** 2 + 2
eng = Engine(dims, max_tokens=B * (ntok_img + 34), max_seqs=B, kv_pages=B * 24, enable_vit=True, enable_gen=False)
eng.fill_synthetic(0)
eng.finalize()
model = Bagel(eng, dims)
tok = dict(ucfg.QWEN25_TOKEN_IDS)
pixels, pos_ids, lens, prompts, _ = bench.synthetic_job(0)
model.vqa_generate(pixels, pos_ids, lens, prompts, tok, 3)
torch.cuda.synchronize()
CAP, NL = 8192, 32
_lib.check(eng.lib.umv_trace_begin(CAP))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if os.environ.get("UMV_NCU"):           # ncu --profile-from-start off: only this job is profiled
    torch.cuda.profiler.start()
ev0.record()
model.vqa_generate(pixels, pos_ids, lens, prompts, tok, 3)
ev1.record()
if os.environ.get("UMV_NCU"):
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
torch.cuda.synchronize()
stamps = np.zeros((CAP, 12), dtype=np.uint64)
names = C.create_string_buffer(CAP * NL)
n = C.c_int32()
_lib.check(eng.lib.umv_trace_read(stamps.ctypes.data_as(C.c_void_p), names, NL, CAP, C.byref(n)))
_lib.check(eng.lib.umv_trace_begin(0))
n = n.value
nm = [names.raw[i * NL:(i + 1) * NL].split(b"\0")[0].decode() for i in range(n)]
t = stamps[:n].astype(np.int64)
out = [f"# VQA prefill timeline: {B} samples x (1024 ViT tokens + 2 markers + 32 prompt tokens), + 2 decode steps; {n} traced "
       f"launches; wall {ev0.elapsed_time(ev1):.2f} ms\n\ncritical-path share = this kernel's last-CTA end minus the later of (the previous "
       "traced kernel's last-CTA end, this kernel's first-CTA start); the GPU-idle time before a kernel starts (host work between the prefill "
       "and the decode loop -- with tracing on, the decode graph is re-captured there, ~7 ms -- and untraced kernels) is the row `(idle before "
       "a kernel starts)`, not charged to the kernel that follows it\n\n| kernel | launches | mean us | total ms | share |\n|---|---:|---:|---:|---:|\n"]
agg = {}
idle = 0.0
for i in range(1, n):
    a = agg.setdefault(nm[i], [0, 0.0])
    a[0] += 1
    gap = max(0, t[i, 0] - t[i - 1, 3])
    idle += gap / 1e3
    a[1] += (t[i, 3] - t[i - 1, 3] - gap) / 1e3
agg["(idle before a kernel starts)"] = [n - 1, idle]
tot = sum(a[1] for a in agg.values())
for k, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    out.append(f"| `{k}` | {c} | {s / c:.1f} | {s / 1e3:.3f} | {100 * s / tot:.1f}% |\n")
ig = [i for i in range(n) if nm[i].startswith(("gemm<256,2,0>", "gemm2<256,2>"))]
if ig:
    i0 = ig[len(ig) // 4]               # a gate/up launch of the image prefill
    out.append("\n## around one image-prefill layer (us; start/wait/end relative to the first row's start)\n\n"
               "| # | kernel | first CTA start | wait passed | last CTA end | share |\n|---|---|---:|---:|---:|---:|\n")
    for i in range(i0 - 6, i0 + 4):
        out.append(f"| {i} | `{nm[i]}` | {(t[i, 0] - t[i0 - 6, 0]) / 1e3:.1f} | {(t[i, 1] - t[i0 - 6, 0]) / 1e3:.1f} | "
                   f"{(t[i, 3] - t[i0 - 6, 0]) / 1e3:.1f} | {(t[i, 3] - t[i - 1, 3]) / 1e3:.1f} |\n")
    i0 = ig[(3 * len(ig)) // 4]         # a gate/up launch of the prompt prefill (8 x 34 rows on the image context)
    out.append("\n## around one prompt-prefill layer (272 rows; us)\n\n"
               "| # | kernel | first CTA start | wait passed | last CTA end | share |\n|---|---|---:|---:|---:|---:|\n")
    for i in range(i0 - 6, i0 + 4):
        out.append(f"| {i} | `{nm[i]}` | {(t[i, 0] - t[i0 - 6, 0]) / 1e3:.1f} | {(t[i, 1] - t[i0 - 6, 0]) / 1e3:.1f} | "
                   f"{(t[i, 3] - t[i0 - 6, 0]) / 1e3:.1f} | {(t[i, 3] - t[i - 1, 3]) / 1e3:.1f} |\n")
iv = [i for i in range(n) if "K4304" in nm[i]]
if iv:
    i0 = iv[len(iv) // 2]               # an fc2 launch in the middle of the ViT
    out.append("\n## around one ViT layer (layernorm launches are not traced: their time shows up as the gap before the next row)\n\n"
               "| # | kernel | first CTA start | wait passed | last CTA end | share |\n|---|---|---:|---:|---:|---:|\n")
    for i in range(i0 - 5, i0 + 4):
        out.append(f"| {i} | `{nm[i]}` | {(t[i, 0] - t[i0 - 5, 0]) / 1e3:.1f} | {(t[i, 1] - t[i0 - 5, 0]) / 1e3:.1f} | "
                   f"{(t[i, 3] - t[i0 - 5, 0]) / 1e3:.1f} | {(t[i, 3] - t[i - 1, 3]) / 1e3:.1f} |\n")
gaps = sorted(((t[i, 0] - t[i - 1, 3]) / 1e3, i) for i in range(1, n))[::-1][:12]
out.append("\n## largest idle gaps (first CTA start minus the previous traced kernel's last CTA end: host work, untraced kernels, copies)\n\n"
           "| # | kernel | after | gap us | at ms |\n|---|---|---|---:|---:|\n")
for g, i in gaps:
    out.append(f"| {i} | `{nm[i]}` | `{nm[i - 1]}` | {g:.1f} | {(t[i, 0] - t[0, 0]) / 1e6:.2f} |\n")
out.append(f"\nsum of all positive gaps: {sum(max(0, (t[i, 0] - t[i - 1, 3])) for i in range(1, n)) / 1e6:.2f} ms of "
           f"{(t[n - 1, 3] - t[0, 0]) / 1e6:.2f} ms first start -> last end\n")
text = "".join(out)
print(text)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
