"""In-graph timeline of one decode forward at the 14B dims (B samples, ctx ~1058): every traced kernel stamps
%globaltimer (first CTA start / dependency wait passed / first CTA end / last CTA end); the slots of the LAST replayed
step are read back and printed per kernel of one middle layer plus per-kernel-class averages over all layers.

    python tools/decode_trace.py [out.md]
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unimedvl_b200 import config as ucfg, _lib  # noqa: E402
from unimedvl_b200.engine import Engine  # noqa: E402

torch.manual_seed(0)
B = int(os.environ.get("B", "8"))
steps = int(os.environ.get("STEPS", "16"))
ctx = int(os.environ.get("CTX", "1058"))
d = ucfg.bagel_7b_mot()
eng = Engine(d, max_tokens=max(1100, ctx), max_seqs=B, kv_pages=B * ((ctx + steps * 4) // 64 + 2), enable_gen=False)
eng.fill_synthetic(0)
eng.finalize()
seqs = [eng.seq_new() for _ in range(B)]
for b in range(B):
    x = (torch.randn(ctx, d.llm.hidden, device="cuda") * 0.05).bfloat16()
    eng.llm_forward(x, [seqs[b]], [ctx], list(range(ctx)), is_causal=True, update_kv=True, want_hidden=False)
eng.generate_text(seqs, [151644] * B, [ctx] * B, 4)
torch.cuda.synchronize()
lib = eng.lib
CAP, NL = 1024, 32
_lib.check(lib.umv_trace_begin(CAP))
eng.generate_text(seqs, [151644] * B, [ctx + 4] * B, steps)
stamps = np.zeros((CAP, 12), dtype=np.uint64)
names = C.create_string_buffer(CAP * NL)
n = C.c_int32()
_lib.check(lib.umv_trace_read(stamps.ctypes.data_as(C.c_void_p), names, NL, CAP, C.byref(n)))
_lib.check(lib.umv_trace_begin(0))
n = n.value
nm = [names.raw[i * NL:(i + 1) * NL].split(b"\0")[0].decode() for i in range(n)]
t = stamps[:n].astype(np.int64)
t0 = t[0, 0]
rows = []
for i in range(n):
    nxt_start = t[i + 1, 1] if i + 1 < n else t[i, 3]
    rows.append((i, nm[i], (t[i, 0] - t0) / 1e3, (t[i, 1] - t[i, 0]) / 1e3, (t[i, 3] - t[i, 1]) / 1e3, (t[i, 3] - t[i, 2]) / 1e3,
                 (t[i, 3] - t0) / 1e3))
out = []
out.append(f"# Decode forward timeline (in-graph, last of {steps} replays), B={B}, ctx={ctx}; {n} traced launches; times in us\n")
out.append(f"total first-start -> last-end: {(t[:, 3].max() - t0) / 1e3:.1f} us\n")
per = 7 if nm.count("attn_decode") else 9
mid = [r for r in rows if nm[r[0]] != ""][:]
lo = n // 2 - (n // 2) % per
out.append("\n## one middle layer (start = first CTA start relative to the step start; pre = start -> dependency wait passed "
           "(linears: all tiles requested); run = wait -> last CTA end; skew = first CTA end -> last CTA end)\n")
out.append("| # | kernel | start | pre | run | skew | end | gap to next end |\n|---|---|---:|---:|---:|---:|---:|---:|\n")
for r in rows[lo:lo + 2 * per]:
    i = r[0]
    gap = (t[i + 1, 3] - t[i, 3]) / 1e3 if i + 1 < n else 0
    out.append(f"| {i} | `{r[1]}` | {r[2]:.1f} | {r[3]:.1f} | {r[4]:.1f} | {r[5]:.1f} | {r[6]:.1f} | {gap:.1f} |\n")
ia = [i for i in range(n) if nm[i] == "attn_decode"]
if ia:
    i = ia[len(ia) // 2]
    out.append("\n## inside attn_decode (first CTA; us after the dependency wait): " +
               ", ".join(f"dbg{k}={(t[i, 4 + k] - t[i, 1]) / 1e3:.1f}" for k in range(8) if t[i, 4 + k] > 0) +
               f", end={(t[i, 2] - t[i, 1]) / 1e3:.1f}\n")
for pref in ("dec<1,1>", "dec<1,2>", "dec<0,3> N3584 K3584"):
    ib = [i for i in range(n) if nm[i].startswith(pref)]
    if ib:
        i = ib[len(ib) // 2]
        out.append(f"\n## inside {nm[i]} (first CTA; us after its start): " +
                   ", ".join(f"dbg{k}={(t[i, 4 + k] - t[i, 0]) / 1e3:.1f}" for k in range(8) if t[i, 4 + k] > 0) +
                   f", all tiles requested={(t[i, 1] - t[i, 0]) / 1e3:.1f}, first CTA end={(t[i, 2] - t[i, 0]) / 1e3:.1f}, "
                   f"last CTA end={(t[i, 3] - t[i, 0]) / 1e3:.1f}; previous kernel's last CTA ended at {(t[i - 1, 3] - t[i, 0]) / 1e3:.1f}\n")
ia = [i for i in range(n) if nm[i] == "add_rmsnorm"]
if ia:
    i = ia[len(ia) // 2]
    out.append("\n## inside add_rmsnorm (first CTA; us after the dependency wait): " +
               ", ".join(f"dbg{k}={(t[i, 4 + k] - t[i, 1]) / 1e3:.1f}" for k in range(8) if t[i, 4 + k] > 0) +
               f", end={(t[i, 2] - t[i, 1]) / 1e3:.1f}\n")
out.append("\n## per kernel class: mean time between the previous kernel's last-CTA end and this kernel's last-CTA end "
           "(the kernel's share of the critical path)\n")
out.append("| kernel | launches | mean us | total us |\n|---|---:|---:|---:|\n")
agg = {}
for i in range(1, n):
    dt = (t[i, 3] - t[i - 1, 3]) / 1e3
    a = agg.setdefault(nm[i], [0, 0.0])
    a[0] += 1
    a[1] += dt
for k, (c, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {c} | {tot / c:.2f} | {tot:.1f} |\n")
text = "".join(out)
print(text)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
