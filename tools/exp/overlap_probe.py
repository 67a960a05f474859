"""Does an HBM-bound weight stream hide under a tensor-bound pair linear on this (power-capped) B200?
Stream A: 28 gate/up linears (3072 x 37888 x 3584, fused SwiGLU).  Stream B: a read-only pass over 272 MB per linear (torch.sum of a bf16
tensor -- small CTAs that co-reside with the linear's one-CTA-per-SM grid).  Prints A alone, B alone, A with B running, and SM clocks."""
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from unimedvl_b200.engine import op_linear  # noqa: E402

M, N, K = 3072, 37888, 3584
x = torch.randn(M, K, device="cuda").bfloat16()
ws = [(torch.randn(N, K, device="cuda") * 0.02).bfloat16() for _ in range(4)]
bg = [torch.randn(N * K // 2, device="cuda").bfloat16() for _ in range(6)]      # 136 MB each, > L2 together
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
clocks = []
stop = False


def sample():
    while not stop:
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.sw_power_cap", "--format=csv,noheader,nounits"],
                           capture_output=True, text=True).stdout.strip()
        clocks.append(r)
        time.sleep(0.05)


def run(a, b, reps=6):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if a:
        with torch.cuda.stream(sa):
            e0.record()
            for r in range(reps):
                for i in range(28):
                    op_linear(x, ws[i % 4], None, None, epi=2)
            e1.record()
    if b:
        with torch.cuda.stream(sb):
            f0.record()
            for r in range(reps):
                for i in range(28):
                    bg[(2 * i) % 6].sum()
                    bg[(2 * i + 1) % 6].sum()
            f1.record()
    torch.cuda.synchronize()
    return (e0.elapsed_time(e1) / reps if a else None, f0.elapsed_time(f1) / reps if b else None)


run(True, True, 1)
th = threading.Thread(target=sample)
th.start()
for name, a, b in (("linears alone", True, False), ("stream alone", False, True), ("both", True, True), ("linears alone", True, False), ("both", True, True)):
    n0 = len(clocks)
    ta, tb = run(a, b)
    print(f"{name:14s}: 28 linears {ta if ta is None else round(ta, 2)} ms   28 x 272 MB reads {tb if tb is None else round(tb, 2)} ms   clocks/power/cap {clocks[n0:][-3:]}")
stop = True
th.join()
