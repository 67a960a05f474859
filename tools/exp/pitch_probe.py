"""Does the 8,608-byte row pitch of the ViT MLP intermediate (4,304 columns) cost the fc2 linear?  Same product with K zero-padded to
4,352 (128-byte aligned rows, 1 % more MMA work), next to cuBLAS on both."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from unimedvl_b200.engine import op_linear  # noqa: E402


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


M, N = 8192, 1152
for K in (4304, 4352, 4096):
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = torch.zeros(N, device="cuda").bfloat16()
    res = torch.zeros(M, N, device="cuda").bfloat16()
    print(f"fc2 K={K}: engine {timeit(lambda: op_linear(x, w, b, res, epi=3)):.1f} us   cuBLAS {timeit(lambda: torch.matmul(x, w.T)):.1f} us")
M, K = 8192, 1152
for N in (4304, 4352):
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = torch.zeros(N, device="cuda").bfloat16()
    print(f"fc1 N={N}: engine {timeit(lambda: op_linear(x, w, b, None, epi=1)):.1f} us (GELU)  {timeit(lambda: op_linear(x, w, b, None, epi=0)):.1f} us (bias only)   cuBLAS {timeit(lambda: torch.matmul(x, w.T)):.1f} us")
