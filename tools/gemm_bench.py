"""Token-major / weight-major linear microbenchmark (TFLOP/s, GB/s) at the shapes of the 14B path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unimedvl_b200.engine import op_linear
shapes = [(8208, 4608, 3584, 0), (8208, 3584, 3584, 3), (8208, 37888, 3584, 2), (8208, 3584, 18944, 3), (3096, 37888, 3584, 2),
          (3096, 3584, 18944, 3), (1032, 37888, 3584, 2), (8192, 3456, 1152, 0), (8192, 4304, 1152, 1), (8192, 1152, 4304, 3),
          (65536, 256, 2304, 0), (8, 37888, 3584, 2), (24, 37888, 3584, 2), (24, 3584, 18944, 0)]
for (M, N, K, epi) in shapes:
    x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = None if epi == 2 else torch.zeros(N, device="cuda").bfloat16()
    res = torch.zeros(M, N, device="cuda").bfloat16() if epi == 3 else None
    for _ in range(3): y = op_linear(x, w, b, res, epi=epi)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    n = 10
    e0.record()
    for _ in range(n): y = op_linear(x, w, b, res, epi=epi)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    ref = torch.matmul(x, w.T)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): ref = torch.matmul(x, w.T)
    e1.record(); torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / n
    print(f"M={M:6d} N={N:6d} K={K:6d} epi={epi}: {ms*1e3:9.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s  {N*K*2/ms/1e6:8.1f} GB/s(w) | cuBLAS {ms_ref*1e3:9.1f} us {2*M*N*K/ms_ref/1e9:8.1f} TFLOP/s", flush=True)
