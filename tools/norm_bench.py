import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unimedvl_b200.engine import op_rmsnorm
for warp, M in [(w, m) for w in ("0", "2", "1") for m in (272, 1026, 3096, 8208, 16416)]:
    os.environ["UMV_NORM_WARP"] = warp
    D = 3584
    x = torch.randn(M, D, device="cuda").bfloat16(); w = torch.ones(D, device="cuda").bfloat16()
    big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3): y = op_rmsnorm(x, w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    ts = []
    for _ in range(5):
        big.zero_()                      # flush L2
        e0.record(); y = op_rmsnorm(x, w); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    e0.record()
    for _ in range(10): y = op_rmsnorm(x, w)
    e1.record(); torch.cuda.synchronize()
    print(f"warp={warp} M={M:6d}: cold {min(ts):7.1f} us  warm {e0.elapsed_time(e1) * 100:7.1f} us   ({M * D * 4 / min(ts) / 1e3:.0f} GB/s cold)")
