import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from util import ulp_stats
from unimedvl_b200.engine import op_linear
torch.manual_seed(0)
for (M, N, K) in [(24, 512, 24), (24, 32, 4608), (24, 512, 64), (24, 512, 32), (64, 512, 24), (100, 512, 24), (24, 32, 512), (24, 128, 512), (24, 64, 512), (200, 32, 512)]:
    x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    b = (torch.randn(N, device="cuda") * 0.1).bfloat16()
    ref = (x.float() @ w.float().T + b.float()).bfloat16()
    for impl in (3, 1, 2):
        if impl == 2 and M > 64: continue
        try:
            y = op_linear(x, w, b, None, epi=0, impl=impl); torch.cuda.synchronize()
            print(M, N, K, impl, {k: round(v, 5) for k, v in ulp_stats(y, ref).items()})
        except Exception as ex:
            print(M, N, K, impl, "EXC", ex)
