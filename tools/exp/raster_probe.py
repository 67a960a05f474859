"""Tile-order groups of the pair linear (gemm_2cta.cu: tile_coords / pick_group_m): time per launch under sustained load for forced
group sizes next to the automatic choice.  UMV_RASTER_G is read per call."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from unimedvl_b200.engine import op_linear  # noqa: E402

shapes = [("prefill q|k|v", 8208, 4608, 3584, 0), ("prefill gate|up", 8208, 37888, 3584, 2), ("prefill down", 8208, 3584, 18944, 3), ("prefill o_proj", 8208, 3584, 3584, 3),
          ("flow gate|up", 3072, 37888, 3584, 2), ("flow down", 3072, 3584, 18944, 3), ("flow o_proj", 3072, 3584, 3584, 3),
          ("flow q|k|v", 3072, 4608, 3584, 0)]
only = os.environ.get("ONLY")
print("| shape | m_pairs | " + " | ".join(f"G={g}" for g in ("auto", "all", 2, 3, 4, 6, 8, 11, 17)) + " |")
for (name, M, N, K, epi) in shapes:
    if only and only not in name:
        continue
    x = torch.randn(M, K, device="cuda").bfloat16()
    ws = [(torch.randn(N, K, device="cuda") * 0.02).bfloat16() for _ in range(3)]
    b = None if epi == 2 else torch.zeros(N, device="cuda").bfloat16()
    res = torch.zeros(M, N, device="cuda").bfloat16() if epi == 3 else None
    mp = (M + 255) // 256
    cfgs = ("auto", "all", 2, 3, 4, 6, 8, 11, 17)
    acc = {g: [] for g in cfgs}
    reps = max(10, int(0.15e6 / (2.0 * M * N * K / 1.4e9)))          # ~0.15 s per figure, 3 rounds over all candidates
    for i in range(int(2e6 / (2.0 * M * N * K / 1.4e9))):            # 2 s of load first: steady power-capped clocks
        op_linear(x, ws[i % 3], b, res, epi=epi)
    for rnd in range(3):
        for g in cfgs:
            if isinstance(g, int) and g >= mp:
                continue
            if g == "auto":
                os.environ.pop("UMV_RASTER_G", None)
            else:
                os.environ["UMV_RASTER_G"] = str(mp if g == "all" else g)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps):
                op_linear(x, ws[i % 3], b, res, epi=epi)
            e1.record()
            torch.cuda.synchronize()
            acc[g].append(e0.elapsed_time(e1) / reps * 1e3)
    row = [f"{sum(acc[g]) / len(acc[g]):.1f}" if acc[g] else "-" for g in cfgs]
    print(f"| {name} {M}x{N}x{K} | {mp} | " + " | ".join(row) + " |", flush=True)
    del x, ws, b, res
    torch.cuda.empty_cache()
