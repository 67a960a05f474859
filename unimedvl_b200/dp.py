"""Data-parallel plumbing: requests are independent, so the global batch is split into
contiguous per-rank shards, every rank runs the whole path on its shard with replicated weights, and the
only collective is one all_gather of the output tokens / images (NCCL over NVLink on GPUs, gloo in CPU tests)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split; the first n_items % world ranks take one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(items: Sequence, rank: int, world: int) -> List:
    lo, hi = shard_bounds(len(items), rank, world)
    return list(items[lo:hi])


def gather_tokens(tokens: torch.Tensor, n_items: int) -> torch.Tensor:
    """tokens: [steps, B_local] on every rank -> [steps, n_items] on every rank (batch order restored).
    Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return tokens
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(n_items, r, world) for r in range(world)]
    bmax = max(hi - lo for lo, hi in sizes)
    steps = tokens.shape[0]
    pad = torch.zeros((steps, bmax), dtype=tokens.dtype, device=tokens.device)
    pad[:, :tokens.shape[1]] = tokens
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous())          # the path's only collective
    return torch.cat([out[r][:, :hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=1)


def gather_images(images: torch.Tensor, n_items: int) -> torch.Tensor:
    """images: uint8 [B_local, H, W, 3] (one size per job, as BASELINE.json's text-to-image configs) on every rank ->
    [n_items, H, W, 3] on every rank in batch order; the text-to-image counterpart of gather_tokens (SURVEY.md section 8e:
    768 KB per rank at 4 x 256 x 256)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return images
    world = dist.get_world_size()
    sizes = [shard_bounds(n_items, r, world) for r in range(world)]
    bmax = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((bmax,) + tuple(images.shape[1:]), dtype=images.dtype, device=images.device)
    pad[:images.shape[0]] = images
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous())
    return torch.cat([out[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
