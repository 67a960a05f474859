"""Data-parallel host logic with world_size 2 over gloo (CPU): sharding + the single output all_gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unimedvl_b200 import dp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = dp.shard_bounds(n_items, rank, world)
    # "decode" of a shard: token[s, i] is a function of the GLOBAL sample id, as on independent requests
    ids = torch.arange(lo, hi)
    toks = torch.stack([ids * 1000 + s for s in range(5)], 0)
    full = dp.gather_tokens(toks, n_items)
    imgs = (ids.view(-1, 1, 1, 1) * 3 + torch.arange(3).view(1, 1, 1, 3)).expand(-1, 4, 6, 3).to(torch.uint8)
    q.put((rank, full, dp.gather_images(imgs.contiguous(), n_items)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 33):
        for w in (1, 2, 3, 8):
            spans = [dp.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_gather_tokens_world2_ragged():
    n_items, world = 7, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    got = {r: t for r, t, _ in res}
    got_img = {r: i for r, _, i in res}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.stack([torch.arange(n_items) * 1000 + s for s in range(5)], 0)
    want_img = (torch.arange(n_items).view(-1, 1, 1, 1) * 3 + torch.arange(3).view(1, 1, 1, 3)).expand(-1, 4, 6, 3).to(torch.uint8)
    for r in range(world):
        assert torch.equal(got[r], want)
        assert torch.equal(got_img[r], want_img)
