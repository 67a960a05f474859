"""Continuous batching (unimedvl_b200/scheduler.py): requests admitted into free slots,
decoded in chunks, retired at their own end token -- every request must return exactly the tokens it gets when it runs
ALONE through the reference-shaped calls (prepare_* / forward_cache_update_* / generate_text, the sequence Bagel.chat
makes, bagel.py:1321-1392), and every KV page must come back."""
import gc

import pytest
import torch
from PIL import Image

from unimedvl_b200 import synth
from util import TOK, tiny_weights

pytestmark = pytest.mark.gpu


class FakeTokenizer:
    def encode(self, text):
        return [(ord(c) * 7 + i * 13) % 2000 for i, c in enumerate(text)]


@pytest.fixture(scope="module")
def stack():
    from unimedvl_b200.bagel import Bagel
    from unimedvl_b200.engine import Engine
    from unimedvl_b200.packing import ImageTransform
    dims, sd, _ = tiny_weights()
    eng = Engine(dims, max_tokens=512, max_seqs=16, kv_pages=64)
    eng.load_state_dict(sd)
    eng.finalize()
    return eng, Bagel(eng, dims), dims, ImageTransform(980, 28, 14)


def _requests():
    sizes = [(70, 98), (56, 56), None, (84, 70), (70, 98), None, (56, 84), (98, 98), (70, 70)]
    reqs = []
    for i, sz in enumerate(sizes):
        img = Image.fromarray(synth.synthetic_image(60 + i, *sz)) if sz else None
        reqs.append(dict(prompt=f"Question number {i}: what does region {i * 3} show?" + " more" * (i % 3), image=img,
                         max_length=[12, 7, 9, 15, 5, 11, 8, 13, 6][i], prefix="You are a careful radiologist." if i % 4 == 1 else None))
    return reqs


def _solo(model, dims, vit_tf, tok, r):
    """The request alone, free-running to its length budget."""
    from unimedvl_b200.cache import NaiveCache
    cache, lens, rope = NaiveCache(dims.llm.layers), [0], [0]
    if r["prefix"]:
        g, lens, rope = model.prepare_prompts(lens, rope, [r["prefix"]], tok, TOK)
        cache = model.forward_cache_update_text(cache, **g)
    if r["image"] is not None:
        g, lens, rope = model.prepare_vit_images(lens, rope, [r["image"]], vit_tf, TOK)
        cache = model.forward_cache_update_vit(cache, **g)
    g, lens, rope = model.prepare_prompts(lens, rope, [r["prompt"]], tok, TOK)
    cache = model.forward_cache_update_text(cache, **g)
    g = model.prepare_start_tokens(lens, rope, TOK)
    return model.generate_text(past_key_values=cache, max_length=r["max_length"], end_token_id=None, **g).cpu()[:, 0]


def test_continuous_batching_matches_solo_runs(stack):
    from unimedvl_b200.scheduler import ContinuousBatcher
    eng, model, dims, vit_tf = stack
    tok = FakeTokenizer()
    reqs = _requests()
    gc.collect()
    free0 = eng.pages_free()
    solo = [_solo(model, dims, vit_tf, tok, r) for r in reqs]
    # an end token that really occurs: the most frequent generated token ends several requests early, others run out of budget
    gen = torch.cat([s[1:] for s in solo])
    vals, counts = torch.unique(gen, return_counts=True)
    eos = int(torch.mode(gen).values)
    for t in vals[torch.argsort(counts, descending=True)].tolist():          # prefer one that some requests never produce
        n_with = sum(bool((s[1:] == t).any()) for s in solo)
        if 2 <= n_with <= len(solo) - 2:
            eos = t
            break
    want = []
    for s in solo:
        hit = torch.nonzero(s[1:] == eos)
        want.append(s[:int(hit[0]) + 1] if hit.numel() else s)
    assert any(len(w) < len(s) for w, s in zip(want, solo))
    gc.collect()
    assert eng.pages_free() == free0

    cb = ContinuousBatcher(model, tok, TOK, vit_tf, max_batch=3, chunk=4, end_token_id=eos)
    ids = [cb.submit(**r) for r in reqs]
    got = cb.run()
    assert sorted(got) == ids
    for i, w in zip(ids, want):
        assert torch.equal(got[i], w), (i, got[i].tolist(), w.tolist())
    st = cb.stats
    assert st["admitted"] == len(reqs) and st["prefill_calls"] > 2          # admitted in several waves, not one batch
    assert st["slot_steps_used"] == sum(len(w) for w in want)
    # a slot is held only until its request ends (at most chunk - 1 wasted steps in its last chunk), not until the batch ends
    assert st["decode_steps"] <= st["slot_steps_used"] + len(reqs) * (4 - 1)
    cb.close()
    gc.collect()
    assert eng.pages_free() == free0


def test_admission_respects_the_page_pool(stack):
    from unimedvl_b200.scheduler import ContinuousBatcher
    eng, model, dims, vit_tf = stack
    tok = FakeTokenizer()
    cb = ContinuousBatcher(model, tok, TOK, vit_tf, max_batch=8, chunk=8, end_token_id=-1)
    cb.capacity = 6                                  # pretend the pool is tiny: 3 pages per request -> two run at a time
    ids = [cb.submit(prompt=f"prompt {i}", max_length=70) for i in range(5)]
    peak = 0
    while cb.waiting or cb.running:
        cb.step()
        peak = max(peak, len(cb.running))
    assert peak == 2 and sorted(cb.finished) == ids and all(len(v) == 70 for v in cb.finished.values())
    with pytest.raises(MemoryError):
        cb.submit(prompt="x", max_length=64 * 6)
        cb.run()
