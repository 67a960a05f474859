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

    for mixed, fused in ((False, False), (True, False), (False, True), (True, True)):
        # mixed=True: the admission forwards (image block / prompt prefill of the NEW requests) carry one decode row per RUNNING request
        # (umv_decode_riders) -- prefill and decode of different requests in one packed forward; same tokens either way
        # fused=True: image block + prompt of an admission group in ONE forward (umv_forward_cache_update_vit_prompt) when every request
        # of the group has an image
        cb = ContinuousBatcher(model, tok, TOK, vit_tf, max_batch=3, chunk=4, end_token_id=eos, mixed=mixed, fused_prefill=fused)
        ids = [cb.submit(**r) for r in reqs]
        got = cb.run()
        assert sorted(got) == ids
        for i, w in zip(ids, want):
            assert torch.equal(got[i], w), (mixed, fused, i, got[i].tolist(), w.tolist())
        st = cb.stats
        assert st["admitted"] == len(reqs) and st["prefill_calls"] > (1 if fused else 2)      # admitted in several waves, not one batch
        assert st["slot_steps_used"] == sum(len(w) for w in want)
        # a slot is held only until its request ends (at most chunk - 1 wasted steps in its last chunk), not until the batch ends
        assert st["decode_steps"] + st["rider_steps"] <= st["slot_steps_used"] + len(reqs) * (4 - 1)
        assert (st["rider_steps"] > 0) == mixed, st
        cb.close()
        gc.collect()
        assert eng.pages_free() == free0


def test_decode_riders_match_a_plain_decode_step(stack):
    """umv_forward_cache_update_{text,vit}_riders: a running request's next token computed inside another request's prefill forward is
    the token a plain decode step gives, and its cache advances by exactly that row."""
    from unimedvl_b200.cache import NaiveCache
    eng, model, dims, vit_tf = stack
    tok = FakeTokenizer()
    L = dims.llm.layers

    def context(prompt):
        g, lens, rope = model.prepare_prompts([0], [0], [prompt], tok, TOK)
        return model.forward_cache_update_text(NaiveCache(L), **g), lens, rope
    ca, la, ra = context("a running request with some context")
    cb_, lb, rb = context("another one, longer than the first request by a few tokens")
    import copy
    # plain decode: 2 steps from the start token
    start = TOK["bos_token_id"]
    fa, fb = copy.deepcopy(ca), copy.deepcopy(cb_)                        # forks: the originals stay at their prefill length
    plain = eng.generate_text(fa._umv.seqs + fb._umv.seqs, [start, start], [ra[0], rb[0]], 2, return_next=True)
    toks, nxt = plain[0].cpu(), plain[1].cpu()
    # the same two steps as riders: step 1 on a text prefill of a third request, step 2 on an image-block prefill of a fourth
    g, _, _ = model.prepare_prompts([0], [0], ["the newcomer's prompt"], tok, TOK)
    riders = ([ca._umv.seqs[0], cb_._umv.seqs[0]], [start, start], [ra[0], rb[0]])
    model.forward_cache_update_text(NaiveCache(L), **g, decode_riders=riders)
    first = model.rider_tokens.cpu()
    assert torch.equal(first, toks[1]), (first.tolist(), toks[1].tolist())
    img = Image.fromarray(synth.synthetic_image(77, 70, 98))
    g, _, _ = model.prepare_vit_images([0], [0], [img], vit_tf, TOK)
    riders = (riders[0], first.tolist(), [ra[0] + 1, rb[0] + 1])
    model.forward_cache_update_vit(NaiveCache(L), **g, decode_riders=riders)
    assert torch.equal(model.rider_tokens.cpu(), nxt), (model.rider_tokens.tolist(), nxt.tolist())
    assert ca._umv.lens() == [la[0] + 2] and cb_._umv.lens() == [lb[0] + 2]


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
