"""Host logic of the continuous batcher (admission under slot / page / prefill-row budgets, chunked decode bookkeeping,
retirement, prefix sharing) on a fake engine: no GPU.  The fake "model" emits token (1000 * request + step), so every
request's expected column is known in closed form; the real-engine version is tests/test_scheduler_gpu.py."""
from types import SimpleNamespace

import pytest
import torch

from unimedvl_b200.scheduler import ContinuousBatcher

TOK = dict(bos_token_id=7, eos_token_id=9, start_of_image=1, end_of_image=2)


class FakeTokenizer:
    def encode(self, text):
        return list(range(len(text)))


class FakeEngine:
    """Sequences are python lists of 'context items'; decode emits 1000 * tag + step where tag = the first prompt's id."""

    def __init__(self, pages=64, max_seqs=16, max_tokens=256):
        self.max_seqs, self.max_tokens, self.total_pages = max_seqs, max_tokens, pages
        self.seqs, self.next, self.tag, self.step = {}, 0, {}, {}
        self.calls = []

    def seq_new(self):
        self.next += 1
        self.seqs[self.next] = 0
        return self.next

    def seq_fork(self, s):
        n = self.seq_new()
        self.seqs[n] = self.seqs[s]
        self.tag[n] = self.tag.get(s)
        return n

    def seq_free(self, s):
        del self.seqs[s]

    def pages_free(self):
        return self.total_pages - sum((n + 63) // 64 for n in self.seqs.values())

    def generate_text(self, seqs, start_tokens, positions, n_steps, return_next=False, **kw):
        assert len(set(seqs)) == len(seqs) and all(s in self.seqs for s in seqs)
        self.calls.append((tuple(seqs), n_steps))
        toks = torch.zeros((n_steps, len(seqs)), dtype=torch.int64)
        nxt = torch.zeros(len(seqs), dtype=torch.int64)
        for b, s in enumerate(seqs):
            assert positions[b] == self.seqs[s], "rope position must follow the context length in this fake"
            cur = start_tokens[b]
            for i in range(n_steps):
                toks[i, b] = cur
                self.step[s] = self.step.get(s, 0) + 1
                cur = 1000 * self.tag[s] + self.step[s]
            nxt[b] = cur
            self.seqs[s] += n_steps
        return toks, nxt


class FakeModel:
    def __init__(self, engine):
        self.engine = engine
        self.config = SimpleNamespace(llm_config=SimpleNamespace(num_hidden_layers=2))
        self.vit_patch_size = 14
        self.prefill_batches = []

    def prepare_prompts(self, curr_kvlens, curr_rope, prompts, tokenizer, new_token_ids):
        n = [len(tokenizer.encode(p)) + 2 for p in prompts]
        return dict(n=n, prompts=prompts), [k + m for k, m in zip(curr_kvlens, n)], [r + m for r, m in zip(curr_rope, n)]

    def forward_cache_update_text(self, cache, n, prompts, decode_riders=None):
        seqs = cache._umv.seqs
        assert len(seqs) == len(n)
        self.prefill_batches.append(len(seqs))
        if decode_riders is not None:           # mixed step: the running requests decode one token inside this prefill forward
            rs, rt, rp = decode_riders
            assert not set(rs) & set(seqs)
            self.rider_tokens = self.engine.generate_text(rs, rt, rp, 1, return_next=True)[1]
            self.rider_calls = getattr(self, "rider_calls", 0) + 1
        for s, m, p in zip(seqs, n, prompts):
            self.engine.seqs[s] += m
            if p.startswith("q"):
                self.engine.tag[s] = int(p[1:])
        return cache


def _expected(i, max_length, eos_at=None):
    col = [TOK["bos_token_id"]] + [1000 * i + k for k in range(1, max_length)]
    return col if eos_at is None else col[:eos_at]


def _batcher(**kw):
    eng = FakeEngine(**{k: kw.pop(k) for k in ("pages", "max_seqs", "max_tokens") if k in kw})
    model = FakeModel(eng)
    return ContinuousBatcher(model, FakeTokenizer(), TOK, None, **kw), eng, model


def test_every_request_gets_its_own_column_and_slots_are_reused():
    cb, eng, model = _batcher(max_batch=3, chunk=4, end_token_id=-1)
    lens = [5, 12, 3, 9, 1, 7, 8]
    ids = [cb.submit(prompt=f"q{i}", max_length=n) for i, n in enumerate(lens)]
    out = cb.run()
    assert sorted(out) == ids
    for i, n in enumerate(lens):
        assert out[i].tolist() == _expected(i, n)
    assert max(len(c[0]) for c in eng.calls) == 3 and not eng.seqs          # never more than max_batch; everything freed
    assert cb.stats["admitted"] == 7 and cb.stats["slot_steps_used"] == sum(lens)
    assert cb.stats["decode_steps"] + cb.stats["rider_steps"] <= sum(lens) + len(lens) * 3
    assert cb.stats["rider_steps"] > 0 and model.rider_calls > 0            # later admissions carried the running requests along
    cb2, eng2, model2 = _batcher(max_batch=3, chunk=4, end_token_id=-1, mixed=False)
    for i, n in enumerate(lens):
        cb2.submit(prompt=f"q{i}", max_length=n)
    out2 = cb2.run()
    assert all(out2[i].tolist() == out[i].tolist() for i in out) and cb2.stats["rider_steps"] == 0
    assert len(model.prefill_batches) > 2 and model.prefill_batches[0] == 3   # first wave packed together, later waves refill


def test_requests_end_at_their_own_end_token():
    cb, eng, _ = _batcher(max_batch=4, chunk=5)
    cb.eos = 2003                                    # request 2 computes it at its 3rd step -> 3 rows; nobody else ever does
    for i in range(4):
        cb.submit(prompt=f"q{i}", max_length=11)
    out = cb.run()
    assert out[2].tolist() == _expected(2, 11, eos_at=3)
    assert all(out[i].tolist() == _expected(i, 11) for i in (0, 1, 3))


def test_page_budget_limits_admission_and_oversized_requests_fail_loudly():
    cb, eng, _ = _batcher(max_batch=8, chunk=8, end_token_id=-1, pages=7)
    for i in range(4):
        cb.submit(prompt=f"q{i}", max_length=100)    # 4 + 100 tokens -> 2 pages + 1 spare = 3 pages: two fit in 7
    peak = 0
    while cb.waiting or cb.running:
        cb.step()
        peak = max(peak, len(cb.running))
    assert peak == 2 and len(cb.finished) == 4
    cb.submit(prompt="q9", max_length=64 * 8)
    with pytest.raises(MemoryError):
        cb.run()


def test_prefill_row_budget_splits_admission_groups():
    cb, eng, model = _batcher(max_batch=8, chunk=8, end_token_id=-1, max_tokens=12)
    for i in range(4):
        cb.submit(prompt=f"q{i}", max_length=4)      # 4 prompt rows each: three fit in 12 rows
    cb.run()
    assert model.prefill_batches[0] == 3 and sum(model.prefill_batches) == 4


def test_shared_prefix_is_prefilled_once_and_forked():
    cb, eng, model = _batcher(max_batch=4, chunk=4, end_token_id=-1)
    for i in range(3):
        cb.submit(prompt=f"q{i}", max_length=6, prefix="system prompt")
    out = cb.run()
    assert model.prefill_batches[0] == 1              # the prefix alone, once
    assert all(out[i].tolist() == _expected(i, 6) for i in range(3))
    assert len(eng.seqs) == 1                         # only the prefix sequence is still held
    cb.close()
    assert not eng.seqs


def test_bad_arguments():
    cb, _, _ = _batcher()
    with pytest.raises(ValueError):
        cb.submit(prompt="q0", max_length=0)
