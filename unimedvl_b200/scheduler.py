"""Continuous batching over the paged KV cache.

The reference decodes one packed batch until SAMPLE 0 emits the end token (bagel.py:1313) and has no scheduler:
a finished sample keeps burning decode steps and a waiting request cannot start until the whole batch ends.  Here
every request owns one engine sequence (a page list), so the set of sequences a decode call runs over can change
between calls:

  * admission  -- waiting requests are prefilled TOGETHER (one packed ViT + image-block forward, one packed prompt
    forward: the same ``prepare_* / forward_cache_update_*`` calls ``Bagel.chat`` makes, bagel.py:1321-1392) as soon as
    a batch slot and their worst-case KV pages are free;
  * decode     -- all running requests advance ``chunk`` steps in one device-resident loop (``umv_generate_text``);
  * retirement -- a request ends at ITS OWN end token or length budget; its pages return to the pool and its slot is
    re-used by the next admission;
  * shared prefixes -- requests naming the same ``prefix`` (e.g. the think-mode system prompt, inferencer.py:574-577)
    fork one prefilled sequence (ref-counted pages, copy-on-write tail) instead of prefilling it again;
  * fused admission -- when every admitted request carries an image, its image block (full mask) and its prompt (causal) are
    prefilled in ONE forward (``umv_forward_cache_update_vit_prompt``): one pass over the weights instead of two;
  * mixed steps -- the admission forwards (image block, prompt) carry the running requests along as DECODE RIDERS
    (``umv_decode_riders``): one query row per running request rides in the same packed forward, so decoding does not stop while
    new requests are prefilled (``mixed=True``, the default).

Samples are independent (SURVEY.md section 8a "Batching semantics verified"), so every request returns exactly the
tokens ``Bagel.chat`` / ``generate_text`` would return for it alone (tests/test_scheduler_gpu.py).
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass, field
from typing import Any, Deque, Dict, List, Optional, Sequence

import torch

from .cache import NaiveCache, PagedKV

_PAGE = 64      # tokens per KV page (include/umv.h: UMV_PAGE_TOKENS)


@dataclass
class Request:
    """One VQA / text request: optional image (PIL, goes through the ViT transform), prompt text, length budget as
    ``generate_text(max_length=...)`` counts it (rows of the returned column, the start token included)."""
    prompt: str
    image: Any = None
    max_length: int = 128
    prefix: Optional[str] = None
    rid: int = -1
    # running state
    seq: int = -1
    kv_len: int = 0
    rope: int = 0
    next_token: int = 0
    inputs: List[int] = field(default_factory=list)
    pages: int = 0


class _Borrowed:
    """A NaiveCache view over sequences the scheduler owns (not freed when the view dies)."""

    def __init__(self, engine, layers: int, seqs: Sequence[int]):
        self.cache = NaiveCache(layers)
        self.cache._umv = PagedKV(engine, seqs=list(seqs))

    def __enter__(self):
        return self.cache

    def __exit__(self, *exc):
        self.cache._umv.seqs = []
        return False


class ContinuousBatcher:
    def __init__(self, model, tokenizer, new_token_ids: Dict[str, int], vit_transform, max_batch: int = 8, chunk: int = 16,
                 end_token_id: Optional[int] = None, max_prefill_tokens: Optional[int] = None, mixed: bool = True,
                 fused_prefill: bool = True):
        self.model, self.engine = model, model.engine
        self.tokenizer, self.tok, self.vit_transform = tokenizer, new_token_ids, vit_transform
        self.max_batch = min(max_batch, self.engine.max_seqs, 64)
        self.chunk = chunk
        self.mixed = mixed
        self.fused_prefill = fused_prefill         # image block + prompt of an admission group in one forward (when every request has an image)
        self._riding: List[Request] = []
        self.eos = new_token_ids["eos_token_id"] if end_token_id is None else end_token_id
        self.max_prefill_tokens = max_prefill_tokens or self.engine.max_tokens
        self.layers = model.config.llm_config.num_hidden_layers
        self.waiting: Deque[Request] = deque()
        self.running: List[Request] = []
        self.finished: Dict[int, torch.Tensor] = {}
        self._prefixes: Dict[str, tuple] = {}           # text -> (seq, kv_len, rope)
        self._next_id = 0
        self.capacity = self.engine.pages_free()
        self.stats = {"prefill_calls": 0, "decode_calls": 0, "decode_steps": 0, "slot_steps_used": 0, "admitted": 0, "rider_steps": 0}

    # ------------------------------------------------------------------ public
    def submit(self, prompt: str, image=None, max_length: int = 128, prefix: Optional[str] = None) -> int:
        if max_length < 1:
            raise ValueError("max_length must be >= 1")
        r = Request(prompt=prompt, image=image, max_length=max_length, prefix=prefix, rid=self._next_id)
        self._next_id += 1
        self.waiting.append(r)
        return r.rid

    def run(self) -> Dict[int, torch.Tensor]:
        """Drain the queue.  Returns {request id: i64 tokens} -- for each request the column ``generate_text`` returns for it
        alone: the start token, then the generated tokens up to (not including) its end token."""
        while self.waiting or self.running:
            self.step()
        out, self.finished = self.finished, {}
        return out

    def close(self):
        for seq, _, _ in self._prefixes.values():
            self.engine.seq_free(seq)
        self._prefixes.clear()

    # ------------------------------------------------------------------ one scheduling iteration
    @torch.no_grad()
    def step(self) -> List[int]:
        self._admit()
        return self._decode_chunk()

    def _rows_and_pages(self, r: Request, prefix_len: int):
        """Worst-case geometry of a request before it is tokenised for real: prefill rows and KV pages."""
        n_txt = len(self.tokenizer.encode(r.prompt)) + 2
        n_img = 0
        if r.image is not None:
            w, h = r.image.size
            tw, th = self.vit_transform.resize_transform.target_size(w, h)
            p = self.model.vit_patch_size
            n_img = (tw // p) * (th // p) + 2
        # a chunk runs min(chunk, longest remaining budget) steps for EVERY running sequence, so a request may grow up to
        # chunk - 1 tokens past its own budget before it is retired; +1: copy-on-write tail of a forked prefix
        total = prefix_len + n_img + n_txt + r.max_length + self.chunk
        rows = n_img + n_txt if self.fused_prefill else max(n_img, n_txt)      # one fused forward, or the larger of the two
        return rows, (total + _PAGE - 1) // _PAGE + 1

    def _prefix(self, text: str):
        if text not in self._prefixes:
            seq = self.engine.seq_new()
            g, lens, rope = self.model.prepare_prompts([0], [0], [text], self.tokenizer, self.tok)
            with _Borrowed(self.engine, self.layers, [seq]) as c:
                self.model.forward_cache_update_text(c, **g)
            self._prefixes[text] = (seq, lens[0], rope[0])
        return self._prefixes[text]

    def _admit(self):
        group: List[Request] = []
        rows = 0
        # pages the running requests may still claim (reserved, not yet held); the pool is re-read at every admission so
        # pages held by other caches on the same engine count (back-pressure instead of a mid-decode "pool exhausted")
        held = lambda r: (r.kv_len + _PAGE - 1) // _PAGE
        committed = sum(max(r.pages - held(r), 0) for r in self.running)
        while self.waiting and len(self.running) + len(group) < self.max_batch:
            r = self.waiting[0]
            plen = self._prefix(r.prefix)[1] if r.prefix else 0
            # free pages right now, capped by this scheduler's own budget (`capacity` minus what its running requests hold)
            avail = min(self.engine.pages_free(), self.capacity - sum(held(q) for q in self.running))
            need_rows, need_pages = self._rows_and_pages(r, plen)
            if need_pages > self.capacity:
                raise MemoryError(f"request {r.rid} needs {need_pages} KV pages, the pool has {self.capacity}")
            if committed + need_pages > avail or (group and rows + need_rows > self.max_prefill_tokens):
                if not self.running and not group:
                    raise MemoryError(f"request {r.rid} needs {need_pages} KV pages, only {avail} are free and nothing is running")
                break
            self.waiting.popleft()
            r.pages = need_pages
            committed += need_pages
            rows += need_rows
            group.append(r)
        if not group:
            return
        m = self.model
        for r in group:
            if r.prefix:
                seq, r.kv_len, r.rope = self._prefix(r.prefix)
                r.seq = self.engine.seq_fork(seq)
            else:
                r.seq, r.kv_len, r.rope = self.engine.seq_new(), 0, 0
        with_img = [r for r in group if r.image is not None]
        L = None
        if self.fused_prefill and len(with_img) == len(group):
            from . import packing
            gv, _, _ = m.prepare_vit_images([r.kv_len for r in group], [r.rope for r in group], [r.image for r in group],
                                            self.vit_transform, self.tok)
            n_img = [int(n) for n in gv["vit_token_seqlens"]]
            L = packing.image_prompt_layout(n_img, [self.tokenizer.encode(r.prompt) for r in group], self.tok,
                                            [r.kv_len for r in group], [r.rope for r in group])
            if sum(L["seq_lens"]) > self.engine.max_tokens:       # the workspace only holds the two forwards one after the other
                L = None
        if L is not None:
            # image block + prompt of every admitted request in ONE forward (umv_forward_cache_update_vit_prompt): one pass over the
            # weights and one rider step instead of two
            riders = self._riders(sum(L["seq_lens"]))
            self.model.rider_tokens = self.engine.forward_cache_update_vit(
                [r.seq for r in group], L["seq_lens"], L["text_ids"], L["text_rows"], gv["packed_vit_tokens"], gv["packed_vit_position_ids"],
                n_img, L["vit_rows"], L["positions"], riders=riders, prompt_lens=L["prompt_lens"])
            self._riders_done(riders)
            self.stats["prefill_calls"] += 1
            lens, ropes = L["kv_lens"], L["rope"]
        else:
            if with_img:
                g, lens, ropes = m.prepare_vit_images([r.kv_len for r in with_img], [r.rope for r in with_img],
                                                      [r.image for r in with_img], self.vit_transform, self.tok)
                riders = self._riders(sum(lens) - sum(r.kv_len for r in with_img))
                with _Borrowed(self.engine, self.layers, [r.seq for r in with_img]) as c:
                    m.forward_cache_update_vit(c, **g, decode_riders=riders)
                self._riders_done(riders)
                for r, l, p in zip(with_img, lens, ropes):
                    r.kv_len, r.rope = l, p
                self.stats["prefill_calls"] += 1
            g, lens, ropes = m.prepare_prompts([r.kv_len for r in group], [r.rope for r in group], [r.prompt for r in group],
                                               self.tokenizer, self.tok)
            riders = self._riders(sum(lens) - sum(r.kv_len for r in group))
            with _Borrowed(self.engine, self.layers, [r.seq for r in group]) as c:
                m.forward_cache_update_text(c, **g, decode_riders=riders)
            self._riders_done(riders)
            self.stats["prefill_calls"] += 1
        for r, l, p in zip(group, lens, ropes):
            r.kv_len, r.rope = l, p
            r.next_token = self.tok["bos_token_id"]             # prepare_start_tokens (bagel.py:1213-1233)
            r.inputs = []
        self.running.extend(group)
        self.stats["admitted"] += len(group)

    # ------------------------------------------------------------------ mixed steps: running requests ride the admission forwards
    def _riders(self, prefill_rows: int):
        """(seqs, tokens, positions) of the running requests for ONE rider step, or None.  A request whose budget is exhausted does not
        ride; the packed forward (prefill rows + riders) must fit the engine's workspace."""
        if not self.mixed:
            return None
        run = [r for r in self.running if len(r.inputs) < r.max_length][:64]
        if not run or prefill_rows + len(run) > self.engine.max_tokens:
            return None
        self._riding = run
        return [r.seq for r in run], [r.next_token for r in run], [r.rope for r in run]

    def _riders_done(self, riders) -> None:
        if riders is None:
            return
        run, self._riding = self._riding, []
        nxt = self.model.rider_tokens.cpu()
        toks = torch.tensor([riders[1]], dtype=torch.int64)          # the tokens that were fed: one executed step per rider
        self.stats["rider_steps"] += len(run)
        self._advance(run, toks, nxt, 1)

    def _decode_chunk(self) -> List[int]:
        if not self.running:
            return []
        run = self.running
        n = min(self.chunk, max(r.max_length - len(r.inputs) for r in run))
        toks, nxt = self.engine.generate_text([r.seq for r in run], [r.next_token for r in run], [r.rope for r in run], n,
                                              return_next=True)
        self.stats["decode_calls"] += 1
        self.stats["decode_steps"] += n * len(run)
        return self._advance(run, toks.cpu(), nxt.cpu(), n)

    def _advance(self, run: List[Request], toks: torch.Tensor, nxt: torch.Tensor, n: int) -> List[int]:
        """Book-keeping after `n` executed decode steps of the requests `run`: toks [n, len(run)] the tokens fed, nxt the token the
        last step computed.  Retires requests at their own end token / budget, returns the retired ids."""
        computed = torch.cat([toks[1:], nxt[None]], dim=0)          # token computed by each executed step, per sample
        done: List[int] = []
        still: List[Request] = []
        for b, r in enumerate(run):
            budget = min(n, r.max_length - len(r.inputs))
            hit = torch.nonzero(computed[:budget, b] == self.eos)
            used = int(hit[0]) + 1 if hit.numel() else budget
            r.inputs.extend(toks[:used, b].tolist())
            self.stats["slot_steps_used"] += used
            if hit.numel() or len(r.inputs) >= r.max_length:
                self.finished[r.rid] = torch.tensor(r.inputs, dtype=torch.int64)
                self.engine.seq_free(r.seq)
                done.append(r.rid)
            else:
                r.next_token = int(nxt[b])
                r.kv_len += n
                r.rope += n
                still.append(r)
        gone = {id(r) for r in run} - {id(r) for r in still}
        self.running = [r for r in self.running if id(r) not in gone]
        return done
