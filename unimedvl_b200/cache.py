"""KV-cache objects of the drop-in boundary.

The reference keeps the cache in ``NaiveCache`` (qwen2_navit.py:207-221): per layer one packed K and
V tensor that is re-materialised on every forward, and forks contexts with ``copy.deepcopy``
(inferencer.py:261,587,600,607).  Here the tensors live in the engine's page pool; a ``PagedKV``
handle (one engine sequence per sample) rides on the cache object and implements ``__deepcopy__`` as a
ref-counted page fork, so the reference's own ``InterleaveInferencer`` code keeps working unchanged.
"""
from __future__ import annotations

from typing import List, Optional


class PagedKV:
    """Engine sequences of one packed batch.  Freed when garbage collected."""

    def __init__(self, engine, n: int = 0, seqs: Optional[List[int]] = None):
        self.engine = engine
        self.seqs = list(seqs) if seqs is not None else [engine.seq_new() for _ in range(n)]

    def lens(self) -> List[int]:
        return [self.engine.seq_len(s) for s in self.seqs]

    def fork(self) -> "PagedKV":
        return PagedKV(self.engine, seqs=[self.engine.seq_fork(s) for s in self.seqs])

    def __deepcopy__(self, memo):
        return self.fork()

    def free(self):
        eng, seqs, self.seqs = self.engine, self.seqs, []
        if eng is not None and getattr(eng, "h", None):
            for s in seqs:
                try:
                    eng.seq_free(s)
                except Exception:
                    pass

    def __del__(self):
        self.free()


class _LayerView:
    """key_cache / value_cache mapping: layer -> packed [sum_ctx, kv_heads, head_dim] tensor, exported on
    access (parity tests); ``None`` for an empty cache exactly as the reference."""

    def __init__(self, cache: "NaiveCache", which: int):
        import weakref
        self._ref, self._which = weakref.ref(cache), which      # no reference cycle: pages are freed by refcount

    @property
    def _cache(self):
        return self._ref()

    def __getitem__(self, layer: int):
        h = self._cache._umv
        if h is None or not h.seqs or sum(h.lens()) == 0:
            return None
        import torch
        parts = [h.engine.seq_export(s, layer)[self._which] for s in h.seqs]
        return torch.cat(parts, dim=0)

    def __len__(self):
        return self._cache._num_layers

    def keys(self):
        return range(self._cache._num_layers)


class NaiveCache:
    """Same constructor and attributes as the reference's NaiveCache (qwen2_navit.py:207-221)."""

    def __init__(self, num_layers: int):
        self._num_layers = num_layers
        self._umv: Optional[PagedKV] = None
        self.key_cache = _LayerView(self, 0)
        self.value_cache = _LayerView(self, 1)

    @property
    def num_layers(self) -> int:
        return self._num_layers

    @property
    def seq_lens(self) -> int:
        return 0 if self._umv is None else sum(self._umv.lens())

    def __deepcopy__(self, memo):
        c = NaiveCache(self._num_layers)
        c._umv = None if self._umv is None else self._umv.fork()
        return c

    @classmethod
    def concat(cls, caches) -> "NaiveCache":
        """One cache over the samples of several (batch order = argument order).  The sequences MOVE: the inputs are left empty.
        Lets a large batch be prefilled in chunks that fit the workspace and decoded as one."""
        caches = list(caches)
        out = cls(caches[0]._num_layers)
        seqs, engine = [], None
        for c in caches:
            if c._umv is not None:
                engine = c._umv.engine
                seqs += c._umv.seqs
                c._umv.seqs = []
        if engine is not None:
            out._umv = PagedKV(engine, seqs=seqs)
        return out


def paged_handle(past_key_values, engine, n_seqs: int) -> PagedKV:
    """PagedKV attached to a cache object (ours or the reference's NaiveCache); created on first use."""
    h = getattr(past_key_values, "_umv", None)
    if h is None or (not h.seqs and n_seqs > 0):
        h = PagedKV(engine, n_seqs)
        past_key_values._umv = h
    if len(h.seqs) != n_seqs:
        raise ValueError(f"cache holds {len(h.seqs)} samples, call has {n_seqs}")
    return h
