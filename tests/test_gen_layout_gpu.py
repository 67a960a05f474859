"""Gen-mode (Mixture-of-Transformers routed, full-mask) forwards run on SEGREGATED rows inside the engine: generation-expert rows of all
samples first, understanding-expert marker rows after them (csrc/engine.cu: gen_rows_segregate / llm_run), so each expert's linears see
one contiguous slice.  The layout is an internal choice: through the C ABI the packed order of the reference (bagel.py:617-694,
qwen2_navit.py:843-902) goes in and comes out.  These tests compare the segregated forward with the packed one (UMV_GEN_SEG=0: generation
expert over every row, marker rows overwritten by gathered understanding-expert linears) on the same inputs:

* hidden states, K/V written to the cache and flow velocities are BIT-IDENTICAL whenever both layouts take the same kernel family for
  every linear (the token-major tcgen05 linears and the attention kernels are batch-invariant; the cache keeps the packed order);
* with few generation rows the segregated slice drops to the weight-major split-K linears (<= 64 rows), whose fp32 summation order
  differs: agreement to accumulation noise (rel-L2 < 1e-2 on hidden states after all layers)."""
import os

import pytest
import torch

from util import tiny_weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from unimedvl_b200.engine import Engine
    dims, sd, _ = tiny_weights()
    e = Engine(dims, max_tokens=1600, max_seqs=8, kv_pages=160)
    e.load_state_dict(sd)
    e.finalize()
    return e, dims


def _both(fn):
    """fn() under the packed layout and under the segregated one."""
    old = os.environ.get("UMV_GEN_SEG")
    try:
        os.environ["UMV_GEN_SEG"] = "0"
        a = fn()
        os.environ["UMV_GEN_SEG"] = "1"
        b = fn()
    finally:
        if old is None:
            os.environ.pop("UMV_GEN_SEG", None)
        else:
            os.environ["UMV_GEN_SEG"] = old
    return a, b


def _gen_prefill(e, dims, n_lat_per_sample, ctx_lens, seed):
    """Text context, then one gen-mode full-mask forward [marker, latents..., marker] per sample with update_kv (the shape of
    forward_cache_update_vae).  Returns hidden states and every layer's K/V."""
    g = torch.Generator().manual_seed(seed)
    D = dims.llm.hidden
    seqs = [e.seq_new() for _ in ctx_lens]
    ctx = (torch.randn(sum(ctx_lens), D, generator=g) * 0.5).bfloat16()
    e.llm_forward(ctx, seqs, ctx_lens, [j for n in ctx_lens for j in range(n)], want_hidden=False)
    q_lens = [n + 2 for n in n_lat_per_sample]
    M = sum(q_lens)
    x = (torch.randn(M, D, generator=g) * 0.5).bfloat16()
    is_gen, pos = [], []
    for n, c in zip(n_lat_per_sample, ctx_lens):
        is_gen += [0] + [1] * n + [0]
        pos += [c] * (n + 2)
    h = e.llm_forward(x, seqs, q_lens, pos, row_is_gen=is_gen, is_causal=False, update_kv=True).cpu()
    kv = [tuple(t.cpu() for t in e.seq_export(s, li)) for s in seqs for li in range(dims.llm.layers)]
    for s in seqs:
        e.seq_free(s)
    return h, kv


def test_segregated_gen_prefill_is_bit_identical_to_packed(eng):
    e, dims = eng
    (h0, kv0), (h1, kv1) = _both(lambda: _gen_prefill(e, dims, [256, 144, 256], [34, 7, 70], seed=1))
    assert torch.isfinite(h0.float()).all()
    assert torch.equal(h0, h1), f"hidden states differ: {(h0.float() - h1.float()).abs().max()}"
    for (k0, v0), (k1, v1) in zip(kv0, kv1):
        assert torch.equal(k0, k1) and torch.equal(v0, v1)


def test_segregated_gen_prefill_few_rows_within_accumulation_noise(eng):
    e, dims = eng
    (h0, kv0), (h1, kv1) = _both(lambda: _gen_prefill(e, dims, [16, 9], [12, 40], seed=2))     # 25 generation rows: weight-major linears
    rel = ((h0.float() - h1.float()).norm() / h0.float().norm()).item()
    assert rel < 1e-2, rel
    for (k0, v0), (k1, v1) in zip(kv0, kv1):
        assert ((k0.float() - k1.float()).norm() / k0.float().norm()).item() < 1e-2
        assert ((v0.float() - v1.float()).norm() / v0.float().norm()).item() < 1e-2


def test_gen_forward_one_expert_only_unchanged(eng):
    """All rows on the generation expert (no marker rows): nothing to segregate, both settings run the same path."""
    e, dims = eng
    D = dims.llm.hidden

    def run():
        s = e.seq_new()
        x = (torch.randn(130, D, generator=torch.Generator().manual_seed(3)) * 0.5).bfloat16()
        h = e.llm_forward(x, [s], [130], [0] * 130, row_is_gen=[1] * 130, is_causal=False, update_kv=False).cpu()
        e.seq_free(s)
        return h
    a, b = _both(run)
    assert torch.equal(a, b)


@pytest.mark.parametrize("renorm", [0, 1, 2])
def test_flow_velocity_segregated_is_bit_identical_to_packed(eng, renorm):
    """Three CFG branches x 2 images x 256 latent tokens: 1,548 packed rows, 1,536 + 12 segregated."""
    e, dims = eng
    g = torch.Generator().manual_seed(4)
    D, C = dims.llm.hidden, dims.patch_latent_dim
    ctx_lens = [30, 18]
    main = [e.seq_new() for _ in ctx_lens]
    cfg_text = [e.seq_new() for _ in ctx_lens]
    ctx = (torch.randn(sum(ctx_lens), D, generator=g) * 0.5).bfloat16()
    e.llm_forward(ctx, main, ctx_lens, [j for n in ctx_lens for j in range(n)], want_hidden=False)
    cfg_img = [e.seq_fork(s) for s in main]
    lat_lens = [256, 256]
    x_t = torch.randn(sum(lat_lens), C, generator=g).cuda().contiguous()
    lat_pos = torch.cat([torch.arange(n) for n in lat_lens])
    markers = [5, 6]                                                     # start / end-of-image ids inside the tiny vocabulary

    def run():
        return e.flow_velocity(x_t, lat_pos, lat_lens, main, ctx_lens, markers, 0.7, cfg_text=(cfg_text, [0, 0]),
                               cfg_img=(cfg_img, ctx_lens), cfg_text_scale=4.0, cfg_img_scale=1.5, renorm_type=renorm).cpu().clone()
    v0, v1 = _both(run)
    assert torch.isfinite(v0).all() and v0.abs().max() > 0
    assert torch.equal(v0, v1), f"velocity differs: {(v0 - v1).abs().max()}"
    assert [e.seq_len(s) for s in main] == ctx_lens                     # update_past_key_values=False: caches untouched
    for s in main + cfg_text + cfg_img:
        e.seq_free(s)


def test_cfg_branch_over_the_same_context_is_evaluated_once(eng, monkeypatch):
    """Pure text-to-image: the image-free CFG context holds the main context's tokens (inferencer.py:578-600), so the cfg_img velocity IS
    the main one.  umv_flow_velocity recognises a fork of the main context (same content id, length, rope position) and runs two branches
    instead of three; the guided velocity is bit-identical to the three-branch evaluation (UMV_CFG_DEDUP=0), the third branch's rows are not computed.  A context that received its own (identical) prefill, or differs in length, is NOT deduplicated."""
    e, dims = eng
    g = torch.Generator().manual_seed(5)
    D, C = dims.llm.hidden, dims.patch_latent_dim
    ctx_lens = [30, 18]
    ctx = (torch.randn(sum(ctx_lens), D, generator=g) * 0.5).bfloat16()
    pos0 = [j for n in ctx_lens for j in range(n)]
    main = [e.seq_new() for _ in ctx_lens]
    e.llm_forward(ctx, main, ctx_lens, pos0, want_hidden=False)
    fork = [e.seq_fork(s) for s in main]
    own = [e.seq_new() for _ in ctx_lens]                     # same tokens, prefilled separately: equal K / V, but not known to be
    e.llm_forward(ctx, own, ctx_lens, pos0, want_hidden=False)
    empty = [e.seq_new() for _ in ctx_lens]
    lat_lens = [256, 256]
    x_t = torch.randn(sum(lat_lens), C, generator=g).cuda().contiguous()
    lat_pos = torch.cat([torch.arange(n) for n in lat_lens])

    def run(img, dedup):
        monkeypatch.setenv("UMV_CFG_DEDUP", dedup)
        v = e.flow_velocity(x_t, lat_pos, lat_lens, main, ctx_lens, [5, 6], 0.7, cfg_text=(empty, [0, 0]), cfg_img=(img, ctx_lens),
                            cfg_text_scale=4.0, cfg_img_scale=1.5, renorm_type=0).cpu().clone()
        return v, e.flow_branches_last()
    v3, n3 = run(fork, "0")
    v2, n2 = run(fork, "1")
    assert (n3, n2) == (3, 2) and torch.equal(v2, v3)
    vo, no = run(own, "1")
    assert no == 3 and torch.equal(vo, v3)                     # not deduplicated; equal all the same (deterministic kernels)
    vs, ns = run(main, "1")                                     # the main context itself as the image-free context
    assert ns == 2 and torch.equal(vs, v3)
    e.llm_forward(ctx[:1], [fork[0]], [1], [ctx_lens[0]], want_hidden=False)      # the fork moves on: no longer the main context
    assert run(fork, "1")[1] == 3
    for s in main + fork + own + empty:
        e.seq_free(s)
