"""Direct op-level parity of the two production attention kernels against the oracle, at the 14B head geometry
(28 query / 4 kv heads x 128, GQA group 7) and the contexts of the BASELINE configs (64 ... 1570 keys, ragged):

  * `attn_tc_kernel`     (tcgen05 flash attention: image / text prefill, flow forwards)         -> path 2
  * `attn_decode_kernel` (fused split-K reduce + q/k-norm + RoPE + KV append + split-KV attention
                          + DSMEM combine, one cluster launch per layer)                         -> path 3
  * `attn_fwd_kernel`    (mma.sync flash kernel: short queries, fall-backs)                      -> path 1

through `umv_op_attention_block`, which runs exactly the attention block of the model path (q/k RMSNorm, RoPE, KV append into
the paged pool, attention over past + new keys) on caller-provided projection outputs.  The oracle side is the restatement
of PackedAttentionMoT.forward_inference (qwen2_navit.py:544-614): oracle.numerics rmsnorm / apply_rope / attention_varlen.

Bars: the appended K/V rows are elementwise work (R4/R5 rounding chain: fp32 sum-of-squares order and cos/sin evaluation are the
only freedom) -> equal except <= 2 ulp on < 0.1 % of the elements (> 1 ulp on < 1e-5); attention
outputs: both sides round P to bf16 and accumulate in fp32, they differ in the order of the online-softmax rescaling ->
rel-L2 < 4e-3 (the bound test_attention_varlen already uses for the mma.sync kernel).
"""
import pytest
import torch

from oracle import numerics as nm
from util import ulp_stats

pytestmark = pytest.mark.gpu

H, HKV, DH = 28, 4, 128
QN = (H + 2 * HKV) * DH


@pytest.fixture(scope="module")
def eng():
    from unimedvl_b200 import config as ucfg
    from unimedvl_b200.engine import Engine
    dims = ucfg.BagelDims(llm=ucfg.LLMDims(hidden=H * DH, heads=H, kv_heads=HKV, inter=128, layers=1, vocab=1024),
                          vit=ucfg.ViTDims(hidden=144, heads=2, inter=328, layers=1))
    e = Engine(dims, max_tokens=4200, max_seqs=16, kv_pages=16 * 26 + 32, enable_vit=False, enable_gen=True)
    e.fill_synthetic(seed=11)
    e.finalize()
    p = "language_model.model.layers.0.self_attn."
    w = {n: e.export_tensor(p + n + ".weight", (DH,)) for n in ("q_norm", "k_norm", "q_norm_moe_gen", "k_norm_moe_gen")}
    yield e, dims, w
    e.close()


def _oracle_block(qkv, positions, w, past_k, past_v, q_lens, causal, is_gen=None):
    """q/k-norm + RoPE (und: bf16 chain R4/R5; gen rows: fp32, one rounding) -> new K/V rows; merged K/V per sample; attention."""
    M = qkv.shape[0]
    q = qkv[:, :H * DH].view(M, H, DH)
    k = qkv[:, H * DH:(H + HKV) * DH].view(M, HKV, DH)
    v = qkv[:, (H + HKV) * DH:].view(M, HKV, DH)
    cos, sin = nm.rope_cos_sin(torch.as_tensor(positions), nm.default_inv_freq(DH, 1e6), torch.bfloat16)
    if is_gen is None:
        qn, kn = nm.rmsnorm(q, w["q_norm"], 1e-6), nm.rmsnorm(k, w["k_norm"], 1e-6)
    else:                                   # mode "gen": everything up-cast, rows routed to the expert's norm weights
        g = torch.as_tensor(is_gen, dtype=torch.bool)
        qf, kf = q.float(), k.float()
        qn, kn = qf.clone(), kf.clone()
        qn[~g], kn[~g] = nm.rmsnorm(qf[~g], w["q_norm"], 1e-6), nm.rmsnorm(kf[~g], w["k_norm"], 1e-6)
        qn[g], kn[g] = nm.rmsnorm(qf[g], w["q_norm_moe_gen"], 1e-6), nm.rmsnorm(kf[g], w["k_norm_moe_gen"], 1e-6)
    qr, kr = nm.apply_rope(qn, kn, cos, sin)
    qr, kr = qr.bfloat16(), kr.bfloat16()
    mk, mv, klens, o = [], [], [], 0
    for b, n in enumerate(q_lens):
        mk += [past_k[b], kr[o:o + n]]
        mv += [past_v[b], v[o:o + n]]
        klens.append(past_k[b].shape[0] + n)
        o += n
    out = nm.attention_varlen(qr, torch.cat(mk), torch.cat(mv), list(q_lens), klens, causal)
    return out.reshape(M, H * DH), kr, v


def _context(e, lens, seed):
    """Sequences holding `lens` random keys/values in layer 0 (written through the same entry point, un-normed inputs do not
    matter: what was stored is exported and handed to the oracle as the past)."""
    seqs = [e.seq_new() for _ in lens]
    g = torch.Generator().manual_seed(seed)
    for s, n in zip(seqs, lens):            # one sequence per call: the packed rows of a call are bounded by max_tokens
        if n > 0:
            x = torch.randn(n, QN, generator=g).bfloat16()
            e.attention_block(0, [s], [n], list(range(n)), qkv=x, is_causal=True, update_kv=True)
    past = [e.seq_export(s, 0) if n > 0 else (torch.zeros(0, HKV, DH, dtype=torch.bfloat16), torch.zeros(0, HKV, DH, dtype=torch.bfloat16))
            for s, n in zip(seqs, lens)]
    return seqs, [p[0].cpu() for p in past], [p[1].cpu() for p in past]


def _free(e, seqs):
    for s in seqs:
        e.seq_free(s)


@pytest.mark.parametrize("name,past,q_lens,causal", [
    ("image prefill, empty cache, full mask", [0, 0, 0], [1026, 730, 258], False),
    ("text prefill on an image context, causal", [1026, 730, 64], [32, 13, 130], True),
    ("second image on a context, full mask", [1058, 300], [1026, 512], False),
    ("long causal chunk", [0], [1570], True),
])
def test_prefill_attention_tcgen05_vs_oracle(eng, name, past, q_lens, causal):
    e, dims, w = eng
    seqs, pk, pv = _context(e, past, seed=len(name))
    g = torch.Generator().manual_seed(7 + len(name))
    M = sum(q_lens)
    qkv = torch.randn(M, QN, generator=g).bfloat16()
    pos = [past[b] // 2 + (j if causal else 0) for b, n in enumerate(q_lens) for j in range(n)]
    out, path = e.attention_block(0, seqs, q_lens, pos, qkv=qkv, is_causal=causal, update_kv=True)
    want, kr, v = _oracle_block(qkv, pos, w, pk, pv, q_lens, causal)
    big = [n * (H // HKV) >= 128 for n in q_lens]
    assert path == (2 if all(big) or any(big) else 1), (name, path)       # the tcgen05 kernel takes calls with 128-row tiles
    s = ulp_stats(out, want)
    assert torch.isfinite(out.float()).all() and s["rel_l2"] < 4e-3, (name, path, s)
    # the appended rows: q/k-norm + RoPE rounding chain, elementwise
    o = 0
    for b, n in enumerate(q_lens):
        k_all, v_all = e.seq_export(seqs[b], 0)
        sk = ulp_stats(k_all[past[b]:], kr[o:o + n])
        assert sk["max_ulp"] <= 2 and sk["frac"] < 1e-3 and sk["frac_gt1"] < 1e-5, (name, b, sk)
        assert torch.equal(v_all[past[b]:].cpu(), v[o:o + n])
        assert torch.equal(k_all[:past[b]].cpu(), pk[b]) and torch.equal(v_all[:past[b]].cpu(), pv[b])      # the past is untouched
        o += n
    _free(e, seqs)


@pytest.mark.parametrize("early", ["0", "2"])
@pytest.mark.parametrize("ctx,splits", [([1058] * 8, 3), ([1570, 1185, 1058, 777, 513, 512, 65, 64], 2), ([1569] * 16, 0), ([63, 1], 1)])
def test_decode_attention_fused_vs_oracle(eng, ctx, splits, early, monkeypatch):
    """One query token per sample on top of `ctx` cached keys -- config 2 (ctx 1058 -> 1185) and config 3 (16 per GPU,
    ctx -> 1569) geometries, ragged, page-boundary lengths -- with the projection outputs arriving as split-K partials + bias
    (what the decode step's weight-major linear produces) or as bf16 rows (splits == 0)."""
    # early = "2": the step state is read and the K/V tiles are requested BEFORE griddepcontrol.wait (what layers >= 1 of the decode
    # step do); "0": after it (layer 0)
    monkeypatch.setenv("UMV_ATTN_EARLY", early)
    e, dims, w = eng
    B = len(ctx)
    seqs, pk, pv = _context(e, ctx, seed=B)
    g = torch.Generator().manual_seed(31 + B)
    pos = [c // 3 + 5 for c in ctx]
    if splits:
        part = torch.randn(splits, B, QN, generator=g) * 0.7
        bias = (torch.randn(QN, generator=g) * 0.3).bfloat16()
        acc = torch.zeros(B, QN)
        for s_ in range(splits):
            acc = acc + part[s_]                  # the kernel sums the partials in index order, then adds the bias, one rounding
        qkv = (acc + bias.float()).bfloat16()
        out, path = e.attention_block(0, seqs, [1] * B, pos, partial=part, bias=bias, is_causal=True, update_kv=True)
    else:
        qkv = torch.randn(B, QN, generator=g).bfloat16()
        out, path = e.attention_block(0, seqs, [1] * B, pos, qkv=qkv, is_causal=True, update_kv=True)
    assert path == 3, path
    want, kr, v = _oracle_block(qkv, pos, w, pk, pv, [1] * B, True)
    s = ulp_stats(out, want)
    assert torch.isfinite(out.float()).all() and s["rel_l2"] < 4e-3, s
    for b in range(B):
        k_all, v_all = e.seq_export(seqs[b], 0)
        assert k_all.shape[0] == ctx[b] + 1
        sk = ulp_stats(k_all[-1:], kr[b:b + 1])
        assert sk["max_ulp"] <= 1, (b, sk)
        assert torch.equal(v_all[-1:].cpu(), v[b:b + 1])
        assert torch.equal(k_all[:-1].cpu(), pk[b])
    _free(e, seqs)


def test_decode_attention_unfused_vs_oracle(eng, monkeypatch):
    """The three-kernel decode path (rope_append + split-KV mma.sync attention + combine) the fused kernel replaced."""
    e, dims, w = eng
    monkeypatch.setenv("UMV_FUSED_ATTN", "0")
    ctx = [1185, 1058, 300, 64]
    seqs, pk, pv = _context(e, ctx, seed=4)
    qkv = torch.randn(len(ctx), QN, generator=torch.Generator().manual_seed(2)).bfloat16()
    pos = [40, 41, 42, 43]
    out, path = e.attention_block(0, seqs, [1] * len(ctx), pos, qkv=qkv, is_causal=True, update_kv=False)
    assert path == 1
    want, _, _ = _oracle_block(qkv, pos, w, pk, pv, [1] * len(ctx), True)
    s = ulp_stats(out, want)
    assert s["rel_l2"] < 4e-3, s
    assert [e.seq_len(s_) for s_ in seqs] == ctx                       # update_kv=False: lengths unchanged
    _free(e, seqs)


def test_flow_attention_gen_mode_vs_oracle(eng):
    """A flow forward's attention: 3 CFG branches x 258 rows (2 marker rows through the understanding q/k norms, 256 latent rows
    through *_moe_gen), fp32 norm + RoPE with ONE rounding (qwen2_navit.py:568-583), full mask over context + block, no cache update."""
    e, dims, w = eng
    past = [34, 0, 34]
    seqs, pk, pv = _context(e, past, seed=9)
    q_lens = [258] * 3
    g = torch.Generator().manual_seed(12)
    qkv = torch.randn(sum(q_lens), QN, generator=g).bfloat16()
    is_gen = ([0] + [1] * 256 + [0]) * 3
    pos = [p for b in range(3) for p in [past[b]] * 258]
    out, path = e.attention_block(0, seqs, q_lens, pos, qkv=qkv, row_is_gen=is_gen, is_causal=False, update_kv=False)
    assert path == 2
    want, _, _ = _oracle_block(qkv, pos, w, pk, pv, q_lens, False, is_gen=is_gen)
    s = ulp_stats(out, want)
    assert s["rel_l2"] < 4e-3, s
    assert [e.seq_len(s_) for s_ in seqs] == past
    _free(e, seqs)
