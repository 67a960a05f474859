"""Op-level parity on the B200, through the C ABI (umv_op_*): every kernel the model path launches vs the
oracle / a plain torch fp32 restatement of the same op.  Tolerances: elementwise ops (norms, epilogue math)
are bit-exact up to rare 1-ulp flips from fp32 reduction order; contractions accumulate in fp32 in a
different order than the oracle, so results agree to <= 1 bf16 ulp on all but a small fraction."""
import pytest
import torch

from oracle import numerics as nm
from util import ulp_stats

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from unimedvl_b200 import engine
    return engine


def _ref_linear(x, w, b, res, epi):
    M, N = x.shape[0], w.shape[0]
    acc = x.float() @ w.float().T
    if epi == 2:
        v = acc.view(M, N // 128, 2, 64)
        g, u = v[:, :, 0].reshape(M, N // 2).bfloat16(), v[:, :, 1].reshape(M, N // 2).bfloat16()
        return torch.nn.functional.silu(g) * u
    y = (acc + b.float()).bfloat16()
    if epi == 1:
        y = torch.nn.functional.gelu(y, approximate="tanh")
    if epi == 3:
        y = y + res
    return y


@pytest.mark.parametrize("M,N,K", [(8, 4608, 3584), (1, 896, 896), (16, 3584, 3584), (33, 1152, 896), (64, 2048, 896),
                                   (8, 3584, 18944), (130, 896, 64), (300, 3456, 1152), (1000, 1152, 4304), (70, 64, 896),
                                   (1026, 3072, 896), (257, 432, 144)])
def test_linear_all_paths(ops, M, N, K):
    torch.manual_seed(M * 7 + N)
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    b = (torch.randn(N, device="cuda") * 0.1).bfloat16()
    res = torch.randn(M, N, device="cuda").bfloat16()
    for epi in (0, 1, 3, 2):
        if epi == 2 and N % 128:
            continue
        ref = _ref_linear(x, w, b, res, epi)
        for impl in (3, 1, 2):                    # simple CUDA-core kernel, tcgen05 token-major, tcgen05 weight-major
            if impl == 2 and M > 64:
                continue
            y = ops.op_linear(x, w, None if epi == 2 else b, res if epi == 3 else None, epi=epi, impl=impl)
            s = ulp_stats(y, ref)
            assert not torch.isnan(y.float()).any()
            # swiglu multiplies two rounded values (and K up to 18944 widens the fp32 order noise): looser ulp budget
            assert s["rel_l2"] < 1e-3 and s["frac_gt1"] < (1e-2 if epi == 2 else 2e-3), (M, N, K, epi, impl, s)


@pytest.mark.parametrize("M,N,K,epi", [(2100, 3584, 2048, 3), (2100, 4096, 1024, 2), (1500, 1152, 4304, 0), (700, 2048, 896, 1)])
def test_linear_tile_order_never_changes_a_bit(ops, M, N, K, epi, monkeypatch):
    """The token-major linears sweep the column tiles inside groups of row tiles (csrc/gemm.cu: pick_tile_group / tile_ab,
    csrc/gemm_2cta.cu: tile_coords).  The order only decides WHEN a tile is computed: every forced group size, the automatic choice and
    the plain all-rows order give the same bytes, on the pair kernel and on the single-CTA kernel, with ragged last tiles."""
    torch.manual_seed(M + N)
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    b = None if epi == 2 else (torch.randn(N, device="cuda") * 0.1).bfloat16()
    res = torch.randn(M, N, device="cuda").bfloat16() if epi == 3 else None
    outs = {}
    for pair in ("1", "0"):
        monkeypatch.setenv("UMV_2CTA", pair)
        for g in ("100000", "1", "2", "3", "5", None):
            if g is None:
                monkeypatch.delenv("UMV_RASTER_G", raising=False)
            else:
                monkeypatch.setenv("UMV_RASTER_G", g)
            outs[pair, g] = ops.op_linear(x, w, b, res, epi=epi, impl=1).clone()
    ref = outs["1", "100000"]
    assert torch.isfinite(ref.float()).all()
    assert ulp_stats(ref, _ref_linear(x, w, b if b is not None else None, res, epi))["rel_l2"] < 1e-3
    for key, y in outs.items():
        assert torch.equal(y, ref), key


def test_linear_rejects_unaligned_rows(ops):
    x = torch.randn(4, 588, device="cuda").bfloat16()
    w = torch.randn(64, 588, device="cuda").bfloat16()
    y = ops.op_linear(x, w, impl=0)              # auto: falls to the CUDA-core kernel (588*2 B rows are not TMA-addressable)
    assert ulp_stats(y, (x.float() @ w.float().T).bfloat16())["rel_l2"] < 1e-3
    with pytest.raises(ValueError):
        ops.op_linear(x, w, impl=1)


@pytest.mark.parametrize("M,D", [(8, 3584), (100, 896), (3, 128), (1, 8), (17, 1152), (50, 144)])
def test_norms(ops, M, D):
    torch.manual_seed(D)
    x = torch.randn(M, D).bfloat16()
    w = (1 + 0.1 * torch.randn(D)).bfloat16()
    b = (0.1 * torch.randn(D)).bfloat16()
    s = ulp_stats(ops.op_rmsnorm(x.cuda(), w.cuda()), nm.rmsnorm(x, w, 1e-6))
    assert s["max_ulp"] <= 1 and s["frac"] < 1e-3, s
    s = ulp_stats(ops.op_layernorm(x.cuda(), w.cuda(), b.cuda()), nm.layernorm(x, w, b, 1e-6, nm.Semantics.cuda).bfloat16())
    assert s["frac"] < 1e-3 and s["rel_l2"] < 1e-4, s


@pytest.mark.parametrize("M,D", [(4100, 3584), (300, 2048), (1027, 1152), (515, 4096)])
def test_rmsnorm_row_kernels_bit_identical(ops, M, D, monkeypatch):
    """The warp-per-row kernel (>= 4096 rows) and the 128-threads-per-row kernel (256 .. 4095 rows) sum the squares in the block
    kernel's order: whichever runs, the bits are the same -- batch invariance (tests/test_fullsize_gpu.py) rests on it."""
    torch.manual_seed(M)
    x = (torch.randn(M, D) * torch.rand(M, 1) * 8).bfloat16().cuda()
    w = (1 + 0.1 * torch.randn(D)).bfloat16().cuda()
    monkeypatch.setenv("UMV_NORM_WARP", "0")
    y_block = ops.op_rmsnorm(x, w)
    monkeypatch.setenv("UMV_NORM_WARP", "1")
    y_warp = ops.op_rmsnorm(x, w)
    monkeypatch.setenv("UMV_NORM_WARP", "2")
    y_half = ops.op_rmsnorm(x, w)
    assert torch.equal(y_block, y_warp) and torch.equal(y_block, y_half)
    s = ulp_stats(y_warp, nm.rmsnorm(x.cpu(), w.cpu(), 1e-6))
    assert s["max_ulp"] <= 1 and s["frac"] < 1e-3, s


def test_argmax_ties_go_to_lowest_index(ops):
    torch.manual_seed(0)
    for V in (2048, 152064):
        lg = torch.randn(8, V).bfloat16()
        lg[3, 100] = lg[3, 7] = 50.0
        lg[5, V - 1] = lg[5, V - 2] = 60.0
        assert torch.equal(ops.op_argmax(lg.cuda()).cpu(), torch.argmax(lg, -1))


@pytest.mark.parametrize("dh,H,Hkv,ql,kl,causal", [
    (128, 7, 1, [5, 70], [5, 70], True), (128, 7, 1, [1, 1], [100, 37], True), (128, 7, 1, [64], [64], False),
    (72, 2, 2, [48, 20, 200], [48, 20, 200], False), (128, 28, 4, [130], [300], True), (128, 28, 4, [130], [300], False),
    (72, 16, 16, [1024], [1024], False), (128, 28, 4, [1, 1, 1], [1058, 1185, 64], True), (128, 7, 1, [34, 12], [1060, 50], True),
    (128, 4, 4, [1], [1], True), (72, 2, 2, [1], [1], False)])
def test_attention_varlen(ops, dh, H, Hkv, ql, kl, causal):
    torch.manual_seed(sum(kl))
    q = torch.randn(sum(ql), H, dh).bfloat16()
    k = torch.randn(sum(kl), Hkv, dh).bfloat16()
    v = torch.randn(sum(kl), Hkv, dh).bfloat16()
    o = ops.op_attention(q.cuda(), k.cuda(), v.cuda(), ql, kl, causal)
    s = ulp_stats(o, nm.attention_varlen(q, k, v, ql, kl, causal))
    # both sides round P to bf16 but the online softmax rescales per 64-key block: <= few ulp, tiny rel error
    assert not torch.isnan(o.float()).any() and s["rel_l2"] < 4e-3, s


def test_attention_tcgen05_matches_mma_sync(monkeypatch):
    """The tcgen05 prefill attention (attention_tc.cu) against the mma.sync kernel on the same paged cache: ragged
    lengths, causal and full masks, a second chunk on top of an existing context (kv_len > q_len), tiles that end inside a
    128-key block.  Both follow flash-attn's rounding points; they differ in fp32 accumulation order and in the key
    granularity of the online softmax, i.e. bf16 ulp noise on the layer outputs."""
    import os
    from unimedvl_b200 import config as ucfg
    from unimedvl_b200.engine import Engine
    dims = ucfg.tiny(llm_layers=2)
    eng = Engine(dims, max_tokens=1200, max_seqs=4, kv_pages=64, enable_vit=False, enable_gen=False)
    eng.fill_synthetic(seed=5)
    eng.finalize()
    torch.manual_seed(11)
    lens1, lens2 = [333, 130, 257], [40, 200, 19]
    for causal in (True, False):
        outs = {}
        for tc in ("0", "1"):
            monkeypatch.setenv("UMV_ATTN_TC", tc)
            seqs = [eng.seq_new() for _ in lens1]
            g = torch.Generator(device="cuda").manual_seed(3)
            x1 = (torch.randn(sum(lens1), dims.llm.hidden, device="cuda", generator=g) * 0.5).bfloat16()
            x2 = (torch.randn(sum(lens2), dims.llm.hidden, device="cuda", generator=g) * 0.5).bfloat16()
            pos1 = [p for n in lens1 for p in range(n)]
            pos2 = [lens1[i] + p for i, n in enumerate(lens2) for p in range(n)]
            h1 = eng.llm_forward(x1, seqs, lens1, pos1, is_causal=causal, update_kv=True, want_hidden=True)
            h2 = eng.llm_forward(x2, seqs, lens2, pos2, is_causal=causal, update_kv=True, want_hidden=True)
            outs[tc] = (h1.float().cpu(), h2.float().cpu())
            for s_ in seqs:
                eng.seq_free(s_)
        for a_, b_ in zip(outs["0"], outs["1"]):
            assert torch.isfinite(b_).all()
            rel = (a_ - b_).norm() / a_.norm()
            assert rel < 5e-3, (causal, float(rel))


@pytest.mark.parametrize("n_seqs,ctx", [(1, 3400), (3, 700), (20, 130)])
def test_decode_attention_fused_matches_unfused(monkeypatch, n_seqs, ctx):
    """Fused decode attention (split-K reduce + q/k norm + RoPE + KV append + TMA-staged split-KV attention + DSMEM combine)
    against the three-kernel path on the same cache: a context long enough that a CTA's TMA ring wraps (more than 6 blocks of
    64 keys per CTA), ragged lengths, and more (sample, kv head) pairs than fit one cluster wave of 8.  Per-step logits
    (teacher forced, so both paths see the same inputs) agree to bf16 accumulation-order noise: the logits are bf16 values
    of magnitude ~0.6, one ulp is 0.4-0.8 % of that, and the two schedules sum in different orders (measured rel-L2 0.010
    at every context length from 380 to 3400, identical argmax)."""
    from unimedvl_b200 import config as ucfg
    from unimedvl_b200.engine import Engine
    dims = ucfg.tiny(llm_layers=2)
    eng = Engine(dims, max_tokens=max(4096, n_seqs * ctx), max_seqs=max(n_seqs, 2), kv_pages=n_seqs * (ctx // 64 + 3) + 8,
                 enable_vit=False, enable_gen=False)
    eng.fill_synthetic(seed=7)
    eng.finalize()
    g = torch.Generator(device="cuda").manual_seed(5)
    lens = [ctx - 5 * i for i in range(n_seqs)]
    seqs = [eng.seq_new() for _ in lens]
    x = (torch.randn(sum(lens), dims.llm.hidden, device="cuda", generator=g) * 0.5).bfloat16()
    eng.llm_forward(x, seqs, lens, [p for n in lens for p in range(n)], is_causal=True, update_kv=True, want_hidden=False)
    steps = 5
    forced = torch.randint(0, dims.llm.vocab, (steps, n_seqs), generator=torch.Generator().manual_seed(1))
    outs = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("UMV_FUSED_ATTN", fused)
        forks = [eng.seq_fork(s_) for s_ in seqs]
        _, logits = eng.generate_text(forks, [1] * n_seqs, lens, steps, forced_tokens=forced, return_logits=True)
        outs[fused] = logits.float().cpu()
        for f in forks:
            eng.seq_free(f)
    assert torch.isfinite(outs["1"]).all()
    rel = (outs["1"] - outs["0"]).norm() / outs["0"].norm()
    assert rel < 2e-2, float(rel)
    assert (outs["1"].argmax(-1) == outs["0"].argmax(-1)).float().mean() > 0.9


def test_vit_padded_head_tcgen05_matches_mma_sync(monkeypatch):
    """ViT tower with its q|k|v written into 128-column zero-padded heads and attended by the tcgen05 kernel (softmax scale
    and output width of the real head_dim 72) against the head_dim-72 mma.sync path: ragged images, one of them shorter than
    a 128-row tile boundary multiple."""
    from unimedvl_b200 import config as ucfg
    from unimedvl_b200.engine import Engine
    dims = ucfg.tiny()
    eng = Engine(dims, max_tokens=1024, max_seqs=2, kv_pages=8, enable_gen=False)
    eng.fill_synthetic(seed=9)
    eng.finalize()
    lens = [400, 130, 256]
    g = torch.Generator().manual_seed(2)
    pixels = torch.randn(sum(lens), dims.vit.patch_dim, generator=g)
    pos = torch.cat([torch.randperm(dims.vit.num_positions, generator=g)[:n] for n in lens])
    outs = {}
    for tc in ("0", "1"):
        monkeypatch.setenv("UMV_ATTN_TC", tc)
        outs[tc] = eng.vit_embed(pixels, pos, lens).float().cpu()
    assert torch.isfinite(outs["1"]).all()
    rel = (outs["1"] - outs["0"]).norm() / outs["0"].norm()
    assert rel < 5e-3, float(rel)


# ------------------------------------------------------------------------------------------------ sampling (bagel.py:1297-1301)
def _ref_probs(logits, T):
    """softmax(pred_logits / temperature) as the reference computes it on CUDA: the division of the bf16 logits by the Python
    scalar is x * (1.0f / T) in fp32 rounded to bf16 (ATen's div-by-CPU-scalar), the softmax is an fp32-policy autocast op."""
    scaled = (logits.float() * (torch.tensor(1.0, dtype=torch.float32) / torch.tensor(T, dtype=torch.float32))).bfloat16()
    return torch.softmax(scaled.float(), dim=-1)


@pytest.mark.parametrize("T", [0.1, 1.0])
def test_sampling_distribution_chi2(ops, T):
    """>= 10k draws from one row of logits (every CTA row hashes (seed, row) to its own u): chi-square of the counts against the fp32
    softmax of the reference -- T = 0.1 is the shipped default of the VQA script (interactive_vqa_inferencer.py:66)."""
    V, N = 64, 40000
    torch.manual_seed(5)
    row = (torch.randn(V) * (0.25 if T < 1 else 1.5)).bfloat16()
    logits = row[None].repeat(N, 1).cuda()
    draws = ops.op_sample(logits, T, seed=1234).cpu()
    assert draws.min() >= 0 and draws.max() < V
    p = _ref_probs(row[None], T)[0].double()
    counts = torch.bincount(draws, minlength=V).double()
    keep = p * N >= 5                                    # pool the rare cells (standard chi-square validity rule)
    exp = torch.cat([p[keep] * N, (p[~keep].sum() * N)[None]])
    obs = torch.cat([counts[keep], counts[~keep].sum()[None]])
    ok = exp > 0
    chi2 = float((((obs - exp) ** 2)[ok] / exp[ok]).sum())
    dof = int(ok.sum()) - 1
    # 99.99 % quantile of chi-square(dof) by Wilson-Hilferty; a wrong CDF (off-by-one span, missing temperature) gives thousands
    z = 3.72
    bound = dof * (1 - 2 / (9 * dof) + z * (2 / (9 * dof)) ** 0.5) ** 3
    assert chi2 < bound, (T, chi2, bound, dof)
    # and the most likely token is drawn most often
    assert int(counts.argmax()) == int(p.argmax())


def test_sampling_seed_determinism_and_step_variation(ops):
    V = 152064
    torch.manual_seed(6)
    logits = torch.randn(8, V).bfloat16().cuda()
    a = ops.op_sample(logits, 1.0, seed=77)
    b = ops.op_sample(logits, 1.0, seed=77)
    c = ops.op_sample(logits, 1.0, seed=78)
    assert torch.equal(a, b)                              # same (seed, step, row) -> same draw
    assert not torch.equal(a, c)
    assert len(set(a.tolist())) > 1                       # rows draw independently from (nearly) the same flat distribution


def test_sampling_full_vocab_edges(ops):
    """V = 152064 with 1024 threads: span 149 -> the last three threads own nothing.  u * total rounding up to total (u_force = 1)
    must land on the last token with mass, never leave out[] unwritten; u = 0 lands on the first token with mass."""
    V = 152064
    logits = torch.full((4, V), -30.0).bfloat16()
    logits[0, 5] = logits[0, V - 1] = 10.0
    logits[1, 100] = 10.0                                  # all mass far from the end: the tail spans hold ~0 probability
    logits[2, V - 1] = 10.0
    logits[3, 0] = logits[3, V - 2] = 10.0
    lg = logits.cuda()
    sentinel = torch.full((4,), -7, dtype=torch.int64, device="cuda")
    from unimedvl_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    for u, name in ((1.0, "u=1"), (0.99999994, "u=1-2^-24"), (0.0, "u=0")):
        out = sentinel.clone()
        _lib.check(lib.umv_op_sample(C.c_void_p(lg.data_ptr()), 4, V, C.c_float(1.0), C.c_uint64(0), C.c_float(u),
                                     C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        got = out.cpu().tolist()
        assert all(0 <= t < V for t in got), (name, got)               # every row written
        if u == 0.0:
            assert got[0] in (0, 5) and got[3] == 0 and got[2] in range(0, V), (name, got)
        else:
            assert got[0] == V - 1 and got[2] == V - 1 and got[3] in (V - 2, V - 1), (name, got)
            assert got[1] >= 100, (name, got)
    # temperature -> 0 limit reproduces argmax
    torch.manual_seed(8)
    rnd = torch.randn(8, V).bfloat16()
    rnd[torch.arange(8), torch.arange(8) * 1000 + 17] = 20.0          # a unique maximum per row (bf16 ties would be a coin flip)
    rnd = rnd.cuda()
    assert torch.equal(ops.op_sample(rnd, 1e-3, seed=3), ops.op_argmax(rnd))


def test_generate_text_sampling_branch(ops):
    """do_sample=True through Bagel.generate_text: same seed -> same tokens, the draws follow the step counter (different tokens per
    step)."""
    from unimedvl_b200 import config as ucfg
    from unimedvl_b200.bagel import Bagel
    from unimedvl_b200.cache import NaiveCache
    from unimedvl_b200.engine import Engine
    from unimedvl_b200 import packing
    dims = ucfg.tiny()
    eng = Engine(dims, max_tokens=256, max_seqs=4, kv_pages=32, enable_vit=False, enable_gen=False)
    eng.fill_synthetic(seed=3)
    eng.finalize()
    model = Bagel(eng, dims)
    tok = dict(bos_token_id=2040, eos_token_id=2041, start_of_image=2042, end_of_image=2043)

    class _Ids:
        def encode(self, i): return [(7 * i + j * 13) % 2000 for j in range(9 + i)]

    def run(**kw):
        g, lens, rope = packing.prepare_prompts([0, 0], [0, 0], [0, 1], _Ids(), tok)
        cache = model.forward_cache_update_text(NaiveCache(dims.llm.layers), **g)
        st = packing.prepare_start_tokens(lens, rope, tok)
        return model.generate_text(past_key_values=cache, max_length=12, end_token_id=None, **st, **kw).cpu()
    greedy = run()
    a = run(do_sample=True, temperature=1.0, seed=5)
    b = run(do_sample=True, temperature=1.0, seed=5)
    c = run(do_sample=True, temperature=1.0, seed=6)
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a, greedy)
    assert a.shape == (12, 2) and int(a.max()) < dims.llm.vocab
