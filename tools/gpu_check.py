"""First-contact GPU diagnostics (run on the B200 box under gpurun).  Every group runs in its own
subprocess with a timeout so that one crashing / hanging kernel cannot hide the others; results go to
gpurun_out/gpu_check_<group>.log.  This is a development aid, the graded parity tests live in tests/.

    python tools/gpu_check.py [group ...]        # groups: linear norms attention model model_simple bench
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")


def stats(a, b):
    from util import ulp_stats
    import torch
    s = ulp_stats(a.cpu(), b.cpu())
    s["nan"] = bool(torch.isnan(a.float()).any().item())
    return {k: (round(v, 6) if isinstance(v, float) else v) for k, v in s.items()}


def group_linear():
    import torch
    from unimedvl_b200.engine import op_linear
    torch.manual_seed(0)
    dev = "cuda"
    shapes = [(8, 512, 256), (8, 4608, 3584), (16, 3584, 3584), (33, 1152, 896), (128, 256, 128), (300, 3456, 1152),
              (1000, 1152, 4304), (200, 896, 1536), (70, 64, 896), (130, 896, 64), (8, 37888, 3584), (8, 3584, 18944),
              (2048, 3584, 3584), (5, 3072, 896), (64, 2048, 896), (1026, 37888, 3584)]
    worst = 0.0
    for (M, N, K) in shapes:
        x = (torch.randn(M, K, device=dev)).bfloat16()
        w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        b = (torch.randn(N, device=dev) * 0.1).bfloat16()
        acc = x.float() @ w.float().T
        for epi in (0, 1, 3, 2):
            if epi == 2 and N % 128:
                continue
            if epi == 2:
                g = acc.view(M, N // 128, 2, 64)[:, :, 0].reshape(M, N // 2).bfloat16()
                u = acc.view(M, N // 128, 2, 64)[:, :, 1].reshape(M, N // 2).bfloat16()
                ref = torch.nn.functional.silu(g) * u
                res = None
                bias = None
            else:
                bias = b
                ref = (acc + b.float()).bfloat16()
                res = None
                if epi == 1:
                    ref = torch.nn.functional.gelu(ref, approximate="tanh")
                if epi == 3:
                    res = torch.randn(M, N, device=dev).bfloat16()
                    ref = ref + res
            for impl in (3, 1, 2):
                if impl == 2 and M > 64:
                    continue
                if impl == 3 and M * N * K > 2e10:
                    continue
                try:
                    t0 = time.time()
                    y = op_linear(x, w, bias, res, epi=epi, impl=impl)
                    torch.cuda.synchronize()
                    s = stats(y, ref)
                    bad = s["nan"] or s["rel_l2"] > 5e-3
                    worst = max(worst, s["rel_l2"])
                    print(("FAIL " if bad else "ok   ") + f"linear M={M} N={N} K={K} epi={epi} impl={impl} "
                          f"{time.time() - t0:.3f}s {s}", flush=True)
                except Exception as ex:  # noqa: BLE001
                    print(f"EXC  linear M={M} N={N} K={K} epi={epi} impl={impl}: {ex}", flush=True)
    print("worst rel_l2", worst)


def group_norms():
    import torch
    from unimedvl_b200.engine import op_rmsnorm, op_layernorm, op_argmax
    from oracle import numerics as nm
    torch.manual_seed(1)
    for (M, D) in [(8, 3584), (100, 896), (3, 128), (17, 1152), (50, 144)]:
        x = torch.randn(M, D).bfloat16()
        w = (1 + 0.1 * torch.randn(D)).bfloat16()
        b = (0.1 * torch.randn(D)).bfloat16()
        y = op_rmsnorm(x.cuda(), w.cuda())
        print("rmsnorm", M, D, stats(y, nm.rmsnorm(x, w, 1e-6)), flush=True)
        y = op_layernorm(x.cuda(), w.cuda(), b.cuda())
        print("layernorm", M, D, stats(y, nm.layernorm(x, w, b, 1e-6, nm.Semantics.cuda).bfloat16()), flush=True)
    for V in (2048, 152064):
        lg = torch.randn(8, V).bfloat16()
        lg[3, 100] = lg[3, 7] = 50.0      # tie -> lowest index
        out = op_argmax(lg.cuda()).cpu()
        print("argmax", V, bool((out == torch.argmax(lg, -1)).all()), out.tolist(), flush=True)


def group_attention():
    import torch
    from unimedvl_b200.engine import op_attention
    from oracle import numerics as nm
    torch.manual_seed(2)
    cases = [
        (128, 7, 1, [5, 70], [5, 70], True), (128, 7, 1, [1, 1], [100, 37], True), (128, 7, 1, [64], [64], False),
        (72, 2, 2, [48, 20, 200], [48, 20, 200], False), (128, 28, 4, [130], [300], True), (128, 28, 4, [130], [300], False),
        (72, 16, 16, [1024], [1024], False), (128, 28, 4, [1, 1, 1], [1058, 1185, 64], True),
        (128, 7, 1, [34, 12], [1060, 50], True),
    ]
    for (dh, H, Hkv, ql, kl, causal) in cases:
        q = torch.randn(sum(ql), H, dh).bfloat16()
        k = torch.randn(sum(kl), Hkv, dh).bfloat16()
        v = torch.randn(sum(kl), Hkv, dh).bfloat16()
        try:
            o = op_attention(q.cuda(), k.cuda(), v.cuda(), ql, kl, causal)
            torch.cuda.synchronize()
            ref = nm.attention_varlen(q, k, v, ql, kl, causal)
            s = stats(o, ref)
            bad = s["nan"] or s["rel_l2"] > 1e-2
            print(("FAIL " if bad else "ok   ") + f"attention dh={dh} H={H}/{Hkv} q={ql} k={kl} causal={causal} {s}", flush=True)
        except Exception as ex:  # noqa: BLE001
            print(f"EXC  attention dh={dh} H={H}/{Hkv} q={ql} k={kl}: {ex}", flush=True)


def group_model():
    import torch
    from util import Golden, make_oracle, tiny_weights, Semantics
    from unimedvl_b200.engine import Engine
    d, sd, _ = tiny_weights()
    eng = Engine(d, max_tokens=512, max_seqs=4, kv_pages=64)
    eng.load_state_dict(sd)
    eng.finalize()
    print("weights", eng.weight_bytes(), flush=True)
    o = make_oracle(Semantics.cuda)
    g = Golden("vqa")
    gi = g.group("vqa.vit_in")
    # --- ViT + connector
    ref = __import__("oracle.vit", fromlist=["x"]).vit_tokens_to_llm(o.sd, o.dims.vit, gi["packed_vit_tokens"],
                                                                      gi["packed_vit_position_ids"], gi["vit_token_seqlens"])
    emb = eng.vit_embed(gi["packed_vit_tokens"], gi["packed_vit_position_ids"], gi["vit_token_seqlens"].tolist())
    torch.cuda.synchronize()
    print("vit_embed", stats(emb, ref), flush=True)
    # --- image prefill (non-causal), using the oracle's embeddings so stages are isolated
    B = 2
    seqs = [eng.seq_new() for _ in range(B)]
    lens = gi["packed_seqlens"].tolist()
    ids = eng.embed_tokens(gi["packed_text_ids"])
    seq = torch.zeros((sum(lens), d.llm.hidden), dtype=torch.bfloat16, device="cuda")
    seq[gi["packed_text_indexes"].cuda()] = ids
    seq[gi["packed_vit_token_indexes"].cuda()] = ref.cuda()
    eng.llm_forward(seq, seqs, lens, gi["packed_position_ids"].tolist(), is_causal=False, update_kv=True, want_hidden=False)
    cache = o.forward_cache_update_vit(o.new_cache(), **gi)
    off = [0, lens[0], lens[0] + lens[1]]
    for li in range(d.llm.layers):
        for b in range(B):
            k, v = eng.seq_export(seqs[b], li)
            print(f"prefill_vit L{li} b{b} K", stats(k, cache.key[li][off[b]:off[b + 1]]), "V",
                  stats(v, cache.value[li][off[b]:off[b + 1]]), flush=True)
    # --- text prefill (causal)
    gt = g.group("vqa.text_in")
    tl = gt["text_token_lens"].tolist()
    x = eng.embed_tokens(gt["packed_text_ids"])
    eng.llm_forward(x, seqs, tl, gt["packed_text_position_ids"].tolist(), is_causal=True, update_kv=True, want_hidden=False)
    cache = o.forward_cache_update_text(cache, **gt)
    kvl = [lens[b] + tl[b] for b in range(B)]
    off = [0, kvl[0], kvl[0] + kvl[1]]
    for li in range(d.llm.layers):
        for b in range(B):
            k, v = eng.seq_export(seqs[b], li)
            print(f"prefill_text L{li} b{b} K", stats(k, cache.key[li][off[b]:off[b + 1]]), flush=True)
    # --- decode: teacher-forced logits, then free-running tokens
    st = g.group("vqa.start")
    forced = g.t("vqa.tokens")
    fork = [eng.seq_fork(s) for s in seqs]
    toks, logits = eng.generate_text(fork, st["packed_start_tokens"].tolist(), st["packed_query_position_ids"].tolist(), 9,
                                     forced_tokens=forced, return_logits=True)
    lg = []
    o.generate_text(cache.clone(), st["packed_key_value_indexes"], st["key_values_lens"], st["packed_start_tokens"],
                    st["packed_query_position_ids"], 9, forced_tokens=forced, logits_out=lg)
    for s_ in range(9):
        a, b_ = logits[s_].cpu(), lg[s_]
        print(f"decode step {s_}", stats(a, b_), "argmax_eq", (a.float().argmax(-1) == b_.float().argmax(-1)).tolist(), flush=True)
    fork2 = [eng.seq_fork(s) for s in seqs]
    toks2 = eng.generate_text(fork2, st["packed_start_tokens"].tolist(), st["packed_query_position_ids"].tolist(), 9)
    ref_t = o.generate_text(cache.clone(), st["packed_key_value_indexes"], st["key_values_lens"], st["packed_start_tokens"],
                            st["packed_query_position_ids"], 9)
    print("free-running tokens engine", toks2.cpu().T.tolist(), "oracle", ref_t.T.tolist(), "golden", g.t("vqa.tokens").T.tolist(), flush=True)
    print("pages free", eng.pages_free(), "launches", eng.launch_count(), flush=True)


def group_bench():
    import torch
    from unimedvl_b200 import config as ucfg
    from unimedvl_b200.engine import Engine
    d = ucfg.bagel_7b_mot()
    t0 = time.time()
    eng = Engine(d, max_tokens=1100, max_seqs=8, kv_pages=400, enable_gen=False)
    eng.fill_synthetic(0)
    eng.finalize()
    torch.cuda.synchronize()
    print(f"engine up in {time.time() - t0:.1f}s, weights {eng.weight_bytes() / 1e9:.2f} GB", flush=True)
    B = 8
    seqs = [eng.seq_new() for _ in range(B)]
    # prefill 1058 tokens per sample, one sample per call
    for b in range(B):
        x = (torch.randn(1058, d.llm.hidden, device="cuda") * 0.05).bfloat16()
        ev0, ev1 = torch.cuda.Event(True), torch.cuda.Event(True)
        ev0.record()
        eng.llm_forward(x, [seqs[b]], [1058], [0] * 1026 + list(range(1, 33)), is_causal=True, update_kv=True, want_hidden=False)
        ev1.record()
        torch.cuda.synchronize()
        print(f"prefill sample {b}: {ev0.elapsed_time(ev1):.2f} ms", flush=True)
    for n_steps in (4, 32, 128):
        fork = [eng.seq_fork(s) for s in seqs]
        ev0, ev1 = torch.cuda.Event(True), torch.cuda.Event(True)
        ev0.record()
        toks = eng.generate_text(fork, [151644] * B, [33] * B, n_steps)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        print(f"decode B={B} steps={n_steps}: {ms:.2f} ms total, {ms / n_steps:.3f} ms/step, {B * n_steps / ms * 1e3:.1f} tok/s "
              f"tokens[:,0]={toks[:6, 0].tolist()}", flush=True)
        for s in fork:
            eng.seq_free(s)


GROUPS = {"linear": group_linear, "norms": group_norms, "attention": group_attention, "model": group_model,
          "bench": group_bench}
ENVS = {"model_simple": ("model", {"UMV_GEMM_IMPL": "3"}), "model_nosplit": ("model", {"UMV_SPLITK": "0"}),
        "model_nograph": ("model", {"UMV_GRAPH": "0"}), "bench_nosplit": ("bench", {"UMV_SPLITK": "0"})}


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) >= 3 and sys.argv[1] == "--run":
        GROUPS[sys.argv[2]]()
        return
    names = sys.argv[1:] or ["norms", "linear", "attention", "model_simple", "model", "model_nosplit", "model_nograph", "bench"]
    summary = {}
    for name in names:
        grp, env = ENVS.get(name, (name, {}))
        log = os.path.join(OUT, f"gpu_check_{name}.log")
        t0 = time.time()
        with open(log, "w") as f:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", grp], stdout=f, stderr=subprocess.STDOUT,
                                   timeout=420, env={**os.environ, **env})
                rc = r.returncode
            except subprocess.TimeoutExpired:
                rc = "timeout"
        txt = open(log).read()
        summary[name] = dict(rc=rc, secs=round(time.time() - t0, 1), fails=txt.count("FAIL "), excs=txt.count("EXC "))
        print(name, summary[name], flush=True)
        print("\n".join(txt.splitlines()[-12:]), flush=True)
    json.dump(summary, open(os.path.join(OUT, "gpu_check_summary.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
