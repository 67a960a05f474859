#!/usr/bin/env python
"""Headline benchmark: VQA decode tokens/s at the 14B dims (BASELINE.json configs[1]:
"VQA batch=8, 448x448, 14B bf16, 128-token greedy decode, 1xB200").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference]

One step = one pass of the hot path over one batch per GPU:
  value : the 128-step greedy decode of 8 samples (ctx 1058 -> 1185) from a KV cache already
          resident in HBM -- Bagel.generate_text through the engine (device loop, CUDA-graph replay);
  e2e   : the whole VQA job through the reference-facing call (Bagel.vqa_generate_images): pinned HOST
          uint8 images / prompt ids -> H2D -> normalise + patchify on the device -> ViT + connector ->
          image prefill -> prompt prefill -> 128-step decode -> D2H of the tokens; tokens/s counts the same
          8 x 128 decoded tokens.
N > 1: one process per GPU (torchrun), requests sharded data-parallel (weak scaling: 8 samples per
GPU), one NCCL all_gather of the output tokens per step; time = max over ranks.
--impl reference: the reference's CPU forward, restated by oracle/ (the reference itself cannot
travel to the GPU box), timed on the host cores on a bounded sample of the same decode step.
Weights are random-init (no checkpoint offline), inputs synthetic.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU = 8
# DRAM traffic of one weight-major gate/up + SwiGLU launch at M=8 (ncu --set full, profiles/r1_decode_kernels_full.md):
# 271,748,096 B read + 3,301,632 B written; the algorithmic figure is 271,941,632 B.
NCU_TRAFFIC_GATE_UP = 275_049_728
DECODE_STEPS = 128
PROMPT_TOKENS = 30
IMG = 448


WORKLOAD = ("VQA batch=8 per GPU, 448x448 (1024 ViT tokens + 2 markers), 32-token prompt, 14B MoT bf16 "
            "(both experts resident), 128-token greedy decode, ctx 1058->1185")      # BASELINE.json configs[1], both arms


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(r[0])) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def synthetic_job(rank: int):
    """Pinned host inputs of one batch: 448x448 uint8 noise images through the reference's transform +
    patchify (host), 30 random prompt ids per sample."""
    import torch
    from PIL import Image
    from unimedvl_b200 import packing, synth
    tf = packing.ImageTransform(980, 378, 14, max_pixels=2_007_040)
    toks, pos, lens, prompts, images = [], [], [], [], []
    for i in range(B_PER_GPU):
        gid = rank * B_PER_GPU + i
        raw = synth.synthetic_image(gid, IMG, IMG)
        images.append(torch.from_numpy(raw.copy()).pin_memory())     # 448x448 passes the reference's resize rule unchanged
        t = tf(Image.fromarray(raw))
        toks.append(packing.patchify(t, 14))
        pos.append(packing.flattened_position_ids(t.size(1), t.size(2), 14, 70))
        lens.append(toks[-1].shape[0])
        prompts.append(synth.synthetic_prompt_ids(gid, PROMPT_TOKENS))
    pixels = torch.cat(toks, 0).pin_memory()
    pos_ids = torch.cat(pos, 0).pin_memory()
    return pixels, pos_ids, lens, prompts, images


# ------------------------------------------------------------------------------------------------
def cpu_decode_baseline(n_layers_sample: int = 2, steps: int = 3, ctx: int = 1058, threads: int | None = None):
    """Reference CPU forward of ONE decode step at the 14B dims (oracle/, native bf16 CPU GEMMs, the
    reference's per-step KV re-materialisation included), on a bounded sample: `n_layers_sample` of the 28
    decoder layers (time scaled by 28/n) + the full lm_head.  Returns tok/s and a description."""
    import torch
    from oracle import llm as ollm
    from unimedvl_b200 import config as ucfg
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    L = ucfg.bagel_7b_mot().llm
    dims = ollm.LLMDims(L.hidden, L.heads, L.kv_heads, L.inter, n_layers_sample, L.vocab, L.rope_theta, L.eps)
    g = torch.Generator().manual_seed(0)
    rnd = lambda *s: (torch.randn(*s, generator=g, dtype=torch.float32) * 0.02).to(torch.bfloat16)
    sd = {"language_model.model.embed_tokens.weight": rnd(4096, L.hidden),      # 4096 rows suffice for the lookup
          "language_model.lm_head.weight": rnd(L.vocab, L.hidden),
          "language_model.model.norm.weight": torch.ones(L.hidden, dtype=torch.bfloat16)}
    dh = L.head_dim
    for i in range(n_layers_sample):
        P = f"language_model.model.layers.{i}."
        sd[P + "self_attn.q_proj.weight"] = rnd(L.heads * dh, L.hidden); sd[P + "self_attn.q_proj.bias"] = rnd(L.heads * dh)
        sd[P + "self_attn.k_proj.weight"] = rnd(L.kv_heads * dh, L.hidden); sd[P + "self_attn.k_proj.bias"] = rnd(L.kv_heads * dh)
        sd[P + "self_attn.v_proj.weight"] = rnd(L.kv_heads * dh, L.hidden); sd[P + "self_attn.v_proj.bias"] = rnd(L.kv_heads * dh)
        sd[P + "self_attn.o_proj.weight"] = rnd(L.hidden, L.heads * dh)
        sd[P + "self_attn.q_norm.weight"] = torch.ones(dh, dtype=torch.bfloat16)
        sd[P + "self_attn.k_norm.weight"] = torch.ones(dh, dtype=torch.bfloat16)
        sd[P + "mlp.gate_proj.weight"] = rnd(L.inter, L.hidden); sd[P + "mlp.up_proj.weight"] = rnd(L.inter, L.hidden)
        sd[P + "mlp.down_proj.weight"] = rnd(L.hidden, L.inter)
        sd[P + "input_layernorm.weight"] = torch.ones(L.hidden, dtype=torch.bfloat16)
        sd[P + "post_attention_layernorm.weight"] = torch.ones(L.hidden, dtype=torch.bfloat16)
    B = B_PER_GPU
    cache = ollm.PackedKV(n_layers_sample)
    for i in range(n_layers_sample):
        cache.key[i] = rnd(B * ctx, L.kv_heads, dh)
        cache.value[i] = rnd(B * ctx, L.kv_heads, dh)
    kv_lens = torch.full((B,), ctx, dtype=torch.int64)
    kvidx = torch.arange(B * ctx)
    pos = torch.full((B,), 33, dtype=torch.int64)
    toks = torch.randint(0, 4096, (B,), generator=g)
    t_layers = t_head = 0.0
    with torch.no_grad():
        for s in range(steps + 1):
            qidx = torch.cumsum(kv_lens, 0) + torch.arange(B)
            parts = list(kvidx.split(kv_lens.tolist()))
            kvidx = torch.cat([p + i for i, p in enumerate(parts)])
            t0 = time.perf_counter()
            h, cache = ollm.forward_inference(sd, dims, ollm.embed(sd, toks), torch.ones(B, dtype=torch.int64), pos, qidx, cache,
                                              kv_lens, kvidx, True, True, exact=False)
            t1 = time.perf_counter()
            logits = ollm.lm_head(sd, h, exact=False)
            nxt = torch.argmax(logits, -1) % 4096
            t2 = time.perf_counter()
            if s > 0:                      # step 0 is the warm-up
                t_layers += t1 - t0
                t_head += t2 - t1
            parts = list(kvidx.split(kv_lens.tolist()))
            kvidx = torch.cat([torch.cat([p, p[-1:] + 1]) for p in parts])
            kv_lens, pos, toks = kv_lens + 1, pos + 1, nxt
    step_s = (t_layers / steps) * (L.layers / n_layers_sample) + t_head / steps
    return {"value": round(B / step_s, 3), "unit": "tok/s", "cores": cores, "kind": "port",
            "sample": f"oracle/ decode forward, B={B}, ctx={ctx}, {n_layers_sample} of {L.layers} layers x{steps} steps "
                      f"(layer time scaled x{L.layers // n_layers_sample}) + full lm_head; native bf16 CPU GEMM",
            "ms_per_step_extrapolated": round(step_s * 1e3, 1)}


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    t0 = time.perf_counter()
    base = None
    for i in range(args.warmup + args.steps):
        r = cpu_decode_baseline(steps=2)
        if i >= args.warmup:
            base = r if base is None or r["value"] > base["value"] else base
    line = {"impl": "reference", "metric": "VQA decode tok/s @14B (B=8/GPU, 448x448, 128-token greedy)", "value": base["value"],
            "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": base["ms_per_step_extrapolated"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": {"workload": WORKLOAD,
                                                             "arm": "CPU forward of the reference algorithm (oracle port), rank 0 only"},
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": round(time.perf_counter() - t0, 1)}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def t2i_secondary(eng, model, dims, tok, n_images: int = 4, H: int = 256, W: int = 256):
    """Secondary metric (BASELINE.json configs[3] per-GPU share): text -> 256x256 image, 4 images per GPU, 50 timesteps,
    dual CFG 4.0 / 1.5 on (0.4, 1], per-image "global" renorm (defaults of inferencer.py:165-178): 131 LLM forwards per
    image batched as 3 CFG branches per step.  Device-timed generate_image (the VAE decode is not resident in this engine
    instance: it adds 0.62 of ~446 TFLOP per image)."""
    import torch
    from unimedvl_b200 import packing, synth
    from unimedvl_b200.cache import NaiveCache
    B = n_images

    class _Ids:
        def encode(self, i): return synth.synthetic_prompt_ids(100 + i, 30)
    g, lens, rope = packing.prepare_prompts([0] * B, [0] * B, list(range(B)), _Ids(), tok)
    ctx = model.forward_cache_update_text(NaiveCache(dims.llm.layers), **g)
    cfg_text = NaiveCache(dims.llm.layers)                      # context without the prompt: empty
    from unimedvl_b200.cache import paged_handle
    paged_handle(cfg_text, eng, B)
    cfg_img = NaiveCache(dims.llm.layers)
    cfg_img = model.forward_cache_update_text(cfg_img, **g)     # text-only context (no image in a pure T2I request)
    torch.manual_seed(42)
    gi = model.prepare_vae_latent(lens, rope, [(H, W)] * B, tok)
    ct = model.prepare_vae_latent_cfg([0] * B, [0] * B, [(H, W)] * B)
    ci = model.prepare_vae_latent_cfg(lens, rope, [(H, W)] * B)

    def run(steps):
        return model.generate_image(
            past_key_values=ctx, cfg_text_past_key_values=cfg_text, cfg_img_past_key_values=cfg_img, num_timesteps=steps,
            timestep_shift=3.0, cfg_text_scale=4.0, cfg_img_scale=1.5, cfg_interval=(0.4, 1.0), cfg_renorm_min=0.0,
            cfg_renorm_type="global", **gi,
            cfg_text_packed_position_ids=ct["cfg_packed_position_ids"], cfg_text_packed_query_indexes=ct["cfg_packed_query_indexes"],
            cfg_text_key_values_lens=ct["cfg_key_values_lens"], cfg_text_packed_key_value_indexes=ct["cfg_packed_key_value_indexes"],
            cfg_img_packed_position_ids=ci["cfg_packed_position_ids"], cfg_img_packed_query_indexes=ci["cfg_packed_query_indexes"],
            cfg_img_key_values_lens=ci["cfg_key_values_lens"], cfg_img_packed_key_value_indexes=ci["cfg_packed_key_value_indexes"])
    run(4)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    lat = run(50)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    flops = 0.0
    for t_on, branches in ((41, 3), (8, 1)):
        flops += t_on * branches * B * (2 * 258 * 6.5253e9 + 4 * 258 * (32 + 258) * 3584 * 28)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0
    return {"metric": "T2I img/s @14B, 256x256, 50 steps, dual CFG", "value": round(B / (ms / 1e3), 4), "unit": "img/s",
            "images_per_gpu": B, "ms_per_batch": round(ms, 1), "algorithmic_tflop_per_image": round(flops / B / 1e12, 1),
            "roofline": {"bound": "tensor", "achieved": round(flops / (ms / 1e3) / 1e12, 1), "peak": peak, "unit": "TFLOP/s",
                         "frac": round(flops / (ms / 1e3) / 1e12 / peak, 4), "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained"},
            "finite": bool(all(torch.isfinite(x).all().item() for x in lat))}


def run_engine(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist
    from copy import deepcopy
    from unimedvl_b200 import config as ucfg, packing, dp
    from unimedvl_b200.bagel import Bagel
    from unimedvl_b200.cache import NaiveCache
    from unimedvl_b200.engine import Engine
    from unimedvl_b200 import _lib
    import ctypes as C

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dims = ucfg.bagel_7b_mot()
    B = B_PER_GPU
    ntok_img = (IMG // 14) ** 2 + 2
    ctx0 = ntok_img + PROMPT_TOKENS + 2
    eng = Engine(dims, max_tokens=B * ntok_img, max_seqs=B, kv_pages=B * 64, enable_vit=True, enable_gen=True)
    eng.fill_synthetic(seed=0)          # random-init weights of the reference architecture, both experts resident
    eng.finalize()
    model = Bagel(eng, dims)
    tok = dict(ucfg.QWEN25_TOKEN_IDS)
    pixels, pos_ids, lens, prompts, images = synthetic_job(rank)

    # ---- resident context for the device-timed decode: the same prefill the e2e path performs
    cache = NaiveCache(dims.llm.layers)
    zeros = [0] * B
    L = packing._image_block_layout(zeros, zeros, lens, tok)
    g = dict(packed_text_ids=torch.as_tensor(L["text_ids"]), packed_text_indexes=torch.as_tensor(L["text_idx"]),
             packed_vit_tokens=pixels.cuda(), packed_vit_token_indexes=torch.as_tensor(L["img_idx"]),
             packed_vit_position_ids=pos_ids.cuda(), vit_token_seqlens=lens, packed_position_ids=torch.as_tensor(L["pos"]),
             packed_seqlens=L["seqlens"], packed_indexes=torch.as_tensor(L["packed_idx"]),
             packed_key_value_indexes=torch.as_tensor(L["kv_indexes"]), key_values_lens=zeros)
    cache = model.forward_cache_update_vit(cache, **g)

    class _Ids:
        def encode(self, i): return list(prompts[i])
    gp, kvl, rope = packing.prepare_prompts(L["seqlens"], [1] * B, list(range(B)), _Ids(), tok)
    cache = model.forward_cache_update_text(cache, **gp)
    start = packing.prepare_start_tokens(kvl, rope, tok)
    assert kvl == [ctx0] * B

    def decode_step():
        c = deepcopy(cache)                       # page fork (the reference deep-copies the KV, inferencer.py:261)
        t = model.generate_text(past_key_values=c, max_length=DECODE_STEPS, end_token_id=None, **start)
        return dp.gather_tokens(t, world * B)       # one NCCL all_gather of the output tokens

    def e2e_step():
        t = model.vqa_generate_images(images, prompts, tok, DECODE_STEPS)
        return dp.gather_tokens(t.cuda(), world * B) if world > 1 else t

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        decode_step()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = eng.launch_count()
    ms = timed(decode_step, args.steps)
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    tokens_per_step = world * B * DECODE_STEPS
    value = tokens_per_step * args.steps / (ms / 1e3)

    # ---- end to end (host buffers, H2D/D2H inside)
    for _ in range(2):
        e2e_step()
    e2e_steps = max(1, min(args.steps, 5))
    ms_e2e = timed(e2e_step, e2e_steps)
    e2e_value = tokens_per_step * e2e_steps / (ms_e2e / 1e3)
    h2d = sum(im.numel() for im in images) + sum(len(p) + 2 for p in prompts) * 8 + B * 2 * 8
    d2h = B * DECODE_STEPS * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: the weight-major gate/up linear (58% of the bytes of a decode step)
    peak, peak_src = _peaks()
    wb = C.c_int64()
    reps = 8
    for layer in range(dims.llm.layers):
        _lib.check(eng.lib.umv_bench_decode_linear(eng.h, 2, layer, B, C.byref(wb), C.c_void_p(eng.stream.cuda_stream)))
    eng.stream.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(eng.stream):
        ev0.record()
        for _ in range(reps):
            for layer in range(dims.llm.layers):      # 28 distinct 271 MB weight blocks: 7.6 GB >> 126 MB L2
                eng.lib.umv_bench_decode_linear(eng.h, 2, layer, B, C.byref(wb), C.c_void_p(eng.stream.cuda_stream))
        ev1.record()
    eng.stream.synchronize()
    k_ms = ev0.elapsed_time(ev1) / (reps * dims.llm.layers)
    D, I = dims.llm.hidden, dims.llm.inter
    alg_bytes = wb.value + B * D * 2 + B * I * 2
    achieved = alg_bytes / (k_ms / 1e3) / 1e9
    # whole decode step against the same roofline (weights + KV read/write + embedding rows)
    ctx_mean = ctx0 + (DECODE_STEPS - 1) / 2
    w_bytes = 2 * (dims.llm.layers * 233_058_048 + dims.llm.vocab * D + D)
    step_bytes = w_bytes + B * ctx_mean * 57_344 + B * 57_344 + B * D * 2
    step_ms = ms / args.steps / DECODE_STEPS
    step_achieved = step_bytes / (step_ms / 1e3) / 1e9

    t2i = t2i_secondary(eng, model, dims, tok) if (world == 1 and args.t2i) else None
    cpu = cpu_decode_baseline() if world == 1 else None
    line = {
        "metric": "VQA decode tok/s @14B (B=8/GPU, 448x448, 128-token greedy)", "value": round(value, 1), "unit": "tok/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "parallelism": f"dp{world}", "l2": "weights streamed per decode forward (14.1 GB) exceed the 126 MB L2",
                   "weights": "random-init, generated on device", "step": "generate_text(max_length=128) for the batch"},
        "ms_per_decode_forward": round(step_ms, 4),
        "e2e": {"value": round(e2e_value, 1), "unit": "tok/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": round(ms_e2e / e2e_steps, 2), "steps": e2e_steps,
                "path": "Bagel.vqa_generate_images: pinned host uint8 images + prompt ids -> H2D -> normalise/patchify on device -> ViT -> "
                        "image+prompt prefill -> 128-step decode -> host tokens"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "gemm_tc_kernel<16,2,true> (weight-major gate/up + SwiGLU, M=8)",
                     "achieved": round(achieved, 1), "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": NCU_TRAFFIC_GATE_UP, "traffic_source": "profiles/r1_decode_kernels_full.md "
                     "(ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum of this kernel, per launch)",
                     "algorithmic_bytes_per_launch": int(alg_bytes),
                     "us_per_launch": round(k_ms * 1e3, 2)},
        "step_roofline": {"bound": "hbm", "algorithmic_bytes_per_decode_forward": int(step_bytes),
                          "achieved": round(step_achieved, 1), "unit": "GB/s", "frac": round(step_achieved / peak, 4),
                          "roofline_tok_s_per_gpu": round(B / (step_bytes / (peak * 1e9)), 1)},
        "clocks": clocks,
    }
    if t2i is not None:
        line["t2i"] = t2i
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--t2i", type=int, default=1, help="also time the secondary metric (text-to-image img/s) at N=1")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_engine(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
