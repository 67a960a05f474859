#!/usr/bin/env python
"""Benchmark of the unified forward path at the 14B dims, every BASELINE.json config, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference] [--legs a,b,...]

Headline (`value`, BASELINE.json configs[1]): VQA decode tokens/s, 8 samples per GPU, 448x448, 128-token greedy decode from a
KV cache resident in HBM (Bagel.generate_text through the engine: device loop, CUDA-graph replay), device-timed, max over
ranks.  `e2e`: the whole VQA job through the reference-facing call (Bagel.vqa_generate_images) from pinned HOST uint8 images
to host tokens, H2D / D2H inside the timed region.  Extra keys, measured at every N (per-GPU share of the config, weak scaling):
  report_gen   configs[2]  16 samples per GPU, 512-token decode over the paged KV (ctx 1058 -> 1569)
  t2i          configs[3]  4 prompts per GPU -> 256x256 images: prompt prefill (3 contexts), 50-step dual-CFG flow loop, VAE
               decode, uint8 conversion on the device, NCCL all_gather of the images, D2H
  interleaved  configs[4]  8 requests per GPU, each: image + question -> 64 greedy tokens, then image + instruction -> 256x256
               image (VAE + ViT context, CFG 4.0 / 2.0 on [0, 1], text_channel renorm): BatchedInferencer end to end
N == 1 only (rank 0):
  gpu_reference  the UNMODIFIED reference (baseline/_ref/codes; real flash-attn, CUDA autocast) timed on the same B200 for
                 configs[1] and a single-image T2I: the like-for-like GPU baseline of BASELINE.md section 5
  cpu_baseline   configs[0]: the unmodified reference through InterleaveInferencer.__call__ on the host cores (one 448x448 VQA,
                 32-token answer, all 28 layers); falls back to the oracle port on a bounded sample when no reference copy exists
--impl reference: the reference's own CPU forward of the headline metric (28 layers, B=8, ctx 1058: Bagel.generate_text steps on
the host cores), rank 0 only.  Weights are random-init (no checkpoint offline), inputs synthetic.
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

B_PER_GPU = 8            # configs[1]
DECODE_STEPS = 128
PROMPT_TOKENS = 30
IMG = 448
B_REPORT, REPORT_STEPS = 16, 512      # configs[2]: 32 samples over 2 GPUs
B_T2I, T2I_SIZE, T2I_STEPS = 4, 256, 50   # configs[3]: 16 images over 4 GPUs
B_INTER, INTER_TOKENS = 8, 64         # configs[4]: 64 requests over 8 GPUs
# DRAM traffic of one weight-major gate/up + SwiGLU launch at M=8 (ncu --set full, profiles/r2_decode_kernels_full.md, second layer):
# 271,748,864 B read + 3,923,712 B written (round 1: 271,748,096 + 3,301,632); the algorithmic figure is 271,941,632 B.
NCU_TRAFFIC_GATE_UP = 275_672_576

METRIC = "VQA decode tok/s @14B (B=8/GPU, 448x448, 128-token greedy)"
WORKLOAD = ("VQA batch=8 per GPU, 448x448 (1024 ViT tokens + 2 markers), 32-token prompt, 14B MoT bf16 "
            "(both experts resident), 128-token greedy decode, ctx 1058->1185")      # BASELINE.json configs[1], both arms


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", 1400.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


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


class IdTokenizer:
    """Prompts are given as token ids (no vocabulary files offline): an id list, or -- where the reference's drivers insist on `str` inputs
    -- the ids written out as a space-separated string.  decode joins the ids (eos / bos as the reference's strings)."""

    def __init__(self, tok):
        self.m = {tok["bos_token_id"]: "<|im_start|>", tok["eos_token_id"]: "<|im_end|>"}

    def encode(self, ids):
        return [int(t) for t in ids.split()] if isinstance(ids, str) else list(ids)

    def decode(self, ids):
        return " ".join(self.m.get(int(i), str(int(i))) for i in ids)


def synthetic_requests(first: int, n: int, size: int = IMG):
    """n synthetic requests: uint8 noise image (pinned host tensor + PIL view) and 30 prompt ids."""
    import torch
    from PIL import Image
    from unimedvl_b200 import synth
    raws = [synth.synthetic_image(first + i, size, size) for i in range(n)]
    pinned = [torch.from_numpy(r.copy()).pin_memory() if torch.cuda.is_available() else torch.from_numpy(r.copy()) for r in raws]
    ids = [synth.synthetic_prompt_ids(first + i, PROMPT_TOKENS) for i in range(n)]
    return pinned, [Image.fromarray(r) for r in raws], ids, [" ".join(str(t) for t in p) for p in ids]


def synthetic_job(rank: int):
    """Host-preprocessed inputs of one configs[1] batch (tools/prefill_trace.py, e2e_phases.py, decode_after_prefill.py): the images of
    synthetic_requests through the reference's transform + patchify on the host, pinned; 30 prompt ids per sample."""
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
    return torch.cat(toks, 0).pin_memory(), torch.cat(pos, 0).pin_memory(), lens, prompts, images


# ================================================================================================ reference (CPU / GPU) helpers
def _refharness():
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import refharness as rh
    return rh if rh.ref_path() is not None else None


def reference_cpu_config1(model, vae, tok):
    """BASELINE.json configs[0] / SURVEY.md section 8d "Config 1": the unmodified reference through InterleaveInferencer.__call__ on the
    host cores -- one 448x448 image + 30-token question -> 32 new tokens (max_think_token_n=33), all 28 layers."""
    import torch
    rh = _refharness()
    R = rh.load()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    _, pil, _, prompts = synthetic_requests(0, 1)
    inf = R.InterleaveInferencer(model, vae, IdTokenizer(tok), R.ImageTransform(1024, 32, 16), R.ImageTransform(980, 378, 14, max_pixels=2_007_040), tok)
    t = {"prefill": 0.0, "decode": 0.0}

    def timed(name, fn):
        def wrap(*a, **k):
            t0 = time.perf_counter()
            r = fn(*a, **k)
            t[name] += time.perf_counter() - t0
            return r
        return wrap
    model.forward_cache_update_vit = timed("prefill", model.forward_cache_update_vit)
    model.forward_cache_update_text = timed("prefill", model.forward_cache_update_text)
    model.generate_text = timed("decode", model.generate_text)
    t0 = time.perf_counter()
    with torch.no_grad(), rh.autocast("cpu"):
        out = inf(image=pil[0], text=prompts[0], understanding_output=True, do_sample=False, max_think_token_n=33)
    wall = time.perf_counter() - t0
    n_tok = max(1, len(out["text"].split()))
    return {"value": round(n_tok / t["decode"], 3), "unit": "tok/s", "cores": cores, "kind": "reference",
            "sample": "BASELINE configs[0] in full: unmodified reference (baseline/_ref/codes) InterleaveInferencer.__call__ on CPU, bf16 params, "
                      f"CPU autocast, 1 x 448x448 + 32-token prompt -> {n_tok} greedy tokens, 28 layers; run once (no warm-up)",
            "batch": 1, "prefill_s": round(t["prefill"], 2), "decode_s": round(t["decode"], 2), "wall_s": round(wall, 2),
            "s_per_token": round(t["decode"] / n_tok, 4)}


def oracle_port_decode_baseline(n_layers_sample: int = 2, steps: int = 3, ctx: int = 1058):
    """Fallback when no reference copy travels with the repo: the oracle port of one decode step, 2 of 28 layers (EXTRAPOLATED x14)."""
    import torch
    from oracle import llm as ollm
    from unimedvl_b200 import config as ucfg
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    L = ucfg.bagel_7b_mot().llm
    dims = ollm.LLMDims(L.hidden, L.heads, L.kv_heads, L.inter, n_layers_sample, L.vocab, L.rope_theta, L.eps)
    g = torch.Generator().manual_seed(0)
    rnd = lambda *s: (torch.randn(*s, generator=g, dtype=torch.float32) * 0.02).to(torch.bfloat16)
    dh = L.head_dim
    sd = {"language_model.model.embed_tokens.weight": rnd(4096, L.hidden), "language_model.lm_head.weight": rnd(L.vocab, L.hidden),
          "language_model.model.norm.weight": torch.ones(L.hidden, dtype=torch.bfloat16)}
    for i in range(n_layers_sample):
        P = f"language_model.model.layers.{i}."
        for n, rows in (("q", L.heads * dh), ("k", L.kv_heads * dh), ("v", L.kv_heads * dh)):
            sd[P + f"self_attn.{n}_proj.weight"] = rnd(rows, L.hidden)
            sd[P + f"self_attn.{n}_proj.bias"] = rnd(rows)
        sd[P + "self_attn.o_proj.weight"] = rnd(L.hidden, L.heads * dh)
        for n in ("q_norm", "k_norm"):
            sd[P + f"self_attn.{n}.weight"] = torch.ones(dh, dtype=torch.bfloat16)
        sd[P + "mlp.gate_proj.weight"] = rnd(L.inter, L.hidden)
        sd[P + "mlp.up_proj.weight"] = rnd(L.inter, L.hidden)
        sd[P + "mlp.down_proj.weight"] = rnd(L.hidden, L.inter)
        for n in ("input_layernorm", "post_attention_layernorm"):
            sd[P + n + ".weight"] = torch.ones(L.hidden, dtype=torch.bfloat16)
    B = B_PER_GPU
    cache = ollm.PackedKV(n_layers_sample)
    for i in range(n_layers_sample):
        cache.key[i], cache.value[i] = rnd(B * ctx, L.kv_heads, dh), rnd(B * ctx, L.kv_heads, dh)
    kv_lens = torch.full((B,), ctx, dtype=torch.int64)
    kvidx = torch.arange(B * ctx)
    pos = torch.full((B,), 33, dtype=torch.int64)
    toks = torch.randint(0, 4096, (B,), generator=g)
    t_layers = t_head = 0.0
    with torch.no_grad():
        for s in range(steps + 1):
            qidx = torch.cumsum(kv_lens, 0) + torch.arange(B)
            kvidx = torch.cat([p + i for i, p in enumerate(kvidx.split(kv_lens.tolist()))])
            t0 = time.perf_counter()
            h, cache = ollm.forward_inference(sd, dims, ollm.embed(sd, toks), torch.ones(B, dtype=torch.int64), pos, qidx, cache,
                                              kv_lens, kvidx, True, True, exact=False)
            t1 = time.perf_counter()
            nxt = torch.argmax(ollm.lm_head(sd, h, exact=False), -1) % 4096
            t2 = time.perf_counter()
            if s > 0:
                t_layers += t1 - t0
                t_head += t2 - t1
            kvidx = torch.cat([torch.cat([p, p[-1:] + 1]) for p in kvidx.split(kv_lens.tolist())])
            kv_lens, pos, toks = kv_lens + 1, pos + 1, nxt
    step_s = (t_layers / steps) * (L.layers / n_layers_sample) + t_head / steps
    return {"value": round(B / step_s, 3), "unit": "tok/s", "cores": cores, "kind": "port",
            "sample": f"EXTRAPOLATED: oracle/ decode forward, B={B}, ctx={ctx}, {n_layers_sample} of {L.layers} layers x{steps} steps "
                      f"(layer time scaled x{L.layers // n_layers_sample}) + full lm_head; native bf16 CPU GEMM",
            "ms_per_step": round(step_s * 1e3, 1)}


def run_reference(args, rank: int, world: int):
    """--impl reference: the unmodified reference's CPU forward of the headline metric.  One step = Bagel.generate_text(max_length=1) --
    one decode forward of all 28 layers + lm_head + argmax -- for 8 samples on a 1058-token context; warm-up + K steps, wall clock."""
    if rank != 0:
        return
    import torch
    t_start = time.perf_counter()
    rh = _refharness()
    if rh is None:
        base = None
        for i in range(args.warmup + args.steps):
            r = oracle_port_decode_baseline(steps=2)
            if i >= args.warmup and (base is None or r["value"] > base["value"]):
                base = r
        value, ms_step = base["value"], base["ms_per_step"]
        arm = "CPU forward of the reference algorithm (oracle port, no reference copy present), rank 0 only"
    else:
        from unimedvl_b200 import config as ucfg
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        dims, tok = ucfg.bagel_7b_mot(), dict(ucfg.QWEN25_TOKEN_IDS)
        model, _ = rh.build_reference(dims, None, None, "cpu", fill=rh.random_fill("cpu"))
        R = rh.load()
        B, ctx, L = B_PER_GPU, (IMG // 14) ** 2 + 2 + PROMPT_TOKENS + 2, dims.llm.layers
        cache = R.NaiveCache(L)
        g = torch.Generator().manual_seed(0)
        for i in range(L):
            cache.key_cache[i] = (torch.randn(B * ctx, dims.llm.kv_heads, dims.llm.head_dim, generator=g) * 0.5).to(torch.bfloat16)
            cache.value_cache[i] = (torch.randn(B * ctx, dims.llm.kv_heads, dims.llm.head_dim, generator=g) * 0.5).to(torch.bfloat16)
        kvl, rope = [ctx] * B, [PROMPT_TOKENS + 3] * B
        times = []
        with torch.no_grad(), rh.autocast("cpu"):
            for i in range(args.warmup + args.steps):
                gi = model.prepare_start_tokens(kvl, rope, tok)
                t0 = time.perf_counter()
                model.generate_text(past_key_values=cache, max_length=1, do_sample=False, end_token_id=None, **gi)
                if i >= args.warmup:
                    times.append(time.perf_counter() - t0)
                kvl, rope = [k + 1 for k in kvl], [r + 1 for r in rope]       # generate_text appended one token per sample
        step = sum(times) / len(times)
        value, ms_step = round(B / step, 3), round(step * 1e3, 1)
        base = {"value": value, "unit": "tok/s", "cores": cores, "kind": "reference",
                "sample": f"unmodified reference Bagel.generate_text(max_length=1) on CPU: one decode forward of all 28 layers + lm_head, B={B}, "
                          f"ctx {ctx}+, bf16 params, CPU autocast; mean of {len(times)} steps after {args.warmup} warm-up",
                "ms_per_step": ms_step}
        arm = "the reference's own CPU forward (baseline/_ref/codes, unmodified), rank 0 only"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": {"workload": WORKLOAD, "arm": arm}, "cpu_baseline": base,
            "e2e": {"value": value, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "wall_s": round(time.perf_counter() - t_start, 1)}
    print(json.dumps(line), flush=True)


def gpu_reference_leg(tok, engine_t2i_b1=None):
    """The unmodified reference on the same B200 (real flash_attn_varlen_func, CUDA autocast): configs[1] through its packed API
    (prepare_vit_images -> forward_cache_update_vit -> prepare_prompts -> forward_cache_update_text -> generate_text) and one
    256x256 text-to-image request through InterleaveInferencer.__call__.  Returns (dict, model, vae) -- the model is reused for the
    CPU baseline."""
    import torch
    from unimedvl_b200 import config as ucfg
    rh = _refharness()
    R = rh.load()
    dims = ucfg.bagel_7b_mot()
    model, vae = rh.build_reference(dims, None, None, "cuda", fill=rh.random_fill("cuda"))
    _, pil, prompts, prompt_strs = synthetic_requests(0, B_PER_GPU)
    B, L = B_PER_GPU, dims.llm.layers
    vit_tf = R.ImageTransform(980, 378, 14, max_pixels=2_007_040)
    itok = IdTokenizer(tok)
    cuda = lambda d: {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()}

    def vqa():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        with torch.no_grad(), rh.autocast("cuda"):
            ev[0].record()
            cache = R.NaiveCache(L)
            gv, lens, rope = model.prepare_vit_images([0] * B, [0] * B, pil, vit_tf, tok)
            cache = model.forward_cache_update_vit(cache, **cuda(gv))
            gt, lens, rope = model.prepare_prompts(lens, rope, prompts, itok, tok)
            cache = model.forward_cache_update_text(cache, **cuda(gt))
            ev[1].record()
            gs = model.prepare_start_tokens(lens, rope, tok)
            toks = model.generate_text(past_key_values=cache, max_length=DECODE_STEPS, do_sample=False, end_token_id=None, **cuda(gs))
            ev[2].record()
        torch.cuda.synchronize()
        assert tuple(toks.shape) == (DECODE_STEPS, B)
        return ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    vqa()                                                   # warm-up (cuBLAS / flash-attn autotune, allocator)
    pre, dec = vqa()
    inf = R.InterleaveInferencer(model, vae, itok, R.ImageTransform(1024, 32, 16), vit_tf, tok)
    kw = dict(understanding_output=False, num_timesteps=T2I_STEPS, image_shapes=(T2I_SIZE, T2I_SIZE), cfg_text_scale=4.0, cfg_img_scale=1.5)

    def t2i():
        torch.manual_seed(42)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad(), rh.autocast("cuda"):
            r = inf(text=prompt_strs[0], **kw)
        torch.cuda.synchronize()
        assert r["image"].size == (T2I_SIZE, T2I_SIZE)
        return time.perf_counter() - t0
    t2i()
    t_img = t2i()
    out = {"what": "unmodified reference (baseline/_ref/codes), real flash_attn_varlen_func, CUDA autocast bf16, random-init 14B weights, 1 x B200",
           "decode_tok_s": round(B * DECODE_STEPS / (dec / 1e3), 1), "ms_per_decode_forward": round(dec / DECODE_STEPS, 3),
           "prefill_ms": round(pre, 1), "e2e_tok_s": round(B * DECODE_STEPS / ((pre + dec) / 1e3), 1),
           "t2i_b1_img_s": round(1.0 / t_img, 4), "t2i_b1_s": round(t_img, 3), "timing": "second of two runs; CUDA events (VQA), wall clock around a synchronize (T2I)"}
    if engine_t2i_b1 is not None:
        out["engine_t2i_b1_img_s"] = engine_t2i_b1
    return out, model, vae


# ================================================================================================ engine legs
def run_engine(args, rank: int, local_rank: int, world: int):
    import ctypes as C
    from copy import deepcopy

    import numpy as np
    import torch
    import torch.distributed as dist
    from unimedvl_b200 import _lib, config as ucfg, dp, packing
    from unimedvl_b200.autoencoder import AutoEncoder
    from unimedvl_b200.bagel import Bagel
    from unimedvl_b200.batched import BatchedInferencer
    from unimedvl_b200.cache import NaiveCache
    from unimedvl_b200.engine import Engine

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    legs = set(args.legs.split(",")) if args.legs else {"report_gen", "t2i", "interleaved", "gpu_reference", "cpu_baseline"}
    dims = ucfg.bagel_7b_mot()
    tok = dict(ucfg.QWEN25_TOKEN_IDS)
    hbm_peak, tensor_peak, peak_src = _peaks()
    B = B_PER_GPU
    ntok_img = (IMG // 14) ** 2 + 2
    ctx0 = ntok_img + PROMPT_TOKENS + 2
    eng = Engine(dims, max_tokens=B * (ntok_img + PROMPT_TOKENS + 2), max_seqs=max(B_REPORT, B), kv_pages=1152, enable_vit=True, enable_gen=True, enable_vae=True)
    eng.fill_synthetic(seed=0)          # random-init weights of the reference architecture: both experts, ViT, VAE resident
    eng.finalize()
    model = Bagel(eng, dims)
    vae = AutoEncoder(eng)
    itok = IdTokenizer(tok)
    vit_tf = packing.ImageTransform(980, 378, 14, max_pixels=2_007_040)
    bi = BatchedInferencer(model, vae, itok, packing.ImageTransform(1024, 32, 16), vit_tf, tok)
    images, pil, prompts, prompt_strs = synthetic_requests(rank * 64, max(B_REPORT, B))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K calls bracketed by barrier + synchronize, CUDA events on the current stream, max over ranks -> ms."""
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

    def prefill_vqa(n):
        """ViT + image prefill + prompt prefill of the first n requests, 8 at a time (the workspace holds 8 x 1026 rows)."""
        parts = []
        for lo in range(0, n, B):
            pix, pos, lens = eng.patchify_u8(images[lo:lo + B])
            k = len(lens)
            L = packing._image_block_layout([0] * k, [0] * k, lens, tok)
            g = dict(packed_text_ids=torch.as_tensor(L["text_ids"]), packed_text_indexes=torch.as_tensor(L["text_idx"]), packed_vit_tokens=pix,
                     packed_vit_token_indexes=torch.as_tensor(L["img_idx"]), packed_vit_position_ids=pos, vit_token_seqlens=lens,
                     packed_position_ids=torch.as_tensor(L["pos"]), packed_seqlens=L["seqlens"], packed_indexes=torch.as_tensor(L["packed_idx"]),
                     packed_key_value_indexes=torch.as_tensor(L["kv_indexes"]), key_values_lens=[0] * k)
            c = model.forward_cache_update_vit(NaiveCache(dims.llm.layers), **g)
            gp, kvl, rope = packing.prepare_prompts(L["seqlens"], [1] * k, prompts[lo:lo + k], itok, tok)
            parts.append((model.forward_cache_update_text(c, **gp), kvl, rope))
        cache = NaiveCache.concat([p[0] for p in parts])
        kvl = [v for p in parts for v in p[1]]
        rope = [v for p in parts for v in p[2]]
        return cache, packing.prepare_start_tokens(kvl, rope, tok), kvl

    # ---------------------------------------------------------------- headline: configs[1] decode from a resident context
    cache, start, kvl = prefill_vqa(B)
    assert kvl == [ctx0] * B

    def decode_step():
        c = deepcopy(cache)                       # page fork (the reference deep-copies the KV, inferencer.py:261)
        t = model.generate_text(past_key_values=c, max_length=DECODE_STEPS, end_token_id=None, **start)
        return dp.gather_tokens(t, world * B)       # one NCCL all_gather of the output tokens

    def e2e_step():
        t = model.vqa_generate_images(images[:B], prompts[:B], tok, DECODE_STEPS)
        return dp.gather_tokens(t.cuda(), world * B) if world > 1 else t

    warm = max(args.warmup, 3)
    for _ in range(warm):
        decode_step()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = eng.launch_count()
    ms = timed(decode_step, args.steps)
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    tokens_per_step = world * B * DECODE_STEPS
    value = tokens_per_step * args.steps / (ms / 1e3)
    for _ in range(2):
        e2e_step()
    e2e_steps = max(1, min(args.steps, 5))
    ms_e2e = timed(e2e_step, e2e_steps)
    e2e_value = tokens_per_step * e2e_steps / (ms_e2e / 1e3)
    h2d = sum(im.numel() for im in images[:B]) + sum(len(p) + 2 for p in prompts[:B]) * 8 + B * 2 * 8
    d2h = B * DECODE_STEPS * 8
    D, I = dims.llm.hidden, dims.llm.inter
    w_bytes = 2 * (dims.llm.layers * 233_058_048 + dims.llm.vocab * D + D)

    def step_bytes(b, ctx_mean):            # weights + KV read + KV write + embedding rows
        return w_bytes + b * ctx_mean * 57_344 + b * 57_344 + b * D * 2

    # ---------------------------------------------------------------- roofline of the dominant kernel (rank 0)
    roof = in_graph = None
    if rank == 0:
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
        alg_bytes = wb.value + B * D * 2 + B * I * 2
        # the same kernel INSIDE the replayed decode graph: %globaltimer stamps of the last step (umv_trace_*), critical-path share =
        # this launch's last-CTA end minus the previous launch's last-CTA end, averaged over the 28 layers
        try:
            CAP, NL = 1024, 32
            _lib.check(eng.lib.umv_trace_begin(CAP))
            c = deepcopy(cache)
            model.generate_text(past_key_values=c, max_length=8, end_token_id=None, **start)
            stamps = np.zeros((CAP, 12), dtype=np.uint64)
            names = C.create_string_buffer(CAP * NL)
            n = C.c_int32()
            _lib.check(eng.lib.umv_trace_read(stamps.ctypes.data_as(C.c_void_p), names, NL, CAP, C.byref(n)))
            _lib.check(eng.lib.umv_trace_begin(0))
            nm = [names.raw[i * NL:(i + 1) * NL].split(b"\0")[0].decode() for i in range(n.value)]
            t = stamps[:n.value].astype(np.int64)
            agg = {}
            for i in range(1, n.value):
                a = agg.setdefault(nm[i], [0, 0.0])
                a[0] += 1
                a[1] += (t[i, 3] - t[i - 1, 3]) / 1e3
            top = max(agg.items(), key=lambda kv: kv[1][1])
            in_graph = {"kernel_trace_name": top[0], "launches": top[1][0], "us_per_launch": round(top[1][1] / top[1][0], 2),
                        "forward_us": round((t[:, 3].max() - t[0, 0]) / 1e3, 1)}
            del c
        except Exception as ex:            # tracing is diagnostic: never fail the bench on it
            in_graph = {"error": str(ex)[:200]}
        roof = {"bound": "hbm", "kernel": "gemm_tc_kernel<16,2,true> (weight-major gate/up + SwiGLU, M=8)",
                "achieved": round(alg_bytes / (k_ms / 1e3) / 1e9, 1), "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": round(alg_bytes / (k_ms / 1e3) / 1e9 / hbm_peak, 4), "traffic": NCU_TRAFFIC_GATE_UP,
                "traffic_source": "profiles/r2_decode_kernels_full.md (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum of this kernel, per launch)",
                "algorithmic_bytes_per_launch": int(alg_bytes), "us_per_launch": round(k_ms * 1e3, 2),
                "us_per_launch_source": "CUDA events over 8 x 28 back-to-back launches on 28 distinct layers' weights", "in_graph": in_graph}
        if in_graph and "us_per_launch" in in_graph:
            roof["frac_in_graph"] = round(alg_bytes / (in_graph["us_per_launch"] / 1e6) / 1e9 / hbm_peak, 4)
    step_ms = ms / args.steps / DECODE_STEPS
    sb = step_bytes(B, ctx0 + (DECODE_STEPS - 1) / 2)
    del cache

    # ---------------------------------------------------------------- configs[2]: report generation, 16 per GPU, 512 steps
    extra = {}
    if "report_gen" in legs:
        cache, start, kvl = prefill_vqa(B_REPORT)

        def report_step():
            c = deepcopy(cache)
            t = model.generate_text(past_key_values=c, max_length=REPORT_STEPS, end_token_id=None, **start)
            return dp.gather_tokens(t, world * B_REPORT)
        report_step()
        ms_r = timed(report_step, 2) / 2
        fwd = ms_r / REPORT_STEPS
        rb = step_bytes(B_REPORT, ctx0 + (REPORT_STEPS - 1) / 2)
        extra["report_gen"] = {
            "config": f"BASELINE configs[2] per-GPU share: {B_REPORT} samples per GPU, {REPORT_STEPS}-token greedy decode over the paged KV, ctx "
                      f"{ctx0}->{ctx0 + REPORT_STEPS - 1}, one NCCL all_gather of the tokens; device-timed, max over ranks, 2 runs after 1 warm-up",
            "value": round(world * B_REPORT * REPORT_STEPS / (ms_r / 1e3), 1), "unit": "tok/s", "ms_per_decode_forward": round(fwd, 4),
            "roofline": {"bound": "hbm", "algorithmic_bytes_per_decode_forward": int(rb), "achieved": round(rb / (fwd / 1e3) / 1e9, 1), "unit": "GB/s",
                         "peak": hbm_peak, "frac": round(rb / (fwd / 1e3) / 1e9 / hbm_peak, 4),
                         "roofline_tok_s_per_gpu": round(B_REPORT / (rb / (hbm_peak * 1e9)), 1)}}
        del cache

    # ---------------------------------------------------------------- configs[3]: text-to-image, 4 per GPU, the whole job
    t2i_flops_img = 0.0
    for steps_on, branches in ((41, 3), (8, 1)):
        t2i_flops_img += steps_on * branches * (2 * 258 * 6.5253e9 + 4 * 258 * (32 + 258) * 3584 * 28)
    t2i_flops_img += 0.62e12                                                   # VAE decode
    t2i_kw = dict(num_timesteps=T2I_STEPS, timestep_shift=3.0, cfg_text_scale=4.0, cfg_img_scale=1.5, cfg_interval=(0.4, 1.0),
                  cfg_renorm_min=0.0, cfg_renorm_type="global")

    def t2i_job(n):
        """prompts -> uint8 images on the host: what interleave_inference does for [text] (inferencer.py:552-638), packed over n."""
        ps = [prompt_strs[i % len(prompt_strs)] for i in range(n)]
        ctx = bi.update_context_text(ps, bi.init_gen_context(n))
        cfg_img = deepcopy(ctx)        # text only: the image-free context holds the same tokens -- a page fork, not a second prefill
        torch.manual_seed(42)
        u8 = bi.gen_image((T2I_SIZE, T2I_SIZE), ctx, cfg_text_precontext=bi.init_gen_context(n), cfg_img_precontext=cfg_img, as_uint8=True,
                          **t2i_kw)
        allu8 = dp.gather_images(u8, world * n)
        host = torch.empty(allu8.shape, dtype=torch.uint8, pin_memory=True)
        host.copy_(allu8, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host
    engine_t2i_b1 = None
    if "t2i" in legs:
        # the three CFG branches are all evaluated here, as the reference does.  (The engine can evaluate the cfg_img branch of a pure
        # text-to-image request once -- its context holds the main context's tokens -- with bit-identical images: reported separately
        # under "branch_dedup", not in `value`.)
        os.environ["UMV_CFG_DEDUP"] = "0"
        t2i_job(B_T2I)
        ms_t = timed(lambda: t2i_job(B_T2I), 2) / 2
        img = t2i_job(B_T2I)
        os.environ["UMV_CFG_DEDUP"] = "1"
        t2i_job(B_T2I)
        ms_d = timed(lambda: t2i_job(B_T2I), 2) / 2
        img_d = t2i_job(B_T2I)
        os.environ.pop("UMV_CFG_DEDUP", None)
        flops_d = (41 * 2 + 8) * (2 * 258 * 6.5253e9 + 4 * 258 * (32 + 258) * 3584 * 28) + 0.62e12
        tf = t2i_flops_img * B_T2I / (ms_t / 1e3) / 1e12
        extra["t2i"] = {
            "config": f"BASELINE configs[3] per-GPU share, the whole job: {B_T2I} prompts per GPU -> {T2I_SIZE}x{T2I_SIZE}: prompt prefill (the cfg_img "
                      "context is a page fork of it), 50 timesteps (41 x 3 CFG branches + 8 x 1), shift 3.0, CFG 4.0 / 1.5 on (0.4, 1], per-image global renorm, VAE decode, "
                      "uint8 conversion on the device, NCCL all_gather of the images, D2H; device-timed, max over ranks, 2 runs after 1 warm-up",
            "value": round(world * B_T2I / (ms_t / 1e3), 4), "unit": "img/s", "ms_per_batch": round(ms_t, 1), "images_per_gpu": B_T2I,
            "algorithmic_tflop_per_image": round(t2i_flops_img / 1e12, 1), "d2h_bytes_per_step": int(img.numel()),
            "roofline": {"bound": "tensor", "achieved": round(tf, 1), "peak": tensor_peak, "unit": "TFLOP/s", "frac": round(tf / tensor_peak, 4),
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained"},
            "image_mean_grey": round(float(img.float().mean()), 2),
            "branch_dedup": {
                "what": "same job with the cfg_img branch evaluated once (umv_flow_velocity: the image-free context is a fork of the main context "
                        "-- pure text-to-image -- so its velocity is the main branch's bit for bit): 41 x 2 + 8 forwards instead of 41 x 3 + 8",
                "value": round(world * B_T2I / (ms_d / 1e3), 4), "unit": "img/s", "ms_per_batch": round(ms_d, 1),
                "images_identical": bool(torch.equal(img, img_d)),
                "algorithmic_tflop_per_image": round(flops_d / 1e12, 1),
                "roofline_frac": round(flops_d * B_T2I / (ms_d / 1e3) / 1e12 / tensor_peak, 4)}}
        if world == 1 and "gpu_reference" in legs:
            t2i_job(1)
            engine_t2i_b1 = round(1.0 / (timed(lambda: t2i_job(1), 1) / 1e3), 4)

    # ---------------------------------------------------------------- configs[4]: interleaved I2T + image-conditioned T2I, 8 per GPU
    if "interleaved" in legs:
        reqs = [[pil[i], prompt_strs[i]] for i in range(B_INTER)]
        edit_kw = dict(understanding_output=False, num_timesteps=T2I_STEPS, image_shapes=(T2I_SIZE, T2I_SIZE), cfg_text_scale=4.0,
                       cfg_img_scale=2.0, cfg_interval=[0.0, 1.0], cfg_renorm_type="text_channel", return_uint8=True)

        def inter_job():
            texts = bi.interleave_inference(reqs, understanding_output=True, do_sample=False, max_think_token_n=INTER_TOKENS + 1)
            torch.manual_seed(43)
            outs = bi.interleave_inference(reqs, **edit_kw)
            u8 = torch.stack([o[-1] for o in outs], 0)
            allu8 = dp.gather_images(u8, world * B_INTER)
            host = torch.empty(allu8.shape, dtype=torch.uint8, pin_memory=True)
            host.copy_(allu8, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return texts, host
        inter_job()
        ms_i = timed(inter_job, 1)
        texts, host = inter_job()
        n_tok = sum(len(t[0].split()) for t in texts)
        extra["interleaved"] = {
            "config": f"BASELINE configs[4] per-GPU share through BatchedInferencer: {B_INTER} requests per GPU, each (1) 448x448 image + 32-token question "
                      f"-> {INTER_TOKENS} greedy tokens, (2) the same image (VAE + ViT context, 1,812 tokens) + instruction -> {T2I_SIZE}x{T2I_SIZE} image, 49 x 3 "
                      "CFG forwards, CFG 4.0 / 2.0 on [0, 1], text_channel renorm, VAE encode + decode; host PIL inputs, H2D, D2H of tokens and uint8 images, "
                      "NCCL all_gather of the images inside; one timed run after one warm-up, max over ranks",
            "value": round(world * B_INTER / (ms_i / 1e3), 4), "unit": "requests/s", "ms_per_batch": round(ms_i, 1), "requests_per_gpu": B_INTER,
            "decoded_tokens": int(n_tok), "images": int(host.shape[0]), "image_mean_grey": round(float(host.float().mean()), 2)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- reference on the same GPU, reference on the host cores (N == 1)
    gpu_ref = cpu = None
    ref_model = ref_vae = None
    if world == 1 and _refharness() is not None and "gpu_reference" in legs:
        try:
            gpu_ref, ref_model, ref_vae = gpu_reference_leg(tok, engine_t2i_b1)
            gpu_ref["engine_decode_tok_s"] = round(value, 1)
            gpu_ref["engine_e2e_tok_s"] = round(e2e_value, 1)
        except Exception as ex:
            gpu_ref = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}
    if world == 1 and "cpu_baseline" in legs:
        rh = _refharness()
        try:
            if rh is None:
                cpu = oracle_port_decode_baseline()
            else:
                if ref_model is not None:                   # weights already drawn on the GPU: bring them over instead of a host RNG pass
                    ref_model, ref_vae = ref_model.to("cpu"), ref_vae.to("cpu")
                    torch.cuda.empty_cache()
                    R = rh.load()
                    R.qn.flash_attn_varlen_func = R.sn.flash_attn_varlen_func = rh.sdpa_varlen      # flash-attn is CUDA-only (shim S3)
                else:
                    ref_model, ref_vae = rh.build_reference(dims, None, None, "cpu", fill=rh.random_fill("cpu"))
                cpu = reference_cpu_config1(ref_model, ref_vae, tok)
        except Exception as ex:
            cpu = {"error": f"{type(ex).__name__}: {str(ex)[:300]}", "kind": "reference", "value": None, "unit": "tok/s", "cores": os.cpu_count()}

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "tok/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "parallelism": f"dp{world}", "l2": "weights streamed per decode forward (14.1 GB) exceed the 126 MB L2",
                   "weights": "random-init, generated on device", "step": "generate_text(max_length=128) for the batch"},
        "ms_per_decode_forward": round(step_ms, 4),
        "e2e": {"value": round(e2e_value, 1), "unit": "tok/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": round(ms_e2e / e2e_steps, 2), "steps": e2e_steps,
                "path": "Bagel.vqa_generate_images: pinned host uint8 images + prompt ids -> H2D -> normalise/patchify on device -> ViT -> "
                        "image+prompt prefill -> 128-step decode -> host tokens"},
        "gpu_launches": int(launches), "roofline": roof,
        "step_roofline": {"bound": "hbm", "algorithmic_bytes_per_decode_forward": int(sb), "achieved": round(sb / (step_ms / 1e3) / 1e9, 1),
                          "unit": "GB/s", "frac": round(sb / (step_ms / 1e3) / 1e9 / hbm_peak, 4),
                          "roofline_tok_s_per_gpu": round(B / (sb / (hbm_peak * 1e9)), 1)},
        "clocks": clocks,
    }
    line.update(extra)
    if gpu_ref is not None:
        line["gpu_reference"] = gpu_ref
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
    ap.add_argument("--legs", default="", help="comma list of extra legs to run (default: all): report_gen,t2i,interleaved,gpu_reference,cpu_baseline")
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
