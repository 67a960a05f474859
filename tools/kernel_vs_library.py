"""The engine's kernels next to the library kernels the reference calls at the same shapes, on the same GPU (SURVEY.md section 2.2's
bars: cuBLAS K1-K4, flash-attn 2 @ sm_100 K8 / K10, cuDNN K17).  Writes a markdown table.

    python tools/kernel_vs_library.py [out.md]

  linears    umv_op_linear (tcgen05, fused epilogue) vs torch.matmul (cuBLAS, no epilogue: bias / activation / residual are extra kernels
             in the reference) at the 14B prefill / flow / ViT / decode shapes
  attention  attn_tc_kernel through umv_op_attention_block (the model path's whole block: q/k-norm + RoPE + KV append + attention, host
             metadata upload included) vs flash_attn_varlen_func (the attention kernel alone) at the same geometry
  VAE        umv_vae_decode / umv_vae_encode_moments vs the reference's AutoEncoder (cuDNN convolutions, bf16, CUDA autocast) at 256 x 256
             -- only when the reference copy (baseline/_ref) is present
Each figure: mean of 10 back-to-back launches after 3 warm-ups, CUDA events.
"""
import os
import sys

import ctypes as C

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))
from unimedvl_b200 import config as ucfg  # noqa: E402
from unimedvl_b200 import _lib  # noqa: E402
from unimedvl_b200.engine import Engine, op_linear  # noqa: E402

out = []


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3          # us


def traced(eng, fn, cap=64, nl=32):
    """One call of fn under the engine's in-kernel %globaltimer trace: {kernel name: us from its first CTA's start to its last CTA's end}."""
    fn()
    torch.cuda.synchronize()
    _lib.check(eng.lib.umv_trace_begin(cap))
    fn()
    torch.cuda.synchronize()
    stamps = np.zeros((cap, 12), dtype=np.uint64)
    names = C.create_string_buffer(cap * nl)
    n = C.c_int32()
    _lib.check(eng.lib.umv_trace_read(stamps.ctypes.data_as(C.c_void_p), names, nl, cap, C.byref(n)))
    _lib.check(eng.lib.umv_trace_begin(0))
    t = stamps[:n.value].astype(np.int64)
    res = {}
    for i in range(n.value):
        nm = names.raw[i * nl:(i + 1) * nl].split(b"\0")[0].decode()
        res[nm] = res.get(nm, 0.0) + (t[i, 3] - t[i, 0]) / 1e3
    return res


# ------------------------------------------------------------------------------------------------ linears
out.append("## linears: `umv_op_linear` vs cuBLAS (`torch.matmul`, no epilogue)\n\n| what | M | N | K | epilogue | engine us | engine TFLOP/s (or GB/s of weights) | cuBLAS us | cuBLAS TFLOP/s | engine / cuBLAS time |\n|---|---:|---:|---:|---|---:|---:|---:|---:|---:|\n")
EPI = {0: "bias", 1: "bias + GELU", 2: "SwiGLU", 3: "bias + residual"}
shapes = [("image prefill q|k|v", 8208, 4608, 3584, 0), ("image prefill o_proj", 8208, 3584, 3584, 3), ("image prefill gate|up", 8208, 37888, 3584, 2),
          ("image prefill down", 8208, 3584, 18944, 3), ("flow step gate|up (4 img x 3)", 3096, 37888, 3584, 2), ("flow step down", 3096, 3584, 18944, 3),
          ("flow step q|k|v", 3096, 4608, 3584, 0), ("ViT q|k|v", 8192, 3456, 1152, 0), ("ViT fc1", 8192, 4304, 1152, 1), ("ViT fc2", 8192, 1152, 4304, 3),
          ("decode gate|up (B=8)", 8, 37888, 3584, 2), ("decode down (B=8)", 8, 3584, 18944, 0), ("decode lm_head (B=8)", 8, 152064, 3584, 0)]
for (name, M, N, K, epi) in shapes:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = None if epi == 2 else torch.zeros(N, device="cuda").bfloat16()
    res = torch.zeros(M, N, device="cuda").bfloat16() if epi == 3 else None
    us = timeit(lambda: op_linear(x, w, b, res, epi=epi))
    us_ref = timeit(lambda: torch.matmul(x, w.T))
    fl = 2.0 * M * N * K
    perf = f"{fl / us / 1e6:.0f} TFLOP/s" if M > 64 else f"{N * K * 2 / us / 1e3:.0f} GB/s"
    out.append(f"| {name} | {M} | {N} | {K} | {EPI[epi]} | {us:.1f} | {perf} | {us_ref:.1f} | {fl / us_ref / 1e6:.0f} | {us / us_ref:.2f} |\n")
    del x, w, b, res
torch.cuda.empty_cache()

# ------------------------------------------------------------------------------------------------ attention
H, HKV, DH = 28, 4, 128
QN = (H + 2 * HKV) * DH
dims = ucfg.BagelDims(llm=ucfg.LLMDims(hidden=H * DH, heads=H, kv_heads=HKV, inter=128, layers=1, vocab=1024), vit=ucfg.ViTDims(hidden=144, heads=2, inter=328, layers=1))
eng = Engine(dims, max_tokens=8300, max_seqs=16, kv_pages=512, enable_vit=False, enable_gen=False)
eng.fill_synthetic(1)
eng.finalize()
try:
    from flash_attn import flash_attn_varlen_func
except Exception as ex:                      # pragma: no cover
    flash_attn_varlen_func = None
    out.append(f"\n(flash_attn not importable: {ex})\n")
out.append("\n## attention: `attn_tc_kernel` (tcgen05) vs `flash_attn_varlen_func` (flash-attn 2.8, sm_100 build), 28 / 4 heads x 128\n\n"
           "| case | samples | q rows each | kv each | causal | engine attention kernel us (in-kernel timer) | TFLOP/s | q/k-norm + RoPE + append kernel us | whole block call us (host metadata + 2 D2D copies + both kernels) | flash-attn us (attention only) | flash-attn TFLOP/s | engine kernel / flash-attn |\n|---|---:|---:|---:|---|---:|---:|---:|---:|---:|---:|---:|\n")
for (name, B, q, past, causal) in [("image prefill (8 x 1026, full mask)", 8, 1026, 0, False), ("image prefill, 1 image", 1, 1026, 0, False),
                                   ("text prefill on image ctx (8 x 34 on 1026, causal)", 8, 34, 1026, True), ("flow step (12 x 258 on 34, full)", 12, 258, 34, False),
                                   ("long causal chunk (1 x 4096)", 1, 4096, 0, True)]:
    seqs = [eng.seq_new() for _ in range(B)]
    if past:
        for s_ in seqs:
            eng.attention_block(0, [s_], [past], list(range(past)), qkv=torch.randn(past, QN).bfloat16(), update_kv=True)
    qkv = torch.randn(B * q, QN, device="cuda").bfloat16()
    pos = [past + (j if causal else 0) for _ in range(B) for j in range(q)]
    blk = timeit(lambda: eng.attention_block(0, seqs, [q] * B, pos, qkv=qkv, is_causal=causal, update_kv=False))
    tr = traced(eng, lambda: eng.attention_block(0, seqs, [q] * B, pos, qkv=qkv, is_causal=causal, update_kv=False))
    k_attn = sum(v for k, v in tr.items() if k.startswith("attn"))
    k_rope = sum(v for k, v in tr.items() if k.startswith("rope"))
    kvn = past + q
    fl = 4.0 * B * q * (kvn if not causal else (past + (q + 1) / 2)) * H * DH
    fa_us = float("nan")
    if flash_attn_varlen_func is not None:
        qq = torch.randn(B * q, H, DH, device="cuda").bfloat16()
        kk = torch.randn(B * kvn, HKV, DH, device="cuda").bfloat16()
        vv = torch.randn(B * kvn, HKV, DH, device="cuda").bfloat16()
        cq = torch.arange(0, (B + 1) * q, q, device="cuda", dtype=torch.int32)
        ck = torch.arange(0, (B + 1) * kvn, kvn, device="cuda", dtype=torch.int32)
        fa_us = timeit(lambda: flash_attn_varlen_func(qq, kk, vv, cq, ck, q, kvn, causal=causal))
    out.append(f"| {name} | {B} | {q} | {kvn} | {causal} | {k_attn:.1f} ({'+'.join(k for k in tr if k.startswith('attn'))}) | {fl / k_attn / 1e6:.0f} | {k_rope:.1f} | {blk:.1f} | {fa_us:.1f} | {fl / fa_us / 1e6:.0f} | {k_attn / fa_us:.2f} |\n")
    for s_ in seqs:
        eng.seq_free(s_)
out.append("\nThe engine kernel column is the attention kernel alone, first CTA start to last CTA end on the in-kernel `%globaltimer` trace (`umv_trace_*`); the "
           "whole-block column is one `umv_op_attention_block` call (host metadata build + upload, two device copies of q|k|v and the output, per-head q/k RMSNorm, "
           "RoPE, K/V append into the paged pool, attention over the pool) and is host-bound at these sizes -- inside a model forward the metadata goes up once "
           "per forward, not per layer; flash-attn's column is the attention kernel alone on contiguous q / k / v -- the reference additionally runs the "
           "norm, RoPE, the dtype casts and the `torch.zeros` + two index_put KV merges (qwen2_navit.py:544-600) as separate kernels.\n")
eng.close()
del eng
torch.cuda.empty_cache()

# ------------------------------------------------------------------------------------------------ VAE
try:
    import refharness as rh
    if rh.ref_path() is None:
        raise RuntimeError("no reference copy (tools/install_ref.py)")
    from unimedvl_b200 import synth
    from unimedvl_b200.autoencoder import AutoEncoder
    d = ucfg.tiny()
    with synth.on_device("cuda"):
        sd = synth.bagel_state_dict(d, 0)
        vsd = synth.vae_state_dict(d.vae, 0)
    ref, rvae = rh.build_reference(d, sd, vsd, "cuda")
    e2 = Engine(d, max_tokens=512, max_seqs=2, kv_pages=8, enable_vae=True)
    e2.load_state_dict(sd)
    vae = AutoEncoder(e2)
    vae.load_state_dict(vsd)
    e2.finalize()
    out.append("\n## VAE (FLUX autoencoder, bf16): engine (im2col + tcgen05 linears) vs the reference's cuDNN convolutions under CUDA autocast\n\n"
               "| op | shape | engine ms | reference (cuDNN) ms | engine / reference |\n|---|---|---:|---:|---:|\n")
    for n in (1, 4):
        z = torch.randn(n, 16, 32, 32, device="cuda").bfloat16()
        x = (torch.rand(n, 3, 256, 256, device="cuda") * 2 - 1).bfloat16()

        def ref_dec():
            with torch.no_grad(), rh.autocast("cuda"):
                return rvae.decode(z)

        def ref_enc():
            with torch.no_grad(), rh.autocast("cuda"):
                return rvae.encoder(x)
        a, b = timeit(lambda: vae.decode(z), 5) / 1e3, timeit(ref_dec, 5) / 1e3
        out.append(f"| decode | {n} x 16x32x32 -> {n} x 3x256x256 | {a:.2f} | {b:.2f} | {a / b:.2f} |\n")
        a, b = timeit(lambda: vae.encode_moments(x), 5) / 1e3, timeit(ref_enc, 5) / 1e3
        out.append(f"| encode (moments) | {n} x 3x256x256 -> {n} x 32x32x32 | {a:.2f} | {b:.2f} | {a / b:.2f} |\n")
    lat = torch.randn(4, 256, 64, device="cuda")
    a = timeit(lambda: e2.decode_image_u8(lat, 16, 16), 5) / 1e3
    out.append(f"| decode_image_u8 (un-patchify + decode + uint8) | 4 x 256 tokens -> 4 x 256x256x3 u8 | {a:.2f} | - | - |\n")
except Exception as ex:
    out.append(f"\n## VAE vs cuDNN: skipped ({type(ex).__name__}: {ex})\n")

text = "# Engine kernels vs the library kernels the reference calls (same B200, same shapes)\n\n" + "".join(out)
print(text)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
