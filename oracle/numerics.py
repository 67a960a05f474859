"""Elementary numerics of the oracle (test infrastructure; see oracle/__init__.py).

All helpers take/return torch CPU tensors.  "bf16" below always means
round-to-nearest-even to bfloat16, the rounding torch applies on ``.to(bfloat16)``.
"""
from __future__ import annotations

import enum
import math

import torch
import torch.nn.functional as F

BF16 = torch.bfloat16
F32 = torch.float32


class Semantics(enum.Enum):
    """Which autocast behaviour of the reference is restated.

    cuda: LayerNorm / GroupNorm / torch.norm / softmax run in fp32 and return
          fp32 (AutocastCUDA fp32-policy ops), SURVEY.md section 8a.
    cpu : those ops stay in bf16 (no AutocastCPU kernel) -- what the committed
          golden fixtures were generated with.
    """
    cuda = "cuda"
    cpu = "cpu"


def bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(BF16)


def linear(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor | None = None,
           exact: bool = True) -> torch.Tensor:
    """Autocast ``F.linear``: operands cast to bf16, fp32 accumulate (+bias in
    fp32), ONE rounding to bf16 (R3).  modeling_qwen2.py:283-286 et al.

    exact=True does the contraction in fp32 tensors (order-independent to fp32
    rounding noise); exact=False uses the native bf16 CPU GEMM (same contract,
    faster -- used by the timed CPU baseline).
    """
    xb = bf16(x)
    if exact:
        y = F.linear(xb.float(), w.float(), None if b is None else b.float())
        return bf16(y)
    return F.linear(xb, bf16(w), None if b is None else bf16(b))


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """Qwen2RMSNorm.forward, modeling_qwen2.py:89-94.

    fp32 normalise -> cast back to the INPUT dtype (bf16 rounds here, R2; fp32
    input as in gen-mode q/k stays fp32) -> multiply by the bf16 weight (type
    promotion: bf16*bf16 -> bf16 (one more rounding), bf16*fp32 -> fp32).
    """
    in_dtype = x.dtype
    h = x.to(F32)
    var = h.pow(2).mean(-1, keepdim=True)
    h = h * torch.rsqrt(var + eps)
    return w * h.to(in_dtype)


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float,
              sem: Semantics) -> torch.Tensor:
    """nn.LayerNorm under autocast (siglip_navit.py:283,296,370).

    cuda: fp32-policy op -> fp32 in, fp32 out.  cpu: runs on the bf16 tensor.
    """
    if sem is Semantics.cuda:
        return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), eps)
    return F.layer_norm(x, (x.shape[-1],), w.to(x.dtype), b.to(x.dtype), eps)


def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    """ACT2FN['gelu_pytorch_tanh'] on a bf16 tensor: fp32 math, one rounding."""
    return F.gelu(x, approximate="tanh")


def silu(x: torch.Tensor) -> torch.Tensor:
    """ACT2FN['silu'] / nn.SiLU on the tensor's own dtype (fp32 math inside)."""
    return F.silu(x)


def rope_cos_sin(position_ids: torch.Tensor, inv_freq: torch.Tensor,
                 out_dtype: torch.dtype) -> tuple[torch.Tensor, torch.Tensor]:
    """Qwen2RotaryEmbedding.forward, modeling_qwen2.py:164-184.

    angles = inv_freq (fp32, never rounded to bf16 -- SURVEY S4) * position in
    fp32; emb = cat(freqs, freqs); cos/sin in fp32; cast to the activation dtype
    (bf16, R5).  attention_scaling == 1.0 for the default rope type.
    """
    freqs = position_ids.to(F32)[:, None] * inv_freq.to(F32)[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(out_dtype), emb.sin().to(out_dtype)


def default_inv_freq(head_dim: int, theta: float) -> torch.Tensor:
    """ROPE_INIT_FUNCTIONS['default'] (transformers 4.49, used at
    modeling_qwen2.py:139-141): 1 / theta^(arange(0,dh,2)/dh), fp32."""
    return 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    """modeling_qwen2.py:188-192."""
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def apply_rope(q: torch.Tensor, k: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor):
    """apply_rotary_pos_emb(unsqueeze_dim=1), modeling_qwen2.py:196-220.

    q,k: [T, heads, dh]; cos/sin: [T, dh] (bf16).  With bf16 q/k every product
    and the sum round to bf16 (three roundings, R5); with fp32 q/k (gen mode)
    torch promotes to fp32 and nothing rounds here.
    """
    c, s = cos[:, None, :], sin[:, None, :]
    return (q * c) + (rotate_half(q) * s), (k * c) + (rotate_half(k) * s)


def attention_varlen(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
                     q_lens: list[int], k_lens: list[int], causal: bool,
                     scale: float | None = None, p_bf16: bool = True) -> torch.Tensor:
    """flash_attn_varlen_func semantics (call sites qwen2_navit.py:605-614,
    siglip_navit.py:232-241; flash-attn 2.x, not vendored in the reference):

    per sample, GQA (q head h reads kv head h // group), softmax(scale * QK^T)
    in fp32, bottom-right aligned causal mask when q_len != k_len, the
    probabilities are rounded to bf16 before the PV product (FA2 feeds P to the
    tensor cores in the input dtype), fp32 accumulation, fp32 row-sum of the
    un-rounded probabilities, one rounding of the output to bf16.

    p_bf16=False keeps the probabilities in fp32 -- what the golden fixtures contain,
    because on CPU the reference could only run with flash-attn replaced by an fp32
    SDPA stand-in (tests/golden/make_golden.py shim S3).
    """
    T, H, D = q.shape
    Hkv = k.shape[1]
    g = H // Hkv
    scale = (1.0 / math.sqrt(D)) if scale is None else scale
    out = torch.empty((T, H, D), dtype=BF16)
    qo = ko = 0
    for lq, lk in zip(q_lens, k_lens):
        qi = q[qo:qo + lq].float().transpose(0, 1)                      # [H, lq, D]
        ki = k[ko:ko + lk].float().transpose(0, 1).repeat_interleave(g, 0)
        vi = v[ko:ko + lk].float().transpose(0, 1).repeat_interleave(g, 0)
        s = torch.matmul(qi, ki.transpose(1, 2)) * scale                # [H, lq, lk]
        if causal:
            keep = torch.ones(lq, lk, dtype=torch.bool).tril(lk - lq)
            s = s.masked_fill(~keep, float("-inf"))
        m = s.max(-1, keepdim=True).values
        p = torch.exp(s - m)
        l = p.sum(-1, keepdim=True)
        o = torch.matmul(p.to(BF16).float() if p_bf16 else p, vi) / l
        out[qo:qo + lq] = o.transpose(0, 1).to(BF16)
        qo += lq
        ko += lk
    return out


def sincos_2d_table(embed_dim: int, grid: int) -> torch.Tensor:
    """get_2d_sincos_pos_embed, modeling_utils.py:23-65 (frozen table of
    PositionEmbedding :126-143).  float64 angles, [sin|cos] per axis, first half
    of the channels from the w coordinate ("w goes first" meshgrid quirk :26-27
    makes grid[0] the column index), second half from the row index.  fp32 out.
    """
    def axis(dim: int, pos: torch.Tensor) -> torch.Tensor:
        omega = torch.arange(dim // 2, dtype=torch.float64) / (dim / 2.0)
        omega = 1.0 / (10000.0 ** omega)
        ang = pos.reshape(-1).to(torch.float64)[:, None] * omega[None, :]
        return torch.cat((ang.sin(), ang.cos()), dim=1)

    rows, cols = torch.meshgrid(torch.arange(grid, dtype=torch.float32),
                                torch.arange(grid, dtype=torch.float32), indexing="ij")
    emb = torch.cat((axis(embed_dim // 2, cols), axis(embed_dim // 2, rows)), dim=1)
    return emb.to(F32)


def timestep_frequencies(t: torch.Tensor, dim: int = 256, max_period: float = 10000.0) -> torch.Tensor:
    """TimestepEmbedder.timestep_embedding, modeling_utils.py:86-104 (fp32)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=F32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat((torch.cos(args), torch.sin(args)), dim=-1)
