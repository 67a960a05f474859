"""Oracle: FLUX-style KL autoencoder (TEST INFRASTRUCTURE).

Follows autoencoder.py:38-65 (AttnBlock), :68-95 (ResnetBlock), :98-119
(Down/Upsample), :122-187 (Encoder), :190-257 (Decoder), :260-272
(DiagonalGaussian), :300-307 (encode/decode), hard-coded params :338-349.
Weights are bf16 (interactive_image_generator.py:193,223) and every conv runs
under autocast (bf16 operands, fp32 accumulate, one rounding).  State-dict keys
are ``encoder.*`` / ``decoder.*`` exactly as ``AutoEncoder.state_dict()``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

from .numerics import BF16, Semantics


@dataclass
class VAEDims:
    ch: int = 128
    ch_mult: tuple = (1, 2, 4, 4)
    num_res_blocks: int = 2
    z_channels: int = 16
    in_channels: int = 3
    out_ch: int = 3
    scale_factor: float = 0.3611
    shift_factor: float = 0.1159
    groups: int = 32
    eps: float = 1e-6


def conv(sd, name, x, stride=1, padding=1):
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    y = F.conv2d(x.to(BF16).float(), w.float(), b.float(), stride=stride, padding=padding)
    return y.to(BF16)


def gn_swish(sd, name, x, dims: VAEDims, sem: Semantics, act=True):
    """GroupNorm(32, eps 1e-6) then swish (autoencoder.py:84-85).  cuda: GroupNorm is
    an fp32-policy autocast op -> fp32 out and swish runs in fp32; cpu: both bf16."""
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    if sem is Semantics.cuda:
        h = F.group_norm(x.float(), dims.groups, w.float(), b.float(), dims.eps)
    else:
        h = F.group_norm(x, dims.groups, w.to(x.dtype), b.to(x.dtype), dims.eps)
    return h * torch.sigmoid(h) if act else h


def resnet_block(sd, p, x, cin, cout, dims, sem):
    h = conv(sd, p + "conv1", gn_swish(sd, p + "norm1", x, dims, sem))
    h = conv(sd, p + "conv2", gn_swish(sd, p + "norm2", h, dims, sem))
    if cin != cout:
        x = conv(sd, p + "nin_shortcut", x, padding=0)
    return x + h


def attn_block(sd, p, x, dims, sem):
    """AttnBlock.forward (:50-65): single head, d = channels, SDPA scale 1/sqrt(c)."""
    h = gn_swish(sd, p + "norm", x, dims, sem, act=False)
    q = conv(sd, p + "q", h, padding=0)
    k = conv(sd, p + "k", h, padding=0)
    v = conv(sd, p + "v", h, padding=0)
    b, c, hh, ww = q.shape
    f = lambda t: t.reshape(b, c, hh * ww).transpose(1, 2).float()          # b (h w) c
    s = torch.matmul(f(q), f(k).transpose(1, 2)) * (c ** -0.5)
    pm = torch.softmax(s, dim=-1)
    o = torch.matmul(pm.to(BF16).float(), f(v)).to(BF16)
    o = o.transpose(1, 2).reshape(b, c, hh, ww)
    return x + conv(sd, p + "proj_out", o, padding=0)


def decoder(sd, z, dims: VAEDims, sem: Semantics = Semantics.cuda, taps=None):
    """Decoder.forward (:240-257)."""
    P = "decoder."
    nres = len(dims.ch_mult)
    block_in = dims.ch * dims.ch_mult[-1]
    h = conv(sd, P + "conv_in", z)
    h = resnet_block(sd, P + "mid.block_1.", h, block_in, block_in, dims, sem)
    h = attn_block(sd, P + "mid.attn_1.", h, dims, sem)
    h = resnet_block(sd, P + "mid.block_2.", h, block_in, block_in, dims, sem)
    if taps is not None:
        taps["mid"] = h.clone()
    for lvl in reversed(range(nres)):
        block_out = dims.ch * dims.ch_mult[lvl]
        for bi in range(dims.num_res_blocks + 1):
            h = resnet_block(sd, f"{P}up.{lvl}.block.{bi}.", h, block_in, block_out, dims, sem)
            block_in = block_out
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = conv(sd, f"{P}up.{lvl}.upsample.conv", h)
        if taps is not None:
            taps[f"up{lvl}"] = h.clone()
    h = gn_swish(sd, P + "norm_out", h, dims, sem)
    return conv(sd, P + "conv_out", h)


def encoder(sd, x, dims: VAEDims, sem: Semantics = Semantics.cuda):
    """Encoder.forward (:169-187).  Returns the 2*z moments tensor."""
    P = "encoder."
    nres = len(dims.ch_mult)
    in_mult = (1,) + tuple(dims.ch_mult)
    h = conv(sd, P + "conv_in", x)
    block_in = dims.ch
    for lvl in range(nres):
        block_in = dims.ch * in_mult[lvl]
        block_out = dims.ch * dims.ch_mult[lvl]
        for bi in range(dims.num_res_blocks):
            h = resnet_block(sd, f"{P}down.{lvl}.block.{bi}.", h, block_in, block_out, dims, sem)
            block_in = block_out
        if lvl != nres - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)               # :105-106 asymmetric pad
            h = conv(sd, f"{P}down.{lvl}.downsample.conv", h, stride=2, padding=0)
    h = resnet_block(sd, P + "mid.block_1.", h, block_in, block_in, dims, sem)
    h = attn_block(sd, P + "mid.attn_1.", h, dims, sem)
    h = resnet_block(sd, P + "mid.block_2.", h, block_in, block_in, dims, sem)
    h = gn_swish(sd, P + "norm_out", h, dims, sem)
    return conv(sd, P + "conv_out", h)


def decode(sd, z, dims: VAEDims = VAEDims(), sem: Semantics = Semantics.cuda, taps=None):
    """AutoEncoder.decode (:305-307)."""
    z = z / dims.scale_factor + dims.shift_factor
    return decoder(sd, z, dims, sem, taps)


def encode(sd, x, dims: VAEDims = VAEDims(), sem: Semantics = Semantics.cuda, noise=None):
    """AutoEncoder.encode (:300-303).  ``noise`` replaces DiagonalGaussian's device
    ``randn_like`` (:270); None -> the mean (sample=False)."""
    m = encoder(sd, x, dims, sem)
    mean, logvar = torch.chunk(m, 2, dim=1)
    z = mean if noise is None else mean + torch.exp(0.5 * logvar) * noise.to(mean.dtype)
    return dims.scale_factor * (z - dims.shift_factor)


def image_to_uint8(img: torch.Tensor) -> torch.Tensor:
    """inferencer.py:253-254: (x*0.5+0.5).clamp(0,1)[0] HWC * 255 -> uint8 (truncation)."""
    x = (img * 0.5 + 0.5).clamp(0, 1)[0].permute(1, 2, 0) * 255
    return x.to(torch.uint8)
