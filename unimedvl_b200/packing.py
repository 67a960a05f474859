"""Host-side packing: builds the ``generation_input`` dicts of the reference's ``Bagel.prepare_*``
methods (same keys, dtypes and values) plus the image transform / patchify that feed them.

Reference: bagel.py:377-409 (prepare_prompts), :460-520 (prepare_vit_images), :617-694
(prepare_vae_images), :809-865 (prepare_vae_latent), :867-898 (prepare_vae_latent_cfg), :1213-1233
(prepare_start_tokens); data/transforms.py:15-115 (ImageTransform); data/data_utils.py:43-58
(patchify, get_flattened_position_ids_extrapolate).  Packed layout per sample: [its cached keys |
its new query tokens], samples concatenated in batch order.  Pure host code (numpy/torch CPU);
verified bit-for-bit against fixtures produced by the reference (tests/test_packing.py).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
from PIL import Image


# ----------------------------------------------------------------------------- image transform
class ResizeRule:
    """MaxLongEdgeMinShortEdgeResize (data/transforms.py:15-87): longest side <= max_size, shortest
    side >= min_size, both sides multiples of stride, area <= max_pixels; bicubic, antialiased."""

    def __init__(self, max_size: int, min_size: int, stride: int, max_pixels: int):
        self.max_size, self.min_size, self.stride, self.max_pixels = max_size, min_size, stride, max_pixels

    def _snap(self, v: float) -> int:
        return max(self.stride, int(round(v / self.stride) * self.stride))

    def _scaled(self, w: int, h: int, s: float) -> Tuple[int, int]:
        return self._snap(round(w * s)), self._snap(round(h * s))

    def target_size(self, width: int, height: int, img_num: int = 1) -> Tuple[int, int]:
        s = min(self.max_size / max(width, height), 1.0)
        s = max(s, self.min_size / min(width, height))
        w, h = self._scaled(width, height, s)
        if w * h > self.max_pixels / img_num:
            w, h = self._scaled(w, h, self.max_pixels / img_num / (w * h))
        if max(w, h) > self.max_size:
            w, h = self._scaled(w, h, self.max_size / max(w, h))
        return w, h

    def __call__(self, img: Image.Image, img_num: int = 1) -> Image.Image:
        w, h = self.target_size(*img.size, img_num=img_num)
        return img.resize((w, h), Image.BICUBIC)       # == torchvision F.resize on a PIL image


class ImageTransform:
    """data/transforms.py:90-115: resize -> ToTensor (CHW float / 255) -> Normalize(0.5, 0.5)."""

    def __init__(self, max_image_size, min_image_size, image_stride, max_pixels=14 * 14 * 9 * 1024,
                 image_mean=(0.5, 0.5, 0.5), image_std=(0.5, 0.5, 0.5)):
        self.stride = image_stride
        self.resize_transform = ResizeRule(max_image_size, min_image_size, image_stride, max_pixels)
        self.mean = torch.tensor(image_mean, dtype=torch.float32).view(3, 1, 1)
        self.std = torch.tensor(image_std, dtype=torch.float32).view(3, 1, 1)

    def __call__(self, img: Image.Image, img_num: int = 1) -> torch.Tensor:
        img = self.resize_transform(img, img_num=img_num)
        a = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy())
        if a.dim() == 2:
            a = a[:, :, None]
        t = a.permute(2, 0, 1).contiguous().to(torch.float32).div(255)
        return t.sub_(self.mean).div_(self.std)


def pil_img2rgb(image: Image.Image) -> Image.Image:
    """data/data_utils.py:116-137: RGBA / palette transparency composited on white, else convert('RGB')."""
    w, h = image.size
    if w * h > 20_000_000:
        raise ValueError(f"Image too large: {w * h} pixels")
    if image.mode == "RGBA" or image.info.get("transparency", None) is not None:
        image = image.convert("RGBA")
        white = Image.new(mode="RGB", size=image.size, color=(255, 255, 255))
        white.paste(image, mask=image.split()[3])
        return white
    return image.convert("RGB")


def patchify(image: torch.Tensor, p: int) -> torch.Tensor:
    """[C,H,W] -> [H/p * W/p, p*p*C] with element order (p_row, p_col, channel) (data_utils.py:43-50)."""
    c, h, w = image.shape
    assert h % p == 0 and w % p == 0
    return image.reshape(c, h // p, p, w // p, p).permute(1, 3, 2, 4, 0).reshape(-1, p * p * c)


def flattened_position_ids(img_h: int, img_w: int, patch: int, max_per_side: int) -> torch.Tensor:
    """row * max_per_side + col over the patch grid (data_utils.py:53-58, the 'extrapolate' variant)."""
    rows = torch.arange(img_h // patch)
    cols = torch.arange(img_w // patch)
    return (rows[:, None] * max_per_side + cols[None, :]).reshape(-1)


# ----------------------------------------------------------------------------- index layout
def _layout(curr_kvlens: Sequence[int], new_lens: Sequence[int]):
    """Offsets of the packed [kv_i | new_i] layout: returns (kv_indexes, list of new-token start offsets)."""
    kv = np.asarray(curr_kvlens, dtype=np.int64)
    nl = np.asarray(new_lens, dtype=np.int64)
    block_start = np.concatenate([[0], np.cumsum(kv + nl)[:-1]]) if len(kv) else np.zeros(0, np.int64)
    kv_idx = [np.arange(s, s + k) for s, k in zip(block_start, kv)]
    kv_indexes = np.concatenate(kv_idx) if kv_idx else np.zeros(0, np.int64)
    return kv_indexes.astype(np.int64), (block_start + kv).tolist()


def _i64(a) -> torch.Tensor:
    return torch.as_tensor(np.asarray(a, dtype=np.int64))


def _i32(a) -> torch.Tensor:
    return torch.as_tensor(np.asarray(a, dtype=np.int32))


def prepare_prompts(curr_kvlens, curr_rope, prompts, tokenizer, new_token_ids):
    ids = [[new_token_ids["bos_token_id"]] + list(tokenizer.encode(p)) + [new_token_ids["eos_token_id"]] for p in prompts]
    lens = [len(t) for t in ids]
    kv_indexes, starts = _layout(curr_kvlens, lens)
    g = {
        "text_token_lens": _i32(lens),
        "packed_text_ids": _i64([t for seq in ids for t in seq]),
        "packed_text_position_ids": _i64(np.concatenate([np.arange(r, r + n) for r, n in zip(curr_rope, lens)]) if lens else []),
        "packed_text_indexes": _i64(np.concatenate([np.arange(s, s + n) for s, n in zip(starts, lens)]) if lens else []),
        "packed_key_value_indexes": _i64(kv_indexes),
        "key_values_lens": _i32(list(curr_kvlens)),
    }
    newlens = [int(k) + n for k, n in zip(curr_kvlens, lens)]
    new_rope = [int(r) + n for r, n in zip(curr_rope, lens)]
    return g, newlens, new_rope


def _image_block_layout(curr_kvlens, curr_rope, n_img_tokens, new_token_ids):
    """Shared by the ViT / VAE / latent packers: every image contributes
    [start_of_image, n image tokens, end_of_image], all at ONE rope position (bagel.py:501-504)."""
    seqlens = [n + 2 for n in n_img_tokens]
    kv_indexes, starts = _layout(curr_kvlens, seqlens)
    qstart = np.concatenate([[0], np.cumsum(seqlens)[:-1]]).astype(np.int64) if seqlens else np.zeros(0, np.int64)
    text_idx, img_idx, packed_idx, pos = [], [], [], []
    for s, q, n, r in zip(starts, qstart, n_img_tokens, curr_rope):
        text_idx += [q, q + n + 1]
        img_idx.append(np.arange(q + 1, q + 1 + n))
        packed_idx.append(np.arange(s, s + n + 2))
        pos.append(np.full(n + 2, int(r), dtype=np.int64))
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.int64)
    text_ids = []
    if new_token_ids is not None:
        text_ids = [new_token_ids["start_of_image"], new_token_ids["end_of_image"]] * len(seqlens)
    return dict(seqlens=seqlens, kv_indexes=kv_indexes, text_ids=text_ids, text_idx=text_idx, img_idx=cat(img_idx),
                packed_idx=cat(packed_idx), pos=cat(pos))


def image_prompt_layout(n_img_tokens, prompt_ids, new_token_ids, curr_kvlens=None, curr_rope=None):
    """Packed rows of VQA requests prefilled in ONE forward (Engine.forward_cache_update_vit(prompt_lens=...)): per sample
    [start_of_image, n image tokens, end_of_image] at the sample's rope position r -- the block prepare_vit_images lays out
    (bagel.py:460-520) -- followed by [bos, prompt ids, eos] at rope positions r + 1, r + 2, ... -- the rows prepare_prompts lays out on top
    of it (bagel.py:377-409).  curr_kvlens / curr_rope: cache length and rope position each sample starts from (a shared prefix; default
    0).  Returns the row lists the engine call takes and the (kv_lens, rope) state after both prefills."""
    B = len(n_img_tokens)
    curr_kvlens = [0] * B if curr_kvlens is None else [int(k) for k in curr_kvlens]
    curr_rope = [0] * B if curr_rope is None else [int(r) for r in curr_rope]
    seq_lens, prompt_lens, text_ids, text_rows, vit_rows, pos = [], [], [], [], [], []
    row = 0
    for n, ids, r in zip(n_img_tokens, prompt_ids, curr_rope):
        n = int(n)
        p = [new_token_ids["bos_token_id"]] + [int(t) for t in ids] + [new_token_ids["eos_token_id"]]
        text_ids += [new_token_ids["start_of_image"], new_token_ids["end_of_image"]] + p
        text_rows += [row, row + n + 1] + list(range(row + n + 2, row + n + 2 + len(p)))
        vit_rows += list(range(row + 1, row + 1 + n))
        pos += [r] * (n + 2) + list(range(r + 1, r + 1 + len(p)))
        seq_lens.append(n + 2 + len(p))
        prompt_lens.append(len(p))
        row += n + 2 + len(p)
    return dict(seq_lens=seq_lens, prompt_lens=prompt_lens, text_ids=text_ids, text_rows=text_rows, vit_rows=vit_rows, positions=pos,
                kv_lens=[k + s_ for k, s_ in zip(curr_kvlens, seq_lens)], rope=[r + 1 + p for r, p in zip(curr_rope, prompt_lens)])


def prepare_vit_images(curr_kvlens, curr_rope, images, transforms, new_token_ids, vit_patch_size=14,
                       vit_max_num_patch_per_side=70):
    tensors = [transforms(im) for im in images]
    tokens = [patchify(t, vit_patch_size) for t in tensors]
    pos_ids = [flattened_position_ids(t.size(1), t.size(2), vit_patch_size, vit_max_num_patch_per_side) for t in tensors]
    n = [int(t.shape[0]) for t in tokens]
    L = _image_block_layout(curr_kvlens, curr_rope, n, new_token_ids)
    g = {
        "packed_text_ids": _i64(L["text_ids"]),
        "packed_text_indexes": _i64(L["text_idx"]),
        "vit_token_seqlens": _i32(n),
        "packed_vit_tokens": torch.cat(tokens, dim=0),
        "packed_vit_position_ids": torch.cat(pos_ids, dim=0),
        "packed_vit_token_indexes": _i64(L["img_idx"]),
        "packed_position_ids": _i64(L["pos"]),
        "packed_seqlens": _i32(L["seqlens"]),
        "packed_indexes": _i64(L["packed_idx"]),
        "packed_key_value_indexes": _i64(L["kv_indexes"]),
        "key_values_lens": _i32(list(curr_kvlens)),
    }
    newlens = [int(k) + s for k, s in zip(curr_kvlens, L["seqlens"])]
    new_rope = [int(r) + 1 for r in curr_rope]
    return g, newlens, new_rope


def prepare_vae_images(curr_kvlens, curr_rope, images, transforms, new_token_ids, latent_downsample=16,
                       max_latent_size=64, timestep=0):
    tensors = [transforms(im) for im in images]
    shapes = [(t.shape[1] // latent_downsample, t.shape[2] // latent_downsample) for t in tensors]
    pos_ids = [flattened_position_ids(t.size(1), t.size(2), latent_downsample, max_latent_size) for t in tensors]
    n = [h * w for h, w in shapes]
    L = _image_block_layout(curr_kvlens, curr_rope, n, new_token_ids)
    max_c = max(t.shape[0] for t in tensors)
    max_h = max(t.shape[1] for t in tensors)
    max_w = max(t.shape[2] for t in tensors)
    padded = torch.zeros((len(tensors), max_c, max_h, max_w))
    for i, t in enumerate(tensors):
        padded[i, :, :t.shape[1], :t.shape[2]] = t
    g = {
        "padded_images": padded,
        "patchified_vae_latent_shapes": shapes,
        "packed_vae_position_ids": torch.cat(pos_ids, dim=0),
        "packed_timesteps": torch.tensor([timestep]),
        "packed_vae_token_indexes": _i64(L["img_idx"]),
        "packed_text_ids": _i64(L["text_ids"]),
        "packed_text_indexes": _i64(L["text_idx"]),
        "packed_position_ids": _i64(L["pos"]),
        "packed_seqlens": _i32(L["seqlens"]),
        "packed_indexes": _i64(L["packed_idx"]),
        "packed_key_value_indexes": _i64(L["kv_indexes"]),
        "key_values_lens": _i32(list(curr_kvlens)),
    }
    newlens = [int(k) + s for k, s in zip(curr_kvlens, L["seqlens"])]
    new_rope = [int(r) + 1 for r in curr_rope]
    return g, newlens, new_rope


def prepare_vae_latent(curr_kvlens, curr_rope, image_sizes, new_token_ids, latent_downsample=16, max_latent_size=64,
                       patch_latent_dim=64):
    """Initial noise is drawn image by image from torch's global CPU generator (bagel.py:835-837), so
    torch.manual_seed(s) before the call reproduces the reference's noise."""
    hw = [(H // latent_downsample, W // latent_downsample) for (H, W) in image_sizes]
    n = [h * w for h, w in hw]
    noises = [torch.randn(k, patch_latent_dim) for k in n]
    pos_ids = [flattened_position_ids(H, W, latent_downsample, max_latent_size) for (H, W) in image_sizes]
    L = _image_block_layout(curr_kvlens, curr_rope, n, new_token_ids)
    return {
        "packed_text_ids": _i64(L["text_ids"]),
        "packed_text_indexes": _i64(L["text_idx"]),
        "packed_init_noises": torch.cat(noises, dim=0),
        "packed_vae_position_ids": torch.cat(pos_ids, dim=0),
        "packed_vae_token_indexes": _i64(L["img_idx"]),
        "packed_seqlens": _i32(L["seqlens"]),
        "packed_position_ids": _i64(L["pos"]),
        "key_values_lens": _i32(list(curr_kvlens)),
        "packed_indexes": _i64(L["packed_idx"]),
        "packed_key_value_indexes": _i64(L["kv_indexes"]),
    }


def prepare_vae_latent_cfg(curr_kvlens, curr_rope, image_sizes, latent_downsample=16):
    n = [(H // latent_downsample) * (W // latent_downsample) for (H, W) in image_sizes]
    L = _image_block_layout(curr_kvlens, curr_rope, n, None)
    return {
        "cfg_packed_position_ids": _i64(L["pos"]),
        "cfg_key_values_lens": _i32(list(curr_kvlens)),
        "cfg_packed_query_indexes": _i64(L["packed_idx"]),
        "cfg_packed_key_value_indexes": _i64(L["kv_indexes"]),
    }


def prepare_start_tokens(curr_kvlens, curr_rope, new_token_ids, device=None):
    kv = np.asarray(curr_kvlens, dtype=np.int64)
    kv_indexes = np.arange(int(kv.sum()), dtype=np.int64)      # decode starts from a gap-free packed cache
    return {
        "packed_start_tokens": _i64([new_token_ids["bos_token_id"]] * len(kv)).to(device),
        "packed_query_position_ids": _i64(list(curr_rope)).to(device),
        "key_values_lens": _i32(list(curr_kvlens)).to(device),
        "packed_key_value_indexes": _i64(kv_indexes).to(device),
    }
