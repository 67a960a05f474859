"""Python handle on the CUDA engine (libumv.so).  torch is used for device memory and streams only;
every compute call goes through the C ABI of include/umv.h."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import torch

from . import _lib
from .config import BagelDims


def _ptr(t: torch.Tensor | None):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(stream: torch.cuda.Stream | None = None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


class Engine:
    """One engine per GPU per process (matches the reference's single-device usage,
    codes/interactive_vqa_inferencer.py:19-20).  Owns weights, the KV page pool and workspaces."""

    def __init__(self, dims: BagelDims, max_tokens: int = 2048, max_seqs: int = 8, kv_pages: int = 256,
                 enable_vit: bool = True, enable_gen: bool = True, enable_vae: bool = False, device: int | None = None):
        if not torch.cuda.is_available():
            raise _lib.UmvError("unimedvl_b200 needs a CUDA device: the engine has no CPU fallback")
        self.lib = _lib.load()
        if device is not None:
            torch.cuda.set_device(device)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.dims = dims
        # the engine replays captured decode steps: use a non-default stream for all work
        self.stream = torch.cuda.Stream(device=self.device)
        l, v = dims.llm, dims.vit
        d = _lib.Dims(
            hidden=l.hidden, heads=l.heads, kv_heads=l.kv_heads, inter=l.inter, layers=l.layers, vocab=l.vocab,
            rope_theta=l.rope_theta, rms_eps=l.eps,
            vit_hidden=v.hidden, vit_heads=v.heads, vit_inter=v.inter, vit_layers=v.layers, vit_patch_dim=v.patch_dim,
            vit_positions=v.num_positions, vit_eps=v.eps, vit_pos_table=dims.vit_max_num_patch_per_side ** 2,
            latent_dim=dims.patch_latent_dim, latent_pos_table=dims.max_latent_size ** 2,
            max_tokens=max_tokens, max_seqs=max_seqs, kv_pages=kv_pages,
            enable_vit=int(enable_vit), enable_gen=int(enable_gen), enable_vae=int(enable_vae))
        self._c_dims = d
        h = C.c_void_p()
        _lib.check(self.lib.umv_create(C.byref(d), C.byref(h)))
        self.h = h
        self.max_tokens, self.max_seqs = max_tokens, max_seqs

    def close(self):
        if getattr(self, "h", None):
            self.lib.umv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd: dict, strict: bool = True) -> None:
        """Reference state_dict keys; tensors may live on CPU or the GPU."""
        for name, t in sd.items():
            t = t.detach()
            if t.dtype not in (torch.bfloat16, torch.float32):
                if not strict:          # e.g. an int64 step counter / position_ids buffer in the shard: nothing the engine holds
                    continue
                raise ValueError(f"{name}: dtype {t.dtype} (bf16 or fp32 expected)")
            if t.dtype == torch.float32 and t.is_cuda:
                t = t.to(torch.bfloat16)
            t = t.contiguous()
            shape = _lib.i64_array(t.shape)
            rc = self.lib.umv_load_tensor(self.h, name.encode(), _ptr(t), _lib.BF16 if t.dtype == torch.bfloat16 else _lib.F32,
                                          t.dim(), shape, int(t.is_cuda))
            if rc == _lib.ERR_INVALID and not strict and "unknown tensor" in self.lib.umv_last_error().decode():
                continue
            _lib.check(rc)
        torch.cuda.synchronize()

    def fill_synthetic(self, seed: int = 0) -> None:
        _lib.check(self.lib.umv_fill_synthetic(self.h, C.c_uint64(seed)))

    def finalize(self) -> None:
        _lib.check(self.lib.umv_finalize(self.h))

    def export_tensor(self, name: str, shape: Sequence[int]) -> torch.Tensor:
        out = torch.empty(tuple(shape), dtype=torch.bfloat16)
        _lib.check(self.lib.umv_export_tensor(self.h, name.encode(), _ptr(out), out.numel() * 2))
        return out

    def weight_bytes(self) -> int:
        n = C.c_int64()
        _lib.check(self.lib.umv_weight_bytes(self.h, C.byref(n)))
        return n.value

    # ---------------------------------------------------------------- sequences
    def seq_new(self) -> int:
        s = C.c_int32()
        _lib.check(self.lib.umv_seq_new(self.h, C.byref(s)))
        return s.value

    def seq_fork(self, src: int) -> int:
        s = C.c_int32()
        _lib.check(self.lib.umv_seq_fork(self.h, src, C.byref(s)))
        return s.value

    def seq_free(self, seq: int) -> None:
        if self.h:
            _lib.check(self.lib.umv_seq_free(self.h, seq))

    def seq_len(self, seq: int) -> int:
        n = C.c_int32()
        _lib.check(self.lib.umv_seq_len(self.h, seq, C.byref(n)))
        return n.value

    def seq_truncate(self, seq: int, length: int) -> None:
        _lib.check(self.lib.umv_seq_truncate(self.h, seq, length))

    def pages_free(self) -> int:
        n = C.c_int32()
        _lib.check(self.lib.umv_pages_free(self.h, C.byref(n)))
        return n.value

    def seq_export(self, seq: int, layer: int):
        """Packed [len, kv_heads, head_dim] K and V of one layer (NaiveCache layout)."""
        n = self.seq_len(seq)
        l = self.dims.llm
        k = torch.empty((n, l.kv_heads, l.head_dim), dtype=torch.bfloat16, device=self.device)
        v = torch.empty_like(k)
        with torch.cuda.stream(self.stream):
            _lib.check(self.lib.umv_seq_export(self.h, seq, layer, _ptr(k), _ptr(v), _stream_ptr(self.stream)))
        self.stream.synchronize()
        return k, v

    # ------------------------------------------------------------------ forward
    def _enter(self):
        """Order the engine stream after whatever the caller queued on the current stream."""
        self.stream.wait_stream(torch.cuda.current_stream())

    def _exit(self):
        torch.cuda.current_stream().wait_stream(self.stream)

    def embed_tokens(self, ids: torch.Tensor) -> torch.Tensor:
        ids = ids.to(self.device, torch.int64).contiguous()
        out = torch.empty((ids.numel(), self.dims.llm.hidden), dtype=torch.bfloat16, device=self.device)
        self._enter()
        _lib.check(self.lib.umv_embed_tokens(self.h, _ptr(ids), ids.numel(), _ptr(out), _stream_ptr(self.stream)))
        self._exit()
        return out

    def patchify_u8(self, images: Sequence[torch.Tensor]):
        """ToTensor + Normalize(0.5, 0.5) + patchify + flattened position ids on the device (data/transforms.py:104-115,
        data/data_utils.py:43-58) for uint8 [H, W, 3] images whose sides are multiples of the ViT patch (resizing stays
        with the caller's ImageTransform).  Host tensors are uploaded as uint8 (a twelfth of the fp32 patch matrix).
        Returns (pixels f32 [n_tokens, patch_dim], pos_ids i64 [n_tokens], seqlens) on the device, bit-identical to the
        host transforms."""
        v = self.dims.vit
        hw, offs, total = [], [], 0
        for im in images:
            assert im.dtype == torch.uint8 and im.dim() == 3 and im.shape[2] == v.channels, "uint8 [H, W, 3] images expected"
            hw += [int(im.shape[0]), int(im.shape[1])]
            offs.append(total)
            total += im.numel()
        flat = torch.empty(total, dtype=torch.uint8, device=self.device)
        for im, o in zip(images, offs):
            flat[o:o + im.numel()].copy_(im.reshape(-1), non_blocking=True)
        lens = [(hw[2 * i] // v.patch) * (hw[2 * i + 1] // v.patch) for i in range(len(images))]
        pixels = torch.empty((sum(lens), v.patch_dim), dtype=torch.float32, device=self.device)
        pos = torch.empty((sum(lens),), dtype=torch.int64, device=self.device)
        _lib.check(self.lib.umv_patchify_u8(_ptr(flat), _lib.i64_array(offs), _lib.i32_array(hw), len(images), v.patch,
                                            self.dims.vit_max_num_patch_per_side, _ptr(pixels), _ptr(pos), _stream_ptr()))
        return pixels, pos, lens

    def resize_u8(self, image: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
        """PIL.Image.resize((out_w, out_h), BICUBIC) -- the resize of MaxLongEdgeMinShortEdgeResize (data/transforms.py:60-87)
        -- on the device, bit-identical to Pillow: uint8 [H, W, 3] (host or device) -> device uint8 [out_h, out_w, 3]."""
        assert image.dtype == torch.uint8 and image.dim() == 3 and image.shape[2] == 3, "uint8 [H, W, 3] image expected"
        src = image.to(self.device, non_blocking=True).contiguous()
        H, W = int(src.shape[0]), int(src.shape[1])
        dst = torch.empty((out_h, out_w, 3), dtype=torch.uint8, device=self.device)
        nbytes = int(self.lib.umv_resize_workspace_bytes(H, W, out_h, out_w))
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.umv_resize_bicubic_u8(_ptr(src), H, W, _ptr(dst), out_h, out_w, _ptr(ws), nbytes, _stream_ptr()))
        return dst

    def vit_embed(self, pixels: torch.Tensor, pos_ids: torch.Tensor, seqlens: Iterable[int]) -> torch.Tensor:
        pixels = pixels.to(self.device, torch.float32).contiguous()
        pos_ids = pos_ids.to(self.device, torch.int64).contiguous()
        lens = [int(x) for x in seqlens]
        out = torch.empty((pixels.shape[0], self.dims.llm.hidden), dtype=torch.bfloat16, device=self.device)
        self._enter()
        _lib.check(self.lib.umv_vit_embed(self.h, _ptr(pixels), _ptr(pos_ids), _lib.i32_array(lens), len(lens), _ptr(out),
                                          _stream_ptr(self.stream)))
        self._exit()
        return out

    def llm_forward(self, x: torch.Tensor, seqs: Sequence[int], q_lens: Sequence[int], positions: Sequence[int],
                    row_is_gen=None, is_causal: bool = True, update_kv: bool = True, want_hidden: bool = True):
        x = x.to(self.device, torch.bfloat16).contiguous()
        M = x.shape[0]
        assert len(positions) == M and sum(int(q) for q in q_lens) == M
        out = torch.empty_like(x) if want_hidden else None
        sel = None
        if row_is_gen is not None:
            sel = (C.c_uint8 * M)(*[int(b) for b in row_is_gen])
        self._enter()
        _lib.check(self.lib.umv_llm_forward(self.h, _ptr(x), len(seqs), _lib.i32_array(seqs), _lib.i32_array(q_lens),
                                            _lib.i32_array(positions), sel, int(is_causal), int(update_kv), _ptr(out),
                                            _stream_ptr(self.stream)))
        self._exit()
        return out

    def attention_block(self, layer: int, seqs: Sequence[int], q_lens: Sequence[int], positions: Sequence[int], qkv=None,
                        partial=None, bias=None, row_is_gen=None, is_causal: bool = True, update_kv: bool = True):
        """umv_op_attention_block: q/k-norm + RoPE + KV append + attention of one layer on given projection outputs.
        Returns (out bf16 [M, heads*head_dim], path) with path 1 mma.sync / 2 tcgen05 / 3 fused decode kernel."""
        M = sum(int(q) for q in q_lens)
        l = self.dims.llm
        out = torch.empty((M, l.heads * l.head_dim), dtype=torch.bfloat16, device=self.device)
        if qkv is not None:
            qkv = qkv.to(self.device, torch.bfloat16).contiguous()
        if partial is not None:
            partial = partial.to(self.device, torch.float32).contiguous()
            bias = bias.to(self.device, torch.bfloat16).contiguous()
        sel = (C.c_uint8 * M)(*[int(b) for b in row_is_gen]) if row_is_gen is not None else None
        path = C.c_int32()
        self._enter()
        _lib.check(self.lib.umv_op_attention_block(
            self.h, layer, _ptr(qkv), None if partial is None else C.c_void_p(partial.data_ptr()),
            0 if partial is None else partial.shape[0], _ptr(bias), len(seqs), _lib.i32_array(seqs), _lib.i32_array(q_lens),
            _lib.i32_array(positions), sel, int(is_causal), int(update_kv), _ptr(out), C.byref(path), _stream_ptr(self.stream)))
        self._exit()
        return out, path.value

    def lm_head(self, hidden: torch.Tensor) -> torch.Tensor:
        hidden = hidden.to(self.device, torch.bfloat16).contiguous()
        out = torch.empty((hidden.shape[0], self.dims.llm.vocab), dtype=torch.bfloat16, device=self.device)
        self._enter()
        _lib.check(self.lib.umv_lm_head(self.h, _ptr(hidden), hidden.shape[0], _ptr(out), _stream_ptr(self.stream)))
        self._exit()
        return out

    def generate_text(self, seqs: Sequence[int], start_tokens: Sequence[int], positions: Sequence[int], n_steps: int,
                      temperature: float = 0.0, seed: int = 0, forced_tokens: torch.Tensor | None = None,
                      return_logits: bool = False, return_next: bool = False):
        """Bagel.generate_text (bagel.py:1236-1317) without the EOS early-exit: returns i64 [n_steps, B]
        (row 0 = start tokens) and optionally bf16 logits [n_steps, B, vocab]."""
        B = len(seqs)
        toks = torch.empty((n_steps, B), dtype=torch.int64, device=self.device)
        logits = torch.empty((n_steps, B, self.dims.llm.vocab), dtype=torch.bfloat16, device=self.device) if return_logits else None
        if forced_tokens is not None:
            forced_tokens = forced_tokens.to(self.device, torch.int64).contiguous()
            assert tuple(forced_tokens.shape) == (n_steps, B)
        nxt = torch.empty((B,), dtype=torch.int64, device=self.device) if return_next else None
        self._enter()
        _lib.check(self.lib.umv_generate_text(self.h, B, _lib.i32_array(seqs), _lib.i64_array(start_tokens),
                                              _lib.i32_array(positions), n_steps, C.c_float(temperature), C.c_uint64(seed),
                                              _ptr(forced_tokens), _ptr(toks), _ptr(logits), _ptr(nxt),
                                              _stream_ptr(self.stream)))
        self._exit()
        out = (toks,) + ((logits,) if return_logits else ()) + ((nxt,) if return_next else ())
        return out if len(out) > 1 else toks

    def flow_velocity(self, x_t: torch.Tensor, lat_pos_ids: torch.Tensor, lat_lens, seqs, positions, marker_ids, timestep: float,
                      cfg_text=None, cfg_img=None, cfg_text_scale: float = 1.0, cfg_img_scale: float = 1.0,
                      cfg_renorm_min: float = 0.0, renorm_type: int = 0, out: torch.Tensor | None = None) -> torch.Tensor:
        """Bagel._forward_flow (bagel.py:989-1211).  x_t fp32 [n_lat, latent_dim] on the device; cfg_text / cfg_img are
        (seqs, positions) of the alternate contexts.  Returns the guided velocity as fp32 values [n_lat, latent_dim]."""
        assert x_t.is_cuda and x_t.dtype == torch.float32 and x_t.is_contiguous()
        lat_pos_ids = lat_pos_ids.to(self.device, torch.int64).contiguous()
        v = out if out is not None else torch.empty_like(x_t)
        keep = [_lib.i32_array(seqs), _lib.i32_array(lat_lens), _lib.i32_array(positions), _lib.i64_array(marker_ids)]
        a = _lib.FlowArgs()
        a.n_seqs = len(seqs)
        a.seqs, a.lat_lens, a.positions, a.marker_ids = keep[0], keep[1], keep[2], keep[3]
        if cfg_text is not None:
            keep += [_lib.i32_array(cfg_text[0]), _lib.i32_array(cfg_text[1])]
            a.cfg_text_seqs, a.cfg_text_positions = keep[-2], keep[-1]
        if cfg_img is not None:
            keep += [_lib.i32_array(cfg_img[0]), _lib.i32_array(cfg_img[1])]
            a.cfg_img_seqs, a.cfg_img_positions = keep[-2], keep[-1]
        a.lat_pos_ids = lat_pos_ids.data_ptr()
        a.timestep, a.cfg_text_scale, a.cfg_img_scale = float(timestep), float(cfg_text_scale), float(cfg_img_scale)
        a.cfg_renorm_min, a.renorm_type = float(cfg_renorm_min), int(renorm_type)
        self._enter()
        _lib.check(self.lib.umv_flow_velocity(self.h, C.byref(a), _ptr(x_t), _ptr(v), _stream_ptr(self.stream)))
        self._exit()
        return v

    def flow_branches_last(self) -> int:
        """CFG branches the last flow_velocity call evaluated (umv_flow_branches_last)."""
        n = C.c_int32()
        _lib.check(self.lib.umv_flow_branches_last(self.h, C.byref(n)))
        return n.value

    def latent_embed(self, x: torch.Tensor, pos_ids: torch.Tensor, timestep: float) -> torch.Tensor:
        x = x.to(self.device, torch.float32).contiguous()
        pos_ids = pos_ids.to(self.device, torch.int64).contiguous()
        out = torch.empty((x.shape[0], self.dims.llm.hidden), dtype=torch.bfloat16, device=self.device)
        self._enter()
        _lib.check(self.lib.umv_latent_embed(self.h, _ptr(x), _ptr(pos_ids), x.shape[0], C.c_float(timestep), _ptr(out),
                                             _stream_ptr(self.stream)))
        self._exit()
        return out

    # ------------------------------------------------------------------ prefill drivers (one C call per reference method)
    def _riders(self, riders, temperature: float, seed: int):
        """riders = (seqs, tokens, positions) of running requests that decode ONE token inside this prefill forward
        (umv_decode_riders).  Returns (struct or None, keep-alive list, device next-token tensor or None)."""
        if not riders or not riders[0]:
            return None, [], None
        seqs, tokens, positions = riders
        nxt = torch.empty((len(seqs),), dtype=torch.int64, device=self.device)
        keep = [_lib.i32_array(seqs), _lib.i64_array(tokens), _lib.i32_array(positions)]
        r = _lib.DecodeRiders()
        r.n, r.seqs, r.tokens, r.positions = len(seqs), keep[0], keep[1], keep[2]
        r.temperature, r.seed, r.next_tokens = float(temperature), int(seed), nxt.data_ptr()
        return C.byref(r), keep + [r], nxt

    def forward_cache_update_text(self, seqs, text_lens, text_ids, positions, riders=None, temperature: float = 0.0, seed: int = 0):
        """umv_forward_cache_update_text[_riders]: host index lists in, KV pages updated.  With `riders` returns their next tokens
        (device i64 [n]): running requests decode one token inside the same packed forward."""
        rd, keep, nxt = self._riders(riders, temperature, seed)
        self._enter()
        _lib.check(self.lib.umv_forward_cache_update_text_riders(self.h, len(seqs), _lib.i32_array(seqs), _lib.i32_array(text_lens),
                                                                 _lib.i64_array(text_ids), _lib.i32_array(positions), rd,
                                                                 _stream_ptr(self.stream)))
        self._exit()
        return nxt

    def forward_cache_update_vit(self, seqs, seq_lens, text_ids, text_rows, pixels, vit_pos_ids, vit_seqlens, vit_rows, positions,
                                 riders=None, temperature: float = 0.0, seed: int = 0, prompt_lens=None):
        """prompt_lens (umv_forward_cache_update_vit_prompt): the last prompt_lens[b] packed rows of sample b are its prompt, prefilled
        causally behind the image block in the same forward."""
        pixels = pixels.to(self.device, torch.float32, non_blocking=True).contiguous()
        vit_pos_ids = vit_pos_ids.to(self.device, torch.int64, non_blocking=True).contiguous()
        rd, keep, nxt = self._riders(riders, temperature, seed)
        self._enter()
        _lib.check(self.lib.umv_forward_cache_update_vit_prompt(
            self.h, len(seqs), _lib.i32_array(seqs), _lib.i32_array(seq_lens),
            _lib.i32_array(prompt_lens) if prompt_lens is not None else None, len(text_ids), _lib.i64_array(text_ids),
            _lib.i32_array(text_rows), _ptr(pixels), _ptr(vit_pos_ids), len(vit_seqlens), _lib.i32_array(vit_seqlens),
            _lib.i32_array(vit_rows), _lib.i32_array(positions), rd, _stream_ptr(self.stream)))
        self._exit()
        return nxt

    def forward_cache_update_vae(self, seqs, seq_lens, text_ids, text_rows, latent, latent_hw, patch, lat_pos_ids, lat_rows, timestep,
                                 positions) -> None:
        latent = latent.to(self.device, torch.bfloat16).contiguous()
        lat_pos_ids = lat_pos_ids.to(self.device, torch.int64).contiguous()
        n, _, Hl, Wl = latent.shape
        hw = [int(v) for pair in latent_hw for v in pair]
        self._enter()
        _lib.check(self.lib.umv_forward_cache_update_vae(
            self.h, len(seqs), _lib.i32_array(seqs), _lib.i32_array(seq_lens), len(text_ids), _lib.i64_array(text_ids),
            _lib.i32_array(text_rows), _ptr(latent), n, Hl, Wl, _lib.i32_array(hw), int(patch), _ptr(lat_pos_ids),
            _lib.i32_array(lat_rows), C.c_float(timestep), _lib.i32_array(positions), _stream_ptr(self.stream)))
        self._exit()

    # ------------------------------------------------------------------ inner modules
    def vit_model(self, pixels: torch.Tensor, pos_ids: torch.Tensor, seqlens: Iterable[int]) -> torch.Tensor:
        pixels = pixels.to(self.device, torch.float32).contiguous()
        pos_ids = pos_ids.to(self.device, torch.int64).contiguous()
        lens = [int(x) for x in seqlens]
        out = torch.empty((pixels.shape[0], self.dims.vit.hidden), dtype=torch.bfloat16, device=self.device)
        self._enter()
        _lib.check(self.lib.umv_vit_model(self.h, _ptr(pixels), _ptr(pos_ids), _lib.i32_array(lens), len(lens), _ptr(out),
                                          _stream_ptr(self.stream)))
        self._exit()
        return out

    def _rows(self, fn, x: torch.Tensor, dtype, cols: int, *extra) -> torch.Tensor:
        x = x.to(self.device, dtype).contiguous()
        out = torch.empty((x.shape[0], cols), dtype=torch.bfloat16, device=self.device)
        self._enter()
        _lib.check(fn(self.h, *extra, _ptr(x), x.shape[0], _ptr(out), _stream_ptr(self.stream)))
        self._exit()
        return out

    def connector(self, x: torch.Tensor) -> torch.Tensor:
        return self._rows(self.lib.umv_connector, x, torch.bfloat16, self.dims.llm.hidden)

    def pos_embed(self, which: int, pos_ids: torch.Tensor) -> torch.Tensor:
        return self._rows(self.lib.umv_pos_embed, pos_ids.reshape(-1), torch.int64, self.dims.llm.hidden, which)

    def vae2llm(self, x: torch.Tensor) -> torch.Tensor:
        return self._rows(self.lib.umv_vae2llm, x, torch.float32, self.dims.llm.hidden)

    def llm2vae(self, h: torch.Tensor) -> torch.Tensor:
        return self._rows(self.lib.umv_llm2vae, h, torch.bfloat16, self.dims.patch_latent_dim)

    def time_embedder(self, t) -> torch.Tensor:
        ts = [float(v) for v in (t.reshape(-1).tolist() if torch.is_tensor(t) else t)]
        out = torch.empty((len(ts), self.dims.llm.hidden), dtype=torch.bfloat16, device=self.device)
        self._enter()
        _lib.check(self.lib.umv_time_embedder(self.h, (C.c_float * len(ts))(*ts), len(ts), _ptr(out), _stream_ptr(self.stream)))
        self._exit()
        return out

    def vae_sample(self, moments: torch.Tensor, noise: torch.Tensor | None) -> torch.Tensor:
        """umv_vae_sample: scale * ((mean + exp(0.5 logvar) * noise) - shift) with torch's bf16 rounding after every op."""
        moments = moments.to(self.device, torch.bfloat16).contiguous()
        n, c2, h, w = moments.shape
        if noise is not None:
            noise = noise.to(self.device, torch.bfloat16).contiguous()
        out = torch.empty((n, c2 // 2, h, w), dtype=torch.bfloat16, device=self.device)
        self._enter()
        _lib.check(self.lib.umv_vae_sample(self.h, _ptr(moments), _ptr(noise), n, h, w, _ptr(out), _stream_ptr(self.stream)))
        self._exit()
        return out

    def decode_image_u8(self, latent_tokens: torch.Tensor, h: int, w: int, out: torch.Tensor | None = None) -> torch.Tensor:
        """umv_decode_image_u8: fp32 latent tokens [n, h*w, patch_latent_dim] (or [h*w, ..]) -> device uint8 [n, 16h, 16w, 3]."""
        x = latent_tokens.to(self.device, torch.float32).contiguous()
        if x.dim() == 2:
            x = x[None]
        p = self.dims.latent_patch_size
        n = x.shape[0]
        assert x.shape[1] == h * w and x.shape[2] == self.dims.patch_latent_dim
        if out is None:
            out = torch.empty((n, 8 * p * h, 8 * p * w, 3), dtype=torch.uint8, device=self.device)
        self._enter()
        _lib.check(self.lib.umv_decode_image_u8(self.h, _ptr(x), n, h, w, p, _ptr(out), _stream_ptr(self.stream)))
        self._exit()
        return out

    def vae_block(self, path: str, x: torch.Tensor) -> torch.Tensor:
        """umv_op_vae_block: one autoencoder block (`path` = the reference's module path) on bf16 [1, C, H, W] or [C, H, W]."""
        x4 = x if x.dim() == 4 else x[None]
        assert x4.shape[0] == 1
        x4 = x4.to(self.device, torch.bfloat16).contiguous()
        _, c, h, w = x4.shape
        out = torch.empty((512 * 4 * h * w,), dtype=torch.bfloat16, device=self.device)
        chw = (C.c_int32 * 3)()
        self._enter()
        _lib.check(self.lib.umv_op_vae_block(self.h, path.encode(), _ptr(x4), c, h, w, _ptr(out), chw, _stream_ptr(self.stream)))
        self._exit()
        co, ho, wo = chw[0], chw[1], chw[2]
        return out[:co * ho * wo].reshape(1, co, ho, wo).clone()

    def vae_decode(self, z: torch.Tensor) -> torch.Tensor:
        """AutoEncoder.decode: bf16 [n, 16, h, w] -> bf16 [n, 3, 8h, 8w]."""
        z = z.to(self.device, torch.bfloat16).contiguous()
        n, c, h, w = z.shape
        out = torch.empty((n, 3, 8 * h, 8 * w), dtype=torch.bfloat16, device=self.device)
        self._enter()
        _lib.check(self.lib.umv_vae_decode(self.h, _ptr(z), n, h, w, _ptr(out), _stream_ptr(self.stream)))
        self._exit()
        return out

    def vae_encode_moments(self, x: torch.Tensor) -> torch.Tensor:
        """Encoder.forward: bf16 [n, 3, H, W] -> bf16 moments [n, 32, H/8, W/8]."""
        x = x.to(self.device, torch.bfloat16).contiguous()
        n, c, H, W = x.shape
        out = torch.empty((n, 2 * self.dims.vae.z_channels, H // 8, W // 8), dtype=torch.bfloat16, device=self.device)
        self._enter()
        _lib.check(self.lib.umv_vae_encode_moments(self.h, _ptr(x), n, H, W, _ptr(out), _stream_ptr(self.stream)))
        self._exit()
        return out

    def flow_euler(self, x_t: torch.Tensor, v: torch.Tensor, dt: float, v_is_bf16: bool) -> None:
        self._enter()
        _lib.check(self.lib.umv_flow_euler(self.h, _ptr(x_t), _ptr(v), x_t.numel(), C.c_float(dt), int(v_is_bf16),
                                           _stream_ptr(self.stream)))
        self._exit()

    def launch_count(self) -> int:
        return int(self.lib.umv_launch_count())


# ------------------------------------------------------------------------ op-level (parity tests)
def op_linear(x, w, bias=None, residual=None, epi: int = 0, impl: int = 0) -> torch.Tensor:
    lib = _lib.load()
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty((M, N // 2 if epi == 2 else N), dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.umv_op_linear(_ptr(x), _ptr(w), _ptr(bias), _ptr(residual), _ptr(y), M, N, K, epi, impl, _stream_ptr()))
    return y


def op_rmsnorm(x, w, eps=1e-6) -> torch.Tensor:
    lib = _lib.load()
    y = torch.empty_like(x)
    _lib.check(lib.umv_op_rmsnorm(_ptr(x), _ptr(w), _ptr(y), x.shape[0], x.shape[1], C.c_float(eps), _stream_ptr()))
    return y


def op_layernorm(x, w, b, eps=1e-6) -> torch.Tensor:
    lib = _lib.load()
    y = torch.empty_like(x)
    _lib.check(lib.umv_op_layernorm(_ptr(x), _ptr(w), _ptr(b), _ptr(y), x.shape[0], x.shape[1], C.c_float(eps), _stream_ptr()))
    return y


def op_attention(q, k, v, q_lens, k_lens, causal: bool) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty_like(q)
    _lib.check(lib.umv_op_attention(_ptr(q), _ptr(k), _ptr(v), _ptr(out), len(q_lens), _lib.i32_array(q_lens),
                                    _lib.i32_array(k_lens), q.shape[1], k.shape[1], q.shape[2], int(causal), _stream_ptr()))
    return out


def op_argmax(logits) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty((logits.shape[0],), dtype=torch.int64, device=logits.device)
    _lib.check(lib.umv_op_argmax(_ptr(logits), logits.shape[0], logits.shape[1], _ptr(out), _stream_ptr()))
    return out


def op_sample(logits, temperature: float, seed: int = 0, u_force: float = -1.0) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty((logits.shape[0],), dtype=torch.int64, device=logits.device)
    _lib.check(lib.umv_op_sample(_ptr(logits), logits.shape[0], logits.shape[1], C.c_float(temperature), C.c_uint64(seed),
                                 C.c_float(u_force), _ptr(out), _stream_ptr()))
    return out
