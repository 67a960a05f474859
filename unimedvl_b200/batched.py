"""The reference's interleaved workflows as BATCHED pipelines on the packed API.

``codes/inferencer.py`` drives one request at a time (``prompts=[text]``, ``images=[image]``, ``image_sizes=[shape]``) although
every ``Bagel.prepare_*`` / ``forward_cache_update_*`` / ``generate_*`` method underneath is packed over samples.  Here a batch
of B requests of the same structure (e.g. B x [image, question]) moves through the same sequence of calls
``interleave_inference`` makes for one (inferencer.py:552-638) -- and through the three VQA + reconstruction variants
(inferencer.py:282-549) and think mode (:574-577, :612-620) -- with every call packed over the batch:

  * contexts are dicts ``{kv_lens: [B], ropes: [B], past_key_values}`` whose cache carries one engine sequence per request;
    ``deepcopy`` is a page fork of all B;
  * ``gen_text`` decodes all requests in one device loop, each stopping at ITS OWN end token (``stop="each"``) -- request b
    gets exactly the text the reference's single-request ``gen_text`` returns for it;
  * ``gen_image`` runs ONE flow loop over B images x 3 CFG branches and decodes the latents to uint8 on the device
    (``umv_decode_image_u8``), same-sized images in one call;
  * samples are independent (SURVEY.md section 8a "Batching semantics verified"; the "global" renorm is per image), so element
    b of every output equals the reference's output for request b alone, given the same noise: the initial latent noise is
    drawn image by image from the CPU generator exactly as ``prepare_vae_latent`` does for a batch (bagel.py:835-837).
"""
from __future__ import annotations

from copy import deepcopy
from typing import Any, Dict, List, Optional, Sequence, Union

import torch
from PIL import Image

from .cache import NaiveCache, PagedKV
from .inferencer import GEN_THINK_SYSTEM_PROMPT, VLM_THINK_SYSTEM_PROMPT, InterleaveInferencer
from .packing import pil_img2rgb

Item = Union[str, Image.Image]


def _kinds(request: Sequence[Item]) -> tuple:
    out = []
    for t in request:
        if isinstance(t, str):
            out.append("text")
        elif isinstance(t, Image.Image):
            out.append("image")
        else:
            raise ValueError(f"Unsupported input type: {type(t)}")          # inferencer.py:610
    return tuple(out)


class BatchedInferencer(InterleaveInferencer):
    """Same constructor as InterleaveInferencer.  Methods take / return one entry per request."""

    # ------------------------------------------------------------------ contexts
    def init_gen_context(self, batch: int = 1) -> Dict[str, Any]:
        return {"kv_lens": [0] * batch, "ropes": [0] * batch,
                "past_key_values": NaiveCache(self.model.config.llm_config.num_hidden_layers)}

    @staticmethod
    def select(ctx: Dict[str, Any], idx: Sequence[int]) -> Dict[str, Any]:
        """The sub-batch `idx` of a context (page forks of the chosen requests)."""
        out = {"kv_lens": [ctx["kv_lens"][i] for i in idx], "ropes": [ctx["ropes"][i] for i in idx],
               "past_key_values": NaiveCache(ctx["past_key_values"].num_layers)}
        h = ctx["past_key_values"]._umv
        if h is not None and h.seqs:
            out["past_key_values"]._umv = PagedKV(h.engine, seqs=[h.engine.seq_fork(h.seqs[i]) for i in idx])
        return out

    @torch.no_grad()
    def update_context_text(self, texts, gen_context):
        texts = [texts] if isinstance(texts, str) else list(texts)
        g, kv_lens, ropes = self.model.prepare_prompts(curr_kvlens=gen_context["kv_lens"], curr_rope=gen_context["ropes"],
                                                       prompts=texts, tokenizer=self.tokenizer, new_token_ids=self.new_token_ids)
        pkv = self.model.forward_cache_update_text(gen_context["past_key_values"], **g)
        gen_context.update(kv_lens=kv_lens, ropes=ropes, past_key_values=pkv)
        return gen_context

    @torch.no_grad()
    def update_context_image(self, images, gen_context, vae: bool = True, vit: bool = True):
        assert vae or vit
        images = [images] if isinstance(images, Image.Image) else list(images)
        pkv, kv_lens, ropes = gen_context["past_key_values"], gen_context["kv_lens"], gen_context["ropes"]
        if vae:
            g, kv_lens, ropes = self.model.prepare_vae_images(curr_kvlens=kv_lens, curr_rope=ropes, images=images,
                                                              transforms=self.vae_transform, new_token_ids=self.new_token_ids)
            pkv = self.model.forward_cache_update_vae(self.vae_model, pkv, **g)
        if vit:
            g, kv_lens, ropes = self.model.prepare_vit_images(curr_kvlens=kv_lens, curr_rope=ropes, images=images,
                                                              transforms=self.vit_transform, new_token_ids=self.new_token_ids)
            pkv = self.model.forward_cache_update_vit(pkv, **g)
        gen_context.update(kv_lens=kv_lens, ropes=ropes, past_key_values=pkv)
        return gen_context

    # ------------------------------------------------------------------ generation
    @torch.no_grad()
    def gen_text(self, gen_context, max_length: int = 500, do_sample: bool = True, temperature: float = 1.0) -> List[str]:
        """inferencer.py:258-279 for every request of the batch; request b ends at its own <|im_end|>."""
        ctx = deepcopy(gen_context)
        g = self.model.prepare_start_tokens(ctx["kv_lens"], ctx["ropes"], self.new_token_ids)
        toks = self.model.generate_text(past_key_values=ctx["past_key_values"], max_length=max_length, do_sample=do_sample,
                                        temperature=temperature, end_token_id=self.new_token_ids["eos_token_id"], stop="each", **g)
        toks = toks.cpu()
        out = []
        for b in range(toks.shape[1]):
            text = self.tokenizer.decode(toks[:, b])
            out.append(text.split("<|im_end|>")[0].split("<|im_start|>")[1])
        return out

    @torch.no_grad()
    def gen_image(self, image_shapes, gen_context, cfg_text_scale=4.0, cfg_img_scale=1.5, cfg_text_precontext=None,
                  cfg_img_precontext=None, cfg_interval=(0.4, 1.0), cfg_renorm_min=0.0, cfg_renorm_type="global", num_timesteps=50,
                  timestep_shift=3.0, as_uint8: bool = False):
        """inferencer.py:164-232 for the whole batch: one flow loop over B images x CFG branches.  `image_shapes`: one (H, W) for
        all or one per request.  Returns a list of PIL images, or with `as_uint8` (all shapes equal) a DEVICE uint8 tensor
        [B, H, W, 3] (what a server hands to its encoder / an all_gather)."""
        B = len(gen_context["kv_lens"])
        shapes = [tuple(image_shapes)] * B if isinstance(image_shapes[0], int) else [tuple(s) for s in image_shapes]
        g = self.model.prepare_vae_latent(curr_kvlens=gen_context["kv_lens"], curr_rope=gen_context["ropes"], image_sizes=shapes,
                                          new_token_ids=self.new_token_ids)
        ct = self.model.prepare_vae_latent_cfg(curr_kvlens=cfg_text_precontext["kv_lens"], curr_rope=cfg_text_precontext["ropes"],
                                               image_sizes=shapes)
        ci = self.model.prepare_vae_latent_cfg(curr_kvlens=cfg_img_precontext["kv_lens"], curr_rope=cfg_img_precontext["ropes"],
                                               image_sizes=shapes)
        latents = self.model.generate_image(
            past_key_values=gen_context["past_key_values"], cfg_text_past_key_values=cfg_text_precontext["past_key_values"],
            cfg_img_past_key_values=cfg_img_precontext["past_key_values"], num_timesteps=num_timesteps,
            cfg_text_scale=cfg_text_scale, cfg_img_scale=cfg_img_scale, cfg_interval=cfg_interval,
            cfg_renorm_min=cfg_renorm_min, cfg_renorm_type=cfg_renorm_type, timestep_shift=timestep_shift, **g,
            cfg_text_packed_position_ids=ct["cfg_packed_position_ids"], cfg_text_packed_query_indexes=ct["cfg_packed_query_indexes"],
            cfg_text_key_values_lens=ct["cfg_key_values_lens"], cfg_text_packed_key_value_indexes=ct["cfg_packed_key_value_indexes"],
            cfg_img_packed_position_ids=ci["cfg_packed_position_ids"], cfg_img_packed_query_indexes=ci["cfg_packed_query_indexes"],
            cfg_img_key_values_lens=ci["cfg_key_values_lens"], cfg_img_packed_key_value_indexes=ci["cfg_packed_key_value_indexes"])
        if as_uint8:
            if len(set(shapes)) != 1:
                raise ValueError("as_uint8 needs one image size for the whole batch")
            m = self.model
            h, w = shapes[0][0] // m.latent_downsample, shapes[0][1] // m.latent_downsample
            return self.vae_model.engine.decode_image_u8(torch.stack([l.reshape(h * w, -1) for l in latents], 0), h, w)
        out: List[Optional[Image.Image]] = [None] * B
        for shape in sorted(set(shapes)):               # same-sized images decode in one call
            idx = [i for i, s in enumerate(shapes) if s == shape]
            for i, im in zip(idx, self.decode_images([latents[i] for i in idx], shape)):
                out[i] = im
        return out

    # ------------------------------------------------------------------ workflows
    @staticmethod
    def _columns(requests: Sequence[Sequence[Item]]):
        kinds = {_kinds(r) for r in requests}
        if len(kinds) != 1:
            raise ValueError("a batch holds requests of one structure (same sequence of text / image inputs)")
        return list(kinds.pop()), [[r[j] for r in requests] for j in range(len(requests[0]))] if requests and requests[0] else []

    @torch.no_grad()
    def interleave_inference(self, requests, think=False, understanding_output=False, max_think_token_n=1000, do_sample=False,
                             text_temperature=0.3, cfg_text_scale=3.0, cfg_img_scale=1.5, cfg_interval=(0.4, 1.0), timestep_shift=3.0,
                             num_timesteps=50, cfg_renorm_min=0.0, cfg_renorm_type="global", image_shapes=(1024, 1024),
                             return_uint8: bool = False):
        """inferencer.py:551-638 for a batch: `requests` is a list of input lists (or one input list: the reference's call).
        Returns one output list per request.  `return_uint8`: generated images come back as DEVICE uint8 [H, W, 3] tensors (one
        image size for the batch) instead of PIL images -- for a server's encoder or the data-parallel image gather."""
        single = bool(requests) and isinstance(requests[0], (str, Image.Image))
        reqs = [list(requests)] if single else [list(r) for r in requests]
        B = len(reqs)
        kinds, cols = self._columns(reqs)
        outs: List[List[Item]] = [[] for _ in range(B)]
        gen_context = self.init_gen_context(B)
        need_cfg = not understanding_output
        cfg_img_context = deepcopy(gen_context) if need_cfg else None
        cfg_text_context = None
        # The image-free context (cfg_img) receives every text the main context receives (inferencer.py:578-600).  Until the first image
        # the two hold the same tokens, hence the same K / V: cfg_img is then a page fork of the main context, not a second prefill.
        same_so_far = True
        if think:
            system_prompt = VLM_THINK_SYSTEM_PROMPT if understanding_output else GEN_THINK_SYSTEM_PROMPT
            gen_context = self.update_context_text([system_prompt] * B, gen_context)
            if need_cfg:
                cfg_img_context = deepcopy(gen_context)
        for kind, col in zip(kinds, cols):
            if kind == "text":
                if need_cfg:
                    cfg_text_context = deepcopy(gen_context)
                gen_context = self.update_context_text(col, gen_context)
                if need_cfg:
                    cfg_img_context = deepcopy(gen_context) if same_so_far else self.update_context_text(col, cfg_img_context)
            else:
                col = [self.vae_transform.resize_transform(pil_img2rgb(im)) for im in col]
                gen_context = self.update_context_image(col, gen_context, vae=not understanding_output)
                same_so_far = False
                if need_cfg:
                    cfg_text_context = deepcopy(gen_context)
        if understanding_output:
            for b, t in enumerate(self.gen_text(gen_context, do_sample=do_sample, temperature=text_temperature, max_length=max_think_token_n)):
                outs[b].append(t)
            return outs[0] if single else outs
        if think:
            plans = self.gen_text(gen_context, do_sample=do_sample, temperature=text_temperature, max_length=max_think_token_n)
            gen_context = self.update_context_text(plans, gen_context)
            for b, t in enumerate(plans):
                outs[b].append(t)
        if cfg_text_context is None:
            cfg_text_context = self.init_gen_context(B)
        imgs = self.gen_image(image_shapes, gen_context, cfg_text_precontext=cfg_text_context, cfg_img_precontext=cfg_img_context,
                              cfg_text_scale=cfg_text_scale, cfg_img_scale=cfg_img_scale, cfg_interval=cfg_interval,
                              timestep_shift=timestep_shift, num_timesteps=num_timesteps, cfg_renorm_min=cfg_renorm_min,
                              cfg_renorm_type=cfg_renorm_type, as_uint8=return_uint8)
        for b in range(B):
            outs[b].append(imgs[b])
        return outs[0] if single else outs

    # ------------------------------------------------------------------ VQA + reconstruction, batched
    def _vqa_answer(self, reqs, need_text_only_context, max_think_token_n, do_sample, text_temperature):
        B = len(reqs)
        kinds, cols = self._columns(reqs)
        vqa_context = self.init_gen_context(B)
        text_only = deepcopy(vqa_context) if need_text_only_context else None
        for kind, col in zip(kinds, cols):
            if kind == "text":
                vqa_context = self.update_context_text(col, vqa_context)
                if text_only is not None:
                    text_only = self.update_context_text(col, text_only)
            else:
                col = [self.vae_transform.resize_transform(pil_img2rgb(im)) for im in col]
                vqa_context = self.update_context_image(col, vqa_context, vae=True, vit=True)
        answers = self.gen_text(vqa_context, do_sample=do_sample, temperature=text_temperature, max_length=max_think_token_n)
        return answers, vqa_context, text_only

    @torch.no_grad()
    def vqa_reconstruction(self, requests, variant: str = "ver1", reconstruct_image=False, max_think_token_n=1000, do_sample=False,
                           text_temperature=0.3, cfg_text_scale=3.0, cfg_img_scale=1.5, cfg_interval=(0.4, 1.0), timestep_shift=3.0,
                           num_timesteps=50, cfg_renorm_min=0.0, cfg_renorm_type="global"):
        """interleave_inference_for_vqa_reconstruction_{ver1, ver0_1, ver0} (inferencer.py:282-549) over a batch of requests of one
        structure.  Returns per request [answer, reconstructed images...].  Requests whose answer is empty are not reconstructed
        (inferencer.py:325, 418, 509)."""
        if variant not in ("ver1", "ver0_1", "ver0"):
            raise ValueError(f"Unsupported inference_ver: {variant}")
        reqs = [list(r) for r in requests]
        images = [[t for t in r if isinstance(t, Image.Image)] for r in reqs]
        n_img = len(images[0]) if images else 0
        answers, vqa_context, text_only = self._vqa_answer(reqs, variant == "ver1" and reconstruct_image and n_img > 0,
                                                           max_think_token_n, do_sample, text_temperature)
        outs: List[List[Item]] = [[a] for a in answers]
        live = [b for b, a in enumerate(answers) if a and a.strip()]
        if not reconstruct_image or not n_img or not live:
            return outs
        ans = [answers[b] for b in live]
        kw = dict(cfg_interval=cfg_interval, timestep_shift=timestep_shift, num_timesteps=num_timesteps, cfg_renorm_min=cfg_renorm_min,
                  cfg_renorm_type=cfg_renorm_type)
        if variant == "ver1":
            cfg_text = self.select(vqa_context, live)
            cfg_img = self.update_context_text(ans, self.select(text_only, live))
            full = self.update_context_text(ans, self.select(vqa_context, live))
            for j in range(n_img):
                shapes = [self._calculate_target_size_with_aspect_ratio(*images[b][j].size) for b in live]
                gen = self.gen_image(shapes, full, cfg_text_precontext=cfg_text, cfg_img_precontext=cfg_img,
                                     cfg_text_scale=cfg_text_scale, cfg_img_scale=cfg_img_scale, **kw)
                for b, im in zip(live, gen):
                    outs[b].append(im)
                again = [self.vae_transform.resize_transform(pil_img2rgb(im)) for im in gen]
                full = self.update_context_image(again, full, vae=True, vit=False)
                cfg_text = self.update_context_image(again, cfg_text, vae=True, vit=False)
            return outs
        for j in range(n_img if variant == "ver0_1" else 1):            # a FRESH [image, answer] context per reconstructed image
            originals = [images[b][j] for b in live]
            shapes = [self._calculate_target_size_with_aspect_ratio(*im.size) for im in originals]
            processed = [self.vae_transform.resize_transform(pil_img2rgb(im)) for im in originals]
            cfg_text = self.update_context_image(processed, self.init_gen_context(len(live)), vae=True, vit=True)
            full = self.update_context_text(ans, deepcopy(cfg_text))
            cfg_img = self.update_context_text(ans, self.init_gen_context(len(live)))
            gen = self.gen_image(shapes, full, cfg_text_precontext=cfg_text, cfg_img_precontext=cfg_img, cfg_text_scale=7.0,
                                 cfg_img_scale=7.0, **kw)
            for b, im in zip(live, gen):
                outs[b].append(im)
        return outs

    def __call__(self, images=None, texts=None, inference_ver=0, **kargs) -> List[Dict[str, Any]]:
        """Batched inferencer.py:640-680: `images` / `texts` hold one entry per request (an entry of `images` may be a list of
        images).  Returns one {"image", "text"} dict per request."""
        n = len(texts) if texts is not None else (len(images) if images is not None else 0)
        if n == 0:
            return []
        reqs = []
        for b in range(n):
            r: list = []
            if images is not None and images[b] is not None:
                r.extend(images[b] if isinstance(images[b], list) else [images[b]])
            if texts is not None and texts[b] is not None:
                r.append(texts[b])
            reqs.append(r)
        if inference_ver == 0:
            items = self.interleave_inference(reqs, **kargs)
        elif inference_ver == 1:
            kargs.pop("think", None), kargs.pop("understanding_output", None), kargs.pop("image_shapes", None)
            items = self.vqa_reconstruction(reqs, "ver1", **kargs)
        else:
            raise ValueError(f"Unsupported inference_ver: {inference_ver}")
        outs = []
        for it in items:
            o: Dict[str, Any] = {"image": None, "text": None}
            for x in it:
                if isinstance(x, Image.Image):
                    o["image"] = (o["image"] or []) + [x]
                elif isinstance(x, str):
                    o["text"] = x
            if isinstance(o["image"], list) and len(o["image"]) == 1:
                o["image"] = o["image"][0]
            outs.append(o)
        return outs
