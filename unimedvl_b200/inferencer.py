"""Session driver with the reference's ``InterleaveInferencer`` surface (codes/inferencer.py:31-680):
same constructor, context dict (``kv_lens`` / ``ropes`` / ``past_key_values``), method names, keyword
arguments and return types, so scripts written against the reference keep working.  The reference's own
class also runs unmodified on top of ``unimedvl_b200.Bagel``; this restatement exists so the
path has no dependency on the reference tree and so the two wasteful steps SURVEY.md section 8a (a14) notes
-- prefilling prompts into a CFG context that an understanding request never reads, and deep-copying the
whole KV for ``gen_text`` -- become no-ops (lazy context, page fork) without changing any output.
"""
from __future__ import annotations

from copy import deepcopy
from typing import Any, Dict, List, Optional, Union

import torch
from PIL import Image

from .cache import NaiveCache
from .packing import pil_img2rgb

VLM_THINK_SYSTEM_PROMPT = '''You should first think about the reasoning process in the mind and then provide the user with the answer.
The reasoning process is enclosed within <think> </think> tags, i.e. <think> reasoning process here </think> answer here'''

GEN_THINK_SYSTEM_PROMPT = '''You should first think about the planning process in your mind, and then generate the image.
The planning process is enclosed within <think> </think> tags; that is, <think> planning process here </think> image here.
'''


class InterleaveInferencer:
    def __init__(self, model, vae_model, tokenizer, vae_transform, vit_transform, new_token_ids):
        self.model = model
        self.vae_model = vae_model
        self.tokenizer = tokenizer
        self.vae_transform = vae_transform
        self.vit_transform = vit_transform
        self.new_token_ids = new_token_ids

    # ------------------------------------------------------------------ contexts
    def init_gen_context(self) -> Dict[str, Any]:
        """inferencer.py:73-80."""
        return {"kv_lens": [0], "ropes": [0],
                "past_key_values": NaiveCache(self.model.config.llm_config.num_hidden_layers)}

    @torch.no_grad()
    def update_context_text(self, text: str, gen_context: Dict[str, Any]) -> Dict[str, Any]:
        """inferencer.py:82-128."""
        g, kv_lens, ropes = self.model.prepare_prompts(curr_kvlens=gen_context["kv_lens"], curr_rope=gen_context["ropes"],
                                                       prompts=[text], tokenizer=self.tokenizer,
                                                       new_token_ids=self.new_token_ids)
        pkv = self.model.forward_cache_update_text(gen_context["past_key_values"], **g)
        gen_context.update(kv_lens=kv_lens, ropes=ropes, past_key_values=pkv)
        return gen_context

    @torch.no_grad()
    def update_context_image(self, image, gen_context: Dict[str, Any], vae: bool = True, vit: bool = True) -> Dict[str, Any]:
        """inferencer.py:130-162: VAE latent tokens (generation expert) and / or ViT tokens (understanding expert)."""
        assert vae or vit
        pkv, kv_lens, ropes = gen_context["past_key_values"], gen_context["kv_lens"], gen_context["ropes"]
        if vae:
            g, kv_lens, ropes = self.model.prepare_vae_images(curr_kvlens=kv_lens, curr_rope=ropes, images=[image],
                                                              transforms=self.vae_transform, new_token_ids=self.new_token_ids)
            pkv = self.model.forward_cache_update_vae(self.vae_model, pkv, **g)
        if vit:
            g, kv_lens, ropes = self.model.prepare_vit_images(curr_kvlens=kv_lens, curr_rope=ropes, images=[image],
                                                              transforms=self.vit_transform, new_token_ids=self.new_token_ids)
            pkv = self.model.forward_cache_update_vit(pkv, **g)
        gen_context.update(kv_lens=kv_lens, ropes=ropes, past_key_values=pkv)
        return gen_context

    # ------------------------------------------------------------------ generation
    @torch.no_grad()
    def gen_image(self, image_shape, gen_context, cfg_text_scale=4.0, cfg_img_scale=1.5, cfg_text_precontext=None,
                  cfg_img_precontext=None, cfg_interval=(0.4, 1.0), cfg_renorm_min=0.0, cfg_renorm_type="global",
                  num_timesteps=50, timestep_shift=3.0):
        """inferencer.py:164-232."""
        g = self.model.prepare_vae_latent(curr_kvlens=gen_context["kv_lens"], curr_rope=gen_context["ropes"],
                                          image_sizes=[image_shape], new_token_ids=self.new_token_ids)
        ct = self.model.prepare_vae_latent_cfg(curr_kvlens=cfg_text_precontext["kv_lens"], curr_rope=cfg_text_precontext["ropes"],
                                               image_sizes=[image_shape])
        ci = self.model.prepare_vae_latent_cfg(curr_kvlens=cfg_img_precontext["kv_lens"], curr_rope=cfg_img_precontext["ropes"],
                                               image_sizes=[image_shape])
        latents = self.model.generate_image(
            past_key_values=gen_context["past_key_values"], cfg_text_past_key_values=cfg_text_precontext["past_key_values"],
            cfg_img_past_key_values=cfg_img_precontext["past_key_values"], num_timesteps=num_timesteps,
            cfg_text_scale=cfg_text_scale, cfg_img_scale=cfg_img_scale, cfg_interval=cfg_interval,
            cfg_renorm_min=cfg_renorm_min, cfg_renorm_type=cfg_renorm_type, timestep_shift=timestep_shift, **g,
            cfg_text_packed_position_ids=ct["cfg_packed_position_ids"], cfg_text_packed_query_indexes=ct["cfg_packed_query_indexes"],
            cfg_text_key_values_lens=ct["cfg_key_values_lens"], cfg_text_packed_key_value_indexes=ct["cfg_packed_key_value_indexes"],
            cfg_img_packed_position_ids=ci["cfg_packed_position_ids"], cfg_img_packed_query_indexes=ci["cfg_packed_query_indexes"],
            cfg_img_key_values_lens=ci["cfg_key_values_lens"], cfg_img_packed_key_value_indexes=ci["cfg_packed_key_value_indexes"])
        return self.decode_image(latents[0], image_shape)

    def decode_image(self, latent: torch.Tensor, image_shape) -> Image.Image:
        """inferencer.py:234-256: un-patchify (nhwpqc -> nchpwq), VAE decode, (x*0.5+0.5).clamp(0,1)*255 -> uint8 PIL."""
        return self.decode_images([latent], image_shape)[0]

    def decode_images(self, latents, image_shape) -> List[Image.Image]:
        """decode_image for a batch of same-sized latents: one umv_decode_image_u8 call (un-patchify, VAE decode and the uint8
        conversion run on the device with the reference's bf16 rounding after every op), one asynchronous D2H of the uint8
        HWC images into pinned memory.  A vae_model that is not the engine's AutoEncoder falls back to its own decode()."""
        H, W = image_shape
        m = self.model
        h, w = H // m.latent_downsample, W // m.latent_downsample
        eng = getattr(self.vae_model, "engine", None)
        if eng is None or not hasattr(eng, "decode_image_u8"):
            p, c = m.latent_patch_size, m.latent_channel
            out = []
            for latent in latents:
                z = latent.reshape(1, h, w, p, p, c).permute(0, 5, 1, 3, 2, 4).reshape(1, c, h * p, w * p)
                probe = next(self.vae_model.parameters())
                image = self.vae_model.decode(z.to(device=probe.device, dtype=probe.dtype))
                image = (image * 0.5 + 0.5).clamp(0, 1)[0].permute(1, 2, 0) * 255
                out.append(Image.fromarray(image.to(torch.uint8).cpu().numpy()))
            return out
        x = torch.stack([l.reshape(h * w, -1) for l in latents], 0)
        u8 = eng.decode_image_u8(x, h, w)
        host = torch.empty(u8.shape, dtype=torch.uint8, pin_memory=True)
        host.copy_(u8, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return [Image.fromarray(host[i].numpy()) for i in range(host.shape[0])]

    @torch.no_grad()
    def gen_text(self, gen_context, max_length: int = 500, do_sample: bool = True, temperature: float = 1.0) -> str:
        """inferencer.py:258-279.  deepcopy == page fork: the caller's context is not advanced."""
        ctx = deepcopy(gen_context)
        g = self.model.prepare_start_tokens(ctx["kv_lens"], ctx["ropes"], self.new_token_ids)
        toks = self.model.generate_text(past_key_values=ctx["past_key_values"], max_length=max_length, do_sample=do_sample,
                                        temperature=temperature, end_token_id=self.new_token_ids["eos_token_id"], **g)
        out = self.tokenizer.decode(toks[:, 0])
        return out.split("<|im_end|>")[0].split("<|im_start|>")[1]

    # ------------------------------------------------------------------ workflows
    @torch.no_grad()
    def interleave_inference(self, input_lists: List[Union[str, Image.Image]], think=False, understanding_output=False,
                             max_think_token_n=1000, do_sample=False, text_temperature=0.3, cfg_text_scale=3.0,
                             cfg_img_scale=1.5, cfg_interval=(0.4, 1.0), timestep_shift=3.0, num_timesteps=50,
                             cfg_renorm_min=0.0, cfg_renorm_type="global", image_shapes=(1024, 1024)) -> List[Union[str, Image.Image]]:
        """inferencer.py:551-638.  The cfg_img context (text only) is maintained only when an image will be generated:
        an understanding request never reads it, so its prefills are skipped."""
        output_list: List[Union[str, Image.Image]] = []
        gen_context = self.init_gen_context()
        need_cfg = not understanding_output
        cfg_img_context = deepcopy(gen_context) if need_cfg else None
        cfg_text_context = None
        # Until the first image the image-free context receives exactly the main context's text (inferencer.py:578-600): it is then a
        # page fork of the main context instead of a second prefill -- same K / V, and the flow step evaluates the two CFG branches
        # over it once (umv_flow_velocity).
        same_so_far = True
        if think:
            system_prompt = VLM_THINK_SYSTEM_PROMPT if understanding_output else GEN_THINK_SYSTEM_PROMPT
            gen_context = self.update_context_text(system_prompt, gen_context)
            if need_cfg:
                cfg_img_context = deepcopy(gen_context)
        for term in input_lists:
            if isinstance(term, str):
                if need_cfg:
                    cfg_text_context = deepcopy(gen_context)           # the context WITHOUT this text
                gen_context = self.update_context_text(term, gen_context)
                if need_cfg:
                    cfg_img_context = deepcopy(gen_context) if same_so_far else self.update_context_text(term, cfg_img_context)
            elif isinstance(term, Image.Image):
                term = self.vae_transform.resize_transform(pil_img2rgb(term))
                gen_context = self.update_context_image(term, gen_context, vae=not understanding_output)
                same_so_far = False
                if need_cfg:
                    cfg_text_context = deepcopy(gen_context)
            else:
                raise ValueError(f"Unsupported input type: {type(term)}")
        if understanding_output:
            output_list.append(self.gen_text(gen_context, do_sample=do_sample, temperature=text_temperature,
                                             max_length=max_think_token_n))
            return output_list
        if think:
            text = self.gen_text(gen_context, do_sample=do_sample, temperature=text_temperature, max_length=max_think_token_n)
            gen_context = self.update_context_text(text, gen_context)
            output_list.append(text)
        output_list.append(self.gen_image(image_shapes, gen_context, cfg_text_precontext=cfg_text_context,
                                          cfg_img_precontext=cfg_img_context, cfg_text_scale=cfg_text_scale,
                                          cfg_img_scale=cfg_img_scale, cfg_interval=cfg_interval, timestep_shift=timestep_shift,
                                          num_timesteps=num_timesteps, cfg_renorm_min=cfg_renorm_min,
                                          cfg_renorm_type=cfg_renorm_type))
        return output_list

    # ------------------------------------------------------------------ VQA + reconstruction workflows
    def _calculate_target_size_with_aspect_ratio(self, original_width: int, original_height: int):
        """inferencer.py:42-71: the (H, W) a reconstruction is generated at -- the input's aspect ratio under the VAE
        transform's max_size / min_size / stride / max_pixels rules."""
        rt = self.vae_transform.resize_transform
        stride = rt.stride

        def snap(width, height, scale):
            return tuple(max(stride, int(round(round(v * scale) / stride) * stride)) for v in (width, height))

        scale = max(min(rt.max_size / max(original_width, original_height), 1.0), rt.min_size / min(original_width, original_height))
        w, h = snap(original_width, original_height, scale)
        if w * h > rt.max_pixels:
            w, h = snap(w, h, rt.max_pixels / (w * h))
        if max(w, h) > rt.max_size:
            w, h = snap(w, h, rt.max_size / max(w, h))
        return h, w

    def _vqa_answer(self, input_lists, need_text_only_context: bool, max_think_token_n, do_sample, text_temperature):
        """Common first half of the three variants (inferencer.py:303-323, 392-416, 488-507): every image enters the VQA
        context through BOTH encoders (VAE latent tokens + ViT tokens), then one answer is decoded.  The text-only context the
        reference prefills alongside is read by ver1's reconstruction only, so it is built only when that will run."""
        vqa_context = self.init_gen_context()
        text_only = deepcopy(vqa_context) if need_text_only_context else None
        for term in input_lists:
            if isinstance(term, str):
                vqa_context = self.update_context_text(term, vqa_context)
                if text_only is not None:
                    text_only = self.update_context_text(term, text_only)
            elif isinstance(term, Image.Image):
                vqa_context = self.update_context_image(self.vae_transform.resize_transform(pil_img2rgb(term)), vqa_context,
                                                        vae=True, vit=True)
            else:
                raise ValueError(f"Unsupported input type: {type(term)}")
        answer = self.gen_text(vqa_context, do_sample=do_sample, temperature=text_temperature, max_length=max_think_token_n)
        return answer, vqa_context, text_only

    @torch.no_grad()
    def interleave_inference_for_vqa_reconstruction_ver1(
            self, input_lists, reconstruct_image=False, think=False, understanding_output=True, max_think_token_n=1000,
            do_sample=False, text_temperature=0.3, cfg_text_scale=3.0, cfg_img_scale=1.5, cfg_interval=(0.4, 1.0),
            timestep_shift=3.0, num_timesteps=50, cfg_renorm_min=0.0, cfg_renorm_type="global", image_shapes=(1024, 1024)):
        """inferencer.py:282-362: answer the question, then regenerate every input image conditioned on
        [inputs, answer]; each generated image joins the context (VAE tokens only) before the next one.
        CFG contexts: text = the VQA context without the answer, img = the text-only context plus the answer."""
        images = [t for t in input_lists if isinstance(t, Image.Image)]
        answer, vqa_context, text_only = self._vqa_answer(input_lists, reconstruct_image and bool(images), max_think_token_n,
                                                          do_sample, text_temperature)
        out: List[Union[str, Image.Image]] = [answer]
        if not reconstruct_image or not answer or not answer.strip() or not images:
            return out
        cfg_text = deepcopy(vqa_context)
        cfg_img = self.update_context_text(answer, text_only)
        full = self.update_context_text(answer, deepcopy(vqa_context))
        for original in images:
            shape = self._calculate_target_size_with_aspect_ratio(*original.size)
            generated = self.gen_image(shape, full, cfg_text_precontext=cfg_text, cfg_img_precontext=cfg_img,
                                       cfg_text_scale=cfg_text_scale, cfg_img_scale=cfg_img_scale, cfg_interval=cfg_interval,
                                       timestep_shift=timestep_shift, num_timesteps=num_timesteps, cfg_renorm_min=cfg_renorm_min,
                                       cfg_renorm_type=cfg_renorm_type)
            out.append(generated)
            again = self.vae_transform.resize_transform(pil_img2rgb(generated))
            full = self.update_context_image(again, full, vae=True, vit=False)
            cfg_text = self.update_context_image(again, cfg_text, vae=True, vit=False)
        return out

    def _reconstruct_from_scratch(self, originals, answer, cfg_interval, timestep_shift, num_timesteps, cfg_renorm_min,
                                  cfg_renorm_type):
        """Second half of ver0 / ver0_1 (inferencer.py:427-462, 521-547): per image a FRESH context [image (VAE + ViT), answer];
        CFG text context = the image alone, CFG img context = the answer alone; both scales fixed at 7.0."""
        out = []
        for original in originals:
            shape = self._calculate_target_size_with_aspect_ratio(*original.size)
            processed = self.vae_transform.resize_transform(pil_img2rgb(original))
            cfg_text = self.update_context_image(processed, self.init_gen_context(), vae=True, vit=True)
            full = self.update_context_text(answer, deepcopy(cfg_text))
            cfg_img = self.update_context_text(answer, self.init_gen_context())
            out.append(self.gen_image(shape, full, cfg_text_precontext=cfg_text, cfg_img_precontext=cfg_img, cfg_text_scale=7.0,
                                      cfg_img_scale=7.0, cfg_interval=cfg_interval, timestep_shift=timestep_shift,
                                      num_timesteps=num_timesteps, cfg_renorm_min=cfg_renorm_min, cfg_renorm_type=cfg_renorm_type))
        return out

    @torch.no_grad()
    def interleave_inference_for_vqa_reconstruction_ver0_1(
            self, input_lists, reconstruct_image=False, think=False, understanding_output=True, max_think_token_n=1000,
            do_sample=False, text_temperature=0.3, cfg_text_scale=3.0, cfg_img_scale=1.5, cfg_interval=(0.4, 1.0),
            timestep_shift=3.0, num_timesteps=50, cfg_renorm_min=0.0, cfg_renorm_type="global", image_shapes=(1024, 1024)):
        """inferencer.py:365-462: answer, then reconstruct EVERY input image from a fresh [image, answer] context."""
        answer, _, _ = self._vqa_answer(input_lists, False, max_think_token_n, do_sample, text_temperature)
        out: List[Union[str, Image.Image]] = [answer]
        images = [t for t in input_lists if isinstance(t, Image.Image)]
        if reconstruct_image and answer and answer.strip() and images:
            out += self._reconstruct_from_scratch(images, answer, cfg_interval, timestep_shift, num_timesteps, cfg_renorm_min,
                                                  cfg_renorm_type)
        return out

    @torch.no_grad()
    def interleave_inference_for_vqa_reconstruction_ver0(
            self, input_lists, reconstruct_image=False, think=False, understanding_output=True, max_think_token_n=1000,
            do_sample=False, text_temperature=0.3, cfg_text_scale=3.0, cfg_img_scale=1.5, cfg_interval=(0.4, 1.0),
            timestep_shift=3.0, num_timesteps=50, cfg_renorm_min=0.0, cfg_renorm_type="global", image_shapes=(1024, 1024)):
        """inferencer.py:465-549: as ver0_1 but only the FIRST input image is reconstructed."""
        answer, _, _ = self._vqa_answer(input_lists, False, max_think_token_n, do_sample, text_temperature)
        out: List[Union[str, Image.Image]] = [answer]
        images = [t for t in input_lists if isinstance(t, Image.Image)][:1]
        if reconstruct_image and answer and answer.strip() and images:
            out += self._reconstruct_from_scratch(images, answer, cfg_interval, timestep_shift, num_timesteps, cfg_renorm_min,
                                                  cfg_renorm_type)
        return out

    def __call__(self, image: Optional[Union[Image.Image, List[Image.Image]]] = None, text: Optional[str] = None,
                 inference_ver=0, **kargs) -> Dict[str, Any]:
        """inferencer.py:640-680."""
        out: Dict[str, Any] = {"image": None, "text": None}
        if image is None and text is None:
            return out
        inputs: list = []
        if image is not None:
            inputs.extend(image if isinstance(image, list) else [image])
        if text is not None:
            inputs.append(text)
        if inference_ver == 0:
            items = self.interleave_inference(inputs, **kargs)
        elif inference_ver == 1:
            items = self.interleave_inference_for_vqa_reconstruction_ver1(inputs, **kargs)
        else:
            raise ValueError(f"Unsupported inference_ver: {inference_ver}")
        for item in items:
            if isinstance(item, Image.Image):
                out["image"] = (out["image"] or []) + [item]
            elif isinstance(item, str):
                out["text"] = item
        if isinstance(out["image"], list) and len(out["image"]) == 1:
            out["image"] = out["image"][0]
        return out
