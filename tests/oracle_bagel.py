"""TEST HARNESS: the oracle (CPU restatement of the reference's arithmetic) dressed in the product's `Bagel` method surface,
so that the product's HOST code -- `unimedvl_b200.InterleaveInferencer` workflows, `Bagel.prepare_*` wrappers, `packing` --
can be exercised end to end on a machine without a GPU and compared with the fixtures produced by the reference's own
`InterleaveInferencer` (tests/golden/make_golden.py).  Nothing here is reachable from the product."""
from __future__ import annotations

import torch

from oracle import vae as ovae
from unimedvl_b200.bagel import Bagel


class OracleCache:
    """Stands in for NaiveCache: holds the oracle's PackedKV; deepcopy clones the tensors (what the reference does)."""

    def __init__(self, num_layers: int):
        self.num_layers = num_layers
        self.kv = None

    def __deepcopy__(self, memo):
        c = OracleCache(self.num_layers)
        c.kv = None if self.kv is None else self.kv.clone()
        return c


class OracleVAE:
    def __init__(self, o):
        self.o = o
        self._p = torch.empty(0, dtype=torch.bfloat16)

    def parameters(self):
        yield self._p

    def decode(self, z):
        return ovae.decode(self.o.vae_sd, z, self.o.dims.vae, self.o.sem)


class OracleBagel(Bagel):
    """Product `Bagel` host methods (prepare_*, attribute surface) over oracle forwards."""

    def __init__(self, o, dims):
        class _NoEngine:
            device = torch.device("cpu")
        super().__init__(_NoEngine(), dims)
        self.o = o

    def _kv(self, cache):
        if cache.kv is None:
            cache.kv = self.o.new_cache()
        return cache.kv

    def forward_cache_update_text(self, past_key_values, **g):
        past_key_values.kv = self.o.forward_cache_update_text(self._kv(past_key_values), **g)
        return past_key_values

    def forward_cache_update_vit(self, past_key_values, **g):
        past_key_values.kv = self.o.forward_cache_update_vit(self._kv(past_key_values), **g)
        return past_key_values

    def forward_cache_update_vae(self, vae_model, past_key_values, **g):
        # DiagonalGaussian.sample (autoencoder.py:270): randn_like(mean) on the default generator -- drawn here exactly as the
        # reference's CPU run draws it, so one torch.manual_seed pins the run
        img = g["padded_images"]
        zc = self.o.dims.vae.z_channels
        noise = torch.randn((img.shape[0], zc, img.shape[2] // 8, img.shape[3] // 8), dtype=torch.bfloat16)
        past_key_values.kv = self.o.forward_cache_update_vae(self._kv(past_key_values), **g, noise=noise)
        return past_key_values

    def generate_text(self, past_key_values, packed_key_value_indexes, key_values_lens, packed_start_tokens,
                      packed_query_position_ids, max_length, do_sample=False, temperature=1.0, end_token_id=None, **kw):
        assert not do_sample
        return self.o.generate_text(self._kv(past_key_values), packed_key_value_indexes, key_values_lens, packed_start_tokens,
                                    packed_query_position_ids, max_length, end_token_id=end_token_id)

    def generate_image(self, past_key_values, cfg_text_past_key_values=None, cfg_img_past_key_values=None, num_timesteps=24,
                       timestep_shift=1.0, cfg_renorm_min=0.0, cfg_renorm_type="global", cfg_interval=(0, 1), cfg_text_scale=1.0,
                       cfg_img_scale=1.0, cfg_type="parallel", **g):
        def branch(prefix, cache):
            if cache is None:
                return None
            d = {k.replace(prefix, "cfg_"): g.pop(k) for k in list(g) if k.startswith(prefix)}
            return dict(d, cache=self._kv(cache))
        ct = branch("cfg_text_", cfg_text_past_key_values)
        ci = branch("cfg_img_", cfg_img_past_key_values)
        return self.o.generate_image(g, self._kv(past_key_values), ct, ci, num_timesteps=num_timesteps,
                                     timestep_shift=timestep_shift, cfg_renorm_min=cfg_renorm_min, cfg_renorm_type=cfg_renorm_type,
                                     cfg_interval=cfg_interval, cfg_text_scale=cfg_text_scale, cfg_img_scale=cfg_img_scale)
