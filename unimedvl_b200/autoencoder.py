"""Stand-in for the reference's ``AutoEncoder`` module (codes/modeling/autoencoder.py:275-322) as callers use
it: ``vae_model.decode(latent)`` (inferencer.py:251), ``vae_model.encode(images)`` (bagel.py:757) and
``next(vae_model.parameters())`` for device / dtype discovery (inferencer.py:243-249)."""
from __future__ import annotations

import torch

from .engine import Engine


class AutoEncoder:
    def __init__(self, engine: Engine):
        self.engine = engine
        v = engine.dims.vae
        self.scale_factor, self.shift_factor = v.scale_factor, v.shift_factor
        self._probe = torch.empty(0, dtype=torch.bfloat16, device=engine.device)
        self.sample = True                  # DiagonalGaussian(sample=True), autoencoder.py:260-272
        self.generator: torch.Generator | None = None
        self.noise_device = "cuda"          # "cpu": draw like a CPU run of the reference (torch.randn_like on the CPU generator)

    def parameters(self):
        yield self._probe

    def eval(self):
        return self

    def load_state_dict(self, sd: dict) -> None:
        """Keys as AutoEncoder.state_dict() (encoder.* / decoder.*)."""
        strip = lambda k: k[len("module."):] if k.startswith("module.") else k       # as load_ae does (autoencoder.py:352-361)
        self.engine.load_state_dict({"vae_model." + strip(k): v for k, v in sd.items()})

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """autoencoder.py:305-307 (the z / scale + shift affine runs inside the engine)."""
        return self.engine.vae_decode(z)

    @torch.no_grad()
    def encode_moments(self, x: torch.Tensor) -> torch.Tensor:
        return self.engine.vae_encode_moments(x)

    @torch.no_grad()
    def encode(self, x: torch.Tensor, noise: torch.Tensor | None = None) -> torch.Tensor:
        """autoencoder.py:300-303: z = mean + exp(0.5 logvar) * randn (device RNG -- not reproducible across
        implementations, so `noise` may be injected), then scale * (z - shift); bf16 tensor ops as in the reference."""
        m = self.encode_moments(x)
        if self.sample and noise is None:           # the draw itself is torch's generator, as in the reference (torch.randn_like)
            shape = (m.shape[0], m.shape[1] // 2, m.shape[2], m.shape[3])
            dev = m.device if self.noise_device == "cuda" else torch.device("cpu")
            noise = torch.randn(shape, dtype=m.dtype, device=dev, generator=self.generator)
        return self.engine.vae_sample(m, noise if self.sample else None)
