"""Stand-in for diffusers==0.25.0 (test infrastructure; see oracle/shims/README.md)."""
import torch
from torch import nn

from . import configuration_utils, models, utils  # noqa: F401
from .autoencoder_kl import AutoencoderKL  # noqa: F401
from .configuration_utils import ConfigMixin


class ModelMixin(nn.Module):
    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def set_use_memory_efficient_attention_xformers(self, valid, attention_op=None):
        def rec(m):
            if hasattr(m, "set_use_memory_efficient_attention_xformers") and m is not self:
                m.set_use_memory_efficient_attention_xformers(valid, attention_op)
            for ch in m.children():
                rec(ch)

        for ch in self.children():
            rec(ch)


class DiffusionPipeline(ConfigMixin):
    def __init__(self):
        pass

    def register_modules(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def device(self):
        return next(self.unet.parameters()).device

    def to(self, *a, **k):
        for name in ("unet", "vae", "text_encoder", "unet_child"):
            m = getattr(self, name, None)
            if isinstance(m, nn.Module):
                m.to(*a, **k)
        return self


class _Scheduler:
    def __init__(self, *a, **k):
        self.config = k

    def set_timesteps(self, n, device=None):
        self.timesteps = torch.arange(int(n or 1))


class DDIMScheduler(_Scheduler):
    pass


class LCMScheduler(_Scheduler):
    pass


class UNet2DConditionModel:  # type hint only (stablemtl_pipeline.py:139)
    pass
