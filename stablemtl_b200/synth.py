"""Synthetic (seeded, random-init) checkpoints and inputs with the reference's state-dict key layout.

There is no network in the build/bench environment, so the SD-2 UNet/VAE weights cannot be downloaded; every
test and benchmark uses these deterministic stand-ins.  Each tensor is drawn from its own
generator seeded by crc32(key) ^ seed, so the oracle and the CUDA path see bit-identical fp32 weights no matter
in which order they ask for them.  Key names follow SURVEY.md Appendix B and are validated by loading them with
strict=True into the reference's own modules (oracle/make_golden.py).
"""
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch

TASKS = ["normal", "depth", "semantic", "optical_flow", "scene_flow", "albedo", "shading"]  # config/train_stablemtl.yaml:30-37
FLOW_TASKS = ("optical_flow", "scene_flow")                                               # stablemtl_pipeline.py:433


@dataclass(frozen=True)
class UNetConfig:
    in_channels: int = 12
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    heads: Tuple[int, ...] = (5, 10, 20, 20)          # SD-2 attention_head_dim, used as the head COUNT (unet.py:135)
    layers_per_block: int = 2
    cross_attention_dim: int = 1024
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    task_q_hidden: int = 640                          # MLPv2 hidden width (util/model.py:130)
    n_attns: int = 4                                  # task-attention heads (train_stablemtl.yaml:23)

    def as_reference_kwargs(self):
        return dict(in_channels=4, out_channels=self.out_channels, block_out_channels=self.block_out_channels,
                    attention_head_dim=self.heads, layers_per_block=self.layers_per_block,
                    cross_attention_dim=self.cross_attention_dim, norm_num_groups=self.norm_num_groups,
                    norm_eps=self.norm_eps, use_linear_projection=True)

    def transformer_dims(self) -> List[int]:
        """width of the 16 (for 4 levels) transformer layers in forward order (util/model.py:67-84)."""
        c = self.block_out_channels
        n = len(c)
        down = [c[i] for i in range(n - 1) for _ in range(self.layers_per_block)]
        up = [c[i] for i in range(n - 2, -1, -1) for _ in range(self.layers_per_block + 1)]
        return down + [c[-1]] + up


@dataclass(frozen=True)
class VAEConfig:
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    latent_channels: int = 4
    norm_num_groups: int = 32


SD2_UNET = UNetConfig()
SD2_VAE = VAEConfig()
TINY_UNET = UNetConfig(block_out_channels=(64, 128, 256, 256), heads=(1, 2, 4, 4), cross_attention_dim=128,
                       task_q_hidden=64)
TINY_VAE = VAEConfig(block_out_channels=(64, 64, 128, 128))


def _draw(key: str, shape, seed: int, kind: str) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    if kind == "weight":
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return torch.randn(shape, generator=g) * (fan_in ** -0.5)
    if kind == "norm_weight":
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    return 0.05 * torch.randn(shape, generator=g)   # biases


class _SD:
    def __init__(self, seed):
        self.seed = seed
        self.sd: Dict[str, torch.Tensor] = {}

    def lin(self, p, cout, cin, bias=True):
        self.sd[p + ".weight"] = _draw(p + ".weight", (cout, cin), self.seed, "weight")
        if bias:
            self.sd[p + ".bias"] = _draw(p + ".bias", (cout,), self.seed, "bias")

    def conv(self, p, cout, cin, k):
        self.sd[p + ".weight"] = _draw(p + ".weight", (cout, cin, k, k), self.seed, "weight")
        self.sd[p + ".bias"] = _draw(p + ".bias", (cout,), self.seed, "bias")

    def norm(self, p, c):
        self.sd[p + ".weight"] = _draw(p + ".weight", (c,), self.seed, "norm_weight")
        self.sd[p + ".bias"] = _draw(p + ".bias", (c,), self.seed, "bias")


def _unet_resnet(b: _SD, p, cin, cout, temb):
    b.norm(p + ".norm1", cin)
    b.conv(p + ".conv1", cout, cin, 3)
    b.lin(p + ".time_emb_proj", cout, temb)
    b.norm(p + ".norm2", cout)
    b.conv(p + ".conv2", cout, cout, 3)
    if cin != cout:
        b.conv(p + ".conv_shortcut", cout, cin, 1)


def _unet_transformer(b: _SD, p, c, cross):
    b.norm(p + ".norm", c)
    b.lin(p + ".proj_in", c, c)
    t = p + ".transformer_blocks.0"
    for n in ("norm1", "norm2", "norm3"):
        b.norm(f"{t}.{n}", c)
    for n in ("to_q", "to_k", "to_v"):
        b.lin(f"{t}.attn1.{n}", c, c, bias=False)
    b.lin(f"{t}.attn1.to_out.0", c, c)
    b.lin(f"{t}.attn2.to_q", c, c, bias=False)
    b.lin(f"{t}.attn2.to_k", c, cross, bias=False)
    b.lin(f"{t}.attn2.to_v", c, cross, bias=False)
    b.lin(f"{t}.attn2.to_out.0", c, c)
    b.lin(f"{t}.ff.net.0.proj", 8 * c, c)
    b.lin(f"{t}.ff.net.2", c, 4 * c)
    b.lin(p + ".proj_out", c, c)


def unet_transformer_prefixes(cfg: UNetConfig) -> List[str]:
    """state-dict prefixes of the transformer blocks in forward order == task-feature layer index."""
    n = len(cfg.block_out_channels)
    out = [f"down_blocks.{i}.attentions.{j}" for i in range(n - 1) for j in range(cfg.layers_per_block)]
    out.append("mid_block.attentions.0")
    out += [f"up_blocks.{i}.attentions.{j}" for i in range(1, n) for j in range(cfg.layers_per_block + 1)]
    return out


def make_unet_state_dict(cfg: UNetConfig = SD2_UNET, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Single-stream / child UNet (UNet3DConditionModel.state_dict() after _replace_unet_conv_in)."""
    b = _SD(seed)
    c = cfg.block_out_channels
    n = len(c)
    temb = c[0] * 4
    b.conv("conv_in", c[0], cfg.in_channels, 3)
    b.lin("time_embedding.linear_1", temb, c[0])
    b.lin("time_embedding.linear_2", temb, temb)
    ch = c[0]
    for i in range(n):
        cin, ch = ch, c[i]
        for j in range(cfg.layers_per_block):
            _unet_resnet(b, f"down_blocks.{i}.resnets.{j}", cin if j == 0 else ch, ch, temb)
            if i < n - 1:
                _unet_transformer(b, f"down_blocks.{i}.attentions.{j}", ch, cfg.cross_attention_dim)
        if i < n - 1:
            b.conv(f"down_blocks.{i}.downsamplers.0.conv", ch, ch, 3)
    _unet_resnet(b, "mid_block.resnets.0", c[-1], c[-1], temb)
    _unet_transformer(b, "mid_block.attentions.0", c[-1], cfg.cross_attention_dim)
    _unet_resnet(b, "mid_block.resnets.1", c[-1], c[-1], temb)
    rev = list(reversed(c))
    out_ch = rev[0]
    for i in range(n):
        prev = out_ch
        out_ch = rev[i]
        in_ch = rev[min(i + 1, n - 1)]
        for j in range(cfg.layers_per_block + 1):
            skip = in_ch if j == cfg.layers_per_block else out_ch
            rin = prev if j == 0 else out_ch
            _unet_resnet(b, f"up_blocks.{i}.resnets.{j}", rin + skip, out_ch, temb)
            if i > 0:
                _unet_transformer(b, f"up_blocks.{i}.attentions.{j}", out_ch, cfg.cross_attention_dim)
        if i < n - 1:
            b.conv(f"up_blocks.{i}.upsamplers.0.conv", out_ch, out_ch, 3)
    b.norm("conv_norm_out", c[0])
    b.conv("conv_out", cfg.out_channels, c[0], 3)
    return b.sd


def make_task_modules_state_dict(cfg: UNetConfig = SD2_UNET, tasks=TASKS, seed: int = 1) -> Dict[str, torch.Tensor]:
    """The multi-stream additions on every attn1 of the MAIN UNet (util/model.py:102-146).  `to_out_task` is
    zero-initialised in the reference; it is randomised here so the task branch is not vacuous."""
    b = _SD(seed)
    for p, c in zip(unet_transformer_prefixes(cfg), cfg.transformer_dims()):
        a = f"{p}.transformer_blocks.0.attn1"
        hq = cfg.task_q_hidden
        for t in tasks:
            for kv in ("k", "v"):
                b.lin(f"{a}.task_to_{kv}.{t}.fc1", c // 2, c)
                b.lin(f"{a}.task_to_{kv}.{t}.fc2", c, c // 2)
            b.lin(f"{a}.task_to_q.{t}.net.0", hq, c)
            b.lin(f"{a}.task_to_q.{t}.net.2", hq, hq)
            b.lin(f"{a}.task_to_q.{t}.net.4", hq, hq)
            b.lin(f"{a}.task_to_q.{t}.net.6", c, hq)
            for kv in ("k", "v", "q"):
                b.norm(f"{a}.task_norm_{kv}.{t}", c)
        b.lin(f"{a}.to_out_task", c, c)
    return b.sd


def _vae_resnet(b: _SD, p, cin, cout):
    b.norm(p + ".norm1", cin)
    b.conv(p + ".conv1", cout, cin, 3)
    b.norm(p + ".norm2", cout)
    b.conv(p + ".conv2", cout, cout, 3)
    if cin != cout:
        b.conv(p + ".conv_shortcut", cout, cin, 1)


def _vae_mid(b: _SD, p, c):
    _vae_resnet(b, p + ".resnets.0", c, c)
    a = p + ".attentions.0"
    b.norm(a + ".group_norm", c)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        b.lin(f"{a}.{n}", c, c)
    _vae_resnet(b, p + ".resnets.1", c, c)


def make_vae_state_dict(cfg: VAEConfig = SD2_VAE, seed: int = 2) -> Dict[str, torch.Tensor]:
    b = _SD(seed)
    c = cfg.block_out_channels
    n = len(c)
    lat = cfg.latent_channels
    b.conv("encoder.conv_in", c[0], 3, 3)
    ch = c[0]
    for i in range(n):
        cin, ch = ch, c[i]
        for j in range(cfg.layers_per_block):
            _vae_resnet(b, f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else ch, ch)
        if i < n - 1:
            b.conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", ch, ch, 3)
    _vae_mid(b, "encoder.mid_block", c[-1])
    b.norm("encoder.conv_norm_out", c[-1])
    b.conv("encoder.conv_out", 2 * lat, c[-1], 3)
    b.conv("quant_conv", 2 * lat, 2 * lat, 1)
    b.conv("post_quant_conv", lat, lat, 1)
    rev = list(reversed(c))
    b.conv("decoder.conv_in", rev[0], lat, 3)
    _vae_mid(b, "decoder.mid_block", rev[0])
    ch = rev[0]
    for i in range(n):
        cin, ch = ch, rev[i]
        for j in range(cfg.layers_per_block + 1):
            _vae_resnet(b, f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else ch, ch)
        if i < n - 1:
            b.conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", ch, ch, 3)
    b.norm("decoder.conv_norm_out", c[0])
    b.conv("decoder.conv_out", 3, c[0], 3)
    return b.sd


TASK_NTOK = {"normal": 3, "depth": 3, "semantic": 3, "optical_flow": 4, "scene_flow": 4, "albedo": 3, "shading": 3}


def make_text_embeddings(cross_dim: int = 1024, tasks=TASKS, seed: int = 3) -> Dict[str, torch.Tensor]:
    """Stand-in for CLIP last_hidden_state of the task-name prompt: [n_tok, cross_dim], n_tok = 3 (one word) or 4
    ("optical flow", "scene flow") incl. BOS/EOS (stablemtl_pipeline.py:464-472).  CLIP is outside the path."""
    return {t: 20.0 * _draw(f"text.{t}", (TASK_NTOK[t], cross_dim), seed, "bias") for t in tasks}   # unit variance


def make_images(batch: int, h: int, w: int, seed: int = 0):
    """rgb and next-frame rgb in [0,255] as floats, NCHW."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    rgb = torch.randint(0, 256, (batch, 3, h, w), generator=g).float()
    g2 = torch.Generator(device="cpu").manual_seed(seed + 1)
    nxt = torch.randint(0, 256, (batch, 3, h, w), generator=g2).float()
    return rgb, nxt
