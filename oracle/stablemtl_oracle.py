"""CPU/fp32 ORACLE for the StableMTL single-step latent pass.  TEST INFRASTRUCTURE -- never imported by the
product path (`stablemtl_b200/`); only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may use it, and only as the checker / the reported baseline.

It is a plain-PyTorch restatement of the reference's algorithm, written against state dicts with the
reference's own key layout.  Every function cites the reference lines it follows
(paths relative to /root/reference).  The third-party arithmetic the reference delegates to
diffusers==0.25.0 / xformers==0.0.27 (absent from /root/reference and from this image) is restated from the
pinned versions' published behaviour.

PINNING: the reference ships no tests or golden vectors for this path, so this file is pinned
against OUTPUTS OF THE REFERENCE ITSELF: `oracle/make_golden.py` imports the unmodified reference modules from
/root/reference (through `oracle/shims/`), loads the same seeded weights, and checks this restatement against them
(float64 run of both: rel-L2 <= 1e-7, so fp32 summation-order noise cannot hide a mismatch) before writing `tests/golden/*.pt`.  The part of the path that lives in diffusers (VAE,
Attention, GEGLU, timestep embedding) is pinned only against the shims' restatement of diffusers -- no copy of
diffusers exists offline -- so for those pieces parity is "pinned to the published algorithm", not to diffusers' code.
"""
import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

TASKS = ["normal", "depth", "semantic", "optical_flow", "scene_flow", "albedo", "shading"]
FLOW_TASKS = ("optical_flow", "scene_flow")
LATENT_SCALE = 0.18215           # src/stablemtl_pipeline.py:134-135
# VKitti2Encoder(n_classes=8).class_color_embeddings: src/dataset/semantic/labels.py:42-55 (vk-cs colours)
# through mappings.py cls08, see SURVEY.md §8 a18
PALETTE = [[128, 64, 128], [70, 70, 70], [153, 153, 153], [250, 170, 30], [220, 220, 0], [107, 142, 35],
           [70, 130, 180], [0, 0, 142]]


def _gn(sd, p, x, groups, eps):
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def _attend(q, k, v, heads, fp16_inputs=False):
    """softmax(q k^T / sqrt(d)) v per head on [B, N, heads*d] tensors.
    Reference: src/model/attention.py:410-425 (head split) + xformers.ops.memory_efficient_attention (:395).
    fp16_inputs=True reproduces the reference's cast of q/k/v to fp16 and of the result back (:392-394,419)."""
    B, Nq, C = q.shape
    d = C // heads

    def split(t):
        return t.reshape(B, t.shape[1], heads, d).permute(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    dt = q.dtype
    if fp16_inputs:
        q, k, v = q.half().to(dt), k.half().to(dt), v.half().to(dt)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d)
    o = torch.matmul(torch.softmax(s, dim=-1), v)
    if fp16_inputs:
        o = o.half().to(dt)
    return o.permute(0, 2, 1, 3).reshape(B, Nq, C)


# ------------------------------------------------------------------------------------------------- UNet
def timestep_embedding(sd, t: int, dim: int):
    """diffusers Timesteps(dim, flip_sin_to_cos=True, freq_shift=0) + TimestepEmbedding (src/model/unet.py:92-95,347-353)."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    e = torch.tensor([float(t)])[:, None] * freqs[None, :]
    emb = torch.cat([torch.cos(e), torch.sin(e)], dim=-1)           # flipped: cos first
    w = sd["time_embedding.linear_1.weight"]
    emb = emb.to(device=w.device, dtype=w.dtype)                    # unet.py:352 (t_emb.to(dtype=self.dtype))
    return _lin(sd, "time_embedding.linear_2", F.silu(_lin(sd, "time_embedding.linear_1", emb)))   # [1, 4*c0]


def resnet_block(sd, p, x, temb, groups, eps):
    """ResnetBlock3D.forward, src/model/resnet.py:174-204 (F = 1, output_scale_factor = 1)."""
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x, groups, eps)))
    if temb is not None:
        h = h + _lin(sd, p + ".time_emb_proj", F.silu(temb))[:, :, None, None]
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h, groups, eps)))
    if (p + ".conv_shortcut.weight") in sd:
        x = _conv(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def task_attention(sd, a, attn_out, feats: Dict[str, torch.Tensor], output_type: str, n_attns: int):
    """Per-pixel cross-task attention, src/model/attention.py:463-600 (eval branch: no random task masking).
    feats: {other task -> [B, N, C] child feature}.  Returns to_out_task(...) to be added to attn_out."""
    B, N, C = attn_out.shape
    keys, vals = [], []
    for t, f in feats.items():                                   # :475-497
        vals.append(_mlp(sd, f"{a}.task_to_v.{t}", _ln(sd, f"{a}.task_norm_v.{t}", f)))
        keys.append(_mlp(sd, f"{a}.task_to_k.{t}", _ln(sd, f"{a}.task_norm_k.{t}", f)))
    k = torch.stack(keys, dim=2)                                  # [B, N, T, C]   (:500-503, each pixel = batch item)
    v = torch.stack(vals, dim=2)
    q = _mlpv2(sd, f"{a}.task_to_q.{output_type}", _ln(sd, f"{a}.task_norm_q.{output_type}", attn_out))   # :512
    dh = C // n_attns
    qh = q.reshape(B, N, n_attns, 1, dh)                          # :517-519
    kh = k.reshape(B, N, -1, n_attns, dh).permute(0, 1, 3, 2, 4)  # [B,N,h,T,dh]
    vh = v.reshape(B, N, -1, n_attns, dh).permute(0, 1, 3, 2, 4)
    s = torch.matmul(qh, kh.transpose(-1, -2)) / math.sqrt(dh)    # xformers default scale, :587-592
    o = torch.matmul(torch.softmax(s, dim=-1), vh).reshape(B, N, C)   # :596-597 (heads merged back)
    return _lin(sd, f"{a}.to_out_task", o)                        # :598


def _mlp(sd, p, x):        # MLP, src/model/attention.py:655-698
    return _lin(sd, p + ".fc2", F.gelu(_lin(sd, p + ".fc1", x)))


def _mlpv2(sd, p, x):      # MLPv2 with num_hidden_layers=2, src/model/attention.py:701-751, util/model.py:126-132
    x = F.gelu(_lin(sd, p + ".net.0", x))
    x = F.gelu(_lin(sd, p + ".net.2", x))
    x = F.gelu(_lin(sd, p + ".net.4", x))
    return _lin(sd, p + ".net.6", x)


def transformer_block(sd, p, x, text, heads, groups, task_feat=None, output_type=None, n_attns=4, fp16_attn=False):
    """Transformer3DModel.forward (src/model/attention.py:174-223, use_linear_projection branch) around
    BasicTransformerBlock.forward (:323-380).  Returns (output, attn1 output = "afterSelfAttn_residual" tap)."""
    B, C, H, W = x.shape
    h = _gn(sd, p + ".norm", x, groups, 1e-6).permute(0, 2, 3, 1).reshape(B, H * W, C)
    h = _lin(sd, p + ".proj_in", h)
    t = p + ".transformer_blocks.0"
    n1 = _ln(sd, t + ".norm1", h)
    a = t + ".attn1"
    attn = _attend(_lin(sd, a + ".to_q", n1), _lin(sd, a + ".to_k", n1), _lin(sd, a + ".to_v", n1), heads, fp16_attn)
    attn = _lin(sd, a + ".to_out.0", attn)                                         # :442-460
    if task_feat is not None:
        attn = attn + task_attention(sd, a, attn, task_feat, output_type, n_attns)  # :463-600
    h = attn + h                                                                    # :347
    n2 = _ln(sd, t + ".norm2", h)
    a2 = t + ".attn2"
    x2 = _attend(_lin(sd, a2 + ".to_q", n2), _lin(sd, a2 + ".to_k", text), _lin(sd, a2 + ".to_v", text), heads)
    h = _lin(sd, a2 + ".to_out.0", x2) + h                                          # :360-364
    n3 = _ln(sd, t + ".norm3", h)
    g = _lin(sd, t + ".ff.net.0.proj", n3)
    val, gate = g.chunk(2, dim=-1)
    h = _lin(sd, t + ".ff.net.2", val * F.gelu(gate)) + h                           # :372-373 (GEGLU)
    h = _lin(sd, p + ".proj_out", h).reshape(B, H, W, C).permute(0, 3, 1, 2)
    return h + x, attn                                                              # :217


def unet_forward(sd, cfg, sample, text, task_feats: Optional[List[Dict[str, torch.Tensor]]] = None,
                 output_type: Optional[str] = None, timestep: int = 999, fp16_attn: bool = False):
    """UNet3DConditionModel.forward, src/model/unet.py:284-445, with the block wiring of
    src/model/unet_blocks.py (CrossAttnDownBlock3D :294-338, DownBlock3D :388-416, UNetMidBlock3DCrossAttn :199-214,
    UpBlock3D :591-614, CrossAttnUpBlock3D :501-541).  sample: [B, 12, h, w] (the F=1 frame axis is dropped);
    text: [B, n_tok, cross_dim].  Returns (sample [B, 4, h, w], list of per-transformer attn1 outputs)."""
    c = cfg.block_out_channels
    n = len(c)
    G, eps = cfg.norm_num_groups, cfg.norm_eps
    temb = timestep_embedding(sd, timestep, c[0]).expand(sample.shape[0], -1)
    feats = []
    li = [0]

    def tf(p, x, heads):
        tfeat = task_feats[li[0]] if task_feats is not None else None
        y, f = transformer_block(sd, p, x, text, heads, G, tfeat, output_type, cfg.n_attns, fp16_attn)
        feats.append(f)
        li[0] += 1
        return y

    x = _conv(sd, "conv_in", sample)
    skips = [x]
    for i in range(n):
        for j in range(cfg.layers_per_block):
            x = resnet_block(sd, f"down_blocks.{i}.resnets.{j}", x, temb, G, eps)
            if i < n - 1:
                x = tf(f"down_blocks.{i}.attentions.{j}", x, cfg.heads[i])
            skips.append(x)
        if i < n - 1:
            x = _conv(sd, f"down_blocks.{i}.downsamplers.0.conv", x, stride=2, padding=1)   # resnet.py:76-107
            skips.append(x)
    x = resnet_block(sd, "mid_block.resnets.0", x, temb, G, eps)
    x = tf("mid_block.attentions.0", x, cfg.heads[-1])
    x = resnet_block(sd, "mid_block.resnets.1", x, temb, G, eps)
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            x = torch.cat([x, skips.pop()], dim=1)                                          # unet_blocks.py:509,597
            x = resnet_block(sd, f"up_blocks.{i}.resnets.{j}", x, temb, G, eps)
            if i > 0:
                x = tf(f"up_blocks.{i}.attentions.{j}", x, cfg.heads[n - 1 - i])
        if i < n - 1:
            # Upsample3D (resnet.py:41-72): nearest to the next skip's size (forward_upsample_size, unet.py:312-320,
            # 415-416) which equals x2 whenever the sizes are even
            x = F.interpolate(x, size=skips[-1].shape[-2:], mode="nearest")
            x = _conv(sd, f"up_blocks.{i}.upsamplers.0.conv", x)
    x = _conv(sd, "conv_out", F.silu(_gn(sd, "conv_norm_out", x, G, eps)))                  # unet.py:438-440
    return x, feats


# ------------------------------------------------------------------------------------------------- VAE
def _vae_resnet(sd, p, x, G):
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x, G, 1e-6)))
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h, G, 1e-6)))
    if (p + ".conv_shortcut.weight") in sd:
        x = _conv(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def _vae_mid(sd, p, x, G):
    x = _vae_resnet(sd, p + ".resnets.0", x, G)
    a = p + ".attentions.0"
    B, C, H, W = x.shape
    h = _gn(sd, a + ".group_norm", x, G, 1e-6).reshape(B, C, H * W).transpose(1, 2)
    o = _attend(_lin(sd, a + ".to_q", h), _lin(sd, a + ".to_k", h), _lin(sd, a + ".to_v", h), 1)
    o = _lin(sd, a + ".to_out.0", o).transpose(1, 2).reshape(B, C, H, W)
    x = x + o
    return _vae_resnet(sd, p + ".resnets.1", x, G)


def vae_encode(sd, vcfg, rgb_norm):
    """StableMTLPipeline.encode_rgb, src/stablemtl_pipeline.py:607-624: diffusers Encoder -> quant_conv -> mean * 0.18215."""
    c, G = vcfg.block_out_channels, vcfg.norm_num_groups
    x = _conv(sd, "encoder.conv_in", rgb_norm)
    for i in range(len(c)):
        for j in range(vcfg.layers_per_block):
            x = _vae_resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", x, G)
        if i < len(c) - 1:
            x = _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", F.pad(x, (0, 1, 0, 1)), stride=2, padding=0)
    x = _vae_mid(sd, "encoder.mid_block", x, G)
    x = _conv(sd, "encoder.conv_out", F.silu(_gn(sd, "encoder.conv_norm_out", x, G, 1e-6)))
    moments = _conv(sd, "quant_conv", x, padding=0)
    mean, _ = torch.chunk(moments, 2, dim=1)
    return mean * LATENT_SCALE


def vae_decode(sd, vcfg, latent):
    """decode_output up to the decoder output, src/stablemtl_pipeline.py:639-643."""
    c, G = vcfg.block_out_channels, vcfg.norm_num_groups
    x = _conv(sd, "post_quant_conv", latent / LATENT_SCALE, padding=0)
    x = _conv(sd, "decoder.conv_in", x)
    x = _vae_mid(sd, "decoder.mid_block", x, G)
    for i in range(len(c)):
        for j in range(vcfg.layers_per_block + 1):
            x = _vae_resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", x, G)
        if i < len(c) - 1:
            x = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", F.interpolate(x, scale_factor=2.0, mode="nearest"))
    return _conv(sd, "decoder.conv_out", F.silu(_gn(sd, "decoder.conv_norm_out", x, G, 1e-6)))


def select_channels(stacked, output_type):
    """decode_output channel handling, src/stablemtl_pipeline.py:645-656."""
    if output_type in ("depth", "shading"):
        return stacked.mean(dim=1, keepdim=True)
    if output_type == "optical_flow":
        return stacked[:, :2]
    return stacked


# ------------------------------------------------------------------------------------------------- pipeline
class Oracle:
    """StableMTLPipeline restated (src/stablemtl_pipeline.py:177-370, 427-452, 475-515, 519-604), deterministic
    input noise, encode_rgb_model="duplicate", t = 999, exclude_mainstream_output_type=True."""

    def __init__(self, ucfg, vcfg, child_sd, vae_sd, text, main_sd=None, tasks=TASKS, fp16_attn=False):
        self.ucfg, self.vcfg = ucfg, vcfg
        self.child, self.main, self.vae, self.text = child_sd, main_sd, vae_sd, text
        self.tasks = list(tasks)
        self.fp16_attn = fp16_attn

    def rgb_latent(self, task, rgb_norm, rgb_next_norm, cache):
        if "rgb" not in cache:
            cache["rgb"] = vae_encode(self.vae, self.vcfg, rgb_norm)
        first = cache["rgb"]
        if task in FLOW_TASKS and rgb_next_norm is not None:            # :433-434
            if "next" not in cache:
                cache["next"] = vae_encode(self.vae, self.vcfg, rgb_next_norm)
            second = cache["next"]
        else:
            second = first                                               # :436-437 (duplicate)
        return torch.cat([first, second, torch.zeros_like(first)], dim=1)   # :447, :557-558, :582-584

    def text_for(self, task, batch):
        return self.text[task][None].expand(batch, -1, -1).to(self.vae["quant_conv.weight"].dtype)

    @torch.no_grad()
    def child_features(self, rgb_norm, rgb_next_norm, cache):
        """create_task_feats (:475-515) for every task, computed once per image (the reference recomputes them per
        main task with identical results)."""
        if "feats" not in cache:
            out = {}
            for t in self.tasks:
                x = self.rgb_latent(t, rgb_norm, rgb_next_norm, cache)
                _, f = unet_forward(self.child, self.ucfg, x, self.text_for(t, x.shape[0]), fp16_attn=self.fp16_attn)
                out[t] = f
            cache["feats"] = out
        return cache["feats"]

    @torch.no_grad()
    def single_infer(self, rgb_norm, rgb_next_norm, output_type, cache=None, return_latent=False):
        """single_infer (:519-604): returns the clipped task map [B, {1,2,3}, H, W]."""
        cache = {} if cache is None else cache
        x = self.rgb_latent(output_type, rgb_norm, rgb_next_norm, cache)
        text = self.text_for(output_type, x.shape[0])
        if self.main is not None:
            feats = self.child_features(rgb_norm, rgb_next_norm, cache)
            others = [t for t in self.tasks if t != output_type]           # :483-484
            nl = len(feats[others[0]])
            task_feats = [{t: feats[t][l] for t in others} for l in range(nl)]
            lat, _ = unet_forward(self.main, self.ucfg, x, text, task_feats, output_type, fp16_attn=self.fp16_attn)
        else:
            lat, _ = unet_forward(self.child, self.ucfg, x, text, fp16_attn=self.fp16_attn)
        out = torch.clip(select_channels(vae_decode(self.vae, self.vcfg, lat), output_type), -1.0, 1.0)   # :599-601
        return (out, lat) if return_latent else out

    @torch.no_grad()
    def predict_all(self, rgb, rgb_next=None, return_latents=False):
        """__call__ (:177-370) for every task on one batch: {task: post-processed map}, semantic -> int64 class ids."""
        rgb_norm = rgb / 255.0 * 2.0 - 1.0                                  # :263
        next_norm = None if rgb_next is None else rgb_next / 255.0 * 2.0 - 1.0
        cache, maps, clipped, latents = {}, {}, {}, {}
        for t in self.tasks:
            out, lat = self.single_infer(rgb_norm, next_norm, t, cache, return_latent=True)
            clipped[t], latents[t] = out, lat
            maps[t] = postprocess(out, t)
        return (maps, clipped, latents) if return_latents else maps


def postprocess(out, task):
    """Per-type post-processing of __call__, src/stablemtl_pipeline.py:297-366, kept batched ([B, C, H, W])."""
    if task in ("depth", "shading", "albedo"):
        return (out + 1.0) / 2.0
    if task == "normal":
        nrm = out.norm(dim=1, keepdim=True)
        nrm = torch.where(nrm == 0, torch.ones_like(nrm), nrm)
        return out / nrm
    if task == "semantic":
        pal = torch.tensor(PALETTE, dtype=out.dtype, device=out.device) / 255.0 * 2.0 - 1.0
        B, _, H, W = out.shape
        d = torch.cdist(out.permute(0, 2, 3, 1).reshape(-1, 3), pal)
        return torch.argmin(d, dim=1).reshape(B, H, W)
    return out
