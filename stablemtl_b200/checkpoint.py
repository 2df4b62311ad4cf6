"""Checkpoint ingest: the files the reference evaluates from -> the state dicts `StableMTLEngine` takes.

Mirrors the load sequence of the reference (paths under /root/reference):
  eval_mtl.py:288-300           SD-2 diffusers directory: `<base_ckpt_dir>/stable-diffusion-2/{unet,vae}/diffusion_pytorch_model.*`
  src/util/model.py:196-199     the 4-channel SD-2 `conv_in` is widened to 12 channels: weights tiled x3, scaled 1/3
  src/util/model.py:205-221     `single_stream_unet.pth` -> child UNet (and main unless main_stream_from_scratch)
  src/trainer/stablemtl_trainer.py:1176-1181   `<run>/checkpoint/latest/unet/diffusion_pytorch_model.bin` -> main UNet incl. task modules
Only tensors are read (torch.load with weights_only=True, or safetensors); nothing here needs diffusers.
"""
import os
from typing import Dict, Optional

import torch

UNET_FILE_CANDIDATES = ("diffusion_pytorch_model.bin", "diffusion_pytorch_model.safetensors",
                        "diffusion_pytorch_model.fp16.safetensors")
# diffusers < 0.20 AutoencoderKL attention names
_LEGACY_VAE_ATTN = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}


def load_tensor_file(path: str) -> Dict[str, torch.Tensor]:
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    sd = torch.load(path, map_location="cpu", weights_only=True)
    if not isinstance(sd, dict) or not all(torch.is_tensor(v) for v in sd.values()):
        raise ValueError(f"{path} is not a flat tensor state dict")
    return sd


def _find(dirname: str) -> str:
    for f in UNET_FILE_CANDIDATES:
        p = os.path.join(dirname, f)
        if os.path.exists(p):
            return p
    raise FileNotFoundError(f"no diffusion_pytorch_model.* under {dirname}")


def widen_conv_in(sd: Dict[str, torch.Tensor], repeat: int = 3) -> Dict[str, torch.Tensor]:
    """`_replace_unet_conv_in` (src/util/model.py:11-27): [320, 4, 3, 3] -> [320, 4*repeat, 3, 3], tiled and scaled by
    1/repeat.  A state dict whose conv_in already has 4*repeat input channels is returned unchanged."""
    w = sd["conv_in.weight"]
    if w.shape[1] == 4 * repeat:
        return sd
    if w.shape[1] != 4:
        raise ValueError(f"conv_in.weight has {w.shape[1]} input channels, expected 4 or {4 * repeat}")
    out = dict(sd)
    out["conv_in.weight"] = w.repeat(1, repeat, 1, 1) * (1.0 / repeat)
    return out


def normalize_vae_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """legacy `mid_block.attentions.0.{query,key,value,proj_attn}` -> `{to_q,to_k,to_v,to_out.0}`; 1x1-conv shaped
    attention weights [C, C, 1, 1] -> [C, C]"""
    out = {}
    for k, v in sd.items():
        parts = k.split(".")
        if "attentions" in parts:
            i = parts.index("attentions")
            if len(parts) > i + 2 and parts[i + 2] in _LEGACY_VAE_ATTN:
                parts[i + 2] = _LEGACY_VAE_ATTN[parts[i + 2]]
                k = ".".join(parts)
            if v.dim() == 4 and k.endswith("weight") and v.shape[2:] == (1, 1) and "attentions" in k:
                v = v.reshape(v.shape[0], v.shape[1])
        out[k] = v
    return out


def load_reference_checkpoints(base_ckpt_dir: str, run_dir: Optional[str] = None,
                               single_stream_path: Optional[str] = None, sd2_name: str = "stable-diffusion-2",
                               main_stream_from_scratch: bool = False, repeat_input: int = 3):
    """Returns (child_sd, main_sd_or_None, vae_sd) exactly as `setup_unet` + `load_checkpoint` would populate
    `model.unet_child`, `model.unet`, `model.vae`.

    single-stream (StableMTL-S): no `single_stream_path` -> the UNet of `run_dir` (or SD-2) is the only UNet.
    multi-stream: child = `single_stream_path`; main = `run_dir` checkpoint (it contains the task modules)."""
    sd2 = os.path.join(base_ckpt_dir, sd2_name)
    vae_sd = normalize_vae_keys(load_tensor_file(_find(os.path.join(sd2, "vae"))))
    base_unet = widen_conv_in(load_tensor_file(_find(os.path.join(sd2, "unet"))), repeat_input)
    run_unet = None
    if run_dir is not None:
        run_unet = load_tensor_file(os.path.join(run_dir, "checkpoint", "latest", "unet", "diffusion_pytorch_model.bin"))
    if single_stream_path is None:
        return (run_unet if run_unet is not None else base_unet), None, vae_sd
    child = widen_conv_in(load_tensor_file(single_stream_path), repeat_input)
    if run_unet is not None:
        main = run_unet
    else:
        main = dict(base_unet if main_stream_from_scratch else child)
    if not any(".task_to_q." in k for k in main):
        raise ValueError("the main UNet checkpoint has no task modules (task_to_q/k/v, to_out_task): multi-stream needs "
                         "a trained run checkpoint (src/util/model.py:102-146 zero-initialises them)")
    return child, main, vae_sd
