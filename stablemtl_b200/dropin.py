"""Drop-in facade: the reference `StableMTLPipeline`'s object surface on top of the B200 engine.

What the reference's callers touch and what answers it here:

    eval_mtl.py / StableMTLTrainer.eval (src/trainer/stablemtl_trainer.py:405-437)
        model.vae.to(device); model.text_encoder.to(device); accelerator.prepare(model.unet[, model.unet_child]);
        model.unet.eval()                                  -> light proxy objects (`_ModuleProxy` and subclasses)
    validate_single_dataset (:697-712)                     -> `StableMTLPipeline.__call__`
    the training loop's feature extraction (:262-305)      -> `.encode_rgb_latent`, `.create_text_condition`,
        `.create_task_feats`, `.unet(...) -> (obj.sample, feats)`, `.encode_rgb`, `.decode_output`

All numerics run in the sm_100a kernels through `StableMTLEngine`; torch is used here for layout changes of the
arguments and results only.  Inference only: the proxies have no parameters and nothing here builds an autograd graph.
"""
import random
from typing import Dict, List, Optional

import numpy as np
import torch

from .engine import F32
from .pipeline import PALETTE, StableMTLEngine
from .synth import FLOW_TASKS, SD2_UNET, SD2_VAE

KNOWN_OUTPUT_TYPES = ("optical_flow", "scene_flow", "depth", "normal", "semantic", "albedo", "shading")


class _Out(dict):
    """Attribute/dict output object standing in for diffusers.utils.BaseOutput (stablemtl_pipeline.py:32-109)."""
    __getattr__ = dict.__getitem__


class _ModuleProxy:
    """What the trainer does to a sub-module of the pipeline before evaluating: .to(), .eval(), accelerator.prepare()
    (a non-nn.Module passes through `prepare` unchanged), .parameters() for the optimizer it builds even for eval."""

    def __init__(self, engine: StableMTLEngine):
        self._engine = engine
        self.training = False

    @property
    def device(self):
        return self._engine.device

    dtype = torch.float32

    def to(self, *args, **kwargs):
        return self

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        if mode:
            raise RuntimeError("stablemtl_b200 is the inference path: the accelerated modules cannot be put in train mode")
        return self

    def requires_grad_(self, flag=False):
        return self

    def parameters(self, recurse=True):
        return iter(())

    def state_dict(self):
        raise RuntimeError("the accelerated modules hold packed 16-bit weights; keep the reference state dicts for saving")


class _UNetProxy(_ModuleProxy):
    """`unet(sample, timestep, encoder_hidden_states, task_feats=..., output_type=...) -> (obj.sample, feats)`
    (src/model/unet.py:284-294,445; positional form of the child call: stablemtl_pipeline.py:508-510)."""

    def __init__(self, engine, which):
        super().__init__(engine)
        self._which = which                # "single" | "child" | "main"

    def __call__(self, sample, timestep, encoder_hidden_states, class_labels=None, attention_mask=None,
                 return_dict=True, task_feats=None, output_type=None):
        t = torch.as_tensor(timestep).reshape(-1)
        if not bool((t == 999).all()):
            raise ValueError("the accelerated UNet is folded at t = 999 (stablemtl_pipeline.py:552); got timestep "
                             f"{t.tolist()}")
        eng = self._engine
        task = eng.task_of_text(encoder_hidden_states)
        if output_type is not None and output_type != eng.tasks[task]:
            raise ValueError(f"output_type {output_type!r} does not match the prompt embedding ({eng.tasks[task]!r})")
        which = self._which
        if which == "main" and task_feats is None:
            raise ValueError("the multi-stream UNet needs task_feats (create_task_feats)")
        out, taps = eng.unet_forward(which, sample, task, task_feats if which == "main" else None)
        if not return_dict:
            return (out,)
        n_layers = len(eng.ucfg.transformer_dims())
        return _Out(sample=out), (taps if taps is not None else [None] * n_layers)

    forward = __call__


class _VAEProxy(_ModuleProxy):
    pass


class _TextEncoderProxy(_ModuleProxy):
    """stands in for CLIPTextModel when the 7 constant prompts come from a text cache (evaluate.save_text_cache)"""


class _Scheduler:
    class _Cfg(dict):
        __getattr__ = dict.__getitem__

    def __init__(self):
        self.config = self._Cfg(prediction_type="sample")      # eval_mtl.py:295-298; never stepped (single pass, t = 999)

    def set_timesteps(self, *a, **k):
        return None


class StableMTLPipeline:
    """Call-compatible with the reference `StableMTLPipeline` (src/stablemtl_pipeline.py:112-658).

    `__call__` / `single_infer`: the first call for an image (pair) computes all task maps with the batched engine and
    caches them; the following calls of the evaluation loop (one per `output_type`, stablemtl_trainer.py:697-712) are
    served from that result -- the deduplication the reference leaves on the table.  The returned
    tensors are clones; the engine's own `predict` returns views of its static plan buffers.

    text_encoder / tokenizer: pass the reference's CLIP objects to keep `encode_text` live; without them the prompts
    are the engine's cached embeddings (one per task) and `create_text_condition` serves those."""

    rgb_latent_scale_factor = 0.18215
    latent_scale_factor = 0.18215

    def __init__(self, engine: StableMTLEngine, input_noise="deterministic", encode_rgb_model="duplicate",
                 text_encoder=None, tokenizer=None, scheduler=None):
        if input_noise != "deterministic":
            raise ValueError("the accelerated path implements input_noise='deterministic' (config/train_base_config.yaml:23)")
        if encode_rgb_model not in ("duplicate", "zero"):
            raise ValueError("encode_rgb_model must be 'duplicate' or 'zero': 'avg' changes the UNet input width "
                             "(stablemtl_pipeline.py:438-447), which the packed conv_in does not cover")
        self.engine = engine
        self.input_noise, self.encode_rgb_model = input_noise, encode_rgb_model
        self.vae = _VAEProxy(engine)
        self.unet = _UNetProxy(engine, "main" if engine.multi else "single")
        self.unet_child = _UNetProxy(engine, "child") if engine.multi else None
        self.text_encoder = text_encoder if text_encoder is not None else _TextEncoderProxy(engine)
        self.tokenizer = tokenizer
        self.scheduler = scheduler if scheduler is not None else _Scheduler()
        self._cache = None

    # ------------------------------------------------------------------------------------------ construction
    @classmethod
    def from_reference(cls, model, output_types, device="cuda", ucfg=SD2_UNET, vcfg=SD2_VAE, text=None, **kw):
        """From the reference's loaded pipeline object (after setup_unet + load_checkpoint, eval_mtl.py:289-337): takes
        its state dicts and embeds the constant prompts once with its own CLIP (stablemtl_pipeline.py:395-408,464-472)."""
        tasks = list(output_types)
        if text is None:
            text = {t: model.encode_text([t.replace("_", " ")])[0].detach().float().cpu() for t in tasks}
        multi = getattr(model, "unet_child", None) is not None
        engine = StableMTLEngine(ucfg, vcfg, child_sd=(model.unet_child if multi else model.unet).state_dict(),
                                 vae_sd=model.vae.state_dict(), text=text,
                                 main_sd=model.unet.state_dict() if multi else None, tasks=tasks, device=device)
        return cls(engine, input_noise=getattr(model, "input_noise", "deterministic"),
                   encode_rgb_model=getattr(model, "encode_rgb_model", "duplicate"),
                   text_encoder=getattr(model, "text_encoder", None), tokenizer=getattr(model, "tokenizer", None),
                   scheduler=getattr(model, "scheduler", None), **kw)

    @classmethod
    def from_pretrained(cls, base_ckpt_dir, text, output_types, run_dir=None, single_stream_path=None, device="cuda",
                        ucfg=SD2_UNET, vcfg=SD2_VAE, **kw):
        """From the reference's on-disk layout (eval_mtl.py:288-293, stablemtl_trainer.py:1176-1181) plus a text cache
        {task: [n_tok, 1024]} (evaluate.load_text_cache): no reference code, diffusers or CLIP needed at run time."""
        from .checkpoint import load_reference_checkpoints
        child, main, vae = load_reference_checkpoints(base_ckpt_dir, run_dir, single_stream_path)
        engine = StableMTLEngine(ucfg, vcfg, child, vae, text, main, tasks=list(output_types), device=device)
        return cls(engine, **kw)

    @property
    def device(self):
        return self.engine.device

    def to(self, *a, **k):
        return self

    def set_progress_bar_config(self, **k):
        return None

    # ------------------------------------------------------------------------------------------ stage methods
    def encode_text(self, prompt):
        """stablemtl_pipeline.py:395-408"""
        if self.tokenizer is not None and not isinstance(self.text_encoder, _ModuleProxy):
            ids = self.tokenizer(prompt, padding="longest", max_length=self.tokenizer.model_max_length, truncation=True,
                                 return_tensors="pt").input_ids.to(self.text_encoder.device)
            return self.text_encoder(ids)[0]
        prompts = [prompt] if isinstance(prompt, str) else list(prompt)
        embs = []
        for pr in prompts:
            t = pr.replace(" ", "_")
            if t not in self.engine.tasks:
                raise ValueError(f"no cached embedding for prompt {pr!r}: pass the reference's text_encoder/tokenizer, "
                                 f"or use one of the task prompts ({', '.join(self.engine.tasks)})")
            i = self.engine.tasks.index(t)
            embs.append(self.engine.child_w.text[i, : self.engine.child_w.ntok[i]])
        if len({e.shape[0] for e in embs}) != 1:
            raise ValueError("prompts of different token counts need the real tokenizer's 'longest' padding")
        return torch.stack(embs).to(self.device)

    def create_text_condition(self, output_types, batch_size):
        """stablemtl_pipeline.py:464-472 -> [(batch n_task), n_tok, 1024]"""
        text_embed = self.encode_text([t.replace("_", " ") for t in output_types]).detach().clone().to(self.device)
        text_embed = text_embed.unsqueeze(0).expand(batch_size, -1, -1, -1)
        return text_embed.reshape(batch_size * text_embed.shape[1], *text_embed.shape[2:])

    def encode_rgb(self, rgb_in: torch.Tensor) -> torch.Tensor:
        """stablemtl_pipeline.py:607-624"""
        return self.engine.encode_rgb(rgb_in)

    def encode_rgb_latent(self, output_type, rgb_norm, rgb_next_norm):
        """stablemtl_pipeline.py:427-452 -> [B, 8, 1, h, w]"""
        assert output_type in KNOWN_OUTPUT_TYPES, f"Unknown output type: {output_type}"
        rgb_in = self.encode_rgb(rgb_norm)
        if output_type in FLOW_TASKS and rgb_next_norm is not None:
            rgb_next_in = self.encode_rgb(rgb_next_norm)
        elif self.encode_rgb_model == "duplicate":
            rgb_next_in = rgb_in
        else:
            rgb_next_in = torch.zeros_like(rgb_in)
        return torch.cat([rgb_in, rgb_next_in], dim=1).unsqueeze(2)

    def create_task_feats(self, rgb_norm, rgb_next_norm, timesteps, output_type, task_output_types, rand_num_generator,
                          exclude_mainstream_output_type, drop_ratio=0.0):
        """stablemtl_pipeline.py:475-515 -> (list of child samples [B,4,1,h,w], list[16] of {task: [B, N_l, C_l]})"""
        if self.unet_child is None:
            return None, None
        if exclude_mainstream_output_type:
            task_output_types = [t for t in task_output_types if t != output_type]
        if drop_ratio > 0.0 and random.random() < drop_ratio:
            task_output_types = list(np.random.choice(task_output_types, size=len(task_output_types) - 1, replace=False))
        batch_size = rgb_norm.shape[0]
        n_layers = len(self.engine.ucfg.transformer_dims())
        list_task_feats = [{} for _ in range(n_layers)]
        outs = []
        for t in task_output_types:
            text = self.create_text_condition([t], batch_size)
            rgb_latent = self.encode_rgb_latent(t, rgb_norm=rgb_norm, rgb_next_norm=rgb_next_norm)
            cat_latents = torch.cat([rgb_latent, torch.zeros_like(rgb_latent[:, :4])], dim=1)
            out, feats = self.unet_child(cat_latents, timesteps, text)
            outs.append(out.sample)
            for i, f in enumerate(feats):
                list_task_feats[i][t] = f
        return outs, list_task_feats

    def decode_output(self, latent: torch.Tensor, output_type: str) -> torch.Tensor:
        """stablemtl_pipeline.py:626-656 (not clipped)"""
        stacked = self.engine.decode_latents(latent)
        if output_type in ("depth", "shading"):
            return stacked.mean(dim=1, keepdim=True)
        if output_type in ("normal", "semantic", "rgb", "scene_flow", "albedo"):
            return stacked
        if output_type == "optical_flow":
            return stacked[:, :2]
        raise ValueError(f"Unknown output type: {output_type}")

    # ------------------------------------------------------------------------------------------ the evaluation call
    def _all_tasks(self, rgb_norm, rgb_next_norm):
        dev = self.device
        a = rgb_norm.to(dev)
        b = None if rgb_next_norm is None else rgb_next_norm.to(dev)
        c = self._cache
        hit = (c is not None and c["rgb"].shape == a.shape and torch.equal(c["rgb"], a) and
               ((c["next"] is None) == (b is None)) and (b is None or torch.equal(c["next"], b)))
        if not hit:
            # the engine takes [0,255]; undo (x/255*2-1) of stablemtl_pipeline.py:263
            rgb = (a.to(F32) + 1.0) / 2.0 * 255.0
            nxt = None if b is None else (b.to(F32) + 1.0) / 2.0 * 255.0
            self.engine.predict(rgb, nxt)
            self._cache = {"rgb": a.clone(), "next": None if b is None else b.clone(),
                           "maps": {t: v.clone() for t, v in self.engine.last.items()}}
        return self._cache["maps"]

    @torch.no_grad()
    def single_infer(self, rgb_norm, num_inference_steps, generator, show_pbar, output_type,
                     exclude_mainstream_output_type, rgb_next_norm=None, task_output_types=[]):
        """stablemtl_pipeline.py:519-604 -> clipped map [B, {1,2,3}, H, W]"""
        if output_type not in self.engine.tasks:
            raise ValueError(f"Unknown output type: {output_type}")
        if self.engine.multi:
            if not exclude_mainstream_output_type:
                raise ValueError("the multi-stream engine is built for exclude_mainstream_output_type=True "
                                 "(config/train_stablemtl.yaml:22)")
            if task_output_types and set(task_output_types) != set(self.engine.tasks):
                raise ValueError(f"task_output_types {list(task_output_types)} differ from the task streams the engine was "
                                 f"built with {self.engine.tasks}: the child streams that attend are fixed at load time")
        return self._all_tasks(rgb_norm, rgb_next_norm)[output_type]

    @torch.no_grad()
    def __call__(self, input_image, exclude_mainstream_output_type, next_input_image=None, denoising_steps=None,
                 ensemble_size=5, processing_res=None, match_input_res=True, resample_method="bilinear", batch_size=0,
                 generator=None, color_map="Spectral", show_progress_bar=True, ensemble_kwargs=None,
                 output_type="depth", task_output_types=[]):
        """stablemtl_pipeline.py:177-370"""
        if processing_res:
            raise ValueError("processing_res > 0 (resize) is outside the accelerated path; eval uses processing_res=0 "
                             "(config/train_base_config.yaml:179-180)")
        rgb, nxt = input_image, next_input_image
        assert rgb.min() >= 0 and rgb.max() <= 255, "Input images should be in [0,255] range"
        rgb_norm = rgb / 255.0 * 2.0 - 1.0
        nxt_norm = None if nxt is None else nxt / 255.0 * 2.0 - 1.0
        out = self.single_infer(rgb_norm, denoising_steps, generator, show_progress_bar, output_type,
                                exclude_mainstream_output_type, nxt_norm, task_output_types)
        pred = out.squeeze().cpu().numpy()                                       # stablemtl_pipeline.py:294-295
        if output_type == "albedo":
            return _Out(albedo_np=(pred + 1.0) / 2.0)
        if output_type == "shading":
            return _Out(shading_np=(pred + 1.0) / 2.0)
        if output_type == "depth":
            return _Out(depth_np=(pred + 1.0) / 2.0, depth_colored=None)
        if output_type == "normal":
            n = np.linalg.norm(pred, axis=0, keepdims=True)
            n[n == 0] = 1.0
            return _Out(normal_np=pred / n, normal_colored=None)
        if output_type == "optical_flow":
            return _Out(optical_flow_np=pred)
        if output_type == "scene_flow":
            return _Out(scene_flow_np=pred)
        if output_type == "semantic":
            pal = np.asarray(PALETTE, dtype=np.float32)
            emb = pal / 255.0 * 2.0 - 1.0
            flat = pred.transpose(1, 2, 0).reshape(-1, 3)
            d = np.sqrt(((flat[:, None, :] - emb[None]) ** 2).sum(-1))
            return _Out(semantic_class_id=d.argmin(1).reshape(pred.shape[1:]), class_color_visualizes=pal)
        raise ValueError(f"Unknown output type: {output_type}")
