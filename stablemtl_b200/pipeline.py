"""Host-side mirror of the reference pipeline for the accelerated path.

`StableMTLEngine` is the batched all-task fast path (deduplicated schedule of SURVEY.md §3.1: 1-2 VAE encodes,
one batched child pass over the 7 task streams, one batched main pass, 7 decodes).
`dropin.StableMTLPipeline` keeps the reference's object surface (`src/stablemtl_pipeline.py:112-658`) on top of
it so `eval_mtl.py` / `StableMTLTrainer` (`src/trainer/stablemtl_trainer.py:405-437, 697-712`) can use it unchanged.  All numerics run in the sm_100a kernels; torch is used for device memory, H2D/D2H copies
and the current stream only.  There is no CPU / PyTorch fallback: constructing either class without a CUDA
device raises.
"""
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib as L
from . import ops
from . import stream_shard as SS
from .engine import F32, Pool, UNetPlan, UNetWeights, VAEDecodePlan, VAEEncodePlan, VAEWeights
from .synth import FLOW_TASKS, TASKS, UNetConfig, VAEConfig

# VKitti2Encoder(n_classes=8).class_color_embeddings (src/dataset/semantic/encoding.py:10-35,92-96; labels.py:42-55)
PALETTE = [[128, 64, 128], [70, 70, 70], [153, 153, 153], [250, 170, 30], [220, 220, 0], [107, 142, 35],
           [70, 130, 180], [0, 0, 142]]
TASK_MODE = {"depth": L.MAP_MEAN1, "shading": L.MAP_MEAN1, "albedo": L.MAP_RGB3, "normal": L.MAP_NORMAL,
             "optical_flow": L.MAP_FLOW2, "scene_flow": L.MAP_FLOW3, "semantic": L.MAP_SEMANTIC}
TASK_CH = {"depth": 1, "shading": 1, "albedo": 3, "normal": 3, "optical_flow": 2, "scene_flow": 3, "semantic": 3}


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise RuntimeError("stablemtl_b200 needs a CUDA (sm_100a) device: there is no CPU fallback")
    return torch.device(device)


class StableMTLEngine:
    def __init__(self, ucfg: UNetConfig, vcfg: VAEConfig, child_sd, vae_sd, text: Dict[str, torch.Tensor],
                 main_sd=None, tasks: List[str] = TASKS, device="cuda", max_decode_batch=16, use_graph=True,
                 stream_shard=False, stream_group=None):
        """stream_shard=True (multi-stream only, torch.distributed initialised): this rank owns a block of the task
        streams and exchanges the child features with the other ranks between the child and the main pass
        (stream_shard.py) -- the low-latency sharding for batches smaller than the number of GPUs."""
        self.device = _require_cuda(device)
        self.use_graph = use_graph
        self.ucfg, self.vcfg, self.tasks = ucfg, vcfg, list(tasks)
        self.multi = main_sd is not None
        self.stream = bool(stream_shard)
        self.sgroup = stream_group
        self.my = list(range(len(self.tasks)))                       # task indices this rank computes
        if self.stream:
            import torch.distributed as dist
            if not self.multi:
                raise ValueError("stream sharding exchanges child features: it needs the multi-stream model")
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("stream_shard=True needs an initialised torch.distributed process group")
            self.sworld, self.srank = dist.get_world_size(stream_group), dist.get_rank(stream_group)
            if self.sworld * SS.slots_per_rank(len(self.tasks), self.sworld) > L.MAX_TASKS:
                raise ValueError(f"{len(self.tasks)} task streams over {self.sworld} ranks need more than "
                                 f"{L.MAX_TASKS} exchange slots (SMTL_MAX_TASKS)")
            lo, hi = SS.task_range(len(self.tasks), self.sworld, self.srank)
            self.my = list(range(lo, hi))
        self.child_w = UNetWeights(child_sd, ucfg, text, self.tasks, self.device)
        self.main_w = UNetWeights(main_sd, ucfg, text, self.tasks, self.device) if self.multi else None
        self.vae_w = VAEWeights(vae_sd, vcfg, self.device)
        self.palette = (torch.tensor(PALETTE, dtype=F32) / 255.0 * 2.0 - 1.0).contiguous().to(self.device)
        self.max_decode_batch = max_decode_batch
        self._plans = {}

    # ------------------------------------------------------------------------------------------ plan construction
    def _build(self, B, H, W, with_next, rgb_dtype=F32):
        dev = self.device
        pool = Pool(dev)
        T = len(self.tasks)
        mine = self.my
        Tm = len(mine)
        n_enc = 2 * B if with_next else B
        xchg = None
        if self.stream:
            # exchange buffers of the 16 tap layers: send [n_max * B * N_l, C_l] (my tasks first), receive slot-major
            h, w = H // 8, W // 8
            n_max = SS.slots_per_rank(T, self.sworld)
            shapes = UNetPlan.tap_shapes(self.ucfg, h, w)
            xchg = dict(send=[torch.zeros(n_max * B * n, c, device=dev, dtype=ops.h16()) for n, c in shapes],
                        recv=[torch.zeros(self.sworld * n_max * B * n, c, device=dev, dtype=ops.h16()) for n, c in shapes],
                        slots=SS.task_slots(T, self.sworld))
            if Tm == 0:                                               # more ranks than tasks: exchange only
                return dict(enc=None, xchg=xchg, unets=[], chunks=[], out={}, launches=0, flops=0, pool=pool, h=h, w=w,
                            hw=h * w)
        enc = VAEEncodePlan(self.vae_w, n_enc, H, W, pool=pool, rgb_dtype=rgb_dtype)
        h, w = enc.h, enc.w
        hw = h * w
        first = torch.arange(B, dtype=torch.int32).repeat(Tm)
        second = first.clone()
        if with_next:
            for gi, ti in enumerate(mine):
                if self.tasks[ti] in FLOW_TASKS:                     # stablemtl_pipeline.py:433-434
                    second[gi * B:(gi + 1) * B] += B
        first, second = first.to(dev), second.to(dev)
        x_in = torch.empty(Tm * B * hw, self.ucfg.in_channels, device=dev, dtype=F32)
        assemble = ops.unet_input(enc.out, first, second, hw, x_in)
        groups = list(mine)
        if self.multi and self.stream:
            child = UNetPlan(self.child_w, B, h, w, groups, mode="child", pool=pool, x_in=x_in, feat_bufs=xchg["send"])
            main = UNetPlan(self.main_w, B, h, w, groups, mode="main", feats=xchg["recv"], src_tasks=xchg["slots"],
                            pool=pool, x_in=x_in)
            unets = [child, main]
            lat = main.out
        elif self.multi:
            child = UNetPlan(self.child_w, B, h, w, groups, mode="child", pool=pool, x_in=x_in)
            main = UNetPlan(self.main_w, B, h, w, groups, mode="main", feats=child.feats_out, src_tasks=groups,
                            pool=pool, x_in=x_in)
            unets = [child, main]
            lat = main.out
        else:
            single = UNetPlan(self.child_w, B, h, w, groups, mode="single", pool=pool, x_in=x_in)
            unets = [single]
            lat = single.out
        n_lat = Tm * B
        bd = max(d for d in range(1, min(self.max_decode_batch, n_lat) + 1) if n_lat % d == 0)
        dec = VAEDecodePlan(self.vae_w, bd, h, w, pool=pool)
        HW = H * W
        out = {}
        for ti in mine:
            t = self.tasks[ti]
            ch = TASK_CH[t]
            out[t] = {"clipped": torch.empty(B, ch, H, W, device=dev, dtype=F32)}
            if t == "semantic":
                out[t]["post"] = torch.empty(B, H, W, device=dev, dtype=torch.int64)
            else:
                out[t]["post"] = torch.empty(B, ch, H, W, device=dev, dtype=F32)
        chunks = []
        for c0 in range(0, n_lat, bd):
            maps = []
            i = c0
            while i < c0 + bd:                                       # split the chunk at task boundaries
                gi, img = divmod(i, B)
                n = min(B - img, c0 + bd - i)
                t = self.tasks[mine[gi]]
                x = dec.out[(i - c0) * HW:(i - c0 + n) * HW]
                o = out[t]
                if t == "semantic":
                    maps.append(ops.task_map(x, n, HW, TASK_MODE[t], out_clipped=o["clipped"][img:img + n],
                                             out_ids=o["post"][img:img + n], palette=self.palette))
                else:
                    maps.append(ops.task_map(x, n, HW, TASK_MODE[t], out_clipped=o["clipped"][img:img + n],
                                             out_post=o["post"][img:img + n]))
                i += n
            chunks.append((c0, maps))
        launches = enc.plan.launches + 1 + sum(u.plan.launches for u in unets) + \
            len(chunks) * dec.plan.launches + sum(len(m) for _, m in chunks)
        flops = enc.plan.flops + sum(u.plan.flops for u in unets) + len(chunks) * dec.plan.flops
        return dict(enc=enc, assemble=assemble, unets=unets, lat=lat, dec=dec, bd=bd, chunks=chunks, out=out, hw=hw,
                    pool=pool, launches=launches, flops=flops, h=h, w=w, xchg=xchg)

    def plan_for(self, B, H, W, with_next=True, rgb_dtype=F32):
        key = (B, H, W, with_next, rgb_dtype)
        if key not in self._plans:
            self._plans[key] = self._build(B, H, W, with_next, rgb_dtype)
        return self._plans[key]

    # ------------------------------------------------------------------------------------------ execution
    @torch.no_grad()
    def predict(self, rgb: torch.Tensor, rgb_next: Optional[torch.Tensor] = None, return_latents=False, gather=False):
        """rgb / rgb_next: float [B,3,H,W] in [0,255] (host or device).  Returns {task: map} with the reference's
        post-processing (stablemtl_pipeline.py:297-366): depth/shading/albedo in [0,1], unit normals, flows in
        [-1,1], semantic class ids (int64 [B,H,W]).  `.last` keeps the clipped single_infer() tensors.
        With stream sharding the dict holds this rank's tasks, or every task after a broadcast when gather=True."""
        B, _, H, W = rgb.shape
        dt = torch.uint8 if rgb.dtype == torch.uint8 else F32        # uint8 images are converted inside the kernel
        p = self.plan_for(B, H, W, rgb_next is not None, dt)
        enc = p["enc"]
        if enc is not None:
            enc.rgb[:B].copy_(rgb, non_blocking=True)                 # H2D (or D2D) copy; dtype cast only if needed
            if rgb_next is not None:
                enc.rgb[B:].copy_(rgb_next, non_blocking=True)
        # The whole pass (~2000 launches over static buffers) is replayed as CUDA graphs: the first call of a plan runs
        # eagerly (it also sets the kernels' shared-memory attributes), the second captures, later ones replay.  With
        # stream sharding there are two graphs, one either side of the feature exchange.
        stages = self._stages(p)
        if enc is None:                            # more ranks than task streams: this rank only joins the collectives
            self._exchange(p)
        elif not self.use_graph or not p.get("warm"):
            for i, st in enumerate(stages):
                if i:
                    self._exchange(p)
                st()
            p["warm"] = True
        else:
            if p.get("graphs") is None:
                torch.cuda.synchronize()
                graphs = []
                for st in stages:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        st()
                    graphs.append(g)
                p["graphs"] = graphs
            for i, g in enumerate(p["graphs"]):
                if i:
                    self._exchange(p)
                g.replay()
        self.last = {t: v["clipped"] for t, v in p["out"].items()}
        res = {t: v["post"] for t, v in p["out"].items()}
        if self.stream and gather:
            like = {}
            for t in self.tasks:
                like[t] = ((B, H, W), torch.int64) if t == "semantic" else ((B, TASK_CH[t], H, W), F32)
            res = SS.gather_task_maps(res, self.tasks, like, self.device, self.sgroup)
        if return_latents:
            lat = p["lat"]
            lats = lat.view(len(self.my), B, p["h"], p["w"], -1).permute(0, 1, 4, 2, 3)
            return res, {self.tasks[ti]: lats[i] for i, ti in enumerate(self.my)}
        return res

    def empty_result(self, H, W):
        """Zero-image result with the shapes/dtypes `predict` returns (a rank whose shard is empty, shard.py)."""
        out = {}
        for t in self.tasks:
            if t == "semantic":
                out[t] = torch.empty(0, H, W, device=self.device, dtype=torch.int64)
            else:
                out[t] = torch.empty(0, TASK_CH[t], H, W, device=self.device, dtype=F32)
        return out

    def _exchange(self, p):
        """child features of every stream -> every rank (the one data-path collective, stream_shard.py)"""
        SS.exchange_taps(p["xchg"]["send"], p["xchg"]["recv"], self.sgroup)

    def _stages(self, p):
        """launch closures; more than one only when a feature exchange sits between the child and the main pass"""
        def decode():
            dec, bd, hw, lat = p["dec"], p["bd"], p["hw"], p["lat"]
            for c0, maps in p["chunks"]:
                dec.latent.copy_(lat[c0 * hw:(c0 + bd) * hw])
                dec.run()
                for m in maps:
                    m.run()

        if not self.stream:
            def whole():
                p["enc"].run()
                p["assemble"].run()
                for u in p["unets"]:
                    u.run()
                decode()
            return [whole]
        if p["enc"] is None:                                           # no task of mine: nothing to launch
            return []

        def before():
            p["enc"].run()
            p["assemble"].run()
            p["unets"][0].run()

        def after():
            p["unets"][1].run()
            decode()
        return [before, after]

    def _launch_all(self, p):
        for i, st in enumerate(self._stages(p)):
            if i:
                self._exchange(p)
            st()

    # ------------------------------------------------------------------------------------------ stage-level entry points
    # The reference pipeline's own building blocks (encode_rgb, decode_output, unet(...), create_task_feats), for callers
    # that drive the stages themselves (stablemtl_trainer.py:262-305); `predict` is the fused all-task schedule.
    def _stage_plan(self, key, make):
        if key not in self._plans:
            self._plans[key] = make()
        return self._plans[key]

    @torch.no_grad()
    def encode_rgb(self, rgb_norm: torch.Tensor) -> torch.Tensor:
        """encode_rgb (stablemtl_pipeline.py:607-624): [B,3,H,W] in [-1,1] -> latent mean * 0.18215, fp32 [B,4,H/8,W/8]"""
        B, _, H, W = rgb_norm.shape
        enc = self._stage_plan(("encode_rgb", B, H, W), lambda: VAEEncodePlan(self.vae_w, B, H, W, normalized=True))
        enc.rgb.copy_(rgb_norm.to(self.device, F32), non_blocking=True)
        enc.run()
        return enc.out.view(B, enc.h, enc.w, -1).permute(0, 3, 1, 2).contiguous()

    @torch.no_grad()
    def decode_latents(self, latent: torch.Tensor) -> torch.Tensor:
        """the decoder of decode_output (:626-643): task latent [B,4,h,w] -> decoder output fp32 [B,3,8h,8w] (unclipped)"""
        B, C, h, w = latent.shape
        bd = min(B, self.max_decode_batch)
        dec = self._stage_plan(("decode", bd, h, w), lambda: VAEDecodePlan(self.vae_w, bd, h, w))
        flat = latent.to(self.device, F32).permute(0, 2, 3, 1).reshape(B * h * w, C)
        H, W = dec.H, dec.Wd
        out = torch.empty(B, 3, H, W, device=self.device, dtype=F32)
        for c0 in range(0, B, bd):
            n = min(bd, B - c0)
            dec.latent[: n * h * w].copy_(flat[c0 * h * w:(c0 + n) * h * w])
            dec.run()
            out[c0:c0 + n].copy_(dec.out[: n * H * W].view(n, H, W, 3).permute(0, 3, 1, 2))
        return out

    def task_of_text(self, encoder_hidden_states: torch.Tensor) -> int:
        """index of the task whose (constant) prompt embedding this is: the cross-attention K/V are folded per task at
        load time (stablemtl_pipeline.py:464-472), so a UNet call is conditioned by naming one of the engine's tasks"""
        e = encoder_hidden_states.detach().to("cpu", F32)
        if e.dim() == 2:
            e = e[None]
        if not all(torch.equal(e[0], e[i]) for i in range(1, e.shape[0])):
            raise ValueError("one prompt per UNet call: every batch row must carry the same text embedding")
        for i, t in enumerate(self.tasks):
            ref = self.child_w.text[i, : self.child_w.ntok[i]]
            if e.shape[1] == ref.shape[0] and torch.allclose(e[0], ref, rtol=1e-4, atol=1e-5):
                return i
        raise ValueError("encoder_hidden_states is not the embedding of one of the engine's task prompts "
                         f"({', '.join(self.tasks)}): the accelerated UNet folds the constant prompts at load time")

    @torch.no_grad()
    def unet_forward(self, which: str, sample: torch.Tensor, task: int, task_feats=None):
        """UNet3DConditionModel.forward (src/model/unet.py:284-445) at t = 999 for one task prompt.
        which: "single" | "child" (also returns the 16 attn1 taps) | "main" (consumes task_feats: list[16] of
        {task name: [B, N_l, C_l]}, attending to every stream given, attention.py:463-600).
        sample [B,12,1,h,w] -> (sample [B,4,1,h,w] fp32, taps list[16] of fp32 [B, N_l, C_l] or None)."""
        if sample.dim() != 5 or sample.shape[2] != 1 or sample.shape[1] != self.ucfg.in_channels:
            raise ValueError(f"sample must be [B, {self.ucfg.in_channels}, 1, h, w], got {tuple(sample.shape)}")
        B, C, _, h, w = sample.shape
        W_ = self.main_w if which == "main" else self.child_w
        if W_ is None:
            raise ValueError("this engine has no main (multi-stream) UNet")
        src = None
        if which == "main":
            names = [t for t in self.tasks if t in task_feats[0]]
            if not names or len(task_feats) != len(UNetPlan.tap_shapes(self.ucfg, h, w)):
                raise ValueError("task_feats must be one {task: features} dict per transformer layer (16 for SD-2)")
            src = [self.tasks.index(t) for t in names]
        key = ("unet", which, task, None if src is None else tuple(src), B, h, w)

        def make():
            feats = None
            if which == "main":
                feats = [torch.zeros(len(src) * B * n, c, device=self.device, dtype=ops.h16())
                         for n, c in UNetPlan.tap_shapes(self.ucfg, h, w)]
            return UNetPlan(W_, B, h, w, [task], mode=which, feats=feats, src_tasks=src, exclude_self=False)
        plan = self._stage_plan(key, make)
        plan.x_in.copy_(sample.to(self.device, F32).squeeze(2).permute(0, 2, 3, 1).reshape(B * h * w, C))
        if which == "main":
            for l, buf in enumerate(plan.feats_in):
                n = buf.shape[0] // len(src)
                for si, ti in enumerate(src):
                    f = task_feats[l][self.tasks[ti]]
                    buf[si * n:(si + 1) * n].copy_(f.to(self.device).reshape(n, -1))
        plan.run()
        out = plan.out.view(B, h, w, -1).permute(0, 3, 1, 2).unsqueeze(2).contiguous()
        taps = None
        if which == "child":
            taps = [f.view(B, -1, f.shape[-1]).float() for f in plan.feats_out]
        return out, taps

    # ------------------------------------------------------------------------------------------ range audit
    @torch.no_grad()
    def audit_range(self, rgb: torch.Tensor, rgb_next: Optional[torch.Tensor] = None, headroom: float = 0.5):
        """IEEE fp16 operands saturate at +-65504 (the kernels clamp instead of producing inf), which would be a SILENT
        error for a checkpoint whose activations leave that range.  This runs one pass op by op (no CUDA graph) and
        looks at every 16-bit tensor the kernels write: it returns [(max |x|, plan, op index, op name)] sorted by
        magnitude and raises OverflowError if any output reaches `headroom` * 65504 -- run it once per checkpoint on
        representative images; activations depend on the weights far more than on the image.  bf16 operands
        (ops.set_precision("bf16")) have the fp32 range and nothing to audit."""
        if self.stream:
            raise NotImplementedError("audit_range runs the unsharded schedule")
        B, _, H, W = rgb.shape
        dt = torch.uint8 if rgb.dtype == torch.uint8 else F32
        p = self.plan_for(B, H, W, rgb_next is not None, dt)
        p["enc"].rgb[:B].copy_(rgb)
        if rgb_next is not None:
            p["enc"].rgb[B:].copy_(rgb_next)
        report = []

        def run_plan(pname, plan):
            for i, op in enumerate(plan.ops):
                op.run()
                for t in op.outs16:
                    if t.dtype == torch.float16:
                        report.append((float(t.abs().max()), pname, i, op.name))
        run_plan("vae_encode", p["enc"].plan)
        p["assemble"].run()
        for ui, u in enumerate(p["unets"]):
            run_plan(f"unet{ui}", u.plan)
        dec, bd, hw, lat = p["dec"], p["bd"], p["hw"], p["lat"]
        for c0, maps in p["chunks"]:
            dec.latent.copy_(lat[c0 * hw:(c0 + bd) * hw])
            run_plan(f"vae_decode[{c0}]", dec.plan)
            for m in maps:
                m.run()
        torch.cuda.synchronize()
        report.sort(key=lambda r: -r[0])
        limit = headroom * 65504.0
        if report and not (report[0][0] < limit):                     # also catches NaN
            worst = ", ".join(f"{n}#{i} {name}: {m:.0f}" for m, n, i, name in report[:5])
            raise OverflowError(f"fp16 range: activations reach {report[0][0]:.0f} (limit {limit:.0f} = {headroom} x 65504) -- "
                                f"{worst}.  Run this checkpoint with ops.set_precision('bf16').")
        return report

    def launches_per_step(self, B, H, W, with_next=True):
        return self.plan_for(B, H, W, with_next)["launches"]

    def flops_per_step(self, B, H, W, with_next=True):
        return self.plan_for(B, H, W, with_next)["flops"]


def __getattr__(name):
    """`from stablemtl_b200.pipeline import StableMTLPipeline` keeps working: the facade lives in dropin.py"""
    if name == "StableMTLPipeline":
        from .dropin import StableMTLPipeline
        return StableMTLPipeline
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
