"""Plan builders: the SD-2 UNet (single-stream / child / multi-stream main) and the SD-2 VAE expressed as
launch lists over the C-ABI kernels (ops.py).  A plan is built once per (batch, resolution) with statically
assigned buffers and replayed natively (`Plan.run()` -> smtl_run_plan).

Layouts: feature maps are pixel-major matrices; the residual stream is fp32 "compact" [B*H*W, C]; conv/GEMM
operands are bf16 (padded zero-halo layout for 3x3 convs).  See include/stablemtl_sm100.h.

Reference structure followed (paths under /root/reference):
  UNet   src/model/unet.py:284-445, src/model/unet_blocks.py, src/model/resnet.py, src/model/attention.py
  VAE    diffusers AutoencoderKL as used by src/stablemtl_pipeline.py:607-656
"""
import math

import torch

from . import _lib as L
from . import ops
from .synth import UNetConfig, VAEConfig
from .weights import conv_weight_matrix, interleave_geglu

F32 = torch.float32
LATENT_SCALE = 0.18215          # stablemtl_pipeline.py:134-135


def down_size(n):               # conv 3x3 stride 2 pad 1 (resnet.py:87)
    return (n - 1) // 2 + 1


class Pool:
    """Plan-time buffer pool: intermediates are carved out of reusable device blocks; a block returns to the free
    list when the builder releases it (stream order makes reuse after the last consumer safe)."""

    def __init__(self, device):
        self.device = device
        self.free = []          # list of uint8 tensors
        self.owner = {}         # data_ptr -> block
        self.total = 0

    def alloc(self, shape, dtype):
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = max(256, (n * torch.empty((), dtype=dtype).element_size() + 255) // 256 * 256)
        best = None
        for i, blk in enumerate(self.free):
            if blk.numel() >= nbytes and (best is None or blk.numel() < self.free[best].numel()):
                best = i
        if best is not None and self.free[best].numel() <= max(2 * nbytes, nbytes + (64 << 20)):
            blk = self.free.pop(best)
        else:
            blk = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.total += nbytes
        t = blk[: n * torch.empty((), dtype=dtype).element_size()].view(dtype).reshape(shape)
        self.owner[t.data_ptr()] = blk
        return t

    def release(self, *tensors):
        for t in tensors:
            if t is None:
                continue
            if isinstance(t, Act):
                t = t.t
            blk = self.owner.pop(t.data_ptr(), None)
            if blk is not None:
                self.free.append(blk)


class Act:
    """A feature map between layers: 16-bit compact [rows, C] plus the per-(image, channel) sum / sum-of-squares
    buffer its producing GEMM fills (the statistics of the NEXT GroupNorm, so no separate reduction pass)."""
    __slots__ = ("t", "stats")

    def __init__(self, t, stats=None):
        self.t, self.stats = t, stats

    @property
    def shape(self):
        return self.t.shape


class StatsArena:
    """All statistics buffers of a plan live in a few contiguous chunks that one memset each zeroes at plan start
    (every buffer is accumulated into by exactly one GEMM per run)."""

    def __init__(self, device, chunk_bytes=32 << 20):
        self.device, self.chunk_bytes = device, chunk_bytes
        self.chunks, self.used = [], []

    def new(self, images, channels):
        rep = max(1, min(ops.STATS_REPLICAS, 64 // images))
        n = rep * images * channels * 4
        n_al = (n + 31) // 32 * 32
        if not self.chunks or self.used[-1] + n_al > self.chunks[-1].numel():
            self.chunks.append(torch.zeros(max(self.chunk_bytes // 8, n_al), device=self.device, dtype=torch.int64))
            self.used.append(0)
        t = self.chunks[-1][self.used[-1]: self.used[-1] + n].view(rep, images, channels, 4)
        self.used[-1] += n_al
        return t

    def memset_ops(self):
        return [ops.memset_zero(c[:u]) for c, u in zip(self.chunks, self.used) if u > 0]


class _PlanBase:
    """buffer helpers shared by the UNet and VAE plan builders"""

    def _new_act(self, images, hw, C):
        return Act(self.pool.alloc((images * hw, C), ops.h16()), self.arena.new(images, C))

    # Statistics granularity of the plan's maps (smtl_gemm_args.stats_group): the largest power of two <= 8 such that every
    # GroupNorm group that will read a map -- alone or inside a channel concat -- is a union of aligned blocks of that
    # many channels.  Set by the plan builders from the channel counts and the group count.
    stats_group = 1

    @staticmethod
    def stats_group_for(channel_counts, groups):
        """Every group of GroupNorm(groups) over any one map, or over a concat of two maps, of these channel counts starts
        at a multiple of gcd(c_i / groups) (concat offsets c_i are multiples of it too)."""
        import math
        g = 8
        for c in channel_counts:
            if c % groups:
                return 1
            g = math.gcd(g, c // groups)
        return g

    def _stats_group_of(self, act):
        return self.stats_group

    def _into(self, act, hw):
        """epilogue kwargs of the GEMM that produces `act`"""
        return dict(out_bf16=act.t, stats=act.stats, stats_rows_per_image=hw, stats_group=self._stats_group_of(act))

    def _finish(self):
        self.plan.ops[0:0] = self.arena.memset_ops()
        self.plan.finalize()


class _DevDict(dict):
    """Packed weights: every tensor is laid out on the HOST (layout changes, 16-bit casts, constant folding) and
    crosses to the device as one plain copy when it is stored here -- weight ingest launches no kernels at all, so the
    first kernels a process launches are the path's own (the driver's launch capture sees them, not ATen's)."""

    def __init__(self, device):
        super().__init__()
        self.device = device

    def __setitem__(self, k, v):
        if isinstance(v, torch.Tensor):
            v = v.contiguous().to(self.device)
        elif isinstance(v, list):
            v = [t.contiguous().to(self.device) for t in v]
        super().__setitem__(k, v)


def w_ln2g(g, t):
    return g(f"{t}.norm2.weight")


def _host_getter(sd):
    return lambda k: sd[k].detach().to("cpu", F32)


# ================================================================================================== UNet weights
class UNetWeights:
    """Reference state dict (Appendix B key layout) -> kernel layouts, with load-time constant folding:
    time embedding -> conv1 bias (t = 999, stablemtl_pipeline.py:552), text -> per-layer cross-attention K/V
    (stablemtl_pipeline.py:464-472), shortcut weights/bias appended to conv2."""

    def __init__(self, sd, cfg: UNetConfig, text, tasks, device, timestep=999):
        self.cfg, self.device, self.tasks = cfg, device, list(tasks)
        self.w = _DevDict(device)
        g = _host_getter(sd)
        c = cfg.block_out_channels
        # --- time embedding (unet.py:347-353; diffusers Timesteps flip_sin_to_cos=True, shift 0)
        half = c[0] // 2
        freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=F32) / half)
        e = torch.tensor([float(timestep)])[:, None] * freqs[None, :]
        emb = torch.cat([torch.cos(e), torch.sin(e)], dim=-1)
        emb = torch.nn.functional.linear(emb, g("time_embedding.linear_1.weight"), g("time_embedding.linear_1.bias"))
        temb = torch.nn.functional.linear(torch.nn.functional.silu(emb), g("time_embedding.linear_2.weight"),
                                          g("time_embedding.linear_2.bias"))
        self.silu_temb = torch.nn.functional.silu(temb)[0]
        # --- text tokens, padded to 4 per task
        self.ntok = [text[t].shape[0] for t in self.tasks]
        for t, n in zip(self.tasks, self.ntok):
            if not (1 <= n <= L.MAX_XATTN_TOKENS and text[t].shape[1] == cfg.cross_attention_dim):
                raise ValueError(f"text embedding of task {t!r} is {tuple(text[t].shape)}: the cross-attention kernel takes "
                                 f"1..{L.MAX_XATTN_TOKENS} tokens of width {cfg.cross_attention_dim}")
        self.ntp = 4 if max(self.ntok) <= 4 else 8                   # padded token count of the cross-attention tables
        txt = torch.zeros(len(self.tasks), self.ntp, cfg.cross_attention_dim)
        for i, t in enumerate(self.tasks):
            txt[i, : self.ntok[i]] = text[t].detach().to("cpu", F32)
        self.text = txt                                               # host
        # --- stem / head
        w_in = g("conv_in.weight")                                   # [c0, 12, 3, 3]
        kin = 9 * cfg.in_channels
        self.kin_pad = (kin + 63) // 64 * 64
        wm = torch.zeros(c[0], self.kin_pad)
        wm[:, :kin] = conv_weight_matrix(w_in)
        self.w["conv_in.w"] = wm.to(ops.h16())
        self.w["conv_in.b"] = g("conv_in.bias")
        self.w["conv_out.w"] = conv_weight_matrix(g("conv_out.weight")).to(ops.h16())
        self.w["conv_out.b"] = g("conv_out.bias")
        self.w["conv_norm_out.g"], self.w["conv_norm_out.b"] = g("conv_norm_out.weight"), g("conv_norm_out.bias")
        self._sd, self._g = sd, g

    # lazily converted per-module weights (cached)
    def resnet(self, p):
        if p + ".w1" not in self.w:
            g, w = self._g, self.w
            w[p + ".n1g"], w[p + ".n1b"] = g(p + ".norm1.weight"), g(p + ".norm1.bias")
            w[p + ".n2g"], w[p + ".n2b"] = g(p + ".norm2.weight"), g(p + ".norm2.bias")
            w[p + ".w1"] = conv_weight_matrix(g(p + ".conv1.weight")).to(ops.h16())
            b1 = g(p + ".conv1.bias")
            if (p + ".time_emb_proj.weight") in self._sd:            # resnet.py:183-186, constant at inference
                b1 = b1 + torch.nn.functional.linear(self.silu_temb, g(p + ".time_emb_proj.weight"),
                                                     g(p + ".time_emb_proj.bias"))
            w[p + ".b1"] = b1.contiguous()
            w2 = conv_weight_matrix(g(p + ".conv2.weight"))
            b2 = g(p + ".conv2.bias")
            if (p + ".conv_shortcut.weight") in self._sd:
                ws = g(p + ".conv_shortcut.weight")
                w2 = torch.cat([w2, ws.reshape(ws.shape[0], ws.shape[1])], dim=1)
                b2 = b2 + g(p + ".conv_shortcut.bias")
                w[p + ".short"] = True
            w[p + ".w2"] = w2.to(ops.h16()).contiguous()
            w[p + ".b2"] = b2.contiguous()
        return self.w

    def conv(self, p):
        if p + ".w" not in self.w:
            self.w[p + ".w"] = conv_weight_matrix(self._g(p + ".weight")).to(ops.h16())
            self.w[p + ".b"] = self._g(p + ".bias")
        return self.w[p + ".w"], self.w[p + ".b"]

    def conv_up(self, p):
        """nearest-2x upsample + 3x3 conv folded into four per-parity 2x2 filters (ops.up2x_weight_matrices)"""
        if p + ".up" not in self.w:
            self.w[p + ".up"] = [m.to(ops.h16()).contiguous() for m in ops.up2x_weight_matrices(self._g(p + ".weight"))]
            self.w[p + ".b"] = self._g(p + ".bias")
        return self.w[p + ".up"], self.w[p + ".b"]

    def transformer(self, p):
        if p + ".qkv" not in self.w:
            g, w = self._g, self.w
            t = p + ".transformer_blocks.0"
            w[p + ".ng"], w[p + ".nb"] = g(p + ".norm.weight"), g(p + ".norm.bias")
            for n in ("proj_in", "proj_out"):
                w[f"{p}.{n}.w"], w[f"{p}.{n}.b"] = g(f"{p}.{n}.weight").to(ops.h16()), g(f"{p}.{n}.bias")
            for i in (1, 2, 3):
                w[f"{p}.ln{i}g"], w[f"{p}.ln{i}b"] = g(f"{t}.norm{i}.weight"), g(f"{t}.norm{i}.bias")
            w[p + ".qkv"] = torch.cat([g(f"{t}.attn1.to_q.weight"), g(f"{t}.attn1.to_k.weight"),
                                       g(f"{t}.attn1.to_v.weight")], dim=0).to(ops.h16()).contiguous()
            w[p + ".o.w"], w[p + ".o.b"] = g(f"{t}.attn1.to_out.0.weight").to(ops.h16()), g(f"{t}.attn1.to_out.0.bias")
            # cross-attention keys/values of the constant task-name tokens: [ntask, ntp, C] fp32
            kc = torch.nn.functional.linear(self.text, g(f"{t}.attn2.to_k.weight"))
            vc = torch.nn.functional.linear(self.text, g(f"{t}.attn2.to_v.weight"))
            wq, wo = g(f"{t}.attn2.to_q.weight"), g(f"{t}.attn2.to_out.0.weight")
            w[p + ".o2.b"] = g(f"{t}.attn2.to_out.0.bias")
            C_ = wq.shape[0]
            H_ = C_ // 64
            if ops.xattn_fused_supported(H_, self.ntp):
                # attn2 collapsed onto the constant prompt (attention.py:355-364; smtl_xattnf_args): 2 * H * ntp vectors
                T_, n = len(self.tasks), self.ntp
                a0 = 0.125 * torch.einsum("tjhd,hdc->thjc", kc.view(T_, n, H_, 64), wq.view(H_, 64, C_))     # scale 1/sqrt(64)
                bm = torch.einsum("tjhd,chd->thjc", vc.view(T_, n, H_, 64), wo.view(C_, H_, 64))
                w[p + ".xf.ap"], w[p + ".xf.ca"], w[p + ".xf.bmt"] = ops.xattn_tables(
                    a0, w_ln2g(g, t), g(f"{t}.norm2.bias"), bm, self.ntok, n)
                w[p + ".xf"] = True
            else:
                w[p + ".q2"] = wq.to(ops.h16())
                w[p + ".kc"], w[p + ".vc"] = kc.contiguous(), vc.contiguous()
                w[p + ".o2.w"] = wo.to(ops.h16())
            wi, bi = interleave_geglu(g(f"{t}.ff.net.0.proj.weight"), g(f"{t}.ff.net.0.proj.bias"))
            w[p + ".ff1.w"], w[p + ".ff1.b"] = wi.to(ops.h16()), bi
            w[p + ".ff2.w"], w[p + ".ff2.b"] = g(f"{t}.ff.net.2.weight").to(ops.h16()), g(f"{t}.ff.net.2.bias")
        return self.w

    def task_modules(self, p):
        """per-task MLPs / norms stacked over tasks (util/model.py:102-146)."""
        if p + ".tq0.w" not in self.w:
            g, w = self._g, self.w
            a = p + ".transformer_blocks.0.attn1"
            T = self.tasks
            st = lambda fmt, dt=F32: torch.stack([g(fmt.format(t=t)) for t in T]).to(dt).contiguous()
            for kv in ("k", "v"):
                w[f"{p}.t{kv}1.w"], w[f"{p}.t{kv}1.b"] = st(a + ".task_to_" + kv + ".{t}.fc1.weight", ops.h16()), st(a + ".task_to_" + kv + ".{t}.fc1.bias")
                w[f"{p}.t{kv}2.w"], w[f"{p}.t{kv}2.b"] = st(a + ".task_to_" + kv + ".{t}.fc2.weight", ops.h16()), st(a + ".task_to_" + kv + ".{t}.fc2.bias")
                w[f"{p}.tn{kv}.g"], w[f"{p}.tn{kv}.b"] = st(a + ".task_norm_" + kv + ".{t}.weight"), st(a + ".task_norm_" + kv + ".{t}.bias")
            for i, n in enumerate((0, 2, 4, 6)):
                w[f"{p}.tq{i}.w"], w[f"{p}.tq{i}.b"] = st(a + ".task_to_q.{t}.net." + str(n) + ".weight", ops.h16()), st(a + ".task_to_q.{t}.net." + str(n) + ".bias")
            w[f"{p}.tnq.g"], w[f"{p}.tnq.b"] = st(a + ".task_norm_q.{t}.weight"), st(a + ".task_norm_q.{t}.bias")
            w[p + ".tout.w"], w[p + ".tout.b"] = g(a + ".to_out_task.weight").to(ops.h16()), g(a + ".to_out_task.bias")
        return self.w


# ================================================================================================== UNet plan
class UNetPlan(_PlanBase):
    """One batched UNet pass over `groups` row groups of `images` images each (group g runs task group_tasks[g]).

    mode "single": plain UNet (StableMTL-S / child without taps)
    mode "child" : also writes the 16 attn1 outputs (bf16) for the task attention of the main pass
    mode "main"  : consumes `feats` (list of 16 bf16 [n_src*images*N_l, C_l], rows grouped by source task)
    """

    @staticmethod
    def tap_shapes(cfg, h, w):
        """[(tokens per image, channels)] of the 16 attn1 taps in execution order (child feats_out / main feats)."""
        c = cfg.block_out_channels
        nlev = len(c)
        sizes = [(h, w)]
        for _ in range(nlev - 1):
            sizes.append((down_size(sizes[-1][0]), down_size(sizes[-1][1])))
        out = []
        for i in range(nlev - 1):
            out += [(sizes[i][0] * sizes[i][1], c[i])] * cfg.layers_per_block
        out.append((sizes[-1][0] * sizes[-1][1], c[-1]))
        for i in range(1, nlev):
            lev = nlev - 1 - i
            out += [(sizes[lev][0] * sizes[lev][1], c[lev])] * (cfg.layers_per_block + 1)
        return out

    def __init__(self, W: UNetWeights, images, h, w, group_tasks, mode="single", feats=None, src_tasks=None,
                 pool=None, x_in=None, feat_bufs=None, exclude_self=True):
        """feat_bufs (child mode): caller-owned 16-bit tap buffers, one per layer, at least [G*images*N_l, C_l] -- the
        send buffers of the task-stream exchange (stablemtl_b200/stream_shard.py).  src_tasks (main mode) may hold -1
        for an empty exchange slot: its rows are skipped by the task attention."""
        cfg = W.cfg
        self.W, self.cfg, self.mode = W, cfg, mode
        dev = W.device
        self.pool = pool or Pool(dev)
        P = self.pool
        self.arena = StatsArena(dev)
        self.plan = ops.Plan()
        self.stats_group = self.stats_group_for(cfg.block_out_channels, cfg.norm_num_groups)   # SD-2: gcd(10, 20, 40) -> 2
        add = self.plan.add
        G = len(group_tasks)
        Be = G * images
        self.Be, self.h, self.w = Be, h, w
        c = cfg.block_out_channels
        nlev = len(c)
        sizes = [(h, w)]
        for _ in range(nlev - 1):
            sizes.append((down_size(sizes[-1][0]), down_size(sizes[-1][1])))
        self.group_tasks = list(group_tasks)
        self.images = images
        self.feats_out = [] if mode == "child" else None
        self.feats_in, self.src_tasks = feats, src_tasks
        self.feat_bufs = feat_bufs
        self.exclude_self = exclude_self        # stablemtl_pipeline.py:483-484; False: attend to every stream given
        self.layer = 0

        # ---- stem: [Be, hw, 12] fp32 -> im2col -> GEMM
        self.x_in = x_in if x_in is not None else torch.zeros(Be * h * w, cfg.in_channels, device=dev, dtype=F32)
        col = P.alloc((Be * h * w, W.kin_pad), ops.h16())
        add(ops.im2col(self.x_in.view(Be, h, w, cfg.in_channels), Be, h, w, col, stride=1, pad_t=1, pad_l=1, oh=h, ow=w))
        x = self._new_act(Be, h * w, c[0])
        add(ops.gemm(col, W.w["conv_in.w"], bias=W.w["conv_in.b"], name="conv_in", **self._into(x, h * w)))
        P.release(col)

        skips = [x]
        ch = c[0]
        for i in range(nlev):
            hh, ww = sizes[i]
            for j in range(cfg.layers_per_block):
                x_new = self.resnet(f"down_blocks.{i}.resnets.{j}", x, None, hh, ww, c[i])
                if x is not skips[-1]:
                    P.release(x)
                x = x_new
                if i < nlev - 1:
                    x_new = self.transformer(f"down_blocks.{i}.attentions.{j}", x, hh, ww, c[i], cfg.heads[i])
                    P.release(x)
                    x = x_new
                skips.append(x)
            if i < nlev - 1:
                h2, w2 = sizes[i + 1]
                wt, bs = W.conv(f"down_blocks.{i}.downsamplers.0.conv")
                col = P.alloc((Be * h2 * w2, 9 * c[i]), ops.h16())
                add(ops.im2col(x.t.view(Be, hh, ww, c[i]), Be, hh, ww, col, stride=2, pad_t=1, pad_l=1, oh=h2, ow=w2))
                x = self._new_act(Be, h2 * w2, c[i])
                add(ops.gemm(col, wt, bias=bs, name="downsample", **self._into(x, h2 * w2)))
                P.release(col)
                skips.append(x)
        hh, ww = sizes[-1]
        x_new = self.resnet("mid_block.resnets.0", x, None, hh, ww, c[-1])
        x = x_new                                   # previous x is still referenced as a skip
        x_new = self.transformer("mid_block.attentions.0", x, hh, ww, c[-1], cfg.heads[-1])
        P.release(x)
        x = x_new
        x_new = self.resnet("mid_block.resnets.1", x, None, hh, ww, c[-1])
        P.release(x)
        x = x_new
        for i in range(nlev):
            lev = nlev - 1 - i
            hh, ww = sizes[lev]
            cout = c[lev]
            for j in range(cfg.layers_per_block + 1):
                skip = skips.pop()
                x_new = self.resnet(f"up_blocks.{i}.resnets.{j}", x, skip, hh, ww, cout)
                P.release(x, skip)
                x = x_new
                if i > 0:
                    x_new = self.transformer(f"up_blocks.{i}.attentions.{j}", x, hh, ww, cout, cfg.heads[lev])
                    P.release(x)
                    x = x_new
            if i < nlev - 1:
                oh, ow = sizes[lev - 1]                       # explicit size of the next skip (unet.py:415-416)
                x_new = self._new_act(Be, oh * ow, cout)
                if (oh, ow) == (2 * hh, 2 * ww):
                    # exact 2x: four 2x2 convs on the low-res map (2.25x fewer FLOPs, no 4x-sized intermediate)
                    wmats, bs = W.conv_up(f"up_blocks.{i}.upsamplers.0.conv")
                    low = P.alloc((Be * (hh + 2) * (ww + 2), cout), ops.h16())
                    add(ops.upsample_pad(x.t.view(Be, hh, ww, cout), Be, hh, ww, hh, ww, low))   # zero-halo copy
                    for o in ops.conv_up2x(low, wmats, Be, hh, ww, bias=bs, name="upsample_conv",
                                           **self._into(x_new, oh * ow)):
                        add(o)
                    P.release(low, x)
                else:
                    wt, bs = W.conv(f"up_blocks.{i}.upsamplers.0.conv")
                    up = P.alloc((Be * (oh + 2) * (ow + 2), cout), ops.h16())
                    add(ops.upsample_pad(x.t.view(Be, hh, ww, cout), Be, hh, ww, oh, ow, up))
                    add(ops.conv3x3(up, wt, Be, oh, ow, bias=bs, name="upsample_conv", **self._into(x_new, oh * ow)))
                    P.release(up, x)
                x = x_new
        # ---- head
        a = P.alloc((Be * (h + 2) * (w + 2), c[0]), ops.h16())
        add(ops.gn_apply(x.t, x.stats, Be, h, w, W.w["conv_norm_out.g"], W.w["conv_norm_out.b"], a, eps=cfg.norm_eps,
                         silu=True, pad_out=True, groups=cfg.norm_num_groups))
        self.out = torch.empty(Be * h * w, cfg.out_channels, device=dev, dtype=F32)
        add(ops.conv3x3(a, W.w["conv_out.w"], Be, h, w, bias=W.w["conv_out.b"], out_f32=self.out, name="conv_out"))
        P.release(a, x)
        self._finish()

    # ---------------------------------------------------------------------------------------------- resnet
    def resnet(self, p, x, skip, h, w, cout):
        """ResnetBlock3D (resnet.py:174-204) on the virtual concat [x, skip] (unet_blocks.py:509,597)."""
        W, P, add, cfg, Be = self.W, self.pool, self.plan.add, self.cfg, self.Be
        wt = W.resnet(p)
        cin = x.shape[1] + (skip.shape[1] if skip is not None else 0)
        short = wt.get(p + ".short", False)
        a1 = P.alloc((Be * (h + 2) * (w + 2), cin), ops.h16())
        raw = P.alloc((Be * (h + 2) * (w + 2), cin), ops.h16()) if short else None
        add(ops.gn_apply(x.t, x.stats, Be, h, w, wt[p + ".n1g"], wt[p + ".n1b"], a1,
                         x1=None if skip is None else skip.t, stats1=None if skip is None else skip.stats,
                         eps=cfg.norm_eps, silu=True, pad_out=True, raw=raw, groups=cfg.norm_num_groups))
        h1 = self._new_act(Be, h * w, cout)
        add(ops.conv3x3(a1, wt[p + ".w1"], Be, h, w, bias=wt[p + ".b1"], name="res.conv1", **self._into(h1, h * w)))
        P.release(a1)
        a2 = P.alloc((Be * (h + 2) * (w + 2), cout), ops.h16())
        add(ops.gn_apply(h1.t, h1.stats, Be, h, w, wt[p + ".n2g"], wt[p + ".n2b"], a2, eps=cfg.norm_eps, silu=True,
                         pad_out=True, groups=cfg.norm_num_groups))
        P.release(h1)
        out = self._new_act(Be, h * w, cout)
        add(ops.conv3x3(a2, wt[p + ".w2"], Be, h, w, a_short=raw, bias=wt[p + ".b2"], res1=None if short else x.t,
                        name="res.conv2", **self._into(out, h * w)))
        P.release(a2, raw)
        return out

    # ---------------------------------------------------------------------------------------------- transformer
    def transformer(self, p, x, h, w, C, heads):
        """Transformer3DModel + BasicTransformerBlock (attention.py:174-223, 323-380) on tokens [Be*N, C]."""
        W, P, add, cfg, Be = self.W, self.pool, self.plan.add, self.cfg, self.Be
        wt = W.transformer(p)
        N = h * w
        M = Be * N
        rpg = self.images * N                      # rows per task group
        xn = P.alloc((M, C), ops.h16())
        add(ops.gn_apply(x.t, x.stats, Be, h, w, wt[p + ".ng"], wt[p + ".nb"], xn, eps=1e-6, silu=False, pad_out=False,
                         groups=cfg.norm_num_groups))
        hs = P.alloc((M, C), F32)
        add(ops.gemm(xn, wt[p + ".proj_in.w"], bias=wt[p + ".proj_in.b"], out_f32=hs, name="proj_in"))
        n1 = xn                                    # reuse as LN output
        add(ops.layer_norm(hs, wt[p + ".ln1g"], wt[p + ".ln1b"], n1))
        qkv = P.alloc((M, 3 * C), ops.h16())
        add(ops.gemm(n1, wt[p + ".qkv"], out_bf16=qkv, name="qkv"))
        att = n1
        add(ops.flash_attn(qkv, Be, N, heads, att, 0, C, 2 * C))
        P.release(qkv)
        if self.mode != "main":
            feat = None
            if self.mode == "child":
                if self.feat_bufs is not None:
                    feat = self.feat_bufs[len(self.feats_out)][:M]
                    assert feat.shape == (M, C) and feat.dtype == ops.h16() and feat.is_contiguous()
                else:
                    feat = torch.empty(M, C, device=W.device, dtype=ops.h16())  # owned by the plan, read by the main pass
                self.feats_out.append(feat)
            # h += to_out(attn); the pre-residual value is the "afterSelfAttn_residual" tap (attention.py:348-349)
            add(ops.gemm(att, wt[p + ".o.w"], bias=wt[p + ".o.b"], res1=hs, out_f32=hs, aux_bf16=feat, name="attn_out"))
        else:
            tw = W.task_modules(p)
            nT = len(W.tasks)
            attn_out = P.alloc((M, C), F32)
            add(ops.gemm(att, wt[p + ".o.w"], bias=wt[p + ".o.b"], out_f32=attn_out, name="attn_out"))
            # q = MLPv2_q[main task](LN_q[main task](attn_out))         attention.py:512
            qn = att
            def rows_of(key, tasks):
                """[len(tasks), ...] rows of a per-task stack in row-group order (a view when that is every task in order)"""
                t = tw[p + key]
                if list(tasks) == list(range(t.shape[0])):
                    return t
                return t.index_select(0, torch.tensor([max(x, 0) for x in tasks], device=t.device)).contiguous()
            gq, bq = rows_of(".tnq.g", self.group_tasks), rows_of(".tnq.b", self.group_tasks)
            add(ops.layer_norm(attn_out, gq, bq, qn, rows_per_group=rpg))
            hq = cfg.task_q_hidden
            q1, q2 = P.alloc((M, hq), ops.h16()), P.alloc((M, hq), ops.h16())
            tq = P.alloc((M, C), ops.h16())
            grouped = True                    # one grouped launch for all task streams (M tiles restart at every group)

            def stacked(key, tasks):
                """[len(tasks) * out, in] weights / [len(tasks) * out] bias of the per-task module, in row-group order"""
                wk, bk = rows_of(key + ".w", tasks), rows_of(key + ".b", tasks)      # -1 = empty exchange slot -> task 0
                return wk.reshape(-1, wk.shape[-1]), bk.reshape(-1), wk.shape[1]

            def task_linear(key, tasks, src, dst, act, name):
                if grouped:
                    wg, bg, nout = stacked(key, tasks)
                    add(ops.gemm(src, wg, n=nout, bias=bg, act=act, out_bf16=dst, group_rows=rpg, name=name))
                else:
                    for gi, t in enumerate(tasks):
                        if t < 0:
                            continue
                        r = slice(gi * rpg, (gi + 1) * rpg)
                        add(ops.gemm(src[r], tw[p + key + ".w"][t], bias=tw[p + key + ".b"][t], act=act, out_bf16=dst[r],
                                     name=name))
            task_linear(".tq0", self.group_tasks, qn, q1, L.ACT_GELU, "task_q0")
            task_linear(".tq1", self.group_tasks, q1, q2, L.ACT_GELU, "task_q1")
            task_linear(".tq2", self.group_tasks, q2, q1, L.ACT_GELU, "task_q2")
            task_linear(".tq3", self.group_tasks, q1, tq, L.ACT_NONE, "task_q3")
            P.release(q1, q2)
            # k_t, v_t = MLP_{k,v}[t](LN_{k,v}[t](feat_t)) once per source stream           attention.py:494-495
            F_l = self.feats_in[self.layer]
            S = len(self.src_tasks)
            Ms = S * rpg
            assert F_l.shape == (Ms, C), (F_l.shape, Ms, C)
            kn, vn = P.alloc((Ms, C), ops.h16()), P.alloc((Ms, C), ops.h16())
            gk, bk = rows_of(".tnk.g", self.src_tasks), rows_of(".tnk.b", self.src_tasks)
            gv, bv = rows_of(".tnv.g", self.src_tasks), rows_of(".tnv.b", self.src_tasks)
            add(ops.layer_norm(F_l, gk, bk, kn, gamma1=gv, beta1=bv, out1=vn, rows_per_group=rpg))
            hk, hv = P.alloc((Ms, C // 2), ops.h16()), P.alloc((Ms, C // 2), ops.h16())
            K, V = P.alloc((Ms, C), ops.h16()), P.alloc((Ms, C), ops.h16())
            task_linear(".tk1", self.src_tasks, kn, hk, L.ACT_GELU, "task_k1")
            task_linear(".tv1", self.src_tasks, vn, hv, L.ACT_GELU, "task_v1")
            task_linear(".tk2", self.src_tasks, hk, K, L.ACT_NONE, "task_k2")
            task_linear(".tv2", self.src_tasks, hv, V, L.ACT_NONE, "task_v2")
            P.release(kn, vn, hk, hv)
            ta = qn
            add(ops.task_attn(tq, K, V, ta, C, cfg.n_attns, self.group_tasks, self.src_tasks, rpg, exclude_self=self.exclude_self))
            P.release(tq, K, V)
            # h += attn_out + to_out_task(task attention)                                   attention.py:598-600, :347
            add(ops.gemm(ta, tw[p + ".tout.w"], bias=tw[p + ".tout.b"], res1=attn_out, res2=hs, out_f32=hs, name="to_out_task"))
            P.release(attn_out)
        self.layer += 1
        # ---- cross-attention on the task-name tokens (attention.py:355-364)
        n3 = att
        if wt.get(p + ".xf", False):
            # collapsed onto the constant prompt, fused with LayerNorm2, the residual add and LayerNorm3: one pass over hs
            add(ops.xattn_fused(hs, wt[p + ".xf.ap"], wt[p + ".xf.ca"], wt[p + ".xf.bmt"], wt[p + ".o2.b"],
                                wt[p + ".ln3g"], wt[p + ".ln3b"], self.group_tasks, rpg, heads, W.ntp, n3))
        else:
            n2 = att
            add(ops.layer_norm(hs, wt[p + ".ln2g"], wt[p + ".ln2b"], n2))
            q2b = P.alloc((M, C), ops.h16())
            add(ops.gemm(n2, wt[p + ".q2"], out_bf16=q2b, name="xattn_q"))
            xa = n2
            add(ops.xattn(q2b, wt[p + ".kc"], wt[p + ".vc"], W.ntok, self.group_tasks, rpg, heads, xa))
            P.release(q2b)
            add(ops.gemm(xa, wt[p + ".o2.w"], bias=wt[p + ".o2.b"], res1=hs, out_f32=hs, name="xattn_out"))
            # ---- GEGLU feed-forward (attention.py:372-373)
            add(ops.layer_norm(hs, wt[p + ".ln3g"], wt[p + ".ln3b"], n3))
        gg = P.alloc((M, 4 * C), ops.h16())
        add(ops.gemm(n3, wt[p + ".ff1.w"], bias=wt[p + ".ff1.b"], act=L.ACT_GEGLU, out_bf16=gg, name="ff1_geglu"))
        hb = n3
        add(ops.gemm(gg, wt[p + ".ff2.w"], bias=wt[p + ".ff2.b"], res1=hs, out_bf16=hb, name="ff2"))
        P.release(gg, hs)
        out = self._new_act(Be, N, C)
        add(ops.gemm(hb, wt[p + ".proj_out.w"], bias=wt[p + ".proj_out.b"], res1=x.t, name="proj_out",
                     **self._into(out, N)))
        P.release(hb)
        return out

    def run(self):
        self.plan.run()


# ================================================================================================== VAE
class VAEWeights:
    def __init__(self, sd, cfg: VAEConfig, device):
        self.cfg, self.device, self._sd = cfg, device, sd
        self.w = _DevDict(device)
        self._g = _host_getter(sd)
        g, w = self._g, self.w
        # encoder stem (3 -> c0), K = 27 padded to 64
        wm = torch.zeros(cfg.block_out_channels[0], 64)
        wm[:, :27] = conv_weight_matrix(g("encoder.conv_in.weight"))
        w["enc.conv_in.w"], w["enc.conv_in.b"] = wm.to(ops.h16()), g("encoder.conv_in.bias")
        # encoder head: conv_out (3x3, C -> 2L) then quant_conv (1x1) then mean half * 0.18215, folded into one conv C -> L
        lat = cfg.latent_channels
        wq = g("quant_conv.weight").reshape(2 * lat, 2 * lat)[:lat]                # [L, 2L]
        wco = conv_weight_matrix(g("encoder.conv_out.weight"))                     # [2L, 9C]
        w["enc.head.w"] = (LATENT_SCALE * (wq @ wco)).to(ops.h16()).contiguous()
        w["enc.head.b"] = (LATENT_SCALE * (wq @ g("encoder.conv_out.bias") + g("quant_conv.bias")[:lat])).contiguous()
        # decoder stem: latent / 0.18215 -> post_quant_conv (1x1) as a channel mix, then conv_in (L -> C), K = 36 -> 64
        w["dec.pq.w"] = (g("post_quant_conv.weight").reshape(lat, lat) / LATENT_SCALE).contiguous()
        w["dec.pq.b"] = g("post_quant_conv.bias")
        cd = cfg.block_out_channels[-1]
        wm = torch.zeros(cd, 64)
        wm[:, : 9 * lat] = conv_weight_matrix(g("decoder.conv_in.weight"))
        w["dec.conv_in.w"], w["dec.conv_in.b"] = wm.to(ops.h16()), g("decoder.conv_in.bias")
        w["dec.head.w"] = ops.head_weight_matrix(g("decoder.conv_out.weight")).to(ops.h16())   # taps folded into N
        w["dec.head.b"] = g("decoder.conv_out.bias")
        for s in ("encoder", "decoder"):
            w[f"{s}.ng"], w[f"{s}.nb"] = g(f"{s}.conv_norm_out.weight"), g(f"{s}.conv_norm_out.bias")

    def resnet(self, p):
        if p + ".w1" not in self.w:
            g, w = self._g, self.w
            w[p + ".n1g"], w[p + ".n1b"] = g(p + ".norm1.weight"), g(p + ".norm1.bias")
            w[p + ".n2g"], w[p + ".n2b"] = g(p + ".norm2.weight"), g(p + ".norm2.bias")
            w[p + ".w1"], w[p + ".b1"] = conv_weight_matrix(g(p + ".conv1.weight")).to(ops.h16()), g(p + ".conv1.bias")
            w2, b2 = conv_weight_matrix(g(p + ".conv2.weight")), g(p + ".conv2.bias")
            if (p + ".conv_shortcut.weight") in self._sd:
                ws = g(p + ".conv_shortcut.weight")
                w2 = torch.cat([w2, ws.reshape(ws.shape[0], ws.shape[1])], dim=1)
                b2 = b2 + g(p + ".conv_shortcut.bias")
                w[p + ".short"] = True
            w[p + ".w2"], w[p + ".b2"] = w2.to(ops.h16()).contiguous(), b2.contiguous()
        return self.w

    def conv(self, p):
        if p + ".w" not in self.w:
            self.w[p + ".w"] = conv_weight_matrix(self._g(p + ".weight")).to(ops.h16())
            self.w[p + ".b"] = self._g(p + ".bias")
        return self.w[p + ".w"], self.w[p + ".b"]

    def conv_up(self, p):
        if p + ".up" not in self.w:
            self.w[p + ".up"] = [m.to(ops.h16()).contiguous() for m in ops.up2x_weight_matrices(self._g(p + ".weight"))]
            self.w[p + ".b"] = self._g(p + ".bias")
        return self.w[p + ".up"], self.w[p + ".b"]

    def attn(self, p):
        if p + ".qk.w" not in self.w:
            g, w = self._g, self.w
            w[p + ".ng"], w[p + ".nb"] = g(p + ".group_norm.weight"), g(p + ".group_norm.bias")
            w[p + ".qk.w"] = torch.cat([g(p + ".to_q.weight"), g(p + ".to_k.weight")]).to(ops.h16()).contiguous()
            w[p + ".qk.b"] = torch.cat([g(p + ".to_q.bias"), g(p + ".to_k.bias")]).contiguous()
            w[p + ".v.w"], w[p + ".v.b"] = g(p + ".to_v.weight").to(ops.h16()), g(p + ".to_v.bias")
            w[p + ".qkv.w"] = torch.cat([g(p + ".to_q.weight"), g(p + ".to_k.weight"), g(p + ".to_v.weight")]).to(ops.h16())
            w[p + ".qkv.b"] = torch.cat([g(p + ".to_q.bias"), g(p + ".to_k.bias"), g(p + ".to_v.bias")])
            w[p + ".o.w"], w[p + ".o.b"] = g(p + ".to_out.0.weight").to(ops.h16()), g(p + ".to_out.0.bias")
        return self.w


class _VAEBase(_PlanBase):
    """`self.padded`: every inter-layer map of the plan stays in the zero-halo padded layout (conv outputs are written
    there by the epilogue: ROWMAP_PAD_KEEP / TO_PAD / UP2_PAD), so a ResNet's shortcut operand and an up-conv's input
    are the raw maps themselves -- no padded copies."""
    padded = False

    def _rows(self, h, w):
        return self.B * ((h + 2) * (w + 2) if self.padded else h * w)

    def _rpi(self, h, w):
        return (h + 2) * (w + 2) if self.padded else h * w

    def _new_map(self, h, w, C):
        return Act(self.pool.alloc((self._rows(h, w), C), ops.h16()), self.arena.new(self.B, C))

    def _stats_group_of(self, act):
        # no channel concats in the VAE: every map is normalised alone, by GroupNorm(norm_num_groups) over its own channels
        return self.stats_group_for([act.shape[1]], self.W.cfg.norm_num_groups)

    def _resnet(self, p, x, h, w, cout):
        W, P, add, B, G = self.W, self.pool, self.plan.add, self.B, self.W.cfg.norm_num_groups
        wt = W.resnet(p)
        cin = x.shape[1]
        short = wt.get(p + ".short", False)
        pd = self.padded
        a1 = P.alloc((B * (h + 2) * (w + 2), cin), ops.h16())
        raw = P.alloc((B * (h + 2) * (w + 2), cin), ops.h16()) if (short and not pd) else None
        add(ops.gn_apply(x.t, x.stats, B, h, w, wt[p + ".n1g"], wt[p + ".n1b"], a1, eps=1e-6, silu=True, pad_out=True,
                         raw=raw, groups=G, x_padded=pd))
        h1 = self._new_map(h, w, cout)
        add(ops.conv3x3(a1, wt[p + ".w1"], B, h, w, bias=wt[p + ".b1"], name="vae.conv1", pad_out=pd,
                        **self._into(h1, self._rpi(h, w))))
        P.release(a1)
        a2 = P.alloc((B * (h + 2) * (w + 2), cout), ops.h16())
        add(ops.gn_apply(h1.t, h1.stats, B, h, w, wt[p + ".n2g"], wt[p + ".n2b"], a2, eps=1e-6, silu=True, pad_out=True,
                         groups=G, x_padded=pd))
        P.release(h1)
        out = self._new_map(h, w, cout)
        a_short = (x.t if pd else raw) if short else None        # padded mode: the raw input map IS the shortcut operand
        add(ops.conv3x3(a2, wt[p + ".w2"], B, h, w, a_short=a_short, bias=wt[p + ".b2"], res1=None if short else x.t,
                        name="vae.conv2", pad_out=pd, **self._into(out, self._rpi(h, w))))
        P.release(a2, raw)
        return out

    def _mid_attn(self, p, x, h, w, C):
        """diffusers Attention(heads=1, dim_head=C, residual_connection=True) inside UNetMidBlock2D (Appendix A):
        unfused QK^T -> softmax -> PV through the GEMM kernel (single head, d = C = 512)."""
        W, P, add, B, G = self.W, self.pool, self.plan.add, self.B, self.W.cfg.norm_num_groups
        wt = W.attn(p)
        N = h * w
        M = B * N
        Np = (N + 7) // 8 * 8
        xn = P.alloc((M, C), ops.h16())
        add(ops.gn_apply(x.t, x.stats, B, h, w, wt[p + ".ng"], wt[p + ".nb"], xn, eps=1e-6, silu=False, pad_out=False,
                         groups=G, x_padded=self.padded))
        if C == 512:
            # one fused QKV projection, then the d = 512 single-head flash kernel: S and P never leave the SM
            qkv = P.alloc((M, 3 * C), ops.h16())
            add(ops.gemm(xn, wt[p + ".qkv.w"], bias=wt[p + ".qkv.b"], out_bf16=qkv, name="vae.qkv"))
            o = P.alloc((M, C), ops.h16())
            fa = ops.flash_attn(qkv, B, N, 1, o, 0, C, 2 * C, scale=float(C) ** -0.5, head_dim=C)
            fa.name = "vae.flash_attn"
            add(fa)
            P.release(qkv, xn)
            return self._mid_attn_out(p, x, o, h, w, C)
        qk = P.alloc((M, 2 * C), ops.h16())
        add(ops.gemm(xn, wt[p + ".qk.w"], bias=wt[p + ".qk.b"], out_bf16=qk, name="vae.qk"))
        o = P.alloc((M, C), ops.h16())
        vT = P.alloc((C, Np), ops.h16())
        S = P.alloc((N, Np), F32)
        Pm = P.alloc((N, Np), ops.h16())
        for b in range(B):
            r = slice(b * N, (b + 1) * N)
            add(ops.gemm(wt[p + ".v.w"], xn[r], bias=wt[p + ".v.b"], bias_per_row=True, out_bf16=vT[:, :N], name="vae.vT"))
            add(ops.gemm(qk[r, :C], qk[r, C:], out_f32=S[:, :N], name="vae.qkT"))
            add(ops.softmax_rows(S[:, :N], Pm[:, :N], float(C) ** -0.5))
            add(ops.gemm(Pm[:, :N], vT[:, :N], out_bf16=o[r], name="vae.pv"))
        P.release(vT, S, Pm, qk, xn)
        return self._mid_attn_out(p, x, o, h, w, C)

    def _mid_attn_out(self, p, x, o, h, w, C):
        """to_out.0 + the residual connection of the mid-block attention"""
        W, P, add = self.W, self.pool, self.plan.add
        wt = W.attn(p)
        out = self._new_map(h, w, C)
        extra = dict(rowmap=L.ROWMAP_TO_PAD, img_hw=(h, w)) if self.padded else {}
        add(ops.gemm(o, wt[p + ".o.w"], bias=wt[p + ".o.b"], res1=x.t, name="vae.attn_out",
                     **self._into(out, self._rpi(h, w)), **extra))
        P.release(o)
        return out

    def _mid(self, p, x, h, w, C):
        P = self.pool
        y = self._resnet(p + ".resnets.0", x, h, w, C)
        P.release(x)
        z = self._mid_attn(p + ".attentions.0", y, h, w, C)
        P.release(y)
        y = self._resnet(p + ".resnets.1", z, h, w, C)
        P.release(z)
        return y

    def run(self):
        self.plan.run()


class VAEEncodePlan(_VAEBase):
    """encode_rgb (stablemtl_pipeline.py:607-624): rgb [B,3,H,W] in [0,255] -> latent mean * 0.18215, fp32 [B*h*w, 4]."""

    def __init__(self, W: VAEWeights, B, H, Wd, pool=None, rgb=None, rgb_dtype=F32, normalized=False):
        """normalized: `rgb` holds rgb / 255 * 2 - 1 already (the argument of the reference's encode_rgb)"""
        if H % 8 or Wd % 8:
            raise ValueError(f"image size {H}x{Wd}: height and width must be multiples of 8 (the VAE halves the map three "
                             "times and the task maps are produced at 8x the latent size; resize or pad the input first)")
        self.W, self.B = W, B
        dev = W.device
        self.pool = P = pool or Pool(dev)
        self.arena = StatsArena(dev)
        self.plan = ops.Plan()
        add = self.plan.add
        cfg = W.cfg
        c = cfg.block_out_channels
        self.rgb = rgb if rgb is not None else torch.zeros(B, 3, H, Wd, device=dev, dtype=rgb_dtype)
        col = P.alloc((B * H * Wd, 64), ops.h16())
        add(ops.rgb_stem(self.rgb, col, normalized=normalized))      # normalise + im2col of the 3-channel stem, one pass
        x = self._new_act(B, H * Wd, c[0])
        add(ops.gemm(col, W.w["enc.conv_in.w"], bias=W.w["enc.conv_in.b"], name="vae.enc.conv_in",
                     **self._into(x, H * Wd)))
        P.release(col)
        h, w = H, Wd
        for i in range(len(c)):
            for j in range(cfg.layers_per_block):
                y = self._resnet(f"encoder.down_blocks.{i}.resnets.{j}", x, h, w, c[i])
                P.release(x)
                x = y
            if i < len(c) - 1:
                # diffusers Downsample2D: F.pad(x, (0,1,0,1)) then 3x3 stride 2 pad 0
                h2, w2 = (h + 1 - 3) // 2 + 1, (w + 1 - 3) // 2 + 1
                wt, bs = W.conv(f"encoder.down_blocks.{i}.downsamplers.0.conv")
                col = P.alloc((B * h2 * w2, 9 * c[i]), ops.h16())
                add(ops.im2col(x.t.view(B, h, w, c[i]), B, h, w, col, stride=2, pad_t=0, pad_l=0, oh=h2, ow=w2))
                y = self._new_act(B, h2 * w2, c[i])
                add(ops.gemm(col, wt, bias=bs, name="vae.enc.down", **self._into(y, h2 * w2)))
                P.release(col, x)
                x, h, w = y, h2, w2
        x = self._mid("encoder.mid_block", x, h, w, c[-1])
        a = P.alloc((B * (h + 2) * (w + 2), c[-1]), ops.h16())
        add(ops.gn_apply(x.t, x.stats, B, h, w, W.w["encoder.ng"], W.w["encoder.nb"], a, eps=1e-6, silu=True,
                         pad_out=True, groups=cfg.norm_num_groups))
        self.out = torch.empty(B * h * w, cfg.latent_channels, device=dev, dtype=F32)
        add(ops.conv3x3(a, W.w["enc.head.w"], B, h, w, bias=W.w["enc.head.b"], out_f32=self.out, name="vae.enc.head"))
        P.release(a, x)
        self.h, self.w = h, w
        self._finish()


class VAEDecodePlan(_VAEBase):
    """decode_output (stablemtl_pipeline.py:626-643): latent fp32 [B*h*w, 4] -> decoder output fp32 [B*H*W, 3]."""
    padded = True

    def __init__(self, W: VAEWeights, B, h, w, pool=None, latent=None):
        self.W, self.B = W, B
        dev = W.device
        self.pool = P = pool or Pool(dev)
        self.arena = StatsArena(dev)
        self.plan = ops.Plan()
        add = self.plan.add
        cfg = W.cfg
        c = list(reversed(cfg.block_out_channels))
        lat = cfg.latent_channels
        self.latent = latent if latent is not None else torch.zeros(B * h * w, lat, device=dev, dtype=F32)
        H, Wd = h * 2 ** (len(c) - 1), w * 2 ** (len(c) - 1)
        z = P.alloc((B * h * w, lat), F32)
        add(ops.chan_mix(self.latent, W.w["dec.pq.w"], W.w["dec.pq.b"], z))
        col = P.alloc((B * h * w, 64), ops.h16())
        add(ops.im2col(z.view(B, h, w, lat), B, h, w, col, stride=1, pad_t=1, pad_l=1, oh=h, ow=w))
        P.release(z)
        x = self._new_map(h, w, c[0])
        add(ops.gemm(col, W.w["dec.conv_in.w"], bias=W.w["dec.conv_in.b"], name="vae.dec.conv_in",
                     rowmap=L.ROWMAP_TO_PAD, img_hw=(h, w), **self._into(x, self._rpi(h, w))))
        P.release(col)
        x = self._mid("decoder.mid_block", x, h, w, c[0])
        for i in range(len(c)):
            for j in range(cfg.layers_per_block + 1):
                y = self._resnet(f"decoder.up_blocks.{i}.resnets.{j}", x, h, w, c[i])
                P.release(x)
                x = y
            if i < len(c) - 1:
                # diffusers Upsample2D (nearest x2, then 3x3 conv) as four per-parity 2x2 convs on the low-res map
                # the ResNet output is already a zero-halo padded map: it is the up-conv's operand as it stands
                wmats, bs = W.conv_up(f"decoder.up_blocks.{i}.upsamplers.0.conv")
                low = x
                x = self._new_map(2 * h, 2 * w, c[i])
                for o in ops.conv_up2x(low.t, wmats, B, h, w, bias=bs, name="vae.dec.up", pad_out=True,
                                       **self._into(x, self._rpi(2 * h, 2 * w))):
                    add(o)
                P.release(low)
                h, w = 2 * h, 2 * w
        a = P.alloc((B * (h + 2) * (w + 2), c[-1]), ops.h16())
        add(ops.gn_apply(x.t, x.stats, B, h, w, W.w["decoder.ng"], W.w["decoder.nb"], a, eps=1e-6, silu=True,
                         pad_out=True, groups=cfg.norm_num_groups, x_padded=True))
        P.release(x)
        self.out = torch.empty(B * h * w, 3, device=dev, dtype=F32)
        part = P.alloc((B * (h + 2) * (w + 2), W.w["dec.head.w"].shape[0]), F32)
        for o in ops.conv_head(a, W.w["dec.head.w"], W.w["dec.head.b"], B, h, w, 3, part, self.out, name="vae.dec.head"):
            add(o)
        P.release(a, part)
        self.H, self.Wd = h, w
        self._finish()
