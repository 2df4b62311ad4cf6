"""Thin host-side wrappers: torch tensors in, C-ABI op structs out.

Every function returns an `Op` (kind, struct) that can be executed immediately with `.run()` or collected
into a `Plan`, which hands the whole launch list to native code in one call (`smtl_run_plan`).
torch is used only for device memory and the current stream.
"""
import ctypes as C

import torch

from . import _lib as L

F32 = torch.float32

# 16-bit operand format of every op built from here on.  fp16 (saturating, fp32 accumulate) is the default: it is what
# keeps the end-to-end maps within 1e-2 relative L2 of the fp32 oracle; bf16 runs at the same tensor rate with
# ~8x the rounding error (see DESIGN.md "Numerics").
PREC = {"fmt": L.FMT_F16, "dtype": torch.float16, "name": "fp16"}


def set_precision(name):
    if name == "fp16":
        PREC.update(fmt=L.FMT_F16, dtype=torch.float16, name="fp16")
    elif name == "bf16":
        PREC.update(fmt=L.FMT_BF16, dtype=torch.bfloat16, name="bf16")
    else:
        raise ValueError(f"unknown precision {name!r}")


def h16():
    return PREC["dtype"]


class _H16:
    """compares equal to the active 16-bit dtype (so `t.dtype == BF16` follows set_precision)."""

    def __eq__(self, other):
        return other == PREC["dtype"]

    def __ne__(self, other):
        return other != PREC["dtype"]

    __hash__ = None


BF16 = _H16()


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_RUNNERS = {
    L.OP_GEMM: "smtl_gemm_run", L.OP_FATTN: "smtl_fattn_run", L.OP_SOFTMAX: "smtl_softmax_run",
    L.OP_XATTN: "smtl_xattn_run", L.OP_TASKATTN: "smtl_taskattn_run",
    L.OP_LN: "smtl_ln_run", L.OP_UPSAMPLE: "smtl_upsample_run", L.OP_IM2COL: "smtl_im2col_run",
    L.OP_RGBPREP: "smtl_rgbprep_run", L.OP_UNETIN: "smtl_unetin_run", L.OP_TASKMAP: "smtl_taskmap_run",
    L.OP_CHANMIX: "smtl_chanmix_run", L.OP_GNAPPLY: "smtl_gnapply_run", L.OP_MEMSET: "smtl_memset_run",
    L.OP_GNFINALIZE: "smtl_gnfinalize_run", L.OP_LSQSUMS: "smtl_lsqsums_run", L.OP_CONFUSION: "smtl_confusion_run",
    L.OP_RGBSTEM: "smtl_rgbstem_run", L.OP_HEADGATHER: "smtl_headgather_run", L.OP_XATTNF: "smtl_xattnf_run",
}


class Op:
    __slots__ = ("kind", "struct", "keep", "flops", "flops_exec", "name", "bytes", "outs16")

    def __init__(self, kind, struct, keep=(), flops=0, name="", nbytes=0, outs16=()):
        self.kind, self.struct, self.keep, self.flops, self.name = kind, struct, keep, flops, name
        self.outs16 = tuple(t for t in outs16 if t is not None)   # 16-bit outputs (StableMTLEngine.audit_range)
        self.flops_exec = flops      # FLOPs the tensor pipe executes (2*m*n*k of the issued GEMM); `flops` is ALGORITHMIC
        self.bytes = nbytes          # algorithmic HBM bytes of a bandwidth-bound op (0: not accounted)

    def run(self):
        fn = getattr(L.lib, _RUNNERS[self.kind])
        L.check(fn(C.byref(self.struct), _stream()), _RUNNERS[self.kind])
        return self


class Plan:
    """An ordered launch list executed natively (one ABI crossing per pass)."""

    def __init__(self):
        self.ops = []
        self._arr = None

    def add(self, op):
        self.ops.append(op)
        self._arr = None
        return op

    def extend(self, ops):
        for o in ops:
            self.add(o)

    def finalize(self):
        arr = (L.OpRef * len(self.ops))()
        for i, o in enumerate(self.ops):
            arr[i].kind = o.kind
            arr[i].op = C.cast(C.pointer(o.struct), C.c_void_p)
        self._arr = arr
        return self

    @property
    def launches(self):
        if self._arr is None:
            self.finalize()
        return L.lib.smtl_plan_launches(self._arr, len(self.ops))

    @property
    def flops(self):
        return sum(o.flops for o in self.ops)

    def run(self):
        if self._arr is None:
            self.finalize()
        L.check(L.lib.smtl_run_plan(self._arr, len(self.ops), _stream()), "smtl_run_plan")


# ------------------------------------------------------------------------------------------------- GEMM / conv
STATS_REPLICAS = 8      # copies of a statistics buffer the producing GEMM's CTAs spread their atomics over
SILU_MODE = 1           # gn_apply(silu=True): 1 = one tanh.approx per element, 2 = ex2 + rcp (~2 ulp)


def new_stats(images, channels, device, replicas=None):
    """int64 [replicas, images, channels, 4]: fixed-point (sum_lo, sum_hi, sq_lo, sq_hi) cells a GEMM epilogue adds
    into with integer atomics (exact, order-independent); must be zeroed before the producer runs."""
    return torch.zeros(replicas or STATS_REPLICAS, images, channels, 4, device=device, dtype=torch.int64)


def stats_encode(values):
    """float64 [..., 2] (sum, sum of squares) -> int64 fixed-point cells [..., 4] the way a producer would add them"""
    v = values.double()
    fine = v.abs() < 2.0 ** 12
    lo = torch.where(fine, torch.round(v * 2.0 ** 32), torch.zeros_like(v)).to(torch.int64)
    hi = torch.where(fine, torch.zeros_like(v), torch.round(v * 2.0 ** 8)).to(torch.int64)
    return torch.stack([lo[..., 0], hi[..., 0], lo[..., 1], hi[..., 1]], -1)


def stats_values(stats):
    """float64 [images, channels, 2] (sum, sum of squares) held by a statistics buffer"""
    s = stats.double()
    s = torch.stack([s[..., 0] * 2.0 ** -32 + s[..., 1] * 2.0 ** -8, s[..., 2] * 2.0 ** -32 + s[..., 3] * 2.0 ** -8], -1)
    return s.sum(0)


def gemm(a0, b, *, m=None, k=None, n=None, a1=None, segs=None, bias=None, bias_per_row=False, act=L.ACT_NONE,
         res1=None, res2=None, out_f32=None, out_bf16=None, aux_bf16=None, rowmap=L.ROWMAP_IDENTITY, img_hw=None,
         block_n=0, name="gemm", stats=None, stats_rows_per_image=0, cta_group=0, up_parity=0, group_rows=0, tile_order=0,
         stats_group=0):
    """D = A @ B^T (+ fused epilogue).  a0/a1: bf16 [rows, cols] (row stride = stride(0)); b: bf16 [n, k].
    segs: list of (row_shift, kblocks, src, a_col0)."""
    assert a0.dtype == BF16 and b.dtype == BF16 and a0.stride(-1) == 1 and b.stride(-1) == 1
    g = L.GemmArgs()
    g.fmt16 = PREC["fmt"]
    g.a0, g.a0_rows, g.a0_cols, g.a0_ld = a0.data_ptr(), a0.shape[0], a0.shape[1], a0.stride(0)
    if a1 is not None:
        assert a1.dtype == BF16 and a1.stride(-1) == 1
        g.a1, g.a1_rows, g.a1_cols, g.a1_ld = a1.data_ptr(), a1.shape[0], a1.shape[1], a1.stride(0)
    g.b, g.ldb = b.data_ptr(), b.stride(0)
    g.n = b.shape[0] if n is None else n
    g.k = b.shape[1] if k is None else k
    g.m = a0.shape[0] if m is None else m
    if segs:
        g.nseg = len(segs)
        for i, (shift, kb, src, col0) in enumerate(segs):
            g.seg[i].row_shift, g.seg[i].kblocks, g.seg[i].src, g.seg[i].a_col0 = shift, kb, src, col0
    if bias is not None:
        assert bias.dtype == F32
        g.bias = bias.data_ptr()
    g.bias_per_row = int(bias_per_row)
    g.act = act
    n_out = g.n // 2 if act == L.ACT_GEGLU else g.n
    res_dt = None
    for r in (res1, res2):
        if r is not None:
            assert (r.dtype == F32 or r.dtype == BF16) and r.stride(-1) == 1
            assert res_dt is None or (r.dtype == res_dt and r.stride(0) == g.ldres), "res1/res2 must share dtype and ld"
            res_dt = r.dtype
            g.ldres = r.stride(0)
    g.res_fmt16 = int(res_dt is not None and res_dt != F32)
    if stats is not None:
        assert stats.dtype == torch.int64 and stats.dim() == 4 and stats.shape[3] == 4 and stats.is_contiguous()
        assert stats.shape[2] == n_out and stats_rows_per_image > 0
        g.stats, g.stats_replicas, g.stats_images = stats.data_ptr(), stats.shape[0], stats.shape[1]
        g.stats_rows_per_image = stats_rows_per_image
        g.stats_group = stats_group
    g.res1, g.res2 = _ptr(res1), _ptr(res2)
    outs = [t for t in (out_f32, out_bf16) if t is not None]
    if out_f32 is not None:
        assert out_f32.dtype == F32 and out_f32.stride(-1) == 1
    if out_bf16 is not None:
        assert out_bf16.dtype == BF16 and out_bf16.stride(-1) == 1
    if len(outs) == 2:
        assert out_f32.stride(0) == out_bf16.stride(0)
    if outs:
        g.ldc = outs[0].stride(0)
        assert outs[0].shape[-1] >= n_out or outs[0].stride(0) >= n_out
    g.out_f32, g.out_bf16 = _ptr(out_f32), _ptr(out_bf16)
    if aux_bf16 is not None:
        assert aux_bf16.dtype == BF16
        g.aux_bf16, g.ld_aux = aux_bf16.data_ptr(), aux_bf16.stride(0)
    g.rowmap = rowmap
    if img_hw is not None:
        g.img_h, g.img_w = img_hw
    g.block_n = block_n
    g.cta_group = cta_group
    g.up_parity = up_parity
    g.group_rows = group_rows
    g.tile_order = tile_order
    op = L.GemmOp()
    L.check(L.lib.smtl_gemm_plan(C.byref(g), C.byref(op)), "smtl_gemm_plan")
    if group_rows:
        assert n is not None and b.shape[0] == (int(g.m) // group_rows) * n, "grouped GEMM: b is [groups * n, k], pass n"
        assert bias is None or bias.numel() == b.shape[0]
    flops = 2 * int(g.m) * int(g.n) * int(g.k)
    return Op(L.OP_GEMM, op, (a0, a1, b, bias, res1, res2, out_f32, out_bf16, aux_bf16, stats), flops, name,
              outs16=(out_bf16, aux_bf16))


def conv3x3_segs(cin, w, shortcut_cin=0):
    """K segments of a 3x3/s1/p1 conv over the padded layout (+ optional fused 1x1 shortcut from source 1)."""
    assert cin % 64 == 0
    wp = w + 2
    segs = [((ky - 1) * wp + (kx - 1), cin // 64, 0, 0) for ky in range(3) for kx in range(3)]
    if shortcut_cin:
        assert shortcut_cin % 64 == 0
        segs.append((0, shortcut_cin // 64, 1, 0))
    return segs


def conv3x3(a_pad, wmat, batch, h, w, *, a_short=None, name="conv3x3", pad_out=False, **epi):
    """a_pad: bf16 [batch*(h+2)*(w+2), cin] zero-halo layout; wmat: bf16 [cout, 9*cin (+ c_short)].
    pad_out: the output (and residuals) are in the padded layout as well (zero halo written)."""
    cin = a_pad.shape[1]
    cs = 0 if a_short is None else a_short.shape[1]
    op = gemm(a_pad, wmat, m=batch * (h + 2) * (w + 2), a1=a_short, segs=conv3x3_segs(cin, w, cs),
              rowmap=L.ROWMAP_PAD_KEEP if pad_out else L.ROWMAP_CONV_PAD, img_hw=(h, w), name=name, **epi)
    op.flops = 2 * batch * h * w * wmat.shape[0] * wmat.shape[1]   # algorithmic (halo rows excluded); flops_exec keeps them
    return op


def up2x_weight_matrices(w):
    """[Cout, Cin, 3, 3] filter of "nearest 2x upsample, then 3x3/p1 conv" -> the four per-output-parity 2x2 filters on
    the LOW-resolution input, each as a [Cout, 4*Cin] matrix (tap = dy*2 + dx), summed in fp32.
    Output row 2y+py reads low-res rows {y-1: W[0], y: W[1]+W[2]} for py = 0 and {y: W[0]+W[1], y+1: W[2]} for py = 1
    (same for columns): 4 taps instead of 9, i.e. 2.25x fewer FLOPs for the identical result."""
    w = w.float()
    rows = {0: [w[:, :, 0], w[:, :, 1] + w[:, :, 2]], 1: [w[:, :, 0] + w[:, :, 1], w[:, :, 2]]}      # [Cout,Cin,3(kx)]
    mats = []
    for py in (0, 1):
        for px in (0, 1):
            taps = []
            for dy in (0, 1):
                r = rows[py][dy]
                cols = [r[:, :, 0], r[:, :, 1] + r[:, :, 2]] if px == 0 else [r[:, :, 0] + r[:, :, 1], r[:, :, 2]]
                taps += cols
            mats.append(torch.cat(taps, dim=1).contiguous())              # [Cout, 4*Cin], tap-major
    return mats


def conv_up2x(a_pad, wmats, batch, h, w, *, name="up2x_conv", pad_out=False, **epi):
    """nearest-2x upsample + 3x3 conv as four 2x2 implicit-GEMM convs over the padded LOW-res map a_pad
    [batch*(h+2)*(w+2), cin]; outputs land in the compact [batch*2h*2w, cout] map.  Returns the four ops."""
    cin = a_pad.shape[1]
    assert cin % 64 == 0
    wp = w + 2
    out = []
    for py in (0, 1):
        for px in (0, 1):
            ys = (-1, 0) if py == 0 else (0, 1)
            xs = (-1, 0) if px == 0 else (0, 1)
            segs = [(dy * wp + dx, cin // 64, 0, 0) for dy in ys for dx in xs]
            op = gemm(a_pad, wmats[py * 2 + px], m=batch * (h + 2) * wp, segs=segs,
                      rowmap=L.ROWMAP_UP2_PAD if pad_out else L.ROWMAP_CONV_PAD_UP2,
                      img_hw=(h, w), up_parity=py * 2 + px, name=name, **epi)
            # algorithmic: 1/4 of the 3x3 conv on the upsampled map; flops_exec = the 4-tap GEMM that actually runs
            op.flops = 2 * batch * h * w * wmats[0].shape[0] * 9 * cin
            out.append(op)
    return out


# ------------------------------------------------------------------------------------------------- attention
def flash_attn(qkv, batch, ntok, heads, out, q_col0, k_col0, v_col0, scale=0.125, head_dim=64):
    """head_dim 64 (UNet self-attention) or 512 with one head (VAE mid-block attention)"""
    a = L.FattnArgs()
    a.head_dim = head_dim
    a.fmt16 = PREC["fmt"]
    a.qkv, a.ld = qkv.data_ptr(), qkv.stride(0)
    a.q_col0, a.k_col0, a.v_col0 = q_col0, k_col0, v_col0
    a.batch, a.ntok, a.heads = batch, ntok, heads
    a.out_bf16, a.ldo, a.scale = out.data_ptr(), out.stride(0), scale
    assert qkv.dtype == BF16 and out.dtype == BF16 and qkv.shape[0] == batch * ntok
    op = L.FattnOp()
    L.check(L.lib.smtl_fattn_plan(C.byref(a), C.byref(op)), "smtl_fattn_plan")
    o = Op(L.OP_FATTN, op, (qkv, out), 4 * batch * heads * ntok * ntok * head_dim, "flash_attn", outs16=(out,))
    if head_dim == 512:
        o.flops_exec = o.flops * 3 // 2          # two passes over the keys: QK^T runs twice
    return o


def softmax_rows(s, p, scale):
    a = L.SoftmaxArgs()
    a.fmt16 = PREC["fmt"]
    a.s, a.rows, a.n, a.lds, a.scale = s.data_ptr(), s.shape[0], s.shape[1], s.stride(0), scale
    a.p_bf16, a.ldp = p.data_ptr(), p.stride(0)
    assert s.dtype == F32 and p.dtype == BF16
    return Op(L.OP_SOFTMAX, a, (s, p), 0, "softmax", outs16=(p,))


def xattn(q, kc, vc, ntok, task_of_group, rows_per_group, heads, out, scale=0.125):
    a = L.XattnArgs()
    a.fmt16 = PREC["fmt"]
    a.q_bf16, a.ldq, a.rows, a.heads = q.data_ptr(), q.stride(0), q.shape[0], heads
    a.kc, a.vc = kc.data_ptr(), vc.data_ptr()
    assert kc.dtype == F32 and vc.dtype == F32 and kc.shape[1] <= L.MAX_XATTN_TOKENS and kc.shape[2] == heads * 64
    a.ntok_pad = kc.shape[1]
    for i in range(L.MAX_TASKS):
        a.ntok[i] = ntok[i] if i < len(ntok) else 0
        a.task_of_group[i] = task_of_group[i] if i < len(task_of_group) else 0
    a.rows_per_group = rows_per_group
    a.out_bf16, a.ldo, a.scale = out.data_ptr(), out.stride(0), scale
    return Op(L.OP_XATTN, a, (q, kc, vc, out), 0, "xattn", outs16=(out,))


def xattn_fused_supported(heads, ntok_pad):
    return bool(L.lib.smtl_xattnf_supported(heads, ntok_pad))


def xattn_tables(a0, gamma2, beta2, bm, ntok, ntok_pad):
    """Host-side packing of the collapsed cross-attention (smtl_xattnf_args): a0 [T, H, ntp, C] = Wq_head^T k / 8 and
    bm [T, H, ntp, C] = Wo[:, head] v, fp32 -> (ap [T, VP, C] 16-bit, ca [T, VP] fp32, bmt [T, C, VP] 16-bit) with the
    vectors padded to VP = a multiple of 16 and the padding tokens masked (zero vector, -inf constant); bmt's vector axis
    is stored in the kernel's fragment order (see `xattn_unpermute`)."""
    T, H, n, C = a0.shape
    assert n == ntok_pad and bm.shape == a0.shape
    V = H * n
    VP = (V + 15) // 16 * 16
    valid = (torch.arange(n)[None, :] < torch.as_tensor(ntok)[:, None])[:, None, :].expand(T, H, n)          # [T, H, n]
    ap = torch.zeros(T, VP, C)
    ap[:, :V] = (a0 * gamma2).masked_fill(~valid[..., None], 0.0).reshape(T, V, C)
    ca = torch.full((T, VP), float("-inf"))
    ca[:, :V] = (a0 * beta2).sum(-1).masked_fill(~valid, float("-inf")).reshape(T, V)
    bmt = torch.zeros(T, C, VP)
    bmt[:, :, :V] = bm.reshape(T, V, C).transpose(1, 2)
    # vector axis in MMA fragment order: stored position 16 b + 4 q + e holds vector 16 b + (2q, 2q+1, 2q+8, 2q+9)[e], so
    # that a lane's four contraction entries of a 16-block are one 8-byte load
    perm = torch.tensor([16 * b + 2 * q + o for b in range(VP // 16) for q in range(4) for o in (0, 1, 8, 9)])
    bmt = bmt[:, :, perm]
    return ap.to(h16()).contiguous(), ca.contiguous(), bmt.to(h16()).contiguous()


def xattn_unpermute(bmt):
    """bmt [..., VP] as packed by xattn_tables -> natural vector order (tests / inspection)"""
    vp = bmt.shape[-1]
    perm = torch.tensor([16 * b + 2 * q + o for b in range(vp // 16) for q in range(4) for o in (0, 1, 8, 9)], device=bmt.device)
    out = torch.empty_like(bmt)
    out[..., perm] = bmt
    return out


def xattn_fused(hs, ap, ca, bmt, bo, gamma3, beta3, task_of_group, rows_per_group, heads, ntok_pad, out, eps2=1e-5, eps3=1e-5):
    """hs += attn2(LayerNorm2(hs), text[task]); out = LayerNorm3(hs) -- the collapsed cross-attention (see the header)."""
    a = L.XattnFArgs()
    c = heads * 64
    ntask, vp, _ = ap.shape
    ntp = ntok_pad
    v = heads * ntp
    assert vp == (v + 15) // 16 * 16
    assert hs.dtype == F32 and hs.shape[1] == c and hs.stride(1) == 1 and out.dtype == BF16 and out.shape == hs.shape
    assert ap.dtype == BF16 and bmt.dtype == BF16 and ap.shape == (ntask, vp, c) and bmt.shape == (ntask, c, vp)
    assert ap.is_contiguous() and bmt.is_contiguous() and ca.dtype == F32 and ca.shape == (ntask, vp) and ca.is_contiguous()
    assert hs.shape[0] % rows_per_group == 0 and len(task_of_group) == hs.shape[0] // rows_per_group
    a.hs, a.ldh, a.heads, a.rows, a.rows_per_group = hs.data_ptr(), hs.stride(0), heads, hs.shape[0], rows_per_group
    for i in range(L.MAX_TASKS):
        a.task_of_group[i] = max(task_of_group[i], 0) if i < len(task_of_group) else 0
    a.ntok_pad, a.fmt16 = ntp, PREC["fmt"]
    a.ap, a.ca, a.bmt, a.bo = ap.data_ptr(), ca.data_ptr(), bmt.data_ptr(), bo.data_ptr()
    a.gamma3, a.beta3, a.out_bf16, a.ldo, a.eps2, a.eps3 = gamma3.data_ptr(), beta3.data_ptr(), out.data_ptr(), out.stride(0), eps2, eps3
    op = Op(L.OP_XATTNF, a, (hs, ap, ca, bmt, bo, gamma3, beta3, out), 4 * hs.shape[0] * c * c, "xattn_fused",
            hs.numel() * 8 + out.numel() * 2, outs16=(out,))
    op.flops_exec = 4 * hs.shape[0] * c * vp           # what runs: two [rows x C] x [C x VP] skinny GEMMs on mma.sync
    return op


def task_attn(q, k, v, out, c, nheads, main_tasks, src_tasks, rows_per_group, exclude_self=True):
    a = L.TaskAttnArgs()
    a.fmt16 = PREC["fmt"]
    a.q_bf16, a.k_bf16, a.v_bf16, a.out_bf16 = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.c, a.nheads, a.n_main, a.n_src = c, nheads, len(main_tasks), len(src_tasks)
    a.rows_per_group = rows_per_group
    for i, t in enumerate(main_tasks):
        a.main_task[i] = t
    for i, t in enumerate(src_tasks):
        a.src_task[i] = t
    a.exclude_self = int(exclude_self)
    a.scale = float((c // nheads) ** -0.5)
    assert q.dtype == BF16 and k.dtype == BF16 and v.dtype == BF16 and out.dtype == BF16
    assert q.shape == (len(main_tasks) * rows_per_group, c) and k.shape == (len(src_tasks) * rows_per_group, c)
    return Op(L.OP_TASKATTN, a, (q, k, v, out), 0, "task_attn", outs16=(out,))


# ------------------------------------------------------------------------------------------------- norms
def gn_apply(x0, stats0, batch, h, w, gamma, beta, out, *, x1=None, stats1=None, eps, silu, pad_out, raw=None,
             groups=32, x_padded=False):
    """GroupNorm(+SiLU) of the (virtually concatenated) compact map [x0 | x1] using the per-(image, channel) sums the
    producing GEMMs left in stats0 / stats1; writes the 16-bit operand of the next conv (padded) or GEMM (compact)."""
    a = L.GnApplyArgs()
    a.fmt16 = PREC["fmt"]
    a.x0, a.c0 = x0.data_ptr(), x0.shape[-1]
    a.x_fmt16 = int(x0.dtype != F32)
    assert x0.is_contiguous() and (x0.dtype == F32 or x0.dtype == BF16)
    assert stats0.shape[1:] == (batch, x0.shape[-1], 4) and stats0.dtype == torch.int64, (stats0.shape, batch, x0.shape)
    a.stats0, a.stats_replicas = stats0.data_ptr(), stats0.shape[0]
    if x1 is not None:
        assert x1.dtype == x0.dtype and x1.is_contiguous() and stats1.shape == (stats0.shape[0], batch, x1.shape[-1], 4)
        a.x1, a.c1, a.stats1 = x1.data_ptr(), x1.shape[-1], stats1.data_ptr()
    a.batch, a.h, a.w, a.groups, a.eps = batch, h, w, groups, eps
    a.x_padded = int(x_padded)
    assert x0.shape[0] == batch * ((h + 2) * (w + 2) if x_padded else h * w)
    a.gamma, a.beta = gamma.data_ptr(), beta.data_ptr()
    a.silu, a.pad_out = (SILU_MODE if silu is True else int(silu)), int(pad_out)
    a.out_bf16, a.raw_bf16 = out.data_ptr(), _ptr(raw)
    assert out.dtype == BF16
    # algorithmic bytes: every interior element read once, every output element (halo included) written once
    nb = batch * h * w * (x0.shape[-1] * x0.element_size() + (0 if x1 is None else x1.shape[-1] * x1.element_size()))
    nb += out.numel() * out.element_size() + (0 if raw is None else raw.numel() * raw.element_size())
    return Op(L.OP_GNAPPLY, a, (x0, x1, stats0, stats1, gamma, beta, out, raw), 0, "gn_apply", nb, outs16=(out, raw))


def gn_finalize(stats, batch, pixels, gamma, beta, ss, *, eps, groups=32):
    """producer-side channel sums -> fp32 [batch, C, 2] (scale, shift): GroupNorm as a per-image affine map"""
    a = L.GnFinalizeArgs()
    c = stats.shape[2]
    assert stats.shape[1:] == (batch, c, 4) and stats.dtype == torch.int64 and ss.shape == (batch, c, 2) and ss.dtype == F32 and ss.is_contiguous()
    a.stats, a.stats_replicas, a.batch, a.c, a.groups = stats.data_ptr(), stats.shape[0], batch, c, groups
    a.pixels, a.eps = pixels, eps
    a.gamma, a.beta, a.ss = gamma.data_ptr(), beta.data_ptr(), ss.data_ptr()
    return Op(L.OP_GNFINALIZE, a, (stats, gamma, beta, ss), 0, "gn_finalize")


def memset_zero(t):
    a = L.MemsetArgs()
    a.ptr, a.bytes, a.value = t.data_ptr(), t.numel() * t.element_size(), 0
    assert t.is_contiguous()
    return Op(L.OP_MEMSET, a, (t,), 0, "memset")


def layer_norm(x, gamma0, beta0, out0, *, gamma1=None, beta1=None, out1=None, rows_per_group=None, eps=1e-5):
    a = L.LnArgs()
    a.fmt16 = PREC["fmt"]
    a.x, a.x_is_bf16, a.c, a.ldx = x.data_ptr(), int(x.dtype != F32), x.shape[1], x.stride(0)
    a.rows, a.eps = x.shape[0], eps
    a.rows_per_group = x.shape[0] if rows_per_group is None else rows_per_group
    a.gamma0, a.beta0, a.out0 = gamma0.data_ptr(), beta0.data_ptr(), out0.data_ptr()
    a.gamma1, a.beta1, a.out1 = _ptr(gamma1), _ptr(beta1), _ptr(out1)
    a.ldo = out0.stride(0)
    assert out0.dtype == BF16 and gamma0.dtype == F32
    nb = x.numel() * x.element_size() + out0.numel() * out0.element_size() + (0 if out1 is None else out1.numel() * out1.element_size())
    return Op(L.OP_LN, a, (x, gamma0, beta0, out0, gamma1, beta1, out1), 0, "layer_norm", nb, outs16=(out0, out1))


# ------------------------------------------------------------------------------------------------- data movement
def upsample_pad(x, batch, h, w, oh, ow, out):
    a = L.UpsampleArgs()
    a.fmt16 = PREC["fmt"]
    a.x, a.batch, a.h, a.w, a.c, a.oh, a.ow, a.out_bf16 = x.data_ptr(), batch, h, w, x.shape[-1], oh, ow, out.data_ptr()
    a.x_fmt16 = int(x.dtype != F32)
    assert (x.dtype == F32 or x.dtype == BF16) and out.dtype == BF16
    return Op(L.OP_UPSAMPLE, a, (x, out), 0, "upsample", outs16=(out,))


def im2col(x, batch, h, w, out, *, stride, pad_t, pad_l, oh, ow):
    a = L.Im2colArgs()
    a.fmt16 = PREC["fmt"]
    a.x, a.batch, a.h, a.w, a.c = x.data_ptr(), batch, h, w, x.shape[-1]
    a.stride, a.pad_t, a.pad_l, a.oh, a.ow, a.kpad = stride, pad_t, pad_l, oh, ow, out.shape[-1]
    a.out_bf16 = out.data_ptr()
    a.x_fmt16 = int(x.dtype != F32)
    assert (x.dtype == F32 or x.dtype == BF16) and out.dtype == BF16 and out.is_contiguous()
    return Op(L.OP_IM2COL, a, (x, out), 0, "im2col", outs16=(out,))


def rgb_prep(rgb_nchw, out_nhwc, normalized=False):
    """normalized=True: the input already is rgb / 255 * 2 - 1 (encode_rgb's argument); layout change only"""
    a = L.RgbprepArgs()
    b, _, h, w = rgb_nchw.shape
    a.rgb_nchw, a.batch, a.h, a.w, a.out_nhwc = rgb_nchw.data_ptr(), b, h, w, out_nhwc.data_ptr()
    a.src_u8 = 2 if normalized else int(rgb_nchw.dtype == torch.uint8)
    assert not (normalized and rgb_nchw.dtype != F32)
    assert rgb_nchw.dtype in (F32, torch.uint8) and rgb_nchw.is_contiguous() and out_nhwc.dtype == F32
    return Op(L.OP_RGBPREP, a, (rgb_nchw, out_nhwc), 0, "rgb_prep")


def rgb_stem(rgb_nchw, col, normalized=False):
    """[0,255] (or already normalised) NCHW rgb -> the 16-bit im2col operand [batch*h*w, 64] of the VAE encoder's stem"""
    a = L.RgbstemArgs()
    b, c, h, w = rgb_nchw.shape
    assert c == 3 and rgb_nchw.is_contiguous() and rgb_nchw.dtype in (F32, torch.uint8) and not (normalized and rgb_nchw.dtype != F32)
    assert col.dtype == BF16 and col.is_contiguous() and col.shape == (b * h * w, 64)
    a.rgb_nchw, a.batch, a.h, a.w = rgb_nchw.data_ptr(), b, h, w
    a.src_mode = 2 if normalized else int(rgb_nchw.dtype == torch.uint8)
    a.out_bf16, a.fmt16 = col.data_ptr(), PREC["fmt"]
    return Op(L.OP_RGBSTEM, a, (rgb_nchw, col), 0, "rgb_stem", rgb_nchw.numel() * rgb_nchw.element_size() + col.numel() * 2)


def unet_input(latents, first_img, second_img, hw, out):
    a = L.UnetinArgs()
    a.latents, a.first_img, a.second_img = latents.data_ptr(), first_img.data_ptr(), second_img.data_ptr()
    a.out_images, a.hw, a.out = first_img.numel(), hw, out.data_ptr()
    assert first_img.dtype == torch.int32 and second_img.dtype == torch.int32 and latents.dtype == F32
    return Op(L.OP_UNETIN, a, (latents, first_img, second_img, out), 0, "unet_input")


def chan_mix(x, w, b, y):
    a = L.ChanmixArgs()
    a.x, a.rows, a.cin, a.cout = x.data_ptr(), x.numel() // w.shape[1], w.shape[1], w.shape[0]
    a.w, a.b, a.y = w.data_ptr(), _ptr(b), y.data_ptr()
    assert x.dtype == F32 and w.dtype == F32 and y.dtype == F32
    return Op(L.OP_CHANMIX, a, (x, w, b, y), 0, "chan_mix")


def head_weight_matrix(w, npad=16):
    """[cout <= 4, Cin, 3, 3] filter of a narrow-output conv -> [npad, 3 * Cin]: row kx * cout + co, column ky * Cin + c
    (the kx taps folded into N, the ky taps stay K segments)"""
    co, ci, kh, kw = w.shape
    assert kh == 3 and kw == 3 and 3 * co <= npad
    m = torch.zeros(npad, 3 * ci, dtype=w.dtype, device=w.device)
    m[: 3 * co] = w.permute(3, 0, 2, 1).reshape(3 * co, 3 * ci)          # [kx, co, ky, c]
    return m


def conv_head(a_pad, wfold, bias, batch, h, w, cout, partial, out, name="conv_head"):
    """3x3 / pad-1 conv with cout <= 4 channels: a 3-segment (ky) implicit GEMM over the padded map with the kx taps on
    the N side (fp32 partials [rows, 16]) followed by the three-tap horizontal gather.  Returns the two ops."""
    cin = a_pad.shape[1]
    assert cin % 64 == 0 and wfold.shape[1] == 3 * cin
    assert partial.dtype == F32 and partial.shape == (a_pad.shape[0], wfold.shape[0]) and out.dtype == F32
    assert a_pad.shape[0] == batch * (h + 2) * (w + 2) and out.shape == (batch * h * w, cout)
    segs = [((ky - 1) * (w + 2), cin // 64, 0, 0) for ky in range(3)]
    g = gemm(a_pad, wfold, segs=segs, out_f32=partial, name=name)
    g.flops = 2 * batch * h * w * cout * 9 * a_pad.shape[1]                 # algorithmic flops of the conv
    a = L.HeadGatherArgs()
    a.partial, a.ldp, a.batch, a.h, a.w, a.cout = partial.data_ptr(), partial.stride(0), batch, h, w, cout
    a.bias, a.out = _ptr(bias), out.data_ptr()
    nb = batch * h * w * (3 * cout + cout) * 4
    return [g, Op(L.OP_HEADGATHER, a, (partial, bias, out), 0, name + ".gather", nb)]


def task_map(x, batch, hw, mode, *, out_clipped=None, out_post=None, out_ids=None, palette=None):
    a = L.TaskmapArgs()
    a.x, a.batch, a.hw, a.mode = x.data_ptr(), batch, hw, mode
    assert x.is_contiguous() and x.numel() >= batch * hw * 3, "task_map: the decoder output is smaller than batch * hw pixels"
    a.out_clipped, a.out_post, a.out_ids = _ptr(out_clipped), _ptr(out_post), _ptr(out_ids)
    if palette is not None:
        a.palette, a.npalette = palette.data_ptr(), palette.shape[0]
    assert x.dtype == F32
    return Op(L.OP_TASKMAP, a, (x, out_clipped, out_post, out_ids, palette), 0, "task_map")


# ------------------------------------------------------------------------------------------------- evaluation pre-reductions
def lsq_sums(pred, gt, valid, sums):
    """accumulates fp64 [batch, 5] = (n, sum p, sum g, sum p*p, sum p*g) over the valid pixels of fp32 [batch, hw] maps"""
    a = L.LsqSumsArgs()
    batch, hw = pred.shape[0], pred[0].numel()
    assert pred.dtype == F32 and gt.dtype == F32 and pred.is_contiguous() and gt.is_contiguous() and gt.shape == pred.shape
    assert sums.dtype == torch.float64 and sums.shape == (batch, 5) and sums.is_contiguous()
    assert valid is None or (valid.dtype in (torch.uint8, torch.bool) and valid.is_contiguous() and valid.numel() == pred.numel())
    a.pred, a.gt, a.valid, a.batch, a.hw, a.sums = pred.data_ptr(), gt.data_ptr(), _ptr(valid), batch, hw, sums.data_ptr()
    return Op(L.OP_LSQSUMS, a, (pred, gt, valid, sums), 0, "lsq_sums")


def confusion(label_true, label_pred, valid, hist, n_classes):
    """accumulates int64 [n_classes^2 + 1]: the confusion matrix (row = true class) + count of out-of-range predictions"""
    a = L.ConfusionArgs()
    assert label_true.dtype == torch.int64 and label_pred.dtype == torch.int64 and label_true.shape == label_pred.shape
    assert label_true.is_contiguous() and label_pred.is_contiguous()
    assert hist.dtype == torch.int64 and hist.numel() == n_classes * n_classes + 1 and hist.is_contiguous()
    assert valid is None or (valid.dtype in (torch.uint8, torch.bool) and valid.is_contiguous() and valid.numel() == label_true.numel())
    a.label_true, a.label_pred, a.valid = label_true.data_ptr(), label_pred.data_ptr(), _ptr(valid)
    a.n, a.n_classes, a.hist = label_true.numel(), n_classes, hist.data_ptr()
    return Op(L.OP_CONFUSION, a, (label_true, label_pred, valid, hist), 0, "confusion")
