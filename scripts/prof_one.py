"""Runs ONE representative kernel a few times (for `ncu -k ... -s N -c 1`): python scripts/prof_one.py <case>"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablemtl_b200 import ops, _lib as L  # noqa: E402
from stablemtl_b200.weights import interleave_geglu  # noqa: E402

DEV = "cuda"
case = sys.argv[1]
rb = lambda *s: (torch.randn(*s, device=DEV) * 0.5).to(ops.h16())
if case == "ff1":            # GEGLU feed-forward, UNet level 0, 112 images
    m, c = 112 * 4800, 320
    a, w, bias = rb(m, c), rb(8 * c, c), torch.randn(8 * c, device=DEV)
    wi, bi = interleave_geglu(w, bias)
    out = torch.empty(m, 4 * c, device=DEV, dtype=ops.h16())
    op = ops.gemm(a, wi, bias=bi, act=L.ACT_GEGLU, out_bf16=out)
elif case == "lin":          # proj_out-like: K = N = 320, 16-bit residual + stats
    m, c = 112 * 4800, 320
    a, w, bias, res = rb(m, c), rb(c, c), torch.randn(c, device=DEV), rb(m, c)
    out = torch.empty(m, c, device=DEV, dtype=ops.h16())
    st = ops.new_stats(112, c, DEV)
    op = ops.gemm(a, w, bias=bias, res1=res, out_bf16=out, stats=st, stats_rows_per_image=4800)
elif case == "lin0":         # bare K = N = 320 linear, 16-bit out only
    m, c = 112 * 4800, 320
    a, w = rb(m, c), rb(c, c)
    out = torch.empty(m, c, device=DEV, dtype=ops.h16())
    op = ops.gemm(a, w, out_bf16=out)
elif case == "conv128":      # VAE full-resolution 128 -> 128 conv, 8 images
    b, h, wd, c = 8, 480, 640, 128
    a, w = rb(b * (h + 2) * (wd + 2), c), rb(c, 9 * c)
    out = torch.empty(b * h * wd, c, device=DEV, dtype=ops.h16())
    st = ops.new_stats(b, c, DEV)
    op = ops.conv3x3(a, w, b, h, wd, bias=torch.zeros(c, device=DEV), out_bf16=out, stats=st, stats_rows_per_image=h * wd)
elif case == "conv512":
    b, h, wd, c = 8, 120, 160, 512
    a, w = rb(b * (h + 2) * (wd + 2), c), rb(c, 9 * c)
    out = torch.empty(b * h * wd, c, device=DEV, dtype=ops.h16())
    st = ops.new_stats(b, c, DEV)
    op = ops.conv3x3(a, w, b, h, wd, bias=torch.zeros(c, device=DEV), out_bf16=out, stats=st, stats_rows_per_image=h * wd)
elif case == "conv512cg2":
    b, h, wd, c = 8, 120, 160, 512
    a, w = rb(b * (h + 2) * (wd + 2), c), rb(c, 9 * c)
    out = torch.empty(b * h * wd, c, device=DEV, dtype=ops.h16())
    st = ops.new_stats(b, c, DEV)
    op = ops.conv3x3(a, w, b, h, wd, bias=torch.zeros(c, device=DEV), out_bf16=out, stats=st, stats_rows_per_image=h * wd,
                     cta_group=2)
elif case == "sq":
    a, w = rb(8192, 8192), rb(8192, 8192)
    out = torch.empty(8192, 8192, device=DEV, dtype=ops.h16())
    op = ops.gemm(a, w, out_bf16=out, cta_group=1)
elif case == "sqcg2":
    a, w = rb(8192, 8192), rb(8192, 8192)
    out = torch.empty(8192, 8192, device=DEV, dtype=ops.h16())
    op = ops.gemm(a, w, out_bf16=out, cta_group=2)
elif case == "head":
    a = rb(8 * 482 * 642, 128); w = rb(3, 9 * 128); out = torch.empty(8 * 480 * 640, 3, device=DEV)
    op = ops.conv3x3(a, w, 8, 480, 640, bias=torch.zeros(3, device=DEV), out_f32=out)
elif case in ("up2", "up1"):  # one output parity of the VAE decoder's "nearest 2x + 3x3 conv" (2x2 taps), 16 images
    b, h, wd, c = (16, 240, 320, 256) if case == "up2" else (16, 120, 160, 512)
    low = rb(b * (h + 2) * (wd + 2), c)
    wmats = [m.to(DEV).to(ops.h16()) for m in ops.up2x_weight_matrices(torch.randn(c, c, 3, 3) * (9 * c) ** -0.5)]
    out = torch.zeros(b * (2 * h + 2) * (2 * wd + 2), c, device=DEV, dtype=ops.h16())
    st = ops.new_stats(b, c, DEV, replicas=4)
    op = ops.conv_up2x(low, wmats, b, h, wd, bias=torch.zeros(c, device=DEV), pad_out=True, out_bf16=out, stats=st,
                       stats_rows_per_image=(2 * h + 2) * (2 * wd + 2))[3]
elif case == "attn":
    batch, ntok, heads = 16, 4800, 5
    c = heads * 64
    qkv = rb(batch * ntok, 3 * c)
    out = torch.empty(batch * ntok, c, device=DEV, dtype=ops.h16())
    op = ops.flash_attn(qkv, batch, ntok, heads, out, 0, c, 2 * c)
elif case in ("xattnf", "xattnf10"):     # collapsed cross-attention + LN3, UNet level 0 / 1, 112 images
    heads = 5 if case == "xattnf" else 10
    rpg, groups, ntp, c = (16 * 4800, 7, 4, 320) if heads == 5 else (16 * 1200, 7, 4, 640)
    hs = torch.randn(groups * rpg, c, device=DEV)
    a0, bm = torch.randn(7, heads, ntp, c) * 0.05, torch.randn(7, heads, ntp, c) * 0.3
    ap, ca, bmt = [t.to(DEV) for t in ops.xattn_tables(a0, torch.ones(c), torch.zeros(c), bm, [3, 3, 3, 4, 4, 3, 3], ntp)]
    vec = lambda: torch.randn(c, device=DEV)
    out = torch.empty(groups * rpg, c, device=DEV, dtype=ops.h16())
    op = ops.xattn_fused(hs, ap, ca, bmt, vec(), vec(), vec(), list(range(7)), rpg, heads, ntp, out)
elif case == "vattn":        # VAE mid-block attention: one head of 512 channels, 16 images of 4800 tokens
    batch, ntok, c = 16, 4800, 512
    qkv = rb(batch * ntok, 3 * c)
    out = torch.empty(batch * ntok, c, device=DEV, dtype=ops.h16())
    op = ops.flash_attn(qkv, batch, ntok, 1, out, 0, c, 2 * c, scale=c ** -0.5, head_dim=c)
elif case == "gnapply":      # GroupNorm + SiLU, VAE decoder half resolution: 16 images, 256 channels, padded in and out
    b, h, wd, c = 16, 240, 320, 256
    x = rb(b * (h + 2) * (wd + 2), c)
    st = ops.new_stats(b, c, DEV, replicas=4)
    st[0, :, :, 2] = int(h * wd * 0.25 * 2 ** 32)
    out = torch.empty_like(x)
    op = ops.gn_apply(x, st, b, h, wd, torch.ones(c, device=DEV), torch.zeros(c, device=DEV), out, eps=1e-6, silu=True,
                      pad_out=True, x_padded=True)
else:
    raise SystemExit("unknown case")
for _ in range(3):
    op.run()
torch.cuda.synchronize()
print("ok", case)
