"""Per-kernel numerics on the B200: each CUDA kernel against a plain PyTorch fp32 restatement of the same op
(the op-level oracle), called through the C ABI.  bf16 operands are rounded first so the comparison isolates
the kernel's arithmetic; tolerances are stated per test."""
import math
import os
import sys

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ops():
    from stablemtl_b200 import ops, _lib
    return ops, _lib


def H16():
    from stablemtl_b200 import ops
    return ops.h16()


@pytest.fixture(autouse=True, params=["fp16", "bf16"])
def precision(request):
    """every kernel test runs in both 16-bit operand formats"""
    from stablemtl_b200 import ops
    ops.set_precision(request.param)
    yield request.param
    ops.set_precision("fp16")


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


@pytest.mark.parametrize("m,n,k,bn", [
    (128, 256, 64, 0), (256, 256, 320, 0), (300, 320, 320, 0), (1000, 640, 1280, 0), (4800, 1280, 320, 0),
    (130, 4, 576, 0), (777, 96, 200, 0), (512, 960, 320, 0), (512, 1920, 640, 0), (20000, 320, 2880, 0),
    (512, 320, 128, 64), (512, 320, 128, 128), (512, 320, 128, 224),
])
def test_gemm_plain(m, n, k, bn):
    ops, L = _ops()
    kp = (k + 7) // 8 * 8
    a = rnd(m, kp, seed=1).to(H16())[:, :k] if kp != k else rnd(m, k, seed=1).to(H16())
    b = rnd(n, kp, scale=k ** -0.5, seed=2).to(H16())
    if kp != k:
        b = b[:, :k]
    bias = rnd(n, seed=3)
    res = rnd(m, n, seed=4)
    out = torch.full((m, n), float("nan"), device=DEV)
    outb = torch.empty(m, n, device=DEV, dtype=H16()) if n % 8 == 0 else None
    ops.gemm(a, b, bias=bias, res1=res, out_f32=out, out_bf16=outb, block_n=bn).run()
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t() + bias + res
    err = rel_l2(out, ref)
    assert err < 2e-5, f"gemm {m}x{n}x{k}: rel-L2 {err}"   # fp32 accumulate of exact bf16 products
    if outb is not None:
        assert rel_l2(outb.float(), ref) < 4e-3


def test_gemm_gelu_aux_two_res():
    ops, L = _ops()
    m, n, k = 700, 640, 640
    a, b = rnd(m, k, seed=1).to(H16()), rnd(n, k, scale=k ** -0.5, seed=2).to(H16())
    bias, r1, r2 = rnd(n, seed=3), rnd(m, n, seed=4), rnd(m, n, seed=5)
    out = torch.empty(m, n, device=DEV)
    aux = torch.empty(m, n, device=DEV, dtype=H16())
    ops.gemm(a, b, bias=bias, act=L.ACT_GELU, res1=r1, res2=r2, out_f32=out, aux_bf16=aux).run()
    torch.cuda.synchronize()
    pre = F.gelu(a.float() @ b.float().t() + bias)
    assert rel_l2(out, pre + r1 + r2) < 2e-5
    assert rel_l2(aux.float(), pre) < 4e-3


def test_gemm_geglu():
    ops, L = _ops()
    from stablemtl_b200.weights import interleave_geglu
    m, c = 600, 320
    a = rnd(m, c, seed=1).to(H16())
    w = rnd(8 * c, c, scale=c ** -0.5, seed=2).to(H16())
    bias = rnd(8 * c, seed=3)
    wi, bi = interleave_geglu(w, bias)
    out = torch.empty(m, 4 * c, device=DEV, dtype=H16())
    ops.gemm(a, wi, bias=bi, act=L.ACT_GEGLU, out_bf16=out).run()
    torch.cuda.synchronize()
    h = a.float() @ w.float().t() + bias
    ref = h[:, :4 * c] * F.gelu(h[:, 4 * c:])
    assert rel_l2(out.float(), ref) < 4e-3


def test_gemm_bias_per_row_swapped():
    ops, L = _ops()
    m, n, k = 512, 300, 512      # D^T = W X^T : rows are output channels
    w, x = rnd(m, k, scale=k ** -0.5, seed=1).to(H16()), rnd(n, k, seed=2).to(H16())
    bias = rnd(m, seed=3)
    out = torch.empty(m, 304, device=DEV, dtype=H16())
    ops.gemm(w, x, bias=bias, bias_per_row=True, out_bf16=out[:, :300]).run()
    torch.cuda.synchronize()
    ref = w.float() @ x.float().t() + bias[:, None]
    assert rel_l2(out[:, :300].float(), ref) < 4e-3


def _pad_layout(x_nhwc):
    b, h, w, c = x_nhwc.shape
    p = torch.zeros(b, h + 2, w + 2, c, device=x_nhwc.device, dtype=x_nhwc.dtype)
    p[:, 1:-1, 1:-1] = x_nhwc
    return p.reshape(-1, c)


@pytest.mark.parametrize("b,h,w,cin,cout,cs", [(2, 8, 10, 64, 64, 0), (3, 15, 20, 320, 640, 0), (2, 30, 40, 640, 320, 960),
                                               (1, 60, 80, 128, 4, 0), (2, 6, 20, 1280, 1280, 0)])
def test_conv3x3_implicit_gemm(b, h, w, cin, cout, cs):
    ops, L = _ops()
    x = rnd(b, h, w, cin, seed=1).to(H16())
    wt = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=2).to(H16())
    bias = rnd(cout, seed=3)
    res = rnd(b * h * w, cout, seed=4)
    wmat = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1)
    xs = None
    if cs:
        xs_nhwc = rnd(b, h, w, cs, seed=5).to(H16())
        ws = rnd(cout, cs, scale=cs ** -0.5, seed=6).to(H16())
        wmat = torch.cat([wmat, ws], dim=1)
        ref = ref + F.conv2d(xs_nhwc.float().permute(0, 3, 1, 2), ws.float()[:, :, None, None])
        xs = _pad_layout(xs_nhwc)
    wmat = wmat.contiguous()
    ref = ref.permute(0, 2, 3, 1).reshape(b * h * w, cout) + res
    out = torch.full((b * h * w, cout), float("nan"), device=DEV)
    ops.conv3x3(_pad_layout(x), wmat, b, h, w, a_short=xs, bias=bias, res1=res, out_f32=out).run()
    torch.cuda.synchronize()
    err = rel_l2(out, ref)
    assert err < 2e-5, f"conv rel-L2 {err}"


@pytest.mark.parametrize("batch,ntok,heads", [(1, 128, 1), (2, 300, 5), (1, 80, 20), (2, 1200, 10), (1, 4800, 5), (3, 468, 2),
                                              (1, 24, 2), (2, 64, 1), (1, 65, 1), (2, 129, 2), (1, 512, 1), (2, 513, 1), (2, 700, 3)])
def test_flash_attention(batch, ntok, heads):
    ops, L = _ops()
    c = heads * 64
    qkv = rnd(batch * ntok, 3 * c, seed=1).to(H16())
    out = torch.zeros(batch * ntok, c, device=DEV, dtype=H16())
    ops.flash_attn(qkv, batch, ntok, heads, out, 0, c, 2 * c).run()
    torch.cuda.synchronize()
    q, k, v = [t.float().reshape(batch, ntok, heads, 64).permute(0, 2, 1, 3) for t in qkv.split(c, dim=1)]
    ref = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1) @ v
    ref = ref.permute(0, 2, 1, 3).reshape(batch * ntok, c)
    err = rel_l2(out.float(), ref)
    assert err < 1e-2, f"flash attention rel-L2 {err}"   # P and O are rounded to bf16


@pytest.mark.parametrize("rows,c,bf16_in", [(1000, 320, False), (77, 640, False), (300, 1280, True)])
def test_layer_norm(rows, c, bf16_in):
    ops, L = _ops()
    x = rnd(rows, c, seed=1) * 3 + 1
    if bf16_in:
        x = x.to(H16())
    ng = 2
    rpg = (rows + 1) // 2
    g0, b0, g1, b1 = rnd(ng, c, seed=2) + 1, rnd(ng, c, seed=3), rnd(ng, c, seed=4) + 1, rnd(ng, c, seed=5)
    o0 = torch.empty(rows, c, device=DEV, dtype=H16())
    o1 = torch.empty(rows, c, device=DEV, dtype=H16())
    ops.layer_norm(x, g0, b0, o0, gamma1=g1, beta1=b1, out1=o1, rows_per_group=rpg).run()
    torch.cuda.synchronize()
    xn = F.layer_norm(x.float(), (c,), eps=1e-5)
    grp = (torch.arange(rows, device=DEV) // rpg)
    assert rel_l2(o0.float(), xn * g0[grp] + b0[grp]) < 4e-3
    assert rel_l2(o1.float(), xn * g1[grp] + b1[grp]) < 4e-3


@pytest.mark.parametrize("h,w,oh,ow", [(8, 10, 15, 20), (15, 20, 30, 40), (6, 20, 12, 39), (4, 4, 8, 8)])
def test_upsample_pad(h, w, oh, ow):
    ops, L = _ops()
    b, c = 2, 64
    x = rnd(b, h, w, c, seed=1)
    out = torch.full((b * (oh + 2) * (ow + 2), c), float("nan"), device=DEV, dtype=H16())
    ops.upsample_pad(x, b, h, w, oh, ow, out).run()
    torch.cuda.synchronize()
    ref = F.interpolate(x.permute(0, 3, 1, 2), size=(oh, ow), mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(out.float(), _pad_layout(ref.to(H16())).float())


@pytest.mark.parametrize("c,stride,pt,pl,kpad", [(64, 2, 1, 1, 576), (128, 2, 0, 0, 1152), (3, 1, 1, 1, 64), (12, 1, 1, 1, 128)])
def test_im2col(c, stride, pt, pl, kpad):
    ops, L = _ops()
    b, h, w = 2, 12, 14
    x = rnd(b, h, w, c, seed=1)
    if stride == 2 and pt == 0:
        oh, ow = (h + 1 - 3) // 2 + 1, (w + 1 - 3) // 2 + 1     # diffusers Downsample2D: pad (0,1,0,1), stride 2
    else:
        oh, ow = (h + 2 * pt - 3) // stride + 1, (w + 2 * pl - 3) // stride + 1
    out = torch.full((b * oh * ow, kpad), float("nan"), device=DEV, dtype=H16())
    ops.im2col(x, b, h, w, out, stride=stride, pad_t=pt, pad_l=pl, oh=oh, ow=ow).run()
    torch.cuda.synchronize()
    xn = x.permute(0, 3, 1, 2)
    if stride == 2 and pt == 0:
        xn = F.pad(xn, (0, 1, 0, 1))
        cols = F.unfold(xn, 3, stride=2)
    else:
        cols = F.unfold(xn, 3, padding=pt, stride=stride)
    cols = cols.reshape(b, c, 9, oh * ow).permute(0, 3, 2, 1).reshape(b * oh * ow, 9 * c)   # [pix, tap*c + ch]
    assert torch.equal(out[:, :9 * c].float(), cols.to(H16()).float())
    assert out[:, 9 * c:].abs().sum() == 0


def test_xattn_small_keys():
    ops, L = _ops()
    heads, rows_per_group, groups = 5, 100, 3
    c = heads * 64
    q = rnd(groups * rows_per_group, c, seed=1).to(H16())
    kc, vc = rnd(7, 4, c, seed=2), rnd(7, 4, c, seed=3)
    ntok = [3, 3, 3, 4, 4, 3, 3]
    tog = [3, 0, 6]
    out = torch.empty_like(q)
    ops.xattn(q, kc, vc, ntok, tog, rows_per_group, heads, out).run()
    torch.cuda.synchronize()
    ref = torch.empty(groups * rows_per_group, c, device=DEV)
    for g, t in enumerate(tog):
        qq = q[g * rows_per_group:(g + 1) * rows_per_group].float().reshape(-1, heads, 64)
        k = kc[t, :ntok[t]].reshape(ntok[t], heads, 64)
        v = vc[t, :ntok[t]].reshape(ntok[t], heads, 64)
        s = torch.einsum("nhd,jhd->nhj", qq, k) / 8.0
        ref[g * rows_per_group:(g + 1) * rows_per_group] = torch.einsum("nhj,jhd->nhd", s.softmax(-1), v).reshape(-1, c)
    assert rel_l2(out.float(), ref) < 4e-3


@pytest.mark.parametrize("c", [320, 640, 1280])
def test_task_attention(c):
    ops, L = _ops()
    rpg, nheads = 150, 4
    main, src = [2, 0, 5], [0, 1, 2, 3, 4, 5, 6]
    q = rnd(len(main) * rpg, c, seed=1).to(H16())
    k = rnd(len(src) * rpg, c, seed=2).to(H16())
    v = rnd(len(src) * rpg, c, seed=3).to(H16())
    out = torch.empty_like(q)
    ops.task_attn(q, k, v, out, c, nheads, main, src, rpg).run()
    torch.cuda.synchronize()
    dh = c // nheads
    ref = torch.empty(len(main) * rpg, c, device=DEV)
    for g, t in enumerate(main):
        sel = [i for i, s in enumerate(src) if s != t]
        qq = q[g * rpg:(g + 1) * rpg].float().reshape(rpg, nheads, dh)
        kk = torch.stack([k[i * rpg:(i + 1) * rpg].float() for i in sel], 1).reshape(rpg, len(sel), nheads, dh)
        vv = torch.stack([v[i * rpg:(i + 1) * rpg].float() for i in sel], 1).reshape(rpg, len(sel), nheads, dh)
        s = torch.einsum("nhd,nthd->nht", qq, kk) / math.sqrt(dh)
        ref[g * rpg:(g + 1) * rpg] = torch.einsum("nht,nthd->nhd", s.softmax(-1), vv).reshape(rpg, c)
    assert rel_l2(out.float(), ref) < 4e-3


@pytest.mark.parametrize("n", [4800, 8192, 77, 4802])      # 128-bit path (n % 4 == 0) and the scalar one
def test_softmax_rows(n):
    ops, L = _ops()
    s = rnd(300, n, seed=1) * 5
    p = torch.empty(300, n, device=DEV, dtype=H16())
    ops.softmax_rows(s, p, 0.25).run()
    torch.cuda.synchronize()
    assert rel_l2(p.float(), torch.softmax(s * 0.25, -1)) < 4e-3
    # a strided view (the VAE mid-attention's padded score matrix)
    big = rnd(64, n + 8, seed=2) * 3
    pb = torch.zeros(64, n + 8, device=DEV, dtype=H16())
    ops.softmax_rows(big[:, :n], pb[:, :n], 0.5).run()
    torch.cuda.synchronize()
    assert rel_l2(pb[:, :n].float(), torch.softmax(big[:, :n] * 0.5, -1)) < 4e-3 and float(pb[:, n:].abs().max()) == 0.0


def test_task_map_modes():
    ops, L = _ops()
    b, hw = 2, 500
    x = rnd(b, hw, 3, seed=1) * 1.2
    pal = (torch.tensor([[128, 64, 128], [70, 70, 70], [153, 153, 153], [250, 170, 30], [220, 220, 0], [107, 142, 35],
                         [70, 130, 180], [0, 0, 142]], dtype=torch.float32, device=DEV) / 255.0 * 2.0 - 1.0)
    xc = x.clamp(-1, 1).permute(0, 2, 1)          # [b,3,hw]
    clip = torch.empty(b, 1, hw, device=DEV); post = torch.empty(b, 1, hw, device=DEV)
    ops.task_map(x, b, hw, L.MAP_MEAN1, out_clipped=clip, out_post=post).run()
    m = x.mean(-1).clamp(-1, 1)[:, None]
    assert torch.allclose(clip, m, atol=1e-6) and torch.allclose(post, (m + 1) / 2, atol=1e-6)
    post3 = torch.empty(b, 3, hw, device=DEV)
    ops.task_map(x, b, hw, L.MAP_NORMAL, out_post=post3).run()
    assert torch.allclose(post3, xc / xc.norm(dim=1, keepdim=True).clamp_min(1e-30), atol=1e-6)
    ops.task_map(x, b, hw, L.MAP_RGB3, out_post=post3).run()
    assert torch.allclose(post3, (xc + 1) / 2, atol=1e-6)
    f2 = torch.empty(b, 2, hw, device=DEV)
    ops.task_map(x, b, hw, L.MAP_FLOW2, out_clipped=f2).run()
    assert torch.equal(f2, xc[:, :2])
    ids = torch.empty(b, hw, device=DEV, dtype=torch.int64)
    ops.task_map(x, b, hw, L.MAP_SEMANTIC, out_ids=ids, palette=pal).run()
    torch.cuda.synchronize()
    ref = torch.cdist(xc.permute(0, 2, 1).reshape(-1, 3), pal).argmin(1).reshape(b, hw)
    assert (ids == ref).float().mean() > 0.999


# ------------------------------------------------------------------ 16-bit activations + producer-side GN statistics
@pytest.mark.parametrize("m,n,k,rpi", [(4800, 320, 320, 1200), (1000, 640, 1280, 250), (600, 128, 576, 600),
                                       (960, 1280, 320, 80), (777, 96, 200, 777)])
def test_gemm_16bit_residual_and_channel_stats(m, n, k, rpi):
    """out16 = A B^T + bias + res16 ; stats[img, col] = (sum, sum of squares) of the fp32 value over the image's rows."""
    ops, L = _ops()
    kp = (k + 7) // 8 * 8
    a = rnd(m, kp, seed=1).to(H16())[:, :k]
    b = rnd(n, kp, scale=k ** -0.5, seed=2).to(H16())[:, :k]
    bias = rnd(n, seed=3)
    res = rnd(m, n, seed=4).to(H16())
    images = (m + rpi - 1) // rpi
    stats = ops.new_stats(images, n, DEV)
    out = torch.empty(m, n, device=DEV, dtype=H16())
    ops.gemm(a, b, bias=bias, res1=res, out_bf16=out, stats=stats, stats_rows_per_image=rpi).run()
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t() + bias + res.float()
    assert rel_l2(out.float(), ref) < 4e-3
    st = ops.stats_values(stats)                                 # replicas summed, fixed point -> float64
    for img in range(images):
        blk = ref[img * rpi:(img + 1) * rpi].double()
        assert rel_l2(st[img, :, 0], blk.sum(0)) < 1e-4, img
        assert rel_l2(st[img, :, 1], (blk * blk).sum(0)) < 1e-4, img


@pytest.mark.parametrize("b,h,w,cin,cout", [(3, 8, 10, 64, 128), (2, 15, 20, 320, 320), (1, 60, 80, 128, 128)])
def test_conv3x3_16bit_out_with_stats(b, h, w, cin, cout):
    """the conv row map drops halo rows: they must not leak into the statistics; images straddle tiles (8x10 maps)."""
    ops, L = _ops()
    x = rnd(b, h, w, cin, seed=1).to(H16())
    wt = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=2).to(H16())
    bias = rnd(cout, seed=3)
    res = rnd(b * h * w, cout, seed=4).to(H16())
    wmat = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(b * h * w, cout)
    ref = ref + res.float()
    out = torch.empty(b * h * w, cout, device=DEV, dtype=H16())
    stats = ops.new_stats(b, cout, DEV)
    ops.conv3x3(_pad_layout(x), wmat, b, h, w, bias=bias, res1=res, out_bf16=out, stats=stats,
                stats_rows_per_image=h * w).run()
    torch.cuda.synchronize()
    assert rel_l2(out.float(), ref) < 4e-3
    st = ops.stats_values(stats)
    blk = ref.reshape(b, h * w, cout).double()
    assert rel_l2(st[:, :, 0], blk.sum(1)) < 1e-4
    assert rel_l2(st[:, :, 1], (blk * blk).sum(1)) < 1e-4


@pytest.mark.parametrize("b,h,w,c0,c1,silu,pad,x16", [
    (2, 8, 10, 320, 0, True, True, True), (1, 15, 20, 1280, 640, True, True, True), (2, 30, 40, 640, 320, True, True, True),
    (2, 6, 20, 320, 0, False, False, True), (1, 64, 96, 128, 0, True, True, True), (2, 8, 10, 320, 0, True, True, False)])
def test_gn_apply_from_channel_stats(b, h, w, c0, c1, silu, pad, x16):
    """GroupNorm(+SiLU) over a virtual concat from per-(image, channel) sums; group 21 of 1920 = 1280 + 640 straddles
    the seam."""
    ops, L = _ops()
    dt = H16() if x16 else torch.float32
    x0 = (rnd(b, h * w, c0, seed=1) * 2 + 0.5).to(dt)
    x1 = (rnd(b, h * w, c1, seed=2) - 0.3).to(dt) if c1 else None

    def stats_of(x):
        st = ops.new_stats(b, x.shape[-1], DEV)
        xd = x.double()
        full = torch.stack([xd.sum(1), (xd * xd).sum(1)], dim=-1)
        if st.shape[0] > 1:                                    # spread over two replicas: the consumer must sum them
            st[0], st[1] = ops.stats_encode(full * 0.25), ops.stats_encode(full * 0.75)
        else:
            st[0] = ops.stats_encode(full)
        return st
    s0 = stats_of(x0)
    s1 = stats_of(x1) if c1 else None
    C = c0 + c1
    gamma, beta = rnd(C, seed=3) + 1, rnd(C, seed=4)
    hp, wp = (h + 2, w + 2) if pad else (h, w)
    out = torch.full((b * hp * wp, C), float("nan"), device=DEV, dtype=H16())
    raw = torch.full((b * hp * wp, C), float("nan"), device=DEV, dtype=H16())
    ops.gn_apply(x0.reshape(b * h * w, c0), s0, b, h, w, gamma, beta, out,
                 x1=None if x1 is None else x1.reshape(b * h * w, c1), stats1=s1, eps=1e-5, silu=silu, pad_out=pad,
                 raw=raw).run()
    torch.cuda.synchronize()
    xc = (torch.cat([x0, x1], dim=-1) if c1 else x0).float()
    ref = F.group_norm(xc.reshape(b, h, w, C).permute(0, 3, 1, 2), 32, gamma, beta, eps=1e-5)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 3, 1)
    rawref = xc.reshape(b, h, w, C)
    if pad:
        ref, rawref = _pad_layout(ref), _pad_layout(rawref)
    assert rel_l2(out.float(), ref.reshape(-1, C)) < 4e-3
    assert rel_l2(raw.float(), rawref.reshape(-1, C)) < 4e-3
    if pad:
        halo = out.reshape(b, hp, wp, C)
        assert halo[:, 0].abs().max() == 0 and halo[:, :, 0].abs().max() == 0 and halo[:, -1].abs().max() == 0


def test_upsample_and_im2col_take_16bit_input():
    ops, L = _ops()
    b, h, w, c = 2, 8, 10, 64
    x = rnd(b, h, w, c, seed=1).to(H16())
    out = torch.full((b * 17 * 22, c), float("nan"), device=DEV, dtype=H16())
    ops.upsample_pad(x, b, h, w, 15, 20, out).run()
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), size=(15, 20), mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(out.float(), _pad_layout(ref))
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    col = torch.full((b * oh * ow, 9 * c), float("nan"), device=DEV, dtype=H16())
    ops.im2col(x, b, h, w, col, stride=2, pad_t=1, pad_l=1, oh=oh, ow=ow).run()
    torch.cuda.synchronize()
    cols = F.unfold(x.float().permute(0, 3, 1, 2), 3, padding=1, stride=2)
    cols = cols.reshape(b, c, 9, oh * ow).permute(0, 3, 2, 1).reshape(b * oh * ow, 9 * c)
    assert torch.equal(col.float(), cols)


# ------------------------------------------------------------------ CTA-pair (tcgen05 cta_group::2) variant
@pytest.mark.parametrize("m,n,k,bn", [(256, 256, 64, 0), (1000, 640, 1280, 0), (4800, 1280, 320, 0), (130, 4, 576, 0),
                                      (777, 96, 200, 0), (20000, 320, 2880, 0), (4096, 512, 512, 128), (600, 128, 576, 0),
                                      (2000, 448, 256, 224), (5000, 384, 384, 192)])
def test_gemm_cta_pair(m, n, k, bn):
    ops, L = _ops()
    kp = (k + 7) // 8 * 8
    a = rnd(m, kp, seed=1).to(H16())[:, :k]
    b = rnd(n, kp, scale=k ** -0.5, seed=2).to(H16())[:, :k]
    bias = rnd(n, seed=3)
    res = rnd(m, n, seed=4)
    out = torch.full((m, n), float("nan"), device=DEV)
    ops.gemm(a, b, bias=bias, res1=res, out_f32=out, block_n=bn, cta_group=2).run()
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t() + bias + res
    err = rel_l2(out, ref)
    assert err < 2e-5, f"gemm(cta pair) {m}x{n}x{k}: rel-L2 {err}"


def test_gemm_cta_pair_geglu_and_stats():
    ops, L = _ops()
    from stablemtl_b200.weights import interleave_geglu
    m, c = 1300, 320
    a = rnd(m, c, seed=1).to(H16())
    w = rnd(8 * c, c, scale=c ** -0.5, seed=2).to(H16())
    bias = rnd(8 * c, seed=3)
    wi, bi = interleave_geglu(w, bias)
    out = torch.empty(m, 4 * c, device=DEV, dtype=H16())
    ops.gemm(a, wi, bias=bi, act=L.ACT_GEGLU, out_bf16=out, cta_group=2).run()
    h = a.float() @ w.float().t() + bias
    assert rel_l2(out.float(), h[:, :4 * c] * F.gelu(h[:, 4 * c:])) < 4e-3
    # statistics + 16-bit residual through the pair kernel
    n, k, rpi = 640, 320, 325
    b2 = rnd(n, k, scale=k ** -0.5, seed=5).to(H16())
    res = rnd(m, n, seed=6).to(H16())
    stats = ops.new_stats(4, n, DEV)
    o2 = torch.empty(m, n, device=DEV, dtype=H16())
    ops.gemm(a, b2, res1=res, out_bf16=o2, stats=stats, stats_rows_per_image=rpi, cta_group=2).run()
    torch.cuda.synchronize()
    ref = a.float() @ b2.float().t() + res.float()
    assert rel_l2(o2.float(), ref) < 4e-3
    st = ops.stats_values(stats)
    for img in range(4):
        blk = ref[img * rpi:(img + 1) * rpi].double()
        assert rel_l2(st[img, :, 0], blk.sum(0)) < 1e-4 and rel_l2(st[img, :, 1], (blk * blk).sum(0)) < 1e-4


@pytest.mark.parametrize("b,h,w,cin,cout,cs", [(2, 8, 10, 64, 64, 0), (3, 15, 20, 320, 640, 0), (2, 30, 40, 640, 320, 960),
                                               (1, 60, 80, 128, 128, 0)])
def test_conv3x3_cta_pair(b, h, w, cin, cout, cs):
    ops, L = _ops()
    x = rnd(b, h, w, cin, seed=1).to(H16())
    wt = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=2).to(H16())
    bias = rnd(cout, seed=3)
    res = rnd(b * h * w, cout, seed=4)
    wmat = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1)
    xs = None
    if cs:
        xs_nhwc = rnd(b, h, w, cs, seed=5).to(H16())
        ws = rnd(cout, cs, scale=cs ** -0.5, seed=6).to(H16())
        wmat = torch.cat([wmat, ws], dim=1)
        ref = ref + F.conv2d(xs_nhwc.float().permute(0, 3, 1, 2), ws.float()[:, :, None, None])
        xs = _pad_layout(xs_nhwc)
    ref = ref.permute(0, 2, 3, 1).reshape(b * h * w, cout) + res
    out = torch.full((b * h * w, cout), float("nan"), device=DEV)
    ops.conv3x3(_pad_layout(x), wmat.contiguous(), b, h, w, a_short=xs, bias=bias, res1=res, out_f32=out, cta_group=2).run()
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 2e-5


# ------------------------------------------------------------------ swapped form (weights on the MMA M side, N <= 128)
@pytest.mark.parametrize("b,h,w,cin,cout,cs,with_res", [(1, 200, 200, 128, 128, 0, True), (2, 150, 140, 128, 128, 0, False),
                                                        (1, 200, 200, 256, 128, 256, False), (1, 210, 190, 64, 96, 0, True)])
def test_conv3x3_swapped_narrow_output(b, h, w, cin, cout, cs, with_res):
    """N <= 128 and enough pixels -> the plan picks smtl_gemmT_kernel (channel-major accumulator, transposing
    epilogue); 16-bit output, optional 16-bit residual, fused 1x1 shortcut segment, per-channel statistics."""
    ops, L = _ops()
    x = rnd(b, h, w, cin, seed=1).to(H16())
    wt = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=2).to(H16())
    bias = rnd(cout, seed=3)
    wmat = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1)
    xs = None
    if cs:
        xs_nhwc = rnd(b, h, w, cs, seed=5).to(H16())
        ws = rnd(cout, cs, scale=cs ** -0.5, seed=6).to(H16())
        wmat = torch.cat([wmat, ws], dim=1)
        ref = ref + F.conv2d(xs_nhwc.float().permute(0, 3, 1, 2), ws.float()[:, :, None, None])
        xs = _pad_layout(xs_nhwc)
    ref = ref.permute(0, 2, 3, 1).reshape(b * h * w, cout)
    res = rnd(b * h * w, cout, seed=4).to(H16()) if with_res else None
    if with_res:
        ref = ref + res.float()
    out = torch.full((b * h * w, cout), float("nan"), device=DEV, dtype=H16())
    stats = ops.new_stats(b, cout, DEV)
    op = ops.conv3x3(_pad_layout(x), wmat.contiguous(), b, h, w, a_short=xs, bias=bias, res1=res, out_bf16=out,
                     stats=stats, stats_rows_per_image=h * w)
    assert op.struct.cta_group == 3, "expected the swapped kernel"
    op.run()
    torch.cuda.synchronize()
    assert rel_l2(out.float(), ref) < 4e-3
    st = ops.stats_values(stats)
    blk = ref.reshape(b, h * w, cout).double()
    assert rel_l2(st[:, :, 0], blk.sum(1)) < 1e-4
    assert rel_l2(st[:, :, 1], (blk * blk).sum(1)) < 1e-4


def test_gemm_swapped_plain():
    ops, L = _ops()
    m, n, k = 50000, 128, 320
    a, b = rnd(m, k, seed=1).to(H16()), rnd(n, k, scale=k ** -0.5, seed=2).to(H16())
    bias = rnd(n, seed=3)
    out = torch.empty(m, n, device=DEV, dtype=H16())
    op = ops.gemm(a, b, bias=bias, out_bf16=out)
    assert op.struct.cta_group == 3
    op.run()
    torch.cuda.synchronize()
    assert rel_l2(out.float(), a.float() @ b.float().t() + bias) < 4e-3


@pytest.mark.parametrize("b,h,w,cin,cout", [(2, 8, 10, 64, 64), (1, 15, 20, 320, 320), (2, 30, 40, 256, 512)])
def test_conv_up2x_parity_decomposition(b, h, w, cin, cout):
    """nearest-2x upsample + 3x3 conv == four 2x2 convs on the low-res map (ops.conv_up2x), incl. statistics"""
    ops, L = _ops()
    x = rnd(b, h, w, cin, seed=1).to(H16())
    wt = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=2)
    bias = rnd(cout, seed=3)
    wmats = [m.to(H16()).contiguous() for m in ops.up2x_weight_matrices(wt)]
    # reference with the SAME (summed, then rounded) weights the kernel sees would hide a decomposition bug, so the
    # reference uses the original filter on the upsampled map; tolerance covers the 16-bit rounding of summed taps
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, wt, bias, padding=1).permute(0, 2, 3, 1).reshape(b * 4 * h * w, cout)
    out = torch.full((b * 4 * h * w, cout), float("nan"), device=DEV, dtype=H16())
    stats = ops.new_stats(b, cout, DEV)
    for op in ops.conv_up2x(_pad_layout(x), wmats, b, h, w, bias=bias, out_bf16=out, stats=stats,
                            stats_rows_per_image=4 * h * w):
        op.run()
    torch.cuda.synchronize()
    assert not torch.isnan(out.float()).any()
    assert rel_l2(out.float(), ref) < 5e-3
    st = ops.stats_values(stats)
    blk = ref.reshape(b, 4 * h * w, cout).double()
    assert rel_l2(st[:, :, 0], blk.sum(1)) < 2e-3
    assert rel_l2(st[:, :, 1], (blk * blk).sum(1)) < 2e-3


@pytest.mark.parametrize("G,R,n,k", [(7, 640, 320, 640), (7, 2400, 1280, 640), (3, 300, 320, 320), (5, 72, 64, 128)])
def test_gemm_grouped_per_task_weights(G, R, n, k):
    """rows [g*R, (g+1)*R) use weight / bias group g (the per-task MLPs of the task attention in one launch); R need not
    be a multiple of the 128-row tile: M tiles restart at every group"""
    ops, L = _ops()
    a = rnd(G * R, k, seed=1).to(H16())
    w = rnd(G, n, k, scale=k ** -0.5, seed=2).to(H16())
    bias = rnd(G, n, seed=3)
    out = torch.full((G * R, n), float("nan"), device=DEV, dtype=H16())
    ops.gemm(a, w.reshape(G * n, k), n=n, bias=bias.reshape(-1), act=L.ACT_GELU, out_bf16=out, group_rows=R).run()
    torch.cuda.synchronize()
    ref = torch.cat([F.gelu(a[g * R:(g + 1) * R].float() @ w[g].float().t() + bias[g]) for g in range(G)])
    assert rel_l2(out.float(), ref) < 4e-3


# ------------------------------------------------------------------ padded-layout outputs (conv chains that stay padded)
def _unpad(x_pad, b, h, w):
    return x_pad.reshape(b, h + 2, w + 2, -1)[:, 1:-1, 1:-1].reshape(b * h * w, -1)


@pytest.mark.parametrize("b,h,w,cin,cout,cs", [(2, 8, 10, 64, 64, 0), (2, 30, 40, 128, 256, 128), (1, 200, 200, 128, 128, 0),
                                               (1, 200, 200, 256, 128, 256)])
def test_conv3x3_padded_output_zero_halo(b, h, w, cin, cout, cs):
    """ROWMAP_PAD_KEEP (normal and swapped kernel): output and 16-bit residual in the padded layout, halo rows of the
    output written as zeros even over a dirty buffer, statistics exclude the halo."""
    ops, L = _ops()
    x = rnd(b, h, w, cin, seed=1).to(H16())
    wt = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=2).to(H16())
    bias = rnd(cout, seed=3)
    wmat = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1)
    xs = None
    if cs:
        xs_nhwc = rnd(b, h, w, cs, seed=5).to(H16())
        ws = rnd(cout, cs, scale=cs ** -0.5, seed=6).to(H16())
        wmat = torch.cat([wmat, ws], dim=1)
        ref = ref + F.conv2d(xs_nhwc.float().permute(0, 3, 1, 2), ws.float()[:, :, None, None])
        xs = _pad_layout(xs_nhwc)
        xs[xs == 0] = 7.0                                        # the shortcut source's halo may hold anything
    res_nhwc = rnd(b, h, w, cout, seed=4).to(H16())
    ref = ref.permute(0, 2, 3, 1).reshape(b * h * w, cout) + res_nhwc.float().reshape(b * h * w, cout)
    res = _pad_layout(res_nhwc)
    out = torch.full((b * (h + 2) * (w + 2), cout), 3.0, device=DEV, dtype=H16())      # dirty
    stats = ops.new_stats(b, cout, DEV)
    ops.conv3x3(_pad_layout(x), wmat.contiguous(), b, h, w, a_short=xs, bias=bias, res1=res, out_bf16=out, stats=stats,
                stats_rows_per_image=(h + 2) * (w + 2), pad_out=True).run()
    torch.cuda.synchronize()
    assert rel_l2(_unpad(out, b, h, w).float(), ref) < 4e-3
    o4 = out.reshape(b, h + 2, w + 2, cout).float()
    assert o4[:, 0].abs().max() == 0 and o4[:, -1].abs().max() == 0 and o4[:, :, 0].abs().max() == 0 and o4[:, :, -1].abs().max() == 0
    st = ops.stats_values(stats)
    blk = ref.reshape(b, h * w, cout).double()
    assert rel_l2(st[:, :, 0], blk.sum(1)) < 1e-4 and rel_l2(st[:, :, 1], (blk * blk).sum(1)) < 1e-4


def test_gemm_to_pad_and_up2x_pad_rowmaps():
    ops, L = _ops()
    b, h, w, k, n = 2, 15, 20, 128, 320
    a = rnd(b * h * w, k, seed=1).to(H16())
    wt = rnd(n, k, scale=k ** -0.5, seed=2).to(H16())
    res_nhwc = rnd(b, h, w, n, seed=3).to(H16())
    out = torch.zeros(b * (h + 2) * (w + 2), n, device=DEV, dtype=H16())
    stats = ops.new_stats(b, n, DEV)
    ops.gemm(a, wt, res1=_pad_layout(res_nhwc), out_bf16=out, rowmap=L.ROWMAP_TO_PAD, img_hw=(h, w), stats=stats,
             stats_rows_per_image=(h + 2) * (w + 2)).run()
    torch.cuda.synchronize()
    ref = a.float() @ wt.float().t() + res_nhwc.float().reshape(-1, n)
    assert rel_l2(_unpad(out, b, h, w).float(), ref) < 4e-3
    assert rel_l2(ops.stats_values(stats)[:, :, 0], ref.reshape(b, h * w, n).double().sum(1)) < 1e-4
    # up2x into a padded 2x map
    cin, cout = 64, 128
    x = rnd(b, h, w, cin, seed=5).to(H16())
    w3 = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=6)
    wm = [m.to(H16()).contiguous() for m in ops.up2x_weight_matrices(w3)]
    up_ref = F.conv2d(F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest"), w3, padding=1)
    up_ref = up_ref.permute(0, 2, 3, 1).reshape(b * 4 * h * w, cout)
    o2 = torch.zeros(b * (2 * h + 2) * (2 * w + 2), cout, device=DEV, dtype=H16())
    for op in ops.conv_up2x(_pad_layout(x), wm, b, h, w, out_bf16=o2, pad_out=True):
        op.run()
    torch.cuda.synchronize()
    assert rel_l2(_unpad(o2, b, 2 * h, 2 * w).float(), up_ref) < 5e-3


def test_gn_apply_reads_padded_input():
    ops, L = _ops()
    b, h, w, c = 2, 8, 10, 128
    x = (rnd(b, h, w, c, seed=1) * 2 + 0.5).to(H16())
    xp = _pad_layout(x)
    xp[xp == 0] = 9.0                                            # halo content must be ignored
    st = ops.new_stats(b, c, DEV)
    xd = x.double().reshape(b, h * w, c)
    st[0] = ops.stats_encode(torch.stack([xd.sum(1), (xd * xd).sum(1)], dim=-1))
    gamma, beta = rnd(c, seed=3) + 1, rnd(c, seed=4)
    out = torch.full((b * (h + 2) * (w + 2), c), float("nan"), device=DEV, dtype=H16())
    ops.gn_apply(xp, st, b, h, w, gamma, beta, out, eps=1e-6, silu=True, pad_out=True, x_padded=True).run()
    torch.cuda.synchronize()
    ref = F.silu(F.group_norm(x.float().permute(0, 3, 1, 2), 32, gamma, beta, eps=1e-6)).permute(0, 2, 3, 1)
    assert rel_l2(out.float(), _pad_layout(ref)) < 4e-3


# ------------------------------------------------------------------------------------------------- evaluation pre-reductions
@pytest.mark.parametrize("masked", [True, False])
def test_lsq_sums_and_scale_shift_vs_oracle(masked):
    """device sums -> scale/shift against the numpy restatement of align_depth_least_square (alignment.py:122-169)"""
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import metrics_oracle as MO
    from stablemtl_b200.evaluate import lsq_scale_shift
    ops, L = _ops()
    b, h, w = 3, 480, 640
    g = torch.Generator().manual_seed(7)
    pred = torch.rand(b, h, w, generator=g)
    gt = 3.0 * pred + 0.25 + 0.1 * torch.randn(b, h, w, generator=g)
    valid = (torch.rand(b, h, w, generator=g) > 0.4) if masked else torch.ones(b, h, w, dtype=torch.bool)
    sums = torch.zeros(b, 5, dtype=torch.float64, device=DEV)
    ops.lsq_sums(pred.to(DEV).reshape(b, -1), gt.to(DEV).reshape(b, -1),
                 valid.to(DEV).to(torch.uint8).reshape(b, -1) if masked else None, sums).run()
    torch.cuda.synchronize()
    scale, shift = lsq_scale_shift(sums)
    for i in range(b):
        assert int(sums[i, 0].item()) == int(valid[i].sum())                 # the count is exact
        _, s, t = MO.align_least_square(gt[i].numpy(), pred[i].numpy(), valid[i].numpy())
        assert abs(scale[i] - float(s[0])) <= 1e-4 * abs(float(s[0])) and abs(shift[i] - float(t[0])) <= 1e-4
    # sums are linear: accumulating the two halves of an image separately gives the whole
    two = torch.zeros(1, 5, dtype=torch.float64, device=DEV)
    half = h * w // 2
    p0, g0 = pred[0].reshape(1, -1).to(DEV), gt[0].reshape(1, -1).to(DEV)
    ops.lsq_sums(p0[:, :half].contiguous(), g0[:, :half].contiguous(), None, two).run()
    ops.lsq_sums(p0[:, half:].contiguous(), g0[:, half:].contiguous(), None, two).run()
    whole = torch.zeros(1, 5, dtype=torch.float64, device=DEV)
    ops.lsq_sums(p0, g0, None, whole).run()
    torch.cuda.synchronize()
    assert torch.allclose(two, whole, rtol=1e-12, atol=0)


def test_confusion_histogram_bit_exact_vs_oracle():
    """integer work: bit-exact against SemanticMetrics.update/_fast_hist (metric_semantic.py:34-50)"""
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import metrics_oracle as MO
    ops, L = _ops()
    g = torch.Generator().manual_seed(11)
    lt = torch.randint(-1, 10, (4, 480, 640), generator=g)                  # -1, 8, 9: ignored labels
    lp = torch.randint(0, 8, (4, 480, 640), generator=g)
    vm = torch.rand(4, 480, 640, generator=g) > 0.25
    hist = torch.zeros(65, dtype=torch.int64, device=DEV)
    ops.confusion(lt.to(DEV), lp.to(DEV), vm.to(DEV).to(torch.uint8), hist, 8).run()
    torch.cuda.synchronize()
    want = MO.confusion(lt.numpy(), lp.numpy(), vm.numpy(), 8)
    assert np.array_equal(hist[:64].cpu().numpy().reshape(8, 8), want.astype(np.int64)) and int(hist[64]) == 0
    # accumulation + no mask + an out-of-range prediction is counted aside, not binned
    lp2 = lp.clone()
    lp2[0, 0, :5] = 8
    lt2 = lt.clone()
    lt2[0, 0, :5] = 3
    ops.confusion(lt2.to(DEV), lp2.to(DEV), None, hist, 8).run()
    torch.cuda.synchronize()
    lt3, lp3 = lt2.numpy().reshape(-1), lp2.numpy().reshape(-1)
    keep = lp3 < 8
    want2 = want + MO.fast_hist(lt3[keep], lp3[keep], 8)
    assert np.array_equal(hist[:64].cpu().numpy().reshape(8, 8), want2.astype(np.int64)) and int(hist[64]) == 5
    # empty / tiny inputs
    h1 = torch.zeros(65, dtype=torch.int64, device=DEV)
    ops.confusion(torch.tensor([2], device=DEV), torch.tensor([2], device=DEV), None, h1, 8).run()
    torch.cuda.synchronize()
    assert int(h1.sum()) == 1 and int(h1[2 * 8 + 2]) == 1


def test_gn_finalize_scale_shift_table():
    """producer-side sums -> per-(image, channel) (scale, shift): GroupNorm(x)[b, c] == x * scale + shift"""
    ops, L = _ops()
    b, hw, c, groups = 3, 500, 320, 32
    x = rnd(b, hw, c, seed=1) * 2 + 0.7
    st = ops.new_stats(b, c, DEV)
    xd = x.double()
    full = torch.stack([xd.sum(1), (xd * xd).sum(1)], dim=-1)
    if st.shape[0] > 1:
        st[0], st[1] = ops.stats_encode(full * 0.25), ops.stats_encode(full * 0.75)
    else:
        st[0] = ops.stats_encode(full)
    gamma, beta = rnd(c, seed=3) + 1, rnd(c, seed=4)
    ss = torch.zeros(b, c, 2, device=DEV)
    ops.gn_finalize(st, b, hw, gamma, beta, ss, eps=1e-6, groups=groups).run()
    torch.cuda.synchronize()
    ref = F.group_norm(x.permute(0, 2, 1).reshape(b, c, hw, 1), groups, gamma, beta, eps=1e-6).reshape(b, c, hw).permute(0, 2, 1)
    got = x * ss[:, None, :, 0] + ss[:, None, :, 1]
    assert rel_l2(got, ref) < 1e-5


@pytest.mark.parametrize("kind", ["conv", "conv_pair", "conv_pair_split", "swapped", "token"])
def test_channel_stats_are_bit_reproducible_and_batch_position_independent(kind):
    """The GroupNorm statistics are integer fixed-point atomics over image-aligned tiles: two runs give the same BITS,
    and an image's cells do not depend on where it sits in the batch (the reference evaluation is deterministic)."""
    ops, L = _ops()
    if kind == "token":
        b, rpi, n, k = 5, 300, 320, 640
        a = rnd(b, rpi, k, seed=1).to(H16())
        wt = rnd(n, k, scale=k ** -0.5, seed=2).to(H16())

        def run(order):
            x = a[order].reshape(b * rpi, k).contiguous()
            st = ops.new_stats(b, n, DEV)
            out = torch.empty(b * rpi, n, device=DEV, dtype=H16())
            ops.gemm(x, wt, out_bf16=out, stats=st, stats_rows_per_image=rpi).run()
            torch.cuda.synchronize()
            return st.sum(0), out.reshape(b, rpi, n)
    else:
        b, h, w, cin, cout = {"conv": (5, 15, 20, 320, 320), "conv_pair": (3, 60, 80, 256, 512),
                              "conv_pair_split": (5, 15, 20, 320, 320), "swapped": (5, 120, 160, 128, 128)}[kind]
        force = dict(cta_group=2) if kind == "conv_pair_split" else {}
        x = rnd(b, h, w, cin, seed=1).to(H16())
        wt = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=2).to(H16())
        wmat = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()

        def run(order):
            st = ops.new_stats(b, cout, DEV)
            out = torch.empty(b * h * w, cout, device=DEV, dtype=H16())
            op = ops.conv3x3(_pad_layout(x[order]), wmat, b, h, w, out_bf16=out, stats=st, stats_rows_per_image=h * w, **force)
            if kind == "conv_pair_split" and force:  # 374 padded rows per image: the pair takes two consecutive 128-row tiles
                assert op.struct.cta_group == 2 and op.struct.pair_split == 1 and op.struct.tiles_m == 8
            if kind == "swapped":
                assert op.struct.cta_group == 3
            if kind == "conv_pair":
                assert op.struct.cta_group == 2
            op.run()
            torch.cuda.synchronize()
            return st.sum(0), out.reshape(b, h * w, cout)
    ident = list(range(b))
    perm = ident[1:] + ident[:1]
    s0, o0 = run(ident)
    s1, o1 = run(ident)
    assert torch.equal(s0, s1) and torch.equal(o0, o1)
    sp, op_ = run(perm)
    assert torch.equal(sp, s0[perm]) and torch.equal(op_, o0[perm])
    if kind == "conv_pair_split":                # same maps and statistics as the single-CTA tiling of the same conv
        force.clear()
        s1, o1 = run(ident)
        assert rel_l2(o0.float(), o1.float()) < 1e-6
        assert ((s0 - s1).double().abs() <= 1e-5 * s0.double().abs() + 2.0 ** 20).all()      # cells are 2^-32 fixed point


@pytest.mark.parametrize("g,cin,cout,h,w", [(2, 320, 320, 15, 20), (4, 128, 256, 24, 40), (8, 256, 512, 60, 80), (8, 64, 96, 9, 7)])
def test_channel_stats_with_grouped_cells(g, cin, cout, h, w):
    """stats_group = g: the sums of g adjacent channels land in the first channel's cell, the other cells stay zero, and
    the block totals equal those of the per-channel run (what a GroupNorm whose groups are unions of such blocks reads)."""
    ops, L = _ops()
    b = 3
    x = rnd(b, h, w, cin, seed=1).to(H16())
    wmat = rnd(cout, 9 * cin, scale=(9 * cin) ** -0.5, seed=2).to(H16())
    bias = rnd(cout, seed=3)

    def run(group):
        st = ops.new_stats(b, cout, DEV)
        out = torch.empty(b * h * w, cout, device=DEV, dtype=H16())
        ops.conv3x3(_pad_layout(x), wmat, b, h, w, bias=bias, out_bf16=out, stats=st, stats_rows_per_image=h * w,
                    stats_group=group).run()
        torch.cuda.synchronize()
        v = ops.stats_values(st)                       # [images, channels, 2]
        return (v[..., 0], v[..., 1]), out
    (s1, q1), o1 = run(0)
    (sg, qg), og = run(g)
    assert torch.equal(o1, og)
    lead = torch.arange(cout, device=DEV) % g == 0
    assert (sg[:, ~lead] == 0).all() and (qg[:, ~lead] == 0).all()
    ref_s = o1.double().reshape(b, h * w, cout // g, g).sum((1, 3))
    ref_q = (o1.double() ** 2).reshape(b, h * w, cout // g, g).sum((1, 3))
    # the statistics are those of the fp32 value BEFORE its 16-bit rounding: a block total differs from the stored map's
    # by the accumulated roundings (~5e-4 relative per element), but agrees with the per-channel run to fp32 summation order
    n_el = h * w * g
    eps16 = 2.0 ** -8 if H16() == torch.bfloat16 else 2.0 ** -11
    assert torch.allclose(sg[:, lead], ref_s, rtol=0, atol=4 * eps16 * n_el ** 0.5), (sg[:, lead] - ref_s).abs().max()
    assert torch.allclose(qg[:, lead], ref_q, rtol=4 * eps16, atol=1e-2), (qg[:, lead] - ref_q).abs().max()
    assert torch.allclose(s1.reshape(b, cout // g, g).sum(-1), sg[:, lead], rtol=1e-5, atol=1e-3)
    assert torch.allclose(q1.reshape(b, cout // g, g).sum(-1), qg[:, lead], rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("cout", [256, 128])
def test_channel_stats_of_a_large_energetic_map_do_not_overflow(cout):
    """480x640 full-resolution map with std ~ 8 (what the VAE decoder's last level carries for some task latents): the
    per-channel sums of squares are ~2e7 and a 32-group GroupNorm adds 4-8 of them -- the fixed-point cells and the
    consumer's group reduction must hold that (a first version summed a group's cells in int64 and wrapped)."""
    ops, L = _ops()
    b, h, w, cin = 1, 480, 640, 64
    x = rnd(b, h, w, cin, seed=1).to(H16())
    wt = (rnd(cout, cin, 3, 3, scale=8.0 * (9 * cin) ** -0.5, seed=2)).to(H16())
    wmat = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    out = torch.empty(b * h * w, cout, device=DEV, dtype=H16())
    st = ops.new_stats(b, cout, DEV)
    op = ops.conv3x3(_pad_layout(x), wmat, b, h, w, out_bf16=out, stats=st, stats_rows_per_image=h * w)
    assert (op.struct.cta_group == 3) == (cout == 128)             # 128 channels: the swapped kernel
    op.run()
    gamma, beta = rnd(cout, seed=3) + 1, rnd(cout, seed=4)
    y = torch.empty(b * h * w, cout, device=DEV, dtype=H16())
    ops.gn_apply(out, st, b, h, w, gamma, beta, y, eps=1e-6, silu=False, pad_out=False).run()
    torch.cuda.synchronize()
    xo = out.float().reshape(b, h * w, cout)
    assert 6.0 < float(xo.std()) < 11.0
    vals = ops.stats_values(st)
    assert rel_l2(vals[:, :, 1], (xo.double() ** 2).sum(1)) < 1e-3
    ref = F.group_norm(xo.permute(0, 2, 1).reshape(b, cout, h, w), 32, gamma, beta, eps=1e-6)
    assert rel_l2(y.float().reshape(b, h, w, cout).permute(0, 3, 1, 2), ref) < 4e-3


@pytest.mark.parametrize("src", ["uint8", "float255", "normalized"])
def test_rgb_stem_equals_normalise_then_im2col(src):
    """the fused stem producer (normalise + 3x3 im2col of the 3 input channels, one pass) gives the BITS of the two-pass
    form it replaces: rgb_prep (stablemtl_pipeline.py:263) then the scalar im2col"""
    ops, L = _ops()
    b, h, w = 2, 24, 40
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (b, 3, h, w), generator=g, dtype=torch.uint8).to(DEV)
    if src == "uint8":
        rgb, norm = u8, False
    elif src == "float255":
        rgb, norm = u8.float(), False
    else:
        rgb, norm = u8.float() / 255.0 * 2.0 - 1.0, True
    xin = torch.empty(b * h * w, 3, device=DEV)
    ops.rgb_prep(rgb, xin, normalized=norm).run()
    ref = torch.full((b * h * w, 64), float("nan"), device=DEV, dtype=H16())
    ops.im2col(xin.view(b, h, w, 3), b, h, w, ref, stride=1, pad_t=1, pad_l=1, oh=h, ow=w).run()
    got = torch.full((b * h * w, 64), float("nan"), device=DEV, dtype=H16())
    ops.rgb_stem(rgb, got, normalized=norm).run()
    torch.cuda.synchronize()
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
    want = F.unfold(u8.float() / 255.0 * 2.0 - 1.0, 3, padding=1).view(b, 3, 9, h * w).permute(0, 3, 2, 1).reshape(b * h * w, 27)
    assert rel_l2(got[:, :27].float(), want) < 2e-3 and float(got[:, 27:].abs().max()) == 0.0


@pytest.mark.parametrize("b,h,w,cin,cout", [(2, 16, 24, 128, 3), (1, 60, 80, 320, 4), (3, 9, 7, 64, 1)])
def test_conv_head_with_taps_folded_into_n(b, h, w, cin, cout):
    """narrow-output 3x3 conv = a 3-segment implicit GEMM over the padded map (N = 3 * cout) + the horizontal gather"""
    ops, L = _ops()
    x = rnd(b, h, w, cin, seed=1).to(H16())
    wt = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=2).to(H16())
    bias = rnd(cout, seed=3)
    wfold = ops.head_weight_matrix(wt)
    a = _pad_layout(x).reshape(-1, cin)
    part = torch.full((a.shape[0], wfold.shape[0]), float("nan"), device=DEV)
    out = torch.full((b * h * w, cout), float("nan"), device=DEV)
    for op in ops.conv_head(a, wfold, bias, b, h, w, cout, part, out):
        op.run()
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(b * h * w, cout)
    assert rel_l2(out, ref) < 2e-5


@pytest.mark.parametrize("heads,ntp,rpg,groups", [(5, 4, 100, 3), (10, 4, 77, 2), (2, 8, 33, 4), (1, 4, 8, 1), (5, 8, 300, 2),
                                                  (10, 8, 16, 1)])
def test_xattn_fused_collapsed_cross_attention(heads, ntp, rpg, groups):
    """hs += bo + sum_v softmax_token(LN2-normalised(hs) . ap + ca)[v] bm[v];  out = LayerNorm3(hs)  -- one kernel, two
    skinny GEMMs on mma.sync; tables packed by ops.xattn_tables (padding tokens / vectors masked)"""
    ops, L = _ops()
    c, ntask = heads * 64, 3
    hs = rnd(groups * rpg, c, seed=1) * 2 + 0.5
    a0 = (rnd(ntask, heads, ntp, c, seed=2) * 0.05).cpu()
    bm = (rnd(ntask, heads, ntp, c, seed=3) * 0.3).cpu()
    g2, b2 = (rnd(c, seed=8) * 0.1 + 1).cpu(), (rnd(c, seed=9) * 0.1).cpu()
    ntok = [ntp, ntp - 1, max(1, ntp - 2)]
    ap, ca, bmt = [t.to(DEV) for t in ops.xattn_tables(a0, g2, b2, bm, ntok, ntp)]
    v = heads * ntp
    bo, g3, b3 = rnd(c, seed=5), rnd(c, seed=6) + 1, rnd(c, seed=7)
    tasks = [(g * 2) % ntask for g in range(groups)]
    ref_h, ref_o = [], []
    for g, t in enumerate(tasks):
        h = hs[g * rpg:(g + 1) * rpg].double()
        mean, rstd = h.mean(1, keepdim=True), (h.var(1, unbiased=False, keepdim=True) + 1e-5).rsqrt()
        score = ((h - mean) * rstd) @ ap[t, :v].double().t() + ca[t, :v].double()
        prob = torch.softmax(score.view(-1, heads, ntp), -1).view(-1, v)
        hn = h + bo.double() + prob @ ops.xattn_unpermute(bmt[t])[:, :v].double().t()
        ref_h.append(hn)
        ref_o.append(F.layer_norm(hn, (c,), g3.double(), b3.double(), 1e-5))
    ref_h, ref_o = torch.cat(ref_h), torch.cat(ref_o)
    out = torch.full((groups * rpg, c), float("nan"), device=DEV, dtype=H16())
    got_h = hs.clone()
    ops.xattn_fused(got_h, ap, ca, bmt, bo, g3, b3, tasks, rpg, heads, ntp, out).run()
    torch.cuda.synchronize()
    # the normalised row and the probabilities enter the tensor cores in 16 bits: the ADDED term carries ~1e-3 relative
    assert rel_l2(got_h - hs, ref_h - hs.double()) < 6e-3
    assert rel_l2(got_h, ref_h) < 1e-3
    assert rel_l2(out.float(), ref_o) < 4e-3


@pytest.mark.parametrize("batch,ntok", [(2, 4800), (3, 300), (1, 128), (2, 77), (1, 8192)])
def test_flash_attention_d512_single_head(batch, ntok):
    """the VAE mid-block attention (one head of 512 channels): S and P stay in tensor memory, two passes over the keys"""
    ops, L = _ops()
    c = 512
    qkv = (rnd(batch * ntok, 3 * c, seed=1) * 1.5).to(H16())
    out = torch.full((batch * ntok, c), float("nan"), device=DEV, dtype=H16())
    ops.flash_attn(qkv, batch, ntok, 1, out, 0, c, 2 * c, scale=c ** -0.5, head_dim=c).run()
    torch.cuda.synchronize()
    q, k, v = [t.float().reshape(batch, ntok, c) for t in qkv.split(c, dim=1)]
    ref = (torch.softmax(q @ k.transpose(-1, -2) * c ** -0.5, dim=-1) @ v).reshape(batch * ntok, c)
    err = rel_l2(out.float(), ref)
    assert err < 1e-2, f"d=512 flash attention rel-L2 {err}"
