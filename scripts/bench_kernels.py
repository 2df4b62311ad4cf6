"""Micro-benchmarks of the hot kernels on representative StableMTL shapes (CUDA events, L2 flushed between
iterations).  Prints one line per case: time, TFLOP/s, fraction of the measured bf16 peak."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablemtl_b200 import ops, _lib as L  # noqa: E402

DEV = "cuda"
PEAK = 1630.0
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(op, iters=5):
    for _ in range(2):
        op.run()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        op.run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, op, flops=None):
    ms = timeit(op)
    fl = op.flops if flops is None else flops
    tf = fl / ms / 1e9
    print(f"{name:56s} {ms:9.3f} ms {tf:8.1f} TFLOP/s  {tf / PEAK * 100:5.1f}% of measured burst peak", flush=True)


def rb(*shape):
    return (torch.randn(*shape, device=DEV) * 0.5).to(ops.h16())


def gemm_case(name, m, n, k, **kw):
    a, b = rb(m, k), rb(n, k)
    out = torch.empty(m, n, device=DEV)
    report(f"gemm {name} m={m} n={n} k={k}", ops.gemm(a, b, out_f32=out, **kw))


def conv_case(name, batch, h, w, cin, cout, stats=True, **kw):
    """the in-pipeline form: 16-bit output + fused GroupNorm statistics"""
    a = rb(batch * (h + 2) * (w + 2), cin)
    wm = rb(cout, 9 * cin)
    out = torch.empty(batch * h * w, cout, device=DEV, dtype=ops.h16())
    bias = torch.zeros(cout, device=DEV)
    skw = dict(stats=ops.new_stats(batch, cout, DEV, replicas=max(1, min(8, 64 // batch))), stats_rows_per_image=h * w) if stats else {}
    report(f"conv3x3 {name} b={batch} {h}x{w} {cin}->{cout}",
           ops.conv3x3(a, wm, batch, h, w, bias=bias, out_bf16=out, **skw, **kw))


def attn_case(batch, ntok, heads):
    c = heads * 64
    qkv = rb(batch * ntok, 3 * c)
    out = torch.empty(batch * ntok, c, device=DEV, dtype=ops.h16())
    report(f"flash_attn b={batch} ntok={ntok} heads={heads}", ops.flash_attn(qkv, batch, ntok, heads, out, 0, c, 2 * c))


def gn_case(batch, h, w, c, pad=True):
    x = rb(batch * h * w, c)
    st = ops.new_stats(batch, c, DEV)
    st[0, :, :, 0] = 0.0
    st[0, :, :, 1] = float(h * w)
    g, b = torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
    rows = batch * (h + 2) * (w + 2) if pad else batch * h * w
    out = torch.empty(rows, c, device=DEV, dtype=ops.h16())
    op = ops.gn_apply(x, st, batch, h, w, g, b, out, eps=1e-6, silu=True, pad_out=pad)
    ms = timeit(op)
    gb = (x.numel() + out.numel()) * 2 / 1e9
    print(f"gn_apply b={batch} {h}x{w} c={c:4d}                              {ms:9.3f} ms {gb / ms * 1e3:8.1f} GB/s  "
          f"{gb / ms * 1e3 / 6552.6 * 100:5.1f}% of measured HBM copy peak", flush=True)


if __name__ == "__main__":
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    if only == "head":
        a = rb(8 * 482 * 642, 128); wm = rb(3, 9 * 128); out = torch.empty(8 * 480 * 640, 3, device=DEV)
        report("conv3x3 vae head b=8 480x640 128->3", ops.conv3x3(a, wm, 8, 480, 640, bias=torch.zeros(3, device=DEV), out_f32=out))
        sys.exit(0)
    if only == "stats":        # what the fused GroupNorm statistics cost, and the tile order that carries them
        for args in [("unet L0", 112, 60, 80, 320, 320), ("unet L1", 112, 30, 40, 640, 640), ("unet L2", 112, 15, 20, 1280, 1280),
                     ("vae 1/2", 16, 240, 320, 256, 256), ("vae 1/4", 16, 120, 160, 512, 512), ("vae 1/1 swapped", 16, 480, 640, 128, 128)]:
            conv_case(args[0] + " no stats", *args[1:], stats=False)
            conv_case(args[0] + " no stats contiguous", *args[1:], stats=False, tile_order=2)
            conv_case(args[0] + " stats round-robin", *args[1:], tile_order=1)
            conv_case(args[0] + " stats contiguous", *args[1:], tile_order=2)
            for g in (2, 8):
                if (args[-1] // 32) % g == 0:
                    conv_case(args[0] + f" stats contiguous, {g}-channel cells", *args[1:], tile_order=2, stats_group=g)
        sys.exit(0)
    if only == "stages":
        conv_case("vae 1/4 512->512", 8, 120, 160, 512, 512)
        conv_case("vae 1/2 256->256", 8, 240, 320, 256, 256)
        conv_case("unet L2 1280->1280", 112, 15, 20, 1280, 1280)
        sys.exit(0)
    if only == "swap":
        conv_case("vae 1/1 (auto)", 8, 480, 640, 128, 128)
        conv_case("vae 1/1 256->128 (auto)", 8, 480, 640, 256, 128)
        conv_case("vae 1/1 (cg1)", 8, 480, 640, 128, 128, cta_group=1)
        conv_case("vae 1/1 256->128 (cg1)", 8, 480, 640, 256, 128, cta_group=1)
        sys.exit(0)
    if only == "gemm":
        for cg in (1, 2):
            print(f"--- cta_group {cg}")
            gemm_case("square", 8192, 8192, 8192, cta_group=cg)
            gemm_case("ff1 L0", 112 * 4800, 2560, 320, cta_group=cg)
            gemm_case("lin L0", 112 * 4800, 320, 320, cta_group=cg)
            conv_case("unet L0", 16, 60, 80, 320, 320, cta_group=cg)
            conv_case("unet L2", 16, 15, 20, 1280, 1280, cta_group=cg)
            conv_case("vae 1/1", 8, 480, 640, 128, 128, cta_group=cg)
            conv_case("vae 1/2", 8, 240, 320, 256, 256, cta_group=cg)
            conv_case("vae 1/2 512->256", 8, 240, 320, 512, 256, cta_group=cg)
            conv_case("vae 1/4", 8, 120, 160, 512, 512, cta_group=cg)
            conv_case("vae 1/8", 8, 60, 80, 512, 512, cta_group=cg)
        sys.exit(0)
    if only == "elem":
        for args in [(8, 480, 640, 128), (8, 240, 320, 256), (8, 120, 160, 512), (8, 60, 80, 512), (112, 60, 80, 320)]:
            gn_case(*args)
        sys.exit(0)
    if only == "xattn":        # collapsed cross-attention + LN2/LN3, UNet level 0 / 1, 112 images
        for heads, rpg in ((5, 16 * 4800), (10, 16 * 1200)):
            c, ntp = heads * 64, 4
            hs = torch.randn(7 * rpg, c, device=DEV)
            a0, bm = torch.randn(7, heads, ntp, c) * 0.05, torch.randn(7, heads, ntp, c) * 0.3
            ap, ca, bmt = [t.to(DEV) for t in ops.xattn_tables(a0, torch.ones(c), torch.zeros(c), bm, [3, 3, 3, 4, 4, 3, 3], ntp)]
            vec = lambda: torch.randn(c, device=DEV)
            out = torch.empty(7 * rpg, c, device=DEV, dtype=ops.h16())
            op = ops.xattn_fused(hs, ap, ca, bmt, vec(), vec(), vec(), list(range(7)), rpg, heads, ntp, out)
            ms = timeit(op)
            gb = op.bytes / 1e9
            print(f"xattn_fused heads={heads} rows={7 * rpg}                       {ms:9.3f} ms {gb / ms * 1e3:8.1f} GB/s  "
                  f"{gb / ms * 1e3 / 6552.6 * 100:5.1f}% of measured HBM copy peak", flush=True)
        sys.exit(0)
    if only == "res":          # decoder ResNet convs as they run in the plan: padded in, padded out, statistics; conv2 adds the 16-bit residual
        for name, b, h, w, cin, cout in (("vae 1/1", 16, 480, 640, 128, 128), ("vae 1/2", 16, 240, 320, 256, 256),
                                         ("vae 1/4", 16, 120, 160, 512, 512), ("vae 1/8", 16, 60, 80, 512, 512)):
            a = rb(b * (h + 2) * (w + 2), cin)
            wm = rb(cout, 9 * cin)
            res = rb(b * (h + 2) * (w + 2), cout)
            out = torch.empty(b * (h + 2) * (w + 2), cout, device=DEV, dtype=ops.h16())
            bias = torch.zeros(cout, device=DEV)
            for tag, kw in (("conv1 (no residual)", {}), ("conv2 (+ 16-bit residual)", dict(res1=res))):
                st = ops.new_stats(b, cout, DEV, replicas=4)
                report(f"{name} {cin}->{cout} padded out, {tag}",
                       ops.conv3x3(a, wm, b, h, w, bias=bias, out_bf16=out, pad_out=True, stats=st,
                                   stats_rows_per_image=(h + 2) * (w + 2), stats_group=min(8, cout // 32), **kw))
        sys.exit(0)
    if only == "up":           # the VAE decoder's three "nearest 2x + 3x3 conv" layers (four per-parity 2x2 convs each), 16 images
        for h, w, cin, cout in ((60, 80, 512, 512), (120, 160, 512, 512), (240, 320, 256, 256)):
            b = 16
            low = rb(b * (h + 2) * (w + 2), cin)
            wmats = [m.to(DEV).to(ops.h16()) for m in ops.up2x_weight_matrices(torch.randn(cout, cin, 3, 3) * (9 * cin) ** -0.5)]
            out = torch.zeros(b * (2 * h + 2) * (2 * w + 2), cout, device=DEV, dtype=ops.h16())
            st = ops.new_stats(b, cout, DEV, replicas=4)
            four = ops.conv_up2x(low, wmats, b, h, w, bias=torch.zeros(cout, device=DEV), pad_out=True, out_bf16=out, stats=st,
                                 stats_rows_per_image=(2 * h + 2) * (2 * w + 2), stats_group=8)
            class Four:
                flops = sum(o.flops_exec for o in four)
                def run(self):
                    for o in four:
                        o.run()
            report(f"up2x conv (4 launches, executed flops) b={b} {h}x{w} {cin}->{cout}", Four())
            conv_case(f"same-shape 3x3 conv at {h}x{w}", b, h, w, cin, cout)
        sys.exit(0)
    if only == "act":          # GEMMs whose epilogue carries a GELU: GEGLU feed-forward and the per-task MLPs (UNet level 0 / 1)
        from stablemtl_b200.weights import interleave_geglu
        for m, c in ((112 * 4800, 320), (112 * 1200, 640)):
            a, w, bias = rb(m, c), rb(8 * c, c), torch.randn(8 * c, device=DEV)
            wi, bi = interleave_geglu(w, bias)
            out = torch.empty(m, 4 * c, device=DEV, dtype=ops.h16())
            report(f"ff1 geglu m={m} k={c} n={8 * c}", ops.gemm(a, wi, bias=bi, act=L.ACT_GEGLU, out_bf16=out))
            w2, b2 = rb(c, c), torch.randn(c, device=DEV)
            out2 = torch.empty(m, c, device=DEV, dtype=ops.h16())
            report(f"task mlp gelu m={m} k={c} n={c}", ops.gemm(a, w2, bias=b2, act=L.ACT_GELU, out_bf16=out2))
            report(f"linear (no act) m={m} k={c} n={c}", ops.gemm(a, w2, bias=b2, out_bf16=out2))
        sys.exit(0)
    if only == "vattn":        # VAE mid-block attention: one head of 512 channels, 4800 tokens
        for batch in (16, 7 * 16):
            c = 512
            qkv = rb(batch * 4800, 3 * c)
            out = torch.empty(batch * 4800, c, device=DEV, dtype=ops.h16())
            report(f"vae attention d=512 b={batch} ntok=4800 (algorithmic flops)",
                   ops.flash_attn(qkv, batch, 4800, 1, out, 0, c, 2 * c, scale=c ** -0.5, head_dim=c))
        sys.exit(0)
    if only == "attn":
        for args in [(16, 4800, 5), (112, 4800, 5), (112, 1200, 10), (112, 300, 20), (112, 80, 20)]:
            attn_case(*args)
        sys.exit(0)
    gemm_case("square", 8192, 8192, 8192)
    gemm_case("square bn128", 8192, 8192, 8192, block_n=128)
    gemm_case("ff1 L0", 16 * 4800, 2560, 320)
    gemm_case("ff2 L0", 16 * 4800, 320, 1280)
    gemm_case("qkv L0", 16 * 4800, 960, 320)
    gemm_case("lin L0", 16 * 4800, 320, 320)
    gemm_case("ff1 L1", 16 * 1200, 5120, 640)
    gemm_case("lin L2", 16 * 300, 1280, 1280)
    gemm_case("ff1 L2", 112 * 300, 10240, 1280)
    conv_case("unet L0", 16, 60, 80, 320, 320)
    conv_case("unet L0 cat", 16, 60, 80, 960, 320)
    conv_case("unet L1", 16, 30, 40, 640, 640)
    conv_case("unet L2", 16, 15, 20, 1280, 1280)
    conv_case("unet L2 cat", 16, 15, 20, 2560, 1280)
    conv_case("unet L3", 112, 8, 10, 1280, 1280)
    conv_case("vae 1/1", 2, 480, 640, 128, 128)
    conv_case("vae 1/2", 2, 240, 320, 256, 256)
    conv_case("vae 1/4", 4, 120, 160, 512, 512)
    conv_case("vae 1/8", 8, 60, 80, 512, 512)
    attn_case(16, 4800, 5)
    attn_case(16, 1200, 10)
    attn_case(16, 300, 20)
