"""Epilogue cost experiment on the K = N = 320 token linear (M = 112 * 4800): which epilogue feature costs what."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablemtl_b200 import ops, _lib as L  # noqa
from scripts.bench_kernels import timeit, rb, DEV  # noqa

m, c = 112 * 4800, 320
a, w, bias = rb(m, c), rb(c, c), torch.randn(c, device=DEV)
res16, res32 = rb(m, c), torch.randn(m, c, device=DEV)
o16 = torch.empty(m, c, device=DEV, dtype=ops.h16())
o32 = torch.empty(m, c, device=DEV)
st = ops.new_stats(112, c, DEV)
cases = {
    "out16": dict(out_bf16=o16),
    "out16+bias": dict(out_bf16=o16, bias=bias),
    "out16+bias+res16": dict(out_bf16=o16, bias=bias, res1=res16),
    "out16+bias+res16+stats": dict(out_bf16=o16, bias=bias, res1=res16, stats=st, stats_rows_per_image=4800),
    "out16+stats": dict(out_bf16=o16, stats=st, stats_rows_per_image=4800),
    "out32": dict(out_f32=o32),
    "out32+bias+res32": dict(out_f32=o32, bias=bias, res1=res32),
    "out16 bn=64": dict(out_bf16=o16, block_n=64),
    "out16 bn=128": dict(out_bf16=o16, block_n=128),
    "out16 bn=192": dict(out_bf16=o16, block_n=192),
}
for name, kw in cases.items():
    op = ops.gemm(a, w, **kw)
    ms = timeit(op)
    print(f"{name:28s} {ms:8.3f} ms  {2 * m * c * c / ms / 1e9:7.1f} TFLOP/s", flush=True)
# K sweep at N = 320: where does the mainloop start to matter
for k in (320, 640, 1280, 2560):
    a2, w2 = rb(m, k), rb(c, k)
    ms = timeit(ops.gemm(a2, w2, out_bf16=o16))
    print(f"k={k:5d} out16               {ms:8.3f} ms  {2 * m * c * k / ms / 1e9:7.1f} TFLOP/s", flush=True)
