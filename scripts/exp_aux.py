import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablemtl_b200 import ops  # noqa: E402
from scripts.bench_kernels import rb, report, DEV  # noqa: E402
m, c = 56 * 4800, 320
a, w = rb(m, c), rb(c, c)
bias = torch.zeros(c, device=DEV)
hs = torch.randn(m, c, device=DEV)
out = torch.empty(m, c, device=DEV)
aux = torch.empty(m, c, device=DEV, dtype=ops.h16())
report("attn_out: res1 + out_f32", ops.gemm(a, w, bias=bias, res1=hs, out_f32=out))
report("attn_out: res1 + out_f32 in place", ops.gemm(a, w, bias=bias, res1=hs, out_f32=hs))
report("attn_out: res1 + out_f32 in place + aux16", ops.gemm(a, w, bias=bias, res1=hs, out_f32=hs, aux_bf16=aux))
report("attn_out: out_f32 only", ops.gemm(a, w, bias=bias, out_f32=out))
report("attn_out: out16 only", ops.gemm(a, w, bias=bias, out_bf16=aux))
report("attn_out: out_f32 + aux16", ops.gemm(a, w, bias=bias, out_f32=out, aux_bf16=aux))
