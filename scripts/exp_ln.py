import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablemtl_b200 import ops  # noqa: E402
from scripts.bench_kernels import timeit, DEV  # noqa: E402
for m, c in [(112 * 4800, 320), (112 * 1200, 640), (112 * 300, 1280)]:
    x = torch.randn(m, c, device=DEV)
    g, b = torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
    out = torch.empty(m, c, device=DEV, dtype=ops.h16())
    op = ops.layer_norm(x, g, b, out)
    ms = timeit(op)
    gb = m * c * 6 / 1e9
    print(f"layer_norm m={m} c={c}: {ms:.3f} ms {gb / ms * 1e3:.0f} GB/s ({gb / ms * 1e3 / 6552.6 * 100:.1f}% of copy peak)", flush=True)
