import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablemtl_b200 import ops  # noqa: E402
from scripts.bench_kernels import rb, report, DEV  # noqa: E402
b, h, w, c = 8, 480, 640, 128
a = rb(b * (h + 2) * (w + 2), c)
for cout, bn in [(32, 32), (128, 128), (256, 256)]:
    wm = rb(cout, 9 * c)
    out = torch.empty(b * h * w, cout, device=DEV, dtype=ops.h16())
    report(f"head-like cout={cout} block_n={bn}", ops.conv3x3(a, wm, b, h, w, bias=torch.zeros(cout, device=DEV), block_n=bn, cta_group=1, out_bf16=out))
