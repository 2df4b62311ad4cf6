"""Why is the Cout = 3 conv (VAE decoder head) 5x off its shared-memory bound?  Sweeps tile width / mainloop form."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablemtl_b200 import ops  # noqa: E402
from scripts.bench_kernels import rb, report, DEV  # noqa: E402

b, h, w, c = 8, 480, 640, 128
a = rb(b * (h + 2) * (w + 2), c)
for cout, bn, f32 in [(3, 0, True), (3, 64, True), (3, 128, True), (32, 32, True), (32, 32, False), (64, 64, False), (128, 128, False)]:
    wm = rb(cout, 9 * c)
    if f32:
        out = torch.empty(b * h * w, cout, device=DEV)
        kw = dict(out_f32=out)
    else:
        out = torch.empty(b * h * w, cout, device=DEV, dtype=ops.h16())
        kw = dict(out_bf16=out)
    report(f"head-like cout={cout} block_n={bn} f32={f32}", ops.conv3x3(a, wm, b, h, w, bias=torch.zeros(cout, device=DEV), block_n=bn, cta_group=1, **kw))
