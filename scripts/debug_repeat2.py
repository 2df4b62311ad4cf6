import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablemtl_b200 import synth, ops
from stablemtl_b200.engine import VAEWeights, VAEEncodePlan

def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

vcfg = synth.TINY_VAE
vae = synth.make_vae_state_dict(vcfg, seed=2)
W = VAEWeights(vae, vcfg, torch.device("cuda"))
rgb, nxt = synth.make_images(3, 64, 96, seed=9)
plan = VAEEncodePlan(W, 3, 64, 96)
plan.rgb.copy_(rgb.cuda())
snaps = []
for rep in range(3):
    cur = []
    for op in plan.plan.ops:
        op.run()
        torch.cuda.synchronize()
        cur.append([None if (t is None or not torch.is_tensor(t)) else t.clone() for t in op.keep])
    snaps.append(cur)
for rep in (1, 2):
    shown = 0
    for i, op in enumerate(plan.plan.ops):
        worst = 0.0
        for a, b in zip(snaps[rep][i], snaps[0][i]):
            if a is not None and a.numel() > 0 and a.dtype in (torch.float32, torch.float16, torch.bfloat16):
                af, bf = a.float(), b.float()
                if torch.isnan(af).any() or torch.isnan(bf).any():
                    continue
                worst = max(worst, rel(af, bf))
        if worst > 1e-6:
            print(f"rep {rep} op {i:3d} {op.name:18s} kind {op.kind} max rel diff among its tensors {worst:.3e}", flush=True)
            shown += 1
            if shown > 6:
                break
