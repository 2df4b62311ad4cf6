"""Stage-by-stage error attribution of the CUDA engine against the fp32 oracle (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import stablemtl_oracle as O
from stablemtl_b200 import synth, ops
from stablemtl_b200.engine import UNetPlan, UNetWeights, VAEDecodePlan, VAEEncodePlan, VAEWeights

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
cfgname = sys.argv[1] if len(sys.argv) > 1 else "tiny"
H, Wd = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (64, 96)
ucfg, vcfg = (synth.TINY_UNET, synth.TINY_VAE) if cfgname == "tiny" else (synth.SD2_UNET, synth.SD2_VAE)
dev = "cuda"
def rl2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()
child = synth.make_unet_state_dict(ucfg, 0); vae = synth.make_vae_state_dict(vcfg, 2)
text = synth.make_text_embeddings(ucfg.cross_attention_dim)
cd = {k: v.to(dev) for k, v in child.items()}; vd = {k: v.to(dev) for k, v in vae.items()}
td = {k: v.to(dev) for k, v in text.items()}
B = 2
rgb, nxt = synth.make_images(B, H, Wd, 0)
rgb = rgb.to(dev)
rn = rgb / 255.0 * 2.0 - 1.0
with torch.no_grad():
    lat_o = O.vae_encode(vd, vcfg, rn)                     # [B,4,h,w]
vw = VAEWeights(vae, vcfg, dev)
enc = VAEEncodePlan(vw, B, H, Wd)
enc.rgb.copy_(rgb); enc.run(); torch.cuda.synchronize()
h, w = enc.h, enc.w
lat_m = enc.out.view(B, h, w, 4).permute(0, 3, 1, 2)
print("VAE encode rel-L2:", rl2(lat_m, lat_o), "std", lat_o.std().item())
# UNet on the oracle's latents
x = torch.cat([lat_o, lat_o, torch.zeros_like(lat_o)], 1)
task = "depth"; ti = synth.TASKS.index(task)
with torch.no_grad():
    out_o, feats_o = O.unet_forward(cd, ucfg, x, td[task][None].expand(B, -1, -1))
uw = UNetWeights(child, ucfg, text, synth.TASKS, dev)
up = UNetPlan(uw, B, h, w, [ti], mode="child")
up.x_in.copy_(x.permute(0, 2, 3, 1).reshape(-1, 12)); up.run(); torch.cuda.synchronize()
out_m = up.out.view(B, h, w, 4).permute(0, 3, 1, 2)
print("UNet output rel-L2:", rl2(out_m, out_o), "std", out_o.std().item())
for i, (fm, fo) in enumerate(zip(up.feats_out, feats_o)):
    print(f"  feat[{i}] {tuple(fo.shape)} rel-L2 {rl2(fm.view(fo.shape), fo):.3e}")
# decode on the oracle's UNet latents
with torch.no_grad():
    dec_o = O.vae_decode(vd, vcfg, out_o)
dp = VAEDecodePlan(vw, B, h, w)
dp.latent.copy_(out_o.permute(0, 2, 3, 1).reshape(-1, 4)); dp.run(); torch.cuda.synchronize()
dec_m = dp.out.view(B, dp.H, dp.Wd, 3).permute(0, 3, 1, 2)
print("VAE decode rel-L2:", rl2(dec_m, dec_o), "std", dec_o.std().item())
