import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablemtl_b200 import synth
from stablemtl_b200.pipeline import StableMTLEngine

def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

for multi in (False, True):
  for graph in (False, True):
    ucfg, vcfg = synth.TINY_UNET, synth.TINY_VAE
    child = synth.make_unet_state_dict(ucfg, seed=0); vae = synth.make_vae_state_dict(vcfg, seed=2)
    text = synth.make_text_embeddings(ucfg.cross_attention_dim)
    main = None
    if multi:
        main = dict(synth.make_unet_state_dict(ucfg, seed=10)); main.update(synth.make_task_modules_state_dict(ucfg, seed=11))
    eng = StableMTLEngine(ucfg, vcfg, child, vae, text, main, use_graph=graph)
    rgb, nxt = synth.make_images(3, 64, 96, seed=9)
    outs = []
    for i in range(4):
        res, lats = eng.predict(rgb.cuda(), nxt.cuda(), return_latents=True)
        torch.cuda.synchronize()
        p = eng.plan_for(3, 64, 96, True, torch.float32)
        outs.append(({t: v.clone() for t, v in res.items()}, {t: v.clone() for t, v in lats.items()}, p["enc"].out.clone()))
    for i in range(1, 4):
        print(f"multi={multi} graph={graph} call {i} vs 0: enc {rel(outs[i][2], outs[0][2]):.2e} lat(normal) {rel(outs[i][1]['normal'], outs[0][1]['normal']):.2e} "
              f"lat(flow) {rel(outs[i][1]['optical_flow'], outs[0][1]['optical_flow']):.2e} map(normal) {rel(outs[i][0]['normal'], outs[0][0]['normal']):.2e} map(depth) {rel(outs[i][0]['depth'], outs[0][0]['depth']):.2e}", flush=True)
