"""Layer-by-layer bisect of the VAE decoder against the fp32 oracle for ONE task latent of the multi-stream pass
(GPU box).   python scripts/debug_decoder.py [--task semantic] [--seed 0]"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="semantic")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--single", action="store_true")
    args = ap.parse_args()
    from oracle import stablemtl_oracle as O
    from stablemtl_b200 import ops, synth
    from stablemtl_b200 import _lib as L
    from stablemtl_b200.pipeline import StableMTLEngine
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ucfg, vcfg = synth.SD2_UNET, synth.SD2_VAE
    child = synth.make_unet_state_dict(ucfg, seed=0)
    vae = synth.make_vae_state_dict(vcfg, seed=2)
    text = synth.make_text_embeddings(ucfg.cross_attention_dim, seed=3)
    main_sd = None
    if not args.single:
        main_sd = dict(synth.make_unet_state_dict(ucfg, seed=10))
        main_sd.update(synth.make_task_modules_state_dict(ucfg, seed=11))
    H, W = 480, 640
    rgb, nxt = synth.make_images(1, H, W, seed=args.seed)
    dev = lambda sd: None if sd is None else {k: v.cuda() for k, v in sd.items()}
    orc = O.Oracle(ucfg, vcfg, dev(child), dev(vae), {k: v.cuda() for k, v in text.items()}, dev(main_sd))
    rn, nn_ = rgb.cuda() / 255.0 * 2.0 - 1.0, nxt.cuda() / 255.0 * 2.0 - 1.0
    _, lat = orc.single_infer(rn, nn_, args.task, {}, return_latent=True)           # [1, 4, h, w]
    sd, G, c = orc.vae, vcfg.norm_num_groups, vcfg.block_out_channels
    # oracle stage list
    stages = []
    x = O._conv(sd, "post_quant_conv", lat / O.LATENT_SCALE, padding=0)
    x = O._conv(sd, "decoder.conv_in", x); stages.append(("conv_in", x))
    p = "decoder.mid_block"
    x = O._vae_resnet(sd, p + ".resnets.0", x, G); stages.append(("mid.res0", x))
    a = p + ".attentions.0"
    B, C, hh, ww = x.shape
    hN = O._gn(sd, a + ".group_norm", x, G, 1e-6).reshape(B, C, hh * ww).transpose(1, 2)
    o = O._attend(O._lin(sd, a + ".to_q", hN), O._lin(sd, a + ".to_k", hN), O._lin(sd, a + ".to_v", hN), 1)
    x = x + O._lin(sd, a + ".to_out.0", o).transpose(1, 2).reshape(B, C, hh, ww); stages.append(("mid.attn", x))
    x = O._vae_resnet(sd, p + ".resnets.1", x, G); stages.append(("mid.res1", x))
    for i in range(len(c)):
        for j in range(vcfg.layers_per_block + 1):
            x = O._vae_resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", x, G); stages.append((f"up{i}.res{j}", x))
        if i < len(c) - 1:
            x = O._conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", F.interpolate(x, scale_factor=2.0, mode="nearest"))
            stages.append((f"up{i}.upconv", x))
    out_ref = O._conv(sd, "decoder.conv_out", F.silu(O._gn(sd, "decoder.conv_norm_out", x, G, 1e-6)))
    stats = [(n, float(t.abs().max()), float(t.mean()), float(t.std())) for n, t in stages]

    eng = StableMTLEngine(ucfg, vcfg, child, vae, text, main_sd, use_graph=False)
    from stablemtl_b200.engine import VAEDecodePlan
    h, w = H // 8, W // 8
    dec = VAEDecodePlan(eng.vae_w, 1, h, w)
    dec.latent.copy_(lat.permute(0, 2, 3, 1).reshape(h * w, -1))
    got = []
    nup = 0
    for op in dec.plan.ops:
        op.run()
        torch.cuda.synchronize()
        if op.kind != L.OP_GEMM:
            continue
        a_ = op.struct.args
        take = op.name in ("vae.dec.conv_in", "vae.conv2", "vae.attn_out")
        if op.name == "vae.dec.up":
            nup += 1
            take = nup % 4 == 0
        if not take:
            continue
        t = [k for k in op.keep if k is not None and k.data_ptr() == a_.out_bf16][0]
        ih, iw = a_.img_h, a_.img_w
        if op.name == "vae.dec.up":
            ih, iw = 2 * ih, 2 * iw
        Cc = t.shape[-1]
        got.append(t.view(1, ih + 2, iw + 2, Cc)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float().clone())
    res = []
    for (n, ref), g in zip(stages, got):
        res.append((n, rel_l2(g, ref)))
    final = dec.out.view(1, H, W, 3).permute(0, 3, 1, 2)
    print(json.dumps({"task": args.task, "n_engine": len(got), "n_oracle": len(stages), "stage_rel_l2": res,
                      "final_rel_l2": rel_l2(final, out_ref), "oracle_stage_absmax_mean_std": stats,
                      "latent_absmax": float(lat.abs().max())}))


if __name__ == "__main__":
    main()
