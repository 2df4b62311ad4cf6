"""Where is the distance to the fp32 oracle made?  (GPU box; the oracle is the checker.)

    python scripts/parity_stages.py [--multi] [--height 480 --width 640] [--batch 1] [--silu 1|2] [--precision fp16|bf16]

Prints one JSON object: per-task relative L2 of the clipped maps, unconditional semantic agreement, and the error of each
stage in isolation -- VAE encode (engine latent vs oracle latent of the same image), UNet (engine task latents vs oracle
task latents), VAE decode alone (the ENGINE's decoder fed the ORACLE's latents vs the oracle's decode) -- plus the
histogram of the oracle's semantic decision margins (what error budget a 99.9 % class-id agreement needs).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PALETTE = [[128, 64, 128], [70, 70, 70], [153, 153, 153], [250, 170, 30], [220, 220, 0], [107, 142, 35],
           [70, 130, 180], [0, 0, 142]]


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--multi", action="store_true")
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--silu", type=int, default=0, help="override ops.SILU_MODE (1 tanh.approx, 2 ex2+rcp)")
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--tiny", action="store_true")
    args = ap.parse_args()
    from oracle import stablemtl_oracle as O
    from stablemtl_b200 import ops, synth
    from stablemtl_b200.pipeline import StableMTLEngine
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ops.set_precision(args.precision)
    if args.silu:
        ops.SILU_MODE = args.silu
    ucfg, vcfg = (synth.TINY_UNET, synth.TINY_VAE) if args.tiny else (synth.SD2_UNET, synth.SD2_VAE)
    child = synth.make_unet_state_dict(ucfg, seed=0)
    vae = synth.make_vae_state_dict(vcfg, seed=2)
    text = synth.make_text_embeddings(ucfg.cross_attention_dim, seed=3)
    main_sd = None
    if args.multi:
        main_sd = dict(synth.make_unet_state_dict(ucfg, seed=10))
        main_sd.update(synth.make_task_modules_state_dict(ucfg, seed=11))
    B, H, W = args.batch, args.height, args.width
    rgb, nxt = synth.make_images(B, H, W, seed=0)
    eng = StableMTLEngine(ucfg, vcfg, child, vae, text, main_sd)
    res, lats = eng.predict(rgb.cuda(), nxt.cuda(), return_latents=True)
    torch.cuda.synchronize()
    lats = {t: v.clone() for t, v in lats.items()}
    clipped_e = {t: v.clone() for t, v in eng.last.items()}
    sem_e = res["semantic"].clone()
    p = eng.plan_for(B, H, W, True)
    enc_lat = p["enc"].out.clone()                                       # [2B*h*w, 4]

    dev = lambda sd: None if sd is None else {k: v.cuda() for k, v in sd.items()}
    orc = O.Oracle(ucfg, vcfg, dev(child), dev(vae), {k: v.cuda() for k, v in text.items()}, dev(main_sd))
    maps, clipped, olat = orc.predict_all(rgb.cuda(), nxt.cuda(), return_latents=True)
    out = {"config": vars(args), "silu_mode": ops.SILU_MODE}
    out["map_rel_l2"] = {t: rel_l2(clipped_e[t], clipped[t]) for t in synth.TASKS}
    out["semantic_agreement"] = (sem_e == maps["semantic"]).float().mean().item()
    out["unet_latent_rel_l2"] = {t: rel_l2(lats[t], olat[t]) for t in synth.TASKS}
    # encode alone
    h, w = p["h"], p["w"]
    both = torch.cat([rgb, nxt]).cuda() / 255.0 * 2.0 - 1.0
    o_enc = O.vae_encode(orc.vae, vcfg, both)                            # [2B, 4, h, w]
    e_enc = enc_lat.view(2 * B, h, w, -1).permute(0, 3, 1, 2)
    out["vae_encode_rel_l2"] = rel_l2(e_enc, o_enc)
    # decode alone: engine decoder on oracle latents
    dec, bd, hw = p["dec"], p["bd"], p["hw"]
    dec_err = {}
    sem_dec_only = None
    for t in synth.TASKS:
        lat_o = olat[t]                                                   # [B, 4, h, w]
        ref = O.vae_decode(orc.vae, vcfg, lat_o)                          # [B, 3, H, W]
        got = []
        flat = lat_o.permute(0, 2, 3, 1).reshape(B * hw, -1).contiguous()
        for c0 in range(0, B, bd):
            n = min(bd, B - c0)
            dec.latent.zero_()
            dec.latent[: n * hw].copy_(flat[c0 * hw:(c0 + n) * hw])
            dec.run()
            torch.cuda.synchronize()
            got.append(dec.out[: n * H * W].view(n, H, W, 3).permute(0, 3, 1, 2).clone())
        got = torch.cat(got)
        dec_err[t] = rel_l2(got, ref)
        if t == "semantic":
            pal = torch.tensor(PALETTE, dtype=torch.float32, device="cuda") / 255.0 * 2.0 - 1.0
            ids = torch.cdist(got.clip(-1, 1).permute(0, 2, 3, 1).reshape(-1, 3), pal).argmin(1).reshape(B, H, W)
            sem_dec_only = (ids == maps["semantic"]).float().mean().item()
    out["vae_decode_only_rel_l2"] = dec_err
    out["semantic_agreement_decode_only"] = sem_dec_only
    # oracle decision margins of the semantic map
    pal = torch.tensor(PALETTE, dtype=torch.float32, device="cuda") / 255.0 * 2.0 - 1.0
    ref3 = clipped["semantic"].float()
    d = torch.cdist(ref3.permute(0, 2, 3, 1).reshape(-1, 3), pal).sort(dim=1).values
    margin = d[:, 1] - d[:, 0]
    out["semantic_margin_cdf"] = {str(m): (margin < m).float().mean().item() for m in (1e-3, 2e-3, 5e-3, 1e-2, 2e-2, 5e-2)}
    out["semantic_map_abs_err_rms"] = (clipped_e["semantic"].double() - clipped["semantic"].double()).pow(2).mean().sqrt().item()
    out["semantic_map_rms"] = clipped["semantic"].double().pow(2).mean().sqrt().item()
    # The same checker against ITSELF at the precision the reference runs at on a GPU with PyTorch's defaults
    # (torch.backends.cudnn.allow_tf32 = True: every cuDNN conv takes TF32 operands; matmuls stay fp32) and with TF32
    # matmuls too: how far a GPU run of the unmodified reference is from its own strict-fp32 result.
    for name, conv_tf32, mm_tf32 in (("reference_tf32_convs_pytorch_default", True, False), ("reference_tf32_convs_and_matmuls", True, True)):
        torch.backends.cudnn.allow_tf32 = conv_tf32
        torch.backends.cuda.matmul.allow_tf32 = mm_tf32
        maps_t, clipped_t, _ = orc.predict_all(rgb.cuda(), nxt.cuda(), return_latents=True)
        out[name] = {"map_rel_l2": {t: rel_l2(clipped_t[t], clipped[t]) for t in synth.TASKS},
                     "semantic_agreement": (maps_t["semantic"] == maps["semantic"]).float().mean().item()}
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    print(json.dumps(out))


if __name__ == "__main__":
    main()
