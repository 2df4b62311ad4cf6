"""Range audit of one pass: runs every op of the plans one at a time and reports the largest magnitude each 16-bit
output reaches (fp16 saturates at 65504) -- scripts/parity_stages.py tells WHERE an error is made, this tells whether
it is a range problem.   python scripts/debug_range.py [--multi] [--batch 1] [--seed 21] [--precision fp16]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--multi", action="store_true")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--seed", type=int, default=21)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--top", type=int, default=12)
    args = ap.parse_args()
    from stablemtl_b200 import ops, synth
    from stablemtl_b200 import _lib as L
    from stablemtl_b200.pipeline import StableMTLEngine
    ops.set_precision(args.precision)
    ucfg, vcfg = synth.SD2_UNET, synth.SD2_VAE
    child = synth.make_unet_state_dict(ucfg, seed=0)
    vae = synth.make_vae_state_dict(vcfg, seed=2)
    text = synth.make_text_embeddings(ucfg.cross_attention_dim, seed=3)
    main_sd = None
    if args.multi:
        main_sd = dict(synth.make_unet_state_dict(ucfg, seed=10))
        main_sd.update(synth.make_task_modules_state_dict(ucfg, seed=11))
    B, H, W = args.batch, args.height, args.width
    rgb, nxt = synth.make_images(B, H, W, seed=args.seed)
    eng = StableMTLEngine(ucfg, vcfg, child, vae, text, main_sd, use_graph=False)
    eng.predict(rgb.cuda(), nxt.cuda())
    torch.cuda.synchronize()
    p = eng.plan_for(B, H, W, True)
    rows = []
    plans = [("enc", p["enc"].plan)] + [(f"unet{i}", u.plan) for i, u in enumerate(p["unets"])]
    p["enc"].rgb[:B].copy_(rgb.cuda())
    p["enc"].rgb[B:].copy_(nxt.cuda())
    for pname, plan in plans:
        if pname == "unet0":
            p["assemble"].run()
        for i, op in enumerate(plan.ops):
            op.run()
            if op.kind != L.OP_GEMM and op.kind != L.OP_GNAPPLY and op.kind != L.OP_LN and op.kind != L.OP_FATTN:
                continue
            torch.cuda.synchronize()
            for t in op.keep:
                if t is None or t.dtype not in (torch.float16, torch.bfloat16):
                    continue
            outs = []
            st = op.struct
            a = st.args if hasattr(st, "args") else st
            for name in ("out_bf16", "aux_bf16", "out0", "out1"):
                ptr = getattr(a, name, None)
                if not ptr:
                    continue
                for t in op.keep:
                    if t is not None and t.data_ptr() == ptr and t.dtype in (torch.float16, torch.bfloat16):
                        outs.append(t)
            for t in outs:
                f = t.float()
                rows.append((float(f.abs().max()), float(torch.isnan(f).any()), pname, i, op.name, tuple(t.shape)))
    rows.sort(key=lambda r: -r[0])
    print(json.dumps({"config": vars(args), "largest_16bit_outputs": rows[: args.top],
                      "n_saturated": sum(1 for r in rows if r[0] >= 65000), "n_checked": len(rows)}))


if __name__ == "__main__":
    main()
