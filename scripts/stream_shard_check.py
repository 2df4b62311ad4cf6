"""Task-stream sharding on N GPUs (run under torchrun): parity of the sharded multi-stream pass against the
unsharded engine on the same rank-0 GPU, and single-image latency both ways.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/stream_shard_check.py [--height 480 --width 640 --batch 1 --iters 5]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablemtl_b200 import synth  # noqa: E402
from stablemtl_b200.pipeline import StableMTLEngine  # noqa: E402


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--tiny", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ucfg, vcfg = (synth.TINY_UNET, synth.TINY_VAE) if args.tiny else (synth.SD2_UNET, synth.SD2_VAE)
    child = synth.make_unet_state_dict(ucfg, 0)
    vae = synth.make_vae_state_dict(vcfg, 2)
    text = synth.make_text_embeddings(ucfg.cross_attention_dim)
    main_sd = dict(synth.make_unet_state_dict(ucfg, 10))
    main_sd.update(synth.make_task_modules_state_dict(ucfg, seed=11))
    B, H, W = args.batch, args.height, args.width
    rgb, nxt = synth.make_images(B, H, W, seed=7)
    rgb, nxt = rgb.to(dev), nxt.to(dev)

    def timed(fn):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.iters], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    sharded = StableMTLEngine(ucfg, vcfg, child, vae, text, main_sd, device=dev, stream_shard=True)
    got = sharded.predict(rgb, nxt, gather=True)
    got = {t: v.clone() for t, v in got.items()}
    ms_sharded = timed(lambda: sharded.predict(rgb, nxt, gather=True))
    ms_sharded_nogather = timed(lambda: sharded.predict(rgb, nxt, gather=False))
    del sharded
    torch.cuda.empty_cache()
    whole = StableMTLEngine(ucfg, vcfg, child, vae, text, main_sd, device=dev)     # every rank: the unsharded pass
    ref = whole.predict(rgb, nxt)
    ms_whole = timed(lambda: whole.predict(rgb, nxt))
    errs = {}
    for t in synth.TASKS:
        if t == "semantic":
            errs[t] = 1.0 - (got[t] == ref[t]).float().mean().item()
        else:
            errs[t] = rel_l2(got[t], ref[t])
    worst = torch.tensor([max(errs.values())], device=dev, dtype=torch.float64)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"check": "stream_shard", "n_gpus": world, "batch": B, "hw": [H, W],
                          "tasks_per_rank": [len(range(*__import__("stablemtl_b200.stream_shard", fromlist=["x"]).task_range(7, world, r))) for r in range(world)],
                          "max_err_vs_unsharded": worst.item(), "errs_rank0": errs,
                          "ms_unsharded_one_gpu": ms_whole, "ms_stream_sharded_gathered": ms_sharded,
                          "ms_stream_sharded_local_maps": ms_sharded_nogather,
                          "latency_speedup": ms_whole / ms_sharded}), flush=True)
    ok = worst.item() < 1e-2       # run-to-run bound of the engine itself (fp32 atomics order, tests/test_pipeline_gpu.py)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
