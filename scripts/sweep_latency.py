"""Latency vs throughput of the multi-stream pass over the GPUs of one box (BASELINE.json configs[4]: 384x1248, batch
sweep; configs[2]: 480x640, batch 64 over 8 GPUs).  Run under torchrun, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        scripts/sweep_latency.py --height 384 --width 1248 --batches 1,8,64

A global batch smaller than the number of GPUs is sharded by TASK STREAM (every rank works on the same images, the
child features are exchanged once: stream_shard.py); larger ones by image (shard.py, no data-path collective).
Prints one JSON line per batch size on rank 0: device-timed (CUDA events), max over ranks."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablemtl_b200 import synth  # noqa: E402
from stablemtl_b200.pipeline import StableMTLEngine  # noqa: E402
from stablemtl_b200.shard import choose_sharding  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=384)
    ap.add_argument("--width", type=int, default=1248)
    ap.add_argument("--batches", default="1,8,64")
    ap.add_argument("--iters", type=int, default=4)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ucfg, vcfg = synth.SD2_UNET, synth.SD2_VAE
    child = synth.make_unet_state_dict(ucfg, 0)
    vae = synth.make_vae_state_dict(vcfg, 2)
    text = synth.make_text_embeddings(ucfg.cross_attention_dim)
    main_sd = dict(synth.make_unet_state_dict(ucfg, 10))
    main_sd.update(synth.make_task_modules_state_dict(ucfg, seed=11))
    H, W = args.height, args.width
    engines = {}

    def engine(stream):
        if stream not in engines:
            engines[stream] = StableMTLEngine(ucfg, vcfg, child, vae, text, main_sd, device=dev, stream_shard=stream)
        return engines[stream]

    for total in [int(b) for b in args.batches.split(",")]:
        stream = choose_sharding(total, world, True) == "streams"
        per_rank = total if stream else max(1, total // world)
        eng = engine(stream)
        g = torch.Generator().manual_seed(100 + (0 if stream else rank))
        rgb = torch.randint(0, 256, (per_rank, 3, H, W), generator=g, dtype=torch.uint8).to(dev)
        nxt = torch.randint(0, 256, (per_rank, 3, H, W), generator=g, dtype=torch.uint8).to(dev)
        for _ in range(3):
            eng.predict(rgb, nxt)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            eng.predict(rgb, nxt)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.iters], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        images = total if stream else per_rank * world
        if rank == 0:
            print(json.dumps({"sweep": "multi-stream all-task maps", "hw": [H, W], "n_gpus": world, "global_batch": images,
                              "sharding": "task streams (child-feature all-gather)" if stream else "images (no collective)",
                              "latency_ms": ms, "images_per_s": images / ms * 1e3}), flush=True)
        eng._plans.clear()
        torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
