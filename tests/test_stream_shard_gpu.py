"""The task-stream-sharded engine on the GPU.  With one visible GPU the exchange degenerates to a world-size-1
all-gather, which still drives the whole mechanism (caller-owned tap buffers, slot-major receive buffers, two CUDA
graphs around the collective); the result must agree with the unsharded engine and the oracle.  With >= 2 GPUs the same
check runs across 2 ranks under torchrun (scripts/stream_shard_check.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_world1_stream_shard_matches_unsharded_and_oracle():
    from oracle import stablemtl_oracle as O
    from stablemtl_b200 import synth
    from stablemtl_b200.pipeline import StableMTLEngine
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        ucfg, vcfg = synth.TINY_UNET, synth.TINY_VAE
        child = synth.make_unet_state_dict(ucfg, 0)
        vae = synth.make_vae_state_dict(vcfg, 2)
        text = synth.make_text_embeddings(ucfg.cross_attention_dim, seed=3)
        main = dict(synth.make_unet_state_dict(ucfg, 10))
        main.update(synth.make_task_modules_state_dict(ucfg, seed=11))
        rgb, nxt = synth.make_images(2, 64, 96, seed=5)
        sharded = StableMTLEngine(ucfg, vcfg, child, vae, text, main, stream_shard=True)
        whole = StableMTLEngine(ucfg, vcfg, child, vae, text, main)
        for _ in range(3):                                   # eager, capture, replay
            got = sharded.predict(rgb.cuda(), nxt.cuda(), gather=True)
            ref = whole.predict(rgb.cuda(), nxt.cuda())
            torch.cuda.synchronize()
            for t in synth.TASKS:
                # same kernels on both sides; the fp32 atomics' order differs from run to run and the random-init net
                # amplifies the flipped 16-bit roundings (see test_repeated_calls_and_batch_change_are_consistent),
                # so the bound is the parity tolerance itself
                if t == "semantic":
                    assert (got[t] == ref[t]).float().mean().item() > 0.99
                else:
                    assert rel_l2(got[t], ref[t]) < 1e-2, t
        orc = O.Oracle(ucfg, vcfg, child, vae, text, main)
        _, clipped, _ = orc.predict_all(rgb, nxt, return_latents=True)
        for t in synth.TASKS:
            assert rel_l2(sharded.last[t], clipped[t]) <= 1e-2, t
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_stream_shard():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "stream_shard_check.py"),
           "--tiny", "--height", "64", "--width", "96", "--batch", "2", "--iters", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
