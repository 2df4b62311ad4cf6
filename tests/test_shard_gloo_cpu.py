"""world_size-2 (and 3, ragged) `gloo` test of the N>1 path's host logic: images are sharded over ranks, every
rank runs the path on its slice, the task maps are gathered back in image order.  The per-rank "engine" here is a
deterministic stand-in (the CUDA engine needs a GPU); what is under test is stablemtl_b200/shard.py."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class FakeEngine:
    """maps depend only on the image (like the real path): depth = mean colour, semantic = parity of the first pixel."""
    device = torch.device("cpu")
    tasks = ["depth", "semantic", "optical_flow"]

    def predict(self, rgb, rgb_next=None):
        nxt = rgb if rgb_next is None else rgb_next
        return {"depth": rgb.mean(dim=1, keepdim=True) / 255.0,
                "semantic": (rgb[:, 0] % 2).to(torch.int64),
                "optical_flow": (nxt - rgb)[:, :2] / 255.0}

    def empty_result(self, H, W):
        return {"depth": torch.empty(0, 1, H, W), "semantic": torch.empty(0, H, W, dtype=torch.int64),
                "optical_flow": torch.empty(0, 2, H, W)}


def _worker(rank, world, port, n_images, dst, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from stablemtl_b200.shard import ShardedEngine
        g = torch.Generator().manual_seed(0)
        rgb = torch.randint(0, 256, (n_images, 3, 6, 8), generator=g).float()
        nxt = torch.randint(0, 256, (n_images, 3, 6, 8), generator=g).float()
        full = FakeEngine().predict(rgb, nxt)
        got = ShardedEngine(FakeEngine()).predict(rgb, nxt, gather=True, dst=dst)
        if dst is not None and rank != dst:
            ok = got is None
        else:
            ok = all(torch.equal(got[t], full[t]) for t in full) and all(got[t].shape[0] == n_images for t in full)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,n_images,dst", [(2, 8, None), (2, 5, 0), (3, 2, None)])
def test_sharded_predict_matches_unsharded(world, n_images, dst):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, dst, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert all(res.values()), res
