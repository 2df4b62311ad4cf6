"""Host logic of the task-stream sharding (stablemtl_b200/stream_shard.py) on CPU with `gloo`, world sizes 2 and 4
(4 leaves an empty exchange slot: 7 streams over 4 ranks = 2,2,2,1).  A stand-in "model" with the same dependency
structure as the real path -- child feature of a stream depends on (image, task); a task's main output depends on
the OTHER streams' features -- is computed sharded (child -> exchange -> main -> broadcast of the maps) and
unsharded, and must agree exactly."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from stablemtl_b200 import stream_shard as SS  # noqa: E402

TASKS = ["normal", "depth", "semantic", "optical_flow", "scene_flow", "albedo", "shading"]
LAYERS = [(6, 4), (3, 8)]            # (rows per stream-image, channels) of the stand-in tap layers


def child_feature(img, task, layer):
    rows, ch = LAYERS[layer]
    g = torch.Generator().manual_seed(1000 * task + 10 * layer)
    return torch.randn(img.shape[0] * rows, ch, generator=g) + img.mean()


def main_output(feats_by_task, task):
    """sum over the other streams of every layer's mean feature"""
    acc = torch.zeros(())
    for t, layers in feats_by_task.items():
        if t != task:
            acc = acc + sum(f.mean() for f in layers)
    return acc.reshape(1, 1).expand(2, 3).contiguous() + task


def test_slot_layout():
    assert SS.task_slots(7, 1) == list(range(7))
    assert SS.task_slots(7, 2) == [0, 1, 2, 3, 4, 5, 6, -1]
    assert SS.task_slots(7, 4) == [0, 1, 2, 3, 4, 5, 6, -1]
    assert SS.task_slots(7, 8) == [0, 1, 2, 3, 4, 5, 6, -1]
    assert SS.task_slots(7, 3) == [0, 1, 2, 3, 4, -1, 5, 6, -1]
    assert [SS.task_range(7, 8, r) for r in (0, 6, 7)] == [(0, 1), (6, 7), (7, 7)]
    assert [SS.owner_of(t, 7, 4) for t in range(7)] == [0, 0, 1, 1, 2, 2, 3]
    for world in (1, 2, 3, 4, 8):
        slots = SS.task_slots(7, world)
        assert sorted(s for s in slots if s >= 0) == list(range(7)) and len(slots) == world * SS.slots_per_rank(7, world)
        assert len(slots) <= 9          # SMTL_MAX_TASKS is 8: worlds that need more slots are rejected by the engine


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T = len(TASKS)
        img = torch.arange(2 * 3 * 4 * 5, dtype=torch.float32).reshape(2, 3, 4, 5)
        lo, hi = SS.task_range(T, world, rank)
        n_max = SS.slots_per_rank(T, world)
        slots = SS.task_slots(T, world)
        send, recv = [], []
        for layer, (rows, ch) in enumerate(LAYERS):
            rpg = img.shape[0] * rows
            s = torch.zeros(n_max * rpg, ch)
            for i, t in enumerate(range(lo, hi)):
                s[i * rpg:(i + 1) * rpg] = child_feature(img, t, layer)
            send.append(s)
            recv.append(torch.full((world * n_max * rpg, ch), float("nan")))
        SS.exchange_taps(send, recv)
        feats = {}
        for si, t in enumerate(slots):
            if t >= 0:
                feats[t] = [recv[layer][si * img.shape[0] * LAYERS[layer][0]:(si + 1) * img.shape[0] * LAYERS[layer][0]]
                            for layer in range(len(LAYERS))]
        local = {TASKS[t]: main_output(feats, t) for t in range(lo, hi)}
        like = {name: ((2, 3), torch.float32) for name in TASKS}
        got = SS.gather_task_maps(local, TASKS, like, torch.device("cpu"))
        full_feats = {t: [child_feature(img, t, layer) for layer in range(len(LAYERS))] for t in range(T)}
        ok = all(torch.equal(got[TASKS[t]], main_output(full_feats, t)) for t in range(T))
        ok = ok and all(torch.equal(feats[t][layer], full_feats[t][layer]) for t in range(T) for layer in range(len(LAYERS)))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4])
def test_stream_sharded_matches_unsharded(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert all(res.values()), res


def test_exchange_rejects_mismatched_buffers():
    # shape check happens before any collective is issued, so it is testable without a process group of size > 1
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        with pytest.raises(ValueError):
            SS.exchange_taps([torch.zeros(4, 2)], [torch.zeros(5, 2)])
        SS.exchange_taps([torch.ones(4, 2)], [r := torch.zeros(4, 2)])
        assert torch.equal(r, torch.ones(4, 2))
    finally:
        dist.destroy_process_group()
