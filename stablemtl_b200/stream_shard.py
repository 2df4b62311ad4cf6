"""Task-stream sharding of the multi-stream pass over the GPUs of one box (second axis).

For small batches (the latency end of BASELINE.json configs[4]) there are not enough images to give every GPU its
own, but the 7 task streams of ONE image are independent up to a single exchange: every child UNet pass depends only
on (image, its task) (`src/stablemtl_pipeline.py:475-515`), and the main pass of a task needs the 16 attn1 taps of the
OTHER tasks' child passes (`src/model/attention.py:463-600`).  So:

    rank r owns a contiguous block of tasks        task_range(T, world, r)
    every rank: VAE-encodes the image pair (replicated: ~2 of ~40 TFLOP), runs the child pass of ITS tasks
    exchange : one all-gather per tap layer of the child features (16-bit, 27 MB per stream-image at 480x640)
               into slot-major buffers -- slot s = r * n_max + i holds task_range(r).lo + i, or nothing (-1)
    every rank: main pass + VAE decode + task-map epilogue of ITS tasks
    optional : the finished maps are broadcast from their owners (`gather=True`)

This is the only place on the path with a real exchange step, hence the only data-path collective (NCCL over
NVLink on the GPU box, gloo in the CPU tests).  The host logic here is torch.distributed only; the CUDA engine that
uses it is `StableMTLEngine(stream_group=...)` (pipeline.py).
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .shard import shard_range


def task_range(n_tasks: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the task indices rank `rank` owns (contiguous; earlier ranks take the extra task)."""
    return shard_range(n_tasks, world, rank)


def slots_per_rank(n_tasks: int, world: int) -> int:
    return (n_tasks + world - 1) // world


def task_slots(n_tasks: int, world: int) -> List[int]:
    """Slot-major layout of the gathered taps: slots[r * n_max + i] = task id, or -1 for an empty slot."""
    n_max = slots_per_rank(n_tasks, world)
    slots = []
    for r in range(world):
        lo, hi = task_range(n_tasks, world, r)
        slots += [lo + i if lo + i < hi else -1 for i in range(n_max)]
    return slots


def owner_of(task: int, n_tasks: int, world: int) -> int:
    for r in range(world):
        lo, hi = task_range(n_tasks, world, r)
        if lo <= task < hi:
            return r
    raise ValueError(f"task {task} out of range")


def exchange_taps(local: Sequence[torch.Tensor], gathered: Sequence[torch.Tensor], group=None) -> None:
    """All-gathers every tap layer: local[l] is this rank's [n_max * rows_l, C_l] send buffer (its tasks first, empty
    slots after), gathered[l] the [world * n_max * rows_l, C_l] slot-major receive buffer."""
    world = dist.get_world_size(group)
    into = getattr(dist, "all_gather_into_tensor", None)
    use_into = into is not None and dist.get_backend(group) == "nccl"
    for src, dst in zip(local, gathered):
        if dst.shape[0] != world * src.shape[0] or dst.shape[1:] != src.shape[1:]:
            raise ValueError(f"tap exchange: receive buffer {tuple(dst.shape)} is not world x {tuple(src.shape)}")
        if use_into:
            into(dst, src, group=group)
        else:
            dist.all_gather(list(dst.chunk(world, dim=0)), src, group=group)


def gather_task_maps(local: Dict[str, torch.Tensor], tasks: Sequence[str], like: Dict[str, Tuple[tuple, torch.dtype]],
                     device, group=None) -> Dict[str, torch.Tensor]:
    """Every rank ends up with every task's map: each map is broadcast from the rank that owns the task.
    `like[task] = (shape, dtype)` describes the maps this rank does not own."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    out = {}
    for ti, t in enumerate(tasks):                     # same order on every rank
        src = owner_of(ti, len(tasks), world)
        if src == rank:
            buf = local[t].contiguous()
        else:
            shape, dtype = like[t]
            buf = torch.empty(shape, dtype=dtype, device=device)
        dist.broadcast(buf, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
        out[t] = buf
    return out
