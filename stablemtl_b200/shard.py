"""Data-parallel sharding of the all-task pass over one 8xB200 box.

The unit of independence is the image (pair): every image's 7 task maps depend on that image only, so a batch is
split evenly over the ranks (one process per GPU), each rank runs the whole path on its slice with replicated
weights, and there is NO data-path collective.  The only exchange is the gather of the finished task maps to the
caller (`gather_maps`), one NCCL all_gather per task map over NVLink -- <1 % of a step at 480x640.

The reference has no inference-time parallelism to mirror (eval is single-GPU); this module is the
host logic of BASELINE.json's "sharding images over GPUs; NCCL only to gather outputs".  It is backend-agnostic
(`nccl` on the GPU box, `gloo` in the CPU tests).
"""
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_images: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the images rank `rank` owns: contiguous, sizes differ by at most one, earlier ranks get the
    extra image (ragged batches: some ranks may get an empty slice when n_images < world)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if n_images < 0:
        raise ValueError("negative batch")
    base, extra = divmod(n_images, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_images: int, world: int) -> List[int]:
    return [hi - lo for lo, hi in (shard_range(n_images, world, r) for r in range(world))]


def shard_batch(x: Optional[torch.Tensor], world: int, rank: int) -> Optional[torch.Tensor]:
    if x is None:
        return None
    lo, hi = shard_range(x.shape[0], world, rank)
    return x[lo:hi]


def gather_maps(local: Dict[str, torch.Tensor], n_images: int, group=None, dst: Optional[int] = None
                ) -> Optional[Dict[str, torch.Tensor]]:
    """Reassembles {task: [n_local, ...]} from every rank into {task: [n_images, ...]} in the original image order.

    dst=None  -> all_gather: every rank gets the full maps.
    dst=r     -> only rank r gets them (others return None); still one collective per task.
    Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(n_images, world)
    biggest = max(sizes)
    out = {}
    for task in sorted(local):                     # same order on every rank
        t = local[task]
        if t.shape[0] != sizes[rank]:
            raise ValueError(f"{task}: rank {rank} holds {t.shape[0]} images, its shard is {sizes[rank]}")
        pad = t
        if t.shape[0] < biggest:
            pad = torch.zeros((biggest,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            pad[: t.shape[0]] = t
        pad = pad.contiguous()
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        if dst is None or dst == rank:
            out[task] = torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0)
    return out if (dst is None or dst == rank) else None


def choose_sharding(n_images: int, world: int, multi_stream: bool, n_tasks: int = 7, max_slots: int = 8) -> str:
    """"images" | "streams": how a global batch of n_images is spread over `world` ranks.

    Images whenever every rank gets at least one (no data-path collective); for smaller batches of the multi-stream
    model the task streams are sharded instead (stream_shard.py: every rank works on all the images, one exchange of
    the child features per pass) -- measured on 8 GPUs at 384x1248: 1 image 35 ms, 4 images 88 ms sharded by stream
    against 98 ms for a batch of 8 sharded by image (profiles/r01c_sweep_multi_384x1248_8gpu.jsonl).  Stream sharding
    needs world * ceil(n_tasks / world) exchange slots <= max_slots (SMTL_MAX_TASKS)."""
    if world <= 1 or n_images >= world or not multi_stream:
        return "images"
    slots = world * ((n_tasks + world - 1) // world)
    return "streams" if slots <= max_slots else "images"


class ShardedEngine:
    """Runs `engine.predict` on this rank's slice of a global batch and (optionally) gathers the maps.

    `engine` is a `stablemtl_b200.pipeline.StableMTLEngine` (or anything with the same `predict`)."""

    def __init__(self, engine, group=None):
        self.engine, self.group = engine, group

    def _world_rank(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group), dist.get_rank(self.group)
        return 1, 0

    def predict(self, rgb: torch.Tensor, rgb_next: Optional[torch.Tensor] = None, gather: bool = True,
                dst: Optional[int] = None):
        world, rank = self._world_rank()
        n = rgb.shape[0]
        mine, mine_next = shard_batch(rgb, world, rank), shard_batch(rgb_next, world, rank)
        local = self.engine.predict(mine, mine_next) if mine.shape[0] > 0 else {}
        if mine.shape[0] == 0:
            # an empty shard still has to take part in the collective with correctly shaped empties
            local = self.engine.empty_result(rgb.shape[2], rgb.shape[3])
        if not gather or world == 1:
            return local
        return gather_maps(local, n, self.group, dst)
