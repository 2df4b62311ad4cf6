"""Host side of the evaluation loop around the accelerated path.

The reference evaluates one image and one task per pipeline call and ships every full-resolution map to numpy before
reducing it (`src/trainer/stablemtl_trainer.py:635-712`, `src/util/alignment.py`, `src/util/metric_semantic.py`).
Here:

* `BatchedEvaluator` pushes whole batches through `StableMTLEngine.predict` with pinned staging buffers, keeps one
  batch in flight (the host consumes batch i-1 while the GPU computes batch i) and has no per-image host sync;
* the data-sized part of the metrics is reduced on the device (`ops.lsq_sums`, `ops.confusion`): a least-squares
  alignment needs 5 sums per image, mIoU a confusion matrix -- the maps themselves never have to leave HBM;
* `save_text_cache` / `load_text_cache` persist the embeddings of the 7 constant task prompts
  (`src/stablemtl_pipeline.py:395-408, 464-472`) so the CLIP text encoder never runs in the loop.
"""
from typing import Dict, Iterable, Iterator, Optional, Tuple

import numpy as np
import torch

from . import ops

F32 = torch.float32


# ------------------------------------------------------------------------------------------------- text-embedding cache
def save_text_cache(path: str, text: Dict[str, torch.Tensor]) -> None:
    """text: {task: [n_tok, 1024]} as `create_text_condition` builds it from the task name (stablemtl_pipeline.py:464-472)"""
    for t, v in text.items():
        if v.dim() != 2:
            raise ValueError(f"text embedding of {t!r} must be [n_tok, dim], got {tuple(v.shape)}")
    torch.save({"format": "stablemtl_b200.text_cache.v1", "text": {t: v.detach().float().cpu() for t, v in text.items()}},
               path)


def load_text_cache(path: str, tasks: Optional[Iterable[str]] = None) -> Dict[str, torch.Tensor]:
    blob = torch.load(path, map_location="cpu", weights_only=True)
    if not isinstance(blob, dict) or blob.get("format") != "stablemtl_b200.text_cache.v1":
        raise ValueError(f"{path} is not a text-embedding cache")
    text = blob["text"]
    if tasks is not None:
        missing = [t for t in tasks if t not in text]
        if missing:
            raise KeyError(f"text cache {path} has no embedding for {missing}")
    return text


# ------------------------------------------------------------------------------------------------- metric finalisation (host)
def lsq_scale_shift(sums) -> Tuple[np.ndarray, np.ndarray]:
    """(n, Sp, Sg, Spp, Spg) per image -> (scale, shift) of lstsq([pred, 1], gt)  (alignment.py:153-157).
    A degenerate system (n < 2 or constant prediction) raises, as numpy's solver would return a rank-deficient fit."""
    s = np.asarray(sums.cpu() if torch.is_tensor(sums) else sums, dtype=np.float64).reshape(-1, 5)
    n, sp, sg, spp, spg = s.T
    det = n * spp - sp * sp
    if np.any(n < 2) or np.any(np.abs(det) <= 1e-12 * np.maximum(n * spp, 1e-300)):
        raise ValueError("least-squares alignment is degenerate (fewer than 2 valid pixels or a constant prediction)")
    scale = (n * spg - sp * sg) / det
    shift = (sg - scale * sp) / n
    return scale, shift


def semantic_scores(hist) -> Tuple[float, float, np.ndarray]:
    """confusion matrix -> (Acc, mIoU, per-class IoU)  (metric_semantic.py:52-70)"""
    h = np.asarray(hist.cpu() if torch.is_tensor(hist) else hist, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        acc = np.diag(h).sum() / h.sum()
        iu = np.diag(h) / (h.sum(axis=1) + h.sum(axis=0) - np.diag(h))
    return float(acc), float(np.nanmean(iu)), iu


class DeviceMetrics:
    """Accumulators living on the GPU; `update_*` launch one reduction kernel each and never synchronise."""

    def __init__(self, device, n_classes: int = 8):
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceMetrics runs on the CUDA device (no CPU fallback)")
        self.device, self.n_classes = torch.device(device), n_classes
        self.hist = torch.zeros(n_classes * n_classes + 1, dtype=torch.int64, device=self.device)

    def depth_alignment_sums(self, pred: torch.Tensor, gt: torch.Tensor, valid: Optional[torch.Tensor]) -> torch.Tensor:
        """pred / gt: fp32 [B, H, W] (or [B, 1, H, W]) on the device; returns fp64 [B, 5] on the device."""
        b = pred.shape[0]
        sums = torch.zeros(b, 5, dtype=torch.float64, device=self.device)
        v = None if valid is None else valid.reshape(b, -1).to(torch.uint8).contiguous()
        ops.lsq_sums(pred.reshape(b, -1).contiguous(), gt.reshape(b, -1).to(F32).contiguous(), v, sums).run()
        return sums

    def update_semantic(self, label_true: torch.Tensor, label_pred: torch.Tensor, valid: Optional[torch.Tensor]) -> None:
        v = None if valid is None else valid.reshape(-1).to(torch.uint8).contiguous()
        ops.confusion(label_true.reshape(-1).to(torch.int64).contiguous(), label_pred.reshape(-1).contiguous(), v,
                      self.hist, self.n_classes).run()

    def confusion_matrix(self) -> np.ndarray:
        h = self.hist.cpu().numpy()
        if h[-1]:
            raise ValueError(f"{int(h[-1])} predicted class ids were outside [0, {self.n_classes})")
        return h[:-1].reshape(self.n_classes, self.n_classes)


# ------------------------------------------------------------------------------------------------- batched driver
class BatchedEvaluator:
    """for (index, maps) in BatchedEvaluator(engine).run(batches): ...

    `batches` yields (rgb, rgb_next_or_None) host tensors [B, 3, H, W] in [0, 255] (uint8 or float).  `maps` is
    {task: numpy array} for the whole batch, identical to what per-image `StableMTLPipeline.__call__`s return
    (depth/shading/albedo in [0,1], unit normals, flows, class ids).  One batch is kept in flight: while the GPU runs
    batch i, the host receives batch i-1 -- the only synchronisation is the event that marks batch i-1's copies done."""

    def __init__(self, engine, keep_on_device: bool = False):
        self.engine, self.keep = engine, keep_on_device
        self._stage = [None, None]                      # two sets of pinned output buffers

    def _pinned_like(self, slot, res):
        if self._stage[slot] is None or any(self._stage[slot][t].shape != v.shape for t, v in res.items()):
            self._stage[slot] = {t: torch.empty(v.shape, dtype=v.dtype).pin_memory() for t, v in res.items()}
        return self._stage[slot]

    def run(self, batches: Iterable[Tuple[torch.Tensor, Optional[torch.Tensor]]]) -> Iterator[Tuple[int, dict]]:
        pending = None                                  # (index, host buffers or device clones, event)
        for i, (rgb, nxt) in enumerate(batches):
            rgb = rgb if rgb.is_cuda or rgb.is_pinned() else rgb.pin_memory()
            if nxt is not None:
                nxt = nxt if nxt.is_cuda or nxt.is_pinned() else nxt.pin_memory()
            res = self.engine.predict(rgb, nxt)         # async: H2D copy + graph replay on the current stream
            if self.keep:
                out = {t: v.clone() for t, v in res.items()}
            else:
                out = self._pinned_like(i & 1, res)
                for t, v in res.items():
                    out[t].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            if pending is not None:
                yield self._finish(pending)
            pending = (i, out, ev)
        if pending is not None:
            yield self._finish(pending)

    def _finish(self, pending):
        i, out, ev = pending
        ev.synchronize()
        if self.keep:
            return i, out
        return i, {t: v.numpy().copy() for t, v in out.items()}
