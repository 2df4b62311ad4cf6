"""CPU oracle (numpy) of the two evaluation reductions that `stablemtl_b200` pre-reduces on the device.
TEST INFRASTRUCTURE ONLY -- imported by tests/ (never by the product path).  Each function restates the reference:

  align_least_square   src/util/alignment.py:122-169  (align_depth_least_square without the optional down-sampling)
  fast_hist            src/util/metric_semantic.py:45-50
  semantic_scores      src/util/metric_semantic.py:52-70

Pinned against the reference itself by tests/test_metrics_cpu.py::test_oracle_matches_reference_source, which imports
the reference's own functions from /root/reference when that tree is present (the build container) and skips on the
GPU box.
"""
import numpy as np


def align_least_square(gt_arr, pred_arr, valid_mask_arr):
    """alignment.py:129-163: X = lstsq([pred, 1], gt) over valid pixels; returns (aligned, scale, shift)."""
    gt = gt_arr.squeeze()
    pred = pred_arr.squeeze()
    valid = valid_mask_arr.squeeze()
    assert gt.shape == pred.shape == valid.shape
    gt_masked = gt[valid].reshape((-1, 1))
    pred_masked = pred[valid].reshape((-1, 1))
    A = np.concatenate([pred_masked, np.ones_like(pred_masked)], axis=-1)
    X = np.linalg.lstsq(A, gt_masked, rcond=None)[0]
    scale, shift = X
    return (pred_arr * scale + shift).reshape(pred_arr.shape), scale, shift


def fast_hist(label_true, label_pred, n_class):
    """metric_semantic.py:45-50"""
    mask = (label_true >= 0) & (label_true < n_class)
    return np.bincount(n_class * label_true[mask].astype(int) + label_pred[mask], minlength=n_class ** 2
                       ).reshape(n_class, n_class)


def confusion(label_trues, label_preds, valid_masks, n_class):
    """SemanticMetrics.update, metric_semantic.py:34-43, summed over the batch"""
    cm = np.zeros((n_class, n_class))
    for lt, lp, vm in zip(label_trues, label_preds, valid_masks):
        cm += fast_hist(lt[vm], lp[vm], n_class)
    return cm


def semantic_scores(hist):
    """metric_semantic.py:52-70 -> (Acc, mIoU, per-class IoU)"""
    with np.errstate(divide="ignore", invalid="ignore"):
        acc = np.diag(hist).sum() / hist.sum()
        iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist))
    return acc, np.nanmean(iu), iu
