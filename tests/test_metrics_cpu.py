"""CPU tests of the evaluation-side host logic (stablemtl_b200/evaluate.py) and of its oracle (oracle/metrics_oracle.py):
the oracle against the reference's own functions (build container only), the host finalisation (sums -> scale/shift,
histogram -> mIoU) against the oracle, the text-embedding cache round trip."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import metrics_oracle as MO  # noqa: E402

REF = "/root/reference/src/util"


def _load_ref(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _case(seed, h=37, w=53):
    rng = np.random.default_rng(seed)
    pred = rng.random((h, w), dtype=np.float32)
    gt = (2.5 * pred + 0.7 + 0.05 * rng.standard_normal((h, w))).astype(np.float32)
    valid = rng.random((h, w)) > 0.3
    return pred, gt, valid


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_oracle_matches_reference_source():
    try:
        ref_sem = _load_ref("metric_semantic")
        ref_align = _load_ref("alignment")
    except Exception as e:                                   # a missing third-party import of the reference module
        pytest.skip(f"reference module not importable here: {e}")
    pred, gt, valid = _case(0)
    _, s_ref, t_ref = ref_align.align_depth_least_square(gt, pred, valid, return_scale_shift=True)
    _, s, t = MO.align_least_square(gt, pred, valid)
    assert np.array_equal(np.asarray(s_ref), np.asarray(s)) and np.array_equal(np.asarray(t_ref), np.asarray(t))
    rng = np.random.default_rng(1)
    lt = rng.integers(-1, 9, (2, 20, 30))
    lp = rng.integers(0, 8, (2, 20, 30))
    vm = rng.random((2, 20, 30)) > 0.2
    m = ref_sem.SemanticMetrics(8)
    m.update(lt, lp, vm)
    assert np.array_equal(m.confusion_matrix, MO.confusion(lt, lp, vm, 8))
    scores, _ = m._metrics()
    acc, miou, _ = MO.semantic_scores(m.confusion_matrix)
    assert scores["Acc"] == acc and scores["mIoU"] == miou


def test_scale_shift_from_sums_matches_lstsq():
    from stablemtl_b200.evaluate import lsq_scale_shift
    rows = []
    want = []
    for seed in range(4):
        pred, gt, valid = _case(seed)
        p, g = pred[valid].astype(np.float64), gt[valid].astype(np.float64)
        rows.append([p.size, p.sum(), g.sum(), (p * p).sum(), (p * g).sum()])
        _, s, t = MO.align_least_square(gt, pred, valid)
        want.append((float(s[0]), float(t[0])))
    scale, shift = lsq_scale_shift(np.array(rows))
    for i, (s, t) in enumerate(want):
        assert abs(scale[i] - s) <= 1e-4 * abs(s) and abs(shift[i] - t) <= 1e-4 * max(abs(t), 1e-3)
    with pytest.raises(ValueError):
        lsq_scale_shift(np.array([[1, 0.5, 0.5, 0.25, 0.25]]))          # one pixel
    with pytest.raises(ValueError):
        lsq_scale_shift(np.array([[10, 5.0, 7.0, 2.5, 3.5]]))           # constant prediction (det = 0)


def test_semantic_scores_match_oracle():
    from stablemtl_b200.evaluate import semantic_scores
    rng = np.random.default_rng(3)
    hist = rng.integers(0, 1000, (8, 8)).astype(np.float64)
    hist[5] = 0
    hist[:, 5] = 0                                                       # an absent class -> NaN IoU, ignored by nanmean
    a, m, iu = semantic_scores(hist)
    a2, m2, iu2 = MO.semantic_scores(hist)
    assert a == a2 and m == m2 and np.array_equal(np.isnan(iu), np.isnan(iu2))


def test_text_cache_round_trip(tmp_path):
    from stablemtl_b200.evaluate import load_text_cache, save_text_cache
    text = {"depth": torch.randn(3, 1024), "optical_flow": torch.randn(4, 1024)}
    path = str(tmp_path / "text.pt")
    save_text_cache(path, text)
    back = load_text_cache(path, ["depth", "optical_flow"])
    assert all(torch.equal(back[t], text[t]) for t in text)
    with pytest.raises(KeyError):
        load_text_cache(path, ["depth", "normal"])
    torch.save({"x": 1}, path)
    with pytest.raises(ValueError):
        load_text_cache(path)
    with pytest.raises(ValueError):
        save_text_cache(path, {"depth": torch.randn(1024)})
