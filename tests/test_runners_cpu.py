"""Host-side reporting of the runners that needs no device: the cross-validation tables (CV4:1272-1364)."""
import importlib

import numpy as np

from conftest import PKG


def test_cv_report_tables_match_reference_aggregation():
    R = importlib.import_module(PKG + ".runners")
    rng = np.random.default_rng(0)
    the_range = np.arange(0.30, 0.80, 0.05)                                  # CV4:1221
    sweeps = [dict(threshold=the_range, f1=rng.random(10), iou=rng.random(10), precision=rng.random(10),
                   recall=rng.random(10)) for _ in range(4)]
    rep = R.cv_report(sweeps)
    total_dices = np.transpose(np.array([s["f1"] for s in sweeps]))          # CV4:1272-1276 verbatim aggregation
    df = rep["tables"]["dice"]
    assert df.shape == (10, 4) and list(df.columns) == [1, 2, 3, 4] and np.allclose(df.index.to_numpy(), the_range)
    assert np.array_equal(df.to_numpy(), total_dices)
    assert rep["maximum"]["dice"] == np.max(total_dices)
    assert np.array_equal(rep["best_per_split"]["dice"], total_dices.max(axis=0))
    assert rep["best_threshold_per_split"]["dice"] == [float(the_range[total_dices[:, k].argmax()]) for k in range(4)]
    assert np.isclose(rep["mean"]["dice"], df.mean().mean()) and rep["mean"]["f1"] == rep["mean"]["dice"]
    for name in ("iou", "precision", "recall"):
        assert rep["tables"][name].shape == (10, 4)
