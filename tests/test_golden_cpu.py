"""The oracle still reproduces the committed golden vectors (tests/golden/make_golden.py), and the
weights regenerated from the seed are the ones the vectors were made with."""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT

FILES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def load(path):
    with np.load(path) as z:
        return {k.replace("__", "/"): z[k] for k in z.files}


def case_of(path):
    g, hw, n = os.path.basename(path)[:-4].split("_")
    return g, int(hw), int(n[1:])


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_reproduces_golden(path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    want = load(path)
    got = mg.build_case(*case_of(path))
    assert set(got) == set(want)
    for k in want:
        assert np.allclose(np.asarray(got[k], np.float64), np.asarray(want[k], np.float64), rtol=1e-6, atol=1e-9), k


def test_golden_present():
    assert len(FILES) >= 4
