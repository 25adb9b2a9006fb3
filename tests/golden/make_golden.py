"""Generates tests/golden/*.npz from the oracle (run in the build container: python tests/golden/make_golden.py).

The reference has no golden vectors, so these pin the ORACLE's outputs: the GPU tests
compare the CUDA path against them without needing to re-run the oracle, and the CPU suite checks the
oracle still reproduces them.  Weights are not stored (31 MB for U-Net): they are regenerated from the
seed, and their checksums are stored to prove the regeneration is identical.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import K, perturbed_params, synth_batch  # noqa: E402

CASES = [("unet", 32, 2), ("unet", 64, 1), ("unetpp", 32, 2), ("classifier", 32, 4)]


def build_case(gname, hw, n):
    seg = gname != "classifier"
    loss = "bce_dice" if seg else "bce"
    params = perturbed_params(gname, hw, seed=11)
    x, t = synth_batch(n, hw, seg=seg, seed=21)
    inf, _ = K.forward(gname, params, x, training=False, dtype=torch.float64)
    r = K.loss_and_grads(gname, params, x, t, dtype=torch.float64, dropout=dict(seed=7, step=2), loss=loss)
    out = dict(x=x, t=t, probs_infer=inf.astype(np.float32), probs_train=r["probs"].astype(np.float32),
               loss=np.float64(r["loss"]), metric=np.float64(r["metric"]))
    out["weight_checksum"] = np.array([float(np.asarray(v, np.float64).sum()) for v in params.values()])
    for k, g in r["grads"].items():
        out["gsum/" + k] = np.float64(g.sum())
        out["gabs/" + k] = np.float64(np.abs(g).sum())
        out["ghead/" + k] = g.reshape(-1)[:8].astype(np.float64)
    for k, v in r["new_moving"].items():
        out["moving/" + k] = v.astype(np.float32)
    return out


if __name__ == "__main__":
    for gname, hw, n in CASES:
        d = build_case(gname, hw, n)
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "%s_%d_n%d.npz" % (gname, hw, n))
        np.savez_compressed(path, **{k.replace("/", "__"): v for k, v in d.items()})
        print(path, os.path.getsize(path))
