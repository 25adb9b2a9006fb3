"""Dice parity protocol: same synthetic train/valid split, same init, same epoch budget, dropout on in
all arms; compare the final soft val Dice (T1H:784-790) and the best thresholded Dice (T1H:1206) between the
oracle trained on the CPU (fp32) and the engine in exact (fp32) and tensor (fp16) mode.  |delta| <= 0.5 pt is the bar.

usage: python tools/dice_parity.py [size] [n_slices] [epochs] [batch]   -> prints one JSON line"""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "one-stop-for-covid-19-infection-and-lung-segmentation-plus-classification_b200"
from oracle import keras_ref as K  # noqa: E402
from sklearn.model_selection import train_test_split  # noqa: E402

size, n_slices, epochs, batch = [int(v) for v in (sys.argv[1:5] + ["96", "48", "12", "16"][len(sys.argv) - 1:])]
G = importlib.import_module(PKG + ".graphs")
M = importlib.import_module(PKG + ".model")
LS = importlib.import_module(PKG + ".losses")
S = importlib.import_module(PKG + ".synthetic")
task = os.environ.get("TASK", "lung")
x, t = S.make_slices(n_slices, size, seed=1234, task=task)
xtr, xva, ttr, tva = train_test_split(x, t, test_size=0.3, random_state=42)
params0, _ = K.init_params("unet", (size, size, 1), seed=42)
thresholds = [0.3, 0.4, 0.5, 0.6, 0.7]


def best_dice(tv, pv):
    return max(K.sm_threshold_metrics(tv, pv, th)["f1"] for th in thresholds)


res = {"config": dict(size=size, n=n_slices, epochs=epochs, batch=batch, task=task, train=len(xtr), valid=len(xva))}
# ---- oracle on the CPU (fp32), same batches (no shuffling), same dropout stream
t0 = time.time()
torch.set_num_threads(os.cpu_count() or 1)
p = {k: v.copy() for k, v in params0.items()}
opt = K.Adam(lr=5e-4)
step = 0
for ep in range(epochs):
    ep_stats = []
    for lo in range(0, len(xtr), batch):
        l_, d_ = K.train_step("unet", p, opt, xtr[lo:lo + batch], ttr[lo:lo + batch], dtype=torch.float32, dropout=dict(seed=7, step=step))
        ep_stats.append((l_, d_, len(xtr[lo:lo + batch])))
        step += 1
w_ = np.array([a[2] for a in ep_stats], float)
o_train_loss = float((np.array([a[0] for a in ep_stats]) * w_).sum() / w_.sum())
o_train_dice = float((np.array([a[1] for a in ep_stats]) * w_).sum() / w_.sum())
pv, _ = K.forward("unet", p, xva, training=False, dtype=torch.float32)
res["oracle_cpu_fp32"] = dict(val_dice=float(K.dice_coeff(torch.from_numpy(tva).double(), torch.from_numpy(pv).double())),
                              best_thr_dice=float(best_dice(tva, pv)), train_loss=o_train_loss, train_dice=o_train_dice,
                              seconds=time.time() - t0)
arms = [("float32", 7), ("float16", 7)] + [("float16", int(sd)) for sd in os.environ.get("EXTRA_SEEDS", "").split(",") if sd] \
    + [("float32", int(sd)) for sd in os.environ.get("EXTRA_SEEDS", "").split(",") if sd]
for prec, dseed in arms:
    t0 = time.time()
    m = M.Model(graph=G.unet(size, 1), precision=prec, dropout_seed=dseed)
    m.set_weights_dict(params0)
    m.compile(optimizer=M.Adam(lr=0.0005), loss=LS.bce_dice_loss, metrics=[LS.dice_coeff])
    h = m.fit(xtr, ttr, batch_size=batch, epochs=epochs, validation_data=(xva, tva), shuffle=False, verbose=0)
    pr = m.predict(xva, batch_size=batch)
    res["engine_" + prec + ("" if dseed == 7 else "_dropseed%d" % dseed)] = dict(val_dice=float(h.history["val_dice_coeff"][-1]), best_thr_dice=float(best_dice(tva, pr)),
                                 train_loss=float(h.history["loss"][-1]), train_dice=float(h.history["dice_coeff"][-1]),
                                 seconds=time.time() - t0)
o = res["oracle_cpu_fp32"]
for key in [k for k in res if k.startswith("engine_")]:
    e = res[key]
    e["delta_val_dice_pt"] = 100 * (e["val_dice"] - o["val_dice"])
    e["delta_best_dice_pt"] = 100 * (e["best_thr_dice"] - o["best_thr_dice"])
    e["delta_train_dice_pt"] = 100 * (e["train_dice"] - o["train_dice"])
print(json.dumps(res))
