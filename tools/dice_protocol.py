"""Dice-parity protocol with discriminating power.

North star: "Dice within +-0.5 pt after the same epoch budget".  Round 1's 180-step runs could neither pass nor fail
it: the BatchNorm moving statistics were still 16 % initial values (momentum 0.99) and single runs scatter by several
points.  This protocol trains long enough for the moving statistics to converge (>= 600 steps), repeats every arm over
several seeds (weight init + dropout stream; data, split and per-epoch shuffles are shared) and compares MEANS:

  phase oracle : the CPU oracle (oracle/keras_ref.py, fp32) -- slow, run once where CPU time is free; its per-seed
                 results are committed as a fixture (tests/golden/dice_protocol_*.json) next to this script
  phase engine : the B200 engine (fp16 storage and exact fp32) on the same seeds -> per-seed results, means, deltas

  python tools/dice_protocol.py oracle --size 128 --slices 160 --epochs 45 --batch 8 --seeds 0,1,2,3,4 --out tests/golden/dice_protocol_128.json
  python tools/dice_protocol.py engine --fixture tests/golden/dice_protocol_128.json [--out profiles/...json]

Semantics follow T1H:1059-1101: model.fit(shuffle=True) with per-epoch validation in INFERENCE mode, final
val_dice_coeff (soft, T1H:784-790) and the best thresholded Dice of the sm.metrics sweep (T1H:1206).
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "one-stop-for-covid-19-infection-and-lung-segmentation-plus-classification_b200"
THRESHOLDS = [0.3, 0.4, 0.5, 0.6, 0.7]


def make_data(cfg):
    from sklearn.model_selection import train_test_split
    S = importlib.import_module(PKG + ".synthetic")
    x, t = S.make_slices(cfg["slices"], cfg["size"], seed=1234, task=cfg["task"])
    return train_test_split(x, t, test_size=0.3, random_state=42)                     # T1H:762


def epoch_permutations(n, epochs):
    rng = np.random.RandomState(1234)             # Model._shuffle_rng: the engine's fit draws the same sequence
    return [rng.permutation(n) for _ in range(epochs)]


def run_oracle(cfg, seed, threads):
    import torch
    from oracle import keras_ref as K
    torch.set_num_threads(threads)
    xtr, xva, ttr, tva = make_data(cfg)
    p, _ = K.init_params("unet", (cfg["size"], cfg["size"], 1), seed=42 + seed)
    opt = K.Adam(lr=5e-4)
    step, b = 0, cfg["batch"]
    t0 = time.time()
    for perm in epoch_permutations(len(xtr), cfg["epochs"]):
        stats = []
        for lo in range(0, len(xtr), b):
            idx = perm[lo:lo + b]
            l_, d_ = K.train_step("unet", p, opt, xtr[idx], ttr[idx], dtype=torch.float32, dropout=dict(seed=7 + seed, step=step))
            stats.append((l_, d_, len(idx)))
            step += 1
    w = np.array([s[2] for s in stats], float)
    pv, _ = K.forward("unet", p, xva, training=False, dtype=torch.float32)
    return dict(seed=seed, steps=step, seconds=time.time() - t0,
                train_dice=float((np.array([s[1] for s in stats]) * w).sum() / w.sum()),
                train_loss=float((np.array([s[0] for s in stats]) * w).sum() / w.sum()),
                val_dice=float(K.dice_coeff(torch.from_numpy(tva).double(), torch.from_numpy(pv).double())),
                best_thr_dice=float(max(K.sm_threshold_metrics(tva, pv, th)["f1"] for th in THRESHOLDS)))


def run_engine(cfg, seed, precision):
    G = importlib.import_module(PKG + ".graphs")
    M = importlib.import_module(PKG + ".model")
    LS = importlib.import_module(PKG + ".losses")
    E = importlib.import_module(PKG + ".engine")
    xtr, xva, ttr, tva = make_data(cfg)
    m = M.Model(graph=G.unet(cfg["size"], 1), precision=precision, dropout_seed=7 + seed)
    # the oracle's initialiser stream (same numbers as engine.init_weights draws for the same seed is NOT assumed):
    from oracle import keras_ref as K
    p0, _ = K.init_params("unet", (cfg["size"], cfg["size"], 1), seed=42 + seed)
    m.set_weights_dict(p0)
    m.compile(optimizer=M.Adam(lr=0.0005), loss=LS.bce_dice_loss, metrics=[LS.dice_coeff])
    t0 = time.time()
    h = m.fit(xtr, ttr, batch_size=cfg["batch"], epochs=cfg["epochs"], validation_data=(xva, tva), shuffle=True, verbose=0)
    sw = m.threshold_sweep(xva, tva, THRESHOLDS, batch_size=cfg["batch"])
    out = dict(seed=seed, seconds=time.time() - t0, train_dice=float(h.history["dice_coeff"][-1]),
               train_loss=float(h.history["loss"][-1]), val_dice=float(h.history["val_dice_coeff"][-1]),
               best_thr_dice=float(np.max(sw["f1"])))
    m.engine.close()
    return out


def summary(rows):
    out = {}
    for k in ("train_dice", "val_dice", "best_thr_dice"):
        v = np.array([r[k] for r in rows], float)
        out[k] = dict(mean=float(v.mean()), std=float(v.std(ddof=1)) if len(v) > 1 else 0.0)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("phase", choices=["oracle", "engine"])
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--slices", type=int, default=160)
    ap.add_argument("--epochs", type=int, default=45)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--task", default="lung")
    ap.add_argument("--seeds", default="0,1,2,3,4")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--fixture", default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if a.phase == "oracle":
        cfg = dict(size=a.size, slices=a.slices, epochs=a.epochs, batch=a.batch, task=a.task)
        res = {"config": cfg, "generator": "python tools/dice_protocol.py oracle " + " ".join(sys.argv[2:]), "oracle_cpu_fp32": []}
        for s in [int(v) for v in a.seeds.split(",")]:
            res["oracle_cpu_fp32"].append(run_oracle(cfg, s, a.threads))
            res["oracle_summary"] = summary(res["oracle_cpu_fp32"])
            if a.out:                                   # written after every seed: a long run can be inspected / resumed
                json.dump(res, open(a.out, "w"), indent=1)
            print(json.dumps(res["oracle_cpu_fp32"][-1]), flush=True)
    else:
        res = json.load(open(a.fixture))
        cfg = res["config"]
        seeds = [r["seed"] for r in res["oracle_cpu_fp32"]]
        for prec in ("float16", "float32"):
            rows = [run_engine(cfg, s, prec) for s in seeds]
            res["engine_" + prec] = rows
            sm = summary(rows)
            for k in sm:
                sm[k]["delta_mean_pt"] = 100.0 * (sm[k]["mean"] - res["oracle_summary"][k]["mean"])
            res["engine_%s_summary" % prec] = sm
        if a.out:
            json.dump(res, open(a.out, "w"), indent=1)
        print(json.dumps({k: v for k, v in res.items() if k.endswith("summary")}))


if __name__ == "__main__":
    main()
