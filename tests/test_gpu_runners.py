"""The reference's runner surface on the GPU: the six public runner names
(/root/reference/Scripts/app.py:36-57) on tiny synthetic data, and the Keras calls around them that no other test
exercises -- Sequential().add(...), class_weight, RocCallback, ModelCheckpoint (best-only save, then load_weights),
CosineAnnealingScheduler inside fit, get_layer(name).output taps, a C_in = 3 classifier (Task-2's real input depth,
BASELINE configs[4]) with AUROC against the oracle on synthetic labels."""
import importlib
import math

import numpy as np
import pytest
import torch

from conftest import PKG
from helpers import K

pytestmark = pytest.mark.gpu

R = importlib.import_module(PKG + ".runners")
M = importlib.import_module(PKG + ".model")
L = importlib.import_module(PKG + ".layers")
LS = importlib.import_module(PKG + ".losses")
S = importlib.import_module(PKG + ".synthetic")
G = importlib.import_module(PKG + ".graphs")

TINY = dict(new_dim=32, n_synthetic=24, batch_size=8, verbose=0)


@pytest.mark.parametrize("name", ["holdout_runner_unet_infection_segmentation",
                                  "holdout_runner_unetplusplus_infection_segmentation", "runner_lung_segmentation"])
def test_segmentation_holdout_runners(name, tmp_path):
    """T1H:6 / UPP / T3:6 -- split 70/30 (seed 42), compile, fit with best-val-Dice checkpoint, reload, evaluate,
    threshold sweep."""
    out = getattr(R, name)(epochs=3, checkpoint_path=str(tmp_path / "best.h5"), cosine=True, **TINY)
    h = out["history"]
    assert len(h["loss"]) == 3 and len(h["val_dice_coeff"]) == 3 and all(np.isfinite(h["loss"]))
    assert h["lr"][0] == pytest.approx(0.0005) and h["lr"][1] == pytest.approx(K.cosine_annealing_lr(1), rel=1e-6)
    assert h["loss"][-1] < h["loss"][0]                                   # it trains
    assert out["x_valid"].shape == (8, 32, 32, 1)                         # sklearn: ceil(0.3 * 24) = 8 held out
    # the checkpoint holds the best epoch: evaluating the reloaded weights reproduces that epoch's validation Dice
    assert out["val_dice_coeff"] == pytest.approx(max(h["val_dice_coeff"]), rel=1e-4)
    sw = out["sweep"]
    assert len(sw["f1"]) == 9 and np.all((sw["iou"] >= 0) & (sw["iou"] <= 1)) and np.all(sw["f1"] >= sw["iou"] - 1e-9)
    assert out["best_dice"] == pytest.approx(float(np.max(sw["f1"])))
    out["model"].engine.close()


@pytest.mark.parametrize("name,folds,epochs", [("three_fold_runner_unet_infection_segmentation", 3, (2, 1, 1)),
                                               ("four_fold_runner_unet_infection_segmentation", 4, 1)])
def test_cross_validation_runners(name, folds, epochs):
    """CV3:6 / CV4:6 -- KFold(shuffle, seed 42), one model per fold, per-fold threshold sweeps, report tables."""
    out = getattr(R, name)(epochs=epochs, thresholds=[0.3, 0.5, 0.7], **TINY)
    assert len(out["folds"]) == folds
    assert out["tables"]["dice"].shape == (3, folds) and list(out["tables"]["iou"].columns) == list(range(1, folds + 1))
    if folds == 3:
        assert [len(f["history"]["loss"]) for f in out["folds"]] == [2, 1, 1]          # CV3: 80 + 20 + 20 epochs pattern
    assert 0.0 <= out["mean"]["dice"] <= 1.0 and out["maximum"]["dice"] >= out["mean"]["dice"]


def test_runner_classification():
    """T2:6 -- stratified split, balanced class weights, bce + f1 metric, RocCallback, AUROC and thresholded metrics."""
    out = R.runner_classification(new_dim=32, n_synthetic=64, epochs=3, batch_size=16, verbose=0)
    h = out["history"]
    assert len(h["loss"]) == 3 and "val_f1" in h and "val_roc_auc" in h and "roc_auc" in h
    assert out["probs"].shape == (20, 1) and np.all((out["probs"] >= 0) & (out["probs"] <= 1))      # ceil(0.3 * 64)
    assert 0.0 <= out["auroc"] <= 1.0 and out["auroc"] == pytest.approx(h["val_roc_auc"][-1], abs=1e-6)
    assert set(out["metrics@0.81"]) == {"accuracy", "precision", "recall", "f1"}
    assert len(out["class_weight"]) == 2 and out["class_weight"][0] > out["class_weight"][1]        # minority class 0 weighs more
    out["model"].engine.close()


def _task2_sequential(hw, cin, **kw):
    """the reference's declaration, T2:747-778, statement for statement"""
    model = M.Sequential(**kw)
    model.add(L.Conv2D(16, (3, 3), activation='relu', padding="same", kernel_initializer="he_normal", input_shape=(hw, hw, cin)))
    model.add(L.BatchNormalization())
    model.add(L.Conv2D(16, (3, 3), padding="same", activation='relu', kernel_initializer="he_normal"))
    model.add(L.BatchNormalization())
    model.add(L.MaxPooling2D(pool_size=(2, 2)))
    model.add(L.Conv2D(32, (3, 3), padding="same", activation='relu', kernel_initializer="he_normal"))
    model.add(L.BatchNormalization())
    model.add(L.Conv2D(32, (3, 3), padding="same", activation='relu', kernel_initializer="he_normal"))
    model.add(L.BatchNormalization())
    model.add(L.MaxPooling2D(pool_size=(2, 2)))
    model.add(L.Conv2D(64, (3, 3), padding="same", activation='relu', kernel_initializer="he_normal"))
    model.add(L.BatchNormalization())
    model.add(L.Conv2D(64, (3, 3), padding="same", activation='relu', kernel_initializer="he_normal"))
    model.add(L.BatchNormalization())
    model.add(L.MaxPooling2D(pool_size=(2, 2)))
    model.add(L.Flatten())
    model.add(L.Dense(32, activation='relu'))
    model.add(L.Dropout(0.4))
    model.add(L.Dense(1, activation='sigmoid'))
    return model


def _class_data(n, hw, cin, seed):
    x1, y = S.make_slices(n, hw, seed=seed, task="class")
    return np.repeat(x1, cin, axis=3).astype(np.float64), y.astype(np.float64)      # gray replicated


@pytest.mark.parametrize("precision", ["float32", "float16"])
def test_sequential_cin3_classifier_fit_predict_auroc_vs_oracle(precision, tmp_path):
    """Sequential API + class_weight + RocCallback + ModelCheckpoint on a C_in = 3 classifier; afterwards the
    inference-mode probabilities and the AUROC on synthetic labels against the oracle run on the SAME trained
    weights (T2:919-926): 1e-3 max-abs on probabilities, AUROC within 1e-3."""
    from sklearn.metrics import roc_auc_score
    hw, cin, n = 32, 3, 96
    x, y = _class_data(n, hw, cin, seed=5)
    xv, yv = _class_data(64, hw, cin, seed=6)
    model = _task2_sequential(hw, cin, precision=precision, dropout_seed=7)
    assert model.count_params() == sum(v.size for v in K.init_params("classifier", (hw, hw, cin))[0].values())
    model.compile(loss='binary_crossentropy', optimizer=M.Adam(lr=0.0005), metrics=[LS.f1, LS.precision, LS.recall])   # T2:828
    cw = np.array([2.0, 0.65])                                                       # ndarray as in T2:801-803
    ck = str(tmp_path / "cls_best.h5")
    roc = M.RocCallback(training_data=(x, y), validation_data=(xv, yv), filepath=str(tmp_path / "best_val_auc_weights.h5"))
    h = model.fit(x, y, batch_size=32, epochs=3, validation_data=(xv, yv), class_weight=cw, verbose=0,
                  callbacks=[roc, M.ModelCheckpoint(ck, monitor="val_loss", save_best_only=True)])              # T2:834-836
    assert set(["loss", "f1", "val_loss", "val_f1", "val_precision", "val_recall", "roc_auc", "val_roc_auc"]) <= set(h.history) or \
        set(["loss", "val_loss", "val_f1", "val_precision", "val_recall", "roc_auc", "val_roc_auc"]) <= set(h.history)
    # inference parity on the trained weights
    w = model.get_weights_dict()
    want, _ = K.forward("classifier", w, xv.astype(np.float32), training=False, dtype=torch.float32)
    got = model.predict_proba(xv)                                                    # T2:726-728
    assert got.shape == (64, 1)
    assert np.abs(got - want).max() < (2e-5 if precision == "float32" else 1e-3)
    # AUROC moves in steps of 1 / (positives * negatives): allow three swapped pairs on top of the 1e-3 of the north star
    pairs = float(yv.sum() * (len(yv) - yv.sum()))
    assert pairs > 0 and abs(roc_auc_score(yv, got) - roc_auc_score(yv, want)) <= 1e-3 + 3.0 / pairs
    ev = model.evaluate(xv, yv, batch_size=32)
    assert len(ev) == 4 and ev[0] == pytest.approx(h.history["val_loss"][-1], rel=1e-5)
    # checkpoint round trip: best-val-loss weights come back and reproduce that epoch's validation loss
    model.load_weights(ck)
    assert model.evaluate(xv, yv, batch_size=32)[0] == pytest.approx(min(h.history["val_loss"]), rel=1e-4)
    model.engine.close()


def test_class_weight_scales_the_loss_like_keras():
    """fit(class_weight=...) multiplies every sample's cross-entropy by the weight of its class and divides by the
    batch size (Keras weighted mean, T2:836): one step with lr = 0 against the oracle's weighted_bce."""
    hw, n = 32, 16
    x, y = _class_data(n, hw, 1, seed=9)
    params, _ = K.init_params("classifier", (hw, hw, 1), seed=3)
    model = M.Model(graph=G.classifier(hw, 1), precision="float32", dropout_seed=7)
    model.set_weights_dict(params)
    model.compile(loss='binary_crossentropy', optimizer=M.Adam(lr=0.0), metrics=[])
    cw = {0: 3.0, 1: 0.5}
    h = model.fit(x, y, batch_size=n, epochs=1, class_weight=cw, shuffle=False, verbose=0)
    sw = np.array([cw[int(v)] for v in y.ravel()], np.float32)
    r = K.loss_and_grads("classifier", params, x.astype(np.float32), y.astype(np.float32), dtype=torch.float64,
                         dropout=dict(seed=7, step=0), loss="bce", sample_weight=sw)
    assert h.history["loss"][0] == pytest.approx(r["loss"], rel=1e-5)
    model.engine.close()


def test_intermediate_layer_outputs_match_oracle_taps():
    """Model(inputs=model.input, outputs=model.get_layer(name).output).predict(x) (T1H:1386-1405)"""
    hw, n = 32, 3
    x, _ = S.make_slices(n, hw, seed=2)
    params, _ = K.init_params("unet", (hw, hw, 1), seed=4)
    model = M.Model(graph=G.unet(hw, 1), precision="float32")
    model.set_weights_dict(params)
    taps = {}
    K.forward("unet", params, x, training=False, dtype=torch.float32, taps=taps)
    for name in ("conv2d_2", "batch_normalization_1", "max_pooling2d_2", "conv2d_10", "conv2d_transpose_1", "conv2d_18"):
        got = model.intermediate(x, name)
        assert got.shape == taps[name].shape, name
        assert np.abs(got - taps[name]).max() < 1e-4 * max(1.0, float(np.abs(taps[name]).max())), name
    assert model.get_layer("conv2d_10").output.shape == (2, 2, 512)
    model.engine.close()


@pytest.mark.parametrize("gname", ["unet", "unetpp"])
def test_save_weights_h5_to_json_model_from_json_load_weights_round_trip(gname, tmp_path):
    """T1H:1079-1095: model.save_weights('....h5') + model.to_json(), then (as a Keras user would) rebuild the network with
    model_from_json and load the weights: the file is real HDF5 in the Keras layout and the rebuilt model predicts the
    same values bit for bit"""
    hw = 32
    x, _ = S.make_slices(3, hw, seed=8)
    a = M.Model(graph=G.GRAPHS[gname](hw, 1), precision="float32", seed=11)
    pa = a.predict(x)
    wpath, jpath = str(tmp_path / "unet_0.8954_cosine_annealer.h5"), str(tmp_path / "unet_0.8954_cosine_annealer.json")
    a.save_weights(wpath)
    with open(jpath, "w") as json_file:
        json_file.write(a.to_json())
    assert open(wpath, "rb").read(8) == b"\x89HDF\r\n\x1a\n"
    H = importlib.import_module(PKG + ".hdf5")
    tree = H.read(wpath)
    assert [n.decode() for n in tree.attrs["layer_names"]] == [l.name for l in a.layers]
    assert tree["conv2d_1"]["conv2d_1"]["kernel:0"].data.shape == (3, 3, 1, 32)
    b = M.model_from_json(open(jpath).read(), precision="float32", seed=99)      # different init: the weights must come from the file
    assert np.abs(b.predict(x) - pa).max() > 1e-4
    b.load_weights(wpath)
    assert np.array_equal(b.predict(x), pa)
    a.engine.close()
    b.engine.close()


def test_cluster_experiment_activation_tap_pca_kmeans_and_per_cluster_scores():
    """N4 (T1H:1386-1496): bottleneck features through the device activation tap, PCA + KMeans(2) as the reference calls
    them, per-cluster validation scores; the feature matrix equals the oracle's taps in the reference's channel-major order"""
    hw = 32
    x, t = S.make_slices(24, hw, seed=21)
    xv, tv = S.make_slices(12, hw, seed=22)
    params, _ = K.init_params("unet", (hw, hw, 1), seed=4)
    model = M.Model(graph=G.unet(hw, 1), precision="float32")
    model.set_weights_dict(params)
    out = R.cluster_experiment(model, x, xv, tv, layer_name="conv2d_9", n_components=8, batch_size=8)
    assert out["train_labels"].shape == (24,) and set(np.unique(out["train_labels"])) <= {0, 1}
    assert out["valid_labels"].shape == (12,) and out["count_cluster_0"] + out["count_cluster_1"] == 12
    assert 0.0 < out["explained_variance"] <= 1.0 + 1e-9 and len(out["score_all"]) == 3
    for k in (0, 1):
        if out["count_cluster_%d" % k]:
            assert len(out["score_cluster_%d" % k]) == 3 and np.isfinite(out["score_cluster_%d" % k]).all()
    # the weighted mean of the per-cluster losses is the loss of the whole set when every cluster is one batch or less
    taps = {}
    K.forward("unet", params, x[:3], training=False, dtype=torch.float32, taps=taps)
    feat = np.transpose(model.intermediate(x[:3], "conv2d_9"), (0, 3, 1, 2)).reshape(3, -1)
    want = np.transpose(taps["conv2d_9"], (0, 3, 1, 2)).reshape(3, -1)
    assert np.abs(feat - want).max() < 1e-4 * max(1.0, float(np.abs(want).max()))
    model.engine.close()


def test_fp16_inference_plan_tracks_weight_changes_and_taps_layer_outputs():
    """fp16 inference plans fold Conv2D -> BatchNormalization into one kernel and run their weight-only ops (operand packing,
    BN scale / shift) only when needed: predictions must follow set_weights, a training step in between, another batch
    size sharing the arenas, and `intermediate` (a tap plan) must still return each layer's OWN output (T1H:1386-1405)."""
    hw, n = 32, 6
    x, t = S.make_slices(n, hw, seed=12)
    pa, _ = K.init_params("unet", (hw, hw, 1), seed=4)
    pb, _ = K.init_params("unet", (hw, hw, 1), seed=9)
    rng = np.random.default_rng(0)
    for p in (pa, pb):                       # non-trivial moving statistics and affine parameters
        for k in p:
            if k.endswith("moving_variance") or k.endswith("gamma"):
                p[k] = (0.5 + rng.random(p[k].shape)).astype(np.float32)
            elif k.endswith("moving_mean") or k.endswith("beta"):
                p[k] = (0.2 * rng.standard_normal(p[k].shape)).astype(np.float32)
    model = M.Model(graph=G.unet(hw, 1), precision="float16")
    model.compile(optimizer=M.Adam(lr=0.0005), loss=LS.bce_dice_loss, metrics=[LS.dice_coeff])
    want = {}
    for name, p in (("a", pa), ("b", pb)):
        want[name], _ = K.forward("unet", p, x, training=False, dtype=torch.float32)
    model.set_weights_dict(pa)
    assert np.abs(model.predict(x, batch_size=n) - want["a"]).max() < 2e-3
    assert np.abs(model.predict(x, batch_size=n) - want["a"]).max() < 2e-3            # second call: prep ops skipped
    assert np.abs(model.predict(x[:4], batch_size=4) - want["a"][:4]).max() < 2e-3    # another plan, same arenas
    assert np.abs(model.predict(x, batch_size=n) - want["a"]).max() < 2e-3
    model.set_weights_dict(pb)
    assert np.abs(model.predict(x, batch_size=n) - want["b"]).max() < 2e-3
    # a training step changes the weights (and overwrites the shared arenas): the next predict must see them
    before = model.predict(x, batch_size=n)
    for _ in range(3):
        model.train_on_batch(x, t)
    after_w = model.get_weights_dict()
    want_after, _ = K.forward("unet", after_w, x, training=False, dtype=torch.float32)
    got_after = model.predict(x, batch_size=n)
    assert np.abs(got_after - want_after).max() < 2e-3 and np.abs(got_after - before).max() > 1e-4
    # taps: conv2d_2's own (pre-BN) output, the BN output, both against the oracle at fp16 accuracy
    taps = {}
    K.forward("unet", after_w, x, training=False, dtype=torch.float32, taps=taps)
    for name in ("conv2d_2", "batch_normalization_1", "conv2d_10"):
        got = model.intermediate(x, name)
        assert np.abs(got - taps[name]).max() < 4e-3 * max(1.0, float(np.abs(taps[name]).max())), name
    b = model.engine.forward_batch(model._to_dev(x), None, n)
    with pytest.raises(ValueError, match="fused with the BatchNormalization"):
        model.engine.layer_output(b, "conv2d_2")
    assert np.abs(model.predict(x, batch_size=n) - want_after).max() < 2e-3           # after the tap plan used the arenas
    model.engine.close()
