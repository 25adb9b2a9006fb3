"""Callback protocol of the Keras-shaped facade (model.py) that needs no device: the callbacks only talk to
`model.optimizer.lr`, `model.save_weights` and `model.predict_proba`, so a stub model is enough.

  CosineAnnealingScheduler   /root/reference/Scripts/task1_preprocessing_plus_unet_with_comments.py:970-996
  ModelCheckpoint            :1044-1047 (monitor='val_dice_coeff', mode='max', save_best_only=True)
  RocCallback                /root/reference/Scripts/task2_covid19_classifcation.py:706-741
"""
import importlib
import math

import numpy as np

from conftest import PKG

M = importlib.import_module(PKG + ".model")


class _Opt:
    lr = 0.0005


class _StubModel:
    def __init__(self):
        self.optimizer = _Opt()
        self.saved = []

    def save_weights(self, path):
        self.saved.append(path)

    def predict_proba(self, x):
        return 1.0 / (1.0 + np.exp(-np.asarray(x, np.float64).reshape(len(x), -1).mean(1, keepdims=True)))


def test_cosine_annealing_lr_sequence_is_the_reference_formula():
    cb = M.CosineAnnealingScheduler(T_max=7, eta_max=0.0005, eta_min=0.0001, verbose=0)      # T1H:996
    m = _StubModel()
    cb.set_model(m)
    seen = []
    for epoch in range(16):
        logs = {}
        cb.on_epoch_begin(epoch, logs)
        seen.append(m.optimizer.lr)
        cb.on_epoch_end(epoch, logs)
        assert logs["lr"] == m.optimizer.lr                                                    # T1H:988-990
    want = [0.0001 + (0.0005 - 0.0001) * (1 + math.cos(math.pi * e / 7)) / 2 for e in range(16)]   # T1H:984
    assert seen == want
    assert seen[0] == 0.0005 and abs(seen[7] - 0.0001) < 1e-18 and abs(seen[14] - 0.0005) < 1e-18  # period 14, no restart jump


def test_model_checkpoint_best_only_max_and_min():
    m = _StubModel()
    cb = M.ModelCheckpoint("best.h5", monitor="val_dice_coeff", mode="max", save_best_only=True)   # T1H:1046
    cb.set_model(m)
    vals = [0.30, 0.50, 0.45, 0.50, 0.61]
    for e, v in enumerate(vals):
        cb.on_epoch_end(e, {"val_dice_coeff": v, "val_loss": 1 - v})
    assert m.saved == ["best.h5"] * 3 and cb.best == 0.61          # improved at epochs 0, 1, 4 (ties do not save)
    m2 = _StubModel()
    cb = M.ModelCheckpoint("w.h5", monitor="val_loss", save_best_only=True)                         # mode auto -> min
    cb.set_model(m2)
    for e, v in enumerate([0.9, 0.95, 0.7, 0.7]):
        cb.on_epoch_end(e, {"val_loss": v})
    assert len(m2.saved) == 2 and cb.mode == "min"
    m3 = _StubModel()
    cb = M.ModelCheckpoint("every.h5")                                                              # not best-only: every epoch
    cb.set_model(m3)
    for e in range(3):
        cb.on_epoch_end(e, {})
    assert len(m3.saved) == 3
    cb = M.ModelCheckpoint("x.h5", monitor="val_dice_coeff", save_best_only=True)
    assert cb.mode == "max"                                         # auto mode recognises dice / acc / auc monitors
    cb.set_model(_StubModel())
    cb.on_epoch_end(0, {"loss": 1.0})                               # monitored value missing: nothing saved, no error
    assert cb.model.saved == []


def test_roc_callback_logs_auc_and_saves_on_best_validation_auc(capsys):
    from sklearn.metrics import roc_auc_score
    rng = np.random.default_rng(0)
    y = (rng.random(64) < 0.7).astype(np.float64).reshape(-1, 1)
    x = rng.standard_normal((64, 4, 4, 1)) * 0.3 + y.reshape(-1, 1, 1, 1) * 0.4
    yv = (rng.random(32) < 0.7).astype(np.float64).reshape(-1, 1)
    xv = rng.standard_normal((32, 4, 4, 1)) * 0.3 + yv.reshape(-1, 1, 1, 1) * 0.4
    m = _StubModel()
    cb = M.RocCallback(training_data=(x, y), validation_data=(xv, yv), filepath="best_val_auc_weights.h5")   # T2:733
    cb.set_model(m)
    logs = {}
    cb.on_epoch_end(0, logs)
    assert logs["roc_auc"] == roc_auc_score(y, m.predict_proba(x))                 # T2:726-729
    assert logs["val_roc_auc"] == roc_auc_score(yv, m.predict_proba(xv))
    assert m.saved == ["best_val_auc_weights.h5"]
    cb.on_epoch_end(1, {})                                                          # same AUC: not an improvement
    assert len(m.saved) == 1
    assert "roc-auc" in capsys.readouterr().out


def test_sequential_add_builds_the_task2_graph():
    """Sequential().add(...) (T2:747-778): the input layer comes first although it is created inside the first add(),
    auto-names restart per model, and the result is the same graph as graphs.classifier (weights map 1:1)."""
    L = importlib.import_module(PKG + ".layers")
    G = importlib.import_module(PKG + ".graphs")
    for _ in range(2):                                   # a second model in the same process starts from conv2d_1 again
        model = M.Sequential()
        model.add(L.Conv2D(16, (3, 3), activation='relu', padding="same", kernel_initializer="he_normal", input_shape=(224, 224, 1)))
        model.add(L.BatchNormalization())
        model.add(L.Conv2D(16, (3, 3), padding="same", activation='relu', kernel_initializer="he_normal"))
        model.add(L.BatchNormalization())
        model.add(L.MaxPooling2D(pool_size=(2, 2)))
        for ch in (32, 64):
            for _k in range(2):
                model.add(L.Conv2D(ch, (3, 3), padding="same", activation='relu', kernel_initializer="he_normal"))
                model.add(L.BatchNormalization())
            model.add(L.MaxPooling2D(pool_size=(2, 2)))
        model.add(L.Flatten())
        model.add(L.Dense(32, activation='relu'))
        model.add(L.Dropout(0.4))
        model.add(L.Dense(1, activation='sigmoid'))
        assert model.layers[0].kind == "input" and model.layers[1].name == "conv2d_1" and model.layers[-1].name == "dense_2"
        assert model.graph.count_params() == (1678385, 1677937, 448)           # NB task2 cell 73 model.summary()
        ref = G.classifier(224, 1)
        assert [(n, s) for n, s, _, _ in model.graph.weight_specs()] == [(n, s) for n, s, _, _ in ref.weight_specs()]
