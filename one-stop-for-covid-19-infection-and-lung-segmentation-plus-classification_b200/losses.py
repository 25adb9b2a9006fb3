"""Loss / metric tokens with the reference's names.

In the reference these are Keras-backend expressions handed to `model.compile`
(/root/reference/Scripts/task1_preprocessing_plus_unet_with_comments.py:784-799, 1053;
task2_covid19_classifcation.py:688-703, 828).  Here `Model.compile` recognises them by name and maps
them onto fused device kernels (csrc/elementwise.cu: bce_dice_sums / head_bwd / threshold_counts);
calling them directly evaluates the same formula with numpy on host arrays (handy for small checks).
"""
import numpy as np

_EPS = 1e-7


def dice_coeff(y_true, y_pred):
    t, p = np.asarray(y_true, np.float64).ravel(), np.asarray(y_pred, np.float64).ravel()
    return (2.0 * (t * p).sum() + 1.0) / (t.sum() + p.sum() + 1.0)


def dice_loss(y_true, y_pred):
    return 1.0 - dice_coeff(y_true, y_pred)


def binary_crossentropy(y_true, y_pred):
    t = np.asarray(y_true, np.float64)
    p = np.clip(np.asarray(y_pred, np.float64), _EPS, 1 - _EPS)
    return (-(t * np.log(p) + (1 - t) * np.log1p(-p))).mean(axis=-1)


def bce_dice_loss(y_true, y_pred):
    return 0.5 * binary_crossentropy(y_true, y_pred).mean() + 0.5 * dice_loss(y_true, y_pred)


# ---- Task-2 batch metrics (T2:688-703): K.round(K.clip(.,0,1)), eps 1e-7 ---------------------------
def recall(y_true, y_pred):
    t, p = np.asarray(y_true, np.float64), np.asarray(y_pred, np.float64)
    tp = np.rint(np.clip(t * p, 0, 1)).sum()
    return tp / (np.rint(np.clip(t, 0, 1)).sum() + _EPS)


def precision(y_true, y_pred):
    t, p = np.asarray(y_true, np.float64), np.asarray(y_pred, np.float64)
    tp = np.rint(np.clip(t * p, 0, 1)).sum()
    return tp / (np.rint(np.clip(p, 0, 1)).sum() + _EPS)


def f1(y_true, y_pred):
    pr, rc = precision(y_true, y_pred), recall(y_true, y_pred)
    return 2 * ((pr * rc) / (pr + rc + _EPS))


# ---- segmentation_models metrics (T1H:1206-1207): objects carrying a threshold ---------------------
class _SMMetric:
    kind = None

    def __init__(self, threshold=None, smooth=1e-5, name=None):
        self.threshold = 0.5 if threshold is None else float(threshold)
        self.smooth = smooth
        self.__name__ = name or self.default_name

    def from_counts(self, tp, sum_pr, sum_gt):
        s = self.smooth
        fp, fn = sum_pr - tp, sum_gt - tp
        if self.kind == "f1":
            return (2 * tp + s) / (2 * tp + fn + fp + s)
        if self.kind == "iou":
            return (tp + s) / (sum_gt + sum_pr - tp + s)
        if self.kind == "precision":
            return (tp + s) / (tp + fp + s)
        return (tp + s) / (tp + fn + s)

    def __call__(self, y_true, y_pred):
        t = np.asarray(y_true, np.float64)
        pr = (np.asarray(y_pred, np.float64) > self.threshold).astype(np.float64)
        return self.from_counts((t * pr).sum(), pr.sum(), t.sum())


class FScore(_SMMetric):
    kind, default_name = "f1", "f1-score"


class IOUScore(_SMMetric):
    kind, default_name = "iou", "iou_score"


class Precision(_SMMetric):
    kind, default_name = "precision", "precision"


class Recall(_SMMetric):
    kind, default_name = "recall", "recall"
