"""Seeded synthetic stand-ins for the Kaggle CT data (no dataset is reachable offline).

Slices are smooth "CT-like" fields: two soft ellipses (lungs) on a dark body plus low-pass noise, in
[0,1]; infection targets are unions of 1-3 blobs inside the lungs rendered as uint8, bilinearly resized
and divided by 255, so their edges are fractional exactly like the reference's masks
(/root/reference/Scripts/task1_preprocessing_plus_unet_with_comments.py:488, 679).  Lung targets
(Task-3) are the filled ellipses; Task-2 labels are Bernoulli(0.765) (1615 / 2112, NB task2 cell 32)
with a class-conditional texture so that AUROC is neither 0.5 nor 1.
"""
import numpy as np


def _smooth_noise(rng, n, size, cells=8):
    coarse = rng.standard_normal((n, cells + 3, cells + 3)).astype(np.float32)
    # bilinear upsample of a coarse grid = cheap low-pass noise
    xs = np.linspace(1, cells + 1, size, dtype=np.float32)
    i0 = np.floor(xs).astype(np.int64)
    f = xs - i0
    rows = coarse[:, i0, :] * (1 - f)[None, :, None] + coarse[:, i0 + 1, :] * f[None, :, None]
    out = rows[:, :, i0] * (1 - f)[None, None, :] + rows[:, :, i0 + 1] * f[None, None, :]
    return out


def _fall(z):
    """1 / (1 + exp(z)) without overflow"""
    return 0.5 * (1.0 - np.tanh(0.5 * z))


def _ellipse(yy, xx, cy, cx, ry, rx):
    return ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2


def make_slices(n, size=512, seed=1234, task="infection"):
    """returns x (n,size,size,1) float32 in [0,1] and target:
    task='infection' -> soft infection mask (n,size,size,1); 'lung' -> lung mask; 'class' -> labels (n,1)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, size, dtype=np.float32), np.linspace(0, 1, size, dtype=np.float32),
                         indexing="ij")
    x = np.empty((n, size, size), np.float32)
    lung = np.empty((n, size, size), np.float32)
    inf = np.zeros((n, size, size), np.float32)
    labels = (rng.random(n) < 0.765).astype(np.float32)
    noise = _smooth_noise(rng, n, size, 12)
    fine = _smooth_noise(rng, n, size, 48)
    for k in range(n):
        j = rng.uniform(-0.03, 0.03, 8)
        e1 = _ellipse(yy, xx, 0.5 + j[0], 0.27 + j[1], 0.30 + j[2], 0.17 + j[3])
        e2 = _ellipse(yy, xx, 0.5 + j[4], 0.73 + j[5], 0.30 + j[6], 0.17 + j[7])
        soft = _fall((np.minimum(e1, e2) - 1) * 12)
        lung[k] = soft
        body = _fall((_ellipse(yy, xx, 0.5, 0.5, 0.46, 0.48) - 1) * 20)
        img = 0.55 * body - 0.40 * soft + 0.08 * noise[k] * body + 0.04 * fine[k]
        has_inf = labels[k] > 0 if task == "class" else True
        if has_inf:
            for _ in range(rng.integers(1, 4)):
                side = 0.27 if rng.random() < 0.5 else 0.73
                cy, cx = 0.5 + rng.uniform(-0.18, 0.18), side + rng.uniform(-0.07, 0.07)
                r = rng.uniform(0.03, 0.09)
                blob = _fall((_ellipse(yy, xx, cy, cx, r, r * rng.uniform(0.7, 1.3)) - 1) * 6)
                inf[k] = np.maximum(inf[k], blob * soft)
            img = img + 0.30 * inf[k] * (0.8 + 0.2 * fine[k])
        x[k] = np.clip(img + 0.12, 0, 1)
    x = np.round(x * 255) / 255                           # uint8-quantised like the reference's /255 inputs
    if task == "class":
        return x[..., None].astype(np.float32), labels[:, None]
    m = inf if task == "infection" else lung
    m = np.round(np.clip(m, 0, 1) * 255) / 255            # soft, fractional edges; never binarised
    return x[..., None].astype(np.float32), m[..., None].astype(np.float32)
