"""Shared helpers for the plan/emulator (CPU) and CUDA (GPU) parity tests."""
import importlib

import numpy as np
import torch

from oracle import keras_ref as K

PKG = "one-stop-for-covid-19-infection-and-lung-segmentation-plus-classification_b200"
P = importlib.import_module(PKG + ".plan")
G = importlib.import_module(PKG + ".graphs")


def perturbed_params(gname, hw, cin=1, seed=3):
    """oracle init with non-trivial biases / BN affine parameters (fresh init has b=0, gamma=1, beta=0)."""
    params, _ = K.init_params(gname, (hw, hw, cin), seed=seed)
    rng = np.random.default_rng(seed + 100)
    for k in params:
        if k.endswith("bias") or k.endswith("beta"):
            params[k] = (rng.standard_normal(params[k].shape) * 0.1).astype(np.float32)
        if k.endswith("gamma"):
            params[k] = (1 + rng.standard_normal(params[k].shape) * 0.1).astype(np.float32)
        if k.endswith("moving_mean"):
            params[k] = (rng.standard_normal(params[k].shape) * 0.1).astype(np.float32)
        if k.endswith("moving_variance"):
            params[k] = (1 + rng.random(params[k].shape)).astype(np.float32)
    return params


def synth_batch(n, hw, cin=1, seg=True, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.random((n, hw, hw, cin)).astype(np.float32)
    if seg:
        t = np.clip(rng.random((n, hw, hw, 1)) * 1.4 - 0.2, 0, 1).astype(np.float32)   # soft targets (T1H:488)
    else:
        t = (rng.random((n, 1)) > 0.4).astype(np.float32)
    return x, t


def grad_errors(got, want, skip_zero_bias=True, norm="max"):
    """worst tensor of max|got-want| / max|want| (norm="max") or ||got-want||_2 / ||want||_2 (norm="l2");
    convT biases feeding a BN are analytically zero and skipped."""
    worst, who = 0.0, None
    for k, v in want.items():
        if skip_zero_bias and "conv2d_transpose" in k and k.endswith("bias"):
            continue
        if norm == "l2":
            d = float(np.linalg.norm((got[k] - v).ravel()) / (np.linalg.norm(np.ravel(v)) + 1e-12))
        else:
            d = float(np.abs(got[k] - v).max() / (np.abs(v).max() + 1e-12))
        if d > worst:
            worst, who = d, k
    return worst, who


def load_emulator(em, plan, params, x, t):
    """weights, inputs, targets and unit sample weights into a tests/emulator.py memory image laid out by `plan`"""
    import emulator as E
    fp, fs = plan.layout.pack(params)
    em.f32(P.Ref("params", 0), fp.size)[:] = fp
    em.f32(P.Ref("state", 0), fs.size)[:] = fs
    xv = plan.x_view
    xin = x.reshape(-1, x.shape[-1])
    if getattr(plan, "x_pad", 0):                       # channel-padded input tensor (inference plans, plan.py)
        xin = np.concatenate([xin, np.zeros((len(xin), plan.x_pad - xin.shape[1]), xin.dtype)], axis=1)
    em.view(xv.ref, xv.ld, xv.c, x.shape[0] * xv.h * xv.w, xv.dt)[:] = xin.astype(E.NPDT[xv.dt])
    em.f32(plan.target, t.size)[:] = t.reshape(-1)
    em.f32(plan.sample_w, x.shape[0])[:] = 1.0
    return fp, fs
