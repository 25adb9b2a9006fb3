"""The oracle against the only facts the reference pins: parameter counts and layer
output shapes of its stored model.summary() tables, plus self-checks of the restated semantics."""
import numpy as np
import pytest
import torch

from oracle import keras_ref as K
from oracle import philox


# NB/task1_crossval_3folds_unet:2091-2093, NB/task1_unet_plus_plus:2124-2126, NB/task2...:1704-1706
@pytest.mark.parametrize("graph,total,trainable,non_trainable", [
    ("unet", 7765281, 7762401, 2880),
    ("unetpp", 2209697, 2207329, 2368),
    ("classifier", 1678385, 1677937, 448),
])
def test_param_counts_match_reference_summaries(graph, total, trainable, non_trainable):
    params, _ = K.init_params(graph, (224, 224, 1))
    assert K.count_params(params) == (total, trainable, non_trainable)


def test_layer_shapes_match_reference_summary():
    # SURVEY Appendix A.1 / notebook summary: skips at 224,112,56,28; bottleneck 14x14x512; output 224x224x1
    _, shapes = K.init_params("unet", (224, 224, 1))
    d = dict(shapes)
    assert d["conv2d_2"] == (1, 224, 224, 32) and d["max_pooling2d_1"] == (1, 112, 112, 32)
    assert d["conv2d_10"] == (1, 14, 14, 512)
    assert d["conv2d_transpose_1"] == (1, 28, 28, 256) and d["concatenate_1"] == (1, 28, 28, 512)
    assert d["concatenate_4"] == (1, 224, 224, 64) and d["conv2d_19"] == (1, 224, 224, 1)
    _, shapes = K.init_params("unetpp", (224, 224, 1))
    d = dict(shapes)
    assert d["concatenate_3"] == (1, 224, 224, 96) and d["concatenate_5"] == (1, 112, 112, 192)
    assert d["concatenate_6"] == (1, 224, 224, 128) and d["conv2d_21"] == (1, 224, 224, 1)
    _, shapes = K.init_params("classifier", (224, 224, 1))
    d = dict(shapes)
    assert d["flatten_1"] == (1, 50176) and d["dense_1"] == (1, 32) and d["dense_2"] == (1, 1)


def test_classifier_rgb_param_count():
    # 1,678,673 with C_in = 3
    params, _ = K.init_params("classifier", (224, 224, 3))
    assert K.count_params(params)[0] == 1678673


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*[np.uint64(c) for c in ctr], key[0], key[1])
        assert tuple(int(g) for g in got) == want


def test_dropout_mask_rate_and_determinism():
    m1 = philox.dropout_keep_mask(200000, 0.25, seed=7, step=3, op_id=1)
    m2 = philox.dropout_keep_mask(200000, 0.25, seed=7, step=3, op_id=1)
    m3 = philox.dropout_keep_mask(200000, 0.25, seed=7, step=4, op_id=1)
    assert (m1 == m2).all() and (m1 != m3).any()
    assert abs(m1.mean() - 0.75) < 5e-3


def test_primitives_known_answers():
    """hand-computed values for the Keras semantics restated in Appendix B"""
    # conv 'same' alignment: a 3x3 kernel that picks the top-left neighbour
    ctx = K.Ctx(params=None, rng=np.random.default_rng(0), dtype=torch.float64)
    x = torch.arange(16, dtype=torch.float64).reshape(1, 1, 4, 4)
    y = K.conv2d(ctx, x, 1, 3, None)
    w = np.zeros((3, 3, 1, 1), np.float32); w[0, 0, 0, 0] = 1
    ctx2 = K.Ctx(params={"conv2d_1/kernel": w, "conv2d_1/bias": np.zeros(1, np.float32)}, dtype=torch.float64)
    y = K.conv2d(ctx2, x, 1, 3, None)[0, 0].numpy()
    assert y[0].tolist() == [0, 0, 0, 0] and y[1].tolist() == [0, 0, 1, 2] and y[3].tolist() == [0, 8, 9, 10]
    # convT 2x2/s2 index map with the (kh,kw,Cout,Cin) kernel: out[2i+a,2j+b] = x[i,j]*W[a,b]
    wt = np.arange(4, dtype=np.float32).reshape(2, 2, 1, 1) + 1
    ctx3 = K.Ctx(params={"conv2d_transpose_1/kernel": wt, "conv2d_transpose_1/bias": np.zeros(1, np.float32)},
                 dtype=torch.float64)
    u = K.conv2d_transpose(ctx3, torch.tensor([[[[1.0, 10.0]]]], dtype=torch.float64), 1)[0, 0].numpy()
    assert u.tolist() == [[1, 2, 10, 20], [3, 4, 30, 40]]
    # fresh BN in inference mode: y = x / sqrt(1 + 1e-3)
    ctx4 = K.Ctx(params=None, rng=np.random.default_rng(0), training=False, dtype=torch.float64)
    v = K.batchnorm(ctx4, torch.ones(1, 2, 2, 2, dtype=torch.float64))
    assert np.allclose(v.numpy(), 1 / np.sqrt(1.001))
    # soft-label BCE with the 1e-7 clip, dice with smooth=1
    t = torch.tensor([0.0, 1.0, 0.5]); p = torch.tensor([0.0, 1.0, 0.5])
    b = K.binary_crossentropy_map(t.double(), p.double()).numpy()
    assert np.allclose(b[:2], -np.log(1 - 1e-7)) and np.isclose(b[2], np.log(2))
    assert np.isclose(float(K.dice_coeff(t.double(), p.double())), (2 * 1.25 + 1) / (1.5 + 1.5 + 1))


def test_adam_first_step_is_lr_sized():
    p = {"w": np.array([1.0, -2.0], np.float32)}
    opt = K.Adam(lr=5e-4)
    opt.step(p, {"w": np.array([0.3, -7.0])})
    assert np.allclose(p["w"], [1.0 - 5e-4, -2.0 + 5e-4], atol=1e-9)


def test_cosine_annealing_values():
    # notebook output "CosineAnnealingScheduler setting learning rate to 0.0005" at epoch 0 (NB ...:2301)
    assert K.cosine_annealing_lr(0) == pytest.approx(5e-4)
    assert K.cosine_annealing_lr(7) == pytest.approx(1e-4)
    assert K.cosine_annealing_lr(14) == pytest.approx(5e-4)


def test_gradients_by_finite_differences():
    torch.manual_seed(0)
    rng = np.random.default_rng(1)
    params, _ = K.init_params("unet", (16, 16, 1), seed=5)
    x = rng.random((2, 16, 16, 1))
    t = np.clip(rng.random((2, 16, 16, 1)) * 1.4 - 0.2, 0, 1)
    r = K.loss_and_grads("unet", params, x, t, dtype=torch.float64, dropout=dict(seed=1, step=0))
    for name, idx in (("conv2d_2/kernel", (1, 1, 3, 5)), ("conv2d_transpose_2/kernel", (0, 1, 7, 9)),
                      ("batch_normalization_3/gamma", (11,)), ("conv2d_19/bias", (0,)), ("conv2d_10/bias", (100,))):
        h = 1e-5
        vals = []
        for s in (+1, -1):
            p2 = {k: v.astype(np.float64).copy() for k, v in params.items()}
            p2[name][idx] += s * h
            vals.append(K.loss_and_grads("unet", p2, x, t, dtype=torch.float64, dropout=dict(seed=1, step=0))["loss"])
        fd = (vals[0] - vals[1]) / (2 * h)
        # ReLU / max-pool kinks inside the +-h interval cost a few 1e-4 relative; a wrong formula costs O(1)
        assert fd == pytest.approx(r["grads"][name][idx], rel=1e-2, abs=1e-9), name


def test_fp32_and_fp64_oracle_agree():
    rng = np.random.default_rng(2)
    params, _ = K.init_params("unet", (32, 32, 1), seed=5)
    x = rng.random((2, 32, 32, 1))
    p64, _ = K.forward("unet", params, x, dtype=torch.float64)
    p32, _ = K.forward("unet", params, x, dtype=torch.float32)
    assert np.abs(p64 - p32).max() < 1e-5


def test_sm_metrics_formula():
    t = np.array([1.0, 1.0, 0.0, 0.5]); p = np.array([0.9, 0.2, 0.8, 0.6])
    m = K.sm_threshold_metrics(t, p, 0.5)
    tp, spr, sgt = 1.5, 3.0, 2.5
    assert m["tp"] == tp and m["sum_pr"] == spr and m["sum_gt"] == sgt
    assert m["iou"] == pytest.approx((tp + 1e-5) / (sgt + spr - tp + 1e-5))
    assert m["f1"] == pytest.approx((2 * tp + 1e-5) / (2 * tp + (sgt - tp) + (spr - tp) + 1e-5))
