"""Whole-path parity on the GPU: the engine (planner + CUDA kernels through the C-ABI) against the
oracle and the committed golden vectors -- training step (loss, Dice, probabilities, every parameter
gradient, Adam update, BN moving statistics), inference forward at the reference's real size (224) and
at BASELINE.json's 512, CUDA-graph replay, and a short training run (Dice parity).

Tolerances (stated by BASELINE.json north_star: 1e-3 max-abs on the sigmoid outputs of the forward):
  exact mode (fp32 storage)  : 2e-5 on probabilities, 2e-3 relative (max-norm) on every gradient tensor
  tensor mode (fp16 storage) : 1e-3 on inference probabilities at the real sizes (224 / 512);
                               training-mode steps are compared at 64x64 (batch statistics over 2x2 or 4x4
                               maps at the bottleneck of a 32x32 input are ill-conditioned for ANY 16-bit
                               storage) with 5e-3 on probabilities and 5e-2 relative L2 per gradient tensor.
"""
import glob
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import PKG, ROOT
from helpers import G, K, P, grad_errors, perturbed_params, synth_batch

pytestmark = pytest.mark.gpu

E = importlib.import_module(PKG + ".engine")
PTOL = {"float32": 2e-5, "float16": 1e-3}
GTOL = {"float32": 2e-3, "float16": 5e-2}


def engine_for(gname, hw, precision, params, cin=1, use_graph=False, **kw):
    loss = "bce" if gname == "classifier" else "bce_dice"
    eng = E.Engine(G.GRAPHS[gname](hw, cin), precision=precision, use_graph=use_graph, loss=loss, dropout_seed=7, **kw)
    eng.set_weights(params)
    return eng


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a.reshape(len(a), -1), dtype=np.float32)).cuda()


TRAIN_CASES = [("float32", "unet", 32, 2), ("float32", "unet", 48, 3), ("float32", "unetpp", 32, 2),
               ("float32", "classifier", 32, 4), ("float16", "unet", 64, 4), ("float16", "unet", 96, 2),
               ("float16", "unetpp", 64, 2), ("float16", "classifier", 64, 8)]


@pytest.mark.parametrize("precision,gname,hw,n", TRAIN_CASES)
def test_train_step_vs_oracle(gname, hw, n, precision):
    seg = gname != "classifier"
    loss = "bce_dice" if seg else "bce"
    params = perturbed_params(gname, hw)
    x, t = synth_batch(n, hw, seg=seg)
    eng = engine_for(gname, hw, precision, params)
    eng._set_fields(step=5)
    sw = torch.ones(n, device="cuda")
    b = eng.train_batch(dev(x).view(n, hw, hw, 1), dev(t), None, n, sw_src=sw)
    eng.stream.synchronize()
    r = K.loss_and_grads(gname, params, x, t, dtype=torch.float64, dropout=dict(seed=7, step=5), loss=loss)
    lo = eng.loss_dev(b).cpu().numpy()
    probs = eng.probs(b).cpu().numpy().reshape(r["probs"].shape)
    perr = np.abs(probs - r["probs"]).max()
    ptol = 2e-5 * (3 if not seg else 1) if precision == "float32" else 3e-2
    ls = eng._cur_ls
    grads = {k: v / ls for k, v in eng.get_grads().items()}
    worst, who = grad_errors(grads, r["grads"], norm="max" if precision == "float32" else "l2")
    print("%s %s %d: max|dp| %.3e, worst grad err %.3e (%s), loss %.6f vs %.6f" % (precision, gname, hw, perr, worst, who, lo[0], r["loss"]))
    assert perr < ptol
    assert lo[0] == pytest.approx(r["loss"], abs=20 * PTOL[precision])
    if seg:
        assert lo[1] == pytest.approx(r["metric"], abs=5 * PTOL[precision])
    if precision == "float32":
        assert worst < GTOL[precision], (who, worst)
    else:
        # 16-bit storage in TRAINING mode: batch-norm divides by the batch sigma, which amplifies the fp16
        # rounding of its input by |mean|/sigma per channel, and ReLU / max-pool decisions flip on the
        # perturbed values (tests/emulator.py with fp16 storage reproduces the same 5-30 % L2 error on the
        # cancelling sums -- biases, BN affine gradients -- on the CPU).  So the gradient is checked as a
        # direction: cosine >= 0.95 on every tensor of >= 512 weights, relative L2 <= 0.5 everywhere.
        for k, g in r["grads"].items():
            if "conv2d_transpose" in k and k.endswith("bias"):
                continue
            a, bb = grads[k].ravel().astype(np.float64), g.ravel()
            rel = np.linalg.norm(a - bb) / (np.linalg.norm(bb) + 1e-30)
            assert rel < 0.5, (k, rel)
            if g.size >= 512:
                cos = float(a @ bb / (np.linalg.norm(a) * np.linalg.norm(bb) + 1e-30))
                assert cos > 0.95, (k, cos)
    assert not eng.overflowed()
    if precision == "float32":
        want = {k: v.copy() for k, v in params.items()}
        K.Adam().step(want, r["grads"])
        want.update(r["new_moving"])
        new = eng.get_weights()
        for k in want:
            if "conv2d_transpose" in k and k.endswith("bias"):
                continue
            assert np.abs(new[k] - want[k]).max() < 5e-5, k
    assert eng._pull_state().step == 6
    eng.close()


@pytest.mark.parametrize("gname,hw,n", [("unet", 64, 4), ("unet", 96, 2), ("unetpp", 64, 2)])
def test_fp16_train_step_vs_fp16_emulator(gname, hw, n):
    """Separates fp16 ROUNDING from kernel bugs.  The engine's own op list is interpreted on the
    CPU by tests/emulator.py with the same rounding points (fp16 activations / gradients / packed conv kernels, fp32
    accumulation).  Rounding decisions and ReLU / max-pool flips still decorrelate any two correct fp16 runs within
    a few layers and the BatchNorm chain amplifies that towards the first layers, so the emulator is run three times
    -- plain, and twice with summation-order-sized jitter (2^-22 relative) on every conv accumulator -- and the spread
    between those runs is the MEASURED noise floor of each gradient tensor.  The GPU must stay within 3x that floor
    (+5e-3) of the plain run: a kernel bug would stick out of the rounding noise, where the fp64-oracle comparison
    (rel L2 < 0.5) could not tell them apart.  Loss / Dice within 3e-4, probabilities within 3x their floor + 1e-3."""
    import emulator as Em
    from helpers import load_emulator
    params = perturbed_params(gname, hw)
    x, t = synth_batch(n, hw, seg=True)
    eng = engine_for(gname, hw, "float16", params)
    eng._set_fields(step=5)
    b = eng.train_batch(dev(x).view(n, hw, hw, 1), dev(t), None, n)
    eng.stream.synchronize()
    plan, ls = b.plan, eng._cur_ls
    lo_g = eng.loss_dev(b).cpu().numpy()
    pg = eng.probs(b).cpu().numpy().reshape(t.shape)
    gg = eng.get_grads()
    eng.close()
    runs = []
    for seed in (None, 1, 2):
        em = Em.Emulator(plan.arena_sizes())
        em.fp16_weights = True
        em.jitter = None if seed is None else np.random.default_rng(seed)
        em.state.update(seed=7, step=5, loss_scale=ls)
        fp, _ = load_emulator(em, plan, params, x, t)
        em.run(plan.train_ops())
        runs.append((em.f32(plan.loss_out, 2).copy(), em.f32(plan.prob, t.size).reshape(t.shape).copy(),
                     plan.layout.unpack(em.f32(P.Ref("grads", 0), fp.size), None)))
    rel = lambda a, bb: float(np.linalg.norm(a.ravel().astype(np.float64) - bb.ravel()) / (np.linalg.norm(bb.ravel().astype(np.float64)) + 1e-30))
    (lo_e, pe, ge) = runs[0]
    pfloor = max(np.abs(runs[1][1] - pe).max(), np.abs(runs[2][1] - pe).max(), np.abs(runs[1][1] - runs[2][1]).max())
    report, bad = [], []
    for k in ge:
        if "conv2d_transpose" in k and k.endswith("bias"):
            continue                      # analytically zero: both sides hold rounding noise only
        floor = max(rel(runs[1][2][k], ge[k]), rel(runs[2][2][k], ge[k]), rel(runs[1][2][k], runs[2][2][k]))
        d = rel(gg[k], ge[k])
        report.append((d, floor, k))
        if d > 3.0 * floor + 5e-3:
            bad.append((k, d, floor))
    report.sort(reverse=True)
    print("fp16 engine vs fp16 emulator %s %d: worst grad rel L2 %.2e (floor %.2e, %s); median %.2e (floor %.2e); "
          "max|dp| %.2e (floor %.2e), dloss %.2e" % (gname, hw, report[0][0], report[0][1], report[0][2],
                                                    report[len(report) // 2][0], report[len(report) // 2][1],
                                                    np.abs(pg - pe).max(), pfloor, abs(lo_g[0] - lo_e[0])))
    assert not bad, bad
    assert np.abs(pg - pe).max() < 3.0 * pfloor + 1e-3
    assert abs(lo_g[0] - lo_e[0]) < 3e-4 and abs(lo_g[1] - lo_e[1]) < 3e-4


def test_unet_512_batch8_fp16_train_step_vs_oracle():
    """The benchmarked configuration itself (BASELINE configs[1]: U-Net 512x512x1, batch 8, fp16 storage): one
    training step against the fp32 oracle -- loss and Dice within 1e-3, BatchNorm moving statistics within 2e-3,
    training-mode probabilities within 3e-2 (batch statistics: see test_train_step_vs_oracle)."""
    gname, hw, n = "unet", 512, 8
    S = importlib.import_module(PKG + ".synthetic")
    params = perturbed_params(gname, hw)
    x, t = S.make_slices(n, hw, seed=1234)
    eng = engine_for(gname, hw, "float16", params, use_graph=True)
    eng._set_fields(step=5)
    b = eng.train_batch(dev(x).view(n, hw, hw, 1), dev(t), None, n)
    eng.stream.synchronize()
    lo = eng.loss_dev(b).cpu().numpy()
    probs = eng.probs(b).cpu().numpy().reshape(n, hw, hw, 1)
    new = eng.get_weights()
    assert not eng.overflowed()
    eng.close()
    r = K.loss_and_grads(gname, params, x, t, dtype=torch.float32, dropout=dict(seed=7, step=5), loss="bce_dice")
    print("unet 512 b8 fp16: loss %.6f vs %.6f, dice %.6f vs %.6f, max|dp| %.2e"
          % (lo[0], r["loss"], lo[1], r["metric"], np.abs(probs - r["probs"]).max()))
    assert lo[0] == pytest.approx(r["loss"], abs=1e-3)
    assert lo[1] == pytest.approx(r["metric"], abs=1e-3)
    assert np.abs(probs - r["probs"]).max() < 3e-2
    for k, v in r["new_moving"].items():
        assert np.abs(new[k] - v).max() < 2e-3 * max(1.0, float(np.abs(v).max())), k


@pytest.mark.parametrize("precision", ["float32", "float16"])
@pytest.mark.parametrize("gname,hw,n", [("unet", 224, 2), ("unet", 512, 1), ("unetpp", 224, 1), ("classifier", 224, 4)])
def test_inference_forward_vs_oracle(gname, hw, n, precision):
    """model.predict parity at the reference's real size (224) and BASELINE's 512."""
    params = perturbed_params(gname, hw)
    S = importlib.import_module(PKG + ".synthetic")
    x, _ = S.make_slices(n, hw, seed=3)
    want, _ = K.forward(gname, params, x, training=False, dtype=torch.float32)
    eng = engine_for(gname, hw, precision, params)
    b = eng.forward_batch(dev(x).view(n, hw, hw, 1), None, n)
    eng.stream.synchronize()
    got = eng.probs(b).cpu().numpy().reshape(want.shape)
    err = np.abs(got - want).max()
    print("%s %d %s: max|dp| = %.3e" % (gname, hw, precision, err))
    assert err < PTOL[precision]
    eng.close()


FILES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


@pytest.mark.parametrize("precision", ["float32", "float16"])
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_against_golden_vectors(path, precision):
    with np.load(path) as z:
        gold = {k.replace("__", "/"): z[k] for k in z.files}
    gname, hw, n = os.path.basename(path)[:-4].split("_")
    hw, n = int(hw), int(n[1:])
    params = perturbed_params(gname, hw, seed=11)
    assert np.allclose([float(np.asarray(v, np.float64).sum()) for v in params.values()], gold["weight_checksum"])
    x, t = gold["x"], gold["t"]
    eng = engine_for(gname, hw, precision, params)
    b = eng.forward_batch(dev(x).view(n, hw, hw, 1), None, n)
    eng.stream.synchronize()
    assert np.abs(eng.probs(b).cpu().numpy().reshape(gold["probs_infer"].shape) - gold["probs_infer"]).max() < PTOL[precision] * 3
    eng._set_fields(step=2)
    b = eng.train_batch(dev(x).view(n, hw, hw, 1), dev(t), None, n, sw_src=torch.ones(n, device="cuda"))
    eng.stream.synchronize()
    assert np.abs(eng.probs(b).cpu().numpy().reshape(gold["probs_train"].shape) - gold["probs_train"]).max() < (6e-5 if precision == "float32" else 1e-2)
    assert eng.loss_dev(b).cpu().numpy()[0] == pytest.approx(float(gold["loss"]), abs=20 * PTOL[precision])
    ls = eng._cur_ls
    for k, g in eng.get_grads().items():
        if precision == "float16" or ("conv2d_transpose" in k and k.endswith("bias")):
            continue
        g = g / ls
        scale = float(gold["gabs/" + k]) / g.size + 1e-12
        assert np.abs(g.reshape(-1)[:8] - gold["ghead/" + k]).max() < GTOL[precision] * 40 * scale + 1e-7, k
        assert abs(float(g.astype(np.float64).sum()) - float(gold["gsum/" + k])) < GTOL[precision] * float(gold["gabs/" + k]) + 1e-7, k
    new = eng.get_weights()
    for k in gold:
        if k.startswith("moving/"):
            assert np.abs(new[k[7:]] - gold[k]).max() < (1e-5 if precision == "float32" else 2e-3), k
    eng.close()


@pytest.mark.parametrize("precision", ["float32", "float16"])
def test_graph_replay_matches_eager_training(precision):
    hw, n = 32, 4
    params = perturbed_params("unet", hw)
    x, t = synth_batch(3 * n, hw)
    outs = []
    for use_graph in (False, True):
        eng = engine_for("unet", hw, precision, params, use_graph=use_graph)
        xd, td = dev(x).view(3 * n, hw, hw, 1), dev(t)
        losses = []
        for s in range(3):
            idx = torch.arange(s * n, (s + 1) * n, dtype=torch.int32, device="cuda")
            b = eng.train_batch(xd, td, idx, n)
            eng.stream.synchronize()
            losses.append(eng.loss_dev(b).cpu().numpy().copy())
        outs.append((np.array(losses), eng.get_weights()))
        eng.close()
    # atomics make the reduction order run-dependent: compare with a tolerance, not bit-for-bit
    assert np.allclose(outs[0][0], outs[1][0], rtol=1e-2 if precision == "float16" else 1e-4)
    for k in outs[0][1]:
        # Adam turns noise-level gradients (atomics -> run-dependent rounding) into +-lr sized updates
        assert np.abs(outs[0][1][k] - outs[1][1][k]).max() < 2 * 3 * 5e-4 + 1e-6, k


def test_partial_batch_and_odd_sizes():
    """last batch of an epoch is partial (1129 mod 32 = 9, SURVEY hard part 5); W != H."""
    hw = 32
    params = perturbed_params("unet", hw)
    eng = engine_for("unet", hw, "float32", params)
    for n in (1, 3):
        x, t = synth_batch(n, hw, seed=n)
        b = eng.forward_batch(dev(x).view(n, hw, hw, 1), None, n)
        eng.stream.synchronize()
        want, _ = K.forward("unet", params, x, training=False, dtype=torch.float32)
        assert np.abs(eng.probs(b).cpu().numpy() - want).max() < 2e-5
    eng.close()


def test_model_facade_fit_matches_oracle_training():
    """BASELINE config 1 in miniature: 16 synthetic slices, train_test_split 70/30 seed 42 (T1H:762),
    one epoch, batch 8 -> steps of 8 and 3; Dice parity with the oracle trained on the same batches."""
    from sklearn.model_selection import train_test_split
    M = importlib.import_module(PKG + ".model")
    LS = importlib.import_module(PKG + ".losses")
    S = importlib.import_module(PKG + ".synthetic")
    hw = 64
    x, t = S.make_slices(16, hw, seed=1234)
    xtr, xva, ttr, tva = train_test_split(x, t, test_size=0.3, random_state=42)
    assert len(xtr) == 11 and len(xva) == 5
    params, _ = K.init_params("unet", (hw, hw, 1), seed=42)
    m = M.Model(graph=G.unet(hw, 1), precision="float32", dropout_seed=7)
    m.set_weights_dict(params)
    m.compile(optimizer=M.Adam(lr=0.0005), loss=LS.bce_dice_loss, metrics=[LS.dice_coeff])
    h = m.fit(xtr, ttr, batch_size=8, epochs=2, validation_data=(xva, tva), shuffle=False, verbose=0)
    # oracle: same batches, same dropout stream, fp32
    opt = K.Adam(lr=0.0005)
    p = {k: v.copy() for k, v in params.items()}
    step, logs = 0, []
    for ep in range(2):
        ls = []
        for lo in (0, 8):
            l, d = K.train_step("unet", p, opt, xtr[lo:lo + 8], ttr[lo:lo + 8], dtype=torch.float32,
                                dropout=dict(seed=7, step=step))
            ls.append((l, d, len(xtr[lo:lo + 8])))
            step += 1
        w = np.array([a[2] for a in ls], float)
        logs.append((float((np.array([a[0] for a in ls]) * w).sum() / w.sum()),
                     float((np.array([a[1] for a in ls]) * w).sum() / w.sum())))
    pv, _ = K.forward("unet", p, xva, training=False, dtype=torch.float32)
    val_dice = float(K.dice_coeff(torch.from_numpy(tva).double(), torch.from_numpy(pv).double()))
    assert h.history["loss"][0] == pytest.approx(logs[0][0], rel=1e-3)
    assert h.history["loss"][1] == pytest.approx(logs[1][0], rel=2e-2)
    assert h.history["dice_coeff"][1] == pytest.approx(logs[1][1], abs=5e-3)       # +-0.5 pt
    assert h.history["val_dice_coeff"][1] == pytest.approx(val_dice, abs=5e-3)
    # evaluate / predict / threshold sweep surface
    ev = m.evaluate(xva, tva, batch_size=8)
    assert ev[1] == pytest.approx(h.history["val_dice_coeff"][1], rel=1e-5)
    pr = m.predict(xva)
    assert pr.shape == (5, hw, hw, 1) and np.abs(pr - pv).max() < 2e-2        # after 4 chaotic training steps
    sw = m.threshold_sweep(xva, tva, [0.3, 0.5, 0.7], batch_size=8)
    for k, th in enumerate([0.3, 0.5, 0.7]):
        want = K.sm_threshold_metrics(tva, pr, th)
        assert sw["f1"][k] == pytest.approx(want["f1"], rel=1e-4) and sw["iou"][k] == pytest.approx(want["iou"], rel=1e-4)


def test_train_on_batch_pipelined_matches_blocking():
    """Model.train_on_batch(wait=False): double-buffered staging on a copy stream; every step must train on ITS batch
    and return ITS loss (lr = 0 so that the steps are independent and comparable)."""
    M = importlib.import_module(PKG + ".model")
    LS = importlib.import_module(PKG + ".losses")
    hw, n = 32, 2
    batches = [synth_batch(n, hw, seg=True) for _ in range(3)]
    batches = [(x + 0.05 * k, t) for k, (x, t) in enumerate(batches)]
    seqs = []
    for pipelined in (False, True):
        m = M.Model(graph=G.unet(hw, 1), precision="float32", seed=42, use_graph=True)
        m.compile(optimizer=M.Adam(lr=0.0005), loss=LS.bce_dice_loss, metrics=[LS.dice_coeff])
        m.engine._set_fields(lr=0.0)
        out, pending = [], None
        for s in range(7):
            x, t = batches[s % 3]
            xt = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).pin_memory()
            tt = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32)).pin_memory()
            if pipelined:
                h = m.train_on_batch(xt, tt, wait=False, dropout=False)
                if pending is not None:
                    out.append(pending.get())
                pending = h
            else:
                out.append(m.train_on_batch(xt, tt, dropout=False))
        if pending is not None:
            out.append(pending.get())
        seqs.append(np.asarray(out))
        m.engine.close()
    assert seqs[0].shape == seqs[1].shape == (7, 2)
    # BN moving statistics do not enter training-mode losses, so step s only depends on batch s % 3
    assert np.allclose(seqs[0], seqs[1], rtol=1e-4, atol=1e-6), (seqs[0], seqs[1])
    assert np.allclose(seqs[0][0], seqs[0][3], rtol=1e-4) and not np.allclose(seqs[0][0], seqs[0][1], rtol=1e-3)


@pytest.mark.parametrize("flags", [dict(fuse_bn_bwd=False), dict(fuse_bn_bwd_wgrad=False), dict(fuse_bn_bwd=False, fuse_bn_bwd_wgrad=False),
                                   dict(fuse_bias_grad=False), dict(fuse_bn_pool=False), dict(fuse_bn_stats=False),
                                   dict(fuse_bn_bwd=False, fuse_bn_stats=False, fuse_bias_grad=False, fuse_bn_pool=False,
                                        fuse_bn_bwd_wgrad=False),
                                   dict(split_concat=0), dict(split_concat=2), dict(fuse_dropout_bn=False),
                                   dict(split_concat=2, fuse_bn_stats=False, fuse_bn_bwd_wgrad=False, fuse_bn_bwd=False)],
                         ids=lambda d: "+".join(sorted(d)))
def test_every_planner_fusion_can_be_switched_off_on_the_gpu(flags):
    """the schedule without a fusion computes the same exact-mode training step (the CPU twin of this test interprets the
    plans with the emulator; here the unfused KERNELS -- bn_bwd_reduce, bn_stats, channel_sum, separate max-pool -- run)"""
    gname, hw, n = "unet", 48, 3                    # (a TRAIN_CASES shape: its exact-mode step has no borderline ReLU flip)
    params = perturbed_params(gname, hw)
    x, t = synth_batch(n, hw, seg=True)
    r = K.loss_and_grads(gname, params, x, t, dtype=torch.float64, dropout=dict(seed=7, step=5), loss="bce_dice")
    eng = engine_for(gname, hw, "float32", params, plan_options=flags)
    eng._set_fields(step=5)
    b = eng.train_batch(dev(x).view(n, hw, hw, 1), dev(t), None, n)
    eng.stream.synchronize()
    lo = eng.loss_dev(b).cpu().numpy()
    grads = eng.get_grads()
    eng.close()
    assert lo[0] == pytest.approx(r["loss"], abs=20 * PTOL["float32"])
    worst, who = grad_errors(grads, r["grads"], norm="max")
    assert worst < GTOL["float32"], (flags, who, worst)


@pytest.mark.parametrize("precision", ["float32", "float16"])
def test_zero_and_tiny_batchnorm_gamma_switch_to_the_unfused_backward(precision):
    """the fused BatchNorm-backward statistics divide by gamma.  With gamma == 0 / 1e-4 in some channels the
    engine must notice (set_weights), switch to the reductions that do not, and still match the oracle -- d(gamma) of
    the zeroed channels included"""
    gname, hw, n = "unet", 48, 3
    params = perturbed_params(gname, hw)
    for name, idx, val in (("batch_normalization_1/gamma", 3, 0.0), ("batch_normalization_6/gamma", 10, 1e-4),
                           ("batch_normalization_8/gamma", 0, 0.0)):
        params[name] = params[name].copy()
        params[name][idx] = val
    x, t = synth_batch(n, hw, seg=True)
    eng = E.Engine(G.unet(hw, 1), precision=precision, use_graph=False, dropout_seed=7)
    with pytest.warns(UserWarning, match="BatchNorm gamma"):
        eng.set_weights(params)
    assert eng.plan_options["fuse_bn_bwd"] is False and eng.plan_options["fuse_bn_bwd_wgrad"] is False
    eng._set_fields(step=5)
    b = eng.train_batch(dev(x).view(n, hw, hw, 1), dev(t), None, n)
    eng.stream.synchronize()
    assert not any(o.kind in (P.OP_BN_BWD_SUMS_WGRAD,) for o in b.plan.bwd) and any(o.kind == P.OP_BN_BWD_REDUCE for o in b.plan.bwd)
    r = K.loss_and_grads(gname, params, x, t, dtype=torch.float64, dropout=dict(seed=7, step=5), loss="bce_dice")
    grads = {k: v / eng._cur_ls for k, v in eng.get_grads().items()}
    eng.close()
    if precision == "float32":
        errs = sorted(((float(np.abs(grads[k] - v).max() / (np.abs(v).max() + 1e-12)), k) for k, v in r["grads"].items()
                       if not ("conv2d_transpose" in k and k.endswith("bias"))), reverse=True)
        print("small-gamma exact mode: largest gradient errors", errs[:6])
        # a zero gamma makes a whole BatchNorm output channel constant, and the step becomes sensitive to rounding at the
        # 1 % level: the ORACLE ITSELF differs by 0.9 % on conv2d_11/kernel between fp32 and fp64 for these weights
        # (0.5 % without the zeros), and tests/emulator.py interpreting this very plan in fp32 shows the same 2.7 %
        # against the fp64 oracle as the GPU.  So: a loose bound on everything, a tight one on what the switch is about
        # -- d(gamma) of the zeroed channels, which the fused statistics would have returned as exactly 0.
        assert errs[0][0] < 6e-2, errs[:6]
    for name, idx in (("batch_normalization_1/gamma", 3), ("batch_normalization_8/gamma", 0)):
        want = r["grads"][name][idx]
        assert abs(grads[name][idx] - want) <= (3e-2 if precision == "float32" else 0.35) * abs(want) + 1e-7, (name, grads[name][idx], want)


@pytest.mark.parametrize("option,value", [("tc_dwmerge", 1), ("tc_dwmerge", 0), ("tc_halo", 0), ("tc_rowstrip", 1), ("tc_rowstrip", 0)])
def test_train_step_kernel_selection_options(option, value):
    """b2u_set_option kernel-selection switches (dw-merged thin-layer kernel for every eligible layer / for none,
    per-tap loads instead of halo tiles) compute the same fp16 training step as the default selection: loss, Dice,
    probabilities and gradients against the fp64 oracle with the default path's tolerances."""
    lib = importlib.import_module(PKG + "._lib").lib()
    gname, hw, n = "unet", 64, 4
    params = perturbed_params(gname, hw)
    x, t = synth_batch(n, hw, seg=True)
    r = K.loss_and_grads(gname, params, x, t, dtype=torch.float64, dropout=dict(seed=7, step=5), loss="bce_dice")
    old = lib.b2u_set_option(option.encode(), value)
    try:
        eng = engine_for(gname, hw, "float16", params)
        eng._set_fields(step=5)
        b = eng.train_batch(dev(x).view(n, hw, hw, 1), dev(t), None, n)
        eng.stream.synchronize()
        lo = eng.loss_dev(b).cpu().numpy()
        probs = eng.probs(b).cpu().numpy().reshape(r["probs"].shape)
        grads = {k: v / eng._cur_ls for k, v in eng.get_grads().items()}
        eng.close()
    finally:
        lib.b2u_set_option(option.encode(), old)
    assert np.abs(probs - r["probs"]).max() < 3e-2
    assert lo[0] == pytest.approx(r["loss"], abs=20 * PTOL["float16"])
    for k, g in r["grads"].items():
        if "conv2d_transpose" in k and k.endswith("bias"):
            continue
        a, bb = grads[k].ravel().astype(np.float64), g.ravel()
        assert np.linalg.norm(a - bb) / (np.linalg.norm(bb) + 1e-30) < 0.5, k


# ---- optional executor features (off by default) come after every default-path test: under `-x` a failure here
# ---- can no longer hide the headline-path tests above
@pytest.mark.parametrize("use_graph", [False, True])
def test_side_stream_weight_gradients_match_single_stream(use_graph):
    """executor option side_stream (off by default): weight-gradient ops on a forked stream (eager and captured) give
    the same gradients, loss and updated weights as the single-stream schedule.

    Two kinds of run-to-run noise exist even in exact (fp32) mode and were measured on B200 in round 2: (i) the
    order of the fp32 / fp64 atomics, ~3e-6 relative on the worst gradient tensor, every run; (ii) rarely, that noise
    flips a ReLU / max-pool decision of an element sitting on the threshold, which moves the small tensors behind it
    (biases, BN affine gradients) by 1e-4 .. 2e-3 in ONE run (this is what failed the fixed 1e-3 bound of round 1:
    2.2e-3 on one tensor, with 3 single-stream runs agreeing to 2.8e-6).  So each schedule runs three times and the
    CLOSEST single-stream / side-stream pair is compared against the closest single-stream pair: a flip in one run
    cannot fail the test, a schedule that really computes something else (a missing stream dependency) still does."""
    lib = importlib.import_module(PKG + "._lib").lib()
    gname, hw, n = "unet", 64, 4
    params = perturbed_params(gname, hw)
    x, t = synth_batch(n, hw, seg=True)
    outs = []
    for side in (0, 1, 0, 1, 0, 1):
        assert lib.b2u_set_option(b"side_stream", side) >= 0
        try:
            eng = engine_for(gname, hw, "float32", params, use_graph=use_graph)
            # lr = 0: Adam's first steps move every weight by +-lr whatever the gradient's magnitude, so the
            # summation-order noise of near-zero gradients would otherwise show up in the second step's loss
            eng._set_fields(step=5, lr=0.0)
            for _ in range(2):                        # second step replays the captured graph
                b = eng.train_batch(dev(x).view(n, hw, hw, 1), dev(t), None, n)
            eng.stream.synchronize()
            outs.append((side, eng.loss_dev(b).cpu().numpy().copy(), {k: v.copy() for k, v in eng.get_grads().items()},
                         eng.get_weights()))
            eng.close()
        finally:
            lib.b2u_set_option(b"side_stream", 0)
    rel = lambda a, b_: float(np.linalg.norm(a.astype(np.float64) - b_) / (np.linalg.norm(a.astype(np.float64)) + 1e-12))
    single, forked = [o for o in outs if o[0] == 0], [o for o in outs if o[0] == 1]
    for o in forked:
        assert np.allclose(single[0][1], o[1], rtol=2e-4, atol=1e-5)
    report = []
    for k in single[0][2]:
        floor = min(rel(a[2][k], b_[2][k]) for i, a in enumerate(single) for b_ in single[i + 1:])
        d = min(rel(a[2][k], b_[2][k]) for a in single for b_ in forked)
        report.append((d, floor, k))
    report.sort(reverse=True)
    print("side stream: worst relative L2 %.2e (noise floor of that tensor %.2e, %s)" % report[0])
    bad = [r for r in report if r[0] > 4.0 * r[1] + 2e-5]
    assert not bad, bad[:12]
    for k in single[0][3]:
        assert min(np.abs(single[0][3][k] - o[3][k]).max() for o in forked) < 1e-5, k
