"""The planner's schedule (fusions, zero-copy concat, gradient routing) interpreted on the CPU by
tests/emulator.py must reproduce the oracle's autograd for all three reference graphs."""
import numpy as np
import pytest
import torch

import emulator as E
from helpers import G, K, P, grad_errors, perturbed_params, synth_batch


def _load(em, plan, params, x, t):
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


@pytest.mark.parametrize("gname,hw,n,loss", [("unet", 32, 2, "bce_dice"), ("unetpp", 16, 2, "bce_dice"),
                                             ("classifier", 32, 4, "bce")])
def test_train_step_matches_oracle(gname, hw, n, loss):
    g = G.GRAPHS[gname](hw, 1)
    plan = P.Plan(g, n, dt=P.F32, training=True, dropout=True, loss=loss)
    params = perturbed_params(gname, hw)
    x, t = synth_batch(n, hw, seg=(loss == "bce_dice"))
    em = E.Emulator(plan.arena_sizes())
    em.state.update(seed=7, step=5)
    fp, fs = _load(em, plan, params, x, t)
    em.run(plan.train_ops())
    r = K.loss_and_grads(gname, params, x, t, dtype=torch.float64, dropout=dict(seed=7, step=5), loss=loss)
    lo = em.f32(plan.loss_out, 2)
    assert lo[0] == pytest.approx(r["loss"], rel=1e-5)
    if loss == "bce_dice":
        assert lo[1] == pytest.approx(r["metric"], rel=1e-5)
    probs = em.f32(plan.prob, t.size).reshape(t.shape)
    assert np.abs(probs - r["probs"]).max() < 2e-5
    grads = plan.layout.unpack(em.f32(P.Ref("grads", 0), fp.size), None)
    worst, who = grad_errors(grads, r["grads"])
    assert worst < 1e-3, (who, worst)
    # Adam + BN moving statistics
    want = {k: v.copy() for k, v in params.items()}
    K.Adam().step(want, r["grads"])
    want.update(r["new_moving"])
    new = plan.layout.unpack(em.f32(P.Ref("params", 0), fp.size), em.f32(P.Ref("state", 0), fs.size))
    for k in want:
        if "conv2d_transpose" in k and k.endswith("bias"):
            continue
        assert np.abs(new[k] - want[k]).max() < 2e-5, k
    assert em.state["step"] == 6


@pytest.mark.parametrize("gname,hw,n", [("unet", 32, 3), ("unetpp", 16, 2), ("classifier", 32, 5)])
def test_inference_forward_matches_oracle(gname, hw, n):
    g = G.GRAPHS[gname](hw, 1)
    loss = "bce" if gname == "classifier" else "bce_dice"
    plan = P.Plan(g, n, dt=P.F32, training=False, loss=loss)
    params = perturbed_params(gname, hw)
    x, t = synth_batch(n, hw, seg=(loss == "bce_dice"))
    em = E.Emulator(plan.arena_sizes())
    _load(em, plan, params, x, t)
    em.run(plan.forward_ops(with_loss=True))
    want, _ = K.forward(gname, params, x, training=False, dtype=torch.float64)
    probs = em.f32(plan.prob, t.size).reshape(t.shape)
    assert np.abs(probs - want).max() < 2e-5
    assert not plan.bwd and not plan.opt


def test_unet_concat_is_zero_copy_and_fused():
    plan = P.Plan(G.unet(32, 1), 2, dt=P.F16, training=True)
    kinds = [o.kind for o in plan.train_ops()]
    assert P.OP_COPY_SLICE not in kinds                  # skips + upsampled halves are written in place
    assert P.OP_DROPOUT_FWD not in kinds                 # dropout folded into the max-pool pass
    assert kinds.count(P.OP_BN_STATS) == 0               # concat BN statistics come from the producers' epilogues
    # BN backward statistics: encoder ones ride on the max-pool backward, decoder ones follow from the weight
    # gradient of the conv that reads the BN output (adjoint identity) -- no reduction pass is left
    assert kinds.count(P.OP_BN_BWD_REDUCE) == 0 and kinds.count(P.OP_BN_BWD_SUMS_WGRAD) == 4
    legacy = [o.kind for o in P.Plan(G.unet(32, 1), 2, dt=P.F16, training=True, fuse_bn_bwd_wgrad=False, split_concat=False).train_ops()]
    assert legacy.count(P.OP_BN_BWD_REDUCE) == 4
    halves = [o.kind for o in P.Plan(G.unet(32, 1), 2, dt=P.F16, training=True, fuse_bn_bwd_wgrad=False, split_concat=2).train_ops()]
    assert halves.count(P.OP_BN_BWD_REDUCE) == 8          # split concats: one reduction per dense half
    assert kinds.count(P.OP_CONV3X3_FWD) == 18 and kinds.count(P.OP_CONVT_FWD) == 4


def test_unetpp_home_is_last_concat():
    plan = P.Plan(G.unetpp(16, 1), 1, dt=P.F32, training=True)
    # c1 feeds three concats (2 copies), conv1_2 two (1 copy), c2 two (1 copy); everything else is in place
    assert sum(1 for o in plan.fwd if o.kind == P.OP_COPY_SLICE) == 2 + 1 + 1
    assert sum(1 for o in plan.bwd if o.kind == P.OP_COPY_SLICE) == 2 + 1 + 1
    assert all(o.i[4] == 1 for o in plan.bwd if o.kind == P.OP_COPY_SLICE)


def test_param_layout_roundtrip_and_counts():
    for gname, want in (("unet", 7762401), ("unetpp", 2207329), ("classifier", 1677937)):
        g = G.GRAPHS[gname](224, 1)
        lay = P.ParamLayout(g)
        tr = sum(int(np.prod(s)) for _, s, _, t in g.weight_specs() if t)
        assert tr == want
        assert lay.n_params >= tr and lay.n_params - tr < 4 * len(lay.specs)
    params = perturbed_params("unet", 32)
    lay = P.ParamLayout(G.unet(32, 1))
    fp, fs = lay.pack(params)
    back = lay.unpack(fp, fs)
    assert all(np.array_equal(back[k], params[k]) for k in params)
    with pytest.raises(ValueError):
        bad = dict(params); bad["conv2d_1/kernel"] = np.zeros((3, 3, 2, 32), np.float32)
        lay.pack(bad)


def test_shape_validation():
    with pytest.raises(ValueError):
        G.unet(30, 1)          # 30 is not divisible by 16: the 2nd pool would see an odd size


@pytest.mark.parametrize("flags", [dict(fuse_bias_grad=False), dict(fuse_bn_bwd_wgrad=False), dict(fuse_bn_pool=False),
                                   dict(fuse_bn_bwd=False, fuse_bn_stats=False), dict(prepack=False), dict(relu_bits=True)])
def test_every_fusion_can_be_switched_off(flags):
    """each planner fusion (producer-side bias gradients, BN backward statistics from the weight gradient / the max-pool
    backward, BN statistics from producer epilogues, BN apply + pool, one-launch weight packing) is optional: the
    schedule without it computes the same training step"""
    gname, hw, n = "unet", 32, 2
    g = G.GRAPHS[gname](hw, 1)
    params = perturbed_params(gname, hw)
    x, t = synth_batch(n, hw, seg=True)
    r = K.loss_and_grads(gname, params, x, t, dtype=torch.float64, dropout=dict(seed=7, step=5), loss="bce_dice")
    for kw in (flags, {}):
        plan = P.Plan(g, n, dt=P.F32, training=True, dropout=True, loss="bce_dice", **kw)
        em = E.Emulator(plan.arena_sizes())
        em.state.update(seed=7, step=5)
        fp, fs = _load(em, plan, params, x, t)
        em.run(plan.train_ops())
        assert em.f32(plan.loss_out, 2)[0] == pytest.approx(r["loss"], rel=1e-5)
        grads = plan.layout.unpack(em.f32(P.Ref("grads", 0), fp.size), None)
        worst, who = grad_errors(grads, r["grads"])
        assert worst < 1e-3, (kw, who, worst)


def test_relu_bits_plan_structure():
    """relu_bits (default with fp16 storage, off in the exact fp32 mode): the nine `a` convs of the U-Net write a
    packed ReLU mask and the data gradients of the nine `b` convs read it"""
    on = P.Plan(G.unet(32, 1), 2, dt=P.F16, training=True)
    off = P.Plan(G.unet(32, 1), 2, dt=P.F16, training=True, relu_bits=False)
    assert not P.Plan(G.unet(32, 1), 2, dt=P.F32, training=True).relu_bits and not P.Plan(G.unet(32, 1), 2, dt=P.F16, training=False).relu_bits
    assert sum(1 for o in on.fwd if o.kind == P.OP_CONV3X3_FWD and len(o.p) > 6 and o.p[6] is not None) == 9
    assert sum(1 for o in on.bwd if o.kind == P.OP_CONV3X3_DGRAD and o.i[5] == P.ACT_RELU_BITS) == 9
    assert all(o.p[6] is None for o in off.fwd if o.kind == P.OP_CONV3X3_FWD)
    assert not any(o.kind == P.OP_CONV3X3_DGRAD and o.i[5] == P.ACT_RELU_BITS for o in off.bwd)
    assert on.act.size > off.act.size


@pytest.mark.parametrize("gname,hw", [("unet", 64), ("unetpp", 32), ("classifier", 32)])
def test_gradient_buckets_cover_the_buffer_and_follow_their_last_writer(gname, hw):
    """data parallel: the flat gradient buffer is all-reduced in contiguous buckets, each placed right behind the last
    backward op that writes into it (executor flag OPF_COMM = side stream), the union is the whole buffer exactly once,
    nothing but Adam reads a bucket after its exchange, and grad_bucket_bytes=0 restores the single all-reduce"""
    loss = "bce" if gname == "classifier" else "bce_dice"
    plan = P.Plan(G.GRAPHS[gname](hw, 1), 4, dt=P.F16, training=True, world=4, loss=loss, grad_bucket_bytes=256 << 10)
    ars = [(k, o) for k, o in enumerate(plan.bwd) if o.kind == P.OP_ALLREDUCE_F32]
    assert len(ars) >= 2 and all(o.dt & P.OPF_COMM for _, o in ars) and not any(o.kind == P.OP_ALLREDUCE_F32 for o in plan.opt)
    spans = sorted((o.p[0].off // 4, o.p[0].off // 4 + o.i[0]) for _, o in ars)
    assert spans[0][0] == 0 and spans[-1][1] == plan.layout.n_params
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))                      # contiguous, no overlap
    for k, o in ars:
        lo, hi = o.p[0].off, o.p[0].off + 4 * o.i[0]
        for later in plan.bwd[k + 1:]:          # nothing touches the bucket after its exchange: no gradient write, and
            for ref in later.p:                 # no read either (BN_BWD_SUMS_WGRAD reads a conv's LOCAL weight gradient)
                if isinstance(ref, P.Ref) and ref.arena == "grads" and later.kind != P.OP_ALLREDUCE_F32:
                    assert not (lo <= ref.off < hi), (o.tag, later)
    assert plan.opt[0].kind == P.OP_ADAM and plan.opt[0].dt & P.OPF_JOIN
    one = P.Plan(G.GRAPHS[gname](hw, 1), 4, dt=P.F16, training=True, world=4, loss=loss, grad_bucket_bytes=0)
    assert [o.kind for o in one.opt][:2] == [P.OP_ALLREDUCE_F32, P.OP_ADAM] and not any(o.kind == P.OP_ALLREDUCE_F32 for o in one.bwd)


def test_inference_plan_pads_a_three_channel_input_for_the_tensor_core_first_conv():
    """Task-2 slices are 224 x 224 x 3 (BASELINE configs[4]): the fp16 inference plan stores the input zero-padded to 16
    channels and runs conv2d_1 as a K = 16 tensor-core conv with a zero-padded packed kernel; the emulated forward equals
    the oracle's, training plans and exact (fp32) plans keep the 3-channel tensor"""
    hw, n, cin = 32, 3, 3
    g = G.classifier(hw, cin)
    params = perturbed_params("classifier", hw, cin=cin)
    rng = np.random.default_rng(0)
    x = rng.random((n, hw, hw, cin)).astype(np.float32)
    plan = P.Plan(g, n, dt=P.F16, training=False, loss="bce")
    assert plan.x_pad == 16 and plan.x_cin == 3 and plan.x_view.c == 16
    op = plan.fwd[0]
    assert op.kind == P.OP_CONV3X3_FWD and op.i[1] == 16 and op.i[9] == 3 and op.p[5] is not None
    assert plan.pack_table()[0].tolist()[3:] == [0, 9, 16, 16, 3]
    em = E.Emulator(plan.arena_sizes())
    _load(em, plan, params, x, np.zeros((n, 1), np.float32))
    em.run(plan.forward_ops())
    want, _ = K.forward("classifier", params, x, training=False, dtype=torch.float32)
    assert np.abs(em.f32(plan.prob, n).reshape(n, 1) - want).max() < 2e-3          # fp16 storage
    assert P.Plan(g, n, dt=P.F16, training=True, loss="bce").x_pad == 0
    assert P.Plan(g, n, dt=P.F32, training=False, loss="bce").x_pad == 0


def test_inference_plan_folds_batchnorm_into_the_conv_epilogue_and_hoists_weight_only_ops():
    """fp16 inference plans (T2:747-778 classifier; U-Net encoder): Conv2D(relu) -> BatchNormalization is one conv op with a
    post-activation affine, BN finalize and operand packing live in `prep_ops()` (run when the weights change, not per
    batch), the per-batch list holds no BN op for the folded layers; same probabilities as the unfused plan and the
    oracle; training plans and exact plans are untouched; `tap` plans (fuse_bn_infer=False) keep every layer output."""
    hw, n, cin = 32, 3, 3
    g = G.classifier(hw, cin)
    params = perturbed_params("classifier", hw, cin=cin)
    x = np.random.default_rng(1).random((n, hw, hw, cin)).astype(np.float32)
    want, _ = K.forward("classifier", params, x, training=False, dtype=torch.float32)
    plan = P.Plan(g, n, dt=P.F16, training=False, loss="bce")
    convs = [o for o in plan.fwd if o.kind == P.OP_CONV3X3_FWD]
    assert len(convs) == 6 and all(o.p[7] is not None and o.p[8] is not None for o in convs)
    assert plan.folded_into_next == {"conv2d_%d" % k for k in range(1, 7)}
    per_batch = plan.forward_ops(prep=False)
    assert not any(o.kind in (P.OP_BN_FINALIZE, P.OP_BN_APPLY, P.OP_BN_APPLY_POOL, P.OP_PACK_WEIGHTS) for o in per_batch)
    assert sum(1 for o in per_batch if o.kind == P.OP_MAXPOOL_FWD) == 3
    prep = plan.prep_ops()
    assert prep[0].kind == P.OP_PACK_WEIGHTS and [o.kind for o in prep[1:]] == [P.OP_BN_FINALIZE] * 6
    assert [repr(o) for o in plan.forward_ops()] == [repr(o) for o in prep + per_batch]
    outs = {}
    for name, pl in (("folded", plan), ("tap", P.Plan(g, n, dt=P.F16, training=False, loss="bce", fuse_bn_infer=False)),
                     ("inline", P.Plan(g, n, dt=P.F16, training=False, loss="bce", hoist_prep=False))):
        em = E.Emulator(pl.arena_sizes())
        _load(em, pl, params, x, np.zeros((n, 1), np.float32))
        em.run(pl.forward_ops())
        outs[name] = em.f32(pl.prob, n).reshape(n, 1).copy()
        assert np.abs(outs[name] - want).max() < 2e-3, name
    assert np.array_equal(outs["folded"], outs["inline"])
    tap = P.Plan(g, n, dt=P.F16, training=False, loss="bce", fuse_bn_infer=False)
    assert not tap.folded_into_next and sum(1 for o in tap.fwd if o.kind in (P.OP_BN_APPLY, P.OP_BN_APPLY_POOL)) == 6
    assert not P.Plan(g, n, dt=P.F32, training=False, loss="bce").folded_into_next
    tr = P.Plan(g, n, dt=P.F16, training=True, loss="bce")
    assert not tr.folded_into_next and tr.prep_ops() == [] and tr.train_ops()[0].kind == P.OP_PACK_WEIGHTS
    # U-Net: only the encoder's conv -> BN pairs fold (the decoder BNs normalise a concat buffer)
    un = P.Plan(G.unet(32, 1), 2, dt=P.F16, training=False)
    assert un.folded_into_next == {"conv2d_2", "conv2d_4", "conv2d_6", "conv2d_8"}
    xs = np.random.default_rng(2).random((2, 32, 32, 1)).astype(np.float32)
    pu = perturbed_params("unet", 32)
    wantu, _ = K.forward("unet", pu, xs, training=False, dtype=torch.float32)
    em = E.Emulator(un.arena_sizes())
    _load(em, un, pu, xs, np.zeros((2, 32 * 32), np.float32))
    em.run(un.forward_ops())
    assert np.abs(em.f32(un.prob, 2 * 32 * 32).reshape(wantu.shape) - wantu).max() < 2e-3


def test_unet_decoder_concats_stay_two_dense_tensors():
    """concatenate([Conv2DTranspose, skip]) -> BatchNormalization -> Conv2D (T1H:887-889 and the three levels below): the
    planner keeps the two inputs as dense tensors (no interleaved 2c-channel buffer) and the BN ops carry both sources /
    both gradient destinations; U-Net++ (concats read by convs, up to five inputs) keeps its buffers; the emulated step
    equals the one planned with split_concat=False bit for bit (same arithmetic, different addresses)."""
    hw, n = 32, 2
    g = G.unet(hw, 1)
    # default: only where an input is narrower than a 128-byte line (32 fp16 channels: the full-resolution level)
    assert P.Plan(g, n, dt=P.F16, training=True).split_concats == ["concatenate_4"]
    assert P.Plan(g, n, dt=P.F16, training=False).split_concats == ["concatenate_4"]
    assert not P.Plan(g, n, dt=P.F32, training=True).split_concats
    plan = P.Plan(g, n, dt=P.F16, training=True, split_concat=2)          # every eligible concat
    assert plan.split_concats == ["concatenate_%d" % k for k in range(1, 5)]
    assert not P.Plan(G.unetpp(hw, 1), n, dt=P.F16, training=True, split_concat=2).split_concats
    two_src = [o for o in plan.fwd if o.kind == P.OP_BN_APPLY and len(o.p) > 5 and o.p[5] is not None]
    two_dst = [o for o in plan.bwd if o.kind == P.OP_BN_BWD_APPLY and len(o.p) > 12 and o.p[12] is not None]
    assert len(two_src) == 4 and len(two_dst) == 4 and all(o.i[5] * 2 == o.i[2] for o in two_src)
    # every tensor that was half of a concat pixel is dense now: transposed-conv outputs and skip tensors have ld == c
    for o in plan.fwd:
        if o.kind == P.OP_CONVT_FWD:
            assert o.i[2] == o.i[3]
        if o.kind == P.OP_BN_APPLY_POOL:
            assert o.i[1] == o.i[2]
    params = perturbed_params("unet", hw)
    x, t = synth_batch(n, hw, seg=True)
    # ... also with the BN fusions off (one statistics / reduction pass per dense half instead of producer epilogues)
    for opts in ({}, dict(fuse_bn_stats=False, fuse_bn_bwd_wgrad=False, fuse_bn_bwd=False)):
        res = []
        for split in (2, 0):
            pl = P.Plan(g, n, dt=P.F16, training=True, split_concat=split, **opts)
            assert bool(pl.split_concats) == bool(split)
            em = E.Emulator(pl.arena_sizes())
            em.fp16_weights = True
            _load(em, pl, params, x, t.reshape(n, -1))
            em.run(pl.train_ops())
            res.append((em.f32(pl.loss_out, 2).copy(), em.mem["grads"].copy(), em.mem["params"].copy()))
        if opts:
            assert sum(1 for o in pl.fwd if o.kind == P.OP_BN_STATS) > 0
        assert np.array_equal(res[0][0], res[1][0]), opts
        tol = 0 if not opts else 1e-6          # two half reductions add their double sums in another order than one pass
        ga, gb = (np.frombuffer(r[1], np.float32, len(r[1]) // 4) for r in res)
        assert np.abs(ga - gb).max() <= tol * max(1.0, np.abs(gb).max()), opts


def test_unetpp_dropout_in_front_of_batchnorm_is_never_materialised():
    """Conv2D -> Dropout -> BatchNormalization (UPP:874-876 and every decoder block): the BN ops regenerate the keep mask and
    read the conv output; 12 of the 16 Dropout layers lose their forward and backward passes; the emulated step still
    reproduces the oracle's autograd with the SAME dropout stream (and equals the unfused plan in exact mode)."""
    hw, n = 16, 2
    g = G.unetpp(hw, 1)
    plan = P.Plan(g, n, dt=P.F32, training=True, dropout=True)
    off = P.Plan(g, n, dt=P.F32, training=True, dropout=True, fuse_dropout_bn=False)
    count = lambda pl, k: sum(1 for o in pl.train_ops() if o.kind == k)
    assert (count(off, P.OP_DROPOUT_FWD), count(off, P.OP_DROPOUT_BWD)) == (16, 16)
    assert (count(plan, P.OP_DROPOUT_FWD), count(plan, P.OP_DROPOUT_BWD)) == (4, 4) and len(plan._lazy_drop) == 12
    assert sum(1 for o in plan.fwd if o.kind == P.OP_BN_STATS and o.f and o.f[0] > 0 and o.p[3] is not None) == 12
    assert sum(1 for o in plan.fwd if o.kind == P.OP_BN_APPLY and o.f and o.f[0] > 0) == 12
    assert sum(1 for o in plan.bwd if o.kind == P.OP_BN_BWD_APPLY and o.f and o.f[0] > 0) == 12
    params = perturbed_params("unetpp", hw)
    x, t = synth_batch(n, hw, seg=True)
    r = K.loss_and_grads("unetpp", params, x, t, dtype=torch.float64, dropout=dict(seed=7, step=5), loss="bce_dice")
    res = []
    for pl in (plan, off):
        em = E.Emulator(pl.arena_sizes())
        em.state.update(seed=7, step=5)
        _load(em, pl, params, x, t.reshape(n, -1))
        em.run(pl.train_ops())
        assert em.f32(pl.loss_out, 1)[0] == pytest.approx(r["loss"], rel=2e-5)
        grads = pl.layout.unpack(em.f32(P.Ref("grads", 0), pl.layout.n_params), None)
        worst, who = grad_errors(grads, r["grads"])
        assert worst < 2e-3, (who, worst)
        res.append(em.f32(P.Ref("grads", 0), pl.layout.n_params).copy())
    assert np.abs(res[0] - res[1]).max() <= 2e-6 * max(1.0, np.abs(res[1]).max())
