"""GPU probe (not a test): per-op device times of one training step under different library options.

  python tools/ab_ops.py --opt wgrad_halo=1,2,3 --kinds conv3x3_wgrad [--workload unet512]

Prints one row per op of the selected kinds with a column per option value (milliseconds, CUDA events between ops
via b2u_run_ops_timed, best of 3 passes), and the per-kind totals."""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "one-stop-for-covid-19-infection-and-lung-segmentation-plus-classification_b200"
WL = {"unet512": ("unet", 512, 8), "unet256": ("unet", 256, 32), "unetpp512": ("unetpp", 512, 4), "unet224": ("unet", 224, 32)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--opt", default="")
    ap.add_argument("--kinds", default="")
    ap.add_argument("--workload", default="unet512")
    ap.add_argument("--plan", default="", help="planner options, e.g. relu_bits=1,fuse_bn_pool=0")
    ap.add_argument("--set", default="", help="library options held fixed for the whole run, e.g. tc_dw_epi8=0,pdl=0")
    args = ap.parse_args()
    name, vals = (args.opt.split("=") + [""])[:2] if args.opt else ("", "")
    vals = [int(v) for v in vals.split(",")] if vals else [None]
    kinds = set(k for k in args.kinds.split(",") if k)
    E = importlib.import_module(PKG + ".engine")
    G = importlib.import_module(PKG + ".graphs")
    P = importlib.import_module(PKG + ".plan")
    S = importlib.import_module(PKG + ".synthetic")
    graph, size, batch = WL[args.workload]
    torch.cuda.set_device(0)
    plan_options = {k: int(v) for k, v in (kv.split("=") for kv in args.plan.split(",") if kv)}
    eng = E.Engine(G.GRAPHS[graph](size, 1), precision="float16", use_graph=False, dropout_seed=7,
                   plan_options=plan_options)
    x, t = S.make_slices(batch, size, seed=1234)
    xd = torch.from_numpy(x).cuda()
    td = torch.from_numpy(t.reshape(batch, -1)).cuda()
    for kv in (kv for kv in args.set.split(",") if kv):
        eng.lib.b2u_set_option(kv.split("=")[0].encode(), int(kv.split("=")[1]))
    cols = []
    for v in vals:
        if name:
            eng.lib.b2u_set_option(name.encode(), v)
        eng.train_batch(xd, td, None, batch)
        eng.stream.synchronize()
        best = None
        for _ in range(3):
            prof = eng.profile_train_ops(batch)
            ms = [m for _, m in prof]
            best = ms if best is None else [min(a, b) for a, b in zip(best, ms)]
        cols.append((prof, best))
    ops = [op for op, _ in cols[0][0]]
    print("%-16s %-22s %s" % ("op", "layer", " ".join("%10s" % ("%s=%s" % (name, v) if name else "ms") for v in vals)))
    tot = {}
    for k, op in enumerate(ops):
        nm = P.OP_NAMES[op.kind][3:].lower()
        row = [c[1][k] for c in cols]
        d = tot.setdefault(nm, [0.0] * len(vals))
        for j, r in enumerate(row):
            d[j] += r
        if not kinds or nm in kinds:
            print("%-16s %-22s %s  %s" % (nm, op.tag, " ".join("%10.4f" % r for r in row), op.i[:8]))
    print("---- totals per kind (ms)")
    for nm, d in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        print("%-16s %s" % (nm, " ".join("%10.4f" % r for r in d)))
    print("%-16s %s" % ("step", " ".join("%10.4f" % sum(d[j] for d in tot.values()) for j in range(len(vals)))))


if __name__ == "__main__":
    main()
