"""2-rank smoke test of the NCCL data-parallel path with progress prints (run under torchrun, short timeout)."""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "one-stop-for-covid-19-infection-and-lung-segmentation-plus-classification_b200"


def say(*a):
    print("[rank %s %.1fs]" % (os.environ.get("RANK"), time.time() - T0), *a, flush=True)


T0 = time.time()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
import torch.distributed as dist
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
say("pg up")
E = importlib.import_module(PKG + ".engine")
G = importlib.import_module(PKG + ".graphs")
S = importlib.import_module(PKG + ".synthetic")
comm = E.Comm(rank, world)
say("comm up")
for use_graph in (False, True):
    eng = E.Engine(G.unet(64, 1), precision="float16", comm=comm, use_graph=use_graph, seed=42)
    say("engine up graph=%s" % use_graph)
    x, t = S.make_slices(4, 64, seed=10 + rank)
    xd, td = torch.from_numpy(x).cuda(), torch.from_numpy(t.reshape(4, -1)).cuda()
    for s in range(3):
        b = eng.train_batch(xd, td, None, 4)
        eng.stream.synchronize()
        say("step", s, eng.loss_dev(b).cpu().numpy())
    w = eng.get_weights()
    chk = torch.tensor([float(sum(float(np.abs(v).sum()) for v in w.values()))], device="cuda", dtype=torch.float64)
    lst = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    say("weight checksums identical across ranks:", all(abs(float(a) - float(lst[0])) < 1e-6 * abs(float(lst[0])) for a in lst))
    eng.close()
comm.close()
dist.destroy_process_group()
say("done")
