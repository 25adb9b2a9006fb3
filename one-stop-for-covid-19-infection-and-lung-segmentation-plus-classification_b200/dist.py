"""Host-side helpers for the data-parallel path (one process per GPU, launched by torchrun).

The reference is single-device; images are independent, so the global batch / dataset is
sharded by rank and the only exchange is the gradient all-reduce enqueued by the op list
(plan.OP_ALLREDUCE_F32) plus, with sync_stats, the fp64 BatchNorm / Dice sums.
"""
import os


def shard_range(n, rank, world):
    """[lo, hi) of rank's contiguous shard of n samples (remainder spread over the first ranks)."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def epoch_batches(perm, batch_size, rank=0, world=1):
    """This rank's mini-batches (index arrays) of one epoch over the permutation `perm` of the whole dataset.

    Every rank must issue the SAME number of steps with the SAME batch sizes (the gradient all-reduce is part of the
    captured step, so a missing step on one rank deadlocks the others): the permutation -- identical on every rank,
    they share the shuffle seed -- is padded by wrapping around to a multiple of the world size and dealt out
    round-robin (rank r takes perm[r::world]), then cut into batches of `batch_size` per rank; the last batch of the
    epoch is partial on all ranks alike.  world = 1 is the reference's single-device epoch (T1H:1059-1061)."""
    import numpy as np
    perm = np.asarray(perm)
    if world > 1:
        if not 0 <= rank < world:
            raise ValueError("rank %d outside world of %d" % (rank, world))
        pad = (-len(perm)) % world
        if pad:
            perm = np.concatenate([perm, perm[:pad]])
        perm = perm[rank::world]
    return [perm[lo:lo + batch_size] for lo in range(0, len(perm), batch_size)]


def mean_over_ranks(value, world):
    """mean of a host scalar over the ranks of the default torch.distributed group (training logs of model.fit)"""
    import torch
    import torch.distributed as dist
    if world == 1 or not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t)
    return float(t.item()) / world


def barrier(world):
    import torch.distributed as dist
    if world > 1 and dist.is_initialized():
        dist.barrier()


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_from_env(backend="nccl"):
    """torch.distributed init from torchrun's environment (127.0.0.1 rendezvous) + the engine's NCCL comm."""
    import torch
    import torch.distributed as dist
    from . import engine as E
    rank, world, local = env_rank_world()
    if world == 1:
        return None
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend, device_id=torch.device("cuda", local) if backend == "nccl" else None)
    return E.Comm(rank, world)
