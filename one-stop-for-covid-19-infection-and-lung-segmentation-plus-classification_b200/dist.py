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
