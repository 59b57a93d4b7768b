"""Multi-GPU plumbing: environments shard trivially (the reference's equivalent is N OS processes,
config/parallel.py:10-25).  One process per GPU; the only collective is an all-gather of fixed-size episode
statistics (torch.distributed: NCCL over NVLink on the GPU box, gloo in CPU tests).  No data-path collective."""
import os

import torch
import torch.distributed as dist

STAT_FIELDS = ("messages", "limit_orders", "cancels", "fills", "spread_queries", "error_envs", "envs")


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1)."""
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shard_range(n_envs_total, rank, world):
    """Contiguous environment index range [lo, hi) of `rank`; per-env seed = base + global index, so results are
    placement invariant."""
    base, rem = divmod(int(n_envs_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_seeds(base_seed, lo, hi):
    import numpy as np
    return (np.arange(lo, hi, dtype=np.uint64) + np.uint64(base_seed))


def summarize(stats):
    """abx_env_stats structured array -> fixed-size int64 summary vector (STAT_FIELDS order)."""
    from . import _lib
    return torch.tensor([int(stats["messages"].sum()), int(stats["limit_orders"].sum()), int(stats["cancels"].sum()),
                         int(stats["fills"].sum()), int(stats["spread_queries"].sum()),
                         int(((stats["flags"] & _lib.F_ERROR_MASK) != 0).sum()), len(stats)], dtype=torch.int64)


def gather_summaries(local_summary, device=None):
    """All-gather the per-rank summary vectors -> [world, len(STAT_FIELDS)] on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_summary.unsqueeze(0).clone()
    t = local_summary.to(device) if device is not None else local_summary
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu()


def max_over_ranks(value, device=None):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
