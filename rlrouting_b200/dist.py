"""Replica sharding across the GPUs of one box + the single collective of the path.

Replicas are independent environments: rank g owns a contiguous block of the global
replica index space, a replica's random streams are keyed by its GLOBAL index (so results do not
depend on the number of ranks), and the only exchange is a final all-reduce (sum) of the integer
statistics — NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import os

import numpy as np


def shard(total, rank, world):
    """Contiguous block [base, base+count) of `total` replicas owned by `rank` of `world`."""
    q, rem = divmod(int(total), int(world))
    count = q + (1 if rank < rem else 0)
    base = rank * q + min(rank, rem)
    return base, count


def shard_loads(loads, replicas_per_load, rank, world):
    """Per-replica load vector of this rank for a (load level x replica) grid laid out level-major:
    global replica i belongs to level i // replicas_per_load."""
    loads = np.asarray(loads, dtype=np.float64)
    base, count = shard(len(loads) * replicas_per_load, rank, world)
    idx = (base + np.arange(count)) // replicas_per_load
    return base, loads[idx], idx


def init_from_env(backend=None):
    """torchrun rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).  Returns (rank, world, local_rank)."""
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def allreduce_sum_(t):
    """In-place all-reduce (sum) of an int64 / float64 statistics tensor; no-op for one rank."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_max_(t):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


def level_stats(counters, level_idx, n_levels):
    """Sums per-replica counters [R, C] into per-load-level rows [n_levels, C] (torch, on device)."""
    import torch
    out = torch.zeros((n_levels, counters.shape[1]), dtype=counters.dtype, device=counters.device)
    out.index_add_(0, torch.as_tensor(level_idx, device=counters.device, dtype=torch.long), counters)
    return out


def global_stats(network):
    """Whole-job sums of a Network's counters: rlr_reduce_stats on the device, then one all-reduce."""
    from . import _native as nat
    s = allreduce_sum_(network.stats_device().clone()).cpu().numpy()
    out = {k: int(s[i]) for k, i in nat.CTR.items()}
    out["active_packets"] = out["all_packets"] - out["end_packets"] - out["drop_packets"]
    out["ave_route_time"] = out["route_time"] / out["end_packets"] if out["end_packets"] else 0.0
    return out
