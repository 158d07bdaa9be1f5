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
    out.index_add_(0, torch.as_tensor(np.array(level_idx), device=counters.device, dtype=torch.long), counters)
    return out


def combine_stats(counters, level_index=None, n_levels=None, q_levels=None, totals=None):
    """The reduction behind `global_stats`, on whatever device the tensors live (NCCL on GPUs, gloo in the CPU tests):
    `counters` [R, 8] int64 of this rank's replicas -> whole-job sums; with `level_index` also per-level sums and mean
    delivery times; `q_levels` [n_levels, 2] fp64 (this rank's per-level {sum Q, sum |Q - Q_prev|}) is all-reduced too."""
    import torch
    from . import _native as nat
    s = allreduce_sum_(totals.clone() if totals is not None else counters.sum(0)).cpu().numpy()
    out = {k: int(s[i]) for k, i in nat.CTR.items()}
    out["active_packets"] = out["all_packets"] - out["end_packets"] - out["drop_packets"]
    out["ave_route_time"] = out["route_time"] / out["end_packets"] if out["end_packets"] else 0.0
    if level_index is None:
        return out
    lv = np.ascontiguousarray(np.broadcast_to(np.asarray(level_index), (counters.shape[0],)), dtype=np.int64)
    L = int(n_levels) if n_levels is not None else int(lv.max()) + 1
    per_level = allreduce_sum_(level_stats(counters, lv, L)).cpu().numpy()
    end = per_level[:, nat.CTR["end_packets"]]
    out["levels"] = {k: per_level[:, i].copy() for k, i in nat.CTR.items()}
    out["levels"]["ave_route_time"] = np.where(end > 0, per_level[:, nat.CTR["route_time"]] / np.maximum(end, 1), 0.0)
    if q_levels is not None:
        q = torch.zeros((L, 2), dtype=torch.float64, device=counters.device)
        q[:q_levels.shape[0]] = q_levels                 # levels this rank holds no replica of stay zero
        q = allreduce_sum_(q).cpu().numpy()
        out["levels"]["q_sum"], out["levels"]["q_abs_delta"] = q[:, 0].copy(), q[:, 1].copy()
        out["q_sum"], out["q_abs_delta"] = float(q[:, 0].sum()), float(q[:, 1].sum())
    return out


def global_stats(network, level_index=None, n_levels=None, q_prev=None, table="Qtable"):
    """Whole-job statistics of a Network — THE collective of the path.

    Always: the sums of the per-replica counters (rlr_reduce_stats on the device, then one int64 all-reduce).
    With `level_index` (this rank's length-R vector of load-level ids, levels numbered globally 0..n_levels-1): the same
    counters per load level ([n_levels, 8] int64, one more all-reduce) and each level's mean delivery time; and, for the
    policies that learn a table, the Q-convergence statistics of rlr_table_stats per level — sum of the table's entries
    and sum of |entry - q_prev entry| (what the TD updates of qroute.py:40-44 moved since the copy `q_prev` was taken) —
    all-reduced in fp64.  Every rank must call it with the same `n_levels`."""
    q_levels = None
    if level_index is not None and table in getattr(network.agent, "_table_names", ()):
        q_levels = network.table_stats(prev=q_prev, level_index=level_index, table=table)["levels"]
    return combine_stats(network._counters, level_index, n_levels, q_levels, totals=network.stats_device())
