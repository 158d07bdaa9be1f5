"""World-size-2 test of the N>1 path on CPU (gloo): replica sharding by global index + the one
all-reduce of integer statistics.  The per-rank 'simulation' here is the CPU oracle standing in for
a GPU's replica block; the sharding / reduction code under test is rlrouting_b200.dist."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from rlrouting_b200 import dist as rdist
from rlrouting_b200 import nets
from rlrouting_b200.topology import Topology

TOTAL, T, LOADS, PER_LEVEL = 12, 120, [1.0, 2.0, 3.0], 4
KEYS = ("all_packets", "end_packets", "drop_packets", "hops", "route_time", "sends")


def _simulate(base, lam, policy="Shortest", qstats=False):
    """Counters [count, 8] of replicas base..base+count-1 (Philox streams keyed by GLOBAL index); with `qstats` also
    each replica's {sum Q, sum |Q - Q0|} after the run (what Network.table_stats computes on the device)."""
    topo = Topology.from_text(nets.net_text("6x6.net"))
    rows, qs = [], []
    for i, l in enumerate(lam):
        e = O.OracleEnv(topo.links, policy, seed=5, replica=base + i)
        if qstats:
            q0 = np.random.Generator(np.random.Philox(key=[5, base + i])).standard_normal(topo.tsize)
            e.set_qtable(q0)
            q0 = e.Q.copy()
        e.run(T, lambd=float(l))
        c = e.counters()
        rows.append([c[k] for k in KEYS] + [0, 0])
        if qstats:
            qs.append([e.Q.sum(), np.abs(e.Q - q0).sum()])
    if qstats:
        return torch.tensor(rows, dtype=torch.int64), np.array(qs)
    return torch.tensor(rows, dtype=torch.int64)


def _worker_q(rank, world, port, out):
    """rdist.combine_stats — the reduction behind global_stats — over gloo: integer counters and fp64 Q statistics per level."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    rdist.init_from_env(backend="gloo")
    base, lam, idx = rdist.shard_loads(LOADS, PER_LEVEL, rank, world)
    counters, qs = _simulate(base, lam, "Qroute", qstats=True)
    q_levels = torch.zeros((int(idx.max()) + 1, 2), dtype=torch.float64)
    q_levels.index_add_(0, torch.as_tensor(idx, dtype=torch.long), torch.from_numpy(qs))
    res = rdist.combine_stats(counters, idx, len(LOADS), q_levels)
    if rank == 0:
        out.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_level_and_q_convergence_statistics():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_q, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lam = np.repeat(LOADS, PER_LEVEL)
    whole, qs = _simulate(0, lam, "Qroute", qstats=True)
    whole = whole.numpy()
    for lvl in range(len(LOADS)):
        sl = slice(lvl * PER_LEVEL, (lvl + 1) * PER_LEVEL)
        assert res["levels"]["route_time"][lvl] == whole[sl, 4].sum() and res["levels"]["end_packets"][lvl] == whole[sl, 1].sum()
        assert res["levels"]["ave_route_time"][lvl] == whole[sl, 4].sum() / whole[sl, 1].sum()
        assert np.isclose(res["levels"]["q_sum"][lvl], qs[sl, 0].sum(), rtol=1e-12)
        assert np.isclose(res["levels"]["q_abs_delta"][lvl], qs[sl, 1].sum(), rtol=1e-12)
    assert res["q_abs_delta"] > 0 and res["sends"] == whole[:, 5].sum()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = rdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    base, lam, idx = rdist.shard_loads(LOADS, PER_LEVEL, rank, world)
    counters = _simulate(base, lam)
    per_level = rdist.level_stats(counters, idx, len(LOADS))
    rdist.allreduce_sum_(per_level)
    total = rdist.allreduce_sum_(counters.sum(0))
    tmax = rdist.allreduce_max_(torch.tensor([float(rank + 1)], dtype=torch.float64))
    if rank == 0:
        out.put((per_level.numpy(), total.numpy(), float(tmax.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_allreduce_match_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    per_level, total, tmax = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lam = np.repeat(LOADS, PER_LEVEL)
    whole = _simulate(0, lam).numpy()
    assert np.array_equal(total, whole.sum(0))
    for lvl in range(len(LOADS)):
        assert np.array_equal(per_level[lvl], whole[lvl * PER_LEVEL:(lvl + 1) * PER_LEVEL].sum(0))
    assert tmax == 2.0                                   # max-over-ranks timing reduction
    # mean delivery time per load level from the reduced integers (order-independent, bit-reproducible)
    assert all(per_level[:, 4] / np.maximum(per_level[:, 1], 1) > 0)
