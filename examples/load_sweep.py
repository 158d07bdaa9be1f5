#!/usr/bin/env python
"""Load sweep over the GPUs of one box (BASELINE config 5 style): every (load level, replica) pair is one
environment, the grid is sharded over the ranks by GLOBAL replica index, and the per-level statistics are all-reduced
once at the end.

    python examples/load_sweep.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/load_sweep.py --policy MaHybridQ
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rlrouting_b200 as rb  # noqa: E402
from rlrouting_b200 import dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--net", default="6x6.net")
    ap.add_argument("--policy", default="Qroute")
    ap.add_argument("--replicas-per-level", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--loads", default="0.5:4.01:0.5")
    a = ap.parse_args()
    lo, hi, st = (float(v) for v in a.loads.split(":"))
    levels = np.arange(lo, hi, st)
    rank, world, local = dist.init_from_env()
    base, lam, idx = dist.shard_loads(levels, a.replicas_per_level, rank, world)
    nw = rb.Network(a.net, replicas=len(lam), rng="philox", seed=1, replica_base=base,
                    queue_capacity=1 << int(np.ceil(np.log2(256 + 0.25 * a.steps * max(levels.max() / 2, 1)))))
    nw.agent = getattr(rb, a.policy)(nw)
    lr = {"q": 0.1, "p": 0.001, "e": 0.1} if a.policy == "MaHybridQ" else {}
    nw.train(a.steps, lam, lr=lr, per_step=False)
    per_level = dist.level_stats(nw._counters, idx, len(levels))            # [levels, 8] int64 on the device
    dist.allreduce_sum_(per_level)                                          # the one collective of the run
    if rank == 0:
        s = per_level.cpu().numpy()
        print(f"{a.net} {a.policy}: {a.replicas_per_level} replicas x {len(levels)} load levels on {world} GPU(s), {a.steps} steps")
        print(" load   packets  delivered  mean delivery time  mean hops")
        for lvl, row in zip(levels, s):
            print(f"{lvl:5.2f} {row[0]:9d} {row[1]:10d} {row[4] / max(row[1], 1):19.4f} {row[3] / max(row[1], 1):10.3f}")
    if world > 1:
        import torch.distributed as td
        td.destroy_process_group()


if __name__ == "__main__":
    main()
