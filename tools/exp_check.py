#!/usr/bin/env python
"""Quick bit-exact check of the 6x6 Qroute step kernel against the CPU oracle, for kernel-tuning builds
(-DRLR_FAST_BUILD holds only that instantiation).  RLR_LIB=<lib.so> python tools/exp_check.py

Philox mode, one load per replica (0.5 .. 4.0), two consecutive train calls (state carried across launches,
several launches per call), counters + per-step vectors + final queues + Q-table bits compared."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rlrouting_b200 as rb              # noqa: E402
from oracle import oracle as O           # noqa: E402


def main():
    R, seed = 48, 5
    lam = np.linspace(0.5, 9.0, R)            # the high loads inject several packets per node and step
    nw = rb.Network("6x6.net", replicas=R, rng="philox", seed=seed, queue_capacity=4096)
    agent = rb.Qroute(nw)
    nw.agent = agent
    q0 = agent.table_array("Qtable").copy()
    parts = [nw.train(T, lam, steps_per_launch=spl)["route_time"] for T, spl in ((333, 100), (280, None))]
    got = np.concatenate(parts, axis=1)
    c = nw.counters()
    bad = 0
    for r in range(R):
        e = O.OracleEnv(nw.topology.links, "Qroute", seed=seed, replica=r)
        e.set_qtable(q0[r])
        o = e.run(333 + 280, lambd=float(lam[r]), lr_q=0.1, lr_p=0.1, lr_e=0.1)
        oc = e.counters()
        ok = all(int(c[k][r]) == oc[k] for k in ("all_packets", "end_packets", "hops", "route_time", "sends"))
        ok &= np.array_equal(agent.table_array("Qtable")[r].view(np.int64), e.Q.view(np.int64))
        with np.errstate(divide="ignore", invalid="ignore"):
            ort = np.where(o["end_packets"] > 0, o["route_time"] / np.maximum(o["end_packets"], 1), 0.0)
        ok &= np.array_equal(got[r], ort)
        for x in range(nw.topology.N):
            ok &= np.array_equal(nw.queue_array(x, r), e.queue(x)) if hasattr(e, "queue") else True
        bad += 0 if ok else 1
    print(f"exp_check {os.environ.get('RLR_LIB', 'default')}: {R - bad}/{R} replicas match the oracle bit for bit", flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
