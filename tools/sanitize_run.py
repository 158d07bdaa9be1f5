#!/usr/bin/env python
"""Small runs of every step-kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_run.py [case ...]

Each case trains a handful of replicas for a few dozen steps through the public API (fused launch, split phases and,
where the policy has one, the general timing model) and checks the counters against the CPU oracle, so a race that
changed a result would also show up as a mismatch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rlrouting_b200 as rb          # noqa: E402
from oracle import oracle as O       # noqa: E402

CASES = {
    "qroute": ("6x6.net", "Qroute", {}, {}),
    "qroute_hbm": ("6x6.net", "Qroute", {"tables_in_smem": 0}, {}),
    "shortest": ("6x6.net", "Shortest", {}, {}),
    "hybridq": ("6x6.net", "HybridQ", {}, {}),
    "hybridq_smem": ("6x6.net", "HybridQ", {"tables_in_smem": 1}, {}),
    "mahybridq": ("6x6.net", "MaHybridQ", {}, {}),
    "cq": ("6x6.net", "CQ", {}, {}),
    "cdrq": ("6x6.net", "CDRQ", {}, {}),
    "globalroute": ("6x6.net", "GlobalRoute", {}, {}),
    "general_qroute": ("6x6.net", "Qroute", {"transtime": 3, "bandwidth": 2}, {}),
    "general_cdrq": ("6x6.net", "CDRQ", {"transtime": 2}, {}),
    "lata_qroute": ("lata.net", "Qroute", {}, {}),
    "lata_hybridq": ("lata.net", "HybridQ", {}, {}),
}


def run(name):
    net, policy, nkw, akw = CASES[name]
    R, T, seed, lam = 5, 40, 3, 2.5
    lr = {"q": 0.1, "p": 0.001 if policy == "MaHybridQ" else 0.1, "e": 0.1}
    nw = rb.Network(net, replicas=R, rng="philox", seed=seed, queue_capacity=256, **nkw)
    agent = getattr(rb, policy)(nw, **akw)
    nw.agent = agent
    q0 = agent.table_array("Qtable").copy() if "Qtable" in getattr(agent, "_table_names", ()) else None
    nw.train(T, lam, lr=lr, steps_per_launch=17)
    rewards = nw.step()                                     # split phases
    agent.learn(rewards, lr=lr) if hasattr(agent, "learn") and policy not in ("Shortest",) else None
    c = nw.counters()
    bad = 0
    for r in range(R):
        e = O.OracleEnv(nw.topology.links, policy, seed=seed, replica=r, bandwidth=nkw.get("bandwidth", 1),
                        transtime=nkw.get("transtime", 1))
        if q0 is not None:
            e.set_qtable(q0[r])
        e.run(T, lambd=lam, lr_q=lr["q"], lr_p=lr["p"], lr_e=lr["e"])
        oc = e.counters()
        # (the extra split step injects nothing, so all_packets is unchanged by it; compare what the fused run fixed)
        bad += int(c["all_packets"][r]) != oc["all_packets"]
    print(f"sanitize_run {name}: {R} replicas x {T} steps + one split step, all_packets mismatches: {bad}", flush=True)
    return bad


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    sys.exit(1 if sum(run(n) for n in names) else 0)
