"""Randomised stress run (GPU): tiny ring capacities so the queues wrap many times, chunked launches, the general timing
model; every replica is compared with the CPU oracle.  python tools/stress_rings.py [seed]"""
import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, pathlib, tempfile
import rlrouting_b200 as rb
from oracle import oracle as O
import test_gpu_parity as t
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
d = pathlib.Path(tempfile.mkdtemp())
pol = ["Shortest", "Qroute", "HybridQ", "CQ", "CDRQ", "GlobalRoute", "MaHybridQ"]
ok = ovf = 0
for trial in range(120):
    policy = pol[trial % len(pol)]
    n = int(rng.integers(4, 30)); path = d / f"g{trial}.net"
    path.write_text(t._random_net(rng, n, int(rng.integers(0, 2 * n)), int(rng.integers(2, 6))))
    nkw = {}
    if rng.random() < 0.7: nkw["transtime"] = int(rng.integers(1, 5))
    if rng.random() < 0.4: nkw["bandwidth"] = int(rng.integers(1, 3))
    cap = int(rng.choice([16, 32, 64]))
    R, T = int(rng.integers(1, 5)), int(rng.integers(300, 900))
    lam = float(rng.uniform(0.2, 0.08 * n + 0.4))
    nw = rb.Network(str(path), replicas=R, rng="philox", seed=trial, replica_base=trial, queue_capacity=cap, **nkw)
    agent = getattr(rb, policy)(nw); nw.agent = agent
    lr = {"q": 0.1, "p": 0.001 if policy == "MaHybridQ" else 0.1, "e": 0.1}
    q0 = agent.table_array("Qtable") if policy not in ("Shortest", "GlobalRoute") else None
    spl = int(rng.integers(20, 200))
    try:
        res = nw.train(T, lam, lr=lr, steps_per_launch=spl)
    except OverflowError:
        ovf += 1; continue
    except ValueError:
        continue
    envs = []
    for r in range(R):
        e = O.OracleEnv(nw.topology.links, policy, seed=trial, replica=trial + r, bandwidth=nkw.get("bandwidth", 1), transtime=nkw.get("transtime", 1))
        if q0 is not None: e.set_qtable(q0[r])
        envs.append(e)
    sends, rt, en = O.run_many(envs, T, lam, lr_q=lr["q"], lr_p=lr["p"], lr_e=lr["e"], metrics=True)
    try:
        t._compare_batch(nw, agent, envs, rt, en, res)
        ok += 1
    except AssertionError as exc:
        print("FAIL", trial, policy, n, nkw, cap, R, T, lam, spl, str(exc)[:300]); break
    maxq = max(len(e.queue(x)) for e in envs for x in range(n))
print("ok", ok, "overflowed", ovf)
