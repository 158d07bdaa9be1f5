"""Randomised stress run (GPU): trace mode against the oracle fed with the same NumPy replay (Python and C schedule paths),
snapshot/restore at random points, sample_route_time.  python tools/stress_modes.py [seed]"""
import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, pathlib, tempfile
import rlrouting_b200 as rb
from rlrouting_b200 import trace
from oracle import oracle as O
import test_gpu_parity as t
seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rng = np.random.default_rng(seed0)
d = pathlib.Path(tempfile.mkdtemp())
pol_trace = ["Shortest", "Qroute", "CQ", "CDRQ", "GlobalRoute"]
okA = okB = okC = 0
for trial in range(40):
    # A. trace mode (R small -> python lists, R >= 64 -> C batch), general timing, freq/slot, vs oracle fed the same replay
    policy = pol_trace[trial % len(pol_trace)]
    n = int(rng.integers(4, 30)); path = d / f"a{trial}.net"
    path.write_text(t._random_net(rng, n, int(rng.integers(0, 2 * n)), int(rng.integers(2, 6))))
    nkw = {}
    if rng.random() < 0.5: nkw["transtime"] = int(rng.integers(1, 4))
    R = int(rng.choice([1, 3, 70]))
    T = int(rng.integers(30, 120)); slot, freq = int(rng.integers(1, 3)), int(rng.integers(1, 3))
    lam = float(rng.uniform(0.2, 0.08 * n + 0.4)); seed = int(rng.integers(1, 1000))
    nw = rb.Network(str(path), replicas=R, rng="trace", seed=seed, queue_capacity=512, **nkw)
    agent = getattr(rb, policy)(nw); nw.agent = agent
    q0 = agent.table_array("Qtable") if policy not in ("Shortest", "GlobalRoute") else None
    res = nw.train(T * slot, lam, slot=slot, freq=freq, steps_per_launch=int(rng.integers(5, 60)))
    c = nw.counters()
    topo = nw.topology
    for r in range(R):
        rs = np.random.RandomState(seed + r)
        if q0 is not None:
            q = topo.pack(trace.draw_qtable_normals(rs, topo))
        inj = trace.draw_injections(rs, n, T, lam * slot)
        e = O.OracleEnv(topo.links, policy, transtime=nkw.get("transtime", 1))
        if q0 is not None: e.set_qtable(q)
        out = e.run(T, inj=inj, slot=slot, freq=freq)
        oc = e.counters()
        assert all(int(c[k][r]) == oc[k] for k in ("all_packets", "end_packets", "hops", "route_time", "sends")), ("A", trial, policy, r, nkw, slot, freq)
        if q0 is not None:
            assert np.array_equal(agent.table_array("Qtable")[r].view(np.int64), e.Q.view(np.int64)), ("A-Q", trial, policy, r)
    okA += 1
for trial in range(30):
    # B. snapshot at a random point + restore into a fresh network == uninterrupted run; philox, all policies
    policy = ["Shortest", "Qroute", "HybridQ", "MaHybridQ", "CQ", "CDRQ", "GlobalRoute"][trial % 7]
    n = int(rng.integers(4, 25)); path = d / f"b{trial}.net"
    path.write_text(t._random_net(rng, n, int(rng.integers(0, 2 * n)), int(rng.integers(2, 6))))
    nkw = {}
    if rng.random() < 0.6: nkw["transtime"] = int(rng.integers(1, 5))
    if rng.random() < 0.3: nkw["is_drop"] = True
    R, T1, T2 = int(rng.integers(1, 5)), int(rng.integers(10, 80)), int(rng.integers(10, 80))
    lam = float(rng.uniform(0.2, 0.08 * n + 0.4))
    lr = {"q": 0.1, "p": 0.001 if policy == "MaHybridQ" else 0.1}
    def make(cap):
        nw = rb.Network(str(path), replicas=R, rng="philox", seed=trial, replica_base=7, queue_capacity=cap, **nkw)
        ag = getattr(rb, policy)(nw); nw.agent = ag
        return nw, ag
    a, aa = make(256)
    try:
        a.train(T1, lam, lr=lr)
        snap = a.snapshot()
        ra = a.train(T2, lam, lr=lr)
    except ValueError:
        continue
    b, bb = make(512)
    b.restore(snap)
    rb_ = b.train(T2, lam, lr=lr)
    assert np.array_equal(ra["route_time"], rb_["route_time"]), ("B", trial, policy, nkw)
    ca, cb = a.counters(), b.counters()
    assert all(np.array_equal(ca[k], cb[k]) for k in ca), ("B-c", trial, policy, nkw)
    for name in aa._dev:
        assert np.array_equal(aa._dev[name].cpu().numpy(), bb._dev[name].cpu().numpy()), ("B-t", trial, policy, name)
    okB += 1
for trial in range(30):
    # C. sample_route_time vs oracle, philox, random size / timing
    policy = ["Shortest", "Qroute", "HybridQ", "CQ", "CDRQ", "GlobalRoute"][trial % 6]
    n = int(rng.integers(4, 25)); path = d / f"c{trial}.net"
    path.write_text(t._random_net(rng, n, int(rng.integers(0, 2 * n)), int(rng.integers(2, 6))))
    nkw = {}
    if rng.random() < 0.5: nkw["transtime"] = int(rng.integers(1, 4))
    size = int(rng.integers(5, 120)); slot, freq = int(rng.integers(1, 3)), int(rng.integers(1, 3))
    lam = float(rng.uniform(0.5, 0.08 * n + 0.6))
    nw = rb.Network(str(path), replicas=1, rng="philox", seed=trial, replica_base=11, queue_capacity=1024, **nkw)
    ag = getattr(rb, policy)(nw); nw.agent = ag
    q0 = ag.table_array("Qtable") if policy not in ("Shortest", "GlobalRoute") else None
    try:
        got = nw.sample_route_time(size, lam, slot=slot, freq=freq)
    except ValueError:
        continue
    e = O.OracleEnv(nw.topology.links, policy, seed=trial, replica=11, transtime=nkw.get("transtime", 1))
    if q0 is not None: e.set_qtable(q0[0])
    exp, steps = O.sample_route_time(e, size, lambd=lam, slot=slot, freq=freq)
    assert [int(v) for v in got] == [int(v) for v in exp], ("C", trial, policy, nkw, size, slot, freq)
    assert nw.clock == e.counters()["clock"], ("C-clock", trial, policy, nw.clock, e.counters()["clock"])
    okC += 1
print("seed", seed0, "trace ok", okA, "snapshot ok", okB, "sample ok", okC)
