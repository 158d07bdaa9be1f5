"""GPU parity tests: the CUDA path (through the C ABI, driven by the Python host layer) against the
CPU oracle and the committed known answers of the unmodified reference.  Bit-exact for integer
state and — because the device shares the oracle's deterministic exp/log2 — for the fp64 tables."""
import numpy as np
import pytest

from helpers import GOLDEN, case_inputs, load_cases, lr_of, run_oracle, sha, topo_of
from oracle import oracle as O

pytestmark = pytest.mark.gpu

CASES = load_cases()
POLICIES = {}


def _classes():
    import rlrouting_b200 as rb
    return {"Shortest": rb.Shortest, "Qroute": rb.Qroute, "HybridQ": rb.HybridQ, "MaHybridQ": rb.MaHybridQ,
            "CQ": rb.CQ, "CDRQ": rb.CDRQ, "GlobalRoute": rb.GlobalRoute}


def device_run(case, **net_kw):
    """Build Network+agent for a golden case and train; returns (nw, agent, result)."""
    import rlrouting_b200 as rb
    topo = topo_of(case["net"])
    kw = dict(case.get("agent_kwargs") or {})
    rng = case.get("rng", "trace")
    net_kw = dict(net_kw, **(case.get("net_kwargs") or {}))
    if rng == "trace":
        nw = rb.Network(case["net"], rng="trace", seed=case["seed"], **net_kw)
        agent = _classes()[case["policy"]](nw, **kw)
    else:
        nw = rb.Network(case["net"], rng="philox", seed=case.get("philox_seed", 0),
                        replica_base=case.get("replica", 0), **net_kw)
        agent = _classes()[case["policy"]](nw, **kw)
        q0, _ = case_inputs(case, topo)              # the reference drew its normals from np.random.seed(seed)
        if q0 is not None:
            agent.set_initial_q(q0, initP=kw.get("initP", 0.0))
    nw.agent = agent
    slot, freq = case.get("slot", 1), case.get("freq", 1)        # T counts iterations of train's loop: duration = T * slot
    if case.get("sample_size"):
        return nw, agent, nw.sample_route_time(case["sample_size"], case["lambd"], slot=slot, freq=freq, lr=case.get("lr") or {})
    exp_err = case["expect"]["error"]
    if exp_err is not None:
        with pytest.raises(ValueError, match="probabilities contain NaN"):
            nw.train(case["T"] * slot, case["lambd"], slot=slot, freq=freq, lr=case.get("lr") or {}, hop=True, sendlog=True)
        return nw, agent, None
    res = nw.train(case["T"] * slot, case["lambd"], slot=slot, freq=freq, lr=case.get("lr") or {}, hop=True, droprate=True,
                   sendlog=True)
    return nw, agent, res


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['id']}-{c['policy']}-{c['net']}-{c.get('rng', 'trace')}")
def test_device_matches_reference_known_answers(case):
    nw, agent, res = device_run(case)
    exp = case["expect"]
    if case.get("sample_size"):                       # Network.sample_route_time (env.py:416-438)
        assert [int(v) for v in res] == exp["sample"] and res.dtype == np.float64
        c = nw.counters()
        for k in ("clock", "all_packets", "end_packets", "drop_packets", "active_packets", "hops", "route_time"):
            assert int(c[k][0]) == exp["counters"][k], k
        qs = [nw.queue_array(x) for x in range(nw.topology.N)]
        assert sha(np.concatenate([q.reshape(-1) for q in qs] + [np.array([len(q) for q in qs])]), "<i4") == exp["queues_sha256"]
        for name in ("Qtable", "Theta", "Trace", "confidence"):
            if name in agent._dev and name + "_sha256" in exp:
                assert sha(agent.table_array(name)[0], "<f8") == exp[name + "_sha256"], name
        return
    if exp["error"] is not None:
        st = nw._status.cpu().numpy()[0]
        assert (int(st[0]), int(st[1])) == (1, exp["error"][1])
        return
    c = nw.counters()
    got = {k: int(c[k][0]) for k in ("clock", "all_packets", "end_packets", "drop_packets", "active_packets",
                                     "hops", "route_time", "sends")}
    assert got == exp["counters"]
    assert nw.ave_route_time == exp["ave_route_time"]
    log = nw.sendlog_rows(0)
    assert len(log) == exp["n_sends"]
    assert sha(log, "<i4") == exp["sendlog_sha256"]
    topo = nw.topology
    qs = [nw.queue_array(x) for x in range(topo.N)]
    assert sha(np.concatenate([q.reshape(-1) for q in qs] + [np.array([len(q) for q in qs])]), "<i4") == exp["queues_sha256"]
    # per-step cumulative metrics (env.py:409-413) against the oracle's integer vectors
    _, env, _, _, ores = run_oracle(case, sendlog=False)
    with np.errstate(divide="ignore", invalid="ignore"):
        ort = np.where(ores["end_packets"] > 0, ores["route_time"] / np.maximum(ores["end_packets"], 1), 0.0)
        ohp = np.where(ores["end_packets"] > 0, ores["hops"] / np.maximum(ores["end_packets"], 1), 0.0)
    assert np.array_equal(res["route_time"], ort)
    assert np.array_equal(res["hop"], ohp)
    assert res["droprate"][-1] == (got["drop_packets"] / got["all_packets"] if got["all_packets"] else 0)   # env.py:448-450
    assert res["droprate"].shape == ort.shape and np.all(np.diff(res["droprate"] * 0 + np.arange(len(ort))) == 1)
    bit_exact = case.get("rng", "trace") == "trace" or case.get("det_math") or case["policy"] in ("Qroute", "CQ", "CDRQ")
    if case["policy"] == "GlobalRoute":
        assert sha(agent.distance, "<f8") == exp["distance_sha256"]
        assert sha(np.concatenate([agent.choice[x].reshape(-1) for x in range(topo.N)]).astype(np.uint8), "u1") == exp["choice_sha256"]
        assert list(agent.queue_size) == exp["queue_size"]
        return
    if case["policy"] == "Shortest":
        assert sha(agent.distance, "<f8") == exp["distance_sha256"]
        assert sha(np.concatenate([agent.choice[x].reshape(-1) for x in range(topo.N)]).astype(np.uint8), "u1") == exp["choice_sha256"]
        return
    for name, oarr in (("Qtable", env.Q), ("Theta", env.Theta), ("Trace", env.Trace), ("confidence", env.Conf)):
        if name not in agent._dev:
            continue
        dev = agent.table_array(name)[0]
        assert np.array_equal(dev.view(np.int64), oarr.view(np.int64)), f"{name} differs from the oracle"
        if bit_exact and name + "_sha256" in exp:
            assert sha(dev, "<f8") == exp[name + "_sha256"], name
        elif not bit_exact:
            ref = np.load(f"{GOLDEN}/np_tables.npz")
            key = f"case{case['id']}_{name}"
            if key in ref:          # reference used NumPy's own exp/log2: north-star tolerance 1e-5 relative
                np.testing.assert_allclose(dev, ref[key], rtol=1e-5, atol=1e-9, err_msg=name)


def _oracle_batch(net, policy, R, T, lam, seed, base, q0, lr, agent_kw=None):
    topo = topo_of(net)
    kw = agent_kw or {}
    envs = []
    for r in range(R):
        e = O.OracleEnv(topo.links, policy, discount=kw.get("discount"), discount_trace=kw.get("discount_trace", 0.6),
                        add_entropy=kw.get("add_entropy", True), seed=seed, replica=base + r)
        if policy not in ("Shortest", "GlobalRoute"):
            e.set_qtable(q0[r], initP=kw.get("initP", 0.0))
        envs.append(e)
    sends, rt, en = O.run_many(envs, T, lam, lr_q=lr["q"], lr_p=lr["p"], lr_e=lr["e"], metrics=True)
    return envs, sends, rt, en


def _compare_batch(nw, agent, envs, rt, en, res):
    c = nw.counters()
    for r, e in enumerate(envs):
        oc = e.counters()
        for k in ("all_packets", "end_packets", "hops", "route_time", "sends", "active_packets"):
            assert int(c[k][r]) == oc[k], (r, k)
    with np.errstate(divide="ignore", invalid="ignore"):
        ort = np.where(en > 0, rt / np.maximum(en, 1), 0.0)
    assert np.array_equal(np.atleast_2d(res["route_time"]), ort)
    for name, attr in (("Qtable", "Q"), ("Theta", "Theta"), ("Trace", "Trace"), ("confidence", "Conf")):
        if name in agent._dev:
            dev = agent.table_array(name)
            for r, e in enumerate(envs):
                assert np.array_equal(dev[r].view(np.int64), getattr(e, attr).view(np.int64)), (name, r)
    for r in (0, len(envs) - 1):
        for x in range(nw.topology.N):
            assert np.array_equal(nw.queue_array(x, r), envs[r].queue(x)), (r, x)


BATCH = [
    # net, policy, R, T, lambda(s), launch config
    ("6x6.net", "Qroute", 37, 400, 2.0, {}),
    ("6x6.net", "Qroute", 37, 400, 2.0, dict(replicas_per_block=7, threads_per_block=256)),
    ("6x6.net", "Qroute", 20, 300, 1.0, dict(tables_in_smem=0, replicas_per_block=3)),
    ("6x6-modified.net", "Qroute", 16, 500, 2.5, dict(replicas_per_block=2)),
    ("lata.net", "Qroute", 6, 300, 2.0, {}),
    ("6x6.net", "Shortest", 33, 600, "sweep", dict(replicas_per_block=4)),
    ("lata.net", "Shortest", 5, 300, 3.0, {}),
    ("6x6.net", "HybridQ", 24, 300, 2.0, {}),
    ("6x6.net", "HybridQ", 10, 300, 2.0, dict(tables_in_smem=1)),          # the default keeps HybridQ's tables in HBM/L2
    ("6x6.net", "HybridQ", 10, 200, 1.0, dict(tables_in_smem=1, replicas_per_block=3, threads_per_block=128)),
    ("6x6-modified.net", "HybridQ", 9, 300, 1.5, dict(replicas_per_block=3)),
    ("lata.net", "HybridQ", 3, 200, 2.0, {}),
    ("6x6.net", "MaHybridQ", 10, 300, 2.0, {}),
    ("6x6.net", "MaHybridQ", 10, 300, 2.0, dict(replicas_per_block=2, threads_per_block=192)),
    ("6x6-modified.net", "MaHybridQ", 5, 200, 1.0, dict(tables_in_smem=0)),
    ("lata.net", "MaHybridQ", 2, 100, 2.0, {}),
    ("6x6.net", "CQ", 12, 400, 2.0, {}),
    ("6x6-modified.net", "CQ", 7, 300, 1.0, dict(replicas_per_block=2, tables_in_smem=0)),
    ("6x6.net", "CDRQ", 12, 500, 2.5, {}),
    ("6x6.net", "CDRQ", 9, 400, 1.0, dict(replicas_per_block=3, threads_per_block=256)),
    ("lata.net", "CDRQ", 3, 250, 3.0, {}),
    ("6x6.net", "GlobalRoute", 11, 300, 3.0, {}),
    ("6x6-modified.net", "GlobalRoute", 6, 250, 4.0, dict(replicas_per_block=2)),
    ("lata.net", "GlobalRoute", 2, 80, 4.0, {}),
]


@pytest.mark.parametrize("net,policy,R,T,lam,cfg", BATCH, ids=lambda v: str(v) if not isinstance(v, dict) else "-".join(f"{k}{x}" for k, x in v.items()) or "auto")
def test_device_batch_matches_oracle(net, policy, R, T, lam, cfg):
    """R Philox replicas on the device == R independent oracle replicas, bit for bit, for every
    launch geometry (replicas per block, threads per block, tables in shared memory or in HBM)."""
    import rlrouting_b200 as rb
    base, seed = 1000, 42
    if lam == "sweep":
        lam = np.linspace(1.0, 4.0, R)
    nw = rb.Network(net, replicas=R, rng="philox", seed=seed, replica_base=base, **cfg)
    agent = _classes()[policy](nw)
    nw.agent = agent
    lr = {"q": 0.1, "p": 0.001 if policy == "MaHybridQ" else 0.1, "e": 0.1}
    q0 = agent.table_array("Qtable") if policy not in ("Shortest", "GlobalRoute") else None
    # the device applied qroute.py:16-20 itself; hand the oracle the raw draws and let it do the same
    res = nw.train(T, lam, lr=lr)
    envs, sends, rt, en = _oracle_batch(net, policy, R, T, lam, seed, base, q0, lr)
    assert int(nw.counters()["sends"].sum()) == sends
    _compare_batch(nw, agent, envs, rt, en, res)


def test_init_tables_structure_matches_oracle():
    import rlrouting_b200 as rb
    for net in ("6x6.net", "6x6-modified.net", "lata.net"):
        nw = rb.Network(net, replicas=3, rng="philox", seed=5)
        agent = rb.MaHybridQ(nw, initQ=-1.5, initP=0.25)
        topo = nw.topology
        rng = np.random.default_rng(0)
        raw = rng.normal(size=(3, topo.tsize))
        agent.set_initial_q(raw, initP=0.25)
        for r in range(3):
            e = O.OracleEnv(topo.links, "MaHybridQ")
            e.set_qtable(raw[r], initP=0.25)
            assert np.array_equal(agent.table_array("Qtable")[r].view(np.int64), e.Q.view(np.int64))
            assert np.array_equal(agent.table_array("Theta")[r], e.Theta)
            assert np.array_equal(agent.table_array("Trace")[r], e.Trace)
        # reference-shaped views (qroute.py:13-20): zero row x, -eye rows at the neighbours
        Q = agent.Qtable
        for x in range(topo.N):
            assert Q[x].shape == (3, topo.N, topo.deg[x])
            assert np.all(Q[x][:, x, :] == 0)
            assert np.array_equal(Q[x][0][topo.links[x]], -np.eye(topo.deg[x]))


def test_shortest_tables_match_oracle_on_random_graphs():
    import os
    import tempfile
    import rlrouting_b200 as rb
    rng = np.random.default_rng(3)
    for trial in range(6):
        n = int(rng.integers(2, 40))
        edges = [(i, int(rng.integers(0, i))) for i in range(1, n)]          # random spanning tree
        extra = int(rng.integers(0, n))
        for _ in range(extra):
            a, b = (int(v) for v in rng.integers(0, n, 2))
            if a != b and (a, b) not in edges and (b, a) not in edges:
                edges.append((a, b))
        text = "".join(f"1000 {i} 0 0 0\n" for i in range(n)) + "".join(f"2000 {a} {b} 0\n" for a, b in edges)
        with tempfile.NamedTemporaryFile("w", suffix=".net", delete=False) as f:
            f.write(text)
        try:
            topo = rb.Topology.from_text(text)
            if topo.max_deg > 16:
                continue
            nw = rb.Network(f.name, replicas=2, rng="philox", seed=trial)
            agent = rb.Shortest(nw)
            e = O.OracleEnv(topo.links, "Shortest")
            assert np.array_equal(agent.distance, e.distance)
            assert np.array_equal(np.concatenate([agent.choice[x].reshape(-1) for x in range(n)]).astype(np.uint8), e.choice)
            nw.agent = agent
            res = nw.train(200, 1.0)
            envs, sends, rt, en = _oracle_batch_links(topo, 2, 200, 1.0, trial)
            _compare_batch(nw, agent, envs, rt, en, res)
        finally:
            os.unlink(f.name)


def _oracle_batch_links(topo, R, T, lam, seed):
    envs = [O.OracleEnv(topo.links, "Shortest", seed=seed, replica=r) for r in range(R)]
    sends, rt, en = O.run_many(envs, T, lam, metrics=True)
    return envs, sends, rt, en


def test_shard_invariance():
    """a replica's trajectory depends on its GLOBAL index only, so 1x24 == 3x8."""
    import rlrouting_b200 as rb
    R, T = 24, 300
    whole = rb.Network("6x6.net", replicas=R, rng="philox", seed=9)
    a = rb.Qroute(whole)
    whole.agent = a
    whole.train(T, 2.0)
    cw = whole.counters()
    tw = a.table_array("Qtable")
    tot = np.zeros(8, dtype=np.int64)
    for s in range(3):
        part = rb.Network("6x6.net", replicas=8, rng="philox", seed=9, replica_base=8 * s, replicas_per_block=3)
        b = rb.Qroute(part)
        part.agent = b
        part.train(T, 2.0)
        cp = part.counters()
        for k in ("all_packets", "end_packets", "hops", "route_time", "sends"):
            assert np.array_equal(cp[k], cw[k][8 * s:8 * s + 8]), k
        assert np.array_equal(b.table_array("Qtable").view(np.int64), tw[8 * s:8 * s + 8].view(np.int64))
        tot += part.stats_device().cpu().numpy()
    assert np.array_equal(tot, whole.stats_device().cpu().numpy())
    assert int(tot[5]) == int(cw["sends"].sum())


def test_chunked_launches_equal_single_launch_and_reset():
    import rlrouting_b200 as rb
    out = []
    for spl in (None, 37):
        nw = rb.Network("6x6.net", replicas=5, rng="philox", seed=3)
        ag = rb.HybridQ(nw)
        nw.agent = ag
        r1 = nw.train(200, 2.0, steps_per_launch=spl)
        r2 = nw.train(150, 1.0, steps_per_launch=spl)              # continues from clock 200
        out.append((np.concatenate([r1["route_time"], r2["route_time"]], axis=1), ag.table_array("Theta"), nw.counters()))
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1].view(np.int64), out[1][1].view(np.int64))
    assert all(np.array_equal(out[0][2][k], out[1][2][k]) for k in out[0][2])
    # reset keeps the tables, zeroes the environment (env.py:267-280)
    q_before = ag.table_array("Qtable")
    nw.reset()
    assert nw.clock == 0 and int(nw.counters()["all_packets"].sum()) == 0
    assert np.array_equal(ag.table_array("Qtable"), q_before)
    assert all(len(nw.queue_array(x, 0)) == 0 for x in range(36))


def test_edge_cases_zero_load_tiny_graph_overflow():
    import os
    import tempfile
    import rlrouting_b200 as rb
    nw = rb.Network("6x6.net", replicas=3, rng="philox", seed=1)
    nw.agent = rb.Qroute(nw)
    res = nw.train(50, 0.0)                                          # no traffic at all
    assert np.all(res["route_time"] == 0) and int(nw.counters()["all_packets"].sum()) == 0
    with tempfile.NamedTemporaryFile("w", suffix=".net", delete=False) as f:
        f.write("1000 0 0 0 0\n1000 1 0 0 0\n2000 1 0 0\n")         # two nodes, one edge
    try:
        nw = rb.Network(f.name, replicas=4, rng="philox", seed=2)
        ag = rb.MaHybridQ(nw)
        nw.agent = ag
        topo = nw.topology
        q0 = ag.table_array("Qtable")
        res = nw.train(300, 0.7, lr={"q": 0.1, "p": 0.001})
        envs = []
        for r in range(4):
            e = O.OracleEnv(topo.links, "MaHybridQ", seed=2, replica=r)
            e.set_qtable(q0[r])
            envs.append(e)
        sends, rt, en = O.run_many(envs, 300, 0.7, lr_q=0.1, lr_p=0.001, metrics=True)
        _compare_batch(nw, ag, envs, rt, en, res)
    finally:
        os.unlink(f.name)
    nw = rb.Network("6x6.net", replicas=2, rng="philox", seed=1, queue_capacity=8)
    nw.agent = rb.Shortest(nw)
    with pytest.raises(OverflowError, match="queue ring overflow"):
        nw.train(2000, 4.0)                                          # congested: rings of 8 must overflow loudly


def test_full_size_conservation_properties():
    """BASELINE-size run (8192 Qroute replicas): size-independent invariants of the simulator."""
    import rlrouting_b200 as rb
    R, T = 8192, 400
    nw = rb.Network("6x6.net", replicas=R, rng="philox", seed=11, queue_capacity=512)
    ag = rb.Qroute(nw)
    nw.agent = ag
    res = nw.train(T, 2.0, per_step=False)
    c = nw.counters()
    qlen = (nw._q_tail - nw._q_head).cpu().numpy().astype(np.int64)
    assert np.all(qlen >= 0)
    assert np.array_equal(qlen.sum(axis=1), c["active_packets"])             # every live packet sits in exactly one queue
    assert np.all(c["sends"] <= T * 36) and np.all(c["hops"] <= c["sends"])
    assert np.all(c["route_time"] >= c["hops"])                              # a hop costs at least one time unit
    assert abs(c["all_packets"].mean() / T - 2.0) < 0.02                      # Poisson(2.0) per step
    assert res["route_time"].shape == (R, 1)
    q = ag.table_array("Qtable")
    assert np.all(np.isfinite(q))
    # replicas are independent streams: no two identical trajectories
    assert len(np.unique(c["route_time"])) > R // 2
    # spot-check 3 replicas against the oracle
    topo = nw.topology
    nw2 = rb.Network("6x6.net", replicas=R, rng="philox", seed=11, queue_capacity=64)
    ag2 = rb.Qroute(nw2)
    q0 = ag2.table_array("Qtable")
    for r in (0, 4097, R - 1):
        e = O.OracleEnv(topo.links, "Qroute", seed=11, replica=r)
        e.set_qtable(q0[r])
        e.run(T, lambd=2.0)
        oc = e.counters()
        assert all(int(c[k][r]) == oc[k] for k in ("all_packets", "end_packets", "hops", "route_time", "sends"))
        assert np.array_equal(q[r].view(np.int64), e.Q.view(np.int64))


def _random_net(rng, n, extra, max_deg):
    deg = np.zeros(n, dtype=int)
    edges = []
    for i in range(1, n):                                   # random spanning tree with a degree cap
        cand = [j for j in range(i) if deg[j] < max(max_deg - 1, 1)] or [j for j in range(i) if deg[j] < max_deg]
        j = cand[int(rng.integers(0, len(cand)))]
        edges.append((i, j)); deg[i] += 1; deg[j] += 1
    have = set(edges) | {(b, a) for a, b in edges}
    for _ in range(extra):
        a, b = (int(v) for v in rng.integers(0, n, 2))
        if a != b and (a, b) not in have and deg[a] < max_deg and deg[b] < max_deg:
            edges.append((a, b)); have |= {(a, b), (b, a)}; deg[a] += 1; deg[b] += 1
    return "".join(f"1000 {i} 0 0 0\n" for i in range(n)) + "".join(f"2000 {a} {b} 0\n" for a, b in edges)


@pytest.mark.parametrize("policy,n,extra,max_deg,R,T,lam,nkw", [
    ("Qroute", 255, 700, 16, 3, 150, 6.0, {}),      # maximum node count and degree: MAXD = 16, MAXW = 9 code paths
    ("Shortest", 255, 300, 12, 2, 200, 8.0, {}),
    ("CDRQ", 97, 200, 9, 2, 150, 4.0, {}),
    ("HybridQ", 70, 120, 7, 3, 120, 3.0, {}),
    ("MaHybridQ", 127, 200, 10, 2, 60, 3.0, {}),    # MaHybridQ's limit (NumPy single pairwise block: n < 128 senders)
    ("Qroute", 2, 0, 1, 5, 300, 0.8, {}),           # smallest graph
    ("CQ", 33, 40, 6, 4, 200, 2.0, {}),
    # degree 5..8 on at most 64 nodes: four links through the loop-free arrival code, packets on further links through the list
    ("Qroute", 40, 30, 7, 4, 300, 3.0, {}),
    ("Shortest", 50, 40, 8, 3, 300, 4.0, {}),
    # 65..128 nodes, degree 9..12: the 128-thread build (four window words) and, for two replicas per CTA, the five-word one
    ("Qroute", 100, 80, 12, 3, 200, 4.0, {}),
    ("HybridQ", 100, 60, 10, 2, 100, 3.0, {}),
    ("GlobalRoute", 150, 250, 11, 2, 60, 6.0, {}),
    # the general timing model (events in flight, busy links, queue scan) on the same kinds of graph
    ("Qroute", 255, 700, 16, 2, 100, 6.0, dict(transtime=2)),
    ("Shortest", 200, 300, 12, 2, 150, 8.0, dict(transtime=3, bandwidth=2)),
    ("HybridQ", 70, 120, 7, 3, 100, 3.0, dict(transtime=2, is_drop=True)),
    ("CDRQ", 97, 200, 9, 2, 120, 4.0, dict(transtime=4, bandwidth=3)),
    ("MaHybridQ", 40, 60, 6, 2, 80, 2.0, dict(transtime=2)),
    ("GlobalRoute", 60, 90, 8, 2, 60, 4.0, dict(transtime=2)),
    ("Qroute", 2, 0, 1, 4, 200, 0.8, dict(transtime=5)),
    ("CQ", 33, 40, 6, 3, 150, 2.0, dict(transtime=2, bandwidth=0)),     # bandwidth 0: nothing can ever be sent
])
def test_random_topologies_match_oracle(policy, n, extra, max_deg, R, T, lam, nkw, tmp_path):
    """Random graphs written as .net text (degree-1 leaves, hubs, up to the supported maxima): device == oracle."""
    import rlrouting_b200 as rb
    rng = np.random.default_rng(n * 1000 + extra)
    path = tmp_path / "g.net"
    path.write_text(_random_net(rng, n, extra, max_deg))
    nw = rb.Network(str(path), replicas=R, rng="philox", seed=17, replica_base=5, queue_capacity=1024, **nkw)
    agent = _classes()[policy](nw)
    nw.agent = agent
    topo = nw.topology
    lr = {"q": 0.1, "p": 0.001 if policy == "MaHybridQ" else 0.1, "e": 0.1}
    q0 = agent.table_array("Qtable") if policy not in ("Shortest", "GlobalRoute") else None
    res = nw.train(T, lam, lr=lr)
    envs = []
    for r in range(R):
        e = O.OracleEnv(topo.links, policy, seed=17, replica=5 + r, is_drop=nkw.get("is_drop", False),
                        bandwidth=nkw.get("bandwidth", 1), transtime=nkw.get("transtime", 1))
        if q0 is not None:
            e.set_qtable(q0[r])
        envs.append(e)
    sends, rt, en = O.run_many(envs, T, lam, lr_q=lr["q"], lr_p=lr["p"], lr_e=lr["e"], metrics=True)
    assert sends > 0 or lam == 0 or nkw.get("bandwidth") == 0
    _compare_batch(nw, agent, envs, rt, en, res)


@pytest.mark.parametrize("policy", ["GlobalRoute", "MaHybridQ", "CQ", "Qroute"])
def test_network_reset_semantics_match_oracle(policy):
    """Network.reset (env.py:267-280) zeroes the environment and calls agent.clean(): tables survive, MaHybridQ's
    traces and CQ's confidences are re-initialised, and GlobalRoute keeps its (now stale) queue_size like the
    reference does.  Train, reset, train again on both sides."""
    import rlrouting_b200 as rb
    R, seed = 4, 21
    nw = rb.Network("6x6.net", replicas=R, rng="philox", seed=seed)
    agent = _classes()[policy](nw)
    nw.agent = agent
    lr = {"q": 0.1, "p": 0.001 if policy == "MaHybridQ" else 0.1, "e": 0.1}
    q0 = agent.table_array("Qtable") if policy != "GlobalRoute" else None
    nw.train(150, 3.0, lr=lr)
    nw.reset()
    assert nw.clock == 0
    res = nw.train(200, 2.0, lr=lr)
    envs = []
    for r in range(R):
        e = O.OracleEnv(nw.topology.links, policy, seed=seed, replica=r)
        if q0 is not None:
            e.set_qtable(q0[r])
        e.run(150, lambd=3.0, lr_q=lr["q"], lr_p=lr["p"], lr_e=lr["e"])
        e.reset()
        envs.append(e)
    sends, rt, en = O.run_many(envs, 200, 2.0, lr_q=lr["q"], lr_p=lr["p"], lr_e=lr["e"], metrics=True)
    _compare_batch(nw, agent, envs, rt, en, res)
    if policy == "GlobalRoute":
        for r in range(R):
            assert list(agent.queue_size[r]) == list(envs[r].queue_size)


def test_sample_route_time_many_replicas_and_desync_guard():
    """R replicas sample until each has its deliveries; every replica equals its own oracle run, and the network
    refuses further training until reset() because the replicas stopped at different clocks."""
    import rlrouting_b200 as rb
    R, size, lam, seed = 9, 150, 2.0, 4
    nw = rb.Network("6x6.net", replicas=R, rng="philox", seed=seed)
    agent = rb.Qroute(nw)
    nw.agent = agent
    q0 = agent.table_array("Qtable")
    out = nw.sample_route_time(size, lam)
    assert out.shape == (R, size)
    clocks = []
    for r in range(R):
        e = O.OracleEnv(nw.topology.links, "Qroute", seed=seed, replica=r)
        e.set_qtable(q0[r])
        smp, steps = O.sample_route_time(e, size, lambd=lam)
        assert np.array_equal(out[r].astype(np.int64), smp), r
        assert np.array_equal(agent.table_array("Qtable")[r].view(np.int64), e.Q.view(np.int64)), r
        clocks.append(steps)
    assert nw.clock == max(clocks)
    if len(set(clocks)) > 1:
        with pytest.raises(RuntimeError, match="different clocks"):
            nw.train(10, lam)
    nw.reset()
    nw.train(10, lam)


TIMING = [
    # net, policy, R, T (iterations), lambda, Network kwargs, slot, freq, launch config
    ("6x6.net", "Qroute", 19, 300, 2.0, dict(transtime=2), 1, 1, {}),
    ("6x6.net", "Qroute", 16, 200, 1.0, dict(transtime=5, bandwidth=2), 2, 2, dict(replicas_per_block=3, threads_per_block=128)),
    ("6x6.net", "Qroute", 9, 200, 1.5, dict(transtime=1), 3, 2, {}),                 # slot > transtime: nothing outlives a step
    ("6x6.net", "Qroute", 8, 250, 2.0, dict(transtime=3, bandwidth=2, is_drop=True), 1, 1, dict(tables_in_smem=0, replicas_per_block=2)),
    ("6x6-modified.net", "Shortest", 12, 300, "sweep", dict(transtime=2), 1, 1, dict(replicas_per_block=4)),
    ("lata.net", "Shortest", 3, 200, 3.0, dict(transtime=3, bandwidth=3), 1, 2, {}),
    ("6x6.net", "HybridQ", 10, 250, 2.0, dict(transtime=2), 1, 1, {}),
    ("6x6.net", "MaHybridQ", 6, 200, 2.0, dict(transtime=3, bandwidth=2), 1, 1, {}),
    ("6x6.net", "CQ", 7, 250, 2.0, dict(transtime=2), 1, 1, {}),
    ("6x6.net", "CDRQ", 7, 300, 2.5, dict(transtime=2), 1, 1, {}),
    ("6x6.net", "CDRQ", 5, 200, 1.0, dict(transtime=4, bandwidth=2), 2, 1, dict(replicas_per_block=2, threads_per_block=256)),
    ("6x6.net", "GlobalRoute", 6, 200, 3.0, dict(transtime=2), 1, 1, {}),
    ("lata.net", "Qroute", 3, 150, 2.0, dict(transtime=2, bandwidth=2), 1, 1, {}),
]


@pytest.mark.parametrize("net,policy,R,T,lam,nkw,slot,freq,cfg", TIMING,
                         ids=lambda v: str(v) if not isinstance(v, dict) else "-".join(f"{k}{x}" for k, x in v.items()) or "auto")
def test_general_timing_model_matches_oracle(net, policy, R, T, lam, nkw, slot, freq, cfg):
    """integer bandwidth / transtime / slot / freq.  With transtime > slot events stay in flight across
    steps and launches (persistent heapq array, busy links, queue scan); R device replicas == R oracle replicas, bit for
    bit, across a chunked run (3 launches), followed by a second train call that continues from the in-flight state."""
    import rlrouting_b200 as rb
    base, seed = 500, 11
    if lam == "sweep":
        lam = np.linspace(1.0, 3.0, R)
    nw = rb.Network(net, replicas=R, rng="philox", seed=seed, replica_base=base, **nkw, **cfg)
    agent = _classes()[policy](nw)
    nw.agent = agent
    lr = {"q": 0.1, "p": 0.001 if policy == "MaHybridQ" else 0.1, "e": 0.1}
    q0 = agent.table_array("Qtable") if policy not in ("Shortest", "GlobalRoute") else None
    T1 = (2 * T) // 3
    res1 = nw.train(T1 * slot, lam, slot=slot, freq=freq, lr=lr, steps_per_launch=max(1, T1 // 3))
    res2 = nw.train((T - T1) * slot, lam, slot=slot, freq=freq, lr=lr)
    res = {"route_time": np.concatenate([np.atleast_2d(res1["route_time"]), np.atleast_2d(res2["route_time"])], axis=1)}
    topo = topo_of(net)
    envs = []
    for r in range(R):
        e = O.OracleEnv(topo.links, policy, seed=seed, replica=base + r, is_drop=nkw.get("is_drop", False),
                        bandwidth=nkw.get("bandwidth", 1), transtime=nkw.get("transtime", 1))
        if q0 is not None:
            e.set_qtable(q0[r])
        envs.append(e)
    sends, rt, en = O.run_many(envs, T, lam, lr_q=lr["q"], lr_p=lr["p"], lr_e=lr["e"], metrics=True, slot=slot, freq=freq)
    assert int(nw.counters()["sends"].sum()) == sends
    assert nw.clock == T * slot * freq == envs[0].counters()["clock"]
    _compare_batch(nw, agent, envs, rt, en, res)
    c = nw.counters()
    for r, e in enumerate(envs):
        assert int(c["drop_packets"][r]) == e.counters()["drop_packets"]
        if nw._general:
            assert int(nw._event_count[r].item()) == O.lib().orc_in_flight(e.h)
    # Network.reset clears the events in flight and the busy links too (env.py:267-280)
    nw.reset()
    for e in envs:
        e.reset()
    res3 = nw.train(40 * slot, lam, slot=slot, freq=freq, lr=lr)
    sends, rt, en = O.run_many(envs, 40, lam, lr_q=lr["q"], lr_p=lr["p"], lr_e=lr["e"], metrics=True, slot=slot, freq=freq)
    _compare_batch(nw, agent, envs, rt, en, res3)


def test_general_timing_manual_step_loop_equals_train():
    """The reference's own loop body (inject(new_packet) / step(duration) / agent.learn) under transtime > duration."""
    import rlrouting_b200 as rb
    out = []
    for manual in (False, True):
        nw = rb.Network("6x6.net", transtime=3, bandwidth=2, replicas=3, rng="philox", seed=4)
        ag = rb.HybridQ(nw)
        nw.agent = ag
        if manual:
            for _ in range(60):
                nw.inject(nw.new_packet(2.0 * 2))
                r = nw.step(2)
                ag.learn(r)
        else:
            nw.train(120, 2.0, slot=2)
        out.append((nw.counters(), ag.table_array("Theta"), ag.table_array("Qtable"), [nw.queue_array(x, 2) for x in range(36)]))
    assert all(np.array_equal(out[0][0][k], out[1][0][k]) for k in out[0][0])
    assert np.array_equal(out[0][1].view(np.int64), out[1][1].view(np.int64))
    assert np.array_equal(out[0][2].view(np.int64), out[1][2].view(np.int64))
    assert all(np.array_equal(a, b) for a, b in zip(out[0][3], out[1][3]))


def test_random_option_combinations_match_oracle(tmp_path):
    """A seeded sweep over combinations the targeted tests do not pair up: policy x timing model x is_drop x launch
    geometry x load, on small random graphs; every replica must match the oracle bit for bit."""
    import rlrouting_b200 as rb
    import os
    rng = np.random.default_rng(int(os.environ.get("RLR_SWEEP_SEED", "20240")))
    policies = ["Shortest", "Qroute", "HybridQ", "MaHybridQ", "CQ", "CDRQ", "GlobalRoute"]
    for trial in range(int(os.environ.get("RLR_SWEEP_TRIALS", "36"))):
        policy = policies[trial % len(policies)]
        n = int(rng.integers(3, 41))
        max_deg = int(rng.integers(2, 7))
        path = tmp_path / f"g{trial}.net"
        path.write_text(_random_net(rng, n, int(rng.integers(0, 2 * n)), max_deg))
        nkw = {}
        if rng.random() < 0.6:
            nkw["transtime"] = int(rng.integers(0, 5))
        if rng.random() < 0.4:
            nkw["bandwidth"] = int(rng.integers(1, 4))
        if rng.random() < 0.3:
            nkw["is_drop"] = True
        slot, freq = int(rng.integers(1, 4)), int(rng.integers(1, 3))
        cfg = {}
        if rng.random() < 0.5:
            g = int(rng.integers(1, 4))
            cfg = dict(replicas_per_block=g)
            if policy in ("MaHybridQ", "CQ", "CDRQ"):
                cfg["threads_per_block"] = min(512, ((g * n + 31) // 32) * 32 + 128)
        if rng.random() < 0.3 and policy not in ("Shortest", "GlobalRoute"):
            cfg["tables_in_smem"] = int(rng.integers(0, 2))
        akw = {}
        if policy in ("Shortest", "GlobalRoute") and rng.random() < 0.5:
            akw = dict(multiway=True, random=bool(rng.random() < 0.5))
        R, T, lam = int(rng.integers(1, 7)), int(rng.integers(20, 90)), float(rng.uniform(0.3, 0.15 * n + 1.0))
        what = (trial, policy, n, nkw, slot, freq, cfg, akw, R, T, lam)
        nw = rb.Network(str(path), replicas=R, rng="philox", seed=trial, replica_base=3 * trial, queue_capacity=1024,
                        **nkw, **cfg)
        try:
            agent = _classes()[policy](nw, **akw)
        except Exception as exc:                       # a forced geometry that cannot hold the tables in shared memory
            assert "shared-memory footprint" in str(exc) and cfg.get("tables_in_smem") == 1, (what, str(exc))
            continue
        nw.agent = agent
        lr = {"q": 0.1, "p": 0.001 if policy == "MaHybridQ" else 0.1, "e": 0.1}
        q0 = agent.table_array("Qtable") if policy not in ("Shortest", "GlobalRoute") else None
        try:
            res = nw.train(T * slot, lam, slot=slot, freq=freq, lr=lr)
        except ValueError as exc:                      # NaN probabilities: the oracle must stop at the same clock
            assert "NaN" in str(exc), what
            res = None
        envs = []
        for r in range(R):
            e = O.OracleEnv(nw.topology.links, policy, seed=trial, replica=3 * trial + r, is_drop=nkw.get("is_drop", False),
                            bandwidth=nkw.get("bandwidth", 1), transtime=nkw.get("transtime", 1),
                            multiway=akw.get("multiway", False), random=akw.get("random", False))
            if q0 is not None:
                e.set_qtable(q0[r])
            envs.append(e)
        sends, rt, en = O.run_many(envs, T, lam, lr_q=lr["q"], lr_p=lr["p"], lr_e=lr["e"], metrics=True, slot=slot, freq=freq)
        if res is None:
            st = nw._status.cpu().numpy()
            for r, e in enumerate(envs):
                assert (int(st[r, 0]) == 1) == (e.status()[0] == 1), what
                if e.status()[0] == 1:
                    assert int(st[r, 1]) == e.status()[1], what
            continue
        try:
            _compare_batch(nw, agent, envs, rt, en, res)
            c = nw.counters()
            for r, e in enumerate(envs):
                assert int(c["drop_packets"][r]) == e.counters()["drop_packets"]
        except AssertionError as exc:
            raise AssertionError(f"{what}: {exc}") from exc
