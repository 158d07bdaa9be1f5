"""The reference's own script flow (train.py:8-57) against the drop-in modules in rlrouting_b200/compat,
checked against the reference's known answers (tests/golden/kat.json)."""
import os
import pickle
import sys

import numpy as np
import pytest

from helpers import load_cases

pytestmark = pytest.mark.gpu
COMPAT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rlrouting_b200", "compat")


@pytest.fixture()
def compat_path():
    sys.path.insert(0, COMPAT)
    saved = {m: sys.modules.pop(m, None) for m in ("env", "config", "qroute", "shortest", "hybrid", "multi_agent", "base_policy")}
    yield
    sys.path.remove(COMPAT)
    for m, v in saved.items():
        sys.modules.pop(m, None)
        if v is not None:
            sys.modules[m] = v


def test_reference_script_flow_config1(compat_path, tmp_path):
    """BASELINE config 1 through the reference's module names, its global-seed convention and the legacy calls
    train.py makes (bind, clean, train(..., lrq=, lrp=) -> tuple, store)."""
    ns = {}
    exec("from env import *\nfrom config import *\nfrom shortest import Shortest\nfrom qroute import Qroute\n"
         "from hybrid import HybridQ\nfrom multi_agent import MaHybridQ\n", ns)
    np.random.seed(1)                                   # train.py:15
    nw = ns["Network"]("6x6.net")                       # train.py:21
    agent = ns["Qroute"](nw)
    nw.bind(agent)                                      # train.py:39
    nw.clean()                                          # train.py:51
    route_time, drop_rate = nw.train(10000, lambd=2.0, lrq=0.1, lrp=0.001)     # train.py:52
    exp = load_cases()[0]["expect"]
    assert route_time.shape == (10000,) and drop_rate.shape == (10000,)
    assert route_time[-1] == exp["ave_route_time"] == 2081.741862022773
    assert (nw.all_packets, nw.end_packets, nw.route_time, nw.hops, nw.active_packets) == (19934, 14930, 31080406, 243694, 5004)
    assert nw.ave_route_time == exp["ave_route_time"] and nw.drop_rate == 0
    assert float(sum(v.sum() for v in agent.Qtable.values())) == pytest.approx(-3395208.639115376, rel=1e-12)
    import pandas as pd
    assert pd.DataFrame({2.0: route_time}).shape == (10000, 1)                  # train.py:56
    # checkpoint interchange (base_policy.py:57-66): same pickle keys and array shapes as the reference
    path = tmp_path / "q.pkl"
    agent.store(str(path))
    blob = pickle.load(open(path, "rb"))
    assert set(blob) == {"links", "Qtable", "discount", "threshold"}
    assert all(blob["Qtable"][x].shape == (36, len(nw.links[x])) for x in range(36))
    fresh = ns["Qroute"](nw)
    fresh.load(str(path))
    assert all(np.array_equal(fresh.Qtable[x], agent.Qtable[x]) for x in range(36))
    assert agent.choose(0, 35) in nw.links[0]
    # nodes / queues are readable like the reference's objects
    assert sum(len(nw.nodes[x].queue) for x in range(36)) == nw.active_packets


def test_load_sweep_with_clean_between_levels(compat_path):
    """train.py:43-52: one agent, loads swept with nw.clean() in between (tables carry over, env resets)."""
    from env import Network
    from shortest import Shortest
    np.random.seed(1)
    nw = Network("6x6.net")
    nw.agent = Shortest(nw)
    finals = {}
    for l in (1.0, 2.0):
        nw.clean()
        rt, _ = nw.train(10000, lambd=l, lrq=0.1, lrp=0.001)
        finals[l] = rt[-1]
    # load 1.0 from a fresh seed is golden case 6 (5.4241584158415845); load 2.0 continues the same stream
    assert finals[1.0] == load_cases()[6]["expect"]["ave_route_time"]
    assert 5.0 < finals[2.0] < 30.0


def test_unsupported_policies_and_calls_raise(compat_path):
    from base_policy import Policy
    from env import Network
    nw = Network("6x6.net")

    class Mine(Policy):
        def choose(self, source, dest):
            return self.links[source][0]
    with pytest.raises(NotImplementedError):
        nw.agent = Mine(nw)
    with pytest.raises(NotImplementedError):
        nw.train(10, 1.0)                       # no device policy bound
    with pytest.raises(NotImplementedError):
        nw.step(1)


@pytest.mark.parametrize("policy,rng,R", [("Qroute", "trace", 1), ("Shortest", "trace", 1), ("HybridQ", "philox", 3),
                                          ("MaHybridQ", "philox", 2), ("Qroute", "philox", 4), ("CQ", "trace", 1),
                                          ("CQ", "philox", 3), ("CDRQ", "trace", 1), ("CDRQ", "philox", 3),
                                          ("GlobalRoute", "philox", 2)])
def test_manual_loop_equals_fused_train(policy, rng, R):
    """The reference's own loop body (env.py:400-408) spelled out by the caller — inject(new_packet(l)); r = step(1);
    agent.learn(r) — must leave exactly the state Network.train leaves, and the rewards must be the send list."""
    import rlrouting_b200 as rb
    T, lam = 120, 2.0
    lr = {"q": 0.1, "p": 0.001, "e": 0.1} if policy == "MaHybridQ" else {}
    kw = dict(replicas=R, rng=rng, seed=5)
    fused = rb.Network("6x6.net", **kw)
    fa = getattr(rb, policy)(fused)
    fused.agent = fa
    fused.train(T, lam, lr=lr, sendlog=True)
    manual = rb.Network("6x6.net", **kw)
    ma = getattr(rb, policy)(manual)
    manual.agent = ma
    sends = [[] for _ in range(R)]
    for i in range(T):
        pk = manual.new_packet(lam)
        manual.inject(pk)
        assert np.all(manual.active_packets >= 0)
        r = manual.step(1)
        for rr, lst in enumerate([r] if R == 1 else r):
            sends[rr] += [(manual.clock, w.source, w.action, w.dest) for w in lst]
            for w in lst:
                assert w.agent_info["t_y"] == (0 if policy == "CDRQ" else 1) and w.agent_info["q_y"] >= 0 and w.packet.dest == w.dest
                if policy == "CDRQ":                   # Node._build_info_dual (env.py:154-160) + CDRQ.get_info (qroute.py:107-115)
                    assert w.agent_info["q_y"] >= 1 and w.agent_info["q_x"] >= 1 and w.agent_info["t_x"] == 0
                    assert set(w.agent_info) >= {"max_Q_b", "max_Q_f", "C_b", "C_f"}
        ma.learn(r, lr=lr) if lr else ma.learn(r)
    cf, cm = fused.counters(), manual.counters()
    for k in cf:
        assert np.array_equal(cf[k], cm[k]), k
    for name in fa._dev:
        if name in ("Qtable", "Theta", "Trace", "confidence"):
            assert np.array_equal(fa.table_array(name).view(np.int64), ma.table_array(name).view(np.int64)), name
    for rr in range(R):
        assert np.array_equal(np.array(sends[rr], dtype=np.int32).reshape(-1, 4), fused.sendlog_rows(rr))
        for x in range(36):
            assert np.array_equal(fused.queue_array(x, rr), manual.queue_array(x, rr))
    if policy != "Shortest":                           # Shortest.learn is the base class's no-op, as in the reference
        with pytest.raises(NotImplementedError):
            ma.learn([])                               # not the list of the latest step


@pytest.mark.parametrize("policy,rng,nkw,akw", [
    ("Qroute", "philox", {}, {}),
    ("HybridQ", "philox", {}, {}),
    ("MaHybridQ", "philox", {}, {}),
    ("CDRQ", "trace", {}, {}),
    ("GlobalRoute", "philox", {}, {}),
    ("Shortest", "trace", {}, {"multiway": True}),
    ("Qroute", "philox", {"transtime": 3, "bandwidth": 2, "is_drop": True}, {}),
    ("Shortest", "trace", {"transtime": 2}, {}),
])
def test_snapshot_restore_continues_bit_for_bit(policy, rng, nkw, akw, tmp_path):
    """environment-state snapshots.  run(T1) -> snapshot -> run(T2) must equal a fresh network that
    loads the snapshot (from a file, into rings of another capacity) and runs T2."""
    import rlrouting_b200 as rb
    cls = getattr(rb, policy)
    lr = {"q": 0.1, "p": 0.001 if policy == "MaHybridQ" else 0.1}
    R, T1, T2, lam = 5, 150, 120, 2.0

    def make(cap):
        nw = rb.Network("6x6.net", replicas=R, rng=rng, seed=7, replica_base=40, queue_capacity=cap, **nkw)
        ag = cls(nw, **akw)
        nw.agent = ag
        return nw, ag

    a, ag_a = make(512)
    a.train(T1, lam, lr=lr)
    a.save_state(tmp_path / "state.pkl")
    res_a = a.train(T2, lam, lr=lr, hop=True)
    b, ag_b = make(1024)
    b.load_state(tmp_path / "state.pkl")
    assert b.clock == T1
    res_b = b.train(T2, lam, lr=lr, hop=True)
    assert np.array_equal(res_a["route_time"], res_b["route_time"]) and np.array_equal(res_a["hop"], res_b["hop"])
    ca, cb = a.counters(), b.counters()
    assert all(np.array_equal(ca[k], cb[k]) for k in ca)
    for name in ag_a._dev:
        assert np.array_equal(ag_a._dev[name].cpu().numpy(), ag_b._dev[name].cpu().numpy()), name
    for r in (0, R - 1):
        for x in range(36):
            assert np.array_equal(a.queue_array(x, r), b.queue_array(x, r))
    if a._general:
        assert np.array_equal(a._event_count.cpu().numpy(), b._event_count.cpu().numpy())
        assert np.array_equal(a._link_sent.cpu().numpy(), b._link_sent.cpu().numpy())
    with pytest.raises(ValueError, match="does not fit"):
        other = rb.Network("6x6.net", replicas=R + 1, rng=rng, seed=7, **nkw)
        other.agent = cls(other, **akw)
        other.load_state(tmp_path / "state.pkl")


@pytest.mark.parametrize("policy,keys,akw", [
    ("Shortest", {"links", "distance", "choice", "mask"}, {}),
    ("Shortest", {"links", "distance", "choice", "mask"}, {"multiway": True, "random": True}),
    ("GlobalRoute", {"links", "distance", "choice", "mask"}, {}),
    ("CQ", {"links", "Qtable", "discount", "threshold", "confidence", "decay"}, {}),
    ("CDRQ", {"links", "Qtable", "discount", "threshold", "confidence", "decay"}, {}),
    ("HybridQ", {"links", "Qtable", "discount", "threshold", "Theta"}, {}),
    ("MaHybridQ", {"links", "Qtable", "discount", "threshold", "Theta", "discount_trace", "Trace"}, {}),
])
def test_store_load_round_trip_every_agent(policy, keys, akw, tmp_path):
    """Policy.store / load (base_policy.py:57-66) for every device agent: the pickle has the reference's keys and
    per-node (N, deg_x) shapes, and an agent that loads it routes exactly like the one that stored it — including the
    agents whose tables are properties (GlobalRoute) or are derived on the device from host attributes (Shortest)."""
    import rlrouting_b200 as rb
    cls = getattr(rb, policy)
    lr = {"q": 0.1, "p": 0.001 if policy == "MaHybridQ" else 0.1}

    def make(seed):
        nw = rb.Network("6x6.net", replicas=1, rng="philox", seed=seed)
        ag = cls(nw, **akw)
        nw.agent = ag
        return nw, ag

    a, ag_a = make(11)
    a.train(300, 2.5, lr=lr)                               # learned tables / congested GlobalRoute tables
    path = tmp_path / "agent.pkl"
    ag_a.store(str(path))
    blob = pickle.load(open(path, "rb"))
    assert keys <= set(blob), (set(blob), keys)
    for k in ("Qtable", "Theta", "Trace", "confidence", "choice"):
        if k in blob:
            assert all(np.asarray(blob[k][x]).shape == (36, len(a.links[x])) for x in range(36)), k
    # a fresh agent on a fresh network of the same seed, given the stored tables, continues like the original would
    # from the same environment state
    snap = a.snapshot()
    b, ag_b = make(11)
    b.restore(snap)
    if policy == "Shortest":                               # derived device tables: wipe them so that only load() can fill them
        for name in ("next_hop", "choice", "choice_mask"):
            ag_b._dev[name].zero_()
    elif policy == "GlobalRoute":
        ag_b._dev["gr_choice_mask"].zero_()
        ag_b._dev["gr_distance"].zero_()
    else:
        for name in ag_b._dev:
            ag_b._dev[name].zero_()
    ag_b.load(str(path))
    for name in ag_a._dev:
        if policy == "GlobalRoute" and name == "gr_qs_off":
            continue
        assert np.array_equal(ag_a._dev[name].cpu().numpy(), ag_b._dev[name].cpu().numpy()), name
    ra, rb_ = a.train(120, 2.5, lr=lr), b.train(120, 2.5, lr=lr)
    assert np.array_equal(ra["route_time"], rb_["route_time"])
    ca, cb = a.counters(), b.counters()
    assert all(np.array_equal(ca[k], cb[k]) for k in ca)


def test_prefetched_schedule_is_keyed_by_timing_and_seed():
    """train() prefetches the injection schedule of the call it expects next; a following call with another `freq`
    (same clock, step count and loads) must not consume it (the cache key lacked freq / slot / seed)."""
    import rlrouting_b200 as rb
    from oracle import oracle as O
    R, seed = 6, 9
    nw = rb.Network("6x6.net", replicas=R, rng="philox", seed=seed)
    agent = rb.Qroute(nw)
    nw.agent = agent
    q0 = agent.table_array("Qtable").copy()
    nw.train(100, 2.0)
    nw.train(34, 2.0, freq=2)                   # clock 100, 68 device steps: the same key as a repeat of ... (68, freq=1)
    c = nw.counters()
    for r in range(R):
        e = O.OracleEnv(nw.topology.links, "Qroute", seed=seed, replica=r)
        e.set_qtable(q0[r])
        e.run(100, lambd=2.0)
        e.run(34, lambd=2.0, freq=2)
        oc = e.counters()
        assert all(int(c[k][r]) == oc[k] for k in ("all_packets", "end_packets", "hops", "route_time", "sends")), r
    with pytest.raises(NotImplementedError, match="inject"):
        nw.inject([[rb.Packet(0, 5, nw.clock)] for _ in range(R)])
        nw.train(10, 2.0)


def test_node_views_of_any_replica():
    """Network.nodes (env.py:244-246) shows replica 0; nodes_of(r) the same read-only views for replica r."""
    import rlrouting_b200 as rb
    nw = rb.Network("6x6.net", replicas=3, rng="philox", seed=4)
    nw.agent = rb.Shortest(nw)
    nw.train(60, np.array([1.0, 3.0, 5.0]))
    for r in range(3):
        nodes = nw.nodes_of(r)
        assert list(nodes) == list(nw.links)
        for x in (0, 7, 35):
            q = nodes[x].queue
            ref = nw.queue_array(x, r)
            assert [(p.dest, p.source, p.hops, p.birth) for p in q] == [tuple(int(v) for v in row[:4]) for row in ref]
    assert [(p.dest, p.source) for p in nw.nodes[7].queue] == [(p.dest, p.source) for p in nw.nodes_of(0)[7].queue]
    with pytest.raises(IndexError):
        nw.nodes_of(3)
