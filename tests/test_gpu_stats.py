"""Reduced statistics that leave the device instead of the per-replica vectors (north star: per-load-level delivery-time
and Q-convergence statistics): `Network.train(level_index=...)` (rlr_reduce_levels), `Network.table_stats`
(rlr_table_stats) and `dist.global_stats` against the CPU oracle, replica by replica."""
import math

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

LOADS = [1.0, 2.5, 4.0]
PER_LEVEL = 5


def _oracle_runs(topo, policy, seed, lam, T, q0=None, lr=None):
    lr = lr or {"q": 0.1, "p": 0.1, "e": 0.1}
    envs, outs = [], []
    for r, l in enumerate(lam):
        e = O.OracleEnv(topo.links, policy, seed=seed, replica=r)
        if q0 is not None:
            e.set_qtable(q0[r])
        outs.append(e.run(T, lambd=float(l), lr_q=lr["q"], lr_p=lr["p"], lr_e=lr["e"]))
        envs.append(e)
    return envs, outs


@pytest.mark.parametrize("policy,freq", [("Qroute", 1), ("Shortest", 1), ("Qroute", 2)])
def test_level_reduced_train_matches_pooled_oracle_counters(policy, freq):
    import rlrouting_b200 as rb
    R, T, seed = len(LOADS) * PER_LEVEL, 260, 21
    lam = np.repeat(LOADS, PER_LEVEL)
    lv = np.repeat(np.arange(len(LOADS)), PER_LEVEL)
    nw = rb.Network("6x6.net", replicas=R, rng="philox", seed=seed)
    agent = getattr(rb, policy)(nw)
    nw.agent = agent
    q0 = agent.table_array("Qtable").copy() if policy == "Qroute" else None
    res = nw.train(T, lam, freq=freq, hop=True, level_index=lv, steps_per_launch=97)
    assert res["route_time"].shape == (len(LOADS), T) and res["hop"].shape == (len(LOADS), T)
    assert nw.last_d2h_bytes < R * T * 8                   # the per-replica vectors did not cross the bus
    topo = nw.topology
    for lvl in range(len(LOADS)):
        rt = np.zeros(T, dtype=np.int64)
        en = np.zeros(T, dtype=np.int64)
        hp = np.zeros(T, dtype=np.int64)
        for r in np.nonzero(lv == lvl)[0]:
            e = O.OracleEnv(topo.links, policy, seed=seed, replica=int(r))
            if q0 is not None:
                e.set_qtable(q0[r])
            o = e.run(T, lambd=float(lam[r]), freq=freq)
            rt += o["route_time"]                            # cumulative counters after each iteration of train's loop
            en += o["end_packets"]
            hp += o["hops"]
        with np.errstate(divide="ignore", invalid="ignore"):
            want_rt = np.where(en > 0, rt / np.maximum(en, 1), 0.0)
            want_hp = np.where(en > 0, hp / np.maximum(en, 1), 0.0)
        assert np.array_equal(res["route_time"][lvl], want_rt), lvl
        assert np.array_equal(res["hop"][lvl], want_hp), lvl


def test_level_reduced_and_per_replica_vectors_describe_the_same_run():
    import rlrouting_b200 as rb
    R, T = 8, 150
    lv = np.array([0, 0, 0, 1, 1, 2, 2, 2])
    lam = np.array([1.0, 1.0, 1.0, 2.0, 2.0, 3.0, 3.0, 3.0])

    def run(**kw):
        nw = rb.Network("6x6.net", replicas=R, rng="philox", seed=4)
        nw.agent = rb.Qroute(nw)
        return nw, nw.train(T, lam, **kw)

    a, ra = run()
    b, rb_ = run(level_index=lv)
    ca, cb = a.counters(), b.counters()
    assert all(np.array_equal(ca[k], cb[k]) for k in ca)
    # one replica per level reduces to the replica's own vector
    c, rc = run(level_index=np.arange(R))
    assert np.array_equal(rc["route_time"], ra["route_time"])
    # the final pooled mean is the counters' ratio
    for lvl in range(3):
        m = lv == lvl
        assert rb_["route_time"][lvl, -1] == ca["route_time"][m].sum() / ca["end_packets"][m].sum()


def test_table_stats_and_global_stats_match_the_oracle_tables():
    import rlrouting_b200 as rb
    from rlrouting_b200 import dist as rdist
    R, T, seed = len(LOADS) * PER_LEVEL, 200, 8
    lam = np.repeat(LOADS, PER_LEVEL)
    lv = np.repeat(np.arange(len(LOADS)), PER_LEVEL)
    nw = rb.Network("6x6.net", replicas=R, rng="philox", seed=seed)
    agent = rb.Qroute(nw)
    nw.agent = agent
    q0 = agent.table_array("Qtable").copy()
    prev = agent.table_device("Qtable").clone()
    nw.train(T, lam)
    st = nw.table_stats(prev=prev, level_index=lv)
    per, lev = st["per_replica"].cpu().numpy(), st["levels"].cpu().numpy()
    envs, _ = _oracle_runs(nw.topology, "Qroute", seed, lam, T, q0)
    for r, e in enumerate(envs):
        assert np.array_equal(agent.table_array("Qtable")[r].view(np.int64), e.Q.view(np.int64))
        assert math.isclose(per[r, 0], math.fsum(e.Q), rel_tol=1e-12, abs_tol=1e-9)
        assert math.isclose(per[r, 1], math.fsum(np.abs(e.Q - q0[r])), rel_tol=1e-12, abs_tol=1e-9)
    for lvl in range(len(LOADS)):
        m = lv == lvl
        assert math.isclose(lev[lvl, 0], math.fsum(per[m, 0]), rel_tol=1e-13, abs_tol=1e-9)
        assert math.isclose(lev[lvl, 1], math.fsum(per[m, 1]), rel_tol=1e-13)
    assert np.array_equal(nw.table_stats(prev=prev, level_index=lv)["levels"].cpu().numpy(), lev)    # reproducible bits
    g = rdist.global_stats(nw, level_index=lv, n_levels=len(LOADS), q_prev=prev)
    c = nw.counters()
    for lvl in range(len(LOADS)):
        m = lv == lvl
        assert g["levels"]["end_packets"][lvl] == c["end_packets"][m].sum()
        assert g["levels"]["ave_route_time"][lvl] == c["route_time"][m].sum() / c["end_packets"][m].sum()
        assert g["levels"]["q_abs_delta"][lvl] == lev[lvl, 1]
    assert g["sends"] == c["sends"].sum() and g["q_abs_delta"] > 0
    with pytest.raises(ValueError):
        nw.table_stats(table="Theta")
