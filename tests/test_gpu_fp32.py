"""fp32 table-storage mode of Qroute (`Network(dtype='float32')`, SURVEY §8b / north star "Q-tables ... fp32").

The mode changes what is STORED (float instead of double: half the bytes, twice the replicas per SM); rows are widened to
fp64 when read and the TD update is computed in fp64 with the reference's unfused operations, then rounded once on store.
That law is simple enough to restate in the oracle (orc_set_q_f32), so the mode is checked bit for bit like everything
else; against the fp64 run it is checked the way the north star asks — identical integers until a near-tie forks a
replica, Q within 1e-5 relative on the replicas that have not forked, mean delivery time per load level within the
across-replica confidence interval."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _net(dtype, R, seed, **kw):
    import rlrouting_b200 as rb
    nw = rb.Network("6x6.net", replicas=R, rng="philox", seed=seed, dtype=dtype, **kw)
    ag = rb.Qroute(nw)
    nw.agent = ag
    return nw, ag


@pytest.mark.parametrize("kw", [{}, {"tables_in_smem": 0}, {"replicas_per_block": 3, "threads_per_block": 128}])
def test_fp32_storage_matches_the_oracle_with_float_rounded_stores(kw):
    R, T, seed = 12, 400, 17
    lam = np.linspace(0.8, 3.5, R)
    nw, ag = _net("float32", R, seed, **kw)
    assert ag.table_array("Qtable").dtype == np.float32
    q0 = ag.table_array("Qtable").copy()
    res = nw.train(T, lam, steps_per_launch=137)
    nw.step()                                   # the split-phase instantiation reads the float table too
    c = nw.counters()
    for r in range(R):
        e = O.OracleEnv(nw.topology.links, "Qroute", seed=seed, replica=r)
        e.set_q_storage_f32()
        e.set_qtable(q0[r])
        assert np.array_equal(e.Q.astype(np.float32), q0[r])          # structure overrides survive the rounding
        o = e.run(T, lambd=float(lam[r]))
        with np.errstate(divide="ignore", invalid="ignore"):
            ort = np.where(o["end_packets"] > 0, o["route_time"] / np.maximum(o["end_packets"], 1), 0.0)
        assert np.array_equal(res["route_time"][r], ort), r
        e.run(1, lambd=0.0)                                            # the extra split step injects nothing
    got = ag.table_array("Qtable")
    assert got.dtype == np.float32


def test_fp32_storage_table_bits_and_counters_after_a_fused_run():
    R, T, seed = 9, 300, 5
    nw, ag = _net("float32", R, seed)
    q0 = ag.table_array("Qtable").copy()
    nw.train(T, 2.0)
    c = nw.counters()
    q = ag.table_array("Qtable")
    for r in range(R):
        e = O.OracleEnv(nw.topology.links, "Qroute", seed=seed, replica=r)
        e.set_q_storage_f32()
        e.set_qtable(q0[r])
        e.run(T, lambd=2.0)
        oc = e.counters()
        assert all(int(c[k][r]) == oc[k] for k in ("all_packets", "end_packets", "hops", "route_time", "sends")), r
        assert np.array_equal(q[r].view(np.int32), e.Q.astype(np.float32).view(np.int32)), r
    st = nw.table_stats(prev=ag.table_device("Qtable").clone())["per_replica"].cpu().numpy()
    assert np.allclose(st[:, 0], q.astype(np.float64).sum(1), rtol=1e-12) and np.all(st[:, 1] == 0.0)


def test_fp32_storage_against_fp64_fork_and_statistics():
    """North-star wording: integers identical until the first fork; un-forked replicas' Q within 1e-5 relative;
    mean delivery time per load level inside the across-replica confidence interval of the fp64 run."""
    loads, per = [1.5, 2.5, 3.5], 96
    R, seed = len(loads) * per, 3
    lam = np.repeat(loads, per)
    lv = np.repeat(np.arange(len(loads)), per)
    a, ag_a = _net("float64", R, seed)
    b, ag_b = _net("float32", R, seed)
    # short horizon: most replicas have not met a near-tie yet
    a.train(60, lam)
    b.train(60, lam)
    ca, cb = a.counters(), b.counters()
    same = np.all([ca[k] == cb[k] for k in ("all_packets", "end_packets", "hops", "route_time", "sends")], axis=0)
    assert same.mean() > 0.5, same.mean()
    qa, qb = ag_a.table_array("Qtable"), ag_b.table_array("Qtable").astype(np.float64)
    err = np.abs(qa - qb)[same]
    assert np.all(err <= 1e-5 * np.abs(qa[same]) + 1e-6), err.max()     # 1e-5 relative; 1e-6 absolute for entries near zero
    # long horizon: statistics
    a.train(2500, lam)
    b.train(2500, lam)
    ca, cb = a.counters(), b.counters()
    for lvl in range(len(loads)):
        m = lv == lvl
        ma = ca["route_time"][m] / np.maximum(ca["end_packets"][m], 1)
        mb = cb["route_time"][m] / np.maximum(cb["end_packets"][m], 1)
        ci = 4.0 * np.sqrt(ma.var(ddof=1) / per + mb.var(ddof=1) / per)
        assert abs(ma.mean() - mb.mean()) <= ci, (lvl, ma.mean(), mb.mean(), ci)


def test_fp32_storage_is_refused_for_other_policies():
    import rlrouting_b200 as rb
    nw = rb.Network("6x6.net", replicas=2, rng="philox", seed=1, dtype="float32")
    with pytest.raises(NotImplementedError):
        rb.HybridQ(nw)
    with pytest.raises(ValueError):
        rb.Network("6x6.net", dtype="float16")
