"""Paged queues (`Network(page_records=P, pages_per_replica=NP)`): the reference's unbounded per-node lists (env.py:100,130)
as linked pages from a per-replica pool instead of one ring per node sized for the worst node of the run (SURVEY §7 hard
part 2, option ii).  Same results as the rings and as the oracle, bit for bit; memory follows the packets alive."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

LR = {"q": 0.1, "p": 0.1, "e": 0.1}


def _make(net, policy, R, seed, **kw):
    import rlrouting_b200 as rb
    nw = rb.Network(net, replicas=R, rng="philox", seed=seed, **kw)
    ag = getattr(rb, policy)(nw)
    nw.agent = ag
    return nw, ag


@pytest.mark.parametrize("net,policy,P,kw", [
    ("6x6.net", "Qroute", 4, {}),
    ("6x6.net", "Qroute", 16, {"replicas_per_block": 3, "threads_per_block": 128}),
    ("6x6.net", "Qroute", 8, {"tables_in_smem": 0}),
    ("6x6.net", "Shortest", 4, {}),
    ("6x6.net", "HybridQ", 8, {}),
    ("6x6.net", "MaHybridQ", 4, {}),
    ("6x6.net", "CQ", 16, {}),
    ("6x6.net", "CDRQ", 4, {}),
    ("6x6.net", "GlobalRoute", 8, {}),
    ("6x6-modified.net", "Qroute", 4, {}),
    ("lata.net", "Qroute", 8, {}),
    ("lata.net", "HybridQ", 4, {}),
])
def test_paged_queues_match_the_oracle(net, policy, P, kw):
    R, seed = 7, 31
    T1, T2 = (150, 110) if net == "lata.net" or policy == "GlobalRoute" else (420, 260)
    lam = np.linspace(0.7, 5.0, R)                      # the high loads inject bursts and build long queues
    lr = dict(LR, p=0.001) if policy == "MaHybridQ" else LR
    nw, ag = _make(net, policy, R, seed, page_records=P, pages_per_replica=4000, **kw)
    q0 = ag.table_array("Qtable").copy() if "Qtable" in getattr(ag, "_table_names", ()) else None
    parts = [nw.train(T1, lam, lr=lr, steps_per_launch=97)["route_time"], nw.train(T2, lam, lr=lr)["route_time"]]
    got = np.concatenate(parts, axis=1)
    c = nw.counters()
    N = nw.topology.N
    for r in range(R):
        e = O.OracleEnv(nw.topology.links, policy, seed=seed, replica=r)
        if q0 is not None:
            e.set_qtable(q0[r])
        o = e.run(T1 + T2, lambd=float(lam[r]), lr_q=lr["q"], lr_p=lr["p"], lr_e=lr["e"])
        oc = e.counters()
        assert all(int(c[k][r]) == oc[k] for k in ("all_packets", "end_packets", "hops", "route_time", "sends")), (r, oc)
        with np.errstate(divide="ignore", invalid="ignore"):
            ort = np.where(o["end_packets"] > 0, o["route_time"] / np.maximum(o["end_packets"], 1), 0.0)
        assert np.array_equal(got[r], ort), r
        if q0 is not None:
            assert np.array_equal(ag.table_array("Qtable")[r].view(np.int64), e.Q.view(np.int64)), r
        for x in range(N):
            assert np.array_equal(nw.queue_array(x, r), e.queue(x)), (r, x)
    # every page is accounted for: on the free stack, a spare, or in exactly one queue
    top = nw._page_top.cpu().numpy()
    lens = (nw._q_tail.cpu().numpy().astype(np.int64) - nw._q_head.cpu().numpy().astype(np.int64)) & 0xFFFFFFFF
    heads = nw._q_head.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    in_queues = ((heads % P + lens) // P + 1).sum(axis=1)
    assert np.array_equal(top + N + in_queues, np.full(R, nw.pages_per_replica))


def test_paged_and_ring_runs_agree_and_snapshots_move_between_them():
    R, T, seed = 40, 500, 2
    lam = np.linspace(1.0, 4.0, R)
    a, ag_a = _make("6x6.net", "Qroute", R, seed)
    b, ag_b = _make("6x6.net", "Qroute", R, seed, page_records=16, pages_per_replica=1500)
    ra, rb_ = a.train(T, lam), b.train(T, lam)
    assert np.array_equal(ra["route_time"], rb_["route_time"])
    assert np.array_equal(ag_a.table_array("Qtable").view(np.int64), ag_b.table_array("Qtable").view(np.int64))
    # paged -> rings and rings -> paged through a snapshot, then both continue like the original
    c, ag_c = _make("6x6.net", "Qroute", R, seed)
    c.restore(b.snapshot())
    d, ag_d = _make("6x6.net", "Qroute", R, seed, page_records=8, pages_per_replica=3000)
    d.restore(a.snapshot())
    outs = [n.train(300, lam)["route_time"] for n in (a, b, c, d)]
    cs = [n.counters() for n in (a, b, c, d)]
    for o, cc in zip(outs[1:], cs[1:]):
        assert np.array_equal(outs[0], o)
        assert all(np.array_equal(cs[0][k], cc[k]) for k in cs[0])
    for x in (0, 17, 35):
        assert np.array_equal(a.queue_array(x, 5), d.queue_array(x, 5))


def test_paged_memory_follows_live_packets_and_an_empty_pool_is_reported():
    import rlrouting_b200 as rb
    R, T = 64, 3000
    nw, ag = _make("6x6.net", "Qroute", R, 1, page_records=16, pages_per_replica=700)
    nw.train(T, 2.0)
    c = nw.counters()
    live = int(c["active_packets"].max())
    lens = (nw._q_tail.cpu().numpy().astype(np.int64) - nw._q_head.cpu().numpy().astype(np.int64)) & 0xFFFFFFFF
    worst = int(lens.max())
    ring_cap = 1 << int(np.ceil(np.log2(worst + 1)))
    ring_bytes = R * 36 * ring_cap * 16
    assert nw.queue_memory_bytes < 0.7 * ring_bytes, (nw.queue_memory_bytes, ring_bytes, worst, live)
    assert nw.queue_memory_bytes / R <= 3.2 * 16 * live + 2 * 36 * 16 * 16, (nw.queue_memory_bytes / R, live)
    small, _ = _make("6x6.net", "Qroute", 4, 1, page_records=16, pages_per_replica=120)
    with pytest.raises(OverflowError):
        small.train(T, 2.0)
    with pytest.raises(NotImplementedError):
        nw.step()
    with pytest.raises(NotImplementedError):
        rb.Network("6x6.net", transtime=2, page_records=16)
