"""Race evidence without compute-sanitizer (closed on this GPU pool, profiles/r2_compute_sanitizer_closed.txt): the kernels
whose phases share state across threads — CDRQ's backward updates onto a neighbour's table, MaHybridQ's helper-warp
reductions, the general timing model's parallel heap append — must give the same bits for every launch geometry
(replicas per CTA, CTA size, tables in shared memory or HBM) and on every repetition.  A data race shows up as a
geometry- or run-dependent table; the values themselves are pinned to the oracle by tests/test_gpu_parity.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    ("CDRQ", {}, [(0, 0, -1), (1, 128, 1), (3, 256, 1), (2, 160, 0), (7, 512, 0)]),
    ("CQ", {}, [(0, 0, -1), (2, 192, 1), (5, 256, 0)]),
    ("MaHybridQ", {}, [(0, 0, -1), (1, 256, 1), (2, 256, 1), (1, 512, 0)]),
    ("HybridQ", {}, [(0, 0, -1), (1, 64, 1), (7, 256, 0), (3, 128, 1)]),
    ("Qroute", {"transtime": 3, "bandwidth": 2}, [(0, 0, -1), (1, 64, 1), (4, 160, 0), (7, 256, 0)]),
    ("CDRQ", {"transtime": 2}, [(0, 0, -1), (1, 128, 1), (5, 256, 0)]),
    ("GlobalRoute", {}, [(0, 0, -1), (1, 64, -1), (3, 128, -1)]),
]


@pytest.mark.parametrize("policy,nkw,geoms", CASES)
def test_same_bits_for_every_launch_geometry_and_repetition(policy, nkw, geoms):
    import rlrouting_b200 as rb
    R, T, lam = 23, 180, np.linspace(0.5, 4.0, 23)
    lr = {"q": 0.1, "p": 0.001 if policy == "MaHybridQ" else 0.1, "e": 0.1}
    ref = None
    for rpb, tpb, tis in geoms:
        for rep in range(2):
            nw = rb.Network("6x6.net", replicas=R, rng="philox", seed=13, replicas_per_block=rpb, threads_per_block=tpb,
                            tables_in_smem=tis, **nkw)
            agent = getattr(rb, policy)(nw)
            nw.agent = agent
            out = nw.train(T, lam, lr=lr, steps_per_launch=61)
            got = {"rt": out["route_time"].copy(), **{k: v for k, v in nw.counters().items()},
                   **{n: t.cpu().numpy().view(np.int64).copy() for n, t in agent._dev.items()},
                   "queues": np.concatenate([nw.queue_array(x, r) for r in (0, R - 1) for x in range(36)])}
            if ref is None:
                ref = got
                continue
            for k in ref:
                assert np.array_equal(ref[k], got[k]), (policy, (rpb, tpb, tis), rep, k)
