"""Runs and times the UNMODIFIED reference staged in oracle/_ref/ (oracle/make_ref.py) — MEASUREMENT INFRASTRUCTURE
ONLY: imported by bench.py's `cpu_baseline` leg and `--impl reference` arm, never by the product.

One episode = what the reference's train.py does for one environment: `np.random.seed(seed)`, `Network(file)`,
`agent = Policy(nw)`, `nw.train(T, load, lr=...)` (reference env.py:367-414), nothing wrapped or patched inside the
timed call.  Hops performed are counted AFTER the call from the state the reference leaves behind: hops of the ended
packets (Network.hops, env.py:330) plus `p.hops` of every packet still queued (transtime = 1: no event outlives a
step).  Episodes run in a pool of forked worker processes, one per host core.
"""
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")

# SURVEY §8(c), first row: unmodified reference, np.random.seed(1), Qroute on 6x6.net, train(10000, 2.0)
KAT = {"net": "6x6.net", "policy": "Qroute", "load": 2.0, "T": 10000, "seed": 1,
       "all_packets": 19934, "end_packets": 14930, "route_time": 31080406, "hops_performed": 344449}


def available():
    return os.path.isfile(os.path.join(REF, "MANIFEST.json"))


_mods = None


def _load():
    global _mods
    if _mods is None:
        import numpy as np
        if not hasattr(np, "int"):
            np.int = int                       # base_policy.py:19 needs it on NumPy >= 1.24 (SURVEY §8c recipe)
        sys.dont_write_bytecode = True
        if REF not in sys.path:
            sys.path.insert(0, REF)
        import env, qroute, shortest, hybrid, multi_agent  # noqa: E401
        _mods = {"Network": env.Network, "Qroute": qroute.Qroute, "Shortest": shortest.Shortest,
                 "HybridQ": hybrid.HybridQ, "MaHybridQ": multi_agent.MaHybridQ, "CQ": qroute.CQ, "CDRQ": qroute.CDRQ,
                 "GlobalRoute": shortest.GlobalRoute}
    return _mods


def episode(job):
    """job = (net, policy, load, T, seed, lr or None) -> counters of the finished run and the seconds `train` took."""
    import numpy as np
    net, policy, load, T, seed, lr = job
    M = _load()
    np.random.seed(seed)
    nw = M["Network"](os.path.join(REF, net))
    nw.agent = M[policy](nw)
    t0 = time.perf_counter()
    nw.train(T, load, lr=lr) if lr else nw.train(T, load)
    dt = time.perf_counter() - t0
    queued = sum(p.hops for node in nw.nodes.values() for p in node.queue)
    return {"all_packets": int(nw.all_packets), "end_packets": int(nw.end_packets), "route_time": int(nw.route_time),
            "hops_performed": int(nw.hops) + int(queued), "seconds": dt,
            "ave_route_time": float(nw.ave_route_time)}


class Pool:
    """`procs` forked workers, each importing the reference once."""

    def __init__(self, procs=None):
        self.procs = procs or os.cpu_count() or 1
        self.pool = mp.get_context("fork").Pool(self.procs, initializer=_load)

    def run(self, jobs):
        """Runs the episodes concurrently; returns (results, wall seconds of the whole batch)."""
        t0 = time.perf_counter()
        res = self.pool.map(episode, jobs, chunksize=1)
        return res, time.perf_counter() - t0

    def close(self):
        self.pool.terminate()
        self.pool.join()


def check_kat(pool=None):
    """Runs the known-answer row of SURVEY §8(c) through the staged reference and asserts its four integers.
    Returns the episode record (its `seconds` is the single-core time of the reference's own 10k-step run)."""
    job = (KAT["net"], KAT["policy"], KAT["load"], KAT["T"], KAT["seed"], None)
    rec = pool.run([job])[0][0] if pool is not None else episode(job)
    for k in ("all_packets", "end_packets", "route_time", "hops_performed"):
        if rec[k] != KAT[k]:
            raise AssertionError(f"staged reference does not reproduce the known answer: {k} = {rec[k]}, expected {KAT[k]}")
    return rec
