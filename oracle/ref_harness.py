"""Drives the UNMODIFIED reference (/root/reference) in-process — TEST INFRASTRUCTURE ONLY.

Only usable in the build container (the GPU box has no /root/reference).  Used by
tests/golden/gen_golden.py to produce the committed fixtures and by the `reference`-marked
CPU tests that compare the C oracle with the live reference.

Recipe: `np.int = int` shim (base_policy.py:19 needs it on NumPy >= 1.24),
sys.path to the reference, absolute .net paths, never import train.py.
"""
import hashlib
import os
import sys

import numpy as np

REF = "/root/reference"


def available():
    return os.path.isdir(REF)


def load():
    if not hasattr(np, "int"):
        np.int = int
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import env, qroute, shortest, hybrid, multi_agent  # noqa: E401
    return {"env": env, "Network": env.Network, "Qroute": qroute.Qroute,
            "Shortest": shortest.Shortest, "HybridQ": hybrid.HybridQ,
            "MaHybridQ": multi_agent.MaHybridQ, "CQ": qroute.CQ, "CDRQ": qroute.CDRQ, "GlobalRoute": shortest.GlobalRoute}


def sendlog_hash(log):
    """sha256 over little-endian int32 rows (clock_after_step, x, y, dest) — SURVEY §8c."""
    return hashlib.sha256(np.ascontiguousarray(log, dtype="<i4").tobytes()).hexdigest()


def pack_tables(tabs, N):
    return np.concatenate([np.asarray(tabs[x], dtype=np.float64).reshape(-1) for x in range(N)])


def _det_math():
    """np.exp / np.log2 replacements backed by the oracle's deterministic exp/log2."""
    import ctypes as C
    from . import oracle as O
    L = O.lib()
    L.orc_exp.restype = C.c_double
    L.orc_exp.argtypes = [C.c_double]
    L.orc_log2.restype = C.c_double
    L.orc_log2.argtypes = [C.c_double]

    def wrap(f):
        return lambda a: np.array([f(float(v)) for v in np.asarray(a, dtype=np.float64).reshape(-1)],
                                  dtype=np.float64).reshape(np.shape(a))
    return wrap(L.orc_exp), wrap(L.orc_log2)


def run(net, policy, T, lambd, seed=1, lr=None, rng="trace", philox_seed=0, replica=0,
        agent_kwargs=None, keep_sendlog=False, det_math=False, net_kwargs=None, sample_size=0, slot=1, freq=1):
    """Runs `Network(net, **net_kwargs); agent = policy(nw); nw.train(T * slot, lambd, slot, freq, lr=lr)` step by
    step (T = number of iterations of train's loop, env.py:400).

    rng='trace'  : np.random.seed(seed) and the reference's own legacy stream (bit-for-bit the
                   reference); injections are recorded by wrapping the instance's new_packet.
    rng='philox' : np.random.seed(seed) still feeds the constructor's normals, but new_packet is
                   replaced by the Poisson-splitting Philox law and np.random.choice by the
                   Philox-uniform inverse-CDF draw (legacy RandomState.choice algorithm).
    """
    from . import philox_twin
    R = load()
    path = net if os.path.isabs(net) else os.path.join(REF, net)
    np.random.seed(seed)
    nw = R["Network"](path, **(net_kwargs or {}))
    agent = R[policy](nw, **(agent_kwargs or {}))
    nw.agent = agent
    N = len(nw.nodes)
    out = {"N": N}
    if hasattr(agent, "Qtable"):
        out["Q0"] = pack_tables(agent.Qtable, N)
    inj_cnt, inj_pairs, log = [], [], []
    orig_new_packet = nw.new_packet
    Packet = R["env"].Packet
    state = {"node": None}

    if rng == "trace":
        def new_packet(l):
            ps = orig_new_packet(l)
            inj_cnt.append(len(ps))
            inj_pairs.extend((p.source, p.dest) for p in ps)
            return ps
    else:
        def new_packet(l):
            enlam = np.exp(-np.float64(l) / np.float64(N))
            pairs = philox_twin.new_packets(philox_seed, replica, nw.clock, N, enlam)
            inj_cnt.append(len(pairs))
            inj_pairs.extend(pairs)
            return [Packet(s, d, nw.clock) for s, d in pairs]
        orig_choose = agent.choose

        def choose(source, dest, *a, **k):
            state["node"] = source
            return orig_choose(source, dest, *a, **k)
        agent.choose = choose
        orig_choice = np.random.choice

        streams = {}

        def stream():
            # node x's sampling stream of this step (purpose 1): a queue scan that calls choose more than once
            # (env.py:174-190, links busy) keeps drawing from it
            key = (nw.clock, state["node"])
            if key not in streams:
                streams.clear()
                streams[key] = philox_twin.Stream(philox_seed, replica, nw.clock, state["node"], 1)
            return streams[key]

        def choice(a, size=None, replace=True, p=None):
            # legacy RandomState.choice with p (numpy/random/mtrand.pyx): validate p, then
            # cdf = p.cumsum(); cdf /= cdf[-1]; idx = cdf.searchsorted(u, side='right')
            if p is None:
                # legacy RandomState.choice(a) without p = a[randint(0, len(a))]; here the index comes from the Philox
                # stream as mulhi(word, count) (Shortest(random=True), shortest.py:40-41); one choice draws nothing
                if len(a) == 1:
                    return a[0]
                return a[(stream().next() * len(a)) >> 32]
            p = np.asarray(p, dtype=np.float64)
            if np.isnan(p.sum()):
                raise ValueError("probabilities contain NaN")
            cdf = p.cumsum()
            cdf /= cdf[-1]
            st = stream()
            hi, lo = st.next() >> 5, st.next() >> 6
            u = (hi * 67108864.0 + lo) / 9007199254740992.0
            return a[int(cdf.searchsorted(u, side="right"))]
        np.random.choice = choice
    nw.new_packet = new_packet
    # det_math: hybrid.py:17,39 look np.exp / np.log2 up at call time, so swapping the module
    # attributes makes the unmodified reference use the oracle's deterministic exp/log2 and
    # therefore bit-comparable with it (see rlr_oracle.c "Deterministic exp / log2").
    orig_exp, orig_log2 = np.exp, np.log2
    if det_math:
        np.exp, np.log2 = _det_math()

    rt = np.zeros(T, dtype=np.int64)
    en = np.zeros(T, dtype=np.int64)
    hp = np.zeros(T, dtype=np.int64)
    sends = 0
    err = None
    sample = None
    try:
        if sample_size:                          # Network.sample_route_time (env.py:416-438) instead of the train loop
            sample = nw.sample_route_time(sample_size, lambd, slot=slot, freq=freq, lr=lr or {})
            T = 0
        for i in range(T):                       # Network.train's loop, env.py:400-413
            nw.inject(nw.new_packet(lambd * slot))
            for _ in range(freq):
                r = nw.step(slot)
                for rw in r:
                    log.append((nw.clock, rw.source, rw.action, rw.dest))
                sends += len(r)
                if lr:
                    agent.learn(r, lr=lr)
                else:
                    agent.learn(r)
            rt[i], en[i], hp[i] = nw.route_time, nw.end_packets, nw.hops
    except ValueError as e:                      # F7: NaN probabilities
        err = (str(e), i)
    finally:
        np.exp, np.log2 = orig_exp, orig_log2
        if rng != "trace":
            np.random.choice = orig_choice
    log = np.array(log, dtype=np.int32).reshape(-1, 4)
    out.update({
        "counters": {"clock": int(nw.clock), "all_packets": int(nw.all_packets),
                     "end_packets": int(nw.end_packets), "drop_packets": int(nw.drop_packets),
                     "active_packets": int(nw.active_packets), "hops": int(nw.hops),
                     "route_time": int(nw.route_time), "sends": int(sends)},
        "ave_route_time": float(nw.ave_route_time),
        "route_time": rt, "end_packets": en, "hops": hp,
        "sendlog_sha256": sendlog_hash(log), "n_sends": len(log),
        "inj_cnt": np.array(inj_cnt, dtype=np.int32),
        "inj_pairs": np.array(inj_pairs, dtype=np.int32).reshape(-1, 2),
        "queues": [np.array([(p.dest, p.source, p.hops, p.birth, p.start_queue)
                             for p in nw.nodes[x].queue], dtype=np.int32).reshape(-1, 5)
                   for x in range(N)],
        "error": err,
        "in_flight": len(nw.event_queue),
    })
    if keep_sendlog:
        out["sendlog"] = log
    if sample is not None:
        out["sample"] = np.asarray(sample, dtype=np.int64)
    for name in ("Qtable", "Theta", "Trace", "confidence"):
        if hasattr(agent, name):
            out[name] = pack_tables(getattr(agent, name), N)
    if policy in ("Shortest", "GlobalRoute"):
        out["distance"] = agent.distance.copy()
        out["choice"] = np.concatenate([agent.choice[x].reshape(-1) for x in range(N)]).astype(np.uint8)
    if policy == "GlobalRoute":
        out["queue_size"] = np.asarray(agent.queue_size, dtype=np.int64).copy()
    return out
