"""Shared helpers for the parity tests: rebuild a golden case's inputs and run the oracle."""
import hashlib
import json
import os

import numpy as np

from oracle import oracle as O
from rlrouting_b200 import nets, trace
from rlrouting_b200.topology import Topology

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LR_DEFAULT = {"q": 0.1, "p": 0.1, "e": 0.1}


def sha(a, dt):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=dt).tobytes()).hexdigest()


def load_cases():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


def topo_of(net):
    return Topology.from_text(nets.net_text(net))


def case_inputs(case, topo):
    """(Q0 flat or None, (cnt, pairs) or None) exactly as the reference's RNG calls produce them."""
    rs = np.random.RandomState(case["seed"])
    kw = case.get("agent_kwargs") or {}
    q0 = None
    if case["policy"] not in ("Shortest", "GlobalRoute"):
        times = 2 if case["policy"] in ("HybridQ", "MaHybridQ") else 1
        q0 = topo.pack(trace.draw_qtable_normals(rs, topo, kw.get("initQ", 0), times))
    inj = None
    if case.get("rng", "trace") == "trace":
        inj = trace.draw_injections(rs, topo.N, case["T"] or 4000, case["lambd"] * case.get("slot", 1))   # sample cases stop on their own
    return q0, inj


def lr_of(case):
    lr = dict(LR_DEFAULT)
    lr.update(case.get("lr") or {})
    return lr


def run_oracle(case, sendlog=True):
    topo = topo_of(case["net"])
    kw = case.get("agent_kwargs") or {}
    env = O.OracleEnv(topo.links, case["policy"], discount=kw.get("discount"),
                      add_entropy=kw.get("add_entropy", True), decay=kw.get("decay", 0.9),
                      is_drop=(case.get("net_kwargs") or {}).get("is_drop", False),
                      bandwidth=(case.get("net_kwargs") or {}).get("bandwidth", 1),
                      transtime=(case.get("net_kwargs") or {}).get("transtime", 1),
                      multiway=kw.get("multiway", False), random=kw.get("random", False),
                      seed=case.get("philox_seed", 0), replica=case.get("replica", 0))
    q0, inj = case_inputs(case, topo)
    if q0 is not None:
        env.set_qtable(q0, initP=kw.get("initP", 0.0))
    lr = lr_of(case)
    if case.get("sample_size"):
        smp, steps = O.sample_route_time(env, case["sample_size"], lambd=case["lambd"], inj=inj, lr_q=lr["q"], lr_p=lr["p"],
                                         lr_e=lr["e"], slot=case.get("slot", 1), freq=case.get("freq", 1))
        return topo, env, q0, inj, {"sample": smp, "steps_done": steps}
    res = env.run(case["T"], lambd=case["lambd"], inj=inj, lr_q=lr["q"], lr_p=lr["p"], lr_e=lr["e"],
                  sendlog=sendlog, slot=case.get("slot", 1), freq=case.get("freq", 1))
    return topo, env, q0, inj, res
