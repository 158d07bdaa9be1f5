"""Generates tests/golden/kat.json and tests/golden/np_tables.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python tests/golden/gen_golden.py
The reference has no tests or golden vectors of its own; these known answers are the
pin for oracle/rlr_oracle.c (tests/test_oracle_golden.py) and, through it, for the CUDA path.

Inputs are reproducible from the case description alone: NumPy's legacy RandomState stream is
frozen, so `seed` regenerates the constructor's normals and (trace mode) the injections; Philox
mode regenerates injections and policy samples from (philox_seed, replica).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def h(a, dt):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=dt).tobytes()).hexdigest()


CASES = [
    # trace mode: bit-for-bit the reference's own RNG stream
    dict(net="6x6.net", policy="Qroute", T=10000, lambd=2.0, seed=1),          # BASELINE config 1
    dict(net="6x6-modified.net", policy="Qroute", T=3000, lambd=2.0, seed=1),
    dict(net="lata.net", policy="Qroute", T=2000, lambd=2.0, seed=1),
    dict(net="6x6.net", policy="Qroute", T=2000, lambd=0.5, seed=1),
    dict(net="6x6.net", policy="Qroute", T=2000, lambd=1.0, seed=5, lr={"q": 0.5}),
    dict(net="6x6.net", policy="Qroute", T=1500, lambd=3.0, seed=2, agent_kwargs={"discount": 0.9, "initQ": -2}),
    dict(net="6x6.net", policy="Shortest", T=10000, lambd=1.0, seed=1),
    dict(net="6x6.net", policy="Shortest", T=10000, lambd=2.0, seed=1),
    dict(net="6x6.net", policy="Shortest", T=10000, lambd=4.0, seed=1),
    dict(net="6x6-modified.net", policy="Shortest", T=2000, lambd=2.5, seed=3),
    dict(net="lata.net", policy="Shortest", T=1000, lambd=3.0, seed=1),
    # Philox mode: reference driven by the Philox twin (oracle/philox_twin.py)
    dict(net="6x6.net", policy="Qroute", T=2000, lambd=1.0, seed=1, rng="philox", philox_seed=7, replica=3),
    dict(net="6x6.net", policy="Shortest", T=2000, lambd=4.0, seed=1, rng="philox", philox_seed=7, replica=3),
    dict(net="lata.net", policy="Qroute", T=800, lambd=2.0, seed=4, rng="philox", philox_seed=0, replica=0),
    # sampled policies; det_math: the reference's np.exp/np.log2 swapped for the deterministic pair
    dict(net="6x6.net", policy="HybridQ", T=3000, lambd=2.0, seed=1, rng="philox", philox_seed=7, replica=3, det_math=True),
    dict(net="lata.net", policy="HybridQ", T=1000, lambd=2.0, seed=1, rng="philox", philox_seed=7, replica=3, det_math=True),
    dict(net="6x6.net", policy="HybridQ", T=1000, lambd=1.0, seed=9, rng="philox", philox_seed=11, replica=(1 << 33) + 5,
         det_math=True, lr={"q": 0.2, "p": 0.05, "e": 0.3}, agent_kwargs={"add_entropy": False, "discount": 0.9}),
    dict(net="6x6.net", policy="MaHybridQ", T=4000, lambd=2.0, seed=1, rng="philox", philox_seed=7, replica=3,
         det_math=True, lr={"q": 0.1, "p": 0.001}),
    dict(net="6x6.net", policy="MaHybridQ", T=3000, lambd=2.0, seed=1, rng="philox", philox_seed=7, replica=3, det_math=True),  # F7: NaN at 2526
    dict(net="6x6-modified.net", policy="MaHybridQ", T=1500, lambd=1.0, seed=1, rng="philox", philox_seed=7, replica=3,
         det_math=True, lr={"q": 0.1, "p": 0.001}),
    # confidence-based and dual Q-routing (qroute.py:56-122; CDRQ runs the network in 'dual' mode)
    dict(net="6x6.net", policy="CQ", T=3000, lambd=2.0, seed=1),
    dict(net="lata.net", policy="CQ", T=800, lambd=2.0, seed=1),
    dict(net="6x6.net", policy="CQ", T=1500, lambd=1.0, seed=1, rng="philox", philox_seed=7, replica=3,
         agent_kwargs={"decay": 0.8, "discount": 0.95}),
    dict(net="6x6.net", policy="CDRQ", T=3000, lambd=2.0, seed=1),
    dict(net="6x6-modified.net", policy="CDRQ", T=2000, lambd=1.0, seed=1),
    dict(net="lata.net", policy="CDRQ", T=800, lambd=3.0, seed=1),
    dict(net="6x6.net", policy="CDRQ", T=1500, lambd=3.0, seed=1, rng="philox", philox_seed=7, replica=3),
    # GlobalRoute — all-pairs distances recomputed every step with node weights 1 + queue size
    dict(net="6x6.net", policy="GlobalRoute", T=300, lambd=2.0, seed=1),
    dict(net="6x6-modified.net", policy="GlobalRoute", T=200, lambd=3.0, seed=1),
    dict(net="lata.net", policy="GlobalRoute", T=40, lambd=3.0, seed=1),
    dict(net="6x6.net", policy="GlobalRoute", T=200, lambd=3.0, seed=1, rng="philox", philox_seed=7, replica=3),
    # Shortest(multiway=True) keeps every optimal neighbour, random=True samples among them
    dict(net="6x6.net", policy="Shortest", T=1500, lambd=2.0, seed=1, agent_kwargs={"multiway": True}),
    dict(net="lata.net", policy="Shortest", T=500, lambd=3.0, seed=1, agent_kwargs={"multiway": True}),
    dict(net="6x6.net", policy="Shortest", T=1500, lambd=3.0, seed=1, rng="philox", philox_seed=7, replica=3,
         agent_kwargs={"multiway": True, "random": True}),
    dict(net="lata.net", policy="Shortest", T=500, lambd=3.0, seed=1, rng="philox", philox_seed=7, replica=3,
         agent_kwargs={"multiway": True, "random": True}),
    # Network.sample_route_time — run until `sample_size` deliveries, routing times in delivery order
    dict(net="6x6.net", policy="Shortest", T=0, lambd=2.0, seed=1, sample_size=500),
    dict(net="6x6.net", policy="Qroute", T=0, lambd=2.0, seed=1, sample_size=300),
    dict(net="6x6.net", policy="CDRQ", T=0, lambd=3.0, seed=1, sample_size=200),
    dict(net="lata.net", policy="Qroute", T=0, lambd=3.0, seed=1, rng="philox", philox_seed=7, replica=3, sample_size=40),
    dict(net="6x6.net", policy="MaHybridQ", T=0, lambd=2.0, seed=1, rng="philox", philox_seed=7, replica=3, det_math=True,
         lr={"q": 0.1, "p": 0.001}, sample_size=60),
    # is_drop=True — a packet arriving with hops >= N is dropped (env.py:355-360)
    dict(net="6x6.net", policy="Qroute", T=3000, lambd=2.0, seed=1, net_kwargs={"is_drop": True}),
    dict(net="6x6.net", policy="CDRQ", T=2000, lambd=3.0, seed=2, net_kwargs={"is_drop": True}),
    dict(net="lata.net", policy="HybridQ", T=600, lambd=2.0, seed=1, rng="philox", philox_seed=7, replica=3, det_math=True,
         net_kwargs={"is_drop": True}),
    # sampled policies with NumPy's own exp/log2: short horizon, tables compared with a tolerance
    dict(net="6x6.net", policy="HybridQ", T=300, lambd=2.0, seed=1, rng="philox", philox_seed=7, replica=3, keep_tables=True),
    dict(net="6x6.net", policy="MaHybridQ", T=300, lambd=2.0, seed=1, rng="philox", philox_seed=7, replica=3,
         lr={"q": 0.1, "p": 0.001}, keep_tables=True),
    # the general timing model — integer slot / freq / transtime / bandwidth.  transtime > slot
    # keeps events in flight across steps: links stay busy and the send loop scans past packets whose link is full
    dict(net="6x6.net", policy="Qroute", T=600, lambd=1.0, seed=1, slot=2),
    dict(net="6x6.net", policy="Qroute", T=600, lambd=2.0, seed=1, freq=2),
    dict(net="6x6.net", policy="Qroute", T=500, lambd=1.0, seed=2, net_kwargs={"transtime": 2, "bandwidth": 3}, slot=3, freq=2),
    dict(net="6x6.net", policy="Qroute", T=800, lambd=2.0, seed=1, net_kwargs={"transtime": 2}),
    dict(net="6x6.net", policy="Qroute", T=800, lambd=2.0, seed=1, net_kwargs={"transtime": 3, "bandwidth": 2}),
    dict(net="6x6.net", policy="Qroute", T=500, lambd=1.0, seed=2, net_kwargs={"transtime": 5, "bandwidth": 2}, slot=2, freq=2),
    dict(net="6x6.net", policy="Shortest", T=800, lambd=1.0, seed=1, net_kwargs={"transtime": 2}),
    dict(net="lata.net", policy="Shortest", T=300, lambd=3.0, seed=1, net_kwargs={"transtime": 3, "bandwidth": 3}, freq=2),
    dict(net="6x6.net", policy="HybridQ", T=500, lambd=2.0, seed=1, rng="philox", philox_seed=7, replica=3, det_math=True,
         net_kwargs={"transtime": 2}),
    dict(net="6x6.net", policy="MaHybridQ", T=500, lambd=2.0, seed=1, rng="philox", philox_seed=7, replica=3, det_math=True,
         lr={"q": 0.1, "p": 0.001}, net_kwargs={"transtime": 3, "bandwidth": 2}),
    dict(net="6x6.net", policy="CQ", T=500, lambd=2.0, seed=1, net_kwargs={"transtime": 2}),
    dict(net="6x6.net", policy="CDRQ", T=500, lambd=2.0, seed=1, net_kwargs={"transtime": 2}),
    dict(net="6x6.net", policy="GlobalRoute", T=150, lambd=2.0, seed=1, net_kwargs={"transtime": 2}),
    dict(net="6x6.net", policy="Shortest", T=500, lambd=3.0, seed=1, rng="philox", philox_seed=7, replica=3,
         agent_kwargs={"multiway": True, "random": True}, net_kwargs={"transtime": 2}),
    dict(net="6x6.net", policy="Qroute", T=600, lambd=2.0, seed=1, net_kwargs={"transtime": 2, "is_drop": True}),
    dict(net="6x6.net", policy="Qroute", T=0, lambd=2.0, seed=1, sample_size=200, net_kwargs={"transtime": 2}),
    dict(net="lata.net", policy="Qroute", T=300, lambd=2.0, seed=1, rng="philox", philox_seed=7, replica=3,
         net_kwargs={"transtime": 2, "bandwidth": 2}),
    dict(net="6x6.net", policy="Qroute", T=400, lambd=2.0, seed=1, net_kwargs={"bandwidth": 0}),
    # GlobalRoute(multiway=True, random=True): every optimal neighbour kept by the per-step recomputation (shortest.py:57-58)
    dict(net="6x6.net", policy="GlobalRoute", T=150, lambd=3.0, seed=1, agent_kwargs={"multiway": True}),
    dict(net="6x6.net", policy="GlobalRoute", T=150, lambd=3.0, seed=1, rng="philox", philox_seed=7, replica=3,
         agent_kwargs={"multiway": True, "random": True}),
    dict(net="lata.net", policy="GlobalRoute", T=30, lambd=3.0, seed=1, rng="philox", philox_seed=7, replica=3,
         agent_kwargs={"multiway": True, "random": True}),
    dict(net="6x6.net", policy="GlobalRoute", T=100, lambd=2.0, seed=1, rng="philox", philox_seed=7, replica=3,
         agent_kwargs={"multiway": True, "random": True}, net_kwargs={"transtime": 2}),
    # SURVEY §8(c) starter rows 2 and 3 at their full horizon (trajectory hashes d99579f950a89093 / 1553ae894a6cd64f)
    dict(net="6x6-modified.net", policy="Qroute", T=10000, lambd=2.0, seed=1),     # BASELINE config 3's network, T = 10,000
    dict(net="lata.net", policy="Qroute", T=10000, lambd=2.0, seed=1),             # BASELINE config 4's network, T = 10,000
]


def main():
    out, arrays = [], {}
    start = int(sys.argv[sys.argv.index("--from") + 1]) if "--from" in sys.argv else 0     # keep records [0, start) as they are
    if start:
        with open(os.path.join(HERE, "kat.json")) as f:
            out = json.load(f)[:start]
        arrays = dict(np.load(os.path.join(HERE, "np_tables.npz")))
    for i, c in enumerate(CASES):
        if i < start:
            assert {k: v for k, v in out[i].items() if k not in ("id", "expect")} == c, f"case {i} changed"
            continue
        kw = dict(c)
        keep = kw.pop("keep_tables", False)
        r = H.run(kw.pop("net"), kw.pop("policy"), kw.pop("T"), kw.pop("lambd"), **kw)
        rec = dict(c)
        rec["id"] = i
        rec["expect"] = {
            "counters": r["counters"], "ave_route_time": r["ave_route_time"],
            "sendlog_sha256": r["sendlog_sha256"], "n_sends": r["n_sends"],
            "inj_sha256": h(np.concatenate([r["inj_cnt"], r["inj_pairs"].reshape(-1)]), "<i4"),
            "route_time_sha256": h(r["route_time"], "<i8"), "end_packets_sha256": h(r["end_packets"], "<i8"),
            "hops_sha256": h(r["hops"], "<i8"),
            "queues_sha256": h(np.concatenate([q.reshape(-1) for q in r["queues"]] + [np.array([len(q) for q in r["queues"]])]), "<i4"),
            "error": r["error"], "in_flight": r["in_flight"],
        }
        for name in ("Q0", "Qtable", "Theta", "Trace", "confidence"):
            if name in r:
                rec["expect"][name + "_sha256"] = h(r[name], "<f8")
                rec["expect"][name + "_sum"] = float(r[name].sum())
                if keep:
                    arrays[f"case{i}_{name}"] = r[name]
        if "distance" in r:
            rec["expect"]["distance_sha256"] = h(r["distance"], "<f8")
            rec["expect"]["choice_sha256"] = h(r["choice"], "u1")
        if "sample" in r:
            rec["expect"]["sample"] = [int(v) for v in r["sample"]]
        if "queue_size" in r:
            rec["expect"]["queue_size"] = [int(v) for v in r["queue_size"]]
        out.append(rec)
        print(i, c["net"], c["policy"], r["counters"], r["error"], flush=True)
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(out, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "np_tables.npz"), **arrays)


if __name__ == "__main__":
    main()
