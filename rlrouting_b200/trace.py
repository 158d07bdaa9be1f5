"""Trace mode: replay of the reference's NumPy legacy RNG stream on the host.

The reference draws everything from NumPy's process-wide legacy `RandomState`
(train.py:15 `np.random.seed(1)`).  For table/argmax policies (Shortest, Qroute) the only
consumers are the agent constructor (qroute.py:13-15) and `Network.new_packet`
(env.py:298-312), and traffic does not depend on network state, so a replica's whole injection
schedule can be generated up front by issuing exactly the reference's call sequence against a
private `RandomState(seed)` — the stream is frozen by NumPy policy — and uploaded to the GPU.

Replica r of a batched `Network` uses seed `base_seed + r`.
"""
import numpy as np


def draw_qtable_normals(rs, topo, initQ=0.0, times=1):
    """The constructor's draws: one `normal(initQ, 1, (N, deg_x))` per node in id order
    (qroute.py:13-15).  HybridQ/MaHybridQ run Qroute.__init__ twice through the MRO
    (hybrid.py:10,44-47); the second draw is the one kept, hence `times`."""
    out = None
    for _ in range(times):
        out = {x: rs.normal(initQ, 1, (topo.N, int(topo.deg[x]))) for x in range(topo.N)}
    return out


def draw_injections(rs, n_nodes, n_steps, lambd):
    """The traffic draws of `n_steps` calls of new_packet(lambd) (env.py:307-310).

    Returns (cnt[int32 n_steps], pairs[int32 total, 2]) with pairs[:,0]=source, [:,1]=dest."""
    cnt = np.zeros(n_steps, dtype=np.int32)
    pairs = []
    poisson, randint = rs.poisson, rs.randint
    for s in range(n_steps):
        n = poisson(lambd)
        cnt[s] = n
        for _ in range(n):
            source, dest = randint(0, n_nodes, size=2)
            while dest == source:
                dest = randint(0, n_nodes)
            pairs.append((source, dest))
    return cnt, np.array(pairs, dtype=np.int32).reshape(-1, 2)
