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


def draw_injections_native(rs, n_nodes, n_steps, lambd):
    """draw_injections through the C replay of the legacy stream (rlr_trace_replay): same draws, same final
    RandomState state, ~100x faster than one NumPy call per draw.  Needs lambd < 10."""
    import ctypes as C
    from . import _native as nat
    name, key, pos, has_gauss, cached = rs.get_state()
    key = np.ascontiguousarray(key, dtype=np.uint32).copy()
    cpos = C.c_int32(int(pos))
    cap = int(lambd * n_steps + 12.0 * np.sqrt(lambd * n_steps + 1.0) + 64)
    cnt = np.zeros(n_steps, dtype=np.int32)
    packed = np.zeros(cap, dtype=np.uint32)
    n_out = C.c_int64(0)
    nat.check(nat.lib().rlr_trace_replay(key.ctypes.data, C.byref(cpos), n_nodes, n_steps, float(lambd),
                                          cnt.ctypes.data, packed.ctypes.data, cap, C.byref(n_out)))
    rs.set_state((name, key, cpos.value, has_gauss, cached))
    packed = packed[:n_out.value]
    return cnt, np.stack([packed & 0xFFFF, packed >> 16], axis=1).astype(np.int32).reshape(-1, 2)


def draw_injections(rs, n_nodes, n_steps, lambd, native=True):
    """The traffic draws of `n_steps` calls of new_packet(lambd) (env.py:307-310).

    Returns (cnt[int32 n_steps], pairs[int32 total, 2]) with pairs[:,0]=source, [:,1]=dest."""
    if native and 0.0 <= lambd < 10.0 and n_nodes <= 65535:
        return draw_injections_native(rs, n_nodes, n_steps, lambd)
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


def _draw_worker(args):
    state, n_nodes, n_steps, lambd = args
    rs = np.random.RandomState()
    rs.set_state(state)
    cnt, pairs = draw_injections(rs, n_nodes, n_steps, lambd)
    return cnt, pairs, rs.get_state()


def _draw_native_many(streams, n_nodes, n_steps, lambdas, workers):
    """All replicas' streams through one call of rlr_trace_replay_many (C threads, no GIL)."""
    import ctypes as C
    from . import _native as nat
    R = len(streams)
    states = [rs.get_state() for rs in streams]
    keys = np.ascontiguousarray(np.stack([st[1] for st in states]), dtype=np.uint32)
    pos = np.array([st[2] for st in states], dtype=np.int32)
    lam = np.ascontiguousarray(np.broadcast_to(lambdas, (R,)), dtype=np.float64)
    m = float(lam.max()) * n_steps
    cap = int(m + 12.0 * np.sqrt(m + 1.0) + 64)
    cnt = np.zeros((R, n_steps), dtype=np.int32)
    packed = np.zeros((R, cap), dtype=np.uint32)
    n_out = np.zeros(R, dtype=np.int64)
    nat.check(nat.lib().rlr_trace_replay_many(keys.ctypes.data, pos.ctypes.data, R, n_nodes, n_steps, lam.ctypes.data,
                                               cnt.ctypes.data, packed.ctypes.data, cap, n_out.ctypes.data, workers))
    out = []
    for r, (rs, st) in enumerate(zip(streams, states)):
        rs.set_state((st[0], keys[r], int(pos[r]), st[3], st[4]))
        pk = packed[r, :n_out[r]]
        out.append((cnt[r], np.stack([pk & 0xFFFF, pk >> 16], axis=1).astype(np.int32).reshape(-1, 2)))
    return out


class TraceBatch:
    """The replayed injections of many replicas.  `schedule(it0, it1, freq)` gives the per-node event lists of the
    iterations [it0, it1) as a uint32 array [R, N, cap] (the layout of rlr_run_params_t.inj_events)."""

    def __init__(self, n_nodes, lists=None, cnt=None, packed=None, workers=1):
        self.n_nodes, self.lists, self.cnt, self.packed, self.workers = n_nodes, lists, cnt, packed, workers

    def schedule(self, it0, it1, freq=1):
        N = self.n_nodes
        if self.lists is None:                       # packed arrays of rlr_trace_replay_many: built in C, all host threads
            import ctypes as C
            from . import _native as nat
            R, T = self.cnt.shape
            cap = C.c_int32(0)
            args = (self.cnt.ctypes.data, self.packed.ctypes.data, self.packed.shape[1], R, N, T, int(it0), int(it1), int(freq))
            nat.check(nat.lib().rlr_trace_schedule(*args, None, 0, C.byref(cap), self.workers))
            cap.value += cap.value & 1                      # even: the fused step kernel reads the lists in 8-byte chunks
            sched = np.empty((R, N, cap.value), dtype=np.uint32)
            nat.check(nat.lib().rlr_trace_schedule(*args, sched.ctypes.data, cap.value, None, self.workers))
            return sched
        from .env import trace_schedule
        evs = [trace_schedule(c, pr, N, it0, it1, freq) for c, pr in self.lists]
        per_node = [np.bincount(node, minlength=N) for _, node in evs]
        cap = int(max(pn.max() for pn in per_node)) + 2
        cap += cap & 1                                      # even (8-byte chunks, see above)
        sched = np.full((len(evs), N, cap), 0xFFFFFFFF, dtype=np.uint32)
        for r, ((e, node), pn) in enumerate(zip(evs, per_node)):
            start = np.concatenate([[0], np.cumsum(pn)[:-1]])
            sched[r, node, np.arange(len(e)) - start[node]] = e
        return sched


def draw_batch(streams, n_nodes, n_steps, lambdas, workers=None):
    """draw_injections_many as a TraceBatch; large batches stay in the packed form of the C replay."""
    import os
    R = len(streams)
    if workers is None:
        workers = min(os.cpu_count() or 1, 32)
    lam = np.asarray(np.broadcast_to(lambdas, (R,)), dtype=np.float64)
    native = (R >= 64 and workers >= 2 and float(lam.max()) < 10.0 and n_nodes <= 255
              and not any(rs is np.random.mtrand._rand for rs in streams))
    if not native:
        return TraceBatch(n_nodes, lists=draw_injections_many(streams, n_nodes, n_steps, lam, workers))
    cnt, packed = _replay_packed(streams, n_nodes, n_steps, lam, workers)
    return TraceBatch(n_nodes, cnt=cnt, packed=packed, workers=workers)


def _replay_packed(streams, n_nodes, n_steps, lam, workers):
    """rlr_trace_replay_many on every stream (their states advance); returns (cnt[R, n_steps], packed[R, cap])."""
    import ctypes as C
    from . import _native as nat
    R = len(streams)
    states = [rs.get_state() for rs in streams]
    keys = np.ascontiguousarray(np.stack([st[1] for st in states]), dtype=np.uint32)
    pos = np.array([st[2] for st in states], dtype=np.int32)
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    m = float(lam.max()) * n_steps
    cap = int(m + 12.0 * np.sqrt(m + 1.0) + 64)
    cnt = np.zeros((R, n_steps), dtype=np.int32)
    packed = np.zeros((R, cap), dtype=np.uint32)
    n_out = np.zeros(R, dtype=np.int64)
    nat.check(nat.lib().rlr_trace_replay_many(keys.ctypes.data, pos.ctypes.data, R, n_nodes, n_steps, lam.ctypes.data,
                                               cnt.ctypes.data, packed.ctypes.data, cap, n_out.ctypes.data, workers))
    for r, (rs, st) in enumerate(zip(streams, states)):
        rs.set_state((st[0], keys[r], int(pos[r]), st[3], st[4]))
    return cnt, packed


def draw_injections_many(streams, n_nodes, n_steps, lambdas, workers=None):
    """draw_injections for every replica's RandomState.  Replicas are independent streams: the C replay releases the
    GIL, so they are replayed on a thread pool; loads NumPy has to draw itself (lambda >= 10) go to worker processes."""
    import os
    R = len(streams)
    if workers is None:
        workers = min(os.cpu_count() or 1, 32)
    lam_max = float(np.max(lambdas)) if R else 0.0
    if R < 64 or workers < 2 or any(rs is np.random.mtrand._rand for rs in streams):
        return [draw_injections(rs, n_nodes, n_steps, lambdas[r]) for r, rs in enumerate(streams)]
    if lam_max < 10.0 and n_nodes <= 65535:
        return _draw_native_many(streams, n_nodes, n_steps, np.asarray(lambdas, dtype=np.float64), workers)
    import multiprocessing as mp
    with mp.get_context("fork").Pool(workers) as pool:
        res = pool.map(_draw_worker, [(rs.get_state(), n_nodes, n_steps, float(lambdas[r])) for r, rs in enumerate(streams)],
                       chunksize=max(1, R // (workers * 8)))
    out = []
    for rs, (cnt, pairs, state) in zip(streams, res):
        rs.set_state(state)
        out.append((cnt, pairs))
    return out
