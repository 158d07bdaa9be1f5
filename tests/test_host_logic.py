"""CPU tests of the host-side logic: topology layout, heap-tie table, injection threshold, API
validation, the C-ABI library's exports, replica sharding."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rlrouting_b200 import _native as nat
from rlrouting_b200 import dist, nets
from rlrouting_b200.env import injection_threshold
from rlrouting_b200.topology import Topology, heap_pop_order, parse_net

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_loads_and_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rlrouting_b200.h")).read()
    declared = set(re.findall(r"\b(rlr_[a-z_]+)\s*\(", hdr))
    assert declared == set(nat.EXPORTS)
    lib = nat.lib()                                  # raises if the .so is missing or a symbol is absent
    for name in declared:
        assert hasattr(lib, name)
    assert lib.rlr_abi_version() == nat.ABI_VERSION
    # argument validation happens before any CUDA call, so it is testable without a GPU
    t = nat.Topology(1, 0, 0, None, None)
    cfg = nat.Config(1, 0, 1, 64, 0, 0, -1, 0, 0, 0)
    h = C.c_void_p()
    assert lib.rlr_create(C.byref(t), C.byref(cfg), C.byref(h)) < 0
    assert b"n_nodes" in lib.rlr_last_error_string()
    assert lib.rlr_run_steps(None, 1, None, None) < 0


def test_struct_sizes_match_header_layout():
    sizes = (C.c_int32 * 4)()
    assert nat.lib().rlr_struct_sizes(sizes) == 0
    assert list(sizes) == [C.sizeof(nat.Topology), C.sizeof(nat.Config), C.sizeof(nat.Buffers), C.sizeof(nat.RunParams)]
    assert C.sizeof(nat.Buffers) == 28 * 8


def test_builtin_topologies_layout():
    sizes = {"6x6.net": (36, 100, 4), "6x6-modified.net": (36, 100, 5), "lata.net": (116, 336, 9)}
    for name, (n, sdeg, maxdeg) in sizes.items():
        t = Topology.from_text(nets.net_text(name))
        assert (t.N, t.sum_deg, t.max_deg, t.tsize) == (n, sdeg, maxdeg, n * sdeg)
        assert t.row_ptr[-1] == sdeg and len(t.col) == sdeg
        for x in range(n):
            for j, y in enumerate(t.links[x]):
                assert t.col[t.row_ptr[x] + j] == y
                assert t.links[y][t.rev[t.row_ptr[x] + j]] == x          # undirected: the reverse slot exists
        tabs = {x: np.arange(n * t.deg[x], dtype=np.float64).reshape(n, t.deg[x]) + 1000 * x for x in range(n)}
        back = t.unpack(t.pack(tabs))
        assert all(np.array_equal(back[x], tabs[x]) for x in range(n))


@pytest.mark.reference
def test_builtin_topologies_equal_reference_files():
    for name in nets.TOPOLOGIES:
        ref, _ = parse_net(open(os.path.join("/root/reference", name)).read())
        assert ref == Topology.from_text(nets.net_text(name)).links


def test_heap_position_table_is_inverse_of_pop_order():
    t = Topology.from_text(nets.net_text("6x6.net"))
    for k in range(1, t.N + 1):
        order = heap_pop_order(k)
        assert sorted(order) == list(range(k))
        for j, rank in enumerate(order):
            assert t.heap_pos[k, rank] == j


def test_injection_threshold_is_exact():
    rng = np.random.default_rng(0)
    lam = np.concatenate([rng.uniform(0, 8, 2000), [0.0, 1e-12, 0.25, 2.0, 4.0, 50.0, 2000.0]])
    for n_nodes in (2, 36, 116):
        en = np.exp(-lam / n_nodes)
        thr = injection_threshold(en)
        for e, t in zip(en, thr):
            t = int(t)
            u = lambda w: (np.float64(w) + 0.5) * (1.0 / 4294967296.0)          # noqa: E731  (the oracle's expression)
            if t <= 0xFFFFFFFF:
                assert u(t) > e
            if t > 0:
                assert not (u(t - 1) > e)
    assert int(injection_threshold(1.0)) == 1 << 32                              # lambda = 0: never inject


def test_unsupported_options_fail_before_touching_the_device():
    import rlrouting_b200 as rb
    for kw in (dict(bandwidth=1.5), dict(transtime=0.5), dict(transtime=2.25, is_drop=True)):     # float clocks
        with pytest.raises(NotImplementedError):
            rb.Network("6x6.net", **kw)
    with pytest.raises(ValueError):
        rb.Network("6x6.net", rng="mt")
    with pytest.raises(FileNotFoundError):
        rb.Network("no-such-file.net")
    with pytest.raises(NotImplementedError):
        Topology({0: [], 1: []})                                                 # isolated nodes
    with pytest.raises(NotImplementedError):
        Topology({0: [0, 1], 1: [0]})                                            # self-loop


def test_sharding_covers_every_replica_once():
    for total in (1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            blocks = [dist.shard(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
            for (b0, c0), (b1, _) in zip(blocks, blocks[1:]):
                assert b0 + c0 == b1
    base, lam, idx = dist.shard_loads([1.0, 2.0, 3.0, 4.0], 6, 1, 4)
    assert base == 6 and list(lam) == [2.0] * 6 and list(idx) == [1] * 6


def test_trace_schedule_orders_events_per_node():
    from rlrouting_b200.env import schedule_capacity, trace_schedule
    from rlrouting_b200 import trace
    rs = np.random.RandomState(3)
    cnt, pairs = trace.draw_injections(rs, 36, 400, 3.0)
    ev, node = trace_schedule(cnt, pairs, 36, 100, 300)
    steps = np.repeat(np.arange(400), cnt)
    for x in range(36):
        want = [((s - 100) << 8) | d for s, (src, d) in zip(steps, pairs) if src == x and 100 <= s < 300]
        assert list(ev[node == x]) == want                       # (step, draw) order inside the node
    assert len(ev) == int(((steps >= 100) & (steps < 300)).sum())
    assert schedule_capacity(4.0, 36, 500) > 4.0 / 36 * 500 + 40


def test_native_trace_replay_equals_numpy_legacy_stream():
    """rlr_trace_replay (C) must consume NumPy's legacy RandomState stream exactly as env.py:307-310 does."""
    from rlrouting_b200 import trace
    for seed in range(12):
        for lam, n_nodes, steps in ((2.0, 36, 400), (0.0, 36, 30), (0.3, 116, 500), (9.5, 36, 150), (4.0, 2, 200), (1.0, 255, 200)):
            a, b = np.random.RandomState(seed), np.random.RandomState(seed)
            for rs in (a, b):
                rs.normal(0, 1, (36, 4))
                rs.normal()                                   # leaves a cached gaussian in the state
            x = trace.draw_injections(a, n_nodes, steps, lam, native=True)
            y = trace.draw_injections(b, n_nodes, steps, lam, native=False)
            assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1])
            assert a.randint(0, 1 << 30) == b.randint(0, 1 << 30) and a.normal() == b.normal()   # stream continues identically
    rs = np.random.RandomState(0)
    cnt, pairs = trace.draw_injections(rs, 36, 100, 12.0)     # lambda >= 10: NumPy's PTRS branch, falls back to NumPy
    assert cnt.sum() == len(pairs) and np.all(pairs[:, 0] != pairs[:, 1])


def test_integration_doc_ctypes_stub_matches_the_library():
    """The binding shown in INTEGRATION.md must describe the structs the shipped library was built with."""
    src = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = src[src.index("class Topology(C.Structure)"):src.index("sizes = (C.c_int32 * 4)()")]
    ns = {"C": C}
    exec(code, ns)
    sizes = (C.c_int32 * 4)()
    nat.lib().rlr_struct_sizes(sizes)
    assert list(sizes) == [C.sizeof(ns[k]) for k in ("Topology", "Config", "Buffers", "RunParams")]
    assert [f[0] for f in ns["RunParams"]._fields_] == [f[0] for f in nat.RunParams._fields_]


def test_integer_timing_arguments():
    """slot / freq / transtime / bandwidth are integers on the device (integer-valued floats are accepted)."""
    from rlrouting_b200.env import _as_int, trace_schedule
    assert _as_int(2, "slot") == 2 and _as_int(3.0, "transtime") == 3 and _as_int(np.int64(4), "freq") == 4
    for bad in (0.5, 2.25, True):
        with pytest.raises(NotImplementedError):
            _as_int(bad, "slot")
    # freq > 1: iteration i of train's loop injects at Network.step call i * freq (env.py:400-403)
    cnt = np.array([1, 0, 2], dtype=np.int32)
    pairs = np.array([[0, 1], [2, 0], [2, 1]], dtype=np.int32)
    ev, node = trace_schedule(cnt, pairs, 3, 0, 3, freq=3)
    assert list(node) == [0, 2, 2] and list(ev >> 8) == [0, 6, 6] and list(ev & 0xFF) == [1, 0, 1]


def test_trace_batch_schedule_native_equals_python():
    """rlr_trace_schedule (C, all host threads) builds the same per-node event lists as the NumPy path."""
    from rlrouting_b200 import trace
    R, N, T = 96, 36, 300
    lam = np.linspace(0.5, 3.0, R)
    a = [np.random.RandomState(7 + r) for r in range(R)]
    b = [np.random.RandomState(7 + r) for r in range(R)]
    nat_batch = trace.draw_batch(a, N, T, lam, workers=4)
    assert nat_batch.lists is None                      # the packed C form
    py_batch = trace.TraceBatch(N, lists=[trace.draw_injections(rs, N, T, lam[r]) for r, rs in enumerate(b)])
    for it0, it1, freq in ((0, T, 1), (37, 211, 3), (T - 1, T, 1), (5, 5, 1)):
        s1, s2 = nat_batch.schedule(it0, it1, freq), py_batch.schedule(it0, it1, freq) if it1 > it0 else None
        if s2 is None:
            assert s1.shape[2] == 2 and np.all(s1 == 0xFFFFFFFF)
            continue
        assert s1.shape == s2.shape and np.array_equal(s1, s2)
    assert all(np.array_equal(x.get_state()[1], y.get_state()[1]) for x, y in zip(a, b))
