"""ctypes wrapper over oracle/librlr_oracle.so (built by oracle/Makefile).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
Parity pin: tests/golden/ (generated from the unmodified reference by gen_golden.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "librlr_oracle.so")

POLICY_ID = {"Shortest": 0, "Qroute": 1, "HybridQ": 2, "MaHybridQ": 3, "CQ": 4, "CDRQ": 5, "GlobalRoute": 6}
_lib = None


def build(force=False):
    src = os.path.join(HERE, "rlr_oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B"])
    return SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        L = C.CDLL(SO)
        p, i32, i64, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_double
        L.orc_create.restype = p
        L.orc_create.argtypes = [i32, p, p, i32]
        L.orc_destroy.argtypes = [p]
        L.orc_set_hparams.argtypes = [p, f64, f64, i32]
        L.orc_set_stream.argtypes = [p, C.c_uint64, C.c_uint64]
        L.orc_set_decay.argtypes = [p, f64]
        L.orc_set_is_drop.argtypes = [p, i32]
        L.orc_set_timing.argtypes = [p, i32, i32, i32, i32]
        L.orc_set_q_f32.argtypes = [p, i32]
        L.orc_in_flight.restype = i32
        L.orc_in_flight.argtypes = [p]
        L.orc_set_shortest_flags.argtypes = [p, i32, i32]
        L.orc_cq_clean.argtypes = [p]
        L.orc_table_size.restype = i64
        L.orc_table_size.argtypes = [p]
        L.orc_table.restype = C.POINTER(C.c_double)
        L.orc_table.argtypes = [p, i32]
        L.orc_choice.restype = C.POINTER(C.c_uint8)
        L.orc_choice.argtypes = [p]
        L.orc_distance.restype = C.POINTER(C.c_double)
        L.orc_distance.argtypes = [p]
        L.orc_status.restype = i32
        L.orc_status.argtypes = [p, p]
        L.orc_counters.argtypes = [p, p]
        L.orc_queue_len.restype = i64
        L.orc_queue_len.argtypes = [p, i32]
        L.orc_queue_get.argtypes = [p, i32, p]
        L.orc_qtable_structure.argtypes = [p]
        L.orc_reset.argtypes = [p]
        L.orc_shortest_build.argtypes = [p]
        L.orc_globalroute_init.argtypes = [p]
        L.orc_queue_size.restype = C.POINTER(C.c_int64)
        L.orc_queue_size.argtypes = [p]
        L.orc_heap_perm.argtypes = [i32, p]
        L.orc_philox4x32.argtypes = [p, p, p]
        L.orc_run.restype = i64
        L.orc_run.argtypes = [p, i64, i32, p, p, f64, f64, f64, f64, p, p, p, p, i64, p]
        L.orc_sample_route_time.restype = i64
        L.orc_sample_route_time.argtypes = [p, i64, i64, p, p, f64, f64, f64, f64, p]
        L.orc_run_many.restype = i64
        L.orc_run_many.argtypes = [p, i32, i64, p, f64, f64, f64, i32, p, p]
        L.orc_max_threads.restype = i32
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


COUNTER_NAMES = ("clock", "all_packets", "end_packets", "drop_packets", "active_packets",
                 "hops", "route_time", "sends")


class OracleEnv:
    """One environment replica + its agent (reference `Network` + `Policy`)."""

    def __init__(self, links, policy, discount=None, discount_trace=0.6, add_entropy=True,
                 seed=0, replica=0, decay=0.9, is_drop=False, multiway=False, random=False,
                 bandwidth=1, transtime=1):
        if discount is None:
            discount = 0.9 if policy in ("CQ", "CDRQ") else 0.99          # qroute.py:59 vs qroute.py:9
        L = lib()
        self.N = len(links)
        self.deg = np.array([len(links[x]) for x in range(self.N)], dtype=np.int32)
        self.row_ptr = np.zeros(self.N + 1, dtype=np.int32)
        self.row_ptr[1:] = np.cumsum(self.deg)
        self.col = np.array([y for x in range(self.N) for y in links[x]], dtype=np.int32)
        self.policy = policy
        self.h = L.orc_create(self.N, _ptr(self.row_ptr), _ptr(self.col), POLICY_ID[policy])
        L.orc_set_hparams(self.h, discount, discount_trace, int(add_entropy))
        L.orc_set_stream(self.h, seed, replica)
        L.orc_set_decay(self.h, decay)
        L.orc_set_is_drop(self.h, int(is_drop))
        self.bandwidth, self.transtime = int(bandwidth), int(transtime)
        L.orc_set_timing(self.h, self.bandwidth, self.transtime, 1, 1)
        L.orc_set_shortest_flags(self.h, int(multiway), int(random))
        self.tsize = L.orc_table_size(self.h)
        self.toff = self.row_ptr.astype(np.int64) * self.N
        self.Q = np.ctypeslib.as_array(L.orc_table(self.h, 0), shape=(self.tsize,))
        self.Theta = np.ctypeslib.as_array(L.orc_table(self.h, 1), shape=(self.tsize,))
        self.Conf = self.Theta                          # CQ.confidence shares the storage
        self.Trace = np.ctypeslib.as_array(L.orc_table(self.h, 2), shape=(self.tsize,))
        self.choice = np.ctypeslib.as_array(L.orc_choice(self.h), shape=(self.tsize,))
        self.distance = np.ctypeslib.as_array(L.orc_distance(self.h), shape=(self.N, self.N))
        if policy == "Shortest":
            L.orc_shortest_build(self.h)
        if policy == "GlobalRoute":
            L.orc_globalroute_init(self.h)
        self.queue_size = np.ctypeslib.as_array(L.orc_queue_size(self.h), shape=(self.N,))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.orc_destroy(self.h)
            self.h = None

    def set_q_storage_f32(self, on=True):
        """Checker of the product's fp32 table-storage mode (Qroute): Q values are rounded to float whenever they are
        stored; call before set_qtable."""
        self.q_f32 = bool(on)
        lib().orc_set_q_f32(self.h, int(on))

    def set_qtable(self, normals, initP=0.0):
        """normals: flat [tsize] N(initQ,1) draws in packed layout; applies qroute.py:16-20."""
        self.Q[:] = np.asarray(normals, dtype=np.float32) if getattr(self, "q_f32", False) else normals
        lib().orc_qtable_structure(self.h)
        self.Theta[:] = initP
        self.Trace[:] = 0.0
        if self.policy in ("CQ", "CDRQ"):
            lib().orc_cq_clean(self.h)                  # CQ.__init__ -> clean(), qroute.py:64

    def reset(self):
        lib().orc_reset(self.h)

    def counters(self):
        out = np.zeros(8, dtype=np.int64)
        lib().orc_counters(self.h, _ptr(out))
        return dict(zip(COUNTER_NAMES, (int(v) for v in out)))

    def queue(self, x):
        n = lib().orc_queue_len(self.h, x)
        out = np.zeros((n, 5), dtype=np.int32)
        if n:
            lib().orc_queue_get(self.h, x, _ptr(out))
        return out

    def status(self):
        clk = C.c_int64(0)
        st = lib().orc_status(self.h, C.byref(clk))
        return st, clk.value

    def set_timing(self, slot=1, freq=1):
        lib().orc_set_timing(self.h, self.bandwidth, self.transtime, int(slot), int(freq))
        self.slot, self.freq = int(slot), int(freq)

    def run(self, n_steps, lambd=None, inj=None, lr_q=0.1, lr_p=0.1, lr_e=0.1, sendlog=False, slot=1, freq=1):
        """inj=(cnt, pairs) -> trace mode; else Philox mode at rate `lambd` (per unit time: Poisson(lambd * slot) per
        iteration, env.py:401).  One iteration = one injection + `freq` calls of step(slot)."""
        self.set_timing(slot, freq)
        rt = np.zeros(n_steps, dtype=np.int64)
        en = np.zeros(n_steps, dtype=np.int64)
        hp = np.zeros(n_steps, dtype=np.int64)
        log = np.zeros((n_steps * freq * self.N, 4), dtype=np.int32) if sendlog else None
        nlog = C.c_int64(0)
        if inj is not None:
            cnt = np.ascontiguousarray(inj[0], dtype=np.int32)
            pairs = np.ascontiguousarray(inj[1], dtype=np.int32)
            assert len(cnt) >= n_steps
            mode, enlam = 0, 0.0
        else:
            cnt = pairs = None
            mode, enlam = 1, float(np.exp(-(np.float64(lambd) * slot) / np.float64(self.N)))
        done = lib().orc_run(self.h, n_steps, mode, _ptr(cnt), _ptr(pairs), enlam, lr_q, lr_p,
                             lr_e, _ptr(rt), _ptr(en), _ptr(hp), _ptr(log),
                             0 if log is None else len(log), C.byref(nlog))
        res = {"steps_done": int(done), "route_time": rt, "end_packets": en, "hops": hp}
        if sendlog:
            res["sendlog"] = log[:nlog.value].copy()
        return res


def sample_route_time(env, size, lambd=None, inj=None, lr_q=0.1, lr_p=0.1, lr_e=0.1, max_steps=1 << 20, slot=1, freq=1):
    """Network.sample_route_time on one OracleEnv; returns (sample[int64 size], steps_run)."""
    out = np.zeros(size, dtype=np.int64)
    env.set_timing(slot, freq)
    if inj is not None:
        cnt = np.ascontiguousarray(inj[0], dtype=np.int32)
        pairs = np.ascontiguousarray(inj[1], dtype=np.int32)
        max_steps, enlam = min(max_steps, len(cnt)), 0.0
    else:
        cnt = pairs = None
        enlam = float(np.exp(-(np.float64(lambd) * slot) / np.float64(env.N)))
    steps = lib().orc_sample_route_time(env.h, size, max_steps, _ptr(cnt), _ptr(pairs), enlam, lr_q, lr_p, lr_e, _ptr(out))
    return out, int(steps)


def run_many(envs, n_steps, lambd, lr_q=0.1, lr_p=0.1, lr_e=0.1, nthreads=0, metrics=False, slot=1, freq=1):
    """Philox-mode batch over independent OracleEnv objects on `nthreads` host threads."""
    n = len(envs)
    for e in envs:
        e.set_timing(slot, freq)
    hs = (C.c_void_p * n)(*[e.h for e in envs])
    lam = np.broadcast_to(np.asarray(lambd, dtype=np.float64), (n,))
    enlam = np.exp(-(lam * slot) / np.float64(envs[0].N))
    rt = np.zeros((n, n_steps), dtype=np.int64) if metrics else None
    en = np.zeros((n, n_steps), dtype=np.int64) if metrics else None
    sends = lib().orc_run_many(hs, n, n_steps, _ptr(np.ascontiguousarray(enlam)), lr_q, lr_p,
                               lr_e, nthreads, _ptr(rt), _ptr(en))
    return (int(sends), rt, en) if metrics else int(sends)


def heap_perm(k):
    out = np.zeros(k, dtype=np.int32)
    lib().orc_heap_perm(k, _ptr(out))
    return out.tolist()


def philox4x32(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32(_ptr(c), _ptr(k), _ptr(out))
    return out
