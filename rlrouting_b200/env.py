"""`Network`: the reference's simulator API (env.py:213-450) over R replicas on one B200.

With no new keywords a `Network(file)` is ONE environment driven by NumPy's global legacy RNG, so
`np.random.seed(1); nw = Network("6x6.net"); nw.agent = Qroute(nw); nw.train(10000, 2.0)` returns
the reference's numbers bit for bit.  The additive keywords (`replicas`, `rng`, `seed`, ...) batch R
independent environments; every counter property then returns an array of length R.

Integer `bandwidth`, `transtime`, `slot` and `freq`, `is_drop`, `sample_route_time` and the 'dual' node mode
(CDRQ) all run on the device.  What the device cannot do raises NotImplementedError when it is requested, never
mid-run, and there is no CPU fallback: float-valued clocks (`slot`, `transtime` ...), mode 'bp',
'dual' for anything but CDRQ, and agents that are not one of the built-in policies.
"""
import ctypes as C
import os
from collections import OrderedDict

import numpy as np

from . import _native as nat
from . import nets, pinned, trace
from .base_policy import Policy
from .topology import Topology


class Packet:
    """Read-back view of a queued packet (reference env.py:9-27)."""
    __slots__ = ("source", "dest", "birth", "hops", "trans_time", "start_queue")

    def __init__(self, source, dest, birth, hops=0, start_queue=0):
        self.source, self.dest, self.birth, self.hops = source, dest, birth, hops
        self.trans_time, self.start_queue = 1, start_queue

    def __repr__(self):
        return f"Packet<{self.source}->{self.dest}>"


class Reward:
    """What `Network.step` returns per send (reference env.py:52-70)."""

    def __init__(self, source, packet, action, agent_info=None):
        self.source, self.dest, self.action = source, packet.dest, action
        self.packet, self.agent_info = packet, agent_info or {}

    def __repr__(self):
        return f"Reward<{self.source}->{self.dest} by {self.action}>"


class Node:
    """Read-only view of one node of one replica (reference env.py:73-210).  `Network.nodes` holds the views of replica 0,
    like the reference's single environment; `Network.nodes_of(r)` gives those of replica r."""

    def __init__(self, ID, network, replica=0):
        self.ID, self.network, self.replica = ID, network, int(replica)

    @property
    def links(self):
        return self.network.links[self.ID]

    @property
    def queue(self):
        return [Packet(int(s), int(d), int(b), int(h), int(q)) for d, s, h, b, q in self.network.queue_array(self.ID, self.replica)]

    @property
    def sent(self):
        """Packets in flight per link (env.py:115-116): all zero between steps unless events outlive a step
        (transtime > slot), in which case the device keeps the counts (general timing model)."""
        nw = self.network
        if not nw._general:
            return dict.fromkeys(self.links, 0)  # every event lands inside its step (env.py:349-353)
        rp = int(nw.topology.row_ptr[self.ID])
        cnt = nw._link_sent[self.replica, rp:rp + len(self.links)].cpu().numpy()
        return {y: int(c) for y, c in zip(self.links, cnt)}

    def __repr__(self):
        return f"Node<{self.ID}, queue: {self.queue}, sent: {self.sent}>"


def _nodes_of(network, replica):
    if not 0 <= int(replica) < network.replicas:
        raise IndexError(f"replica {replica} out of range [0, {network.replicas})")
    return OrderedDict((i, Node(i, network, replica)) for i in network.links.keys())


def injection_threshold(enlam):
    """Smallest integer w in [0, 2^32] with (w + 0.5) * 2^-32 > enlam, exactly (all float64 steps below are
    exact: scaling by 2^33, subtracting 1 and halving a value below 2^33 lose no bits).  The device tests
    `w0 >= threshold` instead of evaluating the first Poisson uniform in fp64."""
    e = np.asarray(enlam, dtype=np.float64) * 8589934592.0
    return (np.floor((e - 1.0) / 2.0) + 1.0).clip(0.0, 4294967296.0).astype(np.uint64)


def schedule_capacity(lam_max, n_nodes, n_steps):
    """Entries per (replica, node) of a Philox-mode injection schedule: the Poisson(lambda/N * steps) count of
    one node, its mean plus 12 standard deviations plus slack (overflow probability < 1e-30), + terminator."""
    m = float(lam_max) / n_nodes * n_steps
    cap = int(m + 12.0 * np.sqrt(m) + 24) + 1
    return cap + (cap & 1)              # even: the fused step kernel reads the lists in 8-byte chunks (cp.async)


def trace_schedule(cnt, pairs, n_nodes, s0, s1, freq=1):
    """Per-node event lists of iterations [s0, s1) of one replica's trace: (events[int], node_of_event[int]) with
    events = (local_step << 8 | dest) in (node, step, draw order); iteration i injects at local step (i - s0) * freq."""
    steps = np.repeat(np.arange(len(cnt), dtype=np.int64), cnt)
    sel = (steps >= s0) & (steps < s1)
    src, dst, st = pairs[sel, 0].astype(np.int64), pairs[sel, 1].astype(np.int64), (steps[sel] - s0) * freq
    order = np.argsort(src, kind="stable")                       # stable: keeps (step, draw) order inside a node
    return ((st[order] << 8) | dst[order]).astype(np.uint32), src[order]


def _as_int(v, name):
    """The device clock is an integer: slot, freq, transtime and bandwidth must be integer-valued."""
    if isinstance(v, (bool, np.bool_)) or int(v) != v:
        raise NotImplementedError(f"{name}={v!r}: the device timing model is integer-valued (the reference's float "
                                  "clock would make route_time a float sum); there is no CPU fallback")
    return int(v)


def _floor_pow2(v):
    return 1 << (int(v).bit_length() - 1)


class Network:
    def __init__(self, file, bandwidth=1, transtime=1, is_drop=False, *, replicas=1, device=None,
                 rng="trace", seed=None, replica_base=0, queue_capacity=None,
                 replicas_per_block=0, threads_per_block=0, tables_in_smem=-1, dtype="float64",
                 page_records=0, pages_per_replica=None):
        bandwidth, transtime = _as_int(bandwidth, "bandwidth"), _as_int(transtime, "transtime")
        if bandwidth < 0 or transtime < 0:
            raise ValueError("bandwidth and transtime must be >= 0")
        if rng not in ("trace", "philox"):
            raise ValueError("rng must be 'trace' or 'philox'")
        if replicas < 1:
            raise ValueError("replicas must be >= 1")
        # Q-table STORAGE type: 'float64' is the reference's precision and the parity mode; 'float32' (Qroute
        # only) halves the table so that twice the replicas fit in an SM — rows are widened to fp64 when read and the TD
        # update is rounded once on store, so trajectories agree with fp64 until the first near-tie forks them
        self.dtype = np.dtype(dtype).name
        if self.dtype not in ("float64", "float32"):
            raise ValueError("dtype must be 'float64' or 'float32'")
        self.bandwidth, self.transtime, self.is_drop = bandwidth, transtime, is_drop
        self.replicas, self.rng, self.replica_base = int(replicas), rng, int(replica_base)
        self.read_network(file)
        self.nodes = OrderedDict((i, Node(i, self)) for i in self.links.keys())
        self.sample, self._sample_idx = [], 0
        # RNG streams.  trace: replica r replays the reference's legacy stream seeded `seed + r`; a
        # single un-seeded replica uses NumPy's GLOBAL stream like the reference itself (train.py:15).
        if rng == "trace":
            if seed is None and self.replicas == 1:
                self._rs = [np.random.mtrand._rand]
            else:
                base = 1 if seed is None else int(seed)
                self._rs = [np.random.RandomState(base + self.replica_base + r) for r in range(self.replicas)]
        self.seed = 0 if seed is None else int(seed)

        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("rlrouting_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        nat.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:           # 'cuda' without an index means the CURRENT device, not device 0
            self.device = torch.device("cuda", torch.cuda.current_device())
        topo, R, N = self.topology, self.replicas, self.topology.N
        if queue_capacity is None:
            # the reference's queues are unbounded; by default the rings get a fifth of the device memory (16-byte
            # records), at most 32,768 entries per queue
            budget = int(torch.cuda.get_device_properties(self.device).total_memory * 0.2) // 16
            queue_capacity = min(max(_floor_pow2(max(budget // (R * N), 1)), 64), 32768)
        if queue_capacity & (queue_capacity - 1):
            raise ValueError("queue_capacity must be a power of two")
        self.queue_capacity = int(queue_capacity)
        self._launch_cfg = (int(replicas_per_block), int(threads_per_block), int(tables_in_smem))
        dev = self.device
        # Paged queues (page_records > 0): the reference's unbounded lists (env.py:100,130) as linked pages of `page_records`
        # packets from a per-replica pool of `pages_per_replica` pages, instead of one ring of `queue_capacity` slots per
        # node — memory follows the packets alive in a replica, not the worst node of the run.  Served by train() /
        # run_device() with default options (the fused kernel); the call-by-call API needs rings.
        self.page_records = int(page_records)
        self._paged = self.page_records > 0
        if self._paged:
            if self.page_records < 4 or self.page_records & (self.page_records - 1):
                raise ValueError("page_records must be a power of two >= 4")
            if self._general_requested(bandwidth, transtime):
                raise NotImplementedError("paged queues do not serve the general timing model (transtime > 1 or bandwidth < 1)")
            if pages_per_replica is None:
                budget = int(torch.cuda.get_device_properties(self.device).total_memory * 0.2) // 16
                pages_per_replica = budget // (R * self.page_records)
            self.pages_per_replica = int(min(max(int(pages_per_replica), 2 * N + 1), 65535))
            NP, P = self.pages_per_replica, self.page_records
            self._pool = torch.empty((R, NP * P, 4), dtype=torch.int32, device=dev)
            self._page_next = torch.zeros((R, NP), dtype=torch.int16, device=dev)
            self._page_free = torch.zeros((R, NP), dtype=torch.int16, device=dev)
            self._page_top = torch.zeros(R, dtype=torch.int32, device=dev)
            self._q_pages = torch.zeros((R, N, 4), dtype=torch.int16, device=dev)
            self._paged_reset_host()
        self._ring = (torch.empty((1, 1, 2, 4), dtype=torch.int32, device=dev) if self._paged
                      else torch.empty((R, N, self.queue_capacity, 4), dtype=torch.int32, device=dev))
        self._q_head = torch.zeros((R, N), dtype=torch.int32, device=dev)
        self._q_tail = torch.zeros((R, N), dtype=torch.int32, device=dev)
        self._counters = torch.zeros((R, nat.NUM_COUNTERS), dtype=torch.int64, device=dev)
        self._status = torch.zeros((R, 2), dtype=torch.int32, device=dev)
        self._row_ptr = torch.from_numpy(topo.row_ptr).to(dev)
        self._col = torch.from_numpy(topo.col).to(dev)
        self._heap_pos = torch.from_numpy(np.ascontiguousarray(topo.heap_pos, dtype=np.uint8)).to(dev)
        self._stats = torch.zeros(nat.NUM_COUNTERS, dtype=torch.int64, device=dev)
        # General timing model: with transtime > 1 an event can outlive a step of slot 1, so the event
        # heap, the per-link in-flight counts and out-of-order queue removals become persistent device state.
        self._general = transtime > 1 or bandwidth < 1
        self._ecap = N * (transtime + 1) if self._general else 0
        if self._ecap > 65535:
            raise NotImplementedError(f"N * (transtime + 1) = {self._ecap} in-flight event slots exceed the device limit of 65535")
        if self._general:
            self._link_sent = torch.zeros((R, topo.sum_deg), dtype=torch.int32, device=dev)
            self._event_heap = torch.zeros((R, self._ecap), dtype=torch.int64, device=dev)
            self._event_count = torch.zeros(R, dtype=torch.int32, device=dev)
            self._event_rec = torch.zeros((R, self._ecap, 4), dtype=torch.int32, device=dev)
            self._q_dead = torch.zeros((R, N), dtype=torch.int32, device=dev)
        self._handles = []
        # Philox injection schedules are built one launch ahead on a side stream by default; False builds each one in-stream
        # right before the step kernel that consumes it (no kernel then runs beside the step kernel)
        self.schedule_prefetch = os.environ.get("RLR_SCHED_INLINE", "") != "1"
        self.launch_count = 0          # kernels of this library launched: step, injection-schedule and ratio kernels
                                       # (bench.py reports the count of its timed region as gpu_launches)
        self.last_h2d_bytes = self.last_d2h_bytes = 0
        self._resident = {}
        self.clock = 0
        self._agent = Policy(self)
        self.last_sendlog = None

    @staticmethod
    def _general_requested(bandwidth, transtime):
        return transtime > 1 or bandwidth < 1

    def _paged_reset_host(self):
        """Initial page lists (what rlr_reset does on the device once a handle exists): node x owns page x and spare N + x."""
        import torch
        N, NP, dev = self.topology.N, self.pages_per_replica, self.device
        ids = torch.arange(NP, device=dev)
        free = torch.where(ids < NP - 2 * N, ids + 2 * N, torch.zeros_like(ids)).to(torch.int16)
        self._page_free.copy_(free.unsqueeze(0).expand(self.replicas, NP))
        self._page_top.fill_(NP - 2 * N)
        qp = torch.zeros((N, 4), dtype=torch.int64, device=dev)
        qp[:, 0] = torch.arange(N, device=dev)
        qp[:, 1] = qp[:, 0]
        qp[:, 2] = qp[:, 0] + N
        self._q_pages.copy_(qp.to(torch.int16).unsqueeze(0).expand(self.replicas, N, 4))

    @property
    def queue_memory_bytes(self):
        """Device bytes that hold queued packets: the rings, or the page pool with its links and free stack."""
        if self._paged:
            return sum(t.numel() * t.element_size() for t in (self._pool, self._page_next, self._page_free, self._q_pages))
        return self._ring.numel() * 4

    def nodes_of(self, replica):
        """`Network.nodes` (env.py:244-246) of replica `replica`: an ordered dict of read-only Node views."""
        return _nodes_of(self, replica)

    # ------------------------------------------------------------------ topology (env.py:282-296)
    def read_network(self, file):
        self.file = file
        self.topology = Topology.from_text(nets.resolve(file))
        self.links = OrderedDict((k, list(v)) for k, v in self.topology.links.items())
        self.proj = self.topology.proj

    # ------------------------------------------------------------------ native handle plumbing
    def _stream_ptr(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _make_handle(self, agent):
        """One native handle per (network, agent): environment buffers from here, tables from the agent."""
        topo = self.topology
        t = nat.Topology(topo.N, topo.sum_deg, topo.max_deg, topo.row_ptr.ctypes.data, topo.col.ctypes.data)
        rpb, tpb, tis = self._launch_cfg
        cfg = nat.Config(self.replicas, self.replica_base, agent._kernel_policy, self.queue_capacity, rpb, tpb, tis,
                         self.device.index, self._ecap, 1 if (self.dtype == "float32" and agent._kernel_policy == nat.POLICY_QROUTE) else 0,
                         self.page_records, self.pages_per_replica if self._paged else 0)
        h = C.c_void_p()
        nat.check(nat.lib().rlr_create(C.byref(t), C.byref(cfg), C.byref(h)))
        d = agent._dev
        alias = getattr(agent, "_dev_alias", {})
        ptr = lambda name: d[alias.get(name, name)].data_ptr() if alias.get(name, name) in d else None  # noqa: E731
        b = nat.Buffers(ptr("Qtable"), ptr("Theta"), ptr("Trace"), self._ring.data_ptr(), self._q_head.data_ptr(),
                        self._q_tail.data_ptr(), self._counters.data_ptr(), self._status.data_ptr(), ptr("next_hop"),
                        ptr("distance"), ptr("choice"), self._row_ptr.data_ptr(), self._col.data_ptr(),
                        self._heap_pos.data_ptr(), ptr("gr_choice_mask"), ptr("gr_distance"), ptr("choice_mask"),
                        ptr("gr_qs_off"), *([t.data_ptr() for t in (self._link_sent, self._event_heap, self._event_count,
                                                                     self._event_rec, self._q_dead)]
                                             if self._general else [None] * 5),
                        *([t.data_ptr() for t in (self._pool, self._page_next, self._page_free, self._page_top, self._q_pages)]
                          if self._paged else [None] * 5))
        nat.check(nat.lib().rlr_bind_buffers(h, C.byref(b)))
        self._handles.append(h)
        return h

    def _scratch(self, tag, shape, dtype):
        """Device scratch buffers reused across calls (PyTorch-owned, lent to the C layer)."""
        import torch
        buf = self._resident.get(tag)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
            buf = self._resident[tag] = torch.empty(shape, dtype=dtype, device=self.device)
        return buf

    # ---- Philox injection schedules, built one chunk ahead on a side stream (double-buffered) ----
    def _schedule(self, h, rp, clock0, K, lam_key, lam_max, next_key=None):
        """Points rp at the injection schedule of steps [clock0, clock0+K) (Network.new_packet ahead of time).
        If the previous call prefetched it on the side stream the compute stream only waits for that event,
        otherwise it is built in-stream.  Call _schedule_done after the step kernel has been enqueued."""
        import torch
        st = self.__dict__.get("_sched")
        if st is None:
            st = self._sched = {"bufs": [None, None], "ready": [None, None], "used": [None, None],
                                "stream": torch.cuda.Stream(device=self.device)}
        cap = schedule_capacity(lam_max, self.topology.N, K)
        shape = (self.replicas, self.topology.N, cap)
        key = (clock0, K, lam_key, cap, int(rp.freq), int(rp.slot), int(rp.seed))     # everything the schedule kernel reads
        compute = torch.cuda.current_stream(self.device)
        slot = next((i for i in (0, 1) if st["ready"][i] is not None and st["ready"][i][0] == key), None)
        if slot is not None:
            compute.wait_event(st["ready"][slot][1])                 # prefetched by the previous call
        else:
            slot = 0
            if st["ready"][slot] is not None:                        # a stale prefetch may still be writing this slot
                compute.wait_event(st["ready"][slot][1])
            self._sched_buf(st, slot, shape)
            rp.inj_events, rp.inj_cap, rp.clock0 = st["bufs"][slot].data_ptr(), cap, clock0
            nat.check(nat.lib().rlr_build_inject_schedule(h, K, C.byref(rp), self._stream_ptr()))
            self.launch_count += 1
        st["ready"][slot] = None
        rp.inj_events, rp.inj_cap, rp.clock0 = st["bufs"][slot].data_ptr(), cap, clock0
        rp.inject_mode = nat.INJECT_TRACE                            # "events provided"
        self._sched_cur = (slot, next_key, lam_key, lam_max)

    def _sched_buf(self, st, slot, shape):
        import torch
        # One flat allocation per slot, grown when a schedule needs more (a call alternates between the shapes of its long
        # and its short launch: re-allocating on every shape change would synchronise the device twice per call).
        # The storage is allocated ON THE SIDE STREAM: the caching allocator hands a block that was just freed on the
        # compute stream (e.g. a per-launch metrics buffer whose ratio kernel is still queued) to the next compute-stream
        # allocation, which is safe in stream order there — but the prefetch writes the schedule from the side stream,
        # and that queued kernel would then scribble over it.
        n = int(np.prod(shape))
        store = st.setdefault("store", [None, None])
        if store[slot] is None or store[slot].numel() < n:
            if store[slot] is not None:
                torch.cuda.synchronize(self.device)                  # nothing may still use the old buffer
            with torch.cuda.stream(st["stream"]):
                store[slot] = torch.empty(n, dtype=torch.int32, device=self.device)
            store[slot].record_stream(torch.cuda.current_stream(self.device))   # ... and it is read on the compute stream
        st["bufs"][slot] = store[slot][:n].view(shape)
        return st["bufs"][slot]

    def _schedule_done(self, h, rp):
        """The step kernel reading the current slot is enqueued: remember that, then prefetch the next chunk's
        schedule into the other slot on the side stream so it overlaps with the step kernel."""
        import torch
        st = self._sched
        slot, next_key, lam_key, lam_max = self._sched_cur
        used = torch.cuda.Event()
        used.record(torch.cuda.current_stream(self.device))
        st["used"][slot] = used
        if next_key is None or not self.schedule_prefetch:
            return
        clock0, K = next_key
        other = 1 - slot
        cap = schedule_capacity(lam_max, self.topology.N, K)
        self._sched_buf(st, other, (self.replicas, self.topology.N, cap))
        side = st["stream"]
        lam_st = self._resident.get("lam")
        if lam_st is not None and lam_st.get("uploaded") is not None:
            side.wait_event(lam_st["uploaded"])                      # the load vectors were copied on the compute stream
        if st["used"][other] is not None:
            side.wait_event(st["used"][other])                       # the step kernel that last read that slot
        rp2 = nat.RunParams()
        C.memmove(C.byref(rp2), C.byref(rp), C.sizeof(nat.RunParams))
        rp2.inj_events, rp2.inj_cap, rp2.clock0 = st["bufs"][other].data_ptr(), cap, clock0
        # the side stream reads the load vectors, which were allocated on the compute stream: tell the caching allocator, so
        # that a Network dropped while its last (speculative) prefetch is still running does not hand them out again
        if lam_st is not None:
            lam_st["d_en"].record_stream(side)
            lam_st["d_thr"].record_stream(side)
        nat.check(nat.lib().rlr_build_inject_schedule(h, K, C.byref(rp2), C.c_void_p(side.cuda_stream)))
        self.launch_count += 1
        done = torch.cuda.Event()
        done.record(side)
        st["ready"][other] = ((clock0, K, lam_key, cap, int(rp.freq), int(rp.slot), int(rp.seed)), done)

    def _lam_tensors(self, lam, upload=True):
        """exp(-lambda_r / N) and the exact integer injection thresholds as device tensors.  The host values are cached
        per load vector in page-locked staging buffers; `upload` copies them to the device again (they are the inputs of
        a train call), otherwise the resident copies are reused.  Returns (key, enlam, thr, bytes uploaded)."""
        import torch
        key = lam.tobytes() if lam.strides[0] else float(lam[0])
        st = self._resident.get("lam")
        fresh = st is None or st["key"] != key
        if fresh:
            if st is None:
                st = self._resident["lam"] = {
                    "h_en": torch.empty(self.replicas, dtype=torch.float64).pin_memory(),
                    "h_thr": torch.empty(self.replicas, dtype=torch.int64).pin_memory(),
                    "d_en": torch.empty(self.replicas, dtype=torch.float64, device=self.device),
                    "d_thr": torch.empty(self.replicas, dtype=torch.int64, device=self.device)}
            en_host = np.exp(-lam / np.float64(self.topology.N))
            st["h_en"].numpy()[:] = en_host
            st["h_thr"].numpy()[:] = injection_threshold(en_host).view(np.int64)
            st["key"] = key
        sent = 0
        if fresh or upload:
            st["d_en"].copy_(st["h_en"], non_blocking=True)
            st["d_thr"].copy_(st["h_thr"], non_blocking=True)
            st["uploaded"] = torch.cuda.Event()                  # the side stream that prefetches schedules waits for it
            st["uploaded"].record(torch.cuda.current_stream(self.device))
            sent = self.replicas * 16
        return key, st["d_en"], st["d_thr"], sent

    def _copy_stream(self):
        import torch
        if getattr(self, "_copier", None) is None:
            self._copier = torch.cuda.Stream(device=self.device)
        return self._copier

    def launch_info(self):
        h = getattr(self._agent, "_handle", None)
        if h is None:
            return None
        v = [C.c_int32() for _ in range(4)]
        nat.check(nat.lib().rlr_launch_info(h, *[C.byref(x) for x in v]))
        return {"threads_per_block": v[0].value, "replicas_per_block": v[1].value, "smem_bytes": v[2].value,
                "tables_in_smem": bool(v[3].value)}

    def __del__(self):
        try:
            for h in self._handles:
                nat.lib().rlr_destroy(h)
        except Exception:
            pass

    # ------------------------------------------------------------------ agent binding (env.py:252-265)
    @property
    def agent(self):
        return self._agent

    @property
    def mode(self):
        return self._agent.mode

    @agent.setter
    def agent(self, new_agent):
        if type(new_agent) is Policy:
            self._agent = new_agent
            return
        if getattr(new_agent, "_kernel_policy", None) is None or getattr(new_agent, "_handle", None) is None:
            raise NotImplementedError(f"{type(new_agent).__name__} is not one of the device policies (Shortest, Qroute, "
                                      "HybridQ, MaHybridQ); custom Python policies would need a CPU path, which this "
                                      "package does not have")
        if new_agent.mode is not None and not (new_agent.mode == "dual" and new_agent._kernel_policy == nat.POLICY_CDRQ):
            raise NotImplementedError(f"mode {new_agent.mode!r} is not implemented on the device (SURVEY §8 rows f1/f2)")
        if new_agent._network is not self:
            raise ValueError("the agent was constructed for a different Network")
        self._agent = new_agent

    def bind(self, agent):
        """Legacy spelling used by the reference's train.py:39."""
        self.agent = agent

    # ------------------------------------------------------------------ reset (env.py:267-280)
    def reset(self):
        self.clock = 0
        self._desync = False
        self._pending = None                    # (per-replica lists are only created by inject())
        self._last_rewards = None
        h = getattr(self._agent, "_handle", None)
        if hasattr(self._agent, "_before_network_reset"):
            self._agent._before_network_reset()
        if h is not None:
            nat.check(nat.lib().rlr_reset(h, self._stream_ptr()))
        else:
            for t in (self._q_head, self._q_tail, self._counters, self._status):
                t.zero_()
            if self._paged:
                self._paged_reset_host()
            if self._general:
                for t in (self._link_sent, self._event_count, self._q_dead):
                    t.zero_()
        self._agent.clean()

    def clean(self):
        """Legacy spelling used by the reference's train.py:51."""
        self.reset()

    # ------------------------------------------------------------------ checkpoint / resume
    # The reference can pickle an agent's tables (Policy.store / load) but not the environment; these snapshots hold
    # both, so a long run can be stopped and continued bit for bit — also on another GPU, or with another ring capacity.
    def snapshot(self):
        """Environment and bound-agent state as a dict of host arrays: clock, counters, status words, every queue's
        packets (head first), events in flight and busy links (general timing model), the agent's device tables and
        the trace-mode RNG states."""
        import torch
        self._check_sync()
        R, N, cap, dev = self.replicas, self.topology.N, self.queue_capacity, self.device
        head = self._q_head.reshape(-1).to(torch.int64) & 0xFFFFFFFF
        lens = (self._q_tail.reshape(-1).to(torch.int64) - head) & 0xFFFFFFFF          # ring entries, tombstones included
        qid = torch.repeat_interleave(torch.arange(R * N, device=dev), lens)
        j = torch.arange(int(lens.sum().item()), device=dev) - (torch.cumsum(lens, 0) - lens)[qid]
        if self._paged:
            recs = self._pool[qid // N, self._paged_record_index(qid, j, head)]
        else:
            recs = self._ring.view(R * N, cap, 4)[qid, (head[qid] + j) & (cap - 1)]
        snap = {"version": 1, "n_nodes": N, "sum_deg": self.topology.sum_deg, "replicas": R,
                "replica_base": self.replica_base, "rng": self.rng, "seed": self.seed, "clock": int(self.clock),
                "bandwidth": self.bandwidth, "transtime": self.transtime, "is_drop": bool(self.is_drop),
                "queue_len": lens.cpu().numpy().reshape(R, N), "queue_rec": recs.cpu().numpy(),
                "counters": self._counters.cpu().numpy(), "status": self._status.cpu().numpy(),
                "pending": [list(pk) for pk in self.__dict__.get("_pending") or []],
                "agent_class": type(self._agent).__name__,
                "agent": {k: v.cpu().numpy() for k, v in self._agent._dev.items()}}
        if self._general:
            snap["general"] = {k: getattr(self, "_" + k).cpu().numpy()
                               for k in ("link_sent", "event_heap", "event_count", "event_rec", "q_dead")}
        if self.rng == "trace":
            snap["rng_state"] = [rs.get_state() for rs in self._rs]
        return snap

    def restore(self, snap):
        """Loads a `snapshot()` into this network and its bound agent (same topology, replica count, timing model and
        agent class); the run then continues exactly as the snapshotted one would have."""
        import torch
        R, N, cap, dev = self.replicas, self.topology.N, self.queue_capacity, self.device
        want = (N, self.topology.sum_deg, R, self.replica_base, self.rng, self.bandwidth, self.transtime,
                bool(self.is_drop), type(self._agent).__name__)
        got = (snap["n_nodes"], snap["sum_deg"], snap["replicas"], snap["replica_base"], snap["rng"], snap["bandwidth"],
               snap["transtime"], bool(snap.get("is_drop", False)), snap["agent_class"])
        if snap.get("version") != 1 or want != got:
            raise ValueError(f"snapshot does not fit this network: {got} vs {want}")
        lens = torch.from_numpy(np.ascontiguousarray(snap["queue_len"], dtype=np.int64)).reshape(-1).to(dev)
        if not self._paged and int(lens.max().item()) > cap:
            raise OverflowError(f"snapshot holds a queue of {int(lens.max().item())} entries; queue_capacity is {cap}")
        qid = torch.repeat_interleave(torch.arange(R * N, device=dev), lens)
        j = torch.arange(int(lens.sum().item()), device=dev) - (torch.cumsum(lens, 0) - lens)[qid]
        recs = torch.from_numpy(np.ascontiguousarray(snap["queue_rec"], dtype=np.int32)).to(dev)
        if self._paged:
            # every queue gets len // P + 1 consecutive pages of its replica (the page under the tail always exists), then
            # the N spares, and what is left goes onto the free stack
            P, NP = self.page_records, self.pages_per_replica
            npg = (lens // P + 1).reshape(R, N)
            start = torch.cumsum(npg, 1) - npg
            used = npg.sum(1)
            if int(used.max().item()) + N > NP:
                raise OverflowError(f"snapshot needs {int(used.max().item()) + N} pages per replica; pages_per_replica is {NP}")
            startf = start.reshape(-1)
            self._pool[qid // N, (startf[qid] + j // P) * P + j % P] = recs
            ids = torch.arange(NP, device=dev).unsqueeze(0).expand(R, NP)
            self._page_next.copy_((ids + 1).to(torch.int16))             # consecutive pages: next = id + 1 inside a queue
            qp = torch.zeros((R, N, 4), dtype=torch.int64, device=dev)
            qp[:, :, 0] = start
            qp[:, :, 1] = start + npg - 1
            qp[:, :, 2] = used.unsqueeze(1) + torch.arange(N, device=dev).unsqueeze(0)
            self._q_pages.copy_(qp.to(torch.int16))
            nfree = NP - used - N
            free = used.unsqueeze(1) + N + ids
            self._page_free.copy_(torch.where(ids < nfree.unsqueeze(1), free, torch.zeros_like(free)).to(torch.int16))
            self._page_top.copy_(nfree.to(torch.int32))
        else:
            self._ring.view(R * N, cap, 4)[qid, j] = recs
        self._q_head.zero_()
        self._q_tail.copy_(lens.reshape(R, N).to(torch.int32))
        self._counters.copy_(torch.from_numpy(np.ascontiguousarray(snap["counters"])))
        self._status.copy_(torch.from_numpy(np.ascontiguousarray(snap["status"])))
        if self._general:
            for k, v in snap["general"].items():
                getattr(self, "_" + k).copy_(torch.from_numpy(np.ascontiguousarray(v)))
        for k, v in snap["agent"].items():
            self._agent._dev[k].copy_(torch.from_numpy(np.ascontiguousarray(v)))
        if self.rng == "trace":
            for rs, st in zip(self._rs, snap["rng_state"]):
                rs.set_state(st)
        self.seed = int(snap["seed"])                               # Philox key of the runs that follow
        self.clock = int(snap["clock"])
        self._pending = [list(map(tuple, pk)) for pk in snap["pending"]] or None
        self._desync, self._last_rewards = False, None
        if self.__dict__.get("_sched") is not None:                 # injection schedules prefetched for the old state
            torch.cuda.synchronize(dev)
            self._sched["ready"] = [None, None]

    def save_state(self, filename):
        """`snapshot()` pickled to a file."""
        import pickle
        with open(filename, "wb") as f:
            pickle.dump(self.snapshot(), f, protocol=4)

    def load_state(self, filename):
        import pickle
        with open(filename, "rb") as f:
            self.restore(pickle.load(f))

    # ------------------------------------------------------------------ the loop's pieces, one call at a time
    # Network.train fuses these on the device; they are also available separately with the reference's
    # semantics (one kernel launch per call, so this path is for inspection and custom loops, not for speed).
    def new_packet(self, lambd):
        """env.py:298-312.  One replica: a list of Packet; R replicas: a list of R lists."""
        import torch
        R, N = self.replicas, self.topology.N
        lam = np.broadcast_to(np.asarray(lambd, dtype=np.float64), (R,))
        out = []
        if self.rng == "trace":
            for r in range(R):
                cnt, pairs = trace.draw_injections(self._rs[r], N, 1, lam[r])
                out.append([Packet(int(sv), int(dv), self.clock) for sv, dv in pairs])
        else:
            h = self._handle_or_raise()
            cap = schedule_capacity(lam.max(), N, 1) + 8
            rp = nat.RunParams()
            en_host = np.exp(-lam / np.float64(N))
            enlam = torch.from_numpy(en_host).to(self.device)
            thr = torch.from_numpy(injection_threshold(en_host).view(np.int64)).to(self.device)
            buf = torch.empty((R, N, cap), dtype=torch.int32, device=self.device)
            rp.enlam, rp.inj_thr, rp.seed, rp.clock0 = enlam.data_ptr(), thr.data_ptr(), self.seed, self.clock
            rp.inj_events, rp.inj_cap = buf.data_ptr(), cap
            nat.check(nat.lib().rlr_build_inject_schedule(h, 1, C.byref(rp), self._stream_ptr()))
            ev = buf.cpu().numpy().view(np.uint32)
            self._raise_on_status(self._status.cpu().numpy())
            for r in range(R):
                pk = []
                for x in range(N):
                    for e in ev[r, x]:
                        if e == 0xFFFFFFFF:
                            break
                        pk.append(Packet(x, int(e & 0xFF), self.clock))
                out.append(pk)
        return out[0] if R == 1 else out

    def inject(self, packets):
        """env.py:314-319.  The packets join their source queues at the next `step` (same clock, so identical state)."""
        per = [packets] if self.replicas == 1 else packets
        if len(per) != self.replicas:
            raise ValueError("inject expects one packet list per replica")
        pend = self.__dict__.get("_pending")
        if pend is None:
            pend = self._pending = [[] for _ in range(self.replicas)]
        for r, pk in enumerate(per):
            for q in pk:
                if q.source == q.dest:
                    raise NotImplementedError("a packet whose source is its destination ends at injection in the "
                                              "reference (env.py:124-126); not supported on the device")
                pend[r].append((int(q.source), int(q.dest)))

    def _handle_or_raise(self):
        h = getattr(self._agent, "_handle", None)
        if h is None:
            raise NotImplementedError("bind one of the device policies first (nw.agent = Qroute(nw) ...)")
        return h

    def _timing(self, rp, slot=1, freq=1):
        """Fills the timing block of the run parameters (env.py:237-239 bandwidth / transtime, env.py:372-373 slot / freq)."""
        slot, freq = _as_int(slot, "slot"), _as_int(freq, "freq")
        if slot < 1 or freq < 1:
            raise ValueError("slot and freq must be >= 1")
        rp.slot, rp.freq, rp.transtime, rp.bandwidth, rp.general = slot, freq, self.transtime, self.bandwidth, int(self._general)
        return slot, freq

    def _run_params(self, agent, lr, merged=False):
        """Run parameters every entry point shares: learning rates (the agent's defaults overridden by `lr`; `merged`: `lr`
        already is the merged dict), the agent's hyper-parameters, Network options, seed and the default timing block."""
        lrd = dict(lr) if merged else dict(getattr(agent, "_lr_default", {}))
        if not merged:
            lrd.update(lr or {})
        rp = nat.RunParams()
        rp.add_entropy = int(bool(getattr(agent, "add_entropy", True)))
        rp.lr_q, rp.lr_p, rp.lr_e = float(lrd.get("q", 0.1)), float(lrd.get("p", 0.1)), float(lrd.get("e", 0.1))
        rp.discount = float(getattr(agent, "discount", 0.99))
        rp.discount_trace = float(getattr(agent, "discount_trace", 0.6))
        rp.conf_decay = float(getattr(agent, "decay", 0.9))
        rp.is_drop = int(bool(self.is_drop))
        rp.shortest_random = int(agent._kernel_policy in (nat.POLICY_SHORTEST, nat.POLICY_GLOBALROUTE) and bool(getattr(agent, "random", False)))
        rp.multiway = int(agent._kernel_policy == nat.POLICY_GLOBALROUTE and bool(getattr(agent, "multiway", False)))
        rp.seed = self.seed
        self._timing(rp)
        return rp

    def step(self, duration=1):
        """env.py:333-365: one send phase + event drain, nothing learned.  Returns the List[Reward] (one replica)
        or a list of R lists, to be passed to `agent.learn`."""
        import torch
        duration = _as_int(duration, "duration")
        self._no_paged("Network.step (the split-phase kernel)")
        h = self._handle_or_raise()
        R, N = self.replicas, self.topology.N
        pend = self.__dict__.get("_pending") or [[] for _ in range(R)]
        per_node = np.zeros((R, N), dtype=np.int64)
        for r in range(R):
            for sv, _ in pend[r]:
                per_node[r, sv] += 1
        cap = int(per_node.max()) + 2
        sched = np.full((R, N, cap), 0xFFFFFFFF, dtype=np.uint32)
        fill = np.zeros((R, N), dtype=np.int64)
        for r in range(R):
            for sv, dv in pend[r]:
                sched[r, sv, fill[r, sv]] = dv              # local step 0
                fill[r, sv] += 1
        d_sched = torch.from_numpy(sched.view(np.int32)).to(self.device)
        rew = self._scratch("rewards", (R, N, 16), torch.int32)
        rp = self._run_params(self._agent, {})
        self._timing(rp, duration, 1)
        rp.inject_mode, rp.inj_events, rp.inj_cap = nat.INJECT_TRACE, d_sched.data_ptr(), cap
        rp.clock0, rp.phase, rp.rewards = self.clock, 1, rew.data_ptr()
        nat.check(nat.lib().rlr_run_steps(h, 1, C.byref(rp), self._stream_ptr()))
        self.launch_count += 1
        self.clock += duration
        self._last_duration = duration
        self._pending = None
        raw = rew.cpu().numpy()
        self._raise_on_status(self._status.cpu().numpy())
        out = []
        for r in range(R):
            lst = []
            for x in range(N):
                rec = raw[r, x]
                meta = int(rec[4]) & 0xFFFFFFFF
                if not meta & 1:
                    continue
                y, dest = (meta >> 9) & 0xFF, meta >> 17
                w0 = int(rec[0]) & 0xFFFFFFFF
                pkt = Packet(w0 >> 16, w0 & 0xFFFF, int(rec[2]), int(rec[1]), int(rec[3]))
                mq = rec[8:16].view(np.float64)
                pol = self._agent._kernel_policy
                info = {"q_y": int(rec[5]), "t_y": 1}
                if pol in (nat.POLICY_QROUTE, nat.POLICY_HYBRIDQ, nat.POLICY_MAHYBRIDQ):
                    info["max_Q_y"] = float(mq[0])
                if pol in (nat.POLICY_HYBRIDQ, nat.POLICY_MAHYBRIDQ):
                    info["max_Q_x_d"] = float(mq[1])
                if pol in (nat.POLICY_CQ, nat.POLICY_CDRQ):                     # qroute.py:72-78, 107-115
                    info["max_Q_f"], info["C_f"] = float(mq[0]), float(mq[1])
                if pol == nat.POLICY_CDRQ:                                      # env.py:154-160 ('dual' mode)
                    info.update({"t_y": 0, "q_x": int(rec[6]), "t_x": 0, "C_b": float(mq[2]), "max_Q_b": float(mq[3])})
                lst.append(Reward(x, pkt, int(y), info))
            out.append(lst)
        result = out[0] if R == 1 else out
        self._last_rewards = (result, self.clock)
        return result

    def _learn(self, agent, rewards, lr):
        """agent.learn(rewards, lr) for a device policy: learns from the rewards of the latest `step`."""
        import torch
        last = self.__dict__.get("_last_rewards")
        if last is None or last[0] is not rewards or last[1] != self.clock or agent is not self._agent:
            raise NotImplementedError("a device policy can only learn from the list returned by the latest Network.step() "
                                      "of the network it is bound to (the rewards themselves stay on the device)")
        h = self._handle_or_raise()
        R, N = self.replicas, self.topology.N
        empty = self._resident.get("empty_sched")
        if empty is None:
            empty = self._resident["empty_sched"] = torch.full((R, N, 2), -1, dtype=torch.int32, device=self.device)
        rp = self._run_params(agent, lr)
        rp.inject_mode, rp.inj_events, rp.inj_cap = nat.INJECT_TRACE, empty.data_ptr(), 2
        self._timing(rp, self.__dict__.get("_last_duration", 1), 1)
        rp.clock0, rp.phase, rp.rewards = self.clock - rp.slot, 2, self._scratch("rewards", (R, N, 16), torch.int32).data_ptr()
        nat.check(nat.lib().rlr_run_steps(h, 1, C.byref(rp), self._stream_ptr()))
        self.launch_count += 1
        self._last_rewards = None
        self._raise_on_status(self._status.cpu().numpy())

    def sample_route_time(self, size, lambd, slot=1, freq=1, lr={}):
        """env.py:416-438: runs like `train` until `size` packets have been delivered and returns their routing times
        in delivery order (one replica: [size]; R replicas: [R, size], each replica stopping at its own clock).
        After a multi-replica call the replicas' clocks differ, so the network must be `reset()` before it is used
        again."""
        import torch
        self._check_sync()
        self._no_paged("sample_route_time")
        agent = self._agent
        h = self._handle_or_raise()
        R, N, dev = self.replicas, self.topology.N, self.device
        size = int(size)
        rp = self._run_params(agent, lr)
        slot, freq = self._timing(rp, slot, freq)
        lam = np.broadcast_to(np.asarray(lambd, dtype=np.float64), (R,)) * slot          # env.py:427
        samp = torch.zeros((R, size), dtype=torch.int32, device=dev)
        sidx = torch.zeros(R, dtype=torch.int32, device=dev)
        sclk = torch.zeros(R, dtype=torch.int32, device=dev)
        rp.sample, rp.sample_idx, rp.sample_clock, rp.sample_size = samp.data_ptr(), sidx.data_ptr(), sclk.data_ptr(), size
        keep = []
        if self.rng == "philox":
            lam_key = lam.tobytes() if lam.strides[0] else float(lam[0])
            en_host = np.exp(-lam / np.float64(N))
            enlam = torch.from_numpy(en_host).to(dev)
            thr = torch.from_numpy(injection_threshold(en_host).view(np.int64)).to(dev)
            rp.enlam, rp.inj_thr = enlam.data_ptr(), thr.data_ptr()
            keep += [enlam, thr]
        # trace mode advances one iteration per launch so the host RNG stream stops exactly where the reference's would
        K = freq * (1 if self.rng == "trace" else 128)                  # Network.step calls per launch
        while True:
            if self.rng == "trace":
                traces = trace.draw_injections_many(self._rs, N, 1, lam)
                evs = [trace_schedule(c, pr, N, 0, 1) for c, pr in traces]
                per_node = [np.bincount(node, minlength=N) for _, node in evs]
                cap = int(max(pn.max() for pn in per_node)) + 2
                sched = np.full((R, N, cap), 0xFFFFFFFF, dtype=np.uint32)
                for r, ((e, node), pn) in enumerate(zip(evs, per_node)):
                    start = np.concatenate([[0], np.cumsum(pn)[:-1]])
                    sched[r, node, np.arange(len(e)) - start[node]] = e
                d_sched = torch.from_numpy(sched.view(np.int32)).to(dev)
                rp.inject_mode, rp.inj_events, rp.inj_cap, rp.clock0 = nat.INJECT_TRACE, d_sched.data_ptr(), cap, self.clock
            else:
                self._schedule(h, rp, self.clock, K, lam_key, lam.max(), (self.clock + K * slot, K))
            nat.check(nat.lib().rlr_run_steps(h, K, C.byref(rp), self._stream_ptr()))
            self.launch_count += 1
            if self.rng == "philox":
                self._schedule_done(h, rp)
            self.clock += K * slot
            self._raise_on_status(self._status.cpu().numpy())
            if bool((sidx >= size).all().item()):
                break
        clocks = sclk.cpu().numpy().astype(np.int64)
        self.clock = int(clocks.max())
        self._desync = bool((clocks != clocks[0]).any())
        out = samp.cpu().numpy().astype(np.float64)
        return out[0] if R == 1 else out

    def _no_paged(self, what):
        if self._paged:
            raise NotImplementedError(f"{what} is not available with paged queues (page_records > 0): they are served by the "
                                      "fused train() / run_device() kernel; build the Network with rings for this call")

    def _check_sync(self):
        if self.__dict__.get("_desync"):
            raise RuntimeError("the replicas stopped at different clocks in sample_route_time; call reset() first")
        if any(self.__dict__.get("_pending") or []):
            raise NotImplementedError("packets passed to inject() are still waiting for the next step(): train(), "
                                      "sample_route_time() and run_device() draw their own traffic and would leave them "
                                      "behind; call step() first (the reference enqueues them immediately, env.py:314-319)")

    # ------------------------------------------------------------------ the hot path (env.py:367-414)
    def train(self, duration, lambd, slot=1, freq=1, lr={}, droprate=False, hop=False, *, lrq=None, lrp=None,
              per_step=None, sendlog=False, steps_per_launch=None, level_index=None, next_clock0=None):
        """Runs int(duration/slot) steps of inject -> step -> learn on the device.

        Returns the reference's dict (`route_time` = cumulative mean delivery time after each step,
        optional `droprate`, `hop`), each a vector of length steps for one replica or [R, steps] for
        R replicas.  `lambd` may be a scalar or a length-R vector (one load per replica).
        The legacy call `train(T, lambd=l, lrq=.., lrp=..)` (reference train.py:52) returns the tuple
        (route_time, drop_rate) instead.

        `level_index` (length R, values in [0, L)) groups the replicas into L load levels: the per-step vectors are then
        pooled per level ON THE DEVICE (rlr_reduce_levels: sum of route_time / sum of end_packets over the level's
        replicas, i.e. env.py:440-446 for the level as one population) and come back as [L, steps] — what a load sweep
        needs, at L/R of the device->host traffic of the per-replica vectors.
        `next_clock0`: clock at which the next train() call will start if not where this one ends (0 when every call is a
        fresh episode after `reset()`); only steers which injection schedule is prefetched.
        """
        import time
        import torch
        self._check_sync()
        _t = [time.perf_counter()]
        agent = self._agent
        h = getattr(agent, "_handle", None)
        if h is None:
            raise NotImplementedError("bind one of the device policies first (nw.agent = Qroute(nw) ...)")
        legacy = lrq is not None or lrp is not None
        lrd = dict(getattr(agent, "_lr_default", {}))
        lrd.update(lr or {})
        if lrq is not None:
            lrd["q"] = lrq
        if lrp is not None:
            lrd["p"] = lrp
        if self._paged and (sendlog or self.is_drop or _as_int(slot, "slot") != 1 or _as_int(freq, "freq") != 1 or self.transtime != 1):
            self._no_paged("train() with sendlog, is_drop, slot, freq or transtime other than the defaults")
        T, R, N, dev = int(duration / _as_int(slot, "slot")), self.replicas, self.topology.N, self.device
        lam = np.broadcast_to(np.asarray(lambd, dtype=np.float64), (R,)) * slot         # Poisson(lambd * slot), env.py:401
        d_lv, L = None, 0
        if level_index is not None:
            lv = np.ascontiguousarray(np.broadcast_to(np.asarray(level_index), (R,)), dtype=np.int32)
            if lv.min() < 0:
                raise ValueError("level_index must be >= 0")
            L = int(lv.max()) + 1
            if droprate and self.is_drop:
                raise NotImplementedError("per-level droprate vectors are not implemented; use the per-replica vectors")
            cached = self._resident.get("level_index")
            if cached is None or not np.array_equal(cached[0], lv):
                cached = self._resident["level_index"] = (lv, torch.from_numpy(np.array(lv)).to(self.device))
            d_lv = cached[1]
            per_step = True
        if per_step is None:
            per_step = R * T <= (1 << 26)
        result = {}
        W = L if d_lv is not None else R                       # width of the per-step vectors that go to the host
        if T == 0:
            result["route_time"] = np.zeros((W, 0))
            if droprate:
                result["droprate"] = np.zeros((W, 0))
            if hop:
                result["hop"] = np.zeros((W, 0))
            return self._finish(result, legacy, R if d_lv is None else 0)

        rp = self._run_params(agent, lrd, merged=True)
        slot, freq = self._timing(rp, slot, freq)
        keep = []
        h2d = 0
        traces = None
        if self.rng == "trace":
            rp.inject_mode = nat.INJECT_TRACE
            traces = trace.draw_batch(self._rs, N, T, lam)
        else:
            rp.inject_mode = nat.INJECT_PHILOX
            lam_key, enlam, thr, sent = self._lam_tensors(lam)
            rp.enlam, rp.inj_thr = enlam.data_ptr(), thr.data_ptr()
            keep += [enlam, thr]
            h2d += sent

        # Launch in chunks and stream each chunk's per-step vectors to pinned host memory on a side stream
        # while the next chunk computes, so the device->host copy of [R, T] doubles hides behind the kernel.
        # Every launch re-stages the tables in shared memory, so the default is two launches, 7/8 + 1/8 of the run:
        # only the copy of the short last chunk is exposed.
        cap = max(1, (1 << 25) // (R * freq))                   # per-launch device buffers stay under ~1 GB
        if steps_per_launch is not None:
            cap = max(1, int(steps_per_launch))
        elif not per_step:
            cap = T
        if self._general:                                       # event times are 16-bit offsets from the launch's clock
            cap = max(1, min(cap, (65535 - self.transtime) // (slot * freq)))
        plan = [cap] * (T // cap) + ([T % cap] if T % cap else [])
        if steps_per_launch is None and per_step and plan[-1] >= 64 and d_lv is None:   # (level curves are tiny: one launch)
            last = plan.pop()
            short = max(last // 8, 32)                          # long enough to cover the copy of the long part
            plan += [last - short, short]
        starts = np.concatenate([[0], np.cumsum(plan)]).astype(np.int64)
        _t.append(time.perf_counter())
        compute = torch.cuda.current_stream(dev)
        copier = self._copy_stream()
        # results land in pooled page-locked buffers handed out zero-copy (rlrouting_b200.pinned)
        host_rt, out_rt = pinned.POOL.take((T, W)) if per_step else (None, None)       # step-major
        host_hop, out_hop = pinned.POOL.take((T, W)) if (per_step and hop) else (None, None)
        want_drop = bool(per_step and droprate and self.is_drop)
        host_drop, out_drop = pinned.POOL.take((T, R)) if want_drop else (None, None)
        log_chunks = []
        stream = self._stream_ptr()
        for ci, K in enumerate(plan):
            s0 = int(starts[ci])
            if traces is not None:                      # host replay of env.py:307-310 -> per-node event lists
                sched = traces.schedule(s0, s0 + K, freq)
                cap = sched.shape[2]
                d_sched = torch.from_numpy(sched.view(np.int32)).to(dev)
                h2d += sched.nbytes
                rp.inj_events, rp.inj_cap = d_sched.data_ptr(), cap
                keep.append(d_sched)
                rp.clock0 = self.clock + s0 * freq * slot
            else:                                       # device Philox schedule, prefetched one chunk ahead
                # the schedule of the next launch is built on the side stream while this one runs; after the last launch
                # of the call, speculate that the next train() call repeats this one (same duration and loads)
                nxt = ((self.clock + (s0 + K) * freq * slot, plan[ci + 1] * freq) if ci + 1 < len(plan)
                       else (self.clock + T * freq * slot if next_clock0 is None else int(next_clock0), plan[0] * freq))
                self._schedule(h, rp, self.clock + s0 * freq * slot, K * freq, lam_key, lam.max(), nxt)
            KS = K * freq                               # Network.step calls of this launch (K iterations of the loop)
            # per-chunk device buffers are kept across calls (a fresh 32 MB torch.empty costs ~0.2 ms each); the call
            # ends with copier.synchronize(), so nothing still reads them when the next call writes them
            keepbuf = len(plan) <= 2                    # long runs in many launches allocate per launch instead
            d_rec = d_rt = None
            if per_step and keepbuf:
                d_rec = self._scratch(("rec", ci), (KS, R, 2), torch.int64)
                d_rt = self._scratch(("rt", ci), (KS, W), torch.float64)
            elif per_step:
                d_rec = torch.empty((KS, R, 2), dtype=torch.int64, device=dev)
                d_rt = torch.empty((KS, W), dtype=torch.float64, device=dev)
            rp.metrics = d_rec.data_ptr() if d_rec is not None else None
            d_hop = torch.empty((KS, W), dtype=torch.float64, device=dev) if host_hop is not None else None
            # per-replica vectors: the ratio kernel runs behind the step kernel; per-level vectors: rlr_reduce_levels below
            rp.route_time_out = d_rt.data_ptr() if (d_rt is not None and d_lv is None) else None
            rp.hop_out = d_hop.data_ptr() if (d_hop is not None and d_lv is None) else None
            d_rec2 = torch.empty((KS, R, 2), dtype=torch.int32, device=dev) if want_drop else None
            d_drop = torch.empty((KS, R), dtype=torch.float64, device=dev) if want_drop else None
            rp.metrics2 = d_rec2.data_ptr() if want_drop else None
            rp.droprate_out = d_drop.data_ptr() if want_drop else None
            log = torch.empty((R, KS, N), dtype=torch.int32, device=dev) if sendlog else None
            rp.sendlog = log.data_ptr() if log is not None else None
            nat.check(nat.lib().rlr_run_steps(h, KS, C.byref(rp), stream))
            if d_lv is not None:
                nat.check(nat.lib().rlr_reduce_levels(h, d_rec.data_ptr(), KS, d_lv.data_ptr(), L, d_rt.data_ptr(),
                                                      d_hop.data_ptr() if d_hop is not None else None, None, stream))
            if freq > 1 and per_step:                   # the vectors hold one entry per iteration (env.py:409-413)
                d_rt = d_rt[freq - 1::freq].contiguous()
                d_hop = d_hop[freq - 1::freq].contiguous() if d_hop is not None else None
                d_drop = d_drop[freq - 1::freq].contiguous() if d_drop is not None else None
            self.launch_count += 2 if per_step else 1           # the step kernel and, with per-step vectors, the ratio kernel
            if traces is None:
                self._schedule_done(h, rp)
            if per_step:
                done = torch.cuda.Event()
                done.record(compute)
                copier.wait_event(done)
                with torch.cuda.stream(copier):
                    host_rt[s0:s0 + K].copy_(d_rt, non_blocking=True)          # contiguous rows on both sides
                    if freq > 1 or not keepbuf:
                        d_rt.record_stream(copier)
                    if d_hop is not None:
                        host_hop[s0:s0 + K].copy_(d_hop, non_blocking=True)
                        d_hop.record_stream(copier)
                    if d_drop is not None:
                        host_drop[s0:s0 + K].copy_(d_drop, non_blocking=True)
                        d_drop.record_stream(copier)
            if log is not None:
                log_chunks.append(log)
        self.clock += T * freq * slot
        self._last_slot = slot
        _t.append(time.perf_counter())
        status = self._status.cpu().numpy()                     # device -> host: synchronises the compute stream
        _t.append(time.perf_counter())
        copier.synchronize()
        _t.append(time.perf_counter())
        self.last_timing = {"setup_ms": 1e3 * (_t[1] - _t[0]), "issue_ms": 1e3 * (_t[2] - _t[1]),
                            "wait_compute_ms": 1e3 * (_t[3] - _t[2]), "wait_copy_ms": 1e3 * (_t[4] - _t[3])}
        self._raise_on_status(status)
        self.last_h2d_bytes = h2d
        self.last_d2h_bytes = status.nbytes + (W * T * 8 * (2 if host_hop is not None else 1) if per_step
                                               else R * nat.NUM_COUNTERS * 8) + (R * T * N * 4 if sendlog else 0)
        if per_step:
            result["route_time"] = out_rt.T                                    # [R, T] view; env.py:409, 444-446
            if hop:
                result["hop"] = out_hop.T                                      # env.py:412-413, 440-442
        else:
            c = self._counters.cpu().numpy()
            rt, end, hops = c[:, 4:5], c[:, 1:2], c[:, 3:4]
            result["route_time"] = np.where(end > 0, rt / np.maximum(end, 1), 0.0)
            if hop:
                result["hop"] = np.where(end > 0, hops / np.maximum(end, 1), 0.0)
        if droprate:                                                                    # env.py:410-411, 448-450
            if want_drop:
                result["droprate"] = out_drop.T
            elif self.is_drop:
                c = self._counters.cpu().numpy()
                result["droprate"] = np.where(c[:, 0:1] > 0, c[:, 2:3] / np.maximum(c[:, 0:1], 1), 0.0)
            else:
                result["droprate"] = np.broadcast_to(np.zeros((W, 1)), result["route_time"].shape)
        if sendlog:
            self.last_sendlog = torch.cat(log_chunks, dim=1).cpu().numpy().view(np.uint32)
        return self._finish(result, legacy, R if d_lv is None else 0)

    def run_device(self, n_steps, lambd, lr={}, next_clock0=None):
        """Device-resident variant of `train` for Philox mode: one rlr_run_steps launch, inputs (the
        exp(-lambda/N) vector) and outputs (the [R, n_steps] per-step records) stay in HBM, nothing is
        copied and the stream is not synchronised.  The records are left in `self.metrics_device`.
        `next_clock0`: clock at which the NEXT call will start (default: where this one ends; 0 when every call is a
        fresh episode after `reset()`), so that its injection schedule is prefetched on the side stream."""
        import torch
        if self.rng != "philox":
            raise NotImplementedError("run_device needs rng='philox' (trace mode uploads host-generated injections)")
        self._check_sync()
        agent = self._agent
        h = getattr(agent, "_handle", None)
        if h is None:
            raise NotImplementedError("bind one of the device policies first")
        R, N, K = self.replicas, self.topology.N, int(n_steps)
        lam = np.broadcast_to(np.asarray(lambd, dtype=np.float64), (R,))
        key, d_en, d_thr, _ = self._lam_tensors(lam, upload=False)
        if self._resident.get("metrics_shape") != (R, K):
            self._resident["metrics"] = torch.empty((K, R, 2), dtype=torch.int64, device=self.device)
            self._resident["metrics_shape"] = (R, K)
        rp = self._run_params(agent, lr)
        rp.inject_mode = nat.INJECT_PHILOX
        rp.enlam = d_en.data_ptr()
        rp.inj_thr = d_thr.data_ptr()
        rp.metrics = self._resident["metrics"].data_ptr()
        nxt = self.clock + K if next_clock0 is None else int(next_clock0)
        self._schedule(h, rp, self.clock, K, key, lam.max(), (nxt, K))                # next call's schedule is prefetched
        nat.check(nat.lib().rlr_run_steps(h, K, C.byref(rp), self._stream_ptr()))
        self.launch_count += 1
        self._schedule_done(h, rp)
        self.clock += K

    @property
    def metrics_device(self):
        return self._resident.get("metrics")

    def check_status(self):
        """Reads the sticky per-replica status words (one device->host copy) and raises the reference's
        exception type: ValueError for NaN probabilities (hybrid.py:20-23), OverflowError for a full ring."""
        self._raise_on_status(self._status.cpu().numpy())

    @staticmethod
    def _finish(result, legacy, R):
        if R == 1:
            result = {k: v[0] for k, v in result.items()}
        if legacy:
            return result["route_time"], result.get("droprate", np.zeros_like(result["route_time"]))
        return result

    def _raise_on_status(self, status):
        bad = np.nonzero(status[:, 0])[0]
        if bad.size == 0:
            return
        r = int(bad[0])
        code, clk = int(status[r, 0]), int(status[r, 1])
        if code == nat.STATUS_NAN_PROB:
            raise ValueError(f"probabilities contain NaN (replica {r}, clock {clk}; {bad.size} replica(s) affected)")
        if code == nat.STATUS_INJECT_OVERFLOW:
            raise OverflowError(f"injection schedule overflow (replica {r}, clock {clk}): more packets at one node "
                                "than the schedule buffer holds")
        raise OverflowError(f"queue ring overflow (replica {r}, clock {clk}; {bad.size} replica(s) affected): the "
                            f"reference's queues are unbounded; re-run with queue_capacity > {self.queue_capacity}")

    def sendlog_rows(self, r=0):
        """Rows (clock_after_step, x, y, dest) of replica r from the last train(..., sendlog=True)."""
        log = self.last_sendlog[r]
        K, N = log.shape
        s, x = np.nonzero(log != 0xFFFFFFFF)
        v = log[s, x]
        slot = self.__dict__.get("_last_slot", 1)
        clock0 = self.clock - K * slot
        return np.stack([clock0 + (s + 1) * slot, x, v & 0xFFFF, v >> 16], axis=1).astype(np.int32)

    # ------------------------------------------------------------------ counters (env.py:271-276, 440-450)
    def _ctr(self, name):
        v = self._counters[:, nat.CTR[name]].cpu().numpy()
        if name == "all_packets" and self.__dict__.get("_pending"):
            v = v + np.array([len(pk) for pk in self._pending], dtype=np.int64)     # injected, joining at the next step
        return int(v[0]) if self.replicas == 1 else v

    all_packets = property(lambda self: self._ctr("all_packets"))
    end_packets = property(lambda self: self._ctr("end_packets"))
    drop_packets = property(lambda self: self._ctr("drop_packets"))
    hops = property(lambda self: self._ctr("hops"))
    route_time = property(lambda self: self._ctr("route_time"))
    sends = property(lambda self: self._ctr("sends"))

    @property
    def active_packets(self):
        return self.all_packets - self.end_packets - self.drop_packets

    @staticmethod
    def _ratio(a, b):
        if np.ndim(b) == 0:
            return a / b if b > 0 else 0
        return np.where(b > 0, a / np.maximum(b, 1), 0.0)

    ave_hops = property(lambda self: self._ratio(self.hops, self.end_packets))
    ave_route_time = property(lambda self: self._ratio(self.route_time, self.end_packets))
    drop_rate = property(lambda self: self._ratio(self.drop_packets, self.all_packets))

    def counters(self):
        """Dict of per-replica int64 arrays, one device->host copy."""
        c = self._counters.cpu().numpy()
        out = {k: c[:, i].copy() for k, i in nat.CTR.items()}
        out["active_packets"] = out["all_packets"] - out["end_packets"] - out["drop_packets"]
        out["clock"] = np.full(self.replicas, self.clock, dtype=np.int64)
        return out

    def stats_device(self):
        """Sum of the per-replica counters as a device int64[8] tensor (rlr_reduce_stats), ready for an
        NCCL all-reduce across ranks (rlrouting_b200.dist)."""
        h = getattr(self._agent, "_handle", None)
        if h is None:
            raise NotImplementedError("bind a device policy first")
        nat.check(nat.lib().rlr_reduce_stats(h, self._stats.data_ptr(), self._stream_ptr()))
        return self._stats

    def table_stats(self, prev=None, level_index=None, table="Qtable"):
        """Q-convergence statistics of the bound agent's table on the device (rlr_table_stats): per replica the sum of
        the entries and the sum of |entry - prev entry| against an earlier copy `prev` (a device tensor shaped like the
        table, e.g. `agent.table_device('Qtable').clone()`), and the same pooled per load level.  Returns device tensors
        {"per_replica": [R, 2], "levels": [L, 2]} (L = 1 without `level_index`), fp64 in a fixed summation order."""
        import torch
        h = self._handle_or_raise()
        agent = self._agent
        names = list(getattr(agent, "_table_names", ()))
        if table not in names:
            raise ValueError(f"{type(agent).__name__} has no table {table!r}")
        tdt = self._agent._dev[table].dtype
        if prev is not None and (prev.dtype != tdt or tuple(prev.shape) != (self.replicas, self.topology.tsize)
                                 or prev.device != self.device or not prev.is_contiguous()):
            raise ValueError("prev must be a contiguous device tensor of the table's dtype and shape [R, tsize]")
        d_lv, L = None, 1
        if level_index is not None:
            lv = np.ascontiguousarray(np.broadcast_to(np.asarray(level_index), (self.replicas,)), dtype=np.int32)
            L = int(lv.max()) + 1
            d_lv = torch.from_numpy(np.array(lv)).to(self.device)
        per = torch.empty((self.replicas, 2), dtype=torch.float64, device=self.device)
        lev = torch.empty((L, 2), dtype=torch.float64, device=self.device)
        nat.check(nat.lib().rlr_table_stats(h, names.index(table), prev.data_ptr() if prev is not None else None, per.data_ptr(),
                                            d_lv.data_ptr() if d_lv is not None else None, L, lev.data_ptr(), self._stream_ptr()))
        self.launch_count += 2
        return {"per_replica": per, "levels": lev}

    def _queue_len_device(self):
        """len(Node.queue) per (replica, node) as a device tensor."""
        n = self._q_tail - self._q_head
        return n - self._q_dead if self._general else n

    def _paged_record_index(self, qid, j, head):
        """Pool index (within its replica) of record j (0 = head) of every queue in `qid` (flat replica * N + node ids), all
        device tensors: follows each queue's page list for j's page distance from the head page."""
        import torch
        P, N = self.page_records, self.topology.N
        off = (head[qid] & (P - 1)) + j                                  # offset from the start of the head page
        hop = off // P
        page = (self._q_pages.view(-1, 4)[qid, 0].to(torch.int64)) & 0xFFFF
        nxt = self._page_next.to(torch.int64) & 0xFFFF
        rep_of = qid // N
        for k in range(int(hop.max().item()) if hop.numel() else 0):
            page = torch.where(hop > k, nxt[rep_of, page], page)
        return page * P + off % P

    def queue_array(self, x, r=0):
        """Queue of node x of replica r, head first: rows (dest, source, hops, birth, start_queue)."""
        head = int(self._q_head[r, x].item()) & 0xFFFFFFFF
        tail = int(self._q_tail[r, x].item()) & 0xFFFFFFFF
        n = (tail - head) & 0xFFFFFFFF
        if self._paged:
            import torch
            dev = self.device
            qid = torch.full((n,), r * self.topology.N + x, dtype=torch.int64, device=dev)
            hd = (self._q_head.reshape(-1).to(torch.int64)) & 0xFFFFFFFF
            idx = self._paged_record_index(qid, torch.arange(n, device=dev), hd)
            rec = self._pool[r][idx].cpu().numpy().view(np.uint32).reshape(n, 4)
        else:
            idx = (head + np.arange(n, dtype=np.int64)) & (self.queue_capacity - 1)
            rec = self._ring[r, x].cpu().numpy()[idx].view(np.uint32).reshape(n, 4)
        out = np.empty((n, 5), dtype=np.int32)
        out[:, 0] = rec[:, 0] & 0xFFFF
        out[:, 1] = rec[:, 0] >> 16
        out[:, 2:] = rec[:, 1:].view(np.int32)
        return out[rec[:, 0] != 0xFFFFFFFF]             # general timing: entries sent from the middle of the queue
