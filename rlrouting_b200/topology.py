"""Topology parsing and the device-side graph layout.

`parse_net` restates `Network.read_network` (reference env.py:282-296): a `1000 id ...` line
declares a node and assigns the next integer ID, a `2000 a b ...` line appends b to links[a]
and a to links[b].  Neighbour order is file order and defines the action index
(base_policy.py:19-23).

`Topology` holds what the CUDA kernels need:
  row_ptr[N+1], col[sum_deg]   CSR adjacency in that order
  toff[x] = N * row_ptr[x]     so table[x][d][k] (the reference's per-node (N, deg_x) array,
                               qroute.py:13-15) lives at toff[x] + d*deg[x] + k  (no padding)
  rev[e]                       for CSR slot e = (y, j) with z = col[e]: the action index of y
                               in links[z]  (so a receiver can test "did z send to me")
  heap_pos[k][rank]            position at which the rank-th pushed of k equal-key events is
                               popped by heapq (reference env.py:48-49,144-145,350-362: Event
                               compares arrive_time only, so ties pop in a fixed non-FIFO
                               permutation that depends on k alone)
"""
from collections import OrderedDict
import heapq

import numpy as np


def parse_net(text):
    links, proj = OrderedDict(), {}
    next_id = 0
    for raw in text.splitlines():
        l = raw.split()
        if not l:
            continue
        if l[0] == "1000":
            proj[l[1]] = next_id
            links[next_id] = []
            next_id += 1
        elif l[0] == "2000":
            src, dst = proj[l[1]], proj[l[2]]
            links[src].append(dst)
            links[dst].append(src)
    return links, proj


class _Tie:
    """An item that, like the reference's Event, never compares less than an equal-time item."""
    __slots__ = ("i",)

    def __init__(self, i):
        self.i = i

    def __lt__(self, other):
        return False


def heap_pop_order(k):
    """Order in which k same-arrive-time events pushed as 0..k-1 leave `heapq`."""
    h = []
    for i in range(k):
        heapq.heappush(h, _Tie(i))
    return [heapq.heappop(h).i for _ in range(k)]


class Topology:
    MAX_NODES = 255     # node ids and heap positions are stored in 8/16-bit fields on the device
    MAX_DEGREE = 16

    def __init__(self, links):
        self.links = OrderedDict((int(k), [int(v) for v in vs]) for k, vs in links.items())
        N = self.N = len(self.links)
        if list(self.links.keys()) != list(range(N)):
            raise ValueError("node ids must be 0..N-1 in declaration order")
        if N < 2 or N > self.MAX_NODES:
            raise NotImplementedError(f"node count {N} outside [2, {self.MAX_NODES}]")
        self.deg = np.array([len(self.links[x]) for x in range(N)], dtype=np.int32)
        if self.deg.min() < 1:
            raise NotImplementedError("isolated node: a packet queued there could never be routed")
        if self.deg.max() > self.MAX_DEGREE:
            raise NotImplementedError(f"degree {self.deg.max()} > {self.MAX_DEGREE}")
        self.row_ptr = np.zeros(N + 1, dtype=np.int32)
        self.row_ptr[1:] = np.cumsum(self.deg)
        self.sum_deg = int(self.row_ptr[N])
        self.max_deg = int(self.deg.max())
        self.col = np.array([y for x in range(N) for y in self.links[x]], dtype=np.int32)
        self.toff = (self.row_ptr.astype(np.int64) * N)
        self.tsize = int(N * self.sum_deg)
        self.action_idx = {x: {y: i for i, y in enumerate(self.links[x])} for x in range(N)}
        for x in range(N):
            if len(self.action_idx[x]) != len(self.links[x]) or x in self.action_idx[x]:
                raise NotImplementedError("duplicate edges and self-loops are not supported")
        self.rev = np.array([self.action_idx[y][x] for x in range(N) for y in self.links[x]],
                            dtype=np.int32)
        pos = np.zeros((N + 1, N), dtype=np.uint8 if N <= 256 else np.uint16)
        for k in range(1, N + 1):
            for j, rank in enumerate(heap_pop_order(k)):
                pos[k, rank] = j
        self.heap_pos = pos

    @classmethod
    def from_text(cls, text):
        links, proj = parse_net(text)
        t = cls(links)
        t.proj = proj
        return t

    # ---- conversions between the reference's dict-of-(N,deg_x) arrays and the packed layout ----
    def pack(self, tables, dtype=np.float64):
        """Dict[x -> (N, deg_x) array]  ->  flat [tsize] array."""
        out = np.empty(self.tsize, dtype=dtype)
        for x in range(self.N):
            out[self.toff[x]:self.toff[x + 1]] = np.asarray(tables[x], dtype=dtype).reshape(-1)
        return out

    def unpack(self, flat):
        """Flat [tsize] array -> Dict[x -> (N, deg_x) array] (copies)."""
        flat = np.asarray(flat)
        return {x: flat[self.toff[x]:self.toff[x + 1]].reshape(self.N, int(self.deg[x])).copy()
                for x in range(self.N)}
