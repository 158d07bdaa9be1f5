"""Shortest-path routing tables (reference shortest.py:6-58), built by the APSP CUDA kernel."""
import numpy as np

from . import _native as nat
from .base_policy import Policy


class Shortest(Policy):
    """`distance`, `choice` and `mask` have the reference's shapes (shortest.py:14-16); the sweep
    of `_calc_distance` (shortest.py:43-58) runs on the device, one thread per destination."""
    attrs = Policy.attrs | set(['distance', 'choice', 'mask'])
    _kernel_policy = nat.POLICY_SHORTEST

    def __init__(self, network, multiway=False, random=False):
        super().__init__(network)
        self.multiway = multiway
        self.random = random
        import torch
        topo, dev = network.topology, network.device
        self._dev = {"next_hop": torch.zeros(topo.N * topo.N, dtype=torch.uint8, device=dev),
                     "distance": torch.zeros(topo.N * topo.N, dtype=torch.float64, device=dev),
                     "choice": torch.zeros(topo.tsize, dtype=torch.uint8, device=dev),
                     "choice_mask": torch.zeros(topo.N * topo.N, dtype=torch.int16, device=dev)}
        self._handle = network._make_handle(self)
        nat.check(nat.lib().rlr_build_shortest_tables(self._handle, int(bool(multiway)), network._stream_ptr()))
        n = topo.N
        self.distance = self._dev["distance"].cpu().numpy().reshape(n, n)
        flat = self._dev["choice"].cpu().numpy().astype(bool)
        self.choice = {x: flat[topo.toff[x]:topo.toff[x + 1]].reshape(n, int(topo.deg[x])).copy() for x in range(n)}
        self.mask = np.ones((n, n), dtype=bool)                                    # shortest.py:25,28-34
        for x, neighbors in self.links.items():
            self.mask[x, x] = False
            self.mask[x, neighbors] = False

    def choose(self, source, dest):
        choices = self.links[source][self.choice[source][dest]]                    # shortest.py:38-41
        return np.random.choice(choices) if self.random else choices[0]

    def _after_load(self):
        """`load` (base_policy.py:62-66) replaced `distance` / `choice` / `mask` on the host: the device routes by its own
        copies (next hop = first True of a choice row, choice masks as bits), so upload them again."""
        import torch
        topo = self._network.topology
        n = topo.N
        flat = np.concatenate([np.asarray(self.choice[x], dtype=bool).reshape(n * int(topo.deg[x])) for x in range(n)])
        nh = np.zeros((n, n), dtype=np.uint8)
        cm = np.zeros((n, n), dtype=np.uint16)
        for x in range(n):
            c = np.asarray(self.choice[x], dtype=bool).reshape(n, int(topo.deg[x]))
            any_true = c.any(axis=1)
            nh[x] = np.where(any_true, c.argmax(axis=1), max(int(topo.deg[x]) - 1, 0))   # the APSP kernel's "no choice" value
            cm[x] = (c.astype(np.uint16) << np.arange(c.shape[1], dtype=np.uint16)[None, :]).sum(axis=1)
        self._dev["choice"].copy_(torch.from_numpy(flat.astype(np.uint8)))
        self._dev["next_hop"].copy_(torch.from_numpy(nh.reshape(-1)))
        self._dev["choice_mask"].copy_(torch.from_numpy(cm.reshape(-1).view(np.int16)))
        self._dev["distance"].copy_(torch.from_numpy(np.ascontiguousarray(self.distance, dtype=np.float64).reshape(-1)))


class GlobalRoute(Shortest):
    """Queue-aware global routing (reference shortest.py:84-104): after every step all distances are recomputed with
    node weight 1 + queue_size[y] and packets follow the new next-hop tables.  Each replica has its own tables on the
    device; `distance`, `choice` and `queue_size` read them back (one replica: the reference's shapes)."""
    _kernel_policy = nat.POLICY_GLOBALROUTE

    def __init__(self, network, multiway=False, random=False):
        Policy.__init__(self, network)
        self.multiway, self.random = multiway, random
        import torch
        topo, dev, R = network.topology, network.device, network.replicas
        n = topo.N
        self._dev = {"gr_choice_mask": torch.zeros((R, n * n), dtype=torch.int16, device=dev),     # bit j = choice[x][d][j]
                     "gr_distance": torch.zeros((R, n * n), dtype=torch.int32, device=dev),
                     "gr_qs_off": torch.zeros((R, n), dtype=torch.int32, device=dev)}
        self._handle = network._make_handle(self)
        nat.check(nat.lib().rlr_build_shortest_tables(self._handle, int(bool(multiway)), network._stream_ptr()))   # shortest.py:85-91
        self.mask = ~np.eye(n, dtype=bool)                                                    # shortest.py:87-88

    INF = 0x3FFFFFFF

    def _one(self, a):
        return a[0] if self._network.replicas == 1 else a

    @property
    def distance(self):
        n = self._network.topology.N
        d = self._dev["gr_distance"].cpu().numpy().reshape(-1, n, n).astype(np.float64)
        d[d >= self.INF] = np.inf
        return self._one(d)

    @distance.setter
    def distance(self, value):                              # Policy.load: back onto the device (every replica or one for all)
        import torch
        n, R = self._network.topology.N, self._network.replicas
        d = np.asarray(value, dtype=np.float64).reshape(-1, n * n)
        d = np.where(np.isfinite(d), d, float(self.INF)).astype(np.int32)
        self._dev["gr_distance"].copy_(torch.from_numpy(np.array(np.broadcast_to(d, (R, n * n)))))

    @property
    def choice(self):
        topo = self._network.topology
        n, R = topo.N, self._network.replicas
        cm = self._dev["gr_choice_mask"].cpu().numpy().view(np.uint16).reshape(R, n, n)
        out = {}
        for x in range(n):
            bits = (cm[:, x, :, None] >> np.arange(int(topo.deg[x]), dtype=np.uint16)[None, None, :]) & 1
            out[x] = self._one(bits.astype(bool))
        return out

    @choice.setter
    def choice(self, value):
        import torch
        topo = self._network.topology
        n, R = topo.N, self._network.replicas
        cm = np.zeros((R, n, n), dtype=np.uint16)
        for x in range(n):
            c = np.asarray(value[x], dtype=bool).reshape(-1, n, int(topo.deg[x]))
            cm[:, x, :] = (c.astype(np.uint16) << np.arange(c.shape[2], dtype=np.uint16)[None, None, :]).sum(axis=2)
        self._dev["gr_choice_mask"].copy_(torch.from_numpy(cm.reshape(R, n * n).view(np.int16)))

    def _after_load(self):
        pass                                                # distance / choice went through the setters above

    @property
    def queue_size(self):
        net = self._network
        q = (net._queue_len_device() + self._dev["gr_qs_off"]).cpu().numpy().astype(np.int64)
        return self._one(q)

    def choose(self, source, dest):
        return self.links[source][self.choice[source][dest]][0]

    def _before_network_reset(self):
        """Network.reset empties the queues but GlobalRoute.queue_size is only ever changed by the receive / send
        callbacks (shortest.py:93-97), so the reference keeps the stale counts; mirror that exactly."""
        net = self._network
        self._dev["gr_qs_off"] += net._queue_len_device()

    def learn(self, rewards, lr={}):
        self._network._learn(self, rewards, lr)
