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
        if multiway or random:
            raise NotImplementedError("Shortest(multiway=True / random=True) is not implemented on the device "
                                      "(SURVEY §8 row f3); there is no CPU fallback")
        super().__init__(network)
        self.multiway = multiway
        self.random = random
        import torch
        topo, dev = network.topology, network.device
        self._dev = {"next_hop": torch.zeros(topo.N * topo.N, dtype=torch.uint8, device=dev),
                     "distance": torch.zeros(topo.N * topo.N, dtype=torch.float64, device=dev),
                     "choice": torch.zeros(topo.tsize, dtype=torch.uint8, device=dev)}
        self._handle = network._make_handle(self)
        nat.check(nat.lib().rlr_build_shortest_tables(self._handle, network._stream_ptr()))
        n = topo.N
        self.distance = self._dev["distance"].cpu().numpy().reshape(n, n)
        flat = self._dev["choice"].cpu().numpy().astype(bool)
        self.choice = {x: flat[topo.toff[x]:topo.toff[x + 1]].reshape(n, int(topo.deg[x])).copy() for x in range(n)}
        self.mask = np.ones((n, n), dtype=bool)                                    # shortest.py:25,28-34
        for x, neighbors in self.links.items():
            self.mask[x, x] = False
            self.mask[x, neighbors] = False

    def choose(self, source, dest):
        return self.links[source][self.choice[source][dest]][0]                    # shortest.py:38-41
