"""`Policy`: the plugin interface every routing agent implements (reference base_policy.py:5-66).

Same class attributes (`mode`, `attrs`), instance attributes (`links`, `action_idx`) and methods as
the reference.  The built-in agents (Shortest, Qroute, HybridQ, MaHybridQ) additionally carry a
kernel policy id and own their tables as device tensors; a `Policy` subclass without a kernel id
cannot be bound to a `Network`, because there is no CPU fallback.
"""
import pickle

import numpy as np


class Policy:
    mode = None
    attrs = set(['links'])
    _kernel_policy = None          # RLR_POLICY_* for agents implemented by the CUDA kernels
    _table_names = ()              # device tables exposed in the reference's dict-of-arrays shape

    def __init__(self, network):
        self.links = {k: np.array(v, dtype=int) for k, v in network.links.items()}        # base_policy.py:19-20
        self.action_idx = {node: {a: i for i, a in enumerate(neighbors)}                   # base_policy.py:21-23
                           for node, neighbors in self.links.items()}
        self._network = network
        self._handle = None
        self._dev = {}

    # ---- the reference's callback surface (base_policy.py:25-55) ----
    def choose(self, source, dest):
        pass

    def get_info(self, source, dest, action):
        return {}

    def learn(self, rewards, lr={}):
        pass

    def clean(self):
        pass

    def receive(self, source, dest):
        pass

    def send(self, source, dest):
        pass

    def reset(self):
        pass

    def drop_penalty(self, penalty):
        pass

    # ---- checkpoint interchange (base_policy.py:57-66): same pickle layout as the reference ----
    def _attr_value(self, k):
        if k in self._table_names or k not in self.__dict__:        # device tables and property-backed attributes
            return getattr(self, k)
        return self.__dict__[k]

    def store(self, filename):
        with open(filename, 'wb') as f:
            pickle.dump({k: self._attr_value(k) for k in self.attrs}, f)

    def load(self, filename):
        with open(filename, 'rb') as f:
            for k, v in pickle.load(f).items():
                if k in self._table_names or isinstance(getattr(type(self), k, None), property):
                    setattr(self, k, v)                 # re-uploads to the device (a read-only property raises)
                else:
                    self.__dict__[k] = v
        self._after_load()

    def _after_load(self):
        """Hook: agents whose device tables are derived from loaded host attributes re-upload them here."""

    # ---- device tables <-> the reference's Dict[x -> (N, deg_x) ndarray] ----
    def table_array(self, name):
        """The whole table as a NumPy array [R, N*sum_deg] (packed layout of Topology.pack)."""
        return self._dev[name].cpu().numpy()

    def table_device(self, name):
        """The device tensor [R, N*sum_deg] behind a table (no copy); `.clone()` it to keep a copy for Network.table_stats."""
        return self._dev[name]

    def _get_table(self, name):
        topo = self._network.topology
        flat = self.table_array(name)
        if flat.shape[0] == 1:
            return topo.unpack(flat[0])
        return {x: flat[:, topo.toff[x]:topo.toff[x + 1]].reshape(flat.shape[0], topo.N, int(topo.deg[x])).copy()
                for x in range(topo.N)}

    def _set_table(self, name, tables):
        import torch
        topo = self._network.topology
        dev = self._dev[name]
        R = dev.shape[0]
        first = np.asarray(tables[0])
        if first.ndim == 2:
            flat = np.repeat(topo.pack(tables)[None, :], R, axis=0)
        else:
            flat = np.concatenate([np.asarray(tables[x], dtype=np.float64).reshape(R, -1) for x in range(topo.N)], axis=1)
        dev.copy_(torch.from_numpy(np.ascontiguousarray(flat, dtype=np.float64)))      # (rounds to the table's storage type)
