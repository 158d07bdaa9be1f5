"""rlrouting_b200 — B200-native batched packet-routing simulator behind the API of xuxife/rlrouting.

    from rlrouting_b200 import Network, Qroute
    nw = Network("6x6.net", replicas=8192, rng="philox", seed=1)
    nw.agent = Qroute(nw)
    out = nw.train(1000, lambd=2.0)

The compute path is the CUDA library csrc/librlrouting_b200.so (C ABI: include/rlrouting_b200.h);
there is no CPU fallback.  Drop-in module names of the reference (`env`, `qroute`, `shortest`,
`hybrid`, `multi_agent`, `base_policy`, `config`) live in rlrouting_b200/compat/.
"""
from .base_policy import Policy
from .env import Network, Node, Packet, Reward
from .hybrid import HybridQ, PolicyGradient
from .multi_agent import MaHybridQ
from .qroute import CDRQ, CQ, Qroute
from .shortest import GlobalRoute, Shortest
from .topology import Topology

__all__ = ["Network", "Node", "Packet", "Reward", "Policy", "Qroute", "CQ", "CDRQ", "Shortest", "GlobalRoute", "PolicyGradient", "HybridQ",
           "MaHybridQ", "Topology"]
