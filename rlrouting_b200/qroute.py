"""Q-routing (reference qroute.py:6-53).  The table lives on the device as [R, N*sum_deg] float64."""
import numpy as np

from . import _native as nat
from . import trace
from .base_policy import Policy


def _draw_initial_q(network, initQ, times):
    """One N(initQ,1) table per replica, consuming each replica's RNG exactly as the reference's
    constructor does (qroute.py:13-15; twice for HybridQ/MaHybridQ)."""
    topo = network.topology
    out = np.empty((network.replicas, topo.tsize), dtype=np.float64)
    if network.rng == "trace":
        for r in range(network.replicas):
            out[r] = topo.pack(trace.draw_qtable_normals(network._rs[r], topo, initQ, times))
    else:
        for r in range(network.replicas):
            g = np.random.Generator(np.random.Philox(key=[network.seed, network.replica_base + r]))
            out[r] = initQ + g.standard_normal(topo.tsize)
    return out


class Qroute(Policy):
    attrs = Policy.attrs | set(['Qtable', 'discount', 'threshold'])
    _kernel_policy = nat.POLICY_QROUTE
    _table_names = ('Qtable',)
    _lr_default = {'q': 0.1}                                                       # qroute.py:46
    _q_draws = 1

    def __init__(self, network, initQ=0, discount=0.99, threshold=0.1):
        Policy.__init__(self, network)
        self.discount = discount
        self.threshold = threshold
        self._alloc_tables(network, initQ, 0.0)

    def _alloc_tables(self, network, initQ, initP):
        import torch
        q0 = _draw_initial_q(network, initQ, self._q_draws)
        if network.dtype == "float32":
            if self._kernel_policy != nat.POLICY_QROUTE:
                raise NotImplementedError(f"dtype='float32' (fp32 table storage) is implemented for Qroute, not for {type(self).__name__}")
            q0 = q0.astype(np.float32)          # the fp64 draws, rounded once
        self._dev = {"Qtable": torch.from_numpy(q0).to(network.device)}
        for name in self._table_names[1:]:
            self._dev[name] = torch.empty_like(self._dev["Qtable"])
        self._handle = network._make_handle(self)
        self._initP = float(initP)
        # qroute.py:16-20 structure overrides; Theta = initP (hybrid.py:13-14); Trace = 0 (multi_agent.py:14-15)
        nat.check(nat.lib().rlr_init_tables(self._handle, float(initP), network._stream_ptr()))

    def reset_tables(self):
        """Back to the tables the constructor produced (the same N(initQ,1) draws, structure overrides, Theta = initP,
        Trace = 0) without leaving the device: what re-running `Policy(network)` with the same seed would give.  Together
        with `Network.reset()` this starts a fresh episode; the first call keeps a device copy of the initial draws."""
        q = self._dev["Qtable"]
        if getattr(self, "_q_initial", None) is None:
            raise RuntimeError("reset_tables() needs the initial draws: call keep_initial_tables() before the first step")
        q.copy_(self._q_initial)
        nat.check(nat.lib().rlr_init_tables(self._handle, self._initP, self._network._stream_ptr()))

    def keep_initial_tables(self):
        """Keeps a device copy of the current Qtable as the state `reset_tables()` returns to."""
        self._q_initial = self._dev["Qtable"].clone()

    def set_initial_q(self, normals, initP=0.0):
        """Replace the constructor's draws: `normals` is [tsize] or [R, tsize] (packed layout); the
        structure overrides of qroute.py:16-20 are re-applied, Theta is reset to initP and Trace to 0."""
        import torch
        net = self._network
        a = np.array(np.broadcast_to(np.asarray(normals, dtype=np.float64), (net.replicas, net.topology.tsize)))
        self._dev["Qtable"].copy_(torch.from_numpy(a))                   # (rounds to the table's storage type)
        nat.check(nat.lib().rlr_init_tables(self._handle, float(initP), net._stream_ptr()))

    Qtable = property(lambda self: self._get_table("Qtable"), lambda self, v: self._set_table("Qtable", v))

    def learn(self, rewards, lr={}):
        """qroute.py:51-53 / hybrid.py:55-62 / multi_agent.py:17-43 on the device, for the rewards of the latest
        Network.step() (Network.train fuses this call into the step kernel)."""
        self._network._learn(self, rewards, lr)

    def choose(self, source, dest, idx=False):
        scores = self.Qtable[source][..., dest, :]
        if idx:
            return np.argmax(scores, axis=-1), scores.max(axis=-1)
        return self.links[source][np.argmax(scores, axis=-1)]


class CQ(Qroute):
    """Confidence-based Q-routing (reference qroute.py:56-101).  `confidence` has the shape of `Qtable`; the
    learning rate of an update is max(C_f, 1 - C_old) and every confidence entry decays by `decay` each step."""
    attrs = Qroute.attrs | set(['decay', 'confidence'])
    _kernel_policy = nat.POLICY_CQ
    _table_names = ('Qtable', 'confidence')
    _lr_default = {}                                                               # qroute.py:95
    _dev_alias = {"Theta": "confidence"}        # the C ABI's `theta` slot carries the confidence table

    def __init__(self, network, decay=0.9, initQ=0, discount=0.9):
        Policy.__init__(self, network)
        self.discount = discount
        self.threshold = 0.1
        self.decay = decay
        self._alloc_tables(network, initQ, 0.0)      # rlr_init_tables also applies CQ.clean's structure (qroute.py:66-71)

    confidence = property(lambda self: self._get_table("confidence"), lambda self, v: self._set_table("confidence", v))

    def clean(self):
        pass        # Network.reset re-initialises the confidence table on the device (rlr_reset), qroute.py:66-71



class CDRQ(CQ):
    """Confidence-based dual reinforcement Q-routing (reference qroute.py:104-122): runs the network in 'dual' mode
    (env.py:92-94,154-160: rewards are queue lengths) and adds a backward update of Q[y][packet.source]."""
    mode = 'dual'
    _kernel_policy = nat.POLICY_CDRQ
