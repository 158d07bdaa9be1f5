"""Policy-gradient mixin and HybridQ (reference hybrid.py:6-62)."""
from . import _native as nat
from .base_policy import Policy
from .qroute import Qroute


class PolicyGradient(Policy):
    """Theta tables and the softmax-sampled choice (hybrid.py:6-39).  On its own it has no learning
    rule in the reference either; it is the base of HybridQ / MaHybridQ."""
    attrs = Policy.attrs | set(['Theta', 'discount'])

    Theta = property(lambda self: self._get_table("Theta"), lambda self, v: self._set_table("Theta", v))


class HybridQ(PolicyGradient, Qroute):
    attrs = Qroute.attrs | PolicyGradient.attrs
    _kernel_policy = nat.POLICY_HYBRIDQ
    _table_names = ('Qtable', 'Theta')
    _lr_default = {'q': 0.1, 'p': 0.1, 'e': 0.1}                                   # hybrid.py:55
    _q_draws = 2       # the reference's MRO runs Qroute.__init__ twice (hybrid.py:10,44-47); 2nd draw kept

    def __init__(self, network, initQ=0, initP=0, add_entropy=True, discount=0.99):
        Policy.__init__(self, network)
        self.add_entropy = add_entropy
        self.discount = discount
        self.threshold = 0.1
        self._alloc_tables(network, initQ, initP)
