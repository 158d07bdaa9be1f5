"""Multi-agent hybrid Q-routing with eligibility traces (reference multi_agent.py:6-50)."""
from . import _native as nat
from .hybrid import HybridQ


class MaHybridQ(HybridQ):
    attrs = HybridQ.attrs | set(['discount_trace', 'Trace'])
    _kernel_policy = nat.POLICY_MAHYBRIDQ
    _table_names = ('Qtable', 'Theta', 'Trace')
    _lr_default = {'q': 0.1, 'p': 0.1}                                             # multi_agent.py:17

    def __init__(self, network, initQ=0, initP=0, discount=0.99, discount_trace=0.6):
        self.discount_trace = discount_trace
        self.reward_shape = 0
        super().__init__(network, initQ=initQ, initP=initP, discount=discount)

    Trace = property(lambda self: self._get_table("Trace"), lambda self, v: self._set_table("Trace", v))

    def clean(self):
        self._dev["Trace"].zero_()                                                 # multi_agent.py:48-50

    def __repr__(self):
        return "<MultiAgent discount:{} discount_trace:{}>".format(self.discount, self.discount_trace)
