"""Drop-in for the reference's `hybrid` module: put rlrouting_b200/compat on sys.path and the
reference's scripts (`from hybrid import ...`, train.py:8-13) resolve to the B200 implementation."""
from rlrouting_b200.hybrid import *  # noqa: F401,F403
from rlrouting_b200.hybrid import HybridQ, PolicyGradient, Qroute  # noqa: F401
