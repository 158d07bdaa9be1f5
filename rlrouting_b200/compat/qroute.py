"""Drop-in for the reference's `qroute` module: put rlrouting_b200/compat on sys.path and the
reference's scripts (`from qroute import ...`, train.py:8-13) resolve to the B200 implementation."""
from rlrouting_b200.qroute import *  # noqa: F401,F403
from rlrouting_b200.qroute import CDRQ, CQ, Qroute  # noqa: F401
