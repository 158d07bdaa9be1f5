"""Drop-in for the reference's `multi_agent` module: put rlrouting_b200/compat on sys.path and the
reference's scripts (`from multi_agent import ...`, train.py:8-13) resolve to the B200 implementation."""
from rlrouting_b200.multi_agent import *  # noqa: F401,F403
from rlrouting_b200.multi_agent import MaHybridQ, HybridQ  # noqa: F401
