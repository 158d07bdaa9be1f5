"""Drop-in for the reference's `base_policy` module: put rlrouting_b200/compat on sys.path and the
reference's scripts (`from base_policy import ...`, train.py:8-13) resolve to the B200 implementation."""
from rlrouting_b200.base_policy import *  # noqa: F401,F403
from rlrouting_b200.base_policy import Policy  # noqa: F401
