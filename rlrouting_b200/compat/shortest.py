"""Drop-in for the reference's `shortest` module: put rlrouting_b200/compat on sys.path and the
reference's scripts (`from shortest import ...`, train.py:8-13) resolve to the B200 implementation."""
from rlrouting_b200.shortest import *  # noqa: F401,F403
from rlrouting_b200.shortest import GlobalRoute, Shortest  # noqa: F401
