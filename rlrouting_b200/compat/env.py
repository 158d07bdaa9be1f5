"""Drop-in for the reference's `env` module: put rlrouting_b200/compat on sys.path and the
reference's scripts (`from env import ...`, train.py:8-13) resolve to the B200 implementation."""
from rlrouting_b200.env import *  # noqa: F401,F403
from rlrouting_b200.env import Network, Node, Packet, Reward  # noqa: F401
