"""Pure-Python twin of the Philox stream law in rlr_oracle.c — TEST INFRASTRUCTURE ONLY.

Used by ref_harness.py to drive the UNMODIFIED reference with the same counter-based random
numbers the oracle and the GPU use (SURVEY §8c "Philox mode"): `Network.new_packet` and the
uniform inside `np.random.choice` are replaced from outside, no reference file is edited.
"""
import numpy as np

M32 = 0xFFFFFFFF


def philox4x32(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


class Stream:
    """Word stream for (seed, replica, clock, node, purpose); block b gives words 4b..4b+3."""

    def __init__(self, seed, replica, clock, node, purpose):
        self.c = [clock & M32, (node | (purpose << 16)) & M32, replica & M32, (replica >> 32) & M32]
        self.k = (seed & M32, (seed >> 32) & M32)
        self.buf, self.block = [], 0

    def next(self):
        if not self.buf:
            c = list(self.c)
            c[1] |= self.block << 24
            self.buf = list(philox4x32(c, self.k))
            self.block += 1
        return self.buf.pop(0)


def new_packets(seed, replica, clock, n_nodes, enlam):
    """Poisson-splitting injection law: [(source, dest), ...] in node order."""
    out = []
    for x in range(n_nodes):
        s = Stream(seed, replica, clock, x, 0)
        n, prod = 0, np.float64(1.0)
        while True:
            u = (np.float64(s.next()) + np.float64(0.5)) * np.float64(1.0 / 4294967296.0)
            prod = prod * u
            if prod > enlam:
                n += 1
            else:
                break
        for _ in range(n):
            d = (s.next() * (n_nodes - 1)) >> 32
            if d >= x:
                d += 1
            out.append((x, d))
    return out


def uniform53(seed, replica, clock, node):
    s = Stream(seed, replica, clock, node, 1)
    a, b = s.next() >> 5, s.next() >> 6
    return (a * 67108864.0 + b) / 9007199254740992.0
