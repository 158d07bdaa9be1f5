#!/usr/bin/env python
"""Host-side breakdown of one end-to-end bench step (reset + reset_tables + train with per-level curves):
where the milliseconds between the device-resident `value` and `e2e` go.  python tools/e2e_breakdown.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rlrouting_b200 as rb      # noqa: E402

R, T = 8192, 5000
nw = rb.Network("6x6.net", replicas=R, rng="philox", seed=1, queue_capacity=2048)
ag = rb.Qroute(nw)
nw.agent = ag
ag.keep_initial_tables()
lv = np.zeros(R, dtype=np.int32)
lam = np.full(R, 2.0)
for _ in range(3):
    ag.reset_tables(); nw.reset(); nw.train(T, lam, level_index=lv, next_clock0=0)
torch.cuda.synchronize()
acc = {}
t_all = time.perf_counter()
for _ in range(8):
    t0 = time.perf_counter()
    ag.reset_tables(); nw.reset()
    t1 = time.perf_counter()
    nw.train(T, lam, level_index=lv, next_clock0=0)
    t2 = time.perf_counter()
    for k, v in dict(nw.last_timing, fresh_ms=1e3 * (t1 - t0), train_ms=1e3 * (t2 - t1)).items():
        acc[k] = acc.get(k, 0.0) + v / 8
torch.cuda.synchronize()
print({k: round(v, 3) for k, v in acc.items()}, "wall per step ms", round(1e3 * (time.perf_counter() - t_all) / 8, 3))
