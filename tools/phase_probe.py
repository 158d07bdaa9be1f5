#!/usr/bin/env python
"""Per-phase cycle accounting of the step kernel (library built with -DRLR_EXP_TIMING): average cycles thread 0
of a CTA spends in each phase of one simulator step.   RLR_LIB=build/exp/t1.so python tools/phase_probe.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rlrouting_b200 as rb                      # noqa: E402
from rlrouting_b200 import _native as nat        # noqa: E402

R, K = 8192, 500
nw = rb.Network("6x6.net", replicas=R, rng="philox", seed=1, queue_capacity=4096)
nw.agent = rb.Qroute(nw)
for _ in range(3):
    nw.run_device(K, 2.0)
torch.cuda.synchronize()
out = (C.c_ulonglong * 8)()
nat.lib().rlr_debug_phase_cycles(out, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
nw.run_device(K, 2.0)
e1.record()
torch.cuda.synchronize()
nat.lib().rlr_debug_phase_cycles(out, 0)
v = np.array(list(out), dtype=np.float64) / (R * K)
names = ["A+B", "barrier1", "C", "D+inject", "barrier2", "C-keys", "A+pop", "general"]
print(os.environ.get("RLR_LIB"), {n: round(x) for n, x in zip(names, v)}, "sum", round(v.sum()), "ms", round(e0.elapsed_time(e1), 3))
