#!/usr/bin/env python
"""Timing probe: per replica-step device time of the fused Qroute kernel on W x H grids (how much of the 6x6 step
is the price of the second, nearly empty warp).  python tools/exp_grid.py 4x8 6x6 8x8"""
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rlrouting_b200 as rb          # noqa: E402


def grid_net(w, h):
    lines = [f"1000 {i} {i % w} {i // w} 0" for i in range(w * h)]
    for y in range(h):
        for x in range(w):
            i = y * w + x
            if x + 1 < w:
                lines.append(f"2000 {i + 1} {i} 0")
            if y + 1 < h:
                lines.append(f"2000 {i + w} {i} 0")
    return "\n".join(lines) + "\n"


for spec in sys.argv[1:]:
    w, h = map(int, spec.split("x"))
    with tempfile.NamedTemporaryFile("w", suffix=".net", delete=False) as f:
        f.write(grid_net(w, h))
    R, K = 8192, 500
    nw = rb.Network(f.name, replicas=R, rng="philox", seed=1, queue_capacity=2048)
    nw.agent = rb.Qroute(nw)
    load = 2.0 * w * h / 36.0                       # same per-node injection rate as 6x6 at load 2.0
    for _ in range(3):
        nw.run_device(K, load)
    torch.cuda.synchronize()
    s0 = nw.stats_device().clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        nw.run_device(K, load)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    hops = int((nw.stats_device() - s0)[5].item()) / 3
    info = nw.launch_info()
    print(f"{spec}: N={w * h} {info} ms/launch {ms:.3f} hops/launch {hops:.3g} hops/s {hops / ms * 1e3:.4g}", flush=True)
    del nw
    torch.cuda.empty_cache()
