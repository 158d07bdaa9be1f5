#!/usr/bin/env python
"""bench.py — simulated packet-hops/sec of the batched routing simulator on N B200s.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank/GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json metric: "65,536 6x6 Q-routing replicas" on 8 GPUs, weak scaling):
  6x6.net, Qroute (fp64 tables), load 2.0, Philox injection, 8,192 replicas PER GPU.  One bench
  "step" is one pass of the hot path: `--sim-steps` simulator steps (inject -> send -> arrive ->
  learn) over every replica in ONE kernel launch.  `value` = packet-hops (sends) of all ranks in the
  timed region / max-over-ranks device time; `e2e` = the same through `Network.train` (host lambda
  in, per-step mean-delivery-time vectors [R, sim_steps] back on the host).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "simulated packet-hops/sec"
UNIT = "packet-hops/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--replicas-per-gpu", type=int, default=8192)
    ap.add_argument("--sim-steps", type=int, default=500, help="simulator steps per bench step (one launch)")
    ap.add_argument("--net", default="6x6.net")
    ap.add_argument("--policy", default="Qroute", choices=["Shortest", "Qroute", "HybridQ", "MaHybridQ", "CQ", "CDRQ", "GlobalRoute"])
    ap.add_argument("--load", type=float, default=2.0)
    ap.add_argument("--queue-capacity", type=int, default=0, help="ring slots per (replica, node); 0 = from the run length")
    ap.add_argument("--replicas-per-block", type=int, default=0)
    ap.add_argument("--threads-per-block", type=int, default=0)
    ap.add_argument("--tables-in-smem", type=int, default=-1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def plan_capacity(a):
    """The reference's queues are unbounded and grow ~linearly under congestion (Qroute 6x6 load 2.0
    reaches 1,040 per node in 10k steps).  Size the rings for the whole run (both measurement passes restart from
    clock 0) and, if that would not fit, shorten the simulator steps per bench step rather than overflow."""
    if a.queue_capacity:
        return
    budget = 40e9                                                   # bytes of ring storage per GPU
    n_nodes = {"6x6.net": 36, "6x6-modified.net": 36, "lata.net": 116}.get(a.net, 64)
    max_cap = 1 << int(math.log2(budget / (a.replicas_per_gpu * n_nodes * 16)))
    total = (a.steps + a.warmup) * a.sim_steps
    need = 512 + 0.2 * total * max(a.load / 2.0, 1.0)
    cap = 1 << max(int(math.ceil(math.log2(need))), 10)
    if cap > max_cap:
        cap = max_cap
        a.sim_steps = max(1, int((cap - 512) / (0.2 * max(a.load / 2.0, 1.0)) / (a.steps + a.warmup)))
    a.queue_capacity = cap


def workload_name(a):
    return (f"{a.net} {a.policy} fp64, load {a.load}, {a.replicas_per_gpu} replicas/GPU, Philox injection, "
            f"{a.sim_steps} simulator steps per bench step")


LR = {"q": 0.1, "p": 0.001, "e": 0.1}          # MaHybridQ needs p=0.001; q is every policy's default


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (test infrastructure used as the measured baseline)
# ------------------------------------------------------------------------------------------------
def oracle_envs(a, n, base=0, seed=1):
    import numpy as np
    from oracle import oracle as O
    from rlrouting_b200 import nets
    from rlrouting_b200.topology import Topology
    topo = Topology.from_text(nets.net_text(a.net))
    envs = []
    for r in range(n):
        e = O.OracleEnv(topo.links, a.policy, seed=seed, replica=base + r)
        if a.policy not in ("Shortest", "GlobalRoute"):
            g = np.random.Generator(np.random.Philox(key=[seed, base + r]))
            e.set_qtable(g.standard_normal(topo.tsize))
        envs.append(e)
    return envs


def cpu_sample(a, seconds):
    """Times the oracle on all host cores on a bounded sample of the workload (about `seconds` of CPU wall time)."""
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    n = 128 * cores
    envs = oracle_envs(a, n)
    O.run_many(envs, 200, a.load, LR["q"], LR["p"], LR["e"], nthreads=cores)   # warm the queues like the GPU's warm-up
    sends, steps, dt = 0, 0, 0.0
    while dt < seconds and steps < 10000:                                       # chunks of the same continuing run
        t0 = time.perf_counter()
        sends += O.run_many(envs, 500, a.load, LR["q"], LR["p"], LR["e"], nthreads=cores)
        dt += time.perf_counter() - t0
        steps += 500
    return {"value": sends / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} replicas x {steps} simulator steps (after 200 warm-up steps) of the same workload, "
                      f"oracle/rlr_oracle.c via OpenMP on {cores} threads, {dt:.1f} s"}


def run_reference(a):
    """--impl reference: the CPU implementation (oracle port; the Python reference cannot travel to
    the GPU box) on all host threads; each bench step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    n = 16 * cores
    envs = oracle_envs(a, n)
    sim = a.sim_steps
    for _ in range(a.warmup):
        O.run_many(envs, sim, a.load, LR["q"], LR["p"], LR["e"], nthreads=cores)
    t0 = time.perf_counter()
    sends = 0
    for _ in range(a.steps):
        sends += O.run_many(envs, sim, a.load, LR["q"], LR["p"], LR["e"], nthreads=cores)
    dt = time.perf_counter() - t0
    v = sends / dt
    sample = f"{n} replicas x {sim} simulator steps per bench step, oracle/rlr_oracle.c via OpenMP on {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "timing": "host wall clock around the CPU loop"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            p = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(a, topo, hops, injections, replica_steps):
    """per node-step head+tail 16 B; per injection a 16 B record; per hop a 16 B record read,
    the policy's rows, its writes and a 16 B record write (counted for every hop)."""
    w, dbar = 8, 4 if topo.max_deg <= 4 else topo.max_deg          # padded-row convention of SURVEY §8(d)
    per_hop = {"Shortest": 16 + 2 + 16, "GlobalRoute": 16 + 1 + 16,
               "Qroute": 16 + w * dbar + w * dbar + w + 16,
               "HybridQ": 16 + w * dbar + w * dbar + w + 16 + 2 * w * dbar,
               "MaHybridQ": 16 + w * dbar + w * dbar + w + 16 + w * dbar,
               "CQ": 16 + w * dbar + w * dbar + w + 16 + 4 * w,                       # + C_f read, C[x][d][k] read+write, Q read
               "CDRQ": 16 + 2 * w * dbar + w * dbar + 2 * w + 16 + 8 * w}[a.policy]  # + backward row, cell reads/writes
    dense = {"GlobalRoute": 5 * topo.N * topo.N,      # distance (int32) + next-hop (u8) tables rewritten every step
             "MaHybridQ": 4 * w * topo.tsize, "CQ": 2 * w * topo.tsize, "CDRQ": 2 * w * topo.tsize}.get(a.policy, 0)
    return hops * per_hop + replica_steps * (16 * topo.N + dense) + injections * 16, per_hop


def run_b200(a):
    import numpy as np
    import torch
    import rlrouting_b200 as rb
    from rlrouting_b200 import dist as rdist
    rank, world, local = rdist.init_from_env()
    if world != a.gpus and world > 1:
        a.gpus = world
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    R = a.replicas_per_gpu
    cls = getattr(rb, a.policy)

    def make():
        nw = rb.Network(a.net, replicas=R, rng="philox", seed=1, replica_base=rank * R, queue_capacity=a.queue_capacity,
                        replicas_per_block=a.replicas_per_block, threads_per_block=a.threads_per_block,
                        tables_in_smem=a.tables_in_smem)
        nw.agent = cls(nw)
        return nw

    sim = a.sim_steps

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`) ----------------
    nw = make()
    info = nw.launch_info()
    for _ in range(a.warmup):
        nw.run_device(sim, a.load, lr=LR)           # same call as the timed loop (allocates its buffers once)
    barrier()
    c0 = nw.stats_device().clone()
    launches0 = nw.launch_count
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    tw0 = time.perf_counter()
    ev0.record()
    marks = []
    for _ in range(a.steps):
        nw.run_device(sim, a.load, lr=LR)           # one rlr_run_steps launch, per-step metrics written on the device
        marks.append(torch.cuda.Event(enable_timing=True))
        marks[-1].record()
    ev1.record()
    barrier()
    tw1 = time.perf_counter()
    step_ms = [round(a_.elapsed_time(b_), 3) for a_, b_ in zip([ev0] + marks[:-1], marks)]
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    rdist.allreduce_max_(ms)
    clocks = sampler.stop(tw0, tw1) if sampler else None
    c1 = nw.stats_device().clone()
    delta = (c1 - c0)
    rdist.allreduce_sum_(delta)
    delta = delta.cpu().numpy()
    hops, inj = int(delta[5]), int(delta[0])
    launches = nw.launch_count - launches0
    ms_total = float(ms.item())
    value = hops / (ms_total * 1e-3)
    if not os.environ.get("RLR_BENCH_NOCHECK"):         # (destructive timing probes of tools/exp_run.sh skip the status check)
        nw.check_status()
    final = rdist.global_stats(nw)
    if os.environ.get("RLR_BENCH_NOCHECK"):
        print(json.dumps({"probe": True, "ms_per_step": ms_total / a.steps, "value": value, "step_ms": step_ms}))
        return

    # ---------------- end to end through Network.train (`e2e`) ----------------
    del nw
    torch.cuda.empty_cache()
    nw = make()
    out = None
    for _ in range(a.warmup):
        out = nw.train(sim, a.load, lr=LR)          # same usage as the timed loop (the previous result stays alive)
    barrier()
    s0 = nw.stats_device().clone()
    ev0.record()
    d2h = 0
    for _ in range(a.steps):
        out = nw.train(sim, a.load, lr=LR)          # host lambda -> device; [R, sim] route_time vectors -> host
        d2h += nw.last_d2h_bytes
        if os.environ.get("RLR_BENCH_DEBUG"):
            print("e2e step", {k: round(v, 2) for k, v in nw.last_timing.items()}, file=sys.stderr)
    ev1.record()
    barrier()
    ms2 = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    rdist.allreduce_max_(ms2)
    d2 = nw.stats_device().clone() - s0
    rdist.allreduce_sum_(d2)
    e2e_value = int(d2.cpu().numpy()[5]) / (float(ms2.item()) * 1e-3)
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nw.last_h2d_bytes,
           "d2h_bytes_per_step": d2h // a.steps,
           "note": "Network.train(sim_steps, load): lambda + threshold vectors H2D; the reference's per-step route_time "
                   "vector for every replica ([sim_steps, R] float64) D2H to pinned memory, overlapped chunk by chunk"}

    if rank != 0:
        return
    peak, peak_src = measured_peak()
    traffic = None
    try:                                            # DRAM bytes per launch from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(f"{a.net}|{a.policy}|{R}|{sim}|{'smem' if info['tables_in_smem'] else 'hbm'}")
        if t:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        pass
    abytes, per_hop = algorithmic_bytes(a, nw.topology, hops, inj, world * R * sim * a.steps)
    per_launch = abytes / max(a.steps * world, 1)                     # one step-kernel launch per GPU per bench step
    achieved = per_launch / (ms_total / a.steps * 1e-3) / 1e9        # GB/s per GPU (one launch per GPU per step)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "replicas_total": world * R, "sim_steps_per_bench_step": sim,
                   "queue_capacity": a.queue_capacity, "launch": info,
                   "l2_policy": "inputs larger than L2: Q tables %.0f MB + queue rings %.1f GB per GPU" % (
                       R * nw.topology.tsize * 8 / 1e6, R * nw.topology.N * a.queue_capacity * 16 / 1e9),
                   "hops_in_timed_region": hops, "mean_delivery_time": final["ave_route_time"],
                   "parallelism": f"replica-parallel x{world}, no data-path collective"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "algorithmic_bytes_per_launch": per_launch, "peak_source": peak_src, "kernel": "rlr_step_kernel",
                     "algorithmic_bytes_per_hop": per_hop,
                     "note": "algorithmic bytes per SURVEY 8(d); tables are staged in shared memory, so DRAM traffic is "
                             "far below the algorithmic figure (profiles/)"},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "step_ms": step_ms,
    }
    if not a.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_sample(a, a.cpu_seconds)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))


def _finalize():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


def main():
    a = parse()
    plan_capacity(a)
    if a.impl == "reference":
        run_reference(a)
        return
    run_b200(a)
    _finalize()


if __name__ == "__main__":
    main()
