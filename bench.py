#!/usr/bin/env python
"""bench.py — simulated packet-hops/sec of the batched routing simulator on N B200s.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank/GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

Headline workload (BASELINE.json metric: "65,536 6x6 Q-routing replicas" on 8 GPUs, weak scaling):
  6x6.net, Qroute (fp64 tables), load 2.0, Philox injection, 8,192 replicas PER GPU.  One bench "step" is one EPISODE
  of the hot path over the whole batch: every replica restarts from an empty network and freshly initialised Q tables
  (Network.reset + Qroute.reset_tables, both on the device and inside the timed region) and runs `--sim-steps`
  simulator steps (inject -> send -> arrive -> learn) in ONE kernel launch.  Every step therefore does the same work —
  the congested learning phase of train.py's run (Q-routing converges after ~12k steps and then carries a third of the
  traffic) — whatever K and W are, and the timed region of the driver's K = 20 is about one second.
  `value` = packet-hops (sends) of all ranks in the timed region / max-over-ranks device time; `e2e` = the same through
  `Network.train` (host lambda in, the per-step mean-delivery-time curve of the load level back on the host).
After the headline the same ranks run BASELINE.json's configs 2-5 briefly (`extra_configs` in the same JSON line).
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
POLICIES = ["Shortest", "Qroute", "HybridQ", "MaHybridQ", "CQ", "CDRQ", "GlobalRoute"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--replicas-per-gpu", type=int, default=8192)
    ap.add_argument("--sim-steps", type=int, default=5000, help="simulator steps per bench step (one episode, one launch)")
    ap.add_argument("--net", default="6x6.net")
    ap.add_argument("--policy", default="Qroute", choices=POLICIES)
    ap.add_argument("--load", type=float, default=2.0)
    ap.add_argument("--queue-capacity", type=int, default=0, help="ring slots per (replica, node); 0 = from the episode length")
    ap.add_argument("--replicas-per-block", type=int, default=0)
    ap.add_argument("--threads-per-block", type=int, default=0)
    ap.add_argument("--tables-in-smem", type=int, default=-1)
    ap.add_argument("--page-records", type=int, default=0, help="> 0: paged queues with pages of this many packets instead of rings")
    ap.add_argument("--pages-per-replica", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip BASELINE.json's configs 2-5 after the headline")
    ap.add_argument("--only-extra", default="", help="run only these extra configs (comma-separated keys), no headline")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline samples")
    return ap.parse_args()


def episode_capacity(sim_steps, load, n_nodes):
    """Ring slots per (replica, node) for ONE episode from an empty network.  The reference's queues are unbounded and grow
    roughly linearly under congestion (Qroute 6x6 load 2.0 reaches ~500 per node in 5k steps, ~1,040 in 10k);
    twice the observed worst node, as a power of two.  An overflow is never silent: the run raises OverflowError."""
    need = 256 + 0.2 * sim_steps * max(load / 2.0, 1.0) * (36.0 / n_nodes) ** 0.5
    return 1 << max(int(math.ceil(math.log2(need))), 8)


def workload_name(net, policy, load, replicas, sim_steps):
    return (f"{net} {policy} fp64, load {load}, {replicas} replicas/GPU, Philox injection, one fresh episode of "
            f"{sim_steps} simulator steps per bench step")


LR = {"q": 0.1, "p": 0.001, "e": 0.1}          # MaHybridQ needs p=0.001; q is every policy's default


# ------------------------------------------------------------------------------------------------
# CPU arms.  (1) the UNMODIFIED reference staged in oracle/_ref (oracle/make_ref.py), one process per host core;
#            (2) the oracle port (oracle/rlr_oracle.c, OpenMP) — test infrastructure used as a second baseline
# ------------------------------------------------------------------------------------------------
def reference_sample(a, pool, episodes_per_proc=1):
    """One fresh episode per worker process through the reference's own Network.train (env.py:367-414)."""
    jobs = [(a.net, a.policy, a.load, a.sim_steps, 1000 + i, None if a.policy != "MaHybridQ" else LR)
            for i in range(pool.procs * episodes_per_proc)]
    res, wall = pool.run(jobs)
    return sum(r["hops_performed"] for r in res), wall, res


def port_sample(a, seconds):
    """The oracle port on all host cores on a bounded sample of the workload (fresh episodes, about `seconds` s)."""
    import numpy as np
    from oracle import oracle as O
    from rlrouting_b200 import nets
    from rlrouting_b200.topology import Topology
    O.build()
    cores = os.cpu_count() or 1
    topo = Topology.from_text(nets.net_text(a.net))
    n = 32 * cores
    sends, dt, rounds = 0, 0.0, 0
    while dt < seconds and rounds < 8:
        envs = []
        for r in range(n):
            e = O.OracleEnv(topo.links, a.policy, seed=1, replica=rounds * n + r)
            if a.policy not in ("Shortest", "GlobalRoute"):
                e.set_qtable(np.random.Generator(np.random.Philox(key=[1, rounds * n + r])).standard_normal(topo.tsize))
            envs.append(e)
        t0 = time.perf_counter()
        sends += O.run_many(envs, a.sim_steps, a.load, LR["q"], LR["p"], LR["e"], nthreads=cores)
        dt += time.perf_counter() - t0
        rounds += 1
    return {"value": sends / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{rounds} x {n} fresh episodes of {a.sim_steps} simulator steps of the same workload, "
                      f"oracle/rlr_oracle.c via OpenMP on {cores} threads, {dt:.1f} s"}


def cpu_baselines(a):
    """cpu_baseline (the real reference when it is staged, else the port) and cpu_baseline_port."""
    from oracle import ref_runner as RR
    port = port_sample(a, min(a.cpu_seconds, 6.0))
    if not RR.available():
        return port, port
    pool = RR.Pool()
    try:
        kat = RR.check_kat(pool)                     # SURVEY §8(c) row 1 through the staged reference: asserted
        hops, wall, res = reference_sample(a, pool)
    finally:
        pool.close()
    ref = {"value": hops / wall, "unit": UNIT, "cores": pool.procs, "kind": "reference",
           "one_core_value": kat["hops_performed"] / kat["seconds"],
           "kat": {k: kat[k] for k in ("all_packets", "end_packets", "route_time", "hops_performed")},
           "sample": f"unmodified reference Network.train (oracle/_ref): {len(res)} fresh episodes of {a.sim_steps} steps of "
                     f"the same workload, one process per host core ({pool.procs}), {wall:.1f} s wall; known answer "
                     f"(seed 1, T=10000: 19934/14930/31080406/344449) reproduced in {kat['seconds']:.1f} s on one core"}
    return ref, port


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores — the unmodified
    Python `Network.train` staged in oracle/_ref, one process per core, each bench step one fresh episode per process
    (a bounded sample of my arm's 8,192 per GPU).  Falls back to the oracle port only if the staging is missing."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import ref_runner as RR
    name = workload_name(a.net, a.policy, a.load, a.replicas_per_gpu, a.sim_steps)
    if RR.available():
        pool = RR.Pool()
        try:
            kat = RR.check_kat(pool)
            for _ in range(min(a.warmup, 1)):           # processes are warm after one episode; more only burns minutes
                reference_sample(a, pool)
            hops, dt = 0, 0.0
            for _ in range(a.steps):
                h, w, _r = reference_sample(a, pool)
                hops += h
                dt += w
        finally:
            pool.close()
        cores, kind = pool.procs, "reference"
        sample = (f"unmodified reference Network.train (oracle/_ref), {cores} processes x 1 fresh episode of {a.sim_steps} "
                  f"simulator steps per bench step; known answer of SURVEY 8(c) reproduced first "
                  f"({kat['hops_performed']} hops in {kat['seconds']:.1f} s on one core)")
    else:
        from oracle import oracle as O
        O.build()
        p = port_sample(a, 1e9 if a.steps > 8 else a.steps * 2.0)
        hops, dt, cores, kind, sample = p["value"], 1.0, p["cores"], "port", p["sample"]
    v = hops / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "sim_steps_per_bench_step": a.sim_steps, "timing": "host wall clock around the CPU loop"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
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
                                          "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        """Summary of the samples taken in [t0, t1] (host clock); the sampler keeps running."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        rows = [r for t, r in list(self.rows) if t0 <= t <= t1]
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

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(policy, topo, hops, injections, replica_steps, esize=8):
    """per node-step head+tail 16 B; per injection a 16 B record; per hop a 16 B record read,
    the policy's rows, its writes and a 16 B record write (counted for every hop).  esize: bytes per table entry."""
    w, dbar = esize, 4 if topo.max_deg <= 4 else topo.max_deg      # padded-row convention of SURVEY §8(d)
    per_hop = {"Shortest": 16 + 2 + 16, "GlobalRoute": 16 + 1 + 16,
               "Qroute": 16 + w * dbar + w * dbar + w + 16,
               "HybridQ": 16 + w * dbar + w * dbar + w + 16 + 2 * w * dbar,
               "MaHybridQ": 16 + w * dbar + w * dbar + w + 16 + w * dbar,
               "CQ": 16 + w * dbar + w * dbar + w + 16 + 4 * w,                       # + C_f read, C[x][d][k] read+write, Q read
               "CDRQ": 16 + 2 * w * dbar + w * dbar + 2 * w + 16 + 8 * w}[policy]    # + backward row, cell reads/writes
    dense = {"GlobalRoute": 5 * topo.N * topo.N,      # distance (int32) + next-hop (u8) tables rewritten every step
             "MaHybridQ": 4 * w * topo.tsize, "CQ": 2 * w * topo.tsize, "CDRQ": 2 * w * topo.tsize}.get(policy, 0)
    return hops * per_hop + replica_steps * (16 * topo.N + dense) + injections * 16, per_hop


# ------------------------------------------------------------------------------------------------
# one workload on this rank's GPU: device-resident episodes (`value`) and Network.train episodes (`e2e`)
# ------------------------------------------------------------------------------------------------
class Workload:
    def __init__(self, key, net, policy, replicas, sim_steps, loads, note="", launch=(0, 0, -1), queue_capacity=0, agent_kw=None,
                 dtype="float64", page_records=0, pages_per_replica=0):
        self.key, self.net, self.policy, self.replicas, self.sim_steps = key, net, policy, int(replicas), int(sim_steps)
        self.loads, self.note, self.launch, self.queue_capacity = list(loads), note, launch, queue_capacity
        self.agent_kw = agent_kw or {}
        self.dtype = dtype
        self.page_records, self.pages_per_replica = int(page_records), int(pages_per_replica)


def measure(w, steps, warmup, rank, world, dev, sampler=None, vectors_steps=0):
    """Times `steps` episodes of workload `w` twice — device-resident and through Network.train with per-level curves
    coming back to the host — and returns the record (rank 0 keeps it; every rank takes part in the reductions)."""
    import numpy as np
    import torch
    import rlrouting_b200 as rb
    from rlrouting_b200 import dist as rdist
    R, T, L = w.replicas, w.sim_steps, len(w.loads)
    per_level = R // L
    lv = np.minimum(np.arange(R) // max(per_level, 1), L - 1).astype(np.int32)
    lam = np.asarray(w.loads, dtype=np.float64)[lv]
    n_nodes = {"6x6.net": 36, "6x6-modified.net": 36, "lata.net": 116}.get(w.net, 64)
    cap = w.queue_capacity or episode_capacity(T, max(w.loads), n_nodes)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    nw = rb.Network(w.net, replicas=R, rng="philox", seed=1, replica_base=rank * R, queue_capacity=cap,
                    replicas_per_block=w.launch[0], threads_per_block=w.launch[1], tables_in_smem=w.launch[2], dtype=w.dtype,
                    page_records=w.page_records, pages_per_replica=w.pages_per_replica or None)
    agent = getattr(rb, w.policy)(nw, **w.agent_kw)
    nw.agent = agent
    learns = hasattr(agent, "keep_initial_tables")
    if learns:
        agent.keep_initial_tables()
    info = nw.launch_info()

    def fresh():
        if learns:
            agent.reset_tables()
        nw.reset()

    # ---------------- device-resident throughput (`value`) ----------------
    for _ in range(warmup):
        fresh()
        nw.run_device(T, lam, lr=LR, next_clock0=0)
    barrier()
    launches0 = nw.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = torch.zeros(8, dtype=torch.int64, device=dev)
    barrier()
    tw0 = time.perf_counter()
    ev0.record()
    marks, kmarks = [], []
    for _ in range(steps):
        fresh()                                     # Network.reset + Qroute.reset_tables: memsets + two small kernels
        ka = torch.cuda.Event(enable_timing=True)
        ka.record()                                 # the step kernel's own launch, bracketed on the stream it runs on
        nw.run_device(T, lam, lr=LR, next_clock0=0)  # one rlr_run_steps launch, per-step records written on the device
        kb = torch.cuda.Event(enable_timing=True)
        kb.record()
        kmarks.append((ka, kb))
        tot += nw.stats_device()                    # this episode's counters (rlr_reduce_stats), summed on the device
        marks.append(torch.cuda.Event(enable_timing=True))
        marks[-1].record()
    ev1.record()
    barrier()
    tw1 = time.perf_counter()
    step_ms = [round(a_.elapsed_time(b_), 3) for a_, b_ in zip([ev0] + marks[:-1], marks)]
    kms = torch.tensor([sum(a_.elapsed_time(b_) for a_, b_ in kmarks) / steps], dtype=torch.float64, device=dev)
    rdist.allreduce_max_(kms)
    kernel_ms = float(kms.item())                   # average duration of one step-kernel launch (max over ranks)
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    rdist.allreduce_max_(ms)
    clocks = sampler.window(tw0, tw1) if sampler else None
    rdist.allreduce_sum_(tot)
    totals = tot.cpu().numpy()
    hops, inj = int(totals[5]), int(totals[0])
    launches = nw.launch_count - launches0 + (2 if learns else 0) * steps       # + the table re-init kernel and the copy
    ms_total = float(ms.item())
    value = hops / (ms_total * 1e-3)
    nw.check_status()
    # per-level delivery time and Q-convergence of the LAST episode: the one collective of the path
    q_prev = agent._q_initial if learns else None
    final = rdist.global_stats(nw, level_index=lv, n_levels=L, q_prev=q_prev)

    # ---------------- end to end through Network.train (`e2e`) ----------------
    out = None
    for _ in range(min(warmup, 2)):
        fresh()
        out = nw.train(T, lam, lr=LR, level_index=lv, next_clock0=0)
    barrier()
    tot.zero_()
    ev0.record()
    d2h = h2d = 0
    for _ in range(steps):
        fresh()
        out = nw.train(T, lam, lr=LR, level_index=lv, next_clock0=0)     # host lambda -> device; [L, T] curves -> host
        d2h += nw.last_d2h_bytes
        h2d += nw.last_h2d_bytes
        tot += nw.stats_device()
    ev1.record()
    barrier()
    ms2 = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    rdist.allreduce_max_(ms2)
    rdist.allreduce_sum_(tot)
    e2e_value = int(tot.cpu().numpy()[5]) / (float(ms2.item()) * 1e-3)
    curve_end = [float(v) for v in out["route_time"][:, -1]]
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // steps, "d2h_bytes_per_step": d2h // steps,
           "note": "Network.reset + reset_tables + Network.train(sim_steps, loads, level_index=...): lambda + threshold vectors "
                   "H2D from pinned memory; the per-step mean-delivery-time curve of every load level ([levels, sim_steps] "
                   "float64, pooled on the device by rlr_reduce_levels) and the status words D2H"}
    vec = None
    if vectors_steps > 0:                            # the reference's own return value for EVERY replica: [R, sim_steps] fp64
        fresh()
        nw.train(T, lam, lr=LR, next_clock0=0)
        barrier()
        tot.zero_()
        ev0.record()
        d2h = 0
        for _ in range(vectors_steps):
            fresh()
            nw.train(T, lam, lr=LR, next_clock0=0)
            d2h += nw.last_d2h_bytes
            tot += nw.stats_device()
        ev1.record()
        barrier()
        ms3 = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        rdist.allreduce_max_(ms3)
        rdist.allreduce_sum_(tot)
        vec = {"value": int(tot.cpu().numpy()[5]) / (float(ms3.item()) * 1e-3), "unit": UNIT, "steps": vectors_steps,
               "d2h_bytes_per_step": d2h // vectors_steps,
               "note": "same, with the per-replica [R, sim_steps] route_time vectors copied to the host instead of the level curves"}
    topo = nw.topology
    ring_gb = nw.queue_memory_bytes / 1e9
    live_peak = int(final["active_packets"]) / max(world * R, 1)        # packets alive per replica at the end of the last episode
    table_mb = R * topo.tsize * (4 if w.dtype == "float32" else 8) * len(getattr(agent, "_table_names", ())) / 1e6
    del nw, agent, out
    import gc
    gc.collect()                                    # Network <-> Node reference cycles: free the rings now, not inside
    torch.cuda.empty_cache()                        # the next workload's timed region
    if rank != 0:
        return None
    peak, peak_src = measured_peak()
    esize = 4 if w.dtype == "float32" else 8
    abytes, per_hop = algorithmic_bytes(w.policy, topo, hops, inj, world * R * T * steps, esize)
    per_launch = abytes / max(steps * world, 1)                       # one step-kernel launch per GPU per bench step
    achieved = per_launch / (kernel_ms * 1e-3) / 1e9                 # GB/s per GPU over the kernel's own launch duration
    achieved_step = per_launch / (ms_total / steps * 1e-3) / 1e9     # ... over the whole bench step (reset, re-init, statistics)
    traffic = None
    try:                                            # DRAM bytes per launch from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(f"{w.net}|{w.policy}|{R}|{T}|{'smem' if info['tables_in_smem'] else 'hbm'}")
        if t:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        pass
    lev = final.get("levels", {})
    return {
        "key": w.key, "table_dtype": w.dtype, "workload": workload_name(w.net, w.policy, w.loads[0] if L == 1 else f"{w.loads[0]}..{w.loads[-1]} ({L} levels)", R, T),
        "note": w.note, "value": value, "unit": UNIT, "n_gpus": world, "replicas_total": world * R, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_total / steps, "step_ms": step_ms, "hops_in_timed_region": hops, "gpu_launches": launches,
        "launch": info, "queue_capacity": cap,
        "queues": ({"kind": "paged", "page_records": w.page_records, "pages_per_replica": w.pages_per_replica} if w.page_records
                   else {"kind": "rings", "slots_per_queue": cap}),
        "queue_memory_bytes_per_replica": int(ring_gb * 1e9 / R), "live_packets_per_replica_at_end": live_peak,
        "l2_policy": "state larger than L2: tables %.0f MB + queues %.1f GB per GPU, every episode starts cold" % (table_mb, ring_gb),
        "e2e": e2e, "e2e_per_replica_vectors": vec, "clocks": clocks,
        "loads": [float(x) for x in w.loads],
        "mean_delivery_time_per_level": [float(x) for x in lev.get("ave_route_time", [])],
        "mean_delivery_time_per_level_e2e_curve_end": curve_end,
        "mean_delivery_time": final["ave_route_time"],
        "q_convergence": ({"sum_abs_delta_q_per_level": [float(x) for x in lev["q_abs_delta"]],
                           "sum_q_per_level": [float(x) for x in lev["q_sum"]],
                           "note": "sum over the level's replicas of |Q - Q_initial| and of Q after the last episode "
                                   "(rlr_table_stats per rank, fp64 all-reduce)"} if "q_abs_delta" in lev else None),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "algorithmic_bytes_per_launch": per_launch, "peak_source": peak_src,
                     "kernel": "rlr_step_kernel", "kernel_ms_per_launch": kernel_ms, "algorithmic_bytes_per_hop": per_hop,
                     "frac_of_whole_step": achieved_step / peak,
                     "timing": "CUDA events around the rlr_run_steps launch on its stream, averaged over the timed steps, max "
                               "over ranks; the next episode's injection-schedule kernel runs beside it on a side stream",
                     "note": "algorithmic bytes per SURVEY 8(d); where the tables are staged in shared memory DRAM traffic is "
                             "far below the algorithmic figure (profiles/)"},
    }


def extra_workloads(world):
    """BASELINE.json configs 2-5 as short secondary workloads on the same ranks (weak scaling: per-GPU shares)."""
    import numpy as np
    sweep = [float(x) for x in np.linspace(1.0, 4.0, 16)]
    return [
        Workload("headline_fp32_storage", "6x6.net", "Qroute", 8192, 5000, [2.0],
                 "NOT the headline (which stays fp64): the headline workload with the Q table STORED in fp32 (Network(dtype='float32')): "
                 "14 instead of 7 replicas per SM; bit-exact against the oracle's float-rounded-store law, statistically equal to fp64",
                 dtype="float32"),
        Workload("headline_paged_queues", "6x6.net", "Qroute", 8192, 5000, [2.0],
                 "NOT the headline: the headline workload with PAGED queues (64-record pages, 320 pages per replica) instead of "
                 "rings of 2048 slots per node: 3.6x less queue memory, bit-identical results, slower (page bookkeeping in the step)",
                 page_records=64, pages_per_replica=320),
        Workload("config2_shortest_sweep", "6x6.net", "Shortest", 4096, 2000, sweep,
                 "BASELINE config 2: shortest-path routing, 4096 replicas, 16 load levels 1.0-4.0 (256 replicas each) per GPU"),
        Workload("config3_6x6mod_qroute", "6x6-modified.net", "Qroute", 16384, 2000, [2.0],
                 "BASELINE config 3: 6x6-modified.net Q-routing, 16384 replicas per GPU (bit-exactness is tests/test_gpu_configs.py)"),
        Workload("config4_lata_qroute", "lata.net", "Qroute", 4096, 400, [2.0],
                 "BASELINE config 4a: lata.net Q-routing, 4096 replicas per GPU = 32768 on 8 GPUs", queue_capacity=1024),
        Workload("config4_lata_hybridq", "lata.net", "HybridQ", 4096, 200, [2.0],
                 "BASELINE config 4b: lata.net hybrid.py policy, 4096 replicas per GPU = 32768 on 8 GPUs", queue_capacity=1024),
        Workload("config5_mahybridq_levels", "6x6.net", "MaHybridQ", 131072, 300, [float(x) for x in np.linspace(0.5, 3.5, 16)],
                 "BASELINE config 5: multi_agent.py on 6x6.net, 16 load levels x 8192 replicas per GPU = 65536 replicas x 16 "
                 "levels on 8 GPUs (the same per-GPU share at 2 and 4 GPUs); paged queues (16-record pages, 256 per replica)",
                 queue_capacity=256, page_records=16, pages_per_replica=256),
    ]


def run_b200(a):
    import torch
    from rlrouting_b200 import dist as rdist
    rank, world, local = rdist.init_from_env()
    if world != a.gpus and world > 1:
        a.gpus = world
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    sampler = ClockSampler(local) if rank == 0 else None          # started before the warm-up; windows are cut per region
    extras = extra_workloads(world)
    if a.only_extra:
        keys = set(a.only_extra.split(","))
        recs = [measure(w, max(a.steps, 1), max(a.warmup, 1), rank, world, dev, sampler) for w in extras if w.key in keys]
        if rank == 0:
            print(json.dumps({"extra_configs": recs}))
        if sampler:
            sampler.stop()
        return
    head = Workload("headline", a.net, a.policy, a.replicas_per_gpu, a.sim_steps, [a.load],
                    launch=(a.replicas_per_block, a.threads_per_block, a.tables_in_smem), queue_capacity=a.queue_capacity,
                    page_records=a.page_records, pages_per_replica=a.pages_per_replica)
    rec = measure(head, a.steps, a.warmup, rank, world, dev, sampler, vectors_steps=min(3, a.steps))
    extra_recs = []
    if not a.no_extra:
        for w in extras:
            try:
                extra_recs.append(measure(w, 3, 2, rank, world, dev, sampler))
            except Exception as e:                  # an extra config never takes the headline line down with it
                if world > 1:
                    raise
                extra_recs.append({"key": w.key, "error": f"{type(e).__name__}: {e}"})
    if sampler:
        sampler.stop()
    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a.net, a.policy, a.load, a.replicas_per_gpu, a.sim_steps),
                   "replicas_total": rec["replicas_total"], "sim_steps_per_bench_step": a.sim_steps,
                   "queue_capacity": rec["queue_capacity"], "launch": rec["launch"], "l2_policy": rec["l2_policy"],
                   "hops_in_timed_region": rec["hops_in_timed_region"], "mean_delivery_time": rec["mean_delivery_time"],
                   "parallelism": f"replica-parallel x{world}, no data-path collective; one all-reduce of the per-level statistics"},
        "roofline": rec["roofline"], "e2e": rec["e2e"], "e2e_per_replica_vectors": rec["e2e_per_replica_vectors"],
        "gpu_launches": rec["gpu_launches"], "clocks": rec["clocks"], "step_ms": rec["step_ms"],
        "q_convergence": rec["q_convergence"],
        "extra_configs": [r for r in extra_recs if r is not None],
    }
    if not a.no_cpu_baseline and world == 1:
        ref, port = cpu_baselines(a)
        line["cpu_baseline"], line["cpu_baseline_port"] = ref, port
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
    if a.impl == "reference":
        run_reference(a)
        return
    run_b200(a)
    _finalize()


if __name__ == "__main__":
    main()
