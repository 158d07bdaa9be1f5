#!/usr/bin/env python
"""Warp-stall table per phase of the fused step kernel, from one `ncu --set full --import-source on` capture.

  ncu -i prof.ncu-rep --page source --csv > src.csv
  python tools/phase_stalls.py src.csv rlrouting_b200/csrc/librlrouting_b200.so 'rlr_step_kernelILi1ELi4ELi3ELb1ELb0ELi0' > profiles/..md

The i-th SASS instruction of the kernel in ncu's source page is matched with the i-th instruction of the same function in
`nvdisasm -g` (line info from -lineinfo).  Instructions inside the step loop (between the loop's first instruction and its
back-branch) are split at the two BAR.SYNC instructions into the send segment, the arrival segment and the step's tail;
within a segment they are grouped by what their source line does (markers in rlr_step.cuh).  Inlined helpers (cp.async,
predicated loads / stores) carry the helper's own line, so they are counted in the segment they execute in.  Per group:
share of executed warp-instructions, share of the stall samples (smsp__pcsamp), and the top stall reasons of the group.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "rlrouting_b200", "csrc", "rlr_step.cuh")


def sass_lines(so, func_pat):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    out = []
    for cub in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
        cur, active = None, False
        for l in txt.splitlines():
            if l.startswith("//----") and ".text." in l:
                active = re.search(func_pat, l) is not None
                continue
            if not active:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = int(m.group(2))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                out.append((int(m.group(1), 16), cur, m.group(2).strip()))
        if out:
            break
    return out


def marker_lines():
    src = open(SRC).read().split("\n")

    def find(pat, start=0):
        for i in range(start, len(src)):
            if pat in src[i]:
                return i + 1
        raise SystemExit(f"marker {pat!r} not found in rlr_step.cuh")
    m = {}
    m["kernel"] = find("rlr_step_kernel(const __grid_constant__ StepParams p)")
    m["append"] = find("auto append = [&]")
    m["inject_fn"] = find("auto inject_step = [&]")
    m["loop"] = find("for (int s = 0; s < p.K; ++s) {")
    m["B"] = find("// ---- B. Node._send_default", m["loop"])
    m["choose"] = find("if (have) {", m["B"])
    m["get_info"] = find("y = s_col[rp + k];", m["choose"])
    m["publish"] = find("RLR_T(0);", m["get_info"])
    m["bar1"] = find("---- barrier 1 ----", m["publish"])
    m["C"] = find("// ---- C. event drain", m["bar1"])
    m["learn_in_C"] = find("if (kLearnInArrivals) {", m["C"])
    m["flat"] = find("if (kFlatArrivals) {", m["learn_in_C"])
    m["flat_deliver"] = find("unsigned ndel = 0u", m["flat"])
    m["flat_append"] = find("const unsigned tail0 = tail;", m["flat_deliver"])
    m["D"] = find("// ---- D. agent.learn", m["flat_append"])
    m["commit"] = find("if (FUSED) cp_async_commit();", m["D"])
    m["bar2"] = find("---- barrier 2 ----", m["commit"])
    m["end"] = find("// ---- write back ----", m["bar2"])
    return m


def classify(line, seg, M):
    """seg: 0 = before barrier 1, 1 = between the barriers, 2 = after barrier 2 (inside the loop)."""
    if line is None:
        return ("?", "no line info")
    if line < M["kernel"]:
        helper = "helpers: cp.async / predicated ld+st / mbarrier"
        return (["send", "arrive", "tail"][seg], helper)
    if M["append"] <= line < M["inject_fn"]:
        return (["send", "arrive", "tail"][seg], "ring append (lambda)")
    if M["inject_fn"] <= line < M["loop"]:
        return (["send", "arrive", "tail"][seg], "next step's injections (inject_step)")
    if line < M["loop"]:
        return (["send", "arrive", "tail"][seg], "heap_pos lookups (heap_pos_at lambda, __ldg through L1)")
    if line < M["B"]:
        return ("send", "loop head, cp.async wait")
    if line < M["choose"]:
        return ("send", "pop: staged head record, request the record D-1 behind")
    if line < M["get_info"]:
        return ("send", "choose: row Q[x][d] scan, first-index argmax")
    if line < M["publish"]:
        return ("send", "get_info: next hop, row Q[y][d] max, publish record")
    if line < M["C"]:
        return (["send", "arrive", "tail"][min(seg, 1)], "ballot, send mask, target|rank, barrier 1")
    if line < M["learn_in_C"]:
        return ("arrive", "window popcounts (heap_pos row, sender ranks)")
    if line < M["flat"]:
        return ("arrive", "learn: Q[x][d][k] update (fp64, unfused)")
    if line < M["flat_deliver"]:
        return ("arrive", "gather: 4 neighbour slots, heap positions, records")
    if line < M["flat_append"]:
        return ("arrive", "deliveries: end_packets / route_time / hops")
    if line < M["D"]:
        return ("arrive", "order + append arrivals (ranks, ring + stage writes)")
    if line < M["commit"]:
        return ("arrive", "learn dispatch (other policies)")
    if line < M["bar2"]:
        return (["send", "arrive", "tail"][seg], "commit group, inject call, overflow check, barrier 2")
    if line < M["end"]:
        return ("tail", "metric thread: per-step record")
    return ("epilogue", "write back")


def main():
    src_csv, so, pat = sys.argv[1:4]
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) >= len(hdr)]
    sass = sass_lines(so, pat)
    if len(sass) != len(body):
        print(f"warning: {len(sass)} SASS instructions in the cubin vs {len(body)} in the ncu listing", file=sys.stderr)
    M = marker_lines()
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    # the step loop = the most-executed contiguous stretch: find the back-branch with the highest execution count
    execs = [int(r[idx["Instructions Executed"]] or 0) for r in body]
    peak = max(execs)
    in_loop = [e >= peak * 0.25 for e in execs]
    first = in_loop.index(True)
    last = len(in_loop) - 1 - in_loop[::-1].index(True)
    groups = collections.OrderedDict()
    seg = 0
    tot_inst = tot_samp = 0
    for i, ((off, line, text), r) in enumerate(zip(sass, body)):
        n, s = execs[i], int(r[idx["# Samples"]] or 0)
        if i < first or i > last:
            key = ("outside the loop", "staging, set-up, write-back")
        else:
            key = classify(line, seg, M)
            if text.startswith("BAR.SYNC"):
                key = (["send", "arrive", "tail"][seg], f"BAR.SYNC #{seg + 1} (wait for the CTA's other warp)")
                seg = min(seg + 1, 2)
        g = groups.setdefault(key, {"inst": 0, "samp": 0, "stall": collections.Counter(), "n": 0})
        g["inst"] += n
        g["samp"] += s
        g["n"] += 1
        for st in stalls:
            g["stall"][st[6:]] += int(r[idx[st]] or 0)
        tot_inst += n
        tot_samp += s
    print(f"kernel: {rows[0][1]}")
    print(f"warp-instructions executed: {tot_inst:,}; stall samples: {tot_samp:,}; step loop = SASS #{first}..#{last} of {len(body)}\n")
    print("| segment | what | SASS | warp-inst % | samples % | top stall reasons (share of the group's samples) |")
    print("|---|---|---|---|---|---|")
    order = {"send": 0, "arrive": 1, "tail": 2}
    for (sg, what), g in sorted(groups.items(), key=lambda kv: (order.get(kv[0][0], 3), -kv[1]["samp"])):
        top = ", ".join(f"{k} {100 * v / max(g['samp'], 1):.0f}%" for k, v in g["stall"].most_common(3) if v)
        print(f"| {sg} | {what} | {g['n']} | {100 * g['inst'] / tot_inst:.1f} | {100 * g['samp'] / max(tot_samp, 1):.1f} | {top} |")
    allst = collections.Counter()
    for g in groups.values():
        allst.update(g["stall"])
    T = sum(allst.values())
    print("\nwhole kernel: " + ", ".join(f"{k} {100 * v / T:.1f}%" for k, v in allst.most_common(10)))
    for sg in ("send", "arrive", "tail", "outside the loop"):
        si = sum(g["inst"] for (a, _), g in groups.items() if a == sg)
        ss = sum(g["samp"] for (a, _), g in groups.items() if a == sg)
        print(f"segment {sg}: {100 * si / tot_inst:.1f}% of the instructions, {100 * ss / max(tot_samp, 1):.1f}% of the samples")


if __name__ == "__main__":
    main()
