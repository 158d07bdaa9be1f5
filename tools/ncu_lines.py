#!/usr/bin/env python
"""Attribute an ncu `--page source --csv` SASS listing to CUDA source lines.

  ncu -i prof.ncu-rep --page source --csv > src.csv
  python tools/ncu_lines.py src.csv rlrouting_b200/csrc/librlrouting_b200.so 'rlr_step_kernelILi1ELi4ELb1E' [top]

The i-th SASS instruction of the kernel in the ncu listing is matched with the i-th instruction of
the same function in `nvdisasm -g` output (which carries `//## File ..., line N` markers), and
executed-instruction counts / stall samples are summed per source line.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so, func_pat):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    out = []
    for cub in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):      # one cubin per translation unit
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


def main():
    src_csv, so, pat = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) >= len(hdr)]
    sass = sass_lines(so, pat)
    if len(sass) != len(body):
        print(f"warning: {len(sass)} SASS instructions in cubin vs {len(body)} in ncu listing", file=sys.stderr)
    inst, samp = collections.Counter(), collections.Counter()
    for (off, line, text), r in zip(sass, body):
        inst[line] += int(r[idx["Instructions Executed"]] or 0)
        samp[line] += int(r[idx["# Samples"]] or 0)
    ti, ts = sum(inst.values()), sum(samp.values())
    srcpath = os.path.join(os.path.dirname(os.path.abspath(so)), "rlr_step.cuh")
    if not os.path.exists(srcpath):
        srcpath = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rlrouting_b200", "csrc", "rlr_step.cuh")
    srcfile = [l.rstrip("\n") for l in open(srcpath)]
    print(f"total warp-instructions {ti}, samples {ts}")
    print(f"{'line':>5} {'inst%':>6} {'samp%':>6}  source")
    for line, n in inst.most_common(top):
        text = srcfile[line - 1].strip()[:110] if line and line <= len(srcfile) else "?"
        print(f"{line!s:>5} {100 * n / ti:6.2f} {100 * samp[line] / max(ts, 1):6.2f}  {text}")


if __name__ == "__main__":
    main()
