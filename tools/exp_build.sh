#!/bin/bash
# Tuning build of one policy's degree-4 instantiations (seconds):  tools/exp_build.sh <name> [-DVARIANT ...]
#   -> build/exp/<name>.so + build/exp/<name>.sass (the fused smem-table kernel) and the size of its step loop
set -e
name=$1; shift
mkdir -p build/exp
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -DRLR_FAST_BUILD "$@" \
     -shared -o build/exp/$name.so rlrouting_b200/csrc/rlr_kernels.cu 2>&1 | grep -v "warning #177\|det_log2\|\^\|^$" || true
cuobjdump -sass build/exp/$name.so 2>/dev/null | awk '/Function :/{f=($0 ~ /rlr_step_kernelILi[0-9]+ELi4ELi3ELb1ELb0ELi0EdLb0E/)} f' \
    | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\/$//' > build/exp/$name.sass
python3 - "$name" <<'PY'
import re, sys
name = sys.argv[1]
lines = open(f"build/exp/{name}.sass").read().splitlines()
addr = lambda l: int(re.search(r"/\*([0-9a-f]{4})\*/", l).group(1), 16)
# the step loop = the shortest backward branch whose span holds two CTA barriers
bars = [addr(l) for l in lines if "BAR.SYNC" in l]
best = (1 << 30, 0, 0)
for l in lines:
    m = re.search(r"BRA(?:\.U)? (0x[0-9a-f]+)", l)
    if m:
        a, tgt = addr(l), int(m.group(1), 16)
        if tgt < a and sum(tgt <= b <= a for b in bars) >= 2 and a - tgt < best[0]:
            best = (a - tgt, tgt, a)
n = best[0] // 16 + 1
body = [l for l in lines if best[1] <= addr(l) <= best[2]]
from collections import Counter
ops = Counter(re.sub(r"^\s*/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T]+\s+)?", "", l).split()[0].split(".")[0] for l in body)
print(f"{name}: kernel {len(lines)} SASS, step loop {n} SASS [{best[1]:#x}..{best[2]:#x}]")
print("   ", ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
PY
cuobjdump -res-usage build/exp/$name.so 2>/dev/null | grep -A1 "ELi4ELi3ELb1ELb0ELi0EdLb0E" | grep -o "REG:[0-9]*" | head -1
