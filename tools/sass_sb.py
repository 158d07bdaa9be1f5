#!/usr/bin/env python
"""Print a kernel's SASS with the scheduling control fields decoded (stall count, yield, the scoreboard an
instruction sets on write / read completion, and the scoreboards it waits for).

  python tools/sass_sb.py lib.so 'rlr_step_kernelILi1ELi4ELi3ELb1ELb0' [lo_hex hi_hex]

Used to find false dependencies between loads that ptxas put on the same scoreboard.
"""
import re
import subprocess
import sys


def main():
    so, pat = sys.argv[1:3]
    lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.split("\n")
    active = False
    p1 = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/")
    p2 = re.compile(r"^\s+/\* (0x[0-9a-f]{16}) \*/")
    i = 0
    while i < len(txt):
        l = txt[i]
        if "Function :" in l:
            active = re.search(pat, l) is not None
        m = p1.match(l) if active else None
        if m and i + 1 < len(txt) and p2.match(txt[i + 1]):
            hiw = int(p2.match(txt[i + 1]).group(1), 16)
            stall, yld, wb, rb, wm = (hiw >> 41) & 0xF, (hiw >> 45) & 1, (hiw >> 46) & 7, (hiw >> 49) & 7, (hiw >> 52) & 0x3F
            a = int(m.group(1), 16)
            if lo <= a <= hi:
                print(f"{m.group(1)} st={stall:2d} y={yld} W={'-' if wb == 7 else wb} R={'-' if rb == 7 else rb} "
                      f"wait={wm:06b}  {m.group(2).strip()}")
            i += 2
            continue
        i += 1


if __name__ == "__main__":
    main()
