#!/usr/bin/env python
"""Stages the UNMODIFIED reference for the timed CPU baseline — TEST / MEASUREMENT INFRASTRUCTURE ONLY.

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference exists)

Copies the reference's six simulator modules and three topologies, byte for byte, from /root/reference into the
git-ignored directory oracle/_ref/ (never into history: .gitignore lists oracle/_ref/; .gpurunignore does not, so
the staged copy travels to the GPU box like the built .so files).  `bench.py --impl reference` and the
`cpu_baseline` leg of bench.py run `Network.train` (reference env.py:367-414) from there, unmodified, through
oracle/ref_runner.py.  Nothing under rlrouting_b200/ imports it.  A MANIFEST with the sha256 of every staged file
is written beside them so a run can state what it timed.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("RLR_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("env.py", "base_policy.py", "qroute.py", "shortest.py", "hybrid.py", "multi_agent.py",
         "6x6.net", "6x6-modified.net", "lata.net")


def stage(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"make_ref: {SRC} is not here; keeping what oracle/_ref/ already holds", file=sys.stderr)
        return os.path.isfile(os.path.join(DST, "MANIFEST.json"))
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for name in FILES:
        shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
        with open(os.path.join(DST, name), "rb") as f:
            manifest[name] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "xuxife/rlrouting (unmodified copies)", "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"make_ref: staged {len(FILES)} files in {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
