#!/bin/bash
# Kernel-tuning loop: for every experiment library given (built here with -DRLR_FAST_BUILD -D<variant>), check the
# 6x6 Qroute kernel against the oracle, then run the headline bench.  Results: gpurun_out/exp_<name>.{check,bench}
#   gpurun -- 'bash tools/exp_run.sh build/exp/v1.so build/exp/v2.so'
mkdir -p gpurun_out
make -C oracle -s
for lib in "$@"; do
    name=$(basename "$lib" .so)
    RLR_LIB=$PWD/$lib timeout 300 python tools/exp_check.py > gpurun_out/exp_$name.check 2>&1
    echo "check $name rc=$?: $(tail -1 gpurun_out/exp_$name.check)"
    RLR_LIB=$PWD/$lib timeout 300 python bench.py --steps ${EXP_STEPS:-10} --warmup 3 --no-cpu-baseline --no-extra ${EXP_ARGS} > gpurun_out/exp_$name.bench 2> gpurun_out/exp_$name.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/exp_$name.bench").read().strip().splitlines()[-1])
    print("bench $name: ms/step %.3f value %.4g e2e %.4g frac %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"]))
except Exception as e:
    print("bench $name failed", e)
PY
done
