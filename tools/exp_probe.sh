#!/bin/bash
# phase probes of -DRLR_EXP_TIMING builds:  gpurun -- 'bash tools/exp_probe.sh build/exp/t0.so ...'
for lib in "$@"; do RLR_LIB=$PWD/$lib timeout 300 python tools/phase_probe.py 2>&1 | tail -1; done
