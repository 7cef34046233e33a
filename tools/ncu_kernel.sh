#!/bin/bash
# usage: tools/ncu_kernel.sh <kernel-regex> <out-name> [launch-count]   (run under gpurun)
mkdir -p gpurun_out
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$1 -c ${3:-1} \
    -o gpurun_out/$2 -f python tools/profile_step.py --steps 1 > gpurun_out/$2.log 2>&1
tail -1 gpurun_out/$2.log
